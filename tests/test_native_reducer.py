"""The NATIVE DDP reducer (ops/csrc/tfy_reducer.cpp) on a CPU box: the production C++ is compiled against a test
double of the CUDA runtime / collective entry points (tests/native/reducer_stub.cpp) and driven through ctypes.
Checked: buckets are launched strictly in index order whatever order the gradient hooks fire in (every rank must
issue the same kernel sequence), unready buckets are flushed by finalize, every launch is fenced ready -> comm stream
-> done, the compute stream waits for the LAST collective only, counters re-arm, errors propagate, fused mode calls
the fused step (reference being replaced: c10d::Reducer under tf_yarn/pytorch/tasks/worker.py:105-107)."""
import ctypes
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tf_yarn_b200", "ops", "csrc")
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    cxx = shutil.which("g++")
    if not cxx or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    out = str(tmp_path_factory.mktemp("reducer") / "libreducer_test.so")
    # -Bsymbolic: the library's own cuda* test doubles win over a real libcudart that torch may have loaded globally
    cmd = [cxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-Wl,-Bsymbolic", "-include",
           os.path.join(ROOT, "tests", "native", "cuda_device_shim.h"), "-I", CUDA_INC, "-I", CSRC,
           os.path.join(CSRC, "tfy_reducer.cpp"), os.path.join(ROOT, "tests", "native", "reducer_stub.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = ctypes.CDLL(out)
    lib.tfy_reducer_create.restype = ctypes.c_void_p
    lib.tfy_reducer_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    for name in ("tfy_reducer_mark_ready", "tfy_reducer_finalize", "tfy_reducer_set_fused"):
        getattr(lib, name).restype = ctypes.c_int
    lib.tfy_reducer_mark_ready.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    lib.tfy_reducer_finalize.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.tfy_reducer_set_fused.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p]
    lib.tfy_reducer_launches.restype = ctypes.c_long
    lib.tfy_reducer_launches.argtypes = [ctypes.c_void_p]
    lib.tfy_reducer_next.argtypes = [ctypes.c_void_p]
    lib.tfy_reducer_destroy.argtypes = [ctypes.c_void_p]
    lib.tfy_reducer_destroy.restype = None
    lib.stub_log.argtypes = [ctypes.c_char_p, ctypes.c_int]
    return lib


CUR = 0xA0                      # the "compute stream" handle the hooks pass in


def _make(lib, world, n_params=(2, 1, 3), offs=(4096, 8192, 16384), ns=(256, 512, 1024)):
    ctx = ctypes.create_string_buffer(4096)             # TfyCommCtx: rank at +?, world ... zero-filled = rank 0
    # world is read by the reducer (ctx.world > 1): write it through the mirror used by the package
    from tf_yarn_b200.ops import native
    real = native.CommCtx()
    real.rank, real.world = 0, world
    ctypes.memmove(ctx, ctypes.byref(real), ctypes.sizeof(real))
    nb = len(n_params)
    c_offs = (ctypes.c_uint64 * nb)(*offs)
    c_ns = (ctypes.c_size_t * nb)(*ns)
    c_dt = (ctypes.c_int * nb)(*([1] * nb))
    c_np = (ctypes.c_int * nb)(*n_params)
    pb = [b for b, n in enumerate(n_params) for _ in range(n)]
    c_pb = (ctypes.c_int * len(pb))(*pb)
    lib.stub_reset(-1)
    h = lib.tfy_reducer_create(ctx, nb, c_offs, c_ns, c_dt, c_np, len(pb), c_pb, 2)
    assert h
    return h, (ctx, c_offs, c_ns, c_dt, c_np, c_pb)


def _log(lib):
    buf = ctypes.create_string_buffer(1 << 16)
    lib.stub_log(buf, len(buf))
    return [ln for ln in buf.value.decode().splitlines() if ln]


def _launch_offsets(lines):
    return [int(ln.split("off=")[1].split()[0]) for ln in lines if ln.startswith(("allreduce", "fused"))]


def test_buckets_launch_in_index_order_whatever_the_hook_order(lib):
    h, keep = _make(lib, world=2)
    # parameters 0,1 -> bucket 0; 2 -> bucket 1; 3,4,5 -> bucket 2.  Bucket 2 completes FIRST, then 1, then 0.
    for p in (3, 4, 5):
        assert lib.tfy_reducer_mark_ready(h, p, CUR) == 0
    assert _launch_offsets(_log(lib)) == []                           # bucket 2 must wait for 0 and 1
    assert lib.tfy_reducer_mark_ready(h, 2, CUR) == 0
    assert _launch_offsets(_log(lib)) == [] and lib.tfy_reducer_next(h) == 0
    assert lib.tfy_reducer_mark_ready(h, 0, CUR) == 0
    assert lib.tfy_reducer_mark_ready(h, 1, CUR) == 0                 # completes bucket 0 -> 0, 1, 2 go out, in order
    lines = _log(lib)
    assert _launch_offsets(lines) == [4096, 8192, 16384] and lib.tfy_reducer_next(h) == 3
    # every launch: record(ready, compute stream) -> comm stream waits -> collective on the comm stream -> record(done)
    launches = [i for i, ln in enumerate(lines) if ln.startswith("allreduce")]
    for i in launches:
        assert lines[i - 2].startswith("record") and "stream=0xa0" in lines[i - 2]
        assert lines[i - 1].startswith("wait stream=0xc0") and lines[i + 1].startswith("record") \
            and "stream=0xc0" in lines[i + 1]
        assert "stream=0xc0" in lines[i] and "scale=0.5000" in lines[i]           # averaged over 2 ranks
    n_before = len(lines)
    assert lib.tfy_reducer_finalize(h, CUR) == 0
    tail = _log(lib)[n_before:]
    assert len(tail) == 1 and tail[0].startswith("wait stream=0xa0")             # compute stream joins the LAST one only
    assert lib.tfy_reducer_launches(h) == 3 and lib.tfy_reducer_next(h) == 0     # re-armed for the next backward
    lib.tfy_reducer_destroy(h)


def test_finalize_flushes_unready_buckets_and_the_next_backward_starts_clean(lib):
    h, keep = _make(lib, world=4)
    lib.tfy_reducer_mark_ready(h, 0, CUR)
    lib.tfy_reducer_mark_ready(h, 1, CUR)                              # bucket 0 complete and launched
    lib.tfy_reducer_mark_ready(h, 4, CUR)                              # bucket 2 partially ready; bucket 1 unused
    assert _launch_offsets(_log(lib)) == [4096]
    assert lib.tfy_reducer_finalize(h, CUR) == 0                       # unused parameters contribute zeros
    assert _launch_offsets(_log(lib)) == [4096, 8192, 16384]
    lib.stub_reset(-1)
    for p in (5, 4, 3, 2, 1, 0):                                       # second backward, reverse order
        lib.tfy_reducer_mark_ready(h, p, CUR)
    lib.tfy_reducer_mark_ready(h, 0, CUR)                              # a duplicate hook call is ignored
    assert _launch_offsets(_log(lib)) == [4096, 8192, 16384]
    lib.tfy_reducer_finalize(h, CUR)
    assert lib.tfy_reducer_launches(h) == 6
    assert lib.tfy_reducer_mark_ready(h, 99, CUR) == -2                # out-of-range parameter index
    lib.tfy_reducer_destroy(h)


def test_fused_mode_single_rank_and_error_propagation(lib):
    h, keep = _make(lib, world=1)
    for p in range(6):
        lib.tfy_reducer_mark_ready(h, p, CUR)
    lib.tfy_reducer_finalize(h, CUR)
    assert _launch_offsets(_log(lib)) == [] and lib.tfy_reducer_launches(h) == 0     # one rank, nothing fused: no kernels
    lib.stub_reset(-1)
    master = (ctypes.c_float * 8)()
    hyper = ctypes.create_string_buffer(64)
    for b, (poff, shard) in enumerate(((100, 256), (200, 512), (300, 1024))):
        assert lib.tfy_reducer_set_fused(h, b, 1, poff, master, master, master, shard, 2, 0, hyper) == 0
    assert lib.tfy_reducer_set_fused(h, 7, 1, 0, master, master, master, 8, 2, 0, hyper) == -2
    for p in range(6):
        lib.tfy_reducer_mark_ready(h, p, CUR)
    lib.tfy_reducer_finalize(h, CUR)
    fused = [ln for ln in _log(lib) if ln.startswith("fused")]
    assert [ln.split("poff=")[1].split()[0] for ln in fused] == ["100", "200", "300"]
    assert all("zero=1" in ln and "advance=1" in ln and "opt=2" in ln for ln in fused)
    assert lib.tfy_reducer_launches(h) == 3
    lib.tfy_reducer_destroy(h)

    h, keep = _make(lib, world=2)
    lib.stub_reset(8192)                                               # the collective of bucket 1 fails to launch
    rcs = [lib.tfy_reducer_mark_ready(h, p, CUR) for p in (0, 1, 2)]
    assert rcs == [0, 0, -7]
    lib.tfy_reducer_destroy(h)
