"""utils/fs.py (the cluster_pack.filesystem verbs the reference uses: model_ckpt.py:19-72, parquet_dataset.py:21-36,
worker.py:145-152) and the host-side parts of utils/profiling.py."""
import os

import pytest

from tf_yarn_b200.utils import fs as fsmod
from tf_yarn_b200.utils import profiling


def test_resolve_local_paths_and_reject_remote_schemes(tmp_path):
    fs, path = fsmod.resolve_filesystem_and_path(f"file://{tmp_path}/a/b")
    assert isinstance(fs, fsmod.LocalFileSystem) and path == f"{tmp_path}/a/b"
    fs2, path2 = fsmod.resolve_filesystem_and_path(str(tmp_path))
    assert path2 == str(tmp_path) and fs2.base_fs is fs2
    for remote in ("hdfs://root/user/x", "viewfs://root/x", "s3://bucket/x"):
        with pytest.raises(ValueError, match="unsupported filesystem"):
            fsmod.resolve_filesystem_and_path(remote)


def test_verbs(tmp_path):
    fs = fsmod.LocalFileSystem()
    d = tmp_path / "ckpt"
    fs.mkdir(f"file://{d}/sub")
    assert fs.exists(str(d)) and fs.isdir(str(d / "sub"))
    src = tmp_path / "local.bin"
    src.write_bytes(b"abc")
    fs.put(str(src), str(d / "model_1.pt"))
    fs.put(str(src), str(d / "sub" / "model_2.pt"))
    assert [os.path.basename(p) for p in fs.ls(str(d))] == ["model_1.pt", "sub"]
    assert [os.path.relpath(p, d) for p in fs.ls(str(d), recursive=True)] == ["model_1.pt", "sub/model_2.pt"]
    assert not [n for n in os.listdir(d) if ".tmp" in n]                 # put() publishes atomically
    with fs.open(str(d / "model_1.pt")) as f:
        assert f.read() == b"abc"
    with fs.open(str(d / "notes.txt"), "w") as f:
        f.write("x")
    back = tmp_path / "back.bin"
    fs.get(str(d / "model_1.pt"), str(back))
    assert back.read_bytes() == b"abc"
    fs.rm(str(d / "notes.txt"))
    assert not fs.exists(str(d / "notes.txt"))
    with pytest.raises(OSError):
        fs.rm(str(d / "sub"))                                            # non-empty directory needs recursive=True
    fs.rm(str(d / "sub"), recursive=True)
    assert not fs.exists(str(d / "sub"))
    fs.rm(str(d / "missing"))                                            # removing nothing is not an error


def test_host_side_profiling_helpers(capsys):
    with profiling.nvtx_range("region"):                                 # a no-op without CUDA
        pass
    assert profiling.max_over_ranks(3.5) == 3.5                          # no process group: identity
    with profiling.catchtime("setup"):
        pass
    assert capsys.readouterr().out.startswith("setup: ")
