"""GPU tests of the B200 train engines (run with `pytest -m gpu` on a B200).

Numerics of every hand-written kernel on the single-GPU path are compared with a plain
PyTorch fp32 reference of the same op.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mnist_like(drop1=0.0, drop2=0.0):
    from tf_yarn_b200 import keras
    from tf_yarn_b200.keras import layers
    m = keras.Sequential(name="mnist_cnn_test")
    m.add(layers.Conv2D(32, (3, 3), activation="relu", input_shape=(28, 28, 1)))
    m.add(layers.Conv2D(64, (3, 3), activation="relu"))
    m.add(layers.MaxPooling2D(pool_size=(2, 2)))
    m.add(layers.Dropout(drop1))
    m.add(layers.Flatten())
    m.add(layers.Dense(128, activation="relu"))
    m.add(layers.Dropout(drop2))
    m.add(layers.Dense(10))
    return m


def _native_loaded():
    from tf_yarn_b200.ops import native
    assert native.load() is not None


def test_native_library_loads():
    _native_loaded()


@pytest.mark.parametrize("kind", ["sgd", "adadelta", "adam", "adagrad"])
def test_fused_step_local_matches_torch_optim(kind):
    """K4 in LOCAL mode (world=1) against torch.optim on fp32 tensors."""
    from tf_yarn_b200.keras.engine import _solo_communicator
    from tf_yarn_b200.parallel.comm import FusedShardedOptimizer, OptimizerSpec
    comm = _solo_communicator(0)
    shapes = [(33, 17), (129,), (64, 32, 3, 3)]
    spec = {"sgd": OptimizerSpec.sgd(0.05, momentum=0.9, nesterov=True),
            "adadelta": OptimizerSpec.adadelta(1.0, 0.95, 1e-7),
            "adam": OptimizerSpec.adam(1e-3, weight_decay=0.01),
            "adagrad": OptimizerSpec.adagrad(0.05, 1e-10, 0.0, 0.1)}[kind]
    fo = FusedShardedOptimizer(comm, shapes, spec, param_dtype=torch.float32, grad_dtype=torch.float32)
    g = torch.Generator(device="cuda").manual_seed(3)
    init = [torch.randn(s, device="cuda", generator=g) * 0.3 for s in shapes]
    fo.init_from(init, broadcast_root=None)
    ref = [t.clone().requires_grad_(True) for t in init]
    opt = {"sgd": lambda: torch.optim.SGD(ref, lr=0.05, momentum=0.9, nesterov=True),
           "adadelta": lambda: torch.optim.Adadelta(ref, lr=1.0, rho=0.95, eps=1e-7),
           "adam": lambda: torch.optim.Adam(ref, lr=1e-3, weight_decay=0.01),
           "adagrad": lambda: torch.optim.Adagrad(ref, lr=0.05, eps=1e-10, initial_accumulator_value=0.1)}[kind]()
    for step in range(5):
        grads = [torch.randn(s, device="cuda", generator=g) * 0.5 for s in shapes]
        for gv, gr, r in zip(fo.grad_views, grads, ref):
            gv.copy_(gr)
            r.grad = gr.clone()
        fo.step()
        opt.step()
        torch.cuda.synchronize()
        for pv, r in zip(fo.param_views, ref):
            torch.testing.assert_close(pv, r.detach(), rtol=2e-5, atol=2e-6)
        assert all(float(gv.abs().max()) == 0.0 for gv in fo.grad_views), "gradients must be zeroed by the step"
    assert fo.step_count == 5


def test_nn_kernels_match_torch():
    """bias+relu+pool(+dropout off), its backward, act/bias backward, conv C_in=1 fwd/wgrad, xent head."""
    import ctypes
    import torch.nn.functional as F
    from tf_yarn_b200.keras import fastpath  # noqa: F401  (declares the argtypes)
    from tf_yarn_b200.ops import native
    lib = native.load()
    s = torch.cuda.current_stream().cuda_stream
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    B, H, W, C = 16, 24, 24, 64
    partial = torch.zeros(592 * 9 * 64, device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)

    z = (torch.randn(B, H, W, C, device=dev, generator=g)).bfloat16()
    bias = (torch.randn(C, device=dev, generator=g) * 0.1).bfloat16()
    p = torch.empty(B, H // 2, W // 2, C, dtype=torch.bfloat16, device=dev)
    code = torch.empty(B, H // 2, W // 2, C, dtype=torch.uint8, device=dev)
    assert lib.tfy_bias_relu_pool_drop_fwd(z.data_ptr(), bias.data_ptr(), p.data_ptr(), code.data_ptr(), B, H, W, C,
                                           0.0, 1, None, s) == 0
    zr = (z.float() + bias.float()).permute(0, 3, 1, 2).requires_grad_(True)
    pr = F.max_pool2d(F.relu(zr), 2)
    torch.testing.assert_close(p.float(), pr.permute(0, 2, 3, 1).detach(), rtol=1e-2, atol=1e-2)
    dp = (torch.randn(B, H // 2, W // 2, C, device=dev, generator=g)).bfloat16()
    dz = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev)
    dbias = torch.empty(C, dtype=torch.bfloat16, device=dev)
    assert lib.tfy_pool_drop_relu_bwd(dp.data_ptr(), code.data_ptr(), dz.data_ptr(), 1.0, B, H, W, C,
                                      partial.data_ptr(), dbias.data_ptr(), counter.data_ptr(), s) == 0
    pr.backward(dp.float().permute(0, 3, 1, 2))
    torch.testing.assert_close(dz.float(), zr.grad.permute(0, 2, 3, 1), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(dbias.float(), zr.grad.sum((0, 2, 3)), rtol=2e-2, atol=5e-2)
    assert int(counter.item()) == 0

    # bias + relu (+ mask) forward, gated backward + bias gradient
    rows, Cd = 128, 128
    zd = torch.randn(rows, Cd, device=dev, generator=g).bfloat16()
    bd = (torch.randn(Cd, device=dev, generator=g) * 0.1).bfloat16()
    y = torch.empty_like(zd)
    mask = torch.empty(rows, Cd, dtype=torch.uint8, device=dev)
    assert lib.tfy_bias_act_drop_fwd(zd.data_ptr(), bd.data_ptr(), y.data_ptr(), mask.data_ptr(), rows, Cd, 1, 0.0, 7,
                                     None, s) == 0
    yr = F.relu(zd.float() + bd.float())
    torch.testing.assert_close(y.float(), yr, rtol=1e-2, atol=1e-2)
    dy = torch.randn(rows, Cd, device=dev, generator=g).bfloat16()
    dzz = torch.empty_like(dy)
    db = torch.empty(Cd, dtype=torch.bfloat16, device=dev)
    assert lib.tfy_act_drop_bwd_bias(dy.data_ptr(), mask.data_ptr(), None, dzz.data_ptr(), 1.0, rows, Cd,
                                     partial.data_ptr(), db.data_ptr(), counter.data_ptr(), s) == 0
    ref = dy.float() * (yr > 0)
    torch.testing.assert_close(dzz.float(), ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(db.float(), ref.sum(0), rtol=2e-2, atol=5e-2)
    # dropout: keep fraction and scaling
    yd = torch.empty_like(zd)
    assert lib.tfy_bias_act_drop_fwd(zd.data_ptr(), None, yd.data_ptr(), mask.data_ptr(), rows, Cd, 0, 0.5, 11,
                                     None, s) == 0
    kept = (yd != 0).float().mean().item()
    assert 0.45 < kept < 0.55, kept
    nz = yd != 0
    torch.testing.assert_close(yd.float()[nz], (zd.float() * 2)[nz], rtol=1e-2, atol=1e-2)

    # conv 3x3 C_in=1 forward + weight gradient
    Bc, Hc, Wc, O = 8, 28, 28, 32
    x = torch.rand(Bc, Hc, Wc, 1, device=dev, generator=g)
    w = (torch.randn(O, 3, 3, 1, device=dev, generator=g) * 0.3).bfloat16()
    bc = (torch.randn(O, device=dev, generator=g) * 0.1).bfloat16()
    yc = torch.empty(Bc, Hc - 2, Wc - 2, O, dtype=torch.bfloat16, device=dev)
    assert lib.tfy_conv3x3_c1_fwd(x.data_ptr(), 1, w.data_ptr(), bc.data_ptr(), yc.data_ptr(), Bc, Hc, Wc, O, s) == 0
    wr = w.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)      # [O, 1, 3, 3]
    ycr = F.relu(F.conv2d(x.permute(0, 3, 1, 2), wr, bc.float()))
    torch.testing.assert_close(yc.float(), ycr.permute(0, 2, 3, 1).detach(), rtol=1e-2, atol=1e-2)
    dzc = torch.randn(Bc, Hc - 2, Wc - 2, O, device=dev, generator=g).bfloat16()
    dw = torch.empty(O, 9, dtype=torch.bfloat16, device=dev)
    assert lib.tfy_conv3x3_c1_wgrad(x.data_ptr(), 1, dzc.data_ptr(), partial.data_ptr(), dw.data_ptr(),
                                    counter.data_ptr(), Bc, Hc, Wc, O, s) == 0
    pre = F.conv2d(x.permute(0, 3, 1, 2), wr)
    pre.backward(dzc.float().permute(0, 3, 1, 2))
    torch.testing.assert_close(dw.float().view(O, 3, 3), wr.grad[:, 0], rtol=2e-2, atol=0.3)

    # softmax cross-entropy head
    Bx, Cx = 128, 10
    lg = torch.randn(Bx, Cx, device=dev, generator=g).bfloat16()
    bx = (torch.randn(Cx, device=dev, generator=g) * 0.1).bfloat16()
    lab = torch.randint(0, Cx, (Bx,), device=dev, generator=g)
    loss = torch.zeros((), device=dev)
    dl = torch.empty(Bx, Cx, dtype=torch.bfloat16, device=dev)
    dbx = torch.empty(Cx, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2, device=dev)
    assert lib.tfy_softmax_xent(lg.data_ptr(), bx.data_ptr(), lab.data_ptr(), loss.data_ptr(), dl.data_ptr(),
                                dbx.data_ptr(), stats.data_ptr(), Bx, Cx, s) == 0
    lr_ = (lg.float() + bx.float()).requires_grad_(True)
    lref = F.cross_entropy(lr_, lab)
    lref.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - lref.item()) < 1e-3
    torch.testing.assert_close(dl.float(), lr_.grad, rtol=2e-2, atol=1e-4)
    torch.testing.assert_close(dbx.float(), lr_.grad.sum(0), rtol=2e-2, atol=1e-3)
    assert stats[1].item() == Bx and stats[0].item() == (lr_.argmax(1) == lab).sum().item()


def _run_steps(fast: bool, steps: int = 4):
    from tf_yarn_b200 import keras
    torch.manual_seed(7)
    m = _mnist_like(0.0, 0.0)
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True),
              optimizer=keras.optimizers.Adadelta(1.0), metrics=["accuracy"], use_fast_path=fast)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(128 * steps, 28, 28, 1, generator=g)
    y = torch.randint(0, 10, (128 * steps,), generator=g)
    losses = []

    class Rec(keras.callbacks.Callback):
        needs_batch_logs = True

        def on_train_batch_end(self, batch, logs=None):
            losses.append(logs["loss"])
    m.fit(x, y, batch_size=128, epochs=1, shuffle=False, verbose=0, callbacks=[Rec()])
    return m, losses


def test_fastpath_matches_autograd_engine():
    """Same init, same batches, dropout off: the fused-kernel plan must track the autograd engine."""
    from tf_yarn_b200.keras.fastpath import FastSequentialEngine
    m_fast, l_fast = _run_steps(True)
    m_ref, l_ref = _run_steps(False)
    assert isinstance(m_fast._engine, FastSequentialEngine)
    assert not isinstance(m_ref._engine, FastSequentialEngine)
    assert len(l_fast) == len(l_ref) == 4
    for a, b in zip(l_fast, l_ref):
        assert math.isfinite(a) and abs(a - b) < 0.03 * max(1.0, abs(b)), (l_fast, l_ref)
    wf, wr = m_fast.get_weights(), m_ref.get_weights()
    for a, b in zip(wf, wr):
        a, b = torch.as_tensor(a), torch.as_tensor(b)
        # Adadelta's early updates are sign-like (|dx| ~ 1.4e-3 per step whatever |g| is), so bf16
        # noise on a near-zero gradient moves a weight by a full update: compare on that scale
        denom = b.abs().max().clamp_min(2e-2)
        assert ((a - b).abs().max() / denom) < 0.15, ((a - b).abs().max(), denom)
    assert m_fast._engine.kernel_launches > 0


def test_fit_learns_and_checkpoints(tmp_path):
    from tf_yarn_b200 import keras
    torch.manual_seed(0)
    m = _mnist_like(0.25, 0.5)
    m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True),
              optimizer=keras.optimizers.Adadelta(1.0), metrics=["accuracy"])
    g = torch.Generator().manual_seed(1)
    y = torch.randint(0, 10, (1024,), generator=g)
    x = torch.rand(1024, 28, 28, 1, generator=g) * 0.5
    for i in range(1024):                       # class-dependent bright patch: easy to learn
        c = int(y[i])
        x[i, (c // 5) * 14:(c // 5) * 14 + 14, (c % 5) * 5:(c % 5) * 5 + 5, 0] += 0.5
    h = m.fit(x, y, batch_size=128, epochs=8, shuffle=True, verbose=0)
    assert h.history["loss"][-1] < h.history["loss"][0]
    assert h.history["accuracy"][-1] > 0.7, h.history
    path = str(tmp_path / "m.ckpt")
    m.save(path)
    m2 = keras.models.load_model(path)
    acc = m2.evaluate(x, y, batch_size=256, return_dict=True)["accuracy"]
    assert acc > 0.7


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


@pytest.mark.parametrize("shape", [(128, 128, 64), (128, 128, 9216), (256, 384, 512), (100, 72, 200), (1000, 10, 128),
                                   (128, 9216, 128)])
def test_tcgen05_gemm_matches_fp32_reference(shape):
    """tcgen05/TMEM/TMA GEMM vs a plain fp32 matmul of the same bf16 inputs (tails, bias, relu)."""
    from tf_yarn_b200.ops.gemm import gemm_bf16
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()
    bias = (torch.randn(N, device="cuda", generator=g)).bfloat16()
    ref = a.float() @ b.float().t()
    out = gemm_bf16(a, b)
    torch.cuda.synchronize()
    tol = 2e-2 * (K ** 0.5) * 0.25 + 1e-2
    assert (out.float() - ref).abs().max().item() < max(tol, 0.01 * ref.abs().max().item()), \
        (out.float() - ref).abs().max().item()
    out2 = gemm_bf16(a, b, bias=bias, relu=True)
    ref2 = torch.relu(ref + bias.float())
    assert (out2.float() - ref2).abs().max().item() < max(tol, 0.01 * ref2.abs().max().item())
    if K >= 256:
        out3 = gemm_bf16(a, b, split_k=4)
        torch.cuda.synchronize()
        assert out3.dtype == torch.float32
        assert (out3 - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("batch", [2, 16])
def test_tcgen05_conv_kernels_match_fp32_reference(batch, monkeypatch, tmp_path):
    """Implicit-GEMM conv fprop(+bias/relu/pool/dropout), dgrad(+gate), wgrad vs fp32 PyTorch references."""
    import importlib.util
    import os
    import sys
    spec = importlib.util.spec_from_file_location(
        "conv_check", os.path.join(os.path.dirname(__file__), "gpu", "conv_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["conv_check.py", str(batch)])
    assert mod.main() == 0, mod.res


@pytest.mark.parametrize("B,K,C", [(128, 128, 10), (100, 256, 16), (37, 512, 3)])
def test_fused_dense_head_matches_autograd(B, K, C):
    """Fused classifier head (fwd + bwd + previous Dense gate/bias-grad) vs fp32 autograd of the same math."""
    import ctypes
    from tf_yarn_b200.keras import fastpath  # noqa: F401  (declares the kernel)
    from tf_yarn_b200.ops import native
    lib = native.load()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(B + K + C)
    h = torch.randn(B, K, device=dev, generator=g).relu().bfloat16()
    w2 = (torch.randn(C, K, device=dev, generator=g) * 0.2).bfloat16()
    b2 = (torch.randn(C, device=dev, generator=g) * 0.1).bfloat16()
    y = torch.randint(0, C, (B,), device=dev, generator=g)
    mask = (torch.rand(B, K, device=dev, generator=g) > 0.4).to(torch.uint8)
    scale = 2.0
    loss = torch.zeros((), device=dev)
    stats = torch.zeros(2, device=dev)
    dw2 = torch.zeros(C, K, dtype=torch.bfloat16, device=dev)
    db2 = torch.zeros(C, dtype=torch.bfloat16, device=dev)
    dh = torch.zeros(B, K, dtype=torch.bfloat16, device=dev)
    db1 = torch.zeros(K, dtype=torch.bfloat16, device=dev)
    scratch = torch.zeros(int(lib.tfy_dense_head_scratch_elems(K, C)), device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    # reference
    hf = h.float().requires_grad_(True)
    wf = w2.float().requires_grad_(True)
    bf = b2.float().requires_grad_(True)
    logits = hf @ wf.t() + bf
    ref_loss = torch.nn.functional.cross_entropy(logits, y)
    ref_loss.backward()
    ref_dh = hf.grad * mask.float() * scale
    for _ in range(2):      # twice: the scratch buffer and the counter must come back clean
        rc = lib.tfy_dense_head_fused(h.data_ptr(), w2.data_ptr(), b2.data_ptr(), y.data_ptr(), mask.data_ptr(),
                                      ctypes.c_float(scale), loss.data_ptr(), None, stats.data_ptr(), dw2.data_ptr(),
                                      db2.data_ptr(), dh.data_ptr(), db1.data_ptr(), scratch.data_ptr(),
                                      counter.data_ptr(), B, K, C, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()

        def close(a, b, tol=0.02):
            scale_ = b.abs().max().item() + 1e-6
            assert (a.float() - b).abs().max().item() <= tol * scale_, ((a.float() - b).abs().max().item(), scale_)
        assert abs(loss.item() - ref_loss.item()) < 1e-3 * max(1.0, abs(ref_loss.item()))
        close(dw2, wf.grad)
        close(db2, bf.grad)
        close(dh, ref_dh)
        close(db1, ref_dh.sum(0))
        assert bool((scratch == 0).all()) and int(counter.item()) == 0
    assert stats[1].item() == 2 * B
    assert stats[0].item() == 2 * (logits.argmax(1) == y).sum().item()


@pytest.mark.parametrize("B,U,I,want_dx", [(128, 128, 9216, True), (128, 128, 9216, False), (64, 128, 1000, True),
                                            (100, 72, 200, True), (16, 8, 64, True),
                                            # batch / width beyond one tile: the tiled kernel (K loops, TMA ring)
                                            (512, 1024, 1680, True), (300, 136, 264, True), (512, 256, 1024, False)])
def test_tcgen05_dense_backward_matches_fp32_reference(B, U, I, want_dx):
    """dW = dh^T x and dx = dh W out of ONE tcgen05 kernel (MN-major operand views of the same tiles, TMA-store
    epilogue) vs fp32 matmuls of the same bf16 inputs; ragged extents exercise the TMA zero-fill / clipping."""
    from tf_yarn_b200.keras import fastpath  # noqa: F401  (declares the kernel)
    from tf_yarn_b200.ops import native
    lib = native.load()
    g = torch.Generator(device="cuda").manual_seed(B * 3 + U * 5 + I)
    dh = (torch.randn(B, U, device="cuda", generator=g) * 0.5).bfloat16()
    x = (torch.randn(B, I, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(U, I, device="cuda", generator=g) * 0.5).bfloat16()
    dw = torch.full((U, I), 7.0, dtype=torch.bfloat16, device="cuda")
    dx = torch.full((B, I), 7.0, dtype=torch.bfloat16, device="cuda") if want_dx else None
    rc = lib.tfy_dense_bwd(dh.data_ptr(), x.data_ptr(), w.data_ptr(), dw.data_ptr(),
                           dx.data_ptr() if want_dx else None, B, U, I, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    ref_dw = dh.float().t() @ x.float()
    assert (dw.float() - ref_dw).abs().max().item() <= 0.01 * ref_dw.abs().max().item() + 0.02, \
        (dw.float() - ref_dw).abs().max().item()
    if want_dx:
        ref_dx = dh.float() @ w.float()
        assert (dx.float() - ref_dx).abs().max().item() <= 0.01 * ref_dx.abs().max().item() + 0.02, \
            (dx.float() - ref_dx).abs().max().item()


def test_overlapped_exchange_matches_single_launch(monkeypatch):
    """The communication CTAs inside the conv weight-gradient kernel + the trailing ranged step train like the classic
    single fused-step launch (same data, same seeds).  Adadelta's first updates are sign-like (|update| ~ 1.4e-3
    whatever |g|), so ulp-level differences in near-zero gradients flip individual updates: the comparison is on the
    loss trajectory and on the aggregate distance of the parameters, not element by element."""
    import numpy as np
    from tf_yarn_b200 import keras

    def run(overlap):
        monkeypatch.setenv("TFY_OVERLAP", "1" if overlap else "0")
        monkeypatch.setenv("TFY_OVERLAP_GROUPS", "20000")
        torch.manual_seed(7)
        m = _mnist_like(0.25, 0.5)
        m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True),
                  optimizer=keras.optimizers.Adadelta(1.0))
        rs = np.random.RandomState(0)
        y = rs.randint(0, 10, 512).astype("int64")
        x = (rs.rand(512, 28, 28, 1) * 0.5 + (y[:, None, None, None] / 20.0)).astype("float32")
        h = m.fit(torch.from_numpy(x), torch.from_numpy(y), batch_size=128, epochs=3, shuffle=False, verbose=0)
        eng = m._engine
        torch.cuda.synchronize()
        return eng, eng.fused.master.clone(), h.history["loss"]

    eng0, m0, l0 = run(False)
    eng1, m1, l1 = run(True)
    assert eng1._ov_gmid > 0, "the overlap role was not selected"
    assert eng1.fused.step_count == eng0.fused.step_count == 12
    assert all(abs(a - b) <= 0.03 * max(abs(a), 1e-3) + 0.01 for a, b in zip(l0, l1)), (l0, l1)
    rel = ((m1 - m0).norm() / m0.norm()).item()
    assert rel < 0.02, rel


def test_fused_splitk_epilogue_matches_two_launch_path(monkeypatch):
    """TFY_SPLITK_FUSE=1 (bias / ReLU / dropout applied by the split-K GEMM's own CTAs after they meet on the tile
    counter) trains like the default GEMM + bias_act_drop pair: same dropout stream, same loss trajectory."""
    import numpy as np
    from tf_yarn_b200 import keras

    def run(fuse):
        monkeypatch.setenv("TFY_SPLITK_FUSE", "1" if fuse else "0")
        torch.manual_seed(7)
        m = _mnist_like(0.25, 0.5)
        m.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True),
                  optimizer=keras.optimizers.Adadelta(1.0))
        rs = np.random.RandomState(0)
        y = rs.randint(0, 10, 512).astype("int64")
        x = (rs.rand(512, 28, 28, 1) * 0.5 + (y[:, None, None, None] / 20.0)).astype("float32")
        h = m.fit(torch.from_numpy(x), torch.from_numpy(y), batch_size=128, epochs=3, shuffle=False, verbose=0)
        torch.cuda.synchronize()
        return m._engine, m._engine.fused.master.clone(), h.history["loss"]

    e0, m0, l0 = run(False)
    e1, m1, l1 = run(True)
    assert e1._launches_per_step == e0._launches_per_step - 1
    assert all(abs(a - b) <= 0.03 * max(abs(a), 1e-3) + 0.01 for a, b in zip(l0, l1)), (l0, l1)
    assert ((m1 - m0).norm() / m0.norm()).item() < 0.02


@pytest.mark.parametrize("shape", [(256, 256, 64), (512, 768, 512), (1000, 520, 264), (4096, 1024, 128), (130, 304, 72)])
def test_tcgen05_2cta_gemm_matches_fp32_reference(shape):
    """Persistent cta_group::2 GEMM (256x256 tiles of a CTA pair, double-buffered TMEM, TMA-store epilogue) vs fp32."""
    from tf_yarn_b200.ops.gemm import gemm_bf16
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 7 * K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g).bfloat16()
    ref = a.float() @ b.float().t()
    out = gemm_bf16(a, b, impl="2cta")
    torch.cuda.synchronize()
    assert (out.float() - ref).abs().max().item() < max(0.02 * K ** 0.5 * 0.25 + 0.01, 0.01 * ref.abs().max().item())
    out2 = gemm_bf16(a, b, bias=bias, relu=True, impl="2cta")
    ref2 = torch.relu(ref + bias.float())
    torch.cuda.synchronize()
    assert (out2.float() - ref2).abs().max().item() < max(0.02 * K ** 0.5 * 0.25 + 0.02, 0.01 * ref2.abs().max().item())


@pytest.mark.parametrize("B,N,T,n_num", [(200, 264, 3, 13), (512, 1024, 26, 13), (64, 8, 1, 0)])
def test_ps_gather_gemm_matches_fp32_reference(B, N, T, n_num):
    """K5: embedding-row gather (A-operand producer) + remote-weight tcgen05 GEMM + bias in one kernel vs the unfused
    fp32 computation (same bf16 rounding of activations and weights); the gathered activations it keeps for the
    backward must equal the gathered rows."""
    from tf_yarn_b200.estimator import ps_hbm  # noqa: F401  (declares the kernel)
    from tf_yarn_b200.ops import native
    lib = native.load()
    V = 1000
    g = torch.Generator(device="cuda").manual_seed(B + N + T)
    tables = [torch.randn(V, 64, device="cuda", generator=g) * 0.3 for _ in range(T)]
    ids = torch.randint(0, V, (T, B), device="cuda", generator=g)
    numeric = torch.randn(B, max(n_num, 1), device="cuda", generator=g)[:, :n_num].contiguous()
    K = 64 * T + n_num
    Kp = (K + 7) // 8 * 8
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.1)
    shadow = torch.zeros(N, Kp, dtype=torch.bfloat16, device="cuda")
    shadow[:, :K] = w
    bias = (torch.randn(N, device="cuda", generator=g) * 0.1).bfloat16()
    ptrs = torch.tensor([t.data_ptr() for t in tables], dtype=torch.int64, device="cuda")
    xbuf = torch.zeros(B, Kp, dtype=torch.bfloat16, device="cuda")
    y = torch.zeros(B, N, dtype=torch.bfloat16, device="cuda")
    rc = lib.tfy_ps_gather_gemm(shadow.data_ptr(), ptrs.data_ptr(), ids.data_ptr(),
                                numeric.data_ptr() if n_num else None, bias.data_ptr(), xbuf.data_ptr(), y.data_ptr(), B,
                                N, T, n_num, Kp, V, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    x = torch.cat([tables[t][ids[t]] for t in range(T)] + ([numeric] if n_num else []), dim=1)
    xb = x.bfloat16()
    assert torch.equal(xbuf[:, :K], xb)
    ref = xb.float() @ shadow[:, :K].float().t() + bias.float()
    assert (y.float() - ref).abs().max().item() <= 0.01 * ref.abs().max().item() + 0.02
