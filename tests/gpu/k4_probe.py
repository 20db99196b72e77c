"""Single-GPU probe: how fast can a FEW CTAs run the fused step (LOCAL mode)?  (sizing of the overlap role)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tf_yarn_b200.keras.engine import _solo_communicator  # noqa: E402
from tf_yarn_b200.parallel.comm import FusedShardedOptimizer, OptimizerSpec, _DT, _stream_ptr  # noqa: E402

comm = _solo_communicator(0)
fo = FusedShardedOptimizer(comm, [(1_199_936,)], OptimizerSpec.adadelta(1.0), torch.bfloat16, zero_grads=False)
fo.init_from([torch.full((1_199_936,), 0.1, device="cuda")])
fo.flat_grads.normal_()
c = comm


def run(g0, g1, grid, block):
    rc = c.lib.tfy_fused_step_shard_range(
        c.arena.ctx_ref, _DT[fo.grad_dtype], _DT[fo.param_dtype], fo.spec.code, c.mode, fo.grad_off, fo.param_off,
        fo.shard_n, fo.master.data_ptr(), fo.s1.data_ptr(), fo.s2.data_ptr(), fo.hyper.data_ptr(), 0, grid, block,
        g0, g1, 1, _stream_ptr(None))
    assert rc == 0, rc


def t(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


total = fo.shard_n // 8
for groups in (2000, 4608, 9216, 20000):
    for grid, block in ((4, 256), (9, 256), (18, 256), (36, 256), (148, 256), (888, 128)):
        us = t(lambda: run(total - groups, total, grid, block))
        print(f"groups {groups:6d} grid {grid:4d} x {block}: {us:7.2f} us  ({groups / us:7.1f} groups/us, "
              f"{groups / us / min(grid, 148):6.1f} per CTA)")
