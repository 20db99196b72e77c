"""The Horovod-shaped facade over a real two-rank gloo group (CPU plumbing path, world_size 2): every collective the
reference's user code reaches (reference: examples/collective_all_reduce_example.py:62-69,
native_keras_with_gloo_example.py:70-78)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["TFY_ROOT"])
from tf_yarn_b200 import hvd, keras
hvd.init()
r, n = hvd.rank(), hvd.size()
assert n == 2 and hvd.local_size() == 2 and hvd.is_initialized()
t = torch.full((4,), float(r + 1))
assert torch.allclose(hvd.allreduce(t), torch.full((4,), 1.5)) and torch.allclose(t, torch.full((4,), float(r + 1)))
assert torch.allclose(hvd.allreduce(t, average=False), torch.full((4,), 3.0))
u = t.clone(); hvd.allreduce_(u); assert torch.allclose(u, torch.full((4,), 1.5))
a, b = torch.full((3,), float(r)), torch.full((2, 2), float(10 * r))
hvd.grouped_allreduce_([a, b], average=True)
assert torch.allclose(a, torch.full((3,), 0.5)) and torch.allclose(b, torch.full((2, 2), 5.0))
v = torch.full((2,), float(r + 7)); assert torch.allclose(hvd.broadcast(v, 1), torch.full((2,), 8.0))
g = hvd.allgather(torch.full((2, 3), float(r)))
assert g.shape == (4, 3) and torch.allclose(g[:2], torch.zeros(2, 3)) and torch.allclose(g[2:], torch.ones(2, 3))
lin = torch.nn.Linear(3, 2)
with torch.no_grad():
    lin.weight.fill_(float(r)); lin.bias.fill_(float(r))
hvd.broadcast_parameters(lin.state_dict(), root_rank=1)
assert float(lin.weight.mean()) == 1.0
hvd.broadcast_variables(list(lin.parameters()), root_rank=0)          # rank 0 now also holds ones
opt = hvd.DistributedOptimizer(torch.optim.SGD(lin.parameters(), lr=1.0, momentum=0.9))
lin.weight.grad = torch.full_like(lin.weight, float(r)); lin.bias.grad = torch.full_like(lin.bias, float(r))
opt.step()                                                               # averaged gradient 0.5 on both ranks
assert torch.allclose(lin.weight.detach(), torch.full((2, 3), 0.5))
hvd.broadcast_optimizer_state(opt, root_rank=0)
cb = hvd.callbacks.MetricAverageCallback()
logs = {"loss": float(r), "accuracy": 1.0 - r}
cb.on_epoch_end(0, logs)
assert logs == {"loss": 0.5, "accuracy": 0.5}
hvd.barrier()
hvd.shutdown()
print("RANK_OK", r, flush=True)
'''


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.slow
def test_collectives_over_gloo_with_two_ranks(tmp_path):
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, TFY_ROOT=ROOT, HOROVOD_RANK=str(rank), HOROVOD_SIZE="2", HOROVOD_LOCAL_RANK=str(rank),
                   HOROVOD_LOCAL_SIZE="2", HOROVOD_GLOO_RENDEZVOUS_ADDR="127.0.0.1",
                   HOROVOD_GLOO_RENDEZVOUS_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {rank}" in out, out[-3000:]
