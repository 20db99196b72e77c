"""Multi-GPU check of the DDP-compatible wrapper (native reducer) against torch DDP + NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
        tests/gpu/ddp_check.py

Same model / data / SGD-momentum on both sides for a few steps: (a) our wrapper + torch.optim.SGD (bucketed NVLS
all-reduce), (b) our wrapper with the optimizer fused into the exchange (K4 per bucket), (c) torch DDP.  Parameters
must agree to fp32 round-off and all ranks must hold identical parameters.
"""
import copy
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def make_model():
    torch.manual_seed(11)
    return nn.Sequential(nn.Conv2d(3, 16, 3, padding=1), nn.ReLU(), nn.Conv2d(16, 32, 3, padding=1), nn.ReLU(),
                         nn.AdaptiveAvgPool2d(4), nn.Flatten(), nn.Linear(512, 700), nn.ReLU(), nn.Linear(700, 10))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from tf_yarn_b200.parallel import runtime
    from tf_yarn_b200.parallel.ddp import DistributedDataParallel
    comm = runtime.get_communicator(device=local)
    base = make_model().cuda()
    ref = torch.nn.parallel.DistributedDataParallel(copy.deepcopy(base), device_ids=[local])
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
    ours = DistributedDataParallel(copy.deepcopy(base), comm, bucket_cap_mb=1)
    ours_opt = torch.optim.SGD(ours.parameters(), lr=0.05, momentum=0.9)
    fused = DistributedDataParallel(copy.deepcopy(base), comm, bucket_cap_mb=1)
    fused_opt = fused.fuse_optimizer("sgd", lr=0.05, momentum=0.9)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    ok = True
    for step in range(6):
        x = torch.randn(16, 3, 8, 8, device="cuda", generator=g)
        y = torch.randint(0, 10, (16,), device="cuda", generator=g)
        for m, o in ((ref, ref_opt), (ours, ours_opt), (fused, fused_opt)):
            o.zero_grad()
            if m is not ref:
                m.zero_grad()
            nn.functional.cross_entropy(m(x), y).backward()
            o.step()
        torch.cuda.synchronize()
    for name, m in (("allreduce", ours), ("fused", fused)):
        worst = 0.0
        for p, q in zip(m.module.parameters(), ref.module.parameters()):
            worst = max(worst, (p.detach() - q.detach()).abs().max().item() / (q.detach().abs().max().item() + 1e-6))
        flat = torch.cat([p.detach().reshape(-1) for p in m.module.parameters()])
        chk = flat.double().sum()
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool((lo == hi).item())
        good = worst < 2e-4 and same
        ok = ok and good
        if rank == 0:
            print(f"[{'PASS' if good else 'FAIL'}] ddp/{name}: max rel diff vs torch DDP {worst:.2e}, ranks identical {same}, "
                  f"kernels launched {m.kernel_launches}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DDP CHECK", "OK" if flag.item() else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
