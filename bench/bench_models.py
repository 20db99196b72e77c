#!/usr/bin/env python
"""Secondary benchmarks (BASELINE configs 4 and 5), one process per GPU under torchrun.

    python -m torch.distributed.run --nproc-per-node N bench/bench_models.py --model resnet50 --impl ours
    python -m torch.distributed.run --nproc-per-node N bench/bench_models.py --model bert --impl ours

resnet50 / ours    : PytorchExperiment-style DDP -- torchvision ResNet-50 wrapped by
                     tf_yarn_b200.parallel.ddp (bucketed NVLS all-reduce overlapped with backward), SGD.
resnet50 / standin : torch DistributedDataParallel + NCCL (what the reference's worker does).
bert / ours        : mini-Keras BERT-base, hvd.DistributedOptimizer(Adam): CUDA-graph step + fused
                     reduce-scatter/Adam/all-gather kernel.
bert / standin     : eager bf16 step, NCCL all-reduce of a fused bf16 buffer, torch.optim.Adam.
Device-timed with CUDA events, max over ranks; prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, steps, warmup, world):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / steps


def resnet50(args, rank, local, world):
    import torchvision
    torch.manual_seed(0)
    model = torchvision.models.resnet50().cuda().to(memory_format=torch.channels_last)
    B = args.batch or 64
    x = torch.randn(B, 3, 224, 224, device="cuda").contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (B,), device="cuda")
    if args.impl == "ours":
        from tf_yarn_b200.parallel import runtime
        from tf_yarn_b200.parallel.ddp import DistributedDataParallel
        comm = runtime.get_communicator(device=local)
        ddp = DistributedDataParallel(model, comm, bucket_cap_mb=25)
    else:
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=25) if world > 1 \
            else model
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)

    def step():
        if args.impl == "ours":
            ddp.zero_grad()
        else:
            opt.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(ddp(x).float(), y)
        loss.backward()
        opt.step()
    ms = timed(step, args.steps, args.warmup, world)
    return {"metric": "images/sec ResNet-50 DDP", "value": world * B / (ms * 1e-3), "ms_per_step": ms,
            "per_gpu_batch": B}


def bert(args, rank, local, world):
    from tf_yarn_b200.models.bert import BertForPreTraining, pretraining_loss, synthetic_batch
    B, S = args.batch or 32, 128
    torch.manual_seed(0)
    batches = [synthetic_batch(B, S, seed=rank * 100 + i) for i in range(4)]
    if args.impl == "ours":
        from tf_yarn_b200 import hvd, keras
        hvd.init()
        model = keras.Model.from_torch(BertForPreTraining(), name="bert_base")
        model.compile(loss=pretraining_loss, optimizer=hvd.DistributedOptimizer(keras.optimizers.Adam(1e-4)))
        pinned = [({k: v.pin_memory() for k, v in x.items()}, {k: v.pin_memory() for k, v in y.items()})
                  for x, y in batches]
        model.fit(x=iter(pinned * 2), steps_per_epoch=4, epochs=1, verbose=0)      # builds + captures
        eng = model._engine
        dev = [({k: v.cuda() for k, v in x.items()}, {k: v.cuda() for k, v in y.items()}) for x, y in batches]
        state = {"i": 0}

        def step():
            xb, yb = dev[state["i"] % 4]
            state["i"] += 1
            eng.launch_step(eng.stage_inputs(xb, yb))
        ms = timed(step, args.steps, args.warmup, world)
        eng.stream.synchronize()
    else:
        net = BertForPreTraining().cuda()
        params = list(net.parameters())
        opt = torch.optim.Adam(params, lr=1e-4, fused=True)
        n_total = sum(p.numel() for p in params)
        fusion = torch.zeros(n_total, dtype=torch.bfloat16, device="cuda")
        dev = [({k: v.cuda() for k, v in x.items()}, {k: v.cuda() for k, v in y.items()}) for x, y in batches]
        state = {"i": 0}

        def step():
            xb, yb = dev[state["i"] % 4]
            state["i"] += 1
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(xb)
            loss = pretraining_loss(yb, out)
            loss.backward()
            if world > 1:
                o = 0
                for p in params:
                    fusion[o:o + p.numel()].copy_(p.grad.reshape(-1))
                    o += p.numel()
                dist.all_reduce(fusion)
                o = 0
                for p in params:
                    p.grad.copy_(fusion[o:o + p.numel()].view_as(p.grad))
                    p.grad.mul_(1.0 / world)
                    o += p.numel()
            opt.step()
        ms = timed(step, args.steps, args.warmup, world)
    return {"metric": "sequences/sec BERT-base pre-training (seq 128)", "value": world * B / (ms * 1e-3),
            "ms_per_step": ms, "per_gpu_batch": B}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True, choices=["resnet50", "bert"])
    ap.add_argument("--impl", default="ours", choices=["ours", "standin"])
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), \
        int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    res = (resnet50 if args.model == "resnet50" else bert)(args, rank, local, world)
    res.update({"impl": args.impl, "n_gpus": world, "steps": args.steps, "dtype": "bf16", "data": "synthetic"})
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
