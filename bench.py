#!/usr/bin/env python
"""Headline benchmark: Keras MNIST-CNN, Horovod path, samples/sec on N B200s.

Metric / config are the ones BASELINE.json names: samples/sec for the whole box
(device-timed with CUDA events, max over ranks) for the Keras MNIST-CNN
Horovod-path config, bf16 compute, synthetic data of the MNIST shape,
random-init weights; weak scaling (128 samples per GPU per step, the
Horovod keras_mnist batch size).

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 200 --warmup 20

Arms:
  --impl ours       (default) mini-Keras + hvd facade; every step = forward/backward + ONE fused
                    reduce-scatter -> Adadelta -> all-gather kernel, replayed from a CUDA graph.
  --impl reference  the unmodified reference from baseline/_ref (not runnable offline: prints why).
  --impl standin    what the reference's Horovod path does on an NCCL build, written with stock
                    PyTorch: eager bf16 step, NCCL all-reduce of a fused gradient buffer, fp32
                    cast/scale, torch.optim.Adadelta.  Not the reference itself -- our yardstick.

`value` is timed on the device with the input pool (256 MiB, larger than the 126 MB L2, rotated
so no batch is re-read from L2) resident in HBM.  `e2e` runs the same number of steps through
the public API (`model.fit`) with per-step H2D copies from pinned host memory and a per-step
D2H read of the loss.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 128
POOL_BATCHES = 672          # 672 * 128 * 784 * 4 B = 269.7 MB  > 126 MB L2
METRIC = "samples/sec, Keras MNIST-CNN Horovod path (whole job)"


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML on a thread while the timed region runs
    (10 ms period, so even a ~100 ms region gets ~10 samples; nvidia-smi -lms cannot go that fast)."""

    REASONS = {
        "hw_slowdown": 0x0000000000000008, "sw_power_cap": 0x0000000000000004,
        "hw_thermal_slowdown": 0x0000000000000040, "sw_thermal_slowdown": 0x0000000000000020,
        "hw_power_brake_slowdown": 0x0000000000000080,
    }

    def __init__(self, gpu_index: int, period_s: float = 0.01):
        self.gpu, self.period = gpu_index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        self._h = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self._h = None
            return
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def stop(self):
        if self._h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self._stop.set()
        self._t.join(2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def _dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def run_reference(args):
    why = ("reference installs into baseline/_ref only with --no-deps and cannot be imported: it needs skein, "
           "cluster_pack, tensorflow and horovod, none of which are in the image or /opt/wheelhouse")
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import tf_yarn  # noqa: F401
        why = "reference imported but its Keras/Horovod path needs TensorFlow + Horovod + a YARN cluster"
    except Exception as exc:  # noqa: BLE001
        why += f" (import error: {type(exc).__name__}: {exc})"
    rank, _, _ = _dist_env()
    if rank == 0:
        print(json.dumps({"impl": "reference", "unavailable": why}))
    return 0


def _max_over_ranks(value: float, world: int) -> float:
    import torch
    import torch.distributed as dist
    if world == 1:
        return value
    t = torch.tensor([value], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world: int):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def make_pool(seed: int):
    import torch
    g = torch.Generator().manual_seed(seed)
    n = POOL_BATCHES * PER_GPU_BATCH
    x = torch.rand((n, 28, 28, 1), generator=g)
    y = torch.randint(0, 10, (n,), generator=g)
    return x, y


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local, world = _dist_env()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL_DEBUG=VERSION makes NCCL print its version banner on stdout, in front of the JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from tf_yarn_b200 import hvd, keras
    from tf_yarn_b200.models.mnist_cnn import keras_mnist_cnn

    hvd.init()
    torch.manual_seed(1234)      # identical init on every rank (BroadcastGlobalVariables also runs)
    model = keras_mnist_cnn(logits=True)
    opt = hvd.DistributedOptimizer(keras.optimizers.Adadelta(1.0 * hvd.size()))
    model.compile(loss=keras.losses.SparseCategoricalCrossentropy(from_logits=True), optimizer=opt)
    callbacks = [hvd.callbacks.BroadcastGlobalVariablesCallback(0)]

    x_host, y_host = make_pool(seed=100 + rank)
    x_host, y_host = x_host.pin_memory(), y_host.pin_memory()
    h2d_bytes = PER_GPU_BATCH * (28 * 28 * 4 + 8)

    # ---- warm-up through the public API (builds the engine, captures the graph) ----------------
    warm = max(3, args.warmup)
    model.fit(x_host[:warm * PER_GPU_BATCH], y_host[:warm * PER_GPU_BATCH], batch_size=PER_GPU_BATCH, epochs=1,
              shuffle=False, verbose=0, callbacks=callbacks)
    eng = model._engine
    torch.cuda.synchronize()

    # ---- device-timed region: inputs resident in HBM, pool larger than L2 ---------------------
    x_dev, y_dev = x_host.cuda(non_blocking=True), y_host.cuda(non_blocking=True)
    torch.cuda.synchronize()

    pool = [(x_dev[b * PER_GPU_BATCH:(b + 1) * PER_GPU_BATCH], y_dev[b * PER_GPU_BATCH:(b + 1) * PER_GPU_BATCH])
            for b in range(POOL_BATCHES)]          # views, made once: the timed loop only launches

    def device_steps(n, start_batch):
        b = start_batch
        ticket = eng.stage_inputs(*pool[b])
        for i in range(n):
            eng.launch_step(ticket)
            if i + 1 < n:
                b = (b + 1) % POOL_BATCHES
                ticket = eng.stage_inputs(*pool[b])
        return (b + 1) % POOL_BATCHES

    stream_ctx = torch.cuda.stream(eng.stream)      # replays are issued from the engine's stream (as fit() does)
    stream_ctx.__enter__()
    nxt = device_steps(warm, 0)
    launches0 = eng.kernel_launches
    sampler = ClockSampler(local)
    _barrier(world)
    if rank == 0:
        sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(eng.stream)
    device_steps(args.steps, nxt)
    end.record(eng.stream)
    stream_ctx.__exit__(None, None, None)
    torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    _barrier(world)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.kernel_launches - launches0
    ms = _max_over_ranks(ms, world)
    ms_per_step = ms / args.steps
    value = world * PER_GPU_BATCH / (ms_per_step * 1e-3)

    # ---- end-to-end region: model.fit with per-step H2D (pinned) + per-step loss D2H ----------
    # exactly K steps through the public API, fed by an iterable that walks the pinned host pool
    class _HostPool:
        cardinality = None

        def __init__(self, first_batch):
            self.b = first_batch

        def __iter__(self):
            while True:
                b = self.b
                self.b = (b + 1) % POOL_BATCHES
                yield (x_host[b * PER_GPU_BATCH:(b + 1) * PER_GPU_BATCH],
                       y_host[b * PER_GPU_BATCH:(b + 1) * PER_GPU_BATCH])

    e2e_steps = args.steps
    _barrier(world)
    t0 = time.perf_counter()
    hist = model.fit(_HostPool(warm % POOL_BATCHES), steps_per_epoch=e2e_steps, epochs=1, verbose=0)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    _barrier(world)
    e2e_ms = _max_over_ranks((t1 - t0) * 1e3, world)
    e2e_value = world * PER_GPU_BATCH * e2e_steps / (e2e_ms * 1e-3)
    final_loss = hist.history["loss"][-1]

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic (MNIST-shaped), random-init weights",
            "impl": "ours",
            "config": {"model": "Keras MNIST-CNN (1,199,882 params), Adadelta(1.0*size), Horovod path",
                       "global_batch": world * PER_GPU_BATCH, "per_gpu_batch": PER_GPU_BATCH,
                       "seq_len": None, "parallelism": f"dp{world}",
                       "l2": "inputs rotate through a 269.7 MB pool (> 126 MB L2); no flush",
                       "comm": ("NVLS multimem" if eng.comm.multicast else ("P2P" if world > 1 else "local")),
                       "cuda_graph": eng.graph is not None},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "final_loss": final_loss},
            "gpu_launches": launches,
        }
        print(json.dumps(out))
    if world > 1:
        hvd.shutdown()
        dist.destroy_process_group()
    return 0


def run_standin(args):
    """Stock-PyTorch rendition of the reference's Horovod path on an NCCL build (our yardstick)."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    rank, local, world = _dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from tf_yarn_b200.models.mnist_cnn import TorchMnistCnn
    torch.manual_seed(1234)
    net = TorchMnistCnn().cuda().to(memory_format=torch.channels_last)
    params = [p for p in net.parameters()]
    opt = torch.optim.Adadelta(params, lr=1.0 * world, rho=0.95, eps=1e-7)
    n_total = sum(p.numel() for p in params)
    fusion = torch.zeros(n_total, dtype=torch.bfloat16, device="cuda")   # Horovod fusion buffer
    x_host, y_host = make_pool(seed=100 + rank)
    x_host = x_host.permute(0, 3, 1, 2).contiguous().pin_memory()
    y_host = y_host.pin_memory()

    def step(xb, yb):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(xb)
        loss = F.nll_loss(out.float(), yb)
        loss.backward()
        if world > 1:
            o = 0
            for p in params:                       # pack (Horovod tensor fusion), bf16 on the wire
                fusion[o:o + p.numel()].copy_(p.grad.reshape(-1))
                o += p.numel()
            dist.all_reduce(fusion)
            o = 0
            for p in params:                       # unpack + cast/scale
                p.grad.copy_(fusion[o:o + p.numel()].view_as(p.grad))
                p.grad.mul_(1.0 / world)
                o += p.numel()
        opt.step()
        return loss

    warm = max(3, args.warmup)
    x_dev, y_dev = x_host.cuda(), y_host.cuda()
    b = 0
    for _ in range(warm):
        step(x_dev[b * PER_GPU_BATCH:(b + 1) * PER_GPU_BATCH], y_dev[b * PER_GPU_BATCH:(b + 1) * PER_GPU_BATCH])
        b = (b + 1) % POOL_BATCHES
    sampler = ClockSampler(local)
    _barrier(world)
    if rank == 0:
        sampler.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        step(x_dev[b * PER_GPU_BATCH:(b + 1) * PER_GPU_BATCH], y_dev[b * PER_GPU_BATCH:(b + 1) * PER_GPU_BATCH])
        b = (b + 1) % POOL_BATCHES
    e.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(s.elapsed_time(e), world)
    _barrier(world)
    clocks = sampler.stop() if rank == 0 else None
    # end to end: H2D from pinned memory + loss read every step
    _barrier(world)
    t0 = time.perf_counter()
    for i in range(args.steps):
        bb = (b + i) % POOL_BATCHES
        xb = x_host[bb * PER_GPU_BATCH:(bb + 1) * PER_GPU_BATCH].cuda(non_blocking=True)
        yb = y_host[bb * PER_GPU_BATCH:(bb + 1) * PER_GPU_BATCH].cuda(non_blocking=True)
        loss = step(xb, yb)
        loss_val = loss.item()
    torch.cuda.synchronize()
    e2e_ms = _max_over_ranks((time.perf_counter() - t0) * 1e3, world)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": world * PER_GPU_BATCH * args.steps / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (MNIST-shaped), random-init weights", "impl": "standin",
            "config": {"model": "torch MNIST-CNN, autocast bf16, NCCL all-reduce of a fused bf16 buffer + "
                                "torch.optim.Adadelta (stock-PyTorch rendition of the reference's Horovod path)",
                       "global_batch": world * PER_GPU_BATCH, "parallelism": f"dp{world}"},
            "clocks": clocks,
            "e2e": {"value": world * PER_GPU_BATCH * args.steps / (e2e_ms * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": PER_GPU_BATCH * (784 * 4 + 8), "d2h_bytes_per_step": 4,
                    "final_loss": loss_val},
            "gpu_launches": 0}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "standin"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "standin":
        return run_standin(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
