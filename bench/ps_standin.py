"""Stand-in for BASELINE config 3: a TF-style parameter-server job written with stock PyTorch over NCCL.

What a TF ParameterServerStrategy job does per step (reference: tf_yarn/tensorflow/cluster.py:41-67 sets up the
gRPC servers; TF then RecvTensor-pulls the variables and Apply*/SparseApply* run on the ps task), expressed with the
tools the reference's stack has on an NCCL build:

  worker -> ps : the ids of the embedding rows it needs            (NCCL send)
  ps -> worker : those rows (gathered on the ps) + its dense vars  (NCCL send)
  worker       : forward / backward (cuBLAS, bf16 autocast)
  worker -> ps : row gradients + dense gradients                   (NCCL send)
  ps           : index_add_ of the row gradients + Adagrad / FTRL on the ps rank

Same model, batch and topology as `bench.py --config wide_deep --impl ours`.  The ps ranks serve the workers in a
fixed order every step (lock-step rounds keep the NCCL point-to-point pairing deterministic, and no worker ever
waits for a straggler it would not also wait for in TF's asynchronous mode on an idle box).

Launched by bench.py (plain python): re-executes itself under torch.distributed.run with one rank per cluster task.

TFY_STANDIN_DEVICE=cpu runs the same protocol on host tensors over gloo (no GPU needed), e.g. the N=8 topology:
    TFY_STANDIN_DEVICE=cpu python -m torch.distributed.run --nproc-per-node 8 bench/ps_standin.py --steps 2 \
        --warmup 1 --trainers 6 --ps 2 --gpus 8
(it completes there, so the rc=1 seen once at N=8 on GPUs is specific to the NCCL send/recv path).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)



_STAGE = False   # gloo (more ranks than GPUs: the N=1 topology) cannot send CUDA tensors here: stage through the host


def _send(t, dst):
    import torch.distributed as dist
    dist.send(t.detach().cpu() if _STAGE else t, dst=dst)


def _recv(t, src):
    import torch
    import torch.distributed as dist
    if _STAGE:
        h = torch.empty(t.shape, dtype=t.dtype)
        dist.recv(h, src=src)
        t.copy_(h)
    else:
        dist.recv(t, src=src)


def _bcast(t, src):
    import torch.distributed as dist
    if _STAGE:
        h = t.detach().cpu()
        dist.broadcast(h, src=src)
        t.copy_(h)
    else:
        dist.broadcast(t, src=src)


def run(args) -> int:
    from bench import wide_deep
    n_chief, n_worker, n_ps = wide_deep.topology(args.gpus)
    world = n_chief + n_worker + n_ps
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29650 + args.gpus), os.path.abspath(__file__), "--steps", str(args.steps),
           "--warmup", str(args.warmup), "--trainers", str(n_chief + n_worker), "--ps", str(n_ps), "--gpus",
           str(args.gpus)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    if res.returncode != 0 or not lines:
        tb = [ln for ln in res.stderr.splitlines() if "]:" in ln or "Error" in ln]
        sys.stderr.write("\n".join(tb[:60]) + "\n" + res.stderr[-3000:])
        print(json.dumps({"impl": "standin", "config": "wide_deep", "error": f"rc={res.returncode}"}))
        return 1
    print(lines[-1], flush=True)
    return 0


def main():
    import argparse
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from bench import common, wide_deep as wd
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--trainers", type=int, required=True)
    ap.add_argument("--ps", type=int, required=True)
    ap.add_argument("--gpus", type=int, required=True)
    a = ap.parse_args()
    rank, local, world = common.dist_env()
    cpu_debug = os.environ.get("TFY_STANDIN_DEVICE") == "cpu"      # protocol debugging without GPUs (gloo, host tensors)
    ngpu = 0 if cpu_debug else torch.cuda.device_count()
    if not cpu_debug:
        torch.cuda.set_device(local % ngpu)
    common.quiet_nccl()
    global _STAGE
    _STAGE = world > ngpu
    dist.init_process_group("gloo" if _STAGE else "nccl")
    dev = torch.device("cpu" if cpu_debug else "cuda")
    T, P = a.trainers, a.ps
    B, V, E, NC, NN = wd.BATCH, wd.VOCAB, wd.EMB, wd.N_CAT, wd.N_NUM
    H = list(wd.HIDDEN)
    is_ps = rank >= T
    ps_id = rank - T
    # variable placement (round-robin like TF's default device setter): deep tables, wide tables, dense vars
    deep_owner = [t % P for t in range(NC)]
    wide_owner = [t % P for t in range(NC)]
    dims = [NC * E + NN] + H
    dense_shapes = [(1, NN), (1,)]                                     # wide numeric weight, wide bias
    for i in range(len(H)):
        dense_shapes += [(H[i], dims[i]), (H[i],)]
    dense_shapes += [(1, H[-1]), (1,)]
    dense_owner = [i % P for i in range(len(dense_shapes))]
    dense_sizes = [int(torch.tensor(s).prod()) for s in dense_shapes]

    def owned(owner, p):
        return [i for i, o in enumerate(owner) if o == p]

    steps, warm = a.steps, max(3, a.warmup)
    lr = 0.05
    if is_ps:
        g = torch.Generator(device=dev.type).manual_seed(7)
        deep_t = {t: torch.randn(V, E, device=dev, generator=g) / E ** 0.5 for t in owned(deep_owner, ps_id)}
        deep_acc = {t: torch.full((V, E), 0.1, device=dev) for t in deep_t}
        wide_t = {t: torch.zeros(V, 1, device=dev) for t in owned(wide_owner, ps_id)}
        wide_n = {t: torch.full((V, 1), 0.1, device=dev) for t in wide_t}
        wide_z = {t: torch.zeros(V, 1, device=dev) for t in wide_t}
        mine = owned(dense_owner, ps_id)
        n_dense = sum(dense_sizes[i] for i in mine)
        dense = torch.randn(n_dense, device=dev, generator=g) * 0.05
        dense_acc = torch.full((n_dense,), 0.1, device=dev)
        nd, nw = len(deep_t), len(wide_t)
        while True:                                          # serve rounds until rank 0 says stop
            for w in range(T):
                ids = torch.empty((NC, B), dtype=torch.int64, device=dev)
                _recv(ids, src=w)
                rows_d = torch.stack([deep_t[t][ids[t]] for t in deep_t]) if nd else torch.empty(0, device=dev)
                rows_w = torch.stack([wide_t[t][ids[t]] for t in wide_t]) if nw else torch.empty(0, device=dev)
                if nd:
                    _send(rows_d, dst=w)
                if nw:
                    _send(rows_w, dst=w)
                _send(dense, dst=w)
                if nd:
                    gd = torch.empty_like(rows_d)
                    _recv(gd, src=w)
                    for k, t in enumerate(deep_t):                 # sparse Adagrad on the touched rows
                        deep_acc[t].index_add_(0, ids[t], gd[k] * gd[k])
                        deep_t[t].index_add_(0, ids[t], -lr * gd[k] / (deep_acc[t][ids[t]].sqrt() + 1e-7))
                if nw:
                    gw = torch.empty_like(rows_w)
                    _recv(gw, src=w)
                    for k, t in enumerate(wide_t):                 # sparse FTRL on the touched rows
                        i_ = ids[t]
                        n_old = wide_n[t][i_]
                        n_new = n_old + gw[k] * gw[k]
                        z = wide_z[t][i_] + gw[k] - (n_new.sqrt() - n_old.sqrt()) / lr * wide_t[t][i_]
                        wide_n[t][i_] = n_new
                        wide_z[t][i_] = z
                        wide_t[t][i_] = -z / (n_new.sqrt() / lr)
                gdense = torch.empty_like(dense)
                _recv(gdense, src=w)
                dense_acc.addcmul_(gdense, gdense)
                dense.addcdiv_(gdense, dense_acc.sqrt().add_(1e-7), value=-lr)
            flag = torch.zeros(1, device=dev)
            _bcast(flag, src=0)
            if flag.item() > 0:
                break
        dist.barrier()
        dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------------------- trainer
    from tf_yarn_b200.models import wide_deep as wdm
    batches = wdm.synthetic_batches(B, 64, V, seed=rank, n_cat=NC, n_num=NN)
    if not cpu_debug:
        batches = [({k: v.pin_memory() for k, v in f.items()}, y.pin_memory()) for f, y in batches]
    state = {"i": 0}

    def step(sync_loss=False):
        feats, y = batches[state["i"] % len(batches)]
        state["i"] += 1
        num = feats["numeric"].to(dev, non_blocking=True)
        yb = y.to(dev, non_blocking=True).float()
        raw = torch.stack([feats[f"c{t}"][:, 0] for t in range(NC)]).to(dev, non_blocking=True)
        ids = (raw * 2654435761 % (2 ** 32)) % V
        rows_d, rows_w, dense_parts = {}, {}, {}
        for p in range(P):
            _send(ids, dst=T + p)
            nd, nw = len(owned(deep_owner, p)), len(owned(wide_owner, p))
            if nd:
                rd = torch.empty((nd, B, E), device=dev)
                _recv(rd, src=T + p)
                rows_d[p] = rd.requires_grad_(True)
            if nw:
                rw = torch.empty((nw, B, 1), device=dev)
                _recv(rw, src=T + p)
                rows_w[p] = rw.requires_grad_(True)
            dn = torch.empty(sum(dense_sizes[i] for i in owned(dense_owner, p)), device=dev)
            _recv(dn, src=T + p)
            dense_parts[p] = dn.requires_grad_(True)
        var = {}
        for p in range(P):
            o = 0
            for i in owned(dense_owner, p):
                var[i] = dense_parts[p][o:o + dense_sizes[i]].view(dense_shapes[i])
                o += dense_sizes[i]
        deep_rows = [None] * NC
        wide_rows = [None] * NC
        for p in range(P):
            for k, t in enumerate(owned(deep_owner, p)):
                deep_rows[t] = rows_d[p][k]
            for k, t in enumerate(owned(wide_owner, p)):
                wide_rows[t] = rows_w[p][k]
        with torch.autocast(dev.type, dtype=torch.bfloat16):
            x = torch.cat([num] + deep_rows, dim=1)
            vi = 2
            for _ in H:
                x = F.relu(F.linear(x, var[vi], var[vi + 1]))
                vi += 2
            deep_logit = F.linear(x, var[vi], var[vi + 1])
        wide_logit = F.linear(num, var[0], var[1]) + torch.stack(wide_rows).sum(0)
        loss = F.binary_cross_entropy_with_logits((deep_logit.float() + wide_logit).reshape(-1), yb)
        loss.backward()
        for p in range(P):
            if p in rows_d:
                _send(rows_d[p].grad, dst=T + p)
            if p in rows_w:
                _send(rows_w[p].grad, dst=T + p)
            _send(dense_parts[p].grad, dst=T + p)
        flag = torch.zeros(1, device=dev)
        _bcast(flag, src=0)                      # rank 0 tells the ps ranks when to stop serving
        return loss.item() if sync_loss else loss

    import time
    for _ in range(warm):
        step()
    if cpu_debug:                                        # host clock only: this mode checks the protocol, not speed
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dev_ms = (time.perf_counter() - t0) * 1e3
    else:
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            step()
        e.record()
        e.synchronize()
        dev_ms = s.elapsed_time(e)
    t0 = time.perf_counter()
    last = 0.0
    for _ in range(steps):
        last = step(sync_loss=True)
    if not cpu_debug:
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    # stop the ps ranks: one more round with the flag set
    feats, y = batches[0]
    raw = torch.stack([feats[f"c{t}"][:, 0] for t in range(NC)]).to(dev)
    ids = (raw * 2654435761 % (2 ** 32)) % V
    for p in range(P):
        _send(ids, dst=T + p)
        nd, nw = len(owned(deep_owner, p)), len(owned(wide_owner, p))
        if nd:
            rd = torch.empty((nd, B, E), device=dev); _recv(rd, src=T + p)
        if nw:
            rw = torch.empty((nw, B, 1), device=dev); _recv(rw, src=T + p)
        dn = torch.empty(sum(dense_sizes[i] for i in owned(dense_owner, p)), device=dev); _recv(dn, src=T + p)
        if nd:
            _send(torch.zeros_like(rd), dst=T + p)
        if nw:
            _send(torch.zeros_like(rw), dst=T + p)
        _send(torch.zeros_like(dn), dst=T + p)
    flag = torch.ones(1, device=dev)
    _bcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        value = T * B * steps / (dev_ms * 1e-3)
        print(json.dumps({
            "metric": wd.METRIC, "value": value, "unit": "samples/s", "n_gpus": a.gpus, "steps": steps, "warmup": warm,
            "repeats": 1, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 autocast, fp32 master on the ps", "data": "synthetic (Criteo-shaped), "
            "random-init weights", "impl": "standin",
            "config": {"model": "wide-and-deep, NCCL send/recv parameter server (rows gathered / index_add_ + Adagrad / "
                                "FTRL applied on the ps ranks), eager", "topology": f"{T} trainers + {P} ps",
                       "global_batch": T * B, "per_gpu_batch": B, "parallelism": f"ps {T} trainers / {P} ps (lock-step)",
                       "timing": "rank 0's region (the ps ranks serve the trainers in lock step, so every trainer "
                                 "advances at the same rate)"},
            "clocks": None,
            "e2e": {"value": T * B * steps / (e2e_ms * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": B * (NN * 4 + NC * 8 + 8), "d2h_bytes_per_step": 4, "steps": steps,
                    "ms_per_step": e2e_ms / steps, "final_loss": last},
            "gpu_launches": 0}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
