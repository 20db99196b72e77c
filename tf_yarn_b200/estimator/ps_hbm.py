"""B200 data plane of the asynchronous parameter server: shards in the ps ranks' HBM.

Every task of the training cluster (chief, workers AND ps) joins one symmetric arena
(:mod:`tf_yarn_b200.parallel.symm`, rendezvous through the launcher's KV store).  A ps rank's
shard is a region of ITS arena; chief/workers reach it through the peer mapping over NVLink:

* dense pull   : ``tfy_ps_pull`` -- peer fp32 master -> local replica, ONE launch for the variables that need a
  local copy (biases, small vectors); Dense weights and embedding rows are never copied;
* GEMM pull    : the weight matrix of a Dense layer stays on the ps rank: its row-padded bf16 shadow is streamed
  by TMA into the tcgen05 GEMM that consumes it (forward ``tfy_gemm_bf16``, backward ``tfy_dense_bwd``: dW and
  dx in one kernel against the same remote shadow);
* sparse pull  : ``tfy_ps_embedding_bag`` / ``tfy_ps_multi_bag`` (all tables of a tower in one launch); for the
  first deep layer the gather IS the A-operand producer of the GEMM (K5, ``tfy_ps_gather_gemm``);
* push         : ``tfy_ps_push`` / ``tfy_ps_push_rows`` / ``tfy_ps_multi_push_rows`` -- gradients applied to the peer
  master with vector red/atom and the variable's own optimizer fused (SGD, Adagrad, Adam, FTRL), asynchronously
  and without locks (K6); row gradients are pushed straight out of column slices of the fused layer's dx.

The ps process only owns memory and waits for the stop barrier: no server thread is on the data path.
Control state (layout, readiness, global step) stays on the KV store / a shared-memory header, as in
the CPU data plane (:mod:`tf_yarn_b200.estimator.ps`).
"""
from __future__ import annotations

import ctypes
import logging
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from tf_yarn_b200 import _task_commons
from tf_yarn_b200.estimator import ps as ps_cpu
from tf_yarn_b200.ops import native
from tf_yarn_b200.parallel.symm import _RawCudaMemory

logger = logging.getLogger(__name__)

_vp, _i, _sz, _f, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float, ctypes.c_longlong
native.declare("tfy_ps_pull", [_vp, _i, _sz, _i, _vp])
native.declare("tfy_ps_push", [_vp, _i, _sz, _i, _f, _vp, _vp])
native.declare("tfy_ps_refresh_shadow", [_vp, _i, _sz, _vp])
native.declare("tfy_dense_bwd", [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp])
native.declare("tfy_ps_embedding_bag", [_vp, _vp, _vp, _i, _i, _i, _i, _ll, _i, _vp])
native.declare("tfy_ps_push_rows", [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _i, _i, _f, _f, _f, _f, _f, _f, _vp,
                                    _i, _vp])
native.declare("tfy_ps_multi_bag", [_vp, _vp, _vp, _i, _i, _i, _i, _ll, _vp])
native.declare("tfy_ps_multi_push_rows", [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _i, _f, _f, _f, _f, _f, _f, _vp, _i,
                                          _i, _vp])
native.declare("tfy_ps_gather_gemm", [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _vp])

OPT_CODES = {"sgd": native.OPT_SGD, "adagrad": native.OPT_ADAGRAD, "adam": native.OPT_ADAM,
             "ftrl": native.OPT_FTRL}


class PsSeg(ctypes.Structure):
    """Mirror of TfyPsSeg (ops/csrc/tfy_ps.cu): one variable of a push / pull, with ITS optimizer."""
    _fields_ = [("remote_w", ctypes.c_uint64), ("remote_s1", ctypes.c_uint64), ("remote_s2", ctypes.c_uint64),
                ("remote_shadow", ctypes.c_uint64), ("local", ctypes.c_uint64), ("n", ctypes.c_uint64),
                ("opt", ctypes.c_int32), ("lr", ctypes.c_float), ("eps", ctypes.c_float), ("wd", ctypes.c_float),
                ("p1", ctypes.c_float), ("p2", ctypes.c_float), ("p3", ctypes.c_float), ("pad", ctypes.c_int32),
                ("cols", ctypes.c_uint32), ("shadow_ld", ctypes.c_uint32)]


def cluster_ranks(cluster) -> Dict[str, int]:
    """Global rank of every cluster task in the symmetric arena: chief, workers, then ps."""
    keys = cluster.trainers() + [f"ps:{i}" for i in range(len(cluster.spec.get("ps", [])))]
    return {k: r for r, k in enumerate(keys)}


def join_arena(cluster):
    """Create this process's Communicator over ALL cluster tasks (collective: every task calls it)."""
    from tf_yarn_b200.parallel import runtime
    ranks = cluster_ranks(cluster)
    me = f"{cluster.task_type}:{cluster.task_id}"
    os.environ["TFY_RANK"] = str(ranks[me])
    os.environ["TFY_WORLD_SIZE"] = str(len(ranks))
    ids = [int(v) for v in os.environ.get("TFY_GPU_IDS", "").split(",") if v.strip() != ""]
    dev = ids[0] if ids else ranks[me] % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    return runtime.get_communicator(device=dev), ranks


def shard_bytes(layout: ps_cpu.Layout) -> int:
    """Bytes every rank reserves in its arena (the largest shard; fp32 master+slots, then bf16 shadows)."""
    worst = 0
    for ps in range(layout.n_ps):
        elems = layout.shard_elems[ps]
        rows = sum(int(s[0]) for (_, s), o in zip(layout.variables, layout.owner) if o == ps and len(s) == 2)
        worst = max(worst, elems * 4 + elems * 2 + rows * 16 + 4096)       # fp32 area + (row-padded) bf16 shadows
    return (worst + 4095) // 4096 * 4096


class HbmConnection:
    """Chief / worker view of the HBM shards."""

    def __init__(self, layout: ps_cpu.Layout, comm, ranks: Dict[str, int], region_off: int, names: List[str],
                 header: ps_cpu.ShmShard, network: nn.Module):
        self.layout, self.comm, self.names, self.header = layout, comm, names, header
        self.lib = native.load()
        bad = sorted({k for k in layout.kinds if k not in OPT_CODES})
        if bad:
            raise ValueError(f"the HBM parameter server fuses SGD, Adagrad, Adam and FTRL; got {bad}")
        self.var_opt = [OPT_CODES[k] for k in layout.kinds]
        arena = comm.arena
        self.ps_base = [arena.peer_base[ranks[f"ps:{i}"]] + region_off for i in range(layout.n_ps)]
        self.region_bytes = shard_bytes(layout)
        dev = torch.device(f"cuda:{comm.device}")
        self.device = dev
        params = dict(network.named_parameters())
        self.params = [params[n] for n in names]
        self.sparse: Dict[int, nn.Module] = {}      # variable index -> EmbeddingBag served sparsely
        self.gemm: Dict[int, nn.Module] = {}        # variable index -> Linear whose weight stays remote
        self._classify(network)
        self.dense_idx = [i for i in range(len(names)) if i not in self.sparse]
        # Dense-layer weights served by the remote-weight GEMMs are NEVER copied to the worker: forward streams
        # the ps rank's bf16 shadow through TMA, backward (dW, dx) reads the same shadow (tfy_dense_bwd)
        self.pull_idx = [i for i in self.dense_idx if i not in self.gemm]
        self._pull_segs = self._make_segs(for_push=False, idx=self.pull_idx)
        self._push_segs = None
        self._push_ptrs = None
        self._max_n = max([self._n4(i) for i in self.dense_idx] + [4])
        self.pull_stream = torch.cuda.Stream(device=dev)
        self._pull_done: Optional[torch.cuda.Event] = None
        # Adam bias correction sqrt(1-b2^t)/(1-b1^t): a device scalar the push kernels read, refreshed from the
        # global step before every push (so a captured CUDA graph sees it advance)
        self._adam = [i for i, k in enumerate(layout.kinds) if k == "adam"]
        self.adam_scale = torch.ones(1, dtype=torch.float32, device=dev)
        self._adam_host = torch.ones(1, dtype=torch.float32).pin_memory() if self._adam else None
        # NVLink traffic / kernel launches of one step, counted on the host while a step runs eagerly (the
        # captured graph replays exactly the same launches)
        self._acct = {"pull_bytes": 0, "push_bytes": 0, "launches_per_step": 0}
        self._acct_last = dict(self._acct)

    # ------------------------------------------------------------------ addressing
    def _n4(self, i: int) -> int:
        return (self.layout.numel[i] + 3) // 4 * 4

    def master_ptr(self, i: int, slot: int = 0) -> int:
        lay = self.layout
        return self.ps_base[lay.owner[i]] + 4 * (lay.offset[i] + slot * lay.padded(i))

    def shadow_ptr(self, i: int) -> int:
        """bf16 shadow of variable i: the shadows of a ps follow its fp32 area (master + slots)."""
        lay = self.layout
        ps = lay.owner[i]
        return self.ps_base[ps] + 4 * lay.shard_elems[ps] + 2 * self._shadow_off(i)

    def shadow_ld(self, i: int) -> int:
        """Row pitch (elements) of variable i's bf16 shadow: 2-D weights are stored with rows padded to a multiple
        of 8 elements so that a TMA tensor map can describe them whatever in_features is (0 = flat)."""
        shape = self.layout.variables[i][1]
        return (int(shape[1]) + 7) // 8 * 8 if len(shape) == 2 else 0

    def _shadow_elems(self, i: int) -> int:
        shape = self.layout.variables[i][1]
        ld = self.shadow_ld(i)
        return (int(shape[0]) * ld + 7) // 8 * 8 if ld else self.layout.padded(i)

    def _shadow_off(self, i: int) -> int:
        """Element offset of variable i inside its ps's bf16 shadow area (layout order)."""
        lay = self.layout
        off = 0
        for j in range(i):
            if lay.owner[j] == lay.owner[i]:
                off += self._shadow_elems(j)
        return off

    def remote_tensor(self, i: int, slot: int = 0) -> torch.Tensor:
        """Zero-copy torch view of a peer region (used for initialisation and checkpoints)."""
        n = self.layout.numel[i]
        mem = _RawCudaMemory(self.master_ptr(i, slot), n * 4, self)
        return torch.as_tensor(mem, device=self.device).view(torch.float32)

    def _classify(self, network: nn.Module) -> None:
        by_param = {id(p): i for i, p in enumerate(self.params)}
        for mod in network.modules():
            if isinstance(mod, nn.EmbeddingBag) and id(mod.weight) in by_param and mod.mode in ("sum", "mean"):
                self.sparse[by_param[id(mod.weight)]] = mod
            elif isinstance(mod, nn.Linear) and id(mod.weight) in by_param and mod.out_features >= 8 \
                    and mod.out_features % 8 == 0:
                self.gemm[by_param[id(mod.weight)]] = mod

    def _make_segs(self, for_push: bool, idx: Optional[List[int]] = None):
        idx = self.dense_idx if idx is None else idx
        segs = (PsSeg * max(1, len(idx)))()
        for k, i in enumerate(idx):
            p = self.params[i]
            h = self.layout.hypers[i]
            segs[k].remote_w = self.master_ptr(i)
            segs[k].remote_s1 = self.master_ptr(i, 1) if self.layout.var_slots[i] >= 1 else 0
            segs[k].remote_s2 = self.master_ptr(i, 2) if self.layout.var_slots[i] >= 2 else 0
            segs[k].remote_shadow = self.shadow_ptr(i) if i in self.gemm else 0
            segs[k].local = (p.grad.data_ptr() if for_push else p.data_ptr())
            segs[k].n = self.layout.numel[i]
            segs[k].opt = self.var_opt[i]
            segs[k].lr, segs[k].eps, segs[k].wd = float(h["lr"]), float(h["eps"]), float(h["wd"])
            # FTRL: l1, l2, beta (OptimizerSpec.ftrl keeps beta in eps) | Adam: beta1, beta2
            segs[k].p1, segs[k].p2 = float(h["p1"]), float(h["p2"])
            segs[k].p3 = float(h["eps"]) if self.layout.kinds[i] == "ftrl" else 0.0
            ld = self.shadow_ld(i) if i in self.gemm else 0
            segs[k].cols = int(self.layout.variables[i][1][1]) if ld else 0
            segs[k].shadow_ld = ld
        dev_segs = torch.frombuffer(bytearray(bytes(segs)), dtype=torch.uint8).to(self.device)
        return dev_segs

    # ------------------------------------------------------------------ install the fused ops
    def install(self, network: nn.Module) -> None:
        """Route EmbeddingBag / Linear modules through the peer-memory kernels."""
        conn = self
        for idx, mod in self.sparse.items():
            mod.forward = _make_bag_forward(conn, idx, mod)
        for idx, mod in self.gemm.items():
            mod.forward = _make_linear_forward(conn, idx, mod)
        self.fused_first = _try_fuse_first_layer(conn, network)
        self.fused_wide = _try_fuse_wide_tower(conn, network)

    # ------------------------------------------------------------------ pull / push
    def pull(self, network: nn.Module) -> None:
        """Dense pull of the variables that need a local replica (biases, small vectors): ONE launch of peer
        loads.  Embedding rows and Dense weights are not pulled (gathered / streamed by their consumers)."""
        self._acct = {"pull_bytes": 0, "push_bytes": 0, "launches_per_step": 0}      # a step starts with the pull
        if not self.pull_idx:
            return
        native.check(self.lib.tfy_ps_pull(self._pull_segs.data_ptr(), len(self.pull_idx), self._max_n, 0,
                                          torch.cuda.current_stream().cuda_stream), "tfy_ps_pull")
        self.account(pull=sum(self.layout.numel[i] * 4 for i in self.pull_idx))

    def account(self, pull: int = 0, push: int = 0, launches: int = 1) -> None:
        a = self._acct
        a["pull_bytes"] += pull
        a["push_bytes"] += push
        a["launches_per_step"] += launches

    def traffic_per_step(self) -> dict:
        """Bytes read from / written to the ps ranks over NVLink and kernels launched by the last eager step."""
        return dict(self._acct_last)

    def push(self, network: nn.Module) -> None:
        if not self.dense_idx:
            return
        for i in self.dense_idx:
            p = self.params[i]
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        ptrs = tuple(self.params[i].grad.data_ptr() for i in self.dense_idx)
        if self._push_segs is None or self._push_ptrs != ptrs:     # gradients normally keep their address
            self._push_segs, self._push_ptrs = self._make_segs(for_push=True), ptrs
        segs = self._push_segs
        native.check(self.lib.tfy_ps_push(segs.data_ptr(), len(self.dense_idx), self._max_n, 0, 1.0,
                                          self.adam_scale.data_ptr(), torch.cuda.current_stream().cuda_stream),
                     "tfy_ps_push")
        # per element: gradient applied to master (+ slots read-modify-written) on the peer
        self.account(push=sum(self.layout.numel[i] * 4 * (1 + self.layout.var_slots[i]) for i in self.dense_idx))
        self._acct_last = dict(self._acct)

    def refresh_adam_scale(self) -> None:
        """Bias correction of the asynchronous Adam from the shared global step (host side, outside any graph)."""
        if not self._adam:
            return
        h = self.layout.hypers[self._adam[0]]
        t = max(self.global_step(), 0) + 1
        self._adam_host[0] = (1.0 - h["p2"] ** t) ** 0.5 / (1.0 - h["p1"] ** t)
        self.adam_scale.copy_(self._adam_host, non_blocking=True)

    def refresh_shadows(self) -> None:
        segs = self._make_segs(for_push=False)
        native.check(self.lib.tfy_ps_refresh_shadow(segs.data_ptr(), len(self.dense_idx), self._max_n,
                                                    torch.cuda.current_stream().cuda_stream), "tfy_ps_refresh_shadow")
        torch.cuda.current_stream().synchronize()

    # ------------------------------------------------------------------ global step / checkpoints
    def increment_global_step(self) -> int:
        return self.header.add_global_step(1)

    def global_step(self) -> int:
        return self.header.global_step()

    def state_dict_from_ps(self, network: nn.Module) -> Dict[str, torch.Tensor]:
        torch.cuda.synchronize()
        with torch.no_grad():
            for i, p in enumerate(self.params):
                p.copy_(self.remote_tensor(i).view(p.shape))
        return network.state_dict()


def _make_bag_forward(conn: HbmConnection, idx: int, mod: nn.EmbeddingBag):
    lib = conn.lib
    V, D = mod.num_embeddings, mod.embedding_dim
    mean = 1 if mod.mode == "mean" else 0

    class _Bag(torch.autograd.Function):
        @staticmethod
        def forward(ctx, ids, anchor):
            ids = ids.contiguous().long()
            B, L = ids.shape
            out = torch.empty((B, D), dtype=torch.float32, device=ids.device)
            native.check(lib.tfy_ps_embedding_bag(conn.master_ptr(idx), ids.data_ptr(), out.data_ptr(), 0, B, L, D, V,
                                                  mean, torch.cuda.current_stream().cuda_stream),
                         "tfy_ps_embedding_bag")
            conn.account(pull=B * L * D * 4)
            ctx.ids = ids
            return out

        @staticmethod
        def backward(ctx, dout):
            ids = ctx.ids
            B, L = ids.shape
            dout = dout.contiguous().float()
            lay = conn.layout
            h = lay.hypers[idx]
            s1 = conn.master_ptr(idx, 1) if lay.var_slots[idx] >= 1 else None
            s2 = conn.master_ptr(idx, 2) if lay.var_slots[idx] >= 2 else None
            p3 = float(h["eps"]) if lay.kinds[idx] == "ftrl" else 0.0
            native.check(lib.tfy_ps_push_rows(conn.master_ptr(idx), s1, s2, ids.data_ptr(), dout.data_ptr(), 0, B, L, D,
                                              V, mean, conn.var_opt[idx], float(h["lr"]), float(h["eps"]),
                                              float(h["p1"]), float(h["p2"]), p3, 1.0, conn.adam_scale.data_ptr(), 0,
                                              torch.cuda.current_stream().cuda_stream), "tfy_ps_push_rows")
            conn.account(push=B * L * D * 4 * (1 + lay.var_slots[idx]))
            return None, None

    anchor = torch.zeros((), device=conn.device, requires_grad=True)

    def forward(ids, offsets=None, per_sample_weights=None):
        return _Bag.apply(ids, anchor)
    return forward


def _make_linear_forward(conn: HbmConnection, idx: int, mod: nn.Linear):
    from tf_yarn_b200.ops.gemm import gemm_bf16
    N, K = mod.out_features, mod.in_features
    Kp = conn.shadow_ld(idx)                 # row pitch of the remote bf16 shadow (K rounded up to 8; pad columns = 0)
    shadow = conn.shadow_ptr(idx)

    class _PSLinear(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, weight, bias):
            if Kp == K:
                xb = x.to(torch.bfloat16).contiguous()
            else:                            # TMA needs 16-byte row pitches: zero-pad the activations like the shadow
                xb = torch.zeros((x.shape[0], Kp), dtype=torch.bfloat16, device=x.device)
                xb[:, :K] = x
            y = gemm_bf16(xb, None, bias=bias.to(torch.bfloat16) if bias is not None else None, b_ptr=shadow,
                          b_rows=N, b_ld=Kp, impl="1cta")       # weights stream from the ps rank over NVLink
            conn.account(pull=N * Kp * 2)
            ctx.save_for_backward(xb, weight)
            ctx.has_bias = bias is not None
            return y.to(x.dtype)

        @staticmethod
        def backward(ctx, dy):
            xb, weight = ctx.saved_tensors
            B = xb.shape[0]
            dyb = dy.to(torch.bfloat16).contiguous()
            dw = torch.empty((N, Kp), dtype=torch.bfloat16, device=dy.device)
            dx = torch.empty((B, Kp), dtype=torch.bfloat16, device=dy.device) if ctx.needs_input_grad[0] else None
            # dW = dy^T x and dx = dy W in one tcgen05 kernel; W is the ps rank's bf16 shadow, read over NVLink
            native.check(conn.lib.tfy_dense_bwd(dyb.data_ptr(), xb.data_ptr(), shadow, dw.data_ptr(),
                                                dx.data_ptr() if dx is not None else None, B, N, Kp,
                                                torch.cuda.current_stream().cuda_stream), "tfy_dense_bwd")
            conn.account(pull=N * Kp * 2 if dx is not None else 0)
            db = dy.float().sum(0) if ctx.has_bias else None
            gx = dx[:, :K].to(dy.dtype) if dx is not None else None
            return gx, dw[:, :K].to(weight.dtype), db

    def forward(x):
        return _PSLinear.apply(x, mod.weight, mod.bias)
    return forward


class _TableGroup:
    """T embedding tables of one tower served by the multi-table kernels (one launch for the whole group)."""

    def __init__(self, conn: HbmConnection, tables_idx: List[int], cat_cols):
        lay = conn.layout
        self.conn, self.idx, self.cols = conn, tables_idx, cat_cols
        first = tables_idx[0]
        self.uniform = all(lay.kinds[i] == lay.kinds[first] and lay.hypers[i] == lay.hypers[first] for i in tables_idx)
        dev = conn.device
        self.tables = torch.tensor([conn.master_ptr(i) for i in tables_idx], dtype=torch.int64, device=dev)
        ns = lay.var_slots[first]
        self.s1 = torch.tensor([conn.master_ptr(i, 1) for i in tables_idx], dtype=torch.int64, device=dev) if ns >= 1 else None
        self.s2 = torch.tensor([conn.master_ptr(i, 2) for i in tables_idx], dtype=torch.int64, device=dev) if ns >= 2 else None
        self.V = cat_cols[0].num_buckets
        self.same_v = all(c.num_buckets == self.V and c.hashed == cat_cols[0].hashed for c in cat_cols)
        h = lay.hypers[first]
        self.h = h
        self.opt = conn.var_opt[first]
        self.p3 = float(h["eps"]) if lay.kinds[first] == "ftrl" else 0.0
        self.slots = ns

    def ids_of(self, features) -> torch.Tensor:
        """int64 [T, B] bucket ids of the batch: ONE hash over the stacked columns (26 separate 3-kernel hashes
        were 40 % of the step's device time in the first profile)."""
        raw = torch.stack([features[c.key].reshape(-1) for c in self.cols]).long()
        if self.cols[0].hashed:
            raw = (raw * 2654435761) % (2 ** 32)
        return (raw % self.V).contiguous()

    def push(self, ids: torch.Tensor, dout: torch.Tensor, D: int, dout_ld: int, col_stride: int) -> None:
        conn, h = self.conn, self.h
        B, T = ids.shape[1], len(self.idx)
        native.check(conn.lib.tfy_ps_multi_push_rows(
            self.tables.data_ptr(), self.s1.data_ptr() if self.s1 is not None else None,
            self.s2.data_ptr() if self.s2 is not None else None, ids.data_ptr(), dout.data_ptr(),
            1 if dout.dtype == torch.bfloat16 else 0, B, T, D, self.V, self.opt, float(h["lr"]), float(h["eps"]),
            float(h["p1"]), float(h["p2"]), self.p3, 1.0, conn.adam_scale.data_ptr(), dout_ld, col_stride,
            torch.cuda.current_stream().cuda_stream), "tfy_ps_multi_push_rows")
        conn.account(push=T * B * D * 4 * (1 + self.slots))


def _try_fuse_wide_tower(conn: HbmConnection, network: nn.Module) -> bool:
    """The wide tower (fc.LinearModel): T per-bucket weight tables [V, units] looked up once per example and summed.
    One multi-table gather kernel forward, one multi-table push (FTRL / ... fused) backward, instead of T launches
    each plus T adds."""
    from tf_yarn_b200.estimator import feature_column as fc
    if os.environ.get("TFY_PS_FUSE_WIDE", "1") == "0":
        return False
    lin = getattr(network, "linear", None)
    if not isinstance(lin, fc.LinearModel) or not lin.tables:
        return False
    by_param = {id(p): i for i, p in enumerate(conn.params)}
    cat_cols, tables_idx = [], []
    for c in lin.columns:
        cc = c.categorical_column if isinstance(c, (fc.EmbeddingColumn, fc.IndicatorColumn)) else c
        if isinstance(cc, fc.CategoricalColumn):
            emb = lin.tables[cc.key]
            ti = by_param.get(id(emb.weight))
            if ti is None or ti not in conn.sparse or emb.mode != "sum":
                return False
            cat_cols.append(cc)
            tables_idx.append(ti)
    group = _TableGroup(conn, tables_idx, cat_cols)
    if not (group.uniform and group.same_v):
        return False
    units = lin.units
    dev = conn.device

    class _MultiBag(torch.autograd.Function):
        @staticmethod
        def forward(ctx, ids, anchor):
            B = ids.shape[1]
            out = torch.empty((B, units), dtype=torch.float32, device=dev)
            native.check(conn.lib.tfy_ps_multi_bag(group.tables.data_ptr(), ids.data_ptr(), out.data_ptr(), 0, B,
                                                   len(tables_idx), units, group.V,
                                                   torch.cuda.current_stream().cuda_stream), "tfy_ps_multi_bag")
            conn.account(pull=B * len(tables_idx) * units * 4)
            ctx.ids = ids
            return out

        @staticmethod
        def backward(ctx, dout):
            group.push(ctx.ids, dout.contiguous().float(), units, units, 0)     # every table gets the same gradient
            return None, None

    anchor = torch.zeros((), device=dev, requires_grad=True)
    num_cols = [c for c in lin.columns if isinstance(c, fc.NumericColumn)]

    def forward(features):
        out = _MultiBag.apply(group.ids_of(features), anchor)
        if lin.numeric is not None and num_cols:
            nums = [features[c.key].reshape(features[c.key].shape[0], -1).float() for c in num_cols]
            x = nums[0] if len(nums) == 1 else torch.cat(nums, dim=1)
            out = out + lin.numeric(x.to(lin.numeric.weight.dtype))
        return out + lin.bias

    lin.forward = forward
    logger.info("wide tower: %d tables served by the multi-table gather / push kernels", len(tables_idx))
    return True


def _try_fuse_first_layer(conn: HbmConnection, network: nn.Module) -> bool:
    """K5: fuse the sparse pull of the embedding rows INTO the first deep GEMM (ops/csrc/tfy_ps_gemm.cu).

    Pattern (the canned wide-and-deep / DNN networks): ``network.dense_features`` concatenates embedding columns
    of dimension 64 with one id per example followed by numeric columns (<= 64 values in total), and
    ``network.hidden[0]`` is a Linear served by the remote-weight GEMM.  The gather of the rows from the ps ranks'
    HBM becomes the A-operand producer of that GEMM; the backward (dW, dx) is one tcgen05 kernel against the same
    remote shadow, and the row gradients are pushed straight out of its dx.  TFY_PS_FUSE_FIRST=0 disables it."""
    from tf_yarn_b200.estimator import feature_column as fc
    if os.environ.get("TFY_PS_FUSE_FIRST", "1") == "0":
        return False
    df, hidden = getattr(network, "dense_features", None), getattr(network, "hidden", None)
    if not isinstance(df, fc.DenseFeatures) or not hidden or not isinstance(hidden[0], nn.Linear):
        return False
    lin = hidden[0]
    by_param = {id(p): i for i, p in enumerate(conn.params)}
    widx = by_param.get(id(lin.weight))
    if widx is None or widx not in conn.gemm or lin.bias is None:
        return False
    emb_cols, num_cols, seen_numeric = [], [], False
    for c in df.columns:
        if isinstance(c, fc.EmbeddingColumn) and not seen_numeric and c.dimension == 64:
            emb_cols.append(c)
        elif isinstance(c, fc.NumericColumn):
            seen_numeric = True
            num_cols.append(c)
        else:
            return False                     # embeddings must come first, all of dimension 64
    T, n_num = len(emb_cols), sum(c.dim for c in num_cols)
    if T == 0 or n_num > 64 or lin.in_features != 64 * T + n_num:
        return False
    tables_idx = []
    for c in emb_cols:
        emb = df.embeddings[c.categorical_column.key]
        ti = by_param.get(id(emb.weight))
        if ti is None or ti not in conn.sparse or emb.mode not in ("sum", "mean"):
            return False
        tables_idx.append(ti)
    V = emb_cols[0].categorical_column.num_buckets
    if any(c.categorical_column.num_buckets != V for c in emb_cols):
        return False
    dev = conn.device
    N, K, Kp = lin.out_features, lin.in_features, conn.shadow_ld(widx)
    shadow = conn.shadow_ptr(widx)
    group = _TableGroup(conn, tables_idx, [c.categorical_column for c in emb_cols])
    table_ptrs = group.tables
    lib, lay = conn.lib, conn.layout

    class _FusedFirst(torch.autograd.Function):
        @staticmethod
        def forward(ctx, ids, numeric, weight, bias):
            B = ids.shape[1]
            xbuf = torch.zeros((B, Kp), dtype=torch.bfloat16, device=dev)
            y = torch.empty((B, N), dtype=torch.bfloat16, device=dev)
            bb = bias.to(torch.bfloat16)
            native.check(lib.tfy_ps_gather_gemm(shadow, table_ptrs.data_ptr(), ids.data_ptr(),
                                                numeric.data_ptr() if numeric is not None else None, bb.data_ptr(),
                                                xbuf.data_ptr(), y.data_ptr(), B, N, T, n_num, Kp, V,
                                                torch.cuda.current_stream().cuda_stream), "tfy_ps_gather_gemm")
            conn.account(pull=B * T * 64 * 4 + N * Kp * 2)
            ctx.save_for_backward(xbuf, ids, weight)
            return y.to(weight.dtype)

        @staticmethod
        def backward(ctx, dy):
            xbuf, ids, weight = ctx.saved_tensors
            B = xbuf.shape[0]
            s = torch.cuda.current_stream().cuda_stream
            dyb = dy.to(torch.bfloat16).contiguous()
            dw = torch.empty((N, Kp), dtype=torch.bfloat16, device=dev)
            dx = torch.empty((B, Kp), dtype=torch.bfloat16, device=dev)
            native.check(lib.tfy_dense_bwd(dyb.data_ptr(), xbuf.data_ptr(), shadow, dw.data_ptr(), dx.data_ptr(), B, N,
                                           Kp, s), "tfy_dense_bwd")
            conn.account(pull=N * Kp * 2)
            if group.uniform:
                # table t takes columns [64 t, 64 t + 64) of dx: ONE launch for all tables, straight out of dx
                group.push(ids, dx, 64, Kp, 64)
                return None, None, dw[:, :K].to(weight.dtype), dy.float().sum(0).to(weight.dtype)
            for t, ti in enumerate(tables_idx):          # (mixed optimizers) table by table
                h = lay.hypers[ti]
                s1 = conn.master_ptr(ti, 1) if lay.var_slots[ti] >= 1 else None
                s2 = conn.master_ptr(ti, 2) if lay.var_slots[ti] >= 2 else None
                p3 = float(h["eps"]) if lay.kinds[ti] == "ftrl" else 0.0
                native.check(lib.tfy_ps_push_rows(
                    conn.master_ptr(ti), s1, s2, ids[t].data_ptr(), dx.data_ptr() + 2 * 64 * t, 1, B, 1, 64, V, 0,
                    conn.var_opt[ti], float(h["lr"]), float(h["eps"]), float(h["p1"]), float(h["p2"]), p3, 1.0,
                    conn.adam_scale.data_ptr(), Kp, s), "tfy_ps_push_rows")
                conn.account(push=B * 64 * 4 * (1 + lay.var_slots[ti]))
            return None, None, dw[:, :K].to(weight.dtype), dy.float().sum(0).to(weight.dtype)

    class _Deferred:
        """What the patched DenseFeatures returns: the raw inputs of the fused first layer."""

        def __init__(self, ids, numeric):
            self.ids, self.numeric = ids, numeric

        def to(self, *a, **k):
            return self

    def df_forward(features):
        if group.same_v:
            ids = group.ids_of(features)
        else:
            ids = torch.stack([fc._ids(c.categorical_column, features[c.categorical_column.key]).reshape(-1)
                               for c in emb_cols]).contiguous()
        numeric = None
        if num_cols:
            parts = [features[c.key].reshape(features[c.key].shape[0], -1).float() for c in num_cols]
            numeric = (parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)).contiguous()
        return _Deferred(ids, numeric)

    def lin_forward(x):
        if isinstance(x, _Deferred):
            return _FusedFirst.apply(x.ids, x.numeric, lin.weight, lin.bias)
        return _make_linear_forward(conn, widx, lin)(x)

    df.forward = df_forward
    lin.forward = lin_forward
    logger.info("K5: embedding gather of %d tables fused into the first deep GEMM (%d -> %d)", T, K, N)
    return True


# ---------------------------------------------------------------------------------------------
# connection set-up (collective over the cluster)
# ---------------------------------------------------------------------------------------------
def connect_worker(network: nn.Module, opt_desc, cluster, is_chief: bool, global_step: int,
                   opt_by_name=None) -> HbmConnection:
    client = _task_commons.TaskClient.from_current()
    kv = client.kv
    n_ps = len(cluster.spec["ps"])
    named = ps_cpu._named_trainables(network)
    names = [n for n, _ in named]
    if is_chief:
        layout = ps_cpu.make_layout(named, n_ps, opt_desc, opt_by_name)
        kv[ps_cpu.KV_LAYOUT] = layout.to_json().encode()
    else:
        layout = ps_cpu.Layout.from_json(kv.wait(ps_cpu.KV_LAYOUT))
    comm, ranks = join_arena(cluster)
    region_off = comm.arena.alloc(shard_bytes(layout), align=4096)
    header_path = kv.wait("ps:0/shard").decode()
    header = ps_cpu.ShmShard(header_path, 8, create=False)
    conn = HbmConnection(layout, comm, ranks, region_off, names, header, network)
    if is_chief:
        with torch.no_grad():
            for i, (_, p) in enumerate(named):
                conn.remote_tensor(i).copy_(p.detach().reshape(-1).float())
                if layout.kinds[i] in ("adagrad", "ftrl"):
                    conn.remote_tensor(i, 1).fill_(layout.hypers[i]["init_s1"])
        conn.refresh_shadows()
        header.set_global_step(global_step)
        torch.cuda.synchronize()
        kv[ps_cpu.KV_READY] = b"1"
        logger.info("HBM parameter servers initialised: %d variables on %d ps (%d sparse, %d remote-GEMM)",
                    len(names), n_ps, len(conn.sparse), len(conn.gemm))
    else:
        kv.wait(ps_cpu.KV_READY)
    conn.install(network)
    return conn


def serve(cluster, poll_secs: float = 0.2) -> None:
    """``ps`` task on B200: join the arena, reserve the shard region, publish the header, idle."""
    import time
    client = _task_commons.TaskClient.from_current()
    kv = client.kv
    idx = cluster.task_id
    layout = ps_cpu.Layout.from_json(kv.wait(ps_cpu.KV_LAYOUT))
    comm, _ = join_arena(cluster)
    off = comm.arena.alloc(shard_bytes(layout), align=4096)
    region = comm.arena.tensor(off, (shard_bytes(layout),), torch.uint8)
    region.zero_()
    torch.cuda.synchronize()
    path = os.path.join(ps_cpu._shm_dir(), f"tfy_ps_{ps_cpu._job_tag()}_{idx}")
    header = ps_cpu.ShmShard(path, 8, create=True)
    kv[f"ps:{idx}/shard"] = path.encode()
    logger.info("ps %d: %d bytes of HBM shard at arena offset %d", idx, shard_bytes(layout), off)
    import atexit
    atexit.register(header.unlink)
    while True:
        time.sleep(poll_secs)
