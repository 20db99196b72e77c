"""Train-step engines behind ``Model.fit``.

``GraphTrainEngine`` (B200) — the flagship path of the Horovod-style Keras
configuration:

* parameters live in ONE flat bf16 buffer in the symmetric arena, gradients in
  another; model parameters / ``.grad`` are views into them; fp32 master
  weights and optimizer state are sharded 1/world per rank;
* one training step = forward + backward (cuDNN/cuBLAS bf16 through torch
  autograd) followed by ONE hand-written kernel that does the cross-GPU
  gradient reduction, the bf16->fp32 cast + 1/world scale, the optimizer update
  on the owned shard and the all-gather of the new bf16 parameters
  (``tfy_fused_step_kernel``: multimem.ld_reduce / multimem.st over NVSwitch);
* the whole step is captured in a CUDA graph and replayed; inputs are staged
  from pinned host memory on a copy stream one step ahead; the loss is read
  back asynchronously into a pinned ring.

``EagerTrainEngine`` (CPU plumbing configuration) — plain torch ops, gloo
all-reduce of the gradients, ``torch.optim`` with the same update formulas.
"""
from __future__ import annotations

import ctypes

import logging
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

logger = logging.getLogger(__name__)


def _to_device_batch(t, device, non_blocking=True):
    if isinstance(t, (tuple, list)):
        return type(t)(_to_device_batch(x, device, non_blocking) for x in t)
    if isinstance(t, dict):
        return {k: _to_device_batch(v, device, non_blocking) for k, v in t.items()}
    return t.to(device, non_blocking=non_blocking)


class EagerTrainEngine:
    """Reference-semantics engine for CPU (and the fp32 numerics oracle for the graph engine)."""

    def __init__(self, net: nn.Module, loss_fn: Callable, optimizer, metric_fns: Sequence[Tuple[str, Callable]],
                 device: torch.device, distributed: bool = False):
        self.net, self.loss_fn, self.metric_fns = net, loss_fn, list(metric_fns)
        self.device = device
        self.distributed = distributed
        self.opt_desc = optimizer
        self.opt = optimizer.to_torch([p for p in net.parameters() if p.requires_grad])
        self.kernel_launches = 0

    # -- distributed helpers (gloo) -------------------------------------------
    def _world(self) -> int:
        import torch.distributed as dist
        return dist.get_world_size() if (self.distributed and dist.is_initialized()) else 1

    def broadcast_variables(self, root: int = 0) -> None:
        import torch.distributed as dist
        if self._world() > 1:
            for t in list(self.net.parameters()) + list(self.net.buffers()):
                dist.broadcast(t.data, src=root)

    def train_step(self, x, y) -> Dict[str, torch.Tensor]:
        import torch.distributed as dist
        self.net.train()
        x, y = _to_device_batch(x, self.device), _to_device_batch(y, self.device)
        self.opt.zero_grad(set_to_none=False)
        out = self.net(x)
        loss = self.loss_fn(y, out)
        loss.backward()
        world = self._world()
        if world > 1:
            grads = [p.grad for p in self.net.parameters() if p.grad is not None]
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat)
            flat /= world
            o = 0
            for g in grads:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()
        self.opt.step()
        logs = {"loss": loss.detach()}
        with torch.no_grad():
            for name, fn in self.metric_fns:
                num, den = fn(y, out)
                logs[name] = num / den
        return logs

    def set_learning_rate(self, lr: float) -> None:
        for g in self.opt.param_groups:
            g["lr"] = lr

    def get_learning_rate(self) -> float:
        return self.opt.param_groups[0]["lr"]

    def sync_params_to_module(self) -> None:
        pass

    def state_dict(self) -> dict:
        return {"kind": "eager", "optimizer": self.opt.state_dict()}

    def load_state_dict(self, state: dict) -> None:
        if state.get("kind") == "eager":
            self.opt.load_state_dict(state["optimizer"])

    # fit() drives the eager engine synchronously
    pipelined = False


class GraphTrainEngine:
    """CUDA-graph train step with the fused reduce-scatter/optimizer/all-gather kernel."""

    pipelined = True

    def __init__(self, net: nn.Module, loss_fn: Callable, optimizer, metric_fns: Sequence[Tuple[str, Callable]],
                 device: torch.device, distributed: bool, example_x, example_y,
                 compute_dtype: torch.dtype = torch.bfloat16, use_graph: bool = True, comm=None):
        from tf_yarn_b200.parallel import runtime
        from tf_yarn_b200.parallel.comm import FusedShardedOptimizer
        self.net, self.loss_fn, self.metric_fns = net, loss_fn, list(metric_fns)
        self.device = device
        self.compute_dtype = compute_dtype
        self.use_graph = use_graph
        self.opt_desc = optimizer
        torch.cuda.set_device(device)
        if comm is None:
            if distributed:
                comm = runtime.get_communicator(device=device.index)
            else:
                from tf_yarn_b200.parallel.comm import Communicator
                from tf_yarn_b200.parallel.symm import SoloRendezvous
                comm = _solo_communicator(device.index)
        self.comm = comm
        self.distributed = distributed and comm.world > 1
        self.params = [p for p in net.parameters() if p.requires_grad]
        # fp32 source of truth, in the byte order of the flat buffers: conv kernels are stored
        # (O, kh, kw, I) so that the parameter seen by cuDNN is a channels_last view of its slice
        init = [p.detach().float().permute(0, 2, 3, 1).contiguous() if p.dim() == 4 else p.detach().float()
                for p in self.params]
        self.fused = FusedShardedOptimizer(comm, [tuple(t.shape) for t in init], optimizer.to_spec(),
                                           param_dtype=compute_dtype, grad_dtype=compute_dtype, zero_grads=True)
        self.fused.init_from(init, broadcast_root=None)
        net.to(compute_dtype)
        # re-point the module's parameters / grads at the flat symmetric buffers
        for p, pv, gv in zip(self.params, self.fused.param_views, self.fused.grad_views):
            if p.dim() == 4:
                pv, gv = pv.permute(0, 3, 1, 2), gv.permute(0, 3, 1, 2)
            p.data, p.grad = pv, gv
        self.stream = torch.cuda.Stream(device=device)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._static_x = None
        self._static_y = None
        self._staging = None            # two device-side input slots filled by the copy stream
        self._slot = 0
        self._slot_free = [None, None]  # event: the step that read staging[slot] has copied it out
        self._loss = torch.zeros((), dtype=torch.float32, device=device)
        self._metric_acc = torch.zeros(max(1, len(self.metric_fns)), 2, dtype=torch.float32, device=device)
        self._example = (example_x, example_y)
        self.kernel_launches = 0
        self._launches_per_step = 1          # the fused K4 kernel
        self._captured = False

    # ------------------------------------------------------------------ variables
    def broadcast_variables(self, root: int = 0) -> None:
        """Make every rank start from rank ``root``'s weights (BroadcastGlobalVariablesCallback)."""
        if not self.distributed:
            return
        full = self._full_master()
        self.fused.load_full_master(full, broadcast_root=root)

    def _full_master(self) -> torch.Tensor:
        return self.fused.gather_state()["master"] if self.comm.world > 1 else self.fused.master.clone()

    def master_tensors(self) -> List[torch.Tensor]:
        """Full-precision parameters in module layout (for checkpoints / evaluation)."""
        full = self._full_master()
        out = []
        for p, t in zip(self.params, self.fused.unflatten(full)):
            out.append(t.permute(0, 3, 1, 2) if p.dim() == 4 else t)
        return out

    def sync_params_to_module(self) -> None:
        pass  # module parameters ARE the flat buffer

    # ------------------------------------------------------------------ step body
    def _forward_backward(self, x, y) -> None:
        xin = x
        if torch.is_tensor(xin) and torch.is_floating_point(xin):
            xin = xin.to(self.compute_dtype)
        out = self.net(xin)
        loss = self.loss_fn(y, out)
        loss.backward()
        self._loss.copy_(loss.detach())
        if self.metric_fns:
            with torch.no_grad():
                for i, (_, fn) in enumerate(self.metric_fns):
                    num, den = fn(y, out)
                    self._metric_acc[i, 0] += num
                    self._metric_acc[i, 1] += den
        self.fused.step()

    def _capture(self, x, y) -> None:
        self.net.train()
        # two input slots, and one captured graph per slot reading it in place: the copy stream fills slot
        # (i+1) % 2 while graph[i % 2] runs, and no staging -> static copy sits in front of every step
        self._staging = [(_clone_struct(x), _clone_struct(y)) for _ in range(2)]
        self._static_x, self._static_y = self._staging[0]
        self._loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._loss_in_host = [False, False]
        self._capture_slot = None
        self._ready_ev = [torch.cuda.Event() for _ in range(2)]
        self._done_ev = [torch.cuda.Event() for _ in range(2)]
        from tf_yarn_b200.ops import native
        native.declare("tfy_memcpy_async", [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p])
        self._lib = native.load()
        self._fast_copy = hasattr(self._lib, "tfy_memcpy_async")
        fused = self.fused
        snap = (fused.master.clone(), fused.s1.clone(), fused.s2.clone(), fused.flat_params.clone(),
                fused.hyper.clone(), self._metric_acc.clone())
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(3):
                self._forward_backward(self._static_x, self._static_y)
        self.stream.synchronize()
        self.graphs = [None, None]
        if self.use_graph:
            # a CUDAGraph of an older engine being garbage-collected while this stream captures would
            # invalidate the capture (cudaGraphExecDestroy is not allowed then): collect now, pause GC
            import gc
            gc.collect()
            was_enabled = gc.isenabled()
            gc.disable()
            try:
                for slot in range(2):
                    g = torch.cuda.CUDAGraph()
                    sx, sy = self._staging[slot]
                    # the graphs never run concurrently (same stream), so they share one memory pool
                    pool = self.graphs[0].pool() if slot else None
                    self._capture_slot = slot
                    with torch.cuda.graph(g, stream=self.stream, pool=pool):
                        self._forward_backward(sx, sy)
                        if not self._loss_in_host[slot]:      # (the fast path's head kernel writes it itself)
                            self._loss_host[slot].copy_(self._loss, non_blocking=True)     # D2H node of the graph
                    self._capture_slot = None
                    self.graphs[slot] = g
                self.graph = self.graphs[0]
            finally:
                if was_enabled:
                    gc.enable()
        # undo the warm-up steps: training starts from the user's initial state
        with torch.cuda.stream(self.stream):
            fused.master.copy_(snap[0]); fused.s1.copy_(snap[1]); fused.s2.copy_(snap[2])
            fused.flat_params.copy_(snap[3]); fused.hyper.copy_(snap[4]); self._metric_acc.copy_(snap[5])
            fused.flat_grads.zero_()
        self.stream.synchronize()
        if self.distributed:
            self.comm.barrier()
            torch.cuda.synchronize()
        self._captured = True

    # ------------------------------------------------------------------ public
    def ensure_captured(self, x, y) -> None:
        if not self._captured:
            self._capture(_to_device_batch(x, self.device, False), _to_device_batch(y, self.device, False))

    def stage_inputs(self, x, y):
        """Asynchronous copy (H2D from pinned memory, or D2D) of the next batch into a staging slot.

        Runs on the copy stream, so it overlaps the step that is currently executing.  Returns a
        ticket for :meth:`launch_step`.
        """
        self.ensure_captured(x, y)
        slot = self._slot
        self._slot ^= 1
        sx, sy = self._staging[slot]
        cs = self.copy_stream
        free = self._slot_free[slot]
        if free is not None:
            cs.wait_event(free)
        if (self._fast_copy and torch.is_tensor(x) and torch.is_tensor(y) and x.dtype == sx.dtype
                and y.dtype == sy.dtype and x.is_contiguous() and y.is_contiguous()
                and x.numel() == sx.numel() and y.numel() == sy.numel()):
            # hot path: two cudaMemcpyAsync calls through ctypes, no stream context manager
            lib, sp = self._lib, cs.cuda_stream
            lib.tfy_memcpy_async(sx.data_ptr(), x.data_ptr(), x.numel() * x.element_size(), sp)
            lib.tfy_memcpy_async(sy.data_ptr(), y.data_ptr(), y.numel() * y.element_size(), sp)
        else:
            with torch.cuda.stream(cs):
                _copy_struct(sx, x)
                _copy_struct(sy, y)
        ev = self._ready_ev[slot]
        ev.record(cs)
        return slot, ev

    def launch_step(self, ticket) -> torch.cuda.Event:
        """Run one captured step on the staged batch; returns the event marking its completion.

        The step's loss is copied to ``loss_host(slot)`` (pinned) by the graph itself."""
        slot, ready = ticket
        st = self.stream
        st.wait_event(ready)
        if torch.cuda.current_stream() == st:
            self._launch_on_current(slot)
        else:
            with torch.cuda.stream(st):
                self._launch_on_current(slot)
        done = self._done_ev[slot]
        done.record(st)
        self._slot_free[slot] = done          # the slot may be refilled once this step has run
        self.kernel_launches += self._launches_per_step
        return done

    def _launch_on_current(self, slot: int) -> None:
        if self.graph is not None:
            self.graphs[slot].replay()
        else:
            sx, sy = self._staging[slot]
            self._forward_backward(sx, sy)
            self._loss_host[slot].copy_(self._loss, non_blocking=True)

    def loss_host(self, slot: int) -> torch.Tensor:
        """Pinned scalar that receives the loss of the steps run on ``slot`` (valid once their event fired)."""
        return self._loss_host[slot]

    def read_loss_async(self, pinned_slot: torch.Tensor) -> None:
        with torch.cuda.stream(self.stream):
            pinned_slot.copy_(self._loss, non_blocking=True)

    def train_step(self, x, y) -> Dict[str, torch.Tensor]:
        """Synchronous convenience wrapper (tests, smoke): one step, returns device scalars."""
        ticket = self.stage_inputs(x, y)
        self.launch_step(ticket)
        self.stream.synchronize()
        return {"loss": self._loss.clone()}

    def pop_metrics(self) -> Dict[str, float]:
        self.stream.synchronize()
        acc = self._metric_acc.cpu()
        self._metric_acc.zero_()
        return {name: float(acc[i, 0] / acc[i, 1].clamp_min(1)) for i, (name, _) in enumerate(self.metric_fns)}

    def set_learning_rate(self, lr: float) -> None:
        with torch.cuda.stream(self.stream):
            self.fused.set_lr(lr)

    def get_learning_rate(self) -> float:
        return float(self.fused._hyper_host.lr)

    def state_dict(self) -> dict:
        st = self.fused.gather_state()
        return {"kind": "fused", **{k: (v.cpu() if torch.is_tensor(v) else v) for k, v in st.items()}}

    def load_state_dict(self, state: dict) -> None:
        if state.get("kind") == "fused":
            dev = self.fused.master.device
            self.fused.load_state({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in state.items()
                                   if k != "kind"})


_solo = {}


def _solo_communicator(device_index: int):
    """Single-GPU communicator (world=1): the fused kernel runs in LOCAL mode, no peers."""
    from tf_yarn_b200.parallel.comm import Communicator
    from tf_yarn_b200.parallel.symm import SoloRendezvous
    if device_index not in _solo:
        _solo[device_index] = Communicator(arena_bytes=1 << 30, fusion_bytes=16 << 20, rdv=SoloRendezvous(),
                                           device=device_index)
    return _solo[device_index]


def _clone_struct(t):
    if isinstance(t, (tuple, list)):
        return type(t)(_clone_struct(x) for x in t)
    if isinstance(t, dict):
        return {k: _clone_struct(v) for k, v in t.items()}
    return t.clone()


def _copy_struct(dst, src) -> None:
    if isinstance(dst, (tuple, list)):
        for d, s in zip(dst, src):
            _copy_struct(d, s)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_struct(dst[k], src[k])
    else:
        dst.copy_(src, non_blocking=True)
