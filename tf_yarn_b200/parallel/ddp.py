"""Data-parallel wrapper driven by the hand-written NVLS / P2P all-reduce kernels.

Drop-in for ``torch.nn.parallel.DistributedDataParallel`` as the reference's
PyTorch worker uses it (reference: tf_yarn/pytorch/tasks/worker.py:105-107,
arguments tf_yarn/pytorch/experiment.py:23-27):

* parameters are bucketed in reverse registration order (first bucket 1 MiB,
  then ``bucket_cap_mb``) into flat gradient buffers that live in the
  symmetric arena; every ``p.grad`` is a view into its bucket;
* a post-accumulate-grad hook counts ready gradients; when a bucket is
  complete its all-reduce kernel (``multimem.ld_reduce`` + ``multimem.st``,
  averaged, in place) is launched on a side stream, overlapping the rest of
  backward;
* at the end of backward the compute stream waits for the side stream;
* parameters (and, each forward, module buffers when ``broadcast_buffers``)
  are broadcast from rank 0 with the K7 kernel.

No NCCL call is issued by this wrapper.  On a CPU-only process group (gloo,
the plumbing configuration) :func:`wrap_model` returns torch's own DDP.
"""
from __future__ import annotations

import contextlib
from typing import List, Optional

import torch
import torch.nn as nn

from tf_yarn_b200.parallel.comm import Communicator

_FIRST_BUCKET_BYTES = 1 << 20


def _view_like(flat: torch.Tensor, off: int, p: torch.Tensor) -> torch.Tensor:
    """View of ``flat[off:off+numel]`` with the shape AND memory layout of ``p``.

    autograd's gradient layout contract wants ``grad.strides == param.strides``; a channels_last conv
    weight therefore gets a channels_last view of its slice (otherwise every accumulation re-lays the
    gradient out and warns)."""
    seg = flat[off:off + p.numel()]
    if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
        n, c, h, w = p.shape
        return seg.view(n, h, w, c).permute(0, 3, 1, 2)
    return seg.view_as(p)


class _Bucket:
    __slots__ = ("params", "offsets", "flat", "pending", "launched", "event", "dtype")

    def __init__(self, dtype):
        self.params: List[nn.Parameter] = []
        self.offsets: List[int] = []
        self.flat: Optional[torch.Tensor] = None
        self.pending = 0
        self.launched = False
        self.event: Optional[torch.cuda.Event] = None
        self.dtype = dtype


class DistributedDataParallel(nn.Module):
    _is_tfy_ddp = True

    def __init__(self, module: nn.Module, comm: Communicator, broadcast_buffers: bool = True,
                 bucket_cap_mb: int = 25, find_unused_parameters: bool = False,
                 gradient_as_bucket_view: bool = False, device_ids=None):
        super().__init__()
        self.module = module
        self.comm = comm
        self.broadcast_buffers = broadcast_buffers
        self.find_unused_parameters = find_unused_parameters
        self.require_backward_grad_sync = True
        self._comm_stream = torch.cuda.Stream(device=comm.device) if comm.world > 1 else None
        self._callback_queued = False
        self._next_bucket = 0            # buckets [0, _next_bucket) have been launched in this backward
        self._buckets: List[_Bucket] = []
        self._param_bucket = {}          # id(param) -> (bucket, index in bucket); tensors compare elementwise
        self._build_buckets(int(bucket_cap_mb) << 20)
        self._sync_params_and_buffers()
        for b in self._buckets:
            for p in b.params:
                p.register_post_accumulate_grad_hook(self._make_hook(p))

    # ------------------------------------------------------------------ set-up
    def _build_buckets(self, cap_bytes: int) -> None:
        params = [p for p in self.module.parameters() if p.requires_grad]
        cur: Optional[_Bucket] = None
        cur_bytes = 0
        limit = _FIRST_BUCKET_BYTES
        for p in reversed(params):
            if p.dtype not in (torch.float32, torch.bfloat16):
                raise TypeError(f"unsupported parameter dtype {p.dtype}: use float32 or bfloat16")
            nbytes = p.numel() * p.element_size()
            if cur is None or cur.dtype != p.dtype or (cur_bytes + nbytes > limit and cur.params):
                if cur is not None:
                    limit = cap_bytes
                cur = _Bucket(p.dtype)
                self._buckets.append(cur)
                cur_bytes = 0
            cur.params.append(p)
            cur.offsets.append(cur_bytes // p.element_size())
            cur_bytes += (nbytes + 15) // 16 * 16
        for b in self._buckets:
            last = b.params[-1]
            n = b.offsets[-1] + (last.numel() + 7) // 8 * 8
            n = self.comm.pad_elems(n, b.dtype)
            _, b.flat = self.comm.arena.empty((n,), b.dtype, align=4096)
            b.flat.zero_()
            for i, (p, o) in enumerate(zip(b.params, b.offsets)):
                p.grad = _view_like(b.flat, o, p)
                self._param_bucket[id(p)] = (b, i)
            b.pending = len(b.params)

    def _sync_params_and_buffers(self) -> None:
        if self.comm.world == 1:
            return
        tensors = [p.data for p in self.module.parameters()] + [b.data for b in self.module.buffers()]
        self.comm.broadcast(tensors, root=0)
        torch.cuda.current_stream().synchronize()

    def _grad_view(self, p: nn.Parameter) -> torch.Tensor:
        b, i = self._param_bucket[id(p)]
        o = b.offsets[i]
        return _view_like(b.flat, o, p)

    # ------------------------------------------------------------------ hooks
    def _make_hook(self, p: nn.Parameter):
        bucket, idx = self._param_bucket[id(p)]
        off = bucket.offsets[idx]
        view = _view_like(bucket.flat, off, p)

        def hook(param: nn.Parameter) -> None:
            g = param.grad
            if g is not None and g.data_ptr() != view.data_ptr():
                # the user dropped the view (zero_grad(set_to_none=True)): fold the fresh grad back in
                view.copy_(g)
                param.grad = view
            if not self.require_backward_grad_sync:
                return
            if not self._callback_queued:
                self._callback_queued = True
                torch.autograd.Variable._execution_engine.queue_callback(self._finalize_backward)
            bucket.pending -= 1
            if bucket.pending == 0:
                self._launch_ready_prefix()

        return hook

    def _launch_ready_prefix(self) -> None:
        """Launch complete buckets STRICTLY in index order (bucket i only after 0..i-1), as torch DDP does: the
        kernels of different ranks are paired by launch order on the per-CTA flag slots, so every rank must issue
        the same sequence even if autograd completes the buckets in a different order on some rank."""
        while self._next_bucket < len(self._buckets):
            b = self._buckets[self._next_bucket]
            if b.pending != 0:
                return
            self._launch(b)
            self._next_bucket += 1

    def _launch(self, bucket: _Bucket) -> None:
        bucket.launched = True
        if self.comm.world == 1:
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self._comm_stream.wait_event(ready)
        with torch.cuda.stream(self._comm_stream):
            self.comm.all_reduce_symm(bucket.flat, average=True, algo=self.comm.pick_algo_inplace())
            bucket.event = torch.cuda.Event()
            bucket.event.record(self._comm_stream)

    def _finalize_backward(self) -> None:
        self._callback_queued = False
        for b in self._buckets[self._next_bucket:]:
            # buckets still waiting for a predecessor, or whose parameters received no gradient this step
            # (they contribute zeros -- what find_unused_parameters=True does in torch DDP), in index order
            self._launch(b)
        self._next_bucket = 0
        cur = torch.cuda.current_stream() if self.comm.world > 1 else None
        for b in self._buckets:
            if b.event is not None:
                cur.wait_event(b.event)
                b.event = None
            b.launched = False
            b.pending = len(b.params)

    # ---------------------------------------------------------------- forward
    def forward(self, *args, **kwargs):
        if self.broadcast_buffers and self.comm.world > 1 and self.require_backward_grad_sync:
            bufs = [b.data for b in self.module.buffers() if b.is_floating_point()]
            if bufs:
                self.comm.broadcast(bufs, root=0)
        return self.module(*args, **kwargs)

    @contextlib.contextmanager
    def no_sync(self):
        """Accumulate gradients locally; the next synced backward reduces the accumulated sum."""
        old = self.require_backward_grad_sync
        self.require_backward_grad_sync = False
        try:
            yield
        finally:
            self.require_backward_grad_sync = old

    def zero_grad(self, set_to_none: bool = False) -> None:  # keep the bucket views alive
        for b in self._buckets:
            b.flat.zero_()
            for p, o in zip(b.params, b.offsets):
                p.grad = _view_like(b.flat, o, p)

    # state_dict()/load_state_dict() are nn.Module's: keys carry the ``module.`` prefix exactly like
    # torch.nn.parallel.DistributedDataParallel, so checkpoints are interchangeable with the reference's
    # (tf_yarn/pytorch/model_ckpt.py saves ``model.module.state_dict()`` when it sees a wrapper).


def wrap_model(model: nn.Module, device, ddp_kwargs: Optional[dict] = None, comm: Optional[Communicator] = None):
    """Data-parallel wrapper appropriate for the process: ours on B200, torch DDP on CPU/gloo."""
    import torch.distributed as dist
    ddp_kwargs = dict(ddp_kwargs or {})
    if isinstance(device, str) and device.startswith("cuda") or isinstance(device, int):
        if comm is None:
            from tf_yarn_b200.parallel import runtime
            comm = runtime.get_communicator()
        return DistributedDataParallel(model, comm, **ddp_kwargs)
    if dist.is_initialized() and dist.get_world_size() > 1:
        from torch.nn.parallel import DistributedDataParallel as TorchDDP
        return TorchDDP(model, **ddp_kwargs)
    return model
