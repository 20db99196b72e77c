"""Host-side logic of the DDP-like wrapper (bucketing, grad views, hooks) with a single-rank fake communicator."""
import torch
import torch.nn as nn

from tf_yarn_b200.parallel.ddp import DistributedDataParallel


class _FakeArena:
    def empty(self, shape, dtype, align=256):
        return 0, torch.zeros(shape, dtype=dtype)


class _FakeComm:
    world, rank, device = 1, 0, 0
    arena = _FakeArena()

    def pad_elems(self, n, dtype):
        return (n + 7) // 8 * 8


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(300, 700), nn.ReLU(), nn.Linear(700, 600), nn.ReLU(), nn.Linear(600, 10))


def test_buckets_are_built_in_reverse_order_with_a_small_first_bucket():
    ddp = DistributedDataParallel(_model(), _FakeComm(), bucket_cap_mb=1)
    sizes = [sum(p.numel() * 4 for p in b.params) for b in ddp._buckets]
    assert len(ddp._buckets) >= 2
    assert sizes[0] <= 1 << 20 or len(ddp._buckets[0].params) == 1
    first_params = ddp._buckets[0].params
    last_layer = list(ddp.module.parameters())[-1]
    assert first_params[0] is last_layer                    # reverse registration order


def test_gradients_land_in_bucket_views_and_match_plain_backward():
    model, ref = _model(), _model()
    ddp = DistributedDataParallel(model, _FakeComm(), bucket_cap_mb=1)
    x = torch.randn(16, 300)
    ddp(x).sum().backward()
    ref(x).sum().backward()
    for p, r in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, r.grad, atol=1e-6)
        b, i = ddp._param_bucket[id(p)]
        assert p.grad.data_ptr() == b.flat[b.offsets[i]:].data_ptr()       # a view into the flat bucket
    assert all(b.pending == len(b.params) and not b.launched for b in ddp._buckets)   # reset after backward


def test_zero_grad_set_to_none_is_recovered_and_no_sync_accumulates():
    model = _model()
    ddp = DistributedDataParallel(model, _FakeComm(), bucket_cap_mb=1)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    x = torch.randn(8, 300)
    ddp(x).sum().backward()
    g1 = [p.grad.clone() for p in model.parameters()]
    opt.zero_grad(set_to_none=True)                       # user code drops the views ...
    ddp(x).sum().backward()                               # ... the hook folds the fresh grads back into them
    for p, g in zip(model.parameters(), g1):
        b, i = ddp._param_bucket[id(p)]
        assert p.grad.data_ptr() == b.flat[b.offsets[i]:].data_ptr()
        assert torch.allclose(p.grad, g, atol=1e-6)
    ddp.zero_grad()
    with ddp.no_sync():
        ddp(x).sum().backward()
    ddp(x).sum().backward()
    for p, g in zip(model.parameters(), g1):
        assert torch.allclose(p.grad, 2 * g, atol=1e-5)


def test_state_dict_keys_carry_the_module_prefix_like_torch_ddp():
    """Checkpoints written from the wrapper are interchangeable with torch DDP's (keys `module.<name>`)."""
    model = _model()
    ddp = DistributedDataParallel(model, _FakeComm())
    assert set(ddp.state_dict().keys()) == {"module." + k for k in model.state_dict().keys()}
    other = DistributedDataParallel(_model(), _FakeComm())
    other.load_state_dict(ddp.state_dict())
    for a, b in zip(other.module.parameters(), model.parameters()):
        assert torch.equal(a, b)


def test_buckets_launch_strictly_in_index_order(monkeypatch):
    """Even when autograd completes a later bucket first, kernels are issued bucket 0, 1, 2, ... (ranks pair the
    collectives by launch order)."""
    model = _model()
    ddp = DistributedDataParallel(model, _FakeComm(), bucket_cap_mb=1)
    order = []
    monkeypatch.setattr(ddp, "_launch", lambda b: (order.append(ddp._buckets.index(b)), setattr(b, "launched", True)))
    # complete the LAST bucket first by hand, then the others
    for b in reversed(ddp._buckets):
        b.pending = 0
        ddp._launch_ready_prefix()
    assert order == list(range(len(ddp._buckets)))


def test_channels_last_parameters_get_channels_last_gradient_views():
    """The gradient view of a channels_last conv weight must have the parameter's strides (autograd's
    gradient layout contract) and still alias the flat bucket."""
    torch.manual_seed(0)
    model = nn.Sequential(nn.Conv2d(3, 8, 3), nn.ReLU(), nn.Conv2d(8, 4, 3)).to(memory_format=torch.channels_last)
    ref = nn.Sequential(nn.Conv2d(3, 8, 3), nn.ReLU(), nn.Conv2d(8, 4, 3))
    ref.load_state_dict(model.state_dict())
    ddp = DistributedDataParallel(model, _FakeComm(), bucket_cap_mb=1)
    for p in model.parameters():
        assert p.grad.stride() == p.stride()
    x = torch.randn(2, 3, 10, 10)
    ddp(x).sum().backward()
    ref(x).sum().backward()
    for p, q in zip(model.parameters(), ref.parameters()):
        assert p.grad.stride() == p.stride()
        assert torch.allclose(p.grad, q.grad, atol=1e-5)
        b, i = ddp._param_bucket[id(p)]
        assert p.grad.untyped_storage().data_ptr() == b.flat.untyped_storage().data_ptr()


def test_fused_optimizer_state_is_sharded_and_foreign_shards_are_not_adopted(caplog):
    """Rank 0 saves ITS shard (the reference saves on rank 0 only).  Rank 0 restores it verbatim; another rank must
    not adopt rank 0's master weights: it rebuilds its own shard from the restored parameters."""
    from types import SimpleNamespace
    import torch
    from tf_yarn_b200.parallel.ddp import _FusedOptimizer
    from tf_yarn_b200.parallel.optspec import OptimizerSpec

    def make(rank):
        shard_n = 8
        pflat = torch.arange(16, dtype=torch.float32)                      # the (restored) parameters, both shards
        b = SimpleNamespace(params=[], shard_n=shard_n, pflat=pflat, master=torch.zeros(shard_n),
                            s1=torch.zeros(shard_n), s2=torch.zeros(shard_n), hyper=torch.zeros(64, dtype=torch.uint8))
        ddp = SimpleNamespace(_buckets=[b], comm=SimpleNamespace(rank=rank, world=2))
        return _FusedOptimizer(ddp, OptimizerSpec.adam(1e-3)), b

    saver, sb = make(0)
    sb.master.copy_(torch.arange(8, dtype=torch.float32) + 0.25)           # fp32 master carries more than the params
    sb.s1.fill_(0.5), sb.s2.fill_(0.125)
    sb.hyper[24:28].copy_(torch.tensor([7], dtype=torch.int32).view(torch.uint8))
    state = saver.state_dict()
    assert state["buckets"][0]["rank"] == 0 and state["buckets"][0]["step"] == 7

    same, b0 = make(0)
    same.load_state_dict(state)
    assert torch.equal(b0.master, sb.master) and float(b0.s1[0]) == 0.5 and float(b0.s2[0]) == 0.125
    assert int(b0.hyper[24:28].view(torch.int32).item()) == 7

    other, b1 = make(1)
    with caplog.at_level("WARNING"):
        other.load_state_dict(state)
    assert torch.equal(b1.master, torch.arange(8, 16, dtype=torch.float32))     # ITS slice of the parameters
    assert float(b1.s1.abs().max()) == 0.0 and float(b1.s2.abs().max()) == 0.0
    assert int(b1.hyper[24:28].view(torch.int32).item()) == 7 and "another rank's shard" in caplog.text


def test_fused_masters_follow_a_loaded_model_checkpoint():
    """After module.load_state_dict(...) the fp32 master shards must be re-derived from the parameters: the fused
    kernel writes parameters FROM the master, so a stale master would undo the load at the next step."""
    from types import SimpleNamespace
    import torch
    from tf_yarn_b200.parallel.ddp import DistributedDataParallel
    pflat = torch.arange(16, dtype=torch.bfloat16)
    b = SimpleNamespace(shard_n=8, pflat=pflat, master=torch.full((8,), -1.0))
    stub = SimpleNamespace(comm=SimpleNamespace(rank=1, world=2), _buckets=[b])
    DistributedDataParallel._resync_fused_masters(stub)
    assert b.master.dtype == torch.float32 and torch.equal(b.master, torch.arange(8, 16, dtype=torch.float32))
    solo = SimpleNamespace(comm=SimpleNamespace(rank=0, world=1), _buckets=[b])
    DistributedDataParallel._resync_fused_masters(solo)
    assert torch.equal(b.master, torch.arange(0, 8, dtype=torch.float32))
