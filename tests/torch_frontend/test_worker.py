import os
from unittest import mock

import torch

from tf_yarn_b200.pytorch import DataLoaderArgs
from tf_yarn_b200.pytorch.tasks import worker

from fakes import FakeClient


def test_get_device_cpu_is_none():
    with mock.patch("torch.cuda.is_available", return_value=False):
        assert worker._get_device(3) is None
        assert worker._get_collective_ops_backend(2) == "gloo"


def test_get_device_follows_launcher_placement(monkeypatch):
    monkeypatch.setenv("TFY_GPU_IDS", "4,5")
    with mock.patch("torch.cuda.is_available", return_value=True), \
            mock.patch("torch.cuda.device_count", return_value=8):
        assert [worker._get_device(i) for i in range(3)] == [4, 5, 4]
        assert worker._get_collective_ops_backend(2) == "nccl"
        assert worker._get_collective_ops_backend(3) == "gloo"     # more processes than GPUs


def test_get_device_without_placement(monkeypatch):
    monkeypatch.delenv("TFY_GPU_IDS", raising=False)
    with mock.patch("torch.cuda.is_available", return_value=True), \
            mock.patch("torch.cuda.device_count", return_value=2):
        assert [worker._get_device(i) for i in range(4)] == [0, 1, 0, 1]


def test_setup_master(monkeypatch):
    monkeypatch.delenv("MASTER_ADDR", raising=False)
    monkeypatch.delenv("MASTER_PORT", raising=False)
    client = FakeClient()
    worker._setup_master(client, 0)
    assert os.environ["MASTER_ADDR"] == client.kv["MASTER_ADDR"].decode()
    port = os.environ["MASTER_PORT"]
    worker._setup_master(client, 1)
    assert os.environ["MASTER_PORT"] == port


def test_create_dataloader_map_style_no_process_group():
    ds = torch.utils.data.TensorDataset(torch.arange(10).float())
    loader = worker._create_dataloader(ds, DataLoaderArgs(batch_size=4, shuffle=False))
    batches = [b[0] for b in loader]
    assert [len(b) for b in batches] == [4, 4]                     # drop_last=True by default


def test_create_dataloader_iterable_gets_no_sampler():
    class It(torch.utils.data.IterableDataset):
        def __iter__(self):
            return iter(range(6))
    loader = worker._create_dataloader(It(), DataLoaderArgs(batch_size=2, shuffle=True))
    assert loader.sampler is None or not isinstance(loader.sampler, torch.utils.data.distributed.DistributedSampler)
    assert len(list(loader)) == 3


def test_drop_last_false_logs_loudly(caplog):
    ds = torch.utils.data.TensorDataset(torch.arange(10).float())
    worker._create_dataloader(ds, DataLoaderArgs(batch_size=4, drop_last=False))
    assert any("drop_last" in r.getMessage() for r in caplog.records)
