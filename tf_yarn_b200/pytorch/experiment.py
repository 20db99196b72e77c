"""Descriptors of a PyTorch experiment (reference: tf_yarn/pytorch/experiment.py:6-56)."""
from __future__ import annotations

from typing import Any, Callable, NamedTuple, Optional

import torch


class DataLoaderArgs(NamedTuple):
    """Arguments forwarded to ``torch.utils.data.DataLoader`` (see torch docs for their meaning)."""
    batch_size: Optional[int] = 1
    num_workers: int = 0
    pin_memory: bool = False
    drop_last: bool = True
    timeout: float = 0
    prefetch_factor: Optional[int] = 2
    shuffle: bool = False
    persistent_workers: bool = False
    collate_fn: Optional[Callable[[Any], Any]] = None


class DistributedDataParallelArgs(NamedTuple):
    """Arguments of the data-parallel wrapper (same names as ``torch.nn.parallel.DistributedDataParallel``)."""
    broadcast_buffers: bool = True
    bucket_cap_mb: int = 25
    find_unused_parameters: bool = False
    gradient_as_bucket_view: bool = False


class PytorchExperiment(NamedTuple):
    # model to train
    model: torch.nn.Module

    # main_fn(model, trainloader, device: "cuda:N" | "cpu", rank: int, tb_writer) -> None
    #   model: the model wrapped for data-parallel training (gradients are averaged across ranks)
    #   trainloader: the rank's shard of ``train_dataset``
    #   device: where the model lives
    #   rank: global rank of the process
    #   tb_writer: torch.utils.tensorboard.SummaryWriter of this rank
    # (the reference documents 4 arguments but calls with these 5: pytorch/tasks/worker.py:113)
    main_fn: Callable[..., None]

    # training set
    train_dataset: Any

    dataloader_args: DataLoaderArgs

    # directory where each rank's tensorboard event files are collected at the end ("worker<rank>")
    tensorboard_hdfs_dir: Optional[str] = None

    ddp_args: Optional[DistributedDataParallelArgs] = None
