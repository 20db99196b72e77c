"""Iterable dataset over a directory of parquet files (reference: tf_yarn/pytorch/parquet_dataset.py:15-72).

Each rank reads a contiguous, equally sized slice of the record batches of
every file; the ragged last batch of a file is dropped so that every rank
performs the same number of all-reduce steps.
"""
from __future__ import annotations

import logging
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional

import pyarrow.parquet as pq
import torch.distributed as dist
from torch.utils.data import IterableDataset

from tf_yarn_b200.utils.fs import resolve_filesystem_and_path

logger = logging.getLogger(__name__)


class ParquetDataset(IterableDataset):
    def __init__(self, dataset_path: str, batch_size: int, num_samples: Optional[int] = None,
                 columns: Optional[List[str]] = None, rank: Optional[int] = None,
                 world_size: Optional[int] = None) -> None:
        self.fs, _ = resolve_filesystem_and_path(dataset_path)
        self.columns = columns
        self.dataset_file_paths = [f for f in self.fs.base_fs.ls(dataset_path) if f.endswith(".parquet")]
        self.num_samples = num_samples if num_samples else _read_num_samples(self.dataset_file_paths)
        self.batch_size = batch_size
        initialized = dist.is_available() and dist.is_initialized()
        self.worker_id = rank if rank is not None else (dist.get_rank() if initialized else 0)
        self.num_workers = world_size if world_size is not None else (dist.get_world_size() if initialized else 1)
        logger.info("worker_id: %d; num_workers: %d", self.worker_id, self.num_workers)

    def __iter__(self):
        for path in self.dataset_file_paths:
            with self.fs.base_fs.open(path) as f:
                pf = pq.ParquetFile(f)
                # the last batch of every file is dropped (it may be ragged and would desynchronise
                # the ranks' all-reduce counts); same rule as the reference, which drops it unconditionally
                all_batches = (pf.metadata.num_rows + self.batch_size - 1) // self.batch_size
                n_batches = all_batches - 1
                per_worker = n_batches // self.num_workers
                assert per_worker > 0, f"{path}: fewer batches ({n_batches}) than workers ({self.num_workers})"
                start = self.worker_id * per_worker
                end = start + per_worker
                # stream: skip batches before `start`, stop after `end` (no need to hold the file in memory)
                for i, batch in enumerate(_rebatch(pf, self.batch_size, self.columns)):
                    if i >= end:
                        break
                    if i >= start:
                        yield batch

    def __len__(self) -> int:
        return self.num_samples // self.batch_size // self.num_workers


def _rebatch(pf: "pq.ParquetFile", batch_size: int, columns):
    """``iter_batches`` may cut at row-group boundaries; pyarrow >= 7 already re-chunks to batch_size."""
    yield from pf.iter_batches(batch_size=batch_size, columns=columns)


def _get_num_rows(path: str) -> int:
    fs, _ = resolve_filesystem_and_path(path)
    with fs.base_fs.open(path) as f:
        return pq.ParquetFile(f).metadata.num_rows


def _read_num_samples(paths: List[str]) -> int:
    with ThreadPoolExecutor(max_workers=5) as pool:
        return sum(pool.map(_get_num_rows, paths))
