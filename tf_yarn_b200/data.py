"""A small ``tf.data``-shaped input pipeline so the reference's examples translate 1:1.

The reference's examples build inputs as
``CsvDataset(...).filter(...).map(...).shuffle(1000).batch(128).repeat()``
(reference: tf_yarn/examples/winequality.py:30-41, keras_example.py:42-53).
``Dataset`` offers those combinators over in-memory tensors or any python
iterable, yielding ``(features, labels)`` batches as torch tensors.
"""
from __future__ import annotations

import csv
import random
from typing import Any, Callable, Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch


def _stack(items: List[Any]):
    first = items[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_stack([it[i] for it in items]) for i in range(len(first)))
    if isinstance(first, dict):
        return {k: _stack([it[k] for it in items]) for k in first}
    if isinstance(first, torch.Tensor):
        return torch.stack(items)
    return torch.as_tensor(np.asarray(items))


class Dataset:
    """Lazy, re-iterable pipeline."""

    def __init__(self, gen_fn: Callable[[], Iterator[Any]], cardinality: Optional[int] = None):
        self._gen_fn = gen_fn
        self.cardinality = cardinality

    def __iter__(self) -> Iterator[Any]:
        return self._gen_fn()

    # -- sources -------------------------------------------------------------
    @staticmethod
    def from_tensor_slices(tensors) -> "Dataset":
        def conv(t):
            if isinstance(t, (tuple, list)):
                return type(t)(conv(x) for x in t)
            if isinstance(t, dict):
                return {k: conv(v) for k, v in t.items()}
            return t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t))
        tensors = conv(tensors)

        def length(t):
            if isinstance(t, (tuple, list)):
                return length(t[0])
            if isinstance(t, dict):
                return length(next(iter(t.values())))
            return t.shape[0]

        def index(t, i):
            if isinstance(t, (tuple, list)):
                return type(t)(index(x, i) for x in t)
            if isinstance(t, dict):
                return {k: index(v, i) for k, v in t.items()}
            return t[i]
        n = length(tensors)
        ds = Dataset(lambda: (index(tensors, i) for i in range(n)), n)
        ds._tensors = tensors       # fast path for batch()
        return ds

    @staticmethod
    def from_generator(gen_fn: Callable[[], Iterable[Any]]) -> "Dataset":
        return Dataset(lambda: iter(gen_fn()))

    @staticmethod
    def range(n: int) -> "Dataset":
        return Dataset(lambda: iter(range(n)), n)

    # -- combinators -----------------------------------------------------------
    def map(self, fn: Callable) -> "Dataset":
        def gen():
            for item in self:
                yield fn(*item) if isinstance(item, tuple) else fn(item)
        return Dataset(gen, self.cardinality)

    def filter(self, pred: Callable) -> "Dataset":
        def gen():
            for item in self:
                keep = pred(*item) if isinstance(item, tuple) else pred(item)
                if bool(keep):
                    yield item
        return Dataset(gen)

    def shuffle(self, buffer_size: int, seed: Optional[int] = None) -> "Dataset":
        def gen():
            rng = random.Random(seed)
            buf: List[Any] = []
            for item in self:
                buf.append(item)
                if len(buf) >= buffer_size:
                    i = rng.randrange(len(buf))
                    buf[i], buf[-1] = buf[-1], buf[i]
                    yield buf.pop()
            rng.shuffle(buf)
            yield from buf
        return Dataset(gen, self.cardinality)

    def batch(self, batch_size: int, drop_remainder: bool = False) -> "Dataset":
        def gen():
            items: List[Any] = []
            for item in self:
                items.append(item)
                if len(items) == batch_size:
                    yield _stack(items)
                    items = []
            if items and not drop_remainder:
                yield _stack(items)
        card = None
        if self.cardinality is not None:
            card = self.cardinality // batch_size if drop_remainder else -(-self.cardinality // batch_size)
        return Dataset(gen, card)

    def repeat(self, count: Optional[int] = None) -> "Dataset":
        def gen():
            n = 0
            while count is None or n < count:
                empty = True
                for item in self:
                    empty = False
                    yield item
                if empty:
                    return
                n += 1
        card = None if count is None or self.cardinality is None else self.cardinality * count
        return Dataset(gen, card)

    def take(self, n: int) -> "Dataset":
        def gen():
            for i, item in enumerate(self):
                if i >= n:
                    return
                yield item
        return Dataset(gen, n if self.cardinality is None else min(n, self.cardinality))

    def skip(self, n: int) -> "Dataset":
        def gen():
            for i, item in enumerate(self):
                if i >= n:
                    yield item
        return Dataset(gen)

    def shard(self, num_shards: int, index: int) -> "Dataset":
        def gen():
            for i, item in enumerate(self):
                if i % num_shards == index:
                    yield item
        return Dataset(gen)

    def prefetch(self, n: int = 1) -> "Dataset":  # host pipelines are synchronous; GPU prefetch lives in fit()
        return self


def CsvDataset(path: str, record_defaults: Sequence[Any], header: bool = False, field_delim: str = ",") -> Dataset:
    """Rows of a CSV file as tuples typed after ``record_defaults`` (float / int / str exemplars or types)."""
    types = [d if isinstance(d, type) else type(d) for d in record_defaults]

    def gen():
        with open(path, newline="") as f:
            reader = csv.reader(f, delimiter=field_delim)
            if header:
                next(reader, None)
            for row in reader:
                if not row:
                    continue
                yield tuple(t(v) if v != "" else t() for t, v in zip(types, row))
    return Dataset(gen)
