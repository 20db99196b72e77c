"""Feature columns (``tf.feature_column`` subset) for the canned estimators.

Used by the reference's examples as ``tf.feature_column.numeric_column(name)``
(reference: tf_yarn/examples/winequality.py:44-45) and by the wide-and-deep
BASELINE config (categorical + embedding columns).
"""
from __future__ import annotations

from typing import Dict, List, NamedTuple, Sequence, Tuple, Union

import torch
import torch.nn as nn


class NumericColumn(NamedTuple):
    key: str
    shape: Tuple[int, ...] = (1,)

    @property
    def dim(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


class CategoricalColumn(NamedTuple):
    key: str
    num_buckets: int
    hashed: bool = False


class EmbeddingColumn(NamedTuple):
    categorical_column: CategoricalColumn
    dimension: int
    combiner: str = "mean"


class IndicatorColumn(NamedTuple):
    categorical_column: CategoricalColumn


def numeric_column(key: str, shape: Sequence[int] = (1,), **_ignored) -> NumericColumn:
    return NumericColumn(key, tuple(shape))


def categorical_column_with_identity(key: str, num_buckets: int, **_ignored) -> CategoricalColumn:
    return CategoricalColumn(key, int(num_buckets), False)


def categorical_column_with_hash_bucket(key: str, hash_bucket_size: int, **_ignored) -> CategoricalColumn:
    return CategoricalColumn(key, int(hash_bucket_size), True)


def embedding_column(categorical_column: CategoricalColumn, dimension: int, combiner: str = "mean",
                     **_ignored) -> EmbeddingColumn:
    return EmbeddingColumn(categorical_column, int(dimension), combiner)


def indicator_column(categorical_column: CategoricalColumn) -> IndicatorColumn:
    return IndicatorColumn(categorical_column)


def _ids(col: CategoricalColumn, t: torch.Tensor) -> torch.Tensor:
    ids = t.long()
    if col.hashed:
        # cheap integer hash (Knuth multiplicative), stable across processes
        ids = (ids * 2654435761) % (2 ** 32)
    return ids % col.num_buckets


class DenseFeatures(nn.Module):
    """Concatenate numeric / indicator / embedding columns into one dense [batch, width] tensor."""

    def __init__(self, columns: Sequence[Union[NumericColumn, EmbeddingColumn, IndicatorColumn]]):
        super().__init__()
        self.columns = list(columns)
        self.embeddings = nn.ModuleDict()
        width = 0
        for c in self.columns:
            if isinstance(c, NumericColumn):
                width += c.dim
            elif isinstance(c, EmbeddingColumn):
                emb = nn.EmbeddingBag(c.categorical_column.num_buckets, c.dimension,
                                      mode="mean" if c.combiner == "mean" else "sum")
                nn.init.normal_(emb.weight, std=1.0 / c.dimension ** 0.5)
                self.embeddings[c.categorical_column.key] = emb
                width += c.dimension
            elif isinstance(c, IndicatorColumn):
                width += c.categorical_column.num_buckets
            else:
                raise TypeError(f"{c!r} cannot feed a dense layer; wrap it in embedding_column/indicator_column")
        self.width = width

    def forward(self, features: Dict[str, torch.Tensor]) -> torch.Tensor:
        parts: List[torch.Tensor] = []
        ref = None
        for c in self.columns:
            if isinstance(c, NumericColumn):
                t = features[c.key]
                parts.append(t.reshape(t.shape[0], -1))
            elif isinstance(c, EmbeddingColumn):
                t = features[c.categorical_column.key]
                ids = _ids(c.categorical_column, t).reshape(t.shape[0], -1)
                parts.append(self.embeddings[c.categorical_column.key](ids))
            else:
                t = features[c.categorical_column.key]
                ids = _ids(c.categorical_column, t).reshape(t.shape[0], -1)
                parts.append(torch.zeros(t.shape[0], c.categorical_column.num_buckets, device=t.device)
                             .scatter_(1, ids, 1.0))
        dtype = next((p.dtype for p in parts if p.is_floating_point()), torch.float32)
        return torch.cat([p.to(dtype) for p in parts], dim=1)


class LinearModel(nn.Module):
    """Wide part: a weight per numeric dimension and per categorical bucket, plus a bias."""

    def __init__(self, columns: Sequence[Union[NumericColumn, CategoricalColumn]], units: int):
        super().__init__()
        self.columns = list(columns)
        self.units = units
        n_numeric = sum(c.dim for c in self.columns if isinstance(c, NumericColumn))
        self.numeric = nn.Linear(n_numeric, units, bias=False) if n_numeric else None
        if self.numeric is not None:
            nn.init.zeros_(self.numeric.weight)
        self.tables = nn.ModuleDict()
        for c in self.columns:
            if isinstance(c, (EmbeddingColumn, IndicatorColumn)):
                c = c.categorical_column
            if isinstance(c, CategoricalColumn):
                emb = nn.EmbeddingBag(c.num_buckets, units, mode="sum")
                nn.init.zeros_(emb.weight)
                self.tables[c.key] = emb
        self.bias = nn.Parameter(torch.zeros(units))

    def forward(self, features: Dict[str, torch.Tensor]) -> torch.Tensor:
        out = None
        nums = [features[c.key].reshape(features[c.key].shape[0], -1).float() for c in self.columns
                if isinstance(c, NumericColumn)]
        if nums:
            x = torch.cat(nums, dim=1)
            out = self.numeric(x.to(self.numeric.weight.dtype))
        for c in self.columns:
            cc = c.categorical_column if isinstance(c, (EmbeddingColumn, IndicatorColumn)) else c
            if isinstance(cc, CategoricalColumn):
                t = features[cc.key]
                y = self.tables[cc.key](_ids(cc, t).reshape(t.shape[0], -1))
                out = y if out is None else out + y
        return out + self.bias
