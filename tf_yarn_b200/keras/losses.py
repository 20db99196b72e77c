"""Loss functions (mean over the batch), computed in fp32."""
from __future__ import annotations

from typing import Callable, Union

import torch
import torch.nn.functional as F


def sparse_categorical_crossentropy(y_true, y_pred, from_logits: bool = False):
    y_pred = y_pred.float()
    y_true = y_true.long().reshape(-1)
    y_pred = y_pred.reshape(-1, y_pred.shape[-1])
    if from_logits:
        return F.cross_entropy(y_pred, y_true)
    return F.nll_loss(torch.log(y_pred.clamp_min(1e-7)), y_true)


def categorical_crossentropy(y_true, y_pred, from_logits: bool = False):
    y_pred = y_pred.float()
    logp = F.log_softmax(y_pred, dim=-1) if from_logits else torch.log(y_pred.clamp_min(1e-7))
    return -(y_true.float() * logp).sum(dim=-1).mean()


def binary_crossentropy(y_true, y_pred, from_logits: bool = False):
    y_pred = y_pred.float().reshape(-1)
    y_true = y_true.float().reshape(-1)
    if from_logits:
        return F.binary_cross_entropy_with_logits(y_pred, y_true)
    return F.binary_cross_entropy(y_pred.clamp(1e-7, 1 - 1e-7), y_true)


def mean_squared_error(y_true, y_pred):
    return F.mse_loss(y_pred.float(), y_true.float().reshape(y_pred.shape))


def mean_absolute_error(y_true, y_pred):
    return F.l1_loss(y_pred.float(), y_true.float().reshape(y_pred.shape))


class SparseCategoricalCrossentropy:
    def __init__(self, from_logits: bool = False):
        self.from_logits = from_logits
        self.__name__ = "sparse_categorical_crossentropy"

    def __call__(self, y_true, y_pred):
        return sparse_categorical_crossentropy(y_true, y_pred, self.from_logits)


class CategoricalCrossentropy:
    def __init__(self, from_logits: bool = False):
        self.from_logits = from_logits
        self.__name__ = "categorical_crossentropy"

    def __call__(self, y_true, y_pred):
        return categorical_crossentropy(y_true, y_pred, self.from_logits)


class BinaryCrossentropy:
    def __init__(self, from_logits: bool = False):
        self.from_logits = from_logits
        self.__name__ = "binary_crossentropy"

    def __call__(self, y_true, y_pred):
        return binary_crossentropy(y_true, y_pred, self.from_logits)


_BY_NAME = {
    "sparse_categorical_crossentropy": sparse_categorical_crossentropy,
    "categorical_crossentropy": categorical_crossentropy,
    "binary_crossentropy": binary_crossentropy,
    "mse": mean_squared_error, "mean_squared_error": mean_squared_error,
    "mae": mean_absolute_error, "mean_absolute_error": mean_absolute_error,
}


def get(identifier: Union[str, Callable]) -> Callable:
    if callable(identifier):
        return identifier
    if identifier not in _BY_NAME:
        raise ValueError(f"unknown loss {identifier!r}")
    return _BY_NAME[identifier]


def name_of(identifier) -> str:
    return identifier if isinstance(identifier, str) else getattr(identifier, "__name__", "loss")


def serialize(identifier):
    """JSON-able description of a loss (None when it is an arbitrary callable)."""
    if identifier is None:
        return None
    if isinstance(identifier, str):
        return {"name": identifier}
    if isinstance(identifier, (SparseCategoricalCrossentropy, CategoricalCrossentropy, BinaryCrossentropy)):
        return {"name": identifier.__name__, "from_logits": identifier.from_logits}
    return None


def deserialize(cfg):
    if not cfg:
        return None
    if "from_logits" in cfg:
        cls = {"sparse_categorical_crossentropy": SparseCategoricalCrossentropy,
               "categorical_crossentropy": CategoricalCrossentropy,
               "binary_crossentropy": BinaryCrossentropy}[cfg["name"]]
        return cls(from_logits=cfg["from_logits"])
    return cfg["name"]
