"""Keras-style optimizer descriptors.

A descriptor only carries hyper-parameters.  On B200 the train engine maps it to
the fused K4 kernel (``OptimizerSpec``); on CPU it is instantiated as the
equivalent ``torch.optim`` optimizer (same update formulas, see
tests/test_optimizer_math.py).
"""
from __future__ import annotations

from typing import Iterable

import torch

from tf_yarn_b200.parallel.optspec import OptimizerSpec


class Optimizer:
    kind = "sgd"

    def __init__(self, learning_rate: float, **extra):
        self.learning_rate = float(learning_rate)
        self.distributed = False        # set by hvd.DistributedOptimizer
        self.extra = extra

    # lr alias used by Keras 2 code
    @property
    def lr(self) -> float:
        return self.learning_rate

    @lr.setter
    def lr(self, v: float) -> None:
        self.learning_rate = float(v)

    def to_spec(self) -> OptimizerSpec:
        raise NotImplementedError

    def to_torch(self, params: Iterable[torch.nn.Parameter]) -> torch.optim.Optimizer:
        raise NotImplementedError

    def get_config(self):
        return {"class_name": type(self).__name__, "learning_rate": self.learning_rate, **self.extra}


class SGD(Optimizer):
    def __init__(self, learning_rate: float = 0.01, momentum: float = 0.0, nesterov: bool = False,
                 weight_decay: float = 0.0, lr: float = None):
        super().__init__(lr if lr is not None else learning_rate, momentum=momentum, nesterov=nesterov,
                         weight_decay=weight_decay)
        self.momentum, self.nesterov, self.weight_decay = momentum, nesterov, weight_decay

    def to_spec(self):
        return OptimizerSpec.sgd(self.learning_rate, self.momentum, 0.0, self.nesterov, self.weight_decay)

    def to_torch(self, params):
        return torch.optim.SGD(params, lr=self.learning_rate, momentum=self.momentum, nesterov=self.nesterov,
                               weight_decay=self.weight_decay)


class Adadelta(Optimizer):
    def __init__(self, learning_rate: float = 0.001, rho: float = 0.95, epsilon: float = 1e-7,
                 weight_decay: float = 0.0, lr: float = None):
        super().__init__(lr if lr is not None else learning_rate, rho=rho, epsilon=epsilon,
                         weight_decay=weight_decay)
        self.rho, self.epsilon, self.weight_decay = rho, epsilon, weight_decay

    def to_spec(self):
        return OptimizerSpec.adadelta(self.learning_rate, self.rho, self.epsilon, self.weight_decay)

    def to_torch(self, params):
        return torch.optim.Adadelta(params, lr=self.learning_rate, rho=self.rho, eps=self.epsilon,
                                    weight_decay=self.weight_decay)


class Adam(Optimizer):
    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999,
                 epsilon: float = 1e-7, weight_decay: float = 0.0, decoupled_weight_decay: bool = False,
                 lr: float = None):
        super().__init__(lr if lr is not None else learning_rate, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon,
                         weight_decay=weight_decay, decoupled_weight_decay=decoupled_weight_decay)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.weight_decay, self.decoupled = weight_decay, decoupled_weight_decay

    def to_spec(self):
        return OptimizerSpec.adam(self.learning_rate, self.beta_1, self.beta_2, self.epsilon, self.weight_decay,
                                  self.decoupled)

    def to_torch(self, params):
        cls = torch.optim.AdamW if self.decoupled else torch.optim.Adam
        return cls(params, lr=self.learning_rate, betas=(self.beta_1, self.beta_2), eps=self.epsilon,
                   weight_decay=self.weight_decay)


class AdamW(Adam):
    def __init__(self, learning_rate: float = 0.001, weight_decay: float = 0.004, **kw):
        super().__init__(learning_rate, weight_decay=weight_decay, decoupled_weight_decay=True, **kw)


class Adagrad(Optimizer):
    def __init__(self, learning_rate: float = 0.001, initial_accumulator_value: float = 0.1,
                 epsilon: float = 1e-7, weight_decay: float = 0.0, lr: float = None):
        super().__init__(lr if lr is not None else learning_rate,
                         initial_accumulator_value=initial_accumulator_value, epsilon=epsilon,
                         weight_decay=weight_decay)
        self.initial_accumulator_value, self.epsilon, self.weight_decay = \
            initial_accumulator_value, epsilon, weight_decay

    def to_spec(self):
        return OptimizerSpec.adagrad(self.learning_rate, self.epsilon, self.weight_decay,
                                     self.initial_accumulator_value)

    def to_torch(self, params):
        return torch.optim.Adagrad(params, lr=self.learning_rate, eps=self.epsilon,
                                   weight_decay=self.weight_decay,
                                   initial_accumulator_value=self.initial_accumulator_value)


class Ftrl(Optimizer):
    """FTRL-proximal (``tf.keras.optimizers.Ftrl`` / ``tf.train.FtrlOptimizer``), TF's default optimizer of linear
    models and the canonical choice for the wide tower of wide-and-deep (learning_rate_power fixed at -0.5)."""
    kind = "ftrl"

    def __init__(self, learning_rate: float = 0.001, learning_rate_power: float = -0.5,
                 initial_accumulator_value: float = 0.1, l1_regularization_strength: float = 0.0,
                 l2_regularization_strength: float = 0.0, beta: float = 0.0, lr: float = None):
        if learning_rate_power != -0.5:
            raise ValueError("Ftrl: only learning_rate_power=-0.5 is implemented")
        super().__init__(lr if lr is not None else learning_rate, learning_rate_power=learning_rate_power,
                         initial_accumulator_value=initial_accumulator_value,
                         l1_regularization_strength=l1_regularization_strength,
                         l2_regularization_strength=l2_regularization_strength, beta=beta)
        self.initial_accumulator_value = initial_accumulator_value
        self.l1, self.l2, self.beta = l1_regularization_strength, l2_regularization_strength, beta

    def to_spec(self):
        return OptimizerSpec.ftrl(self.learning_rate, self.l1, self.l2, self.beta, self.initial_accumulator_value)

    def to_torch(self, params):
        return TorchFtrl(params, lr=self.learning_rate, l1=self.l1, l2=self.l2, beta=self.beta,
                         initial_accumulator_value=self.initial_accumulator_value)


class TorchFtrl(torch.optim.Optimizer):
    """torch.optim rendition of FTRL-proximal (same formulas as the fused / parameter-server kernels)."""

    def __init__(self, params, lr=1e-3, l1=0.0, l2=0.0, beta=0.0, initial_accumulator_value=0.1):
        super().__init__(params, dict(lr=lr, l1=l1, l2=l2, beta=beta, init=initial_accumulator_value))

    @torch.no_grad()
    def step(self, closure=None):
        for grp in self.param_groups:
            lr, l1, l2, beta = grp["lr"], grp["l1"], grp["l2"], grp["beta"]
            for p in grp["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["n"] = torch.full_like(p, grp["init"])
                    st["z"] = torch.zeros_like(p)
                g, n, z = p.grad, st["n"], st["z"]
                n_new = n + g * g
                z.add_(g - (n_new.sqrt() - n.sqrt()) / lr * p)
                n.copy_(n_new)
                w = -(z - torch.sign(z) * l1) / ((beta + n_new.sqrt()) / lr + 2 * l2)
                p.copy_(torch.where(z.abs() <= l1, torch.zeros_like(w), w))


_BY_NAME = {"sgd": SGD, "adadelta": Adadelta, "adam": Adam, "adamw": AdamW, "adagrad": Adagrad, "ftrl": Ftrl}


def get(identifier) -> Optimizer:
    """Resolve a string / descriptor / distributed wrapper into a descriptor."""
    if isinstance(identifier, Optimizer):
        return identifier
    if isinstance(identifier, str):
        if identifier.lower() not in _BY_NAME:
            raise ValueError(f"unknown optimizer {identifier!r}")
        return _BY_NAME[identifier.lower()]()
    inner = getattr(identifier, "_tfy_inner_optimizer", None)
    if inner is not None:
        opt = get(inner)
        opt.distributed = True
        return opt
    raise TypeError(f"cannot interpret optimizer {identifier!r}")


def from_config(cfg) -> Optimizer:
    cfg = dict(cfg)
    cls = _BY_NAME[cfg.pop("class_name").lower()]
    return cls(**cfg)
