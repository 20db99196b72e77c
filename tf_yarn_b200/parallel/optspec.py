"""Fused-optimizer hyper-parameter description (no CUDA dependency)."""
from __future__ import annotations

OPT_SGD, OPT_ADADELTA, OPT_ADAM, OPT_ADAGRAD, OPT_FTRL = 0, 1, 2, 3, 4


class OptimizerSpec:
    """Hyper-parameters of a fused optimizer (device-independent description)."""

    KINDS = {"sgd": OPT_SGD, "adadelta": OPT_ADADELTA, "adam": OPT_ADAM,
             "adamw": OPT_ADAM, "adagrad": OPT_ADAGRAD, "ftrl": OPT_FTRL}

    def __init__(self, kind: str, lr: float, p1: float = 0.0, p2: float = 0.0, eps: float = 1e-7,
                 weight_decay: float = 0.0, flags: int = 0, init_s1: float = 0.0):
        kind = kind.lower()
        if kind not in self.KINDS:
            raise ValueError(f"unknown fused optimizer {kind!r}")
        self.kind, self.lr, self.p1, self.p2, self.eps = kind, float(lr), float(p1), float(p2), float(eps)
        self.weight_decay, self.flags, self.init_s1 = float(weight_decay), int(flags), float(init_s1)
        if kind == "adamw":
            self.flags |= 1

    @property
    def code(self) -> int:
        return self.KINDS[self.kind]

    @property
    def n_states(self) -> int:
        return 2 if self.kind in ("adadelta", "adam", "adamw", "ftrl") else 1

    @staticmethod
    def sgd(lr, momentum=0.0, dampening=0.0, nesterov=False, weight_decay=0.0):
        return OptimizerSpec("sgd", lr, momentum, dampening, 0.0, weight_decay, 1 if nesterov else 0)

    @staticmethod
    def adadelta(lr=1.0, rho=0.95, eps=1e-7, weight_decay=0.0):
        return OptimizerSpec("adadelta", lr, rho, 0.0, eps, weight_decay)

    @staticmethod
    def adam(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decoupled=False):
        return OptimizerSpec("adamw" if decoupled else "adam", lr, beta1, beta2, eps, weight_decay)

    @staticmethod
    def adagrad(lr=1e-2, eps=1e-10, weight_decay=0.0, initial_accumulator_value=0.0):
        return OptimizerSpec("adagrad", lr, 0.0, 0.0, eps, weight_decay, init_s1=initial_accumulator_value)

    @staticmethod
    def ftrl(lr=1e-3, l1=0.0, l2=0.0, beta=0.0, initial_accumulator_value=0.1, weight_decay=0.0):
        """FTRL-proximal with learning_rate_power = -0.5 (TF's FtrlOptimizer default): s1 = accumulator n
        (initial_accumulator_value), s2 = linear z; p1 = l1, p2 = l2, eps carries beta."""
        return OptimizerSpec("ftrl", lr, l1, l2, beta, weight_decay, init_s1=initial_accumulator_value)
