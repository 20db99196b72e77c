"""Wide & deep click model (the BASELINE parameter-server configuration).

Not part of the reference tree: it is the BASELINE.json config that exercises
the async PS path at realistic sizes -- a Criteo-shaped input (13 numeric + 26 hashed categorical
features), embedding tables sharded over the ps ranks, a deep tower of Dense layers whose weights
are pulled (or streamed by the fused GEMM) every step, and a wide linear part.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

N_NUMERIC = 13
N_CATEGORICAL = 26


def feature_columns(vocab: int = 100_000, emb_dim: int = 64, n_cat: int = N_CATEGORICAL, n_num: int = N_NUMERIC):
    from tf_yarn_b200.estimator import feature_column as fc
    numeric = [fc.numeric_column("numeric", shape=(n_num,))]
    cats = [fc.categorical_column_with_hash_bucket(f"c{i}", vocab) for i in range(n_cat)]
    # embeddings first, numeric last: every embedding then starts at a multiple of 64 columns of the first deep
    # layer's input, which is what lets the HBM parameter server fuse the row gather into that layer's GEMM
    deep = [fc.embedding_column(c, emb_dim) for c in cats] + numeric
    wide = numeric + cats
    return wide, deep


def wide_deep_estimator(model_dir: Optional[str] = None, vocab: int = 100_000, emb_dim: int = 64,
                        hidden_units: Sequence[int] = (1024, 512, 256), optimizer=None, config=None,
                        n_cat: int = N_CATEGORICAL, n_num: int = N_NUMERIC, linear_optimizer=None):
    """``DNNLinearCombinedClassifier`` over the Criteo-shaped columns (FTRL wide tower + Adagrad deep tower)."""
    from tf_yarn_b200 import estimator as est
    from tf_yarn_b200 import keras
    wide, deep = feature_columns(vocab, emb_dim, n_cat, n_num)
    opt = optimizer if optimizer is not None else (lambda: keras.optimizers.Adagrad(0.05))
    lin = linear_optimizer if linear_optimizer is not None else (lambda: keras.optimizers.Ftrl(0.05))
    return est.DNNLinearCombinedClassifier(model_dir=model_dir, linear_feature_columns=wide, linear_optimizer=lin,
                                           dnn_feature_columns=deep, dnn_hidden_units=list(hidden_units),
                                           dnn_optimizer=opt, n_classes=2, config=config)


def synthetic_batches(batch_size: int, n_batches: int, vocab: int = 100_000, seed: int = 0,
                      n_cat: int = N_CATEGORICAL, n_num: int = N_NUMERIC) -> List[Tuple[Dict[str, torch.Tensor], torch.Tensor]]:
    """Random Criteo-shaped batches with a learnable label (depends on two features)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        numeric = torch.randn(batch_size, n_num, generator=g)
        feats: Dict[str, torch.Tensor] = {"numeric": numeric}
        for i in range(n_cat):
            feats[f"c{i}"] = torch.randint(0, vocab, (batch_size, 1), generator=g)
        label = ((numeric[:, 0] + (feats["c0"][:, 0] % 2).float() * 1.5 - 0.75) > 0).long()
        out.append((feats, label))
    return out


def input_fn_factory(batch_size: int, n_batches: int, vocab: int = 100_000, seed: int = 0, repeat: bool = True,
                     n_cat: int = N_CATEGORICAL, n_num: int = N_NUMERIC):
    from tf_yarn_b200.data import Dataset
    batches = synthetic_batches(batch_size, n_batches, vocab, seed, n_cat, n_num)

    def input_fn():
        ds = Dataset(lambda: iter(batches), len(batches))
        return ds.repeat() if repeat else ds
    return input_fn
