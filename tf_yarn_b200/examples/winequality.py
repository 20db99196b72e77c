"""Wine-quality inputs shared by the Estimator / Keras examples.

Same schema as the reference's helper (reference: tf_yarn/examples/winequality.py:5-49): 11
numeric features, ';'-separated CSV with a header, integer quality label in [0, 10).  The box
has no network, so :func:`ensure_dataset` writes a synthetic file with that schema when the
real ``winequality-red.csv`` is not present.
"""
import os
import zlib

from tf_yarn_b200 import data

FEATURES = ["fixed_acidity", "volatile_acidity", "citric_acid", "residual_sugar", "chlorides",
            "free_sulfur_dioxide", "total_sulfur_dioxide", "density", "pH", "sulphates", "alcohol"]
LABEL = "quality"


def ensure_dataset(path: str, n_rows: int = 1600, seed: int = 0) -> str:
    """Create a synthetic wine-quality CSV at ``path`` if it does not exist."""
    if os.path.exists(path):
        return path
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_rows, len(FEATURES), generator=g)
    w = torch.randn(len(FEATURES), generator=g)
    score = (x @ w) / w.norm() + 0.3 * torch.randn(n_rows, generator=g)
    label = (score * 1.5 + 5.5).round().clamp(3, 8).long()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write(";".join(f'"{c}"' for c in FEATURES + [LABEL]) + "\n")
        for row, y in zip(x.tolist(), label.tolist()):
            f.write(";".join(f"{v:.5f}" for v in row) + f";{y}\n")
    return path


def get_dataset(path: str, train_fraction: float = 0.7, split: str = "train") -> data.Dataset:
    """Rows as ``({feature: value}, label)``; a stable hash of the row picks the train / test side."""
    def split_label(*row):
        return dict(zip(FEATURES, row[:-1])), row[-1]

    def in_training_set(*row):
        key = "".join(repr(v) for v in row).encode()
        return (zlib.crc32(key) % 1000) < int(train_fraction * 1000)

    rows = data.CsvDataset(path, [0.0] * len(FEATURES) + [0], header=True, field_delim=";")
    if split == "train":
        return rows.filter(in_training_set).map(split_label)
    if split == "test":
        return rows.filter(lambda *row: not in_training_set(*row)).map(split_label)
    raise ValueError("Unknown option split, must be 'train' or 'test'")


def get_feature_columns():
    from tf_yarn_b200.estimator import feature_column as fc
    return [fc.numeric_column(name) for name in FEATURES]


def get_n_classes() -> int:
    return 10
