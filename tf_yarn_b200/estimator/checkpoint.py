"""Estimator checkpoints: ``model.ckpt-<step>`` files + a TF-style ``checkpoint`` index.

The evaluator discovers work through ``get_checkpoint_state(dir).all_model_checkpoint_paths``
(reference: tf_yarn/tensorflow/tasks/evaluator_task.py:130-136).
"""
from __future__ import annotations

import os
import re
from typing import List, NamedTuple, Optional

import torch

INDEX_FILE = "checkpoint"
PREFIX = "model.ckpt-"


class CheckpointState(NamedTuple):
    model_checkpoint_path: str
    all_model_checkpoint_paths: List[str]


def _abs(model_dir: str, p: str) -> str:
    return p if os.path.isabs(p) else os.path.join(model_dir, p)


def get_checkpoint_state(model_dir: str) -> Optional[CheckpointState]:
    index = os.path.join(model_dir, INDEX_FILE)
    if not os.path.exists(index):
        return None
    latest, all_paths = None, []
    with open(index) as f:
        for line in f:
            m = re.match(r'\s*(model_checkpoint_path|all_model_checkpoint_paths):\s*"(.*)"\s*$', line)
            if not m:
                continue
            if m.group(1) == "model_checkpoint_path":
                latest = _abs(model_dir, m.group(2))
            else:
                all_paths.append(_abs(model_dir, m.group(2)))
    if latest is None:
        return None
    return CheckpointState(latest, all_paths or [latest])


def latest_checkpoint(model_dir: str) -> Optional[str]:
    st = get_checkpoint_state(model_dir)
    return st.model_checkpoint_path if st else None


def step_of(path: str) -> int:
    return int(os.path.basename(path).split(PREFIX)[1])


def save_checkpoint(model_dir: str, step: int, payload: dict, keep_max: int = 5) -> str:
    """Atomically write ``model.ckpt-<step>`` and update the index (oldest beyond keep_max deleted)."""
    os.makedirs(model_dir, exist_ok=True)
    name = f"{PREFIX}{step}"
    path = os.path.join(model_dir, name)
    tmp = f"{path}.tmp{os.getpid()}"
    torch.save(payload, tmp)
    os.replace(tmp, path)
    st = get_checkpoint_state(model_dir)
    names = [os.path.basename(p) for p in (st.all_model_checkpoint_paths if st else [])]
    names = [n for n in names if n != name and os.path.exists(os.path.join(model_dir, n))] + [name]
    if keep_max and len(names) > keep_max:
        for old in names[:-keep_max]:
            try:
                os.remove(os.path.join(model_dir, old))
            except OSError:
                pass
        names = names[-keep_max:]
    tmp_index = os.path.join(model_dir, f"{INDEX_FILE}.tmp{os.getpid()}")
    with open(tmp_index, "w") as f:
        f.write(f'model_checkpoint_path: "{name}"\n')
        for n in names:
            f.write(f'all_model_checkpoint_paths: "{n}"\n')
    os.replace(tmp_index, os.path.join(model_dir, INDEX_FILE))
    return path


def load_checkpoint(path: str, map_location="cpu") -> dict:
    return torch.load(path, map_location=map_location, weights_only=False)
