"""Checkpoint helpers for PyTorch training loops (reference: tf_yarn/pytorch/model_ckpt.py:15-77).

Files are named ``model_<epoch>.pt`` and hold ``{'model', 'optimizer', 'epoch', **extras}``.
Saving is atomic (temp file + rename) so an evaluator polling the directory never
reads a partial checkpoint.
"""
from __future__ import annotations

import logging
import os
import re
from tempfile import TemporaryDirectory
from typing import Any, Dict, Optional, Union

import torch

from tf_yarn_b200.utils import fs as filesystem

_logger = logging.getLogger(__name__)
_CKPT_RE = re.compile(r".*model_(\d+)\.pt$")


def find_latest_ckpt(model_dir: str) -> Optional[str]:
    """Path of the checkpoint with the highest epoch in ``model_dir`` (None if there is none)."""
    resolved_fs, _ = filesystem.resolve_filesystem_and_path(model_dir)
    if not resolved_fs.exists(model_dir):
        return None
    best, best_epoch = None, -1
    for path in resolved_fs.ls(model_dir):
        m = _CKPT_RE.match(path)
        if m and int(m.group(1)) > best_epoch:
            best, best_epoch = path, int(m.group(1))
    return best


def load_latest_ckpt(model_dir: str, model: torch.nn.Module, optimizer, device: Union[int, str]
                     ) -> Optional[Dict[Any, Any]]:
    latest = find_latest_ckpt(model_dir)
    if not latest:
        _logger.info("No checkpoint to load")
        return None
    return load_ckpt(latest, model, optimizer, device)


def load_ckpt(model_ckpt_path: str, model: torch.nn.Module, optimizer, device: Union[int, str]) -> Dict[Any, Any]:
    resolved_fs, _ = filesystem.resolve_filesystem_and_path(model_ckpt_path)
    _logger.info("Loading model checkpoint %s", model_ckpt_path)
    with resolved_fs.open(model_ckpt_path, "rb") as fd:
        checkpoint = torch.load(fd, map_location=torch.device(device), weights_only=False)
    _unwrap_model(model).load_state_dict(checkpoint["model"])
    if optimizer is not None and checkpoint.get("optimizer") is not None:
        optimizer.load_state_dict(checkpoint["optimizer"])
    return checkpoint


def save_ckpt(model_dir: str, model: torch.nn.Module, optimizer, epoch: int, **kwargs: Any) -> str:
    state = {
        "model": _unwrap_model(model).state_dict(),
        "optimizer": optimizer.state_dict() if optimizer is not None else None,
        "epoch": epoch,
        **kwargs,
    }
    resolved_fs, _ = filesystem.resolve_filesystem_and_path(model_dir)
    if not resolved_fs.exists(model_dir):
        resolved_fs.mkdir(model_dir)
    model_ckpt_path = os.path.join(model_dir, f"model_{epoch}.pt")
    with TemporaryDirectory() as tmpdir:
        tmp_file = os.path.join(tmpdir, f"model_{epoch}.pt")
        torch.save(state, tmp_file)
        resolved_fs.put(tmp_file, model_ckpt_path)
    return model_ckpt_path


def _unwrap_model(model: torch.nn.Module) -> torch.nn.Module:
    """Strip any data-parallel wrapper (torch DDP or tf_yarn_b200's) that exposes ``.module``."""
    from torch.nn.parallel import DistributedDataParallel as TorchDDP
    if isinstance(model, TorchDDP):
        return model.module
    if getattr(model, "_is_tfy_ddp", False):
        return model.module
    return model
