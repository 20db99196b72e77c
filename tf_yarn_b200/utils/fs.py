"""Filesystem shim: the subset of ``cluster_pack.filesystem`` the framework needs.

The reference resolves local / HDFS / viewfs paths through cluster_pack
(reference: tf_yarn/pytorch/model_ckpt.py:19-72, parquet_dataset.py:21-36,
pytorch/tasks/worker.py:145-152).  On a single box the storage is the local
file system (NVMe / shared mounts); ``file://`` URIs and plain paths are
accepted, other schemes raise a clear error.
"""
from __future__ import annotations

import os
import shutil
from typing import IO, List, Tuple


class LocalFileSystem:
    """Local file system with the cluster_pack ``EnhancedFileSystem`` verbs."""

    @property
    def base_fs(self) -> "LocalFileSystem":
        return self

    @staticmethod
    def _p(path: str) -> str:
        return path[len("file://"):] if path.startswith("file://") else path

    def exists(self, path: str) -> bool:
        return os.path.exists(self._p(path))

    def isdir(self, path: str) -> bool:
        return os.path.isdir(self._p(path))

    def ls(self, path: str, recursive: bool = False) -> List[str]:
        root = self._p(path)
        if recursive:
            out = []
            for d, _, files in os.walk(root):
                out.extend(os.path.join(d, f) for f in files)
            return sorted(out)
        return sorted(os.path.join(root, name) for name in os.listdir(root))

    def mkdir(self, path: str) -> None:
        os.makedirs(self._p(path), exist_ok=True)

    mkdirs = mkdir

    def open(self, path: str, mode: str = "rb") -> IO:
        return open(self._p(path), mode)

    def put(self, local_path: str, remote_path: str) -> None:
        """Atomic publish: copy next to the destination, then rename (readers never see partial files)."""
        dst = self._p(remote_path)
        os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
        tmp = f"{dst}.tmp{os.getpid()}"
        shutil.copyfile(local_path, tmp)
        os.replace(tmp, dst)

    def get(self, remote_path: str, local_path: str) -> None:
        shutil.copyfile(self._p(remote_path), local_path)

    def rm(self, path: str, recursive: bool = False) -> None:
        p = self._p(path)
        if os.path.isdir(p):
            shutil.rmtree(p) if recursive else os.rmdir(p)
        elif os.path.exists(p):
            os.remove(p)


def resolve_filesystem_and_path(path: str) -> Tuple[LocalFileSystem, str]:
    if "://" in path and not path.startswith("file://"):
        raise ValueError(f"unsupported filesystem for {path!r}: tf_yarn_b200 runs on one box and uses local storage")
    return LocalFileSystem(), LocalFileSystem._p(path)
