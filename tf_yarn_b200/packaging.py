"""Packaging helpers (reference: tf_yarn/packaging.py, deprecated pass-throughs to cluster_pack).

On one box there is nothing to ship: tasks run with the launcher's interpreter and see the launcher's
file system.  The functions keep their names so that user scripts importing them keep working.
"""
from __future__ import annotations

import os
import shutil
import sys
import warnings
import zipfile
from typing import Dict, Optional, Tuple


def zip_path(py_dir: str, include_base_name: bool = True, tmp_dir: Optional[str] = None) -> str:
    """Zip a directory (used by the reference to ship code); returns the archive path."""
    import tempfile
    tmp_dir = tmp_dir or tempfile.mkdtemp()
    os.makedirs(tmp_dir, exist_ok=True)
    base = os.path.basename(os.path.normpath(py_dir))
    out = os.path.join(tmp_dir, base + ".zip")
    with zipfile.ZipFile(out, "w", zipfile.ZIP_DEFLATED) as zf:
        for root, _, files in os.walk(py_dir):
            for name in files:
                full = os.path.join(root, name)
                rel = os.path.relpath(full, os.path.dirname(py_dir) if include_base_name else py_dir)
                zf.write(full, rel)
    return out


def upload_env(*args, **kwargs) -> Tuple[str, str]:
    warnings.warn("upload_env is a no-op on a single box: tasks use the current interpreter", DeprecationWarning)
    return sys.executable, os.path.basename(sys.prefix)


upload_env_to_hdfs = upload_env


def upload_zip(zip_file: str, package_path: Optional[str] = None, **_ignored) -> str:
    if package_path and os.path.abspath(package_path) != os.path.abspath(zip_file):
        os.makedirs(os.path.dirname(os.path.abspath(package_path)), exist_ok=True)
        shutil.copyfile(zip_file, package_path)
        return package_path
    return zip_file


def get_editable_requirements(*args, **kwargs) -> Dict[str, str]:
    """Editable installs are already importable by the tasks (same interpreter, same file system)."""
    return {}


def get_default_fs() -> str:
    return "file://"


def detect_archive_names(*args, **kwargs):
    return sys.executable, os.path.basename(sys.prefix), ""
