"""Import-time stand-in for `cluster_pack` (not installable offline), used ONLY by `bench.py --impl reference`.

The reference imports these names at module import; none of them is called on the single-box DDP path that
the bench drives (they ship python environments to HDFS).  Calling one raises."""
import sys
import types


def _unavailable(name):
    def fn(*a, **k):
        raise RuntimeError(f"cluster_pack.{name} is not available offline (bench/shims)")
    fn.__name__ = name
    return fn


class Packer:
    pass


class PythonEnvDescription:
    def __init__(self, path_to_archive="", interpreter_cmd=sys.executable, dest_path="", must_unpack=False):
        self.path_to_archive, self.interpreter_cmd = path_to_archive, interpreter_cmd
        self.dest_path, self.must_unpack = dest_path, must_unpack


for _n in ("zip_path", "upload_zip", "upload_env", "get_editable_requirements", "get_non_editable_requirements",
           "detect_packer_from_file", "get_default_fs", "get_pyenv_usage_from_archive"):
    globals()[_n] = _unavailable(_n)

packaging = types.ModuleType("cluster_pack.packaging")
packaging.PythonEnvDescription = PythonEnvDescription
packaging.Packer = Packer
sys.modules["cluster_pack.packaging"] = packaging

filesystem = types.ModuleType("cluster_pack.filesystem")
filesystem.resolve_filesystem_and_path = _unavailable("filesystem.resolve_filesystem_and_path")
sys.modules["cluster_pack.filesystem"] = filesystem
