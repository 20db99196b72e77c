"""Import-time stand-in for `skein` (not installable offline), used ONLY by `bench.py --impl reference`.

The reference (baseline/_ref/tf_yarn) talks to skein's ApplicationMaster key-value store:
`skein.ApplicationClient.from_current().kv` with `kv.wait(key)` and `kv[key] = bytes`
(reference: tf_yarn/event.py:13-18,70-79, tf_yarn/_task_commons.py:37-53).  Here `kv` is this repo's
KVClient connected to the KV server that rank 0 of the bench hosts (address in $TFY_REF_KV_ADDR).
Only the control plane goes through it; see bench/shims/README.md.
"""
import os
import sys
import types

__version__ = "0.8.2+tfy-shim"


class _Resources:
    """skein.model.Resources(memory, vcores): memory in MiB or a '2 GiB' string."""

    _UNITS = {"": 1, "mib": 1, "mb": 1, "m": 1, "gib": 1024, "gb": 1024, "g": 1024, "kib": 1 / 1024, "tib": 1 << 20}

    def __init__(self, memory, vcores, gpus=0, fpgas=0):
        if isinstance(memory, str):
            s = memory.strip().lower()
            num = s.rstrip("abcdefghijklmnopqrstuvwxyz ")
            memory = int(float(num) * self._UNITS[s[len(num):].strip()])
        self.memory, self.vcores, self.gpus, self.fpgas = int(memory), int(vcores), gpus, fpgas


class _Placeholder:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs
        self.__dict__.update(kwargs)


class ApplicationClient:
    def __init__(self, address=None, app_id="application_tfy_ref_0001"):
        from tf_yarn_b200.kv import KVClient      # control plane only
        self.address = address or os.environ["TFY_REF_KV_ADDR"]
        self.id = app_id
        self.kv = KVClient(self.address)

    @classmethod
    def from_current(cls):
        return cls()

    def get_containers(self, *a, **k):
        return []

    def shutdown(self, *a, **k):
        return None


class Client(_Placeholder):
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class Service(_Placeholder):
    pass


class ApplicationSpec(_Placeholder):
    pass


class Master(_Placeholder):
    pass


class File(_Placeholder):
    pass


class Security(_Placeholder):
    pass


def _submodule(name, **attrs):
    m = types.ModuleType(f"skein.{name}")
    m.__dict__.update(attrs)
    sys.modules[f"skein.{name}"] = m
    return m


class SkeinError(Exception):
    pass


class _ConnectionError(SkeinError, ConnectionError):
    pass


class _FinalStatus:
    SUCCEEDED, FAILED, KILLED, UNDEFINED = "SUCCEEDED", "FAILED", "KILLED", "UNDEFINED"


exceptions = _submodule("exceptions", SkeinError=SkeinError, ConnectionError=_ConnectionError)
model = _submodule("model", Resources=_Resources, FinalStatus=_FinalStatus, ApplicationReport=_Placeholder,
                   ACLs=_Placeholder, ApplicationLogs=dict, Service=Service, ApplicationSpec=ApplicationSpec,
                   File=File, Master=Master, Security=Security)
kv = _submodule("kv", KeyValueStore=_Placeholder)
