import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_reports_unavailable():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line and "skein" in line["unavailable"]


def test_graft_entry_build_is_idempotent():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py")], capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stderr
    assert os.path.exists(os.path.join(ROOT, "tf_yarn_b200", "ops", "lib", "libtfy_b200.so"))
    assert os.path.exists(os.path.join(ROOT, "tf_yarn_b200", "kv", "libtfy_kv.so"))
