"""Static gate that runs where no linter can be installed: tools/lint_names.py (undefined globals, unused imports)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_undefined_names_or_unused_imports():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "lint_names.py"), "tf_yarn_b200", "bench", "tests",
                          "tools", "bench.py", "__graft_entry__.py"], cwd=ROOT, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout


def test_lint_catches_an_undefined_global(tmp_path):
    bad = tmp_path / "bad.py"
    bad.write_text("import os\n\n\ndef f():\n    return dist.get_rank()\n")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "lint_names.py"), str(bad)], capture_output=True,
                         text=True)
    assert res.returncode == 1
    assert "undefined name 'dist'" in res.stdout and "unused import 'os'" in res.stdout
