#!/usr/bin/env python
"""Poor man's pyflakes (no linters in the offline image): report names that a function reads as globals but that the
module never binds, and imports a module never uses.

    python tools/lint_names.py tf_yarn_b200 bench tests bench.py __graft_entry__.py
"""
import ast
import builtins
import os
import symtable
import sys


def module_files(paths):
    for p in paths:
        if os.path.isfile(p):
            yield p
        else:
            for root, _, files in os.walk(p):
                if "_ref" in root or "__pycache__" in root:
                    continue
                for f in files:
                    if f.endswith(".py"):
                        yield os.path.join(root, f)


def check(path):
    src = open(path).read()
    try:
        top = symtable.symtable(src, path, "exec")
        tree = ast.parse(src)
    except SyntaxError as e:
        return [f"{path}:{e.lineno}: syntax error {e.msg}"]
    bound = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    star = any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(tree))
    out = []

    def declared_global(tab):
        for s in tab.get_symbols():
            if s.is_declared_global() and s.is_assigned():
                bound.add(s.get_name())
        for c in tab.get_children():
            declared_global(c)
    declared_global(top)

    def walk(tab):
        for s in tab.get_symbols():
            n = s.get_name()
            if s.is_global() and s.is_referenced() and n not in bound and not hasattr(builtins, n) \
                    and n not in ("__file__", "__name__", "__doc__", "__package__", "__spec__", "__builtins__"):
                if not star:
                    out.append(f"{path}:{tab.get_lineno()}: undefined name '{n}' in {tab.get_name()}")
        for c in tab.get_children():
            walk(c)
    walk(top)
    # unused imports (module level and function level), skipping __init__ re-exports and names in __all__
    if os.path.basename(path) != "__init__.py":
        names = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name)} | \
                {n.value.id for n in ast.walk(tree) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name)}
        text_all = src
        for n in ast.walk(tree):
            if isinstance(n, (ast.Import, ast.ImportFrom)):
                if isinstance(n, ast.ImportFrom) and n.module == "__future__":
                    continue
                for a in n.names:
                    nm = (a.asname or a.name).split(".")[0]
                    if nm == "*" or nm in names:
                        continue
                    line = src.splitlines()[n.lineno - 1]
                    if "noqa" in line or f'"{nm}"' in text_all or f"'{nm}'" in text_all:
                        continue
                    out.append(f"{path}:{n.lineno}: unused import '{nm}'")
    return out


def main():
    msgs = []
    for f in sorted(set(module_files(sys.argv[1:] or ["."]))):
        msgs += check(f)
    print("\n".join(msgs) if msgs else "clean")
    return 1 if msgs else 0


if __name__ == "__main__":
    sys.exit(main())
