"""The oracle is test infrastructure: nothing under reacherdistilation_b200/ (Python or CUDA) may import, include, load or execute anything
under oracle/, and only tests/, __graft_entry__.smoke() and bench.py's CPU legs may.  Checked statically and at import time."""
import ast
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "reacherdistilation_b200")


def test_product_sources_never_mention_the_oracle_modules():
    for dirpath, _, files in os.walk(PKG):
        if os.path.basename(dirpath) in ("build", "__pycache__"):
            continue
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        names = [node.module or ""]
                    assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path
            elif f.endswith((".cu", ".cuh", ".h")):
                code = [line.split("//")[0] for line in open(path).read().splitlines()]         # comments may cite the oracle, code may not
                assert not any("oracle" in line for line in code), path


def test_importing_the_product_does_not_load_the_oracle():
    code = ("import sys; import reacherdistilation_b200 as p; "
            "import reacherdistilation_b200.mlp_train, reacherdistilation_b200.lstm_train, reacherdistilation_b200.main, "
            "reacherdistilation_b200.dataset, reacherdistilation_b200.vf_train; "
            "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; "
            "assert not bad, bad; print('clean')")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-2000:]


def test_only_the_allowed_files_import_the_oracle():
    allowed_dirs = (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"))
    allowed_files = {os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")}
    skip = {".git", "gpurun_out", "__pycache__", "build", "baseline"}
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in skip]
        for f in files:
            path = os.path.join(dirpath, f)
            if not f.endswith(".py") or path.startswith(allowed_dirs) or path in allowed_files:
                continue
            if os.path.basename(dirpath) == "scripts":           # one-off measurement scripts: not shipped, not imported by the product
                continue
            tree = ast.parse(open(path).read())
            for node in ast.walk(tree):
                if isinstance(node, ast.ImportFrom):
                    assert not (node.module or "").split(".")[0] == "oracle", path
                elif isinstance(node, ast.Import):
                    assert not any(a.name.split(".")[0] == "oracle" for a in node.names), path
