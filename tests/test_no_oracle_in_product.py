"""The product package must not import, link or execute anything under oracle/ (or any CPU
fallback): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rindow_math_matrix_b200")


def test_no_oracle_reference_in_product_sources():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        if "build" in dirpath.split(os.sep):
            continue
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) and fn != "Makefile":
                continue
            with open(os.path.join(dirpath, fn), errors="replace") as f:
                for ln, line in enumerate(f, 1):
                    if re.search(r"\b(import|from)\s+oracle\b|liboracle|rindow_oracle|oracle/", line):
                        bad.append("%s:%d: %s" % (os.path.join(dirpath, fn), ln, line.strip()))
    assert not bad, "\n".join(bad)


def test_importing_product_does_not_load_oracle():
    code = ("import sys; sys.path.insert(0, %r); import rindow_math_matrix_b200; "
            "assert 'oracle' not in sys.modules; "
            "maps = open('/proc/self/maps').read(); assert 'liboracle' not in maps; print('ok')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
