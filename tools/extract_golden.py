#!/usr/bin/env python3
"""Extract golden parameter sets from the reference's own PHPUnit suite.

Run in the build container only (needs /root/reference, which does not exist on the
GPU boxes); the output is committed under tests/golden/.

  python tools/extract_golden.py

Currently extracts: providerIm2col2dNormal (the 28 im2col2d parameter sets,
tests/RindowTest/Math/Matrix/LinearAlgebraTest.php:8455-8880), providerIm2col1dNormal (:9431-9612) and
providerIm2col3dNormal (:9860-10627).  The small literal
known-answer vectors of the other ops are transcribed by hand in
tests/golden/known_answers.json with a file:line citation per entry.
"""
import json
import os
import re
import sys

REF = os.environ.get("RINDOW_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "tests/RindowTest/Math/Matrix/LinearAlgebraTest.php")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def php_value(tok: str):
    tok = tok.strip()
    if tok == "null":
        return None
    if tok == "true":
        return True
    if tok == "false":
        return False
    return int(tok)


def extract_provider(text: str, name: str):
    start = text.index("function %s()" % name)
    end = text.index("public static function", start + 10)
    body = text[start:end]
    cases = []
    for m in re.finditer(r"'([^']+)'\s*=>\s*\[\[(.*?)\]\]", body, re.S):
        params = {}
        for k, v in re.findall(r"'(\w+)'\s*=>\s*([^,]+),", m.group(2)):
            params[k] = php_value(v)
        line = text[: start + m.start()].count("\n") + 1
        cases.append({"name": m.group(1), "line": line, "params": params})
    return cases


def main():
    with open(SRC) as f:
        text = f.read()
    os.makedirs(OUT, exist_ok=True)
    for provider, fname in (("providerIm2col2dNormal", "im2col2d_cases.json"),
                            ("providerIm2col1dNormal", "im2col1d_cases.json"),
                            ("providerIm2col3dNormal", "im2col3d_cases.json")):
        cases = extract_provider(text, provider)
        out = {
            "source": "tests/RindowTest/Math/Matrix/LinearAlgebraTest.php " + provider,
            "generated_by": "tools/extract_golden.py",
            "cases": cases,
        }
        with open(os.path.join(OUT, fname), "w") as f:
            json.dump(out, f, indent=1)
        print(fname, len(cases))
    return 0


if __name__ == "__main__":
    sys.exit(main())
