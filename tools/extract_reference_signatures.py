#!/usr/bin/env python
"""Writes tests/golden/reference_signatures.json: the public / protected method signatures of the reference classes
the PHP shim extends (PHP checks an override against its parent when the class is loaded, so a mismatch is a fatal
error the moment the shim is used).  Run in the build container, where /root/reference exists; the fixture travels.

    python tools/extract_reference_signatures.py
"""
import json
import pathlib
import re
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1] / "tests"))
from php_signatures import class_signatures, strip_php  # noqa: E402

REF = pathlib.Path("/root/reference/src")
SOURCES = {
    "PhpMath": "Drivers/MatlibPHP/PhpMath.php",
    "PhpBlas": "Drivers/MatlibPHP/PhpBlas.php",
    "LinearAlgebra": "LinearAlgebra.php",
    "LinearAlgebraCL": "LinearAlgebraCL.php",
    "MatrixOperator": "MatrixOperator.php",
    "NDArrayCL": "NDArrayCL.php",
    "NDArrayPhp": "NDArrayPhp.php",
    "AbstractMatlibService": "Drivers/AbstractMatlibService.php",
}


def main():
    out = {}
    for cls, rel in SOURCES.items():
        sigs = class_signatures((REF / rel).read_text(), cls)
        out[cls] = {"file": "src/" + rel, "methods": sigs}
    # every class / interface / trait the reference declares under src/ (fully qualified), for the shim's `use` lines
    classes = []
    for f in sorted(REF.rglob("*.php")):
        code = strip_php(f.read_text())
        ns = re.search(r"^\s*namespace\s+([\w\\]+)\s*;", code, flags=re.M)
        for m in re.finditer(r"^\s*(?:final\s+|abstract\s+)?(?:class|interface|trait)\s+(\w+)", code, flags=re.M):
            classes.append((ns.group(1) + "\\" if ns else "") + m.group(1))
    out["__classes__"] = sorted(set(classes))
    # methods the reference parents get from their traits (protected helpers the shim calls through $this->)
    out["__traits__"] = {
        "Utils": sorted(class_signatures((REF / "Drivers/MatlibPHP/Utils.php").read_text(), "Utils")),
        "ComplexUtils": sorted(class_signatures((REF / "ComplexUtils.php").read_text(), "ComplexUtils")),
    }
    # every exception message literal of the driver / array / LA classes (the shim repeats the reference's wording)
    msgs = set()
    for rel in ("Drivers/MatlibPHP/PhpMath.php", "Drivers/MatlibPHP/PhpBlas.php", "Drivers/MatlibPHP/Utils.php",
                "Drivers/MatlibPHP/PhpBuffer.php", "NDArrayCL.php", "NDArrayPhp.php", "LinearAlgebra.php",
                "LinearAlgebraCL.php", "MatrixOperator.php", "Drivers/AbstractMatlibService.php"):
        for m in re.finditer(r"throw\s+new\s+\\?\w+\(\s*(['\"])((?:\\.|(?!\1).)*)\1", (REF / rel).read_text()):
            msgs.add(m.group(2))
    out["__messages__"] = sorted(msgs)
    dst = pathlib.Path(__file__).resolve().parents[1] / "tests" / "golden" / "reference_signatures.json"
    dst.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", dst, {k: len(v["methods"]) for k, v in out.items() if not k.startswith("__")}, len(out["__classes__"]), "classes")


if __name__ == "__main__":
    main()
