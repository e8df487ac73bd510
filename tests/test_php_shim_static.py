"""Static checks of the PHP shim (php/MatlibCuda, php/tests).

There is no PHP interpreter in this image, so the shim cannot be executed here.  These checks
catch what a parser and the FFI loader would catch first: unbalanced delimiters, an FFI call whose symbol the header
does not declare, and an FFI call whose argument count differs from the C prototype in include/rindow_b200.h.
"""
import json
import pathlib
import re

import pytest

from php_signatures import class_signatures, strip_php

ROOT = pathlib.Path(__file__).resolve().parents[1]
PHP_FILES = sorted((ROOT / "php").rglob("*.php"))
HEADER = (ROOT / "include" / "rindow_b200.h").read_text()


def c_prototypes() -> dict:
    """symbol -> number of parameters, from the header."""
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"\b(rb200_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        protos[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return protos


def call_arity(code: str, open_paren: int) -> int:
    depth, commas, i = 0, 0, open_paren
    empty = True
    while i < len(code):
        c = code[i]
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
            if depth == 0:
                return 0 if empty else commas + 1
        elif depth == 1:
            if c == ",":
                commas += 1
            elif not c.isspace():
                empty = False
        if depth > 1 and not c.isspace():
            empty = False
        i += 1
    raise AssertionError("unterminated call")


def test_shim_files_exist():
    names = {p.name for p in PHP_FILES}
    for want in ("MatlibCuda.php", "CudaFFI.php", "CudaBuffer.php", "CudaBlas.php", "CudaMath.php", "NDArrayCuda.php",
                 "LinearAlgebraCuda.php", "CudaDeviceBuffer.php", "CudaQueue.php"):
        assert want in names


@pytest.mark.parametrize("path", PHP_FILES, ids=lambda p: str(p.relative_to(ROOT)))
def test_php_delimiters_balance(path):
    src = path.read_text()
    assert src.lstrip().startswith("<?php")
    code = strip_php(src)
    pairs = {")": "(", "]": "[", "}": "{"}
    stack = []
    line = 1
    for ch in code:
        if ch == "\n":
            line += 1
        elif ch in "([{":
            stack.append((ch, line))
        elif ch in pairs:
            assert stack, f"{path.name}:{line}: unmatched '{ch}'"
            top, at = stack.pop()
            assert top == pairs[ch], f"{path.name}:{line}: '{ch}' closes '{top}' opened at line {at}"
    assert not stack, f"{path.name}: unclosed '{stack[-1][0]}' from line {stack[-1][1]}"


@pytest.mark.parametrize("path", PHP_FILES, ids=lambda p: str(p.relative_to(ROOT)))
def test_php_statements_are_terminated(path):
    """A line that ends a call or an assignment and is followed by a new statement keyword must end in ';'."""
    code = strip_php(path.read_text())
    lines = code.split("\n")
    starters = re.compile(r"^\s*(return|if|foreach|for|while|throw|\$\w+\s*=[^=]|CudaFFI::check)\b")
    for k in range(len(lines) - 1):
        cur = lines[k].rstrip()
        if not cur or not starters.match(lines[k + 1]):
            continue
        if cur.endswith(")") and not re.match(r"^\s*(if|foreach|for|while|elseif|}\s*elseif|function|public|private|protected|static|switch|catch|}\s*catch)\b", cur.lstrip()) \
                and cur.count("(") == cur.count(")") and re.match(r"^\s*(\$|return|throw|CudaFFI::|parent::|self::|static::)", cur):
            raise AssertionError(f"{path.name}:{k + 1}: statement not terminated: {cur.strip()}")


def test_ffi_calls_match_the_c_prototypes():
    protos = c_prototypes()
    assert len(protos) > 80
    seen = set()
    for path in PHP_FILES:
        if path.name == "CudaFFI.php":
            continue                        # the CDEF itself is pinned by test_capi_symbols.py
        code = strip_php(path.read_text())
        for m in re.finditer(r"->\s*(rb200_\w+)\s*\(", code):
            name = m.group(1)
            line = code.count("\n", 0, m.start()) + 1
            assert name in protos, f"{path.name}:{line}: {name} is not declared in include/rindow_b200.h"
            got = call_arity(code, m.end() - 1)
            assert got == protos[name], f"{path.name}:{line}: {name} called with {got} arguments, prototype has {protos[name]}"
            seen.add(name)
    # the hot path of SURVEY.md section 8(a) must be reachable from the shim
    for hot in ("rb200_sgemm", "rb200_dgemm", "rb200_add", "rb200_multiply", "rb200_reduce_sum", "rb200_reduce_max",
                "rb200_reduce_argmax", "rb200_softmax", "rb200_im2col2d", "rb200_conv2d_forward", "rb200_buffer_alloc",
                "rb200_buffer_free"):
        assert hot in seen, f"{hot} is never called from php/MatlibCuda"


def test_php_classes_match_their_files():
    for path in PHP_FILES:
        code = strip_php(path.read_text())
        m = re.search(r"^\s*(?:final\s+|abstract\s+)?(?:class|interface|trait)\s+(\w+)", code, flags=re.M)
        assert m, f"{path.name}: no class"
        assert m.group(1) == path.stem, f"{path.name}: declares {m.group(1)}"
        assert re.search(r"^\s*namespace\s+[\w\\]+\s*;", code, flags=re.M), f"{path.name}: no namespace"


def _shim_classes():
    """(file, class, parent) of every shim class that extends something."""
    out = []
    for path in PHP_FILES:
        code = strip_php(path.read_text())
        for m in re.finditer(r"\bclass\s+(\w+)\s+extends\s+([\w\\\\]+)", code):
            out.append((path, m.group(1), m.group(2).split("\\")[-1]))
    return out


def _type_accepts(child: str, parent: str) -> bool:
    """Parameter types are contravariant: the override must accept everything the parent does."""
    if child == "" or child == "mixed" or child == parent:
        return True
    if parent == "":
        return False
    return set(parent.split("|")) <= set(child.split("|"))


def test_overrides_are_compatible_with_the_reference_parents():
    """PHP verifies an overriding method against its parent when the class is loaded (arity, optional parameters,
    static-ness, visibility, parameter and return types); the reference callers also pass every argument by position
    (src/LinearAlgebra.php:2006 etc.).  The parents' signatures are the fixture tests/golden/reference_signatures.json
    (tools/extract_reference_signatures.py)."""
    ref = json.loads((ROOT / "tests" / "golden" / "reference_signatures.json").read_text())
    own = {}
    for path in PHP_FILES:
        code = path.read_text()
        for m in re.finditer(r"\bclass\s+(\w+)", strip_php(code)):
            own[m.group(1)] = class_signatures(code, m.group(1))
    rank = {"private": 0, "protected": 1, "public": 2}
    checked = 0
    for path, cls, parent in _shim_classes():
        parents = ref[parent]["methods"] if parent in ref else own.get(parent)
        if parents is None:
            continue
        # walk further up when the parent is itself a shim class or a reference class with a known parent
        chain = [parents]
        if parent == "NDArrayCL" or parent == "LinearAlgebraCL":
            pass
        for name, sig in own[cls].items():
            if name == "__construct":
                continue
            for table in chain:
                if name not in table:
                    continue
                psig = table[name]
                if psig["visibility"] == "private":
                    continue
                where = "%s::%s overrides %s::%s" % (cls, name, parent, name)
                assert sig["static"] == psig["static"], where + ": static differs"
                assert rank[sig["visibility"]] >= rank[psig["visibility"]], where + ": narrower visibility"
                assert len(sig["params"]) >= len(psig["params"]), where + ": fewer parameters"
                for i, pp in enumerate(psig["params"]):
                    cp = sig["params"][i]
                    assert cp["by_ref"] == pp["by_ref"], "%s: parameter %d by-reference differs" % (where, i + 1)
                    assert cp["variadic"] == pp["variadic"], "%s: parameter %d variadic differs" % (where, i + 1)
                    assert _type_accepts(cp["type"], pp["type"]), \
                        "%s: parameter %d (%s) takes %s, parent takes %s" % (where, i + 1, cp["name"], cp["type"], pp["type"])
                    if pp["default"]:
                        assert cp["default"], "%s: parameter %d (%s) must stay optional" % (where, i + 1, cp["name"])
                for cp in sig["params"][len(psig["params"]):]:
                    assert cp["default"] or cp["variadic"], "%s: extra parameter $%s must be optional" % (where, cp["name"])
                if psig["returns"]:
                    assert sig["returns"], where + ": return type dropped"
                    child_r, parent_r = set(sig["returns"].split("|")), set(psig["returns"].split("|"))
                    ok = child_r <= parent_r or "mixed" in parent_r or (child_r - {"static", "self", cls}) <= parent_r | {parent}
                    assert ok, "%s: returns %s, parent returns %s" % (where, sig["returns"], psig["returns"])
                checked += 1
    assert checked >= 80, checked


def test_hot_path_methods_keep_the_reference_parameter_order():
    """The reference calls its drivers positionally; the overriding hot-path methods keep the parent's parameter
    names in the parent's order (names are not checked by PHP, but a swapped pair of ints would pass every type check)."""
    ref = json.loads((ROOT / "tests" / "golden" / "reference_signatures.json").read_text())
    math = class_signatures((ROOT / "php" / "MatlibCuda" / "CudaMath.php").read_text(), "CudaMath")
    blas = class_signatures((ROOT / "php" / "MatlibCuda" / "CudaBlas.php").read_text(), "CudaBlas")
    for own, parent, names in ((math, "PhpMath", ("add", "multiply", "reduceSum", "reduceMax", "reduceArgMax", "softmax",
                                                  "im2col2d")),
                               (blas, "PhpBlas", ("gemm", "gemv", "axpy", "copy", "scal"))):
        for name in names:
            assert name in own, name
            want = [p["name"].lower() for p in ref[parent]["methods"][name]["params"]]
            got = [p["name"].lower() for p in own[name]["params"]][:len(want)]
            assert got == want, "%s: %s vs reference %s" % (name, got, want)


# interfaces of interop-phpobjects/polite-math the reference itself imports (grep `use Interop\\Polite` over src/)
POLITE = {"NDArray", "Buffer", "LinearBuffer", "DeviceBuffer", "BLAS", "OpenCL"}
PHP_BUILTIN = {"FFI", "ArrayAccess", "Countable", "IteratorAggregate", "Traversable", "ArrayIterator", "ArrayObject",
               "InvalidArgumentException", "RuntimeException", "LogicException", "OutOfRangeException", "TypeError",
               "DomainException", "Throwable", "Exception", "stdClass", "Serializable"}


def _declared_shim_classes():
    out = {}
    for path in PHP_FILES:
        code = strip_php(path.read_text())
        ns = re.search(r"^\s*namespace\s+([\w\\]+)\s*;", code, flags=re.M).group(1)
        for m in re.finditer(r"^\s*(?:final\s+|abstract\s+)?(?:class|interface|trait)\s+(\w+)", code, flags=re.M):
            out[ns + "\\" + m.group(1)] = path
    return out


def test_every_use_line_resolves():
    ref = json.loads((ROOT / "tests" / "golden" / "reference_signatures.json").read_text())
    known = set(ref["__classes__"]) | set(_declared_shim_classes())
    for path in PHP_FILES:
        code = strip_php(path.read_text())
        for m in re.finditer(r"^use\s+(function\s+)?([\w\\]+)(?:\s+as\s+\w+)?\s*;", code, flags=re.M):
            name = m.group(2).lstrip("\\")
            if m.group(1):
                assert name in ("Rindow\\Math\\Matrix\\R", "Rindow\\Math\\Matrix\\C"), f"{path.name}: use function {name}"
            elif name.startswith("Interop\\Polite\\Math\\Matrix\\"):
                assert name.split("\\")[-1] in POLITE, f"{path.name}: {name} is not a polite-math interface the reference uses"
            elif name.startswith("PHPUnit\\") or name.startswith("RindowTest\\"):
                assert "tests" in path.parts, f"{path.name}: test-only import {name}"
            elif "\\" not in name:
                assert name in PHP_BUILTIN, f"{path.name}: unknown global class {name}"
            else:
                assert name in known, f"{path.name}: use {name} resolves to nothing in the reference or the shim"


def test_this_and_parent_calls_resolve_to_a_method():
    """`$this->name(` / `parent::name(` / `self::name(` / `static::name(` must name a method of the class, of a shim
    parent, or of the reference parent (fixture) -- what PHP reports as 'Call to undefined method' at run time."""
    ref = json.loads((ROOT / "tests" / "golden" / "reference_signatures.json").read_text())
    own, parent_of = {}, {}
    for path in PHP_FILES:
        src = path.read_text()
        code = strip_php(src)
        for m in re.finditer(r"\bclass\s+(\w+)(?:\s+extends\s+([\w\\]+))?", code):
            own[m.group(1)] = (class_signatures(src, m.group(1)), path)
            parent_of[m.group(1)] = m.group(2).split("\\")[-1] if m.group(2) else None

    def methods_of(cls):
        names, open_ended = set(), False
        while cls:
            if cls in own:
                names |= set(own[cls][0])
                cls = parent_of[cls]
            elif cls in ref:
                names |= set(ref[cls]["methods"])
                # the reference parents whose own parents / traits are outside the fixture
                open_ended = cls in ("NDArrayCL", "AbstractMatlibService", "PhpBlas", "PhpMath", "LinearAlgebra", "MatrixOperator")
                cls = None
            else:
                open_ended = True           # PHPUnit TestCase, reference test suites
                cls = None
        return names, open_ended

    def signature_of(cls, name):
        while cls:
            if cls in own:
                if name in own[cls][0]:
                    return own[cls][0][name]
                cls = parent_of[cls]
            elif cls in ref:
                return ref[cls]["methods"].get(name)
            else:
                return None
        return None

    # traits the reference parents pull in (src/Drivers/MatlibPHP/Utils.php, src/ComplexUtils.php)
    trait_methods = set(ref["__traits__"]["Utils"]) | set(ref["__traits__"]["ComplexUtils"])
    checked = 0
    for cls, (sigs, path) in own.items():
        names, open_ended = methods_of(cls)
        body = strip_php(path.read_text())
        for m in re.finditer(r"(\$this->|parent::|self::|static::)(\w+)\s*\(", body):
            name = m.group(2)
            if name in names or name in trait_methods:
                checked += 1
                sig = signature_of(cls, name)
                if sig is not None:
                    got = call_arity(body, m.end() - 1)
                    required = sum(1 for q in sig["params"] if not q["default"] and not q["variadic"])
                    most = 10 ** 6 if any(q["variadic"] for q in sig["params"]) else len(sig["params"])
                    line = body.count("\n", 0, m.start()) + 1
                    assert required <= got <= most, \
                        f"{path.name}:{line}: {m.group(1)}{name}() called with {got} arguments, takes {required}..{most}"
                continue
            if m.group(1) == "$this->" and re.search(r"\$" + name + r"\b", body):
                continue                     # a callable property
            line = body.count("\n", 0, m.start()) + 1
            if open_ended and "tests" in path.parts:
                continue                     # PHPUnit assertions and helpers of the reference suites
            assert not True, f"{path.name}:{line}: {m.group(1)}{name}() is not a method of {cls} or its parents"
    assert checked > 100, checked


def _call_args(code: str, open_paren: int):
    """Top-level argument strings of the call whose '(' is at open_paren."""
    depth, cur, args = 0, [], []
    for i in range(open_paren, len(code)):
        c = code[i]
        if c in "([{":
            depth += 1
            if depth == 1:
                continue
        elif c in ")]}":
            depth -= 1
            if depth == 0:
                tail = "".join(cur).strip()
                if tail:
                    args.append(tail)
                return args
        if c == "," and depth == 1:
            args.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
    raise AssertionError("unterminated call")


def test_named_arguments_of_constructor_calls_exist():
    """`new NDArrayPhp($v, dtype:$d, service:$s)`: a misspelt named argument is an Error at run time."""
    ref = json.loads((ROOT / "tests" / "golden" / "reference_signatures.json").read_text())
    ctor = {cls: v["methods"].get("__construct") for cls, v in ref.items() if not cls.startswith("__")}
    parent_of = {}
    for path in PHP_FILES:
        src = path.read_text()
        code = strip_php(src)
        for m in re.finditer(r"\bclass\s+(\w+)(?:\s+extends\s+([\w\\]+))?", code):
            sigs = class_signatures(src, m.group(1))
            parent_of[m.group(1)] = m.group(2).split("\\")[-1] if m.group(2) else None
            ctor[m.group(1)] = sigs.get("__construct")
    for cls in list(parent_of):                  # inherited constructors (NDArrayCuda -> NDArrayCL)
        c = cls
        while ctor.get(cls) is None and parent_of.get(c):
            c = parent_of[c]
            if ctor.get(c) is not None:
                ctor[cls] = ctor[c]
    checked = 0
    for path in PHP_FILES:
        code = strip_php(path.read_text())
        for m in re.finditer(r"\bnew\s+\\?([\w\\]+)\s*\(", code):
            cls = m.group(1).split("\\")[-1]
            sig = ctor.get(cls)
            if sig is None:
                continue
            names = [p["name"] for p in sig["params"]]
            line = code.count("\n", 0, m.start()) + 1
            args = _call_args(code, m.end() - 1)
            positional = 0
            for a in args:
                nm = re.match(r"^(\w+)\s*:(?!:)", a)
                if nm:
                    assert nm.group(1) in names, f"{path.name}:{line}: new {cls}(... {nm.group(1)}: ...) -- parameters are {names}"
                    assert names.index(nm.group(1)) >= positional, f"{path.name}:{line}: {nm.group(1)} given twice"
                    checked += 1
                else:
                    positional += 1
            assert positional <= len(names) or any(p["variadic"] for p in sig["params"]), f"{path.name}:{line}: too many arguments to new {cls}"
    assert checked >= 10, checked


# wording that belongs to this driver only (no counterpart in the reference): device-side constraints and diagnostics
DRIVER_ONLY_MESSAGES = {
    "Index invalid or out of range", "Illegal Operation",                                   # PhpBuffer raises through PHP itself
    "dtype is required", "bytes is not a multiple of the value size",
    "CudaDeviceBuffer has no host mirror: use read()", "CudaDeviceBuffer has no host mirror",
    "copy source must be a buffer of this driver: ", "librindow_b200 status ",
    "librindow_b200.so / B200 not available",
    "LinearAlgebraCuda returns scalars as numbers (return types of LinearAlgebra).",
    # entry points the reference does not have (gemmStridedBatched, conv2dForward, reduceArgMaxPartial)
    "gemmStridedBatched needs device buffers", "kernel buffer size is invalid", "bias buffer size is invalid",
    "reduceArgMaxPartial needs device buffers", "m, n and k must be greater than 0.",
    "records must be an int64 buffer of two words per output",
    # bounds the reference leaves to PHP's own offset errors; a kernel must not be launched out of range
    "Argument m / n must be greater than 0.", "Argument n / inc must be greater than 0.",
    "Argument m / inc must be greater than 0.", "Matrix specification too large for bufferA.",
    "Matrix specification too large for bufferB.", "Matrix specification too large for buffer.",
    "Vector specification too large for bufferX.", "Vector specification too large for bufferY.",
}


def test_exception_wording_is_the_references_or_declared_driver_only():
    """Callers (and the reference's tests) match on exception messages: a message thrown by the shim is either one the
    reference throws (fixture: every literal of PhpMath / PhpBlas / Utils / PhpBuffer / NDArrayCL / LinearAlgebra(CL) /
    MatrixOperator) or listed above as driver-only."""
    ref = json.loads((ROOT / "tests" / "golden" / "reference_signatures.json").read_text())
    known = set(ref["__messages__"])
    seen_ref = 0
    for path in sorted((ROOT / "php" / "MatlibCuda").glob("*.php")):
        for m in re.finditer(r"throw\s+new\s+\\?\w+\(\s*(['\"])((?:\\.|(?!\1).)*)\1", path.read_text()):
            msg = m.group(2)
            if msg in known:
                seen_ref += 1
                continue
            assert msg in DRIVER_ONLY_MESSAGES, f"{path.name}: '{msg}' is neither the reference's wording nor declared driver-only"
    assert seen_ref >= 60, seen_ref


# ---- the PHP calls pass the same arguments, in the same order, as the ctypes mirror (which the GPU suite runs) --------

def _norm_php_arg(a: str) -> str:
    a = a.replace(" ", "")
    a = re.sub(r"\\?FFI::addr\((.*)\)$", r"&\1", a)
    a = re.sub(r"(?:\$ffi|CudaFFI::lib\(\))->cast\('',(.*)\)$", r"\1", a)
    a = re.sub(r"^\(?(.*?)\)?\?1:0$", r"\1", a)
    a = a.replace("$this->", "self.").replace("$", "")
    a = a.replace("->abiDtype()", ".dtype").replace("->devicePtrRW()", ".ptr").replace("->devicePtr()", ".ptr").replace("::", ".")
    return a.lower().replace("_", "")


def _norm_py_arg(a: str) -> str:
    a = " ".join(a.split())
    for pat, rep in ((r"^1 if (.*) else 0$", r"\1"), (r"^(?:float|int|bool)\((.*)\)$", r"\1"),
                     (r"^ctypes\.byref\((.*)\)$", r"&\1"), (r"^_cuda\(\"\w+\", (\w+)\)\.dtype\(\)$", r"\1.dtype"),
                     (r"^self\._float_dtype\((\w+), .*\)$", r"\1.dtype")):
        a = re.sub(pat, rep, a)
    a = a.replace(" ", "").replace(".dtype()", ".dtype").replace(".device_ptr_rw()", ".ptr").replace(".device_ptr()", ".ptr")
    return a.lower().replace("_", "")


# local variable names that differ between the two hosts for the same thing
_LOCAL_ALIASES = {"a": "aptr", "b": "bptr", "c": "cptr", "y": "yptr", "im": "imptr", "co": "coptr", "x.dtype": "dt"}


def test_php_ffi_calls_pass_what_the_ctypes_mirror_passes():
    """The ctypes host (`rindow_math_matrix_b200/cuda_blas.py`, `cuda_math.py`) is what the GPU suite executes; the PHP
    host cannot run here.  For the compute entry points the two must hand the SAME expressions to the SAME positions --
    the one mistake neither the prototype check nor PHP's type system would see is two swapped integers."""
    php, py = {}, {}
    for path in sorted((ROOT / "php" / "MatlibCuda").glob("*.php")):
        if path.name in ("CudaFFI.php", "CudaBuffer.php", "CudaDeviceBuffer.php", "CudaContext.php", "CudaEventList.php",
                         "CudaQueue.php"):
            continue                                   # buffer / event plumbing is structured differently per host
        code = strip_php(path.read_text())
        for m in re.finditer(r"->\s*(rb200_\w+)\s*\(", code):
            php.setdefault(m.group(1), []).append([_norm_php_arg(x) for x in _call_args(code, m.end() - 1)])
    for path in sorted((ROOT / "rindow_math_matrix_b200").glob("cuda_*.py")):
        src = path.read_text()
        for m in re.finditer(r"\.(rb200_\w+)\s*\(", src):
            py.setdefault(m.group(1), []).append([_norm_py_arg(x) for x in _call_args(src, m.end() - 1)])
    must_match = ("rb200_sgemm", "rb200_dgemm", "rb200_add", "rb200_multiply", "rb200_reduce_sum", "rb200_reduce_max",
                  "rb200_reduce_argmax", "rb200_softmax", "rb200_im2col2d", "rb200_im2col3d", "rb200_gemv", "rb200_axpy",
                  "rb200_copy", "rb200_scal", "rb200_swap", "rb200_dot",      # asum / nrm2 / iamax / iamin: `->$fn(` in PHP
                  "rb200_sum", "rb200_imax", "rb200_imin", "rb200_topk", "rb200_cumsum", "rb200_cumsumb",
                  "rb200_searchsorted", "rb200_bandpart", "rb200_masking", "rb200_repeat", "rb200_imagecopy",
                  "rb200_onehot", "rb200_reduce_argmax_partial", "rb200_slice", "rb200_gatherb", "rb200_einsum")
    matched = 0
    for sym in must_match:
        assert sym in php and sym in py, sym
        for a in php[sym]:
            ok = False
            for b in py[sym]:
                if len(a) == len(b) and all(x == y or _LOCAL_ALIASES.get(x) == y for x, y in zip(a, b)):
                    ok = True
            assert ok, "%s: PHP passes %s, the ctypes mirror passes %s" % (sym, a, py[sym])
            matched += 1
    assert matched >= len(must_match)


def test_php_dtype_table_matches_the_header_enum_and_the_mirror():
    """CudaBuffer::table() is the ONE place polite-math dtype constants meet ABI dtype codes: code and element size per
    dtype must equal include/rindow_b200.h and rindow_math_matrix_b200/dtypes.py."""
    from rindow_math_matrix_b200 import dtypes
    enum = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"#define\s+RB200_([A-Z0-9]+)\s+(\d+)\b", HEADER)}
    src = (ROOT / "php" / "MatlibCuda" / "CudaBuffer.php").read_text()
    rows = re.findall(r"NDArray::(\w+)\s*=>\s*\[(\d+),\s*(\d+),\s*'(\w+)'\]", src)
    assert len(rows) == 13
    ctype_bytes = {"uint8_t": 1, "int8_t": 1, "int16_t": 2, "uint16_t": 2, "int32_t": 4, "uint32_t": 4, "int64_t": 8,
                   "uint64_t": 8, "float": 4, "double": 8}
    for name, code, size, ctype in rows:
        code, size = int(code), int(size)
        assert enum[name] == code, name
        py = getattr(dtypes, "bool_" if name == "bool" else name)
        assert py == code and dtypes.VALUE_SIZE[py] == size, name
        assert ctype_bytes[ctype] * (2 if name.startswith("complex") else 1) == size, name
