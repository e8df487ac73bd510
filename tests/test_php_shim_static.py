"""Static checks of the PHP shim (php/MatlibCuda, php/tests).

There is no PHP interpreter in this image, so the shim cannot be executed here.  These checks
catch what a parser and the FFI loader would catch first: unbalanced delimiters, an FFI call whose symbol the header
does not declare, and an FFI call whose argument count differs from the C prototype in include/rindow_b200.h.
"""
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
PHP_FILES = sorted((ROOT / "php").rglob("*.php"))
HEADER = (ROOT / "include" / "rindow_b200.h").read_text()


def strip_php(src: str) -> str:
    """PHP source with comments removed and string literals emptied (delimiters kept)."""
    out = []
    i, n = 0, len(src)
    while i < n:
        c = src[i]
        two = src[i:i + 2]
        if two == "//" or c == "#" and src[i:i + 2] != "#[":
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif two == "/*":
            j = src.find("*/", i + 2)
            i = n if j < 0 else j + 2
        elif src[i:i + 3] == "<<<":
            m = re.match(r"<<<\s*(['\"]?)(\w+)\1\n", src[i:])
            if not m:
                out.append(c)
                i += 1
                continue
            end = re.search(r"\n\s*" + m.group(2) + r"\b", src[i + m.end() - 1:])
            assert end, "unterminated heredoc"
            out.append("''")
            i = i + m.end() - 1 + end.end()
        elif c in "'\"":
            j = i + 1
            while j < n and src[j] != c:
                j += 2 if src[j] == "\\" else 1
            assert j < n, "unterminated string literal"
            out.append(c + c)
            i = j + 1
        else:
            out.append(c)
            i += 1
    return "".join(out)


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
