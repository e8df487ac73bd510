"""A small reader of PHP method signatures (no PHP interpreter in this image): used by
tools/extract_reference_signatures.py to build tests/golden/reference_signatures.json and by
tests/test_php_shim_static.py to read the shim's own classes."""
import re


def strip_php(src: str) -> str:
    """PHP source with comments removed and string literals emptied (delimiters kept)."""
    out = []
    i, n = 0, len(src)
    while i < n:
        c = src[i]
        two = src[i:i + 2]
        if two == "//" or (c == "#" and two != "#["):
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


def _split_top(s: str):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        parts.append(tail)
    return [p.strip() for p in parts if p.strip()]


def _norm_type(t: str) -> str:
    t = t.replace(" ", "")
    if not t:
        return ""
    nullable = t.startswith("?")
    t = t.lstrip("?")
    names = sorted(x.split("\\")[-1] for x in t.split("|"))
    if nullable and "null" not in names:
        names.append("null")
    return "|".join(sorted(names))


def _matching_paren(code: str, open_at: int) -> int:
    depth = 0
    for i in range(open_at, len(code)):
        if code[i] == "(":
            depth += 1
        elif code[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise AssertionError("unterminated parameter list")


def class_body(code: str, cls: str) -> str:
    m = re.search(r"\b(?:class|trait|interface)\s+" + re.escape(cls) + r"\b[^{]*\{", code)
    assert m, "class %s not found" % cls
    depth, i = 1, m.end()
    while depth and i < len(code):
        depth += code[i] == "{"
        depth -= code[i] == "}"
        i += 1
    return code[m.end():i - 1]


def class_signatures(src: str, cls: str) -> dict:
    """method -> {visibility, static, params: [{type, name, default, variadic, by_ref}], returns}"""
    body = class_body(strip_php(src), cls)
    sigs = {}
    depth_at, d = [], 0                  # brace depth before each character: methods of the class sit at depth 0
    for ch in body:
        depth_at.append(d)
        d += (ch == "{") - (ch == "}")
    for m in re.finditer(r"((?:\b(?:public|protected|private|static|final|abstract)\s+)*)function\s+&?\s*(\w+)\s*\(", body):
        if depth_at[m.start()] != 0:
            continue                     # a closure or a method of an anonymous class inside a method body
        mods = m.group(1).split()
        vis = next((v for v in mods if v in ("public", "protected", "private")), "public")
        close = _matching_paren(body, m.end() - 1)
        params = []
        for p in _split_top(body[m.end():close]):
            pm = re.match(r"^(?:(?:public|protected|private|readonly)\s+)*([\w\\|?\s]*?)\s*(&?)\s*(\.\.\.)?\s*\$(\w+)\s*(=.*)?$",
                          p, flags=re.S)
            assert pm, "cannot read parameter '%s' of %s::%s" % (p, cls, m.group(2))
            params.append({"type": _norm_type(pm.group(1)), "by_ref": bool(pm.group(2)), "variadic": bool(pm.group(3)),
                           "name": pm.group(4), "default": pm.group(5) is not None})
        rm = re.match(r"\s*:\s*([\w\\|?\s]+?)\s*[{;]", body[close + 1:close + 200])
        sigs[m.group(2)] = {"visibility": vis, "static": "static" in mods, "params": params,
                            "returns": _norm_type(rm.group(1)) if rm else ""}
    return sigs
