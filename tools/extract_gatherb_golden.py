#!/usr/bin/env python3
"""Extract the gatherb / scatterb / scatterbAdd cases of the reference's own tests
(tests/RindowTest/Math/Matrix/LinearAlgebraTest.php: testGatherbNormal, testScatterbNormal, testScatterbAddNormal)
into tests/golden/gatherb_cases.json, and the gatherND / scatterND / scatterNDAdd cases (testGatherNDNormal,
testScatterNDNormal, testScatterNDAddNormal) into tests/golden/gathernd_cases.json.  Run in the build container (reads /root/reference); the JSON is committed.
Each case: op, params ("arange" over shapeA or a literal), indices, kwargs, shape (scatter), expect, ref line."""
import json
import re
import sys

SRC = "/root/reference/tests/RindowTest/Math/Matrix/LinearAlgebraTest.php"


def php_literal(text):
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r",\s*([\]\)])", r"\1", text)
    return json.loads(text)


def balanced(text, start):
    """the bracketed literal starting at text[start] == '['"""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == "[":
            depth += 1
        elif text[i] == "]":
            depth -= 1
            if depth == 0:
                return text[start:i + 1]
    raise ValueError("unbalanced")


def main():
    lines = open(SRC).read().split("\n")
    src = "\n".join(lines)
    cases = []
    for fn in ("testGatherbNormal", "testScatterbNormal", "testScatterbAddNormal"):
        starts = [m.start() for m in re.finditer(r"public function %s\(\)" % fn, src)]
        begin = starts[0]                                     # the second testScatterbNormal belongs to another suite section
        end = src.index("    public function ", begin + 10)
        body = src[begin:end]
        base_line = src[:begin].count("\n") + 1
        calls = list(re.finditer(r"\$b = \$la->(gatherb|scatterb|scatterbAdd)\((.*?)\);", body, re.S))
        prev_end = 0
        for call in calls:
            block = body[prev_end:call.end()]
            after = body[call.end():]
            prev_end = call.end()
            op, argtext = call.group(1), call.group(2)
            shapeA = None
            if "$shapeA = " in block:
                shapeA = php_literal(balanced(block, block.rindex("$shapeA = ") + len("$shapeA = ")))
            shapeB = None
            if "$shapeB = " in block:
                shapeB = php_literal(balanced(block, block.rindex("$shapeB = ") + len("$shapeB = ")))
            a_pos = block.rindex("$a = $la->array(")
            a_arg = block[a_pos + len("$a = $la->array("):]
            if a_arg.startswith("range("):
                params = {"arange": shapeA}
            else:
                params = {"a": php_literal(balanced(a_arg, 0))}
            x_pos = block.rindex("$x = $la->array(")
            indices = php_literal(balanced(block, x_pos + len("$x = $la->array(")))
            if "$x = $la->reduceArgMax($x,axis:-1" in block[x_pos:]:       # the argmax case feeds its own indices
                indices = [max(range(len(row)), key=lambda q: (row[q], -q)) for row in indices]
            kwargs = {}
            for k, v in re.findall(r"(\w+):\s*([^,]+)", argtext):
                v = v.strip()
                if v == "$a->ndim()":
                    v = len(shapeA) if op == "gatherb" else None
                elif v == "count($shapeB)":
                    v = len(shapeB)
                kwargs[k] = int(v)
            case = {"op": op, "params": params, "indices": indices, "kwargs": kwargs,
                    "ref": "LinearAlgebraTest.php:%d" % (base_line + body[:call.start()].count("\n"))}
            if op != "gatherb":
                case["shape"] = shapeB
            t_pos = after.index("$trues = $la->array(")
            case["expect"] = php_literal(balanced(after, t_pos + len("$trues = $la->array(")))
            cases.append(case)
    nd = []
    for fn in ("testGatherNDNormal", "testScatterNDNormal", "testScatterNDAddNormal"):
        begin = src.index("public function %s()" % fn)
        end = src.index("    public function ", begin + 10)
        body = src[begin:end]
        base_line = src[:begin].count("\n") + 1
        prev_end = 0
        for call in re.finditer(r"\$b = \$la->(gatherND|scatterND|scatterNDAdd)\((.*?)\);", body, re.S):
            block, after = body[prev_end:call.end()], body[call.end():]
            prev_end = call.end()
            op = call.group(1)
            a_arg = block[block.rindex("$a = $la->array(") + len("$a = $la->array("):].lstrip()
            if a_arg.startswith("range("):
                shapeA = php_literal(balanced(block, block.rindex("$shapeA = ") + len("$shapeA = ")))
                params = {"arange": shapeA}
            elif a_arg.startswith("["):
                params = {"a": php_literal(balanced(a_arg, 0))}
            else:
                params = {"a": json.loads(a_arg[:a_arg.index(")")])}            # a scalar: $la->array(1)
            x_arg = block[block.rindex("$x = $la->array(") + len("$x = $la->array("):].lstrip()
            case = {"op": op, "params": params, "indices": php_literal(balanced(x_arg, 0)),
                    "kwargs": {"batchDims": int(re.search(r"batchDims:(\d+)", call.group(2)).group(1))},
                    "ref": "LinearAlgebraTest.php:%d" % (base_line + body[:call.start()].count("\n"))}
            if op != "gatherND":
                case["shape"] = php_literal(balanced(block, block.rindex("$shape = ") + len("$shape = ")))
            # the expectation is the assertEquals(<literal>,$b->toArray()) that follows the shape assertions
            m = re.search(r"\$this->assertEquals\(\s*((?:(?!\$this->assertEquals).)*?)\s*,\s*\$b->toArray\(\)\);", after, re.S)
            case["expect"] = php_literal(m.group(1))
            nd.append(case)
    json.dump({"source": "LinearAlgebraTest.php testGatherNDNormal / testScatterNDNormal / testScatterNDAddNormal",
               "cases": nd}, open("tests/golden/gathernd_cases.json", "w"), indent=1)
    print(len(nd), "cases -> tests/golden/gathernd_cases.json")
    img = []
    for fn in ("testImagecopy", "testImagecopychannelsfirst"):
        begin = src.index("public function %s()" % fn)
        end = src.index("    public function ", begin + 10)
        body = src[begin:end]
        base_line = src[:begin].count("\n") + 1
        events = [(m.start(), "a", m) for m in re.finditer(r"\$a = \$la->array\(", body)]
        events += [(m.start(), "call", m) for m in re.finditer(r"\$b = \$la->imagecopy\(\$a,null,(null|true),(.*?)\);", body, re.S)]
        a = None
        names = ["heightShift", "widthShift", "verticalFlip", "horizontalFlip", "rgbFlip"]
        for pos, kind, m in sorted(events, key=lambda e: e[0]):
            if kind == "a":
                a = php_literal(balanced(body, m.end()))
                continue
            kwargs = {}
            for slot, arg in enumerate(x.strip() for x in m.group(2).split(",")):
                value = arg.split("=")[-1].strip()                       # "$heightShift=1" or a bare value
                if value in ("true", "false"):
                    kwargs[names[slot]] = value == "true"
                elif value != "null":
                    kwargs[names[slot]] = int(value)
            if m.group(1) == "true":
                kwargs["channels_first"] = True
            exp = re.search(r"\$this->assertEquals\(\s*(\[.*?\])\s*,\s*\$b->toArray\(\)\);", body[m.end():], re.S)
            img.append({"a": a, "kwargs": kwargs, "expect": php_literal(exp.group(1)),
                        "ref": "LinearAlgebraTest.php:%d" % (base_line + body[:pos].count("\n"))})
    json.dump({"source": "LinearAlgebraTest.php testImagecopy / testImagecopychannelsfirst", "cases": img},
              open("tests/golden/imagecopy_cases.json", "w"), indent=1)
    print(len(img), "cases -> tests/golden/imagecopy_cases.json")
    out = sys.argv[1] if len(sys.argv) > 1 else "tests/golden/gatherb_cases.json"
    json.dump({"source": "LinearAlgebraTest.php testGatherbNormal / testScatterbNormal / testScatterbAddNormal",
               "cases": cases}, open(out, "w"), indent=1)
    print(len(cases), "cases ->", out)


if __name__ == "__main__":
    main()
