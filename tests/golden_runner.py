"""Runs the golden known-answer cases (tests/golden/known_answers.json) through the host classes
(LinearAlgebra / MatrixOperator / NDArray) over ANY driver service -- the oracle service on CPU,
the CUDA service on the GPU.  The same thing the reference does when LinearAlgebraCLTest re-runs
LinearAlgebraTest over another driver (tests/.../LinearAlgebraCLTest.php:722-758)."""
import json
import math
import os

import numpy as np

from rindow_math_matrix_b200 import InvalidArgumentException, LinearAlgebra, MatrixOperator

HERE = os.path.dirname(os.path.abspath(__file__))


def load_known_answers():
    with open(os.path.join(HERE, "golden", "known_answers.json")) as f:
        return json.load(f)


def _num(v):
    if isinstance(v, str):
        return {"inf": math.inf, "-inf": -math.inf, "nan": math.nan}[v]
    return v


def _arr(x):
    if isinstance(x, list):
        return [_arr(v) for v in x]
    return _num(x)


def _assert_equal(expect, got, tol=0.0):
    e = np.array(_arr(expect), dtype=np.float64)
    g = np.array(got, dtype=np.float64)
    assert e.shape == g.shape, (e.shape, g.shape)
    if tol:
        assert np.allclose(e, g, rtol=0, atol=tol, equal_nan=True), (e, g)
    else:
        assert np.array_equal(e, g, equal_nan=True), (e, g)


_DTYPES = {"bool": 0, "int8": 1, "int16": 2, "int32": 3, "int64": 4, "uint8": 5, "uint16": 6, "uint32": 7,
           "uint64": 8, "float32": 12, "float64": 13}


def _la_arg(la, a):
    if isinstance(a, str) and a.startswith("@dtype:"):
        return _DTYPES[a[7:]]
    if isinstance(a, str) and a.startswith("@ones"):
        return la.ones(la.alloc(json.loads(a[5:])))
    if not isinstance(a, dict):
        return _num(a)
    if "list" in a:                                            # a PHP array of NDArrays (stack / concat values)
        return [_la_arg(la, x) for x in a["list"]]
    dtype = _DTYPES[a["dtype"]] if "dtype" in a else None
    if "ones" in a:
        x = la.ones(la.alloc(a["ones"], dtype=dtype))
    elif "fill" in a:
        x = la.fill(a["fill"], la.alloc(a["shape"], dtype=dtype))
    elif "arange" in a:
        v = np.arange(a["arange"]) + a.get("start", 0)
        for i, val in a.get("set", {}).items():
            v[int(i)] = val
        x = la.array(v, dtype=dtype)
    else:
        v = np.array(_arr(a["a"]))
        if dtype is not None and a["dtype"].startswith("uint"):
            v = v.astype(np.int64).astype({"uint8": np.uint8, "uint16": np.uint16, "uint32": np.uint32,
                                           "uint64": np.uint64}[a["dtype"]])      # PHP ints wrap into the C buffer
        x = la.array(v, dtype=dtype)
    if "view" in a:
        v = a["view"]
        x = x[slice(v[0], v[1])] if isinstance(v, list) else x[v]          # [start, stop): PHP's R(start, stop)
    return x


def run_case(case, service):
    mo = MatrixOperator(service)
    la = LinearAlgebra(service)
    op = case["op"]
    kw = dict(case.get("kwargs", {}))

    def call():
        if op == "gemm":
            A, B = la.array(_arr(case["A"])), la.array(_arr(case["B"]))
            C = la.array(_arr(case["C"])) if "C" in case else None
            out = la.gemm(A, B, C=C, **kw)
            if C is not None:
                assert out is C
            return out
        if op == "cross":
            A, B = mo.array(_arr(case["A"])), mo.array(_arr(case["B"]))
            if "viewA" in case:
                A = A[case["viewA"]]
                assert A.offset() == case["offsetA"]
            if "viewB" in case:
                B = B[case["viewB"]]
                assert B.offset() == case["offsetB"]
            a0, b0 = A.toArray(), B.toArray()
            out = mo.cross(A, B)
            assert A.toArray() == a0 and B.toArray() == b0       # operands untouched
            return out
        if op in ("add", "multiply"):
            X, A = la.array(_arr(case["X"])), la.array(_arr(case["A"]))
            x0 = X.toArray()
            out = getattr(la, op)(X, A, **kw)
            assert out is A and X.toArray() == x0
            return out
        if op in ("reduceSum", "reduceMax", "reduceArgMax"):
            if "x_ones" in case:
                x = la.ones(la.alloc(case["x_ones"]))
            else:
                x = la.array(_arr(case["x"]))
            if "view" in case:
                x = x[case["view"]]
            return getattr(la, op)(x, **kw)
        if op == "softmax":
            return la.softmax(la.array(_arr(case["x"])))
        if op == "scal":
            x = la.array(_arr(case["x"]))
            out = la.scal(case["alpha"], x)
            assert out is x
            return out
        if op == "axpy":
            x, y = la.array(_arr(case["x"])), la.array(_arr(case["y"]))
            x0 = x.toArray()
            out = la.axpy(x, y, case.get("alpha"))
            assert out is y and x.toArray() == x0
            return out
        if op == "copy":
            x = la.array(_arr(case["x"]))
            if case.get("into"):
                y = la.zerosLike(x)
                out = la.copy(x, y)
                assert out is y
            else:
                out = la.copy(x)
            assert out.shape() == x.shape() and out.dtype() == x.dtype()
            return out
        if op == "gemv":
            A, X = la.array(_arr(case["A"])), la.array(_arr(case["X"]))
            return la.gemv(A, X, **kw)
        if op == "matmul":
            A, B = la.array(_arr(case["A"])), la.array(_arr(case["B"]))
            C = la.zeros(la.alloc(case["C_zeros"])) if "C_zeros" in case else None
            return la.matmul(A, B, C=C, **kw)
        if op == "transpose":
            return la.transpose(la.array(_arr(case["x"])))
        if op == "la":
            # generic: method of LinearAlgebra called on NDArray / scalar arguments; "steps" = further in-place calls
            # on the first argument (the reference tests chain them); the result is the last call's return value
            args = [_la_arg(la, a) for a in case["args"]]
            kwa = {k: _la_arg(la, v) for k, v in kw.items()}
            out = getattr(la, case["method"])(*args, **kwa)
            if case.get("returns_arg") is not None:
                assert out is args[case["returns_arg"]]
            for step in case.get("steps", []):
                out = getattr(la, step["method"])(args[0], *[_la_arg(la, a) for a in step.get("args", [])],
                                                  **step.get("kwargs", {}))
            for i, want in case.get("untouched", {}).items():
                _assert_equal(want, args[int(i)].toArray())
            for i, want in case.get("args_after", {}).items():       # in-place results (swap)
                _assert_equal(want, args[int(i)].toArray())
            return out
        raise AssertionError("unknown op " + op)

    if "error" in case:
        try:
            call()
        except InvalidArgumentException as e:
            assert case["error"] in str(e), str(e)
        else:
            raise AssertionError("expected InvalidArgumentException: " + case["error"])
        return
    out = call()
    if "expect_shape" in case:
        assert out.shape() == case["expect_shape"]
    if "expect_dtype" in case:
        assert out.dtype() == _DTYPES[case["expect_dtype"]]
    if "expect_list" in case:                                  # split: an array of NDArrays
        assert len(out) == len(case["expect_list"])
        for want, got in zip(case["expect_list"], out):
            _assert_equal(want, got.toArray())
    if "expect" in case:
        _assert_equal(case["expect"], out.toArray() if hasattr(out, "toArray") else out, case.get("tol", 0.0))
    if "expect_fill" in case:
        got = np.asarray(out.toArray(), dtype=np.float64)
        assert np.all(got == case["expect_fill"])
    if "expect_row0" in case:
        _assert_equal(case["expect_row0"], out.toArray()[0], case["tol"])
        _assert_equal(case["expect_row0"], out.toArray()[-1], case["tol"])


def im2col_reference_loops(params, images_np):
    """The index loops the reference TEST uses to build its expectation
    (LinearAlgebraTest.php:8962-9005 forward, :9190-9221 reverse) -- independent of PhpMath, so
    it pins both the oracle and the CUDA path.  Returns (cols_expect, images_back_expect)."""
    p = {k: (v if v is not None else 0) for k, v in params.items()}
    b_, im_h, im_w, ch = p["batches"], p["im_h"], p["im_w"], p["channels"]
    kh, kw, sh, sw = p["kernel_h"], p["kernel_w"], p["stride_h"], p["stride_w"]
    dh, dw = p["dilation_h"], p["dilation_w"]
    padding, cf, ccf = bool(params["padding"]), bool(params["channels_first"]), bool(params["cols_channels_first"])
    out_h = int((im_h - (kh - 1) * dh - 1) / sh) + 1
    out_w = int((im_w - (kw - 1) * dw - 1) / sw) + 1
    if padding:
        ph = int(((im_h - 1) * sh - im_h + (kh - 1) * dh + 1) / 2)
        pw = int(((im_w - 1) * sw - im_w + (kw - 1) * dw + 1) / 2)
        out_h, out_w = im_h, im_w
    else:
        ph = pw = 0
    shape = (b_, out_h, out_w, ch, kh, kw) if ccf else (b_, out_h, out_w, kh, kw, ch)
    flat_im = images_np.reshape(-1)
    trues = np.zeros(int(np.prod(shape)))
    back = np.zeros(flat_im.size)
    for b in range(b_):
        for c in range(ch):
            for y in range(out_h):
                for x in range(out_w):
                    for ky in range(kh):
                        for kx in range(kw):
                            iy = y * sh + ky * dh - ph
                            ix = x * sw + kx * dw - pw
                            if cf:
                                iid = ((b * ch + c) * im_h + iy) * im_w + ix
                            else:
                                iid = ((b * im_h + iy) * im_w + ix) * ch + c
                            if ccf:
                                cid = ((((b * out_h + y) * out_w + x) * ch + c) * kh + ky) * kw + kx
                            else:
                                cid = ((((b * out_h + y) * out_w + x) * kh + ky) * kw + kx) * ch + c
                            if 0 <= iy < im_h and 0 <= ix < im_w:
                                trues[cid] = flat_im[iid]
                                back[iid] += flat_im[iid]
    return trues.reshape(shape), back.reshape(images_np.shape)


def run_im2col_case(params, service):
    """testIm2col2dNormal (LinearAlgebraTest.php:8903-9221) for one provider row."""
    la = LinearAlgebra(service)
    p = params
    n = p["batches"] * p["im_h"] * p["im_w"] * p["channels"]
    arange = np.arange(n, dtype=np.float32)
    if p["channels_first"]:
        shape = [p["batches"], p["channels"], p["im_h"], p["im_w"]]
    else:
        shape = [p["batches"], p["im_h"], p["im_w"], p["channels"]]
    images = la.array(arange.reshape(shape))
    cols = la.im2col(images, filterSize=[p["kernel_h"], p["kernel_w"]], strides=[p["stride_h"], p["stride_w"]],
                     padding=p["padding"], channels_first=p["channels_first"],
                     dilation_rate=[p["dilation_h"], p["dilation_w"]],
                     cols_channels_first=p["cols_channels_first"])
    trues, back = im2col_reference_loops(p, arange.reshape(shape).astype(np.float64))
    assert cols.shape() == list(trues.shape)
    assert np.array_equal(np.asarray(cols.toArray(), dtype=np.float64), trues)
    newImages = la.zerosLike(images)
    la.col2im(cols, newImages, filterSize=[p["kernel_h"], p["kernel_w"]], strides=[p["stride_h"], p["stride_w"]],
              padding=p["padding"], channels_first=p["channels_first"],
              dilation_rate=[p["dilation_h"], p["dilation_w"]], cols_channels_first=p["cols_channels_first"])
    assert np.array_equal(np.asarray(newImages.toArray(), dtype=np.float64), back)
    # im2col left the images untouched
    assert np.array_equal(np.asarray(images.toArray(), dtype=np.float64), arange.reshape(shape))


def run_gatherb_case(case, service):
    """tests/golden/gatherb_cases.json, gathernd_cases.json (LinearAlgebraTest.php testGatherbNormal / testScatterbNormal /
    testScatterbAddNormal / testGatherNDNormal / testScatterNDNormal / testScatterNDAddNormal): params float32 (range over shapeA or a literal), indices int32; operands untouched"""
    la = LinearAlgebra(service)
    p = case["params"]
    if "arange" in p:
        a = np.arange(int(np.prod(p["arange"])), dtype=np.float32).reshape(p["arange"])
    else:
        a = np.array(p["a"], dtype=np.float32)
    x = np.array(case["indices"], dtype=np.int32)
    A, X = la.array(a), la.array(x, dtype=_DTYPES["int32"])
    kw = case["kwargs"]
    if case["op"] in ("gatherb", "gatherND"):
        out = getattr(la, case["op"])(A, X, **kw)
    else:
        out = getattr(la, case["op"])(X, A, case["shape"], **kw)
    want = np.array(case["expect"], dtype=np.float32)
    assert out.shape() == list(want.shape), (out.shape(), want.shape)
    assert np.array_equal(out.to_numpy().reshape(want.shape), want)
    assert np.array_equal(A.to_numpy().reshape(a.shape), a) and np.array_equal(X.to_numpy().reshape(x.shape), x)


def run_imagecopy_case(case, service):
    """tests/golden/imagecopy_cases.json (LinearAlgebraTest.php testImagecopy / testImagecopychannelsfirst)"""
    la = LinearAlgebra(service)
    a = np.array(case["a"], dtype=np.float32)
    A = la.array(a)
    out = la.imagecopy(A, **case["kwargs"])
    want = np.array(case["expect"], dtype=np.float32)
    assert out.shape() == list(a.shape)
    assert np.array_equal(out.to_numpy().reshape(want.shape), want)
    assert np.array_equal(A.to_numpy().reshape(a.shape), a)


def run_im2col_nd_case(params, service):
    """testIm2col1dNormal / testIm2col3dNormal (LinearAlgebraTest.php:9613-9778, :10653-10882) for one provider row:
    arange-valued images, cols must hold the source index of every in-bounds tap (0 elsewhere), col2im must add every
    tap back.  The expectation loops are the reference TEST's (with its own padding formulas), independent of PhpMath."""
    la = LinearAlgebra(service)
    p = {k: (v if v is not None else 0) for k, v in params.items()}
    three = "im_d" in params
    b_, ch = p["batches"], p["channels"]
    dims = [p["im_d"], p["im_h"], p["im_w"]] if three else [p["im_w"]]
    ks = [p["kernel_d"], p["kernel_h"], p["kernel_w"]] if three else [p["kernel_w"]]
    ss = [p["stride_d"], p["stride_h"], p["stride_w"]] if three else [p["stride_w"]]
    ds = [p["dilation_d"], p["dilation_h"], p["dilation_w"]] if three else [p["dilation_w"]]
    padding, cf, ccf = bool(params["padding"]), bool(params["channels_first"]), bool(params["cols_channels_first"])
    outs = [int((i - (k - 1) * d - 1) / s) + 1 for i, k, s, d in zip(dims, ks, ss, ds)]
    pads = [0] * len(dims)
    if padding:
        pads = [int(((i - 1) * s - i + (k - 1) * d + 1) / 2) for i, k, s, d in zip(dims, ks, ss, ds)]
        outs = list(dims)
    n = b_ * ch * int(np.prod(dims))
    arange = np.arange(n, dtype=np.float32)
    shape = [b_, ch] + dims if cf else [b_] + dims + [ch]
    images = la.array(arange.reshape(shape))
    cols = la.im2col(images, filterSize=ks, strides=ss, padding=params["padding"], channels_first=params["channels_first"],
                     dilation_rate=ds, cols_channels_first=params["cols_channels_first"])
    cshape = [b_] + outs + ([ch] + ks if ccf else ks + [ch])
    assert cols.shape() == cshape
    # vectorised form of the test's nested loops: source coordinate of every (out cell, tap) pair per axis
    img = arange.reshape(shape).astype(np.float64)
    if cf:
        img = np.moveaxis(img, 1, -1)                          # -> [b, *dims, c]
    trues = np.zeros([b_] + outs + ks + [ch])
    back = np.zeros_like(img)
    nd = len(dims)
    grids = np.meshgrid(*[np.arange(o) for o in outs], *[np.arange(k) for k in ks], indexing="ij")
    src = [grids[a] * ss[a] + grids[nd + a] * ds[a] - pads[a] for a in range(nd)]
    ok = np.ones(src[0].shape, bool)
    for a in range(nd):
        ok &= (src[a] >= 0) & (src[a] < dims[a])
    idx = tuple(np.where(ok, src[a], 0) for a in range(nd))
    for b in range(b_):
        vals = img[b][idx]                                     # [*outs, *ks, c]
        trues[b] = np.where(ok[..., None], vals, 0.0)
        np.add.at(back[b], tuple(i[ok] for i in idx), vals[ok])
    if ccf:
        trues = np.moveaxis(trues, -1, 1 + nd)                 # -> [b, *outs, c, *ks]
    if cf:
        back = np.moveaxis(back, -1, 1)
    assert np.array_equal(np.asarray(cols.toArray(), dtype=np.float64), trues)
    newImages = la.zerosLike(images)
    la.col2im(cols, newImages, filterSize=ks, strides=ss, padding=params["padding"], channels_first=params["channels_first"],
              dilation_rate=ds, cols_channels_first=params["cols_channels_first"])
    assert np.array_equal(np.asarray(newImages.toArray(), dtype=np.float64), back)
    assert np.array_equal(np.asarray(images.toArray(), dtype=np.float64), arange.reshape(shape))
