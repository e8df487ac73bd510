"""Host-side logic that needs no GPU: argument validation (exact reference strings), shape folding
in LinearAlgebra / MatrixOperator, NDArray views, service-level diagnosis.  CPU only."""
import numpy as np
import pytest

import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import LinearAlgebra, MatrixOperator, NDArray, dtypes
from rindow_math_matrix_b200.cuda_blas import (assert_matrix_buffer_spec, assert_shape_parameter,
                                               assert_vector_buffer_spec, code_to_trans)


def test_utils_messages():
    """Utils.php:30-65 -- the strings PhpBlasTest.php:1988-2155 asserts."""
    with pytest.raises(rb.InvalidArgumentException, match=r"Argument m must be greater than 0\."):
        assert_shape_parameter("m", 0)
    b = list(range(6))
    with pytest.raises(rb.InvalidArgumentException, match=r"Matrix specification too large for bufferA\."):
        assert_matrix_buffer_spec("A", b, 3, 3, 0, 3)
    with pytest.raises(rb.InvalidArgumentException, match=r"Argument offsetA must be greater than equals 0\."):
        assert_matrix_buffer_spec("A", b, 2, 3, -1, 3)
    with pytest.raises(rb.InvalidArgumentException, match=r"Argument ldA must be greater than 0\."):
        assert_matrix_buffer_spec("A", b, 2, 3, 0, 0)
    assert_matrix_buffer_spec("A", b, 2, 3, 0, 3)
    with pytest.raises(rb.InvalidArgumentException, match=r"Vector specification too large for bufferX\."):
        assert_vector_buffer_spec("X", b, 4, 0, 2)
    with pytest.raises(rb.InvalidArgumentException, match=r"Argument incX must be greater than 0\."):
        assert_vector_buffer_spec("X", b, 4, 0, 0)
    assert code_to_trans(111) == (False, False) and code_to_trans(113) == (True, True)
    with pytest.raises(rb.InvalidArgumentException, match="Unknown Tranpose Code"):
        code_to_trans(5)


def test_ndarray_views_and_reshape(oracle_service):
    a = NDArray(np.arange(24).reshape(2, 3, 4), dtype=dtypes.float32, service=oracle_service)
    assert a.shape() == [2, 3, 4] and a.ndim() == 3 and a.size() == 24 and a.offset() == 0
    v = a[1]
    assert v.offset() == 12 and v.shape() == [3, 4] and v.buffer() is a.buffer()
    assert v[2].offset() == 20 and v[2][3] == 23.0
    assert a.reshape([6, 4])[5].toArray() == [20, 21, 22, 23]
    with pytest.raises(rb.InvalidArgumentException, match="Unmatch size to reshape"):
        a.reshape([5, 5])
    with pytest.raises(rb.OutOfRangeException):
        a[2]
    v[0] = [9, 9, 9, 9]
    assert a.toArray()[1][0] == [9, 9, 9, 9]
    assert a[0:1].shape() == [1, 3, 4]


def test_reduce_axis_errors(oracle_service):
    la = LinearAlgebra(oracle_service)
    x = la.array([[1, 2, 3], [4, 5, 6]])
    with pytest.raises(rb.InvalidArgumentException, match="Invalid axis: 2"):
        la.reduceSum(x, axis=2)
    with pytest.raises(rb.InvalidArgumentException, match="Invalid axis: -3"):
        la.reduceMax(x, axis=-3)
    out = la.alloc([2])
    with pytest.raises(rb.InvalidArgumentException, match=r"Unmatch shape of dimension: \(2,3\),\(2\)"):
        la.reduceSum(x, axis=0, output=out)
    assert la.reduceArgMax(x, axis=1).dtype() == dtypes.int32          # LinearAlgebra.php:3311-3313
    assert la.reduceArgMax(x, axis=1, dtype=dtypes.int64).dtype() == dtypes.int64


def test_broadcast_errors(oracle_service):
    la = LinearAlgebra(oracle_service)
    with pytest.raises(rb.InvalidArgumentException, match=r"Unmatch dimension size for broadcast\.: \[2\] => \[2,3\]"):
        la.add(la.array([1, 2]), la.array([[1, 2, 3], [4, 5, 6]]))
    with pytest.raises(rb.InvalidArgumentException, match=r"Unmatch dimension size for broadcast\.: \[3\] => \[3,2\]"):
        la.multiply(la.array([1, 2, 3]), la.array([[1, 2, 3], [4, 5, 6]]), trans=True)
    with pytest.raises(rb.InvalidArgumentException, match='"X" must be 2-D dimension'):
        la.softmax(la.array([1, 2, 3]))
    with pytest.raises(rb.InvalidArgumentException, match="Dimensions must be 2D-NDArray"):
        la.gemm(la.array([1, 2, 3]), la.array([[1], [2], [3]]))


def test_cross_errors(oracle_service):
    mo = MatrixOperator(oracle_service)
    with pytest.raises(rb.InvalidArgumentException, match=r"Unmatch shape of dimension to cross multiple: \(2,3\),\(2,3\)"):
        mo.cross(mo.array([[1, 2, 3], [4, 5, 6]]), mo.array([[1, 2, 3], [4, 5, 6]]))
    with pytest.raises(rb.InvalidArgumentException, match="unmatch value type"):
        mo.cross(mo.array([[1.0]]), mo.array([[1.0]], dtype=dtypes.float64))


def test_im2col_shape_errors(oracle_service):
    la = LinearAlgebra(oracle_service)
    with pytest.raises(rb.InvalidArgumentException, match="Invalid shape or paramaters."):
        la.im2col(la.alloc([1, 2, 2, 1]), filterSize=[3, 3])
    with pytest.raises(rb.InvalidArgumentException, match="unsuppoted images shape"):
        la.im2col(la.alloc([2, 2]))


def test_service_without_gpu_is_basic_and_refuses_to_compute():
    from rindow_math_matrix_b200 import _capi
    if _capi.device_available():
        pytest.skip("a GPU is visible")
    svc = rb.MatlibCuda()
    assert svc.serviceLevel() == rb.Service.LV_BASIC
    assert "Basic" in svc.info()
    with pytest.raises(rb.LogicException, match="no CPU fallback"):
        svc.blas()
    with pytest.raises(rb.LogicException):
        svc.math(rb.Service.LV_BASIC)
    with pytest.raises(_capi.B200Error):
        rb.CudaBuffer(4, dtypes.float32)


def test_device_tier_without_gpu_fails_loudly():
    """The device-resident tier is importable everywhere and refuses to work without a B200: no queue, no device
    buffers, no accelerated LA (and the accelerator name the reference knows is not ours)."""
    from rindow_math_matrix_b200 import _capi
    assert rb.LinearAlgebraCuda.__mro__[1] is LinearAlgebra and rb.NDArrayCuda.__name__ == "NDArrayCuda"
    if _capi.device_available():
        pytest.skip("a GPU is visible")
    svc = rb.MatlibCuda()
    assert svc.serviceLevel() < rb.Service.LV_ACCELERATED
    with pytest.raises(rb.LogicException):
        svc.createQueue()
    with pytest.raises(rb.LogicException):
        svc.buffer(rb.Service.LV_ACCELERATED)
    mo = MatrixOperator(svc)
    assert not mo.isAccelerated()
    with pytest.raises(rb.LogicException):
        mo.laAccelerated("cuda")
    with pytest.raises(rb.LogicException, match="Unknown accelerator"):
        mo.laAccelerated("clblast")


def test_sharding_geometry_and_merges_single_process():
    """row stripes of the sharded gemm and the cross-rank merges on one process (world 1): identities."""
    import torch
    from rindow_math_matrix_b200 import sharding
    assert sharding.row_range(16384, 8, 3) == (6144, 2048)
    v = torch.tensor([1.0, float("inf"), float("nan"), -float("inf")], dtype=torch.float64)
    m = sharding.merge_max(v.clone())
    assert m[0] == 1.0 and m[1] == float("inf") and torch.isnan(m[2]) and m[3] == -float("inf")
    idx = sharding.merge_argmax(v.clone(), torch.tensor([3, 0, 2, 7]), 10)
    assert idx.tolist() == [13, 10, 12, 17]
    rec = torch.tensor([0, 5, 0, 9], dtype=torch.int64)
    rec[0] = torch.tensor([2.5], dtype=torch.float64).view(torch.int64)[0]
    rec[2] = torch.tensor([-1.0], dtype=torch.float64).view(torch.int64)[0]
    vals, ids = sharding.split_records(rec, 2)
    assert vals.tolist() == [2.5, -1.0] and ids.tolist() == [5, 9]


def test_conv2d_composition_over_any_driver(oracle_service):
    """LinearAlgebra.conv2d without a fused driver entry = im2col -> gemm -> add (the reference calls)."""
    la = LinearAlgebra(oracle_service)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 5, (2, 6, 7, 3)).astype(np.float32)
    W = rng.integers(-2, 3, (3, 2, 3, 4)).astype(np.float32)
    bias = np.arange(4, dtype=np.float32)
    out = la.conv2d(la.array(img), la.array(W), la.array(bias), strides=(1, 2), padding=False)
    ref = np.zeros((2, 4, 3, 4))
    for b in range(2):
        for oy in range(4):
            for ox in range(3):
                patch = img[b, oy:oy + 3, ox * 2:ox * 2 + 2, :]
                ref[b, oy, ox] = np.tensordot(patch, W, axes=([0, 1, 2], [0, 1, 2])) + bias
    assert out.shape() == [2, 4, 3, 4]
    assert np.array_equal(out.to_numpy(), ref)
    with pytest.raises(rb.InvalidArgumentException, match="Unmatch channels"):
        la.conv2d(la.array(img), la.array(W[:, :, :2, :]))


def test_driver_mirrors_keep_the_reference_parameter_lists():
    """CudaMath / CudaBlas take the reference drivers' arguments by position (src/LinearAlgebra.php passes them that
    way): every public method that exists in both has the same parameter names in the same order.  The reference
    side is the fixture tests/golden/reference_signatures.json (tools/extract_reference_signatures.py)."""
    import inspect
    import json
    import os

    from rindow_math_matrix_b200.cuda_blas import CudaBlas
    from rindow_math_matrix_b200.cuda_math import CudaMath

    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_signatures.json")) as f:
        ref = json.load(f)
    compared = 0
    for cls, parent in ((CudaMath, "PhpMath"), (CudaBlas, "PhpBlas")):
        for name, sig in ref[parent]["methods"].items():
            fn = getattr(cls, name, None)
            if sig["visibility"] != "public" or name.startswith("__") or fn is None or not callable(fn):
                continue
            want = [p["name"].lower() for p in sig["params"]]
            got = [p.name.rstrip("_").lower() for p in inspect.signature(fn).parameters.values() if p.name != "self"]
            assert got[:len(want)] == want, "%s.%s: %s vs reference %s" % (cls.__name__, name, got, want)
            for extra in list(inspect.signature(fn).parameters.values())[1 + len(want):]:
                assert extra.default is not inspect.Parameter.empty or extra.kind in (
                    inspect.Parameter.VAR_POSITIONAL, inspect.Parameter.VAR_KEYWORD), (name, extra.name)
            compared += 1
    assert compared >= 70, compared


def test_linear_algebra_shape_helpers_and_isclose(oracle_service):
    """expandDims / squeeze (LinearAlgebraTest.php:195-259), the accessors (:61-93, 153-189) and the reference's own
    comparator isclose (src/LinearAlgebra.php:5439-5464), over the oracle service."""
    import rindow_math_matrix_b200 as rb
    from rindow_math_matrix_b200 import LinearAlgebra, NDArray, dtypes

    la = LinearAlgebra(oracle_service)
    x = la.alloc([2, 3])
    assert [la.expandDims(x, axis=a).shape() for a in (0, 1, 2, -1, -2, -3)] == \
        [[1, 2, 3], [2, 1, 3], [2, 3, 1], [2, 3, 1], [2, 1, 3], [1, 2, 3]]
    s = la.alloc([])
    assert s.shape() == [] and s.size() == 1 and len(s.buffer()) == 1
    assert la.expandDims(s, axis=0).shape() == [1] and la.expandDims(s, axis=-1).shape() == [1]
    with pytest.raises(rb.InvalidArgumentException, match="axis is out of range: -4"):
        la.expandDims(x, axis=-4)
    y = la.alloc([1, 2, 1, 3, 1])
    assert la.squeeze(y).shape() == [2, 3]
    assert [la.squeeze(y, axis=a).shape() for a in (0, 2, 4, -1, -3, -5)] == \
        [[2, 1, 3, 1], [1, 2, 3, 1], [1, 2, 1, 3], [1, 2, 1, 3], [1, 2, 3, 1], [2, 1, 3, 1]]
    with pytest.raises(rb.InvalidArgumentException, match="axis is out of range: -6"):
        la.squeeze(y, axis=-6)
    with pytest.raises(rb.InvalidArgumentException, match=r"Can not squeeze dim\[3\]"):
        la.squeeze(y, axis=3)

    assert la.fp64() and la.getBlas() is la.blas and la.getMath() is la.math
    i = NDArray([1, 2], dtype=dtypes.int32, service=oracle_service)
    assert la.isInt(i) and not la.isFloat(i) and la.isFloat(x) and not la.isInt(x)
    assert la.scalar(3.5) == 3.5 and la.scalar(NDArray([1, 2], service=oracle_service)) == [1, 2]
    assert la.abs(-2.5) == 2.5 and not la.isComplexDtype(dtypes.float32) and la.isComplexDtype(dtypes.complex64)

    a = NDArray([[1.0, 2.0], [3.0, 1000.0]], service=oracle_service)
    assert la.isclose(a, la.copy(a))
    b = la.copy(a)
    b.buffer()[3] = 1000.05                               # 0.05 < 1e-7 + 1e-4 * 1000.05
    assert la.isclose(a, b)
    b.buffer()[0] = 1.2                                   # 0.2 > 0.1
    assert not la.isclose(a, b)
    assert la.isclose(a, b, rtol=1e-3)
    assert not la.isclose(a, NDArray([1.0, 2.0, 3.0, 1000.0], service=oracle_service))       # shapes differ
    assert a.toArray() == [[1.0, 2.0], [3.0, 1000.0]]     # the operands are left alone

    # conjA / conjB on real dtypes select the Conj* codes, which mean the same as the plain ones (PhpBlas.php:883-886)
    A = NDArray([[1, 2, 3], [4, 5, 6]], service=oracle_service)
    B = NDArray([[1, 0], [0, 1], [1, 1]], service=oracle_service)
    want = la.gemm(A, B).toArray()
    assert la.gemm(A, B, conjA=True, conjB=True).toArray() == want
    assert la.gemm(B, A, transA=True, transB=True, conjA=True).toArray() == la.gemm(B, A, transA=True, transB=True).toArray()


def test_linear_algebra_mirror_keeps_the_reference_parameter_lists():
    """Every public LinearAlgebra / MatrixOperator / NDArray method the mirror has takes the reference's parameters,
    by name and order (named arguments are how the reference's tests call them: `axis:`, `dtype:`, `keepdims:` ...)."""
    import inspect
    import json
    import os

    from rindow_math_matrix_b200 import LinearAlgebra, MatrixOperator, NDArray

    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_signatures.json")) as f:
        ref = json.load(f)
    compared = 0
    for cls, parent in ((LinearAlgebra, "LinearAlgebra"), (MatrixOperator, "MatrixOperator"), (NDArray, "NDArrayPhp")):
        for name, sig in ref[parent]["methods"].items():
            fn = getattr(cls, name, None)
            if sig["visibility"] != "public" or name.startswith("__") or fn is None or not callable(fn):
                continue
            want = [p["name"].lower() for p in sig["params"]]
            got = [p.name.rstrip("_").lower() for p in inspect.signature(fn).parameters.values() if p.name != "self"]
            assert got[:len(want)] == want, "%s.%s: %s vs reference %s" % (cls.__name__, name, got, want)
            compared += 1
    assert compared >= 130, compared
