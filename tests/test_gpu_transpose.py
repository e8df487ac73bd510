"""N-dimensional transpose (rb200_transpose) against the CPU oracle: every permutation of small ranks, all element
widths, views with offsets, large shapes through numpy (a pure copy: bit-exact).  pytest -m gpu"""
import itertools

import numpy as np
import pytest

import oracle
from rindow_math_matrix_b200 import CudaBuffer, InvalidArgumentException, LinearAlgebra, dtypes

pytestmark = pytest.mark.gpu

NP = {k: v for k, v in dtypes.NUMPY.items() if k not in dtypes.COMPLEX_TYPES}     # real dtypes: the arithmetic entry points


def _run(cuda_service, a, perm, dt, offA=0, offB=0):
    npdt = NP[dt]
    src = np.concatenate([np.zeros(offA, npdt), a.reshape(-1)])
    A = CudaBuffer(src.size, dt).from_numpy(src)
    B = CudaBuffer(a.size + offB + 1, dt)
    shape = CudaBuffer(a.ndim, dtypes.int32).from_numpy(np.asarray(a.shape, np.int32))
    pm = CudaBuffer(a.ndim, dtypes.int32).from_numpy(np.asarray(perm, np.int32))
    cuda_service.math().transpose(shape, pm, A, offA, B, offB)
    out = B.to_numpy()
    assert out[-1] == 0 and np.all(out[:offB] == 0)
    return out[offB:offB + a.size]


@pytest.mark.parametrize("shape", [(6,), (2, 3), (4, 2, 3), (2, 4, 3, 2), (3, 1, 5, 2), (2, 3, 4, 5, 3)],
                         ids=lambda s: "x".join(map(str, s)))
def test_all_permutations_small(shape, cuda_service):
    a = np.arange(int(np.prod(shape)), dtype=np.float32).reshape(shape)
    for perm in itertools.permutations(range(len(shape))):
        got = _run(cuda_service, a, perm, dtypes.float32, offA=3, offB=5)
        ref = np.zeros(a.size)
        oracle.transpose(shape, perm, oracle.as_buffer(np.concatenate([np.zeros(3), a.reshape(-1)])), 3, ref, 0)
        assert np.array_equal(got.astype(np.float64), ref), perm
        assert np.array_equal(got, a.transpose(perm).reshape(-1)), perm


LARGE = [((64, 12, 128, 64), (0, 2, 1, 3)),      # attention head split: [b, h, s, d] <-> [b, s, h, d] (rows kernel)
         ((64, 128, 768), (0, 2, 1)),            # last axis changes (tile kernel), batch axis
         ((300, 517), (1, 0)),                   # 2-D, odd sizes
         ((64, 224, 224, 3), (0, 3, 1, 2)),      # channels-last -> channels-first (tile kernel, nx = 3)
         ((64, 3, 224, 224), (0, 2, 3, 1)),      # channels-first -> channels-last
         ((33, 7, 5, 65, 3), (4, 2, 0, 3, 1)),
         ((1, 4099, 1, 3), (3, 2, 1, 0)),
         ((8, 16, 32, 48), (1, 0, 2, 3))]        # inner two axes merge: long contiguous runs


@pytest.mark.parametrize("shape,perm", LARGE, ids=[str(c) for c in LARGE])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64, dtypes.uint8, dtypes.int16])
def test_large_shapes_all_widths(shape, perm, dt, cuda_service):
    rng = np.random.default_rng(len(shape) + dt)
    npdt = NP[dt]
    if dt in (dtypes.float32, dtypes.float64):
        a = rng.standard_normal(shape).astype(npdt)
    else:
        a = rng.integers(0, 100, shape).astype(npdt)
    got = _run(cuda_service, a, perm, dt, offA=4, offB=8)
    assert np.array_equal(got, a.transpose(perm).reshape(-1))


def test_linear_algebra_transpose_nd(cuda_service, oracle_service):
    la, lo = LinearAlgebra(cuda_service), LinearAlgebra(oracle_service)
    rng = np.random.default_rng(2)
    x = rng.standard_normal((5, 6, 7, 8)).astype(np.float32)
    for perm in (None, [0, 2, 1, 3], [3, 0, 1, 2]):
        got, ref = la.transpose(la.array(x), perm), lo.transpose(lo.array(x), perm)
        assert got.shape() == ref.shape()
        assert np.array_equal(got.to_numpy(), np.asarray(ref.toArray(), np.float32))
    # a view of a larger buffer (offset != 0) and an explicit output
    big = la.array(rng.standard_normal((3, 4, 5, 6)).astype(np.float32))
    out = la.alloc([6, 5, 4], dtype=dtypes.float32)
    assert la.transpose(big[1], None, out) is out
    assert np.array_equal(out.to_numpy(), big.to_numpy()[1].transpose())
    xi = la.array(np.arange(24).reshape(2, 12), dtype=dtypes.int32)          # 2-D integers go through transposeND
    assert np.array_equal(la.transpose(xi).to_numpy(), np.arange(24).reshape(2, 12).T)
    with pytest.raises(InvalidArgumentException, match="unmatch sourceshape and perm"):
        la.transpose(la.array(x), [0, 1, 2])
    with pytest.raises(InvalidArgumentException, match="duplicate axis in perm option"):
        la.transpose(la.array(x), [0, 1, 1, 3])
    with pytest.raises(InvalidArgumentException, match="output shape must be transpose matrix of input."):
        la.transpose(la.array(x), None, la.alloc([8, 7, 6, 4]))


# ---------------------------------------------------------------------------------------------------
# slice / stick / repeat (PhpMath.php:2691-2776, :1400-1422): bit-exact copies, one rounded add per element
# ---------------------------------------------------------------------------------------------------
SLICE_CASES = [  # m, n, k, size, (start0,size0), (start1,size1), (start2,size2)
    (2, 4, 1, 3, (0, 2), (1, 2), (0, 1)),
    (5, 7, 3, 1, (1, 3), (2, 4), (1, 2)),               # item of one element: flat kernel
    (3, 2, 2, 64, (0, 3), (0, 2), (0, 2)),              # everything taken: one contiguous run
    (4, 6, 5, 16, (1, 2), (0, 6), (0, 5)),              # trailing axes whole: merged into the item
    (8, 33, 1, 1024, (0, 8), (5, 20), (0, 1)),          # long items: runs kernel, 16-byte words
    (2, 3, 4, 257, (1, 1), (1, 2), (2, 2)),             # odd item length: 4-byte words
    (64, 512, 1, 768, (0, 64), (17, 300), (0, 1)),      # 59 MB moved
]


@pytest.mark.parametrize("case", SLICE_CASES, ids=[str(c) for c in SLICE_CASES])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64, dtypes.int32, dtypes.uint8, dtypes.int64])
def test_slice_vs_oracle(case, dt, cuda_service):
    m, n, k, size, (st0, sz0), (st1, sz1), (st2, sz2) = case
    if m * n * k * size > (1 << 22) and dt not in (dtypes.float32, dtypes.uint8):
        pytest.skip("large case once per word size is enough")
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(m * 31 + size)
    offA, offY = 3, 5
    a = np.floor(rng.uniform(0, 100, offA + m * n * k * size)).astype(npdt)
    ny = sz0 * sz1 * sz2 * size
    math = cuda_service.math()
    args = (m, n, k, size)
    tail = (st0, sz0, st1, sz1, st2, sz2)
    for reverse in (False, True):
        for add in (False, True):
            y = np.floor(rng.uniform(0, 100, offY + ny)).astype(npdt)
            A = CudaBuffer(a.size, dt).from_numpy(a)
            Y = CudaBuffer(y.size, dt).from_numpy(y)
            math.slice(reverse, add, *args, A, offA, 1, Y, offY, 1, *tail)
            ao, yo = a.astype(np.float64), y.astype(np.float64)
            oracle.slice(reverse, add, *args, ao, offA, 1, yo, offY, 1, *tail)
            if dt == dtypes.uint8:
                ao, yo = np.mod(ao, 256), np.mod(yo, 256)                # the C buffer wraps
            assert np.array_equal(A.to_numpy().astype(np.float64), ao), (reverse, add)
            assert np.array_equal(Y.to_numpy().astype(np.float64), yo), (reverse, add)


def test_slice_strided_and_errors(cuda_service):
    math = cuda_service.math()
    rng = np.random.default_rng(5)
    a = rng.standard_normal(2 * 3 * 2 * 4 * 2 + 1).astype(np.float32)
    y = np.zeros(2 * 2 * 1 * 4 * 3 + 2, np.float32)
    A, Y = CudaBuffer(a.size, dtypes.float32).from_numpy(a), CudaBuffer(y.size, dtypes.float32).from_numpy(y)
    math.slice(False, False, 2, 3, 2, 4, A, 1, 2, Y, 2, 3, 0, 2, 1, 2, 1, 1)        # incA = 2, incY = 3
    ao, yo = a.astype(np.float64), y.astype(np.float64)
    oracle.slice(False, False, 2, 3, 2, 4, ao, 1, 2, yo, 2, 3, 0, 2, 1, 2, 1, 1)
    assert np.array_equal(Y.to_numpy().astype(np.float64), yo)
    with pytest.raises(InvalidArgumentException, match="Axis1 range is too large for source array."):
        math.slice(False, False, 2, 3, 2, 4, A, 0, 1, Y, 0, 1, 0, 2, 2, 2, 0, 1)
    with pytest.raises(InvalidArgumentException, match="BufferY size is too small"):
        math.slice(False, False, 2, 3, 2, 4, A, 0, 1, Y, 4, 1, 0, 2, 0, 3, 0, 2)
    with pytest.raises(InvalidArgumentException, match="unmatch BufferA size and m,n,k"):
        math.slice(False, False, 5, 3, 2, 4, A, 0, 1, Y, 0, 1, 0, 1, 0, 1, 0, 1)


@pytest.mark.parametrize("m,k,repeats", [(2, 3, 2), (1, 1, 7), (5, 1024, 3), (4096, 128, 16), (3, 257, 5)])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.int64, dtypes.uint8])
def test_repeat_vs_oracle(m, k, repeats, dt, cuda_service):
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(m + k)
    offA, offB = 2, 3
    a = np.floor(rng.uniform(0, 100, offA + m * k)).astype(npdt)
    A = CudaBuffer(a.size, dt).from_numpy(a)
    B = CudaBuffer(offB + m * repeats * k, dt)
    cuda_service.math().repeat(m, k, repeats, A, offA, B, offB)
    bo = np.zeros(offB + m * repeats * k)
    oracle.repeat(m, k, repeats, a.astype(np.float64), offA, bo, offB)
    assert np.array_equal(B.to_numpy().astype(np.float64), bo)
