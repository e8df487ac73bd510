"""Parity of einsum / einsum4p1 (rb200_einsum) through the C ABI against the CPU oracle: bit-exact -- the device adds
the exact float32 products in double in the reference's order (PhpMath.php:3321-3440), one rounding at the store;
float64 uses separate multiply and add like PHP.  pytest -m gpu"""
import numpy as np
import pytest

import oracle
from rindow_math_matrix_b200 import CudaBuffer, InvalidArgumentException, dtypes

pytestmark = pytest.mark.gpu


def _strides(shape):
    st, acc = [], 1
    for d in reversed(shape):
        st.append(acc)
        acc *= d
    return list(reversed(st))


# (sizes of all index axes, ndimC, which axes A has, which axes B has) -- e.g. "ij,jk->ik": axes (i,k,j), ndimC 2
CASES = [
    ([3, 4, 5], 2, [0, 2], [2, 1]),                      # ij,jk->ik
    ([7], 0, [0], [0]),                                   # i,i->   (dot)
    ([6, 5], 2, [0], [1]),                                # i,j->ij (outer)
    ([2, 3, 4, 5, 6], 4, [0, 1, 2, 4], [0, 3, 4]),        # batched attention-like contraction
    ([4, 8, 16, 3, 32], 3, [0, 1, 3, 4], [2, 3, 4]),      # two summed axes
    ([64, 64, 300], 2, [0, 2], [2, 1]),                   # longer sums: order matters in float
    ([5, 1, 9], 2, [0, 2], [1, 2]),
]


@pytest.mark.parametrize("sizes,ndimC,axesA,axesB", CASES)
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64])
def test_einsum_vs_oracle_bit_exact(sizes, ndimC, axesA, axesB, dt, cuda_service):
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(sum(sizes))
    depth = len(sizes)
    shapeA, shapeB = [sizes[a] for a in axesA], [sizes[a] for a in axesB]
    ldA, ldB = [0] * depth, [0] * depth
    for ax, st in zip(axesA, _strides(shapeA)):
        ldA[ax] = st
    for ax, st in zip(axesB, _strides(shapeB)):
        ldB[ax] = st
    offA, offB, offC = 2, 3, 1
    a = rng.standard_normal(offA + int(np.prod(shapeA))).astype(npdt)
    b = rng.standard_normal(offB + int(np.prod(shapeB))).astype(npdt)
    sizeC = int(np.prod(sizes[:ndimC])) if ndimC else 1
    c = np.full(offC + sizeC + 2, 5).astype(npdt)
    A, B, C = (CudaBuffer(v.size, dt).from_numpy(v) for v in (a, b, c))
    cuda_service.math().einsum(sizes, A, offA, ldA, B, offB, ldB, C, offC, ndimC)
    co = c.astype(np.float64)
    oracle.einsum(sizes, a.astype(np.float64), offA, ldA, b.astype(np.float64), offB, ldB, co, offC, ndimC)
    assert np.array_equal(C.to_numpy().astype(np.float64), co.astype(npdt).astype(np.float64))
    # and against numpy's einsum within rounding
    letters = "abcdefgh"
    spec = "".join(letters[x] for x in axesA) + "," + "".join(letters[x] for x in axesB) + "->" + letters[:ndimC]
    ref = np.einsum(spec, a[offA:].astype(np.float64).reshape(shapeA), b[offB:].astype(np.float64).reshape(shapeB))
    got = C.to_numpy()[offC:offC + sizeC].astype(np.float64)
    assert np.allclose(got, ref.reshape(-1), rtol=1e-5 if dt == dtypes.float32 else 1e-12, atol=1e-5 if dt == dtypes.float32 else 1e-12)


@pytest.mark.parametrize("dims", [(2, 3, 4, 5, 6), (1, 1, 1, 1, 1), (4, 8, 2, 16, 64)])
def test_einsum4p1_vs_oracle_bit_exact(dims, cuda_service):
    """einsum4p1 'abce,adce->abcd'-style strides; as in the reference C is stored from element 0 (PhpMath.php:3432)"""
    rng = np.random.default_rng(sum(dims))
    d0, d1, d2, d3, d4 = dims
    # A[d0,d1,d2,d4] lacks axis 3; B[d0,d3,d2,d4] lacks axis 1
    ldA = [d1 * d2 * d4, d2 * d4, d4, 0, 1]
    ldB = [d3 * d2 * d4, 0, d4, d2 * d4, 1]
    a = rng.standard_normal(1 + d0 * d1 * d2 * d4).astype(np.float32)
    b = rng.standard_normal(2 + d0 * d3 * d2 * d4).astype(np.float32)
    sizeC = d0 * d1 * d2 * d3
    c = np.zeros(sizeC + 3, np.float32)
    A, B, C = (CudaBuffer(v.size, dtypes.float32).from_numpy(v) for v in (a, b, c))
    cuda_service.math().einsum4p1(d0, d1, d2, d3, d4, A, 1, *ldA, B, 2, *ldB, C, 3)
    co = c.astype(np.float64)
    oracle.einsum(list(dims), a.astype(np.float64), 1, ldA, b.astype(np.float64), 2, ldB, co, 0, 4)
    assert np.array_equal(C.to_numpy().astype(np.float64), co.astype(np.float32).astype(np.float64))
    with pytest.raises(InvalidArgumentException, match="buffer C"):
        cuda_service.math().einsum4p1(d0, d1, d2, d3, d4, A, 1, *ldA, B, 2, *ldB, C, 4)
    with pytest.raises(InvalidArgumentException, match="buffer A"):
        cuda_service.math().einsum4p1(d0, d1, d2, d3, d4, A, 2, *ldA, B, 2, *ldB, C, 0)
