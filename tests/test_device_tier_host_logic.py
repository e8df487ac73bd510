"""Host logic of the device-resident tier on CPU: NDArrayCuda / LinearAlgebraCuda (mirrors of src/NDArrayCL.php and
src/LinearAlgebraCL.php) run over the oracle service's numpy-backed DeviceBuffer stand-in, so the view / offset /
copy / download rules, the scalar conventions and blocking() are covered without a GPU.  The same assertions run on
the B200 in tests/test_gpu_device_ndarray.py."""
import copy

import numpy as np
import pytest

import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import LinearAlgebraCuda, NDArray, NDArrayCuda, dtypes
from rindow_math_matrix_b200.ndarray_cuda import CL_MEM_COPY_HOST_PTR, CL_MEM_READ_ONLY, CL_MEM_READ_WRITE


def device(svc, host, flags=CL_MEM_READ_WRITE):
    queue = svc.createQueue()
    return queue, NDArrayCuda(queue, host.buffer(), host.dtype(), host.shape(), host.offset(), flags | CL_MEM_COPY_HOST_PTR,
                              service=svc)


def test_offset_get_views_share_the_buffer(oracle_service):        # NDArrayCLTest.php:56-98
    svc = oracle_service
    host = NDArray([[1, 2], [3, 4], [5, 6]], service=svc)
    queue, a = device(svc, host, CL_MEM_READ_ONLY)
    assert a.shape() == [3, 2] and a.dtype() == dtypes.float32 and a.size() == 6 and a.ndim() == 2
    assert a.flags() == CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR and a.buffer().bytes() == 24
    assert a[1].toNDArray().toArray() == [3, 4] and a[1].buffer() is a.buffer() and a[1].offset() == 2
    assert a[1:3].shape() == [2, 2] and a[1:3].toNDArray().toArray() == [[3, 4], [5, 6]]
    assert a[[1, 3]].toArray() == [[3, 4], [5, 6]]
    s = a[1][1]
    assert isinstance(s, NDArrayCuda) and s.shape() == [] and s.toArray() == 4
    with pytest.raises(rb.OutOfRangeException):
        a[3]
    with pytest.raises(rb.OutOfRangeException, match="scalar"):
        s[0]
    assert a.reshape([2, 3]).toArray() == [[1, 2, 3], [4, 5, 6]]
    with pytest.raises(rb.InvalidArgumentException, match="Unmatch size to reshape"):
        a.reshape([5])
    with pytest.raises(rb.InvalidArgumentException, match="No service specified"):
        NDArrayCuda(queue, None, dtypes.float32, [2])
    with pytest.raises(rb.InvalidArgumentException, match="Invalid dimension size"):
        NDArrayCuda(queue, None, dtypes.float32, None, service=svc)
    with pytest.raises(rb.InvalidArgumentException, match="Invalid shape numbers"):
        NDArrayCuda(queue, None, dtypes.float32, [2, 0], service=svc)
    with pytest.raises(rb.InvalidArgumentException, match="Invalid dimension size"):
        NDArrayCuda(queue, a.buffer(), dtypes.float32, [7], service=svc)       # device buffer too small


def test_offset_set_is_a_device_copy(oracle_service):             # NDArrayCLTest.php:100-123
    svc = oracle_service
    queue, a = device(svc, NDArray([[1, 2], [3, 4], [5, 6]], service=svc))
    _, src = device(svc, NDArray([11, 12], service=svc), CL_MEM_READ_ONLY)
    a[1] = src
    assert a.toArray() == [[1, 2], [11, 12], [5, 6]]
    _, s = device(svc, NDArray(np.asarray(4.0), shape=[], service=svc), CL_MEM_READ_ONLY)
    a[1][1] = s
    assert a.toArray() == [[1, 2], [11, 4], [5, 6]]
    with pytest.raises(rb.LogicException):
        a[1][0] = 3.0
    with pytest.raises(rb.LogicException, match="Unmatch shape numbers"):
        a[0] = a
    with pytest.raises(rb.OutOfRangeException):
        a[0:1] = src
    with pytest.raises(rb.LogicException, match="Unsuppored Operation"):
        del a[0]
    with pytest.raises(rb.LogicException):
        a.setPortableSerializeMode(True)


def test_clone_copies_the_buffer(oracle_service):                 # NDArrayCLTest.php:125-144
    svc = oracle_service
    queue, a = device(svc, NDArray([[1, 2], [3, 4]], dtype=dtypes.int32, service=svc), CL_MEM_READ_ONLY)
    b = copy.copy(a)
    a.buffer().write(queue, NDArray([[0, 0], [0, 0]], dtype=dtypes.int32, service=svc).buffer())
    assert a.toArray() == [[0, 0], [0, 0]] and b.toArray() == [[1, 2], [3, 4]]
    assert b.buffer() is not a.buffer() and b.dtype() == dtypes.int32 and queue.finished >= 1
    assert b.flags() & CL_MEM_COPY_HOST_PTR == 0


def test_linear_algebra_cuda_conventions(oracle_service):         # LinearAlgebraCL.php:145-152, 220-253, 321-368, 441-460
    svc = oracle_service
    queue = svc.createQueue()
    la = LinearAlgebraCuda(queue, svc)
    assert la.accelerated() and la.fp64() and la.getQueue() is queue and la.blocking() is False
    x = la.array([[1.0, -4.0], [2.5, 0.5]])
    assert isinstance(x, NDArrayCuda) and la.array(x) is x
    assert isinstance(la.alloc([2, 2]), NDArrayCuda) and isinstance(la.allocHost([2, 2]), NDArray)
    h = la.toNDArray(x)
    assert isinstance(h, NDArray) and h.toArray() == [[1.0, -4.0], [2.5, 0.5]] and la.toNDArray(h) is h
    # every operation is inherited: results stay device arrays
    y = la.add(la.array([1.0, 1.0]), la.copy(x))
    assert isinstance(y, NDArrayCuda) and y.toArray() == [[2.0, -3.0], [3.5, 1.5]]
    s = la.reduceSum(x, axis=1)
    assert isinstance(s, NDArrayCuda) and s.toArray() == [-3.0, 3.0]
    # scalars: rank-0 device arrays by default, numbers with scalarNumeric(true) (LinearAlgebraCL.php:2526-2529)
    r = la.sum(x)
    assert isinstance(r, NDArrayCuda) and r.shape() == [] and la.scalar(r) == 0.0
    i = la.imax(x)
    assert isinstance(i, NDArrayCuda) and i.dtype() == dtypes.int32 and la.scalar(i) == 2
    la.scalarNumeric(True)
    assert la.sum(x) == 0.0 and la.imax(x) == 2 and la.max(x) == 2.5 and la.asum(x) == 8.0 and la.amax(x) == 4.0
    # blocking(true) finishes the queue after every operation
    before = queue.finished
    la.blocking(True)
    la.scal(2.0, x)
    assert queue.finished == before + 1 and x.toArray() == [[2.0, -8.0], [5.0, 1.0]]
    ev = la.newEventList()
    ev.record()
    assert len(ev) == 1
    x[0] = la.array([7.0, 8.0])
    assert x.toArray() == [[7.0, 8.0], [5.0, 1.0]]
