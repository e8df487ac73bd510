"""NDArray -- host-facing N-d array over a driver Buffer.

Mirror of the part of Rindow\\Math\\Matrix\\NDArrayPhp (src/NDArrayPhp.php) the hot path touches:
construction from nested arrays / shape (:60-151: the buffer comes from
`$service->buffer()->Buffer($size,$dtype)` :148-151), views by offset (`$a[$i]` :383-448),
reshape (:304-318), toArray (:320-352).  Over the B200 service the Buffer is a CudaBuffer
(device-resident, host-indexable), which is what the task calls "a device NDArray alongside
NDArrayCL": toNDArray()/to_numpy() are the device->host reads (cf. src/NDArrayCL.php:252-272).
Driver-agnostic by construction -- the parity tests run the same class over the oracle's buffers.
"""
from __future__ import annotations

import numpy as np

from . import dtypes
from .exceptions import InvalidArgumentException, OutOfRangeException


class NDArrayInterface:
    """Interop\\Polite\\Math\\Matrix\\NDArray: what LinearAlgebra accepts -- the host-facing NDArray below and the
    device-resident NDArrayCuda (ndarray_cuda.py) both implement it, as NDArrayPhp and NDArrayCL do in the reference."""


class NDArray(NDArrayInterface):
    def __init__(self, array=None, dtype: int | None = None, shape=None, offset: int | None = None,
                 service=None, buffer=None):
        if service is None:
            raise InvalidArgumentException("No service specified.")           # NDArrayPhp.php:66-68
        self._service = service
        self._dtype = dtypes.float32 if dtype is None else dtype
        if buffer is not None:                      # view / wrap (NDArrayPhp.php:71-80)
            self._buffer = buffer
            self._offset = offset or 0
            self._shape = tuple(int(s) for s in shape)
            if buffer.dtype() != self._dtype:
                raise InvalidArgumentException("Unmatch dtype and buffer dtype")
            if len(buffer) - self._offset < self.size():
                raise InvalidArgumentException("Invalid dimension size")
            return
        if array is not None:
            a = np.asarray(array)
            if a.dtype == object:
                raise InvalidArgumentException("The shape of the dimension is broken")
            if shape is not None and int(np.prod(shape)) != a.size:
                raise InvalidArgumentException("Unmatch shape and array")
            self._shape = tuple(a.shape if shape is None else shape)
            self._offset = 0
            self._buffer = service.buffer().Buffer(int(a.size), self._dtype)
            self._store(a)
        elif shape is not None:
            self._shape = tuple(int(s) for s in shape)
            for s in self._shape:
                if s < 0:
                    raise InvalidArgumentException("Invalid dimension size")
            self._offset = 0
            self._buffer = service.buffer().Buffer(self.size(), self._dtype)
        else:
            raise InvalidArgumentException("Invalid type of array")

    def _store(self, a: np.ndarray) -> None:
        flat = np.ascontiguousarray(a).reshape(-1)
        if hasattr(self._buffer, "from_numpy"):
            self._buffer.from_numpy(flat.astype(dtypes.NUMPY[self._dtype], copy=False), self._offset)
        else:
            for i, v in enumerate(flat.tolist()):
                self._buffer[self._offset + i] = v

    # ---- NDArray interface ---------------------------------------------------------------
    def shape(self) -> list:
        return list(self._shape)

    def ndim(self) -> int:
        return len(self._shape)

    def dtype(self) -> int:
        return self._dtype

    def buffer(self):
        return self._buffer

    def offset(self) -> int:
        return self._offset

    def size(self) -> int:
        return int(np.prod(self._shape, dtype=np.int64)) if self._shape else 1

    def service(self):
        return self._service

    def reshape(self, shape) -> "NDArray":
        shape = tuple(int(s) for s in shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size():
            raise InvalidArgumentException(
                "Unmatch size to reshape: [%s]=>[%s]" % (",".join(map(str, self._shape)), ",".join(map(str, shape))))
        return NDArray(dtype=self._dtype, shape=shape, offset=self._offset, service=self._service,
                       buffer=self._buffer)

    def to_numpy(self) -> np.ndarray:
        n = self.size()
        if hasattr(self._buffer, "read_range"):                  # device buffer: only the view's bytes travel
            flat = self._buffer.read_range(self._offset, n)
        elif hasattr(self._buffer, "to_numpy"):
            flat = self._buffer.to_numpy()[self._offset:self._offset + n]
        else:
            flat = np.array([self._buffer[self._offset + i] for i in range(n)])
        out = np.asarray(flat).reshape(self._shape)
        return out.astype(bool) if self._dtype == dtypes.bool_ else out

    def toArray(self):
        return self.to_numpy().tolist()

    def toNDArray(self) -> "NDArray":          # device -> host, trivially itself (host-indexable buffer)
        return self

    def __len__(self) -> int:
        if not self._shape:
            raise InvalidArgumentException("scalar has no count")
        return self._shape[0]

    def __getitem__(self, i):
        if not self._shape:
            raise OutOfRangeException("This object is scalar.")
        if isinstance(i, slice):
            start, stop, step = i.indices(self._shape[0])
            if step != 1 or stop < start:
                raise OutOfRangeException("Illegal range specification.")
            rest = int(np.prod(self._shape[1:], dtype=np.int64)) if len(self._shape) > 1 else 1
            return NDArray(dtype=self._dtype, shape=(stop - start,) + self._shape[1:],
                           offset=self._offset + start * rest, service=self._service, buffer=self._buffer)
        if i < 0 or i >= self._shape[0]:
            raise OutOfRangeException("Index is out of range")
        if len(self._shape) == 1:
            return self._buffer[self._offset + i]
        rest = int(np.prod(self._shape[1:], dtype=np.int64))
        return NDArray(dtype=self._dtype, shape=self._shape[1:], offset=self._offset + i * rest,
                       service=self._service, buffer=self._buffer)

    def __setitem__(self, i, value) -> None:
        if len(self._shape) == 1 and not isinstance(value, NDArray):
            if i < 0 or i >= self._shape[0]:
                raise OutOfRangeException("Index is out of range")
            self._buffer[self._offset + i] = value
            return
        if len(self._shape) == 1:
            # `$a[$i] = $ndarray` on a rank-1 array: the reference only takes scalars there (NDArrayPhp.php:462-474)
            if i < 0 or i >= self._shape[0]:
                raise OutOfRangeException("Index is out of range")
            raise InvalidArgumentException("Must be scalar type")
        sub = self[i]
        src = value.to_numpy() if isinstance(value, NDArray) else np.asarray(value)
        if tuple(src.shape) != tuple(sub._shape):
            raise InvalidArgumentException("Unmatch shape numbers")
        sub._store(src)
