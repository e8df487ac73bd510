"""NDArrayCuda -- the device-resident N-d array of the B200 driver, "alongside NDArrayCL" (BASELINE north_star).

Mirror of Rindow\\Math\\Matrix\\NDArrayCL (src/NDArrayCL.php): constructor `(queue, buffer|None|host buffer, dtype,
shape, offset, flags, service)` :68-140 -- a DeviceBuffer is wrapped, None allocates, a host LinearBuffer is uploaded
(CL_MEM_COPY_HOST_PTR :120-133); `toNDArray()` is THE device -> host read (:252-272); `$a[$i]` / `$a[R(s,e)]` return
views that share the buffer, rank-0 included (:328-388); `$a[$i] = $deviceArray` is a device -> device copy and
anything else is a LogicException (:390-442); serialisation and offsetUnset are unsupported (:444-476); clone
copies the buffer on the device (:488-515).  Storage is a CudaDeviceBuffer: HBM only, no host mirror.

The queue is a CudaQueue (the library stream): the driver is stream-ordered, so the reference's event lists are
completion markers on that stream (CudaEventList) and `finish()` is rb200_sync.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _capi, dtypes
from .cuda_buffer import CudaBuffer
from .exceptions import InvalidArgumentException, LogicException, OutOfRangeException
from .ndarray import NDArray, NDArrayInterface

# OpenCL memory flags the reference passes around (Interop\\Polite\\Math\\Matrix\\OpenCL); only COPY_HOST_PTR has
# an effect here, the access flags are kept and reported back by flags()
CL_MEM_READ_WRITE, CL_MEM_WRITE_ONLY, CL_MEM_READ_ONLY, CL_MEM_COPY_HOST_PTR = 1, 2, 4, 32


class CudaContext:
    """What `$queue->getContext()` returns: the device the library is initialised on."""

    def __init__(self):
        self.info = _capi.device_info()

    def getInfo(self, what=None):
        return self.info


class CudaEventList:
    """Stream-ordered completion markers, the stand-in for the OpenCL EventList the reference threads through its
    device calls (`$events`, `$waitEvents`): record() marks "everything queued so far", wait() blocks the host."""

    def __init__(self):
        self._events = []

    def record(self) -> None:
        ev = ctypes.c_void_p()
        _capi.check(_capi.load().rb200_event_create(ctypes.byref(ev)))
        _capi.check(_capi.load().rb200_event_record(ev))
        self._events.append(ev)

    def copy(self, other: "CudaEventList") -> None:         # EventList::copy: append the other list's events
        self._events.extend(other._events)

    def count(self) -> int:
        return len(self._events)

    def __len__(self) -> int:
        return len(self._events)

    def wait(self) -> None:
        for ev in self._events:
            _capi.check(_capi.load().rb200_event_synchronize(ev))

    def done(self) -> bool:
        flag = ctypes.c_int(1)
        for ev in self._events:
            _capi.check(_capi.load().rb200_event_query(ev, ctypes.byref(flag)))
            if not flag.value:
                return False
        return True


class CudaQueue:
    """The command queue of the device classes = the library stream (rb200_get_stream)."""

    def __init__(self, options=None):
        _capi.init()
        self._context = CudaContext()
        self.options = options

    def getContext(self) -> CudaContext:
        return self._context

    def finish(self) -> None:
        _capi.sync()

    def flush(self) -> None:
        pass                                                 # launches are submitted eagerly


class CudaDeviceBuffer(CudaBuffer):
    """DeviceBuffer (src/NDArrayCL.php:147-168, 261-270, 416-441, 503-512): bytes() / dtype() / value_size() and
    queue-ordered read / write / copy / fill with BYTE sizes and offsets.  Lives in HBM only: there is no pinned
    mirror, every transfer is explicit.  (Single elements can still be peeked / poked -- `$buf[$i]` costs one small
    d2h / h2d -- which keeps the driver's host-read parameters working; bulk access goes through read / write.)"""

    def __init__(self, context, bytes_: int, flags: int = CL_MEM_READ_WRITE, hostBuffer=None, hostOffset: int = 0,
                 dtype: int | None = None):
        if dtype is None:
            dtype = hostBuffer.dtype() if hostBuffer is not None else dtypes.float32
        vs = dtypes.VALUE_SIZE[dtype]
        if bytes_ % vs:
            raise InvalidArgumentException("bytes is not a multiple of the value size")
        super().__init__(bytes_ // vs, dtype, zero=False)
        self._flags = flags
        if hostBuffer is not None:
            # CL_MEM_COPY_HOST_PTR semantics: the device buffer starts as a copy of hostBuffer[hostOffset:]
            self.write(None, hostBuffer, bytes_, 0, hostOffset * vs)

    # no mirror, ever
    def host(self):
        raise LogicException("CudaDeviceBuffer has no host mirror: use read() / NDArrayCuda::toNDArray()")

    def _ensure_mirror(self) -> None:
        raise LogicException("CudaDeviceBuffer has no host mirror")

    def __getitem__(self, i):
        i = self._index(i)
        one = np.empty(1, self._np)
        _capi.check(self._lib.rb200_buffer_d2h(one.ctypes.data, self._dptr, i * self._np.itemsize, self._np.itemsize))
        return bool(one[0]) if self._dtype == dtypes.bool_ else one[0].item()

    def __setitem__(self, i, value) -> None:
        i = self._index(i)
        one = np.asarray([self._element_value(value)], self._np)
        _capi.check(self._lib.rb200_buffer_h2d(self._dptr, i * self._np.itemsize, one.ctypes.data, self._np.itemsize))

    def dump(self) -> bytes:
        return self.to_numpy().tobytes()

    def load(self, string: bytes) -> None:
        data = np.frombuffer(string, dtype=self._np)
        self.from_numpy(data)

    @staticmethod
    def _host_view(hostBuf):
        """(numpy array or None, CudaBuffer or None) of a transfer partner."""
        if isinstance(hostBuf, np.ndarray):
            return hostBuf, None
        if isinstance(hostBuf, CudaBuffer):
            return None, hostBuf
        raise InvalidArgumentException("host buffer must be a numpy array or a driver Buffer")

    def read(self, queue, hostBuf, size: int = 0, offset: int = 0, hostoffset: int = 0, blocking_read=None,
             events: CudaEventList | None = None, waitEvents: CudaEventList | None = None) -> None:
        """device -> host buffer (NDArrayCL.php:261-270).  A host-indexable CudaBuffer (the LV_ADVANCED buffer of an
        NDArray) receives the bytes on the device side and fetches them on its next host access."""
        if waitEvents is not None:
            waitEvents.wait()
        size = size or self._bytes - offset
        arr, buf = self._host_view(hostBuf)
        if arr is not None:
            _capi.check(self._lib.rb200_buffer_d2h(arr.ctypes.data + hostoffset, self._dptr, offset, size))
        else:
            buf.copy(self, size, offset, hostoffset)
            if blocking_read is None or blocking_read:
                _capi.sync()
        if events is not None:
            events.record()

    def write(self, queue, hostBuf, size: int = 0, offset: int = 0, hostoffset: int = 0, blocking_write=None,
              events: CudaEventList | None = None, waitEvents: CudaEventList | None = None) -> None:
        if waitEvents is not None:
            waitEvents.wait()
        arr, buf = self._host_view(hostBuf)
        if arr is not None:
            size = size or min(arr.nbytes - hostoffset, self._bytes - offset)
            _capi.check(self._lib.rb200_buffer_h2d(self._dptr, offset, arr.ctypes.data + hostoffset, size))
        else:
            size = size or min(buf.bytes() - hostoffset, self._bytes - offset)
            CudaBuffer.copy(self, buf, size, hostoffset, offset)
        if events is not None:
            events.record()

    def copy(self, queue, src=None, size: int = 0, src_offset: int = 0, dst_offset: int = 0,
             events: CudaEventList | None = None, waitEvents: CudaEventList | None = None) -> None:
        """DeviceBuffer::copy(queue, src, size, srcOffset, dstOffset, events, waitEvents) (NDArrayCL.php:416-441)."""
        if isinstance(queue, CudaBuffer):                     # called with the CudaBuffer signature (no queue argument)
            queue, src, size, src_offset, dst_offset = None, queue, (src or 0), size, src_offset
        if waitEvents is not None:
            waitEvents.wait()
        CudaBuffer.copy(self, src, size, src_offset, dst_offset)
        if events is not None:
            events.record()

    def flags(self) -> int:
        return self._flags


class CudaDeviceBufferFactory:
    """`$service->buffer(Service::LV_ACCELERATED)` (AbstractMatlibService.php:316-318): Buffer(context, bytes, flags,
    hostBuffer, hostOffset, dtype), the call NDArrayCL::newBuffer makes (NDArrayCL.php:147-168)."""

    def isAvailable(self) -> bool:
        return _capi.device_available()

    def Buffer(self, context, bytes_: int, flags: int = CL_MEM_READ_WRITE, hostBuffer=None, hostOffset: int = 0,
               dtype: int | None = None) -> CudaDeviceBuffer:
        return CudaDeviceBuffer(context, bytes_, flags, hostBuffer, hostOffset, dtype)


class NDArrayCuda(NDArrayInterface):
    def __init__(self, queue, buffer=None, dtype: int | None = None, shape=None, offset: int | None = None,
                 flags: int | None = None, service=None):
        if service is None:
            raise InvalidArgumentException("No service specified.")                       # NDArrayCL.php:77-79
        self._service = service
        self._factory = service.buffer(getattr(service, "LV_ACCELERATED", 3))
        self._queue = queue
        self._context = queue.getContext()
        dtype = dtypes.float32 if dtype is None else dtype
        offset = 0 if offset is None else offset
        flags = CL_MEM_READ_WRITE if flags is None else flags
        if shape is None:
            raise InvalidArgumentException("Invalid dimension size")
        self._assert_shape(shape)
        self._shape = tuple(int(s) for s in shape)
        self._flags = flags
        self._events = None
        size = self.size()
        if isinstance(buffer, CudaDeviceBuffer) or (hasattr(buffer, "bytes") and hasattr(buffer, "read")
                                                   and getattr(buffer, "is_device_buffer", False)):
            if buffer.bytes() < (size + offset) * dtypes.VALUE_SIZE[dtype]:
                raise InvalidArgumentException("Invalid dimension size")
            self._dtype, self._buffer, self._offset = dtype, buffer, offset
        elif buffer is None:
            self._buffer = self._factory.Buffer(self._context, size * dtypes.VALUE_SIZE[dtype], flags, None, 0, dtype)
            self._dtype, self._offset = dtype, 0
        elif isinstance(buffer, CudaBuffer) or hasattr(buffer, "dtype"):
            if size > len(buffer) - offset:
                raise InvalidArgumentException("host buffer is too small")
            self._buffer = self._factory.Buffer(self._context, size * dtypes.VALUE_SIZE[buffer.dtype()], flags, buffer,
                                                offset, buffer.dtype())
            self._dtype, self._offset = buffer.dtype(), 0
        else:
            raise InvalidArgumentException("Invalid type of array: " + type(buffer).__name__)

    @staticmethod
    def _assert_shape(shape) -> None:                                                   # NDArrayCL.php:170-186
        for num in shape:
            if not isinstance(num, (int, np.integer)) or isinstance(num, bool):
                raise InvalidArgumentException("Invalid shape numbers. It gives " + type(num).__name__)
            if num <= 0:
                raise InvalidArgumentException("Invalid shape numbers. It gives %d" % num)

    def _like(self, shape, offset) -> "NDArrayCuda":
        return NDArrayCuda(self._queue, self._buffer, self._dtype, list(shape), offset, self._flags, self._service)

    # ---- NDArray interface ---------------------------------------------------------------------------------
    def service(self):
        return self._service

    def shape(self) -> list:
        return list(self._shape)

    def ndim(self) -> int:
        return len(self._shape)

    def dtype(self) -> int:
        return self._dtype

    def flags(self) -> int:
        return self._flags

    def buffer(self) -> CudaDeviceBuffer:
        return self._buffer

    def offset(self) -> int:
        return self._offset

    def valueSize(self) -> int:
        return dtypes.VALUE_SIZE[self._dtype]

    def size(self) -> int:
        return int(np.prod(self._shape, dtype=np.int64)) if self._shape else 1

    def reshape(self, shape) -> "NDArrayCuda":
        self._assert_shape(shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size():
            raise InvalidArgumentException(
                "Unmatch size to reshape: [%s]=>[%s]" % (",".join(map(str, self._shape)), ",".join(map(str, shape))))
        return self._like(shape, self._offset)

    def toNDArray(self, blocking_read=None, events=None, waitEvents=None) -> NDArray:      # NDArrayCL.php:252-272
        """The device -> host read: a host NDArray holding a copy of this view."""
        out = NDArray(None, dtype=self._dtype, shape=self._shape, service=self._service)
        vs = self.valueSize()
        self._buffer.read(self._queue, out.buffer(), self.size() * vs, self._offset * vs, 0,
                          True if blocking_read is None else blocking_read, events, waitEvents)
        return out

    def toArray(self):
        return self.toNDArray().toArray()

    def to_numpy(self) -> np.ndarray:
        if not hasattr(self._buffer, "read_range"):               # a non-CUDA device buffer: go through toNDArray
            return self.toNDArray().to_numpy()
        flat = self._buffer.read_range(self._offset, self.size())
        out = flat.reshape(self._shape)
        return out.astype(bool) if self._dtype == dtypes.bool_ else out

    # ---- views (NDArrayCL.php:274-388) --------------------------------------------------------------------
    def _range(self, i):
        if isinstance(i, slice):
            if i.step not in (None, 1):
                raise OutOfRangeException("Illegal range specification.:delta=%s" % i.step)
            start = 0 if i.start is None else i.start
            limit = self._shape[0] if i.stop is None else i.stop
            if start >= limit:
                raise OutOfRangeException("Illegal range specification.:[%d,%d]" % (start, limit))
            return start, limit
        if isinstance(i, (list, tuple)):
            if len(i) != 2 or i[0] > i[1]:
                raise OutOfRangeException("Illegal range specification.")
            return int(i[0]), int(i[1])
        if isinstance(i, (int, np.integer)) and not isinstance(i, bool):
            return int(i), int(i) + 1
        raise InvalidArgumentException("Array offsets must be integers or ranges.")

    def offsetExists(self, i) -> bool:
        if not self._shape:
            return False
        start, limit = self._range(i)
        return start >= 0 and limit <= self._shape[0]

    def __getitem__(self, i) -> "NDArrayCuda":
        if not self._shape:
            raise OutOfRangeException("This object is scalar.")
        start, limit = self._range(i)
        if start < 0 or limit > self._shape[0]:
            raise OutOfRangeException("Index is out of range. range allows (0,%d): %s given." % (self._shape[0], i))
        rest = self._shape[1:]
        item = int(np.prod(rest, dtype=np.int64)) if rest else 1
        if isinstance(i, (int, np.integer)):
            return self._like(rest, self._offset + start * item)           # rank-0 included: still a device array
        return self._like((limit - start,) + rest, self._offset + start * item)

    def __setitem__(self, i, value) -> None:                                # NDArrayCL.php:390-442
        if not self._shape:
            raise OutOfRangeException("This object is scalar.")
        if not isinstance(i, (int, np.integer)) or isinstance(i, bool):
            raise OutOfRangeException("Unsuppored to set for range specification.")
        if i < 0 or i >= self._shape[0]:
            raise OutOfRangeException("Index is out of range")
        rest = self._shape[1:]
        if not isinstance(value, NDArrayCuda):
            raise LogicException("Must be NDArray on CUDA." if not rest else "Unmatch shape numbers")
        vs = dtypes.VALUE_SIZE[value.dtype()]
        if not rest:
            self._buffer.copy(self._queue, value.buffer(), vs, value.offset() * vs, (self._offset + i) * vs)
            return
        if tuple(value.shape()) != tuple(rest):
            raise LogicException("Unmatch shape numbers")
        size = int(np.prod(rest, dtype=np.int64))
        self._buffer.copy(self._queue, value.buffer(), size * vs, value.offset() * vs, (self._offset + i * size) * vs)

    def __delitem__(self, i) -> None:
        raise LogicException("Unsuppored Operation")

    def __len__(self) -> int:
        if not self._shape:
            raise InvalidArgumentException("scalar has no count")
        return self._shape[0]

    def count(self) -> int:
        return len(self)

    def __iter__(self):
        for i in range(self._shape[0]):
            yield self[i]

    # ---- events / serialisation / clone (NDArrayCL.php:444-515) -------------------------------------------
    def setEvents(self, events) -> None:
        self._events = events

    def getEvents(self):
        return self._events

    def setPortableSerializeMode(self, mode: bool) -> None:
        raise LogicException("Unsuppored Operation")

    def getPortableSerializeMode(self) -> bool:
        return False

    def __reduce__(self):
        raise LogicException("Unsuppored Operation")

    def __copy__(self) -> "NDArrayCuda":
        return self.__deepcopy__(None)

    def __deepcopy__(self, memo) -> "NDArrayCuda":
        """`clone $array`: a new device buffer with the same bytes, copied on the device (NDArrayCL.php:488-515)."""
        flags = self._flags & ~CL_MEM_COPY_HOST_PTR
        nb = self._factory.Buffer(self._context, self._buffer.bytes(), flags, None, 0, self._buffer.dtype())
        nb.copy(self._queue, self._buffer, 0, 0, 0)
        self._queue.finish()
        return NDArrayCuda(self._queue, nb, self._dtype, list(self._shape), self._offset, flags, self._service)
