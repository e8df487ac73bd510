"""CudaBuffer -- the Buffer the B200 driver hands to NDArray / LinearAlgebra / MatrixOperator.

Contract (reference): Interop\\Polite\\Math\\Matrix\\Buffer as implemented by PhpBuffer
(src/Drivers/MatlibPHP/PhpBuffer.php:19-190): ArrayAccess + Countable + dtype() + value_size()
+ dump()/load(); created by the factory call `$service->buffer()->Buffer($size,$dtype)`
(src/NDArrayPhp.php:148-151).  The reference's host code indexes buffers element-wise
(MatrixOperator.php:416,848,1179; NDArrayPhp.php:223,407,473), so an LV_ADVANCED buffer must
stay host-indexable while the data lives in HBM:

  * storage of record: device memory (rb200_buffer_alloc), dense, little-endian, typed;
  * a pinned host mirror is created lazily on the first host access;
  * two dirty flags -- host writes are flushed (h2d) before the next kernel touches the
    buffer, kernel writes are fetched (d2h) before the next host read.

Benchmarks and device-resident pipelines never touch elements from the host, so they never
pay for the mirror.  It also plays the DeviceBuffer role of src/NDArrayCL.php:261-270
(bytes(), read(), write(), copy(), fill()) with BYTE offsets, as that interface does.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _capi, dtypes
from .exceptions import InvalidArgumentException, RuntimeException


class CudaBuffer:
    def __init__(self, size: int, dtype: int, *, zero: bool = True, shareable: bool = False):
        if dtype not in dtypes.NUMPY:
            raise InvalidArgumentException("Invalid data type")          # PhpBuffer.php:96-98
        if size < 0:
            raise InvalidArgumentException("Size must be positive")
        _capi.init()
        self._size = int(size)
        self._dtype = int(dtype)
        self._np = np.dtype(dtypes.NUMPY[dtype])
        self._bytes = self._size * self._np.itemsize
        self._lib = _capi.load()
        ptr = ctypes.c_void_p()
        # shareable: plain cudaMalloc memory that rb200_ipc_export can hand to other processes (sharding.PeerRows);
        # everything else comes from the stream-ordered pool
        alloc = self._lib.rb200_buffer_alloc_shareable if shareable else self._lib.rb200_buffer_alloc
        _capi.check(alloc(self._bytes, ctypes.byref(ptr)))
        self._dptr = ptr.value
        self._hptr = None            # pinned mirror (lazy)
        self._host = None            # numpy view of the mirror
        self._host_dirty = False     # host mirror newer than device
        self._dev_dirty = False      # device newer than host mirror
        self._single_reads = 0       # element reads served by one-element d2h since the last device write
        self._upload_event = None    # recorded behind the asynchronous h2d of the mirror (flush)
        self._upload_pending = False # that copy may still be reading the pinned mirror
        if zero and self._size:
            z = np.zeros(1, self._np)
            _capi.check(self._lib.rb200_buffer_fill(self._dtype, self._dptr, 0, self._size, z.ctypes.data))
            self._dev_dirty = True

    @classmethod
    def adopt(cls, device_ptr: int, size: int, dtype: int, owner=None) -> "CudaBuffer":
        """A Buffer over device memory somebody else owns (e.g. a torch symmetric-memory allocation that is mapped
        into every rank and bound to an NVLS multicast object, sharding.SymmetricRows): same interface, never freed
        here; `owner` is kept alive with the buffer."""
        if dtype not in dtypes.NUMPY:
            raise InvalidArgumentException("Invalid data type")
        _capi.init()
        self = cls.__new__(cls)
        self._size, self._dtype = int(size), int(dtype)
        self._np = np.dtype(dtypes.NUMPY[dtype])
        self._bytes = self._size * self._np.itemsize
        self._lib = _capi.load()
        self._dptr, self._owner, self._adopted = int(device_ptr), owner, True
        self._hptr = self._host = None
        self._host_dirty, self._dev_dirty, self._single_reads = False, True, 0
        self._upload_event, self._upload_pending = None, False
        return self

    # ---- lifetime ------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_dptr", None):
                if not getattr(self, "_adopted", False):
                    self._lib.rb200_buffer_free(self._dptr)
                self._dptr = None
            if getattr(self, "_upload_event", None):
                if self._upload_pending:
                    self._lib.rb200_event_synchronize(self._upload_event)
                self._lib.rb200_event_destroy(self._upload_event)
                self._upload_event = None
            if getattr(self, "_hptr", None):
                self._lib.rb200_host_free(self._hptr)
                self._hptr = None
        except Exception:
            pass

    # ---- Buffer interface (PhpBuffer.php) ------------------------------------------------------
    def __len__(self) -> int:                      # Countable::count
        return self._size

    def count(self) -> int:
        return self._size

    def dtype(self) -> int:
        return self._dtype

    def value_size(self) -> int:
        return self._np.itemsize

    def _index(self, i) -> int:
        if not isinstance(i, (int, np.integer)) or i < 0 or i >= self._size:
            raise RuntimeException("Index invalid or out of range")      # SplFixedArray
        return int(i)

    def __getitem__(self, i):
        i = self._index(i)
        # a few scattered reads of a device-current buffer (LinearAlgebra::max reads $XX[$offX+$i] once,
        # LinearAlgebra.php:1644-1645) fetch single elements; a loop over the buffer gets the whole mirror
        if (self._host is None or self._dev_dirty) and not self._host_dirty and self._single_reads < 8:
            self._single_reads += 1
            one = np.empty(1, self._np)
            _capi.check(self._lib.rb200_buffer_d2h(one.ctypes.data, self._dptr, i * self._np.itemsize, self._np.itemsize))
            return bool(one[0]) if self._dtype == dtypes.bool_ else one[0].item()
        v = self.host()[i]
        return bool(v) if self._dtype == dtypes.bool_ else v.item()

    def _element_value(self, value):
        """Complex buffers take complex values / objects with real and imag, like PhpBuffer::offsetSet
        (PhpBuffer.php:124-133); everything else is stored as given."""
        if self._dtype in dtypes.COMPLEX_TYPES:
            if isinstance(value, (bool, int, float, np.integer, np.floating)) or not (hasattr(value, "real") and hasattr(value, "imag")):
                raise InvalidArgumentException("Cannot convert to complex number.: " + type(value).__name__)
            return complex(value.real, value.imag)
        return value

    def __setitem__(self, i, value) -> None:
        i = self._index(i)
        self.host()[i] = self._element_value(value)
        self._host_dirty = True

    def offsetExists(self, i) -> bool:
        return isinstance(i, (int, np.integer)) and 0 <= i < self._size

    def dump(self) -> bytes:                       # PhpBuffer.php:145-162: raw little-endian
        return self.host().tobytes()

    def load(self, string: bytes) -> None:         # PhpBuffer.php:164-189
        data = np.frombuffer(string, dtype=self._np)
        if data.size > self._size:
            raise RuntimeException("Unpack error")
        self.host()[: data.size] = data
        self._host_dirty = True

    # ---- DeviceBuffer face (byte offsets, NDArrayCL.php:261-270, 416-441) ----------------------
    def bytes(self) -> int:
        return self._bytes

    def read(self, host: np.ndarray, size: int = 0, offset: int = 0, host_offset: int = 0) -> None:
        """device -> host array (size/offset in bytes)."""
        self.flush()
        size = size or self._bytes - offset
        dst = host.ctypes.data + host_offset
        _capi.check(self._lib.rb200_buffer_d2h(dst, self._dptr, offset, size))

    def write(self, host: np.ndarray, size: int = 0, offset: int = 0, host_offset: int = 0) -> None:
        """host array -> device (size/offset in bytes)."""
        self.flush()
        size = size or min(host.nbytes - host_offset, self._bytes - offset)
        _capi.check(self._lib.rb200_buffer_h2d(self._dptr, offset, host.ctypes.data + host_offset, size))
        self._dev_dirty = True

    def copy(self, src: "CudaBuffer", size: int = 0, src_offset: int = 0, dst_offset: int = 0) -> None:
        src.flush()
        self.flush()
        size = size or src.bytes() - src_offset
        _capi.check(self._lib.rb200_buffer_d2d(self._dptr, dst_offset, src._dptr, src_offset, size))
        self._dev_dirty = True

    def fill(self, value, n: int | None = None, offset: int = 0) -> None:
        """fill n elements starting at ELEMENT offset with value."""
        self.flush()
        n = self._size - offset if n is None else n
        v = np.asarray([value], dtype=self._np)
        _capi.check(self._lib.rb200_buffer_fill(self._dtype, self._dptr, offset, n, v.ctypes.data))
        self._dev_dirty = True

    # ---- driver-internal -------------------------------------------------------------------------
    def _ensure_mirror(self) -> None:
        if self._host is None:
            ptr = ctypes.c_void_p()
            _capi.check(self._lib.rb200_host_alloc(max(self._bytes, 16), ctypes.byref(ptr)))
            self._hptr = ptr.value
            raw = (ctypes.c_uint8 * max(self._bytes, 1)).from_address(self._hptr)
            self._host = np.frombuffer(raw, dtype=self._np, count=self._size)
            self._dev_dirty = True      # whatever is on the device is the truth

    def host(self) -> np.ndarray:
        """The pinned host mirror, current.  Callers that write through it must call
        mark_host_written()."""
        self._ensure_mirror()
        self._wait_upload()
        if self._dev_dirty:
            if self._host_dirty:
                raise RuntimeException("CudaBuffer: host and device copies diverged")
            if self._bytes:
                _capi.check(self._lib.rb200_buffer_d2h(self._hptr, self._dptr, 0, self._bytes))
            self._dev_dirty = False
        return self._host

    def _wait_upload(self) -> None:
        """flush() queues the h2d of the pinned mirror asynchronously, possibly behind long kernels: the host may
        not modify the mirror before that copy has read it (X[0]=a; op(X); X[0]=b must show `a` to op)."""
        if self._upload_pending:
            _capi.check(self._lib.rb200_event_synchronize(self._upload_event))
            self._upload_pending = False

    def mark_host_written(self) -> None:
        """The caller wrote through host(); it must have obtained host() AFTER the last kernel touched the buffer
        (host() waits for a queued upload of the mirror)."""
        self._wait_upload()
        self._host_dirty = True

    # state the streamed (host-operand) GEMM needs: it moves the data itself, pipelined with the multiply
    def host_pending(self) -> bool:
        """The host mirror holds writes the device has not seen yet."""
        return self._host_dirty

    def has_mirror(self) -> bool:
        return self._host is not None

    def mirror_current(self) -> bool:
        return self._host is not None and not self._dev_dirty

    def raw_pointers(self):
        """(device pointer, pinned host pointer or None) WITHOUT any implied transfer."""
        return self._dptr, self._hptr

    def streamed_uploaded(self) -> None:
        self._host_dirty = False

    def streamed_result(self, downloaded: bool) -> None:
        self._dev_dirty = not downloaded

    def flush(self) -> None:
        """Make the device copy current (h2d of the mirror if the host wrote to it).  The copy is
        asynchronous on the library stream, i.e. ordered before the next kernel."""
        if self._host_dirty:
            if self._bytes:
                _capi.check(self._lib.rb200_buffer_h2d_async(self._dptr, 0, self._hptr, self._bytes))
                if self._upload_event is None:
                    ev = ctypes.c_void_p()
                    _capi.check(self._lib.rb200_event_create(ctypes.byref(ev)))
                    self._upload_event = ev.value
                _capi.check(self._lib.rb200_event_record(self._upload_event))
                self._upload_pending = True
            self._host_dirty = False

    def device_ptr(self) -> int:
        """Device pointer for a kernel that READS the buffer."""
        self.flush()
        return self._dptr

    def device_ptr_rw(self) -> int:
        """Device pointer for a kernel that WRITES the buffer (marks the mirror stale)."""
        self.flush()
        self._dev_dirty = True
        self._single_reads = 0
        return self._dptr

    # bulk host transfer helpers for tests / benches (typed, element offsets)
    def from_numpy(self, a: np.ndarray, offset: int = 0) -> "CudaBuffer":
        a = np.ascontiguousarray(a, dtype=self._np).reshape(-1)
        if offset < 0 or offset + a.size > self._size:
            raise InvalidArgumentException("array does not fit in buffer")
        CudaBuffer.write(self, a, a.nbytes, offset * self._np.itemsize)
        return self

    def to_numpy(self) -> np.ndarray:
        out = np.empty(self._size, self._np)
        if self._host_dirty:
            self.flush()
        _capi.check(self._lib.rb200_buffer_d2h(out.ctypes.data, self._dptr, 0, self._bytes)) if self._bytes else None
        return out

    def read_range(self, offset: int, n: int) -> np.ndarray:
        """Elements [offset, offset + n) as a fresh host array (one d2h of exactly those bytes)."""
        if offset < 0 or n < 0 or offset + n > self._size:
            raise InvalidArgumentException("range does not fit in buffer")
        out = np.empty(n, self._np)
        if n:
            CudaBuffer.read(self, out, n * self._np.itemsize, offset * self._np.itemsize)
        return out

    @property
    def __cuda_array_interface__(self) -> dict:
        self.flush()
        self._dev_dirty = True
        return {"shape": (self._size,), "typestr": self._np.str, "data": (self._dptr, False), "version": 3,
                "strides": None}


class CudaBufferFactory:
    """`bufferFactory` of AbstractMatlibService (src/Drivers/AbstractMatlibService.php:43-75,
    155-192): isAvailable() + Buffer(size, dtype)."""

    def isAvailable(self) -> bool:
        return _capi.device_available()

    def Buffer(self, size: int, dtype: int) -> CudaBuffer:
        return CudaBuffer(size, dtype)
