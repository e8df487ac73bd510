"""ctypes binding of librindow_b200.so -- the same C ABI (include/rindow_b200.h) a PHP-FFI
shim would bind (php/MatlibCuda/CudaFFI.php).  There is deliberately NO fallback: if the
shared library or a B200 is missing every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librindow_b200.so")

OK, E_INVALID, E_CUDA, E_UNSUPPORTED, E_NOTINIT, E_NOMEM, E_RANGE = 0, -1, -2, -3, -4, -5, -6

# rb200_broadcast / rb200_unary / rb200_compare op codes (include/rindow_b200.h)
BC_MAXIMUM, BC_MINIMUM, BC_GREATER, BC_GREATER_EQUAL, BC_LESS, BC_LESS_EQUAL, BC_POW, BC_DUPLICATE = range(8)
(UN_INCREMENT, UN_RECIPROCAL, UN_SQUARE, UN_SQRT, UN_RSQRT, UN_EXP, UN_LOG, UN_TANH, UN_SIN, UN_COS, UN_TAN,
 UN_NOT, UN_NAN2NUM, UN_ISNAN) = range(14)
CMP_EQUAL, CMP_NOT_EQUAL = 0, 1

GEMM_AUTO, GEMM_FP32_SIMT, GEMM_TF32, GEMM_TF32X3 = 0, 1, 2, 3
GEMM_MODE_NAMES = {GEMM_FP32_SIMT: "fp32_simt", GEMM_TF32: "tf32", GEMM_TF32X3: "tf32x3"}

i32, i64, f32, f64, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/rindow_b200.h declares
SIGNATURES = {
    "rb200_abi_version": (i32, []),
    "rb200_last_error": (ctypes.c_char_p, []),
    "rb200_device_count": (i32, [ctypes.POINTER(i32)]),
    "rb200_init": (i32, [i32]),
    "rb200_shutdown": (i32, []),
    "rb200_device_info": (i32, [ctypes.POINTER(i32)] * 4 + [ctypes.POINTER(i64)]),
    "rb200_set_stream": (i32, [vp]),
    "rb200_get_stream": (i32, [ctypes.POINTER(vp)]),
    "rb200_sync": (i32, []),
    "rb200_set_default_gemm_mode": (i32, [i32]),
    "rb200_get_default_gemm_mode": (i32, [ctypes.POINTER(i32)]),
    "rb200_launch_count": (i32, [ctypes.POINTER(i64), i32]),
    "rb200_last_gemm_path": (ctypes.c_char_p, []),
    "rb200_buffer_alloc": (i32, [i64, ctypes.POINTER(vp)]),
    "rb200_buffer_alloc_shareable": (i32, [i64, ctypes.POINTER(vp)]),
    "rb200_buffer_free": (i32, [vp]),
    "rb200_buffer_h2d": (i32, [vp, i64, vp, i64]),
    "rb200_buffer_d2h": (i32, [vp, vp, i64, i64]),
    "rb200_buffer_h2d_async": (i32, [vp, i64, vp, i64]),
    "rb200_buffer_d2h_async": (i32, [vp, vp, i64, i64]),
    "rb200_buffer_d2d": (i32, [vp, i64, vp, i64, i64]),
    "rb200_buffer_fill": (i32, [i32, vp, i64, i64, vp]),
    "rb200_host_alloc": (i32, [i64, ctypes.POINTER(vp)]),
    "rb200_host_free": (i32, [vp]),
    "rb200_event_create": (i32, [ctypes.POINTER(vp)]),
    "rb200_event_destroy": (i32, [vp]),
    "rb200_event_record": (i32, [vp]),
    "rb200_event_synchronize": (i32, [vp]),
    "rb200_event_query": (i32, [vp, ctypes.POINTER(i32)]),
    "rb200_event_wait": (i32, [vp]),
    "rb200_graph_begin": (i32, []),
    "rb200_graph_end": (i32, [ctypes.POINTER(vp)]),
    "rb200_graph_launch": (i32, [vp]),
    "rb200_graph_destroy": (i32, [vp]),
    "rb200_sgemm": (i32, [i32, i32, i32, i64, i64, i64, f32, vp, i64, i64, vp, i64, i64, f32, vp, i64, i64, i32]),
    "rb200_dgemm": (i32, [i32, i32, i32, i64, i64, i64, f64, vp, i64, i64, vp, i64, i64, f64, vp, i64, i64]),
    "rb200_sgemm_streamed": (i32, [i32, i32, i32, i64, i64, i64, f32, vp, vp, i64, i64, vp, vp, i64, i64,
                                   f32, vp, vp, i32, i64, i64, i32]),
    "rb200_gemm_strided_batched": (i32, [i32, i32, i32, i64, i64, i64, f64, vp, i64, i64, i64, vp, i64, i64, i64, f64,
                                         vp, i64, i64, i64, i64, i32]),
    "rb200_ipc_export": (i32, [vp, vp]),
    "rb200_ipc_open": (i32, [vp, ctypes.POINTER(vp)]),
    "rb200_ipc_close": (i32, [vp]),
    "rb200_sgemm_allgather": (i32, [i32, i32, i32, i64, i64, i64, f32, vp, i64, i64, vp, i64, i64, f32, vp, i64, i64, i32,
                                    i32, ctypes.POINTER(vp)]),
    "rb200_sgemm_allgather_mc": (i32, [i32, i32, i32, i64, i64, i64, f32, vp, i64, i64, vp, i64, i64, f32, vp, i64, i64, i32,
                                       vp]),
    "rb200_scal": (i32, [i32, i64, f64, vp, i64, i64]),
    "rb200_axpy": (i32, [i32, i64, f64, vp, i64, i64, vp, i64, i64]),
    "rb200_copy": (i32, [i32, i64, vp, i64, i64, vp, i64, i64]),
    "rb200_swap": (i32, [i32, i64, vp, i64, i64, vp, i64, i64]),
    "rb200_asum": (i32, [i32, i64, vp, i64, i64, ctypes.POINTER(f64)]),
    "rb200_nrm2": (i32, [i32, i64, vp, i64, i64, ctypes.POINTER(f64)]),
    "rb200_dot": (i32, [i32, i64, vp, i64, i64, vp, i64, i64, ctypes.POINTER(f64)]),
    "rb200_iamax": (i32, [i32, i64, vp, i64, i64, ctypes.POINTER(i64)]),
    "rb200_iamin": (i32, [i32, i64, vp, i64, i64, ctypes.POINTER(i64)]),
    "rb200_gemv": (i32, [i32, i32, i32, i64, i64, f64, vp, i64, i64, vp, i64, i64, f64, vp, i64, i64]),
    "rb200_omatcopy": (i32, [i32, i32, i32, i64, i64, f64, vp, i64, i64, vp, i64, i64]),
    "rb200_transpose": (i32, [i32, i32, ctypes.POINTER(i64), ctypes.POINTER(i64), vp, i64, vp, i64]),
    "rb200_gather": (i32, [i32, i32, i32, i32, i64, i64, i64, vp, i64, vp, i64, vp, i64]),
    "rb200_reduce_gather": (i32, [i32, i32, i32, i32, i64, i64, i64, vp, i64, vp, i64, vp, i64]),
    "rb200_add": (i32, [i32, i32, i64, i64, f64, vp, i64, i64, vp, i64, i64]),
    "rb200_multiply": (i32, [i32, i32, i64, i64, vp, i64, i64, vp, i64, i64]),
    "rb200_broadcast": (i32, [i32, i32, i32, i64, i64, vp, i64, i64, vp, i64, i64]),
    "rb200_unary": (i32, [i32, i32, i64, f64, vp, i64, i64, f64]),
    "rb200_compare": (i32, [i32, i32, i64, vp, i64, i64, vp, i64, i64]),
    "rb200_astype": (i32, [i64, i32, vp, i64, i64, i32, vp, i64, i64]),
    "rb200_sum": (i32, [i32, i64, vp, i64, i64, ctypes.POINTER(f64)]),
    "rb200_imax": (i32, [i32, i64, vp, i64, i64, ctypes.POINTER(i64)]),
    "rb200_imin": (i32, [i32, i64, vp, i64, i64, ctypes.POINTER(i64)]),
    "rb200_onehot": (i32, [i32, i32, i64, i64, f64, vp, i64, i64, vp, i64, i64]),
    "rb200_reduce_sum": (i32, [i32, i64, i64, i64, vp, i64, vp, i64]),
    "rb200_reduce_max": (i32, [i32, i64, i64, i64, vp, i64, vp, i64]),
    "rb200_reduce_argmax": (i32, [i32, i64, i64, i64, vp, i64, i32, vp, i64]),
    "rb200_reduce_argmax_partial": (i32, [i32, i64, i64, i64, vp, i64, i32, vp, i64]),
    "rb200_softmax": (i32, [i32, i64, i64, vp, i64, i64]),
    "rb200_conv2d_forward": (i32, [i32, vp, i64, i64, i64, i64, i64, i64, i64, i64, i64, i32, i64, i64,
                                   vp, i64, i64, vp, i64, vp, i64, i32]),
    "rb200_im2col2d": (i32, [i32, i32, vp, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64,
                             i32, i32, i64, i64, i32, vp, i64, i64]),
    "rb200_bandpart": (i32, [i32, i64, i64, i64, vp, i64, i64, i64]),
    "rb200_gatherb": (i32, [i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, vp, i64, vp, i64, vp, i64]),
    "rb200_gathernd": (i32, [i32, i32, i32, i32, i64, i64, i64, i32, vp, vp, i64, vp, vp, i64]),
    "rb200_imagecopy": (i32, [i32, i64, i64, i64, vp, i64, vp, i64, i32, i64, i64, i32, i32, i32]),
    "rb200_einsum": (i32, [i32, i32, i32, vp, vp, i64, vp, vp, i64, vp, vp, i64]),
    "rb200_topk": (i32, [i32, i32, i64, i64, vp, i64, i64, i32, vp, i64, vp, i64]),
    "rb200_cumsumb": (i32, [i32, i64, i64, i64, vp, i64, i32, i32, vp, i64]),
    "rb200_cumsum": (i32, [i32, i64, vp, i64, i64, i32, i32, vp, i64, i64]),
    "rb200_searchsorted": (i32, [i32, i32, i64, i64, vp, i64, i64, vp, i64, i64, i32, vp, i64, i64]),
    "rb200_masking": (i32, [i32, i64, i64, i64, i64, f64, i32, vp, i64, vp, i64]),
    "rb200_slice": (i32, [i32, i32, i32, i64, i64, i64, i64, vp, i64, i64, vp, i64, i64, i64, i64, i64, i64, i64, i64]),
    "rb200_repeat": (i32, [i32, i64, i64, i64, vp, i64, vp, i64]),
    "rb200_im2col3d": (i32, [i32, i32, vp, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i32, i32,
                             i64, i64, i64, i32, vp, i64, i64]),
}


class B200Error(RuntimeError):
    """A non-zero status from the C ABI (the PHP shim turns these into RuntimeException)."""

    def __init__(self, status: int, message: str):
        super().__init__("librindow_b200 status %d: %s" % (status, message))
        self.status = status


_lib = None


def load() -> ctypes.CDLL:
    """dlopen the library and bind every declared symbol.  Works without a GPU (no CUDA call)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(E_NOTINIT, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                       "(make -C rindow_math_matrix_b200/csrc); there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != OK:
        msg = load().rb200_last_error()
        raise B200Error(status, msg.decode("utf-8", "replace") if msg else "")


_initialised_device = None


def init(device: int | None = None) -> int:
    """rb200_init on `device` (default: LOCAL_RANK or 0).  Raises if there is no B200."""
    global _initialised_device
    L = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialised_device != device:
        check(L.rb200_init(device))
        _initialised_device = device
    return device


def is_initialised() -> bool:
    return _initialised_device is not None


def device_available() -> bool:
    """True when a CUDA device is visible (used by the service-level diagnosis only)."""
    try:
        n = i32(0)
        return load().rb200_device_count(ctypes.byref(n)) == OK and n.value > 0
    except (OSError, B200Error):
        return False


def sync() -> None:
    check(load().rb200_sync())


def launch_count(reset: bool = False) -> int:
    n = i64(0)
    check(load().rb200_launch_count(ctypes.byref(n), 1 if reset else 0))
    return n.value


def set_stream(cuda_stream_ptr: int | None) -> None:
    """Adopt a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); None restores the
    library-owned stream."""
    ptr = vp(-1 & 0xFFFFFFFFFFFFFFFF) if cuda_stream_ptr is None else vp(cuda_stream_ptr)
    check(load().rb200_set_stream(ptr))


class Graph:
    """`with Graph() as g: <driver calls>` records the calls' kernels (rb200_graph_begin / end); g.launch() replays them
    with one driver call.  Issue the sequence once before capturing (allocations and synchronising calls cannot be
    captured)."""

    def __init__(self):
        self._exec = None

    def __enter__(self):
        check(load().rb200_graph_begin())
        return self

    def __exit__(self, et, ev, tb):
        ex = vp()
        rc = load().rb200_graph_end(ctypes.byref(ex))
        if et is None:
            check(rc)
            self._exec = ex
        elif rc == 0 and ex:
            load().rb200_graph_destroy(ex)      # the body raised: end the capture, keep nothing
        return False

    def launch(self) -> None:
        check(load().rb200_graph_launch(self._exec))

    def close(self) -> None:
        if self._exec is not None:
            load().rb200_graph_destroy(self._exec)
            self._exec = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def last_gemm_path() -> str:
    return load().rb200_last_gemm_path().decode()


def device_info() -> dict:
    d, sm, maj, mnr, mem = i32(), i32(), i32(), i32(), i64()
    check(load().rb200_device_info(ctypes.byref(d), ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr),
                                   ctypes.byref(mem)))
    return {"device": d.value, "sm_count": sm.value, "cc": (maj.value, mnr.value), "total_mem": mem.value}
