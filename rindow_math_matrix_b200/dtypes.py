"""dtype constants of the driver boundary.

Names follow Interop\\Polite\\Math\\Matrix\\NDArray (used as NDArray::float32 etc. all over the
reference, e.g. src/Drivers/MatlibPHP/PhpBuffer.php:24-42); the numeric values are the C ABI's
own enum (include/rindow_b200.h) -- the PHP shim maps the polite-math constants to them in one
table (php/MatlibCuda/CudaBuffer.php).
"""
import numpy as np

bool_ = 0
int8 = 1
int16 = 2
int32 = 3
int64 = 4
uint8 = 5
uint16 = 6
uint32 = 7
uint64 = 8
float32 = 12
float64 = 13
complex64 = 16          # storage dtypes (PhpBuffer.php:38-41): interleaved real / imag, no device arithmetic
complex128 = 17

NUMPY = {
    bool_: np.uint8, int8: np.int8, int16: np.int16, int32: np.int32, int64: np.int64,
    uint8: np.uint8, uint16: np.uint16, uint32: np.uint32, uint64: np.uint64,
    float32: np.float32, float64: np.float64, complex64: np.complex64, complex128: np.complex128,
}
# PhpBuffer.php:44-62 $valueSize
VALUE_SIZE = {k: np.dtype(v).itemsize for k, v in NUMPY.items()}
INT_TYPES = (int8, int16, int32, int64, uint8, uint16, uint32, uint64)
FLOAT_TYPES = (float32, float64)
COMPLEX_TYPES = (complex64, complex128)
NAMES = {bool_: "bool", int8: "int8", int16: "int16", int32: "int32", int64: "int64", uint8: "uint8",
         uint16: "uint16", uint32: "uint32", uint64: "uint64", float32: "float32", float64: "float64", complex64: "complex64",
         complex128: "complex128"}


def from_numpy(dt) -> int:
    dt = np.dtype(dt)
    if dt == np.bool_:
        return bool_
    for code, npdt in NUMPY.items():
        if code != bool_ and np.dtype(npdt) == dt:
            return code
    raise ValueError("Invalid data type")
