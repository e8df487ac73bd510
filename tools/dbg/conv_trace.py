import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import CudaBuffer, dtypes
svc = rb.MatlibCuda(); math = svc.math()
b, h, w, ch, f = 64, 224, 224, 3, 64
rng = np.random.default_rng(0)
im = CudaBuffer(b*h*w*ch, dtypes.float32, zero=False).from_numpy(rng.uniform(-1, 1, b*h*w*ch).astype(np.float32))
Wk = CudaBuffer(27*f, dtypes.float32, zero=False).from_numpy(rng.uniform(-1, 1, 27*f).astype(np.float32))
out = CudaBuffer(b*222*222*f, dtypes.float32, zero=False)
for _ in range(3):
    math.conv2d_forward(im, 0, b, h, w, ch, 3, 3, 1, 1, False, 1, 1, Wk, 0, f, None, 0, out, 0)
rb._capi.sync()
