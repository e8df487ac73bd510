import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import LinearAlgebra
svc = rb.MatlibCuda()
la = LinearAlgebra(svc)
cases = [(3, 17, 23, 3, 3, 3, 64, (2, 1), True, (1, 2), True), (1, 30, 31, 8, 3, 3, 32, (1, 1), False, (1, 1), False),
         (3, 17, 23, 3, 3, 3, 64, (1, 1), True, (1, 1), False)]
for case in cases:
    b, h, w, c, kh, kw, f, st, pad, dil, use_bias = case
    rng = np.random.default_rng(4001)
    img = rng.uniform(0, 1, (b, h, w, c)).astype(np.float32)
    W = (rng.standard_normal((kh, kw, c, f)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(f).astype(np.float32) if use_bias else None
    out = la.conv2d(la.array(img), la.array(W), None if bias is None else la.array(bias), strides=st, padding=pad, dilation_rate=dil, fused=True)
    path = rb._capi.last_gemm_path()
    out2 = la.conv2d(la.array(img), la.array(W), None if bias is None else la.array(bias), strides=st, padding=pad, dilation_rate=dil)
    a, r = out.to_numpy(), out2.to_numpy()
    bad = np.abs(a - r) > 1e-3
    print(case, path, a.shape, "bad", bad.sum(), "of", bad.size)
    if bad.any():
        idx = np.argwhere(bad)
        print(" first bad", idx[:5].tolist(), "last bad", idx[-3:].tolist())
        print(" bad per oy:", bad.sum(axis=(0, 2, 3)).tolist())
        print(" bad per ox:", bad.sum(axis=(0, 1, 3)).tolist())
        print(" bad per f :", bad.sum(axis=(0, 1, 2)).tolist())
