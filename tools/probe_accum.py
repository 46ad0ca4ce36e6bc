"""Does the tensor core's fp32 accumulation round to nearest or truncate? (all-positive long accumulation)"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _build
from tools.probe_umma import pack_w
lib = ctypes.CDLL(_build.build(probe=True))
lib.ao_umma_probe.restype = ctypes.c_int
lib.ao_umma_probe.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
rng = np.random.default_rng(0)
for nt in (1, 8, 64):
    act = rng.uniform(0, 1, (128, 128)).astype(np.float16)
    W = rng.uniform(0, 0.1, (nt, 128, 128)).astype(np.float16)
    ref = np.zeros((128, 128))
    seq = np.zeros((128, 128), np.float32)
    for t in range(nt):
        ref += act.astype(np.float64) @ W[t].astype(np.float64)
        for j in range(8):  # fp32 round-to-nearest accumulation of exact K=16 partial dot products
            part = act[:, 16 * j:16 * j + 16].astype(np.float64) @ W[t][16 * j:16 * j + 16].astype(np.float64)
            seq = (seq.astype(np.float64) + part).astype(np.float32)
    out = np.zeros((128, 128), np.float32)
    sh = np.zeros(nt, np.int32)
    wp = pack_w(W)
    rc = lib.ao_umma_probe(act.ctypes.data, 128, wp.ctypes.data, None, out.ctypes.data, 0, nt, sh.ctypes.data)
    ulp = np.spacing(ref.astype(np.float32)).mean()
    print(f"ntaps {nt:3d} rc {rc} mean(ref) {ref.mean():8.2f} ulp {ulp:.2e} | device-ref: mean {np.mean(out-ref)/ulp:+8.2f} ulp, max|.| {np.abs(out-ref).max()/ulp:7.2f} ulp"
          f" | fp32-RN-seq - ref: mean {np.mean(seq-ref)/ulp:+6.2f} ulp, max {np.abs(seq-ref).max()/ulp:6.2f} ulp")
