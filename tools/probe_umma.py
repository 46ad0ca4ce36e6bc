"""Hardware probe for the shifted no-swizzle UMMA operand trick (see csrc/probe/umma_probe.cu).

Run on the GPU box:  python tools/probe_umma.py   (prints one line per case, exit 1 on mismatch)
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _build  # noqa: E402


def pack_w(W):
    """W[t][ci][co] -> [t][ci//8][co][ci%8] (K-major no-swizzle canonical B operand)."""
    t, ci, co = W.shape
    return np.ascontiguousarray(W.reshape(t, ci // 8, 8, co).transpose(0, 1, 3, 2))


def run_case(lib, name, rows, row0, shifts, use_init, rng):
    nt = len(shifts)
    act = (rng.standard_normal((rows, 128)) * 0.5).astype(np.float16)
    W = (rng.standard_normal((nt, 128, 128)) * 0.1).astype(np.float16)
    init = rng.standard_normal((128, 128)).astype(np.float32) if use_init else None
    ref = np.zeros((128, 128), np.float64) if init is None else init.astype(np.float64)
    for t, s in enumerate(shifts):
        ref = ref + act[row0 + s: row0 + s + 128].astype(np.float64) @ W[t].astype(np.float64)
    out = np.zeros((128, 128), np.float32)
    wp = pack_w(W)
    sh = np.asarray(shifts, np.int32)
    rc = lib.ao_umma_probe(
        act.ctypes.data_as(ctypes.c_void_p), rows, wp.ctypes.data_as(ctypes.c_void_p),
        init.ctypes.data_as(ctypes.c_void_p) if init is not None else None,
        out.ctypes.data_as(ctypes.c_void_p), row0, nt, sh.ctypes.data_as(ctypes.c_void_p))
    err = float(np.abs(out - ref).max())
    scale = float(np.abs(ref).max())
    ok = rc == 0 and err < 2e-3 * max(scale, 1.0)
    print(f"{name:28s} rc={rc} max_abs_err={err:.3e} ref_scale={scale:.2f} {'OK' if ok else 'MISMATCH'}", flush=True)
    return ok


def main():
    lib = ctypes.CDLL(_build.build(probe=True))
    lib.ao_umma_probe.restype = ctypes.c_int
    lib.ao_umma_probe.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    rng = np.random.default_rng(0)
    ok = True
    ok &= run_case(lib, "gemm_aligned", 128, 0, [0], False, rng)
    ok &= run_case(lib, "gemm_aligned_init", 128, 0, [0], True, rng)
    ok &= run_case(lib, "shift_plus1", 160, 0, [1], False, rng)
    ok &= run_case(lib, "shift_row0_16_minus11", 160, 16, [-11], False, rng)
    ok &= run_case(lib, "shift_plus8", 160, 0, [8], False, rng)
    taps10 = [dy * 10 + dx for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    taps16 = [dy * 16 + dx for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    ok &= run_case(lib, "conv9_S10", 290, 16, taps10, False, rng)
    ok &= run_case(lib, "conv9_S10_init", 290, 16, taps10, True, rng)
    ok &= run_case(lib, "conv9_S16_init", 300, 24, taps16, True, rng)
    print("PROBE", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
