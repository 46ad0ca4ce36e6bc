"""Hardware probe: tcgen05.mma disable-output-lane masks = per-tap row masking (board-edge handling without padding)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi
from tools.probe_umma import pack_w
lib = _cabi.probe_lib()
rng = np.random.default_rng(1)
B = 9
rows, row0 = 16 + 128 + 16, 16
taps = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
taps = [taps[4]] + taps[:4] + taps[5:]          # centre tap first (no disabled rows) so accumulate=0 is safe
shifts = np.asarray([dy * B + dx for dy, dx in taps], np.int32)
act = (rng.standard_normal((rows, 128)) * 0.5).astype(np.float16)
W = (rng.standard_normal((9, 128, 128)) * 0.1).astype(np.float16)
masks = np.zeros((9, 4), np.uint32)
ref = np.zeros((128, 128))
for t, (dy, dx) in enumerate(taps):
    contrib = act[row0 + shifts[t]: row0 + shifts[t] + 128].astype(np.float64) @ W[t].astype(np.float64)
    for r in range(128):
        g, pos = divmod(r, 81)
        y, x = divmod(pos, 9)
        off_board = not (0 <= y + dy < 9 and 0 <= x + dx < 9)
        if off_board:
            masks[t, r // 32] |= np.uint32(1 << (r % 32))
            contrib[r] = 0.0
    ref += contrib
for use_init in (False, True):
    init = rng.standard_normal((128, 128)).astype(np.float32) if use_init else None
    out = np.zeros((128, 128), np.float32)
    wp = pack_w(W)  # keep alive: .ctypes.data of a temporary dangles
    rc = lib.ao_umma_probe_masked(act.ctypes.data, rows, wp.ctypes.data, init.ctypes.data if use_init else None,
                                  out.ctypes.data, row0, 9, shifts.ctypes.data, masks.ctypes.data)
    r = ref + (init if use_init else 0)
    err = np.abs(out - r).max()
    print(f"masked conv (stride 9, no padding) init={use_init}: rc={rc} max_abs_err={err:.3e} scale={np.abs(r).max():.2f}", "OK" if rc == 0 and err < 1e-3 else "MISMATCH")
    bad = np.flatnonzero(np.abs(out - r).max(axis=1) > 1e-3)
    if len(bad):
        print("  bad rows:", bad[:40], "count", len(bad))
        rr = bad[0]
        print("  row", rr, "out", out[rr, :4], "ref", r[rr, :4], "disabled-in-taps", [t for t in range(9) if masks[t, rr // 32] >> (rr % 32) & 1])
