import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi
from tools.probe_umma import pack_w
lib = _cabi.probe_lib()
rng = np.random.default_rng(1)
rows, row0 = 160, 16
for name, nt, maskfn in (("zero-masks 1 tap", 1, lambda t, r: 0), ("zero-masks 9 taps", 9, lambda t, r: 0),
                         ("mask rows<64 on taps>=1", 9, lambda t, r: int(t >= 1 and r < 64)),
                         ("mask row 5 on tap0 only", 2, lambda t, r: int(t == 0 and r == 5))):
    shifts = np.zeros(nt, np.int32)
    act = (rng.standard_normal((rows, 128)) * 0.5).astype(np.float16)
    W = (rng.standard_normal((nt, 128, 128)) * 0.1).astype(np.float16)
    masks = np.zeros((nt, 4), np.uint32)
    ref = np.zeros((128, 128))
    for t in range(nt):
        c = act[row0:row0 + 128].astype(np.float64) @ W[t].astype(np.float64)
        for r in range(128):
            if maskfn(t, r):
                masks[t, r // 32] |= np.uint32(1 << (r % 32))
                c[r] = 0
        ref += c
    out = np.full((128, 128), 7.0, np.float32)
    wp = pack_w(W)  # keep alive: .ctypes.data of a temporary dangles
    rc = lib.ao_umma_probe_masked(act.ctypes.data, rows, wp.ctypes.data, None, out.ctypes.data, row0, nt, shifts.ctypes.data, masks.ctypes.data)
    e = np.abs(out - ref).max(axis=1)
    print(f"{name:28s} rc={rc} max_err={e.max():.3e} bad_rows={int((e>1e-3).sum())} sample out[0,:3]={out[0,:3]} ref={ref[0,:3]}")
