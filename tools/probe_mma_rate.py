"""Raw tcgen05.mma issue rate of the tower's MMA flavours (csrc/probe/umma_probe.cu: ao_umma_rate)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi
L = _cabi.probe_lib()
L.ao_umma_rate.argtypes = [C.c_int, C.c_int, C.c_void_p]
L.ao_umma_rate.restype = C.c_int
names = {0: "cta_group::1 M128 unmasked", 1: "cta_group::1 M128 masked", 2: "cta_group::2 M256 unmasked", 3: "cta_group::2 M256 masked"}
for fl, iters in ((0, 2000), (3, 2000), (3, 20000), (3, 200000), (7, 200000), (16 + 3, 200000), (64 + 2, 100000), (256 + 3, 200000), (512 + 3, 200000), (1024 + 3, 200000), (2048 + 3, 200000), (4096 + 3, 200000), (2048 + 0, 200000)):
    out = (C.c_ulonglong * 3)()
    for rep in range(2):
        rc = L.ao_umma_rate(fl, iters, out)
    assert rc == 0, rc
    N = 64 if fl & 16 else 256 if fl & 64 else 128
    tf = 148 * out[1] * 2 * 128 * N * 16 / (out[2] * 1e-6) / 1e12
    print(f"flavour {fl:3d} ({names[fl & 3]}{', tap-shifted A' if fl & 4 else ''}{', smem store pressure' if fl & 8 else ''}"
          f"{', N=64' if fl & 16 else ''}{', N=256' if fl & 64 else ''}{', rotating accumulators' if fl & 32 else ''}"
          f"{', SW128 operands' if fl & 128 else ''}{', two issuing warps' if fl & 256 else ''}{', four issuing warps' if fl & 512 else ''}{', commit every 8' if fl & 1024 else ''}{', commit+wait every 16' if fl & 2048 else ''}{', commit+wait every 32' if fl & 4096 else ''}): {out[0] / out[1]:.1f} clock64 cycles per K16 MMA ({out[1]} MMAs), "
          f"kernel {out[2]} us -> clock64 rate {out[0] / out[2] / 1e3:.3f} GHz, {tf:.0f} TFLOP/s chip-wide")
