"""tcgen05.mma rate with FOUR issuing warps at the shapes a one-game tower can use (csrc/probe/umma_probe.cu, zero
operands, all 148 SMs): what bounds an MMA once the issue rate is out of the way - the math (floor = M*N/256 cycles per
K16 instruction and SM) or fetching the operands from shared memory (A: rows x 32 B, B: N x 32 B per instruction)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alpha_omok_b200 import _cabi
L = _cabi.probe_lib()
L.ao_umma_rate.argtypes = [C.c_int, C.c_int, C.c_void_p]
L.ao_umma_rate.restype = C.c_int
F4, MASKSHIFT = 512, 5
cases = [("cta_group::1 M128 N32 ", F4 + 8192 + MASKSHIFT, 128, 32, 1),
         ("cta_group::1 M128 N64 ", F4 + 16 + MASKSHIFT, 128, 64, 1),
         ("cta_group::1 M128 N128", F4 + MASKSHIFT, 128, 128, 1),
         ("cta_group::2 M128 N64  (64 rows per CTA)", F4 + 2 + 16384 + 16 + MASKSHIFT, 64, 64, 2),
         ("cta_group::2 M128 N128 (64 rows per CTA)", F4 + 2 + 16384 + MASKSHIFT, 64, 128, 2),
         ("cta_group::2 M256 N64 ", F4 + 2 + 16 + MASKSHIFT, 128, 64, 2),
         ("cta_group::2 M256 N128", F4 + 2 + MASKSHIFT, 128, 128, 2)]
if len(sys.argv) > 1 and sys.argv[1] == "sw128":   # the same shapes with 128-byte-swizzled K-major operands (timing only)
    cases = [(n + " SW128", f + 128, r, N, c) for n, f, r, N, c in cases]
for name, fl, rows, N, cg in cases:
    out = (C.c_ulonglong * 3)()
    for rep in range(2):
        rc = L.ao_umma_rate(fl, 100000, out)
    assert rc == 0, (name, rc)
    cyc = out[0] / out[1]
    a_bytes, b_bytes = rows * 32, (N // cg) * 32
    floor = max(rows, 128 // cg if cg == 2 and rows == 64 else rows) * N / 256.0
    print(f"{name}: {cyc:6.1f} cycles per K16 MMA with 4 issuers (masked, tap-shifted A); per SM: A {a_bytes} B + B {b_bytes} B from shared memory "
          f"= {(a_bytes + b_bytes) / cyc:5.1f} B/cycle; math floor {rows * N / 256.0:.0f} cycles")
