"""Hardware facts the tower kernel is built on, checked through the C ABI probe (csrc/probe/umma_probe.cu):
shifted start addresses of the K-major no-swizzle UMMA operand, TMEM-preloaded accumulation, and per-MMA
disable-output-lane masks (board edges without padding)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pack_w(W):
    t, ci, co = W.shape
    return np.ascontiguousarray(W.reshape(t, ci // 8, 8, co).transpose(0, 1, 3, 2))


def _run(lib, act, W, init, row0, shifts, masks=None):
    wp = _pack_w(W)  # keep alive while the C call reads it
    out = np.zeros((128, 128), np.float32)
    sh = np.asarray(shifts, np.int32)
    rc = lib.ao_umma_probe_masked(act.ctypes.data, act.shape[0], wp.ctypes.data,
                                  None if init is None else init.ctypes.data, out.ctypes.data, row0, len(shifts),
                                  sh.ctypes.data, None if masks is None else masks.ctypes.data)
    assert rc == 0
    return out


@pytest.mark.parametrize("stride,use_init", [(10, False), (10, True), (16, True), (9, True)])
def test_shifted_operand_conv(stride, use_init):
    from alpha_omok_b200 import _cabi
    lib = _cabi.probe_lib()
    rng = np.random.default_rng(stride)
    rows, row0 = 300, 24
    shifts = [dy * stride + dx for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    act = (rng.standard_normal((rows, 128)) * 0.5).astype(np.float16)
    W = (rng.standard_normal((9, 128, 128)) * 0.1).astype(np.float16)
    init = rng.standard_normal((128, 128)).astype(np.float32) if use_init else None
    ref = np.zeros((128, 128)) if init is None else init.astype(np.float64)
    for t, s in enumerate(shifts):
        ref = ref + act[row0 + s: row0 + s + 128].astype(np.float64) @ W[t].astype(np.float64)
    out = _run(lib, act, W, init, row0, shifts)
    assert np.abs(out - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())


def test_disable_output_lane_masks_replace_padding():
    from alpha_omok_b200 import _cabi
    lib = _cabi.probe_lib()
    rng = np.random.default_rng(1)
    B, rows, row0 = 9, 160, 16
    taps = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    taps = [taps[4]] + taps[:4] + taps[5:]  # centre tap first: no disabled rows on the accumulate=0 MMA
    shifts = [dy * B + dx for dy, dx in taps]
    act = (rng.standard_normal((rows, 128)) * 0.5).astype(np.float16)
    W = (rng.standard_normal((9, 128, 128)) * 0.1).astype(np.float16)
    masks = np.zeros((9, 4), np.uint32)
    ref = np.zeros((128, 128))
    for t, (dy, dx) in enumerate(taps):
        c = act[row0 + shifts[t]: row0 + shifts[t] + 128].astype(np.float64) @ W[t].astype(np.float64)
        for r in range(128):
            y, x = divmod(r % 81, 9)
            if not (0 <= y + dy < 9 and 0 <= x + dx < 9):
                masks[t, r // 32] |= np.uint32(1 << (r % 32))
                c[r] = 0.0
        ref += c
    out = _run(lib, act, W, None, row0, shifts, masks)
    assert np.abs(out - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())
