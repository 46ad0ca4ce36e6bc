// Device-side rules of Omok on row bit-masks (one 16-bit mask per board row and colour, bit x = column x), the
// Philox decision stream, and the numpy / CPython arithmetic restatements the search needs for bit-exact parity.
// Reference semantics: utils.py:22-59,139-179 ; numpy pairwise sum ; CPython 3.12 setobject.c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ao {

constexpr unsigned kFull = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// ------------------------------------------------------------------------------------------------ five in a row
// Lane y (< B) passes its row masks; lanes >= B must pass 0.  Returns, warp-uniformly, bit0 = black has >= 5 in a
// row somewhere, bit1 = white has.
__device__ __forceinline__ uint32_t five_mask(uint32_t r) {
  const uint32_t r1 = __shfl_down_sync(kFull, r, 1), r2 = __shfl_down_sync(kFull, r, 2);
  const uint32_t r3 = __shfl_down_sync(kFull, r, 3), r4 = __shfl_down_sync(kFull, r, 4);
  const uint32_t h = r & (r >> 1) & (r >> 2) & (r >> 3) & (r >> 4);
  const uint32_t v = r & r1 & r2 & r3 & r4;
  const uint32_t d = r & (r1 >> 1) & (r2 >> 2) & (r3 >> 3) & (r4 >> 4);
  const uint32_t a = r & (r1 << 1) & (r2 << 2) & (r3 << 3) & (r4 << 4);
  return h | v | d | a;
}

// utils.check_win restated exactly (window scan order + black-before-white inside a window) for boards on which
// BOTH colours have a five; `rows` = shared memory [2][32] row masks (black, white). Warp-collective.
__device__ inline int win_scan_windows(const uint16_t (*rows)[32], int B, int lane) {
  const int nw = B - 4;
  int first = 1 << 30;
  for (int w = lane; w < nw * nw; w += 32) {
    const int r = w / nw, c = w % nw;
    int hit = 0;
#pragma unroll
    for (int col = 0; col < 2; ++col) {
      uint32_t m[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) m[i] = (rows[col][r + i] >> c) & 31u;
      const bool line = (m[0] == 31u) | (m[1] == 31u) | (m[2] == 31u) | (m[3] == 31u) | (m[4] == 31u) |
                        ((m[0] & m[1] & m[2] & m[3] & m[4]) != 0u) |
                        (((m[0]) & (m[1] >> 1) & (m[2] >> 2) & (m[3] >> 3) & (m[4] >> 4)) & 1u) |
                        (((m[0] >> 4) & (m[1] >> 3) & (m[2] >> 2) & (m[3] >> 1) & (m[4])) & 1u);
      if (line && hit == 0) hit = col + 1;
    }
    if (hit) {
      first = (w << 2) | hit;
      break;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) first = min(first, __shfl_xor_sync(kFull, first, o));
  return first == (1 << 30) ? 0 : (first & 3);
}

// 0 playing, 1 black, 2 white, 3 draw.  rb/rw: this lane's row masks (0 for lanes >= B); n_stones: stones on board.
__device__ inline int check_win_rows(uint32_t rb, uint32_t rw, int B, int n_stones, uint16_t (*scratch)[32], int lane) {
  const bool fb = __ballot_sync(kFull, five_mask(rb) != 0u) != 0u;
  const bool fw = __ballot_sync(kFull, five_mask(rw) != 0u) != 0u;
  if (fb && fw) {  // impossible in legal play; arbitrary boards only
    scratch[0][lane] = (uint16_t)rb;
    scratch[1][lane] = (uint16_t)rw;
    __syncwarp();
    const int r = win_scan_windows(scratch, B, lane);
    __syncwarp();
    return r;
  }
  if (fb) return 1;
  if (fw) return 2;
  return n_stones == B * B ? 3 : 0;
}

// ------------------------------------------------------------------------------------------------ numpy pairwise sum
// float64 sum of a[0..n) in numpy's order (8 strided accumulators per <=128 block, pairwise above). Warp-collective,
// result on all lanes. a in shared memory.
__device__ inline double np_block_sum(const double* a, int n, int lane) {
  double res = 0.0;
  if (n < 8) {
    if (lane == 0)
      for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return __shfl_sync(kFull, res, 0);
  }
  const int nb = n - (n & 7);
  double r = 0.0;
  if (lane < 8) {
    r = a[lane];
    for (int i = 8 + lane; i < nb; i += 8) r = __dadd_rn(r, a[i]);
  }
  const double s = __dadd_rn(r, __shfl_down_sync(kFull, r, 1));  // lanes 0,2,4,6
  const double t = __dadd_rn(s, __shfl_down_sync(kFull, s, 2));  // lanes 0,4
  res = __dadd_rn(t, __shfl_down_sync(kFull, t, 4));             // lane 0
  if (lane == 0)
    for (int i = nb; i < n; ++i) res = __dadd_rn(res, a[i]);
  return __shfl_sync(kFull, res, 0);
}
__device__ inline double np_pairwise_sum(const double* a, int n, int lane) {
  if (n <= 128) return np_block_sum(a, n, lane);
  int n2 = n / 2;
  n2 -= n2 % 8;
  const double l = (n2 <= 128) ? np_block_sum(a, n2, lane) : 0.0;  // A <= 256: one split level is enough
  const double r = np_block_sum(a + n2, n - n2, lane);
  return __dadd_rn(l, r);
}

// ------------------------------------------------------------------------------------------------ CPython set order
__device__ inline void cpy_set_insert_clean(int16_t* table, int mask, int key) {
  unsigned perturb = (unsigned)key;
  int i = key & mask;
  while (true) {
    int probes = (i + 9 <= mask) ? 9 : 0;
    int j = i;
    while (true) {
      if (table[j] < 0) {
        table[j] = (int16_t)key;
        return;
      }
      if (probes == 0) break;
      --probes;
      ++j;
    }
    perturb >>= 5;
    i = (i * 5 + 1 + (int)perturb) & mask;
  }
}

// Child order = iteration order of set(range(A)) - set(stones) (utils.py:22-27). occ_rows: shared [32] occupancy
// row masks. Writes order[0..L) (shared) and returns L. Warp-collective; `table` = shared int16[128] scratch.
__device__ inline int legal_order(const uint16_t* occ_rows, int B, int A, uint8_t* order, int16_t* table, int lane) {
  // ascending enumeration (also the answer whenever the final hash table is larger than the largest key)
  int L = 0, max_key = -1;
  for (int base = 0; base < A; base += 32) {
    const int a = base + lane;
    bool legal = false;
    if (a < A) legal = ((occ_rows[a / B] >> (a % B)) & 1u) == 0u;
    const unsigned bal = __ballot_sync(kFull, legal);
    if (legal) order[L + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)a;
    if (bal) max_key = base + 31 - __clz(bal);
    L += __popc(bal);
  }
  __syncwarp();
  const int stones = A - L;
  if ((A >> 2) > stones) return L;  // set_copy_and_difference path
  const int final_mask = L <= 4 ? 7 : (L <= 18 ? 31 : (L <= 76 ? 127 : 511));
  if (final_mask >= max_key) return L;
  if (lane == 0) {  // serial emulation of the insert / resize sequence (L <= 76 here, tables of 8, 32 or 128 slots)
    // order[0..L) holds the ascending source keys; order[128..256) parks the old keys during a resize (L < 128).
    uint8_t* park = order + 128;
    int mask = 7, fill = 0;
    for (int i = 0; i < 8; ++i) table[i] = -1;
    for (int k = 0; k < L; ++k) {
      cpy_set_insert_clean(table, mask, (int)order[k]);
      ++fill;
      if (fill * 5 >= mask * 3) {  // set_table_resize(so, used * 4): re-insert in old slot order
        int newsize = 8;
        while (newsize <= fill * 4) newsize <<= 1;
        int cnt = 0;
        for (int sidx = 0; sidx <= mask; ++sidx)
          if (table[sidx] >= 0) park[cnt++] = (uint8_t)table[sidx];
        for (int q = 0; q < newsize; ++q) table[q] = -1;
        mask = newsize - 1;
        for (int q = 0; q < cnt; ++q) cpy_set_insert_clean(table, mask, (int)park[q]);
      }
    }
    int o = 0;
    for (int sidx = 0; sidx <= mask; ++sidx)
      if (table[sidx] >= 0) order[o++] = (uint8_t)table[sidx];
  }
  __syncwarp();
  return L;
}

}  // namespace ao
