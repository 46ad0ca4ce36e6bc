// Stand-alone batched kernels for the rules / encodings of utils.py (unit-test and facade entry points):
//   check_win (utils.py:30-59), get_state_pt (utils.py:139-168), legal_actions (utils.py:22-27).
// One warp per board / ID.  The search kernels (tree.cu) use the same device functions from rules.cuh.
#include "engine.h"
#include "rules.cuh"

namespace ao {
namespace {

constexpr int kWarps = 4;

__global__ void __launch_bounds__(kWarps * 32)
check_win_kernel(const int8_t* __restrict__ boards, int n, int B, uint8_t* __restrict__ out) {
  __shared__ uint16_t scratch[kWarps][2][32];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarps + wib;
  if (i >= n) return;
  const int8_t* b = boards + (size_t)i * B * B;
  uint32_t rb = 0, rw = 0;
  if (lane < B)
    for (int x = 0; x < B; ++x) {
      const int v = b[lane * B + x];
      if (v > 0) rb |= 1u << x;
      if (v < 0) rw |= 1u << x;
    }
  int stones = __popc(rb) + __popc(rw);
#pragma unroll
  for (int o = 16; o; o >>= 1) stones += __shfl_xor_sync(kFull, stones, o);
  const int w = check_win_rows(rb, rw, B, stones, scratch[wib], lane);
  if (lane == 0) out[i] = (uint8_t)w;
}

// ID -> the five planes as dense float32 [5][B][B]; same plane construction as the NN request in tree.cu
__global__ void __launch_bounds__(kWarps * 32)
encode_state_kernel(const int16_t* __restrict__ ids, const int32_t* __restrict__ lens, int n, int B,
                    float* __restrict__ out) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarps + wib;
  if (i >= n) return;
  const int A = B * B;
  const int16_t* id = ids + (size_t)i * (A + 1);
  const int m = lens[i] - 1;
  uint32_t rb = 0, rw = 0;
  for (int k = 0; k < m; ++k) {
    const int a = id[1 + k];
    if (lane == a / B) {
      if ((k & 1) == 0) rb |= 1u << (a % B);
      else rw |= 1u << (a % B);
    }
  }
  const int l1 = m >= 1 ? id[m] : -1, l2 = m >= 2 ? id[m - 1] : -1;
  const bool black_to_move = (m & 1) == 0;
  const uint32_t own = black_to_move ? rb : rw, opp = black_to_move ? rw : rb;
  uint32_t opp_prev = opp, own_prev = own;
  if (l1 >= 0 && lane == l1 / B) opp_prev &= ~(1u << (l1 % B));
  if (l2 >= 0 && lane == l2 / B) own_prev &= ~(1u << (l2 % B));
  float* o = out + (size_t)i * 5 * A;
  if (lane < B) {
    for (int x = 0; x < B; ++x) {
      o[0 * A + lane * B + x] = (float)((own_prev >> x) & 1u);
      o[1 * A + lane * B + x] = (float)((opp_prev >> x) & 1u);
      o[2 * A + lane * B + x] = (float)((own >> x) & 1u);
      o[3 * A + lane * B + x] = (float)((opp >> x) & 1u);
      o[4 * A + lane * B + x] = black_to_move ? 1.f : 0.f;
    }
  }
}

__global__ void __launch_bounds__(kWarps * 32)
legal_actions_kernel(const int16_t* __restrict__ ids, const int32_t* __restrict__ lens, int n, int B,
                     int16_t* __restrict__ out) {
  __shared__ uint16_t occ[kWarps][32];
  __shared__ uint8_t order[kWarps][256];
  __shared__ int16_t table[kWarps][128];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarps + wib;
  if (i >= n) return;
  const int A = B * B;
  const int16_t* id = ids + (size_t)i * (A + 1);
  const int m = lens[i] - 1;
  uint32_t r = 0;
  for (int k = 0; k < m; ++k) {
    const int a = id[1 + k];
    if (lane == a / B) r |= 1u << (a % B);
  }
  occ[wib][lane] = (uint16_t)r;
  __syncwarp();
  const int L = legal_order(occ[wib], B, A, order[wib], table[wib], lane);
  for (int k = lane; k < A; k += 32) out[(size_t)i * A + k] = k < L ? (int16_t)order[wib][k] : (int16_t)-1;
}

}  // namespace

cudaError_t launch_check_win(const int8_t* boards, int n, int B, uint8_t* out, cudaStream_t s) {
  check_win_kernel<<<(n + kWarps - 1) / kWarps, kWarps * 32, 0, s>>>(boards, n, B, out);
  return cudaGetLastError();
}
cudaError_t launch_encode_state(const int16_t* ids, const int32_t* lens, int n, int B, float* out, cudaStream_t s) {
  encode_state_kernel<<<(n + kWarps - 1) / kWarps, kWarps * 32, 0, s>>>(ids, lens, n, B, out);
  return cudaGetLastError();
}
cudaError_t launch_legal_actions(const int16_t* ids, const int32_t* lens, int n, int B, int16_t* out, cudaStream_t s) {
  legal_actions_kernel<<<(n + kWarps - 1) / kWarps, kWarps * 32, 0, s>>>(ids, lens, n, B, out);
  return cudaGetLastError();
}

}  // namespace ao
