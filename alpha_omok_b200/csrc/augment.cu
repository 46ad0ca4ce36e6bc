// Replay records -> training samples on the device (the consumer of the replay all-gather):
// for every ply of every finished game, the network input planes (utils.get_state_pt, utils.py:139-168), the target
// pi (visit / visit.sum() before TAU_THRES, the played move's one-hot after; main.py:150-166), z (main.py:201-227) and
// the 8 dihedral copies of utils.augment_dataset (utils.py:226-239: for k in 0..3: rot90(k), then its left-right flip),
// written as float32 tensors the PyTorch training step consumes (main.py:263-296 casts to float32 anyway).
// Pure byte shuffling: HBM-bound, 1948 B written per augmented 9x9 sample, nothing re-read.
#include <stdio.h>

#include <string>

#include "engine.h"
#include "rules.cuh"

namespace ao {
namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ size_t rec_voff(int A) { return (4 + (size_t)A * 2 + 3) & ~(size_t)3; }

// offsets[g] = number of samples (plies) of the games before g; offsets[n] = total. One block.
__global__ void sample_offsets_kernel(const uint8_t* __restrict__ slab, size_t rec_bytes, int n, long long* __restrict__ offsets) {
  __shared__ long long s_part[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (n + nt - 1) / nt;
  const int lo = tid * per, hi = min(n, lo + per);
  long long acc = 0;
  for (int g = lo; g < hi; ++g) acc += *reinterpret_cast<const int16_t*>(slab + (size_t)g * rec_bytes);
  s_part[tid] = acc;
  __syncthreads();
  if (tid == 0) {
    long long run = 0;
    for (int i = 0; i < nt; ++i) {
      const long long v = s_part[i];
      s_part[i] = run;
      run += v;
    }
    offsets[n] = run;
  }
  __syncthreads();
  long long run = s_part[tid];
  for (int g = lo; g < hi; ++g) {
    offsets[g] = run;
    run += *reinterpret_cast<const int16_t*>(slab + (size_t)g * rec_bytes);
  }
}

// source cell of output cell (y, x) under augmentation a = 2*k + flip  (np.rot90(m, k) then np.fliplr)
__device__ __forceinline__ int aug_source(int a, int y, int x, int B) {
  const int k = a >> 1;
  if (a & 1) x = B - 1 - x;  // the flip is applied AFTER the rotation: un-flip first
  int sy, sx;
  switch (k) {
    case 0: sy = y; sx = x; break;
    case 1: sy = x; sx = B - 1 - y; break;          // rot90: out[y][x] = in[x][B-1-y]
    case 2: sy = B - 1 - y; sx = B - 1 - x; break;
    default: sy = B - 1 - x; sx = y; break;         // rot270: out[y][x] = in[B-1-x][y]
  }
  return sy * B + sx;
}

// one block per game, one warp per ply (strided)
__global__ void __launch_bounds__(kWarps * 32)
augment_kernel(const uint8_t* __restrict__ slab, size_t rec_bytes, int n, int B, int tau_thres,
               const long long* __restrict__ offsets, long long capacity, float* __restrict__ states,
               float* __restrict__ pis, float* __restrict__ zs) {
  __shared__ float s_pi[kWarps][kMaxA + 3];
  __shared__ uint16_t s_rows[kWarps][4][32];
  const int g = blockIdx.x;
  if (g >= n) return;
  const int A = B * B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* rec = slab + (size_t)g * rec_bytes;
  const int n_moves = *reinterpret_cast<const int16_t*>(rec);
  const int winner = rec[2];
  const int16_t* mv = reinterpret_cast<const int16_t*>(rec + 4);
  const uint32_t* vis = reinterpret_cast<const uint32_t*>(rec + rec_voff(A));
  const float z_black = winner == 1 ? 1.f : (winner == 2 ? -1.f : 0.f);
  for (int t = warp; t < n_moves; t += kWarps) {
    const long long base = offsets[g] + t;
    if ((base + 1) * 8 > capacity) continue;
    // ---- position after t plies as row masks (lane y = board row y)
    uint32_t rb = 0, rw = 0;
    for (int i = 0; i < t; ++i) {
      const int a = mv[i];
      if (lane == a / B) {
        if ((i & 1) == 0) rb |= 1u << (a % B);
        else rw |= 1u << (a % B);
      }
    }
    const int l1 = t >= 1 ? mv[t - 1] : -1, l2 = t >= 2 ? mv[t - 2] : -1;
    const bool black_to_move = (t & 1) == 0;
    const uint32_t own = black_to_move ? rb : rw, opp = black_to_move ? rw : rb;
    uint32_t opp_prev = opp, own_prev = own;
    if (l1 >= 0 && lane == l1 / B) opp_prev &= ~(1u << (l1 % B));
    if (l2 >= 0 && lane == l2 / B) own_prev &= ~(1u << (l2 % B));
    s_rows[warp][0][lane] = (uint16_t)own_prev;
    s_rows[warp][1][lane] = (uint16_t)opp_prev;
    s_rows[warp][2][lane] = (uint16_t)own;
    s_rows[warp][3][lane] = (uint16_t)opp;
    // ---- pi
    const uint32_t* v = vis + (size_t)t * A;
    if (t < tau_thres) {
      double total = 0.0;
      for (int a = lane; a < A; a += 32) total += (double)v[a];
#pragma unroll
      for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
      for (int a = lane; a < A; a += 32) s_pi[warp][a] = (float)__ddiv_rn((double)v[a], total);
    } else {
      const int played = mv[t];
      for (int a = lane; a < A; a += 32) s_pi[warp][a] = a == played ? 1.f : 0.f;
    }
    __syncwarp();
    const float colour = black_to_move ? 1.f : 0.f;
    const float z = black_to_move ? z_black : -z_black;
    // ---- the 8 dihedral copies
    for (int a8 = 0; a8 < 8; ++a8) {
      const long long s_idx = base * 8 + a8;
      float* so = states + (size_t)s_idx * 5 * A;
      float* po = pis + (size_t)s_idx * A;
      for (int c = lane; c < A; c += 32) {
        const int src = aug_source(a8, c / B, c % B, B);
        const int sy = src / B, sx = src % B;
        so[0 * A + c] = (float)((s_rows[warp][0][sy] >> sx) & 1u);
        so[1 * A + c] = (float)((s_rows[warp][1][sy] >> sx) & 1u);
        so[2 * A + c] = (float)((s_rows[warp][2][sy] >> sx) & 1u);
        so[3 * A + c] = (float)((s_rows[warp][3][sy] >> sx) & 1u);
        so[4 * A + c] = colour;
        po[c] = s_pi[warp][src];
      }
      if (lane == 0) zs[s_idx] = z;
    }
    __syncwarp();
  }
}

thread_local std::string g_aug_err;

}  // namespace
}  // namespace ao

// records slab (device) -> augmented float32 training tensors (device). All pointers are DEVICE pointers except
// n_samples_out (host). Returns 0 / negative; *n_samples_out = 8 * total plies (samples actually needed).
extern "C" int ao_augment_records_dev(const void* slab_dev, int n_games, int board_size, int tau_thres, float* states_dev,
                                      float* pi_dev, float* z_dev, long long capacity_samples, long long* n_samples_out,
                                      void* stream) {
  if (!slab_dev || n_games < 1 || board_size < 5 || board_size > ao::kMaxB) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int A = board_size * board_size;
  const size_t rec_bytes = ((4 + (size_t)A * 2 + 3) & ~(size_t)3) + (size_t)A * A * 4;
  long long* d_off = nullptr;
  if (cudaMalloc(&d_off, (size_t)(n_games + 1) * sizeof(long long)) != cudaSuccess) return -2;
  ao::sample_offsets_kernel<<<1, 1024, 0, s>>>(reinterpret_cast<const uint8_t*>(slab_dev), rec_bytes, n_games, d_off);
  long long total = 0;
  cudaError_t e = cudaMemcpyAsync(&total, d_off + n_games, sizeof(long long), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess && states_dev && pi_dev && z_dev && total > 0) {
    ao::augment_kernel<<<n_games, ao::kWarps * 32, 0, s>>>(reinterpret_cast<const uint8_t*>(slab_dev), rec_bytes, n_games,
                                                          board_size, tau_thres, d_off, capacity_samples, states_dev,
                                                          pi_dev, z_dev);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  }
  cudaFree(d_off);
  if (n_samples_out) *n_samples_out = total * 8;
  return e == cudaSuccess ? 0 : -100 - (int)e;
}
