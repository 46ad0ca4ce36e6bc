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
               float* __restrict__ pis, float* __restrict__ zs, long long ring_cap, long long ring_start,
               long long ring_skip) {
  __shared__ float s_pi[kWarps][kMaxA + 3];
  __shared__ uint16_t s_rows[kWarps][4][32];
  __shared__ uint8_t s_cell[kWarps][kMaxA + 3];   // per ply: bit k = plane k of the cell
  __shared__ uint8_t s_src[8][kMaxA + 3];          // per block: source cell of output cell c under augmentation a
  const int g = blockIdx.x;
  if (g >= n) return;
  const int A = B * B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * A; i += kWarps * 32) s_src[i / A][i % A] = (uint8_t)aug_source(i / A, (i % A) / B, (i % A) % B, B);
  __syncthreads();
  const uint8_t* rec = slab + (size_t)g * rec_bytes;
  const int n_moves = *reinterpret_cast<const int16_t*>(rec);
  const int winner = rec[2];
  const int16_t* mv = reinterpret_cast<const int16_t*>(rec + 4);
  const uint32_t* vis = reinterpret_cast<const uint32_t*>(rec + rec_voff(A));
  const float z_black = winner == 1 ? 1.f : (winner == 2 ? -1.f : 0.f);
  for (int t = warp; t < n_moves; t += kWarps) {
    const long long base = offsets[g] + t;
    // linear mode: sample j lands in slot j (bounded by capacity). ring mode (ring_cap > 0) = deque(maxlen).extend:
    // sample j lands in slot (ring_start + j) % ring_cap; the first ring_skip samples would be pushed out again by
    // later ones of the same extend and are not written at all.
    if (ring_cap == 0 && (base + 1) * 8 > capacity) continue;
    if (ring_cap > 0 && (base + 1) * 8 <= ring_skip) continue;
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
    for (int c = lane; c < A; c += 32) {
      const int y = c / B, x = c % B;
      s_cell[warp][c] = (uint8_t)(((s_rows[warp][0][y] >> x) & 1u) | (((s_rows[warp][1][y] >> x) & 1u) << 1) |
                                  (((s_rows[warp][2][y] >> x) & 1u) << 2) | (((s_rows[warp][3][y] >> x) & 1u) << 3));
    }
    __syncwarp();
    const float colour = black_to_move ? 1.f : 0.f;
    const float z = black_to_move ? z_black : -z_black;
    // ---- the 8 dihedral copies
    for (int a8 = 0; a8 < 8; ++a8) {
      long long s_idx = base * 8 + a8;
      if (ring_cap > 0) {
        if (s_idx < ring_skip) continue;
        s_idx = (ring_start + s_idx) % ring_cap;
      }
      float* so = states + (size_t)s_idx * 5 * A;
      float* po = pis + (size_t)s_idx * A;
      for (int c = lane; c < A; c += 32) {
        const int src = s_src[a8][c];
        const uint32_t bits = s_cell[warp][src];
        so[0 * A + c] = (float)(bits & 1u);
        so[1 * A + c] = (float)((bits >> 1) & 1u);
        so[2 * A + c] = (float)((bits >> 2) & 1u);
        so[3 * A + c] = (float)((bits >> 3) & 1u);
        so[4 * A + c] = colour;
        po[c] = s_pi[warp][src];
      }
      if (lane == 0) zs[s_idx] = z;
    }
    __syncwarp();
  }
}

// train_memory = random.sample(rep_memory, k) (main.py:263-264) as a gather: logical index i of the deque (0 = oldest)
// is ring slot (head + i) % cap. One warp per sample, 4-byte coalesced copies (a 9x9 sample is 405 + 81 + 1 floats).
__global__ void __launch_bounds__(256)
replay_gather_kernel(const float* __restrict__ r_states, const float* __restrict__ r_pi, const float* __restrict__ r_z,
                     long long cap, long long head, const long long* __restrict__ idx, long long k, int A,
                     float* __restrict__ o_states, float* __restrict__ o_pi, float* __restrict__ o_z) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int SA = 5 * A;
  for (long long j = warp0; j < k; j += n_warps) {
    const long long slot = (head + idx[j]) % cap;
    const float* ss = r_states + (size_t)slot * SA;
    float* ds = o_states + (size_t)j * SA;
    for (int c = lane; c < SA; c += 32) ds[c] = ss[c];
    const float* sp = r_pi + (size_t)slot * A;
    float* dp = o_pi + (size_t)j * A;
    for (int c = lane; c < A; c += 32) dp[c] = sp[c];
    if (lane == 0) o_z[j] = r_z[slot];
  }
}

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
                                                          pi_dev, z_dev, 0, 0, 0);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  }
  cudaFree(d_off);
  if (n_samples_out) *n_samples_out = total * 8;
  return e == cudaSuccess ? 0 : -100 - (int)e;
}


// rep_memory.extend(utils.augment_dataset(cur_memory, BOARD_SIZE)) with rep_memory = deque(maxlen=MEMORY_SIZE)
// (main.py:66,250) on the device: the augmented samples of the record slab are written straight into the ring
// (states [cap][5][B][B], pi [cap][A], z [cap], DEVICE pointers) - one pass, no intermediate tensors.  *head_io / *len_io
// (host) are the deque's state: logical item i lives in slot (head + i) % cap.  *n_samples_out = samples appended.
extern "C" int ao_replay_extend_dev(const void* slab_dev, int n_games, int board_size, int tau_thres, float* ring_states_dev,
                                    float* ring_pi_dev, float* ring_z_dev, long long ring_cap, long long* head_io,
                                    long long* len_io, long long* n_samples_out, void* stream) {
  if (!slab_dev || n_games < 1 || board_size < 5 || board_size > ao::kMaxB || ring_cap < 1 || !head_io || !len_io ||
      !ring_states_dev || !ring_pi_dev || !ring_z_dev || *head_io < 0 || *head_io >= ring_cap || *len_io < 0 ||
      *len_io > ring_cap)
    return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int A = board_size * board_size;
  const size_t rec_bytes = ((4 + (size_t)A * 2 + 3) & ~(size_t)3) + (size_t)A * A * 4;
  long long* d_off = nullptr;
  if (cudaMalloc(&d_off, (size_t)(n_games + 1) * sizeof(long long)) != cudaSuccess) return -2;
  ao::sample_offsets_kernel<<<1, 1024, 0, s>>>(reinterpret_cast<const uint8_t*>(slab_dev), rec_bytes, n_games, d_off);
  long long plies = 0;
  cudaError_t e = cudaMemcpyAsync(&plies, d_off + n_games, sizeof(long long), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  const long long N = plies * 8;
  if (e == cudaSuccess && N > 0) {
    const long long skip = N > ring_cap ? N - ring_cap : 0;
    const long long start = (*head_io + *len_io) % ring_cap;
    ao::augment_kernel<<<n_games, ao::kWarps * 32, 0, s>>>(reinterpret_cast<const uint8_t*>(slab_dev), rec_bytes, n_games,
                                                          board_size, tau_thres, d_off, 0, ring_states_dev, ring_pi_dev,
                                                          ring_z_dev, ring_cap, start, skip);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) {
      const long long over = *len_io + N - ring_cap;  // items pushed out at the old end
      if (over > 0) *head_io = (*head_io + over) % ring_cap;
      *len_io = *len_io + N > ring_cap ? ring_cap : *len_io + N;
    }
  }
  cudaFree(d_off);
  if (n_samples_out) *n_samples_out = N;
  return e == cudaSuccess ? 0 : -100 - (int)e;
}

// random.sample(rep_memory, k) (main.py:263-264): idx_dev[k] are the LOGICAL deque indices the host drew
// (random.sample(range(len), k) consumes Python's generator exactly like sampling the deque itself); the samples are
// gathered into out_* (DEVICE, [k][5][B][B] / [k][A] / [k]) in that order.
extern "C" int ao_replay_gather_dev(const float* ring_states_dev, const float* ring_pi_dev, const float* ring_z_dev,
                                    long long ring_cap, long long head, const long long* idx_dev, long long k,
                                    int board_size, float* out_states_dev, float* out_pi_dev, float* out_z_dev,
                                    void* stream) {
  if (!ring_states_dev || !ring_pi_dev || !ring_z_dev || !idx_dev || !out_states_dev || !out_pi_dev || !out_z_dev ||
      ring_cap < 1 || head < 0 || head >= ring_cap || k < 0 || board_size < 5 || board_size > ao::kMaxB)
    return -1;
  if (k == 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long blocks_needed = (k + 7) / 8;
  const int blocks = (int)(blocks_needed < 148 * 8 ? blocks_needed : 148 * 8);
  ao::replay_gather_kernel<<<blocks, 256, 0, s>>>(ring_states_dev, ring_pi_dev, ring_z_dev, ring_cap, head, idx_dev, k,
                                                  board_size * board_size, out_states_dev, out_pi_dev, out_z_dev);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -100 - (int)e;
}
