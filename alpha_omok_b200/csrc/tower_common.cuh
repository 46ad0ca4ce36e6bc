// Geometry, pass scheduling and constants shared by the tower kernels (tower.cu, tower_stag.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "engine.h"
#include "sm100_ptx.cuh"

namespace ao {
namespace {

constexpr int kC = 128;
constexpr int kTileRows = 128;
constexpr int kTiles = 2;
constexpr int kEpiThreads = 256;
constexpr int kThreads = 352;  // 8 epilogue warps, producer, MMA issuer(s): warp 9 and, for CTA pairs, warp 10
constexpr int kStageBytes = kC * kC * 2;  // one tap, 128 in-channels: 32 KB
constexpr int kStemStageBytes = 16 * kC * 2;
constexpr int kMaxLayers = 21;            // 1 + 2*10 blocks (bias table lives in shared memory)

template <int B>
struct Geo {
  // Board cells are packed densely: row R of the CTA's position stream = cell (R % A) of game (R / A), row stride
  // S = B.  Taps that fall off the board are dropped with tcgen05.mma's disable-output-lane masks instead of zero
  // padding, so 3 games of 9x9 (243 rows) fill the two 128-row tiles to 95 % (1 game of 15x15: 88 %).
  static constexpr int S = B;
  static constexpr int A = B * B;
  static constexpr int GameRows = A;
  static constexpr int GPC = (kTiles * kTileRows) / GameRows;  // games per CTA pass
  static constexpr int Halo = ((S + 1 + 7) / 8) * 8;           // rows addressed (never used) beyond the stream
  static constexpr int Rows = Halo + kTiles * kTileRows + Halo;
  static constexpr int ActBytes = 16 * Rows * 16;
  static_assert(GPC >= 1, "board too large for a 256-row CTA tile");
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Pass k of this CTA. Full waves: GPC games in both tiles. A ragged last wave of r <= gridDim.x games is spread as
// ONE game (one 128-row tile, half the MMA work) per CTA instead of ceil(r/GPC) full passes, so that e.g. 4096 games
// of 9x9 cost 9 + ~0.55 pass times instead of 10.  All warp roles call this with the same arguments.
template <int GPC, int A>
__device__ __forceinline__ bool pass_of_cta(int b, int k, int n, int& g0, int& ng, int& ntiles) {
  const int grid = (int)gridDim.x;
  const int per_wave = GPC * grid;
  const int W = n / per_wave, r = n - W * per_wave;
  if (k < W) {
    g0 = (k * grid + b) * GPC;
    ng = GPC;
    ntiles = kTiles;
    return true;
  }
  g0 = 0;
  ng = 0;
  const bool one_game_tail = GPC > 1 && A <= kTileRows && r <= grid;
  ntiles = one_game_tail ? 1 : kTiles;
  if (k > W || r == 0) return false;
  if (one_game_tail) {
    if (b >= r) return false;
    g0 = W * per_wave + b;
    ng = 1;
    return true;
  }
  g0 = W * per_wave + b * GPC;
  if (g0 >= n) return false;
  ng = min(GPC, n - g0);
  return true;
}
// How many of the n queued requests this launch serves: everything, unless `defer` and the last wave would be ragged
// (fewer than 60 % of GPC * gridDim.x leaves): a ragged wave costs a whole pass (or 0.55 of one as one-game passes), so
// it is cheaper to let those games wait one round - the chip only ever runs full waves.  Warp-uniform, grid-uniform.
template <int GPC>
__device__ __forceinline__ int queue_serve_count(int n, int defer) {
  const int per_wave = GPC * (int)gridDim.x;
  if (defer && n > per_wave) {
    const int r = n % per_wave;
    if (r != 0 && r * 5 < per_wave * 3) n -= r;
  }
  return n;
}
// Called once per CTA when it is done (also by CTAs that had nothing to do): the last one publishes the new head.
__device__ __forceinline__ void queue_finish(const NNQueue& q, uint32_t qbase, int n_served) {
  if (q.tail == nullptr || threadIdx.x != 0) return;
  __threadfence();
  if (atomicAdd(q.done, 1u) == gridDim.x - 1u) {
    *q.head = qbase + (uint32_t)n_served;
    *q.done = 0u;
    __threadfence();
  }
}

// PAIR: the two CTAs of a cluster issue their MMAs together, so a pass exists for both as soon as either has games
// (the other one then runs it with ng = 0).
template <int GPC, int A, bool PAIR>
__device__ __forceinline__ bool get_pass(int k, int n, int& g0, int& ng, int& ntiles) {
  const int b = (int)blockIdx.x;
  const bool mine = pass_of_cta<GPC, A>(b, k, n, g0, ng, ntiles);
  if (!PAIR || mine) return mine;
  int g0p, ngp, ntp;
  return pass_of_cta<GPC, A>(b ^ 1, k, n, g0p, ngp, ntp);
}

}  // namespace
}  // namespace ao
