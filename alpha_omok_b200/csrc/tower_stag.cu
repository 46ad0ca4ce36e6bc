// PVNet.forward (model.py:97-104): staggered-tile CTA-pair tower kernel (AO_NN_FP16, the default).
//
// Same data layout and MMAs as the CTA-pair kernel of tower.cu (dense board packing, tap = shifted start row of the
// activation buffer, disable-output-lane masks for off-board taps, tcgen05 cta_group::2 M256 N128 K16, fp32 accumulators
// and fp32 residual stream in TMEM), but the two 128-row tiles of a pass run half a layer apart so that the epilogue of
// one tile (TMEM -> bias/ReLU -> fp16 operand rows in smem) runs while the tensor core works on the other tile,
// instead of MMA and epilogue taking turns:
//
//   stream 0 (warp 9), tile T0:  T0[centre, negative-shift taps]  T0[positive-shift taps]   commit -> acc T0, acc T1
//   needs (previous layer):         epilogue of T0                   epilogue of T1
//   stream 1 (warp 8), tile T1:  T1[centre, negative-shift taps]  commit -> acc T0   T1[positive-shift taps]  commit -> acc T1
//   needs (previous layer):         epilogues of T0 and T1
//
// With dense packing the tiles depend on each other only through the +-(B+1) boundary rows: taps with a negative row
// shift read rows below, so T0's centre/negative taps need T0's own epilogue only and its positive taps also T1's first
// rows; T1's negative taps read T0's last rows.  A tile's epilogue overwrites its rows in place, so it may start only
// when the tile is accumulated AND the other stream no longer reads its boundary rows: every acc barrier collects one
// commit from each stream.  All 8 epilogue warps work on ONE tile at a time (64 accumulator columns per thread) and
// signal their two 32-column groups separately, so stream 0 can start the matching k-steps of the next layer early.
// The epilogues stagger the two streams by themselves: while one tile is in its epilogue the other stream has the
// tensor pipe alone.  Both tiles use every tap, so the weight ring holds exactly one layer of this CTA's half of B
// (9 slots x 16 KB, slot = tap); a slot is released by one commit from each stream and the next layer's tap streams
// in behind it.
//
// Warp roles (480 threads): 0-7 epilogue, 8 and 9 the MMA issuers of tile 1 / tile 0 (leader CTA; in the peer CTA warp 9
// forwards weights-landed), 10-13 heads (1x1 conv outputs -> FC/softmax/tanh of the PREVIOUS pass while the tower of the
// next pass is already running), 14 weight/bias producer.
// Two issuing warps on two different scheduler ports: a single thread gets one M256 N128 K16 MMA per ~109 cycles out of
// the tensor pipe, two threads with their own accumulators one per ~87 (tools/probe_mma_rate.py).
#include <cuda_fp16.h>
#include <stdio.h>
#include <string.h>

#include "tower_common.cuh"
#include "tree_device.cuh"

namespace ao {
namespace {

constexpr int kStagThreads = 480;
constexpr int kHeadThreads = 128;
constexpr int kMaxPassesPerCta = 64;  // persistent mode: passes a CTA owns (4096 games of 15x15 on 148 CTAs: 28)

// PERSIST (the persistent self-play kernel): the head scratch is shared with the per-warp scratch of the tree step the
// head warps run after their head job (tree_device.cuh WarpSmem, sized for this board), and the network's answer is
// handed over in shared memory (pol / val) instead of through HBM.
template <int B, bool PERSIST>
struct StagSmem {
  using G = Geo<B>;
  static constexpr int kSlots = 9;
  static constexpr int kSlotBytes = kStageBytes / 2;
  static constexpr int APad = (G::A + 7) / 8 * 8;
  static constexpr int act = 0;
  static constexpr int wring = G::ActBytes;
  static constexpr int bias2 = wring + kSlots * kSlotBytes;           // [2][128] f32: bias of the layer in flight
  static constexpr int headw = bias2 + 2 * kC * 4;                    // [3][128] f32
  static constexpr int feat = headw + 3 * kC * 4;                     // [GPC][3][A] f32
  static constexpr int logits = feat + G::GPC * 3 * G::A * 4;         // [GPC][A]
  static constexpr int hidden = logits + G::GPC * G::A * 4;           // [GPC][128]
  static constexpr int red = hidden + G::GPC * kC * 4;                // [GPC][2]
  static constexpr int head_end = red + G::GPC * 2 * 4;
  // tree-step scratch of head warp g (aliases feat .. red, which are dead once the policy is final):
  // dbuf f64[APad] | dbuf2 f64[APad] | order u8[256] | table i16[128] | rows u16[2][32]
  static constexpr int kTreeWarpBytes = 16 * APad + 256 + 256 + 128;
  static constexpr int tree_end = feat + G::GPC * kTreeWarpBytes;
  static constexpr int scratch_end = PERSIST ? (tree_end > head_end ? tree_end : head_end) : head_end;
  static constexpr int pol = (scratch_end + 15) / 16 * 16;            // PERSIST: [GPC][APad] f32 final policy
  static constexpr int val = pol + (PERSIST ? G::GPC * APad * 4 : 0); // PERSIST: [4] f32 values
  static constexpr int leaf_done = val + (PERSIST ? 16 : 0);          // PERSIST: [kMaxPassesPerCta] u32
  static constexpr int masks = (leaf_done + (PERSIST ? kMaxPassesPerCta * 4 : 0) + 15) / 16 * 16;  // [kTiles][9][4]
  static constexpr int bars = masks + kTiles * 9 * 4 * 4;
  static constexpr int kBars = 3 * kSlots + 4 + 2 + 2 + 2;
  static constexpr int total = bars + kBars * 8 + 16;
  static_assert(G::ActBytes % 1024 == 0, "weight ring alignment");
  static_assert(feat % 8 == 0, "tree scratch holds doubles");
};

__device__ __forceinline__ void head_bar_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// PERSIST = the persistent self-play kernel: `rounds` lock-step rounds in ONE launch.  Game g is always evaluated in
// the same pass of the same CTA (request slot = game slot, TreeParams.static_slots), and as soon as the head warps have
// the policy / value of a pass they run the tree step of those games themselves (consume the answer, expand, back up,
// play the move when the search is complete, select the next leaf and write its request) while the tower of the next
// pass is already running.  Games never interact, so no grid-wide synchronisation is needed: every CTA pair cycles
// through its own games; the tree work, the launch gaps and the per-launch ramp of the two-kernel round disappear
// behind the tensor pipe.
template <int B, bool PERSIST>
__global__ void __launch_bounds__(kStagThreads, 1)
tower_stag_kernel(TowerWeights W, const LeafIn* __restrict__ in, NNQueue q, int n_max,
                  float* __restrict__ policy, float* __restrict__ value, TreeParams P, int rounds, int free_passes,
                  uint32_t* __restrict__ cta_pos) {
  using G = Geo<B>;
  using SL = StagSmem<B, PERSIST>;
  constexpr bool PAIR = true;
  constexpr int NSLOT = SL::kSlots;
  constexpr uint32_t kSlotBytes = SL::kSlotBytes;
  constexpr uint32_t kTapBytes = kStageBytes / 2, kStemTapBytes = kStemStageBytes / 2, kBRows = kC / 2;
  extern __shared__ __align__(1024) uint8_t smem[];

  int n = n_max;
  uint32_t qbase = 0u;  // ring position of request 0 of this launch
  if (!PERSIST && q.tail != nullptr) {
    qbase = *q.head;
    n = (int)(*q.tail - qbase);
    if (n > n_max) n = n_max;
    n = queue_serve_count<G::GPC>(n, q.defer);
  }
  const uint32_t qmask = (!PERSIST && q.tail != nullptr) ? q.mask : 0xFFFFFFFFu;
  if (!PERSIST) rounds = 1;
  // PERSIST, free-running (free_passes > 0; enough games per CTA): CTA b owns the games b, b + grid, b + 2 grid, ... and
  // cycles through them with FULL passes only - pass q holds the games at list positions GPC*q .. GPC*q + GPC-1 (mod
  // the list length) - for `free_passes` passes per launch, the same number on every CTA.  No ragged tail pass, no round
  // structure: a CTA with one game fewer simply comes round to its games a little sooner.  cta_pos[b] carries the
  // cycle position from launch to launch.
  const bool free_run = PERSIST && free_passes > 0;
  if (free_run) rounds = 1;
  const int n_local = free_run ? (n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 1;
  const uint32_t q0 = free_run ? cta_pos[blockIdx.x] : 0u;
  // pass k of this launch: static map (get_pass) or the next full pass of the cycle
  auto next_pass = [&](int k, int& g0, int& ng, int& ntiles) -> bool {
    if (free_run) {
      g0 = 0;
      ng = G::GPC;
      ntiles = kTiles;
      return k < free_passes;
    }
    return get_pass<G::GPC, G::A, PAIR>(k, n, g0, ng, ntiles);
  };
  // list position / game slot of game j of pass k (free-running), request slot of game j otherwise
  auto list_pos = [&](int k, int j) -> int { return (int)(((q0 + (uint32_t)k) * (uint32_t)G::GPC + (uint32_t)j) % (uint32_t)n_local); };
  auto game_slot = [&](int k, int g0, int j) -> uint32_t {
    if (free_run) return (uint32_t)((int)blockIdx.x + (int)gridDim.x * list_pos(k, j));
    return (qbase + (uint32_t)(g0 + j)) & qmask;
  };
  {
    int g0_, ng_, nt_;
    if (!next_pass(0, g0_, ng_, nt_)) {
      if (!PERSIST) queue_finish(q, qbase, n);
      return;
    }
  }
  uint8_t* s_act = smem + SL::act;
  uint8_t* s_w = smem + SL::wring;
  float* s_bias2 = reinterpret_cast<float*>(smem + SL::bias2);
  float* s_headw = reinterpret_cast<float*>(smem + SL::headw);
  float* s_feat = reinterpret_cast<float*>(smem + SL::feat);
  float* s_logits = reinterpret_cast<float*>(smem + SL::logits);
  float* s_hidden = reinterpret_cast<float*>(smem + SL::hidden);
  float* s_red = reinterpret_cast<float*>(smem + SL::red);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + SL::masks);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + SL::bars);
  uint64_t* bar_empty = bar_full + NSLOT;
  uint64_t* bar_peer_full = bar_empty + NSLOT;  // leader only: the peer CTA's half of slot s has landed
  uint64_t* bar_act = bar_peer_full + NSLOT;    // [tile][column group] leader only: epilogue warps of BOTH CTAs -> MMA issuer
  uint64_t* bar_acc = bar_act + 4;              // [tile] MMA -> epilogue: the tile's layer is accumulated
  uint64_t* bar_bias = bar_acc + 2;             // [slot] producer -> epilogue: the layer's bias has landed
  uint64_t* bar_feat_full = bar_bias + 2;       // epilogue -> heads: 1x1-conv sums of the pass are complete
  uint64_t* bar_feat_free = bar_feat_full + 1;  // heads -> epilogue: s_feat is consumed and zeroed again
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_feat_free + 1);
  float* s_pol = reinterpret_cast<float*>(smem + SL::pol);                       // PERSIST only
  float* s_val = reinterpret_cast<float*>(smem + SL::val);
  volatile uint32_t* s_leaf_done = reinterpret_cast<volatile uint32_t*>(smem + SL::leaf_done);
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_layers = W.n_layers;

  // ---------------- one-time setup
  for (int i = tid; i < G::ActBytes / 16; i += kStagThreads) reinterpret_cast<uint4*>(s_act)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 3 * kC; i += kStagThreads) s_headw[i] = W.head_w[i];
  for (int i = tid; i < G::GPC * 3 * G::A; i += kStagThreads) s_feat[i] = 0.f;
  if (PERSIST)
    for (int i = tid; i < kMaxPassesPerCta; i += kStagThreads) s_leaf_done[i] = 0u;
  for (int i = tid; i < kTiles * 9 * 4; i += kStagThreads) {
    const int tile = i / 36, tap = (i / 4) % 9, word = i % 4;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    uint32_t m = 0;
    for (int b = 0; b < 32; ++b) {
      const int pos = (tile * kTileRows + word * 32 + b) % G::A;
      const int y = pos / B + dy, x = pos % B + dx;
      if (y < 0 || y >= B || x < 0 || x >= B) m |= 1u << b;
    }
    s_mask[i] = m;
  }
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 2);  // both MMA streams are done with the tap
      mbar_init(&bar_peer_full[s], 1);
    }
    for (int t = 0; t < 4; ++t) mbar_init(&bar_act[t], 16);  // lane 0 of the 8 epilogue warps of each CTA of the pair
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bar_acc[t], 2);  // own tile accumulated + the other stream no longer reads this tile's boundary rows
      mbar_init(&bar_bias[t], 1);
    }
    mbar_init(bar_feat_full, 8);
    mbar_init(bar_feat_free, 4);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_pair<512>(s_tmem);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  if (warp == 14) {
    // =========================================================== weight + bias producer: slot = tap, one layer in flight
    if (lane == 0) {
      uint32_t lc = 0;
      int g0, ng, ntiles;
      for (int rd = 0; rd < rounds; ++rd)
      for (int k = 0; next_pass(k, g0, ng, ntiles); ++k) {
        size_t off = 0;
        for (int l = 0; l < n_layers; ++l, ++lc) {
          const uint32_t part = AO_XFLAG(W, 8) ? 64u : (l == 0 ? kStemTapBytes : kTapBytes);
          for (int st = 0; st < 9; ++st) {
            mbar_wait(&bar_empty[st], (lc & 1u) ^ 1u);
            if (st == 0) {
              // slot 0 of layer l is free => layer l-1 is fully issued => the epilogue of layer l-2 (last reader of
              // this bias slot) has finished long ago
              mbar_arrive_expect_tx(&bar_bias[lc & 1u], kC * 4);
              bulk_g2s(s_bias2 + (lc & 1u) * kC, W.bias + l * kC, kC * 4, &bar_bias[lc & 1u]);
            }
            mbar_arrive_expect_tx(&bar_full[st], part);
            bulk_g2s(s_w + st * kSlotBytes,
                     reinterpret_cast<const uint8_t*>(W.conv_pair) + off + ((size_t)st * 2 + cta_rank) * part, part,
                     &bar_full[st]);
          }
          off += l == 0 ? 9u * kStemStageBytes : 9u * kStageBytes;
        }
      }
    }
  } else if (warp == 9 || warp == 8) {
    const int stream = 9 - warp;  // MMA stream 0 (warp 9) owns tile 0, stream 1 (warp 8) tile 1: two scheduler ports
    if (!leader) {
      // ---- peer CTA: forward "my half of slot s has landed" to the leader (the operand-ready arrivals go there directly)
      if (stream == 0 && lane == 0) {
        int g0, ng, ntiles, n_k = 0;
        while (next_pass(n_k, g0, ng, ntiles)) ++n_k;
        const uint32_t n_stage = (uint32_t)(n_k * n_layers) * 9u * (uint32_t)rounds;
        for (uint32_t si = 0; si < n_stage; ++si) {
          const uint32_t s = si % 9u;
          mbar_wait(&bar_full[s], (si / 9u) & 1u);
          mbar_arrive_remote_relaxed(&bar_peer_full[s], 0u);
        }
      }
    } else {
      // =========================================================== MMA issuers (leader CTA), one per tile
      const uint32_t idesc = umma_idesc_f16_f32(256, 128);
      const uint32_t lbo_a = (uint32_t)G::Rows * 16u;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(s_act), lbo_a);
      const uint32_t desc_hi = umma_desc_hi(128u);
      constexpr uint32_t kAStep = (2u * (uint32_t)G::Rows * 16u) >> 4;
      constexpr uint32_t kBStep = (2u * kBRows * 16u) >> 4;
      uint32_t lc = 0, act_ph = 0;
      AO_DBG(long long dbg_act_wait = 0, dbg_full_wait = 0; const long long dbg_t0 = W.dbg ? clock64() : 0;)
      int g0, ng, ntiles;
      for (int rd = 0; rd < rounds; ++rd)
      for (int k = 0; next_pass(k, g0, ng, ntiles); ++k) {
        if (stream == 1 && ntiles < kTiles) {  // one-game tail pass: tile 1 holds no board
          lc += (uint32_t)n_layers;
          // tile 0's operand-ready barriers complete n_layers phases in this pass without this warp looking: keep the
          // parities in step (matters in the persistent kernel, where a two-tile pass of the next round follows)
          if (n_layers & 1) act_ph ^= 3u;
          continue;
        }
        for (int l = 0; l < n_layers; ++l, ++lc) {
          const bool to_b = (l & 1) == 0;      // stem and conv2 accumulate in accB (holds the block input x)
          const bool residual = to_b && l > 0;
          const uint32_t full_ph = lc & 1u;
          const int nk = l == 0 ? 1 : kC / 16;
          // taps [st_lo, st_hi) of `tile`, k-steps of column group `grp` (0: k-steps {0,1,4,5} = the accumulator
          // columns every epilogue thread converts first, 1: {2,3,6,7}, 2: all); first_use: wait for the slot's
          // weights; release: hand the slot back
          auto issue = [&](const int tile, const int st_lo, const int st_hi, const int grp, const bool first_use,
                           const int release) {  // release: number of arrivals on the slot's empty barrier (0, 1, 2)
            const bool no_mma = nk == 1 && grp == 1;  // stem: a single k-step (group 0); only the releases remain
            for (int st = st_lo; st < st_hi; ++st) {
              if (first_use) {
                AO_DBG(const long long t_f0 = W.dbg ? clock64() : 0;)
                mbar_wait(&bar_full[st], full_ph);
                mbar_wait_cluster(&bar_peer_full[st], full_ph);
                tc_fence_after_sync();
                AO_DBG(if (W.dbg) dbg_full_wait += clock64() - t_f0;)
              }
              const int t = st == 0 ? 4 : (st <= 4 ? st - 1 : st);  // packed order: centre, 4 negative, 4 positive shifts
              const int shift = AO_XFLAG(W, 1) ? 0 : (t / 3 - 1) * G::S + (t % 3 - 1);
              const uint32_t b_lo0 = umma_desc_lo(smem_u32(s_w + st * kSlotBytes), kBRows * 16u);
              if (elect_one()) {
                const uint32_t* mk = s_mask + (tile * 9 + t) * 4;
                const uint32_t m0 = mk[0], m1 = mk[1], m2 = mk[2], m3 = mk[3];
                const uint32_t d_tmem = tmem + (uint32_t)(tile * 256 + (to_b ? 128 : 0));
                const uint32_t a_lo = a_lo0 + (uint32_t)(G::Halo + tile * kTileRows + shift);
                // the very first MMA of a fresh accumulation is (centre tap, k-step 0): it writes every row
                uint32_t acc = (residual || st > 0 || grp == 1) ? 1u : 0u;
#pragma unroll
                for (int j = 0; j < kC / 16; ++j) {
                  if (j >= nk || no_mma) break;
                  if (grp != 2 && ((j >> 1) & 1) != grp) continue;
                  umma_f16_ss_pair_masked(d_tmem, a_lo + (uint32_t)j * kAStep, b_lo0 + (uint32_t)j * kBStep, desc_hi, idesc,
                                          acc, m0, m1, m2, m3);
                  acc = 1u;
                }
                for (int r = 0; r < release; ++r) umma_commit_pair(&bar_empty[st]);
              }
              __syncwarp();
            }
          };
          auto wait_act = [&](const int idx) {  // idx = tile * 2 + column group
            AO_DBG(const long long t_a0 = W.dbg ? clock64() : 0;)
            mbar_wait_cluster(&bar_act[idx], (act_ph >> idx) & 1u);
            act_ph ^= 1u << idx;
            tc_fence_after_sync();
            AO_DBG(if (W.dbg) dbg_act_wait += clock64() - t_a0;)
          };
          auto commit_acc = [&](const int tile) {
            if (elect_one()) umma_commit_pair(&bar_acc[tile]);
            __syncwarp();
          };
          // Every accumulator row sees the SAME order of its 72 MMAs whatever tile or kind of pass it sits in - taps
          // 0-4 x k-steps {0,1,4,5}, taps 0-4 x {2,3,6,7}, taps 5-8 x {0,1,4,5}, taps 5-8 x {2,3,6,7} - so a leaf's floats do
          // not depend on where the scheduler put it (tensor-core accumulation is order-sensitive in the last bits).
          if (ntiles < kTiles) {          // one-game tail pass: stream 0 alone, arriving for both streams
            wait_act(0);
            wait_act(1);
            issue(0, 0, 5, 0, true, 0);
            issue(0, 0, 5, 1, false, 2);
            issue(0, 5, 9, 0, true, 0);
            issue(0, 5, 9, 1, false, 2);
            commit_acc(0);
            commit_acc(0);
          } else if (stream == 0) {       // T0: centre/negative taps need T0's epilogue only, positive taps also T1's
            wait_act(0);
            issue(0, 0, 5, 0, true, 0);
            wait_act(1);
            issue(0, 0, 5, 1, false, 1);
            wait_act(2);
            issue(0, 5, 9, 0, true, 0);
            wait_act(3);
            issue(0, 5, 9, 1, false, 1);
            commit_acc(0);                // T0 accumulated
            commit_acc(1);                // T0 no longer reads T1's first rows
          } else {                        // T1: every tap reads T1's own rows, the negative ones also T0's last rows
            wait_act(0);
            wait_act(1);
            wait_act(2);
            wait_act(3);
            issue(1, 0, 5, 0, true, 0);
            issue(1, 0, 5, 1, false, 1);
            commit_acc(0);                // T1 no longer reads T0's last rows: T0's epilogue may overwrite them
            issue(1, 5, 9, 0, true, 0);
            issue(1, 5, 9, 1, false, 1);
            commit_acc(1);                // T1 accumulated
          }
        }
      }
#ifdef AO_PROBE
      if (W.dbg && blockIdx.x == 0 && lane == 0 && stream == 0) {
        atomicAdd(&W.dbg[0], (unsigned long long)(clock64() - dbg_t0));
        atomicAdd(&W.dbg[1], (unsigned long long)dbg_act_wait);
        atomicAdd(&W.dbg[2], (unsigned long long)dbg_full_wait);
        atomicAdd(&W.dbg[3], 1ull);
      }
#endif
    }
  } else if (warp < 8) {
    // =========================================================== epilogue: all 8 warps on one tile at a time
    const int q = warp & 3, half = warp >> 2;   // TMEM lane quarter, accumulator column half
    const int r = q * 32 + lane;                // row of the tile this thread owns
    const uint32_t chunk_stride = (uint32_t)G::Rows * 16u;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    // input planes: thread (half, r) writes row r of tile `half` (256 threads <-> 256 rows)
    const int R_in = half * kTileRows + r;
    const int gl_in = R_in / G::A, pos_in = R_in % G::A;
    uint32_t acc_ph0 = 0, acc_ph1 = 0, lc = 0, pass_ph = 0;
    AO_DBG(long long dbg_acc_wait = 0, dbg_heads = 0; const bool dbg_on = W.dbg != nullptr && blockIdx.x == 0 && tid == 0;
           const long long dbg_e0 = dbg_on ? clock64() : 0;)
    // "this warp's rows of tile t are written": all lanes fence, lane 0 arrives on the LEADER's barrier
    auto arrive_act = [&](const int idx) {  // idx = tile * 2 + column group
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&bar_act[idx]);
        else mbar_arrive_remote(&bar_act[idx], 0u);
      }
    };

    int g0, ng, ntiles;
    for (int rd = 0; rd < rounds; ++rd)
    for (int k = 0; next_pass(k, g0, ng, ntiles); ++k) {
      if (PERSIST && !free_run && rd > 0) {
        // the requests of this pass were written by this CTA's head warps one round ago (long done; just make sure)
        while (s_leaf_done[k] < (uint32_t)(rd * ng)) __nanosleep(200);
        __threadfence_block();
      }
      if (free_run && gl_in < G::GPC) {
        // this row's game was last evaluated n_local list positions ago: its next request must have been written since
        const uint32_t t = (uint32_t)k * (uint32_t)G::GPC + (uint32_t)gl_in;  // list position since the launch began
        const uint32_t need = t / (uint32_t)n_local;
        const int li = list_pos(k, gl_in);
        while (s_leaf_done[li] < need) __nanosleep(200);
        __threadfence_block();
      }
      {
        uint4 c0 = make_uint4(0, 0, 0, 0);
        if (gl_in < G::GPC && gl_in < ng) {
          const LeafIn* li = &in[game_slot(k, g0, gl_in)];
          const int yy = pos_in / B, xx = pos_in % B;
          // L1-bypassing loads: in persistent mode the request was stored by another warp of this SM moments ago
          const uint32_t b0 = (__ldcg(&li->plane[0][yy]) >> xx) & 1u, b1 = (__ldcg(&li->plane[1][yy]) >> xx) & 1u;
          const uint32_t b2 = (__ldcg(&li->plane[2][yy]) >> xx) & 1u, b3 = (__ldcg(&li->plane[3][yy]) >> xx) & 1u;
          const uint32_t b4 = __ldcg(&li->colour) & 1u;
          c0.x = (b0 ? 0x3C00u : 0u) | (b1 ? 0x3C000000u : 0u);
          c0.y = (b2 ? 0x3C00u : 0u) | (b3 ? 0x3C000000u : 0u);
          c0.z = (b4 ? 0x3C00u : 0u);
        }
        const uint32_t ro = (uint32_t)(G::Halo + R_in) * 16u;
        *reinterpret_cast<uint4*>(s_act + ro) = c0;
        *reinterpret_cast<uint4*>(s_act + chunk_stride + ro) = make_uint4(0, 0, 0, 0);
      }
      // every warp wrote rows of ONE tile only, but both barriers expect all 16 warps
      arrive_act(0);
      arrive_act(1);
      if (ntiles == kTiles) {
        arrive_act(2);
        arrive_act(3);
      }

      for (int l = 0; l < n_layers; ++l, ++lc) {
        const bool to_b = (l & 1) == 0;
        const bool last = l == n_layers - 1;
        const float* bias = s_bias2 + (lc & 1u) * kC;
        mbar_wait(&bar_bias[lc & 1u], (lc >> 1) & 1u);
        for (int t = 0; t < ntiles; ++t) {
          const int R = t * kTileRows + r;
          const int g_local = R / G::A, pos = R % G::A;
          const bool valid = g_local < G::GPC && g_local < ng;
          const uint32_t row_off = (uint32_t)(G::Halo + R) * 16u;
          AO_DBG(const long long t_w0 = dbg_on ? clock64() : 0;)
          mbar_wait(&bar_acc[t], t ? acc_ph1 : acc_ph0);
          if (t) acc_ph1 ^= 1u;
          else acc_ph0 ^= 1u;
          tc_fence_after_sync();
          AO_DBG(if (dbg_on) dbg_acc_wait += clock64() - t_w0;)
          const uint32_t lane_addr = lane_base + (uint32_t)(t * 256);
          const uint32_t stash_addr = lane_addr + 128u;
          const uint32_t acc_addr = to_b ? stash_addr : lane_addr;
          float hd0 = 0.f, hd1 = 0.f, hd2 = 0.f;
          auto process = [&](uint32_t (&v)[32], const int qd) {  // qd: 32-column group of the 128 channels
            const float4* b4 = reinterpret_cast<const float4*>(bias + qd * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bb = b4[j4];
              const float y0 = fmaxf(__uint_as_float(v[j4 * 4 + 0]) + bb.x, 0.f);
              const float y1 = fmaxf(__uint_as_float(v[j4 * 4 + 1]) + bb.y, 0.f);
              const float y2 = fmaxf(__uint_as_float(v[j4 * 4 + 2]) + bb.z, 0.f);
              const float y3 = fmaxf(__uint_as_float(v[j4 * 4 + 3]) + bb.w, 0.f);
              // rows beyond the pass's games hold finite garbage: no on-board tap of a real row ever reads them
              v[j4 * 4 + 0] = __float_as_uint(y0);
              v[j4 * 4 + 1] = __float_as_uint(y1);
              v[j4 * 4 + 2] = __float_as_uint(y2);
              v[j4 * 4 + 3] = __float_as_uint(y3);
            }
            if (!last) {
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                uint4 pk;
                __half2 h;
                h = __floats2half2_rn(__uint_as_float(v[cc * 8 + 0]), __uint_as_float(v[cc * 8 + 1]));
                pk.x = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(__uint_as_float(v[cc * 8 + 2]), __uint_as_float(v[cc * 8 + 3]));
                pk.y = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(__uint_as_float(v[cc * 8 + 4]), __uint_as_float(v[cc * 8 + 5]));
                pk.z = *reinterpret_cast<uint32_t*>(&h);
                h = __floats2half2_rn(__uint_as_float(v[cc * 8 + 6]), __uint_as_float(v[cc * 8 + 7]));
                pk.w = *reinterpret_cast<uint32_t*>(&h);
                if (!AO_XFLAG(W, 32)) *reinterpret_cast<uint4*>(s_act + (uint32_t)(qd * 4 + cc) * chunk_stride + row_off) = pk;
              }
              if (to_b && !AO_XFLAG(W, 16)) tmem_st32(stash_addr + (uint32_t)(qd * 32), v);  // fp32 block input for the next residual add
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float x = __uint_as_float(v[j]);
                hd0 = fmaf(x, s_headw[0 * kC + qd * 32 + j], hd0);
                hd1 = fmaf(x, s_headw[1 * kC + qd * 32 + j], hd1);
                hd2 = fmaf(x, s_headw[2 * kC + qd * 32 + j], hd2);
              }
            }
          };
          if (!AO_XFLAG(W, 4) || last) {
            uint32_t va[32], vb[32];
            if (!AO_XFLAG(W, 64) || last) {
              tmem_ld32(acc_addr + (uint32_t)(half * 64), va);
              tmem_ld32(acc_addr + (uint32_t)(half * 64 + 32), vb);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) va[j] = vb[j] = (uint32_t)(j + lane) << 20;
            }
            process(va, half * 2);
            if (!last) arrive_act(t * 2);  // k-steps {0,1} / {4,5} of the next layer can start
            process(vb, half * 2 + 1);
          } else if (!last) {
            arrive_act(t * 2);
          }
          if (!last) {
            if (to_b) tmem_st_wait();
            arrive_act(t * 2 + 1);
          } else {
            // heads' 1x1 convolutions (model.py:44-46, 64-66): the two column halves of a row add up in shared memory
            // (two commutative float adds onto 0: the result does not depend on their order)
            if (t == 0) mbar_wait(bar_feat_free, pass_ph ^ 1u);  // the head warps are done with the previous pass
            if (valid) {
              float* f = s_feat + g_local * 3 * G::A;
              atomicAdd(&f[0 * G::A + pos], hd0);
              atomicAdd(&f[1 * G::A + pos], hd1);
              atomicAdd(&f[2 * G::A + pos], hd2);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_feat_full);
      pass_ph ^= 1u;
    }
#ifdef AO_PROBE
    if (dbg_on) {
      atomicAdd(&W.dbg[4], (unsigned long long)(clock64() - dbg_e0));
      atomicAdd(&W.dbg[5], (unsigned long long)dbg_acc_wait);
      atomicAdd(&W.dbg[6], (unsigned long long)dbg_heads);
    }
#endif
  } else {
    // =========================================================== heads (model.py:43-50, 63-73), one pass behind the tower
    const int ht = tid - 10 * 32, hw = warp - 10;
    uint32_t pass_ph = 0;
    int g0, ng, ntiles;
    for (int rd = 0; rd < rounds; ++rd)
    for (int k = 0; next_pass(k, g0, ng, ntiles); ++k) {
      mbar_wait_sleep(bar_feat_full, pass_ph, 2000);  // a whole tower pass (~100 us) between two head jobs
      pass_ph ^= 1u;
      for (int i = ht; i < G::GPC * 3 * G::A; i += kHeadThreads)
        s_feat[i] = fmaxf(s_feat[i] + W.head_b[(i / G::A) % 3], 0.f);
      head_bar_sync();
      for (int o = ht; o < G::GPC * G::A; o += kHeadThreads) {  // policy FC: (game, output)
        const int pg = o / G::A, po = o % G::A;
        if (pg < ng) {
          const float* f = s_feat + pg * 3 * G::A;
          float acc = W.pfc_b[po];
          const float* wt = W.pfc_wT + po;
#pragma unroll 18
          for (int kk = 0; kk < 2 * G::A; ++kk) acc = fmaf(__ldg(wt + (size_t)kk * G::A), f[kk], acc);
          s_logits[o] = acc;
        }
      }
      for (int vi = ht; vi < G::GPC * kC; vi += kHeadThreads) {  // value FC1: (game, hidden unit)
        const int vg = vi / kC, vj = vi % kC;
        if (vg < ng) {
          const float* f = s_feat + vg * 3 * G::A + 2 * G::A;
          float acc = W.vfc1_b[vj];
          const float* wt = W.vfc1_wT + vj;
#pragma unroll 27
          for (int kk = 0; kk < G::A; ++kk) acc = fmaf(__ldg(wt + (size_t)kk * kC), f[kk], acc);
          s_hidden[vi] = fmaxf(acc, 0.f) * W.vfc2_w[vj];
        }
      }
      head_bar_sync();
      if (hw < ng) {  // head warp g: softmax statistics and the value of game g (GPC <= 3 games per pass)
        float mx = -3.0e38f;
        for (int kk = lane; kk < G::A; kk += 32) mx = fmaxf(mx, s_logits[hw * G::A + kk]);
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        float sum = 0.f;
        for (int kk = lane; kk < G::A; kk += 32) sum += expf(s_logits[hw * G::A + kk] - mx);
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        float hv = 0.f;
        for (int kk = lane; kk < kC; kk += 32) hv += s_hidden[hw * kC + kk];
#pragma unroll
        for (int o = 16; o; o >>= 1) hv += __shfl_xor_sync(0xFFFFFFFFu, hv, o);
        if (lane == 0) {
          s_red[hw * 2 + 0] = mx;
          s_red[hw * 2 + 1] = sum;
          const float v = tanhf(hv + W.vfc2_b);
          if (PERSIST) s_val[hw] = v;
          else value[game_slot(k, g0, hw)] = v;
        }
      }
      if (!PERSIST)
        for (int i = ht; i < G::GPC * 3 * G::A; i += kHeadThreads) s_feat[i] = 0.f;  // ready for the next pass's sums
      head_bar_sync();
      for (int o = ht; o < G::GPC * G::A; o += kHeadThreads) {
        const int pg = o / G::A, po = o % G::A;
        if (pg < ng) {
          const float pr = expf(s_logits[o] - s_red[pg * 2]) / s_red[pg * 2 + 1];
          if (PERSIST) s_pol[pg * SL::APad + po] = pr;
          else policy[(size_t)game_slot(k, g0, pg) * G::A + po] = pr;
        }
      }
      head_bar_sync();  // s_logits / s_red are rewritten by the next pass only after this
      if (PERSIST) {
        // ---- the tree step of this pass's games (tree_device.cuh), head warp g <-> game g0 + g: the head scratch is
        // dead now and becomes the warp's tree scratch; the answer is read from s_pol / s_val
        if (hw < ng) {
          uint8_t* ts = smem + SL::feat + hw * SL::kTreeWarpBytes;
          WarpSmem ws;
          ws.dbuf = reinterpret_cast<double*>(ts);
          ws.dbuf2 = ws.dbuf + SL::APad;
          ws.order = reinterpret_cast<uint8_t*>(ws.dbuf2 + SL::APad);
          ws.table = reinterpret_cast<int16_t*>(ws.order + 256);
          ws.rows = reinterpret_cast<uint16_t(*)[32]>(ws.table + 128);
          ws.pol = s_pol + hw * SL::APad;
          (void)tree_step_game<(G::A <= 96 ? 3 : 8)>(P, (int)game_slot(k, g0, hw), &ws, lane, 64, true, s_val[hw]);
          __threadfence_block();
          __syncwarp();
          if (lane == 0) atomicAdd(const_cast<uint32_t*>(s_leaf_done) + (free_run ? list_pos(k, hw) : k), 1u);
        }
        head_bar_sync();
        for (int i = ht; i < G::GPC * 3 * G::A; i += kHeadThreads) s_feat[i] = 0.f;  // ready for the next pass's sums
        head_bar_sync();
      }
      if (lane == 0) mbar_arrive(bar_feat_free);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves (or frees TMEM) while its partner may still signal / use it
  if (warp == 9) tmem_dealloc_pair<512>(tmem);
  if (!PERSIST) queue_finish(q, qbase, n);
  if (free_run && tid == 0) cta_pos[blockIdx.x] = q0 + (uint32_t)free_passes;
}

template <int B, bool PERSIST>
cudaError_t launch_tower_stag_t(const TowerWeights& w, const LeafIn* in, const NNQueue& q, int n_max, float* policy,
                                float* value, int num_sms, const TreeParams& P, int rounds, int free_passes,
                                uint32_t* cta_pos, cudaStream_t s) {
  using SL = StagSmem<B, PERSIST>;
  static_assert(SL::total <= 232448, "staggered tower kernel exceeds 227 KB of shared memory");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tower_stag_kernel<B, PERSIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::total);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = n_max < num_sms ? n_max : num_sms;  // see get_pass: up to one game per CTA in a ragged wave
  if (grid <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((grid + 1) & ~1));  // whole CTA pairs
  cfg.blockDim = dim3(kStagThreads);
  cfg.dynamicSmemBytes = SL::total;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, tower_stag_kernel<B, PERSIST>, w, in, q, n_max, policy, value, P, rounds, free_passes,
                            cta_pos);
}

}  // namespace

cudaError_t launch_tower_stag(const TowerWeights& w, int B, const LeafIn* in, const NNQueue& q, int n_max,
                              float* policy, float* value, int num_sms, cudaStream_t s) {
  if (w.n_layers > kMaxLayers) return cudaErrorInvalidValue;
  TreeParams none;
  memset(&none, 0, sizeof none);
  if (B == 9) return launch_tower_stag_t<9, false>(w, in, q, n_max, policy, value, num_sms, none, 1, 0, nullptr, s);
  if (B == 15) return launch_tower_stag_t<15, false>(w, in, q, n_max, policy, value, num_sms, none, 1, 0, nullptr, s);
  return cudaErrorInvalidValue;
}

// The persistent self-play kernel: `rounds` lock-step rounds over the game slots [0, n_games) in one launch.  Needs
// p.static_slots = 1 and every running game in ST_WAIT_NN with its request in p.nn_in[game] (engine.cu enter_persist).
// cta_pos: device uint32[>= grid] cycle positions (zeroed when the persistent state is entered).  With at least
// 2 * GPC games per CTA the kernel free-runs (full passes only, rounds * ceil(n / grid) / GPC passes per CTA: every game
// gets at least `rounds` simulations, games of CTAs that own one game fewer a few more); smaller batches keep the
// static game -> pass map.
cudaError_t launch_selfplay_persist(const TowerWeights& w, int B, const TreeParams& p, int n_games, int rounds,
                                    int num_sms, uint32_t* cta_pos, cudaStream_t s) {
  if (w.n_layers > kMaxLayers || !p.static_slots) return cudaErrorInvalidValue;
  const int per_pass = B == 9 ? Geo<9>::GPC : Geo<15>::GPC;
  int grid = n_games < num_sms ? n_games : num_sms;
  grid = (grid + 1) & ~1;
  if ((n_games + per_pass * grid - 1) / (per_pass * grid) + 1 > kMaxPassesPerCta) return cudaErrorInvalidValue;
  const int min_local = n_games / grid, max_local = (n_games + grid - 1) / grid;
  int free_passes = 0;
  if (cta_pos != nullptr && min_local >= 2 * per_pass && max_local <= kMaxPassesPerCta)
    free_passes = (int)(((long long)rounds * max_local + per_pass - 1) / per_pass);
  NNQueue none;
  memset(&none, 0, sizeof none);
  if (B == 9)
    return launch_tower_stag_t<9, true>(w, p.nn_in, none, n_games, p.nn_policy, p.nn_value, num_sms, p, rounds, free_passes, cta_pos, s);
  if (B == 15)
    return launch_tower_stag_t<15, true>(w, p.nn_in, none, n_games, p.nn_policy, p.nn_value, num_sms, p, rounds, free_passes, cta_pos, s);
  return cudaErrorInvalidValue;
}

}  // namespace ao
