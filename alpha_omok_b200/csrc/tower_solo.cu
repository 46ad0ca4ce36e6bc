// PVNet.forward + the MCTS tree step for a HANDFUL of games (the reference's own operating point: ZeroAgent.get_pi on
// one game, agents.py:73-111, one leaf per network call): the latency-bound end of the path.
//
// tower_stag.cu packs 6 games into one CTA pair's M = 256 tile; with one game that tile is 2/3 empty and one thread
// issues 72 dependent M256 N128 K16 MMAs per layer (~108 cycles each): 102-120 us per simulation.  Here ONE game is spread
// over a CLUSTER OF FOUR CTAs along the output channels instead:
//
//   * CTA r of the cluster owns output channels [32 r, 32 r + 32): tcgen05 cta_group::1, M = 128 (one tile of board
//     rows; 15x15: two tiles), N = 32, and streams only its quarter of the weights (72 KB per layer, ring of 24 KB groups
//     of three taps, 4-5 groups deep).  Every CTA holds the full fp16 activations (all 128 input channels) of the game.
//   * the K dimension is split over FOUR issuing warps with their own fp32 accumulators in TMEM: issuer i takes input
//     channels [32 i, 32 i + 32) (k-steps 2i, 2i + 1 of every tap) = exactly the channels CTA i of the cluster produces, so
//     it can start as soon as CTA i's epilogue has delivered them.  18 MMAs per issuer and layer instead of 72 in a row.
//   * epilogue (8 warps): sum the four accumulators in a fixed order, + bias, ReLU (conv2: the fp32 block input sits in
//     issuer 0's accumulator, as in the other tower kernels), convert to fp16 and store the CTA's 32 channels into the
//     activation buffer of ALL FOUR CTAs with st.async (DSMEM) that completes bytes on the destination's "channel group r
//     has landed" barrier: no fence and no arrive on the writer's side, the consumer's barrier counts the bytes.
//     The buffer is single: before a CTA overwrites rows that the others may still be reading, the layer has to be
//     accumulated everywhere - every issuer's final tcgen05.commit of a layer is multicast to all four CTAs.
//   * heads: the 1x1 head convolutions are partial sums over each CTA's channels, delivered to every CTA in fixed slots
//     (deterministic order of the float adds); each CTA computes a quarter of the policy FC outputs / value hidden units
//     (9x9: from FC weight slices in its shared memory) and sends them to the leader, whose warp 0 does softmax / tanh,
//     runs tree_step_game (expand, backup, move, select) for the game and publishes the next request to the cluster.
//     All `rounds` simulations of a call are one launch (like tower_stag's PERSIST mode).
//   * X3 (9x9): the hi/lo split-precision mode for trained nets, see tower_solo_kernel below.
//
// What bounds it: shared-memory operand fetch of the MMAs (tools/probe_mma_small_n.py: 48 cycles per M128 N32 K16 MMA
// against a 16-cycle math floor - every CTA reads all activation rows of every tap), DESIGN.md 4.0.
//
// Results: same arithmetic per MMA, but the fp32 accumulation is grouped differently (4 partial sums per output), so a
// leaf's floats differ from the batch kernels' in the last bits - both are within the 1e-4 contract of model.PVNet
// (tests/test_gpu_parity.py), and a search replayed through the logged network outputs is bit-exact either way.
#include <cuda_fp16.h>
#include <stdio.h>
#include <string.h>

#include "tower_common.cuh"
#include "tree_device.cuh"

namespace ao {
namespace {

constexpr int kSoloNC = 4;                         // CTAs per game
constexpr int kSoloN = kC / kSoloNC;               // output channels per CTA
constexpr int kSoloIssuers = 4;                    // MMA-issuing warps per CTA = K groups of 32 input channels
constexpr int kSoloEpiWarps = 8;
constexpr int kSoloThreads = (kSoloEpiWarps + kSoloIssuers + 1) * 32;  // + weight producer
constexpr int kSoloGroupTaps = 3;
constexpr int kSoloParts = 2 * kSoloNC;            // partial head sums per position: (CTA, column half)

template <int B, bool X3>
struct SoloGeo {
  static constexpr int A = B * B;
  static constexpr int NT = (A + kTileRows - 1) / kTileRows;       // 128-row tiles of the one board
  static constexpr int Halo = ((B + 1 + 7) / 8) * 8;
  static constexpr int Rows = Halo + NT * kTileRows + Halo;
  static constexpr int ActBytes = 16 * Rows * 16;                   // one operand buffer (X3: a hi and a lo one)
  static constexpr int ActTotal = (X3 ? 2 : 1) * ActBytes;
  // TMEM columns per tile: S (fp32 block input; single pass: issuer 0's accumulator on even layers), then one block per
  // issuer: A_i (32 columns); X3: [hi | lo] (64 columns: a_hi*w_hi, and a_hi*w_lo + a_lo*w_hi)
  static constexpr int AccCols = X3 ? 2 * kSoloN : kSoloN;
  static constexpr int TileCols = kSoloN + kSoloIssuers * AccCols;
  static constexpr int TmemCols = NT * TileCols <= 256 ? 256 : 512;
  static constexpr int BRows = X3 ? 2 * kSoloN : kSoloN;            // rows of a tap's B image: X3 = [32 x w_hi ; 32 x w_lo]
  static constexpr int TapBytes = (kC / 8) * BRows * 16, StemTapBytes = 2 * BRows * 16;
  static constexpr int GroupBytes = kSoloGroupTaps * TapBytes;      // 24 KB (X3: 48 KB)
  static constexpr int RingGroups = X3 ? 2 : (B <= 9 ? 5 : 4);
  static constexpr int APad = (A + 7) / 8 * 8;
  // heads: every CTA computes a quarter of the policy FC outputs and of the value FC1 hidden units
  static constexpr bool FcSmem = B <= 9;                            // its slices of the FC weights live in shared memory
  static constexpr int OQ = (A + kSoloNC - 1) / kSoloNC;            // policy outputs per CTA
  static constexpr int KSP = 256 / OQ;                              // K splits of the policy FC over the 256 head threads
  static constexpr int KLP = (2 * A + KSP - 1) / KSP;
  static constexpr int VS = 256 / kSoloN;                           // K splits of the value FC1 (32 hidden units per CTA)
  static constexpr int KLV = (A + VS - 1) / VS;
  static constexpr uint32_t ActTx = (uint32_t)A * 2u * 2u * 16u * (X3 ? 2u : 1u);  // bytes one CTA's epilogue delivers per layer and destination
  static constexpr uint32_t FeatTx = (uint32_t)A * kSoloParts * 3u * 4u;
  static constexpr uint32_t FcTx = (uint32_t)(A + kC) * 4u;
  static_assert(NT * TileCols <= 512, "TMEM");
  static_assert(NT <= 2, "board too large");
  static_assert(!X3 || NT == 1, "the split-precision variant holds two operand buffers: 9x9 only");
  static_assert(KSP >= 1 && KSP * OQ <= 256, "policy FC thread map");
};

template <int B, bool X3>
struct SoloSmem {
  using G = SoloGeo<B, X3>;
  static constexpr int act = 0;
  static constexpr int wring = G::ActTotal;
  static constexpr int bias = wring + G::RingGroups * G::GroupBytes;        // [kMaxLayers][32] f32 (this CTA's channels)
  static constexpr int headw = bias + kMaxLayers * kSoloN * 4;              // [3][32] f32
  static constexpr int fcb = headw + 3 * kSoloN * 4;                        // pfc_b slice [OQ] | vfc1_b [32] | vfc2_w [32] | head_b [4]
  static constexpr int pfcw = fcb + (G::OQ + 2 * kSoloN + 4) * 4;           // FcSmem: [2A][OQ] f32
  static constexpr int vfcw = pfcw + (G::FcSmem ? 2 * G::A * G::OQ * 4 : 0);   // FcSmem: [A][32] f32
  static constexpr int featp = vfcw + (G::FcSmem ? G::A * kSoloN * 4 : 0);  // [kSoloParts][3][A] f32 partial 1x1-conv sums
  static constexpr int feat = featp + kSoloParts * 3 * G::A * 4;            // [3][A]
  static constexpr int fcpart = feat + 3 * G::A * 4;                        // [KSP][OQ]
  static constexpr int hpart = fcpart + G::KSP * G::OQ * 4;                 // [VS][32]
  static constexpr int logits = hpart + G::VS * kSoloN * 4;                 // leader: [A]
  static constexpr int hidden = logits + G::A * 4;                          // leader: [128]
  static constexpr int pol = (hidden + kC * 4 + 15) / 16 * 16;              // leader: [APad] f32
  static constexpr int tree = (pol + G::APad * 4 + 15) / 16 * 16;           // dbuf f64[APad] | dbuf2 f64[APad] | order | table | rows
  static constexpr int kTreeBytes = 16 * G::APad + 256 + 256 + 128;
  static constexpr int masks = (tree + kTreeBytes + 15) / 16 * 16;          // [NT][9][4] u32
  static constexpr int bars = masks + G::NT * 9 * 4 * 4;
  static constexpr int kBars = 2 * G::RingGroups + kSoloIssuers + 6;
  static constexpr int total = bars + kBars * 8 + 32;  // + TMEM base address, probe scratch
  static_assert(G::ActBytes % 1024 == 0, "weight ring alignment");
  static_assert(tree % 8 == 0, "tree scratch holds doubles");
};

// ---- PTX not in sm100_ptx.cuh: 16-column TMEM access, DSMEM stores, multicast commit of cta_group::1 MMAs
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
// Asynchronous store into the shared memory of a CTA of the cluster (own CTA included) that completes `bytes` on an
// mbarrier of the SAME destination CTA when the data has landed: the writer needs no fence and no separate arrive (a
// release fence at cluster scope costs a MEMBAR.ALL.GPU that waits for every outstanding remote store), the consumer
// waits for the barrier's transaction count.
__device__ __forceinline__ void st_async_v4(uint32_t caddr, const uint4& v, uint32_t cbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(caddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cbar)
               : "memory");
}
__device__ __forceinline__ void st_async_f32(uint32_t caddr, float v, uint32_t cbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(caddr),
               "r"(__float_as_uint(v)), "r"(cbar)
               : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `mask` once all MMAs issued so far by this thread are complete
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void solo_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// Waits that span the leader's heads + tree step (~10 us): back off between probes so that a dozen polling warps do not
// take issue slots from the ONE warp the whole cluster is waiting for.
__device__ __forceinline__ void solo_wait_sleep_cluster(uint64_t* bar, uint32_t parity, unsigned ns) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    __nanosleep(ns);
    if ((++spins & 0xFFu) == 0u && globaltimer_ns() - t0 > 4000000000ull) __trap();
  }
}

// One cluster of four CTAs per game slot [0, n_games): `rounds` simulations (network evaluation of the pending request
// + tree step) in one launch.  Needs P.static_slots (request of game g in P.nn_in[g]) and every running game waiting
// for its answer, exactly like the persistent kernel of tower_stag.cu.
// X3 = the hi/lo split-precision mode (trained nets, cf. tower.cu): activations and weights as fp16 hi + lo parts, every
// k-step a_hi*w_hi + a_hi*w_lo + a_lo*w_hi.  The MMAs are bound by operand fetch, and the first two terms share their A
// tile: a tap's B image is [32 rows w_hi ; 32 rows w_lo] (TowerWeights.conv_quad_x3), so ONE N = 64 instruction computes
// a_hi*w_hi into the hi columns and a_hi*w_lo into the lo columns of the issuer's accumulator block (A fetched once),
// and a second N = 32 instruction adds a_lo*w_hi (the first 32 rows of the same image) onto the lo columns: 144 instead
// of 216 MMAs per layer, 10 KB instead of 15 KB of operands per k-step.  The small lo terms have columns of their own, so
// tcgen05's truncating fp32 accumulation never adds them to a large accumulator (tower.cu orders them first instead);
// hi + lo, the four issuers' partial sums and the residual x (stashed in S) are round-to-nearest adds in the epilogue.
template <int B, bool X3>
__global__ void __cluster_dims__(kSoloNC, 1, 1) __launch_bounds__(kSoloThreads, 1)
tower_solo_kernel(TowerWeights W, TreeParams P, int rounds) {
  using G = SoloGeo<B, X3>;
  using SL = SoloSmem<B, X3>;
  constexpr int RG = G::RingGroups;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_act = smem + SL::act;
  uint8_t* s_w = smem + SL::wring;
  float* s_bias = reinterpret_cast<float*>(smem + SL::bias);
  float* s_headw = reinterpret_cast<float*>(smem + SL::headw);
  float* s_fcb = reinterpret_cast<float*>(smem + SL::fcb);
  float* s_pfcw = reinterpret_cast<float*>(smem + SL::pfcw);
  float* s_vfcw = reinterpret_cast<float*>(smem + SL::vfcw);
  float* s_featp = reinterpret_cast<float*>(smem + SL::featp);
  float* s_feat = reinterpret_cast<float*>(smem + SL::feat);
  float* s_fcpart = reinterpret_cast<float*>(smem + SL::fcpart);
  float* s_hpart = reinterpret_cast<float*>(smem + SL::hpart);
  float* s_logits = reinterpret_cast<float*>(smem + SL::logits);
  float* s_hidden = reinterpret_cast<float*>(smem + SL::hidden);
  float* s_pol = reinterpret_cast<float*>(smem + SL::pol);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + SL::masks);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + SL::bars);  // [RG] producer -> issuers: the group's taps landed
  uint64_t* bar_empty = bar_full + RG;                                // [RG] issuers -> producer
  uint64_t* bar_act = bar_empty + RG;              // [4] tx: input channels [32 i, 32 i + 32) of the layer have landed in MY buffer
  uint64_t* bar_in = bar_act + kSoloIssuers;       // my epilogue warps have written the request's input planes
  uint64_t* bar_acc = bar_in + 1;                  // the layer is accumulated in ALL CTAs (16 multicast commits)
  uint64_t* bar_tfree = bar_acc + 1;               // my epilogue warps have read the accumulators of the previous layer
  uint64_t* bar_feat = bar_tfree + 1;              // tx: the head partial sums of all CTAs have landed in MY featp
  uint64_t* bar_fc = bar_feat + 1;                 // leader, tx: logits / value hidden units of all CTAs have landed
  uint64_t* bar_req = bar_fc + 1;                  // the leader's tree step has published the next request
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_req + 1);
  const uint32_t rank = cluster_ctarank();
  const int game = (int)blockIdx.x / kSoloNC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_layers = W.n_layers;
  const LeafIn* __restrict__ req = P.nn_in + game;

  // ---------------- one-time setup
  for (int i = tid; i < G::ActTotal / 16; i += kSoloThreads) reinterpret_cast<uint4*>(s_act)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < n_layers * kSoloN; i += kSoloThreads)
    s_bias[i] = W.bias[(i / kSoloN) * kC + (int)rank * kSoloN + i % kSoloN];
  for (int i = tid; i < 3 * kSoloN; i += kSoloThreads) s_headw[i] = W.head_w[(i / kSoloN) * kC + (int)rank * kSoloN + i % kSoloN];
  for (int i = tid; i < G::OQ; i += kSoloThreads) s_fcb[i] = (int)rank * G::OQ + i < G::A ? W.pfc_b[(int)rank * G::OQ + i] : 0.f;
  for (int i = tid; i < kSoloN; i += kSoloThreads) {
    s_fcb[G::OQ + i] = W.vfc1_b[(int)rank * kSoloN + i];
    s_fcb[G::OQ + kSoloN + i] = W.vfc2_w[(int)rank * kSoloN + i];
  }
  if (tid < 3) s_fcb[G::OQ + 2 * kSoloN + tid] = W.head_b[tid];
  if (G::FcSmem) {
    for (int i = tid; i < 2 * G::A * G::OQ; i += kSoloThreads) {
      const int kk = i / G::OQ, o = (int)rank * G::OQ + i % G::OQ;
      s_pfcw[i] = o < G::A ? W.pfc_wT[(size_t)kk * G::A + o] : 0.f;
    }
    for (int i = tid; i < G::A * kSoloN; i += kSoloThreads)
      s_vfcw[i] = W.vfc1_wT[(size_t)(i / kSoloN) * kC + (int)rank * kSoloN + i % kSoloN];
  }
  for (int i = tid; i < G::NT * 9 * 4; i += kSoloThreads) {
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
    for (int s = 0; s < RG; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], kSoloIssuers);
    }
    // remote channel groups arrive as st.async bytes (count 1 = the consumer's arming arrive); the CTA's OWN group is
    // written with plain stores and signalled by its 8 epilogue warps - no trip through the async-store path
    for (int i = 0; i < kSoloIssuers; ++i) mbar_init(&bar_act[i], i == (int)rank ? kSoloEpiWarps : 1);
    mbar_init(bar_in, kSoloEpiWarps);
    mbar_init(bar_acc, kSoloIssuers * kSoloNC);
    mbar_init(bar_tfree, kSoloEpiWarps);
    mbar_init(bar_feat, 1);
    mbar_init(bar_fc, 1);
    mbar_init(bar_req, 1);
    fence_mbar_init();
  }
  if (warp == kSoloEpiWarps) tmem_alloc<G::TmemCols>(s_tmem);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  if (warp == kSoloEpiWarps + kSoloIssuers) {
    // =========================================================== weight producer: this CTA's quarter, groups of 3 taps
    if (lane == 0) {
      uint32_t gc = 0;
      for (int rd = 0; rd < rounds; ++rd) {
        size_t off = 0;
        for (int l = 0; l < n_layers; ++l) {
          const uint32_t tapb = l == 0 ? (uint32_t)G::StemTapBytes : (uint32_t)G::TapBytes;
          const uint8_t* base = reinterpret_cast<const uint8_t*>(X3 ? W.conv_quad_x3 : W.conv_quad) + off + (size_t)rank * 9u * tapb;
          for (int g = 0; g < 3; ++g, ++gc) {
            const uint32_t slot = gc % (uint32_t)RG, ph = (gc / (uint32_t)RG) & 1u;
            mbar_wait_sleep(&bar_empty[slot], ph ^ 1u, 100);  // a prefetch: never latency-critical
            mbar_arrive_expect_tx(&bar_full[slot], kSoloGroupTaps * tapb);
            bulk_g2s(s_w + slot * G::GroupBytes, base + (size_t)g * kSoloGroupTaps * tapb, kSoloGroupTaps * tapb, &bar_full[slot]);
          }
          off += (size_t)kSoloNC * 9u * tapb;
        }
      }
    }
  } else if (warp >= kSoloEpiWarps) {
    // =========================================================== MMA issuer i: input channels [32 i, 32 i + 32)
    const int i = warp - kSoloEpiWarps;
    const uint32_t idesc = umma_idesc_f16_f32(kTileRows, G::BRows);   // X3: N = 64 = [hi | lo] columns
    const uint32_t idesc_lo = umma_idesc_f16_f32(kTileRows, kSoloN);  // X3: a_lo * w_hi onto the lo columns
    const uint32_t lbo_a = (uint32_t)G::Rows * 16u;
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(s_act), lbo_a);
    const uint32_t desc_hi = umma_desc_hi(128u);
    constexpr uint32_t kAStep = (2u * (uint32_t)G::Rows * 16u) >> 4;
    constexpr uint32_t kBStep = (2u * (uint32_t)G::BRows * 16u) >> 4;
    uint32_t lc = 0, gc = 0, act_ph = 0;
    constexpr uint32_t kALoOff = (uint32_t)G::ActBytes >> 4;   // X3: the lo operand buffer sits above the hi one
    const bool own_group = i == (int)rank;
    if (lane == 0 && !own_group) mbar_arrive_expect_tx(&bar_act[i], G::ActTx);  // arm the first delivery of CTA i's channels
    for (int rd = 0; rd < rounds; ++rd)
      for (int l = 0; l < n_layers; ++l, ++lc) {
        const bool to_s = !X3 && (l & 1) == 0;  // stem and conv2: issuer 0 accumulates in S (conv2: onto the block input x)
        const bool residual = to_s && l > 0;
        const int nk = l == 0 ? 1 : kC / 16;
        const uint32_t tapb = l == 0 ? (uint32_t)G::StemTapBytes : (uint32_t)G::TapBytes;
        mbar_wait(bar_tfree, lc & 1u);
        AO_DBG(const long long dbg_i0 = (W.dbg && blockIdx.x == 0 && i == 0) ? clock64() : 0;)
        if (l == 0) {
          mbar_wait_sleep(bar_in, (uint32_t)rd & 1u, 50);  // spans the heads + tree step of the previous round
        } else {
          mbar_wait_cluster(&bar_act[i], act_ph);
          act_ph ^= 1u;
          if (lane == 0 && !own_group) mbar_arrive_expect_tx(&bar_act[i], G::ActTx);  // arm the next delivery
        }
        fence_proxy_async_smem();  // rows written through the generic proxy (acquired above) -> my MMAs' operand reads
        tc_fence_after_sync();
        AO_DBG(if (W.dbg && blockIdx.x == 0 && i == 0 && lane == 0) {
          const long long t = clock64();
          if (l > 0) atomicAdd(&W.dbg[7], (unsigned long long)(t - dbg_i0));   // layers 1..: delivery of CTA i's channels
          *reinterpret_cast<volatile long long*>(s_tmem + 2) = t;
        })
        const uint32_t acc_col = X3 ? (uint32_t)(kSoloN + i * G::AccCols)
                                    : (i == 0 ? (to_s ? 0u : (uint32_t)kSoloN) : (uint32_t)(kSoloN * (1 + i)));
        for (int g = 0; g < 3; ++g, ++gc) {
          const uint32_t slot = gc % (uint32_t)RG, ph = (gc / (uint32_t)RG) & 1u;
          AO_DBG(const long long dbg_f0 = (W.dbg && blockIdx.x == 0 && i == 0 && lane == 0) ? clock64() : 0;)
          mbar_wait(&bar_full[slot], ph);
          tc_fence_after_sync();
          AO_DBG(if (W.dbg && blockIdx.x == 0 && i == 0 && lane == 0) atomicAdd(&g_tree_dbg[11], (unsigned long long)(clock64() - dbg_f0));)
          if (elect_one()) {
#pragma unroll
            for (int tile = 0; tile < G::NT; ++tile) {
#pragma unroll
              for (int tt = 0; tt < kSoloGroupTaps; ++tt) {
                const int st = g * kSoloGroupTaps + tt;
                const int t = st == 0 ? 4 : (st <= 4 ? st - 1 : st);  // packed order: centre, 4 negative, 4 positive shifts
                const int shift = (t / 3 - 1) * B + (t % 3 - 1);
                const uint32_t* mk = s_mask + (tile * 9 + t) * 4;
                const uint32_t m0 = mk[0], m1 = mk[1], m2 = mk[2], m3 = mk[3];
                const uint32_t b_lo0 = umma_desc_lo(smem_u32(s_w + slot * G::GroupBytes + (uint32_t)tt * tapb), (uint32_t)G::BRows * 16u);
                const uint32_t a_lo = a_lo0 + (uint32_t)(G::Halo + tile * kTileRows + shift);
                const uint32_t d_tmem = tmem + (uint32_t)(tile * G::TileCols) + acc_col;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  const int j = 2 * i + kk;
                  if (j < nk) {
                    // an accumulator's first MMA of a layer is (centre tap, its first k-step): it writes every row
                    const uint32_t acc = (st == 0 && kk == 0) ? ((i == 0 && residual) ? 1u : 0u) : 1u;
                    umma_f16_ss_lohi_masked(d_tmem, a_lo + (uint32_t)j * kAStep, b_lo0 + (uint32_t)j * kBStep, desc_hi, idesc,
                                            acc, m0, m1, m2, m3);
                    if (X3 && l > 0)  // a_lo * w_hi (rows 0..31 of the image) onto the lo columns; stem: a_lo == 0
                      umma_f16_ss_lohi_masked(d_tmem + (uint32_t)kSoloN, a_lo + kALoOff + (uint32_t)j * kAStep,
                                              b_lo0 + (uint32_t)j * kBStep, desc_hi, idesc_lo, 1u, m0, m1, m2, m3);
                  }
                }
              }
            }
            umma_commit(&bar_empty[slot]);
            if (g == 2) umma_commit_multicast(bar_acc, (uint16_t)((1u << kSoloNC) - 1u));
          }
          __syncwarp();
        }
      }
  } else {
    // =========================================================== epilogue warps; after the tower: heads, tree step
    const int q = warp & 3, half = warp >> 2;   // TMEM lane quarter, half of this CTA's 32 accumulator columns
    const int r = q * 32 + lane;
    const uint32_t chunk_stride = (uint32_t)G::Rows * 16u;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t peer_act[kSoloNC], peer_bar_act[kSoloNC], peer_featp[kSoloNC], peer_bar_feat[kSoloNC];
#pragma unroll
    for (int c = 0; c < kSoloNC; ++c) {
      peer_act[c] = mapa_cluster(smem_u32(s_act), (uint32_t)c);
      peer_bar_act[c] = mapa_cluster(smem_u32(&bar_act[rank]), (uint32_t)c);
      peer_featp[c] = mapa_cluster(smem_u32(s_featp), (uint32_t)c);
      peer_bar_feat[c] = mapa_cluster(smem_u32(bar_feat), (uint32_t)c);
    }
    const uint32_t leader_logits = mapa_cluster(smem_u32(s_logits), 0u);
    const uint32_t leader_hidden = mapa_cluster(smem_u32(s_hidden), 0u);
    const uint32_t leader_bar_fc = mapa_cluster(smem_u32(bar_fc), 0u);
    uint32_t lc = 0;
    AO_DBG(const bool dbg_on = W.dbg != nullptr && blockIdx.x == 0 && tid == 0; long long dbg_t3 = 0;
           if (dbg_on) { g_tree_dbg_on = 1; for (int k = 0; k < 12; ++k) g_tree_dbg[k] = 0ull; } __syncwarp();)
    // "my accumulator reads are done" for the very first layer of the launch: nothing was read yet
    if (lane == 0) mbar_arrive(bar_tfree);

    for (int rd = 0; rd < rounds; ++rd) {
      if (rd > 0) solo_wait_sleep_cluster(bar_req, (uint32_t)(rd - 1) & 1u, 50);
      AO_DBG(const long long dbg_t0 = dbg_on ? clock64() : 0; if (dbg_on && rd > 0) atomicAdd(&W.dbg[4], (unsigned long long)(dbg_t0 - dbg_t3));)
      if (tid == 0) {  // arm this round's deliveries of the heads
        mbar_arrive_expect_tx(bar_feat, G::FeatTx);
        if (rank == 0) mbar_arrive_expect_tx(bar_fc, G::FcTx);
      }
      // ---- the five input planes of the request (utils.get_state_pt as row bit-masks) -> k-chunks 0 and 1 of my buffer
      {
        const int R_in = (G::NT == 2 ? half * kTileRows : 0) + r;
        if ((G::NT == 2 || half == 0) && R_in < G::A) {
          const int yy = R_in / B, xx = R_in % B;
          uint4 c0 = make_uint4(0, 0, 0, 0);
          const uint32_t b0 = (__ldcg(&req->plane[0][yy]) >> xx) & 1u, b1 = (__ldcg(&req->plane[1][yy]) >> xx) & 1u;
          const uint32_t b2 = (__ldcg(&req->plane[2][yy]) >> xx) & 1u, b3 = (__ldcg(&req->plane[3][yy]) >> xx) & 1u;
          const uint32_t b4 = __ldcg(&req->colour) & 1u;
          c0.x = (b0 ? 0x3C00u : 0u) | (b1 ? 0x3C000000u : 0u);
          c0.y = (b2 ? 0x3C00u : 0u) | (b3 ? 0x3C000000u : 0u);
          c0.z = (b4 ? 0x3C00u : 0u);
          const uint32_t ro = (uint32_t)(G::Halo + R_in) * 16u;
          *reinterpret_cast<uint4*>(s_act + ro) = c0;
          *reinterpret_cast<uint4*>(s_act + chunk_stride + ro) = make_uint4(0, 0, 0, 0);
          if (X3) {  // {0,1} planes are exact in fp16: their low parts are zero
            *reinterpret_cast<uint4*>(s_act + G::ActBytes + ro) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(s_act + G::ActBytes + chunk_stride + ro) = make_uint4(0, 0, 0, 0);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_in);  // release at CTA scope; the issuers fence the proxies after their wait
      }

      for (int l = 0; l < n_layers; ++l, ++lc) {
        const bool stash = (l & 1) == 0;            // stem and conv2 outputs are the next block's fp32 input x
        const bool to_s = !X3 && stash;             // single-pass mode: issuer 0 accumulated in S (onto x for conv2)
        const bool last = l == n_layers - 1;
        const float* bias = s_bias + l * kSoloN + half * 16;
        AO_DBG(const long long dbg_a0 = dbg_on ? clock64() : 0;)
        mbar_wait(bar_acc, lc & 1u);  // arrivals are tcgen05.commits (all CTAs): nothing to acquire but the accumulators
        tc_fence_after_sync();
        AO_DBG(if (dbg_on) {
          const long long t = clock64();
          atomicAdd(&W.dbg[5], (unsigned long long)(t - dbg_a0));
          atomicAdd(&W.dbg[6], (unsigned long long)(t - *reinterpret_cast<volatile long long*>(s_tmem + 2)));
        })
#pragma unroll
        for (int t = 0; t < G::NT; ++t) {
          const int R = t * kTileRows + r;
          const bool valid = R < G::A;
          const uint32_t col0 = lane_base + (uint32_t)(t * G::TileCols + half * 16);
          uint32_t v0[16], v1[16], v2[16], v3[16];
          // single pass: issuer 0's accumulator is S (even layers) or A0; X3: the hi columns of the four issuers' blocks
          const uint32_t blk0 = X3 ? col0 + (uint32_t)kSoloN : col0 + (to_s ? 0u : (uint32_t)kSoloN);
          const uint32_t blk_step = (uint32_t)G::AccCols;
          tmem_ld16(blk0, v0);
          if (l > 0) {  // the stem has a single k-step: issuer 0's accumulator is the whole sum
            tmem_ld16((X3 ? blk0 : col0 + (uint32_t)kSoloN) + blk_step, v1);
            tmem_ld16((X3 ? blk0 : col0 + (uint32_t)kSoloN) + 2u * blk_step, v2);
            tmem_ld16((X3 ? blk0 : col0 + (uint32_t)kSoloN) + 3u * blk_step, v3);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float s = __uint_as_float(v0[j]);
            if (l > 0) s = ((s + __uint_as_float(v1[j])) + __uint_as_float(v2[j])) + __uint_as_float(v3[j]);
            v0[j] = __float_as_uint(s);
          }
          if (X3) {  // + the lo columns (a_hi*w_lo + a_lo*w_hi) of the four blocks, + the fp32 block input on conv2 layers
            uint32_t xr[16];
            tmem_ld16(blk0 + (uint32_t)kSoloN, v1);
            if (l > 0) {
              tmem_ld16(blk0 + (uint32_t)kSoloN + blk_step, v2);
              tmem_ld16(blk0 + (uint32_t)kSoloN + 2u * blk_step, v3);
              tmem_ld16(blk0 + (uint32_t)kSoloN + 3u * blk_step, xr);
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float lo = __uint_as_float(v1[j]);
              if (l > 0) lo = ((lo + __uint_as_float(v2[j])) + __uint_as_float(v3[j])) + __uint_as_float(xr[j]);
              v0[j] = __float_as_uint(__uint_as_float(v0[j]) + lo);
            }
            if (stash && l > 0) {  // out = conv2 + x as a round-to-nearest fp32 add (model.py:29)
              tmem_ld16(col0, xr);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) v0[j] = __float_as_uint(__uint_as_float(v0[j]) + __uint_as_float(xr[j]));
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v0[j] = __float_as_uint(fmaxf(__uint_as_float(v0[j]) + bias[j], 0.f));
          if (!last) {
            uint4 pk[2];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              __half2 h;
              h = __floats2half2_rn(__uint_as_float(v0[cc * 8 + 0]), __uint_as_float(v0[cc * 8 + 1]));
              pk[cc].x = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2half2_rn(__uint_as_float(v0[cc * 8 + 2]), __uint_as_float(v0[cc * 8 + 3]));
              pk[cc].y = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2half2_rn(__uint_as_float(v0[cc * 8 + 4]), __uint_as_float(v0[cc * 8 + 5]));
              pk[cc].z = *reinterpret_cast<uint32_t*>(&h);
              h = __floats2half2_rn(__uint_as_float(v0[cc * 8 + 6]), __uint_as_float(v0[cc * 8 + 7]));
              pk[cc].w = *reinterpret_cast<uint32_t*>(&h);
            }
            if (valid) {  // rows beyond the board stay zero for ever: no on-board tap of a real row reads them
              const uint32_t off0 = (uint32_t)((int)rank * 4 + half * 2) * chunk_stride + (uint32_t)(G::Halo + R) * 16u;
              *reinterpret_cast<uint4*>(s_act + off0) = pk[0];
              *reinterpret_cast<uint4*>(s_act + off0 + chunk_stride) = pk[1];
#pragma unroll
              for (int k = 1; k < kSoloNC; ++k) {  // destinations rotated by rank: no CTA's inbound port is everyone's first
                const int c = ((int)rank + k) % kSoloNC;
                st_async_v4(peer_act[c] + off0, pk[0], peer_bar_act[c]);
                st_async_v4(peer_act[c] + off0 + chunk_stride, pk[1], peer_bar_act[c]);
              }
              if (X3) {  // low parts: fp16(y - fp16(y)), into the lo operand buffers
                uint4 pl[2];
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                  const __half2* hh = reinterpret_cast<const __half2*>(&pk[cc]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(hh[e]);
                    const __half2 l2 = __floats2half2_rn(__uint_as_float(v0[cc * 8 + 2 * e]) - f.x, __uint_as_float(v0[cc * 8 + 2 * e + 1]) - f.y);
                    reinterpret_cast<uint32_t*>(&pl[cc])[e] = *reinterpret_cast<const uint32_t*>(&l2);
                  }
                }
                *reinterpret_cast<uint4*>(s_act + G::ActBytes + off0) = pl[0];
                *reinterpret_cast<uint4*>(s_act + G::ActBytes + off0 + chunk_stride) = pl[1];
#pragma unroll
                for (int k = 1; k < kSoloNC; ++k) {
                  const int c = ((int)rank + k) % kSoloNC;
                  st_async_v4(peer_act[c] + (uint32_t)G::ActBytes + off0, pl[0], peer_bar_act[c]);
                  st_async_v4(peer_act[c] + (uint32_t)G::ActBytes + off0 + chunk_stride, pl[1], peer_bar_act[c]);
                }
              }
            }
            if (stash) tmem_st16(col0, v0);  // fp32 block input for the next residual add
          } else {
            // heads' 1x1 convolutions (model.py:44-46, 64-66) over this thread's 16 channels -> slot (CTA, half) of every
            // CTA's partial-sum table (fixed slots: the order of the float adds does not depend on arrival order)
            const float* hw = s_headw + half * 16;
            float hd0 = 0.f, hd1 = 0.f, hd2 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float x = __uint_as_float(v0[j]);
              hd0 = fmaf(x, hw[0 * kSoloN + j], hd0);
              hd1 = fmaf(x, hw[1 * kSoloN + j], hd1);
              hd2 = fmaf(x, hw[2 * kSoloN + j], hd2);
            }
            if (valid) {
              const uint32_t fo = (uint32_t)((((int)rank * 2 + half) * 3) * G::A + R) * 4u;
#pragma unroll
              for (int c = 0; c < kSoloNC; ++c) {
                st_async_f32(peer_featp[c] + fo, hd0, peer_bar_feat[c]);
                st_async_f32(peer_featp[c] + fo + (uint32_t)G::A * 4u, hd1, peer_bar_feat[c]);
                st_async_f32(peer_featp[c] + fo + 2u * (uint32_t)G::A * 4u, hd2, peer_bar_feat[c]);
              }
            }
          }
        }
        if (!last) {  // my own channel group sits in my buffer (plain stores): tell this CTA's issuer `rank`
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_act[rank]);
        }
        // the issuers may overwrite the accumulators with the next layer (and issuer 0 finds the new block input in S)
        if (!last && stash) tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tfree);
      }

      // =========================================================== heads (model.py:43-50, 63-73): a quarter per CTA
      mbar_wait_cluster(bar_feat, (uint32_t)rd & 1u);
      AO_DBG(const long long dbg_t1 = dbg_on ? clock64() : 0;)
      for (int i = tid; i < 3 * G::A; i += 256) {
        float s = s_featp[i];
#pragma unroll
        for (int pp = 1; pp < kSoloParts; ++pp) s += s_featp[pp * 3 * G::A + i];
        s_feat[i] = fmaxf(s + s_fcb[G::OQ + 2 * kSoloN + i / G::A], 0.f);
      }
      solo_epi_sync();
      if (tid < G::KSP * G::OQ) {  // policy FC outputs [rank OQ, rank OQ + OQ), K split over KSP thread groups
        const int ks = tid / G::OQ, oo = tid % G::OQ;
        const int o = (int)rank * G::OQ + oo;
        const int k0 = ks * G::KLP, k1 = min(2 * G::A, k0 + G::KLP);
        float acc = 0.f;
        if (o < G::A) {
          if (G::FcSmem) {
#pragma unroll 7
            for (int kk = k0; kk < k1; ++kk) acc = fmaf(s_pfcw[kk * G::OQ + oo], s_feat[kk], acc);
          } else {
            const float* wt = W.pfc_wT + o;
#pragma unroll 16
            for (int kk = k0; kk < k1; ++kk) acc = fmaf(__ldg(wt + (size_t)kk * G::A), s_feat[kk], acc);
          }
        }
        s_fcpart[ks * G::OQ + oo] = acc;
      }
      {  // value FC1 hidden units [32 rank, 32 rank + 32), K split over VS thread groups
        const int vs = tid / kSoloN, vj = tid % kSoloN;
        const int k0 = vs * G::KLV, k1 = min(G::A, k0 + G::KLV);
        const float* f = s_feat + 2 * G::A;
        float acc = 0.f;
        if (G::FcSmem) {
#pragma unroll 11
          for (int kk = k0; kk < k1; ++kk) acc = fmaf(s_vfcw[kk * kSoloN + vj], f[kk], acc);
        } else {
          const float* wt = W.vfc1_wT + (int)rank * kSoloN + vj;
#pragma unroll 15
          for (int kk = k0; kk < k1; ++kk) acc = fmaf(__ldg(wt + (size_t)kk * kC), f[kk], acc);
        }
        s_hpart[vs * kSoloN + vj] = acc;
      }
      solo_epi_sync();
      if (tid < G::OQ) {
        if ((int)rank * G::OQ + tid < G::A) {
          float acc = s_fcb[tid];
#pragma unroll
          for (int ks = 0; ks < G::KSP; ++ks) acc += s_fcpart[ks * G::OQ + tid];
          st_async_f32(leader_logits + (uint32_t)((int)rank * G::OQ + tid) * 4u, acc, leader_bar_fc);
        }
      } else if (tid >= 64 && tid < 64 + kSoloN) {
        const int vj = tid - 64;
        float acc = s_fcb[G::OQ + vj];
#pragma unroll
        for (int vs = 0; vs < G::VS; ++vs) acc += s_hpart[vs * kSoloN + vj];
        st_async_f32(leader_hidden + (uint32_t)((int)rank * kSoloN + vj) * 4u, fmaxf(acc, 0.f) * s_fcb[G::OQ + kSoloN + vj], leader_bar_fc);
      }
      if (rank == 0 && warp == 0) {
        // ---- leader: softmax, tanh, then the tree step of the game
        mbar_wait_cluster(bar_fc, (uint32_t)rd & 1u);
        float mx = -3.0e38f;
        for (int kk = lane; kk < G::A; kk += 32) mx = fmaxf(mx, s_logits[kk]);
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        float sum = 0.f;
        for (int kk = lane; kk < G::A; kk += 32) sum += expf(s_logits[kk] - mx);
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        float hv = 0.f;
        for (int kk = lane; kk < kC; kk += 32) hv += s_hidden[kk];
#pragma unroll
        for (int o = 16; o; o >>= 1) hv += __shfl_xor_sync(0xFFFFFFFFu, hv, o);
        for (int kk = lane; kk < G::A; kk += 32) s_pol[kk] = expf(s_logits[kk] - mx) / sum;
        const float v = tanhf(hv + W.vfc2_b);
        __syncwarp();
        AO_DBG(const long long dbg_t2 = dbg_on ? clock64() : 0;)
        // tree_device.cuh: consume the answer, expand, back up, play the move when the search is complete, select the
        // next leaf and write its request to P.nn_in[game]
        uint8_t* ts = smem + SL::tree;
        WarpSmem ws;
        ws.dbuf = reinterpret_cast<double*>(ts);
        ws.dbuf2 = ws.dbuf + G::APad;
        ws.order = reinterpret_cast<uint8_t*>(ws.dbuf2 + G::APad);
        ws.table = reinterpret_cast<int16_t*>(ws.order + 256);
        ws.rows = reinterpret_cast<uint16_t(*)[32]>(ws.table + 128);
        ws.pol = s_pol;
        (void)tree_step_game<(G::A <= 96 ? 3 : 8)>(P, game, &ws, lane, 64, true, v);
        __threadfence();  // the request (global memory) before the arrives below
        __syncwarp();
        AO_DBG(if (dbg_on) {
          dbg_t3 = clock64();
          atomicAdd(&W.dbg[0], 1ull);
          atomicAdd(&W.dbg[1], (unsigned long long)(dbg_t1 - dbg_t0));
          atomicAdd(&W.dbg[2], (unsigned long long)(dbg_t2 - dbg_t1));
          atomicAdd(&W.dbg[3], (unsigned long long)(dbg_t3 - dbg_t2));
        })
        if (lane == 0)
          for (int c = 0; c < kSoloNC; ++c) mbar_arrive_remote_relaxed(bar_req, (uint32_t)c);
      }
    }
  }
  AO_DBG(if (W.dbg != nullptr && blockIdx.x == 0 && tid == 0) {
    g_tree_dbg_on = 0;
    printf("tree step phases, cycles per round: regs+leaf loads %llu | expand: legal order %llu, prior sum %llu, noise + child block %llu, backup %llu"
           " | (expand_and_backup total incl. bookkeeping %llu) | selection walk %llu, win check %llu, terminal backup %llu, request %llu, store regs %llu"
           " || issuer 0 waits for weights %llu\n",
           g_tree_dbg[0] / rounds, g_tree_dbg[2] / rounds, g_tree_dbg[3] / rounds, g_tree_dbg[4] / rounds, g_tree_dbg[5] / rounds,
           g_tree_dbg[1] / rounds, g_tree_dbg[6] / rounds, g_tree_dbg[7] / rounds, g_tree_dbg[8] / rounds, g_tree_dbg[9] / rounds,
           g_tree_dbg[10] / rounds, g_tree_dbg[11] / rounds);
  })
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while another one of the cluster may still signal it or write into its buffers
  if (warp == kSoloEpiWarps) tmem_dealloc<G::TmemCols>(tmem);
}

template <int B, bool X3>
cudaError_t launch_solo_t(const TowerWeights& w, const TreeParams& p, int n_games, int rounds, cudaStream_t s) {
  using SL = SoloSmem<B, X3>;
  static_assert(SL::total <= 232448, "solo tower kernel exceeds 227 KB of shared memory");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tower_solo_kernel<B, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::total);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  tower_solo_kernel<B, X3><<<dim3((unsigned)(n_games * kSoloNC)), dim3(kSoloThreads), SL::total, s>>>(w, p, rounds);
  return cudaGetLastError();
}

}  // namespace

// How many games the solo kernel takes (one cluster of four CTAs each, all clusters resident at once).
int solo_max_games(int num_sms) {
  const int g = (num_sms - 16) / kSoloNC;  // clusters of 4 strand up to 16 of the 148 SMs (GPCs of 16 / 18 / 20 SMs)
  return g < 1 ? 1 : g;
}

// Does the cluster-of-four kernel exist for this board and tower mode?  (split precision: 9x9 only - two operand buffers)
bool solo_supports(int B, int precision) {
  if (precision == AO_NN_FP16) return B == 9 || B == 15;
  if (precision == AO_NN_FP16X3) return B == 9;
  return false;
}

// `rounds` simulations for each of the game slots [0, n_games) in one launch; needs p.static_slots = 1 and every running
// game in ST_WAIT_NN with its request in p.nn_in[game] (engine.cu enter_persist), w.conv_quad (split mode: conv_quad_x3) loaded.
cudaError_t launch_selfplay_solo(const TowerWeights& w, int B, int precision, const TreeParams& p, int n_games, int rounds,
                                 cudaStream_t s) {
  if (w.n_layers > kMaxLayers || !p.static_slots || w.conv_quad == nullptr || n_games < 1 || rounds < 1) return cudaErrorInvalidValue;
  if (!solo_supports(B, precision)) return cudaErrorInvalidValue;
  if (precision == AO_NN_FP16X3) {
    if (w.conv_quad_x3 == nullptr) return cudaErrorInvalidValue;
    return launch_solo_t<9, true>(w, p, n_games, rounds, s);
  }
  if (B == 9) return launch_solo_t<9, false>(w, p, n_games, rounds, s);
  return launch_solo_t<15, false>(w, p, n_games, rounds, s);
}

}  // namespace ao
