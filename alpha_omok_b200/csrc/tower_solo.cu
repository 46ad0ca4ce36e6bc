// PVNet.forward + the MCTS tree step for a HANDFUL of games (the reference's own operating point: ZeroAgent.get_pi on
// one game, agents.py:73-111, one leaf per network call): the latency-bound end of the path.
//
// tower_stag.cu packs 6 games into one CTA pair's M = 256 tile; with one game that tile is 2/3 empty and one thread
// issues 72 dependent M256 N128 K16 MMAs per layer (~108 cycles each): 102-120 us per simulation.  Here ONE game is spread
// over a CLUSTER OF FOUR CTAs along the output channels instead:
//
//   * CTA r of the cluster owns output channels [32 r, 32 r + 32): tcgen05 cta_group::1, M = 128 (one tile of board
//     rows; 15x15: two tiles), N = 32, and streams only its quarter of the weights (72 KB per layer, ring of 24 KB groups
//     of three taps, 1-2 layers deep).  Every CTA holds the full fp16 activations (all 128 input channels) of the game.
//   * the K dimension is split over FOUR issuing warps with their own fp32 accumulators in TMEM: issuer i takes input
//     channels [32 i, 32 i + 32) (k-steps 2i, 2i + 1 of every tap) = exactly the channels CTA i of the cluster produces, so
//     it can start as soon as CTA i's epilogue has delivered them.  18 MMAs per issuer and layer instead of 72 in a row.
//   * epilogue (8 warps): sum the four accumulators in a fixed order, + bias, ReLU (conv2: the fp32 block input sits in
//     issuer 0's accumulator, as in the other tower kernels), convert to fp16 and store the CTA's 32 channels into the
//     activation buffer of ALL FOUR CTAs (st.shared::cluster), then arrive on "channel group r is ready" in each of them.
//     The buffer is single: before a CTA overwrites rows that the others may still be reading, the layer has to be
//     accumulated everywhere - every issuer's final tcgen05.commit of a layer is multicast to all four CTAs.
//   * heads: the 1x1 head convolutions are partial sums over each CTA's channels, written to the leader in fixed slots
//     (deterministic order of the float adds); the leader's epilogue warps then run the FC layers / softmax / tanh and
//     warp 0 runs tree_step_game (expand, backup, move, select) for the game and publishes the next request to the
//     cluster.  All `rounds` simulations of a call are one launch (like tower_stag's PERSIST mode).
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
constexpr int kSoloTapBytes = (kC / 8) * kSoloN * 16;   // [16 k-chunks][32 co][8 halves] = 8 KB
constexpr int kSoloStemTapBytes = 2 * kSoloN * 16;      // stem: K = 16
constexpr int kSoloGroupTaps = 3;
constexpr int kSoloGroupBytes = kSoloGroupTaps * kSoloTapBytes;
constexpr int kSoloParts = 2 * kSoloNC;            // partial head sums per position: (CTA, column half)

template <int B>
struct SoloGeo {
  static constexpr int A = B * B;
  static constexpr int NT = (A + kTileRows - 1) / kTileRows;       // 128-row tiles of the one board
  static constexpr int Halo = ((B + 1 + 7) / 8) * 8;
  static constexpr int Rows = Halo + NT * kTileRows + Halo;
  static constexpr int ActBytes = 16 * Rows * 16;
  static constexpr int TileCols = kSoloN * (1 + kSoloIssuers);      // S (fp32 block input / issuer 0 on even layers), A0..A3
  static constexpr int TmemCols = NT * TileCols <= 256 ? 256 : 512;
  static constexpr int RingGroups = B <= 9 ? 6 : 4;                 // 24 KB each
  static constexpr int APad = (A + 7) / 8 * 8;
  static constexpr int KS = 256 / A >= 1 ? 256 / A : 1;             // K splits of the policy FC over the 256 head threads
  static_assert(NT * TileCols <= 512, "TMEM");
  static_assert(NT <= 2, "board too large");
};

template <int B>
struct SoloSmem {
  using G = SoloGeo<B>;
  static constexpr int act = 0;
  static constexpr int wring = G::ActBytes;
  static constexpr int bias = wring + G::RingGroups * kSoloGroupBytes;      // [kMaxLayers][32] f32 (this CTA's channels)
  static constexpr int headw = bias + kMaxLayers * kSoloN * 4;              // [3][32] f32
  static constexpr int featp = headw + 3 * kSoloN * 4;                      // leader: [kSoloParts][3][A] f32
  static constexpr int feat = featp + kSoloParts * 3 * G::A * 4;            // [3][A]
  static constexpr int fcpart = feat + 3 * G::A * 4;                        // [KS][A]
  static constexpr int hpart = fcpart + G::KS * G::A * 4;                   // [2][128]
  static constexpr int logits = hpart + 2 * kC * 4;                         // [A]
  static constexpr int hidden = logits + G::A * 4;                          // [128]
  static constexpr int pol = (hidden + kC * 4 + 15) / 16 * 16;              // [APad] f32
  static constexpr int val = pol + G::APad * 4;                             // [4] f32
  static constexpr int tree = (val + 16 + 15) / 16 * 16;                    // dbuf f64[APad] | dbuf2 f64[APad] | order | table | rows
  static constexpr int kTreeBytes = 16 * G::APad + 256 + 256 + 128;
  static constexpr int masks = (tree + kTreeBytes + 15) / 16 * 16;          // [NT][9][4] u32
  static constexpr int bars = masks + G::NT * 9 * 4 * 4;
  static constexpr int kBars = 2 * G::RingGroups + kSoloIssuers + 4;
  static constexpr int total = bars + kBars * 8 + 16;
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
__device__ __forceinline__ void st_cluster_v4(uint32_t caddr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t caddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}
// Publishing data to the other CTAs of the cluster: ONE release fence at cluster scope per warp, then relaxed arrives on
// the barriers of all destination CTAs (an arrive.release per destination costs a MEMBAR.ALL.GPU each).
__device__ __forceinline__ void fence_release_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
// arrive on the mbarrier at this offset in every CTA of `mask` once all MMAs issued so far by this thread are complete
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void solo_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// One cluster of four CTAs per game slot [0, n_games): `rounds` simulations (network evaluation of the pending request
// + tree step) in one launch.  Needs P.static_slots (request of game g in P.nn_in[g]) and every running game waiting
// for its answer, exactly like the persistent kernel of tower_stag.cu.
template <int B>
__global__ void __cluster_dims__(kSoloNC, 1, 1) __launch_bounds__(kSoloThreads, 1)
tower_solo_kernel(TowerWeights W, TreeParams P, int rounds) {
  using G = SoloGeo<B>;
  using SL = SoloSmem<B>;
  constexpr int RG = G::RingGroups;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_act = smem + SL::act;
  uint8_t* s_w = smem + SL::wring;
  float* s_bias = reinterpret_cast<float*>(smem + SL::bias);
  float* s_headw = reinterpret_cast<float*>(smem + SL::headw);
  float* s_featp = reinterpret_cast<float*>(smem + SL::featp);
  float* s_feat = reinterpret_cast<float*>(smem + SL::feat);
  float* s_fcpart = reinterpret_cast<float*>(smem + SL::fcpart);
  float* s_hpart = reinterpret_cast<float*>(smem + SL::hpart);
  float* s_logits = reinterpret_cast<float*>(smem + SL::logits);
  float* s_hidden = reinterpret_cast<float*>(smem + SL::hidden);
  float* s_pol = reinterpret_cast<float*>(smem + SL::pol);
  float* s_val = reinterpret_cast<float*>(smem + SL::val);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + SL::masks);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + SL::bars);  // [RG] producer -> issuers: the group's taps landed
  uint64_t* bar_empty = bar_full + RG;                                // [RG] issuers -> producer
  uint64_t* bar_act = bar_empty + RG;              // [4] input channels [32 i, 32 i + 32) of the layer are in MY buffer
  uint64_t* bar_acc = bar_act + kSoloIssuers;      // the layer is accumulated in ALL CTAs (16 multicast commits)
  uint64_t* bar_tfree = bar_acc + 1;               // my epilogue warps have read the accumulators of the previous layer
  uint64_t* bar_feat = bar_tfree + 1;              // leader: the head partial sums of all CTAs have arrived
  uint64_t* bar_req = bar_feat + 1;                // the leader's tree step has published the next request
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_req + 1);
  const uint32_t rank = cluster_ctarank();
  const int game = (int)blockIdx.x / kSoloNC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_layers = W.n_layers;
  const LeafIn* __restrict__ req = P.nn_in + game;

  // ---------------- one-time setup
  for (int i = tid; i < G::ActBytes / 16; i += kSoloThreads) reinterpret_cast<uint4*>(s_act)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < n_layers * kSoloN; i += kSoloThreads)
    s_bias[i] = W.bias[(i / kSoloN) * kC + (int)rank * kSoloN + i % kSoloN];
  for (int i = tid; i < 3 * kSoloN; i += kSoloThreads) s_headw[i] = W.head_w[(i / kSoloN) * kC + (int)rank * kSoloN + i % kSoloN];
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
    for (int i = 0; i < kSoloIssuers; ++i) mbar_init(&bar_act[i], kSoloEpiWarps);
    mbar_init(bar_acc, kSoloIssuers * kSoloNC);
    mbar_init(bar_tfree, kSoloEpiWarps);
    mbar_init(bar_feat, kSoloEpiWarps * kSoloNC);
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
          const uint32_t tapb = l == 0 ? (uint32_t)kSoloStemTapBytes : (uint32_t)kSoloTapBytes;
          const uint8_t* base = reinterpret_cast<const uint8_t*>(W.conv_quad) + off + (size_t)rank * 9u * tapb;
          for (int g = 0; g < 3; ++g, ++gc) {
            const uint32_t slot = gc % (uint32_t)RG, ph = (gc / (uint32_t)RG) & 1u;
            mbar_wait(&bar_empty[slot], ph ^ 1u);
            mbar_arrive_expect_tx(&bar_full[slot], kSoloGroupTaps * tapb);
            bulk_g2s(s_w + slot * kSoloGroupBytes, base + (size_t)g * kSoloGroupTaps * tapb, kSoloGroupTaps * tapb, &bar_full[slot]);
          }
          off += (size_t)kSoloNC * 9u * tapb;
        }
      }
    }
  } else if (warp >= kSoloEpiWarps) {
    // =========================================================== MMA issuer i: input channels [32 i, 32 i + 32)
    const int i = warp - kSoloEpiWarps;
    const uint32_t idesc = umma_idesc_f16_f32(kTileRows, kSoloN);
    const uint32_t lbo_a = (uint32_t)G::Rows * 16u;
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(s_act), lbo_a);
    const uint32_t desc_hi = umma_desc_hi(128u);
    constexpr uint32_t kAStep = (2u * (uint32_t)G::Rows * 16u) >> 4;
    constexpr uint32_t kBStep = (2u * (uint32_t)kSoloN * 16u) >> 4;
    uint32_t lc = 0, gc = 0;
    for (int rd = 0; rd < rounds; ++rd)
      for (int l = 0; l < n_layers; ++l, ++lc) {
        const bool to_s = (l & 1) == 0;       // stem and conv2: issuer 0 accumulates in S (conv2: onto the block input x)
        const bool residual = to_s && l > 0;
        const int nk = l == 0 ? 1 : kC / 16;
        const uint32_t tapb = l == 0 ? (uint32_t)kSoloStemTapBytes : (uint32_t)kSoloTapBytes;
        mbar_wait(bar_tfree, lc & 1u);
        mbar_wait_cluster(&bar_act[i], lc & 1u);
        fence_proxy_async_smem();  // generic-proxy writes of the other CTAs' epilogues (acquired above) -> my MMAs' operand reads
        tc_fence_after_sync();
        const uint32_t acc_col = i == 0 ? (to_s ? 0u : (uint32_t)kSoloN) : (uint32_t)(kSoloN * (1 + i));
        for (int g = 0; g < 3; ++g, ++gc) {
          const uint32_t slot = gc % (uint32_t)RG, ph = (gc / (uint32_t)RG) & 1u;
          mbar_wait(&bar_full[slot], ph);
          tc_fence_after_sync();
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
                const uint32_t b_lo0 = umma_desc_lo(smem_u32(s_w + slot * kSoloGroupBytes + (uint32_t)tt * tapb), (uint32_t)kSoloN * 16u);
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
    // =========================================================== epilogue warps (+ heads and tree step on the leader)
    const int q = warp & 3, half = warp >> 2;   // TMEM lane quarter, half of this CTA's 32 accumulator columns
    const int r = q * 32 + lane;
    const uint32_t chunk_stride = (uint32_t)G::Rows * 16u;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t act_addr = smem_u32(s_act);
    uint32_t peer_act[kSoloNC];
#pragma unroll
    for (int c = 0; c < kSoloNC; ++c) peer_act[c] = mapa_cluster(act_addr, (uint32_t)c);
    const uint32_t featp_leader = mapa_cluster(smem_u32(s_featp), 0u);
    uint32_t lc = 0;
    AO_DBG(const bool dbg_on = W.dbg != nullptr && blockIdx.x == 0 && tid == 0; long long dbg_t3 = 0;)
    // "my accumulator reads are done" for the very first layer of the launch: nothing was read yet
    if (lane == 0) mbar_arrive(bar_tfree);

    for (int rd = 0; rd < rounds; ++rd) {
      if (rd > 0) mbar_wait_cluster(bar_req, (uint32_t)(rd - 1) & 1u);
      AO_DBG(const long long dbg_t0 = dbg_on ? clock64() : 0; if (dbg_on && rd > 0) atomicAdd(&W.dbg[4], (unsigned long long)(dbg_t0 - dbg_t3));)
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
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0)
          for (int c = 0; c < kSoloIssuers; ++c) mbar_arrive(&bar_act[c]);
      }

      for (int l = 0; l < n_layers; ++l, ++lc) {
        const bool to_s = (l & 1) == 0;
        const bool last = l == n_layers - 1;
        const float* bias = s_bias + l * kSoloN + half * 16;
        AO_DBG(const long long dbg_a0 = dbg_on ? clock64() : 0;)
        mbar_wait(bar_acc, lc & 1u);  // arrivals are tcgen05.commits (all CTAs): nothing to acquire but the accumulators
        tc_fence_after_sync();
        AO_DBG(if (dbg_on) atomicAdd(&W.dbg[5], (unsigned long long)(clock64() - dbg_a0));)
        float hd0 = 0.f, hd1 = 0.f, hd2 = 0.f;
#pragma unroll
        for (int t = 0; t < G::NT; ++t) {
          const int R = t * kTileRows + r;
          const bool valid = R < G::A;
          const uint32_t col0 = lane_base + (uint32_t)(t * G::TileCols + half * 16);
          uint32_t v0[16], v1[16], v2[16], v3[16];
          tmem_ld16(col0 + (to_s ? 0u : (uint32_t)kSoloN), v0);
          if (l > 0) {  // the stem has a single k-step: issuer 0's accumulator is the whole sum
            tmem_ld16(col0 + 2u * kSoloN, v1);
            tmem_ld16(col0 + 3u * kSoloN, v2);
            tmem_ld16(col0 + 4u * kSoloN, v3);
          }
          tmem_ld_wait();
          if (t == G::NT - 1) {  // the issuers may overwrite the accumulators with the next layer
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tfree);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float s = __uint_as_float(v0[j]);
            if (l > 0) s = ((s + __uint_as_float(v1[j])) + __uint_as_float(v2[j])) + __uint_as_float(v3[j]);
            v0[j] = __float_as_uint(fmaxf(s + bias[j], 0.f));
          }
          if (!last) {
            if (to_s) tmem_st16(col0, v0);  // fp32 block input for the next residual add
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
#pragma unroll
              for (int c = 0; c < kSoloNC; ++c) {
                if (c == (int)rank) continue;
                st_cluster_v4(peer_act[c] + off0, pk[0]);
                st_cluster_v4(peer_act[c] + off0 + chunk_stride, pk[1]);
              }
              *reinterpret_cast<uint4*>(s_act + off0) = pk[0];
              *reinterpret_cast<uint4*>(s_act + off0 + chunk_stride) = pk[1];
            }
          } else {
            // heads' 1x1 convolutions (model.py:44-46, 64-66) over this thread's 16 channels
            const float* hw = s_headw + half * 16;
            hd0 = hd1 = hd2 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float x = __uint_as_float(v0[j]);
              hd0 = fmaf(x, hw[0 * kSoloN + j], hd0);
              hd1 = fmaf(x, hw[1 * kSoloN + j], hd1);
              hd2 = fmaf(x, hw[2 * kSoloN + j], hd2);
            }
            if (valid) {
              const uint32_t fo = featp_leader + (uint32_t)((((int)rank * 2 + half) * 3) * G::A + R) * 4u;
              st_cluster_f32(fo, hd0);
              st_cluster_f32(fo + (uint32_t)G::A * 4u, hd1);
              st_cluster_f32(fo + 2u * (uint32_t)G::A * 4u, hd2);
            }
          }
        }
        if (!last) {
          if (to_s) tmem_st_wait();
          // the rows were written through the generic proxy (DSMEM stores); the consuming issuer orders them against its
          // MMAs' operand reads with a proxy fence after its acquire
          tc_fence_before_sync();
          fence_release_cluster();
          __syncwarp();
          if (lane == 0)
            for (int c = 0; c < kSoloNC; ++c) mbar_arrive_remote_relaxed(&bar_act[rank], (uint32_t)c);
        } else {
          fence_release_cluster();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote_relaxed(bar_feat, 0u);
        }
      }

      if (rank == 0) {
        // =========================================================== heads (model.py:43-50, 63-73) + tree step, leader only
        mbar_wait_cluster(bar_feat, (uint32_t)rd & 1u);
        AO_DBG(const long long dbg_t1 = dbg_on ? clock64() : 0;)
        for (int i = tid; i < 3 * G::A; i += 256) {
          float s = s_featp[i];
#pragma unroll
          for (int pp = 1; pp < kSoloParts; ++pp) s += s_featp[pp * 3 * G::A + i];
          s_feat[i] = fmaxf(s + W.head_b[i / G::A], 0.f);
        }
        solo_epi_sync();
        if (tid < G::KS * G::A) {  // policy FC, K split over KS thread groups
          constexpr int KL = (2 * G::A + G::KS - 1) / G::KS;
          const int ks = tid / G::A, po = tid % G::A;
          const int k0 = ks * KL, k1 = min(2 * G::A, k0 + KL);
          const float* wt = W.pfc_wT + po;
          float acc = 0.f;
#pragma unroll 18
          for (int kk = k0; kk < k1; ++kk) acc = fmaf(__ldg(wt + (size_t)kk * G::A), s_feat[kk], acc);
          s_fcpart[ks * G::A + po] = acc;
        }
        {  // value FC1, K split in two
          constexpr int KL = (G::A + 1) / 2;
          const int vs = tid >> 7, vj = tid & 127;
          const int k0 = vs * KL, k1 = min(G::A, k0 + KL);
          const float* f = s_feat + 2 * G::A;
          const float* wt = W.vfc1_wT + vj;
          float acc = 0.f;
#pragma unroll 21
          for (int kk = k0; kk < k1; ++kk) acc = fmaf(__ldg(wt + (size_t)kk * kC), f[kk], acc);
          s_hpart[vs * kC + vj] = acc;
        }
        solo_epi_sync();
        if (tid < G::A) {
          float acc = W.pfc_b[tid];
#pragma unroll
          for (int ks = 0; ks < G::KS; ++ks) acc += s_fcpart[ks * G::A + tid];
          s_logits[tid] = acc;
        }
        if (tid < kC) s_hidden[tid] = fmaxf(W.vfc1_b[tid] + s_hpart[tid] + s_hpart[kC + tid], 0.f) * W.vfc2_w[tid];
        solo_epi_sync();
        if (warp == 0) {
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
          // ---- the tree step of the game (tree_device.cuh): consume the answer, expand, back up, play the move when the
          // search is complete, select the next leaf and write its request to P.nn_in[game]
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
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while another one of the cluster may still signal it or write into its buffers
  if (warp == kSoloEpiWarps) tmem_dealloc<G::TmemCols>(tmem);
}

template <int B>
cudaError_t launch_solo_t(const TowerWeights& w, const TreeParams& p, int n_games, int rounds, cudaStream_t s) {
  using SL = SoloSmem<B>;
  static_assert(SL::total <= 232448, "solo tower kernel exceeds 227 KB of shared memory");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tower_solo_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::total);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  tower_solo_kernel<B><<<dim3((unsigned)(n_games * kSoloNC)), dim3(kSoloThreads), SL::total, s>>>(w, p, rounds);
  return cudaGetLastError();
}

}  // namespace

// How many games the solo kernel takes (one cluster of four CTAs each, all clusters resident at once).
int solo_max_games(int num_sms) {
  const int g = (num_sms - 16) / kSoloNC;  // clusters of 4 strand up to 16 of the 148 SMs (GPCs of 16 / 18 / 20 SMs)
  return g < 1 ? 1 : g;
}

// `rounds` simulations for each of the game slots [0, n_games) in one launch; needs p.static_slots = 1 and every running
// game in ST_WAIT_NN with its request in p.nn_in[game] (engine.cu enter_persist), w.conv_quad loaded.
cudaError_t launch_selfplay_solo(const TowerWeights& w, int B, const TreeParams& p, int n_games, int rounds, cudaStream_t s) {
  if (w.n_layers > kMaxLayers || !p.static_slots || w.conv_quad == nullptr || n_games < 1 || rounds < 1) return cudaErrorInvalidValue;
  if (B == 9) return launch_solo_t<9>(w, p, n_games, rounds, s);
  if (B == 15) return launch_solo_t<15>(w, p, n_games, rounds, s);
  return cudaErrorInvalidValue;
}

}  // namespace ao
