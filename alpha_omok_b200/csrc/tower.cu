// PVNet.forward (model.py:97-104): the lock-step tower kernels - the hi/lo split-precision mode every trained net runs
// (AO_NN_FP16X3, CTA pairs) and two comparison variants of the single-pass mode (AO_NN_FP16_LOCKSTEP: CTA pairs,
// AO_NN_FP16_1CTA: cta_group::1).  The default single-pass kernel is tower_stag.cu; data layout and MMAs are shared:
//
// One persistent CTA per SM; the whole residual tower of a pass (3 games of 9x9 or 1 game of 15x15 = two 128-row tiles)
// runs with the activations resident in shared memory and the accumulators in TMEM; only the leaf descriptors (136 B)
// are read from and policy / value written to HBM, the fp16 weights stream from L2 through a 1-D bulk-copy ring.
//
// Implicit GEMM without im2col and without padding: activations live in smem as [16 k-chunks][Rows][8 halves] (K-major,
// no swizzle, 16 B per row and chunk).  Board cells are packed densely (row = game * A + cell, row stride S = B), so tap
// (dy,dx) of the 3x3 stencil is the SAME buffer addressed from a start row shifted by dy*B+dx; taps that fall off the
// board are dropped with tcgen05.mma's disable-output-lane masks (a disabled accumulator row is not updated = a zero
// contribution).  9 taps x 8 k-steps of tcgen05.mma (K = 16) per tile and layer.
// TMEM: per tile 128 columns accA + 128 columns accB.  Single-pass mode: accB keeps the fp32 block input x, so
// `out += residual` (model.py:29) is the accumulate flag of conv2's first MMA.  Split mode: every layer accumulates in a
// fresh accA (low-order terms first, see the MMA issuer), accB is the epilogue's fp32 stash of x and the residual add
// is a round-to-nearest fp32 add in the epilogue.
// BN (eval) is folded: scale into the fp16 weights, shift into the fp32 bias added in the epilogue.
//
// Warp roles: warps 0-7 epilogue (TMEM lane quarter = warp%4, tile = warp/4), warp 8 weight producer, warp 9 (and, for
// CTA pairs, warp 10) MMA issuer + TMEM owner.
#include <cuda_fp16.h>
#include <stdio.h>

#include "engine.h"
#include "sm100_ptx.cuh"

#include "tower_common.cuh"

namespace ao {

namespace {

// X3 = error-compensated mode: activations and weights are split into fp16 hi + lo parts and every k-step issues
// a_hi*w_hi + a_hi*w_lo + a_lo*w_hi (fp32 accumulate) - ~22 significant bits per operand; needed for 1e-4 on
// trained nets.  A weight stage is then half a tap: [hi 16 KB][lo 16 KB].
template <int B, int RING16, bool X3>  // RING16: size of the weight ring in 16 KB units
struct SmemLayout {
  using G = Geo<B>;
  static constexpr int act = 0;
  static constexpr int act_pad = (G::ActBytes + 1023) / 1024 * 1024;
  static constexpr int act_lo = act_pad;                                // only in X3 mode
  static constexpr int wring = X3 ? 2 * act_pad : act_pad;
  static constexpr int bias = wring + RING16 * (kStageBytes / 2);      // [kMaxLayers][128] f32
  static constexpr int headw = bias + kMaxLayers * kC * 4;              // [3][128] f32
  static constexpr int feat = headw + 3 * kC * 4;                       // [GPC][3][A] f32 (p0, p1, v)
  static constexpr int logits = feat + G::GPC * 3 * G::A * 4;           // [GPC][A]
  static constexpr int hidden = logits + G::GPC * G::A * 4;             // [GPC][128]
  static constexpr int red = hidden + G::GPC * kC * 4;                  // [GPC][2]
  static constexpr int masks = (red + G::GPC * 2 * 4 + 15) / 16 * 16;   // [kTiles][9 taps][4] disable-output-lane words
  static constexpr int bars = masks + kTiles * 9 * 4 * 4;               // mbarriers
  static constexpr int total = bars + (3 * RING16 + 2) * 8 + 16;  // full / empty / peer_full per slot + act + acc
};

// PAIR = CTA pairs (cluster of 2, tcgen05 cta_group::2, plain precision only): one M=256 MMA spans both CTAs' tiles,
// each CTA stages only ITS half of the output channels of B (half the smem operand reads and half the L2 weight
// traffic per SM); the leader CTA (cluster rank 0) issues, commits are multicast to both CTAs' barriers, and the
// peer's warp 9 forwards "my weights landed" / "my epilogue is done" to the leader.
template <int B, int RING16, bool X3, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
tower_kernel(TowerWeights W, const LeafIn* __restrict__ in, NNQueue q, int n_max,
             float* __restrict__ policy, float* __restrict__ value) {
  using G = Geo<B>;
  using SL = SmemLayout<B, RING16, X3>;
  extern __shared__ __align__(1024) uint8_t smem[];

  int n = n_max;
  uint32_t qbase = 0u;  // ring position of request 0 of this launch
  if (q.tail != nullptr) {
    qbase = *q.head;
    n = (int)(*q.tail - qbase);
    if (n > n_max) n = n_max;
    n = queue_serve_count<G::GPC>(n, q.defer);
  }
  const uint32_t qmask = q.tail != nullptr ? q.mask : 0xFFFFFFFFu;
  {
    int g0_, ng_, nt_;
    if (!get_pass<G::GPC, G::A, PAIR>(0, n, g0_, ng_, nt_)) {  // nothing to do for this CTA
      queue_finish(q, qbase, n);
      return;
    }
  }

  uint8_t* s_act = smem + SL::act;
  uint8_t* s_act_lo = smem + SL::act_lo;  // X3 only
  uint8_t* s_w = smem + SL::wring;
  float* s_bias = reinterpret_cast<float*>(smem + SL::bias);
  float* s_headw = reinterpret_cast<float*>(smem + SL::headw);
  float* s_feat = reinterpret_cast<float*>(smem + SL::feat);
  float* s_logits = reinterpret_cast<float*>(smem + SL::logits);
  float* s_hidden = reinterpret_cast<float*>(smem + SL::hidden);
  float* s_red = reinterpret_cast<float*>(smem + SL::red);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + SL::masks);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + SL::bars);
  // ring slots: 32 KB each, or 16 KB in the pair mode (a CTA stages only its half of the output channels)
  constexpr int NSLOT = PAIR ? RING16 : RING16 / 2;
  constexpr uint32_t kSlotBytes = PAIR ? kStageBytes / 2 : kStageBytes;
  uint64_t* bar_empty = bar_full + NSLOT;
  uint64_t* bar_act = bar_empty + NSLOT;    // epilogue -> MMA: operand written, accumulators drained
  uint64_t* bar_acc = bar_act + 1;          // MMA -> epilogue: layer accumulated
  uint64_t* bar_peer_full = bar_acc + 1;    // PAIR, leader only: the peer CTA's half of stage s has landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_peer_full + NSLOT);
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  // bytes of one tap of this CTA's B operand (PAIR: its 64 of the 128 output channels)
  constexpr uint32_t kTapBytes = PAIR ? kStageBytes / 2 : kStageBytes;
  constexpr uint32_t kStemTapBytes = PAIR ? kStemStageBytes / 2 : kStemStageBytes;
  constexpr uint32_t kBRows = PAIR ? kC / 2 : kC;  // rows (output channels) of B in this CTA's smem

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_layers = W.n_layers;

  // ---------------- one-time setup
  for (int i = tid; i < G::ActBytes / 16; i += kThreads) reinterpret_cast<uint4*>(s_act)[i] = make_uint4(0, 0, 0, 0);
  if (X3)
    for (int i = tid; i < G::ActBytes / 16; i += kThreads) reinterpret_cast<uint4*>(s_act_lo)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < n_layers * kC; i += kThreads) s_bias[i] = W.bias[i];
  for (int i = tid; i < 3 * kC; i += kThreads) s_headw[i] = W.head_w[i];
  // disable-output-lane masks: bit r of s_mask[tile][tap] set <=> for stream row tile*128+r the tap's source cell is
  // off the board (the MMA then leaves that accumulator row untouched = contributes zero)
  for (int i = tid; i < kTiles * 9 * 4; i += kThreads) {
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
      mbar_init(&bar_empty[s], PAIR ? 2 : 1);  // PAIR: one commit per MMA stream (tile)
    }
    mbar_init(bar_act, (PAIR && leader) ? kEpiThreads + 1 : kEpiThreads);  // + the peer's forwarded arrival
    mbar_init(bar_acc, PAIR ? 2 : 1);
    for (int s = 0; s < NSLOT; ++s) mbar_init(&bar_peer_full[s], 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    if (PAIR) tmem_alloc_pair<512>(s_tmem);
    else tmem_alloc<512>(s_tmem);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;

  if (warp == 8) {
    // =========================================================== weight producer (one lane)
    if (lane == 0) {
      uint32_t it = 0;
      int g0, ng, ntiles;
      for (int k = 0; get_pass<G::GPC, G::A, PAIR>(k, n, g0, ng, ntiles); ++k) {
        size_t off = 0;  // byte offset of the layer inside the packed conv weights (hi and lo buffers share the layout)
        const uint8_t* w_hi = reinterpret_cast<const uint8_t*>(W.conv_hi);
        const uint8_t* w_lo = reinterpret_cast<const uint8_t*>(W.conv_lo);
        for (int l = 0; l < n_layers; ++l) {
          // stages of this layer. plain: 9 taps. X3: stem 9 x [hi|lo]; residual layers 18 half-tap [hi|lo] stages
          // (the small lo-terms are accumulated first, see the MMA issuer) followed by 9 full-tap hi stages.
          const int n_st = !X3 ? 9 : (l == 0 ? 9 : 27);
          const uint32_t layer_bytes = l == 0 ? 9u * kStemStageBytes : 9u * kStageBytes;
          for (int st = 0; st < n_st; ++st, ++it) {
            const int s = it % NSLOT;
            const uint32_t ph = (it / NSLOT) & 1u;
            mbar_wait(&bar_empty[s], ph ^ 1u);
            uint8_t* dst = s_w + s * kSlotBytes;
            if (PAIR && !X3) {  // conv_pair: per tap [cluster rank][16 k-chunks][64 co][8]
              const uint32_t part = l == 0 ? kStemTapBytes : kTapBytes;
              mbar_arrive_expect_tx(&bar_full[s], part);
              bulk_g2s(dst, reinterpret_cast<const uint8_t*>(W.conv_pair) + off + ((size_t)st * 2 + cta_rank) * part, part,
                       &bar_full[s]);
            } else if (!X3) {
              const uint32_t part = l == 0 ? kStemStageBytes : kStageBytes;
              mbar_arrive_expect_tx(&bar_full[s], part);
              bulk_g2s(dst, w_hi + off + (size_t)st * part, part, &bar_full[s]);
            } else if (!PAIR) {
              if (l == 0 || st < 18) {
                const uint32_t part = l == 0 ? kStemStageBytes : kStageBytes / 2;
                mbar_arrive_expect_tx(&bar_full[s], 2 * part);
                bulk_g2s(dst, w_hi + off + (size_t)st * part, part, &bar_full[s]);
                bulk_g2s(dst + part, w_lo + off + (size_t)st * part, part, &bar_full[s]);
              } else {
                mbar_arrive_expect_tx(&bar_full[s], kStageBytes);
                bulk_g2s(dst, w_hi + off + (size_t)(st - 18) * kStageBytes, kStageBytes, &bar_full[s]);
              }
            } else {  // X3 in pair mode: this CTA's 64 output channels; per tap [rank][16 k-chunks][64][8]
              const uint8_t* p_hi = reinterpret_cast<const uint8_t*>(W.conv_pair) + off;
              const uint8_t* p_lo = reinterpret_cast<const uint8_t*>(W.conv_pair_lo) + off;
              if (l == 0) {                      // [hi 2 KB | lo 2 KB]
                const size_t o = ((size_t)st * 2 + cta_rank) * kStemTapBytes;
                mbar_arrive_expect_tx(&bar_full[s], 2 * kStemTapBytes);
                bulk_g2s(dst, p_hi + o, kStemTapBytes, &bar_full[s]);
                bulk_g2s(dst + kStemTapBytes, p_lo + o, kStemTapBytes, &bar_full[s]);
              } else if (st < 18) {              // lo phase: half a tap, [hi 8 KB | lo 8 KB]
                const size_t o = ((size_t)(st >> 1) * 2 + cta_rank) * kTapBytes + (size_t)(st & 1) * (kTapBytes / 2);
                mbar_arrive_expect_tx(&bar_full[s], kTapBytes);
                bulk_g2s(dst, p_hi + o, kTapBytes / 2, &bar_full[s]);
                bulk_g2s(dst + kTapBytes / 2, p_lo + o, kTapBytes / 2, &bar_full[s]);
              } else {                           // hi phase: a whole tap of hi weights, 16 KB
                const size_t o = ((size_t)(st - 18) * 2 + cta_rank) * kTapBytes;
                mbar_arrive_expect_tx(&bar_full[s], kTapBytes);
                bulk_g2s(dst, p_hi + o, kTapBytes, &bar_full[s]);
              }
            }
          }
          off += layer_bytes;
        }
      }
    }
  } else if (warp == 9 || (PAIR && warp == 10)) {
    // =========================================================== MMA issuer(s)
    // The whole warp runs the (warp-uniform) control flow so descriptors live in uniform registers; one elected
    // lane issues tcgen05.mma / commit.  CTA pairs use two issuing warps on two scheduler ports, one per tile (a
    // single thread gets one M256 N128 K16 MMA per ~109 cycles out of the tensor pipe, two threads one per ~87:
    // tools/probe_mma_rate.py); both run the same control flow and commit to the same barriers (count 2).
    const int stream = warp == 9 ? 0 : 1;
    if (PAIR && !leader) {
      // ---- peer CTA of a pair: no MMA issue; one lane forwards this CTA's readiness to the leader's barriers.
      // Two independent event streams (operand-ready per layer, weights-landed per stage) are polled without
      // blocking so that neither delays the other.
      if (stream == 0 && lane == 0) {
        int g0, ng, ntiles, n_k = 0;
        while (get_pass<G::GPC, G::A, PAIR>(n_k, n, g0, ng, ntiles)) ++n_k;
        const uint32_t n_act = (uint32_t)(n_k * n_layers);
        const uint32_t n_stage = X3 ? (uint32_t)n_k * (9u + (uint32_t)(n_layers - 1) * 27u) : n_act * 9u;
        uint32_t ai = 0, si = 0;
        const uint64_t t0 = globaltimer_ns();
        uint32_t spins = 0;
        while (ai < n_act || si < n_stage) {
          bool progress = false;
          if (ai < n_act && mbar_test_wait(bar_act, ai & 1u)) {  // this CTA's 256 epilogue threads wrote the operand
            mbar_arrive_remote(bar_act, 0u);
            ++ai;
            progress = true;
          }
          if (si < n_stage) {
            const uint32_t s = si % NSLOT;
            if (mbar_test_wait(&bar_full[s], (si / NSLOT) & 1u)) {  // this CTA's half of the stage has landed
              mbar_arrive_remote_relaxed(&bar_peer_full[s], 0u);
              ++si;
              progress = true;
            }
          }
          if (!progress && (++spins & 0x3FFu) == 0u && globaltimer_ns() - t0 > 8000000000ull) __trap();
        }
      }
    } else {
      const uint32_t idesc = umma_idesc_f16_f32(PAIR ? 256 : 128, 128);
      const uint32_t lbo_a = (uint32_t)G::Rows * 16u;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(s_act), lbo_a);
      const uint32_t desc_hi = umma_desc_hi(128u);
      constexpr uint32_t kAStep = (2u * (uint32_t)G::Rows * 16u) >> 4;  // two k-chunks further along K
      constexpr uint32_t kBStep = (2u * kBRows * 16u) >> 4;
      uint32_t it = 0, act_phase = 0;
      AO_DBG(long long dbg_act_wait = 0, dbg_full_wait = 0; const long long dbg_t0 = W.dbg ? clock64() : 0;)
      int g0, ng, ntiles;
      for (int k = 0; get_pass<G::GPC, G::A, PAIR>(k, n, g0, ng, ntiles); ++k) {
        for (int l = 0; l < n_layers; ++l) {
          // plain mode: stem and conv2 accumulate in accB, which already holds the block input x (residual);
          // X3 mode: every layer accumulates in a fresh accA and accB is only the epilogue's fp32 stash of x
          const bool to_b = !X3 && (l & 1) == 0;
          const bool residual = to_b && l > 0;
          AO_DBG(const long long t_a0 = W.dbg ? clock64() : 0;)
          if (PAIR) mbar_wait_cluster(bar_act, act_phase);  // 256 local arrivals + the peer's forwarded one
          else mbar_wait(bar_act, act_phase);
          act_phase ^= 1u;
          tc_fence_after_sync();
          AO_DBG(if (W.dbg) dbg_act_wait += clock64() - t_a0;)
          const int n_st = !X3 ? 9 : (l == 0 ? 9 : 27);
          for (int st = 0; st < n_st; ++st, ++it) {
            const int s = it % NSLOT;
            const uint32_t ph = (it / NSLOT) & 1u;
            AO_DBG(const long long t_f0 = W.dbg ? clock64() : 0;)
            mbar_wait(&bar_full[s], ph);
            if (PAIR) mbar_wait_cluster(&bar_peer_full[s], ph);
            tc_fence_after_sync();
            AO_DBG(if (W.dbg) dbg_full_wait += clock64() - t_f0;)
            const bool lo_phase = X3 && l > 0 && st < 18;
            // the weights are packed with the centre tap first: it has no disabled rows, so the first MMA of a fresh
            // accumulation (accumulate = 0) writes every accumulator row
            const int ti = !X3 || l == 0 ? st : (lo_phase ? st >> 1 : st - 18);  // position in the packed tap order
            const int t = ti == 0 ? 4 : (ti <= 4 ? ti - 1 : ti);                 // tap = (dy+1)*3 + (dx+1)
            const int kh = lo_phase ? st & 1 : 0;  // which half of the 128 input channels
            const int shift = (t / 3 - 1) * G::S + (t % 3 - 1);
            const uint32_t b_lo0 = umma_desc_lo(smem_u32(s_w + s * kSlotBytes), kBRows * 16u);
            if (elect_one()) {
#pragma unroll
              for (int tile = 0; tile < kTiles; ++tile) {
                if (tile >= ntiles) break;  // one-game tail pass: tile 1 holds no board
                if (PAIR && tile != stream) continue;
                const uint32_t* mk = s_mask + (tile * 9 + t) * 4;
                const uint32_t m0 = mk[0], m1 = mk[1], m2 = mk[2], m3 = mk[3];
                const uint32_t d_tmem = tmem + (uint32_t)(tile * 256 + (to_b ? 128 : 0));
                const uint32_t a_lo = a_lo0 + (uint32_t)(G::Halo + tile * kTileRows + shift);  // 16 B per row
                // one MMA = M128 of this CTA, or M256 over the same tile of both CTAs of a pair (same geometry / masks)
                auto mma = [&](uint32_t a, uint32_t b, uint32_t acc) {
                  if (PAIR) umma_f16_ss_pair_masked(d_tmem, a, b, desc_hi, idesc, acc, m0, m1, m2, m3);
                  else umma_f16_ss_lohi_masked(d_tmem, a, b, desc_hi, idesc, acc, m0, m1, m2, m3);
                };
                if (!X3) {
                  mma(a_lo, b_lo0, (residual || st > 0) ? 1u : 0u);
                  if (l > 0) {
#pragma unroll
                    for (int j = 1; j < kC / 16; ++j) mma(a_lo + (uint32_t)j * kAStep, b_lo0 + (uint32_t)j * kBStep, 1u);
                  }
                } else {
                  // The tensor core truncates (does not round) its fp32 accumulation: ~1.3 ulp(acc) lost per MMA
                  // (tools/probe_accum.py).  So the lo-terms a_hi*w_lo + a_lo*w_hi are accumulated FIRST, while the
                  // accumulator is still ~2^-11 of its final magnitude, and only the hi*hi MMAs run at full magnitude.
                  constexpr uint32_t kALoOff = (uint32_t)SL::act_pad >> 4;      // lo activations sit above the hi ones
                  constexpr uint32_t kWLoStem = kStemTapBytes >> 4;            // stage = [hi | lo]
                  constexpr uint32_t kWLoHalf = (kTapBytes / 2u) >> 4;
                  if (l == 0) {       // stem: inputs are exact ({0,1}), a_lo == 0
                    mma(a_lo, b_lo0 + kWLoStem, st > 0 ? 1u : 0u);
                    mma(a_lo, b_lo0, 1u);
                  } else if (lo_phase) {
                    const uint32_t a0 = a_lo + (uint32_t)(kh * 4) * kAStep;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const uint32_t aj = a0 + (uint32_t)j * kAStep, bj = b_lo0 + (uint32_t)j * kBStep;
                      mma(aj, bj + kWLoHalf, (st > 0 || j > 0) ? 1u : 0u);
                      mma(aj + kALoOff, bj, 1u);
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < kC / 16; ++j) mma(a_lo + (uint32_t)j * kAStep, b_lo0 + (uint32_t)j * kBStep, 1u);
                  }
                }
              }
              if (PAIR) {
                umma_commit_pair(&bar_empty[s]);
                if (st == n_st - 1) umma_commit_pair(bar_acc);
              } else {
                umma_commit(&bar_empty[s]);
                if (st == n_st - 1) umma_commit(bar_acc);
              }
            }
            __syncwarp();
          }
        }
      }
#ifdef AO_PROBE
      if (W.dbg && blockIdx.x == 0 && lane == 0 && stream == 0) {  // profiling counters (ao_tower_debug): MMA-issuer view of CTA 0
        atomicAdd(&W.dbg[0], (unsigned long long)(clock64() - dbg_t0));
        atomicAdd(&W.dbg[1], (unsigned long long)dbg_act_wait);
        atomicAdd(&W.dbg[2], (unsigned long long)dbg_full_wait);
        atomicAdd(&W.dbg[3], 1ull);
      }
#endif
    }
  } else if (warp < 8) {
    // =========================================================== epilogue warps (256 threads)
    const int tile = tid >> 7, r = tid & 127;
    const int R = tile * kTileRows + r;          // logical row in the CTA's padded position stream
    const int g_local = R / G::A;                // game of the pass
    const int pos = R % G::A;                    // cell
    const int yy = pos / B, xx = pos % B;
    const bool geo_valid = g_local < G::GPC;
    const uint32_t row_off = (uint32_t)(G::Halo + R) * 16u;
    const uint32_t chunk_stride = (uint32_t)G::Rows * 16u;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tile * 256);
    uint32_t acc_phase = 0;
    AO_DBG(long long dbg_acc_wait = 0, dbg_heads = 0; const bool dbg_on = W.dbg != nullptr && blockIdx.x == 0 && tid == 0;
           const long long dbg_e0 = dbg_on ? clock64() : 0;)

    int g0, ng, ntiles;
    for (int k = 0; get_pass<G::GPC, G::A, PAIR>(k, n, g0, ng, ntiles); ++k) {
      const bool valid = geo_valid && g_local < ng;
      // ---- input planes (utils.get_state_pt) -> channels 0..4 of chunk 0; chunk 1 = 0
      {
        uint4 c0 = make_uint4(0, 0, 0, 0);
        if (valid) {
          const LeafIn* li = &in[(qbase + (uint32_t)(g0 + g_local)) & qmask];
          const uint32_t b0 = (li->plane[0][yy] >> xx) & 1u, b1 = (li->plane[1][yy] >> xx) & 1u;
          const uint32_t b2 = (li->plane[2][yy] >> xx) & 1u, b3 = (li->plane[3][yy] >> xx) & 1u;
          const uint32_t b4 = li->colour & 1u;
          c0.x = (b0 ? 0x3C00u : 0u) | (b1 ? 0x3C000000u : 0u);
          c0.y = (b2 ? 0x3C00u : 0u) | (b3 ? 0x3C000000u : 0u);
          c0.z = (b4 ? 0x3C00u : 0u);
        }
        *reinterpret_cast<uint4*>(s_act + row_off) = c0;
        *reinterpret_cast<uint4*>(s_act + chunk_stride + row_off) = make_uint4(0, 0, 0, 0);
        if (X3) {  // {0,1} planes are exact in fp16: low parts are zero
          *reinterpret_cast<uint4*>(s_act_lo + row_off) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(s_act_lo + chunk_stride + row_off) = make_uint4(0, 0, 0, 0);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(bar_act);

      float hd0 = 0.f, hd1 = 0.f, hd2 = 0.f;
      for (int l = 0; l < n_layers; ++l) {
        const bool to_b = (l & 1) == 0;
        const bool last = l == n_layers - 1;
        const float* bias = s_bias + l * kC;
        AO_DBG(const long long t_w0 = dbg_on ? clock64() : 0;)
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1u;
        tc_fence_after_sync();
        AO_DBG(if (dbg_on) dbg_acc_wait += clock64() - t_w0;)
        const uint32_t stash_addr = lane_addr + 128u;                          // accB: fp32 block input x
        const uint32_t acc_addr = (!X3 && to_b) ? stash_addr : lane_addr;      // X3 accumulates in accA only
        // 4 chunks of 32 accumulator columns, TMEM loads double-buffered against the per-chunk math
        auto process = [&](uint32_t (&v)[32], const int qd) {
          const float4* b4 = reinterpret_cast<const float4*>(bias + qd * 32);
          if (X3 && to_b && l > 0) {  // residual add in fp32 round-to-nearest: out = conv2 + x (model.py:29)
            uint32_t xr[32];
            tmem_ld32(stash_addr + (uint32_t)(qd * 32), xr);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(xr[j]));
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bb = b4[j4];
            const float y0 = fmaxf(__uint_as_float(v[j4 * 4 + 0]) + bb.x, 0.f);
            const float y1 = fmaxf(__uint_as_float(v[j4 * 4 + 1]) + bb.y, 0.f);
            const float y2 = fmaxf(__uint_as_float(v[j4 * 4 + 2]) + bb.z, 0.f);
            const float y3 = fmaxf(__uint_as_float(v[j4 * 4 + 3]) + bb.w, 0.f);
            v[j4 * 4 + 0] = __float_as_uint(valid ? y0 : 0.f);
            v[j4 * 4 + 1] = __float_as_uint(valid ? y1 : 0.f);
            v[j4 * 4 + 2] = __float_as_uint(valid ? y2 : 0.f);
            v[j4 * 4 + 3] = __float_as_uint(valid ? y3 : 0.f);
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
              *reinterpret_cast<uint4*>(s_act + (uint32_t)(qd * 4 + cc) * chunk_stride + row_off) = pk;
              if (X3) {  // low parts: fp16(y - fp16(y))
                uint4 pl;
                const __half2* hh = reinterpret_cast<const __half2*>(&pk);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __half22float2(hh[e]);
                  const __half2 l2 = __floats2half2_rn(__uint_as_float(v[cc * 8 + 2 * e]) - f.x,
                                                       __uint_as_float(v[cc * 8 + 2 * e + 1]) - f.y);
                  reinterpret_cast<uint32_t*>(&pl)[e] = *reinterpret_cast<const uint32_t*>(&l2);
                }
                *reinterpret_cast<uint4*>(s_act_lo + (uint32_t)(qd * 4 + cc) * chunk_stride + row_off) = pl;
              }
            }
            if (to_b) tmem_st32(stash_addr + (uint32_t)(qd * 32), v);  // fp32 block input for the next residual add
          } else {
            // heads' 1x1 convolutions (model.py:44-46, 64-66) straight from the fp32 tower output
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = __uint_as_float(v[j]);
              hd0 = fmaf(x, s_headw[0 * kC + qd * 32 + j], hd0);
              hd1 = fmaf(x, s_headw[1 * kC + qd * 32 + j], hd1);
              hd2 = fmaf(x, s_headw[2 * kC + qd * 32 + j], hd2);
            }
          }
        };
        {
          uint32_t va[32], vb[32];
          tmem_ld32(acc_addr, va);
          tmem_ld_wait();
          tmem_ld32(acc_addr + 32u, vb);
          process(va, 0);
          tmem_ld_wait();
          tmem_ld32(acc_addr + 64u, va);
          process(vb, 1);
          tmem_ld_wait();
          tmem_ld32(acc_addr + 96u, vb);
          process(va, 2);
          tmem_ld_wait();
          process(vb, 3);
        }
        if (!last) {
          if (to_b) tmem_st_wait();
          fence_proxy_async_smem();
          tc_fence_before_sync();
          mbar_arrive(bar_act);
        }
      }
      // ---- heads (model.py:43-50, 63-73)
      AO_DBG(const long long t_h0 = dbg_on ? clock64() : 0;)
      if (valid) {
        float* f = s_feat + g_local * 3 * G::A;
        f[0 * G::A + pos] = fmaxf(hd0 + W.head_b[0], 0.f);
        f[1 * G::A + pos] = fmaxf(hd1 + W.head_b[1], 0.f);
        f[2 * G::A + pos] = fmaxf(hd2 + W.head_b[2], 0.f);
      }
      epi_bar_sync();
      float logit = 0.f;
      const int pg = tid / G::A, po = tid % G::A;  // policy FC: thread = (game, output)
      const bool p_thread = tid < G::GPC * G::A && pg < ng;
      if (p_thread) {
        const float* f = s_feat + pg * 3 * G::A;
        float acc = W.pfc_b[po];
        const float* wt = W.pfc_wT + po;
#pragma unroll 18
        for (int k = 0; k < 2 * G::A; ++k) acc = fmaf(__ldg(wt + (size_t)k * G::A), f[k], acc);
        logit = acc;
        s_logits[pg * G::A + po] = acc;
      }
      for (int vi = tid; vi < G::GPC * kC; vi += kEpiThreads) {  // value FC1: (game, hidden unit)
        const int vg = vi / kC, vj = vi % kC;
        if (vg < ng) {
          const float* f = s_feat + vg * 3 * G::A + 2 * G::A;
          float acc = W.vfc1_b[vj];
          const float* wt = W.vfc1_wT + vj;
#pragma unroll 27
          for (int k = 0; k < G::A; ++k) acc = fmaf(__ldg(wt + (size_t)k * kC), f[k], acc);
          s_hidden[vg * kC + vj] = fmaxf(acc, 0.f) * W.vfc2_w[vj];
        }
      }
      epi_bar_sync();
      if (warp < ng) {  // warp g: softmax statistics and the value of game g
        float mx = -3.0e38f;
        for (int k = lane; k < G::A; k += 32) mx = fmaxf(mx, s_logits[warp * G::A + k]);
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        float sum = 0.f;
        for (int k = lane; k < G::A; k += 32) sum += expf(s_logits[warp * G::A + k] - mx);
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        float hv = 0.f;
        for (int k = lane; k < kC; k += 32) hv += s_hidden[warp * kC + k];
#pragma unroll
        for (int o = 16; o; o >>= 1) hv += __shfl_xor_sync(0xFFFFFFFFu, hv, o);
        if (lane == 0) {
          s_red[warp * 2 + 0] = mx;
          s_red[warp * 2 + 1] = sum;
          value[(qbase + (uint32_t)(g0 + warp)) & qmask] = tanhf(hv + W.vfc2_b);
        }
      }
      epi_bar_sync();
      if (p_thread)
        policy[(size_t)((qbase + (uint32_t)(g0 + pg)) & qmask) * G::A + po] = expf(logit - s_red[pg * 2]) / s_red[pg * 2 + 1];
      // s_feat / s_logits are rewritten only after the next pass's 21 layers: no extra barrier needed
      AO_DBG(if (dbg_on) dbg_heads += clock64() - t_h0;)
    }
#ifdef AO_PROBE
    if (dbg_on) {
      atomicAdd(&W.dbg[4], (unsigned long long)(clock64() - dbg_e0));
      atomicAdd(&W.dbg[5], (unsigned long long)dbg_acc_wait);
      atomicAdd(&W.dbg[6], (unsigned long long)dbg_heads);
    }
#endif
  }
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its partner may still signal / use it
  if (warp == 9) {
    if (PAIR) tmem_dealloc_pair<512>(tmem);
    else tmem_dealloc<512>(tmem);
  }
  queue_finish(q, qbase, n);
}

// dense float states [n][C][B][B] -> LeafIn row masks
__global__ void pack_states_kernel(const float* __restrict__ st, int n, int B, int C, LeafIn* __restrict__ out,
                                   int* __restrict__ bad) {
  const int i = blockIdx.x;
  const int t = threadIdx.x;  // one thread per (plane, row)
  if (i >= n) return;
  const float* s = st + (size_t)i * C * B * B;
  if (t < 4 * kRowsPad) {
    const int k = t / kRowsPad, y = t % kRowsPad;
    uint32_t m = 0;
    if (k < C - 1 && y < B)
      for (int x = 0; x < B; ++x) {
        const float v = s[(k * B + y) * B + x];
        if (v != 0.f && v != 1.f) *bad = 1;
        if (v != 0.f) m |= 1u << x;
      }
    out[i].plane[k][y] = (uint16_t)m;
  }
  if (t == 0) {
    const float c = s[(size_t)(C - 1) * B * B];
    for (int k = 0; k < B * B; ++k)
      if (s[(size_t)(C - 1) * B * B + k] != c) *bad = 1;
    if (c != 0.f && c != 1.f) *bad = 1;
    out[i].colour = c != 0.f ? 1u : 0u;
    out[i].game = i;
  }
}

template <int B, int RING16, bool X3, bool PAIR>
cudaError_t launch_tower_t(const TowerWeights& w, const LeafIn* in, const NNQueue& q, int n_max, float* policy,
                           float* value, int num_sms, cudaStream_t s) {
  using SL = SmemLayout<B, RING16, X3>;
  static_assert(SL::total <= 232448, "tower kernel exceeds 227 KB of shared memory");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tower_kernel<B, RING16, X3, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::total);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = n_max < num_sms ? n_max : num_sms;  // see get_pass: up to one game per CTA in a ragged wave
  if (grid <= 0) return cudaSuccess;
  if (!PAIR) {
    tower_kernel<B, RING16, X3, PAIR><<<grid, kThreads, SL::total, s>>>(w, in, q, n_max, policy, value);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((grid + 1) & ~1));  // whole CTA pairs
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = SL::total;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, tower_kernel<B, RING16, X3, PAIR>, w, in, q, n_max, policy, value);
}

}  // namespace

cudaError_t launch_tower(const TowerWeights& w, int B, int precision, const LeafIn* in, const NNQueue& q, int n_max,
                         float* policy, float* value, int num_sms, cudaStream_t s) {
  if (w.n_layers > kMaxLayers) return cudaErrorInvalidValue;
  if (precision == AO_NN_FP16X3) {
    if (B == 9) return launch_tower_t<9, 4, true, true>(w, in, q, n_max, policy, value, num_sms, s);
    if (B == 15) return launch_tower_t<15, 3, true, true>(w, in, q, n_max, policy, value, num_sms, s);
    return cudaErrorInvalidValue;
  }
  if (precision == AO_NN_FP16_1CTA) {  // single-CTA variant (cta_group::1), kept for comparison
    if (B == 9) return launch_tower_t<9, 8, false, false>(w, in, q, n_max, policy, value, num_sms, s);
    if (B == 15) return launch_tower_t<15, 8, false, false>(w, in, q, n_max, policy, value, num_sms, s);
    return cudaErrorInvalidValue;
  }
  if (precision == AO_NN_FP16)  // default: CTA pairs with staggered tiles (tower_stag.cu)
    return launch_tower_stag(w, B, in, q, n_max, policy, value, num_sms, s);
  if (precision != AO_NN_FP16_LOCKSTEP) return cudaErrorInvalidValue;
  // CTA pairs (cta_group::2), MMA and epilogue in lock-step
  if (B == 9) return launch_tower_t<9, 8, false, true>(w, in, q, n_max, policy, value, num_sms, s);
  if (B == 15) return launch_tower_t<15, 8, false, true>(w, in, q, n_max, policy, value, num_sms, s);
  return cudaErrorInvalidValue;
}

cudaError_t launch_pack_states(const float* states_dev, int n, int B, int inplanes, LeafIn* out, int* bad_flag_dev,
                               cudaStream_t s) {
  pack_states_kernel<<<n, 64, 0, s>>>(states_dev, n, B, inplanes, out, bad_flag_dev);
  return cudaGetLastError();
}

}  // namespace ao
