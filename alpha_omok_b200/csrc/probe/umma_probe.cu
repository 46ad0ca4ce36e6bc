// Hardware probe / unit test for the building block of the tower kernel (tower.cu):
//   out[128][128] = init + sum_t  ACT[row0 + shift[t] + r][0:128] * W[t][0:128][co]
// with ACT resident in shared memory in the K-major no-swizzle chunk layout [16 k-chunks][ROWS][8 halves],
// the per-tap A operand being the SAME buffer addressed through a start address shifted by whole rows (16 B each),
// weights streamed by 1-D bulk async copies into a 2-stage ring, accumulators in TMEM (optionally pre-loaded with
// `init` through tcgen05.st = the residual trick), read back with tcgen05.ld.32x32b.
// Exposed through the C ABI as ao_umma_probe (include/alpha_omok_b200.h).
#include <cuda_runtime.h>
#include <stdio.h>

#include "sm100_ptx.cuh"

namespace {

constexpr int kC = 128;           // channels (K per tap and N)
constexpr int kChunks = kC / 8;   // 16-byte k-chunks per row
constexpr int kStageBytes = kC * kC * 2;  // one tap of weights: [16][128][8] halves = 32 KB
constexpr int kStages = 2;

struct ProbeSmem {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __half* __restrict__ act, int rows, const __half* __restrict__ wpacked,
                  const float* __restrict__ init, float* __restrict__ out, int row0, int ntaps,
                  const int* __restrict__ shifts, const uint32_t* __restrict__ masks) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_act = smem;                                   // 16 * rows * 16 B
  uint32_t act_bytes = (uint32_t)(kChunks * rows * 16);
  act_bytes = (act_bytes + 1023u) & ~1023u;
  uint8_t* s_w = smem + act_bytes;                         // kStages * 32 KB
  ProbeSmem* ctl = reinterpret_cast<ProbeSmem*>(s_w + kStages * kStageBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      ao::mbar_init(&ctl->full[s], 1);
      ao::mbar_init(&ctl->empty[s], 1);
    }
    ao::mbar_init(&ctl->done, 1);
    ao::fence_mbar_init();
  }
  if (warp == 0) ao::tmem_alloc<128>(&ctl->tmem_base);

  // activations: global row-major [rows][128] -> smem [chunk][row][8]
  for (int i = tid; i < rows * kChunks; i += blockDim.x) {
    int r = i / kChunks, c = i % kChunks;
    uint4 v = *reinterpret_cast<const uint4*>(act + (size_t)r * kC + c * 8);
    *reinterpret_cast<uint4*>(s_act + ((size_t)c * rows + r) * 16) = v;
  }
  ao::fence_proxy_async_smem();
  ao::tc_fence_before_sync();
  __syncthreads();
  ao::tc_fence_after_sync();
  const uint32_t tmem = ctl->tmem_base;

  if (init != nullptr) {  // preload accumulator: row = tid (TMEM lane), 128 columns
    uint32_t v[32];
    for (int q = 0; q < 4; ++q) {
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(init[(size_t)tid * kC + q * 32 + j]);
      ao::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + q * 32, v);
    }
    ao::tmem_st_wait();
    ao::tc_fence_before_sync();
  }
  __syncthreads();
  ao::tc_fence_after_sync();

  if (warp == 0 && lane == 0) {
    // producer: stream one tap of weights per stage
    for (int t = 0; t < ntaps; ++t) {
      int s = t % kStages;
      uint32_t ph = (uint32_t)(t / kStages) & 1u;
      ao::mbar_wait(&ctl->empty[s], ph ^ 1u);
      ao::mbar_arrive_expect_tx(&ctl->full[s], kStageBytes);
      ao::bulk_g2s(s_w + s * kStageBytes, wpacked + (size_t)t * kC * kC, kStageBytes, &ctl->full[s]);
    }
  } else if (warp == 1 && lane == 0) {
    // MMA issuer
    const uint32_t idesc = ao::umma_idesc_f16_f32(128, 128);
    const uint32_t a_base = ao::smem_u32(s_act);
    const uint32_t lbo_a = (uint32_t)rows * 16u;
    uint32_t acc = init != nullptr ? 1u : 0u;
    for (int t = 0; t < ntaps; ++t) {
      int s = t % kStages;
      uint32_t ph = (uint32_t)(t / kStages) & 1u;
      ao::mbar_wait(&ctl->full[s], ph);
      ao::tc_fence_after_sync();
      const uint32_t b_base = ao::smem_u32(s_w + s * kStageBytes);
      const uint32_t a_row = (uint32_t)(row0 + shifts[t]);
      for (int j = 0; j < kC / 16; ++j) {
        uint64_t da = ao::umma_desc_kmajor_noswz(a_base + (uint32_t)(2 * j) * lbo_a + a_row * 16u, lbo_a, 128u);
        uint64_t db = ao::umma_desc_kmajor_noswz(b_base + (uint32_t)(2 * j) * (kC * 16u), kC * 16u, 128u);
        if (masks == nullptr) {
          ao::umma_f16_ss(tmem, da, db, idesc, acc);
        } else {  // per-tap disable-output-lane mask (4 x 32 bits)
          ao::umma_f16_ss_lohi_masked(tmem, (uint32_t)da, (uint32_t)db, (uint32_t)(da >> 32), idesc, acc,
                                      masks[4 * t + 0], masks[4 * t + 1], masks[4 * t + 2], masks[4 * t + 3]);
        }
        acc = 1u;
      }
      ao::umma_commit(&ctl->empty[s]);
    }
    ao::umma_commit(&ctl->done);
  }
  __syncwarp();
  ao::mbar_wait(&ctl->done, 0);
  ao::tc_fence_after_sync();

  {
    uint32_t v[32];
    for (int q = 0; q < 4; ++q) {
      ao::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + q * 32, v);
      ao::tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(size_t)tid * kC + q * 32 + j] = __uint_as_float(v[j]);
    }
  }
  ao::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ao::tmem_dealloc<128>(tmem);
}

}  // namespace

// Host entry (C ABI). All pointers are HOST pointers; returns 0 or a negative cudaError.
extern "C" int ao_umma_probe_masked(const uint16_t* act_f16, int rows, const uint16_t* wpacked_f16, const float* init,
                                    float* out, int row0, int ntaps, const int* shifts, const uint32_t* masks);
extern "C" int ao_umma_probe(const uint16_t* act_f16, int rows, const uint16_t* wpacked_f16, const float* init,
                             float* out, int row0, int ntaps, const int* shifts) {
  return ao_umma_probe_masked(act_f16, rows, wpacked_f16, init, out, row0, ntaps, shifts, nullptr);
}
// masks: optional [ntaps][4] uint32 disable-output-lane masks (bit r set: output row r is not updated by that tap)
extern "C" int ao_umma_probe_masked(const uint16_t* act_f16, int rows, const uint16_t* wpacked_f16, const float* init,
                                    float* out, int row0, int ntaps, const int* shifts, const uint32_t* masks) {
  if (rows < 128 || rows > 320 || ntaps < 1 || ntaps > 64) return -1;
  __half *d_act = nullptr, *d_w = nullptr;
  float *d_init = nullptr, *d_out = nullptr;
  int* d_sh = nullptr;
  uint32_t* d_mk = nullptr;
  cudaError_t e;
#define CK(x)                     \
  do {                            \
    e = (x);                      \
    if (e != cudaSuccess) {       \
      fprintf(stderr, "ao_umma_probe: %s -> %s\n", #x, cudaGetErrorString(e)); \
      return -(int)e - 1000;      \
    }                             \
  } while (0)
  CK(cudaMalloc(&d_act, (size_t)rows * kC * 2));
  CK(cudaMalloc(&d_w, (size_t)ntaps * kC * kC * 2));
  CK(cudaMalloc(&d_out, 128 * kC * 4));
  CK(cudaMalloc(&d_sh, ntaps * sizeof(int)));
  CK(cudaMemcpy(d_act, act_f16, (size_t)rows * kC * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, wpacked_f16, (size_t)ntaps * kC * kC * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_sh, shifts, ntaps * sizeof(int), cudaMemcpyHostToDevice));
  if (masks) {
    CK(cudaMalloc(&d_mk, ntaps * 4 * sizeof(uint32_t)));
    CK(cudaMemcpy(d_mk, masks, ntaps * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  if (init) {
    CK(cudaMalloc(&d_init, 128 * kC * 4));
    CK(cudaMemcpy(d_init, init, 128 * kC * 4, cudaMemcpyHostToDevice));
  }
  uint32_t act_bytes = ((uint32_t)(kChunks * rows * 16) + 1023u) & ~1023u;
  size_t smem = act_bytes + kStages * kStageBytes + sizeof(ProbeSmem) + 64;
  CK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem>>>(d_act, rows, d_w, d_init, d_out, row0, ntaps, d_sh, d_mk);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, d_out, 128 * kC * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_act); cudaFree(d_w); cudaFree(d_out); cudaFree(d_sh);
  if (d_init) cudaFree(d_init);
  if (d_mk) cudaFree(d_mk);
#undef CK
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Raw tcgen05.mma throughput probe: every CTA (pair) issues `iters` back-to-back M128(/256) N128 K16 kind::f16 MMAs on
// zero-filled shared-memory operands in the tower's layout and reports cycles per MMA.  Flavours:
//   0 cta_group::1 unmasked   1 cta_group::1 masked (zero masks)   2 cta_group::2 unmasked   3 cta_group::2 masked
//   +4: A start row cycles through the nine 3x3 tap shifts (unaligned core-matrix rows) instead of staying aligned
//   +8: 8 warps stream st.shared.v4 into an unrelated smem region meanwhile (epilogue-like smem write pressure)
namespace {

__device__ __forceinline__ void umma_f16_ss_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr int kRateRows = 288;
constexpr int kRateAct = 16 * kRateRows * 16;  // 73728
constexpr int kRateW = 32768;
constexpr int kRateScratch = 32768;

template <bool PAIR>
__global__ void __launch_bounds__(320, 1) umma_rate_kernel(int flavour, int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_act = smem;
  uint8_t* s_w = smem + kRateAct;
  uint8_t* s_scr = s_w + kRateW;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_scr + kRateScratch);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 4);
  volatile int* s_stop = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (kRateAct + kRateW + kRateScratch) / 16; i += 320) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) ao::mbar_init(&bar[i], 1);
    *s_stop = 0;
    ao::fence_mbar_init();
  }
  if (warp == 9) {
    if (PAIR) ao::tmem_alloc_pair<512>(s_tmem);
    else ao::tmem_alloc<512>(s_tmem);
  }
  ao::fence_proxy_async_smem();
  ao::tc_fence_before_sync();
  __syncthreads();
  if (PAIR) ao::cluster_sync_all();
  ao::tc_fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const bool leader = !PAIR || ao::cluster_ctarank() == 0u;
  const bool masked = flavour & 1, shifted = flavour & 4, pressure = flavour & 8;
  // several issuing warps (9, 8, 7, 6), each with its own accumulator and its share of the MMAs
  const int n_issuers = (flavour & 512) ? 4 : ((flavour & 256) ? 2 : 1);
  const bool two_issuers = n_issuers > 1;
  if (warp <= 9 && warp > 9 - n_issuers) {
    if (leader) {
      iters /= n_issuers;
      // 8192: N = 32; 16384: cta_group::2 with M = 128 (64 rows per CTA) instead of 256
      const int N = (flavour & 8192) ? 32 : (flavour & 16) ? 64 : ((flavour & 64) ? 256 : 128);
      const uint32_t idesc = ao::umma_idesc_f16_f32(PAIR ? ((flavour & 16384) ? 128 : 256) : 128, N);
      const uint32_t brows = (uint32_t)(PAIR ? N / 2 : N);
      const uint32_t a_lo0 = ao::umma_desc_lo(ao::smem_u32(s_act), kRateRows * 16u);
      const uint32_t b_lo0 = ao::umma_desc_lo(ao::smem_u32(s_w), brows * 16u);
      const bool sw128 = (flavour & 128) != 0;  // 128-byte swizzled K-major operands (timing only: zero data)
      const uint32_t desc_hi = sw128 ? (ao::umma_desc_hi(1024u) | (2u << 29)) : ao::umma_desc_hi(128u);
      const uint32_t kAStep = (2u * kRateRows * 16u) >> 4, kBStep = (2u * brows * 16u) >> 4;
      const long long t0 = clock64();
      if (ao::elect_one()) {
        uint32_t wph = 0;
        for (int it = 0; it < iters; ++it) {
          // 1024: commit every 8 MMAs (no wait); 2048: commit AND wait for completion every 16 MMAs (shallow queue);
          // 4096: the same every 32 MMAs
          if ((flavour & 1024) && it > 0) {
            if (PAIR) ao::umma_commit_pair(&bar[3]); else ao::umma_commit(&bar[3]);
          }
          if (((flavour & 2048) && it > 0 && (it & 1) == 0) || ((flavour & 4096) && it > 0 && (it & 3) == 0)) {
            if (PAIR) ao::umma_commit_pair(&bar[2]); else ao::umma_commit(&bar[2]);
            ao::mbar_wait(&bar[2], wph);
            wph ^= 1u;
          }
          const int tap = it % 9;
          const int shift = shifted ? (tap / 3 - 1) * 9 + (tap % 3 - 1) : 0;
          const uint32_t a = a_lo0 + (uint32_t)(16 + ((it / 9) & 1) * 128 + shift);
          const uint32_t d0 = n_issuers == 4 ? tmem + (uint32_t)((9 - warp) * 128)
                              : two_issuers ? tmem + (uint32_t)(((warp & 1) * 2 + ((it / 9) & 1)) * 128)
                                          : tmem + (uint32_t)((flavour & 64) ? ((it / 9) & 1) * 256 : ((it / 9) & 3) * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t d = (flavour & 32) ? tmem + (uint32_t)((j & 3) * 128) : d0;
            if (sw128) {  // [k-block of 64][row][128 B]: k-step j = block j/4, +32 B per step; tap shift = whole rows
              const uint32_t as = (ao::smem_u32(s_act) >> 4) + (uint32_t)(j >> 2) * (kRateRows * 8u) + (uint32_t)(j & 3) * 2u +
                                  (uint32_t)(16 + ((it / 9) & 1) * 128 + shift) * 8u;
              const uint32_t bs = (ao::smem_u32(s_w) >> 4) + (uint32_t)(j >> 2) * (brows * 8u) + (uint32_t)(j & 3) * 2u;
              const uint32_t hi = desc_hi | ((as >> 3) & 7u) << 17;  // base_offset for a start that is not 1024 B aligned
              const uint32_t alo = (as & 0x3FFFu) | (1u << 16), blo = (bs & 0x3FFFu) | (1u << 16);
              if (PAIR) {
                if (masked) ao::umma_f16_ss_pair_masked(d, alo, blo, hi, idesc, 1u, 0, 0, 0, 0);
                else umma_f16_ss_pair(d, alo, blo, hi, idesc, 1u);
              } else {
                if (masked) ao::umma_f16_ss_lohi_masked(d, alo, blo, hi, idesc, 1u, 0, 0, 0, 0);
                else ao::umma_f16_ss_lohi(d, alo, blo, hi, idesc, 1u);
              }
              continue;
            }
            if (PAIR) {
              if (masked) ao::umma_f16_ss_pair_masked(d, a + j * kAStep, b_lo0 + j * kBStep, desc_hi, idesc, 1u, 0, 0, 0, 0);
              else umma_f16_ss_pair(d, a + j * kAStep, b_lo0 + j * kBStep, desc_hi, idesc, 1u);
            } else {
              if (masked) ao::umma_f16_ss_lohi_masked(d, a + j * kAStep, b_lo0 + j * kBStep, desc_hi, idesc, 1u, 0, 0, 0, 0);
              else ao::umma_f16_ss_lohi(d, a + j * kAStep, b_lo0 + j * kBStep, desc_hi, idesc, 1u);
            }
          }
        }
        if (PAIR) ao::umma_commit_pair(&bar[9 - warp]);
        else ao::umma_commit(&bar[9 - warp]);
      }
      __syncwarp();
      ao::mbar_wait(&bar[9 - warp], 0);
      const long long t1 = clock64();
      *s_stop = 1;
      if (lane == 0 && blockIdx.x == 0 && warp == 9) {
        out[0] = (unsigned long long)(t1 - t0);
        out[1] = (unsigned long long)iters * 8ull * (unsigned long long)n_issuers;
      }
    }
  } else if (warp < 6 && pressure && leader) {
    uint32_t k = 0;
    while (!*s_stop) {
      *reinterpret_cast<uint4*>(s_scr + ((k * 192u + (uint32_t)tid) * 16u) % kRateScratch) = make_uint4(k, k, k, k);
      ++k;
    }
  }
  ao::tc_fence_before_sync();
  __syncthreads();
  if (PAIR) ao::cluster_sync_all();
  if (warp == 9) {
    if (PAIR) ao::tmem_dealloc_pair<512>(tmem);
    else ao::tmem_dealloc<512>(tmem);
  }
}

}  // namespace

// flavour: see above; out3[0] = cycles of CTA 0's issue loop incl. drain, out3[1] = MMAs issued, out3[2] = kernel us
extern "C" int ao_umma_rate(int flavour, int iters, unsigned long long* out2) {
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 16) != cudaSuccess) return -2;
  cudaMemset(d, 0, 16);
  cudaEvent_t ev0, ev1;
  cudaEventCreate(&ev0);
  cudaEventCreate(&ev1);
  cudaEventRecord(ev0, 0);
  const int smem = kRateAct + kRateW + kRateScratch + 64;
  cudaError_t e;
  if (flavour & 2) {
    e = cudaFuncSetAttribute(umma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, umma_rate_kernel<true>, flavour, iters, d);
  } else {
    e = cudaFuncSetAttribute(umma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) umma_rate_kernel<false><<<148, 320, smem>>>(flavour, iters, d);
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  cudaEventRecord(ev1, 0);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out2, d, 16, cudaMemcpyDeviceToHost);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev0, ev1);
  out2[2] = (unsigned long long)(ms * 1000.f);  // kernel time in microseconds (CUDA events)
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  cudaFree(d);
  return e == cudaSuccess ? 0 : -100 - (int)e;
}
