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
