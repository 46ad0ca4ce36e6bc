/* alpha_omok_b200 - PROBE entry points (test / profiling instrumentation, NOT part of the product library).
 *
 * Built only into alpha_omok_b200/libalpha_omok_b200_probe.so (the product sources compiled with -DAO_PROBE plus
 * csrc/probe/umma_probe.cu).  The probe library also exports every product entry point of alpha_omok_b200.h, with the
 * tower kernels' cycle counters and the AO_TOWER_XFLAGS timing experiments compiled in; libalpha_omok_b200.so contains
 * none of this and never reads AO_TOWER_XFLAGS.
 */
#ifndef ALPHA_OMOK_B200_PROBE_H
#define ALPHA_OMOK_B200_PROBE_H

#include "alpha_omok_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Profiling aid: cycle counters of CTA 0 of the tower kernel (see csrc/engine.cu); enable=1 starts / resets them. */
int ao_tower_debug(ao_engine* h, int enable, uint64_t* out8);

/* tcgen05 building-block probe (csrc/probe/umma_probe.cu). */
int ao_umma_probe(const uint16_t* act_f16, int rows, const uint16_t* wpacked_f16, const float* init, float* out,
                  int row0, int ntaps, const int* shifts);
/* same with per-tap disable-output-lane masks [ntaps][4] (bit r set: output row r is not updated by that tap) */
int ao_umma_probe_masked(const uint16_t* act_f16, int rows, const uint16_t* wpacked_f16, const float* init, float* out,
                         int row0, int ntaps, const int* shifts, const uint32_t* masks);

/* raw tcgen05.mma throughput probe (csrc/probe/umma_probe.cu): cycles of `iters`*8 back-to-back MMAs of one flavour */
int ao_umma_rate(int flavour, int iters, unsigned long long* out2);

#ifdef __cplusplus
}
#endif
#endif /* ALPHA_OMOK_B200_PROBE_H */
