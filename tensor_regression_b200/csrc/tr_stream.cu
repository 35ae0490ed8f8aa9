// tr_stream.cu — instantiations of the two streaming kernels (k_fwd, k_grad), one group of
// channel counts per translation unit (-DTR_PART=0..7) so `make -j` builds them in parallel.
//
// (channels RK, 16-byte chunks per lane E, samples in flight U): registers ~ E*VEC*RK
// coefficients / accumulators + U*E*VEC streamed values; U * next_pow2(RK) <= 32 for the
// transposed warp reduction.  Ranks not listed run on the next larger entry with zero-padded
// channels.
#include "tr_dispatch.h"

#ifndef TR_PART
#error "compile with -DTR_PART=0..7"
#endif

#define TR_DEFINE(NAME, T, ...)                                        \
    const KEntry<T>* NAME(int* n) {                                    \
        static const KEntry<T> t[] = {__VA_ARGS__};                    \
        *n = (int)(sizeof(t) / sizeof(t[0]));                          \
        return t;                                                      \
    }

#if TR_PART == 0
TR_DEFINE(tr_entries_f32_0, float, TR_ENTRY(float, 1, 4, 4), TR_ENTRY(float, 2, 4, 2))
#elif TR_PART == 1
TR_DEFINE(tr_entries_f32_1, float, TR_ENTRY(float, 4, 2, 4), TR_ENTRY(float, 6, 2, 2))
#elif TR_PART == 2
TR_DEFINE(tr_entries_f32_2, float, TR_ENTRY(float, 8, 2, 2))
#elif TR_PART == 3
TR_DEFINE(tr_entries_f32_3, float, TR_ENTRY(float, 12, 1, 2), TR_ENTRY(float, 16, 1, 2))
#elif TR_PART == 4
TR_DEFINE(tr_entries_f64_0, double, TR_ENTRY(double, 1, 4, 4), TR_ENTRY(double, 2, 4, 2))
#elif TR_PART == 5
TR_DEFINE(tr_entries_f64_1, double, TR_ENTRY(double, 4, 2, 2), TR_ENTRY(double, 6, 1, 2))
#elif TR_PART == 6
TR_DEFINE(tr_entries_f64_2, double, TR_ENTRY(double, 8, 1, 2))
#elif TR_PART == 7
TR_DEFINE(tr_entries_f64_3, double, TR_ENTRY(double, 12, 1, 1), TR_ENTRY(double, 16, 1, 1))
#endif
