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

// tuning knobs for the benchmark shapes (overridable with -D for variant experiments)
#ifndef TR_E1
#define TR_E1 4
#endif
#ifndef TR_UF1
#define TR_UF1 4
#endif
#ifndef TR_UG1
#define TR_UG1 4
#endif
#ifndef TR_E4
#define TR_E4 2
#endif
#ifndef TR_UF4
#define TR_UF4 4
#endif
#ifndef TR_UG4
#define TR_UG4 4
#endif
#ifndef TR_E6
#define TR_E6 2
#endif
#ifndef TR_UF6
#define TR_UF6 4
#endif
#ifndef TR_UG6
#define TR_UG6 3
#endif

#if TR_PART == 0
TR_DEFINE(tr_entries_f32_0, float, TR_ENTRY(float, 1, TR_E1, TR_UF1, TR_UG1), TR_ENTRY(float, 2, 4, 2, 2))
#elif TR_PART == 1
TR_DEFINE(tr_entries_f32_1, float, TR_ENTRY(float, 4, TR_E4, TR_UF4, TR_UG4), TR_ENTRY(float, 6, TR_E6, TR_UF6, TR_UG6))
#elif TR_PART == 2
TR_DEFINE(tr_entries_f32_2, float, TR_ENTRY(float, 8, 2, 2, 2))
#elif TR_PART == 3
TR_DEFINE(tr_entries_f32_3, float, TR_ENTRY(float, 12, 1, 2, 2), TR_ENTRY(float, 16, 1, 2, 2), TR_ENTRY_WIDE(float, 24, 1, 1, 1),
          TR_ENTRY_WIDE(float, 32, 1, 1, 1))
#elif TR_PART == 4
TR_DEFINE(tr_entries_f64_0, double, TR_ENTRY(double, 1, 4, 4, 4), TR_ENTRY(double, 2, 4, 2, 2))
#elif TR_PART == 5
TR_DEFINE(tr_entries_f64_1, double, TR_ENTRY(double, 4, 2, 2, 2), TR_ENTRY(double, 6, 1, 2, 2))
#elif TR_PART == 6
TR_DEFINE(tr_entries_f64_2, double, TR_ENTRY(double, 8, 1, 2, 2))
#elif TR_PART == 7
TR_DEFINE(tr_entries_f64_3, double, TR_ENTRY(double, 12, 1, 1, 1), TR_ENTRY(double, 16, 1, 1, 1), TR_ENTRY_WIDE(double, 24, 1, 1, 1),
          TR_ENTRY_WIDE(double, 32, 1, 1, 1))
#endif
