// tr_fusedmn.cu — instantiations of the single-pass multinomial cluster kernel (tr_fused_mn.cuh), split over
// translation units (-DTR_MPART=0..3) so that `make -j` builds them in parallel.
// (IKC = 16-byte chunks per row of the innermost mode, RKS = channels = rank rounded up to an even number);
// gradient group A takes the first (IKC+1)/2 chunks of a row, group B the rest: each keeps at most 72 (fp32) running sums.
#include "tr_fused_mn.h"

#ifndef TR_MPART
#error "compile with -DTR_MPART=0..3"
#endif

#define TRM_K(T, IKC, RKS) { IKC, RKS, (const void*)k_fused_mn<T, IKC, RKS, (IKC + 1) / 2>, (const void*)trm_c_f3 }

#if TR_MPART == 0
static const TrmEntry tab[] = {TRM_K(float, 5, 6), TRM_K(float, 5, 4), TRM_K(float, 2, 2), TRM_K(float, 3, 2), TRM_K(float, 4, 2),
                               TRM_K(float, 5, 2), TRM_K(float, 6, 2), TRM_K(float, 7, 2), TRM_K(float, 8, 2)};
const TrmEntry* trm_entries_f32_0(int* n) { *n = (int)(sizeof(tab) / sizeof(tab[0])); return tab; }
#elif TR_MPART == 1
static const TrmEntry tab[] = {TRM_K(float, 2, 4), TRM_K(float, 3, 4), TRM_K(float, 4, 4), TRM_K(float, 6, 4), TRM_K(float, 7, 4),
                               TRM_K(float, 8, 4)};
const TrmEntry* trm_entries_f32_1(int* n) { *n = (int)(sizeof(tab) / sizeof(tab[0])); return tab; }
#elif TR_MPART == 2
static const TrmEntry tab[] = {TRM_K(float, 2, 6), TRM_K(float, 3, 6), TRM_K(float, 4, 6), TRM_K(float, 6, 6), TRM_K(float, 2, 8),
                               TRM_K(float, 3, 8), TRM_K(float, 4, 8)};
const TrmEntry* trm_entries_f32_2(int* n) { *n = (int)(sizeof(tab) / sizeof(tab[0])); return tab; }
#elif TR_MPART == 3
static const TrmEntry tab[] = {TRM_K(double, 2, 2), TRM_K(double, 3, 2), TRM_K(double, 4, 2), TRM_K(double, 5, 2), TRM_K(double, 6, 2),
                               TRM_K(double, 2, 4), TRM_K(double, 3, 4), TRM_K(double, 4, 4), TRM_K(double, 5, 4), TRM_K(double, 6, 4)};
const TrmEntry* trm_entries_f64_0(int* n) { *n = (int)(sizeof(tab) / sizeof(tab[0])); return tab; }
#endif
