// tr_specsingle.cu — instantiations of the single-pass spectral kernel (tr_spectral_single.cuh), in their own
// translation unit so that `make -j` builds them next to tr_api.cu.
#define TR_TEMPLATES_ONLY   // the non-template kernels of the shared headers live in tr_api.cu
#include "tr_spectral_single.h"

#define TRSS_CASES(T, VEC)                                           \
    switch (QT) {                                                    \
        case 1: return (const void*)k_spec_single<T, 1, VEC>;        \
        case 2: return (const void*)k_spec_single<T, 2, VEC>;        \
        case 3: return (const void*)k_spec_single<T, 3, VEC>;        \
        case 4: return (const void*)k_spec_single<T, 4, VEC>;        \
        case 5: return (const void*)k_spec_single<T, 5, VEC>;        \
        case 6: return (const void*)k_spec_single<T, 6, VEC>;        \
        case 7: return (const void*)k_spec_single<T, 7, VEC>;        \
        case 8: return (const void*)k_spec_single<T, 8, VEC>;        \
    }                                                                \
    return nullptr;

const void* trss_kernel_f32(int QT) { TRSS_CASES(float, 4) }
const void* trss_kernel_f64(int QT) { TRSS_CASES(double, 2) }
