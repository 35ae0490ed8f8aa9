// tr_spectral_single.h — table of k_spec_single instantiations (tr_spectral_single.cuh) shared by tr_api.cu and
// tr_specsingle.cu: the kernel for QT channels, void (*)(SpecSingleArgs<T>), 16-byte lanes (VEC = 16 / sizeof(T)).
#pragma once
#include "tr_spectral_single.cuh"

const void* trss_kernel_f32(int QT);
const void* trss_kernel_f64(int QT);
