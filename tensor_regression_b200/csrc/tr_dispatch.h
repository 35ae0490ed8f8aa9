// tr_dispatch.h — table of streaming-kernel instantiations shared by tr_api.cu and tr_stream.cu.
#pragma once
#include "tr_kernels.cuh"
// The experimental single-launch dataflow kernel (tr_flow.cuh, DESIGN 4b: correct, measured 2-6 x slower than the
// two-pass kernels, never auto-selected) is compiled only with -DTR_WITH_FLOW (make FLOW=1).
#ifdef TR_WITH_FLOW
#include "tr_flow.cuh"
#define TR_FLOW_ENTRY(T, RK, E, UF, UG) k_flow<T, RK, E, UF, UG, VN<T>::v>
#else
template <typename T> struct FlowArgs;
#define TR_FLOW_ENTRY(T, RK, E, UF, UG) nullptr
#endif

template <typename T> struct VN;
template <> struct VN<float> { static constexpr int v = 4; };
template <> struct VN<double> { static constexpr int v = 2; };

template <typename T>
struct KEntry {
    int RK, E, Uf, Ug;                    // E in native 16-byte chunks per lane; samples in flight (fwd, grad)
    void (*fwd_vec)(FwdArgs<T>);
    void (*fwd_sc)(FwdArgs<T>);
    void (*grad_vec)(GradArgs<T>);
    void (*grad_sc)(GradArgs<T>);
    void (*flow_vec)(FlowArgs<T>);        // single-launch dataflow kernel (tr_flow.cuh), 16-byte loads only
};

#define TR_ENTRY(T, RK, E, UF, UG)                                                                  \
    { RK, E, UF, UG, k_fwd<T, RK, E, UF, VN<T>::v>, k_fwd<T, RK, E * VN<T>::v, UF, 1>,              \
      k_grad<T, RK, E, UG, VN<T>::v>, k_grad<T, RK, E * VN<T>::v, UG, 1>,                           \
      TR_FLOW_ENTRY(T, RK, E, UF, UG) }

// channel counts whose coefficients / accumulators need more than 128 registers: one resident block per SM
#define TR_ENTRY_WIDE(T, RK, E, UF, UG)                                                             \
    { RK, E, UF, UG, k_fwd<T, RK, E, UF, VN<T>::v, 1>, k_fwd<T, RK, E * VN<T>::v, UF, 1, 1>,        \
      k_grad<T, RK, E, UG, VN<T>::v, 1>, k_grad<T, RK, E * VN<T>::v, UG, 1, 1>, nullptr }

const KEntry<float>* tr_entries_f32_0(int* n);
const KEntry<float>* tr_entries_f32_1(int* n);
const KEntry<float>* tr_entries_f32_2(int* n);
const KEntry<float>* tr_entries_f32_3(int* n);
const KEntry<double>* tr_entries_f64_0(int* n);
const KEntry<double>* tr_entries_f64_1(int* n);
const KEntry<double>* tr_entries_f64_2(int* n);
const KEntry<double>* tr_entries_f64_3(int* n);
