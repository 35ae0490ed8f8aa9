// tr_fused_mn.cuh — single-pass fused forward + gradient kernel for the MULTINOMIAL model (SURVEY H8 ii,
// VERDICT r1 item 3): X is read from HBM once per fit iteration, as in k_fused_std, but the R gradient
// channels do not fit a dense D x R accumulator per cluster, so the CP structure is used on both sides.
//
// Split the feature axis into ROWS x INNER:  inner = the last feature mode (I_k = IKC 16-byte chunks),
// rows = all other modes (NR = D / I_k).  With K[i,r] = F12[row,r] * F3[i_k,r]  (F12 = product of the other
// modes' factor rows, F3 = the last mode's factor):
//
//   forward   t[row,r]  = sum_{i_k} X[n,row,i_k] F3[i_k,r]            (thread-private: a thread owns whole rows)
//             u[n,r]    = sum_row  t[row,r] F12[row,r]                 (warp shuffle + cluster exchange)
//   epilogue  P = softmax(Z), Q = softmax(P), loss, dZ, v[n,r]          (tr_epi.cuh, once per sample)
//   gradient  A[i_k,r]  += X[n,row,i_k] * (v[n,r] F12[row,r])           -> dFt_k   (I_k x R sums per thread)
//             S[row,r]  += v[n,r] * t[row,r]                            -> dFt_m, m < k, by a (k-1)-mode MTTKRP of S
//
// so a cluster keeps I_k*R + NR*R running sums instead of D*R, at R FMA per element in each phase — the same
// arithmetic per element as the two-pass kernels.  t[row,:] stays in shared memory next to the sample until the
// gradient phase has used it.
//
// Build option -DTRM_T_IN_TMEM (make TR_NVCC_EXTRA=-DTRM_T_IN_TMEM): t[row,:] waits in TENSOR MEMORY instead — the
// forward thread parks it with tcgen05.st in its own TMEM lane and the gradient-B thread with the same warp % 4 and
// lane (the only TMEM lanes a warp may touch) fetches it with tcgen05.ld, so the X stage is the only per-sample
// shared-memory buffer and cfg 3 gets 8 stages instead of 6.  Parity-green and MEASURED SLOWER on cfg 3 (6.53 ms
// against 6.00 ms; cfg 5: 24.1-24.6 against 25.0 ms): the tcgen05.st on the forward warps — the role that bounds the
// pipeline — costs more than the two extra stages return (without the stores, wrong results: 5.69 ms; nine x2 stores
// or one x16 + one x2 make no difference).  Kept as an option with its tests; the default build does not use it.
//
// One CLUSTER holds one sample: CTA c owns a contiguous range of rows, NS stages deep.  512 threads per CTA:
//   warps 0-3    forward          (wait full[s]; t -> t stage, thread partials of u -> sUp, arrive redA[s])
//   warps 4-7    gradient A       (chunks [0,QA) of every row -> A;              wait rready[s], arrive empty[s])
//   warps 8-11   gradient B       (chunks [QA,IKC) of every row -> A, and S;     wait rready[s], arrive empty[s])
//   warp 12      TMA producer     (cp.async.bulk of the CTA's rows of sample j into stage j % NS)
//   warp 13      reducer          (sums the 128 forward threads' partials, sends the CTA partial to the sample's OWNER CTA)
//   warp 14      epilogue         (for the samples this CTA owns, i mod CL == rank: sums the CL partials, runs the
//                                  per-sample epilogue in fp64, broadcasts v[n,:] to every CTA of the cluster)
// All hand-offs are mbarriers; partials and v travel through DSMEM with st.async (data + complete_tx together).
// Deterministic: fixed summation orders, no atomics.
#pragma once
#include "tr_fused.cuh"
#include "tr_epi.cuh"

// CTA-local mbarrier waits carry a suspend-time hint (nanoseconds; 0 = plain try_wait polling): the waiting warps are
// parked instead of spending issue slots on the polling loop.  Measured on one box, two rounds: cfg 3 6.18 -> 6.06 ms,
// cfg 5 unchanged (25.5 ms); the same hint on k_fused_std costs 11 % (tr_fused.cuh: TRF_WAIT_NS), so it is set here only.
#ifndef TRM_HINT
#define TRM_HINT 20000
#endif
#if TRM_HINT > 0
#define TRM_WAIT(bar, ph) trf::mbar_wait_hint(bar, ph, TRM_HINT)
#else
#define TRM_WAIT(bar, ph) trf::mbar_wait(bar, ph)
#endif
#define TRM_NWF 4                       // forward warps 0-3 (a fifth one on warp 15 was measured: its sub-partition then hosts
                                        // two forward warps and lags, cfg 3 went from 6.8 to 7.0 ms)
#define TRM_NFT (TRM_NWF * 32)          // forward threads: row of thread t, iteration j = t + j * TRM_NFT
#define TRM_NWG 4
#define TRM_NCT 128
#define TRM_NT 512
#define TRM_GMAX 3                      // rows per thread: a CTA holds at most TRM_GMAX * 128 rows
#define TRM_F12_ROWS (TRM_GMAX * TRM_NFT)   // rows of the zero-padded F12 table in shared memory
#define TRM_MAX_NS 8
#define TRM_MAX_CL 16
#define TRM_QO (TRM_MAX_NS + 2)         // partial slots at the owner (> NS / CL + 1 owned samples can be in flight)
#define TRM_RKMAX 8

template <typename T>
struct FusedMnArgs {
    const T* X;
    const long long* y;
    const T* class_w;
    long long N;
    const T* FtT;            // softplus-ed factors (T)
    const double* Ft64;      // same in double (class factor for the epilogue)
    const T* w;              // rank weights
    Geo geo;
    int NR;                  // rows per sample = D / I_k
    T* Apart;                // (NC * nchunk, CL, TRM_NWG, I_k * RKS)
    T* Spart;                // (NC * nchunk, RKS, NR)
    T* P;                    // (N, C) or null
    T* u_ws;                 // (N, R)
    T* dZ_ws;                // (N, C)
    double* losspart;        // (NC * CL)
    int CL, NC, NS, nchunk;
    long long spc;
    unsigned stage_x_bytes;  // stride of an X stage (>= rows_max * I_k * sizeof(T), multiple of 128)
    unsigned stage_t_bytes;  // stride of a t stage (>= rows_max * RKS * sizeof(T), multiple of 16)
    unsigned head_bytes;     // offset of X stage 0 in the dynamic shared memory
    unsigned piece;          // bytes per bulk-copy instruction
    long long* trace;        // debug timeline (build with -DTRM_TRACE, tools/fused_mn_trace.py), else null
    int dbg;                 // TRM_TRACE builds only: 1 = gradient warps skip their work, 2 = forward warps skip theirs,
                             // 4 = the epilogue skips its math (timing experiments: WRONG results)
};

// The last mode's factor F3 (I_k x RKS values, <= 1 KB) is the one coefficient set every forward thread needs for
// every sample and it is warp-uniform: it lives in CONSTANT memory, so the forward FMAs take it as a constant-bank
// operand instead of 6 broadcast LDS.128 per 16-byte chunk of a row (shared-memory bandwidth is what bounds this
// kernel).  One buffer per translation unit; the host copies F3 into it (stream-ordered, device to device) before
// each launch and orders launches from different streams with an event (tr_api.cu).
#define TRM_F3_MAX (8 * 4 * TRM_RKMAX)          // IKC <= 8 chunks x 4 floats x 8 channels
static __constant__ __align__(16) double trm_c_f3[TRM_F3_MAX];       // holds T values (float kernels use the first half)

template <typename T>
struct FusedMnCtl {
    uint64_t full[TRM_MAX_NS];
    uint64_t empty[TRM_MAX_NS];
    uint64_t redA[TRM_MAX_NS];
    uint64_t rready[TRM_MAX_NS];
    uint64_t cready[TRM_QO];
    uint64_t pfree[2];                           // the reducer is done with slot (i & 1) of the forward threads' partials
    uint32_t tmem_base;                          // tcgen05.alloc result
    uint32_t pad_;
    T vbuf[TRM_MAX_NS][TRM_RKMAX];               // written remotely (owner -> every CTA)
    T cpart[TRM_QO][TRM_MAX_CL][TRM_RKMAX];      // written remotely (every CTA -> owner)
    int dims[TR_MAX_MODES];
    int foff[TR_MAX_MODES + 2];
};

// debug timeline: clock64 stamps of cluster 0 / CTA rank TRM_TRACE_RANK for TRM_TRACE_N samples from TRM_TRACE_I0, kept in
// shared memory during the run (a global store before an mbarrier.arrive would distort the timeline)
#ifndef TRM_TRACE_I0
#define TRM_TRACE_I0 320
#endif
#define TRM_TRACE_N 48
#define TRM_TRACE_EV 16
#ifdef TRM_TRACE
#define TRM_DBG(bit) (a.dbg & (bit))
#define TRM_STAMP(ev, i)                                                                                   \
    do {                                                                                                   \
        if (a.trace && cid == 0 && crank == 0 && (i) >= TRM_TRACE_I0 && (i) < TRM_TRACE_I0 + TRM_TRACE_N)  \
            strace[((i) - TRM_TRACE_I0) * TRM_TRACE_EV + (ev)] = clock64();                                \
    } while (0)
#else
#define TRM_DBG(bit) 0
#define TRM_STAMP(ev, i) do { } while (0)
#endif

namespace trf {
__device__ __forceinline__ void st_async_val(uint32_t remote_addr, float v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void st_async_val(uint32_t remote_addr, double v, uint32_t remote_bar) {
    st_async_f64(remote_addr, v, remote_bar);
}
// acc[r] += x * c[r], r < RKS: FFMA2 over channel pairs for float (same bits as scalar FMAs)
template <typename T, int RKS>
__device__ __forceinline__ void fma_row(T (&acc)[RKS], T x, const T (&c)[RKS]) {
#pragma unroll
    for (int r = 0; r < RKS; ++r) acc[r] = tr_fma<T>(x, c[r], acc[r]);
}
template <int RKS>
__device__ __forceinline__ void fma_row(float (&acc)[RKS], float x, const float (&c)[RKS]) {
#pragma unroll
    for (int r = 0; r + 1 < RKS; r += 2) tr_ffma2(acc[r], acc[r + 1], x, x, c[r], c[r + 1]);
    if (RKS & 1) acc[RKS - 1] = fmaf(x, c[RKS - 1], acc[RKS - 1]);
}
// RKS consecutive T values from shared memory (base 8-byte aligned; RKS even): 8-byte loads
template <typename T, int RKS> struct SVec;
template <int RKS> struct SVec<float, RKS> {
    static __device__ __forceinline__ void ld(const float* p, float (&o)[RKS]) {
#pragma unroll
        for (int r = 0; r < RKS; r += 2) {
            const float2 v = *reinterpret_cast<const float2*>(p + r);
            o[r] = v.x; o[r + 1] = v.y;
        }
    }
    static __device__ __forceinline__ void st(float* p, const float (&o)[RKS]) {
#pragma unroll
        for (int r = 0; r < RKS; r += 2) *reinterpret_cast<float2*>(p + r) = make_float2(o[r], o[r + 1]);
    }
};
template <int RKS> struct SVec<double, RKS> {
    static __device__ __forceinline__ void ld(const double* p, double (&o)[RKS]) {
#pragma unroll
        for (int r = 0; r < RKS; ++r) o[r] = p[r];
    }
    static __device__ __forceinline__ void st(double* p, const double (&o)[RKS]) {
#pragma unroll
        for (int r = 0; r < RKS; ++r) p[r] = o[r];
    }
};
}  // namespace trf

#ifdef TRM_T_IN_TMEM
#define TRM_TMEM 1
#else
#define TRM_TMEM 0
#endif
#define TRM_TMEM_COLS 512
#ifndef TRM_ROTB
#define TRM_ROTB 2                      // row rotation of gradient group B (warps) for its X part
#endif
#ifndef TRM_ROTA
#define TRM_ROTA 2
#endif

namespace trf {
// Tensor memory as per-thread scratch: shape 32x32b = thread i of the warp accesses TMEM lane (32 * (warp % 4) + i),
// xN = N consecutive 32-bit columns.  All tcgen05 instructions are warp-wide (.sync.aligned): call them convergently.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(TRM_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(TRM_TMEM_COLS) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}
// wide variants: one instruction moves 4 / 8 / 16 consecutive columns (a tcgen05.st / ld costs about the same issue
// time whatever its width: measured, nine x2 stores per sample cost the forward warps 15 % of their time)
__device__ __forceinline__ void tmem_st4(uint32_t t, const uint32_t* w) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t t, const uint32_t* w) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(t), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t t, const uint32_t* w) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(t), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
                   "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t t, uint32_t* w) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(t) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t t, uint32_t* w) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(t) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t t, uint32_t* w) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]),
                   "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "r"(t) : "memory");
}
// N (even) consecutive columns of the calling thread's lane, in as few instructions as the power-of-two widths allow
template <int N>
__device__ __forceinline__ void tmem_st_words(uint32_t t, const uint32_t* w) {
    static_assert(N >= 0 && N % 2 == 0, "even word counts only");
    if constexpr (N >= 16) { tmem_st16(t, w); tmem_st_words<N - 16>(t + 16, w + 16); }
    else if constexpr (N >= 8) { tmem_st8(t, w); tmem_st_words<N - 8>(t + 8, w + 8); }
    else if constexpr (N >= 4) { tmem_st4(t, w); tmem_st_words<N - 4>(t + 4, w + 4); }
    else if constexpr (N >= 2) { tmem_st2(t, w[0], w[1]); }
}
template <int N>
__device__ __forceinline__ void tmem_ld_words(uint32_t t, uint32_t* w) {
    static_assert(N >= 0 && N % 2 == 0, "even word counts only");
    if constexpr (N >= 16) { tmem_ld16(t, w); tmem_ld_words<N - 16>(t + 16, w + 16); }
    else if constexpr (N >= 8) { tmem_ld8(t, w); tmem_ld_words<N - 8>(t + 8, w + 8); }
    else if constexpr (N >= 4) { tmem_ld4(t, w); tmem_ld_words<N - 4>(t + 4, w + 4); }
    else if constexpr (N >= 2) { tmem_ld2(t, w[0], w[1]); }
}
__device__ __forceinline__ void tmem_tie(uint32_t& x) { asm volatile("" : "+r"(x)); }
// bit casts between NV values of T and 32-bit words
template <int NV> __device__ __forceinline__ void to_words(const float (&v)[NV], uint32_t (&w)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) w[i] = __float_as_uint(v[i]);
}
template <int NV> __device__ __forceinline__ void to_words(const double (&v)[NV], uint32_t (&w)[2 * NV]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const long long b = __double_as_longlong(v[i]);
        w[2 * i] = (uint32_t)b; w[2 * i + 1] = (uint32_t)(b >> 32);
    }
}
__device__ __forceinline__ float from_words(const uint32_t* w, float) { return __uint_as_float(w[0]); }
__device__ __forceinline__ double from_words(const uint32_t* w, double) {
    return __longlong_as_double((long long)(((unsigned long long)w[1] << 32) | w[0]));
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers written by tcgen05.ld are defined only after tcgen05.wait::ld: re-defining them in an (empty) volatile asm
// placed after the wait keeps the compiler from scheduling a consumer above it
__device__ __forceinline__ void tmem_tie(float& x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ void tmem_tie(double& x) { asm volatile("" : "+d"(x)); }
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// RKS values of T <-> consecutive TMEM columns of the calling thread's lane (RKS even; a double takes two columns)
template <typename T, int RKS> struct TVec;
template <int RKS> struct TVec<float, RKS> {
    static constexpr int COLS = RKS;
    static __device__ __forceinline__ void st(uint32_t taddr, const float (&o)[RKS]) {
#pragma unroll
        for (int r = 0; r < RKS; r += 2) tmem_st2(taddr + r, __float_as_uint(o[r]), __float_as_uint(o[r + 1]));
    }
    static __device__ __forceinline__ void ld(uint32_t taddr, float (&o)[RKS]) {
#pragma unroll
        for (int r = 0; r < RKS; r += 2) {
            uint32_t a, b;
            tmem_ld2(taddr + r, a, b);
            o[r] = __uint_as_float(a); o[r + 1] = __uint_as_float(b);
        }
    }
};
template <int RKS> struct TVec<double, RKS> {
    static constexpr int COLS = 2 * RKS;
    static __device__ __forceinline__ void st(uint32_t taddr, const double (&o)[RKS]) {
#pragma unroll
        for (int r = 0; r < RKS; ++r) {
            const long long b = __double_as_longlong(o[r]);
            tmem_st2(taddr + 2 * r, (uint32_t)b, (uint32_t)(b >> 32));
        }
    }
    static __device__ __forceinline__ void ld(uint32_t taddr, double (&o)[RKS]) {
#pragma unroll
        for (int r = 0; r < RKS; ++r) {
            uint32_t a, b;
            tmem_ld2(taddr + 2 * r, a, b);
            o[r] = __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
        }
    }
};
}  // namespace trf

template <typename T> __device__ __forceinline__ T trm_exp(T x);
template <> __device__ __forceinline__ float trm_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double trm_exp<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T trm_log(T x);
template <> __device__ __forceinline__ float trm_log<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double trm_log<double>(double x) { return log(x); }

// Per-sample epilogue of the single-pass kernel, one warp, lane = class (C <= 32), arithmetic in the model dtype T
// like the reference's own (mn:180-187 softmax, mn:364-366 / 448-450 CrossEntropyLoss on the probabilities = second
// softmax, autograd down to v[n,:]).  It sits on the critical path of the cluster pipeline (the sample stays in shared
// memory until v is known), so it is written for latency: reductions over the next power of two >= C lanes only, no
// max-subtraction in the second softmax (its arguments are probabilities in [0,1]; exp(P)/sum exp(P) is the same
// quotient), one table sFCw[c][r] = w_r * FC[c,r] for both the logits and v.  The fp64 epilogue of the two-pass path
// (epi_mn_core) agrees with it to a few ulp of T; both are checked against the oracle at the north-star tolerance.
// a / b for b > 0: one reciprocal + one multiply for float (2 ulp; the IEEE sequence has a divergent slow path)
__device__ __forceinline__ float trm_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double trm_div(double a, double b) { return a / b; }

template <typename T, int RKS, int STEPS>
__device__ __forceinline__ void trm_epilogue(const EpiMnArgs<T>& a, long long n, const T (&u)[RKS], int lane, const T* sFCw,
                                             int yn, T omega, T (&vout)[RKS], T& P_out, T& qs_out, T& dZ_out) {
    // reductions run over groups of 2^STEPS lanes (2^STEPS >= C, fully unrolled butterflies): every lane of group 0 —
    // the classes, and the lanes that send v to the peers when STEPS >= 4 — ends with the group totals; the other
    // groups (STEPS = 4 only) hold inactive lanes whose sums are patched to 1 so that nothing divides by zero
    const int C = a.C, R = a.R;
    const bool act = lane < C;
    T fc[RKS];
    trf::SVec<T, RKS>::ld(sFCw + (size_t)(act ? lane : 0) * RKS, fc);
    T z = (T)0;
#pragma unroll
    for (int r = 0; r < RKS; ++r) z = tr_fma<T>(u[r], fc[r], z);
    T zmax = act ? z : -INFINITY;
#pragma unroll
    for (int st = STEPS - 1; st >= 0; --st) zmax = fmax(zmax, __shfl_xor_sync(TR_FULL, zmax, 1 << st));
    const T e = act ? trm_exp<T>(z - zmax) : (T)0;
    T zs = e;
#pragma unroll
    for (int st = STEPS - 1; st >= 0; --st) zs += __shfl_xor_sync(TR_FULL, zs, 1 << st);
    const T P = trm_div(e, zs > (T)0 ? zs : (T)1);
    const T qe = act ? trm_exp<T>(P) : (T)0;
    T qs = qe;
#pragma unroll
    for (int st = STEPS - 1; st >= 0; --st) qs += __shfl_xor_sync(TR_FULL, qs, 1 << st);
    qs = qs > (T)0 ? qs : (T)1;
    const T q = trm_div(qe, qs);
    const T dP = act ? omega * (q - (lane == yn ? (T)1 : (T)0)) : (T)0;
    T dot = dP * P;
#pragma unroll
    for (int st = STEPS - 1; st >= 0; --st) dot += __shfl_xor_sync(TR_FULL, dot, 1 << st);
    const T dZ = act ? P * (dP - dot) : (T)0;
    // v[r] = w_r sum_c dZ[c] FC[c,r]: RKS independent butterflies
#pragma unroll
    for (int r = 0; r < RKS; ++r) vout[r] = dZ * fc[r];
#pragma unroll
    for (int st = STEPS - 1; st >= 0; --st) {
#pragma unroll
        for (int r = 0; r < RKS; ++r) vout[r] += __shfl_xor_sync(TR_FULL, vout[r], 1 << st);
    }
    P_out = P; qs_out = qs; dZ_out = dZ;
}

// What the gradient warps do not wait for: outputs and the loss term, after v[n,:] is on its way to the peers
template <typename T, int RKS>
__device__ __forceinline__ void trm_epilogue_tail(const EpiMnArgs<T>& a, long long n, const T (&u)[RKS], int lane, int yn,
                                                  T omega, T P, T qs, T dZ, double& loss) {
    const int C = a.C, R = a.R;
    const bool act = lane < C;
    if (act && a.P) a.P[n * C + lane] = P;
    if (act && a.dZ_ws) a.dZ_ws[n * C + lane] = dZ;
    if (act && lane == yn) loss += (double)(-omega * (P - trm_log<T>(qs)));          // -omega log Q[n,y_n]
    if (lane < R && a.u_ws) {
        T ur = (T)0;
#pragma unroll
        for (int r = 0; r < RKS; ++r) if (r == lane) ur = u[r];
        a.u_ws[n * R + lane] = ur;
    }
}

// S[row,r] += v[r] t[row,r] over the rows of the FORWARD thread that shares this thread's TMEM lane (threadIdx & 127 +
// j * 128; njs row iterations, warp-uniform), whatever rotation the X part of the gradient group uses
template <typename T, int RKS, int NJS>
__device__ __forceinline__ void trm_s_rows(T (&S)[TRM_GMAX][RKS], const T (&v)[RKS], uint32_t tt) {
#if TRM_TMEM
    constexpr int WPT = (int)sizeof(T) / 4;
    uint32_t w[NJS * RKS * WPT];
    trf::tmem_ld_words<NJS * RKS * WPT>(tt, w);
    trf::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < NJS * RKS * WPT; ++i) trf::tmem_tie(w[i]);
#pragma unroll
    for (int j = 0; j < NJS; ++j)
#pragma unroll
        for (int r = 0; r < RKS; ++r) S[j][r] = tr_fma<T>(v[r], trf::from_words(&w[(j * RKS + r) * WPT], (T)0), S[j][r]);
#endif
}
template <typename T, int RKS>
__device__ __forceinline__ void trm_s_only(T (&S)[TRM_GMAX][RKS], const T (&v)[RKS], uint32_t tt, int njs) {
    static_assert(TRM_GMAX == 3, "row iterations are dispatched explicitly");
    if (njs == 3) trm_s_rows<T, RKS, 3>(S, v, tt);
    else if (njs == 2) trm_s_rows<T, RKS, 2>(S, v, tt);
    else if (njs == 1) trm_s_rows<T, RKS, 1>(S, v, tt);
}

// One gradient group: chunks [Q0, Q0 + QN) of every row of this CTA (and the row sums S when WITH_S).
// One sample of the forward role for a warp with NJ row iterations (rows tid + j * TRM_NFT): t[row,:] to shared memory,
// the lane's partial of u[n,:] in vals.  Branch-free like trm_gradient_sample.
template <typename T, int IKC, int RKS, int NJ>
__device__ __forceinline__ void trm_forward_sample(T (&vals)[TRM_RKMAX], const T* cF3, const T (&f12)[TRM_GMAX][RKS], const T* xs,
                                                   T* ts, uint32_t tt, int tid, int nrows, const int (&rlc)[TRM_GMAX], bool store_t) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int IK = IKC * VEC;
    T t[NJ][RKS];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int r = 0; r < RKS; ++r) t[j][r] = (T)0;
#pragma unroll
    for (int q = 0; q < IKC; ++q) {
        T f3[VEC][RKS];                                    // F3 rows of this chunk: constant-bank operands
#pragma unroll
        for (int vv = 0; vv < VEC; ++vv)
#pragma unroll
            for (int r = 0; r < RKS; ++r) f3[vv][r] = cF3[(q * VEC + vv) * RKS + r];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            T x[VEC];
            trf::SLoad<T, VEC>::ld(xs + (size_t)rlc[j] * IK + q * VEC, x);
#pragma unroll
            for (int vv = 0; vv < VEC; ++vv) trf::fma_row(t[j], x[vv], f3[vv]);
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#if !TRM_TMEM
        if (tid + j * TRM_NFT < nrows && store_t) trf::SVec<T, RKS>::st(ts + (size_t)(tid + j * TRM_NFT) * RKS, t[j]);
#endif
#pragma unroll
        for (int r = 0; r < RKS; ++r) vals[r] = tr_fma<T>(t[j][r], f12[j][r], vals[r]);     // f12 = 0 past nrows
    }
#if TRM_TMEM
    {
        // t of the thread's NJ rows -> NJ * RKS consecutive TMEM columns of its lane, in one or two wide stores; every
        // lane stores (warp-wide instruction), rows past nrows are never read back
        constexpr int WPT = (int)sizeof(T) / 4;
        T flat[NJ * RKS];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int r = 0; r < RKS; ++r) flat[j * RKS + r] = t[j][r];
        uint32_t w[NJ * RKS * WPT];
        trf::to_words<NJ * RKS>(flat, w);
        trf::tmem_st_words<NJ * RKS * WPT>(tt, w);
    }
#endif
}

// One sample of a gradient group for a warp with NJ row iterations.  No branches: lanes whose row of iteration j lies
// past nrows read this CTA's last row (always written by the TMA) against F12 = 0, i.e. contribute exact zeros, so all
// shared-memory loads of the sample can be issued ahead of the FMAs (one warp per role and sub-partition: nothing else
// hides the LDS latency).
template <typename T, int IKC, int RKS, int Q0, int QN, bool WITH_S, int NJ>
__device__ __forceinline__ void trm_gradient_sample(T (&acc)[QN * (16 / (int)sizeof(T))][RKS], T (&S)[WITH_S ? TRM_GMAX : 1][RKS],
                                                    const T (&v)[RKS], const T* sF12, const T* xs, const T* ts, uint32_t tt, int tb,
                                                    const int (&rlc)[TRM_GMAX], int njs) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int IK = IKC * VEC;
    T c[NJ][RKS];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        trf::SVec<T, RKS>::ld(sF12 + (size_t)(tb + j * TRM_NCT) * RKS, c[j]);
#pragma unroll
        for (int r = 0; r < RKS; ++r) c[j][r] *= v[r];
    }
#if TRM_TMEM

#else
    if (WITH_S) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            T tv[RKS];
            trf::SVec<T, RKS>::ld(ts + (size_t)rlc[j] * RKS, tv);
#pragma unroll
            for (int r = 0; r < RKS; ++r) S[j][r] = tr_fma<T>(v[r], tv[r], S[j][r]);
        }
    }
#endif
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#pragma unroll
        for (int qq = 0; qq < QN; ++qq) {
            T x[VEC];
            trf::SLoad<T, VEC>::ld(xs + (size_t)rlc[j] * IK + (Q0 + qq) * VEC, x);
#pragma unroll
            for (int vv = 0; vv < VEC; ++vv) trf::fma_row(acc[qq * VEC + vv], x[vv], c[j]);
        }
    }
#if TRM_TMEM
    // S runs over the rows of the FORWARD thread with this thread's TMEM lane (threadIdx & 127 + j * 128; njs row
    // iterations, warp-uniform), whatever rotation the X part above uses: t[row,:] comes from that lane with
    // tcgen05.ld.  After the X loop, when the registers of c and x are free again.
    if constexpr (WITH_S) trm_s_only<T, RKS>(S, v, tt, njs);
#endif
}

// ROT rotates the thread -> row assignment by ROT warps: with nrows between 2 and 3 rows per thread the first warps of
// a role carry one row more than the last ones; rotating the gradient roles puts their heavy warps on the
// sub-partitions where the forward role has its light ones (warp w issues on sub-partition w % 4).
template <typename T, int IKC, int RKS, int Q0, int QN, bool WITH_S, int ROT>
__device__ __forceinline__ void trm_gradient_role(const FusedMnArgs<T>& a, FusedMnCtl<T>* ctl, const T* sF12,
                                                  const unsigned char* stageX0, const unsigned char* stageT0,
                                                  int cnt, int cid, unsigned crank, int nrows, int row0, long long* strace) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int IK = IKC * VEC;
    const int lane = threadIdx.x & 31;
    // TMEM address of this thread's lane quarter (lanes 32 * (warp % 4) ..): bits 31..16 lane, 15..0 column
    const uint32_t tmem_lane = ctl->tmem_base + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16);
    const int gw = (threadIdx.x & (TRM_NCT - 1)) >> 5;                          // physical warp of the role: flush slot
    const int tb = (threadIdx.x + ROT * 32) & (TRM_NCT - 1);                    // role-relative thread id for the row assignment
    const int NS = a.NS;
    T acc[QN * VEC][RKS];
#pragma unroll
    for (int e = 0; e < QN * VEC; ++e)
#pragma unroll
        for (int r = 0; r < RKS; ++r) acc[e][r] = (T)0;
    T S[WITH_S ? TRM_GMAX : 1][RKS];
#pragma unroll
    for (int j = 0; j < (WITH_S ? TRM_GMAX : 1); ++j)
#pragma unroll
        for (int r = 0; r < RKS; ++r) S[j][r] = (T)0;

    int rlc[TRM_GMAX];
#pragma unroll
    for (int j = 0; j < TRM_GMAX; ++j) rlc[j] = min(tb + j * TRM_NCT, nrows - 1);
    int nj = 0;
#pragma unroll
    for (int j = 0; j < TRM_GMAX; ++j) nj += ((tb & ~31) + j * TRM_NCT < nrows) ? 1 : 0;
    // rows of the S sums: with t in TMEM those of the forward thread that shares this thread's TMEM lane
    const int ts_id = TRM_TMEM ? (int)(threadIdx.x & (TRM_NCT - 1)) : tb;
    int njs = 0;
#pragma unroll
    for (int j = 0; j < TRM_GMAX; ++j) njs += ((ts_id & ~31) + j * TRM_NCT < nrows) ? 1 : 0;
    int s = 0;
    unsigned ph = 0;
    int left = (int)(a.spc < cnt ? a.spc : cnt);
    int chunk_idx = 0;
    for (int i = 0; i < cnt; ++i) {
        TRM_WAIT(&ctl->rready[s], ph);                       // v[n,:] of this sample arrived (all lanes wait)
        if (WITH_S) TRM_WAIT(&ctl->redA[s], ph);             // acquire the forward warps' t[row,:] stores
        __syncwarp();
#if TRM_TMEM
        if (WITH_S) trf::tmem_fence_after();
        const uint32_t tt = tmem_lane + (uint32_t)(s * TRM_GMAX) * trf::TVec<T, RKS>::COLS;
#else
        const uint32_t tt = 0;
#endif
        if (tb == 0) TRM_STAMP(WITH_S ? 10 : 8, i);
        T v[RKS];
        trf::SVec<T, RKS>::ld(&ctl->vbuf[s][0], v);
        const T* xs = reinterpret_cast<const T*>(stageX0 + (size_t)s * a.stage_x_bytes);
        const T* ts = reinterpret_cast<const T*>(stageT0 + (size_t)s * a.stage_t_bytes);
        // No branches in the row loop: rows past nrows read this CTA's last row (always written by the TMA, every CTA
        // has at least one row) against F12 = 0, i.e. contribute exact zeros, so that all shared-memory loads of a
        // sample can be issued ahead of the FMAs (one warp per role and sub-partition: nothing else hides LDS latency).
        // row iterations this WARP really has (warp-uniform): a warp whose rows of iteration j all lie past nrows skips it
        if (!TRM_DBG(1)) {
            if (nj == 3) trm_gradient_sample<T, IKC, RKS, Q0, QN, WITH_S, 3>(acc, S, v, sF12, xs, ts, tt, tb, rlc, njs);
            else if (nj == 2) trm_gradient_sample<T, IKC, RKS, Q0, QN, WITH_S, 2>(acc, S, v, sF12, xs, ts, tt, tb, rlc, njs);
            else if (nj == 1) trm_gradient_sample<T, IKC, RKS, Q0, QN, WITH_S, 1>(acc, S, v, sF12, xs, ts, tt, tb, rlc, njs);
            else if constexpr (WITH_S) { if (njs > 0) trm_s_only<T, RKS>(S, v, tt, njs); }
        }
        __syncwarp();
        if (tb == 0) TRM_STAMP(WITH_S ? 11 : 9, i);
#if TRM_TMEM
        if (WITH_S) trf::tmem_fence_before();                      // the TMEM reads above are ordered before the stage's release
#endif
        if (lane == 0) trf::mbar_arrive(&ctl->empty[s]);           // this warp's reads of the stage are done
        if (++s == NS) { s = 0; ph ^= 1u; }
        if (--left == 0) {
            // chunk boundary: flush the running sums (bounds the length of every fp32 sum); lanes hold different rows
            const size_t slot = (size_t)cid * a.nchunk + chunk_idx;
            T* ap = a.Apart + ((slot * a.CL + crank) * TRM_NWG + gw) * (size_t)(IK * RKS);
#pragma unroll
            for (int e = 0; e < QN * VEC; ++e)
#pragma unroll
                for (int r = 0; r < RKS; ++r) {
                    T sum = acc[e][r];
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(TR_FULL, sum, off);
                    if (lane == 0) ap[(Q0 * VEC + e) * RKS + r] = sum;
                    acc[e][r] = (T)0;
                }
            if (WITH_S) {
                T* sp = a.Spart + slot * (size_t)RKS * a.NR;
#pragma unroll
                for (int j = 0; j < TRM_GMAX; ++j) {
                    const int rl = ts_id + j * TRM_NCT;
                    if (rl < nrows) {
#pragma unroll
                        for (int r = 0; r < RKS; ++r) {
                            sp[(size_t)r * a.NR + row0 + rl] = S[j][r];
                            S[j][r] = (T)0;
                        }
                    }
                }
            }
            ++chunk_idx;
            const int rem = cnt - (i + 1);
            left = (int)(a.spc < rem ? a.spc : rem);
        }
    }
}

template <typename T, int IKC, int RKS, int QA>
__global__ void __launch_bounds__(TRM_NT, 1) k_fused_mn(const FusedMnArgs<T> a) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int IK = IKC * VEC;
    static_assert(RKS % 2 == 0 && RKS <= TRM_RKMAX, "channel count must be even and <= 8");
    static_assert(QA >= 1 && QA < IKC, "chunk split: both gradient groups need at least one chunk of a row");
    extern __shared__ __align__(128) unsigned char trm_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned crank = trf::cluster_ctarank();
    const int cid = blockIdx.x / a.CL;
    const int NS = a.NS, CL = a.CL;
    const int k = a.geo.k, R = a.geo.R, C = a.geo.C;

    // layout: [control | F12 (GMAX*128, RKS) | w * class factor (C, RKS), class weights (C) | pad | X stages | t stages]
    FusedMnCtl<T>* ctl = reinterpret_cast<FusedMnCtl<T>*>(trm_smem);
    size_t off = (sizeof(FusedMnCtl<T>) + 15) / 16 * 16;
    T* sF12 = reinterpret_cast<T*>(trm_smem + off);
    off += (size_t)TRM_F12_ROWS * RKS * sizeof(T);
    off = (off + 15) / 16 * 16;
    T* sFCw = reinterpret_cast<T*>(trm_smem + off);                 // (C, RKS): w_r * FC[c,r], zero-padded channels
    T* sCW = sFCw + (size_t)C * RKS;                                // (C): class weights omega[c]
    // (2, 128, RKS): the forward threads' partials of u[n,:] of samples i and i+1 (summed by the reducer warp)
    T* sUp = reinterpret_cast<T*>(trm_smem + (off + (size_t)C * (RKS + 1) * sizeof(T) + 15) / 16 * 16);
    const unsigned char* stageX0 = trm_smem + a.head_bytes;
    const unsigned char* stageT0 = stageX0 + (size_t)NS * a.stage_x_bytes;
#ifdef TRM_TRACE
    long long* strace = reinterpret_cast<long long*>(const_cast<unsigned char*>(stageT0) + (size_t)NS * a.stage_t_bytes);
    for (int i = threadIdx.x; i < TRM_TRACE_N * TRM_TRACE_EV; i += TRM_NT) strace[i] = 0;
#else
    long long* strace = nullptr;
#endif

    // rows of this CTA: the first (NR mod CL) ranks hold one more
    const int rq = a.NR / CL, rrem = a.NR % CL;
    const int nrows = rq + ((int)crank < rrem ? 1 : 0);
    const int row0 = (int)crank * rq + ((int)crank < rrem ? (int)crank : rrem);

    if (threadIdx.x < TR_MAX_MODES) ctl->dims[threadIdx.x] = a.geo.dims[threadIdx.x];
    if (threadIdx.x < TR_MAX_MODES + 2) ctl->foff[threadIdx.x] = a.geo.foff[threadIdx.x];
    for (int idx = threadIdx.x; idx < TRM_F12_ROWS * RKS; idx += TRM_NT) {
        const int rl = idx / RKS, r = idx % RKS;
        T p = (T)0;
        if (rl < nrows && r < R) {
            p = (T)1;
            unsigned rem = (unsigned)(row0 + rl);
            for (int m = k - 2; m >= 0; --m) {
                const unsigned d = (unsigned)a.geo.dims[m];
                const unsigned q = rem / d;
                p *= a.FtT[a.geo.foff[m] + (int)(rem - q * d) * R + r];
                rem = q;
            }
        }
        sF12[idx] = p;
    }
    for (int idx = threadIdx.x; idx < C * RKS; idx += TRM_NT) {
        const int c = idx / RKS, r = idx % RKS;
        sFCw[idx] = r < R ? (T)((double)a.w[r] * a.Ft64[a.geo.pfeat + c * R + r]) : (T)0;
    }
    for (int idx = threadIdx.x; idx < C; idx += TRM_NT) sCW[idx] = a.class_w ? a.class_w[idx] : (T)1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            trf::mbar_init(&ctl->full[s], 1);
            trf::mbar_init(&ctl->empty[s], 2 * TRM_NWG);
            trf::mbar_init(&ctl->redA[s], TRM_NWF);
            trf::mbar_init(&ctl->rready[s], 1);
        }
        for (int q = 0; q < TRM_QO; ++q) trf::mbar_init(&ctl->cready[q], 1);
        trf::mbar_init(&ctl->pfree[0], 1);
        trf::mbar_init(&ctl->pfree[1], 1);
        trf::fence_mbar_init();
    }
#if TRM_TMEM
    static_assert(TRM_MAX_NS * TRM_GMAX * TRM_RKMAX * (int)(sizeof(T) / 4) <= TRM_TMEM_COLS, "t does not fit the TMEM columns");
    if (warp == TRM_NT / 32 - 1) trf::tmem_alloc(&ctl->tmem_base);      // one warp allocates (1 CTA per SM: never contended)
    trf::tmem_fence_before();
#endif
    __syncthreads();
#if TRM_TMEM
    trf::tmem_fence_after();
#endif
    trf::cluster_arrive();
    trf::cluster_wait();

    // samples of this cluster: n = cid + i*NC, i = 0..cnt-1; sample i is OWNED by CTA (i mod CL)
    const int cnt = cid < a.N ? (int)((a.N - cid + a.NC - 1) / a.NC) : 0;
    const size_t sample_stride = (size_t)a.NC * (size_t)a.geo.D;

    if (warp < TRM_NWF) {
        // ===================================== forward =====================================
        const int fw = warp;
        const int tid = fw * 32 + lane;                           // 0..127
        const T* cF3 = reinterpret_cast<const T*>(trm_c_f3);
        T f12[TRM_GMAX][RKS];
#pragma unroll
        for (int j = 0; j < TRM_GMAX; ++j) trf::SVec<T, RKS>::ld(sF12 + (size_t)(tid + j * TRM_NFT) * RKS, f12[j]);   // zeros past nrows
        int rlc[TRM_GMAX];
#pragma unroll
        for (int j = 0; j < TRM_GMAX; ++j) rlc[j] = min(tid + j * TRM_NFT, nrows - 1);
        int nj = 0;                                               // row iterations of this warp (warp-uniform)
#pragma unroll
        for (int j = 0; j < TRM_GMAX; ++j) nj += (fw * 32 + j * TRM_NFT < nrows) ? 1 : 0;
        int s = 0;
        unsigned ph = 0;
        for (int i = 0; i < cnt; ++i) {
            TRM_WAIT(&ctl->full[s], ph);                     // every lane waits on the barrier itself (tr_fused.cuh)
            // slot i & 1 of the partial buffer was last used by sample i - 2: its use number is (i >> 1) - 1
            if (i >= 2) TRM_WAIT(&ctl->pfree[i & 1], (unsigned)(((i >> 1) - 1) & 1));
            __syncwarp();
            if (tid == 0) TRM_STAMP(1, i);
            const T* xs = reinterpret_cast<const T*>(stageX0 + (size_t)s * a.stage_x_bytes);
            T* ts = reinterpret_cast<T*>(const_cast<unsigned char*>(stageT0) + (size_t)s * a.stage_t_bytes);
#if TRM_TMEM
            const uint32_t tt = ctl->tmem_base + ((uint32_t)(fw * 32) << 16) + (uint32_t)(s * TRM_GMAX) * trf::TVec<T, RKS>::COLS;
#else
            const uint32_t tt = 0;
#endif
            T vals[TRM_RKMAX];
#pragma unroll
            for (int r = 0; r < TRM_RKMAX; ++r) vals[r] = (T)0;
            if (!TRM_DBG(2)) {
                if (nj == 3) trm_forward_sample<T, IKC, RKS, 3>(vals, cF3, f12, xs, ts, tt, tid, nrows, rlc, !TRM_DBG(16));
                else if (nj == 2) trm_forward_sample<T, IKC, RKS, 2>(vals, cF3, f12, xs, ts, tt, tid, nrows, rlc, !TRM_DBG(16));
                else if (nj == 1) trm_forward_sample<T, IKC, RKS, 1>(vals, cF3, f12, xs, ts, tt, tid, nrows, rlc, !TRM_DBG(16));
            }
            if (tid == 0) TRM_STAMP(13, i);
            // the thread's partial of u[n,:] goes to shared memory as it is: the cross-thread sum is the reducer warp's
            // job (a shuffle reduction here would sit on the forward warps, the busiest role of the pipeline)
            {
                T pv[RKS];
#pragma unroll
                for (int r = 0; r < RKS; ++r) pv[r] = vals[r];
                trf::SVec<T, RKS>::st(sUp + ((size_t)(i & 1) * TRM_NFT + tid) * RKS, pv);
            }
#if TRM_TMEM
            trf::tmem_wait_st();                                   // t is in TMEM ...
            trf::tmem_fence_before();                              // ... and ordered before the arrival below
#endif
            __syncwarp();
            if (tid == 0) TRM_STAMP(2, i);
            if (lane == 0) trf::mbar_arrive(&ctl->redA[s]);
            if (++s == NS) { s = 0; ph ^= 1u; }
        }
    } else if (warp < 4 + TRM_NWG) {
        trm_gradient_role<T, IKC, RKS, 0, QA, false, TRM_ROTA>(a, ctl, sF12, stageX0, stageT0, cnt, cid, crank, nrows, row0, strace);
    } else if (warp < 4 + 2 * TRM_NWG) {
        // group B: X part on rotated rows like group A; its S sums run over the rows of the forward thread it shares TMEM lanes with
        trm_gradient_role<T, IKC, RKS, QA, IKC - QA, true, TRM_ROTB>(a, ctl, sF12, stageX0, stageT0, cnt, cid, crank, nrows, row0, strace);
    } else if (warp == 4 + 2 * TRM_NWG) {
        // ================================== TMA producer ==================================
        if (lane == 0) {
            const unsigned my_bytes = (unsigned)nrows * (unsigned)IK * (unsigned)sizeof(T);
            const T* src = a.X + (size_t)cid * (size_t)a.geo.D + (size_t)row0 * IK;
            int s = 0;
            unsigned ph = 0;
            for (int j = 0; j < cnt; ++j) {
                if (j >= NS) TRM_WAIT(&ctl->empty[s], ph);
                TRM_STAMP(0, j);
                trf::mbar_arrive_expect_tx(&ctl->full[s], my_bytes);
                const unsigned char* sp = reinterpret_cast<const unsigned char*>(src);
                unsigned char* dst = const_cast<unsigned char*>(stageX0) + (size_t)s * a.stage_x_bytes;
                for (unsigned o2 = 0; o2 < my_bytes; o2 += a.piece) {
                    const unsigned len = my_bytes - o2 < a.piece ? my_bytes - o2 : a.piece;
                    trf::bulk_g2s(dst + o2, sp + o2, len, &ctl->full[s]);
                }
                src += sample_stride;
                if (++s == NS) { s = 0; if (j >= NS) ph ^= 1u; }
            }
        }
    } else if (warp == 4 + 2 * TRM_NWG + 1) {
        // ===================== reducer: CTA partial of u[n,:] -> owner CTA =====================
        int s = 0, owner = 0, slot = 0;
        unsigned ph = 0;
        for (int i = 0; i < cnt; ++i) {
            TRM_WAIT(&ctl->redA[s], ph);
            __syncwarp();
            if (lane == 0) TRM_STAMP(3, i);
            // the previous use of rready[s] (sample i - NS) has completed: its gradient phase released the stage
            // before this sample could be loaded.  Arm it for v[n,:] of this sample.
            if (lane == 0) trf::mbar_arrive_expect_tx(&ctl->rready[s], (unsigned)(RKS * sizeof(T)));
            {
                // lane l sums the partials of forward threads l, l + 32, l + 64, l + 96 (fixed order), then a transposed
                // butterfly leaves the total of channel c in lane 4 c
                T pc[TRM_RKMAX];
#pragma unroll
                for (int r = 0; r < TRM_RKMAX; ++r) pc[r] = (T)0;
                const T* up = sUp + ((size_t)(i & 1) * TRM_NFT + lane) * RKS;
#pragma unroll
                for (int w4 = 0; w4 < TRM_NWF; ++w4) {
                    T pv[RKS];
                    trf::SVec<T, RKS>::ld(up + (size_t)w4 * 32 * RKS, pv);
#pragma unroll
                    for (int r = 0; r < RKS; ++r) pc[r] += pv[r];
                }
                warp_reduce_transpose<T, TRM_RKMAX, 0>(pc, lane);
                if ((lane & 3) == 0 && (lane >> 2) < RKS)
                    trf::st_async_val(trf::mapa(trf::smem_u32(&ctl->cpart[slot][crank][lane >> 2]), (unsigned)owner), pc[0],
                                      trf::mapa(trf::smem_u32(&ctl->cready[slot]), (unsigned)owner));
            }
            __syncwarp();
            if (lane == 0) trf::mbar_arrive(&ctl->pfree[i & 1]);        // the partial slot may be rewritten (sample i + 2)
            if (++s == NS) { s = 0; ph ^= 1u; }
            if (++owner == CL) { owner = 0; if (++slot == TRM_QO) slot = 0; }
        }
    } else if (warp == 4 + 2 * TRM_NWG + 2) {
        // ============== epilogue of the samples this CTA owns; v[n,:] -> every CTA ==============
        EpiMnArgs<T> ea;
        ea.partial = nullptr; ea.WT = 0; ea.RKs = RKS; ea.N = a.N; ea.R = R; ea.C = C;
        ea.FC = nullptr; ea.w = a.w; ea.y = a.y; ea.dP_in = nullptr; ea.class_w = a.class_w;
        ea.P = a.P; ea.pred = nullptr; ea.V = nullptr; ea.u_ws = a.u_ws; ea.dZ_ws = a.dZ_ws; ea.part = nullptr;
        double loss = 0.0;
        int slot = 0;
        unsigned phc = 0;
        int s = (int)crank % NS;
        const int sstep = CL % NS;
        // the label of the NEXT owned sample is fetched one iteration ahead: a global load issued when the partials
        // arrive would sit on the critical path (~2 000 cycles under a saturated HBM), class weights are in shared memory
        long long y_next = (int)crank < cnt ? __ldg(a.y + cid + (long long)crank * a.NC) : 0;
        for (int i = (int)crank; i < cnt; i += CL) {
            const long long n = (long long)cid + (long long)i * a.NC;
            const int yn = (int)y_next;
            if (i + CL < cnt) y_next = __ldg(a.y + n + (long long)CL * a.NC);
            const T omega = sCW[yn];
            if (lane == 0) trf::mbar_arrive_expect_tx(&ctl->cready[slot], (unsigned)(CL * RKS * sizeof(T)));
            if (lane == 0) TRM_STAMP(4, i);
            TRM_WAIT(&ctl->cready[slot], phc);
            __syncwarp();
            if (lane == 0) TRM_STAMP(5, i);
            // u[n,:] = sum of the CL CTA partials: lane c (and its twin c + 16) holds CTA c's partial, a fixed 4-step
            // butterfly sums them (same tree on every launch)
            T u[RKS], vv[RKS];
            {
                const int cc = lane & (TRM_MAX_CL - 1);
                trf::SVec<T, RKS>::ld(&ctl->cpart[slot][cc < CL ? cc : 0][0], u);
#pragma unroll
                for (int r = 0; r < RKS; ++r) u[r] = cc < CL ? u[r] : (T)0;
#pragma unroll
                for (int st = 3; st >= 0; --st) {
#pragma unroll
                    for (int r = 0; r < RKS; ++r) u[r] += __shfl_xor_sync(TR_FULL, u[r], 1 << st);
                }
            }
            if (lane == 0) TRM_STAMP(6, i);
            T eP = (T)0, eqs = (T)1, edZ = (T)0;
            if (TRM_DBG(4)) {
#pragma unroll
                for (int r = 0; r < RKS; ++r) vv[r] = u[r];
            } else if (C <= 16) {
                trm_epilogue<T, RKS, 4>(ea, n, u, lane, sFCw, yn, omega, vv, eP, eqs, edZ);
            } else {
                trm_epilogue<T, RKS, 5>(ea, n, u, lane, sFCw, yn, omega, vv, eP, eqs, edZ);
            }
            if (lane == 0) TRM_STAMP(7, i);
            if (lane < CL) {
#pragma unroll
                for (int r = 0; r < RKS; ++r)
                    trf::st_async_val(trf::mapa(trf::smem_u32(&ctl->vbuf[s][r]), (unsigned)lane), vv[r],
                                      trf::mapa(trf::smem_u32(&ctl->rready[s]), (unsigned)lane));
            }
            trm_epilogue_tail<T, RKS>(ea, n, u, lane, yn, omega, eP, eqs, edZ, loss);
            __syncwarp();
            if (++slot == TRM_QO) { slot = 0; phc ^= 1u; }
            s += sstep;
            if (s >= NS) s -= NS;
        }
        loss = warp_sum(loss);
        if (lane == 0) a.losspart[(size_t)cid * CL + crank] = loss;
    }
#ifdef TRM_TRACE
    else if (warp == 4 + 2 * TRM_NWG + 3) {
        // trace builds only: an observer on the otherwise idle warp 15 stamps the moment a sample's bytes have landed
        if (lane == 0) {
            int s = 0;
            unsigned ph = 0;
            for (int i = 0; i < cnt; ++i) {
                TRM_WAIT(&ctl->full[s], ph);
                TRM_STAMP(14, i);
                if (++s == NS) { s = 0; ph ^= 1u; }
            }
        }
    }
#endif
    // no CTA may exit while a peer can still write into its shared memory
    __syncwarp();
#if TRM_TMEM
    trf::tmem_fence_before();
#endif
    trf::cluster_arrive();
    trf::cluster_wait();
#if TRM_TMEM
    trf::tmem_fence_after();
    if (warp == TRM_NT / 32 - 1) trf::tmem_dealloc(ctl->tmem_base);     // every thread of the cluster is past its last TMEM access
#endif
#ifdef TRM_TRACE
    __syncthreads();
    if (a.trace && cid == 0 && crank == 0)
        for (int i = threadIdx.x; i < TRM_TRACE_N * TRM_TRACE_EV; i += TRM_NT) a.trace[i] = strace[i];
#endif
}

// F3 in the layout of the constant buffer: (I_k, RKS), channels r >= R zero
template <typename T>
__global__ void k_pack_f3(const T* __restrict__ F3, int IK, int R, int RKS, T* __restrict__ out) {
    for (int idx = threadIdx.x; idx < IK * RKS; idx += blockDim.x) {
        const int i3 = idx / RKS, r = idx % RKS;
        out[idx] = r < R ? F3[i3 * R + r] : (T)0;
    }
}

// dFt_k[i_k, r] = sum over (slot, CTA, warp) of the flushed A partials, in double; one block per (i_k, r)
template <typename T>
__global__ void __launch_bounds__(128) k_reduce_A(const T* __restrict__ Apart, int nparts, int IK, int RKS, int R,
                                                  double* __restrict__ out /* gradsum + foff[k-1] */) {
    __shared__ double sbuf[32];
    const int i3 = blockIdx.x / R, r = blockIdx.x % R;
    double s = 0.0;
    for (int p = threadIdx.x; p < nparts; p += blockDim.x) s += (double)Apart[(size_t)p * IK * RKS + i3 * RKS + r];
    s = block_sum(s, sbuf);
    if (threadIdx.x == 0) out[i3 * R + r] = s;
}
