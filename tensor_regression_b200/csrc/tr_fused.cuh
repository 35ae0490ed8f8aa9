// tr_fused.cuh — single-pass fused forward + gradient kernel for the standard model (SURVEY H8 ii).
//
// One thread-block CLUSTER holds one sample of X in its distributed shared memory: CTA c of a
// CL-CTA cluster owns the contiguous slice [c*Dc, (c+1)*Dc) of the feature axis, NS stages deep.
// X is streamed from HBM ONCE per fit iteration; the second pass (gradient) reads the sample from
// shared memory.  Everything is asynchronous and mbarrier-driven — no block-wide or cluster-wide
// barrier sits in the sample loop, so the warps of a CTA drift freely within the NS-stage window:
//
//   producer thread   TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) of the CTA's slice of
//                     sample j into stage j % NS as soon as all compute warps released it (empty[s]).
//   8 forward warps   phase A(i), as soon as sample i landed: dot the warp's 16-byte chunks (shared
//                     memory) with the register-resident CP coefficients B[i] -> warp partial of y_hat
//                     -> pA[i%NS][warp], arrive on redA.  They never wait for anything but data.
//   8 gradient warps  phase B(i), as soon as res_i is known: G += res * x from the still-resident
//                     stage (register accumulators), then release the stage to the producer.
//   reducer thread    sums the 16 warp partials, publishes the CTA partial to EVERY CTA of the cluster
//                     through DSMEM (st.shared::cluster + remote mbarrier arrive), waits until all CL
//                     partials of the sample arrived, forms y_hat_n and res_n in a fixed order
//                     (bit-identical in every CTA), hands res to the compute warps (rready).
//
// Outputs have the same layout as the two-pass path (Gpart slots, loss partials, y_hat), so the
// reduction / MTTKRP / finish kernels are shared.  Deterministic: no atomics, fixed summation orders.
#pragma once
#include "tr_kernels.cuh"

#define TR_FUSED_NWC 8                                // warps per compute role (forward / gradient)
#define TR_FUSED_NCT (TR_FUSED_NWC * 32)              // threads per compute role
#define TR_FUSED_NT (2 * TR_FUSED_NCT + 64)           // forward + gradient + producer warp + reducer warp
#define TR_FUSED_MAX_CL 16
#define TR_FUSED_MAX_NS 6

template <typename T>
struct FusedArgs {
    const T* X;
    const T* y;
    long long N;
    const T* FtT;
    const T* w;
    const T* theta;
    int bias_off;
    Geo geo;
    T* Gpart;            // (NC * nchunk, 1, Dpad)
    long long Dpad;
    T* yhat;             // may be null
    T* res;              // (N) residuals y_hat - y out (loss sums are formed from it by k_ressum)
    int CL;              // CTAs per cluster
    int NC;              // clusters in the grid
    int Dc;              // feature elements of the LARGEST CTA slice: ceil(D/VEC / CL) * VEC (slices are ragged when
                         // D/VEC is not a multiple of CL; the shared-memory stage stride is this size)
    int NS;              // shared-memory stages per CTA (>= 3)
    int nchunk;          // G is flushed to a fresh slot every spc samples (bounds fp32 sum length)
    long long spc;
    unsigned stage_bytes;  // Dc * sizeof(T), multiple of 16 (stage stride; a CTA loads its own slice's bytes)
    unsigned piece;        // bytes per cp.async.bulk instruction (multiple of 16)
    int pace;              // minimum cycles between two TMA issues of a CTA (anti-bunching), 0 = off
    long long* trace;      // debug (tools/fused_trace.cu): clock64 stamps of cluster 0 / rank 0, else null
};

namespace trf {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster; release at cluster scope
// orders the preceding st.shared::cluster of this thread before the arrival
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
#ifdef TRF_WAIT_NS
    // build option: suspend-time hint on every wait of the cluster kernels (see mbar_try_wait_hint below)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"((unsigned)TRF_WAIT_NS) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded waits: a pipeline bug must trap, never hang the GPU (try_wait suspends the thread in hardware)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    for (unsigned it = 0; !mbar_try_wait(bar, parity); ++it)
        if (it > (1u << 24)) __trap();
}
// the same wait with a suspend-time hint: the hardware parks the thread for up to `ns` nanoseconds per try instead of
// returning after its (short) default time limit, so a warp that waits for microseconds does not spend the issue slots
// of its sub-partition on the polling loop (k_spec_single: a third of all executed instructions were this loop)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, unsigned parity, unsigned ns) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, unsigned parity, unsigned ns = 20000u) {
    for (unsigned it = 0; !mbar_try_wait_hint(bar, parity, ns); ++it)
        if (it > (1u << 22)) __trap();
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, unsigned parity) {
    for (unsigned it = 0; !mbar_try_wait_cluster(bar, parity); ++it)
        if (it > (1u << 24)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// 8-byte store into a peer CTA's shared memory that signals complete_tx(8) on the peer's mbarrier:
// data and notification travel together, no fence needed on the sender
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(remote_addr), "l"(__double_as_longlong(v)), "r"(remote_bar) : "memory");
}
template <typename T, int VEC> struct SLoad;
template <> struct SLoad<float, 4> {
    static __device__ __forceinline__ void ld(const float* p, float (&x)[4]) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
};
template <> struct SLoad<double, 2> {
    static __device__ __forceinline__ void ld(const double* p, double (&x)[2]) {
        const double2 v = *reinterpret_cast<const double2*>(p);
        x[0] = v.x; x[1] = v.y;
    }
};
}  // namespace trf

// debug timeline: event e of sample i (cluster 0, CTA rank 0 only), TR_TRACE_N samples from TR_TRACE_I0
#ifndef TR_TRACE_I0
#define TR_TRACE_I0 64
#endif
#define TR_TRACE_N 24
#define TR_TRACE_EV 48
// stamps go to shared memory (a global store before an mbarrier.arrive would make the arrive's release
// wait for the store to be performed and distort the timeline) and are copied out at kernel end
#ifdef TR_FUSED_TRACE
#define TR_TRACE(ev, i)                                                                              \
    do {                                                                                             \
        if (a.trace && cid == 0 && crank == 0 && (i) >= TR_TRACE_I0 && (i) < TR_TRACE_I0 + TR_TRACE_N) \
            strace[((i) - TR_TRACE_I0) * TR_TRACE_EV + (ev)] = clock64();                            \
    } while (0)
#else
#define TR_TRACE(ev, i) do { } while (0)
#endif

// control block in the dynamic shared memory, after the NS stages
struct FusedCtl {
    uint64_t full[TR_FUSED_MAX_NS];                      // TMA landed (tx barrier, 1 arrival)
    uint64_t empty[TR_FUSED_MAX_NS];                     // stage released by the NWC compute warps
    uint64_t redA[TR_FUSED_MAX_NS];                      // NWC warp partials of a sample written
    uint64_t rready[TR_FUSED_MAX_NS];                    // res of a sample available
    uint64_t cready[2 * TR_FUSED_MAX_NS];                // CL cluster partials of a sample arrived
    double pA[TR_FUSED_MAX_NS][TR_FUSED_NWC];            // fp32 kernels store floats in the low halves
    double resv[TR_FUSED_MAX_NS];
    double cpart[2 * TR_FUSED_MAX_NS][TR_FUSED_MAX_CL];  // written remotely (DSMEM)
    int dims[TR_MAX_MODES];
    int foff[TR_MAX_MODES + 2];
};

template <typename T, int E>
__global__ void __launch_bounds__(TR_FUSED_NT, 1) k_fused_std(const FusedArgs<T> a) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int NCT = TR_FUSED_NCT;
    constexpr int NWC = TR_FUSED_NWC;
    extern __shared__ __align__(128) unsigned char tr_smem_fused[];
    const int lane = threadIdx.x & 31;
    // role -> warp id.  Warp w issues on sub-partition w % 4, and each sub-partition has its own
    // in-order shared-memory (MIO) instruction queue: forward warps sit on sub-partitions 0,1 and
    // gradient warps on 2,3 so that the forward warps' dependent shuffle chain never queues behind
    // the gradient warps' bursts of LDS.128 (measured: ~2400 cycles per sample, tools/fused_trace).
    const int hw_wid = threadIdx.x >> 5;
    const int wid = hw_wid >= 2 * TR_FUSED_NWC ? hw_wid                                       // producer 16, reducer 17
                    : ((hw_wid & 3) < 2 ? (hw_wid >> 2) * 2 + (hw_wid & 3)                    // forward: 0..7
                                        : TR_FUSED_NWC + (hw_wid >> 2) * 2 + (hw_wid & 3) - 2);  // gradient: 8..15
    const int tid = wid * 32 + lane;                                                          // role-relative thread id
    const unsigned crank = trf::cluster_ctarank();
    const int cid = blockIdx.x / a.CL;
    const int NS = a.NS, QC = 2 * a.NS;

    // layout: [control block | factor rows + rank weights | pad to 1024 | NS stages]
    const int k = a.geo.k, R = a.geo.R, pfeat = a.geo.pfeat;
    FusedCtl* ctl = reinterpret_cast<FusedCtl*>(tr_smem_fused);
    T* sF = reinterpret_cast<T*>(tr_smem_fused + ((sizeof(FusedCtl) + 15) / 16) * 16);
    const size_t head = ((((sizeof(FusedCtl) + 15) / 16) * 16 + (size_t)(pfeat + R) * sizeof(T)) + 1023) / 1024 * 1024;
    T* stage0 = reinterpret_cast<T*>(tr_smem_fused + head);
#ifdef TR_FUSED_TRACE
    long long* strace = reinterpret_cast<long long*>(tr_smem_fused + head + (size_t)NS * a.stage_bytes);
    for (int i = threadIdx.x; i < TR_TRACE_N * TR_TRACE_EV; i += TR_FUSED_NT) strace[i] = 0;
#endif

    for (int i = threadIdx.x; i < pfeat + R; i += TR_FUSED_NT) sF[i] = i < pfeat ? a.FtT[i] : a.w[i - pfeat];
    if (threadIdx.x < TR_MAX_MODES) ctl->dims[threadIdx.x] = a.geo.dims[threadIdx.x];
    if (threadIdx.x < TR_MAX_MODES + 2) ctl->foff[threadIdx.x] = a.geo.foff[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            trf::mbar_init(&ctl->full[s], 1);
            trf::mbar_init(&ctl->empty[s], NWC);
            trf::mbar_init(&ctl->redA[s], NWC);
            trf::mbar_init(&ctl->rready[s], 1);
        }
        for (int q = 0; q < QC; ++q) trf::mbar_init(&ctl->cready[q], 1);     // 1 local arrive + CL*8 bytes of tx
        trf::fence_mbar_init();
    }
    __syncthreads();
    // every CTA of the cluster must have initialised its barriers before any remote arrive / store
    trf::cluster_arrive();
    trf::cluster_wait();

    // samples of this cluster: n = cid + j*NC, j = 0..cnt-1.  All ring positions and mbarrier phase
    // bits are carried incrementally (no 64-bit division in the sample loop).
    const int cnt = cid < a.N ? (int)((a.N - cid + a.NC - 1) / a.NC) : 0;
    // this CTA's slice of the feature axis, in 16-byte chunks: the first (D/VEC mod CL) ranks hold one more
    const int chunks_total = (int)(a.geo.D / VEC);
    const int cq = chunks_total / a.CL, crem = chunks_total % a.CL;
    const int my_chunks = cq + ((int)crank < crem ? 1 : 0);
    const long long my_off = ((long long)crank * cq + ((int)crank < crem ? (int)crank : crem)) * VEC;   // elements
    const int my_elems = my_chunks * VEC;
    const unsigned my_bytes = (unsigned)my_elems * (unsigned)sizeof(T);
    const size_t sample_stride = (size_t)a.NC * (size_t)a.geo.D;      // elements between this cluster's samples

    if (wid == 2 * NWC) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            const T* src = a.X + (size_t)cid * (size_t)a.geo.D + (size_t)my_off;
            int s = 0;
            unsigned ph = 0;                       // parity of the empty-phase to wait for (from the 2nd lap on)
            long long t_next = clock64();
            for (int j = 0; j < cnt; ++j) {
                if (j >= NS) trf::mbar_wait(&ctl->empty[s], ph);
                // pacing: three stages that free up together would otherwise load together and then
                // compute together ("bunching"), leaving HBM idle half of the time
                if (a.pace > 0) {
                    while (clock64() < t_next) __nanosleep(64);
                    t_next = clock64() + a.pace;
                }
                TR_TRACE(0, j);
                trf::mbar_arrive_expect_tx(&ctl->full[s], my_bytes);
                const unsigned char* sp = reinterpret_cast<const unsigned char*>(src);
                unsigned char* dst = reinterpret_cast<unsigned char*>(stage0) + (size_t)s * a.stage_bytes;
                for (unsigned off = 0; off < my_bytes; off += a.piece) {
                    const unsigned len = my_bytes - off < a.piece ? my_bytes - off : a.piece;
                    trf::bulk_g2s(dst + off, sp + off, len, &ctl->full[s]);
                }
                src += sample_stride;
                if (++s == NS) { s = 0; if (j >= NS) ph ^= 1u; }
            }
        }
    } else if (wid == 2 * NWC + 1) {
        // ========================= reducer / cluster exchange =========================
        // all 32 lanes take part: lane c sends this CTA's partial to peer c (st.async), lane 0 owns
        // the waits and the scalar math
        const T bias = a.theta[a.bias_off];
        int q = 0, qc = 0;
        unsigned phq = 0, phc = 0;
        const T* yp = a.y + cid;
        T* yhp = a.yhat ? a.yhat + cid : nullptr;
        T* rp = a.res + cid;                                            // residuals out: loss sums are formed later, in fp64
        for (int i = 0; i < cnt; ++i) {
            T yn = (T)0;
            if (lane == 0) {
                yn = __ldg(yp);                                         // in flight while we wait below
                trf::mbar_arrive_expect_tx(&ctl->cready[qc], (unsigned)(a.CL * sizeof(double)));
            }
            trf::mbar_wait(&ctl->redA[q], phq);                         // all 32 lanes wait on the barrier itself
            __syncwarp();
            if (lane == 0) TR_TRACE(3, i);
            T pc = (T)0;
#pragma unroll
            for (int w8 = 0; w8 < NWC; ++w8) pc += reinterpret_cast<const T*>(&ctl->pA[q][w8])[0];   // same order in every lane
            if (lane < a.CL) {
                double slot = 0.0;
                reinterpret_cast<T*>(&slot)[0] = pc;
                trf::st_async_f64(trf::mapa(trf::smem_u32(&ctl->cpart[qc][crank]), (unsigned)lane), slot,
                                  trf::mapa(trf::smem_u32(&ctl->cready[qc]), (unsigned)lane));
            }
            if (lane == 0) {
                trf::mbar_wait(&ctl->cready[qc], phc);
                TR_TRACE(4, i);
                T yh = bias;
                for (int c = 0; c < a.CL; ++c) yh += reinterpret_cast<const T*>(&ctl->cpart[qc][c])[0];
                const T res = yh - yn;
                reinterpret_cast<T*>(&ctl->resv[q])[0] = res;
                trf::mbar_arrive(&ctl->rready[q]);
                if (crank == 0) {
                    if (yhp) *yhp = yh;
                    *rp = res;
                }
            }
            __syncwarp();
            yp += a.NC;
            rp += a.NC;
            if (yhp) yhp += a.NC;
            if (++q == NS) { q = 0; phq ^= 1u; }
            if (++qc == QC) { qc = 0; phc ^= 1u; }
        }
    } else if (wid < NWC) {
        // ============================ forward warps: phase A ============================
        T coef[E][VEC];
        unsigned cmask = 0;
        const int chunks = my_chunks;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int ch = j * NCT + tid;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                T tmp[1] = {(T)0};
                if (ch < chunks) {
                    cmask |= 1u << j;
                    const unsigned i = (unsigned)(my_off + (long long)ch * VEC + v);
                    tr_coef_at<T, 1>(sF, sF + pfeat, ctl->dims, ctl->foff, k, R, 0, i, tmp);
                }
                coef[j][v] = tmp[0];
            }
        }
        const unsigned stage_elems = a.stage_bytes / (unsigned)sizeof(T);
        const T* xthread = stage0 + (size_t)tid * VEC;                  // this thread's first chunk of stage 0
        // one pass of phase A over the stage at xs: <x, coef> over this thread's chunks, warp-reduced in T
        auto phaseA = [&](const T* xs, unsigned m) -> T {
            T p[4] = {(T)0, (T)0, (T)0, (T)0};
#pragma unroll
            for (int j = 0; j < E; ++j) {
                if ((m >> j) & 1u) {
                    T x[VEC];
                    trf::SLoad<T, VEC>::ld(xs + (size_t)j * NCT * VEC, x);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) p[j & 3] = tr_fma<T>(x[v], coef[j][v], p[j & 3]);
                }
            }
            T pw = (p[0] + p[1]) + (p[2] + p[3]);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) pw += __shfl_xor_sync(TR_FULL, pw, off);
            return pw;
        };
        int sA = 0;
        unsigned phA = 0;
        for (int i = 0; i < cnt; ++i) {
            // EVERY lane waits on the mbarrier itself (hardware-suspended try_wait).  Do not let one lane poll
            // while the other 31 park in __syncwarp(): after a wait of a few thousand cycles the parked lanes
            // take ~2 300 cycles to resume, which made the first forward / gradient phase after every idle
            // spell 3x slower and cost 30 % of the kernel (measured with tools/fused_trace.cu).
            trf::mbar_wait(&ctl->full[sA], phA);
            __syncwarp();
            if (tid == 0) TR_TRACE(1, i);
            if (lane == 0) TR_TRACE(8 + wid * 2, i);
            // warp reduction in T: for fp32 the whole per-sample chain stays off the fp64 pipe
            const T pw = phaseA(xthread + (size_t)sA * stage_elems, cmask);
            if (lane == 0) {
                reinterpret_cast<T*>(&ctl->pA[sA][wid])[0] = pw;
                trf::mbar_arrive(&ctl->redA[sA]);
            }
            if (tid == 0) TR_TRACE(2, i);
            if (lane == 0) TR_TRACE(9 + wid * 2, i);
            if (++sA == NS) { sA = 0; phA ^= 1u; }
        }
    } else {
        // =========================== gradient warps: phase B ===========================
        const int tb = tid - NCT;                                       // 0 .. NCT-1
        T acc[E][VEC];
        unsigned cmask = 0;
        const int chunks = my_chunks;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            if (j * NCT + tb < chunks) cmask |= 1u << j;
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[j][v] = (T)0;
        }
        const unsigned stage_elems = a.stage_bytes / (unsigned)sizeof(T);
        const T* xthread = stage0 + (size_t)tb * VEC;
        auto phaseB = [&](const T* xs, unsigned m, T r) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                if ((m >> j) & 1u) {
                    T x[VEC];
                    trf::SLoad<T, VEC>::ld(xs + (size_t)j * NCT * VEC, x);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[j][v] = tr_fma<T>(r, x[v], acc[j][v]);
                }
            }
        };
        int sB = 0;
        unsigned phB = 0;
        int left = (int)(a.spc < cnt ? a.spc : cnt);                    // samples until the next flush of G
        T* gp = a.Gpart + (size_t)cid * a.nchunk * (size_t)a.Dpad + (size_t)my_off + (size_t)tb * VEC;
        for (int i = 0; i < cnt; ++i) {
            trf::mbar_wait(&ctl->rready[sB], phB);                     // all lanes wait (see phase A)
            __syncwarp();
            if (tb == 0) TR_TRACE(5, i);
            if (lane == 0) TR_TRACE(24 + (wid - NWC) * 2, i);
            const T r = reinterpret_cast<const T*>(&ctl->resv[sB])[0];
            phaseB(xthread + (size_t)sB * stage_elems, cmask, r);
            __syncwarp();
            if (lane == 0) trf::mbar_arrive(&ctl->empty[sB]);          // all lanes' reads of the stage are done
            if (tb == 0) TR_TRACE(6, i);
            if (lane == 0) TR_TRACE(25 + (wid - NWC) * 2, i);
            if (++sB == NS) { sB = 0; phB ^= 1u; }
            // chunk boundary: flush G to its slot (bounds the length of every fp32 running sum)
            if (--left == 0) {
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    if ((cmask >> j) & 1u) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            gp[(size_t)j * NCT * VEC + v] = acc[j][v];
                            acc[j][v] = (T)0;
                        }
                    }
                }
                gp += a.Dpad;
                const int rem = cnt - (i + 1);
                left = (int)(a.spc < rem ? a.spc : rem);
            }
        }
        // slots of chunks this cluster never reached must still be defined for the reduction
        const int used = cnt > 0 ? (int)((cnt - 1) / a.spc) + 1 : 0;
        for (int c2 = used; c2 < a.nchunk; ++c2) {
            T* gz = a.Gpart + ((size_t)cid * a.nchunk + c2) * (size_t)a.Dpad + (size_t)my_off;
            for (int e = tb; e < my_elems; e += NCT) gz[e] = (T)0;
        }
    }
    // no CTA may exit while a peer can still write into its shared memory
    __syncwarp();
    trf::cluster_arrive();
    trf::cluster_wait();
#ifdef TR_FUSED_TRACE
    __syncthreads();
    if (a.trace && cid == 0 && crank == 0)
        for (int i = threadIdx.x; i < TR_TRACE_N * TR_TRACE_EV; i += TR_FUSED_NT) a.trace[i] = strace[i];
#endif
}
