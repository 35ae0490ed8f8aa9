// tr_fused.cuh — single-pass fused forward + gradient kernel for the standard model (SURVEY H8 ii).
//
// One thread-block CLUSTER holds one sample of X in its distributed shared memory: CTA c of a
// CL-CTA cluster owns the contiguous slice [c*Dc, (c+1)*Dc) of the feature axis.  Per sample:
//   1. the slice arrives in shared memory by TMA bulk copies (cp.async.bulk + mbarrier
//      complete_tx), NS stages deep, issued one sample ahead of the compute;
//   2. phase A: every thread dots its 16-byte chunks (read from shared memory) with its register-
//      resident slice of the CP coefficient B[i]; block reduction -> the CTA's partial of y_hat;
//   3. the CL partials are exchanged through DSMEM (st.shared::cluster to every peer) and a
//      cluster barrier; every CTA forms y_hat_n, res_n in the same fixed order;
//   4. phase B: G[i] += res_n * X[n,i] from the SAME shared-memory stage — the second pass over
//      X never touches HBM (that is the whole point: X is streamed from HBM once per iteration).
// The cluster barrier of sample i is overlapped with phase B of sample i-1 and phase A of i+1.
// Outputs have the same layout as the two-pass path (Gpart slots, loss partials, y_hat), so the
// reduction / MTTKRP / finish kernels are shared.
#pragma once
#include "tr_kernels.cuh"

#define TR_FUSED_NT 512
#define TR_FUSED_MAX_CL 16

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
    double* part;        // (NC, 2): sum res, sum res^2
    int CL;              // CTAs per cluster
    int NC;              // clusters in the grid
    int Dc;              // feature elements per CTA slice (D / CL)
    int NS;              // shared-memory stages per CTA
    int nchunk;          // G is flushed to a fresh slot every spc samples (bounds fp32 sum length)
    long long spc;
    unsigned stage_bytes;  // Dc * sizeof(T), multiple of 16
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
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a pipeline bug must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    for (unsigned it = 0; !mbar_try_wait(bar, parity); ++it)
        if (it > (1u << 26)) __trap();
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

// shared memory carve-up (bytes):
//   [0, NS*stage_bytes)                 X stages (128-byte aligned)
//   then  full[NS] mbarriers (8 B each, padded to 128)
//   then  sred[2][16] doubles, part[2][TR_FUSED_MAX_CL] doubles
//   then  factor rows + rank weights (T), dims/offset ints
template <typename T, int E>
__global__ void __launch_bounds__(TR_FUSED_NT, 1) k_fused_std(const FusedArgs<T> a) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int NT = TR_FUSED_NT;
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char tr_smem_fused[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned crank = trf::cluster_ctarank();
    const int cid = blockIdx.x / a.CL;

    unsigned char* sp = tr_smem_fused;
    T* stage0 = reinterpret_cast<T*>(sp);
    sp += (size_t)a.NS * a.stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sp);
    sp += 128;
    double* sred = reinterpret_cast<double*>(sp);            // [2][NW]
    sp += 2 * NW * sizeof(double);
    double* part = reinterpret_cast<double*>(sp);            // [2][TR_FUSED_MAX_CL]
    sp += 2 * TR_FUSED_MAX_CL * sizeof(double);
    T* sF = reinterpret_cast<T*>(sp);
    const int k = a.geo.k, R = a.geo.R, pfeat = a.geo.pfeat;
    sp += (size_t)((pfeat + R) * sizeof(T) + 15) / 16 * 16;
    int* sDims = reinterpret_cast<int*>(sp);
    int* sOff = sDims + TR_MAX_MODES;

    for (int i = tid; i < pfeat + R; i += NT) sF[i] = i < pfeat ? a.FtT[i] : a.w[i - pfeat];
    if (tid < TR_MAX_MODES) sDims[tid] = a.geo.dims[tid];
    if (tid < TR_MAX_MODES + 2) sOff[tid] = a.geo.foff[tid];
    if (tid == 0) {
        for (int s = 0; s < a.NS; ++s) trf::mbar_init(&full[s], 1);
        trf::fence_mbar_init();
    }
    __syncthreads();

    // samples of this cluster: n = cid + j*NC, j = 0..cnt-1
    const long long cnt = cid < a.N ? (a.N - cid + a.NC - 1) / a.NC : 0;
    const T* xslice = a.X + (long long)crank * a.Dc;          // + n*D per sample

    auto issue = [&](long long j) {                            // called by tid 0 only
        const int s = (int)(j % a.NS);
        const long long n = cid + j * a.NC;
        trf::mbar_arrive_expect_tx(&full[s], a.stage_bytes);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(xslice + n * a.geo.D);
        unsigned char* dst = reinterpret_cast<unsigned char*>(stage0) + (size_t)s * a.stage_bytes;
        for (unsigned off = 0; off < a.stage_bytes; off += 16384u) {
            const unsigned len = a.stage_bytes - off < 16384u ? a.stage_bytes - off : 16384u;
            trf::bulk_g2s(dst + off, src + off, len, &full[s]);
        }
    };
    if (tid == 0)
        for (long long j = 0; j < a.NS && j < cnt; ++j) issue(j);
    // every CTA of the cluster must be running before its shared memory is written remotely
    trf::cluster_arrive();
    trf::cluster_wait();

    // register-resident coefficient slice and gradient accumulators
    T coef[E][VEC], acc[E][VEC];
    unsigned cmask = 0;
    const int chunks = a.Dc / VEC;
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int ch = j * NT + tid;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            T tmp[1] = {(T)0};
            if (ch < chunks) {
                cmask |= 1u << j;
                const unsigned i = (unsigned)((long long)crank * a.Dc + (long long)ch * VEC + v);
                tr_coef_at<T, 1>(sF, sF + pfeat, sDims, sOff, k, R, 0, i, tmp);
            }
            coef[j][v] = tmp[0];
            acc[j][v] = (T)0;
        }
    }
    const double bias = (double)a.theta[a.bias_off];
    const uint32_t part_u32 = trf::smem_u32(part);

    double l1 = 0.0, l2 = 0.0;         // sum res, sum res^2 (rank 0, thread 0)
    double res_prev = 0.0;
    for (long long i = 0; i <= cnt; ++i) {
        double pc = 0.0;
        if (i < cnt) {
            // ---- phase A (compute part): partial of y_hat over this CTA's slice ----
            const int s = (int)(i % a.NS);
            trf::mbar_wait(&full[s], (unsigned)((i / a.NS) & 1));
            const T* xs = stage0 + (size_t)s * (a.stage_bytes / sizeof(T));
            T p0 = (T)0, p1 = (T)0;
#pragma unroll
            for (int j = 0; j < E; ++j) {
                if ((cmask >> j) & 1u) {
                    T x[VEC];
                    trf::SLoad<T, VEC>::ld(xs + (size_t)(j * NT + tid) * VEC, x);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if (j & 1) p1 = tr_fma<T>(x[v], coef[j][v], p1);
                        else p0 = tr_fma<T>(x[v], coef[j][v], p0);
                    }
                }
            }
            double pw = warp_sum((double)p0 + (double)p1);
            if (lane == 0) sred[(i & 1) * NW + wid] = pw;
        }
        __syncthreads();                                        // (S1) sred visible; everyone is past B(i-2)
        if (i < cnt) {
#pragma unroll
            for (int w8 = 0; w8 < NW; ++w8) pc += sred[(i & 1) * NW + w8];
        }
        // ---- barrier of sample i-1 completes: its partials from all CTAs are in part[(i-1)&1] ----
        if (i > 0) {
            trf::cluster_wait();
            double yh = bias;
            for (int c = 0; c < a.CL; ++c) yh += part[((i - 1) & 1) * TR_FUSED_MAX_CL + c];
            const long long n = cid + (i - 1) * a.NC;
            const T yhT = (T)yh;
            const double res = (double)yhT - (double)__ldg(a.y + n);
            res_prev = res;
            if (crank == 0 && tid == 0) {
                if (a.yhat) a.yhat[n] = yhT;
                l1 += res;
                l2 += res * res;
            }
        }
        // ---- publish this CTA's partial of sample i to every CTA of the cluster ----
        if (i < cnt) {
            if (tid < a.CL)
                trf::st_cluster_f64(trf::mapa(part_u32 + (uint32_t)(((i & 1) * TR_FUSED_MAX_CL + crank) * sizeof(double)),
                                              (unsigned)tid), pc);
            trf::cluster_arrive();
        }
        // ---- phase B of sample i-1: G += res * x, from the stage that is still resident ----
        if (i > 0) {
            const int s = (int)((i - 1) % a.NS);
            const T* xs = stage0 + (size_t)s * (a.stage_bytes / sizeof(T));
            const T r = (T)res_prev;
#pragma unroll
            for (int j = 0; j < E; ++j) {
                if ((cmask >> j) & 1u) {
                    T x[VEC];
                    trf::SLoad<T, VEC>::ld(xs + (size_t)(j * NT + tid) * VEC, x);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[j][v] = tr_fma<T>(r, x[v], acc[j][v]);
                }
            }
            // chunk boundary: flush G to its slot (bounds the length of every fp32 running sum)
            const long long jj = i - 1;
            if ((jj + 1) % a.spc == 0 || jj == cnt - 1) {
                const long long slot = (long long)cid * a.nchunk + jj / a.spc;
                T* gp = a.Gpart + slot * a.Dpad + (long long)crank * a.Dc;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    if ((cmask >> j) & 1u) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            gp[(size_t)(j * NT + tid) * VEC + v] = acc[j][v];
                            acc[j][v] = (T)0;
                        }
                    }
                }
            }
        }
        __syncthreads();                                        // (S2) stage (i-1)%NS is free
        if (tid == 0 && i > 0 && i - 1 + a.NS < cnt) issue(i - 1 + a.NS);
    }
    if (crank == 0 && tid == 0) { a.part[cid * 2 + 0] = l1; a.part[cid * 2 + 1] = l2; }
    // slots of chunks this cluster never reached must still be defined for the reduction
    {
        const long long used = cnt > 0 ? (cnt - 1) / a.spc + 1 : 0;
        for (long long c2 = used; c2 < a.nchunk; ++c2) {
            T* gp = a.Gpart + ((long long)cid * a.nchunk + c2) * a.Dpad + (long long)crank * a.Dc;
            for (int e = tid; e < a.Dc; e += NT) gp[e] = (T)0;
        }
    }
}
