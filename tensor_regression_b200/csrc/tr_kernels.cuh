// tr_kernels.cuh — sm_100a kernels of the CP tensor-regression fit iteration.
//
// Data layout in HBM:
//   X        (N, I_1..I_k) row-major, sample stride D = prod I_m; streamed, never copied
//   theta    flat parameter vector [F_0 | .. | F_{k-1} | F_C | bias], (I_m, R) row-major blocks
//   partial  (N, WT, RK)   per-sample, per-warp-tile, per-channel partial inner products (channels of a
//                          tile contiguous: one warp store of pass 1 touches U lines, not U*RK)
//   V        (N, RK)       per-sample weights of the gradient pass (residual / v[n,r])
//   Gpart    (slots, RK, Dpad) split-N partial sums of  sum_n V[n,c] * X[n,:]
//   Gred     (RK, D) double   = sum over slots
//
// Both streaming kernels are warp-autonomous: one warp owns a TILE = 32*E*VEC element slice of
// the feature axis (E 16-byte chunks per lane, interleaved across the 32 lanes so every warp
// load instruction touches 512 contiguous bytes) and walks a strided subset of the samples.
// The per-element CP coefficients of the slice (std: B[i] = sum_r w_r prod_m Ft_m[i_m,r];
// mn: K[i,r] = prod_m Ft_m[i_m,r]) are built once per warp from the factor rows staged in
// shared memory and then live in registers — no dense coefficient tensor exists in memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tr_b200.h"

#define TR_TPB 256
#ifndef TR_MINB
#define TR_MINB 2      // resident blocks per SM the streaming kernels are compiled for (<= 128 registers)
#endif
#define TR_WPB (TR_TPB / 32)
#define TR_FULL 0xffffffffu

struct Geo {
    int k, R, C;
    int dims[TR_MAX_MODES];
    int foff[TR_MAX_MODES + 2];   // offset of factor m in theta (foff[k] = class factor / end of features)
    long long D;
    int pfeat;                    // R * sum I_m
    int pf;                       // pfeat + C*R
};

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int tr_next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }
__host__ __device__ constexpr int tr_log2(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

template <typename T, int VEC> struct XLoad;
template <> struct XLoad<float, 4> {
    static __device__ __forceinline__ void ld(const float* p, float (&x)[4]) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]) : "l"(p));
    }
};
template <> struct XLoad<float, 1> {
    static __device__ __forceinline__ void ld(const float* p, float (&x)[1]) {
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(x[0]) : "l"(p));
    }
};
template <> struct XLoad<double, 2> {
    static __device__ __forceinline__ void ld(const double* p, double (&x)[2]) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(x[0]), "=d"(x[1]) : "l"(p));
    }
};
template <> struct XLoad<double, 1> {
    static __device__ __forceinline__ void ld(const double* p, double (&x)[1]) {
        asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(x[0]) : "l"(p));
    }
};

template <typename T> __device__ __forceinline__ T tr_fma(T a, T b, T c);
template <> __device__ __forceinline__ float tr_fma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double tr_fma<double>(double a, double b, double c) { return fma(a, b, c); }

// Two independent fp32 FMAs in one instruction (sm_100 FFMA2, PTX fma.rn.f32x2): d.{x,y} = a.{x,y} * b.{x,y} +
// d.{x,y}, each lane rounded like fmaf — same bits as two scalar FMAs, half the issue slots.  Used in the
// forward contraction with several channels (issue-slot / latency limited); in the gradient pass it
// measured slower (k_grad<6>: 4.71 ms vs 3.97 ms on cfg 3) and is not used there.
__device__ __forceinline__ void tr_ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long d, a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

// vals[c] += x[v] * coef[v][c] for one 16-byte chunk: FFMA2 over the channel pairs (c, c+1) for float
// (x[v] duplicated), scalar otherwise.  Every vals[c] sees the same sequence of FMAs either way.
template <typename T, int VEC, int RK>
struct ChunkDot {
    static __device__ __forceinline__ void run(T* vals, const T (&x)[VEC], const T (&coef)[VEC][RK]) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int c = 0; c < RK; ++c) vals[c] = tr_fma<T>(x[v], coef[v][c], vals[c]);
    }
};
template <int RK>
struct ChunkDot<float, 4, RK> {
    static __device__ __forceinline__ void run(float* vals, const float (&x)[4], const float (&coef)[4][RK]) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
#pragma unroll
            for (int c = 0; c + 1 < RK; c += 2) tr_ffma2(vals[c], vals[c + 1], x[v], x[v], coef[v][c], coef[v][c + 1]);
            if (RK & 1) vals[RK - 1] = fmaf(x[v], coef[v][RK - 1], vals[RK - 1]);
        }
    }
};

// Sum M values (M a power of two <= 32) across the 32 lanes with M-1+(5-log2 M) shuffles instead
// of 5*M: each halving step trades half of the values with the xor-partner.  On return v[0] in
// lane l is the total of value index (l >> (5 - log2 M)).  Deterministic.
//
// NOSEL: the first NOSEL halving steps need no per-lane selects because the caller stored its values
// PERMUTED: a lane whose bit (16 >> s) is set, s < NOSEL, keeps logical value i at physical index
// i ^ (M >> (s+1)) (all such bits xor-ed together, see warp_perm_mask).  Then every lane sends its
// physical upper half and keeps its lower half; the pairing of lanes, and therefore the summation
// order and the bits of the result, are the same as with NOSEL = 0.
template <typename T, int M, int NOSEL = 0>
__device__ __forceinline__ void warp_reduce_transpose(T (&v)[M], int lane) {
    constexpr int LG = tr_log2(M);
#pragma unroll
    for (int s = 0; s < LG; ++s) {
        const int half = (M >> s) >> 1;
        const int off = 16 >> s;
        if (s < NOSEL) {
#pragma unroll
            for (int i = 0; i < half; ++i) v[i] = v[i] + __shfl_xor_sync(TR_FULL, v[i + half], off);
        } else {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const T send = up ? v[i] : v[i + half];
                const T keep = up ? v[i + half] : v[i];
                v[i] = keep + __shfl_xor_sync(TR_FULL, send, off);
            }
        }
    }
#pragma unroll
    for (int off = (16 >> LG); off >= 1; off >>= 1) v[0] += __shfl_xor_sync(TR_FULL, v[0], off);
}

// index permutation of a lane for the NOSEL leading steps of warp_reduce_transpose<T, M, NOSEL>
template <int M, int NOSEL>
__device__ __forceinline__ int warp_perm_mask(int lane) {
    int m = 0;
#pragma unroll
    for (int s = 0; s < NOSEL; ++s)
        if (lane & (16 >> s)) m |= (M >> s) >> 1;
    return m;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(TR_FULL, v, off);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmax(v, __shfl_xor_sync(TR_FULL, v, off));
    return v;
}

// block-wide sum of one double per thread; result valid in thread 0 (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ double block_sum(double v, double* sbuf /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sbuf[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < nw ? sbuf[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

template <typename T> __device__ __forceinline__ T tr_softplus(T x, T beta, T thr);
template <> __device__ __forceinline__ float tr_softplus<float>(float x, float beta, float thr) {
    const float bx = x * beta;
    return bx > thr ? x : log1pf(expf(bx)) / beta;
}
template <> __device__ __forceinline__ double tr_softplus<double>(double x, double beta, double thr) {
    const double bx = x * beta;
    return bx > thr ? x : log1p(exp(bx)) / beta;
}

// ---------------------------------------------------------------------------------------------
// Pass 1: forward contraction.  partial[n, t, c] = sum_{i in tile t} X[n,i] * coef[i,c]
// (lin_model std:123-130 / model mn:181-186 without ever forming the dense B).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct FwdArgs {
    const T* X;
    long long N;
    const T* FtT;      // softplus-ed factors (feature modes first)
    const T* w;        // rank weights (R)
    Geo geo;
    T* partial;        // (N, WT, RK)
    int WT;            // warp tiles per sample
    int Gn;            // sample groups: group g owns samples g, g+Gn, g+2Gn, ...
    int mode;          // 0: one channel, coefficient B[i] (standard); 1: R channels K[i,r] (multinomial)
};

// CP coefficient(s) of feature element i, from the factor rows in shared memory.  Kept out of
// line (and un-unrolled) on purpose: it runs once per warp tile, the streaming loop is what must
// stay tight.  mode 0: out[0] = sum_r w_r prod_m Ft_m[i_m,r];  mode 1: out[c] = prod_m Ft_m[i_m,c].
template <typename T, int RK>
__device__ __noinline__ void tr_coef_at(const T* sF, const T* sW, const int* sDims, const int* sOff, int k, int R,
                                        int mode, unsigned i, T* out) {
    int idx[TR_MAX_MODES];
#pragma unroll 1
    for (int m = k - 1; m >= 0; --m) {
        const unsigned d = (unsigned)sDims[m];
        const unsigned q = i / d;
        idx[m] = (int)(i - q * d);
        i = q;
    }
    if (mode == 0) {
        T s = (T)0;
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            T p = sW[r];
#pragma unroll 1
            for (int m = 0; m < k; ++m) p *= sF[sOff[m] + idx[m] * R + r];
            s += p;
        }
        out[0] = s;
    } else {
#pragma unroll 1
        for (int c = 0; c < RK; ++c) {
            T p = (T)0;
            if (c < R) {
                p = (T)1;
#pragma unroll 1
                for (int m = 0; m < k; ++m) p *= sF[sOff[m] + idx[m] * R + c];
            }
            out[c] = p;
        }
    }
}

template <typename T, int RK, int E, int U, int VEC, int MINB = TR_MINB>
__global__ void __launch_bounds__(TR_TPB, MINB) k_fwd(const FwdArgs<T> a) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    __shared__ int sDims[TR_MAX_MODES], sOff[TR_MAX_MODES + 2];
    T* sF = reinterpret_cast<T*>(tr_smem);
    const int k = a.geo.k, R = a.geo.R, pfeat = a.geo.pfeat;
    for (int i = threadIdx.x; i < pfeat + R; i += TR_TPB) sF[i] = i < pfeat ? a.FtT[i] : a.w[i - pfeat];
    if (threadIdx.x < TR_MAX_MODES) sDims[threadIdx.x] = a.geo.dims[threadIdx.x];
    if (threadIdx.x < TR_MAX_MODES + 2) sOff[threadIdx.x] = a.geo.foff[threadIdx.x];
    __syncthreads();
    const T* sW = sF + pfeat;

    constexpr int TILE = 32 * E * VEC;
    constexpr int RKR = tr_next_pow2(RK);
    constexpr int M = tr_next_pow2(U * RKR);          // padded with zeros when U is not a power of two
    constexpr int LGM = tr_log2(M);
    static_assert(M <= 32, "next_pow2(U * next_pow2(RK)) must be <= 32");
    // With several channels the shuffle reduction is a large share of the instructions.  Its steps that
    // halve the SAMPLE index run without selects when each lane loads sample (u ^ mu) into slot u
    // (U a power of two): same lanes paired in the same order, same bits.
    constexpr int NOSEL = (RK > 1 && (U & (U - 1)) == 0) ? tr_log2(U) : 0;

    const int lane = threadIdx.x & 31;
    const int mu = warp_perm_mask<M, NOSEL>(lane) / RKR;     // sample-slot permutation of this lane
    const long long warp_global = (long long)blockIdx.x * TR_WPB + (threadIdx.x >> 5);
    const long long wtot = (long long)gridDim.x * TR_WPB;
    const long long D = a.geo.D;
    const long long items = (long long)a.WT * a.Gn;

    for (long long item = warp_global; item < items; item += wtot) {
        const int t = (int)(item % a.WT);
        const int g = (int)(item / a.WT);
        const long long tile_base = (long long)t * TILE;

        // ---- per-warp coefficient slice, built from the factor rows in shared memory ----
        T coef[E][VEC][RK];
        unsigned cmask = 0;
#pragma unroll
        for (int j = 0; j < E; ++j) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const long long i = tile_base + (long long)(j * 32 + lane) * VEC + v;
                T tmp[RK];
#pragma unroll
                for (int c = 0; c < RK; ++c) tmp[c] = (T)0;
                if (i < D) {
                    if (v == 0) cmask |= 1u << j;
                    tr_coef_at<T, RK>(sF, sW, sDims, sOff, k, R, a.mode, (unsigned)i, tmp);
                }
#pragma unroll
                for (int c = 0; c < RK; ++c) coef[j][v][c] = tmp[c];
            }
        }

        // ---- stream the group's samples ----
        const T* xbase = a.X + tile_base + (long long)lane * VEC;
        T x[U][E][VEC];
        // loads of the batch of U samples that starts at sample n0 (zeros past the end / past D)
        auto issue = [&](long long n0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long n = n0 + (long long)(u ^ mu) * a.Gn;
                const T* xp = xbase + n * D;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    if (n < a.N && ((cmask >> j) & 1u)) {
                        XLoad<T, VEC>::ld(xp + j * 32 * VEC, x[u][j]);
                    } else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) x[u][j][v] = (T)0;
                    }
                }
            }
        };
        const long long nstep = (long long)U * a.Gn;
        if (g < a.N) issue(g);
        for (long long n0 = g; n0 < a.N; n0 += nstep) {
            T vals[M];
#pragma unroll
            for (int q = 0; q < M; ++q) vals[q] = (T)0;
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < E; ++j) ChunkDot<T, VEC, RK>::run(&vals[u * RKR], x[u][j], coef[j]);
            // the next batch's loads go out as soon as the FMAs have consumed x: they are in flight during
            // the shuffle reduction and the partial stores below
            if (n0 + nstep < a.N) issue(n0 + nstep);
            warp_reduce_transpose<T, M, NOSEL>(vals, lane);
            if ((lane & ((1 << (5 - LGM)) - 1)) == 0) {
                const int q = lane >> (5 - LGM);
                const int u = q / RKR, c = q % RKR;
                const long long n = n0 + (long long)u * a.Gn;
                if (u < U && c < RK && n < a.N) a.partial[(n * a.WT + t) * RK + c] = vals[0];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Pass 2: residual-weighted accumulation  Gpart[slot, c, i] = sum_{n in slot} V[n,c] * X[n,i]
// (what MmBackward0 computes, std:372,462 / mn:361,457) — registers hold the warp's slice of G.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct GradArgs {
    const T* X;
    long long N;
    long long D;
    const T* V;        // (N, RK)
    T* Gpart;          // (nchunk*Gn, RK, Dpad)
    long long Dpad;
    int WT, Gn;
    int nchunk;        // each group's sample list is cut into nchunk slots (bounds fp32 summation length)
    long long spc;     // samples per chunk
};

template <typename T, int RK, int E, int U, int VEC, int MINB = TR_MINB>
__global__ void __launch_bounds__(TR_TPB, MINB) k_grad(const GradArgs<T> a) {
    constexpr int TILE = 32 * E * VEC;
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * TR_WPB + (threadIdx.x >> 5);
    const long long wtot = (long long)gridDim.x * TR_WPB;
    const long long items = (long long)a.WT * a.Gn;

    for (long long item = warp_global; item < items; item += wtot) {
        const int t = (int)(item % a.WT);
        const int g = (int)(item / a.WT);
        const long long tile_base = (long long)t * TILE;
        unsigned cmask = 0;
#pragma unroll
        for (int j = 0; j < E; ++j)
            if (tile_base + (long long)(j * 32 + lane) * VEC < a.D) cmask |= 1u << j;
        const T* xbase = a.X + tile_base + (long long)lane * VEC;
        const long long Sg = g < a.N ? (a.N - g + a.Gn - 1) / a.Gn : 0;   // samples in this group

        for (int ch = 0; ch < a.nchunk; ++ch) {
            T acc[E][VEC][RK];
#pragma unroll
            for (int j = 0; j < E; ++j)
#pragma unroll
                for (int v = 0; v < VEC; ++v)
#pragma unroll
                    for (int c = 0; c < RK; ++c) acc[j][v][c] = (T)0;
            const long long s0 = (long long)ch * a.spc;
            const long long s1 = (s0 + a.spc < Sg) ? s0 + a.spc : Sg;
            for (long long s = s0; s < s1; s += U) {
                T x[U][E][VEC];
                T vv[U][RK];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool ok = s + u < s1;
                    const long long n = g + (s + u) * a.Gn;
                    const T* xp = xbase + n * a.D;
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        if (ok && ((cmask >> j) & 1u)) {
                            XLoad<T, VEC>::ld(xp + j * 32 * VEC, x[u][j]);
                        } else {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) x[u][j][v] = (T)0;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < RK; ++c) vv[u][c] = ok ? __ldg(a.V + n * RK + c) : (T)0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < E; ++j)
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
#pragma unroll
                            for (int c = 0; c < RK; ++c)
                                acc[j][v][c] = tr_fma<T>(vv[u][c], x[u][j][v], acc[j][v][c]);
            }
            // every slot is written (zeros when empty) so the reduction needs no masks
            T* gp = a.Gpart + ((long long)(ch * a.Gn + g) * RK) * a.Dpad + tile_base + (long long)lane * VEC;
#pragma unroll
            for (int c = 0; c < RK; ++c)
#pragma unroll
                for (int j = 0; j < E; ++j)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) gp[c * a.Dpad + j * 32 * VEC + v] = acc[j][v][c];
        }
    }
}
