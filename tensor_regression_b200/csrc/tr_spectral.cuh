// tr_spectral.cuh — the fit iteration of spectral_tensor_regression.py (second half): the standard CP
// model on X (T, W, D) with outputs y (T, NO), plus "spectral" rank components whose first-mode factor has a complex
// axis: the contraction over W is followed by a norm over that axis before the second contraction
// (stepwise_spectral_model, spectral:339-390).  Included by tr_api.cu only.
//
// Both parts contract the window axis W first, so ONE pass over X serves them:
//
//   a[t,q,d]  = sum_w X[t,w,d] G[w,q]       q < Rn: normal component r = q,        G[w,q] = Fn0[w,r]
//                                           q >= Rn: spectral (r,c), q = Rn + r CC + c, G[w,q] = Fc0[w,r,c]
//   normal    s_n[t,r] = sum_d a[t,r,d] Fn1[d,r]                          (lin_model, spectral:118-165)
//   spectral  m[t,r,d] = sqrt(sum_c a[t,(r,c),d]^2)   s_s[t,r] = sum_d m[t,r,d] Fc1[d,r]       (spectral:385-387)
//   output    yhat[t,n] = sum_r w_r s_n[t,r] Fn2[n,r] + sum_r s_s[t,r] Fc2[n,r] + nb bias[n]   (spectral:577-578;
//             nb = number of non-empty parts: both lin_model and stepwise_spectral_model add the bias)
//   loss      MSE over (t, n)                                             (spectral:581-586)
//
// and the backward pass is the same chain reversed (what autograd does for the reference): res = yhat - y,
// ds_n = w (res Fn2), ds_s = res Fc2, da[t,q,d] = ds_n[r] Fn1[d,r]  or  ds_s[r] Fc1[d,r] a[t,q,d] / m[t,r,d], and the
// one large sum  dG[w,q] = sum_t sum_d X[t,w,d] da[t,q,d]  is the second pass over X.  Per-element arithmetic: Q FMA
// in each pass, HBM-bound like the two-pass kernels of the other models (algorithmic bytes 2 W D sizeof(T) per
// sample; the a / da arrays add 2 Q / W of that).
#pragma once
#include "tr_small.cuh"

#define TRS_MAXQ 8        // channels per launch of the streaming kernels (more: several passes over X)
#define TRS_MAXR 16       // rank_normal + rank_spectral
#define TRS_WT 8          // window rows per warp in the gradient pass

struct SpecGeo {
    int W, D, NO, Rn, Rs, CC, Q, RT;      // Q = Rn + Rs*CC channels, RT = Rn + Rs
    int off[7];                            // theta offsets: Fn0, Fn1, Fn2, Fc0, Fc1, Fc2, bias
};

// ---------------------------------------------------------------------------------------------
// pass 1: a[t,q,d] = sum_w X[t,w,d] G[w,q] for the channels [q0, q0 + QT) of this launch.
// A warp owns an item = (sample, tile of 32*VEC features); lanes run along d (coalesced 16-byte loads), the window
// axis is the loop: QT*VEC register accumulators, UW rows of X in flight per lane, G rows broadcast from shared memory.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SpecFwdArgs {
    const T* X;
    long long N;
    const T* FtT;          // softplus-ed parameters (T), theta layout
    T* A;                  // (N, Q, D)
    SpecGeo g;
    int q0;                // first channel of this launch
};

// VEC consecutive, VEC*sizeof(T)-aligned elements of an ordinary (cached) array
template <typename T, int VEC> struct SpecVec;
template <> struct SpecVec<float, 4> {
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
};
template <> struct SpecVec<double, 2> {
    static __device__ __forceinline__ void st(double* p, const double (&v)[2]) { *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); }
    static __device__ __forceinline__ void ld(const double* p, double (&v)[2]) { const double2 t = __ldg(reinterpret_cast<const double2*>(p)); v[0] = t.x; v[1] = t.y; }
};
template <typename T> struct SpecVec<T, 1> {
    static __device__ __forceinline__ void st(T* p, const T (&v)[1]) { p[0] = v[0]; }
    static __device__ __forceinline__ void ld(const T* p, T (&v)[1]) { v[0] = __ldg(p); }
};

// a row of QP values (QP a multiple of the 16-byte chunk) from shared memory with 16-byte loads (warp-uniform address: broadcast)
template <typename T> struct VECG;
template <> struct VECG<float> {
    static constexpr int v = 4;
    template <int QP> static __device__ __forceinline__ void ld(const float* p, float (&o)[QP]) {
#pragma unroll
        for (int i = 0; i < QP; i += 4) {
            const float4 t = *reinterpret_cast<const float4*>(p + i);
            o[i] = t.x; o[i + 1] = t.y; o[i + 2] = t.z; o[i + 3] = t.w;
        }
    }
};
template <> struct VECG<double> {
    static constexpr int v = 2;
    template <int QP> static __device__ __forceinline__ void ld(const double* p, double (&o)[QP]) {
#pragma unroll
        for (int i = 0; i < QP; i += 2) {
            const double2 t = *reinterpret_cast<const double2*>(p + i);
            o[i] = t.x; o[i + 1] = t.y;
        }
    }
};

// VEC consecutive, aligned elements of shared memory
template <typename T, int VEC> struct SpecSm;
template <> struct SpecSm<float, 4> {
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
};
template <> struct SpecSm<double, 2> {
    static __device__ __forceinline__ void st(double* p, const double (&v)[2]) { *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); }
    static __device__ __forceinline__ void ld(const double* p, double (&v)[2]) { const double2 t = *reinterpret_cast<const double2*>(p); v[0] = t.x; v[1] = t.y; }
};
template <typename T> struct SpecSm<T, 1> {
    static __device__ __forceinline__ void st(T* p, const T (&v)[1]) { p[0] = v[0]; }
    static __device__ __forceinline__ void ld(const T* p, T (&v)[1]) { v[0] = p[0]; }
};

template <typename T>
__device__ __forceinline__ T spec_G(const T* FtT, const SpecGeo& g, int w, int q) {
    return q < g.Rn ? FtT[g.off[0] + w * g.Rn + q] : FtT[g.off[3] + w * (g.Rs * g.CC) + (q - g.Rn)];
}

template <typename T, int QT, int VEC, int UW>
__global__ void __launch_bounds__(TR_TPB) k_spec_fwd(const SpecFwdArgs<T> a) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    T* sG = reinterpret_cast<T*>(tr_smem);                                   // (W, QT)
    const SpecGeo& g = a.g;
    for (int i = threadIdx.x; i < g.W * QT; i += TR_TPB) {
        const int w = i / QT, q = a.q0 + i % QT;
        sG[i] = q < g.Q ? spec_G(a.FtT, g, w, q) : (T)0;
    }
    __syncthreads();
    constexpr int TILE = 32 * VEC;
    const int lane = threadIdx.x & 31;
    const long long DT = (g.D + TILE - 1) / TILE;
    const long long items = a.N * DT;
    const long long wtot = (long long)gridDim.x * TR_WPB;
    const size_t WD = (size_t)g.W * g.D;
    for (long long item = (long long)blockIdx.x * TR_WPB + (threadIdx.x >> 5); item < items; item += wtot) {
        const long long t = item / DT;
        const int d0 = (int)(item % DT) * TILE + lane * VEC;
        const bool act = d0 < g.D;
        T acc[VEC][QT];
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int q = 0; q < QT; ++q) acc[v][q] = (T)0;
        const T* xp = a.X + (size_t)t * WD + (act ? d0 : 0);
        for (int w = 0; w < g.W; w += UW) {
            T x[UW][VEC];
#pragma unroll
            for (int u = 0; u < UW; ++u) {
                if (act && w + u < g.W) XLoad<T, VEC>::ld(xp + (size_t)(w + u) * g.D, x[u]);
                else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[u][v] = (T)0;
                }
            }
#pragma unroll
            for (int u = 0; u < UW; ++u) {
                const T* gr = sG + (size_t)(w + u < g.W ? w + u : 0) * QT;
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    const T gq = gr[q];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[v][q] = tr_fma<T>(x[u][v], gq, acc[v][q]);
                }
            }
        }
        if (act) {
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                if (a.q0 + q < g.Q) {
                    T out[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) out[v] = acc[v][q];
                    SpecVec<T, VEC>::st(a.A + ((size_t)t * g.Q + a.q0 + q) * g.D + d0, out);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// per-sample epilogue: one warp per sample, lanes along d, sums in double
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SpecEpiArgs {
    const T* A;            // (N, Q, D)
    const double* Ft64;    // softplus-ed parameters (double), theta layout
    const T* theta;        // raw parameters (bias)
    const T* w;            // rank weights (RT); the spectral part does not use them (spectral:339-390)
    const T* y;            // (N, NO) or null (forward only)
    long long N;
    SpecGeo g;
    double nb;             // bias multiplicity (1 or 2)
    // outputs (any may be null)
    T* yhat;               // (N, NO) model of the fit (lin_model + stepwise_spectral_model)
    T* res;                // (N, NO) yhat - y
    T* U;                  // (N, RT + 1): [ s_n | s_s | 1 ]
    T* dS;                 // (N, RT): ds_n | ds_s
    T* Mc;                 // (N, Rs, D): m[t,r,d]
    T* DA;                 // (N, Q, D): da
    double* part;          // (blocks) sum of res^2 per block
};

template <typename T>
__global__ void __launch_bounds__(TR_TPB) k_spec_epi(const SpecEpiArgs<T> a) {
    __shared__ double sloss[TR_WPB];
    const SpecGeo& g = a.g;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double* Fn1 = a.Ft64 + g.off[1];
    const double* Fn2 = a.Ft64 + g.off[2];
    const double* Fc1 = a.Ft64 + g.off[4];
    const double* Fc2 = a.Ft64 + g.off[5];
    double loss = 0.0;
    const long long wtot = (long long)gridDim.x * TR_WPB;
    for (long long t = (long long)blockIdx.x * TR_WPB + wid; t < a.N; t += wtot) {
        const T* At = a.A + (size_t)t * g.Q * g.D;
        double s[TRS_MAXR];                            // s_n | s_s
#pragma unroll
        for (int r = 0; r < TRS_MAXR; ++r) s[r] = 0.0;
        for (int d = lane; d < g.D; d += 32) {
#pragma unroll
            for (int r = 0; r < TRS_MAXR; ++r) {
                if (r < g.Rn) {
                    s[r] += (double)At[(size_t)r * g.D + d] * Fn1[d * g.Rn + r];
                } else if (r < g.RT) {
                    const int rs = r - g.Rn;
                    double ss = 0.0;
                    for (int c = 0; c < g.CC; ++c) {
                        const double v = (double)At[(size_t)(g.Rn + rs * g.CC + c) * g.D + d];
                        ss += v * v;
                    }
                    const double m = sqrt(ss);
                    if (a.Mc) a.Mc[((size_t)t * g.Rs + rs) * g.D + d] = (T)m;
                    s[r] += m * Fc1[d * g.Rs + rs];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < TRS_MAXR; ++r) {
            if (r < g.RT) s[r] = warp_sum(s[r]);
        }
        if (a.U && lane <= g.RT) {
            T uv = (T)1;
#pragma unroll
            for (int r = 0; r < TRS_MAXR; ++r) if (r == lane && r < g.RT) uv = (T)s[r];
            a.U[t * (g.RT + 1) + lane] = uv;
        }
        // outputs: lanes along n
        double ds[TRS_MAXR];
#pragma unroll
        for (int r = 0; r < TRS_MAXR; ++r) ds[r] = 0.0;
        for (int n = lane; n < g.NO; n += 32) {
            double yl = 0.0, ysp = 0.0;
#pragma unroll
            for (int r = 0; r < TRS_MAXR; ++r) {
                if (r < g.Rn) yl += (double)a.w[r] * s[r] * Fn2[n * g.Rn + r];
                else if (r < g.RT) ysp += s[r] * Fc2[n * g.Rs + (r - g.Rn)];
            }
            const double b = (double)a.theta[g.off[6] + n];
            const double yh = yl + ysp + a.nb * b;
            if (a.yhat) a.yhat[t * g.NO + n] = (T)yh;
            if (a.y) {
                const double rr = yh - (double)a.y[t * g.NO + n];
                if (a.res) a.res[t * g.NO + n] = (T)rr;
                loss += rr * rr;
#pragma unroll
                for (int r = 0; r < TRS_MAXR; ++r) {
                    if (r < g.Rn) ds[r] += rr * Fn2[n * g.Rn + r];
                    else if (r < g.RT) ds[r] += rr * Fc2[n * g.Rs + (r - g.Rn)];
                }
            }
        }
        if (a.y && a.DA) {
#pragma unroll
            for (int r = 0; r < TRS_MAXR; ++r) {
                if (r < g.RT) { ds[r] = warp_sum(ds[r]); if (r < g.Rn) ds[r] *= (double)a.w[r]; }
            }
            if (a.dS && lane < g.RT) {
#pragma unroll
                for (int r = 0; r < TRS_MAXR; ++r) if (r == lane) a.dS[t * g.RT + r] = (T)ds[r];
            }
            T* Dt = a.DA + (size_t)t * g.Q * g.D;
            for (int d = lane; d < g.D; d += 32) {
#pragma unroll
                for (int r = 0; r < TRS_MAXR; ++r) {
                    if (r < g.Rn) {
                        Dt[(size_t)r * g.D + d] = (T)(ds[r] * Fn1[d * g.Rn + r]);
                    } else if (r < g.RT) {
                        const int rs = r - g.Rn;
                        // d||a|| / da = a / ||a||, 0 at the origin (torch.norm's subgradient)
                        double ss = 0.0;
                        for (int c = 0; c < g.CC; ++c) {
                            const double v = (double)At[(size_t)(g.Rn + rs * g.CC + c) * g.D + d];
                            ss += v * v;
                        }
                        const double m = sqrt(ss);
                        const double k = m > 0.0 ? ds[r] * Fc1[d * g.Rs + rs] / m : 0.0;
                        for (int c = 0; c < g.CC; ++c) {
                            const size_t qi = (size_t)(g.Rn + rs * g.CC + c) * g.D + d;
                            Dt[qi] = (T)(k * (double)At[qi]);
                        }
                    }
                }
            }
        }
    }
    if (a.part) {
        loss = warp_sum(loss);
        if (lane == 0) sloss[wid] = loss;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int i = 0; i < TR_WPB; ++i) tot += sloss[i];
            a.part[blockIdx.x] = tot;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// second-mode factor gradients: dF1[d, c] = sum_t dS[t,c] * M[t,c,d],  M = a (normal channels) or m (spectral).
// Block (x, slab) owns a slab of samples, threads run along d; double accumulators; part (slabs, RT, D).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SpecDf1Args {
    const T* A; const T* Mc; const T* dS;
    long long N; SpecGeo g; int slabs; double* part;
};

template <typename T>
__global__ void __launch_bounds__(TR_TPB) k_spec_df1(const SpecDf1Args<T> a) {
    const SpecGeo& g = a.g;
    const int slab = blockIdx.y;
    const long long per = (a.N + a.slabs - 1) / a.slabs;
    const long long t0 = (long long)slab * per, t1 = t0 + per < a.N ? t0 + per : a.N;
    const int d = blockIdx.x * TR_TPB + threadIdx.x;
    if (d >= g.D) return;
    double acc[TRS_MAXR];
#pragma unroll
    for (int r = 0; r < TRS_MAXR; ++r) acc[r] = 0.0;
    for (long long t = t0; t < t1; ++t) {
#pragma unroll
        for (int r = 0; r < TRS_MAXR; ++r) {
            if (r < g.Rn) acc[r] += (double)__ldg(a.dS + t * g.RT + r) * (double)a.A[((size_t)t * g.Q + r) * g.D + d];
            else if (r < g.RT) acc[r] += (double)__ldg(a.dS + t * g.RT + r) * (double)a.Mc[((size_t)t * g.Rs + (r - g.Rn)) * g.D + d];
        }
    }
#pragma unroll
    for (int r = 0; r < TRS_MAXR; ++r)
        if (r < g.RT) a.part[((size_t)slab * g.RT + r) * g.D + d] = acc[r];
}

// ---------------------------------------------------------------------------------------------
// pass 2: dG[w,q] = sum_t sum_d X[t,w,d] da[t,q,d] for the channels [q0, q0 + QT).
// Warp (wt, grp) owns the TRS_WT window rows [wt*8, wt*8+8) and the samples grp, grp + G, ...: per (sample, d-tile) a
// lane loads its 16-byte chunk of the 8 rows of X and of the QT rows of da (the same da chunk is read by the warps of
// the other window tiles at about the same time: L2), 8*QT*VEC FMAs into 8*QT register sums.  The sums are folded
// across lanes and added in double to the warp's own slot every `spc` samples (bounds every fp32 running sum; one
// owner per slot, no atomics, fixed order).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SpecGradArgs {
    const T* X; const T* DA; long long N; SpecGeo g; int q0;
    int WTN, Gn; long long spc;
    double* part;          // (WTN * Gn, TRS_WT, QT)
};

template <typename T, int QT, int VEC>
__global__ void __launch_bounds__(TR_TPB) k_spec_grad(const SpecGradArgs<T> a) {
    const SpecGeo& g = a.g;
    constexpr int TILE = 32 * VEC;
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * TR_WPB + (threadIdx.x >> 5);
    const long long wtot = (long long)gridDim.x * TR_WPB;
    const long long items = (long long)a.WTN * a.Gn;
    const int DT = (g.D + TILE - 1) / TILE;
    const size_t WD = (size_t)g.W * g.D;
    for (long long item = warp_global; item < items; item += wtot) {
        const int wt = (int)(item % a.WTN);
        const int grp = (int)(item / a.WTN);
        const int w0 = wt * TRS_WT;
        T acc[TRS_WT][QT];
#pragma unroll
        for (int i = 0; i < TRS_WT; ++i)
#pragma unroll
            for (int q = 0; q < QT; ++q) acc[i][q] = (T)0;
        double* slot = a.part + (size_t)item * TRS_WT * QT;
        if (lane < TRS_WT * QT) slot[lane] = 0.0;
        if (lane + 32 < TRS_WT * QT) slot[lane + 32] = 0.0;
        long long left = a.spc;
        for (long long t = grp; t < a.N; t += a.Gn) {
            const T* xt = a.X + (size_t)t * WD;
            const T* dt = a.DA + (size_t)t * g.Q * g.D;
            for (int tile = 0; tile < DT; ++tile) {
                const int d0 = tile * TILE + lane * VEC;
                if (d0 < g.D) {
                    T x[TRS_WT][VEC], da[QT][VEC];
#pragma unroll
                    for (int i = 0; i < TRS_WT; ++i) {
                        if (w0 + i < g.W) XLoad<T, VEC>::ld(xt + (size_t)(w0 + i) * g.D + d0, x[i]);
                        else {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) x[i][v] = (T)0;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < QT; ++q) {
                        if (a.q0 + q < g.Q) {
                            SpecVec<T, VEC>::ld(dt + (size_t)(a.q0 + q) * g.D + d0, da[q]);
                        } else {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) da[q][v] = (T)0;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < TRS_WT; ++i)
#pragma unroll
                        for (int q = 0; q < QT; ++q)
#pragma unroll
                            for (int v = 0; v < VEC; ++v) acc[i][q] = tr_fma<T>(x[i][v], da[q][v], acc[i][q]);
                }
            }
            if (--left == 0 || t + a.Gn >= a.N) {
                // fold across lanes and add to the warp's slot (double)
                __syncwarp();
#pragma unroll
                for (int i = 0; i < TRS_WT; ++i)
#pragma unroll
                    for (int q = 0; q < QT; ++q) {
                        double sv = (double)acc[i][q];
                        sv = warp_sum(sv);
                        if (lane == 0) slot[i * QT + q] += sv;
                        acc[i][q] = (T)0;
                    }
                left = a.spc;
            }
        }
    }
}

// third-mode factors and bias from the (NO, RT + 1) matrix  M[n, c] = wcat_c sum_t res[t,n] U[t,c]  (k_dfc + k_colsum):
// c < Rn -> dFn2[n,c], c < RT -> dFc2[n, c - Rn], c = RT -> nb * sum_t res[t,n] (bias), and the loss sum
#ifndef TR_TEMPLATES_ONLY
__global__ void k_spec_scatter(const double* __restrict__ M, const double* __restrict__ losspart, int nloss, SpecGeo g,
                               double nb, double* __restrict__ gradsum) {
    const int total = g.NO * (g.RT + 1);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int n = e / (g.RT + 1), c = e % (g.RT + 1);
        const double v = M[e];
        if (c < g.Rn) gradsum[g.off[2] + n * g.Rn + c] = v;
        else if (c < g.RT) gradsum[g.off[5] + n * g.Rs + (c - g.Rn)] = v;
        else gradsum[g.off[6] + n] = nb * v;
    }
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nloss; ++i) s += losspart[i];
        gradsum[g.off[6] + g.NO] = s;
    }
}
#endif

// wcat = [ rank weights of the normal components | 1 for the spectral components and the bias column ]
template <typename T>
__global__ void k_spec_wcat(const T* __restrict__ w, int Rn, int n, T* __restrict__ wcat) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) wcat[i] = i < Rn ? w[i] : (T)1;
}

// ---------------------------------------------------------------------------------------------
// post-hoc outputs of the estimator (predict / predict_latents, spectral:895-1034), one warp per sample:
//   sq[q]        = sum_d a[t,q,d] F1[d, r(q)]                              (second contraction of every channel)
//   latents[t,r] = sq[r], r < Rn                                           (stepwise_latents_model, spectral:284-337)
//   yhat_lin     = sum_r w_r sq[r] Fn2[n,r] + bias[n]                      (lin_model, spectral:118-165)
//   spec_pred    = sqrt(sum_c (sum_r w_{Rn+r} sq[(r,c)] Fc2[n,r])^2) + bias[n]   (spectral_model, spectral:168-221: the
//                  norm over the complex axis is taken of the complete CP contraction, unlike the model of the fit)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SpecPredArgs {
    const T* A; const double* Ft64; const T* theta; const T* w; long long N; SpecGeo g;
    T* yhat_lin; T* spec_pred; T* latents;
};

template <typename T>
__global__ void __launch_bounds__(TR_TPB) k_spec_pred(const SpecPredArgs<T> a) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    const SpecGeo& g = a.g;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* sq = reinterpret_cast<double*>(tr_smem) + (size_t)wid * g.Q;
    const long long wtot = (long long)gridDim.x * TR_WPB;
    for (long long t = (long long)blockIdx.x * TR_WPB + wid; t < a.N; t += wtot) {
        const T* At = a.A + (size_t)t * g.Q * g.D;
        __syncwarp();
        for (int q = 0; q < g.Q; ++q) {
            const bool nrm = q < g.Rn;
            const int r = nrm ? q : (q - g.Rn) / g.CC;
            const double* F1 = nrm ? a.Ft64 + g.off[1] + r : a.Ft64 + g.off[4] + r;
            const int ld = nrm ? g.Rn : g.Rs;
            double v = 0.0;
            for (int d = lane; d < g.D; d += 32) v += (double)At[(size_t)q * g.D + d] * F1[(size_t)d * ld];
            v = warp_sum(v);
            if (lane == 0) sq[q] = v;
        }
        __syncwarp();
        if (a.latents && lane < g.Rn) a.latents[t * g.Rn + lane] = (T)sq[lane];
        for (int n = lane; n < g.NO; n += 32) {
            const double b = (double)a.theta[g.off[6] + n];
            if (a.yhat_lin) {
                double yl = 0.0;
                for (int r = 0; r < g.Rn; ++r) yl += (double)a.w[r] * sq[r] * a.Ft64[g.off[2] + n * g.Rn + r];
                a.yhat_lin[t * g.NO + n] = (T)(yl + b);
            }
            if (a.spec_pred) {
                double acc = 0.0;
                for (int c = 0; c < g.CC; ++c) {
                    double z = 0.0;
                    for (int r = 0; r < g.Rs; ++r)
                        z += (double)a.w[g.Rn + r] * sq[g.Rn + r * g.CC + c] * a.Ft64[g.off[5] + n * g.Rs + r];
                    acc += z * z;
                }
                a.spec_pred[t * g.NO + n] = (T)(sqrt(acc) + b);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 1 with the per-sample epilogue fused in, for samples whose D features fit ONE warp tile (D <= 32 * VEC) and
// Q <= TRS_MAXQ channels: the warp that contracted the window axis holds a[t,:,d] of its sample in registers, so the
// second contraction, the outputs, the residual, ds and da = d loss / d a follow in the same warp — a, m and the
// second-mode gradient never go through memory.  Per sample the kernel reads X[t] (and y[t]) and writes da[t] (Q / W
// of the sample), res[t], [s_n | s_s | 1]; the second-mode gradient sum_t ds[t,r] m[t,r,d] accumulates in registers
// (the lane owns the same d for every sample it sees) and is folded into the warp's double slot every `spc` samples.
// The channel sums go through a per-warp shared-memory scratch once, so the channels of a component are addressed at run
// time (q = Rn + r CC + c) while every register array is indexed statically by the component r.
// ---------------------------------------------------------------------------------------------
// sqrt(ss) and 1 / sqrt(ss) (0 at ss = 0) with one reciprocal square root: for float the hardware rsqrt (2 ulp) plus one
// Newton step on the product — well inside the 1e-5 tolerance and an order of magnitude fewer instructions than
// sqrtf + an IEEE division; exact operations for double
__device__ __forceinline__ void spec_norm(float ss, float& nrm, float& ri) {
    float r = ss > 0.0f ? rsqrtf(ss) : 0.0f;
    r = r * fmaf(-0.5f * ss * r, r, 1.5f);                 // one Newton step: relative error ~1e-7 -> ~1e-14 (then fp32 rounding)
    ri = r;
    nrm = ss * r;
}
__device__ __forceinline__ void spec_norm(double ss, double& nrm, double& ri) {
    nrm = sqrt(ss);
    ri = nrm > 0.0 ? 1.0 / nrm : 0.0;
}

template <typename T>
struct SpecFusedArgs {
    const T* X; const T* y; long long N;
    const T* FtT; const T* theta; const T* w;
    SpecGeo g; double nb;
    T* DA; T* res; T* U; T* yhat;
    double* df1part;       // (warps, QT, 32 * VEC) doubles: slot of warp (blockIdx * TR_WPB + wid)
    double* losspart;      // (blocks)
    long long spc;
};

template <typename T, int QT, int VEC, int UW>
__global__ void __launch_bounds__(TR_TPB, 2) k_spec_fused(const SpecFusedArgs<T> a) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    __shared__ double sloss[TR_WPB];
    const SpecGeo& g = a.g;
    constexpr int QP = (QT + VECG<T>::v - 1) / VECG<T>::v * VECG<T>::v;      // row stride of sG: whole 16-byte chunks
    T* sG = reinterpret_cast<T*>(tr_smem);                                   // (W, QP)
    T* sF2 = sG + (size_t)g.W * QP;                                          // (NO, QT): w_r Fn2[n,r] | Fc2[n,r]
    T* sB = sF2 + (size_t)g.NO * QT;                                         // (NO): nb * bias
    // per-warp scratch (QT, 32 * VEC): the window sums a[q][d] of the warp's current sample, so that the channels of a
    // component can be addressed at run time (q = Rn + r * CC + c) while every register array stays statically indexed
    T* sA = reinterpret_cast<T*>(tr_smem + ((((size_t)g.W * QP + (size_t)g.NO * QT + g.NO) * sizeof(T) + 15) / 16) * 16)
            + (size_t)(threadIdx.x >> 5) * QT * 32 * VEC + (threadIdx.x & 31) * VEC;
    for (int i = threadIdx.x; i < g.W * QP; i += TR_TPB) {
        const int w = i / QP, q = i % QP;
        sG[i] = q < g.Q ? spec_G(a.FtT, g, w, q) : (T)0;
    }
    for (int i = threadIdx.x; i < g.NO * QT; i += TR_TPB) {
        const int n = i / QT, r = i % QT;
        T v = (T)0;
        if (r < g.Rn) v = a.w[r] * a.FtT[g.off[2] + n * g.Rn + r];
        else if (r < g.RT) v = a.FtT[g.off[5] + n * g.Rs + (r - g.Rn)];
        sF2[i] = v;
    }
    for (int i = threadIdx.x; i < g.NO; i += TR_TPB) sB[i] = (T)(a.nb * (double)a.theta[g.off[6] + i]);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int d0 = lane * VEC;
    const bool act = d0 < g.D;
    T f1[VEC][QT];                                                           // second-mode factor of component r at the lane's d
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int r = 0; r < QT; ++r) {
            T val = (T)0;
            if (act && r < g.Rn) val = a.FtT[g.off[1] + (d0 + v) * g.Rn + r];
            else if (act && r < g.RT) val = a.FtT[g.off[4] + (d0 + v) * g.Rs + (r - g.Rn)];
            f1[v][r] = val;
        }
    T accF[VEC][QT];                                                         // sum_t ds[t,r] m[t,r,d] since the last fold
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int r = 0; r < QT; ++r) accF[v][r] = (T)0;
    const long long wslot = (long long)blockIdx.x * TR_WPB + wid;
    double* slot = a.df1part + (size_t)wslot * QT * 32 * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int r = 0; r < QT; ++r) slot[(size_t)r * 32 * VEC + d0 + v] = 0.0;
    double loss = 0.0;
    long long left = a.spc;
    const long long wtot = (long long)gridDim.x * TR_WPB;
    const size_t WD = (size_t)g.W * g.D;
    for (long long t = wslot; t < a.N; t += wtot) {
        // targets of this sample: one per lane and round, fetched before the window loop
        T yv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) yv[k] = (lane + 32 * k < g.NO) ? __ldg(a.y + t * g.NO + lane + 32 * k) : (T)0;
        T acc[VEC][QT];
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int q = 0; q < QT; ++q) acc[v][q] = (T)0;
        const T* xp = a.X + (size_t)t * WD + (act ? d0 : 0);
        for (int w = 0; w < g.W; w += UW) {
            T x[UW][VEC];
#pragma unroll
            for (int u = 0; u < UW; ++u) {
                if (act && w + u < g.W) XLoad<T, VEC>::ld(xp + (size_t)(w + u) * g.D, x[u]);
                else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[u][v] = (T)0;
                }
            }
#pragma unroll
            for (int u = 0; u < UW; ++u) {
                T gq[QP];
                VECG<T>::template ld<QP>(sG + (size_t)(w + u < g.W ? w + u : 0) * QP, gq);
#pragma unroll
                for (int q = 0; q < QT; ++q) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[v][q] = tr_fma<T>(x[u][v], gq[q], acc[v][q]);
                }
            }
        }
        // window sums -> the warp's scratch (one 16-byte store per channel; read back only by the lane that wrote them)
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            T out[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) out[v] = acc[v][q];
            SpecSm<T, VEC>::st(sA + (size_t)q * 32 * VEC, out);
        }
        __syncwarp();
        // m[v][r]: a itself (normal component) or the norm over the component's complex channels; rinv = 1 / m (0 at 0)
        T m[VEC][QT], rinv[VEC][QT];
#pragma unroll
        for (int r = 0; r < QT; ++r) {
            if (r < g.Rn) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { m[v][r] = acc[v][r]; rinv[v][r] = (T)1; }
            } else if (r < g.RT) {
                T ss[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) ss[v] = (T)0;
                const T* ap = sA + (size_t)(g.Rn + (r - g.Rn) * g.CC) * 32 * VEC;
                for (int c = 0; c < g.CC; ++c) {
                    T av[VEC];
                    SpecSm<T, VEC>::ld(ap + (size_t)c * 32 * VEC, av);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) ss[v] = tr_fma<T>(av[v], av[v], ss[v]);
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) spec_norm(ss[v], m[v][r], rinv[v][r]);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { m[v][r] = (T)0; rinv[v][r] = (T)0; }
            }
        }
        // second contraction: s[r] = sum_d m[d,r] F1[d,r]  (lane partial, then an all-reduce over the warp)
        T s[QT];
#pragma unroll
        for (int r = 0; r < QT; ++r) {
            T p = (T)0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) p = tr_fma<T>(m[v][r], f1[v][r], p);
            s[r] = p;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int r = 0; r < QT; ++r) s[r] += __shfl_xor_sync(TR_FULL, s[r], off);
        if (a.U && lane <= g.RT) {
            T uv = (T)1;
#pragma unroll
            for (int r = 0; r < QT; ++r) if (r == lane && r < g.RT) uv = s[r];
            a.U[t * (g.RT + 1) + lane] = uv;
        }
        // outputs and residuals: lanes along n; ds[r] = sum_n res[n] F2[n,r]
        T ds[QT];
#pragma unroll
        for (int r = 0; r < QT; ++r) ds[r] = (T)0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int n = lane + 32 * k;
            if (n < g.NO) {
                const T* f2 = sF2 + (size_t)n * QT;
                T yh = sB[n];
#pragma unroll
                for (int r = 0; r < QT; ++r) yh = tr_fma<T>(s[r], f2[r], yh);
                const T rr = yh - yv[k];
                if (a.yhat) a.yhat[t * g.NO + n] = yh;
                a.res[t * g.NO + n] = rr;
                loss += (double)rr * (double)rr;
#pragma unroll
                for (int r = 0; r < QT; ++r) ds[r] = tr_fma<T>(rr, f2[r], ds[r]);
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int r = 0; r < QT; ++r) ds[r] += __shfl_xor_sync(TR_FULL, ds[r], off);
        // second-mode gradient (registers) and da (global):
        // da of a normal channel r = ds[r] F1[d,r]; of the channels (r, c) of a spectral component = ds[r] F1[d,r] / m[d,r] * a
#pragma unroll
        for (int r = 0; r < QT; ++r) {
            T kf[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                accF[v][r] = tr_fma<T>(ds[r], m[v][r], accF[v][r]);
                kf[v] = ds[r] * f1[v][r] * rinv[v][r];
            }
            if (act && r < g.Rn) {
                SpecVec<T, VEC>::st(a.DA + ((size_t)t * g.Q + r) * g.D + d0, kf);
            } else if (act && r < g.RT) {
                const int qb = g.Rn + (r - g.Rn) * g.CC;
                for (int c = 0; c < g.CC; ++c) {
                    T av[VEC];
                    SpecSm<T, VEC>::ld(sA + (size_t)(qb + c) * 32 * VEC, av);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) av[v] *= kf[v];
                    SpecVec<T, VEC>::st(a.DA + ((size_t)t * g.Q + qb + c) * g.D + d0, av);
                }
            }
        }
        __syncwarp();
        if (--left == 0) {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int r = 0; r < QT; ++r) {
                    slot[(size_t)r * 32 * VEC + d0 + v] += (double)accF[v][r];
                    accF[v][r] = (T)0;
                }
            left = a.spc;
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int r = 0; r < QT; ++r) slot[(size_t)r * 32 * VEC + d0 + v] += (double)accF[v][r];
    loss = warp_sum(loss);
    if (lane == 0) sloss[wid] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < TR_WPB; ++i) tot += sloss[i];
        a.losspart[blockIdx.x] = tot;
    }
}

// gradsum[Fn1 | Fc1] = sum over the warps' slots of df1part (warps, QT, TILE); one block per (r, d)
#ifndef TR_TEMPLATES_ONLY
__global__ void __launch_bounds__(128) k_spec_df1_fold(const double* __restrict__ part, int nslots, int QT, int TILE, SpecGeo g,
                                                       double* __restrict__ gradsum) {
    __shared__ double sbuf[32];
    const int r = blockIdx.x / g.D, d = blockIdx.x % g.D;
    double s = 0.0;
    for (int b = threadIdx.x; b < nslots; b += blockDim.x) s += part[((size_t)b * QT + r) * TILE + d];
    s = block_sum(s, sbuf);
    if (threadIdx.x == 0) {
        if (r < g.Rn) gradsum[g.off[1] + d * g.Rn + r] = s;
        else gradsum[g.off[4] + d * g.Rs + (r - g.Rn)] = s;
    }
}
#endif

// gradsum[Fn0 | Fc0] from the gradient pass's slots (WTN * Gn, TRS_WT, QT): one block per (w, q)
#ifndef TR_TEMPLATES_ONLY
__global__ void __launch_bounds__(128) k_spec_dg_fold(const double* __restrict__ part, int WTN, int Gn, int QT, int q0, SpecGeo g,
                                                      double* __restrict__ gradsum) {
    __shared__ double sbuf[32];
    const int w = blockIdx.x / QT, ql = blockIdx.x % QT, q = q0 + ql;
    if (q >= g.Q) return;
    const int wt = w / TRS_WT, i = w % TRS_WT;
    double s = 0.0;
    for (int grp = threadIdx.x; grp < Gn; grp += blockDim.x) s += part[(((size_t)grp * WTN + wt) * TRS_WT + i) * QT + ql];
    s = block_sum(s, sbuf);
    if (threadIdx.x == 0) {
        if (q < g.Rn) gradsum[g.off[0] + w * g.Rn + q] = s;
        else gradsum[g.off[3] + w * (g.Rs * g.CC) + (q - g.Rn)] = s;
    }
}
#endif
