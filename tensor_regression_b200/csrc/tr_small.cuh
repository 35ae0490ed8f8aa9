// tr_small.cuh — the small kernels around the two streaming passes: factor preparation,
// per-sample epilogues, split-N reduction, all-mode MTTKRP, finish (normalise + penalty) and Adam.
// Included by tr_api.cu only.
#pragma once
#include "tr_kernels.cuh"

// ---------------------------------------------------------------------------------------------
// k_prep: Ft = softplus?(theta) for the Pf factor entries, in T (for the streaming kernels) and
// in double (for the small finishing kernels).  non_neg_fn, std:53-85 / mn:116-146.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_prep(const T* __restrict__ theta, Geo g, uint32_t nn_mask, double beta, double thr,
                       T* __restrict__ FtT, double* __restrict__ Ft64) {
    const int nfac = g.k + (g.C > 0 ? 1 : 0);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < g.pf; p += gridDim.x * blockDim.x) {
        int m = 0;
        while (m + 1 < nfac && p >= g.foff[m + 1]) ++m;
        T x = theta[p];
        if ((nn_mask >> m) & 1u) x = tr_softplus<T>(x, (T)beta, (T)thr);
        FtT[p] = x;
        Ft64[p] = (double)x;
    }
}

// Gred[c, i] = sum_slots Gpart[slot, c, i]  (double accumulation; Gpart is L2-resident)
template <typename T>
__global__ void k_reduce_G(const T* __restrict__ Gpart, int slots, int RK, long long D, long long Dpad,
                           double* __restrict__ Gred) {
    const long long total = (long long)RK * D;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e / D);
        const long long i = e % D;
        // slots in order (deterministic); four loads in flight per thread
        double s = 0.0;
        const T* gp = Gpart + (long long)c * Dpad + i;
        const long long sstr = (long long)RK * Dpad;
        int sl = 0;
        for (; sl + 4 <= slots; sl += 4) {
            const T v0 = __ldcg(gp + (long long)sl * sstr), v1 = __ldcg(gp + (long long)(sl + 1) * sstr);
            const T v2 = __ldcg(gp + (long long)(sl + 2) * sstr), v3 = __ldcg(gp + (long long)(sl + 3) * sstr);
            s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
        }
        for (; sl < slots; ++sl) s += (double)__ldcg(gp + (long long)sl * sstr);
        Gred[e] = s;
    }
}

#include "tr_epi.cuh"

// dFC_part[b, c*R + r] = w_r * sum_{n in block b's range} dZ[n,c] * u[n,r]   (class-factor gradient).
// Thread (q, j): sample lane q = tid / 64 of 4, (c,r) pair j = tid % 64 (+64, ...); the four sample lanes
// are combined in a fixed order through shared memory (4 * C * R doubles of dynamic shared memory).
template <typename T>
__global__ void __launch_bounds__(TR_TPB) k_dfc(const T* __restrict__ dZ, const T* __restrict__ u,
                                                const T* __restrict__ w, long long N, int C, int R,
                                                double* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    double* sq = reinterpret_cast<double*>(tr_smem);                // [4][C*R]
    const long long per = (N + gridDim.x - 1) / gridDim.x;
    const long long n0 = (long long)blockIdx.x * per;
    const long long n1 = n0 + per < N ? n0 + per : N;
    const int q = threadIdx.x >> 6, CR = C * R;
    for (int j = threadIdx.x & 63; j < CR; j += 64) {
        const int c = j / R, r = j % R;
        double s = 0.0;
#pragma unroll 4
        for (long long n = n0 + q; n < n1; n += 4) s += (double)__ldg(dZ + n * C + c) * (double)__ldg(u + n * R + r);
        sq[q * CR + j] = s;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < CR; j += TR_TPB)
        part[(long long)blockIdx.x * CR + j] = (((sq[j] + sq[CR + j]) + sq[2 * CR + j]) + sq[3 * CR + j]) * (double)w[j % R];
}

// out[j] = sum_b part[b, j] for a (rows, cols) double matrix; one block per column (deterministic)
#ifndef TR_TEMPLATES_ONLY
__global__ void __launch_bounds__(128) k_colsum(const double* __restrict__ part, int rows, int cols,
                                                 double* __restrict__ out) {
    __shared__ double sbuf[32];
    const int j = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < rows; b += blockDim.x) s += part[(long long)b * cols + j];
    s = block_sum(s, sbuf);
    if (threadIdx.x == 0) out[j] = s;
}
#endif

// part[b] = { sum res, sum res^2 } over block b's grid-stride share of a residual vector (fp64)
template <typename T>
__global__ void __launch_bounds__(256) k_ressum(const T* __restrict__ res, long long N, double* __restrict__ part) {
    __shared__ double sbuf[32];
    double l1 = 0.0, l2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double r = (double)res[i];
        l1 += r;
        l2 += r * r;
    }
    const double t1 = block_sum(l1, sbuf);
    const double t2 = block_sum(l2, sbuf);
    if (threadIdx.x == 0) { part[blockIdx.x * 2 + 0] = t1; part[blockIdx.x * 2 + 1] = t2; }
}

// sum of a T vector (for tr_backward_std's dbias) -> out[0]; single block
template <typename T>
__global__ void __launch_bounds__(1024) k_vecsum(const T* __restrict__ v, long long N, double* __restrict__ out) {
    __shared__ double sbuf[32];
    double s = 0.0;
    for (long long i = threadIdx.x; i < N; i += blockDim.x) s += (double)v[i];
    s = block_sum(s, sbuf);
    if (threadIdx.x == 0) out[0] = s;
}

// ---------------------------------------------------------------------------------------------
// All-mode MTTKRP of the reduced G against the factors (the cp_to_tensor backward, K6).
// One block per (mode m, row i_m):  dFt_m[i_m, r] = scale_r * sum_{i : i_m fixed} G(r)[i] prod_{j!=m} Ft_j[i_j, r]
// ---------------------------------------------------------------------------------------------
struct MtArgs {
    const double* G;      // (RKs, D) if per_rank else (D)
    const double* Ft64;
    const void* w;        // rank weights (T) — applied when !per_rank
    int w_is_f64;
    int per_rank;
    Geo geo;
    double* gradsum;
};

#ifndef TR_TEMPLATES_ONLY
__global__ void __launch_bounds__(TR_TPB) k_mttkrp(const MtArgs a) {
    __shared__ double sbuf[32];
    const int k = a.geo.k, R = a.geo.R;
    int b = blockIdx.x, m = 0;
    while (b >= a.geo.dims[m]) { b -= a.geo.dims[m]; ++m; }
    const int im = b;
    long long stride[TR_MAX_MODES];
    stride[k - 1] = 1;
    for (int j = k - 2; j >= 0; --j) stride[j] = stride[j + 1] * a.geo.dims[j + 1];
    const long long D = a.geo.D;
    const long long S = D / a.geo.dims[m];
    for (int rb = 0; rb < R; rb += 8) {
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.0;
        for (unsigned s = threadIdx.x; s < (unsigned)S; s += TR_TPB) {     // D < 2^31: 32-bit index arithmetic
            unsigned rem = s;
            long long lin = (long long)im * stride[m];
            int idx[TR_MAX_MODES];
            for (int j = k - 1; j >= 0; --j) {
                if (j == m) continue;
                const unsigned dj = (unsigned)a.geo.dims[j];
                const unsigned q = rem / dj;
                idx[j] = (int)(rem - q * dj);
                rem = q;
                lin += idx[j] * stride[j];
            }
            const double g0 = a.per_rank ? 0.0 : a.G[lin];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = rb + q;
                if (r < R) {
                    double p = a.per_rank ? a.G[(long long)r * D + lin] : g0;
                    for (int j = 0; j < k; ++j)
                        if (j != m) p *= a.Ft64[a.geo.foff[j] + idx[j] * R + r];
                    acc[q] += p;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = rb + q;
            if (r < R) {                                   // uniform across the block
                const double tot = block_sum(acc[q], sbuf);
                if (threadIdx.x == 0) {
                    double sc = 1.0;
                    if (!a.per_rank)
                        sc = a.w_is_f64 ? ((const double*)a.w)[r] : (double)((const float*)a.w)[r];
                    a.gradsum[a.geo.foff[m] + im * R + r] = tot * sc;
                }
            }
        }
    }
}
#endif

// ---------------------------------------------------------------------------------------------
// Finish: normalisation, softplus chain rule, penalty gradient and value (single block).
// L2_penalty std:180-196 (sum of un-squared Frobenius norms of the RAW factors; bias excluded).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) k_finish(const double* __restrict__ gradsum, int n_gs, int nbias, double grad_scale,
                                                 double loss_scale, const T* __restrict__ theta, Geo g,
                                                 double lambda, uint32_t nn_mask, double beta, double thr,
                                                 T* __restrict__ grad, double* __restrict__ loss) {
    __shared__ double sbuf[32];
    __shared__ double snorm[TR_MAX_MODES + 1];
    const int nfac = g.k + (g.C > 0 ? 1 : 0);
    for (int m = 0; m < nfac; ++m) {
        const int p0 = g.foff[m], p1 = (m + 1 < nfac) ? g.foff[m + 1] : g.pf;
        double s = 0.0;
        for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) { const double x = (double)theta[p]; s += x * x; }
        s = block_sum(s, sbuf);
        if (threadIdx.x == 0) snorm[m] = sqrt(s);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < g.pf; p += blockDim.x) {
        int m = 0;
        while (m + 1 < nfac && p >= g.foff[m + 1]) ++m;
        const double x = (double)theta[p];
        double d = gradsum[p] * grad_scale;
        if ((nn_mask >> m) & 1u) {
            const double bx = x * beta;
            if (!(bx > thr)) d *= 1.0 / (1.0 + exp(-bx));
        }
        if (lambda != 0.0) d += lambda * x / snorm[m];   // lambda == 0: no penalty term (also keeps the VJP path finite at F == 0)
        grad[p] = (T)d;
    }
    // bias entries (one for the standard model, n_out for the spectral one; none for the multinomial model)
    for (int b = threadIdx.x; b < nbias; b += blockDim.x) grad[g.pf + b] = (T)(gradsum[g.pf + b] * grad_scale);
    if (threadIdx.x == 0) {
        double pen = 0.0;
        for (int m = 0; m < nfac; ++m) pen += snorm[m];
        const double ld = gradsum[n_gs - 1] * loss_scale;
        loss[0] = ld;
        loss[1] = ld + lambda * pen;
    }
}

// ---------------------------------------------------------------------------------------------
// Adam (torch/optim/adam.py _single_tensor_adam; arithmetic in the parameter dtype like torch)
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T tr_sqrt(T x);
template <> __device__ __forceinline__ float tr_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double tr_sqrt<double>(double x) { return sqrt(x); }

// Per-parameter-group step sizes (torch.optim.Adam with one parameter group per factor, e.g. the three groups of
// hier:436-440): group g = factor g of theta (boundaries seg_end[], the bias is the last group of the standard
// model); n_seg == 0 means one group with step size `step_size`.
struct AdamGroups {
    int n_seg;
    int seg_end[TR_MAX_MODES + 2];
    double step_size[TR_MAX_MODES + 2];
};

template <typename T>
__global__ void k_adam(T* __restrict__ theta, const T* __restrict__ grad, T* __restrict__ m, T* __restrict__ v,
                       T* __restrict__ vmax, long long P, double beta1, double beta2, double eps, double wd,
                       double step_size, double bc2_sqrt, const AdamGroups groups) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P;
         p += (long long)gridDim.x * blockDim.x) {
        if (groups.n_seg > 0) {
            int sgi = 0;
            while (sgi + 1 < groups.n_seg && p >= groups.seg_end[sgi]) ++sgi;
            step_size = groups.step_size[sgi];
        }
        T g = grad[p];
        const T th = theta[p];
        if (wd != 0.0) g = g + (T)wd * th;
        T mm = m[p];
        mm = mm + (T)(1.0 - beta1) * (g - mm);                       // lerp_(grad, 1-beta1)
        T vv = v[p] * (T)beta2;
        vv = vv + (T)(1.0 - beta2) * g * g;                           // mul_(beta2).addcmul_(g, g, 1-beta2)
        m[p] = mm;
        v[p] = vv;
        T dn;
        if (vmax) {
            T vm = vmax[p];
            vm = vm > vv ? vm : vv;
            vmax[p] = vm;
            dn = tr_sqrt<T>(vm) / (T)bc2_sqrt + (T)eps;
        } else {
            dn = tr_sqrt<T>(vv) / (T)bc2_sqrt + (T)eps;
        }
        theta[p] = th + (T)(-step_size) * (mm / dn);                  // addcdiv_(m, denom, value=-step_size)
    }
}


// max that propagates NaN like torch's flat_grad.abs().max(): a NaN gradient must not read as "max|g| = 0"
// (lbfgs.py would then stop with "optimality reached" on a corrupted state; torch compares NaN <= tol -> False)
__device__ __forceinline__ double tr_nanmax(double a, double b) { return (a != a || b != b) ? NAN : fmax(a, b); }

// ---------------------------------------------------------------------------------------------
// L-BFGS on the flat parameter vector (torch/optim/lbfgs.py:333-536, used by std:366,392 and
// mn:355,381): history, two-loop recursion and every dot product stay on the device; only the few
// scalars the strong-Wolfe control flow branches on are read back by the host.
// lstate (double): [0]=H_diag, [1]=num_old, [2]=head (ring start), [3]=unused, [4 .. 4+hist)=ro
// ---------------------------------------------------------------------------------------------
#define TR_LBFGS_MAX_HIST 256

__device__ __forceinline__ double block_sum_bcast(double v, double* sbuf, double* sb) {
    v = block_sum(v, sbuf);
    if (threadIdx.x == 0) *sb = v;
    __syncthreads();
    const double r = *sb;
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_max_bcast(double v, double* sbuf, double* sb) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = tr_nanmax(v, __shfl_xor_sync(TR_FULL, v, off));
    __syncthreads();
    if (lane == 0) sbuf[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double r = lane < nw ? sbuf[lane] : 0.0;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) r = tr_nanmax(r, __shfl_xor_sync(TR_FULL, r, off));
        if (lane == 0) *sb = r;
    }
    __syncthreads();
    const double r = *sb;
    __syncthreads();
    return r;
}

// One "compute the search direction" step (lbfgs.py:395-440) in a single block.
//   first != 0 : d = -g, history cleared, H_diag = 1
//   otherwise  : y = g - prev_g, s = d*t; if y.s > 1e-10 push (y, s, 1/y.s) and H_diag = y.s / y.y;
//                two-loop recursion -> d
// then prev_g = g and scal = { g.d, sum|g|, max|g|, max|d| }.
template <typename T>
__global__ void __launch_bounds__(1024) k_lbfgs_direction(const T* __restrict__ g, T* __restrict__ prev_g,
                                                          T* __restrict__ d, double t, int first, T* __restrict__ S,
                                                          T* __restrict__ Y, double* __restrict__ ls, int hist,
                                                          long long P, double* __restrict__ scal) {
    __shared__ double sbuf[32];
    __shared__ double sb;
    __shared__ double s_al[TR_LBFGS_MAX_HIST];
    double* ro = ls + 4;
    int num_old = (int)ls[1], head = (int)ls[2];
    double H_diag = ls[0];
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();                                   // every thread read the state before it is rewritten
    if (first) {
        for (long long p = tid; p < P; p += nt) d[p] = -g[p];
        num_old = 0; head = 0; H_diag = 1.0;
    } else {
        double ys = 0.0, yy = 0.0;
        for (long long p = tid; p < P; p += nt) {
            const T y = g[p] - prev_g[p];
            const T s = d[p] * (T)t;
            ys += (double)y * (double)s;
            yy += (double)y * (double)y;
        }
        ys = block_sum_bcast(ys, sbuf, &sb);
        yy = block_sum_bcast(yy, sbuf, &sb);
        if (ys > 1e-10) {
            int slot;
            if (num_old == hist) { slot = head; head = (head + 1) % hist; }
            else { slot = (head + num_old) % hist; ++num_old; }
            for (long long p = tid; p < P; p += nt) {
                Y[(long long)slot * P + p] = g[p] - prev_g[p];
                S[(long long)slot * P + p] = d[p] * (T)t;
            }
            if (tid == 0) ro[slot] = 1.0 / ys;
            H_diag = ys / yy;
            __syncthreads();
        }
        // two-loop recursion; q lives in d
        for (long long p = tid; p < P; p += nt) d[p] = -g[p];
        for (int i = num_old - 1; i >= 0; --i) {
            const long long o = (long long)((head + i) % hist) * P;
            double a = 0.0;
            for (long long p = tid; p < P; p += nt) a += (double)S[o + p] * (double)d[p];
            a = block_sum_bcast(a, sbuf, &sb) * ro[(head + i) % hist];
            if (tid == 0) s_al[i] = a;
            for (long long p = tid; p < P; p += nt) d[p] = d[p] - (T)a * Y[o + p];
        }
        for (long long p = tid; p < P; p += nt) d[p] = d[p] * (T)H_diag;
        __syncthreads();
        for (int i = 0; i < num_old; ++i) {
            const long long o = (long long)((head + i) % hist) * P;
            double be = 0.0;
            for (long long p = tid; p < P; p += nt) be += (double)Y[o + p] * (double)d[p];
            be = block_sum_bcast(be, sbuf, &sb) * ro[(head + i) % hist];
            const double c = s_al[i] - be;
            for (long long p = tid; p < P; p += nt) d[p] = d[p] + (T)c * S[o + p];
        }
    }
    double gtd = 0.0, g1 = 0.0, gm = 0.0, dm = 0.0;
    for (long long p = tid; p < P; p += nt) {
        const double gv = (double)g[p], dv = (double)d[p];
        prev_g[p] = g[p];
        gtd += gv * dv;
        g1 += fabs(gv);
        gm = tr_nanmax(gm, fabs(gv));
        dm = tr_nanmax(dm, fabs(dv));
    }
    gtd = block_sum_bcast(gtd, sbuf, &sb);
    g1 = block_sum_bcast(g1, sbuf, &sb);
    gm = block_max_bcast(gm, sbuf, &sb);
    dm = block_max_bcast(dm, sbuf, &sb);
    if (tid == 0) {
        scal[0] = gtd; scal[1] = g1; scal[2] = gm; scal[3] = dm;
        ls[0] = H_diag; ls[1] = (double)num_old; ls[2] = (double)head;
    }
}

// out = x + t * d   (lbfgs.py:325-330 _directional_evaluate / _add_grad)
template <typename T>
__global__ void k_axpy_out(T* __restrict__ out, const T* __restrict__ x, double t, const T* __restrict__ d, long long P) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x)
        out[p] = x[p] + (T)t * d[p];
}

// scal = { g.d, max|g| }  (line-search directional derivative, optimality condition)
template <typename T>
__global__ void __launch_bounds__(1024) k_lbfgs_gtd(const T* __restrict__ g, const T* __restrict__ d, long long P,
                                                    double* __restrict__ scal) {
    __shared__ double sbuf[32];
    __shared__ double sb;
    double gtd = 0.0, gm = 0.0;
    for (long long p = threadIdx.x; p < P; p += blockDim.x) {
        const double gv = (double)g[p];
        gtd += gv * (d ? (double)d[p] : 0.0);
        gm = tr_nanmax(gm, fabs(gv));
    }
    gtd = block_sum_bcast(gtd, sbuf, &sb);
    gm = block_max_bcast(gm, sbuf, &sb);
    if (threadIdx.x == 0) { scal[0] = gtd; scal[1] = gm; }
}
