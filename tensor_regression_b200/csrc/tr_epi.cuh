// tr_epi.cuh — per-sample epilogues (one warp per sample): the sum over the warp-tile partials of
// pass 1 followed by everything the reference does between the contraction and the backward
// matmul (bias / softmax / second softmax / class-weighted CE / dP -> dZ -> v).  The per-sample
// bodies are device functions so that the stand-alone epilogue kernels (two-pass path) and the
// dataflow kernel (tr_flow.cuh, where the gradient warps run them in-line) share one arithmetic.
#pragma once
#include "tr_kernels.cuh"

#define TR_JC (TR_MAX_CLASSES / 32)
#define TR_MAX_RANK_MN 32

// Gradient weights are stored with NaNs canonicalised: the dataflow kernel (tr_flow.cuh) uses the all-ones
// bit pattern — also a NaN — as "not written yet".
__device__ __forceinline__ float tr_canon(float v) { return v != v ? __uint_as_float(0x7fffffffu) : v; }
__device__ __forceinline__ double tr_canon(double v) { return v != v ? __longlong_as_double(0x7ff8000000000000ll) : v; }

// Where an epilogue finds the (tile t, channel r) partial of its sample.  PlainPartials: the (WT, RKs)
// array written by k_fwd (complete when the epilogue kernel starts).  The dataflow kernel has its own
// reader (tr_flow.cuh) whose words carry a tag and are waited for one by one.
// get() never blocks: it returns false when the word is not there yet, and the epilogue repeats its
// (fully batched) read of the sample — a loop that waited inside get() would serialise ~80 L2 round trips.
template <typename T>
struct PlainPartials {
    const T* p;       // this sample's (WT, RKs) block
    int RKs;
    __device__ __forceinline__ bool get(int t, int r, T& out) const {
        out = __ldg(p + (long long)t * RKs + r);       // the RKs loads of a tile share sectors: keep them in L1
        return true;
    }
};

template <typename T>
struct EpiStdArgs {
    const T* partial; int WT; long long N;
    const T* theta; int bias_off;
    const T* y;        // may be null (forward only)
    T* yhat;           // may be null
    T* V;              // residual out (N) or null
    double* part;      // (gridDim.x, 2): sum res, sum res^2
};

// lin_model's "+ bias" (std:130) and the MSE residual (std:371,461) of sample n; p = the sample's WT partials
template <typename T, typename Reader>
__device__ __forceinline__ void epi_std_sample(const EpiStdArgs<T>& a, long long n, const Reader& rd, int lane,
                                               double bias, double& l1, double& l2) {
    double s;
    bool ok;
    do {
        s = 0.0;
        ok = true;
#pragma unroll 4
        for (int t = lane; t < a.WT; t += 32) {
            T v;
            ok &= rd.get(t, 0, v);
            s += (double)v;
        }
    } while (!__all_sync(TR_FULL, ok));
    s = warp_sum(s);
    if (lane == 0) {
        const T yh = (T)(s + bias);                // yhat in the model dtype, as the reference returns it
        if (a.yhat) a.yhat[n] = yh;
        if (a.y) {
            const double res = (double)yh - (double)a.y[n];
            if (a.V) a.V[n] = tr_canon((T)res);
            l2 += res * res;
            l1 += res;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(TR_TPB) k_epi_std(const EpiStdArgs<T> a) {
    __shared__ double sbuf[32];
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * TR_WPB + (threadIdx.x >> 5);
    const long long wtot = (long long)gridDim.x * TR_WPB;
    const double bias = (double)a.theta[a.bias_off];
    double l2 = 0.0, l1 = 0.0;
    for (long long n = warp_global; n < a.N; n += wtot) {
        const PlainPartials<T> rd{a.partial + n * a.WT, 1};
        epi_std_sample<T>(a, n, rd, lane, bias, l1, l2);
    }
    const double t2 = block_sum(l2, sbuf);
    const double t1 = block_sum(l1, sbuf);
    if (threadIdx.x == 0 && a.part) { a.part[blockIdx.x * 2 + 0] = t1; a.part[blockIdx.x * 2 + 1] = t2; }
}

template <typename T>
struct EpiMnArgs {
    const T* partial; int WT; int RKs; long long N;
    int R, C;
    const double* FC;   // class factor (C,R), softplus-ed, double
    const T* w;         // rank weights
    const long long* y; // may be null (forward only, or backward with dP_in)
    const T* dP_in;     // (N,C) upstream gradient wrt P (tr_backward_mn) or null
    const T* class_w;   // (C) or null
    T* P;               // (N,C) or null
    long long* pred;    // (N) or null
    T* V;               // (N,RKs) or null
    T* u_ws;            // (N,R) or null
    T* dZ_ws;           // (N,C) or null
    double* part;       // (gridDim.x): sum -omega log Q
};

// Everything after the contraction, given the sample's u[r] (already summed over the feature axis, replicated in
// every lane): logits, softmax, second softmax, loss, dP -> dZ -> v.  yn / omega = the sample's label and class
// weight (loaded by the caller so that the loads can be in flight early; ignored when a.y == nullptr).
// vout (may be null): v[r] = w_r sum_c dZ[c] FC[c,r] for r < R, replicated in every lane.
template <typename T, int RMAX = TR_MAX_RANK_MN, int JC = TR_JC>
__device__ __forceinline__ void epi_mn_core(const EpiMnArgs<T>& a, long long n, const double (&u)[RMAX], int lane,
                                            const double* sFC, const double* sW, int yn, double omega, double& loss,
                                            double* vout) {
    const int R = a.R, C = a.C;
    // logits of this lane's classes, softmax
    double z[JC], P[JC];
    double zmax = -INFINITY;
#pragma unroll
    for (int jc = 0; jc < JC; ++jc) {
        const int c = lane + 32 * jc;
        z[jc] = -INFINITY;
        if (c < C) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < RMAX; ++r)
                if (r < R) s += sW[r] * u[r] * sFC[c * R + r];
            z[jc] = s;
            zmax = fmax(zmax, s);
        }
    }
    zmax = warp_max(zmax);
    double zs = 0.0;
#pragma unroll
    for (int jc = 0; jc < JC; ++jc) {
        const int c = lane + 32 * jc;
        P[jc] = c < C ? exp(z[jc] - zmax) : 0.0;
        zs += P[jc];
    }
    zs = warp_sum(zs);
    double pmax = -INFINITY;
#pragma unroll
    for (int jc = 0; jc < JC; ++jc) {
        const int c = lane + 32 * jc;
        P[jc] = P[jc] / zs;
        if (c < C) {
            P[jc] = (double)(T)P[jc];          // probabilities in the model dtype, as the reference holds them
            pmax = fmax(pmax, P[jc]);
            if (a.P) a.P[n * C + c] = (T)P[jc];
        }
    }
    pmax = warp_max(pmax);
    if (a.pred) {                                // first index of the maximum (np.argmax, mn:527)
        int best = 1 << 30;
#pragma unroll
        for (int jc = 0; jc < JC; ++jc) {
            const int c = lane + 32 * jc;
            if (c < C && P[jc] == pmax && c < best) best = c;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) best = min(best, __shfl_xor_sync(TR_FULL, best, off));
        if (lane == 0) a.pred[n] = best;
    }
    if (a.y == nullptr && a.dP_in == nullptr) return;

    double dP[JC], dot = 0.0;
    if (a.dP_in) {
        // vector-Jacobian product for an arbitrary upstream gradient wrt P (autograd of model, mn:180-187)
#pragma unroll
        for (int jc = 0; jc < JC; ++jc) {
            const int c = lane + 32 * jc;
            dP[jc] = c < C ? (double)a.dP_in[n * C + c] : 0.0;
            dot += dP[jc] * P[jc];
        }
    } else {
        // second softmax (CrossEntropyLoss applied to probabilities, mn:364-366 / 448-450)
        double Q[JC];
        double qs = 0.0;
#pragma unroll
        for (int jc = 0; jc < JC; ++jc) {
            const int c = lane + 32 * jc;
            Q[jc] = c < C ? exp(P[jc] - pmax) : 0.0;
            qs += Q[jc];
        }
        qs = warp_sum(qs);
#pragma unroll
        for (int jc = 0; jc < JC; ++jc) {
            const int c = lane + 32 * jc;
            dP[jc] = 0.0;
            if (c < C) {
                const double q = Q[jc] / qs;
                if (c == yn) loss += -omega * ((P[jc] - pmax) - log(qs));
                dP[jc] = omega * (q - (c == yn ? 1.0 : 0.0));
                dot += dP[jc] * P[jc];
            }
        }
    }
    dot = warp_sum(dot);
    double dZ[JC];
#pragma unroll
    for (int jc = 0; jc < JC; ++jc) {
        const int c = lane + 32 * jc;
        dZ[jc] = c < C ? P[jc] * (dP[jc] - dot) : 0.0;
        if (c < C && a.dZ_ws) a.dZ_ws[n * C + c] = (T)dZ[jc];
    }
    // v[r] = w_r sum_c dZ[c] FC[c,r]
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        if (r < R) {
            double s = 0.0;
#pragma unroll
            for (int jc = 0; jc < JC; ++jc) {
                const int c = lane + 32 * jc;
                if (c < C) s += dZ[jc] * sFC[c * R + r];
            }
            s = warp_sum(s) * sW[r];
            if (vout) vout[r] = s;
            if (lane == 0) {
                if (a.V) a.V[n * a.RKs + r] = tr_canon((T)s);
                if (a.u_ws) a.u_ws[n * R + r] = (T)u[r];
            }
        }
    }
    if (a.V && lane >= R && lane < a.RKs) a.V[n * a.RKs + lane] = (T)0;   // padding channels
}

// model()'s softmax (mn:180-187), CrossEntropyLoss on the probabilities (second softmax, mn:364-366 /
// 448-450) and its backward down to v[n,r], for sample n.  rd = reader of the sample's tile partials;
// sFC (C,R) / sW (R) in shared memory (double).
// RMAX >= R and 32*JC >= C are compile-time bounds of the unrolled loops (register arrays u[RMAX], z/P/dP/dZ[JC]):
// the kernels are instantiated for (8,1), (16,1) and (TR_MAX_RANK_MN, TR_JC); the arithmetic on the live
// entries — and therefore every bit of the result — does not depend on the bounds.
template <typename T, typename Reader, int RMAX = TR_MAX_RANK_MN, int JC = TR_JC>
__device__ __forceinline__ void epi_mn_sample(const EpiMnArgs<T>& a, long long n, const Reader& rd, int lane,
                                              const double* sFC, const double* sW, double& loss) {
    const int R = a.R;
    // u[r] = sum over warp tiles: lane t-strided, all channels of a tile are contiguous, so every
    // lane keeps RKs independent accumulators and several tiles' loads in flight
    double u[RMAX];
    bool ok;
    do {
        ok = true;
#pragma unroll
        for (int r = 0; r < RMAX; ++r) u[r] = 0.0;
        // four tiles per step: their partials are pre-summed in T in a fixed order, ((a+b)+(c+d)), and only
        // the sum is widened — the T -> double conversions (XU pipe) were 36 % of this kernel's issue slots
        for (int t0 = lane; t0 < a.WT; t0 += 128) {
#pragma unroll
            for (int r = 0; r < RMAX; ++r)
                if (r < R) {
                    T v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        v[q] = (T)0;
                        if (t0 + 32 * q < a.WT) ok &= rd.get(t0 + 32 * q, r, v[q]);
                    }
                    u[r] += (double)((v[0] + v[1]) + (v[2] + v[3]));
                }
        }
    } while (!__all_sync(TR_FULL, ok));
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
        if (r < R) u[r] = warp_sum(u[r]);
    int yn = 0;
    double omega = 1.0;
    if (a.y != nullptr && a.dP_in == nullptr) {
        yn = (int)a.y[n];
        omega = a.class_w ? (double)a.class_w[yn] : 1.0;
    }
    epi_mn_core<T, RMAX, JC>(a, n, u, lane, sFC, sW, yn, omega, loss, nullptr);
}

template <typename T, int RMAX = TR_MAX_RANK_MN, int JC = TR_JC>
__global__ void __launch_bounds__(TR_TPB, 3) k_epi_mn(const EpiMnArgs<T> a) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    double* sFC = reinterpret_cast<double*>(tr_smem);      // C*R
    double* sW = sFC + a.C * a.R;                            // R
    __shared__ double sbuf[32];
    for (int i = threadIdx.x; i < a.C * a.R + a.R; i += TR_TPB)
        sFC[i] = i < a.C * a.R ? a.FC[i] : (double)a.w[i - a.C * a.R];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * TR_WPB + (threadIdx.x >> 5);
    const long long wtot = (long long)gridDim.x * TR_WPB;
    double loss = 0.0;
    for (long long n = warp_global; n < a.N; n += wtot) {
        const PlainPartials<T> rd{a.partial + n * a.WT * a.RKs, a.RKs};
        epi_mn_sample<T, PlainPartials<T>, RMAX, JC>(a, n, rd, lane, sFC, sW, loss);
    }
    const double tl = block_sum(loss, sbuf);
    if (threadIdx.x == 0 && a.part) a.part[blockIdx.x] = tl;
}
