// tr_flow.cuh — single-launch dataflow kernel: forward contraction, per-sample epilogue and gradient
// accumulation of one fit iteration in ONE persistent kernel, with the second read of X served
// from L2 instead of HBM.
//
// The two-pass path reads X twice from HBM because the gradient weight of sample n (residual /
// v[n,r]) needs that sample's complete inner products.  Here both passes run concurrently, a
// bounded number of samples apart:
//
//   * every block has TR_FLOW_PAIRS forward warps and as many gradient warps; pair p of the grid owns
//     one (warp tile t, sample group g) item exactly like the two-pass kernels, with the CP
//     coefficients (forward warp) / the G accumulators (gradient warp) of the tile in registers;
//   * a forward warp streams its tile of samples g, g+Gn, ... from HBM and writes each tile partial
//     into a small ring (L2 resident) as a 64-bit word {value bits, tag = sample index + 1}: one
//     atomic store, so a reader that sees the tag sees the value — no fence (MEMBAR.GPU costs
//     microseconds under load) and no separate flag.  A relaxed fire-and-forget RED bumps the
//     sample's arrival counter, which is only a HINT that tells the owner when to look;
//   * when the hint says all WT tiles of a sample have arrived, the gradient warp that OWNS the
//     sample (tile == sample index mod WT: a fixed assignment, so the loss partial sums are
//     deterministic) runs the per-sample epilogue (tr_epi.cuh) over the tagged words (a word whose
//     tag is not there yet is simply re-read) and writes V[n,:];
//   * V is pre-filled with an all-ones bit pattern (a NaN no epilogue ever stores): a gradient warp
//     loads V[n,:] TOGETHER with its tile of X[n] — `lag` samples behind its forward warp, i.e.
//     while the lines are still in the 126 MB L2 — and only if a word still holds the pattern does
//     it fall into a polling path.  The data word is its own ready flag, so the streaming loop has
//     no dependent flag load and no fence; then G += v[n,:] * x;
//   * the forward warp never runs more than `lag` samples ahead of its gradient warp (progress word
//     in shared memory), which bounds the window of X that has to stay cached to
//     lag * Gn * D * sizeof(T) bytes, and makes ring-slot reuse safe (ring >= lag + UF: a slot is
//     rewritten only after the paired gradient warp consumed V of the slot's previous sample, i.e.
//     after that sample's epilogue finished reading).
//
// All waits point to strictly earlier samples (or to arrivals that do not wait on anything), so the
// kernel cannot deadlock as long as every block is resident: the host launches at most
// (resident blocks per SM) x (SMs) blocks.  A spin that lasts seconds traps instead of hanging.
#pragma once
#include "tr_kernels.cuh"
#include "tr_epi.cuh"

#define TR_FLOW_PAIRS (TR_WPB / 2)
#define TR_FLOW_SPIN_SMEM (1u << 28)     // polls of a shared-memory word (seconds) before trapping
#define TR_FLOW_SPIN_GMEM (1u << 24)     // polls of a global word (seconds) before trapping

template <typename T>
struct FlowArgs {
    const T* X;
    long long N;
    const T* FtT;        // softplus-ed feature factors
    const T* w;          // rank weights
    Geo geo;
    int mode;            // 0: standard (one channel), 1: multinomial (R channels)
    int WT, Gn;          // WT * Gn items <= gridDim.x * TR_FLOW_PAIRS
    unsigned long long* partial;   // ring (Gn, ring, WT, RK, sizeof(T)/4) of tagged words, zero at launch
    int ring;            // ring slots per group: multiple of UF, >= lag + UF
    unsigned* cnt;       // (N) tile arrivals per sample (a hint, see above), zero at launch
    int lag;             // group-samples the forward warp may lead its gradient warp
    T* V;                // (N, RK) gradient weights, all-ones bit pattern at launch; written by the epilogue
    T* Gpart;            // (nchunk*Gn, RK, Dpad)
    long long Dpad;
    int nchunk;
    long long spc;
    EpiStdArgs<T> es;
    EpiMnArgs<T> em;
    double* losspart;    // (gridDim.x * TR_WPB, 2)
    int dbg;             // timing experiments only (results are wrong when non-zero): 2 epilogue writes zeros,
                         // 4 gradient warps do not wait for V, 8 no forward throttle
};

// V words double as ready flags: relaxed gpu-scope loads (served by L2), all-ones = not written yet
template <typename T> struct FlowWord;
template <> struct FlowWord<float> {
    static __device__ __forceinline__ float ld(const float* p) {
        float v;
        asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
        return v;
    }
    static __device__ __forceinline__ bool pending(float v) { return __float_as_uint(v) == 0xffffffffu; }
};
template <> struct FlowWord<double> {
    static __device__ __forceinline__ double ld(const double* p) {
        double v;
        asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
        return v;
    }
    static __device__ __forceinline__ bool pending(double v) {
        return (unsigned long long)__double_as_longlong(v) == 0xffffffffffffffffull;
    }
};
__device__ __forceinline__ unsigned flow_ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void flow_red_add(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v));
}

// Tagged partials: sizeof(T)/4 64-bit words per value, each {32 value bits, 32-bit tag}, stored and
// loaded with single 64-bit accesses (single-copy atomic).
template <typename T> struct FlowTagged;
template <> struct FlowTagged<float> {
    static constexpr int W = 1;
    static __device__ __forceinline__ void st(unsigned long long* p, float v, unsigned tag) {
        const unsigned long long w = ((unsigned long long)tag << 32) | __float_as_uint(v);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w));
    }
    static __device__ __forceinline__ bool ld(const unsigned long long* p, unsigned tag, float& out) {
        unsigned long long w;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p));
        out = __uint_as_float((unsigned)w);
        return (unsigned)(w >> 32) == tag;
    }
};
template <> struct FlowTagged<double> {
    static constexpr int W = 2;
    static __device__ __forceinline__ void st(unsigned long long* p, double v, unsigned tag) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v);
        const unsigned long long w0 = ((unsigned long long)tag << 32) | (b & 0xffffffffull);
        const unsigned long long w1 = ((unsigned long long)tag << 32) | (b >> 32);
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w0));
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p + 1), "l"(w1));
    }
    static __device__ __forceinline__ bool ld(const unsigned long long* p, unsigned tag, double& out) {
        unsigned long long w0, w1;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0) : "l"(p));
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w1) : "l"(p + 1));
        out = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
        return (unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag;
    }
};

// epilogue-side view of one ring slot (tr_epi.cuh reader concept)
template <typename T>
struct TaggedPartials {
    const unsigned long long* p;   // slot base: (WT, RK, W)
    int RK;
    unsigned tag;
    __device__ __forceinline__ bool get(int t, int r, T& out) const {
        return FlowTagged<T>::ld(p + ((long long)t * RK + r) * FlowTagged<T>::W, tag, out);
    }
};

// ---- forward role ---------------------------------------------------------------------------
template <typename T, int RK, int E, int U, int VEC>
__device__ __forceinline__ void flow_forward(const FlowArgs<T>& a, const T* sF, const T* sW, const int* sDims,
                                             const int* sOff, volatile int* gprog, volatile int* fprog, int t, int g,
                                             int Sg, int lane) {
    constexpr int TILE = 32 * E * VEC;
    constexpr int RKR = tr_next_pow2(RK);
    constexpr int M = tr_next_pow2(U * RKR);
    constexpr int LGM = tr_log2(M);
    static_assert(M <= 32, "next_pow2(U * next_pow2(RK)) must be <= 32");
    const long long D = a.geo.D;
    const long long tile_base = (long long)t * TILE;

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
                tr_coef_at<T, RK>(sF, sW, sDims, sOff, a.geo.k, a.geo.R, a.mode, (unsigned)i, tmp);
            }
#pragma unroll
            for (int c = 0; c < RK; ++c) coef[j][v][c] = tmp[c];
        }
    }

    const T* xbase = a.X + tile_base + (long long)lane * VEC;
    const long long sstride = (long long)a.Gn * D;               // elements between consecutive group samples
    constexpr int W = FlowTagged<T>::W;
    unsigned long long* pring = a.partial + ((long long)g * a.ring * a.WT * RK + (long long)t * RK) * W;
    const int slot_stride = a.WT * RK * W;
    int slot0 = 0;
    const T* xs = xbase + (long long)g * D;
    unsigned* cntp = a.cnt + g;

    T x[U][E][VEC];
    // issue the loads of the batch that starts at group-sample s (never more than `lag` ahead of the
    // paired gradient warp)
    auto issue = [&](int s) {
        const int lim = s + U - a.lag;
        if (lim > 0 && !(a.dbg & 8)) {
            unsigned spins = 0;
            while (*gprog < lim) {
                __nanosleep(64);
                if (++spins > TR_FLOW_SPIN_SMEM) __trap();
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const T* xp = xs + (long long)u * sstride;
#pragma unroll
            for (int j = 0; j < E; ++j) {
                if (s + u < Sg && ((cmask >> j) & 1u)) {
                    XLoad<T, VEC>::ld(xp + j * 32 * VEC, x[u][j]);
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[u][j][v] = (T)0;
                }
            }
        }
        xs += (long long)U * sstride;
    };

    if (Sg > 0) issue(0);
    for (int s = 0; s < Sg; s += U) {
        T vals[M];
#pragma unroll
        for (int q = 0; q < M; ++q) vals[q] = (T)0;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < E; ++j)
#pragma unroll
                for (int v = 0; v < VEC; ++v)
#pragma unroll
                    for (int c = 0; c < RK; ++c)
                        vals[u * RKR + c] = tr_fma<T>(x[u][j][v], coef[j][v][c], vals[u * RKR + c]);
        // next batch in flight before this one is reduced, stored and announced (the fence below then
        // waits for loads the warp would wait for anyway)
        if (s + U < Sg) issue(s + U);
        warp_reduce_transpose<T, M>(vals, lane);
        if ((lane & ((1 << (5 - LGM)) - 1)) == 0) {
            const int q = lane >> (5 - LGM);
            const int u = q / RKR, c = q % RKR;
            if (u < U && c < RK && s + u < Sg)
                FlowTagged<T>::st(pring + (long long)(slot0 + u) * slot_stride + c * W, vals[0], (unsigned)(s + u + 1));
        }
        __syncwarp();
        if (lane == 0) {
            // arrival hints: relaxed, no fence — the tagged words carry the actual synchronisation
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (s + u < Sg) flow_red_add(cntp + (long long)u * a.Gn, 1u);
            *fprog = s + U;
        }
        slot0 += U;
        if (slot0 >= a.ring) slot0 = 0;
        cntp += (long long)U * a.Gn;
    }
}

// ---- gradient role (+ the per-sample epilogues this warp owns) -------------------------------
// The epilogue argument blocks live in shared memory (copied once per block): taking the address of
// the kernel parameter struct for this out-of-line call would make the compiler keep a local-memory
// copy of ALL kernel parameters and read them from there in the streaming loops.
template <typename T>
struct FlowEpiShared {
    EpiStdArgs<T> es;
    EpiMnArgs<T> em;
};

template <typename T>
__device__ __noinline__ void flow_epilogue(const FlowEpiShared<T>* se, int mode, const unsigned long long* p, int RK,
                                           unsigned tag, long long n, int lane, const double* sFC, const double* sWd,
                                           double* sLoss /* [2][32] of this warp */) {
    double l1 = 0.0, l2 = 0.0;
    const TaggedPartials<T> rd{p, RK, tag};
    if (mode == 0) epi_std_sample<T>(se->es, n, rd, lane, (double)se->es.theta[se->es.bias_off], l1, l2);
    else epi_mn_sample<T>(se->em, n, rd, lane, sFC, sWd, l1);
    sLoss[lane] += l1;                                           // per-lane running sums, fixed order: deterministic
    sLoss[32 + lane] += l2;
}

template <typename T, int RK, int E, int U, int VEC>
__device__ __forceinline__ void flow_gradient(const FlowArgs<T>& a, const FlowEpiShared<T>* sEpi, const double* sFC,
                                              const double* sWd,
                                              volatile int* gprog, volatile int* fprog, int t, int g, int Sg, int lane,
                                              double* sLoss, T* sPark /* E*VEC*RK*32 of this warp */) {
    constexpr int TILE = 32 * E * VEC;
    const long long D = a.geo.D;
    const long long tile_base = (long long)t * TILE;
    unsigned cmask = 0;
#pragma unroll
    for (int j = 0; j < E; ++j)
        if (tile_base + (long long)(j * 32 + lane) * VEC < D) cmask |= 1u << j;
    const T* xbase = a.X + tile_base + (long long)lane * VEC + (long long)g * D;
    const long long sstride = (long long)a.Gn * D;
    int next_epi = t;                                            // group-samples t, t+WT, ... are this warp's

    T acc[E][VEC][RK];
    // Run the next owned epilogue.  The accumulators are parked in shared memory across the out-of-line
    // call (rare path), otherwise the register allocator keeps them in local memory for the whole
    // streaming loop.
    auto run_epilogue = [&]() {
        const long long ne = (long long)g + (long long)next_epi * a.Gn;
#pragma unroll
        for (int j = 0; j < E; ++j)
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int c = 0; c < RK; ++c) sPark[((j * VEC + v) * RK + c) * 32 + lane] = acc[j][v][c];
        const int slot = next_epi % a.ring;
        const unsigned long long* p = a.partial + ((long long)g * a.ring + slot) * a.WT * RK * FlowTagged<T>::W;
        if (a.dbg & 2) { if (lane < RK) a.V[ne * RK + lane] = (T)0; }
        else flow_epilogue<T>(sEpi, a.mode, p, RK, (unsigned)(next_epi + 1), ne, lane, sFC, sWd, sLoss);
#pragma unroll
        for (int j = 0; j < E; ++j)
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int c = 0; c < RK; ++c) acc[j][v][c] = sPark[((j * VEC + v) * RK + c) * 32 + lane];
        next_epi += a.WT;
    };
    // polling form: true when an epilogue was run
    auto service = [&]() -> bool {
        if (next_epi >= Sg) return false;
        if (flow_ld_relaxed(a.cnt + ((long long)g + (long long)next_epi * a.Gn)) != (unsigned)a.WT) return false;
        run_epilogue();
        return true;
    };

    for (int ch = 0; ch < a.nchunk; ++ch) {
#pragma unroll
        for (int j = 0; j < E; ++j)
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int c = 0; c < RK; ++c) acc[j][v][c] = (T)0;
        const long long s0l = (long long)ch * a.spc;
        const int s0 = (int)(s0l < Sg ? s0l : Sg);
        const int s1 = (int)(s0l + a.spc < Sg ? s0l + a.spc : Sg);
        for (int s = s0; s < s1; s += U) {
            // An owned sample the paired forward warp has already passed is (nearly) complete: look at its
            // arrival counter.  The load is issued here and its value used after the FMAs, so the
            // streaming loop never stalls on it.
            const bool look = next_epi < Sg && next_epi < *fprog;
            const int looked_at = next_epi;
            unsigned arrived = 0;
            if (look) arrived = flow_ld_relaxed(a.cnt + ((long long)g + (long long)next_epi * a.Gn));
            T x[U][E][VEC];
            T vv[U][RK];
            auto load_x = [&]() {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const T* xp = xbase + (long long)(s + u) * sstride;
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        if (s + u < s1 && ((cmask >> j) & 1u)) {
                            XLoad<T, VEC>::ld(xp + j * 32 * VEC, x[u][j]);
                        } else {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) x[u][j][v] = (T)0;
                        }
                    }
                }
            };
            // V[n,:] of the batch; returns true while any word still holds the "not written" pattern
            auto load_v = [&]() -> bool {
                bool pend = false;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long long n = (long long)g + (long long)(s + u) * a.Gn;
#pragma unroll
                    for (int c = 0; c < RK; ++c) {
                        vv[u][c] = (s + u < s1) ? FlowWord<T>::ld(a.V + n * RK + c) : (T)0;
                        pend |= FlowWord<T>::pending(vv[u][c]);
                    }
                }
                return pend;
            };
            load_x();
            if (load_v() && !(a.dbg & 4)) {
                // rare: the batch's epilogues are not all done.  Poll (serving owned epilogues meanwhile),
                // then re-issue the X loads: nothing but the parked accumulators survives the calls.
                unsigned spins = 0;
                do {
                    if (!service() && ++spins > TR_FLOW_SPIN_GMEM) __trap();
                } while (load_v());
                load_x();
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
            if (lane == 0) *gprog = (s + U < s1) ? s + U : s1;
            if (look && arrived == (unsigned)a.WT && next_epi == looked_at) run_epilogue();   // (the polling path may have run it)
        }
        T* gp = a.Gpart + ((long long)(ch * a.Gn + g) * RK) * a.Dpad + tile_base + (long long)lane * VEC;
#pragma unroll
        for (int c = 0; c < RK; ++c)
#pragma unroll
            for (int j = 0; j < E; ++j)
#pragma unroll
                for (int v = 0; v < VEC; ++v) gp[c * a.Dpad + j * 32 * VEC + v] = acc[j][v][c];
    }
    // every owned sample was needed by this warp's own gradient steps, so nothing is left; kept as a guard
    unsigned spins = 0;
    while (next_epi < Sg)
        if (!service() && ++spins > TR_FLOW_SPIN_GMEM) __trap();
}

template <typename T, int RK, int E, int UF, int UG, int VEC>
__global__ void __launch_bounds__(TR_TPB, TR_MINB) k_flow(const FlowArgs<T> a) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    __shared__ int sDims[TR_MAX_MODES], sOff[TR_MAX_MODES + 2];
    __shared__ volatile int sGprog[TR_FLOW_PAIRS], sFprog[TR_FLOW_PAIRS];
    __shared__ double sLoss[TR_FLOW_PAIRS][64];
    __shared__ FlowEpiShared<T> sEpi;
    T* sF = reinterpret_cast<T*>(tr_smem);
    const int R = a.geo.R, pfeat = a.geo.pfeat, C = a.geo.C;
    const int nF = pfeat + R;
    double* sFC = reinterpret_cast<double*>(tr_smem + (((size_t)nF * sizeof(T) + 15) / 16) * 16);   // C*R
    double* sWd = sFC + C * R;                                                                         // R
    T* sParkAll = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(sWd + R) + ((C * R + R) & 1) * 8);   // 16-byte aligned
    for (int i = threadIdx.x; i < nF; i += TR_TPB) sF[i] = i < pfeat ? a.FtT[i] : a.w[i - pfeat];
    for (int i = threadIdx.x; i < C * R + R; i += TR_TPB)
        sFC[i] = i < C * R ? a.em.FC[i] : (double)a.w[i - C * R];
    if (threadIdx.x < TR_MAX_MODES) sDims[threadIdx.x] = a.geo.dims[threadIdx.x];
    if (threadIdx.x < TR_MAX_MODES + 2) sOff[threadIdx.x] = a.geo.foff[threadIdx.x];
    if (threadIdx.x < TR_FLOW_PAIRS) { sGprog[threadIdx.x] = 0; sFprog[threadIdx.x] = 0; }
    if (threadIdx.x == 0) { sEpi.es = a.es; sEpi.em = a.em; }
    __syncthreads();

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pair = wid % TR_FLOW_PAIRS, role = wid / TR_FLOW_PAIRS;
    const long long item = (long long)blockIdx.x * TR_FLOW_PAIRS + pair;
    if (role == 1) { sLoss[pair][lane] = 0.0; sLoss[pair][32 + lane] = 0.0; }
    __syncwarp();
    if (item < (long long)a.WT * a.Gn) {
        const int t = (int)(item % a.WT);
        const int g = (int)(item / a.WT);
        const int Sg = g < a.N ? (int)((a.N - g + a.Gn - 1) / a.Gn) : 0;
        if (role == 0)
            flow_forward<T, RK, E, UF, VEC>(a, sF, sF + pfeat, sDims, sOff, &sGprog[pair], &sFprog[pair], t, g, Sg, lane);
        else
            flow_gradient<T, RK, E, UG, VEC>(a, &sEpi, sFC, sWd, &sGprog[pair], &sFprog[pair], t, g, Sg, lane, sLoss[pair],
                                             sParkAll + (size_t)pair * (E * VEC * RK * 32));
    }
    __syncwarp();
    double l1 = role == 1 ? sLoss[pair][lane] : 0.0;            // the CE term is added by the lane of the true class
    double l2 = role == 1 ? sLoss[pair][32 + lane] : 0.0;
    l1 = warp_sum(l1);
    l2 = warp_sum(l2);
    if (lane == 0) {
        const long long wg = (long long)blockIdx.x * TR_WPB + wid;
        a.losspart[wg * 2 + 0] = l1;
        a.losspart[wg * 2 + 1] = l2;
    }
}
