// tr_spectral_single.cuh — single-pass fit iteration of the spectral variant (spectral_tensor_regression.py:541-762):
// X is read from HBM ONCE per iteration.
//
// The two-pass path (tr_spectral.cuh) reads X a second time for the first-mode gradient
//     dG[w,q] = sum_t sum_d X[t,w,d] da[t,q,d]
// because da[t] needs the whole sample's outputs.  A spectral sample (W x D elements, 32 KB on the bench workload) fits
// the shared memory of ONE SM several times over, so here a block keeps a ring of NS whole samples in shared memory
// (TMA bulk copies, cp.async.bulk + mbarrier::complete_tx) and both contractions read the sample from there:
//
//   producer lane      sample j of the block -> stage j % NS as soon as the gradient warps released it (empty[s])
//   NF forward warps   warp f owns the block's samples f, f + NF, ...: window contraction a[q,d] = sum_w X[w,d] G[w,q]
//                      (lanes along d, Q * VEC register sums, G rows broadcast from shared memory), then the per-sample
//                      epilogue of k_spec_fused (norm over the complex axis, second contraction, outputs, residual, ds,
//                      second-mode gradient in registers) and da[q,d] -> shared memory next to the stage (ready[s])
//   NG gradient warps  warp k owns the window rows [8k, 8k+8) of EVERY sample of the block: 8 * Q register sums
//                      acc[i][q] += X[w0+i, d] da[q, d] from the still-resident stage, then releases it (empty[s])
//
// No block-wide barrier in the sample loop; every hand-off is an mbarrier.  da, a and m never touch global memory; the
// kernel writes res[t], [s | 1] (third-mode factors / bias, k_dfc) and its partial sums.  Deterministic: fixed orders,
// one owner per slot, no atomics.  Eligible when the sample fits one warp tile (D <= 32 * VEC, 16-byte rows),
// Q <= TRS_MAXQ, W <= 8 * NG and at least two stages fit (tr_api.cu: spec_single_plan).
//
// Each role is a separate (not inlined) device function: the register allocation of one role does not see the other
// roles' live ranges (as one function body the kernel spilled loop-carried values of every role, and with ~220 KB of
// the SM's memory configured as shared memory the remaining L1 does not hold the spill slots of 13+ warps).
#pragma once
#include "tr_spectral.cuh"
#include "tr_fused.cuh"

#ifndef TRSS_NF
#define TRSS_NF 6                                      // forward warps
#endif
#ifndef TRSS_FFMA2
#define TRSS_FFMA2 0                                   // float window loop: 1 = FFMA2 over element pairs (G values stored twice), 0 = scalar FMAs
#endif
// every wait carries a suspend-time hint: a warp that waits for microseconds is parked by the hardware instead of polling
#define TRSS_WAIT(bar, parity) trf::mbar_wait_hint(bar, parity)
#define TRSS_NG 8                                      // gradient warps (TRS_WT window rows each)
#define TRSS_NT ((TRSS_NF + TRSS_NG + 1) * 32)         // + the producer warp
#define TRSS_MAX_NS 12
#define TRSS_HDR 512                                   // barriers (4 per stage) + loss scratch

template <typename T>
struct SpecSingleArgs {
    const T* X; const T* y; long long N;
    const T* FtT; const T* theta; const T* w;
    SpecGeo g; double nb;
    T* res; T* U; T* yhat;
    double* df1part;       // (grid * TRSS_NF, QT, 32 * VEC): second-mode gradient slots of the forward warps
    double* dgpart;        // (grid, WTN, TRS_WT, QT): first-mode gradient slots of the gradient warps
    double* losspart;      // (grid)
    long long spc;         // samples between two folds of the fp32 running sums into the double slots
    int NS;                // stages
    unsigned stage_bytes;  // W * D * sizeof(T), multiple of 16
    unsigned piece;        // bytes per bulk-copy instruction
    long long* trace;      // debug timeline of block 0 (builds with -DTRSS_TRACE only), else null
    const T* gtab;         // first-mode table in the kernel's layout (W rows of TrssG::STRIDE values), packed by k_spec_pack_g
};

// debug timeline: clock64 stamp of event e of the block's sample j (block 0, TRSS_TRACE_N samples from TRSS_TRACE_J0)
#define TRSS_TRACE_J0 200
#define TRSS_TRACE_N 48
#define TRSS_TRACE_EV 16
#ifdef TRSS_TRACE
#define TRSS_STAMP(j, e)                                                                                     \
    do {                                                                                                     \
        if (blockIdx.x == 0 && lane == 0 && (j) >= TRSS_TRACE_J0 && (j) < TRSS_TRACE_J0 + TRSS_TRACE_N)        \
            trace[((j) - TRSS_TRACE_J0) * TRSS_TRACE_EV + (e)] = clock64();                                  \
    } while (0)
#else
#define TRSS_STAMP(j, e) do {} while (0)
#endif

// ---------------------------------------------------------------------------------------------
// rows of the first-mode table G in shared memory (packed by k_spec_pack_g, copied in by the kernel's prologue) and the
// row update acc[q][v] += x[v] * G[w][q].  Build option -DTRSS_FFMA2=1 (float): every value stored twice, (g, g), so that
// one FFMA2 (fma.rn.f32x2) updates the element pair (v, v+1) of a 16-byte load with no register shuffling — same bits,
// half the FMA issue slots, but three 16-byte row loads instead of two: measured 5 % slower on spec1, not the default.
// ---------------------------------------------------------------------------------------------
template <typename T, int QT> struct TrssG {
    static constexpr int CH = 16 / (int)sizeof(T);
    static constexpr int STRIDE = (QT + CH - 1) / CH * CH;
    template <int VEC>
    static __device__ __forceinline__ void fma_row(T (&acc)[QT][VEC], const T (&x)[VEC], const T* row) {
        T gq[STRIDE];
        VECG<T>::template ld<STRIDE>(row, gq);
#pragma unroll
        for (int q = 0; q < QT; ++q)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[q][v] = tr_fma<T>(x[v], gq[q], acc[q][v]);
    }
};
#if TRSS_FFMA2
template <int QT> struct TrssG<float, QT> {
    static constexpr int STRIDE = (2 * QT + 3) / 4 * 4;
    template <int VEC>
    static __device__ __forceinline__ void fma_row(float (&acc)[QT][VEC], const float (&x)[VEC], const float* row) {
        static_assert(VEC == 4, "float rows are 16-byte chunks");
        float gq[STRIDE];
        VECG<float>::template ld<STRIDE>(row, gq);
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            tr_ffma2(acc[q][0], acc[q][1], x[0], x[1], gq[2 * q], gq[2 * q + 1]);
            tr_ffma2(acc[q][2], acc[q][3], x[2], x[3], gq[2 * q], gq[2 * q + 1]);
        }
    }
};
#endif
// row stride of the table for a run-time channel count (host side and the packing kernel); float values stored twice
template <typename T> __host__ __device__ inline bool trss_g_twice() { return sizeof(T) == 4 && TRSS_FFMA2; }
template <typename T> __host__ __device__ inline int trss_g_stride(int QT) {
    const int ch = 16 / (int)sizeof(T);
    return trss_g_twice<T>() ? (2 * QT + 3) / 4 * 4 : (QT + ch - 1) / ch * ch;
}

// shared-memory loads that keep their program order (the compiler otherwise sinks the batch of row loads of the window
// loop next to their first use, which exposes the full shared-memory latency once per row)
template <typename T, int VEC> struct TrssLd;
template <> struct TrssLd<float, 4> {
    static __device__ __forceinline__ void ld(uint32_t addr, float (&x)[4]) {
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]) : "r"(addr));
    }
};
template <> struct TrssLd<double, 2> {
    static __device__ __forceinline__ void ld(uint32_t addr, double (&x)[2]) {
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x[0]), "=d"(x[1]) : "r"(addr));
    }
};

// sqrt(ss) for ss >= 0 (0 at 0): one flush-to-zero reciprocal square root plus a Newton step on the product for float
// (2 ulp -> fp32 rounding; an order of magnitude fewer instructions than sqrtf), the exact square root for double
__device__ __forceinline__ float trss_norm(float ss) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(ss));
    r = ss > 0.0f ? r : 0.0f;
    r = r * fmaf(-0.5f * ss * r, r, 1.5f);
    return ss * r;
}
__device__ __forceinline__ double trss_norm(double ss) { return sqrt(ss); }
// 1 / m for m > 0: the hardware reciprocal (1 ulp) plus one Newton step for float — an IEEE-rounded reciprocal
// (__frcp_rn) is a ~30-instruction sequence and was a quarter of the epilogue's stall samples —, an exact division for double
__device__ __forceinline__ float spec_recip(float m) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(m));
    return fmaf(r, fmaf(-m, r, 1.0f), r);
}
__device__ __forceinline__ double spec_recip(double m) { return 1.0 / m; }

struct SpecSingleLayout { size_t tab, sF1, sDA, stage, total; };

// shared-memory carve-up, the same on host (size) and device (pointers)
template <typename T>
__host__ __device__ inline SpecSingleLayout spec_single_layout(const SpecGeo& g, int QT, int VEC, int NS, size_t stage_bytes) {
    SpecSingleLayout L;
    L.tab = TRSS_HDR;
    const size_t tabB = ((size_t)g.W * trss_g_stride<T>(QT) + (size_t)g.NO * QT + g.NO) * sizeof(T);
    L.sF1 = L.tab + (tabB + 15) / 16 * 16;                                   // (QT, 32 * VEC): second-mode factor rows
    L.sDA = L.sF1 + (size_t)QT * 32 * VEC * sizeof(T);
    L.stage = (L.sDA + (size_t)NS * QT * 32 * VEC * sizeof(T) + 127) / 128 * 128;
    L.total = L.stage + (size_t)NS * stage_bytes;
    return L;
}

// what every role needs: views into the block's dynamic shared memory, derived from the launch arguments in every role
// function itself (pointers handed over through memory would lose their address space: generic loads instead of LDS)
template <typename T>
struct TrssCtx {
    uint64_t* full; uint64_t* ready; uint64_t* empty; double* sloss;
    T* sG; T* sF2; T* sB; T* sF1; T* sDA; unsigned char* stages;
    long long nj;          // samples of this block: t = blockIdx.x + j * gridDim.x
};

extern __shared__ __align__(128) unsigned char trss_smem[];

template <typename T, int QT, int VEC>
__device__ __forceinline__ TrssCtx<T> trss_ctx(const SpecGeo& g, int NS, unsigned stage_bytes, long long N) {
    const SpecSingleLayout L = spec_single_layout<T>(g, QT, VEC, NS, stage_bytes);
    TrssCtx<T> c;
    // full[(round & 1) * MAX_NS + s]: the bytes of the stage's sample of that round landed.  Two barriers per stage: a forward
    // warp may start waiting for sample j while sample j - NS (same stage, previous round) is still on its way; with one
    // barrier its parity wait would alias to the round before that.  With two, the barrier's previous use is sample
    // j - 2 NS, whose landing precedes the issue of the load the warp's previous sample came with.
    c.full = reinterpret_cast<uint64_t*>(trss_smem);
    c.ready = c.full + 2 * TRSS_MAX_NS;                                      // [NS] da of the sample is in shared memory
    c.empty = c.ready + TRSS_MAX_NS;                                         // [NS] gradient warps are done with the stage
    c.sloss = reinterpret_cast<double*>(c.empty + TRSS_MAX_NS);              // [NF]
    c.sG = reinterpret_cast<T*>(trss_smem + L.tab);                          // (W, TrssG::STRIDE)
    c.sF2 = c.sG + (size_t)g.W * TrssG<T, QT>::STRIDE;   // (NO, QT): w_r Fn2[n,r] | Fc2[n,r]
    c.sB = c.sF2 + (size_t)g.NO * QT;                                        // (NO): nb * bias
    c.sF1 = reinterpret_cast<T*>(trss_smem + L.sF1);                         // (QT, TILE): F1[d, r] of component r (0 beyond D / RT)
    c.sDA = reinterpret_cast<T*>(trss_smem + L.sDA);                         // (NS, QT, TILE): a, then da of the stage's sample
    c.stages = trss_smem + L.stage;
    const long long grid = gridDim.x;
    // the block's samples: t = blockIdx.x + j * grid, j < nj (strided: all SMs sweep one moving window of X)
    c.nj = (long long)blockIdx.x < N ? (N - blockIdx.x + grid - 1) / grid : 0;
    return c;
}

// ---------------- producer (one lane) ----------------
template <typename T, int QT, int VEC>
__device__ __noinline__ void trss_producer(const SpecSingleArgs<T>* __restrict__ ap) {
    const int NS = ap->NS;
    const unsigned stage_bytes = ap->stage_bytes, piece = ap->piece;
    const TrssCtx<T> c = trss_ctx<T, QT, VEC>(ap->g, NS, stage_bytes, ap->N);
    const size_t WD = (size_t)ap->g.W * ap->g.D;
    const T* X = ap->X;
    const long long grid = gridDim.x;
#ifdef TRSS_TRACE
    long long* trace = ap->trace;
    const int lane = 0;
#endif
    int s = 0; unsigned round = 0;
    for (long long j = 0; j < c.nj; ++j) {
        if (round > 0) TRSS_WAIT(&c.empty[s], (round - 1) & 1);
        TRSS_STAMP(j, 0);
        uint64_t* fb = &c.full[(round & 1) * TRSS_MAX_NS + s];
        trf::mbar_arrive_expect_tx(fb, stage_bytes);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(X + (size_t)(blockIdx.x + j * grid) * WD);
        unsigned char* dst = c.stages + (size_t)s * stage_bytes;
        for (unsigned off = 0; off < stage_bytes; off += piece) {
            const unsigned nb = stage_bytes - off < piece ? stage_bytes - off : piece;
            trf::bulk_g2s(dst + off, src + off, nb, fb);
        }
        if (++s == NS) { s = 0; ++round; }
    }
}

// ---------------- forward + per-sample epilogue (one warp per sample) ----------------
template <typename T, int QT, int VEC>
__device__ __noinline__ void trss_forward(const SpecSingleArgs<T>* __restrict__ ap) {
    using GR = TrssG<T, QT>;
    constexpr int TILE = 32 * VEC;
    const SpecGeo g = ap->g;
    const int NS = ap->NS;
    const unsigned stage_bytes = ap->stage_bytes;
    const TrssCtx<T> c = trss_ctx<T, QT, VEC>(g, NS, stage_bytes, ap->N);
    const T* __restrict__ yp = ap->y;
    T* __restrict__ resp = ap->res; T* __restrict__ Up = ap->U; T* __restrict__ yhatp = ap->yhat;
    const long long grid = gridDim.x;
#ifdef TRSS_TRACE
    long long* trace = ap->trace;
#endif
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int d0 = lane * VEC;
    const bool act = d0 < g.D;
    // a warp waits for the landing of sample j + nfa while the barrier of that stage may still be one phase behind
    // only if the stage's previous sample is not younger than j: at most NS forward warps take part
    const int nfa = NS < TRSS_NF ? NS : TRSS_NF;
    const T* f1p = c.sF1 + d0;                                               // F1[d0.., r] at f1p + r * TILE
    T accF[QT][VEC];                                                         // sum_t ds[t,r] m[t,r,d] since the last fold
#pragma unroll
    for (int r = 0; r < QT; ++r)
#pragma unroll
        for (int v = 0; v < VEC; ++v) accF[r][v] = (T)0;
    double* slot = ap->df1part + ((size_t)blockIdx.x * TRSS_NF + wid) * QT * TILE + d0;
#pragma unroll
    for (int r = 0; r < QT; ++r)
#pragma unroll
        for (int v = 0; v < VEC; ++v) slot[(size_t)r * TILE + v] = 0.0;
    double loss = 0.0;
    T lossp = (T)0;                                                          // squared residuals of up to 16 samples of this lane
    const int spc = (int)(ap->spc < (1LL << 30) ? ap->spc : (1LL << 30));
    int left = spc;
    int s = wid; unsigned round = 0;
    long long t = blockIdx.x + (long long)wid * grid;
    const long long tstep = (long long)nfa * grid;
    for (long long j = wid < nfa ? wid : c.nj; j < c.nj; j += nfa, t += tstep) {
        T yv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) yv[k] = (lane + 32 * k < g.NO) ? __ldg(yp + t * g.NO + lane + 32 * k) : (T)0;
        T acc[QT][VEC];
#pragma unroll
        for (int q = 0; q < QT; ++q)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[q][v] = (T)0;
        TRSS_STAMP(j, 1);
        TRSS_WAIT(&c.full[(round & 1) * TRSS_MAX_NS + s], (round >> 1) & 1);
        TRSS_STAMP(j, 2);
        uint32_t xa = trf::smem_u32(c.stages + (size_t)s * stage_bytes) + (uint32_t)((act ? d0 : 0) * sizeof(T));
        const uint32_t rowb = (uint32_t)(g.D * sizeof(T));
        const T* gr = c.sG;
        constexpr int UW = 8;
        int w = 0;
        for (; w + UW <= g.W; w += UW) {
            T x[UW][VEC];
#pragma unroll
            for (int u = 0; u < UW; ++u) TrssLd<T, VEC>::ld(xa + u * rowb, x[u]);
#pragma unroll
            for (int u = 0; u < UW; ++u) GR::template fma_row<VEC>(acc, x[u], gr + u * GR::STRIDE);
            xa += UW * rowb;
            gr += UW * GR::STRIDE;
        }
        for (; w < g.W; ++w) {
            T x[VEC];
            TrssLd<T, VEC>::ld(xa, x);
            GR::template fma_row<VEC>(acc, x, gr);
            xa += rowb;
            gr += GR::STRIDE;
        }
        if (!act) {
#pragma unroll
            for (int q = 0; q < QT; ++q)
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[q][v] = (T)0;
        }
        TRSS_STAMP(j, 3);
        // Channel q belongs to component r when  r < Rn: q == r  (normal)  |  Rn <= r < RT: qb(r) <= q < qb(r) + CC, qb(r) = Rn + (r - Rn) CC
        // — all warp-uniform.  The epilogue runs over STATIC (r, q) pairs guarded by uniform branches that contain arithmetic
        // (and stores) only: every shared-memory load is issued in a batch outside the branches, so no ~100-cycle load latency
        // is serialised per component in the warp that is the critical path of its sample.
        // m[r] = a itself (normal component) or the norm over the component's complex channels is parked in row r of the stage's
        // da slot (free until this warp releases it; a lane reads back only what it wrote) and re-read in batches: next to
        // the window sums and the second-mode gradient sums it would not fit the registers
        T* das = c.sDA + (size_t)s * QT * TILE + d0;
#pragma unroll
        for (int r = 0; r < QT; ++r) {
            T m[VEC];
            if (r < g.Rn) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) m[v] = acc[r][v];
            } else if (r < g.RT) {
                const int qb = g.Rn + (r - g.Rn) * g.CC;
                T ss[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) ss[v] = (T)0;
#pragma unroll
                for (int q = r; q < QT; ++q) {                               // qb(r) >= r
                    if (q >= qb && q < qb + g.CC) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) ss[v] = tr_fma<T>(acc[q][v], acc[q][v], ss[v]);
                    }
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) m[v] = trss_norm(ss[v]);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) m[v] = (T)0;
            }
            SpecSm<T, VEC>::st(das + (size_t)r * TILE, m);
        }
        TRSS_STAMP(j, 9);
        // second contraction: s[r] = sum_d m[d,r] F1[d,r]  (lane partial, then an all-reduce over the warp)
        T sr[QT];
#pragma unroll
        for (int r = 0; r < QT; ++r) {
            T p = (T)0, f1[VEC], m[VEC];
            SpecSm<T, VEC>::ld(f1p + (size_t)r * TILE, f1);
            SpecSm<T, VEC>::ld(das + (size_t)r * TILE, m);
#pragma unroll
            for (int v = 0; v < VEC; ++v) p = tr_fma<T>(m[v], f1[v], p);
            sr[r] = p;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int r = 0; r < QT; ++r) sr[r] += __shfl_xor_sync(TR_FULL, sr[r], off);
        TRSS_STAMP(j, 10);
        if (lane <= g.RT) {
            T uv = (T)1;
#pragma unroll
            for (int r = 0; r < QT; ++r) if (r == lane && r < g.RT) uv = sr[r];
            Up[t * (g.RT + 1) + lane] = uv;
        }
        // outputs and residuals: lanes along n; ds[r] = sum_n res[n] F2[n,r]
        T ds[QT];
#pragma unroll
        for (int r = 0; r < QT; ++r) ds[r] = (T)0;
        T l2 = (T)0;                                                         // sum of this lane's squared residuals (at most 4)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int n = lane + 32 * k;
            if (n < g.NO) {
                const T* f2 = c.sF2 + (size_t)n * QT;
                T yh = c.sB[n];
#pragma unroll
                for (int r = 0; r < QT; ++r) yh = tr_fma<T>(sr[r], f2[r], yh);
                const T rr = yh - yv[k];
                if (yhatp) yhatp[t * g.NO + n] = yh;
                resp[t * g.NO + n] = rr;
                l2 = tr_fma<T>(rr, rr, l2);
#pragma unroll
                for (int r = 0; r < QT; ++r) ds[r] = tr_fma<T>(rr, f2[r], ds[r]);
            }
        }
        lossp += l2;
        if ((left & 15) == 0) { loss += (double)lossp; lossp = (T)0; }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int r = 0; r < QT; ++r) ds[r] += __shfl_xor_sync(TR_FULL, ds[r], off);
        TRSS_STAMP(j, 11);
        // second-mode gradient (registers) and da -> shared memory next to the stage.
        // da of a normal channel r = ds[r] F1[d,r]; of the channels (r, c) of a spectral component = ds[r] F1[d,r] / m[d,r] * a
        // (0 where the norm is 0: torch.norm's subgradient)
        {
            T m[QT][VEC];                                                    // all rows first: da overwrites them below
#pragma unroll
            for (int r = 0; r < QT; ++r) SpecSm<T, VEC>::ld(das + (size_t)r * TILE, m[r]);
#pragma unroll
            for (int r = 0; r < QT; ++r) {
                T kf[VEC], f1[VEC];
                SpecSm<T, VEC>::ld(f1p + (size_t)r * TILE, f1);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    accF[r][v] = tr_fma<T>(ds[r], m[r][v], accF[r][v]);
                    kf[v] = ds[r] * f1[v];
                }
                if (r < g.Rn) {
                    SpecSm<T, VEC>::st(das + (size_t)r * TILE, kf);
                } else if (r < g.RT) {
                    const int qb = g.Rn + (r - g.Rn) * g.CC;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) kf[v] = m[r][v] > (T)0 ? kf[v] * spec_recip(m[r][v]) : (T)0;
#pragma unroll
                    for (int q = r; q < QT; ++q) {
                        if (q >= qb && q < qb + g.CC) {
                            T out[VEC];
#pragma unroll
                            for (int v = 0; v < VEC; ++v) out[v] = acc[q][v] * kf[v];
                            SpecSm<T, VEC>::st(das + (size_t)q * TILE, out);
                        }
                    }
                }
            }
        }
        __syncwarp();                                                        // the lanes' stores, then one release for the warp
        if (lane == 0) trf::mbar_arrive(&c.ready[s]);
        TRSS_STAMP(j, 4);
        if (--left == 0) {
#pragma unroll
            for (int r = 0; r < QT; ++r)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    slot[(size_t)r * TILE + v] += (double)accF[r][v];
                    accF[r][v] = (T)0;
                }
            left = spc;
        }
        s += nfa;
        if (s >= NS) { s -= NS; ++round; }
    }
#pragma unroll
    for (int r = 0; r < QT; ++r)
#pragma unroll
        for (int v = 0; v < VEC; ++v) slot[(size_t)r * TILE + v] += (double)accF[r][v];
    loss = warp_sum(loss + (double)lossp);
    if (lane == 0) c.sloss[wid] = loss;
}

// ---------------- first-mode gradient (warp = 8 window rows of every sample) ----------------
template <typename T, int QT, int VEC>
__device__ __noinline__ void trss_gradient(const SpecSingleArgs<T>* __restrict__ ap) {
    constexpr int TILE = 32 * VEC;
    const int W = ap->g.W, D = ap->g.D;
    const int NS = ap->NS;
    const unsigned stage_bytes = ap->stage_bytes;
    const TrssCtx<T> c = trss_ctx<T, QT, VEC>(ap->g, NS, stage_bytes, ap->N);
#ifdef TRSS_TRACE
    long long* trace = ap->trace;
#endif
    const int lane = threadIdx.x & 31;
    const int wt = (threadIdx.x >> 5) - TRSS_NF;
    const int d0 = lane * VEC;
    const int w0 = wt * TRS_WT;
    const bool rows = w0 < W;
    // rows of X this lane works on: the warp's tile clipped to the window; none for lanes beyond D
    const int nrow = d0 >= D ? 0 : (W - w0 < TRS_WT ? W - w0 : TRS_WT);
    T acc[TRS_WT][QT];
#pragma unroll
    for (int i = 0; i < TRS_WT; ++i)
#pragma unroll
        for (int q = 0; q < QT; ++q) acc[i][q] = (T)0;
    const int WTN = (W + TRS_WT - 1) / TRS_WT;
    double* slot = ap->dgpart + ((size_t)blockIdx.x * WTN + wt) * TRS_WT * QT;
    if (rows) {
        if (lane < TRS_WT * QT) slot[lane] = 0.0;
        if (lane + 32 < TRS_WT * QT) slot[lane + 32] = 0.0;
    }
    const unsigned xoff = (unsigned)(((size_t)w0 * D + d0) * sizeof(T));         // the lane's first element within a stage
    const int nji = (int)c.nj;
    const int spc = (int)(ap->spc < (long long)nji ? ap->spc : (long long)(nji > 0 ? nji : 1));
    int s = 0; unsigned round = 0;
    for (int j0 = 0; j0 < nji; j0 += spc) {
        const int j1 = j0 + spc < nji ? j0 + spc : nji;
        for (int j = j0; j < j1; ++j) {
            if (wt == 0) TRSS_STAMP(j, 5);
            TRSS_WAIT(&c.ready[s], round & 1);
            if (wt == 0) TRSS_STAMP(j, 6);
            const T* xs = reinterpret_cast<const T*>(c.stages + (size_t)s * stage_bytes + xoff);
            const T* das = c.sDA + (size_t)s * QT * TILE + d0;
            if (sizeof(T) == 4 && nrow == TRS_WT) {
                // full tile (float: the registers allow it): all loads first, their latencies overlap, then 8 * QT * VEC FMAs
                T da[QT][VEC], x[TRS_WT][VEC];
#pragma unroll
                for (int q = 0; q < QT; ++q) SpecSm<T, VEC>::ld(das + (size_t)q * TILE, da[q]);
#pragma unroll
                for (int i = 0; i < TRS_WT; ++i) SpecSm<T, VEC>::ld(xs + (size_t)i * D, x[i]);
#pragma unroll
                for (int i = 0; i < TRS_WT; ++i)
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
#pragma unroll
                        for (int q = 0; q < QT; ++q) acc[i][q] = tr_fma<T>(x[i][v], da[q][v], acc[i][q]);
            } else if (nrow > 0) {
                T da[QT][VEC];
#pragma unroll
                for (int q = 0; q < QT; ++q) SpecSm<T, VEC>::ld(das + (size_t)q * TILE, da[q]);
#pragma unroll
                for (int i = 0; i < TRS_WT; ++i) {
                    if (i < nrow) {
                        T x[VEC];
                        SpecSm<T, VEC>::ld(xs + (size_t)i * D, x);
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
#pragma unroll
                            for (int q = 0; q < QT; ++q) acc[i][q] = tr_fma<T>(x[v], da[q][v], acc[i][q]);
                    }
                }
            }
            __syncwarp();                                                    // every lane has read the stage
            if (lane == 0) trf::mbar_arrive(&c.empty[s]);
            if (wt == 0) TRSS_STAMP(j, 7);
            if (wt == TRSS_NG - 1) TRSS_STAMP(j, 8);
            if (++s == NS) { s = 0; ++round; }
        }
        if (rows) {
            // fold across lanes and add to the warp's slot (double); bounds every fp32 running sum
#pragma unroll
            for (int i = 0; i < TRS_WT; ++i)
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    double sv = (double)acc[i][q];
                    sv = warp_sum(sv);
                    if (lane == 0) slot[i * QT + q] += sv;
                    acc[i][q] = (T)0;
                }
        }
    }
}

template <typename T, int QT, int VEC>
__global__ void __launch_bounds__(TRSS_NT, 1) k_spec_single(const __grid_constant__ SpecSingleArgs<T> a) {
    const SpecGeo& g = a.g;
    constexpr int TILE = 32 * VEC;
    const int NS = a.NS;
    const TrssCtx<T> c = trss_ctx<T, QT, VEC>(g, NS, a.stage_bytes, a.N);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            trf::mbar_init(&c.full[s], 1);
            trf::mbar_init(&c.full[TRSS_MAX_NS + s], 1);
            trf::mbar_init(&c.ready[s], 1);
            trf::mbar_init(&c.empty[s], TRSS_NG);
        }
        trf::fence_mbar_init();
    }
    for (int i = threadIdx.x; i < g.W * TrssG<T, QT>::STRIDE; i += TRSS_NT) c.sG[i] = a.gtab[i];
    for (int i = threadIdx.x; i < g.NO * QT; i += TRSS_NT) {
        const int n = i / QT, r = i % QT;
        T v = (T)0;
        if (r < g.Rn) v = a.w[r] * a.FtT[g.off[2] + n * g.Rn + r];
        else if (r < g.RT) v = a.FtT[g.off[5] + n * g.Rs + (r - g.Rn)];
        c.sF2[i] = v;
    }
    for (int i = threadIdx.x; i < g.NO; i += TRSS_NT) c.sB[i] = (T)(a.nb * (double)a.theta[g.off[6] + i]);
    for (int i = threadIdx.x; i < QT * TILE; i += TRSS_NT) {
        const int r = i / TILE, d = i % TILE;
        T val = (T)0;
        if (d < g.D && r < g.Rn) val = a.FtT[g.off[1] + d * g.Rn + r];
        else if (d < g.D && r < g.RT) val = a.FtT[g.off[4] + d * g.Rs + (r - g.Rn)];
        c.sF1[i] = val;
    }
    __syncthreads();

    const int wid = threadIdx.x >> 5;
    if (wid == TRSS_NF + TRSS_NG) {
        if ((threadIdx.x & 31) == 0) trss_producer<T, QT, VEC>(&a);
    } else if (wid < TRSS_NF) {
        trss_forward<T, QT, VEC>(&a);
    } else {
        trss_gradient<T, QT, VEC>(&a);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < TRSS_NF; ++i) tot += c.sloss[i];
        a.losspart[blockIdx.x] = tot;
    }
}
