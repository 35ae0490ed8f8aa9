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
// Q <= TRS_MAXQ, W <= 8 * NG and at least three stages fit (tr_api.cu: spec_single_plan).
#pragma once
#include "tr_spectral.cuh"
#include "tr_fused.cuh"

#define TRSS_NF 4                                      // forward warps
#define TRSS_NG 8                                      // gradient warps (TRS_WT window rows each)
#define TRSS_NT ((TRSS_NF + TRSS_NG + 1) * 32)         // + the producer warp
#define TRSS_MAX_NS 8
#define TRSS_HDR 384                                   // barriers + loss scratch

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
};

// debug timeline: clock64 stamp of event e of the block's sample j (block 0, TRSS_TRACE_N samples from TRSS_TRACE_J0)
#define TRSS_TRACE_J0 200
#define TRSS_TRACE_N 48
#define TRSS_TRACE_EV 16
#ifdef TRSS_TRACE
#define TRSS_STAMP(j, e)                                                                                     \
    do {                                                                                                     \
        if (blockIdx.x == 0 && lane == 0 && (j) >= TRSS_TRACE_J0 && (j) < TRSS_TRACE_J0 + TRSS_TRACE_N)        \
            a.trace[((j) - TRSS_TRACE_J0) * TRSS_TRACE_EV + (e)] = clock64();                                \
    } while (0)
#else
#define TRSS_STAMP(j, e) do {} while (0)
#endif

struct SpecSingleLayout { size_t tab, sF1, sDA, stage, total; };

// shared-memory carve-up, the same on host (size) and device (pointers)
template <typename T>
__host__ __device__ inline SpecSingleLayout spec_single_layout(const SpecGeo& g, int QT, int VEC, int NS, size_t stage_bytes) {
    const int CH = 16 / (int)sizeof(T);
    const int QP = (QT + CH - 1) / CH * CH;
    SpecSingleLayout L;
    L.tab = TRSS_HDR;
    const size_t tabB = ((size_t)g.W * QP + (size_t)g.NO * QT + g.NO) * sizeof(T);
    L.sF1 = L.tab + (tabB + 15) / 16 * 16;                                   // (QT, 32 * VEC): second-mode factor rows
    L.sDA = L.sF1 + (size_t)QT * 32 * VEC * sizeof(T);
    L.stage = (L.sDA + (size_t)NS * QT * 32 * VEC * sizeof(T) + 127) / 128 * 128;
    L.total = L.stage + (size_t)NS * stage_bytes;
    return L;
}

template <typename T, int QT, int VEC>
__global__ void __launch_bounds__(TRSS_NT, 1) k_spec_single(const SpecSingleArgs<T> a) {
    extern __shared__ __align__(128) unsigned char trss_smem[];
    unsigned char* const tr_smem = trss_smem;
    const SpecGeo& g = a.g;
    constexpr int QP = (QT + VECG<T>::v - 1) / VECG<T>::v * VECG<T>::v;
    constexpr int TILE = 32 * VEC;
    const int NS = a.NS;
    const SpecSingleLayout L = spec_single_layout<T>(g, QT, VEC, NS, a.stage_bytes);
    // full[(round & 1) * MAX_NS + s]: the bytes of the stage's sample of that round landed.  Two barriers per stage: a forward
    // warp may start waiting for sample j while sample j - NS (same stage, previous round) is still on its way; with one
    // barrier its parity wait would alias to the round before that.  With two, the barrier's previous use is sample
    // j - 2 NS, whose landing precedes the issue of the load the warp's previous sample came with.
    uint64_t* full = reinterpret_cast<uint64_t*>(tr_smem);
    uint64_t* ready = full + 2 * TRSS_MAX_NS;                                  // [NS] da of the sample is in shared memory
    uint64_t* empty = ready + TRSS_MAX_NS;                                   // [NS] gradient warps are done with the stage
    double* sloss = reinterpret_cast<double*>(empty + TRSS_MAX_NS);          // [NF]
    T* sG = reinterpret_cast<T*>(tr_smem + L.tab);                           // (W, QP)
    T* sF2 = sG + (size_t)g.W * QP;                                          // (NO, QT): w_r Fn2[n,r] | Fc2[n,r]
    T* sB = sF2 + (size_t)g.NO * QT;                                         // (NO): nb * bias
    T* sF1 = reinterpret_cast<T*>(tr_smem + L.sF1);                          // (QT, TILE): F1[d, r] of component r (0 beyond D / RT)
    T* sDA = reinterpret_cast<T*>(tr_smem + L.sDA);                          // (NS, QT, TILE): a, then da of the stage's sample
    unsigned char* stages = tr_smem + L.stage;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            trf::mbar_init(&full[s], 1);
            trf::mbar_init(&full[TRSS_MAX_NS + s], 1);
            trf::mbar_init(&ready[s], 1);
            trf::mbar_init(&empty[s], TRSS_NG);
        }
        trf::fence_mbar_init();
    }
    for (int i = threadIdx.x; i < g.W * QP; i += TRSS_NT) {
        const int w = i / QP, q = i % QP;
        sG[i] = q < g.Q ? spec_G(a.FtT, g, w, q) : (T)0;
    }
    for (int i = threadIdx.x; i < g.NO * QT; i += TRSS_NT) {
        const int n = i / QT, r = i % QT;
        T v = (T)0;
        if (r < g.Rn) v = a.w[r] * a.FtT[g.off[2] + n * g.Rn + r];
        else if (r < g.RT) v = a.FtT[g.off[5] + n * g.Rs + (r - g.Rn)];
        sF2[i] = v;
    }
    for (int i = threadIdx.x; i < g.NO; i += TRSS_NT) sB[i] = (T)(a.nb * (double)a.theta[g.off[6] + i]);
    for (int i = threadIdx.x; i < QT * TILE; i += TRSS_NT) {
        const int r = i / TILE, d = i % TILE;
        T val = (T)0;
        if (d < g.D && r < g.Rn) val = a.FtT[g.off[1] + d * g.Rn + r];
        else if (d < g.D && r < g.RT) val = a.FtT[g.off[4] + d * g.Rs + (r - g.Rn)];
        sF1[i] = val;
    }
    __syncthreads();

    const int wid = threadIdx.x >> 5;
    // every role derives its lane-dependent values from an opaque copy of the lane index, so that nothing is computed
    // (and kept alive, i.e. spilled) across the other roles' code
#define TRSS_ROLE_LOCALS()                                        \
    int lane = threadIdx.x & 31;                                  \
    asm volatile("" : "+r"(lane));                                \
    const int d0 = lane * VEC;                                    \
    const bool act = d0 < g.D;
    const size_t WD = (size_t)g.W * g.D;
    const long long grid = gridDim.x;
    // the block's samples: t = blockIdx.x + j * grid, j < nj (strided: all SMs sweep one moving window of X)
    const long long nj = (long long)blockIdx.x < a.N ? (a.N - blockIdx.x + grid - 1) / grid : 0;

    if (wid == TRSS_NF + TRSS_NG) {
        // ---------------- producer ----------------
        int lane = threadIdx.x & 31;
        if (lane == 0) {
            int s = 0; unsigned round = 0;
            for (long long j = 0; j < nj; ++j) {
                if (round > 0) trf::mbar_wait(&empty[s], (round - 1) & 1);
                TRSS_STAMP(j, 0);
                uint64_t* fb = &full[(round & 1) * TRSS_MAX_NS + s];
                trf::mbar_arrive_expect_tx(fb, a.stage_bytes);
                const unsigned char* src = reinterpret_cast<const unsigned char*>(a.X + (size_t)(blockIdx.x + j * grid) * WD);
                unsigned char* dst = stages + (size_t)s * a.stage_bytes;
                for (unsigned off = 0; off < a.stage_bytes; off += a.piece) {
                    const unsigned nb = a.stage_bytes - off < a.piece ? a.stage_bytes - off : a.piece;
                    trf::bulk_g2s(dst + off, src + off, nb, fb);
                }
                if (++s == NS) { s = 0; ++round; }
            }
        }
    } else if (wid < TRSS_NF) {
        // ---------------- forward + per-sample epilogue ----------------
        TRSS_ROLE_LOCALS()
        // a warp waits for the landing of sample j + nfa while the barrier of that stage may still be one phase behind
        // only if the stage's previous sample is not younger than j: at most NS forward warps take part
        const int nfa = NS < TRSS_NF ? NS : TRSS_NF;
        const T* f1p = sF1 + d0;                                             // F1[d0.., r] at f1p + r * TILE
        T accF[VEC][QT];                                                     // sum_t ds[t,r] m[t,r,d] since the last fold
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int r = 0; r < QT; ++r) accF[v][r] = (T)0;
        double* slot = a.df1part + ((size_t)blockIdx.x * TRSS_NF + wid) * QT * TILE;
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int r = 0; r < QT; ++r) slot[(size_t)r * TILE + d0 + v] = 0.0;
        double loss = 0.0;
        long long left = a.spc;
        int s = wid; unsigned round = 0;
        for (long long j = wid < nfa ? wid : nj; j < nj; j += nfa) {
            const long long t = blockIdx.x + j * grid;
            T yv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) yv[k] = (lane + 32 * k < g.NO) ? __ldg(a.y + t * g.NO + lane + 32 * k) : (T)0;
            T acc[VEC][QT];
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int q = 0; q < QT; ++q) acc[v][q] = (T)0;
            TRSS_STAMP(j, 1);
            trf::mbar_wait(&full[(round & 1) * TRSS_MAX_NS + s], (round >> 1) & 1);
            TRSS_STAMP(j, 2);
            const T* xs = reinterpret_cast<const T*>(stages + (size_t)s * a.stage_bytes) + (act ? d0 : 0);
            constexpr int UW = 8;
            int w = 0;
            for (; w + UW <= g.W; w += UW) {
                T x[UW][VEC];
#pragma unroll
                for (int u = 0; u < UW; ++u) SpecSm<T, VEC>::ld(xs + (size_t)(w + u) * g.D, x[u]);
#pragma unroll
                for (int u = 0; u < UW; ++u) {
                    T gq[QP];
                    VECG<T>::template ld<QP>(sG + (size_t)(w + u) * QP, gq);
#pragma unroll
                    for (int q = 0; q < QT; ++q)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[v][q] = tr_fma<T>(x[u][v], gq[q], acc[v][q]);
                }
            }
            for (; w < g.W; ++w) {
                T x[VEC], gq[QP];
                SpecSm<T, VEC>::ld(xs + (size_t)w * g.D, x);
                VECG<T>::template ld<QP>(sG + (size_t)w * QP, gq);
#pragma unroll
                for (int q = 0; q < QT; ++q)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[v][q] = tr_fma<T>(x[v], gq[q], acc[v][q]);
            }
            if (!act) {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
#pragma unroll
                    for (int q = 0; q < QT; ++q) acc[v][q] = (T)0;
            }
            TRSS_STAMP(j, 3);
            // window sums -> the stage's da slot (free until this warp releases it; a lane reads back only what it wrote), so
            // that the channels of a component can be addressed at run time; da overwrites them in place below
            T* das = sDA + (size_t)s * QT * TILE + d0;
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                T out[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) out[v] = acc[v][q];
                SpecSm<T, VEC>::st(das + (size_t)q * TILE, out);
            }
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
                    const T* ap = das + (size_t)(g.Rn + (r - g.Rn) * g.CC) * TILE;
                    for (int c = 0; c < g.CC; ++c) {
                        T av[VEC];
                        SpecSm<T, VEC>::ld(ap + (size_t)c * TILE, av);
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
            T sr[QT];
#pragma unroll
            for (int r = 0; r < QT; ++r) {
                T p = (T)0, f1[VEC];
                SpecSm<T, VEC>::ld(f1p + (size_t)r * TILE, f1);
#pragma unroll
                for (int v = 0; v < VEC; ++v) p = tr_fma<T>(m[v][r], f1[v], p);
                sr[r] = p;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
                for (int r = 0; r < QT; ++r) sr[r] += __shfl_xor_sync(TR_FULL, sr[r], off);
            if (lane <= g.RT) {
                T uv = (T)1;
#pragma unroll
                for (int r = 0; r < QT; ++r) if (r == lane && r < g.RT) uv = sr[r];
                a.U[t * (g.RT + 1) + lane] = uv;
            }
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
                    for (int r = 0; r < QT; ++r) yh = tr_fma<T>(sr[r], f2[r], yh);
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
            // second-mode gradient (registers) and da -> shared memory next to the stage
#pragma unroll
            for (int r = 0; r < QT; ++r) {
                T kf[VEC], f1[VEC];
                SpecSm<T, VEC>::ld(f1p + (size_t)r * TILE, f1);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    accF[v][r] = tr_fma<T>(ds[r], m[v][r], accF[v][r]);
                    kf[v] = ds[r] * f1[v] * rinv[v][r];
                }
                if (r < g.Rn) {
                    SpecSm<T, VEC>::st(das + (size_t)r * TILE, kf);
                } else if (r < g.RT) {
                    const int qb = g.Rn + (r - g.Rn) * g.CC;
                    for (int c = 0; c < g.CC; ++c) {
                        T av[VEC];
                        SpecSm<T, VEC>::ld(das + (size_t)(qb + c) * TILE, av);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) av[v] *= kf[v];
                        SpecSm<T, VEC>::st(das + (size_t)(qb + c) * TILE, av);
                    }
                }
            }
            __syncwarp();                                                    // the lanes' stores, then one release for the warp
            if (lane == 0) trf::mbar_arrive(&ready[s]);
            TRSS_STAMP(j, 4);
            if (--left == 0) {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
#pragma unroll
                    for (int r = 0; r < QT; ++r) {
                        slot[(size_t)r * TILE + d0 + v] += (double)accF[v][r];
                        accF[v][r] = (T)0;
                    }
                left = a.spc;
            }
            s += nfa;
            if (s >= NS) { s -= NS; ++round; }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int r = 0; r < QT; ++r) slot[(size_t)r * TILE + d0 + v] += (double)accF[v][r];
        loss = warp_sum(loss);
        if (lane == 0) sloss[wid] = loss;
    } else {
        // ---------------- first-mode gradient ----------------
        TRSS_ROLE_LOCALS()
        const int wt = wid - TRSS_NF;
        const int w0 = wt * TRS_WT;
        const bool rows = w0 < g.W;
        // rows of X this lane works on: the warp's tile clipped to the window; none for lanes beyond D
        const int nrow = !act ? 0 : (g.W - w0 < TRS_WT ? g.W - w0 : TRS_WT);
        T acc[TRS_WT][QT];
#pragma unroll
        for (int i = 0; i < TRS_WT; ++i)
#pragma unroll
            for (int q = 0; q < QT; ++q) acc[i][q] = (T)0;
        const int WTN = (g.W + TRS_WT - 1) / TRS_WT;
        double* slot = a.dgpart + ((size_t)blockIdx.x * WTN + wt) * TRS_WT * QT;
        if (rows) {
            if (lane < TRS_WT * QT) slot[lane] = 0.0;
            if (lane + 32 < TRS_WT * QT) slot[lane + 32] = 0.0;
        }
        const unsigned xoff = (unsigned)(((size_t)w0 * g.D + d0) * sizeof(T));   // the lane's first element within a stage
        const int nji = (int)nj;
        const int spc = (int)(a.spc < (long long)nji ? a.spc : (long long)(nji > 0 ? nji : 1));
        int s = 0; unsigned round = 0;
        for (int j0 = 0; j0 < nji; j0 += spc) {
            const int j1 = j0 + spc < nji ? j0 + spc : nji;
            for (int j = j0; j < j1; ++j) {
                if (wt == 0) TRSS_STAMP(j, 5);
                trf::mbar_wait(&ready[s], round & 1);
                if (wt == 0) TRSS_STAMP(j, 6);
                if (nrow > 0) {
                    const T* xs = reinterpret_cast<const T*>(stages + (size_t)s * a.stage_bytes + xoff);
                    const T* das = sDA + (size_t)s * QT * TILE + d0;
                    T da[QT][VEC];
#pragma unroll
                    for (int q = 0; q < QT; ++q) SpecSm<T, VEC>::ld(das + (size_t)q * TILE, da[q]);
#pragma unroll
                    for (int i = 0; i < TRS_WT; ++i) {
                        if (i < nrow) {
                            T x[VEC];
                            SpecSm<T, VEC>::ld(xs + (size_t)i * g.D, x);
#pragma unroll
                            for (int q = 0; q < QT; ++q)
#pragma unroll
                                for (int v = 0; v < VEC; ++v) acc[i][q] = tr_fma<T>(x[v], da[q][v], acc[i][q]);
                        }
                    }
                }
                __syncwarp();                                                // every lane has read the stage
                if (lane == 0) trf::mbar_arrive(&empty[s]);
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
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < TRSS_NF; ++i) tot += sloss[i];
        a.losspart[blockIdx.x] = tot;
    }
}
