// tr_api.cu — C ABI (include/tr_b200.h) over the kernels in tr_kernels.cuh.
// Host-side planning only: tile / sample-group geometry, workspace, kernel dispatch, launches.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tr_kernels.cuh"
#include "tr_dispatch.h"
#include "tr_small.cuh"
#include "tr_spectral.cuh"
#include "tr_spectral_single.h"
#include "tr_fused.cuh"
#include "tr_fused_mn.h"

namespace {

thread_local std::string g_create_error;

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

// streaming-kernel tables are compiled in separate translation units (tr_stream.cu, one per
// TR_PART) so the build parallelises; merged here on first use.
template <typename T> std::vector<KEntry<T>>& merged();
template <> std::vector<KEntry<float>>& merged<float>() {
    static std::vector<KEntry<float>> v = [] {
        std::vector<KEntry<float>> t;
        int n; const KEntry<float>* p;
        p = tr_entries_f32_0(&n); t.insert(t.end(), p, p + n);
        p = tr_entries_f32_1(&n); t.insert(t.end(), p, p + n);
        p = tr_entries_f32_2(&n); t.insert(t.end(), p, p + n);
        p = tr_entries_f32_3(&n); t.insert(t.end(), p, p + n);
        std::sort(t.begin(), t.end(), [](const KEntry<float>& a, const KEntry<float>& b) { return a.RK < b.RK; });
        return t;
    }();
    return v;
}
template <> std::vector<KEntry<double>>& merged<double>() {
    static std::vector<KEntry<double>> v = [] {
        std::vector<KEntry<double>> t;
        int n; const KEntry<double>* p;
        p = tr_entries_f64_0(&n); t.insert(t.end(), p, p + n);
        p = tr_entries_f64_1(&n); t.insert(t.end(), p, p + n);
        p = tr_entries_f64_2(&n); t.insert(t.end(), p, p + n);
        p = tr_entries_f64_3(&n); t.insert(t.end(), p, p + n);
        std::sort(t.begin(), t.end(), [](const KEntry<double>& a, const KEntry<double>& b) { return a.RK < b.RK; });
        return t;
    }();
    return v;
}

template <typename T>
const KEntry<T>* pick_entry(int rk) {
    for (const KEntry<T>& e : merged<T>())
        if (e.RK >= rk) return &e;
    return nullptr;
}

struct Plan {
    int RKs;            // instantiated channel count (>= needed)
    int tile;           // elements per warp tile
    int WT;             // warp tiles per sample
    long long Dpad;
    int Gn_f, grid_f;
    int Gn_g, grid_g;
    int nchunk;
    long long spc;
    int vec;            // 1 = 16-byte loads, 0 = element loads
    size_t smem_f;
};

}  // namespace

struct tr_handle {
    int dtype = 0, device = 0, sms = 0;
    size_t l2_bytes = 0;
    size_t elt = 4;
    Geo geo;
    std::string err;
    std::map<const void*, int> occ;     // kernel -> resident blocks per SM
    Buf FtT, Ft64, partial, V, u_ws, dZ_ws, Gpart, Gred, epi_part, dfc_part;
    long long info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int launches = 0;
    // optional in-stream timing of the two streaming kernels (tr_profile_*)
    bool prof = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // fwd, grad, fused: begin/end
    bool ev_set[3] = {false, false, false};
    double prof_ms[3] = {0.0, 0.0, 0.0};
    long long prof_n[3] = {0, 0, 0};
    int fused_mode = -1;                // -1 auto, 0 never, 1 always (error when not eligible)
    int fused_pace = 0;                 // cycles between TMA issues (tuning knob, option "fused_pace")
    int fused_piece = 32768;            // bytes per bulk-copy instruction (option "fused_piece")
    int last_fused = 0;
    int fused_ns = 0;                   // testing knob (option "fused_ns", env TR_B200_FUSED_NS): cap on the shared-memory stages of k_fused_mn
    int fused_cl = 0;                   // testing knob (option "fused_cl", env TR_B200_FUSED_CL): force the cluster size of the single-pass cluster kernels
    std::map<const void*, int> occ_clusters;
    int flow_mode = 0;                  // dataflow kernel (tr_flow.cuh), experimental: 0 never (default), 1 always (error when
                                        // not eligible); -1 is accepted and currently means 0 (it measured slower)
    long long flow_window_mb = 32;      // bytes of X the forward warps may lead the gradient warps by (option "flow_window_mb")
    Buf flow_ring, flow_sync;
    Buf Apart, Spart;                   // single-pass multinomial kernel: flushed A / S partial sums
    Buf trace;                          // debug timeline of k_fused_mn (builds with -DTRM_TRACE only)
    Buf f3_stage;                       // last-mode factor in the constant buffer's layout (copied to the constant bank per launch)
    int flow_debug = 0;
    // spectral_tensor_regression.py handles (tr_spec_create): kind 2, geo describes the six factor blocks of theta
    int kind = 0;
    SpecGeo sg;
    Buf spA, spDA, spMc, spU, spDS, spRes, spPart, spDf1;
    int spec_single = -1;               // single-pass spectral kernel (tr_spectral_single.cuh): -1 auto, 0 never, 1 always (error when not eligible)
    int spec_single_ns = 0;             // testing knob (option "spec_single_ns"): cap on its shared-memory stages
    int smem_optin = 0;                 // cudaDevAttrMaxSharedMemoryPerBlockOptin (queried at first use)
};

// trainable scalars after the factor entries: the scalar bias of the standard model, the (n_out) bias of the spectral one
static inline int tr_nbias(const tr_handle* h) { return h->kind == 2 ? h->sg.NO : (h->geo.C == 0 ? 1 : 0); }

namespace {

int fail(tr_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define TR_CUDA(h, call)                                                                          \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(h, TR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};

int ensure(tr_handle* h, Buf& b, size_t bytes) {
    if (bytes <= b.cap) return TR_OK;
    if (b.p) { TR_CUDA(h, cudaDeviceSynchronize()); TR_CUDA(h, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    const size_t want = bytes + bytes / 16 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(h, TR_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e)); }
    b.cap = want;
    return TR_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (device, kernel), not of a handle: two handles
// that use the same instantiation with different shared-memory sizes must never LOWER it under each other.
// The library keeps a process-wide high-water mark per (device, kernel) and only ever raises the attribute;
// every launch site calls this right before its launch.
template <typename K>
int raise_smem_limit(tr_handle* h, K kern, size_t smem) {
    if (smem <= 48 * 1024) return TR_OK;
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> high;
    std::lock_guard<std::mutex> lk(mu);
    size_t& cur = high[std::make_pair(h->device, (const void*)kern)];
    if (smem > cur) {
        TR_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    return TR_OK;
}

template <typename K>
int occupancy(tr_handle* h, K kern, size_t smem, int* out) {
    // occupancy depends on the shared-memory size of THIS handle's launches: keyed by (kernel, smem)
    const void* key = (const void*)((uintptr_t)kern ^ ((uintptr_t)smem << 20));
    auto it = h->occ.find(key);
    if (it != h->occ.end()) { *out = it->second; return TR_OK; }
    { int rc = raise_smem_limit(h, kern, smem); if (rc) return rc; }
    int nb = 0;
    TR_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, TR_TPB, smem));
    if (nb < 1) return fail(h, TR_ERR_UNSUPPORTED, "kernel does not fit on an SM (dynamic smem %zu bytes)", smem);
    h->occ[key] = nb;
    *out = nb;
    return TR_OK;
}

template <typename T>
int make_plan(tr_handle* h, long long N, int rk_needed, bool vec, Plan* pl, const KEntry<T>** ent) {
    const KEntry<T>* e = pick_entry<T>(rk_needed);
    if (!e) return fail(h, TR_ERR_UNSUPPORTED, "rank %d exceeds the %d channels the streaming kernels are built for", rk_needed, TR_MAX_RANK_MN);
    *ent = e;
    pl->RKs = e->RK;
    pl->vec = vec ? 1 : 0;
    pl->tile = 32 * e->E * VN<T>::v;
    const long long D = h->geo.D;
    pl->WT = (int)((D + pl->tile - 1) / pl->tile);
    pl->Dpad = (long long)pl->WT * pl->tile;
    pl->smem_f = (size_t)(h->geo.pfeat + h->geo.R) * sizeof(T);
    if (pl->smem_f > 200 * 1024)
        return fail(h, TR_ERR_UNSUPPORTED, "factor rows (%zu bytes) do not fit in shared memory", pl->smem_f);
    int occ_f = 0, occ_g = 0, rc;
    if ((rc = occupancy(h, vec ? e->fwd_vec : e->fwd_sc, pl->smem_f, &occ_f))) return rc;
    if ((rc = occupancy(h, vec ? e->grad_vec : e->grad_sc, 0, &occ_g))) return rc;
    auto groups = [&](int occ, int* Gn, int* grid) {
        const long long wtot = (long long)h->sms * occ * TR_WPB;
        if (pl->WT <= wtot) {
            long long g = wtot / pl->WT;
            if (g > N) g = N;
            if (g < 1) g = 1;
            *Gn = (int)g;
            *grid = (int)(((long long)pl->WT * g + TR_WPB - 1) / TR_WPB);
        } else {
            *Gn = 1;
            *grid = h->sms * occ;
        }
    };
    groups(occ_f, &pl->Gn_f, &pl->grid_f);
    groups(occ_g, &pl->Gn_g, &pl->grid_g);
    // bound the length of each fp32 running sum in the gradient pass
    const long long Sg = (N + pl->Gn_g - 1) / pl->Gn_g;
    const long long target = sizeof(T) == 4 ? 2048 : (1LL << 40);
    long long nchunk = std::max<long long>(1, (Sg + target - 1) / target);
    const size_t slot_bytes = (size_t)pl->RKs * (size_t)pl->Dpad * sizeof(T);
    const size_t cap_bytes = (size_t)1 << 30;
    while (nchunk > 1 && (size_t)nchunk * pl->Gn_g * slot_bytes > cap_bytes) --nchunk;
    pl->nchunk = (int)nchunk;
    pl->spc = std::max<long long>(1, (Sg + nchunk - 1) / nchunk);
    return TR_OK;
}

template <typename T>
int reserve_for(tr_handle* h, long long N, const Plan& pl) {
    int rc;
    const Geo& g = h->geo;
    if ((rc = ensure(h, h->FtT, (size_t)g.pf * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->Ft64, (size_t)g.pf * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->partial, (size_t)N * pl.WT * pl.RKs * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->V, (size_t)N * pl.RKs * sizeof(T)))) return rc;
    if (g.C > 0) {
        if ((rc = ensure(h, h->u_ws, (size_t)N * g.R * sizeof(T)))) return rc;
        if ((rc = ensure(h, h->dZ_ws, (size_t)N * g.C * sizeof(T)))) return rc;
        if ((rc = ensure(h, h->dfc_part, (size_t)h->sms * 2 * g.C * g.R * sizeof(double)))) return rc;
    }
    if ((rc = ensure(h, h->Gpart, (size_t)pl.nchunk * pl.Gn_g * pl.RKs * pl.Dpad * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->Gred, (size_t)pl.RKs * g.D * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->epi_part, (size_t)h->sms * 8 * 2 * sizeof(double)))) return rc;
    return TR_OK;
}

inline bool vec_ok(const void* X, long long D, size_t elt) {
    return ((uintptr_t)X % 16 == 0) && ((D * (long long)elt) % 16 == 0);
}

#define TR_LAUNCH_CHECK(h)                                                                        \
    do {                                                                                          \
        cudaError_t e_ = cudaPeekAtLastError();                                                   \
        if (e_ != cudaSuccess) { cudaGetLastError(); return fail(h, TR_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); } \
        ++(h)->launches;                                                                          \
    } while (0)

// Adds the elapsed time of the previous (already finished) recorded launches to the running sums.
int prof_fold(tr_handle* h, bool wait) {
    for (int i = 0; i < 3; ++i) {
        if (!h->ev_set[i]) continue;
        if (wait) TR_CUDA(h, cudaEventSynchronize(h->ev[2 * i + 1]));
        else if (cudaEventQuery(h->ev[2 * i + 1]) != cudaSuccess) { cudaGetLastError(); TR_CUDA(h, cudaEventSynchronize(h->ev[2 * i + 1])); }
        float ms = 0.f;
        TR_CUDA(h, cudaEventElapsedTime(&ms, h->ev[2 * i], h->ev[2 * i + 1]));
        h->prof_ms[i] += ms;
        h->prof_n[i] += 1;
        h->ev_set[i] = false;
    }
    return TR_OK;
}

// prep + pass 1 (+ nothing else): fills h->partial
template <typename T>
int run_forward(tr_handle* h, const T* X, long long N, const T* theta, const T* w, uint32_t nn_mask,
                double beta, double thr, const Plan& pl, const KEntry<T>* e, cudaStream_t st) {
    const Geo& g = h->geo;
    k_prep<T><<<std::max(1, std::min(64, (g.pf + 255) / 256)), 256, 0, st>>>(theta, g, nn_mask, beta, thr,
                                                                           (T*)h->FtT.p, (double*)h->Ft64.p);
    TR_LAUNCH_CHECK(h);
    FwdArgs<T> fa;
    fa.X = X; fa.N = N; fa.FtT = (const T*)h->FtT.p; fa.w = w; fa.geo = g;
    fa.partial = (T*)h->partial.p; fa.WT = pl.WT; fa.Gn = pl.Gn_f; fa.mode = g.C > 0 ? 1 : 0;
    auto kern = pl.vec ? e->fwd_vec : e->fwd_sc;
    { int rc = raise_smem_limit(h, kern, pl.smem_f); if (rc) return rc; }
    if (h->prof) { int rc = prof_fold(h, false); if (rc) return rc; TR_CUDA(h, cudaEventRecord(h->ev[0], st)); }
    kern<<<pl.grid_f, TR_TPB, pl.smem_f, st>>>(fa);
    TR_LAUNCH_CHECK(h);
    if (h->prof) { TR_CUDA(h, cudaEventRecord(h->ev[1], st)); h->ev_set[0] = true; }
    return TR_OK;
}

// pass 2 + reduction + all-mode MTTKRP: consumes V (N, RKs), fills gradsum[0 .. pfeat)
template <typename T>
int run_gradient(tr_handle* h, const T* X, long long N, const T* V, const T* w, const Plan& pl,
                 const KEntry<T>* e, double* gradsum, cudaStream_t st) {
    const Geo& g = h->geo;
    GradArgs<T> ga;
    ga.X = X; ga.N = N; ga.D = g.D; ga.V = V; ga.Gpart = (T*)h->Gpart.p; ga.Dpad = pl.Dpad;
    ga.WT = pl.WT; ga.Gn = pl.Gn_g; ga.nchunk = pl.nchunk; ga.spc = pl.spc;
    auto kern = pl.vec ? e->grad_vec : e->grad_sc;
    if (h->prof) TR_CUDA(h, cudaEventRecord(h->ev[2], st));
    kern<<<pl.grid_g, TR_TPB, 0, st>>>(ga);
    TR_LAUNCH_CHECK(h);
    if (h->prof) { TR_CUDA(h, cudaEventRecord(h->ev[3], st)); h->ev_set[1] = true; }
    const long long tot = (long long)pl.RKs * g.D;
    const int rgrid = (int)std::min<long long>((tot + 255) / 256, (long long)h->sms * 8);
    k_reduce_G<T><<<rgrid, 256, 0, st>>>((const T*)h->Gpart.p, pl.nchunk * pl.Gn_g, pl.RKs, g.D, pl.Dpad,
                                         (double*)h->Gred.p);
    TR_LAUNCH_CHECK(h);
    MtArgs ma;
    ma.G = (const double*)h->Gred.p; ma.Ft64 = (const double*)h->Ft64.p; ma.w = w;
    ma.w_is_f64 = sizeof(T) == 8; ma.per_rank = g.C > 0 ? 1 : 0; ma.geo = g; ma.gradsum = gradsum;
    int rows = 0;
    for (int m = 0; m < g.k; ++m) rows += g.dims[m];
    k_mttkrp<<<rows, TR_TPB, 0, st>>>(ma);
    TR_LAUNCH_CHECK(h);
    return TR_OK;
}

// ---------------------------------------------------------------------------------------------
// single-pass fused path (standard model): plan + launch
// ---------------------------------------------------------------------------------------------
struct FusedPlan {
    int CL, E, NS, NC, nchunk;
    long long spc;
    unsigned stage_bytes;
    size_t smem;
    int Dc;
};

template <typename T> using FusedKern = void (*)(FusedArgs<T>);
template <typename T>
FusedKern<T> fused_kernel(int E) {
    switch (E) {
        case 1: return k_fused_std<T, 1>;
        case 2: return k_fused_std<T, 2>;
        case 3: return k_fused_std<T, 3>;
        case 4: return k_fused_std<T, 4>;
        case 5: return k_fused_std<T, 5>;
        case 6: return k_fused_std<T, 6>;
        case 7: return k_fused_std<T, 7>;
        case 8: return k_fused_std<T, 8>;
        case 9: case 10: return k_fused_std<T, 10>;
        case 11: case 12: return k_fused_std<T, 12>;
        case 13: case 14: return k_fused_std<T, 14>;
        case 15: case 16: return k_fused_std<T, 16>;
    }
    return nullptr;
}

// returns TR_OK and fp->CL > 0 when the geometry fits a cluster's shared memory, fp->CL = 0 otherwise
template <typename T>
int plan_fused(tr_handle* h, long long N, const void* X, FusedPlan* fp) {
    fp->CL = 0;
    const Geo& g = h->geo;
    constexpr int VEC = 16 / (int)sizeof(T);
    if (g.C != 0 || !vec_ok(X, g.D, sizeof(T)) || N < 1) return TR_OK;
    const size_t fixed = ((((sizeof(FusedCtl) + 15) / 16) * 16 + (size_t)(g.pfeat + g.R) * sizeof(T)) + 1023) / 1024 * 1024;
    const size_t budget = 226 * 1024;
    // The kernel takes any cluster size (option "fused_cl"); how many clusters of a size are resident depends on the GPC
    // layout (B200, one CTA per SM: 15 clusters of 8 = 120 SMs, 15 clusters of 9 = 135 SMs).  Measured on cfg 2: 9 instead
    // of 8 changes nothing (16.1-16.3 ms either way: the kernel is bound by HBM, not by the SMs it covers), so the automatic
    // choice stays with the powers of two — the size that covers the most SMs, ties to the smaller cluster.
    long long best_cover = 0;
    for (int CL = 1; CL <= TR_FUSED_MAX_CL; ++CL) {
        if (h->fused_cl > 0 ? CL != h->fused_cl : (CL & (CL - 1)) != 0) continue;
        if (CL > g.D / VEC) break;                                           // every CTA needs at least one 16-byte chunk
        // ragged slices: the first (D/VEC mod CL) CTAs of a cluster hold one 16-byte chunk more
        const long long chunks = (g.D / VEC + CL - 1) / CL;
        const long long Dc = chunks * VEC;
        const int E = (int)((chunks + TR_FUSED_NCT - 1) / TR_FUSED_NCT);
        if (E > 16) continue;
        const size_t stage = (size_t)Dc * sizeof(T);
        if (fixed + 3 * stage > budget) continue;
        int NS = (int)((budget - fixed) / stage);
        if (NS > TR_FUSED_MAX_NS) NS = TR_FUSED_MAX_NS;
        auto kern = fused_kernel<T>(E);
        const size_t smem = fixed + (size_t)NS * stage;
        // resident clusters of this (kernel, cluster size): queried once per handle
        const void* key = (const void*)(((uintptr_t)kern + (uintptr_t)CL) ^ ((uintptr_t)smem << 20));
        int NC = 0;
        auto it = h->occ_clusters.find(key);
        if (it != h->occ_clusters.end()) {
            NC = it->second;
        } else {
            { int rc = raise_smem_limit(h, kern, smem); if (rc) return rc; }
            if (CL > 8) {
                if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); h->occ_clusters[key] = 0; continue; }
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(CL * h->sms), 1, 1);
            cfg.blockDim = dim3(TR_FUSED_NT, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&NC, kern, &cfg);
            if (e != cudaSuccess) { cudaGetLastError(); NC = 0; }
            h->occ_clusters[key] = NC;
        }
        if (NC < 1) continue;
        if ((long long)NC * CL * 10 < (long long)h->sms * 6) continue;       // would leave > 40 % of the SMs idle
        if ((long long)NC * CL <= best_cover) continue;
        best_cover = (long long)NC * CL;
        if (NC > N) NC = (int)N;
        fp->CL = CL; fp->E = E; fp->NS = NS; fp->NC = NC; fp->stage_bytes = (unsigned)stage; fp->smem = smem; fp->Dc = (int)Dc;
        const long long cnt = (N + NC - 1) / NC;
        const long long target = sizeof(T) == 4 ? 2048 : (1LL << 40);
        long long nchunk = std::max<long long>(1, (cnt + target - 1) / target);
        const size_t slot_bytes = (size_t)g.D * sizeof(T);
        while (nchunk > 1 && (size_t)nchunk * NC * slot_bytes > ((size_t)1 << 30)) --nchunk;
        fp->nchunk = (int)nchunk;
        fp->spc = std::max<long long>(1, (cnt + nchunk - 1) / nchunk);
    }
    return TR_OK;
}

template <typename T>
int run_fused_std(tr_handle* h, const T* X, const T* y, long long N, const T* theta, const T* w, uint32_t nn_mask,
                  double beta, double thr, const FusedPlan& fp, double* gradsum, T* yhat, cudaStream_t st) {
    const Geo& g = h->geo;
    int rc;
    if ((rc = ensure(h, h->FtT, (size_t)g.pf * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->Ft64, (size_t)g.pf * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->Gpart, (size_t)fp.NC * fp.nchunk * g.D * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->Gred, (size_t)g.D * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->epi_part, (size_t)h->sms * 8 * 2 * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->V, (size_t)N * sizeof(T)))) return rc;
    k_prep<T><<<std::max(1, std::min(64, (g.pf + 255) / 256)), 256, 0, st>>>(theta, g, nn_mask, beta, thr,
                                                                           (T*)h->FtT.p, (double*)h->Ft64.p);
    TR_LAUNCH_CHECK(h);
    FusedArgs<T> fa;
    fa.X = X; fa.y = y; fa.N = N; fa.FtT = (const T*)h->FtT.p; fa.w = w; fa.theta = theta; fa.bias_off = g.pf; fa.geo = g;
    fa.Gpart = (T*)h->Gpart.p; fa.Dpad = g.D; fa.yhat = yhat; fa.res = (T*)h->V.p;
    fa.CL = fp.CL; fa.NC = fp.NC; fa.Dc = fp.Dc; fa.NS = fp.NS; fa.nchunk = fp.nchunk; fa.spc = fp.spc;
    fa.stage_bytes = fp.stage_bytes;
    fa.trace = nullptr;
    fa.pace = h->fused_pace;
    fa.piece = (unsigned)h->fused_piece;
    auto kern = fused_kernel<T>(fp.E);
    if ((rc = raise_smem_limit(h, kern, fp.smem))) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(fp.CL * fp.NC), 1, 1);
    cfg.blockDim = dim3(TR_FUSED_NT, 1, 1);
    cfg.dynamicSmemBytes = fp.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)fp.CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (h->prof) { rc = prof_fold(h, false); if (rc) return rc; TR_CUDA(h, cudaEventRecord(h->ev[4], st)); }
    TR_CUDA(h, cudaLaunchKernelEx(&cfg, kern, fa));
    TR_LAUNCH_CHECK(h);
    if (h->prof) { TR_CUDA(h, cudaEventRecord(h->ev[5], st)); h->ev_set[2] = true; }
    const int sgrid = (int)std::min<long long>((N + 255) / 256, (long long)h->sms * 2);
    k_ressum<T><<<sgrid, 256, 0, st>>>((const T*)h->V.p, N, (double*)h->epi_part.p);
    TR_LAUNCH_CHECK(h);
    k_colsum<<<2, 128, 0, st>>>((const double*)h->epi_part.p, sgrid, 2, gradsum + g.pf);
    TR_LAUNCH_CHECK(h);
    const int rgrid = (int)std::min<long long>((g.D + 255) / 256, (long long)h->sms * 8);
    k_reduce_G<T><<<rgrid, 256, 0, st>>>((const T*)h->Gpart.p, fp.NC * fp.nchunk, 1, g.D, g.D, (double*)h->Gred.p);
    TR_LAUNCH_CHECK(h);
    MtArgs ma;
    ma.G = (const double*)h->Gred.p; ma.Ft64 = (const double*)h->Ft64.p; ma.w = w;
    ma.w_is_f64 = sizeof(T) == 8; ma.per_rank = 0; ma.geo = g; ma.gradsum = gradsum;
    int rows = 0;
    for (int m = 0; m < g.k; ++m) rows += g.dims[m];
    k_mttkrp<<<rows, TR_TPB, 0, st>>>(ma);
    TR_LAUNCH_CHECK(h);
    h->info[0] = h->launches; h->info[1] = fp.CL * fp.NC; h->info[2] = fp.CL; h->info[3] = fp.NS;
    h->info[4] = fp.NC; h->info[5] = fp.nchunk; h->info[6] = 1; h->info[7] = -(16 / (int)sizeof(T));
    return TR_OK;
}


// ---------------------------------------------------------------------------------------------
// single-pass fused path (multinomial model, tr_fused_mn.cuh): plan + launch
// ---------------------------------------------------------------------------------------------
struct FusedMnPlan {
    int ok;
    int CL, NS, NC, nchunk, IKC, RKS, rows_max, NR;
    long long spc;
    unsigned stage_x, stage_t, head;
    size_t smem;
    const void* kern;
    const void* f3_symbol;
};

// Launches that share a constant buffer (one per translation unit of tr_fusedmn.cu) are ordered: the buffer is
// rewritten before every launch, so a launch on another stream must wait for the previous reader.
struct ConstUse { cudaStream_t stream = nullptr; cudaEvent_t done = nullptr; };
std::mutex g_const_mu;
std::map<std::pair<int, const void*>, ConstUse> g_const_use;

template <typename T> const TrmEntry* trm_find(int IKC, int RKS);
static const TrmEntry* trm_scan(const TrmEntry* t, int n, int IKC, int RKS) {
    for (int i = 0; i < n; ++i)
        if (t[i].IKC == IKC && t[i].RKS == RKS) return &t[i];
    return nullptr;
}
template <> const TrmEntry* trm_find<float>(int IKC, int RKS) {
    int n; const TrmEntry* t; const TrmEntry* e;
    t = trm_entries_f32_0(&n); if ((e = trm_scan(t, n, IKC, RKS))) return e;
    t = trm_entries_f32_1(&n); if ((e = trm_scan(t, n, IKC, RKS))) return e;
    t = trm_entries_f32_2(&n); if ((e = trm_scan(t, n, IKC, RKS))) return e;
    return nullptr;
}
template <> const TrmEntry* trm_find<double>(int IKC, int RKS) {
    int n; const TrmEntry* t = trm_entries_f64_0(&n);
    return trm_scan(t, n, IKC, RKS);
}

// fp->ok = 1 when the geometry fits: last feature mode = 2..8 16-byte chunks, rank <= 8, classes <= 32, the rows of
// a sample spread over a cluster with <= TRM_GMAX*128 rows per CTA and >= 4 shared-memory stages
template <typename T>
int plan_fused_mn(tr_handle* h, long long N, const void* X, FusedMnPlan* fp) {
    fp->ok = 0;
    const Geo& g = h->geo;
    constexpr int VEC = 16 / (int)sizeof(T);
    if (g.C < 1 || g.C > 32 || g.R > TRM_RKMAX || g.k < 2 || !vec_ok(X, g.D, sizeof(T)) || N < 1) return TR_OK;
    const int IK = g.dims[g.k - 1];
    if (IK % VEC != 0) return TR_OK;
    const int IKC = IK / VEC;
    const TrmEntry* ent = nullptr;
    int RKS = std::max(2, (g.R + 1) / 2 * 2);
    for (; RKS <= TRM_RKMAX && !(ent = trm_find<T>(IKC, RKS)); RKS += 2) {}
    if (!ent) return TR_OK;
    const int NR = (int)(g.D / IK);
    size_t head = (sizeof(FusedMnCtl<T>) + 15) / 16 * 16;
    head += (size_t)TRM_F12_ROWS * RKS * sizeof(T);
    head = (head + 15) / 16 * 16;
    head += (size_t)g.C * (RKS + 1) * sizeof(T);
    head = (head + 15) / 16 * 16;
    head += (size_t)2 * TRM_NFT * RKS * sizeof(T);          // the forward threads' partials of u, two samples deep
    head = (head + 127) / 128 * 128;
    #ifdef TRM_TRACE
    const size_t budget = 218 * 1024;       // room for the timeline stamps
#else
    const size_t budget = 226 * 1024;
#endif
    int bestCL = 0, bestNS = 0;
    unsigned best_sx = 0, best_st = 0;
    for (int CL = 1; CL <= TRM_MAX_CL; CL *= 2) {
        if (h->fused_cl > 0 && CL != h->fused_cl) continue;
        const int rows_max = (NR + CL - 1) / CL;
        if (rows_max > TRM_GMAX * TRM_NCT || NR < CL) continue;    // every CTA of the cluster needs at least one row
        const size_t sx = ((size_t)rows_max * IK * sizeof(T) + 127) / 128 * 128;
#if TRM_TMEM
        const size_t st = 0;                                             // t[row,:] waits in tensor memory
#else
        const size_t st = ((size_t)rows_max * RKS * sizeof(T) + 15) / 16 * 16;
#endif
        if (head + 4 * (sx + st) > budget) continue;
        int NS = (int)((budget - head) / (sx + st));
        if (NS > TRM_MAX_NS) NS = TRM_MAX_NS;
        if (h->fused_ns >= 4 && NS > h->fused_ns) NS = h->fused_ns;
        // the smallest cluster that still gives a deep pipeline (>= 6 stages); otherwise the deepest one
        const bool better = bestCL == 0 || (bestNS < 6 && NS > bestNS);
        if (better) { bestCL = CL; bestNS = NS; best_sx = (unsigned)sx; best_st = (unsigned)st; }
    }
    if (bestCL == 0) return TR_OK;
    const int CL = bestCL, NS = bestNS;
    const size_t smem = head + (size_t)NS * (best_sx + best_st);
    const void* key = (const void*)(((uintptr_t)ent->kern + (uintptr_t)CL) ^ ((uintptr_t)smem << 20));
    int NC = 0;
    auto it = h->occ_clusters.find(key);
    if (it != h->occ_clusters.end()) {
        NC = it->second;
    } else {
        { int rc = raise_smem_limit(h, ent->kern, smem); if (rc) return rc; }
        bool okc = true;
        if (CL > 8 && cudaFuncSetAttribute(ent->kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); okc = false; }
        if (okc) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(CL * h->sms), 1, 1);
            cfg.blockDim = dim3(TRM_NT, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&NC, ent->kern, &cfg);
            if (e != cudaSuccess) { cudaGetLastError(); NC = 0; }
        }
        h->occ_clusters[key] = NC;
    }
    if (NC < 1) return TR_OK;
    if (h->fused_cl == 0 && (long long)NC * CL * 10 < (long long)h->sms * 6) return TR_OK;   // would leave > 40 % of the SMs idle
    if (NC > N) NC = (int)N;
    fp->CL = CL; fp->NS = NS; fp->NC = NC; fp->IKC = IKC; fp->RKS = RKS; fp->NR = NR;
    fp->rows_max = (NR + CL - 1) / CL;
    fp->stage_x = best_sx; fp->stage_t = best_st; fp->head = (unsigned)head; fp->smem = smem; fp->kern = ent->kern;
    fp->f3_symbol = ent->f3_symbol;
    const long long cnt = (N + NC - 1) / NC;
    const long long target = sizeof(T) == 4 ? 2048 : (1LL << 40);
    long long nchunk = std::max<long long>(1, (cnt + target - 1) / target);
    fp->nchunk = (int)nchunk;
    fp->spc = std::max<long long>(1, (cnt + nchunk - 1) / nchunk);
    fp->ok = 1;
    return TR_OK;
}

template <typename T> int launch_dfc(tr_handle* h, int dgrid, const T* w, long long N, cudaStream_t st);

template <typename T>
int run_fused_mn(tr_handle* h, const T* X, const long long* y, const T* class_w, long long N, const T* theta, const T* w,
                 uint32_t nn_mask, double beta, double thr, const FusedMnPlan& fp, double* gradsum, T* P, cudaStream_t st) {
    const Geo& g = h->geo;
    constexpr int VEC = 16 / (int)sizeof(T);
    const int IK = fp.IKC * VEC;
    int rc;
    const size_t slots = (size_t)fp.NC * fp.nchunk;
    const size_t a_bytes = slots * fp.CL * TRM_NWG * (size_t)IK * fp.RKS * sizeof(T);
    const size_t s_bytes = slots * (size_t)fp.RKS * fp.NR * sizeof(T);
    if ((rc = ensure(h, h->FtT, (size_t)g.pf * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->Ft64, (size_t)g.pf * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->u_ws, (size_t)N * g.R * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->dZ_ws, (size_t)N * g.C * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->dfc_part, (size_t)h->sms * 2 * g.C * g.R * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->epi_part, std::max((size_t)fp.NC * fp.CL, (size_t)h->sms * 8) * 2 * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->Apart, a_bytes))) return rc;
    if ((rc = ensure(h, h->Spart, s_bytes))) return rc;
    if ((rc = ensure(h, h->Gred, (size_t)fp.RKS * fp.NR * sizeof(double)))) return rc;
    k_prep<T><<<std::max(1, std::min(64, (g.pf + 255) / 256)), 256, 0, st>>>(theta, g, nn_mask, beta, thr,
                                                                           (T*)h->FtT.p, (double*)h->Ft64.p);
    TR_LAUNCH_CHECK(h);
    // slots of chunks a cluster never reaches must be defined for the reductions
    TR_CUDA(h, cudaMemsetAsync(h->Apart.p, 0, a_bytes, st));
    TR_CUDA(h, cudaMemsetAsync(h->Spart.p, 0, s_bytes, st));
    FusedMnArgs<T> fa;
    memset(&fa, 0, sizeof(fa));
    fa.X = X; fa.y = y; fa.class_w = class_w; fa.N = N; fa.FtT = (const T*)h->FtT.p; fa.Ft64 = (const double*)h->Ft64.p;
    fa.w = w; fa.geo = g; fa.NR = fp.NR; fa.Apart = (T*)h->Apart.p; fa.Spart = (T*)h->Spart.p; fa.P = P;
    fa.u_ws = (T*)h->u_ws.p; fa.dZ_ws = (T*)h->dZ_ws.p; fa.losspart = (double*)h->epi_part.p;
    fa.CL = fp.CL; fa.NC = fp.NC; fa.NS = fp.NS; fa.nchunk = fp.nchunk; fa.spc = fp.spc;
    fa.stage_x_bytes = fp.stage_x; fa.stage_t_bytes = fp.stage_t; fa.head_bytes = fp.head;
    fa.piece = (unsigned)h->fused_piece;
    size_t smem_launch = fp.smem;
#ifdef TRM_TRACE
    if ((rc = ensure(h, h->trace, (size_t)TRM_TRACE_N * TRM_TRACE_EV * sizeof(long long)))) return rc;
    fa.trace = (long long*)h->trace.p;
    fa.dbg = h->flow_debug;                       // timing experiments (option "flow_debug"), trace builds only
    smem_launch += (size_t)TRM_TRACE_N * TRM_TRACE_EV * sizeof(long long);
#endif
    if ((rc = raise_smem_limit(h, fp.kern, smem_launch))) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(fp.CL * fp.NC), 1, 1);
    cfg.blockDim = dim3(TRM_NT, 1, 1);
    cfg.dynamicSmemBytes = smem_launch;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)fp.CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    {
        // last-mode factor -> the kernel's constant buffer (stream-ordered); launches sharing the buffer are serialised
        const size_t f3_bytes = (size_t)IK * fp.RKS * sizeof(T);
        if ((rc = ensure(h, h->f3_stage, f3_bytes))) return rc;
        k_pack_f3<T><<<1, 256, 0, st>>>((const T*)h->FtT.p + g.foff[g.k - 1], IK, g.R, fp.RKS, (T*)h->f3_stage.p);
        TR_LAUNCH_CHECK(h);
        std::lock_guard<std::mutex> lk(g_const_mu);
        ConstUse& cu = g_const_use[std::make_pair(h->device, fp.f3_symbol)];
        if (cu.done && cu.stream != st) TR_CUDA(h, cudaStreamWaitEvent(st, cu.done, 0));
        TR_CUDA(h, cudaMemcpyToSymbolAsync(fp.f3_symbol, h->f3_stage.p, f3_bytes, 0, cudaMemcpyDeviceToDevice, st));
        if (h->prof) { rc = prof_fold(h, false); if (rc) return rc; TR_CUDA(h, cudaEventRecord(h->ev[4], st)); }
        void* kargs[] = {(void*)&fa};
        TR_CUDA(h, cudaLaunchKernelExC(&cfg, fp.kern, kargs));
        TR_LAUNCH_CHECK(h);
        if (h->prof) { TR_CUDA(h, cudaEventRecord(h->ev[5], st)); h->ev_set[2] = true; }
        if (!cu.done) TR_CUDA(h, cudaEventCreateWithFlags(&cu.done, cudaEventDisableTiming));
        TR_CUDA(h, cudaEventRecord(cu.done, st));
        cu.stream = st;
    }
    // loss, class-factor gradient
    k_colsum<<<1, 128, 0, st>>>((const double*)h->epi_part.p, fp.NC * fp.CL, 1, gradsum + g.pf);
    TR_LAUNCH_CHECK(h);
    const int dgrid = (int)std::min<long long>((N + 63) / 64, (long long)h->sms * 2);
    if ((rc = launch_dfc<T>(h, dgrid, w, N, st))) return rc;
    k_colsum<<<g.C * g.R, 128, 0, st>>>((const double*)h->dfc_part.p, dgrid, g.C * g.R, gradsum + g.pfeat);
    TR_LAUNCH_CHECK(h);
    // last feature mode: sum of the flushed A partials
    k_reduce_A<T><<<IK * g.R, 128, 0, st>>>((const T*)h->Apart.p, (int)(slots * fp.CL * TRM_NWG), IK, fp.RKS, g.R,
                                            gradsum + g.foff[g.k - 1]);
    TR_LAUNCH_CHECK(h);
    // other feature modes: S (rows x channels) summed over the slots, then its (k-1)-mode MTTKRP
    const long long tot = (long long)fp.RKS * fp.NR;
    const int rgrid = (int)std::min<long long>((tot + 255) / 256, (long long)h->sms * 8);
    k_reduce_G<T><<<rgrid, 256, 0, st>>>((const T*)h->Spart.p, (int)slots, fp.RKS, fp.NR, fp.NR, (double*)h->Gred.p);
    TR_LAUNCH_CHECK(h);
    MtArgs ma;
    ma.G = (const double*)h->Gred.p; ma.Ft64 = (const double*)h->Ft64.p; ma.w = w;
    ma.w_is_f64 = sizeof(T) == 8; ma.per_rank = 1; ma.geo = g; ma.gradsum = gradsum;
    ma.geo.k = g.k - 1;
    ma.geo.D = fp.NR;
    int rows = 0;
    for (int m = 0; m < g.k - 1; ++m) rows += g.dims[m];
    k_mttkrp<<<rows, TR_TPB, 0, st>>>(ma);
    TR_LAUNCH_CHECK(h);
    h->info[0] = h->launches; h->info[1] = fp.CL * fp.NC; h->info[2] = fp.CL; h->info[3] = fp.NS;
    h->info[4] = fp.NC; h->info[5] = fp.nchunk; h->info[6] = fp.RKS; h->info[7] = -(16 / (int)sizeof(T));
    h->last_fused = 1;
    return TR_OK;
}

// class-factor gradient partial sums (dZ_ws, u_ws -> dfc_part)
template <typename T>
int launch_dfc(tr_handle* h, int dgrid, const T* w, long long N, cudaStream_t st) {
    const Geo& g = h->geo;
    const size_t smem = (size_t)4 * g.C * g.R * sizeof(double);
    { int rc = raise_smem_limit(h, k_dfc<T>, smem); if (rc) return rc; }
    k_dfc<T><<<dgrid, TR_TPB, smem, st>>>((const T*)h->dZ_ws.p, (const T*)h->u_ws.p, w, N, g.C, g.R, (double*)h->dfc_part.p);
    TR_LAUNCH_CHECK(h);
    return TR_OK;
}

#ifdef TR_WITH_FLOW
// ---------------------------------------------------------------------------------------------
// single-launch dataflow path (tr_flow.cuh): plan + launch
// ---------------------------------------------------------------------------------------------
struct FlowPlan {
    int ok;             // 1 when the geometry is eligible
    int RKs, tile, WT, Gn, grid, lag, ring, nchunk;
    long long Dpad, spc;
    size_t smem;
};

template <typename T>
int plan_flow(tr_handle* h, long long N, int rk_needed, const void* X, FlowPlan* fp, const KEntry<T>** ent) {
    fp->ok = 0;
    const Geo& g = h->geo;
    if (!vec_ok(X, g.D, sizeof(T)) || N < 1) return TR_OK;
    const KEntry<T>* e = pick_entry<T>(rk_needed);
    if (!e || !e->flow_vec) return TR_OK;
    *ent = e;
    fp->RKs = e->RK;
    fp->tile = 32 * e->E * VN<T>::v;
    fp->WT = (int)((g.D + fp->tile - 1) / fp->tile);
    fp->Dpad = (long long)fp->WT * fp->tile;
    // factor rows (T) | class factor + rank weights (double) | accumulator parking area of the gradient warps
    fp->smem = (((size_t)(g.pfeat + g.R) * sizeof(T) + 15) / 16) * 16 +
               (((size_t)(g.C * g.R + g.R) * sizeof(double) + 15) / 16) * 16 +
               (size_t)TR_FLOW_PAIRS * e->E * VN<T>::v * e->RK * 32 * sizeof(T);
    if (fp->smem > 200 * 1024) return TR_OK;
    int occ = 0, rc;
    if ((rc = occupancy(h, e->flow_vec, fp->smem, &occ))) return rc;
    const long long pairs = (long long)h->sms * occ * TR_FLOW_PAIRS;
    if (fp->WT > pairs) return TR_OK;                       // every item needs its own resident warp pair
    long long Gn = pairs / fp->WT;
    if (Gn > N) Gn = N;
    fp->Gn = (int)Gn;
    fp->grid = (int)(((long long)fp->WT * Gn + TR_FLOW_PAIRS - 1) / TR_FLOW_PAIRS);
    const long long Sg = (N + Gn - 1) / Gn;
    if (Sg >= (1LL << 30)) return TR_OK;
    const int umin = 2 * (e->Uf + e->Ug);
    const double per = (double)Gn * (double)g.D * sizeof(T);
    long long lag = (long long)((double)h->flow_window_mb * 1048576.0 / per);
    lag = std::min<long long>(lag, Sg);                      // no point in a window longer than the sample list
    lag = std::max<long long>(lag, umin);
    fp->lag = (int)lag;
    fp->ring = (int)(((lag + e->Uf + e->Uf - 1) / e->Uf) * e->Uf);
    const long long target = sizeof(T) == 4 ? 2048 : (1LL << 40);
    long long nchunk = std::max<long long>(1, (Sg + target - 1) / target);
    const size_t slot_bytes = (size_t)fp->RKs * (size_t)fp->Dpad * sizeof(T);
    while (nchunk > 1 && (size_t)nchunk * Gn * slot_bytes > ((size_t)1 << 30)) --nchunk;
    fp->nchunk = (int)nchunk;
    fp->spc = std::max<long long>(1, (Sg + nchunk - 1) / nchunk);
    fp->ok = 1;
    return TR_OK;
}

// forward + epilogue + gradient in one launch, then the shared tail (loss sums, class-factor gradient,
// split-N reduction, all-mode MTTKRP).  Fills gradsum completely.
template <typename T>
int run_flow(tr_handle* h, const T* X, const void* y, const T* class_w, long long N, const T* theta, const T* w,
             uint32_t nn_mask, double beta, double thr, const FlowPlan& fp, const KEntry<T>* e, double* gradsum,
             T* out /* yhat (N) or P (N,C), may be null */, cudaStream_t st) {
    const Geo& g = h->geo;
    int rc;
    const bool mn = g.C > 0;
    if ((rc = ensure(h, h->FtT, (size_t)g.pf * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->Ft64, (size_t)g.pf * sizeof(double)))) return rc;
    const size_t ring_bytes = (size_t)fp.Gn * fp.ring * fp.WT * fp.RKs * (sizeof(T) / 4) * sizeof(unsigned long long);
    if ((rc = ensure(h, h->flow_ring, ring_bytes))) return rc;
    if ((rc = ensure(h, h->flow_sync, (size_t)N * sizeof(unsigned)))) return rc;
    if ((rc = ensure(h, h->V, (size_t)N * fp.RKs * sizeof(T)))) return rc;
    if (mn) {
        if ((rc = ensure(h, h->u_ws, (size_t)N * g.R * sizeof(T)))) return rc;
        if ((rc = ensure(h, h->dZ_ws, (size_t)N * g.C * sizeof(T)))) return rc;
        if ((rc = ensure(h, h->dfc_part, (size_t)h->sms * 2 * g.C * g.R * sizeof(double)))) return rc;
    }
    if ((rc = ensure(h, h->Gpart, (size_t)fp.nchunk * fp.Gn * fp.RKs * fp.Dpad * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->Gred, (size_t)fp.RKs * g.D * sizeof(double)))) return rc;
    const int nwarps = fp.grid * TR_WPB;
    if ((rc = ensure(h, h->epi_part, std::max((size_t)nwarps, (size_t)h->sms * 8) * 2 * sizeof(double)))) return rc;

    k_prep<T><<<std::max(1, std::min(64, (g.pf + 255) / 256)), 256, 0, st>>>(theta, g, nn_mask, beta, thr,
                                                                           (T*)h->FtT.p, (double*)h->Ft64.p);
    TR_LAUNCH_CHECK(h);
    TR_CUDA(h, cudaMemsetAsync(h->flow_sync.p, 0, (size_t)N * sizeof(unsigned), st));
    TR_CUDA(h, cudaMemsetAsync(h->flow_ring.p, 0, ring_bytes, st));                     // tag 0 = never written
    TR_CUDA(h, cudaMemsetAsync(h->V.p, 0xff, (size_t)N * fp.RKs * sizeof(T), st));      // "not written yet" pattern

    FlowArgs<T> fa;
    memset(&fa, 0, sizeof(fa));
    fa.X = X; fa.N = N; fa.FtT = (const T*)h->FtT.p; fa.w = w; fa.geo = g; fa.mode = mn ? 1 : 0;
    fa.WT = fp.WT; fa.Gn = fp.Gn; fa.partial = (unsigned long long*)h->flow_ring.p; fa.ring = fp.ring;
    fa.cnt = (unsigned*)h->flow_sync.p; fa.lag = fp.lag;
    fa.V = (T*)h->V.p; fa.Gpart = (T*)h->Gpart.p; fa.Dpad = fp.Dpad; fa.nchunk = fp.nchunk; fa.spc = fp.spc;
    fa.losspart = (double*)h->epi_part.p;
    fa.dbg = h->flow_debug;
    if (!mn) {
        fa.es.partial = nullptr; fa.es.WT = fp.WT; fa.es.N = N; fa.es.theta = theta; fa.es.bias_off = g.pf;
        fa.es.y = (const T*)y; fa.es.yhat = out; fa.es.V = (T*)h->V.p; fa.es.part = nullptr;
    } else {
        fa.em.partial = nullptr; fa.em.WT = fp.WT; fa.em.RKs = fp.RKs; fa.em.N = N; fa.em.R = g.R; fa.em.C = g.C;
        fa.em.w = w; fa.em.y = (const long long*)y; fa.em.dP_in = nullptr; fa.em.class_w = class_w;
        fa.em.P = out; fa.em.pred = nullptr; fa.em.V = (T*)h->V.p; fa.em.u_ws = (T*)h->u_ws.p;
        fa.em.dZ_ws = (T*)h->dZ_ws.p; fa.em.part = nullptr;
    }
    fa.em.FC = (const double*)h->Ft64.p + g.pfeat;           // read by the kernel prologue only when C > 0
    if (h->prof) { rc = prof_fold(h, false); if (rc) return rc; TR_CUDA(h, cudaEventRecord(h->ev[4], st)); }
    {
        // blocks wait on one another: a cooperative launch guarantees that the whole grid is resident
        if ((rc = raise_smem_limit(h, e->flow_vec, fp.smem))) return rc;
        void* kargs[] = {(void*)&fa};
        TR_CUDA(h, cudaLaunchCooperativeKernel((const void*)e->flow_vec, dim3((unsigned)fp.grid), dim3(TR_TPB), kargs,
                                               fp.smem, st));
    }
    TR_LAUNCH_CHECK(h);
    if (h->prof) { TR_CUDA(h, cudaEventRecord(h->ev[5], st)); h->ev_set[2] = true; }

    if (!mn) {
        k_colsum<<<2, 128, 0, st>>>((const double*)h->epi_part.p, nwarps, 2, gradsum + g.pf);
        TR_LAUNCH_CHECK(h);
    } else {
        k_colsum<<<1, 128, 0, st>>>((const double*)h->epi_part.p, nwarps, 2, gradsum + g.pf);
        TR_LAUNCH_CHECK(h);
        const int dgrid = (int)std::min<long long>((N + 63) / 64, (long long)h->sms * 2);
        if ((rc = launch_dfc<T>(h, dgrid, w, N, st))) return rc;
        k_colsum<<<g.C * g.R, 128, 0, st>>>((const double*)h->dfc_part.p, dgrid, g.C * g.R, gradsum + g.pfeat);
        TR_LAUNCH_CHECK(h);
    }
    const long long tot = (long long)fp.RKs * g.D;
    const int rgrid = (int)std::min<long long>((tot + 255) / 256, (long long)h->sms * 8);
    k_reduce_G<T><<<rgrid, 256, 0, st>>>((const T*)h->Gpart.p, fp.nchunk * fp.Gn, fp.RKs, g.D, fp.Dpad,
                                         (double*)h->Gred.p);
    TR_LAUNCH_CHECK(h);
    MtArgs ma;
    ma.G = (const double*)h->Gred.p; ma.Ft64 = (const double*)h->Ft64.p; ma.w = w;
    ma.w_is_f64 = sizeof(T) == 8; ma.per_rank = mn ? 1 : 0; ma.geo = g; ma.gradsum = gradsum;
    int rows = 0;
    for (int m = 0; m < g.k; ++m) rows += g.dims[m];
    k_mttkrp<<<rows, TR_TPB, 0, st>>>(ma);
    TR_LAUNCH_CHECK(h);
    h->info[0] = h->launches; h->info[1] = fp.grid; h->info[2] = 0; h->info[3] = fp.lag;
    h->info[4] = fp.Gn; h->info[5] = fp.nchunk; h->info[6] = fp.RKs; h->info[7] = -(16 / (int)sizeof(T));
    h->last_fused = 2;
    return TR_OK;
}

#endif  // TR_WITH_FLOW

void set_info(tr_handle* h, const Plan& pl) {
    h->info[0] = h->launches; h->info[1] = pl.grid_f; h->info[2] = pl.grid_g; h->info[3] = pl.WT;
    h->info[4] = pl.Gn_f; h->info[5] = pl.Gn_g; h->info[6] = pl.RKs; h->info[7] = pl.vec ? (int)(16 / h->elt) : 1;
}

template <typename T>
int forward_std_t(tr_handle* h, const void* X, long long N, const void* theta, const void* w, uint32_t nn_mask,
                  double beta, double thr, void* yhat, cudaStream_t st) {
    Plan pl; const KEntry<T>* e; int rc;
    if ((rc = make_plan<T>(h, N, 1, vec_ok(X, h->geo.D, sizeof(T)), &pl, &e))) return rc;
    if ((rc = reserve_for<T>(h, N, pl))) return rc;
    h->launches = 0;
    if ((rc = run_forward<T>(h, (const T*)X, N, (const T*)theta, (const T*)w, nn_mask, beta, thr, pl, e, st))) return rc;
    EpiStdArgs<T> ea;
    ea.partial = (const T*)h->partial.p; ea.WT = pl.WT; ea.N = N; ea.theta = (const T*)theta;
    ea.bias_off = h->geo.pf; ea.y = nullptr; ea.yhat = (T*)yhat; ea.V = nullptr; ea.part = nullptr;
    const int egrid = (int)std::min<long long>((N + TR_WPB - 1) / TR_WPB, (long long)h->sms * 8);
    k_epi_std<T><<<egrid, TR_TPB, 0, st>>>(ea);
    TR_LAUNCH_CHECK(h);
    set_info(h, pl);
    return TR_OK;
}

template <typename T>
int fwd_grad_std_t(tr_handle* h, const void* X, const void* y, long long N, const void* theta, const void* w,
                   uint32_t nn_mask, double beta, double thr, double* gradsum, void* yhat, cudaStream_t st,
                   bool backward_only) {
    Plan pl; const KEntry<T>* e; int rc;
    const Geo& g = h->geo;
    h->last_fused = 0;
    if (!backward_only && h->fused_mode != 0) {
        FusedPlan fp;
        if ((rc = plan_fused<T>(h, N, X, &fp))) return rc;
        // auto: the single-pass kernel pays once every cluster has a few samples to pipeline AND X is well beyond
        // L2: up to ~3 x L2 the second pass of the two-pass kernels is served largely from L2 and beats the cluster
        // kernel's prologue (measured on (N, 20,30,40): crossover between 192 MB and 768 MB of X)
        const bool big = (size_t)N * (size_t)g.D * sizeof(T) >= 3 * h->l2_bytes;
        const bool want = fp.CL > 0 && (h->fused_mode == 1 || (N >= 8LL * fp.NC && big));
        if (h->fused_mode == 1 && fp.CL == 0)
            return fail(h, TR_ERR_UNSUPPORTED, "fused=1 requested but this geometry / alignment is not eligible for the single-pass kernel");
        if (want) {
            h->launches = 0;
            h->last_fused = 1;
            return run_fused_std<T>(h, (const T*)X, (const T*)y, N, (const T*)theta, (const T*)w, nn_mask, beta, thr,
                                    fp, gradsum, (T*)yhat, st);
        }
    }
#ifdef TR_WITH_FLOW
    if (!backward_only && h->flow_mode == 1) {
        FlowPlan fl; const KEntry<T>* fe = nullptr;
        if ((rc = plan_flow<T>(h, N, 1, X, &fl, &fe))) return rc;
        if (h->flow_mode == 1 && !fl.ok)
            return fail(h, TR_ERR_UNSUPPORTED, "flow=1 requested but this geometry / alignment is not eligible for the dataflow kernel");
        if (fl.ok) {
            h->launches = 0;
            return run_flow<T>(h, (const T*)X, y, (const T*)nullptr, N, (const T*)theta, (const T*)w, nn_mask, beta, thr,
                               fl, fe, gradsum, (T*)yhat, st);
        }
    }
#endif
    if ((rc = make_plan<T>(h, N, 1, vec_ok(X, g.D, sizeof(T)), &pl, &e))) return rc;
    if ((rc = reserve_for<T>(h, N, pl))) return rc;
    h->launches = 0;
    const int egrid = (int)std::min<long long>((N + TR_WPB - 1) / TR_WPB, (long long)h->sms * 8);
    const T* V;
    if (!backward_only) {
        if ((rc = run_forward<T>(h, (const T*)X, N, (const T*)theta, (const T*)w, nn_mask, beta, thr, pl, e, st))) return rc;
        EpiStdArgs<T> ea;
        ea.partial = (const T*)h->partial.p; ea.WT = pl.WT; ea.N = N; ea.theta = (const T*)theta;
        ea.bias_off = g.pf; ea.y = (const T*)y; ea.yhat = (T*)yhat; ea.V = (T*)h->V.p; ea.part = (double*)h->epi_part.p;
        k_epi_std<T><<<egrid, TR_TPB, 0, st>>>(ea);
        TR_LAUNCH_CHECK(h);
        // gradsum[pf] = sum res, gradsum[pf+1] = sum res^2
        k_colsum<<<2, 128, 0, st>>>((const double*)h->epi_part.p, egrid, 2, gradsum + g.pf);
        TR_LAUNCH_CHECK(h);
        V = (const T*)h->V.p;
    } else {
        // y carries the upstream gradient dyhat; prep is still needed for the MTTKRP factors
        k_prep<T><<<std::max(1, std::min(64, (g.pf + 255) / 256)), 256, 0, st>>>((const T*)theta, g, nn_mask, beta, thr,
                                                                               (T*)h->FtT.p, (double*)h->Ft64.p);
        TR_LAUNCH_CHECK(h);
        V = (const T*)y;
        k_vecsum<T><<<1, 1024, 0, st>>>(V, N, gradsum + g.pf);
        TR_LAUNCH_CHECK(h);
        TR_CUDA(h, cudaMemsetAsync(gradsum + g.pf + 1, 0, sizeof(double), st));
    }
    if ((rc = run_gradient<T>(h, (const T*)X, N, V, (const T*)w, pl, e, gradsum, st))) return rc;
    set_info(h, pl);
    return TR_OK;
}

template <typename T>
int mn_t(tr_handle* h, const void* X, const long long* y, const void* class_w, long long N, const void* theta,
         const void* w, uint32_t nn_mask, double beta, double thr, double* gradsum, void* P, long long* pred,
         cudaStream_t st, const void* dP_in = nullptr) {
    Plan pl; const KEntry<T>* e; int rc;
    const Geo& g = h->geo;
    h->last_fused = 0;
    if (y != nullptr && dP_in == nullptr && pred == nullptr && h->fused_mode != 0) {
        FusedMnPlan fp;
        if ((rc = plan_fused_mn<T>(h, N, X, &fp))) return rc;
        if (h->fused_mode == 1 && !fp.ok)
            return fail(h, TR_ERR_UNSUPPORTED, "fused=1 requested but this geometry / alignment is not eligible for the single-pass multinomial kernel");
        // auto: as for the standard model — every cluster has samples to pipeline and X is well beyond L2; rows whose
        // 16-byte chunk count is even make the row-per-thread shared-memory reads bank-conflicted: two-pass is kept
        const bool big = (size_t)N * (size_t)g.D * sizeof(T) >= 3 * h->l2_bytes;
        // row pitch in 16-byte chunks: odd = conflict-free row-per-thread reads; 6 = two-way conflicts, still 15 % ahead of the
        // two-pass kernels (4.5 against 5.3 ms on 16 GB of 100 x 50 x 24); 2 and 4 lose to them (tools/mn_even_pitch.py)
        const bool pitch_ok = (fp.IKC & 1) || fp.IKC == 6;
        if (fp.ok && (h->fused_mode == 1 || (N >= 8LL * fp.NC && big && pitch_ok))) {
            h->launches = 0;
            return run_fused_mn<T>(h, (const T*)X, y, (const T*)class_w, N, (const T*)theta, (const T*)w, nn_mask, beta, thr,
                                   fp, gradsum, (T*)P, st);
        }
    }
#ifdef TR_WITH_FLOW
    if (y != nullptr && dP_in == nullptr && pred == nullptr && h->flow_mode == 1) {
        FlowPlan fl; const KEntry<T>* fe = nullptr;
        if ((rc = plan_flow<T>(h, N, g.R, X, &fl, &fe))) return rc;
        if (h->flow_mode == 1 && !fl.ok)
            return fail(h, TR_ERR_UNSUPPORTED, "flow=1 requested but this geometry / alignment is not eligible for the dataflow kernel");
        if (fl.ok) {
            h->launches = 0;
            return run_flow<T>(h, (const T*)X, y, (const T*)class_w, N, (const T*)theta, (const T*)w, nn_mask, beta, thr,
                               fl, fe, gradsum, (T*)P, st);
        }
    }
#endif
    if ((rc = make_plan<T>(h, N, g.R, vec_ok(X, g.D, sizeof(T)), &pl, &e))) return rc;
    if ((rc = reserve_for<T>(h, N, pl))) return rc;
    h->launches = 0;
    if ((rc = run_forward<T>(h, (const T*)X, N, (const T*)theta, (const T*)w, nn_mask, beta, thr, pl, e, st))) return rc;
    const bool train = (y != nullptr) || (dP_in != nullptr);
    EpiMnArgs<T> ea;
    ea.partial = (const T*)h->partial.p; ea.WT = pl.WT; ea.RKs = pl.RKs; ea.N = N; ea.R = g.R; ea.C = g.C;
    ea.FC = (const double*)h->Ft64.p + g.pfeat; ea.w = (const T*)w; ea.y = y; ea.class_w = (const T*)class_w;
    ea.dP_in = (const T*)dP_in;
    ea.P = (T*)P; ea.pred = pred;
    ea.V = train ? (T*)h->V.p : nullptr; ea.u_ws = train ? (T*)h->u_ws.p : nullptr;
    ea.dZ_ws = train ? (T*)h->dZ_ws.p : nullptr; ea.part = train ? (double*)h->epi_part.p : nullptr;
    const int egrid = (int)std::min<long long>((N + TR_WPB - 1) / TR_WPB, (long long)h->sms * 8);
    const size_t esmem = (size_t)(g.C * g.R + g.R) * sizeof(double);
    // loop bounds of the epilogue sized to the model (same arithmetic, fewer dead iterations and registers)
    if (g.C <= 32 && g.R <= 8) k_epi_mn<T, 8, 1><<<egrid, TR_TPB, esmem, st>>>(ea);
    else if (g.C <= 32 && g.R <= 16) k_epi_mn<T, 16, 1><<<egrid, TR_TPB, esmem, st>>>(ea);
    else if (g.R <= 16) k_epi_mn<T, 16, TR_JC><<<egrid, TR_TPB, esmem, st>>>(ea);
    else if (g.C <= 32) k_epi_mn<T, TR_MAX_RANK_MN, 1><<<egrid, TR_TPB, esmem, st>>>(ea);
    else k_epi_mn<T, TR_MAX_RANK_MN, TR_JC><<<egrid, TR_TPB, esmem, st>>>(ea);
    TR_LAUNCH_CHECK(h);
    if (train) {
        k_colsum<<<1, 128, 0, st>>>((const double*)h->epi_part.p, egrid, 1, gradsum + g.pf);
        TR_LAUNCH_CHECK(h);
        const int dgrid = (int)std::min<long long>((N + 63) / 64, (long long)h->sms * 2);
        if ((rc = launch_dfc<T>(h, dgrid, (const T*)w, N, st))) return rc;
        k_colsum<<<g.C * g.R, 128, 0, st>>>((const double*)h->dfc_part.p, dgrid, g.C * g.R,
                                                          gradsum + g.pfeat);
        TR_LAUNCH_CHECK(h);
        if ((rc = run_gradient<T>(h, (const T*)X, N, (const T*)h->V.p, (const T*)w, pl, e, gradsum, st))) return rc;
    }
    set_info(h, pl);
    return TR_OK;
}


// ---------------------------------------------------------------------------------------------
// spectral model (tr_spectral.cuh): plan + launch
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC> using SpecFwdKern = void (*)(SpecFwdArgs<T>);
template <typename T, int VEC> using SpecGradKern = void (*)(SpecGradArgs<T>);

template <typename T, int VEC>
SpecFwdKern<T, VEC> spec_fwd_kernel(int QT) {
    switch (QT) {
        case 1: return k_spec_fwd<T, 1, VEC, 8>;
        case 2: return k_spec_fwd<T, 2, VEC, 8>;
        case 3: return k_spec_fwd<T, 3, VEC, 4>;
        case 4: return k_spec_fwd<T, 4, VEC, 4>;
        case 5: return k_spec_fwd<T, 5, VEC, 4>;
        case 6: return k_spec_fwd<T, 6, VEC, 4>;
        case 7: return k_spec_fwd<T, 7, VEC, 4>;
        case 8: return k_spec_fwd<T, 8, VEC, 4>;
    }
    return nullptr;
}
template <typename T, int VEC>
SpecGradKern<T, VEC> spec_grad_kernel(int QT) {
    switch (QT) {
        case 1: return k_spec_grad<T, 1, VEC>;
        case 2: return k_spec_grad<T, 2, VEC>;
        case 3: return k_spec_grad<T, 3, VEC>;
        case 4: return k_spec_grad<T, 4, VEC>;
        case 5: return k_spec_grad<T, 5, VEC>;
        case 6: return k_spec_grad<T, 6, VEC>;
        case 7: return k_spec_grad<T, 7, VEC>;
        case 8: return k_spec_grad<T, 8, VEC>;
    }
    return nullptr;
}

template <typename T, int VEC> using SpecFusedKern = void (*)(SpecFusedArgs<T>);
template <typename T, int VEC>
SpecFusedKern<T, VEC> spec_fused_kernel(int QT) {
    switch (QT) {
        case 1: return k_spec_fused<T, 1, VEC, 8>;
        case 2: return k_spec_fused<T, 2, VEC, 8>;
        case 3: return k_spec_fused<T, 3, VEC, 8>;
        case 4: return k_spec_fused<T, 4, VEC, 8>;
        case 5: return k_spec_fused<T, 5, VEC, 8>;
        case 6: return k_spec_fused<T, 6, VEC, 8>;
        case 7: return k_spec_fused<T, 7, VEC, 4>;
        case 8: return k_spec_fused<T, 8, VEC, 4>;
    }
    return nullptr;
}

// channels are processed in groups of at most TRS_MAXQ per pass over X, the groups as equal as possible
static inline void spec_groups(int Q, int* ngroups, int* qt) {
    *ngroups = (Q + TRS_MAXQ - 1) / TRS_MAXQ;
    *qt = (Q + *ngroups - 1) / *ngroups;
}

template <typename T, int VEC>
int spec_pass1(tr_handle* h, const T* X, long long N, cudaStream_t st) {
    const SpecGeo& sg = h->sg;
    int ng, QT;
    spec_groups(sg.Q, &ng, &QT);
    auto kern = spec_fwd_kernel<T, VEC>(QT);
    const size_t smem = (size_t)sg.W * QT * sizeof(T);
    if (smem > 200 * 1024) return fail(h, TR_ERR_UNSUPPORTED, "first-mode factor rows (%zu bytes) do not fit in shared memory", smem);
    int rc, occ = 0;
    if ((rc = occupancy(h, kern, smem, &occ))) return rc;
    const long long DT = (sg.D + 32 * VEC - 1) / (32 * VEC);
    const long long items = N * DT;
    const int grid = (int)std::max<long long>(1, std::min<long long>((items + TR_WPB - 1) / TR_WPB, (long long)h->sms * occ));
    for (int gi = 0; gi < ng; ++gi) {
        SpecFwdArgs<T> fa;
        fa.X = X; fa.N = N; fa.FtT = (const T*)h->FtT.p; fa.A = (T*)h->spA.p; fa.g = sg; fa.q0 = gi * QT;
        if ((rc = raise_smem_limit(h, kern, smem))) return rc;
        if (h->prof && gi == 0) { rc = prof_fold(h, false); if (rc) return rc; TR_CUDA(h, cudaEventRecord(h->ev[0], st)); }
        kern<<<grid, TR_TPB, smem, st>>>(fa);
        TR_LAUNCH_CHECK(h);
        if (h->prof && gi == ng - 1) { TR_CUDA(h, cudaEventRecord(h->ev[1], st)); h->ev_set[0] = true; }
    }
    h->info[1] = grid; h->info[3] = (long long)DT; h->info[6] = QT; h->info[7] = VEC;
    return TR_OK;
}

template <typename T, int VEC>
int spec_pass2(tr_handle* h, const T* X, long long N, double* gradsum, cudaStream_t st) {
    const SpecGeo& sg = h->sg;
    int ng, QT;
    spec_groups(sg.Q, &ng, &QT);
    auto kern = spec_grad_kernel<T, VEC>(QT);
    int rc, occ = 0;
    if ((rc = occupancy(h, kern, 0, &occ))) return rc;
    const int WTN = (sg.W + TRS_WT - 1) / TRS_WT;
    const long long wtot = (long long)h->sms * occ * TR_WPB;
    long long Gn = std::max<long long>(1, wtot / WTN);
    if (Gn > N) Gn = N;
    const long long items = (long long)WTN * Gn;
    const int grid = (int)std::max<long long>(1, std::min<long long>((items + TR_WPB - 1) / TR_WPB, (long long)h->sms * occ));
    if ((rc = ensure(h, h->spPart, (size_t)items * TRS_WT * QT * sizeof(double)))) return rc;
    const long long per = (N + Gn - 1) / Gn;
    const long long DT = (sg.D + 32 * VEC - 1) / (32 * VEC);
    // at most ~2048 (sample, tile) steps between two folds of the fp32 sums into the double slots
    const long long spc = sizeof(T) == 4 ? std::max<long long>(1, std::min<long long>(per, 2048 / std::max<long long>(1, DT))) : per;
    for (int gi = 0; gi < ng; ++gi) {
        SpecGradArgs<T> ga;
        ga.X = X; ga.DA = (const T*)h->spDA.p; ga.N = N; ga.g = sg; ga.q0 = gi * QT; ga.WTN = WTN; ga.Gn = (int)Gn;
        ga.spc = spc; ga.part = (double*)h->spPart.p;
        if (h->prof && gi == 0) { rc = prof_fold(h, false); if (rc) return rc; TR_CUDA(h, cudaEventRecord(h->ev[2], st)); }
        kern<<<grid, TR_TPB, 0, st>>>(ga);
        TR_LAUNCH_CHECK(h);
        if (h->prof && gi == ng - 1) { TR_CUDA(h, cudaEventRecord(h->ev[3], st)); h->ev_set[1] = true; }
        k_spec_dg_fold<<<sg.W * QT, 128, 0, st>>>((const double*)h->spPart.p, WTN, (int)Gn, QT, gi * QT, sg, gradsum);
        TR_LAUNCH_CHECK(h);
    }
    h->info[2] = grid; h->info[5] = Gn;
    return TR_OK;
}

// pass 1 + epilogue + second-mode gradient in one kernel (samples of one warp tile); fills DA, res, U, gradsum[F*1], loss parts
template <typename T, int VEC>
int spec_fused_pass1(tr_handle* h, const T* X, const T* y, long long N, const T* theta, const T* w, double nb,
                     double* gradsum, T* yhat, int* nloss, cudaStream_t st) {
    const SpecGeo& sg = h->sg;
    const int QT = sg.Q;
    auto kern = spec_fused_kernel<T, VEC>(QT);
    constexpr int CH = 16 / (int)sizeof(T);
    const int QP = (QT + CH - 1) / CH * CH;                                   // row stride of the G table (whole 16-byte chunks)
    const size_t smem = ((((size_t)sg.W * QP + (size_t)sg.NO * QT + sg.NO) * sizeof(T) + 15) / 16) * 16
                        + (size_t)TR_WPB * QT * 32 * VEC * sizeof(T);                 // tables + the warps' scratch
    int rc, occ = 0;
    if ((rc = occupancy(h, kern, smem, &occ))) return rc;
    const int grid = (int)std::max<long long>(1, std::min<long long>((N + TR_WPB - 1) / TR_WPB, (long long)h->sms * occ));
    const size_t nslots = (size_t)grid * TR_WPB;
    const int TILE = 32 * VEC;
    if ((rc = ensure(h, h->spDf1, nslots * QT * TILE * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->epi_part, (size_t)grid * sizeof(double)))) return rc;
    SpecFusedArgs<T> fa;
    fa.X = X; fa.y = y; fa.N = N; fa.FtT = (const T*)h->FtT.p; fa.theta = theta; fa.w = w; fa.g = sg; fa.nb = nb;
    fa.DA = (T*)h->spDA.p; fa.res = (T*)h->spRes.p; fa.U = (T*)h->spU.p; fa.yhat = yhat;
    fa.df1part = (double*)h->spDf1.p; fa.losspart = (double*)h->epi_part.p;
    const long long per = (N + (long long)nslots - 1) / (long long)nslots;
    fa.spc = sizeof(T) == 4 ? std::max<long long>(1, std::min<long long>(per, 2048)) : std::max<long long>(1, per);
    if ((rc = raise_smem_limit(h, kern, smem))) return rc;
    if (h->prof) { rc = prof_fold(h, false); if (rc) return rc; TR_CUDA(h, cudaEventRecord(h->ev[0], st)); }
    kern<<<grid, TR_TPB, smem, st>>>(fa);
    TR_LAUNCH_CHECK(h);
    if (h->prof) { TR_CUDA(h, cudaEventRecord(h->ev[1], st)); h->ev_set[0] = true; }
    k_spec_df1_fold<<<sg.RT * sg.D, 128, 0, st>>>((const double*)h->spDf1.p, (int)nslots, QT, TILE, sg, gradsum);
    TR_LAUNCH_CHECK(h);
    *nloss = grid;
    h->info[1] = grid; h->info[3] = 1; h->info[6] = QT; h->info[7] = VEC;
    return TR_OK;
}

// ---- single-pass spectral kernel (tr_spectral_single.cuh): plan + launch ----
// first-mode table G[w][q] in the layout the kernel keeps in shared memory: rows of `stride` values (TrssG)
template <typename T>
__global__ void k_spec_pack_g(const T* __restrict__ FtT, SpecGeo g, int QT, int stride, T* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.W * stride; i += gridDim.x * blockDim.x) {
        const int w = i / stride, e = i % stride;
        const int q = trss_g_twice<T>() ? e / 2 : e;
        out[i] = (q < QT && q < g.Q) ? spec_G(FtT, g, w, q) : (T)0;
    }
}

struct SpecSinglePlan { bool ok; int NS; size_t smem; unsigned stage_bytes; const void* kern; };

template <typename T>
int spec_single_plan(tr_handle* h, const void* X, SpecSinglePlan* sp) {
    const SpecGeo& sg = h->sg;
    constexpr int VEC = 16 / (int)sizeof(T);
    sp->ok = false; sp->NS = 0; sp->smem = 0; sp->kern = nullptr;
    if (!vec_ok(X, sg.D, sizeof(T)) || sg.Q > TRS_MAXQ || sg.D > 32 * VEC || sg.W > TRSS_NG * TRS_WT) return TR_OK;
    const size_t stage = (size_t)sg.W * sg.D * sizeof(T);
    if (stage % 16 != 0 || stage > (1u << 20)) return TR_OK;
    if (h->smem_optin == 0) TR_CUDA(h, cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    int NS = TRSS_MAX_NS;
    if (h->spec_single_ns > 0) NS = std::min(NS, h->spec_single_ns);
    while (NS >= 2 && spec_single_layout<T>(sg, sg.Q, VEC, NS, stage).total + 1024 > (size_t)h->smem_optin) --NS;
    if (NS < 2) return TR_OK;
    sp->kern = sizeof(T) == 4 ? trss_kernel_f32(sg.Q) : trss_kernel_f64(sg.Q);
    if (!sp->kern) return TR_OK;
    sp->ok = true; sp->NS = NS; sp->stage_bytes = (unsigned)stage;
    sp->smem = spec_single_layout<T>(sg, sg.Q, VEC, NS, stage).total;
    return TR_OK;
}

// forward + epilogue + both factor gradients of the spectral model with X read once; fills res, U, gradsum[F*0 | F*1], loss parts
template <typename T>
int spec_single_pass(tr_handle* h, const SpecSinglePlan& sp, const T* X, const T* y, long long N, const T* theta, const T* w,
                     double nb, double* gradsum, T* yhat, int* nloss, cudaStream_t st) {
    const SpecGeo& sg = h->sg;
    constexpr int VEC = 16 / (int)sizeof(T);
    const int QT = sg.Q, TILE = 32 * VEC;
    const int grid = (int)std::max<long long>(1, std::min<long long>(N, (long long)h->sms));
    const int WTN = (sg.W + TRS_WT - 1) / TRS_WT;
    const size_t nslots = (size_t)grid * TRSS_NF;
    int rc;
    if ((rc = ensure(h, h->spDf1, nslots * QT * TILE * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->spPart, (size_t)grid * WTN * TRS_WT * QT * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->epi_part, (size_t)grid * sizeof(double)))) return rc;
    SpecSingleArgs<T> fa;
    fa.X = X; fa.y = y; fa.N = N; fa.FtT = (const T*)h->FtT.p; fa.theta = theta; fa.w = w; fa.g = sg; fa.nb = nb;
    fa.res = (T*)h->spRes.p; fa.U = (T*)h->spU.p; fa.yhat = yhat;
    fa.df1part = (double*)h->spDf1.p; fa.dgpart = (double*)h->spPart.p; fa.losspart = (double*)h->epi_part.p;
    const long long per = (N + grid - 1) / grid;
    fa.spc = sizeof(T) == 4 ? std::max<long long>(1, std::min<long long>(per, 2048)) : std::max<long long>(1, per);
    fa.NS = sp.NS; fa.stage_bytes = sp.stage_bytes;
    fa.piece = (unsigned)std::max(16, h->fused_piece / 16 * 16);
    fa.trace = nullptr;
#ifdef TRSS_TRACE
    if ((rc = ensure(h, h->trace, (size_t)TRSS_TRACE_N * TRSS_TRACE_EV * sizeof(long long)))) return rc;
    TR_CUDA(h, cudaMemsetAsync(h->trace.p, 0, (size_t)TRSS_TRACE_N * TRSS_TRACE_EV * sizeof(long long), st));
    fa.trace = (long long*)h->trace.p;
#endif
    if ((rc = raise_smem_limit(h, sp.kern, sp.smem))) return rc;
    {
        // first-mode table in the kernel's layout (every block copies it into its shared memory)
        const int stride = trss_g_stride<T>(QT);
        if ((rc = ensure(h, h->f3_stage, (size_t)sg.W * stride * sizeof(T)))) return rc;
        k_spec_pack_g<T><<<4, 256, 0, st>>>((const T*)h->FtT.p, sg, QT, stride, (T*)h->f3_stage.p);
        TR_LAUNCH_CHECK(h);
        fa.gtab = (const T*)h->f3_stage.p;
        if (h->prof) { rc = prof_fold(h, false); if (rc) return rc; TR_CUDA(h, cudaEventRecord(h->ev[4], st)); }
        void* kargs[] = {(void*)&fa};
        TR_CUDA(h, cudaLaunchKernel(sp.kern, dim3(grid), dim3(TRSS_NT), kargs, sp.smem, st));
        TR_LAUNCH_CHECK(h);
        if (h->prof) { TR_CUDA(h, cudaEventRecord(h->ev[5], st)); h->ev_set[2] = true; }
    }
    k_spec_df1_fold<<<sg.RT * sg.D, 128, 0, st>>>((const double*)h->spDf1.p, (int)nslots, QT, TILE, sg, gradsum);
    TR_LAUNCH_CHECK(h);
    k_spec_dg_fold<<<sg.W * QT, 128, 0, st>>>((const double*)h->spPart.p, WTN, grid, QT, 0, sg, gradsum);
    TR_LAUNCH_CHECK(h);
    *nloss = grid;
    h->info[1] = grid; h->info[2] = grid; h->info[3] = 1; h->info[5] = sp.NS; h->info[6] = QT; h->info[7] = VEC;
    return TR_OK;
}

// forward (+ gradient when y and gradsum are given) of the spectral model
template <typename T>
int run_spec(tr_handle* h, const T* X, const T* y, long long N, const T* theta, const T* w, uint32_t nn_mask, double beta,
             double thr, double* gradsum, T* yhat, T* yhat_lin, T* spec_pred, T* latents, cudaStream_t st) {
    const SpecGeo& sg = h->sg;
    const Geo& g = h->geo;
    const bool grad = gradsum != nullptr;
    int rc;
    h->launches = 0;
    if ((rc = ensure(h, h->FtT, (size_t)g.pf * sizeof(T)))) return rc;
    if ((rc = ensure(h, h->Ft64, (size_t)g.pf * sizeof(double)))) return rc;
    const int egrid = (int)std::max<long long>(1, std::min<long long>((N + TR_WPB - 1) / TR_WPB, (long long)h->sms * 8));
    const int dgrid = (int)std::max<long long>(1, std::min<long long>(N / 64 + 1, (long long)h->sms * 2));
    const int CR = sg.NO * (sg.RT + 1);
    const int slabs = (int)std::max<long long>(1, std::min<long long>(N / 32 + 1, std::max<long long>(1, (long long)h->sms * 4 / ((sg.D + TR_TPB - 1) / TR_TPB))));
    if (grad) {
        if ((rc = ensure(h, h->spU, (size_t)N * (sg.RT + 1) * sizeof(T) + (size_t)(sg.RT + 1) * sizeof(T)))) return rc;
        if ((rc = ensure(h, h->spRes, (size_t)N * sg.NO * sizeof(T)))) return rc;
        if ((rc = ensure(h, h->epi_part, (size_t)egrid * sizeof(double)))) return rc;
        if ((rc = ensure(h, h->dfc_part, ((size_t)dgrid + 1) * CR * sizeof(double)))) return rc;
    }
    k_prep<T><<<std::max(1, std::min(64, (g.pf + 255) / 256)), 256, 0, st>>>(theta, g, nn_mask, beta, thr, (T*)h->FtT.p, (double*)h->Ft64.p);
    TR_LAUNCH_CHECK(h);
    const bool vec = vec_ok(X, sg.D, sizeof(T));
    constexpr int VEC = 16 / (int)sizeof(T);
    const double nbias_mult = (double)((sg.Rn > 0 ? 1 : 0) + (sg.Rs > 0 ? 1 : 0));
    // fit iterations of samples that fit one warp tile: window contraction, epilogue and second-mode gradient in ONE kernel
    const bool can_fuse = grad && sg.Q <= TRS_MAXQ && sg.D <= 32 * (vec ? VEC : 1);
    if (h->fused_mode == 1 && grad && !can_fuse)
        return fail(h, TR_ERR_UNSUPPORTED, "option fused=1: the fused spectral pass needs Q <= %d channels and D <= %d features", TRS_MAXQ, 32 * (vec ? VEC : 1));
    // ... and when a ring of whole samples fits the shared memory of an SM, the first-mode gradient as well: X is read once
    SpecSinglePlan sp;
    sp.ok = false;
    if (grad && can_fuse && h->spec_single != 0) { if ((rc = spec_single_plan<T>(h, X, &sp))) return rc; }
    if (h->spec_single == 1 && grad && !sp.ok)
        return fail(h, TR_ERR_UNSUPPORTED, "option spec_single=1: the single-pass spectral kernel needs 16-byte rows, Q <= %d channels, D <= %d features, "
                    "W <= %d window rows and two whole samples in shared memory", TRS_MAXQ, 32 * VEC, TRSS_NG * TRS_WT);
    const bool single = sp.ok && (h->spec_single == 1 || (h->fused_mode != 0 && N >= 4LL * h->sms));
    const bool fuse = single || (can_fuse && h->fused_mode != 0);
    if (!fuse) {
        if ((rc = ensure(h, h->spA, (size_t)N * sg.Q * sg.D * sizeof(T)))) return rc;
        if (grad) {
            if ((rc = ensure(h, h->spMc, (size_t)N * std::max(1, sg.Rs) * sg.D * sizeof(T)))) return rc;
            if ((rc = ensure(h, h->spDS, (size_t)N * sg.RT * sizeof(T)))) return rc;
            if ((rc = ensure(h, h->spDf1, (size_t)slabs * sg.RT * sg.D * sizeof(double)))) return rc;
        }
    }
    if (grad && !single) {
        if ((rc = ensure(h, h->spDA, (size_t)N * sg.Q * sg.D * sizeof(T)))) return rc;
    }
    int nloss = egrid;
    if (single) {
        if ((rc = spec_single_pass<T>(h, sp, X, y, N, theta, w, nbias_mult, gradsum, yhat, &nloss, st))) return rc;
    } else if (fuse) {
        if ((rc = vec ? spec_fused_pass1<T, VEC>(h, X, y, N, theta, w, nbias_mult, gradsum, yhat, &nloss, st)
                      : spec_fused_pass1<T, 1>(h, X, y, N, theta, w, nbias_mult, gradsum, yhat, &nloss, st))) return rc;
    } else {
        if ((rc = vec ? spec_pass1<T, VEC>(h, X, N, st) : spec_pass1<T, 1>(h, X, N, st))) return rc;
    }
    SpecEpiArgs<T> ea;
    memset(&ea, 0, sizeof(ea));
    ea.A = (const T*)h->spA.p; ea.Ft64 = (const double*)h->Ft64.p; ea.theta = theta; ea.w = w; ea.y = y; ea.N = N; ea.g = sg;
    ea.nb = nbias_mult;
    ea.yhat = yhat;
    if (grad) {
        ea.res = (T*)h->spRes.p; ea.U = (T*)h->spU.p; ea.dS = (T*)h->spDS.p; ea.Mc = (T*)h->spMc.p; ea.DA = (T*)h->spDA.p;
        ea.part = (double*)h->epi_part.p;
    }
    if (!fuse && (grad || yhat)) {
        k_spec_epi<T><<<egrid, TR_TPB, 0, st>>>(ea);
        TR_LAUNCH_CHECK(h);
    }
    if (yhat_lin || spec_pred || latents) {
        SpecPredArgs<T> pa;
        pa.A = (const T*)h->spA.p; pa.Ft64 = (const double*)h->Ft64.p; pa.theta = theta; pa.w = w; pa.N = N; pa.g = sg;
        pa.yhat_lin = yhat_lin; pa.spec_pred = spec_pred; pa.latents = latents;
        k_spec_pred<T><<<egrid, TR_TPB, (size_t)TR_WPB * sg.Q * sizeof(double), st>>>(pa);
        TR_LAUNCH_CHECK(h);
    }
    if (grad) {
        // third-mode factors + bias:  M[n,c] = wcat_c sum_t res[t,n] U[t,c],  wcat = [w_normal | 1 ... 1 | 1]
        T* wcat = (T*)h->spU.p + (size_t)N * (sg.RT + 1);
        k_spec_wcat<T><<<1, 32, 0, st>>>(w, sg.Rn, sg.RT + 1, wcat);
        TR_LAUNCH_CHECK(h);
        const size_t smem = (size_t)4 * CR * sizeof(double);
        if ((rc = raise_smem_limit(h, k_dfc<T>, smem))) return rc;
        k_dfc<T><<<dgrid, TR_TPB, smem, st>>>((const T*)h->spRes.p, (const T*)h->spU.p, wcat, N, sg.NO, sg.RT + 1, (double*)h->dfc_part.p);
        TR_LAUNCH_CHECK(h);
        double* M = (double*)h->dfc_part.p + (size_t)dgrid * CR;
        k_colsum<<<CR, 128, 0, st>>>((const double*)h->dfc_part.p, dgrid, CR, M);
        TR_LAUNCH_CHECK(h);
        k_spec_scatter<<<1, 256, 0, st>>>(M, (const double*)h->epi_part.p, nloss, sg, ea.nb, gradsum);
        TR_LAUNCH_CHECK(h);
        if (!fuse) {
            // second-mode factors
            SpecDf1Args<T> da;
            da.A = (const T*)h->spA.p; da.Mc = (const T*)h->spMc.p; da.dS = (const T*)h->spDS.p; da.N = N; da.g = sg; da.slabs = slabs;
            da.part = (double*)h->spDf1.p;
            k_spec_df1<T><<<dim3((sg.D + TR_TPB - 1) / TR_TPB, slabs), TR_TPB, 0, st>>>(da);
            TR_LAUNCH_CHECK(h);
            k_spec_df1_fold<<<sg.RT * sg.D, 128, 0, st>>>((const double*)h->spDf1.p, slabs, sg.RT, sg.D, sg, gradsum);
            TR_LAUNCH_CHECK(h);
        }
        // first-mode factors: the second pass over X
        if (!single) {
            if ((rc = vec ? spec_pass2<T, VEC>(h, X, N, gradsum, st) : spec_pass2<T, 1>(h, X, N, gradsum, st))) return rc;
        }
    }
    h->info[0] = h->launches; h->info[4] = single ? -1 : (fuse ? 0 : slabs);
    return TR_OK;
}

int check_common(tr_handle* h, const void* X, long long N, const void* theta, const void* w) {
    if (!h) return TR_ERR_INVALID;
    if (h->kind != 0) return fail(h, TR_ERR_INVALID, "this entry point does not take a spectral handle (use tr_spec_*)");
    if (N < 0) return fail(h, TR_ERR_INVALID, "N must be >= 0 (got %lld)", N);
    if ((N > 0 && !X) || !theta || !w) return fail(h, TR_ERR_INVALID, "null X / theta / w pointer");
    if ((uintptr_t)X % h->elt != 0) return fail(h, TR_ERR_INVALID, "X is not aligned to its element size");
    return TR_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int tr_version(void) { return TR_B200_VERSION; }

int tr_create(tr_handle** out, int dtype, int k, const int64_t* dims, int R, int C, int device) {
    if (!out) return fail(nullptr, TR_ERR_INVALID, "out is null");
    *out = nullptr;
    if (dtype != TR_F32 && dtype != TR_F64) return fail(nullptr, TR_ERR_INVALID, "dtype must be TR_F32 or TR_F64");
    if (k < 1 || k > TR_MAX_MODES) return fail(nullptr, TR_ERR_UNSUPPORTED, "k=%d feature modes (supported: 1..%d)", k, TR_MAX_MODES);
    if (R < 1) return fail(nullptr, TR_ERR_INVALID, "rank must be >= 1");
    if (C < 0 || C > TR_MAX_CLASSES) return fail(nullptr, TR_ERR_UNSUPPORTED, "n_classes=%d (supported: up to %d)", C, TR_MAX_CLASSES);
    if (C > 0 && R > TR_MAX_RANK_MN) return fail(nullptr, TR_ERR_UNSUPPORTED, "multinomial rank %d (supported: up to %d)", R, TR_MAX_RANK_MN);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, TR_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return fail(nullptr, TR_ERR_INVALID, "device %d out of range (have %d)", device, ndev);
    tr_handle* h = new tr_handle();
    h->dtype = dtype; h->device = device; h->elt = dtype == TR_F32 ? 4 : 8;
    Geo& g = h->geo;
    memset(&g, 0, sizeof(g));
    g.k = k; g.R = R; g.C = C; g.D = 1;
    long long sumI = 0;
    for (int m = 0; m < k; ++m) {
        if (dims[m] < 1 || dims[m] > (1LL << 30)) { delete h; return fail(nullptr, TR_ERR_INVALID, "dims[%d]=%lld invalid", m, (long long)dims[m]); }
        g.dims[m] = (int)dims[m];
        g.foff[m] = (int)(sumI * R);
        sumI += dims[m];
        g.D *= dims[m];
        if (g.D > (1LL << 31) - 4096) { delete h; return fail(nullptr, TR_ERR_UNSUPPORTED, "one sample has more than 2^31 elements"); }
    }
    if ((sumI + C) * R > (1LL << 30)) { delete h; return fail(nullptr, TR_ERR_UNSUPPORTED, "too many parameters"); }
    g.pfeat = (int)(sumI * R);
    g.foff[k] = g.pfeat;
    g.pf = g.pfeat + C * R;
    g.foff[k + 1] = g.pf;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete h; return fail(nullptr, TR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    if (prop.major < 10) { delete h; return fail(nullptr, TR_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); }
    h->sms = prop.multiProcessorCount;
    h->l2_bytes = (size_t)prop.l2CacheSize;
    if (const char* ev = getenv("TR_B200_FUSED")) h->fused_mode = atoi(ev) < 0 ? -1 : (atoi(ev) > 0 ? 1 : 0);
    if (const char* ev = getenv("TR_B200_FUSED_NS")) h->fused_ns = atoi(ev);
    if (const char* ev = getenv("TR_B200_FUSED_CL")) { const int v = atoi(ev); if (v >= 0 && v <= 16) h->fused_cl = v; }
    if (const char* ev = getenv("TR_B200_SPEC_SINGLE")) h->spec_single = atoi(ev) < 0 ? -1 : (atoi(ev) > 0 ? 1 : 0);
    if (const char* ev = getenv("TR_B200_FUSED_PACE")) h->fused_pace = atoi(ev);
    if (const char* ev = getenv("TR_B200_FUSED_PIECE")) { const int v = atoi(ev); if (v >= 16 && v % 16 == 0) h->fused_piece = v; }
    *out = h;
    return TR_OK;
}

int tr_destroy(tr_handle* h) {
    if (!h) return TR_OK;
    DeviceGuard dg(h->device);
    cudaDeviceSynchronize();
    Buf* bufs[] = {&h->FtT, &h->Ft64, &h->partial, &h->V, &h->u_ws, &h->dZ_ws, &h->Gpart, &h->Gred, &h->epi_part, &h->dfc_part,
                   &h->flow_ring, &h->flow_sync, &h->Apart, &h->Spart, &h->trace, &h->f3_stage,
                   &h->spA, &h->spDA, &h->spMc, &h->spU, &h->spDS, &h->spRes, &h->spPart, &h->spDf1};
    for (Buf* b : bufs) if (b->p) cudaFree(b->p);
    for (int i = 0; i < 6; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    delete h;
    return TR_OK;
}

const char* tr_last_error(tr_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int tr_param_count(tr_handle* h, int64_t* P, int64_t* Pf) {
    if (!h) return TR_ERR_INVALID;
    if (P) *P = h->geo.pf + tr_nbias(h);
    if (Pf) *Pf = h->geo.pf;
    return TR_OK;
}

int tr_gradsum_count(tr_handle* h, int64_t* count) {
    if (!h || !count) return TR_ERR_INVALID;
    *count = h->geo.pf + tr_nbias(h) + 1;
    return TR_OK;
}

int tr_reserve(tr_handle* h, int64_t N) {
    if (!h) return TR_ERR_INVALID;
    if (N < 1) return TR_OK;
    DeviceGuard dg(h->device);
    const int rk = h->geo.C > 0 ? h->geo.R : 1;
    int rc;
    for (int vec = 0; vec < 2; ++vec) {
        Plan pl;
        if (h->dtype == TR_F32) {
            const KEntry<float>* e;
            if ((rc = make_plan<float>(h, N, rk, vec != 0, &pl, &e))) return rc;
            if ((rc = reserve_for<float>(h, N, pl))) return rc;
        } else {
            const KEntry<double>* e;
            if ((rc = make_plan<double>(h, N, rk, vec != 0, &pl, &e))) return rc;
            if ((rc = reserve_for<double>(h, N, pl))) return rc;
        }
    }
    return TR_OK;
}

int tr_forward_std(tr_handle* h, const void* X, int64_t N, const void* theta, const void* w, uint32_t nn_mask,
                   double sp_beta, double sp_thr, void* yhat, void* stream) {
    int rc = check_common(h, X, N, theta, w);
    if (rc) return rc;
    if (h->geo.C != 0) return fail(h, TR_ERR_INVALID, "tr_forward_std on a multinomial handle");
    if (N == 0) return TR_OK;
    if (!yhat) return fail(h, TR_ERR_INVALID, "yhat is null");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return h->dtype == TR_F32 ? forward_std_t<float>(h, X, N, theta, w, nn_mask, sp_beta, sp_thr, yhat, st)
                              : forward_std_t<double>(h, X, N, theta, w, nn_mask, sp_beta, sp_thr, yhat, st);
}

int tr_forward_mn(tr_handle* h, const void* X, int64_t N, const void* theta, const void* w, uint32_t nn_mask,
                  double sp_beta, double sp_thr, void* P, int64_t* pred, void* stream) {
    int rc = check_common(h, X, N, theta, w);
    if (rc) return rc;
    if (h->geo.C == 0) return fail(h, TR_ERR_INVALID, "tr_forward_mn on a standard handle");
    if (N == 0) return TR_OK;
    if (!P && !pred) return fail(h, TR_ERR_INVALID, "P and pred are both null");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    return h->dtype == TR_F32
               ? mn_t<float>(h, X, nullptr, nullptr, N, theta, w, nn_mask, sp_beta, sp_thr, nullptr, P, (long long*)pred, st)
               : mn_t<double>(h, X, nullptr, nullptr, N, theta, w, nn_mask, sp_beta, sp_thr, nullptr, P, (long long*)pred, st);
}

static int zero_gradsum(tr_handle* h, double* gradsum, cudaStream_t st) {
    const size_t n = (size_t)h->geo.pf + tr_nbias(h) + 1;
    TR_CUDA(h, cudaMemsetAsync(gradsum, 0, n * sizeof(double), st));
    return TR_OK;
}

int tr_fwd_grad_std(tr_handle* h, const void* X, const void* y, int64_t N, const void* theta, const void* w,
                    uint32_t nn_mask, double sp_beta, double sp_thr, double* gradsum, void* yhat, void* stream) {
    int rc = check_common(h, X, N, theta, w);
    if (rc) return rc;
    if (h->geo.C != 0) return fail(h, TR_ERR_INVALID, "tr_fwd_grad_std on a multinomial handle");
    if (!gradsum || (N > 0 && !y)) return fail(h, TR_ERR_INVALID, "null y / gradsum pointer");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) return zero_gradsum(h, gradsum, st);
    return h->dtype == TR_F32
               ? fwd_grad_std_t<float>(h, X, y, N, theta, w, nn_mask, sp_beta, sp_thr, gradsum, yhat, st, false)
               : fwd_grad_std_t<double>(h, X, y, N, theta, w, nn_mask, sp_beta, sp_thr, gradsum, yhat, st, false);
}

int tr_backward_std(tr_handle* h, const void* X, const void* dyhat, int64_t N, const void* theta, const void* w,
                    uint32_t nn_mask, double sp_beta, double sp_thr, double* gradsum, void* stream) {
    int rc = check_common(h, X, N, theta, w);
    if (rc) return rc;
    if (h->geo.C != 0) return fail(h, TR_ERR_INVALID, "tr_backward_std on a multinomial handle");
    if (!gradsum || (N > 0 && !dyhat)) return fail(h, TR_ERR_INVALID, "null dyhat / gradsum pointer");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) return zero_gradsum(h, gradsum, st);
    return h->dtype == TR_F32
               ? fwd_grad_std_t<float>(h, X, dyhat, N, theta, w, nn_mask, sp_beta, sp_thr, gradsum, nullptr, st, true)
               : fwd_grad_std_t<double>(h, X, dyhat, N, theta, w, nn_mask, sp_beta, sp_thr, gradsum, nullptr, st, true);
}

int tr_fwd_grad_mn(tr_handle* h, const void* X, const int64_t* y, const void* class_w, int64_t N, const void* theta,
                   const void* w, uint32_t nn_mask, double sp_beta, double sp_thr, double* gradsum, void* P,
                   void* stream) {
    int rc = check_common(h, X, N, theta, w);
    if (rc) return rc;
    if (h->geo.C == 0) return fail(h, TR_ERR_INVALID, "tr_fwd_grad_mn on a standard handle");
    if (!gradsum || (N > 0 && (!y || !class_w))) return fail(h, TR_ERR_INVALID, "null y / class_w / gradsum pointer");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) return zero_gradsum(h, gradsum, st);
    return h->dtype == TR_F32
               ? mn_t<float>(h, X, (const long long*)y, class_w, N, theta, w, nn_mask, sp_beta, sp_thr, gradsum, P, nullptr, st)
               : mn_t<double>(h, X, (const long long*)y, class_w, N, theta, w, nn_mask, sp_beta, sp_thr, gradsum, P, nullptr, st);
}

int tr_backward_mn(tr_handle* h, const void* X, const void* dP, int64_t N, const void* theta, const void* w,
                   uint32_t nn_mask, double sp_beta, double sp_thr, double* gradsum, void* stream) {
    int rc = check_common(h, X, N, theta, w);
    if (rc) return rc;
    if (h->geo.C == 0) return fail(h, TR_ERR_INVALID, "tr_backward_mn on a standard handle");
    if (!gradsum || (N > 0 && !dP)) return fail(h, TR_ERR_INVALID, "null dP / gradsum pointer");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) return zero_gradsum(h, gradsum, st);
    return h->dtype == TR_F32
               ? mn_t<float>(h, X, nullptr, nullptr, N, theta, w, nn_mask, sp_beta, sp_thr, gradsum, nullptr, nullptr, st, dP)
               : mn_t<double>(h, X, nullptr, nullptr, N, theta, w, nn_mask, sp_beta, sp_thr, gradsum, nullptr, nullptr, st, dP);
}


// ---------------------------------------------------------------------------------------------
// spectral_tensor_regression.py (tr_spectral.cuh)
// ---------------------------------------------------------------------------------------------
int tr_spec_create(tr_handle** out, int dtype, int64_t W, int64_t D, int64_t n_out, int rank_normal, int rank_spectral,
                   int complex_dim, int device) {
    if (!out) return fail(nullptr, TR_ERR_INVALID, "out is null");
    *out = nullptr;
    if (W < 1 || D < 1 || n_out < 1 || W > (1 << 20) || D > (1 << 24))
        return fail(nullptr, TR_ERR_INVALID, "invalid geometry W=%lld D=%lld n_out=%lld", (long long)W, (long long)D, (long long)n_out);
    if (n_out > TR_MAX_CLASSES) return fail(nullptr, TR_ERR_UNSUPPORTED, "n_out=%lld outputs (supported: up to %d)", (long long)n_out, TR_MAX_CLASSES);
    if (rank_normal < 0 || rank_spectral < 0 || rank_normal + rank_spectral < 1 || rank_normal + rank_spectral > TRS_MAXR)
        return fail(nullptr, TR_ERR_UNSUPPORTED, "rank_normal=%d rank_spectral=%d (supported: 1 <= sum <= %d)", rank_normal, rank_spectral, TRS_MAXR);
    if (complex_dim < 1 || complex_dim > 16) return fail(nullptr, TR_ERR_UNSUPPORTED, "complex_dim=%d (supported: 1..16)", complex_dim);
    const int64_t one = 1;
    tr_handle* h = nullptr;
    int rc = tr_create(&h, dtype, 1, &one, 1, 0, device);       // device / dtype checks, SM count
    if (rc) return rc;
    h->kind = 2;
    SpecGeo& sg = h->sg;
    sg.W = (int)W; sg.D = (int)D; sg.NO = (int)n_out; sg.Rn = rank_normal; sg.Rs = rank_spectral; sg.CC = complex_dim;
    sg.Q = rank_normal + rank_spectral * complex_dim; sg.RT = rank_normal + rank_spectral;
    const long long sizes[6] = {W * rank_normal, D * rank_normal, n_out * rank_normal,
                                W * rank_spectral * complex_dim, D * rank_spectral, n_out * rank_spectral};
    long long off = 0;
    Geo& g = h->geo;
    memset(&g, 0, sizeof(g));
    g.k = 6; g.R = 1; g.C = 0; g.D = W * D;
    for (int i = 0; i < 6; ++i) {
        sg.off[i] = (int)off; g.foff[i] = (int)off; g.dims[i] = (int)sizes[i];
        off += sizes[i];
        if (off > (1LL << 30)) { tr_destroy(h); return fail(nullptr, TR_ERR_UNSUPPORTED, "too many parameters"); }
    }
    sg.off[6] = (int)off;
    g.foff[6] = (int)off; g.foff[7] = (int)off;
    g.pfeat = (int)off; g.pf = (int)off;
    *out = h;
    return TR_OK;
}

static int spec_check(tr_handle* h, const void* X, int64_t N, const void* theta, const void* w) {
    if (!h) return TR_ERR_INVALID;
    if (h->kind != 2) return fail(h, TR_ERR_INVALID, "tr_spec_* needs a handle made by tr_spec_create");
    if (N < 0) return fail(h, TR_ERR_INVALID, "N must be >= 0 (got %lld)", (long long)N);
    if ((N > 0 && !X) || !theta || !w) return fail(h, TR_ERR_INVALID, "null X / theta / w pointer");
    if ((uintptr_t)X % h->elt != 0) return fail(h, TR_ERR_INVALID, "X is not aligned to its element size");
    return TR_OK;
}

int tr_spec_forward(tr_handle* h, const void* X, int64_t N, const void* theta, const void* w, uint32_t nn_mask,
                    double sp_beta, double sp_thr, void* yhat, void* yhat_lin, void* spec_pred, void* latents, void* stream) {
    int rc = spec_check(h, X, N, theta, w);
    if (rc) return rc;
    if (N == 0) return TR_OK;
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (h->dtype == TR_F32)
        return run_spec<float>(h, (const float*)X, nullptr, N, (const float*)theta, (const float*)w, nn_mask, sp_beta, sp_thr,
                               nullptr, (float*)yhat, (float*)yhat_lin, (float*)spec_pred, (float*)latents, st);
    return run_spec<double>(h, (const double*)X, nullptr, N, (const double*)theta, (const double*)w, nn_mask, sp_beta, sp_thr,
                            nullptr, (double*)yhat, (double*)yhat_lin, (double*)spec_pred, (double*)latents, st);
}

int tr_spec_fwd_grad(tr_handle* h, const void* X, const void* y, int64_t N, const void* theta, const void* w,
                     uint32_t nn_mask, double sp_beta, double sp_thr, double* gradsum, void* yhat, void* stream) {
    int rc = spec_check(h, X, N, theta, w);
    if (rc) return rc;
    if (!gradsum || (N > 0 && !y)) return fail(h, TR_ERR_INVALID, "null y / gradsum pointer");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = zero_gradsum(h, gradsum, st))) return rc;
    if (N == 0) return TR_OK;
    if (h->dtype == TR_F32)
        return run_spec<float>(h, (const float*)X, (const float*)y, N, (const float*)theta, (const float*)w, nn_mask, sp_beta,
                               sp_thr, gradsum, (float*)yhat, nullptr, nullptr, nullptr, st);
    return run_spec<double>(h, (const double*)X, (const double*)y, N, (const double*)theta, (const double*)w, nn_mask, sp_beta,
                            sp_thr, gradsum, (double*)yhat, nullptr, nullptr, nullptr, st);
}

int tr_finish_grad(tr_handle* h, const double* gradsum, double grad_scale, double loss_scale, const void* theta,
                   double lambda_L2, uint32_t nn_mask, double sp_beta, double sp_thr, void* grad, double* loss,
                   void* stream) {
    if (!h) return TR_ERR_INVALID;
    if (!gradsum || !theta || !grad || !loss) return fail(h, TR_ERR_INVALID, "null pointer argument");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_gs = h->geo.pf + tr_nbias(h) + 1;
    if (h->dtype == TR_F32)
        k_finish<float><<<1, 1024, 0, st>>>(gradsum, n_gs, tr_nbias(h), grad_scale, loss_scale, (const float*)theta, h->geo,
                                            lambda_L2, nn_mask, sp_beta, sp_thr, (float*)grad, loss);
    else
        k_finish<double><<<1, 1024, 0, st>>>(gradsum, n_gs, tr_nbias(h), grad_scale, loss_scale, (const double*)theta, h->geo,
                                             lambda_L2, nn_mask, sp_beta, sp_thr, (double*)grad, loss);
    TR_LAUNCH_CHECK(h);
    return TR_OK;
}

static int adam_launch(tr_handle* h, void* theta, const void* grad, void* m, void* v, void* vmax, int64_t step,
                       double lr, const double* lr_groups, int n_groups, double beta1, double beta2, double eps,
                       double weight_decay, void* stream) {
    if (!h) return TR_ERR_INVALID;
    if (!theta || !grad || !m || !v) return fail(h, TR_ERR_INVALID, "null pointer argument");
    if (step < 1) return fail(h, TR_ERR_INVALID, "step is 1-based (got %lld)", (long long)step);
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const Geo& g = h->geo;
    const long long P = g.pf + tr_nbias(h);
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    const double bc2_sqrt = sqrt(bc2);
    AdamGroups ag;
    memset(&ag, 0, sizeof(ag));
    if (lr_groups && h->kind != 0) return fail(h, TR_ERR_UNSUPPORTED, "tr_adam_step_groups: not available for spectral handles");
    if (lr_groups) {
        // one group per factor in theta order (feature factors, class factor) and, standard model, the bias
        const int nfac = g.k + (g.C > 0 ? 1 : 0);
        const int want = nfac + (g.C == 0 ? 1 : 0);
        if (n_groups != want)
            return fail(h, TR_ERR_INVALID, "tr_adam_step_groups: %d learning rates given, this model has %d parameter groups", n_groups, want);
        ag.n_seg = want;
        for (int i = 0; i < nfac; ++i) ag.seg_end[i] = (i + 1 < nfac) ? g.foff[i + 1] : g.pf;
        if (g.C == 0) ag.seg_end[nfac] = g.pf + 1;
        for (int i = 0; i < want; ++i) ag.step_size[i] = lr_groups[i] / bc1;
    }
    const double step_size = lr / bc1;
    const int grid = (int)std::min<long long>((P + 255) / 256, 1024);
    if (h->dtype == TR_F32)
        k_adam<float><<<grid, 256, 0, st>>>((float*)theta, (const float*)grad, (float*)m, (float*)v, (float*)vmax, P,
                                            beta1, beta2, eps, weight_decay, step_size, bc2_sqrt, ag);
    else
        k_adam<double><<<grid, 256, 0, st>>>((double*)theta, (const double*)grad, (double*)m, (double*)v, (double*)vmax,
                                             P, beta1, beta2, eps, weight_decay, step_size, bc2_sqrt, ag);
    TR_LAUNCH_CHECK(h);
    return TR_OK;
}

int tr_adam_step(tr_handle* h, void* theta, const void* grad, void* m, void* v, void* vmax, int64_t step,
                 double lr, double beta1, double beta2, double eps, double weight_decay, void* stream) {
    return adam_launch(h, theta, grad, m, v, vmax, step, lr, nullptr, 0, beta1, beta2, eps, weight_decay, stream);
}

int tr_adam_step_groups(tr_handle* h, void* theta, const void* grad, void* m, void* v, void* vmax, int64_t step,
                        const double* lr_groups, int n_groups, double beta1, double beta2, double eps,
                        double weight_decay, void* stream) {
    if (h && !lr_groups) return fail(h, TR_ERR_INVALID, "lr_groups is null");
    return adam_launch(h, theta, grad, m, v, vmax, step, 0.0, lr_groups, n_groups, beta1, beta2, eps, weight_decay, stream);
}

// ---------------------------------------------------------------------------------------------
// Cross-GPU sum of the packed gradient sums: NCCL, resolved at run time so that
// the library neither links against nor requires NCCL on single-GPU hosts.
// ---------------------------------------------------------------------------------------------
namespace {
struct TrNcclId { char internal[128]; };     // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct NcclApi {
    void* lib = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, TrNcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string err;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("TR_B200_NCCL_LIB");
        // prefer the NCCL the host process already loaded (torch's bundled one), then the system library
        if (env && *env) api.lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) { api.err = std::string("NCCL library not found (libnccl.so.2; set TR_B200_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return; }
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
        api.GetVersion = (decltype(api.GetVersion))dlsym(api.lib, "ncclGetVersion");
        if (!api.AllReduce || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy) api.err = "NCCL library lacks ncclAllReduce / ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy";
    });
    return &api;
}
const char* nccl_str(NcclApi* a, int rc) { return a->GetErrorString ? a->GetErrorString(rc) : "?"; }
}  // namespace

int tr_comm_unique_id(void* id128) {
    if (!id128) return fail(nullptr, TR_ERR_INVALID, "id128 is null");
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return fail(nullptr, TR_ERR_UNSUPPORTED, "%s", a->err.c_str());
    const int rc = a->GetUniqueId(id128);
    if (rc != 0) return fail(nullptr, TR_ERR_CUDA, "ncclGetUniqueId failed: %s", nccl_str(a, rc));
    return TR_OK;
}

int tr_comm_create(void** comm, const void* id128, int rank, int world, int device) {
    if (!comm || !id128) return fail(nullptr, TR_ERR_INVALID, "null comm / id128 pointer");
    *comm = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(nullptr, TR_ERR_INVALID, "rank %d of world %d", rank, world);
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return fail(nullptr, TR_ERR_UNSUPPORTED, "%s", a->err.c_str());
    DeviceGuard dg(device);
    TrNcclId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = a->CommInitRank(comm, world, id, rank);
    if (rc != 0) return fail(nullptr, TR_ERR_CUDA, "ncclCommInitRank failed: %s", nccl_str(a, rc));
    return TR_OK;
}

int tr_comm_destroy(void* comm) {
    if (!comm) return TR_OK;
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return fail(nullptr, TR_ERR_UNSUPPORTED, "%s", a->err.c_str());
    const int rc = a->CommDestroy(comm);
    if (rc != 0) return fail(nullptr, TR_ERR_CUDA, "ncclCommDestroy failed: %s", nccl_str(a, rc));
    return TR_OK;
}

int tr_allreduce(tr_handle* h, double* buf, int64_t count, void* nccl_comm, void* stream) {
    if (!h) return TR_ERR_INVALID;
    if (!buf || count < 0) return fail(h, TR_ERR_INVALID, "null buffer / negative count");
    if (!nccl_comm) return fail(h, TR_ERR_INVALID, "nccl_comm is null (single GPU: do not call tr_allreduce)");
    if (count == 0) return TR_OK;
    NcclApi* a = nccl_api();
    if (!a->err.empty()) return fail(h, TR_ERR_UNSUPPORTED, "%s", a->err.c_str());
    DeviceGuard dg(h->device);
    const int rc = a->AllReduce(buf, buf, (size_t)count, /* ncclFloat64 */ 8, /* ncclSum */ 0, nccl_comm, (cudaStream_t)stream);
    if (rc != 0) return fail(h, TR_ERR_CUDA, "ncclAllReduce failed: %s", nccl_str(a, rc));
    return TR_OK;
}

int tr_profile_enable(tr_handle* h, int enable) {
    if (!h) return TR_ERR_INVALID;
    DeviceGuard dg(h->device);
    if (enable && !h->ev[0])
        for (int i = 0; i < 6; ++i) TR_CUDA(h, cudaEventCreate(&h->ev[i]));
    if (h->prof && !enable) { int rc = prof_fold(h, true); if (rc) return rc; }
    h->prof = enable != 0;
    if (enable) for (int i = 0; i < 3; ++i) { h->prof_ms[i] = 0.0; h->prof_n[i] = 0; h->ev_set[i] = false; }
    return TR_OK;
}

int tr_profile_read(tr_handle* h, double* out6) {
    if (!h || !out6) return TR_ERR_INVALID;
    DeviceGuard dg(h->device);
    int rc = prof_fold(h, true);
    if (rc) return rc;
    for (int i = 0; i < 3; ++i) { out6[2 * i] = h->prof_ms[i]; out6[2 * i + 1] = (double)h->prof_n[i]; }
    return TR_OK;
}

int tr_set_option(tr_handle* h, const char* name, int64_t value) {
    if (!h || !name) return TR_ERR_INVALID;
    if (strcmp(name, "fused") == 0) {
        if (value < -1 || value > 1) return fail(h, TR_ERR_INVALID, "option fused: -1 (auto), 0 (two-pass), 1 (single-pass)");
        h->fused_mode = (int)value;
        return TR_OK;
    }
    if (strcmp(name, "flow") == 0) {
#ifndef TR_WITH_FLOW
        if (value == 1) return fail(h, TR_ERR_UNSUPPORTED, "option flow=1: this build does not contain the experimental dataflow kernel (make FLOW=1)");
#endif
        if (value < -1 || value > 1) return fail(h, TR_ERR_INVALID, "option flow: 0 (never, default), 1 (always the experimental dataflow kernel), -1 (auto = 0 for now)");
        h->flow_mode = (int)value;
        return TR_OK;
    }
    if (strcmp(name, "flow_debug") == 0) { h->flow_debug = (int)value; return TR_OK; }
    if (strcmp(name, "flow_window_mb") == 0) {
        if (value < 1 || value > 4096) return fail(h, TR_ERR_INVALID, "flow_window_mb must be in 1..4096");
        h->flow_window_mb = value;
        return TR_OK;
    }
    if (strcmp(name, "spec_single") == 0) {
        if (value < -1 || value > 1) return fail(h, TR_ERR_INVALID, "option spec_single: -1 (auto), 0 (never), 1 (always the single-pass spectral kernel)");
        h->spec_single = (int)value;
        return TR_OK;
    }
    if (strcmp(name, "spec_single_ns") == 0) { h->spec_single_ns = (int)value; return TR_OK; }
    if (strcmp(name, "fused_pace") == 0) { h->fused_pace = (int)value; return TR_OK; }
    if (strcmp(name, "fused_ns") == 0) { h->fused_ns = (int)value; h->occ_clusters.clear(); return TR_OK; }
    if (strcmp(name, "fused_cl") == 0) {
        if (value < 0 || value > 16)
            return fail(h, TR_ERR_INVALID, "fused_cl must be 0 (auto) or a cluster size 1..16 (the multinomial kernel takes 1, 2, 4, 8, 16)");
        h->fused_cl = (int)value;
        h->occ_clusters.clear();
        return TR_OK;
    }
    if (strcmp(name, "fused_piece") == 0) {
        if (value < 16 || value % 16) return fail(h, TR_ERR_INVALID, "fused_piece must be a positive multiple of 16");
        h->fused_piece = (int)value;
        return TR_OK;
    }
    return fail(h, TR_ERR_INVALID, "unknown option '%s'", name);
}

int tr_lbfgs_direction(tr_handle* h, const void* g, void* prev_g, void* d, double t, int first, void* S, void* Y,
                       double* lstate, int history, double* scal4, void* stream) {
    if (!h) return TR_ERR_INVALID;
    if (!g || !prev_g || !d || !S || !Y || !lstate || !scal4) return fail(h, TR_ERR_INVALID, "null pointer argument");
    if (history < 1 || history > TR_LBFGS_MAX_HIST)
        return fail(h, TR_ERR_UNSUPPORTED, "history_size %d (supported: 1..%d)", history, TR_LBFGS_MAX_HIST);
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = h->geo.pf + tr_nbias(h);
    if (h->dtype == TR_F32)
        k_lbfgs_direction<float><<<1, 1024, 0, st>>>((const float*)g, (float*)prev_g, (float*)d, t, first, (float*)S,
                                                     (float*)Y, lstate, history, P, scal4);
    else
        k_lbfgs_direction<double><<<1, 1024, 0, st>>>((const double*)g, (double*)prev_g, (double*)d, t, first,
                                                      (double*)S, (double*)Y, lstate, history, P, scal4);
    TR_LAUNCH_CHECK(h);
    return TR_OK;
}

int tr_lbfgs_point(tr_handle* h, void* out, const void* x, double t, const void* d, void* stream) {
    if (!h) return TR_ERR_INVALID;
    if (!out || !x || !d) return fail(h, TR_ERR_INVALID, "null pointer argument");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = h->geo.pf + tr_nbias(h);
    const int grid = (int)std::min<long long>((P + 255) / 256, 1024);
    if (h->dtype == TR_F32) k_axpy_out<float><<<grid, 256, 0, st>>>((float*)out, (const float*)x, t, (const float*)d, P);
    else k_axpy_out<double><<<grid, 256, 0, st>>>((double*)out, (const double*)x, t, (const double*)d, P);
    TR_LAUNCH_CHECK(h);
    return TR_OK;
}

int tr_lbfgs_gtd(tr_handle* h, const void* g, const void* d, double* scal2, void* stream) {
    if (!h) return TR_ERR_INVALID;
    if (!g || !scal2) return fail(h, TR_ERR_INVALID, "null pointer argument");
    DeviceGuard dg(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = h->geo.pf + tr_nbias(h);
    if (h->dtype == TR_F32) k_lbfgs_gtd<float><<<1, 1024, 0, st>>>((const float*)g, (const float*)d, P, scal2);
    else k_lbfgs_gtd<double><<<1, 1024, 0, st>>>((const double*)g, (const double*)d, P, scal2);
    TR_LAUNCH_CHECK(h);
    return TR_OK;
}

#if defined(TRSS_TRACE) && !defined(TRM_TRACE)
// debug builds only (not part of include/tr_b200.h): the timeline of the last k_spec_single launch
int tr_debug_trace(tr_handle* h, long long* out, int n) {
    if (!h || !out || !h->trace.p) return TR_ERR_INVALID;
    DeviceGuard dg(h->device);
    TR_CUDA(h, cudaDeviceSynchronize());
    TR_CUDA(h, cudaMemcpy(out, h->trace.p, (size_t)std::min(n, TRSS_TRACE_N * TRSS_TRACE_EV) * sizeof(long long), cudaMemcpyDeviceToHost));
    return TR_OK;
}
#endif
#ifdef TRM_TRACE
// debug builds only (not part of include/tr_b200.h): the timeline of the last k_fused_mn launch, TRM_TRACE_N x TRM_TRACE_EV stamps
int tr_debug_trace(tr_handle* h, long long* out, int n) {
    if (!h || !out || !h->trace.p) return TR_ERR_INVALID;
    DeviceGuard dg(h->device);
    TR_CUDA(h, cudaDeviceSynchronize());
    TR_CUDA(h, cudaMemcpy(out, h->trace.p, (size_t)std::min(n, TRM_TRACE_N * TRM_TRACE_EV) * sizeof(long long), cudaMemcpyDeviceToHost));
    return TR_OK;
}
#endif

int tr_last_launch_info(tr_handle* h, int64_t* info8) {
    if (!h || !info8) return TR_ERR_INVALID;
    for (int i = 0; i < 8; ++i) info8[i] = h->info[i];
    return TR_OK;
}

}  // extern "C"
