// fused_trace.cu — per-sample timeline (clock64) of k_fused_std for cluster 0 / CTA rank 0.
// Debug probe; not part of the library.  nvcc -I../tensor_regression_b200/csrc -I../include
#include <cstdio>
#include <vector>
#define TR_FUSED_TRACE 1
#include "tr_fused.cuh"

int main(int argc, char** argv) {
    const int CL = argc > 1 ? atoi(argv[1]) : 8;
    const int NSreq = argc > 2 ? atoi(argv[2]) : 3;
    const int pace = argc > 3 ? atoi(argv[3]) : 0;
    const unsigned piece = argc > 4 ? (unsigned)atoi(argv[4]) : 32768u;
    Geo g; memset(&g, 0, sizeof(g));
    g.k = 3; g.R = 8; g.C = 0; g.dims[0] = 64; g.dims[1] = 64; g.dims[2] = 32; g.D = 131072;
    g.foff[0] = 0; g.foff[1] = 512; g.foff[2] = 1024; g.pfeat = 1280; g.foff[3] = 1280; g.pf = 1280; g.foff[4] = 1280;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    typedef float T;
    const int Dc = (int)(g.D / CL);
    const int E = (Dc / 4 + TR_FUSED_NCT - 1) / TR_FUSED_NCT;
    const unsigned stage = Dc * sizeof(T);
    const size_t fixed = ((((sizeof(FusedCtl) + 15) / 16) * 16 + (size_t)(g.pfeat + g.R) * sizeof(T)) + 1023) / 1024 * 1024;
    int NS = (int)((226 * 1024 - fixed) / stage); if (NS > TR_FUSED_MAX_NS) NS = TR_FUSED_MAX_NS; if (NS > NSreq) NS = NSreq;
    const size_t smem = fixed + (size_t)NS * stage + TR_TRACE_N * TR_TRACE_EV * 8;
    auto kern = E == 16 ? k_fused_std<T, 16> : (E == 8 ? k_fused_std<T, 8> : k_fused_std<T, 4>);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (CL > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * prop.multiProcessorCount); cfg.blockDim = dim3(TR_FUSED_NT); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int NC = 0; cudaOccupancyMaxActiveClusters(&NC, kern, &cfg);
    const long long N = (long long)NC * 400;
    printf("CL %d E %d NS %d NC %d smem %zu N %lld\n", CL, E, NS, NC, smem, N);
    T *X, *y, *Ft, *w, *theta, *Gp; double* part; long long* trace;
    cudaMalloc(&X, N * g.D * sizeof(T)); cudaMemset(X, 0, N * g.D * sizeof(T));
    cudaMalloc(&y, N * sizeof(T)); cudaMemset(y, 0, N * sizeof(T));
    std::vector<T> hf(g.pf + 1, 0.01f), hw(g.R, 1.f);
    cudaMalloc(&Ft, hf.size() * sizeof(T)); cudaMemcpy(Ft, hf.data(), hf.size() * sizeof(T), cudaMemcpyHostToDevice);
    cudaMalloc(&theta, hf.size() * sizeof(T)); cudaMemcpy(theta, hf.data(), hf.size() * sizeof(T), cudaMemcpyHostToDevice);
    cudaMalloc(&w, g.R * sizeof(T)); cudaMemcpy(w, hw.data(), g.R * sizeof(T), cudaMemcpyHostToDevice);
    cudaMalloc(&Gp, (size_t)NC * g.D * sizeof(T)); cudaMalloc(&part, NC * 2 * sizeof(double));
    cudaMalloc(&trace, TR_TRACE_N * TR_TRACE_EV * sizeof(long long)); cudaMemset(trace, 0, TR_TRACE_N * TR_TRACE_EV * sizeof(long long));
    FusedArgs<T> a; a.X = X; a.y = y; a.N = N; a.FtT = Ft; a.w = w; a.theta = theta; a.bias_off = g.pf; a.geo = g;
    a.Gpart = Gp; a.Dpad = g.D; a.yhat = nullptr; a.res = y; a.CL = CL; a.NC = NC; a.Dc = Dc; a.NS = NS; a.nchunk = 1;
    a.spc = 1 << 30; a.stage_bytes = stage; a.trace = trace; a.pace = pace; a.piece = piece;
    cfg.gridDim = dim3(CL * NC);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); cudaLaunchKernelEx(&cfg, kern, a); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d: %.3f ms, %.1f GB/s, %.0f ns/sample/cluster (%s)\n", rep, ms, N * g.D * 4.0 / ms / 1e6, ms * 1e6 / 400,
               cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<long long> t(TR_TRACE_N * TR_TRACE_EV);
    cudaMemcpy(t.data(), trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const long long t0 = t[0];
    printf("sample: issue  full_seen  A_done  redA_done  cready  res_seen  B_done   (cycles since issue of sample %d; deltas vs own issue)\n", TR_TRACE_I0);
    for (int i = 0; i < TR_TRACE_N; ++i) {
        long long* e = &t[i * TR_TRACE_EV];
        printf("%3d: issue@%8lld | full +%6lld  fma +%6lld Adone +%6lld  redA +%6lld  cready +%6lld  res +%6lld  Bdone +%6lld\n", i + TR_TRACE_I0,
               e[0] - t0, e[1] - e[0], e[7] - e[0], e[2] - e[0], e[3] - e[0], e[4] - e[0], e[5] - e[0], e[6] - e[0]);
    }
    printf("per-warp phase A [start,dur] and phase B [start,dur] relative to the sample's issue:\n");
    for (int i = 0; i < 12; ++i) {
        long long* e = &t[i * TR_TRACE_EV];
        printf("%3d A:", i + TR_TRACE_I0);
        for (int w = 0; w < 8; ++w) printf(" [%5lld,%4lld]", e[8 + 2 * w] - e[0], e[9 + 2 * w] - e[8 + 2 * w]);
        printf("\n    B:");
        for (int w = 0; w < 8; ++w) printf(" [%5lld,%4lld]", e[24 + 2 * w] - e[0], e[25 + 2 * w] - e[24 + 2 * w]);
        printf("\n");
    }
    return 0;
}
