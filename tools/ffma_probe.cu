// ffma_probe.cu — issue rate of FFMA (3-register) vs FFMA2 (fma.rn.f32x2) on sm_100a, in the operand pattern of the
// streaming kernels: acc[e][r] += x[e] * c[r] (x broadcast over a channel pair).  One block per SM, W warps per block.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma_probe tools/ffma_probe.cu && tools/ffma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long d, a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

template <int MODE>
__global__ void probe(float* out, const float* in, int iters, long long* cyc) {
    float acc[8][6], x[8], c[6];
#pragma unroll
    for (int e = 0; e < 8; ++e) { x[e] = in[threadIdx.x + e]; for (int r = 0; r < 6; ++r) acc[e][r] = 0.f; }
#pragma unroll
    for (int r = 0; r < 6; ++r) c[r] = in[64 + r + (threadIdx.x & 7)];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (MODE == 0) {
#pragma unroll
                for (int r = 0; r < 6; ++r) acc[e][r] = fmaf(x[e], c[r], acc[e][r]);
            } else {
#pragma unroll
                for (int r = 0; r < 6; r += 2) ffma2(acc[e][r], acc[e][r + 1], x[e], x[e], c[r], c[r + 1]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) for (int r = 0; r < 6; ++r) s += acc[e][r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {1, 2, 4, 8, 16}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<148, warps * 32>>>(out, in, iters, cyc); else probe<1><<<148, warps * 32>>>(out, in, iters, cyc);
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            const double fma_per_thread = (double)iters * 48;
            const double per_smsp_warps = warps / 4.0 > 1 ? warps / 4.0 : 1;     // warps sharing one sub-partition (blockDim warps spread over 4)
            printf("%s warps/SM=%2d cycles=%lld  cycles per warp-instruction per SMSP: %.2f  FMA lanes/clk/SM: %.1f\n",
                   mode ? "FFMA2" : "FFMA ", warps, h, (double)h / (iters * (mode ? 24 : 48)) / per_smsp_warps,
                   fma_per_thread * warps * 32 / h);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
