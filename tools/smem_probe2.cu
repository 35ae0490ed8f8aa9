// smem_probe2.cu — LDS.128 sweep time over a 64 KB window at different shared-memory offsets (probe).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) probe(unsigned base_off, int sweeps, long long* out, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 256) reinterpret_cast<float*>(smem)[i] = 1.0f;
    __syncthreads();
    const uint32_t a0 = s32(smem + base_off) + threadIdx.x * 16;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int sw = 0; sw < sweeps; ++sw) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float x, y, z, w;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a0 + j * 4096) : "memory");
            acc += x + y + z + w;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345.f) sink[0] = acc;
}
int main() {
    long long* out; cudaMalloc(&out, 148 * 8); float* sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int sweeps = 1000;
    for (unsigned off : {0u, 32768u, 65536u, 98304u, 131072u, 139264u}) {
        probe<<<148, 256, 200 * 1024>>>(off, sweeps, out, sink);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i];
        printf("offset %6u KB: %.0f cycles per 64 KB sweep (8 warps x 16 LDS.128; ideal 512)  (%s)\n", off >> 10, avg / 148 / sweeps,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
