// smem_probe3.cu — LDS.128 sweep time (8 warps x 16 LDS.128 over 64 KB) with / without a TMA bulk stream
// into two other 64 KB stages of the same CTA.  All SMs run (HBM saturated when tma=1).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void __launch_bounds__(288, 1) probe(const unsigned char* X, long long per_cta, int tma, int sweeps, int gap,
                                               long long* out, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];     // stage 0: LDS target; stages 1,2: TMA ring
    __shared__ uint64_t full[2];
    __shared__ volatile int stop;
    const unsigned stage = 65536;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    for (int i = threadIdx.x; i < 16384; i += 288) reinterpret_cast<float*>(smem)[i] = 1.0f;
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid == 8) {
        if (lane == 0 && tma) {
            const unsigned char* src = X + (long long)blockIdx.x * per_cta;
            const long long n = per_cta / stage;
            long long j = 0;
            for (; j < n && !stop; ++j) {
                const int s = (int)(j % 2);
                if (j >= 2) { while (!try_wait(&full[s], (unsigned)(((j / 2) - 1) & 1))) {} }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(stage) : "memory");
                for (unsigned off = 0; off < stage; off += 32768)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s32(smem + (size_t)(1 + s) * stage + off)), "l"(src + j * stage + off), "r"(32768u), "r"(s32(&full[s])) : "memory");
            }
            for (long long q = (j >= 2 ? j - 2 : 0); q < j; ++q) { const int s = (int)(q % 2); while (!try_wait(&full[s], (unsigned)((q / 2) & 1))) {} }
            out[gridDim.x + blockIdx.x] = j;
        }
    } else {
        const long long tstart = clock64();
        while (clock64() - tstart < 100000) {}
        const uint32_t a0 = s32(smem) + threadIdx.x * 16;
        float acc = 0.f;
        long long busy = 0;
        for (int sw = 0; sw < sweeps; ++sw) {
            const long long t0 = clock64();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x, y, z, w;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a0 + j * 4096) : "memory");
                acc += x + y + z + w;
            }
            busy += clock64() - t0;
            const long long tw = clock64(); while (clock64() - tw < gap) {}
        }
        if (threadIdx.x == 0) out[blockIdx.x] = busy;
        if (acc == 12345.f) sink[0] = acc;
        __syncwarp();
        if (threadIdx.x == 0) stop = 1;
    }
}
int main() {
    const long long per_cta = 128LL << 20;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    unsigned char* X; cudaMalloc(&X, per_cta * sms); cudaMemset(X, 1, per_cta * sms);
    long long* out; cudaMalloc(&out, 2 * sms * 8); float* sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sweeps = 300;
    for (int gap : {0, 2000})
        for (int tma = 0; tma < 2; ++tma) {
            cudaMemset(out, 0, 2 * sms * 8);
            cudaEventRecord(e0);
            probe<<<sms, 288, 3 * 65536>>>(X, per_cta, tma, sweeps, gap, out, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[2 * 148]; cudaMemcpy(h, out, 2 * sms * 8, cudaMemcpyDeviceToHost);
            double avg = 0, st = 0; for (int i = 0; i < sms; ++i) { avg += h[i]; st += h[sms + i]; }
            printf("gap %4d tma=%d: LDS sweep of 64 KB by 8 warps: %.0f cycles (warp 0's 16 LDS.128) | TMA %.0f GB/s | %.3f ms (%s)\n", gap, tma,
                   avg / sms / sweeps, tma ? st * 65536.0 / ms / 1e6 : 0.0, ms, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
