// smem_probe.cu — latency of thread-issued shared-memory ops while a TMA bulk copy streams into the
// same SM's shared memory (design probe for tr_fused.cuh).  1 CTA per SM, all SMs stream (so each
// 64 KB copy takes a few thousand cycles).  Warp 0 lane 0 issues the copies; probe warps 1..4
// (sub-partitions 1,2,3,0) time a dependent chain of 8 ops of one kind: LDS, STS, SHFL, mbarrier.arrive.
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
__global__ void __launch_bounds__(160, 1) probe(const unsigned char* X, long long per_cta, int tma, int kind, long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];     // 2 x 64 KB ring + 4 KB scratch
    __shared__ uint64_t full[2];
    __shared__ uint64_t dummy[8];
    __shared__ volatile int stop;
    const unsigned stage = 65536;
    float* scratch = reinterpret_cast<float*>(smem + 2 * stage);
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
        for (int s = 0; s < 8; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(s32(&dummy[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wid == 0) {
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
                                 ::"r"(s32(smem + (size_t)s * stage + off)), "l"(src + j * stage + off), "r"(32768u), "r"(s32(&full[s])) : "memory");
            }
            for (long long q = (j >= 2 ? j - 2 : 0); q < j; ++q) { const int s = (int)(q % 2); while (!try_wait(&full[s], (unsigned)((q / 2) & 1))) {} }
        }
    } else {
        // let the stream reach steady state
        const long long tstart = clock64();
        while (clock64() - tstart < 200000) {}
        long long worst = 0, sum = 0;
        const int reps = 200;
        float v = (float)lane;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (kind == 0) {            // dependent LDS chain
                    float w; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(s32(scratch + ((int)v & 31) + wid * 64)) : "memory"); v = w * 0.f + (float)lane;
                } else if (kind == 1) {     // STS then LDS of the same word (forces the store to complete)
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s32(scratch + lane + wid * 64)), "f"(v) : "memory");
                    float w; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(s32(scratch + lane + wid * 64)) : "memory"); v = w;
                } else if (kind == 2) {     // dependent shuffle chain
                    v += __shfl_xor_sync(0xffffffffu, v, 1 << (q % 5));
                } else {                    // mbarrier.arrive returning the state (dependent through the token)
                    unsigned long long tok;
                    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(tok) : "r"(s32(&dummy[wid])) : "memory");
                    v += (float)(tok & 1);
                }
            }
            const long long t1 = clock64();
            const long long d = t1 - t0;
            sum += d; if (d > worst) worst = d;
            const long long tw = clock64(); while (clock64() - tw < 500) {}
        }
        if (lane == 0) { out[(blockIdx.x * 4 + (wid - 1)) * 2 + 0] = sum / reps; out[(blockIdx.x * 4 + (wid - 1)) * 2 + 1] = worst; }
        if (v == 1234567.f) out[0] = 0;
        __syncwarp();
        if (wid == 1 && lane == 0) stop = 1;
    }
}
int main() {
    const long long per_cta = 128LL << 20;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    unsigned char* X; cudaMalloc(&X, per_cta * sms); cudaMemset(X, 1, per_cta * sms);
    long long* out; cudaMalloc(&out, sms * 8 * sizeof(long long));
    const size_t smem = 2 * 65536 + 4096;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[] = {"LDS chain x8", "STS+LDS x8", "SHFL chain x8", "mbarrier.arrive x8"};
    for (int kind = 0; kind < 4; ++kind)
        for (int tma = 0; tma < 2; ++tma) {
            cudaMemset(out, 0, sms * 8 * sizeof(long long));
            probe<<<sms, 160, smem>>>(X, per_cta, tma, kind, out);
            cudaDeviceSynchronize();
            long long h[148 * 8]; cudaMemcpy(h, out, sms * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
            printf("%-20s tma=%d :", names[kind], tma);
            for (int w = 0; w < 4; ++w) {
                double avg = 0; long long worst = 0;
                for (int b = 0; b < sms; ++b) { avg += h[(b * 4 + w) * 2]; if (h[(b * 4 + w) * 2 + 1] > worst) worst = h[(b * 4 + w) * 2 + 1]; }
                printf("  warp%d(smsp%d) avg %6.0f worst %6lld |", w + 1, (w + 1) % 4, avg / sms, worst);
            }
            printf(" (%s)\n", cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
