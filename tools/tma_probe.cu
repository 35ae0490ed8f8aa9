// tma_probe.cu — how many bytes can cp.async.bulk keep in flight per SM?  (design probe for tr_fused.cuh)
// One CTA per SM, one thread issues bulk copies of `stage` bytes (in `piece`-byte instructions) with NS
// copies in flight; no consumer work.  Prints GB/s for a few (stage, piece, NS, issuing threads).
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

__global__ void probe(const unsigned char* X, long long per_cta_bytes, unsigned stage, unsigned piece, int NS, int issuers) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(issuers));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= issuers) return;
    const unsigned char* src = X + (long long)blockIdx.x * per_cta_bytes;
    const long long n = per_cta_bytes / stage;
    const unsigned share = stage / issuers;           // bytes of a stage this thread issues
    for (long long j = 0; j < n; ++j) {
        const int s = (int)(j % NS);
        if (j >= NS) { while (!try_wait(&full[s], (unsigned)(((j / NS) - 1) & 1))) {} }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(share) : "memory");
        for (unsigned off = 0; off < share; off += piece) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(smem + (size_t)s * stage + t * share + off)), "l"(src + j * stage + t * share + off),
                           "r"(piece), "r"(s32(&full[s])) : "memory");
        }
    }
    for (int s = 0; s < NS && s < n; ++s) {
        const long long last = ((n - 1 - s) / NS) * NS + s;
        while (!try_wait(&full[s], (unsigned)((last / NS) & 1))) {}
    }
}

int main() {
    const long long per_cta = 256LL << 20;      // 256 MiB per CTA -> 37 GiB total over 148 CTAs
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    unsigned char* X; cudaMalloc(&X, per_cta * sms); cudaMemset(X, 1, per_cta * sms);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { unsigned stage, piece; int NS, issuers, ctas; } cfgs[] = {
        {65536, 16384, 1, 1, 148}, {65536, 16384, 2, 1, 148}, {65536, 16384, 3, 1, 148},
        {65536, 65536, 3, 1, 148}, {65536, 4096, 3, 1, 148}, {65536, 2048, 3, 4, 148}, {65536, 1024, 3, 16, 148},
        {32768, 16384, 6, 1, 148}, {32768, 8192, 6, 4, 148}, {16384, 16384, 12, 1, 148}, {16384, 4096, 12, 4, 148},
        {65536, 16384, 3, 1, 120}, {65536, 16384, 2, 1, 120}, {65536, 4096, 3, 16, 120},
    };
    for (auto c : cfgs) {
        const size_t smem = (size_t)c.NS * c.stage;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            probe<<<c.ctas, 128, smem>>>(X, per_cta, c.stage, c.piece, c.NS, c.issuers);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 1)
                printf("stage %6u piece %6u NS %2d issuers %2d ctas %3d inflight/SM %4zu KB : %8.1f GB/s  (%s)\n", c.stage, c.piece,
                       c.NS, c.issuers, c.ctas, smem >> 10, (double)per_cta * c.ctas / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
