// Resident clusters per cluster size (cudaOccupancyMaxActiveClusters) for a 512-thread kernel with the shared-memory
// footprint of the single-pass kernels: shows how many SMs a launch with cluster size CL can occupy on this die
// (GPC floor-sweeping decides it).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k_dummy(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s: %d SMs\n", pr.name, pr.multiProcessorCount);
    cudaFuncSetAttribute(k_dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const size_t smems[3] = {64 * 1024, 110 * 1024, 212 * 1024};
    for (size_t smem : smems) {
        cudaFuncSetAttribute(k_dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        printf("smem %zu KB:", smem / 1024);
        for (int CL = 1; CL <= 16; ++CL) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(CL * pr.multiProcessorCount, 1, 1);
            cfg.blockDim = dim3(512, 1, 1);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int NC = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&NC, k_dummy, &cfg);
            if (e != cudaSuccess) { cudaGetLastError(); NC = -1; }
            printf(" CL%d:%d(%d)", CL, NC, NC * CL);
        }
        printf("\n");
    }
    return 0;
}
