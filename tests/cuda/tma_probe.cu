// tma_probe.cu -- which (box, origin) combinations of an FP64 tensor-map copy does this GPU accept?  (diagnostic; run on the GPU box)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; ./tma_probe NX NY NZ B0 B1 B2 C0 C1 C2
//   (the binary is built here and travels to the GPU box with the snapshot; it is git-ignored.  Results: profiles/tma_probe_r02.log)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, unsigned bytes, double *out, int n)
{
    extern __shared__ __align__(128) double sm[];
    __shared__ __align__(8) uint64_t bar;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(&bar), sd = (unsigned)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(sd),
                     "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(sb)
                     : "memory");
    }
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(sb) : "memory");
    }
    for (int q = threadIdx.x; q < n; q += blockDim.x) out[q] = sm[q];
}

int main(int argc, char **argv)
{
    if (argc < 10) return 2;
    const int NX = atoi(argv[1]), NY = atoi(argv[2]), NZ = atoi(argv[3]), B0 = atoi(argv[4]), B1 = atoi(argv[5]), B2 = atoi(argv[6]);
    const int C0 = atoi(argv[7]), C1 = atoi(argv[8]), C2 = atoi(argv[9]);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encoder\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const size_t N = (size_t)NX * NY * NZ;
    std::vector<double> hsrc(N);
    for (size_t e = 0; e < N; e++) hsrc[e] = (double)e + 1.0;
    double *d, *o;
    cudaMalloc(&d, N * 8);
    cudaMemcpy(d, hsrc.data(), N * 8, cudaMemcpyHostToDevice);
    const int nbox = B0 * B1 * B2;
    cudaMalloc(&o, nbox * 8);
    cudaMemset(o, 0xff, nbox * 8);
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)NX, (cuuint64_t)NY, (cuuint64_t)NZ};
    const cuuint64_t strides[2] = {(cuuint64_t)NX * 8, (cuuint64_t)NX * NY * 8};
    const cuuint32_t box[3] = {(cuuint32_t)B0, (cuuint32_t)B1, (cuuint32_t)B2}, es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("dims %dx%dx%d box %dx%dx%d : ENCODE FAILED %d\n", NX, NY, NZ, B0, B1, B2, (int)r); return 0; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<1, 128, nbox * 8>>>(m, C0, C1, C2, (unsigned)(nbox * 8), o, nbox);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("dims %dx%dx%d box %dx%dx%d at (%d,%d,%d): RUN FAILED %s\n", NX, NY, NZ, B0, B1, B2, C0, C1, C2, cudaGetErrorString(e)); return 0; }
    std::vector<double> h(nbox);
    cudaMemcpy(h.data(), o, nbox * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int z = 0; z < B2; z++)
        for (int y = 0; y < B1; y++)
            for (int x = 0; x < B0; x++) {
                const long gx = C0 + x, gy = C1 + y, gz = C2 + z;
                const bool in = gx >= 0 && gx < NX && gy >= 0 && gy < NY && gz >= 0 && gz < NZ;
                const double want = in ? hsrc[gx + (size_t)NX * (gy + (size_t)NY * gz)] : 0.0;
                if (h[x + B0 * (y + B1 * z)] != want) bad++;
            }
    printf("dims %dx%dx%d box %dx%dx%d at (%d,%d,%d): ok, %d wrong elements\n", NX, NY, NZ, B0, B1, B2, C0, C1, C2, bad);
    return 0;
}
