// tma.cuh -- tensor-map asynchronous copies (the TMA engine: cp.async.bulk.tensor, SASS UTMALDG) with mbarrier completion, used to
// stage whole operand boxes global -> shared with ONE instruction per box instead of one LDGSTS per thread and element, plus the
// producer / consumer handshake around the staging slots.
#pragma once

#include <stdint.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp (elect.sync): the form ptxas recognises as "single thread", so the bulk-copy instructions
// behind it are emitted once instead of each inside its own elect/vote loop (which `lane == 0` would cause).
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make freshly initialised barriers visible to the asynchronous proxy (the TMA engine) before the first bulk copy
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order this thread's earlier generic-proxy accesses to shared memory before later asynchronous-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// ---- tensor-map (tiled) copies: one instruction fetches a whole box of a strided array (SASS UTMALDG) ----------------
// Requirements, the last one MEASURED on B200 (tests/cuda/tma_probe.cu, profiles/tma_probe_r02.log):
//   - 16-byte aligned base and strides that are multiples of 16 bytes, i.e. an even row length ni+2 for the FP64 arrays of this
//     path (the driver uses the LDGSTS kernels for blocks with an odd ni or unaligned bases);
//   - the box must START on a 16-byte boundary: an odd first coordinate of an FP64 box raises "illegal instruction" (u8 boxes are
//     started at a multiple of 16 accordingly).  Negative / out-of-range coordinates are fine.
// Elements of a box that fall outside the array are filled with zeros and still count as transaction bytes.
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
