// tma.cuh -- 1-D bulk asynchronous copies (the TMA engine's non-tensor form: cp.async.bulk, SASS UBLKCP) with mbarrier
// completion, used to stage whole operand rows global -> shared with ONE instruction per row instead of one LDGSTS per
// thread, plus the producer/consumer handshake around the staging slots.
//
// A bulk copy needs 16-byte aligned source and destination and a size that is a multiple of 16 bytes.  The caller's arrays
// keep the reference's dimensioning (isd:ied, jsd:jed, nk), so the first element a tile needs sits at an arbitrary element
// offset: a row is therefore fetched from the 16-byte boundary at or below it (`bulk_span`) and every consumer adds the
// row's `shift` (0 or 1 doubles; 0..15 mask bytes) when it reads its element back.  Base pointers must be 16-byte aligned
// (cudaMalloc / torch allocations are; the driver falls back to the per-thread LDGSTS kernels otherwise).
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

// global -> shared bulk copy; completion is signalled on `bar` as `bytes` transaction bytes
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// The 16-byte aligned span that covers elements [e0, e0 + n) of an array of `total` elements of ESZ bytes whose base is
// 16-byte aligned.  Returns the byte count (a multiple of 16; 0 if nothing is left), sets `first` to the first element of
// the span; element e0 then sits at index (e0 - first) of the staged row.  The span is clipped at the last 16-byte boundary
// inside the array: callers never need the elements beyond it (documented where the spans are built).
template <int ESZ>
__device__ __forceinline__ unsigned bulk_span(unsigned e0, unsigned n, unsigned total, unsigned &first)
{
    constexpr unsigned PER = 16 / ESZ;                 // elements per 16 bytes
    first = e0 & ~(PER - 1);
    unsigned last = (e0 + n + PER - 1) & ~(PER - 1);  // one past the span
    const unsigned lim = total & ~(PER - 1);
    if (last > lim) last = lim;
    return last > first ? (last - first) * ESZ : 0u;
}

// ---- tensor-map (tiled) copies: one instruction fetches a whole box of a strided array (SASS UTMALDG) ----------------
// The box origin is given in ELEMENT coordinates, so the caller's halo-1 dimensioning needs no alignment shift; what the
// tensor map does require is a 16-byte aligned base and strides that are multiples of 16 bytes, i.e. an even row length
// ni+2 for the FP64 arrays of this path (the driver uses the LDGSTS kernels for blocks with an odd ni or unaligned bases).
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
