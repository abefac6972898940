// tma.cuh -- 2-D TMA tensor copies (cp.async.bulk.tensor.2d, SASS UTMALDG) shared by the gather
// kernels of rectify (gather.cu) and reproject (reproject.cu): a CTA pulls the box of source pixels
// its target tile reaches into shared memory, one box per band, through an mbarrier ring.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace xrs {

// host (gather.cu): driver entry point looked up at run time, no link-time dependency on libcuda
bool tma_available();
bool tma_encode_2d(CUtensorMap *map, int elem_size, const void *base, uint64_t width, uint64_t height,
                   uint64_t pitch_bytes, uint32_t box_w, uint32_t box_h);

__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

// base + index * sizeof(T) as ONE wide multiply-add (the compiler's shift-and-add pair costs two
// issue slots per store in the band loops)
template <typename T>
__device__ __forceinline__ T *elem_ptr(T *base, uint32_t index) {
    uint64_t out;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(out) : "r"(index), "n"(sizeof(T)), "l"(reinterpret_cast<uint64_t>(base)));
    return reinterpret_cast<T *>(out);
}

}  // namespace xrs
