// TMA tile loads (cp.async.bulk.tensor.2d + mbarrier) used by the image kernels. A level's frames are contiguous
// in the level-major pyramid (frame stride = pitch * h), so each level is ONE 2-D u8 tensor of pitch x (h * batch)
// bytes with row stride pitch; a tile of frame f at (x, y) is the box at (x, f * h + y). Measured on B200: the
// box must start on a 16-byte boundary of the row (any other x raises "illegal instruction"); rows/columns
// outside the tensor are zero-filled, negative coordinates are fine.
#pragma once
#include <cuda.h>
#include <stdint.h>

// arm `bar` (8-byte aligned shared memory) for one arrival + `bytes`, then issue the box copy into `dst`
// (128-byte aligned shared memory). Call from ONE thread; everybody else waits with tma_wait after a barrier.
static __device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* tmap, int x, int y, void* bar_ptr,
                                                     uint32_t bytes) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bar_ptr);
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(d),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(bar)
               : "memory");
}

// wait for phase 0 of the barrier armed by tma_load_tile (single-use barrier)
static __device__ __forceinline__ void tma_wait(void* bar_ptr) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bar_ptr);
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(bar)
                 : "memory");
  } while (!done);
}
