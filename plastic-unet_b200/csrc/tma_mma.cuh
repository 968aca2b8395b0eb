// tma_mma.cuh — device helpers shared by the TMA-fed, warp-level tensor-core (mma.sync) parameter-gradient kernels
// (conv3x3_wgrad_tma.cu, convT_dw_tma.cu): mbarrier ring, tensor-map loads, m16n8k8 TF32 MMA, vector reduction.
#pragma once
#include <cuda.h>
#include "pu_common.cuh"

namespace pu {

// tensor maps over an (oy, ox)-offset H x W window of an NHWC tensor (conv3x3_tc.cu); out-of-window coordinates zero-fill
int tma_make_window_map(CUtensorMap* tm, const View& v, int B, int H, int W, int cb, int bw, int bh, int bn);
int tma_make_window_map_merged(CUtensorMap* tm, const View& v, int B, int H, int W, int bw, int bh, int bn);
int tma_make_plane_map(CUtensorMap* tm, const View& v, int B, int H, int W, int bw, int bh, int bn);  // one-channel tensor, 3-D

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wg_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void wg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wg_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a lost TMA transaction must not hang the GPU
__device__ __forceinline__ void wg_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  printf("pu TMA gradient kernel: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
  asm volatile("trap;");
}
__device__ __forceinline__ void wg_tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void wg_tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void wg_red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// 32-bit shared-memory load at [addr + OFF] (shared-space byte address, compile-time offset): the fragment loops walk the
// operand planes with ONE 32-bit address per plane and immediate offsets instead of 64-bit generic-pointer arithmetic per load
// (ncu of the first version: 32 % of all instructions were IMADs).  volatile: ordered after the mbarrier wait.
template <int OFF>
__device__ __forceinline__ unsigned wg_lds(uint32_t addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}
__device__ __forceinline__ void wg_mma(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

}  // namespace
}  // namespace pu
