// Which shared-memory elements does tcgen05.mma kind::tf32 read through a K-major SWIZZLE_{32,64,128}B descriptor whose
// start address is shifted by an arbitrary number of rows?  (Design probe for the conv3x3 tap shifts; not product code.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probes/umma_swizzle_probe scripts/probes/umma_swizzle_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Variant { int swz_bytes, shift, k0, use_base_off, sbo_rows; };

__global__ void __launch_bounds__(128, 1) probe(const Variant* vars, int nvar, int fill, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [0, 32K): A buffer; [32K, 33K): B (16 x 8 identity, K-major SWIZZLE_NONE: [kchunk][n][4]); then barrier + tmem slot
  uint8_t* sA = smem;
  float* sB = reinterpret_cast<float*>(smem + 32768);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 16 * 8; i += 128) {
    const int n = i / 8, k = i % 8;
    sB[((k / 4) * 16 + n) * 4 + (k % 4)] = (n == k) ? 1.f : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  uint32_t phase = 0;
  for (int v = 0; v < nvar; ++v) {
    const Variant var = vars[v];
    const int rowbytes = var.swz_bytes, nrows = 32768 / rowbytes, cpr = rowbytes / 4;
    const uint32_t mask = (uint32_t)(rowbytes / 16 - 1);  // 1, 3, 7
    for (int i = tid; i < nrows * cpr; i += 128) {
      const int R = i / cpr, kc = i % cpr;
      uint32_t lin = (uint32_t)(R * rowbytes + kc * 4);
      lin ^= ((lin >> 7) & mask) << 4;  // what TMA SWIZZLE_xB writes (Swizzle<B,4,3> on the byte address)
      *reinterpret_cast<float*>(sA + lin) = fill == 0 ? (float)R : (float)kc;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const uint32_t start = smem_u32(sA) + (uint32_t)(var.shift * rowbytes + var.k0 * 4);
      const uint64_t lt = var.swz_bytes == 32 ? 6 : (var.swz_bytes == 64 ? 4 : 2);
      uint64_t adesc = (uint64_t)((start >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(((uint32_t)(var.sbo_rows * rowbytes) >> 4) & 0x3FFF) << 32) |
                       (1ull << 46) | (lt << 61);
      if (var.use_base_off) adesc |= (uint64_t)((start >> 7) & 7) << 49;
      const uint64_t bdesc = (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF) | ((uint64_t)((16 * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 22) && !done; ++it)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    if (!done) { asm volatile("trap;"); }
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[((size_t)v * 128 + tid) * 16 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<Variant> vars;
  const int shifts[] = {0, 1, 2, 3, 5, 8, 9, 17, 66, 132};
  for (int swz : {32, 64, 128})
    for (int sbo : {8, 6})
      for (int k0 = 0; k0 < swz / 4; k0 += 8)
        for (int s : shifts) vars.push_back({swz, s, k0, 0, sbo});
  const int nv = (int)vars.size();
  Variant* dv; float *d0, *d1;
  cudaMalloc(&dv, nv * sizeof(Variant)); cudaMemcpy(dv, vars.data(), nv * sizeof(Variant), cudaMemcpyHostToDevice);
  cudaMalloc(&d0, (size_t)nv * 128 * 16 * 4); cudaMalloc(&d1, (size_t)nv * 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  probe<<<1, 128, 40960>>>(dv, nv, 0, d0);
  probe<<<1, 128, 40960>>>(dv, nv, 1, d1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> h0((size_t)nv * 128 * 16), h1(h0.size());
  cudaMemcpy(h0.data(), d0, h0.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h1.data(), d1, h1.size() * 4, cudaMemcpyDeviceToHost);
  for (int v = 0; v < nv; ++v) {
    int bad = 0, first_bad = -1;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 8; ++n) {
        const int R = (int)h0[((size_t)v * 128 + r) * 16 + n], kc = (int)h1[((size_t)v * 128 + r) * 16 + n];
        if (R != (r / 8) * vars[v].sbo_rows + (r % 8) + vars[v].shift || kc != vars[v].k0 + n) { if (first_bad < 0) first_bad = r * 8 + n; ++bad; }
      }
    printf("swz=%3d sbo_rows=%d k0=%2d shift=%3d : %s (%d/1024 wrong)", vars[v].swz_bytes, vars[v].sbo_rows, vars[v].k0, vars[v].shift,
           bad == 0 ? "OK " : "BAD", bad);
    if (bad) {
      printf("  rows 0..9 read (R,kc0,kc4): ");
      for (int r = 0; r < 10; ++r)
        printf("(%d,%d,%d) ", (int)h0[((size_t)v * 128 + r) * 16], (int)h1[((size_t)v * 128 + r) * 16], (int)h1[((size_t)v * 128 + r) * 16 + 4]);
    }
    printf("\n");
  }
  return 0;
}
