// tcgen05.mma kind::tf32 issue/throughput probe: cycles per MMA (M=128, K=8) as a function of N, the number of
// concurrently issuing warps, whether consecutive MMAs share an accumulator, and the operand layouts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probes/umma_rate_probe scripts/probes/umma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Variant { int n, warps, same_acc, a_swz /*0 none,32,128*/, b_swz /*0,32*/, kind /*0 tf32, 1 bf16*/, sbo_rows, shift_rows, commit_every; };

__global__ void __launch_bounds__(256, 1) probe(const Variant* vars, int nvar, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 98304);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 98304 / 4; i += 256) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 16 + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  uint32_t phase = 0;  // per warp: flips only when this warp took part
  for (int v = 0; v < nvar; ++v) {
    const Variant var = vars[v];
    __syncthreads();
    const long long t0 = clock64();
    if (warp < var.warps && lane == 0) {
      // A: 8 KB region per warp; B: at 64 KB
      const uint32_t a_addr = smem_u32(smem) + (warp & 3) * 16384;
      const uint32_t b_addr = smem_u32(smem) + 90112;
      uint64_t adesc, bdesc;
      if (var.a_swz == 0) adesc = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      else if (var.a_swz == 32) adesc = (uint64_t)(((a_addr + var.shift_rows * 32) >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((var.sbo_rows * 32) >> 4) << 32) | (1ull << 46) | (6ull << 61);
      else adesc = (uint64_t)(((a_addr + var.shift_rows * 128) >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((var.sbo_rows * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
      if (var.b_swz == 0) bdesc = (uint64_t)((b_addr >> 4) & 0x3FFF) | ((uint64_t)((var.n * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      else bdesc = (uint64_t)((b_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
      const uint32_t fmt = var.kind == 0 ? 2u : 1u;  // tf32 : bf16
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(var.n >> 3) << 17) | ((128u >> 4) << 24);
      // accumulator columns: each warp its own 512/warps range; within it rotate unless same_acc
      const int span = 512 / (var.warps == 7 ? 8 : var.warps);
      const int nslots = var.same_acc ? 1 : (span / var.n > 0 ? span / var.n : 1);
      int sidx = 0;
      for (int i = 0; i < iters; ++i) {
        int col = warp * span + sidx * var.n;
        if (col > 512 - var.n) col = 512 - var.n;
        const uint32_t d = tmem + (uint32_t)col;
        if (var.kind == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
        if (++sidx == nslots) sidx = 0;
        if (var.commit_every > 0 && (i + 1) % var.commit_every == 0)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 16 + warp)) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + warp)) : "memory");
    }
    const long long t1 = clock64();
    if (warp < var.warps) {
      uint32_t done = 0;
      for (uint32_t it = 0; it < (1u << 20) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar + warp)), "r"(phase) : "memory");
      if (!done) asm volatile("trap;");
      phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
    const long long t2 = clock64();
    if (tid == 0) { out[2 * v] = t1 - t0; out[2 * v + 1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<Variant> vars;
  for (int n : {32, 96})
    for (int w : {1, 2, 4, 7})
      for (int ce : {0, 24, 6, 3, 1})
        vars.push_back({n, w, 0, 32, 32, 0, 6, 0, ce});
  const int nv = (int)vars.size(), iters = 512;
  Variant* dv; long long* dout;
  cudaMalloc(&dv, nv * sizeof(Variant)); cudaMemcpy(dv, vars.data(), nv * sizeof(Variant), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, nv * 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100352);
  probe<<<1, 256, 100352>>>(dv, nv, iters, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> h(2 * nv);
  cudaMemcpy(h.data(), dout, nv * 16, cudaMemcpyDeviceToHost);
  for (int v = 0; v < nv; ++v) {
    const double total = (double)vars[v].warps * iters;
    printf("%s N=%3d warps=%d same_acc=%d a_swz=%3d b_swz=%2d commit_every=%2d : issue %.1f clk/MMA/warp, complete %.1f clk per MMA (SM-wide)\n",
           vars[v].kind ? "bf16" : "tf32", vars[v].n, vars[v].warps, vars[v].same_acc, vars[v].a_swz, vars[v].b_swz, vars[v].commit_every,
           (double)h[2 * v] / iters, (double)h[2 * v + 1] / total);
  }
  return 0;
}
