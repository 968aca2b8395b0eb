// Shared helpers for the plastic-unet B200 kernel library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "plastic_unet_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "plastic-unet_b200 kernels are written for sm_100a only"
#endif

namespace pu {

// ---- error / accounting (defined in pu_api.cu) -------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int post_launch(const char* what);  // cudaGetLastError -> status, counts one launch

#define PU_REQUIRE(cond, code, ...)        \
  do {                                     \
    if (!(cond)) {                         \
      pu::set_error(__VA_ARGS__);          \
      return (code);                       \
    }                                      \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;  // B200

// A window of a [B,Hs,Ws,C] NHWC tensor; op pixel (y,x) maps to tensor pixel (y+oy, x+ox).
struct View {
  const float* p;
  int Hs, Ws, C, oy, ox;
};
struct ViewW {
  float* p;
  int Hs, Ws, C, oy, ox;
};

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// Hot kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's launch (and the
// scheduling of its CTAs as SM resources free up) overlaps this kernel's tail instead of waiting for a full
// kernel-boundary drain (~3 us each, ~100 kernels per train step).  Protocol: every such kernel calls pdl_prologue()
// before touching global memory — griddepcontrol.wait blocks until ALL prerequisite grids have completed and flushed,
// so correctness does not depend on where the predecessor triggers; launch_dependents right after lets the successor
// start launching immediately.  Kernels launched without the attribute execute both as no-ops.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

bool pdl_enabled();  // PU_PDL=1 enables (defined in pu_api.cu)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers ----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// round-to-nearest to TF32 precision (10-bit mantissa).  tcgen05 kind::tf32 TRUNCATES fp32 operands, so every
// tensor that feeds a tensor-core conv is stored already rounded by its producer's epilogue.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// streaming (read-once) 128-bit load that does not allocate in L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace pu
