// optim.cu — train-step tail: BCE loss + its gradient in one pass, and a
// single-launch Adam over the flat parameter arena.  reference train.py:66-70,101-112
// (nn.BCELoss, torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay).
#include "pu_common.cuh"

namespace pu {

// loss += sum_i -(t*max(log s,-100) + (1-t)*max(log(1-s),-100)) / n ;  gS = (s-t)/max(s(1-s),1e-12)/n
__global__ void bce_kernel(const float* __restrict__ S, const float* __restrict__ T, float* __restrict__ loss, float* __restrict__ gS,
                           long long n) {
  __shared__ float red[8];
  const float inv_n = 1.f / (float)n;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = S[i], t = T[i];
    const float l1 = fmaxf(logf(s), -100.f), l0 = fmaxf(log1pf(-s), -100.f);
    acc -= t * l1 + (1.f - t) * l0;
    if (gS != nullptr) gS[i] = (s - t) / fmaxf(s * (1.f - s), 1e-12f) * inv_n;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int u = 0; u < (int)blockDim.x / 32; ++u) s += red[u];
    atomicAdd(loss, s * inv_n);
  }
}

__global__ void step_inc_kernel(float* step) { *step += 1.f; }

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float b1, float b2, float eps, float gscale,
                                         float bc2_sqrt, float step_size) {
  const float gi = g * gscale;
  const float mi = b1 * m + (1.f - b1) * gi;
  const float vi = b2 * v + (1.f - b2) * gi * gi;
  m = mi;
  v = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p -= step_size * (mi / denom);
}

// 128-bit form (16-byte aligned arenas, n4 = n / 4): the whole UNetp arena (66 k float4) is ONE round of loads for 66 k threads —
// the scalar grid-stride loop below took four dependent rounds at the very end of every step.  Same arithmetic per element.
__global__ void adam_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                 const float* __restrict__ step_p, const float* __restrict__ lr_p, float b1, float b2, float eps,
                                 float gscale, long long n4) {
  const float step = __ldg(step_p);
  const float lr = __ldg(lr_p);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pi = p[i], mi = m[i], vi = v[i];
    const float4 gi = g[i];
    adam_one(pi.x, gi.x, mi.x, vi.x, b1, b2, eps, gscale, bc2_sqrt, step_size);
    adam_one(pi.y, gi.y, mi.y, vi.y, b1, b2, eps, gscale, bc2_sqrt, step_size);
    adam_one(pi.z, gi.z, mi.z, vi.z, b1, b2, eps, gscale, bc2_sqrt, step_size);
    adam_one(pi.w, gi.w, mi.w, vi.w, b1, b2, eps, gscale, bc2_sqrt, step_size);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ step_p, const float* __restrict__ lr_p, float b1, float b2, float eps,
                            float gscale, long long n) {
  const float step = __ldg(step_p);
  const float lr = __ldg(lr_p);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}


// ---- fused gradient exchange + Adam over NVLink peer memory (data parallel) ------------------------------------------------
// Every rank's flat gradient arena lives in symmetric (peer-mapped) memory.  ONE kernel per step and rank:
//   1. block-wise barrier with the same block of every peer ("my gradients are complete": flags in symmetric memory,
//      system-scope release / acquire);
//   2. g[i] = sum over ranks j = 0..W-1 of peer_grad[j][i], read straight from the peers' HBM over NVLink with 128-bit loads
//      (all W loads of an element in flight together), in the SAME order on every rank => bit-identical sums, replicas
//      stay bit-identical; then the Adam update of the local parameters — the all-reduced gradient never exists in memory;
//   3. a second barrier ("I have finished reading yours") before anyone may overwrite its arena in the next step.
// Replaces ncclAllReduce(1.06 MB) + adam_kernel: the exchange is latency-bound (one-shot: one NVLink round trip instead of a
// ring / tree schedule plus a separate kernel).  All spins are bounded (trap, never hang).
constexpr int kArThreads = 512;

__device__ __forceinline__ void ar_put(int* flag) {  // set a peer's flag 0 -> 1
  unsigned spins = 0;
  while (true) {
    int old;
    asm volatile("atom.release.sys.global.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(flag) : "memory");
    if (old == 0) return;
    if (++spins > (1u << 26)) {
      printf("pu adam_allreduce: peer flag still set (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void ar_wait(int* flag) {  // wait for my flag to become 1, reset it to 0
  unsigned spins = 0;
  while (true) {
    int old;
    asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(flag) : "memory");
    if (old == 1) return;
    if (++spins > (1u << 26)) {
      printf("pu adam_allreduce: timed out waiting for a peer (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      asm volatile("trap;");
    }
  }
}
// flags layout (per rank, symmetric): [phase 2][block][world] ints, zero-initialised
__device__ __forceinline__ void ar_barrier(const long long* __restrict__ flag_ptrs, int rank, int world, int phase) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    int* theirs = reinterpret_cast<int*>(flag_ptrs[peer]) + ((size_t)phase * gridDim.x + blockIdx.x) * world + rank;
    int* mine = reinterpret_cast<int*>(flag_ptrs[rank]) + ((size_t)phase * gridDim.x + blockIdx.x) * world + peer;
    ar_put(theirs);
    ar_wait(mine);
  }
  __syncthreads();
}

__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}

template <int W>
__global__ void __launch_bounds__(kArThreads) adam_allreduce_kernel(float* __restrict__ p, const long long* __restrict__ grad_ptrs,
                                                                    const long long* __restrict__ flag_ptrs, int rank, float* __restrict__ m,
                                                                    float* __restrict__ v, const float* __restrict__ step_p,
                                                                    const float* __restrict__ lr_p, float b1, float b2, float eps, float gscale,
                                                                    long long n4, float* __restrict__ extra_out, long long n4_extra) {
  ar_barrier(flag_ptrs, rank, W, 0);
  const float* g[W];
#pragma unroll
  for (int j = 0; j < W; ++j) g[j] = reinterpret_cast<const float*>(grad_ptrs[j]);
  const float step = __ldg(step_p);
  const float lr = __ldg(lr_p);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 part[W];
#pragma unroll
    for (int j = 0; j < W; ++j) part[j] = ld_peer4(g[j] + 4 * i);
    float4 s = part[0];
#pragma unroll
    for (int j = 1; j < W; ++j) { s.x += part[j].x; s.y += part[j].y; s.z += part[j].z; s.w += part[j].w; }
    float gi[4] = {s.x * gscale, s.y * gscale, s.z * gscale, s.w * gscale};
    float4 pm = *reinterpret_cast<float4*>(m + 4 * i), pv = *reinterpret_cast<float4*>(v + 4 * i), pp = *reinterpret_cast<float4*>(p + 4 * i);
    float mm[4] = {pm.x, pm.y, pm.z, pm.w}, vv[4] = {pv.x, pv.y, pv.z, pv.w}, ww[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mm[e] = b1 * mm[e] + (1.f - b1) * gi[e];
      vv[e] = b2 * vv[e] + (1.f - b2) * gi[e] * gi[e];
      const float denom = sqrtf(vv[e]) / bc2_sqrt + eps;
      ww[e] -= step_size * (mm[e] / denom);
    }
    *reinterpret_cast<float4*>(m + 4 * i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + 4 * i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *reinterpret_cast<float4*>(p + 4 * i) = make_float4(ww[0], ww[1], ww[2], ww[3]);
  }
  // the floats that follow the gradients in every rank's arena (the plastic-trace delta of the step) are only summed — same order
  // on every rank — and handed back: the trace all-reduce rides on the same NVLink round trip and the same two barriers
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_extra; i += (long long)gridDim.x * blockDim.x) {
    float4 s = ld_peer4(g[0] + 4 * (n4 + i));
#pragma unroll
    for (int j = 1; j < W; ++j) {
      const float4 t = ld_peer4(g[j] + 4 * (n4 + i));
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(extra_out + 4 * i) = s;
  }
  ar_barrier(flag_ptrs, rank, W, 1);
}

// flat[dst_off[t] + i] = src[t][i]: gathers the per-parameter gradient tensors autograd produced into the flat
// gradient arena with ONE launch (blockIdx.y = tensor), instead of one accumulate kernel per parameter.
__global__ void gather_flat_kernel(const long long* __restrict__ table, int n, float* __restrict__ flat, float* step_inc) {
  // step_inc (may be null): the optimizer's step counter, incremented here so that Adam needs no launch of its own for it
  if (step_inc != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *step_inc += 1.f;
  const int t = blockIdx.y;
  if (t >= n) return;
  const float* src = reinterpret_cast<const float*>(table[3 * t]);
  const long long off = table[3 * t + 1], size = table[3 * t + 2];
  float* dst = flat + off;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {  // 128-bit body + scalar tail
    const long long n4 = size >> 2;
    for (long long i = i0; i < n4; i += stride) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
    for (long long i = 4 * n4 + i0; i < size; i += stride) dst[i] = src[i];
  } else {
    for (long long i = i0; i < size; i += stride) dst[i] = src[i];
  }
}

// Single-GPU training step: Adam straight from the per-parameter gradient tensors (no gather launch, no pass through the flat
// gradient arena, no step-counter launch).  blockIdx.y = tensor; arena elements that no table row covers (parameters without a
// gradient, alignment gaps) have g = m = v = 0 and would not move under adam_kernel either.  Every block reads the OLD step
// count and uses step + 1; the last block to finish (ticket counter, re-zeroed for the next launch) writes it back.
__device__ unsigned int g_adam_ticket = 0;  // launches that share it must be stream-ordered (the steps of a TrainStep are)

__global__ void adam_table_kernel(const long long* __restrict__ table, int n, float* __restrict__ p, float* __restrict__ m,
                                  float* __restrict__ v, float* step_p, const float* __restrict__ lr_p, float b1, float b2, float eps,
                                  float gscale) {
  const int t = blockIdx.y;
  const float step = *reinterpret_cast<volatile const float*>(step_p) + 1.f;
  if (t < n) {
    const float* __restrict__ g = reinterpret_cast<const float*>(table[3 * t]);
    const long long off = table[3 * t + 1], size = table[3 * t + 2];
    const float lr = __ldg(lr_p);
    const float bc1 = 1.f - powf(b1, step);
    const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
    const float step_size = lr / bc1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (long long)gridDim.x * blockDim.x) {
      const float gi = g[i] * gscale;
      const float mi = b1 * m[off + i] + (1.f - b1) * gi;
      const float vi = b2 * v[off + i] + (1.f - b2) * gi * gi;
      m[off + i] = mi;
      v[off + i] = vi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      p[off + i] -= step_size * (mi / denom);
    }
  }
  __syncthreads();  // every thread of the block has read the step count
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    if (atomicAdd(&g_adam_ticket, 1u) == total - 1) {
      *step_p = step;
      g_adam_ticket = 0u;
    }
  }
}

// Two device-to-device copies in ONE launch (the step's image and mask batches into its static buffers): two cudaMemcpyAsync
// calls cost two launch gaps in front of every graph replay (3.1 + 2.8 us of copies over 7.5 us of stream time).
__global__ void copy2_kernel(const float4* __restrict__ s0, float4* __restrict__ d0, long long n0, const float4* __restrict__ s1,
                             float4* __restrict__ d1, long long n1) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += stride) {
    if (i < n0) d0[i] = s0[i];
    else d1[i - n0] = s1[i - n0];
  }
}

// dst[b] = src[idx[b]] placed at (oy, ox) of a zero-filled Hd x Wd canvas: the device-resident dataset's batch assembly +
// the 101 -> 128 zero padding of BASELINE configs[0..1] in one pass (reference train.py:94-95 converts and copies one
// image per step from host numpy arrays; utils/data_set.py:43-44 holds them as float64 [n, 1, 101, 101]).
__global__ void gather_pad_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ dst, int B,
                                  long long plane_src, int Hs, int Ws, int Hd, int Wd, int oy, int ox, int planes) {
  const long long n = (long long)B * planes * Hd * Wd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wd);
    long long r = i / Wd;
    const int y = (int)(r % Hd);
    r /= Hd;
    const int c = (int)(r % planes);
    const int b = (int)(r / planes);
    const int sy = y - oy, sx = x - ox;
    float v = 0.f;
    if (sy >= 0 && sy < Hs && sx >= 0 && sx < Ws) v = __ldg(src + (idx[b] * planes + c) * plane_src + (long long)sy * Ws + sx);
    dst[i] = v;
  }
}

}  // namespace pu

extern "C" {

int pu_bce_fwd_bwd(const float* S, const float* T, float* loss, float* gS, long long n, void* stream) {
  PU_REQUIRE(S && T && loss && n > 0, PU_ERR_BAD_ARG, "pu_bce_fwd_bwd: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (e != cudaSuccess) {
    pu::set_error("pu_bce_fwd_bwd memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  int g = (int)((n + 1023) / 1024);
  g = g < 1 ? 1 : (g > 4 * pu::kNumSMs ? 4 * pu::kNumSMs : g);
  pu::bce_kernel<<<g, 256, 0, st>>>(S, T, loss, gS, n);
  return pu::post_launch("pu_bce_fwd_bwd");
}

int pu_copy2(const float* src0, float* dst0, long long n0, const float* src1, float* dst1, long long n1, void* stream) {
  PU_REQUIRE(src0 && dst0 && src1 && dst1 && n0 > 0 && n1 > 0 && n0 % 4 == 0 && n1 % 4 == 0, PU_ERR_BAD_ARG,
             "pu_copy2: bad argument (element counts must be multiples of 4)");
  PU_REQUIRE(pu::aligned16(src0) && pu::aligned16(dst0) && pu::aligned16(src1) && pu::aligned16(dst1), PU_ERR_BAD_ARG,
             "pu_copy2: pointers must be 16-byte aligned");
  const long long n4 = (n0 + n1) / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 16LL * pu::kNumSMs) blocks = 16LL * pu::kNumSMs;
  pu::copy2_kernel<<<(unsigned)blocks, 256, 0, pu::as_stream(stream)>>>(reinterpret_cast<const float4*>(src0), reinterpret_cast<float4*>(dst0),
                                                                        n0 / 4, reinterpret_cast<const float4*>(src1),
                                                                        reinterpret_cast<float4*>(dst1), n1 / 4);
  return pu::post_launch("pu_copy2");
}

int pu_gather_flat(const long long* table, int n, float* flat, void* stream) {
  PU_REQUIRE(table && flat && n > 0 && n <= 65535, PU_ERR_BAD_ARG, "pu_gather_flat: bad argument");
  dim3 grid(32, n);
  pu::gather_flat_kernel<<<grid, 256, 0, pu::as_stream(stream)>>>(table, n, flat, nullptr);
  return pu::post_launch("pu_gather_flat");
}

int pu_gather_flat_inc(const long long* table, int n, float* flat, float* step_count, void* stream) {
  PU_REQUIRE(table && flat && step_count && n > 0 && n <= 65535, PU_ERR_BAD_ARG, "pu_gather_flat_inc: bad argument");
  dim3 grid(32, n);
  pu::gather_flat_kernel<<<grid, 256, 0, pu::as_stream(stream)>>>(table, n, flat, step_count);
  return pu::post_launch("pu_gather_flat_inc");
}

// One block per SM with 512 threads: 75.8 k threads cover the 66 k float4 of the UNetp arena in ONE pass, i.e. one NVLink round trip
// for all peer loads (with 32 blocks the four dependent passes made the kernel 15 us slower than ncclAllReduce + adam at 4 GPUs).
// Block b only ever waits for block b of its peers, and nothing else runs on the GPU at this point of the step.
int pu_adam_allreduce_blocks(void) { return pu::kNumSMs; }

int pu_adam_allreduce_step(float* param, const long long* peer_grad_ptrs, const long long* peer_flag_ptrs, int rank, int world, float* exp_avg,
                           float* exp_avg_sq, float* step_count, const float* lr, float beta1, float beta2, float eps, float grad_scale,
                           long long n, float* extra_sum, long long n_extra, void* stream) {
  PU_REQUIRE(param && peer_grad_ptrs && peer_flag_ptrs && exp_avg && exp_avg_sq && step_count && lr && n > 0, PU_ERR_BAD_ARG,
             "pu_adam_allreduce_step: bad argument");
  PU_REQUIRE(n_extra >= 0 && n_extra % 4 == 0 && (n_extra == 0 || (extra_sum != nullptr && pu::aligned16(extra_sum))), PU_ERR_BAD_ARG,
             "pu_adam_allreduce_step: extra_sum must be 16-byte aligned and n_extra a multiple of 4");
  PU_REQUIRE(n % 4 == 0 && pu::aligned16(param) && pu::aligned16(exp_avg) && pu::aligned16(exp_avg_sq), PU_ERR_BAD_ARG,
             "pu_adam_allreduce_step: the arenas must be 16-byte aligned and a multiple of 4 floats long");
  PU_REQUIRE(rank >= 0 && rank < world, PU_ERR_BAD_ARG, "pu_adam_allreduce_step: rank %d outside world %d", rank, world);
  cudaStream_t st = pu::as_stream(stream);
  pu::step_inc_kernel<<<1, 1, 0, st>>>(step_count);
  int rc = pu::post_launch("pu_adam_allreduce_step inc");
  if (rc) return rc;
  const int grid = pu_adam_allreduce_blocks();  // the same on every rank: block b pairs up with block b of every peer
  const long long n4 = n / 4;
#define PU_AR_LAUNCH(W_)                                                                                                                  \
  pu::adam_allreduce_kernel<W_><<<grid, pu::kArThreads, 0, st>>>(param, peer_grad_ptrs, peer_flag_ptrs, rank, exp_avg, exp_avg_sq, step_count, \
                                                                  lr, beta1, beta2, eps, grad_scale, n4, extra_sum, n_extra / 4)
  switch (world) {
    case 2: PU_AR_LAUNCH(2); break;
    case 4: PU_AR_LAUNCH(4); break;
    case 8: PU_AR_LAUNCH(8); break;
    default:
      pu::set_error("pu_adam_allreduce_step: world size %d not supported (2, 4, 8)", world);
      return PU_ERR_UNSUPPORTED;
  }
#undef PU_AR_LAUNCH
  return pu::post_launch("pu_adam_allreduce_step");
}

int pu_gather_pad(const float* src, const long long* idx, float* dst, int B, int planes, int Hs, int Ws, int Hd, int Wd, int oy, int ox,
                  void* stream) {
  PU_REQUIRE(src && idx && dst && B > 0 && planes > 0 && Hs > 0 && Ws > 0 && oy >= 0 && ox >= 0 && oy + Hs <= Hd && ox + Ws <= Wd,
             PU_ERR_BAD_ARG, "pu_gather_pad: bad argument (the %dx%d source must fit at (%d,%d) of the %dx%d canvas)", Hs, Ws, oy, ox, Hd, Wd);
  const long long n = (long long)B * planes * Hd * Wd;
  int g = (int)((n + 1023) / 1024);
  g = g < 1 ? 1 : (g > 8 * pu::kNumSMs ? 8 * pu::kNumSMs : g);
  pu::gather_pad_kernel<<<g, 256, 0, pu::as_stream(stream)>>>(src, idx, dst, B, (long long)Hs * Ws, Hs, Ws, Hd, Wd, oy, ox, planes);
  return pu::post_launch("pu_gather_pad");
}

int pu_adam_table_step(const long long* table, int n, float* param, float* exp_avg, float* exp_avg_sq, float* step_count, const float* lr,
                       float beta1, float beta2, float eps, float grad_scale, void* stream) {
  PU_REQUIRE(table && param && exp_avg && exp_avg_sq && step_count && lr && n > 0 && n <= 65535, PU_ERR_BAD_ARG,
             "pu_adam_table_step: bad argument");
  dim3 grid(64, n);
  pu::adam_table_kernel<<<grid, 256, 0, pu::as_stream(stream)>>>(table, n, param, exp_avg, exp_avg_sq, step_count, lr, beta1, beta2, eps,
                                                                 grad_scale);
  return pu::post_launch("pu_adam_table_step");
}

static int adam_step_impl(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step_count, const float* lr, float beta1,
                          float beta2, float eps, float grad_scale, long long n, bool inc, void* stream);

int pu_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step_count, const float* lr, float beta1,
                 float beta2, float eps, float grad_scale, long long n, void* stream) {
  return adam_step_impl(param, grad, exp_avg, exp_avg_sq, step_count, lr, beta1, beta2, eps, grad_scale, n, true, stream);
}

int pu_adam_step_counted(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* step_count, const float* lr,
                         float beta1, float beta2, float eps, float grad_scale, long long n, void* stream) {
  return adam_step_impl(param, grad, exp_avg, exp_avg_sq, const_cast<float*>(step_count), lr, beta1, beta2, eps, grad_scale, n, false, stream);
}

static int adam_step_impl(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step_count, const float* lr, float beta1,
                          float beta2, float eps, float grad_scale, long long n, bool inc, void* stream) {
  PU_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_count && lr && n > 0, PU_ERR_BAD_ARG, "pu_adam_step: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  int rc = PU_OK;
  if (inc) {
    pu::step_inc_kernel<<<1, 1, 0, st>>>(step_count);
    rc = pu::post_launch("pu_adam_step inc");
    if (rc) return rc;
  }
  if (n % 4 == 0 && pu::aligned16(param) && pu::aligned16(grad) && pu::aligned16(exp_avg) && pu::aligned16(exp_avg_sq)) {
    const long long n4 = n / 4;
    long long g4 = (n4 + 255) / 256;
    if (g4 > 8LL * pu::kNumSMs) g4 = 8LL * pu::kNumSMs;
    pu::adam_vec4_kernel<<<(unsigned)g4, 256, 0, st>>>(reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad),
                                                       reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), step_count,
                                                       lr, beta1, beta2, eps, grad_scale, n4);
    return pu::post_launch("pu_adam_step");
  }
  int g = (int)((n + 1023) / 1024);
  g = g < 1 ? 1 : (g > 8 * pu::kNumSMs ? 8 * pu::kNumSMs : g);
  pu::adam_kernel<<<g, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, step_count, lr, beta1, beta2, eps, grad_scale, n);
  return pu::post_launch("pu_adam_step");
}

}  // extern "C"
