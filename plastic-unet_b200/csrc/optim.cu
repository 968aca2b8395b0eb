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

// flat[dst_off[t] + i] = src[t][i]: gathers the per-parameter gradient tensors autograd produced into the flat
// gradient arena with ONE launch (blockIdx.y = tensor), instead of one accumulate kernel per parameter.
__global__ void gather_flat_kernel(const long long* __restrict__ table, int n, float* __restrict__ flat) {
  const int t = blockIdx.y;
  if (t >= n) return;
  const float* src = reinterpret_cast<const float*>(table[3 * t]);
  const long long off = table[3 * t + 1], size = table[3 * t + 2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (long long)gridDim.x * blockDim.x)
    flat[off + i] = src[i];
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

int pu_gather_flat(const long long* table, int n, float* flat, void* stream) {
  PU_REQUIRE(table && flat && n > 0 && n <= 65535, PU_ERR_BAD_ARG, "pu_gather_flat: bad argument");
  dim3 grid(32, n);
  pu::gather_flat_kernel<<<grid, 256, 0, pu::as_stream(stream)>>>(table, n, flat);
  return pu::post_launch("pu_gather_flat");
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

int pu_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step_count, const float* lr, float beta1,
                 float beta2, float eps, float grad_scale, long long n, void* stream) {
  PU_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_count && lr && n > 0, PU_ERR_BAD_ARG, "pu_adam_step: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  pu::step_inc_kernel<<<1, 1, 0, st>>>(step_count);
  int rc = pu::post_launch("pu_adam_step inc");
  if (rc) return rc;
  int g = (int)((n + 1023) / 1024);
  g = g < 1 ? 1 : (g > 8 * pu::kNumSMs ? 8 * pu::kNumSMs : g);
  pu::adam_kernel<<<g, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, step_count, lr, beta1, beta2, eps, grad_scale, n);
  return pu::post_launch("pu_adam_step");
}

}  // extern "C"
