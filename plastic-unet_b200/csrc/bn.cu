// bn.cu — BatchNorm2d over NHWC fp32 (optional in the reference: unet_p.py:106,109; unet_p_res.py:151,175;
// off in every reference caller).  Training forward = per-channel batch statistics
// (fp32 block partials -> fp64 atomics for a stable E[x^2]-E[x]^2), apply with optional fused ReLU.
#include "pu_common.cuh"

namespace pu {

// acc[0..C) = sum x, acc[C..2C) = sum x^2 (double, pre-zeroed).  blockDim.x multiple of C.
__global__ void bn_sums_kernel(const float* __restrict__ x, double* __restrict__ acc, long long npix, int C) {
  extern __shared__ float red[];  // [2][blockDim.x]
  const int c = threadIdx.x % C, sub = threadIdx.x / C, nsub = blockDim.x / C;
  float s = 0.f, s2 = 0.f;
  for (long long p = (long long)blockIdx.x * nsub + sub; p < npix; p += (long long)gridDim.x * nsub) {
    const float v = __ldg(x + p * C + c);
    s += v;
    s2 = fmaf(v, v, s2);
  }
  red[threadIdx.x] = s;
  red[blockDim.x + threadIdx.x] = s2;
  __syncthreads();
  if (sub == 0) {
    double a = 0.0, b = 0.0;
    for (int u = 0; u < nsub; ++u) {
      a += (double)red[u * C + c];
      b += (double)red[blockDim.x + u * C + c];
    }
    atomicAdd(acc + c, a);
    atomicAdd(acc + C + c, b);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ acc, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                                   long long npix, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc[c] / (double)npix;
  double var = acc[C + c] / (double)npix - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
  if (running_var != nullptr) {
    const double unbiased = npix > 1 ? var * (double)npix / (double)(npix - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// mode 0: mean/invstd given; mode 1: running_mean/running_var given (invstd computed on the fly)
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ mean, const float* __restrict__ stat2, float* __restrict__ y, float eps,
                                long long n, int C, int flags, int mode) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float is = mode == 0 ? __ldg(stat2 + c) : rsqrtf(__ldg(stat2 + c) + eps);
    float v = (x[i] - __ldg(mean + c)) * is * __ldg(gamma + c) + __ldg(beta + c);
    if (flags & PU_FLAG_RELU) v = fmaxf(v, 0.f);
    if (flags & PU_FLAG_ROUND_TF32) v = round_tf32(v);
    y[i] = v;
  }
}

// acc[0..C) = sum g, acc[C..2C) = sum g*xhat  with g = relu ? dy*(y>0) : dy
__global__ void bn_bwd_sums_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                   const float* __restrict__ mean, const float* __restrict__ invstd, double* __restrict__ acc,
                                   long long npix, int C, int relu) {
  extern __shared__ float red[];
  const int c = threadIdx.x % C, sub = threadIdx.x / C, nsub = blockDim.x / C;
  const float m = __ldg(mean + c), is = __ldg(invstd + c);
  float s = 0.f, s2 = 0.f;
  for (long long p = (long long)blockIdx.x * nsub + sub; p < npix; p += (long long)gridDim.x * nsub) {
    float g = __ldg(dy + p * C + c);
    if (relu && !(__ldg(y + p * C + c) > 0.f)) g = 0.f;
    s += g;
    s2 = fmaf(g, (__ldg(x + p * C + c) - m) * is, s2);
  }
  red[threadIdx.x] = s;
  red[blockDim.x + threadIdx.x] = s2;
  __syncthreads();
  if (sub == 0) {
    double a = 0.0, b = 0.0;
    for (int u = 0; u < nsub; ++u) {
      a += (double)red[u * C + c];
      b += (double)red[blockDim.x + u * C + c];
    }
    atomicAdd(acc + c, a);
    atomicAdd(acc + C + c, b);
  }
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const double* __restrict__ acc, float* __restrict__ dx, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, long long npix, int C, int relu, int train) {
  const long long n = npix * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float g = dy[i];
    if (relu && !(y[i] > 0.f)) g = 0.f;
    const float is = __ldg(invstd + c), ga = __ldg(gamma + c);
    float v;
    if (train) {
      const float xhat = (x[i] - __ldg(mean + c)) * is;
      const float sg = (float)(acc[c] / (double)npix), sgx = (float)(acc[C + c] / (double)npix);
      v = ga * is * (g - sg - xhat * sgx);
    } else {
      v = ga * is * g;
    }
    dx[i] = v;
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta != nullptr) dbeta[c] = (float)acc[c];
      if (dgamma != nullptr) dgamma[c] = (float)acc[C + c];
    }
  }
}

// running-stat update from the saved batch statistics (unbiased variance recovered from invstd)
__global__ void bn_update_running_kernel(const float* __restrict__ mean, const float* __restrict__ invstd,
                                         float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                         float eps, long long npix, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double is = (double)invstd[c];
  double var = 1.0 / (is * is) - (double)eps;
  if (var < 0.0) var = 0.0;
  const double unbiased = npix > 1 ? var * (double)npix / (double)(npix - 1) : var;
  running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean[c];
  running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
}

__global__ void invstd_from_var_kernel(const float* __restrict__ var, float* __restrict__ invstd, float eps, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) invstd[c] = rsqrtf(var[c] + eps);
}

static inline int ew_grid(long long n) {
  long long g = (n + 1023) / 1024;
  if (g < 1) g = 1;
  if (g > 16LL * kNumSMs) g = 16LL * kNumSMs;
  return (int)g;
}

static int red_cfg(int C, long long npix, int* bs, int* blocks) {
  PU_REQUIRE(C <= 1024, PU_ERR_UNSUPPORTED, "bn: C=%d > 1024", C);
  *bs = C >= 256 ? C : (256 / C) * C;
  const int nsub = *bs / C;
  long long b = (npix + (long long)nsub * 32 - 1) / ((long long)nsub * 32);
  if (b < 1) b = 1;
  if (b > 4 * kNumSMs) b = 4 * kNumSMs;
  *blocks = (int)b;
  return PU_OK;
}

// eval-mode BatchNorm folded into the preceding convolution (a free fusion: no activation pass at all):
//   BN(conv_w(x) + b) = conv_{w'}(x) + b',  s = gamma / sqrt(var + eps),  w'[co] = w[co] * s[co],  b' = (b - mean) * s + beta
__global__ void bn_fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                    float* __restrict__ w_out, float* __restrict__ b_out, int Cout, int per_co) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)Cout * per_co;
  if (i < n) {
    const int co = (int)(i / per_co);
    w_out[i] = w[i] * (gamma[co] * rsqrtf(var[co] + eps));
  } else if (i < n + Cout) {
    const int co = (int)(i - n);
    const float s = gamma[co] * rsqrtf(var[co] + eps);
    b_out[co] = ((b != nullptr ? b[co] : 0.f) - mean[co]) * s + beta[co];
  }
}

}  // namespace pu

extern "C" {

// ws: caller-provided scratch of 2*C doubles (fp64 sum / sum-of-squares accumulators).
int pu_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* y, float* save_mean, float* save_invstd,
                       float* running_mean, float* running_var, double* ws, float momentum, float eps, long long npix, int C,
                       int flags, void* stream) {
  PU_REQUIRE(x && gamma && beta && y && save_mean && save_invstd && ws && npix > 0 && C > 0, PU_ERR_BAD_ARG, "pu_bn_train_fwd: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st);
  if (e != cudaSuccess) {
    pu::set_error("pu_bn_train_fwd memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  int bs, blocks;
  int rc = pu::red_cfg(C, npix, &bs, &blocks);
  if (rc) return rc;
  pu::bn_sums_kernel<<<blocks, bs, 2 * bs * sizeof(float), st>>>(x, ws, npix, C);
  rc = pu::post_launch("pu_bn_train_fwd sums");
  if (rc) return rc;
  pu::bn_finalize_kernel<<<pu::cdiv(C, 128), 128, 0, st>>>(ws, save_mean, save_invstd, running_mean, running_var, momentum, eps, npix, C);
  rc = pu::post_launch("pu_bn_train_fwd finalize");
  if (rc) return rc;
  pu::bn_apply_kernel<<<pu::ew_grid(npix * C), 256, 0, st>>>(x, gamma, beta, save_mean, save_invstd, y, eps, npix * C, C, flags, 0);
  return pu::post_launch("pu_bn_train_fwd apply");
}

int pu_bn_eval_fwd(const float* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                   float* y, float eps, long long npix, int C, int flags, void* stream) {
  PU_REQUIRE(x && gamma && beta && running_mean && running_var && y && npix > 0 && C > 0, PU_ERR_BAD_ARG, "pu_bn_eval_fwd: bad argument");
  pu::bn_apply_kernel<<<pu::ew_grid(npix * C), 256, 0, pu::as_stream(stream)>>>(x, gamma, beta, running_mean, running_var, y, eps,
                                                                               npix * C, C, flags, 1);
  return pu::post_launch("pu_bn_eval_fwd");
}

int pu_bn_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* mean, const float* invstd,
                 float* dx, float* dgamma, float* dbeta, double* ws, long long npix, int C, int relu, int train, void* stream) {
  PU_REQUIRE(x && dy && gamma && mean && invstd && dx && ws && npix > 0 && C > 0 && (y || !relu), PU_ERR_BAD_ARG, "pu_bn_bwd: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st);
  if (e != cudaSuccess) {
    pu::set_error("pu_bn_bwd memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  int bs, blocks;
  int rc = pu::red_cfg(C, npix, &bs, &blocks);
  if (rc) return rc;
  pu::bn_bwd_sums_kernel<<<blocks, bs, 2 * bs * sizeof(float), st>>>(x, y, dy, mean, invstd, ws, npix, C, relu);
  rc = pu::post_launch("pu_bn_bwd sums");
  if (rc) return rc;
  pu::bn_bwd_apply_kernel<<<pu::ew_grid(npix * C), 256, 0, st>>>(x, y, dy, gamma, mean, invstd, ws, dx, dgamma, dbeta, npix, C, relu, train);
  return pu::post_launch("pu_bn_bwd apply");
}

int pu_bn_update_running(const float* mean, const float* invstd, float* running_mean, float* running_var, float momentum,
                         float eps, long long npix, int C, void* stream) {
  PU_REQUIRE(mean && invstd && running_mean && running_var && npix > 0 && C > 0, PU_ERR_BAD_ARG, "pu_bn_update_running: bad argument");
  pu::bn_update_running_kernel<<<pu::cdiv(C, 128), 128, 0, pu::as_stream(stream)>>>(mean, invstd, running_mean, running_var, momentum,
                                                                                eps, npix, C);
  return pu::post_launch("pu_bn_update_running");
}

int pu_bn_fold_conv(const float* w, const float* b, const float* gamma, const float* beta, const float* running_mean,
                    const float* running_var, float eps, float* w_out, float* b_out, int Cout, int per_co, void* stream) {
  PU_REQUIRE(w && gamma && beta && running_mean && running_var && w_out && b_out && Cout > 0 && per_co > 0, PU_ERR_BAD_ARG,
             "pu_bn_fold_conv: bad argument");
  const long long n = (long long)Cout * per_co;
  pu::bn_fold_conv_kernel<<<pu::cdiv(n + Cout, 256), 256, 0, pu::as_stream(stream)>>>(w, b, gamma, beta, running_mean, running_var, eps, w_out,
                                                                                      b_out, Cout, per_co);
  return pu::post_launch("pu_bn_fold_conv");
}

int pu_bn_invstd(const float* running_var, float* invstd, float eps, int C, void* stream) {
  PU_REQUIRE(running_var && invstd && C > 0, PU_ERR_BAD_ARG, "pu_bn_invstd: bad argument");
  pu::invstd_from_var_kernel<<<pu::cdiv(C, 128), 128, 0, pu::as_stream(stream)>>>(running_var, invstd, eps, C);
  return pu::post_launch("pu_bn_invstd");
}

}  // extern "C"
