// pool.cu — MaxPool2d(2) (+ fused Dropout2d channel scale) and bilinear x2 upsampling, NHWC fp32.
// HBM-bound elementwise kernels: one float4 (4 channels) per thread access, grid-stride.
// reference unet_p.py:139,153; unet_p_res.py:240-253.
#include "pu_common.cuh"

namespace pu {

// V = 4 (float4 path, C % 4 == 0) or 1 (scalar)
// code (optional): one byte per pooled element — bits 0-1 = position of the maximum inside the window (ATen's tie-break: the
// first element, row-major, that is strictly greater or NaN), bit 2 = (maximum > 0).  The backward pass then routes the
// gradient from the code alone instead of re-reading the four inputs of every window (a third of its HBM traffic).
template <int V>
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, float* __restrict__ y,
                                    unsigned char* __restrict__ code, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, CV = C / V;
  const long long n = (long long)B * Ho * Wo * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long p = i / CV;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const float* xp = x + (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + cv * V;
    float m[V];
    int arg[V];
#pragma unroll
    for (int u = 0; u < V; ++u) { m[u] = -INFINITY; arg[u] = 0; }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const float* q = xp + ((size_t)dy * W + dx) * C;
        if (V == 4) {
          const float4 v = ldg4(q);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int u = 0; u < V; ++u)
            if (vv[u] > m[u] || vv[u] != vv[u]) { m[u] = vv[u]; arg[u] = 2 * dy + dx; }
        } else {
          const float v = __ldg(q);
          if (v > m[0] || v != v) { m[0] = v; arg[0] = 2 * dy + dx; }
        }
      }
    if (code != nullptr) {
      if (V == 4) {
        uint32_t c4 = 0;
#pragma unroll
        for (int u = 0; u < V; ++u) c4 |= (uint32_t)(arg[u] | (m[u] > 0.f ? 4 : 0)) << (8 * u);
        *reinterpret_cast<uint32_t*>(code + i * V) = c4;
      } else {
        code[i] = (unsigned char)(arg[0] | (m[0] > 0.f ? 4 : 0));
      }
    }
    if (scale != nullptr) {
#pragma unroll
      for (int u = 0; u < V; ++u) m[u] *= __ldg(scale + (size_t)b * C + cv * V + u);
    }
    float* yp = y + i * V;
    if (V == 4) *reinterpret_cast<float4*>(yp) = make_float4(m[0], m[1 % V], m[2 % V], m[3 % V]);
    else yp[0] = m[0];
  }
}

// one thread per pooling window (and channel vector): recompute the arg-max with ATen's tie-break
// (first element in (dy,dx) row-major scan that is strictly greater / NaN) and route dy*scale there.
template <int V>
__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, const unsigned char* __restrict__ code, const float* __restrict__ scale,
                                    const float* __restrict__ dy, const float* __restrict__ acc, float* __restrict__ dx, int B, int H, int W,
                                    int C, int mask_in) {
  const int Ho = H / 2, Wo = W / 2, CV = C / V;
  const long long n = (long long)B * Ho * Wo * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long p = i / CV;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const size_t base = (((size_t)b * H + 2 * oy) * W + 2 * ox) * C + cv * V;
    float m[V];
    int arg[V];
    bool pos[V];  // the routed-to input is > 0 (mask_in: its ReLU mask)
    if (code != nullptr) {  // the forward pass recorded where the maximum sits: no re-read of x
      uint32_t c4;
      if (V == 4) c4 = __ldg(reinterpret_cast<const uint32_t*>(code + i * V));
      else c4 = code[i];
#pragma unroll
      for (int u = 0; u < V; ++u) {
        arg[u] = (int)((c4 >> (8 * u)) & 3u);
        pos[u] = ((c4 >> (8 * u)) & 4u) != 0;
      }
    } else {
#pragma unroll
      for (int u = 0; u < V; ++u) { m[u] = -INFINITY; arg[u] = 0; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* q = x + base + ((size_t)(k >> 1) * W + (k & 1)) * C;
        float vals[V];
        if (V == 4) {
          const float4 v = ldg4(q);
          vals[0] = v.x; vals[1 % V] = v.y; vals[2 % V] = v.z; vals[3 % V] = v.w;
        } else {
          vals[0] = __ldg(q);
        }
#pragma unroll
        for (int u = 0; u < V; ++u)
          if (vals[u] > m[u] || vals[u] != vals[u]) { m[u] = vals[u]; arg[u] = k; }
      }
#pragma unroll
      for (int u = 0; u < V; ++u) pos[u] = m[u] > 0.f;
    }
    float g[V];
    if (V == 4) {
      const float4 v = ldg4(dy + i * V);
      g[0] = v.x; g[1 % V] = v.y; g[2 % V] = v.z; g[3 % V] = v.w;
    } else {
      g[0] = __ldg(dy + i);
    }
    if (scale != nullptr) {
#pragma unroll
      for (int u = 0; u < V; ++u) g[u] *= __ldg(scale + (size_t)b * C + cv * V + u);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[V];
#pragma unroll
      for (int u = 0; u < V; ++u) o[u] = (arg[u] == k && !(mask_in && !pos[u])) ? g[u] : 0.f;
      const size_t qoff = base + ((size_t)(k >> 1) * W + (k & 1)) * C;
      if (acc != nullptr) {  // the other gradient of x (skip connection), accumulated here instead of by a separate add pass
        if (V == 4) {
          const float4 a = ldg4(acc + qoff);
          o[0] += a.x; o[1 % V] += a.y; o[2 % V] += a.z; o[3 % V] += a.w;
        } else {
          o[0] += __ldg(acc + qoff);
        }
      }
      float* q = dx + qoff;
      if (V == 4) *reinterpret_cast<float4*>(q) = make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]);
      else q[0] = o[0];
    }
  }
}

// bilinear x2, align_corners=True: src = dst * (in-1)/(out-1)
__global__ void bilinear2x_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W;
  const float ry = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float rx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const long long n = (long long)B * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const float sy = ry * oy, sx = rx * ox;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = sy - y0, lx = sx - x0;
    const float* xb = x + (size_t)b * H * W * C + c;
    const float v00 = __ldg(xb + ((size_t)y0 * W + x0) * C), v01 = __ldg(xb + ((size_t)y0 * W + x1) * C);
    const float v10 = __ldg(xb + ((size_t)y1 * W + x0) * C), v11 = __ldg(xb + ((size_t)y1 * W + x1) * C);
    y[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

__global__ void bilinear2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W;
  const float ry = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float rx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const long long n = (long long)B * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const float sy = ry * oy, sx = rx * ox;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = sy - y0, lx = sx - x0;
    const float g = dy[i];
    float* db = dx + (size_t)b * H * W * C + c;
    atomicAdd(db + ((size_t)y0 * W + x0) * C, (1.f - ly) * (1.f - lx) * g);
    atomicAdd(db + ((size_t)y0 * W + x1) * C, (1.f - ly) * lx * g);
    atomicAdd(db + ((size_t)y1 * W + x0) * C, ly * (1.f - lx) * g);
    atomicAdd(db + ((size_t)y1 * W + x1) * C, ly * lx * g);
  }
}

static inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  if (g < 1) g = 1;
  if (g > 16LL * kNumSMs) g = 16LL * kNumSMs;
  return (int)g;
}

// ---- zero insertion (stride-2 transposed convolution as a stride-1 convolution) --------------------------------------
// z[b, r, s, :] = x[b, (r+oy-1)/2, (s+ox-1)/2, :] where r+oy and s+ox are odd (and the source pixel exists), else 0:
// the (oy, ox, Ho, Wo) window of the (2H+1) x (2W+1) canvas with x at the odd coordinates.  A 3x3 stride-1 pad-1
// convolution of that canvas with the flipped kernel equals ConvTranspose2d(k=3, s=2, p=0) (reference unet_p_res.py:207).
__global__ void zero_insert2x_kernel(const float4* __restrict__ x, float4* __restrict__ z, int B, int H, int W, int C4, int Ho, int Wo,
                                     int oy, int ox) {
  const long long n = (long long)B * Ho * Wo * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long t = i / C4;
    const int s = (int)(t % Wo);
    t /= Wo;
    const int r = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int fy = r + oy, fx = s + ox;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((fy & 1) && (fx & 1)) {
      const int iy = (fy - 1) >> 1, ix = (fx - 1) >> 1;
      if (iy < H && ix < W) v = __ldg(x + (((size_t)b * H + iy) * W + ix) * C4 + c);
    }
    z[i] = v;
  }
}

__global__ void zero_insert2x_bwd_kernel(const float4* __restrict__ dz, float4* __restrict__ dx, int B, int H, int W, int C4, int Ho,
                                         int Wo, int oy, int ox) {
  const long long n = (long long)B * H * W * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long t = i / C4;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H);
    const int b = (int)(t / H);
    const int r = 2 * iy + 1 - oy, s = 2 * ix + 1 - ox;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0 && r < Ho && s >= 0 && s < Wo) v = __ldg(dz + (((size_t)b * Ho + r) * Wo + s) * C4 + c);
    dx[i] = v;
  }
}

}  // namespace pu

extern "C" {

int pu_zero_insert2x_fwd(const float* x, float* z, int B, int H, int W, int C, int Ho, int Wo, int oy, int ox, void* stream) {
  PU_REQUIRE(x && z && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && Ho > 0 && Wo > 0 && oy >= 0 && ox >= 0 && oy + Ho <= 2 * H + 1 &&
                 ox + Wo <= 2 * W + 1,
             PU_ERR_BAD_ARG, "pu_zero_insert2x_fwd: bad argument");
  PU_REQUIRE(pu::aligned16(x) && pu::aligned16(z), PU_ERR_BAD_ARG, "pu_zero_insert2x_fwd: pointers not 16-byte aligned");
  const long long n = (long long)B * Ho * Wo * (C / 4);
  long long blocks = (n + 255) / 256;
  if (blocks > 16LL * pu::kNumSMs) blocks = 16LL * pu::kNumSMs;
  pu::zero_insert2x_kernel<<<(unsigned)blocks, 256, 0, pu::as_stream(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(z),
                                                                              B, H, W, C / 4, Ho, Wo, oy, ox);
  return pu::post_launch("pu_zero_insert2x_fwd");
}

int pu_zero_insert2x_bwd(const float* dz, float* dx, int B, int H, int W, int C, int Ho, int Wo, int oy, int ox, void* stream) {
  PU_REQUIRE(dz && dx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && Ho > 0 && Wo > 0 && oy >= 0 && ox >= 0, PU_ERR_BAD_ARG,
             "pu_zero_insert2x_bwd: bad argument");
  PU_REQUIRE(pu::aligned16(dz) && pu::aligned16(dx), PU_ERR_BAD_ARG, "pu_zero_insert2x_bwd: pointers not 16-byte aligned");
  const long long n = (long long)B * H * W * (C / 4);
  long long blocks = (n + 255) / 256;
  if (blocks > 16LL * pu::kNumSMs) blocks = 16LL * pu::kNumSMs;
  pu::zero_insert2x_bwd_kernel<<<(unsigned)blocks, 256, 0, pu::as_stream(stream)>>>(reinterpret_cast<const float4*>(dz),
                                                                                  reinterpret_cast<float4*>(dx), B, H, W, C / 4, Ho, Wo, oy, ox);
  return pu::post_launch("pu_zero_insert2x_bwd");
}


static int maxpool2_fwd_impl(const float* x, const float* chan_scale, float* y, unsigned char* code, int B, int H, int W, int C,
                             void* stream) {
  PU_REQUIRE(x && y && B > 0 && H >= 2 && W >= 2 && C > 0, PU_ERR_BAD_ARG, "pu_maxpool2_fwd: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  const long long nwin = (long long)B * (H / 2) * (W / 2);
  if (C % 4 == 0 && pu::aligned16(x) && pu::aligned16(y) && (reinterpret_cast<uintptr_t>(code) & 3u) == 0)
    pu::maxpool2_fwd_kernel<4><<<pu::grid_for(nwin * (C / 4)), 256, 0, st>>>(x, chan_scale, y, code, B, H, W, C);
  else
    pu::maxpool2_fwd_kernel<1><<<pu::grid_for(nwin * C), 256, 0, st>>>(x, chan_scale, y, code, B, H, W, C);
  return pu::post_launch("pu_maxpool2_fwd");
}

int pu_maxpool2_fwd(const float* x, const float* chan_scale, float* y, int B, int H, int W, int C, void* stream) {
  return maxpool2_fwd_impl(x, chan_scale, y, nullptr, B, H, W, C, stream);
}

int pu_maxpool2_fwd_code(const float* x, const float* chan_scale, float* y, unsigned char* code, int B, int H, int W, int C, void* stream) {
  PU_REQUIRE(code != nullptr, PU_ERR_BAD_ARG, "pu_maxpool2_fwd_code: code is NULL");
  return maxpool2_fwd_impl(x, chan_scale, y, code, B, H, W, C, stream);
}

static int maxpool2_bwd_impl(const float* x, const unsigned char* code, const float* chan_scale, const float* dy, const float* acc, float* dx,
                             int B, int H, int W, int C, int flags, void* stream);

int pu_maxpool2_bwd(const float* x, const float* chan_scale, const float* dy, const float* acc, float* dx, int B, int H, int W, int C,
                    int flags, void* stream) {
  PU_REQUIRE(x != nullptr, PU_ERR_BAD_ARG, "pu_maxpool2_bwd: bad argument");
  return maxpool2_bwd_impl(x, nullptr, chan_scale, dy, acc, dx, B, H, W, C, flags, stream);
}

int pu_maxpool2_bwd_code(const unsigned char* code, const float* chan_scale, const float* dy, const float* acc, float* dx, int B, int H,
                         int W, int C, int flags, void* stream) {
  PU_REQUIRE(code != nullptr, PU_ERR_BAD_ARG, "pu_maxpool2_bwd_code: code is NULL");
  return maxpool2_bwd_impl(nullptr, code, chan_scale, dy, acc, dx, B, H, W, C, flags, stream);
}

static int maxpool2_bwd_impl(const float* x, const unsigned char* code, const float* chan_scale, const float* dy, const float* acc, float* dx,
                             int B, int H, int W, int C, int flags, void* stream) {
  const int mask_in = (flags & PU_FLAG_MASK_IN) ? 1 : 0;
  PU_REQUIRE((x || code) && dy && dx && B > 0 && H >= 2 && W >= 2 && C > 0, PU_ERR_BAD_ARG, "pu_maxpool2_bwd: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  if ((H & 1) || (W & 1)) {  // floor mode leaves the last row/column unpooled: their gradient is zero (or just acc)
    const size_t bytes = sizeof(float) * (size_t)B * H * W * C;
    cudaError_t e = acc != nullptr ? cudaMemcpyAsync(dx, acc, bytes, cudaMemcpyDeviceToDevice, st) : cudaMemsetAsync(dx, 0, bytes, st);
    if (e != cudaSuccess) {
      pu::set_error("pu_maxpool2_bwd memset: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
  }
  const long long nwin = (long long)B * (H / 2) * (W / 2);
  if (C % 4 == 0 && (x == nullptr || pu::aligned16(x)) && (reinterpret_cast<uintptr_t>(code) & 3u) == 0 && pu::aligned16(dy) && pu::aligned16(dx) &&
      (acc == nullptr || pu::aligned16(acc)))
    pu::maxpool2_bwd_kernel<4><<<pu::grid_for(nwin * (C / 4)), 256, 0, st>>>(x, code, chan_scale, dy, acc, dx, B, H, W, C, mask_in);
  else
    pu::maxpool2_bwd_kernel<1><<<pu::grid_for(nwin * C), 256, 0, st>>>(x, code, chan_scale, dy, acc, dx, B, H, W, C, mask_in);
  return pu::post_launch("pu_maxpool2_bwd");
}

int pu_bilinear2x_fwd(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  PU_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0, PU_ERR_BAD_ARG, "pu_bilinear2x_fwd: bad argument");
  pu::bilinear2x_fwd_kernel<<<pu::grid_for((long long)B * 4 * H * W * C), 256, 0, pu::as_stream(stream)>>>(x, y, B, H, W, C);
  return pu::post_launch("pu_bilinear2x_fwd");
}

int pu_bilinear2x_bwd(const float* dy, float* dx, int B, int H, int W, int C, void* stream) {
  PU_REQUIRE(dy && dx && B > 0 && H > 0 && W > 0 && C > 0, PU_ERR_BAD_ARG, "pu_bilinear2x_bwd: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * H * W * C, st);
  if (e != cudaSuccess) {
    pu::set_error("pu_bilinear2x_bwd memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  pu::bilinear2x_bwd_kernel<<<pu::grid_for((long long)B * 4 * H * W * C), 256, 0, st>>>(dy, dx, B, H, W, C);
  return pu::post_launch("pu_bilinear2x_bwd");
}

}  // extern "C"
