// pointwise.cu — layout changes, 1x1 convolutions (with analytic CoordConv channels), channel scale,
// crop+concat, elementwise add.  All HBM-bound: coalesced, 128-bit vectorised where alignment allows.
#include "pu_common.cuh"

namespace pu {

// ---- NCHW <-> NHWC (32x32 smem transpose of the [C, HW] matrix per image) ----------------------
__global__ void transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
  // per image: x is [rows, cols], y is [cols, rows]; blockIdx.z = image
  __shared__ float tile[32][33];
  const size_t img = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = x[img + (size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) y[img + (size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

// ---- AddCoords channel values (coord_conv_script.py:69-96): k=0 -> xx (varies along width),
// k=1 -> yy (varies along height), k=2 -> rr = sqrt((xx-0.5)^2 + (yy-0.5)^2)
__device__ __forceinline__ float coord_val(int k, int i, int j, int H, int W) {
  const float xx = (W > 1) ? 2.f * (float)j / (float)(W - 1) - 1.f : -1.f;
  const float yy = (H > 1) ? 2.f * (float)i / (float)(H - 1) - 1.f : -1.f;
  if (k == 0) return xx;
  if (k == 1) return yy;
  return sqrtf((xx - 0.5f) * (xx - 0.5f) + (yy - 0.5f) * (yy - 0.5f));
}

// y[p][co] = act( sum_ci x[p][ci] w[co][ci] + sum_k coord_k(p) w[co][Cin+k] + bias[co] )
// thread = (pixel, block of 8 co); weights of the co block staged in smem as ws[ci][8]
__global__ void conv1x1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                   float* __restrict__ y, long long npix, int H, int W, int Cin, int Cout, int coords, int flags) {
  extern __shared__ float ws[];  // [(Cin+coords)][8]
  const int K = Cin + coords;
  const int co0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < K * 8; i += blockDim.x) {
    const int j = i & 7, k = i >> 3;
    ws[i] = (co0 + j < Cout) ? w[(size_t)(co0 + j) * K + k] : 0.f;
  }
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (bias != nullptr && co0 + j < Cout) ? bias[co0 + j] : 0.f;
  const float* xp = x + p * Cin;
  if (Cin % 4 == 0) {
    for (int c = 0; c < Cin; c += 4) {
      const float4 v = ldg4(xp + c);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(vv[u], ws[(c + u) * 8 + j], acc[j]);
    }
  } else {
    for (int c = 0; c < Cin; ++c) {
      const float v = __ldg(xp + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, ws[c * 8 + j], acc[j]);
    }
  }
  if (coords > 0) {
    const int hw = (int)(p % ((long long)H * W));
    const int i = hw / W, jx = hw - i * W;
    for (int k = 0; k < coords; ++k) {
      const float v = coord_val(k, i, jx, H, W);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, ws[(Cin + k) * 8 + j], acc[j]);
    }
  }
  float* yp = y + p * Cout + co0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (co0 + j < Cout) {
      float v = (flags & PU_FLAG_RELU) ? fmaxf(acc[j], 0.f) : acc[j];
      if (flags & PU_FLAG_ROUND_TF32) v = round_tf32(v);
      yp[j] = v;
    }
  }
}

// The output conv of the U-Net (unet_p.py:173: 8 -> 1 channels): a thread owns FOUR consecutive pixels — eight 128-bit loads of
// 128 contiguous bytes, one 128-bit store; same summation order as the generic kernel (bias, then channels ascending).
__global__ void __launch_bounds__(256) conv1x1_c8to1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ y, long long nquad,
                                                                int flags) {
  float wc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) wc[c] = __ldg(w + c);
  const float b0 = bias != nullptr ? __ldg(bias) : 0.f;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += (long long)gridDim.x * blockDim.x) {
    const float* xp = x + q * 32;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ldg4(xp + 4 * i);  // L1-allocating: two loads share every 32-byte sector
    float o[4];
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const float4 lo = v[2 * px], hi = v[2 * px + 1];
      float a = b0;
      a = fmaf(lo.x, wc[0], a); a = fmaf(lo.y, wc[1], a); a = fmaf(lo.z, wc[2], a); a = fmaf(lo.w, wc[3], a);
      a = fmaf(hi.x, wc[4], a); a = fmaf(hi.y, wc[5], a); a = fmaf(hi.z, wc[6], a); a = fmaf(hi.w, wc[7], a);
      if (flags & PU_FLAG_RELU) a = fmaxf(a, 0.f);
      if (flags & PU_FLAG_ROUND_TF32) a = round_tf32(a);
      o[px] = a;
    }
    *reinterpret_cast<float4*>(y + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// dx[p][ci] = sum_co g[p][co] w[co][ci]     (thread = pixel; Cout small)
__global__ void conv1x1_dx_kernel(const float* __restrict__ g, const float* __restrict__ w, float* __restrict__ dx,
                                  long long npix, int Cin, int Cout, int K) {
  extern __shared__ float ws[];  // [Cout][Cin]
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) ws[i] = w[(size_t)(i / Cin) * K + (i % Cin)];
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const float* gp = g + p * Cout;
  float* dp = dx + p * Cin;
  for (int c = 0; c < Cin; ++c) {
    float s = 0.f;
    for (int co = 0; co < Cout; ++co) s = fmaf(__ldg(gp + co), ws[co * Cin + c], s);
    dp[c] = s;
  }
}

// dw[co][k] = sum_p g[p][co] * in_k(p)  with in_k = x channels, then coords, then the constant 1 (bias).
// thread = (pair (co,k), pixel sub-stream); pairs <= 256.  Result accumulated with atomics into
// dwb = [Cout*(K+1)] (pre-zeroed), laid out [co][K+1].
__global__ void conv1x1_dw_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ dwb,
                                  long long npix, int H, int W, int Cin, int Cout, int coords, int P, int nsub) {
  __shared__ float red[256];
  const int K1 = Cin + coords + 1;
  const int pair = threadIdx.x % P;  // valid if pair < Cout*K1
  const int sub = threadIdx.x / P;
  const int npairs = Cout * K1;
  const int co = pair / K1, k = pair - co * K1;
  float acc = 0.f;
  if (pair < npairs && sub < nsub) {
    const long long stride = (long long)gridDim.x * nsub;
    for (long long p = (long long)blockIdx.x * nsub + sub; p < npix; p += stride) {
      const float gv = __ldg(g + p * Cout + co);
      float iv;
      if (k < Cin) {
        iv = __ldg(x + p * Cin + k);
      } else if (k < Cin + coords) {
        const int hw = (int)(p % ((long long)H * W));
        const int i = hw / W;
        iv = coord_val(k - Cin, i, hw - i * W, H, W);
      } else {
        iv = 1.f;
      }
      acc = fmaf(gv, iv, acc);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (sub == 0 && pair < npairs) {
    float s = 0.f;
    for (int u = 0; u < nsub; ++u) s += red[u * P + pair];
    atomicAdd(dwb + pair, s);
  }
}

// outc backward (Cout == 1, no coords), one pass: dx[p][k] = g[p]*w[k]; dw[k] = sum_p g[p]*x[p][k]; db = sum_p g[p].
// thread = pixel (grid-stride), 128-bit loads/stores, per-thread accumulators reduced by warp shuffles + atomics
// into dwb = [CIN + 1] (pre-zeroed).
template <int CIN>
__global__ void __launch_bounds__(256) conv1x1_bwd_c1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ g, float* __restrict__ dx,
                                                             float* __restrict__ dwb, long long npix, int mask_in) {
  __shared__ float red[8][CIN + 1];
  float wv[CIN];
#pragma unroll
  for (int k = 0; k < CIN; ++k) wv[k] = __ldg(w + k);
  float acc[CIN + 1];
#pragma unroll
  for (int k = 0; k <= CIN; ++k) acc[k] = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const float gv = __ldg(g + p);
    acc[CIN] += gv;
#pragma unroll
    for (int k = 0; k < CIN; k += 4) {
      const float4 xv = ldg4(x + p * CIN + k);
      acc[k] = fmaf(gv, xv.x, acc[k]);
      acc[k + 1] = fmaf(gv, xv.y, acc[k + 1]);
      acc[k + 2] = fmaf(gv, xv.z, acc[k + 2]);
      acc[k + 3] = fmaf(gv, xv.w, acc[k + 3]);
      if (dx != nullptr) {
        float4 o = make_float4(gv * wv[k], gv * wv[k + 1], gv * wv[k + 2], gv * wv[k + 3]);
        if (mask_in) {  // ReLU mask of the layer that produced x
          o.x = xv.x > 0.f ? o.x : 0.f; o.y = xv.y > 0.f ? o.y : 0.f; o.z = xv.z > 0.f ? o.z : 0.f; o.w = xv.w > 0.f ? o.w : 0.f;
        }
        *reinterpret_cast<float4*>(dx + p * CIN + k) = o;
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= CIN; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[wid][k] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x <= CIN) {
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) sum += red[u][threadIdx.x];
    atomicAdd(dwb + threadIdx.x, sum);
  }
}

// scatter the [co][K+1] accumulator into dw [Cout][K] and db [Cout]
__global__ void conv1x1_dw_finish_kernel(const float* __restrict__ dwb, float* __restrict__ dw, float* __restrict__ db, int Cout, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * (K + 1)) return;
  const int co = i / (K + 1), k = i - co * (K + 1);
  if (k < K) dw[co * K + k] = dwb[i];
  else if (db != nullptr) db[co] = dwb[i];
}

// ---- channel scale / concat -------------------------------------------------------------------
__global__ void chan_scale_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y,
                                  int B, long long hw, int C) {
  const long long n = (long long)B * hw * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int b = (int)(i / (hw * C));
    y[i] = x[i] * __ldg(s + (size_t)b * C + c);
  }
}

__global__ void concat_scale_kernel(View s0, View s1, const float* __restrict__ scale, float* __restrict__ y, int B, int H, int W,
                                    int flags) {
  const int C = s0.C + s1.C;
  const long long n = (long long)B * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int xx = (int)(p % W);
    p /= W;
    const int yy = (int)(p % H);
    const int b = (int)(p / H);
    float v;
    if (c < s0.C) v = __ldg(s0.p + (((size_t)b * s0.Hs + yy + s0.oy) * s0.Ws + xx + s0.ox) * s0.C + c);
    else v = __ldg(s1.p + (((size_t)b * s1.Hs + yy + s1.oy) * s1.Ws + xx + s1.ox) * s1.C + (c - s0.C));
    if (scale != nullptr) v *= __ldg(scale + (size_t)b * C + c);
    if (flags & PU_FLAG_ROUND_TF32) v = round_tf32(v);
    y[i] = v;
  }
}

__global__ void concat_scale_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ scale, ViewW d0, ViewW d1,
                                        int B, int H, int W) {
  const int C = d0.C + d1.C;
  const long long n = (long long)B * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int xx = (int)(p % W);
    p /= W;
    const int yy = (int)(p % H);
    const int b = (int)(p / H);
    float v = dy[i];
    if (scale != nullptr) v *= __ldg(scale + (size_t)b * C + c);
    if (c < d0.C) {
      if (d0.p != nullptr) d0.p[(((size_t)b * d0.Hs + yy + d0.oy) * d0.Ws + xx + d0.ox) * d0.C + c] = v;
    } else {
      if (d1.p != nullptr) d1.p[(((size_t)b * d1.Hs + yy + d1.oy) * d1.Ws + xx + d1.ox) * d1.C + (c - d0.C)] = v;
    }
  }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n) {
  const long long n4 = n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 u = reinterpret_cast<const float4*>(a)[i];
    const float4 v = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(o)[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) o[i] = a[i] + b[i];
}

static inline int ew_grid(long long n, int per_thread = 4) {
  long long g = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  if (g < 1) g = 1;
  if (g > 16LL * kNumSMs) g = 16LL * kNumSMs;
  return (int)g;
}

}  // namespace pu

extern "C" {

int pu_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  PU_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, PU_ERR_BAD_ARG, "pu_nchw_to_nhwc: bad argument");
  const int rows = C, cols = H * W;
  dim3 grid(pu::cdiv(cols, 32), pu::cdiv(rows, 32), B), block(32, 8);
  PU_REQUIRE(B <= 65535 && grid.y <= 65535, PU_ERR_UNSUPPORTED, "pu_nchw_to_nhwc: grid too large");
  pu::transpose_kernel<<<grid, block, 0, pu::as_stream(stream)>>>(x, y, rows, cols);
  return pu::post_launch("pu_nchw_to_nhwc");
}

int pu_nhwc_to_nchw(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  PU_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, PU_ERR_BAD_ARG, "pu_nhwc_to_nchw: bad argument");
  const int rows = H * W, cols = C;
  dim3 grid(pu::cdiv(cols, 32), pu::cdiv(rows, 32), B), block(32, 8);
  PU_REQUIRE(B <= 65535 && grid.y <= 65535, PU_ERR_UNSUPPORTED, "pu_nhwc_to_nchw: grid too large");
  pu::transpose_kernel<<<grid, block, 0, pu::as_stream(stream)>>>(x, y, rows, cols);
  return pu::post_launch("pu_nhwc_to_nchw");
}

int pu_conv1x1_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin, int Cout,
                   int coords, int flags, void* stream) {
  PU_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, PU_ERR_BAD_ARG, "pu_conv1x1_fwd: bad argument");
  PU_REQUIRE(coords == 0 || coords == 2 || coords == 3, PU_ERR_BAD_ARG, "pu_conv1x1_fwd: coords must be 0, 2 or 3");
  const size_t smem = (size_t)(Cin + coords) * 8 * sizeof(float);
  PU_REQUIRE(smem <= 48 * 1024, PU_ERR_UNSUPPORTED, "pu_conv1x1_fwd: Cin=%d too large", Cin);
  PU_REQUIRE(Cin % 4 != 0 || pu::aligned16(x), PU_ERR_BAD_ARG, "pu_conv1x1_fwd: x not 16-byte aligned");
  const long long npix = (long long)B * H * W;
  if (Cout == 1 && Cin == 8 && coords == 0 && npix % 4 == 0 && pu::aligned16(x) && pu::aligned16(y)) {
    const long long nquad = npix / 4;
    long long blocks = (nquad + 255) / 256;
    if (blocks > 16 * pu::kNumSMs) blocks = 16 * pu::kNumSMs;
    pu::conv1x1_c8to1_fwd_kernel<<<(unsigned)blocks, 256, 0, pu::as_stream(stream)>>>(x, w, bias, y, nquad, flags);
    return pu::post_launch("pu_conv1x1_fwd c8to1");
  }
  dim3 grid((unsigned)((npix + 255) / 256), pu::cdiv(Cout, 8));
  pu::conv1x1_fwd_kernel<<<grid, 256, smem, pu::as_stream(stream)>>>(x, w, bias, y, npix, H, W, Cin, Cout, coords, flags);
  return pu::post_launch("pu_conv1x1_fwd");
}

int pu_conv1x1_bwd(const float* x, const float* w, const float* g, float* dx, float* dw, float* db, float* ws, int B, int H, int W,
                   int Cin, int Cout, int coords, int flags, void* stream) {
  PU_REQUIRE(x && w && g && dw && ws && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, PU_ERR_BAD_ARG, "pu_conv1x1_bwd: bad argument");
  PU_REQUIRE(coords == 0 || coords == 2 || coords == 3, PU_ERR_BAD_ARG, "pu_conv1x1_bwd: coords must be 0, 2 or 3");
  cudaStream_t st = pu::as_stream(stream);
  const long long npix = (long long)B * H * W;
  const int K = Cin + coords;
  const int npairs = Cout * (K + 1);
  PU_REQUIRE(npairs <= 256, PU_ERR_UNSUPPORTED, "pu_conv1x1_bwd: Cout*(Cin+coords+1)=%d > 256", npairs);
  float* scratch = ws;  // caller-provided [Cout*(Cin+coords+1)] accumulator
  cudaError_t e = cudaMemsetAsync(scratch, 0, npairs * sizeof(float), st);
  if (e != cudaSuccess) {
    pu::set_error("pu_conv1x1_bwd memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  const int mask_in = (flags & PU_FLAG_MASK_IN) ? 1 : 0;
  const bool fused_ok = Cout == 1 && coords == 0 && (Cin == 8 || Cin == 16 || Cin == 32 || Cin == 64) && pu::aligned16(x) && (dx == nullptr || pu::aligned16(dx));
  PU_REQUIRE(!mask_in || fused_ok, PU_ERR_UNSUPPORTED, "pu_conv1x1_bwd: PU_FLAG_MASK_IN needs the fused output-conv kernel (Cout=1, Cin in 8/16/32/64)");
  if (fused_ok) {
    // fused single pass for the output conv: dx, dw and db together
    long long blocks = (npix + 256 * 4 - 1) / (256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > 8 * pu::kNumSMs) blocks = 8 * pu::kNumSMs;
    switch (Cin) {
      case 8: pu::conv1x1_bwd_c1_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(x, w, g, dx, scratch, npix, mask_in); break;
      case 16: pu::conv1x1_bwd_c1_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(x, w, g, dx, scratch, npix, mask_in); break;
      case 32: pu::conv1x1_bwd_c1_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(x, w, g, dx, scratch, npix, mask_in); break;
      default: pu::conv1x1_bwd_c1_kernel<64><<<(unsigned)blocks, 256, 0, st>>>(x, w, g, dx, scratch, npix, mask_in); break;
    }
    int rc1 = pu::post_launch("pu_conv1x1_bwd fused");
    if (rc1) return rc1;
    if (flags & PU_FLAG_DEFER_FINISH) return PU_OK;
    pu::conv1x1_dw_finish_kernel<<<pu::cdiv(npairs, 256), 256, 0, st>>>(scratch, dw, db, Cout, K);
    return pu::post_launch("pu_conv1x1_bwd finish");
  }
  if (dx != nullptr) {
    const size_t smem = (size_t)Cout * Cin * sizeof(float);
    pu::conv1x1_dx_kernel<<<(unsigned)((npix + 255) / 256), 256, smem, st>>>(g, w, dx, npix, Cin, Cout, K);
    int rc = pu::post_launch("pu_conv1x1_bwd dx");
    if (rc) return rc;
  }
  int P = 1;
  while (P < npairs) P <<= 1;  // pad pairs to a power of two <= 256
  const int nsub = 256 / P;
  long long blocks = (npix + (long long)nsub * 64 - 1) / ((long long)nsub * 64);
  if (blocks < 1) blocks = 1;
  if (blocks > 4 * pu::kNumSMs) blocks = 4 * pu::kNumSMs;
  pu::conv1x1_dw_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, g, scratch, npix, H, W, Cin, Cout, coords, P, nsub);
  int rc = pu::post_launch("pu_conv1x1_bwd dw");
  if (rc) return rc;
  if (flags & PU_FLAG_DEFER_FINISH) return PU_OK;
  pu::conv1x1_dw_finish_kernel<<<pu::cdiv(npairs, 256), 256, 0, st>>>(scratch, dw, db, Cout, K);
  return pu::post_launch("pu_conv1x1_bwd finish");
}

int pu_conv1x1_dw_finish(const float* ws, float* dw, float* db, int Cin, int Cout, int coords, void* stream) {
  PU_REQUIRE(ws && dw && Cin > 0 && Cout > 0 && (coords == 0 || coords == 2 || coords == 3), PU_ERR_BAD_ARG, "pu_conv1x1_dw_finish: bad argument");
  const int K = Cin + coords, npairs = Cout * (K + 1);
  pu::conv1x1_dw_finish_kernel<<<pu::cdiv(npairs, 256), 256, 0, pu::as_stream(stream)>>>(ws, dw, db, Cout, K);
  return pu::post_launch("pu_conv1x1_dw_finish");
}

int pu_chan_scale(const float* x, const float* s, float* y, int B, long long hw, int C, void* stream) {
  PU_REQUIRE(x && s && y && B > 0 && hw > 0 && C > 0, PU_ERR_BAD_ARG, "pu_chan_scale: bad argument");
  const long long n = (long long)B * hw * C;
  pu::chan_scale_kernel<<<pu::ew_grid(n), 256, 0, pu::as_stream(stream)>>>(x, s, y, B, hw, C);
  return pu::post_launch("pu_chan_scale");
}

int pu_concat_scale_fwd(const float* src0, int H0, int W0, int C0, int oy0, int ox0, const float* src1, int H1, int W1, int C1,
                        int oy1, int ox1, const float* chan_scale, float* y, int B, int H, int W, int flags, void* stream) {
  PU_REQUIRE(src0 && src1 && y && B > 0 && H > 0 && W > 0 && C0 > 0 && C1 > 0, PU_ERR_BAD_ARG, "pu_concat_scale_fwd: bad argument");
  PU_REQUIRE(oy0 >= 0 && ox0 >= 0 && oy0 + H <= H0 && ox0 + W <= W0 && oy1 >= 0 && ox1 >= 0 && oy1 + H <= H1 && ox1 + W <= W1,
             PU_ERR_BAD_ARG, "pu_concat_scale_fwd: window exceeds source");
  pu::View s0{src0, H0, W0, C0, oy0, ox0}, s1{src1, H1, W1, C1, oy1, ox1};
  const long long n = (long long)B * H * W * (C0 + C1);
  pu::concat_scale_kernel<<<pu::ew_grid(n), 256, 0, pu::as_stream(stream)>>>(s0, s1, chan_scale, y, B, H, W, flags);
  return pu::post_launch("pu_concat_scale_fwd");
}

int pu_concat_scale_bwd(const float* dy, const float* chan_scale, float* dx0, int H0, int W0, int C0, int oy0, int ox0,
                        float* dx1, int H1, int W1, int C1, int oy1, int ox1, int B, int H, int W, void* stream) {
  PU_REQUIRE(dy && (dx0 || dx1) && B > 0 && H > 0 && W > 0 && C0 > 0 && C1 > 0, PU_ERR_BAD_ARG, "pu_concat_scale_bwd: bad argument");
  PU_REQUIRE(oy0 >= 0 && ox0 >= 0 && oy0 + H <= H0 && ox0 + W <= W0 && oy1 >= 0 && ox1 >= 0 && oy1 + H <= H1 && ox1 + W <= W1,
             PU_ERR_BAD_ARG, "pu_concat_scale_bwd: window exceeds destination");
  pu::ViewW d0{dx0, H0, W0, C0, oy0, ox0}, d1{dx1, H1, W1, C1, oy1, ox1};
  const long long n = (long long)B * H * W * (C0 + C1);
  pu::concat_scale_bwd_kernel<<<pu::ew_grid(n), 256, 0, pu::as_stream(stream)>>>(dy, chan_scale, d0, d1, B, H, W);
  return pu::post_launch("pu_concat_scale_bwd");
}

int pu_add(const float* a, const float* b, float* out, long long n, void* stream) {
  PU_REQUIRE(a && b && out && n > 0, PU_ERR_BAD_ARG, "pu_add: bad argument");
  PU_REQUIRE(pu::aligned16(a) && pu::aligned16(b) && pu::aligned16(out), PU_ERR_BAD_ARG, "pu_add: pointers not 16-byte aligned");
  pu::add_kernel<<<pu::ew_grid(n, 8), 256, 0, pu::as_stream(stream)>>>(a, b, out, n);
  return pu::post_launch("pu_add");
}

}  // extern "C"
