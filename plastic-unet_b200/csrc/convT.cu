// convT.cu — transposed convolutions of the decoder (NHWC fp32), sm_100a CUDA-core kernels.
//   2x2 stride 2 (reference unet_p.py:155):    a per-pixel [Cin]->[4*Cout] product + pixel-shuffle store.
//   3x3 stride 2 pad 0 (unet_p_res.py:207):    gather form over the (2H+1)x(2W+1) output, with the
//                                              F.pad crop (unet_p_res.py:215-217) fused as a window.
// These are ~4 % of the model FLOPs; they are written for coalesced NHWC traffic,
// weights staged in shared memory per 8-channel output block.
#include <stdlib.h>
#include "pu_common.cuh"

namespace pu {

// ---------------- 2x2 stride 2 -----------------------------------------------------------------
// thread = (output pixel, block of 8 co).  w is [Cin][Cout][2][2].
__global__ void convT2x2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                    float* __restrict__ y, int B, int H, int W, int Cin, int Cout, int flags) {
  extern __shared__ __align__(16) float ws[];  // [4][Cin][8]
  const int co0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < 4 * Cin * 8; i += blockDim.x) {
    const int j = i & 7, ci = (i >> 3) % Cin, ac = i / (8 * Cin);
    ws[i] = (co0 + j < Cout) ? w[((size_t)ci * Cout + co0 + j) * 4 + ac] : 0.f;
  }
  __syncthreads();
  const int Ho = 2 * H, Wo = 2 * W;
  const long long npix = (long long)B * Ho * Wo;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int ox = (int)(p % Wo);
  const int oy = (int)((p / Wo) % Ho);
  const int b = (int)(p / ((long long)Wo * Ho));
  const int ac = (oy & 1) * 2 + (ox & 1);
  const float* xp = x + (((size_t)b * H + (oy >> 1)) * W + (ox >> 1)) * Cin;
  const float* wp = ws + ac * Cin * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (bias != nullptr && co0 + j < Cout) ? bias[co0 + j] : 0.f;
  if (Cin % 4 == 0) {
    float2 a2[4] = {make_float2(acc[0], acc[1]), make_float2(acc[2], acc[3]), make_float2(acc[4], acc[5]), make_float2(acc[6], acc[7])};
    for (int c = 0; c < Cin; c += 4) {
      const float4 v = ldg4(xp + c);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 wa = *reinterpret_cast<const float4*>(wp + (c + u) * 8);
        const float4 wb = *reinterpret_cast<const float4*>(wp + (c + u) * 8 + 4);
        const float2 xv = make_float2(vv[u], vv[u]);
        a2[0] = __ffma2_rn(xv, make_float2(wa.x, wa.y), a2[0]);
        a2[1] = __ffma2_rn(xv, make_float2(wa.z, wa.w), a2[1]);
        a2[2] = __ffma2_rn(xv, make_float2(wb.x, wb.y), a2[2]);
        a2[3] = __ffma2_rn(xv, make_float2(wb.z, wb.w), a2[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[2 * j] = a2[j].x; acc[2 * j + 1] = a2[j].y; }
  } else {
    for (int c = 0; c < Cin; ++c) {
      const float v = __ldg(xp + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[c * 8 + j], acc[j]);
    }
  }
  if (flags & PU_FLAG_ROUND_TF32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = round_tf32(acc[j]);
  }
  float* yp = y + p * Cout + co0;
  if (Cout % 4 == 0 && co0 + 8 <= Cout) {
    *reinterpret_cast<float4*>(yp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(yp + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (co0 + j < Cout) yp[j] = acc[j];
  }
}

// dx[b,i,j,ci] = sum_{a,c,co} dy[b,2i+a,2j+c,co] w[ci][co][a][c];  thread = (input pixel, block of 8 ci)
__global__ void convT2x2_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                   int B, int H, int W, int Cin, int Cout, const float* __restrict__ xmask) {
  extern __shared__ __align__(16) float ws[];  // [4][Cout][8 ci]
  const int ci0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < 4 * Cout * 8; i += blockDim.x) {
    const int j = i & 7, co = (i >> 3) % Cout, ac = i / (8 * Cout);
    ws[i] = (ci0 + j < Cin) ? w[((size_t)(ci0 + j) * Cout + co) * 4 + ac] : 0.f;
  }
  __syncthreads();
  const long long npix = (long long)B * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int jx = (int)(p % W);
  const int iy = (int)((p / W) % H);
  const int b = (int)(p / ((long long)W * H));
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (Cout % 4 == 0) {
    float2 a2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    for (int ac = 0; ac < 4; ++ac) {
      const float* gp = dy + (((size_t)b * 2 * H + 2 * iy + (ac >> 1)) * 2 * W + 2 * jx + (ac & 1)) * Cout;
      const float* wp = ws + ac * Cout * 8;
      for (int co = 0; co < Cout; co += 4) {
        const float4 gv = ldg4(gp + co);
        const float vv[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 wa = *reinterpret_cast<const float4*>(wp + (co + u) * 8);
          const float4 wb = *reinterpret_cast<const float4*>(wp + (co + u) * 8 + 4);
          const float2 xv = make_float2(vv[u], vv[u]);
          a2[0] = __ffma2_rn(xv, make_float2(wa.x, wa.y), a2[0]);
          a2[1] = __ffma2_rn(xv, make_float2(wa.z, wa.w), a2[1]);
          a2[2] = __ffma2_rn(xv, make_float2(wb.x, wb.y), a2[2]);
          a2[3] = __ffma2_rn(xv, make_float2(wb.z, wb.w), a2[3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[2 * j] = a2[j].x; acc[2 * j + 1] = a2[j].y; }
  } else {
    for (int ac = 0; ac < 4; ++ac) {
      const float* gp = dy + (((size_t)b * 2 * H + 2 * iy + (ac >> 1)) * 2 * W + 2 * jx + (ac & 1)) * Cout;
      const float* wp = ws + ac * Cout * 8;
      for (int co = 0; co < Cout; ++co) {
        const float v = __ldg(gp + co);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[co * 8 + j], acc[j]);
      }
    }
  }
  float* dp = dx + p * Cin + ci0;
  if (xmask != nullptr) {  // ReLU mask of the layer that produced x
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (ci0 + j < Cin && !(__ldg(xmask + p * Cin + ci0 + j) > 0.f)) acc[j] = 0.f;
  }
  if (Cin % 4 == 0 && ci0 + 8 <= Cin) {
    *reinterpret_cast<float4*>(dp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(dp + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (ci0 + j < Cin) dp[j] = acc[j];
  }
}

// dw[ci][co][a][c] = sum_{b,i,j} x[b,i,j,ci] dy[b,2i+a,2j+c,co].
// A [Cin x P] . [P x 4*Cout] contraction over the P input pixels.  Block = 256 threads, each owning one
// (ci, ac, 4 consecutive co) slice = 1024 outputs per block; the block walks its pixel range in chunks
// staged through shared memory (x rows and the four dy rows of every pixel, coalesced 128-bit loads), so each
// thread does 1 LDS + 1 LDS.128 per 4 FMA.  Pixel range split over blockIdx.y, fp32 atomics at the end.
__global__ void __launch_bounds__(256) convT2x2_dw_tiled_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                float* __restrict__ dw, int B, int H, int W, int Cin, int Cout,
                                                                int pch, int nrep) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                 // [pch][Cin]
  float* ds = sm + pch * Cin;     // [pch][4][Cout]
  const int co4n = Cout >> 2;
  // a block owns `slices` = min(256, Cin*Cout) four-output slices; when there are fewer than 256 of them the block's
  // threads are replicated nrep times over the pixels of a chunk (rep = pixel residue class)
  const int slices = 256 / nrep;
  const int sl = threadIdx.x % slices, rep = threadIdx.x / slices;
  const long long e = (long long)blockIdx.x * slices + sl;  // ((ci*4 + ac)*co4n + co4)
  const int co4 = (int)(e % co4n);
  const int ac = (int)((e / co4n) & 3);
  const int ci = (int)(e / (4 * co4n));
  const bool valid = ci < Cin;
  const long long npix = (long long)B * H * W;
  const long long per = (npix + gridDim.y - 1) / gridDim.y;
  const long long p0 = (long long)blockIdx.y * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int cin4 = Cin >> 2;
  for (long long pc = p0; pc < p1; pc += pch) {
    const int np = (int)((p1 - pc) < pch ? (p1 - pc) : pch);
    __syncthreads();
    // stage x rows: np * Cin floats, contiguous in global memory
    for (int i = threadIdx.x; i < np * cin4; i += 256)
      reinterpret_cast<float4*>(xs)[i] = ldg4(x + pc * Cin + 4 * (size_t)i);
    // stage the 4 dy rows of every pixel
    for (int i = threadIdx.x; i < np * 4 * co4n; i += 256) {
      const int c4 = i % co4n;
      const int a = (i / co4n) & 3;
      const int pp = i / (4 * co4n);
      const long long p = pc + pp;
      const int jx = (int)(p % W);
      const int iy = (int)((p / W) % H);
      const int b = (int)(p / ((long long)W * H));
      reinterpret_cast<float4*>(ds)[i] =
          ldg4(dy + (((size_t)b * 2 * H + 2 * iy + (a >> 1)) * 2 * W + 2 * jx + (a & 1)) * Cout + 4 * c4);
    }
    __syncthreads();
    if (valid) {
#pragma unroll 4
      for (int pp = rep; pp < np; pp += nrep) {
        const float xv = xs[pp * Cin + ci];
        const float4 g = *reinterpret_cast<const float4*>(ds + ((size_t)pp * 4 + ac) * Cout + 4 * co4);
        acc.x = fmaf(xv, g.x, acc.x);
        acc.y = fmaf(xv, g.y, acc.y);
        acc.z = fmaf(xv, g.z, acc.z);
        acc.w = fmaf(xv, g.w, acc.w);
      }
    }
  }
  if (valid) {
    float* o = dw + ((size_t)ci * Cout + 4 * co4) * 4 + ac;
    atomicAdd(o + 0, acc.x);
    atomicAdd(o + 4, acc.y);
    atomicAdd(o + 8, acc.z);
    atomicAdd(o + 12, acc.w);
  }
}

// generic fallback (ragged channel counts): one thread per output element, pixel range split over blockIdx.y
__global__ void convT2x2_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                   int B, int H, int W, int Cin, int Cout) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int nout = Cin * Cout * 4;
  if (e >= nout) return;
  const int co = e % Cout;
  const int ac = (e / Cout) & 3;
  const int ci = e / (4 * Cout);
  const long long npix = (long long)B * H * W;
  const long long per = (npix + gridDim.y - 1) / gridDim.y;
  const long long p0 = (long long)blockIdx.y * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  float acc = 0.f;
  for (long long p = p0; p < p1; ++p) {
    const int jx = (int)(p % W);
    const int iy = (int)((p / W) % H);
    const int b = (int)(p / ((long long)W * H));
    const float xv = __ldg(x + p * Cin + ci);
    const float gv = __ldg(dy + (((size_t)b * 2 * H + 2 * iy + (ac >> 1)) * 2 * W + 2 * jx + (ac & 1)) * Cout + co);
    acc = fmaf(xv, gv, acc);
  }
  atomicAdd(dw + ((size_t)ci * Cout + co) * 4 + ac, acc);
}

// ---------------- 3x3 stride 2 pad 0, cropped window ----------------------------------------------
// thread = (window pixel, block of 8 co).  w is [Cin][Cout][3][3]; full output (fy,fx) = (oy+wy, ox+wx);
// contributing taps have (fy-ky) even and 0 <= (fy-ky)/2 < H.
__global__ void convT3x3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                    const float* __restrict__ scale, float* __restrict__ y, int B, int H, int W, int Cin, int Cout,
                                    int Ho, int Wo, int oy, int ox, int flags) {
  extern __shared__ float ws[];  // [9][CK][8], CK = channel chunk
  constexpr int CK = 32;
  const int co0 = blockIdx.y * 8;
  const long long npix = (long long)B * Ho * Wo;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < npix;
  int wx = 0, wy = 0, b = 0;
  if (active) {
    wx = (int)(p % Wo);
    wy = (int)((p / Wo) % Ho);
    b = (int)(p / ((long long)Wo * Ho));
  }
  const int fy = wy + oy, fx = wx + ox;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += CK) {
    const int cc = min(CK, Cin - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * CK * 8; i += blockDim.x) {
      const int j = i & 7, ci = (i >> 3) % CK, tap = i / (8 * CK);
      ws[i] = (ci < cc && co0 + j < Cout) ? w[((size_t)(c0 + ci) * Cout + co0 + j) * 9 + tap] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = fy - ky;
      if (ty < 0 || (ty & 1) || (ty >> 1) >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = fx - kx;
        if (tx < 0 || (tx & 1) || (tx >> 1) >= W) continue;
        const float* xp = x + (((size_t)b * H + (ty >> 1)) * W + (tx >> 1)) * Cin + c0;
        const float* wp = ws + (ky * 3 + kx) * CK * 8;
        for (int c = 0; c < cc; ++c) {
          const float v = __ldg(xp + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[c * 8 + j], acc[j]);
        }
      }
    }
  }
  if (!active) return;
  float* yp = y + p * Cout + co0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (co0 + j < Cout) {
      float v = acc[j] + (bias != nullptr ? bias[co0 + j] : 0.f);
      if (scale != nullptr) v *= __ldg(scale + (size_t)b * Cout + co0 + j);
      if (flags & PU_FLAG_ROUND_TF32) v = round_tf32(v);
      yp[j] = v;
    }
  }
}

// dx[b,i,j,ci] = sum_{ky,kx,co} dywin[b,2i+ky-oy,2j+kx-ox,co] * s[b,co] * w[ci][co][ky][kx]
__global__ void convT3x3_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ scale,
                                   float* __restrict__ dx, int B, int H, int W, int Cin, int Cout, int Ho, int Wo, int oy, int ox) {
  extern __shared__ float ws[];  // [9][CK co][8 ci]
  constexpr int CK = 32;
  const int ci0 = blockIdx.y * 8;
  const long long npix = (long long)B * H * W;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < npix;
  int jx = 0, iy = 0, b = 0;
  if (active) {
    jx = (int)(p % W);
    iy = (int)((p / W) % H);
    b = (int)(p / ((long long)W * H));
  }
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int c0 = 0; c0 < Cout; c0 += CK) {
    const int cc = min(CK, Cout - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * CK * 8; i += blockDim.x) {
      const int j = i & 7, co = (i >> 3) % CK, tap = i / (8 * CK);
      ws[i] = (co < cc && ci0 + j < Cin) ? w[((size_t)(ci0 + j) * Cout + c0 + co) * 9 + tap] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    for (int ky = 0; ky < 3; ++ky) {
      const int wy = 2 * iy + ky - oy;
      if (wy < 0 || wy >= Ho) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int wx = 2 * jx + kx - ox;
        if (wx < 0 || wx >= Wo) continue;
        const float* gp = dy + (((size_t)b * Ho + wy) * Wo + wx) * Cout + c0;
        const float* sp = scale != nullptr ? scale + (size_t)b * Cout + c0 : nullptr;
        const float* wp = ws + (ky * 3 + kx) * CK * 8;
        for (int c = 0; c < cc; ++c) {
          float v = __ldg(gp + c);
          if (sp != nullptr) v *= __ldg(sp + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[c * 8 + j], acc[j]);
        }
      }
    }
  }
  if (!active) return;
  float* dp = dx + p * Cin + ci0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (ci0 + j < Cin) dp[j] = acc[j];
}

// dw[ci][co][ky][kx] = sum_{b,i,j} x[b,i,j,ci] * dywin[b,2i+ky-oy,2j+kx-ox,co]*s[b,co]
// thread = output element e = ((ci*9 + tap)*Cout + co); input-pixel range split over blockIdx.y.
__global__ void convT3x3_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ scale,
                                   float* __restrict__ dw, int B, int H, int W, int Cin, int Cout, int Ho, int Wo, int oy, int ox) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nout = (long long)Cin * Cout * 9;
  if (e >= nout) return;
  const int co = (int)(e % Cout);
  const int tap = (int)((e / Cout) % 9);
  const int ci = (int)(e / (9LL * Cout));
  const int ky = tap / 3, kx = tap - ky * 3;
  const long long npix = (long long)B * H * W;
  const long long per = (npix + gridDim.y - 1) / gridDim.y;
  const long long p0 = (long long)blockIdx.y * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  float acc = 0.f;
  for (long long p = p0; p < p1; ++p) {
    const int jx = (int)(p % W);
    const int iy = (int)((p / W) % H);
    const int b = (int)(p / ((long long)W * H));
    const int wy = 2 * iy + ky - oy, wx = 2 * jx + kx - ox;
    if (wy < 0 || wy >= Ho || wx < 0 || wx >= Wo) continue;
    float gv = __ldg(dy + (((size_t)b * Ho + wy) * Wo + wx) * Cout + co);
    if (scale != nullptr) gv *= __ldg(scale + (size_t)b * Cout + co);
    acc = fmaf(__ldg(x + p * Cin + ci), gv, acc);
  }
  atomicAdd(dw + ((size_t)ci * Cout + co) * 9 + tap, acc);
}

// Tiled version for channel counts that are multiples of 4: per tap a [Cin x P] . [P x Cout] contraction over the input
// pixels.  Block = 64 ci x 64 co outputs of ONE tap over a slice of the pixels; 256 threads x (4 ci x 4 co) register
// tiles; x rows and the tap's (gathered, scaled, zero-filled) dy rows are staged 16 pixels at a time.
__global__ void __launch_bounds__(256) convT3x3_dw_tiled_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                const float* __restrict__ scale, float* __restrict__ dw, int B, int H,
                                                                int W, int Cin, int Cout, int Ho, int Wo, int oy, int ox, int nsplit) {
  constexpr int PC = 16;
  __shared__ __align__(16) float xs[PC][64];
  __shared__ __align__(16) float ds[PC][64];
  const int tid = threadIdx.x;
  const int ci_blk = blockIdx.x * 64, co_blk = blockIdx.y * 64;
  const int tap = blockIdx.z % 9, split = blockIdx.z / 9;
  const int ky = tap / 3, kx = tap - 3 * ky;
  const int tci = (tid & 15) * 4, tco = (tid >> 4) * 4;  // this thread's 4x4 outputs inside the 64x64 tile
  const long long npix = (long long)B * H * W;
  const long long per = (npix + nsplit - 1) / nsplit;
  const long long p0 = (long long)split * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long pc = p0; pc < p1; pc += PC) {
    __syncthreads();
    {  // stage: 16 pixels x 16 float4 for x and for dy = 512 float4, two per thread
      const int pp = tid >> 4, c4 = (tid & 15) * 4;
      const long long p = pc + pp;
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < p1) {
        const int jx = (int)(p % W);
        const int iy = (int)((p / W) % H);
        const int b = (int)(p / ((long long)W * H));
        if (ci_blk + c4 < Cin) xv = ldg4(x + p * Cin + ci_blk + c4);
        const int wy = 2 * iy + ky - oy, wx = 2 * jx + kx - ox;
        if (wy >= 0 && wy < Ho && wx >= 0 && wx < Wo && co_blk + c4 < Cout) {
          dv = ldg4(dy + (((size_t)b * Ho + wy) * Wo + wx) * Cout + co_blk + c4);
          if (scale != nullptr) {
            const float4 sv = ldg4(scale + (size_t)b * Cout + co_blk + c4);
            dv.x *= sv.x; dv.y *= sv.y; dv.z *= sv.z; dv.w *= sv.w;
          }
        }
      }
      *reinterpret_cast<float4*>(&xs[pp][c4]) = xv;
      *reinterpret_cast<float4*>(&ds[pp][c4]) = dv;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < PC; ++pp) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[pp][tci]);
      const float4 d = *reinterpret_cast<const float4*>(&ds[pp][tco]);
      const float av[4] = {a.x, a.y, a.z, a.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], dv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci_blk + tci + i;
    if (ci >= Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co_blk + tco + j;
      if (co < Cout) atomicAdd(dw + ((size_t)ci * Cout + co) * 9 + tap, acc[i][j]);
    }
  }
}

// db[co] = sum_{b,pixels} dy[b,p,co] * s[b,co]
__global__ void bias_grad_scaled_kernel(const float* __restrict__ dy, const float* __restrict__ scale, float* __restrict__ db,
                                        int B, long long hw, int C) {
  // block handles a slab of pixels; thread -> channel (threadIdx.x % C), sub-stream threadIdx.x / C
  extern __shared__ float red[];
  const int c = threadIdx.x % C;
  const int sub = threadIdx.x / C;
  const int nsub = blockDim.x / C;
  const long long npix = (long long)B * hw;
  float acc = 0.f;
  if (sub < nsub) {
    for (long long p = (long long)blockIdx.x * nsub + sub; p < npix; p += (long long)gridDim.x * nsub) {
      float v = __ldg(dy + p * C + c);
      if (scale != nullptr) v *= __ldg(scale + (p / hw) * C + c);
      acc += v;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (sub == 0) {
    float s = 0.f;
    for (int u = 0; u < nsub; ++u) s += red[u * C + c];
    atomicAdd(db + c, s);
  }
}

static int launch_bias_grad(const float* dy, const float* scale, float* db, int B, long long hw, int C, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * C, st);
  if (e != cudaSuccess) {
    set_error("bias_grad memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  PU_REQUIRE(C <= 1024, PU_ERR_UNSUPPORTED, "bias_grad: C=%d > 1024", C);
  int bs = C >= 256 ? C : (256 / C) * C;
  const int nsub = bs / C;
  const long long npix = (long long)B * hw;
  long long blocks = (npix + (long long)nsub * 32 - 1) / ((long long)nsub * 32);
  if (blocks < 1) blocks = 1;
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  bias_grad_scaled_kernel<<<(unsigned)blocks, bs, bs * sizeof(float), st>>>(dy, scale, db, B, hw, C);
  return post_launch("bias_grad");
}

static int split_for(long long nout_blocks, long long npix) {
  long long want = (4LL * kNumSMs + nout_blocks - 1) / nout_blocks;
  if (want < 1) want = 1;
  long long maxsplit = (npix + 15) / 16;  // at least 16 pixels per split
  if (maxsplit < 1) maxsplit = 1;
  if (want > maxsplit) want = maxsplit;
  if (want > 65535) want = 65535;
  return (int)want;
}

// convT_mma.cu
bool convT2x2_mma_ok(const float* x, const float* dy_or_y, int Cin, int Cout, long long npix);
int convT2x2_fwd_mma(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin, int Cout, int round_out,
                     cudaStream_t st);
int convT2x2_dx_mma(const float* x, const float* w, const float* dy, float* dx, int B, int H, int W, int Cin, int Cout, int mask_in,
                    cudaStream_t st);
int convT2x2_dw_mma(const float* x, const float* dy, float* dw, float* db, int B, int H, int W, int Cin, int Cout, cudaStream_t st);
// convT_dw_tma.cu
bool convT2x2_dw_tma_ok(int Cin, int Cout);
int convT2x2_dw_tma(const float* x, const float* dy, float* dw, float* db, int B, int H, int W, int Cin, int Cout, cudaStream_t st);

}  // namespace pu

extern "C" {

int pu_convT2x2s2_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin, int Cout, int flags,
                      void* stream) {
  PU_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, PU_ERR_BAD_ARG, "pu_convT2x2s2_fwd: bad argument");
  const size_t smem = (size_t)4 * Cin * 8 * sizeof(float);
  PU_REQUIRE(smem <= 48 * 1024, PU_ERR_UNSUPPORTED, "pu_convT2x2s2_fwd: Cin=%d > 384", Cin);
  PU_REQUIRE(pu::aligned16(x) && pu::aligned16(y), PU_ERR_BAD_ARG, "pu_convT2x2s2_fwd: pointers not 16-byte aligned");
  if ((flags & PU_FLAG_TF32_MATH) && pu::convT2x2_mma_ok(x, y, Cin, Cout, (long long)B * H * W))
    return pu::convT2x2_fwd_mma(x, w, bias, y, B, H, W, Cin, Cout, (flags & PU_FLAG_ROUND_TF32) ? 1 : 0, pu::as_stream(stream));
  const long long npix = (long long)B * 4 * H * W;
  dim3 grid((unsigned)((npix + 255) / 256), pu::cdiv(Cout, 8));
  pu::convT2x2_fwd_kernel<<<grid, 256, smem, pu::as_stream(stream)>>>(x, w, bias, y, B, H, W, Cin, Cout, flags);
  return pu::post_launch("pu_convT2x2s2_fwd");
}

int pu_convT2x2s2_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int B, int H, int W,
                      int Cin, int Cout, int flags, void* stream) {
  PU_REQUIRE(x && w && dy && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, PU_ERR_BAD_ARG, "pu_convT2x2s2_bwd: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  const long long npix = (long long)B * H * W;
  const bool mma = (flags & PU_FLAG_TF32_MATH) && pu::convT2x2_mma_ok(x, dy, Cin, Cout, npix) && (dx == nullptr || pu::aligned16(dx));
  if (mma) {
    if (dx != nullptr) {
      int rc = pu::convT2x2_dx_mma(x, w, dy, dx, B, H, W, Cin, Cout, (flags & PU_FLAG_MASK_IN) ? 1 : 0, st);
      if (rc) return rc;
    }
    if (dw != nullptr) {
      cudaError_t e = cudaSuccess;
      if (!(flags & PU_FLAG_ACCUM_GRADS)) {
        e = cudaMemsetAsync(dw, 0, sizeof(float) * Cin * Cout * 4, st);
        if (e == cudaSuccess && db != nullptr) e = cudaMemsetAsync(db, 0, sizeof(float) * Cout, st);
      }
      if (e != cudaSuccess) {
        pu::set_error("pu_convT2x2s2_bwd memset: %s", cudaGetErrorString(e));
        return PU_ERR_CUDA;
      }
      static int ver = -1;  // PU_CONVT_DW_V=1: the first, cp.async-fed kernels (A/B measurements)
      if (ver < 0) {
        const char* e_ = getenv("PU_CONVT_DW_V");
        ver = e_ ? atoi(e_) : 2;
      }
      if (ver != 1 && pu::convT2x2_dw_tma_ok(Cin, Cout)) return pu::convT2x2_dw_tma(x, dy, dw, db, B, H, W, Cin, Cout, st);
      return pu::convT2x2_dw_mma(x, dy, dw, db, B, H, W, Cin, Cout, st);
    }
    if (db != nullptr) {
      PU_REQUIRE(!(flags & PU_FLAG_ACCUM_GRADS), PU_ERR_UNSUPPORTED, "pu_convT2x2s2_bwd: PU_FLAG_ACCUM_GRADS needs dw and db together");
      return pu::launch_bias_grad(dy, nullptr, db, B, 4LL * H * W, Cout, st);
    }
    return PU_OK;
  }
  PU_REQUIRE(!(flags & PU_FLAG_ACCUM_GRADS) || (dw == nullptr && db == nullptr), PU_ERR_UNSUPPORTED,
             "pu_convT2x2s2_bwd: PU_FLAG_ACCUM_GRADS needs the TF32 tensor-core path for this shape");
  if (dx != nullptr) {
    const size_t smem = (size_t)4 * Cout * 8 * sizeof(float);
    PU_REQUIRE(smem <= 48 * 1024, PU_ERR_UNSUPPORTED, "pu_convT2x2s2_bwd: Cout=%d > 384", Cout);
    dim3 grid((unsigned)((npix + 255) / 256), pu::cdiv(Cin, 8));
    pu::convT2x2_dx_kernel<<<grid, 256, smem, st>>>(dy, w, dx, B, H, W, Cin, Cout, (flags & PU_FLAG_MASK_IN) ? x : nullptr);
    int rc = pu::post_launch("pu_convT2x2s2_bwd dx");
    if (rc) return rc;
  }
  if (dw != nullptr) {
    const int nout = Cin * Cout * 4;
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * nout, st);
    if (e != cudaSuccess) {
      pu::set_error("pu_convT2x2s2_bwd memset: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
    if (Cin % 4 == 0 && Cout % 4 == 0 && pu::aligned16(x) && pu::aligned16(dy)) {
      int pch = 10240 / (Cin + 4 * Cout);
      pch = pch > 64 ? 64 : (pch < 4 ? 4 : pch);
      const size_t smem = (size_t)pch * (Cin + 4 * Cout) * sizeof(float);
      PU_REQUIRE(smem <= 48 * 1024, PU_ERR_UNSUPPORTED, "pu_convT2x2s2_bwd: Cin+4*Cout=%d too large", Cin + 4 * Cout);
      int nrep = 1;
      while (nrep < 8 && (long long)Cin * Cout * nrep * 2 <= 256) nrep *= 2;  // replicate small output sets over the pixels
      const int slices = 256 / nrep;
      const int nb = pu::cdiv((long long)Cin * Cout, slices);  // 4 outputs per thread
      long long splits = (4LL * pu::kNumSMs + nb - 1) / nb;
      const long long maxsplit = (npix + pch - 1) / pch;
      if (splits > maxsplit) splits = maxsplit;
      if (splits < 1) splits = 1;
      if (splits > 65535) splits = 65535;
      dim3 grid(nb, (unsigned)splits);
      pu::convT2x2_dw_tiled_kernel<<<grid, 256, smem, st>>>(x, dy, dw, B, H, W, Cin, Cout, pch, nrep);
    } else {
      const int nb = pu::cdiv(nout, 256);
      dim3 grid(nb, pu::split_for(nb, npix));
      pu::convT2x2_dw_kernel<<<grid, 256, 0, st>>>(x, dy, dw, B, H, W, Cin, Cout);
    }
    int rc = pu::post_launch("pu_convT2x2s2_bwd dw");
    if (rc) return rc;
  }
  if (db != nullptr) return pu::launch_bias_grad(dy, nullptr, db, B, 4LL * H * W, Cout, st);
  return PU_OK;
}

int pu_convT3x3s2_fwd(const float* x, const float* w, const float* bias, const float* chan_scale, float* y, int B, int H, int W,
                      int Cin, int Cout, int Ho, int Wo, int oy, int ox, int flags, void* stream) {
  PU_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, PU_ERR_BAD_ARG, "pu_convT3x3s2_fwd: bad argument");
  PU_REQUIRE(oy >= 0 && ox >= 0 && Ho > 0 && Wo > 0 && oy + Ho <= 2 * H + 1 && ox + Wo <= 2 * W + 1, PU_ERR_BAD_ARG,
             "pu_convT3x3s2_fwd: window (%d+%d,%d+%d) exceeds %dx%d", oy, Ho, ox, Wo, 2 * H + 1, 2 * W + 1);
  const size_t smem = (size_t)9 * 32 * 8 * sizeof(float);
  const long long npix = (long long)B * Ho * Wo;
  dim3 grid((unsigned)((npix + 127) / 128), pu::cdiv(Cout, 8));
  pu::convT3x3_fwd_kernel<<<grid, 128, smem, pu::as_stream(stream)>>>(x, w, bias, chan_scale, y, B, H, W, Cin, Cout, Ho, Wo, oy, ox, flags);
  return pu::post_launch("pu_convT3x3s2_fwd");
}

int pu_convT3x3s2_bwd(const float* x, const float* w, const float* dy, const float* chan_scale, float* dx, float* dw, float* db,
                      int B, int H, int W, int Cin, int Cout, int Ho, int Wo, int oy, int ox, void* stream) {
  PU_REQUIRE(x && w && dy && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, PU_ERR_BAD_ARG, "pu_convT3x3s2_bwd: bad argument");
  PU_REQUIRE(oy >= 0 && ox >= 0 && Ho > 0 && Wo > 0 && oy + Ho <= 2 * H + 1 && ox + Wo <= 2 * W + 1, PU_ERR_BAD_ARG,
             "pu_convT3x3s2_bwd: window exceeds output");
  cudaStream_t st = pu::as_stream(stream);
  const long long npix = (long long)B * H * W;
  if (dx != nullptr) {
    const size_t smem = (size_t)9 * 32 * 8 * sizeof(float);
    dim3 grid((unsigned)((npix + 127) / 128), pu::cdiv(Cin, 8));
    pu::convT3x3_dx_kernel<<<grid, 128, smem, st>>>(dy, w, chan_scale, dx, B, H, W, Cin, Cout, Ho, Wo, oy, ox);
    int rc = pu::post_launch("pu_convT3x3s2_bwd dx");
    if (rc) return rc;
  }
  if (dw != nullptr) {
    const long long nout = (long long)Cin * Cout * 9;
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * nout, st);
    if (e != cudaSuccess) {
      pu::set_error("pu_convT3x3s2_bwd memset: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
    if (Cin % 4 == 0 && Cout % 4 == 0 && pu::aligned16(x) && pu::aligned16(dy)) {
      const int gx = pu::cdiv(Cin, 64), gy = pu::cdiv(Cout, 64);
      long long nsplit = (4LL * pu::kNumSMs + 9LL * gx * gy - 1) / (9LL * gx * gy);
      const long long maxsplit = (npix + 63) / 64;
      if (nsplit > maxsplit) nsplit = maxsplit;
      if (nsplit < 1) nsplit = 1;
      if (9 * nsplit > 65535) nsplit = 65535 / 9;
      dim3 grid(gx, gy, (unsigned)(9 * nsplit));
      pu::convT3x3_dw_tiled_kernel<<<grid, 256, 0, st>>>(x, dy, chan_scale, dw, B, H, W, Cin, Cout, Ho, Wo, oy, ox, (int)nsplit);
    } else {
      const long long nb = (nout + 255) / 256;
      dim3 grid((unsigned)nb, pu::split_for(nb, npix));
      pu::convT3x3_dw_kernel<<<grid, 256, 0, st>>>(x, dy, chan_scale, dw, B, H, W, Cin, Cout, Ho, Wo, oy, ox);
    }
    int rc = pu::post_launch("pu_convT3x3s2_bwd dw");
    if (rc) return rc;
  }
  if (db != nullptr) return pu::launch_bias_grad(dy, chan_scale, db, B, (long long)Ho * Wo, Cout, st);
  return PU_OK;
}

}  // extern "C"
