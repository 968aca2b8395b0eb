// stem.cu — the network's first 3x3 convolution (one input channel: reference unet_p.py:105 via inconv, unet_p.py:124-132).
//
// With Cin = 1 there is no reduction to feed a tensor core (K = 9): the layer is a pure streaming op — read 4 bytes,
// write 4*Cout bytes per pixel forward; read x and g once for the weight gradient.  The generic kernels (shared-memory
// channel planes / pixel-major tiles built for Cin >= 8) spent 22 us (forward) and 42 + 13 us (wgrad + separate bias
// pass) on it; these one-thread-per-pixel kernels stay close to the HBM time of the 33.5 MB activation.
#include "conv3x3.cuh"

namespace pu {

// y[p][co] = act(bias[co] + sum_tap x[p + tap] * w[co][0][tap]);  CO output channels per pixel, thread = pixel
template <int CO>
__global__ void __launch_bounds__(256) conv3x3_c1_fwd_kernel(const View s0, const float* __restrict__ w, const float* __restrict__ bias,
                                                             const ViewW d0, unsigned char* __restrict__ mask_out, int B, int H, int W,
                                                             int relu, int round_out) {
  __shared__ __align__(16) float ws[9][CO];
  __shared__ __align__(16) float bs[CO];
  for (int i = threadIdx.x; i < 9 * CO; i += blockDim.x) {
    const int co = i / 9, tap = i - co * 9;
    ws[tap][co] = w[i];
  }
  for (int i = threadIdx.x; i < CO; i += blockDim.x) bs[i] = bias != nullptr ? bias[i] : 0.f;
  __syncthreads();
  const long long npix = (long long)B * H * W;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const long long b = p / ((long long)W * H);
    float xv[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int gy = y + ky - 1, gx = x + kx - 1;
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
        xv[ky * 3 + kx] = ok ? __ldg(s0.p + ((size_t)b * s0.Hs + (gy + s0.oy)) * s0.Ws + (gx + s0.ox)) : 0.f;
      }
    float acc[CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[j] = bs[j];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int j = 0; j < CO; ++j) acc[j] = fmaf(xv[tap], ws[tap][j], acc[j]);
    const size_t opix = ((size_t)b * d0.Hs + (y + d0.oy)) * d0.Ws + (x + d0.ox);
    float* o = d0.p + opix * d0.C;
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < CO; j += 4) {
      float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
      *reinterpret_cast<float4*>(o + j) = v;
      m |= ((v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u)) << j;
    }
    if (mask_out != nullptr) {  // packed ReLU mask of the output, one byte per 8 channels
#pragma unroll
      for (int g8 = 0; g8 < CO / 8; ++g8) mask_out[opix * (CO / 8) + g8] = (unsigned char)((m >> (8 * g8)) & 0xffu);
    }
  }
}

// dw[co][0][tap] = sum_p x[p + tap] * g[p][co],  db[co] = sum_p g[p][co];  thread = strip of pixels, 10*CO partial sums in
// registers, warp shuffle + shared-memory reduction, one atomic per output and CTA (dw/db zeroed by the caller)
template <int CO>
__global__ void __launch_bounds__(128) conv3x3_c1_wgrad_kernel(const View s0, const float* __restrict__ g, float* __restrict__ dw,
                                                               float* __restrict__ db, int B, int H, int W) {
  float acc[10][CO];  // 9 taps + the bias row
#pragma unroll
  for (int t = 0; t < 10; ++t)
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[t][j] = 0.f;
  const long long npix = (long long)B * H * W;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const long long b = p / ((long long)W * H);
    float gv[CO];
#pragma unroll
    for (int j = 0; j < CO; j += 4) {
      const float4 v = ldg4(g + p * CO + j);
      gv[j] = v.x; gv[j + 1] = v.y; gv[j + 2] = v.z; gv[j + 3] = v.w;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int gy = y + ky - 1, gx = x + kx - 1;
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const float xv = ok ? __ldg(s0.p + ((size_t)b * s0.Hs + (gy + s0.oy)) * s0.Ws + (gx + s0.ox)) : 0.f;
#pragma unroll
        for (int j = 0; j < CO; ++j) acc[ky * 3 + kx][j] = fmaf(xv, gv[j], acc[ky * 3 + kx][j]);
      }
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[9][j] += gv[j];
  }
  __shared__ float red[4][10 * CO];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 10; ++t)
#pragma unroll
    for (int j = 0; j < CO; ++j) {
      const float s = warp_sum(acc[t][j]);
      if (lane == 0) red[warp][t * CO + j] = s;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * CO; i += blockDim.x) {
    const float s = red[0][i] + red[1][i] + red[2][i] + red[3][i];
    const int t = i / CO, j = i - t * CO;
    if (t < 9) atomicAdd(dw + j * 9 + t, s);
    else if (db != nullptr) atomicAdd(db + j, s);
  }
}

bool conv3x3_c1_ok(int Cin, int Cout) { return Cin == 1 && (Cout == 8 || Cout == 16); }

int conv3x3_c1_fwd(const Conv3x3Args& a, cudaStream_t st) {
  const long long npix = (long long)a.B * a.H * a.W;
  long long blocks = (npix + 255) / 256;
  if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
  if (a.Cout == 8) conv3x3_c1_fwd_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(a.s0, a.wp, a.bias, a.d0, a.mask_out, a.B, a.H, a.W, a.relu, a.round_out);
  else conv3x3_c1_fwd_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(a.s0, a.wp, a.bias, a.d0, a.mask_out, a.B, a.H, a.W, a.relu, a.round_out);
  return post_launch("pu_conv3x3_fwd (stem)");
}

// dw and db are zeroed here (unless the caller accumulates: PU_MATH_ACCUM)
int conv3x3_c1_wgrad(const WgradArgs& a, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  if (!a.accum) {
    e = cudaMemsetAsync(a.dw, 0, sizeof(float) * a.Cout * 9, st);
    if (e == cudaSuccess && a.db != nullptr) e = cudaMemsetAsync(a.db, 0, sizeof(float) * a.Cout, st);
  }
  if (e != cudaSuccess) {
    set_error("conv3x3_wgrad (stem) memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  const long long npix = (long long)a.B * a.H * a.W;
  long long blocks = (npix + 128 * 16 - 1) / (128 * 16);  // >= 16 pixels per thread amortise the 10*CO-value reduction
  if (blocks > 4LL * kNumSMs) blocks = 4LL * kNumSMs;
  if (blocks < 1) blocks = 1;
  if (a.Cout == 8) conv3x3_c1_wgrad_kernel<8><<<(unsigned)blocks, 128, 0, st>>>(a.s0, a.g, a.dw, a.db, a.B, a.H, a.W);
  else conv3x3_c1_wgrad_kernel<16><<<(unsigned)blocks, 128, 0, st>>>(a.s0, a.g, a.dw, a.db, a.B, a.H, a.W);
  return post_launch("pu_conv3x3_wgrad (stem)");
}

}  // namespace pu
