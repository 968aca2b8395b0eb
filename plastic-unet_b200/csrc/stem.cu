// stem.cu — the network's first 3x3 convolution (one input channel: reference unet_p.py:105 via inconv, unet_p.py:124-132).
//
// With Cin = 1 there is no reduction to feed a tensor core (K = 9): the layer is a pure streaming op — read 4 bytes,
// write 4*Cout bytes per pixel forward; read x and g once for the weight gradient.  The generic kernels (shared-memory
// channel planes / pixel-major tiles built for Cin >= 8) spent 22 us (forward) and 42 + 13 us (wgrad + separate bias
// pass) on it; these one-thread-per-pixel kernels stay close to the HBM time of the 33.5 MB activation.
#include <stdlib.h>
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
  // the destination is the whole tensor and warps never straddle the end: a warp's 32 consecutive pixels are contiguous memory
  const bool warp_rows = d0.Hs == H && d0.Ws == W && d0.oy == 0 && d0.ox == 0 && d0.C == CO && (npix & 31) == 0;
  // (b, y, x) of the thread's pixel advance incrementally: three 64-bit divisions per pixel cost more than the 72 FMAs
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int sx = (int)(stride % W), sy = (int)((stride / W) % H), sb = (int)(stride / ((long long)W * H));
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
  for (; p < npix; p += stride, x += sx, y += sy, b += sb) {
    if (x >= W) { x -= W; ++y; }
    if (y >= H) { y -= H; ++b; }
    float xv[9];
    const float* ctr = s0.p + ((size_t)b * s0.Hs + (y + s0.oy)) * s0.Ws + (x + s0.ox);
    if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {  // interior pixel: nine unchecked loads around one pointer
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) xv[ky * 3 + kx] = __ldg(ctr + (ky - 1) * s0.Ws + (kx - 1));
    } else {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int gy = y + ky - 1, gx = x + kx - 1;
          const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
          xv[ky * 3 + kx] = ok ? __ldg(ctr + (ky - 1) * s0.Ws + (kx - 1)) : 0.f;
        }
    }
    // packed FFMA2: two output channels per instruction, each lane pair is an ordinary fp32 FMA (same rounding, same order)
    float2 acc2[CO / 2];
#pragma unroll
    for (int j = 0; j < CO / 2; ++j) acc2[j] = make_float2(bs[2 * j], bs[2 * j + 1]);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float2 xx = make_float2(xv[tap], xv[tap]);
#pragma unroll
      for (int j = 0; j < CO / 2; ++j) acc2[j] = __ffma2_rn(xx, *reinterpret_cast<const float2*>(&ws[tap][2 * j]), acc2[j]);
    }
    float acc[CO];
#pragma unroll
    for (int j = 0; j < CO / 2; ++j) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
    const size_t opix = ((size_t)b * d0.Hs + (y + d0.oy)) * d0.Ws + (x + d0.ox);
    float* o = d0.p + opix * d0.C;
    unsigned m = 0;
    float4 vq[CO / 4];
#pragma unroll
    for (int j = 0; j < CO; j += 4) {
      float4 v = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
      vq[j / 4] = v;
      m |= ((v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u)) << j;
    }
    if (CO == 8 && warp_rows) {
      // The 32 pixels of a warp are 1 KB of contiguous output.  Written straight from the lanes, every 128-bit store instruction
      // fills HALF of each 32-byte sector (lane stride 32 B); a shuffle transpose lets each of the two instructions write 512
      // contiguous bytes: lane l stores float4 number l (pixel l / 2, half l % 2), then number 32 + l.
      const int lane = threadIdx.x & 31, src = lane >> 1;
      const bool odd = lane & 1;
      float4 lo, hi;
#define PU_STEM_PICK(dst, comp, from)                                         \
      {                                                                       \
        const float t0 = __shfl_sync(0xffffffffu, vq[0].comp, from);          \
        const float t1 = __shfl_sync(0xffffffffu, vq[CO / 4 - 1].comp, from); \
        dst.comp = odd ? t1 : t0;                                             \
      }
      PU_STEM_PICK(lo, x, src) PU_STEM_PICK(lo, y, src) PU_STEM_PICK(lo, z, src) PU_STEM_PICK(lo, w, src)
      PU_STEM_PICK(hi, x, 16 + src) PU_STEM_PICK(hi, y, 16 + src) PU_STEM_PICK(hi, z, 16 + src) PU_STEM_PICK(hi, w, 16 + src)
#undef PU_STEM_PICK
      float* wbase = o - (size_t)lane * CO;  // the warp's first pixel
      *reinterpret_cast<float4*>(wbase + 4 * lane) = lo;
      *reinterpret_cast<float4*>(wbase + 128 + 4 * lane) = hi;
    } else {
#pragma unroll
      for (int j = 0; j < CO; j += 4) *reinterpret_cast<float4*>(o + j) = vq[j / 4];
    }
    if (mask_out != nullptr) {  // packed ReLU mask of the output, one byte per 8 channels
#pragma unroll
      for (int g8 = 0; g8 < CO / 8; ++g8) mask_out[opix * (CO / 8) + g8] = (unsigned char)((m >> (8 * g8)) & 0xffu);
    }
  }
}

// dw[co][0][tap] = sum_p x[p + tap] * g[p][co],  db[co] = sum_p g[p][co];  thread = strip of pixels, 10*CO partial sums in
// registers, warp shuffle + shared-memory reduction, then ONE vector reduction (red.global.add.v4.f32) per four outputs and CTA
// (dw/db zeroed by the caller).  Two CTAs per SM (8 channels): with 4 CTAs of 128 threads per SM (first version) the 592 x 80 scalar atomics
// on three cache lines cost ~8 us of a 28 us kernel, and this kernel is the LAST weight gradient of the backward pass — nothing
// is left to overlap it with.  The next pixel's g vector is requested before the current one is consumed.
template <int CO>
__global__ void __launch_bounds__(256, CO == 8 ? 2 : 1) conv3x3_c1_wgrad_kernel(const View s0, const float* __restrict__ g, float* __restrict__ dw,
                                                                  float* __restrict__ db, int B, int H, int W, int vec4) {
  float acc[10][CO];  // 9 taps + the bias row
#pragma unroll
  for (int t = 0; t < 10; ++t)
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[t][j] = 0.f;
  const long long npix = (long long)B * H * W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long p0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // PF pixels of g in flight per thread (the loads of pixel i + PF are issued before pixel i is consumed): with 512 threads per SM
  // one 32-byte vector each is 16 KB in flight per SM, a third of what the HBM latency needs
  constexpr int PF = 3;
  float gn[PF][CO];
#pragma unroll
  for (int d = 0; d < PF; ++d) {
    const long long q = p0 + d * stride;
    if (q < npix) {
#pragma unroll
      for (int j = 0; j < CO; j += 4) {
        const float4 v = ldg4_stream(g + q * CO + j);
        gn[d][j] = v.x; gn[d][j + 1] = v.y; gn[d][j + 2] = v.z; gn[d][j + 3] = v.w;
      }
    }
  }
  // (b, y, x) advance incrementally with the pixel index: no 64-bit divisions in the loop
  const int sx = (int)(stride % W), sy = (int)((stride / W) % H), sb = (int)(stride / ((long long)W * H));
  int x = (int)(p0 % W), y = (int)((p0 / W) % H), b = (int)(p0 / ((long long)W * H));
  for (long long p = p0; p < npix; p += PF * stride) {
#pragma unroll
    for (int d = 0; d < PF; ++d, x += sx, y += sy, b += sb) {
      const long long q = p + d * stride;
      if (x >= W) { x -= W; ++y; }
      if (y >= H) { y -= H; ++b; }
      if (q < npix) {
        float xv[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int gy = y + ky - 1, gx = x + kx - 1;
            const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
            xv[ky * 3 + kx] = ok ? __ldg(s0.p + ((size_t)b * s0.Hs + (gy + s0.oy)) * s0.Ws + (gx + s0.ox)) : 0.f;
          }
        float gv[CO];
#pragma unroll
        for (int j = 0; j < CO; ++j) gv[j] = gn[d][j];
        const long long qn = q + PF * stride;
        if (qn < npix) {
#pragma unroll
          for (int j = 0; j < CO; j += 4) {
            const float4 v = ldg4_stream(g + qn * CO + j);
            gn[d][j] = v.x; gn[d][j + 1] = v.y; gn[d][j + 2] = v.z; gn[d][j + 3] = v.w;
          }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int j = 0; j < CO; ++j) acc[t][j] = fmaf(xv[t], gv[j], acc[t][j]);
#pragma unroll
        for (int j = 0; j < CO; ++j) acc[9][j] += gv[j];
      }
    }
  }
  __shared__ __align__(16) float red[8][10 * CO];  // per warp, in memory order: dw[j][tap] = [j * 9 + tap], then db[j] at 9 * CO + j
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 10; ++t)
#pragma unroll
    for (int j = 0; j < CO; ++j) {
      const float s = warp_sum(acc[t][j]);
      if (lane == 0) red[warp][t < 9 ? j * 9 + t : 9 * CO + j] = s;
    }
  __syncthreads();
  if (vec4) {
    for (int i = threadIdx.x; i < 10 * CO / 4; i += blockDim.x) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float4 v = *reinterpret_cast<const float4*>(&red[w][4 * i]);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      float* dst = 4 * i < 9 * CO ? dw + 4 * i : (db != nullptr ? db + (4 * i - 9 * CO) : nullptr);  // 9 * CO is a multiple of 4
      if (dst != nullptr) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
    }
  } else {
    for (int i = threadIdx.x; i < 10 * CO; i += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][i];
      if (i < 9 * CO) atomicAdd(dw + i, s);
      else if (db != nullptr) atomicAdd(db + (i - 9 * CO), s);
    }
  }
}

bool conv3x3_c1_ok(int Cin, int Cout) { return Cin == 1 && (Cout == 8 || Cout == 16); }

int conv3x3_c1_fwd(const Conv3x3Args& a, cudaStream_t st) {
  const long long npix = (long long)a.B * a.H * a.W;
  long long blocks = (npix + 255) / 256;
  // blocks per SM of the grid-stride loop (PU_STEM_CAP: tuning override).  Measured on B200 (1 -> 8 @128x128, B = 64, pack + conv
  // per launch): 4: 13.8 us, 8: 14.5, 16: 16.2, 28 and more (one pixel per thread): 19.8 — the write stream of this 33 MB-out /
  // 4 MB-in layer prefers few resident warps
  static long long cap = -1;
  if (cap < 0) {
    const char* e = getenv("PU_STEM_CAP");
    cap = e != nullptr && atoi(e) > 0 ? atoi(e) : 4;
  }
  if (blocks > cap * kNumSMs) blocks = cap * kNumSMs;
  if (a.Cout == 8) conv3x3_c1_fwd_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(a.s0, a.wp, a.bias, a.d0, a.mask_out, a.B, a.H, a.W, a.relu, a.round_out);
  else conv3x3_c1_fwd_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(a.s0, a.wp, a.bias, a.d0, a.mask_out, a.B, a.H, a.W, a.relu, a.round_out);
  return post_launch("pu_conv3x3_fwd (stem)");
}

// dw and db are zeroed here (unless the caller accumulates: PU_MATH_ACCUM)
int conv3x3_c1_wgrad(const WgradArgs& a, cudaStream_t st, int math) {
  cudaError_t e = cudaSuccess;
  if (!a.accum) {
    e = cudaMemsetAsync(a.dw, 0, sizeof(float) * a.Cout * 9, st);
    if (e == cudaSuccess && a.db != nullptr) e = cudaMemsetAsync(a.db, 0, sizeof(float) * a.Cout, st);
  }
  if (e != cudaSuccess) {
    set_error("conv3x3_wgrad (stem) memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  {
    static int ver = -1;  // PU_STEM_WGRAD_V=1: always the streaming kernel below (A/B measurements)
    if (ver < 0) {
      const char* e_ = getenv("PU_STEM_WGRAD_V");
      ver = e_ ? atoi(e_) : 2;
    }
    // TF32 mode only: the strict-fp32 mode keeps the FFMA kernel below
    if (math == PU_MATH_TF32 && ver != 1 && conv3x3_c1_wgrad_tma_ok(a)) return conv3x3_c1_wgrad_tma(a, st);
  }
  const long long npix = (long long)a.B * a.H * a.W;
  long long blocks = (npix + 256 * 16 - 1) / (256 * 16);  // >= 16 pixels per thread amortise the 10*CO-value reduction
  const long long per_sm = a.Cout == 8 ? 2 : 1;  // 16 output channels: 160 accumulators per thread, one CTA per SM
  if (blocks > per_sm * kNumSMs) blocks = per_sm * kNumSMs;
  if (blocks < 1) blocks = 1;
  const int vec4 = ((reinterpret_cast<uintptr_t>(a.dw) | (a.db != nullptr ? reinterpret_cast<uintptr_t>(a.db) : 0)) & 15u) == 0 ? 1 : 0;
  if (a.Cout == 8) conv3x3_c1_wgrad_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(a.s0, a.g, a.dw, a.db, a.B, a.H, a.W, vec4);
  else conv3x3_c1_wgrad_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(a.s0, a.g, a.dw, a.db, a.B, a.H, a.W, vec4);
  return post_launch("pu_conv3x3_wgrad (stem)");
}

}  // namespace pu
