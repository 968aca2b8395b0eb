// conv3x3_ffma.cu — strict-fp32 CUDA-core 3x3 convolution family (NHWC), sm_100a.
//
// This is the parity path (PU_MATH_FP32) and the small-shape path (C_in = 1 stem, ragged channel
// counts, 6x6 / 8x8 bottlenecks).  Forward, dgrad (same kernel, transposed+flipped packed weights)
// and wgrad.  Replaces nn.Conv2d(k=3,pad=1) + ReLU + residual add + cat/crop at
// reference unet_p.py:105-116,161-166; unet_p_res.py:150-158,186-189,215-219,230,264.
//
// Forward tiling: one CTA = TH x TW output pixels x 8 output channels.  Input channels are staged
// 8 at a time into shared memory as channel planes [ci][hy][hx] (halo included, zero padded), so the
// per-lane reads are stride-1 (bank-conflict free); weights [tap][ci][8 co] are read as two
// broadcast LDS.128.  Each thread owns RP rows x 1 column x 8 channels = RP*8 fp32 accumulators:
// per input channel 3*(RP+2) scalar LDS + 18 LDS.128 feed 72*RP FFMA.
#include "pu_common.cuh"
#include "conv3x3.cuh"

namespace pu {

template <int TW, int RP, int NT>
__global__ void __launch_bounds__(NT) conv3x3_ffma_kernel(const Conv3x3Args a) {
  constexpr int TH = (NT / TW) * RP;
  constexpr int HW_ = TW + 2;  // halo width
  constexpr int HH_ = TH + 2;
  constexpr int PLANE = HH_ * HW_;
  __shared__ float in_s[8 * PLANE];
  __shared__ __align__(16) float w_s[9 * 8 * 8];

  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int tx = t % a.tilesX;
  t /= a.tilesX;
  const int ty = t % a.tilesY;
  const int b = t / a.tilesY;
  const int x0 = tx * TW, y0 = ty * TH;
  const int co0 = blockIdx.y * 8;

  const int px = tid % TW;
  const int r0 = (tid / TW) * RP;

  // packed fp32 FMA (FFMA2: two FMAs per issue slot on sm_100): accumulators as float2 pairs of output channels
  float2 acc2[RP][4];
#pragma unroll
  for (int r = 0; r < RP; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[r][j] = make_float2(0.f, 0.f);

  int cbase = 0;  // channel index in the concatenated input
  for (int s = 0; s < 2; ++s) {
    const View v = s == 0 ? a.s0 : a.s1;
    if (v.p == nullptr || v.C == 0) continue;
    const bool vec = (v.C % 4 == 0);
    for (int c0 = 0; c0 < v.C; c0 += 8) {
      const int cc = min(8, v.C - c0);
      __syncthreads();  // previous chunk fully consumed
      // ---- stage the input halo tile as channel planes
      for (int i = tid; i < PLANE; i += NT) {
        const int hy = i / HW_, hx = i - hy * HW_;
        const int gy = y0 + hy - 1, gx = x0 + hx - 1;
        float vals[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) vals[c] = 0.f;
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const float* p = v.p + (((size_t)b * v.Hs + (gy + v.oy)) * v.Ws + (gx + v.ox)) * v.C + c0;
          if (vec) {
            const float4 q0 = ldg4(p);
            vals[0] = q0.x; vals[1] = q0.y; vals[2] = q0.z; vals[3] = q0.w;
            if (cc > 4) {
              const float4 q1 = ldg4(p + 4);
              vals[4] = q1.x; vals[5] = q1.y; vals[6] = q1.z; vals[7] = q1.w;
            }
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < cc) vals[c] = __ldg(p + c);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) in_s[c * PLANE + i] = vals[c];
      }
      // ---- stage the weights of this chunk: w_s[tap][ci][j]
      for (int i = tid; i < 9 * 8 * 8; i += NT) {
        const int j = i & 7, ci = (i >> 3) & 7, tap = i >> 6;
        float wv = 0.f;
        if (ci < cc && co0 + j < a.Cout) {
          const int cig = cbase + c0 + ci, cog = co0 + j;
          if (a.wfmt == 0) wv = __ldg(a.wp + ((size_t)tap * a.Cin + cig) * a.Cout + cog);           // packed [9][Cin][Cout]
          else if (a.wfmt == 1) wv = __ldg(a.wp + ((size_t)cog * a.Cin + cig) * 9 + tap);           // raw OIHW, forward
          else wv = __ldg(a.wp + ((size_t)cig * a.Cout + cog) * 9 + (8 - tap));                     // raw OIHW, dgrad
        }
        w_s[i] = wv;
      }
      __syncthreads();
      // ---- accumulate
      for (int ci = 0; ci < cc; ++ci) {
        float xin[RP + 2][3];
        const float* ip = in_s + ci * PLANE + r0 * HW_ + px;
#pragma unroll
        for (int i = 0; i < RP + 2; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) xin[i][j] = ip[i * HW_ + j];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 wa = *reinterpret_cast<const float4*>(w_s + ((ky * 3 + kx) * 8 + ci) * 8);
            const float4 wb = *reinterpret_cast<const float4*>(w_s + ((ky * 3 + kx) * 8 + ci) * 8 + 4);
            const float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w);
            const float2 w45 = make_float2(wb.x, wb.y), w67 = make_float2(wb.z, wb.w);
#pragma unroll
            for (int r = 0; r < RP; ++r) {
              const float2 xv = make_float2(xin[r + ky][kx], xin[r + ky][kx]);
              acc2[r][0] = __ffma2_rn(xv, w01, acc2[r][0]);
              acc2[r][1] = __ffma2_rn(xv, w23, acc2[r][1]);
              acc2[r][2] = __ffma2_rn(xv, w45, acc2[r][2]);
              acc2[r][3] = __ffma2_rn(xv, w67, acc2[r][3]);
            }
          }
      }
    }
    cbase += v.C;
  }

  // ---- epilogue: bias + residual + ReLU, channel-split store
  const int gx = x0 + px;
  if (gx >= a.W) return;
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = (a.bias != nullptr && co0 + j < a.Cout) ? __ldg(a.bias + co0 + j) : 0.f;
  const bool vec_res = (a.Cout % 4 == 0);
#pragma unroll
  for (int r = 0; r < RP; ++r) {
    const int gy = y0 + r0 + r;
    if (gy >= a.H) continue;
    float o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[2 * j] = acc2[r][j].x + bv[2 * j];
      o[2 * j + 1] = acc2[r][j].y + bv[2 * j + 1];
    }
    if (a.res != nullptr) {
      const float* rp = a.res + (((size_t)b * a.H + gy) * a.W + gx) * a.Cout + co0;
      if (vec_res) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (co0 + 4 * h < a.Cout) {
            const float4 q = ldg4(rp + 4 * h);
            o[4 * h + 0] += q.x; o[4 * h + 1] += q.y; o[4 * h + 2] += q.z; o[4 * h + 3] += q.w;
          }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (co0 + j < a.Cout) o[j] += __ldg(rp + j);
      }
    }
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    if (a.round_out) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = round_tf32(o[j]);
    }
    if (a.mask_out != nullptr) {  // packed mask of this op's own output (single destination, Cout % 8 == 0: checked by the API)
      unsigned m = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) m |= (o[j] > 0.f ? 1u : 0u) << j;
      a.mask_out[(((size_t)b * a.d0.Hs + (gy + a.d0.oy)) * a.d0.Ws + (gx + a.d0.ox)) * (a.d0.C / 8) + (co0 >> 3)] = (unsigned char)m;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int co = co0 + 4 * h;
      if (co >= a.Cout) continue;
      // destination select (channel split)
      const bool first = co < a.d0.C;
      const ViewW d = first ? a.d0 : a.d1;
      const int cd = first ? co : co - a.d0.C;
      const size_t doff = (((size_t)b * d.Hs + (gy + d.oy)) * d.Ws + (gx + d.ox)) * d.C + cd;
      float* dp = d.p + doff;
      if ((d.C % 4 == 0) && (a.d0.C % 4 == 0) && (co + 3 < a.Cout)) {
        const unsigned char* mk = first ? a.mask0 : a.mask1;
        if (mk != nullptr) {  // packed ReLU mask: byte (pixel, cd / 8), bits cd % 8 ..
          const unsigned m = (unsigned)__ldg(mk + (doff - cd) / 8 + (cd >> 3)) >> (cd & 7);
          o[4 * h] = (m & 1u) ? o[4 * h] : 0.f;
          o[4 * h + 1] = (m & 2u) ? o[4 * h + 1] : 0.f;
          o[4 * h + 2] = (m & 4u) ? o[4 * h + 2] : 0.f;
          o[4 * h + 3] = (m & 8u) ? o[4 * h + 3] : 0.f;
        }
        *reinterpret_cast<float4*>(dp) = make_float4(o[4 * h], o[4 * h + 1], o[4 * h + 2], o[4 * h + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = co + j;
          if (c >= a.Cout) continue;
          const bool f2 = c < a.d0.C;
          const ViewW d2 = f2 ? a.d0 : a.d1;
          const int c2 = f2 ? c : c - a.d0.C;
          const size_t off2 = (((size_t)b * d2.Hs + (gy + d2.oy)) * d2.Ws + (gx + d2.ox)) * d2.C + c2;
          const unsigned char* mk2 = f2 ? a.mask0 : a.mask1;
          float ov = o[4 * h + j];
          if (mk2 != nullptr && !((__ldg(mk2 + (off2 - c2) / 8 + (c2 >> 3)) >> (c2 & 7)) & 1u)) ov = 0.f;
          d2.p[off2] = ov;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: dw[co][ci][ky][kx] = sum_{b,y,x} g[b,y,x,co] * in[b,y+ky-1,x+kx-1,ci]
// CTA = (pixel-tile group, 8-channel ci chunk, 8-channel co chunk).  Thread = (ci, co quad, row
// partition): 36 accumulators (9 taps x 4 co), sliding 3x3 input window along x.
template <int TW, int TH>
__global__ void __launch_bounds__(256) conv3x3_wgrad_ffma_kernel(const WgradArgs a) {
  constexpr int HW_ = TW + 2;
  constexpr int HH_ = TH + 2;
  constexpr int PLANE0 = HH_ * HW_;
  constexpr int PLANE = PLANE0 + ((4 - (PLANE0 % 32)) + 32) % 32;  // PLANE % 32 == 4: 8 planes hit distinct banks
  constexpr int NPART = 16;
  constexpr int RED_LD = 8 * 2 * 36 + 1;
  constexpr int STAGE_F = 8 * PLANE + TH * TW * 8;
  constexpr int SMEM_F = STAGE_F > 8 * RED_LD ? STAGE_F : 8 * RED_LD;
  static_assert((8 * PLANE) % 4 == 0, "g_s must stay 16-byte aligned");
  static_assert(SMEM_F * 4 <= 48 * 1024, "static shared memory budget");
  __shared__ __align__(16) float smem[SMEM_F];
  float* in_s = smem;
  float* g_s = smem + 8 * PLANE;
  float (*red_s)[RED_LD] = reinterpret_cast<float (*)[RED_LD]>(smem);  // aliases the staging buffers after the main loop

  const int tid = threadIdx.x;
  const int ci = tid & 7;
  const int quad = (tid >> 3) & 1;
  const int part = tid >> 4;  // 0..15

  // which input chunk
  int cchunk = blockIdx.y;
  const int nchunk0 = (a.s0.C + 7) / 8;
  const View v = cchunk < nchunk0 ? a.s0 : a.s1;
  const int cbase = cchunk < nchunk0 ? 0 : a.s0.C;
  if (cchunk >= nchunk0) cchunk -= nchunk0;
  const int c0 = cchunk * 8;
  const int cc = min(8, v.C - c0);
  const bool vec = (v.C % 4 == 0);
  const int co0 = blockIdx.z * 8;
  const bool gvec = (a.Cout % 4 == 0);

  float2 acc2[9][2];  // FFMA2: (co, co+1) pairs
#pragma unroll
  for (int t = 0; t < 9; ++t) acc2[t][0] = acc2[t][1] = make_float2(0.f, 0.f);

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    int t = tile;
    const int tx = t % a.tilesX;
    t /= a.tilesX;
    const int ty = t % a.tilesY;
    const int b = t / a.tilesY;
    const int x0 = tx * TW, y0 = ty * TH;
    __syncthreads();
    for (int i = tid; i < PLANE0; i += 256) {
      const int hy = i / HW_, hx = i - hy * HW_;
      const int gy = y0 + hy - 1, gx = x0 + hx - 1;
      float vals[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) vals[c] = 0.f;
      if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
        const float* p = v.p + (((size_t)b * v.Hs + (gy + v.oy)) * v.Ws + (gx + v.ox)) * v.C + c0;
        if (vec) {
          const float4 q0 = ldg4(p);
          vals[0] = q0.x; vals[1] = q0.y; vals[2] = q0.z; vals[3] = q0.w;
          if (cc > 4) {
            const float4 q1 = ldg4(p + 4);
            vals[4] = q1.x; vals[5] = q1.y; vals[6] = q1.z; vals[7] = q1.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < cc) vals[c] = __ldg(p + c);
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) in_s[c * PLANE + i] = vals[c];
    }
    for (int i = tid; i < TH * TW; i += 256) {
      const int yy = i / TW, xx = i - yy * TW;
      const int gy = y0 + yy, gx = x0 + xx;
      float vals[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) vals[c] = 0.f;
      if (gy < a.H && gx < a.W) {
        const float* p = a.g + (((size_t)b * a.H + gy) * a.W + gx) * a.Cout + co0;
        if (gvec) {
          const float4 q0 = ldg4(p);
          vals[0] = q0.x; vals[1] = q0.y; vals[2] = q0.z; vals[3] = q0.w;
          if (co0 + 4 < a.Cout) {
            const float4 q1 = ldg4(p + 4);
            vals[4] = q1.x; vals[5] = q1.y; vals[6] = q1.z; vals[7] = q1.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (co0 + c < a.Cout) vals[c] = __ldg(p + c);
        }
      }
      *reinterpret_cast<float4*>(g_s + i * 8) = make_float4(vals[0], vals[1], vals[2], vals[3]);
      *reinterpret_cast<float4*>(g_s + i * 8 + 4) = make_float4(vals[4], vals[5], vals[6], vals[7]);
    }
    __syncthreads();
    // rows of this partition
    for (int yy = part; yy < TH; yy += NPART) {
      const float* ip = in_s + ci * PLANE + yy * HW_;
      float win[3][3];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        win[ky][1] = ip[ky * HW_ + 0];
        win[ky][2] = ip[ky * HW_ + 1];
      }
#pragma unroll 4
      for (int xx = 0; xx < TW; ++xx) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          win[ky][0] = win[ky][1];
          win[ky][1] = win[ky][2];
          win[ky][2] = ip[ky * HW_ + xx + 2];
        }
        const float4 gv = *reinterpret_cast<const float4*>(g_s + (yy * TW + xx) * 8 + quad * 4);
        const float2 g01 = make_float2(gv.x, gv.y), g23 = make_float2(gv.z, gv.w);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float2 xv = make_float2(win[ky][kx], win[ky][kx]);
            acc2[ky * 3 + kx][0] = __ffma2_rn(xv, g01, acc2[ky * 3 + kx][0]);
            acc2[ky * 3 + kx][1] = __ffma2_rn(xv, g23, acc2[ky * 3 + kx][1]);
          }
      }
    }
  }

  // ---- reduce the 16 row partitions (two rounds through shared memory), then one atomic per output
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    acc[t][0] = acc2[t][0].x; acc[t][1] = acc2[t][0].y; acc[t][2] = acc2[t][1].x; acc[t][3] = acc2[t][1].y;
  }
  const int slot = (ci * 2 + quad) * 36;
  __syncthreads();
  if (part >= 8) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int q = 0; q < 4; ++q) red_s[part - 8][slot + t * 4 + q] = acc[t][q];
  }
  __syncthreads();
  if (part < 8) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[t][q] += red_s[part][slot + t * 4 + q];
  }
  __syncthreads();
  if (part < 8) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int q = 0; q < 4; ++q) red_s[part][slot + t * 4 + q] = acc[t][q];
  }
  __syncthreads();
  for (int i = tid; i < 8 * 2 * 36; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) s += red_s[p][i];
    const int q = i & 3, tap = (i >> 2) % 9, cq = i / 36;
    const int ci_ = cq >> 1, quad_ = cq & 1;
    const int co = co0 + quad_ * 4 + q;
    if (ci_ < cc && co < a.Cout) atomicAdd(a.dw + ((size_t)co * a.Cin + cbase + c0 + ci_) * 9 + tap, s);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad, pipelined version for channel counts that are multiples of 8: both operand tiles stay PIXEL-major in
// shared memory ([pixel][8 channels], exactly as they sit in the NHWC tensors), so they are staged with plain
// 16-byte cp.async (zero-fill outside the image) into a two-stage ring: tile n+1 streams in while tile n is
// reduced.  Thread = (ci, co quad, row partition) as above; the 8 ci lanes of a pixel read 8 consecutive words
// and the two partitions of a warp sit (TW+2)*8 words apart (16 banks for TW = 32, 16), i.e. conflict-free.
template <int TW, int TH>
__global__ void __launch_bounds__(256, 2) conv3x3_wgrad_pipe_kernel(const WgradArgs a) {
  constexpr int HW_ = TW + 2, HH_ = TH + 2;
  constexpr int XPIX = HH_ * HW_, GPIX = TH * TW;
  constexpr int STAGE_F = (XPIX + GPIX) * 8;
  constexpr int HALF_W = TW / 2;
  constexpr int RED_LD = 8 * 72 + 1;  // 8 ci x (9 taps x 8 co); 256 threads = 8 ci x 32 (row, column half) partitions
  extern __shared__ __align__(16) float dsm[];  // [2][STAGE_F]; the reduction buffer aliases it at the end
  static_assert(TH == 16, "one partition per (row, half)");
  static_assert(2 * STAGE_F >= 16 * RED_LD, "reduction buffer must fit in the staging ring");

  const int tid = threadIdx.x;
  const int ci = tid & 7, part = tid >> 3;  // thread = one input channel x all 8 output channels x 9 taps
  const int prow = part & 15, phalf = part >> 4;
  int cchunk = blockIdx.y;
  const int nchunk0 = a.s0.C / 8;
  const View v = cchunk < nchunk0 ? a.s0 : a.s1;
  const int cbase = cchunk < nchunk0 ? 0 : a.s0.C;
  if (cchunk >= nchunk0) cchunk -= nchunk0;
  const int c0 = cchunk * 8;
  const int co0 = blockIdx.z * 8;

  auto issue = [&](int tile, int stage) {
    int t = tile;
    const int tx = t % a.tilesX;
    t /= a.tilesX;
    const int ty = t % a.tilesY;
    const int b = t / a.tilesY;
    const int x0 = tx * TW, y0 = ty * TH;
    float* xs = dsm + stage * STAGE_F;
    float* gs = xs + XPIX * 8;
    for (int i = tid; i < XPIX * 2; i += 256) {  // two 16-byte halves per halo pixel
      const int pix = i >> 1, half = i & 1;
      const int hy = pix / HW_, hx = pix - hy * HW_;
      const int gy = y0 + hy - 1, gx = x0 + hx - 1;
      const bool ok = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
      const float* src = ok ? v.p + (((size_t)b * v.Hs + (gy + v.oy)) * v.Ws + (gx + v.ox)) * v.C + c0 + half * 4 : v.p;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(xs + pix * 8 + half * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    for (int i = tid; i < GPIX * 2; i += 256) {
      const int pix = i >> 1, half = i & 1;
      const int yy = pix / TW, xx = pix - yy * TW;
      const int gy = y0 + yy, gx = x0 + xx;
      const bool ok = gy < a.H && gx < a.W;
      const float* src = ok ? a.g + (((size_t)b * a.H + gy) * a.W + gx) * a.Cout + co0 + half * 4 : a.g;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(gs + pix * 8 + half * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float2 acc2[9][4];  // FFMA2 accumulators: 9 taps x (co pairs 01,23,45,67)
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc2[t][q] = make_float2(0.f, 0.f);

  int stage = 0;
  if ((int)blockIdx.x < a.ntiles) issue(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, stage ^= 1) {
    const int next = tile + gridDim.x;
    if (next < a.ntiles) {
      issue(next, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* xs = dsm + stage * STAGE_F;
    const float* gs = xs + XPIX * 8;
    {
      const int xbeg = phalf * HALF_W;
      const float* ip = xs + (prow * HW_ + xbeg) * 8 + ci;
      const float* gp = gs + (prow * TW + xbeg) * 8;
      float win[3][3];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        win[ky][1] = ip[(ky * HW_ + 0) * 8];
        win[ky][2] = ip[(ky * HW_ + 1) * 8];
      }
#pragma unroll 4
      for (int xx = 0; xx < HALF_W; ++xx) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          win[ky][0] = win[ky][1];
          win[ky][1] = win[ky][2];
          win[ky][2] = ip[(ky * HW_ + xx + 2) * 8];
        }
        const float4 ga = *reinterpret_cast<const float4*>(gp + xx * 8);
        const float4 gb = *reinterpret_cast<const float4*>(gp + xx * 8 + 4);
        const float2 g01 = make_float2(ga.x, ga.y), g23 = make_float2(ga.z, ga.w);
        const float2 g45 = make_float2(gb.x, gb.y), g67 = make_float2(gb.z, gb.w);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float2 xv = make_float2(win[ky][kx], win[ky][kx]);
            acc2[ky * 3 + kx][0] = __ffma2_rn(xv, g01, acc2[ky * 3 + kx][0]);
            acc2[ky * 3 + kx][1] = __ffma2_rn(xv, g23, acc2[ky * 3 + kx][1]);
            acc2[ky * 3 + kx][2] = __ffma2_rn(xv, g45, acc2[ky * 3 + kx][2]);
            acc2[ky * 3 + kx][3] = __ffma2_rn(xv, g67, acc2[ky * 3 + kx][3]);
          }
      }
    }
    __syncthreads();  // everyone is done with `stage` before the next iteration refills it
  }

  // ---- reduce the 32 partitions through shared memory (16 slots, two folding rounds), then one atomic per output
  float (*red_s)[RED_LD] = reinterpret_cast<float (*)[RED_LD]>(dsm);
  const int slot = ci * 72;
  __syncthreads();
  if (part >= 16) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        red_s[part - 16][slot + t * 8 + 2 * q] = acc2[t][q].x;
        red_s[part - 16][slot + t * 8 + 2 * q + 1] = acc2[t][q].y;
      }
  }
  __syncthreads();
  if (part < 16) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc2[t][q].x += red_s[part][slot + t * 8 + 2 * q];
        acc2[t][q].y += red_s[part][slot + t * 8 + 2 * q + 1];
      }
  }
  __syncthreads();
  if (part < 16) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        red_s[part][slot + t * 8 + 2 * q] = acc2[t][q].x;
        red_s[part][slot + t * 8 + 2 * q + 1] = acc2[t][q].y;
      }
  }
  __syncthreads();
  for (int i = tid; i < 8 * 72; i += 256) {
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < 16; ++p) sum += red_s[p][i];
    const int co = co0 + (i & 7), tap = (i >> 3) % 9, ci_ = i / 72;
    atomicAdd(a.dw + ((size_t)co * a.Cin + cbase + c0 + ci_) * 9 + tap, sum);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad in TF32 mode: the contraction dimension is the PIXEL index, the outputs are tiny (8 ci x 8 co x 9 taps per
// CTA), which does not fit tcgen05 (M >= 64; K-major operands would need pixel-contiguous = transposed tiles, and
// MN-major tf32 with SWIZZLE_NONE returns zeros on B200).  The warp-level mma.sync m16n8k8 TF32 path fits exactly:
// A tile = (2 taps x 8 ci) x 8 pixels, B tile = 8 pixels x 8 co, both read conflict-free from the same pixel-major
// [pixel][8] shared tiles the cp.async ring of conv3x3_wgrad_pipe_kernel provides (bank = 8*t + g).  Each warp owns
// a share of the 8-pixel groups of a tile; 5 MMAs (tap pairs) per group; partial sums reduced across the 8 warps
// at the end, one fp32 atomic per output and CTA.  Operands are already TF32-rounded by their producers.
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NCO = 8-channel co tiles per CTA (1, 2 or 4): the A fragments of an 8-pixel group (18 LDS: nine shifted windows of the
// x halo tile) are reused for NCO x 5 MMAs, and the x tile is re-read Cout/(8*NCO) times instead of Cout/8 times
// (with NCO = 1 the wide layers were L2-bandwidth bound: 270 MB of L2 reads for a 42 MB problem).
template <int TW, int TH, int NCO>
__global__ void __launch_bounds__(256, 2) conv3x3_wgrad_mma_kernel(const WgradArgs a) {
  constexpr int HW_ = TW + 2, HH_ = TH + 2;
  constexpr int XPIX = HH_ * HW_, GPIX = TH * TW;
  constexpr int GP = NCO == 1 ? 8 : 8 * NCO + 8;  // g-tile pixel pitch: == 8 (mod 32) or 8/24 -> B-fragment reads conflict-free
  constexpr int STAGE_F = XPIX * 8 + GPIX * GP;
  constexpr int GROUPS = TH * (TW / 8);  // 8-pixel groups (along x) per tile
  constexpr int NACC = 20 * NCO;
  extern __shared__ __align__(16) float dsm[];  // [2][STAGE_F]; the reduction buffer aliases it at the end
  static_assert(2 * STAGE_F >= 8 * 32 * NACC, "reduction buffer must fit in the staging ring");

  pdl_prologue();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;  // mma fragment coordinates: groupID, threadID_in_group
  int cchunk = blockIdx.y;
  const int nchunk0 = a.s0.C / 8;
  const View v = cchunk < nchunk0 ? a.s0 : a.s1;
  const int cbase = cchunk < nchunk0 ? 0 : a.s0.C;
  if (cchunk >= nchunk0) cchunk -= nchunk0;
  const int c0 = cchunk * 8;
  const int co0 = blockIdx.z * 8 * NCO;

  auto issue = [&](int tile, int stage) {
    int t = tile;
    const int tx = t % a.tilesX;
    t /= a.tilesX;
    const int ty = t % a.tilesY;
    const int b = t / a.tilesY;
    const int x0 = tx * TW, y0 = ty * TH;
    float* xs = dsm + stage * STAGE_F;
    float* gs = xs + XPIX * 8;
    for (int i = tid; i < XPIX * 2; i += 256) {
      const int pix = i >> 1, half = i & 1;
      const int hy = pix / HW_, hx = pix - hy * HW_;
      const int gy = y0 + hy - 1, gx = x0 + hx - 1;
      const bool ok = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
      const float* src = ok ? v.p + (((size_t)b * v.Hs + (gy + v.oy)) * v.Ws + (gx + v.ox)) * v.C + c0 + half * 4 : v.p;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(xs + pix * 8 + half * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    for (int i = tid; i < GPIX * 2 * NCO; i += 256) {
      const int pix = i / (2 * NCO), q = i - pix * (2 * NCO);  // q: 16-byte unit inside the pixel's 8*NCO channels
      const int yy = pix / TW, xx = pix - yy * TW;
      const int gy = y0 + yy, gx = x0 + xx;
      const bool ok = gy < a.H && gx < a.W;
      const float* src = ok ? a.g + (((size_t)b * a.H + gy) * a.W + gx) * a.Cout + co0 + q * 4 : a.g;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(gs + pix * GP + q * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // accumulators of the 5 tap-pair tiles per co tile: acc[n][p] = {(tap 2p, ci=gq, co=8n+2tq), (tap 2p, gq, +1), (tap 2p+1, ...), ...}
  float acc[NCO][5][4];
#pragma unroll
  for (int n = 0; n < NCO; ++n)
#pragma unroll
    for (int p = 0; p < 5; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[n][p][j] = 0.f;

  int stage = 0;
  if ((int)blockIdx.x < a.ntiles) issue(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, stage ^= 1) {
    const int next = tile + gridDim.x;
    if (next < a.ntiles) {
      issue(next, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const unsigned* xs = reinterpret_cast<const unsigned*>(dsm + stage * STAGE_F);
    const unsigned* gs = xs + XPIX * 8;
    for (int grp = warp; grp < GROUPS && !(a.debug & 4); grp += 8) {
      const int yy = grp / (TW / 8), xg = (grp - yy * (TW / 8)) * 8;
      // A fragments: row = (tap within pair, ci = gq), col = pixel (tq, tq+4); the halo tile is offset by (+1,+1)
      const unsigned* xr = xs + ((yy * HW_) + xg + tq) * 8 + gq;
      unsigned av[9][2];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          av[ky * 3 + kx][0] = xr[(ky * HW_ + kx) * 8];
          av[ky * 3 + kx][1] = xr[(ky * HW_ + kx + 4) * 8];
        }
#pragma unroll
      for (int n = 0; n < NCO; ++n) {
        // B fragment: k = pixel (tq, tq+4), n = co (gq)
        const unsigned b0 = gs[(yy * TW + xg + tq) * GP + 8 * n + gq];
        const unsigned b1 = gs[(yy * TW + xg + tq + 4) * GP + 8 * n + gq];
        if (a.debug & 2) {  // (experiment) fragment loads without MMAs
#pragma unroll
          for (int p = 0; p < 9; ++p) acc[n][p >> 1][p & 3] += __uint_as_float(av[p][0] ^ av[p][1] ^ b0 ^ b1);
          continue;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) mma_tf32_16x8x8(acc[n][p], av[2 * p][0], av[2 * p + 1][0], av[2 * p][1], av[2 * p + 1][1], b0, b1);
        // rows 8..15 of the fifth tile are free: feeding ones there makes them the column sums of G = the bias gradient
        mma_tf32_16x8x8(acc[n][4], av[8][0], 0x3f800000u, av[8][1], 0x3f800000u, b0, b1);
      }
    }
    __syncthreads();
  }

  // ---- reduce the 8 warps, then one atomic per output
  float* red = dsm;  // [8 warps][32 lanes][NACC]
  __syncthreads();
#pragma unroll
  for (int n = 0; n < NCO; ++n)
#pragma unroll
    for (int p = 0; p < 5; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(warp * 32 + lane) * NACC + n * 20 + p * 4 + j] = acc[n][p][j];
  __syncthreads();
  for (int i = tid; i < 32 * NACC; i += 256) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w * 32 * NACC + i];
    const int ln = i / NACC, r = i - ln * NACC;
    const int n = r / 20, r2 = r - n * 20;
    const int p = r2 >> 2, j = r2 & 3;
    const int tap = 2 * p + (j >> 1);
    const int ci_ = ln >> 2, co = co0 + 8 * n + 2 * (ln & 3) + (j & 1);
    if ((a.debug & 1) && sum != 12345.678f) continue;  // (experiment) no atomics
    if (tap > 8) {  // the ones rows: every ci_ row holds the same sum_pixels g[.][co]
      if (a.db != nullptr && ci_ == 0 && blockIdx.y == 0) atomicAdd(a.db + co, sum);
      continue;
    }
    atomicAdd(a.dw + ((size_t)co * a.Cin + cbase + c0 + ci_) * 9 + tap, sum);
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void pack_w3x3_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int transpose) {
  const int n = Cout * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int tap = i % 9;
    const int ci = (i / 9) % Cin;
    const int co = i / (9 * Cin);
    const float v = w[i];
    if (!transpose) {
      out[((size_t)tap * Cin + ci) * Cout + co] = v;  // [9][Cin][Cout]
    } else {
      out[((size_t)(8 - tap) * Cout + co) * Cin + ci] = v;  // [9 flipped][Cout][Cin]
    }
  }
}

__global__ void relu_bwd_bias_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ g,
                                     float* __restrict__ dbias, long long npix, int C, int flags) {
  const int relu = flags & PU_FLAG_RELU, rnd = flags & PU_FLAG_ROUND_TF32;
  // generic scalar version: thread -> channel c = idx % C over a strided set of pixels
  extern __shared__ float red[];
  const int c4n = C;  // scalar granularity
  const long long total = npix * c4n;
  const long long stride = (long long)gridDim.x * blockDim.x;  // multiple of C by construction
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    float v = dy[i];
    if (relu) v = y[i] > 0.f ? v : 0.f;
    s += v;
    if (rnd) v = round_tf32(v);
    if (g != nullptr) g[i] = v;
  }
  if (dbias == nullptr) return;
  red[threadIdx.x] = s;
  __syncthreads();
  // threads with equal (threadIdx.x % C) own the same channel
  if ((int)threadIdx.x < C) {
    float t = 0.f;
    for (int j = threadIdx.x; j < (int)blockDim.x; j += C) t += red[j];
    atomicAdd(dbias + threadIdx.x, t);
  }
}

__global__ void relu_bwd_bias_vec4_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, float4* __restrict__ g,
                                          float* __restrict__ dbias, long long n4, int C4, int flags) {
  const int relu = flags & PU_FLAG_RELU, rnd = flags & PU_FLAG_ROUND_TF32;
  // blockDim.x * gridDim.x is a multiple of C4 => each thread stays on one channel quad
  __shared__ float4 red[256];
  const long long stride = (long long)gridDim.x * blockDim.x;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = dy[i];
    if (relu) {
      const float4 m = y[i];
      v.x = m.x > 0.f ? v.x : 0.f;
      v.y = m.y > 0.f ? v.y : 0.f;
      v.z = m.z > 0.f ? v.z : 0.f;
      v.w = m.w > 0.f ? v.w : 0.f;
    }
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    if (rnd) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    if (g != nullptr) g[i] = v;
  }
  if (dbias == nullptr) return;
  red[threadIdx.x] = s;
  __syncthreads();
  if ((int)threadIdx.x < C4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = threadIdx.x; j < (int)blockDim.x; j += C4) {
      const float4 r = red[j];
      t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
    }
    atomicAdd(dbias + threadIdx.x * 4 + 0, t.x);
    atomicAdd(dbias + threadIdx.x * 4 + 1, t.y);
    atomicAdd(dbias + threadIdx.x * 4 + 2, t.z);
    atomicAdd(dbias + threadIdx.x * 4 + 3, t.w);
  }
}

// host-side launchers ---------------------------------------------------------------------------
int conv3x3_fwd_ffma(const Conv3x3Args& a0, cudaStream_t st) {
  if (conv3x3_c1_ok(a0.Cin, a0.Cout) && a0.s1.p == nullptr && a0.wfmt == 1 && a0.res == nullptr && a0.d1.p == nullptr &&
      a0.mask0 == nullptr && a0.d0.C == a0.Cout)
    return conv3x3_c1_fwd(a0, st);  // streaming stem kernel
  Conv3x3Args a = a0;
  const int cog = cdiv(a.Cout, 8);
  if (a.W > 16) {
    a.tilesX = cdiv(a.W, 32); a.tilesY = cdiv(a.H, 32);
    dim3 grid(a.tilesX * a.tilesY * a.B, cog);
    conv3x3_ffma_kernel<32, 4, 256><<<grid, 256, 0, st>>>(a);
  } else if (a.W > 8) {
    a.tilesX = cdiv(a.W, 16); a.tilesY = cdiv(a.H, 16);
    dim3 grid(a.tilesX * a.tilesY * a.B, cog);
    conv3x3_ffma_kernel<16, 2, 128><<<grid, 128, 0, st>>>(a);
  } else {
    a.tilesX = cdiv(a.W, 8); a.tilesY = cdiv(a.H, 8);
    dim3 grid(a.tilesX * a.tilesY * a.B, cog);
    conv3x3_ffma_kernel<8, 1, 64><<<grid, 64, 0, st>>>(a);
  }
  return post_launch("conv3x3_fwd_ffma");
}

int conv3x3_wgrad_ffma(const WgradArgs& a0, cudaStream_t st, int math) {
  if (conv3x3_c1_ok(a0.Cin, a0.Cout) && (a0.s1.p == nullptr || a0.s1.C == 0)) return conv3x3_c1_wgrad(a0, st, math);  // stem kernels (dw + db)
  WgradArgs a = a0;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e_ = getenv("PU_WG_DEBUG");
      dbg = e_ ? atoi(e_) : 0;
    }
    a.debug = dbg;
  }
  const int nci = cdiv(a.s0.C, 8) + ((a.s1.p != nullptr && a.s1.C > 0) ? cdiv(a.s1.C, 8) : 0);
  const int nco = cdiv(a.Cout, 8);
  const bool have1 = a.s1.p != nullptr && a.s1.C > 0;
  const bool mma_path = math == PU_MATH_TF32 && a.s0.C % 8 == 0 && (!have1 || a.s1.C % 8 == 0) && a.Cout % 8 == 0;
  if (a.accum && !mma_path) {
    set_error("pu_conv3x3_wgrad: PU_MATH_ACCUM needs the TF32 path (channel counts that are multiples of 8) or the one-channel stem");
    return PU_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaSuccess;
  if (!a.accum) {
    e = cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.Cout * a.Cin * 9, st);
    if (e != cudaSuccess) {
      set_error("conv3x3_wgrad memset: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
  }
  if (a.db != nullptr) {
    if (mma_path) {  // accumulated by the ones rows of the MMA kernel
      if (!a.accum) {
        e = cudaMemsetAsync(a.db, 0, sizeof(float) * a.Cout, st);
        if (e != cudaSuccess) {
          set_error("conv3x3_wgrad db memset: %s", cudaGetErrorString(e));
          return PU_ERR_CUDA;
        }
      }
    } else {  // separate read-only reduction pass over g
      int rc = pu_relu_bwd_bias(a.g, nullptr, nullptr, a.db, (long long)a.B * a.H * a.W, a.Cout, 0, st);
      if (rc) return rc;
    }
  }
  if (a.s0.C % 8 == 0 && (!have1 || a.s1.C % 8 == 0) && a.Cout % 8 == 0) {
    if (math == PU_MATH_TF32) {
      // TMA-fed kernel (conv3x3_wgrad_tma.cu); PU_WGRAD_V=1 selects the first, cp.async-fed version below (A/B measurements)
      static int ver = -1;
      if (ver < 0) {
        const char* e_ = getenv("PU_WGRAD_V");
        ver = e_ ? atoi(e_) : 2;
      }
      if (ver != 1 && conv3x3_wgrad_tma_ok(a)) return conv3x3_wgrad_tma(a, st);
      // warp-level TF32 MMAs over the same cp.async ring; NCO co tiles (8 channels each) per CTA
      const int nco_t = a.Cout % 32 == 0 ? 4 : (a.Cout % 16 == 0 ? 2 : 1);
      const int ncoz = a.Cout / (8 * nco_t);
      auto stage_bytes = [](int tw, int th, int n) { return (size_t)2 * ((th + 2) * (tw + 2) * 8 + th * tw * (n == 1 ? 8 : 8 * n + 8)) * sizeof(float); };
      static bool attr = false;
      if (!attr) {
        cudaFuncSetAttribute(conv3x3_wgrad_mma_kernel<32, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes(32, 16, 1));
        cudaFuncSetAttribute(conv3x3_wgrad_mma_kernel<16, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes(16, 16, 1));
        cudaFuncSetAttribute(conv3x3_wgrad_mma_kernel<16, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes(16, 16, 2));
        cudaFuncSetAttribute(conv3x3_wgrad_mma_kernel<16, 16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes(16, 16, 4));
        attr = true;
      }
      if (a.W > 16 && nco_t == 1) {
        a.tilesX = cdiv(a.W, 32); a.tilesY = cdiv(a.H, 16);
        a.ntiles = a.tilesX * a.tilesY * a.B;
        const int gx = max(1, min(a.ntiles, (3 * kNumSMs) / max(1, nci * ncoz)));  // 72 KB + 72 regs per CTA: three fit an SM
        launch_pdl(conv3x3_wgrad_mma_kernel<32, 16, 1>, dim3(gx, nci, ncoz), dim3(256), stage_bytes(32, 16, 1), st, a);
      } else {
        a.tilesX = cdiv(a.W, 16); a.tilesY = cdiv(a.H, 16);
        a.ntiles = a.tilesX * a.tilesY * a.B;
        const int per_sm = nco_t == 4 ? 2 : 4;
        const int gx = max(1, min(a.ntiles, (per_sm * kNumSMs) / max(1, nci * ncoz)));
        const dim3 grid(gx, nci, ncoz);
        if (nco_t == 4) launch_pdl(conv3x3_wgrad_mma_kernel<16, 16, 4>, grid, dim3(256), stage_bytes(16, 16, 4), st, a);
        else if (nco_t == 2) launch_pdl(conv3x3_wgrad_mma_kernel<16, 16, 2>, grid, dim3(256), stage_bytes(16, 16, 2), st, a);
        else launch_pdl(conv3x3_wgrad_mma_kernel<16, 16, 1>, grid, dim3(256), stage_bytes(16, 16, 1), st, a);
      }
      return post_launch("conv3x3_wgrad_mma");
    }
    // pipelined cp.async kernel (2 CTAs / SM, 72 accumulators per thread)
    if (a.W > 16) {
      constexpr int TW = 32, TH = 16;
      a.tilesX = cdiv(a.W, TW); a.tilesY = cdiv(a.H, TH);
      a.ntiles = a.tilesX * a.tilesY * a.B;
      const size_t smem = 2 * ((TH + 2) * (TW + 2) + TH * TW) * 8 * sizeof(float);
      static bool attr = false;
      if (!attr) {
        cudaFuncSetAttribute(conv3x3_wgrad_pipe_kernel<TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
      }
      const int gx = max(1, min(a.ntiles, (2 * kNumSMs) / max(1, nci * nco)));
      dim3 grid(gx, nci, nco);
      conv3x3_wgrad_pipe_kernel<TW, TH><<<grid, 256, smem, st>>>(a);
    } else {
      constexpr int TW = 16, TH = 16;
      a.tilesX = cdiv(a.W, TW); a.tilesY = cdiv(a.H, TH);
      a.ntiles = a.tilesX * a.tilesY * a.B;
      const size_t smem = 2 * ((TH + 2) * (TW + 2) + TH * TW) * 8 * sizeof(float);
      static bool attr = false;
      if (!attr) {
        cudaFuncSetAttribute(conv3x3_wgrad_pipe_kernel<TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
      }
      const int gx = max(1, min(a.ntiles, (4 * kNumSMs) / max(1, nci * nco)));
      dim3 grid(gx, nci, nco);
      conv3x3_wgrad_pipe_kernel<TW, TH><<<grid, 256, smem, st>>>(a);
    }
    return post_launch("conv3x3_wgrad_pipe");
  }
  if (a.W > 16) {
    a.tilesX = cdiv(a.W, 32); a.tilesY = cdiv(a.H, 16);
    a.ntiles = a.tilesX * a.tilesY * a.B;
    const int gx = max(1, min(a.ntiles, (4 * kNumSMs) / max(1, nci * nco)));
    dim3 grid(gx, nci, nco);
    conv3x3_wgrad_ffma_kernel<32, 16><<<grid, 256, 0, st>>>(a);
  } else {
    a.tilesX = cdiv(a.W, 16); a.tilesY = cdiv(a.H, 16);
    a.ntiles = a.tilesX * a.tilesY * a.B;
    const int gx = max(1, min(a.ntiles, (8 * kNumSMs) / max(1, nci * nco)));
    dim3 grid(gx, nci, nco);
    conv3x3_wgrad_ffma_kernel<16, 16><<<grid, 256, 0, st>>>(a);
  }
  return post_launch("conv3x3_wgrad_ffma");
}

}  // namespace pu

extern "C" {

long long pu_pack_w3x3_floats(int Cout, int Cin, int transpose, int math, int C0) {
  if (math == PU_MATH_TF32 || math == PU_MATH_TF32_FLAT) {
    const bool flat = math == PU_MATH_TF32_FLAT;
    const long long n = transpose ? pu::conv3x3_tc_weight_floats(Cout, 0, Cin, flat) : pu::conv3x3_tc_weight_floats(C0, Cin - C0, Cout, flat);
    if (n > 0) return n;
  }
  return 9LL * Cin * Cout;
}

int pu_pack_w3x3(const float* w, float* out, int Cout, int Cin, int transpose, int math, int C0, void* stream) {
  PU_REQUIRE(w && out && Cout > 0 && Cin > 0, PU_ERR_BAD_ARG, "pu_pack_w3x3: bad argument");
  PU_REQUIRE(math == PU_MATH_FP32 || math == PU_MATH_TF32 || math == PU_MATH_TF32_FLAT, PU_ERR_BAD_ARG, "pu_pack_w3x3: unknown math mode %d", math);
  if (math != PU_MATH_FP32) return pu::conv3x3_tc_pack(w, out, Cout, Cin, transpose, C0, math == PU_MATH_TF32_FLAT, pu::as_stream(stream));
  const int n = Cout * Cin * 9;
  pu::pack_w3x3_kernel<<<pu::cdiv(n, 256) > 592 ? 592 : pu::cdiv(n, 256), 256, 0, pu::as_stream(stream)>>>(w, out, Cout, Cin, transpose);
  return pu::post_launch("pu_pack_w3x3");
}

int pu_relu_bwd_bias(const float* dy, const float* y, float* g, float* dbias, long long npix, int C, int flags, void* stream) {
  const int relu = flags;  // forwarded as the flag word
  PU_REQUIRE(dy && npix > 0 && C > 0 && (y || !(flags & PU_FLAG_RELU)), PU_ERR_BAD_ARG, "pu_relu_bwd_bias: bad argument");
  PU_REQUIRE(g || dbias, PU_ERR_BAD_ARG, "pu_relu_bwd_bias: nothing to compute");
  cudaStream_t st = pu::as_stream(stream);
  if (dbias) {
    cudaError_t e = cudaMemsetAsync(dbias, 0, sizeof(float) * C, st);
    if (e != cudaSuccess) {
      pu::set_error("pu_relu_bwd_bias memset: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
  }
  const bool pow2 = (C & (C - 1)) == 0;
  if (C % 4 == 0 && pow2 && C / 4 <= 256 && pu::aligned16(dy) && (!y || pu::aligned16(y)) && (!g || pu::aligned16(g))) {
    const long long n4 = npix * (C / 4);
    int grid = (int)((n4 + 256 * 8 - 1) / (256 * 8));
    grid = grid < 1 ? 1 : (grid > 8 * pu::kNumSMs ? 8 * pu::kNumSMs : grid);
    pu::relu_bwd_bias_vec4_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y),
                                                        reinterpret_cast<float4*>(g), dbias, n4, C / 4, relu);
  } else {
    // block size = multiple of C so that a thread keeps its channel
    PU_REQUIRE(C <= 1024, PU_ERR_UNSUPPORTED, "pu_relu_bwd_bias: C=%d > 1024 with ragged channel count", C);
    int bs = (256 / C) * C;
    if (bs == 0) bs = C;
    const long long total = npix * C;
    int grid = (int)((total + (long long)bs * 8 - 1) / ((long long)bs * 8));
    grid = grid < 1 ? 1 : (grid > 8 * pu::kNumSMs ? 8 * pu::kNumSMs : grid);
    pu::relu_bwd_bias_kernel<<<grid, bs, bs * sizeof(float), st>>>(dy, y, g, dbias, npix, C, relu);
  }
  return pu::post_launch("pu_relu_bwd_bias");
}

}  // extern "C"
