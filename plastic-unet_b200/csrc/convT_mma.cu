// convT_mma.cu — ConvTranspose2d(k=2, s=2) (reference unet_p.py:155) on the warp-level TF32 tensor cores.
//
// A 2x2 stride-2 transposed convolution is three skinny GEMMs over the P = B*H*W input pixels:
//     forward   Y'[P, 4*Cout] = X[P, Cin] . Wm[Cin, (a,c,co)]                + pixel-shuffle store
//     dgrad     dX[P, Cin]    = sum_a dY_a[P, (c,co)] . Wm_a^T[(c,co), Cin]   (dY_a = output row 2h+a: 2*Cout contiguous floats)
//     wgrad     dWm[Cin, 4*Cout] = X^T[Cin, P] . dY'[P, 4*Cout],  db = column sums of dY'
// They move 40+ MB per layer at the top of the decoder and only 67 MFMA, so they are memory-bound; the CUDA-core
// versions in convT.cu were LDS- and index-arithmetic-bound instead (19-75 us per launch, 22% of a training step).
// Here pixel tiles stream through a 2-stage cp.async ring (rows are contiguous in NHWC memory), the small weight
// matrix sits in shared memory in B-fragment order, and mma.sync.m16n8k8 (tf32 in, fp32 accumulate) does the math;
// fragment reads are bank-conflict free by choice of the row pitch.  tcgen05 does not fit: M or N would be 8..64
// with K = 8..64, the operands change every 64 pixels, and the outputs need per-row scatter.
//
// Used when PU_FLAG_TF32_MATH is set (the TF32 model mode); operands are rounded to TF32 (RN) when fragments are
// built, accumulation is fp32.  Other shapes and the fp32 mode keep the convT.cu kernels.
#include "pu_common.cuh"

namespace pu {

__device__ __forceinline__ void ct_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32b(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void cp16(float* dst, const float* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kCtTile = 64;  // pixels per tile
constexpr int kCtStages = 2; // cp.async ring depth (2, 3 and 4 measured within 2% of each other; deeper rings cost CTAs per SM)

struct CtArgs {
  const float* x;      // [B,H,W,Cin]
  const float* w;      // [Cin][Cout][2][2]
  const float* bias;   // [Cout] | null
  const float* dy;     // [B,2H,2W,Cout]
  float* y;            // forward output
  float* dx;
  float* dw;
  float* db;
  int B, H, W, Cin, Cout, round_out, mask_in;
  int wsh, hsh;        // log2(W), log2(H) when they are powers of two, else -1
  long long npix;      // < 2^31 (checked on the host)
};

// pixel index -> (image, row, column) with 32-bit arithmetic; shifts when the sizes are powers of two
__device__ __forceinline__ void ct_split(const CtArgs& a, int p, int& b, int& hy, int& wx) {
  int q;
  if (a.wsh >= 0) { q = p >> a.wsh; wx = p & (a.W - 1); } else { q = p / a.W; wx = p - q * a.W; }
  if (a.hsh >= 0) { b = q >> a.hsh; hy = q & (a.H - 1); } else { b = q / a.H; hy = q - b * a.H; }
}

__host__ __device__ inline int ct_pitch_b(int n) { return (n + 31) / 32 * 32 + 8; }  // == 8 (mod 32): B-fragment reads conflict-free

// ---- forward (MODE 0) and dgrad (MODE 1): pixel-major A tile x resident weight matrix -------------------------------
// 128 threads: warp w owns pixels 16w..16w+15 of the tile.  KS = K/8 k-steps per pass (A fragments live in registers).
// gridDim.y splits the N columns (forward: one (a,c) quadrant per CTA; dgrad: NT n-tiles of Cin per CTA) so that the
// per-CTA weight matrix stays small for the wide layers.
template <int MODE, int KS, int NT>
__global__ void __launch_bounds__(128) convT2x2_px_mma_kernel(const CtArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int K = 8 * KS;                                  // forward: Cin; dgrad: 2*Cout per output-row parity
  constexpr int PA = K + 4;                                  // == 4 (mod 8): A-fragment reads conflict-free
  const int N = (MODE == 0 ? 4 * a.Cout : a.Cin) / (int)gridDim.y;  // columns of this CTA
  const int n_base = blockIdx.y * N;
  constexpr int NPASS = MODE == 0 ? 1 : 2;
  const int pB = ct_pitch_b(N);
  float* Wm = sm;                                            // [NPASS][K][pB], tf32
  float* At = sm + NPASS * K * pB;                           // [kCtStages][64][PA]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  // weights -> shared, linear (coalesced) read of [Cin][Cout][a][c]
  if (MODE == 0 && gridDim.y == 4) {  // one quadrant: Wm[ci][co] = w[ci][co][blockIdx.y]
    const int total = a.Cin * a.Cout;
    for (int i = tid; i < total; i += 128) {
      const int ci = i / a.Cout, co = i - ci * a.Cout;
      Wm[ci * pB + co] = round_tf32(__ldg(a.w + 4 * (size_t)i + blockIdx.y));
    }
  } else {  // forward, all quadrants: Wm[ci][(a,c,co)]; dgrad: Wm[a][(c,co)][ci - n_base] for this CTA's contiguous ci range
    const int n4 = 4 * a.Cout;
    const int total = (MODE == 0 ? a.Cin : N) * n4;
    const float* wsrc = a.w + (MODE == 0 ? 0 : (size_t)n_base * n4);
    for (int i = tid; i < total; i += 128) {
      const int ci = i / n4, r = i - ci * n4, co = r >> 2, ac = r & 3;
      const float v = round_tf32(__ldg(wsrc + i));
      if (MODE == 0) Wm[ci * pB + ac * a.Cout + co] = v;
      else Wm[((ac >> 1) * K + (ac & 1) * a.Cout + co) * pB + ci] = v;
    }
  }
  const long long ntiles = (a.npix + kCtTile - 1) / kCtTile;
  // this CTA's work list: unit v = (its (v / NPASS)-th tile, pass v % NPASS); both passes of a tile stay in one CTA
  const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long units = my_tiles * NPASS;
  // staging: two threads per pixel row
  const int srow = tid >> 1, shalf = tid & 1;
  auto issue = [&](long long v, int buf) {
    const long long tile = blockIdx.x + (v / NPASS) * gridDim.x;
    const int pass = (int)(v % NPASS);
    const long long p = tile * kCtTile + srow;
    const bool ok = p < a.npix;
    const float* src;
    if (MODE == 0) {
      src = a.x + (ok ? p : 0) * K;
    } else {
      int b, hy, wx;
      ct_split(a, ok ? (int)p : 0, b, hy, wx);
      src = a.dy + (((size_t)b * 2 * a.H + 2 * hy + pass) * 2 * a.W + 2 * wx) * a.Cout;
    }
    float* dst = At + ((size_t)buf * kCtTile + srow) * PA;
#pragma unroll
    for (int cu = 0; cu < K / 8; ++cu) cp16(dst + 4 * (2 * cu + shalf), src + 4 * (2 * cu + shalf), ok);
    cp_commit();
  };

  for (int i = 0; i < kCtStages - 1; ++i) {
    if (i < units) issue(i, i);
    else cp_commit();
  }
  float acc1[MODE == 1 ? NT : 1][4];  // dgrad accumulators over the two passes
  for (long long v = 0; v < units; ++v) {
    const int it = (int)(v % kCtStages);
    cp_wait<kCtStages - 2>();  // unit v has landed (groups complete in order)
    __syncthreads();           // ... for every thread; everyone is done with unit v-1 (also orders the weight build)
    if (v + kCtStages - 1 < units) issue(v + kCtStages - 1, (int)((v + kCtStages - 1) % kCtStages));  // refills unit v-1's buffer
    else cp_commit();
    const long long tile = blockIdx.x + (v / NPASS) * gridDim.x;
    const int pass = (int)(v % NPASS);
    const float* At_ = At + (size_t)it * kCtTile * PA + (size_t)(warp * 16) * PA;
    uint32_t af[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      af[ks][0] = tf32b(At_[g * PA + 8 * ks + t]);
      af[ks][1] = tf32b(At_[(g + 8) * PA + 8 * ks + t]);
      af[ks][2] = tf32b(At_[g * PA + 8 * ks + t + 4]);
      af[ks][3] = tf32b(At_[(g + 8) * PA + 8 * ks + t + 4]);
    }
    // the two pixel rows this thread holds in its accumulators
    const long long p0 = tile * kCtTile + warp * 16 + g, p1 = p0 + 8;
    if (MODE == 0) {
      size_t ob[2];
      bool okr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const long long p = r ? p1 : p0;
        okr[r] = p < a.npix;
        int b, hy, wx;
        ct_split(a, okr[r] ? (int)p : 0, b, hy, wx);
        ob[r] = (((size_t)b * 2 * a.H + 2 * hy) * 2 * a.W + 2 * wx) * a.Cout;  // output pixel (2h, 2w)
      }
      const int ntl = N >> 3;
      for (int nt = 0; nt < ntl; nt += 2) {  // two n-tiles in flight
        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = n_base + 8 * (nt + j) + 2 * t;
          const int co = n % a.Cout;
          const float b0 = a.bias != nullptr ? __ldg(a.bias + co) : 0.f, b1 = a.bias != nullptr ? __ldg(a.bias + co + 1) : 0.f;
          acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b0; acc[j][3] = b1;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float* bp = Wm + (8 * ks + t) * pB + 8 * (nt + j) + g;
            ct_mma(acc[j], af[ks][0], af[ks][1], af[ks][2], af[ks][3], __float_as_uint(bp[0]), __float_as_uint(bp[4 * pB]));
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = n_base + 8 * (nt + j) + 2 * t;
          const int ac = n / a.Cout, co = n - ac * a.Cout;
          const size_t qo = ((size_t)(ac >> 1) * 2 * a.W + (ac & 1)) * a.Cout + co;  // (a, c) offset inside the 2x2 block
          if (a.round_out) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = round_tf32(acc[j][e]);
          }
          if (okr[0]) *reinterpret_cast<float2*>(a.y + ob[0] + qo) = make_float2(acc[j][0], acc[j][1]);
          if (okr[1]) *reinterpret_cast<float2*>(a.y + ob[1] + qo) = make_float2(acc[j][2], acc[j][3]);
        }
      }
    } else {
      if (pass == 0) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc1[nt][e] = 0.f;
      }
      const float* Wp = Wm + (size_t)pass * K * pB;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float* bp = Wp + (8 * ks + t) * pB + 8 * nt + g;
          ct_mma(acc1[nt], af[ks][0], af[ks][1], af[ks][2], af[ks][3], __float_as_uint(bp[0]), __float_as_uint(bp[4 * pB]));
        }
      }
      if (pass == 1) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int ci = n_base + 8 * nt + 2 * t;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const long long p = r ? p1 : p0;
            if (p >= a.npix) continue;
            float2 o = make_float2(acc1[nt][2 * r], acc1[nt][2 * r + 1]);
            if (a.mask_in) {  // ReLU mask of the layer that produced x
              const float2 xv = *reinterpret_cast<const float2*>(a.x + p * a.Cin + ci);
              o.x = xv.x > 0.f ? o.x : 0.f;
              o.y = xv.y > 0.f ? o.y : 0.f;
            }
            *reinterpret_cast<float2*>(a.dx + p * a.Cin + ci) = o;
          }
        }
      }
    }
  }
}

// ---- wgrad + bias gradient ---------------------------------------------------------------------------------------------
// 256 threads.  D[ci][n] (MT m16 tiles x N/8 n8 tiles) is split over the warps: up to 16 tiles per warp; when there are
// fewer than 8 x 16 tiles the remaining warps split the pixels (k-steps) of every stage and are reduced at the end.
// gridDim.y == 4: one (a,c) quadrant of the columns per CTA (wide layers: fewer atomics per CTA, more CTAs).
template <int TPW>
__global__ void __launch_bounds__(256) convT2x2_dw_mma_kernel(const CtArgs a) {
  extern __shared__ __align__(16) float sm[];
  const bool quadrant = gridDim.y == 4;
  const int N = quadrant ? a.Cout : 4 * a.Cout;  // columns of this CTA
  const int n_base = quadrant ? blockIdx.y * a.Cout : 0;
  const int pX = ct_pitch_b(a.Cin < 16 ? 16 : a.Cin), pD = ct_pitch_b(N);
  const int stage_f = kCtTile * (pX + pD);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nt8 = N >> 3, mt_n = (a.Cin + 15) >> 4;
  const int tiles = mt_n * nt8;
  const int nchunk = (tiles + 15) / 16;           // warps needed to cover D once (1, 2, 4 or 8)
  const int ksplit = 8 / nchunk;                  // warps sharing a chunk split the 8 k-steps of a stage
  const int chunk = warp % nchunk, kpart = warp / nchunk;
  // this warp's tiles: ti = chunk*16 + j, j < TPW, with (m-tile, n-tile) = (ti / nt8, ti % nt8)

  const long long nstage = (a.npix + kCtTile - 1) / kCtTile;
  const int prow = tid >> 2, q4 = tid & 3;  // staging: four threads per pixel
  auto issue = [&](long long s, int buf) {
    float* Xs = sm + (size_t)buf * stage_f;
    float* Ds = Xs + kCtTile * pX;
    const long long p = s * kCtTile + prow;
    const bool ok = p < a.npix;
    const long long pp = ok ? p : 0;
    const float* xs = a.x + pp * a.Cin;
    for (int cu = q4; cu < a.Cin / 4; cu += 4) cp16(Xs + prow * pX + 4 * cu, xs + 4 * cu, ok);
    int b, hy, wx;
    ct_split(a, (int)pp, b, hy, wx);
    if (quadrant) {  // Cout contiguous floats of output pixel (2h + a, 2w + c)
      const float* ds = a.dy + (((size_t)b * 2 * a.H + 2 * hy + (blockIdx.y >> 1)) * 2 * a.W + 2 * wx + (blockIdx.y & 1)) * a.Cout;
      for (int cu = q4; cu < a.Cout / 4; cu += 4) cp16(Ds + prow * pD + 4 * cu, ds + 4 * cu, ok);
    } else {         // both output rows: 2 * Cout contiguous floats each
      const int ar = q4 >> 1;
      const float* ds = a.dy + (((size_t)b * 2 * a.H + 2 * hy + ar) * 2 * a.W + 2 * wx) * a.Cout;
      float* dd = Ds + prow * pD + ar * 2 * a.Cout;
      for (int cu = (q4 & 1); cu < a.Cout / 2; cu += 2) cp16(dd + 4 * cu, ds + 4 * cu, ok);
    }
    cp_commit();
  };
  if (a.Cin < 16) {  // rows Cin..15 of the single m-tile: keep the staged pad columns finite (zero)
    for (int i = tid; i < kCtStages * kCtTile * 8; i += 256) {
      const int buf = i / (kCtTile * 8), r = (i / 8) % kCtTile, c = 8 + (i & 7);
      sm[(size_t)buf * stage_f + r * pX + c] = 0.f;
    }
  }
  float acc[TPW][4];
#pragma unroll
  for (int j = 0; j < TPW; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  // bias gradient: thread = (column quad, row group); N/4 quads, 256 / (N/4) row groups
  const int nquad = N >> 2;
  const int rgroups = 256 / nquad, quad = tid % nquad, rg = tid / nquad;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);

  const long long units = blockIdx.x < nstage ? (nstage - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;  // this CTA's stages
  for (int i = 0; i < kCtStages - 1; ++i) {
    if (i < units) issue(blockIdx.x + (long long)i * gridDim.x, i);
    else cp_commit();
  }
  for (long long v = 0; v < units; ++v) {
    cp_wait<kCtStages - 2>();
    __syncthreads();
    if (v + kCtStages - 1 < units) issue(blockIdx.x + (v + kCtStages - 1) * gridDim.x, (int)((v + kCtStages - 1) % kCtStages));
    else cp_commit();
    const float* Xs = sm + (size_t)(v % kCtStages) * stage_f;
    const float* Ds = Xs + kCtTile * pX;
    for (int ks = kpart; ks < 8; ks += ksplit) {
      const float* xk = Xs + (8 * ks + t) * pX + g;
      const float* dk = Ds + (8 * ks + t) * pD + g;
      uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      int ti = chunk * 16, mtj = ti / nt8, ntj = ti - mtj * nt8;
#pragma unroll
      for (int j = 0; j < TPW; ++j) {
        if (j == 0 || ntj == 0) {  // new m-tile: reload the A fragments
          const float* xr = xk + mtj * 16;
          a0 = tf32b(xr[0]); a1 = tf32b(xr[8]); a2 = tf32b(xr[4 * pX]); a3 = tf32b(xr[4 * pX + 8]);
        }
        ct_mma(acc[j], a0, a1, a2, a3, tf32b(dk[8 * ntj]), tf32b(dk[4 * pD + 8 * ntj]));
        if (++ntj == nt8) { ntj = 0; ++mtj; }
      }
    }
    if (a.db != nullptr && rg < rgroups) {
      for (int r = rg; r < kCtTile; r += rgroups) {
        const float4 dv = *reinterpret_cast<const float4*>(Ds + r * pD + 4 * quad);
        bsum.x += dv.x; bsum.y += dv.y; bsum.z += dv.z; bsum.w += dv.w;
      }
    }
  }
  cp_wait<0>();
  // ---- reduce the k-split warps through shared memory, then one atomic per output and CTA
  float* red = sm;  // [8 warps][TPW][32 lanes][4]
  __syncthreads();
#pragma unroll
  for (int j = 0; j < TPW; ++j)
    *reinterpret_cast<float4*>(red + (((size_t)warp * TPW + j) * 32 + lane) * 4) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
  __syncthreads();
  for (int i = tid; i < nchunk * TPW * 128; i += 256) {
    const int e = i & 3, ln = (i >> 2) & 31, j = (i >> 7) % TPW, ch = i / (128 * TPW);
    float sacc = 0.f;
    for (int kp = 0; kp < ksplit; ++kp) sacc += red[(((size_t)(kp * nchunk + ch) * TPW + j) * 32 + ln) * 4 + e];
    const int mtc = (ch * 16 + j) / nt8, ntc = (ch * 16 + j) % nt8;
    const int ci = mtc * 16 + (ln >> 2) + (e >> 1) * 8;
    const int n = n_base + 8 * ntc + 2 * (ln & 3) + (e & 1);
    if (ci < a.Cin && mtc < mt_n) {
      const int ac = n / a.Cout, co = n - ac * a.Cout;
      atomicAdd(a.dw + ((size_t)ci * a.Cout + co) * 4 + ac, sacc);
    }
  }
  if (a.db != nullptr) {  // row groups (and the (a,c) quadrants held by this CTA) -> Cout sums per CTA, one atomic each
    __syncthreads();
    float* bred = sm;  // [rgroups][N]
    if (rg < rgroups) *reinterpret_cast<float4*>(bred + rg * N + 4 * quad) = bsum;
    __syncthreads();
    for (int co = tid; co < a.Cout; co += 256) {
      float sacc = 0.f;
      for (int r = 0; r < rgroups; ++r)
        for (int q = 0; q < N / a.Cout; ++q) sacc += bred[r * N + q * a.Cout + co];
      atomicAdd(a.db + co, sacc);
    }
  }
}

// ---- wgrad for the 8 -> 8 layer at the top of the decoder ---------------------------------------------------------------
// 8 x 32 outputs from 262 k pixels: the MMA kernel above is synchronisation-bound there (64-pixel stages, 4 MMAs per warp
// and stage).  Streaming version: thread = (input pixel, 2x2 quadrant): 8 x-values x 8 dy-values -> 64 + 8 partial sums in
// registers over a strip of pixels, warp-shuffle + shared-memory reduction, one atomic per output and CTA.
__global__ void __launch_bounds__(128) convT2x2_dw_c8_kernel(const CtArgs a) {
  float acc[8][8], bacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bacc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }
  const int q = threadIdx.x & 3;  // quadrant (a, c) = (q >> 1, q & 1): fixed per thread, blockDim.x % 4 == 0
  const long long nitems = a.npix * 4;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < nitems; it += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(it >> 2);
    int b, hy, wx;
    ct_split(a, p, b, hy, wx);
    const float4 x0 = ldg4(a.x + (size_t)p * 8), x1 = ldg4(a.x + (size_t)p * 8 + 4);
    const float* dp = a.dy + (((size_t)b * 2 * a.H + 2 * hy + (q >> 1)) * 2 * a.W + 2 * wx + (q & 1)) * 8;
    const float4 g0 = ldg4(dp), g1 = ldg4(dp + 4);
    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv[i], gv[j], acc[i][j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) bacc[j] += gv[j];
  }
  // lanes with equal (lane & 3) hold the same quadrant: reduce over lane bits 2..4, then over the warps
  __shared__ float red[4][4][72];  // [warp][quadrant][ci*8+co | 64+co]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = i < 8 ? acc[i][j] : bacc[j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) red[warp][lane][i * 8 + j] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * 72; i += blockDim.x) {
    const int qq = i / 72, r = i - qq * 72;
    const float v = red[0][qq][r] + red[1][qq][r] + red[2][qq][r] + red[3][qq][r];
    if (r < 64) atomicAdd(a.dw + ((size_t)(r >> 3) * 8 + (r & 7)) * 4 + qq, v);  // dw[ci][co][a][c]
    else if (a.db != nullptr) atomicAdd(a.db + (r - 64), v);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static int ct_log2(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return (1 << s) == v ? s : -1;
}

static bool ct_shape_ok(int Cin, int Cout) {
  return Cin == Cout && (Cin == 8 || Cin == 16 || Cin == 32 || Cin == 64);
}

template <typename Kern>
static int ct_launch(Kern kern, dim3 grid, int threads, size_t smem, cudaStream_t st, const CtArgs& a, const char* what) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
  }
  kern<<<grid, threads, smem, st>>>(a);
  return post_launch(what);
}

bool convT2x2_mma_ok(const float* x, const float* dy_or_y, int Cin, int Cout, long long npix) {
  return ct_shape_ok(Cin, Cout) && aligned16(x) && aligned16(dy_or_y) && npix < (1LL << 31) - 64;
}

int convT2x2_fwd_mma(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin, int Cout, int round_out,
                     cudaStream_t st) {
  CtArgs a{};
  a.x = x; a.w = w; a.bias = bias; a.y = y; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.round_out = round_out;
  a.npix = (long long)B * H * W;
  a.wsh = ct_log2(W); a.hsh = ct_log2(H);
  const long long ntiles = (a.npix + kCtTile - 1) / kCtTile;
  const size_t smem = ((size_t)Cin * ct_pitch_b(4 * Cout) + 2 * kCtTile * (Cin + 4)) * sizeof(float);
  const int nsplit = Cin >= 32 ? 4 : 1;  // one (a,c) quadrant per CTA for the wide layers
  const size_t smem2 = ((size_t)Cin * ct_pitch_b(4 * Cout / nsplit) + (size_t)kCtStages * kCtTile * (Cin + 4)) * sizeof(float);
  const long long cap = 16LL * kNumSMs / nsplit;  // tiles are tiny: many CTAs per SM hide the load latency
  dim3 grid((unsigned)(ntiles < cap ? ntiles : cap), nsplit);
  (void)smem;
  switch (Cin) {
    case 8: return ct_launch(convT2x2_px_mma_kernel<0, 1, 1>, grid, 128, smem2, st, a, "pu_convT2x2s2_fwd (mma)");
    case 16: return ct_launch(convT2x2_px_mma_kernel<0, 2, 1>, grid, 128, smem2, st, a, "pu_convT2x2s2_fwd (mma)");
    case 32: return ct_launch(convT2x2_px_mma_kernel<0, 4, 1>, grid, 128, smem2, st, a, "pu_convT2x2s2_fwd (mma)");
    default: return ct_launch(convT2x2_px_mma_kernel<0, 8, 1>, grid, 128, smem2, st, a, "pu_convT2x2s2_fwd (mma)");
  }
}

int convT2x2_dx_mma(const float* x, const float* w, const float* dy, float* dx, int B, int H, int W, int Cin, int Cout, int mask_in,
                    cudaStream_t st) {
  CtArgs a{};
  a.x = x; a.w = w; a.dy = dy; a.dx = dx; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.mask_in = mask_in;
  a.npix = (long long)B * H * W;
  a.wsh = ct_log2(W); a.hsh = ct_log2(H);
  const long long ntiles = (a.npix + kCtTile - 1) / kCtTile;
  const int K = 2 * Cout;
  const int nsplit = Cin >= 32 ? Cin / 16 : 1;  // 16 input channels (two n-tiles) per CTA for the wide layers
  const size_t smem = ((size_t)2 * K * ct_pitch_b(Cin / nsplit) + (size_t)kCtStages * kCtTile * (K + 4)) * sizeof(float);
  const long long cap = 16LL * kNumSMs / nsplit;
  dim3 grid((unsigned)(ntiles < cap ? ntiles : cap), nsplit);
  switch (Cout) {
    case 8: return ct_launch(convT2x2_px_mma_kernel<1, 2, 1>, grid, 128, smem, st, a, "pu_convT2x2s2_bwd dx (mma)");
    case 16: return ct_launch(convT2x2_px_mma_kernel<1, 4, 2>, grid, 128, smem, st, a, "pu_convT2x2s2_bwd dx (mma)");
    case 32: return ct_launch(convT2x2_px_mma_kernel<1, 8, 2>, grid, 128, smem, st, a, "pu_convT2x2s2_bwd dx (mma)");
    default: return ct_launch(convT2x2_px_mma_kernel<1, 16, 2>, grid, 128, smem, st, a, "pu_convT2x2s2_bwd dx (mma)");
  }
}

// dw and db must be zeroed by the caller (fp32 atomics across CTAs)
int convT2x2_dw_mma(const float* x, const float* dy, float* dw, float* db, int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  CtArgs a{};
  a.x = x; a.dy = dy; a.dw = dw; a.db = db; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
  a.npix = (long long)B * H * W;
  a.wsh = ct_log2(W); a.hsh = ct_log2(H);
  if (Cin == 8 && Cout == 8) {  // streaming kernel: >= 16 items per thread amortise the 72-value reduction
    long long blocks = (a.npix * 4 + 128 * 16 - 1) / (128 * 16);
    if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
    if (blocks < 1) blocks = 1;
    convT2x2_dw_c8_kernel<<<(unsigned)blocks, 128, 0, st>>>(a);
    return post_launch("pu_convT2x2s2_bwd dw (c8)");
  }
  const long long nstage = (a.npix + kCtTile - 1) / kCtTile;
  const int nsplit = Cin >= 32 ? 4 : 1;  // one (a,c) quadrant of the columns per CTA for the wide layers
  const int N = 4 * Cout / nsplit;
  const int pX = ct_pitch_b(Cin < 16 ? 16 : Cin), pD = ct_pitch_b(N);
  const int tiles = ((Cin + 15) / 16) * (N / 8);
  const int tpw = tiles < 16 ? tiles : 16;
  size_t smem = (size_t)kCtStages * kCtTile * (pX + pD) * sizeof(float);
  const size_t red = (size_t)8 * tpw * 128 * sizeof(float);
  if (smem < red) smem = red;
  // The stages are tiny (64 pixels), so the kernel is latency-bound unless several CTAs share an SM; every CTA ends with
  // Cin*N atomics, which bounds their number for the wide layers; one wave (<= 4 CTAs per SM).
  long long ncta = nstage / 2;
  const long long by_atomics = 200000LL / ((long long)Cin * N);
  if (ncta > by_atomics) ncta = by_atomics;
  if (ncta > 4LL * kNumSMs / nsplit) ncta = 4LL * kNumSMs / nsplit;
  if (ncta < 1) ncta = 1;
  dim3 grid((unsigned)ncta, nsplit);
  switch (tpw) {
    case 4: return ct_launch(convT2x2_dw_mma_kernel<4>, grid, 256, smem, st, a, "pu_convT2x2s2_bwd dw (mma)");
    case 8: return ct_launch(convT2x2_dw_mma_kernel<8>, grid, 256, smem, st, a, "pu_convT2x2s2_bwd dw (mma)");
    default: return ct_launch(convT2x2_dw_mma_kernel<16>, grid, 256, smem, st, a, "pu_convT2x2s2_bwd dw (mma)");
  }
}

}  // namespace pu
