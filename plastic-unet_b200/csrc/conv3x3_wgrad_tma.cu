// conv3x3_wgrad_tma.cu — weight gradient of the 3x3 convolutions in TF32 mode: TMA-fed, warp-level tensor-core MMAs.
//
//   dw[co][ci][ky][kx] = sum_p x[p + (ky-1, kx-1)][ci] * g[p][co],   db[co] = sum_p g[p][co]
//
// (the parameter gradients autograd computes for nn.Conv2d(k=3, p=1) of reference unet_p.py:105-116,161-166 and
// unet_p_res.py:150-158,186-189,215-219.)
//
// Why not tcgen05: the contraction runs over PIXELS, and a tcgen05.mma contracts 8 tf32 elements (32 bytes of K) per
// instruction.  The outputs are tiny (8..64 ci x 8..64 co per tap), so one instruction would do a 64 x 64 x 8 product at
// best: 64->64 @16x16, B = 64 needs 18.4 k of them = 124 per SM at ~58 clk (measured issue cost) = 3.7 us —
// no better than the 4.5 us the same MACs take as warp-level mma.sync.m16n8k8 (measured 8 clk per SM sub-partition), and
// for the 8-channel layers (M = 64 minimum, 8 useful rows) an order of magnitude worse.  Every conv layer of the U-Net
// has the same MAC count, so 4.5 us of mma.sync time per layer is the floor, far below the HBM time of the big layers.
//
// What bounded the first mma.sync kernel (conv3x3_wgrad_mma_kernel, ncu: profiles/r2_wgrad_ncu_summary.md) was not the
// tensor pipe (23 % active) but instruction issue: 28 M warp instructions for 1.3 M MMAs — 25 % IMAD/LEA address
// arithmetic of the per-thread cp.async staging, 22 % shared-memory fragment loads (20 per 5 MMAs), and an fp32-atomic
// tail of up to 760 k atomics on ~1 k addresses.  This kernel removes all three:
//   * staging is TMA: one elected thread issues one 4-D box (8 channels, x, y, images) per operand plane into a
//     multi-stage mbarrier ring — no per-thread index arithmetic, zero fill outside the image (= conv padding = crop);
//   * every operand plane is [pixel][8 channels] (32-byte rows): the m16n8k8 A fragments (rows = ci, k = pixels) and
//     B fragments (k = pixels, n = co) are conflict-free 32-bit loads (bank = 8 * pixel + channel);
//   * a warp walks DOWN an 8-pixel-wide strip and keeps the three input rows of its 3x3 window in registers: 6 new A
//     loads per 8-pixel group instead of 18;
//   * one CTA per SM (16 MMA warps + 1 producer warp): 148 x (outputs) atomics instead of ~600 x.
// Work split: a CTA owns NCI 8-channel input chunks x NCO 8-channel output tiles (template), warps = NCI chunk groups x
// (16 / NCI) pixel strips; grid.y enumerates the channel groups, grid.x splits the pixels (tiles dealt round-robin).
// The bias gradient comes from the unused rows 8..15 of the fifth tap-pair tile (A = ones there).
#include <cuda.h>
#include <stdlib.h>
#include "conv3x3.cuh"
#include "tma_mma.cuh"

namespace pu {

namespace {

constexpr int kWgWarps = 16;                     // MMA warps
constexpr int kWgThreads = 32 * (kWgWarps + 1);  // + the TMA producer warp
constexpr int kWgMaxStages = 4;

struct WgTmaArgs {
  float* dw;
  float* db;  // | null
  int B, H, W, Cin, Cout;
  int n0;                    // 8-channel chunks of source 0 (the rest belong to source 1)
  int ncg;                   // channel-chunk groups (Cin / 8 / NCI); blockIdx.y = cg + ncg * (co group)
  int TW, TH, NB;            // tile: NB images x TH x TW output pixels, TW % 8 == 0, TH % 8 == 0
  int tilesX, tilesY, tilesB, ntiles;
  int xplane, gplane;        // bytes of one staged plane (128-byte multiples)
  int xbox, gbox;            // bytes one TMA box delivers
  int nstages;
  int merged;                // bit 0 / 1 / 2: the map of source 0 / source 1 / g is the 3-D (channel, x)-merged form (C == 8)
  int vec4;                  // dw is 16-byte aligned: red.global.add.v4.f32
  int debug;
};

template <int NCI, int NCO>
__global__ void __launch_bounds__(kWgThreads, 1) conv3x3_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmx0,
                                                                          const __grid_constant__ CUtensorMap tmx1,
                                                                          const __grid_constant__ CUtensorMap tmg, const WgTmaArgs a) {
  constexpr int PS = kWgWarps / NCI;  // 8x8-pixel strips of a tile (one per warp of a chunk group)

  extern __shared__ uint8_t wg_smem_raw[];
  uint8_t* smem = wg_smem_raw + ((128u - (smem_addr(wg_smem_raw) & 127u)) & 127u);
  const int stage_bytes = NCI * a.xplane + NCO * a.gplane;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.nstages * stage_bytes);  // full[kWgMaxStages], empty[kWgMaxStages]
  const uint32_t bar0 = smem_addr(bars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cg = (int)blockIdx.y % a.ncg, cog = (int)blockIdx.y / a.ncg;
  const int co0 = cog * 8 * NCO;

  if (tid == 0) {
    for (int i = 0; i < kWgMaxStages; ++i) {
      wg_mbar_init(bar0 + 8u * i, 1);                          // full: the producer's expect_tx arrival
      wg_mbar_init(bar0 + 8u * (kWgMaxStages + i), kWgWarps);  // empty: one arrival per MMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmg)) : "memory");
  }
  __syncthreads();
  pdl_prologue();  // x and g are written by the preceding kernels of the stream

  float acc[NCO][5][4];
#pragma unroll
  for (int n = 0; n < NCO; ++n)
#pragma unroll
    for (int p = 0; p < 5; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[n][p][j] = 0.f;

  const int tiles_img = a.tilesX * a.tilesY;
  if (warp == kWgWarps) {
    // ================= TMA producer =================
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int tb = tile / tiles_img, tr = tile - tb * tiles_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH, b0 = tb * a.NB;
      const int st = k % a.nstages;
      const uint32_t ph = (uint32_t)(k / a.nstages) & 1;
      wg_mbar_wait(bar0 + 8u * (kWgMaxStages + st), ph ^ 1);  // passes immediately on a fresh barrier
      if (lane == 0) {
        const uint32_t full = bar0 + 8u * st;
        wg_mbar_expect_tx(full, (uint32_t)(NCI * a.xbox + NCO * a.gbox));
        const uint32_t sS = smem_addr(smem + (size_t)st * stage_bytes);
#pragma unroll
        for (int c = 0; c < NCI; ++c) {
          const int chunk = cg * NCI + c;
          const bool second = chunk >= a.n0;
          if (a.merged & (second ? 2 : 1))  // an 8-channel source: its single chunk as (channel, x)-merged rows
            wg_tma_load_3d(sS + c * a.xplane, second ? &tmx1 : &tmx0, full, 8 * (x0 - 1), y0 - 1, b0);
          else
            wg_tma_load_4d(sS + c * a.xplane, second ? &tmx1 : &tmx0, full, 8 * (second ? chunk - a.n0 : chunk), x0 - 1, y0 - 1, b0);
        }
        if (a.merged & 4) {
          wg_tma_load_3d(sS + NCI * a.xplane, &tmg, full, 8 * x0, y0, b0);
        } else {
#pragma unroll
          for (int n = 0; n < NCO; ++n) wg_tma_load_4d(sS + NCI * a.xplane + n * a.gplane, &tmg, full, co0 + 8 * n, x0, y0, b0);
        }
      }
      __syncwarp();
    }
  } else {
    // ================= MMA warps =================
    const int gq = lane >> 2, tq = lane & 3;  // mma fragment coordinates: groupID, threadID_in_group
    const int pg = warp / PS, ps = warp - pg * PS;
    const int tws = a.TW >> 3, ths = a.TH >> 3;
    const int sx = ps % tws, sy = (ps / tws) % ths, nb = ps / (tws * ths);
    const int PW = a.TW + 2;
    // word offsets of this lane's first fragment elements inside an x plane / a g plane
    const int xoff = ((nb * (a.TH + 2) + sy * 8) * PW + sx * 8 + tq) * 8 + gq;
    const int goff = ((nb * a.TH + sy * 8) * a.TW + sx * 8 + tq) * 8 + gq;
    const int xrow = PW * 8, grow = a.TW * 8;  // words per plane row
    const int gpl = a.gplane >> 2;
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int tb = tile / tiles_img, tr = tile - tb * tiles_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int st = k % a.nstages;
      const uint32_t ph = (uint32_t)(k / a.nstages) & 1;
      wg_mbar_wait(bar0 + 8u * st, ph);
      // strips that lie entirely outside the image (ragged sizes, batch tail) hold zeros only
      const bool live = tx * a.TW + sx * 8 < a.W && ty * a.TH + sy * 8 < a.H && tb * a.NB + nb < a.B && !(a.debug & 4);
      if (live) {
        const unsigned* xs = reinterpret_cast<const unsigned*>(smem + (size_t)st * stage_bytes + pg * a.xplane) + xoff;
        const unsigned* gs = reinterpret_cast<const unsigned*>(smem + (size_t)st * stage_bytes + NCI * a.xplane) + goff;
        // win[r % 3][kx][half]: the A-fragment elements of halo row r: pixel (tq + kx) and (tq + kx + 4), channel gq
        unsigned win[3][3][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            win[r][kx][0] = xs[r * xrow + kx * 8];
            win[r][kx][1] = xs[r * xrow + (kx + 4) * 8];
          }
#pragma unroll
        for (int yy = 0; yy < 8; ++yy) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            win[(yy + 2) % 3][kx][0] = xs[(yy + 2) * xrow + kx * 8];
            win[(yy + 2) % 3][kx][1] = xs[(yy + 2) * xrow + (kx + 4) * 8];
          }
#pragma unroll
          for (int n = 0; n < NCO; ++n) {
            // B fragment: k = pixel (tq, tq + 4), n = co (gq)
            const unsigned b0 = gs[n * gpl + yy * grow];
            const unsigned b1 = gs[n * gpl + yy * grow + 32];
            if (a.debug & 2) {  // (experiment) fragment loads without MMAs
#pragma unroll
              for (int t = 0; t < 9; ++t)
                acc[n][t >> 1][t & 3] += __uint_as_float(win[(yy + t / 3) % 3][t % 3][0] ^ win[(yy + t / 3) % 3][t % 3][1] ^ b0 ^ b1);
              continue;
            }
            // tile p: rows 0-7 = tap 2p, rows 8-15 = tap 2p + 1 (ci = gq), k = the 8 pixels of this group
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const int t0 = 2 * p, t1 = 2 * p + 1;
              wg_mma(acc[n][p], win[(yy + t0 / 3) % 3][t0 % 3][0], win[(yy + t1 / 3) % 3][t1 % 3][0], win[(yy + t0 / 3) % 3][t0 % 3][1],
                     win[(yy + t1 / 3) % 3][t1 % 3][1], b0, b1);
            }
            // rows 8..15 of the fifth tile are free: ones there make them the column sums of g = the bias gradient
            wg_mma(acc[n][4], win[(yy + 2) % 3][2][0], 0x3f800000u, win[(yy + 2) % 3][2][1], 0x3f800000u, b0, b1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) wg_mbar_arrive(bar0 + 8u * (kWgMaxStages + st));
    }
  }

  // ---- reduce the PS warps of every chunk group through shared memory, then one atomic per output and CTA
  __syncthreads();  // every staged tile has been consumed: the ring is free
  float* red = reinterpret_cast<float*>(smem);  // [NACC][kWgWarps * 32]: conflict-free writes
  if (warp < kWgWarps) {
#pragma unroll
    for (int n = 0; n < NCO; ++n)
#pragma unroll
      for (int p = 0; p < 5; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[(n * 20 + p * 4 + j) * (kWgWarps * 32) + warp * 32 + lane] = acc[n][p][j];
  }
  __syncthreads();
  if (a.debug & 1) return;  // (experiment) no atomics
  // consecutive threads -> consecutive dw addresses: i = ((pg * NCO + n) * 8 + co) * 72 + ci * 9 + tap; the 72 floats of a
  // (co, chunk) row are contiguous and 16-byte aligned (72 = 4 * 18): one vector reduction per 4 outputs
  if (a.vec4) {
    for (int i = tid; i < NCI * NCO * 8 * 18; i += kWgThreads) {
      const int v4 = i % 18, q = i / 18;
      const int co = q & 7, n = (q >> 3) % NCO, pg = q / (8 * NCO);
      float sum[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = 4 * v4 + e;
        const int ci = r / 9, tap = r - ci * 9;
        const int ln = ci * 4 + (co >> 1), p = tap >> 1, j = ((tap & 1) << 1) | (co & 1);
        float t = 0.f;
#pragma unroll
        for (int s = 0; s < PS; ++s) t += red[(n * 20 + p * 4 + j) * (kWgWarps * 32) + (pg * PS + s) * 32 + ln];
        sum[e] = t;
      }
      wg_red_add_v4(a.dw + ((size_t)(co0 + 8 * n + co) * a.Cin + (cg * NCI + pg) * 8) * 9 + 4 * v4, sum[0], sum[1], sum[2], sum[3]);
    }
  } else
  for (int i = tid; i < NCI * NCO * 576; i += kWgThreads) {
    const int r = i % 72, q = i / 72;
    const int co = q & 7, n = (q >> 3) % NCO, pg = q / (8 * NCO);
    const int ci = r / 9, tap = r - ci * 9;
    const int ln = ci * 4 + (co >> 1), p = tap >> 1, j = ((tap & 1) << 1) | (co & 1);
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < PS; ++s) sum += red[(n * 20 + p * 4 + j) * (kWgWarps * 32) + (pg * PS + s) * 32 + ln];
    atomicAdd(a.dw + ((size_t)(co0 + 8 * n + co) * a.Cin + (cg * NCI + pg) * 8 + ci) * 9 + tap, sum);
  }
  if (a.db != nullptr && cg == 0 && tid < 8 * NCO) {  // the ones rows of chunk 0: every ci row holds sum_pixels g[.][co]
    const int co = tid & 7, n = tid >> 3;
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < PS; ++s) sum += red[(n * 20 + 16 + 2 + (co & 1)) * (kWgWarps * 32) + s * 32 + (co >> 1)];
    atomicAdd(a.db + co0 + 8 * n + co, sum);
  }
}

template <int NCI, int NCO>
int launch_wg(const CUtensorMap& tmx0, const CUtensorMap& tmx1, const CUtensorMap& tmg, const WgTmaArgs& wa, dim3 grid, size_t smem,
              cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_tma_kernel<NCI, NCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("conv3x3_wgrad_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaError_t le = launch_pdl(conv3x3_wgrad_tma_kernel<NCI, NCO>, grid, dim3(kWgThreads), smem, st, tmx0, tmx1, tmg, wa);
  if (le != cudaSuccess) {
    set_error("conv3x3_wgrad_tma launch: %s", cudaGetErrorString(le));
    return PU_ERR_CUDA;
  }
  return post_launch("conv3x3_wgrad_tma");
}

inline int pow2_le(int v, int cap) {
  int p = cap;
  while (p > 1 && v % p != 0) p >>= 1;
  return p;
}

}  // namespace

// Host-side plan, shared with pu_conv3x3_wgrad_plan (tests): channel split and tile geometry.
struct WgPlan {
  int nci, nco, TW, TH, NB, tilesX, tilesY, tilesB, nstages, gx, gy;
  size_t smem;
};

static bool wg_plan(int B, int H, int W, int C0, int C1, int Cout, WgPlan* p) {
  if (C0 < 8 || C0 % 8 != 0 || C1 < 0 || C1 % 8 != 0 || Cout < 8 || Cout % 8 != 0 || B < 1 || H < 1 || W < 1) return false;
  const int nchunks = (C0 + C1) / 8, ncot = Cout / 8;
  p->nci = pow2_le(nchunks, 4);
  p->nco = pow2_le(ncot, 2);  // 17 warps: 5 on one SM sub-partition (16 K registers) -> 96 registers per thread; 4 co tiles (80 accumulators) would spill
  const int PS = kWgWarps / p->nci;
  // tile = NB images x (8 ths) x (8 tws) pixels with tws * ths * NB == PS: least staged bytes over the whole problem
  double best = 1e30;
  p->TW = 0;
  for (int tws = 1; tws <= PS; tws <<= 1)
    for (int ths = 1; tws * ths <= PS; ths <<= 1) {
      const int nb = PS / (tws * ths);
      if (nb > 16) continue;  // box dimensions stay small
      // (channel, x)-merged boxes of the 8-channel tensors: a box row holds <= 256 elements
      if ((C0 == 8 || C1 == 8) && 8 * (8 * tws + 2) > 256) continue;
      if (Cout == 8 && 8 * 8 * tws > 256) continue;
      const int tw = 8 * tws, th = 8 * ths;
      const long long tiles = (long long)cdiv(W, tw) * cdiv(H, th) * cdiv(B, nb);
      const double cost = (double)tiles * ((double)(tw + 2) * (th + 2) * nb * p->nci + (double)tw * th * nb * p->nco + 256.0);
      if (cost < best) {
        best = cost;
        p->TW = tw; p->TH = th; p->NB = nb;
      }
    }
  if (p->TW == 0) return false;
  p->tilesX = cdiv(W, p->TW); p->tilesY = cdiv(H, p->TH); p->tilesB = cdiv(B, p->NB);
  const int xplane = ((p->TW + 2) * (p->TH + 2) * p->NB * 32 + 127) / 128 * 128;
  const int gplane = (p->TW * p->TH * p->NB * 32 + 127) / 128 * 128;
  const size_t stage = (size_t)p->nci * xplane + (size_t)p->nco * gplane;
  const size_t red = (size_t)kWgWarps * 32 * 20 * p->nco * sizeof(float);  // the final reduction reuses the ring
  const size_t budget = 218 * 1024;
  const long long ntiles = (long long)p->tilesX * p->tilesY * p->tilesB;
  p->gy = (nchunks / p->nci) * (ncot / p->nco);
  long long gx = kNumSMs / p->gy;
  if (gx < 1) gx = 1;
  if (gx > ntiles) gx = ntiles;
  p->gx = (int)gx;
  const int per_cta = (int)((ntiles + gx - 1) / gx);
  int ns = (int)(budget / stage);
  if (ns > kWgMaxStages) ns = kWgMaxStages;
  if (ns > per_cta) ns = per_cta;
  if (ns < 1) return false;
  p->nstages = ns;
  size_t ring = (size_t)ns * stage;
  if (ring < red) ring = red;
  p->smem = ring + 2 * kWgMaxStages * 8 + 128;  // barriers + alignment slack
  return p->smem <= 227 * 1024;
}

bool conv3x3_wgrad_tma_ok(const WgradArgs& a) {
  WgPlan p;
  const int C1 = (a.s1.p != nullptr) ? a.s1.C : 0;
  return wg_plan(a.B, a.H, a.W, a.s0.C, C1, a.Cout, &p);
}

int conv3x3_wgrad_tma(const WgradArgs& a, cudaStream_t st) {
  WgPlan p;
  const int C1 = (a.s1.p != nullptr) ? a.s1.C : 0;
  if (!wg_plan(a.B, a.H, a.W, a.s0.C, C1, a.Cout, &p)) {
    set_error("pu_conv3x3_wgrad: shape (C %d|%d -> %d, %dx%d) does not fit the TMA weight-gradient kernel", a.s0.C, C1, a.Cout, a.H, a.W);
    return PU_ERR_UNSUPPORTED;
  }
  CUtensorMap tmx0, tmx1, tmg;
  static int allow_merged = -1;
  if (allow_merged < 0) {
    const char* e_ = getenv("PU_WG_MERGED");  // 0: always the 4-D maps with 32-byte box rows (A/B measurements)
    allow_merged = (e_ != nullptr && e_[0] == '0') ? 0 : 1;
  }
  int merged = 0;
  int rc;
  if (a.s0.C == 8 && allow_merged) {
    merged |= 1;
    rc = tma_make_window_map_merged(&tmx0, a.s0, a.B, a.H, a.W, p.TW + 2, p.TH + 2, p.NB);
  } else {
    rc = tma_make_window_map(&tmx0, a.s0, a.B, a.H, a.W, 8, p.TW + 2, p.TH + 2, p.NB);
  }
  if (rc) return rc;
  if (C1 > 0) {
    if (C1 == 8 && allow_merged) {
      merged |= 2;
      rc = tma_make_window_map_merged(&tmx1, a.s1, a.B, a.H, a.W, p.TW + 2, p.TH + 2, p.NB);
    } else {
      rc = tma_make_window_map(&tmx1, a.s1, a.B, a.H, a.W, 8, p.TW + 2, p.TH + 2, p.NB);
    }
    if (rc) return rc;
  } else {
    tmx1 = tmx0;
  }
  const View gv{a.g, a.H, a.W, a.Cout, 0, 0};
  if (a.Cout == 8 && allow_merged) {
    merged |= 4;
    rc = tma_make_window_map_merged(&tmg, gv, a.B, a.H, a.W, p.TW, p.TH, p.NB);
  } else {
    rc = tma_make_window_map(&tmg, gv, a.B, a.H, a.W, 8, p.TW, p.TH, p.NB);
  }
  if (rc) return rc;
  WgTmaArgs wa;
  wa.dw = a.dw; wa.db = a.db;
  wa.B = a.B; wa.H = a.H; wa.W = a.W; wa.Cin = a.Cin; wa.Cout = a.Cout;
  wa.n0 = a.s0.C / 8;
  wa.ncg = (a.Cin / 8) / p.nci;
  wa.TW = p.TW; wa.TH = p.TH; wa.NB = p.NB;
  wa.tilesX = p.tilesX; wa.tilesY = p.tilesY; wa.tilesB = p.tilesB;
  wa.ntiles = p.tilesX * p.tilesY * p.tilesB;
  wa.xbox = (p.TW + 2) * (p.TH + 2) * p.NB * 32;
  wa.gbox = p.TW * p.TH * p.NB * 32;
  wa.xplane = (wa.xbox + 127) / 128 * 128;
  wa.gplane = (wa.gbox + 127) / 128 * 128;
  wa.nstages = p.nstages;
  wa.merged = merged;
  wa.vec4 = (reinterpret_cast<uintptr_t>(a.dw) & 15u) == 0 ? 1 : 0;
  wa.debug = a.debug;
  if (a.debug & 8)
    fprintf(stderr, "conv3x3_wgrad_tma plan: %d|%d->%d %dx%d B=%d: NCI=%d NCO=%d tile %dx%dx%d tiles %dx%dx%d stages=%d grid %dx%d smem=%zu\n",
            a.s0.C, C1, a.Cout, a.H, a.W, a.B, p.nci, p.nco, p.NB, p.TH, p.TW, p.tilesB, p.tilesY, p.tilesX, p.nstages, p.gx, p.gy, p.smem);
  const dim3 grid(p.gx, p.gy);
#define PU_WG_CASE(I, O) \
  if (p.nci == I && p.nco == O) return launch_wg<I, O>(tmx0, tmx1, tmg, wa, grid, p.smem, st);
  PU_WG_CASE(1, 1) PU_WG_CASE(1, 2)
  PU_WG_CASE(2, 1) PU_WG_CASE(2, 2)
  PU_WG_CASE(4, 1) PU_WG_CASE(4, 2)
#undef PU_WG_CASE
  set_error("conv3x3_wgrad_tma: no kernel for NCI=%d NCO=%d", p.nci, p.nco);
  return PU_ERR_UNSUPPORTED;
}

}  // namespace pu

// Host-only: the plan pu_conv3x3_wgrad (PU_MATH_TF32) uses — {NCI, NCO, TW, TH, NB, tilesX, tilesY, tilesB, stages, grid x, grid y, smem}
extern "C" int pu_conv3x3_wgrad_plan(int B, int H, int W, int C0, int C1, int Cout, int* out12) {
  pu::WgPlan p;
  if (out12 == nullptr || !pu::wg_plan(B, H, W, C0, C1, Cout, &p)) {
    pu::set_error("pu_conv3x3_wgrad_plan: shape (C %d|%d -> %d, %dx%d) does not fit the TMA weight-gradient kernel", C0, C1, Cout, H, W);
    return PU_ERR_UNSUPPORTED;
  }
  const int v[12] = {p.nci, p.nco, p.TW, p.TH, p.NB, p.tilesX, p.tilesY, p.tilesB, p.nstages, p.gx, p.gy, (int)p.smem};
  for (int i = 0; i < 12; ++i) out12[i] = v[i];
  return PU_OK;
}
