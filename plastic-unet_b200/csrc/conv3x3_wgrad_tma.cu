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
        // shared-space byte addresses of this lane's first fragment element in its x plane (halo row 0) and in the g planes
        const uint32_t sbase = smem_addr(smem) + (uint32_t)st * (uint32_t)stage_bytes;
        uint32_t xa = sbase + (uint32_t)(pg * a.xplane) + 4u * (uint32_t)xoff;
        uint32_t ga[NCO];
#pragma unroll
        for (int n = 0; n < NCO; ++n) ga[n] = sbase + (uint32_t)(NCI * a.xplane + n * a.gplane) + 4u * (uint32_t)goff;
        const uint32_t xrow_b = 4u * (uint32_t)xrow, grow_b = 4u * (uint32_t)grow;
        // win[r % 3][kx][half]: the A-fragment elements of halo row r: pixel (tq + kx) and (tq + kx + 4), channel gq
        unsigned win[3][3][2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          win[r][0][0] = wg_lds<0>(xa);   win[r][0][1] = wg_lds<128>(xa);
          win[r][1][0] = wg_lds<32>(xa);  win[r][1][1] = wg_lds<160>(xa);
          win[r][2][0] = wg_lds<64>(xa);  win[r][2][1] = wg_lds<192>(xa);
          xa += xrow_b;
        }
#pragma unroll
        for (int yy = 0; yy < 8; ++yy) {
          win[(yy + 2) % 3][0][0] = wg_lds<0>(xa);   win[(yy + 2) % 3][0][1] = wg_lds<128>(xa);
          win[(yy + 2) % 3][1][0] = wg_lds<32>(xa);  win[(yy + 2) % 3][1][1] = wg_lds<160>(xa);
          win[(yy + 2) % 3][2][0] = wg_lds<64>(xa);  win[(yy + 2) % 3][2][1] = wg_lds<192>(xa);
          xa += xrow_b;
#pragma unroll
          for (int n = 0; n < NCO; ++n) {
            // B fragment: k = pixel (tq, tq + 4), n = co (gq)
            const unsigned b0 = wg_lds<0>(ga[n]);
            const unsigned b1 = wg_lds<128>(ga[n]);
            ga[n] += grow_b;
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

// ---- the one-input-channel stem (reference unet_p.py:105 via inconv) ----------------------------------------------------------
// dw[co][0][tap] = sum_p x[p + tap] g[p][co], db[co] = sum_p g[p][co]: ONE MMA per 8 pixels and co tile — A rows 0..8 are the nine
// taps (row r, k = pixel j: x[p_j + tap_r], 32-bit loads from the [y][x] plane of the input), row 9 is ones (bias gradient),
// B = the g plane.  The plane rows are XW = TW + 8 (20 for TW = 8) words apart, so the three tap rows of a fragment load fall
// into disjoint banks; the box starts at x0 - 4 (16-byte aligned in memory), pixel x0 - 1 is column 3.  This is the LAST weight gradient of the backward pass (its g is the data gradient of the second conv):
// nothing overlaps it.  The thread-per-pixel kernel in stem.cu (kept for unaligned inputs) took 30 us for these 38 MB: 14 M warp
// instructions, four warps per scheduler waiting on dependent loads (ncu: profiles/r2_wgrad_ncu_summary.md).
struct WgStemArgs {
  float* dw;
  float* db;
  int B, H, W, Cout;
  int TW, TH, NB, XW;
  int tilesX, tilesY, tilesB, ntiles;
  int xplane, gplane, xbox, gbox;
  int nstages, gmerged, vec4;
};

__device__ __forceinline__ unsigned tf32_bits(unsigned x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(__uint_as_float(x)));
  return r;
}

template <int NCO>
__global__ void __launch_bounds__(kWgThreads, 1) conv3x3_c1_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                             const __grid_constant__ CUtensorMap tmg, const WgStemArgs a) {
  extern __shared__ uint8_t wg_smem_raw[];
  uint8_t* smem = wg_smem_raw + ((128u - (smem_addr(wg_smem_raw) & 127u)) & 127u);
  const int stage_bytes = a.xplane + NCO * a.gplane;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.nstages * stage_bytes);
  const uint32_t bar0 = smem_addr(bars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = (int)blockIdx.y * 8 * NCO;
  if (tid == 0) {
    for (int i = 0; i < kWgMaxStages; ++i) {
      wg_mbar_init(bar0 + 8u * i, 1);
      wg_mbar_init(bar0 + 8u * (kWgMaxStages + i), kWgWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmg)) : "memory");
  }
  __syncthreads();
  pdl_prologue();
  float acc[NCO][4];
#pragma unroll
  for (int n = 0; n < NCO; ++n)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[n][j] = 0.f;
  const int tiles_img = a.tilesX * a.tilesY;
  if (warp == kWgWarps) {
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int tb = tile / tiles_img, tr = tile - tb * tiles_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH, b0 = tb * a.NB;
      const int st = k % a.nstages;
      const uint32_t ph = (uint32_t)(k / a.nstages) & 1;
      wg_mbar_wait(bar0 + 8u * (kWgMaxStages + st), ph ^ 1);
      if (lane == 0) {
        const uint32_t full = bar0 + 8u * st;
        wg_mbar_expect_tx(full, (uint32_t)(a.xbox + NCO * a.gbox));
        const uint32_t sS = smem_addr(smem + (size_t)st * stage_bytes);
        wg_tma_load_3d(sS, &tmx, full, x0 - 4, y0 - 1, b0);  // the box starts 16-byte aligned in memory: column 3 = pixel x0 - 1
        if (a.gmerged) {
          wg_tma_load_3d(sS + a.xplane, &tmg, full, 8 * x0, y0, b0);
        } else {
#pragma unroll
          for (int n = 0; n < NCO; ++n) wg_tma_load_4d(sS + a.xplane + n * a.gplane, &tmg, full, co0 + 8 * n, x0, y0, b0);
        }
      }
      __syncwarp();
    }
  } else {
    const int gq = lane >> 2, tq = lane & 3;
    const int tws = a.TW >> 3, ths = a.TH >> 3;
    const int sx = warp % tws, sy = (warp / tws) % ths, nb = warp / (tws * ths);
    // rows 0..7 of A: tap gq = (ky, kx); row 8 (lane group 0): tap 8; row 9 (lane group 1): ones; rows 10..15: zero
    const int xoff = (nb * (a.TH + 2) + sy * 8 + gq / 3) * a.XW + sx * 8 + tq + gq % 3 + 3;
    const int xoff8 = (nb * (a.TH + 2) + sy * 8 + 2) * a.XW + sx * 8 + tq + 2 + 3;
    const int goff = ((nb * a.TH + sy * 8) * a.TW + sx * 8 + tq) * 8 + gq;
    const int grow = a.TW * 8, gpl = a.gplane >> 2;
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int tb = tile / tiles_img, tr = tile - tb * tiles_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int st = k % a.nstages;
      const uint32_t ph = (uint32_t)(k / a.nstages) & 1;
      wg_mbar_wait(bar0 + 8u * st, ph);
      const bool live = tx * a.TW + sx * 8 < a.W && ty * a.TH + sy * 8 < a.H && tb * a.NB + nb < a.B;
      if (live) {
        const unsigned* xs = reinterpret_cast<const unsigned*>(smem + (size_t)st * stage_bytes);
        const unsigned* gs = reinterpret_cast<const unsigned*>(smem + (size_t)st * stage_bytes + a.xplane) + goff;
#pragma unroll
        for (int yy = 0; yy < 8; ++yy) {
          // the network input is not stored TF32-rounded (the forward stem is an fp32 FFMA kernel): round here, RN like every
          // other tensor-core operand (the MMA itself would truncate)
          const unsigned a0 = tf32_bits(xs[xoff + yy * a.XW]), a2 = tf32_bits(xs[xoff + yy * a.XW + 4]);
          const unsigned t0 = tf32_bits(xs[xoff8 + yy * a.XW]), t2 = tf32_bits(xs[xoff8 + yy * a.XW + 4]);
          const unsigned a1 = gq == 0 ? t0 : (gq == 1 ? 0x3f800000u : 0u);
          const unsigned a3 = gq == 0 ? t2 : (gq == 1 ? 0x3f800000u : 0u);
#pragma unroll
          for (int n = 0; n < NCO; ++n) wg_mma(acc[n], a0, a1, a2, a3, gs[n * gpl + yy * grow], gs[n * gpl + yy * grow + 32]);
        }
      }
      __syncwarp();
      if (lane == 0) wg_mbar_arrive(bar0 + 8u * (kWgMaxStages + st));
    }
  }
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);  // [NCO * 4][kWgWarps * 32]
  if (warp < kWgWarps) {
#pragma unroll
    for (int n = 0; n < NCO; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(n * 4 + j) * (kWgWarps * 32) + warp * 32 + lane] = acc[n][j];
  }
  __syncthreads();
  // output e of this CTA: e < 72 NCO: dw[(co0 + e / 9)][tap e % 9] (contiguous floats); then db[co0 + e - 72 NCO]
  auto total = [&](int e) {
    int row, col;
    if (e < 72 * NCO) { col = e / 9; row = e - col * 9; } else { col = e - 72 * NCO; row = 9; }
    const int n = col >> 3, c = col & 7;
    const int ln = (row & 7) * 4 + (c >> 1), j = ((row >> 3) << 1) | (c & 1);
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWgWarps; ++w) t += red[(n * 4 + j) * (kWgWarps * 32) + w * 32 + ln];
    return t;
  };
  const int nout = (a.db != nullptr ? 80 : 72) * NCO;
  if (a.vec4) {
    for (int i = tid; i < nout / 4; i += kWgThreads) {
      const int e = 4 * i;
      float* dst = e < 72 * NCO ? a.dw + (size_t)co0 * 9 + e : a.db + co0 + (e - 72 * NCO);
      wg_red_add_v4(dst, total(e), total(e + 1), total(e + 2), total(e + 3));
    }
  } else {
    for (int e = tid; e < nout; e += kWgThreads) atomicAdd(e < 72 * NCO ? a.dw + (size_t)co0 * 9 + e : a.db + co0 + (e - 72 * NCO), total(e));
  }
}

template <int NCO>
int launch_wg_stem(const CUtensorMap& tmx, const CUtensorMap& tmg, const WgStemArgs& wa, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_c1_wgrad_tma_kernel<NCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("conv3x3_c1_wgrad_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaError_t le = launch_pdl(conv3x3_c1_wgrad_tma_kernel<NCO>, grid, dim3(kWgThreads), smem, st, tmx, tmg, wa);
  if (le != cudaSuccess) {
    set_error("conv3x3_c1_wgrad_tma launch: %s", cudaGetErrorString(le));
    return PU_ERR_CUDA;
  }
  return post_launch("conv3x3_wgrad_tma (stem)");
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

// The stem through TMA + MMA: needs a 16-byte aligned one-channel input window (else the streaming kernel of stem.cu runs).
bool conv3x3_c1_wgrad_tma_ok(const WgradArgs& a) {
  if (a.s0.C != 1 || (a.s1.p != nullptr && a.s1.C > 0) || a.Cout % 8 != 0 || a.Cout > 64) return false;
  const float* base = a.s0.p + (size_t)a.s0.oy * a.s0.Ws + a.s0.ox;
  return (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && a.s0.Ws % 4 == 0 && (reinterpret_cast<uintptr_t>(a.g) & 15u) == 0;
}

// dw / db zeroed by the caller
int conv3x3_c1_wgrad_tma(const WgradArgs& a, cudaStream_t st) {
  const int nco = (a.Cout / 8) % 2 == 0 ? 2 : 1;
  WgStemArgs wa;
  wa.dw = a.dw; wa.db = a.db; wa.B = a.B; wa.H = a.H; wa.W = a.W; wa.Cout = a.Cout;
  // tile: NB images x (8 ths) x (8 tws) pixels, 16 strips; least staged bytes
  double best = 1e30;
  wa.TW = 0;
  for (int tws = 1; tws <= kWgWarps; tws <<= 1)
    for (int ths = 1; tws * ths <= kWgWarps; ths <<= 1) {
      const int nb = kWgWarps / (tws * ths);
      if (nb > 16) continue;
      const int tw = 8 * tws, th = 8 * ths;
      if (a.Cout == 8 && 8 * tw > 256) continue;  // merged g rows
      const int xw = tw == 8 ? 20 : tw + 8;  // >= TW + 5 (the box starts at x0 - 4), rows of the three taps in disjoint banks
      const long long tiles = (long long)cdiv(a.W, tw) * cdiv(a.H, th) * cdiv(a.B, nb);
      const double cost = (double)tiles * ((double)xw * (th + 2) * nb / 8.0 + (double)tw * th * nb * nco + 256.0);
      if (cost < best) {
        best = cost;
        wa.TW = tw; wa.TH = th; wa.NB = nb; wa.XW = xw;
      }
    }
  wa.tilesX = cdiv(a.W, wa.TW); wa.tilesY = cdiv(a.H, wa.TH); wa.tilesB = cdiv(a.B, wa.NB);
  wa.ntiles = wa.tilesX * wa.tilesY * wa.tilesB;
  wa.xbox = wa.XW * (wa.TH + 2) * wa.NB * 4;
  wa.gbox = wa.TW * wa.TH * wa.NB * 32;
  wa.xplane = (wa.xbox + 127) / 128 * 128;
  wa.gplane = (wa.gbox + 127) / 128 * 128;
  const size_t stage = (size_t)wa.xplane + (size_t)nco * wa.gplane;
  int gx = wa.ntiles < kNumSMs ? wa.ntiles : kNumSMs;
  const int gy = a.Cout / (8 * nco);
  if (gy > 1) gx = gx / gy > 0 ? gx / gy : 1;
  const int per_cta = (wa.ntiles + gx - 1) / gx;
  int ns = (int)((size_t)218 * 1024 / stage);
  if (ns > kWgMaxStages) ns = kWgMaxStages;
  if (ns > per_cta) ns = per_cta;
  if (ns < 1) ns = 1;
  wa.nstages = ns;
  size_t ring = (size_t)ns * stage;
  const size_t red = (size_t)kWgWarps * 32 * 4 * nco * sizeof(float);
  if (ring < red) ring = red;
  const size_t smem = ring + 2 * kWgMaxStages * 8 + 128;
  CUtensorMap tmx, tmg;
  int rc = tma_make_plane_map(&tmx, a.s0, a.B, a.H, a.W, wa.XW, wa.TH + 2, wa.NB);
  if (rc) return rc;
  const View gv{a.g, a.H, a.W, a.Cout, 0, 0};
  wa.gmerged = a.Cout == 8 ? 1 : 0;
  if (wa.gmerged) rc = tma_make_window_map_merged(&tmg, gv, a.B, a.H, a.W, wa.TW, wa.TH, wa.NB);
  else rc = tma_make_window_map(&tmg, gv, a.B, a.H, a.W, 8, wa.TW, wa.TH, wa.NB);
  if (rc) return rc;
  wa.vec4 = ((reinterpret_cast<uintptr_t>(a.dw) | (a.db != nullptr ? reinterpret_cast<uintptr_t>(a.db) : 0)) & 15u) == 0 ? 1 : 0;
  const dim3 grid(gx, gy);
  if (nco == 2) return launch_wg_stem<2>(tmx, tmg, wa, grid, smem, st);
  return launch_wg_stem<1>(tmx, tmg, wa, grid, smem, st);
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
