// convT_dw_tma.cu — parameter gradients of ConvTranspose2d(k=2, s=2) (reference unet_p.py:155) in TF32 mode: TMA-fed,
// warp-level tensor-core MMAs, the scheme of conv3x3_wgrad_tma.cu without a halo.
//
//   dW[ci][co][a][c] = sum_{b,h,w} X[b,h,w,ci] * dY[b, 2h+a, 2w+c, co],      db[co] = sum dY[.., co]
//
// Row parity a of the output gradient is a strided VIEW of dY: dY_a[b,h,w,(c,co)] = dY[b, 2h+a, 2w+c, co] is the NHWC tensor
// [B, H, W, 2 Cout] with pixel pitch 2 Cout floats and row pitch 2 * (2W) * Cout floats — a plain 4-D tensor map, so the two
// parities arrive in shared memory as [pixel][8 channels] planes exactly like the operands of the 3x3 weight gradient.  One
// m16n8k8 MMA per 8 input pixels contracts a 16-row A tile — rows 0-7: 8 (c,co) channels of dY_0, rows 8-15: the same
// channels of dY_1 — with an 8-column B tile (8 input channels of X): D[(a, cc)][ci].  A second MMA against B = ones gives
// the column sums of dY, i.e. the bias gradient, on the CTAs of input-channel group 0.
//
// CTA = 16 MMA warps + 1 TMA producer warp, one per SM: warps = NPC (c,co)-chunks x (16 / NPC) strips of 8x8 pixels;
// every warp multiplies its chunk with NQ input-channel chunks.  grid.y = channel groups, grid.x splits the pixels.
// The first version (convT2x2_dw_mma_kernel / convT2x2_dw_c8_kernel: cp.async staging with per-thread index arithmetic,
// 64-pixel stages) took 18-30 us per layer for 5-42 MB of operands.
#include <cuda.h>
#include <stdlib.h>
#include "pu_common.cuh"
#include "tma_mma.cuh"

namespace pu {

namespace {

constexpr int kCdWarps = 16;
constexpr int kCdThreads = 32 * (kCdWarps + 1);
constexpr int kCdMaxStages = 4;

struct CdArgs {
  float* dw;
  float* db;  // | null
  int B, H, W, Cin, Cout;
  int npg;                   // (c,co)-chunk groups: blockIdx.y = pgroup + npg * qgroup
  int TW, TH, NB, tilesX, tilesY, tilesB, ntiles;
  int plane;                 // bytes of one staged plane (TW * TH * NB * 32, a 128-byte multiple)
  int nstages;
};

template <int NPC, int NQ>
__global__ void __launch_bounds__(kCdThreads, 1) convT2x2_dw_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                        const __grid_constant__ CUtensorMap tmy0,
                                                                        const __grid_constant__ CUtensorMap tmy1, const CdArgs a) {
  constexpr int PS = kCdWarps / NPC;
  extern __shared__ uint8_t cd_smem_raw[];
  uint8_t* smem = cd_smem_raw + ((128u - (smem_addr(cd_smem_raw) & 127u)) & 127u);
  const int stage_bytes = (2 * NPC + NQ) * a.plane;  // [parity][chunk] dY planes, then the X planes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.nstages * stage_bytes);
  const uint32_t bar0 = smem_addr(bars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pgroup = (int)blockIdx.y % a.npg, qgroup = (int)blockIdx.y / a.npg;

  if (tid == 0) {
    for (int i = 0; i < kCdMaxStages; ++i) {
      wg_mbar_init(bar0 + 8u * i, 1);
      wg_mbar_init(bar0 + 8u * (kCdMaxStages + i), kCdWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmx)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmy0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmy1)) : "memory");
  }
  __syncthreads();
  pdl_prologue();

  float acc[NQ + 1][4];
#pragma unroll
  for (int n = 0; n <= NQ; ++n)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[n][j] = 0.f;

  const int tiles_img = a.tilesX * a.tilesY;
  const bool want_bias = a.db != nullptr && qgroup == 0;
  if (warp == kCdWarps) {
    // ================= TMA producer =================
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int tb = tile / tiles_img, tr = tile - tb * tiles_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH, b0 = tb * a.NB;
      const int st = k % a.nstages;
      const uint32_t ph = (uint32_t)(k / a.nstages) & 1;
      wg_mbar_wait(bar0 + 8u * (kCdMaxStages + st), ph ^ 1);
      if (lane == 0) {
        const uint32_t full = bar0 + 8u * st;
        wg_mbar_expect_tx(full, (uint32_t)((2 * NPC + NQ) * a.plane));
        const uint32_t sS = smem_addr(smem + (size_t)st * stage_bytes);
#pragma unroll
        for (int c = 0; c < NPC; ++c) {
          const int cc0 = 8 * (pgroup * NPC + c);
          wg_tma_load_4d(sS + c * a.plane, &tmy0, full, cc0, x0, y0, b0);
          wg_tma_load_4d(sS + (NPC + c) * a.plane, &tmy1, full, cc0, x0, y0, b0);
        }
#pragma unroll
        for (int n = 0; n < NQ; ++n) wg_tma_load_4d(sS + (2 * NPC + n) * a.plane, &tmx, full, 8 * (qgroup * NQ + n), x0, y0, b0);
      }
      __syncwarp();
    }
  } else {
    // ================= MMA warps =================
    const int gq = lane >> 2, tq = lane & 3;
    const int pg = warp / PS, ps = warp - pg * PS;
    const int tws = a.TW >> 3, ths = a.TH >> 3;
    const int sx = ps % tws, sy = (ps / tws) % ths, nb = ps / (tws * ths);
    const int off = ((nb * a.TH + sy * 8) * a.TW + sx * 8 + tq) * 8 + gq;  // word offset of this lane's first element in a plane
    const int row = a.TW * 8;
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int tb = tile / tiles_img, tr = tile - tb * tiles_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int st = k % a.nstages;
      const uint32_t ph = (uint32_t)(k / a.nstages) & 1;
      wg_mbar_wait(bar0 + 8u * st, ph);
      const bool live = tx * a.TW + sx * 8 < a.W && ty * a.TH + sy * 8 < a.H && tb * a.NB + nb < a.B;
      if (live) {
        // shared-space byte addresses of this lane's first element in its two dY planes and in the X planes (tma_mma.cuh: wg_lds)
        const uint32_t sbase = smem_addr(smem) + (uint32_t)st * (uint32_t)stage_bytes + 4u * (uint32_t)off;
        uint32_t y0a = sbase + (uint32_t)(pg * a.plane), y1a = sbase + (uint32_t)((NPC + pg) * a.plane);
        uint32_t xa[NQ];
#pragma unroll
        for (int n = 0; n < NQ; ++n) xa[n] = sbase + (uint32_t)((2 * NPC + n) * a.plane);
        const uint32_t row_b = 4u * (uint32_t)row;
#pragma unroll
        for (int yy = 0; yy < 8; ++yy) {
          // A: rows 0-7 = dY_0 channels (gq), rows 8-15 = dY_1; k = pixels (tq, tq + 4)
          const unsigned a0 = wg_lds<0>(y0a), a1 = wg_lds<0>(y1a), a2 = wg_lds<128>(y0a), a3 = wg_lds<128>(y1a);
          y0a += row_b;
          y1a += row_b;
#pragma unroll
          for (int n = 0; n < NQ; ++n) {
            wg_mma(acc[n], a0, a1, a2, a3, wg_lds<0>(xa[n]), wg_lds<128>(xa[n]));
            xa[n] += row_b;
          }
          if (want_bias) wg_mma(acc[NQ], a0, a1, a2, a3, 0x3f800000u, 0x3f800000u);  // B = ones: column sums of dY
        }
      }
      __syncwarp();
      if (lane == 0) wg_mbar_arrive(bar0 + 8u * (kCdMaxStages + st));
    }
  }

  // ---- reduce the PS strips of every chunk through shared memory, one atomic per output and CTA
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);  // [NACC][kCdWarps * 32]
  if (warp < kCdWarps) {
#pragma unroll
    for (int n = 0; n <= NQ; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(n * 4 + j) * (kCdWarps * 32) + warp * 32 + lane] = acc[n][j];
  }
  __syncthreads();
  // i = ((pg * NQ + n) * 16 + r) * 8 + col: D row r = (a, channel r % 8 of chunk pg), column = input channel col of chunk n
  for (int i = tid; i < NPC * NQ * 128; i += kCdThreads) {
    const int col = i & 7, r = (i >> 3) & 15, n = (i >> 7) % NQ, pg = i / (128 * NQ);
    const int ln = (r & 7) * 4 + (col >> 1), j = ((r >> 3) << 1) | (col & 1);
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < PS; ++s) sum += red[(n * 4 + j) * (kCdWarps * 32) + (pg * PS + s) * 32 + ln];
    const int cc = 8 * (pgroup * NPC + pg) + (r & 7), par = r >> 3;
    const int c = cc / a.Cout, co = cc - c * a.Cout;
    const int ci = 8 * (qgroup * NQ + n) + col;
    atomicAdd(a.dw + (((size_t)ci * a.Cout + co) * 2 + par) * 2 + c, sum);
  }
  if (want_bias && tid < NPC * 16) {  // bias tile: every column holds the row's sum over the pixels
    const int r = tid & 15, pg = tid >> 4;
    const int ln = (r & 7) * 4, j = (r >> 3) << 1;
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < PS; ++s) sum += red[(NQ * 4 + j) * (kCdWarps * 32) + (pg * PS + s) * 32 + ln];
    const int cc = 8 * (pgroup * NPC + pg) + (r & 7);
    atomicAdd(a.db + cc % a.Cout, sum);
  }
}

template <int NPC, int NQ>
int launch_cd(const CUtensorMap& tmx, const CUtensorMap& tmy0, const CUtensorMap& tmy1, const CdArgs& ca, dim3 grid, size_t smem,
              cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(convT2x2_dw_tma_kernel<NPC, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("convT2x2_dw_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaError_t le = launch_pdl(convT2x2_dw_tma_kernel<NPC, NQ>, grid, dim3(kCdThreads), smem, st, tmx, tmy0, tmy1, ca);
  if (le != cudaSuccess) {
    set_error("convT2x2_dw_tma launch: %s", cudaGetErrorString(le));
    return PU_ERR_CUDA;
  }
  return post_launch("pu_convT2x2s2_bwd dw (tma)");
}

inline int cd_pow2_le(int v, int cap) {
  int p = cap;
  while (p > 1 && v % p != 0) p >>= 1;
  return p;
}

}  // namespace

bool convT2x2_dw_tma_ok(int Cin, int Cout) { return Cin >= 8 && Cin % 8 == 0 && Cout >= 4 && Cout % 4 == 0; }

// dw / db must be zeroed by the caller (pu_convT2x2s2_bwd does, unless PU_FLAG_ACCUM_GRADS)
int convT2x2_dw_tma(const float* x, const float* dy, float* dw, float* db, int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  const int npch = 2 * Cout / 8, nqch = Cin / 8;  // (c,co) chunks, input-channel chunks
  const int npc = cd_pow2_le(npch, 4), nq = cd_pow2_le(nqch, 4);
  const int PS = kCdWarps / npc;
  // tile = NB images x (8 ths) x (8 tws) pixels, tws * ths * NB == PS: fewest tiles (no halo), then the widest rows
  int TW = 0, TH = 0, NB = 0;
  long long best = -1;
  for (int tws = 1; tws <= PS; tws <<= 1)
    for (int ths = 1; tws * ths <= PS; ths <<= 1) {
      const int nb = PS / (tws * ths);
      if (nb > 16) continue;
      const long long tiles = (long long)cdiv(W, 8 * tws) * cdiv(H, 8 * ths) * cdiv(B, nb);
      if (best < 0 || tiles < best || (tiles == best && 8 * tws > TW)) {
        best = tiles;
        TW = 8 * tws; TH = 8 * ths; NB = nb;
      }
    }
  CdArgs ca;
  ca.dw = dw; ca.db = db; ca.B = B; ca.H = H; ca.W = W; ca.Cin = Cin; ca.Cout = Cout;
  ca.npg = npch / npc;
  ca.TW = TW; ca.TH = TH; ca.NB = NB;
  ca.tilesX = cdiv(W, TW); ca.tilesY = cdiv(H, TH); ca.tilesB = cdiv(B, NB);
  ca.ntiles = ca.tilesX * ca.tilesY * ca.tilesB;
  ca.plane = TW * TH * NB * 32;  // 64 * PS * 32: a multiple of 128
  const size_t stage = (size_t)(2 * npc + nq) * ca.plane;
  const int gy = ca.npg * (nqch / nq);
  int gx = kNumSMs / gy;
  if (gx < 1) gx = 1;
  if (gx > ca.ntiles) gx = ca.ntiles;
  const int per_cta = (ca.ntiles + gx - 1) / gx;
  int ns = (int)((size_t)218 * 1024 / stage);
  if (ns > kCdMaxStages) ns = kCdMaxStages;
  if (ns > per_cta) ns = per_cta;
  if (ns < 1) {
    set_error("pu_convT2x2s2_bwd: a stage of the TMA weight-gradient kernel does not fit in shared memory (Cin %d, Cout %d)", Cin, Cout);
    return PU_ERR_UNSUPPORTED;
  }
  ca.nstages = ns;
  size_t ring = (size_t)ns * stage;
  const size_t red = (size_t)kCdWarps * 32 * (4 * nq + 4) * sizeof(float);
  if (ring < red) ring = red;
  const size_t smem = ring + 2 * kCdMaxStages * 8 + 128;
  CUtensorMap tmx, tmy0, tmy1;
  const View xv{x, H, W, Cin, 0, 0};
  int rc = tma_make_window_map(&tmx, xv, B, H, W, 8, TW, TH, NB);
  if (rc) return rc;
  // output rows of parity a as the NHWC tensor [B, H, 2W / 2, 2 Cout]: pixel pitch 2 Cout, row pitch 2 * 2W * Cout
  const View y0v{dy, H, 2 * W, 2 * Cout, 0, 0};
  const View y1v{dy + (size_t)2 * W * Cout, H, 2 * W, 2 * Cout, 0, 0};
  rc = tma_make_window_map(&tmy0, y0v, B, H, W, 8, TW, TH, NB);
  if (rc) return rc;
  rc = tma_make_window_map(&tmy1, y1v, B, H, W, 8, TW, TH, NB);
  if (rc) return rc;
  const dim3 grid(gx, gy);
#define PU_CD_CASE(P, Q) \
  if (npc == P && nq == Q) return launch_cd<P, Q>(tmx, tmy0, tmy1, ca, grid, smem, st);
  PU_CD_CASE(1, 1) PU_CD_CASE(1, 2) PU_CD_CASE(1, 4)
  PU_CD_CASE(2, 1) PU_CD_CASE(2, 2) PU_CD_CASE(2, 4)
  PU_CD_CASE(4, 1) PU_CD_CASE(4, 2) PU_CD_CASE(4, 4)
#undef PU_CD_CASE
  set_error("convT2x2_dw_tma: no kernel for NPC=%d NQ=%d", npc, nq);
  return PU_ERR_UNSUPPORTED;
}

}  // namespace pu
