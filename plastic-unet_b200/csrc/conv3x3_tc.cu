// conv3x3_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution for sm_100a (PU_MATH_TF32).
//
//   D[pixel, co] += A[pixel, ci] * W[ci, co]   per filter tap, kind::tf32, fp32 accumulation in TMEM.
//
// Layout trick ("flattened padded rows"): the CTA's input halo tile (TH+2 rows x PW pixels, PW = tile width
// + halo rounded up to 8) is brought in by ONE 5-D TMA box per source with the tensor map dims ordered
// (4 floats of a channel group, x, y, channel group, image), so shared memory holds channel-group planes
// [cg][pixel][4 floats] and out-of-bounds coordinates zero-fill — that is the conv's zero padding and the crop
// window for free.  In that layout the A operand of output-pixel block r..r+127 for tap (ky,kx) is the
// canonical no-swizzle K-major UMMA layout starting at pixel r + ky*PW + kx: core matrices (8 pixels x 16 B)
// are contiguous (SBO = 128 B) and the two 16-byte K chunks of a K=8 MMA are one plane apart (LBO = plane
// bytes).  Every tap therefore re-reads the SAME staged tile through a shifted descriptor — the input
// crosses L2->SMEM once, not nine times — and any image width works (garbage columns x >= TW are computed
// and dropped in the epilogue).  Skip-concat is a second tensor map whose planes land behind the first.
//
// CTA = 128 threads: thread 0 issues TMA + bulk weight copy, waits the full barrier, issues all
// tcgen05.mma (one accumulator of N columns per 128-pixel block, up to 256 TMEM columns) and commits;
// the 4 warps then drain TMEM with tcgen05.ld (warp w owns lanes 32w..32w+31 = pixel rows) and apply the
// fused epilogue: bias + residual + ReLU (+ RN rounding to TF32 so that the next layer's operands are exact
// TF32 values: the MMA truncates fp32 operands, unrounded inputs would bias every product towards zero) and a
// channel-split, fully coalesced NHWC store.  Two CTAs per SM overlap one tile's epilogue with the next
// tile's loads.  All waits are bounded (trap instead of hanging the GPU).
//
// Replaces nn.Conv2d(k=3,p=1)+ReLU(+add, +cat/crop) of reference unet_p.py:105-116,161-166 and
// unet_p_res.py:150-158,186-189,215-219 for channel counts that are multiples of 8; dgrad is the same kernel
// on transposed/flipped packed weights.
#include <cuda.h>
#include <mutex>
#include "conv3x3.cuh"

namespace pu {

constexpr int kMaxChunks = 48;
constexpr int kCoBlk = 64;  // output channels per CTA (grid.z splits larger Cout)

struct TcChunk {
  int cgA, nA;  // planes [0, nA): channel groups cgA.. of source srcA
  int cgB, nB;  // planes [nA, nA+nB): channel groups cgB.. of source 1 (only in a combined chunk)
  int srcA;
  unsigned w_off;  // byte offset of this chunk's weights inside one co-block of the packed buffer
};

struct TcArgs {
  const float* wpk;
  const float* bias;
  const float* res;
  ViewW d0, d1;
  int B, H, W, Cout, relu, round_out;
  int TH, TW, PW, tilesX, tilesY;
  int nmb, nmma, plane_bytes, a_bytes, w_bytes_max, tmem_cols, nchunks;
  unsigned w_coblk_stride;  // bytes
  TcChunk chunks[kMaxChunks];
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a lost TMA transaction or MMA commit must not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  printf("pu conv3x3_tc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
  asm volatile("trap;");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// SWIZZLE_NONE K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ---- the kernel ---------------------------------------------------------------------------------
// Persistent and warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM lane
// quarter = warp % 4).  kStages shared-memory stages (A chunk planes + that chunk's weights) cycle between
// producer and MMA warp through full/empty mbarriers; two TMEM accumulator buffers cycle between the MMA warp
// and the epilogue through tmem_full/tmem_empty, so tile i's epilogue overlaps tile i+1's MMAs and tile i+2's
// loads.  Tiles are scheduled statically: tile = blockIdx.x + k * gridDim.x.
constexpr int kStages = 2;
constexpr int kTcThreads = 192;

template <int COLS>
__global__ void __launch_bounds__(kTcThreads, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm0,
                                                                   const __grid_constant__ CUtensorMap tm1, const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stage_bytes = a.a_bytes + a.w_bytes_max;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  // bars: [0,kStages) full, [kStages,2kStages) empty, then tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int st) { return bar0 + 8u * st; };
  auto empty_bar = [&](int st) { return bar0 + 8u * (kStages + st); };
  auto tfull_bar = [&](int as) { return bar0 + 8u * (2 * kStages + as); };
  auto tempty_bar = [&](int as) { return bar0 + 8u * (2 * kStages + 2 + as); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int coblk = blockIdx.y;
  const int tiles_per_img = a.tilesX * a.tilesY;
  const int ntiles = tiles_per_img * a.B;
  const int acc_cols = a.nmb * a.nmma;  // TMEM columns of one accumulator buffer

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) {
        mbar_init(full_bar(i), 1);
        mbar_init(empty_bar(i), 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(tfull_bar(i), 1);
        mbar_init(tempty_bar(i), 4);  // one arrival per epilogue warp
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    const uint8_t* wblk = reinterpret_cast<const uint8_t*>(a.wpk) + (size_t)coblk * a.w_coblk_stride;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH;
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const int st = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(empty_bar(st), ph ^ 1);  // passes immediately on a fresh barrier
        if (elect_one()) {
          const TcChunk ch = a.chunks[c];
          const int ncg = ch.nA + ch.nB;
          const uint32_t w_bytes = (uint32_t)(9 * ncg * a.nmma * 16);
          const uint32_t sA = smem_u32(smem + st * stage_bytes);
          mbar_expect_tx(full_bar(st), (uint32_t)(ncg * a.plane_bytes) + w_bytes);
          tma_load_5d(sA, ch.srcA == 0 ? &tm0 : &tm1, full_bar(st), 0, x0 - 1, y0 - 1, ch.cgA, b);
          if (ch.nB > 0) tma_load_5d(sA + ch.nA * a.plane_bytes, &tm1, full_bar(st), 0, x0 - 1, y0 - 1, ch.cgB, b);
          bulk_load(sA + a.a_bytes, wblk + ch.w_off, w_bytes, full_bar(st));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp runs the warp-uniform loop and one elected lane executes each tcgen05 instruction, so the
    // descriptors live in uniform registers.  M blocks are innermost: consecutive MMAs write different
    // accumulators and are not serialised on the accumulate dependency of one small TMEM tile.
    // instruction descriptor: D=f32, A=B=tf32, K-major both, N = nmma, M = 128 (cute::UMMA::InstrDescriptor)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.nmma >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_kstep = (uint32_t)(2 * a.plane_bytes) >> 4;  // two channel-group planes per K=8 step (16-byte units)
    const uint32_t b_kstep = (uint32_t)(2 * a.nmma);
    const bool leader = elect_one();  // elected once: the issue loop must stay a handful of instructions per MMA
    const uint32_t nmma = (uint32_t)a.nmma;
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(tempty_bar(as), aph ^ 1);  // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(as * acc_cols);
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const int st = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        const int ncg = a.chunks[c].nA + a.chunks[c].nB;
        mbar_wait(full_bar(st), ph);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + st * stage_bytes);
        const uint64_t a_base = umma_desc(sA, (uint32_t)a.plane_bytes, 128);
        const uint64_t b_base = umma_desc(sA + a.a_bytes, (uint32_t)(a.nmma * 16), 128);
        const int ksteps = ncg >> 1;
        const uint32_t b_tap = (uint32_t)(ncg * a.nmma);
        uint32_t b_off = 0;
        for (int ky = 0; ky < 3; ++ky) {
          for (int kx = 0; kx < 3; ++kx) {
            const uint32_t a_tap = (uint32_t)(ky * a.PW + kx);
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t bd = b_base + (uint64_t)(b_off + ks * b_kstep);
              uint64_t ad = a_base + (uint64_t)(a_tap + ks * a_kstep);
              uint32_t d = d0;
              const uint32_t acc = (c | ky | kx | ks) ? 1u : 0u;
#pragma unroll 4
              for (int mb = 0; mb < a.nmb; ++mb) {
                if (leader) umma_tf32(d, ad, bd, idesc, acc);
                d += nmma;
                ad += 128;  // next 128-pixel block: 128 * 16 B, in 16-byte descriptor units
              }
            }
            b_off += b_tap;
          }
        }
        if (leader) tc_commit(empty_bar(st));  // frees the stage once these MMAs have read it
        __syncwarp();
      }
      if (leader) tc_commit(tfull_bar(as));  // accumulator complete
      __syncwarp();
    }
  } else {
    // ================= epilogue: TMEM -> registers -> bias/residual/ReLU -> NHWC global =================
    const int quarter = warp & 3;           // TMEM lanes [32*quarter, 32*quarter+32)
    const int row = quarter * 32 + lane;    // row of the 128-pixel block owned by this thread
    const int co_base = coblk * kCoBlk;
    float bv[COLS];
#pragma unroll
    for (int j = 0; j < COLS; ++j) bv[j] = (a.bias != nullptr && co_base + j < a.Cout) ? __ldg(a.bias + co_base + j) : 0.f;
    const int step_y = 128 / a.PW, step_x = 128 - step_y * a.PW;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH;
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      int yy = row / a.PW, xx = row - yy * a.PW;
      for (int mb = 0; mb < a.nmb; ++mb) {
        uint32_t v[COLS];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * acc_cols + mb * a.nmma);
        if (COLS == 8) {
          tmem_ld8(taddr, v);
        } else {
#pragma unroll
          for (int q = 0; q < COLS / 16; ++q) tmem_ld16(taddr + 16 * q, v + 16 * q);
        }
        tmem_ld_wait();
        const int gy = y0 + yy, gx = x0 + xx;
        if (yy < a.TH && xx < a.TW && gy < a.H && gx < a.W) {
          const float* rp = a.res != nullptr ? a.res + (((size_t)b * a.H + gy) * a.W + gx) * a.Cout + co_base : nullptr;
#pragma unroll
          for (int q = 0; q < COLS / 4; ++q) {
            const int co = co_base + 4 * q;
            if (co < a.Cout) {
              float o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = __uint_as_float(v[4 * q + j]) + bv[4 * q + j];
              if (rp != nullptr) {
                const float4 rr = ldg4(rp + 4 * q);
                o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
              }
              if (a.relu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
              }
              if (a.round_out) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = round_tf32(o[j]);
              }
              const bool first = co < a.d0.C;
              const ViewW dd = first ? a.d0 : a.d1;
              const int cd = first ? co : co - a.d0.C;
              float* dp = dd.p + (((size_t)b * dd.Hs + (gy + dd.oy)) * dd.Ws + (gx + dd.ox)) * dd.C + cd;
              *reinterpret_cast<float4*>(dp) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        }
        // next 128-pixel block: advance (yy, xx) without a division
        yy += step_y;
        xx += step_x;
        if (xx >= a.PW) { xx -= a.PW; ++yy; }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(as)) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
}

// ---- weight packing for the tensor-core path ------------------------------------------------------
// out[coblk][chunk][tap][plane][n][4]  (n < nmma rows, zero beyond the co-block's valid channels), RN-rounded to TF32
struct TcPackArgs {
  const float* w;  // OIHW [Cout_w][Cin_w][3][3]
  float* out;
  int Cout_w, Cin_w, transpose;  // transpose: conv computes dgrad (input channels = Cout_w, output = Cin_w, taps flipped)
  int Cin, Cout, C0, nmma, nchunks, ncoblk;
  unsigned w_coblk_stride;
  TcChunk chunks[kMaxChunks];
};

__global__ void pack_w3x3_tc_kernel(const TcPackArgs a) {
  const int coblk = blockIdx.y, c = blockIdx.z;
  const TcChunk ch = a.chunks[c];
  const int ncg = ch.nA + ch.nB;
  const int n_el = 9 * ncg * a.nmma * 4;
  float* out = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.out) + (size_t)coblk * a.w_coblk_stride + ch.w_off);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += gridDim.x * blockDim.x) {
    const int j = i & 3;
    const int n = (i >> 2) % a.nmma;
    const int pl = (i / (4 * a.nmma)) % ncg;
    const int tap = i / (4 * a.nmma * ncg);
    int ci;  // channel in the concatenated input
    if (pl < ch.nA) ci = (ch.srcA == 0 ? 0 : a.C0) + (ch.cgA + pl) * 4 + j;
    else ci = a.C0 + (ch.cgB + pl - ch.nA) * 4 + j;
    const int co = coblk * kCoBlk + n;
    float v = 0.f;
    if (co < a.Cout && n < kCoBlk && ci < a.Cin) {
      if (!a.transpose) v = a.w[((size_t)co * a.Cin_w + ci) * 9 + tap];
      else v = a.w[((size_t)ci * a.Cin_w + co) * 9 + (8 - tap)];
    }
    out[i] = round_tf32(v);
  }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static int g_tc_state = -1;  // -1 unknown, 0 unavailable, 1 available
static std::mutex g_tc_mu;

static bool tc_init() {
  std::lock_guard<std::mutex> lk(g_tc_mu);
  if (g_tc_state >= 0) return g_tc_state == 1;
  g_tc_state = 0;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (prop.major != 10) return false;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
      qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  g_tc_state = 1;
  return true;
}

struct TcPlan {
  int TH, TW, PW, tilesX, tilesY, nmb, nmma, cols, plane_bytes, a_bytes, w_bytes_max, tmem_cols, nchunks, ncoblk, kcg0, kcg1;
  unsigned w_coblk_stride;
  size_t smem_bytes;
  TcChunk chunks[kMaxChunks];
};

static int next_pow2_cols(int c) {
  int p = 32;
  while (p < c) p <<= 1;
  return p;
}

// returns false if the shape does not fit the tensor-core path
static bool tc_plan(int B, int H, int W, int C0, int C1, int Cout, TcPlan* p) {
  if (C0 < 8 || C0 % 8 != 0 || C1 % 8 != 0 || Cout % 8 != 0) return false;
  const int cg0 = C0 / 4, cg1 = C1 / 4;
  // chunks of at most 8 channel groups (32 channels); small concat pairs share one chunk
  p->nchunks = 0;
  p->kcg0 = cg0 < 8 ? cg0 : 8;
  p->kcg1 = cg1 == 0 ? 0 : (cg1 < 8 ? cg1 : 8);
  if (cg0 % p->kcg0 != 0 || (cg1 > 0 && cg1 % p->kcg1 != 0)) return false;
  const int cout_blk = Cout < kCoBlk ? Cout : kCoBlk;
  p->cols = cout_blk <= 8 ? 8 : (cout_blk <= 16 ? 16 : (cout_blk <= 32 ? 32 : 64));
  p->nmma = p->cols < 16 ? 16 : p->cols;
  p->ncoblk = (Cout + kCoBlk - 1) / kCoBlk;
  unsigned woff = 0;
  int max_ncg = 0;
  auto add = [&](int srcA, int cgA, int nA, int cgB, int nB) {
    TcChunk& ch = p->chunks[p->nchunks++];
    ch.srcA = srcA; ch.cgA = cgA; ch.nA = nA; ch.cgB = cgB; ch.nB = nB; ch.w_off = woff;
    woff += (unsigned)(9 * (nA + nB) * p->nmma * 16);
    if (nA + nB > max_ncg) max_ncg = nA + nB;
  };
  if (cg1 > 0 && cg0 + cg1 <= 8) {
    add(0, 0, cg0, 0, cg1);
  } else {
    if (cg0 / p->kcg0 + (cg1 ? cg1 / p->kcg1 : 0) > kMaxChunks) return false;
    for (int c = 0; c < cg0; c += p->kcg0) add(0, c, p->kcg0, 0, 0);
    for (int c = 0; c < cg1; c += p->kcg1) add(1, c, p->kcg1, 0, 0);
  }
  p->w_coblk_stride = woff;
  p->w_bytes_max = ((9 * max_ncg * p->nmma * 16) + 127) / 128 * 128;
  // tile geometry
  p->tilesX = (W + 247) / 248;
  p->TW = (W + p->tilesX - 1) / p->tilesX;
  p->PW = (p->TW + 2 + 7) / 8 * 8;
  if (p->PW > 256) return false;
  // one CTA per SM: kStages stages of (A planes + weights) and two TMEM accumulator buffers of <= 256 columns
  const size_t smem_soft = (216 * 1024) / kStages, smem_hard = (216 * 1024) / kStages;
  int best = 0;
  for (int pass = 0; pass < 2 && best == 0; ++pass) {
    const size_t lim = pass == 0 ? smem_soft : smem_hard;
    for (int th = (H < 48 ? H : 48); th >= 1; --th) {
      const int nmb = (th * p->PW + 127) / 128;
      if (nmb * p->nmma > 256) continue;
      const int plane_pix = (th + 2) * p->PW;
      int tail = nmb * 128 + 2 * p->PW + 2 - plane_pix;
      if (tail < 0) tail = 0;
      const size_t a_bytes = ((size_t)max_ncg * plane_pix * 16 + (size_t)tail * 16 + 127) / 128 * 128;
      const size_t total = a_bytes + p->w_bytes_max;
      if (total > lim) continue;
      // prefer the largest tile that still gives every SM at least two tiles; never shrink below two M blocks
      const long long tiles = (long long)B * ((H + th - 1) / th) * p->tilesX;
      best = th;
      if (tiles >= 2LL * kNumSMs || th * p->PW <= 256) break;
    }
  }
  if (best == 0) return false;
  p->TH = best;
  p->tilesY = (H + best - 1) / best;
  p->nmb = (best * p->PW + 127) / 128;
  p->plane_bytes = (best + 2) * p->PW * 16;
  int tail = p->nmb * 128 + 2 * p->PW + 2 - (best + 2) * p->PW;
  if (tail < 0) tail = 0;
  p->a_bytes = (int)(((size_t)max_ncg * p->plane_bytes + (size_t)tail * 16 + 127) / 128 * 128);
  p->tmem_cols = next_pow2_cols(2 * p->nmb * p->nmma);
  p->smem_bytes = (size_t)kStages * ((size_t)p->a_bytes + p->w_bytes_max) + 128;
  return true;
}

static int make_tmap(CUtensorMap* tm, const View& v, int B, int H, int W, int PW, int TH, int kcg) {
  // dims, innermost first: (4 floats, x, y, channel group, image) over the HxW window of the tensor
  const float* base = v.p + ((size_t)v.oy * v.Ws + v.ox) * v.C;
  cuuint64_t dims[5] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(v.C / 4), (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)v.C * 4, (cuuint64_t)v.Ws * v.C * 4, 16, (cuuint64_t)v.Hs * v.Ws * v.C * 4};
  cuuint32_t box[5] = {4, (cuuint32_t)PW, (cuuint32_t)(TH + 2), (cuuint32_t)kcg, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv3x3_tc: cuTensorMapEncodeTiled failed (CUresult %d) for window %dx%d C=%d box (4,%d,%d,%d,1)", (int)r, H, W, v.C,
              PW, TH + 2, kcg);
    return PU_ERR_CUDA;
  }
  return PU_OK;
}

bool conv3x3_tc_ok(int C0, int C1, int Cout, int Cd0, int Cd1) {
  if (!tc_init()) return false;
  TcPlan p;
  if (!tc_plan(1, 8, 8, C0, C1, Cout, &p)) return false;
  if (Cd0 % 4 != 0 || Cd1 % 4 != 0 || Cd0 + Cd1 != Cout) return false;
  return true;
}

long long conv3x3_tc_weight_floats(int C0, int C1, int Cout) {
  TcPlan p;
  if (!tc_plan(1, 8, 8, C0, C1, Cout, &p)) return 0;
  return (long long)p.w_coblk_stride * p.ncoblk / 4;
}

int conv3x3_tc_pack(const float* w_oihw, float* out, int Cout_w, int Cin_w, int transpose, int C0, cudaStream_t st) {
  // roles of the conv that will consume the packed weights
  const int cin = transpose ? Cout_w : Cin_w;
  const int cout = transpose ? Cin_w : Cout_w;
  const int c0 = transpose ? Cout_w : C0;
  TcPlan p;
  if (!tc_plan(1, 8, 8, c0, cin - c0, cout, &p)) {
    set_error("pu_pack_w3x3: channels (%d|%d -> %d) do not fit the tcgen05 path", c0, cin - c0, cout);
    return PU_ERR_UNSUPPORTED;
  }
  TcPackArgs pa;
  pa.w = w_oihw; pa.out = out; pa.Cout_w = Cout_w; pa.Cin_w = Cin_w; pa.transpose = transpose;
  pa.Cin = cin; pa.Cout = cout; pa.C0 = c0; pa.nmma = p.nmma; pa.nchunks = p.nchunks; pa.ncoblk = p.ncoblk;
  pa.w_coblk_stride = p.w_coblk_stride;
  for (int i = 0; i < p.nchunks; ++i) pa.chunks[i] = p.chunks[i];
  dim3 g(4, p.ncoblk, p.nchunks);
  pack_w3x3_tc_kernel<<<g, 256, 0, st>>>(pa);
  return post_launch("pu_pack_w3x3 (tc)");
}

template <int COLS>
static int launch_tc(const CUtensorMap& tm0, const CUtensorMap& tm1, const TcArgs& ta, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) {
      set_error("conv3x3_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
    attr_set = true;
  }
  conv3x3_tc_kernel<COLS><<<grid, kTcThreads, smem, st>>>(tm0, tm1, ta);
  return post_launch("conv3x3_tc");
}

int conv3x3_fwd_tc(const Conv3x3Args& a, cudaStream_t st) {
  if (!tc_init()) {
    set_error("pu_conv3x3_fwd: PU_MATH_TF32 requested but the tcgen05/TMA path is unavailable on this device");
    return PU_ERR_UNSUPPORTED;
  }
  TcPlan p;
  const int C1 = (a.s1.p != nullptr) ? a.s1.C : 0;
  if (!tc_plan(a.B, a.H, a.W, a.s0.C, C1, a.Cout, &p) || a.d0.C % 4 != 0 || (a.d1.p != nullptr && a.d1.C % 4 != 0) || a.B > 65535) {
    set_error("pu_conv3x3_fwd: PU_MATH_TF32 does not support this shape (C %d|%d -> %d, %dx%d); check pu_conv3x3_tc_ok", a.s0.C, C1,
              a.Cout, a.H, a.W);
    return PU_ERR_UNSUPPORTED;
  }
  CUtensorMap tm0, tm1;
  int rc = make_tmap(&tm0, a.s0, a.B, a.H, a.W, p.PW, p.TH, p.kcg0);
  if (rc) return rc;
  if (C1 > 0) {
    rc = make_tmap(&tm1, a.s1, a.B, a.H, a.W, p.PW, p.TH, p.kcg1);
    if (rc) return rc;
  } else {
    tm1 = tm0;
  }
  TcArgs ta;
  ta.wpk = a.wp; ta.bias = a.bias; ta.res = a.res; ta.d0 = a.d0; ta.d1 = a.d1;
  ta.B = a.B; ta.H = a.H; ta.W = a.W; ta.Cout = a.Cout; ta.relu = a.relu; ta.round_out = a.round_out;
  ta.TH = p.TH; ta.TW = p.TW; ta.PW = p.PW; ta.tilesX = p.tilesX; ta.tilesY = p.tilesY;
  ta.nmb = p.nmb; ta.nmma = p.nmma; ta.plane_bytes = p.plane_bytes; ta.a_bytes = p.a_bytes; ta.w_bytes_max = p.w_bytes_max;
  ta.tmem_cols = p.tmem_cols; ta.nchunks = p.nchunks; ta.w_coblk_stride = p.w_coblk_stride;
  for (int i = 0; i < p.nchunks; ++i) ta.chunks[i] = p.chunks[i];
  const int ntiles = p.tilesX * p.tilesY * a.B;
  dim3 grid(ntiles < kNumSMs ? ntiles : kNumSMs, p.ncoblk);
  switch (p.cols) {
    case 8: return launch_tc<8>(tm0, tm1, ta, grid, p.smem_bytes, st);
    case 16: return launch_tc<16>(tm0, tm1, ta, grid, p.smem_bytes, st);
    case 32: return launch_tc<32>(tm0, tm1, ta, grid, p.smem_bytes, st);
    default: return launch_tc<64>(tm0, tm1, ta, grid, p.smem_bytes, st);
  }
}

}  // namespace pu

extern "C" int pu_tc_available(void) { return pu::tc_init() ? 1 : 0; }
