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
#include <stdlib.h>
#include <mutex>
#include "conv3x3.cuh"

namespace pu {

constexpr int kMaxChunks = 48;
constexpr int kCoBlk = 64;  // output channels per CTA (grid.y splits larger Cout)
constexpr unsigned kMaxResidentW = 40 * 1024;  // largest weight image kept resident in shared memory (per co block)

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
  const float* mask0;  // | null: geometry of d0; the stored value is zeroed where mask0 <= 0
  const float* mask1;
  int B, H, W, Cout, relu, round_out;
  int TH, TW, PW, tilesX, tilesY;
  int nmb, nmma, plane_bytes, a_bytes, w_bytes_max, tmem_cols, nchunks;
  int wfmt;       // 0: wpk holds the packed B tiles (streamed per stage); 1/2: wpk is the raw OIHW weight (forward / dgrad)
                  // and the B tiles are built in shared memory once per CTA (resident)
  int Cin, C0;    // concatenated input channels and the split point (for the in-kernel weight build)
  int w_res_bytes;  // bytes of the resident weight image (0 in streamed mode)
  View s0, s1;      // the two sources (cp.async loader path)
  int loader;       // 1: one 5-D TMA box per source (default); 0: 4 loader warps with 16-byte cp.async
  int debug;  // PU_TC_DEBUG experiments: 1 = skip MMAs, 2 = skip epilogue stores, 4 = load only the first stage
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
__device__ __forceinline__ void ldg8(const float* p, float* r) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg8(float* p, const float* r) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]),
               "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ---- the kernel ---------------------------------------------------------------------------------
// Persistent and warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM lane
// quarter = warp % 4).  kStages shared-memory stages (A chunk planes + that chunk's weights) cycle between
// producer and MMA warp through full/empty mbarriers; two TMEM accumulator buffers cycle between the MMA warp
// and the epilogue through tmem_full/tmem_empty, so tile i's epilogue overlaps tile i+1's MMAs and tile i+2's
// loads.  Tiles are scheduled statically: tile = blockIdx.x + k * gridDim.x.
constexpr int kStages = 2;
constexpr int kLoadWarps = 4;                       // loader warps (cp.async path); warp 0 alone drives the TMA path
constexpr int kMmaWarps = 4;                        // MMA-issuing warps: warp m owns the 128-pixel blocks mb = m (mod 4)
constexpr int kEpiWarps = 8;                        // two warps per TMEM lane quarter, alternating 128-pixel blocks
constexpr int kMmaWarp0 = kLoadWarps;
constexpr int kEpiWarp0 = kLoadWarps + kMmaWarps;   // first epilogue warp (multiple of 4: quarter = warp & 3)
constexpr int kTcThreads = 32 * (kLoadWarps + kMmaWarps + kEpiWarps);

template <int COLS>
__global__ void __launch_bounds__(kTcThreads, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm0,
                                                                   const __grid_constant__ CUtensorMap tm1, const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int stage_bytes = a.a_bytes + a.w_bytes_max;  // w_bytes_max == 0 when the weights are resident
  uint8_t* smWres = smem + kStages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smWres + a.w_res_bytes);
  // bars: [0,kStages) full, [kStages,2kStages) empty, then tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int st) { return bar0 + 8u * st; };
  auto empty_bar = [&](int st) { return bar0 + 8u * (kStages + st); };
  auto tfull_bar = [&](int as) { return bar0 + 8u * (2 * kStages + as); };
  auto tempty_bar = [&](int as) { return bar0 + 8u * (2 * kStages + 2 + as); };

  pdl_prologue();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int coblk = blockIdx.y;
  const int tiles_per_img = a.tilesX * a.tilesY;
  const int ntiles = tiles_per_img * a.B;
  const int acc_cols = a.nmb * a.nmma;  // TMEM columns of one accumulator buffer

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) {
        mbar_init(full_bar(i), a.loader == 0 ? 32 * kLoadWarps : 1);
        mbar_init(empty_bar(i), kMmaWarps);  // one tcgen05.commit per MMA warp
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(tfull_bar(i), kMmaWarps);
        mbar_init(tempty_bar(i), kEpiWarps);  // one arrival per epilogue warp
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (a.wfmt != 0) {
    // Build the tf32 B-operand tiles of every K chunk straight from the OIHW weight tensor (no pack kernel):
    // image layout [chunk][tap][channel-group plane][N rows][4], identical to pu_pack_w3x3's.
    const float* __restrict__ w = a.wpk;
    const int co_base_w = coblk * kCoBlk;
    for (int c = 0; c < a.nchunks; ++c) {
      const TcChunk ch = a.chunks[c];
      const int ncg = ch.nA + ch.nB;
      const int units = ncg * a.nmma;  // 16-byte units per tap: [plane][n]
      float4* out = reinterpret_cast<float4*>(smWres + ch.w_off);
      const int nsh = 31 - __clz(a.nmma);  // nmma is 16, 32 or 64
      for (int u = tid; u < units; u += kTcThreads) {
        const int n = u & (a.nmma - 1), pl = u >> nsh;
        int ci0;
        if (pl < ch.nA) ci0 = (ch.srcA == 0 ? 0 : a.C0) + (ch.cgA + pl) * 4;
        else ci0 = a.C0 + (ch.cgB + pl - ch.nA) * 4;
        const int co = co_base_w + n;
        const bool ok = co < a.Cout && n < kCoBlk;
        // forward: w[co][ci][tap]; dgrad: w[ci][co][8 - tap]  (ci = conv input channel, co = conv output channel)
        const float* base = a.wfmt == 1 ? w + ((size_t)co * a.Cin + ci0) * 9 : w + ((size_t)ci0 * a.Cout + co) * 9;
        const size_t cstride = a.wfmt == 1 ? 9 : (size_t)a.Cout * 9;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int t = a.wfmt == 1 ? tap : 8 - tap;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) {
            v.x = round_tf32(__ldg(base + t));
            v.y = round_tf32(__ldg(base + cstride + t));
            v.z = round_tf32(__ldg(base + 2 * cstride + t));
            v.w = round_tf32(__ldg(base + 3 * cstride + t));
          }
          out[tap * units + u] = v;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kLoadWarps && a.loader == 0) {
    // ================= cp.async loaders (4 warps) =================
    // 16-byte LDGSTS per thread: consecutive threads read consecutive 16 bytes of an NHWC halo row (fully coalesced,
    // every 32-byte sector used once) and scatter them into the channel-group planes; out-of-image pixels are
    // zero-filled by src-size 0.  Measured on B200 this path is SLOWER than the 5-D TMA box (43.6 vs 25.4 us for
    // 8->8 @128x128 B=64: the per-copy index arithmetic of 128 threads costs more than TMA's 16-byte rows), so it
    // is only kept as an alternative loader (PU_TC_LOADER=cpasync).
    const int ltid = tid;  // 0..127
    const uint8_t* wblk = reinterpret_cast<const uint8_t*>(a.wpk) + (size_t)coblk * a.w_coblk_stride;
    const int halo_rows = a.TH + 2;
    uint32_t it = 0;
    int prev_st = -1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH;
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const int st = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(empty_bar(st), ph ^ 1);
        const TcChunk ch = a.chunks[c];
        const uint32_t sA = smem_u32(smem + st * stage_bytes);
        if (!((a.debug & 4) && it >= (uint32_t)kStages)) {
#pragma unroll 1
          for (int part = 0; part < 2; ++part) {
            const int nX = part == 0 ? ch.nA : ch.nB;
            if (nX == 0) continue;
            const View v = (part == 0 && ch.srcA == 0) ? a.s0 : a.s1;
            const int cgX = part == 0 ? ch.cgA : ch.cgB;
            const uint32_t dplane = sA + (uint32_t)((part == 0 ? 0 : ch.nA) * a.plane_bytes);
            const int units = a.PW * nX;  // 16-byte units per halo row
            const bool pow2 = (nX & (nX - 1)) == 0;
            const int sh = 31 - __clz(nX);
            for (int hy = 0; hy < halo_rows; ++hy) {
              const int gy = y0 - 1 + hy;
              const bool rowok = gy >= 0 && gy < a.H;
              const float* rowp = v.p + (((size_t)b * v.Hs + (rowok ? gy + v.oy : 0)) * v.Ws + v.ox) * v.C + cgX * 4;
              const uint32_t drow = dplane + (uint32_t)(hy * a.PW * 16);
              for (int u = ltid; u < units; u += 32 * kLoadWarps) {
                const int hx = pow2 ? (u >> sh) : (u / nX);
                const int cgl = u - hx * nX;
                const int gx = x0 - 1 + hx;
                const bool ok = rowok && gx >= 0 && gx < a.W;
                const float* src = ok ? rowp + (size_t)gx * v.C + cgl * 4 : v.p;
                const uint32_t dst = drow + (uint32_t)(cgl * a.plane_bytes + hx * 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
              }
            }
          }
          if (a.wfmt == 0) {  // streamed weights: plain 16-byte cp.async as well
            const int ncg = ch.nA + ch.nB;
            const int wunits = 9 * ncg * a.nmma;
            const uint8_t* wsrc = wblk + ch.w_off;
            for (int u = ltid; u < wunits; u += 32 * kLoadWarps)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(sA + a.a_bytes + u * 16), "l"(wsrc + (size_t)u * 16) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (prev_st >= 0) {  // the previous stage's copies have landed: publish it to the MMA warps
          asm volatile("cp.async.wait_group 1;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_bar(prev_st)) : "memory");
        }
        prev_st = st;
      }
    }
    if (prev_st >= 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_bar(prev_st)) : "memory");
    }
  } else if (warp == 0) {
    // ================= TMA producer (loader == 1) =================
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
          if ((a.debug & 4) && it >= (uint32_t)kStages) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_bar(st)) : "memory");
          } else {
          mbar_expect_tx(full_bar(st), (uint32_t)(ncg * a.plane_bytes) + (a.wfmt == 0 ? w_bytes : 0u));
          tma_load_5d(sA, ch.srcA == 0 ? &tm0 : &tm1, full_bar(st), 0, x0 - 1, y0 - 1, ch.cgA, b);
          if (ch.nB > 0) tma_load_5d(sA + ch.nA * a.plane_bytes, &tm1, full_bar(st), 0, x0 - 1, y0 - 1, ch.cgB, b);
          if (a.wfmt == 0) bulk_load(sA + a.a_bytes, wblk + ch.w_off, w_bytes, full_bar(st));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= kMmaWarp0 && warp < kEpiWarp0) {
    // ================= MMA issuers =================
    // A tcgen05.mma of this conv is tiny (N = 16..64 columns, 8-32 tensor-pipe cycles) and a single warp needs
    // ~30 cycles of scalar work per issue, so the 128-pixel blocks are dealt round-robin to kMmaWarps issuing warps.
    const int mw = warp - kMmaWarp0;
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
        const uint64_t b_base = umma_desc(a.wfmt == 0 ? sA + a.a_bytes : smem_u32(smWres) + a.chunks[c].w_off, (uint32_t)(a.nmma * 16), 128);
        const int ksteps = ncg >> 1;
        const uint32_t b_tap = (uint32_t)(ncg * a.nmma);
        uint32_t b_off = 0;
        for (int ky = 0; ky < 3; ++ky) {
          for (int kx = 0; kx < 3; ++kx) {
            const uint32_t a_tap = (uint32_t)(ky * a.PW + kx);
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t bd = b_base + (uint64_t)(b_off + ks * b_kstep);
              uint64_t ad = a_base + (uint64_t)(a_tap + ks * a_kstep + mw * 128);
              uint32_t d = d0 + mw * nmma;
              const uint32_t acc = (c | ky | kx | ks) ? 1u : 0u;
#pragma unroll 4
              for (int mb = mw; mb < a.nmb; mb += kMmaWarps) {
                if (leader && !(a.debug & 1)) umma_tf32(d, ad, bd, idesc, acc);
                d += kMmaWarps * nmma;
                ad += kMmaWarps * 128;  // this warp's next 128-pixel block: 128 * 16 B each, in 16-byte descriptor units
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
  } else if (warp >= kEpiWarp0) {
    // ================= epilogue: TMEM -> registers -> bias/residual/ReLU -> NHWC global =================
    // 8 warps: warp e handles TMEM lane quarter (warp % 4) of the 128-pixel blocks mb = set, set+2, ... (set = e / 4);
    // G blocks are fetched per tcgen05.wait::ld so that the TMEM read latency is paid once per group.
    constexpr int G = COLS <= 8 ? 4 : (COLS == 16 ? 2 : 1);  // G * COLS = 32 accumulator + 32 aux registers
    const int quarter = warp & 3;
    const int set = (warp - kEpiWarp0) >> 2;
    const int row = quarter * 32 + lane;
    const int co_base = coblk * kCoBlk;
    float bv[COLS];
#pragma unroll
    for (int j = 0; j < COLS; ++j) bv[j] = (a.bias != nullptr && co_base + j < a.Cout) ? __ldg(a.bias + co_base + j) : 0.f;
    const int step_y = 256 / a.PW, step_x = 256 - step_y * a.PW;  // this warp advances two 128-pixel blocks at a time
    // 256-bit stores need 32-byte aligned channel groups in both destinations
    const bool vec8 = (a.d0.C % 8 == 0) && (a.d1.p == nullptr || a.d1.C % 8 == 0);
    const int row0 = set * 128 + row;
    const int yy0 = row0 / a.PW, xx0 = row0 - yy0 * a.PW;
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
      int yy = yy0, xx = xx0;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * acc_cols);
      // `aux` = the one extra tensor the epilogue reads per element: the dgrad ReLU masks if present, else the residual.
      // Its loads are issued BEFORE the TMEM wait of a group so that their latency overlaps the tcgen05.ld's.
      const bool has_mask = a.mask0 != nullptr || a.mask1 != nullptr;
      constexpr bool kPrefetch = COLS <= 32;
      for (int mb0 = set; mb0 < a.nmb; mb0 += 2 * G) {
        uint32_t v[G][COLS];
        float aux[kPrefetch ? G : 1][kPrefetch ? COLS : 1];
        int gyv[G], gxv[G];
        bool okv[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int mb = mb0 + 2 * g;
          gyv[g] = y0 + yy;
          gxv[g] = x0 + xx;
          okv[g] = mb < a.nmb && yy < a.TH && xx < a.TW && gyv[g] < a.H && gxv[g] < a.W && !(a.debug & 2);
          // next block of this warp (two 128-pixel blocks further): advance (yy, xx) without a division
          yy += step_y;
          xx += step_x;
          if (xx >= a.PW) { xx -= a.PW; ++yy; }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int mb = mb0 + 2 * g;
          if (mb < a.nmb) {  // warp-uniform
            const uint32_t taddr = tbase + (uint32_t)(mb * a.nmma);
            if (COLS == 8) {
              tmem_ld8(taddr, v[g]);
            } else {
#pragma unroll
              for (int q = 0; q < COLS / 16; ++q) tmem_ld16(taddr + 16 * q, v[g] + 16 * q);
            }
          }
        }
        if (kPrefetch && vec8 && (has_mask || a.res != nullptr)) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (!okv[g]) continue;
#pragma unroll
            for (int q = 0; q < COLS / 8; ++q) {
              const int co = co_base + 8 * q;
              if (co >= a.Cout) continue;
              const float* src;
              if (has_mask) {
                const bool first = co < a.d0.C;
                const ViewW dd = first ? a.d0 : a.d1;
                const float* mk = first ? a.mask0 : a.mask1;
                src = mk == nullptr ? nullptr
                                    : mk + (((size_t)b * dd.Hs + (gyv[g] + dd.oy)) * dd.Ws + (gxv[g] + dd.ox)) * dd.C + (first ? co : co - a.d0.C);
              } else {
                src = a.res + (((size_t)b * a.H + gyv[g]) * a.W + gxv[g]) * a.Cout + co;
              }
              if (src != nullptr) {
                ldg8(src, &aux[kPrefetch ? g : 0][kPrefetch ? 8 * q : 0]);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) aux[kPrefetch ? g : 0][kPrefetch ? 8 * q + j : 0] = 1.f;  // "mask" that keeps everything
              }
            }
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (!okv[g]) continue;
          const int gy = gyv[g], gx = gxv[g];
          const float* rp = a.res != nullptr ? a.res + (((size_t)b * a.H + gy) * a.W + gx) * a.Cout + co_base : nullptr;
#pragma unroll
          for (int q = 0; q < COLS / 8; ++q) {  // 8 channels = one 32-byte sector per 256-bit access
            const int co = co_base + 8 * q;
            if (co >= a.Cout) continue;
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[g][8 * q + j]) + bv[8 * q + j];
            if (rp != nullptr) {
              if (kPrefetch && vec8 && !has_mask) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] += aux[kPrefetch ? g : 0][kPrefetch ? 8 * q + j : 0];
              } else {
                float rr[8];
                ldg8(rp + 8 * q, rr);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] += rr[j];
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
            // an 8-channel group may straddle the d0|d1 split only at a multiple of 4
            if (vec8) {
              const bool first = co < a.d0.C;
              const ViewW dd = first ? a.d0 : a.d1;
              const int cd = first ? co : co - a.d0.C;
              const size_t doff = (((size_t)b * dd.Hs + (gy + dd.oy)) * dd.Ws + (gx + dd.ox)) * dd.C + cd;
              if (has_mask) {  // dgrad: ReLU mask of the layer that produced this source
                if (kPrefetch) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) o[j] = aux[kPrefetch ? g : 0][kPrefetch ? 8 * q + j : 0] > 0.f ? o[j] : 0.f;
                } else {
                  const float* mk = first ? a.mask0 : a.mask1;
                  if (mk != nullptr) {
                    float mv[8];
                    ldg8(mk + doff, mv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = mv[j] > 0.f ? o[j] : 0.f;
                  }
                }
              }
              stg8(dd.p + doff, o);
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int c4 = co + 4 * h;
                const bool first = c4 < a.d0.C;
                const ViewW dd = first ? a.d0 : a.d1;
                const int cd = first ? c4 : c4 - a.d0.C;
                const size_t doff = (((size_t)b * dd.Hs + (gy + dd.oy)) * dd.Ws + (gx + dd.ox)) * dd.C + cd;
                const float* mk = first ? a.mask0 : a.mask1;
                if (mk != nullptr) {
                  const float4 m = ldg4(mk + doff);
                  o[4 * h] = m.x > 0.f ? o[4 * h] : 0.f;
                  o[4 * h + 1] = m.y > 0.f ? o[4 * h + 1] : 0.f;
                  o[4 * h + 2] = m.z > 0.f ? o[4 * h + 2] : 0.f;
                  o[4 * h + 3] = m.w > 0.f ? o[4 * h + 3] : 0.f;
                }
                *reinterpret_cast<float4*>(dd.p + doff) = make_float4(o[4 * h], o[4 * h + 1], o[4 * h + 2], o[4 * h + 3]);
              }
            }
          }
        }
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
  const int units = ncg * a.nmma;  // 16-byte units per tap: [plane][n]
  float4* out = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a.out) + (size_t)coblk * a.w_coblk_stride + ch.w_off);
  const int nsh = 31 - __clz(a.nmma);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 9 * units; i += gridDim.x * blockDim.x) {
    const int tap = i / units, u = i - tap * units;
    const int n = u & (a.nmma - 1), pl = u >> nsh;
    int ci0;
    if (pl < ch.nA) ci0 = (ch.srcA == 0 ? 0 : a.C0) + (ch.cgA + pl) * 4;
    else ci0 = a.C0 + (ch.cgB + pl - ch.nA) * 4;
    const int co = coblk * kCoBlk + n;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (co < a.Cout && n < kCoBlk) {
      // forward: w[co][ci][tap]; dgrad: w[ci][co][8 - tap] of the OIHW tensor [Cout_w][Cin_w]
      const float* base = !a.transpose ? a.w + ((size_t)co * a.Cin_w + ci0) * 9 + tap : a.w + ((size_t)ci0 * a.Cin_w + co) * 9 + (8 - tap);
      const size_t cs = !a.transpose ? 9 : (size_t)a.Cin_w * 9;
      v.x = round_tf32(__ldg(base));
      v.y = round_tf32(__ldg(base + cs));
      v.z = round_tf32(__ldg(base + 2 * cs));
      v.w = round_tf32(__ldg(base + 3 * cs));
    }
    out[i] = v;
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
  int TH, TW, PW, tilesX, tilesY, nmb, nmma, cols, plane_bytes, a_bytes, w_bytes_max, w_res_bytes, tmem_cols, nchunks, ncoblk, kcg0, kcg1;
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
static bool tc_plan_k(int B, int H, int W, int C0, int C1, int Cout, TcPlan* p, bool resident, int maxcg);

// K chunks of at most 8 channel groups (32 channels); if a stage does not fit in shared memory (wide images with
// streamed weights) retry with 16- and 8-channel chunks
static bool tc_plan(int B, int H, int W, int C0, int C1, int Cout, TcPlan* p, bool resident = false) {
  for (int maxcg = 8; maxcg >= 2; maxcg >>= 1)
    if (tc_plan_k(B, H, W, C0, C1, Cout, p, resident, maxcg)) return true;
  return false;
}

static bool tc_plan_k(int B, int H, int W, int C0, int C1, int Cout, TcPlan* p, bool resident, int maxcg) {
  if (C0 < 8 || C0 % 8 != 0 || C1 % 8 != 0 || Cout % 8 != 0) return false;
  const int cg0 = C0 / 4, cg1 = C1 / 4;
  // chunks of at most 8 channel groups (32 channels); small concat pairs share one chunk
  p->nchunks = 0;
  p->kcg0 = cg0 < maxcg ? cg0 : maxcg;
  p->kcg1 = cg1 == 0 ? 0 : (cg1 < maxcg ? cg1 : maxcg);
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
  if (cg1 > 0 && cg0 + cg1 <= maxcg) {
    add(0, 0, cg0, 0, cg1);
  } else {
    if (cg0 / p->kcg0 + (cg1 ? cg1 / p->kcg1 : 0) > kMaxChunks) return false;
    for (int c = 0; c < cg0; c += p->kcg0) add(0, c, p->kcg0, 0, 0);
    for (int c = 0; c < cg1; c += p->kcg1) add(1, c, p->kcg1, 0, 0);
  }
  p->w_coblk_stride = woff;
  p->w_bytes_max = ((9 * max_ncg * p->nmma * 16) + 127) / 128 * 128;
  p->w_res_bytes = 0;
  if (resident) {  // weights built in shared memory once per CTA: no per-stage weight slot
    if (woff > kMaxResidentW) return false;
    p->w_res_bytes = (int)((woff + 127) / 128 * 128);
    p->w_bytes_max = 0;
  }
  // tile geometry
  p->tilesX = (W + 247) / 248;
  p->TW = (W + p->tilesX - 1) / p->tilesX;
  p->PW = (p->TW + 2 + 7) / 8 * 8;
  if (p->PW > 256) return false;
  // one CTA per SM: kStages stages of (A planes + weights) and two TMEM accumulator buffers of <= 256 columns
  const size_t smem_soft = (216 * 1024 - (size_t)p->w_res_bytes) / kStages, smem_hard = smem_soft;
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
  p->smem_bytes = (size_t)kStages * ((size_t)p->a_bytes + p->w_bytes_max) + p->w_res_bytes + 128;
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

bool conv3x3_tc_resident(int C0, int C1, int Cout, int H, int W) {
  TcPlan p;
  return tc_init() && tc_plan(1, H, W, C0, C1, Cout, &p, true);
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
  dim3 g(8, p.ncoblk, p.nchunks);
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
  cudaError_t le = launch_pdl(conv3x3_tc_kernel<COLS>, grid, dim3(kTcThreads), smem, st, tm0, tm1, ta);
  if (le != cudaSuccess) {
    set_error("conv3x3_tc launch: %s", cudaGetErrorString(le));
    return PU_ERR_CUDA;
  }
  return post_launch("conv3x3_tc");
}

int conv3x3_fwd_tc(const Conv3x3Args& a, cudaStream_t st) {
  if (!tc_init()) {
    set_error("pu_conv3x3_fwd: PU_MATH_TF32 requested but the tcgen05/TMA path is unavailable on this device");
    return PU_ERR_UNSUPPORTED;
  }
  TcPlan p;
  const int C1 = (a.s1.p != nullptr) ? a.s1.C : 0;
  if (!tc_plan(a.B, a.H, a.W, a.s0.C, C1, a.Cout, &p, a.wfmt != 0) || a.d0.C % 4 != 0 || (a.d1.p != nullptr && a.d1.C % 4 != 0)) {
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
  ta.mask0 = a.mask0; ta.mask1 = a.mask1;
  ta.B = a.B; ta.H = a.H; ta.W = a.W; ta.Cout = a.Cout; ta.relu = a.relu; ta.round_out = a.round_out;
  ta.TH = p.TH; ta.TW = p.TW; ta.PW = p.PW; ta.tilesX = p.tilesX; ta.tilesY = p.tilesY;
  ta.nmb = p.nmb; ta.nmma = p.nmma; ta.plane_bytes = p.plane_bytes; ta.a_bytes = p.a_bytes; ta.w_bytes_max = p.w_bytes_max;
  ta.tmem_cols = p.tmem_cols; ta.nchunks = p.nchunks; ta.w_coblk_stride = p.w_coblk_stride;
  ta.wfmt = a.wfmt; ta.Cin = a.Cin; ta.C0 = a.s0.C; ta.w_res_bytes = p.w_res_bytes;
  ta.s0 = a.s0; ta.s1 = a.s1;
  {
    const char* ld = getenv("PU_TC_LOADER");
    ta.loader = (ld != nullptr && ld[0] == 'c') ? 0 : 1;  // default: one 5-D TMA box per source; PU_TC_LOADER=cpasync: 4 LDGSTS warps
  }
  {
    const char* dbg = getenv("PU_TC_DEBUG");
    ta.debug = dbg ? atoi(dbg) : 0;
  }
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
