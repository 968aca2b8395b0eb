// conv3x3_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution for sm_100a (PU_MATH_TF32).
//
//   E[pixel, (kx, co)] += A[pixel + ky*PW, ci] * W[ky][ci, (kx, co)]      kind::tf32, fp32 accumulation in TMEM
//   out[pixel, co]      = E[pixel, (0,co)] + E[pixel+1, (1,co)] + E[pixel+2, (2,co)]   (epilogue, warp shuffles)
//
// Shared-memory layout ("flattened halo rows", NHWC): the CTA's input halo tile ((TH+2) rows x PW = TW+2 pixels) of
// a channel region (8, 16 or 32 channels of one source) is ONE 4-D TMA box (channels, x, y, image) written with the
// hardware 32/64/128-byte swizzle, i.e. shared memory holds [pixel][channels] rows exactly as NHWC memory does and
// out-of-bounds coordinates zero-fill — the conv's zero padding and the skip-connection crop window for free.  That is
// the canonical K-major SWIZZLE_xB UMMA operand layout, and (measured on B200, scripts/probes/umma_swizzle_probe.cu)
// the tensor core applies the swizzle to the ABSOLUTE shared address: a descriptor whose start address is moved by any
// number of pixel rows, and whose 8-row core groups are only 6 rows apart (SBO = 6 rows), reads exactly the rows it
// names.  So
//   * the three ky taps are descriptor start offsets of ky*PW rows into the SAME staged tile (input crosses L2->SMEM once),
//   * the three kx taps are folded into the N dimension: one MMA computes all three column taps of 128 rows, and the
//     one/two-row realignment happens in the epilogue.  A tcgen05.mma of this shape is bound by its A-operand read
//     (128 rows x 32 B at 64 B/clk = 64 clk, independent of N <= 128: measured), so folding cuts the MMA time 3x;
//   * to keep the realignment inside a warp, the 128 rows of an MMA are 16 groups of 8 rows that START 6 pixels apart:
//     rows 6,7 of a group duplicate rows 0,1 of the next one, every output pixel finds its +1/+2 neighbours in lanes
//     l+1, l+2 of its own 8-lane group, and an MMA block yields 96 output pixels.
// Skip-concat is a second tensor map whose regions follow the first in K.
//
// Persistent, warp-specialised CTA (one per SM): warp 0 = TMA producer, warps 1-3 = MMA issuers (blocks dealt
// round-robin, one elected lane issues), warps 4-11 = epilogue (TMEM lane quarter = warp % 4, two warps per quarter).
// nstages shared-memory stages cycle through full/empty mbarriers, two TMEM accumulator buffers through
// tmem_full/tmem_empty, so tile i's epilogue overlaps tile i+1's MMAs and tile i+2's loads.  The epilogue applies
// bias + residual + ReLU (+ RN rounding to TF32 so that the next layer's operands are exact TF32 values: the MMA
// truncates fp32 operands) or, for dgrad, the ReLU mask of the producing layer, and stores NHWC with 256-bit accesses.
// All waits are bounded (trap instead of hanging the GPU).
//
// Replaces nn.Conv2d(k=3,p=1)+ReLU(+add, +cat/crop) of reference unet_p.py:105-116,161-166 and
// unet_p_res.py:150-158,186-189,215-219 for channel counts that are multiples of 8; dgrad is the same kernel
// on transposed/flipped weights.
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "conv3x3.cuh"

namespace pu {

constexpr int kMaxChunks = 40;
constexpr int kCoBlk = 64;                     // output channels per CTA (grid.y splits larger Cout)
constexpr unsigned kMaxResidentW = 40 * 1024;  // largest weight image kept resident in shared memory (per co block)
constexpr int kBlkPix = 96;                    // output pixels per 128-row MMA block (16 groups x 6)
constexpr int kMaxStages = 4;

struct TcRegion {
  int src, c_off, cb;  // cb (8/16/32) channels of source src starting at channel c_off; shared-memory row = cb*4 bytes
  unsigned off;        // byte offset of the region inside a stage (1024-aligned)
};
struct TcChunk {       // one pipeline stage worth of K: one region, or the two small regions of a concat pair
  int nreg, ncg;       // ncg = 4-channel groups of the chunk (K / 4)
  TcRegion reg[2];
  unsigned w_off;      // byte offset of this chunk's weights inside one co-block of the weight image
  unsigned tx_bytes;   // bytes the TMA boxes of this chunk deliver
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
  int nmb, a_bytes, w_bytes_max, tmem_cols, nchunks, nstages;
  int wfmt;         // 0: wpk holds the packed B tiles (streamed per stage); 1/2: wpk is the raw OIHW weight (forward / dgrad)
                    // and the B tiles are built in shared memory once per CTA (resident)
  int Cin, C0;      // concatenated input channels and the split point (for the in-kernel weight build)
  int w_res_bytes;  // bytes of the resident weight image (0 in streamed mode)
  int debug;        // PU_TC_DEBUG experiments: 1 = skip MMAs, 2 = skip epilogue stores, 4 = load only the first stages
  unsigned w_coblk_stride;  // bytes
  TcChunk chunks[kMaxChunks];
};

__host__ __device__ constexpr int tc_n3(int cols) { return (3 * cols + 15) / 16 * 16; }  // MMA N: (kx, co) columns

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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout, version 1).
// layout: 0 = SWIZZLE_NONE (interleaved 8x16B core matrices), 6 / 4 / 2 = SWIZZLE_32B / 64B / 128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
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
constexpr int kMmaWarps = 3;
constexpr int kMmaWarp0 = 1;
constexpr int kEpiWarp0 = 4;  // first epilogue warp (multiple of 4: TMEM lane quarter = warp & 3)
constexpr int kEpiWarps = 8;
constexpr int kTcThreads = 32 * (kEpiWarp0 + kEpiWarps);

// B-operand tile of one chunk: [ky][4-channel group][n = kx*COLS + co][4 floats], SWIZZLE_NONE K-major
// (8x16B core matrices 128 B apart along N, the two K halves of an MMA N3*16 B apart), RN-rounded to TF32.
template <int COLS>
__device__ __forceinline__ void build_w_tile(const TcChunk& ch, float4* out, const float* __restrict__ w, int wfmt, int Cin, int Cout,
                                             int C0, int co_base, int u0, int ustep) {
  constexpr int N3 = tc_n3(COLS);
  const int units = ch.ncg * N3;
  for (int u = u0; u < units; u += ustep) {
    const int n = u % N3, kc = u / N3;
    const int kk = kc * 4;
    const TcRegion rg = kk < ch.reg[0].cb ? ch.reg[0] : ch.reg[1];
    const int ci0 = (rg.src ? C0 : 0) + rg.c_off + (kk < ch.reg[0].cb ? kk : kk - ch.reg[0].cb);
    const int kx = n / COLS, j = n - kx * COLS;
    const int co = co_base + j;
    const bool ok = n < 3 * COLS && co < Cout;
    // forward: w[co][ci][tap]; dgrad: w[ci][co][8 - tap]  (ci = conv input channel, co = conv output channel)
    const float* base = wfmt == 1 ? w + ((size_t)co * Cin + ci0) * 9 : w + ((size_t)ci0 * Cout + co) * 9;
    const size_t cs = wfmt == 1 ? 9 : (size_t)Cout * 9;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int tap = ky * 3 + kx;
      const int t = wfmt == 1 ? tap : 8 - tap;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        v.x = round_tf32(__ldg(base + t));
        v.y = round_tf32(__ldg(base + cs + t));
        v.z = round_tf32(__ldg(base + 2 * cs + t));
        v.w = round_tf32(__ldg(base + 3 * cs + t));
      }
      out[(ky * ch.ncg + kc) * N3 + n] = v;
    }
  }
}

template <int COLS>
__global__ void __launch_bounds__(kTcThreads, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm0,
                                                                   const __grid_constant__ CUtensorMap tm1, const TcArgs a) {
  constexpr int N3 = tc_n3(COLS);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle patterns repeat every 1024 B
  const int nst = a.nstages;
  const int stage_bytes = a.a_bytes + a.w_bytes_max;  // w_bytes_max == 0 when the weights are resident
  uint8_t* smWres = smem + nst * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smWres + a.w_res_bytes);
  // bars: [0,kMaxStages) full, [kMaxStages,2kMaxStages) empty, then tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int st) { return bar0 + 8u * st; };
  auto empty_bar = [&](int st) { return bar0 + 8u * (kMaxStages + st); };
  auto tfull_bar = [&](int as) { return bar0 + 8u * (2 * kMaxStages + as); };
  auto tempty_bar = [&](int as) { return bar0 + 8u * (2 * kMaxStages + 2 + as); };

  pdl_prologue();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int coblk = blockIdx.y;
  const int co_base = coblk * kCoBlk;
  const int tiles_per_img = a.tilesX * a.tilesY;
  const int ntiles = tiles_per_img * a.B;
  const int acc_cols = a.nmb * N3;  // TMEM columns of one accumulator buffer

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < kMaxStages; ++i) {
        mbar_init(full_bar(i), 1);
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
    // Build the tf32 B-operand tiles of every K chunk straight from the OIHW weight tensor (no pack kernel).
    for (int c = 0; c < a.nchunks; ++c)
      build_w_tile<COLS>(a.chunks[c], reinterpret_cast<float4*>(smWres + a.chunks[c].w_off), a.wpk, a.wfmt, a.Cin, a.Cout, a.C0, co_base,
                         tid, kTcThreads);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    const uint8_t* wblk = reinterpret_cast<const uint8_t*>(a.wpk) + (size_t)coblk * a.w_coblk_stride;
    uint32_t it = 0;
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH;
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        mbar_wait(empty_bar(st), ph ^ 1);  // passes immediately on a fresh barrier
        if (elect_one()) {
          const TcChunk& ch = a.chunks[c];
          const uint32_t w_bytes = (uint32_t)(3 * ch.ncg * N3 * 16);
          const uint32_t sS = smem_u32(smem + st * stage_bytes);
          if ((a.debug & 4) && it >= (uint32_t)nst) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_bar(st)) : "memory");
          } else {
            mbar_expect_tx(full_bar(st), ch.tx_bytes + (a.wfmt == 0 ? w_bytes : 0u));
            tma_load_4d(sS + ch.reg[0].off, ch.reg[0].src == 0 ? &tm0 : &tm1, full_bar(st), ch.reg[0].c_off, x0 - 1, y0 - 1, b);
            if (ch.nreg > 1)
              tma_load_4d(sS + ch.reg[1].off, ch.reg[1].src == 0 ? &tm0 : &tm1, full_bar(st), ch.reg[1].c_off, x0 - 1, y0 - 1, b);
            if (a.wfmt == 0) bulk_load(sS + a.a_bytes, wblk + ch.w_off, w_bytes, full_bar(st));
          }
        }
        __syncwarp();
        if (++st == nst) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp >= kMmaWarp0 && warp < kEpiWarp0) {
    // ================= MMA issuers =================
    // The whole warp runs the warp-uniform loop and one elected lane executes each tcgen05 instruction, so the
    // descriptors live in uniform registers.  M blocks are innermost: consecutive MMAs write different accumulators
    // and are not serialised on the accumulate dependency of one small TMEM tile.
    const int mw = warp - kMmaWarp0;
    // instruction descriptor: D=f32, A=B=tf32, K-major both, N = N3, M = 128 (cute::UMMA::InstrDescriptor)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N3 >> 3) << 17) | ((128u >> 4) << 24);
    const bool leader = elect_one();  // elected once: the issue loop must stay a handful of instructions per MMA
    uint32_t tcount = 0;
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(tempty_bar(as), aph ^ 1);  // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(as * acc_cols);
      for (int c = 0; c < a.nchunks; ++c) {
        const TcChunk& ch = a.chunks[c];
        mbar_wait(full_bar(st), ph);
        tc_fence_after();
        const uint32_t sS = smem_u32(smem + st * stage_bytes);
        const uint64_t b_base = umma_desc(a.wfmt == 0 ? sS + a.a_bytes : smem_u32(smWres) + ch.w_off, (uint32_t)(N3 * 16), 128, 0);
        for (int ky = 0; ky < 3; ++ky) {
          uint32_t kc = 0;
          for (int r = 0; r < ch.nreg; ++r) {
            const uint32_t rb = (uint32_t)ch.reg[r].cb * 4;  // row bytes = swizzle span
            const uint32_t layout = rb == 32 ? 6u : (rb == 64 ? 4u : 2u);
            const uint64_t a_base = umma_desc(sS + ch.reg[r].off + (uint32_t)(ky * a.PW) * rb, 16, 6 * rb, layout);
            const uint32_t mb_step = 6 * rb;  // kBlkPix rows, in 16-byte descriptor units
            const int ksteps = ch.reg[r].cb >> 3;
            for (int ks = 0; ks < ksteps; ++ks, kc += 2) {
              const uint64_t bd = b_base + (uint64_t)((ky * ch.ncg + kc) * N3);
              uint64_t ad = a_base + (uint64_t)(2 * ks + mw * mb_step);
              uint32_t d = d0 + mw * N3;
              const uint32_t acc = (c | ky | (int)kc) ? 1u : 0u;
#pragma unroll 3
              for (int mb = mw; mb < a.nmb; mb += kMmaWarps) {
                if (leader && !(a.debug & 1)) umma_tf32(d, ad, bd, idesc, acc);
                d += kMmaWarps * N3;
                ad += kMmaWarps * mb_step;
              }
            }
          }
        }
        if (leader) tc_commit(empty_bar(st));  // frees the stage once these MMAs have read it
        __syncwarp();
        if (++st == nst) { st = 0; ph ^= 1; }
      }
      if (leader) tc_commit(tfull_bar(as));  // accumulator complete
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ================= epilogue: TMEM -> registers -> kx realignment -> bias/residual/ReLU/mask -> NHWC global =================
    // warp e handles TMEM lane quarter (warp % 4) of the MMA blocks mb = set, set+2, ... (set = e / 4).  A unit is
    // 8 output channels of one block: three tcgen05.ld.x8 (the kx = 0,1,2 column groups), two shuffles per channel.
    // Two units are fetched per tcgen05.wait::ld and their bias / residual / mask vectors are requested before the
    // wait.  The loop is issue-bound (measured), so everything that is per-tile or per-thread constant is hoisted:
    // per-tile base pointers are warp-uniform, per-pixel offsets are 32-bit, (yy, xx) advance without a division.
    constexpr int NQ = COLS / 8;
    const int quarter = warp & 3;
    const int set = (warp - kEpiWarp0) >> 2;
    const int R = quarter * 32 + lane;
    const int gi = R & 7;                     // row inside its 8-row group; rows 6,7 duplicate the next group's 0,1
    const int prow = (R >> 3) * 6 + gi;       // pixel offset of this row inside an MMA block
    const int p0 = set * kBlkPix + prow;
    const int yy0 = p0 / a.PW, xx0 = p0 - yy0 * a.PW;
    const int step_y = (2 * kBlkPix) / a.PW, step_x = 2 * kBlkPix - step_y * a.PW;  // this warp advances two blocks at a time
    const bool has_mask = a.mask0 != nullptr || a.mask1 != nullptr;
    const bool has_res = a.res != nullptr && !has_mask;
    const int relu = a.relu, round_out = a.round_out;
    const int nq_valid = (a.Cout - co_base + 7) / 8 < NQ ? (a.Cout - co_base + 7) / 8 : NQ;  // 8-channel groups of this co block
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH;
      const int ymax = (a.H - y0) < a.TH ? (a.H - y0) : a.TH, xmax = (a.W - x0) < a.TW ? (a.W - x0) : a.TW;
      // warp-uniform bases of this tile in every tensor the epilogue touches
      float* const d0b = a.d0.p + (((size_t)b * a.d0.Hs + (y0 + a.d0.oy)) * a.d0.Ws + (x0 + a.d0.ox)) * a.d0.C;
      float* const d1b = a.d1.p == nullptr ? nullptr : a.d1.p + (((size_t)b * a.d1.Hs + (y0 + a.d1.oy)) * a.d1.Ws + (x0 + a.d1.ox)) * a.d1.C;
      const ptrdiff_t m0d = a.mask0 == nullptr ? 0 : a.mask0 - a.d0.p;  // masks share the geometry of their destination
      const ptrdiff_t m1d = a.mask1 == nullptr ? 0 : a.mask1 - a.d1.p;
      const float* const rsb = has_res ? a.res + (((size_t)b * a.H + y0) * a.W + x0) * a.Cout + co_base : nullptr;
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * acc_cols);

      // one pair of units: (taddr, ok, dst, aux source, bias source) x 2; unit B may be absent (warp-uniform)
      auto pair = [&](uint32_t tA, bool okA, float* dA, const float* xA, const float* bA, bool haveB, uint32_t tB, bool okB, float* dB,
                      const float* xB, const float* bB) {
        uint32_t v[2][24];
        float aux[2][8], bias[2][8];
        tmem_ld8(tA, v[0]);
        tmem_ld8(tA + COLS, v[0] + 8);
        tmem_ld8(tA + 2 * COLS, v[0] + 16);
        if (haveB) {
          tmem_ld8(tB, v[1]);
          tmem_ld8(tB + COLS, v[1] + 8);
          tmem_ld8(tB + 2 * COLS, v[1] + 16);
        }
        if (bA != nullptr) {  // parameters may live in a packed arena: only 4-byte alignment is guaranteed
          if (okA) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bias[0][j] = __ldg(bA + j);
          }
          if (haveB && okB) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bias[1][j] = __ldg(bB + j);
          }
        }
        if (xA != nullptr && okA) ldg8(xA, aux[0]);
        if (haveB && xB != nullptr && okB) ldg8(xB, aux[1]);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k == 1 && !haveB) break;  // warp-uniform: the shuffles below need the whole warp
          const bool ok = k == 0 ? okA : okB;
          float* const dst = k == 0 ? dA : dB;
          const bool has_aux = (k == 0 ? xA : xB) != nullptr;
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float e1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[k][8 + j]), 1);
            const float e2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[k][16 + j]), 2);
            o[j] = (__uint_as_float(v[k][j]) + e1) + e2;
          }
          if (!ok) continue;
          if (bA != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += bias[k][j];
          }
          if (has_aux) {
            if (has_mask) {  // dgrad: ReLU mask of the layer that produced this source
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = aux[k][j] > 0.f ? o[j] : 0.f;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] += aux[k][j];
            }
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
          }
          if (round_out) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = round_tf32(o[j]);
          }
          stg8(dst, o);
        }
      };
      // per-unit operands: q selects the destination tensor (warp-uniform), (yy, xx) the pixel (per lane)
      struct Unit { uint32_t t; bool ok; float* d; const float* x; const float* bsrc; };
      auto unit = [&](int mb, int q, int yy, int xx) {
        Unit un;
        un.t = tbase + (uint32_t)(mb * N3 + 8 * q);
        un.ok = gi < 6 && yy < ymax && xx < xmax && q < nq_valid && !(a.debug & 2);
        const int co = co_base + 8 * q;
        const bool first = co < a.d0.C;
        const int off = first ? (yy * a.d0.Ws + xx) * a.d0.C + co : (yy * a.d1.Ws + xx) * a.d1.C + (co - a.d0.C);
        un.d = (first ? d0b : d1b) + off;
        un.x = nullptr;
        if (has_mask) {
          if ((first ? a.mask0 : a.mask1) != nullptr) un.x = un.d + (first ? m0d : m1d);
        } else if (has_res) {
          un.x = rsb + (yy * a.W + xx) * a.Cout + 8 * q;
        }
        un.bsrc = a.bias == nullptr ? nullptr : a.bias + co;
        return un;
      };
      int yy = yy0, xx = xx0;
      auto advance = [&]() {
        yy += step_y;
        xx += step_x;
        if (xx >= a.PW) { xx -= a.PW; ++yy; }
      };
      if (NQ == 1) {
        for (int mb = set; mb < a.nmb; mb += 4) {  // two of this warp's blocks per TMEM wait
          const Unit ua = unit(mb, 0, yy, xx);
          advance();
          const bool haveB = mb + 2 < a.nmb;
          const Unit ub = unit(mb + 2, 0, yy, xx);
          advance();
          pair(ua.t, ua.ok, ua.d, ua.x, ua.bsrc, haveB, ub.t, ub.ok, ub.d, ub.x, ub.bsrc);
        }
      } else {
        for (int mb = set; mb < a.nmb; mb += 2) {
#pragma unroll
          for (int q = 0; q < NQ; q += 2) {
            const Unit ua = unit(mb, q, yy, xx);
            const Unit ub = unit(mb, q + 1, yy, xx);
            pair(ua.t, ua.ok, ua.d, ua.x, ua.bsrc, true, ub.t, ub.ok, ub.d, ub.x, ub.bsrc);
          }
          advance();
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

// ---- weight packing for the streamed-weights mode ---------------------------------------------------
// out[coblk][chunk][ky][4-channel group][n][4]  (n = kx*COLS + co, zero beyond the co-block's valid channels)
struct TcPackArgs {
  const float* w;  // OIHW [Cout_w][Cin_w][3][3]
  float* out;
  int wfmt;        // 1: forward, 2: dgrad (conv input channels = Cout_w, outputs = Cin_w, taps flipped)
  int Cin, Cout, C0, nchunks;
  unsigned w_coblk_stride;
  TcChunk chunks[kMaxChunks];
};

template <int COLS>
__global__ void pack_w3x3_tc_kernel(const TcPackArgs a) {
  const int coblk = blockIdx.y, c = blockIdx.z;
  float4* out = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a.out) + (size_t)coblk * a.w_coblk_stride + a.chunks[c].w_off);
  build_w_tile<COLS>(a.chunks[c], out, a.w, a.wfmt, a.Cin, a.Cout, a.C0, coblk * kCoBlk, blockIdx.x * blockDim.x + threadIdx.x,
                     gridDim.x * blockDim.x);
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
  int TH, TW, PW, tilesX, tilesY, nmb, cols, n3, a_bytes, w_bytes_max, w_res_bytes, tmem_cols, nchunks, ncoblk, nstages, cb0, cb1;
  unsigned w_coblk_stride;
  size_t smem_bytes;
  TcChunk chunks[kMaxChunks];
};

static int next_pow2_cols(int c) {
  int p = 32;
  while (p < c) p <<= 1;
  return p;
}

static int region_channels(int C) { return C % 32 == 0 ? 32 : (C % 16 == 0 ? 16 : 8); }

// channel plan: K chunks, weight image layout.  false if the channel counts do not fit the tensor-core path.
static bool tc_plan_channels(int C0, int C1, int Cout, TcPlan* p) {
  if (C0 < 8 || C0 % 8 != 0 || C1 < 0 || C1 % 8 != 0 || Cout < 8 || Cout % 8 != 0) return false;
  const int cout_blk = Cout < kCoBlk ? Cout : kCoBlk;
  p->cols = cout_blk <= 8 ? 8 : (cout_blk <= 16 ? 16 : (cout_blk <= 32 ? 32 : 64));
  p->n3 = tc_n3(p->cols);
  p->ncoblk = (Cout + kCoBlk - 1) / kCoBlk;
  p->cb0 = region_channels(C0);
  p->cb1 = C1 > 0 ? region_channels(C1) : 0;
  p->nchunks = 0;
  unsigned woff = 0;
  auto add = [&](int nreg, TcRegion r0, TcRegion r1) {
    TcChunk& ch = p->chunks[p->nchunks++];
    ch.nreg = nreg; ch.reg[0] = r0; ch.reg[1] = r1;
    ch.ncg = (r0.cb + (nreg > 1 ? r1.cb : 0)) / 4;
    ch.w_off = woff;
    ch.tx_bytes = 0;
    woff += (unsigned)(3 * ch.ncg * p->n3 * 16);
  };
  const TcRegion none = {0, 0, 0, 0};
  if (C1 > 0 && C0 == p->cb0 && C1 == p->cb1 && C0 + C1 <= 32) {  // a small concat pair shares one stage
    add(2, TcRegion{0, 0, p->cb0, 0}, TcRegion{1, 0, p->cb1, 0});
  } else {
    if (C0 / p->cb0 + (C1 ? C1 / p->cb1 : 0) > kMaxChunks) return false;
    for (int c = 0; c < C0; c += p->cb0) add(1, TcRegion{0, c, p->cb0, 0}, none);
    for (int c = 0; c < C1; c += p->cb1) add(1, TcRegion{1, c, p->cb1, 0}, none);
  }
  p->w_coblk_stride = woff;
  return true;
}

// tile geometry, pipeline depth and shared-memory layout for one problem size
static bool tc_plan(int B, int H, int W, int C0, int C1, int Cout, TcPlan* p, bool resident = false) {
  if (!tc_plan_channels(C0, C1, Cout, p)) return false;
  int max_ncg = 0;
  for (int i = 0; i < p->nchunks; ++i) max_ncg = p->chunks[i].ncg > max_ncg ? p->chunks[i].ncg : max_ncg;
  p->w_bytes_max = ((3 * max_ncg * p->n3 * 16) + 1023) / 1024 * 1024;
  p->w_res_bytes = 0;
  if (resident) {  // weights built in shared memory once per CTA: no per-stage weight slot
    if (p->w_coblk_stride > kMaxResidentW) return false;
    p->w_res_bytes = (int)((p->w_coblk_stride + 127) / 128 * 128);
    p->w_bytes_max = 0;
  }
  const size_t budget = 222 * 1024 - (size_t)p->w_res_bytes;  // minus barriers and the 1024-byte alignment slack below
  const int nmb_max = 256 / p->n3;  // two accumulator buffers of <= 256 TMEM columns
  auto stage_a_bytes = [&](int th, int pw, int nmb) {
    // rows a region must hold: the halo tile, and whatever the last MMA block's shifted reads touch beyond it
    int rows = (th + 2) * pw;
    const int reach = (nmb - 1) * kBlkPix + 98 + 2 * pw;
    if (reach > rows) rows = reach;
    size_t worst = 0;
    for (int i = 0; i < p->nchunks; ++i) {
      size_t s = 0;
      for (int r = 0; r < p->chunks[i].nreg; ++r) s += ((size_t)rows * p->chunks[i].reg[r].cb * 4 + 1023) / 1024 * 1024;
      if (s > worst) worst = s;
    }
    return worst;
  };
  double best_cost = 1e30;
  int best_th = 0, best_tx = 0;
  const int tx_min = (W + 253) / 254;
  for (int tilesX = tx_min; tilesX <= tx_min + 7; ++tilesX) {
    const int tw = (W + tilesX - 1) / tilesX;
    if (tilesX > tx_min && tw < 16) break;
    const int pw = tw + 2;
    for (int th = (H < 254 ? H : 254); th >= 1; --th) {
      const int nmb = (th * pw + kBlkPix - 1) / kBlkPix;
      if (nmb > nmb_max) continue;
      const size_t stage = stage_a_bytes(th, pw, nmb) + p->w_bytes_max;
      if (2 * stage > budget) continue;
      const long long ntiles = (long long)B * ((H + th - 1) / th) * ((W + tw - 1) / tw);
      const long long waves = (ntiles + kNumSMs - 1) / kNumSMs;
      // per-SM time ~ waves x (MMA rows + staged rows + a fixed per-tile hand-off cost), in units of one pixel row
      const double cost = (double)waves * (nmb * 128 + (th + 2) * pw + 256);
      if (cost < best_cost) { best_cost = cost; best_th = th; best_tx = tilesX; }
    }
  }
  if (best_th == 0) return false;
  {  // tuning overrides (experiments only)
    const char* eth = getenv("PU_TC_TH");
    const char* etx = getenv("PU_TC_TX");
    if (eth != nullptr && etx != nullptr) { best_th = atoi(eth) < H ? atoi(eth) : H; best_tx = atoi(etx); }
  }
  p->tilesX = best_tx;
  p->TW = (W + best_tx - 1) / best_tx;
  p->tilesX = (W + p->TW - 1) / p->TW;
  p->PW = p->TW + 2;
  p->TH = best_th;
  p->tilesY = (H + best_th - 1) / best_th;
  p->nmb = (p->TH * p->PW + kBlkPix - 1) / kBlkPix;
  int rows = (p->TH + 2) * p->PW;
  const int reach = (p->nmb - 1) * kBlkPix + 98 + 2 * p->PW;
  if (reach > rows) rows = reach;
  for (int i = 0; i < p->nchunks; ++i) {
    TcChunk& ch = p->chunks[i];
    unsigned off = 0;
    ch.tx_bytes = 0;
    for (int r = 0; r < ch.nreg; ++r) {
      ch.reg[r].off = off;
      off += (unsigned)(((size_t)rows * ch.reg[r].cb * 4 + 1023) / 1024 * 1024);
      ch.tx_bytes += (unsigned)((p->TH + 2) * p->PW * ch.reg[r].cb * 4);
    }
  }
  p->a_bytes = (int)stage_a_bytes(p->TH, p->PW, p->nmb);
  const size_t stage = (size_t)p->a_bytes + p->w_bytes_max;
  p->nstages = (int)(budget / stage);
  if (p->nstages > kMaxStages) p->nstages = kMaxStages;
  p->tmem_cols = next_pow2_cols(2 * p->nmb * p->n3);
  p->smem_bytes = (size_t)p->nstages * stage + p->w_res_bytes + 256 + 1024;
  return p->nstages >= 2 && p->tmem_cols <= 512;
}

static int make_tmap(CUtensorMap* tm, const View& v, int B, int H, int W, int PW, int TH, int cb) {
  // dims, innermost first: (channel, x, y, image) over the HxW window of the stored tensor; box = one halo tile of a region
  const float* base = v.p + ((size_t)v.oy * v.Ws + v.ox) * v.C;
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)v.C * 4, (cuuint64_t)v.Ws * v.C * 4, (cuuint64_t)v.Hs * v.Ws * v.C * 4};
  cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)PW, (cuuint32_t)(TH + 2), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = cb == 8 ? CU_TENSOR_MAP_SWIZZLE_32B : (cb == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv3x3_tc: cuTensorMapEncodeTiled failed (CUresult %d) for window %dx%d C=%d box (%d,%d,%d,1)", (int)r, H, W, v.C, cb, PW,
              TH + 2);
    return PU_ERR_CUDA;
  }
  return PU_OK;
}

bool conv3x3_tc_ok(int C0, int C1, int Cout, int Cd0, int Cd1) {
  if (!tc_init()) return false;
  TcPlan p;
  if (!tc_plan_channels(C0, C1, Cout, &p)) return false;
  if (Cd0 % 8 != 0 || Cd1 % 8 != 0 || Cd0 + Cd1 != Cout) return false;
  return true;
}

bool conv3x3_tc_resident(int C0, int C1, int Cout, int H, int W) {
  TcPlan p;
  return tc_init() && tc_plan(1, H, W, C0, C1, Cout, &p, true);
}

long long conv3x3_tc_weight_floats(int C0, int C1, int Cout) {
  TcPlan p;
  if (!tc_plan_channels(C0, C1, Cout, &p)) return 0;
  return (long long)p.w_coblk_stride * p.ncoblk / 4;
}

int conv3x3_tc_pack(const float* w_oihw, float* out, int Cout_w, int Cin_w, int transpose, int C0, cudaStream_t st) {
  // roles of the conv that will consume the packed weights
  const int cin = transpose ? Cout_w : Cin_w;
  const int cout = transpose ? Cin_w : Cout_w;
  const int c0 = transpose ? Cout_w : C0;
  TcPlan p;
  if (!tc_plan_channels(c0, cin - c0, cout, &p)) {
    set_error("pu_pack_w3x3: channels (%d|%d -> %d) do not fit the tcgen05 path", c0, cin - c0, cout);
    return PU_ERR_UNSUPPORTED;
  }
  TcPackArgs pa;
  pa.w = w_oihw; pa.out = out; pa.wfmt = transpose ? 2 : 1;
  pa.Cin = cin; pa.Cout = cout; pa.C0 = c0; pa.nchunks = p.nchunks;
  pa.w_coblk_stride = p.w_coblk_stride;
  for (int i = 0; i < p.nchunks; ++i) pa.chunks[i] = p.chunks[i];
  dim3 g(8, p.ncoblk, p.nchunks);
  switch (p.cols) {
    case 8: pack_w3x3_tc_kernel<8><<<g, 256, 0, st>>>(pa); break;
    case 16: pack_w3x3_tc_kernel<16><<<g, 256, 0, st>>>(pa); break;
    case 32: pack_w3x3_tc_kernel<32><<<g, 256, 0, st>>>(pa); break;
    default: pack_w3x3_tc_kernel<64><<<g, 256, 0, st>>>(pa); break;
  }
  return post_launch("pu_pack_w3x3 (tc)");
}

template <int COLS>
static int launch_tc(const CUtensorMap& tm0, const CUtensorMap& tm1, const TcArgs& ta, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
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
  if (!tc_plan(a.B, a.H, a.W, a.s0.C, C1, a.Cout, &p, a.wfmt != 0) || a.d0.C % 8 != 0 || (a.d1.p != nullptr && a.d1.C % 8 != 0)) {
    set_error("pu_conv3x3_fwd: PU_MATH_TF32 does not support this shape (C %d|%d -> %d, %dx%d); check pu_conv3x3_tc_ok", a.s0.C, C1,
              a.Cout, a.H, a.W);
    return PU_ERR_UNSUPPORTED;
  }
  CUtensorMap tm0, tm1;
  int rc = make_tmap(&tm0, a.s0, a.B, a.H, a.W, p.PW, p.TH, p.cb0);
  if (rc) return rc;
  if (C1 > 0) {
    rc = make_tmap(&tm1, a.s1, a.B, a.H, a.W, p.PW, p.TH, p.cb1);
    if (rc) return rc;
  } else {
    tm1 = tm0;
  }
  TcArgs ta;
  ta.wpk = a.wp; ta.bias = a.bias; ta.res = a.res; ta.d0 = a.d0; ta.d1 = a.d1;
  ta.mask0 = a.mask0; ta.mask1 = a.mask1;
  ta.B = a.B; ta.H = a.H; ta.W = a.W; ta.Cout = a.Cout; ta.relu = a.relu; ta.round_out = a.round_out;
  ta.TH = p.TH; ta.TW = p.TW; ta.PW = p.PW; ta.tilesX = p.tilesX; ta.tilesY = p.tilesY;
  ta.nmb = p.nmb; ta.a_bytes = p.a_bytes; ta.w_bytes_max = p.w_bytes_max; ta.nstages = p.nstages;
  ta.tmem_cols = p.tmem_cols; ta.nchunks = p.nchunks; ta.w_coblk_stride = p.w_coblk_stride;
  ta.wfmt = a.wfmt; ta.Cin = a.Cin; ta.C0 = a.s0.C; ta.w_res_bytes = p.w_res_bytes;
  {
    const char* dbg = getenv("PU_TC_DEBUG");
    ta.debug = dbg ? atoi(dbg) : 0;
    if (ta.debug & 8)
      fprintf(stderr, "conv3x3_tc plan: %d|%d->%d %dx%d B=%d: TH=%d TW=%d tiles %dx%d nmb=%d n3=%d stages=%d a_bytes=%d w_stage=%d w_res=%d smem=%zu\n",
              a.s0.C, C1, a.Cout, a.H, a.W, a.B, p.TH, p.TW, p.tilesX, p.tilesY, p.nmb, p.n3, p.nstages, p.a_bytes, p.w_bytes_max,
              p.w_res_bytes, p.smem_bytes);
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
