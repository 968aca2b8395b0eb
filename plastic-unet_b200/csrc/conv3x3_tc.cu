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
// Persistent, warp-specialised CTA (one per SM): warp 0 = TMA producer, warps 1-3 and 12 = MMA issuers (blocks dealt
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
#include <atomic>
#include <mutex>
#include "conv3x3.cuh"

namespace pu {

constexpr int kMaxChunks = 64;  // K chunks per conv (16-channel chunks in flat mode: Cin <= 1024)
constexpr int kCoBlk = 64;                     // output channels per CTA (grid.y splits larger Cout)
constexpr unsigned kMaxResidentW = 40 * 1024;  // largest weight image kept resident in shared memory (per co block)
constexpr int kBlkPix = 96;                    // output pixels per 128-row MMA block, kx folded into N (16 groups x 6)
constexpr int kBlkPixFlat = 128;               // ... and with one MMA per tap (FOLD = false): 128 consecutive flattened halo pixels
constexpr int kMaxStages = 4;
constexpr int kMaxAcc = 4;                     // TMEM accumulator buffers (tiles in flight between the MMA and the epilogue warps)
constexpr int kTileQ = 4;                      // depth of the producer -> consumers tile-index queue (dynamic scheduler)
constexpr int kMaxCoBlk = 16;                  // co blocks per launch slot of the tile counters

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
  const unsigned char* mask0;  // | null: packed ReLU mask, geometry of d0 with C/8 bytes per pixel; clear bit -> stored value 0
  const unsigned char* mask1;
  unsigned char* mask_out;     // | null: packed mask of this conv's own output (geometry of d0, Cout/8 bytes per pixel)
  int B, H, W, Cout, relu, round_out;
  int TH, TW, PW, tilesX, tilesY;
  int nmb, a_bytes, w_bytes_max, tmem_cols, nchunks, nstages;
  int nacc;         // accumulator buffers in TMEM (2..kMaxAcc): nacc * nmb * NACC columns
  int wfmt;         // 0: wpk holds the packed B tiles (streamed per stage); 1/2: wpk is the raw OIHW weight (forward / dgrad)
                    // and the B tiles are built in shared memory once per CTA (resident)
  int Cin, C0;      // concatenated input channels and the split point (for the in-kernel weight build)
  int w_res_bytes;  // bytes of the resident weight image (0 in streamed mode)
  unsigned int* tile_ctr;  // dynamic tile scheduler: [co block] next-tile counters of this launch's slot (zero between launches)
  unsigned int* done_ctr;  // CTAs of this launch that have finished (the last one re-zeroes the slot)
  int dynamic;             // 0: static round-robin tiles (tile = blockIdx.x + k * gridDim.x), 1: the dynamic scheduler
  int debug;        // PU_TC_DEBUG experiments: 1 = skip MMAs, 2 = skip epilogue stores, 4 = load only the first stages
  unsigned w_coblk_stride;  // bytes
  TcChunk chunks[kMaxChunks];
};

__host__ __device__ constexpr int tc_n3(int cols) { return (3 * cols + 15) / 16 * 16; }  // MMA N: (kx, co) columns

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a lost TMA transaction or MMA commit must not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
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
// Measured on B200 (scripts/probes/umma_rate_probe.cu): one thread can issue a tcgen05.mma (M=128, K=8, tf32 or bf16)
// only every ~128 clk whatever N <= 128 is, two threads reach 64 clk/MMA, four or more saturate the tensor pipe at
// ~42 + 0.25 N clk/MMA; operand swizzle mode, row-shifted start addresses and accumulator reuse make no difference,
// and a tcgen05.commit costs the issuing thread about as much as an MMA.  So the MMAs of a tile are issued by three
// warps (the blocks dealt round-robin).  Variants with 4 or 7 issuing warps, two warp groups alternating tiles, or the
// stage barrier doubling as the "accumulators complete" signal were all measured within +-8% of this one: per tile the
// chain wait -> 6..8 issues -> completion -> epilogue -> release of the TMEM buffer is latency-bound.
// NB a waiter that falls TWO phases behind an mbarrier never returns from a parity wait (and one that is two phases
// ahead passes spuriously): every barrier here has waiters that consume each of its phases in order.
constexpr int kMmaWarps = 3;
constexpr int kEpiWarp0 = 4;  // first epilogue warp (multiple of 4: TMEM lane quarter = warp & 3)
constexpr int kEpiWarps = 8;
constexpr int kEpiSets = kEpiWarps / 4;  // warps per TMEM lane quarter
constexpr int kTcThreads = 32 * (kEpiWarp0 + kEpiWarps);

// B-operand tile of one chunk: [ky][K step of 8 channels][n = kx*COLS + co][8 floats], K-major SWIZZLE_32B (32-byte
// rows, the two 16-byte halves of a row exchanged where bit 7 of the byte address is set), RN-rounded to TF32.
// (A SWIZZLE_NONE B tile costs the tensor core ~1 clk per N row: measured 150 clk per MMA at N = 96.)
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
      out[((ky * (ch.ncg >> 1) + (kc >> 1)) * N3 + n) * 2 + ((kc & 1) ^ ((n >> 2) & 1))] = v;
    }
  }
}

// FOLD = true: the three kx taps are folded into N (N = 3*COLS, one MMA per ky and K step, 96 output pixels per 128-row
// block, realignment in the epilogue) — the right shape for the narrow, HBM-bound layers where the fixed cost per MMA
// dominates.  FOLD = false ("flat"): one MMA per tap (N = COLS, nine per K step) on 128 CONSECUTIVE flattened halo pixels:
// every row is an output pixel, the accumulators need COLS instead of 3*COLS TMEM columns per block, so a tile holds 3x the
// pixels per weight byte streamed — the shape for the wide (>= 64 channel), tensor-bound layers (SURVEY.md §8d "TC demo").
// KS = true ("tap-row split", flat mode, small tiles of the deep layers): ONE thread issues a tcgen05.mma only every ~128 clk, so a
// 64-channel block (72 MMAs) costs its issuing warp 4.8 us while the other MMA warps idle when a tile has one or two blocks.
// With KS the three MMA warps each issue ONE tap row (ky = warp) of every block into their own partial accumulator
// (block mb, partial ky at TMEM column mb*3*NACC + ky*NACC) and the epilogue adds the three partials.
template <int COLS, bool FOLD, bool KS = false>
__global__ void __launch_bounds__(kTcThreads, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm0,
                                                                   const __grid_constant__ CUtensorMap tm1, const TcArgs a) {
  constexpr int N3 = tc_n3(COLS);
  // flat MMA width: COLS, but an M = 128 MMA needs N % 16 == 0: an 8-channel layer computes 8 unused extra columns (the next
  // tap's rows of the B tile; MMAs are free in these HBM-bound layers)
  constexpr int NACC = FOLD ? N3 : (COLS < 16 ? 16 : COLS);  // TMEM columns of one MMA block's accumulators
  constexpr int BLK = FOLD ? kBlkPix : kBlkPixFlat;  // output pixels per MMA block
  constexpr int NBLK = KS ? 3 * NACC : NACC;         // TMEM columns per block (KS: three partial accumulators)
  static_assert(!KS || (!FOLD && COLS >= 16), "tap-row split is a flat-mode variant");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle patterns repeat every 1024 B
  const int nst = a.nstages;
  const int stage_bytes = a.a_bytes + a.w_bytes_max;  // w_bytes_max == 0 when the weights are resident
  uint8_t* smWres = smem + nst * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smWres + a.w_res_bytes);
  // bars: [0,kMaxStages) full, [kMaxStages,2kMaxStages) empty, then tmem_full[kMaxAcc], tmem_empty[kMaxAcc], tile queue full[kTileQ], empty[kTileQ]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kMaxAcc + 2 * kTileQ);
  volatile int* tileq = reinterpret_cast<volatile int*>(tmem_slot + 2);  // [kTileQ] tile indices handed from the producer to the other roles
  int* wtab = reinterpret_cast<int*>(tmem_slot + 2 + kTileQ);  // resident-weight build: (offset, ky stride) per 4-channel group, <= 1 KB
  // PU_TC_DEBUG & 64: per-tile timeline of CTA 0 (cycles since kernel start): [event][tile < 12]
  long long* trace = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(tmem_slot) + 1024);  // 10 events x 12 tiles
  const bool tracing = (a.debug & 64) && blockIdx.x == 0 && blockIdx.y == 0;
  const long long t_start = clock64();
  auto stamp = [&](int ev, int k) {
    if (tracing && k < 12) trace[ev * 12 + k] = clock64() - t_start;
  };
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int st) { return bar0 + 8u * st; };
  auto empty_bar = [&](int st) { return bar0 + 8u * (kMaxStages + st); };
  auto tfull_bar = [&](int as) { return bar0 + 8u * (2 * kMaxStages + as); };
  auto tempty_bar = [&](int as) { return bar0 + 8u * (2 * kMaxStages + kMaxAcc + as); };
  auto tqfull_bar = [&](int q) { return bar0 + 8u * (2 * kMaxStages + 2 * kMaxAcc + q); };
  auto tqempty_bar = [&](int q) { return bar0 + 8u * (2 * kMaxStages + 2 * kMaxAcc + kTileQ + q); };
  // Dynamic tile scheduler.  The weight-gradient kernels of the backward pass run concurrently on side streams, so a
  // persistent CTA may get its SM late (or share the memory system unevenly); with a static round-robin assignment the
  // kernel then lasts as long as its unluckiest CTA.  Instead the producer warp draws tiles from a global counter (the first
  // tile is blockIdx.x, later ones gridDim.x + atomicAdd) and hands each index to the MMA and epilogue warps through a
  // small shared-memory queue guarded by mbarriers; -1 ends the kernel.  Consumers: kMmaWarps + kEpiWarps warps.
  // Measured on B200 (UNetp step, B = 64): the queue hand-off costs the latency-bound narrow layers ~4 % (1.24 -> 1.29 ms
  // per step) and the contention it was built for is not relieved by it (the side-stream kernels time-slice whole SMs), so
  // it is opt-in (PU_TC_DYNAMIC=1); the default is the static assignment.
  auto next_tile = [&](int k) -> int {  // consumer side: k-th tile of this CTA (every consumer warp calls it for every k in order)
    if (!a.dynamic) {
      const int t = (int)blockIdx.x + k * (int)gridDim.x;
      return t < a.tilesX * a.tilesY * a.B ? t : -1;
    }
    const int q = k % kTileQ;
    mbar_wait(tqfull_bar(q), (uint32_t)(k / kTileQ) & 1);
    const int t = tileq[q];
    __syncwarp();
    if (lane_id() == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tqempty_bar(q)) : "memory");
    return t;
  };

  // Programmatic dependent launch (PU_PDL=1): let the NEXT kernel's CTAs be scheduled as soon as ours exist; our own wait
  // for the predecessor grid (pdl_wait below) sits AFTER the prologue — barrier init, TMEM allocation, tensor-map prefetch,
  // the in-kernel weight-tile build and the bias load touch only parameters (written by the optimizer, many kernels ago),
  // so they overlap the predecessor's tail.  Every role waits before its first access to activation / gradient memory.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  auto pdl_wait = [&]() { asm volatile("griddepcontrol.wait;" ::: "memory"); };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int coblk = blockIdx.y;
  const int co_base = coblk * kCoBlk;
  const int tiles_per_img = a.tilesX * a.tilesY;
  const int ntiles = tiles_per_img * a.B;
  const int acc_cols = a.nmb * NBLK;  // TMEM columns of one accumulator buffer

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < kMaxStages; ++i) {
        mbar_init(full_bar(i), 1);
        mbar_init(empty_bar(i), kMmaWarps);  // one tcgen05.commit per MMA warp
      }
      for (int i = 0; i < kMaxAcc; ++i) {
        mbar_init(tfull_bar(i), kMmaWarps);
        mbar_init(tempty_bar(i), kEpiWarps);  // one arrival per epilogue warp
      }
      for (int i = 0; i < kTileQ; ++i) {
        mbar_init(tqfull_bar(i), 1);
        mbar_init(tqempty_bar(i), kMmaWarps + kEpiWarps);  // one arrival per consumer warp
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm0)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm1)) : "memory");
    }
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();  // barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) stamp(5, 0);

  if (a.wfmt != 0 && warp != 0) {
    // Resident weights: warps 1.. build the tf32 B-operand tiles of every K chunk straight from the OIHW tensor while
    // warp 0 already streams the first input tiles (a naive per-element build with runtime index arithmetic took
    // 6 us for a 37 KB image: more than the rest of a small layer).
    constexpr int kBuilders = kTcThreads - 32;
    const int bt = tid - 32;
    auto build_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kBuilders) : "memory"); };
    if (!(a.debug & 16)) {
      // k-group table: byte offset of group g's [n][4] block for ky = 0, and the ky stride, per 4-channel group
      for (int g = bt; g < a.Cin / 4; g += kBuilders) {
        const int ci = g * 4;
        int off = 0, kys = 0;
        for (int c = 0; c < a.nchunks; ++c) {
          const TcChunk& ch = a.chunks[c];
          int kk = 0;
          for (int r = 0; r < ch.nreg; ++r) {
            const int c_lo = (ch.reg[r].src ? a.C0 : 0) + ch.reg[r].c_off;
            if (ci >= c_lo && ci < c_lo + ch.reg[r].cb) {
              const int kc = (kk + ci - c_lo) / 4;
              off = (int)ch.w_off + (kc >> 1) * N3 * 32 + (kc & 1) * 16;
              kys = (ch.ncg >> 1) * N3 * 32;
            }
            kk += ch.reg[r].cb;
          }
        }
        wtab[2 * g] = off;
        wtab[2 * g + 1] = kys;
      }
      build_sync();
      // One thread per (output channel, 4-input-channel group): its 4 x 9 weights are 36 consecutive floats of the
      // OIHW tensor in the forward case (nine 128-bit loads), four 9-float runs for dgrad; every tap then is ONE
      // 16-byte shared-memory store, and consecutive lanes (consecutive co) write consecutive 16-byte slots.
      const int cob = (a.Cout - co_base) < kCoBlk ? (a.Cout - co_base) : kCoBlk;  // output channels of this co block
      const bool vec = (reinterpret_cast<uintptr_t>(a.wpk) & 15) == 0;
      constexpr int kColsShift = COLS == 8 ? 3 : (COLS == 16 ? 4 : (COLS == 32 ? 5 : 6));
      for (int q = bt; q < COLS * (a.Cin / 4); q += kBuilders) {
        const int co_l = q & (COLS - 1), g = q >> kColsShift;
        uint8_t* const dst = smWres + ((wtab[2 * g] + co_l * 32) ^ (((co_l >> 2) & 1) << 4));  // SWIZZLE_32B
        const int kys = wtab[2 * g + 1];
        float x[36];
#pragma unroll
        for (int j = 0; j < 36; ++j) x[j] = 0.f;
        if (co_l < cob) {
          if (a.wfmt == 1) {
            const float* src = a.wpk + ((size_t)(co_base + co_l) * a.Cin + 4 * g) * 9;
            if (vec) {
#pragma unroll
              for (int j = 0; j < 9; ++j) {
                const float4 t = ldg4(src + 4 * j);
                x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 36; ++j) x[j] = __ldg(src + j);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float* src = a.wpk + ((size_t)(4 * g + e) * a.Cout + (co_base + co_l)) * 9;
#pragma unroll
              for (int t = 0; t < 9; ++t) x[e * 9 + t] = __ldg(src + t);
            }
          }
        }
        if (a.wfmt == 1) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
            *reinterpret_cast<float4*>(dst + (tap / 3) * kys + (tap % 3) * COLS * 32) =
                make_float4(round_tf32(x[tap]), round_tf32(x[9 + tap]), round_tf32(x[18 + tap]), round_tf32(x[27 + tap]));
        } else {  // dgrad: taps flipped
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
            *reinterpret_cast<float4*>(dst + (tap / 3) * kys + (tap % 3) * COLS * 32) =
                make_float4(round_tf32(x[8 - tap]), round_tf32(x[17 - tap]), round_tf32(x[26 - tap]), round_tf32(x[35 - tap]));
        }
        if (3 * COLS < N3) {  // COLS == 8: N rows 24..31 of every (ky, k group) are padding
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) *reinterpret_cast<float4*>(dst + ky * kys + 3 * COLS * 32) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
    build_sync();
  }

  if (a.debug & 32) {
    // (experiment) prologue only
  } else if (warp == 0) {
    // ================= TMA producer =================
    pdl_wait();
    const uint8_t* wblk = reinterpret_cast<const uint8_t*>(a.wpk) + (size_t)coblk * a.w_coblk_stride;
    int tile_next = blockIdx.x;  // lane 0: the tile drawn one iteration ahead (the atomic's round trip hides behind the loads)
    for (int k = 0;; ++k) {
      // publish the tile drawn during the previous iteration to the consumer warps, and draw the one after it
      int tile;
      if (a.dynamic) {
        tile = __shfl_sync(0xffffffffu, tile_next, 0);
        if (tile >= ntiles) tile = -1;
        if (tile >= 0 && lane == 0) tile_next = (int)gridDim.x + (int)atomicAdd(a.tile_ctr + coblk, 1u);
      } else {
        tile = (int)blockIdx.x + k * (int)gridDim.x;
        if (tile >= ntiles) tile = -1;
      }
      if (a.dynamic) {
        const int q = k % kTileQ;
        mbar_wait(tqempty_bar(q), ((uint32_t)(k / kTileQ) & 1) ^ 1);  // passes immediately on a fresh barrier
        if (lane == 0) {
          tileq[q] = tile;
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tqfull_bar(q)) : "memory");
        }
        __syncwarp();
      }
      if (tile < 0) break;
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH;
      for (int c = 0; c < a.nchunks; ++c) {
        const int it = k * a.nchunks + c;
        const int st = it % nst;
        const uint32_t ph = (uint32_t)(it / nst) & 1;
        mbar_wait(empty_bar(st), ph ^ 1);  // passes immediately on a fresh barrier
        if (elect_one()) {
          const TcChunk& ch = a.chunks[c];
          const uint32_t w_bytes = (uint32_t)(3 * ch.ncg * N3 * 16);
          const uint32_t sS = smem_u32(smem + st * stage_bytes);
          if ((a.debug & 4) && k * a.nchunks + c >= nst) {
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
      }
      if (lane == 0) stamp(0, k);
    }
  } else if (warp < kEpiWarp0) {
    // ================= MMA issuers (warps 1-3) =================
    // The warps run the same warp-uniform loop in lock-step over the tiles (every warp consumes every barrier phase in
    // order) and split the 96-pixel blocks of a tile round-robin; one elected lane executes each tcgen05 instruction,
    // so the descriptors live in uniform registers.
    const int mw = warp - 1;
    // instruction descriptor: D=f32, A=B=tf32, K-major both, N = N3, M = 128 (cute::UMMA::InstrDescriptor)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NACC >> 3) << 17) | ((128u >> 4) << 24);
    const bool leader = elect_one();  // elected once: the issue loop must stay a handful of instructions per MMA
    int it = 0;
    for (int k = 0;; ++k) {
      if (next_tile(k) < 0) break;
      const int as = k % a.nacc;
      const uint32_t aph = (uint32_t)(k / a.nacc) & 1;
      mbar_wait(tempty_bar(as), aph ^ 1);  // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(as * acc_cols);
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const TcChunk& ch = a.chunks[c];
        const int st = it % nst;
        const uint32_t ph = (uint32_t)(it / nst) & 1;
        // Everything the issue loop needs from the chunk table goes into registers BEFORE the wait for the data: one thread
        // issues a tcgen05.mma only every ~128 clk and whatever sits between two issues (constant-bank loads of the chunk table,
        // descriptor bit arithmetic) adds to that — measured 211 clk per MMA on the narrow flat layers with the arithmetic inside
        // the loop.  A descriptor is affine in (ky, kx, k step, block): base descriptor + offset / 16.
        const uint32_t sS = smem_u32(smem + st * stage_bytes);
        const int nreg = ch.nreg, ncg_half = ch.ncg >> 1;
        uint64_t a0[2];
        uint32_t rbq[2];  // row bytes / 16: the descriptor step of one pixel row
        int ksn[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t rb = (uint32_t)ch.reg[r < nreg ? r : 0].cb * 4;  // row bytes = swizzle span
          const uint32_t layout = rb == 32 ? 6u : (rb == 64 ? 4u : 2u);
          // 8-row core groups: 6 rows apart when folded (rows 6,7 duplicate the next group's 0,1), canonical 8 when flat
          a0[r] = umma_desc(sS + ch.reg[r < nreg ? r : 0].off, 16, (FOLD ? 6 : 8) * rb, layout);
          rbq[r] = rb >> 4;
          ksn[r] = ch.reg[r < nreg ? r : 0].cb >> 3;
        }
        const uint64_t b_base = umma_desc(a.wfmt == 0 ? sS + a.a_bytes : smem_u32(smWres) + ch.w_off, 16, 256, 6);
        mbar_wait(full_bar(st), ph);
        tc_fence_after();
        if (mw == 0 && lane == 0 && c == 0) stamp(1, k);
        if (KS) {
          // tap-row split: this warp issues the three kx taps of tap row ky = mw for every K step and block (nmb <= 2: the two
          // blocks alternate tap by tap, consecutive MMAs into the same accumulator wait for each other)
          const int ky = mw;
          const bool two = a.nmb > 1;
          const uint32_t d = d0 + (uint32_t)(ky * NACC);
          const uint32_t dB = d + (uint32_t)NBLK;
          uint32_t kc = 0;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            if (r >= nreg) break;
            const uint64_t a_blk = a0[r] + (uint64_t)((uint32_t)(ky * a.PW) * rbq[r]);
            const uint64_t blk_step = (uint64_t)((uint32_t)kBlkPixFlat * rbq[r]);
            const uint32_t kxs = rbq[r];
            for (int ks = 0; ks < ksn[r]; ++ks, kc += 2) {
              const uint64_t bd0 = b_base + (uint64_t)((ky * ncg_half + (kc >> 1)) * N3 * 2);
              const uint64_t ad0 = a_blk + (uint64_t)(2 * ks);
              const uint32_t first = (c | (int)kc) ? 1u : 0u;
              if (leader && !(a.debug & 1)) {
                if (two) {
                  const uint64_t ad1 = ad0 + blk_step;
                  umma_tf32(d, ad0, bd0, idesc, first);
                  umma_tf32(dB, ad1, bd0, idesc, first);
                  umma_tf32(d, ad0 + kxs, bd0 + COLS * 2, idesc, 1u);
                  umma_tf32(dB, ad1 + kxs, bd0 + COLS * 2, idesc, 1u);
                  umma_tf32(d, ad0 + 2 * kxs, bd0 + 2 * COLS * 2, idesc, 1u);
                  umma_tf32(dB, ad1 + 2 * kxs, bd0 + 2 * COLS * 2, idesc, 1u);
                } else {
                  umma_tf32(d, ad0, bd0, idesc, first);
                  umma_tf32(d, ad0 + kxs, bd0 + COLS * 2, idesc, 1u);
                  umma_tf32(d, ad0 + 2 * kxs, bd0 + 2 * COLS * 2, idesc, 1u);
                }
              }
            }
          }
        } else if (!FOLD) {
          // flat: every MMA warp owns whole 128-pixel blocks (the planner keeps nmb <= kMmaWarps, so one each) and issues
          // the nine taps of each 8-channel K step back to back — immediate descriptor offsets, no per-MMA loop control
          // Consecutive MMAs into the SAME accumulator wait for each other in the tensor pipe, so a warp that owns two blocks
          // (narrow layers: nmb up to 6) alternates between them tap by tap (PU_TC_ILV=0: one block after the other)
          const int mb2 = mw + kMmaWarps;
          const bool two = mb2 < a.nmb && !(a.debug & 256);
          for (int mb = mw; mb < a.nmb; mb += (two ? 2 : 1) * kMmaWarps) {
            const uint32_t d = d0 + (uint32_t)(mb * NACC);
            const uint32_t dB = d + (uint32_t)(kMmaWarps * NACC);
            for (int ky = 0; ky < 3; ++ky) {
              uint32_t kc = 0;
#pragma unroll
              for (int r = 0; r < 2; ++r) {  // static indices: a0 / rbq / ksn stay in registers
                if (r >= nreg) break;
                const uint64_t a_blk = a0[r] + (uint64_t)((uint32_t)(ky * a.PW + mb * kBlkPixFlat) * rbq[r]);
                const uint64_t blk_step = (uint64_t)((uint32_t)(kMmaWarps * kBlkPixFlat) * rbq[r]);
                const uint32_t kxs = rbq[r];
                for (int ks = 0; ks < ksn[r]; ++ks, kc += 2) {
                  const uint64_t bd0 = b_base + (uint64_t)((ky * ncg_half + (kc >> 1)) * N3 * 2);
                  const uint64_t ad0 = a_blk + (uint64_t)(2 * ks);
                  const uint32_t first = (c | ky | (int)kc) ? 1u : 0u;
                  if (leader && !(a.debug & 1)) {
                    if (two) {
                      const uint64_t ad1 = ad0 + blk_step;
                      umma_tf32(d, ad0, bd0, idesc, first);
                      umma_tf32(dB, ad1, bd0, idesc, first);
                      umma_tf32(d, ad0 + kxs, bd0 + COLS * 2, idesc, 1u);
                      umma_tf32(dB, ad1 + kxs, bd0 + COLS * 2, idesc, 1u);
                      umma_tf32(d, ad0 + 2 * kxs, bd0 + 2 * COLS * 2, idesc, 1u);
                      umma_tf32(dB, ad1 + 2 * kxs, bd0 + 2 * COLS * 2, idesc, 1u);
                    } else {
                      umma_tf32(d, ad0, bd0, idesc, first);
                      umma_tf32(d, ad0 + kxs, bd0 + COLS * 2, idesc, 1u);
                      umma_tf32(d, ad0 + 2 * kxs, bd0 + 2 * COLS * 2, idesc, 1u);
                    }
                  }
                }
              }
            }
          }
        } else
        for (int ky = 0; ky < 3; ++ky) {
          uint32_t kc = 0;
#pragma unroll
          for (int r = 0; r < 2; ++r) {  // static indices: a0 / rbq / ksn stay in registers
            if (r >= nreg) break;
            const uint64_t a_base = a0[r] + (uint64_t)((uint32_t)(ky * a.PW) * rbq[r]);
            const uint32_t mb_step = 6 * 16 * rbq[r];  // BLK rows (96 when folded), in 16-byte descriptor units: 6 * rb
            for (int ks = 0; ks < ksn[r]; ++ks, kc += 2) {
              const uint64_t bd = b_base + (uint64_t)((ky * ncg_half + (kc >> 1)) * N3 * 2);
              uint64_t ad = a_base + (uint64_t)(2 * ks + mw * mb_step);
              uint32_t d = d0 + mw * NACC;
              const uint32_t acc = (c | ky | (int)kc) ? 1u : 0u;
              // blocks innermost: consecutive MMAs write different accumulators
#pragma unroll 3
              for (int mb = mw; mb < a.nmb; mb += kMmaWarps) {
                if (leader && !(a.debug & 1)) umma_tf32(d, ad, bd, idesc, acc);
                d += kMmaWarps * NACC;
                ad += kMmaWarps * mb_step;
              }
            }
          }
        }
        if (leader) tc_commit(empty_bar(st));  // frees the stage once these MMAs have read it
        __syncwarp();
      }
      if (leader) tc_commit(tfull_bar(as));  // accumulators complete
      __syncwarp();
      if (mw == 0 && lane == 0) stamp(2, k);
    }
  } else {
    // ================= epilogue: TMEM -> registers -> kx realignment -> bias/residual/ReLU/mask -> NHWC global =================
    // warp e handles TMEM lane quarter (warp % 4) of the MMA blocks mb = set, set + kEpiSets, ... (set = e / 4).
    // A unit is 8 output channels of one block: three tcgen05.ld.x8 (the kx = 0,1,2 column groups), two shuffles per
    // channel.  Two units are fetched per tcgen05.wait::ld and their residual / mask vectors are requested before the
    // wait.  The loop is issue-bound (measured: ~170 SASS instructions per unit in the first version), so everything
    // per-tile is warp-uniform, per-pixel offsets are 32-bit, (yy, xx) advance without a division and the bias of an
    // 8-channel co block lives in registers.
    constexpr int NQ = COLS / 8;
    const int quarter = warp & 3;
    const int set = (warp - kEpiWarp0) >> 2;
    const int R = quarter * 32 + lane;
    const int gi = R & 7;                     // row inside its 8-row group; folded: rows 6,7 duplicate the next group's 0,1
    const int prow = FOLD ? (R >> 3) * 6 + gi : R;  // pixel offset of this row inside an MMA block
    const int p0 = set * BLK + prow;
    const int yy0 = p0 / a.PW, xx0 = p0 - yy0 * a.PW;
    const int step_y = (kEpiSets * BLK) / a.PW, step_x = kEpiSets * BLK - step_y * a.PW;  // to this warp's next block
    const bool has_mask = a.mask0 != nullptr || a.mask1 != nullptr;
    const bool has_res = a.res != nullptr && !has_mask;
    const bool has_bias = a.bias != nullptr;
    const bool bias32 = (reinterpret_cast<uintptr_t>(a.bias) & 31) == 0;  // parameters may sit in a packed arena
    const int relu = a.relu, round_out = a.round_out;
    const int nq_valid = (a.Cout - co_base + 7) / 8 < NQ ? (a.Cout - co_base + 7) / 8 : NQ;  // 8-channel groups of this co block
    const bool live = (!FOLD || gi < 6) && !(a.debug & 2);
    auto load_bias = [&](int co, float* bb) {
      if (bias32) {
        ldg8(a.bias + co, bb);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = __ldg(a.bias + co + j);
      }
    };
    float bias0[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bias0[j] = 0.f;
    if (NQ == 1 && has_bias) load_bias(co_base, bias0);

    // kx realignment + fused pointwise tail of one unit.  mbits: packed ReLU mask of the destination's 8 channels (dgrad;
    // 0xff = keep everything); mdst: where to store the packed mask of this unit's own output (forward; may be null)
    auto finish = [&](const uint32_t* v, bool ok, float* dst, bool has_aux, const float* aux, const float* bb, uint32_t mbits,
                      unsigned char* mdst) {
      float o[8];
      if (FOLD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float e1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[8 + j]), 1);
          const float e2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[16 + j]), 2);
          o[j] = (__uint_as_float(v[j]) + e1) + e2;
        }
      } else if (KS) {  // the three tap-row partials
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (__uint_as_float(v[j]) + __uint_as_float(v[8 + j])) + __uint_as_float(v[16 + j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[j]);
      }
      if (!ok) return;
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += bb[j];
      }
      if (has_aux) {  // residual
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += aux[j];
      }
      if (has_mask) {  // dgrad: ReLU mask of the layer that produced this source
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (mbits >> j) & 1u ? o[j] : 0.f;
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
      }
      if (round_out) {  // round-to-nearest (ties away) to TF32, same result as cvt.rna.tf32.f32 for finite values
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __uint_as_float((__float_as_uint(o[j]) + 0x1000u) & 0xffffe000u);
      }
      stg8(dst, o);
      if (mdst != nullptr) {
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) m |= (o[j] > 0.f ? 1u : 0u) << j;
        *mdst = (unsigned char)m;
      }
    };
    // this warp's MMA blocks of a tile are mb = set + i * kEpiSets, i < kMaxI (compile-time bound so that the packed-mask
    // bytes prefetched before the accumulators are ready stay in registers)
    constexpr int kMaxI = (256 / NACC + kEpiSets - 1) / kEpiSets;

    pdl_wait();
    uint32_t tcount = 0;
    for (;; ++tcount) {
      const int tile = next_tile((int)tcount);
      if (tile < 0) break;
      const int b = tile / tiles_per_img;
      const int tr = tile - b * tiles_per_img;
      const int ty = tr / a.tilesX, tx = tr - ty * a.tilesX;
      const int x0 = tx * a.TW, y0 = ty * a.TH;
      const int ymax = (a.H - y0) < a.TH ? (a.H - y0) : a.TH, xmax = (a.W - x0) < a.TW ? (a.W - x0) : a.TW;
      // warp-uniform bases of this tile in every tensor the epilogue touches (masks share their destination's geometry)
      const size_t o0 = (((size_t)b * a.d0.Hs + (y0 + a.d0.oy)) * a.d0.Ws + (x0 + a.d0.ox)) * a.d0.C;
      const size_t o1 = a.d1.p == nullptr ? 0 : (((size_t)b * a.d1.Hs + (y0 + a.d1.oy)) * a.d1.Ws + (x0 + a.d1.ox)) * a.d1.C;
      float* const d0b = a.d0.p + o0;
      float* const d1b = a.d1.p == nullptr ? nullptr : a.d1.p + o1;
      const unsigned char* const m0b = a.mask0 == nullptr ? nullptr : a.mask0 + o0 / 8;  // o0 is a multiple of d0.C, d0.C of 8
      const unsigned char* const m1b = a.mask1 == nullptr ? nullptr : a.mask1 + o1 / 8;
      unsigned char* const mob = a.mask_out == nullptr ? nullptr : a.mask_out + o0 / 8 + (co_base >> 3);
      const float* const rsb = has_res ? a.res + (((size_t)b * a.H + y0) * a.W + x0) * a.Cout + co_base : nullptr;
      const int as = (int)(tcount % (uint32_t)a.nacc);
      // ---- packed ReLU masks of this tile's units: requested BEFORE waiting for the accumulators, so that their global
      // latency hides behind the MMAs (one byte per unit instead of re-reading 32 bytes of fp32 activation per unit)
      uint32_t mk[kMaxI][NQ];
      if (has_mask) {
        int yy_ = yy0, xx_ = xx0;
#pragma unroll
        for (int i = 0; i < kMaxI; ++i) {
          const bool ok_ = live && (set + i * kEpiSets) < a.nmb && yy_ < ymax && xx_ < xmax;
          const int p0_ = (yy_ * a.d0.Ws + xx_) * (a.d0.C >> 3);
          const int p1_ = a.d1.p == nullptr ? 0 : (yy_ * a.d1.Ws + xx_) * (a.d1.C >> 3);
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int co = co_base + 8 * q;
            const unsigned char* mp = co < a.d0.C ? (m0b == nullptr ? nullptr : m0b + p0_ + (co >> 3))
                                                  : (m1b == nullptr ? nullptr : m1b + p1_ + ((co - a.d0.C) >> 3));
            mk[i][q] = (ok_ && mp != nullptr && q < nq_valid) ? (uint32_t)__ldg(mp) : 0xffu;
          }
          yy_ += step_y;
          xx_ += step_x;
          if (xx_ >= a.PW) { xx_ -= a.PW; ++yy_; }
        }
      }
      mbar_wait(tfull_bar(as), (tcount / (uint32_t)a.nacc) & 1);
      tc_fence_after();
      if (warp == kEpiWarp0 && lane == 0) stamp(3, tcount);
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * acc_cols);
      int yy = yy0, xx = xx0;
      auto advance = [&]() {
        yy += step_y;
        xx += step_x;
        if (xx >= a.PW) { xx -= a.PW; ++yy; }
      };
      if (a.debug & 128) {
        // (experiment) no TMEM reads at all
      } else if (NQ == 1) {
        // 8-channel co block: everything goes to d0; two of this warp's blocks per TMEM wait
        const bool has_aux = has_res;
        const float* const xb = rsb;
        const int xs_y = a.W * a.Cout, xs_x = a.Cout;
#pragma unroll
        for (int i = 0; i < kMaxI; i += 2) {
          const int mb = set + i * kEpiSets;
          if (mb >= a.nmb) break;
          uint32_t v[2][24];
          float aux[2][8];
          const bool okA = live && yy < ymax && xx < xmax;
          const int pixA = yy * a.d0.Ws + xx, xoA = yy * xs_y + xx * xs_x;
          advance();
          const bool haveB = (i + 1 < kMaxI) && mb + kEpiSets < a.nmb;
          const bool okB = haveB && live && yy < ymax && xx < xmax;
          const int pixB = yy * a.d0.Ws + xx, xoB = yy * xs_y + xx * xs_x;
          advance();
          const uint32_t tA = tbase + (uint32_t)(mb * NBLK);
          tmem_ld8(tA, v[0]);
          if (FOLD || KS) {
            tmem_ld8(tA + COLS, v[0] + 8);
            tmem_ld8(tA + 2 * COLS, v[0] + 16);
          }
          if (haveB) {
            const uint32_t tB = tA + (uint32_t)(kEpiSets * NBLK);
            tmem_ld8(tB, v[1]);
            if (FOLD || KS) {
              tmem_ld8(tB + COLS, v[1] + 8);
              tmem_ld8(tB + 2 * COLS, v[1] + 16);
            }
          }
          if (has_aux) {
            if (okA) ldg8(xb + xoA, aux[0]);
            if (okB) ldg8(xb + xoB, aux[1]);
          }
          tmem_ld_wait();
          finish(v[0], okA, d0b + pixA * a.d0.C, has_aux, aux[0], bias0, has_mask ? mk[i][0] : 0xffu, mob == nullptr ? nullptr : mob + pixA);
          if (haveB)
            finish(v[1], okB, d0b + pixB * a.d0.C, has_aux, aux[1], bias0, has_mask ? mk[(i + 1) % kMaxI][0] : 0xffu,
                   mob == nullptr ? nullptr : mob + pixB);
        }
      } else {
#pragma unroll
        for (int i = 0; i < kMaxI; ++i) {
          const int mb = set + i * kEpiSets;
          if (mb >= a.nmb) break;
          const bool ok = live && yy < ymax && xx < xmax;
          const int pix0 = yy * a.d0.Ws + xx;
          const int off0 = pix0 * a.d0.C;
          const int off1 = d1b == nullptr ? 0 : (yy * a.d1.Ws + xx) * a.d1.C;
          const int offr = (yy * a.W + xx) * a.Cout;
          advance();
#pragma unroll
          for (int q = 0; q < NQ; q += 2) {
            uint32_t v[2][24];
            float aux[2][8], bb[2][8];
            const uint32_t tA = tbase + (uint32_t)(mb * NBLK + 8 * q);
            tmem_ld8(tA, v[0]);
            tmem_ld8(tA + 8, v[1]);
            if (FOLD || KS) {
              tmem_ld8(tA + COLS, v[0] + 8);
              tmem_ld8(tA + 2 * COLS, v[0] + 16);
              tmem_ld8(tA + 8 + COLS, v[1] + 8);
              tmem_ld8(tA + 8 + 2 * COLS, v[1] + 16);
            }
            float* dst[2];
            bool okq[2];
            unsigned char* mdst[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int co = co_base + 8 * (q + k);  // warp-uniform: which destination tensor this channel group belongs to
              const bool first = co < a.d0.C;
              okq[k] = ok && (q + k) < nq_valid;
              dst[k] = first ? d0b + off0 + co : d1b + off1 + (co - a.d0.C);
              mdst[k] = mob == nullptr ? nullptr : mob + pix0 * (a.d0.C >> 3) + (q + k);
              if (has_res && okq[k]) ldg8(rsb + offr + 8 * (q + k), aux[k]);
              if (has_bias && okq[k]) load_bias(co, bb[k]);
            }
            tmem_ld_wait();
            finish(v[0], okq[0], dst[0], has_res, aux[0], bb[0], has_mask ? mk[i][q] : 0xffu, mdst[0]);
            finish(v[1], okq[1], dst[1], has_res, aux[1], bb[1], has_mask ? mk[i][(q + 1) % NQ] : 0xffu, mdst[1]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(as)) : "memory");
      if (warp == kEpiWarp0 && lane == 0) stamp(4, tcount);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tracing && tid == 0) {
    const long long t_end = clock64() - t_start;
    const int nt = (ntiles - 1) / (int)gridDim.x + 1;
    printf("tc trace (cycles): kernel %lld (prologue %lld), %d tiles on CTA 0; per tile: loads issued | data landed | MMAs issued | acc ready | epilogue done\n", t_end, trace[60], nt);
    for (int k = 0; k < nt && k < 12; ++k)
      printf("  tile %2d: %7lld %7lld %7lld %7lld %7lld \n", k, trace[k], trace[12 + k], trace[24 + k], trace[36 + k], trace[48 + k]);
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
  }
  if (tid == 0 && a.dynamic) {
    // this CTA draws no more tiles (its producer, this very thread, is done): the last CTA to get here re-zeroes the slot
    __threadfence();
    const unsigned total = gridDim.x * gridDim.y;
    if (atomicAdd(a.done_ctr, 1u) == total - 1) {
      for (unsigned i = 0; i < gridDim.y; ++i) a.tile_ctr[i] = 0u;
      __threadfence();
      *a.done_ctr = 0u;
    }
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
// tile-scheduler counters: kTcSlots launch slots x (kMaxCoBlk tile counters + 1 done counter), used round-robin by
// successive launches (a captured graph keeps the slot of each of its kernel nodes; a slot is re-zeroed by its own kernel)
constexpr int kTcSlots = 256;
__device__ unsigned int g_tc_counters[kTcSlots * (kMaxCoBlk + 1)];
static unsigned int* g_tc_counters_dev = nullptr;
static std::atomic<unsigned> g_tc_slot{0};
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
  void* ctr = nullptr;
  if (cudaGetSymbolAddress(&ctr, g_tc_counters) != cudaSuccess || cudaMemset(ctr, 0, sizeof(unsigned int) * kTcSlots * (kMaxCoBlk + 1)) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  g_tc_counters_dev = reinterpret_cast<unsigned int*>(ctr);
  g_tc_state = 1;
  return true;
}

struct TcPlan {
  int TH, TW, PW, tilesX, tilesY, nmb, cols, n3, a_bytes, w_bytes_max, w_res_bytes, tmem_cols, nchunks, ncoblk, nstages, cb0, cb1;
  int nacc;  // accumulator buffers
  int fold;  // 1: kx folded into N (96-pixel blocks); 0: flat, one MMA per tap (128-pixel blocks), wide layers
  int ksplit;  // flat mode, tiles of <= 2 blocks: the MMA warps split the tap rows of a block (three partial accumulators)
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

// When to use the flat mode (one MMA per tap).  PU_TC_FLAT=0/1 forces it off/on.
static bool tc_want_flat(long long npix, int C0, int C1, int Cout) {
  if (const char* e = getenv("PU_TC_FLAT")) return atoi(e) != 0;
  if (Cout >= 64 && Cout % 64 == 0 && C0 % 16 == 0 && C1 % 16 == 0 && C0 + C1 >= 64 && ((npix + 383) / 384) * (Cout / 64) >= 96)
    return true;  // wide, tensor-bound layers with enough (384-pixel tile, co block) work items to fill the SMs
  // narrow, HBM-bound layers: the flat epilogue does ~0.45x the work per pixel (no kx realignment, no duplicated rows) but the
  // tile needs 3x the MMAs, and one thread issues a tcgen05.mma only every ~128 clk: measured on B200 (B = 64) flat wins
  // for 16..64 output channels when C_in <= C_out (8->16 @64^2: 18.4 -> 14.0 us, 32->64 @16^2: 14.6 -> 11.9 us) and loses
  // for 8 output channels (N padded to 16) and for the concat layers (8|8->8 @128^2: 31.9 -> 44.7 us)
  return Cout >= 16 && Cout <= 64 && C0 + C1 <= Cout;
}

// channel plan: K chunks, weight image layout.  false if the channel counts do not fit the tensor-core path.
static bool tc_plan_channels(int C0, int C1, int Cout, TcPlan* p, bool flat = false) {
  if (C0 < 8 || C0 % 8 != 0 || C1 < 0 || C1 % 8 != 0 || Cout < 8 || Cout % 8 != 0) return false;
  const int cout_blk = Cout < kCoBlk ? Cout : kCoBlk;
  p->cols = cout_blk <= 8 ? 8 : (cout_blk <= 16 ? 16 : (cout_blk <= 32 ? 32 : 64));
  p->n3 = tc_n3(p->cols);
  p->ncoblk = (Cout + kCoBlk - 1) / kCoBlk;
  p->fold = flat ? 0 : 1;
  // flat, wide layers: 16-channel K chunks keep a 384-pixel stage (35 KB of pixels + 37 KB of weights) small enough for two stages
  const bool cap16 = flat && Cout >= 64;
  p->cb0 = region_channels(C0);
  p->cb1 = C1 > 0 ? region_channels(C1) : 0;
  if (cap16 && p->cb0 > 16) p->cb0 = 16;
  if (cap16 && p->cb1 > 16) p->cb1 = 16;
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
static bool tc_plan1(int B, int H, int W, int C0, int C1, int Cout, TcPlan* p, bool resident, bool want_flat);

// flat < 0: flat mode where tc_want_flat says so and the plan fits, else folded
static bool tc_plan(int B, int H, int W, int C0, int C1, int Cout, TcPlan* p, bool resident = false, int flat = -1) {
  const bool want_flat = flat < 0 ? tc_want_flat((long long)B * H * W, C0, C1, Cout) : (flat != 0);
  if (want_flat && tc_plan1(B, H, W, C0, C1, Cout, p, resident, true)) return true;
  return tc_plan1(B, H, W, C0, C1, Cout, p, resident, false);
}

static bool tc_plan1(int B, int H, int W, int C0, int C1, int Cout, TcPlan* p, bool resident, bool want_flat) {
  if (!tc_plan_channels(C0, C1, Cout, p, want_flat)) return false;
  const int blk = p->fold ? kBlkPix : kBlkPixFlat;
  const int tail = p->fold ? 98 : 130;  // rows the last block's shifted reads touch beyond its first row: 127 + kx 2 + 1 (flat)
  int max_ncg = 0;
  for (int i = 0; i < p->nchunks; ++i) max_ncg = p->chunks[i].ncg > max_ncg ? p->chunks[i].ncg : max_ncg;
  p->w_bytes_max = ((3 * max_ncg * p->n3 * 16) + 1023) / 1024 * 1024;
  p->w_res_bytes = 0;
  if (resident) {  // weights built in shared memory once per CTA: no per-stage weight slot
    if (p->w_coblk_stride > kMaxResidentW) return false;
    p->w_res_bytes = (int)((p->w_coblk_stride + 127) / 128 * 128);
    p->w_bytes_max = 0;
  }
  const size_t budget = 220 * 1024 - (size_t)p->w_res_bytes;  // minus barriers and the 1024-byte alignment slack below
  // two accumulator buffers of <= 256 TMEM columns; flat: one 128-pixel block per MMA-issuing warp (balanced issue)
  const int nacc = p->fold ? p->n3 : (p->cols < 16 ? 16 : p->cols);
  const int flat_cap = p->cols >= 64 ? kMmaWarps : 2 * kMmaWarps;  // wide: one block per issuing warp; narrow (HBM-bound): two
  const int nmb_max = p->fold ? 256 / nacc : (256 / nacc < flat_cap ? 256 / nacc : flat_cap);
  auto stage_a_bytes = [&](int th, int pw, int nmb) {
    // rows a region must hold: the halo tile, and whatever the last MMA block's shifted reads touch beyond it
    int rows = (th + 2) * pw;
    const int reach = (nmb - 1) * blk + tail + 2 * pw;
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
  // per-tile hand-off cost in pixel-row units (measured: ~1.8 k clk of barrier waits, descriptor set-up and drain per tile
  // against ~2.2 clk per MMA row)
  int tile_fixed = 800;
  if (const char* e = getenv("PU_TC_TILE_FIXED")) tile_fixed = atoi(e);
  const int tx_min = (W + 253) / 254;
  for (int tilesX = tx_min; tilesX <= tx_min + 7; ++tilesX) {
    const int tw = (W + tilesX - 1) / tilesX;
    if (tilesX > tx_min && tw < 16) break;
    const int pw = tw + 2;
    for (int th = (H < 254 ? H : 254); th >= 1; --th) {
      const int nmb = (th * pw + blk - 1) / blk;
      if (nmb > nmb_max) continue;
      const size_t stage = stage_a_bytes(th, pw, nmb) + p->w_bytes_max;
      if (2 * stage > budget) continue;
      const long long ntiles = (long long)B * ((H + th - 1) / th) * ((W + tw - 1) / tw);
      const int ctas_x = (kNumSMs + p->ncoblk - 1) / p->ncoblk;  // the co blocks share the SMs: grid = (ctas_x, ncoblk)
      const long long waves = (ntiles + ctas_x - 1) / ctas_x;
      // per-SM time ~ waves x (MMA rows + staged rows + a fixed per-tile hand-off cost), in units of one pixel row
      const double cost = (double)waves * (nmb * 128 + (th + 2) * pw + tile_fixed);
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
  p->nmb = (p->TH * p->PW + blk - 1) / blk;
  int rows = (p->TH + 2) * p->PW;
  const int reach = (p->nmb - 1) * blk + tail + 2 * p->PW;
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
  {
    // as many accumulator buffers as TMEM holds (2..kMaxAcc): the chain MMA issue -> completion -> epilogue -> release of a tile is
    // latency-bound, more tiles in flight hide it (PU_TC_NACC overrides)
    int per_buf = p->nmb * (p->fold ? p->n3 : (p->cols < 16 ? 16 : p->cols));
    // tap-row split (kernel template KS): tiles of one or two 128-pixel blocks of a >= 32-channel flat layer — the deep 16x16 / 8x8
    // levels, whose time is the MMA issue chain of ONE thread per block.  Three partial accumulators per block; a single
    // accumulator buffer is enough when every CTA has one tile.
    p->ksplit = 0;
    {
      static int ks_env = -1;
      if (ks_env < 0) {
        const char* e = getenv("PU_TC_KSPLIT");
        ks_env = (e != nullptr && e[0] == '0') ? 0 : 1;
      }
      const long long ntiles = (long long)B * p->tilesX * p->tilesY;
      const int ctas_x = (kNumSMs + p->ncoblk - 1) / p->ncoblk;
      if (ks_env && !p->fold && p->cols >= 32 && p->nmb <= 2 && (2 * 3 * per_buf <= 512 || (3 * per_buf <= 512 && ntiles <= ctas_x))) {
        p->ksplit = 1;
        per_buf *= 3;
      }
    }
    int nacc = 512 / per_buf;
    if (nacc > kMaxAcc) nacc = kMaxAcc;
    if (const char* e = getenv("PU_TC_NACC")) nacc = atoi(e) < nacc ? atoi(e) : nacc;
    if (nacc < 2 && !p->ksplit) nacc = 2;
    if (nacc < 1) nacc = 1;
    p->nacc = nacc;
    p->tmem_cols = next_pow2_cols(nacc * per_buf);
  }
  p->smem_bytes = (size_t)p->nstages * stage + p->w_res_bytes + 256 + 2048 + 1024;  // barriers, k-group table + trace, alignment
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

// A 4-D fp32 tensor map over an (oy, ox)-offset H x W window of an NHWC tensor, SWIZZLE_NONE, box = (cb, bw, bh, bn): used by the
// TMA-fed weight-gradient kernel (conv3x3_wgrad_tma.cu); out-of-window coordinates zero-fill.
int tma_make_window_map(CUtensorMap* tm, const View& v, int B, int H, int W, int cb, int bw, int bh, int bn) {
  if (!tc_init()) {
    set_error("tensor maps are unavailable on this device (cuTensorMapEncodeTiled)");
    return PU_ERR_UNSUPPORTED;
  }
  const float* base = v.p + ((size_t)v.oy * v.Ws + v.ox) * v.C;
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)v.C * 4, (cuuint64_t)v.Ws * v.C * 4, (cuuint64_t)v.Hs * v.Ws * v.C * 4};
  cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for window %dx%d C=%d box (%d,%d,%d,%d)", (int)r, H, W, v.C, cb, bw, bh, bn);
    return PU_ERR_CUDA;
  }
  return PU_OK;
}

// The same window for an 8-channel tensor with (channel, x) merged into ONE dimension of 8 W elements: a box row is then
// 32 (bw) contiguous bytes instead of bw separate 32-byte rows (measured: 32-byte box rows cap the TMA feed at ~4.5 TB/s).
// 3-D: (8 W, H, B), box (8 bw, bh, bn); out-of-window elements of the merged dimension zero-fill exactly like x < 0 / x >= W.
int tma_make_window_map_merged(CUtensorMap* tm, const View& v, int B, int H, int W, int bw, int bh, int bn) {
  if (!tc_init()) {
    set_error("tensor maps are unavailable on this device (cuTensorMapEncodeTiled)");
    return PU_ERR_UNSUPPORTED;
  }
  if (v.C != 8 || 8 * bw > 256) {
    set_error("tma_make_window_map_merged: needs C == 8 and a box row of <= 256 elements (C %d, box width %d)", v.C, bw);
    return PU_ERR_BAD_ARG;
  }
  const float* base = v.p + ((size_t)v.oy * v.Ws + v.ox) * 8;
  cuuint64_t dims[3] = {(cuuint64_t)8 * W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)v.Ws * 32, (cuuint64_t)v.Hs * v.Ws * 32};
  cuuint32_t box[3] = {(cuuint32_t)(8 * bw), (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for merged window %dx%d box (%d,%d,%d)", (int)r, H, W, 8 * bw, bh, bn);
    return PU_ERR_CUDA;
  }
  return PU_OK;
}

// A 3-D fp32 tensor map over the H x W window of a ONE-channel NHWC tensor ([B, Hs, Ws]), box = (bw, bh, bn): the stem's input
// for its TMA-fed weight gradient.  Needs a 16-byte aligned window origin and row pitch (PU_ERR_UNSUPPORTED otherwise).
int tma_make_plane_map(CUtensorMap* tm, const View& v, int B, int H, int W, int bw, int bh, int bn) {
  if (!tc_init()) {
    set_error("tensor maps are unavailable on this device (cuTensorMapEncodeTiled)");
    return PU_ERR_UNSUPPORTED;
  }
  const float* base = v.p + (size_t)v.oy * v.Ws + v.ox;
  if (v.C != 1 || (reinterpret_cast<uintptr_t>(base) & 15u) != 0 || v.Ws % 4 != 0 || bw % 4 != 0 || bw > 256) {
    set_error("tma_make_plane_map: needs C == 1, a 16-byte aligned window and row pitch, and a box row of 4..256 elements");
    return PU_ERR_UNSUPPORTED;
  }
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)v.Ws * 4, (cuuint64_t)v.Hs * v.Ws * 4};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for plane window %dx%d box (%d,%d,%d)", (int)r, H, W, bw, bh, bn);
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

bool conv3x3_tc_flat(int B, int H, int W, int C0, int C1, int Cout) {
  TcPlan p;
  return tc_init() && tc_plan(B, H, W, C0, C1, Cout, &p, false) && !p.fold;
}

long long conv3x3_tc_weight_floats(int C0, int C1, int Cout, bool flat) {
  TcPlan p;
  if (!tc_plan_channels(C0, C1, Cout, &p, flat)) return 0;
  return (long long)p.w_coblk_stride * p.ncoblk / 4;
}

int conv3x3_tc_pack(const float* w_oihw, float* out, int Cout_w, int Cin_w, int transpose, int C0, bool flat, cudaStream_t st) {
  // roles of the conv that will consume the packed weights
  const int cin = transpose ? Cout_w : Cin_w;
  const int cout = transpose ? Cin_w : Cout_w;
  const int c0 = transpose ? Cout_w : C0;
  TcPlan p;
  if (!tc_plan_channels(c0, cin - c0, cout, &p, flat)) {
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

template <int COLS, bool FOLD = true, bool KS = false>
static int launch_tc(const CUtensorMap& tm0, const CUtensorMap& tm1, const TcArgs& ta, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<COLS, FOLD, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("conv3x3_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaError_t le = launch_pdl(conv3x3_tc_kernel<COLS, FOLD, KS>, grid, dim3(kTcThreads), smem, st, tm0, tm1, ta);
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
  ta.mask0 = a.mask0; ta.mask1 = a.mask1; ta.mask_out = a.mask_out;
  ta.B = a.B; ta.H = a.H; ta.W = a.W; ta.Cout = a.Cout; ta.relu = a.relu; ta.round_out = a.round_out;
  ta.TH = p.TH; ta.TW = p.TW; ta.PW = p.PW; ta.tilesX = p.tilesX; ta.tilesY = p.tilesY;
  ta.nmb = p.nmb; ta.a_bytes = p.a_bytes; ta.w_bytes_max = p.w_bytes_max; ta.nstages = p.nstages;
  ta.tmem_cols = p.tmem_cols; ta.nchunks = p.nchunks; ta.w_coblk_stride = p.w_coblk_stride;
  ta.nacc = p.nacc;
  ta.wfmt = a.wfmt; ta.Cin = a.Cin; ta.C0 = a.s0.C; ta.w_res_bytes = p.w_res_bytes;
  if (p.ncoblk > kMaxCoBlk) {
    set_error("pu_conv3x3_fwd: more than %d output-channel blocks (Cout %d)", kMaxCoBlk, a.Cout);
    return PU_ERR_UNSUPPORTED;
  }
  {
    static int dyn = -1;
    if (dyn < 0) {
      const char* e = getenv("PU_TC_DYNAMIC");
      dyn = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    ta.dynamic = dyn;
    const unsigned slot = g_tc_slot.fetch_add(1) % kTcSlots;
    ta.tile_ctr = g_tc_counters_dev + (size_t)slot * (kMaxCoBlk + 1);
    ta.done_ctr = ta.tile_ctr + kMaxCoBlk;
  }
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
  const int ctas_x = (kNumSMs + p.ncoblk - 1) / p.ncoblk;  // one persistent CTA per SM in total
  dim3 grid(ntiles < ctas_x ? ntiles : ctas_x, p.ncoblk);
  if (!p.fold && p.ksplit) {
    if (p.cols == 32) return launch_tc<32, false, true>(tm0, tm1, ta, grid, p.smem_bytes, st);
    return launch_tc<64, false, true>(tm0, tm1, ta, grid, p.smem_bytes, st);
  }
  if (!p.fold) {
    switch (p.cols) {
      case 8: return launch_tc<8, false>(tm0, tm1, ta, grid, p.smem_bytes, st);
      case 16: return launch_tc<16, false>(tm0, tm1, ta, grid, p.smem_bytes, st);
      case 32: return launch_tc<32, false>(tm0, tm1, ta, grid, p.smem_bytes, st);
      default: return launch_tc<64, false>(tm0, tm1, ta, grid, p.smem_bytes, st);
    }
  }
  switch (p.cols) {
    case 8: return launch_tc<8>(tm0, tm1, ta, grid, p.smem_bytes, st);
    case 16: return launch_tc<16>(tm0, tm1, ta, grid, p.smem_bytes, st);
    case 32: return launch_tc<32>(tm0, tm1, ta, grid, p.smem_bytes, st);
    default: return launch_tc<64>(tm0, tm1, ta, grid, p.smem_bytes, st);
  }
}

}  // namespace pu

extern "C" int pu_tc_available(void) { return pu::tc_init() ? 1 : 0; }

extern "C" int pu_conv3x3_tc_flat(int B, int H, int W, int C0, int C1, int Cout) { return pu::conv3x3_tc_flat(B, H, W, C0, C1, Cout) ? 1 : 0; }

// Host-only: the tile plan pu_conv3x3_fwd would use (no device needed; for tests and tuning).
extern "C" int pu_conv3x3_tc_plan(int B, int H, int W, int C0, int C1, int Cout, int resident, int* out17) {
  if (out17 == nullptr || B <= 0 || H <= 0 || W <= 0) {
    pu::set_error("pu_conv3x3_tc_plan: bad argument");
    return PU_ERR_BAD_ARG;
  }
  pu::TcPlan p;
  if (!pu::tc_plan(B, H, W, C0, C1, Cout, &p, resident != 0)) {
    pu::set_error("pu_conv3x3_tc_plan: shape (C %d|%d -> %d, %dx%d) does not fit the tcgen05 path%s", C0, C1, Cout, H, W,
                  resident ? " with resident weights" : "");
    return PU_ERR_UNSUPPORTED;
  }
  const int v[17] = {p.TH, p.TW, p.PW, p.tilesX, p.tilesY, p.nmb, p.cols, p.n3, p.a_bytes, p.w_bytes_max, p.w_res_bytes, p.tmem_cols,
                     p.nchunks, p.ncoblk, p.nstages, (int)p.smem_bytes, p.fold};
  for (int i = 0; i < 17; ++i) out17[i] = v[i];
  return PU_OK;
}
