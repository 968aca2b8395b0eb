// head.cu — the plastic output head and its Hebb / Oja trace (reference unet_p.py:70-88 ==
// unet_p_res.py:116-134), batched over B maps of N x N.
//
//   Weff = w + alpha*hebb ; A = X @ Weff ; S = sigmoid(A)                       (forward)
//   gA = gS*S*(1-S) ; gX = gA @ Weff^T ; gWeff = X^T @ gA (split-K) ;
//   gw = gWeff ; galpha = gWeff*hebb ; ghebb = gWeff*alpha                      (backward)
//   trace: out = decay*hebb + eta/K * pre^T @ post, decay = 1-eta (Hebb) or 1 - eta*q_j/K (Oja),
//          one fused pass (contraction + epilogue), exactly the reference at K == 1.
//
// The GEMMs of the general form are strict-fp32 shared-memory-tiled FFMA kernels (64x64 tiles, 4x4 per thread, K staged in whole
// 128-deep chunks): the head's logits decide the thresholded masks, so it stays in full fp32.  The training step of the TF32
// mode uses ONE fused kernel instead (head_bce_fused_kernel below: forward + BCE loss + backward, logits on mma.sync with an
// error-compensated TF32 split that keeps them at fp32 level) and a split-K tensor-core GEMM for the parameter gradients.
#include <stdlib.h>
#include "pu_common.cuh"

namespace pu {

enum { EPI_STORE = 0, EPI_SIGMOID = 1, EPI_ATOMIC = 2 };

// ---- strict-fp32 GEMM, whole-K-chunk staging --------------------------------------------------------------------------
// C[M,N] (row-major, ldc) (+)= sum_k A(m,k) B(k,n), A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].
// 64x64 tile per CTA, 4x4 outputs per thread, K consumed in chunks of 128: ALL loads of a chunk (8 x 128-bit per thread
// and operand) are in flight at once, then one barrier, then 128 k-steps from shared memory.  The head's K is 128
// (nbf), so a CTA pays ONE global round trip instead of the eight dependent 16-deep rounds of gemm_ffma_kernel
// (measured 18-22 us for a 0.27 GFLOP GEMM: latency-, not FMA-bound).  Operands that are contiguous along k are
// transposed on the way into shared memory (As[k][m], Bs[k][n]: conflict-free float4 reads in the k loop).
constexpr int G2_BM = 64, G2_BN = 64, G2_BK = 128;
constexpr int G2_SMEM = 2 * G2_BK * (G2_BM + 4) * (int)sizeof(float);

template <int ROWS>  // stage a ROWS(=64) x 128 tile into S[k][r] (row stride ROWS + 4)
__device__ __forceinline__ void g2_stage(float* __restrict__ S, const float* __restrict__ P, long long s_r, long long s_k, int r0, int k0,
                                         int r_end, int k_end, bool vec) {
  constexpr int LD = ROWS + 4;
  const int tid = threadIdx.x;
  if (s_k == 1) {
    // contiguous along k: thread -> (row r = i % 64, k quad kq = i / 64)
#pragma unroll
    for (int it = 0; it < (ROWS * G2_BK / 4) / 256; ++it) {
      const int i = tid + it * 256;
      const int r = i % ROWS, kq = i / ROWS;
      const int gr = r0 + r, gk = k0 + 4 * kq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < r_end) {
        const float* src = P + gr * s_r + gk;
        if (vec && gk + 3 < k_end) {
          v = ldg4(src);
        } else {
          if (gk < k_end) v.x = __ldg(src);
          if (gk + 1 < k_end) v.y = __ldg(src + 1);
          if (gk + 2 < k_end) v.z = __ldg(src + 2);
          if (gk + 3 < k_end) v.w = __ldg(src + 3);
        }
      }
      S[(4 * kq + 0) * LD + r] = v.x;
      S[(4 * kq + 1) * LD + r] = v.y;
      S[(4 * kq + 2) * LD + r] = v.z;
      S[(4 * kq + 3) * LD + r] = v.w;
    }
  } else {
    // contiguous along the row index (s_r == 1): thread -> (row quad rq = i % 16, k = i / 16)
#pragma unroll
    for (int it = 0; it < (ROWS * G2_BK / 4) / 256; ++it) {
      const int i = tid + it * 256;
      const int rq = i % (ROWS / 4), k = i / (ROWS / 4);
      const int gr = r0 + 4 * rq, gk = k0 + k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gk < k_end) {
        const float* src = P + gk * s_k + gr * s_r;
        if (vec && s_r == 1 && gr + 3 < r_end) {
          v = ldg4(src);
        } else {
          if (gr < r_end) v.x = __ldg(src);
          if (gr + 1 < r_end) v.y = __ldg(src + s_r);
          if (gr + 2 < r_end) v.z = __ldg(src + 2 * s_r);
          if (gr + 3 < r_end) v.w = __ldg(src + 3 * s_r);
        }
      }
      *reinterpret_cast<float4*>(S + k * LD + 4 * rq) = v;
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(256) gemm_chunk_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ Bm,
                                                         long long sbk, long long sbn, float* __restrict__ C, long long ldc, int M, int N,
                                                         int K, int kper, int vecA, int vecB) {
  extern __shared__ __align__(16) float g2_smem[];
  float* As = g2_smem;                          // [128][64 + 4]
  float* Bs = g2_smem + G2_BK * (G2_BM + 4);    // [128][64 + 4]
  constexpr int LD = G2_BM + 4;
  const int tid = threadIdx.x;
  const int m_blk = blockIdx.y * G2_BM, n_blk = blockIdx.x * G2_BN;
  const int k_begin = blockIdx.z * kper;
  const int k_end = min(K, k_begin + kper);
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += G2_BK) {
    if (k0 > k_begin) __syncthreads();
    g2_stage<G2_BM>(As, A, sam, sak, m_blk, k0, M, k_end, vecA != 0);
    g2_stage<G2_BN>(Bs, Bm, sbn, sbk, n_blk, k0, N, k_end, vecB != 0);
    __syncthreads();
    const int kn = min(G2_BK, k_end - k0);
    // software-pipelined k loop: the fragments of step k+1 are requested from shared memory before the 16 FMAs of step k
    // (a CTA has only two warps per scheduler; without this every step paid the full LDS latency)
    float4 a = *reinterpret_cast<const float4*>(As + ty * 4);
    float4 b = *reinterpret_cast<const float4*>(Bs + tx * 4);
#pragma unroll 8
    for (int k = 0; k < kn; ++k) {
      const int kp = k + 1 < G2_BK ? k + 1 : k;  // the row after the last valid one is still inside the tile
      const float4 an = *reinterpret_cast<const float4*>(As + kp * LD + ty * 4);
      const float4 bn = *reinterpret_cast<const float4*>(Bs + kp * LD + tx * 4);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      a = an;
      b = bn;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m_blk + ty * 4 + i;
    if (gm >= M) continue;
    const int gn = n_blk + tx * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = EPI == EPI_SIGMOID ? 1.f / (1.f + expf(-acc[i][j])) : acc[i][j];
    if (EPI == EPI_ATOMIC) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < N) atomicAdd(C + gm * ldc + gn + j, v[j]);
    } else if (gn + 3 < N && ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
      *reinterpret_cast<float4*>(C + gm * ldc + gn) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gn + j < N) C[gm * ldc + gn + j] = v[j];
    }
  }
}

template <int EPI>
static int launch_gemm_chunk(const float* A, long long sam, long long sak, const float* Bm, long long sbk, long long sbn, float* C, long long ldc,
                             int M, int N, int K, int kper, int splits, cudaStream_t st, const char* what) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_chunk_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM);
    // three 68 KB CTAs per SM: ask for the largest shared-memory carve-out (the default picks one that fits a single CTA,
    // which left the 256-CTA head GEMM at two waves of one latency-bound CTA per SM: 15 us for 0.27 GFLOP)
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_chunk_kernel<EPI>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
    attr_set = true;
  }
  // 128-bit loads need 16-byte aligned rows: base pointer and the non-unit stride
  const int vecA = aligned16(A) && ((sak == 1 ? sam : sak) % 4 == 0);
  const int vecB = aligned16(Bm) && ((sbk == 1 ? sbn : sbk) % 4 == 0);
  dim3 grid(cdiv(N, G2_BN), cdiv(M, G2_BM), splits);
  gemm_chunk_kernel<EPI><<<grid, 256, G2_SMEM, st>>>(A, sam, sak, Bm, sbk, sbn, C, ldc, M, N, K, kper, vecA, vecB);
  return post_launch(what);
}

__global__ void weff_kernel(const float* __restrict__ w, const float* __restrict__ alpha, const float* __restrict__ hebb,
                            float* __restrict__ weff, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) weff[i] = fmaf(alpha[i], hebb[i], w[i]);
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ S, const float* __restrict__ gS, float* __restrict__ gA, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = S[i];
    gA[i] = gS[i] * s * (1.f - s);
  }
}

__global__ void head_param_grads_kernel(const float* __restrict__ gw, const float* __restrict__ alpha, const float* __restrict__ hebb,
                                        float* __restrict__ galpha, float* __restrict__ ghebb, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = gw[i];
  if (galpha != nullptr) galpha[i] = g * hebb[i];
  if (ghebb != nullptr) ghebb[i] = g * alpha[i];
}

// ---- trace --------------------------------------------------------------------------------------
// mode 0: fused update (out = decay*hebb + eta/K*delta); mode 1: write delta_q only (DP split form)
__global__ void trace_contract_kernel(const float* __restrict__ hebb, const float* __restrict__ pre, const float* __restrict__ post,
                                      long long ld, int K, const float* __restrict__ eta_p, int rule, float* __restrict__ out,
                                      float* __restrict__ delta_q, int N, int Kdiv, int mode) {
  // block = (32, 8): the K pre/post rows of the block's 8 i-columns and 32 j-columns are staged through shared memory in
  // chunks of 64 (all loads of a chunk in flight at once: a per-k dependent global-load loop cost 19 us for K = 64)
  __shared__ float sp[64][8];
  __shared__ float sq[64][32];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int i0 = blockIdx.y * 8, j0 = blockIdx.x * 32;
  float d = 0.f, q = 0.f;
  for (int kc = 0; kc < K; kc += 64) {
    const int kn = min(64, K - kc);
    __syncthreads();
    for (int e = tid; e < kn * 40; e += 256) {
      const int k = e / 40, c = e - k * 40;
      if (c < 8) sp[k][c] = (i0 + c < N) ? __ldg(pre + (size_t)(kc + k) * ld + i0 + c) : 0.f;
      else sq[k][c - 8] = (j0 + c - 8 < N) ? __ldg(post + (size_t)(kc + k) * ld + j0 + c - 8) : 0.f;
    }
    __syncthreads();
    for (int k = 0; k < kn; ++k) {
      const float a = sp[k][threadIdx.y];
      const float b = sq[k][threadIdx.x];
      d = fmaf(a, b, d);
      q = fmaf(b, b, q);
    }
  }
  if (i >= N || j >= N) return;
  if (mode == 1) {
    delta_q[(size_t)i * N + j] = d;
    if (i == 0) delta_q[(size_t)N * N + j] = q;
    return;
  }
  const float eta = __ldg(eta_p);
  const float invK = 1.f / (float)Kdiv;
  const float h = hebb[(size_t)i * N + j];
  const float decay = rule == PU_RULE_HEBB ? 1.f - eta : 1.f - eta * q * invK;
  out[(size_t)i * N + j] = fmaf(decay, h, eta * d * invK);
}


// ---- trace contraction on tensor cores (rows = 'all': K = B*N pre/post pairs) ------------------------------------------
// delta[i][j] += sum_k pre[k][i] * post[k][j],  q[j] += sum_k post[k][j]^2   over this CTA's K slice.
// The literal "every pixel-row" reading of the plastic update (the reference's bmm computes all N outer products before
// keeping [0], unet_p.py:82): a [N x K] x [K x N] GEMM with K = B*N (8192 at B = 64, N = 128; 134 M MACs).  Warp-level
// mma.sync.m16n8k8 TF32 with the 3xTF32 split (a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, a_lo = a - tf32(a)): the trace
// keeps fp32-level accuracy (~1e-6 relative) while the MACs run on the tensor cores.  CTA = 64 x 64 outputs x one K
// slice, 4 warps x (32 x 32), operands staged through shared memory in 32-row chunks, partial sums merged with fp32 atomics.
__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int TD_T = 64, TD_KC = 32, TD_LD = TD_T + 8;  // row pitch 72: fragment reads (k = t, col = g) hit 32 distinct banks

template <bool WITH_Q, int TERMS = 3>
__global__ void __launch_bounds__(128) trace_delta_mma_kernel(const float* __restrict__ pre, const float* __restrict__ post, long long ld, int K,
                                                              int kper, float* __restrict__ delta_q, int N) {
  __shared__ __align__(16) float td_sm[2 * TD_KC * TD_LD];  // the two operand tiles; afterwards the 64 x (64 + 4) output tile
  float (*Ps)[TD_LD] = reinterpret_cast<float (*)[TD_LD]>(td_sm);
  float (*Qs)[TD_LD] = reinterpret_cast<float (*)[TD_LD]>(td_sm + TD_KC * TD_LD);
  const int i0 = blockIdx.y * TD_T, j0 = blockIdx.x * TD_T;
  const int k_begin = blockIdx.z * kper, k_end = min(K, k_begin + kper);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wi = (warp >> 1) * 32, wj = (warp & 1) * 32;  // this warp's 32 x 32 sub-tile
  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;
  float qsum = 0.f;
  // 128-bit path (aligned rows): the next 32-row chunk is fetched into registers (8 x LDG.128 per thread) while the current one
  // is multiplied — the scalar, unpipelined staging paid a full HBM round trip per chunk
  const bool vec = (N & 3) == 0 && (ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(post)) & 15u) == 0;
  float4 pa[4], qa[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + u * 128;
      const int k = e >> 4, c = (e & 15) * 4;
      const bool kin = k0 + k < k_end;
      pa[u] = (kin && i0 + c < N) ? ldg4(pre + (size_t)(k0 + k) * ld + i0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      qa[u] = (kin && j0 + c < N) ? ldg4(post + (size_t)(k0 + k) * ld + j0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (vec) fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += TD_KC) {
    __syncthreads();
    // stage 32 rows x 64 columns of pre (columns i0..) and post (columns j0..): coalesced along the row
    if (vec) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = tid + u * 128;
        const int k = e >> 4, c = (e & 15) * 4;
        *reinterpret_cast<float4*>(&Ps[k][c]) = pa[u];
        *reinterpret_cast<float4*>(&Qs[k][c]) = qa[u];
      }
      if (k0 + TD_KC < k_end) fetch(k0 + TD_KC);
    } else {
      for (int e = tid; e < TD_KC * TD_T; e += 128) {
        const int k = e / TD_T, c = e - k * TD_T;
        const bool kin = k0 + k < k_end;
        Ps[k][c] = (kin && i0 + c < N) ? __ldg(pre + (size_t)(k0 + k) * ld + i0 + c) : 0.f;
        Qs[k][c] = (kin && j0 + c < N) ? __ldg(post + (size_t)(k0 + k) * ld + j0 + c) : 0.f;
      }
    }
    __syncthreads();
    if (WITH_Q && blockIdx.y == 0 && tid < TD_T) {  // q_j = sum_k post_kj^2, once per column tile
#pragma unroll 8
      for (int k = 0; k < TD_KC; ++k) qsum = fmaf(Qs[k][tid], Qs[k][tid], qsum);
    }
#pragma unroll
    for (int kk = 0; kk < TD_KC; kk += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {  // A = pre^T: A[m = i][k] = Ps[k][i]
        const int m = wi + a * 16 + g;
        const float v[4] = {Ps[kk + t][m], Ps[kk + t][m + 8], Ps[kk + t + 4][m], Ps[kk + t + 4][m + 8]};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ah[a][e] = f2tf32(v[e]);
          al[a][e] = TERMS == 3 ? f2tf32(v[e] - __uint_as_float(ah[a][e])) : 0u;
        }
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {  // B[k][n = j] = Qs[k][j]
        const int n = wj + b * 8 + g;
        const float v[2] = {Qs[kk + t][n], Qs[kk + t + 4][n]};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          bh[b][e] = f2tf32(v[e]);
          bl[b][e] = TERMS == 3 ? f2tf32(v[e] - __uint_as_float(bh[b][e])) : 0u;
        }
      }
      // term-outer order: consecutive MMAs write different accumulators (no stall on the previous MMA's result)
#pragma unroll
      for (int term = (TERMS == 3 ? 0 : 2); term < 3; ++term)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            mma_tf32_16x8x8(acc[a][b], term == 0 ? al[a] : ah[a], term == 1 ? bl[b] : bh[b]);  // small terms first
    }
  }
  if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(delta_q) & 15u) == 0) {
    // merge through shared memory: one 16-byte vector reduction per four outputs (1024 per CTA instead of 4096 scalar atomics —
    // the scalar version was bound by the L2 atomic rate: 1.2 M atomics on 16 K addresses)
    constexpr int OLD = TD_T + 4;
    static_assert(TD_T * OLD <= 2 * TD_KC * TD_LD, "output tile must fit the operand tiles");
    __syncthreads();  // every warp is done with the operand tiles
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          *reinterpret_cast<float2*>(td_sm + (wi + a * 16 + g + h * 8) * OLD + wj + b * 8 + 2 * t) =
              make_float2(acc[a][b][2 * h], acc[a][b][2 * h + 1]);
    __syncthreads();
    for (int e = tid; e < TD_T * (TD_T / 4); e += 128) {
      const int r = e >> 4, q = e & 15;
      const int i = i0 + r, j = j0 + 4 * q;
      if (i < N && j < N) {
        const float4 v = *reinterpret_cast<const float4*>(td_sm + r * OLD + 4 * q);
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(delta_q + (size_t)i * N + j), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = i0 + wi + a * 16 + g + (e >> 1) * 8, j = j0 + wj + b * 8 + 2 * t + (e & 1);
          if (i < N && j < N) atomicAdd(delta_q + (size_t)i * N + j, acc[a][b][e]);
        }
  }
  if (WITH_Q && blockIdx.y == 0 && tid < TD_T && j0 + tid < N) atomicAdd(delta_q + (size_t)N * N + j0 + tid, qsum);
}

// ---- fused head for the training step: forward + BCE loss + sigmoid/BCE backward + gX in ONE kernel ------------------------
// (TF32 mode of TrainStep; the strict-fp32 path keeps the FFMA GEMMs above.)  A CTA owns 64 (or 32) rows of X [M = B*N, N]:
//   phase 1  Z = X_tile @ Weff, Weff = w + alpha*hebb built on the way into shared memory, or copied in if it was computed beforehand
//   epilogue S = sigmoid(Z) -> global; loss -= t*log(S) + (1-t)*log(1-S) (torch.nn.BCELoss: logs clamped at -100, mean);
//            gA = dLoss/dZ = ((S-t)/max(S(1-S),1e-12)/n) * S * (1-S) -> global (the parameter gradients read it) and back into
//            the X tile's shared memory
//   phase 2  gX = gA_tile @ Weff^T from the SAME shared-memory copy of Weff
// mma.sync.m16n8k8 TF32.  Phase 1 uses the three-term error-compensated split (x = hi + lo, hi = the top 19 bits:
// x*y ~ hi*hi' + hi*lo' + lo*hi'; the dropped terms are 2^-20 relative): the logits decide the thresholded masks, the loss and
// the trace, and keep fp32-level accuracy.  Phase 2 is a plain TF32 product of round-to-nearest operands: gX enters the
// tensor-core data-gradient convs, which round it to TF32 anyway.
// Weff sits in shared memory ONCE, unpadded, with the column index XOR-swizzled by s(r) = 8*(r&3) + 4*((r>>2)&1): the B
// fragments of phase 1 (rows k+t, columns n+g) and of phase 2 (rows n+g, columns k+t) are then both bank-conflict free.
// The loss is reduced deterministically and without a memset: every CTA publishes its partial sum, the last one to arrive
// (ticket counter in `scratch`, re-zeroed for the next launch) adds them in block order.
constexpr int HF_NP = 128, HF_LDX = HF_NP + 4, HF_THREADS = 512;
// rows of X per CTA: 64 (one CTA per SM, 16 warps as 4 x 4 tiles of 16 x 32) or 32 (two co-resident CTAs per SM, 2 x 8 tiles of
// 16 x 16: one CTA's staging and pointwise phases overlap the other's MMAs)
constexpr int hf_smem(int bm) { return (HF_NP * HF_NP + bm * HF_LDX + bm * HF_NP) * (int)sizeof(float); }

__device__ __forceinline__ int hf_swz(int r) { return ((r & 3) << 3) | (r & 4); }
__device__ __forceinline__ void hf_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;                                  // top 19 bits: a valid TF32 operand
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;            // exact residual, truncated to TF32
}
__device__ __forceinline__ uint32_t hf_rn(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }  // RN (ties away) to TF32

// acc = As[row tile][k] @ B, B[k][n] = TRANS_B ? Ws[n][k] : Ws[k][n]; this warp's 16 x (8 NT) sub-tile.
// TERMS = 3: error-compensated split; TERMS = 1: plain TF32.
template <bool TRANS_B, int TERMS, int NT>
__device__ __forceinline__ void hf_gemm(float (&acc)[NT][4], const float* __restrict__ As, const float* __restrict__ Ws, int wm, int wn, int g,
                                        int t, int nk, int N) {
#pragma unroll
  for (int b = 0; b < NT; ++b)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[b][e] = 0.f;
#pragma unroll 2
  for (int ks = 0; ks < nk; ++ks) {
    const int kk = ks * 8;
    uint32_t ah[4], al[4], bh[NT][2], bl[NT][2];
    {
      const float* r0 = As + (wm + g) * HF_LDX + kk + t;
      const float v[4] = {r0[0], r0[8 * HF_LDX], r0[4], r0[8 * HF_LDX + 4]};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (TERMS == 3) {
          hf_split(v[e], ah[e], al[e]);
        } else {
          ah[e] = hf_rn(v[e]);
          al[e] = 0u;
        }
      }
    }
#pragma unroll
    for (int b = 0; b < NT; ++b) {
      const int n0 = wn + b * 8;
      float v[2] = {0.f, 0.f};
      if (n0 < N) {  // warp-uniform
        if (!TRANS_B) {  // B[k][n] = Ws[k][n]
          const int r = kk + t, c = n0 + g;
          v[0] = Ws[r * HF_NP + (c ^ hf_swz(r))];
          v[1] = Ws[(r + 4) * HF_NP + (c ^ hf_swz(r + 4))];
        } else {  // B[k][n] = Ws[n][k]
          const int r = n0 + g, c = kk + t, sw = hf_swz(r);
          v[0] = Ws[r * HF_NP + (c ^ sw)];
          v[1] = Ws[r * HF_NP + ((c + 4) ^ sw)];
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (TERMS == 3) {
          hf_split(v[e], bh[b][e], bl[b][e]);
        } else {
          bh[b][e] = hf_rn(v[e]);
          bl[b][e] = 0u;
        }
      }
    }
    // term-outer order: consecutive MMAs write DIFFERENT accumulators (an mma.sync waits for the previous one into the same
    // accumulator: with the three terms of a tile back to back the warp stalled two MMA latencies per tile)
#pragma unroll
    for (int term = (TERMS == 3 ? 0 : 2); term < 3; ++term)
#pragma unroll
      for (int b = 0; b < NT; ++b) {
        if (wn + b * 8 >= N) continue;
        mma_tf32_16x8x8(acc[b], term == 0 ? al : ah, term == 1 ? bl[b] : bh[b]);  // small terms first
      }
  }
}

// one BCE element: s = sigmoid(z) (exact division: s decides the masks); loss term as torch.nn.BCELoss evaluates it on the fp32 s
// (logs clamped at -100; log(1 - s) for log1p(-s): within 6e-8 absolute; hardware log2 — the loss scalar only, mean of B*N*N terms);
// ga = dloss/dz = ((s - t) / max(s (1 - s), 1e-12) / n) * s (1 - s), i.e. (s - t) / n unless the clamp of the reference is active
__device__ __forceinline__ void hf_bce(float z, float tt, float inv_n, float& s, float& ga, float& lsum) {
  s = 1.f / (1.f + expf(-z));
  const float l1 = fmaxf(__logf(s), -100.f), l0 = fmaxf(__logf(1.f - s), -100.f);
  lsum -= tt * l1 + (1.f - tt) * l0;
  const float p = s * (1.f - s);
  ga = p >= 1e-12f ? (s - tt) * inv_n : (s - tt) / 1e-12f * inv_n * p;
}

__device__ __forceinline__ void hf_cp16(float* dst_smem, const float* src, bool ok) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src), "r"(ok ? 16 : 0)
               : "memory");
}

// The kernel is executed ONCE per warp (128 CTAs, no tile loop), so its code size is its cost: the first version — pointwise
// epilogue unrolled over the 32 accumulator registers of a thread, ~10 k instructions — spent 46 % of its issue slots waiting
// for instruction fetches (ncu: stall_no_instruction).  Hence the pointwise work is a ROLLED loop over the tile in shared
// memory (logits out of the accumulators, gA back in for phase 2), with coalesced 128-bit global accesses; 16 warps per CTA
// (8 in the first versions) for memory-level parallelism in the staging and pointwise phases.  weff != NULL: Weff was computed beforehand
// (pu_head_weff, off the critical path) and arrives by asynchronous copies — building it in the kernel from w, alpha and hebb
// triples the L2 traffic of the staging phase (30 % of the kernel in the ncu capture of that version).
template <int HF_BM>
__global__ void __launch_bounds__(HF_THREADS, HF_BM == 32 ? 2 : 1) head_bce_fused_kernel(const float* __restrict__ X, const float* __restrict__ w,
                                                                    const float* __restrict__ alpha, const float* __restrict__ hebb,
                                                                    const float* __restrict__ weff, const float* __restrict__ T,
                                                                    float* __restrict__ S, float* __restrict__ loss, float* __restrict__ gA,
                                                                    float* __restrict__ gX, float* __restrict__ scratch, int M, int N,
                                                                    float inv_n, int vec) {
  extern __shared__ __align__(16) float hf_smem[];
  __shared__ float red[HF_THREADS / 32];
  float* Ws = hf_smem;                   // [128][128], swizzled columns
  float* Xs = hf_smem + HF_NP * HF_NP;   // [HF_BM][132]: the X tile, then the logits, then the gA tile
  float* Ts = Xs + HF_BM * HF_LDX;       // [HF_BM][128]: the targets of the tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int m_blk = blockIdx.x * HF_BM;
  constexpr int WM = HF_BM / 16, WN = (HF_THREADS / 32) / WM, NT = HF_NP / (8 * WN);  // warps along m / n, n-tiles per warp
  const int wm = (warp / WN) * 16, wn = (warp % WN) * (8 * NT);
  const int nq = N >> 2;
  // ---- stage: X tile and targets (HBM; asynchronous copies), then Weff (L2 after the first CTA); zero outside [M x N] / [N x N]
  if (vec) {
#pragma unroll
    for (int it = 0; it < HF_BM * (HF_NP / 4) / HF_THREADS; ++it) {
      const int e = tid + it * HF_THREADS;
      const int r = e >> 5, q = e & 31;
      const bool ok = m_blk + r < M && q < nq;
      const size_t o = ok ? (size_t)(m_blk + r) * N + 4 * q : 0;
      hf_cp16(Xs + r * HF_LDX + 4 * q, X + o, ok);
      hf_cp16(Ts + r * HF_NP + 4 * q, T + o, ok);
    }
    if (weff != nullptr) {
#pragma unroll
      for (int it = 0; it < HF_NP * (HF_NP / 4) / HF_THREADS; ++it) {
        const int e = tid + it * HF_THREADS;
        const int r = e >> 5, q = e & 31;
        const bool ok = r < N && q < nq;
        hf_cp16(Ws + r * HF_NP + ((4 * q) ^ hf_swz(r)), weff + (ok ? (size_t)r * N + 4 * q : 0), ok);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
      asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 8
      for (int it = 0; it < HF_NP * (HF_NP / 4) / HF_THREADS; ++it) {
        const int e = tid + it * HF_THREADS;
        const int r = e >> 5, q = e & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < N && q < nq) {
          const size_t o = (size_t)r * N + 4 * q;
          const float4 a = ldg4(alpha + o), h = ldg4(hebb + o), b = ldg4(w + o);
          v = make_float4(fmaf(a.x, h.x, b.x), fmaf(a.y, h.y, b.y), fmaf(a.z, h.z, b.z), fmaf(a.w, h.w, b.w));
        }
        *reinterpret_cast<float4*>(Ws + r * HF_NP + ((4 * q) ^ hf_swz(r))) = v;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
#pragma unroll 4
    for (int e = tid; e < HF_BM * HF_NP; e += HF_THREADS) {
      const int r = e >> 7, c = e & 127;
      const bool ok = m_blk + r < M && c < N;
      const size_t o = ok ? (size_t)(m_blk + r) * N + c : 0;
      Xs[r * HF_LDX + c] = ok ? __ldg(X + o) : 0.f;
      Ts[r * HF_NP + c] = ok ? __ldg(T + o) : 0.f;
    }
#pragma unroll 4
    for (int e = tid; e < HF_NP * HF_NP; e += HF_THREADS) {
      const int r = e >> 7, c = e & 127;
      float v = 0.f;
      if (r < N && c < N) {
        const size_t o = (size_t)r * N + c;
        v = weff != nullptr ? __ldg(weff + o) : fmaf(__ldg(alpha + o), __ldg(hebb + o), __ldg(w + o));
      }
      Ws[r * HF_NP + (c ^ hf_swz(r))] = v;
    }
  }
  __syncthreads();
  const int nk = (N + 7) >> 3;
  float acc[NT][4];
  hf_gemm<false, 3, NT>(acc, Xs, Ws, wm, wn, g, t, nk, N);
  __syncthreads();  // every warp is done reading the X tile
  // logits -> shared memory (columns of skipped n-tiles are >= N and never read)
#pragma unroll
  for (int b = 0; b < NT; ++b) {
    if (wn + b * 8 >= N) continue;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow)
      *reinterpret_cast<float2*>(Xs + (wm + g + hrow * 8) * HF_LDX + wn + b * 8 + 2 * t) = make_float2(acc[b][hrow * 2], acc[b][hrow * 2 + 1]);
  }
  __syncthreads();
  // ---- pointwise pass (rolled): sigmoid, BCE, gradient of the logits; the tile in shared memory becomes gA
  float lsum = 0.f;
  if (vec) {
#pragma unroll 1
    for (int it = 0; it < HF_BM * (HF_NP / 4) / HF_THREADS; ++it) {
      const int e = tid + it * HF_THREADS;
      const int r = e >> 5, q = e & 31;
      float4 ga4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m_blk + r < M && q < nq) {
        const float4 z4 = *reinterpret_cast<const float4*>(Xs + r * HF_LDX + 4 * q);
        const float4 t4 = *reinterpret_cast<const float4*>(Ts + r * HF_NP + 4 * q);
        float4 s4;
        hf_bce(z4.x, t4.x, inv_n, s4.x, ga4.x, lsum);
        hf_bce(z4.y, t4.y, inv_n, s4.y, ga4.y, lsum);
        hf_bce(z4.z, t4.z, inv_n, s4.z, ga4.z, lsum);
        hf_bce(z4.w, t4.w, inv_n, s4.w, ga4.w, lsum);
        const size_t o = (size_t)(m_blk + r) * N + 4 * q;
        *reinterpret_cast<float4*>(S + o) = s4;
        *reinterpret_cast<float4*>(gA + o) = ga4;
      }
      *reinterpret_cast<float4*>(Xs + r * HF_LDX + 4 * q) = ga4;
    }
  } else {
#pragma unroll 1
    for (int e = tid; e < HF_BM * HF_NP; e += HF_THREADS) {
      const int r = e >> 7, c = e & 127;
      float ga = 0.f;
      if (m_blk + r < M && c < N) {
        float s;
        hf_bce(Xs[r * HF_LDX + c], Ts[r * HF_NP + c], inv_n, s, ga, lsum);
        const size_t o = (size_t)(m_blk + r) * N + c;
        S[o] = s;
        gA[o] = ga;
      }
      Xs[r * HF_LDX + c] = ga;
    }
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();  // the gA tile is complete
  if (tid == 0) {
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < HF_THREADS / 32; ++u) s += red[u];
    if (scratch == nullptr) {
      atomicAdd(loss, s * inv_n);  // the caller zeroed *loss
    } else {
      // scratch[0] = ticket counter (zero between launches), scratch[1 + b] = partial sum of CTA b; the ticket is drawn at the
      // very end of the kernel (nothing of this CTA waits for the atomic's round trip)
      scratch[1 + blockIdx.x] = s;
      __threadfence();
    }
  }
  if (gX != nullptr) {
    // ---- phase 2: gX = gA @ Weff^T
    hf_gemm<true, 1, NT>(acc, Xs, Ws, wm, wn, g, t, nk, N);
    const bool pair = (N & 1) == 0;  // (row*N + col) is even for even col: 8-byte stores
#pragma unroll
    for (int b = 0; b < NT; ++b)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int row = m_blk + wm + g + hrow * 8;
        const int col = wn + b * 8 + 2 * t;
        if (row >= M || col >= N) continue;
        const size_t o = (size_t)row * N + col;
        if (pair) {
          *reinterpret_cast<float2*>(gX + o) = make_float2(acc[b][hrow * 2], acc[b][hrow * 2 + 1]);
        } else {
          gX[o] = acc[b][hrow * 2];
          if (col + 1 < N) gX[o + 1] = acc[b][hrow * 2 + 1];
        }
      }
  }
  // ---- loss: the last CTA to arrive adds the partial sums in a fixed order (warp 0: lane-strided, then the shuffle tree)
  if (scratch != nullptr && warp == 0) {
    unsigned ticket = 0;
    if (lane == 0) ticket = atomicAdd(reinterpret_cast<unsigned int*>(scratch), 1u);  // after this thread's own fenced partial store
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      const volatile float* part = scratch + 1;
      float tot = 0.f;
      for (unsigned b = lane; b < gridDim.x; b += 32) tot += part[b];
      tot = warp_sum(tot);
      if (lane == 0) {
        *loss = tot * inv_n;
        *reinterpret_cast<unsigned int*>(scratch) = 0u;
      }
    }
  }
}

__global__ void trace_apply_kernel(const float* __restrict__ hebb, const float* __restrict__ delta_q, int Kdiv,
                                   const float* __restrict__ eta_p, int rule, float* __restrict__ out, int N) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= N || j >= N) return;
  const float eta = __ldg(eta_p);
  const float invK = 1.f / (float)Kdiv;
  const float d = delta_q[(size_t)i * N + j];
  const float q = delta_q[(size_t)N * N + j];
  const float h = hebb[(size_t)i * N + j];
  const float decay = rule == PU_RULE_HEBB ? 1.f - eta : 1.f - eta * q * invK;
  out[(size_t)i * N + j] = fmaf(decay, h, eta * d * invK);
}

// backward, part 1: ghebb and geta
__global__ void trace_bwd_hebb_eta_kernel(const float* __restrict__ hebb, const float* __restrict__ pre, const float* __restrict__ post,
                                          long long ld, int K, const float* __restrict__ eta_p, int rule, const float* __restrict__ D,
                                          float* __restrict__ ghebb, float* __restrict__ geta, int N) {
  __shared__ float red[8];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  float contrib = 0.f;
  if (i < N && j < N) {
    float d = 0.f, q = 0.f;
    for (int k = 0; k < K; ++k) {
      const float a = __ldg(pre + k * ld + i);
      const float b = __ldg(post + k * ld + j);
      d = fmaf(a, b, d);
      q = fmaf(b, b, q);
    }
    const float eta = __ldg(eta_p);
    const float invK = 1.f / (float)K;
    const float h = hebb[(size_t)i * N + j];
    const float g = D[(size_t)i * N + j];
    const float decay = rule == PU_RULE_HEBB ? 1.f - eta : 1.f - eta * q * invK;
    if (ghebb != nullptr) ghebb[(size_t)i * N + j] = g * decay;
    contrib = rule == PU_RULE_HEBB ? g * (d * invK - h) : g * (d * invK - h * q * invK);
  }
  if (geta == nullptr) return;
  contrib = warp_sum(contrib);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if ((tid & 31) == 0) red[tid >> 5] = contrib;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int u = 0; u < (int)(blockDim.x * blockDim.y) / 32; ++u) s += red[u];
    atomicAdd(geta, s);
  }
}

// gpre[k][i] = eta/K * sum_j D[i][j] post[k][j]
__global__ void trace_bwd_pre_kernel(const float* __restrict__ post, long long ld, int K, const float* __restrict__ eta_p,
                                     const float* __restrict__ D, float* __restrict__ gpre, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (i >= N) return;
  float s = 0.f;
  for (int j = 0; j < N; ++j) s = fmaf(__ldg(D + (size_t)i * N + j), __ldg(post + k * ld + j), s);
  gpre[(size_t)k * N + i] = __ldg(eta_p) / (float)K * s;
}

// gpost[k][j] = eta/K * sum_i D[i][j] * (pre[k][i] - (oja ? 2*hebb[i][j]*post[k][j] : 0))
__global__ void trace_bwd_post_kernel(const float* __restrict__ hebb, const float* __restrict__ pre, const float* __restrict__ post,
                                      long long ld, int K, const float* __restrict__ eta_p, int rule, const float* __restrict__ D,
                                      float* __restrict__ gpost, int N) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (j >= N) return;
  const float pj = __ldg(post + k * ld + j);
  float s = 0.f;
  for (int i = 0; i < N; ++i) {
    const float g = __ldg(D + (size_t)i * N + j);
    float t = __ldg(pre + k * ld + i);
    if (rule == PU_RULE_OJA) t -= 2.f * __ldg(hebb + (size_t)i * N + j) * pj;
    s = fmaf(g, t, s);
  }
  gpost[(size_t)k * N + j] = __ldg(eta_p) / (float)K * s;
}

}  // namespace pu

extern "C" {

int pu_plastic_head_fwd(const float* X, const float* w, const float* alpha, const float* hebb, float* weff_out, float* S, int B,
                        int N, void* stream) {
  PU_REQUIRE(X && w && alpha && hebb && weff_out && S && B > 0 && N > 0, PU_ERR_BAD_ARG, "pu_plastic_head_fwd: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  pu::weff_kernel<<<pu::cdiv((long long)N * N, 256), 256, 0, st>>>(w, alpha, hebb, weff_out, N * N);
  int rc = pu::post_launch("pu_plastic_head_fwd weff");
  if (rc) return rc;
  const int M = B * N;
  PU_REQUIRE(pu::cdiv(M, 64) <= 65535, PU_ERR_UNSUPPORTED, "pu_plastic_head_fwd: B*N too large");
  return pu::launch_gemm_chunk<pu::EPI_SIGMOID>(X, N, 1, weff_out, N, 1, S, N, M, N, N, N, 1, st, "pu_plastic_head_fwd gemm");
}

int pu_plastic_head_bwd(const float* X, const float* S, const float* gS, const float* weff, const float* alpha, const float* hebb,
                        float* gA_ws, float* gX, float* gw, float* galpha, float* ghebb, int B, int N, void* stream) {
  // gS == NULL: gA_ws already holds gA (a previous call computed it); gw == NULL: skip the parameter gradients.  The two
  // halves can then be issued separately (gX on the critical path, the parameter gradients on a side stream).
  PU_REQUIRE(X && S && weff && alpha && hebb && gA_ws && (gS || gw) && B > 0 && N > 0, PU_ERR_BAD_ARG, "pu_plastic_head_bwd: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  const int M = B * N;
  const long long n = (long long)M * N;
  int rc = PU_OK;
  if (gS != nullptr) {
    int g = (int)((n + 1023) / 1024);
    g = g < 1 ? 1 : (g > 16 * pu::kNumSMs ? 16 * pu::kNumSMs : g);
    pu::sigmoid_bwd_kernel<<<g, 256, 0, st>>>(S, gS, gA_ws, n);
    rc = pu::post_launch("pu_plastic_head_bwd gA");
    if (rc) return rc;
  }
  if (gX != nullptr) {
    // gX[m][n] = sum_k gA[m][k] * weff[n][k]
    rc = pu::launch_gemm_chunk<pu::EPI_STORE>(gA_ws, N, 1, weff, 1, N, gX, N, M, N, N, N, 1, st, "pu_plastic_head_bwd gX");
    if (rc) return rc;
  }
  if (gw == nullptr) return PU_OK;
  cudaError_t e = cudaMemsetAsync(gw, 0, sizeof(float) * N * N, st);
  if (e != cudaSuccess) {
    pu::set_error("pu_plastic_head_bwd memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  {
    // gW[i][j] = sum_m X[m][i] * gA[m][j], split over m
    const int tiles = pu::cdiv(N, 64) * pu::cdiv(N, 64);
    int splits = (pu::kNumSMs + tiles - 1) / tiles;
    int kper = (M + splits - 1) / splits;
    kper = ((kper + 127) / 128) * 128;  // whole 128-deep chunks per split
    splits = (M + kper - 1) / kper;
    rc = pu::launch_gemm_chunk<pu::EPI_ATOMIC>(X, 1, N, gA_ws, N, 1, gw, N, N, N, M, kper, splits, st, "pu_plastic_head_bwd gW");
    if (rc) return rc;
  }
  if (galpha != nullptr || ghebb != nullptr) {
    pu::head_param_grads_kernel<<<pu::cdiv((long long)N * N, 256), 256, 0, st>>>(gw, alpha, hebb, galpha, ghebb, N * N);
    rc = pu::post_launch("pu_plastic_head_bwd param grads");
    if (rc) return rc;
  }
  return PU_OK;
}

int pu_head_weff(const float* w, const float* alpha, const float* hebb, float* weff, int N, void* stream) {
  PU_REQUIRE(w && alpha && hebb && weff && N > 0, PU_ERR_BAD_ARG, "pu_head_weff: bad argument");
  pu::weff_kernel<<<pu::cdiv((long long)N * N, 256), 256, 0, pu::as_stream(stream)>>>(w, alpha, hebb, weff, N * N);
  return pu::post_launch("pu_head_weff");
}

int pu_plastic_head_bce(const float* X, const float* w, const float* alpha, const float* hebb, const float* weff, const float* target,
                        float* S, float* loss, float* gA, float* gX, float* scratch, int B, int N, void* stream) {
  PU_REQUIRE(X && (weff || (w && alpha && hebb)) && target && S && loss && gA && B > 0 && N > 0, PU_ERR_BAD_ARG,
             "pu_plastic_head_bce: bad argument");
  PU_REQUIRE(N <= pu::HF_NP, PU_ERR_UNSUPPORTED, "pu_plastic_head_bce: N=%d > %d", N, pu::HF_NP);
  cudaStream_t st = pu::as_stream(stream);
  static int bm = 0;  // rows per CTA (PU_HEAD_BM=32|64)
  if (bm == 0) {
    const char* env = getenv("PU_HEAD_BM");
    const int want = (env != nullptr && atoi(env) == 32) ? 32 : 64;
    cudaError_t e = cudaFuncSetAttribute(pu::head_bce_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, pu::hf_smem(64));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pu::head_bce_fused_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, pu::hf_smem(32));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pu::head_bce_fused_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) {
      pu::set_error("pu_plastic_head_bce: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
    bm = want;
  }
  if (scratch == nullptr) {  // no ticket scratch: atomics into a zeroed *loss
    cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
    if (e != cudaSuccess) {
      pu::set_error("pu_plastic_head_bce memset: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
  }
  const int M = B * N;
  const int vec = (N % 4 == 0) && pu::aligned16(X) && pu::aligned16(target) && pu::aligned16(S) && pu::aligned16(gA) &&
                  (weff != nullptr ? pu::aligned16(weff) : (pu::aligned16(w) && pu::aligned16(alpha) && pu::aligned16(hebb)));
  const bool al8 = ((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(gA) |
                     reinterpret_cast<uintptr_t>(gX)) & 7u) == 0;
  PU_REQUIRE(al8, PU_ERR_BAD_ARG, "pu_plastic_head_bce: target, S, gA, gX must be 8-byte aligned");
  const float inv_n = 1.f / ((float)M * (float)N);
  if (bm == 32)
    pu::head_bce_fused_kernel<32><<<pu::cdiv(M, 32), pu::HF_THREADS, pu::hf_smem(32), st>>>(X, w, alpha, hebb, weff, target, S, loss, gA, gX,
                                                                                           scratch, M, N, inv_n, vec);
  else
    pu::head_bce_fused_kernel<64><<<pu::cdiv(M, 64), pu::HF_THREADS, pu::hf_smem(64), st>>>(X, w, alpha, hebb, weff, target, S, loss, gA, gX,
                                                                                           scratch, M, N, inv_n, vec);
  return pu::post_launch("pu_plastic_head_bce");
}

int pu_plastic_head_wgrad_tc(const float* X, const float* gA, const float* alpha, const float* hebb, float* gw, float* galpha,
                             float* ghebb, int B, int N, int terms, void* stream) {
  PU_REQUIRE(X && gA && alpha && hebb && gw && B > 0 && N > 0, PU_ERR_BAD_ARG, "pu_plastic_head_wgrad_tc: bad argument");
  PU_REQUIRE(terms == 1 || terms == 3, PU_ERR_BAD_ARG, "pu_plastic_head_wgrad_tc: terms must be 1 (TF32) or 3 (3xTF32)");
  cudaStream_t st = pu::as_stream(stream);
  cudaError_t e = cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)N * N, st);
  if (e != cudaSuccess) {
    pu::set_error("pu_plastic_head_wgrad_tc memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  // gW[i][j] = sum_m X[m][i] * gA[m][j]: the trace-delta contraction with pre = X, post = gA, K = B*N rows
  const int K = B * N;
  const int tiles = pu::cdiv(N, pu::TD_T);
  int splits = pu::cdiv(2 * pu::kNumSMs, tiles * tiles);
  int kper = pu::cdiv(K, splits);
  kper = pu::cdiv(kper, pu::TD_KC) * pu::TD_KC;
  splits = pu::cdiv(K, kper);
  PU_REQUIRE(splits <= 65535, PU_ERR_UNSUPPORTED, "pu_plastic_head_wgrad_tc: B*N too large");
  if (terms == 3)
    pu::trace_delta_mma_kernel<false, 3><<<dim3(tiles, tiles, splits), 128, 0, st>>>(X, gA, N, K, kper, gw, N);
  else
    pu::trace_delta_mma_kernel<false, 1><<<dim3(tiles, tiles, splits), 128, 0, st>>>(X, gA, N, K, kper, gw, N);
  int rc = pu::post_launch("pu_plastic_head_wgrad_tc");
  if (rc) return rc;
  if (galpha != nullptr || ghebb != nullptr) {
    pu::head_param_grads_kernel<<<pu::cdiv((long long)N * N, 256), 256, 0, st>>>(gw, alpha, hebb, galpha, ghebb, N * N);
    rc = pu::post_launch("pu_plastic_head_wgrad_tc param grads");
  }
  return rc;
}

int pu_trace_update_fwd(const float* hebb, const float* pre, const float* post, long long ld, int K, const float* eta, int rule,
                        float* out, int N, void* stream) {
  PU_REQUIRE(hebb && pre && post && eta && out && K > 0 && N > 0 && ld >= N, PU_ERR_BAD_ARG, "pu_trace_update_fwd: bad argument");
  PU_REQUIRE(rule == PU_RULE_HEBB || rule == PU_RULE_OJA, PU_ERR_BAD_ARG, "pu_trace_update_fwd: unknown rule %d", rule);
  dim3 block(32, 8), grid(pu::cdiv(N, 32), pu::cdiv(N, 8));
  pu::trace_contract_kernel<<<grid, block, 0, pu::as_stream(stream)>>>(hebb, pre, post, ld, K, eta, rule, out, nullptr, N, K, 0);
  return pu::post_launch("pu_trace_update_fwd");
}

int pu_trace_delta(const float* pre, const float* post, long long ld, int K, float* delta_q, int N, void* stream) {
  PU_REQUIRE(pre && post && delta_q && K > 0 && N > 0 && ld >= N, PU_ERR_BAD_ARG, "pu_trace_delta: bad argument");
  dim3 block(32, 8), grid(pu::cdiv(N, 32), pu::cdiv(N, 8));
  pu::trace_contract_kernel<<<grid, block, 0, pu::as_stream(stream)>>>(nullptr, pre, post, ld, K, nullptr, 0, nullptr, delta_q, N, K, 1);
  return pu::post_launch("pu_trace_delta");
}

int pu_trace_delta_tc(const float* pre, const float* post, long long ld, int K, float* delta_q, int N, void* stream) {
  PU_REQUIRE(pre && post && delta_q && K > 0 && N > 0 && ld >= N, PU_ERR_BAD_ARG, "pu_trace_delta_tc: bad argument");
  cudaStream_t st = pu::as_stream(stream);
  cudaError_t e = cudaMemsetAsync(delta_q, 0, sizeof(float) * ((size_t)N * N + N), st);
  if (e != cudaSuccess) {
    pu::set_error("pu_trace_delta_tc memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  const int tiles = pu::cdiv(N, pu::TD_T);
  int splits = pu::cdiv(2 * pu::kNumSMs, tiles * tiles);
  int kper = pu::cdiv(K, splits);
  kper = pu::cdiv(kper, pu::TD_KC) * pu::TD_KC;
  splits = pu::cdiv(K, kper);
  PU_REQUIRE(splits <= 65535, PU_ERR_UNSUPPORTED, "pu_trace_delta_tc: K too large");
  pu::trace_delta_mma_kernel<true><<<dim3(tiles, tiles, splits), 128, 0, st>>>(pre, post, ld, K, kper, delta_q, N);
  return pu::post_launch("pu_trace_delta_tc");
}

int pu_trace_apply(const float* hebb, const float* delta_q, int K_global, const float* eta, int rule, float* out, int N, void* stream) {
  PU_REQUIRE(hebb && delta_q && eta && out && K_global > 0 && N > 0, PU_ERR_BAD_ARG, "pu_trace_apply: bad argument");
  PU_REQUIRE(rule == PU_RULE_HEBB || rule == PU_RULE_OJA, PU_ERR_BAD_ARG, "pu_trace_apply: unknown rule %d", rule);
  dim3 block(32, 8), grid(pu::cdiv(N, 32), pu::cdiv(N, 8));
  pu::trace_apply_kernel<<<grid, block, 0, pu::as_stream(stream)>>>(hebb, delta_q, K_global, eta, rule, out, N);
  return pu::post_launch("pu_trace_apply");
}

int pu_trace_update_bwd(const float* hebb, const float* pre, const float* post, long long ld, int K, const float* eta, int rule,
                        const float* gout, float* ghebb, float* gpre, float* gpost, float* geta, int N, void* stream) {
  PU_REQUIRE(hebb && pre && post && eta && gout && K > 0 && N > 0 && ld >= N, PU_ERR_BAD_ARG, "pu_trace_update_bwd: bad argument");
  PU_REQUIRE(rule == PU_RULE_HEBB || rule == PU_RULE_OJA, PU_ERR_BAD_ARG, "pu_trace_update_bwd: unknown rule %d", rule);
  PU_REQUIRE(K <= 65535, PU_ERR_UNSUPPORTED, "pu_trace_update_bwd: K=%d > 65535", K);
  cudaStream_t st = pu::as_stream(stream);
  int rc;
  if (ghebb != nullptr || geta != nullptr) {
    if (geta != nullptr) {
      cudaError_t e = cudaMemsetAsync(geta, 0, sizeof(float), st);
      if (e != cudaSuccess) {
        pu::set_error("pu_trace_update_bwd memset: %s", cudaGetErrorString(e));
        return PU_ERR_CUDA;
      }
    }
    dim3 block(32, 8), grid(pu::cdiv(N, 32), pu::cdiv(N, 8));
    pu::trace_bwd_hebb_eta_kernel<<<grid, block, 0, st>>>(hebb, pre, post, ld, K, eta, rule, gout, ghebb, geta, N);
    rc = pu::post_launch("pu_trace_update_bwd hebb/eta");
    if (rc) return rc;
  }
  if (gpre != nullptr) {
    dim3 grid(pu::cdiv(N, 128), K);
    pu::trace_bwd_pre_kernel<<<grid, 128, 0, st>>>(post, ld, K, eta, gout, gpre, N);
    rc = pu::post_launch("pu_trace_update_bwd pre");
    if (rc) return rc;
  }
  if (gpost != nullptr) {
    dim3 grid(pu::cdiv(N, 128), K);
    pu::trace_bwd_post_kernel<<<grid, 128, 0, st>>>(hebb, pre, post, ld, K, eta, rule, gout, gpost, N);
    rc = pu::post_launch("pu_trace_update_bwd post");
    if (rc) return rc;
  }
  return PU_OK;
}

}  // extern "C"
