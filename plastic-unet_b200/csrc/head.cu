// head.cu — the plastic output head and its Hebb / Oja trace (reference unet_p.py:70-88 ==
// unet_p_res.py:116-134), batched over B maps of N x N.
//
//   Weff = w + alpha*hebb ; A = X @ Weff ; S = sigmoid(A)                       (forward)
//   gA = gS*S*(1-S) ; gX = gA @ Weff^T ; gWeff = X^T @ gA (split-K) ;
//   gw = gWeff ; galpha = gWeff*hebb ; ghebb = gWeff*alpha                      (backward)
//   trace: out = decay*hebb + eta/K * pre^T @ post, decay = 1-eta (Hebb) or 1 - eta*q_j/K (Oja),
//          one fused pass (contraction + epilogue), exactly the reference at K == 1.
//
// The GEMMs are strict-fp32 shared-memory-tiled FFMA kernels (64x64 tiles, 4x4 per thread, K staged in whole
// 128-deep chunks): the head's logits decide the thresholded masks, so it stays in full fp32.
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

__global__ void __launch_bounds__(128) trace_delta_mma_kernel(const float* __restrict__ pre, const float* __restrict__ post, long long ld, int K,
                                                              int kper, float* __restrict__ delta_q, int N) {
  __shared__ __align__(16) float Ps[TD_KC][TD_LD];
  __shared__ __align__(16) float Qs[TD_KC][TD_LD];
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
  for (int k0 = k_begin; k0 < k_end; k0 += TD_KC) {
    __syncthreads();
    // stage 32 rows x 64 columns of pre (columns i0..) and post (columns j0..): coalesced along the row
    for (int e = tid; e < TD_KC * TD_T; e += 128) {
      const int k = e / TD_T, c = e - k * TD_T;
      const bool kin = k0 + k < k_end;
      Ps[k][c] = (kin && i0 + c < N) ? __ldg(pre + (size_t)(k0 + k) * ld + i0 + c) : 0.f;
      Qs[k][c] = (kin && j0 + c < N) ? __ldg(post + (size_t)(k0 + k) * ld + j0 + c) : 0.f;
    }
    __syncthreads();
    if (blockIdx.y == 0 && tid < TD_T) {  // q_j = sum_k post_kj^2, once per column tile
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
          al[a][e] = f2tf32(v[e] - __uint_as_float(ah[a][e]));
        }
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {  // B[k][n = j] = Qs[k][j]
        const int n = wj + b * 8 + g;
        const float v[2] = {Qs[kk + t][n], Qs[kk + t + 4][n]};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          bh[b][e] = f2tf32(v[e]);
          bl[b][e] = f2tf32(v[e] - __uint_as_float(bh[b][e]));
        }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          mma_tf32_16x8x8(acc[a][b], al[a], bh[b]);  // small terms first
          mma_tf32_16x8x8(acc[a][b], ah[a], bl[b]);
          mma_tf32_16x8x8(acc[a][b], ah[a], bh[b]);
        }
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i0 + wi + a * 16 + g + (e >> 1) * 8, j = j0 + wj + b * 8 + 2 * t + (e & 1);
        if (i < N && j < N) atomicAdd(delta_q + (size_t)i * N + j, acc[a][b][e]);
      }
  if (blockIdx.y == 0 && tid < TD_T && j0 + tid < N) atomicAdd(delta_q + (size_t)N * N + j0 + tid, qsum);
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
  pu::trace_delta_mma_kernel<<<dim3(tiles, tiles, splits), 128, 0, st>>>(pre, post, ld, K, kper, delta_q, N);
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
