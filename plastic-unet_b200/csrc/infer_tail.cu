// infer_tail.cu — the inference tail of the hot path: integer / byte work, bit-exact.
//
//   threshold sweep + confusion counts   reference eval.py:48-52 (31 thresholds), utils/iou_metric.py:6-24 (fast_iou_metric),
//                                         utils/iou_metric.py:26-43 (histogram2d over bins [0,0.5,1])
//   mask threshold + run-length encoding  reference infer.py:81,88,99 (mask > mask_threshold), utils/rle_encode.py:6-17
//                                         (column-major = Fortran order, 1-based starts, (start, length) pairs)
//
// Everything here produces INTEGERS (counts, run positions, mask bytes); the float arithmetic that turns counts into IoU
// scores is a handful of flops per image and stays on the host, written exactly as the reference writes it, so the scores
// are bit-identical too (pu_b200/infer_tail.py).  Comparisons are done in double: numpy compares the float32 predictions
// with float64 thresholds in eval.py:52 (thresholds come from a float64 ndarray) and with a float32-rounded threshold in
// infer.py:81 (python scalar) — the caller passes the threshold already rounded the way the reference's numpy would.
#include "pu_common.cuh"

namespace pu {

// ---- threshold sweep ------------------------------------------------------------------------------------------------
// thresholds sorted ascending: { j : pred > thr[j] } is a prefix [0, idx) of the sorted list, so one histogram over idx per
// label class gives the counts of ALL thresholds by a suffix sum — each prediction is read once for the whole sweep.
// label classes: mode 0 = np.histogram bins [0,0.5,1] (iou_metric.py:34-36): [0,0.5) -> 0, [0.5,1] -> 1, else dropped (2);
//                mode 1 = `A > 0` (iou_metric.py:10).
constexpr int kMaxThr = 64;

__global__ void __launch_bounds__(256) threshold_hist_kernel(const float* __restrict__ pred, const float* __restrict__ label,
                                                              const double* __restrict__ thr, int T, int label_mode, long long npix,
                                                              int* __restrict__ hist /* [B][3][T+1], zeroed */) {
  __shared__ double s_thr[kMaxThr];
  __shared__ int s_hist[3 * (kMaxThr + 1)];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < T; i += blockDim.x) s_thr[i] = thr[i];
  for (int i = threadIdx.x; i < 3 * (T + 1); i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const float* p = pred + (long long)b * npix;
  const float* l = label + (long long)b * npix;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)__ldg(p + i);
    const float lv = __ldg(l + i);
    // idx = number of thresholds with v > thr (binary search on the ascending list; NaN compares false -> 0)
    int lo = 0, hi = T;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (v > s_thr[mid]) lo = mid + 1; else hi = mid;
    }
    int cls;
    if (label_mode == 0) cls = (lv >= 0.f && lv < 0.5f) ? 0 : ((lv >= 0.5f && lv <= 1.f) ? 1 : 2);
    else cls = lv > 0.f ? 1 : 0;
    atomicAdd(&s_hist[cls * (T + 1) + lo], 1);
  }
  __syncthreads();
  int* h = hist + (long long)b * 3 * (T + 1);
  for (int i = threadIdx.x; i < 3 * (T + 1); i += blockDim.x)
    if (s_hist[i]) atomicAdd(&h[i], s_hist[i]);
}

// counts[b][j] = {c00, c01, c10, c11, np1, n}: c(true class)(pred class) over pixels with a valid label, np1 = predicted
// ones over ALL pixels (np.histogram(y_pred) sees every pixel), n = pixels
__global__ void threshold_counts_kernel(const int* __restrict__ hist, int T, int B, long long npix, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int b = i / T, j = i % T;
  const int* h = hist + (long long)b * 3 * (T + 1);
  int tot[3] = {0, 0, 0}, one[3] = {0, 0, 0};
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k <= T; ++k) {
      const int v = h[c * (T + 1) + k];
      tot[c] += v;
      if (k > j) one[c] += v;  // pred > thr[j]  <=>  idx > j
    }
  int* o = counts + (long long)i * 6;
  o[0] = tot[0] - one[0];
  o[1] = one[0];
  o[2] = tot[1] - one[1];
  o[3] = one[1];
  o[4] = one[0] + one[1] + one[2];
  o[5] = (int)npix;
}

// ---- mask threshold + column-major run-length encoding --------------------------------------------------------------
// One CTA per image.  Stage 1: coalesced row-major reads, 32x32 tiles transposed with ballots into a bit array in
// Fortran order (k = c*R + r) in shared memory (+ optional uint8 mask, row-major, infer.py:88).  Stage 2: transitions
// d[k] = px[k] ^ px[k-1] over k = 0..n (px[-1] = px[n] = 0, rle_encode.py:14) -> popcount, block scan, ordered emit of
// k+1 (rle_encode.py:15), then every odd entry becomes a length (rle_encode.py:16).
constexpr int kRleThreads = 1024;

__global__ void __launch_bounds__(kRleThreads) mask_rle_kernel(const float* __restrict__ pred, double thr, int R, int C,
                                                               unsigned char* __restrict__ mask /* nullable */,
                                                               int* __restrict__ runs, int cap, int* __restrict__ count) {
  extern __shared__ unsigned int s_bits[];  // nw words + 32 scan slots
  const int n = R * C;
  const int nw = (n + 1 + 31) / 32 + 1;
  unsigned int* s_scan = s_bits + nw;
  const int b = blockIdx.x;
  const float* p = pred + (long long)b * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kRleThreads / 32;
  for (int i = threadIdx.x; i < nw; i += kRleThreads) s_bits[i] = 0u;
  __syncthreads();
  const int tr_n = (R + 31) / 32, tc_n = (C + 31) / 32;
  for (int t = warp; t < tr_n * tc_n; t += nwarps) {
    const int tr = t / tc_n, tc = t % tc_n;
    const int c = tc * 32 + lane;
    unsigned int colw = 0u;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const int r = tr * 32 + i;
      bool v = false;
      if (r < R && c < C) {
        v = (double)__ldg(p + (long long)r * C + c) > thr;
        if (mask != nullptr) mask[(long long)b * n + (long long)r * C + c] = v ? 1 : 0;
      }
      const unsigned int w = __ballot_sync(0xffffffffu, v);
      colw |= ((w >> lane) & 1u) << i;
    }
    if (c < C && colw) {
      const int k0 = c * R + tr * 32;
      const int sh = k0 & 31;
      atomicOr(&s_bits[k0 >> 5], colw << sh);
      if (sh && (colw >> (32 - sh))) atomicOr(&s_bits[(k0 >> 5) + 1], colw >> (32 - sh));
    }
  }
  __syncthreads();
  // stage 2
  const int nwt = (n + 1 + 31) / 32;  // words that hold positions 0..n
  const int wpt = (nwt + kRleThreads - 1) / kRleThreads;
  const int w0 = min(threadIdx.x * wpt, nwt), w1 = min(w0 + wpt, nwt);
  unsigned int cnt = 0;
  for (int wi = w0; wi < w1; ++wi) {
    const unsigned int w = s_bits[wi];
    const unsigned int carry = wi ? (s_bits[wi - 1] >> 31) : 0u;
    cnt += __popc(w ^ ((w << 1) | carry));
  }
  // block exclusive scan of cnt
  unsigned int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) s_scan[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned int v = s_scan[lane];
    unsigned int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += u;
    }
    s_scan[lane] = s - v;  // exclusive warp offsets
    if (lane == 31) s_scan[32] = s;  // total
  }
  __syncthreads();
  unsigned int off = s_scan[warp] + inc - cnt;
  const unsigned int total = s_scan[32];
  int* out = runs + (long long)b * cap;
  for (int wi = w0; wi < w1; ++wi) {
    const unsigned int w = s_bits[wi];
    const unsigned int carry = wi ? (s_bits[wi - 1] >> 31) : 0u;
    unsigned int d = w ^ ((w << 1) | carry);
    while (d) {
      const int bit = __ffs(d) - 1;
      d &= d - 1;
      if (off < (unsigned int)cap) out[off] = wi * 32 + bit + 1;
      ++off;
    }
  }
  __syncthreads();
  const unsigned int m = total < (unsigned int)cap ? total : (unsigned int)cap;
  for (unsigned int j = 2 * threadIdx.x + 1; j < m; j += 2 * kRleThreads) out[j] -= out[j - 1];
  if (threadIdx.x == 0) count[b] = total <= (unsigned int)cap ? (int)total : -(int)total;
}

}  // namespace pu

extern "C" {

int pu_threshold_counts(const float* pred, const float* label, const double* thr_sorted, int T, int label_mode, int B, long long npix,
                        int* hist_ws, int* counts, void* stream) {
  PU_REQUIRE(pred && label && thr_sorted && hist_ws && counts, PU_ERR_BAD_ARG, "pu_threshold_counts: null pointer");
  PU_REQUIRE(T >= 1 && T <= pu::kMaxThr && B >= 1 && B <= 65535 && npix >= 1 && npix < (1ll << 31) && (label_mode == 0 || label_mode == 1),
             PU_ERR_BAD_ARG, "pu_threshold_counts: need 1 <= T <= %d, 1 <= B <= 65535, npix < 2^31, label_mode in {0,1}", pu::kMaxThr);
  cudaStream_t st = pu::as_stream(stream);
  cudaError_t e = cudaMemsetAsync(hist_ws, 0, sizeof(int) * (size_t)B * 3 * (T + 1), st);
  if (e != cudaSuccess) {
    pu::set_error("pu_threshold_counts memset: %s", cudaGetErrorString(e));
    return PU_ERR_CUDA;
  }
  // enough CTAs to fill the machine: B images x chunks, each thread >= 8 pixels
  int chunks = pu::cdiv(npix, 256 * 8);
  const int want = pu::cdiv(4 * pu::kNumSMs, B);
  chunks = chunks < 1 ? 1 : (chunks > want ? want : chunks);
  pu::threshold_hist_kernel<<<dim3(chunks, B), 256, 0, st>>>(pred, label, thr_sorted, T, label_mode, npix, hist_ws);
  int rc = pu::post_launch("pu_threshold_counts hist");
  if (rc) return rc;
  pu::threshold_counts_kernel<<<pu::cdiv((long long)B * T, 128), 128, 0, st>>>(hist_ws, T, B, npix, counts);
  return pu::post_launch("pu_threshold_counts");
}

long long pu_mask_rle_smem_bytes(int R, int C) {
  const long long n = (long long)R * C;
  return (long long)sizeof(unsigned int) * ((n + 1 + 31) / 32 + 1 + 33);
}

int pu_mask_rle(const float* pred, double thr, int B, int R, int C, unsigned char* mask, int* runs, int cap, int* count, void* stream) {
  PU_REQUIRE(pred && runs && count, PU_ERR_BAD_ARG, "pu_mask_rle: null pointer");
  PU_REQUIRE(B >= 1 && R >= 1 && C >= 1 && cap >= 2, PU_ERR_BAD_ARG, "pu_mask_rle: bad dims");
  const long long smem = pu_mask_rle_smem_bytes(R, C);
  PU_REQUIRE(smem <= 227 * 1024, PU_ERR_UNSUPPORTED, "pu_mask_rle: image %dx%d needs %lld B of shared memory (max 227 KB: <= 1.8 M pixels)", R, C,
             smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pu::mask_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      pu::set_error("pu_mask_rle smem attribute: %s", cudaGetErrorString(e));
      return PU_ERR_CUDA;
    }
  }
  pu::mask_rle_kernel<<<B, pu::kRleThreads, (size_t)smem, pu::as_stream(stream)>>>(pred, thr, R, C, mask, runs, cap, count);
  return pu::post_launch("pu_mask_rle");
}

}  // extern "C"
