// C-ABI entry points of the 3x3 convolution family: argument validation + math-mode dispatch.
#include "conv3x3.cuh"

namespace {

int check_view(const char* name, const float* p, int Hs, int Ws, int C, int oy, int ox, int H, int W) {
  PU_REQUIRE(p != nullptr, PU_ERR_BAD_ARG, "%s: null pointer", name);
  PU_REQUIRE(C > 0 && Hs > 0 && Ws > 0, PU_ERR_BAD_ARG, "%s: non-positive dims", name);
  PU_REQUIRE(oy >= 0 && ox >= 0 && oy + H <= Hs && ox + W <= Ws, PU_ERR_BAD_ARG,
             "%s: window (%d+%d,%d+%d) exceeds tensor %dx%d", name, oy, H, ox, W, Hs, Ws);
  PU_REQUIRE(pu::aligned16(p), PU_ERR_BAD_ARG, "%s: pointer not 16-byte aligned", name);
  return PU_OK;
}

}  // namespace

extern "C" {

int pu_conv3x3_fwd(const float* src0, int H0, int W0, int C0, int oy0, int ox0,
                   const float* src1, int H1, int W1, int C1, int oy1, int ox1,
                   const float* wp, const float* bias, const float* res, int flags,
                   float* dst0, int Hd0, int Wd0, int Cd0, int oyd0, int oxd0,
                   float* dst1, int Hd1, int Wd1, int Cd1, int oyd1, int oxd1,
                   const unsigned char* mask0, const unsigned char* mask1, unsigned char* mask_out,
                   int B, int H, int W, int Cout, int math, int wfmt, void* stream) {
  PU_REQUIRE(B > 0 && H > 0 && W > 0 && Cout > 0 && wp != nullptr, PU_ERR_BAD_ARG, "pu_conv3x3_fwd: bad dims");
  PU_REQUIRE(wfmt >= 0 && wfmt <= 2, PU_ERR_BAD_ARG, "pu_conv3x3_fwd: unknown weight format %d", wfmt);
  int rc = check_view("pu_conv3x3_fwd src0", src0, H0, W0, C0, oy0, ox0, H, W);
  if (rc) return rc;
  if (src1 != nullptr) {
    rc = check_view("pu_conv3x3_fwd src1", src1, H1, W1, C1, oy1, ox1, H, W);
    if (rc) return rc;
  } else {
    C1 = 0;
  }
  rc = check_view("pu_conv3x3_fwd dst0", dst0, Hd0, Wd0, Cd0, oyd0, oxd0, H, W);
  if (rc) return rc;
  if (dst1 != nullptr) {
    rc = check_view("pu_conv3x3_fwd dst1", dst1, Hd1, Wd1, Cd1, oyd1, oxd1, H, W);
    if (rc) return rc;
  } else {
    Cd1 = 0;
  }
  PU_REQUIRE(Cd0 + Cd1 == Cout, PU_ERR_BAD_ARG, "pu_conv3x3_fwd: destination channels %d+%d != Cout %d", Cd0, Cd1, Cout);
  PU_REQUIRE(math == PU_MATH_FP32 || math == PU_MATH_TF32, PU_ERR_BAD_ARG, "pu_conv3x3_fwd: unknown math mode %d", math);

  pu::Conv3x3Args a;
  a.s0 = pu::View{src0, H0, W0, C0, oy0, ox0};
  a.s1 = pu::View{src1, H1, W1, C1, oy1, ox1};
  a.wp = wp;
  a.bias = bias;
  a.res = res;
  a.d0 = pu::ViewW{dst0, Hd0, Wd0, Cd0, oyd0, oxd0};
  a.d1 = pu::ViewW{dst1, Hd1, Wd1, Cd1, oyd1, oxd1};
  PU_REQUIRE((mask0 == nullptr || Cd0 % 8 == 0) && (mask1 == nullptr || (dst1 != nullptr && Cd1 % 8 == 0)), PU_ERR_BAD_ARG,
             "pu_conv3x3_fwd: packed masks need destination channel counts that are multiples of 8");
  PU_REQUIRE(mask_out == nullptr || (dst1 == nullptr && Cout % 8 == 0), PU_ERR_BAD_ARG,
             "pu_conv3x3_fwd: mask_out needs a single destination with Cout %% 8 == 0");
  a.mask0 = mask0;
  a.mask1 = mask1;
  a.mask_out = mask_out;
  a.B = B; a.H = H; a.W = W; a.Cin = C0 + C1; a.Cout = Cout;
  a.relu = (flags & PU_FLAG_RELU) ? 1 : 0;
  a.round_out = (flags & PU_FLAG_ROUND_TF32) ? 1 : 0;
  a.wfmt = wfmt;
  a.tilesX = a.tilesY = 0;
  // no silent fallback: PU_MATH_TF32 means the tcgen05 kernel (wp must be in its layout) or an error
  if (math == PU_MATH_TF32) return pu::conv3x3_fwd_tc(a, pu::as_stream(stream));
  return pu::conv3x3_fwd_ffma(a, pu::as_stream(stream));
}

int pu_conv3x3_wgrad(const float* src0, int H0, int W0, int C0, int oy0, int ox0,
                     const float* src1, int H1, int W1, int C1, int oy1, int ox1,
                     const float* g, float* dw_oihw, float* db, int B, int H, int W, int Cout, int math, void* stream) {
  PU_REQUIRE(B > 0 && H > 0 && W > 0 && Cout > 0 && g != nullptr && dw_oihw != nullptr, PU_ERR_BAD_ARG, "pu_conv3x3_wgrad: bad dims");
  PU_REQUIRE(pu::aligned16(g), PU_ERR_BAD_ARG, "pu_conv3x3_wgrad: g not 16-byte aligned");
  int rc = check_view("pu_conv3x3_wgrad src0", src0, H0, W0, C0, oy0, ox0, H, W);
  if (rc) return rc;
  if (src1 != nullptr) {
    rc = check_view("pu_conv3x3_wgrad src1", src1, H1, W1, C1, oy1, ox1, H, W);
    if (rc) return rc;
  } else {
    C1 = 0;
  }
  const int accum = (math & PU_MATH_ACCUM) ? 1 : 0;
  math &= ~PU_MATH_ACCUM;
  PU_REQUIRE(math == PU_MATH_FP32 || math == PU_MATH_TF32, PU_ERR_BAD_ARG, "pu_conv3x3_wgrad: unknown math mode %d", math);
  pu::WgradArgs a;
  a.accum = accum;
  a.debug = 0;
  a.s0 = pu::View{src0, H0, W0, C0, oy0, ox0};
  a.s1 = pu::View{src1, H1, W1, C1, oy1, ox1};
  a.g = g;
  a.dw = dw_oihw;
  a.db = db;
  a.B = B; a.H = H; a.W = W; a.Cin = C0 + C1; a.Cout = Cout;
  a.tilesX = a.tilesY = a.ntiles = 0;
  // PU_MATH_FP32: FFMA2 on the CUDA cores.  PU_MATH_TF32: TMA-fed warp-level TF32 MMAs (mma.sync m16n8k8) for channel
  // counts that are multiples of 8 (conv3x3_wgrad_tma.cu, which also explains why tcgen05 does not fit this contraction).
  return pu::conv3x3_wgrad_ffma(a, pu::as_stream(stream), math);
}

int pu_conv3x3_tc_ok(int C0, int C1, int Cout, int Cd0, int Cd1) { return pu::conv3x3_tc_ok(C0, C1, Cout, Cd0, Cd1) ? 1 : 0; }

int pu_conv3x3_tc_resident(int C0, int C1, int Cout, int H, int W) { return pu::conv3x3_tc_resident(C0, C1, Cout, H, W) ? 1 : 0; }

}  // extern "C"
