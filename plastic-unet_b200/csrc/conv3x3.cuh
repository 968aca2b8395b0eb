// Argument blocks shared by the conv3x3 kernel families (FFMA fp32 and tcgen05 tf32).
#pragma once
#include "pu_common.cuh"

namespace pu {

struct Conv3x3Args {
  View s0, s1;        // concatenated sources (s1.p may be null)
  const float* wp;    // packed weights [9][Cin][Cout]
  const float* bias;  // [Cout] | null
  const float* res;   // [B,H,W,Cout] | null, added before the ReLU
  ViewW d0, d1;       // channel-split destinations (d1.p may be null)
  // packed ReLU masks, one BYTE per (pixel, 8-channel group), bit j = channel 8g+j was > 0:
  const unsigned char* mask0;  // | null: geometry of d0 with C/8 bytes per pixel; the stored value is zeroed where the bit is clear
  const unsigned char* mask1;  // | null: same for d1
  unsigned char* mask_out;     // | null: geometry of d0 (needs d1.p == null, Cout % 8 == 0): bit = (stored output > 0)
  int B, H, W, Cin, Cout, relu, round_out;
  int wfmt;  // 0: wp packed by pu_pack_w3x3; 1: wp is the raw OIHW weight (forward); 2: raw OIHW, conv is the dgrad
  int tilesX, tilesY;
};

struct WgradArgs {
  View s0, s1;
  const float* g;  // [B,H,W,Cout]
  float* dw;       // OIHW
  float* db;       // [Cout] | null
  int B, H, W, Cin, Cout;
  int tilesX, tilesY, ntiles;
  int accum;  // 1: add to dw / db (zeroed by the caller) instead of overwriting them
  int debug;  // PU_WG_DEBUG experiments: 1 = skip the final atomics, 2 = skip the MMAs, 4 = staging only (no fragment loads, no MMAs)
};

int conv3x3_fwd_ffma(const Conv3x3Args& a, cudaStream_t st);
// one-input-channel stem (stem.cu)
bool conv3x3_c1_ok(int Cin, int Cout);
int conv3x3_c1_fwd(const Conv3x3Args& a, cudaStream_t st);
int conv3x3_c1_wgrad(const WgradArgs& a, cudaStream_t st, int math);
int conv3x3_wgrad_ffma(const WgradArgs& a, cudaStream_t st, int math);
// TMA-fed mma.sync weight gradient (conv3x3_wgrad_tma.cu): TF32 mode, channel counts that are multiples of 8
bool conv3x3_wgrad_tma_ok(const WgradArgs& a);
int conv3x3_wgrad_tma(const WgradArgs& a, cudaStream_t st);
bool conv3x3_c1_wgrad_tma_ok(const WgradArgs& a);               // the one-channel stem through TMA + MMA (aligned inputs)
int conv3x3_c1_wgrad_tma(const WgradArgs& a, cudaStream_t st);  // dw / db zeroed by the caller
// tcgen05 path (conv3x3_tc.cu); returns PU_ERR_UNSUPPORTED when the shape does not fit it
int conv3x3_fwd_tc(const Conv3x3Args& a, cudaStream_t st);
bool conv3x3_tc_ok(int C0, int C1, int Cout, int Cd0, int Cd1);
bool conv3x3_tc_resident(int C0, int C1, int Cout, int H, int W);
bool conv3x3_tc_flat(int B, int H, int W, int C0, int C1, int Cout);  // will pu_conv3x3_fwd use the flat (one MMA per tap) mode?
long long conv3x3_tc_weight_floats(int C0, int C1, int Cout, bool flat);
int conv3x3_tc_pack(const float* w_oihw, float* out, int Cout_w, int Cin_w, int transpose, int C0, bool flat, cudaStream_t st);

}  // namespace pu
