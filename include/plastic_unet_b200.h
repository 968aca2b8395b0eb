/*
 * plastic_unet_b200.h — C-ABI of the B200-native Plastic U-Net hot path.
 *
 * Drop-in boundary: the reference has no FFI; its hot path sits behind the
 * Python nn.Module surface of package `unet` (reference src/unet/__init__.py:1-2).  Every
 * entry point below replaces one stock-PyTorch call site inside that package; the reference
 * file:line it replaces is cited per function.  The Python side (plastic-unet_b200/pu_b200)
 * binds these with ctypes and wraps them as torch.library custom ops.
 *
 * Conventions
 *  - All tensors are fp32, device-resident, **NHWC** (channels contiguous) unless stated.
 *  - The caller (PyTorch) owns every buffer; the library allocates nothing persistent.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*); no internal
 *    synchronisation, safe under CUDA-graph capture.
 *  - Return 0 on success, negative pu_status otherwise; pu_last_error() has the text.
 *    No exception or abort crosses this boundary.
 *  - A "view" (H?,W?,oy?,ox?) describes a tensor [B,H?,W?,C?] of which the op uses the window
 *    starting at pixel (oy?,ox?) of extent H x W (crop fused as a pointer offset;
 *    reference unet_p.py:161-165, unet_p_res.py:215-218).
 */
#ifndef PLASTIC_UNET_B200_H_
#define PLASTIC_UNET_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PU_OK = 0,
  PU_ERR_BAD_ARG = -1,      /* null pointer / non-positive dim / misaligned */
  PU_ERR_UNSUPPORTED = -2,  /* shape outside what the kernel family handles  */
  PU_ERR_CUDA = -3,         /* CUDA runtime/driver error captured at launch  */
  PU_ERR_NO_DEVICE = -4
} pu_status;

#define PU_RULE_HEBB 0
#define PU_RULE_OJA 1

/* op flags */
#define PU_FLAG_RELU 1       /* fused ReLU in the epilogue                                            */
#define PU_FLAG_ROUND_TF32 2 /* round the op's output to TF32 (RN): producers of tensor-core operands  */
#define PU_FLAG_MASK_IN 4    /* backward kernels: zero the input gradient where the op's input x <= 0, i.e. apply the
                                ReLU mask of the PRODUCER of x (premasked-gradient protocol)          */

#define PU_FLAG_ACCUM_GRADS 16 /* backward kernels: ADD the parameter gradients to dw / db instead of overwriting them (the caller
                                 zeroed them, e.g. the slots of a flat gradient arena): no memset launches.  Supported by the
                                 TF32 paths of pu_conv3x3_wgrad (as PU_MATH_ACCUM in `math`) and pu_convT2x2s2_bwd.          */
#define PU_FLAG_TF32_MATH 8  /* transposed convolutions: TF32 tensor-core math (mma.sync), fp32 accumulate; the model's TF32
                                mode.  Without it (or for unsupported shapes) the fp32 CUDA-core kernels run.            */

/* conv3x3 weight operand formats */
#define PU_W_PACKED 0
#define PU_W_OIHW 1
#define PU_W_OIHW_DGRAD 2

/* conv3x3 math modes */
#define PU_MATH_FP32 0 /* CUDA-core FFMA, strict fp32 (parity mode, any shape)        */
#define PU_MATH_TF32 1 /* tcgen05 kind::tf32 implicit GEMM, fp32 accumulate in TMEM   */
#define PU_MATH_ACCUM 0x100 /* pu_conv3x3_wgrad only, OR-ed into math: accumulate into dw / db (see PU_FLAG_ACCUM_GRADS)          */
#define PU_MATH_TF32_FLAT 2 /* pu_pack_w3x3 only: the weight image of a PU_MATH_TF32 conv for which pu_conv3x3_tc_flat() is 1 */

/* ---- library management --------------------------------------------------------------- */
int pu_version(void);
const char* pu_last_error(void);
/* number of kernels this library has launched (or captured) since load / last reset */
long long pu_launch_count(void);
void pu_reset_launch_count(void);
/* 1 if the tcgen05/TMA conv path is usable on the current device (sm_100, driver entry points) */
int pu_tc_available(void);

/* ---- layout -----------------------------------------------------------------------------
 * NCHW <-> NHWC (module entry/exit when n_channels > 1; reference keeps NCHW throughout). */
int pu_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, void* stream);
int pu_nhwc_to_nchw(const float* x, float* y, int B, int C, int H, int W, void* stream);

/* ---- 3x3 convolution, stride 1, pad 1 -----------------------------------------------------
 * replaces nn.Conv2d(k=3,padding=1) (+ReLU, +residual add, +torch.cat/F.pad crop) at
 * reference unet_p.py:105-116,161-166 and unet_p_res.py:150-158,186-189,215-219,230,264.
 *
 * pu_pack_w3x3: OIHW weight [Cout,Cin,3,3] -> the operand layout of the conv that will consume it:
 *   transpose=0: the forward conv (input channels Cin split C0 | Cin-C0 over the two sources, output Cout);
 *   transpose=1: its dgrad (a conv with Cout input channels, Cin output channels, taps flipped; C0 ignored).
 *   math=PU_MATH_FP32: [9][Cin'][Cout'];  math=PU_MATH_TF32: the tcgen05 B-operand tiles
 *   [co block][K chunk][tap][channel group][N][4], rounded to TF32 (requires pu_conv3x3_tc_ok).          */
int pu_pack_w3x3(const float* w_oihw, float* w_packed, int Cout, int Cin, int transpose, int math, int C0, void* stream);
/* floats needed for the packed buffer of the call above */
long long pu_pack_w3x3_floats(int Cout, int Cin, int transpose, int math, int C0);
/* 1 if a conv with these source / destination channel counts can run on the tcgen05 path (PU_MATH_TF32) */
int pu_conv3x3_tc_ok(int C0, int C1, int Cout, int Cd0, int Cd1);
/* 1 if pu_conv3x3_fwd(PU_MATH_TF32) will run this problem in the "flat" mode (one MMA per tap, N = 64, 512-pixel tiles:
 * the wide, tensor-bound layers); its packed weights must then be produced with math = PU_MATH_TF32_FLAT.             */
int pu_conv3x3_tc_flat(int B, int H, int W, int C0, int C1, int Cout);
/* 1 if, additionally, the tcgen05 kernel can build its weight tiles from the raw OIHW tensor (they fit in shared memory) */
int pu_conv3x3_tc_resident(int C0, int C1, int Cout, int H, int W);
/* Host-only (no device needed): the tile plan pu_conv3x3_fwd would use for this problem, for tests and tuning.
 * out17 = {TH, TW, PW, tilesX, tilesY, blocks per tile, COLS, N3 (B-tile rows), stage A bytes, stage weight bytes, resident weight
 * bytes, TMEM columns, K chunks, co blocks, stages, dynamic shared memory bytes, fold (1: kx folded into N, 96-pixel blocks;
 * 0: flat, one MMA per tap, 128-pixel blocks)}.  resident != 0: the weights are built
 * in shared memory from the raw OIHW tensor (PU_W_OIHW / PU_W_OIHW_DGRAD).  PU_ERR_UNSUPPORTED if the shape does not fit. */
int pu_conv3x3_tc_plan(int B, int H, int W, int C0, int C1, int Cout, int resident, int* out17);
/* Host-only: the plan of the TMA-fed weight-gradient kernel behind pu_conv3x3_wgrad(PU_MATH_TF32) for this problem.
 * out12 = {input chunks per CTA, output tiles per CTA, tile width, tile height, images per tile, tilesX, tilesY, image tiles,
 * pipeline stages, grid x, grid y, dynamic shared memory bytes}.  PU_ERR_UNSUPPORTED if the channel counts are not multiples of 8. */
int pu_conv3x3_wgrad_plan(int B, int H, int W, int C0, int C1, int Cout, int* out12);

/* y = act( conv3x3(cat[src0,src1]) + bias + res ), written channel-split into dst0|dst1 (flags: PU_FLAG_*).
 * wfmt selects what `wp` points at: PU_W_PACKED (output of pu_pack_w3x3 for this math mode), PU_W_OIHW (the raw
 * [Cout, C0+C1, 3, 3] weight: operand tiles are built inside the kernel, no pack launch) or PU_W_OIHW_DGRAD (the
 * raw [C0, Cout, 3, 3] weight of the forward conv whose dgrad this call computes).  With PU_MATH_TF32 the raw
 * formats need pu_conv3x3_tc_resident(C0, C1, Cout, H, W).
 * src1, bias, res, dst1 may be NULL.  wp is the packed weight [9][C0+C1][Cout].
 * dst views may be larger than HxW (their border is NOT written — caller zero-fills).
 * mask0 / mask1 (may be NULL): PACKED ReLU masks with the geometry of dst0 / dst1: one byte per (pixel, 8-channel group),
 * bit j set = channel 8g+j of the tensor this gradient belongs to was > 0; where the bit is clear the stored value is 0
 * (dgrad calls: the ReLU mask of the layer that produced the source, applied in the epilogue — 1/32 of the bytes of
 * re-reading the fp32 activation).  mask_out (may be NULL; needs dst1 == NULL and Cout % 8 == 0): the forward conv
 * writes the packed mask of its own (post-ReLU) output there, geometry of dst0 with Cout/8 bytes per pixel.       */
int pu_conv3x3_fwd(const float* src0, int H0, int W0, int C0, int oy0, int ox0,
                   const float* src1, int H1, int W1, int C1, int oy1, int ox1,
                   const float* wp, const float* bias, const float* res, int flags,
                   float* dst0, int Hd0, int Wd0, int Cd0, int oyd0, int oxd0,
                   float* dst1, int Hd1, int Wd1, int Cd1, int oyd1, int oxd1,
                   const unsigned char* mask0, const unsigned char* mask1, unsigned char* mask_out,
                   int B, int H, int W, int Cout, int math, int wfmt, void* stream);

/* dw_oihw[Cout, C0+C1, 3, 3] = sum_{b,y,x} g[b,y,x,co] * cat[src0,src1][b,y+ky-1,x+kx-1,ci]
 * (overwrites dw).  g is [B,H,W,Cout] dense.  db (may be NULL) receives the bias gradient sum_pixels g.
 * math | PU_MATH_ACCUM: dw and db are ADDED to (no memset launches); PU_ERR_UNSUPPORTED where the shape's kernel cannot.  */
int pu_conv3x3_wgrad(const float* src0, int H0, int W0, int C0, int oy0, int ox0,
                     const float* src1, int H1, int W1, int C1, int oy1, int ox1,
                     const float* g, float* dw_oihw, float* db, int B, int H, int W, int Cout,
                     int math, void* stream);

/* g = (flags & RELU) ? dy * (y > 0) : dy [rounded to TF32 if flags & ROUND]; dbias[c] = sum_pixels g  (g may alias
 * dy; dbias may be NULL; y may be NULL iff no RELU flag).  Autograd of the fused ReLU + bias epilogue above.          */
int pu_relu_bwd_bias(const float* dy, const float* y, float* g, float* dbias,
                     long long npix, int C, int flags, void* stream);

/* ---- 1x1 convolution (reference unet_p.py:173, unet_p_res.py:194; CoordConv stem
 * coord_conv_script.py:104-126).  coords != 0 appends the AddCoords channels
 * (coord_conv_script.py:69-96): xx = 2*j/(W-1)-1, yy = 2*i/(H-1)-1, [rr if coords == 3]
 * analytically — they are never materialised.  w is [Cout, Cin + coords].                  */
int pu_conv1x1_fwd(const float* x, const float* w, const float* bias, float* y,
                   int B, int H, int W, int Cin, int Cout, int coords, int flags, void* stream);
/* g is the (already ReLU-masked) output gradient.  dx and db may be NULL. dw [Cout,Cin+coords], db [Cout]
 * overwritten.  ws: caller-provided scratch of Cout*(Cin+coords+1) floats (<= 256).              */
int pu_conv1x1_bwd(const float* x, const float* w, const float* g, float* dx, float* dw, float* db, float* ws,
                   int B, int H, int W, int Cin, int Cout, int coords, int flags, void* stream);
/* flags & PU_FLAG_DEFER_FINISH: pu_conv1x1_bwd leaves the parameter gradients in ws ([Cout][Cin+coords+1], bias last) and the
 * caller scatters them into dw / db with this call — e.g. on a side stream, off the data-gradient chain of the backward pass. */
#define PU_FLAG_DEFER_FINISH 32
int pu_conv1x1_dw_finish(const float* ws, float* dw, float* db, int Cin, int Cout, int coords, void* stream);

/* ---- transposed convolutions ---------------------------------------------------------------
 * 2x2 stride 2 (reference unet_p.py:155): w is PyTorch [Cin,Cout,2,2]; y is [B,2H,2W,Cout]. */
int pu_convT2x2s2_fwd(const float* x, const float* w, const float* bias, float* y,
                      int B, int H, int W, int Cin, int Cout, int flags, void* stream);
int pu_convT2x2s2_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db,
                      int B, int H, int W, int Cin, int Cout, int flags, void* stream);
/* 3x3 stride 2 pad 0 (reference unet_p_res.py:207): full output is (2H+1)x(2W+1); the op writes
 * only the window [oy,oy+Ho) x [ox,ox+Wo) of it (the F.pad crop of unet_p_res.py:215-217 fused).
 * chan_scale (may be NULL) is a per-(b,co) multiplier [B,Cout] applied after bias (Dropout2d). */
int pu_convT3x3s2_fwd(const float* x, const float* w, const float* bias, const float* chan_scale, float* y,
                      int B, int H, int W, int Cin, int Cout, int Ho, int Wo, int oy, int ox, int flags, void* stream);
int pu_convT3x3s2_bwd(const float* x, const float* w, const float* dy, const float* chan_scale,
                      float* dx, float* dw, float* db,
                      int B, int H, int W, int Cin, int Cout, int Ho, int Wo, int oy, int ox, void* stream);

/* Zero insertion: the (oy, ox, Ho, Wo) window of the (2H+1)x(2W+1) canvas holding x at the odd coordinates.  With it
   ConvTranspose2d(k=3, s=2, p=0) + the reference's crop (unet_p_res.py:207,214-217) = pu_conv3x3_fwd with PU_W_OIHW_DGRAD
   on the tcgen05 path (TF32 mode).  _bwd gathers the gradient back.  C % 4 == 0. */
int pu_zero_insert2x_fwd(const float* x, float* z, int B, int H, int W, int C, int Ho, int Wo, int oy, int ox, void* stream);
int pu_zero_insert2x_bwd(const float* dz, float* dx, int B, int H, int W, int C, int Ho, int Wo, int oy, int ox, void* stream);

/* ---- pooling / resampling --------------------------------------------------------------------
 * MaxPool2d(2) floor mode (reference unet_p.py:139, unet_p_res.py:247) with the Dropout2d of
 * pool_drop (unet_p_res.py:248) fused as an optional per-(b,c) scale [B,C].                   */
int pu_maxpool2_fwd(const float* x, const float* chan_scale, float* y, int B, int H, int W, int C, void* stream);
/* dx gets dy*scale at the first maximum (row-major scan, ATen tie-break) and 0 elsewhere.  acc (may be NULL, shape of
 * x): a second gradient of x — the skip connection's (unet_p.py:165) — added in the same pass: dx = route(dy) + acc.   */
int pu_maxpool2_bwd(const float* x, const float* chan_scale, const float* dy, const float* acc, float* dx,
                    int B, int H, int W, int C, int flags, void* stream);
/* The same pair with an arg-max code: the forward pass also writes one byte per pooled element (bits 0-1: position of the
 * maximum in the 2x2 window with ATen's tie-break, bit 2: maximum > 0) and the backward pass routes from it without re-reading
 * x — a third less HBM traffic for the pooling backward of the training step (x [B,H,W,C] vs code [B,H/2,W/2,C] bytes). */
int pu_maxpool2_fwd_code(const float* x, const float* chan_scale, float* y, unsigned char* code, int B, int H, int W, int C,
                         void* stream);
int pu_maxpool2_bwd_code(const unsigned char* code, const float* chan_scale, const float* dy, const float* acc, float* dx,
                         int B, int H, int W, int C, int flags, void* stream);
/* nn.Upsample(scale_factor=2, bilinear, align_corners=True) (reference unet_p.py:153)        */
int pu_bilinear2x_fwd(const float* x, float* y, int B, int H, int W, int C, void* stream);
int pu_bilinear2x_bwd(const float* dy, float* dx, int B, int H, int W, int C, void* stream);

/* ---- concat / channel scale (only materialised in Dropout2d training mode) ------------------- */
int pu_concat_scale_fwd(const float* src0, int H0, int W0, int C0, int oy0, int ox0,
                        const float* src1, int H1, int W1, int C1, int oy1, int ox1,
                        const float* chan_scale, float* y, int B, int H, int W, int flags, void* stream);
/* inverse: dx0|dx1 windows receive dy[..., :C0]*scale | dy[..., C0:]*scale (either may be NULL; borders
 * outside the window are NOT written — caller zero-fills). */
int pu_concat_scale_bwd(const float* dy, const float* chan_scale,
                        float* dx0, int H0, int W0, int C0, int oy0, int ox0,
                        float* dx1, int H1, int W1, int C1, int oy1, int ox1, int B, int H, int W, void* stream);
/* y[b,p,c] = x[b,p,c] * scale[b,c]  (y may alias x) */
int pu_chan_scale(const float* x, const float* chan_scale, float* y, int B, long long hw, int C, void* stream);

/* ---- BatchNorm2d (optional, reference unet_p.py:106,109; unet_p_res.py:151,175) -------------- */
/* training forward: batch statistics; writes mean/invstd [C] (saved for backward) and updates the running
 * stats (may be NULL); y = (x-mean)*invstd*gamma+beta, optional fused ReLU.  ws: scratch of 2*C doubles. */
int pu_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* y,
                    float* save_mean, float* save_invstd, float* running_mean, float* running_var, double* ws,
                    float momentum, float eps, long long npix, int C, int flags, void* stream);
int pu_bn_eval_fwd(const float* x, const float* gamma, const float* beta, const float* running_mean,
                   const float* running_var, float* y, float eps, long long npix, int C, int flags, void* stream);
/* dy is grad wrt y (post-ReLU if relu); y needed iff relu. train!=0 uses the batch-statistics backward,
 * train==0 treats mean/invstd as constants (eval).  dgamma/dbeta may be NULL.  ws: 2*C doubles.          */
int pu_bn_bwd(const float* x, const float* y, const float* dy, const float* gamma,
              const float* mean, const float* invstd, float* dx, float* dgamma, float* dbeta, double* ws,
              long long npix, int C, int relu, int train, void* stream);
/* running_mean/var <- (1-momentum)*running + momentum*(batch mean / unbiased batch var recovered from invstd) */
int pu_bn_update_running(const float* mean, const float* invstd, float* running_mean, float* running_var,
                         float momentum, float eps, long long npix, int C, void* stream);
/* Eval-mode BatchNorm folded into the convolution in front of it (reference unet_p.py:105-110 in net.eval()):
 * w_out[co][...] = w[co][...] * s[co], b_out[co] = (b[co] - running_mean[co]) * s[co] + beta[co], s = gamma / sqrt(var + eps);
 * per_co = C_in * kh * kw weights per output channel; b may be NULL.  The conv then runs with its fused ReLU epilogue and
 * the normalisation costs no pass over the activations.                                                                */
int pu_bn_fold_conv(const float* w, const float* b, const float* gamma, const float* beta, const float* running_mean,
                    const float* running_var, float eps, float* w_out, float* b_out, int Cout, int per_co, void* stream);
/* invstd[c] = rsqrt(running_var[c] + eps) (eval-mode backward helper) */
int pu_bn_invstd(const float* running_var, float* invstd, float eps, int C, void* stream);

/* ---- plastic head (reference unet_p.py:70-79 == unet_p_res.py:116-125) ------------------------
 * X [B*N, N] (the B output maps viewed as N x N), Weff = w + alpha*hebb, A = X @ Weff, S = sigmoid(A).
 * weff_out [N,N] is scratch that backward re-uses.                                              */
int pu_plastic_head_fwd(const float* X, const float* w, const float* alpha, const float* hebb,
                        float* weff_out, float* S, int B, int N, void* stream);
/* gA = gS*S*(1-S); gX = gA @ Weff^T; gWeff = X^T @ gA; gw = gWeff; galpha = gWeff*hebb; ghebb = gWeff*alpha.
 * gA_ws is [B*N,N] scratch. galpha/ghebb may be NULL.  gw == NULL: only gA and gX; gS == NULL: gA_ws already
 * holds gA from such a call and only the parameter gradients are computed (lets a caller put them on a side stream). */
int pu_plastic_head_bwd(const float* X, const float* S, const float* gS, const float* weff,
                        const float* alpha, const float* hebb, float* gA_ws,
                        float* gX, float* gw, float* galpha, float* ghebb, int B, int N, void* stream);

/* Training-step form of the head (TF32 mode of TrainStep): forward + the reference's loss + their backward in ONE launch.
 * Replaces, for `loss = nn.BCELoss()(net(x, hebb)[0], target); loss.backward()` (train.py:99-104 over unet_p.py:70-79):
 *   S = sigmoid(X @ (w + alpha*hebb)); *loss = mean BCE(S, target) (logs clamped at -100 as torch does);
 *   gA = dloss/d(logits) [B*N, N]; gX = gA @ Weff^T (NULL: skipped).
 * mma.sync TF32: the logits with the three-term error-compensated split (fp32-level: they decide masks, loss and trace),
 * gX as a plain TF32 product (it feeds the TF32 data-gradient convs).  N <= 128.
 * scratch: NULL (then *loss is zeroed by a memset node and summed with atomics) or 1 + ceil(B*N/64) floats, zero before
 * the FIRST launch and left zeroed by every launch: deterministic block-ordered loss sum, no memset.
 * weff: NULL (Weff = w + alpha*hebb is built inside the kernel) or the matrix pu_head_weff computed beforehand — off the
 * critical path, e.g. at the start of the step: the parameters and the trace are known then (w, alpha, hebb may be NULL). */
int pu_plastic_head_bce(const float* X, const float* w, const float* alpha, const float* hebb, const float* weff,
                        const float* target, float* S, float* loss, float* gA, float* gX, float* scratch, int B, int N,
                        void* stream);
/* weff = w + alpha*hebb [N, N] (unet_p.py:73-76; 'free' and 'yoked' alike) */
int pu_head_weff(const float* w, const float* alpha, const float* hebb, float* weff, int N, void* stream);
/* The head's parameter gradients from gA (either form): gw = X^T @ gA (split-K mma.sync, fp32 atomics; terms = 3:
 * error-compensated 3xTF32, terms = 1: plain TF32 as the conv weight gradients of the TF32 mode),
 * galpha = gw*hebb, ghebb = gw*alpha (each may be NULL).                                                         */
int pu_plastic_head_wgrad_tc(const float* X, const float* gA, const float* alpha, const float* hebb,
                             float* gw, float* galpha, float* ghebb, int B, int N, int terms, void* stream);

/* ---- plastic trace update (reference unet_p.py:81-84 == unet_p_res.py:127-130) ----------------
 * pre/post are the stacked pre-/post-synaptic rows [K, N] (parity mode: K = B, row 0 of every
 * map, given as pointers to map b's first row with row stride `ld` floats).
 * One fused pass:  delta = pre^T @ post (contraction over K), q = sum_k post^2,
 *   hebb: out = (1-eta)*hebb + eta*delta/K
 *   oja : out = hebb*(1 - eta*q/K) + eta*delta/K
 * which is exactly the reference at K == 1.  eta is a device scalar.                            */
int pu_trace_update_fwd(const float* hebb, const float* pre, const float* post, long long ld, int K,
                        const float* eta, int rule, float* out, int N, void* stream);
/* data-parallel split form: delta_q = [N*N + N] floats = (sum_k outer, sum_k post^2) — all-reduced
 * over ranks by the caller — then the epilogue with K_global.                                   */
int pu_trace_delta(const float* pre, const float* post, long long ld, int K, float* delta_q, int N, void* stream);
/* The same payload for LARGE K on the tensor cores (opt-in rows='all' mode: K = B*N, every row of every map is a
 * (pre, post) pair — what the reference's bmm computes before it keeps [0], unet_p.py:82): split-K mma.sync TF32 GEMM
 * with the 3xTF32 error-compensated split (fp32-level accuracy), partial sums merged with fp32 atomics.            */
int pu_trace_delta_tc(const float* pre, const float* post, long long ld, int K, float* delta_q, int N, void* stream);
int pu_trace_apply(const float* hebb, const float* delta_q, int K_global, const float* eta, int rule,
                   float* out, int N, void* stream);
/* backward of the fused update w.r.t. hebb, pre, post, eta (closed forms) */
int pu_trace_update_bwd(const float* hebb, const float* pre, const float* post, long long ld, int K,
                        const float* eta, int rule, const float* gout,
                        float* ghebb, float* gpre, float* gpost, float* geta, int N, void* stream);

/* ---- inference tail: integer / byte work, bit-exact ------------------------------- */
/* Threshold sweep: confusion counts of (label class, pred > thr[j]) for every image b and EVERY threshold j in one pass
 * over the predictions (reference eval.py:48-52 sweeps 31 thresholds, re-reading the predictions 31 times;
 * utils/iou_metric.py:34-36 bins labels and predictions with np.histogram(bins=[0,0.5,1]); utils/iou_metric.py:10,23 use
 * `A > 0` / `pred > 0.5`).  pred, label: [B, npix] fp32.  thr_sorted: T <= 64 thresholds, ASCENDING, device doubles
 * (the comparison is (double)pred > thr, as numpy does against a float64 threshold array).
 * label_mode 0: histogram bins [0,0.5) -> 0, [0.5,1] -> 1, anything else dropped; 1: label > 0.
 * hist_ws: [B*3*(T+1)] int32 workspace.  counts: [B][T][6] int32 = {c00, c01, c10, c11, pred_ones, npix}
 * (c<true><pred> over pixels with a valid label; pred_ones over all pixels).                                */
int pu_threshold_counts(const float* pred, const float* label, const double* thr_sorted, int T, int label_mode, int B, long long npix,
                        int* hist_ws, int* counts, void* stream);
/* Mask threshold + run-length encoding (reference infer.py:81,88,99 `mask > mask_threshold`; utils/rle_encode.py:6-17:
 * column-major flattening, 1-based run starts, "start length" pairs).  pred: [B, R, C] fp32 row-major.  mask (nullable):
 * [B, R, C] uint8 = pred > thr.  runs: [B][cap] int32, image b gets count[b] ints = (start, length) pairs in order;
 * count[b] = -(needed ints) if cap is too small (cap = R*C + 2 always suffices).  One CTA per image; R*C <= ~1.8 M.    */
int pu_mask_rle(const float* pred, double thr, int B, int R, int C, unsigned char* mask, int* runs, int cap, int* count, void* stream);
long long pu_mask_rle_smem_bytes(int R, int C);

/* ---- train-step tail (reference train.py:66-70,101-112) ------------------ */
/* loss = mean BCE(S, T) with log clamped at -100 (nn.BCELoss); gS = dloss/dS. loss is a device scalar. */
int pu_bce_fwd_bwd(const float* S, const float* T, float* loss, float* gS, long long n, void* stream);
/* Adam over one flat arena (torch.optim.Adam defaults; step_count is a device float scalar that is incremented) */
int pu_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step_count,
                 const float* lr, float beta1, float beta2, float eps, float grad_scale, long long n, void* stream);

/* table: n rows of (source device pointer, destination offset in floats, element count) as int64 on the device;
 * copies every source tensor into flat[offset ...] with one launch (gradient tensors -> flat gradient arena). */
int pu_gather_flat(const long long* table, int n, float* flat, void* stream);
/* dst0[0..n0) = src0, dst1[0..n1) = src1 (device memory, 16-byte aligned, counts multiples of 4): the step's image and mask
 * batches (train.py:94-95) into its static buffers with one launch instead of two copies.                              */
int pu_copy2(const float* src0, float* dst0, long long n0, const float* src1, float* dst1, long long n1, void* stream);
/* The same launch also increments *step_count, and pu_adam_step_counted is pu_adam_step for a counter that already holds the
 * number of THIS step: the optimizer step costs two launches instead of three.                                         */
int pu_gather_flat_inc(const long long* table, int n, float* flat, float* step_count, void* stream);
int pu_adam_step_counted(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* step_count,
                         const float* lr, float beta1, float beta2, float eps, float grad_scale, long long n, void* stream);
/* Adam straight from the per-parameter gradient tensors: table rows as for pu_gather_flat, (pointer, offset into the flat
 * parameter / moment arenas, element count).  One launch replaces pu_gather_flat + pu_adam_step on a single GPU (the
 * arithmetic per element is pu_adam_step's; *step_count is incremented by the launch's last block).  Arena elements outside
 * every row are left alone (zero gradient, zero moments).  Launches on one device must be stream-ordered.          */
int pu_adam_table_step(const long long* table, int n, float* param, float* exp_avg, float* exp_avg_sq, float* step_count,
                       const float* lr, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* Data parallel: gradient exchange FUSED with the optimizer over NVLink peer memory.  peer_grad_ptrs / peer_flag_ptrs: device
 * arrays of `world` pointers (as int64) to every rank's flat gradient arena (n floats) and flag buffer
 * (2 * pu_adam_allreduce_blocks() * world int32, zero-initialised), all in symmetric (peer-mapped) memory.  One launch per rank and
 * step: barrier with the peers, g = sum_j peer_grad[j] read over NVLink (same order on every rank: bit-identical replicas),
 * Adam update of the local arena with grad_scale * g, barrier.  Replaces ncclAllReduce + pu_adam_step (train.py:110-111 under
 * data parallelism).  world in {2, 4, 8}; every rank must call it once per step.  n_extra (multiple of 4, may be 0): that many
 * floats FOLLOW the n gradients in every rank's arena and are only summed over the ranks into extra_sum (local memory) — the
 * plastic-trace delta (pu_trace_delta) of the step, so that its all-reduce shares the round trip; apply it with pu_trace_apply. */
int pu_adam_allreduce_step(float* param, const long long* peer_grad_ptrs, const long long* peer_flag_ptrs, int rank, int world, float* exp_avg,
                           float* exp_avg_sq, float* step_count, const float* lr, float beta1, float beta2, float eps, float grad_scale,
                           long long n, float* extra_sum, long long n_extra, void* stream);
int pu_adam_allreduce_blocks(void);

/* Input pipeline: batch assembly from a DEVICE-resident dataset src [n, planes, Hs, Ws] + zero padding
 * to Hd x Wd at offset (oy, ox) in one pass: dst[b] = pad(src[idx[b]]).  idx: B int64 sample indices on the device.
 * Replaces the per-step host conversion + H2D copy of reference train.py:94-95 (and defines the 101 -> 128 padding of
 * BASELINE configs: oy = ox = 13).                                                                              */
int pu_gather_pad(const float* src, const long long* idx, float* dst, int B, int planes, int Hs, int Ws, int Hd, int Wd, int oy, int ox,
                  void* stream);

/* elementwise helpers for autograd glue */
int pu_add(const float* a, const float* b, float* out, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLASTIC_UNET_B200_H_ */
