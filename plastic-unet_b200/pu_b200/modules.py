"""Plastic U-Net modules — the reference's nn.Module surface over B200 custom ops.

Drop-in for reference ``src/unet`` (``from unet import UNetp, UNetpRes``): identical constructor
signatures, ``forward(x, hebb) -> (activout, hebb')``, ``initialZeroHebb()``, attribute names and
``state_dict`` keys/shapes, so ``train.py`` / ``eval.py`` / ``infer.py`` run unchanged
and ``.pth`` checkpoints interchange.  The sub-module tree (``inc.conv.conv.0`` ...) is kept as the
*parameter container* — the stock ``nn.Conv2d`` / ``nn.ConvTranspose2d`` / ``nn.BatchNorm2d`` objects
are never called; forward runs only ``pu_b200.ops`` (hand-written sm_100a kernels, NHWC fp32).

Extensions beyond the reference (all default-off / default-identical):
  * ``batched=True``  — accept B > 1: every map uses the shared trace ``hebb``; the trace update is the
    mean over the batch of the per-sample reference updates (exactly the reference at B == 1).
  * ``depth=``        — encoder depth (4 = reference), ``base=`` first-level width (8 = reference).
  * ``UNetpCoord``    — CoordConv stem + plastic head hybrid (coord_conv_script.py topology).
  * ``dp_group``      — data-parallel trace: all-reduce of the trace delta (see dp.py).
"""
import os

import torch
import torch.nn as nn

from . import ops

_MATH = {"fp32": ops.MATH_FP32, "tf32": ops.MATH_TF32, "mixed": ops.MATH_MIXED}
_TC_BWD = (ops.MATH_TF32, ops.MATH_MIXED)  # modes whose backward runs on the tensor cores (premasked-gradient protocol)


def _default_math() -> str:
    m = os.environ.get("PU_CONV_MATH", "fp32").lower()
    if m not in _MATH:
        raise ValueError("PU_CONV_MATH must be 'fp32', 'tf32' or 'mixed'")
    return m


# --------------------------------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------------------------------
def _c3(x0, conv, relu, x1=None, res=None, H=None, W=None, off0=(0, 0), off1=(0, 0), math=ops.MATH_FP32,
        m0=None, m1=None, premasked=False, emit_mask=False):
    """conv3x3 over cat[x0, x1] windows with fused bias / residual / ReLU using the parameters held by `conv`.
    emit_mask: -> (y, packed ReLU mask of y) instead of y."""
    if H is None:
        H, W = x0.shape[1], x0.shape[2]
    y, ym = ops.conv3x3_m(x0, x1, conv.weight, conv.bias, res, relu, H, W, off0[0], off0[1], off1[0], off1[1], math,
                          m0, m1, premasked, emit_mask)
    return (y, ym) if emit_mask else y


class Masked:
    """A ReLU output together with its packed mask (premasked-gradient protocol)."""
    __slots__ = ("t", "m")

    def __init__(self, t, m=None):
        self.t, self.m = t, m


def _bn(x, bn, relu, math=ops.MATH_FP32):
    """BatchNorm2d over NHWC with the parameters/buffers held by `bn` (+ optional fused ReLU).  Follows nn.BatchNorm2d:
    batch statistics when training or when no running statistics are kept; momentum=None = cumulative moving average."""
    track = bn.track_running_stats and bn.running_mean is not None
    train = bn.training or not track
    momentum = 0.1 if bn.momentum is None else bn.momentum
    if train and track and bn.training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if bn.momentum is None:  # exponential_average_factor = 1 / num_batches_tracked
            momentum = 1.0 / float(bn.num_batches_tracked)
    C = x.shape[-1]
    # the op schema takes tensors: placeholders when the module keeps no running statistics (never read in that case)
    rm = bn.running_mean if track else x.new_zeros(C)
    rv = bn.running_var if track else x.new_ones(C)
    y, mean, invstd = ops.batchnorm(x, bn.weight, bn.bias, rm, rv, train, momentum, bn.eps, relu, math == ops.MATH_TF32)
    if train and track and bn.training:
        ops.bn_update_running(mean.detach(), invstd.detach(), bn.running_mean, bn.running_var, momentum, bn.eps,
                              x.numel() // x.shape[-1])
    return y


# Test hook: a list of [B, C] Dropout2d scale tensors consumed front to back instead of drawing from the
# device RNG (lets the parity tests replay the noise the CPU reference drew).  None = draw normally.
DROPOUT_MASK_QUEUE = None


def _feature_noise(x_like_B, C, p, device):
    """Dropout2d noise exactly as ATen's feature_dropout draws it: [B,C,1,1].bernoulli_(1-p)/(1-p)  -> [B, C]."""
    if DROPOUT_MASK_QUEUE is not None:
        m = DROPOUT_MASK_QUEUE.pop(0).to(device=device, dtype=torch.float32).contiguous()
        if tuple(m.shape) != (x_like_B, C):
            raise RuntimeError("injected dropout mask has shape %s, expected %s" % (tuple(m.shape), (x_like_B, C)))
        return m
    noise = torch.empty((x_like_B, C, 1, 1), device=device, dtype=torch.float32).bernoulli_(1 - p).div_(1 - p)
    return noise.view(x_like_B, C)


class _PlasticBase(nn.Module):
    """Plastic head + trace shared by every variant (reference unet_p.py:69-94 == unet_p_res.py:115-140)."""

    def _init_plastic(self, n_channels, n_classes, device, alfa_type, rule, nbf, batched):
        self.n_classes = n_classes
        self.n_channels = n_channels
        self.nbf = nbf
        self.torch_dev = device
        self.alfa_type = alfa_type
        self.rule = rule
        self.batched = batched
        self.conv_math = _default_math()
        # 'row0' (default) = the reference: only row 0 of each map enters the trace (unet_p.py:82 keeps [0] of the bmm).
        # 'all' (opt-in extension, no reference semantics): EVERY row of every map is a (pre, post) pair, K = B*nbf — the
        # full contraction the reference's bmm computes before discarding all but one row — on the tensor cores; the trace
        # is detached in this mode (as train.py:99 does every step anyway).
        self.trace_rows = 'row0'
        self.premask = True   # TF32 mode: fold each ReLU mask into its consumers' backward epilogues (UNetp / UNetpCoord)
        self.dp_group = None  # set by pu_b200.dp.attach() for the data-parallel trace all-reduce
        self.dp_defer = False  # TrainStep: the trace update (and its DP all-reduce) runs on a side stream, off the critical path
        self.dp_side = None
        self.dp_external = None  # TrainStep (fused peer-memory exchange): callback(delta_q, K_global, rule) that takes over the trace reduction
        self.dp_world = 1
        self._head_bce = None   # TrainStep (TF32 mode): the target tensor -> _plastic fuses head + BCE loss + their backward
        self._head_loss = None  # ... and leaves the loss (autograd root of the step) here
        self._head_weff = None  # TrainStep: (w + alpha*hebb, event) computed off the critical path at the start of the step
        # TrainStep (data parallel): a callback fired in the backward pass once the gradient w.r.t. the input of encoder level
        # `_bucket_level` exists, i.e. when every parameter gradient of the decoder and of the deeper encoder levels has been
        # launched — the trainer all-reduces that bucket on a communication stream while the shallow levels still run
        self._bucket_hook = None
        self._bucket_level = 0
        # same creation order and RNG consumption as the reference (unet_p.py:30-32)
        self.w = torch.nn.Parameter((.01 * torch.randn(self.nbf, self.nbf, device=self.torch_dev)), requires_grad=True)
        self.alpha = torch.nn.Parameter((.01 * torch.rand(self.nbf, self.nbf, device=self.torch_dev)), requires_grad=True)
        self.eta = torch.nn.Parameter((.01 * torch.ones(1, device=self.torch_dev)), requires_grad=True)

    @property
    def _math(self):
        return _MATH[self.conv_math]

    def _to_nhwc(self, x):
        if x.dim() != 4:
            raise ValueError("expected a [B, C, H, W] input")
        x = x.contiguous()
        B, C, H, W = x.shape
        if C == 1:
            return x.view(B, H, W, 1)
        return ops.nchw_to_nhwc(x)

    def _plastic(self, o, hebb):
        """o: [B, H, W, 1] output map of the 1x1 conv  ->  (activout, hebb')."""
        N = self.nbf
        B = o.shape[0]
        if self.alfa_type not in ('free', 'yoked'):
            raise ValueError("Must select one plasticity coefficient type ('free' or 'yoked')")
        if o.numel() != B * N * N:
            raise RuntimeError("shape '[%d, %d]' is invalid for input of size %d" % (N, N, o.numel() // B))
        X = o.view(B * N, N)
        hebb = hebb.contiguous()
        # 'free' and 'yoked' are numerically identical (alpha is always [nbf, nbf])
        self._head_loss = None
        tgt = getattr(self, "_head_bce", None)
        if tgt is not None and X.is_cuda and N <= 128 and self.conv_math == 'tf32' and torch.is_grad_enabled():
            # TrainStep, TF32 mode: head forward + BCE loss + the backward of both in ONE launch (3xTF32 tensor-core GEMMs);
            # the loss leaves through self._head_loss, and its backward() enters the U-Net with the finished gX
            wf = getattr(self, "_head_weff", None)  # (Weff, event): computed by TrainStep at the start of the step
            if wf is not None:
                torch.cuda.current_stream().wait_event(wf[1])
            S, self._head_loss, _, _ = ops.plastic_head_bce(X, self.w, self.alpha, hebb, tgt, X.requires_grad,
                                                            wf[0] if wf is not None else None)
        else:
            S, _ = ops.plastic_head(X, self.w, self.alpha, hebb)
        if self.rule == 'hebb':
            rule = ops.RULE_HEBB
        elif self.rule == 'oja':
            rule = ops.RULE_OJA
        else:
            raise ValueError("Must select one learning rule ('hebb' or 'oja')")
        if self.trace_rows not in ('row0', 'all'):
            raise ValueError("trace_rows must be 'row0' (reference) or 'all'")
        all_rows = self.trace_rows == 'all'
        if all_rows and not (self.dp_group is not None and self.dp_world > 1):
            delta_q = ops.trace_delta_tc(X.detach(), S.detach(), N, N, B * N)
            hebb_new = ops.trace_apply(hebb.detach(), delta_q, self.eta.detach(), rule, B * N)
            return (S.view(N, N) if B == 1 else S.view(B, N, N)), hebb_new
        if self.dp_group is not None and self.dp_world > 1:
            import torch.distributed as dist
            # data-parallel: all-reduce (sum_k outer, sum_k post^2), then the identical epilogue on every rank
            kloc = B * N if all_rows else B  # (pre, post) pairs of this rank

            def local_delta():
                return ops.trace_delta_tc(X.detach(), S.detach(), N, N, B * N) if all_rows else ops.trace_delta(X.detach(), S.detach(), N, N * N, B)

            if getattr(self, "dp_external", None) is not None and getattr(self, "dp_defer", False) and S.is_cuda:
                # TrainStep with the peer-memory exchange: the trainer sums the delta over the ranks inside its fused
                # gradient-exchange kernel at the end of the step and applies it there; the trace returned here is a placeholder.
                # The local delta (and its copy into the exchange arena) run on the side stream, off the critical path.
                if getattr(self, "dp_side", None) is None:
                    self.dp_side = torch.cuda.Stream()
                main = torch.cuda.current_stream()
                self.dp_side.wait_stream(main)
                with torch.cuda.stream(self.dp_side):
                    delta_q = local_delta()
                    self.dp_external(delta_q, kloc * self.dp_world, rule)
                X.record_stream(self.dp_side)
                S.record_stream(self.dp_side)
                return (S.view(N, N) if B == 1 else S.view(B, N, N)), hebb.detach()
            delta_q = local_delta()
            if getattr(self, "dp_defer", False) and delta_q.is_cuda:
                # The new trace is only needed at the end of the step: run its all-reduce + epilogue on a side stream so
                # that they overlap the backward pass.  The caller (TrainStep) joins `self.dp_side` before it reads hebb_new.
                if getattr(self, "dp_side", None) is None:
                    self.dp_side = torch.cuda.Stream()
                main = torch.cuda.current_stream()
                self.dp_side.wait_stream(main)
                with torch.cuda.stream(self.dp_side):
                    dist.all_reduce(delta_q, op=dist.ReduceOp.SUM, group=self.dp_group)
                    hebb_new = ops.trace_apply(hebb.detach(), delta_q, self.eta.detach(), rule, kloc * self.dp_world)
                delta_q.record_stream(self.dp_side)
                hebb_new.record_stream(main)
            else:
                dist.all_reduce(delta_q, op=dist.ReduceOp.SUM, group=self.dp_group)
                hebb_new = ops.trace_apply(hebb.detach(), delta_q, self.eta.detach(), rule, kloc * self.dp_world)
        elif getattr(self, "dp_defer", False) and S.is_cuda:
            # TrainStep (single GPU): the loss does not depend on the new trace and the step detaches it (train.py:99), so the
            # update leaves the critical path: side stream, joined by TrainStep before the optimizer (it reads eta)
            if getattr(self, "dp_side", None) is None:
                self.dp_side = torch.cuda.Stream()
            main = torch.cuda.current_stream()
            self.dp_side.wait_stream(main)
            with torch.cuda.stream(self.dp_side):
                hebb_new = ops.trace_update(hebb.detach(), X.detach(), S.detach(), self.eta.detach(), rule, N * N, B)
            X.record_stream(self.dp_side)
            S.record_stream(self.dp_side)
            hebb_new.record_stream(main)
        else:
            # rows k of pre/post = row 0 of map k (reference keeps only [0] of the bmm)
            hebb_new = ops.trace_update(hebb, X, S, self.eta, rule, N * N, B)
        activout = S.view(N, N) if B == 1 else S.view(B, N, N)
        return activout, hebb_new

    def initialZeroHebb(self):
        """Creates variable to store Hebbian plasticity coefficients (reference unet_p.py:90-94)."""
        return torch.zeros(self.nbf, self.nbf, dtype=torch.float, device=self.torch_dev)


# --------------------------------------------------------------------------------------------------
# UNetp  (reference src/unet/unet_p.py)
# --------------------------------------------------------------------------------------------------
class double_conv(nn.Module):
    """(Conv3x3 => [BN] => ReLU) * 2  — parameter container of reference unet_p.py:96-122."""

    def __init__(self, in_ch, out_ch, batch_norm):
        super(double_conv, self).__init__()
        self.batch_norm = batch_norm
        if batch_norm:
            self.conv = nn.Sequential(
                nn.Conv2d(in_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True),
                nn.Conv2d(out_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True))
        else:
            self.conv = nn.Sequential(
                nn.Conv2d(in_ch, out_ch, 3, padding=1), nn.ReLU(inplace=True),
                nn.Conv2d(out_ch, out_ch, 3, padding=1), nn.ReLU(inplace=True))

    def run(self, x0, math, x1=None, H=None, W=None, off0=(0, 0), off1=(0, 0), m0=None, m1=None, premask=False):
        """premask: premasked-gradient protocol (TF32 mode, no BN): the ReLU mask of each conv is applied by the backward
        epilogue of its consumers instead of a separate pass; every conv emits the packed mask of its output from its
        forward epilogue (1 bit per element) and m0/m1 are the packed masks of sources that are such outputs.
        Returns the output tensor, or Masked(output, packed mask) when premask."""
        if self.batch_norm:
            bn1, bn2 = self.conv[1], self.conv[4]
            if not (bn1.training or bn2.training) and bn1.running_mean is not None and bn2.running_mean is not None \
                    and not torch.is_grad_enabled():
                # eval mode (eval.py:79-80, infer.py:38-39 run under no_grad): BN folds into the conv weights and bias, the
                # ReLU into the conv epilogue — conv + BN + ReLU is ONE kernel and no extra pass over the activations
                Ho, Wo = (x0.shape[1], x0.shape[2]) if H is None else (H, W)
                w1, b1 = ops.bn_fold_conv(self.conv[0].weight, self.conv[0].bias, bn1.weight, bn1.bias, bn1.running_mean,
                                          bn1.running_var, bn1.eps)
                y = ops.conv3x3(x0, x1, w1, b1, None, True, Ho, Wo, off0[0], off0[1], off1[0], off1[1], math)
                w2, b2 = ops.bn_fold_conv(self.conv[3].weight, self.conv[3].bias, bn2.weight, bn2.bias, bn2.running_mean,
                                          bn2.running_var, bn2.eps)
                return ops.conv3x3(y, None, w2, b2, None, True, Ho, Wo, 0, 0, 0, 0, math)
            y = _c3(x0, self.conv[0], False, x1=x1, H=H, W=W, off0=off0, off1=off1, math=math)
            y = _bn(y, self.conv[1], True, math)
            y = _c3(y, self.conv[3], False, math=math)
            return _bn(y, self.conv[4], True, math)
        if not premask:
            y = _c3(x0, self.conv[0], True, x1=x1, H=H, W=W, off0=off0, off1=off1, math=math)
            return _c3(y, self.conv[2], True, math=math)
        y, ym = _c3(x0, self.conv[0], True, x1=x1, H=H, W=W, off0=off0, off1=off1, math=math, m0=m0, m1=m1, premasked=True,
                    emit_mask=True)
        z, zm = _c3(y, self.conv[2], True, math=math, m0=ym, premasked=True, emit_mask=True)
        return Masked(z, zm)


class inconv(nn.Module):
    def __init__(self, in_ch, out_ch, batch_norm=True):
        super(inconv, self).__init__()
        self.conv = double_conv(in_ch, out_ch, batch_norm)

    def run(self, x, math, premask=False):
        return self.conv.run(x, math, premask=premask)


class down(nn.Module):
    def __init__(self, in_ch, out_ch, batch_norm=True):
        super(down, self).__init__()
        self.mpconv = nn.Sequential(nn.MaxPool2d(2), double_conv(in_ch, out_ch, batch_norm))

    def run(self, x, math, premask=False):
        """x: tensor, or Masked when premask (the pool's backward then applies the ReLU mask of x itself: it reads x anyway)."""
        xt = x.t if premask else x
        return self.mpconv[1].run(ops.maxpool2(xt, None, premask), math, premask=premask)

    def run_pooled(self, pooled, math, premask=False):
        """The double conv on an already pooled tensor (UNetp.forward pools through ops.pool_skip)."""
        return self.mpconv[1].run(pooled, math, premask=premask)


class up(nn.Module):
    """ConvTranspose2d(k2,s2) (or bilinear x2) on the deep tensor, crop the skip, cat [skip, up], double_conv
    (reference unet_p.py:148-167)."""

    def __init__(self, in_ch, out_ch, bilinear=True, batch_norm=True):
        super(up, self).__init__()
        self.bilinear = bilinear
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        else:
            self.up = nn.ConvTranspose2d(in_ch // 2, in_ch // 2, 2, stride=2)
        self.conv = double_conv(in_ch, out_ch, batch_norm)

    def run(self, x1, x2, math, premask=False):
        m2 = None
        if premask:  # Masked inputs: x1 (deep tensor; the transposed conv's dgrad masks with x1 itself), x2 (skip) with its packed mask
            x1, x2, m2 = x1.t, x2.t, x2.m
        u = ops.bilinear2x(x1) if self.bilinear else ops.convT2x2s2(x1, self.up.weight, self.up.bias, math == ops.MATH_TF32, premask,
                                                                    math == ops.MATH_MIXED)
        diffX = u.shape[1] - x2.shape[1]  # reference names: size()[2] == H
        diffY = u.shape[2] - x2.shape[2]
        # F.pad(x2, (diffX//2, int(diffX/2), diffY//2, int(diffY/2))): first pair pads W, second pair pads H
        if diffX > 0 or diffY > 0:
            raise RuntimeError("up: upsampled tensor larger than the skip connection is not supported")
        w_new = x2.shape[2] + diffX // 2 + int(diffX / 2)
        h_new = x2.shape[1] + diffY // 2 + int(diffY / 2)
        if h_new != u.shape[1] or w_new != u.shape[2]:
            raise RuntimeError("Sizes of tensors must match except in dimension 1")
        off_skip = (-(diffY // 2), -(diffX // 2))  # (oy, ox) of the crop window in the skip tensor
        return self.conv.run(x2, math, x1=u, H=u.shape[1], W=u.shape[2], off0=off_skip, off1=(0, 0), m0=m2, premask=premask)


class outconv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super(outconv, self).__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, 1)

    def run(self, x, premask=False):
        w = self.conv.weight
        return ops.conv1x1(x.t if premask else x, w.view(w.shape[0], w.shape[1]), self.conv.bias, 0, False, False, premask)


class UNetp(_PlasticBase):
    def __init__(self, n_channels, n_classes, device, alfa_type='free', rule='hebb', nbf=128, batch_norm=False,
                 bilinear_upsample=False, batched=False, depth=4, base=8):
        """
        Plastic U-Net (reference unet_p.py:8-52).  Arguments as in the reference:
            n_channels, n_classes, device, alfa_type ['free','yoked'], rule ['hebb','oja'], nbf, batch_norm,
            bilinear_upsample.  Extensions: batched (accept B>1), depth (encoder levels), base (first width).
        """
        super(UNetp, self).__init__()
        self._init_plastic(n_channels, n_classes, device, alfa_type, rule, nbf, batched)
        self.depth = depth
        c = [base * (2 ** i) for i in range(depth)]
        c.append(c[-1])  # the deepest level keeps its width (reference: down4 = 64 -> 64)
        self.inc = inconv(n_channels, c[0], batch_norm=batch_norm)
        for k in range(1, depth + 1):
            setattr(self, "down%d" % k, down(c[k - 1], c[k], batch_norm=batch_norm))
        for j in range(1, depth + 1):
            skip_c = c[depth - j]
            out_c = c[depth - j - 1] if depth - j - 1 >= 0 else c[0]
            setattr(self, "up%d" % j, up(2 * skip_c, out_c, batch_norm=batch_norm, bilinear=bilinear_upsample))
        self.outc = outconv(c[0], n_classes)
        self.to(device)
        print("UNet plastic model with plastic rule [%s] initialized" % self.rule)

    def forward(self, x, hebb):
        if x.shape[0] != 1 and not self.batched:
            raise ValueError("Only batch size: 1 is supported, but was: %d" % x.shape[0])
        m = self._math
        x = self._to_nhwc(x)
        # premasked-gradient protocol (backward only, TF32 mode without BN / bilinear / ragged output conv): see DESIGN.md 4.2
        pm = (m in _TC_BWD and self.premask and not self.inc.conv.batch_norm and not self.up1.bilinear
              and self.outc.conv.weight.shape[1] in (8, 16, 32, 64) and self.n_classes == 1 and self.n_channels == 1)
        # each encoder level output feeds the pool of the next level AND a skip connection: ops.pool_skip returns both so
        # that the two gradients are summed inside the pooling backward kernel (no separate accumulation pass)
        feats = []
        f = self.inc.run(x, m, pm)
        for k in range(1, self.depth + 1):
            ft, fm = (f.t, f.m) if pm else (f, None)
            pooled, skip = ops.pool_skip(ft, pm)
            if self._bucket_hook is not None and k == self._bucket_level and pooled.requires_grad:
                pooled.register_hook(self._bucket_hook)
            feats.append(Masked(skip, fm) if pm else skip)
            f = getattr(self, "down%d" % k).run_pooled(pooled, m, pm)
        feats.append(f)
        y = feats[-1]
        for j in range(1, self.depth + 1):
            y = getattr(self, "up%d" % j).run(y, feats[self.depth - j], m, pm)
        o = self.outc.run(y, pm)
        return self._plastic(o, hebb)


# --------------------------------------------------------------------------------------------------
# UNetpRes  (reference src/unet/unet_p_res.py)
# --------------------------------------------------------------------------------------------------
class conv_module(nn.Module):
    """Conv2d(C, C, k) [+BN] [+ReLU] — parameter container of reference unet_p_res.py:142-164."""

    def __init__(self, out_ch, kernel_size, stride=1, padding=1, activation=True, batch_norm=False):
        super(conv_module, self).__init__()
        self.batch_norm = batch_norm
        if batch_norm == True:  # noqa: E712 (reference spelling)
            self.conv = nn.Sequential(
                nn.Conv2d(out_ch, out_ch, kernel_size=kernel_size, stride=stride, padding=padding), nn.BatchNorm2d(out_ch))
        else:
            self.conv = nn.Conv2d(out_ch, out_ch, kernel_size=kernel_size, stride=stride, padding=padding)
        self.activation = activation
        if activation == True:  # noqa: E712
            self.activ = nn.ReLU(inplace=True)

    def run(self, x, math, res=None, relu_after_res=False):
        if self.batch_norm:
            y = _c3(x, self.conv[0], False, math=math)
            y = _bn(y, self.conv[1], self.activation, math)
            if res is not None:
                raise RuntimeError("conv_module with BN cannot fuse a residual")
            return y
        return _c3(x, self.conv, self.activation or relu_after_res, res=res, math=math)


class residual_block(nn.Module):
    """r = relu_(x); y = conv(relu(conv([BN] r))) + r   (reference unet_p_res.py:166-189; the in-place leading
    ReLU means the skip adds relu(input)).  The leading ReLU is fused into the producer
    of `x`, and the ReLU that follows this block (next block's leading ReLU or the trailing nn.ReLU of
    down/middle) is fused into this block's second conv epilogue."""

    def __init__(self, out_ch, batch_norm=False):
        super(residual_block, self).__init__()
        self.batch_norm = batch_norm
        if batch_norm == True:  # noqa: E712
            self.conv = nn.Sequential(nn.ReLU(inplace=True), nn.BatchNorm2d(out_ch),
                                      conv_module(out_ch, kernel_size=3), conv_module(out_ch, kernel_size=3, activation=False))
        else:
            self.conv = nn.Sequential(nn.ReLU(inplace=True),
                                      conv_module(out_ch, kernel_size=3), conv_module(out_ch, kernel_size=3, activation=False))

    def run(self, r, math):
        """r is already relu(input).  Returns relu(conv2(relu(conv1([BN] r))) + r)."""
        if self.batch_norm:
            a = self.conv[2].run(_bn(r, self.conv[1], False, math), math)
            return self.conv[3].run(a, math, res=r, relu_after_res=True)
        a = self.conv[1].run(r, math)
        return self.conv[2].run(a, math, res=r, relu_after_res=True)


class _res_stack(nn.Module):
    """Conv3x3(in->out) -> residual_block x2 -> ReLU (both `down` and `middle` of the reference)."""

    def _run_stack(self, seq, x0, math, x1=None, H=None, W=None, off0=(0, 0), off1=(0, 0)):
        r = _c3(x0, seq[0], True, x1=x1, H=H, W=W, off0=off0, off1=off1, math=math)  # conv + leading ReLU of block 1
        r = seq[1].run(r, math)  # ... + leading ReLU of block 2
        return seq[2].run(r, math)  # ... + trailing nn.ReLU


class res_down(_res_stack):
    def __init__(self, in_ch, out_ch, batch_norm=False):
        super(res_down, self).__init__()
        self.dconv = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=3, padding=1),
                                   residual_block(out_ch=out_ch, batch_norm=batch_norm),
                                   residual_block(out_ch=out_ch, batch_norm=batch_norm), nn.ReLU(inplace=True))

    def run(self, x, math):
        return self._run_stack(self.dconv, x, math)


class middle(_res_stack):
    def __init__(self, in_ch, out_ch, batch_norm=False):
        super(middle, self).__init__()
        self.mconv = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=3, padding=1),
                                   residual_block(out_ch=out_ch, batch_norm=batch_norm),
                                   residual_block(out_ch=out_ch, batch_norm=batch_norm), nn.ReLU(inplace=True))

    def run(self, x0, math, **kw):
        return self._run_stack(self.mconv, x0, math, **kw)


class pool_drop(nn.Module):
    """MaxPool2d(2) + Dropout2d(p) fused (reference unet_p_res.py:240-253)."""

    def __init__(self, dropout_ratio):
        super(pool_drop, self).__init__()
        self.dpool = nn.Sequential(nn.MaxPool2d(2), nn.Dropout2d(p=dropout_ratio, inplace=True))

    def run(self, x):
        p = self.dpool[1].p
        scale = None
        if self.training and p > 0:
            scale = _feature_noise(x.shape[0], x.shape[3], p, x.device)
        return ops.maxpool2(x, scale)


class res_up(nn.Module):
    """ConvTranspose2d(k3,s2,p0) -> crop to the skip size -> cat [up, skip] -> Dropout2d -> middle
    (reference unet_p_res.py:200-220).  Crop is fused into the transposed conv; cat is fused into the
    first conv3x3 of `middle` unless Dropout2d is active (training, p > 0), in which case cat+crop+scale
    is one materialising pass."""

    def __init__(self, in_ch, out_ch, dropout_ratio, batch_norm=False):
        super(res_up, self).__init__()
        self.dconv = nn.ConvTranspose2d(in_ch, out_ch, kernel_size=3, stride=2, padding=0)
        self.uconv = nn.Sequential(nn.Dropout2d(p=dropout_ratio, inplace=True), middle(in_ch, out_ch, batch_norm=False))

    def run(self, x1, x2, math):
        Hu, Wu = 2 * x1.shape[1] + 1, 2 * x1.shape[2] + 1
        diffX = x2.shape[1] - Hu
        diffY = x2.shape[2] - Wu
        if diffX > 0 or diffY > 0:
            raise RuntimeError("res_up: skip connection larger than the upsampled tensor is not supported")
        # F.pad(x, (diffX//2, int(diffX/2), diffY//2, int(diffY/2))): first pair crops W, second pair crops H
        Wo = Wu + diffX // 2 + int(diffX / 2)
        Ho = Hu + diffY // 2 + int(diffY / 2)
        if Ho != x2.shape[1] or Wo != x2.shape[2]:
            raise RuntimeError("Sizes of tensors must match except in dimension 1")
        oy, ox = -(diffY // 2), -(diffX // 2)
        Cin_t, Cout_t = self.dconv.weight.shape[0], self.dconv.weight.shape[1]
        if math == ops.MATH_TF32 and ops.convT3x3s2_tc_ok(Cin_t, Cout_t, x1.shape[1], x1.shape[2], Ho, Wo, oy, ox):
            # zero insertion + tcgen05 conv3x3 with the flipped kernel (4x the MACs, >5x faster than the CUDA-core kernel)
            u = ops.convT3x3s2_tc(x1, self.dconv.weight, self.dconv.bias, Ho, Wo, oy, ox)
        else:
            u = ops.convT3x3s2(x1, self.dconv.weight, self.dconv.bias, None, Ho, Wo, oy, ox, math == ops.MATH_TF32)
        p = self.uconv[0].p
        if self.training and p > 0:
            scale = _feature_noise(u.shape[0], u.shape[3] + x2.shape[3], p, u.device)
            cat = ops.concat_scale(u, x2, scale, Ho, Wo, 0, 0, 0, 0, math == ops.MATH_TF32)
            return self.uconv[1].run(cat, math)
        return self.uconv[1].run(u, math, x1=x2, H=Ho, W=Wo)


class UNetpRes(_PlasticBase):
    def __init__(self, n_channels, n_classes, device, neurons=16, dropout_ratio=0.5, alfa_type='free', rule='hebb', nbf=128,
                 batch_norm=False, bilinear_upsample=False, batched=False, depth=4):
        """
        Residual plastic U-Net (reference unet_p_res.py:9-69).  Arguments as in the reference (bilinear_upsample is
        accepted and ignored there too).  Extensions: batched (accept B>1), depth (encoder levels).
        """
        super(UNetpRes, self).__init__()
        self._init_plastic(n_channels, n_classes, device, alfa_type, rule, nbf, batched)
        self.depth = depth
        ch_in = n_channels
        for k in range(1, depth + 1):
            ch_out = neurons * (2 ** (k - 1))
            setattr(self, "conv%d" % k, res_down(ch_in, ch_out, batch_norm=batch_norm))
            setattr(self, "pool%d" % k, pool_drop(dropout_ratio=dropout_ratio / 2 if k == 1 else dropout_ratio))
            ch_in = ch_out
        self.mid = middle(ch_in, ch_in * 2, batch_norm=batch_norm)
        for k in range(depth, 0, -1):
            setattr(self, "uconv%d" % k, res_up(neurons * (2 ** k), neurons * (2 ** (k - 1)), dropout_ratio=dropout_ratio,
                                                batch_norm=batch_norm))
        self.outc = outconv(neurons, n_classes)
        self.to(device)
        print("UNet plastic model with plastic rule [%s] initialized" % self.rule)

    def forward(self, x, hebb):
        if x.shape[0] != 1 and not self.batched:
            raise RuntimeError("shape '[%d, %d]' is invalid for input of size %d (batch size 1 only; pass batched=True)"
                               % (self.nbf, self.nbf, x.numel()))
        m = self._math
        x = self._to_nhwc(x)
        skips = []
        for k in range(1, self.depth + 1):
            xc = getattr(self, "conv%d" % k).run(x, m)
            skips.append(xc)
            x = getattr(self, "pool%d" % k).run(xc)
        y = self.mid.run(x, m)
        for k in range(self.depth, 0, -1):
            y = getattr(self, "uconv%d" % k).run(y, skips[k - 1], m)
        o = self.outc.run(y)
        return self._plastic(o, hebb)


# --------------------------------------------------------------------------------------------------
# UNetpCoord — CoordConv stem (coord_conv_script.py:61-126,146-200) + plastic head
# --------------------------------------------------------------------------------------------------
class coord_stem(nn.Module):
    """AddCoords (+2 channels, +3 with_r) -> 1x1 conv -> ReLU.  The coordinate channels are generated
    inside the kernel, never materialised."""

    def __init__(self, in_ch, out_ch, with_r=False):
        super(coord_stem, self).__init__()
        self.coords = 3 if with_r else 2
        self.conv = nn.Conv2d(in_ch + self.coords, out_ch, 1)

    def run(self, x, math=ops.MATH_FP32):
        w = self.conv.weight
        return ops.conv1x1(x, w.view(w.shape[0], w.shape[1]), self.conv.bias, self.coords, True, math == ops.MATH_TF32)


class coord_up(nn.Module):
    """Conv2DTranspose(C/2, 2x2, s2) -> concatenate([up, skip]) -> (Conv3x3 + ReLU) x2  (coord_conv_script.py:174-192)."""

    def __init__(self, in_ch, out_ch):
        super(coord_up, self).__init__()
        self.up = nn.ConvTranspose2d(in_ch, out_ch, 2, stride=2)
        self.conv = double_conv(2 * out_ch, out_ch, False)

    def run(self, x1, x2, math, premask=False):
        m2 = None
        if premask:  # Masked inputs (see `up.run`): the transposed conv's dgrad masks with x1 itself, the skip brings its packed mask
            x1, x2, m2 = x1.t, x2.t, x2.m
        u = ops.convT2x2s2(x1, self.up.weight, self.up.bias, math == ops.MATH_TF32, premask, math == ops.MATH_MIXED)
        if u.shape[1] != x2.shape[1] or u.shape[2] != x2.shape[2]:
            raise RuntimeError("UNetpCoord needs H, W divisible by 16 (Keras 'same' padding has no crop)")
        return self.conv.run(u, math, x1=x2, m1=m2, premask=premask)


class UNetpCoord(_PlasticBase):
    def __init__(self, n_channels, n_classes, device, alfa_type='free', rule='oja', nbf=128, with_r=False, batched=False,
                 base=8, depth=4):
        """Plastic coord-conv U-Net: CoordConv(1x1, base, relu) stem, widths base*[1,2,4,8,16], cat [up, skip],
        1x1 output conv, then the plastic head of UNetp in place of the Keras sigmoid."""
        super(UNetpCoord, self).__init__()
        self._init_plastic(n_channels, n_classes, device, alfa_type, rule, nbf, batched)
        self.depth = depth
        self.stem = coord_stem(n_channels, base, with_r=with_r)
        c = [base * (2 ** i) for i in range(depth + 1)]
        self.enc0 = double_conv(base, c[0], False)
        for k in range(1, depth + 1):
            setattr(self, "enc%d" % k, down(c[k - 1], c[k], batch_norm=False))
        for k in range(depth, 0, -1):
            setattr(self, "dec%d" % k, coord_up(c[k], c[k - 1]))
        self.outc = outconv(c[0], n_classes)
        self.to(device)
        print("UNet plastic model with plastic rule [%s] initialized" % self.rule)

    def forward(self, x, hebb):
        if x.shape[0] != 1 and not self.batched:
            raise ValueError("Only batch size: 1 is supported, but was: %d" % x.shape[0])
        m = self._math
        x = self._to_nhwc(x)
        # premasked-gradient protocol + pool_skip as in UNetp.forward (the CoordConv stem keeps its own ReLU-mask pass: the
        # 1x1 kernel emits no packed mask)
        pm = (m in _TC_BWD and self.premask and self.outc.conv.weight.shape[1] in (8, 16, 32, 64) and self.n_classes == 1
              and self.stem.conv.weight.shape[0] % 8 == 0)
        feats = []
        f = self.enc0.run(self.stem.run(x, m), m, premask=pm)
        for k in range(1, self.depth + 1):
            ft, fm = (f.t, f.m) if pm else (f, None)
            pooled, skip = ops.pool_skip(ft, pm)
            feats.append(Masked(skip, fm) if pm else skip)
            f = getattr(self, "enc%d" % k).run_pooled(pooled, m, pm)
        feats.append(f)
        y = feats[-1]
        for k in range(self.depth, 0, -1):
            y = getattr(self, "dec%d" % k).run(y, feats[k - 1], m, pm)
        o = self.outc.run(y, pm)
        return self._plastic(o, hebb)
