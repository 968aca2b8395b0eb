"""torch.library custom ops over the C-ABI kernels (namespace ``pu``).

Every op inside the Plastic U-Net modules is one of these: a ``torch.library.custom_op`` with a fake
(meta) implementation and an autograd formula whose backward is again a registered custom op.  All
tensors are fp32 CUDA tensors of logical shape [B, H, W, C] (NHWC, contiguous).  There is no CPU,
cuDNN or Triton fallback: a non-CUDA tensor raises.
"""
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib

import functools
import os

MATH_FP32 = 0
MATH_TF32 = 1
RULE_HEBB = 0
RULE_OJA = 1
FLAG_RELU = 1
FLAG_ROUND_TF32 = 2
FLAG_MASK_IN = 4
FLAG_TF32_MATH = 8


@functools.lru_cache(maxsize=None)
def _tc_ok(C0: int, C1: int, Cout: int, Cd0: int, Cd1: int) -> bool:
    """Can a conv3x3 with these source/destination channel splits run on the tcgen05 path?"""
    return bool(_lib.load().pu_conv3x3_tc_ok(C0, C1, Cout, Cd0, Cd1))


@functools.lru_cache(maxsize=None)
def _tc_resident(C0: int, C1: int, Cout: int, H: int, W: int) -> bool:
    return bool(_lib.load().pu_conv3x3_tc_resident(C0, C1, Cout, H, W))


W_PACKED, W_OIHW, W_OIHW_DGRAD = 0, 1, 2

# set by pu_b200.trainer.TrainStep: CUDA streams on which the backward ops run their parameter-gradient kernels, used
# round-robin (None = everything on the current stream)
WGRAD_SIDE_STREAMS = None
_side_rr = 0


# debug hook of TrainStep: when a set, every tensor a backward op WRITES on a side stream registers its data_ptr here, so
# that the trainer can assert that autograd handed exactly those tensors over as .grad (stolen, not copied on the main
# stream while the side stream is still writing)
SIDE_OUTPUTS = None


def _note_side(*ts):
    if SIDE_OUTPUTS is not None:
        for t in ts:
            if t is not None and t.numel() > 0:
                SIDE_OUTPUTS.add(t.data_ptr())


# set by pu_b200.trainer.TrainStep: {weight.data_ptr(): (flat gradient arena, weight offset, bias offset | -1, bias numel)}.
# A backward op whose kernels can ACCUMULATE (PU_MATH_ACCUM / PU_FLAG_ACCUM_GRADS) then writes its parameter gradients straight
# into the arena slots (zeroed once per step by the trainer): no memset launches per op, no gradient gather before Adam.  The
# custom ops receive the slot ADDRESSES (ints; their tensor outputs may not alias one another's storage) and the autograd
# wrappers hand views of the slots to autograd.
GRAD_SINK = None
FLAG_ACCUM_GRADS = 16
FLAG_DEFER_FINISH = 32
MATH_ACCUM = 0x100


def _sink(weight: Tensor, want_bias: bool):
    """-> (dw view, db view | None) inside the trainer's gradient arena, or None.  Fresh view objects every call, so that
    autograd's AccumulateGrad takes them over as .grad (no copy kernel)."""
    if GRAD_SINK is None:
        return None
    ent = GRAD_SINK.get(weight.data_ptr())
    if ent is None:
        return None
    arena, woff, boff, bn = ent
    if want_bias and boff < 0:
        return None
    dw = arena[woff:woff + weight.numel()].view(weight.shape)
    db = arena[boff:boff + bn] if want_bias else None
    return dw, db


def _fork_point():
    """An event on the current stream at the point where every input of an op's parameter-gradient kernels exists.  The side
    stream waits for THIS (not for the whole main stream), so the weight gradient of a layer does not queue behind the layer's
    own data-gradient kernel, which the op enqueues on the main stream first (critical path first)."""
    if not WGRAD_SIDE_STREAMS:
        return None
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    return ev


def _side_stream():
    global _side_rr
    if not WGRAD_SIDE_STREAMS:
        return None
    _side_rr = (_side_rr + 1) % len(WGRAD_SIDE_STREAMS)
    return WGRAD_SIDE_STREAMS[_side_rr]



MATH_TF32_FLAT = 2  # pu_pack_w3x3 only
# Python-level mode (never passed to the C-ABI): strict-fp32 FORWARD kernels (exact activations and ReLU masks), TF32
# tensor-core BACKWARD kernels (tcgen05 dgrad, mma.sync wgrad / transposed convs) — see DESIGN.md §2
MATH_MIXED = 4


@functools.lru_cache(maxsize=None)
def _tc_flat(B: int, H: int, W: int, C0: int, C1: int, Cout: int) -> bool:
    """Will the tcgen05 conv run this problem in its flat mode (wide layers)?  Its packed weights differ then."""
    return bool(_lib.load().pu_conv3x3_tc_flat(B, H, W, C0, C1, Cout))


def _weight_operand(weight: Tensor, transpose: int, math: int, C0: int, C1: int, Cout_conv: int, H: int, W: int, B: int = 1):
    """-> (tensor, wfmt): the raw OIHW weight whenever the kernel can build its operand tiles itself (every FFMA
    conv, and tcgen05 convs whose weight image fits in shared memory), else the pu_pack_w3x3 buffer."""
    if math == MATH_FP32:
        return weight, (W_OIHW_DGRAD if transpose else W_OIHW)
    if _tc_resident(C0, C1, Cout_conv, H, W):
        return weight, (W_OIHW_DGRAD if transpose else W_OIHW)
    if _tc_flat(B, H, W, C0, C1, Cout_conv):  # the wide layers' K chunking differs: their weight image is packed accordingly
        return _pack_w(weight, transpose, MATH_TF32_FLAT, C0), W_PACKED
    return _pack_w(weight, transpose, math, C0), W_PACKED


# Weight images that do not fit in shared memory are packed by a separate kernel (pu_pack_w3x3).  The weights only change
# in the optimizer, so TrainStep packs them all at the START of the step on a side stream, off the critical path:
#   PACK_LOG   (a list, during the first warm-up step) records which (weight, transpose, math, C0) the step packs;
#   PACK_CACHE {(data_ptr, transpose, math, C0): (packed tensor, event)} is then filled by TrainStep every step.
PACK_LOG = None
PACK_CACHE = None


def _pack_w(weight: Tensor, transpose: int, math: int, C0: int) -> Tensor:
    key = (weight.data_ptr(), transpose, math, C0)
    if PACK_CACHE is not None and key in PACK_CACHE:
        wp, ev = PACK_CACHE[key]
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        return wp
    if PACK_LOG is not None:
        PACK_LOG.append((weight, transpose, math, C0))
    return _pack_w_now(weight, transpose, math, C0)


def _pack_w_now(weight: Tensor, transpose: int, math: int, C0: int) -> Tensor:
    Cout, Cin = weight.shape[0], weight.shape[1]
    n = int(_lib.load().pu_pack_w3x3_floats(Cout, Cin, transpose, math, C0))
    wp = torch.empty(n, device=weight.device, dtype=torch.float32)
    _lib.call("pu_pack_w3x3", weight.data_ptr(), wp.data_ptr(), Cout, Cin, transpose, math, C0, _s())
    return wp


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _chk(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("pu_b200 ops run only on CUDA tensors (sm_100a); there is no CPU fallback")
        if t.dtype != torch.float32:
            raise RuntimeError(f"pu_b200 ops are fp32; got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError("pu_b200 ops need contiguous NHWC tensors")


def _e(dev) -> Tensor:
    """A fresh empty placeholder (custom-op outputs may not alias each other)."""
    return torch.empty(0, device=dev, dtype=torch.float32)


def _dims(t: Optional[Tensor]) -> Tuple[int, int, int]:
    """(H, W, C) of an NHWC tensor, zeros for None."""
    if t is None:
        return 0, 0, 0
    return t.shape[1], t.shape[2], t.shape[3]


# =================================================================================================
# conv3x3 (+ concat/crop of two sources, + bias, + residual, + ReLU)
# =================================================================================================
def mask_like(y: Tensor) -> Tensor:
    """An (uninitialised) packed ReLU-mask buffer for the NHWC tensor y: uint8 [B, H, W, C/8], bit j of byte g = channel 8g+j > 0."""
    if y.shape[-1] % 8 != 0:
        raise RuntimeError("packed ReLU masks need a channel count that is a multiple of 8")
    return torch.empty(y.shape[:-1] + (y.shape[-1] // 8,), device=y.device, dtype=torch.uint8)


def _chk_mask(m: Optional[Tensor], like: Optional[Tensor], what: str) -> None:
    if m is None:
        return
    if like is None or not m.is_cuda or m.dtype != torch.uint8 or not m.is_contiguous() or \
            tuple(m.shape) != tuple(like.shape[:-1]) + (like.shape[-1] // 8,):
        raise RuntimeError("conv3x3: %s must be a contiguous uint8 CUDA tensor [B, H, W, C/8] matching its tensor" % what)


@torch.library.custom_op("pu::conv3x3_m", mutates_args=())
def conv3x3_m(x0: Tensor, x1: Optional[Tensor], weight: Tensor, bias: Optional[Tensor], res: Optional[Tensor],
              relu: bool, H: int, W: int, oy0: int, ox0: int, oy1: int, ox1: int, math: int,
              m0: Optional[Tensor] = None, m1: Optional[Tensor] = None, premasked: bool = False,
              emit_mask: bool = False) -> Tuple[Tensor, Tensor]:
    """-> (y, ymask).  m0/m1/premasked/emit_mask implement the premasked-gradient protocol.  `premasked` =
    every consumer of this op's ReLU output multiplies the gradient it sends back by (y > 0), so this op's backward skips
    its own mask pass; `m0`/`m1` = the PACKED ReLU mask (mask_like) of source 0/1, which is such a ReLU output: this
    op's dgrad applies it in its epilogue (backward only); `emit_mask`: the forward epilogue also writes the packed mask
    of this op's own output (ymask = mask_like(y); an empty tensor otherwise), to be handed to its consumers as m0/m1."""
    _chk(x0, x1, weight, bias, res)
    _chk_mask(m0, x0, "m0")
    _chk_mask(m1, x1, "m1")
    B = x0.shape[0]
    Cout, Cin = weight.shape[0], weight.shape[1]
    H0, W0, C0 = _dims(x0)
    H1, W1, C1 = _dims(x1)
    if C0 + C1 != Cin:
        raise RuntimeError(f"conv3x3: weight expects {Cin} input channels, sources have {C0}+{C1}")
    # PU_MATH_TF32: tcgen05 kernel where the channel counts allow, fp32 FFMA kernel (output rounded to TF32) elsewhere
    m = MATH_TF32 if (math == MATH_TF32 and _tc_ok(C0, C1, Cout, Cout, 0)) else MATH_FP32
    flags = (FLAG_RELU if relu else 0) | (FLAG_ROUND_TF32 if math == MATH_TF32 else 0)
    wp, wfmt = _weight_operand(weight, 0, m, C0, C1, Cout, H, W, B)
    y = torch.empty((B, H, W, Cout), device=x0.device, dtype=torch.float32)
    ymask = mask_like(y) if emit_mask else torch.empty(0, device=x0.device, dtype=torch.uint8)
    _lib.call("pu_conv3x3_fwd", x0.data_ptr(), H0, W0, C0, oy0, ox0, _p(x1), H1, W1, C1, oy1, ox1,
              wp.data_ptr(), _p(bias), _p(res), flags,
              y.data_ptr(), H, W, Cout, 0, 0, None, 0, 0, 0, 0, 0, None, None, ymask.data_ptr() if emit_mask else None,
              B, H, W, Cout, m, wfmt, _s())
    return y, ymask


@conv3x3_m.register_fake
def _(x0, x1, weight, bias, res, relu, H, W, oy0, ox0, oy1, ox1, math, m0=None, m1=None, premasked=False, emit_mask=False):
    y = x0.new_empty((x0.shape[0], H, W, weight.shape[0]))
    return y, (x0.new_empty((x0.shape[0], H, W, weight.shape[0] // 8), dtype=torch.uint8) if emit_mask
               else x0.new_empty(0, dtype=torch.uint8))


def conv3x3(x0: Tensor, x1: Optional[Tensor], weight: Tensor, bias: Optional[Tensor], res: Optional[Tensor],
            relu: bool, H: int, W: int, oy0: int, ox0: int, oy1: int, ox1: int, math: int,
            m0: Optional[Tensor] = None, m1: Optional[Tensor] = None, premasked: bool = False) -> Tensor:
    """conv3x3 without the packed-mask output (see conv3x3_m)."""
    return conv3x3_m(x0, x1, weight, bias, res, relu, H, W, oy0, ox0, oy1, ox1, math, m0, m1, premasked, False)[0]


@torch.library.custom_op("pu::conv3x3_bwd", mutates_args=())
def conv3x3_bwd(dy: Tensor, y: Tensor, x0: Tensor, x1: Optional[Tensor], weight: Tensor, has_bias: bool, relu: bool,
                H: int, W: int, oy0: int, ox0: int, oy1: int, ox1: int, math: int,
                need_dx: bool, need_dw: bool, m0: Optional[Tensor] = None, m1: Optional[Tensor] = None,
                premasked: bool = False, dw_ptr: int = 0, db_ptr: int = 0) -> List[Tensor]:
    """-> [g, dx0, dx1, dw, db]; g = dy masked by the fused ReLU (== dy when relu is False).  dw_ptr / db_ptr != 0: the
    (zeroed) gradient-arena slots the weight-gradient kernel accumulates into (see GRAD_SINK); dw / db are then returned empty."""
    _chk(dy, y, x0, x1, weight)
    dev = dy.device
    B = x0.shape[0]
    Cout, Cin = weight.shape[0], weight.shape[1]
    H0, W0, C0 = _dims(x0)
    H1, W1, C1 = _dims(x1)
    npix = B * H * W
    tf32 = math in (MATH_TF32, MATH_MIXED)
    premasked = premasked and relu
    fresh_g = (relu or tf32) and not premasked  # tensor-core operands are stored rounded to TF32 by their producer
    db_in_wgrad = premasked and has_bias and need_dw
    db_sunk = db_in_wgrad and db_ptr != 0
    db = torch.empty(Cout, device=dev, dtype=torch.float32) if (has_bias and not db_sunk) else _e(dev)
    dbp = db_ptr if db_sunk else (db.data_ptr() if db_in_wgrad else None)
    if premasked:
        g = dy  # every consumer already applied (y > 0) (and rounded) in its own backward epilogue
        if has_bias and not need_dw:
            _lib.call("pu_relu_bwd_bias", dy.data_ptr(), None, None, db.data_ptr(), npix, Cout, 0, _s())
    elif fresh_g:
        g = torch.empty_like(dy)
        _lib.call("pu_relu_bwd_bias", dy.data_ptr(), y.data_ptr() if relu else None, g.data_ptr(), _p(db) if has_bias else None,
                  npix, Cout, (FLAG_RELU if relu else 0) | (FLAG_ROUND_TF32 if tf32 else 0), _s())
    else:
        g = dy
        if has_bias:
            _lib.call("pu_relu_bwd_bias", dy.data_ptr(), None, None, db.data_ptr(), npix, Cout, 0, _s())
    dx0, dx1 = _e(dev), _e(dev)
    fork = _fork_point() if need_dw else None  # g is complete here
    if need_dx:
        md = MATH_TF32 if (tf32 and _tc_ok(Cout, 0, Cin, C0, C1)) else MATH_FP32
        wpt, wfmt = _weight_operand(weight, 1, md, Cout, 0, Cin, H, W, B)
        full0 = (H0 == H and W0 == W)
        dx0 = (torch.empty if full0 else torch.zeros)((B, H0, W0, C0), device=dev, dtype=torch.float32)
        if x1 is not None:
            full1 = (H1 == H and W1 == W)
            dx1 = (torch.empty if full1 else torch.zeros)((B, H1, W1, C1), device=dev, dtype=torch.float32)
        # the packed masks share the geometry of the gradient tensors they gate (the op's window is a pointer offset)
        _lib.call("pu_conv3x3_fwd", g.data_ptr(), H, W, Cout, 0, 0, None, 0, 0, 0, 0, 0,
                  wpt.data_ptr(), None, None, FLAG_ROUND_TF32 if tf32 else 0,
                  dx0.data_ptr(), H0, W0, C0, oy0, ox0,
                  dx1.data_ptr() if x1 is not None else None, H1, W1, C1, oy1, ox1,
                  _p(m0), _p(m1) if x1 is not None else None, None,
                  B, H, W, Cin, md, wfmt, _s())
    dw = _e(dev)
    if need_dw:
        wmath = (MATH_TF32 if tf32 else MATH_FP32) | (MATH_ACCUM if dw_ptr else 0)
        side = _side_stream()
        if side is not None:
            # The weight gradient is not needed before the optimizer: run it on a side stream so that it overlaps the
            # dgrad chain (the caller joins WGRAD_SIDE_STREAMS before it reads any parameter gradient).
            main = torch.cuda.current_stream()
            side.wait_event(fork)
            with torch.cuda.stream(side):
                if not dw_ptr:
                    dw = torch.empty_like(weight)
                _lib.call("pu_conv3x3_wgrad", x0.data_ptr(), H0, W0, C0, oy0, ox0, _p(x1), H1, W1, C1, oy1, ox1,
                          g.data_ptr(), dw_ptr if dw_ptr else dw.data_ptr(), dbp, B, H, W, Cout, wmath, _s())
            for t in (x0, x1, g, db if (db_in_wgrad and not db_sunk) else None):
                if t is not None:
                    t.record_stream(side)
            if not dw_ptr:
                dw.record_stream(main)
            _note_side(dw if not dw_ptr else None, db if (db_in_wgrad and not db_sunk) else None)
        else:
            if not dw_ptr:
                dw = torch.empty_like(weight)
            _lib.call("pu_conv3x3_wgrad", x0.data_ptr(), H0, W0, C0, oy0, ox0, _p(x1), H1, W1, C1, oy1, ox1,
                      g.data_ptr(), dw_ptr if dw_ptr else dw.data_ptr(), dbp, B, H, W, Cout, wmath, _s())
    g_out = g if fresh_g else _e(dev)  # never return an alias of an input
    return [g_out, dx0, dx1, dw, db]


@conv3x3_bwd.register_fake
def _(dy, y, x0, x1, weight, has_bias, relu, H, W, oy0, ox0, oy1, ox1, math, need_dx, need_dw, m0=None, m1=None,
      premasked=False, dw_ptr=0, db_ptr=0):
    e = dy.new_empty(0)
    return [torch.empty_like(dy) if ((relu or math in (MATH_TF32, MATH_MIXED)) and not (premasked and relu)) else e,
            torch.empty_like(x0) if need_dx else e,
            torch.empty_like(x1) if (need_dx and x1 is not None) else e,
            torch.empty_like(weight) if (need_dw and not dw_ptr) else e,
            dy.new_empty(weight.shape[0]) if (has_bias and not (db_ptr and premasked and relu and need_dw)) else e]


def _conv3x3_setup(ctx, inputs, output):
    x0, x1, weight, bias, res, relu, H, W, oy0, ox0, oy1, ox1, math, m0, m1, premasked, _emit = inputs
    ctx.save_for_backward(x0, x1, weight, output[0], m0, m1)
    ctx.mark_non_differentiable(output[1])
    ctx.set_materialize_grads(False)  # no zero-filled "gradient" tensor for the packed mask on every backward
    ctx.cfg = (bias is not None, res is not None, relu, H, W, oy0, ox0, oy1, ox1, math, premasked)


def _conv3x3_backward(ctx, dy, _dmask=None):
    x0, x1, weight, y, m0, m1 = ctx.saved_tensors
    has_bias, has_res, relu, H, W, oy0, ox0, oy1, ox1, math, premasked = ctx.cfg
    need = ctx.needs_input_grad
    if dy is None:  # (materialize_grads is off) nothing flows back through this conv
        return (None,) * 17
    need_dx = need[0] or (x1 is not None and need[1])
    dy = dy.contiguous()
    # gradient arena of the trainer: the TF32 kernels (channel multiples of 8, or the one-channel stem) accumulate in place
    sink = None
    want_b = bool(has_bias and need[3])
    if need[2] and math in (MATH_TF32, MATH_MIXED) and GRAD_SINK is not None:
        C0, C1, Cout = x0.shape[3], (x1.shape[3] if x1 is not None else 0), weight.shape[0]
        if (C0 % 8 == 0 and C1 % 8 == 0 and Cout % 8 == 0) or (C0 == 1 and C1 == 0 and Cout in (8, 16)):
            db_in_wgrad = want_b and premasked and relu
            sink = _sink(weight, db_in_wgrad)
            if sink is None and db_in_wgrad:
                sink = _sink(weight, False)
    g, dx0, dx1, dw, db = conv3x3_bwd(dy, y, x0, x1, weight, want_b, relu, H, W, oy0, ox0, oy1, ox1, math,
                                      need_dx, need[2], m0, m1, premasked,
                                      sink[0].data_ptr() if sink is not None else 0,
                                      sink[1].data_ptr() if (sink is not None and sink[1] is not None) else 0)
    if sink is not None:
        dw = sink[0]
        if sink[1] is not None:
            db = sink[1]
        if WGRAD_SIDE_STREAMS:
            _note_side(dw, sink[1])
    gres = None
    if has_res and need[4]:
        gres = g if ((relu or math in (MATH_TF32, MATH_MIXED)) and not (premasked and relu)) else dy
    return (dx0 if need[0] else None, dx1 if (x1 is not None and need[1]) else None, dw if need[2] else None,
            db if (has_bias and need[3]) else None, gres, None, None, None, None, None, None, None, None, None, None, None, None)


conv3x3_m.register_autograd(_conv3x3_backward, setup_context=_conv3x3_setup)


# =================================================================================================
# conv1x1 (+ analytic CoordConv channels, + ReLU)
# =================================================================================================
@torch.library.custom_op("pu::conv1x1", mutates_args=())
def conv1x1(x: Tensor, weight: Tensor, bias: Optional[Tensor], coords: int, relu: bool, round_out: bool = False,
            mask_in: bool = False) -> Tensor:
    _chk(x, weight, bias)
    B, H, W, Cin = x.shape
    Cout = weight.shape[0]
    if weight.shape[1] != Cin + coords:
        raise RuntimeError("conv1x1: weight/in-channel mismatch")
    y = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.float32)
    _lib.call("pu_conv1x1_fwd", x.data_ptr(), weight.data_ptr(), _p(bias), y.data_ptr(), B, H, W, Cin, Cout, coords,
              (FLAG_RELU if relu else 0) | (FLAG_ROUND_TF32 if round_out else 0), _s())
    return y


@conv1x1.register_fake
def _(x, weight, bias, coords, relu, round_out=False, mask_in=False):
    return x.new_empty((x.shape[0], x.shape[1], x.shape[2], weight.shape[0]))


@torch.library.custom_op("pu::conv1x1_bwd", mutates_args=())
def conv1x1_bwd(dy: Tensor, y: Tensor, x: Tensor, weight: Tensor, coords: int, relu: bool, need_dx: bool,
                mask_in: bool = False) -> List[Tensor]:
    _chk(dy, y, x, weight)
    B, H, W, Cin = x.shape
    Cout = weight.shape[0]
    dev = x.device
    g = dy
    if relu:
        g = torch.empty_like(dy)
        _lib.call("pu_relu_bwd_bias", dy.data_ptr(), y.data_ptr(), g.data_ptr(), None, B * H * W, Cout, 1, _s())
    dx = torch.empty_like(x) if need_dx else _e(dev)
    dw = torch.empty_like(weight)
    db = torch.empty(Cout, device=dev, dtype=torch.float32)
    ws = torch.empty(Cout * (Cin + coords + 1), device=dev, dtype=torch.float32)
    side = _side_stream()
    flags = (FLAG_MASK_IN if mask_in else 0) | (FLAG_DEFER_FINISH if side is not None else 0)
    _lib.call("pu_conv1x1_bwd", x.data_ptr(), weight.data_ptr(), g.data_ptr(), dx.data_ptr() if need_dx else None,
              dw.data_ptr(), db.data_ptr(), ws.data_ptr(), B, H, W, Cin, Cout, coords, flags, _s())
    if side is not None:
        # the scatter of the parameter gradients out of the scratch leaves the data-gradient chain (TrainStep's side streams)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _lib.call("pu_conv1x1_dw_finish", ws.data_ptr(), dw.data_ptr(), db.data_ptr(), Cin, Cout, coords, _s())
        for t in (ws, dw, db):
            t.record_stream(side)
        _note_side(dw, db)
    return [dx, dw, db]


@conv1x1_bwd.register_fake
def _(dy, y, x, weight, coords, relu, need_dx, mask_in=False):
    return [torch.empty_like(x) if need_dx else x.new_empty(0), torch.empty_like(weight), x.new_empty(weight.shape[0])]


def _conv1x1_setup(ctx, inputs, output):
    x, weight, bias, coords, relu, _round, mask_in = inputs
    ctx.save_for_backward(x, weight, output)
    ctx.cfg = (bias is not None, coords, relu, mask_in)


def _conv1x1_backward(ctx, dy):
    x, weight, y = ctx.saved_tensors
    has_bias, coords, relu, mask_in = ctx.cfg
    need = ctx.needs_input_grad
    dx, dw, db = conv1x1_bwd(dy.contiguous(), y, x, weight, coords, relu, need[0], mask_in)
    return dx if need[0] else None, dw if need[1] else None, db if (has_bias and need[2]) else None, None, None, None, None


conv1x1.register_autograd(_conv1x1_backward, setup_context=_conv1x1_setup)


# =================================================================================================
# transposed convolutions
# =================================================================================================
@torch.library.custom_op("pu::convT2x2s2", mutates_args=())
def convT2x2s2(x: Tensor, weight: Tensor, bias: Optional[Tensor], round_out: bool = False, mask_in: bool = False,
               bwd_tf32: bool = False) -> Tensor:
    """round_out: the model's TF32 mode (TF32 forward math, output rounded, TF32 backward); bwd_tf32: MIXED mode (strict-fp32
    forward, TF32 mma.sync backward)."""
    _chk(x, weight, bias)
    B, H, W, Cin = x.shape
    Cout = weight.shape[1]
    y = torch.empty((B, 2 * H, 2 * W, Cout), device=x.device, dtype=torch.float32)
    _lib.call("pu_convT2x2s2_fwd", x.data_ptr(), weight.data_ptr(), _p(bias), y.data_ptr(), B, H, W, Cin, Cout,
              (FLAG_ROUND_TF32 | FLAG_TF32_MATH) if round_out else 0, _s())  # round_out == the model's TF32 mode
    return y


@convT2x2s2.register_fake
def _(x, weight, bias, round_out=False, mask_in=False, bwd_tf32=False):
    return x.new_empty((x.shape[0], 2 * x.shape[1], 2 * x.shape[2], weight.shape[1]))


@torch.library.custom_op("pu::convT2x2s2_bwd", mutates_args=())
def convT2x2s2_bwd(dy: Tensor, x: Tensor, weight: Tensor, need_dx: bool, need_dw: bool, need_db: bool,
                   mask_in: bool = False, tf32: bool = False, dw_ptr: int = 0, db_ptr: int = 0) -> List[Tensor]:
    """dw_ptr and db_ptr != 0: the (zeroed) gradient-arena slots the kernels accumulate into (GRAD_SINK); dw / db are returned empty."""
    _chk(dy, x, weight)
    B, H, W, Cin = x.shape
    Cout = weight.shape[1]
    dx = torch.empty_like(x) if need_dx else _e(x.device)
    sunk = bool(dw_ptr and db_ptr and need_dw and need_db)
    dw = torch.empty_like(weight) if (need_dw and not sunk) else _e(x.device)
    db = torch.empty(Cout, device=x.device, dtype=torch.float32) if (need_db and not sunk) else _e(x.device)
    flags = (FLAG_MASK_IN if mask_in else 0) | (FLAG_TF32_MATH if tf32 else 0) | (FLAG_ACCUM_GRADS if sunk else 0)
    dwp = dw_ptr if sunk else (dw.data_ptr() if need_dw else None)
    dbp = db_ptr if sunk else (db.data_ptr() if need_db else None)
    side = _side_stream()
    if side is not None and (need_dw or need_db):
        # dx on the current stream; the parameter gradients (not needed before the optimizer) on the side stream
        fork = _fork_point()
        if need_dx:
            _lib.call("pu_convT2x2s2_bwd", x.data_ptr(), weight.data_ptr(), dy.data_ptr(), dx.data_ptr(), None, None,
                      B, H, W, Cin, Cout, flags, _s())
        main = torch.cuda.current_stream()
        side.wait_event(fork)
        with torch.cuda.stream(side):
            _lib.call("pu_convT2x2s2_bwd", x.data_ptr(), weight.data_ptr(), dy.data_ptr(), None, dwp, dbp, B, H, W, Cin, Cout, flags, _s())
        for t in (x, dy, dw if (need_dw and not sunk) else None, db if (need_db and not sunk) else None):
            if t is not None:
                t.record_stream(side)
        _note_side(dw if (need_dw and not sunk) else None, db if (need_db and not sunk) else None)
    else:
        _lib.call("pu_convT2x2s2_bwd", x.data_ptr(), weight.data_ptr(), dy.data_ptr(), dx.data_ptr() if need_dx else None,
                  dwp, dbp, B, H, W, Cin, Cout, flags, _s())
    return [dx, dw, db]


@convT2x2s2_bwd.register_fake
def _(dy, x, weight, need_dx, need_dw, need_db, mask_in=False, tf32=False, dw_ptr=0, db_ptr=0):
    e = x.new_empty(0)
    sunk = bool(dw_ptr and db_ptr and need_dw and need_db)
    return [torch.empty_like(x) if need_dx else e, torch.empty_like(weight) if (need_dw and not sunk) else e,
            x.new_empty(weight.shape[1]) if (need_db and not sunk) else e]


def _convT2_setup(ctx, inputs, output):
    x, weight, bias, round_out, mask_in, bwd_tf32 = inputs
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None
    ctx.mask_in = mask_in
    ctx.tf32 = bool(round_out) or bool(bwd_tf32)


def _convT2_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    need = ctx.needs_input_grad
    Cin, Cout = weight.shape[0], weight.shape[1]
    sink = None
    if ctx.tf32 and need[1] and ctx.has_bias and need[2] and Cin == Cout and Cin in (8, 16, 32, 64):  # the tensor-core kernels' shapes
        sink = _sink(weight, True)
    dx, dw, db = convT2x2s2_bwd(dy.contiguous(), x, weight, need[0], need[1], ctx.has_bias and need[2], ctx.mask_in, ctx.tf32,
                                sink[0].data_ptr() if sink is not None else 0, sink[1].data_ptr() if sink is not None else 0)
    if sink is not None:
        dw, db = sink
        if WGRAD_SIDE_STREAMS:
            _note_side(dw, db)
    return dx if need[0] else None, dw if need[1] else None, db if (ctx.has_bias and need[2]) else None, None, None, None


convT2x2s2.register_autograd(_convT2_backward, setup_context=_convT2_setup)


# -------------------------------------------------------------------------------------------------
# ConvTranspose2d(k=3, s=2, p=0) + crop on the tcgen05 path (TF32 mode): zero insertion + conv3x3 with the flipped kernel
# -------------------------------------------------------------------------------------------------
def zero_window_ok(o: int, n: int, full: int) -> bool:
    """1-D condition for the window [o, o+n) of the zero-inserted canvas (x at the odd coordinates of [0, full)): a
    zero-padded 3-tap convolution of the window equals the window of the full convolution iff the canvas is zero at
    o-1 and at o+n (outside the canvas, or an even coordinate)."""
    return (o == 0 or (o - 1) % 2 == 0) and (o + n == full or (o + n) % 2 == 0)


def convT3x3s2_tc_ok(Cin: int, Cout: int, H: int, W: int, Ho: int, Wo: int, oy: int, ox: int) -> bool:
    """The zero-padded window equals the cropped transposed conv only if the canvas is zero just outside the window:
    before it (index oy-1 / ox-1: outside the canvas or an even coordinate) and after it (end of the canvas or an even
    coordinate).  True for the reference's crops of 0 or 1 (unet_p_res.py:214-217)."""
    return (zero_window_ok(oy, Ho, 2 * H + 1) and zero_window_ok(ox, Wo, 2 * W + 1)
            and _tc_ok(Cin, 0, Cout, Cout, 0) and _tc_ok(Cout, 0, Cin, Cin, 0))


@torch.library.custom_op("pu::convT3x3s2_tc", mutates_args=())
def convT3x3s2_tc(x: Tensor, weight: Tensor, bias: Optional[Tensor], Ho: int, Wo: int, oy: int, ox: int) -> Tensor:
    _chk(x, weight, bias)
    B, H, W, Cin = x.shape
    Cout = weight.shape[1]
    z = torch.empty((B, Ho, Wo, Cin), device=x.device, dtype=torch.float32)
    _lib.call("pu_zero_insert2x_fwd", x.data_ptr(), z.data_ptr(), B, H, W, Cin, Ho, Wo, oy, ox, _s())
    # the conv whose "dgrad" operand is the IOHW transposed-conv weight: input channels = dim 0, outputs = dim 1, taps flipped
    wp, wfmt = _weight_operand(weight, 1, MATH_TF32, Cin, 0, Cout, Ho, Wo, B)
    y = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
    _lib.call("pu_conv3x3_fwd", z.data_ptr(), Ho, Wo, Cin, 0, 0, None, 0, 0, 0, 0, 0,
              wp.data_ptr(), _p(bias), None, FLAG_ROUND_TF32,
              y.data_ptr(), Ho, Wo, Cout, 0, 0, None, 0, 0, 0, 0, 0, None, None, None, B, Ho, Wo, Cout, MATH_TF32, wfmt, _s())
    return y


@convT3x3s2_tc.register_fake
def _(x, weight, bias, Ho, Wo, oy, ox):
    return x.new_empty((x.shape[0], Ho, Wo, weight.shape[1]))


@torch.library.custom_op("pu::convT3x3s2_tc_bwd", mutates_args=())
def convT3x3s2_tc_bwd(dy: Tensor, x: Tensor, weight: Tensor, has_bias: bool, need_dx: bool, need_dw: bool,
                      Ho: int, Wo: int, oy: int, ox: int) -> List[Tensor]:
    _chk(dy, x, weight)
    dev = x.device
    B, H, W, Cin = x.shape
    Cout = weight.shape[1]
    npix = B * Ho * Wo
    db = torch.empty(Cout, device=dev, dtype=torch.float32) if has_bias else _e(dev)
    g = torch.empty_like(dy)  # gradient rounded to TF32 (tensor-core operand) + bias gradient, one pass
    _lib.call("pu_relu_bwd_bias", dy.data_ptr(), None, g.data_ptr(), _p(db) if has_bias else None, npix, Cout, FLAG_ROUND_TF32, _s())
    dx, dw = _e(dev), _e(dev)
    if need_dx:
        # dz = conv3x3(g) with the weight read as OIHW = [Cin][Cout]: input channels = dim 1, outputs = dim 0, taps as stored
        wp, wfmt = _weight_operand(weight, 0, MATH_TF32, Cout, 0, Cin, Ho, Wo, B)
        dz = torch.empty((B, Ho, Wo, Cin), device=dev, dtype=torch.float32)
        _lib.call("pu_conv3x3_fwd", g.data_ptr(), Ho, Wo, Cout, 0, 0, None, 0, 0, 0, 0, 0,
                  wp.data_ptr(), None, None, FLAG_ROUND_TF32,
                  dz.data_ptr(), Ho, Wo, Cin, 0, 0, None, 0, 0, 0, 0, 0, None, None, None, B, Ho, Wo, Cin, MATH_TF32, wfmt, _s())
        dx = torch.empty_like(x)
        _lib.call("pu_zero_insert2x_bwd", dz.data_ptr(), dx.data_ptr(), B, H, W, Cin, Ho, Wo, oy, ox, _s())
    if need_dw:
        def wgrad():
            z = torch.empty((B, Ho, Wo, Cin), device=dev, dtype=torch.float32)  # recomputed: cheaper than keeping it alive
            _lib.call("pu_zero_insert2x_fwd", x.data_ptr(), z.data_ptr(), B, H, W, Cin, Ho, Wo, oy, ox, _s())
            dwc = torch.empty((Cout, Cin, 3, 3), device=dev, dtype=torch.float32)  # gradient of the equivalent conv's OIHW weight
            _lib.call("pu_conv3x3_wgrad", z.data_ptr(), Ho, Wo, Cin, 0, 0, None, 0, 0, 0, 0, 0,
                      g.data_ptr(), dwc.data_ptr(), None, B, Ho, Wo, Cout, MATH_TF32, _s())
            return dwc.flip(2, 3).permute(1, 0, 2, 3).contiguous()  # w_conv[co][ci][k] = w[ci][co][2-k]
        side = _side_stream()
        if side is not None:  # off the critical path: see conv3x3_bwd
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dw = wgrad()
            x.record_stream(side)
            g.record_stream(side)
            dw.record_stream(main)
            _note_side(dw)
        else:
            dw = wgrad()
    return [dx, dw, db]


@convT3x3s2_tc_bwd.register_fake
def _(dy, x, weight, has_bias, need_dx, need_dw, Ho, Wo, oy, ox):
    e = x.new_empty(0)
    return [torch.empty_like(x) if need_dx else e, torch.empty_like(weight) if need_dw else e,
            x.new_empty(weight.shape[1]) if has_bias else e]


def _convT3tc_setup(ctx, inputs, output):
    x, weight, bias, Ho, Wo, oy, ox = inputs
    ctx.save_for_backward(x, weight)
    ctx.cfg = (bias is not None, Ho, Wo, oy, ox)


def _convT3tc_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    has_bias, Ho, Wo, oy, ox = ctx.cfg
    need = ctx.needs_input_grad
    dx, dw, db = convT3x3s2_tc_bwd(dy.contiguous(), x, weight, has_bias and need[2], need[0], need[1], Ho, Wo, oy, ox)
    return dx if need[0] else None, dw if need[1] else None, db if (has_bias and need[2]) else None, None, None, None, None


convT3x3s2_tc.register_autograd(_convT3tc_backward, setup_context=_convT3tc_setup)


@torch.library.custom_op("pu::convT3x3s2", mutates_args=())
def convT3x3s2(x: Tensor, weight: Tensor, bias: Optional[Tensor], chan_scale: Optional[Tensor],
               Ho: int, Wo: int, oy: int, ox: int, round_out: bool = False) -> Tensor:
    _chk(x, weight, bias, chan_scale)
    B, H, W, Cin = x.shape
    Cout = weight.shape[1]
    y = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
    _lib.call("pu_convT3x3s2_fwd", x.data_ptr(), weight.data_ptr(), _p(bias), _p(chan_scale), y.data_ptr(),
              B, H, W, Cin, Cout, Ho, Wo, oy, ox, FLAG_ROUND_TF32 if round_out else 0, _s())
    return y


@convT3x3s2.register_fake
def _(x, weight, bias, chan_scale, Ho, Wo, oy, ox, round_out=False):
    return x.new_empty((x.shape[0], Ho, Wo, weight.shape[1]))


@torch.library.custom_op("pu::convT3x3s2_bwd", mutates_args=())
def convT3x3s2_bwd(dy: Tensor, x: Tensor, weight: Tensor, chan_scale: Optional[Tensor], oy: int, ox: int,
                   need_dx: bool, need_dw: bool, need_db: bool) -> List[Tensor]:
    _chk(dy, x, weight, chan_scale)
    B, H, W, Cin = x.shape
    Cout = weight.shape[1]
    Ho, Wo = dy.shape[1], dy.shape[2]
    dx = torch.empty_like(x) if need_dx else _e(x.device)
    dw = torch.empty_like(weight) if need_dw else _e(x.device)
    db = torch.empty(Cout, device=x.device, dtype=torch.float32) if need_db else _e(x.device)
    _lib.call("pu_convT3x3s2_bwd", x.data_ptr(), weight.data_ptr(), dy.data_ptr(), _p(chan_scale),
              dx.data_ptr() if need_dx else None, dw.data_ptr() if need_dw else None, db.data_ptr() if need_db else None,
              B, H, W, Cin, Cout, Ho, Wo, oy, ox, _s())
    return [dx, dw, db]


@convT3x3s2_bwd.register_fake
def _(dy, x, weight, chan_scale, oy, ox, need_dx, need_dw, need_db):
    e = x.new_empty(0)
    return [torch.empty_like(x) if need_dx else e, torch.empty_like(weight) if need_dw else e,
            x.new_empty(weight.shape[1]) if need_db else e]


def _convT3_setup(ctx, inputs, output):
    x, weight, bias, chan_scale, Ho, Wo, oy, ox, _round = inputs
    ctx.save_for_backward(x, weight, chan_scale)
    ctx.cfg = (bias is not None, oy, ox)


def _convT3_backward(ctx, dy):
    x, weight, chan_scale = ctx.saved_tensors
    has_bias, oy, ox = ctx.cfg
    need = ctx.needs_input_grad
    dx, dw, db = convT3x3s2_bwd(dy.contiguous(), x, weight, chan_scale, oy, ox, need[0], need[1], has_bias and need[2])
    return dx if need[0] else None, dw if need[1] else None, db if (has_bias and need[2]) else None, None, None, None, None, None, None


convT3x3s2.register_autograd(_convT3_backward, setup_context=_convT3_setup)


# =================================================================================================
# pooling / resampling
# =================================================================================================
@torch.library.custom_op("pu::maxpool2", mutates_args=())
def maxpool2(x: Tensor, chan_scale: Optional[Tensor], mask_in: bool = False) -> Tensor:
    _chk(x, chan_scale)
    B, H, W, C = x.shape
    y = torch.empty((B, H // 2, W // 2, C), device=x.device, dtype=torch.float32)
    _lib.call("pu_maxpool2_fwd", x.data_ptr(), _p(chan_scale), y.data_ptr(), B, H, W, C, _s())
    return y


@maxpool2.register_fake
def _(x, chan_scale, mask_in=False):
    return x.new_empty((x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3]))


@torch.library.custom_op("pu::maxpool2_bwd", mutates_args=())
def maxpool2_bwd(dy: Tensor, x: Tensor, chan_scale: Optional[Tensor], mask_in: bool = False, acc: Optional[Tensor] = None) -> Tensor:
    """acc: a second gradient of x (the skip connection's) summed in the same pass: dx = route(dy) + acc."""
    _chk(dy, x, chan_scale, acc)
    B, H, W, C = x.shape
    if acc is not None and acc.shape != x.shape:
        raise RuntimeError("maxpool2_bwd: acc must have the shape of x")
    dx = torch.empty_like(x)
    _lib.call("pu_maxpool2_bwd", x.data_ptr(), _p(chan_scale), dy.data_ptr(), _p(acc), dx.data_ptr(), B, H, W, C,
              FLAG_MASK_IN if mask_in else 0, _s())
    return dx


@maxpool2_bwd.register_fake
def _(dy, x, chan_scale, mask_in=False, acc=None):
    return torch.empty_like(x)


@torch.library.custom_op("pu::maxpool2_code", mutates_args=())
def maxpool2_code(x: Tensor) -> Tuple[Tensor, Tensor]:
    """-> (maxpool2(x), code): code [B, H/2, W/2, C] uint8 records where each maximum sits (bits 0-1) and whether it is > 0
    (bit 2), so that the backward pass does not re-read x."""
    _chk(x)
    B, H, W, C = x.shape
    y = torch.empty((B, H // 2, W // 2, C), device=x.device, dtype=torch.float32)
    code = torch.empty((B, H // 2, W // 2, C), device=x.device, dtype=torch.uint8)
    _lib.call("pu_maxpool2_fwd_code", x.data_ptr(), None, y.data_ptr(), code.data_ptr(), B, H, W, C, _s())
    return y, code


@maxpool2_code.register_fake
def _(x):
    shp = (x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3])
    return x.new_empty(shp), x.new_empty(shp, dtype=torch.uint8)


@torch.library.custom_op("pu::maxpool2_bwd_code", mutates_args=())
def maxpool2_bwd_code(dy: Tensor, code: Tensor, H: int, W: int, mask_in: bool = False, acc: Optional[Tensor] = None) -> Tensor:
    """maxpool2_bwd from the arg-max code of maxpool2_code (x itself is not read); H, W: the size of x."""
    _chk(dy, acc)
    B, Ho, Wo, C = dy.shape
    if not code.is_cuda or code.dtype != torch.uint8 or not code.is_contiguous() or tuple(code.shape) != (B, Ho, Wo, C):
        raise RuntimeError("maxpool2_bwd_code: code must be the contiguous uint8 CUDA tensor maxpool2_code returned")
    if (H // 2, W // 2) != (Ho, Wo) or (acc is not None and tuple(acc.shape) != (B, H, W, C)):
        raise RuntimeError("maxpool2_bwd_code: shape mismatch")
    dx = torch.empty((B, H, W, C), device=dy.device, dtype=torch.float32)
    _lib.call("pu_maxpool2_bwd_code", code.data_ptr(), None, dy.data_ptr(), _p(acc), dx.data_ptr(), B, H, W, C,
              FLAG_MASK_IN if mask_in else 0, _s())
    return dx


@maxpool2_bwd_code.register_fake
def _(dy, code, H, W, mask_in=False, acc=None):
    return dy.new_empty((dy.shape[0], H, W, dy.shape[3]))


POOL_CODE = os.environ.get("PU_POOL_CODE", "1") == "1"


class _PoolSkip(torch.autograd.Function):
    """x -> (maxpool2(x), x): the pooled tensor for the next encoder level and x itself as the skip connection
    (unet_p.py:59-66).  Both gradients of x then reach ONE backward call, and the skip gradient is accumulated inside the
    pooling backward kernel instead of by a separate elementwise add over the whole tensor."""

    @staticmethod
    def forward(ctx, x, mask_in):
        ctx.mask_in = mask_in
        ctx.set_materialize_grads(False)
        if POOL_CODE and x.requires_grad:
            # the backward pass routes from one code byte per pooled element instead of re-reading x (33 of its 107 MB at 128x128)
            y, code = maxpool2_code(x)
            ctx.save_for_backward(code)
            ctx.hw = (x.shape[1], x.shape[2])
            return y, x.view_as(x)
        ctx.save_for_backward(x)
        ctx.hw = None
        return maxpool2(x, None, mask_in), x.view_as(x)

    @staticmethod
    def backward(ctx, gy, gskip):
        (x,) = ctx.saved_tensors
        if gy is None:
            return gskip, None
        acc = None if gskip is None else gskip.contiguous()
        if ctx.hw is not None:
            return maxpool2_bwd_code(gy.contiguous(), x, ctx.hw[0], ctx.hw[1], ctx.mask_in, acc), None
        return maxpool2_bwd(gy.contiguous(), x, None, ctx.mask_in, acc), None


def pool_skip(x: Tensor, mask_in: bool = False) -> Tuple[Tensor, Tensor]:
    return _PoolSkip.apply(x, mask_in)


def _pool_setup(ctx, inputs, output):
    x, chan_scale, mask_in = inputs
    ctx.save_for_backward(x, chan_scale)
    ctx.mask_in = mask_in


def _pool_backward(ctx, dy):
    x, chan_scale = ctx.saved_tensors
    return maxpool2_bwd(dy.contiguous(), x, chan_scale, ctx.mask_in, None), None, None


maxpool2.register_autograd(_pool_backward, setup_context=_pool_setup)


@torch.library.custom_op("pu::bilinear2x", mutates_args=())
def bilinear2x(x: Tensor) -> Tensor:
    _chk(x)
    B, H, W, C = x.shape
    y = torch.empty((B, 2 * H, 2 * W, C), device=x.device, dtype=torch.float32)
    _lib.call("pu_bilinear2x_fwd", x.data_ptr(), y.data_ptr(), B, H, W, C, _s())
    return y


@bilinear2x.register_fake
def _(x):
    return x.new_empty((x.shape[0], 2 * x.shape[1], 2 * x.shape[2], x.shape[3]))


@torch.library.custom_op("pu::bilinear2x_bwd", mutates_args=())
def bilinear2x_bwd(dy: Tensor) -> Tensor:
    _chk(dy)
    B, Ho, Wo, C = dy.shape
    dx = torch.empty((B, Ho // 2, Wo // 2, C), device=dy.device, dtype=torch.float32)
    _lib.call("pu_bilinear2x_bwd", dy.data_ptr(), dx.data_ptr(), B, Ho // 2, Wo // 2, C, _s())
    return dx


@bilinear2x_bwd.register_fake
def _(dy):
    return dy.new_empty((dy.shape[0], dy.shape[1] // 2, dy.shape[2] // 2, dy.shape[3]))


bilinear2x.register_autograd(lambda ctx, dy: bilinear2x_bwd(dy.contiguous()), setup_context=lambda ctx, inputs, output: None)


# =================================================================================================
# concat + crop + channel scale (materialised only in Dropout2d training mode)
# =================================================================================================
@torch.library.custom_op("pu::concat_scale", mutates_args=())
def concat_scale(x0: Tensor, x1: Tensor, chan_scale: Optional[Tensor], H: int, W: int,
                 oy0: int, ox0: int, oy1: int, ox1: int, round_out: bool = False) -> Tensor:
    _chk(x0, x1, chan_scale)
    B = x0.shape[0]
    H0, W0, C0 = _dims(x0)
    H1, W1, C1 = _dims(x1)
    y = torch.empty((B, H, W, C0 + C1), device=x0.device, dtype=torch.float32)
    _lib.call("pu_concat_scale_fwd", x0.data_ptr(), H0, W0, C0, oy0, ox0, x1.data_ptr(), H1, W1, C1, oy1, ox1,
              _p(chan_scale), y.data_ptr(), B, H, W, FLAG_ROUND_TF32 if round_out else 0, _s())
    return y


@concat_scale.register_fake
def _(x0, x1, chan_scale, H, W, oy0, ox0, oy1, ox1, round_out=False):
    return x0.new_empty((x0.shape[0], H, W, x0.shape[3] + x1.shape[3]))


@torch.library.custom_op("pu::concat_scale_bwd", mutates_args=())
def concat_scale_bwd(dy: Tensor, chan_scale: Optional[Tensor], shape0: List[int], shape1: List[int],
                     oy0: int, ox0: int, oy1: int, ox1: int) -> List[Tensor]:
    _chk(dy, chan_scale)
    B, H, W, _ = dy.shape
    dev = dy.device
    _, H0, W0, C0 = shape0
    _, H1, W1, C1 = shape1
    dx0 = (torch.empty if (H0 == H and W0 == W) else torch.zeros)(shape0, device=dev, dtype=torch.float32)
    dx1 = (torch.empty if (H1 == H and W1 == W) else torch.zeros)(shape1, device=dev, dtype=torch.float32)
    _lib.call("pu_concat_scale_bwd", dy.data_ptr(), _p(chan_scale), dx0.data_ptr(), H0, W0, C0, oy0, ox0,
              dx1.data_ptr(), H1, W1, C1, oy1, ox1, B, H, W, _s())
    return [dx0, dx1]


@concat_scale_bwd.register_fake
def _(dy, chan_scale, shape0, shape1, oy0, ox0, oy1, ox1):
    return [dy.new_empty(shape0), dy.new_empty(shape1)]


def _cat_setup(ctx, inputs, output):
    x0, x1, chan_scale, H, W, oy0, ox0, oy1, ox1, _round = inputs
    ctx.save_for_backward(chan_scale)
    ctx.cfg = (list(x0.shape), list(x1.shape), oy0, ox0, oy1, ox1)


def _cat_backward(ctx, dy):
    (chan_scale,) = ctx.saved_tensors
    s0, s1, oy0, ox0, oy1, ox1 = ctx.cfg
    dx0, dx1 = concat_scale_bwd(dy.contiguous(), chan_scale, s0, s1, oy0, ox0, oy1, ox1)
    return dx0, dx1, None, None, None, None, None, None, None, None


concat_scale.register_autograd(_cat_backward, setup_context=_cat_setup)


# =================================================================================================
# BatchNorm2d (optional)
# =================================================================================================
@torch.library.custom_op("pu::batchnorm", mutates_args=())
def batchnorm(x: Tensor, gamma: Tensor, beta: Tensor, running_mean: Tensor, running_var: Tensor,
              train: bool, momentum: float, eps: float, relu: bool, round_out: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """-> [y, mean, invstd] (the statistics the backward needs)."""
    _chk(x, gamma, beta, running_mean, running_var)
    C = x.shape[-1]
    npix = x.numel() // C
    y = torch.empty_like(x)
    mean = torch.empty(C, device=x.device, dtype=torch.float32)
    invstd = torch.empty(C, device=x.device, dtype=torch.float32)
    fl = (FLAG_RELU if relu else 0) | (FLAG_ROUND_TF32 if round_out else 0)
    if train:
        ws = torch.empty(2 * C, device=x.device, dtype=torch.float64)
        _lib.call("pu_bn_train_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(),
                  invstd.data_ptr(), None, None, ws.data_ptr(),
                  float(momentum), float(eps), npix, C, fl, _s())
    else:
        _lib.call("pu_bn_eval_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
                  running_var.data_ptr(), y.data_ptr(), float(eps), npix, C, fl, _s())
        mean.copy_(running_mean)
        _lib.call("pu_bn_invstd", running_var.data_ptr(), invstd.data_ptr(), float(eps), C, _s())
    return y, mean, invstd


@torch.library.custom_op("pu::bn_update_running", mutates_args=("running_mean", "running_var"))
def bn_update_running(mean: Tensor, invstd: Tensor, running_mean: Tensor, running_var: Tensor,
                      momentum: float, eps: float, npix: int) -> None:
    """Running-statistics side effect of a training-mode BatchNorm (kept out of the functional op)."""
    _chk(mean, invstd, running_mean, running_var)
    _lib.call("pu_bn_update_running", mean.data_ptr(), invstd.data_ptr(), running_mean.data_ptr(), running_var.data_ptr(),
              float(momentum), float(eps), npix, mean.shape[0], _s())


@batchnorm.register_fake
def _(x, gamma, beta, running_mean, running_var, train, momentum, eps, relu, round_out=False):
    C = x.shape[-1]
    return torch.empty_like(x), x.new_empty(C), x.new_empty(C)


@torch.library.custom_op("pu::batchnorm_bwd", mutates_args=())
def batchnorm_bwd(dy: Tensor, x: Tensor, y: Tensor, gamma: Tensor, mean: Tensor, invstd: Tensor,
                  relu: bool, train: bool) -> List[Tensor]:
    _chk(dy, x, y, gamma, mean, invstd)
    C = x.shape[-1]
    npix = x.numel() // C
    dx = torch.empty_like(x)
    dgamma = torch.empty(C, device=x.device, dtype=torch.float32)
    dbeta = torch.empty(C, device=x.device, dtype=torch.float32)
    ws = torch.empty(2 * C, device=x.device, dtype=torch.float64)
    _lib.call("pu_bn_bwd", x.data_ptr(), y.data_ptr(), dy.data_ptr(), gamma.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
              dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), npix, C, int(relu), int(train), _s())
    return [dx, dgamma, dbeta]


@batchnorm_bwd.register_fake
def _(dy, x, y, gamma, mean, invstd, relu, train):
    C = x.shape[-1]
    return [torch.empty_like(x), x.new_empty(C), x.new_empty(C)]


def _bn_setup(ctx, inputs, output):
    x, gamma, beta, running_mean, running_var, train, momentum, eps, relu, _round = inputs
    y, mean, invstd = output
    ctx.save_for_backward(x, y, gamma, mean, invstd)
    ctx.cfg = (relu, train)


def _bn_backward(ctx, dy, _dmean, _dinvstd):
    x, y, gamma, mean, invstd = ctx.saved_tensors
    relu, train = ctx.cfg
    dx, dgamma, dbeta = batchnorm_bwd(dy.contiguous(), x, y, gamma, mean, invstd, relu, train)
    return dx, dgamma, dbeta, None, None, None, None, None, None, None


batchnorm.register_autograd(_bn_backward, setup_context=_bn_setup)


@torch.library.custom_op("pu::bn_fold_conv", mutates_args=())
def bn_fold_conv(weight: Tensor, bias: Optional[Tensor], gamma: Tensor, beta: Tensor, running_mean: Tensor, running_var: Tensor,
                 eps: float) -> Tuple[Tensor, Tensor]:
    """Eval-mode BatchNorm folded into the conv in front of it -> (w', b') (forward only: no autograd formula)."""
    _chk(weight, bias, gamma, beta, running_mean, running_var)
    Cout = weight.shape[0]
    w2 = torch.empty_like(weight)
    b2 = torch.empty(Cout, device=weight.device, dtype=torch.float32)
    _lib.call("pu_bn_fold_conv", weight.data_ptr(), _p(bias), gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
              running_var.data_ptr(), float(eps), w2.data_ptr(), b2.data_ptr(), Cout, weight.numel() // Cout, _s())
    return w2, b2


@bn_fold_conv.register_fake
def _(weight, bias, gamma, beta, running_mean, running_var, eps):
    return torch.empty_like(weight), weight.new_empty(weight.shape[0])


# =================================================================================================
# layout
# =================================================================================================
@torch.library.custom_op("pu::nchw_to_nhwc", mutates_args=())
def nchw_to_nhwc(x: Tensor) -> Tensor:
    _chk(x)
    B, C, H, W = x.shape
    y = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
    _lib.call("pu_nchw_to_nhwc", x.data_ptr(), y.data_ptr(), B, C, H, W, _s())
    return y


@nchw_to_nhwc.register_fake
def _(x):
    B, C, H, W = x.shape
    return x.new_empty((B, H, W, C))


@torch.library.custom_op("pu::nhwc_to_nchw", mutates_args=())
def nhwc_to_nchw(x: Tensor) -> Tensor:
    _chk(x)
    B, H, W, C = x.shape
    y = torch.empty((B, C, H, W), device=x.device, dtype=torch.float32)
    _lib.call("pu_nhwc_to_nchw", x.data_ptr(), y.data_ptr(), B, C, H, W, _s())
    return y


@nhwc_to_nchw.register_fake
def _(x):
    B, H, W, C = x.shape
    return x.new_empty((B, C, H, W))


nchw_to_nhwc.register_autograd(lambda ctx, dy: nhwc_to_nchw(dy.contiguous()), setup_context=lambda ctx, inputs, output: None)
nhwc_to_nchw.register_autograd(lambda ctx, dy: nchw_to_nhwc(dy.contiguous()), setup_context=lambda ctx, inputs, output: None)


# =================================================================================================
# plastic head + trace
# =================================================================================================
@torch.library.custom_op("pu::plastic_head", mutates_args=())
def plastic_head(X: Tensor, w: Tensor, alpha: Tensor, hebb: Tensor) -> Tuple[Tensor, Tensor]:
    """X [B*N, N] -> [S = sigmoid(X @ (w + alpha*hebb)) [B*N, N], Weff [N, N]]   (unet_p.py:70-79)"""
    _chk(X, w, alpha, hebb)
    N = w.shape[0]
    B = X.shape[0] // N
    S = torch.empty_like(X)
    weff = torch.empty_like(w)
    _lib.call("pu_plastic_head_fwd", X.data_ptr(), w.data_ptr(), alpha.data_ptr(), hebb.data_ptr(), weff.data_ptr(),
              S.data_ptr(), B, N, _s())
    return S, weff


@plastic_head.register_fake
def _(X, w, alpha, hebb):
    return torch.empty_like(X), torch.empty_like(w)


@torch.library.custom_op("pu::plastic_head_bwd", mutates_args=())
def plastic_head_bwd(gS: Tensor, X: Tensor, S: Tensor, weff: Tensor, alpha: Tensor, hebb: Tensor,
                     need_gx: bool, need_galpha: bool, need_ghebb: bool) -> List[Tensor]:
    _chk(gS, X, S, weff, alpha, hebb)
    N = weff.shape[0]
    B = X.shape[0] // N
    gA = torch.empty_like(X)
    gX = torch.empty_like(X) if need_gx else _e(X.device)
    gw = torch.empty_like(weff)
    galpha = torch.empty_like(weff) if need_galpha else _e(X.device)
    ghebb = torch.empty_like(weff) if need_ghebb else _e(X.device)
    side = _side_stream()
    if side is not None:
        # gA and gX on the critical path; the parameter gradients (gw, galpha, ghebb) on the side stream
        _lib.call("pu_plastic_head_bwd", X.data_ptr(), S.data_ptr(), gS.data_ptr(), weff.data_ptr(), alpha.data_ptr(), hebb.data_ptr(),
                  gA.data_ptr(), gX.data_ptr() if need_gx else None, None, None, None, B, N, _s())
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _lib.call("pu_plastic_head_bwd", X.data_ptr(), S.data_ptr(), None, weff.data_ptr(), alpha.data_ptr(), hebb.data_ptr(),
                      gA.data_ptr(), None, gw.data_ptr(), galpha.data_ptr() if need_galpha else None,
                      ghebb.data_ptr() if need_ghebb else None, B, N, _s())
        for t in (X, S, weff, gA, alpha, hebb, gw, galpha if need_galpha else None, ghebb if need_ghebb else None):
            if t is not None:
                t.record_stream(side)
        _note_side(gw, galpha if need_galpha else None)
    else:
        _lib.call("pu_plastic_head_bwd", X.data_ptr(), S.data_ptr(), gS.data_ptr(), weff.data_ptr(), alpha.data_ptr(), hebb.data_ptr(),
                  gA.data_ptr(), gX.data_ptr() if need_gx else None, gw.data_ptr(), galpha.data_ptr() if need_galpha else None,
                  ghebb.data_ptr() if need_ghebb else None, B, N, _s())
    return [gX, gw, galpha, ghebb]


@plastic_head_bwd.register_fake
def _(gS, X, S, weff, alpha, hebb, need_gx, need_galpha, need_ghebb):
    e = X.new_empty(0)
    return [torch.empty_like(X) if need_gx else e, torch.empty_like(weff),
            torch.empty_like(weff) if need_galpha else e, torch.empty_like(weff) if need_ghebb else e]


def _head_setup(ctx, inputs, output):
    X, w, alpha, hebb = inputs
    S, weff = output
    ctx.save_for_backward(X, S, weff, alpha, hebb)


def _head_backward(ctx, gS, _gweff):
    X, S, weff, alpha, hebb = ctx.saved_tensors
    need = ctx.needs_input_grad
    gX, gw, galpha, ghebb = plastic_head_bwd(gS.contiguous(), X, S, weff, alpha, hebb, need[0], need[2], need[3])
    return gX if need[0] else None, gw if need[1] else None, galpha if need[2] else None, ghebb if need[3] else None


plastic_head.register_autograd(_head_backward, setup_context=_head_setup)


# ---- training-step form of the head: forward + BCE loss + backward of both in one launch (TrainStep, TF32 mode) ------------
# UNIT_GRAD: the gradient tensor TrainStep seeds `loss.backward()` with.  The fused kernel has already folded dloss/dloss = 1
# into gA and gX; any other seed (a caller scaling the loss) is applied with plain tensor multiplies.
UNIT_GRAD = None
_HEAD_SCRATCH = {}
_HEAD_SCRATCH_BLOCKS = 4096
# terms of the head's parameter-gradient GEMM: 1 = plain TF32 (the precision of every conv weight gradient of the TF32 mode),
# 3 = error-compensated 3xTF32 (fp32 level)
HEAD_WGRAD_TERMS = int(os.environ.get("PU_HEAD_WGRAD_TERMS", "1"))


@torch.library.custom_op("pu::head_weff", mutates_args=())
def head_weff(w: Tensor, alpha: Tensor, hebb: Tensor) -> Tensor:
    """Weff = w + alpha*hebb (unet_p.py:73-76), for plastic_head_bce(weff=...): TrainStep computes it at the start of the step."""
    _chk(w, alpha, hebb)
    out = torch.empty_like(w)
    _lib.call("pu_head_weff", w.data_ptr(), alpha.data_ptr(), hebb.data_ptr(), out.data_ptr(), w.shape[0], _s())
    return out


@head_weff.register_fake
def _(w, alpha, hebb):
    return torch.empty_like(w)


@torch.library.custom_op("pu::plastic_head_bce", mutates_args=())
def plastic_head_bce(X: Tensor, w: Tensor, alpha: Tensor, hebb: Tensor, target: Tensor,
                     need_gx: bool, weff: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (S, loss, gA, gX): S = sigmoid(X @ (w + alpha*hebb)) (unet_p.py:70-79), loss = nn.BCELoss()(S, target) (train.py:100-103),
    gA = dloss/dlogits, gX = dloss/dX — one launch, 3xTF32 tensor-core GEMMs (fp32-level logits)."""
    _chk(X, w, alpha, hebb, target, weff)
    N = w.shape[0]
    B = X.shape[0] // N
    if weff is not None and tuple(weff.shape) != (N, N):
        raise RuntimeError("plastic_head_bce: weff must be [%d, %d]" % (N, N))
    if target.numel() != X.numel():
        raise RuntimeError("plastic_head_bce: target has %d elements, the output %d" % (target.numel(), X.numel()))
    S = torch.empty_like(X)
    gA = torch.empty_like(X)
    gX = torch.empty_like(X) if need_gx else _e(X.device)
    loss = torch.empty(1, device=X.device, dtype=torch.float32)
    nblk = (B * N + 63) // 64
    scratch = None
    if nblk <= _HEAD_SCRATCH_BLOCKS:
        # ticket counter + per-CTA partial sums of the loss: zero once, every launch leaves it zeroed (one fused head at a time
        # per device: launches that share it must be stream-ordered, as the steps of a TrainStep are)
        scratch = _HEAD_SCRATCH.get(X.device.index)
        if scratch is None:
            scratch = _HEAD_SCRATCH[X.device.index] = torch.zeros(1 + _HEAD_SCRATCH_BLOCKS, device=X.device, dtype=torch.float32)
    _lib.call("pu_plastic_head_bce", X.data_ptr(), w.data_ptr(), alpha.data_ptr(), hebb.data_ptr(), _p(weff), target.data_ptr(),
              S.data_ptr(), loss.data_ptr(), gA.data_ptr(), gX.data_ptr() if need_gx else None, _p(scratch), B, N, _s())
    return S, loss, gA, gX


@plastic_head_bce.register_fake
def _(X, w, alpha, hebb, target, need_gx, weff=None):
    return torch.empty_like(X), X.new_empty(1), torch.empty_like(X), torch.empty_like(X) if need_gx else X.new_empty(0)


@torch.library.custom_op("pu::plastic_head_wgrad", mutates_args=())
def plastic_head_wgrad(X: Tensor, gA: Tensor, alpha: Tensor, hebb: Tensor, need_galpha: bool, need_ghebb: bool) -> List[Tensor]:
    """-> [gw, galpha, ghebb] from gA (on a weight-gradient side stream when TrainStep provides one)."""
    _chk(X, gA, alpha, hebb)
    N = alpha.shape[0]
    B = X.shape[0] // N
    gw = torch.empty_like(alpha)
    galpha = torch.empty_like(alpha) if need_galpha else _e(X.device)
    ghebb = torch.empty_like(alpha) if need_ghebb else _e(X.device)

    def run():
        _lib.call("pu_plastic_head_wgrad_tc", X.data_ptr(), gA.data_ptr(), alpha.data_ptr(), hebb.data_ptr(), gw.data_ptr(),
                  galpha.data_ptr() if need_galpha else None, ghebb.data_ptr() if need_ghebb else None, B, N, HEAD_WGRAD_TERMS, _s())

    side = _side_stream()
    if side is not None:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        for t in (X, gA, alpha, hebb, gw, galpha if need_galpha else None, ghebb if need_ghebb else None):
            if t is not None:
                t.record_stream(side)
        _note_side(gw, galpha if need_galpha else None)
    else:
        run()
    return [gw, galpha, ghebb]


@plastic_head_wgrad.register_fake
def _(X, gA, alpha, hebb, need_galpha, need_ghebb):
    e = X.new_empty(0)
    return [torch.empty_like(alpha), torch.empty_like(alpha) if need_galpha else e, torch.empty_like(alpha) if need_ghebb else e]


def _head_bce_setup(ctx, inputs, output):
    X, w, alpha, hebb, target, need_gx, _weff = inputs
    S, loss, gA, gX = output
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(X, alpha, hebb, gA, gX)
    ctx.need_gx = need_gx


def _head_bce_backward(ctx, gS, gloss, _ga, _gx):
    X, alpha, hebb, gA, gX = ctx.saved_tensors
    need = ctx.needs_input_grad
    if gS is not None:
        raise RuntimeError("plastic_head_bce: the fused head owns the loss; a gradient through its sigmoid output is not supported "
                           "(detach it, as train.py:99 does with the trace)")
    if gloss is None:
        return None, None, None, None, None, None, None
    if need[0] and not ctx.need_gx:
        raise RuntimeError("plastic_head_bce: built with need_gx=False but X requires a gradient")
    if UNIT_GRAD is None or gloss.data_ptr() != UNIT_GRAD.data_ptr():
        gA = gA * gloss
        gX = gX * gloss if need[0] else gX
    gw = galpha = ghebb = None
    if need[1] or need[2] or need[3]:
        gw, galpha, ghebb = plastic_head_wgrad(X, gA, alpha, hebb, need[2], need[3])
    return (gX if need[0] else None, gw if need[1] else None, galpha if need[2] else None, ghebb if need[3] else None, None, None, None)


plastic_head_bce.register_autograd(_head_bce_backward, setup_context=_head_bce_setup)


@torch.library.custom_op("pu::trace_update", mutates_args=())
def trace_update(hebb: Tensor, pre: Tensor, post: Tensor, eta: Tensor, rule: int, ld: int, K: int) -> Tensor:
    """Fused Hebb/Oja update over K (pre, post) row pairs, rows k at pre[k*ld : k*ld+N]  (unet_p.py:81-84)."""
    _chk(hebb, pre, post, eta)
    N = hebb.shape[0]
    if pre.numel() < (K - 1) * ld + N or post.numel() < (K - 1) * ld + N:
        raise RuntimeError("trace_update: pre/post too small for (K, ld)")
    out = torch.empty_like(hebb)
    _lib.call("pu_trace_update_fwd", hebb.data_ptr(), pre.data_ptr(), post.data_ptr(), ld, K, eta.data_ptr(), rule,
              out.data_ptr(), N, _s())
    return out


@trace_update.register_fake
def _(hebb, pre, post, eta, rule, ld, K):
    return torch.empty_like(hebb)


@torch.library.custom_op("pu::trace_update_bwd", mutates_args=())
def trace_update_bwd(gout: Tensor, hebb: Tensor, pre: Tensor, post: Tensor, eta: Tensor, rule: int, ld: int, K: int) -> List[Tensor]:
    """-> [ghebb [N,N], gpre (shape of pre, zero outside the K rows), gpost (same), geta [1]]"""
    _chk(gout, hebb, pre, post, eta)
    N = hebb.shape[0]
    dev = hebb.device
    ghebb = torch.empty_like(hebb)
    geta = torch.empty(1, device=dev, dtype=torch.float32)
    gpre_rows = torch.empty((K, N), device=dev, dtype=torch.float32)
    gpost_rows = torch.empty((K, N), device=dev, dtype=torch.float32)
    _lib.call("pu_trace_update_bwd", hebb.data_ptr(), pre.data_ptr(), post.data_ptr(), ld, K, eta.data_ptr(), rule,
              gout.data_ptr(), ghebb.data_ptr(), gpre_rows.data_ptr(), gpost_rows.data_ptr(), geta.data_ptr(), N, _s())
    return [ghebb, gpre_rows, gpost_rows, geta]


@trace_update_bwd.register_fake
def _(gout, hebb, pre, post, eta, rule, ld, K):
    N = hebb.shape[0]
    return [torch.empty_like(hebb), hebb.new_empty((K, N)), hebb.new_empty((K, N)), hebb.new_empty(1)]


def _trace_setup(ctx, inputs, output):
    hebb, pre, post, eta, rule, ld, K = inputs
    ctx.save_for_backward(hebb, pre, post, eta)
    ctx.cfg = (rule, ld, K)


def _scatter_rows(rows: Tensor, like: Tensor, ld: int, K: int) -> Tensor:
    """Place the K gradient rows at stride ld inside a zero tensor shaped like the flat pre/post buffer."""
    N = rows.shape[1]
    flat = torch.zeros(like.numel(), device=like.device, dtype=like.dtype)
    if ld == N:
        flat[: K * N] = rows.reshape(-1)
    else:
        flat.as_strided((K, N), (ld, 1)).copy_(rows)
    return flat.view(like.shape)


def _trace_backward(ctx, gout):
    hebb, pre, post, eta = ctx.saved_tensors
    rule, ld, K = ctx.cfg
    need = ctx.needs_input_grad
    ghebb, gpre_rows, gpost_rows, geta = trace_update_bwd(gout.contiguous(), hebb, pre, post, eta, rule, ld, K)
    gpre = _scatter_rows(gpre_rows, pre, ld, K) if need[1] else None
    gpost = _scatter_rows(gpost_rows, post, ld, K) if need[2] else None
    return (ghebb if need[0] else None, gpre, gpost, geta.view(eta.shape) if need[3] else None, None, None, None)


trace_update.register_autograd(_trace_backward, setup_context=_trace_setup)


# ---- data-parallel split form (no autograd: the trace is detached between steps, train.py:99) ----
@torch.library.custom_op("pu::trace_delta", mutates_args=())
def trace_delta(pre: Tensor, post: Tensor, N: int, ld: int, K: int) -> Tensor:
    """-> [N*N + N] = (sum_k outer(pre_k, post_k), sum_k post_k^2): the all-reduce payload."""
    _chk(pre, post)
    out = torch.empty(N * N + N, device=pre.device, dtype=torch.float32)
    _lib.call("pu_trace_delta", pre.data_ptr(), post.data_ptr(), ld, K, out.data_ptr(), N, _s())
    return out


@trace_delta.register_fake
def _(pre, post, N, ld, K):
    return pre.new_empty(N * N + N)


@torch.library.custom_op("pu::trace_delta_tc", mutates_args=())
def trace_delta_tc(pre: Tensor, post: Tensor, N: int, ld: int, K: int) -> Tensor:
    """trace_delta on the tensor cores (3xTF32 mma.sync split-K GEMM) for large K (rows='all': K = B*N)."""
    _chk(pre, post)
    if pre.numel() < (K - 1) * ld + N or post.numel() < (K - 1) * ld + N:
        raise RuntimeError("trace_delta_tc: pre/post too small for (K, ld)")
    out = torch.empty(N * N + N, device=pre.device, dtype=torch.float32)
    _lib.call("pu_trace_delta_tc", pre.data_ptr(), post.data_ptr(), ld, K, out.data_ptr(), N, _s())
    return out


@trace_delta_tc.register_fake
def _(pre, post, N, ld, K):
    return pre.new_empty(N * N + N)


@torch.library.custom_op("pu::trace_apply", mutates_args=())
def trace_apply(hebb: Tensor, delta_q: Tensor, eta: Tensor, rule: int, K_global: int) -> Tensor:
    _chk(hebb, delta_q, eta)
    out = torch.empty_like(hebb)
    _lib.call("pu_trace_apply", hebb.data_ptr(), delta_q.data_ptr(), K_global, eta.data_ptr(), rule, out.data_ptr(), hebb.shape[0], _s())
    return out


@trace_apply.register_fake
def _(hebb, delta_q, eta, rule, K_global):
    return torch.empty_like(hebb)
