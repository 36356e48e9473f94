"""
FusedBatchNormAct2d: BatchNorm2d (+ residual add) (+ ReLU) as ONE op, backed by
the sm_100a NHWC kernels in ``csrc/bn_kernels.cu``.

Why it exists: an ncu launch list of the ResNet-50 training step at batch 256
(``profiles/launches_r1_bs256_before_fused_bn.csv``) attributes ~59 % of the
GPU time to the framework's channels-last BatchNorm kernels and ~13 % to
stand-alone ReLU / add / cast kernels, versus ~15 % for the tensor-core
convolutions.  All of that work is memory-bound, so it is rewritten as the
minimum number of 16-byte-vectorised passes over HBM (forward 2 reads + 1
write instead of 5 reads + 3 writes; backward recomputes the ReLU mask from the
saved input instead of reading a mask/output tensor).

The module is a drop-in ``nn.BatchNorm2d`` subclass (same parameters, buffers
and ``state_dict`` keys); ``forward(x, residual=None, relu=False)``.  Anything
the kernels do not cover (CPU tensors, NCHW layout, C % 8 != 0, fp16/fp64)
runs the reference composition ``relu(batch_norm(x) + residual)``, which is
also the oracle in the numerics tests.

``conv_bn_act`` / ``conv_bn_act_split`` go one step further for 1x1 / stride-1
convolutions: the convolution itself runs as the hand-written tcgen05 GEMM of
``csrc/conv1x1_kernels.cu`` (TMA operand ring, TMEM accumulators), whose
epilogue produces the BatchNorm statistics (no separate statistics pass) and,
in backward, adds the skip-branch gradient of a residual block inside the
dgrad GEMM (no autograd accumulation pass).  Also here: the NHWC max-pool and
the tensor-core stem convolution wrappers.
"""

from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native


FORCE_REFERENCE = False      # debugging / A-B switch: route everything to the PyTorch composition

# 1x1 convolutions feeding a training-mode BatchNorm run as the tcgen05 GEMM with the
# statistics fused into its epilogue (csrc/conv1x1_kernels.cu).  SGP_B200_CONV1X1=0 routes
# them back to the library convolution + stand-alone statistics pass.
USE_TCGEN05_CONV1X1 = os.environ.get('SGP_B200_CONV1X1', '1') != '0'


# the tensor-core stem kernels (csrc/stem_kernels.cu); SGP_B200_STEM=0 -> library convolution
USE_STEM_KERNELS = os.environ.get('SGP_B200_STEM', '1') != '0'


def _can_fuse(x: torch.Tensor) -> bool:
    if FORCE_REFERENCE or not x.is_cuda or not native.available():
        return False
    return bool(native.load().bn_can_fuse(x))


class _FusedBNAct(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, nbt,
                training, momentum, eps, relu):
        C = native.load()
        y, coef = C.bn_forward(x, residual, weight, bias, running_mean, running_var, nbt,
                               training, momentum, eps, relu)
        ctx.relu = relu
        ctx.add = residual is not None
        ctx.training = training
        if ctx.add:
            ctx.save_for_backward(x, y, coef, weight)
        else:
            ctx.save_for_backward(x, coef, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        C = native.load()
        if ctx.add:
            x, y, coef, weight = ctx.saved_tensors
        else:
            x, coef, weight = ctx.saved_tensors
            y = None
        if not ctx.training:
            # eval-mode BN is an affine map: dx = scale * dz
            scale = coef[2].to(dy.dtype).view(1, -1, 1, 1)
            dz = dy
            if ctx.relu:
                out = y if ctx.add else torch.relu(F.batch_norm(x, None, None, training=False))
                dz = dy * (out > 0)
            return dz * scale, (dz if ctx.add else None), None, None, None, None, None, \
                None, None, None, None
        dx, dz, ggamma, gbeta = C.bn_backward(dy, x, y, coef, ctx.relu, ctx.add)
        return dx, (dz if ctx.add else None), ggamma, gbeta, None, None, None, None, None, None, None


def fused_bn_act(x, weight, bias, running_mean, running_var, num_batches_tracked=None,
                 residual=None, relu=False, training=True, momentum=0.1, eps=1e-5):
    """Functional form; falls back to plain PyTorch when the kernels cannot run."""
    fusable = _can_fuse(x) and weight is not None and weight.dtype == torch.float32 \
        and (residual is None or (relu and residual.dtype == x.dtype
                                  and residual.stride() == x.stride()))
    if fusable and (training or not torch.is_grad_enabled() or not x.requires_grad):
        return _FusedBNAct.apply(x, residual, weight, bias, running_mean, running_var,
                                 num_batches_tracked, training, momentum, eps, relu)
    return reference_bn_act(x, weight, bias, running_mean, running_var, num_batches_tracked,
                            residual, relu, training, momentum, eps)


def reference_bn_act(x, weight, bias, running_mean, running_var, num_batches_tracked=None,
                     residual=None, relu=False, training=True, momentum=0.1, eps=1e-5):
    if training and num_batches_tracked is not None:
        num_batches_tracked.add_(1)
    y = F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


class FusedBatchNormAct2d(nn.BatchNorm2d):
    """``y = relu?(bn(x) (+ residual)?)`` -- see module docstring."""

    def forward(self, x, residual=None, relu=False):
        training = self.training or (self.running_mean is None)
        momentum = 0.1 if self.momentum is None else self.momentum
        return fused_bn_act(
            x, self.weight, self.bias,
            self.running_mean if self.track_running_stats else None,
            self.running_var if self.track_running_stats else None,
            self.num_batches_tracked if (self.track_running_stats and training) else None,
            residual=residual, relu=relu, training=training, momentum=momentum, eps=self.eps)


# --------------------------------------------------------------------------- #
# 1x1 convolution + BatchNorm (+ residual) (+ ReLU)
# --------------------------------------------------------------------------- #
class _Conv1x1BNAct(torch.autograd.Function):
    """``act(bn(conv1x1(x, w)) [+ residual])`` in training mode.  Forward: tcgen05 GEMM
    whose epilogue also produces the BatchNorm statistics, finalize, apply.  Backward: the
    fused BN backward, then the library's dgrad / wgrad for the convolution."""

    @staticmethod
    def forward(ctx, x, w, residual, gamma, beta, running_mean, running_var, nbt, momentum, eps, relu):
        C = native.load()
        w16 = w if w.dtype == x.dtype else w.detach().to(x.dtype)      # compute-dtype weights
        yraw, out, coef = C.conv1x1_bn_forward(x, w16, residual, gamma, beta, running_mean, running_var,
                                               nbt, momentum, eps, relu)
        ctx.relu = relu
        ctx.add = residual is not None
        ctx.w_dtype = w.dtype
        if ctx.add:
            ctx.save_for_backward(x, w16, yraw, coef, out)
        else:
            ctx.save_for_backward(x, w16, yraw, coef)
        return out

    @staticmethod
    def backward(ctx, dy):
        C = native.load()
        x, w16, yraw, coef = ctx.saved_tensors[:4]
        out = ctx.saved_tensors[4] if ctx.add else None
        dyraw, dz, ggamma, gbeta = C.bn_backward(dy, yraw, out, coef, ctx.relu, ctx.add)
        dx, dw, _ = torch.ops.aten.convolution_backward(
            dyraw, x, w16, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1,
            [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        if dw is not None and dw.dtype != ctx.w_dtype:
            dw = dw.to(ctx.w_dtype)
        return dx, dw, (dz if ctx.add else None), ggamma, gbeta, None, None, None, None, None, None


class _Conv1x1BNActSplit(torch.autograd.Function):
    """``(act(bn(conv1x1(x, w))), x)`` for a block input that also feeds the skip connection.
    Returning ``x`` through the op routes the skip-branch gradient into ``backward``, where the
    dgrad GEMM adds it in its epilogue (fp32, before rounding) instead of autograd running a
    separate read-read-write accumulation pass over the widest activation of the block."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, running_mean, running_var, nbt, momentum, eps, relu):
        C = native.load()
        w16 = w if w.dtype == x.dtype else w.detach().to(x.dtype)
        yraw, out, coef = C.conv1x1_bn_forward(x, w16, None, gamma, beta, running_mean, running_var,
                                               nbt, momentum, eps, relu)
        ctx.relu = relu
        ctx.w_dtype = w.dtype
        ctx.save_for_backward(x, w16, yraw, coef)
        ctx.set_materialize_grads(False)
        return out, x.view_as(x)

    @staticmethod
    def backward(ctx, d_out, d_skip):
        C = native.load()
        x, w16, yraw, coef = ctx.saved_tensors
        if d_out is None:
            d_out = torch.zeros_like(yraw)
        dyraw, _, ggamma, gbeta = C.bn_backward(d_out, yraw, None, coef, ctx.relu, False)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # dX = dY . W: the same GEMM with the transposed weight as its K-major B operand
            wt = w16.reshape(w16.shape[0], w16.shape[1]).t().contiguous()
            res = d_skip
            if res is not None and (res.dtype != dyraw.dtype
                                    or not res.is_contiguous(memory_format=torch.channels_last)):
                res = res.to(dyraw.dtype).contiguous(memory_format=torch.channels_last)
            if C.conv1x1_can_fuse(dyraw, wt):
                dx = C.conv1x1_forward(dyraw, wt, False, res)
            else:
                dx = torch.ops.aten.convolution_backward(
                    dyraw, x, w16, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False])[0]
                if res is not None:
                    dx = dx + res
        if ctx.needs_input_grad[1]:
            dw = torch.ops.aten.convolution_backward(
                dyraw, x, w16, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [False, True, False])[1]
            if dw.dtype != ctx.w_dtype:
                dw = dw.to(ctx.w_dtype)
        return dx, dw, ggamma, gbeta, None, None, None, None, None, None


def _conv1x1_operands(conv, bn, x, residual, relu):
    """(x, fusable): x possibly cast the way autocast would; fusable says whether the tcgen05
    conv + BN path covers this call."""
    if not (USE_TCGEN05_CONV1X1 and not FORCE_REFERENCE and bn.training and bn.track_running_stats
            and x.is_cuda and native.available() and x.dim() == 4
            and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0)
            and conv.groups == 1 and conv.bias is None and bn.weight is not None
            and bn.weight.dtype == torch.float32 and (residual is None or relu)):
        return x, False
    w = conv.weight
    if x.dtype == torch.float32 and torch.is_autocast_enabled():
        x = x.to(torch.bfloat16)
    if x.dtype == torch.float32:
        # fp32 activations: the GEMM multiplies in TF32 (what the library convolution does by
        # default); honour a user who switched TF32 off
        dtype_ok = w.dtype == torch.float32 and torch.backends.cudnn.allow_tf32
    else:
        dtype_ok = x.dtype == torch.bfloat16 and (w.dtype == torch.bfloat16 or torch.is_autocast_enabled())
    if not (dtype_ok and x.is_contiguous(memory_format=torch.channels_last)
            and (residual is None or (residual.dtype == x.dtype
                                      and residual.is_contiguous(memory_format=torch.channels_last)))):
        return x, False
    wk = w if w.dtype == x.dtype else w.detach().to(x.dtype)
    return x, bool(native.load().conv1x1_can_fuse(x, wk)) and conv.out_channels % 8 == 0


def conv_bn_act(conv: nn.Conv2d, bn: 'FusedBatchNormAct2d', x, residual=None, relu=False):
    """``bn(conv(x), residual=residual, relu=relu)``; 1x1 / stride-1 convolutions of NHWC bf16
    activations in training mode take the fused tcgen05 path."""
    xk, ok = _conv1x1_operands(conv, bn, x, residual, relu)
    if ok:
        momentum = 0.1 if bn.momentum is None else bn.momentum
        return _Conv1x1BNAct.apply(xk, conv.weight, residual, bn.weight, bn.bias, bn.running_mean,
                                   bn.running_var, bn.num_batches_tracked, momentum, bn.eps, relu)
    return bn(conv(x), residual=residual, relu=relu)


def conv_bn_act_split(conv: nn.Conv2d, bn: 'FusedBatchNormAct2d', x, relu=False):
    """``(bn(conv(x), relu=relu), x)`` for an ``x`` that is also consumed by a skip connection:
    use the returned ``x`` for the skip branch and its gradient is added inside the dgrad GEMM
    of ``conv`` (see :class:`_Conv1x1BNActSplit`).  Falls back to ``(conv_bn_act(...), x)``."""
    if x.requires_grad and torch.is_grad_enabled():
        xk, ok = _conv1x1_operands(conv, bn, x, None, relu)
        if ok and xk is x:
            momentum = 0.1 if bn.momentum is None else bn.momentum
            return _Conv1x1BNActSplit.apply(x, conv.weight, bn.weight, bn.bias, bn.running_mean,
                                            bn.running_var, bn.num_batches_tracked, momentum, bn.eps, relu)
    return conv_bn_act(conv, bn, x, relu=relu), x


# --------------------------------------------------------------------------- #
# NHWC max-pool
# --------------------------------------------------------------------------- #
class _MaxPoolNHWC(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, k, s, p):
        C = native.load()
        y, code = C.maxpool_forward(x, k, s, p)
        ctx.save_for_backward(code)
        ctx.geom = (x.shape[2], x.shape[3], k, s, p)
        return y

    @staticmethod
    def backward(ctx, dy):
        C = native.load()
        (code,) = ctx.saved_tensors
        H, W, k, s, p = ctx.geom
        return C.maxpool_backward(dy, code, H, W, k, s, p), None, None, None


class MaxPool2dNHWC(nn.MaxPool2d):
    """``nn.MaxPool2d`` whose channels-last CUDA path is the sm_100a gather
    kernel pair in ``csrc/pool_kernels.cu`` (1-byte argmax codes instead of
    int64 indices, atomic-free backward).  Other inputs use the stock op."""

    def forward(self, x):
        k, s, p = self.kernel_size, self.stride, self.padding
        simple = all(isinstance(v, int) for v in (k, s, p)) and self.dilation == 1 \
            and not self.ceil_mode and not self.return_indices
        if simple and not FORCE_REFERENCE and x.is_cuda and native.available() \
                and native.load().pool_can_fuse(x, k, s, p):
            return _MaxPoolNHWC.apply(x, k, s, p)
        return super().forward(x)


# --------------------------------------------------------------------------- #
# ResNet stem convolution
# --------------------------------------------------------------------------- #
class _StemConv(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight):
        C = native.load()
        w16 = weight                   # compute-dtype weights: bf16, or fp32 (TF32 tensor-core math)
        if w16.dtype != x.dtype or not w16.is_contiguous(memory_format=torch.channels_last):
            w16 = weight.detach().to(x.dtype).contiguous(memory_format=torch.channels_last)
        ctx.save_for_backward(x)
        ctx.w_dtype = weight.dtype
        return C.stem_forward(x, w16)

    @staticmethod
    def backward(ctx, dy):
        C = native.load()
        (x,) = ctx.saved_tensors
        dw = C.stem_wgrad(x, dy)
        return None, dw.to(ctx.w_dtype)


def stem_conv(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """``conv(x)`` for the ResNet stem (3 -> 64 channels, 7x7, stride 2, pad 3, no bias).
    NHWC bf16 CUDA inputs run the tensor-core implicit-GEMM kernels in
    ``csrc/stem_kernels.cu`` (the library needs 1.48 ms + 0.77 ms for this layer at batch
    256; its HBM traffic is worth ~0.1 ms); anything else runs ``conv`` unchanged."""
    ok = (USE_STEM_KERNELS and not FORCE_REFERENCE and x.is_cuda and native.available() and not x.requires_grad
          and conv.in_channels == 3 and conv.out_channels == 64 and conv.kernel_size == (7, 7)
          and conv.stride == (2, 2) and conv.padding == (3, 3) and conv.dilation == (1, 1)
          and conv.groups == 1 and conv.bias is None)
    if ok and x.dtype == torch.float32 and torch.is_autocast_enabled():
        x = x.to(torch.bfloat16)          # what autocast would do inside the convolution
    dtype_ok = x.dtype == torch.bfloat16 or (x.dtype == torch.float32 and conv.weight.dtype == torch.float32
                                             and torch.backends.cudnn.allow_tf32)
    if ok and dtype_ok and x.dim() == 4 \
            and x.is_contiguous(memory_format=torch.channels_last) and x.shape[2] >= 7 and x.shape[3] >= 7:
        return _StemConv.apply(x, conv.weight)
    return conv(x)
