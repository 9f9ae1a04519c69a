"""`project.models.op` surface: FusedLeakyReLU, fused_leaky_relu, upfirdn2d.

Same signatures, defaults and derivative structure (first and second order) as the
reference's wrappers (project/models/op/fused_act.py:19-118, upfirdn2d.py:18-154), on top
of the sm_100a kernels behind e3_fused_bias_act / e3_upfirdn2d.

Deliberate differences from the reference (SURVEY.md §7 "quirks"):
  * no CPU branch — CPU tensors raise (the reference's CPU branch also ignores its
    `negative_slope` argument, fused_act.py:111-115; that quirk is not reproduced);
  * fp32 only (the reference never runs these ops in another dtype).
"""
import torch
from torch import nn
from torch.autograd import Function

from . import _lib


def _bias_act(x, bias, refer, act, grad, alpha, scale):
    """y = fused_bias_act(x, bias, refer, act, grad, alpha, scale) — reference pybind ABI."""
    lib = _lib.load()
    x = _lib.as_f32c(x)
    y = torch.empty_like(x)
    step_b = 1
    for d in x.shape[2:]:
        step_b *= d
    b = _lib.as_f32c(bias) if (bias is not None and bias.numel()) else None
    r = _lib.as_f32c(refer) if (refer is not None and refer.numel()) else None
    if b is not None and x.ndim < 2:
        raise RuntimeError("fused_bias_act: bias needs an input with a channel dim (dim 1)")
    _lib.check(lib.e3_fused_bias_act(_lib.ptr(x), _lib.ptr(b), _lib.ptr(r), _lib.ptr(y), x.numel(),
                                     step_b, b.numel() if b is not None else 0, act, grad,
                                     float(alpha), float(scale), _lib.cur_stream()),
               "e3_fused_bias_act")
    return y


class _FusedLeakyReLUBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, out, has_bias, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        grad_input = _bias_act(grad_output, None, out, 3, 1, negative_slope, scale)
        if has_bias:
            dims = [0] + list(range(2, grad_input.ndim))
            grad_bias = grad_input.sum(dims).detach()
        else:
            grad_bias = grad_output.new_empty(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gg_bias = gradgrad_bias if (gradgrad_bias is not None and gradgrad_bias.numel()) else None
        gradgrad_out = _bias_act(gradgrad_input, gg_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None, None


class _FusedLeakyReLU(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        ctx.has_bias = bias is not None
        out = _bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope, ctx.scale = negative_slope, scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = _FusedLeakyReLUBackward.apply(grad_output, out, ctx.has_bias,
                                                              ctx.negative_slope, ctx.scale)
        return grad_input, (grad_bias if ctx.has_bias else None), None, None


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    """leaky_relu(input + bias[dim 1], negative_slope) * scale  (fused_act.py:106-118)."""
    if not input.is_cuda:
        raise RuntimeError("e3dge_b200.fused_leaky_relu: CUDA tensor required (no CPU path)")
    return _FusedLeakyReLU.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    """fused_act.py:87-103 — `bias` parameter [channel], state_dict key `bias`."""

    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def _upfirdn2d_raw(x4, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """x4 [major,in_h,in_w,minor] -> [major,out_h,out_w,minor]  (upfirdn2d.cpp:12-23)."""
    lib = _lib.load()
    x4 = _lib.as_f32c(x4)
    kernel = _lib.as_f32c(kernel)
    major, in_h, in_w, minor = x4.shape
    kh, kw = kernel.shape
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) // down_y + 1
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) // down_x + 1
    y = torch.empty(major, out_h, out_w, minor, device=x4.device, dtype=torch.float32)
    _lib.check(lib.e3_upfirdn2d(_lib.ptr(x4), _lib.ptr(kernel), _lib.ptr(y), major, in_h, in_w,
                                minor, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0,
                                pad_y1, _lib.cur_stream()), "e3_upfirdn2d")
    return y


class _UpFirDn2dBackward(Function):
    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        up_x, up_y = up
        down_x, down_y = down
        g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1 = g_pad
        grad_output = grad_output.reshape(-1, out_size[0], out_size[1], 1)
        # adjoint = upfirdn2d with the flipped kernel and up/down swapped (upfirdn2d.py:31-42)
        grad_input = _upfirdn2d_raw(grad_output, grad_kernel, down_x, down_y, up_x, up_y, g_pad_x0,
                                    g_pad_x1, g_pad_y0, g_pad_y1)
        grad_input = grad_input.reshape(in_size[0], in_size[1], in_size[2], in_size[3])
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.in_size, ctx.out_size = in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gradgrad_input = gradgrad_input.reshape(-1, ctx.in_size[2], ctx.in_size[3], 1)
        gradgrad_out = _upfirdn2d_raw(gradgrad_input, kernel, ctx.up[0], ctx.up[1], ctx.down[0],
                                      ctx.down[1], *ctx.pad)
        gradgrad_out = gradgrad_out.reshape(ctx.in_size[0], ctx.in_size[1], ctx.out_size[0],
                                            ctx.out_size[1])
        return gradgrad_out, None, None, None, None, None, None, None, None


class _UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        batch, channel, in_h, in_w = input.shape
        ctx.in_size = input.shape
        x4 = input.reshape(-1, in_h, in_w, 1)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        out_h = (in_h * up_y + pad_y0 + pad_y1 - kernel_h) // down_y + 1
        out_w = (in_w * up_x + pad_x0 + pad_x1 - kernel_w) // down_x + 1
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        # padding of the adjoint (upfirdn2d.py:101-106)
        ctx.g_pad = (kernel_w - pad_x0 - 1, in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                     kernel_h - pad_y0 - 1, in_h * up_y - out_h * down_y + pad_y0 - up_y + 1)
        out = _upfirdn2d_raw(x4, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
        return out.reshape(-1, channel, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = _UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down,
                                              ctx.pad, ctx.g_pad, ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """input [N,C,H,W], kernel [kh,kw]; pad=(p0,p1) on both axes  (upfirdn2d.py:145-154)."""
    if not input.is_cuda:
        raise RuntimeError("e3dge_b200.upfirdn2d: CUDA tensor required (no CPU path)")
    return _UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))


# ---- op-level route: the reference's own Python wrappers over this library ----------
# The reference JIT-builds two pybind modules at import (`fused = load("fused", ...)`, fused_act.py:10-16;
# `upfirdn2d_op = load("upfirdn2d", ...)`, upfirdn2d.py:9-15).  These two objects have their call shapes
# (fused_bias_act.cpp:11-20, upfirdn2d.cpp:12-23), so a maintainer replaces the two `load(...)` calls by
# `from e3dge_b200.op import fused` / `upfirdn2d_op` and keeps every line of the reference's autograd code.
class fused:
    fused_bias_act = staticmethod(_bias_act)


class upfirdn2d_op:
    upfirdn2d = staticmethod(_upfirdn2d_raw)
