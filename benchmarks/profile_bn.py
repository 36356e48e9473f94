"""Driver for ncu captures / event timing of the fused BN and max-pool kernels at a
ResNet-50 layer1 shape (batch 256: [256, 256, 56, 56] bf16 NHWC = 411 MB per tensor)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochastic_gradient_push_b200.ops.fused_bn import fused_bn_act, MaxPool2dNHWC   # noqa: E402

dev = 'cuda'
DT = torch.float32 if '--fp32' in sys.argv else torch.bfloat16
N, C, H, W = 256, 256, 56, 56
x = torch.randn(N, C, H, W, device=dev).to(DT).contiguous(memory_format=torch.channels_last)
res = torch.randn_like(x)
w = torch.ones(C, device=dev, requires_grad=True)
b = torch.zeros(C, device=dev, requires_grad=True)
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
xs = x.clone().requires_grad_(True)
rs = res.clone().requires_grad_(True)
dy = torch.randn_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bytes_t = x.numel() * x.element_size()


def timed(fn, iters=10):
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def fwd_relu():
    return fused_bn_act(xs, w, b, rm, rv, None, relu=True)


def fwd_add():
    return fused_bn_act(xs, w, b, rm, rv, None, residual=rs, relu=True)


y = fwd_relu()
t_f = timed(fwd_relu)
t_b = timed(lambda: torch.autograd.grad(fwd_relu(), [xs, w, b], dy)) - t_f
ya = fwd_add()
t_fa = timed(fwd_add)
t_ba = timed(lambda: torch.autograd.grad(fwd_add(), [xs, rs, w, b], dy)) - t_fa
gb = bytes_t / 1e6
print('tensor %.0f MB' % gb)
print('BN+ReLU      fwd %.3f ms (3 passes -> %.2f TB/s)   bwd %.3f ms (5 passes -> %.2f TB/s)'
      % (t_f, 3 * gb / t_f / 1e3, t_b, 5 * gb / t_b / 1e3))
print('BN+add+ReLU  fwd %.3f ms (4 passes -> %.2f TB/s)   bwd %.3f ms (7 passes -> %.2f TB/s)'
      % (t_fa, 4 * gb / t_fa / 1e3, t_ba, 7 * gb / t_ba / 1e3))

xp = torch.randn(256, 64, 112, 112, device=dev).to(DT).contiguous(
    memory_format=torch.channels_last).requires_grad_(True)
mp = MaxPool2dNHWC(3, 2, 1)
yp = mp(xp)
gp = torch.randn_like(yp)
t_pf = timed(lambda: mp(xp))
t_pb = timed(lambda: torch.autograd.grad(mp(xp), [xp], gp)) - t_pf
ref_f = timed(lambda: torch.nn.functional.max_pool2d(xp, 3, 2, 1))
ref_b = timed(lambda: torch.autograd.grad(torch.nn.functional.max_pool2d(xp, 3, 2, 1), [xp], gp)) - ref_f
print('maxpool 3x3/2 [256,64,112,112]: ours fwd %.3f bwd %.3f ms | framework fwd %.3f bwd %.3f ms'
      % (t_pf, t_pb, ref_f, ref_b))
