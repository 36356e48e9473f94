"""ResNet stem convolution (3->64, 7x7/2) at batch 256, bf16 NHWC: cuDNN vs the
tensor-core implicit-GEMM kernels in ops/csrc/stem_kernels.cu (fwd and wgrad)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochastic_gradient_push_b200.ops import native       # noqa: E402

torch.backends.cudnn.benchmark = True
C = native.load()
dev = 'cuda'
B = 256
DT = torch.float32 if '--fp32' in sys.argv else torch.bfloat16     # fp32: TF32 mma.sync kernels vs cuDNN TF32
x = torch.randn(B, 3, 224, 224, device=dev).to(DT).contiguous(memory_format=torch.channels_last)
w = torch.randn(64, 3, 7, 7, device=dev).to(DT).contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    ts = []
    for _ in range(iters + 2):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2]


y = F.conv2d(x, w, None, 2, 3)
dy = torch.randn_like(y)
wr = w.clone().requires_grad_(True)
t_cudnn_f = timed(lambda: F.conv2d(x, w, None, 2, 3))
t_cudnn_fb = timed(lambda: torch.autograd.grad(F.conv2d(x, wr, None, 2, 3), [wr], dy))
t_cudnn_w = timed(lambda: torch.ops.aten.convolution_backward(
    dy, x, w, None, [2, 2], [3, 3], [1, 1], False, [0, 0], 1, [False, True, False]))
t_ours_f = timed(lambda: C.stem_forward(x, w))
t_ours_b = timed(lambda: C.stem_wgrad(x, dy))
io = (x.numel() + y.numel()) * x.element_size() / 1e6
print('stem conv batch %d: cuDNN fwd %.3f ms, wgrad-only %.3f ms, fwd+wgrad %.3f ms | ours fwd %.3f ms (%.2f TB/s), wgrad %.3f ms'
      % (B, t_cudnn_f, t_cudnn_w, t_cudnn_fb, t_ours_f, io / t_ours_f / 1e3, t_ours_b))
