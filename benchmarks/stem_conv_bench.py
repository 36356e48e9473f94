"""Does padding the 3 input channels of the ResNet stem (7x7/2 conv) to 4 or 8 buy a
faster cuDNN kernel in bf16 NHWC?  fwd + wgrad (the input needs no gradient)."""
import torch
import torch.nn.functional as F

torch.backends.cudnn.benchmark = True
dev = 'cuda'
B = 256


def bench(cin):
    x = torch.randn(B, cin, 224, 224, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, cin, 7, 7, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    w.requires_grad_(True)
    g = None
    for it in range(12):
        if it == 4:
            torch.cuda.synchronize()
            e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e0.record()
        y = F.conv2d(x, w, None, 2, 3)
        if it == 11:
            e1.record()
        if g is None:
            g = torch.randn_like(y)
        y.backward(g)
        w.grad = None
    e2.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e2) / 8


for c in (3, 4, 8):
    print('cin=%d  fwd+wgrad %.3f ms' % (c, bench(c)))
