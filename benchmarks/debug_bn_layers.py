"""Per-layer A/B of the fused BN kernels inside a real ResNet-50 pass: for every
FusedBatchNormAct2d call compare the fused output with the PyTorch composition on the
SAME input (forward), and dx / dgamma / dbeta on the SAME upstream gradient."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stochastic_gradient_push_b200.models import resnet50
from stochastic_gradient_push_b200.ops import fused_bn
from stochastic_gradient_push_b200.ops.fused_bn import FusedBatchNormAct2d, fused_bn_act, reference_bn_act

torch.manual_seed(0)
dtype = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == 'bf16' else torch.float32
net = resnet50().cuda().to(memory_format=torch.channels_last)
x = torch.randn(16, 3, 128, 128, device='cuda').contiguous(memory_format=torch.channels_last)
rows = []
orig = FusedBatchNormAct2d.forward


def patched(self, inp, residual=None, relu=False):
    with torch.enable_grad():
        xs = [inp.detach().clone().requires_grad_(True) for _ in range(2)]
        rs = [None if residual is None else residual.detach().clone().requires_grad_(True) for _ in range(2)]
        ws = [self.weight.detach().clone().requires_grad_(True) for _ in range(2)]
        bs = [self.bias.detach().clone().requires_grad_(True) for _ in range(2)]
        outs = []
        for i, fn in enumerate((fused_bn_act, reference_bn_act)):
            xi = xs[i] if i == 0 else xs[i].float()
            ri = rs[i] if (i == 0 or rs[i] is None) else rs[i].float()
            outs.append(fn(xi, ws[i], bs[i], None, None, None, residual=ri, relu=relu, training=True))
        g = torch.randn_like(outs[1])
        outs[0].backward(g.to(outs[0].dtype))
        outs[1].backward(g)
        def rel(a, b):
            return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
        rows.append((tuple(inp.shape), relu, residual is not None, rel(outs[0], outs[1]),
                     rel(xs[0].grad, xs[1].grad), rel(ws[0].grad, ws[1].grad), rel(bs[0].grad, bs[1].grad),
                     (outs[0].float() - outs[1].float()).abs().max().item()))
    return orig(self, inp, residual, relu)


FusedBatchNormAct2d.forward = patched
with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
    y = net(x)
print('%-22s relu add   rel(y)    rel(dx)   rel(dw)   rel(db)   maxabs(y)' % 'shape')
for r in rows:
    print('%-22s %-4s %-4s %.2e  %.2e  %.2e  %.2e  %.2e' % (str(r[0]), r[1], r[2], r[3], r[4], r[5], r[6], r[7]))
