"""Runs one fused 1x1-conv + BN-statistics GEMM a few times (target for `ncu -k regex:c1_gemm`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stochastic_gradient_push_b200.ops import native   # noqa: E402

if __name__ == '__main__':
    hw, cin, cout, batch = [int(v) for v in (sys.argv[1:5] + ['56', '64', '256', '256'][len(sys.argv) - 1:])]
    mode = int(sys.argv[5]) if len(sys.argv) > 5 else 1      # 0 plain, 1 + BN statistics, 2 + residual
    C = native.load()
    M = batch * hw * hw
    x = torch.randn(M, cin, device='cuda').to(torch.bfloat16)
    w = (torch.randn(cout, cin, device='cuda') * cin ** -0.5).to(torch.bfloat16)
    gamma, beta = torch.ones(cout, device='cuda'), torch.zeros(cout, device='cuda')
    rm, rv = torch.zeros(cout, device='cuda'), torch.ones(cout, device='cuda')
    nbt = torch.zeros((), dtype=torch.long, device='cuda')
    res = torch.randn(M, cout, device='cuda').to(torch.bfloat16) if mode == 2 else None
    for _ in range(4):
        if mode == 1:
            C.conv1x1_bn_forward(x, w, None, gamma, beta, rm, rv, nbt, 0.1, 1e-5, True)
        else:
            C.conv1x1_forward(x, w, False, res)
    torch.cuda.synchronize()
