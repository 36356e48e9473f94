"""Single-GPU driver for `ncu --set full` captures of the fused gossip kernel.

ncu serialises and replays kernels, so two co-operating ranks would dead-lock;
instead ONE rank is given a self-loop in-neighbour (device table row
``n_in=1, in[0]=self``): the kernel runs its complete code path -- SGD, publish,
flag acquire, weighted "peer" loads through the pointer table, mix, de-bias --
with the peer traffic landing on local HBM instead of NVLink.  NVLink numbers
come from benchmarks/mix_bench.py (CUDA events, multi-GPU).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stochastic_gradient_push_b200 as sgp                                # noqa: E402
from stochastic_gradient_push_b200.ops.peer_mix import GossipEngine        # noqa: E402
from stochastic_gradient_push_b200.parallel.symmetric import LocalWorld    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--numel', type=int, default=25559040 // 4096 * 4096 + 4096)
    ap.add_argument('--iters', type=int, default=6)
    ap.add_argument('--bf16', action='store_true')
    ap.add_argument('--mode', default='mix', choices=['mix', 'local', 'gather'])
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    n = args.numel
    graph = sgp.NPeerDynamicDirectedExponentialGraph(0, 1)
    z = torch.randn(n, device=dev)
    grad = torch.randn(n, device=dev)
    if args.bf16:
        grad = grad.bfloat16()
    shadow = torch.zeros(n, device=dev, dtype=torch.bfloat16) if args.bf16 else None
    eng = GossipEngine(LocalWorld(1).view(0), z, graph, sgp.UniformMixing(graph, dev), grad=grad,
                       momentum=torch.zeros(n, device=dev), shadow=shadow, timeout_s=5.0,
                       with_residual=True, gather_grid=32)
    C = eng.C
    if args.mode in ('mix', 'gather'):
        table = torch.full((1, C.TABLE_ROW), -1, dtype=torch.int32, device=dev)
        table[0, 0] = 1      # n_in
        table[0, 1] = 1      # n_out
        table[0, 2] = 0      # in[0]  = self
        table[0, 2 + C.MAX_PEERS] = 0
        wtable = torch.zeros((1, C.WTABLE_ROW), dtype=torch.float32, device=dev)
        wtable[0, 0] = 0.5
        wtable[0, 1] = 0.5
        eng.ctx.set_schedule(table, wtable)
    eng.set_hyper(1e-3, 0.9, 1e-4, True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(args.iters):
        flush.zero_()
        if args.mode == 'mix':
            eng.mix(sgd=True, zero_grad=True)
        elif args.mode == 'gather':
            eng.publish(sgd=False)
            flush.zero_()
            eng.gather()
        else:
            eng.local(sgd=True, zero_grad=True)
        torch.cuda.synchronize()
    eng.check()
    print('ok', eng.device_step)


if __name__ == '__main__':
    main()
