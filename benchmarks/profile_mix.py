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
    ap.add_argument('--two-gpu', action='store_true',
                    help='ONE process, two GPUs (LocalWorld over plain peer access): rank 0 on cuda:0 pulls '
                         "rank 1's outbox from cuda:1 over NVLink.  The launches are serialised (r1.publish -> "
                         'r0.mix -> r1.gather) so that the capture works under ncu, which runs one kernel of '
                         'the process at a time: every flag r0.mix waits for is already set when it starts.')
    ap.add_argument('--no-pipe', action='store_true')
    args = ap.parse_args()
    if args.two_gpu:
        return two_gpu(args)
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


def two_gpu(args):
    n = args.numel
    lw = LocalWorld(2, [0, 1])
    engs = []
    for r in range(2):
        dev = torch.device('cuda', r)
        torch.cuda.set_device(dev)
        graph = sgp.NPeerDynamicDirectedExponentialGraph(r, 2)
        z = torch.randn(n, device=dev)
        grad = torch.randn(n, device=dev)
        eng = GossipEngine(lw.view(r), z, graph, sgp.UniformMixing(graph, dev), grad=grad,
                           momentum=torch.zeros(n, device=dev), timeout_s=5.0, with_residual=True, gather_grid=32)
        eng.ctx.set_pipe(not args.no_pipe)
        eng.set_hyper(1e-3, 0.9, 1e-4, True)
        engs.append(eng)
    flush = [torch.empty(256 << 20, dtype=torch.uint8, device=torch.device('cuda', r)) for r in range(2)]
    ms = []
    for it in range(args.iters):
        for r in range(2):
            torch.cuda.set_device(r)
            flush[r].zero_()
        torch.cuda.set_device(1)
        engs[1].publish(sgd=False)                      # rank 1's outbox + flags for this step
        torch.cuda.synchronize(1)
        torch.cuda.set_device(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engs[0].mix(sgd=True, zero_grad=False)          # the profiled launch: SGD + publish + NVLink pull + mix
        e1.record()
        torch.cuda.synchronize(0)
        ms.append(e0.elapsed_time(e1))
        torch.cuda.set_device(1)
        engs[1].gather()                                # rank 1 pulls rank 0's outbox and acks it
        torch.cuda.synchronize(1)
    for e in engs:
        e.check()
    ms = sorted(ms[2:]) if len(ms) > 2 else ms
    print('two-gpu mix (flags pre-set, no peer skew): median %.4f ms -> %.1f GB/s pulled over NVLink; steps %d / %d'
          % (ms[len(ms) // 2], n * 4 / ms[len(ms) // 2] / 1e6, engs[0].device_step, engs[1].device_step))


if __name__ == '__main__':
    main()
