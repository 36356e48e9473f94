"""Microbenchmark of the fused gossip kernels at ResNet-50 size (25.56 M fp32).

    torchrun --nproc-per-node N benchmarks/mix_bench.py [--numel 25559040] [--ppi 1]

Device-timed with CUDA events (max over ranks), L2 flushed between iterations.
Reports ms per launch, NVLink GB/s pulled per rank and HBM GB/s (algorithmic).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stochastic_gradient_push_b200 as sgp                                # noqa: E402
from stochastic_gradient_push_b200.ops.peer_mix import GossipEngine        # noqa: E402
from stochastic_gradient_push_b200.parallel.symmetric import SymmetricWorld, LocalWorld  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--numel', type=int, default=25559040 // 4096 * 4096 + 4096)
    ap.add_argument('--ppi', type=int, default=1)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--grid', type=int, default=None)
    ap.add_argument('--bf16', action='store_true')
    ap.add_argument('--segments', type=int, default=4)
    ap.add_argument('--mode', default='mix', choices=['mix', 'mix_nosgd', 'publish', 'gather', 'local'])
    ap.add_argument('--no-pipe', action='store_true',
                    help='register-staged sgp_step_kernel instead of the warp-specialised TMA kernel')
    args = ap.parse_args()

    multi = 'RANK' in os.environ
    if multi:
        rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
        dist.init_process_group('nccl')
        sw = SymmetricWorld()
    else:
        rank, world = 0, 1
        sw = LocalWorld(1).view(0)
    dev = torch.device('cuda', torch.cuda.current_device())
    n = args.numel
    graph = sgp.NPeerDynamicDirectedExponentialGraph(rank, world, peers_per_itr=args.ppi)
    z = torch.randn(n, device=dev)
    grad = torch.randn(n, device=dev)
    if args.bf16:
        grad = grad.bfloat16()
    shadow = torch.zeros(n, device=dev, dtype=torch.bfloat16) if args.bf16 else None
    mom = torch.zeros(n, device=dev)
    eng = GossipEngine(sw, z, graph, sgp.UniformMixing(graph, dev), grad=grad, momentum=mom,
                       shadow=shadow, with_residual=True, grid=args.grid, timeout_s=20.0,
                       segments=args.segments)
    eng.ctx.set_pipe(not args.no_pipe)
    eng.set_hyper(1e-3, 0.9, 1e-4, True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def launch():
        if args.mode == 'mix':
            eng.mix(sgd=True, zero_grad=False)
        elif args.mode == 'mix_nosgd':
            eng.mix(sgd=False)
        elif args.mode == 'publish':
            eng.publish(sgd=True, fold=True, zero_grad=False)
            e1.record()
            eng.gather()            # releases the outbox (acks); not timed
        elif args.mode == 'gather':
            eng.publish(sgd=False)
            eng.gather()
        else:
            eng.local(sgd=True, zero_grad=False)

    e0 = e1 = None
    times = []
    for it in range(args.iters + 3):
        flush.zero_()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.mode == 'gather':
            eng.publish(sgd=False)
            torch.cuda.synchronize()
            if multi:
                dist.barrier()
            e0.record()
            eng.gather()
            e1.record()
        else:
            e0.record()
            launch()
            if args.mode != 'publish':
                e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
    eng.check()
    t = torch.tensor([sorted(times)[len(times) // 2], min(times)], device=dev)
    if multi:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    med, best = t.tolist()
    n_in = len(graph.get_peers()[1])
    nvlink_bytes = n_in * n * 4
    if rank == 0:
        print(json.dumps({'mode': args.mode, 'world': world, 'ppi': args.ppi, 'numel': n,
                          'grid': eng.grid, 'segments': args.segments, 'pipe': not args.no_pipe, 'ms_median': round(med, 4), 'ms_best': round(best, 4),
                          'nvlink_GBps_per_rank': round(nvlink_bytes / med / 1e6, 1),
                          'bf16': args.bf16}))
    if multi:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
