"""Where does the Overlap-SGP gather run relative to forward/backward?

    torchrun --nproc-per-node 2 benchmarks/osgp_trace.py [--algo osgp|sgp] [--batch-size 64] [--no-graph]

Runs the flagship trainer, records a CUPTI timeline (torch.profiler) of a few steps and prints,
per step: the publish kernel, the gather kernel (side stream), the first / last model kernel and
how long the main stream idles between the last backward kernel and the next publish.  This is
the diagnosis tool behind the OSGP numbers in profiles/README.md (no nsys in this image).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stochastic_gradient_push_b200 as sgp                                        # noqa: E402
from stochastic_gradient_push_b200 import models                                   # noqa: E402
from stochastic_gradient_push_b200.optim import FusedGossipSGD                     # noqa: E402
from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel  # noqa: E402
from stochastic_gradient_push_b200.parallel.trainer import GossipTrainer           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--algo', default='osgp')
    ap.add_argument('--batch-size', type=int, default=64)
    ap.add_argument('--dtype', default='fp32')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    dev = torch.device('cuda', torch.cuda.current_device())
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.backends.cudnn.benchmark = True
    net = models.init_imagenet_in_1hr(models.resnet50()).to(dev).to(memory_format=torch.channels_last)
    graph = sgp.NPeerDynamicDirectedExponentialGraph(rank, world)
    amp = torch.bfloat16 if args.dtype == 'bf16' else None
    model = GossipDataParallel(net, graph=graph, overlap=(args.algo == 'osgp'), rank=rank, world_size=world,
                               heartbeat_timeout=60, compute_dtype=amp)
    opt = FusedGossipSGD(model, lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=True)
    tr = GossipTrainer(model, opt, amp_dtype=amp, use_cuda_graph=not args.no_graph)
    x = torch.randn(args.batch_size, 3, 224, 224).pin_memory()
    y = torch.randint(0, 1000, (args.batch_size,)).pin_memory()
    for _ in range(8):
        tr.step(x, y)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            tr.step_resident()
        torch.cuda.synchronize()
    tr.finish()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    if rank == 0:
        gossip = [k for k in ks if 'sgp_' in k[2]]
        t0 = ks[0][0] if ks else 0
        print('rank0: %d device events, %d gossip kernels' % (len(ks), len(gossip)))
        pubs = [k for k in gossip if 'sgp_step' in k[2]]
        rows = []
        for i, p in enumerate(pubs):
            nxt = pubs[i + 1][0] if i + 1 < len(pubs) else None
            inside = [k for k in ks if k[0] >= p[1] and (nxt is None or k[0] < nxt) and 'sgp_' not in k[2]]
            gat = [k for k in gossip if 'gather' in k[2] and k[0] >= p[0] and (nxt is None or k[0] < nxt)]
            row = {'step_kernel': p[2][:40], 'step_kernel_us': round(p[1] - p[0], 1),
                   'first_model_kernel_after_us': round(inside[0][0] - p[1], 1) if inside else None,
                   'model_span_us': round(inside[-1][1] - inside[0][0], 1) if inside else None,
                   'gap_to_next_step_kernel_us': round(nxt - inside[-1][1], 1) if (inside and nxt) else None}
            if gat:
                row['gather_start_after_publish_end_us'] = round(gat[0][0] - p[1], 1)
                row['gather_us'] = round(gat[0][1] - gat[0][0], 1)
                row['gather_end_before_model_end_us'] = round(inside[-1][1] - gat[0][1], 1) if inside else None
            rows.append(row)
            print(json.dumps(row))
        if args.out:
            with open(args.out, 'w') as f:
                json.dump(rows, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
