#!/usr/bin/env python
"""
bench.py -- ResNet-50 gossip-SGD throughput (BASELINE.json headline metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--algo sgp|osgp|dpsgd|ar]
                    [--batch-size B] [--impl ours|reference]

N > 1 is launched by the driver as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
(one rank per GPU; RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* come from the env).

What is timed (our arm):
  * `value`  : K replays of the captured training step -- forward (bf16 autocast,
    NHWC) + loss + backward + ONE fused sm_100a kernel (SGD-momentum + push-sum
    publish + P2P pull over NVLink + mix + de-bias) -- inputs resident on the
    device, CUDA events on the launching stream, barrier + synchronize on both
    sides, max over ranks.  Whole-job images/s.
  * `e2e`    : the same K steps through the public API (`GossipTrainer.step`) with
    the step's inputs copied from pinned host memory every step (prefetch stream)
    and the step's loss copied back to pinned host memory every step.
Synthetic 3x224x224 fp32 images, random-init ResNet-50 (no network / datasets).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--algo', default='sgp', choices=['sgp', 'osgp', 'dpsgd', 'ar', 'adpsgd'])
    ap.add_argument('--batch-size', '--batch_size', dest='batch_size', type=int, default=256,
                    help='per-agent batch; 256 = every shipped job script of the reference '
                         '(job_scripts/submit_*.sh: --batch_size 256 per gossip agent); one agent '
                         'per B200 here')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--model', default='resnet50')
    ap.add_argument('--ppi', type=int, default=1)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--autocast', action='store_true',
                    help='bf16 via torch.autocast instead of the bf16 shadow-weight twin')
    ap.add_argument('--skip-e2e', action='store_true')
    return ap.parse_args()


# --------------------------------------------------------------------------- #
# clocks sampling (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------- #
class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                 '-lms', '100', '-i', str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None,
                'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# --------------------------------------------------------------------------- #
def run_ours(args):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import stochastic_gradient_push_b200 as sgp
    from stochastic_gradient_push_b200 import models
    from stochastic_gradient_push_b200.optim import FusedGossipSGD
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    from stochastic_gradient_push_b200.parallel.trainer import GossipTrainer

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1 + rank)

    bs, K, W = args.batch_size, args.steps, args.warmup
    amp = torch.bfloat16 if args.dtype == 'bf16' else None
    net = models.MODEL_ZOO[args.model]()
    models.init_imagenet_in_1hr(net)
    net = net.to(dev).to(memory_format=torch.channels_last)

    if args.algo == 'adpsgd':
        return run_adpsgd(args, net, rank, world, dev, amp)
    if args.algo == 'ar':
        from stochastic_gradient_push_b200.parallel.allreduce import AllReduceDataParallel, ARTrainer
        model = AllReduceDataParallel(net)
        trainer = ARTrainer(model, lr=0.1 * bs * world / 256, momentum=0.9, weight_decay=1e-4,
                            nesterov=True, amp_dtype=amp, use_cuda_graph=not args.no_graph)
        kernels_per_step = 1
        graph_name = 'all-reduce'
    else:
        if args.algo == 'dpsgd':
            graph = sgp.RingGraph(rank, world, peers_per_itr=args.ppi)
            graph_name = 'static ring'
        else:
            graph = sgp.NPeerDynamicDirectedExponentialGraph(rank, world, peers_per_itr=args.ppi)
            graph_name = 'n-peer dynamic directed exponential'
        model = GossipDataParallel(net, graph=graph, push_sum=(args.algo != 'dpsgd'),
                                   overlap=(args.algo == 'osgp'), rank=rank, world_size=world,
                                   verbose=False, heartbeat_timeout=60,
                                   compute_dtype=(torch.bfloat16 if (amp is not None and not args.autocast)
                                                  else None))
        opt = FusedGossipSGD(model, lr=0.1 * bs * world / 256, momentum=0.9,
                             weight_decay=1e-4, nesterov=True)
        trainer = GossipTrainer(model, opt, amp_dtype=amp, use_cuda_graph=not args.no_graph)
        kernels_per_step = 2 if (args.algo == 'osgp' and world > 1) else 1
    # our own sm_100a kernels per training step: 6 per fused BatchNorm (stats, finalize,
    # apply | reduce, finalize, dx), 2 per NHWC max-pool, plus the gossip kernel(s)
    from stochastic_gradient_push_b200.ops.fused_bn import FusedBatchNormAct2d, MaxPool2dNHWC
    n_bn = sum(isinstance(m, FusedBatchNormAct2d) for m in net.modules())
    n_pool = sum(isinstance(m, MaxPool2dNHWC) for m in net.modules())
    kernels_per_step += 6 * n_bn + 2 * n_pool
    # (for 1x1 convolutions the "stats" launch is the tcgen05 GEMM whose epilogue produces them);
    # + stem convolution forward / wgrad, + one dgrad GEMM (skip gradient folded in) per bottleneck
    from stochastic_gradient_push_b200.ops import fused_bn as _fb
    from stochastic_gradient_push_b200.models.resnet import Bottleneck
    kernels_per_step += 2
    if _fb.USE_TCGEN05_CONV1X1:
        kernels_per_step += sum(isinstance(m, Bottleneck) for m in net.modules())

    # synthetic data: a small pool of pinned host batches (the loader's output)
    g = torch.Generator().manual_seed(1234 + rank)
    pool = [(torch.randn(bs, 3, 224, 224, generator=g).pin_memory(),
             torch.randint(0, 1000, (bs,), generator=g).pin_memory()) for _ in range(4)]
    h2d_bytes = pool[0][0].numel() * 4 + pool[0][1].numel() * 8

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: eager iterations, graph capture, a few replays ------------
    for i in range(max(W, 5)):
        trainer.step(*pool[i % len(pool)])
    sync_all()

    # ---- device-only timed region -------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(trainer.stream)
    for i in range(K):
        trainer.step_resident()
    e1.record(trainer.stream)
    sync_all()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    value = bs * world * K / (ms / 1e3)

    # ---- end-to-end timed region (public API, H2D + D2H every step) ---------
    e2e = None
    if not args.skip_e2e:
        trainer.prefetch(*pool[0])
        sync_all()
        e0.record(trainer.stream)
        slots = []
        for i in range(K):
            nxt = pool[(i + 1) % len(pool)]
            slots.append(trainer.step(None, None, nxt[0], nxt[1]))
        e1.record(trainer.stream)
        sync_all()
        losses = [float(trainer.loss_ring[s]) for s in slots]     # D2H results, all K read
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        ms2 = ms2.item()
        e2e = {'value': round(bs * world * K / (ms2 / 1e3), 2), 'unit': 'images/s',
               'ms_per_step': round(ms2 / K, 4),
               'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
               'last_loss': round(losses[-1], 4)}
    trainer.finish()

    if rank == 0:
        out = {
            'metric': 'resnet50_%s_images_per_sec' % args.algo, 'value': round(value, 2),
            'unit': 'images/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': round(ms / K, 4), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic', 'impl': 'ours',
            'config': {'model': args.model, 'algorithm': args.algo, 'graph': graph_name,
                       'peers_per_itr': args.ppi, 'per_gpu_batch': bs, 'global_batch': bs * world,
                       'image': '3x224x224', 'parallelism': 'dp%d-gossip' % world,
                       'master_weights': 'fp32 flat arena', 'layout': 'NHWC',
                       'bf16_path': ('autocast' if (args.autocast or args.algo == 'ar') else
                                     'shadow weights written by the gossip kernel'),
                       'cuda_graph': not args.no_graph,
                       'conv1x1': ('tcgen05 GEMM (TMA/TMEM), BN statistics and skip gradient fused '
                                   'into its epilogues' if _fb.USE_TCGEN05_CONV1X1 else 'library'),
                       'l2': 'per-step working set (activations+weights > 1 GB) exceeds the '
                             '126 MB L2; no explicit flush'},
            'clocks': clocks, 'e2e': e2e,
            'gpu_launches': kernels_per_step * K,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_adpsgd(args, net, rank, world, dev, amp):
    """AD-PSGD (BilatGossipDataParallel): asynchronous by construction, so the step is
    the reference-style eager loop (forward, backward hook = push grads + pull model,
    local step); gossip + the gossip-side fused SGD run on the low-priority stream."""
    import torch
    import torch.distributed as dist
    import stochastic_gradient_push_b200 as sgp
    from stochastic_gradient_push_b200.parallel.ad_psgd import BilatGossipDataParallel
    bs, K, W = args.batch_size, args.steps, args.warmup
    lr = 0.1 * bs * world / 256
    model = BilatGossipDataParallel(net, rank=rank, world_size=world,
                                    graph_class=sgp.DynamicBipartiteExponentialGraph,
                                    mixing_class=sgp.UniformMixing, lr=lr, momentum=0.9,
                                    weight_decay=1e-4, nesterov=True, verbose=False,
                                    heartbeat_timeout=60)
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4, nesterov=True)
    crit = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(1234 + rank)
    pool = [(torch.randn(bs, 3, 224, 224, generator=g).pin_memory(),
             torch.randint(0, 1000, (bs,), generator=g).pin_memory()) for _ in range(4)]
    loss_host = torch.zeros(K + W + 8).pin_memory()
    model.train()
    model.enable_gossip()

    def step(i):
        x, y = pool[i % 4]
        x = x.to(dev, non_blocking=True).contiguous(memory_format=torch.channels_last)
        y = y.to(dev, non_blocking=True)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp is not None):
            loss = crit(model(x).float(), y)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=False)
        loss_host[i:i + 1].copy_(loss.detach().view(1), non_blocking=True)

    for i in range(max(W, 5)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(int(os.environ.get('LOCAL_RANK', 0))) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rounds0 = model.rounds_completed
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    rounds = torch.tensor([float(model.rounds_completed - rounds0)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rounds, op=dist.ReduceOp.MIN)
    ms = ms.item()
    value = bs * world * K / (ms / 1e3)
    model.disable_gossip()
    if rank == 0:
        print(json.dumps({
            'metric': 'resnet50_adpsgd_images_per_sec', 'value': round(value, 2), 'unit': 'images/s',
            'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': round(ms / K, 4),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype,
            'data': 'synthetic', 'impl': 'ours', 'value_is_e2e': True,
            'config': {'model': args.model, 'algorithm': 'adpsgd', 'graph': 'dynamic bipartite exponential',
                       'per_gpu_batch': bs, 'global_batch': bs * world, 'image': '3x224x224',
                       'parallelism': 'dp%d-bilateral-gossip' % world, 'cuda_graph': False,
                       'gossip_rounds_in_timed_region_min_over_ranks': int(rounds.item()),
                       'l2': 'per-step working set exceeds the 126 MB L2; no explicit flush'},
            'clocks': clocks,
            'e2e': {'value': round(value, 2), 'unit': 'images/s',
                    'h2d_bytes_per_step': pool[0][0].numel() * 4 + bs * 8, 'd2h_bytes_per_step': 4},
            'gpu_launches': None}))
    if world > 1:
        dist.barrier()
    model.shutdown()
    if world > 1:
        dist.destroy_process_group()


def relaunch_under_torchrun(args):
    """`python bench.py --gpus N` without a launcher: start one rank per GPU ourselves."""
    import socket
    sock = socket.socket()
    sock.bind(('127.0.0.1', 0))
    port = sock.getsockname()[1]
    sock.close()
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node',
           str(args.gpus), '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.abspath(__file__)] + sys.argv[1:]
    sys.exit(subprocess.call(cmd))


def main():
    args = parse()
    if args.gpus > 1 and 'RANK' not in os.environ:
        relaunch_under_torchrun(args)
    if args.impl == 'reference':
        sys.path.insert(0, os.path.join(ROOT, 'baseline'))
        try:
            import run_reference
        except Exception as e:       # pragma: no cover
            print(json.dumps({'impl': 'reference', 'unavailable': 'shim import failed: %r' % e}))
            return
        try:
            run_reference.main(args)
        except SystemExit:
            raise
        except Exception as e:
            import traceback
            traceback.print_exc()
            if int(os.environ.get('RANK', 0)) == 0:
                print(json.dumps({'impl': 'reference',
                                  'unavailable': 'reference run failed: %s' % str(e)[:200]}))
        return
    run_ours(args)


if __name__ == '__main__':
    main()
