"""
Reference arm of bench.py: runs the UNMODIFIED reference
(facebookresearch/stochastic_gradient_push, installed under ``baseline/_ref``)
through its own entry point -- ``gossip_sgd.py``'s ``main()`` -> ``train()``
loop with ``gossip.GossipDataParallel`` / ``torch.optim.SGD`` / its KL-div
criterion / its ``accuracy`` + ``.item()`` logging -- on the same workload as
our arm.  Nothing from ``stochastic_gradient_push_b200`` is imported here.

The only substitutions are the ones SURVEY.md 7.4 lists as unavoidable:

* SLURM env vars are faked (``SLURM_PROCID`` / ``SLURM_NTASKS`` / ``HOSTNAME``)
  because the script reads rank / world / master from them;
* one GPU per process via ``CUDA_VISIBLE_DEVICES`` (the reference drives every
  visible GPU from one process, ``gossip/distributed.py:49-52``);
* ``make_dataloader`` is replaced by a synthetic loader yielding pinned-host
  fp32 3x224x224 batches (there is no ImageNet on the box).  The loader is also
  the stopwatch: it records CUDA events (after a device sync + barrier) when
  batch W and batch W+K are requested, so exactly K iterations of the stock
  loop are timed, end to end (H2D of the batch, fwd/bwd, optimizer.step,
  transfer_params, the loop's own accuracy/.item() reads).
* ``accuracy()`` gets a one-word torch>=1.7 fix (``.view`` -> ``.reshape``);
* world size 1: gossip graphs are undefined for n=1 in the reference (math
  domain error in ``graph_manager``), so N=1 runs its AllReduce-SGD path
  (``--all_reduce True``, DistributedDataParallel); DDP without ``device_ids``
  does not move inputs, so in that mode the loader performs the H2D copy itself
  (still inside the timed region).
"""

from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def unavailable(why: str):
    print(json.dumps({'impl': 'reference', 'unavailable': why}))
    sys.exit(0)


class _Stopwatch(object):
    def __init__(self, warmup, steps):
        self.warmup, self.steps = warmup, steps
        self.t0 = self.t1 = None
        self.ev0 = self.ev1 = None
        self.ms = None

    def mark(self, i, torch, dist):
        if i == self.warmup:
            torch.cuda.synchronize()
            if dist.is_initialized() and dist.get_world_size() > 1:
                dist.barrier()
            torch.cuda.synchronize()
            self.ev0 = torch.cuda.Event(enable_timing=True)
            self.ev1 = torch.cuda.Event(enable_timing=True)
            self.ev0.record()
            self.t0 = time.time()
            return False
        if i == self.warmup + self.steps:
            self.ev1.record()
            torch.cuda.synchronize()
            self.t1 = time.time()
            self.ms = self.ev0.elapsed_time(self.ev1)
            return True
        return False


class _TimedRegionDone(Exception):
    pass


def main_adpsgd(args, rank, world, master_port, torch, dist):
    """AD-PSGD comparator: the reference's own ``gossip_sgd_adpsgd.py`` (copied unmodified into
    ``baseline/_ref/bin`` next to ``gossip_sgd.py``; its setup.py installs only the latter) ->
    ``main()`` -> ``train()`` with ``gossip.BilatGossipDataParallel``: two worlds per rank, the
    training process on ``master_port + 1`` and the forked gossip PROCESS on ``master_port``
    (``gossip_sgd_adpsgd.py:695``, ``gossip/ad_psgd.py:280-284``), CUDA-IPC shared parameter /
    gradient tensors between them.  Substitutions: synthetic loader (= stopwatch), the one-word
    ``accuracy`` fix, the ``forkserver`` start method its ``__main__`` block sets, and a stub for
    its NIC-name probe (needs the `ip` binary, which the image lacks; the name is unused here)."""
    import torch.multiprocessing as mp
    script = os.path.join(REF, 'bin', 'gossip_sgd_adpsgd.py')
    if not os.path.isfile(script):
        unavailable('baseline/_ref/bin/gossip_sgd_adpsgd.py missing (cp from the reference tree)')
    mp.set_start_method('forkserver', force=True)
    # two worlds, both hosted by THEIR rank 0 (training world on port+1, gossip world on port): the
    # launcher's agent store only serves the launcher's own port, so the scripts must create their
    # TCP stores themselves on a port pair next to it
    # torch-version shim for the gossip child (see baseline/compat/sitecustomize.py) and for us
    compat = os.path.join(HERE, 'compat')
    os.environ['PYTHONPATH'] = compat + os.pathsep + os.environ.get('PYTHONPATH', '')
    if not hasattr(dist, '_backend'):
        dist._backend = -1
    os.environ.pop('TORCHELASTIC_USE_AGENT_STORE', None)
    master_port = str(int(master_port) + 17)
    spec = importlib.util.spec_from_file_location('ref_gossip_sgd_adpsgd', script)
    ref = importlib.util.module_from_spec(spec)
    ckpt = tempfile.mkdtemp(prefix='ref_ckpt_') + '/'
    shared = '/tmp/ref_adpsgd_counter_%s.txt' % master_port
    sys.argv = ['gossip_sgd_adpsgd.py', '--bilat', 'True', '--graph_type', '1', '--shared_fpath', shared,
                '--batch_size', str(args.batch_size), '--lr', '0.1', '--num_dataloader_workers', '0',
                '--num_epochs', '1', '--nesterov', 'True', '--warmup', 'True', '--seed', '1',
                '--schedule', '30', '0.1', '60', '0.1', '80', '0.1', '--print_freq', '100',
                '--verbose', 'False', '--train_fast', 'True', '--checkpoint_dir', ckpt,
                '--dataset_dir', '/nonexistent', '--backend', 'nccl',
                '--network_interface_type', 'infiniband', '--master_port', master_port, '--tag', 'ref_']
    spec.loader.exec_module(ref)
    warmup, steps, bs = args.warmup, args.steps, args.batch_size
    watch = _Stopwatch(warmup, steps)
    h2d_bytes = bs * 3 * 224 * 224 * 4 + bs * 8

    class SyntheticLoader(object):
        def __init__(self):
            g = torch.Generator().manual_seed(1234 + rank)
            self.pool = [(torch.randn(bs, 3, 224, 224, generator=g).pin_memory(),
                          torch.randint(0, 1000, (bs,), generator=g).pin_memory()) for _ in range(4)]

        def __len__(self):
            return 5005

        def __iter__(self):
            for i in range(warmup + steps + 1):
                if watch.mark(i, torch, dist):
                    raise _TimedRegionDone()      # leave the reference's epoch loop
                yield self.pool[i % len(self.pool)]

    class _Sampler(object):
        def set_epoch(self, e):
            pass

    ref.make_dataloader = lambda a, train=True: ((SyntheticLoader(), _Sampler()) if train else [])
    # the script probes the NIC name with `ip link show up` unconditionally (gossip_sgd_adpsgd.py:174);
    # the image has no `ip` binary and the name is unused with --network_interface_type infiniband
    # (no NCCL_SOCKET_IFNAME is exported), so the probe is stubbed
    ref.get_tcp_interface_name = lambda network_interface_type='ethernet': 'lo'

    def accuracy(output, target, topk=(1,)):
        with torch.no_grad():
            maxk = max(topk)
            _, pred = output.topk(maxk, 1, True, True)
            pred = pred.t()
            correct = pred.eq(target.view(1, -1).expand_as(pred))
            return [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0))
                    for k in topk]

    ref.accuracy = accuracy
    t_all = time.time()

    # the reference's training process waits forever on its gossip process's flags
    # (gossip/ad_psgd.py:237-249); if that child dies (it does on current PyTorch without the shim
    # above) or nothing completes in time, report `unavailable` instead of hanging the harness
    limit_s = float(os.environ.get('SGP_REF_ADPSGD_LIMIT_S', 600))

    def _watchdog():
        seen_child = False
        while watch.ms is None:
            time.sleep(1.0)
            kids = mp.active_children()
            seen_child = seen_child or bool(kids)
            why = None
            if seen_child and not kids:
                why = 'the reference gossip process exited before the timed region completed'
            elif time.time() - t_all > limit_s:
                why = 'reference AD-PSGD run did not complete within %d s' % int(limit_s)
            if why is not None:
                if rank == 0:
                    print(json.dumps({'impl': 'reference', 'unavailable': why}))
                    sys.stdout.flush()
                for child in kids:
                    child.terminate()
                os._exit(0)

    import threading
    threading.Thread(target=_watchdog, name='ref-adpsgd-watchdog', daemon=True).start()
    try:
        ref.main()
    except _TimedRegionDone:
        pass
    torch.cuda.synchronize()
    if watch.ms is None:
        unavailable('reference AD-PSGD loop ended before the timed region completed')
    ms = torch.tensor([watch.ms], device='cuda')
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    value = bs * world * steps / (ms / 1e3)
    if rank == 0:
        print(json.dumps({
            'impl': 'reference', 'metric': 'resnet50_adpsgd_images_per_sec', 'value': round(value, 2),
            'unit': 'images/s', 'n_gpus': world, 'steps': steps, 'warmup': warmup,
            'ms_per_step': round(ms / steps, 4), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'fp32 (cuDNN TF32 conv default)', 'data': 'synthetic',
            'value_is_e2e': True,
            'config': {'model': 'resnet50 (torchvision, reference init_model)', 'algorithm': 'adpsgd',
                       'per_gpu_batch': bs, 'global_batch': bs * world, 'image': '3x224x224',
                       'parallelism': 'dp%d' % world,
                       'entry': 'baseline/_ref/bin/gossip_sgd_adpsgd.py main()->train(), stock; gossip '
                                'process + second process group as in the reference'},
            'e2e': {'value': round(value, 2), 'unit': 'images/s', 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': 12},
            'gpu_launches': 0, 'wall_s': round(time.time() - t_all, 1)}))
    sys.stdout.flush()
    try:
        dist.barrier()
    except Exception:
        pass
    for child in mp.active_children():           # the forked gossip process never exits on its own
        child.terminate()
    os._exit(0)


def main(args):
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    if not os.path.isdir(os.path.join(REF, 'gossip')):
        unavailable('baseline/_ref missing (pip install --target baseline/_ref of the reference)')

    # one GPU per process, set before CUDA initialises
    os.environ['CUDA_VISIBLE_DEVICES'] = os.environ.get(
        'SGP_REF_VISIBLE', str(local_rank))
    os.environ['SLURM_PROCID'] = str(rank)
    os.environ['SLURM_NTASKS'] = str(world)
    os.environ['SLURM_LOCALID'] = '0'            # (gossip_sgd_adpsgd.py:627; one visible GPU per process)
    os.environ['HOSTNAME'] = os.environ.get('MASTER_ADDR', '127.0.0.1')
    master_port = os.environ.get('MASTER_PORT', '40100')

    sys.path.insert(0, REF)
    sys.path = [p for p in sys.path if os.path.abspath(p or '.') != os.path.dirname(HERE)]
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        unavailable('no CUDA device')
    import gossip
    assert os.path.abspath(gossip.__file__).startswith(REF), gossip.__file__

    algo = args.algo
    if world == 1:
        algo = 'ar'
    if algo == 'adpsgd':
        return main_adpsgd(args, rank, world, master_port, torch, dist)
    spec = importlib.util.spec_from_file_location(
        'ref_gossip_sgd', os.path.join(REF, 'bin', 'gossip_sgd.py'))
    ref = importlib.util.module_from_spec(spec)
    ckpt = tempfile.mkdtemp(prefix='ref_ckpt_') + '/'
    argv = ['gossip_sgd.py', '--batch_size', str(args.batch_size), '--lr', '0.1',
            '--num_dataloader_workers', '0', '--num_epochs', '1',
            '--nesterov', 'True', '--warmup', 'True', '--seed', '1',
            '--schedule', '30', '0.1', '60', '0.1', '80', '0.1',
            '--print_freq', '100', '--verbose', 'False', '--train_fast', 'True',
            '--checkpoint_dir', ckpt, '--dataset_dir', '/nonexistent',
            '--backend', 'nccl', '--network_interface_type', 'infiniband',
            '--master_port', master_port, '--num_itr_ignore', '0', '--tag', 'ref_']
    if algo == 'ar':
        argv += ['--all_reduce', 'True', '--graph_type', '-1']
    elif algo == 'sgp':
        argv += ['--push_sum', 'True', '--graph_type', '5']
    elif algo == 'osgp':
        argv += ['--push_sum', 'True', '--graph_type', '5', '--overlap', 'True']
    elif algo == 'dpsgd':
        argv += ['--push_sum', 'False', '--graph_type', '4']
    else:
        unavailable('unknown algo ' + algo)
    sys.argv = argv
    spec.loader.exec_module(ref)

    warmup, steps, bs = args.warmup, args.steps, args.batch_size
    watch = _Stopwatch(warmup, steps)
    to_device = (algo == 'ar')
    h2d_bytes = bs * 3 * 224 * 224 * 4 + bs * 8

    class SyntheticLoader(object):
        """pinned-host fp32 batches, like DataLoader(pin_memory=True)."""

        def __init__(self, n_batches):
            self.n = n_batches
            g = torch.Generator().manual_seed(1234 + rank)
            self.pool = [(torch.randn(bs, 3, 224, 224, generator=g).pin_memory(),
                          torch.randint(0, 1000, (bs,), generator=g).pin_memory())
                         for _ in range(4)]

        def __len__(self):
            return 5005         # ~ImageNet iterations/epoch at 256; only feeds the LR warm-up

        def __iter__(self):
            for i in range(self.n + 1):
                if watch.mark(i, torch, dist):
                    return
                x, y = self.pool[i % len(self.pool)]
                if to_device:
                    x = x.cuda(non_blocking=True)
                yield x, y

    class _Sampler(object):
        def set_epoch(self, e):
            pass

    def make_dataloader(a, train=True):
        if train:
            return SyntheticLoader(warmup + steps), _Sampler()
        return []          # --train_fast: one (empty) validation pass at the end

    ref.make_dataloader = make_dataloader

    # torch >= 1.7 compatibility: the reference's accuracy() calls .view(-1) on a
    # non-contiguous slice (gossip_sgd.py:486) and raises; same maths with
    # .reshape(-1), as upstream pytorch/examples fixed it.
    def accuracy(output, target, topk=(1,)):
        with torch.no_grad():
            maxk = max(topk)
            batch_size = target.size(0)
            _, pred = output.topk(maxk, 1, True, True)
            pred = pred.t()
            correct = pred.eq(target.view(1, -1).expand_as(pred))
            res = []
            for k in topk:
                correct_k = correct[:k].reshape(-1).float().sum(0, keepdim=True)
                res.append(correct_k.mul_(100.0 / batch_size))
            return res

    ref.accuracy = accuracy
    t_all = time.time()
    ref.main()
    torch.cuda.synchronize()
    if watch.ms is None:
        unavailable('reference loop ended before the timed region completed')

    ms = torch.tensor([watch.ms], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    value = bs * world * steps / (ms / 1e3)
    if rank == 0:
        out = {
            'impl': 'reference', 'metric': 'resnet50_%s_images_per_sec' % args.algo,
            'value': round(value, 2), 'unit': 'images/s', 'n_gpus': world, 'steps': steps,
            'warmup': warmup, 'ms_per_step': round(ms / steps, 4), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32 (cuDNN TF32 conv default)',
            'data': 'synthetic', 'value_is_e2e': True,
            'config': {'model': 'resnet50 (torchvision, reference init_model)',
                       'algorithm': algo + (' (gossip undefined at n=1 in the reference)'
                                            if world == 1 and args.algo != 'ar' else ''),
                       'per_gpu_batch': bs, 'global_batch': bs * world,
                       'image': '3x224x224', 'parallelism': 'dp%d' % world,
                       'l2': 'per-step working set (activations+weights > 1 GB) exceeds L2',
                       'entry': 'baseline/_ref/bin/gossip_sgd.py main()->train(), stock'},
            'e2e': {'value': round(value, 2), 'unit': 'images/s',
                    'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 12},
            'gpu_launches': 0,
            'wall_s': round(time.time() - t_all, 1),
        }
        print(json.dumps(out))
    if dist.is_initialized():
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass
    sys.stdout.flush()
    os._exit(0)       # the reference leaves daemon gossip threads behind
