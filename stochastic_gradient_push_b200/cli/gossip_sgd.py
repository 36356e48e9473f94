"""
ResNet / ImageNet trainer for AR-SGD, SGP, Overlap-SGP and D-PSGD.

Functional parity with the reference's ``gossip_sgd.py`` (``main`` :163,
``train`` :346, ``validate`` :440, ``accuracy`` :474, ``update_state`` :491,
``update_peers_per_itr`` :497, ``update_learning_rate`` :508,
``make_dataloader`` :539, ``parse_args`` :586, ``init_model`` :693): same flags,
same LR / peers-per-itr schedules, same per-rank CSV, same checkpoint layout.

Launch (one rank per GPU):
    torchrun --nproc-per-node 8 gossip_sgd.py --batch_size 256 --lr 0.1 \
        --num_epochs 90 --nesterov True --warmup True --push_sum True --graph_type 0 \
        --schedule 30 0.1 60 0.1 80 0.1 --checkpoint_dir ./ckpt/ [--dataset_dir /imagenet]
or under SLURM with the scripts in ``job_scripts/`` (SLURM_PROCID / SLURM_NTASKS).

B200-native execution: parameters in a flat arena, SGD fused into the gossip
kernel (``--fused True``), bf16 NHWC compute (``--amp/--channels_last``),
device-side metrics read out once per ``--print_freq`` instead of three host
syncs per iteration, optional whole-step CUDA graph (``--cuda_graph True``).
"""

from __future__ import annotations

import os
import socket
import time

import torch
import torch.distributed as dist
import torch.nn as nn

from . import common
from .common import (CSVLog, Meter, accuracy, build_parser, finalize_args,
                     fresh_state, init_model, make_dataloader, update_learning_rate,
                     update_peers_per_itr, update_state)
from ..experiment import ClusterManager, make_logger
from ..utils import tracing


def parse_args(argv=None):
    parser = build_parser(adpsgd=False)
    parser.add_argument('--cuda_graph', default=False, type=common.str2bool, nargs='?', const=True,
                        help='capture forward+backward+gossip once and replay (fused path only)')
    args = parser.parse_args(argv)
    return finalize_args(args, adpsgd=False)


def build_model_and_optimizer(args, log):
    from ..optim import FusedGossipSGD
    from ..parallel.distributed import GossipDataParallel
    net = init_model(args)
    if args.all_reduce:
        if args.device == 'cuda' and args.transport != 'c10d':
            from ..parallel.allreduce import AllReduceDataParallel
            model = AllReduceDataParallel(net, rank=args.rank, world_size=args.world_size)
            optimizer = _AROptimizer(model, args)
        else:
            model = torch.nn.parallel.DistributedDataParallel(net) if args.world_size > 1 else net
            optimizer = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum,
                                        weight_decay=args.weight_decay, nesterov=args.nesterov)
        return model, optimizer
    model = GossipDataParallel(
        net, graph=args.graph, mixing=args.mixing, comm_device=args.comm_device,
        push_sum=args.push_sum, overlap=args.overlap, synch_freq=args.synch_freq,
        verbose=args.verbose, use_streams=not args.no_cuda_streams, rank=args.rank,
        world_size=args.world_size, transport=args.transport, nprocs_per_node=args.nprocs_per_node,
        # the captured fast path trains through the bf16 shadow-weight twin (no autocast casts)
        compute_dtype=(torch.bfloat16 if (args.cuda_graph and args.amp and args.fused
                                          and args.device == 'cuda' and args.transport != 'c10d')
                       else None))
    if args.fused:
        optimizer = FusedGossipSGD(model, lr=args.lr, momentum=args.momentum,
                                   weight_decay=args.weight_decay, nesterov=args.nesterov)
    else:
        optimizer = torch.optim.SGD(model.parameters(), lr=args.lr, momentum=args.momentum,
                                    weight_decay=args.weight_decay, nesterov=args.nesterov)
    log.info('transport: %s, fused optimizer: %s' % (model.transport, args.fused))
    return model, optimizer


class _AROptimizer(object):
    """optimizer facade over AllReduceDataParallel's fused all-reduce+SGD kernel"""

    def __init__(self, model, args):
        self.model = model
        self.param_groups = [dict(lr=args.lr, momentum=args.momentum,
                                  weight_decay=args.weight_decay, nesterov=args.nesterov)]

    def step(self):
        g = self.param_groups[0]
        self.model.set_hyper(g['lr'], g['momentum'], g['weight_decay'], g['nesterov'])
        self.model.allreduce_step()

    def zero_grad(self, set_to_none=False):
        pass          # the step clears the (symmetric) gradient buffer

    def state_dict(self):
        return {'param_groups': self.param_groups, 'momentum': self.model.momentum.cpu()}

    def load_state_dict(self, sd):
        self.param_groups[0].update(sd['param_groups'][0])
        self.model.momentum.copy_(sd['momentum'])


def main(argv=None):
    args = parse_args(argv)
    log = make_logger(args.rank, args.verbose)
    log.info('args: {}'.format({k: v for k, v in vars(args).items()
                                if k not in ('graph', 'mixing')}))
    log.info(socket.gethostname())
    torch.manual_seed(args.seed)
    if args.device == 'cuda':
        torch.cuda.manual_seed(args.seed)
        torch.backends.cudnn.benchmark = True

    if args.trace_file:
        tracing.enable(args.trace_file, rank=args.rank)
    model, optimizer = build_model_and_optimizer(args, log)
    criterion = nn.CrossEntropyLoss()      # == KLDiv(log_softmax, one_hot) of the reference
    optimizer.zero_grad()
    args._trainer = None
    if args.cuda_graph and not args.all_reduce and args.fused and getattr(model, '_kernel', None):
        from ..parallel.trainer import GossipTrainer
        # criterion=None: the trainer's fused softmax-xent + prec@1/5 kernel (same value as
        # nn.CrossEntropyLoss; tests/test_fused_loss_gpu.py)
        args._trainer = GossipTrainer(model, optimizer, None,
                                      amp_dtype=torch.bfloat16 if args.amp else None)

    state = fresh_state(model.state_dict(), optimizer.state_dict())
    cmanager = ClusterManager(rank=args.rank, world_size=args.world_size, model_tag=args.tag,
                              state=state, all_workers=args.checkpoint_all)
    if args.resume and os.path.isfile(cmanager.checkpoint_fpath):
        log.info("=> loading checkpoint '{}'".format(cmanager.checkpoint_fpath))
        ckpt = torch.load(cmanager.checkpoint_fpath, map_location='cpu', weights_only=False)
        update_state(state, {k: ckpt[k] for k in (
            'epoch', 'itr', 'best_prec1', 'state_dict', 'optimizer', 'elapsed_time',
            'batch_meter', 'data_meter', 'nn_meter')})
        state['is_best'] = False
        model.load_state_dict(ckpt['state_dict'])
        optimizer.load_state_dict(ckpt['optimizer'])
        log.info("=> loaded checkpoint (epoch {}; itr {})".format(ckpt['epoch'], ckpt['itr']))
    elif args.resume:
        log.info("=> no checkpoint found at '{}'".format(cmanager.checkpoint_fpath))

    batch_meter = Meter(state['batch_meter'])
    data_meter = Meter(state['data_meter'])
    nn_meter = Meter(state['nn_meter'])
    csv = CSVLog(args.out_fname, args.world_size, args.num_dataloader_workers, args.batch_size)

    loader, sampler = make_dataloader(args, train=True)
    val_loader = None if args.train_fast else make_dataloader(args, train=False)

    start_itr, start_epoch = state['itr'], state['epoch']
    elapsed_time = state['elapsed_time']
    begin_time = time.time() - elapsed_time
    best_val_prec1 = state.get('best_prec1', 0)
    for epoch in range(start_epoch, args.num_epochs):
        sampler.set_epoch(epoch + args.seed * 90)
        if not args.all_reduce:
            update_peers_per_itr(args, model, epoch)
            model.block()
        train(args, model, criterion, optimizer, batch_meter, data_meter, nn_meter, loader,
              epoch, start_itr, csv, log)
        start_itr = 0
        if not args.train_fast:
            elapsed_time = time.time() - begin_time
            update_state(state, {
                'epoch': epoch + 1, 'itr': 0, 'is_best': False,
                'state_dict': model.state_dict(), 'optimizer': optimizer.state_dict(),
                'elapsed_time': elapsed_time, 'batch_meter': batch_meter.__dict__,
                'data_meter': data_meter.__dict__, 'nn_meter': nn_meter.__dict__})
            prec1 = validate(args, val_loader, model, criterion, log)
            csv.val_row(epoch, batch_meter, nn_meter, data_meter, prec1)
            if prec1 > best_val_prec1:
                update_state(state, {'is_best': True, 'best_prec1': prec1})
                best_val_prec1 = prec1
            epoch_id = None if args.overwrite_checkpoints else epoch
            cmanager.save_checkpoint(epoch_id, requeue_on_signal=(epoch != args.num_epochs - 1))

    if args.train_fast:
        val_loader = make_dataloader(args, train=False)
        prec1 = validate(args, val_loader, model, criterion, log)
        log.info('Test accuracy: {}'.format(prec1))
    log.info('elapsed_time {0}'.format(time.time() - begin_time))
    if tracing.get_tracer().enabled:            # shorter run than --trace_iters
        log.info('trace written to %s' % tracing.disable().dump())
    if dist.is_initialized():
        dist.barrier()
    return state


def _autocast(args):
    return torch.autocast('cuda', dtype=torch.bfloat16, enabled=bool(args.amp and args.device == 'cuda'))


def _drain_metrics(trainer, pending):
    """[loss, prec@1, prec@5] of every pending iteration with one synchronisation.  Graph path:
    the rows already sit in the trainer's pinned ring (slots); eager path: device tensors.
    Also the place where the gossip kernels' health word is polled (a peer that stopped
    publishing raises here, like the reference's 'Gossip flag timeout')."""
    if trainer is not None:
        trainer.stream.synchronize()
        trainer.check()
        return [trainer.metrics_ring[slot].tolist() for slot, _ in pending]
    return torch.stack([v for v, _ in pending]).cpu().tolist()


def train(args, model, criterion, optimizer, batch_meter, data_meter, nn_meter, loader, epoch,
          itr, csv, log):
    losses, top1, top5 = Meter(ptag='Loss'), Meter(ptag='Prec@1'), Meter(ptag='Prec@5')
    model.train()
    dev = args.device
    ignore = getattr(args, 'num_itr_ignore', 0) if epoch == 0 else 0
    limit = getattr(args, 'num_iterations_per_training_epoch', None)
    # device-side metric accumulators: [loss, prec1, prec5] per iteration, read
    # back once per print interval (the reference syncs 3x per iteration)
    pending = []
    trainer = getattr(args, '_trainer', None)
    it = iter(loader)
    for _ in range(itr):            # resume mid-epoch: skip what was already consumed
        next(it, None)
    t_batch = time.time()
    traced = 0
    for i, (batch, target) in enumerate(it, start=itr):
        if tracing.get_tracer().enabled:
            traced += 1
            if traced > args.trace_iters:       # bounded trace: dump once, then stop recording
                log.info('trace written to %s' % tracing.disable().dump())
        target = target.to(dev, non_blocking=True)
        batch = common.device_batch(args, batch, train=True)       # (uint8 shard batches only)
        if dev == 'cuda' and not batch.is_cuda and args.all_reduce:
            batch = batch.to(dev, non_blocking=True)
        t_data = time.time() - t_batch
        tracing.counter('data_ms', t_data * 1e3)
        t_nn = time.time()
        if i % 100 == 0:
            update_learning_rate(args, optimizer, epoch, itr=i, itr_per_epoch=len(loader))
        if trainer is not None:
            # whole step = one CUDA-graph replay (forward+backward+fused gossip kernel);
            # the metrics are computed on the trainer's stream, before the next replay
            # can overwrite the static output buffers
            # (loss, prec@1, prec@5 come out of the fused loss kernel inside the graph and land
            # in the trainer's pinned metrics ring: 12 bytes per step, no host sync)
            with tracing.span('step[graph]', itr=i):
                slot = trainer.step(batch, target)
            pending.append((slot, batch.size(0)))
        else:
            with tracing.span('forward', itr=i), _autocast(args):
                output = model(batch)
                loss = criterion(output.float(), target)
            with tracing.span('backward'):
                loss.backward()
            with tracing.span('optimizer'):
                optimizer.step()
                optimizer.zero_grad()
            if not args.overlap and not args.all_reduce:
                model.transfer_params()
            with torch.no_grad():
                p1, p5 = accuracy(output, target, topk=(1, 5))
                pending.append((torch.stack([loss.detach().float().reshape(()), p1[0], p5[0]]),
                                batch.size(0)))
        t_nn = time.time() - t_nn
        if ignore == 0:
            data_meter.update(t_data)
            nn_meter.update(t_nn)
            batch_meter.update(time.time() - t_batch)
        else:
            ignore -= 1
        t_batch = time.time()

        last = (limit not in (None, -1) and i + 1 == limit)
        if i % args.print_freq == 0 or last or len(pending) >= 512:
            vals = _drain_metrics(trainer, pending)                # ONE sync per interval
            for (l, a1, a5), (_, n) in zip(vals, pending):
                losses.update(l, n)
                top1.update(a1, n)
                top5.update(a5, n)
            pending = []
            csv.train_row(epoch, i, batch_meter, nn_meter, data_meter, losses, top1, top5)
        if last:
            break
    if trainer is not None:
        # end of epoch: land the deferred SGD / gathered residual (overlap) so that validation
        # and state_dict() see every update; eval / checkpointing run on the default stream
        trainer.finish()
    if pending:
        vals = _drain_metrics(trainer, pending)
        for (l, a1, a5), (_, n) in zip(vals, pending):
            losses.update(l, n)
            top1.update(a1, n)
            top5.update(a5, n)
    csv.train_row(epoch, i, batch_meter, nn_meter, data_meter, losses, top1, top5)
    return losses.avg


def validate(args, val_loader, model, criterion, log):
    losses, top1, top5 = Meter(ptag='Loss'), Meter(ptag='Prec@1'), Meter(ptag='Prec@5')
    model.eval()
    dev = args.device
    with torch.no_grad():
        for features, target in val_loader:
            target = target.to(dev, non_blocking=True)
            features = common.device_batch(args, features, train=False).to(dev, non_blocking=True)
            if args.channels_last and dev == 'cuda':
                features = features.contiguous(memory_format=torch.channels_last)
            with _autocast(args):
                output = model(features)
                loss = criterion(output.float(), target)
            p1, p5 = accuracy(output, target, topk=(1, 5))
            n = features.size(0)
            losses.update(loss.item(), n)
            top1.update(p1.item(), n)
            top5.update(p5.item(), n)
    log.info(' * Prec@1 {top1.avg:.3f} Prec@5 {top5.avg:.3f}'.format(top1=top1, top5=top5))
    return top1.avg


if __name__ == '__main__':
    main()
