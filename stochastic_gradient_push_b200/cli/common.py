"""
Shared machinery of the two training entry points (``gossip_sgd.py`` and
``gossip_sgd_adpsgd.py``): flag parsing, environment bootstrap, schedules,
data loaders, CSV logging, accuracy, checkpoint state.

Flag surface = the reference's (``gossip_sgd.py:72-159``,
``gossip_sgd_adpsgd.py:69-144``): same names, same defaults, booleans accepted
as the strings ``'True'/'False'`` so the shipped SLURM job scripts keep working.
Additions are opt-in: ``--synthetic`` (no dataset on the box), ``--fused``
(FusedGossipSGD instead of torch.optim.SGD), ``--amp`` (bf16 autocast),
``--channels_last``, ``--transport``; rank / world / master are also read from
torchrun's ``RANK / WORLD_SIZE / MASTER_ADDR`` when SLURM/OMPI variables are
absent.

The per-rank CSV is byte-compatible with the reference's (4 header lines, the
``Epoch,itr,BT(s),...,val`` column row; ``gossip_sgd.py:262-274``) so
``visualization/plotting.py`` reads either.
"""

from __future__ import annotations

import argparse
import copy
import os
import time

import torch
import torch.distributed as dist

from .. import (GRAPH_TOPOLOGIES, MIXING_STRATEGIES)
from ..experiment import ClusterManager, Meter, get_tcp_interface_name, make_logger


# --------------------------------------------------------------------------- #
# flags
# --------------------------------------------------------------------------- #
def str2bool(v):
    if isinstance(v, bool):
        return v
    if str(v).lower() in ('true', '1', 'yes', 'y'):
        return True
    if str(v).lower() in ('false', '0', 'no', 'n'):
        return False
    raise argparse.ArgumentTypeError('expected True/False, got %r' % (v,))


def build_parser(adpsgd: bool = False) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description='Gossip SGD (B200-native)')
    B = dict(type=str2bool, nargs='?', const=True)
    p.add_argument('--all_reduce', default=False, **B, help='all-reduce instead of gossip')
    p.add_argument('--batch_size', default=32, type=int, help='per-agent batch size')
    p.add_argument('--lr', default=0.1, type=float,
                   help='reference learning rate (for a 256-sample batch)')
    p.add_argument('--num_dataloader_workers', default=10, type=int)
    p.add_argument('--num_epochs', default=90, type=int)
    p.add_argument('--momentum', default=0.9, type=float)
    p.add_argument('--weight_decay', default=1e-4, type=float)
    p.add_argument('--nesterov', default=False, **B)
    p.add_argument('--push_sum', default=True, **B, help='push-sum (SGP) or push-pull (D-PSGD)')
    p.add_argument('--graph_type', default=5, type=int, choices=list(GRAPH_TOPOLOGIES))
    p.add_argument('--mixing_strategy', default=0, type=int, choices=list(MIXING_STRATEGIES))
    p.add_argument('--schedule', nargs='+', default=None, type=float,
                   help='epoch factor epoch factor ... (default 30 0.1 60 0.1 80 0.1)')
    p.add_argument('--peers_per_itr_schedule', nargs='+', type=int,
                   help='epoch num_peers epoch num_peers ...; must contain epoch 0')
    p.add_argument('--overlap', default=False, **B)
    p.add_argument('--synch_freq', default=0, type=int)
    p.add_argument('--warmup', default=False, **B, help='5-epoch linear LR warm-up')
    p.add_argument('--seed', default=47, type=int)
    p.add_argument('--resume', default=False, **B)
    p.add_argument('--backend', default='nccl', choices=['nccl', 'gloo', 'mpi'])
    p.add_argument('--tag', default='', type=str)
    p.add_argument('--print_freq', default=10, type=int)
    p.add_argument('--verbose', default=True, **B)
    p.add_argument('--train_fast', default=False, **B)
    p.add_argument('--checkpoint_all', default=True, **B)
    p.add_argument('--master_port', default='40100', type=str)
    p.add_argument('--checkpoint_dir', type=str, default='./checkpoints/')
    p.add_argument('--network_interface_type', default='infiniband',
                   choices=['infiniband', 'ethernet'])
    p.add_argument('--dataset_dir', type=str, default=None)
    if adpsgd:
        p.add_argument('--bilat', default=True, **B)
        p.add_argument('--shared_fpath', default='', type=str,
                       help='file on a shared FS used as the global iteration counter')
        p.add_argument('--bs_fpath', default='', type=str, help='(accepted, unused; parity)')
    else:
        p.add_argument('--num_iterations_per_training_epoch', default=None, type=int,
                       help='testing only: leave the training loop early')
        p.add_argument('--overwrite_checkpoints', default=True, **B)
        p.add_argument('--num_itr_ignore', type=int, default=10)
        p.add_argument('--no_cuda_streams', action='store_true')
    # ---- B200-native additions (all optional) ----
    p.add_argument('--synthetic', default=None, **B,
                   help='synthetic 3x224x224 data (default: True when --dataset_dir is unset)')
    p.add_argument('--synthetic_len', default=1281167, type=int, help='images per synthetic epoch')
    p.add_argument('--data_format', default='folder', choices=['folder', 'shards'],
                   help="folder: torchvision ImageFolder + DataLoader workers (the reference's pipeline); "
                        "shards: pre-decoded uint8 shards under DATASET_DIR/{train,val} (python -m "
                        "stochastic_gradient_push_b200.data.make_shards) gathered by a background thread, "
                        "crop / flip / normalisation on the GPU (data/shards.py)")
    p.add_argument('--fused', default=True, **B, help='FusedGossipSGD (SGD inside the gossip kernel)')
    p.add_argument('--amp', default=True, **B, help='bf16 autocast for forward/backward')
    p.add_argument('--channels_last', default=True, **B)
    p.add_argument('--transport', default='auto', choices=['auto', 'nvlink', 'c10d'])
    p.add_argument('--nprocs_per_node', default=1, type=int,
                   help='hierarchical mode (GossipDataParallel(nprocs_per_node=K), reference '
                        'gossip/distributed.py:62-80; it has no flag for it): K consecutive ranks form a '
                        'node -- only its first rank gossips, on a graph over the NODES; parameters are '
                        'broadcast and gradients averaged inside the node every iteration (NVLS kernels on '
                        'one NVLink domain).  E.g. N hosts x 8 GPUs: --nprocs_per_node 8')
    p.add_argument('--model', default='resnet50')
    p.add_argument('--num_classes', default=1000, type=int)
    p.add_argument('--image_size', default=224, type=int)
    p.add_argument('--device', default=None, help="'cuda' / 'cpu' (default: cuda if available)")
    p.add_argument('--trace_file', default='', type=str,
                   help='write a Chrome trace (PREFIX_r<rank>.json; NVTX ranges on CUDA) of the '
                        'first --trace_iters training iterations')
    p.add_argument('--trace_iters', default=50, type=int)
    return p


def pairs_to_dict(flat, cast=float):
    """[e0, v0, e1, v1, ...] -> {e0: v0, e1: v1, ...}"""
    flat = list(flat)
    assert len(flat) % 2 == 0, 'schedule needs (epoch, value) pairs'
    return {int(flat[i]): cast(flat[i + 1]) for i in range(0, len(flat), 2)}


def resolve_env(args):
    """rank / world / master from SLURM, OpenMPI or torchrun variables."""
    env = os.environ
    if args.backend == 'mpi' and 'OMPI_COMM_WORLD_RANK' in env:
        args.rank = int(env['OMPI_COMM_WORLD_RANK'])
        args.world_size = int(env.get('OMPI_UNIVERSE_SIZE', env.get('OMPI_COMM_WORLD_SIZE', 1)))
    elif 'SLURM_PROCID' in env and 'RANK' not in env:
        args.rank = int(env['SLURM_PROCID'])
        args.world_size = int(env['SLURM_NTASKS'])
    else:
        args.rank = int(env.get('RANK', 0))
        args.world_size = int(env.get('WORLD_SIZE', 1))
    args.local_rank = int(env.get('LOCAL_RANK', env.get('SLURM_LOCALID', 0)))
    args.master_addr = env.get('MASTER_ADDR') or env.get('HOSTNAME') or '127.0.0.1'
    if 'MASTER_PORT' in env and 'SLURM_PROCID' not in env:
        args.master_port = env['MASTER_PORT']
    return args


def resolve_backend(args) -> str:
    """``--backend mpi`` (reference ``gossip_sgd.py:127-129, 600-602``): ranks come from the
    ``mpirun`` environment (``OMPI_COMM_WORLD_*``, see :func:`resolve_env`).  A PyTorch built with
    MPI uses the MPI process group as in the reference; otherwise MPI stays the *launcher* and the
    control plane is a TCP rendezvous (MASTER_ADDR / MASTER_PORT, exported by the job script) over
    nccl (CUDA) or gloo (CPU).  Either way the data plane on one NVLink domain is the kernel
    transport, which never touches the process group."""
    if args.backend != 'mpi':
        return args.backend
    if dist.is_mpi_available():
        return 'mpi'
    return 'nccl' if (args.device == 'cuda' and dist.is_nccl_available()) else 'gloo'


def finalize_args(args, adpsgd=False):
    """Derived settings + process-group / graph / mixing construction."""
    resolve_env(args)
    if args.checkpoint_dir and not args.checkpoint_dir.endswith('/'):
        args.checkpoint_dir += '/'
    os.makedirs(args.checkpoint_dir, exist_ok=True)
    ClusterManager.set_checkpoint_dir(args.checkpoint_dir)
    args.out_fname = '{}{}out_r{}_n{}.csv'.format(args.checkpoint_dir, args.tag, args.rank,
                                                  args.world_size)
    if args.device is None:
        args.device = 'cuda' if torch.cuda.is_available() else 'cpu'
    if args.device == 'cpu' and args.backend == 'nccl':
        args.backend = 'gloo'
    if args.synthetic is None:
        args.synthetic = args.dataset_dir is None
    args.cpu_comm = (args.backend == 'gloo' and not args.push_sum and not args.all_reduce) \
        or args.device == 'cpu'
    args.comm_device = torch.device('cpu') if args.cpu_comm else torch.device('cuda')
    args.lr_schedule = pairs_to_dict(args.schedule if args.schedule is not None
                                     else [30, 0.1, 60, 0.1, 80, 0.1], float)
    args.ppi_schedule = pairs_to_dict(args.peers_per_itr_schedule
                                      if args.peers_per_itr_schedule is not None else [0, 1], int)
    assert 0 in args.ppi_schedule, 'peers_per_itr_schedule must define epoch 0'
    if args.all_reduce:
        assert args.graph_type == -1, '--all_reduce needs --graph_type -1'

    if args.backend == 'gloo' and args.network_interface_type == 'ethernet':
        try:
            os.environ['GLOO_SOCKET_IFNAME'] = get_tcp_interface_name('ethernet')
        except Exception:
            pass
    elif args.network_interface_type == 'ethernet' and args.backend == 'nccl':
        os.environ['NCCL_SOCKET_IFNAME'] = get_tcp_interface_name('ethernet')
        os.environ['NCCL_IB_DISABLE'] = '1'

    if args.device == 'cuda':
        torch.cuda.set_device(args.local_rank % max(torch.cuda.device_count(), 1))
    os.environ['MASTER_ADDR'] = str(args.master_addr)
    # one control-plane world is enough (the reference's AD-PSGD script needs
    # master_port and master_port+1 because its gossip PROCESS owns a second world)
    os.environ['MASTER_PORT'] = str(args.master_port)
    if not dist.is_initialized() and args.world_size > 1:
        dist.init_process_group(backend=resolve_backend(args), world_size=args.world_size, rank=args.rank)
        dist.barrier()        # create the communicator now, not >5 min apart (reference :678-682)

    args.graph, args.mixing = None, None
    graph_class = GRAPH_TOPOLOGIES[args.graph_type]
    args.graph_class = graph_class
    args.mixing_class = MIXING_STRATEGIES[args.mixing_strategy]
    k = max(1, int(getattr(args, 'nprocs_per_node', 1)))
    args.nprocs_per_node = k
    if k > 1:
        assert not adpsgd and not args.all_reduce, '--nprocs_per_node applies to SGP / OSGP / D-PSGD'
        assert args.world_size % k == 0, '--nprocs_per_node must divide the number of ranks'
    if graph_class is not None and not adpsgd:
        # (hierarchical mode: the graph connects NODES; edges address each node's first rank)
        args.graph = graph_class(args.rank // k, args.world_size // k, nprocs_per_node=k,
                                 local_rank=args.rank % k, peers_per_itr=args.ppi_schedule[0])
        if args.mixing_class is not None:
            args.mixing = args.mixing_class(args.graph, args.comm_device)
    return args


# --------------------------------------------------------------------------- #
# schedules
# --------------------------------------------------------------------------- #
def learning_rate_at(args, epoch, itr=None, itr_per_epoch=None, scale=1):
    """Goyal et al. policy (``gossip_sgd.py:508-536``): target = lr*bs*ws/256,
    optional 5-epoch linear warm-up from ``lr``, then multiplicative steps."""
    target = args.lr * args.batch_size * scale * args.world_size / 256
    if args.warmup and epoch < 5:
        if target <= args.lr:
            return target
        assert itr is not None and itr_per_epoch is not None
        count = epoch * itr_per_epoch + itr + 1
        return args.lr + (target - args.lr) * (count / (5 * itr_per_epoch))
    lr = target
    for e, factor in args.lr_schedule.items():
        if epoch >= e:
            lr *= factor
    return lr


def update_learning_rate(args, optimizer, epoch, itr=None, itr_per_epoch=None, scale=1):
    lr = learning_rate_at(args, epoch, itr, itr_per_epoch, scale)
    for group in optimizer.param_groups:
        group['lr'] = lr
    return lr


def peers_per_itr_at(schedule, epoch):
    """Latest schedule entry whose epoch is <= ``epoch``."""
    best, ppi = -1, None
    for e, v in schedule.items():
        if best <= e <= epoch:
            best, ppi = e, v
    return ppi


def update_peers_per_itr(args, model, epoch):
    model.update_gossiper('peers_per_itr', peers_per_itr_at(args.ppi_schedule, epoch))


# --------------------------------------------------------------------------- #
# data
# --------------------------------------------------------------------------- #
class SyntheticLoader(object):
    """Epoch-length stream of (pinned-host) random batches; a small pool is
    generated once and cycled, like a page-cache-resident dataset."""

    def __init__(self, n_images, batch_size, world_size, rank, image_size=224,
                 num_classes=1000, seed=0, pool=8, pin=True):
        self.batch_size = batch_size
        self.n_batches = max(1, n_images // (batch_size * world_size))
        g = torch.Generator().manual_seed(seed * 1000003 + rank)
        self.pool = []
        for _ in range(min(pool, self.n_batches)):
            x = torch.randn(batch_size, 3, image_size, image_size, generator=g)
            y = torch.randint(0, num_classes, (batch_size,), generator=g)
            if pin and torch.cuda.is_available():
                x, y = x.pin_memory(), y.pin_memory()
            self.pool.append((x, y))
        self.sampler = self

    def set_epoch(self, epoch):
        self._epoch = epoch

    def __len__(self):
        return self.n_batches

    def __iter__(self):
        for i in range(self.n_batches):
            yield self.pool[i % len(self.pool)]


def make_dataloader(args, train=True):
    """(loader, sampler) for training, loader for validation."""
    if args.synthetic:
        n = args.synthetic_len if train else max(args.batch_size * args.world_size * 4, 1)
        loader = SyntheticLoader(n, args.batch_size, args.world_size if train else 1, args.rank,
                                 args.image_size, args.num_classes, seed=args.seed + (0 if train else 1))
        return (loader, loader) if train else loader
    if getattr(args, 'data_format', 'folder') == 'shards':
        from ..data import ShardLoader
        if train:
            loader = ShardLoader(os.path.join(args.dataset_dir, 'train'), args.batch_size, args.world_size,
                                 args.rank, shuffle=True, drop_last=True, seed=args.seed)
            return loader, loader
        return ShardLoader(os.path.join(args.dataset_dir, 'val'), args.batch_size, 1, 0, shuffle=False,
                           drop_last=False, seed=args.seed)
    import torchvision.datasets as datasets
    import torchvision.transforms as transforms
    norm = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    if train:
        ds = datasets.ImageFolder(os.path.join(args.dataset_dir, 'train'), transforms.Compose([
            transforms.RandomResizedCrop(args.image_size), transforms.RandomHorizontalFlip(),
            transforms.ToTensor(), norm]))
        sampler = torch.utils.data.distributed.DistributedSampler(
            ds, num_replicas=args.world_size, rank=args.rank)
        loader = torch.utils.data.DataLoader(
            ds, batch_size=args.batch_size, shuffle=False, sampler=sampler,
            num_workers=args.num_dataloader_workers, pin_memory=True, drop_last=True,
            persistent_workers=args.num_dataloader_workers > 0)
        return loader, sampler
    ds = datasets.ImageFolder(os.path.join(args.dataset_dir, 'val'), transforms.Compose([
        transforms.Resize(int(args.image_size * 256 / 224)), transforms.CenterCrop(args.image_size),
        transforms.ToTensor(), norm]))
    return torch.utils.data.DataLoader(ds, batch_size=args.batch_size, shuffle=False,
                                       num_workers=args.num_dataloader_workers, pin_memory=True)


def device_batch(args, batch, train=True):
    """uint8 shard batches (``--data_format shards``): copy the bytes to the device and run the
    crop / flip / normalisation there (``data.GpuAugment``); anything else passes through."""
    if batch.dtype != torch.uint8:
        return batch
    aug = getattr(args, '_gpu_augment', None)
    if aug is None:
        from ..data import GpuAugment
        aug = args._gpu_augment = GpuAugment(out_size=args.image_size, seed=args.seed * 7919 + args.rank)
    return aug(batch.to(args.device, non_blocking=True), train=train)


# --------------------------------------------------------------------------- #
# metrics / logging / state
# --------------------------------------------------------------------------- #
def accuracy(output, target, topk=(1,)):
    """precision@k in percent as 1-element tensors (no host sync)."""
    with torch.no_grad():
        maxk = max(topk)
        _, pred = output.topk(maxk, 1, True, True)
        hit = pred.eq(target.view(-1, 1))
        return [hit[:, :k].any(dim=1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0))
                for k in topk]


CSV_COLUMNS = ('Epoch,itr,BT(s),avg:BT(s),std:BT(s),NT(s),avg:NT(s),std:NT(s),'
               'DT(s),avg:DT(s),std:DT(s),Loss,avg:Loss,Prec@1,avg:Prec@1,Prec@5,avg:Prec@5,val')


class CSVLog(object):

    def __init__(self, fname, world_size, num_workers, batch_size):
        self.fname = fname
        if not os.path.exists(fname):
            with open(fname, 'w') as f:
                print('BEGIN-TRAINING\nWorld-Size,{}\nNum-DLWorkers,{}\nBatch-Size,{}\n{}'.format(
                    world_size, num_workers, batch_size, CSV_COLUMNS), file=f)

    def train_row(self, epoch, itr, bt, nt, dt, loss, top1, top5):
        with open(self.fname, 'a') as f:
            print('{},{},{},{},{},{:.4f},{:.4f},{:.3f},{:.3f},{:.3f},{:.3f},-1'.format(
                epoch, itr, bt, nt, dt, loss.val, loss.avg, top1.val, top1.avg,
                top5.val, top5.avg), file=f)

    def val_row(self, epoch, bt, nt, dt, prec1):
        with open(self.fname, 'a') as f:
            print('{},-1,{},{},{},-1,-1,-1,-1,-1,-1,{}'.format(epoch, bt, nt, dt, prec1), file=f)


def update_state(state, update_dict):
    for key, value in update_dict.items():
        state[key] = copy.deepcopy(value)


def fresh_state(model_sd, optim_sd):
    return {'epoch': 0, 'itr': 0, 'best_prec1': 0, 'is_best': True, 'state_dict': model_sd,
            'optimizer': optim_sd, 'elapsed_time': 0,
            'batch_meter': Meter(ptag='Time').__dict__, 'data_meter': Meter(ptag='Data').__dict__,
            'nn_meter': Meter(ptag='Forward/Backward').__dict__}


def init_model(args):
    """ResNet (default resnet50) initialised as in "ImageNet in 1 hour"."""
    from .. import models
    kw = {} if args.model == 'tiny' else {'num_classes': args.num_classes}
    net = models.MODEL_ZOO[args.model](**kw)
    if isinstance(net, models.ResNet):
        models.init_imagenet_in_1hr(net)
    net = net.to(args.device)
    if args.channels_last and args.device == 'cuda':
        net = net.to(memory_format=torch.channels_last)
    return net
