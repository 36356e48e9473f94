"""
BilatGossipDataParallel: asynchronous bilateral gossip (AD-PSGD).

API parity with ``gossip/ad_psgd.py:36-418``: same constructor arguments,
``forward``, ``update_lr``, ``enable_gossip`` / ``disable_gossip``, ``block``
(no-op), ``sync_comms``, ``train`` / ``eval``, ``_pull_model``,
``_transfer_grads``, ``communicator_warmup``.

The reference forks a separate gossip *process* that owns a second
``torch.distributed`` world, a shared-memory (CUDA-IPC / pinned) copy of the
model + gradients and its own ``torch.optim.SGD``, and loops forever
{apply new gradients; bilateral average with the next peer}.  On B200 none of
that needs a process:

* the "gossip copy" is a second flat arena on the same GPU; gradients reach it
  with one flat D2D copy, the model is pulled back with one flat D2D copy;
* the gossip-side optimizer is the fused SGD kernel (``engine.local``);
* the bilateral average is the pull kernel over NVSwitch peer memory, gated by
  a device-side flag handshake: ``publish_only`` (snapshot + release flags),
  ``probe`` (has my partner published this round?  -- the reference's
  ``_pending_req.is_completed()``), ``pull_only`` (x <- (x + x_partner)/2, ack).
  Active ranks publish unconditionally, passive ranks only once their partner's
  snapshot is visible (``gossip/gossiper.py:290-316``); no kernel ever spins;
* a daemon *thread* drives that loop on a dedicated low-priority stream with a
  small grid, so training kernels keep the SMs.

On CPU tensors / gloo the same loop runs over ``BilatPushPull`` (isend/irecv on
a dedicated process group) with plain torch SGD -- the oracle for the kernels.
"""

from __future__ import annotations

import threading
import time

import torch
import torch.distributed as dist
from torch.autograd import Variable
from torch.nn.modules import Module

from ..gossiper import BilatPushPull, C10dTransport
from ..ops import oracle
from ..utils.arena import FlatArena
from ..utils.helpers import make_logger
from ..utils.metering import Meter


class BilatGossipDataParallel(Module):
    """Distributed bilateral-gossip model wrapper (AD-PSGD)."""

    def __init__(self, module, device_ids=None, master_addr=None, master_port=None,
                 backend=None, world_size=None, rank=None, graph_class=None,
                 mixing_class=None, num_peers=1, comm_device=None, lr=0.1, momentum=0.9,
                 weight_decay=1e-4, nesterov=True, verbose=True,
                 network_interface_type=None, tcp_interface_name=None,
                 transport='auto', poll_interval=2e-4, gossip_grid=32,
                 heartbeat_timeout=300.0, max_rounds_per_update=4):
        super(BilatGossipDataParallel, self).__init__()
        first = next(module.parameters())
        on_cuda = first.is_cuda
        if device_ids is None:
            device_ids = [first.device.index] if on_cuda else []
        self.device_ids = list(device_ids)
        self.output_device = self.device_ids[0] if self.device_ids else None
        assert len(self.device_ids) <= 1, \
            'one process per GPU: wrap each replica in its own rank'
        self.module = module
        self._module_copies = [self.module]

        # control plane: reuse the caller's process group, or create the one the
        # reference's gossip process would have created (ad_psgd.py:280-284)
        if not dist.is_initialized() and world_size is not None and world_size > 1:
            import os
            if master_addr is not None:
                os.environ['MASTER_ADDR'] = str(master_addr)
            if master_port is not None:
                os.environ['MASTER_PORT'] = str(master_port)
            dist.init_process_group(backend=backend or ('nccl' if on_cuda else 'gloo'),
                                    world_size=world_size, rank=rank)
        if rank is None or world_size is None:
            if dist.is_initialized():
                rank, world_size = dist.get_rank(), dist.get_world_size()
            else:
                rank, world_size = 0, 1

        if comm_device is None:
            comm_device = first.device if on_cuda else torch.device('cpu')
        comm_device = torch.device(comm_device)
        self.__cpu_comm = comm_device.type == 'cpu'
        self.dist_config = {
            'verbose': verbose, 'graph_class': graph_class, 'master_addr': master_addr,
            'master_port': master_port, 'backend': backend, 'world_size': world_size,
            'rank': rank, 'mixing_class': mixing_class, 'lr': lr, 'momentum': momentum,
            'nesterov': nesterov, 'weight_decay': weight_decay, 'comm_device': comm_device,
            'network_interface_type': network_interface_type, 'num_peers': num_peers,
        }
        self.num_updates = 0
        self.logger = make_logger(rank, verbose)
        self.gossip_enable = True
        self._poll = float(poll_interval)
        # the reference's gossip process averages as fast as it can, even when no new
        # gradient arrived; on NVLink that is ~18 rounds (1.8 GB of pulls) per training
        # step competing with the model for HBM and SMs.  Bound it (None = unbounded).
        self.max_rounds_per_update = max_rounds_per_update
        self._rounds_since_update = 0
        self._timeout_s = float(heartbeat_timeout)

        # graph / mixing are given as CLASSES (instantiated here, like the
        # reference instantiates them inside its gossip process)
        from ..topology.graph_manager import DynamicBipartiteExponentialGraph
        from ..mixing_manager import UniformMixing
        graph_class = graph_class or DynamicBipartiteExponentialGraph
        mixing_class = mixing_class or UniformMixing
        self.graph = graph_class(rank, world_size, peers_per_itr=num_peers)
        self.mixing = mixing_class(self.graph, comm_device)
        self.dist_config['graph'] = self.graph
        self.dist_config['mixing'] = self.mixing

        # train copy: flat arena (+ flat gradient)
        params = list(module.parameters())
        assert all(p.dtype == first.dtype for p in params), 'single-dtype models only'
        self.arena = FlatArena(params, device=first.device)
        self.arena.adopt(params)
        self.grad_flat = self.arena.new_buffer()
        self.arena.bind_grads(params, self.grad_flat)

        # gossip copy + its gradient / momentum buffers
        self.gossip_flat = self.arena.new_buffer()
        self.gossip_flat.copy_(self.arena.flat)
        self.gossip_params = self.arena.views_of(self.gossip_flat)
        self.gossip_grad_flat = self.arena.new_buffer()
        self.gossip_grads = self.arena.views_of(self.gossip_grad_flat)
        self.momentum_flat = self.arena.new_buffer()

        self.gossip_lock = threading.Lock()
        self.gossip_enable_flag = threading.Event()
        self.train_write_flag = threading.Event()    # train thread wrote new grads
        self.gossip_read_flag = threading.Event()    # gossip thread consumed them
        self.gossip_update_flag = threading.Event()  # learning rate changed
        self._stop = threading.Event()
        self._lr = float(lr)
        self.rounds_completed = 0
        self.grads_applied = 0
        self._error = None

        use_kernels = False
        if transport == 'auto':
            from .distributed import _native_ok, _single_nvlink_domain
            use_kernels = on_cuda and not self.__cpu_comm and first.dtype == torch.float32 \
                and _native_ok()
            # kernels need one NVLink domain (all ranks on one host); else the c10d loop
            use_kernels = _single_nvlink_domain(world_size, use_kernels)
        elif transport in ('nvlink', 'kernel', 'peer'):
            use_kernels = True
        self.transport = 'nvlink' if use_kernels else 'c10d'

        self.engine = None
        self.gossiper = None
        if use_kernels:
            from ..ops.peer_mix import GossipEngine
            from .symmetric import LocalWorld, SymmetricWorld
            sw = SymmetricWorld(first.device) if world_size > 1 \
                else LocalWorld(1, [first.device.index]).view(0)
            self.engine = GossipEngine(sw, self.gossip_flat, self.graph, self.mixing,
                                       grad=self.gossip_grad_flat, momentum=self.momentum_flat,
                                       grid=gossip_grid, timeout_s=self._timeout_s, name='adpsgd')
            lo, hi = torch.cuda.Stream.priority_range()
            self.gossip_stream = torch.cuda.Stream(device=first.device, priority=lo)
            self._host_flag = self.engine.C.pinned_flag()
            self.engine.set_hyper(lr, momentum, weight_decay, nesterov)
            torch.cuda.synchronize(first.device)
        else:
            self.gossip_stream = None
            group = None
            if dist.is_initialized() and world_size > 1:
                group = dist.new_group(list(range(world_size)))   # gossip-only channel
            self._gossip_group = group
            if world_size > 1:
                self.gossiper = BilatPushPull(self.gossip_flat, graph=self.graph, mixing=self.mixing,
                                              rank=rank, world_size=world_size,
                                              transport=C10dTransport(group), logger=None)

        self.model_meter = Meter(ptag='Model', stateful=True, csv_format=False)
        self.gossip_meter = Meter(ptag='Gossip', stateful=True, csv_format=False)
        self.gossip_read_flag.set()
        self.gossip_thread = threading.Thread(target=self._gossip_target, daemon=True,
                                              name='Gossip-Thread')
        self.gossip_thread.start()
        self.__register_hooks()

    # ------------------------------------------------------------------ #
    # public API
    # ------------------------------------------------------------------ #
    def update_lr(self, lr):
        if self._lr == lr:
            return
        self._lr = float(lr)
        self.gossip_update_flag.set()

    def forward(self, *inputs, **kwargs):
        if self.device_ids:
            from torch.nn.parallel.scatter_gather import scatter_kwargs
            inputs, kwargs = scatter_kwargs(inputs, kwargs, self.device_ids, dim=0)
            return self.module(*inputs[0], **kwargs[0])
        return self.module(*inputs, **kwargs)

    def train(self, mode=True):
        super(BilatGossipDataParallel, self).train(mode)
        return self

    def eval(self):
        super(BilatGossipDataParallel, self).eval()
        self._pull_model()
        return self

    def enable_gossip(self):
        self.gossip_enable = True
        self.gossip_enable_flag.set()

    def disable_gossip(self):
        self.gossip_enable = False
        self.gossip_enable_flag.clear()

    def block(self):
        return          # the reference's barrier is unreachable too (ad_psgd.py:212-215)

    def sync_comms(self):
        self._pull_model()

    def communicator_warmup(self):
        if dist.is_initialized():
            dist.barrier()
            time.sleep(0.1)
            dist.barrier()

    def shutdown(self):
        self._stop.set()
        self.gossip_enable_flag.set()
        self.gossip_thread.join(timeout=10)

    # ------------------------------------------------------------------ #
    # train-thread <-> gossip-thread hand-offs
    # ------------------------------------------------------------------ #
    def _check(self):
        if self._error is not None:
            raise RuntimeError('gossip thread died: %r' % (self._error,))

    def _pull_model(self):
        """train copy <- gossip copy (one flat copy under the gossip lock)."""
        self._check()
        with self.gossip_lock:
            if self.engine is not None:
                cur = torch.cuda.current_stream(self.arena.flat.device)
                cur.wait_stream(self.gossip_stream)
            self.arena.flat.copy_(self.gossip_flat, non_blocking=False)
            if self.engine is not None:
                torch.cuda.current_stream(self.arena.flat.device).synchronize()
        return True

    def _transfer_grads(self):
        """gossip-side gradient buffer <- this step's gradients."""
        self._check()
        if not self.gossip_read_flag.wait(timeout=self._timeout_s):
            raise RuntimeError('gossip thread did not consume the previous gradients')
        params = list(self.module.parameters())
        g0 = params[0].grad
        if g0 is not None and g0.data_ptr() == self.grad_flat.data_ptr():
            self.gossip_grad_flat.copy_(self.grad_flat, non_blocking=False)   # one flat copy
        else:
            # an optimizer replaced / dropped the flat views (zero_grad(set_to_none=True)):
            # per-tensor copies, then re-bind the views for the next iteration
            for p, g in zip(params, self.gossip_grads):
                if p.requires_grad and p.grad is not None:
                    g.copy_(p.grad)
                else:
                    g.zero_()
        if self.engine is not None:
            torch.cuda.current_stream(self.arena.flat.device).synchronize()
        self.gossip_read_flag.clear()
        self.train_write_flag.set()
        return True

    # ------------------------------------------------------------------ #
    # gossip thread
    # ------------------------------------------------------------------ #
    def _gossip_target(self):
        try:
            if self.engine is not None:
                torch.cuda.set_device(self.arena.flat.device)
                with torch.cuda.stream(self.gossip_stream):
                    self._loop_kernels()
            else:
                with torch.no_grad():
                    self._loop_c10d()
        except Exception as e:           # surfaced to the train thread
            self._error = e
            self.gossip_read_flag.set()

    def _apply_pending(self, cfg):
        """learning-rate updates and fresh gradients -> gossip-side SGD step"""
        if self.gossip_update_flag.is_set():
            cfg['lr'] = self._lr
            if self.engine is not None:
                self.engine.set_hyper(cfg['lr'], cfg['momentum'], cfg['weight_decay'],
                                      cfg['nesterov'])
            self.gossip_update_flag.clear()
        if self.train_write_flag.is_set():
            bt = time.time()
            with self.gossip_lock:
                if self.engine is not None:
                    self.engine.local(sgd=True, zero_grad=False)
                    self.gossip_stream.synchronize()
                else:
                    x, m = oracle.sgd_momentum(self.gossip_flat, self.gossip_grad_flat,
                                               self.momentum_flat, cfg['lr'], cfg['momentum'],
                                               cfg['weight_decay'], cfg['nesterov'])
                    self.gossip_flat.copy_(x)
                    self.momentum_flat.copy_(m)
            self.grads_applied += 1
            self._rounds_since_update = 0
            self.train_write_flag.clear()
            self.gossip_read_flag.set()
            self.model_meter.update(time.time() - bt)

    def _throttled(self, round_in_flight):
        """True when this rank has done enough rounds since its last gradient and
        is not in the middle of one (a round in flight is always completed)."""
        k = self.max_rounds_per_update
        return (k is not None) and (not round_in_flight) and (self._rounds_since_update >= k) \
            and self.training

    def _loop_kernels(self):
        cfg = dict(self.dist_config)
        e = self.engine
        passive = self.graph.is_passive()
        alone = self.dist_config['world_size'] < 2
        published = False
        while not self._stop.is_set():
            if not self.gossip_enable_flag.wait(timeout=0.05):
                continue
            self._apply_pending(cfg)
            if alone or self._throttled(published):
                time.sleep(self._poll)
                continue
            bt = time.time()
            e.probe(self._host_flag)
            self.gossip_stream.synchronize()
            ready = int(self._host_flag[0]) == 1        # partner's snapshot is visible
            may_publish = int(self._host_flag[1]) == 1  # our outbox buffer has been released
            # active ranks publish unconditionally, passive ranks only once their
            # partner showed up; never before the readers of round r-2 have acked
            if not published and may_publish and (ready or not passive):
                with self.gossip_lock:
                    e.publish_only()
                published = True
            if ready and published:
                with self.gossip_lock:
                    e.pull_only()
                    self.gossip_stream.synchronize()
                e.check()
                published = False
                self.rounds_completed += 1
                self._rounds_since_update += 1
                self.gossip_meter.update(time.time() - bt)
            else:
                time.sleep(self._poll)

    def _loop_c10d(self):
        """Same protocol over isend/irecv.  "Has my partner published?" is a key in
        the c10d store (the analogue of the probe kernel): a receive is only posted
        once the matching send exists, so no rank ever blocks on -- or exits with --
        a dangling receive, whatever its partner is doing (finished, validating,
        slow).  The reference's active rank blocks inside ``mix`` here
        (``gossip/gossiper.py:290-299``)."""
        cfg = dict(self.dist_config)
        alone = self.gossiper is None
        g = self.gossiper
        tr = g.transport if g is not None else None
        passive = self.graph.is_passive()
        store = None
        if not alone:
            from torch.distributed.distributed_c10d import _get_default_store
            store = _get_default_store()
        rnd = 0
        sent = None
        inflight = []
        wait = self._poll
        t_round = time.time()

        def key(r, src, dst):
            return 'adpsgd/%d/%d>%d' % (r, src, dst)

        def publish():
            with self.gossip_lock:
                snap = self.gossip_flat.clone()
            out = g.out_edges[0]
            req = tr.post_sends([snap], [out])[0]
            store.set(key(rnd, out.src, out.dest), '1')
            return (req, snap)

        while not self._stop.is_set():
            if not self.gossip_enable_flag.wait(timeout=0.05):
                continue
            self._apply_pending(cfg)
            if alone or self._throttled(sent is not None):
                time.sleep(self._poll)
                continue
            if sent is None:
                t_round = time.time()
                if not passive:
                    sent = publish()
            in_edge = g.in_edges[0]
            in_key = key(rnd, in_edge.src, in_edge.dest)
            if not store.check([in_key]):
                # back off exponentially (x1.5 up to 20 ms): a quiet partner must not cost the
                # store on rank 0 a constant 5 kHz of polls from every rank
                time.sleep(wait)
                wait = min(wait * 1.5, 0.02)
                continue
            wait = self._poll
            try:
                store.delete_key(in_key)       # consumed: the store does not grow with the run
            except Exception:                  # (a store without delete support keeps the key)
                pass
            if sent is None:          # passive: answer now that the partner showed up
                sent = publish()
            tr.post_recvs([g.in_msg_buffer], [in_edge])[0].wait()
            with self.gossip_lock:
                self.gossip_flat.add_(g.in_msg_buffer.to(self.gossip_flat.device)).mul_(0.5)
            # never block on OUR send: the partner may have gone quiet before
            # posting its receive.  A send is certainly complete once the partner
            # has answered two later rounds, so only those are reaped.
            inflight.append(sent)
            while len(inflight) > 2:
                inflight.pop(0)[0].wait()
            sent = None
            g.refresh_peers_()
            rnd += 1
            self.rounds_completed += 1
            self._rounds_since_update += 1
            self.gossip_meter.update(time.time() - t_round)

    # ------------------------------------------------------------------ #
    def __register_hooks(self):
        self.register_full_backward_pre_hook(self.__make_backward_hook())

    def __make_backward_hook(self):
        def hook(*unused):
            self._transfer_grads()
            self._pull_model()

        def queue_hook(*unused):
            Variable._execution_engine.queue_callback(hook)
        return queue_hook
