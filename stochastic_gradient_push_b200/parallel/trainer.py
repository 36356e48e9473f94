"""
GossipTrainer: the flagship training step, captured ONCE as a CUDA graph.

One replay = forward (bf16 autocast, NHWC) + loss + backward + ONE fused
gossip kernel (SGD-momentum + publish + P2P pull + mix + de-bias), i.e. the
reference's whole hot loop (``gossip_sgd.py:369-393`` + the pre-forward /
backward hooks + the gossip thread) with zero Python on the critical path:

* the time-varying graph is a device table the kernel indexes with its own
  device-side step counter; the learning rate lives in device memory -- so the
  captured graph never needs re-capturing;
* Overlap-SGP forks the gather kernel onto ``gossip_stream`` INSIDE the graph:
  it pulls the peers' parameters over NVLink while forward/backward run;
* inputs arrive from pinned host memory through a prefetch stream (H2D of step
  k+1 overlaps compute of step k); the loss is read back through a pinned ring
  without a per-step host synchronisation.

Modes: ``'sgp'`` (push-sum, directed graph), ``'dpsgd'`` (push-pull, symmetric
graph) -- same kernel, the graph/mixing differ --, ``'osgp'`` (overlap), and
``'local'`` (world size 1 / gossip disabled).
"""

from __future__ import annotations

import contextlib
import os
from typing import Callable, Optional

import torch

from ..ops.fused_loss import FusedCrossEntropyWithAccuracy
from ..optim import FusedGossipSGD
from .distributed import GossipDataParallel


class GossipTrainer(object):

    def __init__(self, model: GossipDataParallel, optimizer: FusedGossipSGD,
                 criterion: Optional[Callable] = None, amp_dtype=torch.bfloat16,
                 use_cuda_graph: bool = True, warmup_iters: int = 3,
                 channels_last: bool = True):
        assert model._kernel is not None, 'GossipTrainer drives the nvlink kernel transport'
        self.model = model
        self.opt = optimizer
        self.engine = model._kernel.engine
        self.k = model._kernel
        self.overlap = model.overlap
        self.gossip = model.dist_config['world_size'] > 1
        self._init_runtime(self.engine.device, criterion, amp_dtype, use_cuda_graph, warmup_iters,
                           channels_last)

    def _init_runtime(self, device, criterion, amp_dtype, use_cuda_graph, warmup_iters, channels_last):
        """state shared by every trainer flavour (gossip / all-reduce / bilateral)"""
        # default: fused softmax cross-entropy that also yields prec@1 / prec@5 (the reference loop
        # measures both every iteration, gossip_sgd.py:394-399) in the same launch
        self.criterion = criterion or FusedCrossEntropyWithAccuracy()
        self._fused_loss = isinstance(self.criterion, FusedCrossEntropyWithAccuracy)
        self.amp_dtype = amp_dtype
        self.use_graph = use_cuda_graph
        self.warmup_iters = warmup_iters
        self.channels_last = channels_last
        self.device = device
        self.graph = None
        self.static_in = None
        self.static_tgt = None
        self.static_loss = None
        self.static_metrics = None
        self.static_out = None
        self._eager_steps = 0
        # see _backward(); SGP_B200_BATCHED_GRAD_COPY=0 restores per-parameter accumulation
        self.batched_grad_copy = os.environ.get('SGP_B200_BATCHED_GRAD_COPY', '1') != '0'
        self._grad_slots = None
        # every launch of the step runs on ONE dedicated side stream: autograd's
        # AccumulateGrad nodes are then created on the stream the graph is captured on
        self.stream = torch.cuda.Stream(device=self.device)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._stage = None
        self._stage_tgt = None
        self._stage_ready = torch.cuda.Event()
        self._stage_free = torch.cuda.Event()
        self._prefetched = False
        self._loss_ring = None
        self._loss_slot = 0
        self._skip_next_sgd = False
        # kernels of OUR extension in one training step (counted over the graph capture, or over
        # the last eager step): what bench.py reports as gpu_launches / step
        self.own_launches_per_step = None
        self._graphs = {}               # overlap + DMA gather: one captured graph per (row, parity)
        self._pool = None

    # ------------------------------------------------------------------ #
    def _autocast(self):
        if self.amp_dtype is None or self.amp_dtype == torch.float32:
            return contextlib.nullcontext()
        return torch.autocast('cuda', dtype=self.amp_dtype)

    def _fwd_bwd(self):
        twin = getattr(self.model, '_twin', None)
        net = None
        if twin:
            # bf16 twin: weights already live in bf16 (shadow arena), input buffer is bf16
            net = twin[0]
            net.train(self.model.module.training)
            out = net(self.static_in)
            loss = self.criterion(out if self._fused_loss else out.float(), self.static_tgt)
        else:
            with self._autocast():
                out = self.model.module(self.static_in)
                loss = self.criterion(out if self._fused_loss else out.float(), self.static_tgt)
        self._backward(loss, net if twin else self.model.module)
        metrics = getattr(self.criterion, 'metrics', None)
        if metrics is not None:
            self.static_metrics.copy_(metrics)          # [loss, prec@1 %, prec@5 %]
        else:
            self.static_metrics[0].copy_(loss.detach())
        self.static_out = out.detach()

    def _backward(self, loss, net):
        """``loss.backward()`` without ~160 per-parameter accumulate kernels: the flat gradient
        buffers are zero at this point (the fused SGD kernel clears them), so instead of letting
        autograd run ``p.grad += g`` once per parameter, the ``.grad`` views are detached for
        the backward pass (autograd then just keeps each fresh gradient) and the results are
        written into the arena with one multi-tensor copy per dtype."""
        if not self.batched_grad_copy:
            loss.backward()
            return
        if self._grad_slots is None:
            self._grad_slots = [(p, p.grad) for p in net.parameters()
                                if p.requires_grad and p.grad is not None]
        for p, _ in self._grad_slots:
            p.grad = None
        loss.backward()
        by_dtype = {}
        for p, view in self._grad_slots:
            g = p.grad
            p.grad = view
            if g is None:
                continue                                   # no gradient this step: slot stays zero
            if g.dtype != view.dtype or g.shape != view.shape:
                view.copy_(g)                              # foreign layout / dtype: plain copy
                continue
            dst, src = by_dtype.setdefault(view.dtype, ([], []))
            dst.append(view)
            src.append(g)
        for dst, src in by_dtype.values():
            torch._foreach_copy_(dst, src)

    def _step_sync(self):
        """forward/backward, then ONE kernel: SGD + publish + pull + mix + de-bias
        (SGD only when there is nobody to gossip with)."""
        self._fwd_bwd()
        if self.gossip and self.model.gossip_enable:
            self.engine.mix(sgd=True)
        else:
            self.engine.local(sgd=True)

    def _step_overlap(self, first: bool):
        """publish(k) -> [gather(k) on gossip_stream || fwd/bwd(k)] ; the SGD of
        step k is fused into publish(k+1)."""
        e, k = self.engine, self.k
        main = torch.cuda.current_stream(self.device)
        side = self.model.gossip_stream
        e.publish(sgd=not first, fold=not first)
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            e.gather()
            join = torch.cuda.Event()
            join.record(side)
        self._fwd_bwd()
        main.wait_event(join)

    def _one_step(self, first=False):
        if self.overlap and self.gossip and self.model.gossip_enable:
            self._step_overlap(first)
        else:
            self._step_sync()

    # ------------------------------------------------------------------ #
    def _ensure_static(self, batch, target):
        if self.static_in is not None:
            return
        fmt = torch.channels_last if (self.channels_last and batch.dim() == 4) \
            else torch.contiguous_format
        in_dtype = torch.bfloat16 if getattr(self.model, '_twin', None) else batch.dtype
        self.static_in = torch.empty(batch.shape, dtype=in_dtype, device=self.device
                                     ).contiguous(memory_format=fmt)
        self.static_tgt = torch.empty(target.shape, dtype=target.dtype, device=self.device)
        # staging keeps the HOST layout (plain pinned memcpy); the device-side
        # stage -> static copy performs the NCHW -> NHWC permute
        self._stage = torch.empty(batch.shape, dtype=batch.dtype, device=self.device)
        self._stage_tgt = torch.empty_like(self.static_tgt)
        self.static_metrics = torch.full((3,), float('nan'), dtype=torch.float32, device=self.device)
        self.static_loss = self.static_metrics[0]
        self._metrics_ring = torch.zeros(1024, 3, dtype=torch.float32).pin_memory()
        self._loss_ring = self._metrics_ring[:, 0]
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        self._stage_free.record(self.stream)

    def _set_lr(self):
        g = self.opt.param_groups[0]
        # overlap: the SGD of step k runs inside publish(k+1).  After finish() applied it early
        # (end of epoch / checkpoint) the next replay's publish must not apply it again: the
        # captured kernel reads `do_sgd` from device memory
        do_sgd = not self._skip_next_sgd
        self._skip_next_sgd = False
        self.engine.set_hyper(g['lr'], g['momentum'], g['weight_decay'], g['nesterov'],
                              do_sgd=do_sgd, grad_scale=self.opt.grad_scale)

    def prefetch(self, batch_cpu, target_cpu):
        """Start the H2D copy of the NEXT step's inputs on the copy stream."""
        self._ensure_static(batch_cpu, target_cpu)
        cs = self._copy_stream
        cs.wait_event(self._stage_free)
        on_device = [t for t in (batch_cpu, target_cpu) if t.is_cuda]
        if on_device:
            # inputs produced on the GPU (e.g. data.GpuAugment output, a target moved by the caller):
            # order the copy after the stream that produced them and keep their memory alive
            cs.wait_stream(torch.cuda.current_stream(self.device))
            for t in on_device:
                t.record_stream(cs)
        with torch.cuda.stream(cs):
            self._stage.copy_(batch_cpu, non_blocking=True)
            self._stage_tgt.copy_(target_cpu, non_blocking=True)
            self._stage_ready.record(cs)
        self._prefetched = True

    def _load_inputs(self, batch, target):
        main = self.stream
        if batch is not None and not self._prefetched:
            self.prefetch(batch, target)
        main.wait_event(self._stage_ready)
        with torch.cuda.stream(main):
            self.static_in.copy_(self._stage, non_blocking=True)
            self.static_tgt.copy_(self._stage_tgt, non_blocking=True)
            self._stage_free.record(main)
        self._prefetched = False

    def step(self, batch=None, target=None, next_batch=None, next_target=None,
             read_loss: bool = True):
        """One training iteration through the public path.

        ``batch``/``target``: (pinned) host or device tensors for THIS step;
        pass ``None`` if a previous call already staged them via
        ``next_batch``/``prefetch``.  ``next_batch``/``next_target``: inputs of
        the FOLLOWING step; their H2D copy is started on the prefetch stream as
        soon as this step's inputs have left the staging buffer, so it overlaps
        this step's compute.  Returns the slot of :attr:`loss_ring` that holds
        this step's loss -- and of :attr:`metrics_ring` that holds ``[loss, prec@1 %, prec@5 %]``
        -- once the stream has drained (no host sync here)."""
        if batch is not None:
            self._ensure_static(batch, target)
        self._load_inputs(batch, target)
        if next_batch is not None:
            self.prefetch(next_batch, next_target)
        self._run()
        slot = self._loss_slot
        if read_loss:
            with torch.cuda.stream(self.stream):
                self._metrics_ring[slot].copy_(self.static_metrics, non_blocking=True)   # 12 bytes D2H
            self._loss_slot = (slot + 1) % self._metrics_ring.shape[0]
        return slot

    def step_resident(self):
        """Replay the step on whatever is resident in the static input buffers
        (no H2D, no loss read-back): the device-only number of bench.py."""
        assert self.static_in is not None
        self._run()

    def _run(self):
        with torch.cuda.stream(self.stream):
            self._run_on_stream()

    def _graph_key(self):
        """Overlap-SGP with the copy-engine gather: the source address of the DMA copy (which
        in-neighbour, which outbox half) is baked into a captured graph, so there is one graph per
        (schedule row, parity) -- lcm(period, 2) of them, sharing one memory pool."""
        if self.overlap and self.gossip and self.model.gossip_enable \
                and getattr(self.engine, 'gather_dma', False):
            return self.engine.dma_key()
        return None

    def _run_on_stream(self):
        self._set_lr()
        key = self._graph_key()
        if key is not None and self._graphs:
            if key not in self._graphs:           # first visit of this (row, parity): capture it
                self.graph = None
                self._capture()
            self.graph = self._graphs[key]
        if self.graph is not None:
            self.graph.replay()
            self._after_replay()
        elif self.use_graph and self._eager_steps >= self.warmup_iters:
            self._capture()
            self.graph.replay()
            self._after_replay()
        else:
            c0 = self.engine.C.launch_count()
            self._one_step(first=(self._eager_steps == 0))
            self.own_launches_per_step = self.engine.C.launch_count() - c0
            self._eager_steps += 1

    def _after_replay(self):
        if self.gossip and self.model.gossip_enable:
            self.engine.steps += 1

    @property
    def loss_ring(self):
        return self._loss_ring

    @property
    def metrics_ring(self):
        """pinned ``[1024, 3]`` ring: row ``slot`` = ``[loss, prec@1 %, prec@5 %]`` of that step"""
        return self._metrics_ring

    def _capture(self):
        torch.cuda.synchronize(self.device)
        steps_before = self.engine.steps
        self.graph = torch.cuda.CUDAGraph()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()      # shared by the per-(row, parity) graphs
        key = self._graph_key()
        c0 = self.engine.C.launch_count()
        with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode='thread_local',
                              pool=self._pool):
            self._one_step(first=False)
        self.own_launches_per_step = self.engine.C.launch_count() - c0
        if key is not None:
            self._graphs[key] = self.graph
        # capture does not execute: undo the host-side step mirror advance
        self.engine.steps = steps_before
        torch.cuda.synchronize(self.device)

    def finish(self):
        """Drain: apply the last deferred SGD / residual (overlap), sync, and poll the kernels'
        health word.  Call at the end of an epoch, before validation / ``state_dict()``: the
        parameters then hold every update and all the push-sum mass received so far.  Training
        may simply continue afterwards (the next step skips the SGD that was applied here)."""
        if self.overlap and self.gossip and self._eager_steps + (self.graph is not None) > 0:
            with torch.cuda.stream(self.stream):
                self.stream.wait_stream(self.model.gossip_stream)
                self._set_lr()
                self.engine.local(sgd=True, fold=True)
                self.engine.residual.zero_()       # already folded: the next publish adds 0
            self._skip_next_sgd = True
        torch.cuda.synchronize(self.device)
        self.engine.check()

    def check(self):
        """Raise if a gossip kernel reported a heartbeat / ack timeout (reads one word)."""
        self.engine.check()
