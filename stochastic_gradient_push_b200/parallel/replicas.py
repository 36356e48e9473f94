"""
LocalReplicas: the reference's single-process multi-GPU mode (one process drives several GPUs;
``gossip/ad_psgd.py:57-69, 148-191, 378-404`` and ``gossip/distributed.py:87-99, 253-276, 523-549``)
as a small stand-alone helper.

The reference replicates the module with ``torch.nn.parallel.replicate``, broadcasts every
parameter with ``broadcast_coalesced`` before each forward and sums the replicas' gradients with
``reduce_add_coalesced`` in the backward hook (N10).  Here every replica owns ONE flat parameter
arena and ONE flat gradient buffer, so

* parameter sync  = one peer-to-peer DMA copy of the master arena per replica (+ the buffers),
* gradient reduce = ONE kernel on the master GPU that sums the replicas' flat gradients with
  16-byte P2P loads (``_C.peer_reduce_`` -> ``sgp_peer_reduce_kernel``),

and the usual scatter -> parallel_apply (one thread per GPU) -> gather run the forward pass.
Gossip always happens on the master (``device_ids[0]``) copy, as in the reference.

``GossipDataParallel`` has the same logic built in (``_build_local_replicas``); this helper serves
``BilatGossipDataParallel``.
"""

from __future__ import annotations

import copy
from typing import List, Sequence

import torch

from ..utils.arena import FlatArena


class LocalReplicas(object):

    def __init__(self, module: torch.nn.Module, device_ids: Sequence[int], master_arena: FlatArena,
                 master_params: List[torch.nn.Parameter]):
        assert len(device_ids) > 1
        self.module = module
        self.device_ids = list(device_ids)
        self.master_arena = master_arena
        self.master_params = master_params
        self.copies = [module]
        self.arenas = []            # (arena, grad_flat) of replicas 1..
        for dev_idx in self.device_ids[1:]:
            dev = torch.device('cuda', dev_idx)
            rep = copy.deepcopy(module).to(dev)
            params = list(rep.parameters())
            arena = FlatArena(params, device=dev)
            arena.adopt(params)
            grad = arena.new_buffer()
            arena.bind_grads(params, grad)
            rep.train(module.training)
            self.copies.append(rep)
            self.arenas.append((arena, grad))

    # ------------------------------------------------------------------ #
    def train(self, mode=True):
        for m in self.copies[1:]:
            m.train(mode)

    def sync_params(self):
        """master parameters / buffers -> every replica (one DMA copy of the flat arena each)"""
        src = self.master_arena.flat
        cur = torch.cuda.current_stream(src.device)
        for (arena, _), rep in zip(self.arenas, self.copies[1:]):
            with torch.cuda.device(arena.flat.device):
                s = torch.cuda.current_stream(arena.flat.device)
                s.wait_stream(cur)
                arena.flat.copy_(src, non_blocking=True)
                for b_m, b_r in zip(self.module.buffers(), rep.buffers()):
                    b_r.copy_(b_m, non_blocking=True)

    def forward(self, inputs, kwargs, output_device):
        """``inputs`` / ``kwargs``: already scattered over ``device_ids``"""
        from torch.nn.parallel.parallel_apply import parallel_apply
        from torch.nn.parallel.scatter_gather import gather
        self.sync_params()
        n = min(len(inputs), len(self.copies))
        outs = parallel_apply(self.copies[:n], inputs[:n], kwargs[:n], self.device_ids[:n])
        return gather(outs, output_device, dim=0)

    def reduce_grads(self, master_grad_flat: torch.Tensor):
        """master_grad_flat += sum of the replicas' flat gradients (in place, one kernel), then the
        replicas' gradients are cleared.  Falls back to per-tensor adds if an optimizer replaced the
        master's ``.grad`` views."""
        from ..ops import native
        dev0 = master_grad_flat.device
        cur = torch.cuda.current_stream(dev0)
        for _, g in self.arenas:
            cur.wait_stream(torch.cuda.current_stream(g.device))
        grads = [g for _, g in self.arenas]
        in_place = False
        for p, v in zip(self.master_params, self.master_arena.views_of(master_grad_flat)):
            if p.requires_grad:
                in_place = p.grad is not None and p.grad.data_ptr() == v.data_ptr()
                break
        fused = native.available() and master_grad_flat.is_cuda
        if in_place:
            if fused:
                native.load().peer_reduce_(master_grad_flat, [master_grad_flat] + grads, 1.0)
            else:
                for g in grads:
                    master_grad_flat.add_(g.to(dev0))
        else:
            scratch = self.master_arena.new_buffer()
            if fused:
                native.load().peer_reduce_(scratch, grads, 1.0)
            else:
                for g in grads:
                    scratch.add_(g.to(dev0))
            for p, v in zip(self.master_params, self.master_arena.views_of(scratch)):
                if p.requires_grad:
                    p.grad = v.clone() if p.grad is None else p.grad.add_(v)
        for _, g in self.arenas:
            with torch.cuda.device(g.device):
                torch.cuda.current_stream(g.device).wait_stream(cur)
                g.zero_()
