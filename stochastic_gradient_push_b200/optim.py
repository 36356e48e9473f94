"""
FusedGossipSGD: SGD-with-momentum whose update runs INSIDE the gossip kernel.

The reference calls ``torch.optim.SGD.step()`` (P x 3..5 elementwise launches
in torch 1.x, ``gossip_sgd.py:200-205, 389``) and then ``transfer_params()``.
Here ``optimizer.step()`` only records the hyper-parameters (device-resident,
so a captured CUDA graph picks up learning-rate changes) and marks the update
pending; the next gossip launch of the wrapped
:class:`~.parallel.distributed.GossipDataParallel` -- ``transfer_params()`` in
sync mode, the forward pre-hook in overlap mode -- applies

    d = g + wd*x ; m = mu*m + d ; x -= lr*(nesterov ? d + mu*m : m)

to the push-sum *numerator* in the same pass that publishes / mixes it (the
reference also steps the numerator: its backward hook re-biases before
``optimizer.step``, ``gossip/distributed.py:564-565``).  Gradients live in one
flat buffer (``p.grad`` are views, autograd accumulates in place) that the
kernel zeroes after consuming, so ``zero_grad()`` is free.

On the c10d transport (CPU / gloo) the same update runs as plain torch ops on
the flat arena -- that path is the oracle for the kernel.
"""

from __future__ import annotations

import torch

from .ops import oracle


class FusedGossipSGD(object):

    def __init__(self, model, lr=0.1, momentum=0.0, weight_decay=0.0, nesterov=False,
                 grad_scale=1.0):
        from .parallel.distributed import GossipDataParallel
        assert isinstance(model, GossipDataParallel)
        if nesterov and momentum <= 0:
            raise ValueError('Nesterov momentum requires a momentum')
        self.model = model
        self.defaults = dict(lr=lr, momentum=momentum, weight_decay=weight_decay,
                             nesterov=nesterov, dampening=0)
        self.param_groups = [dict(self.defaults, params=[p for p in model.module.parameters()])]
        self.grad_scale = grad_scale
        self._arenas = model._arenas
        self.grad_flat, self.momentum_flat = {}, {}
        for dtype, arena in self._arenas.items():
            hier = getattr(model, '_hier_grad', None)
            # hierarchical mode on the kernel plane: keep accumulating into the symmetric
            # (multicast-bound) flat gradient so the local-node average stays one NVLS kernel
            self.grad_flat[dtype] = hier if (hier is not None and dtype == torch.float32) \
                else arena.new_buffer()
            self.momentum_flat[dtype] = arena.new_buffer(dtype=torch.float32
                                                         if dtype != torch.float64 else dtype)
            arena.bind_grads(model._params_by_dtype[dtype], self.grad_flat[dtype])
        self.grad_lo = None
        if model._kernel is not None and model._twin:
            # bf16 compute twin: its conv/linear gradients accumulate into a bf16 flat
            # buffer, the normalisation gradients (and anything computed through the fp32
            # master module) into the fp32 one; the kernel adds the two
            arena = self._arenas[torch.float32]
            self.grad_lo = arena.new_buffer(dtype=torch.bfloat16)
            lo_views = arena.views_of(self.grad_lo)
            hi_views = arena.views_of(self.grad_flat[torch.float32])
            for p, lo, hi in zip(model._twin[0].parameters(), lo_views, hi_views):
                if p.requires_grad:
                    p.grad = lo if p.dtype == torch.bfloat16 else hi
            model._kernel.attach_sgd(self.grad_lo, self.momentum_flat[torch.float32],
                                     grad2=self.grad_flat[torch.float32])
        elif model._kernel is not None:
            model._kernel.attach_sgd(self.grad_flat[torch.float32],
                                     self.momentum_flat[torch.float32])
        model._fused_optimizer = self
        self._steps = 0

    # -- torch.optim-like surface -------------------------------------------- #
    def _hyper(self):
        g = self.param_groups[0]
        return g['lr'], g['momentum'], g['weight_decay'], g['nesterov']

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lr, mu, wd, nest = self._hyper()
        k = self.model._kernel
        if k is not None:
            k.engine.set_hyper(lr, mu, wd, nest, do_sgd=True, grad_scale=self.grad_scale)
            k.sgd_pending = True
        else:
            m = self.model
            m.ps_numerator()                       # step the numerator (reference :564)
            for dtype, arena in self._arenas.items():
                g = self.grad_flat[dtype]
                if self.grad_scale != 1.0:
                    g = g * self.grad_scale
                x, mom = oracle.sgd_momentum(arena.flat, g, self.momentum_flat[dtype],
                                             lr, mu, wd, nest)
                arena.flat.copy_(x)
                self.momentum_flat[dtype].copy_(mom)
        self._steps += 1
        return loss

    def zero_grad(self, set_to_none=False):
        """Free on the kernel path (the fused kernel clears the flat gradient
        right after reading it); explicit memset otherwise."""
        k = self.model._kernel
        if k is not None and k.sgd_pending:
            return
        for g in self.grad_flat.values():
            g.zero_()
        if self.grad_lo is not None:
            self.grad_lo.zero_()

    def state_dict(self):
        self.model._flush_pending()
        groups = [{k: v for k, v in g.items() if k != 'params'} for g in self.param_groups]
        return {'param_groups': groups, 'steps': self._steps,
                'momentum': {str(dt): m.detach().cpu().clone()
                             for dt, m in self.momentum_flat.items()}}

    def load_state_dict(self, sd):
        for g, saved in zip(self.param_groups, sd['param_groups']):
            g.update(saved)
        self._steps = sd.get('steps', 0)
        for dt, m in self.momentum_flat.items():
            m.copy_(sd['momentum'][str(dt)])
