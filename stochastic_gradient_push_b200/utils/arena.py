"""
Flat parameter arena.

The reference flattens the model with ``torch.cat`` into a fresh 102 MB buffer
every iteration and scatters the result back with one copy per parameter
(``gossip/distributed.py:441, 450-455``).  Here the parameters
*live* in one flat, 256-byte-aligned buffer for the whole run: every
``nn.Parameter.data`` is a view into it, so "flatten" and "unflatten" are the
identity, the gossip/SGD kernels stream the arena with 16-byte vector accesses
and -- when the buffer comes from the symmetric allocator -- NVSwitch peers can
read it directly.
"""

from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence

import torch

ALIGN_ELEMS = 64          # every tensor starts on a 256-byte boundary (fp32)
PAD_ELEMS = 4096          # arena length is a multiple of the kernels' chunk


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class FlatArena(object):
    """Layout of a list of same-dtype tensors inside one flat buffer."""

    def __init__(self, tensors: Sequence[torch.Tensor], device=None,
                 allocator: Optional[Callable] = None,
                 align: int = ALIGN_ELEMS, pad_to: int = PAD_ELEMS):
        tensors = list(tensors)
        assert len(tensors) > 0, 'arena needs at least one tensor'
        self.dtype = tensors[0].dtype
        assert all(t.dtype == self.dtype for t in tensors), \
            'one arena per dtype (see group_by_dtype)'
        self.device = torch.device(device) if device is not None else tensors[0].device
        self.shapes = [tuple(t.shape) for t in tensors]
        self.numels = [t.numel() for t in tensors]
        # 4-D tensors already stored channels-last keep that layout inside the
        # arena (cuDNN NHWC kernels then read the weights without a transpose)
        self.channels_last = [
            t.dim() == 4 and not t.is_contiguous()
            and t.is_contiguous(memory_format=torch.channels_last) for t in tensors]
        self.offsets: List[int] = []
        off = 0
        for n in self.numels:
            self.offsets.append(off)
            off = _round_up(off + n, align)
        self.used = off                       # extent incl. inter-tensor padding
        self.total = _round_up(max(off, 1), pad_to)
        self._allocator = allocator or self._default_alloc
        self.flat = self._allocator(self.total, self.dtype, self.device)
        assert self.flat.numel() >= self.total and self.flat.is_contiguous()
        self.flat = self.flat[:self.total]
        self.flat.zero_()
        self.views = self.views_of(self.flat)

    @staticmethod
    def _default_alloc(numel, dtype, device):
        return torch.empty(numel, dtype=dtype, device=device)

    # ------------------------------------------------------------------ #
    @property
    def payload(self) -> int:
        """number of real (non-padding) elements"""
        return sum(self.numels)

    def views_of(self, flat: torch.Tensor) -> List[torch.Tensor]:
        """Per-tensor views of any buffer that shares this layout."""
        assert flat.numel() >= self.used
        out = []
        for o, n, s, cl in zip(self.offsets, self.numels, self.shapes, self.channels_last):
            v = flat.narrow(0, o, n)
            if cl:      # (N,C,H,W) logical shape over an (N,H,W,C) physical layout
                v = v.view(s[0], s[2], s[3], s[1]).permute(0, 3, 1, 2)
            else:
                v = v.view(s)
            out.append(v)
        return out

    def new_buffer(self, dtype=None, allocator=None, zero=True) -> torch.Tensor:
        """Another flat buffer with the same layout (grads, momentum, shadow)."""
        alloc = allocator or self._default_alloc
        buf = alloc(self.total, dtype or self.dtype, self.device)[:self.total]
        if zero:
            buf.zero_()
        return buf

    # ------------------------------------------------------------------ #
    @torch.no_grad()
    def adopt(self, params: Iterable[torch.Tensor]):
        """Copy the current values in and re-point ``p.data`` at the views."""
        for p, v in zip(params, self.views):
            v.copy_(p.detach().to(v.device, v.dtype))
            p.data = v

    @torch.no_grad()
    def bind_grads(self, params: Iterable[torch.Tensor], grad_flat: torch.Tensor):
        """Make ``p.grad`` a view of ``grad_flat`` so autograd accumulates in
        place into a flat gradient the fused SGD kernel can stream."""
        for p, g in zip(params, self.views_of(grad_flat)):
            if p.requires_grad:
                p.grad = g

    def pack(self, tensors: Iterable[torch.Tensor], out: torch.Tensor = None):
        """Gather arbitrary (non-arena) tensors into a flat buffer of this
        layout -- the slow path used only for foreign tensors."""
        out = self.new_buffer() if out is None else out
        for t, v in zip(tensors, self.views_of(out)):
            v.copy_(t.view(v.shape) if t.shape != v.shape else t)
        return out
