"""GossipTrainer._backward (batched gradient copy into the flat arenas) on CPU: the trainer itself
needs the CUDA engine, but this method only touches autograd and the parameter <-> arena-view
bookkeeping, so it is exercised here on a bare instance."""
import torch
import torch.nn as nn

from stochastic_gradient_push_b200.parallel.trainer import GossipTrainer
from stochastic_gradient_push_b200.utils.arena import FlatArena


def _bare_trainer(batched=True):
    t = GossipTrainer.__new__(GossipTrainer)
    t.batched_grad_copy = batched
    t._grad_slots = None
    return t


def _net():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU(),
                         nn.Flatten(), nn.Linear(8 * 6 * 6, 5))


def _bound(net, dtype=torch.float32):
    """parameters stay where they are; their .grad become views of one flat buffer (as in the arena)"""
    params = [p for p in net.parameters() if p.requires_grad]
    flat = torch.zeros(sum(p.numel() for p in params), dtype=dtype)
    off, views = 0, []
    for p in params:
        v = flat[off:off + p.numel()].view_as(p)
        p.grad = v
        views.append(v)
        off += p.numel()
    return flat, views


def test_batched_copy_equals_plain_backward():
    x = torch.randn(4, 3, 6, 6)
    ref = _net()
    ref(x).square().mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])

    net = _net()
    flat, views = _bound(net)
    t = _bare_trainer()
    for step in range(3):                      # slots are cached after the first call
        flat.zero_()                           # what the fused SGD kernel does after consuming them
        t._backward(net(x).square().mean(), net)
        torch.testing.assert_close(flat, want)
        for p, v in zip(net.parameters(), views):
            assert p.grad is v                 # views are re-bound: the next kernel reads the arena
    assert len(t._grad_slots) == len(views)


def test_parameter_without_gradient_keeps_its_zero_slot():
    net = _net()
    extra = nn.Parameter(torch.ones(3))        # never used in the loss
    net.register_parameter('unused', extra)
    flat, views = _bound(net)
    slot = extra.grad                          # the arena view bound to the unused parameter
    t = _bare_trainer()
    t._backward(net(torch.randn(2, 3, 6, 6)).sum(), net)
    assert extra.grad is slot and float(slot.abs().sum()) == 0.0
    assert float(flat.abs().sum()) > 0


class _TwoDtypes(nn.Module):
    """fp32 layer -> bf16 layer, like the bf16 twin (bf16 conv / linear weights, fp32 BN parameters)."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(1)
        self.a = nn.Linear(6, 8)
        self.b = nn.Linear(8, 4).to(torch.bfloat16)

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)).to(torch.bfloat16)).float()


def test_mixed_dtype_slots_and_fallback_switch():
    """bf16 gradient slots for bf16 weights, fp32 for the rest: one multi-tensor copy per dtype;
    SGP_B200_BATCHED_GRAD_COPY=0 keeps autograd's per-parameter accumulation.  Same result."""
    x = torch.randn(5, 6)
    ref = _TwoDtypes()
    ref(x).square().mean().backward()
    for batched in (True, False):
        net = _TwoDtypes()
        slots = []
        for p in net.parameters():
            buf = torch.zeros(p.numel(), dtype=p.dtype)
            p.grad = buf.view_as(p)
            slots.append(buf)
        assert {b.dtype for b in slots} == {torch.float32, torch.bfloat16}
        t = _bare_trainer(batched)
        t._backward(net(x).square().mean(), net)
        for buf, p_ref, p in zip(slots, ref.parameters(), net.parameters()):
            assert p.grad.data_ptr() == buf.data_ptr()
            torch.testing.assert_close(buf, p_ref.grad.reshape(-1))
