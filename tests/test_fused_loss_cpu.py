"""CPU fallback of the fused loss module == CrossEntropyLoss + the reference's accuracy()."""
import torch
import torch.nn.functional as F

from stochastic_gradient_push_b200.ops import fused_loss


def test_cpu_fallback_matches_reference_composition():
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(40, 17, generator=g, requires_grad=True)
    target = torch.randint(0, 17, (40,), generator=g)
    crit = fused_loss.FusedCrossEntropyWithAccuracy()
    loss = crit(logits, target)
    loss.backward()
    ref = logits.detach().clone().requires_grad_(True)
    want = F.cross_entropy(ref, target)
    want.backward()
    torch.testing.assert_close(loss, want)
    torch.testing.assert_close(logits.grad, ref.grad)
    pred = logits.detach().topk(5, 1)[1]
    p1 = (pred[:, 0] == target).float().mean() * 100
    p5 = (pred == target[:, None]).any(1).float().mean() * 100
    torch.testing.assert_close(crit.metrics[1], p1)
    torch.testing.assert_close(crit.metrics[2], p5)
    torch.testing.assert_close(crit.metrics[0], want.detach())
