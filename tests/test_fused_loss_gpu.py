"""Fused softmax cross-entropy + prec@1/5 kernel (csrc/loss_kernels.cu) vs the plain PyTorch
composition the reference loop runs (CrossEntropyLoss + topk accuracy)."""
import pytest
import torch
import torch.nn.functional as F

from stochastic_gradient_push_b200.ops import fused_loss, native

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,C', [(256, 1000), (32, 1000), (7, 10), (1, 5), (64, 33), (16, 4099)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_loss_metrics_and_gradient(B, C, dtype):
    g = torch.Generator(device='cuda').manual_seed(B * 31 + C)
    logits = (torch.randn(B, C, device='cuda', generator=g) * 3).to(dtype).requires_grad_(True)
    target = torch.randint(0, C, (B,), device='cuda', generator=g)
    crit = fused_loss.FusedCrossEntropyWithAccuracy()
    loss = crit(logits, target)
    assert native.load().xent_can_fuse(logits.detach(), target)
    (loss * 1.7).backward()
    torch.cuda.synchronize()

    ref_logits = logits.detach().float().requires_grad_(True)
    want, p1, p5 = fused_loss.reference_loss_and_accuracy(ref_logits, target)
    (want * 1.7).backward()
    torch.testing.assert_close(loss.detach(), want.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(crit.metrics[0], want.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(crit.metrics[1], p1, rtol=0, atol=1e-3)
    torch.testing.assert_close(crit.metrics[2], p5, rtol=0, atol=1e-3)
    tol = 1e-6 if dtype == torch.float32 else 4e-3 / B
    torch.testing.assert_close(logits.grad.float(), ref_logits.grad, rtol=1e-2 if dtype != torch.float32 else 1e-4,
                               atol=tol)


def test_perfect_and_worst_predictions():
    C = 1000
    target = torch.arange(0, 64, device='cuda')
    logits = torch.zeros(64, C, device='cuda')
    logits[torch.arange(64), target] = 10.0
    _, m = fused_loss.fused_cross_entropy(logits, target)
    assert m[1].item() == pytest.approx(100.0) and m[2].item() == pytest.approx(100.0)
    logits = -logits
    _, m = fused_loss.fused_cross_entropy(logits, target)
    assert m[1].item() == 0.0 and m[2].item() == 0.0


def test_cuda_graph_capture():
    """the forward memset + kernels are stream-ordered and capturable (they run inside the
    captured training step)"""
    logits = torch.randn(32, 100, device='cuda', requires_grad=True)
    target = torch.randint(0, 100, (32,), device='cuda')
    out = torch.zeros(3, device='cuda')
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            loss, m = fused_loss.fused_cross_entropy(logits, target)
            loss.backward()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    logits.grad = None
    with torch.cuda.graph(graph):
        loss, m = fused_loss.fused_cross_entropy(logits, target)
        loss.backward()
        out.copy_(m)
    with torch.no_grad():
        logits.copy_(torch.randn(32, 100, device='cuda'))
    graph.replay()
    torch.cuda.synchronize()
    want = F.cross_entropy(logits.detach(), target)
    torch.testing.assert_close(out[0], want, rtol=1e-5, atol=1e-5)
