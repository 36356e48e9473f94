"""Fused NHWC BatchNorm(+add)(+ReLU) kernels vs the plain PyTorch fp32 composition."""
import pytest
import torch

from stochastic_gradient_push_b200.ops.fused_bn import (FusedBatchNormAct2d, fused_bn_act,
                                                        reference_bn_act, _can_fuse)

pytestmark = pytest.mark.gpu


def assert_mostly_close(got, want, rtol, atol, max_bad=2e-5):
    """allclose up to a vanishing fraction of elements: a pre-activation within
    rounding distance of 0 may legitimately land on the other side of the ReLU."""
    bad = (got - want).abs() > atol + rtol * want.abs()
    frac = bad.float().mean().item()
    assert frac <= max_bad, 'mismatch fraction %.3g (max abs diff %.4g)' % (
        frac, (got - want).abs().max().item())


def _run(fn, x, res, w, b, relu, dtype):
    x = x.detach().clone().to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = None
    if res is not None:
        r = res.detach().clone().to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = w.detach().clone().requires_grad_(True)
    b = b.detach().clone().requires_grad_(True)
    C = w.numel()
    rm, rv = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
    nbt = torch.zeros((), dtype=torch.long, device='cuda')
    y = fn(x, w, b, rm, rv, nbt, residual=r, relu=relu, training=True, momentum=0.1, eps=1e-5)
    g = torch.Generator(device='cuda').manual_seed(5)
    dy = torch.randn(y.shape, device='cuda', generator=g).to(dtype).contiguous(
        memory_format=torch.channels_last)
    y.backward(dy)
    return dict(y=y.detach().float(), dx=x.grad.float(), dres=None if r is None else r.grad.float(),
                dw=w.grad, db=b.grad, rm=rm, rv=rv, nbt=int(nbt))


@pytest.mark.parametrize('shape', [(8, 64, 14, 14), (4, 256, 7, 9), (3, 2048, 4, 4), (5, 24, 6, 5),
                                   (16, 512, 2, 2), (32, 64, 56, 56), (8, 2048, 2, 2)])
@pytest.mark.parametrize('relu,add', [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_fused_bn_matches_reference(shape, relu, add, dtype):
    torch.manual_seed(0)
    N, C, H, W = shape
    x = torch.randn(shape, device='cuda') * 2 + 0.5
    res = torch.randn(shape, device='cuda') if add else None
    w = torch.rand(C, device='cuda') + 0.5
    b = torch.randn(C, device='cuda') * 0.1
    assert _can_fuse(x.to(dtype).contiguous(memory_format=torch.channels_last))
    got = _run(fused_bn_act, x, res, w, b, relu, dtype)
    # oracle: same (possibly bf16-rounded) inputs, fp32 math
    xq = x.to(dtype).float()
    rq = None if res is None else res.to(dtype).float()
    want = _run(reference_bn_act, xq, rq, w, b, relu, torch.float32)
    tol = dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    assert_mostly_close(got['y'], want['y'], **tol)
    assert_mostly_close(got['dx'], want['dx'], **tol)
    if add:
        assert_mostly_close(got['dres'], want['dres'], **tol)
    stat_tol = dict(rtol=1e-3, atol=1e-3) if dtype == torch.float32 else dict(rtol=3e-2, atol=3e-2)
    M = N * H * W
    torch.testing.assert_close(got['dw'], want['dw'], rtol=stat_tol['rtol'], atol=stat_tol['atol'] * M ** 0.5)
    torch.testing.assert_close(got['db'], want['db'], rtol=stat_tol['rtol'], atol=stat_tol['atol'] * M ** 0.5)
    torch.testing.assert_close(got['rm'], want['rm'], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(got['rv'], want['rv'], rtol=1e-3, atol=1e-3)
    assert got['nbt'] == 1


def test_variance_is_stable_when_mean_dominates():
    """|mean| >> std: a naive E[x^2]-E[x]^2 in fp32 loses the variance."""
    torch.manual_seed(0)
    x = (1000.0 + 0.05 * torch.randn(16, 64, 8, 8, device='cuda')).contiguous(
        memory_format=torch.channels_last)
    w, b = torch.ones(64, device='cuda'), torch.zeros(64, device='cuda')
    y = fused_bn_act(x, w, b, None, None, None, training=True)
    ref = torch.nn.functional.batch_norm(x.double(), None, None, w.double(), b.double(), True)
    torch.testing.assert_close(y.double(), ref, rtol=5e-3, atol=5e-3)


def test_module_eval_and_fallbacks():
    bn = FusedBatchNormAct2d(32).cuda()
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2.0)
    ref = torch.nn.BatchNorm2d(32).cuda()
    ref.load_state_dict(bn.state_dict())
    x = torch.randn(4, 32, 5, 5, device='cuda').contiguous(memory_format=torch.channels_last)
    bn.eval(), ref.eval()
    with torch.no_grad():
        torch.testing.assert_close(bn(x, relu=True), torch.relu(ref(x)), rtol=1e-5, atol=1e-5)
    # NCHW input -> reference path, identical semantics
    bn.train(), ref.train()
    xn = torch.randn(4, 32, 5, 5, device='cuda')
    torch.testing.assert_close(bn(xn), ref(xn), rtol=1e-5, atol=1e-5)
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked) == 1


def test_resnet50_eval_matches_torchvision():
    """eval mode (running statistics): the fused net IS torchvision's function."""
    import torchvision
    from stochastic_gradient_push_b200.models import resnet50
    torch.manual_seed(0)
    tv = torchvision.models.resnet50().cuda().to(memory_format=torch.channels_last).eval()
    ours = resnet50().cuda().to(memory_format=torch.channels_last).eval()
    ours.load_state_dict(tv.state_dict())
    x = torch.randn(4, 3, 96, 96, device='cuda').contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        yo, yt = ours(x), tv(x)
    scale = yt.abs().max().item()
    assert (yo - yt).abs().max().item() < 2e-3 * scale


@pytest.mark.parametrize('amp', [False, True])
def test_resnet50_every_bn_layer_matches_composition(amp, monkeypatch):
    """Per-layer A/B inside a real ResNet-50 pass: every FusedBatchNormAct2d call
    is compared with the PyTorch composition on the SAME input and the SAME upstream
    gradient.  (Comparing two whole networks end to end is meaningless here: batch
    statistics over the tiny late-stage feature maps amplify rounding noise
    chaotically; measured per-layer agreement is ~1e-7 in fp32.)"""
    from stochastic_gradient_push_b200.models import resnet50
    from stochastic_gradient_push_b200.ops import fused_bn
    # this test is about the stand-alone BN op: keep the 1x1 convolutions on the path that calls it
    # (the conv + BN fusion has its own A/B tests in test_conv1x1_gpu.py)
    monkeypatch.setattr(fused_bn, 'USE_TCGEN05_CONV1X1', False)
    torch.manual_seed(0)
    net = resnet50().cuda().to(memory_format=torch.channels_last)
    x = torch.randn(16, 3, 128, 128, device='cuda').contiguous(memory_format=torch.channels_last)
    rows = []
    orig = FusedBatchNormAct2d.forward

    def rel(a, b):
        return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()

    def patched(self, inp, residual=None, relu=False):
        with torch.enable_grad():
            xs = [inp.detach().clone().requires_grad_(True) for _ in range(2)]
            rs = [None if residual is None else residual.detach().clone().requires_grad_(True)
                  for _ in range(2)]
            ws = [self.weight.detach().clone().requires_grad_(True) for _ in range(2)]
            bs = [self.bias.detach().clone().requires_grad_(True) for _ in range(2)]
            assert _can_fuse(inp)
            y0 = fused_bn_act(xs[0], ws[0], bs[0], None, None, None, residual=rs[0], relu=relu)
            y1 = reference_bn_act(xs[1].float(), ws[1], bs[1], None, None, None,
                                  residual=None if rs[1] is None else rs[1].float(), relu=relu)
            g = torch.randn_like(y1)
            y0.backward(g.to(y0.dtype))
            y1.backward(g)
            rows.append((rel(y0, y1), rel(xs[0].grad, xs[1].grad), rel(ws[0].grad, ws[1].grad),
                         rel(bs[0].grad, bs[1].grad)))
        return orig(self, inp, residual, relu)

    FusedBatchNormAct2d.forward = patched
    try:
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
            net(x)
    finally:
        FusedBatchNormAct2d.forward = orig
    assert len(rows) == 53
    tol_y, tol_g = (1e-2, 2e-2) if amp else (1e-5, 5e-3)     # dx tolerates a rare ReLU-mask flip
    for ry, rdx, rdw, rdb in rows:
        assert ry < tol_y and rdx < tol_g and rdw < tol_g and rdb < tol_g, (ry, rdx, rdw, rdb)


@pytest.mark.parametrize('shape,k,s,p', [((4, 64, 28, 28), 3, 2, 1), ((2, 16, 9, 11), 3, 2, 1),
                                         ((3, 8, 8, 8), 2, 2, 0), ((2, 24, 7, 7), 3, 1, 1)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_maxpool_nhwc_matches_torch(shape, k, s, p, dtype):
    from stochastic_gradient_push_b200.ops.fused_bn import MaxPool2dNHWC
    torch.manual_seed(0)
    x = torch.randn(shape, device='cuda').to(dtype).contiguous(memory_format=torch.channels_last)
    x1 = x.clone().requires_grad_(True)
    x2 = x.clone().requires_grad_(True)
    y1 = MaxPool2dNHWC(k, s, p)(x1)
    y2 = torch.nn.functional.max_pool2d(x2, k, s, p)
    assert y1.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(y1, y2, rtol=0, atol=0)
    dy = torch.randn_like(y2)
    y1.backward(dy)
    y2.backward(dy)
    torch.testing.assert_close(x1.grad.float(), x2.grad.float(), rtol=1e-2, atol=1e-2)


def test_maxpool_ties_route_to_first_max():
    """post-ReLU inputs are full of exact ties (zeros): one winner per window."""
    from stochastic_gradient_push_b200.ops.fused_bn import MaxPool2dNHWC
    x = torch.relu(torch.randn(2, 8, 12, 12, device='cuda')).contiguous(memory_format=torch.channels_last)
    x1 = x.clone().requires_grad_(True)
    x2 = x.clone().requires_grad_(True)
    MaxPool2dNHWC(3, 2, 1)(x1).sum().backward()
    torch.nn.functional.max_pool2d(x2, 3, 2, 1).sum().backward()
    torch.testing.assert_close(x1.grad, x2.grad, rtol=0, atol=0)


@pytest.mark.parametrize('shape', [(4, 3, 64, 64), (3, 3, 224, 224), (2, 3, 70, 90), (1, 3, 33, 47)])
def test_stem_conv_matches_conv2d(shape):
    """tensor-core stem convolution (fwd + wgrad) vs F.conv2d in fp32 on the same
    bf16-rounded operands."""
    from stochastic_gradient_push_b200.ops.fused_bn import stem_conv
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).cuda().to(memory_format=torch.channels_last)
    x = torch.randn(shape, device='cuda').bfloat16().contiguous(memory_format=torch.channels_last)
    y = stem_conv(conv, x)
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    wq = conv.weight.detach().bfloat16().float().requires_grad_(True)
    y_ref = torch.nn.functional.conv2d(x.float(), wq, None, 2, 3)
    assert y.shape == y_ref.shape
    torch.testing.assert_close(y.float(), y_ref, rtol=2e-2, atol=2e-2)
    dy = torch.randn_like(y_ref).bfloat16()
    y.backward(dy)
    y_ref.backward(dy.float())
    scale = wq.grad.abs().max().item()
    assert (conv.weight.grad - wq.grad).abs().max().item() < 1e-2 * scale


@pytest.mark.parametrize('shape', [(4, 3, 64, 64), (3, 3, 224, 224), (2, 3, 70, 90), (1, 3, 33, 47)])
def test_stem_conv_tf32_matches_conv2d(shape):
    """fp32 operands, TF32 tensor-core math (mma.sync.m16n8k8.tf32): forward and wgrad against an
    fp64 convolution of the tf32-truncated operands (tight) and of the fp32 operands (TF32-level)."""
    from stochastic_gradient_push_b200.ops.fused_bn import stem_conv

    def tf32(t):
        return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)

    torch.manual_seed(0)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).cuda().to(memory_format=torch.channels_last)
    x = torch.randn(shape, device='cuda').contiguous(memory_format=torch.channels_last)
    y = stem_conv(conv, x)
    assert y.dtype == torch.float32 and y.is_contiguous(memory_format=torch.channels_last)
    assert y.grad_fn is not None and type(y.grad_fn).__name__.startswith('_StemConv')
    w = conv.weight.detach()
    exact = torch.nn.functional.conv2d(tf32(x).double(), tf32(w).double(), None, 2, 3).float()
    torch.testing.assert_close(y, exact, rtol=1e-4, atol=1e-4)
    full = torch.nn.functional.conv2d(x.double(), w.double(), None, 2, 3).float()
    torch.testing.assert_close(y, full, rtol=5e-3, atol=5e-3)
    dy = torch.randn_like(y)
    y.backward(dy)
    wd = w.double().requires_grad_(True)
    torch.nn.functional.conv2d(tf32(x).double(), wd, None, 2, 3).backward(tf32(dy).double())
    scale = wd.grad.abs().max().item()
    assert (conv.weight.grad.double() - wd.grad).abs().max().item() < 2e-4 * scale + 1e-4


def test_stem_conv_inside_resnet_training_step():
    """the whole bf16 twin path (stem conv + fused BN + max-pool) produces finite,
    decreasing losses -- see test_flagship_gpu for the trainer-level checks."""
    from stochastic_gradient_push_b200.models import resnet50
    from stochastic_gradient_push_b200.ops import fused_bn
    net = resnet50().cuda().to(memory_format=torch.channels_last)
    x = torch.randn(8, 3, 96, 96, device='cuda').contiguous(memory_format=torch.channels_last)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        y1 = net(x)
    fused_bn.FORCE_REFERENCE = True
    try:
        net2 = resnet50().cuda().to(memory_format=torch.channels_last)
        net2.load_state_dict(net.state_dict())
        with torch.autocast('cuda', dtype=torch.bfloat16):
            s1 = fused_bn.stem_conv(net2.conv1, x)          # reference path: cuDNN
    finally:
        fused_bn.FORCE_REFERENCE = False
    with torch.autocast('cuda', dtype=torch.bfloat16):
        s2 = fused_bn.stem_conv(net.conv1, x)
    torch.testing.assert_close(s2.float(), s1.float(), rtol=3e-2, atol=3e-2)
    assert torch.isfinite(y1).all()
