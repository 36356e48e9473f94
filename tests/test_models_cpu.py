"""Model zoo on CPU: the fused ops must degrade to the plain PyTorch composition, parameter
order / count must match torchvision's ResNets (checkpoints and gossip message sizes of the
reference, ``gossip_sgd.py:693-707``), and the reference initialisation must be applied."""
import pytest
import torch
import torch.nn as nn

from stochastic_gradient_push_b200.models import resnet as R
from stochastic_gradient_push_b200.ops.fused_bn import (FusedBatchNormAct2d, conv_bn_act,
                                                        conv_bn_act_split)


def test_resnet50_parameter_inventory():
    net = R.resnet50()
    params = list(net.parameters())
    assert len(params) == 161
    assert sum(p.numel() for p in params) == 25557032
    names = [n for n, _ in net.named_parameters()]
    assert names[0] == 'conv1.weight' and names[-1] == 'fc.bias'
    assert 'layer1.0.downsample.0.weight' in names and 'layer4.2.bn3.bias' in names
    # state_dict keys of a BatchNorm are the stock ones (drop-in for torchvision checkpoints)
    sd = net.state_dict()
    for k in ('bn1.running_mean', 'bn1.running_var', 'bn1.num_batches_tracked'):
        assert k in sd


@pytest.mark.parametrize('name,n_params', [('resnet18', 11689512), ('resnet34', 21797672),
                                           ('resnet101', 44549160), ('resnet152', 60192808)])
def test_other_resnets_match_torchvision_sizes(name, n_params):
    net = R.MODEL_ZOO[name]()
    assert sum(p.numel() for p in net.parameters()) == n_params


def test_imagenet_in_1hr_init():
    net = R.init_imagenet_in_1hr(R.resnet50())
    for m in net.modules():
        if isinstance(m, R.Bottleneck):
            assert float(m.bn3.weight.detach().abs().sum()) == 0.0
            assert float(m.bn1.weight.detach().min()) == 1.0
    assert abs(float(net.fc.weight.detach().std()) - 0.01) < 2e-3


def test_bottleneck_cpu_matches_plain_composition():
    """On CPU conv_bn_act / conv_bn_act_split are exactly bn(conv(x)) (+ residual) (+ ReLU)."""
    torch.manual_seed(0)
    blk = R.Bottleneck(32, 8, 1, nn.Sequential(R._conv1x1(32, 32), R._BN(32)))
    x = torch.randn(4, 32, 6, 6, requires_grad=True)
    # plain reference with stock modules sharing the same parameters / fresh running stats
    import copy
    ref = copy.deepcopy(blk)
    for m in ref.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.reset_running_stats()
    for m in blk.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.reset_running_stats()
    y = blk(x)
    x2 = x.detach().clone().requires_grad_(True)

    def plain(b, t):
        idt = nn.functional.batch_norm(b.downsample[0](t), None, None, b.downsample[1].weight,
                                       b.downsample[1].bias, True)
        o = torch.relu(nn.functional.batch_norm(b.conv1(t), None, None, b.bn1.weight, b.bn1.bias, True))
        o = torch.relu(nn.functional.batch_norm(b.conv2(o), None, None, b.bn2.weight, b.bn2.bias, True))
        o = nn.functional.batch_norm(b.conv3(o), None, None, b.bn3.weight, b.bn3.bias, True)
        return torch.relu(o + idt)

    y2 = plain(ref, x2)
    torch.testing.assert_close(y, y2, rtol=1e-5, atol=1e-5)
    g = torch.randn_like(y)
    y.backward(g)
    y2.backward(g)
    torch.testing.assert_close(x.grad, x2.grad, rtol=1e-4, atol=1e-5)
    for (n, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(p.grad, q.grad, rtol=1e-4, atol=1e-5, msg=n)


def test_split_op_falls_back_to_identity_alias_on_cpu():
    conv = nn.Conv2d(8, 16, 1, bias=False)
    bn = FusedBatchNormAct2d(16)
    x = torch.randn(2, 8, 5, 5, requires_grad=True)
    out, skip = conv_bn_act_split(conv, bn, x, relu=True)
    assert skip is x
    torch.testing.assert_close(out, conv_bn_act(conv, bn, x.detach(), relu=True))
    # eval mode uses the running statistics
    bn.eval()
    out_eval = conv_bn_act(conv, bn, x, relu=False)
    want = nn.functional.batch_norm(conv(x), bn.running_mean, bn.running_var, bn.weight, bn.bias, False)
    torch.testing.assert_close(out_eval, want)


def test_tiny_and_resnet18_train_step_cpu():
    for net, shape in ((R.TinyConvNet(), (4, 3, 16, 16)), (R.resnet18(num_classes=10), (2, 3, 32, 32))):
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        x = torch.randn(*shape)
        loss0 = None
        for _ in range(3):
            opt.zero_grad()
            loss = nn.functional.cross_entropy(net(x), torch.zeros(shape[0], dtype=torch.long))
            loss.backward()
            opt.step()
            loss0 = loss0 if loss0 is not None else float(loss.detach())
        assert float(loss.detach()) < loss0
