"""Single-GPU flagship path: smoke(), AR comparator at world 1, trainer on ResNet-50."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graft_smoke():
    import __graft_entry__ as ge
    ge.smoke()


def test_trainer_resnet50_loss_decreases_world1():
    import stochastic_gradient_push_b200 as sgp
    from stochastic_gradient_push_b200 import models
    from stochastic_gradient_push_b200.optim import FusedGossipSGD
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    from stochastic_gradient_push_b200.parallel.trainer import GossipTrainer
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    net = models.init_imagenet_in_1hr(models.resnet50()).to(dev).to(memory_format=torch.channels_last)
    model = GossipDataParallel(net, graph=sgp.NPeerDynamicDirectedExponentialGraph(0, 1),
                               rank=0, world_size=1, heartbeat_timeout=20)
    opt = FusedGossipSGD(model, lr=0.02, momentum=0.9, weight_decay=1e-4, nesterov=True)
    tr = GossipTrainer(model, opt, use_cuda_graph=True, warmup_iters=2)
    x = torch.randn(8, 3, 96, 96).pin_memory()
    y = torch.randint(0, 1000, (8,)).pin_memory()
    slots = [tr.step(x, y) for _ in range(12)]
    tr.finish()
    losses = [float(tr.loss_ring[s]) for s in slots]
    assert tr.graph is not None                    # captured
    assert losses[-1] < losses[0]                  # memorises the fixed batch
    assert all(l == l for l in losses)


def test_allreduce_world1_matches_torch_sgd():
    from stochastic_gradient_push_b200.parallel.allreduce import AllReduceDataParallel
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 8)).to(dev)
    import copy
    ref = copy.deepcopy(net)
    ar = AllReduceDataParallel(net, rank=0, world_size=1)
    opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=True)
    ar.set_hyper(0.1, 0.9, 1e-4, True)
    for _ in range(4):
        x = torch.randn(16, 64, device=dev)
        ar(x).square().mean().backward()
        ar.allreduce_step()
        opt.zero_grad()
        ref(x).square().mean().backward()
        opt.step()
    torch.cuda.synchronize()
    ar.check()
    for p, q in zip(net.parameters(), ref.parameters()):
        torch.testing.assert_close(p, q, rtol=1e-4, atol=1e-5)


def test_single_process_multi_gpu_is_rejected_with_guidance():
    import stochastic_gradient_push_b200 as sgp
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    net = torch.nn.Linear(4, 4).cuda(0)
    with pytest.raises(NotImplementedError, match='one rank per GPU'):
        GossipDataParallel(net, device_ids=[0, 1], rank=0, world_size=1,
                           graph=sgp.NPeerDynamicDirectedExponentialGraph(0, 1))
