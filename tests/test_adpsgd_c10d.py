"""BilatGossipDataParallel over the c10d transport (gloo, CPU)."""
import time

import torch
import torch.nn as nn

import stochastic_gradient_push_b200 as sgp

from dist_utils import run_distributed


def _flat(m):
    return torch.cat([p.detach().reshape(-1) for p in m.parameters()])


def _consensus_worker(rank, world, seconds):
    from stochastic_gradient_push_b200.parallel.ad_psgd import BilatGossipDataParallel
    torch.manual_seed(rank)
    net = nn.Linear(8, 4)
    with torch.no_grad():
        for p in net.parameters():
            p.fill_(float(rank))
    model = BilatGossipDataParallel(net, rank=rank, world_size=world,
                                    graph_class=sgp.DynamicBipartiteExponentialGraph,
                                    mixing_class=sgp.UniformMixing, lr=0.0, momentum=0.0,
                                    weight_decay=0.0, nesterov=False, verbose=False)
    assert model.transport == 'c10d'
    model.enable_gossip()
    t0 = time.time()
    while time.time() - t0 < seconds and model.rounds_completed < 12:
        time.sleep(0.01)
    model.disable_gossip()
    time.sleep(0.2)
    model.sync_comms()
    rounds = model.rounds_completed
    val = _flat(model.module)
    import torch.distributed as dist
    dist.barrier()
    model.shutdown()
    return val.mean().item(), (val.max() - val.min()).item(), rounds


def test_bilateral_gossip_reaches_consensus_without_gradients():
    world = 4
    out = run_distributed(_consensus_worker, world, 20.0, timeout=180)
    rounds = [o[2] for o in out]
    assert min(rounds) >= 2, rounds
    vals = [o[0] for o in out]
    assert all(o[1] < 1e-6 for o in out)              # every element moved identically
    assert max(vals) - min(vals) < 1.6                # contracted from the initial spread of 3.0
    assert 0.0 <= min(vals) and max(vals) <= 3.0


def _train_worker(rank, world, steps):
    from stochastic_gradient_push_b200.parallel.ad_psgd import BilatGossipDataParallel
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(6, 16), nn.Tanh(), nn.Linear(16, 1))
    model = BilatGossipDataParallel(net, rank=rank, world_size=world,
                                    graph_class=sgp.DynamicBipartiteExponentialGraph,
                                    mixing_class=sgp.UniformMixing, lr=0.05, momentum=0.9,
                                    weight_decay=0.0, nesterov=True, verbose=False)
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True)
    g = torch.Generator().manual_seed(rank)
    w_true = torch.arange(6.) / 6
    model.train()
    model.enable_gossip()
    losses = []
    for s in range(steps):
        x = torch.randn(32, 6, generator=g)
        y = (x @ w_true).unsqueeze(1)
        loss = ((model(x) - y) ** 2).mean()
        loss.backward()                      # hook: push grads to the gossip side, pull model
        opt.step()
        opt.zero_grad()                      # set_to_none=True: exercises the re-bind path
        losses.append(loss.item())
        time.sleep(0.002)
    model.update_lr(0.01)
    model.eval()
    applied = model.grads_applied
    model.disable_gossip()
    import torch.distributed as dist
    dist.barrier()
    model.shutdown()
    return losses[0], sum(losses[-5:]) / 5, applied


def test_adpsgd_trains_through_the_gossip_side_optimizer():
    out = run_distributed(_train_worker, 2, 60, timeout=240)
    for first, last, applied in out:
        assert last < 0.5 * first, (first, last)
        assert applied >= 55


def _split_worker(rank, world, seconds):
    import os
    os.environ['SGP_B200_C10D_SPLIT'] = '1'       # what the NCCL backend selects: one group per direction
    out = _consensus_worker(rank, world, seconds)
    return out


def test_bilateral_gossip_with_one_process_group_per_direction():
    """the NCCL arrangement of the c10d loop (messages from a higher to a lower rank on a second
    group, gossiper.C10dTransport) run over gloo: same protocol, same contraction"""
    world = 4
    out = run_distributed(_split_worker, world, 6.0, timeout=180)
    assert min(o[2] for o in out) >= 2
    vals = [o[0] for o in out]
    assert all(o[1] < 1e-6 for o in out)
    assert max(vals) - min(vals) < 1.6 and 0.0 <= min(vals) and max(vals) <= 3.0


def test_bilateral_update_keeps_gradients_applied_during_a_round():
    from stochastic_gradient_push_b200.parallel.ad_psgd import BilatGossipDataParallel as B
    g = torch.Generator().manual_seed(0)
    own, partner, delta = (torch.randn(9, generator=g) for _ in range(3))
    # nothing applied since the snapshot: the plain pairwise average of the reference
    torch.testing.assert_close(B._bilateral_update(own.clone(), own, partner), 0.5 * (own + partner))
    # a gradient step delta landed between publish and pull: it survives in full
    torch.testing.assert_close(B._bilateral_update(own + delta, own, partner), 0.5 * (own + partner) + delta)
