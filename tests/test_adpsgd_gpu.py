"""AD-PSGD on the kernel data plane, loop-back: two virtual ranks (LocalWorld) on ONE GPU, each
with its own native gossip daemon (csrc/bindings.cpp::BilatDaemon) driving the device-side round
state machine (sgp_bilat_decide_kernel + SGP_F_FROM_STATE worker launches).  No process group, so
the driver's single-GPU run exercises the whole bilateral protocol."""
import time

import pytest
import torch

import stochastic_gradient_push_b200 as sgp

pytestmark = pytest.mark.gpu


def _pair(lr=0.0, max_rounds=None, fill=None, seed=0):
    from stochastic_gradient_push_b200.parallel.ad_psgd import BilatGossipDataParallel
    from stochastic_gradient_push_b200.parallel.symmetric import LocalWorld
    dev = torch.device('cuda', 0)
    lw = LocalWorld(2)
    models = []
    for r in range(2):
        torch.manual_seed(seed)
        net = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.Tanh(), torch.nn.Linear(32, 1)).to(dev)
        if fill is not None:
            with torch.no_grad():
                for p in net.parameters():
                    p.fill_(fill[r])
        m = BilatGossipDataParallel(net, rank=r, world_size=2, graph_class=sgp.DynamicBipartiteExponentialGraph,
                                    mixing_class=sgp.UniformMixing, lr=lr, momentum=0.9, weight_decay=0.0,
                                    nesterov=True, verbose=False, heartbeat_timeout=20, transport='nvlink',
                                    symmetric_world=lw.view(r), max_rounds_per_update=max_rounds, gossip_grid=4)
        m.daemon.start()         # (world_size > 1 without a process group: started by hand)
        models.append(m)
    return models


def _flat(m):
    return torch.cat([p.detach().reshape(-1) for p in m.module.parameters()])


def test_daemons_reach_consensus_and_conserve_mass():
    ms = _pair(fill=[0.0, 4.0])
    try:
        for m in ms:
            m.enable_gossip()
        t0 = time.time()
        while time.time() - t0 < 20 and min(m.rounds_completed for m in ms) < 6:
            time.sleep(0.01)
        for m in ms:
            m.disable_gossip()
        time.sleep(0.2)
        torch.cuda.synchronize()
        for m in ms:
            m.sync_comms()
            m._check()
        torch.cuda.synchronize()
        a, b = _flat(ms[0]), _flat(ms[1])
        assert min(m.rounds_completed for m in ms) >= 2
        # every parameter of a rank holds the same value; the pair average is conserved (2.0) and
        # the ranks have contracted towards it
        assert (a.max() - a.min()).item() < 1e-6 and (b.max() - b.min()).item() < 1e-6
        assert abs((a[0] + b[0]).item() / 2 - 2.0) < 1e-5
        assert abs(a[0].item() - b[0].item()) < 1e-5          # one exact average is enough for two ranks
        assert ms[0].daemon.pairs_enqueued() > 0 and not ms[0].daemon.error()
    finally:
        for m in ms:
            m.shutdown()


def test_budget_bounds_rounds_per_gradient():
    """without gradients a rank may only START max_rounds_per_update rounds"""
    ms = _pair(fill=[1.0, 3.0], max_rounds=2)
    try:
        for m in ms:
            m.enable_gossip()
        time.sleep(1.0)
        torch.cuda.synchronize()
        r = [m.rounds_completed for m in ms]
        assert max(r) <= 2 and min(r) >= 1, r
        idle = ms[0].daemon.idle_polls()
        time.sleep(0.2)
        assert ms[0].daemon.idle_polls() > idle              # the daemon is alive and backing off
    finally:
        for m in ms:
            m.shutdown()


def test_training_through_hooks_and_fast_handoff():
    """reference-style loop (backward hook: transfer grads + pull, own optimizer.step) on rank 0's
    data and the fused hand-off on rank 1's: both learn the regression target while gossiping"""
    ms = _pair(lr=0.05, max_rounds=4, seed=1)
    try:
        dev = torch.device('cuda', 0)
        w_true = (torch.arange(6.) / 6).to(dev)
        opts = [torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9, nesterov=True) for m in ms]
        for m in ms:
            m.train()
            m.enable_gossip()
        first, last = [None, None], [None, None]
        g = torch.Generator(device='cuda').manual_seed(3)
        for s in range(60):
            for r, m in enumerate(ms):
                x = torch.randn(64, 6, device=dev, generator=g)
                y = (x @ w_true).unsqueeze(1)
                loss = ((m(x) - y) ** 2).mean()
                loss.backward()          # hook: _transfer_grads + _pull_model
                opts[r].step()
                opts[r].zero_grad(set_to_none=False)
                if s == 0:
                    first[r] = loss.item()
                last[r] = loss.item()
        for m in ms:
            m.eval()
            m.disable_gossip()
            m._check()
        assert all(m.grads_applied >= 59 for m in ms)
        assert min(m.rounds_completed for m in ms) >= 5
        for r in range(2):
            assert last[r] < 0.5 * first[r], (first, last)
        # the two gossip copies stay close (they average every few steps)
        torch.cuda.synchronize()
        d = (ms[0].gossip_flat - ms[1].gossip_flat).abs().max().item()
        assert d < 0.5, d
    finally:
        for m in ms:
            m.shutdown()
