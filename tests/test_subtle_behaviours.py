"""the reference behaviours that are easy to lose in a re-design, pinned on the c10d
(gloo, CPU) data plane with two ranks.

1. SGD (and therefore weight decay) acts on the push-sum NUMERATOR x, not on the de-biased z
   (gossip/distributed.py:564-565) -- only visible when the push-sum weight is not 1.
2. BatchNorm buffers are never gossiped (gossip/distributed.py:151): only module.parameters().
3. transfer_params() refuses (returns False) while the previous round has not been mixed (:397-400).
4. update_gossiper() with the value a gossiper already has is a no-op (:201-207).
5. regular graphs never transmit the push-sum weight; the receiver assumes own_w * ppi
   (gossip/gossiper.py:166-167) -- the c10d message is exactly the parameter count long.
"""
import torch
import torch.nn as nn

import stochastic_gradient_push_b200 as sgp
from stochastic_gradient_push_b200.ops import oracle

from dist_utils import run_distributed

LR, MU, WD = 0.1, 0.9, 0.1        # a large weight decay makes numerator-vs-z visible


def _net(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(4, 3), nn.Tanh(), nn.Linear(3, 2))


def _flat(m):
    return torch.cat([p.detach().reshape(-1) for p in m.parameters()])


def _batch(rank, step):
    g = torch.Generator().manual_seed(77 * rank + step)
    return torch.randn(5, 4, generator=g), torch.randn(5, 2, generator=g)


SELF_W = [0.3, 0.7]         # per-rank self weight: column-stochastic, NOT doubly stochastic -> w drifts from 1


def _numerator_worker(rank, world, steps):
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    graph = sgp.RingGraph(rank, world)
    mixing = sgp.SelfWeightedMixing(graph, 'cpu', self_weight=SELF_W)
    model = GossipDataParallel(_net(rank), graph=graph, mixing=mixing, push_sum=True,
                               rank=rank, world_size=world)
    assert not model.lazy_mixing          # (lazy mixing keeps w == 1 and skips the re-bias, as the reference)
    opt = torch.optim.SGD(model.parameters(), lr=LR, momentum=MU, weight_decay=WD)
    model.train()
    for step in range(steps):
        x, y = _batch(rank, step)
        ((model(x) - y) ** 2).mean().backward()
        opt.step()
        opt.zero_grad()
        model.transfer_params()
    model.sync_comms()
    model.unbias()
    return _flat(model.module).tolist(), float(model.state_dict()['ps_weight'])


def _simulate_numerator(world, steps, on_numerator=True):
    models = [_net(r) for r in range(world)]
    graphs = [sgp.RingGraph(r, world) for r in range(world)]
    mix = [sgp.SelfWeightedMixing(g, 'cpu', self_weight=SELF_W) for g in graphs]
    ws = [1.0] * world
    xs = [_flat(m) * w for m, w in zip(models, ws)]
    moms = [torch.zeros_like(x) for x in xs]
    for step in range(steps):
        grads = []
        for i, m in enumerate(models):
            off = 0
            for p in m.parameters():
                p.data.copy_((xs[i] / ws[i])[off:off + p.numel()].view_as(p))
                off += p.numel()
            m.zero_grad()
            x, y = _batch(i, step)
            ((m(x) - y) ** 2).mean().backward()
            grads.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]))
        for i in range(world):
            if on_numerator:
                xs[i], moms[i] = oracle.sgd_momentum(xs[i], grads[i], moms[i], LR, MU, WD, False)
            else:       # the tempting wrong design: step on z, then re-bias
                z, moms[i] = oracle.sgd_momentum(xs[i] / ws[i], grads[i], moms[i], LR, MU, WD, False)
                xs[i] = z * ws[i]
        xs, ws = oracle.mix_columns(xs, ws, graphs, mix)
        oracle.rotate_all(graphs)
    return [x / w for x, w in zip(xs, ws)], ws


def test_sgd_and_weight_decay_act_on_the_numerator():
    steps = 4
    got = run_distributed(_numerator_worker, 2, steps)
    want, ws = _simulate_numerator(2, steps, on_numerator=True)
    wrong, _ = _simulate_numerator(2, steps, on_numerator=False)
    for r in range(2):
        z = torch.tensor(got[r][0])
        assert abs(got[r][1] - ws[r]) < 1e-5
        assert abs(ws[r] - 1.0) > 1e-2                      # the weights really are not 1
        torch.testing.assert_close(z, want[r], rtol=1e-4, atol=1e-5)
        assert (z - wrong[r]).abs().max() > 1e-3            # and the test can tell the difference


class _BNNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(4, 6)
        self.bn = nn.BatchNorm1d(6)
        self.out = nn.Linear(6, 2)

    def forward(self, x):
        return self.out(torch.relu(self.bn(self.fc(x))))


def _bn_worker(rank, world):
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    torch.manual_seed(0)                         # identical initial weights
    net = _BNNet()
    model = GossipDataParallel(net, graph=sgp.RingGraph(rank, world), push_sum=True,
                               rank=rank, world_size=world)
    n_params = sum(p.numel() for p in net.parameters())
    n_gossiped = sum(a.flat.numel() for a in model._arenas.values())
    opt = torch.optim.SGD(model.parameters(), lr=0.0)         # lr 0: parameters only move by mixing
    model.train()
    for step in range(3):
        g = torch.Generator().manual_seed(step)
        x = torch.randn(16, 4, generator=g) + 5.0 * rank      # rank-dependent input statistics
        model(x).sum().backward()
        opt.step()
        opt.zero_grad()
        model.transfer_params()
    model.sync_comms()
    model.unbias()
    return n_params, n_gossiped, net.bn.running_mean.tolist(), _flat(net).tolist()


def test_batchnorm_buffers_are_not_gossiped():
    r0, r1 = run_distributed(_bn_worker, 2)
    assert r0[0] <= r0[1] < r0[0] + 4096          # the arena holds the parameters (+ padding), no buffers
    # parameters agree (same init, lr 0, mixing keeps them equal); running statistics do not
    torch.testing.assert_close(torch.tensor(r0[3]), torch.tensor(r1[3]), rtol=1e-5, atol=1e-6)
    assert (torch.tensor(r0[2]) - torch.tensor(r1[2])).abs().max() > 0.1


def _protocol_worker(rank, world):
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    net = _net(rank)
    graph = sgp.NPeerDynamicDirectedExponentialGraph(rank, world)
    model = GossipDataParallel(net, graph=graph, push_sum=True, rank=rank, world_size=world)
    model.train()
    x, y = _batch(rank, 0)
    ((model(x) - y) ** 2).mean().backward()
    first = model.transfer_params()
    second = model.transfer_params()          # previous round not mixed yet -> refused
    mixed_flag = model.params_mixed
    model._query_gossip_queue()               # mix it
    after = model.params_mixed
    # update_gossiper with the current value changes nothing (no schedule rebuild, no rotation)
    peers_before = graph.get_peers()
    model.update_gossiper('peers_per_itr', graph.peers_per_itr)
    peers_after = graph.get_peers()
    # regular graph: the message carries parameters only, never the push-sum weight
    n_params = sum(p.numel() for p in net.parameters())
    gossiper = list(model.dist_config['gossipers'].values())[0] if model.dist_config['gossipers'] else None
    msg_len = int(gossiper.in_msg_buffer.numel()) if gossiper is not None and hasattr(gossiper, 'in_msg_buffer') \
        else None
    model.sync_comms()
    return first, second, mixed_flag, after, peers_before == peers_after, n_params, msg_len, \
        float(model.state_dict()['ps_weight'])


def test_transfer_params_refuses_until_mixed_and_update_gossiper_is_idempotent():
    for first, second, mixed_flag, after, same_peers, n_params, msg_len, w in run_distributed(_protocol_worker, 2):
        assert first is True and second is False
        assert mixed_flag is False and after is True
        assert same_peers
        assert abs(w - 1.0) < 1e-6                    # own_w * ppi assumed by the receiver: stays exactly 1
        if msg_len is not None:
            assert msg_len == n_params                # no appended ps-weight slot on a regular graph
