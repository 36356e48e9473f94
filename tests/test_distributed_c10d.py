"""GossipDataParallel over the c10d (gloo, CPU) transport vs a single-process
simulation of the whole world (SURVEY 4.2 'Dist-CPU' tier)."""
import copy

import pytest
import torch
import torch.nn as nn

import stochastic_gradient_push_b200 as sgp
from stochastic_gradient_push_b200.ops import oracle

from dist_utils import run_distributed

LR, MU, WD = 0.05, 0.9, 1e-3


def _model(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3))


def _batch(rank, step):
    g = torch.Generator().manual_seed(1000 * rank + step)
    return torch.randn(4, 6, generator=g), torch.randn(4, 3, generator=g)


def _flat(model):
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()])


def _worker(rank, world, graph_name, ppi, steps, push_sum, overlap, fused, nesterov,
            ppi_switch=None, jitter_ms=0):
    import random
    import time
    rng = random.Random(97 * rank + 5)
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    from stochastic_gradient_push_b200.optim import FusedGossipSGD
    graph = getattr(sgp, graph_name)(rank, world, peers_per_itr=ppi)
    model = GossipDataParallel(_model(rank), graph=graph, push_sum=push_sum, overlap=overlap,
                               rank=rank, world_size=world, verbose=False)
    assert model.transport == 'c10d'
    if fused:
        opt = FusedGossipSGD(model, lr=LR, momentum=MU, weight_decay=WD, nesterov=nesterov)
    else:
        opt = torch.optim.SGD(model.parameters(), lr=LR, momentum=MU, weight_decay=WD,
                              nesterov=nesterov)
    model.train()
    for step in range(steps):
        if ppi_switch is not None and step == ppi_switch[0]:
            model.update_gossiper('peers_per_itr', ppi_switch[1])
        x, y = _batch(rank, step)
        if jitter_ms:                    # randomised per-rank delays (SURVEY 5.2 stress test)
            time.sleep(rng.uniform(0, jitter_ms) / 1e3)
        loss = ((model(x) - y) ** 2).mean()
        loss.backward()
        opt.step()
        opt.zero_grad()
        if jitter_ms:
            time.sleep(rng.uniform(0, jitter_ms) / 1e3)
        if not overlap:
            model.transfer_params()
    model.sync_comms()
    model.unbias()
    sd = model.state_dict()
    return _flat(model.module).tolist(), float(sd['ps_weight']), sd['is_ps_numerator']


def _simulate(world, graph_name, ppi, steps, overlap, nesterov, ppi_switch=None):
    """All ranks in one process, plain tensors, straight from the algebra in
    SURVEY 3.3/3.4."""
    models = [_model(r) for r in range(world)]
    graphs = [getattr(sgp, graph_name)(r, world, peers_per_itr=ppi) for r in range(world)]
    mix = [sgp.UniformMixing(g, 'cpu') for g in graphs]
    ws = [1.0] * world
    moms = [torch.zeros_like(_flat(m)) for m in models]
    xs = [_flat(m).clone() for m in models]          # numerators
    res = [torch.zeros_like(x) for x in xs]
    wres = [0.0] * world

    def load(m, flat):
        off = 0
        for p in m.parameters():
            n = p.numel()
            p.data.copy_(flat[off:off + n].view_as(p))
            off += n

    for step in range(steps):
        if ppi_switch is not None and step == ppi_switch[0]:
            if overlap:          # a schedule change drains the in-flight gossip first
                xs = [x + r for x, r in zip(xs, res)]
                ws = [w + wr for w, wr in zip(ws, wres)]
                res = [torch.zeros_like(x) for x in xs]
                wres = [0.0] * world
            for g in graphs:
                g.peers_per_itr = ppi_switch[1]
        if overlap:
            # pre-forward: fold residual, pre-scale, snapshot & send
            xs = [x + r for x, r in zip(xs, res)]
            ws = [w + wr for w, wr in zip(ws, wres)]
            cols = [mix[j].scalar_weights(graphs[j].get_peers()[0]) for j in range(world)]
            snap, snapw = [x.clone() for x in xs], list(ws)
            for i in range(world):
                _, ins = graphs[i].get_peers()
                res[i] = sum(cols[j][1][i] * snap[j] for j in ins)
                wres[i] = sum(cols[j][1][i] * snapw[j] for j in ins)
                xs[i] = cols[i][0] * xs[i]
                ws[i] = cols[i][0] * ws[i]
            oracle.rotate_all(graphs)
        grads = []
        for i, m in enumerate(models):
            load(m, xs[i] / ws[i])
            m.zero_grad()
            x, y = _batch(i, step)
            ((m(x) - y) ** 2).mean().backward()
            grads.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]))
        for i in range(world):
            xs[i], moms[i] = oracle.sgd_momentum(xs[i], grads[i], moms[i], LR, MU, WD, nesterov)
        if not overlap:
            xs, ws = oracle.mix_columns(xs, ws, graphs, mix)
            oracle.rotate_all(graphs)
    if overlap:
        xs = [x + r for x, r in zip(xs, res)]
        ws = [w + wr for w, wr in zip(ws, wres)]
    return [x / w for x, w in zip(xs, ws)], ws


@pytest.mark.parametrize('graph_name,ppi,push_sum,fused,nesterov', [
    ('NPeerDynamicDirectedExponentialGraph', 1, True, False, True),
    ('NPeerDynamicDirectedExponentialGraph', 1, True, True, True),
    ('DynamicDirectedExponentialGraph', 2, True, True, False),
    ('RingGraph', 1, False, False, False),
    ('DynamicBipartiteExponentialGraph', 1, False, True, True),
])
def test_sync_gossip_matches_simulation(graph_name, ppi, push_sum, fused, nesterov):
    world, steps = 4, 5
    out = run_distributed(_worker, world, graph_name, ppi, steps, push_sum, False, fused, nesterov)
    want, ws = _simulate(world, graph_name, ppi, steps, False, nesterov)
    for r in range(world):
        got, w, is_num = out[r]
        torch.testing.assert_close(torch.tensor(got), want[r], rtol=1e-4, atol=1e-5)
        assert abs(w - ws[r]) < 1e-5 and is_num is False


@pytest.mark.parametrize('overlap', [False, True])
def test_random_per_rank_delays_do_not_change_the_result(overlap):
    """Race stress: ranks stall for random times before forward and before the
    exchange; synchronous SGP and Overlap-SGP must still reproduce the deterministic simulation
    bit-for-tolerance (no message applied twice, early or to the wrong step)."""
    world, steps = 4, 8
    name = 'NPeerDynamicDirectedExponentialGraph'
    out = run_distributed(_worker, world, name, 1, steps, True, overlap, True, True, None, 30)
    want, ws = _simulate(world, name, 1, steps, overlap, True)
    for r in range(world):
        got, w, _ = out[r]
        torch.testing.assert_close(torch.tensor(got), want[r], rtol=1e-4, atol=1e-5)
        assert abs(w - ws[r]) < 1e-5


@pytest.mark.parametrize('fused', [False, True])
def test_overlap_adds_residual_one_step_late(fused):
    world, steps = 4, 5
    out = run_distributed(_worker, world, 'NPeerDynamicDirectedExponentialGraph', 1, steps,
                          True, True, fused, False)
    want, ws = _simulate(world, 'NPeerDynamicDirectedExponentialGraph', 1, steps, True, False)
    for r in range(world):
        got, w, _ = out[r]
        torch.testing.assert_close(torch.tensor(got), want[r], rtol=1e-4, atol=1e-5)
        assert abs(w - ws[r]) < 1e-5


def _state_roundtrip(rank, world):
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    graph = sgp.RingGraph(rank, world)
    model = GossipDataParallel(_model(rank), graph=graph, rank=rank, world_size=world)
    model.ps_weight = 0.75
    sd = model.state_dict()
    assert set(sd) == {'state_dict', 'ps_weight', 'is_ps_numerator'}
    assert all(k.startswith('module.') for k in sd['state_dict'])
    other = GossipDataParallel(_model(99), graph=sgp.RingGraph(rank, world), rank=rank,
                               world_size=world)
    other.load_state_dict(copy.deepcopy(sd))
    assert abs(float(other.ps_weight) - 0.75) < 1e-7
    assert torch.equal(_flat(other.module), _flat(model.module))
    # parameters are views of one arena
    arena = other.arena
    p0 = next(other.module.parameters())
    assert p0.data_ptr() == arena.flat.data_ptr()
    # transfer_params refuses while the previous gossip is un-mixed
    assert model.transfer_params() is True
    assert model.transfer_params() is False
    model.sync_comms()
    assert model.params_mixed and not model.gossiping
    # eval drains and disables, train re-enables
    model.eval()
    assert model.gossip_enable is False and model.transfer_params() is False
    model.train()
    assert model.gossip_enable is True
    return True


def test_state_dict_and_flags():
    assert all(run_distributed(_state_roundtrip, 2))


def _ppi_update(rank, world):
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    graph = sgp.DynamicDirectedExponentialGraph(rank, world)
    model = GossipDataParallel(_model(rank), graph=graph, rank=rank, world_size=world)
    model.transfer_params()
    model.update_gossiper('peers_per_itr', 2)
    assert graph.peers_per_itr == 2 and graph._group_indices == [0, 1]
    model.update_gossiper('peers_per_itr', 2)      # no-op second time
    for _ in range(6):
        model.transfer_params()
        model.sync_comms()
    model.unbias()
    return _flat(model.module).tolist()


def test_update_gossiper_peers_per_itr_and_consensus():
    world = 4
    out = run_distributed(_ppi_update, world)
    mean = sum(_flat(_model(r)) for r in range(world)) / world
    for r in range(world):
        torch.testing.assert_close(torch.tensor(out[r]), mean, rtol=0, atol=2e-3)


def _async_worker(rank, world, synch_freq, steps):
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    graph = sgp.NPeerDynamicDirectedExponentialGraph(rank, world)
    model = GossipDataParallel(_model(rank), graph=graph, overlap=True, synch_freq=synch_freq,
                               rank=rank, world_size=world)
    assert model.asynch and not model.lazy_mixing
    x = torch.zeros(2, 6)
    for _ in range(steps):
        model(x)                       # pre-forward hook: poll / fold / launch next gossip
        assert model.num_updates <= synch_freq
    model.sync_comms()
    model.unbias()
    return _flat(model.module).tolist(), float(model.ps_weight)


def test_bounded_staleness_conserves_mass_and_mixes():
    world = 4
    out = run_distributed(_async_worker, world, 2, 12)
    x0 = torch.stack([_flat(_model(r)) for r in range(world)])
    # sum_i w_i z_i == sum_i x_i(0): push-sum mass is conserved without gradients
    mass = sum(torch.tensor(z) * w for z, w in out)
    torch.testing.assert_close(mass, x0.sum(0), rtol=1e-4, atol=1e-4)
    assert abs(sum(w for _, w in out) - world) < 1e-4
    spread0 = (x0 - x0.mean(0)).abs().max().item()
    zs = torch.stack([torch.tensor(z) for z, _ in out])
    assert (zs - zs.mean(0)).abs().max().item() < 0.2 * spread0


def _hier_worker(rank, world, steps):
    """2 nodes x 2 processes: only node masters gossip; locals mirror them."""
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    torch.manual_seed(rank)
    model = GossipDataParallel(_model(rank), nprocs_per_node=2, rank=rank, world_size=world)
    assert model.dist_config['world_size'] == 2 and model.dist_config['rank'] == rank // 2
    assert model.is_local_master == (rank % 2 == 0)
    opt = torch.optim.SGD(model.parameters(), lr=LR)
    for step in range(steps):
        x, y = _batch(rank, step)
        ((model(x) - y) ** 2).mean().backward()
        opt.step()
        opt.zero_grad()
        model.transfer_params()
    model.sync_comms()
    model.unbias()
    x, _ = _batch(0, 99)
    model(x)                           # forward re-broadcasts the master's parameters
    return _flat(model.module).tolist(), [p.grad is None for p in model.parameters()]


def test_hierarchical_nprocs_per_node():
    out = run_distributed(_hier_worker, 4, 3)
    p = [torch.tensor(o[0]) for o in out]
    torch.testing.assert_close(p[0], p[1], rtol=0, atol=0)      # same node -> identical
    torch.testing.assert_close(p[2], p[3], rtol=0, atol=0)
    # two nodes on a 1-peer graph average exactly at every mix
    torch.testing.assert_close(p[0], p[2], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('overlap', [False, True])
def test_peers_per_itr_switch_mid_training_matches_simulation(overlap):
    world, steps, switch = 4, 6, (3, 2)
    name = 'DynamicDirectedExponentialGraph'
    out = run_distributed(_worker, world, name, 1, steps, True, overlap, True, False, switch)
    want, ws = _simulate(world, name, 1, steps, overlap, False, switch)
    for r in range(world):
        got, w, _ = out[r]
        torch.testing.assert_close(torch.tensor(got), want[r], rtol=1e-4, atol=1e-5)
        assert abs(w - ws[r]) < 1e-5


def _heartbeat_worker(rank, world):
    """Fault injection: rank 1 is 3 s late.  Rank 0's 1 s heartbeat must fire as the reference's
    ``NameError('Gossip flag timeout')`` instead of hanging, and the exchange must
    still be completable once the slow peer shows up (nothing was lost or double-counted)."""
    import time
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    graph = sgp.RingGraph(rank, world)
    model = GossipDataParallel(_model(rank), graph=graph, rank=rank, world_size=world,
                               heartbeat_timeout=1.0)
    before = _flat(model.module).clone()
    if rank == 1:
        time.sleep(3.0)
    t0 = time.time()
    model.transfer_params()
    caught, waited = None, 0.0
    try:
        model.sync_comms()
    except NameError as e:
        caught, waited = str(e), time.time() - t0
        assert model.gossiping and not model.params_mixed      # state intact: retry is possible
        model._timeout_s = 60.0
        model.sync_comms()
    model.unbias()
    return caught, waited, before.tolist(), _flat(model.module).tolist()


def test_heartbeat_timeout_detects_a_slow_peer_and_recovers():
    out = run_distributed(_heartbeat_worker, 2, timeout=120)
    caught0, waited0, b0, a0 = out[0]
    caught1, _, b1, a1 = out[1]
    assert caught0 == 'Gossip flag timeout' and 0.9 <= waited0 < 2.5
    assert caught1 is None
    mean = [(x + y) / 2 for x, y in zip(b0, b1)]
    assert torch.allclose(torch.tensor(a0), torch.tensor(mean), atol=1e-6)
    assert torch.allclose(torch.tensor(a1), torch.tensor(mean), atol=1e-6)


# --------------------------------------------------------------------------- #
# soft retry of an interrupted gossip round (reference gossip/distributed.py:358-364, 494-504)
# --------------------------------------------------------------------------- #
class _BrokenOnce(object):
    """transport wrapper: during round `bad_round` nothing is sent and every receive fails with a
    RuntimeError -- what a communicator error looks like to the gossip state machine"""

    class _Failed(object):
        def wait(self):
            raise RuntimeError('injected communicator failure')

        def is_completed(self):
            return True

    class _Done(object):
        def wait(self):
            return True

        def is_completed(self):
            return True

    def __init__(self, inner, bad_round):
        self.inner, self.bad_round, self.round = inner, bad_round, 0
        self.group = inner.group

    def post_recvs(self, buffers, in_edges):
        self.round += 1
        if self.round == self.bad_round:
            return [self._Failed() for _ in buffers]
        return self.inner.post_recvs(buffers, in_edges)

    def post_sends(self, msgs, out_edges):
        if self.round == self.bad_round:
            return [self._Done() for _ in msgs]
        return self.inner.post_sends(msgs, out_edges)

    def post_polled_recv(self, buf, in_edge):
        return self.post_recvs([buf], [in_edge])[0]


def _retry_worker(rank, world, bad_round, rounds):
    import torch.distributed as dist
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    net = nn.Linear(4, 4, bias=False)
    with torch.no_grad():
        net.weight.fill_(float(rank + 1))
    graph = sgp.RingGraph(rank, world, peers_per_itr=1)
    model = GossipDataParallel(net, graph=graph, rank=rank, world_size=world, verbose=False, push_sum=True)
    model._c10d.transport = _BrokenOnce(model._c10d.transport, bad_round)
    model.train()
    for _ in range(rounds):
        model.transfer_params()
        model._query_gossip_queue()
        if model.gossiping:                   # the interrupted round was re-queued: finish the retry
            model._query_gossip_queue()
    model.unbias()
    total = net.weight.detach().clone() * float(model.ps_weight)
    dist.all_reduce(total)
    wsum = torch.tensor([float(model.ps_weight)])
    dist.all_reduce(wsum)
    return model.gossip_retries, float(total[0, 0]), float(wsum), float(net.weight[0, 0])


def test_interrupted_gossip_round_is_requeued_and_mass_is_conserved():
    world, rounds = 2, 4
    out = run_distributed(_retry_worker, world, 2, rounds)
    for retries, total, wsum, _ in out:
        assert retries == 1
        assert abs(total - (1.0 + 2.0)) < 1e-5          # sum of push-sum numerators: conserved
        assert abs(wsum - world) < 1e-6                 # sum of push-sum weights: conserved
    # and the ranks still contract to the average
    assert abs(out[0][3] - 1.5) < 0.2 and abs(out[1][3] - 1.5) < 0.2
