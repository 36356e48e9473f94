"""Gossipers over the c10d (gloo) transport; golden values from SURVEY 4.4."""
import pytest
import torch

import stochastic_gradient_push_b200 as sgp
from stochastic_gradient_push_b200.gossiper import BilatPushPull, PushPull, PushSum

from dist_utils import run_distributed


def _pushsum_trace(rank, world, graph_name, ppi, steps, residual):
    graph = getattr(sgp, graph_name)(rank, world, peers_per_itr=ppi)
    x = torch.full((5,), 10.0 * rank)
    g = PushSum(x, graph, rank=rank, world_size=world)
    w = torch.ones(1)
    trace = []
    for _ in range(steps):
        if residual:     # caller pre-scales by lo and folds the residual in
            lo = g.mixing_weights['lo'].item() if False else 1.0 / (len(g.out_edges) + 1)
            x = x * lo
            w = w * lo
            r, wr = g.mix(x.clone(), w, residual=True)
            x = x + r
            w = w + wr
        else:
            x, _ = g.mix(x.clone(), w, residual=False)
            x = x.clone()
        trace.append((round(x[0].item(), 4), round(float(w), 6)))
    return trace


def test_two_rank_ring_pushsum_converges_to_mean():
    out = run_distributed(_pushsum_trace, 2, 'RingGraph', 1, 1, False)
    assert out[0][0][0] == 5.0 and out[1][0][0] == 5.0


@pytest.mark.parametrize('residual', [False, True])
def test_pushsum_npdde_n4_golden(residual):
    out = run_distributed(_pushsum_trace, 4, 'NPeerDynamicDirectedExponentialGraph', 1, 3, residual)
    want = {0: [15, 15, 15], 1: [5, 15, 15], 2: [15, 15, 15], 3: [25, 15, 15]}
    for r in range(4):
        assert [v for v, _ in out[r]] == want[r]
        assert all(abs(w - 1.0) < 1e-6 for _, w in out[r])


def test_pushsum_npdde_n8_golden():
    out = run_distributed(_pushsum_trace, 8, 'NPeerDynamicDirectedExponentialGraph', 1, 3, True)
    want = {0: [35, 45, 35], 1: [5, 35, 35], 2: [15, 25, 35], 3: [25, 15, 35],
            4: [35, 25, 35], 5: [45, 35, 35], 6: [55, 45, 35], 7: [65, 55, 35]}
    for r in range(8):
        assert [v for v, _ in out[r]] == want[r]


def _pushpull_trace(rank, world, graph_name, steps):
    graph = getattr(sgp, graph_name)(rank, world)
    x = torch.full((3,), 10.0 * rank)
    g = PushPull(x, graph, rank=rank, world_size=world)
    trace = []
    for _ in range(steps):
        lo = 1.0 / (len(g.out_edges) + 1)
        x = x * lo
        r, _ = g.mix(x.clone(), torch.ones(1), residual=True)
        x = x + r
        trace.append(round(x[0].item(), 4))
    return trace


def test_pushpull_dbe_n4_golden():
    out = run_distributed(_pushpull_trace, 4, 'DynamicBipartiteExponentialGraph', 3)
    assert out == [[15, 10, 10], [5, 10, 15], [15, 20, 20], [25, 20, 15]]


def test_pushpull_ring_does_not_deadlock():
    """The reference deadlocks here on gloo (all ranks active, blocking sends)."""
    out = run_distributed(_pushpull_trace, 4, 'RingGraph', 2)
    # static ring, ppi=1: x_i <- (x_i + x_{i-1}) / 2
    assert out[0][0] == 15.0 and out[1][0] == 5.0


def _column_stochastic(rank, world, ppi, steps):
    graph = sgp.DynamicDirectedExponentialGraph(rank, world, peers_per_itr=ppi)
    mixing = sgp.SelfWeightedMixing(graph, self_weight=0.5)
    torch.manual_seed(rank)
    x = torch.randn(7)
    g = PushSum(x, graph, mixing=mixing, rank=rank, world_size=world)
    assert not g.regular
    w = torch.ones(1)
    sums = []
    import torch.distributed as dist
    for _ in range(steps):
        x, w = g.mix(x.clone(), w, residual=False)
        x, w = x.clone(), w.clone().reshape(1)
        tot = torch.cat([x, w])
        dist.all_reduce(tot)
        sums.append(tot.tolist())
    return sums, (x / w).tolist()


def test_irregular_mixing_conserves_mass_and_debiases():
    world = 4
    out = run_distributed(_column_stochastic, world, 1, 12)
    x0 = torch.stack([torch.randn(7, generator=torch.Generator().manual_seed(r))
                      for r in range(world)])
    total = x0.sum(0)
    for sums, _ in out:
        for tot in sums:
            tot = torch.tensor(tot)
            assert torch.allclose(tot[:-1], total, atol=1e-4)
            assert abs(tot[-1].item() - world) < 1e-4          # sum of ps-weights
    for _, z in out:                                            # z -> global mean
        assert torch.allclose(torch.tensor(z), total / world, atol=5e-2)


def _bilat(rank, world, rounds):
    import time
    graph = sgp.DynamicBipartiteExponentialGraph(rank, world)
    x = torch.full((4,), float(rank))
    g = BilatPushPull(x, graph, rank=rank, world_size=world)
    polls, done = 0, 0
    while done < rounds:
        out = g.mix(x.clone())
        if isinstance(out[1], bool) and out[1] is False:
            polls += 1
            assert g.passive
            time.sleep(0.001)
            continue
        in_msg, _ = out
        x = (x + in_msg) * 0.5
        done += 1
    return x[0].item(), polls, g.passive


def test_bilat_pushpull_handshake():
    out = run_distributed(_bilat, 4, 2)
    passive = [o[2] for o in out]
    assert passive == [True, False, True, False]
    vals = [o[0] for o in out]
    assert abs(sum(vals) - 6.0) < 1e-5     # pairwise averaging conserves the sum


def test_uniform_mixing_scales_out_msg_in_place():
    graph = sgp.RingGraph(0, 2)
    g = PushSum(torch.zeros(2), graph, rank=0, world_size=2)
    msg = torch.ones(2)
    outs = list(g.mix_out_msg_(msg, 1.0, residual=False))
    assert torch.allclose(outs[0], torch.full((2,), 0.5))   # loop-back copy
    assert outs[1] is msg and torch.allclose(msg, torch.full((2,), 0.5))
