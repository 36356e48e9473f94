"""C10dTransport's NCCL-oriented modes, exercised over gloo (the control flow is backend-neutral;
what NCCL adds is the grouped launch inside ``batch_isend_irecv``):

* ``batched=True``: receives and sends of one exchange go through ONE ``batch_isend_irecv`` call
  (on NCCL: no "irecv then isend on both sides" deadlock for symmetric exchanges);
* ``reverse_group``: messages from a higher to a lower rank use a second process group, so sends
  and receives that are NOT issued together (AD-PSGD loop) never share a communicator per direction.
"""
import pytest
import torch
import torch.distributed as dist

import stochastic_gradient_push_b200 as sgp
from stochastic_gradient_push_b200.gossiper import BilatPushPull, C10dTransport, PushPull, PushSum

from dist_utils import run_distributed


def _pushsum(rank, world, graph_name, ppi, steps, batched, split):
    graph = getattr(sgp, graph_name)(rank, world, peers_per_itr=ppi)
    rev = dist.new_group(list(range(world))) if split else None
    tr = C10dTransport(None, batched=batched, reverse_group=rev)
    assert tr.batched == batched
    x = torch.full((5,), 10.0 * rank)
    g = PushSum(x, graph, rank=rank, world_size=world, transport=tr)
    w = torch.ones(1)
    trace = []
    for _ in range(steps):
        lo = 1.0 / (len(g.out_edges) + 1)
        x, w = x * lo, w * lo
        r, wr = g.mix(x.clone(), w, residual=True)
        x, w = x + r, w + wr
        trace.append(round(x[0].item(), 4))
    return trace


@pytest.mark.parametrize('batched,split', [(True, False), (False, True), (True, True)])
def test_pushsum_golden_values_in_every_transport_mode(batched, split):
    out = run_distributed(_pushsum, 8, 'NPeerDynamicDirectedExponentialGraph', 1, 3, batched, split)
    want = {0: [35, 45, 35], 1: [5, 35, 35], 2: [15, 25, 35], 3: [25, 15, 35],
            4: [35, 25, 35], 5: [45, 35, 35], 6: [55, 45, 35], 7: [65, 55, 35]}
    for r in range(8):
        assert out[r] == want[r]


def test_batched_symmetric_exchange_two_ranks_and_duplicate_edges():
    """in-peer == out-peer (the NCCL deadlock shape) and ppi = 2 on two ranks (duplicate edges
    between one pair + a self-edge): same numbers as the op-by-op path"""
    a = run_distributed(_pushsum, 2, 'DynamicDirectedExponentialGraph', 2, 3, True, False)
    b = run_distributed(_pushsum, 2, 'DynamicDirectedExponentialGraph', 2, 3, False, False)
    assert a == b


def _pushpull(rank, world, batched):
    graph = sgp.DynamicBipartiteExponentialGraph(rank, world)
    x = torch.full((3,), 10.0 * rank)
    g = PushPull(x, graph, rank=rank, world_size=world, transport=C10dTransport(None, batched=batched))
    trace = []
    for _ in range(3):
        lo = 1.0 / (len(g.out_edges) + 1)
        x = x * lo
        r, _ = g.mix(x.clone(), torch.ones(1), residual=True)
        x = x + r
        trace.append(round(x[0].item(), 4))
    return trace


def test_batched_pushpull_golden():
    out = run_distributed(_pushpull, 4, True)
    assert out == [[15, 10, 10], [5, 10, 15], [15, 20, 20], [25, 20, 15]]


def _bilat(rank, world, batched, split):
    import time
    graph = sgp.DynamicBipartiteExponentialGraph(rank, world)
    rev = dist.new_group(list(range(world))) if split else None
    x = torch.full((4,), float(rank))
    g = BilatPushPull(x, graph, rank=rank, world_size=world,
                      transport=C10dTransport(None, batched=batched, reverse_group=rev))
    done, t0 = None, time.time()
    while done is None and time.time() - t0 < 60:
        in_msg, ok = g.mix(x.clone())
        if ok is not False:
            done = in_msg.clone()
        else:
            time.sleep(0.005)
    return None if done is None else float(done[0])


@pytest.mark.parametrize('batched,split', [(False, True), (True, False)])
def test_bilateral_handshake_in_every_transport_mode(batched, split):
    # (batched=True over gloo still uses the polled receive of the passive side: `post_polled_recv`
    # returns a raw request only when the transport is batched, and gloo requests cannot be polled
    # -- so the passive rank of this test runs un-batched)
    out = run_distributed(_bilat_mixed if batched else _bilat, 2, batched, split)
    assert out[0] == 1.0 and out[1] == 0.0          # each rank received its partner's value


def _bilat_mixed(rank, world, batched, split):
    graph = sgp.DynamicBipartiteExponentialGraph(rank, world)
    return _bilat(rank, world, batched and not graph.is_passive(), split)


def _gdp(rank, world, batched):
    import test_distributed_c10d as sim
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    from stochastic_gradient_push_b200.optim import FusedGossipSGD
    graph = sgp.DynamicDirectedExponentialGraph(rank, world, peers_per_itr=2)
    model = GossipDataParallel(sim._model(rank), graph=graph, push_sum=True, rank=rank, world_size=world)
    model._c10d.transport = C10dTransport(None, batched=batched)
    opt = FusedGossipSGD(model, lr=sim.LR, momentum=sim.MU, weight_decay=sim.WD, nesterov=True)
    model.train()
    for step in range(5):
        x, y = sim._batch(rank, step)
        ((model(x) - y) ** 2).mean().backward()
        opt.step()
        opt.zero_grad()
        model.transfer_params()
    model.sync_comms()
    model.unbias()
    return sim._flat(model.module).tolist(), float(model.state_dict()['ps_weight'])


def test_gossip_data_parallel_batched_exchange_matches_simulation():
    import test_distributed_c10d as sim
    world = 4
    got = run_distributed(_gdp, world, True)
    want, ws = sim._simulate(world, 'DynamicDirectedExponentialGraph', 2, 5, False, True)
    for r in range(world):
        torch.testing.assert_close(torch.tensor(got[r][0]), want[r], rtol=1e-4, atol=1e-5)
        assert abs(got[r][1] - ws[r]) < 1e-5
