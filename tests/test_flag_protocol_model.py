"""Random-skew exploration of the publish / pull / ack protocol of the gossip kernels on its
executable model (``ops/flag_model.py``), over every shipped topology -- SURVEY 5.2 (race detection)
at the protocol level: no torn reads, no premature overwrite, no deadlock, bounded skew."""
import random

import pytest

import stochastic_gradient_push_b200 as sgp
from stochastic_gradient_push_b200.ops.flag_model import FlagRank

TOPOLOGIES = ['NPeerDynamicDirectedExponentialGraph', 'DynamicDirectedExponentialGraph',
              'DynamicBipartiteExponentialGraph', 'DynamicDirectedLinearGraph', 'DynamicBipartiteLinearGraph',
              'RingGraph']


def _peer_fn(name, world, ppi):
    graphs = [getattr(sgp, name)(r, world, peers_per_itr=ppi) for r in range(world)]
    tables = [g.phases() for g in graphs]

    def peers(step, rank):
        ph = tables[rank]
        outs, ins = ph[step % len(ph)]
        return list(outs), list(ins)
    return peers


@pytest.mark.parametrize('name', TOPOLOGIES)
@pytest.mark.parametrize('world,ppi', [(2, 1), (4, 1), (8, 1), (8, 2), (6, 1)])
@pytest.mark.parametrize('overlap', [False, True])
def test_random_skew_never_tears_an_outbox(name, world, ppi, overlap):
    if 'Bipartite' in name and world % 2:
        pytest.skip('bipartite graphs need an even world')
    try:
        peers = _peer_fn(name, world, ppi)
    except Exception as e:                       # (a topology may reject a world / ppi combination)
        pytest.skip(str(e))
    rng = random.Random(hash((name, world, ppi, overlap)) & 0xFFFF)
    ranks = [FlagRank(r, peers) for r in range(world)]
    target, idle = 40, 0
    while min(r.pulled for r in ranks) < target:
        r = rng.choice(ranks)
        moved = False
        # each scheduler tick lets ONE rank take ONE micro-step, in random preference order
        for action in rng.sample(['publish', 'pull'], 2):
            if action == 'publish' and r.step < target and r.can_publish(ranks, overlap):
                r.publish(ranks)
                moved = True
                break
            if action == 'pull' and r.can_pull(ranks):
                r.pull(ranks)
                moved = True
                break
        idle = 0 if moved else idle + 1
        assert idle < 20000, 'deadlock: steps %s pulled %s' % ([x.step for x in ranks], [x.pulled for x in ranks])
        # bounded drift: every pull needs the in-neighbour's publish of the same step and every
        # publish needs the own previous pull, so the skew is bounded by the graph's reach
        assert max(x.step for x in ranks) - min(x.pulled for x in ranks) <= 2 * world + 2
    for r in ranks:
        assert not r.errors, r.errors[:3]
