"""Protocol-level checks of the device-side AD-PSGD handshake on its executable model
(``ops/bilat_model.py`` mirrors ``sgp_bilat_decide_kernel`` + the ``SGP_F_FROM_STATE`` worker): random
interleavings of the ranks' launches, random partner delays, random gradient arrivals."""
import random

import pytest

from stochastic_gradient_push_b200.ops.bilat_model import BilatRank


def _bipartite_partner(world):
    """dynamic bipartite exponential pairing: even ranks are active, odd ranks passive; round r pairs
    an even rank i with the odd rank (i + 1 + 2*(r % (world/2))) % world -- symmetric by construction"""
    half = world // 2

    def partner(r, rank):
        shift = 1 + 2 * (r % half)
        return (rank + shift) % world if rank % 2 == 0 else (rank - shift) % world
    return partner


def _make(world, budget, seed):
    rng = random.Random(seed)
    part = _bipartite_partner(world)
    ranks = [BilatRank(r, part, passive=(r % 2 == 1), x=rng.uniform(-5, 5), budget=budget) for r in range(world)]
    return ranks, rng


@pytest.mark.parametrize('world', [2, 4, 8])
@pytest.mark.parametrize('seed', range(6))
def test_random_interleavings_average_pairwise_and_conserve_mass(world, seed):
    ranks, rng = _make(world, None, seed)
    total = sum(r.x for r in ranks)
    for _ in range(4000):
        r = rng.choice(ranks)
        # a daemon launch pair of ONE rank; the partner may or may not show up within the bounded wait
        r.decide(ranks, partner_ready_within_wait=rng.random() < 0.7)
        if rng.random() < 0.9:                  # (the worker always follows in stream order; other ranks interleave)
            r.work(ranks)
        else:
            other = rng.choice(ranks)
            other.decide(ranks)
            other.work(ranks)
            r.work(ranks)
    assert all(r.overwrites_while_unread == 0 for r in ranks)           # WAR fence holds
    assert abs(sum(r.x for r in ranks) - total) < 1e-9 * max(1.0, abs(total)) + 1e-9   # mass conserved
    steps = [r.step for r in ranks]
    assert min(steps) >= 5                                              # progress: nobody starves
    # every completed round was executed by BOTH partners with each other's round-r snapshot
    done = {(rec[0], r.rank): rec for r in ranks for rec in r.rounds_done}
    for (rnd, rank), (_, partner, _, _) in done.items():
        if (rnd, partner) in done:
            assert done[(rnd, partner)][1] == rank
    # consensus: the spread shrinks
    assert max(r.x for r in ranks) - min(r.x for r in ranks) < 1e-3


@pytest.mark.parametrize('seed', range(4))
def test_budget_bounds_rounds_between_gradients(seed):
    world, k = 4, 2
    ranks, rng = _make(world, k, seed)
    for _ in range(2000):
        r = rng.choice(ranks)
        r.decide(ranks)
        r.work(ranks)
    # without gradients a rank STARTS at most k rounds; a started round may complete later
    assert all(r.step <= k for r in ranks), [r.step for r in ranks]
    assert all(r.budget == 0 for r in ranks if not r.passive)
    before = [r.step for r in ranks]
    for r in ranks:
        r.apply_gradient(0.25, k)               # a gradient arrives everywhere: budget refilled
    for _ in range(2000):
        r = rng.choice(ranks)
        r.decide(ranks)
        r.work(ranks)
    assert all(b < r.step <= b + k for b, r in zip(before, ranks))


def test_slow_partner_never_blocks_the_fast_one_and_catches_up():
    ranks, rng = _make(2, None, 0)
    a, b = ranks
    for _ in range(50):                          # rank 1 (passive) is away: rank 0 keeps launching
        a.decide(ranks, partner_ready_within_wait=False)
        a.work(ranks)
    assert a.published and a.step == 0           # published once, never stuck in a kernel, no progress
    b.decide(ranks)
    b.work(ranks)                                # the passive rank shows up: publishes AND pulls in one launch
    assert b.step == 1
    a.decide(ranks)
    a.work(ranks)
    assert a.step == 1 and abs(a.x - b.x) < 1e-12
