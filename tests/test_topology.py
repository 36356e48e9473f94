"""Golden tables from SURVEY.md section 4.4 (captured from the reference's
GraphManager with a stubbed Edge)."""
import pytest

import stochastic_gradient_push_b200 as sgp
from stochastic_gradient_push_b200.topology import make_schedule, MAX_PEERS_PER_ITR

DDE, DBE = sgp.DynamicDirectedExponentialGraph, sgp.DynamicBipartiteExponentialGraph
DDL, DBL = sgp.DynamicDirectedLinearGraph, sgp.DynamicBipartiteLinearGraph
RING, NP = sgp.RingGraph, sgp.NPeerDynamicDirectedExponentialGraph

GOLDEN = {
    2: {DDE: [1, 1], DBE: [1, 1], DDL: [1, 1], DBL: [1, 1], RING: [1, 1]},
    4: {DDE: [1, 3, 2, 2], DBE: [1, 3, 3, 1], DDL: [1, 3, 3, 1],
        DBL: [1, 3, 3, 1], RING: [1, 3]},
    8: {DDE: [1, 7, 2, 6, 4, 4], DBE: [1, 7, 3, 5, 5, 3],
        DDL: [1, 7, 3, 5, 5, 3, 7, 1], DBL: [1, 7, 3, 5, 5, 3, 7, 1],
        RING: [1, 7]},
}
GOLDEN_NP = {(2, 1): [1], (2, 2): [1, 0], (4, 1): [1, 2], (4, 2): [1, 2, 3, 2],
             (8, 1): [1, 2, 4], (8, 2): [1, 2, 3, 6]}


def book0(g):
    return [e.dest for e in g.phone_book[0]]


@pytest.mark.parametrize('n', [2, 4, 8])
def test_phone_books_match_reference(n):
    for cls, want in GOLDEN[n].items():
        assert book0(cls(0, n)) == want, cls.__name__


@pytest.mark.parametrize('n,ppi', sorted(GOLDEN_NP))
def test_npdde_phone_book(n, ppi):
    assert book0(NP(0, n, peers_per_itr=ppi)) == GOLDEN_NP[(n, ppi)]


def test_books_are_circulant_for_even_n():
    for cls in (DDE, DBE, DDL, DBL, RING, NP):
        for n in (2, 4, 8, 16):
            g = cls(0, n)
            offs = [(e.dest - 0) % n for e in g.phone_book[0]]
            for r in range(n):
                assert [(e.dest - r) % n for e in g.phone_book[r]] == offs


def test_rank0_sequence_npdde_n8():
    g = NP(0, 8, peers_per_itr=1)
    seq = [g.get_peers()]
    for _ in range(3):
        seq.append(g.get_peers(rotate=True))
    assert seq == [([1], [7]), ([2], [6]), ([4], [4]), ([1], [7])]
    g2 = NP(0, 8, peers_per_itr=2)
    assert g2.get_peers() == ([1, 2], [7, 6])
    assert g2.get_peers(rotate=True) == ([3, 6], [5, 2])


def test_ring_static_and_ppi2():
    g = RING(0, 8)
    assert g.get_peers() == ([1], [7])
    assert not g.is_dynamic_graph()
    g.peers_per_itr = 2
    assert g.get_peers() == ([1, 7], [7, 1])
    assert g.period == 1 and len(g.phases()) == 1


def test_in_out_duality():
    """j in out(i, t)  <=>  i in in(j, t), at every phase."""
    for cls in (DDE, DBE, DDL, DBL, RING, NP):
        n = 8
        graphs = [cls(r, n) for r in range(n)]
        for _ in range(graphs[0].period + 1):
            peers = [g.get_peers() for g in graphs]
            for i in range(n):
                for j in peers[i][0]:
                    if j != i:
                        assert i in peers[j][1]
                for j in peers[i][1]:
                    assert i in peers[j][0]
            for g in graphs:
                if g.is_dynamic_graph():
                    g.get_peers(rotate=True)


def test_peers_per_itr_setter_resets_window():
    g = DDE(0, 8)
    g.get_peers(rotate=True)
    assert g._group_indices == [1]
    g.peers_per_itr = 2
    assert g._group_indices == [0, 1]
    assert g.get_peers() == ([1, 7], [7, 1])


def test_predicates():
    for cls in (DDE, DBE, DDL, DBL, RING, NP):
        assert cls(0, 4).is_regular_graph()
    assert DBE(0, 4).is_bipartite_graph() and DBL(0, 4).is_bipartite_graph()
    assert not DDE(0, 4).is_bipartite_graph()
    assert DBE(0, 4).is_passive() and not DBE(1, 4).is_passive()
    assert DBE(1, 4).is_passive(rank=2)
    assert not NP(0, 4).is_passive()


def test_dedupe_flag_gives_intended_schedule():
    g = DDE(0, 8, dedupe=True)
    assert book0(g) == [1, 7, 2, 6, 4]
    assert g.period == 5
    assert DDE(0, 8).period == 6          # reference quirk: 4 visited twice


def test_world_size_one_is_peerless():
    g = NP(0, 1)
    assert g.get_peers() == ([], [])
    assert g.device_table().shape == (1, 2 + 2 * MAX_PEERS_PER_ITR)
    assert g.device_table()[0, :2].tolist() == [0, 0]


def test_device_table_matches_python_rotation():
    for cls in (DDE, NP, DBE, RING):
        for ppi in (1, 2):
            g = cls(3, 8, peers_per_itr=ppi)
            tab = g.device_table()
            assert tab.shape[0] == g.period
            for t in range(g.period):
                outs, ins = g.get_peers()
                n_in, n_out = tab[t, 0].item(), tab[t, 1].item()
                assert tab[t, 2:2 + n_in].tolist() == ins
                assert tab[t, 2 + MAX_PEERS_PER_ITR:2 + MAX_PEERS_PER_ITR + n_out].tolist() == outs
                assert g.phase_index() == t
                if g.is_dynamic_graph():
                    g.get_peers(rotate=True)


def test_nprocs_per_node_scales_process_ranks():
    g = NP(1, 4, nprocs_per_node=2)
    outs, ins = g.get_peers()
    assert outs == [4] and ins == [0]
    assert g.phone_book[1][0].src == 2 and g.phone_book[1][0].dest == 4


def test_schedule_pure_math_no_torch_distributed():
    s = make_schedule('npdde', 0, 8, 1)
    assert s.books[0] == [1, 2, 4]
    assert s.period == 3
