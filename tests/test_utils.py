import math

import pytest
import torch

import stochastic_gradient_push_b200 as sgp
from stochastic_gradient_push_b200.utils import (
    FlatArena, Meter, communicate, contiguous_span, flatten_tensors,
    group_by_dtype, is_power_of, make_logger, unflatten_tensors)


def test_flatten_unflatten_roundtrip_and_aliasing():
    ts = [torch.randn(3, 4), torch.randn(5), torch.randn(2, 2, 2)]
    flat = flatten_tensors(ts)
    assert flat.shape == (12 + 5 + 8,)
    views = unflatten_tensors(flat, ts)
    for v, t in zip(views, ts):
        assert torch.equal(v, t)
    views[0].zero_()                      # views alias the flat buffer
    assert flat[:12].abs().sum() == 0
    single = flatten_tensors([ts[1]])
    single.zero_()                        # single tensor -> clone, not alias
    assert ts[1].abs().sum() > 0


def test_group_by_dtype():
    g = group_by_dtype([torch.zeros(1), torch.zeros(1).half(), torch.ones(2)])
    assert len(g[torch.float32]) == 2 and len(g[torch.float16]) == 1


def test_communicate_scatter_back():
    ts = [torch.ones(3), torch.ones(2, 2)]
    communicate(ts, lambda tensor: tensor.mul_(3))
    assert all(torch.all(t == 3) for t in ts)


def test_is_power_of():
    assert is_power_of(8, 2) and is_power_of(1, 2) and is_power_of(27, 3)
    assert not is_power_of(6, 2) and not is_power_of(5, 1)
    assert is_power_of(3 ** 20, 3)


def test_meter_stats_and_str():
    m = Meter(ptag='T')
    vals = [1.0, 2.0, 4.0, 7.0]
    for v in vals:
        m.update(v)
    mean = sum(vals) / 4
    std = math.sqrt(sum((v - mean) ** 2 for v in vals) / 3)
    assert m.val == 7.0 and abs(m.avg - mean) < 1e-12 and abs(m.std - std) < 1e-9
    assert str(m) == '7.000,%.3f,%.3f' % (mean, std)
    s = Meter(ptag='G', stateful=True, csv_format=False)
    for v in vals:
        s.update(v)
    mad = sum(abs(v - mean) for v in vals) / 4
    assert abs(s.mad - mad) < 1e-12
    assert str(s) == 'G: 7.000 (%.3f +- %.3f)' % (mean, mad)


def test_meter_rehydrate_from_dict():
    m = Meter(ptag='Time')
    m.update(3.0)
    m.update(5.0)
    m2 = Meter(m.__dict__)
    assert (m2.avg, m2.count, m2.ptag) == (m.avg, m.count, 'Time')
    m2.update(1.0)
    assert m2.count == 3


def test_logger_format(capsys):
    lg = make_logger(7, verbose=True, name='test-logger-fmt')
    lg.info('hello')
    out = capsys.readouterr().out
    assert out.startswith('7: INFO -- MainThread -- hello')


def test_arena_adopts_params_as_views():
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    ref = [p.detach().clone() for p in model.parameters()]
    params = list(model.parameters())
    arena = FlatArena(params)
    arena.adopt(params)
    assert arena.total % 4096 == 0 and arena.payload == sum(p.numel() for p in ref)
    for p, r, off in zip(params, ref, arena.offsets):
        assert torch.equal(p, r)
        assert off % 64 == 0
        assert p.data_ptr() == arena.flat.data_ptr() + off * 4
    arena.flat.mul_(2)                    # the model sees arena writes
    assert torch.equal(params[0], ref[0] * 2)
    out = model(torch.randn(2, 5))
    out.sum().backward()                  # autograd still works on the views
    assert params[0].grad is not None
    span = contiguous_span([arena.views[0]])
    assert span.data_ptr() == arena.flat.data_ptr()


def test_arena_flat_grads_accumulate_in_place():
    model = torch.nn.Linear(4, 4)
    params = list(model.parameters())
    arena = FlatArena(params)
    arena.adopt(params)
    gflat = arena.new_buffer()
    arena.bind_grads(params, gflat)
    model(torch.ones(1, 4)).sum().backward()
    assert gflat.abs().sum() > 0
    assert params[0].grad.data_ptr() == gflat.data_ptr()
    before = gflat.clone()
    model(torch.ones(1, 4)).sum().backward()
    assert torch.allclose(gflat, 2 * before)


def test_mixing_weights():
    g = sgp.NPeerDynamicDirectedExponentialGraph(0, 8, peers_per_itr=2)
    m = sgp.UniformMixing(g, torch.device('cpu'))
    w = m.get_mixing_weights(residual_adjusted=False)
    assert abs(w['lo'].item() - 1 / 3) < 1e-7 and abs(w['uniform'].item() - 1 / 3) < 1e-7
    assert set(k for k in w if isinstance(k, int)) == {1, 2}
    wr = m.get_mixing_weights(residual_adjusted=True)
    assert abs(wr['uniform'].item() - 1.0) < 1e-6 and abs(wr['lo'].item() - 1 / 3) < 1e-7
    assert m.is_regular() and m.is_uniform()
    s = sgp.SelfWeightedMixing(g, self_weight=0.6)
    lo, edges = s.scalar_weights()
    assert abs(lo + sum(edges.values()) - 1.0) < 1e-12 and not s.is_regular()
