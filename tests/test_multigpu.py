"""One process per GPU: IPC rendezvous + P2P kernels over NVLink."""
import pytest
import torch

import stochastic_gradient_push_b200 as sgp
from stochastic_gradient_push_b200.ops import oracle

from dist_utils import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _sgp_worker(rank, world, graph_name, ppi, steps, numel):
    import torch.distributed as dist
    from stochastic_gradient_push_b200.parallel.symmetric import SymmetricWorld
    from stochastic_gradient_push_b200.ops.peer_mix import GossipEngine
    dev = torch.device('cuda', rank)
    torch.manual_seed(1234 + rank)
    graph = getattr(sgp, graph_name)(rank, world, peers_per_itr=ppi)
    mixing = sgp.UniformMixing(graph, dev)
    z = torch.randn(numel, device=dev)
    grad = torch.randn(numel, device=dev)
    mom = torch.zeros(numel, device=dev)
    sw = SymmetricWorld(dev)
    eng = GossipEngine(sw, z, graph, mixing, grad=grad, momentum=mom, timeout_s=20.0)
    lr, mu, wd, nest = 0.1, 0.9, 1e-4, True
    eng.set_hyper(lr, mu, wd, nest)

    # oracle state: every rank simulates the whole world from gathered tensors
    def gather(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.double() for o in out]

    zs, ms, ws = gather(z), gather(mom), [1.0] * world
    ogs = [getattr(sgp, graph_name)(r, world, peers_per_itr=ppi) for r in range(world)]
    oms = [sgp.UniformMixing(g, 'cpu') for g in ogs]
    for step in range(steps):
        grad.normal_()
        gs = gather(grad)
        torch.cuda.synchronize()
        eng.mix(sgd=True)
        torch.cuda.synchronize()
        eng.check()
        xs = []
        for i in range(world):
            x, ms[i] = oracle.sgd_momentum(zs[i] * ws[i], gs[i], ms[i], lr, mu, wd, nest)
            xs.append(x)
        xs, ws = oracle.mix_columns(xs, ws, ogs, oms)
        oracle.rotate_all(ogs)
        zs = [x / w for x, w in zip(xs, ws)]
        torch.testing.assert_close(z.double(), zs[rank], rtol=1e-5, atol=1e-5)
    return True


@pytest.mark.parametrize('graph_name,ppi', [
    ('NPeerDynamicDirectedExponentialGraph', 1),
    ('DynamicDirectedExponentialGraph', 2),
    ('RingGraph', 1),
])
def test_sgp_mix_over_nvlink(graph_name, ppi):
    n = min(_ngpu(), 8)
    out = run_distributed(_sgp_worker, n, graph_name, ppi, 6, 64 * 4096,
                          backend='nccl', timeout=300)
    assert all(out)


def _skew_worker(rank, world, steps):
    """Ranks arrive with random delays; flags + acks must keep data consistent."""
    import time
    import torch.distributed as dist
    from stochastic_gradient_push_b200.parallel.symmetric import SymmetricWorld
    from stochastic_gradient_push_b200.ops.peer_mix import GossipEngine
    dev = torch.device('cuda', rank)
    graph = sgp.NPeerDynamicDirectedExponentialGraph(rank, world)
    z = torch.full((32 * 4096,), float(rank), device=dev)
    eng = GossipEngine(SymmetricWorld(dev), z, graph, sgp.UniformMixing(graph, dev), timeout_s=20.0)
    gen = torch.Generator().manual_seed(rank)
    total0 = torch.tensor([z.double().sum().item()], device=dev)
    dist.all_reduce(total0)
    for _ in range(steps):
        time.sleep(float(torch.rand(1, generator=gen)) * 0.01)
        eng.mix(sgd=False)
    torch.cuda.synchronize()
    eng.check()
    assert (z - z[0]).abs().max().item() == 0.0          # every element mixed identically
    total = torch.tensor([z.double().sum().item()], device=dev)
    dist.all_reduce(total)
    assert abs(total.item() - total0.item()) < 1e-3 * abs(total0.item()) + 1e-3
    return z[0].item()


def test_random_skew_conserves_mass():
    n = min(_ngpu(), 8)
    out = run_distributed(_skew_worker, n, 40, backend='nccl', timeout=300)
    mean = sum(range(n)) / n
    assert all(abs(v - mean) < 1e-3 for v in out)         # 40 steps >> log2(n): consensus


# --------------------------------------------------------------------------- #
# GossipDataParallel on the nvlink kernel transport vs the world simulation
# --------------------------------------------------------------------------- #
def _gdp_worker(rank, world, graph_name, ppi, steps, overlap, fused, nesterov, ppi_switch=None):
    import test_distributed_c10d as sim
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    from stochastic_gradient_push_b200.optim import FusedGossipSGD
    dev = torch.device('cuda', rank)
    graph = getattr(sgp, graph_name)(rank, world, peers_per_itr=ppi)
    net = sim._model(rank).to(dev)
    model = GossipDataParallel(net, graph=graph, overlap=overlap, rank=rank, world_size=world,
                               heartbeat_timeout=20)
    assert model.transport == 'nvlink'
    if fused:
        opt = FusedGossipSGD(model, lr=sim.LR, momentum=sim.MU, weight_decay=sim.WD,
                             nesterov=nesterov)
    else:
        opt = torch.optim.SGD(model.parameters(), lr=sim.LR, momentum=sim.MU,
                              weight_decay=sim.WD, nesterov=nesterov)
    model.train()
    for step in range(steps):
        if ppi_switch is not None and step == ppi_switch[0]:
            model.update_gossiper('peers_per_itr', ppi_switch[1])
        x, y = sim._batch(rank, step)
        loss = ((model(x.to(dev)) - y.to(dev)) ** 2).mean()
        loss.backward()
        opt.step()
        opt.zero_grad()
        if not overlap:
            model.transfer_params()
    model.sync_comms()
    model.unbias()
    torch.cuda.synchronize()
    model.engine.check()
    return sim._flat(model.module).cpu().tolist(), float(model.ps_weight)


@pytest.mark.parametrize('graph_name,ppi,overlap,fused,nesterov', [
    ('NPeerDynamicDirectedExponentialGraph', 1, False, True, True),
    ('NPeerDynamicDirectedExponentialGraph', 1, False, False, True),
    ('DynamicDirectedExponentialGraph', 2, False, True, False),
    ('NPeerDynamicDirectedExponentialGraph', 1, True, True, False),
    ('NPeerDynamicDirectedExponentialGraph', 1, True, False, False),
    ('RingGraph', 1, False, True, True),
])
def test_gossip_data_parallel_kernels_match_simulation(graph_name, ppi, overlap, fused, nesterov):
    import test_distributed_c10d as sim
    n = min(_ngpu(), 4)
    n = n if n % 2 == 0 else n - 1
    steps = 5
    out = run_distributed(_gdp_worker, n, graph_name, ppi, steps, overlap, fused, nesterov,
                          backend='nccl', timeout=300)
    want, ws = sim._simulate(n, graph_name, ppi, steps, overlap, nesterov)
    for r in range(n):
        got, w = out[r]
        torch.testing.assert_close(torch.tensor(got), want[r], rtol=1e-4, atol=1e-5)
        assert abs(w - ws[r]) < 1e-5


def _gdp_c10d_nccl_worker(rank, world, graph_name, ppi, steps, overlap):
    import test_distributed_c10d as sim
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    from stochastic_gradient_push_b200.optim import FusedGossipSGD
    dev = torch.device('cuda', rank)
    graph = getattr(sgp, graph_name)(rank, world, peers_per_itr=ppi)
    model = GossipDataParallel(sim._model(rank).to(dev), graph=graph, overlap=overlap, rank=rank,
                               world_size=world, transport='c10d')
    assert model.transport == 'c10d' and model._c10d.transport.batched       # NCCL: grouped launches
    opt = FusedGossipSGD(model, lr=sim.LR, momentum=sim.MU, weight_decay=sim.WD, nesterov=True)
    model.train()
    for step in range(steps):
        x, y = sim._batch(rank, step)
        ((model(x.to(dev)) - y.to(dev)) ** 2).mean().backward()
        opt.step()
        opt.zero_grad()
        if not overlap:
            model.transfer_params()
    model.sync_comms()
    model.unbias()
    torch.cuda.synchronize()
    return sim._flat(model.module).cpu().tolist(), float(model.ps_weight)


@pytest.mark.parametrize('graph_name,ppi,overlap', [
    ('NPeerDynamicDirectedExponentialGraph', 1, False),      # n = 2: in-peer == out-peer every step
    ('RingGraph', 1, False),
    ('NPeerDynamicDirectedExponentialGraph', 1, True),
])
def test_gossip_data_parallel_c10d_over_nccl_matches_simulation(graph_name, ppi, overlap):
    """the multi-host data plane (isend / irecv of GPU tensors over NCCL) on the GPUs of one host.
    Symmetric exchanges deadlock on NCCL unless a rank's receives and sends are launched as one
    group (gossiper.C10dTransport.exchange).  [written after the round's GPU budget was spent:
    not run on GPUs yet; its control flow is covered over gloo by tests/test_c10d_transport_modes.py]"""
    import test_distributed_c10d as sim
    n, steps = 2, 5
    out = run_distributed(_gdp_c10d_nccl_worker, n, graph_name, ppi, steps, overlap,
                          backend='nccl', timeout=300)
    want, ws = sim._simulate(n, graph_name, ppi, steps, overlap, True)
    for r in range(n):
        got, w = out[r]
        torch.testing.assert_close(torch.tensor(got), want[r], rtol=1e-4, atol=1e-5)
        assert abs(w - ws[r]) < 1e-5


def _trainer_worker(rank, world, algo, use_graph, steps):
    from stochastic_gradient_push_b200 import models
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    from stochastic_gradient_push_b200.optim import FusedGossipSGD
    from stochastic_gradient_push_b200.parallel.trainer import GossipTrainer
    dev = torch.device('cuda', rank)
    torch.manual_seed(7 + rank)
    net = models.TinyConvNet().to(dev).to(memory_format=torch.channels_last)
    graph = sgp.NPeerDynamicDirectedExponentialGraph(rank, world)
    model = GossipDataParallel(net, graph=graph, overlap=(algo == 'osgp'), rank=rank,
                               world_size=world, heartbeat_timeout=20)
    opt = FusedGossipSGD(model, lr=0.05, momentum=0.9, weight_decay=1e-4, nesterov=True)
    tr = GossipTrainer(model, opt, amp_dtype=None, use_cuda_graph=use_graph, warmup_iters=2)
    g = torch.Generator().manual_seed(100 + rank)
    losses = []
    for s in range(steps):
        x = torch.randn(8, 3, 32, 32, generator=g).pin_memory()
        y = torch.randint(0, 10, (8,), generator=g).pin_memory()
        slot = tr.step(x, y)
        torch.cuda.synchronize()
        losses.append(float(tr.loss_ring[slot]))
    tr.finish()
    return model.arena.flat.cpu().tolist(), losses, model.engine.device_step


@pytest.mark.parametrize('algo', ['sgp', 'osgp'])
def test_graphed_trainer_equals_eager(algo):
    n = 2
    steps = 7
    eager = run_distributed(_trainer_worker, n, algo, False, steps, backend='nccl', timeout=300)
    graphed = run_distributed(_trainer_worker, n, algo, True, steps, backend='nccl', timeout=300)
    for r in range(n):
        torch.testing.assert_close(torch.tensor(graphed[r][0]), torch.tensor(eager[r][0]),
                                   rtol=1e-4, atol=1e-5)
        assert graphed[r][2] == eager[r][2] == steps
        assert all(l == l for l in graphed[r][1])


# --------------------------------------------------------------------------- #
# AD-PSGD: device-side bilateral handshake
# --------------------------------------------------------------------------- #
def _adpsgd_consensus(rank, world, seconds):
    import time
    import torch.distributed as dist
    from stochastic_gradient_push_b200.parallel.ad_psgd import BilatGossipDataParallel
    dev = torch.device('cuda', rank)
    net = torch.nn.Linear(64, 64).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.fill_(float(rank))
    model = BilatGossipDataParallel(net, rank=rank, world_size=world,
                                    graph_class=sgp.DynamicBipartiteExponentialGraph,
                                    mixing_class=sgp.UniformMixing, lr=0.0, momentum=0.0,
                                    weight_decay=0.0, nesterov=False, verbose=False,
                                    heartbeat_timeout=20, max_rounds_per_update=None)
    assert model.transport == 'nvlink'
    total0 = torch.tensor([float(rank)], device=dev)
    dist.all_reduce(total0)
    model.enable_gossip()
    t0 = time.time()
    while time.time() - t0 < seconds and model.rounds_completed < 16:
        time.sleep(0.005)
    model.disable_gossip()
    time.sleep(0.3)
    dist.barrier()
    model.sync_comms()
    model._check()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    rounds = model.rounds_completed
    model.shutdown()
    return flat.mean().item(), (flat.max() - flat.min()).item(), rounds


def test_adpsgd_device_handshake_consensus():
    n = min(_ngpu(), 4)
    n = n if n % 2 == 0 else n - 1
    out = run_distributed(_adpsgd_consensus, n, 30.0, backend='nccl', timeout=300)
    assert min(o[2] for o in out) >= 2, out
    assert all(o[1] < 1e-6 for o in out)
    vals = [o[0] for o in out]
    assert max(vals) - min(vals) < 0.6 * (n - 1) + 1e-6, vals    # contracted
    assert min(vals) >= -1e-6 and max(vals) <= n - 1 + 1e-6


def _adpsgd_train(rank, world, steps):
    import time
    import torch.distributed as dist
    from stochastic_gradient_push_b200.parallel.ad_psgd import BilatGossipDataParallel
    dev = torch.device('cuda', rank)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.Tanh(), torch.nn.Linear(32, 1)).to(dev)
    model = BilatGossipDataParallel(net, rank=rank, world_size=world,
                                    graph_class=sgp.DynamicBipartiteExponentialGraph,
                                    mixing_class=sgp.UniformMixing, lr=0.05, momentum=0.9,
                                    weight_decay=0.0, nesterov=True, verbose=False,
                                    heartbeat_timeout=20)
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True)
    g = torch.Generator().manual_seed(rank)
    w_true = (torch.arange(6.) / 6).to(dev)
    model.train()
    model.enable_gossip()
    losses = []
    for s in range(steps):
        x = torch.randn(64, 6, generator=g).to(dev)
        y = (x @ w_true).unsqueeze(1)
        loss = ((model(x) - y) ** 2).mean()
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=False)
        losses.append(loss.item())
    model.eval()
    model.disable_gossip()
    applied, rounds = model.grads_applied, model.rounds_completed
    dist.barrier()
    model._check()
    model.shutdown()
    return losses[0], sum(losses[-5:]) / 5, applied, rounds


def test_adpsgd_trains_on_gpu():
    out = run_distributed(_adpsgd_train, 2, 80, backend='nccl', timeout=300)
    for first, last, applied, rounds in out:
        assert last < 0.5 * first, (first, last)
        assert applied >= 70


# --------------------------------------------------------------------------- #
# stand-alone gossipers on the nvlink transport (README usage of the reference)
# --------------------------------------------------------------------------- #
def _standalone_pushsum(rank, world, residual):
    from stochastic_gradient_push_b200.gossiper import PushSum
    dev = torch.device('cuda', rank)
    graph = sgp.NPeerDynamicDirectedExponentialGraph(rank, world)
    x = torch.full((5000,), 10.0 * rank, device=dev)
    g = PushSum(x, graph, rank=rank, world_size=world, transport='nvlink')
    w = torch.ones(1, device=dev)
    trace = []
    for _ in range(3):
        if residual:
            lo = 1.0 / (len(g.out_edges) + 1)
            x, w = x * lo, w * lo
            r, wr = g.mix(x.clone(), w, residual=True)
            x, w = x + r, w + wr
        else:
            x, w = g.mix(x.clone(), w, residual=False)
            x, w = x.clone(), w.clone()
        trace.append(round(x[0].item(), 3))
    assert abs(float(w) - 1.0) < 1e-5
    return trace


@pytest.mark.parametrize('residual', [False, True])
def test_standalone_pushsum_on_peer_memory(residual):
    n = 4 if _ngpu() >= 4 else 2
    out = run_distributed(_standalone_pushsum, n, residual, backend='nccl', timeout=300)
    if n == 4:
        want = {0: [15, 15, 15], 1: [5, 15, 15], 2: [15, 15, 15], 3: [25, 15, 15]}
    else:
        want = {0: [5, 5, 5], 1: [5, 5, 5]}
    for r in range(n):
        assert out[r] == want[r], out


@pytest.mark.parametrize('overlap', [False, True])
def test_device_schedule_swap_on_peers_per_itr_change(overlap):
    """update_gossiper('peers_per_itr') re-emits the device tables mid-training
    (phase_base / ack_from bookkeeping) -- must match the world simulation."""
    import test_distributed_c10d as sim
    n = 4 if _ngpu() >= 4 else 2
    steps, switch, name = 6, (3, 2), 'DynamicDirectedExponentialGraph'
    out = run_distributed(_gdp_worker, n, name, 1, steps, overlap, True, False, switch,
                          backend='nccl', timeout=300)
    want, ws = sim._simulate(n, name, 1, steps, overlap, False, switch)
    for r in range(n):
        got, w = out[r]
        torch.testing.assert_close(torch.tensor(got), want[r], rtol=1e-4, atol=1e-5)
        assert abs(w - ws[r]) < 1e-5


def _async_gpu_worker(rank, world, synch_freq, steps):
    import test_distributed_c10d as sim
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    dev = torch.device('cuda', rank)
    graph = sgp.NPeerDynamicDirectedExponentialGraph(rank, world)
    model = GossipDataParallel(sim._model(rank).to(dev), graph=graph, overlap=True,
                               synch_freq=synch_freq, rank=rank, world_size=world,
                               heartbeat_timeout=30)
    x = torch.zeros(2, 6, device=dev)
    for _ in range(steps):
        model(x)
    model.sync_comms()
    model.unbias()
    torch.cuda.synchronize()
    model.engine.check()
    return sim._flat(model.module).cpu().tolist(), float(model.ps_weight)


def test_bounded_staleness_on_kernels_conserves_mass():
    import test_distributed_c10d as sim
    n = 4 if _ngpu() >= 4 else 2
    out = run_distributed(_async_gpu_worker, n, 2, 12, backend='nccl', timeout=300)
    x0 = torch.stack([sim._flat(sim._model(r)) for r in range(n)])
    mass = sum(torch.tensor(z) * w for z, w in out)
    torch.testing.assert_close(mass, x0.sum(0), rtol=1e-4, atol=1e-4)
    assert abs(sum(w for _, w in out) - n) < 1e-4


def _barrier_worker(rank, world):
    import time
    from stochastic_gradient_push_b200.parallel.symmetric import SymmetricWorld
    from stochastic_gradient_push_b200.ops.peer_mix import GossipEngine
    dev = torch.device('cuda', rank)
    graph = sgp.RingGraph(rank, world)
    eng = GossipEngine(SymmetricWorld(dev), torch.zeros(4096, device=dev), graph,
                       sgp.UniformMixing(graph, dev), timeout_s=20.0, name='bar')
    stamps = []
    for i in range(3):
        time.sleep(0.05 * rank)                  # ranks arrive staggered
        eng.barrier()
        torch.cuda.synchronize()
        stamps.append(time.time())
    eng.check()
    return stamps


def test_device_barrier_kernel():
    n = min(_ngpu(), 4)
    out = run_distributed(_barrier_worker, n, backend='nccl', timeout=200)
    for i in range(3):                            # nobody leaves a barrier before the last arrives
        leave = [out[r][i] for r in range(n)]
        assert max(leave) - min(leave) < 0.04, leave


# --------------------------------------------------------------------------- #
# NVLS: VMM symmetric memory + multicast, multimem.* kernels
# --------------------------------------------------------------------------- #
def _nvls_supported():
    from stochastic_gradient_push_b200.parallel.symmetric import VmmSymmetricWorld
    return _ngpu() >= 2 and VmmSymmetricWorld.supported(0)


def _vmm_world_worker(rank, world):
    from stochastic_gradient_push_b200.parallel.symmetric import VmmSymmetricWorld
    dev = torch.device('cuda', rank)
    sw = VmmSymmetricWorld(dev)
    buf = sw.alloc('t', 1 << 20)
    mine = buf.local.view(torch.float32)
    mine.fill_(float(rank + 1))
    torch.cuda.synchronize()
    sw.barrier()
    # unicast P2P view of every peer + the multicast view exist and alias the right memory
    seen = [float(buf.peers[r].view(torch.float32)[5].item()) for r in range(world)]
    assert buf.mc is not None and buf.mc.numel() == buf.local.numel()
    sw.barrier()
    return seen


def test_vmm_symmetric_world_maps_peers_and_multicast():
    if not _nvls_supported():
        pytest.skip('no NVSwitch multicast')
    n = min(_ngpu(), 4)
    out = run_distributed(_vmm_world_worker, n, backend='nccl', timeout=200)
    for seen in out:
        assert seen == [float(r + 1) for r in range(n)]


def _nvls_ar_worker(rank, world, grad_dtype, steps):
    import copy
    import torch.distributed as dist
    from stochastic_gradient_push_b200.parallel.allreduce import AllReduceDataParallel
    dev = torch.device('cuda', rank)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.ReLU(), torch.nn.Linear(256, 40)).to(dev)
    ref = copy.deepcopy(net)
    ar = AllReduceDataParallel(net, rank=rank, world_size=world, transport='nvls')
    assert ar.transport == 'nvls'
    n = ar.arena.total
    g16 = None
    if grad_dtype == 'bf16':
        # bf16 gradient buffers (what a bf16 compute twin accumulates into): a second symmetric
        # allocation whose multicast view is reduced with multimem.ld_reduce.add.acc::f32.v4.bf16x2
        g16 = ar.world.alloc('ar.g16', n * 2)
    opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4, nesterov=True)
    ar.set_hyper(0.1, 0.9, 1e-4, True)
    g = torch.Generator(device='cuda').manual_seed(100 + rank)
    for _ in range(steps):
        x = torch.randn(32, 64, device=dev, generator=g)
        ar(x).square().mean().backward()
        if g16 is not None:
            local16 = g16.local.view(torch.bfloat16)[:n]
            local16.copy_(ar.grad_flat)                  # this rank's gradient, rounded to bf16
            ar.grad_flat.zero_()
            gflat = local16.float().clone()
        else:
            gflat = ar.grad_flat.clone()
        # reference: average of the ranks' gradients (as the kernel will see them), plain SGD
        dist.all_reduce(gflat)
        gflat /= world
        if g16 is not None:
            torch.cuda.synchronize()
            dist.barrier()
            ar.C.nvls_allreduce(ar.arena.flat, ar._z_mc, g16.mc.view(torch.bfloat16)[:n], ar.momentum,
                                ar.pad.table, ar.state, ar.hyper, ar.world.rank, ar.world.world, 30.0, 1.0,
                                True, ar.grid)
        else:
            ar.allreduce_step()
        opt.zero_grad()
        for p, v in zip(ref.parameters(), ar.arena.views_of(gflat)):
            p.grad = v.clone()
        opt.step()
        torch.cuda.synchronize()
        ar.check()
        cleared = g16.local.view(torch.bfloat16)[:n] if g16 is not None else ar.grad_flat
        assert float(cleared.float().abs().max()) == 0.0            # cleared on every rank by multimem.st
    tol = dict(rtol=1e-4, atol=1e-5) if grad_dtype == 'fp32' else dict(rtol=1e-3, atol=1e-4)
    for p, q in zip(net.parameters(), ref.parameters()):
        torch.testing.assert_close(p, q, **tol)
    # replicas are bit-identical (every rank received the same multicast values)
    flat = ar.arena.flat.clone()
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    assert all(torch.equal(o, others[0]) for o in others)
    return True


@pytest.mark.parametrize('grad_dtype', ['fp32', 'bf16'])
def test_nvls_allreduce_sgd_matches_torch_sgd(grad_dtype):
    if not _nvls_supported():
        pytest.skip('no NVSwitch multicast')
    n = min(_ngpu(), 8)
    assert all(run_distributed(_nvls_ar_worker, n, grad_dtype, 5, backend='nccl', timeout=300))


def _hier_nvls_worker(rank, world, steps):
    """world = ONE node of `world` local ranks (nprocs_per_node = world): the gossip world has a
    single (peer-less) agent, so the result must equal single-process SGD on the mean gradient."""
    import copy
    import torch.distributed as dist
    from stochastic_gradient_push_b200.parallel.distributed import GossipDataParallel
    dev = torch.device('cuda', rank)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(32, 128), torch.nn.Tanh(), torch.nn.Linear(128, 8)).to(dev)
    ref = copy.deepcopy(net)
    if rank != 0:
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)                  # only the node master's parameters count (broadcast)
    model = GossipDataParallel(net, rank=rank, world_size=world, nprocs_per_node=world,
                               graph=sgp.NPeerDynamicDirectedExponentialGraph(0, 1), heartbeat_timeout=30)
    assert model._hier is not None, 'hierarchical mode did not get its NVLS group'
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
    gens = [torch.Generator(device='cuda').manual_seed(7 + r) for r in range(world)]
    for _ in range(steps):
        xs = [torch.randn(16, 32, device=dev, generator=g) for g in gens]
        model(xs[rank]).square().mean().backward()
        opt.step()
        opt.zero_grad(set_to_none=False)
        model.transfer_params()
        model._query_gossip_queue()
        torch.stack([ref(x).square().mean() for x in xs]).mean().backward()
        ropt.step()
        ropt.zero_grad()
    # one more forward: parameters of the master have been multicast to everyone
    model(xs[rank])
    torch.cuda.synchronize()
    for p, q in zip(net.parameters(), ref.parameters()):
        torch.testing.assert_close(p, q, rtol=1e-4, atol=1e-5)
    return True


def test_hierarchical_mode_runs_on_nvls_kernels():
    if not _nvls_supported():
        pytest.skip('no NVSwitch multicast')
    assert all(run_distributed(_hier_nvls_worker, 2, 4, backend='nccl', timeout=300))


def test_adpsgd_single_process_multi_gpu_replicas():
    """reference launch mode for AD-PSGD (one process, several GPUs; gossip/ad_psgd.py:57-69,
    148-191, 378-404): the replicas' gradients are summed into the master's by the P2P kernel before
    they are handed to the gossip side, whose fused SGD must then equal plain SGD on the whole batch."""
    import copy
    from stochastic_gradient_push_b200.parallel.ad_psgd import BilatGossipDataParallel
    if _ngpu() < 2:
        pytest.skip('needs 2 GPUs')
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4)).to(dev)
    ref = copy.deepcopy(net)
    model = BilatGossipDataParallel(net, device_ids=[0, 1], rank=0, world_size=1, lr=0.05, momentum=0.9,
                                    weight_decay=0.0, nesterov=True, verbose=False, heartbeat_timeout=20)
    assert model._replicas is not None and len(model._module_copies) == 2
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, nesterov=True)
    model.train()
    g = torch.Generator(device='cuda').manual_seed(1)
    try:
        for _ in range(5):
            x = torch.randn(24, 16, device=dev, generator=g)
            y = torch.randn(24, 4, device=dev, generator=g)
            out = model(x)
            assert out.shape == (24, 4) and out.device == dev
            (((out - y) ** 2).sum() / 24).backward()
            opt.step()
            opt.zero_grad(set_to_none=False)
            (((ref(x) - y) ** 2).sum() / 24).backward()
            ropt.step()
            ropt.zero_grad()
        model.eval()                       # pulls the gossip copy into the training copy
        torch.cuda.synchronize()
        model._check()
        assert model.grads_applied == 5
        for p, q in zip(net.parameters(), ref.parameters()):
            torch.testing.assert_close(p, q, rtol=1e-4, atol=1e-5)
    finally:
        model.shutdown()
