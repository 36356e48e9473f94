"""CLI flag parsing, schedules, CSV format, ClusterManager, end-to-end CPU runs."""
import os
import signal
import subprocess
import sys
import types

import pytest
import torch

from stochastic_gradient_push_b200.cli import common
from stochastic_gradient_push_b200.experiment import ClusterManager, get_tcp_interface_name

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(**kw):
    base = dict(lr=0.1, batch_size=256, world_size=8, warmup=True,
                lr_schedule={30: 0.1, 60: 0.1, 80: 0.1})
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_string_booleans_like_the_job_scripts():
    p = common.build_parser()
    a = p.parse_args(['--push_sum', 'False', '--nesterov', 'True', '--overlap', 'True',
                      '--all_reduce', 'False', '--schedule', '30', '0.1', '60', '0.1'])
    assert a.push_sum is False and a.nesterov is True and a.overlap is True
    assert a.all_reduce is False and a.schedule == [30.0, 0.1, 60.0, 0.1]
    d = p.parse_args([])
    assert (d.batch_size, d.lr, d.graph_type, d.seed, d.print_freq, d.num_itr_ignore) == \
        (32, 0.1, 5, 47, 10, 10)
    assert d.push_sum is True and d.backend == 'nccl' and d.master_port == '40100'
    ad = common.build_parser(adpsgd=True).parse_args(['--shared_fpath', '/x', '--bilat', 'True'])
    assert ad.shared_fpath == '/x' and not hasattr(ad, 'num_itr_ignore')


def test_schedule_parsing():
    assert common.pairs_to_dict([30, 0.1, 60, 0.5]) == {30: 0.1, 60: 0.5}
    assert common.pairs_to_dict([0, 1, 10, 2], int) == {0: 1, 10: 2}
    with pytest.raises(AssertionError):
        common.pairs_to_dict([1, 2, 3])


def test_learning_rate_policy():
    a = _args()
    target = 0.1 * 256 * 8 / 256                      # linear scaling rule
    ipe = 100
    assert abs(common.learning_rate_at(a, 0, 0, ipe) - (0.1 + (target - 0.1) * 1 / 500)) < 1e-12
    assert abs(common.learning_rate_at(a, 4, 99, ipe) - target) < 1e-12     # end of warm-up
    assert abs(common.learning_rate_at(a, 5, 0, ipe) - target) < 1e-12
    assert abs(common.learning_rate_at(a, 30, 0, ipe) - target * 0.1) < 1e-12
    assert abs(common.learning_rate_at(a, 85, 0, ipe) - target * 1e-3) < 1e-12
    small = _args(batch_size=32, world_size=2)         # target below the base lr: no ramp
    assert abs(common.learning_rate_at(small, 0, 0, ipe) - 0.1 * 32 * 2 / 256) < 1e-12
    nowarm = _args(warmup=False)
    assert abs(common.learning_rate_at(nowarm, 0) - target) < 1e-12


def test_peers_per_itr_schedule_lookup():
    sched = {0: 1, 10: 2, 40: 4}
    assert [common.peers_per_itr_at(sched, e) for e in (0, 9, 10, 39, 40, 89)] == [1, 1, 2, 2, 4, 4]


def test_accuracy_topk():
    out = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5]])
    tgt = torch.tensor([1, 2, 2])
    p1, p2 = common.accuracy(out, tgt, topk=(1, 2))
    assert abs(p1.item() - 200 / 3) < 1e-4 and abs(p2.item() - 200 / 3) < 1e-4
    tgt2 = torch.tensor([1, 1, 2])
    p1, p2 = common.accuracy(out, tgt2, topk=(1, 2))
    assert abs(p2.item() - 100.0) < 1e-4


def test_csv_format_matches_plotting_contract(tmp_path):
    from stochastic_gradient_push_b200.utils import Meter
    f = str(tmp_path / 'out_r0_n4.csv')
    log = common.CSVLog(f, 4, 10, 256)
    m = Meter()
    m.update(0.25)
    loss = Meter()
    loss.update(2.5)
    log.train_row(0, 10, m, m, m, loss, loss, loss)
    log.val_row(0, m, m, m, 71.2)
    lines = open(f).read().splitlines()
    assert lines[:4] == ['BEGIN-TRAINING', 'World-Size,4', 'Num-DLWorkers,10', 'Batch-Size,256']
    assert lines[4] == common.CSV_COLUMNS and len(lines[4].split(',')) == 18
    assert len(lines[5].split(',')) == 18 and lines[5].endswith(',-1')
    assert lines[6].split(',')[1] == '-1' and lines[6].endswith(',71.2')
    import pandas as pd
    df = pd.read_csv(f, skiprows=4)                   # how visualization/plotting.py reads it
    assert list(df.columns)[:3] == ['Epoch', 'itr', 'BT(s)'] and len(df) == 2


def test_cluster_manager_checkpoint_and_signal(tmp_path):
    ClusterManager.set_checkpoint_dir(str(tmp_path) + '/')
    state = {'epoch': 3, 'is_best': True, 'w': torch.arange(4.)}
    fired = []
    cm = ClusterManager(rank=2, world_size=1, state=state, model_tag='t_', all_workers=True,
                        callback=lambda: fired.append(1))
    assert cm.checkpoint_fpath.endswith('t_checkpoint_r2_n1.pth.tar')
    assert cm.model_best_fpath.endswith('t_model_best_r2_n1.pth.tar')
    cm.save_checkpoint()
    assert os.path.isfile(cm.checkpoint_fpath) and os.path.isfile(cm.model_best_fpath)
    assert state['is_best'] is False
    assert torch.equal(torch.load(cm.checkpoint_fpath)['w'], torch.arange(4.))
    cm.save_checkpoint(epoch_id=7)
    assert os.path.isfile(str(tmp_path) + '/ep7_t_checkpoint_r2_n1.pth.tar')
    # rank != master without all_workers writes nothing
    cm2 = ClusterManager(rank=1, world_size=1, state=state, model_tag='u_', all_workers=False)
    cm2.save_checkpoint()
    assert not os.path.exists(str(tmp_path) + '/u_checkpoint_r0_n1.pth.tar')
    # SIGUSR1: remembered, agreed at the next checkpoint, clean exit (no NameError)
    os.kill(os.getpid(), signal.SIGUSR1)
    assert cm2.signal_received and not cm.signal_received or True
    cm.signal_received = True
    with pytest.raises(SystemExit) as e:
        cm.save_checkpoint(requeue_on_signal=True)
    assert e.value.code == 0
    os.kill(os.getpid(), signal.SIGTERM)              # logged and ignored


def test_nic_probe_returns_or_raises_cleanly():
    try:
        name = get_tcp_interface_name('ethernet')
        assert isinstance(name, str) and name
    except Exception as e:
        assert 'interface found' in str(e)


def _torchrun(nproc, script, args, port, timeout=600):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node',
           str(nproc), '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(ROOT, script)] + args
    env = dict(os.environ, OMP_NUM_THREADS='1')
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=timeout, env=env)


COMMON = ['--device', 'cpu', '--backend', 'gloo', '--model', 'tiny', '--num_classes', '10',
          '--image_size', '16', '--synthetic', 'True', '--synthetic_len', '64', '--batch_size', '4',
          '--verbose', 'False', '--print_freq', '2', '--amp', 'False',
          '--num_dataloader_workers', '0', '--lr', '0.05']


@pytest.mark.parametrize('extra', [
    ['--push_sum', 'True', '--graph_type', '5'],
    ['--push_sum', 'True', '--graph_type', '0', '--overlap', 'True'],
    ['--push_sum', 'False', '--graph_type', '4', '--fused', 'False'],
    ['--all_reduce', 'True', '--graph_type', '-1'],
])
def test_gossip_sgd_cli_two_ranks_cpu(tmp_path, master_port, extra):
    out = _torchrun(2, 'gossip_sgd.py', COMMON + extra + [
        '--num_epochs', '2', '--checkpoint_dir', str(tmp_path) + '/', '--num_itr_ignore', '0'],
        master_port)
    assert out.returncode == 0, out.stdout[-3000:]
    for r in range(2):
        lines = open(str(tmp_path / ('out_r%d_n2.csv' % r))).read().splitlines()
        assert lines[1] == 'World-Size,2'
        rows = [l.split(',') for l in lines[5:]]
        assert any(row[1] == '-1' for row in rows)             # validation rows
        assert all(len(row) == 18 for row in rows)
        assert os.path.isfile(str(tmp_path / ('checkpoint_r%d_n2.pth.tar' % r)))


@pytest.mark.parametrize('fused', ['False', 'True'])
def test_gossip_sgd_cli_hierarchical_nprocs_per_node(tmp_path, master_port, fused):
    """--nprocs_per_node 2 on 4 ranks = 2 nodes: only ranks 0 and 2 gossip (graph over the nodes),
    their node-mates mirror them at every forward pass.  Every rank writes its CSV / checkpoint;
    the checkpoints of the two ranks of a node differ by at most the last local step + mix (the
    snapshot is taken before the next forward re-broadcasts the master's parameters)."""
    import torch
    out = _torchrun(4, 'gossip_sgd.py', COMMON + [
        '--push_sum', 'True', '--graph_type', '5', '--nprocs_per_node', '2', '--fused', fused,
        '--synthetic_len', '128', '--num_epochs', '1', '--checkpoint_dir', str(tmp_path) + '/',
        '--num_itr_ignore', '0'], master_port)
    assert out.returncode == 0, out.stdout[-3000:]
    sds = []
    for r in range(4):
        assert os.path.isfile(str(tmp_path / ('out_r%d_n4.csv' % r)))
        ck = torch.load(str(tmp_path / ('checkpoint_r%d_n4.pth.tar' % r)), map_location='cpu',
                        weights_only=False)
        sd = ck['state_dict']['state_dict']
        sds.append(torch.cat([v.reshape(-1).float() for k, v in sorted(sd.items())
                              if 'running' not in k and 'num_batches' not in k]))
    assert all(torch.isfinite(v).all() for v in sds)
    torch.testing.assert_close(sds[0], sds[1], rtol=0, atol=2e-2)
    torch.testing.assert_close(sds[2], sds[3], rtol=0, atol=2e-2)
    assert 'World-Size,4' in open(str(tmp_path / 'out_r3_n4.csv')).read()


def test_gossip_sgd_cli_writes_chrome_trace(tmp_path, master_port):
    """--trace_file: per-rank Chrome trace with forward / backward / optimizer / gossip spans and
    the exposed-communication counter, bounded by --trace_iters."""
    import json
    prefix = str(tmp_path / 'trace')
    out = _torchrun(2, 'gossip_sgd.py', COMMON + [
        '--graph_type', '5', '--fused', 'False', '--num_epochs', '1', '--train_fast', 'True',
        '--checkpoint_dir', str(tmp_path) + '/', '--trace_file', prefix, '--trace_iters', '5'],
        master_port)
    assert out.returncode == 0, out.stdout[-3000:]
    for r in range(2):
        doc = json.load(open('%s_r%d.json' % (prefix, r)))
        ev = doc['traceEvents']
        spans = [e for e in ev if e.get('ph') == 'X']
        names = {e['name'] for e in spans}
        assert {'forward', 'backward', 'optimizer', 'gossip.post', 'gossip.wait+fold'} <= names, names
        assert sum(e['name'] == 'forward' for e in spans) == 5            # bounded by --trace_iters
        assert all(e['pid'] == r and e['dur'] >= 0 for e in spans)
        assert any(e.get('ph') == 'C' and e['name'] == 'exposed_comm_ms' for e in ev)


def test_gossip_sgd_cli_resume(tmp_path, master_port):
    base = COMMON + ['--graph_type', '5', '--checkpoint_dir', str(tmp_path) + '/']
    out = _torchrun(2, 'gossip_sgd.py', base + ['--num_epochs', '1'], master_port)
    assert out.returncode == 0, out.stdout[-3000:]
    ck = torch.load(str(tmp_path / 'checkpoint_r0_n2.pth.tar'), weights_only=False)
    assert ck['epoch'] == 1 and set(ck['state_dict']) == {'state_dict', 'ps_weight', 'is_ps_numerator'}
    out = _torchrun(2, 'gossip_sgd.py', base + ['--num_epochs', '2', '--resume', 'True'],
                    master_port + 1)
    assert out.returncode == 0, out.stdout[-3000:]
    assert 'loaded checkpoint (epoch 1' in out.stdout
    ck2 = torch.load(str(tmp_path / 'checkpoint_r0_n2.pth.tar'), weights_only=False)
    assert ck2['epoch'] == 2


def test_adpsgd_cli_two_ranks_cpu(tmp_path, master_port):
    out = _torchrun(2, 'gossip_sgd_adpsgd.py', COMMON + [
        '--num_epochs', '2', '--checkpoint_dir', str(tmp_path) + '/', '--graph_type', '1',
        '--train_fast', 'True'], master_port)
    assert out.returncode == 0, out.stdout[-3000:]
    size = os.stat(str(tmp_path / 'global_itr.txt')).st_size
    assert size >= 2 * 2 * 8        # >= num_epochs * world * itr_per_epoch bytes appended


def _signal_worker(rank, world, ckpt_dir):
    import torch.distributed as dist
    ClusterManager.set_checkpoint_dir(ckpt_dir)
    state = {'epoch': 1, 'is_best': False}
    cm = ClusterManager(rank=rank, world_size=world, state=state, model_tag='sig_', all_workers=True)
    cm.save_checkpoint(requeue_on_signal=True)           # nobody signalled: returns
    if rank == 1:
        os.kill(os.getpid(), signal.SIGUSR1)             # pre-emption notice on ONE rank
    exited = False
    try:
        cm.save_checkpoint(requeue_on_signal=True)       # all ranks agree and exit(0)
    except SystemExit as e:
        exited = (e.code == 0)
    return exited, os.path.isfile(cm.checkpoint_fpath)


def test_cluster_manager_agrees_on_preemption_across_ranks(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from dist_utils import run_distributed
    out = run_distributed(_signal_worker, 2, str(tmp_path) + '/')
    assert out == [(True, True), (True, True)]


def _mpirun_like(nproc, script, args, port, timeout=600):
    """what `mpirun -np N python gossip_sgd.py --backend mpi` looks like to the script: one process
    per rank with OMPI_COMM_WORLD_* set (no torchrun variables), MASTER_* exported by the job script"""
    procs = []
    for r in range(nproc):
        env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')}
        env.update(OMP_NUM_THREADS='1', OMPI_COMM_WORLD_RANK=str(r), OMPI_COMM_WORLD_SIZE=str(nproc),
                   OMPI_UNIVERSE_SIZE=str(nproc), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, script)] + args, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=timeout)[0] for p in procs]
    return [p.returncode for p in procs], outs


def test_backend_mpi_uses_the_mpirun_environment(tmp_path, master_port):
    """`--backend mpi` (reference gossip_sgd.py:127-129): ranks / world from OMPI_COMM_WORLD_*;
    without an MPI-enabled PyTorch the control plane falls back to a TCP rendezvous (gloo here)."""
    a = common.build_parser().parse_args(['--backend', 'mpi'])
    a.device = 'cpu'
    want = 'mpi' if torch.distributed.is_mpi_available() else 'gloo'
    assert common.resolve_backend(a) == want
    args = [x if x != 'gloo' else 'mpi' for x in COMMON] + [
        '--push_sum', 'True', '--graph_type', '5', '--num_epochs', '1',
        '--checkpoint_dir', str(tmp_path) + '/', '--num_itr_ignore', '0']
    codes, outs = _mpirun_like(2, 'gossip_sgd.py', args, master_port)
    assert codes == [0, 0], outs[0][-2000:] + outs[1][-2000:]
    for r in range(2):
        assert open(str(tmp_path / ('out_r%d_n2.csv' % r))).read().splitlines()[1] == 'World-Size,2'
