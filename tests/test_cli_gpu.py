"""The user-facing entry points on real GPUs (one rank per GPU, nvlink transport)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script, args, port, timeout=600):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node',
           str(nproc), '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(ROOT, script)] + args
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=timeout)


COMMON = ['--model', 'tiny', '--num_classes', '10', '--image_size', '32', '--synthetic', 'True',
          '--synthetic_len', '256', '--batch_size', '16', '--verbose', 'False', '--print_freq', '2',
          '--num_dataloader_workers', '0', '--lr', '0.05', '--num_epochs', '2']


@pytest.mark.parametrize('extra', [
    ['--push_sum', 'True', '--graph_type', '5'],
    ['--push_sum', 'True', '--graph_type', '5', '--cuda_graph', 'True'],
    ['--push_sum', 'True', '--graph_type', '0', '--overlap', 'True'],
    ['--push_sum', 'False', '--graph_type', '4', '--fused', 'False'],
    ['--all_reduce', 'True', '--graph_type', '-1'],
])
def test_gossip_sgd_cli_on_gpus(tmp_path, master_port, extra):
    out = _torchrun(2, 'gossip_sgd.py', COMMON + extra + [
        '--checkpoint_dir', str(tmp_path) + '/', '--num_itr_ignore', '0'], master_port)
    assert out.returncode == 0, out.stdout[-3000:]
    if '--all_reduce' not in extra:
        assert 'transport: nvlink' in out.stdout
    for r in range(2):
        rows = open(str(tmp_path / ('out_r%d_n2.csv' % r))).read().splitlines()[5:]
        assert len(rows) >= 6 and all(len(x.split(',')) == 18 for x in rows)
        losses = [float(x.split(',')[12]) for x in rows if x.split(',')[1] != '-1']
        assert all(l == l for l in losses)          # no NaNs
        assert os.path.isfile(str(tmp_path / ('checkpoint_r%d_n2.pth.tar' % r)))


def test_adpsgd_cli_on_gpus(tmp_path, master_port):
    out = _torchrun(2, 'gossip_sgd_adpsgd.py', COMMON + [
        '--checkpoint_dir', str(tmp_path) + '/', '--graph_type', '1', '--train_fast', 'True'],
        master_port)
    assert out.returncode == 0, out.stdout[-3000:]
    assert os.stat(str(tmp_path / 'global_itr.txt')).st_size >= 2 * 2 * 8
