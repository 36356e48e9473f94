"""
ClusterManager: checkpoint files + SLURM pre-emption protocol.

API parity with ``experiment_utils/cluster_manager.py:24-141``: class attribute
``CHECKPOINT_DIR`` / ``set_checkpoint_dir``, constructor ``(rank, world_size,
state, model_tag='', callback=None, all_workers=False)``, ``save_checkpoint(
epoch_id=None, requeue_on_signal=True)``, ``checkpoint_fpath`` /
``model_best_fpath``, SIGUSR1 / SIGTERM handlers.  File names are identical
(``{dir}{tag}checkpoint_r{rank}_n{ws}.pth.tar``, ``ep{N}_`` prefix for
per-epoch files, ``model_best_r..`` copies) so checkpoints interchange.

Fixes relative to the reference: the requeue path exits cleanly (the
reference calls ``sys.exit`` without importing ``sys``), the
``callback`` argument is honoured, the "was anybody signalled?" all-reduce
runs on a CPU (gloo) tensor when the backend allows it -- a 1-element control
message has no business on the GPU stream --, and writes are atomic
(tmp file + rename) so a pre-emption mid-save cannot corrupt the checkpoint.
"""

from __future__ import annotations

import os
import shutil
import signal
import sys

import torch
import torch.distributed as dist

from ..utils.helpers import make_logger


class ClusterManager(object):

    MASTER_RANK = 0
    CHECKPOINT_DIR = None

    @staticmethod
    def set_checkpoint_dir(checkpoint_dir):
        ClusterManager.CHECKPOINT_DIR = checkpoint_dir

    def __init__(self, rank, world_size, state, model_tag='', callback=None,
                 all_workers=False):
        assert ClusterManager.CHECKPOINT_DIR is not None
        self.rank = rank
        self.world_size = world_size
        self.state = state
        self.all_workers = all_workers
        self.main_pid = os.getpid()
        self.callback = callback
        self.signal_received = False
        self.logger = make_logger(rank)
        self.model_tag = model_tag

        model_rank = rank if all_workers else ClusterManager.MASTER_RANK
        suffix = '_r{}_n{}.pth.tar'.format(model_rank, world_size)
        self.checkpoint_fname = 'checkpoint' + suffix
        self.model_best_fname = 'model_best' + suffix
        self.checkpoint_fpath = ClusterManager.CHECKPOINT_DIR + model_tag + self.checkpoint_fname
        self.model_best_fpath = ClusterManager.CHECKPOINT_DIR + model_tag + self.model_best_fname

        self.signal_handlers_installed = False
        self.install_signal_handlers()

        self.process_group = None
        self._signal_device = 'cpu'
        if self.world_size > 1:
            assert dist.is_initialized()
            self.process_group = dist.new_group(list(range(self.world_size)))
            if dist.get_backend(self.process_group) == 'nccl':
                self._signal_device = 'cuda'
        self.signal_tensor = torch.zeros(1, device=self._signal_device)

    # ------------------------------------------------------------------ #
    def _epoch_fpath(self, epoch_id):
        if epoch_id is None:
            return self.checkpoint_fpath
        return (ClusterManager.CHECKPOINT_DIR + 'ep' + str(epoch_id) + '_'
                + self.model_tag + self.checkpoint_fname)

    def save_checkpoint(self, epoch_id=None, requeue_on_signal=True):
        if self.signal_received:
            self.signal_tensor[0] = 1
        if requeue_on_signal and self.world_size > 1:
            dist.all_reduce(self.signal_tensor, group=self.process_group)

        self.logger.info('Saving checkpoint')
        if self.all_workers or self.rank == ClusterManager.MASTER_RANK:
            fpath = self._epoch_fpath(epoch_id)
            tmp = fpath + '.tmp.%d' % os.getpid()
            torch.save(self.state, tmp)
            os.replace(tmp, fpath)
            if self.state.get('is_best', False):
                shutil.copyfile(fpath, self.model_best_fpath)
                self.state['is_best'] = False

        if requeue_on_signal and float(self.signal_tensor[0]) > 0:
            self.logger.info('At least 1 process received SIGUSR1. Terminating')
            if self.rank == 0 and os.getpid() == self.main_pid:
                job = os.environ.get('SLURM_JOB_ID')
                if job is not None:
                    command = 'scontrol requeue ' + job
                    self.logger.info('Relaunching: ' + command)
                    if os.system(command):
                        raise RuntimeError('requeue failed')
                    self.logger.info('New job submitted to the queue')
            self.logger.info('Terminating')
            sys.exit(0)

    # ------------------------------------------------------------------ #
    def install_signal_handlers(self):
        try:
            signal.signal(signal.SIGUSR1, self.SIGUSR1Handler)
            signal.signal(signal.SIGTERM, self.SIGTERMHandler)
            self.signal_handlers_installed = True
            self.logger.info('Signal handlers installed')
        except ValueError:        # not the main thread
            self.logger.warning('signal handlers NOT installed (not in main thread)')

    def SIGTERMHandler(self, signum, frame):
        """SIGTERM precedes SIGUSR1 under SLURM pre-emption: log and carry on."""
        self.logger.info('Received SIGTERM')

    def SIGUSR1Handler(self, signum, frame):
        """Remember the signal; the next save_checkpoint() agrees on it globally."""
        self.logger.info('Received SIGUSR1')
        if self.callback is not None:
            self.callback()
        self.signal_received = True
