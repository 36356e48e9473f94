"""Spawn N local processes with a gloo world on 127.0.0.1 and run `fn(rank, world, *args)`
in each; exceptions (with traceback) are re-raised in the parent."""
import os
import sys
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _entry(rank, world, port, fn, args, q, backend):
    try:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        os.environ['RANK'] = str(rank)
        os.environ['WORLD_SIZE'] = str(world)
        os.environ['LOCAL_RANK'] = str(rank)
        torch.set_num_threads(1)
        if backend == 'nccl':
            torch.cuda.set_device(rank)
        dist.init_process_group(backend, rank=rank, world_size=world)
        out = fn(rank, world, *args)
        q.put((rank, None, out))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:                                   # noqa
        q.put((rank, traceback.format_exc(), None))


def run_distributed(fn, world, *args, backend='gloo', timeout=120):
    from conftest import free_port
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, args, q, backend))
             for r in range(world)]
    for p in procs:
        p.start()
    results, err = {}, None
    try:
        for _ in range(world):
            rank, tb, out = q.get(timeout=timeout)
            if tb is not None:
                err = 'rank %d failed:\n%s' % (rank, tb)
                break
            results[rank] = out
    finally:
        for p in procs:
            p.join(timeout=5 if err else timeout)
            if p.is_alive():
                p.terminate()
    if err:
        raise AssertionError(err)
    return [results[r] for r in range(world)]
