"""_C.FdChannel: file descriptors passed between processes over abstract unix-domain sockets
(SCM_RIGHTS) -- the rendezvous primitive of the VMM / multicast symmetric memory
(csrc/vmm_symm.cpp, parallel/symmetric.py::VmmSymmetricWorld).  Pure host code: runs without a GPU."""
import multiprocessing as mp
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, job, q):
    sys.path.insert(0, ROOT)
    import time
    from stochastic_gradient_push_b200.ops import native
    C = native.load()
    ch = C.FdChannel(job, rank, world)
    time.sleep(0.2 * rank)                       # staggered start: senders retry until the peer listens
    r, w = os.pipe()
    os.write(w, ('rank %d says hi' % rank).encode())
    for peer in range(world):                    # everybody sends first ...
        if peer != rank:
            ch.send(peer, 40 + rank, r)
    got = {}
    for _ in range(world - 1):                   # ... then everybody receives: must not deadlock
        src, tag, fd = ch.recv()
        assert tag == 40 + src
        got[src] = fd
    # the received descriptors are dups of the SAME pipe read end: each message is read exactly once
    q.put((rank, sorted(got)))
    for fd in got.values():
        C.close_fd(fd)


def test_descriptors_travel_between_processes():
    try:
        from stochastic_gradient_push_b200.ops import native
        native.load()
    except Exception as e:                       # no compiler on this host
        pytest.skip('native extension unavailable: %s' % e)
    world = 3
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    job = 'sgp_b200_test.%d' % os.getpid()
    procs = [ctx.Process(target=_worker, args=(r, world, job, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert out == [(r, [s for s in range(world) if s != r]) for r in range(world)]


def test_payload_is_readable_through_the_received_descriptor():
    try:
        from stochastic_gradient_push_b200.ops import native
        C = native.load()
    except Exception as e:
        pytest.skip('native extension unavailable: %s' % e)
    job = 'sgp_b200_self.%d' % os.getpid()
    a, b = C.FdChannel(job, 0, 2), C.FdChannel(job, 1, 2)
    r, w = os.pipe()
    os.write(w, b'payload')
    a.send(1, 7, r)
    src, tag, fd = b.recv()
    assert (src, tag) == (0, 7) and fd != r
    assert os.read(fd, 16) == b'payload'
    C.close_fd(fd)


def test_recv_times_out_instead_of_hanging_when_a_peer_never_sends():
    import time
    try:
        from stochastic_gradient_push_b200.ops import native
        C = native.load()
    except Exception as e:
        pytest.skip('native extension unavailable: %s' % e)
    ch = C.FdChannel('sgp_b200_timeout.%d' % os.getpid(), 0, 2)
    t0 = time.time()
    with pytest.raises(RuntimeError, match='timed out'):
        ch.recv(timeout_s=0.3)
    assert 0.25 < time.time() - t0 < 5.0
