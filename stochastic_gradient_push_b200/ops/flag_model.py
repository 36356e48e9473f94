"""
Executable model of the flag / ack protocol of the synchronous and overlapped gossip kernels
(``csrc/sgp_common.cuh``, ``sgp_step_kernel`` / ``sgp_step_pipe_kernel`` / ``sgp_gather*``):

* every rank owns a double-buffered outbox (``outbox[s & 1]``), a publish sequence ``pub_seq``
  (released after the outbox of step s is written) and one ack sequence per reader
  (``ack_seq[r] == s + 1``: rank r finished reading our outbox of step s);
* **publish(s)** may overwrite ``outbox[s & 1]`` only after every out-neighbour of step s-2 acked
  it (the WAR fence, ``ack_seq[o] >= s - 1``);
* **pull(s)** may read an in-neighbour's ``outbox[s & 1]`` only after its ``pub_seq >= s``, and
  acks afterwards;
* in the synchronous step publish(s) and pull(s) are one kernel, so publish(s) follows pull(s - 1);
  ``overlap=True`` explores the WEAKER ordering of the side-stream gather with bounded staleness
  (``synch_freq > 0``): the gather of step s - 1 may still be in flight when step s is published --
  only the WAR fence and the publish flags protect the outbox then.

``tests/test_flag_protocol_model.py`` drives it with a random scheduler (arbitrary rank skew) over
every shipped topology and checks that a reader always finds the data of the step it asked for (no
torn / overwritten outbox), that the schedule never deadlocks and that the ranks' step counters
never drift more than the protocol allows.  The model is for tests / documentation only.
"""

from __future__ import annotations

from typing import Callable, List, Tuple


class FlagRank(object):

    def __init__(self, rank: int, peers: Callable[[int, int], Tuple[List[int], List[int]]]):
        self.rank = rank
        self.peers = peers                    # (step, rank) -> (out_peers, in_peers)
        self.step = 0                         # next step to publish
        self.pulled = 0                       # next step to pull
        self.pub_seq = -1
        self.ack_seq = {}                     # reader -> highest (step + 1) acked
        self.outbox = [None, None]            # [parity] -> step tag of the data in it
        self.errors = []

    # -- phase 1 ---------------------------------------------------------------- #
    def can_publish(self, world, overlap: bool) -> bool:
        s = self.step
        if self.pulled < (s if not overlap else s - 1):
            return False                      # sync: pull(s-1) finished; overlap: at most one gather in flight
        if s >= 2:
            outs, _ = self.peers(s - 2, self.rank)
            for o in outs:
                if o != self.rank and self.ack_seq.get(o, 0) < s - 1:
                    return False              # WAR fence
        return True

    def publish(self, world):
        s = self.step
        old = self.outbox[s & 1]
        if old is not None:
            outs, _ = self.peers(old, self.rank)
            for o in outs:
                if o != self.rank and self.ack_seq.get(o, 0) < old + 1:
                    self.errors.append('rank %d overwrote outbox of step %d before %d read it' % (self.rank, old, o))
        self.outbox[s & 1] = s
        self.pub_seq = s
        self.step = s + 1

    # -- phase 2 ---------------------------------------------------------------- #
    def can_pull(self, world) -> bool:
        s = self.pulled
        if s >= self.step:
            return False                      # own publish(s) comes first (same kernel / stream order)
        _, ins = self.peers(s, self.rank)
        return all(world[j].pub_seq >= s for j in ins)

    def pull(self, world):
        s = self.pulled
        _, ins = self.peers(s, self.rank)
        for j in ins:
            tag = world[j].outbox[s & 1]
            if tag != s:
                self.errors.append('rank %d read step %s data of rank %d while pulling step %d' % (self.rank, tag, j, s))
            if j != self.rank:
                world[j].ack_seq[self.rank] = s + 1
        self.pulled = s + 1
