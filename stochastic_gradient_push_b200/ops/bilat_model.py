"""
Executable model of the device-side AD-PSGD round protocol (``csrc/sgp_kernels.cu``:
``sgp_bilat_decide_kernel`` + ``sgp_step_kernel`` launched with ``SGP_F_FROM_STATE``).

It mirrors, flag for flag, what the two kernels read and write -- the per-rank round counter
(``step``), ``bilat_published`` / ``bilat_budget`` / ``bilat_enabled``, the partner-visible publish
sequence (``pub_seq``), the ack sequence and the double-buffered outbox -- but on plain Python
floats, so that the *protocol* can be explored on a CPU: ``tests/test_bilat_protocol_model.py`` runs
it under random interleavings of the ranks' {decide, work} launches, random partner delays and random
gradient arrivals and checks the properties the kernels rely on (every round is a pairwise average
executed exactly once by both partners, an outbox half is never overwritten while its reader may
still need it, the sum over ranks is conserved when no gradients arrive, nobody waits forever).

This is a model for tests and documentation, not a fallback data plane (the portable data plane of
AD-PSGD is the c10d loop in ``parallel/ad_psgd.py``).
"""

from __future__ import annotations

from typing import List, Optional

PUBLISH, PULL = 1, 2


class BilatRank(object):
    """State of ONE rank: what lives in its ``SgpState`` / signal pad / outbox."""

    def __init__(self, rank: int, partners_by_round, passive: bool, x: float, budget: Optional[int]):
        self.rank = rank
        self.partner_of = partners_by_round      # round -> partner rank (bilateral: symmetric)
        self.passive = passive
        self.x = float(x)                         # the gossip copy (one scalar stands for the arena)
        self.step = 0                             # rounds completed (SgpState::step)
        self.published = False                    # SgpState::bilat_published
        self.budget = budget                      # SgpState::bilat_budget (None = unbounded)
        self.enabled = True
        self.cmd = 0                              # SgpState::bilat_cmd
        # peer-visible
        self.pub_seq = -1                         # highest round whose snapshot is in the outbox
        self.ack_seq = {}                         # reader rank -> rounds of OURS it has finished reading
        self.outbox = [None, None]                # [parity] -> (round, value)
        # bookkeeping for the checks
        self.rounds_done = []                     # (round, partner, own value used, partner value used)
        self.overwrites_while_unread = 0

    # -- sgp_bilat_decide_kernel ------------------------------------------------ #
    def decide(self, world: List['BilatRank'], partner_ready_within_wait: bool = True):
        s = self.step
        partner = world[self.partner_of(s, self.rank)]
        may_start = self.enabled and (self.budget is None or self.budget > 0)
        engaged = self.published or may_start
        ready = engaged and partner_ready_within_wait and partner.pub_seq >= s
        acks_ok = True
        if s >= 2:
            reader = world[self.partner_of(s - 2, self.rank)]
            acks_ok = self.ack_seq.get(reader.rank, 0) >= s - 1
        do_publish = (not self.published) and may_start and acks_ok and (ready or not self.passive)
        do_pull = ready and (self.published or do_publish)
        self.cmd = (PUBLISH if do_publish else 0) | (PULL if do_pull else 0)
        if do_publish and self.budget is not None:
            self.budget -= 1
        return self.cmd

    # -- worker launch (sgp_step_kernel, SGP_F_FROM_STATE) ---------------------- #
    def work(self, world: List['BilatRank']):
        cmd, self.cmd = self.cmd, 0
        s = self.step
        if cmd & PUBLISH:
            slot = s & 1
            old = self.outbox[slot]
            if old is not None:
                reader = world[self.partner_of(old[0], self.rank)]
                if self.ack_seq.get(reader.rank, 0) < old[0] + 1:
                    self.overwrites_while_unread += 1          # WAR hazard (must never happen)
            self.outbox[slot] = (s, self.x)                    # snapshot (KEEP_Z: x itself untouched)
            self.pub_seq = s
            self.published = True
        if cmd & PULL:
            partner = world[self.partner_of(s, self.rank)]
            rnd, val = partner.outbox[s & 1]
            assert rnd == s, 'rank %d pulled round %d data while in round %d' % (self.rank, rnd, s)
            self.rounds_done.append((s, partner.rank, self.x, val))
            self.x = 0.5 * self.x + 0.5 * val                 # SELF_FROM_Z: the CURRENT x, not the snapshot
            partner.ack_seq[self.rank] = s + 1                 # release the partner's outbox half
            self.published = False
            self.step = s + 1

    def apply_gradient(self, delta: float, budget: Optional[int]):
        """what the training thread enqueues under the daemon lock: fused SGD + budget refill"""
        self.x += delta
        self.budget = budget
