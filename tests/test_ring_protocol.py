"""Model check of the tensor kernel's smem-ring handshake with several MMA-issuing warps
(domain_rag_system_b200/csrc/dense_umma.cu, DESIGN.md 4.1 "three issuing warps").

The kernel's rule: stage `st` lives in ring slot `st % n_stages` and is issued by warp `st % n_issuers`; a warp
waits ONLY for its own stages' `full` barriers and tests them by phase PARITY (mbarrier.try_wait.parity).  The
simulation below replays that protocol with random TMA / MMA latencies and reports a violation whenever a warp
passes its wait although the slot does not (yet) hold the stage it is about to multiply.  It shows why the plan
trims the ring to a multiple of the issuer count: then a slot is always consumed by the same warp, which has
observed every earlier phase of that slot, and parity is unambiguous."""
import random

import pytest


class Barrier:
    """mbarrier with one pending arrival per phase; try_wait(parity) == "the phase with this parity completed"."""

    def __init__(self):
        self.phase = 0

    def complete(self):
        self.phase += 1

    def try_wait(self, parity):
        return (self.phase & 1) != parity


def simulate(n_stages, n_issuers, n_total, seed):
    rng = random.Random(seed)
    full = [Barrier() for _ in range(n_stages)]
    empty = [Barrier() for _ in range(n_stages)]
    content = [None] * n_stages              # stage whose bytes the slot holds (None while a load is in flight)
    events = []                              # (time, kind, slot, stage)
    prod_st = 0
    iss_st = list(range(n_issuers))          # next stage of each issuing warp
    violations, t, done = 0, 0, 0
    while done < n_total and t < 500_000:
        t += 1
        for ev in [e for e in events if e[0] <= t]:
            events.remove(ev)
            _, kind, slot, stage = ev
            if kind == "landed":
                content[slot] = stage
                full[slot].complete()
            else:                                # "mma_done": tcgen05.commit arrives on empty[slot]
                empty[slot].complete()
                done += 1
        # producer: waits empty[slot] with parity (lap & 1) ^ 1, then issues the TMA load (completes later)
        if prod_st < n_total:
            slot, lap = prod_st % n_stages, prod_st // n_stages
            if empty[slot].try_wait((lap & 1) ^ 1):
                content[slot] = None
                events.append((t + rng.randint(20, 400), "landed", slot, prod_st))   # completions may reorder
                prod_st += 1
        # issuers: each waits full[slot] of ITS next stage with parity (lap & 1)
        for w in range(n_issuers):
            st = iss_st[w]
            if st >= n_total:
                continue
            slot, lap = st % n_stages, st // n_stages
            if full[slot].try_wait(lap & 1):
                if content[slot] != st:
                    violations += 1              # stale phase: the slot holds another stage or is being filled
                events.append((t + rng.randint(30, 60), "mma_done", slot, st))
                iss_st[w] = st + n_issuers
    return violations, done


@pytest.mark.parametrize("n_stages,n_issuers", [(4, 1), (3, 1), (4, 2), (3, 3), (6, 3), (6, 2)])
def test_ring_is_sound_when_every_slot_has_one_consumer(n_stages, n_issuers):
    for seed in range(8):
        violations, done = simulate(n_stages, n_issuers, 300, seed)
        assert violations == 0 and done == 300


def test_ring_breaks_when_slots_change_hands():
    """3 issuers on a 4-stage ring (the configuration that faulted on the GPU): some interleaving lets a warp read
    a slot on a stale phase."""
    assert sum(simulate(4, 3, 300, seed)[0] for seed in range(20)) > 0
    assert sum(simulate(5, 2, 300, seed)[0] for seed in range(20)) > 0


def test_plan_rule_matches_the_model():
    """dense_umma.cu make_plan: ns = (ns / ni) * ni."""
    for ns in range(3, 13):
        for ni in (2, 3):
            trimmed = (ns // ni) * ni
            assert trimmed % ni == 0 and 0 < trimmed <= ns
