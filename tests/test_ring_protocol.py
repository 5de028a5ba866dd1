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


def simulate_stream_ring(n_stages, n_consumers, n_chunks, seed):
    """dense_stream_kernel (csrc/dense_simt.cu): ONE producer lane, n_consumers warps that each read EVERY chunk.
    full[s] completes when the bulk copy of the chunk has landed (one arrive.expect_tx + the copy's bytes); empty[s]
    completes when all n_consumers warps have arrived.  Producer: waits empty[s] with parity ((c / S) - 1) & 1 for
    c >= S only (the first lap finds the ring empty).  Consumer: waits full[s] with parity (c / S) & 1.
    A violation = a consumer reads a slot that does not hold its chunk, or the producer overwrites a slot some warp
    has not finished reading."""
    rng = random.Random(seed)
    full_phase = [0] * n_stages
    empty_phase = [0] * n_stages
    empty_arrivals = [0] * n_stages
    content = [None] * n_stages
    readers_done = [n_consumers] * n_stages          # warps that finished the slot's current content
    events = []
    prod_c = 0
    cons_c = [0] * n_consumers
    busy_until = [0] * n_consumers
    violations, t = 0, 0
    while min(cons_c) < n_chunks and t < 400_000:
        t += 1
        for ev in [e for e in events if e[0] <= t]:
            events.remove(ev)
            _, kind, slot, chunk, w = ev
            if kind == "landed":
                content[slot] = chunk
                readers_done[slot] = 0
                full_phase[slot] += 1
            else:                                         # a warp finished reading: arrives on empty[slot]
                readers_done[slot] += 1
                empty_arrivals[slot] += 1
                if empty_arrivals[slot] == n_consumers:
                    empty_arrivals[slot] = 0
                    empty_phase[slot] += 1
        if prod_c < n_chunks:
            s, lap = prod_c % n_stages, prod_c // n_stages
            ok = True if prod_c < n_stages else ((empty_phase[s] & 1) != ((lap - 1) & 1))
            if ok:
                if readers_done[s] != n_consumers:
                    violations += 1                       # overwriting a slot that is still being read
                content[s] = None
                events.append((t + rng.randint(50, 600), "landed", s, prod_c, -1))
                prod_c += 1
        for w in range(n_consumers):
            c = cons_c[w]
            if c >= n_chunks or busy_until[w] > t:
                continue
            s, lap = c % n_stages, c // n_stages
            if (full_phase[s] & 1) != (lap & 1):          # try_wait.parity(lap & 1) succeeded
                if content[s] != c:
                    violations += 1
                d = rng.randint(20, 300)
                busy_until[w] = t + d
                events.append((t + d, "read_done", s, c, w))
                cons_c[w] = c + 1
    return violations, min(cons_c)


@pytest.mark.parametrize("n_stages", [2, 3, 4, 8])
def test_stream_ring_never_reads_stale_or_overwrites_live_slots(n_stages):
    for seed in range(4):
        violations, done = simulate_stream_ring(n_stages, 8, 60, seed)
        assert violations == 0 and done == 60
