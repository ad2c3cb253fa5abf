"""CPU tier: a model of the classed kernel's phase gates (csrc/drt_render.cuh: gate(), LOCKSTEP).

Every warp of a CTA runs the sequence  claim a pixel -> [gate, phase 1, gate, phase 2] per batch -> claim ...  and leaves with
mbarrier.arrive_drop when the pixels run out; pixels that need no tracing (culled, out of the band) are claimed without passing a gate.
The gates are ONE mbarrier with one arrival per warp per gate and parity waits.  This model plays that sequence with the mbarrier
semantics (arrive: pending -= 1; arrive_drop: expected -= 1 as well; a phase completes when pending reaches 0 and re-arms with the
current expected count; a parity wait returns once the phase of that parity has completed) under random phase durations, warp counts,
task counts, batch counts and cull rates, and checks what the kernel relies on: every warp always leaves (no schedule deadlocks), and at
any time all warps inside a phase are in the SAME phase (the point of the gates: one phase's code in the instruction cache)."""
import heapq
import random

import pytest


def simulate(nwarps, ntasks, batches, p_cull, seed):
    rnd = random.Random(seed)
    expected = pending = nwarps
    phase = 0
    counter = 0
    parity = [0] * nwarps
    state = ["claim"] * nwarps
    left = [0] * nwarps
    waiting = {}
    running = {}                     # warp -> "p1" / "p2" while it is inside a phase
    done = 0
    ready = [(0.0, w) for w in range(nwarps)]
    heapq.heapify(ready)

    def arrive(drop):
        nonlocal expected, pending, phase
        if drop:
            expected -= 1
        pending -= 1
        if pending == 0:
            phase += 1
            pending = expected
            return True
        return False

    def release(now):
        for w, par in list(waiting.items()):
            if phase % 2 != par:
                del waiting[w]
                heapq.heappush(ready, (now, w))

    steps = 0
    while ready:
        now, w = heapq.heappop(ready)
        steps += 1
        assert steps < 5_000_000
        s = state[w]
        if s == "claim":
            running.pop(w, None)
            if counter >= ntasks:
                state[w] = "exit"
                done += 1
                if arrive(True):
                    release(now)
                continue
            counter += 1
            if rnd.random() < p_cull:                      # a pixel that needs no tracing: no gate
                heapq.heappush(ready, (now + 0.01, w))
                continue
            left[w] = batches
            state[w] = "gateA"
            heapq.heappush(ready, (now + 0.01, w))
        elif s in ("gateA", "gateB"):
            running.pop(w, None)
            completed = arrive(False)
            mine = parity[w]
            parity[w] ^= 1
            state[w] = "p1" if s == "gateA" else "p2"
            if phase % 2 != mine:
                heapq.heappush(ready, (now, w))
            else:
                waiting[w] = mine
            if completed:
                release(now)
        else:                                              # entering phase 1 or phase 2
            running[w] = s
            assert len(set(running.values())) == 1, "warps of one CTA in different phases"
            if s == "p1":
                state[w] = "gateB"
                heapq.heappush(ready, (now + rnd.expovariate(1.0), w))
            else:
                left[w] -= 1
                state[w] = "gateA" if left[w] > 0 else "claim"
                heapq.heappush(ready, (now + rnd.expovariate(2.0), w))
    return done == nwarps and not waiting


@pytest.mark.parametrize("seed", range(400))
def test_every_warp_leaves_and_phases_never_mix(seed):
    r = random.Random(1000 + seed)
    assert simulate(r.choice([1, 2, 4, 8, 16]), r.choice([0, 1, 3, 17, 64, 200]), r.choice([1, 2, 4, 32]), r.choice([0.0, 0.3, 0.9]), seed)
