"""Host-side model of the lockstep launch of the default E kernel (thunder_b200/csrc/thb_expect7.cuh): one arrival counter per
(wave, barrier), the counter a CTA waits for, partial last waves, finished particles that arrive at all barriers of their wave at
once.  The model restates the kernel's integer arithmetic and checks, under random scheduling, that (a) every CTA finishes (no
deadlock, whatever the window), (b) a CTA that has passed the wait of barrier J knows that every CTA of that barrier's wave has
arrived at barrier J - window - i.e. no CTA is ever more than window + 1 barriers ahead of the slowest one.  (A single running
counter compared with (J - window + 1) x CTAs, the first version of the kernel, only has property (b) for window 0: with a window,
fast CTAs' arrivals at later barriers are counted for slow ones - this test is what showed it.)
The GPU tests check the kernel itself (tests/test_gpu_hotpath.py::test_expect_lockstep_radial_order_equals_free_running)."""
import random

import pytest


@pytest.mark.parametrize("G,nAct,K,W", [(4, 10, 5, 0), (4, 10, 5, 2), (8, 8, 3, 1), (8, 21, 4, 6), (3, 2, 7, 1), (16, 50, 2, 3)])
@pytest.mark.parametrize("frac_inactive", [0.0, 0.3])
def test_lockstep_counter_model(G, nAct, K, W, frac_inactive):
    rng = random.Random(G * 1000 + nAct * 10 + K + W)
    active = [rng.random() >= frac_inactive for _ in range(nAct)]
    grid = min(G, nAct)
    nWaves = (nAct + grid - 1) // grid
    ctr = [0] * (nWaves * K)                                   # lockCtr[wave * K + j]
    cta = [dict(it=c, j=0, st=0, done=False) for c in range(grid)]     # st 0: about to arrive at barrier j, 1: waiting
    progress = lambda c: (c["it"] // grid) * K + c["j"] + c["st"]     # barriers this CTA has arrived at
    steps = 0
    while not all(c["done"] for c in cta):
        steps += 1
        assert steps < 10 ** 6, "no progress"
        movable = []
        for i, c in enumerate(cta):
            if c["done"]:
                continue
            if c["st"] == 0 or not active[c["it"]]:
                movable.append(i)
                continue
            Jg = (c["it"] // grid) * K + c["j"] - W
            if Jg < 0 or ctr[Jg] >= min(grid, nAct - (Jg // K) * grid):
                movable.append(i)
        assert movable, "every CTA is waiting: deadlock"
        c = cta[rng.choice(movable)]
        wave = c["it"] // grid
        if not active[c["it"]]:
            for j in range(K):                                 # a finished particle: all barriers of its wave at once
                ctr[wave * K + j] += 1
            c["it"] += grid
            c["j"] = 0
        elif c["st"] == 0:
            ctr[wave * K + c["j"]] += 1                        # arrive
            c["st"] = 1
        else:
            Jg = wave * K + c["j"] - W
            if Jg >= 0:                                        # passing the wait: every CTA of that wave has arrived at barrier Jg
                for o in cta:
                    if o["it"] < nAct and o["it"] // grid <= Jg // K:
                        assert progress(o) > Jg or not active[o["it"]], (progress(o), Jg)
            c["st"] = 0
            c["j"] += 1
            if c["j"] == K:
                c["it"] += grid
                c["j"] = 0
        if c["it"] >= nAct:
            c["done"] = True
    for w in range(nWaves):
        assert all(ctr[w * K + j] == min(grid, nAct - w * grid) for j in range(K))
