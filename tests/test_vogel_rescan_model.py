"""Executable statement of VogelApproximate's incremental re-scan rule (rs_device.cuh vogel_approximate): a line (a row
over the slices with quota left, or a column over the free RBGs) reports (e1, first, e2) exactly as the reference's scan
computes them (downlink-transport-scheduler.cpp:392-414):

    if (e1 == -1 || e > e1) { first = k; e1 = e; continue; }
    if (e2 == -1 || e > e2) e2 = e;

When the line loses element x, the device re-scans it only if key(x) >= e2 AND NOT (key(x) == e1, x is not `first`, and
the line holds at least three elements equal to e1).  This model checks on random lines that every removal the rule
skips leaves (e1, first, e2) unchanged, and that the stored count of maxima stays a valid lower bound.  No GPU."""
import numpy as np


def scan(keys, alive):
    e1 = e2 = first = -1
    for k, e in enumerate(keys):
        if not alive[k]:
            continue
        if e1 == -1 or e > e1:
            first, e1 = k, e
            continue
        if e2 == -1 or e > e2:
            e2 = e
    return e1, first, e2


def test_skipped_removals_leave_the_candidate_unchanged():
    rng = np.random.default_rng(103)
    skipped = rescanned = 0
    for case in range(3000):
        n = int(rng.integers(2, 65))
        hi = int(rng.choice([2, 4, 16]))                       # few distinct keys: many ties, as with winner CQIs
        keys = rng.integers(16 - hi, 16, n)
        alive = rng.random(n) < 0.8
        if alive.sum() < 2:
            continue
        e1, first, e2 = scan(keys, alive)
        n_max = min(int(((keys == e1) & alive).sum()), 7)      # what the candidate word carries (saturating)
        while alive.sum() > 1:
            x = int(rng.choice(np.flatnonzero(alive)))
            kx = int(keys[x])
            alive[x] = False
            if kx >= e2:                                        # e2 == -1 ("no second") lands here too
                if kx == e1 and x != first and n_max >= 3:
                    n_max -= 1
                    skipped += 1
                else:
                    e1, first, e2 = scan(keys, alive)           # the device re-scans
                    n_max = min(int(((keys == e1) & alive).sum()), 7)
                    rescanned += 1
                    continue
            else:
                skipped += 1
            assert (e1, first, e2) == scan(keys, alive), (case, keys.tolist(), alive.tolist(), x)
            assert n_max <= int(((keys == e1) & alive).sum())
    assert skipped > 1000 and rescanned > 1000
