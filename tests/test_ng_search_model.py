"""Executable statement of why the RBG-bitmask form of NVS non-greedy's sample score (rs_device.cuh ng_search) is exact:
the reference (downlink-nvs-scheduler.cpp:447-460) adds, RBG by RBG, the largest metric among the users whose drawn
MCS their CQI on that RBG reaches (a user's metric does not depend on the RBG); the device sorts the users by metric,
gives every RBG to the first user in that order whose "CQI >= MCS" bit is set, and adds the same doubles in the same
order.  Pure Python / numpy, no GPU."""
import numpy as np


def score_reference(metric, mcs, cqi):
    """metric [Q][16] doubles, mcs [Q] ints, cqi [Q][G] ints -> the sample's score as the reference accumulates it."""
    Q, G = cqi.shape
    pf = 0.0
    for g in range(G):
        highest = -1.0
        for q in range(Q):
            m = float(metric[q, mcs[q]]) if mcs[q] <= cqi[q, g] else 0.0
            if highest < m:
                highest = m
        pf = pf + highest
    return pf


def score_bitmask(metric, mcs, cqi):
    Q, G = cqi.shape
    masks = [sum(1 << g for g in range(G) if cqi[q, g] >= mcs[q]) for q in range(Q)]
    vals = [float(metric[q, mcs[q]]) for q in range(Q)]
    order = sorted(range(Q), key=lambda q: -vals[q])          # any order among equal metrics: only the value is added
    seen, owned = 0, []
    for q in order:
        owned.append((vals[q], masks[q] & ~seen))
        seen |= masks[q]
    pf = 0.0
    for g in range(G):
        v = 0.0
        for val, m in owned:
            if (m >> g) & 1:
                v = val
        pf = pf + v
    return pf


def test_bitmask_score_equals_reference_score():
    rng = np.random.default_rng(11)
    for case in range(400):
        Q, G = int(rng.integers(1, 9)), 64
        cqi = rng.integers(1, 16, (Q, G))
        if case % 5 == 0:
            cqi[:] = cqi[:, :1]                                # flat rows: whole-row ties
        metric = rng.random((Q, 16)) * rng.choice([1e-3, 1.0, 1e6])
        if case % 7 == 0:
            metric[1:] = metric[0]                             # users with identical metrics
        metric[:, 0] = 0.0
        hc = cqi.max(axis=1)
        mcs = np.maximum(hc - rng.integers(0, 4, Q), 1)        # max(best CQI - rand() % 4, 1)
        assert score_bitmask(metric, mcs, cqi) == score_reference(metric, mcs, cqi), case
