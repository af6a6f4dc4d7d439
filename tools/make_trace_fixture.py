#!/usr/bin/env python3
"""Generate tests/golden/traces/trace_subset.npz from the reference's cqi-traces-noise0 directory.

The golden *_trace records (tools/make_golden.py) hold the CQI the unmodified reference ingested from
`cqi-traces-noise0/ue<id>.log` through `mapping1.config` (enb-mac-entity.cc:42-56, 160-193).  The
full trace set is 93 MB of text and does not travel to the GPU box, so this fixture keeps what the
trace-ingest parity tests need: the mapping file's trace ids and the first ROWS lines of every trace
a UE of those records replays (one value per RBG: the shipped traces are constant inside each 8-RB
group, which this script verifies over ALL lines of ALL 158 files).

Usage: python tools/make_trace_fixture.py   (only where /root/reference is mounted)
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_DIR = "/root/reference/cqi-traces-noise0"
ROWS = 8          # the golden records run 60-100 TTIs from t = 0.1 s: lines 2..4 are in force
N_UES = 204       # largest UE count of any shipped backlogged config (exp-backlogged-20slicesdiffw)


def main():
    mapping = np.array([int(l.split()[1]) for l in open(os.path.join(TRACE_DIR, "mapping1.config")) if l.strip()],
                       dtype=np.int32)
    ids = sorted(int(f[2:-4]) for f in os.listdir(TRACE_DIR) if f.startswith("ue") and f.endswith(".log"))
    hist = np.zeros(16, dtype=np.int64)
    full = {}
    for i in ids:
        t = np.loadtxt(os.path.join(TRACE_DIR, f"ue{i}.log"), dtype=np.int32)
        assert t.shape[1] == 512 and t.shape[0] >= 475, t.shape
        t = t[:475]
        assert t.min() >= 1 and t.max() <= 15
        g = t.reshape(475, 64, 8)
        assert (g == g[:, :, :1]).all(), f"ue{i}.log varies inside an RBG"
        hist += np.bincount(g[:, :, 0].ravel(), minlength=16)
        full[i] = g[:, :, 0].astype(np.uint8)
    used = sorted(set(int(mapping[u % len(mapping)]) for u in range(N_UES)))
    rows = np.stack([full[i][:ROWS] for i in used])
    out = os.path.join(ROOT, "tests", "golden", "traces", "trace_subset.npz")
    np.savez_compressed(out, mapping=mapping, trace_ids=np.asarray(used, dtype=np.int32), rows=rows,
                        n_rows_full=np.int32(475), histogram=hist)
    print(f"wrote {out}: {len(used)} traces x {ROWS} rows x 64 RBGs, mapping of {len(mapping)} lines")
    print("per-RBG CQI histogram over all 158 traces (CQI 1..15):", hist[1:].tolist())


if __name__ == "__main__":
    main()
