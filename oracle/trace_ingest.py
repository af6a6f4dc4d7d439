"""CPU restatement of the reference's trace-driven CQI ingest.  TEST INFRASTRUCTURE ONLY (see
oracle/pyoracle.py): the product path (radiosaber_b200/) never imports this module.

Follows EnbMacEntity under USE_REAL_TRACE, src/protocolStack/mac/enb-mac-entity.cc:
  :42-56    constructor: mapping.config is read as "uid tid" pairs; only tid is kept, in file order
  :160-193  ReceiveCqiIdealControlMessage: UE u replays trace mapping[u % len(mapping)]; the file
            ue<tid>.log is read as MAX_TTI_TRACE = 475 lines of nb_rbs integers; every report sets
            the UE record's CQI vector to line (int)(Now*1000 / CQI_INTERVAL) % 475
and ENodeB::UserEquipmentRecord's constructor (src/device/ENodeB.cpp:207-217): CQI 10 on every RB
until the first report.

Pinned by tests/test_trace_ingest.py against the CQI vectors the unmodified reference ingested
(tests/golden/*_trace.npz, recorded by oracle/ref_harness.cpp).
"""
from __future__ import annotations

import numpy as np

CQI_INTERVAL = 40     # enb-mac-entity.cc:38
MAX_TTI_TRACE = 475   # enb-mac-entity.cc:40
INITIAL_CQI = 10      # ENodeB.cpp:207-217


def read_mapping(path) -> np.ndarray:
    """enb-mac-entity.cc:48-55: `while (ifs >> uid >> tid) m_userMapping.push_back(tid);`"""
    toks = open(path).read().split()
    out = []
    for k in range(0, len(toks) - 1, 2):
        try:
            int(toks[k])
            out.append(int(toks[k + 1]))
        except ValueError:
            break
    return np.asarray(out, dtype=np.int32)


def read_trace(path, n_rows=MAX_TTI_TRACE, n_rbs=512) -> np.ndarray:
    """enb-mac-entity.cc:169-187: getline per row, `iss >> cqi` n_rbs times (a failed extraction leaves
    the previous value in place)."""
    out = np.empty((n_rows, n_rbs), dtype=np.uint8)
    cqi = 0
    with open(path) as f:
        for i in range(n_rows):
            toks = f.readline().split()
            ok = True
            for j in range(n_rbs):
                if ok and j < len(toks):
                    try:
                        cqi = int(toks[j])
                    except ValueError:
                        ok = False
                else:
                    ok = False
                out[i, j] = cqi
    return out


def trace_of_ue(mapping, ue_id) -> int:
    """enb-mac-entity.cc:164: m_userMapping[user_id % m_userMapping.size()]"""
    return int(mapping[ue_id % len(mapping)])


def trace_row(now, n_rows=MAX_TTI_TRACE) -> int:
    """enb-mac-entity.cc:189-191: `int time_stamp = Now()*1000 / CQI_INTERVAL; ... [time_stamp % size]`"""
    return int(np.float64(now) * np.float64(1000) / np.float64(CQI_INTERVAL)) % n_rows


def rows_for_run(now, first_report_tti=0, interval=CQI_INTERVAL, n_rows=MAX_TTI_TRACE) -> np.ndarray:
    """Line in force at each TTI when reports arrive at TTIs first_report_tti + k*interval; -1 = none yet."""
    rows = np.full(len(now), -1, dtype=np.int32)
    cur = -1
    for t in range(len(now)):
        if t >= first_report_tti and (t - first_report_tti) % interval == 0:
            cur = trace_row(now[t], n_rows)
        rows[t] = cur
    return rows


def cqi_at(traces, ue_trace, row) -> np.ndarray:
    """The CQI vectors the UE records hold: traces [n][rows][C], ue_trace [B][U] -> uint8 [B][U][C]."""
    traces = np.asarray(traces)
    ue_trace = np.asarray(ue_trace)
    if row < 0:
        return np.full(ue_trace.shape + (traces.shape[2],), INITIAL_CQI, dtype=np.uint8)
    out = np.ascontiguousarray(traces[np.maximum(ue_trace, 0), row]).astype(np.uint8)
    out[ue_trace < 0] = INITIAL_CQI   # no report ever received from this UE: the record's initial vector
    return out
