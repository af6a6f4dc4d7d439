"""GPU: radiosaber_b200/rs_batch, the C++ host program for batch runs (radiosaber_b200/host/rs_batch_main.cpp).
It links only the C ABI; what it prints must equal what the ctypes mirror computes from the same seeded inputs,
and its --log-cell text must equal the writer fed from the Python side."""
import json
import os
import subprocess

import numpy as np
import pytest

from radiosaber_b200 import sched, workload
from tests.helpers import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "radiosaber_b200", "rs_batch")
FIX = os.path.join(ROOT, "tests", "golden", "traces", "trace_subset.npz")


def _config(tmp_path, groups, ues):
    cfg = {"slices": [{"n_slices": n, "weight": w, "algo_alpha": a, "algo_beta": b, "algo_epsilon": e, "algo_psi": p}
                      for n, w, (a, b, e, p) in groups], "ues_per_slice": ues}
    path = tmp_path / "cfg.json"
    path.write_text(json.dumps(cfg, indent=2))
    return str(path)


def _run(*args):
    assert os.path.exists(BIN), "build with make -C radiosaber_b200/csrc"
    r = subprocess.run([BIN, *[str(a) for a in args]], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("algo", [9, 8, 7, 1, 10, 11, 101, 103])
def test_synthetic_batch_equals_python_mirror(algo, tmp_path):
    B, T, seed = 37, 35, 4
    path = _config(tmp_path, [(3, 0.2, (0, 0, 1, 1)), (2, 0.2, (0, 0, 1, 0))], [4, 2, 6, 1, 3])
    extra = ["--log-cell", 11, "--log-prefix", tmp_path / "cell11"]
    got = _run("--algo", algo, "--config", path, "--cells", B, "--ttis", T, "--seed", seed, *extra)
    w, p, u2s = sched.load_slice_config(path)
    S, U, G = len(w), len(u2s), 64
    g = sched.Scheduler(algo, w, p, u2s, B, cqi_per_rb=2)
    cqi = workload.synth_cqi(seed, 0, B, 0, T, U, G)
    if algo == 11:
        draws = workload.synth_rand_draws(seed, 0, B, 0, T, S, g.rand_stride)
    else:
        draws = workload.synth_rand2(seed, 0, B, 0, T, S)
    _, dts = workload.tti_clock(T)
    res = g.run_host(sched.pack_cqi(cqi), draws, dts, want_aux=True, ttis_per_launch=16)
    st = g.get_stats()
    assert got["slice_bytes"] == [int(x) for x in st[0]]
    assert got["slice_rbs"] == [int(x) for x in st[1]]
    assert got["cells"] == B and got["ttis"] == T and got["slices"] == S and got["ues"] == U
    assert int(st[1].sum()) > 0
    lw = sched.LogWriter(algo, u2s, S, cqi_per_rb=2)
    for t in range(T):
        if algo == 10:
            lw.tti_grants(100 + t, sched.pack_cqi(cqi[t, 11]), int(res["alloc_n"][t, 11]), res["alloc_ue"][t, 11],
                          res["alloc_rbg"][t, 11], res["tbs_bits"][t, 11], res["final_cqi"][t, 11],
                          res["slice_target"][t, 11], res["slice_quota"][t, 11])
            continue
        lw.tti(100 + t, sched.pack_cqi(cqi[t, 11]), res["rbg_to_ue"][t, 11], res["tbs_bits"][t, 11],
               res["final_cqi"][t, 11] if algo != 1 else None,
               res["slice_target"][t, 11] if algo in (8, 9, 101, 103) else None,
               res["slice_quota"][t, 11] if algo in (8, 9, 101, 103) else None)
    assert (tmp_path / "cell11.stdout").read_text() == lw.stdout
    assert (tmp_path / "cell11.stderr").read_text() == lw.stderr
    assert lw.stderr.count("\n") > T
    g.close()


@pytest.mark.parametrize("algo", [9, 7])
def test_trace_replay_batch_equals_python_mirror(algo, tmp_path):
    """cqi-traces-noise0 text files + mapping file -> rs_batch == set_traces/run_traces_host on the same arrays."""
    z = np.load(FIX)
    mapping, ids, rows = z["mapping"], z["trace_ids"], z["rows"]     # rows [K][n_rows][64]
    K, n_rows = 6, rows.shape[1]
    tdir = tmp_path / "traces"
    tdir.mkdir()
    for k in range(K):
        per_rb = np.repeat(rows[k], 8, axis=1)
        (tdir / f"ue{int(ids[k])}.log").write_text("\n".join(" ".join(str(int(v)) for v in r) for r in per_rb) + "\n")
    rng = np.random.default_rng(2)
    mp = ids[rng.integers(0, K, 23)]
    (tmp_path / "mapping.config").write_text("".join(f"{i} {int(t)}\n" for i, t in enumerate(mp)))
    B, T, seed = 29, 90, 9
    path = _config(tmp_path, [(4, 0.25, (0, 0, 1, 1))], [3, 5, 2, 4])
    got = _run("--algo", algo, "--config", path, "--cells", B, "--ttis", T, "--seed", seed,
               "--traces", tdir, "--mapping", tmp_path / "mapping.config", "--trace-rows", n_rows,
               "--log-cell", 5, "--log-prefix", tmp_path / "cell5")
    assert got["cqi"] == "trace replay"
    w, p, u2s = sched.load_slice_config(path)
    S, U = len(w), len(u2s)
    g = sched.Scheduler(algo, w, p, u2s, B, cqi_per_rb=2)
    n_tr = int(mp.max()) + 1                                          # the program indexes traces by file id
    traces = np.full((n_tr, n_rows, 512), 10, dtype=np.uint8)
    for k in range(K):
        traces[int(ids[k])] = np.repeat(rows[k], 8, axis=1)
    ue_trace = np.array([[mp[(u + 7 * b) % len(mp)] for u in range(U)] for b in range(B)], dtype=np.int32)
    g.set_traces(traces, ue_trace)
    now, dts = workload.tti_clock(T)
    tr = sched.trace_rows_for_run(now, 0, n_rows=n_rows)
    res = g.run_traces_host(tr, workload.synth_rand2(seed, 0, B, 0, T, S), dts, want_aux=True, ttis_per_launch=16)
    st = g.get_stats()
    assert got["slice_bytes"] == [int(x) for x in st[0]]
    assert got["slice_rbs"] == [int(x) for x in st[1]]
    assert int(st[0].sum()) > 0
    # the reference's log text of cell 5 in replay mode (the flag used to be accepted and ignored): the
    # runner rebuilds the cell's CQI rows from the traces; here they come from the same arrays
    lw = sched.LogWriter(algo, u2s, S, cqi_per_rb=2)
    for t in range(T):
        rows_t = np.stack([traces[ue_trace[5, u], tr[t], ::8] if tr[t] >= 0 else np.full(64, 10, np.uint8) for u in range(U)])
        lw.tti(100 + t, sched.pack_cqi(rows_t), res["rbg_to_ue"][t, 5], res["tbs_bits"][t, 5], res["final_cqi"][t, 5],
               res["slice_target"][t, 5] if algo == 9 else None, res["slice_quota"][t, 5] if algo == 9 else None)
    assert (tmp_path / "cell5.stdout").read_text() == lw.stdout
    assert (tmp_path / "cell5.stderr").read_text() == lw.stderr
    assert lw.stderr.count("\n") > T
    g.close()


def test_mapping_with_a_bad_trace_id_is_refused(tmp_path):
    path = _config(tmp_path, [(2, 0.5, (0, 0, 1, 1))], [2, 2])
    (tmp_path / "mapping.config").write_text("0 -3\n1 2\n")
    r = subprocess.run([BIN, "--algo", "9", "--config", path, "--cells", "4", "--ttis", "2", "--traces", str(tmp_path),
                        "--mapping", str(tmp_path / "mapping.config")], capture_output=True, text=True)
    assert r.returncode == 1 and "trace id -3" in r.stderr


def test_errors_are_loud(tmp_path):
    path = _config(tmp_path, [(2, 0.5, (0, 0, 1, 1))], [2, 2])
    r = subprocess.run([BIN, "--algo", "12", "--config", path, "--cells", "4", "--ttis", "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "scheduler id" in r.stderr
    r = subprocess.run([BIN, "--algo", "9", "--config", str(tmp_path / "none.json"), "--cells", "4", "--ttis", "2"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "configuration file" in r.stderr
    r = subprocess.run([BIN, "--algo", "9"], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
