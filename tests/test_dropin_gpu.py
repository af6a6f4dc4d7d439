"""GPU: the drop-in itself.  oracle/_ref/ref_harness_gpu is the reference's own LTE-Sim
(SingleCellWithI scenario, unmodified sources) with the product's host plug-in
(radiosaber_b200/host/rs_gpu_scheduler.h -> C ABI -> CUDA) installed in the eNB instead of the
reference scheduler class (ids 1, 7, 8, 9, 10, 11, 101, 103).  It must reproduce, record for record, what the reference classes
produced on the same CQI / rand() inputs (tests/golden)."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests.helpers import ROOT, load_golden
from tools import golden_io

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness_gpu")

CASES = ["a9_fix20x5_synth", "a9_diffw_synth", "a9_diffw_trace", "a9_small_synth", "a8_fix20x5_synth", "a8_small_synth",
         "a7_fix20x5_synth", "a7_mix20_synth", "a7_small_synth", "a1_fix20x5_synth", "a1_small_synth",
         # finite queues and head-of-line delays (internet flows): the plug-in reads them from the bearers
         "a9_qif_synth", "a8_qif_synth", "a7_qif_synth", "a1_qif_synth",
         # the other inter-slice algorithms of DownlinkTransportScheduler: UpperBound (grants as a list), SubOpt, Vogel
         "a10_fix20x5_synth", "a10_diffw_synth", "a10_small_synth", "a10_qif_synth",
         "a101_fix20x5_synth", "a101_diffw_synth", "a101_small_synth", "a101_qif_synth",
         "a103_fix20x5_synth", "a103_diffw_synth", "a103_small_synth", "a103_qif_synth",
         # NVS non-greedy: the plug-in takes 300 x (listed users of the served slice) rand() values per TTI from the process
         "a11_fix20x5_synth", "a11_small_synth", "a11_qif_synth"]


@pytest.mark.parametrize("name", CASES)
def test_plugin_inside_lte_sim_matches_reference(name, tmp_path):
    if not os.path.exists(HARNESS):
        pytest.fail("oracle/_ref/ref_harness_gpu missing: run __graft_entry__.build() where /root/reference is mounted")
    rec = load_golden(name)
    S, U, T = int(rec["S"]), int(rec["U"]), int(rec["T"])
    cfg = {"slices": [], "ues_per_slice": [int((rec["ue_to_slice"] == s).sum()) for s in range(S)]}
    for s in range(S):
        a, b, e, p = (int(x) for x in rec["params"][s])
        cfg["slices"].append({"n_slices": 1, "weight": float(rec["weight"][s]), "video_app": 0, "video_bitrate": 0,
                              "internet_flow": 0, "if_bitrate": 0, "backlog_flow": 1, "algo_alpha": a,
                              "algo_beta": b, "algo_epsilon": e, "algo_psi": p})
    if "queue" in rec:   # the traffic mix is part of the scenario: the recorded run's own config
        cfg = json.loads(str(rec["config_json"]))
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    assert int(rec["cqi_per_rb"]) == 0
    np.ascontiguousarray(rec["cqi"], dtype=np.uint8).tofile(tmp_path / "cqi.bin")
    np.ascontiguousarray(rec["rand2"], dtype=np.int32).tofile(tmp_path / "rand.bin")
    out = tmp_path / "rec.bin"
    algo = int(rec["algo"])
    extra = ["--alloc-log", str(tmp_path / "alloc.bin")] if algo == 10 else []
    if algo == 11:
        extra = ["--rand-log", str(tmp_path / "draws.bin")]
    r = subprocess.run([HARNESS, "--gpu", "--algo", str(algo), "--config", str(tmp_path / "cfg.json"),
                        "--ttis", str(T), "--cqi", str(tmp_path / "cqi.bin"), "--rand", str(tmp_path / "rand.bin"),
                        "--out", str(out), "--seed", str(int(rec["seed"]))] + extra, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-500:])
    got = golden_io.compact(golden_io.parse_record_stream(str(out)))
    assert int(got["T"]) == T
    fields = ["rbg_to_ue", "bits", "final_cqi", "avg_before", "avg_after", "tx_after", "cum_bytes", "cum_rbs",
              "state_before", "state_after", "dt"]
    if algo in (8, 9, 10, 101, 103):
        fields += ["target", "quota", "rand2"]
    if algo in (7, 11):
        fields += ["nvs_slice"]
    for f in fields:
        assert np.array_equal(np.asarray(got[f]), np.asarray(rec[f])), (name, f)
    if algo == 11:   # every rand() value drawn inside DoSchedule: the same count and the same values as the reference's
        raw = np.fromfile(tmp_path / "draws.bin", dtype="<i4")
        pos = 0
        for t in range(T):
            k = int(raw[pos])
            assert k == int(rec["rand_ng_n"][t]), (name, t)
            assert np.array_equal(raw[pos + 1:pos + 1 + k], rec["rand_ng"][t, :k]), (name, t)
            pos += 1 + k
        assert pos == len(raw)
    if algo == 10:   # every (user, RBG) grant of every TTI, in the order of the users' RB lists
        raw = np.fromfile(tmp_path / "alloc.bin", dtype="<i2")
        pos = 0
        for t in range(T):
            k = int(raw[pos]) | (int(raw[pos + 1]) << 16)
            pairs = raw[pos + 2:pos + 2 + 2 * k].reshape(k, 2)
            pos += 2 + 2 * k
            assert k == int(rec["alloc_n"][t]), (name, t)
            assert np.array_equal(pairs[:, 0], rec["alloc_ue"][t, :k]) and np.array_equal(pairs[:, 1], rec["alloc_rbg"][t, :k]), (name, t)
        assert pos == len(raw)


# ---- two bearers per UE (internet_flow: 2) ---------------------------------------------------------------------
# The record format is per UE, so here the golden is the reference's own log text (tools/make_golden_logs_two_bearers.py):
# the plug-in inside LTE-Sim must print the same allocation dump (stdout) and the same per-bearer cumu_bytes / cumu_rbs
# lines (stderr) -- which only happens if every TTI's allocation, every bearer's share of the bytes and, through the
# bearers' own EWMA, every later TTI are identical.
TWO_BEARER_LOGS = os.path.join(ROOT, "tests", "golden", "logs_two_bearers")


def _scheduler_lines(text):
    """Drop what is not scheduler output: the RLC's "ipflow end ..." lines and the inter-slice function's all_bytes."""
    return [l for l in text.splitlines() if not l.startswith("ipflow ") and not l.startswith("all_bytes")]


@pytest.mark.parametrize("algo", [9, 8, 7, 10, 101, 103])
def test_two_bearers_per_ue_inside_lte_sim(algo, tmp_path):
    from tools import make_golden_logs_two_bearers as mb
    if not os.path.exists(HARNESS):
        pytest.fail("oracle/_ref/ref_harness_gpu missing: run __graft_entry__.build() where /root/reference is mounted")
    cfg = json.load(open(mb.CFG))
    cqi, rnd = mb.write_inputs(str(tmp_path), cfg)
    cmd = mb.command(HARNESS, algo, cqi, rnd, str(tmp_path / "got"), cfg)
    r = subprocess.run(cmd[:1] + ["--gpu"] + cmd[1:], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-500:])
    for ext in ("stdout", "stderr"):
        want = _scheduler_lines(open(os.path.join(TWO_BEARER_LOGS, f"a{algo}.{ext}")).read())
        got = _scheduler_lines((tmp_path / f"got.{ext}").read_text())
        assert len(want) > mb.TTIS
        for k, (a, b) in enumerate(zip(want, got)):
            assert a == b, (algo, ext, k, a, b)
        assert len(got) == len(want), (algo, ext)
