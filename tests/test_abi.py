"""CPU: the C-ABI library loads without a GPU and exports every symbol include/rs_sched.h declares;
the host-side config parser reads the reference's JSON slice format."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from radiosaber_b200 import sched

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rs_sched.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(sched.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(sched.LIB_PATH), "build with make -C radiosaber_b200/csrc"
    L = ctypes.CDLL(sched.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(L, name), name
    assert sched.lib().rs_abi_version() == 3


def test_create_rejects_bad_configs_without_touching_the_gpu():
    w = np.full(2, 0.5)
    p = np.array([[0, 0, 1, 1], [0, 0, 1, 2]], dtype=np.int32)  # psi = 2 is not bit-exact on device
    u2s = np.array([0, 1], dtype=np.int32)
    with pytest.raises(sched.RsError, match="psi"):
        sched.Scheduler(9, w, p, u2s, 1)
    p[1, 3] = 1
    with pytest.raises(sched.RsError, match="scheduler id"):
        sched.Scheduler(12, w, p, u2s, 1)
    with pytest.raises(sched.RsError, match="multiple"):
        sched.Scheduler(9, w, p, u2s, 1, n_rbs=25, rbg_size=2)
    with pytest.raises(sched.RsError, match="n_slices"):
        sched.Scheduler(9, np.full(65, 1 / 65), np.tile([0, 0, 1, 1], (65, 1)), np.arange(65), 1)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    w = np.full(2, 0.5)
    p = np.tile(np.array([0, 0, 1, 1], dtype=np.int32), (2, 1))
    with pytest.raises(sched.RsError, match="rs error 3"):
        sched.Scheduler(9, w, p, np.array([0, 1], dtype=np.int32), 1)


def test_load_slice_config(tmp_path):
    cfg = {"slices": [{"n_slices": 2, "weight": 0.3, "algo_alpha": 0, "algo_beta": 0, "algo_epsilon": 1,
                       "algo_psi": 1},
                      {"n_slices": 1, "weight": 0.4, "algo_alpha": 0, "algo_beta": 0, "algo_epsilon": 1,
                       "algo_psi": 0}],
           "ues_per_slice": [2, 1, 3]}
    path = tmp_path / "c.json"
    path.write_text(json.dumps(cfg))
    w, p, u2s = sched.load_slice_config(str(path))
    assert w.tolist() == [0.3, 0.3, 0.4]
    assert p.tolist() == [[0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 1, 0]]
    assert u2s.tolist() == [0, 0, 1, 2, 2, 2]


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "radiosaber_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "rs_oracle" not in txt and "librs_oracle" not in txt, f


def test_batch_runner_fails_loudly_without_a_gpu(tmp_path):
    """radiosaber_b200/rs_batch (C++ over the C ABI): config errors read like the reference's, and without a CUDA
    device it exits non-zero with the library's message instead of computing anything on the host."""
    import subprocess
    import torch
    exe = os.path.join(ROOT, "radiosaber_b200", "rs_batch")
    assert os.path.exists(exe), "build with make -C radiosaber_b200/csrc"
    r = subprocess.run([exe, "--algo", "9", "--config", str(tmp_path / "missing.json"), "--cells", "4", "--ttis", "2"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "Fail to open configuration file." in r.stderr   # downlink-transport-scheduler.cpp:58-60
    r = subprocess.run([exe, "--cells", "4"], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "--algo", "9", "--config", os.path.join(ROOT, "tests", "data", "cfg20x5.json"), "--cells", "4",
                        "--ttis", "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "rs_create" in r.stderr and r.stdout == ""


def test_nccl_library_exports_every_declared_symbol():
    """include/rs_sched_nccl.h -> radiosaber_b200/librs_nccl.so (loads without a GPU; no calls made)."""
    import torch  # noqa: F401  -- before the library: a process that loads the system libnccl first cannot import torch later
    src = open(os.path.join(ROOT, "include", "rs_sched_nccl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", src)))
    assert names == ["rs_comm_destroy", "rs_comm_init_all", "rs_comm_init_rank", "rs_comm_unique_id",
                     "rs_nccl_last_error", "rs_reduce_stats"]
    path = os.path.join(ROOT, "radiosaber_b200", "librs_nccl.so")
    assert os.path.exists(path), "build with make -C radiosaber_b200/csrc"
    L = ctypes.CDLL(path)
    for name in names:
        assert hasattr(L, name), name
