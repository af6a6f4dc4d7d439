"""Reader for the record stream written by oracle/ref_harness.cpp and the
compact .npz golden fixtures derived from it (tests/golden/*.npz).

Test infrastructure only.
"""
from __future__ import annotations

import numpy as np

MAGIC = b"RSGOLD1\x00"
TTI_MARK = 0x54544921


def parse_record_stream(path: str) -> dict:
    """Parse a ref_harness --out file into a dict of stacked numpy arrays."""
    raw = open(path, "rb").read()
    if raw[:8] != MAGIC:
        raise ValueError(f"{path}: bad magic")
    pos = 8

    def take(dtype, n):
        nonlocal pos
        dt = np.dtype(dtype)
        arr = np.frombuffer(raw, dtype=dt, count=n, offset=pos)
        pos += dt.itemsize * n
        return arr

    algo, S, U, R, rbg = (int(x) for x in take("<i4", 5))
    G = R // rbg
    out = {
        "algo": algo, "S": S, "U": U, "R": R, "rbg_size": rbg, "G": G,
        "weight": take("<f8", S).copy(),
        "params": take("<i4", 4 * S).reshape(S, 4).copy(),
        "ue_to_slice": take("<i4", U).copy(),
    }
    fields = [
        ("now", "<f8", 1), ("avg_before", "<f8", U), ("tx_before", "<i4", U), ("last_update", "<f8", U),
        ("state_before", "<f8", S), ("cqi_rb", "u1", U * R), ("active", "u1", U), ("rand2", "<i4", 2),
        ("rbg_to_ue", "<i2", G), ("bits", "<i4", U), ("final_cqi", "u1", U),
        ("target", "<i4", S), ("quota", "<i4", S), ("nvs_slice", "<i4", 1),
        ("avg_after", "<f8", U), ("tx_after", "<i4", U), ("cum_bytes", "<u8", U), ("cum_rbs", "<u8", U),
        ("state_after", "<f8", S),
    ]
    cols = {name: [] for name, _, _ in fields}
    t = 0
    while pos < len(raw):
        mark, tti = (int(x) for x in take("<i4", 2))
        if mark != TTI_MARK or tti != t:
            raise ValueError(f"{path}: bad TTI header at byte {pos}")
        for name, dt, n in fields:
            cols[name].append(take(dt, n).copy())
        t += 1
    for name, _, _ in fields:
        out[name] = np.stack(cols[name]) if cols[name] else np.zeros((0,))
    out["T"] = t
    out["cqi_rb"] = out["cqi_rb"].reshape(t, U, R)
    out["now"] = out["now"].reshape(t)
    out["nvs_slice"] = out["nvs_slice"].reshape(t)
    return out


def compact(rec: dict) -> dict:
    """Reduce per-RB CQI to per-RBG when every RBG is constant (true for all shipped traces)."""
    T, U, R, rbg, G = rec["T"], rec["U"], rec["R"], rec["rbg_size"], rec["G"]
    c = rec["cqi_rb"].reshape(T, U, G, rbg)
    out = {k: v for k, v in rec.items() if k != "cqi_rb"}
    if (c == c[..., :1]).all():
        out["cqi"] = np.ascontiguousarray(c[..., 0])
        out["cqi_per_rb"] = 0
    else:
        out["cqi"] = rec["cqi_rb"]
        out["cqi_per_rb"] = 1
    # dt the EWMA saw: Now - lastUpdate (identical for every bearer in backlogged runs)
    dt = rec["now"][:, None] - rec["last_update"]
    if not (dt == dt[:, :1]).all():
        raise ValueError("bearers disagree on lastUpdate")
    out["dt"] = np.ascontiguousarray(dt[:, 0])
    del out["last_update"]
    return out


def save_npz(path: str, rec: dict) -> None:
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in rec.items()})


def load_npz(path: str) -> dict:
    z = np.load(path)
    out = {}
    for k in z.files:
        v = z[k]
        out[k] = v.item() if v.ndim == 0 else v
    return out
