"""How far into the sorted (rbg, slice) list does MaximizeCell's first-fit scan look?  

Runs the CPU oracle on the headline workload in steady state and prints the histogram of the position of the
last accepted entry, in 64ths of the list length (oracle diagnostic rso_diag_*).  CPU only.
"""
import argparse, ctypes as C, json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle
from radiosaber_b200 import workload

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=64)
ap.add_argument("--ttis", type=int, default=600)
ap.add_argument("--skip", type=int, default=240)
ap.add_argument("--slices", type=int, default=20)
ap.add_argument("--ues", type=int, default=5)
ap.add_argument("--mix", action="store_true")
a = ap.parse_args()
S, B, T = a.slices, a.cells, a.ttis
u2s = np.repeat(np.arange(S), a.ues).astype(np.int32)
U, G = len(u2s), 64
w = np.full(S, 1.0 / S)
p = np.tile(np.array([0, 0, 1, 1], dtype=np.int32), (S, 1))
if a.mix:
    p[1::2] = [0, 0, 1, 0]
o = pyoracle.OracleScheduler(9, w, p, u2s, B, n_threads=8)
L = pyoracle.lib()
_, dts = workload.tti_clock(T)
for t in range(T):
    if t == a.skip:
        L.rso_diag_enable(1)
    cqi = workload.synth_cqi(1, 0, B, t, 1, U, G)[0]
    r2 = workload.synth_rand2(1, 0, B, t, 1, S)[0]
    o.step(cqi, r2, dt=float(dts[t]))
h = np.zeros(66, dtype=np.uint64)
L.rso_diag_stop_hist(h.ctypes.data_as(C.c_void_p))
L.rso_diag_enable(0)
tot = int(h.sum())
cum = np.cumsum(h) / tot
print(json.dumps({"cells": B, "ttis": T - a.skip, "samples": tot,
                  "median_64ths": int(np.searchsorted(cum, 0.5)), "p90_64ths": int(np.searchsorted(cum, 0.9)),
                  "p99_64ths": int(np.searchsorted(cum, 0.99)), "max_64ths": int(np.nonzero(h)[0].max()),
                  "hist": h.tolist()}))
