#!/usr/bin/env python3
"""BASELINE.json configs[2] and configs[3] on one GPU (device-resident inputs, CUDA events):

  configs[2]: the 4096-cell 20 x 5 batch under ids 1 / 7 / 8 / 9 (and 10 UpperBound, 11 NVS non-greedy) with PF enterprise schedulers and with the
              Sec 6.2-style mix (per-slice parameters cycling PF (eps 1, psi 1) / MT == max-CI (eps 1, psi 0))
  configs[3]: slices x UEs-per-slice sweep, weights proportional to 1 + (s mod 3), RadioSaber (9), the batch
              size chosen so that cells x UEs ~ 409 600

One JSON line per point on stdout; `--out FILE` also writes them to FILE.  Not the bench contract (bench.py
is); these are the parity-test shapes measured for DESIGN.md / profiles/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiosaber_b200 import sched, workload  # noqa: E402

G = 64


def measure(algo, w, p, u2s, B, ttis, launches, label, warm=5, per_launch=16, layout=0):
    import torch
    dev = torch.device("cuda", 0)
    S, U = len(w), len(u2s)
    g = sched.Scheduler(algo, w, p, u2s, B, cqi_per_rb=layout)
    row = G // 2 if layout == 2 else G   # bytes of CQI per UE: one value per RBG, or two RBGs per byte
    stream = torch.cuda.Stream(dev)   # a real stream: handle 0 (the default stream) means "the handle's own" to rs_set_stream
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)
    d_cqi = torch.empty((ttis, B, U, row), dtype=torch.uint8, device=dev)
    d_r2 = torch.empty((ttis, B, max(g.rand_stride, 2)), dtype=torch.int32, device=dev)
    g.synth_cqi(1, 0, 0, ttis, d_cqi.data_ptr())
    g.synth_rand2(1, 0, 0, ttis, d_r2.data_ptr())
    d_rbg = torch.empty((ttis, B, G), dtype=torch.int16, device=dev)
    d_bits = torch.empty((ttis, B, U), dtype=torch.int32, device=dev)
    outs = {"rbg_to_ue": d_rbg.data_ptr(), "tbs_bits": d_bits.data_ptr()}
    _, dts = workload.tti_clock(ttis * (launches + warm))

    def step(k):
        g.run_device(ttis, d_cqi.data_ptr(), B * U * row, d_r2.data_ptr(), dts[k * ttis:(k + 1) * ttis], outs,
                     ttis_per_launch=per_launch)

    for k in range(warm):     # past the start-up transient (all bearers begin at the same average rate): steady state
        step(k)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(launches):
        step(warm + k)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    alg = g.algorithmic_bytes_per_cell_tti
    value = B * ttis * launches / (ms * 1e-3)
    rbg = d_rbg.cpu().numpy()
    line = {"label": label, "scheduler_id": algo, "slices": S, "ues_per_slice": U // S, "ues": U, "cells": B,
            "ttis_per_call": ttis, "ttis_per_launch": per_launch, "calls": launches, "warmup_ttis": ttis * warm, "cell_ttis_per_s": value, "ue_ttis_per_s": value * U,
            "smem_bytes_per_cta": g.smem_bytes, "cqi_layout": layout, "algorithmic_bytes_per_cell_tti": alg,
            "algorithmic_GBps": value * alg / 1e9, "rbgs_allocated_frac": float((rbg >= 0).mean())}
    g.close()
    del d_cqi, d_r2, d_rbg, d_bits
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--ttis", type=int, default=48, help="TTIs per call (three 16-TTI launches: part-batches overlap inside a call)")
    ap.add_argument("--launches", type=int, default=3, help="timed calls")
    ap.add_argument("--only", default=None, choices=[None, "ids", "sweep"])
    ap.add_argument("--points", default=None, help='sweep points "S,n;S,n;..." instead of the full grid')
    ap.add_argument("--ids", default=None, help='scheduler ids "9,8,..." instead of all')
    ap.add_argument("--layout", type=int, default=0, choices=[0, 2], help="CQI layout: 0 = u8 per RBG, 2 = two RBGs per byte")
    args = ap.parse_args()
    lines = []

    def emit(line):
        lines.append(line)
        print(json.dumps(line), flush=True)

    if args.only in (None, "ids"):
        S, n = 20, 5
        u2s = np.repeat(np.arange(S), n).astype(np.int32)
        w = np.full(S, 1.0 / S)
        pf = np.tile(np.array([0, 0, 1, 1], dtype=np.int32), (S, 1))
        mix = pf.copy()
        mix[1::2, 3] = 0      # every other slice MT (max-CI): eps 1, psi 0
        for algo in ([int(x) for x in args.ids.split(",")] if args.ids else (9, 8, 7, 1, 11, 10, 101, 103)):
            emit(measure(algo, w, pf, u2s, 4096, args.ttis, args.launches, f"configs[2] id {algo} PF", layout=args.layout))
            if algo not in (1, 11, 10, 101, 103):
                emit(measure(algo, w, mix, u2s, 4096, args.ttis, args.launches, f"configs[2] id {algo} PF/MT mix"))
    if args.only in (None, "sweep"):
        grid = [(S, n) for S in (5, 10, 15, 20, 30, 40, 50) for n in (2, 5, 10, 15, 20, 30, 40)]
        if args.points:
            grid = [tuple(int(x) for x in pt.split(",")) for pt in args.points.split(";")]
        for S, n in grid:
            if True:
                U = S * n
                B = max(148, int(round(409600 / U)))
                w = 1.0 + (np.arange(S) % 3)
                w = w / w.sum()
                p = np.tile(np.array([0, 0, 1, 1], dtype=np.int32), (S, 1))
                u2s = np.repeat(np.arange(S), n).astype(np.int32)
                emit(measure(9, w, p, u2s, B, args.ttis, args.launches, f"configs[3] {S} slices x {n} UEs"))
    if args.out:
        with open(args.out, "w") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
