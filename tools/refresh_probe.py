"""Development probe: device-resident throughput of the headline cell with CQI refreshed every TTI and every 40 TTIs, in the
u8 and 4-bit layouts, over four consecutive calls that share ONE continuing TTI clock (restarting the clock per call
feeds every call a 7e-17 s first TTI: a burst in the EWMA rates, 33 instead of 48 UEs served per TTI, 8 % slower)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radiosaber_b200 import sched, workload
S, B, T = 20, 4096, 80
u2s = np.repeat(np.arange(S), 5).astype(np.int32); U, G = 100, 64
w, p = np.full(S, 0.05), np.tile(np.array([0, 0, 1, 1], np.int32), (S, 1))
dev = torch.device("cuda:0")
_, dts_all = workload.tti_clock(4 * T)
for layout, refresh in ((0, 1), (2, 1), (0, 40), (2, 40)):
    row = G if layout == 0 else G // 2
    g = sched.Scheduler(9, w, p, u2s, B, cqi_per_rb=layout)
    n_slabs = -(-T // refresh)
    d_cqi = torch.empty((n_slabs, B, U, row), dtype=torch.uint8, device=dev)
    g.synth_cqi(1, 0, 0, n_slabs, d_cqi.data_ptr())
    d_r2 = torch.empty((T, B, 2), dtype=torch.int32, device=dev)
    g.synth_rand2(1, 0, 0, T, d_r2.data_ptr())
    d_rbg = torch.empty((T, B, G), dtype=torch.int16, device=dev); d_bits = torch.empty((T, B, U), dtype=torch.int32, device=dev)
    d_mcs = torch.empty((T, B, U), dtype=torch.uint8, device=dev)
    dout = {"rbg_to_ue": d_rbg.data_ptr(), "tbs_bits": d_bits.data_ptr(), "mcs": d_mcs.data_ptr()}
    g.sync()
    for rep in range(4):
        dts = dts_all[rep * T:(rep + 1) * T]
        t0 = time.perf_counter()
        g.run_device(T, d_cqi.data_ptr(), B * U * row, d_r2.data_ptr(), dts, dout, ttis_per_launch=16, cqi_refresh=refresh)
        g.sync()
        dtm = time.perf_counter() - t0
        served = int((d_bits[-1] > 0).sum().item())
        print(f"layout {layout} smem {g.smem_bytes} refresh {refresh:2d} rep {rep}: {B * T / dtm / 1e6:6.2f} M cell-TTIs/s; UEs served in last TTI per cell {served / B:.1f}", flush=True)
    g.close()
