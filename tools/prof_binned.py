import os, sys, time, warnings, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfb_b200 as sfb
from sfb_b200 import configs
wl = configs.Workload(4); n = wl.lnnsize; win = wl.win
nr, npix = win.shape
hw = torch.empty((npix, nr), dtype=torch.float64).pin_memory(); hw.copy_(torch.from_numpy(np.ascontiguousarray(win.T))); host_win = hw.numpy().T
wt, vv = sfb.bandpower_binning_weights(wl.cmodes, dl=4); bc = sfb.ClnnBinnedModes(wt, vv, wl.cmodes)
outN_t = torch.empty((vv.shape[1], wt.shape[0]), dtype=torch.float64).pin_memory(); outN = outN_t.numpy().T
with warnings.catch_warnings():
    warnings.simplefilter("ignore", RuntimeWarning)
    sfb.power_win_mix(host_win, wt, vv, wl.wmodes, bc, out=outN)
    pr = cProfile.Profile(); pr.enable()
    t0 = time.perf_counter()
    for _ in range(3):
        sfb.power_win_mix(host_win, wt, vv, wl.wmodes, bc, out=outN)
    print("per call ms", (time.perf_counter() - t0) / 3 * 1e3)
    pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
