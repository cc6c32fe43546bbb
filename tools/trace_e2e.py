import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfb_b200 as sfb
from sfb_b200 import configs
wl = configs.Workload(4); n = wl.lnnsize; win = wl.win
nr, npix = win.shape
hw = torch.empty((npix, nr), dtype=torch.float64).pin_memory(); hw.copy_(torch.from_numpy(np.ascontiguousarray(win.T))); host_win = hw.numpy().T
out_t = torch.empty((n, n), dtype=torch.float64).pin_memory(); out = out_t.numpy().T
wt, vv = sfb.bandpower_binning_weights(wl.cmodes, dl=4); bc = sfb.ClnnBinnedModes(wt, vv, wl.cmodes)
outN_t = torch.empty((vv.shape[1], wt.shape[0]), dtype=torch.float64).pin_memory(); outN = outN_t.numpy().T
warnings.simplefilter("ignore")
for rep in range(3):
    t0 = time.perf_counter(); sfb.power_win_mix(host_win, wl.wmodes, wl.cmodes, out=out); print("unbinned call %.1f ms" % (1e3*(time.perf_counter()-t0)), file=sys.stderr)
    t0 = time.perf_counter(); sfb.power_win_mix(host_win, wt, vv, wl.wmodes, bc, out=outN); print("binned call %.1f ms" % (1e3*(time.perf_counter()-t0)), file=sys.stderr)
