"""SFB_TRACE=1 python tools/trace_md.py NDEV: phase timings of the multi-device host path (cfg4, page-locked buffers)."""
import sys
import time
import warnings

sys.path.insert(0, ".")
import numpy as np

import sfb_b200 as sfb
from sfb_b200 import _lib, configs

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl = configs.Workload(4)
n = wl.lnnsize
warnings.simplefilter("ignore")
sfb.set_devices(ndev)
win = sfb.pinned_empty(wl.win.shape)
win[...] = wl.win
out = sfb.pinned_empty((n, n))
wt, vv = sfb.bandpower_binning_weights(wl.cmodes, dl=4)
bc = sfb.ClnnBinnedModes(wt, vv, wl.cmodes)
outN = sfb.pinned_empty((wt.shape[0], vv.shape[1]))
for rep in range(3):
    print(f"---- unbinned call {rep}", file=sys.stderr)
    t0 = time.perf_counter()
    sfb.power_win_mix(win, wl.wmodes, wl.cmodes, out=out)
    print(f"python wall {1e3 * (time.perf_counter() - t0):.2f} ms", _lib.timings(), file=sys.stderr)
for rep in range(3):
    print(f"---- binned call {rep}", file=sys.stderr)
    t0 = time.perf_counter()
    sfb.power_win_mix(win, wt, vv, wl.wmodes, bc, out=outN)
    print(f"python wall {1e3 * (time.perf_counter() - t0):.2f} ms", _lib.timings(), file=sys.stderr)
sfb.set_devices(1)
t0 = time.perf_counter()
N1 = sfb.power_win_mix(win, wt, vv, wl.wmodes, bc)
print("1-device binned equals:", float(np.abs(N1 - outN).max()), file=sys.stderr)
