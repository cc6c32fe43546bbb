"""End-to-end power_win_mix at cfg4 with pageable host arrays (plain numpy): the library's staged copies against the driver's
(SFB_NO_STAGING=1), for a few host-thread counts.  Usage: python tools/try_pageable.py"""
import os, subprocess, sys
code = r'''
import os, sys, time, warnings
sys.path.insert(0, os.getcwd())
import numpy as np
import sfb_b200 as sfb
from sfb_b200 import configs
wl = configs.Workload(4); n = wl.lnnsize
out = np.empty((n, n), order="F")
warnings.simplefilter("ignore")
for rep in range(4):
    t0 = time.perf_counter(); sfb.power_win_mix(wl.win, wl.wmodes, wl.cmodes, out=out); dt = time.perf_counter() - t0
print("%.1f ms  checksum %.15g" % (dt * 1e3, out[::n // 97, ::n // 89].sum()))
'''
for env in ({"SFB_NO_STAGING": "1"}, {"SFB_COPY_THREADS": "4"}, {"SFB_COPY_THREADS": "8"}, {"SFB_COPY_THREADS": "16"},
            {"SFB_COPY_THREADS": "16", "SFB_NO_NT_COPY": "1"}, {"SFB_COPY_THREADS": "32"}):
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True)
    print(env, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
