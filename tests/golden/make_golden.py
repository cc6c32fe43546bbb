#!/usr/bin/env python
"""Regenerates tests/golden/small_case.npz from the CPU oracle.

The reference is Julia and cannot run in this image (no julia binary), so these vectors are NOT outputs of the
reference itself: they freeze the oracle's answers on a seeded small case so that (a) a later edit to the oracle
that changes its numbers is caught on CPU, and (b) the CUDA path is also checked against committed numbers.
The reference's own known answers for this path (full sky => M = I, analytic l=0 block, sep == insep, ...) are
asserted directly in tests/test_oracle_*.py.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import modes as om  # noqa: E402
from oracle import windows as ow  # noqa: E402


def build():
    a = om.AnlmModes(0.022, 500.0, 1000.0)       # lmax 18, ragged nmax_l, nside 32
    wm = ow.ConfigurationSpaceModes(500.0, 1000.0, 20, 8)   # nside-8 window: exercises udgrade 8 -> 32
    c = om.ClnnModes(a)
    rng = np.random.default_rng(20240517)
    win = rng.random((wm.nr, wm.npix)) * np.exp(-(wm.r / 550.0) ** 2)[:, None]
    win[:, rng.random(wm.npix) < 0.3] = 0.0
    win /= win.max()
    LMAX = 2 * a.lmax
    Wr_lm = ow.calc_Wr_lm(win, LMAX, a.nside)
    M = ow.power_win_mix(win, win, wm, c)
    Md = ow.power_win_mix(win, win, wm, c, div2Lp1=True, interchange=True, lnn_min=7)
    wt, v = om.bandpower_binning_weights(c, dl=3)
    Wlnn = ow.win_lnn(win, wm, c)
    return dict(Wlnn=Wlnn, kmax=0.022, nr=20, win_nside=8, win=win, lnn=c.lnn, nmax_l=a.nmax_l, knl=a.knl, Wr_lm=Wr_lm, M=M,
                M_div_interchange_min7=Md, N_binned_dl3=wt @ M @ v)


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_case.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, os.path.getsize(out), "bytes")
