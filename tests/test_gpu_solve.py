"""On-device window deconvolution: X = N \\ B by blocked LU with partial pivoting on the GPU, against
numpy's LAPACK solve, and the reference's own identities for the binned matrices (test/test_windows.jl:570-586)."""
import time
import warnings

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,nrhs", [(1, 1), (7, 3), (64, 1), (65, 2), (200, 5), (517, 130), (1500, 1)])
def test_solve_matches_lapack(n, nrhs):
    import sfb_b200 as sfb
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))                       # generic matrix: pivoting matters
    B = rng.standard_normal((n, nrhs))
    X = sfb.solve(A, B)
    ref = np.linalg.solve(A, B)
    assert X.shape == ref.shape
    assert relerr(X, ref) < 1e-9 * max(1.0, np.linalg.cond(A) / 1e4)
    assert relerr(A @ X, B) < 1e-10 * max(1.0, np.linalg.cond(A) / 1e3)
    x1 = sfb.solve(A, B[:, 0])
    assert x1.shape == (n,) and relerr(x1, ref[:, 0]) < 1e-9 * max(1.0, np.linalg.cond(A) / 1e4)


def test_solve_needs_pivoting_and_flags_singular():
    import sfb_b200 as sfb
    from sfb_b200 import _lib
    A = np.array([[0.0, 2.0, 1.0], [1.0, 1.0, 0.0], [3.0, 0.0, 1.0]])       # zero leading pivot
    b = np.array([1.0, 2.0, 3.0])
    assert np.allclose(sfb.solve(A, b), np.linalg.solve(A, b), rtol=1e-13)
    with pytest.raises(_lib.SFBError, match="SingularException"):
        sfb.solve(np.ones((4, 4)), np.ones(4))


def test_binned_deconvolution_identities():
    # test/test_windows.jl:570-586: N = w̃ M v, w = inv(N) w̃ M, w v ≈ I; inv(Nmix) * w̃M ≈ w at rtol 1e-10
    import sfb_b200 as sfb
    a = sfb.AnlmModes(0.03, 500.0, 1000.0)
    c = sfb.ClnnModes(a)
    wm = sfb.ConfigurationSpaceModes(a, 40)
    rng = np.random.default_rng(3)
    mask = (rng.random(wm.npix) > 0.4).astype(float)
    win = np.outer(np.exp(-(wm.r / 550.0) ** 2), mask)
    wt, v = sfb.bandpower_binning_weights(c, dl=3)
    bc = sfb.ClnnBinnedModes(wt, v, c)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        M = sfb.power_win_mix(win, wm, c)
        N = sfb.power_win_mix(win, wt, v, wm, bc)
        wM = sfb.power_win_mix(win, wt, None, wm, bc)
        cobs = rng.random(M.shape[0])
        X, N2 = sfb.power_win_mix_solve(win, wt, v, wm, bc, wt @ cobs, return_N=True)    # C = bcmix \ (w̃ Cobs)
        W = sfb.power_win_mix_solve(win, wt, v, wm, bc, wM)                            # w = inv(N) w̃M, on the device
    assert relerr(N2, N) < 1e-14
    assert relerr(X, np.linalg.solve(N, wt @ cobs)) < 1e-10
    w_ref = np.linalg.inv(N) @ (wt @ M)
    assert relerr(W, w_ref) < 1e-10
    assert np.allclose(W @ v, np.eye(N.shape[0]), atol=1e-10)                          # w v ≈ I


def test_cfg4_binned_deconvolution_on_device():
    # cfg4's named output (Δl = 4, LNN = 5545): the whole chain window -> N -> C = N \ (w̃ Cobs) with only vectors leaving
    import sfb_b200 as sfb
    from sfb_b200 import configs
    wl = configs.Workload(4)
    wt, vv = sfb.bandpower_binning_weights(wl.cmodes, dl=4)
    bc = sfb.ClnnBinnedModes(wt, vv, wl.cmodes)
    rng = np.random.default_rng(0)
    rhs = wt @ rng.random(wl.lnnsize)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sfb.power_win_mix_solve(wl.win, wt, vv, wl.wmodes, bc, rhs)                    # warm-up (plans, workspaces)
        t0 = time.perf_counter()
        X, N = sfb.power_win_mix_solve(wl.win, wt, vv, wl.wmodes, bc, rhs, return_N=True)
        dt = time.perf_counter() - t0
    assert relerr(N @ X, rhs) < 1e-10
    t0 = time.perf_counter()
    ref = np.linalg.solve(N, rhs)
    dt_cpu = time.perf_counter() - t0
    assert relerr(X, ref) < 1e-9
    print(f"cfg4 binned deconvolution LNN={N.shape[0]}: GPU chain {dt * 1e3:.1f} ms (incl. H2D, stage 1-3, N out), "
          f"LAPACK solve alone on the host {dt_cpu * 1e3:.1f} ms")
