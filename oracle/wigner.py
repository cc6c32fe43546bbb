"""TEST INFRASTRUCTURE (see oracle/__init__.py): Wigner 3j families by three-term recursion.

The reference obtains all (j1 j2 j3; m1 m2 m3), j1 = jmin..jmax, from WignerFamilies 1.0.2 (`WignerF`, `wigner3j_f!`,
call sites src/windows.jl:434-464, src/window_chains.jl:530-551).  That package is not vendored under /root/reference
(Manifest.toml:990-994); it implements the Schulten-Gordon / Luscombe-Luban recursion restated here from the published
algorithm (K. Schulten, R. G. Gordon, J. Math. Phys. 16 (1975) 1961; J. H. Luscombe, M. Luban, Phys. Rev. E 57 (1998)
7274):

    j A(j+1) f(j+1) + B(j) f(j) + (j+1) A(j) f(j-1) = 0,      f(j) = (j j2 j3; m1 m2 m3),  m1 = -m2-m3
    A(j) = sqrt[(j² - (j2-j3)²) ((j2+j3+1)² - j²) (j² - m1²)]
    B(j) = -(2j+1) [j2(j2+1) m1 - j3(j3+1) m1 - j(j+1)(m3-m2)]
    Σ_j (2j+1) f(j)² = 1,    sign f(jmax) = (-1)^(j2-j3-m1)

run forward from jmin and backward from jmax and matched where both are stable.  Pinned against sympy's exact
`wigner_3j` in tests/test_oracle_windows.py.  Integer angular momenta only (all the path needs).
"""
from __future__ import annotations

import math

import numpy as np


def _A(j, j2, j3, m1):
    return math.sqrt(max(0.0, (j * j - (j2 - j3) ** 2) * ((j2 + j3 + 1) ** 2 - j * j) * (j * j - m1 * m1)))


def _B(j, j2, j3, m1, m2, m3):
    return -(2 * j + 1) * (j2 * (j2 + 1) * m1 - j3 * (j3 + 1) * m1 - j * (j + 1) * (m3 - m2))


def wigner3j_family(j2, j3, m2, m3):
    """(jmin, f) with f[k] = (jmin+k  j2  j3; -m2-m3  m2  m3), like WignerFamilies.wigner3j_f(j2, j3, m2, m3).
    Returns (jmin, empty array) when |m2| > j2 or |m3| > j3."""
    m1 = -m2 - m3
    if abs(m2) > j2 or abs(m3) > j3:
        return 0, np.zeros(0)
    jmin, jmax = max(abs(j2 - j3), abs(m1)), j2 + j3
    n = jmax - jmin + 1
    if n == 1:
        f = np.array([1.0 / math.sqrt(2 * jmin + 1)])
    else:
        # forward from jmin (for jmin = 0 the first step degenerates to 0 = 0: the backward run alone is used)
        fw = np.full(n, np.nan)
        if jmin > 0:
            fw[0] = 1.0
            for k in range(n - 1):
                j = jmin + k
                prev = fw[k - 1] if k > 0 else 0.0
                fw[k + 1] = -(_B(j, j2, j3, m1, m2, m3) * fw[k] + (j + 1) * _A(j, j2, j3, m1) * prev) / (
                    j * _A(j + 1, j2, j3, m1))
        # backward from jmax
        bw = np.zeros(n)
        bw[-1] = 1.0
        for k in range(n - 1, 0, -1):
            j = jmin + k
            nxt = bw[k + 1] if k + 1 < n else 0.0
            bw[k - 1] = -(_B(j, j2, j3, m1, m2, m3) * bw[k] + j * _A(j + 1, j2, j3, m1) * nxt) / ((j + 1) * _A(j, j2, j3, m1))
        if not np.all(np.isfinite(fw)):
            f = bw
        else:
            # match at the index where both runs are largest (classically allowed region)
            score = np.abs(fw) / np.max(np.abs(fw)) * np.abs(bw) / np.max(np.abs(bw))
            km = int(np.argmax(score))
            f = np.concatenate([fw[:km] * (bw[km] / fw[km]), bw[km:]])
        js = jmin + np.arange(n)
        f = f / math.sqrt(np.sum((2 * js + 1) * f * f))
    sign = -1.0 if (j2 - j3 - m1) % 2 else 1.0
    if f[-1] * sign < 0:
        f = -f
    return jmin, f


def wigner3j(j1, j2, j3, m1, m2, m3):
    if m1 + m2 + m3 != 0:
        return 0.0
    jmin, f = wigner3j_family(j2, j3, m2, m3)
    k = j1 - jmin
    return float(f[k]) if 0 <= k < f.size else 0.0
