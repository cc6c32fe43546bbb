"""CPU oracle for the SuperFaB window coupling-matrix path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`sphericalfourierbesseldecompositions.jl_b200/`) imports this.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import, call, link or execute anything in here, and there only
as the checker / baseline.

Each function restates the algorithm of hsgg/SphericalFourierBesselDecompositions.jl
v0.5.19 (Julia) and cites the reference file:line it follows (paths relative to
the reference checkout).  The reference itself cannot run here (no Julia in the
image), and the spherical-harmonic transform it calls lives in un-vendored
dependencies (Healpix.jl 4.2.2 -> Libsharp.jl 0.2.0 -> libsharp2_jll 1.0.2+2,
pinned in the reference's Manifest.toml), so:

* mode / index tables, the radial basis, `calc_Wrl_Wrl`, `wigner3j000`,
  `calc_cmix`, the binned and separable variants are restated line by line from
  the in-tree Julia source and pinned by the reference's own identities
  (tests/test_oracle_*.py);
* `map2alm(niter=3)` / `udgrade` / `alm2cl` are restated from the published
  HEALPix / libsharp definitions.  The reference's tests pin this step only to
  1e-3..1e-6 (test/test_windows.jl:36,213,300), so at the 1e-10 level the SHT
  boundary is **parity unpinned** (see DESIGN.md).
"""
