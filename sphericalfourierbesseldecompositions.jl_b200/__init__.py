"""B200-native (sm_100a) drop-in for the window coupling-matrix hot path of
hsgg/SphericalFourierBesselDecompositions.jl: `calc_Wr_lm` + `power_win_mix`.

Import as `sfb_b200` (see ../sfb_b200.py; the directory name carries the reference's name and is not a
valid Python identifier).  Layout:

  csrc/        CUDA kernels (stage 1 SHT, stage 2+3 coupling matrix, binning) and the C ABI (include/sfb_b200.h)
  _lib.py      ctypes binding of the C ABI; fails loudly when the library is missing (no CPU fallback)
  modes.py     AnlmModes / ClnnModes / ClnnBinnedModes mirrors (inputs of the path)
  windows.py   ConfigurationSpaceModes, calc_Wr_lm, power_win_mix mirrors (the path itself)
  separable.py SeparableArray
  device.py    device-resident pipeline + row sharding over torch.distributed (multi-GPU, bench)
"""
from . import _lib
from .modes import (AnlmModes, ClnnBinnedModes, ClnnModes, bandpower_binning_weights, estimate_nside, getidx, getlkk,
                    getlmsize, getlnn, getlnnsize, getnlm, getnlmsize)
from .separable import SeparableArray
from .cat2anlm import amln2clnn, anlm2field, cat2amln, field2anlm, win_rhat_ln
from .windows import (ConfigurationSpaceModes, calc_Wr_lm, calc_wmix, calc_wmix_all, check_nsamp, get_devices, optimize_Wr_lm_layout, pinned_empty,
                      power_win_mix, power_win_mix_from_wrlm, power_win_mix_solve, solve, precompute_gnlr, rsdrgnlr, set_devices, win_lnn, window_r)

__all__ = [
    "AnlmModes", "ClnnModes", "ClnnBinnedModes", "bandpower_binning_weights", "estimate_nside", "getidx", "getlkk",
    "getlmsize", "getlnn", "getlnnsize", "getnlm", "getnlmsize", "SeparableArray", "ConfigurationSpaceModes",
    "calc_Wr_lm", "check_nsamp", "optimize_Wr_lm_layout", "power_win_mix", "power_win_mix_from_wrlm",
    "precompute_gnlr", "rsdrgnlr", "win_lnn", "window_r", "set_devices", "get_devices", "pinned_empty", "calc_wmix", "calc_wmix_all",
    "solve", "power_win_mix_solve", "field2anlm", "anlm2field", "win_rhat_ln", "cat2amln", "amln2clnn",
]
