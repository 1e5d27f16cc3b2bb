"""B200-native column-radiation engine for the RRTMGP.jl `update_fluxes!` hot path.

Host-side mirror of the reference's Layer-2 interface (`RRTMGPSolver`, `update_fluxes!`,
getters; `solver.py`) over a C-ABI CUDA library (`csrc/`, `include/rrtmgp_b200.h`)."""
import importlib as _importlib

from . import hdf5min, lutpack, synthetic, tables  # noqa: F401
from ._lib import RRTMGPB200Error, build_ext  # noqa: F401


def __getattr__(name):
    # solver.py imports torch; keep `import rrtmgp_b200` light for pure-host uses (LUT packs)
    if name.startswith("__") or name == "solver":
        raise AttributeError(name)
    _solver = _importlib.import_module(__name__ + ".solver")
    try:
        return getattr(_solver, name)
    except AttributeError:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}") from None
