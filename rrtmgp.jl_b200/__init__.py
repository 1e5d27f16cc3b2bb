"""B200-native column-radiation engine for the RRTMGP.jl `update_fluxes!` hot path.

Host-side mirror of the reference's Layer-2 interface (`RRTMGPSolver`, `update_fluxes!`,
getters) over a C-ABI CUDA library (`csrc/`, `include/rrtmgp_b200.h`)."""
from . import lutpack, synthetic  # noqa: F401
