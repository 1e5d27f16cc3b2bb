"""Column sharding across the GPUs of one box.

Columns are independent (every reference kernel indexes a single `gcol`), `ncol` is the slowest
axis of every input and output array (`AtmosphericStates.jl:60-66`, `BCs.jl:14-15,34-38`,
`Fluxes.jl:355-374`), so a rank's shard is one contiguous block and the gathered `(nlev, ncol)`
flux views are a plain concatenation: one `all_gather_into_tensor` per flux array (NCCL over
NVLink on GPUs, gloo in the CPU tests).  The reference has no multi-GPU path to compare with.
The McICA draws are keyed by the GLOBAL column index (`col_offset`), so results do not depend
on how columns are sharded.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

FLUX_KEYS = ("lw_flux_up", "lw_flux_dn", "lw_flux_net", "sw_flux_up", "sw_flux_dn", "sw_flux_net",
             "sw_flux_dn_dir", "net_flux")


def shard_range(ncol_global: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of `rank`'s contiguous column block; blocks differ by at most one column."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(ncol_global, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_state(state: Dict, ncol_global: int, rank: int, world: int) -> Dict:
    """Slice every per-column array of a host state dict (keys of `synthetic.make_atmosphere`)."""
    a, b = shard_range(ncol_global, rank, world)
    out = {}
    for k, v in state.items():
        if k == "inc_flux_lw":          # (ncol, ngpt) in the reference == [ngpt][ncol] here
            out[k] = v[:, a:b].copy()
        elif getattr(v, "ndim", 0) >= 1 and v.shape[0] == ncol_global:
            out[k] = v[a:b]
        else:
            out[k] = v
    return out


def all_gather_fluxes(local: Dict, keys: Sequence[str] = FLUX_KEYS, group=None) -> Dict:
    """Gather equal-size `[ncol_local, nlev]` flux tensors of every rank into `[ncol_global, nlev]`.

    Equal shard sizes are required (pad `ncol` to a multiple of the world size, as a host would);
    all collectives are issued before any is waited on, so they overlap."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out, works = {}, []
    for k in keys:
        t = local[k]
        g = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        works.append(dist.all_gather_into_tensor(g, t.contiguous(), group=group, async_op=True))
        out[k] = g
    for w in works:
        w.wait()
    return out
