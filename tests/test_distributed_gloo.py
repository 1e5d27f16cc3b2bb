"""N > 1 host logic on CPU: world_size-2 gloo run of the column-shard + all-gather plumbing
(rrtmgp.jl_b200/sharding.py).  The per-shard compute is the CPU oracle here (no GPU in this
container); on the GPU box bench.py drives the same plumbing with NCCL and the CUDA engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rrtmgp_b200 as R
from rrtmgp_b200.sharding import FLUX_KEYS, all_gather_fluxes, shard_range, shard_state

_KEYMAP = {"lw_flux_up": "lw_up", "lw_flux_dn": "lw_dn", "lw_flux_net": "lw_net", "sw_flux_up": "sw_up",
           "sw_flux_dn": "sw_dn", "sw_flux_net": "sw_net", "sw_flux_dn_dir": "sw_dir", "net_flux": "net"}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ncol, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import Oracle
        pack = R.synthetic.make_lut_pack(seed=11, dims=R.synthetic.SMALL_DIMS)
        st = R.synthetic.make_atmosphere(ncol, 16, dtype=np.float64, n_bnd_lw=3, n_bnd_sw=3, cld_frac=None)
        lo, hi = shard_range(ncol, rank, world)
        r = Oracle(pack, np.float64).update_fluxes(shard_state(st, ncol, rank, world), seed=77, col_offset=lo, nthreads=1)
        local = {k: torch.from_numpy(r[_KEYMAP[k]]) for k in FLUX_KEYS}
        g = all_gather_fluxes(local)
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"), **{k: v.numpy() for k, v in g.items()})
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n, w in ((100000, 8), (10, 3), (7, 8), (64, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def test_two_rank_gather_equals_single_process(tmp_path):
    ncol, world = 24, 2
    mp.spawn(_worker, args=(world, _free_port(), ncol, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npz")
    from oracle import Oracle
    pack = R.synthetic.make_lut_pack(seed=11, dims=R.synthetic.SMALL_DIMS)
    st = R.synthetic.make_atmosphere(ncol, 16, dtype=np.float64, n_bnd_lw=3, n_bnd_sw=3, cld_frac=None)
    ref = Oracle(pack, np.float64).update_fluxes(st, seed=77, nthreads=1)
    for k in FLUX_KEYS:
        np.testing.assert_array_equal(got[k], ref[_KEYMAP[k]])


def _id_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        from rrtmgp_b200 import _lib
        box = [None]
        if rank == 0:
            buf = (C.c_char * _lib.UNIQUE_ID_BYTES)()
            st = _lib.lib().rrtmgp_b200_comm_unique_id(buf, _lib.UNIQUE_ID_BYTES)
            box = [(st, bytes(buf))]
        dist.broadcast_object_list(box, src=0)      # how bench.py ships the id to every rank
        with open(os.path.join(out_dir, f"id{rank}.bin"), "wb") as f:
            f.write(bytes([box[0][0]]) + box[0][1])
    finally:
        dist.destroy_process_group()


def test_comm_unique_id_reaches_every_rank(tmp_path):
    """`rrtmgp_b200_comm_unique_id` (rank 0) -> broadcast -> every rank holds the same 128 bytes to hand to
    `rrtmgp_b200_comm_init`; the gathered layout is rank r -> rows [r ncol, (r + 1) ncol) (shard_range with equal
    shards).  NCCL itself needs GPUs; without libnccl the entry point answers UNSUPPORTED."""
    from rrtmgp_b200 import _lib
    world = 2
    mp.spawn(_id_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (open(tmp_path / f"id{r}.bin", "rb").read() for r in range(world))
    assert a == b and len(a) == 1 + _lib.UNIQUE_ID_BYTES
    assert a[0] in (_lib.OK, _lib.ERR_UNSUPPORTED)
    if a[0] == _lib.OK:
        assert any(a[1:])                            # a real id, not zeros
    for r in range(4):
        assert shard_range(4 * 1000, r, 4) == (r * 1000, (r + 1) * 1000)
