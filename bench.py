#!/usr/bin/env python
"""bench.py -- columns/s for all-sky-with-aerosols `update_fluxes!` (ncol = 1e5 per GPU, nlay = 64, Float32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--ncol C]

One process per GPU (torchrun for N > 1; RANK / LOCAL_RANK / WORLD_SIZE from the env).  A "step" is one
full `update_fluxes!` (prepare_atmosphere! + LW two-stream + SW two-stream + net flux, clouds with McICA,
15-species aerosols) over one batch of synthetic columns (seeded generator).  Columns shard
across ranks (weak scaling: `--ncol` columns PER GPU); for N > 1 the step is `update_fluxes_gathered` of the C ABI:
the all-gather of the eight (nlev, ncol) flux views (north star) pushed over NVLink by the copy
engines while the shortwave kernel still runs, framed by two one-element NCCL all-reduces.

Rank 0 prints ONE JSON line.  `value` = whole-job columns/s with inputs resident in HBM; `e2e` = the same
step driven from pinned HOST buffers (H2D of every input + D2H of every flux inside the timed region);
`roofline` = algorithmic HBM bytes of the dominant kernel / its CUDA-event time vs the measured HBM peak
(this path is FP32/LUT-gather bound, not HBM bound -- see DESIGN.md; the FP32 view is in `roofline_fp32`);
`cpu_baseline` = the restated reference (oracle/, OpenMP, all host cores) on a bounded column sample.
`--impl reference` times that CPU restatement alone (the reference itself is Julia and cannot run here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its version
# banner there at communicator creation on the GPU boxes), so descriptor 1 is pointed at stderr for the whole run and
# the JSON line goes to a private duplicate of the original stdout.
sys.stdout.flush()
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())

METRIC = "columns/sec for all-sky update_fluxes! (ncol=1e5, nlay=64) at 1/2/4/8 B200"
PARAMS = dict(grav=9.80665, molmass_dryair=0.028964, molmass_water=0.018016)
ALGO_FLOPS_PER_COL = 6.8e6      # SURVEY.md §8d (+-30 %)
FP32_NOMINAL_TFLOPS = 74.4      # 148 SM x 128 lanes x 2 x 1.965 GHz (nominal) -- reported, not used as the peak


def workload_name(ncol, nlay):
    return (f"all_sky_with_aerosols: ncol={ncol}/GPU nlay={nlay} Float32, 256 LW + 224 SW g-points, two-stream LW+SW, "
            "cld_frac=1 (McICA), 15-species aerosols, cos_zenith=0.86, synthetic LUT pack seed 7")


def config_of(ncol, nlay, world):
    """The `config` object of both arms (engine and `--impl reference`): same keys, same workload string."""
    in_bytes = ncol * ((nlay * 4 + nlay * 2 + nlay * 5 + 2 * nlay * 15) * 4 + 400)
    return {"workload": workload_name(ncol, nlay), "global_columns": world * ncol,
            "parallelism": f"column shards x{world}" + (", all-gather of 8 flux views (copy-engine pushes over NVLink + NCCL fences, "
                                                        "rrtmgp_b200_update_fluxes_gathered)" if world > 1 else ""),
            "l2_policy": f"inputs {in_bytes / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every 10 ms
    from a thread (the timed region is a fraction of a second), `nvidia-smi -lms` as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    BITS = (0x8, 0x40, 0x20, 0x4)   # nvmlClocksEventReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        try:
            self.index = int(vis.split(",")[index]) if vis else index
        except Exception:
            self.index = index
        self.rows, self.proc, self.t, self.stop = [], None, None, threading.Event()

    def _nvml_loop(self, nv, h):
        while not self.stop.is_set():
            try:
                sm, mx = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for b in self.BITS])
            except Exception:
                pass
            self.stop.wait(0.01)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
            return self
        except Exception:
            self.t = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.t:
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(self.NAMES, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Host threads this process may use (affinity mask, not the machine total)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def make_state(ncol, nlay, rank):
    import rrtmgp_b200 as R
    return R.synthetic.make_atmosphere(ncol, nlay, seed=20260101 + rank, cld_frac=1.0, cos_zenith=0.86)


def run_reference(args, rank, world):
    """`--impl reference`: the restated reference CPU path (oracle/), all host threads, bounded sample."""
    if rank != 0:
        return
    import rrtmgp_b200 as R
    from oracle import Oracle
    pack = R.synthetic.make_lut_pack(seed=7)
    o = Oracle(pack, np.float32)
    cores = host_threads()   # explicit: torchrun exports OMP_NUM_THREADS=1 to its workers
    # bounded sample: a short calibration run sizes it so that warmup + steps take about 90 s of CPU time
    cal = make_state(2048, args.nlay, 0)
    t0 = time.perf_counter()
    o.update_fluxes(cal, seed=1, params=PARAMS, nthreads=cores)
    rate = 2048 / max(time.perf_counter() - t0, 1e-6)
    budget_cols = int(90.0 * rate / max(1, args.steps + args.warmup))
    ncol_s = max(2048, min(args.ncol, args.cpu_sample, budget_cols))
    st = make_state(ncol_s, args.nlay, 0)
    for _ in range(args.warmup):
        o.update_fluxes(st, seed=1, params=PARAMS, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.update_fluxes(st, seed=1, params=PARAMS, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = ncol_s / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "columns/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args.ncol, args.nlay, max(world, 1)),
            "cpu_baseline": {"value": v, "unit": "columns/s", "cores": cores, "kind": "port",
                             "sample": f"{ncol_s} of {args.ncol} columns per step (cost is linear in ncol); "
                                       "C++ restatement of the Julia reference, OpenMP over columns"},
            "e2e": {"value": v, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--ncol", type=int, default=100000, help="columns per GPU")
    ap.add_argument("--nlay", type=int, default=64)
    ap.add_argument("--cpu-sample", type=int, default=32768, help="columns of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import rrtmgp_b200 as R

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ncol, nlay, nlev = args.ncol, args.nlay, args.nlay + 1
    pack = R.synthetic.make_lut_pack(seed=7)
    st = make_state(ncol, nlay, rank)
    gp = R.RRTMGPGridParams(FT=np.float32, domain_nlay=nlay, ncol=ncol, device=local_rank)
    rm = R.AllSkyRadiation(aerosol_radiation=True, reset_rng_seed=True)
    s = R.RRTMGPSolver(gp, rm, R.default_parameters(**PARAMS), pack, col_offset=rank * ncol)
    s.set_state(st)
    R.compute_relative_humidity(s)

    gather_how = None

    def join(solver):
        """Every rank joins the engine's own communicator (NCCL id created by rank 0, shipped through torch.distributed).
        Returns False -- on EVERY rank -- if any rank could not (no libnccl to dlopen, CUDA IPC not permitted ...): the
        step then falls back to torch.distributed's all-gather after the kernels, as in round 1."""
        err = "RRTMGP_BENCH_TORCH_GATHER set" if os.environ.get("RRTMGP_BENCH_TORCH_GATHER") else None
        try:
            box = [R.comm_unique_id() if (rank == 0 and err is None) else None]
        except Exception as e:   # noqa: BLE001
            box, err = [None], repr(e)
        dist.broadcast_object_list(box, src=0)
        if box[0] is not None:
            try:
                R.comm_init(solver, box[0], rank, world)
            except Exception as e:   # noqa: BLE001
                err = repr(e)
        else:
            err = err or "rank 0 could not create the NCCL id"
        bad = torch.tensor([0.0 if err is None else 1.0], device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if bad.item() > 0:
            if rank == 0:
                print(f"bench.py: rrtmgp_b200_comm_init failed on some rank ({err}); using torch.distributed all-gather", file=sys.stderr)
            return False
        return True

    fallback_keys = ("lw_flux_up", "lw_flux_dn", "lw_flux_net", "sw_flux_up", "sw_flux_dn", "sw_flux_net", "sw_flux_dn_dir", "net_flux")

    def make_step(solver, engine_gather):
        if world == 1:
            return lambda seed: R.update_fluxes(solver, seed)
        if engine_gather:   # compute + all-gather of the eight (nlev, ncol) views, overlapped, behind the C ABI
            return lambda seed: R.update_fluxes_gathered(solver, seed)
        n = solver.grid_params.ncol
        solver.gathered = {k: torch.empty(world * n, nlev, dtype=torch.float32, device=dev) for k in fallback_keys}

        def fb(seed):
            R.update_fluxes(solver, seed)
            works = [dist.all_gather_into_tensor(solver.gathered[k], solver.buffers[k], async_op=True) for k in fallback_keys]
            for w in works:
                w.wait()
        return fb

    engine_gather = join(s) if world > 1 else False
    if world > 1:
        gather_how = ("rrtmgp_b200_update_fluxes_gathered (copy-engine pushes over NVLink + NCCL fences)" if engine_gather
                      else "torch.distributed all_gather_into_tensor after the kernels (comm_init unavailable)")
    step = make_step(s, engine_gather)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    launches = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        for i in range(args.steps):
            step(100 + i)
            launches += s.last_launch_count
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * ncol / (ms * 1e-3)

    # --- N > 1: every rank must hold every rank's columns (checksum of each rank's rows against that rank's own) ---
    gather_check = None
    if world > 1:
        from rrtmgp_b200.sharding import FLUX_KEYS as GK
        torch.cuda.synchronize()
        mine = torch.stack([s.buffers[k].double().sum() for k in GK])
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        ok = all(torch.allclose(torch.stack([s.gathered[k][r * ncol:(r + 1) * ncol].double().sum() for k in GK]), sums[r],
                                rtol=1e-12, atol=0) for r in range(world))
        okt = torch.tensor([1.0 if ok and float(mine.abs().sum()) > 0 else 0.0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        gather_check = "ok: every rank holds every rank's rows (8 views, checksums)" if okt.item() == 1.0 else "MISMATCH"

    # --- dominant kernel: per-kernel CUDA-event timing on the launching stream ---
    def time_call(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    k_steps = max(2, min(args.steps, 5))
    ms_prep = time_call(lambda i: R.prepare_atmosphere(s), k_steps)
    ms_lw = time_call(lambda i: R.update_lw_fluxes(s, 200 + i), k_steps)
    ms_sw = time_call(lambda i: R.update_sw_fluxes(s, 200 + i), k_steps)
    esz = 4
    in_common = (nlay * 4 + nlay * 2 + nlay * 5 + 2 * nlay * 15) * esz    # layerdata, h2o+o3, cloud, aerosols
    lw_bytes = in_common + (nlev + 1 + 16) * esz + 3 * nlev * esz           # + t_lev, t_sfc, emis; out up/dn/net
    sw_bytes = in_common + (2 + 2 * 14) * esz + (4 + 1) * nlev * esz + nlev * esz  # + mu0, toa, albedos; out 4 + net (+ lw_net read)
    dom = "LW" if ms_lw >= ms_sw else "SW"
    dom_ms, dom_bytes = (ms_lw, lw_bytes) if dom == "LW" else (ms_sw, sw_bytes)
    hbm_peak, peak_src = peaks()
    achieved = ncol * dom_bytes / (dom_ms * 1e-3) / 1e9
    # measured DRAM traffic of the same kernel: one `ncu --set full` capture (profiles/traffic.json, written by
    # profiles/summarize.py from dram__bytes_read.sum + dram__bytes_write.sum), scaled per column
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic = tj["lw_dram_bytes" if dom == "LW" else "sw_dram_bytes"] / tj["ncol"] * ncol
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic,
                "kernel": f"solve_kernel_fast<{'LW_2STREAM,256' if dom == 'LW' else 'SW_2STREAM,224'},NG=1,cloud,aerosol>",
                "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": ncol * dom_bytes, "peak_source": peak_src,
                "note": "fused path is FP32/LUT-gather bound by construction (SURVEY.md §8d): the HBM fraction is "
                        "<< 1; see roofline_fp32"}
    cols_per_s_gpu = ncol / ((ms_prep + ms_lw + ms_sw) * 1e-3)
    # FP32 peak MEASURED in this run (rrtmgp_b200_measure_fp32_peak: independent scalar FFMA / packed FFMA2 streams on
    # every SM); the fused kernels are scalar FP32, so the scalar-FFMA figure is the denominator
    import ctypes
    ffma, ffma2 = ctypes.c_double(0), ctypes.c_double(0)
    R._lib.check(R._lib.lib().rrtmgp_b200_measure_fp32_peak(local_rank, ctypes.byref(ffma), ctypes.byref(ffma2)))
    fp32_achieved = cols_per_s_gpu * ALGO_FLOPS_PER_COL / 1e12
    roofline_fp32 = {"bound": "fp32", "achieved": fp32_achieved, "peak": ffma.value, "unit": "TFLOP/s",
                     "frac": fp32_achieved / ffma.value,
                     "peak_source": "measured in this run: scalar FFMA stream, 16 warps/SM (rrtmgp_b200_measure_fp32_peak)",
                     "peak_ffma2": ffma2.value, "frac_of_ffma2_peak": fp32_achieved / ffma2.value,
                     "peak_nominal": FP32_NOMINAL_TFLOPS, "algorithmic_flops_per_column": ALGO_FLOPS_PER_COL}
    try:   # the restated reference algorithm's own count (tools/opcount.py: every +, -, *, / of the oracle, per column)
        with open(os.path.join(ROOT, "profiles", "oracle_opcount.json")) as f:
            roofline_fp32["reference_algorithm_ops_per_column"] = json.load(f)["per_column"]["total"]["flops_add_mul_div"]
    except Exception:
        pass
    try:   # what the kernels actually execute (FADD + FMUL + 2 FFMA per thread, counted by ncu for this workload)
        ex = (tj["lw_fp32_flops_executed"] + tj["sw_fp32_flops_executed"]) / tj["ncol"]
        roofline_fp32.update({"executed_flops_per_column": ex, "achieved_executed": cols_per_s_gpu * ex / 1e12,
                              "frac_executed": cols_per_s_gpu * ex / 1e12 / ffma.value,
                              "executed_source": "profiles/traffic.json (" + tj.get("fp32_flops_note", "") + ")"})
    except Exception:
        pass

    # The two limits that actually bind: warp-instruction issue (4 schedulers x 1 instr/clk per SM) and
    # the L1 data pipe (one 128-byte wavefront/clk per SM, shared + global).  Counts per launch are properties of
    # the workload measured once with ncu (profiles/traffic.json); the time is this run's.
    roofline_issue = None
    try:
        tag = "lw" if dom == "LW" else "sw"
        sms, ghz = 148, (clk.summary()["sm_mhz"] or 1965.0) * 1e-3
        scale = ncol / tj["ncol"]
        cyc = dom_ms * 1e-3 * ghz * 1e9
        roofline_issue = {"bound": "issue", "kernel": dom,
                          "warp_instructions_per_launch": tj[f"{tag}_warp_instructions"] * scale,
                          "issue_frac": tj[f"{tag}_warp_instructions"] * scale / (sms * 4 * cyc),
                          "l1_data_pipe_frac": tj[f"{tag}_l1_data_wavefronts_per_sm"] * scale / cyc,
                          "source": "instruction / wavefront counts: " + tj["source"]}
    except Exception:
        pass

    # --- e2e: the state lives in pinned HOST memory; every step copies every per-column input H2D, runs
    # update_fluxes! and copies every flux / diagnostic D2H (HostPipeline: column chunks on 3 streams) ---
    e2e, pipe = None, None
    if not args.no_e2e:
        pipe = R.HostPipeline(s, n_chunks=args.e2e_chunks)
        pipe.load_host_inputs(st)
        pipe.host_in["layerdata"].copy_(s.buffers["layerdata"].cpu())   # rel_hum computed above
        def e2e_step(seed):
            pipe.update_fluxes(seed)            # H2D of every input, update_fluxes!, D2H of every flux (column chunks)
            if world > 1:
                if engine_gather:
                    R.all_gather_fluxes(s)      # + the gather of the eight views (grouped ncclAllGather)
                else:
                    for w in [dist.all_gather_into_tensor(s.gathered[k], s.buffers[k], async_op=True) for k in fallback_keys]:
                        w.wait()
                torch.cuda.synchronize()
        for i in range(2):
            e2e_step(i)
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e2e_step(300 + i)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # the host-side ceiling of this leg: all ranks copy their inputs H2D at the same time, nothing else running
        barrier()
        t0 = time.perf_counter()
        for k in pipe.in_keys:
            s.buffers[k].copy_(pipe.host_in[k], non_blocking=True)
        torch.cuda.synchronize()
        h2d = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(h2d, op=dist.ReduceOp.MAX)
        try:
            cpus = sorted(os.sched_getaffinity(0))
            affinity = f"{cpus[0]}-{cpus[-1]} ({len(cpus)} cpus)"
        except Exception:
            affinity = None
        e2e = {"value": world * ncol / float(dt.item()), "unit": "columns/s", "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": float(dt.item()) * 1e3,
               "includes_gather": world > 1,
               "h2d_only_ms": float(h2d.item()) * 1e3,
               "h2d_only_gbs_per_rank": pipe.h2d_bytes / float(h2d.item()) / 1e9,
               "h2d_only_gbs_all_ranks": world * pipe.h2d_bytes / float(h2d.item()) / 1e9,
               "host_cpu_affinity_rank0": affinity,
               "how": f"HostPipeline: {len(pipe.chunks)} column chunks on {len(pipe.streams)} CUDA streams, pinned host buffers"
                      + ("; then the grouped ncclAllGather of the 8 views" if world > 1 else "")}

    # --- SURVEY §8d variants of the same workload, reported beside the headline (device-resident, rank 0 of a
    # 1-GPU run only): day/night mix cos_zenith ~ U(-0.2, 1) (about 17 % night columns skip the SW solve) and
    # partial cloudiness cld_frac ~ U(0, 1) (McICA masks differ per g-point) ---
    variants = None
    if world == 1 and not args.no_variants:
        variants = {}
        base_cz, base_cf = s.buffers["cos_zenith"].clone(), s.buffers["cld_frac"].clone()
        rng = np.random.default_rng(5)
        for name, key, arr in (("cos_zenith~U(-0.2,1)", "cos_zenith", rng.uniform(-0.2, 1.0, ncol)),
                               ("cld_frac~U(0,1)", "cld_frac", np.where(base_cf.cpu().numpy() > 0, rng.uniform(0.0, 1.0, (ncol, nlay)), 0.0))):
            s.buffers[key].copy_(torch.as_tensor(arr.astype(np.float32)))
            ms_v = time_call(lambda i: R.update_fluxes(s, 300 + i), 3)
            variants[name] = {"value": ncol / (ms_v * 1e-3), "unit": "columns/s", "ms_per_step": ms_v}
            s.buffers["cos_zenith"].copy_(base_cz); s.buffers["cld_frac"].copy_(base_cf)

    # --- BASELINE config 5: weak-scaling sweep ncol/GPU = 1e4 .. 1e6 (device-timed, inputs resident, gather included
    # for N > 1); the 1e5-column synthetic state is tiled on the device for the larger points ---
    sweep = None
    if not args.no_sweep:
        sweep = []
        del pipe
        base = {k: v for k, v in s.buffers.items() if v is not None and v.dim() >= 1 and v.shape[0] == ncol}
        for n_sw in (10_000, 30_000, 100_000, 300_000, 1_000_000):
            if n_sw == ncol:
                sweep.append({"ncol_per_gpu": n_sw, "value": value, "unit": "columns/s", "ms_per_step": ms})
                continue
            gp2 = R.RRTMGPGridParams(FT=np.float32, domain_nlay=nlay, ncol=n_sw, device=local_rank)
            s2 = R.RRTMGPSolver(gp2, rm, R.default_parameters(**PARAMS), pack, col_offset=rank * n_sw)
            for k, v in base.items():
                dst = s2.buffers[k]
                for a in range(0, n_sw, ncol):
                    m = min(ncol, n_sw - a)
                    dst[a:a + m].copy_(v[:m])
            s2.buffers["vmr"].copy_(s.buffers["vmr"])
            eg2 = join(s2) if (world > 1 and engine_gather) else False
            step2 = make_step(s2, eg2)
            fn = lambda i: step2(400 + i)
            fn(0)
            barrier()
            ms_s = torch.tensor([time_call(fn, 3)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms_s, op=dist.ReduceOp.MAX)
            sweep.append({"ncol_per_gpu": n_sw, "value": world * n_sw / (float(ms_s.item()) * 1e-3), "unit": "columns/s",
                          "ms_per_step": float(ms_s.item())})
            if eg2:
                R.comm_destroy(s2)
            del s2, step2
            torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import Oracle
        ncs = min(ncol, args.cpu_sample)
        sub = {k: (v[:ncs] if (v.ndim >= 1 and v.shape[0] == ncol) else v) for k, v in st.items()}
        o = Oracle(pack, np.float32)
        cores = host_threads()
        o.update_fluxes(sub, seed=1, params=PARAMS, nthreads=cores)
        t0 = time.perf_counter()
        o.update_fluxes(sub, seed=1, params=PARAMS, nthreads=cores)
        dtc = time.perf_counter() - t0
        cpu_baseline = {"value": ncs / dtc, "unit": "columns/s", "cores": cores, "kind": "port",
                        "sample": f"first {ncs} of {ncol} columns, one update_fluxes (cost is linear in ncol); C++ "
                                  "restatement of the Julia reference CPU path, OpenMP over columns, Float32"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(ncol, nlay, world),
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "roofline_fp32": roofline_fp32, "roofline_issue": roofline_issue, "cpu_baseline": cpu_baseline,
                "kernel_ms": {"prepare": ms_prep, "lw": ms_lw, "sw": ms_sw}, "variants": variants, "sweep": sweep, "gather_check": gather_check, "gather": gather_how}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
