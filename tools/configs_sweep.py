#!/usr/bin/env python
"""Device-timed `update_fluxes!` for every BASELINE.json configuration that fits one B200, plus the kernel
variants the headline bench does not exercise (clear sky, no-scattering LW, Float64, per-band fluxes, taller
columns).  One JSON line per case on stdout (and into `--out`); inputs resident in HBM, CUDA events on the
launching stream, 3 warm-up + `--steps` timed steps per case.  This is a measurement script, not a test:
parity of every case is covered by tests/test_gpu_parity.py.

    python tools/configs_sweep.py [--out gpurun_out/configs.jsonl] [--steps 3] [--quick]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PARAMS = dict(grav=9.80665, molmass_dryair=0.028964, molmass_water=0.018016)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--quick", action="store_true", help="skip the 3e5 / 1e6 column cases")
    ap.add_argument("--only", default=None, help="run only the cases whose name contains this substring")
    args = ap.parse_args()
    import torch
    import rrtmgp_b200 as R
    assert torch.cuda.is_available()
    pack = R.synthetic.make_lut_pack(seed=7)

    def run(name, ncol, nlay, dtype, method, aerosols, *, clouds=True, lw_noscat=False, n_gauss=1, spectral=False,
            cld_frac=1.0, baseline_config=None):
        if args.only and args.only not in name:
            return
        st = R.synthetic.make_atmosphere(ncol, nlay, dtype=dtype, cld_frac=cld_frac, cos_zenith=0.86,
                                         clouds=clouds, aerosols=aerosols)
        gp = R.RRTMGPGridParams(FT=dtype, domain_nlay=nlay, ncol=ncol)
        rm = {"clear_sky": lambda: R.ClearSkyRadiation(aerosol_radiation=aerosols),
              "all_sky": lambda: R.AllSkyRadiation(aerosol_radiation=aerosols, reset_rng_seed=True)}[method]()
        s = R.RRTMGPSolver(gp, rm, R.default_parameters(**PARAMS), pack,
                           op_lw="one_scalar" if lw_noscat else "two_stream", n_gauss_angles=n_gauss,
                           spectral_fluxes=spectral)
        s.set_state(st)
        R.compute_relative_humidity(s)

        def timed(fn):
            for i in range(3):
                fn(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for i in range(args.steps):
                fn(10 + i)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / args.steps
        ms = timed(lambda i: R.update_fluxes(s, i))
        ms_lw = timed(lambda i: R.update_lw_fluxes(s, i))
        ms_sw = timed(lambda i: R.update_sw_fluxes(s, i))
        line = {"case": name, "baseline_config": baseline_config, "ncol": ncol, "nlay": nlay,
                "dtype": "f64" if dtype == np.float64 else "f32", "method": method, "aerosols": aerosols,
                "lw_solver": f"noscat x{n_gauss}" if lw_noscat else "two_stream", "spectral_fluxes": spectral,
                "ms_per_step": ms, "ms_lw": ms_lw, "ms_sw": ms_sw, "columns_per_s": ncol / (ms * 1e-3)}
        print(json.dumps(line), flush=True)
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")
        del s
        torch.cuda.empty_cache()

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").close()
    f32, f64 = np.float32, np.float64
    run("clear_sky two-stream f64", 128, 64, f64, "clear_sky", False, clouds=False, baseline_config=1)
    run("cloudy_sky two-stream + McICA f32", 4096, 64, f32, "all_sky", False, cld_frac=None, baseline_config=2)
    for n in (10_000, 30_000, 100_000):
        run("all_sky_with_aerosols f32", n, 64, f32, "all_sky", True, baseline_config=3 if n == 100_000 else 4)
    run("clear_sky two-stream f32", 100_000, 64, f32, "clear_sky", False, clouds=False)
    run("all_sky_with_aerosols f32, per-band fluxes", 100_000, 64, f32, "all_sky", True, spectral=True)
    run("all_sky_with_aerosols f32, nlay 95", 50_000, 95, f32, "all_sky", True)
    run("all_sky_with_aerosols f32, noscat LW 1 angle", 100_000, 64, f32, "all_sky", True, lw_noscat=True)
    run("all_sky_with_aerosols f32, noscat LW 3 angles", 50_000, 64, f32, "all_sky", True, lw_noscat=True, n_gauss=3)
    run("all_sky_with_aerosols f64", 20_000, 64, f64, "all_sky", True)
    run("clear_sky noscat f64", 20_000, 64, f64, "clear_sky", False, clouds=False, lw_noscat=True)
    for n in ([] if args.quick else [300_000, 1_000_000]):   # last: the big host-side generations
        run("all_sky_with_aerosols f32", n, 64, f32, "all_sky", True, baseline_config=4)


if __name__ == "__main__":
    main()
