#!/usr/bin/env python
"""Timing of BASELINE.json config 3: three-level zoom -- base n^3 over L = 50 Mpc/h plus two 2x zoom
grids of n^3 -- with an overdensity and a filtered-variance modification, through the multi-level host layer
(genetic_b200/multilevel.py) on one B200; optionally the CPU oracle on the same paramfile-equivalent at a reduced size.

  python tools/zoom_bench.py [--n 256] [--cpu-n 64] [--seed 1512351] [--no-cpu]

Equivalent paramfile (tests/profiler_test/paramfile.txt of the reference, modernised as in SURVEY 8d):
  base_grid 50 n; centre 25 25 25; select_sphere 5; zoom_grid 2 n; select_sphere 3; zoom_grid 2 n;
  centre 25 25 25.1; select_sphere 2; modify overdensity relative 1; filtering_scale 0.5; modify variance relative 3;
  random_seed_serial <seed>; done
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CAMB_FILE = os.path.join(ROOT, "tests", "golden", "planck_2018_transfer_out.dat")
COSMO = dict(Om=0.3158, Ol=0.6842, s8=0.8117, z=200.0, ns=0.9660, hubble=0.6732)


def run_device(n, seed, timing=False):
    import genetic_b200 as gb
    from genetic_b200 import multilevel as GM, zoom as Z, lib as L
    lib = L.load()
    lib.gx_set_device(0)
    cos = gb.CosmologicalParameters(COSMO["Om"], COSMO["Ol"], COSMO["s8"], COSMO["z"], ns=COSMO["ns"], hubble=COSMO["hubble"])
    spectra = {s: gb.CAMB(cos, CAMB_FILE, s) for s in ("dm", "all")}

    def once():
        ctx = GM.MultiLevelGrid()
        ctx.addLevel(gb.Grid(50.0, n))
        for k, v in spectra.items():
            ctx.setPowerspectrumGenerator(k, v)
        Z.selectSphere(ctx, (25, 25, 25), 5)
        Z.initZoomGrid(ctx, 2, n)
        Z.selectSphere(ctx, (25, 25, 25), 3)
        Z.initZoomGrid(ctx, 2, n)
        Z.selectSphere(ctx, (25, 25, 25.1), 2)
        t = {}
        lib.gx_synchronize(None)
        t0 = time.perf_counter()
        field = GM.draw(ctx, seed, drawInFourierSpace=True)
        lib.gx_synchronize(None)
        t["draw"] = time.perf_counter() - t0
        mm = GM.ModificationManager(ctx, cos, field)
        mm.addModificationToList("overdensity", "relative", 1.0)
        mm.addModificationToList("variance", "relative", 3.0, filterscale=0.5)
        mm.applyModifications()
        lib.gx_synchronize(None)
        t["modifications"] = time.perf_counter() - t0 - t["draw"]
        field.applyPowerSpectrumFor(GM.ALL)
        GM.MultiLevelParticleGenerator(field)
        lib.gx_synchronize(None)
        t["total"] = time.perf_counter() - t0
        t["combine"] = t["total"] - t["draw"] - t["modifications"]
        sums = []
        for l in range(ctx.getNumLevels()):      # outside the timed region: a value to compare runs by
            d = field.getFieldForLevel(l)
            d.toReal()
            sums.append(float((d.getDataVector().astype("float64") ** 2).sum()))
        t["sumsq"] = sums
        return t
    once()
    l0 = lib.gx_launch_count()
    t = once()
    t["launches"] = int(lib.gx_launch_count() - l0)
    if os.environ.get("GX_ZOOM_CPROFILE"):       # host-side profile of one more run (development)
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        once()
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(35)
    if timing:                                   # a third run with the library's per-class CUDA-event timing on
        import ctypes as C
        lib.gx_timing_enable(1)
        once()
        buf = C.create_string_buffer(1 << 16)
        lib.gx_timing_report(buf, len(buf))
        lib.gx_timing_enable(0)
        t["classes_ms"] = {a: (int(b), round(float(c), 2)) for a, b, c in (ln.split() for ln in buf.value.decode().splitlines())}
    return t


def run_oracle(n, seed):
    from oracle import model as M, multilevel as ML
    cos = M.Cosmology(COSMO["Om"], COSMO["Ol"], COSMO["s8"], COSMO["z"], ns=COSMO["ns"], hubble=COSMO["hubble"])
    camb = M.CAMB(CAMB_FILE, cos)
    t0 = time.perf_counter()
    grids = [ML.Grid(50.0, n)]
    ML.select_sphere(grids, (25, 25, 25), 5)
    ML.add_zoom_level(grids, 2, n)
    ML.select_sphere(grids, (25, 25, 25), 3)
    ML.add_zoom_level(grids, 2, n)
    ML.select_sphere(grids, (25, 25, 25.1), 2)
    f = ML.draw_levels(grids, seed, "fourier_serial")
    cdm = [camb.covariance(g.size, g.n, "dm") for g in grids]
    call = [camb.covariance(g.size, g.n, "all") for g in grids]
    a = ML.overdensity_covector(grids, cdm)
    v = ML.inner_product(grids, a, f)
    var0 = ML.variance_value(grids, cdm, f, 0.5)
    f, vecs = ML.apply_linear_modifications(grids, f, [a], [v])
    f = ML.apply_variance_modification(grids, cdm, f, vecs, 3.0 * var0, 0.5)
    ML.combine_levels(grids, [M.apply_transfer(x, c, 0.5) for x, c in zip(f, call)])
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--cpu-n", type=int, default=64)
    ap.add_argument("--seed", type=int, default=1512351)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--timing", action="store_true", help="add per-kernel-class device time of one more run")
    a = ap.parse_args()
    t = run_device(a.n, a.seed, a.timing)
    cells = 3 * a.n ** 3
    line = {"workload": "three-level zoom %d^3 x 3, overdensity + variance modification, random_seed_serial %d" % (a.n, a.seed),
            "seconds": t["total"], "cells_per_s": cells / t["total"], "breakdown_s": {k: t[k] for k in ("draw", "modifications", "combine")},
            "gpu_launches": t["launches"], "sumsq_per_level": t["sumsq"]}
    if "classes_ms" in t:
        line["classes_ms"] = t["classes_ms"]
        line["device_ms"] = round(sum(v[1] for v in t["classes_ms"].values()), 1)
    if not a.no_cpu:
        tc = run_oracle(a.cpu_n, a.seed)
        line["cpu_oracle"] = {"n": a.cpu_n, "seconds": tc, "cells_per_s": 3 * a.cpu_n ** 3 / tc, "cores": os.cpu_count(),
                              "kind": "port (numpy/pocketfft + C spline/RNG)"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
