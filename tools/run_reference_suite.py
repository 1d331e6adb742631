#!/usr/bin/env python
"""Runs the reference's own test suite (genetIC/tests/test_* and mapper_test_*) through `genetic_b200.icgen` and compares
with the reference's goldens using the tolerances of the reference's tools/compare.py (particles: 1e-4 abs on float32
pos/vel, masses 1e-6; grids 1e-4 abs; power spectra rtol 1e-4; reference.txt lines present in the log; id files exact).

Development tool for the build container: needs /root/reference (never read by tests, smoke or bench).
  python tools/run_reference_suite.py [--mock] [test_dir ...]
--mock runs the host layer against tests/host/mock_device.py (numpy stand-in for the C ABI) when there is no GPU.
"""
import glob
import os
import shutil
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("GX_REFERENCE_TESTS", "/root/reference/genetIC/tests")   # a copy of the test DATA on a GPU box


def read_particles(fn):
    from genetic_b200 import output as OUT
    return OUT.read_particles(fn)


def read_multifile(first):
    """A split gadget snapshot as one particle set, type-major like a single file (what pynbody presents)."""
    from genetic_b200 import output as OUT
    parts, i = [], 0
    while os.path.exists(first[:-1] + str(i)):
        parts.append(OUT.read_gadget(first[:-1] + str(i)))
        i += 1
    out = {}
    for k in ("pos", "vel", "ids", "mass"):
        if k not in parts[0]:
            continue
        chunks = []
        for t in range(6):
            for p in parts:
                o = sum(p["npart"][:t])
                chunks.append(p[k][o:o + p["npart"][t]])
        out[k] = np.concatenate(chunks)
    return out


def compare_dir(d, log):
    msgs = []
    ref_txt = os.path.join(d, "reference.txt")
    if os.path.exists(ref_txt) and not os.path.basename(d).startswith("mapper_test"):
        lines = [l.strip() for l in log]
        for s in open(ref_txt).read().splitlines():
            assert s.strip() in lines, "log line missing: %r" % s
        msgs.append("log")
    if os.path.basename(d).startswith("mapper_test"):
        assert open(ref_txt).read().split() == open(os.path.join(d, "output.txt")).read().split(), "id output differs"
        msgs.append("ids")
    out = [f for f in glob.glob(d + "/*.tipsy") + glob.glob(d + "/*.gadget?") + glob.glob(d + "/*.gadget?.0")]
    if os.path.exists(os.path.join(d, "reference_output")) and os.path.getsize(os.path.join(d, "reference_output")) > 100 \
            and not (glob.glob(d + "/*.grafic_*") and not out):    # compare.py:190-191: grafic output without a particle file is not compared
        assert len(out) == 1, "expected one particle file, found %s" % out
        a = read_particles(os.path.join(d, "reference_output"))
        b = read_multifile(out[0]) if out[0].endswith(".0") else read_particles(out[0])
        for k in a:
            if k == "u" and k not in b:
                continue
            assert k in b, "block %s missing" % k
            assert a[k].shape == b[k].shape, "%s: shape %s vs %s" % (k, a[k].shape, b[k].shape)
            tol = {"pos": 1e-4, "vel": 1e-4, "mass": 1e-6, "eps": 1e-6}.get(k, 0)
            if k == "vel" and "ids" in a:
                tol = 1e-4 * max(1.0, np.abs(a[k]).max() / 10)   # gadget km/s velocities: float32 decimal=4 on O(100) values
            err = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max() if a[k].size else 0
            assert err <= tol, "%s differs by %g" % (k, err)
        msgs.append("particles")
        ro = os.path.join(d, "reference_output.overdensity")
        if os.path.exists(ro):
            mine = glob.glob(d + "/*.overdensity") + glob.glob(d + "/*.overden")
            mine = [m for m in mine if not m.endswith("reference_output.overdensity")]
            assert len(mine) == 1, "expected one dump_tipsy array"
            x, y = np.fromfile(ro, dtype=np.float32, offset=4), np.fromfile(mine[0], dtype=np.float32, offset=4)
            assert x.shape == y.shape and np.abs(x - y).max() < 1e-5 * max(1.0, np.abs(x).max()), "overdensity array differs"
            msgs.append("overdensity")
    rg = os.path.join(d, "reference_grid")
    if os.path.isdir(rg):
        for f in sorted(glob.glob(rg + "/grid-?.npy")):
            x, y = np.load(f), np.load(os.path.join(d, os.path.basename(f)))
            assert x.shape == y.shape and np.abs(x - y).max() < 1e-4, "%s differs by %g" % (os.path.basename(f), np.abs(x - y).max())
        msgs.append("grids")
    rm = os.path.join(d, "reference_mask")
    if os.path.isdir(rm):
        masks = sorted(glob.glob(rm + "/mask-?.npy"))
        assert masks, "no reference masks"
        for f in masks:
            x, y = np.load(f), np.load(os.path.join(d, os.path.basename(f)))
            assert x.shape == y.shape and np.abs(x - y).max() < 1e-4, "%s differs" % os.path.basename(f)
        msgs.append("masks")
    rgf = os.path.join(d, "reference_grafic")
    if os.path.isdir(rgf):
        from genetic_b200 import grafic as GF
        ref_levels = sorted(glob.glob(rgf + "/*.grafic_*"), key=lambda p: int(p.rsplit("_", 1)[1]))
        my_levels = sorted(glob.glob(d + "/*.grafic_*"), key=lambda p: int(p.rsplit("_", 1)[1]))
        assert [p.rsplit("_", 1)[1] for p in ref_levels] == [p.rsplit("_", 1)[1] for p in my_levels], "grafic resolutions differ"
        for a, b in zip(ref_levels, my_levels):
            x, y = GF.read_grafic_level(a), GF.read_grafic_level(b)
            assert sorted(x) == sorted(y), "grafic files differ in %s" % os.path.basename(a)
            assert np.allclose(x["header"], y["header"], rtol=1e-6), "grafic header differs in %s: %s vs %s" % (os.path.basename(a), x["header"], y["header"])
            for k in x:
                if k == "header":
                    continue
                if k == "ic_particle_ids":
                    assert np.array_equal(x[k], y[k]), "ids differ in " + os.path.basename(a)
                else:
                    err = np.abs(x[k].astype(np.float64) - y[k].astype(np.float64)).max()
                    assert err < 1e-4 * max(1.0, np.abs(x[k]).max() / 10), "%s/%s differs by %g" % (os.path.basename(a), k, err)
        msgs.append("grafic")
    if not msgs and glob.glob(d + "/*.grafic_*"):
        msgs.append("ran (the reference's compare.py has no check for this case)")
    rp = os.path.join(d, "reference_ps")
    if os.path.isdir(rp):
        for f in sorted(glob.glob(rp + "/*.ps")):
            mine = [g for g in glob.glob(d + "/*.ps") if g.endswith(f[-5:])]
            assert mine, "no power spectrum for " + f
            np.testing.assert_allclose(np.loadtxt(f), np.loadtxt(mine[0]), rtol=1e-4)
        msgs.append("ps")
    assert msgs, "nothing to compare"
    return "+".join(msgs)


def main():
    args = sys.argv[1:]
    mock = "--mock" in args
    cpp = "--cpp" in args          # through the C++ driver host/genetic_b200_modify instead of the Python command layer
    args = [a for a in args if a not in ("--mock", "--cpp")]
    if mock:
        import importlib.util
        spec = importlib.util.spec_from_file_location("gx_mock_device", os.path.join(ROOT, "tests", "host", "mock_device.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)

        class MP:
            def setattr(self, o, n, v):
                setattr(o, n, v)
        m.install(MP())
        import genetic_b200.icgen as I
        I.CAMB = m.MockSpectrum
    from genetic_b200 import icgen, lib as L
    work = "/tmp/gx_suite"
    os.makedirs(work, exist_ok=True)
    for f in glob.glob(REF + "/*.dat"):
        shutil.copy(f, work)
    tests = args or sorted(os.path.basename(p) for p in glob.glob(REF + "/test_*") + glob.glob(REF + "/mapper_test_*"))
    results = []
    for t in tests:
        d = os.path.join(work, t)
        shutil.rmtree(d, ignore_errors=True)
        shutil.copytree(os.path.join(REF, t), d)
        log = []
        t0 = time.time()
        try:
            if cpp:
                import subprocess
                from genetic_b200 import build as gbuild
                r = subprocess.run([gbuild.build_host(program="genetic_b200_modify"), os.path.join(d, "paramfile.txt")],
                                   capture_output=True, text=True, cwd=d, timeout=600)
                log.extend(r.stdout.splitlines())
                if r.returncode != 0:
                    raise L.GxError((r.stderr.strip().splitlines() or ["exit %d" % r.returncode])[-1])
            else:
                icgen.run_paramfile(os.path.join(d, "paramfile.txt"), log=lambda s: log.extend(str(s).splitlines() or [""]))
            status = "PASS " + compare_dir(d, log)
        except L.GxError as e:
            msg = str(e)
            if "error" in t and "Error" in msg:
                status = "PASS (expected error)"
            elif "not supported" in msg or "Unknown output format" in msg or (cpp and ("Unknown command" in msg or "in the C++ layer" in msg)):
                status = "SKIP " + msg[:110]
            else:
                status = "FAIL " + msg[:160]
        except AssertionError as e:
            status = "FAIL " + str(e)[:160]
        except Exception as e:
            status = "FAIL " + "".join(traceback.format_exception_only(type(e), e)).strip()[:160] + " @ " + traceback.format_tb(e.__traceback__)[-1].strip().splitlines()[0][-70:]
        results.append((t, status, time.time() - t0))
        print("%-45s %6.1fs  %s" % (t, time.time() - t0, status), flush=True)
    n = {k: sum(1 for r in results if r[1].startswith(k)) for k in ("PASS", "SKIP", "FAIL")}
    print("passed %d, skipped (outside scope) %d, failed %d of %d" % (n["PASS"], n["SKIP"], n["FAIL"], len(results)))


if __name__ == "__main__":
    main()
