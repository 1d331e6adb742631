"""The C++ host layer for zoom ICs (include/genetic_b200_zoom.hpp, host/genetic_b200_zoom.cpp) against the Python layer on
CPU: `--layout-only` runs a paramfile's geometry commands (selection, zoom-window placement, trimming, flag transfer,
mapper construction with gas and gadget types, id files) and dumps the mapper's layout before anything touches the device.
The Python layer is itself pinned on the reference binary and the reference's goldens (tests/test_mapper.py,
tests/test_gpu_vs_reference.py)."""
import os
import subprocess

import numpy as np
import pytest

from genetic_b200 import build as gbuild, icgen, mapper as PM

COSMO = "Om 0.279\nOl 0.721\ns8 0.817\nzin 99\n"
CASES = {
    "two_levels_tipsy": "outformat tipsy\nbasegrid 60.0 32\ncentre 20 33 41\nselect_sphere 7\nzoomgrid 2 32\n",
    "three_levels_gadget": ("outformat gadget3\nbasegrid 60.0 32\ncentre 20 33 41\nselect_sphere 7\nzoomgrid 2 32\n"
                            "centre 20 33 41\nselect_sphere 3\nzoomgrid 2 32\n"),
    "wrapping_cube_zoom3": "outformat gadget2\nbasegrid 50.0 24\ncentre 1 10 49\nselect_cube 9\nzoomgrid 3 24\n",
    "gadget_types": ("outformat gadget3\nbasegrid 60.0 32\ngadget_particle_type 2\ncentre 20 33 41\nselect_sphere 7\n"
                     "zoomgrid 2 32\ngadget_particle_type 1\n"),
    "trimmed_origin": "outformat gadget3\nbasegrid 60.0 32\ncentre 20 33 41\nselect_sphere 12\nzoomgrid_with_origin_at 6 9 13 2 32\n",
    "gas_gadget": "Ob 0.04\noutformat gadget3\nbasegrid 60.0 32\ncentre 20 33 41\nselect_sphere 7\nzoomgrid 2 32\n",
    "gas_tipsy_all_levels": ("Ob 0.04\noutformat tipsy\nbaryons_all_levels\nbasegrid 60.0 32\ncentre 20 33 41\nselect_sphere 7\n"
                             "zoomgrid 2 32\n"),
    "flagged_particle_type": ("outformat gadget3\nbasegrid 60.0 32\ncentre 20 33 41\nselect_sphere 7\nzoomgrid 2 32\n"
                              "centre 20 33 41\nselect_sphere 2\ngadget_flagged_particle_type 3\n"),
    "nearest_then_ids": ("outformat tipsy\nbasegrid 50.0 32\ncentre 24.219 24.219 24.219\nselect_nearest\nzoomgrid 4 32\n"
                         "idfile ids.txt\ndump_idfile out_ids.txt\n"),
}


def _python_layout(text, workdir):
    gen = icgen.ICGenerator(log=lambda s: None)
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        for line in text.splitlines():
            if line.strip() and not line.startswith("dump_idfile"):
                gen.run_line(line)
        m = gen._mapper()
        flagged = m.getFlaggedParticles()
    finally:
        os.chdir(cwd)
    lay, cell = m.layout()
    layers = m.layers()
    level = np.array([layers[l].level for l in lay], dtype=np.int32)
    mass, eps = np.empty(len(lay)), np.empty(len(lay))
    for k, one in layers.items():
        w = lay == k
        mass[w] = PM.particle_mass(one.grid, gen.cosmology)
        eps[w] = PM.particle_eps(one.grid)
    gas = m.isGas() if hasattr(m, "isGas") else np.zeros(len(lay), dtype=bool)
    return level, cell, mass, eps, m.attrs()[0], gas, flagged


@pytest.mark.parametrize("case", sorted(CASES))
def test_cpp_layout_matches_python(tmp_path, case):
    exe = gbuild.build_host(program="genetic_b200_zoom")
    text = COSMO + CASES[case]
    if "ids.txt" in text:
        np.savetxt(str(tmp_path / "ids.txt"), np.array([3, 77, 4000, 20000, 32766, 32767, 32770, 32800, 32830]), fmt="%d")
    (tmp_path / "paramfile.txt").write_text(text + "done\n")
    out = tmp_path / "layout.bin"
    r = subprocess.run([exe, str(tmp_path / "paramfile.txt"), "--layout-only", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = out.read_bytes()
    n = int(np.frombuffer(raw, dtype=np.int64, count=1)[0])
    off = 8
    arrays = []
    for dt in (np.int32, np.int64, np.float64, np.float64, np.int32, np.uint8):
        arrays.append(np.frombuffer(raw, dtype=dt, count=n, offset=off))
        off += n * np.dtype(dt).itemsize
    level, cell, mass, eps, gtype, gas, flagged = _python_layout(text, str(tmp_path))
    assert n == len(level)
    assert np.array_equal(arrays[0], level) and np.array_equal(arrays[1], cell)      # bit-exact particle ordering
    assert np.array_equal(arrays[2], mass) and np.array_equal(arrays[3], eps)        # same double arithmetic
    assert np.array_equal(arrays[4], gtype) and np.array_equal(arrays[5].astype(bool), gas)
    if "ids.txt" in text:
        assert np.array_equal(np.loadtxt(str(tmp_path / "out_ids.txt"), dtype=np.int64), flagged)


def test_cpp_zoom_driver_errors(tmp_path):
    exe = gbuild.build_host(program="genetic_b200_zoom")
    p = tmp_path / "paramfile.txt"
    p.write_text("Om 0.3\nzoomgrid 2 64\n")
    r = subprocess.run([exe, str(p), "--layout-only", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and 'on line 2 ("zoomgrid 2 64")' in r.stderr and "base grid" in r.stderr
    p.write_text("basegrid 50 16\nzoomgrid 2 16\n")
    r = subprocess.run([exe, str(p), "--layout-only", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "without marking particles" in r.stderr
    p.write_text("basegrid 50 16\nsupersample 2\n")
    r = subprocess.run([exe, str(p), "--layout-only", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown command supersample" in r.stderr


@pytest.mark.gpu
def test_cpp_zoom_driver_matches_the_reference_binary(tmp_path, golden_dir):
    """host/genetic_b200_zoom on the device: a two-level zoom written as gadget3, byte for byte the file the reference
    binary writes for the same paramfile (geometry, mapper order, level combination, particle evaluation, writer in C++)."""
    import shutil
    import oracle
    if oracle.ref_binary() is None:
        pytest.skip("oracle/_ref/genetIC not built")
    exe = gbuild.build_host(program="genetic_b200_zoom")
    camb = os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat")
    pf = (COSMO + "camb camb_transfer_kmax40_z0.dat\nrandom_seed_real_space 90210\noutname t\noutdir ./\noutformat gadget3\n"
          "basegrid 60.0 32\ncentre 20 33 41\nselect_sphere 7\nzoomgrid 2 32\ndone\n")
    oracle.run_ref(pf, str(tmp_path / "ref"), [camb])
    ours = tmp_path / "ours"
    ours.mkdir()
    shutil.copy(camb, ours)
    (ours / "paramfile.txt").write_text(pf)
    r = subprocess.run([exe, str(ours / "paramfile.txt")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (ours / "t.gadget3").read_bytes() == (tmp_path / "ref" / "t.gadget3").read_bytes()


WRITER_CASES = {
    "tipsy_two_levels": CASES["two_levels_tipsy"],
    "tipsy_gas_all_levels": CASES["gas_tipsy_all_levels"],
    "gadget3_gas": CASES["gas_gadget"],
    "gadget3_types_two_files": "gadget_num_files 2\n" + CASES["gadget_types"],
    "gadget2_three_levels": CASES["three_levels_gadget"].replace("gadget3", "gadget2"),
}


@pytest.mark.parametrize("case", sorted(WRITER_CASES))
def test_cpp_writers_match_python(tmp_path, case):
    """The C++ tipsy / gadget writers (gas records, per-type ordering, variable masses, multi-file split, unit conversions)
    fed with the same particles as the Python writers (which are byte-identical to the reference binary's files):
    `--particles-from` replaces the device evaluation by a file, so this runs without a GPU."""
    from genetic_b200 import output as OUT
    exe = gbuild.build_host(program="genetic_b200_zoom")
    text = COSMO + WRITER_CASES[case] + "outname t\noutdir ./\n"
    (tmp_path / "paramfile.txt").write_text(text + "done\n")
    gen = icgen.ICGenerator(log=lambda s: None)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        for line in text.splitlines():
            if line.strip():
                gen.run_line(line)
        m = gen._mapper()
    finally:
        os.chdir(cwd)
    lay, cell = m.layout()
    N = len(lay)
    rng = np.random.default_rng(11)
    pos, vel = rng.uniform(0, 60, (N, 3)), rng.normal(0, 300, (N, 3))
    np.concatenate([pos.reshape(-1), vel.reshape(-1)]).tofile(str(tmp_path / "particles.bin"))

    class P:
        pass
    P.pos, P.vel = pos, vel
    P.mass, P.eps = np.empty(N), np.empty(N)
    for k, one in m.layers().items():
        P.mass[lay == k] = PM.particle_mass(one.grid, gen.cosmology)
        P.eps[lay == k] = PM.particle_eps(one.grid)
    (P.gadgetType,) = m.attrs()
    P.isGas = m.isGas() if hasattr(m, "isGas") else np.zeros(N, dtype=bool)
    py = tmp_path / "py"
    py.mkdir()
    if gen.outputFormat == "tipsy":
        OUT.saveTipsy(str(py / "t.tipsy"), 60.0, P, gen.cosmology)
        names = ["t.tipsy"]
    else:
        P.pos, P.vel = pos.astype(np.float32), vel.astype(np.float32)
        names = [os.path.basename(f) for f in OUT.saveGadget(str(py / "t.gadget"), 60.0, P, gen.cosmology,
                                                             int(gen.outputFormat[-1]), gen.nGadgetFiles)]
    r = subprocess.run([exe, str(tmp_path / "paramfile.txt"), "--particles-from", str(tmp_path / "particles.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for name in names:
        a, b = (py / name).read_bytes(), (tmp_path / name).read_bytes()
        if name.endswith(".tipsy"):
            assert a[:28] == b[:28] and a[32:] == b[32:]
        else:
            assert a == b
