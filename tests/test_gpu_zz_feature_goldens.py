"""Mapper / writer features of the reference through the command layer, against subsampled reference_output goldens
(tests/golden/*/particles_sample.npz, make_golden.py). Kept in its own file, collected after the other GPU files: these
cases were added after round 1's GPU minutes were spent (host logic verified on the mock device and against the reference
binary; the kernels they use are covered by the earlier files)."""
import importlib.util
import os
import shutil

import numpy as np
import pytest

from genetic_b200 import icgen, output as OUT

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("gx_paramfile_helpers", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_gpu_paramfile.py"))
_pf = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_pf)
_run = _pf._run

_spec2 = importlib.util.spec_from_file_location("gx_reference_helpers", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_gpu_vs_reference.py"))
_vr = importlib.util.module_from_spec(_spec2)
_spec2.loader.exec_module(_vr)


@pytest.mark.skipif(_vr.oracle.ref_binary() is None, reason="oracle/_ref/genetIC not built")
@pytest.mark.parametrize("case", _vr.ADDED_LATER)
def test_stray_and_autopad_particle_files_match_the_reference_binary(tmp_path, golden_dir, case):
    """Stray particles (with and without supersampling) and autopad with gas: files byte-identical to the reference binary's."""
    _vr.zoom_particle_files_case(tmp_path, golden_dir, case)


@pytest.mark.parametrize("test", [
    "test_04a_supersampling", "test_04b_subsampling", "test_04c_doublezoom_supersampling", "test_04d_gas_supersampling",
    "test_06a_stray_particles", "test_06b_stray_supersample", "test_06c_stray_subsampling", "test_06d_stray_wrapping_zoom",
    "test_10d_gadget_flagged_particles", "test_10e_gadget_zoom_ptype_resample", "test_10i_gadget_zoom_ptype_autopad",
    "test_18_intermediate_levels", "test_01f_two_zooms_adapt_mask", "test_03a_baryons_basegrid",
    "test_10h_gadget_zoom_ptype_gas", "test_01g_zoom_grid_vel_offset", "test_17_reversal"])
def test_mapper_and_writer_feature_goldens(golden_dir, tmp_path, test):
    """Sub/supersampling (dark matter and gas), stray particles, flagged and per-level gadget types, autopad, adapt_mask,
    velocity offset and field reversal: the reference's paramfiles against ~4000 evenly spaced particles of its
    reference_output files, with the tolerances of its tools/compare.py (1e-4 on float32 positions / velocities, 1e-6
    on masses)."""
    d, _ = _run(golden_dir, tmp_path, test)
    out = [f for f in os.listdir(d) if f.endswith(".tipsy") or f.endswith(".gadget2") or f.endswith(".gadget3")]
    assert len(out) == 1
    ours = OUT.read_particles(str(d / out[0]))
    gold = np.load(str(d / "particles_sample.npz"))
    assert len(ours["pos"]) == int(gold["count"][0])
    idx = gold["index"]
    if "n" in gold.files:
        assert np.array_equal(ours["n"], gold["n"])              # gas / dark counts
    if "ids" in gold.files:
        assert np.array_equal(ours["ids"][idx], gold["ids"])
    assert np.abs(ours["pos"][idx] - gold["pos"]).max() < 1e-4
    assert np.abs(ours["vel"][idx] - gold["vel"]).max() < 1e-4 * max(1.0, np.abs(gold["vel"]).max() / 10)
    if "mass" in gold.files and len(gold["mass"]) == len(idx):
        assert np.abs(ours["mass"][idx] - gold["mass"]).max() < 1e-6 * max(1.0, np.abs(gold["mass"]).max())
    if "eps" in gold.files:
        assert np.abs(ours["eps"][idx] - gold["eps"]).max() < 1e-6


@pytest.mark.parametrize("fmt,extra", [("tipsy", ""), ("gadget3", ""), ("gadget2", "velocity_offset 10 -20 30\n"),
                                       ("tipsy", "velocity_offset 5 0 -7\n")])
def test_streaming_unigrid_writer_equals_the_mapped_path(tmp_path, golden_dir, monkeypatch, fmt, extra):
    """Large single-level ICs are written component by component through a memory-mapped file (two fields on the device,
    nothing of size n^3 on the host; automatic from 512^3): forced on here at 32^3, the files must equal the mapper path's."""
    files = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("GX_STREAM_UNIGRID", mode)
        d = tmp_path / mode
        d.mkdir()
        shutil.copy(os.path.join(golden_dir, "camb_transfer_kmax40_z0.dat"), d)
        (d / "paramfile.txt").write_text("Om 0.279\nOl 0.721\ns8 0.817\nzin 99\ncamb camb_transfer_kmax40_z0.dat\n"
                                         "random_seed 4242\noutname t\noutdir ./\noutformat %s\n%sbase_grid 70.0 32\ndone\n" % (fmt, extra))
        icgen.run_paramfile(str(d / "paramfile.txt"), log=lambda s: None)
        name = "t.tipsy" if fmt == "tipsy" else "t." + fmt
        files[mode] = (d / name).read_bytes()
        if fmt == "tipsy":
            files[mode + "param"] = (d / "tipsy.param").read_text()
    a, b = files["0"], files["1"]
    if fmt == "tipsy":
        assert a[:28] == b[:28] and files["0param"] == files["1param"]
        x, y = np.frombuffer(a[32:], dtype=np.float32), np.frombuffer(b[32:], dtype=np.float32)
    else:
        assert a[:264] == b[:264] and len(a) == len(b)
        x, y = np.frombuffer(a[264:-4], dtype=np.float32), np.frombuffer(b[264:-4], dtype=np.float32)
    # the fused transform and the unfused one (operator sweep + plain inverse FFT) round differently in the last bits
    ok = np.isfinite(x) & np.isfinite(y)
    assert np.array_equal(ok, np.isfinite(x))
    assert np.abs(x[ok] - y[ok]).max() <= 2e-6 * np.abs(x[ok]).max()


def test_profiler_test_config3_against_the_reference_binary_sample(golden_dir, tmp_path):
    """SURVEY 8(d) config 3 -- genetIC/tests/profiler_test modernised: 256^3 base, two x2 zoom levels of 256^3, an
    overdensity and a filtered-variance modification (serial Fourier seed), 18.1 M particles -- through the command layer,
    against ~6000 evenly spaced particles of the file the reference binary wrote for the same paramfile (36 minutes on the
    build container's 8 threads; tests/golden/make_golden.py). f64 fields, float32 file; tolerances an order tighter than the
    reference's tools/compare.py (the variance iterations amplify rounding differences between FFT implementations)."""
    d, log = _run(golden_dir, tmp_path, "profiler_test_config3")
    ours = OUT.read_particles(str(d / "test_9.tipsy"))
    gold = np.load(str(d / "particles_sample.npz"))
    assert len(ours["pos"]) == int(gold["count"][0]) and np.array_equal(ours["n"], gold["n"])
    idx = gold["index"]
    assert np.array_equal(ours["mass"][idx], gold["mass"]) and np.array_equal(ours["eps"][idx], gold["eps"])
    assert np.abs(ours["pos"][idx] - gold["pos"]).max() < 1e-5
    assert np.abs(ours["vel"][idx] - gold["vel"]).max() < 1e-4 * np.abs(gold["vel"]).max()


def test_config2_512_unigrid_whole_chain_against_the_reference_binary(tmp_path, golden_dir):
    """BASELINE config 2 / SURVEY 8(d) item 2: single-level 512^3, double precision, Fourier-space (parallel) seed, Zeldovich
    displacement + velocity, L = 200 Mpc/h, Planck-2018 transfer -- the paramfile run through the command layer on the GPU
    (streaming writer: staged noise, shared-transform Zeldovich, natural-order c2r with the particle emit) and through the
    reference binary on the host cores (~20 s); the two tipsy files are compared particle by particle."""
    import oracle
    from genetic_b200 import icgen
    if oracle.ref_binary() is None:
        pytest.skip("oracle/_ref/genetIC not built")
    n, box, seed = 512, 200.0, 8896131
    camb = os.path.join(golden_dir, "planck_2018_transfer_out.dat")
    pf = ("Om 0.3158\nOl 0.6842\ns8 0.8117\nns 0.9660\nhubble 0.6732\nz 200\ncamb %s\nrandom_seed %d\noutname c2\noutdir ./\n"
          "outformat tipsy\nbase_grid %r %d\ndone\n" % (os.path.basename(camb), seed, box, n))
    ref_dir, our_dir = tmp_path / "ref", tmp_path / "ours"
    oracle.run_ref(pf, str(ref_dir), [camb], threads=len(os.sched_getaffinity(0)))
    os.makedirs(our_dir)
    shutil.copy(camb, our_dir)
    (our_dir / "paramfile.txt").write_text(pf)
    icgen.run_paramfile(str(our_dir / "paramfile.txt"), log=lambda s: None)
    hdr = 32
    a = np.memmap(str(ref_dir / "c2.tipsy"), dtype=np.float32, mode="r", offset=hdr).reshape(n ** 3, 9)
    b = np.memmap(str(our_dir / "c2.tipsy"), dtype=np.float32, mode="r", offset=hdr).reshape(n ** 3, 9)
    # time, counts: 28 bytes; the last 4 are struct padding the reference writes uninitialised (io/tipsy.hpp:62-71)
    assert open(str(ref_dir / "c2.tipsy"), "rb").read(28) == open(str(our_dir / "c2.tipsy"), "rb").read(28)
    worst_pos = worst_vel = 0.0
    exact = 0
    vmax = float(np.abs(a[::997, 4:7]).max())
    for c0 in range(0, n ** 3, 1 << 24):
        x, y = np.asarray(a[c0:c0 + (1 << 24)]), np.asarray(b[c0:c0 + (1 << 24)])
        assert np.array_equal(x[:, 0], y[:, 0]) and np.array_equal(x[:, 7], y[:, 7])       # mass, softening
        worst_pos = max(worst_pos, float(np.abs(x[:, 1:4] - y[:, 1:4]).max()))
        worst_vel = max(worst_vel, float(np.abs(x[:, 4:7] - y[:, 4:7]).max()))
        exact += int(np.count_nonzero(np.all(x == y, axis=1)))
    # float32 file: positions in box units (|pos| <= 0.5), f64 fields behind them agree to ~1e-13
    assert worst_pos <= np.spacing(np.float32(0.5)) and worst_vel <= 2 * np.spacing(np.float32(vmax))
    assert exact > 0.98 * n ** 3
