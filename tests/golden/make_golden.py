#!/usr/bin/env python
"""Regenerates the fixtures under tests/golden/ from the read-only reference checkout.

Run in the build container only (needs /root/reference). The GPU box never reads /root/reference:
tests use the committed copies.

  * verbatim copies of reference DATA files (no source code): CAMB transfer tables, paramfiles,
    tipsy / text goldens of test_13, test_05a, test_02k, test_02m
  * test_01a_basic_grid/reference_output (9.4 MB, 64^3 tipsy) -> every 61st particle, as .npz
  * test_19_rounding_errors/reference_grid/grid-0.npy (2 MB float64 64^3) -> [::4, ::4, ::4] subsample
  * multi-level goldens: test_01c (zoom), test_02c (two linear modifications with a zoom), test_02i
    (overdensity + variance modification on two levels) reference_grid/grid-{0,1}.npy
    -> [::4, ::4, ::4] subsamples + the id files that flag the zoom region; test_24 (splicing, 16^3) grid-0.npy whole;
    test_02f / test_02g reference.txt (variance / overdensity values); test_14 (vx modification, 8^3) and test_22 (Ly
    modification, 32^3, every 7th particle) gadget3 outputs -> pos/vel arrays
  * particle mappers / writers: test_10b (zoom x3, variable-mass gadget3) and test_10c (zoom with gadget particle types
    1 and 0) reference_output -> header[:208] + every 53rd particle (pos, vel, ids, mass) and the per-type counts;
    mapper_test_01a_zoom (id file round trip through a two-level mapper) paramfile, zoomtest.txt, reference.txt;
    mapper_test_01b / 03a / 03b / 05 / 07 (ids of one setup re-expressed in another, `mapper_relative_to`), 04a-c (stray
    particles), 06 (sub-sampling + adapt_mask + autopad in the input setup): text files
  * mapper / writer feature tests 04a-d, 06a-d, 10d, 10e, 10h, 10i, 18, 01f, 01g, 03a, 17: paramfile and ~4000 evenly spaced
    particles of reference_output (particles_sample.npz)
  * profiler_test_config3: SURVEY 8(d) config 3 run through the reference binary built here (oracle/_ref/genetIC), sampled
"""
import os
import shutil
import struct
import sys
import numpy as np

REF = "/root/reference/genetIC/tests"
EX = "/root/reference/example"
HERE = os.path.dirname(os.path.abspath(__file__))


def cp(src, dst):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    os.chmod(dst, 0o644)


def read_tipsy(fn):
    raw = open(fn, "rb").read()
    a, n, ndim, ngas, ndark, nstar = struct.unpack("<diiiii", raw[:28])
    return np.frombuffer(raw, dtype=np.float32, count=ndark * 9, offset=32).reshape(ndark, 9)


for f in ("camb_transfer_kmax40_z0.dat", "camb_transfer_kmax40_z500.dat"):
    cp(f"{REF}/{f}", f"{HERE}/{f}")
cp(f"{EX}/planck_2018_transfer_out.dat", f"{HERE}/planck_2018_transfer_out.dat")
cp(f"{EX}/paramfile.txt", f"{HERE}/example_paramfile.txt")
for t, files in {
    "test_13_parallel_seed": ("paramfile.txt", "reference_output"),
    "test_05a_seed_old": ("paramfile.txt", "reference_output"),
    "test_02k_chisquare_onelevel": ("paramfile.txt", "reference.txt"),
    "test_02m_chisquare_twolevels": ("paramfile.txt", "reference.txt"),
    "test_01a_basic_grid": ("paramfile.txt",),
    "test_19_rounding_errors": ("paramfile.txt",),
    "test_01c_zoom_grid": ("paramfile.txt", "zoomtest.txt"),
    "test_02c_two_modifs": ("paramfile.txt", "zoomtest.txt"),
    "test_02f_onelevelvariance": ("paramfile.txt", "reference.txt"),
    "test_02g_multilevelvariance": ("paramfile.txt", "reference.txt"),
    "test_24_splicing": ("paramfile.txt",),
    "test_02i_multilevel_algorithm": ("paramfile.txt",),
}.items():
    for f in files:
        cp(f"{REF}/{t}/{f}", f"{HERE}/{t}/{f}")

d = read_tipsy(f"{REF}/test_01a_basic_grid/reference_output")
idx = np.arange(0, d.shape[0], 61)
np.savez_compressed(f"{HERE}/test_01a_basic_grid/particles_every61.npz", index=idx, particles=d[idx])
g = np.load(f"{REF}/test_19_rounding_errors/reference_grid/grid-0.npy")
np.savez_compressed(f"{HERE}/test_19_rounding_errors/grid0_every4.npz", grid=g[::4, ::4, ::4])
for t in ("test_01c_zoom_grid", "test_02c_two_modifs", "test_02i_multilevel_algorithm"):
    gs = {"grid%d" % l: np.load(f"{REF}/{t}/reference_grid/grid-{l}.npy")[::4, ::4, ::4] for l in (0, 1)}
    np.savez_compressed(f"{HERE}/{t}/grids_every4.npz", **gs)
np.savez_compressed(f"{HERE}/test_24_splicing/grid0.npz", grid0=np.load(f"{REF}/test_24_splicing/reference_grid/grid-0.npy"))


def read_gadget_pos_vel(fn, npart):
    """gadget2/3 single file, Fortran record markers (io/gadget.hpp:513-552): header, POS (float32 n x 3), VEL, ID."""
    raw = open(fn, "rb").read()
    pos = np.frombuffer(raw, dtype=np.float32, count=3 * npart, offset=4 + 256 + 4 + 4).reshape(-1, 3)
    vel = np.frombuffer(raw, dtype=np.float32, count=3 * npart, offset=4 + 256 + 4 + 4 + 12 * npart + 4 + 4).reshape(-1, 3)
    return pos, vel


cp(f"{REF}/test_14_velocity_modif/paramfile.txt", f"{HERE}/test_14_velocity_modif/paramfile.txt")
cp(f"{REF}/test_14_velocity_modif/ids.txt", f"{HERE}/test_14_velocity_modif/ids.txt")
pos, vel = read_gadget_pos_vel(f"{REF}/test_14_velocity_modif/reference_output", 8 ** 3)
np.savez_compressed(f"{HERE}/test_14_velocity_modif/particles.npz", pos=pos, vel=vel)
cp(f"{REF}/test_22_angular_momentum_modif/paramfile.txt", f"{HERE}/test_22_angular_momentum_modif/paramfile.txt")
pos, vel = read_gadget_pos_vel(f"{REF}/test_22_angular_momentum_modif/reference_output", 32 ** 3)
idx = np.arange(0, 32 ** 3, 7)
np.savez_compressed(f"{HERE}/test_22_angular_momentum_modif/particles_every7.npz", index=idx, pos=pos[idx], vel=vel[idx])


def read_gadget_blocks(fn):
    raw = open(fn, "rb").read()
    out, off = [], 0
    while off < len(raw):
        n = struct.unpack_from("<i", raw, off)[0]
        out.append(raw[off + 4:off + 4 + n])
        off += n + 8
    return out


for t in ("test_10b_gadget_zoom_varimass", "test_10c_gadget_zoom_ptype"):
    cp(f"{REF}/{t}/paramfile.txt", f"{HERE}/{t}/paramfile.txt")
    cp(f"{REF}/{t}/zoomtest.txt", f"{HERE}/{t}/zoomtest.txt")
    blocks = read_gadget_blocks(f"{REF}/{t}/reference_output")
    n = sum(struct.unpack_from("<6i", blocks[0], 0))
    idx = np.arange(0, n, 53)
    arrays = dict(header=np.frombuffer(blocks[0][:208], dtype=np.uint8), index=idx,
                  pos=np.frombuffer(blocks[1], dtype=np.float32).reshape(n, 3)[idx],
                  vel=np.frombuffer(blocks[2], dtype=np.float32).reshape(n, 3)[idx],
                  ids=np.frombuffer(blocks[3], dtype=np.uint64)[idx])
    if len(blocks) > 4:
        arrays["mass"] = np.frombuffer(blocks[4], dtype=np.float32)[idx]
    np.savez_compressed(f"{HERE}/{t}/particles_every53.npz", **arrays)
for f in ("paramfile.txt", "unzoomed_paramfile.txt", "zoomtest.txt", "reference.txt"):
    cp(f"{REF}/mapper_test_01a_zoom/{f}", f"{HERE}/mapper_test_01a_zoom/{f}")
# id files through `mapper_relative_to` (ids of one setup re-expressed in another): whole directories, text only
for t in ("mapper_test_01b_zoom_gas", "mapper_test_03a_zoom_linkage", "mapper_test_03b_zoom_to_coarse",
          "mapper_test_05_gadget_ptype", "mapper_test_07_ids_from_two_inputs", "mapper_test_04a_stray",
          "mapper_test_04b_stray_linkage", "mapper_test_04c_stray_wrapping", "mapper_test_06_intermediate_levels"):
    for f in os.listdir(f"{REF}/{t}"):
        cp(f"{REF}/{t}/{f}", f"{HERE}/{t}/{f}")

# particle goldens of the mapper / writer features (sub/supersampling, gas, strays, flagged and per-level gadget types,
# autopad, adapt_mask, velocity offset, reversal): paramfile (+ id file) and every k-th particle of reference_output
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from genetic_b200 import output as OUT  # noqa: E402  (reader only)
for t in ("test_04a_supersampling", "test_04b_subsampling", "test_04c_doublezoom_supersampling", "test_04d_gas_supersampling",
          "test_06a_stray_particles", "test_06b_stray_supersample", "test_06c_stray_subsampling", "test_06d_stray_wrapping_zoom",
          "test_10d_gadget_flagged_particles", "test_10e_gadget_zoom_ptype_resample", "test_10i_gadget_zoom_ptype_autopad",
          "test_18_intermediate_levels", "test_01f_two_zooms_adapt_mask", "test_03a_baryons_basegrid",
          "test_10h_gadget_zoom_ptype_gas", "test_01g_zoom_grid_vel_offset", "test_17_reversal"):
    cp(f"{REF}/{t}/paramfile.txt", f"{HERE}/{t}/paramfile.txt")
    if os.path.exists(f"{REF}/{t}/zoomtest.txt"):
        cp(f"{REF}/{t}/zoomtest.txt", f"{HERE}/{t}/zoomtest.txt")
    d = OUT.read_particles(f"{REF}/{t}/reference_output")
    n = len(d["pos"])
    idx = np.arange(0, n, max(1, n // 4000))
    arrays = {"index": idx, "count": np.array([n])}
    for k in ("pos", "vel", "mass", "eps", "ids", "u"):
        if k in d:
            arrays[k] = d[k][idx] if len(d[k]) == n else d[k][:1]
    if "n" in d:
        arrays["n"] = d["n"]
    np.savez_compressed(f"{HERE}/{t}/particles_sample.npz", **arrays)

# SURVEY 8(d) config 3 (profiler_test, modernised): ~6000 evenly spaced particles of the output of the reference BINARY
# (oracle/_ref/genetIC; 36 minutes on 8 threads, so only regenerated on request: GX_GOLDEN_SLOW=1)
cfg3 = f"{HERE}/profiler_test_config3"
cp(os.path.join(os.path.dirname(os.path.dirname(HERE)), "tools", "paramfiles", "profiler_test_modernised.txt"), f"{cfg3}/paramfile.txt")
if os.environ.get("GX_GOLDEN_SLOW") == "1":
    import tempfile
    import oracle
    with tempfile.TemporaryDirectory() as tmp:
        oracle.run_ref(open(f"{cfg3}/paramfile.txt").read(), tmp, [f"{HERE}/camb_transfer_kmax40_z0.dat"], timeout=4 * 3600)
        d = OUT.read_particles(os.path.join(tmp, "test_9.tipsy"))
    n = len(d["pos"])
    idx = np.arange(0, n, n // 6000)
    np.savez_compressed(f"{cfg3}/particles_sample.npz", index=idx, count=np.array([n]), pos=d["pos"][idx], vel=d["vel"][idx],
                        mass=d["mass"][idx], eps=d["eps"][idx], n=d["n"])
print("ok")


# ---- grafic output, grafic mapper ids, refinement masks, import_level (round 2) ---------------------------------------------
def read_grafic_dir(path):
    """The ten Fortran-record files of one level directory of the reference's grafic output."""
    out = {}
    for name in sorted(os.listdir(path)):
        raw = open(os.path.join(path, name), "rb").read()
        hdr = struct.unpack("<iiiffffffff", raw[4:48])
        n = hdr[0]
        dtype = np.uint64 if name == "ic_particle_ids" else np.float32
        plane = n * n * np.dtype(dtype).itemsize
        body = np.frombuffer(raw, dtype=np.uint8, offset=52).reshape(n, plane + 8)
        out[name] = np.ascontiguousarray(body[:, 4:4 + plane]).view(dtype).reshape(n, n, n)
        out["header"] = np.array(hdr, dtype=np.float64)
    return out


cp(f"{REF}/camb_transfer_kmax40_z0_post2015.dat", f"{HERE}/camb_transfer_kmax40_z0_post2015.dat")
for t in ("test_12_grafic", "test_08e_mask_multiple_virtual_grids", "test_08f_mask_multiple_supersample"):
    cp(f"{REF}/{t}/paramfile.txt", f"{HERE}/{t}/paramfile.txt")
    packed = {}
    for d in sorted(os.listdir(f"{REF}/{t}/reference_grafic")):
        res = d.rsplit("_", 1)[1]
        for k, v in read_grafic_dir(f"{REF}/{t}/reference_grafic/{d}").items():
            packed["%s/%s" % (res, k)] = v
    np.savez_compressed(f"{HERE}/{t}/reference_grafic.npz", **packed)
for t, extra in {"test_08a_mask_lagrange_volume": ("lagrange.txt",), "test_08b_mask_virtual_architecture": (),
                 "test_08c_centering_masks": (), "test_08d_mask_flag_interaction": ("lagrange.txt",),
                 "test_08e_mask_multiple_virtual_grids": (), "test_08f_mask_multiple_supersample": ()}.items():
    for f in ("paramfile.txt",) + extra:
        cp(f"{REF}/{t}/{f}", f"{HERE}/{t}/{f}")
    masks = {os.path.basename(f)[:-4]: np.load(f"{REF}/{t}/reference_mask/{f}").astype(np.uint8)
             for f in sorted(os.listdir(f"{REF}/{t}/reference_mask")) if f.endswith(".npy")}
    np.savez_compressed(f"{HERE}/{t}/reference_mask.npz", **masks)
for t, files in {"mapper_test_02a_grafic": ("paramfile.txt", "test_ids.txt", "reference.txt"),
                 "mapper_test_02b_grafic_zoom_to_coarse": ("paramfile.txt", "test_ids.txt", "reference.txt"),
                 "mapper_test_02c_grafic_relative": ("paramfile.txt", "test_ids.txt", "reference.txt", "morezoomed.txt"),
                 "mapper_test_02d_grafic_rel_subsupersample": ("paramfile.txt", "test_ids.txt", "reference.txt", "morezoomed.txt"),
                 "test_01h_zoom_grid_pos_offset": ("paramfile.txt",), "test_01i_zoom_pos_offset_wrap": ("paramfile.txt",),
                 "test_11_import": ("paramfile.txt", "reference_grid/grid-0.npy")}.items():
    for f in files:
        cp(f"{REF}/{t}/{f}", f"{HERE}/{t}/{f}")

# test_01d (two nested zooms): the three power-spectrum dumps -- modes sit exactly on bin edges there (found on the B200)
cp(f"{REF}/test_01d_two_zooms/paramfile.txt", f"{HERE}/test_01d_two_zooms/paramfile.txt")
for l in (0, 1, 2):
    cp(f"{REF}/test_01d_two_zooms/reference_ps/test_01d_{l}.ps", f"{HERE}/test_01d_two_zooms/reference_ps/test_01d_{l}.ps")
