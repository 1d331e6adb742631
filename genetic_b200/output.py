"""Particle file writers fed from mapper.MappedParticles (host arrays in mapper order).

Reference (paths relative to genetIC/src/io/):
  tipsy::TipsyOutput::operator()     tipsy.hpp:207-271   units, header, dark particle = 9 float32, photogenic.txt, tipsy.param
  gadget::GadgetOutput::operator()   gadget.hpp:340-556  Fortran records: header (256 B), pos, vel, ids, [mass]; by gadget type
  createGadget2Header / 3Header      gadget.hpp:116-239
  tools::MemMapFileWriter            tools/memmap.hpp:169-215  32-bit record markers

Gas particles (OmegaBaryons0 > 0: mapper.AddGasMapper) are written as tipsy gas records / gadget type 0 with the dummy
internal-energy block. HDF5/swift/grafic are not written.
"""
import math
import os
import struct

import numpy as np


def _sequential_sum(values, dtype=np.float64):
    """tot_mass += mass in mapper order (tipsy.hpp:217-225): sequential, not pairwise, and in the generator's FloatType --
    `FloatType mass, tot_mass` is float under genetIC -f."""
    return float(np.cumsum(np.asarray(values, dtype=dtype), dtype=dtype)[-1])


def _wrap_int32(v):
    """A particle count as the reference's `int npart[6]` holds it (gadget.hpp:28,122-127): the low 32 bits, signed."""
    v = int(v) & 0xffffffff
    return v - (1 << 32) if v >= (1 << 31) else v


def saveTipsy(filename, boxLength, particles, cosmology, aux_dir=None):
    """tipsy::save. `particles` must hold positions/velocities in the grid's floating type (the conversions below are
    done in double before the float32 store, like the reference). Writes tipsy.param and, when masses differ,
    photogenic.txt next to the file (the reference writes them into its working directory)."""
    ftype = np.float32 if np.asarray(particles.mass).dtype == np.float32 else np.float64
    mass = particles.mass.astype(np.float64)
    min_mass, max_mass = float(mass.min()), float(mass.max())
    mass_factor = float(cosmology.OmegaM0) / _sequential_sum(particles.mass, ftype)
    pos_factor = 1. / float(boxLength)
    dKpcUnit = float(boxLength) * 1000 / float(cosmology.hubble)
    dMsolUnit = 1e10 / float(cosmology.hubble) / mass_factor
    dKmsUnit = float(np.sqrt(4.3022682e-6 * dMsolUnit / dKpcUnit))
    vel_factor = float(cosmology.scalefactor) ** -0.5 / dKmsUnit
    N = len(mass)
    gas = np.asarray(getattr(particles, "isGas", np.zeros(N, dtype=bool)))
    ngas = int(gas.sum())
    if ngas and not gas[:ngas].all():
        raise ValueError("tipsy output needs the gas particles first (AddGasMapper with gasFirst)")

    def records(sl, width, eps_col):
        rec = np.zeros((sl.stop - sl.start, width), dtype=np.float32)
        rec[:, 0] = mass[sl] * mass_factor
        rec[:, 1:4] = particles.pos[sl].astype(np.float64) * pos_factor - 0.5
        rec[:, 4:7] = particles.vel[sl].astype(np.float64) * vel_factor
        rec[:, eps_col] = particles.eps[sl].astype(np.float64) * pos_factor
        return rec

    # gas: mass, x, y, z, vx, vy, vz, rho, temp, eps, metals, phi; dark: mass, x, y, z, vx, vy, vz, eps, phi
    grec = records(slice(0, ngas), 12, 9)
    grec[:, 8] = _temperature(cosmology)
    rec = records(slice(ngas, N), 9, 7)
    aux_dir = aux_dir if aux_dir is not None else (os.path.dirname(filename) or ".")
    if min_mass != max_mass:
        with open(os.path.join(aux_dir, "photogenic.txt"), "w") as f:
            for i in np.flatnonzero(mass == min_mass):
                f.write("%d\n" % i)
    with open(os.path.join(aux_dir, "tipsy.param"), "w") as f:
        f.write("dKpcUnit = %s\ndMsolUnit = %s\ndHubble0 = %s\nbComove = 1 \n"
                % (_g6(dKpcUnit), _g6(dMsolUnit), _g6(0.1 * float(cosmology.hubble) * dKpcUnit / dKmsUnit)))
    with open(filename, "wb") as f:
        # (int n, ngas, ndark in the reference's header struct: counts beyond 2^31 wrap, tipsy.hpp:20-27)
        f.write(struct.pack("<diiiii4x", float(cosmology.scalefactor), _wrap_int32(N), 3, _wrap_int32(ngas),
                            _wrap_int32(N - ngas), 0))
        grec.tofile(f)
        rec.tofile(f)


def _temperature(cosmology):
    """cosmology::getTemperature (parameters.hpp:36-42); baryon decoupling at a = 1/150 (ic.hpp:206)."""
    a, adec = float(cosmology.scalefactor), 1. / 150.
    if a > adec:
        return float(cosmology.TCMB) * adec / (a * a)
    return float(cosmology.TCMB) / a


def _g6(x):
    """operator<<(ostream, double) with the default precision (6 significant digits, %g)."""
    return "%g" % x


def _fortran(f, payload):
    n = len(payload)
    marker = struct.pack("<i", n if n <= 0x7fffffff else 0x7fffffff)
    f.write(marker)
    f.write(payload)
    f.write(marker)


def _gadget_header(version, masses, npart, npartTotal, nFiles, boxLength, cosmology, doubleprecision):
    h = bytearray(256)
    struct.pack_into("<6i", h, 0, *[_wrap_int32(v) for v in npart])
    struct.pack_into("<6d", h, 24, *[float(m) for m in masses])
    struct.pack_into("<dd", h, 72, float(cosmology.scalefactor), float(cosmology.redshift))
    baryons = 1 if npartTotal[0] > 0 else 0
    struct.pack_into("<ii", h, 88, baryons, baryons)                              # flag_sfr, flag_feedback
    struct.pack_into("<6I", h, 96, *[int(v) & 0xffffffff for v in npartTotal])
    struct.pack_into("<ii", h, 120, baryons, int(nFiles))                         # flag_cooling, num_files
    struct.pack_into("<4d", h, 128, float(boxLength), float(cosmology.OmegaM0), float(cosmology.OmegaLambda0),
                     float(cosmology.hubble))
    struct.pack_into("<ii", h, 160, 0, 0)                                         # flag_stellarage, flag_metals
    struct.pack_into("<6I", h, 168, *[int(v) >> 32 for v in npartTotal])
    struct.pack_into("<i", h, 192, 0)                                             # flag_entropy_instead_u
    if version == 3:
        struct.pack_into("<iif", h, 196, 1 if doubleprecision else 0, 1, 0.0)     # flag_doubleprecision, flag_ic_info, lpt
    return bytes(h)


def saveGadget(name, boxLength, particles, cosmology, gadgetVersion=3, nFiles=1, out_dtype=np.float32):
    """gadget::save<float>: files `name + "2"/"3"` (+ ".i" when split). Particles are written by gadget type 0..5, within a
    type in mapper order; ids are mapper indices (unsigned long)."""
    if gadgetVersion not in (2, 3):
        raise ValueError("Unknown gadget format")
    order = np.argsort(particles.gadgetType, kind="stable")
    types = particles.gadgetType[order]
    nPartPerType = [int(np.count_nonzero(types == t)) for t in range(6)]
    nTotal = len(order)
    ordered = np.array_equal(order, np.arange(nTotal))    # already type-ordered (one type, or gas last): no gather
    massT = particles.mass if ordered else particles.mass[order]
    masses = [0.0] * 6
    variableMass = False
    for t in range(6):
        if nPartPerType[t]:
            m = massT[types == t]
            if m.min() != m.max():
                variableMass = True
            masses[t] = float(m.min())
    if variableMass:
        masses = [0.0] * 6
    nPartPerFile = [nTotal // nFiles] * (nFiles - 1) + [nTotal - (nTotal // nFiles) * (nFiles - 1)]
    if ordered:
        pos, vel = (np.asarray(a, dtype=out_dtype) for a in (particles.pos, particles.vel))
    else:
        pos, vel = (a[order].astype(out_dtype) for a in (particles.pos, particles.vel))
    mass = massT.astype(out_dtype)
    ids = order.astype(np.uint64)
    remainingPerType = list(nPartPerType)
    start = 0
    written = []
    for i in range(nFiles):
        room = nPartPerFile[i]
        thisFile = [0] * 6
        for t in range(6):
            thisFile[t] = min(remainingPerType[t], room)
            room -= thisFile[t]
            remainingPerType[t] -= thisFile[t]
        fn = name + str(gadgetVersion) + ("" if nFiles == 1 else "." + str(i))
        sl = slice(start, start + nPartPerFile[i])
        with open(fn, "wb") as f:
            _fortran(f, _gadget_header(gadgetVersion, masses, thisFile, nPartPerType, nFiles, boxLength, cosmology,
                                       np.dtype(out_dtype) == np.float64))
            _fortran(f, pos[sl].tobytes())
            _fortran(f, vel[sl].tobytes())
            _fortran(f, ids[sl].tobytes())
            if variableMass:
                _fortran(f, mass[sl].tobytes())
            if float(cosmology.OmegaBaryons0) > 0:
                _fortran(f, np.zeros(thisFile[0], dtype=out_dtype).tobytes())        # dummy internal energy, gadget.hpp:545-552
        start += nPartPerFile[i]
        written.append(fn)
    return written


def read_gadget(filename):
    """Reader for the files above (tests): header fields and the blocks present."""
    with open(filename, "rb") as f:
        raw = f.read()

    def block(off):
        n = struct.unpack_from("<i", raw, off)[0]
        return raw[off + 4:off + 4 + n], off + 8 + n

    h, off = block(0)
    npart = struct.unpack_from("<6i", h, 0)
    out = {"npart": npart, "mass_table": struct.unpack_from("<6d", h, 24), "time": struct.unpack_from("<d", h, 72)[0],
           "redshift": struct.unpack_from("<d", h, 80)[0], "nPartTotal": struct.unpack_from("<6I", h, 96),
           "num_files": struct.unpack_from("<i", h, 124)[0], "BoxSize": struct.unpack_from("<d", h, 128)[0],
           "Omega0": struct.unpack_from("<d", h, 136)[0], "OmegaLambda": struct.unpack_from("<d", h, 144)[0],
           "HubbleParam": struct.unpack_from("<d", h, 152)[0], "header": h}
    n = sum(npart)
    b, off = block(off)
    out["pos"] = np.frombuffer(b, dtype=np.float32).reshape(n, 3)
    b, off = block(off)
    out["vel"] = np.frombuffer(b, dtype=np.float32).reshape(n, 3)
    b, off = block(off)
    out["ids"] = np.frombuffer(b, dtype=np.uint64)
    for _ in range(2):
        if off < len(raw):
            b, off = block(off)
            out["mass" if len(b) == 4 * n and "mass" not in out else "u"] = np.frombuffer(b, dtype=np.float32)
    return out


def writePowerSpectrum(bins, grid, filename, nBins=100):
    """The text file of cosmology::dumpPowerSpectrum (cosmology/parameters.hpp:172-205) from the per-bin sums the device
    reduction returns (gx_power_spectrum_bins: count, sum k, sum |delta_k|^2, sum C(k) for 100 logarithmic bins between the
    fundamental and the Nyquist wavenumber); columns: bin centre, mean k, mean theory P, mean |delta_k|^2 / norm, count."""
    n = grid.size
    inBin, kbin, Gx, Px = (np.asarray(bins[i], dtype=np.float64) for i in range(4))
    boxLength = float(grid.thisGridSize)
    kmax, kmin = math.pi / boxLength * n, 2.0 * math.pi / boxLength
    dklog, kw = math.log10(kmax / kmin) / nBins, 2.0 * math.pi / boxLength
    psnorm = 1.0 / ((kw * kw * kw / float(np.float32(2.0 * np.float32(math.pi)) ** np.float32(3.0))) * float(n) ** 3)
    with open(filename, "w") as f:
        for i in range(nBins):
            if inBin[i] > 0:
                f.write("".join("%16s" % ("%g" % v) for v in (
                    10. ** (math.log10(kmin) + dklog * (i + 0.5)), kbin[i] / inBin[i], Px[i] / inBin[i] * psnorm,
                    Gx[i] / inBin[i] * psnorm, inBin[i])) + "\n")



def saveFieldTipsyArray(filename, values):
    """io::tipsy::saveFieldTipsyArray (tipsy.hpp:38-54), `dump_tipsy`: a 4-byte count, then the field value of every
    particle's cell in mapper order as float32 (mapper.field_on_mapper)."""
    out = np.asarray(values).astype(np.float32)
    with open(filename, "wb") as f:
        f.write(struct.pack("<I", len(out) & 0xffffffff))
        out.tofile(f)


def read_particles(filename):
    """Reader for the files written above (tests and tools): tipsy (gas + dark records) or a single gadget2/3 file, as a
    dict of arrays (pos, vel, mass, and ids / u for gadget, eps and the counts for tipsy)."""
    with open(filename, "rb") as f:
        head = f.read(4)
    if head == b"\x00\x01\x00\x00":
        g = read_gadget(filename)
        return {k: g[k] for k in ("pos", "vel", "ids", "mass", "u") if k in g}
    with open(filename, "rb") as f:
        b = f.read()
    a, n, ndim, ngas, ndark, nstar = struct.unpack("<diiiii", b[:28])
    gas = np.frombuffer(b, dtype=np.float32, count=ngas * 12, offset=32).reshape(ngas, 12)
    dark = np.frombuffer(b, dtype=np.float32, count=ndark * 9, offset=32 + ngas * 48).reshape(ndark, 9)
    return {"a": np.array([a]), "pos": np.concatenate([gas[:, 1:4], dark[:, 1:4]]), "vel": np.concatenate([gas[:, 4:7], dark[:, 4:7]]),
            "mass": np.concatenate([gas[:, 0], dark[:, 0]]), "eps": np.concatenate([gas[:, 9], dark[:, 7]]),
            "n": np.array([ngas, ndark])}
