#!/usr/bin/env python
"""bench.py -- IC particles/sec for a Zeldovich unigrid IC generated from a seed (BASELINE.json metric).

One "step" = the whole hot path for one IC: draw white noise from the seed (GSL-exact, parallel Fourier order)
-> delta_k = sqrt(C(k)) * noise -> psi_xyz = c2r(i k_j delta_k / k^2) -> particle positions and velocities.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n GRID] [--impl ours|reference|reference-cufft]

  value : N^3 / step time, device timed (CUDA events), results left in HBM (staging buffer)
  e2e   : the same through the C-ABI call with HOST buffers: the transfer table is uploaded and the particles are
          copied to pinned host memory inside the timed region
  roofline     : the dominant kernel's algorithmic GB/s (2*B per FFT pass) against MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle port (oracle/) timed on this box's host cores on a bounded sample
`--impl reference` times the reference's own CPU implementation: oracle/_ref/genetIC, the unmodified reference sources
compiled against stand-in FFTW3/GSL/HDF5 headers (oracle/build_ref.sh), or the numpy/C oracle port if that
binary was not built, with the same metric/config keys.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
CAMB_FILE = os.path.join(GOLD, "planck_2018_transfer_out.dat")
SEED = 8896131
COSMO = dict(Om=0.3158, Ol=0.6842, s8=0.8117, z=200.0, ns=0.9660, hubble=0.6732)  # example/paramfile.txt:4-10


def box_for(n):
    """512^3 -> L = 200, 2048^3 -> L = 1000, 4096^3 -> L = 2000 Mpc/h (k_Nyq stays inside the CAMB table)."""
    return 1000.0 * n / 2048.0 if n >= 2048 else 200.0 * n / 512.0


def workload_name(n, dtype):
    return "%d^3 unigrid %s, L=%g Mpc/h, Planck-2018 CAMB transfer, random_seed %d (parallel Fourier), " \
           "Zeldovich displacement+velocity" % (n, dtype, box_for(n), SEED)


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
def cpu_port_step(n, threads):
    """One IC through the CPU oracle port (the reference's algorithm on the host cores). Returns seconds.
    C/OpenMP: per-shell ranlxs2+ziggurat draw (randomfieldgenerator.hpp:226-243), covariance field and transfer
    (camb.hpp:101-119, field.hpp:394-412). numpy/scipy(pocketfft, `threads` workers): ik/k^2 and the three c2r."""
    import numpy as np
    import scipy.fft
    from oracle import model as M, crng
    box = box_for(n)
    cos = M.Cosmology(COSMO["Om"], COSMO["Ol"], COSMO["s8"], COSMO["z"], ns=COSMO["ns"], hubble=COSMO["hubble"])
    t0 = time.perf_counter()
    buf = crng.draw_fourier_parallel_shells(SEED, n, 0, n // 2)
    W = buf.view(np.complex128).reshape(n, n, n // 2 + 1)
    camb = M.CAMB(CAMB_FILE, cos)
    cov = camb.covariance(box, n, "all")
    D = M.apply_transfer(W, cov, 0.5)
    psis_k = M.zeldovich_offsets_fourier(D, box)
    psis = []
    for p in psis_k:
        M.mirror_fast(p)
        psis.append(scipy.fft.irfftn(p, s=(n, n, n), axes=(0, 1, 2), workers=threads) * (float(n) ** 1.5))
    pos, vel, mass, eps = M.particles_unigrid(psis, box, cos)
    return time.perf_counter() - t0


def host_threads():
    """Threads given to the reference: every host core of the box, set explicitly (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_binary_step(n, workdir=None, cufft=False):
    """One IC of the same workload through the reference ITSELF (oracle/_ref/genetIC: the unmodified sources compiled
    against the stand-in FFTW3/GSL/HDF5 headers of oracle/shims/): whole-process wall clock, including the tipsy file it
    has to write (the reference has no output-less mode). Returns (seconds, threads) or None if the binary is absent."""
    import shutil
    import tempfile
    import oracle
    if oracle.ref_binary(cufft) is None:
        return None
    d = tempfile.mkdtemp(prefix="gx_ref_", dir=workdir)
    try:
        pf = ("Om %(Om)r\nOl %(Ol)r\ns8 %(s8)r\nns %(ns)r\nhubble %(hubble)r\nz %(z)r\n" % COSMO +
              "camb %s\nrandom_seed %d\noutname bench\noutdir ./\noutformat tipsy\nbase_grid %r %d\ndone\n"
              % (os.path.basename(CAMB_FILE), SEED, box_for(n), n))
        threads = host_threads()
        t0 = time.perf_counter()
        oracle.run_ref(pf, d, [CAMB_FILE], cufft=cufft, threads=threads)
        return time.perf_counter() - t0, threads
    finally:
        shutil.rmtree(d, ignore_errors=True)


REF_SAMPLE = ("%d^3 unigrid of the same paramfile through the reference binary (unmodified genetIC sources compiled against "
              "stand-in headers: radix-2 OpenMP FFT instead of FFTW, restated GSL ranlxs2/ziggurat/cspline), %d OpenMP threads, "
              "whole process including its tipsy file write")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    n = args.cpu_n
    kind = "port"
    cufft = args.impl == "reference-cufft"   # the reference's own GPU path (USE_CUFFT=1: FFTs through cuFFTW, rest on the host)
    try:
        first = reference_binary_step(min(n, 64), cufft=cufft)  # warm-up (page cache, OpenMP pool) and the availability check
    except RuntimeError as e:
        if not cufft:
            raise
        first = None
        print("reference-cufft failed: %s" % str(e).strip().splitlines()[-1], file=sys.stderr)
    if cufft and first is None:
        print(json.dumps({"impl": "reference-cufft", "unavailable": "oracle/_ref/genetIC_cufft not built, or no GPU for cuFFTW"}))
        return 0
    if first is not None:
        kind = "reference"
        times = [reference_binary_step(n, cufft=cufft)[0] for _ in range(max(1, min(args.steps, 2)))]
        sample = REF_SAMPLE % (n, threads) + (" (OMP_NUM_THREADS=%d set for the subprocess; the %d^3 workload needs ~0.4 TB of "
                                              "host memory in the reference's allocation pattern and hours of CPU time)"
                                              % (threads, args.n))
        if cufft:
            sample = sample.replace("radix-2 OpenMP FFT instead of FFTW", "USE_CUFFT=1: FFTs through the toolkit's cuFFTW on the GPU")
    else:
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_port_step(min(n, 128), threads)
        times = [cpu_port_step(n, threads) for _ in range(max(1, min(args.steps, 3)))]
        sample = "%d^3 unigrid (the %d^3 workload does not fit the CPU time budget); CPU oracle port: C/OpenMP draw + " \
                 "covariance, numpy/pocketfft FFTs with %d workers (reference binary oracle/_ref/genetIC not built)" % (
                     n, args.n, threads)
    t = statistics.mean(times)
    v = n ** 3 / t
    line = {"impl": args.impl, "metric": "IC particles/sec (Zeldovich %d^3)" % args.n, "value": v,
            "unit": "particles/s", "n_gpus": args.gpus, "steps": len(times), "warmup": 1, "ms_per_step": t * 1e3,
            "sample_n": n, "omp_num_threads": threads,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.n, "f64"), "sample": sample},
            "cpu_baseline": {"value": v, "unit": "particles/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------
def particle_checksum(lib, torch, step, dist=None):
    """One extra IC (outside every timed region) with the emit hook on: per axis sum(pos), sum(vel), sum(vel^2) over all
    particles, summed over the ranks. The N-rank and the 1-rank run of the same workload must agree to rounding."""
    acc = torch.zeros(9, dtype=torch.float64, device="cuda")
    lib.gx_emit_checksum(C.c_void_p(acc.data_ptr()))
    try:
        step()
        torch.cuda.synchronize()
    finally:
        lib.gx_emit_checksum(None)
    if dist is not None:
        dist.all_reduce(acc)
    a = acc.cpu().numpy().reshape(3, 3)
    return {"pos_sum": [float("%.13e" % v) for v in a[:, 0]], "vel_sum": [float("%.13e" % v) for v in a[:, 1]],
            "vel_sumsq": [float("%.13e" % v) for v in a[:, 2]],
            "note": "per axis, double accumulation over every emitted particle of one extra IC; compare across n_gpus"}


def cufft_crosscheck(lib, torch, n, dt, planes=8):
    """The three hand-written pass kinds at the bench's line length against cuFFT (torch.fft on the same device) on a thin
    slab of `planes` planes of random data, outside the timed loop (the north star allows cuFFT for exactly this)."""
    if dt != 1:   # F64 only: the comparison is at 1e-13
        cdt, rdt, tol = torch.complex64, torch.float32, 2e-5
    else:
        cdt, rdt, tol = torch.complex128, torch.float64, 1e-12
    nc, P = n // 2 + 1, min(planes, n)
    world = n // P
    g = torch.Generator(device="cuda").manual_seed(7)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())

    out = {"n": n, "planes": P, "tolerance": tol}
    # y-pass: x-slab [P][n][nc], inverse along axis 1
    x = torch.randn(P, n, nc, 2, dtype=rdt, device="cuda", generator=g)
    y = torch.empty_like(x)
    lib.gx_slab_fft_y(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), n, dt, world, 1, 0, 0, st)
    ref = torch.fft.ifft(torch.view_as_complex(x), dim=1, norm="forward")
    out["y_pass_rel_err"] = rel(torch.view_as_complex(y), ref)
    # z-pass: c2r rows scaled by n^-1.5, in place, padded pitch n + 2 (Hermitian-consistent input: cuFFT's single-precision
    # c2r does not ignore the imaginary parts of the DC and Nyquist entries the way FFTW and our pass do)
    x[:, :, 0, 1] = 0
    x[:, :, n // 2, 1] = 0
    z = x.clone()
    lib.gx_slab_fft_z_c2r(C.c_void_p(z.data_ptr()), n, dt, world, st)
    ref = torch.fft.irfft(torch.view_as_complex(x), n=n, dim=2, norm="forward") / float(n) ** 1.5
    out["z_pass_rel_err"] = rel(z.view(P, n, n + 2)[:, :, :n], ref)
    # x-pass: ky-slab [n][P][nc], inverse along axis 0, in place, unscaled
    w = torch.randn(n, P, nc, 2, dtype=rdt, device="cuda", generator=g)
    ref = torch.fft.ifft(torch.view_as_complex(w), dim=0, norm="forward")
    lib.gx_slab_fft_x(C.c_void_p(w.data_ptr()), n, dt, world, 1, 0, st)
    out["x_pass_rel_err"] = rel(torch.view_as_complex(w), ref)
    torch.cuda.synchronize()
    out["ok"] = all(out[k] < tol for k in ("x_pass_rel_err", "y_pass_rel_err", "z_pass_rel_err"))
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from genetic_b200 import lib as L
    import genetic_b200 as gb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    lib = L.load()
    lib.gx_set_device(local_rank)
    if world > 1:
        from genetic_b200 import slab
        return slab.bench_main(args, lib, world, rank, local_rank)

    n = args.n
    dtype = np.float64
    dt = L.F64
    box = box_for(n)
    cos = gb.CosmologicalParameters(COSMO["Om"], COSMO["Ol"], COSMO["s8"], COSMO["z"], ns=COSMO["ns"],
                                    hubble=COSMO["hubble"])
    grid = gb.Grid(box, n, dtype=dtype)
    spectrum = gb.CAMB(cos, CAMB_FILE, "all", dtype)
    fbytes = lib.gx_field_bytes(n, dt)
    delta, psi, stage = C.c_void_p(), C.c_void_p(), C.c_void_p()
    lib.gx_malloc(C.byref(delta), fbytes)
    lib.gx_malloc(C.byref(psi), fbytes)
    osz = 4  # the reference's writers store float32 (ic.hpp:1294-1305)
    n3 = n ** 3
    # staging (two halves of pos+vel): 16 GiB when the fields leave room -- the device->host copies of one component then
    # keep the PCIe link busy through the ~80 ms of column passes of the next one (2 GiB drained in 40 ms and left it idle)
    free_b, total_b = torch.cuda.mem_get_info()
    stage_cap = (16 << 30) if free_b > (24 << 30) else (2 << 30)
    stage_bytes = min(2 * 2 * n3 * osz, stage_cap)
    stage_bytes = max(stage_bytes, 2 * 2 * n * n * osz)
    lib.gx_malloc(C.byref(stage), stage_bytes)
    # host arrays: full component-major arrays if they fit 24 GiB, else a ring
    cap = n3 if 6 * n3 * osz <= (24 << 30) else (8 << 30) // (3 * osz)
    hpos, hvel = C.c_void_p(), C.c_void_p()
    with L.numa_local(torch.cuda.current_device()) as numa:      # pinned pages on the GPU's NUMA node
        lib.gx_host_alloc_pinned(C.byref(hpos), 3 * cap * osz)
        lib.gx_host_alloc_pinned(C.byref(hvel), 3 * cap * osz)
    pinned_cpus = numa.cpus

    prm = L.IcParams()
    prm.n, prm.dtype, prm.out_dtype = n, dt, L.F32
    prm.noise_mode, prm.reverse_order, prm.fix_power = L.NOISE_FOURIER_PARALLEL, 0, 0
    prm.seed = SEED
    prm.grid_length, prm.cell_size, prm.domain_size = float(grid.thisGridSize), float(grid.cellSize), float(
        grid.periodicDomainSize)
    prm.offset_lower[0] = prm.offset_lower[1] = prm.offset_lower[2] = 0.0
    prm.vel_ratio = cos.zeldovichVelocityToOffsetRatio(dtype)

    def device_step():
        lib.gx_unigrid_ic(C.byref(prm), spectrum.handle, delta, psi, stage, stage_bytes, None, None, 0, None)

    def e2e_step():
        sp = gb.CAMB(cos, CAMB_FILE, "all", dtype)  # host spline + normalisation, table upload (H2D)
        lib.gx_unigrid_ic(C.byref(prm), sp.handle, delta, psi, stage, stage_bytes, hpos, hvel, cap, None)
        return sp

    def barrier():
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.gx_timing_enable(1)
    l0 = lib.gx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = lib.gx_launch_count() - l0
    buf = C.create_string_buffer(1 << 16)
    lib.gx_timing_report(buf, len(buf))
    lib.gx_timing_enable(0)
    clocks = sampler.stop()
    classes = {}
    for line in buf.value.decode().strip().split("\n"):
        if line:
            nm, cnt, tot = line.split()
            classes[nm] = (int(cnt), float(tot))

    # e2e: host buffers, copies inside the timed region (wall clock around synchronous calls + device sync)
    if args.no_e2e:   # development switch for kernel A/B runs; the line then carries no e2e number
        e2e_s = float("nan")
    else:
        for _ in range(min(args.warmup, 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
    tab_bytes = 3 * 8 * 213
    checksum = particle_checksum(lib, torch, device_step)
    crosscheck = cufft_crosscheck(lib, torch, n, dt)

    # roofline of the dominant kernel class
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # Algorithmic bytes per step of each pass class: a column pass reads and writes the field once
    # (2 B) per component; the z-pass reads the field (B) and writes pos+vel of every particle (2 * osz bytes each). A
    # class may run as several launches per step (the z-pass is launched once per staging slab), so the rate is
    # class bytes / class time and "per launch" is the mean over the class's launches.
    def n_launch(k):
        return classes[k][0] / args.steps if k in classes else 0

    # (a column pass is one launch: the shared-transform schedule runs 3 x-passes and 2 y-passes per IC)
    n_gauss = 8 * sum(3 * ks * ks for ks in range(1, n // 2))   # 24 ks^2 gaussians per shell
    class_bytes = {"fft_cols_x": n_launch("fft_cols_x") * 2 * fbytes, "fft_cols_y": n_launch("fft_cols_y") * 2 * fbytes,
                   "fft_rows_c2r": 3 * (fbytes + 2 * osz * n3), "apply_transfer": 2 * fbytes,
                   "rng_pass_b": fbytes,
                   # staged noise: gen writes 4 bytes per accepted draw (it is instruction-bound: ~1250 integer instructions
                   # of ranlxs2 per 24 raw outputs); deliver reads them once and writes every mode of the field once
                   "rng_gen": 4 * n_gauss, "rng_deliver": 4 * n_gauss + fbytes}
    passes = {}
    for k, (cnt, tot) in classes.items():
        if k in class_bytes and tot > 0:
            gbs = class_bytes[k] / (tot / args.steps * 1e-3) / 1e9
            passes[k] = {"GB/s": gbs, "frac": gbs / peak, "ms_per_step": tot / args.steps,
                         "launches_per_step": cnt / args.steps}
    # the dominant kernel class: the largest share of the step among ALL classes (noise included)
    cand = {k: v for k, v in classes.items() if k in class_bytes and class_bytes[k] > 0}
    dom = max(cand, key=lambda k: cand[k][1]) if cand else None
    roof = None
    if dom:
        cnt, tot = classes[dom]
        avg_ms = tot / cnt
        bytes_per_launch = class_bytes[dom] * args.steps / cnt
        achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        # dram bytes per launch of the same kernel from ncu (profiles/r2_ncu_dram_2048.json, written by
        # tools/ncu_dram_summary.py from a dram__bytes_{read,write}.sum pass over this very command), if committed
        traffic, traffic_note = None, "no ncu capture committed for this kernel at this size"
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_dram_%d.json" % n)))
            if dom in tr:
                traffic = tr[dom]["dram_bytes_per_launch"]
                traffic_note = tr[dom].get("note", "")
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms,
                "per_class_ms_per_step": {k: v[1] / args.steps for k, v in sorted(classes.items())},
                "passes": passes, "share_of_step": tot / args.steps / ms}

    cpu = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        ref = reference_binary_step(args.cpu_n)
        if ref is not None:
            cpu = {"value": args.cpu_n ** 3 / ref[0], "unit": "particles/s", "cores": threads, "kind": "reference",
                   "sample": "one IC (%.1f s): " % ref[0] + REF_SAMPLE % (args.cpu_n, threads)}
        else:
            cpu_port_step(64, threads)
            tcpu = cpu_port_step(args.cpu_n, threads)
            cpu = {"value": args.cpu_n ** 3 / tcpu, "unit": "particles/s", "cores": threads, "kind": "port",
                   "sample": "one %d^3 IC (%.1f s) through the CPU oracle port: C/OpenMP draw+covariance, pocketfft with "
                             "%d workers (reference binary oracle/_ref/genetIC not built)" % (args.cpu_n, tcpu, threads)}

    line = {"metric": "IC particles/sec (Zeldovich %d^3)" % n, "value": n3 / (ms * 1e-3), "unit": "particles/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, "f64"), "output": "float32 pos+vel, component-major",
                       "l2": "inputs larger than L2 (each field %.2f GB)" % (fbytes / 1e9),
                       "e2e_host_buffers": "full component-major arrays" if cap == n3 else
                       "pinned RING of %d particles per component (6 x n^3 x 4 bytes = %.0f GB does not fit the pinned "
                       "budget): every byte is copied, the IC is not retained" % (cap, 6 * n3 * osz / 1e9),
                       "e2e_pinned_on_cpus": ("%d-%d (GPU-local NUMA node, NVML)" % (pinned_cpus[0], pinned_cpus[-1])) if pinned_cpus else "default"},
            "clocks": clocks,
            "e2e": None if args.no_e2e else {"value": n3 / e2e_s, "unit": "particles/s", "h2d_bytes_per_step": tab_bytes,
                                             "d2h_bytes_per_step": 6 * n3 * osz, "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "checksum": checksum, "cufft_crosscheck": crosscheck}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cufft"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("GX_BENCH_N", "2048")))
    ap.add_argument("--dtype", default=os.environ.get("GX_BENCH_DTYPE", "f64"), choices=["f64", "f32"],
                    help="field precision: f64 (default) or f32 = the reference's -f build (BASELINE config 5: 4096^3 on 8 GPUs)")
    ap.add_argument("--cpu-n", type=int, default=512)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl in ("reference", "reference-cufft"):
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
