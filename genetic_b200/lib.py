"""ctypes binding of libgenetic_b200.so -- one prototype per function declared in include/genetic_b200.h."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "libgenetic_b200.so")
F32, F64 = 0, 1
LAYOUT_REAL, LAYOUT_FOURIER = 0, 1
NOISE_REAL, NOISE_FOURIER_SERIAL, NOISE_FOURIER_PARALLEL = 0, 1, 2
FILTER_CONST, FILTER_FERMI_LP, FILTER_FERMI_HP_COV, FILTER_FERMI_HP = 0, 1, 2, 3


class GxError(RuntimeError):
    pass


class IcParams(C.Structure):
    _fields_ = [("n", C.c_int), ("dtype", C.c_int), ("out_dtype", C.c_int), ("noise_mode", C.c_int),
                ("reverse_order", C.c_int), ("fix_power", C.c_int), ("seed", C.c_ulong), ("grid_length", C.c_double),
                ("cell_size", C.c_double), ("domain_size", C.c_double), ("offset_lower", C.c_double * 3),
                ("vel_ratio", C.c_double)]


class GridDesc(C.Structure):
    _fields_ = [("n", C.c_int), ("cell_size", C.c_double), ("offset_lower", C.c_double * 3), ("domain_size", C.c_double)]


class FilterFactor(C.Structure):
    _fields_ = [("kind", C.c_int), ("kcut", C.c_double), ("temperature", C.c_double), ("power", C.c_int)]


vp, ci, cd, cl, cll, csz = C.c_void_p, C.c_int, C.c_double, C.c_long, C.c_longlong, C.c_size_t
cul, cull = C.c_ulong, C.c_ulonglong
P = C.POINTER

# name -> (restype, argtypes); int-returning functions are status codes
PROTOTYPES = {
    "gx_last_error": (C.c_char_p, []),
    "gx_version": (ci, []),
    "gx_device_count": (ci, [P(ci)]),
    "gx_set_device": (ci, [ci]),
    "gx_synchronize": (ci, [vp]),
    "gx_workspace_release": (ci, []),
    "gx_launch_count": (cull, []),
    "gx_field_bytes": (csz, [ci, ci]),
    "gx_malloc": (ci, [P(vp), csz]),
    "gx_free": (ci, [vp]),
    "gx_memset_zero": (ci, [vp, csz, vp]),
    "gx_host_alloc_pinned": (ci, [P(vp), csz]),
    "gx_host_free_pinned": (ci, [vp]),
    "gx_ipc_get_handle": (ci, [vp, vp]),
    "gx_ipc_open": (ci, [vp, P(vp)]),
    "gx_ipc_close": (ci, [vp]),
    "gx_memcpy_d2d": (ci, [vp, vp, csz, vp]),
    "gx_memcpy_h2d": (ci, [vp, vp, csz, vp]),
    "gx_memcpy_d2h": (ci, [vp, vp, csz, vp]),
    "gx_field_upload": (ci, [vp, vp, ci, ci, ci, vp]),
    "gx_field_download": (ci, [vp, vp, ci, ci, ci, vp]),
    "gx_fft_forward": (ci, [vp, ci, ci, vp]),
    "gx_fft_inverse": (ci, [vp, ci, ci, vp]),
    "gx_mirror": (ci, [vp, ci, ci, vp]),
    "gx_noise_draw": (ci, [vp, ci, ci, cul, ci, ci, P(cull), vp]),
    "gx_rng_raw": (ci, [cul, cull, cl, vp]),
    "gx_rng_gaussians": (ci, [cul, cl, vp, P(cull)]),
    "gx_spectrum_create_camb": (ci, [vp, vp, vp, ci, cd, cd, cd, cd, cd, ci, P(vp)]),
    "gx_spectrum_create_powerlaw": (ci, [cd, cd, ci, P(vp)]),
    "gx_spectrum_destroy": (ci, [vp]),
    "gx_spectrum_amplitude": (cd, [vp]),
    "gx_spectrum_eval": (cd, [vp, cd]),
    "gx_covariance_field": (ci, [vp, ci, ci, cd, vp, vp]),
    "gx_power_spectrum_bins": (ci, [vp, vp, ci, ci, cd, ci, vp, vp]),
    "gx_apply_transfer": (ci, [vp, ci, ci, cd, vp, cd, vp]),
    "gx_apply_transfer_field": (ci, [vp, vp, ci, ci, cd, vp]),
    "gx_apply_filter": (ci, [vp, ci, ci, cd, P(FilterFactor), ci, vp]),
    "gx_unit_variance": (ci, [vp, ci, ci, vp]),
    "gx_reverse_small_k": (ci, [vp, ci, ci, cd, cd, vp]),
    "gx_scale": (ci, [vp, ci, ci, cd, vp]),
    "gx_axpy": (ci, [vp, vp, ci, ci, cd, vp]),
    "gx_axpy_fourier": (ci, [vp, vp, ci, ci, cd, vp]),
    "gx_mul_mask": (ci, [vp, vp, ci, ci, vp]),
    "gx_dot_fourier": (ci, [vp, vp, ci, ci, P(cd), vp]),
    "gx_dot_real": (ci, [vp, vp, ci, ci, P(cd), vp]),
    "gx_get_mode": (ci, [vp, ci, ci, ci, ci, ci, P(cd), P(cd)]),
    "gx_set_mode": (ci, [vp, ci, ci, ci, ci, ci, cd, cd]),
    "gx_interp_add": (ci, [vp, P(GridDesc), vp, P(GridDesc), ci, vp]),
    "gx_deinterp_add": (ci, [vp, P(GridDesc), vp, P(GridDesc), ci, cd, vp]),
    "gx_zero_window": (ci, [vp, P(GridDesc), ci, P(GridDesc), ci, vp]),
    "gx_masked_sum": (ci, [vp, vp, ci, ci, P(cd), vp]),
    "gx_masked_shift_scale": (ci, [vp, vp, ci, ci, cd, cd, vp]),
    "gx_scale_inv_k2": (ci, [vp, ci, ci, cd, cd, ci, vp]),
    "gx_zeldovich_fourier": (ci, [vp, vp, ci, ci, cd, ci, vp]),
    "gx_zeldovich_offset": (ci, [vp, vp, ci, ci, cd, ci, vp]),
    "gx_emit_component": (ci, [vp, ci, ci, ci, cd, cd, cd, cd, vp, vp, ci, cl, cll, cll, vp]),
    "gx_emit_cells": (ci, [vp, ci, ci, ci, cd, cd, cd, cd, vp, cll, vp, vp, ci, cl, vp]),
    "gx_unigrid_ic": (ci, [P(IcParams), vp, vp, vp, vp, csz, vp, vp, csz, vp]),
    "gx_slab_bytes": (csz, [ci, ci, ci]),
    "gx_slab_noise_draw": (ci, [vp, ci, ci, ci, ci, cul, ci, ci, P(cull), vp]),
    "gx_slab_noise_fn_bytes": (csz, [ci, ci]),
    "gx_slab_noise_pass_a": (ci, [vp, ci, ci, ci, ci, cul, vp]),
    "gx_slab_noise_finish": (ci, [vp, vp, ci, ci, ci, ci, cul, ci, vp]),
    "gx_slab_noise_staged": (ci, [vp, ci, ci, ci, ci, cul, ci, P(vp), vp, cd, vp, csz, vp]),
    "gx_slab_zero_dc": (ci, [vp, ci, vp]),
    "gx_emit_checksum": (ci, [vp]),
    "gx_slab_unit_variance": (ci, [vp, ci, ci, ci, ci, vp]),
    "gx_slab_apply_transfer": (ci, [vp, ci, ci, ci, ci, cd, vp, cd, ci, vp]),
    "gx_slab_extract_planes": (ci, [vp, vp, ci, ci, ci, vp]),
    "gx_slab_fft_x_mirror": (ci, [vp, vp, vp, ci, ci, ci, ci, cd, ci, ci, vp]),
    "gx_slab_fft_x_mirror_push": (ci, [vp, vp, P(vp), ci, ci, ci, ci, cd, ci, ci, vp]),
    "gx_slab_fft_y_op": (ci, [vp, vp, ci, ci, ci, ci, cd, ci, vp]),
    "gx_slab_fft_x": (ci, [vp, ci, ci, ci, ci, ci, vp]),
    "gx_slab_fft_y": (ci, [vp, vp, ci, ci, ci, ci, ci, ci, vp]),
    "gx_slab_fft_z_r2c": (ci, [vp, ci, ci, ci, vp]),
    "gx_slab_fft_z_c2r": (ci, [vp, ci, ci, ci, vp]),
    "gx_slab_z_emit": (ci, [vp, ci, ci, ci, ci, ci, cd, cd, cd, cd, ci, ci, vp, vp, ci, vp]),
    "gx_timing_enable": (ci, [ci]),
    "gx_timing_report": (ci, [C.c_char_p, ci]),
}
NOT_STATUS = {"gx_timing_report", "gx_last_error", "gx_version", "gx_launch_count", "gx_field_bytes", "gx_slab_bytes", "gx_slab_noise_fn_bytes", "gx_spectrum_amplitude",
              "gx_spectrum_eval"}

_lib = None


class _Checked:
    """Wraps a status-returning entry point: non-zero -> GxError(gx_last_error()), the reference's
    std::runtime_error behaviour (tools/parser.hpp:202-207)."""

    def __init__(self, lib, name, fn):
        self._lib, self._name, self._fn = lib, name, fn

    def __call__(self, *args):
        rc = self._fn(*args)
        if rc != 0:
            raise GxError("%s: %s" % (self._name, self._lib.gx_last_error().decode()))
        return rc


class _Lib:
    def __init__(self, cdll):
        self.cdll = cdll
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(cdll, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn if name in NOT_STATUS else _Checked(self, name, fn))


def load():
    """Loads the CUDA library. Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise GxError("genetic_b200: %s is missing -- run `python -m genetic_b200.build` "
                          "(there is no CPU fallback)" % LIBPATH)
        _lib = _Lib(C.CDLL(LIBPATH))
    return _lib


class numa_local:
    """Context manager: run the body on the CPUs NVML reports as local to CUDA device `device`, then restore the affinity.
    Pinned host buffers allocated inside land on the GPU's NUMA node (first touch), which is what the PCIe copies of the
    end-to-end path want; outside, host work keeps all the cores. A box without NVML or with one NUMA node is left alone."""

    def __init__(self, device=0):
        self.device, self.saved, self.cpus = device, None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.device
            if vis:
                tok = vis.split(",")[self.device].strip()
                index = int(tok) if tok.isdigit() else self.device
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            have = os.sched_getaffinity(0)
            use = local & have
            if use and use != have:
                self.saved = have
                os.sched_setaffinity(0, use)
            self.cpus = sorted(use) if use else None
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False
