"""ctypes binding of libsasktran2_b200.so (the C ABI declared in include/sasktran2_b200.h).

The library must be built in-tree (`python -c "import __graft_entry__ as g; g.build()"` or
`make -C sasktran2_b200/csrc`).  There is no CPU fallback: if the library is missing, import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libsasktran2_b200.so"

_lib = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class LibraryMissing(ImportError):
    pass


def _declare(lib):
    vp = C.c_void_p
    i = C.c_int
    d = C.c_double

    def f(name, res, *args):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = list(args)

    f("sk_config_create", vp)
    f("sk_config_destroy", None, vp)
    for n in ("num_stokes", "multiple_scatter_source", "single_scatter_source", "num_streams", "num_threads",
              "threading_model", "wavelength_batch_size", "num_singlescatter_moments", "apply_delta_scaling",
              "num_do_sza", "num_do_forced_azimuth", "do_backprop", "emission_source", "occultation_source",
              "solar_refraction", "wf_enabled", "wf_precision", "input_validation_mode", "log_level",
              "output_los_optical_depth"):
        f(f"sk_config_get_{n}", i, vp, c_int_p)
        f(f"sk_config_set_{n}", i, vp, i)
    f("sk_geometry1d_create", vp, d, d, d, c_double_p, i, i, i)
    f("sk_geometry1d_destroy", None, vp)
    f("sk_geometry1d_get_num_altitudes", i, vp)
    f("sk_geometry1d_get_altitudes", i, vp, c_double_p)
    f("sk_viewing_geometry_create", vp)
    f("sk_viewing_geometry_destroy", None, vp)
    f("sk_viewing_geometry_add_ground_viewing_solar", None, vp, d, d, d, d)
    f("sk_viewing_geometry_num_rays", i, vp, c_int_p)
    f("sk_viewing_geometry_num_flux_observers", i, vp, c_int_p)
    f("sk_atmosphere_storage_create", vp, i, i, i, i, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p)
    f("sk_atmosphere_storage_destroy", None, vp)
    f("sk_atmosphere_storage_get_derivative_mapping", i, vp, C.c_char_p, C.POINTER(vp))
    f("sk_atmosphere_storage_get_derivative_mapping_by_index", i, vp, i, C.POINTER(vp))
    f("sk_atmosphere_storage_get_num_derivative_mappings", i, vp, c_int_p)
    f("sk_atmosphere_storage_get_derivative_mapping_name", i, vp, i, C.POINTER(C.c_char_p))
    f("sk_atmosphere_storage_finalize_scattering_derivatives", i, vp)
    f("sk_atmosphere_storage_set_zero", i, vp)
    f("sk_atmosphere_create", vp, vp, vp, i, i)
    f("sk_atmosphere_destroy", None, vp)
    f("sk_atmosphere_apply_delta_m_scaling", i, vp, i)
    f("sk_surface_create", vp, i, i, c_double_p)
    f("sk_surface_destroy", None, vp)
    f("sk_surface_set_brdf", i, vp, vp, c_double_p)
    f("sk_surface_get_derivative_mapping", i, vp, C.c_char_p, C.POINTER(vp))
    f("sk_surface_get_num_derivative_mappings", i, vp, c_int_p)
    f("sk_surface_get_derivative_mapping_name", i, vp, i, C.POINTER(C.c_char_p))
    f("sk_surface_set_zero", i, vp)
    f("sk_brdf_create_lambertian", vp, i)
    f("sk_brdf_create_modis", vp, i)
    f("sk_brdf_create_kokhanovsky", vp, i)
    f("sk_brdf_get_num_deriv", i, vp, c_int_p)
    f("sk_brdf_get_num_args", i, vp, c_int_p)
    f("sk_brdf_destroy", None, vp)
    f("sk_deriv_mapping_destroy", i, vp)
    f("sk_deriv_mapping_set_zero", i, vp)
    for n in ("d_ssa", "d_extinction", "scat_factor", "d_legendre"):
        f(f"sk_deriv_mapping_get_{n}", i, vp, C.POINTER(c_double_p))
    f("sk_deriv_mapping_get_scat_deriv_index", i, vp, c_int_p)
    f("sk_deriv_mapping_set_scat_deriv_index", i, vp, i)
    for n in ("num_location", "num_wavel", "num_legendre", "num_output", "log_radiance_space"):
        f(f"sk_deriv_mapping_get_{n}", i, vp, c_int_p)
    f("sk_deriv_mapping_is_scattering_derivative", i, vp, c_int_p)
    f("sk_deriv_mapping_set_interp_dim", i, vp, C.c_char_p)
    f("sk_deriv_mapping_set_assign_name", i, vp, C.c_char_p)
    f("sk_deriv_mapping_set_log_radiance_space", i, vp, i)
    f("sk_deriv_mapping_get_assign_name", i, vp, C.POINTER(C.c_char_p))
    f("sk_deriv_mapping_get_interp_dim", i, vp, C.POINTER(C.c_char_p))
    f("sk_deriv_mapping_set_interpolator", i, vp, c_double_p, i, i)
    f("sk_deriv_mapping_clear_interpolator", i, vp)
    f("sk_deriv_mapping_get_interpolator", i, vp, C.POINTER(c_double_p), c_int_p, c_int_p)
    f("sk_surface_deriv_mapping_get_num_wavel", i, vp, c_int_p)
    f("sk_surface_deriv_mapping_get_num_brdf_args", i, vp, c_int_p)
    f("sk_surface_deriv_mapping_get_d_brdf", i, vp, C.POINTER(c_double_p))
    f("sk_surface_deriv_mapping_set_zero", i, vp)
    f("sk_surface_deriv_mapping_destroy", i, vp)
    f("sk_output_create", vp, c_double_p, i, i, c_double_p, i)
    f("sk_output_destroy", None, vp)
    f("sk_output_assign_derivative_memory", i, vp, C.c_char_p, c_double_p, i, i, i)
    f("sk_output_assign_surface_derivative_memory", i, vp, C.c_char_p, c_double_p, i, i)
    f("sk_engine_create", vp, vp, vp, vp)
    f("sk_engine_destroy", None, vp)
    f("sk_engine_calculate_radiance", i, vp, vp, vp, i)
    f("sk_engine_calculate_radiance_block_thread", i, vp, vp, i, i, i)
    f("sk_engine_effective_wavelength_batch_size", i, vp, i)
    f("sk_engine_supports_linearization", i, vp, i, c_int_p)
    f("sk_engine_linearization_backend", i, vp, i, c_int_p)
    f("sk_openmp_support_enabled", i)
    f("sk_b200_last_error", C.c_char_p)
    f("sk_b200_device_count", i)
    f("sk_b200_set_device", i, i)
    f("sk_b200_engine_stage_atmosphere", i, vp, vp, vp, i, i)
    f("sk_b200_engine_solve_staged", i, vp)
    f("sk_b200_engine_fetch_output", i, vp, vp)
    f("sk_b200_engine_get_timings", i, vp, c_double_p, i)
    f("sk_b200_engine_kernel_launches", C.c_longlong, vp)
    f("sk_b200_engine_info", i, vp, c_int_p, c_int_p, c_double_p)
    f("sk_b200_engine_set_workspace_gb", i, vp, d)
    f("sk_b200_measure_fp64_tflops", d)
    f("sk_b200_adjoint_reuses_factors", i, i, i)
    f("sk_b200_engine_debug_copy", C.c_longlong, vp, C.c_char_p, c_double_p, C.c_longlong)
    f("sk_b200_host_register", i, vp, C.c_size_t)
    f("sk_b200_host_unregister", i, vp)
    f("sk_b200_comm_unique_id", i, C.c_char_p, i)
    f("sk_b200_comm_init", i, C.c_char_p, i, i)
    f("sk_b200_comm_destroy", i)
    f("sk_b200_engine_gather_output", i, vp, vp, i, c_int_p, c_int_p, i, c_double_p)
    f("sk_viewing_geometry_add_tangent_altitude_solar", i, vp, d, d, d, d)
    f("sk_output_get_los_optical_depth", i, vp, C.POINTER(C.POINTER(C.c_double)))
    f("sk_b200_limb_plan_check", i, vp, vp, i, i, c_double_p, c_double_p, c_int_p, c_double_p)
    f("sk_b200_host_alloc", vp, C.c_size_t)
    f("sk_b200_host_free", None, vp)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise LibraryMissing(
                f"{LIB_PATH} is missing: build it with `make -C sasktran2_b200/csrc` (there is no CPU fallback)")
        _lib = C.CDLL(str(LIB_PATH))
        _declare(_lib)
    return _lib


def last_error() -> str:
    return lib().sk_b200_last_error().decode()


class SasktranError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise SasktranError(f"{what} failed with code {rc}: {last_error()}")


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


class _PinnedBlock:
    """Owns one sk_b200_host_alloc block; numpy views keep it alive through their .base chain."""

    def __init__(self, nbytes: int):
        self.ptr = lib().sk_b200_host_alloc(C.c_size_t(max(nbytes, 8)))
        if not self.ptr:
            raise MemoryError("sk_b200_host_alloc failed")
        self.nbytes = nbytes

    def __del__(self):
        try:
            lib().sk_b200_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape):
    """float64 C-ordered array in page-locked host memory (plain memory when no CUDA device is present)."""
    import numpy as np

    n = int(np.prod(shape))
    blk = _PinnedBlock(8 * n)
    buf = (C.c_double * max(n, 1)).from_address(blk.ptr)
    buf._owner = blk  # keeps the block alive as long as any view of `buf` lives
    return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
