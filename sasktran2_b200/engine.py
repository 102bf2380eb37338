"""Mirror of sasktran2.Engine (src/sasktran2/engine.py:72-166, 474-560): same constructor and
calculate_radiance(atmosphere) contract, results as numpy arrays with the reference's dimension order
(radiance [wavelength, los, stokes]; weighting functions [<interp_dim>, wavelength, los, stokes])."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

TIMING_NAMES = ("h2d", "optics", "layer", "bvp", "radiance", "d2h", "kernels_total", "wf", "wf_adjoint", "wf_layer",
                "wf_chain", "wf_map", "limb_source", "limb_integrate")


class Result(dict):
    """dict of numpy arrays; `dims` gives the dimension names of every entry."""

    def __init__(self):
        super().__init__()
        self.dims = {}

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class Engine:
    def __init__(self, config, model_geometry, viewing_geo):
        self._config = config
        self._geometry = model_geometry
        self._viewing_geometry = viewing_geo
        self._engine = _lib.lib().sk_engine_create(config._config, model_geometry._geometry,
                                                   viewing_geo._viewing_geometry)
        if not self._engine:
            raise _lib.SasktranError(f"sk_engine_create failed: {_lib.last_error()}")
        self._keepalive = None
        self._staged_out = None
        # Output arrays live in page-locked host memory (D2H at PCIe speed).  With reuse_output_buffers = True the
        # arrays of the previous call with the same shapes are handed out again instead of fresh ones - for
        # pipelines that consume a result before asking for the next one (bench.py's e2e loop).
        self.reuse_output_buffers = False
        self._pool = {}
        self._given = None

    def __del__(self):
        try:
            if self._staged_out is not None:
                _lib.lib().sk_output_destroy(self._staged_out[0])
            _lib.lib().sk_engine_destroy(self._engine)
        except Exception:
            pass

    # ---- the reference call -------------------------------------------------------------------------
    def _buffer(self, key, shape):
        if self._given is not None and key in self._given:   # caller-owned result arrays (e.g. shared memory)
            buf = self._given[key]
            assert buf.shape == tuple(shape) and buf.flags["C_CONTIGUOUS"] and buf.dtype == np.float64, (key, buf.shape, shape)
            return buf
        if self.reuse_output_buffers:
            buf = self._pool.get(key)
            if buf is None or buf.shape != tuple(shape):
                buf = self._pool[key] = _lib.pinned_empty(shape)
            return buf
        buf = _lib.pinned_empty(shape)
        buf[...] = 0.0
        return buf

    def result_shapes(self, atmosphere) -> dict:
        """Shapes of the arrays a call fills, keyed like the `buffers` argument of calculate_radiance / stage."""
        nw, nlos = atmosphere.num_wavel, self._viewing_geometry.num_rays
        shapes = {"radiance": (nw, nlos, 1)}
        if atmosphere.calculate_derivatives and self._config.wf_enabled:
            for name in atmosphere.storage.derivative_mapping_names:
                shapes["wf:" + name] = (atmosphere.storage.get_derivative_mapping(name).num_output, nw, nlos, 1)
            for name in atmosphere.surface._mapping_names:
                shapes["surf:" + name] = (1, nw, nlos, 1)
        return shapes

    def _make_output(self, atmosphere, radiance_buffer=None, buffers=None):
        self._given = buffers
        if buffers is not None and radiance_buffer is None:
            radiance_buffer = buffers.get("radiance")
        nw = atmosphere.num_wavel
        nlos = self._viewing_geometry.num_rays
        rad = radiance_buffer if radiance_buffer is not None else self._buffer("radiance", (nw, nlos, 1))
        assert rad.shape == (nw, nlos, 1) and rad.flags["C_CONTIGUOUS"]
        out = _lib.lib().sk_output_create(_lib.dptr(rad), nw * nlos, 1, None, 0)
        res = Result()
        res["radiance"] = rad
        res.dims["radiance"] = ("wavelength", "los", "stokes")
        if atmosphere.calculate_derivatives and self._config.wf_enabled:
            for name in atmosphere.storage.derivative_mapping_names:
                m = atmosphere.storage.get_derivative_mapping(name)
                nout = m.num_output
                buf = self._buffer("wf:" + name, (nout, nw, nlos, 1))
                _lib.check(_lib.lib().sk_output_assign_derivative_memory(out, name.encode(), _lib.dptr(buf),
                                                                         nw * nlos, 1, nout))
                key = m.assign_name or name
                res[key] = buf
                res.dims[key] = (m.interp_dim, "wavelength", "los", "stokes")
            for name in atmosphere.surface._mapping_names:
                buf = self._buffer("surf:" + name, (1, nw, nlos, 1))
                _lib.check(_lib.lib().sk_output_assign_surface_derivative_memory(out, name.encode(),
                                                                                 _lib.dptr(buf), nw * nlos, 1))
                res[name] = buf[0]
                res.dims[name] = ("wavelength", "los", "stokes")
                res["_" + name + "_buf"] = buf
        return out, res

    def calculate_radiance(self, atmosphere, radiance_buffer=None, buffers=None, wavelength_block=None) -> Result:
        """sasktran2.Engine.calculate_radiance.  `buffers` (optional): caller-owned result arrays keyed as in
        result_shapes().  `wavelength_block = (start, count)` solves only that block of the spectrum and writes it at
        its place in the full-size result arrays: the reference's Rayon entry points
        (sk_engine_calculate_radiance(only_initialize = 1) + sk_engine_calculate_radiance_block_thread,
        rust/sasktran2-rs/src/bindings/engine.rs:314-395), which is also how the ranks of a wavelength-sharded run
        fill one shared result."""
        out, res = self._make_output(atmosphere, radiance_buffer, buffers)
        try:
            if wavelength_block is None:
                rc = _lib.lib().sk_engine_calculate_radiance(self._engine, atmosphere.internal_object(), out, 0)
                _lib.check(rc, "sk_engine_calculate_radiance")
            else:
                rc = _lib.lib().sk_engine_calculate_radiance(self._engine, atmosphere.internal_object(), out, 1)
                _lib.check(rc, "sk_engine_calculate_radiance(only_initialize)")
                rc = _lib.lib().sk_engine_calculate_radiance_block_thread(self._engine, out, int(wavelength_block[0]),
                                                                          int(wavelength_block[1]), 0)
                _lib.check(rc, "sk_engine_calculate_radiance_block_thread")
            if self._config.output_los_optical_depth and self._geometry.geometry_type == 2:
                # Output::los_optical_depth (src/sasktran2/engine.py: "los_optical_depth" [wavelength, los])
                od = C.POINTER(C.c_double)()
                _lib.check(_lib.lib().sk_output_get_los_optical_depth(out, C.byref(od)), "sk_output_get_los_optical_depth")
                nw, nlos = res["radiance"].shape[:2]
                res["los_optical_depth"] = np.ctypeslib.as_array(od, shape=(nlos, nw)).T.copy()
                res.dims["los_optical_depth"] = ("wavelength", "los")
        finally:
            _lib.lib().sk_output_destroy(out)
        if atmosphere.wavelengths_nm is not None:
            res["wavelength"] = atmosphere.wavelengths_nm
        return res

    # ---- device-resident extension (bench / pipelines that keep the atmosphere on the GPU) -----------
    def stage(self, atmosphere, wavelength_start: int = 0, wavelength_count: int = -1, radiance_buffer=None,
              buffers=None) -> None:
        """Copy the atmosphere (and its derivative mappings when weighting functions are on) to the device."""
        self._keepalive = atmosphere
        if self._staged_out is not None:
            _lib.lib().sk_output_destroy(self._staged_out[0])
        self._staged_out = self._make_output(atmosphere, radiance_buffer, buffers)
        _lib.check(_lib.lib().sk_b200_engine_stage_atmosphere(self._engine, atmosphere.internal_object(),
                                                              self._staged_out[0], wavelength_start, wavelength_count),
                   "stage")

    def solve_staged(self) -> None:
        _lib.check(_lib.lib().sk_b200_engine_solve_staged(self._engine), "solve_staged")

    def fetch(self, atmosphere=None, radiance_buffer=None) -> Result:
        """Copy the results of the last staged solve into the output buffers created by stage()."""
        out, res = self._staged_out
        _lib.check(_lib.lib().sk_b200_engine_fetch_output(self._engine, out), "fetch")
        return res

    def gather(self, block_starts, block_counts, nw_total: int, root: int = 0):
        """Collective of a wavelength-sharded run (sk_b200_comm_init on every rank first): after solve_staged() on every
        rank, collect all blocks on `root` over NCCL into the result arrays given to stage().  Returns
        (result on root / None elsewhere, (nccl_ms, d2h_ms) of this rank)."""
        out, res = self._staged_out
        starts = (C.c_int * len(block_starts))(*[int(v) for v in block_starts])
        counts = (C.c_int * len(block_counts))(*[int(v) for v in block_counts])
        ms = np.zeros(2)
        _lib.check(_lib.lib().sk_b200_engine_gather_output(self._engine, out, int(root), starts, counts, int(nw_total),
                                                           _lib.dptr(ms)), "gather")
        return res, (float(ms[0]), float(ms[1]))

    def timings_ms(self) -> dict:
        buf = np.zeros(len(TIMING_NAMES))
        _lib.check(_lib.lib().sk_b200_engine_get_timings(self._engine, _lib.dptr(buf), buf.size))
        return dict(zip(TIMING_NAMES, buf.tolist()))

    def kernel_launches(self) -> int:
        return int(_lib.lib().sk_b200_engine_kernel_launches(self._engine))

    def info(self) -> dict:
        a, c, mb = C.c_int(0), C.c_int(0), C.c_double(0)
        _lib.check(_lib.lib().sk_b200_engine_info(self._engine, C.byref(a), C.byref(c), C.byref(mb)))
        return {"num_azimuth": a.value, "chunk_wavelengths": c.value, "workspace_mb_per_wavelength": mb.value}

    def set_workspace_gb(self, gb: float) -> None:
        _lib.check(_lib.lib().sk_b200_engine_set_workspace_gb(self._engine, float(gb)))
