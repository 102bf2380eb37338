"""Mirror of sasktran2.Config (src/sasktran2/config.py) over the C ABI's sk_config_* accessors."""
from __future__ import annotations

import ctypes as C

from . import _lib
from .enums import EmissionSource, MultipleScatterSource, SingleScatterSource, ThreadingModel, WeightingFunctionPrecision


def _int_prop(name, wrap=int, doc=None):
    def getter(self):
        v = C.c_int(0)
        _lib.check(getattr(_lib.lib(), f"sk_config_get_{name}")(self._config, C.byref(v)), f"get {name}")
        return wrap(v.value)

    def setter(self, value):
        _lib.check(getattr(_lib.lib(), f"sk_config_set_{name}")(self._config, int(value)), f"set {name}")

    return property(getter, setter, doc=doc)


class Config:
    """Defaults follow the reference (cpp/lib/config/config.cpp:5-33): 16 streams, single scatter Exact,
    multiple scatter NoSource — the discrete-ordinates path needs both sources set to DiscreteOrdinates
    (or single scatter NoSource)."""

    def __init__(self):
        self._config = _lib.lib().sk_config_create()

    def __del__(self):
        try:
            _lib.lib().sk_config_destroy(self._config)
        except Exception:
            pass

    num_stokes = _int_prop("num_stokes")
    num_streams = _int_prop("num_streams")
    num_threads = _int_prop("num_threads")
    wavelength_batch_size = _int_prop("wavelength_batch_size")
    multiple_scatter_source = _int_prop("multiple_scatter_source", MultipleScatterSource)
    single_scatter_source = _int_prop("single_scatter_source", SingleScatterSource)
    emission_source = _int_prop("emission_source", EmissionSource)
    threading_model = _int_prop("threading_model", ThreadingModel)
    num_forced_azimuth = _int_prop("num_do_forced_azimuth")
    do_backprop = _int_prop("do_backprop", bool)
    num_sza = _int_prop("num_do_sza")
    num_singlescatter_moments = _int_prop("num_singlescatter_moments")
    delta_m_scaling = _int_prop("apply_delta_scaling", bool)
    solar_refraction = _int_prop("solar_refraction", bool)
    wf_enabled = _int_prop("wf_enabled", bool)
    wf_precision = _int_prop("wf_precision", WeightingFunctionPrecision)
    log_level = _int_prop("log_level")
    output_los_optical_depth = _int_prop("output_los_optical_depth", bool)
    input_validation_mode = _int_prop("input_validation_mode")
