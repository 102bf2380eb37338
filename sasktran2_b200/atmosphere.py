"""Mirror of the low-level part of sasktran2.Atmosphere (src/sasktran2/atmosphere.py): the storage arrays that
cross the C ABI, the Lambertian surface and the derivative mappings.  Constituents / climatologies / optical
property databases are upstream of the solve and out of scope."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _view(ptr, shape):
    n = int(np.prod(shape))
    arr = np.ctypeslib.as_array(ptr, shape=(n,))
    return arr.reshape(shape, order="F")


class DerivativeMapping:
    """View of one named mapping owned by the storage (cpp/include/sasktran2/derivative_mapping.h:82-300).
    Arrays are Fortran-ordered views of library-owned memory, allocated on first access like upstream
    (cpp/c_api/deriv_mapping.cpp:25-90)."""

    def __init__(self, handle, nleg, nloc, nwavel, name):
        self._h = handle
        self._shape2 = (nloc, nwavel)
        self._shape3 = (nleg, nloc, nwavel)
        self.name = name

    def __del__(self):
        try:
            _lib.lib().sk_deriv_mapping_destroy(self._h)
        except Exception:
            pass

    def _get(self, which, shape):
        p = _lib.c_double_p()
        _lib.check(getattr(_lib.lib(), f"sk_deriv_mapping_get_{which}")(self._h, C.byref(p)), which)
        return _view(p, shape)

    @property
    def d_ssa(self):
        return self._get("d_ssa", self._shape2)

    @property
    def d_extinction(self):
        return self._get("d_extinction", self._shape2)

    @property
    def scat_factor(self):
        return self._get("scat_factor", self._shape2)

    @property
    def d_leg_coeff(self):
        return self._get("d_legendre", self._shape3)

    @property
    def is_scattering_derivative(self) -> bool:
        v = C.c_int(0)
        _lib.check(_lib.lib().sk_deriv_mapping_is_scattering_derivative(self._h, C.byref(v)))
        return bool(v.value)

    @property
    def num_output(self) -> int:
        v = C.c_int(0)
        _lib.check(_lib.lib().sk_deriv_mapping_get_num_output(self._h, C.byref(v)))
        return v.value

    @property
    def interp_dim(self) -> str:
        p = C.c_char_p()
        _lib.check(_lib.lib().sk_deriv_mapping_get_interp_dim(self._h, C.byref(p)))
        return p.value.decode()

    @interp_dim.setter
    def interp_dim(self, name: str):
        _lib.check(_lib.lib().sk_deriv_mapping_set_interp_dim(self._h, name.encode()))

    @property
    def assign_name(self) -> str:
        p = C.c_char_p()
        _lib.check(_lib.lib().sk_deriv_mapping_get_assign_name(self._h, C.byref(p)))
        return p.value.decode()

    @assign_name.setter
    def assign_name(self, name: str):
        _lib.check(_lib.lib().sk_deriv_mapping_set_assign_name(self._h, name.encode()))

    @property
    def interpolator(self):
        p = _lib.c_double_p()
        d1, d2 = C.c_int(0), C.c_int(0)
        _lib.check(_lib.lib().sk_deriv_mapping_get_interpolator(self._h, C.byref(p), C.byref(d1), C.byref(d2)))
        if not p:
            return None
        return _view(p, (d1.value, d2.value))

    @interpolator.setter
    def interpolator(self, mat):
        if mat is None:
            _lib.check(_lib.lib().sk_deriv_mapping_clear_interpolator(self._h))
            return
        m = np.asfortranarray(mat, dtype=np.float64)
        _lib.check(_lib.lib().sk_deriv_mapping_set_interpolator(self._h, _lib.dptr(m), m.shape[0], m.shape[1]))


class AtmosphereStorage:
    def __init__(self, nloc: int, nwavel: int, nleg: int):
        self.ssa = np.zeros((nloc, nwavel), order="F")
        self.total_extinction = np.zeros((nloc, nwavel), order="F")
        self.emission_source = np.zeros((nloc, nwavel), order="F")
        self.leg_coeff = np.zeros((nleg, nloc, nwavel), order="F")
        self.solar_irradiance = np.ones(nwavel)
        self._nleg, self._nloc, self._nwavel = nleg, nloc, nwavel
        self._h = _lib.lib().sk_atmosphere_storage_create(nloc, nwavel, nleg, 1, _lib.dptr(self.ssa),
                                                          _lib.dptr(self.total_extinction),
                                                          _lib.dptr(self.emission_source), _lib.dptr(self.leg_coeff),
                                                          _lib.dptr(self.solar_irradiance))
        if not self._h:
            raise _lib.SasktranError(_lib.last_error())

    def __del__(self):
        try:
            _lib.lib().sk_atmosphere_storage_destroy(self._h)
        except Exception:
            pass

    def get_derivative_mapping(self, name: str) -> DerivativeMapping:
        h = C.c_void_p()
        _lib.check(_lib.lib().sk_atmosphere_storage_get_derivative_mapping(self._h, name.encode(), C.byref(h)))
        return DerivativeMapping(h, self._nleg, self._nloc, self._nwavel, name)

    @property
    def derivative_mapping_names(self):
        n = C.c_int(0)
        _lib.check(_lib.lib().sk_atmosphere_storage_get_num_derivative_mappings(self._h, C.byref(n)))
        out = []
        for i in range(n.value):
            p = C.c_char_p()
            _lib.check(_lib.lib().sk_atmosphere_storage_get_derivative_mapping_name(self._h, i, C.byref(p)))
            out.append(p.value.decode())
        return out

    def finalize_scattering_derivatives(self):
        _lib.check(_lib.lib().sk_atmosphere_storage_finalize_scattering_derivatives(self._h))

    def set_zero(self):
        """Zero the optical arrays, the derivative mappings and the delta-M state (sk_atmosphere_storage_set_zero)."""
        _lib.check(_lib.lib().sk_atmosphere_storage_set_zero(self._h))

    def _fingerprint(self):
        # strided sample of the arrays the delta-M pass rewrites in place (cheap even for GB-size spectra)
        return tuple(float(a.ravel(order="K")[::4099].sum()) for a in (self.ssa, self.total_extinction, self.leg_coeff))


class Surface:
    """Lambertian surface (sk_surface_create + sk_brdf_create_lambertian + sk_surface_set_brdf)."""

    def __init__(self, nwavel: int):
        self.albedo = np.zeros(nwavel)
        self.emission = np.zeros(nwavel)
        self._h = _lib.lib().sk_surface_create(nwavel, 1, _lib.dptr(self.emission))
        self._brdf = _lib.lib().sk_brdf_create_lambertian(1)
        _lib.check(_lib.lib().sk_surface_set_brdf(self._h, self._brdf, _lib.dptr(self.albedo)))
        self._nwavel = nwavel
        self._mapping_names = []

    def __del__(self):
        try:
            _lib.lib().sk_surface_destroy(self._h)
            _lib.lib().sk_brdf_destroy(self._brdf)
        except Exception:
            pass

    def use_modis(self, isotropic, volumetric, geometric):
        """Kernel-based MODIS BRDF (sk_brdf_create_modis; reference: sasktran2.constituent.MODIS): isotropic +
        Ross-thick volumetric + Li-sparse-R geometric kernel weights, scalars or [nwavel] arrays.  Radiances only."""
        self.brdf_args = np.zeros((3, self._nwavel), order="F")
        self.brdf_args[0], self.brdf_args[1], self.brdf_args[2] = isotropic, volumetric, geometric
        old = self._brdf
        self._brdf = _lib.lib().sk_brdf_create_modis(1)
        _lib.check(_lib.lib().sk_surface_set_brdf(self._h, self._brdf, _lib.dptr(self.brdf_args)))
        _lib.lib().sk_brdf_destroy(old)

    def use_snow_kokhanovsky(self, arg):
        """Snow BRDF of Kokhanovsky (sk_brdf_create_kokhanovsky; reference: sasktran2.constituent.SnowKokhanovsky):
        arg = (chi + M) / wavelength * L, scalar or [nwavel].  Radiances only."""
        self.brdf_args = np.zeros((1, self._nwavel), order="F")
        self.brdf_args[0] = arg
        old = self._brdf
        self._brdf = _lib.lib().sk_brdf_create_kokhanovsky(1)
        _lib.check(_lib.lib().sk_surface_set_brdf(self._h, self._brdf, _lib.dptr(self.brdf_args)))
        _lib.lib().sk_brdf_destroy(old)

    def enable_albedo_derivative(self, name: str = "wf_albedo"):
        """Registers a surface mapping with d_brdf = 1 (d radiance / d albedo)."""
        h = C.c_void_p()
        _lib.check(_lib.lib().sk_surface_get_derivative_mapping(self._h, name.encode(), C.byref(h)))
        p = _lib.c_double_p()
        _lib.check(_lib.lib().sk_surface_deriv_mapping_get_d_brdf(h, C.byref(p)))
        np.ctypeslib.as_array(p, shape=(self._nwavel,))[:] = 1.0
        _lib.lib().sk_surface_deriv_mapping_destroy(h)
        if name not in self._mapping_names:
            self._mapping_names.append(name)


    def enable_brdf_argument_derivative(self, name: str, arg: int, num_args: int = 3):
        """Registers a surface mapping with d_brdf[:, arg] = 1: d radiance / d (BRDF argument `arg`), e.g. the weight of
        one MODIS kernel (upstream: surface.derivative_mappings[name].d_brdf is [nwavel, num_args]).  Call after
        use_modis()."""
        h = C.c_void_p()
        _lib.check(_lib.lib().sk_surface_get_derivative_mapping(self._h, name.encode(), C.byref(h)))
        p = _lib.c_double_p()
        _lib.check(_lib.lib().sk_surface_deriv_mapping_get_d_brdf(h, C.byref(p)))
        d = np.ctypeslib.as_array(p, shape=(num_args, self._nwavel))   # column-major [nwavel, num_args]
        d[:] = 0.0
        d[arg] = 1.0
        _lib.lib().sk_surface_deriv_mapping_destroy(h)
        if name not in self._mapping_names:
            self._mapping_names.append(name)


class Atmosphere:
    def __init__(self, model_geometry, config, wavelengths_nm=None, numwavel=None, calculate_derivatives=True,
                 num_legendre=None):
        if wavelengths_nm is not None:
            self.wavelengths_nm = np.atleast_1d(wavelengths_nm).astype(float)
            numwavel = self.wavelengths_nm.size
        else:
            self.wavelengths_nm = None
        if numwavel is None:
            raise ValueError("one of wavelengths_nm / numwavel is required")
        nloc = model_geometry.altitudes().size
        nleg = int(num_legendre) if num_legendre is not None else max(int(config.num_streams), 1)
        self.model_geometry = model_geometry
        self.storage = AtmosphereStorage(nloc, numwavel, nleg)
        self.surface = Surface(numwavel)
        self.calculate_derivatives = bool(calculate_derivatives)
        self._config = config
        self._applied_delta_m_order = None
        self._scaled_fingerprint = None
        self._h = None

    @property
    def num_wavel(self) -> int:
        return self.storage._nwavel

    def internal_object(self):
        if self._h is None:
            self._h = _lib.lib().sk_atmosphere_create(self.storage._h, self.surface._h,
                                                      int(self.calculate_derivatives), 0)
            if not self._h:
                raise _lib.SasktranError(_lib.last_error())
        self.storage.finalize_scattering_derivatives()
        # delta-M scaling is applied by the atmosphere, once, with order = num_streams
        # (src/sasktran2/atmosphere.py:846-856): in place on the storage arrays and the derivative mappings
        if self._config is not None and self._config.delta_m_scaling:
            if self._applied_delta_m_order is None:
                if self._config.num_streams != self.storage._nleg:
                    _lib.check(_lib.lib().sk_atmosphere_apply_delta_m_scaling(self._h, int(self._config.num_streams)))
                    self._applied_delta_m_order = int(self._config.num_streams)
                    self._scaled_fingerprint = self.storage._fingerprint()
            elif self._scaled_fingerprint != self.storage._fingerprint():
                # upstream rebuilds the storage from its constituents and rescales on every internal_object()
                # (src/sasktran2/atmosphere.py:655-662, 846-856); here the caller owns the arrays, so refilling them
                # without zero_storage() would pair unscaled optics with the stale truncation fractions
                raise _lib.SasktranError("storage arrays were modified after delta-M scaling was applied to them: call "
                                         "atmosphere.zero_storage() before refilling ssa / total_extinction / leg_coeff")
        return self._h

    def zero_storage(self) -> None:
        """Reset storage, mappings and surface to zero for the next fill (sasktran2.Atmosphere._zero_storage,
        src/sasktran2/atmosphere.py:655-662); delta-M scaling is applied again by the next internal_object()."""
        self.storage.set_zero()
        self.surface.albedo[:] = 0.0
        self._applied_delta_m_order = None
        self._scaled_fingerprint = None

    def __del__(self):
        try:
            if self._h:
                _lib.lib().sk_atmosphere_destroy(self._h)
        except Exception:
            pass

    @classmethod
    def from_scenario(cls, scenario, model_geometry, config, calculate_derivatives=None, total_wavelengths=None,
                      wavelength_start=0):
        """Fill an atmosphere from a sasktran2_b200.scenarios.Scenario (synthetic inputs).

        With `total_wavelengths` the storage is sized for a full spectrum of which `scenario` is the block starting at
        `wavelength_start` (one rank of a wavelength-sharded run holds the caller's array shapes but only fills - and
        only ever reads - its own block; untouched pages of the zero-initialised arrays cost nothing)."""
        wf = bool(scenario.mappings) if calculate_derivatives is None else calculate_derivatives
        nw = scenario.nwavel if total_wavelengths is None else int(total_wavelengths)
        sl = slice(wavelength_start, wavelength_start + scenario.nwavel)
        atm = cls(model_geometry, config, numwavel=nw, calculate_derivatives=wf,
                  num_legendre=scenario.leg_coeff.shape[0])
        atm.storage.ssa[:, sl] = scenario.ssa
        atm.storage.total_extinction[:, sl] = scenario.total_extinction
        atm.storage.leg_coeff[:, :, sl] = scenario.leg_coeff
        atm.storage.solar_irradiance[sl] = scenario.solar_irradiance
        atm.surface.albedo[sl] = scenario.albedo
        if wf:
            for name, mp in scenario.mappings.items():
                m = atm.storage.get_derivative_mapping(name)
                m.d_extinction[:, sl] = mp["d_extinction"]
                m.d_ssa[:, sl] = mp["d_ssa"]
                if "d_legendre" in mp:
                    m.d_leg_coeff[:, :, sl] = mp["d_legendre"]
                    m.scat_factor[:, sl] = mp["scat_factor"]
                if mp.get("interpolator") is not None:
                    m.interpolator = mp["interpolator"]
        return atm
