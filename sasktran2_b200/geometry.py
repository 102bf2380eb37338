"""Mirror of sasktran2.Geometry1D (src/sasktran2/geometry.py:13-56)."""
from __future__ import annotations

import numpy as np

from . import _lib
from .enums import GeometryType, InterpolationMethod


class Geometry1D:
    def __init__(self, cos_sza: float, solar_azimuth: float, earth_radius_m: float, altitude_grid_m,
                 interpolation_method: InterpolationMethod = InterpolationMethod.LinearInterpolation,
                 geometry_type: GeometryType = GeometryType.Spherical):
        self._alt = np.ascontiguousarray(np.atleast_1d(altitude_grid_m), dtype=np.float64)
        self.cos_sza = float(cos_sza)
        self.solar_azimuth = float(solar_azimuth)
        self.earth_radius_m = float(earth_radius_m)
        self.interpolation_method = InterpolationMethod(interpolation_method)
        self.geometry_type = GeometryType(geometry_type)
        self._geometry = _lib.lib().sk_geometry1d_create(self.cos_sza, self.solar_azimuth, self.earth_radius_m,
                                                         _lib.dptr(self._alt), self._alt.size,
                                                         int(self.interpolation_method), int(self.geometry_type))
        if not self._geometry:
            raise _lib.SasktranError(_lib.last_error())

    def __del__(self):
        try:
            _lib.lib().sk_geometry1d_destroy(self._geometry)
        except Exception:
            pass

    def altitudes(self) -> np.ndarray:
        n = _lib.lib().sk_geometry1d_get_num_altitudes(self._geometry)
        out = np.zeros(n)
        _lib.check(_lib.lib().sk_geometry1d_get_altitudes(self._geometry, _lib.dptr(out)), "get_altitudes")
        return out
