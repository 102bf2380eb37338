"""Mirror of sasktran2.ViewingGeometry / GroundViewingSolar (src/sasktran2/viewinggeo/wrappers.py)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from . import _lib


@dataclass
class GroundViewingSolar:
    """A line of sight that ends at the ground, defined by solar angles at the ground point
    (cpp/lib/viewinggeometry/groundviewing.cpp:7-60)."""
    cos_sza: float
    relative_azimuth: float
    cos_viewing_zenith: float
    observer_altitude_m: float


@dataclass
class TangentAltitudeSolar:
    """A limb line of sight through a tangent point given by its altitude and solar angles (spherical geometry only,
    cpp/lib/viewinggeometry/tangentaltitudesolar.cpp:5-62)."""
    tangent_altitude_m: float
    relative_azimuth: float
    observer_altitude_m: float
    cos_sza: float


class ViewingGeometry:
    def __init__(self):
        self._viewing_geometry = _lib.lib().sk_viewing_geometry_create()
        self.observer_rays = []
        self.flux_observers = []

    def __del__(self):
        try:
            _lib.lib().sk_viewing_geometry_destroy(self._viewing_geometry)
        except Exception:
            pass

    def add_ray(self, ray) -> None:
        if isinstance(ray, TangentAltitudeSolar):
            _lib.check(_lib.lib().sk_viewing_geometry_add_tangent_altitude_solar(
                self._viewing_geometry, float(ray.tangent_altitude_m), float(ray.relative_azimuth),
                float(ray.observer_altitude_m), float(ray.cos_sza)), "add_tangent_altitude_solar")
        elif isinstance(ray, GroundViewingSolar):
            _lib.lib().sk_viewing_geometry_add_ground_viewing_solar(
                self._viewing_geometry, float(ray.cos_sza), float(ray.relative_azimuth), float(ray.observer_altitude_m),
                float(ray.cos_viewing_zenith))
        else:
            raise NotImplementedError("the B200 path supports GroundViewingSolar and TangentAltitudeSolar rays")
        self.observer_rays.append(ray)

    @property
    def num_rays(self) -> int:
        n = C.c_int(0)
        _lib.check(_lib.lib().sk_viewing_geometry_num_rays(self._viewing_geometry, C.byref(n)), "num_rays")
        return n.value
