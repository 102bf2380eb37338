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
        if not isinstance(ray, GroundViewingSolar):
            raise NotImplementedError("the B200 DO path supports GroundViewingSolar rays only")
        _lib.lib().sk_viewing_geometry_add_ground_viewing_solar(
            self._viewing_geometry, float(ray.cos_sza), float(ray.relative_azimuth), float(ray.observer_altitude_m),
            float(ray.cos_viewing_zenith))
        self.observer_rays.append(ray)

    @property
    def num_rays(self) -> int:
        n = C.c_int(0)
        _lib.check(_lib.lib().sk_viewing_geometry_num_rays(self._viewing_geometry, C.byref(n)), "num_rays")
        return n.value
