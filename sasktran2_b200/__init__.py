"""sasktran2_b200 — B200-native discrete-ordinates radiance engine behind the SASKTRAN2 API.

Python host layer mirroring the reference's `sasktran2` package for this path:
    Engine(config, model_geometry, viewing_geo).calculate_radiance(atmosphere)
Everything numeric happens in libsasktran2_b200.so (hand-written CUDA for sm_100a behind the reference's
`sk_*` C ABI, include/sasktran2_b200.h).  There is no CPU fallback.
"""
from .enums import (EmissionSource, GeometryType, InterpolationMethod, MultipleScatterSource, SingleScatterSource, ThreadingModel,
                    WeightingFunctionPrecision)
from ._lib import SasktranError, LibraryMissing
from .config import Config
from .geometry import Geometry1D
from .viewinggeo import GroundViewingSolar, TangentAltitudeSolar, ViewingGeometry
from .atmosphere import Atmosphere
from .engine import Engine
from . import scenarios


def engine_for_scenario(sc, num_threads: int = 1, do_backprop: bool = True, input_validation: bool = True):
    """Build (config, geometry, viewing geometry, engine, atmosphere) for a synthetic Scenario."""
    cfg = Config()
    cfg.num_streams = sc.nstr
    cfg.num_stokes = 1
    cfg.multiple_scatter_source = MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = SingleScatterSource.DiscreteOrdinates
    cfg.do_backprop = do_backprop
    if not input_validation:
        cfg.input_validation_mode = 2   # InputValidationMode::disabled
    geo = Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, InterpolationMethod(sc.interp),
                     GeometryType(sc.geotype))
    view = ViewingGeometry()
    for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
        view.add_ray(GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
    eng = Engine(cfg, geo, view)
    atm = Atmosphere.from_scenario(sc, geo, cfg)
    return cfg, geo, view, eng, atm
