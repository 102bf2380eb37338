"""Input validation pre-pass on the device (Sasktran2::validate_input_atmosphere,
cpp/lib/engine/engine.cpp:481-540): bad extinction / single-scatter albedo values are refused with the reference's
messages instead of producing NaN radiances."""
import numpy as np
import pytest

import sasktran2_b200 as sk
from sasktran2_b200 import _lib, scenarios


def _engine(sc, mode=None):
    cfg = sk.Config()
    cfg.num_streams = sc.nstr
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    if mode is not None:
        cfg.input_validation_mode = mode
    geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp), sk.GeometryType(sc.geotype))
    view = sk.ViewingGeometry()
    for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
        view.add_ray(sk.GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
    return cfg, geo, view, sk.Engine(cfg, geo, view)


@pytest.mark.gpu
@pytest.mark.parametrize("field,value,needle", [
    ("total_extinction", np.nan, "total extinction contains non-finite"),
    ("total_extinction", -1e-6, "total extinction contains values less than 0"),
    ("ssa", np.inf, "single scatter albedo contains non-finite"),
    ("ssa", -0.1, "single scatter albedo contains values less than 0"),
    ("ssa", 1.5, "single scatter albedo contains values greater than 1"),
])
def test_cuda_input_validation_refuses_bad_values(field, value, needle):
    sc = scenarios.small_wf_case(nstr=4, nlayers=6, nwavel=5, nlos=2)
    sc.mappings = {}
    cfg, geo, view, eng = _engine(sc)
    atm = sk.Atmosphere.from_scenario(sc, geo, cfg, calculate_derivatives=False)
    getattr(atm.storage, field)[3, 2] = value
    with pytest.raises(_lib.SasktranError) as err:
        eng.calculate_radiance(atm)
    assert needle in str(err.value)
    # the engine stays usable after a refused call
    getattr(atm.storage, field)[3, 2] = getattr(sc, field)[3, 2]
    rad = eng.calculate_radiance(atm)["radiance"]
    assert np.all(np.isfinite(rad)) and np.all(rad > 0)


@pytest.mark.gpu
def test_cuda_input_validation_can_be_disabled():
    sc = scenarios.small_wf_case(nstr=4, nlayers=6, nwavel=3, nlos=1)
    sc.mappings = {}
    cfg, geo, view, eng = _engine(sc, mode=2)   # InputValidationMode::disabled
    atm = sk.Atmosphere.from_scenario(sc, geo, cfg, calculate_derivatives=False)
    atm.storage.ssa[2, 1] = 1.0 + 1e-12         # harmless excursion the validation would refuse
    rad = eng.calculate_radiance(atm)["radiance"]
    assert np.all(np.isfinite(rad))
