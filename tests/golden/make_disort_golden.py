#!/usr/bin/env python
"""Extract the DISORT-verified known-answer tables of the reference's legacy scalar DO tests into
tests/golden/disort_scalar.json.

Source (read-only, only present in the authoring container):
  /root/reference/cpp/lib/tests/sktran_disco/legacy/test_scalar.cpp   (20 active cases x 35 LOS, 12 digits)
Only DATA (inputs + expected radiances + tolerance) is extracted; no reference code is copied.
The driver semantics these numbers assume (cpp/lib/tests/sktran_disco/test_util.cpp:26-115):
  plane-parallel, altitude grid 0..nlyr (unit thickness => extinction == layer optical depth),
  `lower` interpolation, layer l (0 = TOA) stored at grid index nlyr-1-l, HG moments (2l+1) g^l,
  GroundViewingSolar(csz, -saz + az, coszen, top + 1), SS = MS = DiscreteOrdinates, result * direct.
Run:  python tests/golden/make_disort_golden.py
"""
import json
import re
import sys
from pathlib import Path

SRC = Path("/root/reference/cpp/lib/tests/sktran_disco/legacy/test_scalar.cpp")
OUT = Path(__file__).with_name("disort_scalar.json")

EPS = {"SKDO_FPC_EPS": 1e-8, "SKDO_FPC_EPS_LOW_PRECISION": 1e-6}


def numbers(s):
    return [float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", s)]


def parse_layers(block):
    # TestLayerSpecHG({od, ssa, g})
    out = []
    for m in re.finditer(r"TestLayerSpecHG\(\s*\{([^}]*)\}\s*\)", block):
        od, ssa, g = numbers(m.group(1))
        out.append([od, ssa, g])
    return out


def parse_los(block):
    out = []
    for m in re.finditer(r"\{\s*([0-9.]+)\s*,\s*([^}]*?)\}", block):
        cz = float(m.group(1))
        az_expr = m.group(2).strip()
        mm = re.match(r"(\d+)\s*\*\s*PI\s*/\s*(\d+)", az_expr)
        if mm:
            az = ["pi_frac", int(mm.group(1)), int(mm.group(2))]
        else:
            az = ["rad", float(az_expr)]
        out.append([cz, az])
    return out


def main():
    text = SRC.read_text()
    # strip block comments (the BRDF case body is commented out in the reference)
    text_nc = re.sub(r"/\*.*?\*/", "", text, flags=re.S)

    head = text_nc[: text_nc.index("TEST_CASE")]
    m = re.search(r"default_atmo\s*=\s*std::vector<TestLayerSpecHG>\(\{(.*?)\}\);", head, re.S)
    default_atmo = parse_layers(m.group(1))
    m = re.search(r"default_los\s*=\s*\{(.*?)\};", head, re.S)
    default_los = parse_los(m.group(1))
    default_sun = {"csz": 0.8, "saz": 0.0, "direct": 1.0}

    cases = []
    parts = re.split(r"TEST_CASE\(", text_nc)[1:]
    for part in parts:
        name = re.match(r'\s*"([^"]*)"', part).group(1)
        if "correct_radiances = {" not in part:
            continue  # "Scalar Boundary Conditions" has no table (self-consistency test)
        mrad = re.search(r"correct_radiances\s*=\s*\{(.*?)\};", part, re.S)
        rad = numbers(mrad.group(1))
        mtc = re.search(r"TestCase<1>\s+testcase\(\s*(\d+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,\s*([0-9.]+)\s*,\s*(\w+)\s*,", part, re.S)
        if mtc is None:
            continue  # BRDF case (non-Lambertian callable) - body commented out upstream
        nstr = int(mtc.group(1))
        sun_name, atmo_name, albedo, los_name = mtc.group(2), mtc.group(3), float(mtc.group(4)), mtc.group(5)
        if sun_name == "default_sun":
            sun = dict(default_sun)
        else:
            ms = re.search(sun_name + r"\s*=\s*\{\s*([0-9.]+)\s*,\s*([0-9.]+)\s*,\s*\{\s*([0-9.]+)\s*,\s*([0-9.]+)\s*\}\s*\}", part)
            sun = {"csz": float(ms.group(1)), "saz": float(ms.group(2)), "direct": float(ms.group(3))}
        if atmo_name == "default_atmo":
            layers = default_atmo
        else:
            ma = re.search(atmo_name + r"\s*=\s*std::vector<TestLayerSpecHG>\(\s*\{(.*?)\}\s*\);", part, re.S)
            layers = parse_layers(ma.group(1))
        if los_name == "default_los":
            los = default_los
        else:
            ml = re.search(los_name + r"\s*=\s*\{(.*?)\};", part, re.S)
            los = parse_los(ml.group(1))
        mt = re.search(r"REQUIRE\(diff\s*<\s*(\w+)\)", part)
        tol = EPS[mt.group(1)]
        assert len(rad) == len(los), (name, len(rad), len(los))
        cases.append({"name": name, "nstr": nstr, "sun": sun, "albedo": albedo,
                      "layers_od_ssa_g": layers, "los_coszen_az": los,
                      "radiance": rad, "abs_tol": tol})
    doc = {
        "source": "usask-arg/sasktran2 cpp/lib/tests/sktran_disco/legacy/test_scalar.cpp (DISORT-verified tables)",
        "driver": "cpp/lib/tests/sktran_disco/test_util.cpp:26-115",
        "cases": cases,
    }
    OUT.write_text(json.dumps(doc, indent=1))
    print(f"wrote {OUT} with {len(cases)} cases", file=sys.stderr)


if __name__ == "__main__":
    main()
