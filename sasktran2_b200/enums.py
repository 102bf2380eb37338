"""Enumerations with the integer values of the reference C ABI (cpp/include/sasktran2/config.h:41-71,
cpp/include/c_api/geometry.h:10-13)."""
from enum import IntEnum


class MultipleScatterSource(IntEnum):
    DiscreteOrdinates = 0
    SuccessiveOrdersLegacy = 1
    TwoStream = 2
    NoSource = 3
    SuccessiveOrders = 4


class SingleScatterSource(IntEnum):
    Exact = 0
    Table = 1
    DiscreteOrdinates = 2
    NoSource = 3


class EmissionSource(IntEnum):
    Standard = 0
    NoSource = 1
    DiscreteOrdinates = 2
    VolumeEmissionRate = 3
    TwoStream = 4


class InterpolationMethod(IntEnum):
    ShellInterpolation = 0
    LinearInterpolation = 1
    LowerInterpolation = 2


class GeometryType(IntEnum):
    PlaneParallel = 0
    PseudoSpherical = 1
    Spherical = 2
    Ellipsoidal = 3


class ThreadingModel(IntEnum):
    Wavelength = 0
    Source = 1


class WeightingFunctionPrecision(IntEnum):
    Full = 0
    Reduced = 1
    Limited = 2
