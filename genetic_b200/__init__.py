"""genetic_b200 -- B200-native implementation of genetIC's Gaussian-random-field + Zeldovich hot path.

The compute lives in genetic_b200/libgenetic_b200.so (hand-written sm_100a CUDA behind the C ABI declared in
include/genetic_b200.h). This package is the thin Python host layer used by the tests and the bench; it
mirrors the reference's class surface (grids::Grid, fields::Field, fields::RandomFieldGenerator,
cosmology::CAMB, particle::ZeldovichParticleGenerator). There is NO CPU fallback: every operation raises if
the CUDA library is missing or no device is present.
"""
from .lib import load, GxError, F32, F64  # noqa: F401
from .host import (Grid, Field, CAMB, PowerLawPowerSpectrum, CosmologicalParameters, RandomFieldGenerator,  # noqa: F401
                   ZeldovichParticleGenerator, generate_unigrid_ic, read_camb_table, LowPassFermiFilter,
                   FilterProduct)
