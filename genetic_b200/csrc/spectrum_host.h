// Host side of the power spectrum: spline set-up and overall normalisation, then a device table.
#pragma once
#include "spectral.cuh"
struct gx_spectrum;
gx::SpecDev gx_spectrum_device_view(const gx_spectrum *s);
// cosmology::PowerSpectrum::getPowerSpectrumNormalizationForGrid (camb.hpp:127-139), in the field's type
double gx_spectrum_grid_norm(const gx_spectrum *s, double grid_length, int n);
