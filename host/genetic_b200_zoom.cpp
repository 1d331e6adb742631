// genetic_b200_zoom paramfile.txt [-f] [--layout-only out.bin]
// The `genetIC paramfile.txt` equivalent for zoom ICs through the C++ host layer (include/genetic_b200_zoom.hpp): geometry,
// particle mappers, multi-level particle generation on the device, tipsy / gadget2 / gadget3 output.
// --layout-only stops at `done` before anything touches the device and dumps the mapper's layout (N; level, cell, mass, eps,
// gadget type, gas flag per particle) so the host logic can be checked on a machine without a GPU (tests/test_cpp_zoom.py);
// --particles-from file.bin takes the particles' positions / velocities from a file instead of the device and runs the
// writers only (same purpose).
#include "../include/genetic_b200_zoom.hpp"

#include <cstdio>
#include <iostream>

template <typename T> static int run(const std::string &paramfile, const std::string &layoutFile, const std::string &particlesFile) {
  gx::ic::ZoomICGenerator<T> gen;
  gen.layoutOnly = !layoutFile.empty();
  gen.particlesFile = particlesFile;
  gen.runParamFile(paramfile);
  if (!gen.layoutOnly) return 0;
  if (!gen.pMapper) gen.pMapper = gen.mapper();
  const auto &lay = gen.pMapper->layout();
  const auto layers = gen.pMapper->layers();
  gx::particle::MappedParticles<T, float> P(*gen.pMapper, nullptr, gen.cosmo);
  const long long N = (long long) lay.size();
  std::vector<int32_t> level(N);
  for (long long i = 0; i < N; ++i) level[i] = (int32_t) layers.at(lay.layer[i])->level;
  std::vector<double> mass(P.mass.begin(), P.mass.end()), eps(P.eps.begin(), P.eps.end());
  FILE *f = std::fopen(layoutFile.c_str(), "wb");
  if (!f) throw std::runtime_error("Could not open " + layoutFile);
  std::fwrite(&N, sizeof(N), 1, f);
  std::fwrite(level.data(), 4, N, f);
  std::fwrite(lay.cell.data(), 8, N, f);
  std::fwrite(mass.data(), 8, N, f);
  std::fwrite(eps.data(), 8, N, f);
  std::fwrite(P.gadgetType.data(), 4, N, f);
  std::fwrite(P.isGas.data(), 1, N, f);
  std::fclose(f);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 2) {
    std::cerr << "Usage: genetic_b200_zoom paramfile [-f] [--layout-only out.bin] [--particles-from in.bin]" << std::endl;
    return 1;
  }
  bool useFloat = false;
  std::string layoutFile, particlesFile;
  for (int i = 2; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-f") useFloat = true;
    else if (a == "--layout-only" && i + 1 < argc) layoutFile = argv[++i];
    else if (a == "--particles-from" && i + 1 < argc) particlesFile = argv[++i];
    else { std::cerr << "Error: unknown option " << a << std::endl; return 1; }
  }
  try {
    return useFloat ? run<float>(argv[1], layoutFile, particlesFile) : run<double>(argv[1], layoutFile, particlesFile);
  } catch (std::exception &e) {
    std::cerr << "Error: " << e.what() << std::endl;   // Dispatch::run_loop: log + exit(1), parser.hpp:202-207
    return 1;
  }
}
