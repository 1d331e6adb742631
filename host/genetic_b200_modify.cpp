// genetIC paramfile driver with the modification commands (include/genetic_b200_modif.hpp): geometry, seeds, cosmology,
// calculate / modify / filtering_scale / chi2 / splice / dump_grid / done -- the C++ counterpart of `python -m genetic_b200`
// for the north-star classes (ModificationManager, linear + variance modifications, spliceOneLevel / conjugate gradient).
//   host/genetic_b200_modify paramfile.txt [-f]
#include "../include/genetic_b200_modif.hpp"
#include <cstring>

template <typename T> static int run(const char *paramfile) {
  gx::ic::ModifICGenerator<T> gen;
  gen.runParamFile(paramfile);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s paramfile [-f]\n", argv[0]); return 2; }
  const bool useFloat = argc > 2 && std::strcmp(argv[2], "-f") == 0;
  try {
    return useFloat ? run<float>(argv[1]) : run<double>(argv[1]);
  } catch (std::exception &e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}
