// genetic_b200_ic <paramfile> [device] -- runs the single-level hot path of a genetIC paramfile on a B200 through the
// C++ host layer (include/genetic_b200.hpp): the reference's `genetIC paramfile.txt` for the commands that layer covers
// (main.cpp:174-260 of the reference: read the paramfile, dispatch each command, exit(1) with the message on error).
#include <genetic_b200.hpp>

#include <cstdio>
#include <cstdlib>
#include <iostream>

int main(int argc, char **argv) {
  if (argc < 2) {
    std::cerr << "Usage: genetic_b200_ic paramfile [device]" << std::endl;
    return 1;
  }
  try {
    gx::check(gx_set_device(argc > 2 ? std::atoi(argv[2]) : 0));
    gx::ic::ICGenerator<double> generator;
    generator.runParamFile(argv[1]);
    if (!generator.haveOutput) std::cerr << "Warning: the paramfile did not contain `done`; nothing was generated" << std::endl;
    else
      std::cout << "Wrote " << generator.n * generator.n * generator.n << " particles (" << gx_launch_count()
                << " kernel launches)" << std::endl;
  } catch (std::exception &e) {
    std::cerr << "Error: " << e.what() << std::endl;  // tools/parser.hpp:202-207 behaviour
    return 1;
  }
  return 0;
}
