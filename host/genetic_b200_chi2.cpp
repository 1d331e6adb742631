// genetic_b200_chi2 <levels.txt> <seed> -- the reference's `chi2` command on a multi-level grid through the C++ host
// layer: one line per level "simsize n cellsize x0 y0 z0" (the zoom placement is control-plane logic and comes from the
// caller), white noise drawn in the real-space order (one gsl stream across the levels), then
// MultiLevelField::getChi2 (multilevelfield.hpp:461-487) printed as ICGenerator::getFieldChi2 prints it (ic.hpp:1592-1597).
#include <genetic_b200.hpp>

#include <cstdlib>
#include <iomanip>
#include <iostream>

int main(int argc, char **argv) {
  if (argc < 3) {
    std::cerr << "Usage: genetic_b200_chi2 levels.txt seed" << std::endl;
    return 1;
  }
  try {
    gx::check(gx_set_device(0));
    gx::multilevelgrid::MultiLevelGrid<double> context;
    std::ifstream f(argv[1]);
    if (!f.is_open()) throw std::runtime_error(std::string("File ") + argv[1] + " not found");
    double simsize, dx, x0, y0, z0;
    size_t n;
    while (f >> simsize >> n >> dx >> x0 >> y0 >> z0)
      context.addLevel(std::make_shared<gx::grids::Grid<double>>(simsize, n, dx, x0, y0, z0));
    if (context.getNumLevels() == 0) throw std::runtime_error("no levels");
    gx::fields::MultiLevelField<double> field(context);
    gx::fields::RandomFieldGenerator<double> rfg(field.getFields());
    rfg.seed(std::strtoul(argv[2], nullptr, 10));
    rfg.draw();
    const double chi2 = field.getChi2();
    std::cout << "Calculated chi^2 = " << std::setprecision(10) << chi2 << " (dof = " << context.getNumDof() << ")" << std::endl;
  } catch (std::exception &e) {
    std::cerr << "Error: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
