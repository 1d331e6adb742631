// TEST INFRASTRUCTURE: stand-in for <highfive/highfive.hpp> (oracle/shims/README.md). HDF5 is not installed; the classes
// the reference's gadgethdf / swift writers name exist so that the translation unit compiles, and throw if such an output
// format is actually requested.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

namespace HighFive {
struct Exception : std::runtime_error { using std::runtime_error::runtime_error; };
[[noreturn]] inline void unavailable() { throw Exception("HDF5 output is not available in this build (oracle/shims)"); }
struct SilenceHDF5 {};
struct DataSpace {
  enum DataspaceType { dataspace_scalar, dataspace_null };
  DataSpace(size_t) {}
  DataSpace(size_t, size_t) {}
  DataSpace(DataspaceType) {}
};
struct Attribute {
  template <typename T> void write(const T &) { unavailable(); }
};
struct DataSet {
  template <typename T> void write_raw(const T *) { unavailable(); }
};
struct Group {
  template <typename T> DataSet createDataSet(const std::string &, const DataSpace &) { unavailable(); }
  template <typename T> Attribute createAttribute(const std::string &, const DataSpace &) { unavailable(); }
};
struct File {
  enum { Truncate = 1 };
  File(const std::string &, int) { unavailable(); }
  Group getGroup(const std::string &) { unavailable(); }
  Group createGroup(const std::string &) { unavailable(); }
};
}  // namespace HighFive
