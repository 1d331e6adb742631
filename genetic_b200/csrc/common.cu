#include "gx_common.cuh"

namespace gx {
static thread_local std::string g_last_error;
unsigned long long g_launch_count = 0;
void set_error(const std::string &msg) { g_last_error = msg; }
int fail(const std::string &msg) { g_last_error = msg; return 1; }
const char *last_error_cstr() { return g_last_error.c_str(); }
}  // namespace gx
