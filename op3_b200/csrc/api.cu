// Library-wide state of the C ABI: version, thread-local error string, launch counter.
#include <stdarg.h>
#include "common.cuh"

namespace op3 {
static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace op3

extern "C" const char* op3_version(void) { return "op3_b200 0.1.0 (sm_100a)"; }
extern "C" const char* op3_last_error(void) { return op3::g_err; }
extern "C" unsigned long long op3_launch_count(void) { return op3::g_launch_count; }
