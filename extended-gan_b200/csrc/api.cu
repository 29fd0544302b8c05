// Library identity and per-thread error message for the C ABI (include/cgat_b200.h).
#include "common.cuh"

namespace cgat {
char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
}  // namespace cgat

extern "C" const char* cgat_version(void) { return "cgat_b200 0.1 (sm_100a)"; }
extern "C" const char* cgat_last_error(void) { return cgat::last_error_buf(); }
