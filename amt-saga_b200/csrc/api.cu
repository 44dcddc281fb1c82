// Error channel, version and launch counter of the saga_b200 C ABI.
#include <cstring>

#include "saga_common.cuh"

namespace saga {
std::atomic<int64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = "";
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace saga

extern "C" const char* saga_last_error_string(void) { return saga::err_buf(); }
extern "C" int saga_abi_version(void) { return 2; }   // 2: + saga_short_window_batch_exec, saga_gather_frames_exec, saga_cqt_frames_exec
extern "C" int64_t saga_launch_count(void) { return saga::g_launches.load(); }
