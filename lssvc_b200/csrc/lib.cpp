// Library-level state of the C-ABI: last error text and the kernel-launch counter.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/lssvc_b200.h"

namespace lssvc {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace lssvc

extern "C" {
int32_t lssvc_abi_version(void) { return 4; }
const char *lssvc_last_error(void) { return lssvc::g_error; }
int64_t lssvc_launch_count(void) { return lssvc::g_launches.load(std::memory_order_relaxed); }
void lssvc_launch_count_add(int64_t n) { lssvc::count_launch(static_cast<int>(n)); }
}
