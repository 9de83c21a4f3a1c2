// Error plumbing, version and launch accounting for libb200rl.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace b200rl {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace b200rl

extern "C" {

const char* b200rl_last_error(void) { return b200rl::g_err; }
int b200rl_version(void) { return 100; }
int64_t b200rl_launch_count(void) { return b200rl::g_launches.load(); }

}  // extern "C"
