// Error string, ABI version and launch accounting of liboisat.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace oisat {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace oisat

extern "C" {

const char* oisat_last_error(void) { return oisat::g_err; }

int oisat_abi_version(void) { return OISAT_ABI_VERSION; }

int64_t oisat_launch_count(void) { return oisat::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
