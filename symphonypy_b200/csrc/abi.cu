// Library-level exports: version, last error, launch counter.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace sym {
static thread_local char g_err[512] = {0};
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace sym

extern "C" {
int sym_version(void) { return 100; }
const char *sym_last_error(void) { return sym::g_err; }
int64_t sym_launch_count(void) { return sym::g_launches.load(std::memory_order_relaxed); }
int sym_compiled_arch(void) { return 100; }
}
