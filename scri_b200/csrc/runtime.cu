// Error reporting / bookkeeping entry points of libscrib200.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace scrib200 {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace scrib200

extern "C" {

int scrib200_version(void) { return 100; }

const char* scrib200_last_error(void) { return scrib200::g_error; }

int64_t scrib200_launch_count(void) { return scrib200::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
