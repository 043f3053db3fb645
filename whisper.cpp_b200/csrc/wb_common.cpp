// wb_common.cpp -- error text, launch accounting, logging sink.
#include "wb_common.h"
#include <atomic>
#include <cstring>
#include <mutex>

namespace wb {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char * fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    logf(LOG_ERROR, "%s\n", g_err);
}
const char * last_error() { return g_err; }

void     count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
uint64_t launch_count()           { return g_launches.load(std::memory_order_relaxed); }

// log sink: installed by whisper_log_set (wb_api.cpp); signature mirrors ggml_log_callback
typedef void (*log_cb_t)(int level, const char * text, void * user);
static void default_log(int level, const char * text, void *) { (void) level; fputs(text, stderr); fflush(stderr); }
static log_cb_t g_log_cb = default_log;
static void *   g_log_ud = nullptr;
void set_log_sink(log_cb_t cb, void * ud) { g_log_cb = cb ? cb : default_log; g_log_ud = ud; }

void logf(int level, const char * fmt, ...) {
    char buf[2048];
    va_list ap; va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_log_cb(level, buf, g_log_ud);
}

} // namespace wb
