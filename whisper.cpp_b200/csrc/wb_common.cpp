// wb_common.cpp -- error text, launch accounting, logging sink.
#include <nvtx3/nvToolsExt.h>
#include <cstdlib>
#include "wb_common.h"
#include <utility>
#include <map>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

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

static std::mutex g_cnt_mu; static double g_cnt[16] = { 0 };
void   counter_add(int idx, double v) { std::lock_guard<std::mutex> lk(g_cnt_mu); if (idx >= 0 && idx < 16) g_cnt[idx] += v; }
double counter_get(int idx) { std::lock_guard<std::mutex> lk(g_cnt_mu); return (idx >= 0 && idx < 16) ? g_cnt[idx] : 0.0; }

static std::atomic<uint64_t> g_h2d{0}, g_d2h{0};
void count_h2d(uint64_t n) { g_h2d.fetch_add(n, std::memory_order_relaxed); }
void count_d2h(uint64_t n) { g_d2h.fetch_add(n, std::memory_order_relaxed); }
uint64_t h2d_bytes() { return g_h2d.load(); }
uint64_t d2h_bytes() { return g_d2h.load(); }

// ---- profiling -------------------------------------------------------------------------------------------------------
struct ProfEntry { int cls; cudaEvent_t e0, e1; double bytes, flops; };
static std::atomic<bool> g_prof{false};
static std::mutex g_prof_mu;
static std::vector<ProfEntry> g_prof_entries;
void prof_enable(bool on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto & e : g_prof_entries) { cudaEventDestroy(e.e0); cudaEventDestroy(e.e1); }
    g_prof_entries.clear();
    g_prof.store(on);
}
bool prof_enabled() { return g_prof.load(std::memory_order_relaxed); }
ProfScope::ProfScope(int cls, cudaStream_t stream, double bytes, double flops) : st(stream) {
    if (!prof_enabled()) return;
    ProfEntry e; e.cls = cls; e.bytes = bytes; e.flops = flops;
    if (cudaEventCreate(&e.e0) != cudaSuccess || cudaEventCreate(&e.e1) != cudaSuccess) return;
    cudaEventRecord(e.e0, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    idx = (int) g_prof_entries.size();
    g_prof_entries.push_back(e);
}
ProfScope::~ProfScope() {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (idx < (int) g_prof_entries.size()) cudaEventRecord(g_prof_entries[idx].e1, st);
}
void prof_collect(double * ms, uint64_t * launches, double * bytes, double * flops) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int c = 0; c < PC_COUNT; ++c) { ms[c] = 0; launches[c] = 0; bytes[c] = 0; flops[c] = 0; }
    for (auto & e : g_prof_entries) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, e.e0, e.e1) != cudaSuccess) continue;
        ms[e.cls] += t; launches[e.cls]++; bytes[e.cls] += e.bytes; flops[e.cls] += e.flops;
    }
    if (const char * path = getenv("WB200_PROF_DUMP")) {         // every scope on its own line: class, ms, bytes, flops (launch order)
        if (FILE * f = fopen(path, "w")) {
            for (auto & e : g_prof_entries) {
                float t = 0.0f;
                if (cudaEventElapsedTime(&t, e.e0, e.e1) == cudaSuccess) fprintf(f, "%d %.6f %.0f %.0f\n", e.cls, t, e.bytes, e.flops);
            }
            fclose(f);
        }
    }
}

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

static bool nvtx_enabled() { static const bool on = [] { const char * e = getenv("WB200_NVTX"); return e && atoi(e) != 0; }(); return on; }
NvtxRange::NvtxRange(const char * name) : on(nvtx_enabled()) { if (on) nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { if (on) nvtxRangePop(); }

cudaError_t ensure_dyn_smem(const void * kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    size_t & have = done[std::make_pair(dev, kernel)];
    if (have >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

} // namespace wb

namespace wb {
// language codes in whisper id order (src/whisper.cpp:278-381)
extern const char * const g_lang_codes[100] = {
    "en","zh","de","es","ru","ko","fr","ja","pt","tr","pl","ca","nl","ar","sv","it","id","hi","fi","vi",
    "he","uk","el","ms","cs","ro","da","hu","ta","no","th","ur","hr","bg","lt","la","mi","ml","cy","sk",
    "te","fa","lv","bn","sr","az","sl","kn","et","mk","br","eu","is","hy","ne","mn","bs","kk","sq","sw",
    "gl","mr","pa","si","km","sn","yo","so","af","oc","ka","be","tg","sd","gu","am","yi","lo","uz","fo",
    "ht","ps","tk","nn","mt","sa","lb","my","bo","tl","mg","as","tt","haw","ln","ha","ba","jw","su","yue",
};
extern const char * const g_lang_names[100] = {
    "english","chinese","german","spanish","russian","korean","french","japanese","portuguese","turkish",
    "polish","catalan","dutch","arabic","swedish","italian","indonesian","hindi","finnish","vietnamese",
    "hebrew","ukrainian","greek","malay","czech","romanian","danish","hungarian","tamil","norwegian",
    "thai","urdu","croatian","bulgarian","lithuanian","latin","maori","malayalam","welsh","slovak",
    "telugu","persian","latvian","bengali","serbian","azerbaijani","slovenian","kannada","estonian","macedonian",
    "breton","basque","icelandic","armenian","nepali","mongolian","bosnian","kazakh","albanian","swahili",
    "galician","marathi","punjabi","sinhala","khmer","shona","yoruba","somali","afrikaans","occitan",
    "georgian","belarusian","tajik","sindhi","gujarati","amharic","yiddish","lao","uzbek","faroese",
    "haitian creole","pashto","turkmen","nynorsk","maltese","sanskrit","luxembourgish","myanmar","tibetan","tagalog",
    "malagasy","assamese","tatar","hawaiian","lingala","hausa","bashkir","javanese","sundanese","cantonese",
};
} // namespace wb
