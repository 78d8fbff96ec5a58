// Error bookkeeping of the C ABI (include/fi_b200.h): thread-local status + message, never exit().
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "fi_common.cuh"

namespace fi {

static thread_local int g_status = FI_OK;
static thread_local char g_message[512] = "";

void set_error(int status, const char *fmt, ...) {
    g_status = status;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_message, sizeof(g_message), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;   // relaxed counter; exactness under concurrent callers is not required
void count_launch() { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); }

// ---- options: one int each, initialised once from the environment (std::call_once), changed by fi_set_option ----
static int g_options[FI_OPT_COUNT];
static std::once_flag g_options_once;
int g_deterministic_init = 0;          // FI_BWD=exact: initial value of the deterministic switch (roi_align_bwd.cu)

static void init_options() {
    for (int i = 0; i < FI_OPT_COUNT; ++i) g_options[i] = 0;
    const char *v;
    if ((v = getenv("FI_BWD_ACC")) && v[0] == 's') g_options[FI_OPT_BWD_FORM] = 1;
    if ((v = getenv("FI_BWD_TILE")) && v[0] == 'f') g_options[FI_OPT_BWD_FORM] = 2;
    if ((v = getenv("FI_BWD")) && v[0] == 'r') g_options[FI_OPT_BWD_FORM] = 3;
    if ((v = getenv("FI_BWD")) && v[0] == 'e') g_deterministic_init = 1;
    if ((v = getenv("FI_TILE")) && !strcmp(v, "4x4")) g_options[FI_OPT_TILE_SHAPE] = 1;
    if ((v = getenv("FI_TILE")) && !strcmp(v, "2x8")) g_options[FI_OPT_TILE_SHAPE] = 2;
    if ((v = getenv("FI_SINKHORN_GENERIC"))) g_options[FI_OPT_SINKHORN_GENERIC] = v[0] == '2' ? 2 : 1;
    if ((v = getenv("FI_PIX_CFG")) && v[0] >= '0' && v[0] <= '3') g_options[FI_OPT_PIX_CFG] = v[0] - '0';
    if ((v = getenv("FI_FWD_FORM")) && v[0] >= '0' && v[0] <= '6') g_options[FI_OPT_FWD_FORM] = v[0] - '0';
    if ((v = getenv("FI_FWD_CHUNK")) && v[0] >= '0' && v[0] <= '6') g_options[FI_OPT_FWD_CHUNK] = v[0] - '0';
    if ((v = getenv("FI_FWD_SCHED")) && v[0] >= '0' && v[0] <= '2') g_options[FI_OPT_FWD_SCHED] = v[0] - '0';
    if ((v = getenv("FI_FWD_PAIR")) && v[0] >= '0' && v[0] <= '2') g_options[FI_OPT_FWD_PAIR] = v[0] - '0';
    g_options[FI_OPT_PIX_GROUP] = 1;
    if ((v = getenv("FI_PIX_GROUP")) && v[0] >= '0' && v[0] <= '7') g_options[FI_OPT_PIX_GROUP] = v[0] - '0';
}

int option(int key) {
    std::call_once(g_options_once, init_options);
    return __atomic_load_n(&g_options[key], __ATOMIC_RELAXED);
}

int ok() {
    g_status = FI_OK;
    g_message[0] = 0;
    return FI_OK;
}

}  // namespace fi

FI_API int fi_abi_version(void) { return 1; }
FI_API const char *fi_last_error(void) { return fi::g_message; }
FI_API int fi_last_status(void) { return fi::g_status; }
FI_API unsigned long long fi_kernel_launches(void) { return __atomic_load_n(&fi::g_launches, __ATOMIC_RELAXED); }

FI_API int fi_get_option(int opt) {
    if (opt < 0 || opt >= FI_OPT_COUNT) { fi::set_error(FI_ERR_INVALID, "fi_get_option: unknown option %d", opt); return FI_ERR_INVALID; }
    return fi::option(opt);
}
FI_API int fi_set_option(int opt, int value) {
    static const int limit[FI_OPT_COUNT] = {3, 2, 6, 2, 3, 7, 6, 2, 2};
    if (opt < 0 || opt >= FI_OPT_COUNT || value < 0 || value > limit[opt]) {
        fi::set_error(FI_ERR_INVALID, "fi_set_option: option %d value %d", opt, value);
        return FI_ERR_INVALID;
    }
    const int old = fi::option(opt);
    __atomic_store_n(&fi::g_options[opt], value, __ATOMIC_RELAXED);
    return old;
}
