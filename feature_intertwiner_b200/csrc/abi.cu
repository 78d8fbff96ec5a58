// Error bookkeeping of the C ABI (include/fi_b200.h): thread-local status + message, never exit().
#include <stdarg.h>
#include <string.h>

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
