#include "afs_common.cuh"

#include <atomic>

namespace afs {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

char *error_buffer() { return g_err; }

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count()
{
    // per device: a process may drive several GPUs (the plans carry their device)
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace afs

extern "C" {
const char *afs_last_error(void) { return afs::error_buffer(); }
const char *afs_version(void) { return "libafsync 0.1 sm_100a"; }
int64_t afs_launch_count(void) { return afs::g_launches.load(std::memory_order_relaxed); }
}
