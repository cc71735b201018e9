// Debug aid in place of compute-sanitizer (closed on the GPU pool this was developed on): with UPOL_GUARD=1 in the
// environment every device buffer of the library is allocated with a 256-byte guard zone on either side, filled
// with a byte pattern; upol_debug_check_guards() reads all guard zones back and counts the damaged ones.  An
// out-of-bounds store next to any buffer (a column split too many, a row past the padding, a mailbox slot too far)
// shows up as a non-zero count after the test sequence (tests/test_gpu_guards.py).  Off by default: dmalloc/dfree
// are then plain cudaMalloc/cudaFree.  Buffers exported through CUDA IPC cannot carry guards (the peer maps the
// base address), so the peer-memory exchange refuses to start in guard mode.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace upol {

namespace {
constexpr size_t kGuardBytes = 256;
constexpr unsigned char kGuardByte = 0xA5;
struct GuardEntry { char *base; size_t bytes; };
std::mutex g_guard_mutex;
std::unordered_map<void *, GuardEntry> g_guarded;
}  // namespace

bool guard_enabled() {
    static const bool on = [] { const char *e = getenv("UPOL_GUARD"); return e != nullptr && atoi(e) != 0; }();
    return on;
}

cudaError_t dmalloc(void **p, size_t bytes) {
    if (!guard_enabled()) return cudaMalloc(p, bytes);
    *p = nullptr;
    const size_t padded = (bytes + 255) / 256 * 256;
    char *base = nullptr;
    cudaError_t e = cudaMalloc((void **)&base, padded + 2 * kGuardBytes);
    if (e != cudaSuccess) return e;
    e = cudaMemset(base, kGuardByte, padded + 2 * kGuardBytes);
    if (e != cudaSuccess) { cudaFree(base); return e; }
    *p = base + kGuardBytes;
    std::lock_guard<std::mutex> lock(g_guard_mutex);
    g_guarded[*p] = GuardEntry{base, bytes};
    return cudaSuccess;
}

cudaError_t dfree(void *p) {
    if (p == nullptr) return cudaSuccess;
    if (guard_enabled()) {
        std::lock_guard<std::mutex> lock(g_guard_mutex);
        auto it = g_guarded.find(p);
        if (it != g_guarded.end()) {
            char *base = it->second.base;
            g_guarded.erase(it);
            return cudaFree(base);
        }
    }
    return cudaFree(p);
}

}  // namespace upol

// Number of damaged guard zones over all live buffers (0 = clean; -1 = guard mode is off; < -1 = CUDA error).
extern "C" int upol_debug_check_guards(void) {
    using namespace upol;
    if (!guard_enabled()) return -1;
    if (cudaDeviceSynchronize() != cudaSuccess) return UPOL_ERR_CUDA - 1;
    std::lock_guard<std::mutex> lock(g_guard_mutex);
    int bad = 0;
    std::vector<unsigned char> h(kGuardBytes);
    for (auto &kv : g_guarded) {
        const GuardEntry &g = kv.second;
        // the zone in front of the buffer, and everything from the end of the requested bytes to the end of the block
        const size_t padded = (g.bytes + 255) / 256 * 256;
        if (cudaMemcpy(h.data(), g.base, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return UPOL_ERR_CUDA - 1;
        bool hit = false;
        for (unsigned char c : h) hit = hit || c != kGuardByte;
        std::vector<unsigned char> t(padded - g.bytes + kGuardBytes);
        if (cudaMemcpy(t.data(), g.base + kGuardBytes + g.bytes, t.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return UPOL_ERR_CUDA - 1;
        for (unsigned char c : t) hit = hit || c != kGuardByte;
        bad += hit ? 1 : 0;
    }
    return bad;
}

// Self-test of the mechanism: one store past the end of a guarded buffer must be counted.  Returns 1 when it is
// (and the count is back to its old value after the buffer is freed), 0 otherwise, -1 with guard mode off.
extern "C" int upol_debug_guard_selftest(void) {
    using namespace upol;
    if (!guard_enabled()) return -1;
    const int before = upol_debug_check_guards();
    char *p = nullptr;
    if (dmalloc((void **)&p, 1000) != cudaSuccess) return 0;
    cudaMemset(p + 1000, 0, 1);                     // first byte past the requested size
    const int during = upol_debug_check_guards();
    dfree(p);
    const int after = upol_debug_check_guards();
    return (during == before + 1 && after == before) ? 1 : 0;
}
