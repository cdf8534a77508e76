// ABI version and error strings for libpsi_b200.
#include "common.cuh"
#include <atomic>
#include <map>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <utility>

namespace psi {
static std::atomic<unsigned long long> g_launches{0};
// Ablation timing: with PSI_SKIP_KERNEL=<name> every launch tagged <name> is skipped, so the captured
// iteration shows that kernel's marginal cost inside the graph (eager per-launch timings overstate the
// short kernels).  The results of such a run are meaningless; bench.py never sets it.
bool skip_kernel(const char *name) {
    static const char *skip = getenv("PSI_SKIP_KERNEL");
    return skip && name && strcmp(skip, name) == 0;
}
int pdl_mode() {
    // PSI_PDL: programmatic dependent launch (griddepcontrol) between consecutive kernels of a chain.
    //   3 (default) = only kernels whose grid has at most PSI_PDL_MAX (296) CTAs -- the per-body chain (decoder GEMMs,
    //       pose kernels, fit_step) and the persistent blend GEMMs: their constant-only prologues (TMA of weights and
    //       of Jdirs, barrier / TMEM set-up) run under the predecessor's tail.  797 vs 767 bodies/s.
    //   0 = plain graph edges;  1 = every kernel;  2 = only the NN walk and the vertex backward kernel;
    //   4 = 3 plus the NN walk and the vertex backward kernel (794.6 vs 796.9: nothing gained).
    // Programmatic launch of the BIG grids is what made mode 1 slower every time it was measured (460 vs 535, 593 vs
    // 630, 640 vs 663, 732 vs 767 bodies/s): with the skinning kernel (1344 CTAs) included the gain turns into a
    // loss (734), without it (limits 100 ... 400) the loop runs at 786 ... 797.
    static const int mode = [] { const char *e = getenv("PSI_PDL"); return e ? atoi(e) : 3; }();
    return mode;
}
bool pdl_enabled() { return pdl_mode() == 1; }
int pdl_max_ctas() {
    static const int v = [] { const char *e = getenv("PSI_PDL_MAX"); return e ? atoi(e) : 296; }();
    return v;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: a process that drives several GPUs
// through the C ABI has to opt every kernel in on every device it launches on.  (device, function) pairs that
// already hold at least `bytes` are remembered; the map is the only mutable process-wide state besides the
// launch counter and is guarded by a mutex.
int ensure_max_dyn_smem(const void *func, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, int> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find({dev, func});
    if (it != done.end() && it->second >= bytes) return PSI_OK;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    done[{dev, func}] = bytes;
    return PSI_OK;
}
static thread_local LaunchRecorder *g_rec = nullptr;
void recorder_set(LaunchRecorder *r) { g_rec = r; }
void count_launch(const char *name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    LaunchRecorder *r = g_rec;
    if (r && r->n < 64) {
        r->names[r->n] = name;
        cudaEventRecord(r->ev[r->n + 1], r->st);
        ++r->n;
    }
}
}  // namespace psi

extern "C" {

int psi_abi_version(void) { return PSI_ABI_VERSION; }

unsigned long long psi_launch_count(void) { return psi::g_launches.load(std::memory_order_relaxed); }

const char *psi_error_string(int code) {
    switch (code) {
        case PSI_OK: return "ok";
        case PSI_ERR_BAD_ARG: return "psi: bad argument (null pointer, negative size or misaligned buffer)";
        case PSI_ERR_WORKSPACE: return "psi: workspace too small";
        case PSI_ERR_UNSUPPORTED: return "psi: shape not supported by this build";
        case PSI_ERR_ALLOC: return "psi: device allocation failed";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "psi: unknown error";
}

}  // extern "C"
