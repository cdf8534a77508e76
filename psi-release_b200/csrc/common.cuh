// Shared device helpers for libpsi_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <nvtx3/nvToolsExt.h>
#include "../../include/psi_b200.h"

#define PSI_RETURN_IF_LAUNCH_FAILED()                  \
    do {                                               \
        cudaError_t e__ = cudaGetLastError();          \
        if (e__ != cudaSuccess) return (int)e__;       \
    } while (0)

// counts the launch (and, under psi_fit_profile, records an event behind it), then reports a
// launch-time error
#define PSI_LAUNCHED_K(name)                           \
    do {                                               \
        psi::count_launch(name);                       \
        PSI_RETURN_IF_LAUNCH_FAILED();                 \
    } while (0)
#define PSI_LAUNCHED() PSI_LAUNCHED_K("kernel")

#define PSI_NUM_SMS 148  // B200: 2 dies x 74 SMs

struct psi_lbs_model;

namespace psi {

// NVTX range around a C-ABI entry point (header-only NVTX3: a no-op unless a profiler is attached)
struct Range {
    explicit Range(const char *name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range &) = delete;
    Range &operator=(const Range &) = delete;
};

void count_launch(const char *name);   // api.cu
// per-launch timing for psi_fit_profile: while a recorder is active on this thread, every counted
// launch is followed by an event on `st`
struct LaunchRecorder {
    cudaStream_t st;
    const char *names[64];
    cudaEvent_t ev[65];
    int n;
};
void recorder_set(LaunchRecorder *r);

// ---- operand layout shared by the tensor-core GEMMs (LBS blend, VPoser decoder) --------------
constexpr int kBG = 64;         // bodies per CTA = M of the tensor-core tiles
constexpr int kKC = 32;         // reduction elements per pipeline stage: one 128-byte row per operand row
// Operand rows are 32 floats = 8 chunks of 16 bytes; chunk c of row r is stored at chunk
// c ^ (r & 7), so the 8 rows of one ldmatrix 8x4 block fall into 8 different bank groups.
__host__ __device__ inline int swz(int row, int col) { return ((((col >> 2) ^ row) & 7) << 2) | (col & 3); }
// element (body b, reduction index k) of an A operand [body group][chunk k/32][64 bodies][32 k]
__host__ __device__ inline size_t a_index(int b, int k, int kpad) {
    return (size_t)(b / kBG) * kpad * kBG + (size_t)(k / kKC) * (kBG * kKC) + (size_t)(b % kBG) * kKC +
           swz(b % kBG, k % kKC);
}

// ---- BF16x3 operand layout (the default blend GEMM path, lbs.cu) ------------------------------------------
// x = b1 + b2 + b3 with three bfloat16 terms (8 + 8 + 8 mantissa bits: every FP32 value exactly, up to 2^-24
// relative).  Operand rows are 64 bf16 = 128 bytes = 8 chunks of 16 bytes, chunk c of row r stored at c ^ (r & 7)
// (the 128-byte swizzle of tcgen05 / TMA).  A per-body operand tile holds the three terms of 64 bodies as 192 rows:
// [body group][chunk k/64][term][64 bodies][64 k].
constexpr int kKC3 = 64;        // reduction elements per 128-byte bf16 row
__host__ __device__ inline int swz16(int row, int col) { return ((((col >> 3) ^ row) & 7) << 3) | (col & 7); }
// element (body b, term t, reduction index k) of a bf16x3 per-body operand with reduction length kpad (multiple of 64)
__host__ __device__ inline size_t a3_index(int b, int t, int k, int kpad) {
    return ((size_t)(b / kBG) * (kpad / kKC3) + (size_t)(k / kKC3)) * (3 * kBG * kKC3) + (size_t)(t * kBG + b % kBG) * kKC3 +
           swz16(b % kBG, k % kKC3);
}
// the three bf16 terms of x (round to nearest even at every step; the residuals are exact in FP32)
__host__ __device__ inline void split_bf16x3(float x, unsigned short &b1, unsigned short &b2, unsigned short &b3) {
    auto rn = [](float v) -> unsigned short {
        unsigned int u;
#ifdef __CUDA_ARCH__
        u = __float_as_uint(v);
#else
        union { float f; unsigned int i; } c; c.f = v; u = c.i;
#endif
        if ((u & 0x7f800000u) == 0x7f800000u) return (unsigned short)(u >> 16);        // inf / nan: truncate
        return (unsigned short)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
    };
    auto up = [](unsigned short h) -> float {
        const unsigned int u = (unsigned int)h << 16;
#ifdef __CUDA_ARCH__
        return __uint_as_float(u);
#else
        union { float f; unsigned int i; } c; c.i = u; return c.f;
#endif
    };
    b1 = rn(x);
    const float r1 = x - up(b1);
    b2 = rn(r1);
    const float r2 = r1 - up(b2);
    b3 = rn(r2);
}

// lbs.cu, used by the fused fitting loop (fit.cu); SdfFuse / VGradFuse: fit_fuse.cuh
struct SdfFuse;
struct VGradFuse;
int lbs_fwd_impl(const psi_lbs_model *m, int B, const float *betas, const float *pose, const float *transl,
                 const float *cam, long cam_bstride, const float *rot_in, const float *rot6d, int num_rot,
                 float *verts, float *joints, float *saved, const SdfFuse *sdf, cudaStream_t st);
int lbs_bwd_impl(const psi_lbs_model *m, int B, const float *pose, const float *cam, long cam_bstride,
                 const float *saved, const float *grad_verts, const VGradFuse *vg, const float *grad_joints,
                 float *grad_betas, float *grad_pose, float *grad_transl, float *grad_rot, int num_rot,
                 const float *rot6d, float *g6_root, float *g6A, int g6_kpad, void *workspace,
                 size_t workspace_bytes, cudaStream_t st);
int lbs_vertex_chunks(const psi_lbs_model *m);   // 256-vertex chunks = rows of SdfFuse::partial / VGradFuse::cpart

// ---- programmatic dependent launch -----------------------------------------------------------
// The fitting iteration is a chain of 15 short dependent kernels; with a plain stream (or graph) dependency
// each one pays the full launch latency after its predecessor has drained.  Launched with the
// programmatic-stream-serialization attribute, a kernel's CTAs are scheduled once every CTA of the
// predecessor has passed pdl_launch_dependents (placed in the kernels' tails) and block in pdl_wait until
// the predecessor has COMPLETED and its writes are visible -- so everything that touches a predecessor's
// output (or overwrites its input) comes after pdl_wait; only barrier initialisation and loads of constants
// precede it.  Opt-in (PSI_PDL=1): it measured slower than plain graph edges, see api.cu.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// opt `func` in to `bytes` of dynamic shared memory on the CURRENT device (once per device; api.cu); 0 or a cudaError_t
int ensure_max_dyn_smem(const void *func, int bytes);
template <typename... KArgs>
inline int ensure_max_dyn_smem(void (*kernel)(KArgs...), size_t bytes) {
    return ensure_max_dyn_smem(reinterpret_cast<const void *>(kernel), (int)bytes);
}

bool skip_kernel(const char *name);   // api.cu: measurement aid, PSI_SKIP_KERNEL=<name> drops that launch (results invalid)
bool pdl_enabled();   // api.cu: true only when PSI_PDL=1 (measured slower, see api.cu)
int pdl_max_ctas();   // api.cu: PSI_PDL_MAX (mode 3's grid-size limit, default 296)
int pdl_mode();       // api.cu: PSI_PDL (0 off, 1 every kernel, 2 only launch_pdl_sel sites, 3 only grids of <= 296 CTAs, 4 = 2 + 3)
// launch `kernel`; it MUST call pdl_wait() before touching global memory other than constants
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // PSI_PDL=3: only the small kernels of the per-body chain (their constant prologues -- TMA of weights, barrier
    // initialisation -- then run under the predecessor's tail)
    cfg.numAttrs = (pdl_enabled() || ((pdl_mode() == 3 || pdl_mode() == 4) && (size_t)grid.x * grid.y * grid.z <= (size_t)pdl_max_ctas())) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// the same for a kernel with a long constant-only prologue: also programmatic under PSI_PDL=2
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_sel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_mode() == 1 || pdl_mode() == 2 || pdl_mode() == 4) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) --------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> this CTA's shared memory, completion counted in bytes on `bar`.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// The same copy with an L2 eviction hint.  The two blend bases (2 x 64 MB) stream through the 126 MB L2 once per
// iteration and cannot stay; loaded evict-first they stop pushing out what is re-used (scene SDF tiles, vertices,
// gradients, the NN index, decoder weights).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ---- tensor-core helpers: legacy mma.sync path (HMMA.1688.F32.TF32), FP32 accumulate ----
// ldmatrix moves 8x8 b16 matrices = 8 rows x 16 bytes; with 32-bit elements a matrix is 8 rows x 4
// words and lane l receives word (l % 4) of row (l / 4): exactly the m16n8k8 TF32 fragments when
// both operands are stored with the reduction index contiguous.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(saddr) : "memory");
}
// x = hi + lo exactly: hi keeps the TF32 bits of x (sign, exponent, 10 mantissa bits), lo the 13
// bits below (the tensor core reads the top 11 of them).  One LOP3 + one FADD per element;
// cvt.rna.tf32 would cost five instructions (it is emulated, with NaN/Inf handling).
// 3xTF32: a*b ~= lo_a*hi_b + hi_a*lo_b + hi_a*hi_b, relative error < 2^-19, FP32 accumulate.
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t &hi, uint32_t &lo) {
    hi = x & 0xffffe000u;
    lo = __float_as_uint(__uint_as_float(x) - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
    mma_tf32(d, al, bh0, bh1);      // small terms first
    mma_tf32(d, ah, bl0, bl1);
    mma_tf32(d, ah, bh0, bh1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace psi
