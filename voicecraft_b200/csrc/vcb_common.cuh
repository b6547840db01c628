// Common device helpers for the sm_100a kernels: mbarrier / TMA / tcgen05 PTX wrappers, small math utils.
// Everything here is hand-written PTX for Blackwell (no CUTLASS dependency).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vcb {

// ------------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------------
#define VCB_CUDA_OK(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            vcb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)

void set_error(const char* fmt, ...);
const char* get_error();

// ------------------------------------------------------------------------------------------------
// bf16 split:  x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi).  Residual error <= 2^-17 |x|.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// shared-memory address, mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ------------------------------------------------------------------------------------------------
// TMA: tiled tensor load (2D) and plain bulk copy, both completing on an mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// c0 = innermost coordinate (elements), c1 = row
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// contiguous global -> shared, bytes % 16 == 0, both 16B aligned
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// L2 cache policies / prefetch.  Streams that are touched once per step (KV pages, weight tiles) are loaded with
// evict_first so they do not push the *prefetched* next-kernel weights out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// asynchronous HBM -> L2 prefetch of a contiguous range (bytes % 16 == 0)
__device__ __forceinline__ void tma_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes) : "memory");
}
// this CTA's share of a [ptr, ptr+bytes) range, issued by one thread in <= 32 KB pieces
__device__ __forceinline__ void prefetch_l2_slice(const void* ptr, unsigned long long bytes, unsigned int part, unsigned int nparts) {
    if (ptr == nullptr || bytes == 0) return;
    unsigned long long chunk = ((bytes + nparts - 1) / nparts + 127ull) & ~127ull;
    unsigned long long beg = chunk * part;
    if (beg >= bytes) return;
    unsigned long long end = beg + chunk < bytes ? beg + chunk : bytes;
    const char* p = static_cast<const char*>(ptr);
    for (unsigned long long o = beg; o < end; o += 32768ull) {
        const unsigned long long n = (end - o < 32768ull ? end - o : 32768ull) & ~15ull;
        if (n) tma_prefetch_l2(p + o, static_cast<uint32_t>(n));
    }
}

// Optional device-side timeline (debug): CTA 0 of instrumented kernels appends (tag, globaltimer ns) records.
// Each translation unit has its own copy of the pointer; the host sets them through vcb_timeline_set().
static __constant__ unsigned long long* g_tl_buf = nullptr;   // constant bank: the disabled check costs no L2 round trip
static __constant__ unsigned int* g_tl_cnt = nullptr;
__device__ __forceinline__ void tl_mark(unsigned int tag) {
#ifndef VCB_TIMELINE
    (void)tag;                                   // compiled out by default: even the disabled check costs 2.7 % of a decode step
    return;                                      // (profiles/r02_exp_timeline_marks.json); `make TIMELINE=1` builds it in
#endif
    if (g_tl_buf != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned int i = atomicAdd(g_tl_cnt, 1u);
        if (i < 65536u) {
            g_tl_buf[2 * i] = tag;
            g_tl_buf[2 * i + 1] = t;
        }
    }
}

// Every CTA records (vcb_timeline(2, ...)): tag | 0x8000 | cta << 16 -- used to see launch skew and stragglers.
__device__ __forceinline__ void tl_mark_all(unsigned int tag) {
#ifndef VCB_TIMELINE
    (void)tag;
    return;
#endif
    if (g_tl_buf != nullptr && g_tl_cnt[1] != 0u) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned int i = atomicAdd(g_tl_cnt, 1u);
        if (i < 65536u) {
            g_tl_buf[2 * i] = tag | 0x8000u | ((blockIdx.x + gridDim.x * blockIdx.y) << 16);
            g_tl_buf[2 * i + 1] = t;
        }
    }
}

// Programmatic dependent launch (PDL): wait for the producer grid / let the consumer grid start.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, TMEM load
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc];  bf16 inputs, fp32 accumulate, single-CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (base_lane + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major operand tile in shared memory, rows of 128 bytes (64 bf16), 128B swizzle (what TMA SWIZZLE_128B writes):
// 8-row x 128B swizzle atoms, atoms stacked every 1024 B.  Descriptor layout = cute::UMMA::SmemDescriptor (sm100):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 | [46,48) version=1 |
//   [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;                       // LBO (ignored), canonical value 1
    d |= static_cast<uint64_t>(1024 >> 4) << 32;               // SBO: 8 rows * 128 B
    d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: bf16 x bf16 -> fp32, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(int M, int N) {
    return (1u << 4)                                  // c_format = F32
           | (1u << 7)                                // a_format = BF16
           | (1u << 10)                               // b_format = BF16
           | (static_cast<uint32_t>(N >> 3) << 17)    // n_dim
           | (static_cast<uint32_t>(M >> 4) << 24);   // m_dim
}

}  // namespace vcb
