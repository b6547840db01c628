// Row-major ("rows as M") GEMM of the codec-LM for many token rows at once: prompt prefill.
//
//   out[r][n] = epilogue( sum_k (Xhi[r,k] + Xlo[r,k]) * W[n,k] + bias[n] )        r = token row, n = output feature
//
// gemm_tcgen05.cu keeps the WEIGHTS as the 128-lane operand and at most 128 token rows as the UMMA N dimension: right
// for decode (every weight byte is used once per step), wrong for a prompt of thousands of rows, where it streams the
// whole matrix once per 128 rows and spends most of each launch in the per-row epilogue.  Here a CTA owns a
// [128 rows] x [BN features] output tile:
//   A operand = activations, bf16 [2][rcap][K]: plane 0 = hi parts, plane 1 = lo parts (x ~= hi + lo, split_bf16);
//               both 128 x 64 tiles of a k-block are multiplied with the SAME weight tile into the SAME accumulator
//               (D += Ahi.B^T ; D += Alo.B^T), so the second pass costs no extra weight traffic.
//   B operand = the pre-tiled weights of gemm_tcgen05.cu, unchanged: BN/128 consecutive 16 KB blocks of one k-block,
//               stacked in shared memory, are exactly a K-major SWIZZLE_128B operand of BN rows.
//   D         = fp32 in TMEM, lane = token row, column = feature (BN <= 256 columns).
// No split-K, no cluster: with >= 512 rows there are enough tiles (e.g. QKV at d = 2048: 24 x rows/128 CTAs), and the
// weights (<= 34 MB per matrix) stay L2-resident across the row tiles.
// Warp roles as in gemm_tcgen05.cu: w0 TMA producer, w1 TMEM alloc + MMA issuer, w2..w9 epilogue (two sets of four,
// each set covers the 128 TMEM lanes and takes half of the columns).
// Epilogues: EPI_QKV (q -> fp32 rows, k/v -> paged KV cache), EPI_RESID (x += y + b), EPI_ACT (ReLU/GELU -> hi/lo
// planes), EPI_LOGITS (plain fp32 rows; bring-up tests).  A thread owns one token row and 32 consecutive features
// per step, so every access is a run of 16-byte vectors.
#include "vcb_internal.h"

#include <algorithm>
#include <cstdio>

namespace vcb {

static constexpr int RG_BM = 128;     // token rows per CTA (UMMA M)
static constexpr int RG_BK = 64;      // K elements per stage (one 128-byte swizzle row of bf16)
static constexpr int RG_EPI_WARPS = 8;
static constexpr int RG_THREADS = 64 + 32 * RG_EPI_WARPS;

template <int BN, int STAGES>
struct RowsSmem {
    static constexpr int A_BYTES = RG_BM * RG_BK * 2;               // one of the hi / lo tiles
    static constexpr int B_BYTES = BN * RG_BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16;
};

__device__ __forceinline__ float act_fn(float v, int kind) {
    if (kind == 1) return fmaxf(v, 0.f);
    if (kind == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    return v;
}

// 32 consecutive features [f0, f0+32) of token row `row`, accumulator values in v[]
__device__ __forceinline__ void rows_epilogue32(const GemmEpilogue& ep, int row, int f0, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += ep.bias[f0 + j];         // same address in every lane: broadcast
    switch (ep.mode) {
        case EPI_QKV: {
            const int part = f0 / ep.d, cc = f0 - part * ep.d;
            if (part == 0) {
                float4* dst = reinterpret_cast<float4*>(ep.qbuf + static_cast<size_t>(row) * ep.d + cc);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                return;
            }
            const int pos = ep.row_pos[row];
            if (pos < 0) return;
            const int page = ep.row_page ? ep.row_page[row] : ep.page_table[ep.row_slot[row] * ep.max_pages + pos / ep.page_size];
            const int h = cc / ep.hd, e0 = cc - h * ep.hd;          // hd % 32 == 0: the 32 features share a head
            const size_t off = ((static_cast<size_t>(page) * ep.H + h) * ep.page_size + pos % ep.page_size) * ep.hd + e0;
            void* pool = (part == 1) ? ep.kpool : ep.vpool;
            if (ep.kv_fp32) {
                float4* dst = reinterpret_cast<float4*>(static_cast<float*>(pool) + off);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
                uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(pool) + off);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
                    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
                    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
                    __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
                    dst[j] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                        *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
                }
            }
            break;
        }
        case EPI_RESID: {
            float4* px = reinterpret_cast<float4*>(ep.x + static_cast<size_t>(row) * ep.ld_out + f0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 t = px[j];
                t.x += v[4 * j];
                t.y += v[4 * j + 1];
                t.z += v[4 * j + 2];
                t.w += v[4 * j + 3];
                px[j] = t;
            }
            break;
        }
        case EPI_ACT: {
            uint4* ph = reinterpret_cast<uint4*>(ep.act + static_cast<size_t>(row) * ep.ld_out + f0);
            uint4* pl = reinterpret_cast<uint4*>(ep.act + static_cast<size_t>(row + ep.bpad_out) * ep.ld_out + f0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(act_fn(v[8 * j + 2 * u], ep.act_kind), h0, l0);
                    split_bf16(act_fn(v[8 * j + 2 * u + 1], ep.act_kind), h1, l1);
                    __nv_bfloat162 hh = __halves2bfloat162(h0, h1), ll = __halves2bfloat162(l0, l1);
                    hw[u] = *reinterpret_cast<uint32_t*>(&hh);
                    lw[u] = *reinterpret_cast<uint32_t*>(&ll);
                }
                ph[j] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                pl[j] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            break;
        }
        default: {
            float4* dst = reinterpret_cast<float4*>(ep.out + static_cast<size_t>(row) * ep.ld_out + ep.col_off + f0);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(RG_THREADS)
gemm_rows_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const GemmEpilogue ep,
                 int rows, int rcap, int Nout, int total_kb) {
    using L = RowsSmem<BN, STAGES>;
    extern __shared__ __align__(1024) uint8_t smem[];       // SWIZZLE_128B tiles need 1024-byte alignment
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN;                         // first output feature of this tile
    const int r0 = blockIdx.y * RG_BM;                      // first token row
    const int mt0 = n0 / 128;                               // first 128-feature weight block
    const int pre = min(total_kb, STAGES);

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
        // weights never depend on the previous kernel: the first stages' weight tiles go in flight before the wait
        for (int i = 0; i < pre; ++i) {
            mbar_arrive_expect_tx(&full_bar[i], L::STAGE_BYTES);
            uint8_t* b = smem + i * L::STAGE_BYTES + 2 * L::A_BYTES;
#pragma unroll
            for (int j = 0; j < BN / 128; ++j)
                tma_load_2d(b + j * (128 * RG_BK * 2), &tmW, &full_bar[i], 0, ((mt0 + j) * total_kb + i) * 128);
        }
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer ==========================================================================
        if (lane == 0) {
            pdl_wait();                                     // the activation planes come from the previous kernel
            for (int i = 0; i < pre; ++i) {
                uint8_t* a = smem + i * L::STAGE_BYTES;
                tma_load_2d(a, &tmX, &full_bar[i], i * RG_BK, r0);
                tma_load_2d(a + L::A_BYTES, &tmX, &full_bar[i], i * RG_BK, rcap + r0);
            }
            int stage = 0, phase = 0;
            for (int i = pre; i < total_kb; ++i) {
                mbar_wait(&empty_bar[stage], phase);        // the MMAs released this slot
                mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                uint8_t* a = smem + stage * L::STAGE_BYTES;
                tma_load_2d(a, &tmX, &full_bar[stage], i * RG_BK, r0);
                tma_load_2d(a + L::A_BYTES, &tmX, &full_bar[stage], i * RG_BK, rcap + r0);
#pragma unroll
                for (int j = 0; j < BN / 128; ++j)
                    tma_load_2d(a + 2 * L::A_BYTES + j * (128 * RG_BK * 2), &tmW, &full_bar[stage], 0,
                                ((mt0 + j) * total_kb + i) * 128);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer ============================================================================
        constexpr uint32_t idesc = umma_idesc_bf16_f32(RG_BM, BN);
        int stage = 0, phase = 0;
        for (int i = 0; i < total_kb; ++i) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
                const uint64_t hi_desc = umma_desc_kmajor_sw128(a_addr);
                const uint64_t lo_desc = umma_desc_kmajor_sw128(a_addr + L::A_BYTES);
                const uint64_t b_desc = umma_desc_kmajor_sw128(a_addr + 2 * L::A_BYTES);
#pragma unroll
                for (int k = 0; k < RG_BK / 16; ++k) {      // +32 B per 16 K-elements inside the swizzle row
                    umma_bf16(tmem_base, hi_desc + 2 * k, b_desc + 2 * k, idesc, (i | k) != 0);
                    umma_bf16(tmem_base, lo_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
                }
                umma_commit(&empty_bar[stage]);             // frees the smem slot when the MMAs retire
                if (i == total_kb - 1) umma_commit(tmem_full);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> fused epilogue, one token row per thread ================
        const int q = warp & 3;                             // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;                   // which half of the columns
        const int row = r0 + q * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        pdl_wait();                                         // residual rows / KV positions come from earlier kernels
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
            float v[32];
            tmem_ld_32x32(lane_addr + c, v);                // warp-collective: outside the row / feature guards
            if (row < rows && n0 + c < Nout) rows_epilogue32(ep, row, n0 + c, v);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

template <int BN, int STAGES>
static int launch_rows(const RowsGemmCall& g, cudaStream_t st) {
    using L = RowsSmem<BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        VCB_CUDA_OK(cudaFuncSetAttribute(gemm_rows_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((g.Nout + BN - 1) / BN, (g.rows + RG_BM - 1) / RG_BM, 1);
    cfg.blockDim = dim3(RG_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g.pdl ? 1 : 0;
    VCB_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_rows_kernel<BN, STAGES>, *g.tmX, *g.tmW, g.ep, g.rows, g.rcap, g.Nout,
                                   g.Kdim / RG_BK));
    return 0;
}

// Shapes this kernel takes: K a multiple of 64, Nout a multiple of 128 (whole weight blocks), hd a multiple of 32 for
// the QKV epilogue, rcap (rows per plane) a multiple of 128.
bool gemm_rows_supported(int Nout, int Kdim, int hd) {
    return Kdim % RG_BK == 0 && Nout % 128 == 0 && (hd == 0 || hd % 32 == 0);
}

int gemm_rows_launch(const RowsGemmCall& g, cudaStream_t st) {
    if (!gemm_rows_supported(g.Nout, g.Kdim, g.ep.mode == EPI_QKV ? g.ep.hd : 0) || g.rcap % RG_BM || g.rows < 1 ||
        g.rows > g.rcap) {
        set_error("gemm_rows: unsupported shape N=%d K=%d rows=%d rcap=%d", g.Nout, g.Kdim, g.rows, g.rcap);
        return -1;
    }
    // 256-feature tiles halve the activation re-reads; an odd number of 128-feature blocks runs 128-wide tiles
    if ((g.Nout / 128) % 2 == 0) return launch_rows<256, 3>(g, st);
    return launch_rows<128, 4>(g, st);
}

}  // namespace vcb
