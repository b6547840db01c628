// Weight-streaming GEMM for the codec-LM: partial[z][j][m] = sum_{k in split z} W[m,k] * (Xhi[j,k] + Xlo[j,k])
//
//   A operand = weight matrix W [Nout, Kdim] bf16 row-major (K-major), tile 128 x 64, TMA SWIZZLE_128B
//   B operand = activations  X [2*Bpad, Kdim] bf16: rows [0,Bpad) = hi parts, rows [Bpad,2*Bpad) = lo parts
//               (x ~= hi + lo, see split_bf16) -> one UMMA of N = 2*Bpad columns covers both, the epilogue
//               adds column j and column j+Bpad.  The tensor pipe is idle >80% of the time in this
//               HBM-bound regime, so the second half is free and buys ~16 mantissa bits on activations.
//   D         = fp32 accumulator in TMEM, 128 lanes (= output features) x 2*Bpad columns
//
// One CTA = one 128-feature tile x one K split.  Warp roles: w0 TMA producer, w1 TMEM alloc + MMA issuer
// (single elected thread issues tcgen05.mma), w2..w5 epilogue (tcgen05.ld -> coalesced fp32 partial stores).
// Split-K partials are reduced deterministically (fixed z order) by the consumer kernels in lm_kernels.cu.
//
// Replaces in the reference: F.linear at models/modules/activation.py:86 (packed QKV), :637 (out_proj),
// models/modules/transformer.py:387 (FFN linear1/linear2) and models/voicecraft.py:181-185,1085 (logit heads).
#include "vcb_internal.h"

#include <algorithm>

namespace vcb {

static constexpr int GEMM_BM = 128;   // output features per CTA (UMMA M)
static constexpr int GEMM_BK = 64;    // K elements per pipeline stage (= 128 B of bf16 = one swizzle row)
static constexpr int GEMM_THREADS = 192;

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_w_xT_splitk(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 float* __restrict__ partial, int Nout, int ldp, int total_kb, int kb_per_split,
                 int b_col_off, int nvalid) {
    using L = GemmSmem<BN, STAGES>;
    constexpr int BPAD = BN / 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * GEMM_BM;
    const int z = blockIdx.z;
    const int kb0 = z * kb_per_split;
    const int nkb = min(kb_per_split, total_kb - kb0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer ==========================================================================
        if (lane == 0) {
            // Weights never depend on the previous kernel: under PDL their first STAGES tiles are in
            // flight before the producer grid has drained; activations wait for griddepcontrol.wait.
            const int pre = min(nkb, STAGES);
            for (int i = 0; i < pre; ++i) {
                mbar_arrive_expect_tx(&full_bar[i], L::STAGE_BYTES);
                tma_load_2d(smem + i * L::STAGE_BYTES, &tmA, &full_bar[i], (kb0 + i) * GEMM_BK, m0);
            }
            pdl_wait();
            for (int i = 0; i < pre; ++i)
                tma_load_2d(smem + i * L::STAGE_BYTES + L::A_BYTES, &tmB, &full_bar[i],
                            b_col_off + (kb0 + i) * GEMM_BK, 0);
            int stage = 0, phase = 0;                       // state after the first `pre` fills
            for (int i = pre; i < nkb; ++i) {
                // stage `stage` was filled in the previous round; wait until the MMA released it
                mbar_wait(&empty_bar[stage], phase);
                mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                uint8_t* a = smem + stage * L::STAGE_BYTES;
                tma_load_2d(a, &tmA, &full_bar[stage], (kb0 + i) * GEMM_BK, m0);
                tma_load_2d(a + L::A_BYTES, &tmB, &full_bar[stage], b_col_off + (kb0 + i) * GEMM_BK, 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer ============================================================================
        constexpr uint32_t idesc = umma_idesc_bf16_f32(GEMM_BM, BN);
        int stage = 0, phase = 0;
        for (int i = 0; i < nkb; ++i) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
                const uint64_t a_desc = umma_desc_kmajor_sw128(a_addr);
                const uint64_t b_desc = umma_desc_kmajor_sw128(a_addr + L::A_BYTES);
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                    // advance 16 K-elements = 32 bytes inside the 128B swizzle row: +2 in the (>>4) address field
                    umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (i | k) != 0);
                }
                umma_commit(&empty_bar[stage]);                 // frees the smem slot when the MMAs retire
                if (i == nkb - 1) umma_commit(tmem_full);       // accumulator complete
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> fp32 partials =========================================
        pdl_launch_dependents();
        const int q = warp & 3;                                 // TMEM lane quarter owned by this warp
        const int m = m0 + q * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float* out = partial + (static_cast<size_t>(z) * BPAD) * ldp + m;
        constexpr int CH = BPAD < 32 ? 16 : 32;
#pragma unroll 1
        for (int c = 0; c < BPAD; c += CH) {
            float hi[CH], lo[CH];
            if constexpr (CH == 32) {
                tmem_ld_32x32(lane_addr + c, hi);
                tmem_ld_32x32(lane_addr + BPAD + c, lo);
            } else {
                tmem_ld_32x16(lane_addr + c, hi);
                tmem_ld_32x16(lane_addr + BPAD + c, lo);
            }
            if (m < Nout) {
#pragma unroll
                for (int j = 0; j < CH; ++j)
                    if (c + j < nvalid) out[static_cast<size_t>(c + j) * ldp] = hi[j] + lo[j];
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
    }
}

// ---------------------------------------------------------------------------------------------------
// Bring-up / cross-check kernel: same contract on CUDA cores (one warp per output feature).  Selected
// with VCB_GEMM_IMPL=simt; never the default.  It exists so a tcgen05 descriptor bug can be told apart
// from a bug anywhere else in the step.
// ---------------------------------------------------------------------------------------------------
__global__ void gemm_w_xT_simt(const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ X,
                               float* __restrict__ partial, int Nout, int Kdim, int ldx, int ldp, int bpad,
                               int total_kb, int kb_per_split, int b_col_off, int nvalid) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int z = blockIdx.z;
    if (warp >= Nout) return;
    const int k0 = z * kb_per_split * GEMM_BK;
    const int k1 = min(Kdim, (z * kb_per_split + min(kb_per_split, total_kb - z * kb_per_split)) * GEMM_BK);
    for (int j = 0; j < nvalid; ++j) {
        float acc = 0.f;
        for (int k = k0 + lane; k < k1; k += 32) {
            const float w = __bfloat162float(W[static_cast<size_t>(warp) * Kdim + k]);
            const float xh = __bfloat162float(X[static_cast<size_t>(j) * ldx + b_col_off + k]);
            const float xl = __bfloat162float(X[static_cast<size_t>(j + bpad) * ldx + b_col_off + k]);
            acc = fmaf(w, xh, acc);
            acc = fmaf(w, xl, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) partial[(static_cast<size_t>(z) * bpad + j) * ldp + warp] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128B swizzle, OOB -> 0
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return -1;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * 2};
    cuuint32_t box[2] = {GEMM_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: %d (rows=%llu cols=%llu ld=%llu box_rows=%u)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows);
        return -1;
    }
    return 0;
}

template <int BN, int STAGES>
static int launch_one(const GemmCall& g, cudaStream_t st) {
    using L = GemmSmem<BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        VCB_CUDA_OK(cudaFuncSetAttribute(gemm_w_xT_splitk<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((g.Nout + GEMM_BM - 1) / GEMM_BM, 1, g.splits);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g.pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int total_kb = g.Kdim / GEMM_BK;
    const int kbps = (total_kb + g.splits - 1) / g.splits;
    VCB_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_w_xT_splitk<BN, STAGES>, *g.tmA, *g.tmB, g.partial, g.Nout, g.ldp,
                                   total_kb, kbps, g.b_col_off, g.nvalid));
    return 0;
}

int gemm_launch(const GemmCall& g, cudaStream_t st) {
    if (g.Kdim % GEMM_BK != 0) {
        set_error("gemm: K=%d not a multiple of %d", g.Kdim, GEMM_BK);
        return -1;
    }
    const int total_kb = g.Kdim / GEMM_BK;
    const int kbps = (total_kb + g.splits - 1) / g.splits;
    if ((g.splits - 1) * kbps >= total_kb) {
        set_error("gemm: splits=%d leaves an empty split for %d k-blocks", g.splits, total_kb);
        return -1;
    }
    if (g.simt) {
        dim3 grid((g.Nout * 32 + 255) / 256, 1, g.splits);
        gemm_w_xT_simt<<<grid, 256, 0, st>>>(g.W, g.X, g.partial, g.Nout, g.Kdim, g.ldx, g.ldp, g.bpad, total_kb,
                                             kbps, g.b_col_off, g.nvalid);
        VCB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    switch (g.bpad) {
        case 16: return launch_one<32, 4>(g, st);
        case 32: return launch_one<64, 4>(g, st);     // 4 x 24 KB: two CTAs per SM stay resident (PDL overlap)
        case 64: return launch_one<128, 3>(g, st);
        case 128: return launch_one<256, 4>(g, st);
        default: set_error("gemm: unsupported bpad %d", g.bpad); return -1;
    }
}

int gemm_pick_splits(int Nout, int Kdim, int num_sms) {
    // ~1.5 CTAs per SM so every SM streams weights, but at least 4 k-blocks (32 KB of weights) per CTA
    const int tiles = (Nout + GEMM_BM - 1) / GEMM_BM;
    const int total_kb = Kdim / GEMM_BK;
    int s = (3 * num_sms / 2 + tiles - 1) / tiles;
    s = std::min(s, std::max(1, total_kb / 4));
    s = std::max(1, std::min(s, 16));
    while (s > 1 && (s - 1) * ((total_kb + s - 1) / s) >= total_kb) --s;
    return s;
}

}  // namespace vcb
