// Weight-streaming GEMM of the codec-LM with cluster split-K and fused epilogues.
//
//   out[j][m] = epilogue( sum_k W[m,k] * (Xhi[j,k] + Xlo[j,k]) + bias[m] )        j = token row, m = output feature
//
//   A operand = weight matrix W [Nout, Kdim] bf16, PRE-TILED in HBM: tile (mt, kb) = W[mt*128.., kb*64..] is one
//               contiguous 16 KB block [128 rows][64 cols] at block index mt*KB + kb (pack_weight_tiles), so every
//               TMA box is a single sequential DRAM burst (a row-major matrix would scatter a box over 128 DRAM
//               pages -- measured 18% DRAM utilisation in ncu).  TMA SWIZZLE_128B into smem.
//   B operand = activations  X [2*Bpad, Kdim] bf16: rows [0,Bpad) = hi parts, rows [Bpad,2*Bpad) = lo parts
//               (x ~= hi + lo, see split_bf16): one UMMA of N = 2*Bpad columns covers both; the tensor pipe is
//               >80% idle in this HBM-bound regime, so the second half is free and buys ~16 mantissa bits.
//   D         = fp32 accumulator in TMEM, 128 lanes (= output features) x 2*Bpad columns
//
// Split-K without a workspace: the S CTAs of a thread-block cluster each stream one K slice of the same 128-feature
// weight tile (so >=128 CTAs pull HBM even for a 2048 x 2048 matrix), then reduce-scatter their accumulators
// through distributed shared memory: CTA z owns token rows [z*R, (z+1)*R), every CTA writes its partial of those
// rows into the owner's smem with asynchronous stores that count their bytes on the owner's mbarrier (st.async ...
// mbarrier::complete_tx): the owner waits on its own barrier -- no cluster barrier after the start-up one -- then sums
// the S partials in fixed order (deterministic) and applies the fused epilogue:
//     EPI_QKV    q -> fp32 buffer, k/v -> appended to the paged KV cache    (activation.py:86-88, 626-631)
//     EPI_RESID  x += y + bias                                            (transformer.py:321-329)
//     EPI_ACT    ReLU / exact GELU -> bf16 hi/lo rows of the next GEMM      (transformer.py:387, voicecraft.py:183)
//     EPI_LOGITS fp32 logits                                              (voicecraft.py:1085)
// Warp roles: w0 TMA producer, w1 TMEM alloc + MMA issuer (one elected thread issues tcgen05.mma), w2..w5 epilogue
// (plus w6..w9 for tiles with >= 64 token rows).  Prompt batches use the rows-as-M kernel in gemm_rows.cu instead.
#include "vcb_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace vcb {

static constexpr int GEMM_BM = 128;   // output features per CTA (UMMA M)
static constexpr int GEMM_BK = 64;    // K elements per pipeline stage (= 128 B of bf16 = one swizzle row)
// threads per CTA: w0 TMA, w1 MMA, then 4 epilogue warps (one per TMEM lane quarter); tiles with >= 64 token rows get a
// second set of 4 that takes the other half of the rows (the epilogue, not the weight stream, dominates those launches)
constexpr int gemm_epi_warps(int bn) { return bn >= 128 ? 8 : 4; }
constexpr int gemm_threads(int bn) { return 64 + 32 * gemm_epi_warps(bn); }

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int RED_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int RED_BYTES = (BN / 2) * GEMM_BM * 4;          // [S][R][128] fp32 with S*R = Bpad
    static constexpr int BAR_OFFSET = RED_OFFSET + RED_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 2) * 8 + 8;       // full[S] empty[S] tmem_full red_full + tmem slot
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(cta));
    return remote;
}
// Asynchronous store into a peer CTA's shared memory that counts its bytes on the peer's mbarrier: the owner of the
// rows learns that every partial has landed from its own barrier -- no cluster-wide barrier (and no GPU-scope
// membar, which is what barrier.cluster.arrive.release costs) between the accumulators and the epilogue.
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, uint32_t remote_bar, float v) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(remote_addr), "f"(v),
                 "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void st_async_f32x4(uint32_t remote_addr, uint32_t remote_bar, float a, float b, float c, float d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     remote_addr),
                 "f"(a), "f"(b), "f"(c), "f"(d), "r"(remote_bar)
                 : "memory");
}
// bounded wait: a byte-count mismatch would otherwise hang the GPU; ~1 s of polling, then trap (surfaces as a CUDA error)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    for (unsigned int spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}

// Fused epilogue for up to 4 token rows of one output feature m.  All loads of the group are issued before the first
// dependent use (the per-row chains position -> page -> address would otherwise serialise on L2 latency).
__device__ __forceinline__ void apply_epilogue4(const GemmEpilogue& ep, int row0, int nrows, int m, const float (&sum)[4],
                                                float bias, float (&xnew)[4], int colx = 0, const float* xpre = nullptr) {
    switch (ep.mode) {
        case EPI_QKV: {
            int pos[4], slot[4], page[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                pos[u] = (u < nrows) ? ep.row_pos[row0 + u] : -1;
                slot[u] = (u < nrows && !ep.row_page) ? ep.row_slot[row0 + u] : 0;
                page[u] = (u < nrows && ep.row_page) ? ep.row_page[row0 + u] : 0;      // same batch of loads as pos
            }
            const int part = m / ep.d, cc = m - part * ep.d;
            if (part == 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (pos[u] >= 0) ep.qbuf[static_cast<size_t>(row0 + u) * ep.d + cc] = sum[u] + bias;
                return;
            }
            if (!ep.row_page) {         // (precomputed by step_prep / prefill otherwise: one L2 round trip instead of two)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    page[u] = (pos[u] >= 0) ? ep.page_table[slot[u] * ep.max_pages + pos[u] / ep.page_size] : 0;
            }
            const int h = cc / ep.hd, e = cc - h * ep.hd;
            void* pool = (part == 1) ? ep.kpool : ep.vpool;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (pos[u] < 0) continue;
                const size_t off = ((static_cast<size_t>(page[u]) * ep.H + h) * ep.page_size + pos[u] % ep.page_size) * ep.hd + e;
                const float val = sum[u] + bias;
                if (ep.kv_fp32) static_cast<float*>(pool)[off] = val;
                else static_cast<__nv_bfloat16*>(pool)[off] = __float2bfloat16_rn(val);
            }
            break;
        }
        case EPI_RESID: {
            float xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                xv[u] = xpre ? xpre[u] : ((u < nrows) ? ep.x[static_cast<size_t>(row0 + u) * ep.ld_out + m] : 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                xnew[u] = xv[u] + (sum[u] + bias);
                if (u < nrows) ep.x[static_cast<size_t>(row0 + u) * ep.ld_out + m] = xnew[u];
            }
            break;
        }
        case EPI_ACT: {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u >= nrows) continue;
                float v = sum[u] + bias;
                if (ep.act_kind == 1) v = fmaxf(v, 0.f);
                else if (ep.act_kind == 2) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
                __nv_bfloat16 hi, lo;
                split_bf16(v, hi, lo);
                ep.act[static_cast<size_t>(row0 + u) * ep.ld_out + m] = hi;
                ep.act[static_cast<size_t>(row0 + u + ep.bpad_out) * ep.ld_out + m] = lo;
            }
            break;
        }
        default:
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (u < nrows) ep.out[static_cast<size_t>(row0 + u) * ep.ld_out + ep.col_off + colx + m] = sum[u] + bias;
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(gemm_threads(BN))
gemm_w_xT_cluster(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmEpilogue ep, int Nout, int total_kb, int kb_per_split, int b_col_off, int nvalid,
                  const void* pf_ptr, unsigned long long pf_bytes, const GemmGroup grp) {
    using L = GemmSmem<BN, STAGES>;
    constexpr int BPAD = BN / 2;
    constexpr int HALVES = gemm_epi_warps(BN) / 4;          // epilogue warp sets, each covering all 128 TMEM lanes
    constexpr int EPI_THREADS = 32 * gemm_epi_warps(BN);
    extern __shared__ __align__(1024) uint8_t smem[];       // SWIZZLE_128B tiles need 1024-byte alignment
    float* red = reinterpret_cast<float*>(smem + L::RED_OFFSET);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* red_full = tmem_full + 1;                     // all S partials of my R rows have landed in `red`
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(red_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = static_cast<int>(cluster_nctarank());      // K splits = cluster size
    const int z = static_cast<int>(cluster_ctarank());
    const int mt = blockIdx.x / S;                          // 128-feature tile
    const int m0 = mt * GEMM_BM;
    // grouped launch (blockIdx.y = group): same shapes, per-group weight map / bias / column offsets (the K logit heads)
    const int grp_i = blockIdx.y;
    const CUtensorMap* pA = grp.tmA ? grp.tmA + grp_i : &tmA;
    b_col_off += grp_i * grp.b_stride;
    const int colx = grp_i * grp.col_stride;
    const int kb0 = z * kb_per_split;
    const int nkb = max(0, min(kb_per_split, total_kb - kb0));
    const int pre = min(nkb, STAGES);

    pdl_launch_dependents();        // dependents may be scheduled now; their griddepcontrol.wait still orders the data
    if (threadIdx.x == 0) { tl_mark(0x100 + ep.mode); tl_mark_all(0x100 + ep.mode); }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(pA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(red_full, 1);
        mbar_fence_init();
        // every CTA of the cluster sends all of my R rows x 128 features (fp32), armed before anyone can send
        mbar_arrive_expect_tx(red_full, static_cast<uint32_t>(BPAD) * GEMM_BM * 4);
        // Weights never depend on the previous kernel: their first STAGES tiles go in flight right away
        // (under PDL: while the producer grid is still draining); activations wait for griddepcontrol.wait.
        const uint64_t pol = l2_policy_evict_first();      // weight tiles are read once per step
        for (int i = 0; i < pre; ++i) {
            mbar_arrive_expect_tx(&full_bar[i], L::STAGE_BYTES);
            tma_load_2d_hint(smem + i * L::STAGE_BYTES, pA, &full_bar[i], 0, (mt * total_kb + kb0 + i) * GEMM_BM, pol);
        }
        // keep HBM busy across the kernel boundary: pull the NEXT GEMM's weights into L2 while this one runs
        // (bit 63 of pf_bytes: issue it after this CTA's last weight load instead -- HBM idles during the epilogue)
        if (!(pf_bytes >> 63)) prefetch_l2_slice(pf_ptr, pf_bytes, blockIdx.x, gridDim.x);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    cluster_sync_all();             // CTA-wide sync + "every CTA of the cluster has started and armed its barrier"
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // per-feature epilogue constants (bias / LN-fold vector / next gamma) are model weights, never written by a kernel:
    // the epilogue warps fetch them while the mainloop runs, off the critical tail
    float w_bias = 0.f, w_cv = 0.f, w_gnext = 0.f;
    float e_mean = 0.f, e_rstd = 0.f;                       // LayerNorm statistics of row z*R + lane (BPAD <= 32, see below)
    float e_x[4] = {0.f, 0.f, 0.f, 0.f};                    // residual rows fetched early (R == 4)
    bool e_x_valid = false;
    if (warp >= 2) {
        const int m = m0 + (warp & 3) * 32 + lane;
        if (m < Nout) {
            const float* bias_ptr = grp.bias ? grp.bias[grp_i] : ep.bias;
            w_bias = bias_ptr[m];
            if (ep.ln_fold) w_cv = ep.cvec[m];
            if (ep.emit) w_gnext = ep.next_gamma[m];
        }
    }

    if (warp == 0) {
        // ===== TMA producer ==========================================================================
        if (lane == 0) {
            pdl_wait();
            tl_mark(0x110 + ep.mode);
            tl_mark_all(0x110 + ep.mode);
            for (int i = 0; i < pre; ++i)
                tma_load_2d(smem + i * L::STAGE_BYTES + L::A_BYTES, &tmB, &full_bar[i],
                            b_col_off + (kb0 + i) * GEMM_BK, 0);
            int stage = 0, phase = 0;                       // state after the first `pre` fills
            const uint64_t pol = l2_policy_evict_first();
            for (int i = pre; i < nkb; ++i) {
                mbar_wait(&empty_bar[stage], phase);        // the MMA released this slot
                mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                uint8_t* a = smem + stage * L::STAGE_BYTES;
                tma_load_2d_hint(a, pA, &full_bar[stage], 0, (mt * total_kb + kb0 + i) * GEMM_BM, pol);
                tma_load_2d(a + L::A_BYTES, &tmB, &full_bar[stage], b_col_off + (kb0 + i) * GEMM_BK, 0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (pf_bytes >> 63) prefetch_l2_slice(pf_ptr, pf_bytes & ~(1ull << 63), blockIdx.x, gridDim.x);
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
                for (int k = 0; k < GEMM_BK / 16; ++k)     // +32 B per 16 K-elements inside the swizzle row
                    umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (i | k) != 0);
                umma_commit(&empty_bar[stage]);             // frees the smem slot when the MMAs retire
                if (i == nkb - 1) umma_commit(tmem_full);   // accumulator complete
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ===== epilogue part 1: TMEM -> registers -> reduce-scatter over the cluster (DSMEM) =========
        const int q = warp & 3;                             // TMEM lane quarter owned by this warp
        const int half = (warp - 2) >> 2;                   // which set of 4 epilogue warps (0 unless HALVES == 2)
        const int ml = q * 32 + lane;                       // feature inside the tile
        const int R = BPAD / S;                             // token rows owned by each CTA (power of two)
        const int shR = 31 - __clz(R);
        // While the mainloop runs: everything the epilogue needs from EARLIER kernels.  With a folded LayerNorm that is the
        // mean / rstd of my R rows, from the per-tile partial sums the producer kernel left (fixed tile order): lane r of every
        // warp computes row r (R <= 32 here) and the row groups below fetch it with a shuffle -- no shared memory (the scratch
        // area aliases a pipeline stage that is still in use now), no barrier, and the L2 round trip is off the critical tail.
        if constexpr (BPAD <= 32) {
            pdl_wait();
            if (ep.ln_fold && lane < R) {
                const int row = z * R + lane;
                if (row < nvalid) {
                    float s1 = 0.f, s2 = 0.f;
                    for (int t0 = 0; t0 < ep.stats_tiles; t0 += 16) {
                        float2 v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            v[i] = (t0 + i < ep.stats_tiles)
                                       ? *reinterpret_cast<const float2*>(ep.stats + (static_cast<size_t>(t0 + i) * STATS_ROWS + row) * 2)
                                       : make_float2(0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            s1 += v[i].x;
                            s2 += v[i].y;
                        }
                    }
                    e_mean = s1 * ep.inv_d;
                    e_rstd = 1.0f / sqrtf(fmaxf(s2 * ep.inv_d - e_mean * e_mean, 0.f) + ep.ln_eps);
                }
            }
            // ... and, when this CTA owns a single group of 4 rows (the out-projection and FFN2 at B = 32: R = 4), the
            // residual rows it is going to update
            if (ep.mode == EPI_RESID && R == 4 && m0 + ml < Nout) {
                const int row0 = z * 4;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    e_x[u] = (row0 + u < nvalid) ? ep.x[static_cast<size_t>(row0 + u) * ep.ld_out + m0 + ml] : 0.f;
                e_x_valid = true;
            }
        }
        if (nkb > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        if (threadIdx.x == 64) tl_mark(0x120 + ep.mode);
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t red_addr = smem_u32(red);
        const uint32_t bar_addr = smem_u32(red_full);
        constexpr int CH = BPAD < 32 ? 16 : 32;
        // red layout in the owner: [writer z][feature ml][R rows] -- a thread's R partials for one owner are contiguous
        // (16-byte vector stores when R >= 4).  All Bpad rows are sent, valid or not: the byte count is a constant.
#pragma unroll 1
        for (int c = half * (BPAD / HALVES); c < (half + 1) * (BPAD / HALVES); c += CH) {
            float hi[CH], lo[CH];
            if (nkb > 0) {
                if constexpr (CH == 32) {
                    tmem_ld_32x32(lane_addr + c, hi);
                    tmem_ld_32x32(lane_addr + BPAD + c, lo);
                } else {
                    tmem_ld_32x16(lane_addr + c, hi);
                    tmem_ld_32x16(lane_addr + BPAD + c, lo);
                }
            } else {
#pragma unroll
                for (int j = 0; j < CH; ++j) hi[j] = lo[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
                const int row = c + j;
                if (S == 1) {
                    // no split: the partials stay in this CTA -- plain shared stores, handed over by a named barrier
                    // (a 1-CTA cluster has no peer to st.async to; compute-sanitizer rejects the remote form there)
                    *reinterpret_cast<float4*>(red + (static_cast<size_t>(ml) << shR) + row) =
                        make_float4(hi[j] + lo[j], hi[j + 1] + lo[j + 1], hi[j + 2] + lo[j + 2], hi[j + 3] + lo[j + 3]);
                } else if (R >= 4) {
                    const int owner = row >> shR, rr = row & (R - 1);
                    const uint32_t off = static_cast<uint32_t>((((z << 7) + ml) << shR) + rr) << 2;
                    st_async_f32x4(mapa_u32(red_addr + off, owner), mapa_u32(bar_addr, owner), hi[j] + lo[j],
                                   hi[j + 1] + lo[j + 1], hi[j + 2] + lo[j + 2], hi[j + 3] + lo[j + 3]);
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int owner = (row + u) >> shR, rr = (row + u) & (R - 1);
                        const uint32_t off = static_cast<uint32_t>((((z << 7) + ml) << shR) + rr) << 2;
                        st_async_f32(mapa_u32(red_addr + off, owner), mapa_u32(bar_addr, owner), hi[j + u] + lo[j + u]);
                    }
                }
            }
        }
        tc_fence_before();
        if (threadIdx.x == 64) tl_mark(0x140 + ep.mode);
    }
    if (warp >= 2) {
        if (S == 1) asm volatile("bar.sync 4, %0;" ::"n"(EPI_THREADS) : "memory");   // every epilogue thread stored its partials
        else mbar_wait_bounded(red_full, 0);                // all S partials of my rows have landed (async stores counted)
        if (threadIdx.x == 64) tl_mark(0x150 + ep.mode);
        // ===== epilogue part 2: fixed-order sum of the S partials of my R rows + fused epilogue =======
        // scratch aliases pipeline stage 0: every TMA write / MMA read of this CTA's stages has retired (tmem_full), and
        // peers only ever write into `red`.  (Static __shared__ here would cost the second resident CTA per SM.)
        float* s_mean = reinterpret_cast<float*>(smem);
        float* s_rstd = s_mean + GEMM_BM;
        float (*s_part_all)[4][4][2] = reinterpret_cast<float (*)[4][4][2]>(s_rstd + GEMM_BM);
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        float (*s_part)[4][2] = s_part_all[half];
        const int ml = q * 32 + lane;
        const int et = static_cast<int>(threadIdx.x) - 64;  // index among the epilogue threads
        const int m = m0 + ml;
        const int R = BPAD / S;
        const bool valid_m = m < Nout;
        pdl_wait();                                         // x / stats of earlier kernels are read below
        if (BPAD > 32 && ep.ln_fold) {
            // mean / rstd of my rows from the per-tile partial sums the producer kernel left (fixed tile order)
            for (int rr = et; rr < R; rr += EPI_THREADS) {
                const int row = z * R + rr;
                float mean = 0.f, rstd = 0.f;
                if (row < nvalid) {
                    // 16 tiles = one batch of independent 8-byte loads (a single L2 round trip), summed in tile order
                    float s1 = 0.f, s2 = 0.f;
                    for (int t0 = 0; t0 < ep.stats_tiles; t0 += 16) {
                        float2 v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            v[i] = (t0 + i < ep.stats_tiles)
                                       ? *reinterpret_cast<const float2*>(ep.stats + (static_cast<size_t>(t0 + i) * STATS_ROWS + row) * 2)
                                       : make_float2(0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            s1 += v[i].x;
                            s2 += v[i].y;
                        }
                    }
                    mean = s1 * ep.inv_d;
                    rstd = 1.0f / sqrtf(fmaxf(s2 * ep.inv_d - mean * mean, 0.f) + ep.ln_eps);
                }
                s_mean[rr] = mean;
                s_rstd[rr] = rstd;
            }
            asm volatile("bar.sync 4, %0;" ::"n"(EPI_THREADS) : "memory");
        }
        const float bias = w_bias, cv = w_cv, gnext = w_gnext;
        // the two warp sets (if any) alternate over the groups of 4 rows; each has its own named barrier (2 + half),
        // so they may run a different number of groups
        for (int rr0 = 4 * half; rr0 < R; rr0 += 4 * HALVES) {
            const int row0 = z * R + rr0;
            const int nrows = min(min(4, R - rr0), nvalid - row0);
            if (nrows <= 0) break;
            float sum[4], xnew[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float a = 0.f;
                if (u < nrows) {
                    for (int zz = 0; zz < S; ++zz) a += red[(zz * GEMM_BM + ml) * R + rr0 + u];
                }
                if (ep.ln_fold) {
                    float mean, rstd;
                    if constexpr (BPAD <= 32) {             // (uniform: every lane of the warp takes part in the shuffle)
                        mean = __shfl_sync(0xffffffffu, e_mean, (rr0 + u) & 31);
                        rstd = __shfl_sync(0xffffffffu, e_rstd, (rr0 + u) & 31);
                    } else {
                        mean = s_mean[(rr0 + u) & (R - 1)];
                        rstd = s_rstd[(rr0 + u) & (R - 1)];
                    }
                    if (u < nrows) a = rstd * (a - mean * cv);
                }
                sum[u] = a;
            }
            if (valid_m) apply_epilogue4(ep, row0, nrows, m, sum, bias, xnew, colx, e_x_valid ? e_x : nullptr);
            if (ep.emit) {
                // next GEMM's operand gamma_next * x_new (hi/lo) and this tile's (sum x, sum x^2) per row
                float p1[4], p2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool ok = valid_m && u < nrows;
                    if (ok) {
                        __nv_bfloat16 hi, lo;
                        split_bf16(gnext * xnew[u], hi, lo);
                        ep.next_act[static_cast<size_t>(row0 + u) * ep.next_ld + m] = hi;
                        ep.next_act[static_cast<size_t>(row0 + u + ep.next_bpad) * ep.next_ld + m] = lo;
                    }
                    p1[u] = warp_sum(ok ? xnew[u] : 0.f);
                    p2[u] = warp_sum(ok ? xnew[u] * xnew[u] : 0.f);
                }
                if (lane == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        s_part[q][u][0] = p1[u];
                        s_part[q][u][1] = p2[u];
                    }
                }
                if (half == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
                else asm volatile("bar.sync 3, 128;" ::: "memory");
                if (ml < 8) {
                    const int u = ml >> 1, w = ml & 1;
                    if (u < nrows)
                        ep.stats_out[(static_cast<size_t>(mt) * STATS_ROWS + row0 + u) * 2 + w] =
                            s_part[0][u][w] + s_part[1][u][w] + s_part[2][u][w] + s_part[3][u][w];
                }
                if (half == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
                else asm volatile("bar.sync 3, 128;" ::: "memory");
            }
        }
        if (threadIdx.x == 64) tl_mark(0x160 + ep.mode);
    }
    __syncthreads();
    if (threadIdx.x == 0) { tl_mark(0x130 + ep.mode); tl_mark_all(0x130 + ep.mode); }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
    }
}

void gemm_timeline_set(unsigned long long* buf, unsigned int* cnt) {
    cudaMemcpyToSymbol(g_tl_buf, &buf, sizeof(buf));
    cudaMemcpyToSymbol(g_tl_cnt, &cnt, sizeof(cnt));
}

// ---------------------------------------------------------------------------------------------------
// Bring-up / cross-check kernel: same contract on CUDA cores (one warp per output feature, no split).
// Selected with VCB_GEMM_IMPL=simt; never the default.  It exists so a tcgen05 descriptor bug can be told
// apart from a bug anywhere else in the step.
// ---------------------------------------------------------------------------------------------------
// element (m, k) of the pre-tiled weight layout
__host__ __device__ inline size_t packed_index(int m, int k, int Kdim) {
    const int KB = Kdim / GEMM_BK;
    return ((static_cast<size_t>(m / GEMM_BM) * KB + k / GEMM_BK) * GEMM_BM + (m % GEMM_BM)) * GEMM_BK + (k % GEMM_BK);
}

__global__ void gemm_w_xT_simt(const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ X,
                               const GemmEpilogue ep, int Nout, int Kdim, int ldx, int bpad, int b_col_off, int nvalid) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= Nout) return;
    for (int j = 0; j < nvalid; ++j) {
        float acc = 0.f;
        for (int k = lane; k < Kdim; k += 32) {
            const float w = __bfloat162float(W[packed_index(warp, k, Kdim)]);
            const float xh = __bfloat162float(X[static_cast<size_t>(j) * ldx + b_col_off + k]);
            const float xl = __bfloat162float(X[static_cast<size_t>(j + bpad) * ldx + b_col_off + k]);
            acc = fmaf(w, xh, acc);
            acc = fmaf(w, xl, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            if (ep.ln_fold) {
                float s1 = 0.f, s2 = 0.f;
                for (int t = 0; t < ep.stats_tiles; ++t) {
                    s1 += ep.stats[(static_cast<size_t>(t) * STATS_ROWS + j) * 2];
                    s2 += ep.stats[(static_cast<size_t>(t) * STATS_ROWS + j) * 2 + 1];
                }
                const float mean = s1 * ep.inv_d;
                const float rstd = 1.0f / sqrtf(fmaxf(s2 * ep.inv_d - mean * mean, 0.f) + ep.ln_eps);
                acc = rstd * (acc - mean * ep.cvec[warp]);
            }
            const float sum[4] = {acc, 0.f, 0.f, 0.f};
            float xnew[4];
            apply_epilogue4(ep, j, 1, warp, sum, ep.bias[warp], xnew);
            if (ep.emit) {
                __nv_bfloat16 hi, lo;
                split_bf16(ep.next_gamma[warp] * xnew[0], hi, lo);
                ep.next_act[static_cast<size_t>(j) * ep.next_ld + warp] = hi;
                ep.next_act[static_cast<size_t>(j + ep.next_bpad) * ep.next_ld + warp] = lo;
                // one "tile" per feature group of 128, accumulated with atomics (cross-check path only)
                atomicAdd(&ep.stats_out[(static_cast<size_t>(warp / GEMM_BM) * STATS_ROWS + j) * 2], xnew[0]);
                atomicAdd(&ep.stats_out[(static_cast<size_t>(warp / GEMM_BM) * STATS_ROWS + j) * 2 + 1], xnew[0] * xnew[0]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// fp32 row-major [N, K] -> bf16 tiles [ceil(N/128)][K/64][128][64], rows beyond N zero-filled
__global__ void pack_weight_tiles_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int Kdim) {
    const size_t total = static_cast<size_t>((N + GEMM_BM - 1) / GEMM_BM) * GEMM_BM * Kdim;
    const int KB = Kdim / GEMM_BK;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % GEMM_BK);
        const int r = static_cast<int>((i / GEMM_BK) % GEMM_BM);
        const size_t blk = i / (GEMM_BK * GEMM_BM);
        const int kb = static_cast<int>(blk % KB);
        const int mt = static_cast<int>(blk / KB);
        const int m = mt * GEMM_BM + r, k = kb * GEMM_BK + c;
        out[i] = (m < N) ? __float2bfloat16_rn(in[static_cast<size_t>(m) * Kdim + k]) : __float2bfloat16_rn(0.f);
    }
}

// LayerNorm folding vectors from the (bf16, pre-tiled) weight: cvec[m] = sum_k gamma[k] W[m,k],
// bprime[m] = bias[m] + sum_k beta[k] W[m,k]; one warp per output feature, fp32.
__global__ void ln_fold_vectors_kernel(const __nv_bfloat16* __restrict__ Wp, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ bias,
                                       float* __restrict__ cvec, float* __restrict__ bprime, int N, int Kdim) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (m >= N) return;
    float c = 0.f, b = 0.f;
    for (int k = lane; k < Kdim; k += 32) {
        const float w = __bfloat162float(Wp[packed_index(m, k, Kdim)]);
        c = fmaf(gamma[k], w, c);
        b = fmaf(beta[k], w, b);
    }
    c = warp_sum(c);
    b = warp_sum(b);
    if (lane == 0) {
        cvec[m] = c;
        bprime[m] = bias[m] + b;
    }
}

int ln_fold_vectors(const __nv_bfloat16* Wp, const float* gamma, const float* beta, const float* bias, float* cvec,
                    float* bprime, int N, int Kdim) {
    ln_fold_vectors_kernel<<<(N * 32 + 255) / 256, 256>>>(Wp, gamma, beta, bias, cvec, bprime, N, Kdim);
    VCB_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t packed_weight_elems(int N, int Kdim) { return static_cast<size_t>((N + GEMM_BM - 1) / GEMM_BM) * GEMM_BM * Kdim; }

int pack_weight(const float* w_f32_dev, __nv_bfloat16* out, int N, int Kdim, CUtensorMap* tm) {
    if (Kdim % GEMM_BK) {
        set_error("pack_weight: K=%d not a multiple of %d", Kdim, GEMM_BK);
        return -1;
    }
    pack_weight_tiles_kernel<<<1024, 256>>>(w_f32_dev, out, N, Kdim);
    VCB_CUDA_OK(cudaGetLastError());
    const uint64_t rows = packed_weight_elems(N, Kdim) / GEMM_BK;       // 128-byte rows, 128 per tile
    return make_tmap_bf16_2d(tm, out, rows, GEMM_BK, GEMM_BK, GEMM_BM);
}

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
        VCB_CUDA_OK(cudaFuncSetAttribute(gemm_w_xT_cluster<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL));
        attr_set = true;
        if (getenv("VCB_DEBUG_OCC")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gemm_w_xT_cluster<BN, STAGES>, gemm_threads(BN), L::TOTAL);
            fprintf(stderr, "[vcb] gemm<%d,%d>: %d B smem, %d CTAs/SM (occupancy API)\n", BN, STAGES, L::TOTAL, nb);
        }
    }
    const int tiles = (g.Nout + GEMM_BM - 1) / GEMM_BM;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles * g.splits, g.grp.tmA ? g.groups : 1, 1);
    cfg.blockDim = dim3(gemm_threads(BN));
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = g.splits;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g.pdl ? 2 : 1;
    const int total_kb = g.Kdim / GEMM_BK;
    const int kbps = (total_kb + g.splits - 1) / g.splits;
    VCB_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_w_xT_cluster<BN, STAGES>, *g.tmA, *g.tmB, g.ep, g.Nout, total_kb, kbps,
                                   g.b_col_off, g.nvalid, g.pf_ptr, static_cast<unsigned long long>(g.pf_bytes), g.grp));
    return 0;
}

int gemm_launch(const GemmCall& g, cudaStream_t st) {
    if (g.Kdim % GEMM_BK != 0) {
        set_error("gemm: K=%d not a multiple of %d", g.Kdim, GEMM_BK);
        return -1;
    }
    if (g.simt) {
        dim3 grid((g.Nout * 32 + 255) / 256, 1, 1);
        gemm_w_xT_simt<<<grid, 256, 0, st>>>(g.W, g.X, g.ep, g.Nout, g.Kdim, g.ldx, g.bpad, g.b_col_off, g.nvalid);
        VCB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const int total_kb = g.Kdim / GEMM_BK;
    const int kbps = (total_kb + g.splits - 1) / g.splits;
    if (g.splits < 1 || g.splits > 8 || (g.splits & (g.splits - 1)) || g.bpad % g.splits ||
        (g.splits - 1) * kbps >= total_kb) {
        set_error("gemm: bad split count %d for %d k-blocks, bpad %d", g.splits, total_kb, g.bpad);
        return -1;
    }
    switch (g.bpad) {
        case 16: return launch_one<32, 4>(g, st);
        case 32:
            if (g.stages == 2) return launch_one<64, 2>(g, st);
            if (g.stages == 3) return launch_one<64, 3>(g, st);
            if (g.stages == 6) return launch_one<64, 6>(g, st);
            if (g.stages == 8) return launch_one<64, 8>(g, st);
            return launch_one<64, 4>(g, st);     // 4 x 24 KB + 16 KB: two CTAs per SM stay resident (PDL overlap)
        case 64: return launch_one<128, 3>(g, st);
        case 128: return launch_one<256, 3>(g, st);
        default: set_error("gemm: unsupported bpad %d", g.bpad); return -1;
    }
}

// Cluster size (= K splits), a power of two <= 8.  Measured on B200 (scripts/bench_gemm.py, profiles/r01_gemm_micro.txt):
// the kernel has a ~7 us latency floor, so the grid should reach >= ~1.3 CTAs per SM in ONE wave of co-resident CTAs
// (2 per SM) while every split keeps >= 4 k-blocks.
int gemm_pick_splits(int Nout, int Kdim, int num_sms) {
    const int tiles = (Nout + GEMM_BM - 1) / GEMM_BM;
    const int total_kb = Kdim / GEMM_BK;
    int s = 1;
    while (s < 8 && tiles * s < num_sms && total_kb / (2 * s) >= 4) s *= 2;
    while (s > 1 && (s - 1) * ((total_kb + s - 1) / s) >= total_kb) s /= 2;
    return s;
}

}  // namespace vcb
