// Persistent multi-phase GEMM chain for the decode step: up to 8 dependent GEMMs (out-proj -> FFN1 -> FFN2 -> next
// layer's QKV, or the logit heads) run inside ONE kernel on a resident grid of 8-CTA clusters.
//
// Why: the device timeline (profiles/r01_timeline_v4_lnfold.txt) shows that a chain of separate GEMM kernels, even with
// programmatic dependent launch, pays ~2-3 us of completion->wait latency plus a pipeline ramp per kernel -- more than the
// 1.3-5 us of HBM streaming each GEMM needs.  Here the kernel boundary becomes a device-wide barrier (one atomic + an
// acquire spin), the TMA producer runs ahead across tiles AND phases (the next phase's weight tiles are already in
// flight while the barrier resolves), TMEM accumulators are double buffered so the MMAs of tile t+1 overlap the epilogue of
// tile t, and barriers / TMEM / tensor-map fetches are paid once per chain instead of once per GEMM.
//
// Per phase the math is exactly gemm_w_xT_cluster's (gemm_tcgen05.cu): weight tile 128 x 64 via TMA SWIZZLE_128B from the
// pre-tiled layout, activations hi/lo as 2*Bpad UMMA columns, fp32 accumulation in TMEM, split-K over the 8 CTAs of a cluster
// with a DSMEM reduce-scatter (fixed-order sums), fused epilogues incl. folded LayerNorm.  Differences:
//   * cluster-wide barriers inside the loop are replaced by per-buffer mbarriers signalled with remote arrives
//     (mbarrier.arrive.release.cluster), so only the epilogue warps synchronise across the cluster;
//   * `red` (the DSMEM landing zone) and the TMEM accumulator are double buffered by tile parity.
#include "vcb_internal.h"

#include <algorithm>

namespace vcb {

static constexpr int CH_BM = 128, CH_BK = 64, CH_THREADS = 192, CH_CLUSTER = 8;

template <int BN, int STAGES>
struct ChainSmem {
    static constexpr int A_BYTES = CH_BM * CH_BK * 2;
    static constexpr int B_BYTES = BN * CH_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int RED_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int RED_BYTES = (BN / 2) * CH_BM * 4;                 // one buffer: [8][R][128] fp32
    static constexpr int BAR_OFFSET = RED_OFFSET + RED_BYTES;             // single landing buffer (keeps 2 CTAs / SM)
    static constexpr int NBARS = 2 * STAGES + 6;                           // full, empty, tfull[2], tempty[2], redfull, redempty
    static constexpr int TOTAL = BAR_OFFSET + NBARS * 8 + 16;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Device-wide barrier wait with a watchdog: if the grid is not fully resident (the barrier can never complete) the wait
// gives up after ~0.5 s and raises a flag (ctr[1]) instead of hanging the GPU; the host checks it in vcb_poll.
__device__ __forceinline__ void grid_wait(unsigned int* ctr, unsigned int target) {
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (static_cast<int>(ld_acquire_u32(ctr) - target) < 0) {
        __nanosleep(20);
        if ((++spins & 0x3ff) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 500000000ull) {
                atomicExch(ctr + 1, 1u);
                break;
            }
        }
    }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_remote_f32(uint32_t remote_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint32_t ch_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void ch_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// identical semantics to apply_epilogue4 in gemm_tcgen05.cu (kept local: separate translation unit, no -rdc)
__device__ __forceinline__ void chain_epilogue4(const GemmEpilogue& ep, int row0, int nrows, int m, const float (&sum)[4],
                                                float bias, float (&xnew)[4]) {
    switch (ep.mode) {
        case EPI_QKV: {
            int pos[4], slot[4], page[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                pos[u] = (u < nrows) ? ep.row_pos[row0 + u] : -1;
                slot[u] = (u < nrows) ? ep.row_slot[row0 + u] : 0;
            }
            const int part = m / ep.d, cc = m - part * ep.d;
            if (part == 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (pos[u] >= 0) ep.qbuf[static_cast<size_t>(row0 + u) * ep.d + cc] = sum[u] + bias;
                return;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                page[u] = (pos[u] >= 0) ? ep.page_table[slot[u] * ep.max_pages + pos[u] / ep.page_size] : 0;
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
            for (int u = 0; u < 4; ++u) xv[u] = (u < nrows) ? ep.x[static_cast<size_t>(row0 + u) * ep.ld_out + m] : 0.f;
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
                if (u < nrows) ep.out[static_cast<size_t>(row0 + u) * ep.ld_out + ep.col_off + m] = sum[u] + bias;
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(CH_THREADS) gemm_chain_kernel(const __grid_constant__ ChainArgs A) {
    using L = ChainSmem<BN, STAGES>;
    constexpr int BPAD = BN / 2;
    constexpr int S = CH_CLUSTER;
    constexpr int R = BPAD / S;                                  // token rows owned by each CTA of the cluster
    static_assert(R >= 1 && R <= 4, "chain kernel: Bpad in {8..32} with 8-CTA clusters");
    extern __shared__ __align__(1024) uint8_t smem[];
    float* red = reinterpret_cast<float*>(smem + L::RED_OFFSET);           // [S*R][128]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull = empty_bar + STAGES;                                    // [2] accumulator ready
    uint64_t* tempty = tfull + 2;                                            // [2] accumulator drained
    uint64_t* redfull = tempty + 2;                                          // all 8 partials of my rows landed
    uint64_t* redempty = redfull + 1;                                        // all 8 owners consumed my previous partials
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(redempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int z = static_cast<int>(ch_cluster_ctarank());
    const int cid = blockIdx.x / S, ncl = gridDim.x / S;
    const int nvalid = A.nvalid;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);
        }
        mbar_init(redfull, S * 4);                               // one arrive per epilogue warp of every CTA in the cluster
        mbar_init(redempty, S * 4);
        mbar_fence_init();
        for (int p = 0; p < A.nphases; ++p) tma_prefetch_desc(&A.ph[p].tmA);
        for (int b = 0; b < 4; ++b) tma_prefetch_desc(&A.tmB[b]);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 2 * BN < 32 ? 32 : 2 * BN);
        tmem_relinquish();
    }
    tc_fence_before();
    ch_cluster_sync();            // CTA sync + every CTA of the cluster is running and its mbarriers are initialised
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto my_tiles = [&](int tiles) { return tiles > cid ? (tiles - cid + ncl - 1) / ncl : 0; };

    if (warp == 0) {
        // ===== TMA producer: streams weight / activation tiles for every tile of every phase, in order ===================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            int it = 0;
            for (int p = 0; p < A.nphases; ++p) {
                const ChainPhase& P = A.ph[p];
                const int tiles = (P.Nout + CH_BM - 1) / CH_BM;
                const int kbps = P.total_kb / S;
                const int n = my_tiles(tiles) * kbps;
                const int pre = min(n, STAGES);
                auto a_row = [&](int i) { return ((cid + (i / kbps) * ncl) * P.total_kb + z * kbps + (i % kbps)) * CH_BM; };
                auto b_col = [&](int i) { return P.b_col_off + (z * kbps + (i % kbps)) * CH_BK; };
                // weights of the first tiles go in flight before this phase's activations exist
                for (int i = 0; i < pre; ++i) {
                    const int j = it + i, s = j % STAGES;
                    if (j >= STAGES) mbar_wait(&empty_bar[s], ((j / STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
                    tma_load_2d_hint(smem + s * L::STAGE_BYTES, &P.tmA, &full_bar[s], 0, a_row(i), pol);
                }
                if (p == 0) pdl_wait();
                else grid_wait(A.ctr, A.epoch + static_cast<unsigned int>(p) * gridDim.x);
                fence_proxy_async_all();         // other CTAs' generic-proxy stores -> visible to this thread's TMA loads
                for (int i = 0; i < pre; ++i) {
                    const int s = (it + i) % STAGES;
                    tma_load_2d(smem + s * L::STAGE_BYTES + L::A_BYTES, &A.tmB[P.b_map], &full_bar[s], b_col(i), 0);
                }
                for (int i = pre; i < n; ++i) {
                    const int j = it + i, s = j % STAGES;
                    if (j >= STAGES) mbar_wait(&empty_bar[s], ((j / STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
                    tma_load_2d_hint(smem + s * L::STAGE_BYTES, &P.tmA, &full_bar[s], 0, a_row(i), pol);
                    tma_load_2d(smem + s * L::STAGE_BYTES + L::A_BYTES, &A.tmB[P.b_map], &full_bar[s], b_col(i), 0);
                }
                it += n;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer ======================================================================================================
        constexpr uint32_t idesc = umma_idesc_bf16_f32(CH_BM, BN);
        int it = 0, tc = 0;
        for (int p = 0; p < A.nphases; ++p) {
            const ChainPhase& P = A.ph[p];
            const int nt = my_tiles((P.Nout + CH_BM - 1) / CH_BM);
            const int kbps = P.total_kb / S;
            for (int t = 0; t < nt; ++t, ++tc) {
                const int buf = tc & 1;
                if (tc >= 2) mbar_wait(&tempty[buf], ((tc >> 1) - 1) & 1);          // epilogue drained this accumulator
                tc_fence_after();
                for (int kb = 0; kb < kbps; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
                        const uint64_t a_desc = umma_desc_kmajor_sw128(a_addr);
                        const uint64_t b_desc = umma_desc_kmajor_sw128(a_addr + L::A_BYTES);
#pragma unroll
                        for (int k = 0; k < CH_BK / 16; ++k)
                            umma_bf16(tmem_base + buf * BN, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                        umma_commit(&empty_bar[s]);
                        if (kb == kbps - 1) umma_commit(&tfull[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue warps ==================================================================================================
        const int q = warp & 3;
        const int ml = q * 32 + lane;
        int tc = 0;
        for (int p = 0; p < A.nphases; ++p) {
            const ChainPhase& P = A.ph[p];
            const GemmEpilogue& ep = P.ep;
            const int tiles = (P.Nout + CH_BM - 1) / CH_BM;
            const int nt = my_tiles(tiles);
            // data written by earlier phases / kernels (x, stats, page tables) must be visible to THESE threads
            if (p == 0) pdl_wait();
            else grid_wait(A.ctr, A.epoch + static_cast<unsigned int>(p) * gridDim.x);
            // LayerNorm statistics of my R rows (fixed tile order), identical for every tile of the phase
            float mean[R], rstd[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                mean[rr] = 0.f;
                rstd[rr] = 0.f;
                const int row = z * R + rr;
                if (ep.ln_fold && row < nvalid) {
                    float s1 = 0.f, s2 = 0.f;
                    for (int t = 0; t < ep.stats_tiles; ++t) {
                        s1 += ep.stats[(static_cast<size_t>(t) * STATS_ROWS + row) * 2];
                        s2 += ep.stats[(static_cast<size_t>(t) * STATS_ROWS + row) * 2 + 1];
                    }
                    mean[rr] = s1 * ep.inv_d;
                    rstd[rr] = 1.0f / sqrtf(fmaxf(s2 * ep.inv_d - mean[rr] * mean[rr], 0.f) + ep.ln_eps);
                }
            }
            for (int t = 0; t < nt; ++t, ++tc) {
                const int buf = tc & 1;
                const int mt = cid + t * ncl;
                const int m = mt * CH_BM + ml;
                const bool valid_m = m < P.Nout;
                // operands that do not depend on the reduction: issue the loads early
                const float bias = valid_m ? ep.bias[m] : 0.f;
                const float cv = (ep.ln_fold && valid_m) ? ep.cvec[m] : 0.f;
                const float gnext = (ep.emit && valid_m) ? ep.next_gamma[m] : 0.f;
                // ---- TMEM -> registers ------------------------------------------------------------------------------------
                mbar_wait(&tfull[buf], (tc >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16);
                float part[BPAD];
                if constexpr (BPAD == 32) {
                    float hi[32], lo[32];
                    tmem_ld_32x32(taddr, hi);
                    tmem_ld_32x32(taddr + BPAD, lo);
#pragma unroll
                    for (int j = 0; j < 32; ++j) part[j] = hi[j] + lo[j];
                } else {
                    float hi[16], lo[16];
                    tmem_ld_32x16(taddr, hi);
                    tmem_ld_32x16(taddr + BPAD, lo);
#pragma unroll
                    for (int j = 0; j < 16; ++j) part[j] = hi[j] + lo[j];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);                       // the MMA warp may reuse this accumulator
                // ---- reduce-scatter over the cluster: my partial of owner o's rows -> o's red ------------------------------
                if (tc >= 1) mbar_wait_cluster(redempty, (tc - 1) & 1);         // every owner is done with my previous partials
                const uint32_t red_local = smem_u32(red);
#pragma unroll
                for (int o = 0; o < S; ++o) {
                    const uint32_t base = mapa_u32(red_local, o) + static_cast<uint32_t>(((z * R) * CH_BM + ml) * 4);
#pragma unroll
                    for (int rr = 0; rr < R; ++rr)
                        if (o * R + rr < nvalid) st_remote_f32(base + rr * CH_BM * 4, part[o * R + rr]);
                }
                __syncwarp();
                if (lane == 0) {
                    const uint32_t bar_local = smem_u32(redfull);
#pragma unroll
                    for (int o = 0; o < S; ++o) mbar_arrive_remote(mapa_u32(bar_local, o));
                }
                // ---- all 8 partials of my rows have landed: fixed-order sum + fused epilogue ------------------------------------
                mbar_wait_cluster(redfull, tc & 1);
                const float* rbuf = red;
                const int row0 = z * R;
                const int nrows = min(R, nvalid - row0);
                if (nrows <= 0) {                                               // idle owner: still release the landing buffer
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t eb = smem_u32(redempty);
#pragma unroll
                        for (int o = 0; o < S; ++o) mbar_arrive_remote(mapa_u32(eb, o));
                    }
                } else {
                    float sum[4] = {0.f, 0.f, 0.f, 0.f}, xnew[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int u = 0; u < R; ++u) {
                        float a = 0.f;
                        if (u < nrows) {
#pragma unroll
                            for (int zz = 0; zz < S; ++zz) a += rbuf[(zz * R + u) * CH_BM + ml];
                            if (ep.ln_fold) a = rstd[u] * (a - mean[u] * cv);
                        }
                        sum[u] = a;
                    }
                    __syncwarp();
                    if (lane == 0) {                                            // my reads of red are done: writers may overwrite it
                        const uint32_t eb = smem_u32(redempty);
#pragma unroll
                        for (int o = 0; o < S; ++o) mbar_arrive_remote(mapa_u32(eb, o));
                    }
                    if (valid_m) chain_epilogue4(ep, row0, nrows, m, sum, bias, xnew);
                    if (ep.emit) {
                        // next GEMM's operand gamma_next * x_new (hi/lo) and this tile's (sum x, sum x^2) per row.
                        // Cross-warp combine through a 4x4x2 scratch at the end of this buffer's red area is not needed:
                        // each warp publishes its own partial into stats_out via a 4-slot sub-tile (q), summed by the reader.
#pragma unroll
                        for (int u = 0; u < R; ++u) {
                            const bool ok = valid_m && u < nrows;
                            if (ok) {
                                __nv_bfloat16 hi, lo;
                                split_bf16(gnext * xnew[u], hi, lo);
                                ep.next_act[static_cast<size_t>(row0 + u) * ep.next_ld + m] = hi;
                                ep.next_act[static_cast<size_t>(row0 + u + ep.next_bpad) * ep.next_ld + m] = lo;
                            }
                            const float p1 = warp_sum(ok ? xnew[u] : 0.f);
                            const float p2 = warp_sum(ok ? xnew[u] * xnew[u] : 0.f);
                            if (lane == 0 && u < nrows) {
                                // sub-tile index = 4*mt + q: the consumer sums 4*tiles entries in fixed order
                                float* so = ep.stats_out + (static_cast<size_t>(4 * mt + q) * STATS_ROWS + row0 + u) * 2;
                                so[0] = p1;
                                so[1] = p2;
                            }
                        }
                    }
                }
            }
            // ---- phase end: publish this CTA's global writes, arrive on the device-wide barrier -------------------------------
            __threadfence();
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (threadIdx.x == 64) {
                fence_proxy_async_all();
                atomicAdd(A.ctr, 1u);
            }
        }
    }
    // no CTA may exit while a peer can still touch its shared memory (the trailing redempty arrives have no waiter)
    ch_cluster_sync();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN < 32 ? 32 : 2 * BN);
    }
}

template <int BN, int STAGES>
static int chain_launch_t(const ChainArgs& args, int nclusters, int pdl, cudaStream_t st) {
    using L = ChainSmem<BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        VCB_CUDA_OK(cudaFuncSetAttribute(gemm_chain_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        VCB_CUDA_OK(cudaFuncSetAttribute(gemm_chain_kernel<BN, STAGES>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nclusters * CH_CLUSTER);
    cfg.blockDim = dim3(CH_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CH_CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 2 : 1;
    VCB_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_chain_kernel<BN, STAGES>, args));
    return 0;
}

// How many 8-CTA clusters of the chain kernel can be resident at once (the device-wide barrier needs all of them).
int chain_max_clusters(int bpad) {
    int n = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(32 * CH_CLUSTER);
    cfg.blockDim = dim3(CH_THREADS);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CH_CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (bpad == 32) {
        using L = ChainSmem<64, 4>;
        cudaFuncSetAttribute(gemm_chain_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        cudaFuncSetAttribute(gemm_chain_kernel<64, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cfg.dynamicSmemBytes = L::TOTAL;
        if (cudaOccupancyMaxActiveClusters(&n, gemm_chain_kernel<64, 4>, &cfg) != cudaSuccess) n = 0;
    } else {
        using L = ChainSmem<32, 4>;
        cudaFuncSetAttribute(gemm_chain_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        cudaFuncSetAttribute(gemm_chain_kernel<32, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cfg.dynamicSmemBytes = L::TOTAL;
        if (cudaOccupancyMaxActiveClusters(&n, gemm_chain_kernel<32, 4>, &cfg) != cudaSuccess) n = 0;
    }
    return n;
}

int chain_launch(const ChainArgs& args, int bpad, int nclusters, int pdl, cudaStream_t st) {
    for (int p = 0; p < args.nphases; ++p)
        if (args.ph[p].total_kb % CH_CLUSTER) {
            set_error("chain: K blocks (%d) of phase %d not divisible by the cluster size", args.ph[p].total_kb, p);
            return -1;
        }
    if (bpad == 32) return chain_launch_t<64, 4>(args, nclusters, pdl, st);
    if (bpad == 16) return chain_launch_t<32, 4>(args, nclusters, pdl, st);
    set_error("chain: unsupported bpad %d", bpad);
    return -1;
}

}  // namespace vcb
