// EXPERIMENTAL (off by default, VCB_PREFILL_ATT_GROUP=4; written after the round's last GPU session, not yet run on
// hardware): prefill attention over the paged KV cache with G consecutive prompt rows per work item.
//
// attn_rows_kernel (lm_kernels.cuh) treats every prompt row as its own work item, so in the wide prefill path each of
// the 231 rows of an utterance streams the same K/V pages again: 15 GB of L2 traffic per layer for a 7392-row chunk
// (measured: 23 of the 52 ms of a prefill).  Here an item is (G consecutive rows, head): the pages of the item's longest
// context are streamed once through the same TMA-bulk / mbarrier ring and every page is scored against the G query rows
// (per-row causal mask key <= pos[row]), i.e. 1/G of the traffic.  Rows of a group that belong to different utterances
// (a group straddling a prompt boundary) are handled as consecutive "segments" of equal slot, each with its own page walk.
// Numerics per row are those of attn_rows_kernel: fp32 QK, online softmax per 64-key page, fp32 PV, same page order.
#pragma once

#include "lm_kernels.cuh"

namespace vcb {

template <typename KVT, int HD, int G>
struct GroupAttSmem {
    static constexpr int PAGE_BYTES = KV_PAGE * HD * sizeof(KVT);
    static constexpr int OFF_V = ATT_STAGES * PAGE_BYTES;
    static constexpr int OFF_SC = 2 * ATT_STAGES * PAGE_BYTES;
    static constexpr int OFF_PW = OFF_SC + 2 * G * KV_PAGE * 4;                 // scores [parity][G][64]
    static constexpr int OFF_RED = OFF_PW + ATT_CWARPS * G * KV_PAGE * 4;       // exp values [warp][G][64]
    static constexpr int OFF_BAR = OFF_RED + ATT_CWARPS * G * HD * 4;           // partial outputs [warp][G][HD]
    static constexpr int TOTAL = OFF_BAR + 2 * ATT_STAGES * 8 + 128;
};

// Segment walk shared by the producer and the consumers: rows [a, b) of the group share a slot; returns b (== a when the
// row at `a` is not a live row) and the largest position of the segment.
template <int G>
__device__ __forceinline__ int group_segment(const int (&slot)[G], const int (&pos)[G], int a, int& pmax) {
    pmax = -1;
    if (pos[a] < 0) return a;
    int b = a;
#pragma unroll
    for (int j = 0; j < G; ++j)
        if (j >= a && j == b && pos[j] >= 0 && slot[j] == slot[a]) {
            pmax = max(pmax, pos[j]);
            b = j + 1;
        }
    return b;
}

template <typename KVT, int HD, int G>
__global__ void __launch_bounds__(ATT_THREADS + 32)
attn_group_kernel(const float* __restrict__ qbuf, const KVT* __restrict__ kpool, const KVT* __restrict__ vpool,
                  const int* __restrict__ page_table, int max_pages, const int* __restrict__ row_slot,
                  const int* __restrict__ row_pos, int H, __nv_bfloat16* __restrict__ act, int ld_act, int bpad,
                  float scale, int rows) {
    using L = GroupAttSmem<KVT, HD, G>;
    constexpr int LPT = HD / 8;          // lanes per key in QK
    constexpr int TPW = 32 / LPT;        // keys per warp iteration
    constexpr int DPT = HD / 32;         // output dims per lane in PV
    constexpr int KPW = KV_PAGE / ATT_CWARPS;
    extern __shared__ __align__(128) uint8_t gatt_smem[];
    KVT* sK = reinterpret_cast<KVT*>(gatt_smem);
    KVT* sV = reinterpret_cast<KVT*>(gatt_smem + L::OFF_V);
    float* sc_all = reinterpret_cast<float*>(gatt_smem + L::OFF_SC);
    float* pw = reinterpret_cast<float*>(gatt_smem + L::OFF_PW);
    float* red = reinterpret_cast<float*>(gatt_smem + L::OFF_RED);
    uint64_t* full = reinterpret_cast<uint64_t*>(gatt_smem + L::OFF_BAR);
    uint64_t* empty = full + ATT_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < ATT_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ATT_CWARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    const int n_groups = (rows + G - 1) / G;
    const int n_items = n_groups * H;

    if (warp == ATT_CWARPS) {
        // ===== producer: the pages of every segment of every item of this CTA, in order ========================
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int g = item / H, h = item - g * H;
                int slot[G], pos[G];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int r = g * G + j;
                    pos[j] = r < rows ? row_pos[r] : -1;
                    slot[j] = r < rows ? row_slot[r] : -1;
                }
                for (int a = 0; a < G;) {
                    int pmax;
                    const int b = group_segment<G>(slot, pos, a, pmax);
                    if (b == a) { ++a; continue; }
                    const int* pt = page_table + slot[a] * max_pages;
                    const int npages = pmax / KV_PAGE + 1;
                    for (int p = 0; p < npages; ++p, ++it) {
                        const int s = it % ATT_STAGES;
                        if (it >= ATT_STAGES) mbar_wait(&empty[s], ((it / ATT_STAGES) - 1) & 1);
                        const size_t off = (static_cast<size_t>(pt[p]) * H + h) * KV_PAGE * HD;
                        mbar_arrive_expect_tx(&full[s], 2 * L::PAGE_BYTES);
                        tma_bulk_g2s(sK + s * KV_PAGE * HD, kpool + off, L::PAGE_BYTES, &full[s]);
                        tma_bulk_g2s(sV + s * KV_PAGE * HD, vpool + off, L::PAGE_BYTES, &full[s]);
                    }
                    a = b;
                }
            }
        }
        return;
    }

    // ===== consumers ==================================================================================
    const int sub = lane % LPT;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int g = item / H, h = item - g * H;
        int slot[G], pos[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int r = g * G + j;
            pos[j] = r < rows ? row_pos[r] : -1;
            slot[j] = r < rows ? row_slot[r] : -1;
        }
        for (int a = 0; a < G;) {
            int pmax;
            const int b = group_segment<G>(slot, pos, a, pmax);
            if (b == a) { ++a; continue; }
            // rows [a, b) are live in this segment
            float q[G][8], m_run[G], l_run[G], acc[G][DPT];
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const bool on = j >= a && j < b;
                const float* qp = qbuf + (static_cast<size_t>(g * G + (on ? j : a)) * H + h) * HD + sub * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) q[j][i] = qp[i];
                m_run[j] = -INFINITY;
                l_run[j] = 0.f;
#pragma unroll
                for (int i = 0; i < DPT; ++i) acc[j][i] = 0.f;
            }
            const int npages = pmax / KV_PAGE + 1;
            for (int p = 0; p < npages; ++p, ++it) {
                const int s = it % ATT_STAGES;
                mbar_wait(&full[s], (it / ATT_STAGES) & 1);
                const KVT* K = sK + s * KV_PAGE * HD;
                const KVT* V = sV + s * KV_PAGE * HD;
                float* sc = sc_all + (it & 1) * G * KV_PAGE;
                // ---- scores of this warp's KPW keys against the G query rows
#pragma unroll
                for (int itq = 0; itq < KPW / TPW; ++itq) {
                    const int t = warp * KPW + itq * TPW + lane / LPT;
                    float kv[8];
                    load_kv_vec<KVT, 8>(K + t * HD + sub * 8, kv);
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        float dsum = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) dsum = fmaf(q[j][i], kv[i], dsum);
#pragma unroll
                        for (int o = LPT / 2; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
                        if (sub == 0) sc[j * KV_PAGE + t] = (j >= a && j < b && p * KV_PAGE + t <= pos[j]) ? dsum * scale : -INFINITY;
                    }
                }
                named_bar_sync(1, ATT_THREADS);
                // ---- online softmax per live row (every warp redundantly over all 64 scores: identical m, l)
                float corr[G];
                float* mypw = pw + warp * G * KV_PAGE;
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    corr[j] = 1.f;
                    if (j < a || j >= b) continue;                      // warp-uniform
                    const float s0 = sc[j * KV_PAGE + lane], s1 = sc[j * KV_PAGE + lane + 32];
                    const float m_new = fmaxf(m_run[j], warp_max(fmaxf(s0, s1)));   // finite from page 0 on (key 0 is live)
                    corr[j] = expf(m_run[j] - m_new);
                    const float e0 = expf(s0 - m_new), e1 = expf(s1 - m_new);
                    mypw[j * KV_PAGE + lane] = e0;
                    mypw[j * KV_PAGE + lane + 32] = e1;
                    l_run[j] = l_run[j] * corr[j] + warp_sum(e0 + e1);
                    m_run[j] = m_new;
                }
                __syncwarp();
                // ---- PV for this warp's KPW keys
#pragma unroll
                for (int j = 0; j < G; ++j)
#pragma unroll
                    for (int i = 0; i < DPT; ++i) acc[j][i] *= corr[j];
#pragma unroll
                for (int tt = 0; tt < KPW; ++tt) {
                    const int t = warp * KPW + tt;
                    float vv[DPT];
                    load_kv_vec<KVT, DPT>(V + t * HD + lane * DPT, vv);
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        if (j < a || j >= b) continue;
                        const float pt_ = mypw[j * KV_PAGE + t];
#pragma unroll
                        for (int i = 0; i < DPT; ++i) acc[j][i] = fmaf(pt_, vv[i], acc[j][i]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);                  // this warp is done with stage s
            }
            // ---- combine the warps' partial outputs of the segment's rows
#pragma unroll
            for (int j = 0; j < G; ++j)
#pragma unroll
                for (int i = 0; i < DPT; ++i) red[(warp * G + j) * HD + lane * DPT + i] = acc[j][i];
            named_bar_sync(1, ATT_THREADS);
            const size_t ocol = static_cast<size_t>(h) * HD;
            for (int idx = threadIdx.x; idx < (b - a) * HD; idx += ATT_THREADS) {
                const int j = a + idx / HD, dd = idx - (idx / HD) * HD;
                float osum = 0.f;
#pragma unroll
                for (int w = 0; w < ATT_CWARPS; ++w) osum += red[(w * G + j) * HD + dd];
                float lj = l_run[0];
#pragma unroll
                for (int jj = 1; jj < G; ++jj) lj = (jj == j) ? l_run[jj] : lj;
                const float o = osum / lj;
                __nv_bfloat16 hi, lo;
                split_bf16(o, hi, lo);
                const int r = g * G + j;
                act[static_cast<size_t>(r) * ld_act + ocol + dd] = hi;
                act[static_cast<size_t>(r + bpad) * ld_act + ocol + dd] = lo;
            }
            named_bar_sync(1, ATT_THREADS);                             // red[] / scores are reused by the next segment
            a = b;
        }
    }
}

}  // namespace vcb
