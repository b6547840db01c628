// Device kernels of the codec-LM decode path (everything except the tcgen05 GEMM).
// All of them are HBM/L2-bound integer or fp32 work: coalesced 16-byte accesses, warp-level reductions,
// TMA bulk copies for the paged KV cache.  Included only by lm_engine.cu.
#pragma once
#include "vcb_internal.h"

namespace vcb {

static constexpr int KV_PAGE = 64;          // tokens per KV page
static constexpr int ATT_CWARPS = 8;            // consumer warps per CTA
static constexpr int ATT_THREADS = ATT_CWARPS * 32;
static constexpr int ATT_STAGES = 3;

__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

// ---------------------------------------------------------------------------------------------------
// Prompt embedding:  text rows  x = E_text[id] + alpha_t * PE[i]          (voicecraft.py:950-951)
//                    audio rows x = sum_k E_k[tok_k] (or mask_embedding[r]) + alpha_a * PE[j]   (:978-985, 311-320)
// One CTA per row of the prefill chunk.  fp32 throughout; alpha*pe is rounded before the add, like eager torch.
// ---------------------------------------------------------------------------------------------------
struct EmbedSeq {
    const long long* text_ids;
    const long long* y_tokens;   // [y_len][K]
    const int* mask_rows;        // [y_len] or null
    int x_len, y_len;
};

__global__ void embed_rows_kernel(const EmbedSeq* __restrict__ seqs, const int* __restrict__ row_seq,
                                  const int* __restrict__ row_pos, float* __restrict__ x_rows, int d, int K,
                                  const float* __restrict__ E_text, const float* const* __restrict__ E_audio,
                                  const float* __restrict__ mask_emb, const float* __restrict__ pe, float alpha_t,
                                  float alpha_a) {
    const int r = blockIdx.x;
    const EmbedSeq s = seqs[row_seq[r]];
    const int pos = row_pos[r];
    float* out = x_rows + static_cast<size_t>(r) * d;
    if (pos < s.x_len) {
        const float* e = E_text + static_cast<size_t>(s.text_ids[pos]) * d;
        const float* p = pe + static_cast<size_t>(pos) * d;
        for (int c = threadIdx.x; c < d; c += blockDim.x) out[c] = __fadd_rn(e[c], __fmul_rn(alpha_t, p[c]));
    } else {
        const int j = pos - s.x_len;
        const float* p = pe + static_cast<size_t>(j) * d;
        const int mr = s.mask_rows ? s.mask_rows[j] : -1;
        for (int c = threadIdx.x; c < d; c += blockDim.x) {
            float acc;
            if (mr >= 0) {
                acc = mask_emb[static_cast<size_t>(mr) * d + c];
            } else {
                acc = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float v = E_audio[k][static_cast<size_t>(s.y_tokens[static_cast<size_t>(j) * K + k]) * d + c];
                    acc = (k == 0) ? v : __fadd_rn(acc, v);
                }
            }
            out[c] = __fadd_rn(acc, __fmul_rn(alpha_a, p[c]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm of the fp32 residual stream -> bf16 hi / lo rows of the next GEMM's B operand.
// Replaces F.layer_norm (transformer.py:62-75); the residual adds live in the GEMM epilogue (EPI_RESID).
// One CTA per row; two-pass variance in registers.
// ---------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256)
ln_rows_kernel(const float* __restrict__ x_in, const int* __restrict__ src_index, const float* __restrict__ gamma,
               const float* __restrict__ beta, __nv_bfloat16* __restrict__ act, int ld_act, int bpad, int d, float eps) {
    __shared__ float red[8];
    pdl_launch_dependents();          // let the consumer GEMM start prefetching its weights right away
    if (threadIdx.x == 0) tl_mark(0x200);
    pdl_wait();
    if (threadIdx.x == 0) tl_mark(0x210);
    const int r = blockIdx.x;
    const int src = src_index ? src_index[r] : r;
    float v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int c = threadIdx.x + j * 256;
        v[j] = (c < d) ? x_in[static_cast<size_t>(src) * d + c] : 0.f;
        s += v[j];
    }
    const float mean = block_sum_256(s, red) / d;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int c = threadIdx.x + j * 256;
        if (c < d) {
            const float dv = v[j] - mean;
            q += dv * dv;
        }
    }
    const float var = block_sum_256(q, red) / d;
    const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
        const int c = threadIdx.x + j * 256;
        if (c < d) {
            const float y = (v[j] - mean) * rstd * gamma[c] + beta[c];
            __nv_bfloat16 hi, lo;
            split_bf16(y, hi, lo);
            act[static_cast<size_t>(r) * ld_act + c] = hi;
            act[static_cast<size_t>(r + bpad) * ld_act + c] = lo;
        }
    }
    if (threadIdx.x == 0) tl_mark(0x230);
}

// hidden state of the last token of each utterance: out[out_index[row]] = x[row]   (prefill -> first sampling step)
__global__ void gather_rows_kernel(const float* __restrict__ x_in, float* __restrict__ out,
                                   const int* __restrict__ out_index, int d) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x;
    const int dst = out_index[r];
    if (dst < 0) return;
    for (int c = threadIdx.x; c < d; c += blockDim.x)
        out[static_cast<size_t>(dst) * d + c] = x_in[static_cast<size_t>(r) * d + c];
}

// fp32 rows -> bf16 hi/lo rows (bring-up hook vcb_debug_gemm only)
__global__ void split_rows_kernel(const float* __restrict__ x, int N, __nv_bfloat16* __restrict__ act, int ld_act,
                                  int bpad) {
    const int r = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    __nv_bfloat16 hi, lo;
    split_bf16(x[static_cast<size_t>(r) * N + c], hi, lo);
    act[static_cast<size_t>(r) * ld_act + c] = hi;
    act[static_cast<size_t>(r + bpad) * ld_act + c] = lo;
}

// ---------------------------------------------------------------------------------------------------
// Attention over the paged KV cache, one CTA per (row, head).  q-length 1 per row (decode step, or one
// token of a prefill chunk): keys 0..pos of the row's own utterance -- the causal mask over [text;audio]
// of voicecraft.py:419-447 without ever materialising it.  K/V pages are staged with TMA bulk copies
// (cp.async.bulk -> UBLKCP) into a 2-stage shared-memory ring guarded by mbarriers; math is fp32 on CUDA
// cores (1 FLOP/byte: HBM-bound).  Replaces F.scaled_dot_product_attention at activation.py:634.
//   QK : LPT = hd/8 lanes per key, each lane one 16-byte chunk of the key row (conflict-free), xor-shuffle reduce
//   PV : warp w owns keys [16w,16w+16) of the page, lane owns hd/32 output dims
// Output: bf16 hi/lo rows for the out-projection GEMM.
// ---------------------------------------------------------------------------------------------------
template <typename KVT, int HD>
struct AttSmem {
    static constexpr int PAGE_BYTES = KV_PAGE * HD * sizeof(KVT);
    static constexpr int OFF_V = ATT_STAGES * PAGE_BYTES;
    static constexpr int OFF_SC = 2 * ATT_STAGES * PAGE_BYTES;
    static constexpr int OFF_PW = OFF_SC + 2 * KV_PAGE * 4;          // scores double-buffered by page parity
    static constexpr int OFF_RED = OFF_PW + ATT_CWARPS * KV_PAGE * 4;
    static constexpr int OFF_BAR = OFF_RED + ATT_CWARPS * HD * 4;
    static constexpr int TOTAL = OFF_BAR + 2 * ATT_STAGES * 8 + 128;
};

template <typename KVT, int N>
__device__ __forceinline__ void load_kv_vec(const KVT* p, float (&out)[N]) {
    if constexpr (sizeof(KVT) == 2) {
        if constexpr (N == 8) {
            const uint4 u = *reinterpret_cast<const uint4*>(p);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                out[2 * i] = __uint_as_float(w[i] << 16);
                out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            }
        } else if constexpr (N == 4) {
            const uint2 u = *reinterpret_cast<const uint2*>(p);
            out[0] = __uint_as_float(u.x << 16);
            out[1] = __uint_as_float(u.x & 0xffff0000u);
            out[2] = __uint_as_float(u.y << 16);
            out[3] = __uint_as_float(u.y & 0xffff0000u);
        } else {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
            out[0] = __uint_as_float(u << 16);
            out[1] = __uint_as_float(u & 0xffff0000u);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            const float2 f = *reinterpret_cast<const float2*>(p + i);
            out[i] = f.x;
            out[i + 1] = f.y;
        }
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Persistent, warp-specialised version: grid = a few CTAs per SM; each CTA walks a static list of work items
// (row*head, context chunk).  Warp 4 is the TMA producer: it runs ahead ACROSS items, so the HBM stream never drains
// at an item boundary (short CTAs with a cold start were the measured loss: 45% DRAM utilisation in ncu).
// Warps 0..ATT_CWARPS-1 consume pages: scores -> online softmax -> PV, release the stage through an mbarrier.
template <typename KVT, int HD>
__global__ void __launch_bounds__(ATT_THREADS + 32)
attn_rows_kernel(const float* __restrict__ qbuf, const KVT* __restrict__ kpool, const KVT* __restrict__ vpool,
                 const int* __restrict__ page_table, int max_pages, const int* __restrict__ row_slot,
                 const int* __restrict__ row_pos, int H, __nv_bfloat16* __restrict__ act, int ld_act, int bpad,
                 float scale, float* __restrict__ ws, int* __restrict__ cnt, int maxch, int chunk_pages, int n_rh,
                 int n_chunks, const int* __restrict__ row_pages) {
    using L = AttSmem<KVT, HD>;
    constexpr int LPT = HD / 8;          // lanes per key in QK
    constexpr int TPW = 32 / LPT;        // keys per warp iteration
    constexpr int DPT = HD / 32;         // output dims per lane in PV
    extern __shared__ __align__(128) uint8_t att_smem[];
    KVT* sK = reinterpret_cast<KVT*>(att_smem);
    KVT* sV = reinterpret_cast<KVT*>(att_smem + L::OFF_V);
    float* sc_all = reinterpret_cast<float*>(att_smem + L::OFF_SC);
    float* pw = reinterpret_cast<float*>(att_smem + L::OFF_PW);
    float* red = reinterpret_cast<float*>(att_smem + L::OFF_RED);
    uint64_t* full = reinterpret_cast<uint64_t*>(att_smem + L::OFF_BAR);
    uint64_t* empty = full + ATT_STAGES;
    __shared__ int s_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < ATT_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ATT_CWARPS);
        }
        mbar_fence_init();
        tl_mark(0x300);
        tl_mark_all(0x300);
    }
    __syncthreads();
    pdl_wait();
    if (threadIdx.x == 0) { tl_mark(0x310); tl_mark_all(0x310); }
    const int n_items = n_rh * n_chunks;

    if (warp == ATT_CWARPS) {
        // ===== producer: one lane streams the K/V pages of every item of this CTA, in order ================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();          // KV pages stream through L2 once per step
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int chunk = item / n_rh, rh = item - chunk * n_rh;
                const int r = rh / H, h = rh - r * H;
                const int pos = row_pos[r];
                if (pos < 0) continue;
                const int npages = pos / KV_PAGE + 1;
                const int p0 = chunk * chunk_pages;
                if (p0 >= npages) continue;
                const int p1 = min(npages, p0 + chunk_pages);
                // decode steps: step_prep left a per-row copy of the slot's page list, so the first TMA issue is one
                // L2 round trip away (row -> pages) instead of two (row -> slot -> pages)
                const int* pt = row_pages ? row_pages + r * max_pages : page_table + row_slot[r] * max_pages;
                for (int p = p0; p < p1; ++p, ++it) {
                    const int s = it % ATT_STAGES;
                    if (it >= ATT_STAGES) mbar_wait(&empty[s], ((it / ATT_STAGES) - 1) & 1);
                    const size_t off = (static_cast<size_t>(pt[p]) * H + h) * KV_PAGE * HD;
                    mbar_arrive_expect_tx(&full[s], 2 * L::PAGE_BYTES);
                    tma_bulk_g2s_hint(sK + s * KV_PAGE * HD, kpool + off, L::PAGE_BYTES, &full[s], pol);
                    tma_bulk_g2s_hint(sV + s * KV_PAGE * HD, vpool + off, L::PAGE_BYTES, &full[s], pol);
                }
            }
        }
        return;
    }

    // ===== consumers ================================================================================
    const int sub = lane % LPT;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int chunk = item / n_rh, rh = item - chunk * n_rh;
        const int r = rh / H, h = rh - r * H;
        const int pos = row_pos[r];
        if (pos < 0) continue;
        const int npages = pos / KV_PAGE + 1;
        const int p0 = chunk * chunk_pages;
        if (p0 >= npages) continue;
        const int p1 = min(npages, p0 + chunk_pages);
        const int nch = (npages + chunk_pages - 1) / chunk_pages;
        float q[8];
        {
            const float* qp = qbuf + (static_cast<size_t>(r) * H + h) * HD + sub * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = qp[i];
        }
        float m_run = -INFINITY, l_run = 0.f;
        float acc[DPT];
#pragma unroll
        for (int i = 0; i < DPT; ++i) acc[i] = 0.f;

        for (int p = p0; p < p1; ++p, ++it) {
            const int s = it % ATT_STAGES;
            mbar_wait(&full[s], (it / ATT_STAGES) & 1);
            const KVT* K = sK + s * KV_PAGE * HD;
            const KVT* V = sV + s * KV_PAGE * HD;
            float* sc = sc_all + (it & 1) * KV_PAGE;
            // ---- scores for this warp's KPW keys
            constexpr int KPW = KV_PAGE / ATT_CWARPS;
#pragma unroll
            for (int itq = 0; itq < KPW / TPW; ++itq) {
                const int t = warp * KPW + itq * TPW + lane / LPT;
                float kv[8];
                load_kv_vec<KVT, 8>(K + t * HD + sub * 8, kv);
                float dsum = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) dsum = fmaf(q[i], kv[i], dsum);
#pragma unroll
                for (int o = LPT / 2; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
                if (sub == 0) sc[t] = (p * KV_PAGE + t <= pos) ? dsum * scale : -INFINITY;
            }
            named_bar_sync(1, ATT_THREADS);
            // ---- online softmax bookkeeping (every warp redundantly over all 64 scores: identical m, l)
            const float s0 = sc[lane], s1 = sc[lane + 32];
            const float m_new = fmaxf(m_run, warp_max(fmaxf(s0, s1)));
            const float corr = expf(m_run - m_new);
            const float e0 = expf(s0 - m_new), e1 = expf(s1 - m_new);
            float* mypw = pw + warp * KV_PAGE;
            mypw[lane] = e0;
            mypw[lane + 32] = e1;
            l_run = l_run * corr + warp_sum(e0 + e1);
            m_run = m_new;
            __syncwarp();
            // ---- PV for this warp's KPW keys
#pragma unroll
            for (int i = 0; i < DPT; ++i) acc[i] *= corr;
#pragma unroll
            for (int tt = 0; tt < KPW; ++tt) {
                const int t = warp * KPW + tt;
                const float pt_ = mypw[t];
                float vv[DPT];
                load_kv_vec<KVT, DPT>(V + t * HD + lane * DPT, vv);
#pragma unroll
                for (int i = 0; i < DPT; ++i) acc[i] = fmaf(pt_, vv[i], acc[i]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);            // this warp is done with stage s
        }
        // ---- combine the 4 warps' partial outputs of this item
#pragma unroll
        for (int i = 0; i < DPT; ++i) red[warp * HD + lane * DPT + i] = acc[i];
        named_bar_sync(1, ATT_THREADS);
        const size_t ocol = static_cast<size_t>(h) * HD;
        if (nch == 1) {
            for (int dd = threadIdx.x; dd < HD; dd += ATT_THREADS) {
                float osum = 0.f;
#pragma unroll
                for (int w = 0; w < ATT_CWARPS; ++w) osum += red[w * HD + dd];
                const float o = osum / l_run;
                __nv_bfloat16 hi, lo;
                split_bf16(o, hi, lo);
                act[static_cast<size_t>(r) * ld_act + ocol + dd] = hi;
                act[static_cast<size_t>(r + bpad) * ld_act + ocol + dd] = lo;
            }
            named_bar_sync(1, ATT_THREADS);                   // red[] is reused by the next item
            continue;
        }
        // ---- split context: publish (o, m, l); the last chunk to finish merges all chunks in chunk order
        float* myws = ws + (static_cast<size_t>(rh) * maxch + chunk) * (HD + 2);
        for (int dd = threadIdx.x; dd < HD; dd += ATT_THREADS) {
            float osum = 0.f;
#pragma unroll
            for (int w = 0; w < ATT_CWARPS; ++w) osum += red[w * HD + dd];
            myws[dd] = osum;
        }
        if (threadIdx.x == 0) {
            myws[HD] = m_run;
            myws[HD + 1] = l_run;
        }
        __threadfence();
        named_bar_sync(1, ATT_THREADS);
        if (threadIdx.x == 0) s_last = (atomicAdd(&cnt[rh], 1) == nch - 1);
        named_bar_sync(1, ATT_THREADS);
        if (s_last) {
            __threadfence();
            const volatile float* base = ws + static_cast<size_t>(rh) * maxch * (HD + 2);
            float M = -INFINITY;
            for (int c = 0; c < nch; ++c) M = fmaxf(M, base[c * (HD + 2) + HD]);
            for (int dd = threadIdx.x; dd < HD; dd += ATT_THREADS) {
                float Lsum = 0.f, O = 0.f;
                for (int c = 0; c < nch; ++c) {
                    const float w = expf(base[c * (HD + 2) + HD] - M);
                    Lsum += base[c * (HD + 2) + HD + 1] * w;
                    O += base[c * (HD + 2) + dd] * w;
                }
                const float o = O / Lsum;
                __nv_bfloat16 hi, lo;
                split_bf16(o, hi, lo);
                act[static_cast<size_t>(r) * ld_act + ocol + dd] = hi;
                act[static_cast<size_t>(r + bpad) * ld_act + ocol + dd] = lo;
            }
            if (threadIdx.x == 0) cnt[rh] = 0;
        }
        named_bar_sync(1, ATT_THREADS);                       // s_last / red[] reused by the next item
    }
    if (threadIdx.x == 0) { tl_mark(0x330); tl_mark_all(0x330); }
}

// ---------------------------------------------------------------------------------------------------
// Step prologue: assign each row its position, advance the per-slot counters, fetch the input embedding
// produced by the previous sampler call.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
step_prep_kernel(const int* __restrict__ slots, int n, SlotState* __restrict__ st, const GroupState* __restrict__ gr,
                 int* __restrict__ row_slot, int* __restrict__ row_pos, int* __restrict__ row_last,
                 const float* __restrict__ x_slot, float* __restrict__ x_rows, int d,
                 const float* __restrict__ gamma0, __nv_bfloat16* __restrict__ act, int bpad, float* __restrict__ stats,
                 const int* __restrict__ page_table, int max_pages, int* __restrict__ row_page,
                 int* __restrict__ row_pages, int* __restrict__ row_forced, unsigned int* __restrict__ phase_flags,
                 int n_phase_flags, unsigned int* __restrict__ tile_counters, int n_tile_counters,
                 __nv_bfloat16* __restrict__ act_tiled) {
    __shared__ float red[8];
    pdl_launch_dependents();
    pdl_wait();
    const int r = blockIdx.x;
    if (r == 0)                                  // completion counters of the persistent step kernel (mega_step.cu)
        for (int i = threadIdx.x; i < n_phase_flags; i += blockDim.x) phase_flags[i] = 0u;
    // ... and its split-K arrival counters, spread over the CTAs of this launch
    for (int i = r * blockDim.x + threadIdx.x; i < n_tile_counters; i += gridDim.x * blockDim.x) tile_counters[i] = 0u;
    const int slot = slots[r];
    __shared__ int s_pos;
    if (threadIdx.x == 0) {
        SlotState& S = st[slot];
        const bool on = S.active && !gr[S.group].done;
        s_pos = on ? S.seq_len : -1;
        row_slot[r] = slot;
        row_pos[r] = s_pos;
        row_last[r] = on ? slot : -1;
        row_page[r] = on ? page_table[slot * max_pages + s_pos / KV_PAGE] : 0;
        row_forced[r] = S.forced;           // snapshot for the sampler's K CTAs of this slot (they are not ordered)
        if (on) {
            S.seq_len += 1;
            S.y_len += 1;
        }
    }
    for (int j = threadIdx.x; j < max_pages; j += blockDim.x)       // the attention producer reads these by row
        row_pages[static_cast<size_t>(r) * max_pages + j] = page_table[slot * max_pages + j];
    __syncthreads();
    if (s_pos < 0) return;
    // x row + (LayerNorm folding) gamma0 * x as hi/lo rows and the row statistics for layer 0's QKV GEMM
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float v = x_slot[static_cast<size_t>(slot) * d + c];
        x_rows[static_cast<size_t>(r) * d + c] = v;
        if (gamma0) {
            __nv_bfloat16 hi, lo;
            split_bf16(gamma0[c] * v, hi, lo);
            act[static_cast<size_t>(r) * d + c] = hi;
            act[static_cast<size_t>(r + bpad) * d + c] = lo;
            if (act_tiled) {       // persistent step kernel: the same operand as the pre-swizzled image of its UMMA tiles (mega_step.cu)
                const int kk = c & 63, rows2 = 2 * bpad;
                const size_t t0 = static_cast<size_t>(c >> 6) * rows2;
                act_tiled[(t0 + r) * 64 + ((((kk >> 3) ^ (r & 7)) << 3) | (kk & 7))] = hi;
                act_tiled[(t0 + r + bpad) * 64 + ((((kk >> 3) ^ ((r + bpad) & 7)) << 3) | (kk & 7))] = lo;
            }
            s1 += v;
            s2 += v * v;
        }
    }
    if (gamma0) {
        s1 = block_sum_256(s1, red);
        s2 = block_sum_256(s2, red);
        if (threadIdx.x == 0) {
            stats[static_cast<size_t>(r) * 2] = s1;
            stats[static_cast<size_t>(r) * 2 + 1] = s2;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused sampler: one CTA per (utterance, codebook) row of V logits.  Reproduces, without a host sync,
//   * the in-place logit edits of sample_helper (voicecraft.py:1018-1067 tts, :718-787 edit, :1269-1325 batch)
//   * temperature, top-k (strict '<' vs the k-th largest value), top-p (sorted cumulative softmax, shift by one),
//     softmax and torch.multinomial(.,1) == argmax(p / q) with caller-provided q ~ Exp(1)   (:26-86)
//   * empty-token forcing, end-token trigger (sample / argmax / length cap), silence bookkeeping, the K-1 step
//     end cascade, best-of-N `keep`, multi-span hand-over; and it emits the next input embedding
//     sum_k E_k[tok_k] + alpha * PE[t]                                                   (:1102-1116)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2key(float x) {        // order-preserving float -> uint
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---------------------------------------------------------------------------------------------------
// Exp(1) noise of `torch.multinomial` generated in place (voicecraft.py:85: multinomial(p, 1) == argmax(p / q),
// q = empty_like(p).exponential_(1)).  Bit-identical to ATen's CUDA path for a draw of `numel` fp32 elements from a
// Philox generator at (seed, offset): distribution_nullary_kernel launches T = 256 * grid threads, thread `idx` runs
// curand_init(seed, idx, offset) and element li of loop iteration `it` takes component (li % 4T) / T of the it-th
// curand_uniform4 of thread (li % T); exponential_ maps u -> -log(u) with the u ~ 1 guard of
// ATen/core/TransformationHelper.h.  (offset is a multiple of 4 in torch: one 128-bit Philox counter per call.)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float torch_exponential_at(unsigned long long seed, unsigned long long offset,
                                                      unsigned int threads, unsigned long long li) {
    const unsigned long long per_iter = 4ull * threads;
    const unsigned long long it = li / per_iter, rem = li - it * per_iter;
    const unsigned int comp = static_cast<unsigned int>(rem / threads);
    const unsigned long long idx = rem - static_cast<unsigned long long>(comp) * threads;
    const unsigned long long ctr = offset / 4ull + it;          // curand_init skipahead(offset) + one counter per curand4
    const uint4 o = philox4x32_10(make_uint4(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32),
                                             static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32)),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    const uint32_t x = comp == 0 ? o.x : comp == 1 ? o.y : comp == 2 ? o.z : o.w;
    // curand_uniform: x * 2^-32 + 2^-33 (the product is exact, one rounding in the add), range (0, 1]
    const float u = __fadd_rn(__fmul_rn(static_cast<float>(x), 2.3283064365386963e-10f), 1.16415321826934814453e-10f);
    // at::log is __logf in device code (ATen/NumericUtils.h), which is why the transform guards u ~ 1
    const float lg = (u >= 1.0f - 1.1920928955078125e-07f / 2) ? -1.1920928955078125e-07f / 2 : __logf(u);
    return -lg;                                                 // (-1 / lambda) * log with lambda = 1
}
// philox offset consumed by one draw of numel elements (ATen calc_execution_policy: counter_offset)
__host__ __device__ inline unsigned long long torch_draw_offset(unsigned long long numel, unsigned int threads) {
    return ((numel - 1) / (4ull * threads) + 1) * 4ull;
}

__global__ void debug_exponential_kernel(float* out, unsigned long long numel, unsigned long long seed,
                                         unsigned long long offset, unsigned int threads) {
    for (unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; i < numel;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
        out[i] = torch_exponential_at(seed, offset, threads, i);
}

struct SamplerArgs {
    const int* slots;
    const int* row_forced;    // per listed slot: SlotState::forced as of the step prologue (null: read the slot)
    int n;
    SlotState* st;
    GroupState* gr;
    const float* logits;      // [n][ldl] fp32 (bias included), column = k*Vpad + v
    int ldl;
    const float* noise;       // [n*K][V], or null: generated from the group's Philox stream (GroupState::rng_*)
    float* dbg_logits;        // [n*K][V] or null
    int* tok_log;             // [max_slots][max_steps][K]
    int max_steps, max_seq;
    float* x_slot;            // [max_slots][d]
    const float* const* E_audio;
    const float* mask_emb;
    const float* pe;
    float alpha_a;
    int d, K, V, Vpad;
    int empty_token, eog, eos, encodec_sr;
    SamplingParams sp;
};

static constexpr int SAMP_THREADS = 256;
static constexpr int SAMP_MAXV = 12;       // V <= 3072
static constexpr int SAMP_SORT_N = 4096;

__device__ void sampler_finish_slot(const SamplerArgs& a, int slot, float* sred);

__global__ void __launch_bounds__(SAMP_THREADS) sampler_kernel(const SamplerArgs a) {
    __shared__ float sred[8];
    __shared__ int sidx[8];
    __shared__ int hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int s_kleft;
    __shared__ int s_flag;
    extern __shared__ unsigned long long sort_buf[];     // SAMP_SORT_N entries (only used when top_p < 1)

    pdl_launch_dependents();
    if (threadIdx.x == 0) tl_mark(0x400);
    pdl_wait();
    if (threadIdx.x == 0) tl_mark(0x410);
    const int i = blockIdx.x / a.K, k = blockIdx.x % a.K;
    const int slot = a.slots[i];
    SlotState& S = a.st[slot];
    if (!S.active) return;
    GroupState& G = a.gr[S.group];
    if (G.done) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.K, V = a.V;

    // `forced` as of the start of this step: the k == 0 CTA decrements S.forced below, and the K CTAs of a slot are not
    // ordered against each other (they need not even be co-resident), so the decision is taken on a snapshot
    const int forced_now = a.row_forced ? a.row_forced[i] : S.forced;
    if (forced_now > 0) {
        // edit-mode hand-over to the next span (voicecraft.py:838-858): this forward fed a forced embedding,
        // nothing is sampled; prepare the next forced input.
        if (k != 0) return;
        const int f = forced_now;               // 2: next input = mask embedding, 1: next input = empty-token embedding
        const float* pe = a.pe + static_cast<size_t>(S.y_len) * a.d;
        for (int c = tid; c < a.d; c += SAMP_THREADS) {
            float acc;
            if (f == 2) {
                acc = a.mask_emb[static_cast<size_t>(G.more_mask[0]) * a.d + c];
            } else {
                acc = 0.f;
                for (int kk = 0; kk < K; ++kk) {
                    const float v = a.E_audio[kk][static_cast<size_t>(a.empty_token) * a.d + c];
                    acc = kk == 0 ? v : __fadd_rn(acc, v);
                }
            }
            a.x_slot[static_cast<size_t>(slot) * a.d + c] = __fadd_rn(acc, __fmul_rn(a.alpha_a, pe[c]));
        }
        __syncthreads();
        if (tid == 0) {
            if (f == 2) {                       // consume the mask row
                for (int j = 0; j < 7; ++j) G.more_mask[j] = G.more_mask[j + 1];
            }
            S.forced = f - 1;
        }
        return;
    }

    const bool tts = G.mode == 0;
    const int E = tts ? (a.eos > 0 ? a.eos : a.eog) : a.eog;
    const int n_eog = G.n_eog, cur = G.cur_num_gen;
    const int row = i * K + k;

    // ---- load logits (split-K reduce + bias), apply the reference's in-place edits -----------------
    float l[SAMP_MAXV];
#pragma unroll
    for (int j = 0; j < SAMP_MAXV; ++j) {
        const int v = tid + j * SAMP_THREADS;
        float u = -INFINITY;
        if (v < V) {
            u = a.logits[static_cast<size_t>(i) * a.ldl + k * a.Vpad + v];
            if (a.dbg_logits) a.dbg_logits[static_cast<size_t>(row) * V + v] = u;
            if (a.eos > 0 && v == (tts ? a.eog : a.eos)) u = -10000.f;                   // :1091-1093 / :816-818
            if (n_eog == 0) {
                if (k >= 1 && (v == E || v == a.empty_token)) u = -10000.f;                // :1021-1023
                if (k == 0 && tts && cur <= a.encodec_sr / 5 && v == E) u = -10000.f;     // :1024-1025
                if (k == 0 && a.sp.stop_repetition > 0 && v == S.prev_token && S.consec > a.sp.stop_repetition) {
                    bool sil = false;
                    for (int t = 0; t < a.sp.n_silence; ++t) sil |= (a.sp.silence_tokens[t] == v);
                    if (sil) {                                                             // :1027-1031
                        const float f = static_cast<float>(S.consec - (a.sp.stop_repetition - 1));
                        u = (u < 0.f) ? u * f : u / f;
                    }
                }
            } else {
                if (k > n_eog && (v == E || v == a.empty_token)) u = -10000.f;            // :1056-1058
            }
        }
        l[j] = u;
    }

    // the Exp(1) draws are independent of everything below: fetch them now, not after the softmax
    float nz[SAMP_MAXV];
    if (a.noise) {
#pragma unroll
        for (int j = 0; j < SAMP_MAXV; ++j) {
            const int v = tid + j * SAMP_THREADS;
            nz[j] = (v < V) ? a.noise[static_cast<size_t>(row) * V + v] : 1.f;
        }
    } else {
        // this group's own generator: the draw has the reference's shape [size*K, V], row = member*K + k
        const unsigned long long seed = (static_cast<unsigned long long>(G.seed_hi) << 32) | G.seed_lo;
        const unsigned long long off = (static_cast<unsigned long long>(G.off_hi) << 32) | G.off_lo;
        const unsigned long long base = (static_cast<unsigned long long>(S.member) * K + k) * V;
#pragma unroll
        for (int j = 0; j < SAMP_MAXV; ++j) {
            const int v = tid + j * SAMP_THREADS;
            nz[j] = (v < V) ? torch_exponential_at(seed, off, G.rng_threads, base + v) : 1.f;
        }
    }

    // ---- argmax of the edited logits (first index wins), needed for the end-token trigger ----------
    float bm = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < SAMP_MAXV; ++j) {
        const int v = tid + j * SAMP_THREADS;
        if (v < V && (l[j] > bm || (l[j] == bm && v < bi))) { bm = l[j]; bi = v; }
    }
    auto block_argmax = [&](float& val, int& idx) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, val, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
        }
        __syncthreads();
        if (lane == 0) { sred[warp] = val; sidx[warp] = idx; }
        __syncthreads();
        val = sred[0]; idx = sidx[0];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            if (sred[w] > val || (sred[w] == val && sidx[w] < idx)) { val = sred[w]; idx = sidx[w]; }
    };
    block_argmax(bm, bi);
    const int argmax_raw = bi;

    // ---- temperature (:80-81) -----------------------------------------------------------------------
    if (a.sp.temperature != 1.0f) {
#pragma unroll
        for (int j = 0; j < SAMP_MAXV; ++j) l[j] = __fdiv_rn(l[j], a.sp.temperature);
        bm = __fdiv_rn(bm, a.sp.temperature);
    }

    // ---- top-k: exact k-th largest by 4-pass radix select on order-preserving keys (:38-44) ----------
    if (a.sp.top_k > 0) {
        const int kk = min(max(a.sp.top_k, 1), V);
        if (tid == 0) { s_prefix = 0; s_kleft = kk; }
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            hist[tid] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
#pragma unroll
            for (int j = 0; j < SAMP_MAXV; ++j) {
                const int v = tid + j * SAMP_THREADS;
                // warp-aggregated histogram update: logits share a handful of exponent bytes, so plain atomics would
                // serialise ~32-way on the same shared-memory word in the first passes
                const uint32_t key = v < V ? f2key(l[j]) : 0u;
                const bool in = v < V && (key & pmask) == prefix;
                const unsigned act = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const int bin = (key >> shift) & 0xff;
                    const unsigned peers = __match_any_sync(act, bin);
                    if (lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
                }
            }
            __syncthreads();
            if (warp == 0) {
                // descending scan over the 256 bins: lane l owns bins [255-8l-7, 255-8l]
                int loc[8], lsum = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    loc[u] = hist[255 - 8 * lane - u];
                    lsum += loc[u];
                }
                int incl = lsum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                const int left0 = s_kleft;
                const int before = incl - lsum;                     // keys in higher bins than this lane's
                const unsigned hit = __ballot_sync(0xffffffffu, incl >= left0);
                const int owner = hit ? __ffs(hit) - 1 : 31;
                if (lane == owner) {
                    int left = left0 - before, b = 255 - 8 * lane;
                    for (int u = 0; u < 7; ++u) {
                        if (loc[u] >= left) break;
                        left -= loc[u];
                        --b;
                    }
                    s_kleft = left;
                    s_prefix = prefix | (static_cast<uint32_t>(b) << shift);
                }
            }
            __syncthreads();
        }
        const float thr = key2f(s_prefix);
#pragma unroll
        for (int j = 0; j < SAMP_MAXV; ++j)
            if (l[j] < thr) l[j] = -INFINITY;
        __syncthreads();
    }

    // ---- top-p (:46-67): sort descending, softmax, cumulative sum, keep ranks < j0 ---------------------
    if (a.sp.top_p < 1.0f) {
        for (int s = tid; s < SAMP_SORT_N; s += SAMP_THREADS) sort_buf[s] = 0ull;   // pads sort last (key 0 < any real key)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SAMP_MAXV; ++j) {
            const int v = tid + j * SAMP_THREADS;
            if (v < V)   // descending by value; ties -> lower index first
                sort_buf[v] = (static_cast<unsigned long long>(f2key(l[j])) << 32) | static_cast<uint32_t>(0xffffffffu - v);
        }
        __syncthreads();
        for (int size = 2; size <= SAMP_SORT_N; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < SAMP_SORT_N / 2; t += SAMP_THREADS) {
                    const int lo = 2 * t - (t & (stride - 1));
                    const int hi = lo + stride;
                    const bool desc = ((lo & size) == 0);
                    const unsigned long long x = sort_buf[lo], y = sort_buf[hi];
                    if ((x < y) == desc) { sort_buf[lo] = y; sort_buf[hi] = x; }
                }
                __syncthreads();
            }
        }
        // softmax over the sorted row (max = first element) and inclusive scan in rank order
        const float smax = key2f(static_cast<uint32_t>(sort_buf[0] >> 32));
        constexpr int PER = SAMP_SORT_N / SAMP_THREADS;      // 16 consecutive ranks per thread
        float e[PER];
        float loc = 0.f;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int rnk = tid * PER + j;
            const unsigned long long ent = sort_buf[rnk];
            e[j] = (rnk < V) ? expf(key2f(static_cast<uint32_t>(ent >> 32)) - smax) : 0.f;
            loc += e[j];
        }
        const float total = block_sum_256(loc, sred);
        // exclusive prefix of `loc` across threads
        float incl = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) sred[warp] = incl;
        __syncthreads();
        float base = 0.f;
        for (int w = 0; w < warp; ++w) base += sred[w];
        float run = base + incl - loc;
        int cnt = 0;                                       // ranks j with cum_j <= top_p
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            run += e[j];
            const int rnk = tid * PER + j;
            if (rnk < V && !(run / total > a.sp.top_p)) cnt++;
        }
        __syncthreads();
        const int j0 = static_cast<int>(block_sum_256(static_cast<float>(cnt), sred) + 0.5f) + 1;   // kept ranks: [0, j0)
        // cum is monotone, so "cum_j <= top_p" holds exactly for ranks [0, j0-1): rank of a value = its sorted position
        __syncthreads();
        // mark removed: write per-index flag through the sorted order
        // reuse hist as nothing; flags go into the low bit trick: store rank into a dense array (aliasing sort_buf upper half)
        int* rank_of = reinterpret_cast<int*>(sort_buf + SAMP_SORT_N);      // V ints after the sort area
        for (int rnk = tid; rnk < V; rnk += SAMP_THREADS) {
            const uint32_t idx = 0xffffffffu - static_cast<uint32_t>(sort_buf[rnk] & 0xffffffffull);
            rank_of[idx] = rnk;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SAMP_MAXV; ++j) {
            const int v = tid + j * SAMP_THREADS;
            if (v < V && rank_of[v] >= j0) l[j] = -INFINITY;
        }
        __syncthreads();
    }

    // ---- softmax + multinomial == argmax(p / q)  (:85) ---------------------------------------------------
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SAMP_MAXV; ++j) mx = fmaxf(mx, l[j]);
    mx = warp_max(mx);
    __syncthreads();
    if (lane == 0) sred[warp] = mx;
    __syncthreads();
    mx = sred[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, sred[w]);
    float ev[SAMP_MAXV];
    float esum = 0.f;
#pragma unroll
    for (int j = 0; j < SAMP_MAXV; ++j) {
        const int v = tid + j * SAMP_THREADS;
        ev[j] = (v < V) ? expf(l[j] - mx) : 0.f;
        esum += ev[j];
    }
    const float tot = block_sum_256(esum, sred);
    float best = -1.f;
    int besti = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < SAMP_MAXV; ++j) {
        const int v = tid + j * SAMP_THREADS;
        if (v < V) {
            const float p = ev[j] / tot;
            const float sc = p / nz[j];
            if (sc > best || (sc == best && v < besti)) { best = sc; besti = v; }
        }
    }
    block_argmax(best, besti);
    int tok = besti;

    // ---- forced values / end-token logic ----------------------------------------------------------------
    if (tid == 0) {
        if (n_eog == 0) {
            if (cur < K - 1 && k > cur) tok = a.empty_token;                                 // :1037-1039
            if (k == 0) {
                const int cap = tts ? S.x_len * (a.encodec_sr / 5) : S.x_len * 10;
                if (tok == E || argmax_raw == E || S.y_len > cap) {                          // :1041-1045
                    tok = E;
                    atomicMax(&G.trig_keep, S.member + 1);
                }
            }
        } else if (G.size == 1 || S.member == G.keep) {                                      // :1063-1066 / :1321-1323
            if (k < n_eog) tok = a.empty_token;
            else if (k == n_eog) tok = E;
        }
        a.tok_log[(static_cast<size_t>(slot) * a.max_steps + S.n_steps) * K + k] = tok;
        __threadfence();
        s_flag = (atomicAdd(&S.arrive, 1) == K - 1);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    sampler_finish_slot(a, slot, sred);
    if (threadIdx.x == 0) tl_mark(0x430);
}

// Last codebook row of a slot: next input embedding + silence bookkeeping; last slot of a group: group state.
__device__ void sampler_finish_slot(const SamplerArgs& a, int slot, float* sred) {
    SlotState& S = a.st[slot];
    GroupState& G = a.gr[S.group];
    const int tid = threadIdx.x, K = a.K;
    const volatile int* toks = a.tok_log + (static_cast<size_t>(slot) * a.max_steps + S.n_steps) * K;
    const float* pe = a.pe + static_cast<size_t>(S.y_len) * a.d;
    for (int c = tid; c < a.d; c += SAMP_THREADS) {
        float acc = 0.f;
        for (int kk = 0; kk < K; ++kk) {
            const float v = a.E_audio[kk][static_cast<size_t>(toks[kk]) * a.d + c];
            acc = kk == 0 ? v : __fadd_rn(acc, v);
        }
        a.x_slot[static_cast<size_t>(slot) * a.d + c] = __fadd_rn(acc, __fmul_rn(a.alpha_a, pe[c]));
    }
    __syncthreads();
    if (tid != 0) return;
    if (G.n_eog == 0) {                                                                     // :1047-1051
        const int t0 = toks[0];
        bool sil = false;
        for (int t = 0; t < a.sp.n_silence; ++t) sil |= (a.sp.silence_tokens[t] == t0);
        S.consec = (sil && t0 == S.prev_token) ? S.consec + 1 : 0;
        S.prev_token = t0;
    }
    S.n_steps += 1;
    S.arrive = 0;
    __threadfence();
    if (atomicAdd(&G.arrive, 1) != G.size - 1) return;
    __threadfence();
    // ---- group finalize
    G.arrive = 0;
    if (G.rng_threads) {                    // one draw of [size*K, V] consumed, whether or not the caller supplied noise
        unsigned long long off = (static_cast<unsigned long long>(G.off_hi) << 32) | G.off_lo;
        off += torch_draw_offset(static_cast<unsigned long long>(G.size) * K * a.V, G.rng_threads);
        G.off_lo = static_cast<unsigned int>(off);
        G.off_hi = static_cast<unsigned int>(off >> 32);
    }
    if (G.n_eog == 0) {
        if (G.trig_keep > 0) {
            G.n_eog = 1;
            G.keep = G.trig_keep - 1;            // the last member that triggered wins (:1302)
        }
    } else {
        G.n_eog += 1;
    }
    G.trig_keep = 0;
    G.cur_num_gen += 1;
    if (S.n_steps >= a.max_steps - 1 || S.seq_len >= a.max_seq - 2) {     // token log / KV capacity exhausted: stop, flag it
        G.done = 2;
        return;
    }
    if (G.n_eog == K) {                                                                      // span finished
        if (G.n_spans_done < 8) G.span_ends[G.n_spans_done] = S.n_steps;
        G.n_spans_done += 1;
        if (G.mode == 1 && G.spans_left > 0) {
            G.spans_left -= 1;
            G.n_eog = 0;
            G.cur_num_gen = 0;
            for (int mI = 0; mI < G.size; ++mI) {
                SlotState& M = a.st[G.first_slot + mI];
                M.forced = 2;
                M.prev_token = -1;
                M.consec = 0;
            }
        } else {
            G.done = 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Delayed codebook pattern gather (integer): out[b,k,s] = z[b,k,s-1-k] if 0 <= s-1-k < T else special.
// codebooks_patterns.py:151-176 with the DelayedPatternProvider layout (:336-352, delays = 0..K-1).
// ---------------------------------------------------------------------------------------------------
__global__ void delay_pattern_kernel(const long long* __restrict__ z, long long* __restrict__ out, int K, int T,
                                     long long special) {
    const int S = T + K;
    const size_t bk = blockIdx.y;
    const int k = static_cast<int>(bk % K);
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const int t = s - 1 - k;
        out[bk * S + s] = (t >= 0 && t < T) ? z[bk * T + t] : special;
    }
}

}  // namespace vcb
