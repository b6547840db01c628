// EnCodec SEANet decoder on the tensor cores (tcgen05), round 2.
//
// Replaces, for the configurations it covers, the CUDA-core kernels of encodec.cu behind the same entry point
// (enc_decode <- AudioTokenizer.decode, reference data/tokenizer.py:131-133 -> audiocraft EncodecModel.decode).
//
// Every layer of the decoder is one launch of ONE kernel, an implicit GEMM with the time steps as the UMMA M dimension:
//
//   activations  channels-last bf16 "planes": x ~= hi + lo (split_bf16), tensor [2 planes][rows][C], a row = one time step of
//                one utterance, every utterance preceded by `halo` rows that hold its left padding (reflect or zero), so a
//                causal convolution tap is the same 128-row TMA box shifted up by (k-1-j)*dilation rows -- no im2col, no
//                gather.  C is padded to a multiple of 64 (one 128-byte swizzle row) with zero channels.
//   weights      fp32 -> (hi, lo) bf16, pre-tiled [n-tile][k-block][hi|lo][BN][64]: one TMA box per k-block.
//   arithmetic   D += Ahi*Bhi + Alo*Bhi + Ahi*Blo in fp32 TMEM: the "3-pass" product, relative error ~2^-16 per term, i.e.
//                fp32-level parity with the CUDA-core path (tests/test_codec.py states the waveform tolerance).
//   ConvTranspose1d (stride r, kernel 2r, causal trim): y[t*r+p] = W[:,:,p] x[t] + W[:,:,p+r] x[t-1]: a 2-tap convolution
//                with N = r*Cout whose output row [r*Cout] IS r consecutive channels-last output rows.
//   residual block  conv k3 -> hidden; then conv k1 (hidden) and the 1x1 shortcut (block input) are ONE GEMM with
//                K = hidden + C (two A sources), bias = b2 + bs.
//   epilogue     TMEM -> registers: + bias, optional ELU, split to planes (raw and/or ELU'd form, as the consumers need),
//                mirror rows 1..pad into the halo (reflect) or zero it; or fp32 rows (LSTM pre-activations, waveform).
//   LSTM         W_ih for all steps is one GEMM; each step is one launch of the same kernel over h_{t-1} planes with the
//                gates interleaved (n = 4*unit + gate) so the epilogue owns complete cells: c/h update, h_t planes for the
//                next step, and for the last layer ELU(h + skip) straight into the first ConvTranspose's input.
//                800 steps x 2 layers = 1 600 dependent launches for the WHOLE batch (was: per 16 utterances), W_hh (16.8 MB
//                as planes) stays in L2; chained with programmatic dependent launch, weights in flight before the wait.
#include "codec_tc.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "vcb_internal.h"

namespace vcb {

static constexpr int TC_BM = 128;
static constexpr int TC_BK = 64;
static constexpr int TC_EPI_WARPS = 16;
static constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
static constexpr int TC_MAX_KB = 64;
enum { TC_MODE_CONV = 0, TC_MODE_LSTM = 1 };

struct TcTap {
    short src;     // which A tensor map (0 / 1)
    short shift;   // rows above the output row
    int coff;      // first channel of this k-block
};

struct TcCall {
    int total_kb, ntiles, mtiles, rows_total, row_base;
    int rcap[2];                       // rows per plane of A source 0 / 1 (lo plane = + rcap rows)
    int in_store_halo;                 // 1: halo rows (t >= -in_halo) are computed and stored too
    int in_div, in_tm, in_halo, T_in, B;   // A row -> (b, t): utterance-major (row / Tp, row % Tp - halo) or time-major
    int mode, up, Cout, Nstore;        // GEMM column n -> (phase p = n / Cout, channel n % Cout); columns >= Nstore are padding
    const float* bias;
    __nv_bfloat16* raw;                // optional output planes, raw / ELU'd form; row(b, t') = b*o_sb + t'*o_st + o_off
    __nv_bfloat16* elu;
    long long o_plane, o_sb, o_st, o_off;
    int o_ld, o_halo, o_halo_zero;     // halo rows t' = -1..-o_halo: mirror of rows 1..o_halo (reflect) or zeros
    float* f32;                        // optional fp32 rows
    long long f_sb, f_st, f_off;
    int f_ld, f_valid, f_scalar;
    // LSTM step
    const float* pre;                  // [T][Bcap][4H] gate pre-activations (interleaved), bias included
    float* cst;                        // [Bcap][H]
    const float* skip;                 // [T][Bcap][H] or null
    __nv_bfloat16* hseq;               // planes [T+1][Bcap][H]; slot t+1 = h_t
    long long h_plane;
    int t_step, Bcap, H;
    TcTap taps[TC_MAX_KB];
};

template <int BN, int STAGES>
struct TcSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 2;
    static constexpr int B_BYTES = BN * TC_BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16;
};

// ELU for the planes the next layer reads: exp through ex2.approx (absolute error ~2e-7, far below the 2^-17 relative
// error of the hi/lo split it feeds)
__device__ __forceinline__ float tc_elu(float x) { return x > 0.f ? x : __expf(x) - 1.f; }
__device__ __forceinline__ float tc_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// two fp32 -> packed bf16x2 hi parts and bf16x2 lo parts (same roundings as split_bf16)
__device__ __forceinline__ void tc_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t*>(&h);
    __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ void tc_split8(const float* v, uint4& hi, uint4& lo) {
    tc_split2(v[0], v[1], hi.x, lo.x);
    tc_split2(v[2], v[3], hi.y, lo.y);
    tc_split2(v[4], v[5], hi.z, lo.z);
    tc_split2(v[6], v[7], hi.w, lo.w);
}

// 32 bytes per lane in one instruction (STG.256, sm_100): every lane writes a whole 32-byte sector
__device__ __forceinline__ void tc_st256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// CW consecutive channels of one row -> both planes (and the mirrored halo row, if any).
// (Measured and rejected, same box A/B `gpurun_out/r2q`: a branch-free ex2.approx ELU + 16-column chunks everywhere + the bias
// fetched after the accumulator wait (L1-prefetched) made the narrow stages 20-50 % slower: 147 -> 156 ms at B = 256.)
template <int CW>
__device__ __forceinline__ void tc_store_planes(__nv_bfloat16* base, long long plane, int ld, long long row, long long mirror,
                                                int co0, const float (&v)[CW], bool elu, bool zero_mirror) {
    uint4 hi[CW / 8], lo[CW / 8];
#pragma unroll
    for (int j = 0; j < CW / 8; ++j) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = elu ? tc_elu(v[8 * j + u]) : v[8 * j + u];
        tc_split8(w, hi[j], lo[j]);
    }
    __nv_bfloat16* ph = base + row * ld + co0;
    __nv_bfloat16* pl = ph + plane;
#pragma unroll
    for (int j = 0; j < CW / 16; ++j) {
        tc_st256(ph + 16 * j, hi[2 * j], hi[2 * j + 1]);
        tc_st256(pl + 16 * j, lo[2 * j], lo[2 * j + 1]);
    }
    if (mirror >= 0) {
        __nv_bfloat16* mh = base + mirror * ld + co0;
        __nv_bfloat16* ml = mh + plane;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < CW / 16; ++j) {
            tc_st256(mh + 16 * j, zero_mirror ? z : hi[2 * j], zero_mirror ? z : hi[2 * j + 1]);
            tc_st256(ml + 16 * j, zero_mirror ? z : lo[2 * j], zero_mirror ? z : lo[2 * j + 1]);
        }
    }
}

template <int CW>
__device__ __forceinline__ void tc_epilogue_conv(const TcCall& c, int b, int t, int n0, float (&v)[CW], const float4 (&bias)[CW / 4]) {
    const int p = n0 / c.Cout, co0 = n0 - p * c.Cout;
    const int t_out = t * c.up + p;
#pragma unroll
    for (int j = 0; j < CW / 4; ++j) {
        v[4 * j] += bias[j].x; v[4 * j + 1] += bias[j].y; v[4 * j + 2] += bias[j].z; v[4 * j + 3] += bias[j].w;
    }
    if (c.f32 != nullptr) {
        float* dst = c.f32 + (b * c.f_sb + t_out * c.f_st + c.f_off) * c.f_ld + co0;
        if (c.f_scalar) {
#pragma unroll
            for (int j = 0; j < CW; ++j)
                if (co0 + j < c.f_valid) dst[j] = v[j];
        } else if (co0 < c.f_valid) {
            float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
            for (int j = 0; j < CW / 4; ++j)
                if (co0 + 4 * j < c.f_valid) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }
    if (c.raw != nullptr || c.elu != nullptr) {
        const long long row = b * c.o_sb + t_out * c.o_st + c.o_off;
        const long long mirror = (t_out >= 1 && t_out <= c.o_halo) ? row - 2ll * t_out * c.o_st : -1ll;
        if (c.raw != nullptr) tc_store_planes<CW>(c.raw, c.o_plane, c.o_ld, row, mirror, co0, v, false, c.o_halo_zero != 0);
        if (c.elu != nullptr) tc_store_planes<CW>(c.elu, c.o_plane, c.o_ld, row, mirror, co0, v, true, c.o_halo_zero != 0);
    }
}

// columns [n0, n0+CW) = gates (i, f, g, o) of hidden units [n0/4, n0/4 + CW/4) of utterance b at step c.t_step
template <int CW>
__device__ __forceinline__ void tc_epilogue_lstm(const TcCall& c, int b, int n0, float (&v)[CW]) {
    constexpr int U = CW / 4;
    const int j0 = n0 >> 2;
    const size_t tb = static_cast<size_t>(c.t_step) * c.Bcap + b;
    const float4* pr = reinterpret_cast<const float4*>(c.pre + tb * (4 * static_cast<size_t>(c.H)) + n0);
    float* cp = c.cst + static_cast<size_t>(b) * c.H + j0;
    float cs[U], h[U];
#pragma unroll
    for (int u = 0; u < U; u += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(cp + u);
        cs[u] = c4.x; cs[u + 1] = c4.y; cs[u + 2] = c4.z; cs[u + 3] = c4.w;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float4 g = pr[u];
        const float ig = tc_sigmoid(v[4 * u] + g.x), fg = tc_sigmoid(v[4 * u + 1] + g.y);
        const float gg = tanhf(v[4 * u + 2] + g.z), og = tc_sigmoid(v[4 * u + 3] + g.w);
        cs[u] = fg * cs[u] + ig * gg;
        h[u] = og * tanhf(cs[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u += 4) *reinterpret_cast<float4*>(cp + u) = make_float4(cs[u], cs[u + 1], cs[u + 2], cs[u + 3]);
    __nv_bfloat16* hp = c.hseq + (tb + c.Bcap) * c.H + j0;          // slot t+1
    const float* sk = c.elu != nullptr ? c.skip + tb * c.H + j0 : nullptr;
    __nv_bfloat16* op = c.elu != nullptr ? c.elu + (b * c.o_sb + c.t_step * c.o_st + c.o_off) * c.o_ld + j0 : nullptr;
#pragma unroll
    for (int u = 0; u < U; u += 4) {
        uint2 hi, lo;
        tc_split2(h[u], h[u + 1], hi.x, lo.x);
        tc_split2(h[u + 2], h[u + 3], hi.y, lo.y);
        *reinterpret_cast<uint2*>(hp + u) = hi;
        *reinterpret_cast<uint2*>(hp + c.h_plane + u) = lo;
        if (op != nullptr) {
            const float4 s4 = *reinterpret_cast<const float4*>(sk + u);
            tc_split2(tc_elu(h[u] + s4.x), tc_elu(h[u + 1] + s4.y), hi.x, lo.x);
            tc_split2(tc_elu(h[u + 2] + s4.z), tc_elu(h[u + 3] + s4.w), hi.y, lo.y);
            *reinterpret_cast<uint2*>(op + u) = hi;
            *reinterpret_cast<uint2*>(op + c.o_plane + u) = lo;
        }
    }
}

template <int CW>
__device__ __forceinline__ void tc_tmem_ld(uint32_t taddr, float (&v)[CW]) {
    if constexpr (CW == 32) tmem_ld_32x32(taddr, v);
    else tmem_ld_32x16(taddr, v);
}

// Persistent: CTA i works on tiles i, i + grid, ... (tile = m-tile * ntiles + n-tile, n fastest so the CTAs that run
// together share their A rows in L2).  The TMA producer and the MMA issuer run ahead across tile boundaries through the
// shared-memory ring; the accumulator is double-buffered in TMEM so the 16 epilogue warps drain tile i while tile i+1 is
// being multiplied.
template <int BN, int STAGES, int CW>
__global__ void __launch_bounds__(TC_THREADS)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ TcCall c) {
    using L = TcSmem<BN, STAGES>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull = empty_bar + STAGES;                  // [2] accumulator ready
    uint64_t* tempty = tfull + 2;                          // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_kb = c.total_kb;
    const int total_tiles = c.mtiles * c.ntiles;
    const int pre = min(total_kb, STAGES);
    constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;

    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], TC_EPI_WARPS);
        }
        mbar_fence_init();
        if (static_cast<int>(blockIdx.x) < total_tiles) {   // weights never depend on the previous kernel
            const int ntile = blockIdx.x % c.ntiles;
            for (int i = 0; i < pre; ++i) {
                mbar_arrive_expect_tx(&full_bar[i], L::STAGE_BYTES);
                tma_load_2d(smem + i * L::STAGE_BYTES + 2 * L::A_BYTES, &tmW, &full_bar[i], 0, (ntile * total_kb + i) * 2 * BN);
            }
        }
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            pdl_wait();                                     // the activation planes come from the previous kernel
            int g = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int ntile = tile % c.ntiles, mtile = tile / c.ntiles;
                const int r0 = c.row_base + mtile * TC_BM;
                for (int i = 0; i < total_kb; ++i, ++g) {
                    const int stage = g % STAGES, use = g / STAGES;
                    uint8_t* a = smem + stage * L::STAGE_BYTES;
                    if (g >= pre) {                         // (the first `pre` k-blocks were armed above, with their weights)
                        if (use > 0) mbar_wait(&empty_bar[stage], (use - 1) & 1);
                        mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                        tma_load_2d(a + 2 * L::A_BYTES, &tmW, &full_bar[stage], 0, (ntile * total_kb + i) * 2 * BN);
                    }
                    const TcTap tp = c.taps[i];
                    const CUtensorMap* m = tp.src ? &tmA1 : &tmA0;
                    const int row = r0 - tp.shift;
                    tma_load_2d(a, m, &full_bar[stage], tp.coff, row);
                    tma_load_2d(a + L::A_BYTES, m, &full_bar[stage], tp.coff, c.rcap[tp.src] + row);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16_f32(TC_BM, BN);
        int g = 0, it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            if (it >= 2) {                                  // the epilogue warps have drained this accumulator
                mbar_wait(&tempty[acc], ((it >> 1) - 1) & 1);
                tc_fence_after();
            }
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int i = 0; i < total_kb; ++i, ++g) {
                const int stage = g % STAGES;
                mbar_wait(&full_bar[stage], (g / STAGES) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint64_t ahi = umma_desc_kmajor_sw128(a_addr);
                    const uint64_t alo = umma_desc_kmajor_sw128(a_addr + L::A_BYTES);
                    const uint64_t bhi = umma_desc_kmajor_sw128(a_addr + 2 * L::A_BYTES);
                    const uint64_t blo = umma_desc_kmajor_sw128(a_addr + 2 * L::A_BYTES + L::B_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        umma_bf16(d_tmem, ahi + 2 * k, bhi + 2 * k, idesc, (i | k) != 0);
                        umma_bf16(d_tmem, alo + 2 * k, bhi + 2 * k, idesc, 1u);
                        umma_bf16(d_tmem, ahi + 2 * k, blo + 2 * k, idesc, 1u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (i == total_kb - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;                             // TMEM lane quarter this warp may read
        const int grp = (warp - 2) >> 2;                    // which column chunks
        pdl_wait();                                         // what the epilogue overwrites may still be read upstream
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int ntile = tile % c.ntiles, mtile = tile / c.ntiles;
            const int acc = it & 1;
            const int row = c.row_base + mtile * TC_BM + q * 32 + lane;
            int b, t;
            bool valid;
            if (c.mode == TC_MODE_LSTM) {
                b = row - c.row_base;
                t = c.t_step;
                valid = b < c.B;
            } else {
                const int qd = row / c.in_div, rm = row - qd * c.in_div;
                if (c.in_tm) { t = qd - c.in_halo; b = rm; }
                else { b = qd; t = rm - c.in_halo; }
                valid = row < c.rows_total && b < c.B && t >= (c.in_store_halo ? -c.in_halo : 0) && t < c.T_in;
            }
            // BN / CW <= 4 chunks and 4 warp groups: a warp owns at most one chunk of every tile.  Its bias is fetched while
            // the accumulator is still being produced.
            static_assert(BN / CW <= TC_EPI_WARPS / 4, "one column chunk per epilogue warp");
            const bool has_chunk = grp < BN / CW;
            const int n0 = ntile * BN + grp * CW;
            float4 bias[CW / 4];
            if (has_chunk) {
                const float4* b4 = reinterpret_cast<const float4*>(c.bias + n0);
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) bias[j] = b4[j];
            }
            mbar_wait(&tfull[acc], (it >> 1) & 1);
            tc_fence_after();
            if (has_chunk) {
                float v[CW];
                tc_tmem_ld<CW>(tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16) + grp * CW, v);
                if (valid && n0 < c.Nstore) {
                    if (c.mode == TC_MODE_LSTM) tc_epilogue_lstm<CW>(c, b, n0, v);
                    else tc_epilogue_conv<CW>(c, b, t, n0, v, bias);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// codes [B][K][T] -> latent planes (row (b, t) = b*Tp + halo + t, D channels of ld), halo mirrored / zeroed
__global__ void __launch_bounds__(128)
tc_rvq_planes_kernel(const long long* __restrict__ codes, const float* const* __restrict__ embed, __nv_bfloat16* out,
                     long long plane, int K, int D, int ld, int T, int Tp, int halo, int halo_zero) {
    const int b = blockIdx.y, t0 = blockIdx.x * 16;
    for (int tt = 0; tt < 16; ++tt) {
        const int t = t0 + tt;
        if (t >= T) break;
        for (int ch = threadIdx.x; ch < ld; ch += blockDim.x) {
            float acc = 0.f;
            if (ch < D)
                for (int q = 0; q < K; ++q)
                    acc += embed[q][static_cast<size_t>(codes[(static_cast<size_t>(b) * K + q) * T + t]) * D + ch];
            __nv_bfloat16 hi, lo;
            split_bf16(acc, hi, lo);
            const long long row = static_cast<long long>(b) * Tp + halo + t;
            out[row * ld + ch] = hi;
            out[plane + row * ld + ch] = lo;
            if (t >= 1 && t <= halo) {
                const long long mr = row - 2ll * t;
                out[mr * ld + ch] = halo_zero ? __float2bfloat16_rn(0.f) : hi;
                out[plane + mr * ld + ch] = halo_zero ? __float2bfloat16_rn(0.f) : lo;
            }
        }
    }
}

// Final convolution (C channels -> 1 channel, k taps), split so that the activation planes are read ONCE:
//   P[row][j] = sum_c w[c][j] * x[row][c]          one k = 1 GEMM with N = k columns (conv_tc_kernel, fp32 rows of 8)
//   out[t]    = bias + sum_j P[t - (k-1-j)][j]     this kernel: k shifted diagonals, staged through shared memory
// (as a k-tap implicit GEMM it would pull k shifted copies of every 128-row tile through shared memory for one output column).
static constexpr int DS_ROWS = 256;
static constexpr int DS_LD = 8;                            // floats per P row (k <= 8)
__global__ void __launch_bounds__(DS_ROWS)
tc_diag_sum_kernel(const float* __restrict__ P, int k, float bias, float* __restrict__ out, int rows_total, int Tp, int halo,
                   int T, int B) {
    __shared__ float ps[(DS_ROWS + DS_LD) * (DS_LD + 1)];
    const long long r0 = static_cast<long long>(blockIdx.x) * DS_ROWS - (k - 1);
    pdl_wait();
    const int W = DS_ROWS + k - 1;
    for (int i = threadIdx.x; i < W * DS_LD; i += DS_ROWS) {  // contiguous floats of P: coalesced
        const long long row = r0 + i / DS_LD;
        ps[(i / DS_LD) * (DS_LD + 1) + (i % DS_LD)] = (row >= 0 && row < rows_total) ? P[row * DS_LD + (i % DS_LD)] : 0.f;
    }
    __syncthreads();
    const long long row = static_cast<long long>(blockIdx.x) * DS_ROWS + threadIdx.x;
    if (row >= rows_total) return;
    const int b = static_cast<int>(row / Tp), t = static_cast<int>(row - static_cast<long long>(b) * Tp) - halo;
    if (b >= B || t < 0 || t >= T) return;
    float acc = bias;
    for (int j = 0; j < k; ++j) acc += ps[(threadIdx.x + j) * (DS_LD + 1) + j];   // window row threadIdx.x + j = t - (k-1-j)
    out[static_cast<size_t>(b) * T + t] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
namespace {

inline int cpad(int c) { return (c + 63) / 64 * 64; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline uint16_t f2bf(float f) {                                // round to nearest even, like __float2bfloat16_rn
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}
inline float bf2f(uint16_t h) {
    uint32_t u = static_cast<uint32_t>(h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

struct TcGemm {
    __nv_bfloat16* tiles = nullptr;
    float* bias = nullptr;
    int N = 0, BN = 0, ntiles = 0, total_kb = 0;
    int Cout = 0, up = 1;              // column n -> (n / Cout, n % Cout)
    CUtensorMap tmW;
    std::vector<TcTap> taps;
};

struct Plane {                         // an activation tensor in the workspace
    __nv_bfloat16* raw = nullptr;
    __nv_bfloat16* elu = nullptr;
    int C = 0;                         // padded channels (= ld)
    int rcap = 0;                      // rows per plane
    long long sb = 0, st = 1, off = 0; // row(b, t)
    int Tp = 0, halo = 0, halo_zero = 0, T = 0, tm = 0;
    long long plane() const { return static_cast<long long>(rcap) * C; }
};

}  // namespace

struct TcCodec {
    enc_config cfg;
    int hop = 1, D = 0, Dp = 0, ch0 = 0, num_sms = 148;
    const float** d_embed = nullptr;
    TcGemm conv_in, conv_out;
    TcGemm conv_out_p;                 // final conv as per-tap partial products (N = k), summed by tc_diag_sum_kernel
    bool co_split = false;
    int lstm_wide = -1;                // VCB_CODEC_LSTM_WIDE: force 64- (0) or 128-column (1) step tiles
    float co_bias = 0.f;
    std::vector<TcGemm> pre, step, step_wide, up;   // step: 64-column tiles (<= 128 utterances), step_wide: 128-column tiles
    std::vector<std::vector<TcGemm>> res1, res2;
    std::vector<void*> owned;
    uint8_t* ws = nullptr;
    size_t ws_bytes = 0;
    size_t ws_limit = 0;
    int min_T = 8;
    bool profile = false;
    std::vector<std::pair<std::string, float>> prof;
    struct Dbg { Plane p; const __nv_bfloat16* ptr; int B; };
    std::map<std::string, Dbg> dbg;      // tensors of the last decoded chunk (debug read-back)
};

namespace {

inline int tc_num_sms(const TcCodec* tc) { return tc->num_sms; }

int upload_gemm(TcCodec* tc, TcGemm& g, const std::vector<float>& W, const std::vector<float>& bias, int N, int Ktot, int bn_hint) {
    if (Ktot % TC_BK || Ktot / TC_BK > TC_MAX_KB) {
        set_error("codec_tc: K = %d outside the kernel's range", Ktot);
        return -1;
    }
    g.N = N;
    g.total_kb = Ktot / TC_BK;
    g.BN = bn_hint ? bn_hint : (N >= 128 ? 128 : (N >= 64 ? 64 : 32));
    g.ntiles = (N + g.BN - 1) / g.BN;
    const int Npad = g.ntiles * g.BN;
    std::vector<uint16_t> t(static_cast<size_t>(Npad) * Ktot * 2);
    size_t o = 0;
    for (int nt = 0; nt < g.ntiles; ++nt)
        for (int kb = 0; kb < g.total_kb; ++kb)
            for (int part = 0; part < 2; ++part)
                for (int r = 0; r < g.BN; ++r) {
                    const int n = nt * g.BN + r;
                    for (int kk = 0; kk < TC_BK; ++kk) {
                        const float w = n < N ? W[static_cast<size_t>(n) * Ktot + kb * TC_BK + kk] : 0.f;
                        const uint16_t hi = f2bf(w);
                        t[o++] = part == 0 ? hi : f2bf(w - bf2f(hi));
                    }
                }
    VCB_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g.tiles), t.size() * 2));
    tc->owned.push_back(g.tiles);
    VCB_CUDA_OK(cudaMemcpy(g.tiles, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
    std::vector<float> bp(Npad, 0.f);
    for (int n = 0; n < N && n < static_cast<int>(bias.size()); ++n) bp[n] = bias[n];
    VCB_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&g.bias), bp.size() * 4));
    tc->owned.push_back(g.bias);
    VCB_CUDA_OK(cudaMemcpy(g.bias, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
    return make_tmap_bf16_2d(&g.tmW, g.tiles, static_cast<uint64_t>(g.ntiles) * g.total_kb * 2 * g.BN, TC_BK, TC_BK, 2 * g.BN);
}

struct HostW {
    const std::map<std::string, float*>& dev;
    const std::map<std::string, std::vector<int64_t>>& shapes;
    int get(const std::string& name, std::vector<float>& out, std::vector<int64_t>* shape = nullptr) const {
        auto it = dev.find(name);
        auto is = shapes.find(name);
        if (it == dev.end() || is == shapes.end()) {
            set_error("codec: missing weight %s", name.c_str());
            return -1;
        }
        size_t n = 1;
        for (auto s : is->second) n *= static_cast<size_t>(s);
        out.resize(n);
        VCB_CUDA_OK(cudaMemcpy(out.data(), it->second, n * 4, cudaMemcpyDeviceToHost));
        if (shape) *shape = is->second;
        return 0;
    }
};

// Conv1d weight w[Cout][Cin][k] (stride 1, dilation dil, causal) -> GEMM [Cout_pad][k * Cin_pad], tap j reads row t - (k-1-j)*dil
int build_conv(TcCodec* tc, TcGemm& g, const HostW& hw, const std::string& name, int Cin, int Cout, int k, int dil, int bn_hint = 0) {
    std::vector<float> w, b;
    if (hw.get(name + ".weight", w) || hw.get(name + ".bias", b)) return -1;
    const int Cip = cpad(Cin), Cop = Cout == 1 ? 1 : cpad(Cout);
    const int Ktot = k * Cip;
    std::vector<float> W(static_cast<size_t>(Cop) * Ktot, 0.f);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int j = 0; j < k; ++j) W[static_cast<size_t>(co) * Ktot + j * Cip + ci] = w[(static_cast<size_t>(co) * Cin + ci) * k + j];
    for (int j = 0; j < k; ++j)
        for (int cb = 0; cb < Cip / TC_BK; ++cb) g.taps.push_back(TcTap{0, static_cast<short>((k - 1 - j) * dil), cb * TC_BK});
    g.Cout = Cop == 1 ? 32 : Cop;
    g.up = 1;
    return upload_gemm(tc, g, W, b, Cop, Ktot, bn_hint);
}

// ConvTranspose1d weight w[Cin][Cout][2r], stride r, causal trim of the last r samples -> GEMM [r * Cout_pad][2 * Cin_pad]
int build_convtr(TcCodec* tc, TcGemm& g, const HostW& hw, const std::string& name, int Cin, int Cout, int r) {
    std::vector<float> w, b;
    if (hw.get(name + ".weight", w) || hw.get(name + ".bias", b)) return -1;
    const int Cip = cpad(Cin), Cop = cpad(Cout), Ktot = 2 * Cip, N = r * Cop;
    std::vector<float> W(static_cast<size_t>(N) * Ktot, 0.f), bias(N, 0.f);
    for (int p = 0; p < r; ++p)
        for (int co = 0; co < Cout; ++co) {
            const int n = p * Cop + co;
            bias[n] = b[co];
            for (int ci = 0; ci < Cin; ++ci) {
                const float* src = &w[(static_cast<size_t>(ci) * Cout + co) * 2 * r];
                W[static_cast<size_t>(n) * Ktot + ci] = src[p];               // x[t]
                W[static_cast<size_t>(n) * Ktot + Cip + ci] = src[p + r];     // x[t-1]
            }
        }
    for (int tap = 0; tap < 2; ++tap)
        for (int cb = 0; cb < Cip / TC_BK; ++cb) g.taps.push_back(TcTap{0, static_cast<short>(tap), cb * TC_BK});
    g.Cout = Cop;
    g.up = r;
    return upload_gemm(tc, g, W, bias, N, Ktot, 0);
}

// residual block tail: conv2 (k = 1, on ELU(hidden)) + shortcut (k = 1, on the raw block input) as one GEMM
int build_res_tail(TcCodec* tc, TcGemm& g, const HostW& hw, const std::string& prefix, int C, int hidden) {
    std::vector<float> w2, b2, ws, bs;
    if (hw.get(prefix + ".conv2.weight", w2) || hw.get(prefix + ".conv2.bias", b2) || hw.get(prefix + ".shortcut.weight", ws) ||
        hw.get(prefix + ".shortcut.bias", bs))
        return -1;
    const int Cp = cpad(C), Hp = cpad(hidden), Ktot = Hp + Cp;
    std::vector<float> W(static_cast<size_t>(Cp) * Ktot, 0.f), bias(Cp, 0.f);
    for (int co = 0; co < C; ++co) {
        bias[co] = b2[co] + bs[co];
        for (int ci = 0; ci < hidden; ++ci) W[static_cast<size_t>(co) * Ktot + ci] = w2[static_cast<size_t>(co) * hidden + ci];
        for (int ci = 0; ci < C; ++ci) W[static_cast<size_t>(co) * Ktot + Hp + ci] = ws[static_cast<size_t>(co) * C + ci];
    }
    for (int cb = 0; cb < Hp / TC_BK; ++cb) g.taps.push_back(TcTap{0, 0, cb * TC_BK});
    for (int cb = 0; cb < Cp / TC_BK; ++cb) g.taps.push_back(TcTap{1, 0, cb * TC_BK});
    g.Cout = Cp;
    g.up = 1;
    return upload_gemm(tc, g, W, bias, Cp, Ktot, 0);
}

// LSTM matrix [4H][H] (gate-major rows i, f, g, o) -> gate-interleaved rows n = 4*unit + gate
int build_lstm(TcCodec* tc, TcGemm& g, const HostW& hw, const std::string& wname, const std::string& b1, const std::string& b2,
               int H, int bn_hint) {
    std::vector<float> w, bi, bh;
    if (hw.get(wname, w)) return -1;
    std::vector<float> bias;
    if (!b1.empty()) {
        if (hw.get(b1, bi) || hw.get(b2, bh)) return -1;
        bias.assign(4 * H, 0.f);
    }
    std::vector<float> W(static_cast<size_t>(4) * H * H);
    for (int j = 0; j < H; ++j)
        for (int gt = 0; gt < 4; ++gt) {
            memcpy(&W[(static_cast<size_t>(4) * j + gt) * H], &w[(static_cast<size_t>(gt) * H + j) * H], static_cast<size_t>(H) * 4);
            if (!bias.empty()) bias[4 * j + gt] = bi[gt * H + j] + bh[gt * H + j];
        }
    for (int cb = 0; cb < H / TC_BK; ++cb) g.taps.push_back(TcTap{0, 0, cb * TC_BK});
    g.Cout = 4 * H;
    g.up = 1;
    return upload_gemm(tc, g, W, bias, 4 * H, H, bn_hint);
}

template <int BN, int STAGES, int CW>
int tc_launch_t(const CUtensorMap& a0, const CUtensorMap& a1, const TcGemm& g, const TcCall& c, int num_sms, cudaStream_t st) {
    using L = TcSmem<BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        VCB_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    const long long tiles = static_cast<long long>(c.mtiles) * c.ntiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(std::min<long long>(tiles, num_sms)), 1, 1);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    VCB_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, STAGES, CW>, a0, a1, g.tmW, c));
    return 0;
}

int tc_launch(TcCodec* tc, const CUtensorMap& a0, const CUtensorMap& a1, const TcGemm& g, TcCall& c, int mtiles, cudaStream_t st) {
    c.total_kb = g.total_kb;
    c.ntiles = g.ntiles;
    c.mtiles = mtiles;
    c.bias = g.bias;
    c.up = g.up;
    c.Cout = g.Cout;
    for (int i = 0; i < g.total_kb; ++i) c.taps[i] = g.taps[i];
    if (static_cast<long long>(mtiles) * g.ntiles > 0x7fffffffll) {
        set_error("codec_tc: too many tiles");
        return -1;
    }
    const int sms = tc_num_sms(tc);
    if (g.BN == 128) return tc_launch_t<128, 3, 32>(a0, a1, g, c, sms, st);
    if (g.BN == 64) return tc_launch_t<64, 4, 16>(a0, a1, g, c, sms, st);
    return tc_launch_t<32, 5, 16>(a0, a1, g, c, sms, st);
}

// A-side tensor map of an activation tensor form (raw / elu): [2 * rcap rows][C]
int plane_map(CUtensorMap* tm, const __nv_bfloat16* base, const Plane& p) {
    return make_tmap_bf16_2d(tm, base, 2ull * p.rcap, p.C, p.C, TC_BM);
}

void set_in(TcCall& c, const Plane& p, int B) {
    c.rows_total = p.rcap;
    c.rcap[0] = p.rcap;
    c.in_tm = p.tm;
    c.in_div = p.tm ? static_cast<int>(p.st) : p.Tp;
    c.in_halo = p.halo;
    c.T_in = p.T;
    c.B = B;
}

void set_out(TcCall& c, const Plane& p, bool raw, bool elu) {
    c.raw = raw ? p.raw : nullptr;
    c.elu = elu ? p.elu : nullptr;
    c.o_plane = p.plane();
    c.o_ld = p.C;
    c.o_sb = p.sb;
    c.o_st = p.st;
    c.o_off = p.off;
    c.o_halo = p.halo;
    c.o_halo_zero = p.halo_zero;
}

struct Bump {
    uint8_t* base;
    size_t off = 0;
    void* take(size_t bytes) {
        void* p = base ? base + off : nullptr;
        off += align_up(bytes, 1024);
        return p;
    }
};

// utterance-major tensor: every utterance = halo rows + T rows
Plane make_um(int C, int B, int T, int halo, int halo_zero) {
    Plane p;
    p.C = C;
    p.T = T;
    p.Tp = T + halo;
    p.halo = halo;
    p.halo_zero = halo_zero;
    p.rcap = std::max(B * p.Tp, TC_BM);               // (a TMA box never taller than its tensor)
    p.sb = p.Tp;
    p.st = 1;
    p.off = halo;
    p.tm = 0;
    return p;
}
// time-major tensor: row = (t + halo) * Bcap + b
Plane make_tm(int C, int Bcap, int T, int halo) {
    Plane p;
    p.C = C;
    p.T = T;
    p.Tp = T + halo;
    p.halo = halo;
    p.halo_zero = 1;
    p.rcap = (T + halo) * Bcap;
    p.sb = 1;
    p.st = Bcap;
    p.off = static_cast<long long>(halo) * Bcap;
    p.tm = 1;
    return p;
}
size_t plane_bytes(const Plane& p, int forms) { return static_cast<size_t>(p.rcap) * p.C * 2 * 2 * forms; }

struct Prof {
    TcCodec* tc;
    cudaStream_t st;
    std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> ev;
    void begin(const char* name) {
        if (!tc->profile) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
        ev.push_back({name, {a, b}});
    }
    void end() {
        if (!tc->profile) return;
        cudaEventRecord(ev.back().second.second, st);
    }
    void finish() {
        if (!tc->profile) return;
        cudaStreamSynchronize(st);
        for (auto& e : ev) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e.second.first, e.second.second);
            tc->prof.push_back({e.first, ms});
            cudaEventDestroy(e.second.first);
            cudaEventDestroy(e.second.second);
        }
    }
};

// One chunk of B utterances.  dry = only measure the workspace (returns bytes through *need).
int decode_chunk_tc(TcCodec* tc, const int64_t* codes, float* wav, int B, int T, cudaStream_t st, int64_t* launches, bool dry,
                    size_t* need) {
    const enc_config& cf = tc->cfg;
    const int Bcap = (B + 127) / 128 * 128;
    const int H = tc->ch0;
    const int nl = cf.lstm;
    const int kin = cf.kernel_size, kres = cf.residual_kernel_size, kout = cf.last_kernel_size;
    const int hz = cf.pad_reflect ? 0 : 1;
    Bump ws{dry ? nullptr : tc->ws};
    auto take_planes = [&](Plane& p, bool raw, bool elu) {
        if (raw) p.raw = static_cast<__nv_bfloat16*>(ws.take(plane_bytes(p, 1)));
        if (elu) p.elu = static_cast<__nv_bfloat16*>(ws.take(plane_bytes(p, 1)));
    };
    // ---- fixed tensors
    Plane Z = make_um(tc->Dp, B, T, kin - 1, hz);
    take_planes(Z, true, false);
    Plane U0 = make_um(cpad(H), B, T, 1, 1);                       // input of the first ConvTranspose: ELU'd, zero halo
    take_planes(U0, false, true);
    Plane X0 = make_tm(H, Bcap, T, 0);                             // conv_in output (LSTM input), time-major
    Plane HS[2] = {make_tm(H, Bcap, T, 1), make_tm(H, Bcap, T, 1)};
    float *x0f = nullptr, *pre = nullptr, *cst = nullptr;
    if (nl > 0) {
        take_planes(X0, true, false);
        take_planes(HS[0], true, false);
        if (nl > 1) take_planes(HS[1], true, false);
        x0f = static_cast<float*>(ws.take(static_cast<size_t>(T) * Bcap * H * 4));
        pre = static_cast<float*>(ws.take(static_cast<size_t>(T) * Bcap * 4 * H * 4));
        cst = static_cast<float*>(ws.take(static_cast<size_t>(nl) * Bcap * H * 4));
    }
    // ---- arenas of the up-sampling stages: block input X (raw + ELU), hidden, block output
    size_t arenaX = 0, arenaH = 0;
    {
        int ch = H, t = T;
        for (int i = 0; i < cf.n_ratios; ++i) {
            t *= cf.ratios[i];
            ch /= 2;
            int maxhalo = std::max(kout - 1, 1);
            for (int j = 0, d = 1; j < cf.n_residual_layers; ++j, d *= cf.dilation_base) maxhalo = std::max(maxhalo, (kres - 1) * d);
            const size_t rows = static_cast<size_t>(B) * (t + maxhalo);
            arenaX = std::max(arenaX, align_up(rows * cpad(ch) * 4, 1024) * 2);
            arenaH = std::max(arenaH, align_up(rows * cpad(ch / cf.compress) * 4, 1024));
        }
    }
    uint8_t* ar[2];
    ar[0] = static_cast<uint8_t*>(ws.take(arenaX));
    ar[1] = static_cast<uint8_t*>(ws.take(arenaX));
    uint8_t* arh = static_cast<uint8_t*>(ws.take(arenaH));
    float* copart = nullptr;                                       // per-tap partial products of the final conv
    if (tc->co_split) {
        int t = T;
        for (int i = 0; i < cf.n_ratios; ++i) t *= cf.ratios[i];
        copart = static_cast<float*>(ws.take((static_cast<size_t>(B) * (t + std::max(kout - 1, 1)) + TC_BM) * DS_LD * 4));
    }
    if (need) *need = ws.off;
    if (dry) return 0;
    auto place = [&](Plane& p, uint8_t* base, bool raw, bool elu) {
        Bump b{base};
        if (raw) p.raw = static_cast<__nv_bfloat16*>(b.take(plane_bytes(p, 1)));
        if (elu) p.elu = static_cast<__nv_bfloat16*>(b.take(plane_bytes(p, 1)));
    };

    Prof pf{tc, st};
    CUtensorMap mA, mB;
    tc->dbg.clear();
    auto note = [&](const std::string& name, const Plane& p, const __nv_bfloat16* ptr) { tc->dbg[name] = TcCodec::Dbg{p, ptr, B}; };
    note("z", Z, Z.raw);
    note("u0", U0, U0.elu);
    if (nl > 0) {
        note("x0", X0, X0.raw);
        note("hs0", HS[0], HS[0].raw);
        if (nl > 1) note("hs1", HS[1], HS[1].raw);
    }
    // state that the kernels only ever read: zero halos / initial LSTM state
    VCB_CUDA_OK(cudaMemset2DAsync(U0.elu, static_cast<size_t>(U0.Tp) * U0.C * 2, 0, static_cast<size_t>(U0.C) * 2, B, st));
    VCB_CUDA_OK(cudaMemset2DAsync(U0.elu + U0.plane(), static_cast<size_t>(U0.Tp) * U0.C * 2, 0, static_cast<size_t>(U0.C) * 2, B, st));
    if (nl > 0) {
        VCB_CUDA_OK(cudaMemsetAsync(cst, 0, static_cast<size_t>(nl) * Bcap * H * 4, st));
        for (int l = 0; l < std::min(nl, 2); ++l) {
            VCB_CUDA_OK(cudaMemsetAsync(HS[l].raw, 0, static_cast<size_t>(Bcap) * H * 2, st));
            VCB_CUDA_OK(cudaMemsetAsync(HS[l].raw + HS[l].plane(), 0, static_cast<size_t>(Bcap) * H * 2, st));
        }
    }
    // ---- RVQ decode -> latent planes
    pf.begin("rvq");
    tc_rvq_planes_kernel<<<dim3((T + 15) / 16, B), 128, 0, st>>>(reinterpret_cast<const long long*>(codes), tc->d_embed, Z.raw,
                                                                 Z.plane(), cf.n_q, tc->D, Z.C, T, Z.Tp, Z.halo, Z.halo_zero);
    VCB_CUDA_OK(cudaGetLastError());
    ++*launches;
    pf.end();
    // ---- conv_in
    pf.begin("conv_in");
    {
        TcCall c;
        memset(&c, 0, sizeof(c));
        set_in(c, Z, B);
        c.Nstore = cpad(H);
        if (nl > 0) {
            set_out(c, X0, true, false);
            c.f32 = x0f;
            c.f_ld = H; c.f_valid = H; c.f_sb = 1; c.f_st = Bcap; c.f_off = 0;
        } else {
            set_out(c, U0, false, true);
        }
        if (plane_map(&mA, Z.raw, Z)) return -1;
        if (tc_launch(tc, mA, mA, tc->conv_in, c, (Z.rcap + TC_BM - 1) / TC_BM, st)) return -1;
        ++*launches;
    }
    pf.end();
    // ---- LSTM stack with skip
    for (int l = 0; l < nl; ++l) {
        const Plane& in = l == 0 ? X0 : HS[(l - 1) & 1];
        Plane& hs = HS[l & 1];
        if (l >= 2) {
            VCB_CUDA_OK(cudaMemsetAsync(hs.raw, 0, static_cast<size_t>(Bcap) * H * 2, st));
            VCB_CUDA_OK(cudaMemsetAsync(hs.raw + hs.plane(), 0, static_cast<size_t>(Bcap) * H * 2, st));
        }
        pf.begin("lstm_ih");
        {
            TcCall c;
            memset(&c, 0, sizeof(c));
            set_in(c, in, B);
            c.Nstore = 4 * H;
            c.f32 = pre;
            c.f_ld = 4 * H; c.f_valid = 4 * H; c.f_sb = 1; c.f_st = Bcap; c.f_off = 0;
            if (plane_map(&mA, in.raw, in)) return -1;
            if (tc_launch(tc, mA, mA, tc->pre[l], c, (in.rcap + TC_BM - 1) / TC_BM, st)) return -1;
            ++*launches;
        }
        pf.end();
        pf.begin("lstm_steps");
        {
            TcCall c;
            memset(&c, 0, sizeof(c));
            c.mode = TC_MODE_LSTM;
            c.rcap[0] = hs.rcap;
            c.rows_total = hs.rcap;
            c.B = B;
            c.Nstore = 4 * H;
            c.pre = pre;
            c.cst = cst + static_cast<size_t>(l) * Bcap * H;
            c.hseq = hs.raw;
            c.h_plane = hs.plane();
            c.Bcap = Bcap;
            c.H = H;
            if (l == nl - 1) {
                c.skip = x0f;
                set_out(c, U0, false, true);
            }
            if (plane_map(&mA, hs.raw, hs)) return -1;
            const int mt = (B + TC_BM - 1) / TC_BM;
            // 128-column tiles would move fewer bytes through L2 per step (65 vs 98 MB at B = 256) but measured slower
            // (18.6 vs 15.1 ms per 800 steps): the step is bound by the 16-k-block pipeline of each CTA, not by L2 throughput
            const bool wide = tc->lstm_wide > 0;
            const TcGemm& sg = wide ? tc->step_wide[l] : tc->step[l];
            for (int t = 0; t < T; ++t) {
                c.t_step = t;
                c.row_base = t * Bcap;
                if (tc_launch(tc, mA, mA, sg, c, mt, st)) return -1;
            }
            *launches += T;
        }
        pf.end();
    }
    // ---- up-sampling stages
    Plane cur = U0;                                                // ELU'd input of the next ConvTranspose
    int ch = H, t_cur = T, side = 0;
    for (int i = 0; i < cf.n_ratios; ++i) {
        const int r = cf.ratios[i];
        const int cout = ch / 2, hidden = cout / cf.compress;
        const bool last_stage = i == cf.n_ratios - 1;
        // ConvTranspose -> X (raw + ELU, halo for the first residual conv)
        Plane X = make_um(cpad(cout), B, t_cur * r, cf.n_residual_layers > 0 ? (kres - 1) : (last_stage ? kout - 1 : 1),
                          cf.n_residual_layers > 0 ? hz : (last_stage ? hz : 1));
        place(X, ar[side], cf.n_residual_layers > 0, true);
        note("x" + std::to_string(i + 1) + ".elu", X, X.elu);
        if (X.raw) note("x" + std::to_string(i + 1) + ".raw", X, X.raw);
        pf.begin("convtr");
        {
            TcCall c;
            memset(&c, 0, sizeof(c));
            set_in(c, cur, B);
            c.Nstore = r * cpad(cout);
            set_out(c, X, cf.n_residual_layers > 0, true);
            if (plane_map(&mA, cur.elu, cur)) return -1;
            if (tc_launch(tc, mA, mA, tc->up[i], c, (cur.rcap + TC_BM - 1) / TC_BM, st)) return -1;
            ++*launches;
        }
        pf.end();
        ch = cout;
        t_cur *= r;
        int dil = 1;
        for (int j = 0; j < cf.n_residual_layers; ++j, dil *= cf.dilation_base) {
            const bool last_res = j == cf.n_residual_layers - 1;
            // conv1 (k3, dilated) on ELU(X) -> hidden, stored ELU'd, in X's row geometry (it is an A source next to X)
            Plane Hd = X;
            Hd.C = cpad(hidden);
            Hd.raw = nullptr;
            Hd.halo_zero = 0;
            place(Hd, arh, false, true);
            note("h" + std::to_string(i + 1) + "." + std::to_string(j), Hd, Hd.elu);
            pf.begin("res_conv1");
            {
                TcCall c;
                memset(&c, 0, sizeof(c));
                set_in(c, X, B);
                c.Nstore = cpad(hidden);
                set_out(c, Hd, false, true);
                c.o_halo = 0;
                if (plane_map(&mA, X.elu, X)) return -1;
                if (tc_launch(tc, mA, mA, tc->res1[i][j], c, (X.rcap + TC_BM - 1) / TC_BM, st)) return -1;
                ++*launches;
            }
            pf.end();
            // conv2 (k1 on ELU(hidden)) + shortcut (k1 on raw X) -> next tensor
            Plane O;
            if (!last_res) O = make_um(cpad(ch), B, t_cur, (kres - 1) * dil * cf.dilation_base, hz);
            else if (last_stage) O = make_um(cpad(ch), B, t_cur, kout - 1, hz);
            else O = make_um(cpad(ch), B, t_cur, 1, 1);
            place(O, ar[side ^ 1], !last_res, true);
            note("o" + std::to_string(i + 1) + "." + std::to_string(j), O, O.elu);
            pf.begin("res_conv2");
            {
                TcCall c;
                memset(&c, 0, sizeof(c));
                set_in(c, Hd, B);
                c.rcap[1] = X.rcap;
                c.Nstore = cpad(ch);
                set_out(c, O, !last_res, true);
                if (plane_map(&mA, Hd.elu, Hd) || plane_map(&mB, X.raw, X)) return -1;
                if (tc_launch(tc, mA, mB, tc->res2[i][j], c, (X.rcap + TC_BM - 1) / TC_BM, st)) return -1;
                ++*launches;
            }
            pf.end();
            X = O;
            side ^= 1;
        }
        cur = X;
        side ^= 1;                                                 // the next ConvTranspose must not write over `cur`
    }
    // ---- conv_out -> waveform
    pf.begin("conv_out");
    if (tc->co_split) {
        TcCall c;
        memset(&c, 0, sizeof(c));
        set_in(c, cur, B);
        c.in_store_halo = 1;
        c.Nstore = 16;
        c.f32 = copart;
        c.f_ld = DS_LD; c.f_valid = DS_LD; c.f_sb = cur.sb; c.f_st = 1; c.f_off = cur.off;
        if (plane_map(&mA, cur.elu, cur)) return -1;
        if (tc_launch(tc, mA, mA, tc->conv_out_p, c, (cur.rcap + TC_BM - 1) / TC_BM, st)) return -1;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((cur.rcap + DS_ROWS - 1) / DS_ROWS);
        lc.blockDim = dim3(DS_ROWS);
        lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at;
        lc.numAttrs = 1;
        VCB_CUDA_OK(cudaLaunchKernelEx(&lc, tc_diag_sum_kernel, static_cast<const float*>(copart), kout, tc->co_bias, wav, cur.rcap,
                                       cur.Tp, cur.halo, t_cur, B));
        *launches += 2;
    } else {
        TcCall c;
        memset(&c, 0, sizeof(c));
        set_in(c, cur, B);
        c.Nstore = 32;
        c.f32 = wav;
        c.f_ld = 1; c.f_valid = 1; c.f_scalar = 1; c.f_sb = t_cur; c.f_st = 1; c.f_off = 0;
        if (plane_map(&mA, cur.elu, cur)) return -1;
        if (tc_launch(tc, mA, mA, tc->conv_out, c, (cur.rcap + TC_BM - 1) / TC_BM, st)) return -1;
        ++*launches;
    }
    pf.end();
    pf.finish();
    return 0;
}

}  // namespace

int tc_codec_build(const enc_config& cfg, const std::map<std::string, float*>& w_dev,
                   const std::map<std::string, std::vector<int64_t>>& shapes, TcCodec** out, const char** reason) {
    *out = nullptr;
    const int ch0 = cfg.n_filters << cfg.n_ratios;
    auto no = [&](const char* why) {
        *reason = why;
        return 1;
    };
    if (!cfg.causal || cfg.trim_right_ratio != 1.0f) return no("non-causal / partial right trim");
    if (cfg.true_skip) return no("identity skip");
    if (cfg.channels != 1) return no("multi-channel output");
    if (cfg.lstm > 0 && ch0 % 64) return no("LSTM width not a multiple of 64");
    if (cfg.lstm == 0 && ch0 % 64) return no("first stage narrower than one k-block");
    if ((ch0 >> cfg.n_ratios) < 1 || cfg.compress < 1) return no("channel plan");
    if (2 * cpad(ch0) / TC_BK > TC_MAX_KB || cfg.kernel_size * cpad(cfg.dimension) / TC_BK > TC_MAX_KB ||
        cfg.residual_kernel_size * cpad(ch0 / 2) / TC_BK > TC_MAX_KB || cfg.last_kernel_size * cpad(cfg.n_filters) / TC_BK > TC_MAX_KB)
        return no("reduction deeper than 64 k-blocks");
    if (getenv("VCB_CODEC_TC") && atoi(getenv("VCB_CODEC_TC")) == 0) return no("disabled by VCB_CODEC_TC=0");
    TcCodec* tc = new TcCodec();
    tc->cfg = cfg;
    tc->D = cfg.dimension;
    tc->Dp = cpad(cfg.dimension);
    tc->ch0 = ch0;
    for (int i = 0; i < cfg.n_ratios; ++i) tc->hop *= cfg.ratios[i];
    {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            tc->num_sms = sms;
        if (getenv("VCB_CODEC_GRID")) tc->num_sms = std::max(1, atoi(getenv("VCB_CODEC_GRID")));
    }
    tc->profile = getenv("VCB_CODEC_PROFILE") && atoi(getenv("VCB_CODEC_PROFILE")) != 0;
    if (getenv("VCB_CODEC_LSTM_WIDE")) tc->lstm_wide = atoi(getenv("VCB_CODEC_LSTM_WIDE"));
    const char* lim = getenv("VCB_CODEC_WS_GB");
    tc->ws_limit = static_cast<size_t>((lim ? atof(lim) : 100.0) * (1ull << 30));
    HostW hw{w_dev, shapes};
    char nm[160];
    int rc = 0;
    {
        std::vector<const float*> emb(cfg.n_q);
        for (int q = 0; q < cfg.n_q && !rc; ++q) {
            snprintf(nm, sizeof(nm), "vq.%d.embed", q);
            auto it = w_dev.find(nm);
            if (it == w_dev.end()) {
                set_error("codec: missing weight %s", nm);
                rc = -1;
            } else {
                emb[q] = it->second;
            }
        }
        if (!rc && (cudaMalloc(reinterpret_cast<void**>(&tc->d_embed), cfg.n_q * sizeof(float*)) != cudaSuccess ||
                    cudaMemcpy(tc->d_embed, emb.data(), cfg.n_q * sizeof(float*), cudaMemcpyHostToDevice) != cudaSuccess)) {
            set_error("codec_tc: codebook pointer table");
            rc = -1;
        }
        if (!rc) tc->owned.push_back(tc->d_embed);
    }
    if (!rc) rc = build_conv(tc, tc->conv_in, hw, "dec.conv_in", cfg.dimension, ch0, cfg.kernel_size, 1);
    tc->pre.resize(cfg.lstm);
    tc->step.resize(cfg.lstm);
    tc->step_wide.resize(cfg.lstm);
    for (int l = 0; l < cfg.lstm && !rc; ++l) {
        char a[96], b[96], c2[96];
        snprintf(a, sizeof(a), "dec.lstm.weight_ih_l%d", l);
        snprintf(b, sizeof(b), "dec.lstm.bias_ih_l%d", l);
        snprintf(c2, sizeof(c2), "dec.lstm.bias_hh_l%d", l);
        rc = build_lstm(tc, tc->pre[l], hw, a, b, c2, ch0, 0);
        snprintf(a, sizeof(a), "dec.lstm.weight_hh_l%d", l);
        if (!rc) rc = build_lstm(tc, tc->step[l], hw, a, "", "", ch0, 64);
        if (!rc) rc = build_lstm(tc, tc->step_wide[l], hw, a, "", "", ch0, 128);
    }
    tc->up.resize(cfg.n_ratios);
    tc->res1.resize(cfg.n_ratios);
    tc->res2.resize(cfg.n_ratios);
    int ch = ch0;
    for (int i = 0; i < cfg.n_ratios && !rc; ++i) {
        snprintf(nm, sizeof(nm), "dec.up%d.convtr", i);
        rc = build_convtr(tc, tc->up[i], hw, nm, ch, ch / 2, cfg.ratios[i]);
        ch /= 2;
        tc->res1[i].resize(cfg.n_residual_layers);
        tc->res2[i].resize(cfg.n_residual_layers);
        int dil = 1;
        for (int j = 0; j < cfg.n_residual_layers && !rc; ++j, dil *= cfg.dilation_base) {
            snprintf(nm, sizeof(nm), "dec.up%d.res%d.conv1", i, j);
            rc = build_conv(tc, tc->res1[i][j], hw, nm, ch, ch / cfg.compress, cfg.residual_kernel_size, dil);
            snprintf(nm, sizeof(nm), "dec.up%d.res%d", i, j);
            if (!rc) rc = build_res_tail(tc, tc->res2[i][j], hw, nm, ch, ch / cfg.compress);
            tc->min_T = std::max(tc->min_T, (cfg.residual_kernel_size - 1) * dil * cfg.dilation_base + 2);
        }
    }
    if (!rc) rc = build_conv(tc, tc->conv_out, hw, "dec.conv_out", ch, 1, cfg.last_kernel_size, 1, 32);
    if (!rc && cfg.last_kernel_size <= DS_LD && !(getenv("VCB_CODEC_CONVOUT_TC") && atoi(getenv("VCB_CODEC_CONVOUT_TC")))) {
        std::vector<float> w, b;
        rc = hw.get("dec.conv_out.weight", w) || hw.get("dec.conv_out.bias", b) ? -1 : 0;
        if (!rc) {
            const int Cp = cpad(ch), k = cfg.last_kernel_size;
            std::vector<float> W(static_cast<size_t>(k) * Cp, 0.f);          // GEMM row j = tap j
            for (int ci = 0; ci < ch; ++ci)
                for (int j = 0; j < k; ++j) W[static_cast<size_t>(j) * Cp + ci] = w[static_cast<size_t>(ci) * k + j];
            TcGemm& g = tc->conv_out_p;
            for (int cb = 0; cb < Cp / TC_BK; ++cb) g.taps.push_back(TcTap{0, 0, cb * TC_BK});
            g.Cout = 32;
            g.up = 1;
            rc = upload_gemm(tc, g, W, std::vector<float>(), k, Cp, 32);
            tc->co_bias = b[0];
            tc->co_split = rc == 0;
        }
    }
    tc->min_T = std::max(tc->min_T, std::max(cfg.kernel_size, cfg.last_kernel_size) + 1);
    if (rc) {
        tc_codec_destroy(tc);
        return -1;
    }
    *out = tc;
    return 0;
}

bool tc_codec_accepts(const TcCodec* c, int B, int T) { return c != nullptr && B >= 1 && T >= c->min_T; }

int tc_codec_decode(TcCodec* tc, const int64_t* codes, float* wav, int B, int T, cudaStream_t st, int64_t* launches) {
    tc->prof.clear();
    // chunk the batch so the workspace stays under the limit -- and halve the chunk again if the device cannot give that much
    int chunk = B;
    size_t need = 0;
    for (;;) {
        int64_t dummy = 0;
        if (decode_chunk_tc(tc, nullptr, nullptr, chunk, T, st, &dummy, true, &need)) return -1;
        if (need > tc->ws_limit && chunk > 1) {
            chunk = (chunk + 1) / 2;
            continue;
        }
        if (need <= tc->ws_bytes) break;
        if (tc->ws) {
            VCB_CUDA_OK(cudaStreamSynchronize(st));
            cudaFree(tc->ws);
            tc->ws = nullptr;
            tc->ws_bytes = 0;
        }
        const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&tc->ws), need);
        if (e == cudaSuccess) {
            tc->ws_bytes = need;
            break;
        }
        cudaGetLastError();                                // clear the sticky allocation error
        tc->ws = nullptr;
        if (chunk == 1) {
            set_error("codec_tc: cannot allocate a %.2f GB workspace for one utterance of %d frames (%s)", need / 1073741824.0, T,
                      cudaGetErrorString(e));
            return -1;
        }
        chunk = (chunk + 1) / 2;
    }
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        if (decode_chunk_tc(tc, codes + static_cast<size_t>(b0) * tc->cfg.n_q * T, wav + static_cast<size_t>(b0) * T * tc->hop, nb, T, st,
                            launches, false, nullptr))
            return -1;
    }
    return 0;
}

void tc_codec_destroy(TcCodec* tc) {
    if (!tc) return;
    for (auto p : tc->owned) cudaFree(p);
    cudaFree(tc->ws);
    delete tc;
}

const std::vector<std::pair<std::string, float>>& tc_codec_profile(const TcCodec* c) { return c->prof; }

int tc_codec_debug_tensor(TcCodec* tc, const char* name, float* host_out, int64_t cap, int32_t* dims) {
    auto it = tc->dbg.find(name);
    if (it == tc->dbg.end()) {
        set_error("codec_tc: no tensor '%s' in the last decode", name);
        return -1;
    }
    const Plane& p = it->second.p;
    const int B = it->second.B;
    dims[0] = B; dims[1] = p.C; dims[2] = p.Tp; dims[3] = p.halo;
    const int64_t n = static_cast<int64_t>(B) * p.C * p.Tp;
    if (host_out == nullptr) return 0;
    if (cap < n) {
        set_error("codec_tc: debug buffer too small");
        return -1;
    }
    VCB_CUDA_OK(cudaDeviceSynchronize());
    std::vector<uint16_t> h(static_cast<size_t>(p.rcap) * p.C * 2);
    VCB_CUDA_OK(cudaMemcpy(h.data(), it->second.ptr, h.size() * 2, cudaMemcpyDeviceToHost));
    const size_t pl = static_cast<size_t>(p.plane());
    for (int b = 0; b < B; ++b)
        for (int t = -p.halo; t < p.T; ++t) {
            const long long row = b * p.sb + t * p.st + p.off;
            for (int ch = 0; ch < p.C; ++ch)
                host_out[(static_cast<size_t>(b) * p.C + ch) * p.Tp + (t + p.halo)] = bf2f(h[row * p.C + ch]) + bf2f(h[pl + row * p.C + ch]);
        }
    return 0;
}

}  // namespace vcb
