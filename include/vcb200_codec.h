/* vcb200_codec.h -- C ABI of the EnCodec decoder (token -> waveform) and encoder (waveform -> token) in libvcb200.so.
 *
 * Replaces AudioTokenizer.decode (reference data/tokenizer.py:131-133), i.e. audiocraft's
 * EncodecModel.decode = ResidualVectorQuantizer.decode + SEANetDecoder, with hand-written sm_100a kernels
 * (RVQ gather-sum, implicit-GEMM Conv1d / ConvTranspose1d with fused ELU / bias / residual, LSTM).
 * Same conventions as vcb200.h: plain pointers, 0 on success, vcb_last_error() for the message.
 */
#ifndef VCB200_CODEC_H_
#define VCB200_CODEC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct enc_engine enc_engine;

/* SEANet / RVQ hyper-parameters (audiocraft config `seanet.*`, `rvq.*`; defaults of the 16 kHz / 50 Hz / 4x2048 codec) */
typedef struct {
    int32_t n_q, bins, dimension, n_filters;
    int32_t n_ratios, ratios[8];
    int32_t kernel_size, last_kernel_size, residual_kernel_size, dilation_base, n_residual_layers, compress;
    int32_t lstm;          /* LSTM layers (0 = none) */
    int32_t causal;        /* causal convolutions (left padding / right trim) */
    int32_t pad_reflect;   /* 1 = 'reflect' padding, 0 = zeros */
    int32_t true_skip;     /* 1 = identity skip in residual blocks, 0 = 1x1 conv shortcut */
    int32_t channels;
    float trim_right_ratio;
    int32_t device;
} enc_config;

int enc_create(const enc_config* cfg, enc_engine** out);
int enc_destroy(enc_engine* e);
/* name = "vq.{q}.embed", "dec.conv_in.weight", "dec.lstm.weight_ih_l0", "dec.up{i}.convtr.weight",
 * "dec.up{i}.res{j}.conv1.weight", ... (weight-norm already folded: w = g * v / ||v||), fp32 row-major. */
int enc_load_weight(enc_engine* e, const char* name, const float* data, const int64_t* shape, int32_t ndim,
                    int32_t is_device_ptr);
int enc_finalize(enc_engine* e);
/* codes [B][n_q][T] int64 (device) -> wav [B][channels][T * hop] fp32 (device) */
int enc_decode(enc_engine* e, const int64_t* codes_dev, float* wav_dev, int32_t B, int32_t T, void* stream);
/* wav [B][channels][N] fp32 (device) -> codes [B][n_q][T] int64 (device), T = N down-sampled by every ratio (rounded up).
 * Replaces AudioTokenizer.encode (reference data/tokenizer.py:127-129 -> audiocraft EncodecModel.encode = SEANetEncoder +
 * ResidualVectorQuantizer.encode).  Needs the "enc.*" weights: "enc.conv_in.weight", "enc.down{i}.res{j}.conv1.weight",
 * "enc.down{i}.conv.weight" (strided), "enc.lstm.*", "enc.conv_out.weight". */
int enc_encode(enc_engine* e, const float* wav_dev, int64_t* codes_dev, int32_t B, int32_t N, void* stream);
int64_t enc_counter(enc_engine* e, const char* name); /* "launches", "hop", "flops_per_frame", "tc_enabled", "tc_decodes" */
/* Debug / tests: an intermediate tensor of the last enc_decode on the tensor-core path ("z", "x0", "u0", "x1.raw", "x1.elu",
 * "h1.0", "o1.0", ...), reassembled from its bf16 (hi, lo) planes as fp32 [B][C][halo + T] on the host.  dims = {B, C, halo + T,
 * halo}; host_out == NULL only queries dims.  The up-sampling stages share two workspace arenas, so after a full decode only
 * the tensors of the last stage (and "z", "x0", "u0", "hs*") still hold their values: see scripts/codec_tc_debug.py. */
int enc_debug_tensor(enc_engine* e, const char* name, float* host_out, int64_t cap, int32_t* dims);

#ifdef __cplusplus
}
#endif
#endif /* VCB200_CODEC_H_ */
