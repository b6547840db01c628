/* vcb200.h -- C ABI of libvcb200.so, the B200 (sm_100a) codec-LM decode + EnCodec decode engine.
 *
 * Drop-in boundary for the VoiceCraft hot path.  The reference has no FFI of its
 * own (it is pure PyTorch); each entry point below names the reference code it replaces.  Plain pointers
 * and sizes only -- no torch types.  All `dev` pointers are CUDA device pointers owned by the caller;
 * `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Every function returns 0 on
 * success and a negative value on error; vcb_last_error() then describes it.  Nothing throws across the ABI.
 *
 * Threading: one engine per device; calls on one engine must be serialised by the caller (the reference is
 * single-threaded Python, inference_tts_scale.py:42).  All launches are asynchronous on `stream`; only
 * vcb_poll / vcb_read_tokens synchronise.
 */
#ifndef VCB200_H_
#define VCB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vcb_engine vcb_engine;

enum { VCB_MODE_TTS = 0, VCB_MODE_EDIT = 1 };
enum { VCB_KV_BF16 = 0, VCB_KV_FP32 = 1 };

/* Model hyper-parameters: the argparse Namespace the reference model is built from
 * (reference config.py:50-84, models/voicecraft.py:106-195). */
typedef struct {
    int32_t d_model, nhead, num_layers, n_codebooks;
    int32_t audio_vocab_size, n_special, text_vocab_rows; /* text_vocab_size + 1 */
    int32_t empty_token, eog, audio_pad_token, eos;       /* eos <= 0: unused */
    int32_t encodec_sr, max_n_spans;
    int32_t max_slots;      /* concurrently open utterances */
    int32_t max_seq_len;    /* text + audio columns per utterance, upper bound */
    int32_t max_new_tokens; /* token-log capacity per utterance */
    int32_t kv_dtype;       /* VCB_KV_BF16 (default) or VCB_KV_FP32 */
    int32_t device;         /* CUDA device ordinal */
} vcb_config;

/* Sampling arguments of inference_tts / inference / inference_tts_batch (voicecraft.py:908-920). */
typedef struct {
    int32_t top_k;
    float top_p;
    float temperature;
    int32_t stop_repetition;
    int32_t n_silence;
    int32_t silence_tokens[8];
} vcb_sampling;

/* One utterance (or one best-of-N group) to prefill.  `y_tokens` is the already arranged prompt:
 * the delayed pattern of voicecraft.py:961-972 (TTS) or the segment/placeholder layout of :615-683 (edit),
 * [y_len][n_codebooks] int64 on the device. */
typedef struct {
    int32_t slot;            /* first slot; a group occupies slot .. slot+n_copies-1 */
    int32_t n_copies;        /* 1, or batch_size of inference_tts_batch (voicecraft.py:1329-1343) */
    int32_t mode;            /* VCB_MODE_TTS / VCB_MODE_EDIT */
    int32_t x_len;
    const int64_t* text_ids_dev;    /* [x_len] */
    int32_t y_len;
    const int64_t* y_tokens_dev;    /* [y_len][K] */
    const int32_t* mask_rows_dev;   /* [y_len] or NULL: >= 0 -> column embedding = mask_embedding[row] (:311-320) */
    int32_t n_more_spans;           /* edit: masked spans after the first (voicecraft.py:681, 838-858) */
    int32_t more_mask_rows[8];      /* mask_embedding row of each further span */
} vcb_prompt;

/* Per-slot status returned by vcb_poll. */
typedef struct {
    int32_t done;      /* 1: generation finished (all codebooks ended, all spans); 2: stopped, token-log / KV capacity exhausted */
    int32_t forced;    /* >0: the next decode step feeds a forced embedding and consumes no noise */
    int32_t n_steps;   /* sampling steps recorded so far */
    int32_t keep;      /* best-of-N: member index whose tokens are the result (-1 while undecided) */
    int32_t n_spans_done;
    int32_t span_ends[8];
} vcb_status;

const char* vcb_last_error(void);
int vcb_version(void);

/* ---- life cycle: replaces VoiceCraft.__init__ / load_state_dict (voicecraft.py:106-195) ------------ */
int vcb_create(const vcb_config* cfg, vcb_engine** out);
int vcb_destroy(vcb_engine* e);
/* key = reference state_dict key, data = fp32 device or host pointer, row-major. */
int vcb_load_weight(vcb_engine* e, const char* key, const float* data, const int64_t* shape, int32_t ndim,
                    int32_t is_device_ptr);
/* sinusoidal table of SinePositionalEmbedding (embedding.py:67-92), fp32 [rows][d_model] */
int vcb_load_pe(vcb_engine* e, const float* data, int32_t rows, int32_t is_device_ptr);
int vcb_finalize_weights(vcb_engine* e);   /* packs bf16 GEMM operands, builds TMA descriptors */

/* ---- decode: replaces dec_forward + the sampling loop (voicecraft.py:406-470, 1018-1120) ----------- */
int vcb_prefill(vcb_engine* e, const vcb_prompt* prompts, int32_t n, void* stream);
/* final LayerNorm + logit heads + fused sampler on the last hidden state of each listed slot.
 * exp_noise_dev: [n * K][V] fp32 Exp(1) noise, the draw torch.multinomial makes (voicecraft.py:85). */
int vcb_sample(vcb_engine* e, const int32_t* slots, int32_t n, const float* exp_noise_dev,
               const vcb_sampling* sp, void* stream);
/* one transformer step on the embeddings produced by the previous sample, then vcb_sample. */
int vcb_decode_step(vcb_engine* e, const int32_t* slots, int32_t n, const float* exp_noise_dev,
                    const vcb_sampling* sp, void* stream);
int vcb_poll(vcb_engine* e, const int32_t* slots, int32_t n, vcb_status* out_host, void* stream);
/* copies the raw (still delayed) sampled tokens [n_steps][K] int32 to host memory */
int vcb_read_tokens(vcb_engine* e, int32_t slot, int32_t* out_host, int32_t max_steps, void* stream);
int vcb_release(vcb_engine* e, int32_t slot, int32_t n_copies);

/* debug / parity hooks */
int vcb_debug_logits(vcb_engine* e, float* out_dev, int32_t n_rows);   /* last sampled logits [n*K][V] (pre-edit) */
int vcb_debug_gemm(const float* W_dev /*[N][K]*/, const float* X_dev /*[B][K]*/, float* out_dev /*[B][N]*/, int32_t N,
                   int32_t K, int32_t B, int32_t splits /*<=0: auto*/, int32_t simt);
/* same check for the rows-as-M prefill GEMM (csrc/gemm_rows.cu): any number of rows, N % 128 == 0, K % 64 == 0 */
int vcb_debug_gemm_rows(const float* W_dev /*[N][K]*/, const float* X_dev /*[rows][K]*/, float* out_dev /*[rows][N]*/,
                        int32_t N, int32_t K, int32_t rows);
/* debug timeline: enable=1 starts recording (tag, globaltimer ns) pairs from CTA 0 of each kernel; enable=0 stops and
 * copies up to max_records pairs to out_host */
int vcb_timeline(int32_t enable, uint64_t* out_host, int32_t max_records, int32_t* n_out);
/* micro-benchmark of the GEMM kernel alone (HBM-resident weights): average microseconds per launch */
int vcb_bench_gemm(int32_t N, int32_t K, int32_t B, int32_t splits, int32_t stages, int32_t pdl, int32_t iters,
                   int32_t ncopies, float* us_out);
int vcb_set_option(vcb_engine* e, const char* name, int32_t value);   /* "gemm_simt", "pdl", "profile" */
/* profile mode: summed device ms and launch counts per kernel class since the last read
 * (0 gemm, 1 attention, 2 layernorm/reduce, 3 bias/act/qkv finish, 4 sampler, 5 misc) */
int vcb_profile_read(vcb_engine* e, double* ms_by_class, int64_t* count_by_class, int32_t n_classes);
int64_t vcb_counter(vcb_engine* e, const char* name);                 /* "launches", "kv_bytes", ... */

/* ---- delayed codebook pattern on the device: Pattern.build_pattern_sequence
 *      (codebooks_patterns.py:151-176 with DelayedPatternProvider :336-352, delays = 0..K-1) -------- */
int vcb_delay_pattern(const int64_t* z_dev /*[B][K][T]*/, int64_t* out_dev /*[B][K][T+K]*/, int32_t B, int32_t K,
                      int32_t T, int64_t special_token, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VCB200_H_ */
