/*
 * dsb200.h — C ABI of libdsb200.so: the B200-native (sm_100a) implementation of
 * deepsignal's call_mods hot path.
 *
 * The reference (bioinfomaticsCSU/deepsignal v0.2.0) has no FFI / plugin API; the
 * boundary this library replaces is two Python call sites:
 *
 *   forward     deepsignal/call_modifications.py:168-178  (sess.run inside _call_mods,
 *               graph = deepsignal/model.py:25-108 + deepsignal/layers.py:20-264)
 *   extraction  deepsignal/extract_features.py:215-286    (_extract_features per-read math:
 *               :211-212 rescale, :143-151 normalise, :154-190 window, :254-276 site loop;
 *               motif scan deepsignal/utils/process_utils.py:95-110)
 *
 * Conventions
 *  - plain C, no C++ exceptions cross the ABI; every function returns 0 on success or a
 *    negative dsb_status; dsb_last_error(ctx) gives the message.
 *  - one dsb_ctx per GPU; a ctx is not thread-safe, independent ctxs are.
 *  - "_dev" entry points take DEVICE pointers owned by the caller and are asynchronous on
 *    the given cudaStream_t (passed as void*; NULL = legacy default stream).  "_host" entry
 *    points take HOST pointers, stage through pinned buffers, and return after the results
 *    are on the host.  The library never frees caller memory.
 *  - there is no CPU fallback and no backend dispatch: without a CUDA device (or with a
 *    non-sm_100 device for the tensor-core path) calls fail with DSB_ERR_CUDA / DSB_ERR_ARCH.
 */
#ifndef DSB200_H_
#define DSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSB_ABI_VERSION 2

#if defined(__GNUC__)
#define DSB_API __attribute__((visibility("default")))
#else
#define DSB_API
#endif

typedef struct dsb_ctx dsb_ctx;

typedef enum dsb_status {
  DSB_OK = 0,
  DSB_ERR_INVALID = -1,   /* bad argument                                  */
  DSB_ERR_CUDA = -2,      /* CUDA runtime / driver error                   */
  DSB_ERR_WEIGHTS = -3,   /* missing / mis-shaped checkpoint tensor        */
  DSB_ERR_STATE = -4,     /* e.g. forward before load_weights              */
  DSB_ERR_CAPACITY = -5,  /* caller-provided output capacity too small     */
  DSB_ERR_ARCH = -6,      /* tensor-core path needs an sm_100 device       */
  DSB_ERR_NOMEM = -7,
  DSB_ERR_NONFINITE = -8  /* a 16-bit activation overflowed: some logits are inf/NaN (outputs are written;
                             rerun those samples with DSB_PREC_FP32)          */
} dsb_status;

/* model-variant and precision flags of dsb_load_weights (reference flags --is_cnn/--is_rnn/
 * --is_base, deepsignal/deepsignal.py:246-257; model.py:59-95) */
#define DSB_FLAG_CNN 1u
#define DSB_FLAG_RNN 2u
#define DSB_FLAG_BASE 4u
#define DSB_FLAG_DEFAULT (DSB_FLAG_CNN | DSB_FLAG_RNN | DSB_FLAG_BASE)

typedef enum dsb_precision {
  DSB_PREC_FP32 = 0,      /* CUDA-core fp32 path, 1e-4 abs parity target   */
  DSB_PREC_BF16 = 1,      /* tcgen05 kind::f16 path, bf16 operands, fp32 accumulate */
  DSB_PREC_FP16 = 2,      /* same kernels and rate with fp16 operands (10-bit mantissa: 8x smaller
                             rounding error than bf16; activations must stay below 65504)       */
  DSB_PREC_FP16X2 = 3,    /* split-operand tensor-core path: every CNN activation is carried as hi + lo fp16
                             parts (22 mantissa bits), each conv is [xh | xl] . [W ; W] on tcgen05 with fp16
                             weights and fp32 accumulation.  This is the 16-bit mode that meets the 2e-3 /
                             99.9 % tolerance on spread-logit checkpoints; the LSTM runs
                             as in DSB_PREC_FP16.  The residual error is the fp16 rounding of the CNN weights
                             and therefore depends on the checkpoint (0.6 - 2.4e-3 over random-init ones). */
  DSB_PREC_FP16X3 = 4     /* as FP16X2 plus the third product term xh . wl (the CNN weights are hi + lo fp16 too):
                             ~1e-5 whatever the checkpoint, on the one-launch-per-conv tcgen05 plan (the fused
                             kernels keep one copy of the weights in shared memory), roughly half the rate of
                             FP16X2.  The tensor-core mode whose tolerance does not depend on the weights. */
} dsb_precision;

/* One checkpoint tensor, host float32, keyed by its TF-1.x variable name (the names a
 * tf.train.Saver bundle of the reference model holds; restore site
 * call_modifications.py:210-211).  The library copies and folds; caller keeps ownership. */
typedef struct dsb_tensor {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} dsb_tensor;

/* ---- context -------------------------------------------------------------------------- */
DSB_API int dsb_abi_version(void);
DSB_API int dsb_create(int device, dsb_ctx** out);
DSB_API int dsb_destroy(dsb_ctx* ctx);
DSB_API const char* dsb_last_error(const dsb_ctx* ctx);   /* ctx may be NULL: last create error */

/* ---- model: replaces Model(...) + Saver.restore + sess.run ------------------------------ */
/* kmer_len / signal_len: reference --kmer_len / --cent_signals_len (17 / 360). */
DSB_API int dsb_load_weights(dsb_ctx* ctx, const dsb_tensor* tensors, int32_t n_tensors,
                     uint32_t flags, int32_t kmer_len, int32_t signal_len);

/* Largest number of samples processed per internal pass (workspace grows with it: ~330 KB per sample on the 16-bit
 * paths).  Default 37888 = 148 SMs x 256 samples, the size at which the persistent LSTM kernel is perfectly balanced;
 * smaller values (multiples of 128) trade throughput for memory. */
DSB_API int dsb_set_max_chunk(dsb_ctx* ctx, int64_t max_chunk);

/* How the forward plan is built.  Defaults (all fused) are the product path; the un-fused variants are the
 * references the fused kernels are tested against, and the ablate_* fields make a kernel skip part of its work for
 * timing experiments (wrong results by design).  Takes effect for plans built after the call (plans are rebuilt). */
typedef struct dsb_plan_options {
  int32_t fused_blocks;   /* 1: inception kernels A + B per block; 0: one GEMM launch per conv + pool kernels   */
  int32_t fused_front;    /* 1: conv1 + pool + conv2 + conv3 in one kernel (single 16-bit modes); 0: stem + 2 GEMMs */
  int32_t lstm_seq;       /* 1: whole BiLSTM in one persistent kernel; 0: one GEMM launch per (dir, layer, step) */
  int32_t ablate_a, ablate_b, ablate_gemm, ablate_lstm;
  int32_t lstm_sms;       /* 0: every kernel sizes its grid to the whole GPU (the streams time-slice the SMs);
                             k > 0: the LSTM kernel runs on k SMs beside the CNN kernels on the other SMs          */
} dsb_plan_options;
DSB_API int dsb_get_plan_options(const dsb_ctx* ctx, dsb_plan_options* out);
DSB_API int dsb_set_plan_options(dsb_ctx* ctx, const dsb_plan_options* opts);
DSB_API size_t dsb_workspace_bytes(const dsb_ctx* ctx, int64_t n, int32_t precision);

/* Forward of n samples (the feed_dict / fetches of call_modifications.py:168-178):
 *   kmer    int32 [n, kmer_len]   base codes 0..4 (A C G T N)      -> model.base_int
 *   means   f32   [n, kmer_len]                                     -> model.means
 *   stds    f32   [n, kmer_len]                                     -> model.stds
 *   lens    f32   [n, kmer_len]   signal counts per base            -> model.sanums
 *   signals f32   [n, signal_len]                                   -> model.signals
 *   probs   f32   [n, 2]          sigmoid(logits) = model.activation_logits
 *   pred    int32 [n]             argmax(probs), ties -> 0 = model.prediction
 */
DSB_API int dsb_forward_dev(dsb_ctx* ctx, int32_t precision, int64_t n,
                    const int32_t* kmer, const float* means, const float* stds, const float* lens,
                    const float* signals, float* probs, int32_t* pred, void* cuda_stream);
DSB_API int dsb_forward_host(dsb_ctx* ctx, int32_t precision, int64_t n,
                     const int32_t* kmer, const float* means, const float* stds, const float* lens,
                     const float* signals, float* probs, int32_t* pred);
/* Samples whose logits were not finite since the last call of this function (the counter is reset).  dsb_forward_host
 * checks it itself and returns DSB_ERR_NONFINITE; after dsb_forward_dev / dsb_extract_call_dev call this once the
 * stream has been synchronised. */
DSB_API int dsb_nonfinite_count(dsb_ctx* ctx, int64_t* count);

/* Per-kernel-class timing of the calls made since profiling was switched on (CUDA events on the launch
 * stream around every launch of the class; the forward runs single-stream while profiling).  Classes:
 * 0 generic tensor-core / fp32 GEMM launches, 1 LSTM kernel(s), 2 elementwise helpers (pools, feature rows, head),
 * 3 inception kernel A, 4 inception kernel B, 5 fused front end (conv1..conv3);
 * extraction chain: 6 read_norm, 7 site count / scan / emit, 8 LUT, 9 unused, 10 site_feature (incl. the per-base statistics). */
DSB_API int dsb_set_profiling(dsb_ctx* ctx, int32_t enabled);
DSB_API int dsb_get_profile(dsb_ctx* ctx, int32_t cls, double* ms_total, int64_t* launches, double* flops);
DSB_API int64_t dsb_launch_count(const dsb_ctx* ctx);   /* kernels launched by this ctx so far */

/* ---- feature extraction: replaces the math of _extract_features ---------------------------- */
typedef struct dsb_reads {
  int64_t n_reads;
  const int16_t* raw;          /* [raw_off[n_reads]] concatenated Raw/Reads/ * /Signal         */
  const int64_t* raw_off;      /* [n_reads+1], raw_off[0] == 0                                 */
  const double* scaling;       /* [n_reads] range / digitisation (extract_features.py:198-201) */
  const double* offset;        /* [n_reads] channel offset                                     */
  const int32_t* ev_start;     /* [ev_off[n_reads]] Events.start + read_start_rel_to_raw        */
  const int32_t* ev_len;       /* Events.length                                                */
  const uint8_t* ev_base;      /* Events.base (ASCII)                                          */
  const int64_t* ev_off;       /* [n_reads+1], ev_off[0] == 0                                  */
  const int64_t* chrom_start;  /* [n_reads] Alignment.mapped_start                             */
  const uint8_t* strand_minus; /* [n_reads] 1 if Alignment.mapped_strand == '-'                */
  const int64_t* chrom_len;    /* [n_reads] contig length, -1 without --reference_path         */
} dsb_reads;

typedef struct dsb_extract_params {
  int32_t kmer_len;            /* odd, <= 33 (default 17)                                      */
  int32_t signal_len;          /* default 360                                                  */
  int32_t normalize_method;    /* 0 = mad, 1 = zscore  (--normalize_method)                    */
  int32_t mod_loc;             /* --mod_loc                                                    */
  int32_t n_motifs;            /* expanded motif sequences, all of motif_len                   */
  int32_t motif_len;           /* <= 16                                                        */
  const char* motifs;          /* HOST pointer, n_motifs * motif_len ASCII bytes               */
} dsb_extract_params;

/* Caller-allocated per-site outputs, `capacity` sites each; any pointer may be NULL to skip
 * that output.  Sites are ordered by read, then by location in the read. */
typedef struct dsb_sites {
  int64_t capacity;
  int32_t* site_read;          /* [cap]      index of the read                                  */
  int32_t* site_loc;           /* [cap]      loc_in_read                                        */
  int64_t* pos;                /* [cap]      extract_features.py:255-260                        */
  int64_t* pos_in_strand;      /* [cap]                                                         */
  int32_t* kmer;               /* [cap, k]   base codes (model.base_int); 255 = not ACGTN       */
  float* means;                /* [cap, k]   float32 cast of the float64 means (model feed)     */
  float* stds;                 /* [cap, k]                                                      */
  float* lens;                 /* [cap, k]   slice lengths as float32 (model.sanums)            */
  float* signals;              /* [cap, signal_len]                                             */
  double* means64;             /* [cap, k]   float64, for the feature-file text                 */
  double* stds64;              /* [cap, k]                                                      */
  int32_t* lens_i32;           /* [cap, k]                                                      */
  double* signals64;           /* [cap, signal_len]                                             */
  int32_t* window;             /* [cap, 3]   mode (0 short / 1 centre subsample / 2 window), start, n_real */
} dsb_sites;

/* All array pointers in `reads` and `sites` are DEVICE pointers.  n_sites (HOST pointer) gets the
 * number of sites found; if it exceeds sites->capacity the call returns DSB_ERR_CAPACITY after
 * writing the first `capacity` sites.  read_shift / read_scale (device, [n_reads], may be NULL)
 * receive the per-read normalisation constants.  Synchronises the stream once (site count). */
DSB_API int dsb_extract_dev(dsb_ctx* ctx, const dsb_reads* reads, const dsb_extract_params* params,
                    const dsb_sites* sites, int64_t* n_sites, double* read_shift, double* read_scale,
                    void* cuda_stream);
/* Same with HOST pointers everywhere (stages through device buffers owned by the ctx). */
DSB_API int dsb_extract_host(dsb_ctx* ctx, const dsb_reads* reads, const dsb_extract_params* params,
                     const dsb_sites* sites, int64_t* n_sites, double* read_shift, double* read_scale);
/* Upper bound on the number of sites of a read batch (= number of events), for sizing. */
DSB_API int64_t dsb_extract_max_sites(const dsb_reads* reads_host_offsets_only);

/* ---- fused end-to-end: extraction straight into the forward (fast5 -> call mode,
 *      call_modifications.py:94-122,264-278), features never leave HBM ------------------------ */
DSB_API int dsb_extract_call_dev(dsb_ctx* ctx, int32_t precision, const dsb_reads* reads,
                         const dsb_extract_params* params, const dsb_sites* sites, int64_t* n_sites,
                         float* probs, int32_t* pred, void* cuda_stream);

/* ---- site-frequency aggregation of call_mods results -------------------------------
 * Replaces the per-line dict update of the reference's scripts/call_modification_frequency.py:17-49 (ModRecord /
 * SiteStats of scripts/txt_formater.py): records are keyed by (chrom id, pos); for every key among the records with
 * use[i] != 0 it returns the index of the first record, the float64 sums of p0 / p1 accumulated in record order (bit-
 * identical to the reference's +=), the number of records with label[i] != 0 and the coverage.  HOST pointers in and out;
 * sites come in no particular order (sort by `first` for the reference's first-appearance order).  If more than
 * `capacity` sites exist, *n_sites holds the count and DSB_ERR_CAPACITY is returned. */
DSB_API int dsb_site_frequency_host(dsb_ctx* ctx, int64_t n, const int32_t* chrom_id, const int64_t* pos, const double* p0,
                                    const double* p1, const uint8_t* label, const uint8_t* use, int64_t capacity,
                                    int64_t* n_sites, int64_t* first, double* sum_p0, double* sum_p1, int64_t* met, int64_t* coverage);

/* ---- text codec around the path (host only) ---------------------------------------
 * dsb_tsv_parse_features: parse whole lines of a 12-column feature file (reference
 * extract_features.py:289-303, reader call_modifications.py:35-91) from buf[0:len) into the model-input
 * arrays; info_off[i] = byte range of the first six columns of sample i.  Returns the number of samples
 * (<= max_samples), *consumed = bytes of complete lines eaten; -(1000+i) if line i is malformed.
 * dsb_tsv_format_calls: render the 10-column call_mods lines (call_modifications.py:183-190) with
 * numpy-identical float32 text; returns bytes written. */
DSB_API int64_t dsb_tsv_parse_features(const char* buf, int64_t len, int32_t kmer_len, int32_t signal_len,
                                       int64_t max_samples, int32_t* kmer, float* means, float* stds, float* lens,
                                       float* signals, int32_t* labels, int64_t* info_off, int64_t* consumed,
                                       int32_t nthreads);
DSB_API int64_t dsb_tsv_format_calls(const char* info_buf, const int64_t* info_off, const int32_t* kmer,
                                     int32_t kmer_len, const float* probs, const int32_t* pred, int64_t n,
                                     char* out, int64_t out_cap);

#ifdef __cplusplus
}
#endif
#endif /* DSB200_H_ */
