/* defer_b200.h - C-ABI of libdefer_b200.so: the stage operator behind DEFER's node loop.
 *
 * Drop-in boundary.  The reference's compute node does, per stage,
 *     part = model_from_json(json); part.set_weights(ws)      (src/node.py:31,34)
 *     out  = part.predict(inpt)                                (src/node.py:105-106)
 *     socket_send(lz4(zfp(out)), next_node)                    (src/node.py:107-108)
 * and the dispatcher feeds / drains the chain (src/dispatcher.py:85-105).  This library replaces
 * exactly that: a *stage* is a fused-op plan + weights resident in HBM on one B200; its forward
 * pass is hand-written sm_100a kernels captured in a CUDA graph per in-flight lane; the hop is a
 * copy-engine transfer (or, optionally, the last kernel's own stores) of the stage output into the
 * next stage's input slot over NVLink (peer or CUDA-IPC mapped) followed by a release flag - no host
 * round trip, no codec (the reference codec is lossless, src/node.py:76-79, so a raw copy is
 * bit-equivalent).  Queue items stay single samples; a stage built with batch = G x item-batch runs
 * G in-flight items per launch (defer_stage_submit_part / _parts).
 *
 * Conventions: every entry point is extern "C", returns 0 on success and a negative defer_status
 * on failure; defer_last_error() gives a thread-local message.  No exception, torch type or C++
 * type crosses the boundary: plain pointers, sizes and POD structs only.  A stage handle is used
 * by one host thread at a time; different stages are independent.
 */
#ifndef DEFER_B200_H_
#define DEFER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEFER_ABI_VERSION 1

#if defined(DEFER_BUILD)
#define DEFER_API __attribute__((visibility("default")))
#else
#define DEFER_API
#endif

typedef enum defer_status {
  DEFER_OK = 0,
  DEFER_ERR_INVALID = -1,   /* bad argument / unsupported plan */
  DEFER_ERR_CUDA = -2,      /* CUDA runtime or driver error (message has the CUDA string) */
  DEFER_ERR_TIMEOUT = -3,   /* a device-side flag wait gave up (peer stage dead?) */
  DEFER_ERR_STATE = -4      /* call sequence error (e.g. step before link) */
} defer_status;

/* Activation storage format inside a stage and across hops. */
typedef enum defer_fmt {
  DEFER_FMT_F32 = 0,     /* IEEE fp32, SIMT FFMA contractions (exact-order fp32)                       */
  DEFER_FMT_BF16X2 = 1,  /* fp32 carried as two bf16 planes (hi + lo, 4 B/elem): tcgen05 bf16x3 MMAs,  */
                         /* fp32 accumulate - ~2^-16 relative per layer, meets the 1e-3 fp32 parity bar */
  DEFER_FMT_BF16 = 2     /* one bf16 plane, tcgen05 bf16 MMA, fp32 accumulate (the bf16 configs)       */
} defer_fmt;

typedef enum defer_op_kind {
  DEFER_OP_CONV = 1,     /* [zero-pad +] conv + per-channel scale/shift (bias+BN folded) [+ residual] [+ relu] */
  DEFER_OP_MAXPOOL = 2,  /* [zero-pad +] max-pool, 'valid'                                                    */
  DEFER_OP_GAP = 3,      /* global average pool  (H,W,C) -> (C)                                               */
  DEFER_OP_DENSE = 4,    /* x @ W + b [+ relu]   (softmax is its own op)                                      */
  DEFER_OP_SOFTMAX = 5,  /* row softmax over C                                                                */
  DEFER_OP_AFFINE = 6,   /* standalone BN: per-channel scale/shift [+ relu]                                   */
  DEFER_OP_RELU = 7,     /* standalone Activation('relu')                                                     */
  DEFER_OP_ADD = 8,      /* standalone Add of two tensors [+ relu]                                            */
  DEFER_OP_PAD = 9,      /* standalone ZeroPadding2D                                                          */
  DEFER_OP_COPY = 10     /* identity / Flatten / format cast                                                  */
} defer_op_kind;

#define DEFER_FLAG_RELU      1u   /* apply relu at the end of the op                 */
#define DEFER_FLAG_RESIDUAL  2u   /* CONV: add buffer in1 before the (optional) relu */

/* Buffer element type. */
#define DEFER_BUF_ACT 0   /* stage activation format (defer_fmt of the stage) */
#define DEFER_BUF_F32 1   /* plain fp32 regardless of the stage format (image in, probabilities out) */

/* One logical tensor of the plan.  Shapes are per sample, NHWC; vectors use h = w = 1. */
typedef struct defer_buf_desc {
  int32_t h, w, c;
  int32_t elem;            /* DEFER_BUF_ACT | DEFER_BUF_F32 */
} defer_buf_desc;

/* One fused op of the plan.  Buffer ids index the defer_buf_desc array; weight ids index the
 * weight-pointer array given to defer_stage_create (-1 = absent). */
typedef struct defer_op_desc {
  int32_t kind;            /* defer_op_kind */
  int32_t in0, in1, out;   /* buffer ids; in1 = residual / second addend, -1 if unused */
  int32_t kh, kw, sh, sw;  /* CONV / MAXPOOL window and stride */
  int32_t pad_t, pad_l, pad_b, pad_r; /* explicit zero padding applied to in0 (fused ZeroPadding2D / 'same') */
  uint32_t flags;          /* DEFER_FLAG_* */
  int32_t w_kernel;        /* CONV: fp32 HWIO kernel;  DENSE: fp32 (in,out) kernel */
  int32_t w_scale;         /* CONV / AFFINE: fp32 per-channel scale (NULL id -1 = ones) */
  int32_t w_shift;         /* CONV / AFFINE: fp32 per-channel shift;  DENSE: bias */
  int32_t reserved;
} defer_op_desc;

typedef struct defer_stage_config {
  int32_t abi_version;     /* DEFER_ABI_VERSION */
  int32_t device;          /* CUDA ordinal (reference: a node IP, src/dispatcher.py:45-55) */
  int32_t fmt;             /* defer_fmt */
  int32_t batch;           /* samples per microbatch (reference: 1, test/test.py:22) */
  int32_t depth;           /* in-flight microbatches = input slots = lanes (>= 1) */
  int32_t input_buf;       /* buffer id of the stage input  */
  int32_t output_buf;      /* buffer id of the stage output */
  int32_t is_first;        /* 1: input arrives from host via defer_stage_submit */
  int32_t is_last;         /* 1: output is read back by defer_stage_result */
  int32_t conv_backend;    /* 0 auto, 1 SIMT only, 2 tcgen05 where eligible (error if fmt == F32) */
  int32_t use_graph;       /* 1: capture each lane's chain into a CUDA graph (default), 0: eager launches */
  int32_t wait_timeout_ms; /* device-side flag wait budget; 0 = default (4000 ms) */
} defer_stage_config;

typedef struct defer_stage_s* defer_stage_t;

/* Size in bytes of the opaque link token produced by defer_stage_export_link. */
#define DEFER_LINK_TOKEN_BYTES 256

/* ---- library ------------------------------------------------------------------------------ */
DEFER_API const char* defer_last_error(void);
DEFER_API int defer_abi_version(void);
DEFER_API int defer_device_count(int* count);
/* name (<=255 chars), SM count, compute capability major*10+minor, total HBM bytes */
DEFER_API int defer_device_info(int device, char* name, int name_len, int* sm_count, int* cc, uint64_t* hbm_bytes);

/* ---- stage life cycle  (replaces model_from_json + set_weights, src/node.py:31-38) ---------- */
/* weights: host pointers to fp32 arrays, copied; the caller keeps ownership of host memory. */
DEFER_API int defer_stage_create(const defer_stage_config* cfg,
                       const defer_buf_desc* bufs, int n_bufs,
                       const defer_op_desc* ops, int n_ops,
                       const void* const* weight_ptrs, const uint64_t* weight_nbytes, int n_weights,
                       defer_stage_t* out);
DEFER_API int defer_stage_destroy(defer_stage_t s);
/* Human-readable plan / kernel choice dump (replaces plot_model, src/node.py:39). */
DEFER_API int defer_stage_describe(defer_stage_t s, char* buf, size_t buf_len);
/* bytes of one microbatch entering / leaving the stage (hop payload) */
DEFER_API int defer_stage_io_bytes(defer_stage_t s, uint64_t* in_bytes, uint64_t* out_bytes);

/* ---- wiring the chain  (replaces next-hop hand-off, src/dispatcher.py:51-55,63) ------------- */
/* Same process: enable peer access both ways and hand prod the consumer's slots + flags. */
DEFER_API int defer_stage_link(defer_stage_t prod, defer_stage_t cons);
/* Other process (one rank per GPU): the consumer exports a token, the producer imports it, and
 * vice versa for the back-pressure flags.  role: 0 = "my input side" (give to my producer),
 * 1 = "my output side" (give to my consumer). */
DEFER_API int defer_stage_export_link(defer_stage_t s, int role, void* token /* DEFER_LINK_TOKEN_BYTES */);
DEFER_API int defer_stage_import_link(defer_stage_t s, int role, const void* token);
/* Drop the mappings of the neighbours' arenas (CUDA-IPC close).  In a multi-process pipeline every rank
 * calls this, then all ranks synchronise, then each destroys its stage - so no arena is freed while a
 * neighbour still maps it.  The stage cannot step afterwards. */
DEFER_API int defer_stage_unlink(defer_stage_t s);
/* Finish wiring: builds the per-lane CUDA graphs.  Must be called once after linking (also for
 * a single-stage pipeline). */
DEFER_API int defer_stage_finalize(defer_stage_t s);

/* ---- steady state  (replaces the recv -> predict -> send loop, src/node.py:88-91,103-108) --- */
/* First stage only: enqueue the H2D copy of microbatch `seq` from host memory (pinned => async). */
DEFER_API int defer_stage_submit(defer_stage_t s, uint64_t seq, const void* host_in, uint64_t nbytes);
/* First stage only, coalesced ingress: the reference's queue items are single samples (test/test.py:22,47-49); a stage
 * built with batch = G x item-batch runs G in-flight items per launch.  Copies `count` consecutive samples of
 * microbatch `seq`, starting at sample `index`, from host memory (nbytes = count x bytes of one sample). */
DEFER_API int defer_stage_submit_part(defer_stage_t s, uint64_t seq, int index, int count, const void* host_in, uint64_t nbytes);
/* The same for a run of queue items in ONE call (the dispatcher's feeder gathers the items of a group, then ships them):
 * item i (samples_per_item samples, nbytes_per_item bytes at host_ptrs[i]) goes to samples
 * [first_index + i * samples_per_item, ...). */
DEFER_API int defer_stage_submit_parts(defer_stage_t s, uint64_t seq, int first_index, int n_items, int samples_per_item,
                             const void* const* host_ptrs, uint64_t nbytes_per_item);
/* Enqueue microbatch `seq` on lane seq % depth: wait-input -> kernel chain -> hop -> flags. Async. */
DEFER_API int defer_stage_step(defer_stage_t s, uint64_t seq);
/* Last stage only: block until microbatch `seq` is complete and copy its fp32 output to host. */
DEFER_API int defer_stage_result(defer_stage_t s, uint64_t seq, void* host_out, uint64_t nbytes);
/* Convenience for single-stage use: submit + step + result. */
DEFER_API int defer_stage_predict(defer_stage_t s, const void* host_in, uint64_t in_bytes, void* host_out, uint64_t out_bytes);
DEFER_API int defer_stage_sync(defer_stage_t s);
/* Sticky device-side status: 0 ok, DEFER_ERR_TIMEOUT if a flag wait expired. */
DEFER_API int defer_stage_status(defer_stage_t s);
/* Device time of the last completed step on `lane` in microseconds (CUDA events on the lane's stream). */
DEFER_API int defer_stage_last_step_us(defer_stage_t s, int lane, float* us);

/* Whole-job device timing across all lanes (CUDA events, no host clock): start records T0 on lane 0's
 * stream (call it when the stage is idle, right before the first step of the timed region); stop makes
 * lane 0 wait for every lane, records T1, synchronises and returns T1 - T0 in milliseconds. */
DEFER_API int defer_stage_timer_start(defer_stage_t s);
DEFER_API int defer_stage_timer_stop(defer_stage_t s, float* ms);

/* Steady-state timing without draining the pipeline (the reference protocol counts results inside a window while the
 * chain stays flooded, test/test.py:25-36): call mark(seq, slot) right after step(seq); it records a CUDA event behind
 * that microbatch on its lane.  mark_elapsed = device time from the completion of the slot-0 microbatch to the
 * completion of the slot-1 microbatch. */
DEFER_API int defer_stage_mark(defer_stage_t s, uint64_t seq, int slot /* 0 | 1 */);
DEFER_API int defer_stage_mark_elapsed(defer_stage_t s, float* ms);

/* ---- introspection for tests and benches --------------------------------------------------- */
DEFER_API int defer_stage_num_kernels(defer_stage_t s, int* per_step);           /* kernel launches per step */
/* copy any plan buffer of `lane` to host as fp32 NHWC (decodes the stage format) */
DEFER_API int defer_stage_read_buffer(defer_stage_t s, int lane, int buf_id, float* host_out, uint64_t n_floats);
/* stream / event handles for external timing (cudaStream_t as void*) */
DEFER_API int defer_stage_stream(defer_stage_t s, int lane, void** stream);
/* time `iters` back-to-back launches of op `op_index` alone on lane 0 (CUDA events); microseconds per launch */
DEFER_API int defer_stage_time_op(defer_stage_t s, int op_index, int iters, int flush_l2, float* us_per_launch);
/* algorithmic bytes and flops of op `op_index` (SURVEY.md 8d formula) and its kernel name */
DEFER_API int defer_stage_op_info(defer_stage_t s, int op_index, double* alg_bytes, double* alg_flops,
                        char* kernel_name, int name_len);

/* ---- host memory helpers -------------------------------------------------------------------- */
DEFER_API int defer_host_alloc(void** ptr, uint64_t nbytes);     /* pinned */
DEFER_API int defer_host_free(void* ptr);
DEFER_API int defer_host_register(void* ptr, uint64_t nbytes);   /* page-lock caller memory in place */
DEFER_API int defer_host_unregister(void* ptr);

/* ---- per-kernel entry points (raw device pointers, e.g. torch.Tensor.data_ptr(); NHWC) ------ */
/* Each runs ONE kernel on `stream` (cudaStream_t as void*, NULL = default) so it can be parity-
 * tested and ncu-profiled alone.  Activation tensors are in `fmt`; planes of BF16X2 are
 * [hi | lo], lo at element offset n*h*w*c. Weights: fp32 HWIO + fp32 scale/shift (may be NULL). */
/* backend: 1 SIMT FFMA | 2 tcgen05, one tile per CTA (conv_umma_kernel) | 3 round-1 persistent grid (conv_mega_kernel) |
 *          4 / 5 streaming persistent kernel (conv_stream_kernel) with 64- / 128-wide N tiles |
 *          6 / 7 the same with the per-thread (peer-memory capable) epilogue */
DEFER_API int defer_k_conv(int fmt, int backend,
                 const void* x, int x_is_f32, const float* w_hwio, const float* scale, const float* shift,
                 const void* residual, void* y,
                 int n, int h, int w, int cin, int cout, int kh, int kw, int sh, int sw,
                 int pad_t, int pad_l, int pad_b, int pad_r, uint32_t flags, void* stream);
DEFER_API int defer_k_maxpool(int fmt, const void* x, void* y, int n, int h, int w, int c,
                    int ph, int pw, int sh, int sw, int pad_t, int pad_l, int pad_b, int pad_r, void* stream);
DEFER_API int defer_k_gap(int fmt, const void* x, void* y, int n, int h, int w, int c, void* stream);
DEFER_API int defer_k_dense(int fmt, const void* x, const float* w_io, const float* bias, void* y, int y_is_f32,
                  int n, int in_features, int units, uint32_t flags, void* stream);
DEFER_API int defer_k_softmax(const float* x, float* y, int n, int c, void* stream);
DEFER_API int defer_k_eltwise(int fmt, int kind /* AFFINE | RELU | ADD */, const void* a, const void* b,
                    const float* scale, const float* shift, void* y, int n, int h, int w, int c,
                    uint32_t flags, void* stream);
/* fp32 <-> stage-format conversion of a whole tensor (device pointers) */
DEFER_API int defer_k_encode(int fmt, const float* x_f32, void* y_act, uint64_t n_elems, void* stream);
DEFER_API int defer_k_decode(int fmt, const void* x_act, float* y_f32, uint64_t n_elems, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEFER_B200_H_ */
