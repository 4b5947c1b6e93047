/* threebz_cuda.h — C ABI of libthreebz_cuda.so, the B200 (sm_100a) inflate engine that sits
 * behind 3bz's Common Lisp API.
 *
 * The reference (3b/3bz, pure Common Lisp) has no FFI seam of its own: its boundary is the
 * exported Lisp API (package.lisp:13-27).  Each entry point below names the reference
 * function(s) it replaces; lisp/ holds the CFFI bindings and the re-hosted API, INTEGRATION.md
 * shows how a maintainer wires them in.
 *
 * Conventions
 *   - every function returns int32: 0 = ok, < 0 = engine failure (CUDA error, OOM, bad
 *     argument).  Problems with a *stream* are not failures: they are per-member verdicts.
 *   - plain pointers and sizes only; no exceptions, callbacks or longjmp cross the boundary.
 *   - a tbz_ctx belongs to one host thread at a time; distinct ctxs may run concurrently.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     TBZ_E_NO_DEVICE.
 */
#ifndef THREEBZ_CUDA_H
#define THREEBZ_CUDA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TBZ_ABI_VERSION 1

/* ---- status codes (return values) ---- */
enum {
  TBZ_OK = 0,
  TBZ_E_CUDA = -1,        /* a CUDA runtime call failed; see tbz_ctx_last_error */
  TBZ_E_NO_DEVICE = -2,   /* no usable CUDA device */
  TBZ_E_ARG = -3,         /* bad argument */
  TBZ_E_NOMEM = -4,       /* host or device allocation failed */
  TBZ_E_BUFFER_SWITCH = -5, /* replace-output-buffer refused (api.lisp:13-18) */
  TBZ_E_STATE = -6        /* session used after an error / after :done (gzip.lisp:279-286) */
};

/* ---- formats: decompress-vector's :format (api.lisp:31-34) ---- */
enum { TBZ_DEFLATE = 0, TBZ_ZLIB = 1, TBZ_GZIP = 2 };

/* ---- per-member verdicts.  0..2 are the three status flags of deflate-state
 * (deflate.lisp:53-62); >= 16 are the places where the reference signals a Lisp error. ---- */
enum {
  TBZ_FINISHED = 0,
  TBZ_INPUT_UNDERRUN = 1,
  TBZ_OUTPUT_OVERFLOW = 2,
  TBZ_ERR_BLOCK_TYPE = 16,      /* deflate.lisp:521-528 */
  TBZ_ERR_STORED_LEN = 17,      /* deflate.lisp:535 */
  TBZ_ERR_OVERSUBSCRIBED = 18,  /* huffman-tree.lisp:116-118 */
  TBZ_ERR_INCOMPLETE = 19,      /* huffman-tree.lisp:119-122 */
  TBZ_ERR_REPEAT_NO_PREV = 20,  /* deflate.lisp:642-643 */
  TBZ_ERR_REPEAT_OVERRUN = 21,  /* deflate.lisp:645,656 */
  TBZ_ERR_INVALID_SYMBOL = 22,  /* deflate.lisp:438,481,679 (ecase on an invalid node) */
  TBZ_ERR_DISTANCE_TOO_FAR = 23,/* deflate.lisp:343-345 */
  TBZ_ERR_ZLIB_FCHECK = 24,     /* zlib.lisp:20-22 */
  TBZ_ERR_ZLIB_METHOD = 25,     /* zlib.lisp:23-28 */
  TBZ_ERR_ZLIB_WINDOW = 26,     /* zlib.lisp:29-32 */
  TBZ_ERR_ZLIB_DICT = 27,       /* zlib.lisp:33-36 */
  TBZ_ERR_GZIP_MAGIC = 28,      /* gzip.lisp:120-121 */
  TBZ_ERR_GZIP_METHOD = 29,     /* gzip.lisp:130-132 */
  TBZ_ERR_GZIP_RESERVED = 30,   /* gzip.lisp:133-134 */
  TBZ_ERR_GZIP_HCRC = 31,       /* gzip.lisp:255 */
  TBZ_ERR_CHECKSUM = 32,        /* zlib.lisp:94, gzip.lisp:92 */
  TBZ_ERR_TREE_TOO_LARGE = 35   /* huffman-tree.lisp:208-216 (1444-node array overrun) */
};

/* where an INPUT_UNDERRUN happened: the gzip wrapper returns 0 instead of the output offset
 * when it runs dry inside its header or trailer (gzip.lisp:86,99,117) */
enum { TBZ_AT_HEADER = 0, TBZ_AT_BODY = 1, TBZ_AT_TRAILER = 2 };

/* ---- flags ---- */
enum {
  TBZ_FLAG_DEVICE_PTRS = 1u << 0, /* tbz_member.in/out are device pointers (no PCIe traffic) */
  TBZ_FLAG_NO_FASTPATH = 1u << 1, /* force the sequential kernel (testing / triage) */
  TBZ_FLAG_NO_SPLIT = 1u << 2     /* never use the speculative split decode for big members */
};

typedef struct tbz_ctx tbz_ctx;         /* one per (host thread, device) */
typedef struct tbz_batch tbz_batch;     /* a prepared, re-launchable batch */
typedef struct tbz_session tbz_session; /* backs one deflate-/zlib-/gzip-state */

typedef struct tbz_member {
  const uint8_t *in;  uint64_t in_len;   /* octet-vector-context / octet-pointer-context range */
  uint8_t *out;       uint64_t out_cap;  /* :output buffer */
} tbz_member;

typedef struct tbz_result {
  uint64_t out_len;    /* bytes produced (value `decompress` would return) */
  uint64_t in_used;    /* compressed bytes consumed when finished (end of trailer) */
  uint32_t checksum;   /* adler32 (zlib) / crc32 (gzip) of out[0,out_len); 0 for raw deflate */
  int32_t  verdict;    /* TBZ_FINISHED ... */
  uint32_t where;      /* TBZ_AT_* for TBZ_INPUT_UNDERRUN */
  uint32_t path;       /* which kernel produced it: 0 seq, 1 fast, 2 split (diagnostic) */
} tbz_result;

/* ---- library / context ---- */
int32_t tbz_abi_version(void);
int32_t tbz_device_count(int32_t *n);
int32_t tbz_ctx_create(int32_t device, uint64_t flags, tbz_ctx **ctx);
int32_t tbz_ctx_destroy(tbz_ctx *ctx);
const char *tbz_strerror(int32_t status);
const char *tbz_verdict_name(int32_t verdict);
const char *tbz_ctx_last_error(tbz_ctx *ctx);
int32_t tbz_ctx_synchronize(tbz_ctx *ctx);
/* the cudaStream_t all work of this ctx is enqueued on (for event timing by the caller) */
int32_t tbz_ctx_stream(tbz_ctx *ctx, void **cuda_stream);
/* CUDA-event stopwatch on that stream: start, enqueue work, stop -> elapsed device ms */
int32_t tbz_ctx_timer_start(tbz_ctx *ctx);
int32_t tbz_ctx_timer_stop(tbz_ctx *ctx, float *ms);
/* per-kernel CUDA-event times of the batched path: enable, launch, read the last launch's
 * {decode, resolve, sequential} milliseconds (measurement only; adds one event sync per launch) */
int32_t tbz_ctx_kernel_timing(tbz_ctx *ctx, int32_t enable);
int32_t tbz_ctx_last_kernel_ms(tbz_ctx *ctx, float *ms3);
/* number of engine kernels launched by this ctx so far */
int32_t tbz_ctx_launch_count(tbz_ctx *ctx, uint64_t *n);

/* ---- memory: io-mmap.lisp's octet-pointer ranges become pinned / registered host memory ---- */
int32_t tbz_host_alloc(uint64_t n, void **p);                     /* pinned */
int32_t tbz_host_free(void *p);
int32_t tbz_host_register(void *p, uint64_t n, uint32_t flags);   /* with-octet-pointer (io-mmap.lisp:26-40) */
int32_t tbz_host_unregister(void *p);
int32_t tbz_device_alloc(tbz_ctx *ctx, uint64_t n, void **p);
int32_t tbz_device_free(tbz_ctx *ctx, void *p);
int32_t tbz_memcpy_h2d(tbz_ctx *ctx, void *dst, const void *src, uint64_t n);
int32_t tbz_memcpy_d2h(tbz_ctx *ctx, void *dst, const void *src, uint64_t n);

/* ---- one-shot: decompress-vector with :output (api.lisp:23-48), many members at once.
 * Synchronous.  Caller memory only needs to stay valid during the call. ---- */
int32_t tbz_inflate_batch(tbz_ctx *ctx, int32_t format, const tbz_member *m, uint64_t n,
                          tbz_result *r, uint32_t flags, float *device_ms);
int32_t tbz_inflate_single(tbz_ctx *ctx, int32_t format, const uint8_t *in, uint64_t in_len,
                           uint8_t *out, uint64_t out_cap, tbz_result *r, uint32_t flags,
                           float *device_ms);
/* decompress-vector without :output (api.lisp:50-65): the engine sizes the result itself.
 * *out is malloc'ed; release it with tbz_free. */
int32_t tbz_inflate_alloc(tbz_ctx *ctx, int32_t format, const uint8_t *in, uint64_t in_len,
                          uint8_t **out, tbz_result *r);
void tbz_free(void *p);

/* prepared batch: upload once, launch many times (bench's device-resident timing loop) */
int32_t tbz_batch_prepare(tbz_ctx *ctx, int32_t format, const tbz_member *m, uint64_t n,
                          uint32_t flags, tbz_batch **b);
int32_t tbz_batch_launch(tbz_batch *b);                 /* asynchronous on the ctx stream */
int32_t tbz_batch_finish(tbz_batch *b, tbz_result *r);  /* waits, fetches results (+ outputs if host) */
int32_t tbz_batch_destroy(tbz_batch *b);

/* multi-GPU: partitions members host-side over ctxs[0..g) (longest first, least-loaded device),
 * one host thread per ctx, gathers results.  No collective, no NCCL. */
int32_t tbz_inflate_batch_multi(tbz_ctx *const *ctxs, int32_t g, int32_t format,
                                const tbz_member *m, uint64_t n, tbz_result *r,
                                uint32_t flags, float *device_ms_per_gpu);
/* the partition alone (host logic, no GPU needed): owner[i] in [0,g) */
int32_t tbz_partition(const uint64_t *in_len, uint64_t n, int32_t g, int32_t *owner);

/* ---- chunked API: decompress / replace-output-buffer / the status readers (api.lisp:3-21,67-72)
 * over a session that keeps the decoded member resident on the device. ---- */
int32_t tbz_session_create(tbz_ctx *ctx, int32_t format, tbz_session **s);
int32_t tbz_session_destroy(tbz_session *s);
/* (make-*-state :output-buffer out)  or  (setf ds-output-buffer) + (setf ds-output-offset 0) */
int32_t tbz_session_set_output(tbz_session *s, uint8_t *out, uint64_t cap);
/* same buffer, new address: a managed host (SBCL) pins its vectors only for the duration of one
 * foreign call, so the shim passes the current address before every decompress */
int32_t tbz_session_rebind_output(tbz_session *s, uint8_t *out);
/* replace-output-buffer: TBZ_E_BUFFER_SWITCH unless the old buffer is untouched or overflowed */
int32_t tbz_session_replace_output(tbz_session *s, uint8_t *out, uint64_t cap);
/* decompress: hands the session the unread octets [in, in+n) of the caller's context (the shim
 * then advances the context to its end), decodes as far as input and output allow.
 * *ret = the value the Lisp function returns; *verdict = TBZ_FINISHED / _INPUT_UNDERRUN /
 * _OUTPUT_OVERFLOW, or an error verdict where the Lisp would signal. */
int32_t tbz_session_decompress(tbz_session *s, const uint8_t *in, uint64_t n,
                               int64_t *ret, int32_t *verdict);
int32_t tbz_session_flags(tbz_session *s, int32_t *finished, int32_t *input_underrun,
                          int32_t *output_overflow);
/* How many of the octets handed over by the LAST tbz_session_decompress call the stream consumed: all of them
 * unless it finished inside them — the reference leaves the context's offset just past the consumed octets
 * (io.lisp:17-58; %resync-file-stream, io-common.lisp:60-63, seeks the stream there). */
int32_t tbz_session_consumed(tbz_session *s, uint64_t *n);

/* ---- gzip member metadata: the gzip-state slots flags / extra / name / comment / operating-system /
 * mtime/unix / compression-level that decompress-gzip fills while it reads the header
 * (gzip.lisp:17-28, :113-260).  Host-side only: no device work.  Offsets are from `in`. ---- */
enum { TBZ_GZ_TEXT = 1, TBZ_GZ_HCRC = 2, TBZ_GZ_EXTRA = 4, TBZ_GZ_NAME = 8, TBZ_GZ_COMMENT = 16 };
typedef struct tbz_gzip_header {
  int32_t  verdict;      /* TBZ_FINISHED = header complete; TBZ_INPUT_UNDERRUN; TBZ_ERR_GZIP_* */
  uint32_t flags;        /* TBZ_GZ_* (gzip.lisp:136-144) */
  uint32_t mtime;        /* mtime/unix; 0 = not set (gzip.lisp:154-157) */
  uint32_t xfl;          /* 2 = :maximum, 4 = :fastest (gzip.lisp:166-168) */
  uint32_t os;           /* 0..13 index the reference's keyword table, else (:unknown os) (gzip.lisp:169-176) */
  uint32_t header_crc;   /* the stored CRC16 when TBZ_GZ_HCRC */
  uint64_t extra_off, extra_len;     /* FEXTRA payload */
  uint64_t name_off, name_len;       /* FNAME without the terminating zero */
  uint64_t comment_off, comment_len; /* FCOMMENT without the terminating zero */
  uint64_t header_len;   /* the deflate body starts here */
} tbz_gzip_header;
int32_t tbz_gzip_header_parse(const uint8_t *in, uint64_t in_len, tbz_gzip_header *h);

/* Concatenated gzip members (RFC 1952 2.2).  The reference stops after the first member and reports
 * :done (gzip.lisp:279-286); that stays the default everywhere else.  This *new* entry point walks
 * them: member i is decoded to out + sum of the earlier out_len, r[i].in_used is its compressed size.
 * Stops at the first member that is not TBZ_FINISHED (its result is the last one written) or when
 * max_members are done.  *n_members = results written; *in_used = bytes of `in` consumed by finished members. */
int32_t tbz_inflate_gzip_members(tbz_ctx *ctx, const uint8_t *in, uint64_t in_len, uint8_t *out,
                                 uint64_t out_cap, tbz_result *r, uint64_t max_members,
                                 uint64_t *n_members, uint64_t *in_used);

#ifdef __cplusplus
}
#endif
#endif
