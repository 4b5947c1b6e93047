/* oracle3bz.h — CPU restatement of 3bz's inflate path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for the B200 engine: a plain-C restatement of the
 * algorithm in the reference's deflate.lisp / huffman-tree.lisp / zlib.lisp /
 * gzip.lisp / checksums.lisp / api.lisp.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libthreebz_cuda.so and the 3bz_b200 package) never links or calls it.
 *
 * Parity status: PINNED against the reference's own fixtures — test.deflated
 * (size + sha256), the 12 known-answer vectors and the 25 error-class vectors of
 * deflate-test.lisp — and cross-checked against system libz on valid streams.
 * The reference itself (Common Lisp) cannot run in this image; zlib/gzip wrapper
 * parity is pinned by restated rules, not by a reference fixture (none exists).
 */
#ifndef ORACLE3BZ_H
#define ORACLE3BZ_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { O3_DEFLATE = 0, O3_ZLIB = 1, O3_GZIP = 2 };

/* verdict numbering is shared with include/threebz_cuda.h */
enum {
  O3_FINISHED = 0,
  O3_INPUT_UNDERRUN = 1,
  O3_OUTPUT_OVERFLOW = 2,
  O3_ERR_BLOCK_TYPE = 16,     /* deflate.lisp:521-528 ecase on BTYPE 3 */
  O3_ERR_STORED_LEN = 17,     /* deflate.lisp:535 assert */
  O3_ERR_OVERSUBSCRIBED = 18, /* huffman-tree.lisp:116-118 */
  O3_ERR_INCOMPLETE = 19,     /* huffman-tree.lisp:119-122 */
  O3_ERR_REPEAT_NO_PREV = 20, /* deflate.lisp:642-643 */
  O3_ERR_REPEAT_OVERRUN = 21, /* deflate.lisp:645,656 */
  O3_ERR_INVALID_SYMBOL = 22, /* invalid node reached: ecase in deflate.lisp:481,438,679 */
  O3_ERR_DISTANCE_TOO_FAR = 23, /* deflate.lisp:343-345 "no window?" */
  O3_ERR_ZLIB_FCHECK = 24,    /* zlib.lisp:20-22 */
  O3_ERR_ZLIB_METHOD = 25,    /* zlib.lisp:23-28 */
  O3_ERR_ZLIB_WINDOW = 26,    /* zlib.lisp:29-32 */
  O3_ERR_ZLIB_DICT = 27,      /* zlib.lisp:33-36 */
  O3_ERR_GZIP_MAGIC = 28,     /* gzip.lisp:120-121 */
  O3_ERR_GZIP_METHOD = 29,    /* gzip.lisp:130-132 */
  O3_ERR_GZIP_RESERVED = 30,  /* gzip.lisp:133-134 */
  O3_ERR_GZIP_HCRC = 31,      /* gzip.lisp:255 */
  O3_ERR_CHECKSUM = 32,       /* zlib.lisp:94, gzip.lisp:92 */
  O3_ERR_BUFFER_SWITCH = 33,  /* api.lisp:13-18 */
  O3_ERR_STATE = 34,          /* calling decompress on a :done gzip state etc. */
  O3_ERR_TREE_TOO_LARGE = 35  /* node array (1444) overrun: bounds error in huffman-tree.lisp:208-216 */
};

typedef struct o3bz_context {   /* io-common.lisp:8-14,36-45 */
  const uint8_t *p;
  size_t start, end, offset;
} o3bz_context;

typedef struct o3bz_stats {
  uint64_t total_out;       /* U: bytes produced over the whole stream */
  uint64_t literals;        /* literal tokens */
  uint64_t matches;         /* length/distance tokens */
  uint64_t match_bytes;     /* B: sum of match lengths */
  uint64_t stored_bytes;    /* payload of BTYPE 0 blocks */
  uint64_t blocks[3];       /* per BTYPE */
  uint64_t header_bits;     /* dynamic header bits */
} o3bz_stats;

typedef struct o3bz_state o3bz_state;

o3bz_state *o3bz_state_new(int format);
void o3bz_state_free(o3bz_state *s);
/* (make-*-state :output-buffer buf) / (setf ds-output-buffer) + (setf ds-output-offset 0) */
void o3bz_set_output(o3bz_state *s, uint8_t *buf, size_t cap);
/* api.lisp:12-21; returns 0 or O3_ERR_BUFFER_SWITCH */
int o3bz_replace_output_buffer(o3bz_state *s, uint8_t *buf, size_t cap);
void o3bz_context_init(o3bz_context *c, const uint8_t *p, size_t start, size_t end);
/* api.lisp:3-10.  Returns what the Lisp `decompress` returns (an output offset, or 0 from
 * the gzip header/trailer underrun sites), or -1 when the reference would signal an error;
 * then o3bz_error() says which. */
int64_t o3bz_decompress(o3bz_context *c, o3bz_state *s);
int o3bz_finished(const o3bz_state *s);
int o3bz_input_underrun(const o3bz_state *s);
int o3bz_output_overflow(const o3bz_state *s);
int o3bz_error(const o3bz_state *s);
uint32_t o3bz_checksum(const o3bz_state *s); /* running adler32 (zlib) / crc32 (gzip) */
void o3bz_get_stats(const o3bz_state *s, o3bz_stats *st);

/* One-shot with a caller buffer: decompress-vector ... :output (api.lisp:36-48).
 * verdict = O3_FINISHED / O3_INPUT_UNDERRUN ("incomplete stream") / O3_OUTPUT_OVERFLOW
 * ("not enough space") / an error code.  *out_len = bytes produced. */
int o3bz_decompress_vector(const uint8_t *in, size_t start, size_t end, int format,
                           uint8_t *out, size_t out_cap, size_t *out_len,
                           uint32_t *checksum, o3bz_stats *st);
/* decompress-vector without :output (api.lisp:50-65): doubling buffers, concatenated.
 * Returns malloc'ed buffer in *out (free with o3bz_free). */
int o3bz_decompress_vector_grow(const uint8_t *in, size_t start, size_t end, int format,
                                uint8_t **out, size_t *out_len);
void o3bz_free(void *p);

/* checksums.lisp restated: running forms, same argument meaning as the Lisp */
void o3bz_adler32(const uint8_t *buf, size_t end, uint32_t *s1, uint32_t *s2);
uint32_t o3bz_crc32(const uint8_t *buf, size_t end, uint32_t crc);

/* batch helper for CPU baselines: members i in [lo,hi) ; returns number of non-finished */
int o3bz_batch(const uint8_t *const *in, const size_t *in_len, uint8_t *const *out,
               const size_t *out_cap, size_t *out_len, int *verdict, uint64_t *match_bytes,
               int format, size_t lo, size_t hi);

#ifdef __cplusplus
}
#endif
#endif
