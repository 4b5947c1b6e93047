/* oracle3bz.c — CPU restatement of 3bz's inflate path.  TEST INFRASTRUCTURE ONLY
 * (see oracle3bz.h for who may use it and for the parity-pinning statement).
 *
 * Written from the behaviour of the reference, not transliterated from it.  Citations are
 * file:line into /root/reference.
 */
#include "oracle3bz.h"
#include <stdlib.h>
#include <string.h>

/* ---- constants.lisp ------------------------------------------------------------------ */
#define MAX_TREE 1444              /* constants.lisp:4-7  (852 + 592) */
#define HT_MAX_BITS 28             /* constants.lisp:11 */
#define WINDOW_SIZE 32768          /* deflate.lisp:122 */
enum { T_LIT = 0, T_LINK = 1, T_LENDIST = 2, T_INVALID = 3 }; /* constants.lisp:14-17 */
#define NODE_INVALID 0xffffu       /* huffman-tree.lisp:74 */
#define NODE_END 0x0001u           /* huffman-tree.lisp:75 */
#define LEN_OFFSET 32              /* constants.lisp:26 */

/* constants.lisp:36-43: distance extra bits at 0..29, length extra bits at 32..60 */
static const uint8_t k_extra_bits[61] = {
  0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13, 0,0,
  0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0 };
/* constants.lisp:48-61 */
static const uint16_t k_bases[61] = {
  1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,
  8193,12289,16385,24577, 0,0,
  3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258 };
/* constants.lisp:65-68 */
static const uint8_t k_len_code_order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
/* constants.lisp:70-73 */
static const uint8_t k_len_code_extra[19] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,3,7};

enum { TREE_LITLEN, TREE_DIST, TREE_DHTLEN };

/* huffman-tree.lisp:78-84 */
typedef struct {
  int start_bits, max_bits;
  int empty;                       /* all lengths zero: the Lisp leaves a stale table (huffman-tree.lisp:156-157);
                                      in a fresh state that table is all-invalid, which is what we model */
  uint16_t nodes[MAX_TREE];
} tree;

static void tree_init(tree *t) {
  t->start_bits = 0; t->max_bits = 0; t->empty = 1;
  for (int i = 0; i < MAX_TREE; i++) t->nodes[i] = NODE_INVALID;
}

static unsigned bit_rev(unsigned v, int bits) {  /* util.lisp:59-69 */
  unsigned r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (v & 1); v >>= 1; }
  return r;
}

typedef struct {
  tree *t;
  uint16_t term[320];
  int counts[16], offsets[16];
  int next_subtable;
  int overflow;
} builder;

static int next_len(const builder *b, int l) {
  for (int i = l; i < 16; i++) if (b->counts[i] > 0) return i;
  return -1;
}

/* huffman-tree.lisp:190-212: one entry of a (sub)table whose prefix is prefix_bits long */
static uint16_t fill_entry(builder *b, int prefix_bits) {
  int entry_bits = b->counts[prefix_bits] ? prefix_bits : next_len(b, prefix_bits);
  if (entry_bits < 0) return NODE_INVALID;
  if (entry_bits == prefix_bits) {
    uint16_t n = b->term[b->offsets[entry_bits]++];
    b->counts[entry_bits]--;
    return n;
  }
  int start = b->next_subtable, nb = entry_bits - prefix_bits;
  if (start + (1 << nb) > MAX_TREE) { b->overflow = 1; return NODE_INVALID; }
  b->next_subtable += 1 << nb;
  for (unsigned i = 0; i < (1u << nb); i++)
    b->t->nodes[start + bit_rev(i, nb)] = fill_entry(b, entry_bits);
  return (uint16_t)(T_LINK | (nb << 2) | (start << 6));   /* huffman-tree.lisp:26-29 */
}

/* huffman-tree.lisp:99-218.  Returns 0 or an O3_ERR_* code. */
static int build_tree_part(tree *t, const uint8_t *table, int type, int start, int end,
                           const uint8_t *extra_bits) {
  builder b;
  int counts[16];
  memset(counts, 0, sizeof counts);
  for (int x = start; x < end; x++) counts[table[x]]++;
  /* Kraft check, huffman-tree.lisp:112-122 */
  long s = 1;
  for (int i = 1; i < 16; i++) {
    s <<= 1;
    if (counts[i] > s) return O3_ERR_OVERSUBSCRIBED;
    s -= counts[i];
  }
  if (s > 0 && (end - start) - counts[0] > 1) return O3_ERR_INCOMPLETE;
  counts[0] = 0;
  int c = 0, min = -1, max = 0;
  for (int i = 0; i < 16; i++) {
    b.offsets[i] = counts[i] ? c : 0;
    c += counts[i];
    b.counts[i] = counts[i];
    if (counts[i]) { if (min < 0) min = i; max = i; }
  }
  t->max_bits = max + (type == TREE_DIST ? 13 : type == TREE_LITLEN ? 5 : 7);
  if (min < 0) {                   /* huffman-tree.lisp:156-157 */
    t->start_bits = 0; t->max_bits = 0; t->empty = 1;
    return 0;
  }
  /* sort symbols by (length, symbol) into terminals, huffman-tree.lisp:159-183 */
  int off_tmp[16];
  memcpy(off_tmp, b.offsets, sizeof off_tmp);
  for (int i = 0, to = start; to < end; i++, to++) {
    int l = table[to];
    if (!l) continue;
    int o = off_tmp[l]++;
    uint16_t n;
    if (type != TREE_LITLEN)
      n = i <= 29 ? (uint16_t)(T_LENDIST | (i << 6) | (extra_bits[i] << 2)) : NODE_INVALID;
    else if (i > 285) n = NODE_INVALID;
    else if (i >= 257) {
      int v = LEN_OFFSET + (i - 257);
      n = (uint16_t)(T_LENDIST | (v << 6) | (extra_bits[v] << 2));
    } else if (i == 256) n = NODE_END;
    else n = (uint16_t)(T_LIT | (i << 6));
    b.term[o] = n;
  }
  b.t = t;
  b.next_subtable = 1 << min;      /* huffman-tree.lisp:213-217 */
  /* The node array has 1444 slots (constants.lisp:4-7).  A complete code over <= 288 symbols
   * always fits; a lone symbol of length >= 11 needs a 2^len root and the Lisp's bounds-checked
   * (setf aref) signals an error. */
  if (b.next_subtable > MAX_TREE) return O3_ERR_TREE_TOO_LARGE;
  b.overflow = 0;
  for (unsigned i = 0; i < (1u << min); i++)
    t->nodes[bit_rev(i, min)] = fill_entry(&b, min);
  if (b.overflow) return O3_ERR_TREE_TOO_LARGE;
  t->start_bits = min;
  t->empty = 0;
  return 0;
}

/* ht-constants.lisp:9-32 + huffman-tree.lisp:89-97 */
static tree g_static_lit, g_static_dist;
static int g_static_ready = 0;
static void static_trees(void) {
  if (g_static_ready) return;
  uint8_t lit[288], dist[32];
  for (int i = 0; i < 288; i++) lit[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
  for (int i = 0; i < 32; i++) dist[i] = 5;
  tree_init(&g_static_lit); tree_init(&g_static_dist);
  build_tree_part(&g_static_lit, lit, TREE_LITLEN, 0, 288, k_extra_bits);
  build_tree_part(&g_static_dist, dist, TREE_DIST, 0, 32, k_extra_bits);
  g_static_ready = 1;
}

/* ---- checksums.lisp ------------------------------------------------------------------ */
void o3bz_adler32(const uint8_t *buf, size_t end, uint32_t *ps1, uint32_t *ps2) {
  /* checksums.lisp:18-62: 64-bit accumulators, reduced before they can wrap */
  uint64_t s1 = *ps1, s2 = *ps2;
  size_t i = 0;
  while (i < end) {
    size_t n = end - i;
    if (n > 380368439u) n = 380368439u - (380368439u % 32);
    for (size_t e = i + n; i < e; i++) { s1 += buf[i]; s2 += s1; }
    s1 %= 65521; s2 %= 65521;
  }
  *ps1 = (uint32_t)s1; *ps2 = (uint32_t)s2;
}

static uint32_t g_crc_table[256];
static int g_crc_ready = 0;
static void crc_table(void) {      /* checksums.lisp:177-189 */
  if (g_crc_ready) return;
  for (uint32_t n = 0; n < 256; n++) {
    uint32_t c = n;
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
    g_crc_table[n] = c;
  }
  g_crc_ready = 1;
}
uint32_t o3bz_crc32(const uint8_t *buf, size_t end, uint32_t crc) { /* checksums.lisp:196-210 */
  crc_table();
  crc ^= 0xffffffffu;
  for (size_t i = 0; i < end; i++) crc = (crc >> 8) ^ g_crc_table[(crc ^ buf[i]) & 0xff];
  return crc ^ 0xffffffffu;
}

/* ---- deflate.lisp -------------------------------------------------------------------- */
enum { /* state tags, deflate.lisp:517-726 */
  ST_START_OF_BLOCK, ST_UNCOMPRESSED_BLOCK, ST_COPY_BLOCK, ST_DYNAMIC_HUFFMAN_BLOCK,
  ST_DHT_LEN_TABLE, ST_DHT_LEN_TABLE_DATA, ST_DECODE_COMPRESSED_DATA, ST_CONTINUE_COPY_HISTORY,
  ST_OUT_BYTE, ST_BLOCK_END, ST_DONE };
enum { R_OK = 0, R_EOI = 1, R_EOO = 2, R_ERR = 3 };
enum { ZS_NONE, ZS_HEADER, ZS_HEADER2, ZS_ADLER };
enum { GS_HEADER, GS_HEADER2, GS_MTIME, GS_HEADER3, GS_EXTRA, GS_NAME, GS_COMMENT, GS_HCRC,
       GS_DEFLATE, GS_FINAL_CRC, GS_FINAL_LEN, GS_DONE };

struct o3bz_state {                /* deflate.lisp:4-62, zlib.lisp:3-12, gzip.lisp:3-28 */
  int format;
  int cur;
  int last_block;
  tree dyn_lit, dyn_dist, len_tree;
  const tree *lit, *dist;
  int hlit, hlit_hdist, hclen;
  uint8_t len_codes[19];
  uint8_t lld[320];
  int lld_index, last_len;
  uint32_t bytes_to_copy, copy_offset;
  uint64_t bits; int nbits;        /* partial-bits / bits-remaining */
  size_t out_off, out_cap; uint8_t *out;
  uint8_t *window;                 /* 32768 + 8, created at the first overflow */
  uint64_t base_total;             /* bytes produced into earlier output buffers */
  int finished, overflow, underrun, error;
  /* zlib */
  int zstate; uint32_t s1, s2;
  /* gzip */
  int gstate, gflags, have_extra; uint32_t extra_len, extra_got;
  uint8_t *hdr; size_t hdr_len, hdr_cap; int hdr_keep;
  uint32_t crc, isize;
  o3bz_stats st;
};

static int need(o3bz_state *s, o3bz_context *c, int n) {
  /* word64/word32 refills (io.lisp:17-58) restated bytewise: same bits, same order */
  if (s->nbits < n && c->end - c->offset >= 8) {   /* word64, io.lisp:17-37 */
    uint64_t w;
    int take = (63 - s->nbits) >> 3;
    memcpy(&w, c->p + c->offset, 8);
    s->bits |= (w & ((1ull << (8 * take)) - 1)) << s->nbits;
    s->nbits += 8 * take;
    c->offset += (size_t)take;
  }
  while (s->nbits < n && s->nbits <= 56 && c->offset < c->end) {
    s->bits |= (uint64_t)c->p[c->offset++] << s->nbits;
    s->nbits += 8;
  }
  return s->nbits >= n;
}
static inline uint32_t take(o3bz_state *s, int n) {      /* %bits, deflate.lisp:180-183 */
  uint32_t v = (uint32_t)(s->bits & ((n >= 32) ? 0xffffffffu : ((1u << n) - 1)));
  s->bits >>= n; s->nbits -= n;
  return v;
}
static inline void byte_align(o3bz_state *s) {           /* deflate.lisp:142-146 */
  int r = s->nbits & 7;
  if (r) { s->bits >>= r; s->nbits -= r; }
}

/* decode-huffman / decode-huffman-full (deflate.lisp:361-509) as one peeking walk:
 * nothing is consumed unless the whole symbol (code + extra bits) is available. */
static inline int walk(const tree *t, uint64_t bits, int avail, unsigned *node_out,
                       unsigned *extra_out, int *used_out) {
  if (t->empty) return R_ERR;
  int hb = t->start_bits, off = 0;
  unsigned node = 0, extra = 0;
  for (;;) {
    if (hb > avail - off) return R_EOI;
    node = t->nodes[node + ((bits >> off) & ((1u << hb) - 1))];
    off += hb;
    unsigned ty = node & 3;
    if (ty == T_LINK) {
      if (node == NODE_END) break;
      hb = (node >> 2) & 15; node >>= 6;
    } else if (ty == T_LIT) break;
    else if (ty == T_LENDIST) {
      int x = (node >> 2) & 15;
      if (x) {
        if (x > avail - off) return R_EOI;
        extra = (unsigned)((bits >> off) & ((1u << x) - 1));
        off += x;
      }
      break;
    } else return R_ERR;
  }
  *node_out = node; *extra_out = extra; *used_out = off;
  return R_OK;
}

/* eoo, deflate.lisp:121-137 */
static int eoo(o3bz_state *s) {
  if (!s->window) s->window = (uint8_t *)calloc(WINDOW_SIZE + 8, 1);
  if (s->out_off < WINDOW_SIZE)
    memmove(s->window, s->window + s->out_off, WINDOW_SIZE + 8 - s->out_off);
  size_t s1 = s->out_off < WINDOW_SIZE ? WINDOW_SIZE - s->out_off : 0;
  size_t s2 = s->out_off > WINDOW_SIZE ? s->out_off - WINDOW_SIZE : 0;
  size_t n = s->out_cap - s2;
  if (n > WINDOW_SIZE + 8 - s1) n = WINDOW_SIZE + 8 - s1;
  if (n) memcpy(s->window + s1, s->out + s2, n);
  s->overflow = 1;
  return R_EOO;
}

/* copy-history / %copy-history, deflate.lisp:244-359 — bytewise semantics; the Lisp's
 * word-sized fast paths give the same first `count` bytes. */
static int copy_history(o3bz_state *s, uint32_t count, uint32_t offset) {
  size_t d = s->out_off, e = s->out_cap;
  if ((uint64_t)offset > s->base_total + d) {   /* deflate.lisp:343-345 (see DESIGN: once a
                                                    window exists the Lisp read is unchecked) */
    s->error = O3_ERR_DISTANCE_TOO_FAR; return R_ERR;
  }
  uint32_t n = count;
  while (n && d < e) {
    long si = (long)d - (long)offset;
    s->out[d] = si < 0 ? s->window[WINDOW_SIZE + si] : s->out[si];
    d++; n--;
  }
  s->out_off = d;
  if (n) {                                       /* deflate.lisp:263-269 */
    s->bytes_to_copy = n; s->copy_offset = offset;
    s->cur = ST_CONTINUE_COPY_HISTORY;
    return eoo(s);
  }
  return R_OK;
}

static int decompress_deflate(o3bz_context *c, o3bz_state *s) {
  s->overflow = 0; s->underrun = 0;              /* deflate.lisp:102-103 */
  for (;;) switch (s->cur) {
  case ST_START_OF_BLOCK: {                      /* deflate.lisp:518-528 */
    if (!need(s, c, 3)) goto eoi;
    uint32_t b = take(s, 3);
    s->last_block = b & 1;
    switch (b >> 1) {
    case 0: s->cur = ST_UNCOMPRESSED_BLOCK; s->st.blocks[0]++; break;
    case 1: s->lit = &g_static_lit; s->dist = &g_static_dist;
            s->cur = ST_DECODE_COMPRESSED_DATA; s->st.blocks[1]++; break;
    case 2: s->lit = &s->dyn_lit; s->dist = &s->dyn_dist;
            s->cur = ST_DYNAMIC_HUFFMAN_BLOCK; s->st.blocks[2]++; break;
    default: s->error = O3_ERR_BLOCK_TYPE; return R_ERR;
    }
    break; }
  case ST_UNCOMPRESSED_BLOCK: {                  /* deflate.lisp:532-537 */
    byte_align(s);
    if (!need(s, c, 32)) goto eoi;
    uint32_t v = take(s, 32);
    if ((v >> 16) != ((~v) & 0xffff)) { s->error = O3_ERR_STORED_LEN; return R_ERR; }
    s->bytes_to_copy = v & 0xffff;
    s->cur = ST_COPY_BLOCK;
    break; }
  case ST_COPY_BLOCK:                            /* deflate.lisp:538-573 */
    while (s->bytes_to_copy) {
      if (s->out_off >= s->out_cap) return eoo(s);
      if (!need(s, c, 8)) goto eoi;
      s->out[s->out_off++] = (uint8_t)take(s, 8);
      s->bytes_to_copy--; s->st.stored_bytes++;
    }
    s->cur = ST_BLOCK_END;
    break;
  case ST_DYNAMIC_HUFFMAN_BLOCK: {               /* deflate.lisp:577-595 */
    if (!need(s, c, 26)) goto eoi;
    uint32_t v = take(s, 26);
    memset(s->len_codes, 0, sizeof s->len_codes);
    s->len_codes[16] = (v >> 14) & 7; s->len_codes[17] = (v >> 17) & 7;
    s->len_codes[18] = (v >> 20) & 7; s->len_codes[0] = (v >> 23) & 7;
    s->hlit = (v & 31) + 257;
    s->hlit_hdist = s->hlit + ((v >> 5) & 31) + 1;
    s->hclen = (v >> 10) & 15;
    s->lld_index = 0;
    s->st.header_bits += 26;
    s->cur = ST_DHT_LEN_TABLE;
    break; }
  case ST_DHT_LEN_TABLE: {                       /* deflate.lisp:597-624 */
    int n = s->hclen * 3;
    if (!need(s, c, n)) goto eoi;
    uint64_t v = s->bits; s->bits >>= n; s->nbits -= n;
    for (int i = 0; i < s->hclen; i++) s->len_codes[k_len_code_order[4 + i]] = (v >> (3 * i)) & 7;
    s->st.header_bits += n;
    int err = build_tree_part(&s->len_tree, s->len_codes, TREE_DHTLEN, 0, 19, k_len_code_extra);
    if (err) { s->error = err; return R_ERR; }
    s->last_len = 0xff;
    s->cur = ST_DHT_LEN_TABLE_DATA;
    break; }
  case ST_DHT_LEN_TABLE_DATA: {                  /* deflate.lisp:626-669 */
    while (s->lld_index < s->hlit_hdist) {
      unsigned node, extra; int used;
      need(s, c, HT_MAX_BITS);
      int r = walk(&s->len_tree, s->bits, s->nbits, &node, &extra, &used);
      if (r == R_EOI) goto eoi;
      if (r == R_ERR) { s->error = O3_ERR_INVALID_SYMBOL; return R_ERR; }
      s->bits >>= used; s->nbits -= used; s->st.header_bits += used;
      int code = node >> 6;
      if (code < 16) { s->lld[s->lld_index++] = (uint8_t)code; s->last_len = code; }
      else if (code == 16) {
        if (s->last_len >= 16) { s->error = O3_ERR_REPEAT_NO_PREV; return R_ERR; }
        int e = s->lld_index + extra + 3;
        if (e > s->hlit_hdist) { s->error = O3_ERR_REPEAT_OVERRUN; return R_ERR; }
        while (s->lld_index < e) s->lld[s->lld_index++] = (uint8_t)s->last_len;
      } else {
        int e = s->lld_index + extra + (code == 17 ? 3 : 11);
        if (e > s->hlit_hdist) { s->error = O3_ERR_REPEAT_OVERRUN; return R_ERR; }
        while (s->lld_index < e) s->lld[s->lld_index++] = 0;
        s->last_len = 0;
      }
    }
    /* build-trees*, huffman-tree.lisp:272-287 */
    int err = build_tree_part(&s->dyn_lit, s->lld, TREE_LITLEN, 0, s->hlit, k_extra_bits);
    if (!err) err = build_tree_part(&s->dyn_dist, s->lld, TREE_DIST, s->hlit, s->lld_index, k_extra_bits);
    if (err) { s->error = err; return R_ERR; }
    s->cur = ST_DECODE_COMPRESSED_DATA;
    break; }
  case ST_DECODE_COMPRESSED_DATA:                /* deflate.lisp:673-702 */
    for (;;) {
      unsigned node, extra, dnode, dextra; int used, dused;
      if (s->nbits < 48) need(s, c, 48);
      int r = walk(s->lit, s->bits, s->nbits, &node, &extra, &used);
      if (r == R_EOI) goto eoi;
      if (r == R_ERR) { s->error = O3_ERR_INVALID_SYMBOL; return R_ERR; }
      unsigned ty = node & 3;
      if (ty == T_LIT) {
        s->bits >>= used; s->nbits -= used;
        if (s->out_off >= s->out_cap) {          /* deflate.lisp:693-697 */
          s->cur = ST_OUT_BYTE; s->bytes_to_copy = node >> 6;
          return eoo(s);
        }
        s->out[s->out_off++] = (uint8_t)(node >> 6);
        s->st.literals++;
      } else if (ty == T_LENDIST) {
        uint32_t octets = extra + k_bases[node >> 6];
        /* a failed distance read leaves the length bits unread, deflate.lisp:411-426 */
        r = walk(s->dist, s->bits >> used, s->nbits - used, &dnode, &dextra, &dused);
        if (r == R_EOI) goto eoi;
        if (r == R_ERR) { s->error = O3_ERR_INVALID_SYMBOL; return R_ERR; }
        s->bits >>= used + dused; s->nbits -= used + dused;
        s->st.matches++; s->st.match_bytes += octets;
        r = copy_history(s, octets, k_bases[dnode >> 6] + dextra);
        if (r != R_OK) return r;
      } else {                                   /* end of block */
        s->bits >>= used; s->nbits -= used;
        s->cur = ST_BLOCK_END;
        break;
      }
    }
    break;
  case ST_CONTINUE_COPY_HISTORY: {               /* deflate.lisp:705-707 */
    s->cur = ST_DECODE_COMPRESSED_DATA;
    int r = copy_history(s, s->bytes_to_copy, s->copy_offset);
    if (r != R_OK) return r;
    break; }
  case ST_OUT_BYTE:                              /* deflate.lisp:709-716 */
    if (s->out_off >= s->out_cap) return eoo(s);
    s->out[s->out_off++] = (uint8_t)s->bytes_to_copy;
    s->st.literals++;
    s->cur = ST_DECODE_COMPRESSED_DATA;
    break;
  case ST_BLOCK_END:                             /* deflate.lisp:719-722 */
    s->cur = s->last_block ? ST_DONE : ST_START_OF_BLOCK;
    break;
  case ST_DONE:                                  /* deflate.lisp:725-726 */
    s->finished = 1;
    return R_OK;
  }
eoi:                                             /* deflate.lisp:114-120 */
  s->underrun = 1;
  return R_EOI;
}

/* ---- zlib.lisp ----------------------------------------------------------------------- */
static int zlib_adler(o3bz_context *c, o3bz_state *s) {  /* zlib.lisp:80-96 */
  if (!need(s, c, 32)) { s->underrun = 1; return R_EOI; }
  uint32_t a = take(s, 8) << 24; a |= take(s, 8) << 16; a |= take(s, 8) << 8; a |= take(s, 8);
  if (a != (s->s1 | (s->s2 << 16))) { s->error = O3_ERR_CHECKSUM; return R_ERR; }
  s->finished = 1;
  return R_OK;
}

static int64_t decompress_zlib(o3bz_context *c, o3bz_state *s) { /* zlib.lisp:39-144 */
  s->underrun = 0;
  if (s->zstate != ZS_NONE) {
    switch (s->zstate) {
    case ZS_HEADER: {
      if (!need(s, c, 16)) { s->underrun = 1; return 0; }
      uint32_t cmf = take(s, 8), flg = take(s, 8);
      /* check-zlib-header, zlib.lisp:14-37: errors in this order */
      if ((cmf * 256 + flg) % 31) { s->error = O3_ERR_ZLIB_FCHECK; return -1; }
      if ((cmf & 15) != 8) { s->error = O3_ERR_ZLIB_METHOD; return -1; }
      if ((cmf >> 4) > 7) { s->error = O3_ERR_ZLIB_WINDOW; return -1; }
      if (flg & 32) { s->zstate = ZS_HEADER2; s->error = O3_ERR_ZLIB_DICT; return -1; }
      break; }
    case ZS_HEADER2: s->error = O3_ERR_ZLIB_DICT; return -1;
    case ZS_ADLER: {
      int r = zlib_adler(c, s);
      if (r == R_ERR) return -1;
      if (r == R_EOI) return (int64_t)s->out_off;
      s->zstate = ZS_NONE;
      return (int64_t)s->out_off; }
    }
    s->zstate = ZS_NONE;
  }
  int r = decompress_deflate(c, s);
  if (r == R_ERR) return -1;
  if (s->finished || s->overflow) o3bz_adler32(s->out, s->out_off, &s->s1, &s->s2);
  if (s->finished) { byte_align(s); s->zstate = ZS_ADLER; s->finished = 0; }
  if (s->zstate == ZS_ADLER && zlib_adler(c, s) == R_ERR) return -1;
  return (int64_t)s->out_off;
}

/* ---- gzip.lisp ----------------------------------------------------------------------- */
static uint32_t hdr_byte(o3bz_state *s) {        /* gzip.lisp:69-74 */
  uint32_t b = take(s, 8);
  if (s->hdr_keep) {
    if (s->hdr_len == s->hdr_cap) {
      s->hdr_cap = s->hdr_cap ? s->hdr_cap * 2 : 64;
      s->hdr = (uint8_t *)realloc(s->hdr, s->hdr_cap);
    }
    s->hdr[s->hdr_len++] = (uint8_t)b;
  }
  return b;
}

static int64_t decompress_gzip(o3bz_context *c, o3bz_state *s) { /* gzip.lisp:30-287 */
  s->underrun = 0;
  while (!(s->finished || s->overflow || s->underrun)) switch (s->gstate) {
  case GS_HEADER: {                              /* gzip.lisp:113-122 */
    if (!need(s, c, 16)) { s->underrun = 1; return 0; }
    uint32_t id1 = hdr_byte(s), id2 = hdr_byte(s);
    if (id1 != 0x1f || id2 != 0x8b) { s->error = O3_ERR_GZIP_MAGIC; return -1; }
    s->gstate = GS_HEADER2; break; }
  case GS_HEADER2: {                             /* gzip.lisp:123-143 */
    if (!need(s, c, 16)) { s->underrun = 1; return 0; }
    uint32_t cm = hdr_byte(s), flg = hdr_byte(s);
    if (cm != 8) { s->error = O3_ERR_GZIP_METHOD; return -1; }
    if (flg >> 5) { s->error = O3_ERR_GZIP_RESERVED; return -1; }
    s->gflags = (int)flg;
    if (!(flg & 2)) s->hdr_keep = 0;
    s->gstate = GS_MTIME; break; }
  case GS_MTIME:                                 /* gzip.lisp:144-157 */
    if (!need(s, c, 32)) { s->underrun = 1; return 0; }
    for (int i = 0; i < 4; i++) hdr_byte(s);
    s->gstate = GS_HEADER3; break;
  case GS_HEADER3:                               /* gzip.lisp:159-177 */
    if (!need(s, c, 16)) { s->underrun = 1; return 0; }
    hdr_byte(s); hdr_byte(s);
    s->gstate = GS_EXTRA; break;
  case GS_EXTRA:                                 /* gzip.lisp:178-197 */
    if (s->gflags & 4) {
      if (!s->have_extra) {
        if (!need(s, c, 16)) { s->underrun = 1; return 0; }
        s->extra_len = hdr_byte(s); s->extra_len |= hdr_byte(s) << 8;
        s->extra_got = 0; s->have_extra = 1;
      }
      while (s->extra_got < s->extra_len) {
        if (!need(s, c, 8)) { s->underrun = 1; return 0; }
        hdr_byte(s); s->extra_got++;
      }
    }
    s->gstate = GS_NAME; break;
  case GS_NAME: case GS_COMMENT:                 /* gzip.lisp:198-241 */
    if (s->gflags & (s->gstate == GS_NAME ? 8 : 16))
      for (;;) {
        if (!need(s, c, 8)) { s->underrun = 1; return 0; }
        if (!hdr_byte(s)) break;
      }
    s->gstate++; break;
  case GS_HCRC:                                  /* gzip.lisp:242-266 */
    if (s->gflags & 2) {
      if (!need(s, c, 16)) { s->underrun = 1; return 0; }
      uint32_t crc = take(s, 8); crc |= take(s, 8) << 8;
      if (crc != (o3bz_crc32(s->hdr, s->hdr_len, 0) & 0xffff)) { s->error = O3_ERR_GZIP_HCRC; return -1; }
    }
    s->gstate = GS_DEFLATE; break;
  case GS_DEFLATE: {                             /* gzip.lisp:267-276 */
    int r = decompress_deflate(c, s);
    if (r == R_ERR) return -1;
    if (s->finished || s->overflow) s->crc = o3bz_crc32(s->out, s->out_off, s->crc);
    if (s->finished) { byte_align(s); s->gstate = GS_FINAL_CRC; s->finished = 0; }
    break; }
  case GS_FINAL_CRC: {                           /* gzip.lisp:82-94 */
    if (!need(s, c, 32)) { s->underrun = 1; return 0; }
    uint32_t crc = hdr_byte(s); crc |= hdr_byte(s) << 8; crc |= hdr_byte(s) << 16; crc |= hdr_byte(s) << 24;
    if (crc != s->crc) { s->error = O3_ERR_CHECKSUM; return -1; }
    s->gstate = GS_FINAL_LEN; break; }
  case GS_FINAL_LEN:                             /* gzip.lisp:95-106,277-286: ISIZE read, not checked */
    if (!need(s, c, 32)) { s->underrun = 1; return 0; }
    s->isize = hdr_byte(s); s->isize |= hdr_byte(s) << 8; s->isize |= hdr_byte(s) << 16; s->isize |= hdr_byte(s) << 24;
    s->finished = 1; s->gstate = GS_DONE;
    break;
  default:                                       /* ecase on :done */
    s->error = O3_ERR_STATE; return -1;
  }
  return (int64_t)s->out_off;
}

/* ---- api.lisp ------------------------------------------------------------------------ */
o3bz_state *o3bz_state_new(int format) {
  static_trees(); crc_table();
  o3bz_state *s = (o3bz_state *)calloc(1, sizeof *s);
  s->format = format;
  s->cur = ST_START_OF_BLOCK;
  tree_init(&s->dyn_lit); tree_init(&s->dyn_dist); tree_init(&s->len_tree);
  s->lit = &g_static_lit; s->dist = &g_static_dist;
  s->last_len = 0xff;
  s->zstate = format == O3_ZLIB ? ZS_HEADER : ZS_NONE;
  s->s1 = 1; s->s2 = 0;
  s->gstate = GS_HEADER; s->hdr_keep = 1;
  return s;
}
void o3bz_state_free(o3bz_state *s) {
  if (!s) return;
  free(s->window); free(s->hdr); free(s);
}
void o3bz_set_output(o3bz_state *s, uint8_t *buf, size_t cap) {
  s->out = buf; s->out_cap = cap; s->out_off = 0;
}
int o3bz_replace_output_buffer(o3bz_state *s, uint8_t *buf, size_t cap) { /* api.lisp:12-21 */
  if (!(s->out_off == 0 || s->overflow)) return O3_ERR_BUFFER_SWITCH;
  s->base_total += s->out_off;
  s->out = buf; s->out_cap = cap; s->out_off = 0; s->overflow = 0;
  return 0;
}
void o3bz_context_init(o3bz_context *c, const uint8_t *p, size_t start, size_t end) {
  c->p = p; c->start = start; c->end = end; c->offset = start;
}
int64_t o3bz_decompress(o3bz_context *c, o3bz_state *s) {   /* api.lisp:3-10 */
  if (s->error) return -1;
  switch (s->format) {
  case O3_GZIP: return decompress_gzip(c, s);
  case O3_ZLIB: return decompress_zlib(c, s);
  default: return decompress_deflate(c, s) == R_ERR ? -1 : (int64_t)s->out_off;
  }
}
int o3bz_finished(const o3bz_state *s) { return s->finished; }
int o3bz_input_underrun(const o3bz_state *s) { return s->underrun; }
int o3bz_output_overflow(const o3bz_state *s) { return s->overflow; }
int o3bz_error(const o3bz_state *s) { return s->error; }
uint32_t o3bz_checksum(const o3bz_state *s) {
  return s->format == O3_ZLIB ? (s->s1 | (s->s2 << 16)) : s->format == O3_GZIP ? s->crc : 0;
}
void o3bz_get_stats(const o3bz_state *s, o3bz_stats *st) {
  *st = s->st;
  st->total_out = s->base_total + s->out_off;
}

int o3bz_decompress_vector(const uint8_t *in, size_t start, size_t end, int format,
                           uint8_t *out, size_t out_cap, size_t *out_len,
                           uint32_t *checksum, o3bz_stats *st) {   /* api.lisp:36-48 */
  o3bz_state *s = o3bz_state_new(format);
  o3bz_context c;
  o3bz_context_init(&c, in, start, end);
  o3bz_set_output(s, out, out_cap);
  o3bz_decompress(&c, s);
  int v = s->error ? s->error : s->finished ? O3_FINISHED
        : s->underrun ? O3_INPUT_UNDERRUN : s->overflow ? O3_OUTPUT_OVERFLOW : O3_ERR_STATE;
  if (out_len) *out_len = s->out_off;
  if (checksum) {
    /* after an overflow/finish the running checksum covers out[0,out_off); otherwise compute it */
    uint32_t ck = 0;
    if (format == O3_ZLIB) { uint32_t a = 1, b = 0; o3bz_adler32(out, s->out_off, &a, &b); ck = a | (b << 16); }
    else if (format == O3_GZIP) ck = o3bz_crc32(out, s->out_off, 0);
    *checksum = ck;
  }
  if (st) o3bz_get_stats(s, st);
  o3bz_state_free(s);
  return v;
}

int o3bz_decompress_vector_grow(const uint8_t *in, size_t start, size_t end, int format,
                                uint8_t **out, size_t *out_len) {   /* api.lisp:50-65 */
  o3bz_state *s = o3bz_state_new(format);
  o3bz_context c;
  o3bz_context_init(&c, in, start, end);
  size_t cap = end - start < 32768 ? end - start : 32768, total = 0, acc_cap = 0;
  uint8_t *acc = NULL, *buf = NULL;
  int v = O3_FINISHED;
  for (;;) {
    buf = (uint8_t *)malloc(cap ? cap : 1);
    o3bz_replace_output_buffer(s, buf, cap);
    int64_t n = o3bz_decompress(&c, s);
    if (s->error) { v = s->error; free(buf); break; }
    if (s->underrun) { v = O3_INPUT_UNDERRUN; free(buf); break; }   /* (assert (not input-underrun)) */
    if (n > 0) {
      if (total + (size_t)n > acc_cap) { acc_cap = (total + (size_t)n) * 2; acc = (uint8_t *)realloc(acc, acc_cap); }
      memcpy(acc + total, buf, (size_t)n); total += (size_t)n;
    }
    free(buf);
    if (s->finished) break;
    cap = cap ? cap * 2 : 1;   /* (* 2 0) would never grow; the Lisp would loop forever on empty input */
  }
  o3bz_state_free(s);
  if (v != O3_FINISHED) { free(acc); acc = NULL; total = 0; }
  *out = acc; *out_len = total;
  return v;
}
void o3bz_free(void *p) { free(p); }

int o3bz_batch(const uint8_t *const *in, const size_t *in_len, uint8_t *const *out,
               const size_t *out_cap, size_t *out_len, int *verdict, uint64_t *match_bytes, int format,
               size_t lo, size_t hi) {
  int bad = 0;
  for (size_t i = lo; i < hi; i++) {
    size_t n = 0;
    o3bz_stats st;
    int v = o3bz_decompress_vector(in[i], 0, in_len[i], format, out[i], out_cap[i], &n, NULL, &st);
    if (match_bytes) match_bytes[i] = st.match_bytes;
    if (out_len) out_len[i] = n;
    if (verdict) verdict[i] = v;
    bad += v != O3_FINISHED;
  }
  return bad;
}
