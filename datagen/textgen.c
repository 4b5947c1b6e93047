/* textgen.c — deterministic synthetic text for the bench/test configurations.
 * Integer-only so every build produces the same bytes.  Not part of the product. */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

static inline uint64_t splitmix(uint64_t *s) {
  *s += 0x9E3779B97F4A7C15ull;
  uint64_t z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

#define NWORDS 4096
static char g_words[NWORDS][12];
static uint8_t g_wlen[NWORDS];
static int g_ready = 0;

static void vocab(void) {
  if (g_ready) return;
  static const char letters[] = "etaoinshrdlcumwfgypbvkjxqz";
  uint64_t s = 0x3B2;
  for (int w = 0; w < NWORDS; w++) {
    uint64_t r = splitmix(&s);
    int len = 2 + (int)(r % 9);
    for (int i = 0; i < len; i++) {
      r = splitmix(&s);
      unsigned idx = (unsigned)(((r & 0xff) * ((r >> 8) & 0xff)) >> 11);
      g_words[w][i] = letters[idx > 25 ? 25 : idx];
    }
    g_wlen[w] = (uint8_t)len;
  }
  g_ready = 1;
}

void tbzgen_init(void) { vocab(); }

/* Zipf-ish word stream, punctuation and line breaks; exactly `size` bytes. */
void tbzgen_text(uint8_t *dst, size_t size, uint64_t seed) {
  vocab();
  uint64_t s = seed;
  size_t n = 0;
  int line_words = 8, in_line = 0;
  char tmp[16];
  while (n < size) {
    uint64_t r = splitmix(&s);
    unsigned k = (unsigned)(r % 12);
    unsigned rank = ((1u << k) - 1) + (unsigned)((r >> 8) & ((1u << k) - 1));
    int len = g_wlen[rank];
    memcpy(tmp, g_words[rank], (size_t)len);
    if ((r >> 40) % 32 == 0) tmp[0] = (char)(tmp[0] - 32);
    unsigned p = (unsigned)((r >> 48) % 16);
    if (p == 0) tmp[len++] = ',';
    else if (p == 1) tmp[len++] = '.';
    if (++in_line == line_words) { tmp[len++] = '\n'; in_line = 0; line_words = 8 + (int)((r >> 56) % 8); }
    else tmp[len++] = ' ';
    size_t c = (size_t)len < size - n ? (size_t)len : size - n;
    memcpy(dst + n, tmp, c);
    n += c;
  }
}

/* incompressible bytes (config 5) */
void tbzgen_random(uint8_t *dst, size_t size, uint64_t seed) {
  uint64_t s = seed;
  size_t n = 0;
  while (n < size) {
    uint64_t r = splitmix(&s);
    for (int i = 0; i < 8 && n < size; i++) dst[n++] = (uint8_t)(r >> (8 * i));
  }
}
