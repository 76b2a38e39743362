/*
 * sjoracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See sjoracle.h.
 *
 * Scalar C restatement of the reference's hot path.  With -mavx2 -mpclmul the
 * five per-64-byte mask routines switch to an intrinsics restatement of the same
 * assembly (same instruction recipe: VPCMPEQB/VPMOVMSKB, PCLMULQDQ by all-ones,
 * two VPSHUFB nibble look-ups), which is what the CPU baseline times; the
 * scalar versions stay compiled in (sjo_*_scalar) so tests can compare both.
 *
 * Third-party arithmetic outside the reference checkout: Go's standard library
 * strconv.ParseInt / ParseUint / ParseFloat (Go >= 1.22, go.mod:3), called from
 * parse_number.go:105,114,130.  Restated here as: exact base-10 integer
 * accumulation with range checks, Go's readFloat syntax, and glibc strtod
 * (correctly rounded, round-half-even -- the same result ParseFloat guarantees).
 */
#define _GNU_SOURCE
#include "sjoracle.h"

#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#if defined(__AVX512BW__) && defined(__AVX512F__) && defined(__PCLMUL__) && !defined(SJO_FORCE_SCALAR) && defined(SJO_WANT_AVX512)
#include <immintrin.h>
#define SJO_SIMD 2 /* 64-byte ZMM loads, compares straight into k-mask registers: the reference's *_avx512 routines
                      (find_structural_bits_avx512_amd64.s:51-164, selected at stage1_find_marks_amd64.go:42) */
#elif defined(__AVX2__) && defined(__PCLMUL__) && !defined(SJO_FORCE_SCALAR)
#include <immintrin.h>
#define SJO_SIMD 1
#else
#define SJO_SIMD 0
#endif

/* which mask routines this build uses (bench.py prints it next to the CPU arm's number) */
const char *sjo_isa(void) { return SJO_SIMD == 2 ? "avx512bw+pclmul" : SJO_SIMD == 1 ? "avx2+pclmul" : "scalar"; }

/* ======================================================================= */
/* stage 1: per-64-byte mask routines                                       */
/* ======================================================================= */

static inline uint64_t mask_eq_scalar(const uint8_t *in, uint8_t c) {
    uint64_t m = 0;
    for (int i = 0; i < 64; i++) m |= (uint64_t)(in[i] == c) << i;
    return m;
}

#if SJO_SIMD == 2
/* VPCMPEQB zmm -> k (find_quote_mask_and_bits_amd64.s:117, find_odd_backslash_sequences_amd64.s:84-88) */
static inline uint64_t mask_eq(const uint8_t *in, uint8_t c) {
    return (uint64_t)_mm512_cmpeq_epi8_mask(_mm512_loadu_si512((const void *)in), _mm512_set1_epi8((char)c));
}
#elif SJO_SIMD
static inline uint64_t mask_eq(const uint8_t *in, uint8_t c) {
    __m256i lo = _mm256_loadu_si256((const __m256i *)in);
    __m256i hi = _mm256_loadu_si256((const __m256i *)(in + 32));
    __m256i k = _mm256_set1_epi8((char)c);
    uint64_t a = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(lo, k));
    uint64_t b = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(hi, k));
    return a | (b << 32);
}
#else
#define mask_eq mask_eq_scalar
#endif

/* find_odd_backslash_sequences_amd64.s:24-61 (mask), :27-58 (bit algebra) */
static inline uint64_t odd_backslash_from_mask(uint64_t bs, uint64_t *prev) {
    const uint64_t even_bits = 0x5555555555555555ULL, odd_bits = 0xAAAAAAAAAAAAAAAAULL;
    uint64_t p = *prev;
    uint64_t starts = bs & ~(bs << 1);
    uint64_t even_starts = starts & (even_bits ^ p);
    uint64_t odd_starts = starts & (odd_bits ^ p);
    uint64_t even_carries = bs + even_starts;
    uint64_t odd_carries = bs + odd_starts;
    uint64_t overflow = odd_carries < bs; /* SETCS */
    odd_carries |= p;
    *prev = overflow;
    uint64_t nbs = ~bs;
    return (even_carries & nbs & odd_bits) | (odd_carries & nbs & even_bits);
}

uint64_t sjo_find_odd_backslash_sequences(const uint8_t *in, uint64_t *prev) {
    return odd_backslash_from_mask(mask_eq(in, '\\'), prev);
}

static inline uint64_t prefix_xor(uint64_t x) {
#if SJO_SIMD
    /* find_quote_mask_and_bits_amd64.s:66 VPCLMULQDQ by all-ones */
    __m128i r = _mm_clmulepi64_si128(_mm_set_epi64x(0, (long long)x), _mm_set1_epi8((char)0xFF), 0);
    return (uint64_t)_mm_cvtsi128_si64(r);
#else
    x ^= x << 1;
    x ^= x << 2;
    x ^= x << 4;
    x ^= x << 8;
    x ^= x << 16;
    x ^= x << 32;
    return x;
#endif
}

static inline uint64_t mask_le_1f(const uint8_t *in) {
#if SJO_SIMD == 2
    /* find_quote_mask_and_bits_amd64.s:127-128: VPXORD 0x80, VPCMPGTB against 0xA0 into a k register */
    __m512i v = _mm512_xor_si512(_mm512_loadu_si512((const void *)in), _mm512_set1_epi8((char)0x80));
    return (uint64_t)_mm512_cmpgt_epi8_mask(_mm512_set1_epi8((char)0xA0), v);
#elif SJO_SIMD
    /* :67-80  (in ^ 0x80) <s 0xA0 */
    __m256i lo = _mm256_loadu_si256((const __m256i *)in);
    __m256i hi = _mm256_loadu_si256((const __m256i *)(in + 32));
    __m256i flip = _mm256_set1_epi8((char)0x80), lim = _mm256_set1_epi8((char)0xA0);
    uint64_t a = (uint32_t)_mm256_movemask_epi8(_mm256_cmpgt_epi8(lim, _mm256_xor_si256(lo, flip)));
    uint64_t b = (uint32_t)_mm256_movemask_epi8(_mm256_cmpgt_epi8(lim, _mm256_xor_si256(hi, flip)));
    return a | (b << 32);
#else
    uint64_t m = 0;
    for (int i = 0; i < 64; i++) m |= (uint64_t)(in[i] <= 0x1F) << i;
    return m;
#endif
}

/* find_quote_mask_and_bits_amd64.s:49-84 */
uint64_t sjo_find_quote_mask_and_bits(const uint8_t *in, uint64_t odd_ends, uint64_t *prev_inside, uint64_t *quote_bits,
                                      uint64_t *error_mask) {
    uint64_t qb = mask_eq(in, '"') & ~odd_ends;
    *quote_bits = qb;
    uint64_t qm = prefix_xor(qb) ^ *prev_inside;
    *error_mask |= mask_le_1f(in) & qm;
    *prev_inside = (uint64_t)((int64_t)qm >> 63);
    return qm;
}

/* find_whitespace_and_structurals_amd64.s:6-29 (tables), :62-103 */
#if !SJO_SIMD
static const uint8_t LO_NIBBLE[16] = {16, 0, 0, 0, 0, 0, 0, 0, 0, 8, 12, 1, 2, 9, 0, 0};
static const uint8_t HI_NIBBLE[16] = {8, 0, 18, 4, 0, 1, 0, 1, 0, 0, 0, 3, 2, 1, 0, 0};
#endif

void sjo_find_whitespace_and_structurals(const uint8_t *in, uint64_t *whitespace, uint64_t *structurals) {
    uint64_t ws = 0, st = 0;
#if SJO_SIMD == 2
    /* find_whitespace_and_structurals_amd64.s:136-148: two VPSHUFB zmm look-ups, VPANDD, VPCMPEQB against zero -> k, KNOTQ */
    const __m512i lo_tbl = _mm512_broadcast_i32x4(_mm_setr_epi8(16, 0, 0, 0, 0, 0, 0, 0, 0, 8, 12, 1, 2, 9, 0, 0));
    const __m512i hi_tbl = _mm512_broadcast_i32x4(_mm_setr_epi8(8, 0, 18, 4, 0, 1, 0, 1, 0, 0, 0, 3, 2, 1, 0, 0));
    const __m512i v = _mm512_loadu_si512((const void *)in);
    const __m512i l = _mm512_shuffle_epi8(lo_tbl, v);
    const __m512i hn = _mm512_and_si512(_mm512_srli_epi32(v, 4), _mm512_set1_epi8(0x7f));
    const __m512i c = _mm512_and_si512(l, _mm512_shuffle_epi8(hi_tbl, hn));
    st = (uint64_t)_mm512_test_epi8_mask(c, _mm512_set1_epi8(0x07));
    ws = (uint64_t)_mm512_test_epi8_mask(c, _mm512_set1_epi8(0x18));
#elif SJO_SIMD
    __m256i lo_tbl = _mm256_setr_epi8(16, 0, 0, 0, 0, 0, 0, 0, 0, 8, 12, 1, 2, 9, 0, 0, 16, 0, 0, 0, 0, 0, 0, 0, 0, 8, 12, 1,
                                      2, 9, 0, 0);
    __m256i hi_tbl = _mm256_setr_epi8(8, 0, 18, 4, 0, 1, 0, 1, 0, 0, 0, 3, 2, 1, 0, 0, 8, 0, 18, 4, 0, 1, 0, 1, 0, 0, 0, 3, 2,
                                      1, 0, 0);
    __m256i m7f = _mm256_set1_epi8(0x7f), zero = _mm256_setzero_si256();
    __m256i smask = _mm256_set1_epi8(0x07), wmask = _mm256_set1_epi8(0x18);
    for (int h = 0; h < 2; h++) {
        __m256i v = _mm256_loadu_si256((const __m256i *)(in + 32 * h));
        __m256i l = _mm256_shuffle_epi8(lo_tbl, v); /* bit 7 set -> 0 */
        __m256i hn = _mm256_and_si256(_mm256_srli_epi32(v, 4), m7f);
        __m256i hh = _mm256_shuffle_epi8(hi_tbl, hn);
        __m256i c = _mm256_and_si256(l, hh);
        uint64_t s = (uint32_t)~_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(c, smask), zero));
        uint64_t w = (uint32_t)~_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_and_si256(c, wmask), zero));
        st |= s << (32 * h);
        ws |= w << (32 * h);
    }
#else
    for (int i = 0; i < 64; i++) {
        uint8_t b = in[i];
        /* VPSHUFB zeroes the lane when bit 7 of the index byte is set; the high
         * nibble is masked to 0x7f first, so bytes >= 0x80 give v = 0 via lo. */
        uint8_t lo = (b & 0x80) ? 0 : LO_NIBBLE[b & 15];
        uint8_t hi = HI_NIBBLE[(b >> 4) & 15];
        uint8_t v = lo & hi;
        st |= (uint64_t)((v & 0x07) != 0) << i;
        ws |= (uint64_t)((v & 0x18) != 0) << i;
    }
#endif
    *whitespace = ws;
    *structurals = st;
}

/* finalize_structurals_amd64.s:19-36 */
uint64_t sjo_finalize_structurals(uint64_t structurals, uint64_t whitespace, uint64_t quote_mask, uint64_t quote_bits,
                                  uint64_t *prev_pseudo) {
    structurals &= ~quote_mask;
    structurals |= quote_bits;
    uint64_t pseudo_pred = structurals | whitespace;
    uint64_t shifted = (pseudo_pred << 1) | *prev_pseudo;
    *prev_pseudo = pseudo_pred >> 63;
    uint64_t pseudo = shifted & ~whitespace & ~quote_mask;
    structurals |= pseudo;
    structurals &= ~(quote_bits & ~quote_mask);
    return structurals;
}

/* find_newline_delimiters_amd64.s:16-28 */
uint64_t sjo_find_newline_delimiters(const uint8_t *in, uint64_t quote_mask) { return mask_eq(in, '\n') & ~quote_mask; }

/* flatten_bits_amd64.s:26-60 */
void sjo_flatten_bits_incremental(uint32_t *base, int *index, uint64_t mask, uint64_t *carried, uint64_t *position) {
    uint64_t shifts = 0;
    int idx = *index;
    int first = 1;
    while (mask) {
        uint64_t zeros = (uint64_t)__builtin_ctzll(mask);
        /* two shifts: zeros+1 may be 64 */
        mask >>= 1;
        mask >>= zeros;
        zeros += 1;
        shifts += zeros;
        if (first) {
            zeros += *carried;
            *carried = 0;
            first = 0;
        }
        base[idx++] = (uint32_t)zeros;
        *position += zeros;
    }
    *index = idx;
    *carried += 64 - shifts;
}

/* find_structural_bits_amd64.s:3-36 (single block, no flatten) */
uint64_t sjo_find_structural_bits(const uint8_t *in, uint64_t *prev_odd, uint64_t *prev_inside, uint64_t *error_mask,
                                  uint64_t structurals_in, uint64_t *prev_pseudo) {
    (void)structurals_in;
    uint64_t qb = 0, ws = 0, st = 0;
    uint64_t odd_ends = sjo_find_odd_backslash_sequences(in, prev_odd);
    uint64_t qm = sjo_find_quote_mask_and_bits(in, odd_ends, prev_inside, &qb, error_mask);
    sjo_find_whitespace_and_structurals(in, &ws, &st);
    return sjo_finalize_structurals(st, ws, qm, qb, prev_pseudo);
}

static inline void one_block(const uint8_t *in, uint64_t *prev_odd, uint64_t *prev_inside, uint64_t *error_mask,
                             uint64_t *prev_pseudo, uint32_t *indexes, int *index, uint64_t *carried, uint64_t *position,
                             uint64_t ndjson) {
    uint64_t qb = 0, ws = 0, st = 0;
    uint64_t odd_ends = sjo_find_odd_backslash_sequences(in, prev_odd);
    uint64_t qm = sjo_find_quote_mask_and_bits(in, odd_ends, prev_inside, &qb, error_mask);
    sjo_find_whitespace_and_structurals(in, &ws, &st);
    uint64_t s = sjo_finalize_structurals(st, ws, qm, qb, prev_pseudo);
    if (ndjson) s |= sjo_find_newline_delimiters(in, qm);
    sjo_flatten_bits_incremental(indexes, index, s, carried, position);
}

/* find_structural_bits_amd64.s:49-155 */
uint64_t sjo_find_structural_bits_in_slice(const uint8_t *buf, uint64_t len, uint64_t *prev_odd, uint64_t *prev_inside,
                                           uint64_t *error_mask, uint64_t *prev_pseudo, uint32_t *indexes, int *index,
                                           uint64_t *carried, uint64_t *position, uint64_t ndjson) {
    if (len == 0) return 0; /* find_subroutines_amd64.go:157 */
    uint64_t ax = 0, cx = len & ~63ULL;
    while (ax < cx) {
        one_block(buf + ax, prev_odd, prev_inside, error_mask, prev_pseudo, indexes, index, carried, position, ndjson);
        ax += 64;
        if (*index >= SJO_INDEX_SIZE_SAFETY) return ax; /* :111-112 */
    }
    uint64_t rem = len & 63;
    if (rem) { /* masking: bytes past the end become 0x20 (:134-155) */
        uint8_t pad[64];
        memset(pad, 0x20, 64);
        memcpy(pad, buf + ax, rem);
        one_block(pad, prev_odd, prev_inside, error_mask, prev_pseudo, indexes, index, carried, position, ndjson);
        ax += rem;
    }
    return ax;
}

static const uint8_t JSON_MARKUP[256] = {['{'] = 1, ['}'] = 1, ['['] = 1, [']'] = 1, [','] = 1, [':'] = 1};

/* stage1_find_marks_amd64.go:41-148, chunk loop restated; the per-chunk
 * channel hand-off becomes an append to the flat output. */
static int stage1_driver(const uint8_t *msg, size_t len, int ndjson, uint32_t *out, size_t cap, size_t *n_out) {
    uint64_t prev_odd = 0, prev_inside = 0, prev_pseudo = 1, error_mask = 0;
    uint64_t carried = 0, position = ~0ULL, stripped_index = ~0ULL;
    size_t index_total = 0;
    const uint8_t *buf = msg;
    size_t blen = len;
    int overflow = 0;
    uint32_t indexes[SJO_INDEX_SIZE];

    while (blen > 0) {
        int ilen = 0;
        if (stripped_index != ~0ULL) {
            position += stripped_index;
            indexes[0] = (uint32_t)stripped_index;
            ilen = 1;
            stripped_index = ~0ULL;
        }
        uint64_t processed = sjo_find_structural_bits_in_slice(buf, blen & ~(size_t)63, &prev_odd, &prev_inside,
                                                               &error_mask, &prev_pseudo, indexes, &ilen, &carried,
                                                               &position, (uint64_t)ndjson);
        if (blen - processed <= 64) {
            uint8_t padded[128];
            memset(padded, 0, sizeof padded);
            size_t pb = blen - processed;
            memcpy(padded, buf + processed, pb);
            processed += sjo_find_structural_bits_in_slice(padded, pb, &prev_odd, &prev_inside, &error_mask, &prev_pseudo,
                                                           indexes, &ilen, &carried, &position, (uint64_t)ndjson);
        }
        if (ilen == 0) {
            error_mask = ~0ULL;
            break;
        }
        if (blen == processed) {
            if (prev_inside != 0 || position >= blen || !(buf[position] == '}' || buf[position] == ']')) {
                error_mask = ~0ULL;
                break;
            }
        } else if (!JSON_MARKUP[buf[position]]) {
            stripped_index = indexes[ilen - 1];
            position -= stripped_index;
            ilen -= 1;
        }
        if (out) {
            if (index_total + (size_t)ilen > cap)
                overflow = 1;
            else
                memcpy(out + index_total, indexes, (size_t)ilen * sizeof(uint32_t));
        }
        index_total += (size_t)ilen;
        buf += processed;
        blen -= processed;
        position -= processed;
    }
    *n_out = index_total;
    if (overflow) return -1;
    return error_mask == 0 && index_total > 0;
}

int sjo_find_structural_indices(const uint8_t *msg, size_t len, int ndjson, uint32_t *deltas, size_t cap, size_t *n) {
    return stage1_driver(msg, len, ndjson, deltas, cap, n);
}

size_t sjo_stage1_count(const uint8_t *msg, size_t len, int ndjson, int *ok) {
    size_t n = 0;
    int r = stage1_driver(msg, len, ndjson, NULL, 0, &n);
    if (ok) *ok = r;
    return n;
}

/* ======================================================================= */
/* stage 2: strings                                                          */
/* ======================================================================= */

/* parse_string_amd64.s:4-69: digittoval (+0x40) and escape_map (+0x140).
 * Entries for bytes 0x00..0x2F of digittoval carry no DATA line => 0. */
static int8_t DIGIT_TO_VAL[256];
static uint8_t ESCAPE_MAP[256];
static int tables_ready = 0;

static void init_tables(void) {
    if (tables_ready) return;
    for (int i = 0; i < 256; i++) DIGIT_TO_VAL[i] = (i < 0x30) ? 0 : -1;
    for (int i = 0; i < 10; i++) DIGIT_TO_VAL['0' + i] = (int8_t)i;
    for (int i = 0; i < 6; i++) {
        DIGIT_TO_VAL['A' + i] = (int8_t)(10 + i);
        DIGIT_TO_VAL['a' + i] = (int8_t)(10 + i);
    }
    memset(ESCAPE_MAP, 0, sizeof ESCAPE_MAP);
    ESCAPE_MAP['"'] = 0x22;
    ESCAPE_MAP['/'] = 0x2f;
    ESCAPE_MAP['\\'] = 0x5c;
    ESCAPE_MAP['b'] = 0x08;
    ESCAPE_MAP['f'] = 0x0c;
    ESCAPE_MAP['n'] = 0x0a;
    ESCAPE_MAP['r'] = 0x0d;
    ESCAPE_MAP['t'] = 0x09;
    __atomic_store_n(&tables_ready, 1, __ATOMIC_RELEASE);
}

typedef struct {
    const uint8_t *p;
    size_t n;
} bytes_t;

static inline uint8_t at(const bytes_t *b, uint64_t i) { return i < b->n ? b->p[i] : 0; }

static inline void window_masks(const bytes_t *s, uint64_t p, uint32_t *bs, uint32_t *q) {
    uint32_t b = 0, qq = 0;
#if SJO_SIMD
    if (p + 32 <= s->n) { /* parse_string_amd64.s:92-100: one YMM load, two compares */
        __m256i v = _mm256_loadu_si256((const __m256i *)(s->p + p));
        *bs = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('\\')));
        *q = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, _mm256_set1_epi8('"')));
        return;
    }
#endif
    for (int i = 0; i < 32; i++) {
        uint8_t c = at(s, p + (uint64_t)i);
        b |= (uint32_t)(c == '\\') << i;
        qq |= (uint32_t)(c == '"') << i;
    }
    *bs = b;
    *q = qq;
}

static inline void copy_bytes(const bytes_t *s, uint64_t p, uint8_t *dst, uint32_t n) {
    if (p + n <= s->n) {
        memcpy(dst, s->p + p, n);
        return;
    }
    for (uint32_t i = 0; i < n; i++) dst[i] = at(s, p + i);
}

static inline int32_t hex4(const bytes_t *s, uint64_t p) {
    int32_t a = DIGIT_TO_VAL[at(s, p)], b = DIGIT_TO_VAL[at(s, p + 1)];
    int32_t c = DIGIT_TO_VAL[at(s, p + 2)], d = DIGIT_TO_VAL[at(s, p + 3)];
    return (int32_t)(((uint32_t)a << 12) | ((uint32_t)b << 8) | ((uint32_t)c << 4) | (uint32_t)d);
}

/* One escape/unicode step shared by validate and copy.  `p` is the window
 * start, `b` the offset of the first backslash in the window, `q` the quote mask
 * of the window.  Returns 0 on failure; otherwise *adv = source bytes consumed
 * from the backslash, *cp = code point or escape byte, *nbytes = UTF-8 length
 * (0 => single escape byte). parse_string_amd64.s:101-229 / :283-466 */
static int escape_step(const bytes_t *s, uint64_t p, uint32_t b, uint32_t q, uint32_t *adv, uint32_t *cp_out,
                       uint32_t *nbytes) {
    uint8_t e = at(s, p + b + 1);
    if (e != 'u') {
        if (ESCAPE_MAP[e] == 0) return 0;
        *cp_out = ESCAPE_MAP[e];
        *adv = 2;
        *nbytes = 0;
        return 1;
    }
    uint32_t dist;
    if (q != 0) {
        dist = (uint32_t)__builtin_ctz(q) - b;
    } else if (b < 21) {
        dist = 32 - b;
    } else {
        uint32_t q2 = 0;
        for (int i = 0; i < 32; i++) q2 |= (uint32_t)(at(s, p + b - 20 + (uint64_t)i) == '"') << i;
        dist = (q2 ? (uint32_t)__builtin_ctz(q2) : 32u) + b - 20 - b;
    }
    if (dist < 6) return 0;
    uint64_t bsl = p + b; /* position of the backslash */
    uint32_t cp = (uint32_t)hex4(s, bsl + 2);
    uint32_t a = 6;
    if ((cp & 0xFFFFFC00u) == 0xD800u) {
        if (dist < 12) return 0;
        if (at(s, bsl + 6) != '\\' || at(s, bsl + 7) != 'u') return 0;
        uint32_t cp2 = (uint32_t)hex4(s, bsl + 8);
        if ((cp | cp2) > 0xFFFFu) return 0;
        cp = (((cp << 10) + 0xFCA00000u) | (cp2 + 0xFFFF2400u)) + 0x10000u;
        a = 12;
    }
    uint32_t n;
    if (cp < 0x80u)
        n = 1;
    else if (cp < 0x800u)
        n = 2;
    else if (cp < 0x10000u)
        n = 3;
    else if (cp <= 0x10FFFFu)
        n = 4;
    else
        return 0;
    *cp_out = cp;
    *adv = a;
    *nbytes = n;
    return 1;
}

/* parse_string_amd64.s:72-258; wrapper parse_string_amd64.go:33-46 */
int sjo_parse_string_validate_only(const uint8_t *buf, size_t avail, uint64_t max_string_size, uint64_t *src_len,
                                   uint64_t *dst_len) {
    init_tables();
    bytes_t s = {buf + 1, avail ? avail - 1 : 0};
    if (max_string_size == 0) return 0;
    uint64_t p = 0, dl = 0;
    for (;;) {
        uint32_t bs, q;
        window_masks(&s, p, &bs, &q);
        if (((bs - 1) & q) != 0) {
            uint32_t t = (uint32_t)__builtin_ctz(q);
            *src_len = p + t;
            *dst_len = dl + t;
            return 1;
        }
        if (((q - 1) & bs) == 0) {
            p += 32;
            dl += 32;
        } else {
            uint32_t b = (uint32_t)__builtin_ctz(bs), adv, cp, n;
            if (!escape_step(&s, p, b, q, &adv, &cp, &n)) return 0;
            dl += b + (n ? n : 1);
            p += b + adv;
        }
        if (!(p < max_string_size)) return 0;
    }
}

/* parse_string_amd64.s:260-479; wrapper parse_string_amd64.go:48-59.  Only the
 * bytes that end up below the final length are observable, so the 32-byte
 * over-writes of the assembly are not reproduced. */
int sjo_parse_string(const uint8_t *buf, size_t avail, uint8_t *dst, uint64_t *dst_len) {
    init_tables();
    bytes_t s = {buf + 1, avail ? avail - 1 : 0};
    uint64_t p = 0, dl = 0;
    for (;;) {
        uint32_t bs, q;
        if (p > s.n + 64) return 0; /* the assembly would run away; never reached after validate */
        window_masks(&s, p, &bs, &q);
        if (((bs - 1) & q) != 0) {
            uint32_t t = (uint32_t)__builtin_ctz(q);
            copy_bytes(&s, p, dst + dl, t);
            *dst_len = dl + t;
            return 1;
        }
        if (((q - 1) & bs) == 0) {
            copy_bytes(&s, p, dst + dl, 32);
            p += 32;
            dl += 32;
            continue;
        }
        uint32_t b = (uint32_t)__builtin_ctz(bs), adv, cp, n;
        if (!escape_step(&s, p, b, q, &adv, &cp, &n)) return 0;
        copy_bytes(&s, p, dst + dl, b);
        dl += b;
        if (n <= 1) {
            dst[dl++] = (uint8_t)cp;
        } else if (n == 2) {
            dst[dl++] = (uint8_t)(0xC0 + (cp >> 6));
            dst[dl++] = (uint8_t)(0x80 | (cp & 63));
        } else if (n == 3) {
            dst[dl++] = (uint8_t)(0xE0 + (cp >> 12));
            dst[dl++] = (uint8_t)(0x80 | ((cp >> 6) & 63));
            dst[dl++] = (uint8_t)(0x80 | (cp & 63));
        } else {
            dst[dl++] = (uint8_t)(0xF0 + (cp >> 18));
            dst[dl++] = (uint8_t)(0x80 | ((cp >> 12) & 63));
            dst[dl++] = (uint8_t)(0x80 | ((cp >> 6) & 63));
            dst[dl++] = (uint8_t)(0x80 | (cp & 63));
        }
        p += b + adv;
    }
}

/* ======================================================================= */
/* stage 2: numbers                                                          */
/* ======================================================================= */

enum { F_PART = 1, F_FLOAT_ONLY = 2, F_MINUS = 4, F_EOV = 8, F_DIGIT = 16, F_MUST_DIGIT = 32 };

/* parse_number.go:36-60 */
static uint8_t number_rune(uint8_t c) {
    if (c >= '0' && c <= '9') return F_PART | F_DIGIT;
    switch (c) {
    case '.': return F_PART | F_FLOAT_ONLY | F_MUST_DIGIT;
    case '+': return F_PART;
    case '-': return F_PART | F_MINUS | F_MUST_DIGIT;
    case 'e':
    case 'E': return F_PART | F_FLOAT_ONLY;
    case ',':
    case '}':
    case ']':
    case ' ':
    case '\t':
    case '\r':
    case '\n':
    case ':': return F_EOV;
    default: return 0;
    }
}

/* Go strconv.ParseInt(s, 10, 64): 0 ok, 1 ErrSyntax, 2 ErrRange */
static int go_parse_int64(const uint8_t *s, size_t n, int64_t *out) {
    if (n == 0) return 1;
    int neg = 0;
    size_t i = 0;
    if (s[0] == '+' || s[0] == '-') {
        neg = s[0] == '-';
        i = 1;
        if (n == 1) return 1;
    }
    uint64_t v = 0;
    for (; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return 1;
        uint64_t d = (uint64_t)(s[i] - '0');
        if (v > (UINT64_MAX - d) / 10) return 2; /* ParseUint returns ErrRange at the first overflow */
        v = v * 10 + d;
    }
    if (!neg && v > (uint64_t)INT64_MAX) return 2;
    if (neg && v > (uint64_t)INT64_MAX + 1) return 2;
    *out = neg ? (int64_t)(0 - v) : (int64_t)v;
    return 0;
}

static int go_parse_uint64(const uint8_t *s, size_t n, uint64_t *out) {
    if (n == 0) return 1;
    uint64_t v = 0;
    for (size_t i = 0; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return 1;
        uint64_t d = (uint64_t)(s[i] - '0');
        if (v > (UINT64_MAX - d) / 10) return 2;
        v = v * 10 + d;
    }
    *out = v;
    return 0;
}

/* Go strconv.ParseFloat(s, 64) restricted to the bytes parseNumber lets
 * through (0-9 . + - e E): readFloat's syntax, then a correctly rounded
 * conversion.  Returns 0 ok, 1 syntax, 2 range (overflow to +-Inf). */
static int go_parse_float64(const uint8_t *s, size_t n, double *out) {
    size_t i = 0;
    if (n == 0) return 1;
    int neg = 0;
    if (s[i] == '+' || s[i] == '-') {
        neg = s[i] == '-';
        i++;
    }
    int sawdot = 0, sawdigits = 0;
    uint64_t mant = 0;
    int nd = 0, dp = 0, trunc = 0, ndmant = 0;
    for (; i < n; i++) {
        uint8_t c = s[i];
        if (c == '.') {
            if (sawdot) break;
            sawdot = 1;
            dp = nd;
            continue;
        }
        if (c >= '0' && c <= '9') {
            sawdigits = 1;
            if (c == '0' && nd == 0) {
                dp--;
                continue;
            }
            nd++;
            if (ndmant < 19) {
                mant = mant * 10 + (uint64_t)(c - '0');
                ndmant++;
            } else if (c != '0') {
                trunc = 1;
            }
            continue;
        }
        break;
    }
    if (!sawdigits) return 1;
    if (!sawdot) dp = nd;
    long exp10 = 0;
    if (i < n && (s[i] == 'e' || s[i] == 'E')) {
        i++;
        if (i >= n) return 1;
        int esign = 1;
        if (s[i] == '+')
            i++;
        else if (s[i] == '-') {
            i++;
            esign = -1;
        }
        if (i >= n || s[i] < '0' || s[i] > '9') return 1;
        long e = 0;
        for (; i < n && s[i] >= '0' && s[i] <= '9'; i++)
            if (e < 10000) e = e * 10 + (s[i] - '0');
        exp10 = e * esign;
    }
    if (i != n) return 1;
    /* Clinger fast path (exact): mantissa < 2^53 and |exponent| <= 22 */
    if (!trunc && mant != 0) {
        long e = exp10 + dp - ndmant;
        static const double P10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                     1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        if (nd == ndmant && mant < (1ULL << 53) && e >= -22 && e <= 22) {
            double d = (double)mant;
            d = e < 0 ? d / P10[-e] : d * P10[e];
            *out = neg ? -d : d;
            return 0;
        }
    }
    char stackbuf[128];
    char *tmp = n + 1 <= sizeof stackbuf ? stackbuf : (char *)malloc(n + 1);
    memcpy(tmp, s, n);
    tmp[n] = 0;
    errno = 0;
    char *endp = NULL;
    double d = strtod(tmp, &endp);
    int bad = (size_t)(endp - tmp) != n;
    if (tmp != stackbuf) free(tmp);
    if (bad) return 1;
    if (isinf(d)) return 2;
    *out = d;
    return 0;
}

/* parse_number.go:65-135 */
uint64_t sjo_parse_number(const uint8_t *buf, size_t len, uint64_t *val) {
    size_t pos = 0;
    uint8_t found = 0;
    for (size_t i = 0; i < len; i++) {
        uint8_t t = number_rune(buf[i]);
        if (t == 0) return 0;
        if (t == F_EOV) break;
        if (t & F_MUST_DIGIT) {
            if (len < i + 2 || !(number_rune(buf[i + 1]) & F_DIGIT)) return 0;
        }
        found |= t;
        pos = i + 1;
    }
    if (pos == 0) return 0;
    uint64_t float_tag = (uint64_t)'d' << 56;
    if (!(found & F_FLOAT_ONLY) && pos <= 20) {
        if (!(found & F_MINUS)) {
            if (pos > 1 && buf[0] == '0') return 0;
        } else {
            if (pos > 2 && buf[1] == '0') return 0;
        }
        int64_t i64;
        int r = go_parse_int64(buf, pos, &i64);
        if (r == 0) {
            *val = (uint64_t)i64;
            return (uint64_t)'l' << 56;
        }
        if (r == 2) float_tag |= 1;
        if (!(found & F_MINUS)) {
            uint64_t u64;
            r = go_parse_uint64(buf, pos, &u64);
            if (r == 0) {
                *val = u64;
                return (uint64_t)'u' << 56;
            }
            if (r == 2) float_tag |= 1;
        }
    } else if (!(found & F_FLOAT_ONLY)) {
        float_tag |= 1;
    }
    if (pos > 1 && buf[0] == '0' && !(number_rune(buf[1]) & F_FLOAT_ONLY)) return 0;
    double d;
    if (go_parse_float64(buf, pos, &d) != 0) return 0;
    memcpy(val, &d, 8);
    return float_tag;
}

/* ======================================================================= */
/* stage 2: atoms, state machine                                             */
/* ======================================================================= */

/* stage2_build_tape_amd64.go:455-476: 0 for structural / whitespace / NUL */
static inline int not_structural_or_ws(uint8_t c) {
    switch (c) {
    case 0:
    case '\t':
    case '\n':
    case '\r':
    case ' ':
    case ',':
    case ':':
    case '[':
    case ']':
    case '{':
    case '}': return 0;
    default: return 1;
    }
}

/* stage2_build_tape_amd64.go:124-158 */
int sjo_is_valid_true_atom(const uint8_t *b, size_t n) {
    return n >= 5 && memcmp(b, "true", 4) == 0 && !not_structural_or_ws(b[4]);
}
int sjo_is_valid_false_atom(const uint8_t *b, size_t n) {
    return n >= 6 && memcmp(b, "false", 5) == 0 && !not_structural_or_ws(b[5]);
}
int sjo_is_valid_null_atom(const uint8_t *b, size_t n) {
    return n >= 5 && memcmp(b, "null", 4) == 0 && !not_structural_or_ws(b[4]);
}

/* Go bytes.TrimSpace + unicode.IsSpace fall-back */
static int is_unicode_space(uint32_t r) {
    switch (r) {
    case '\t':
    case '\n':
    case '\v':
    case '\f':
    case '\r':
    case ' ':
    case 0x85:
    case 0xA0:
    case 0x1680:
    case 0x2028:
    case 0x2029:
    case 0x202F:
    case 0x205F:
    case 0x3000: return 1;
    default: return r >= 0x2000 && r <= 0x200A;
    }
}

static int is_ascii_space(uint8_t c) { return c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r' || c == ' '; }

/* utf8.DecodeRune: returns width, *r = rune (0xFFFD, width 1 when invalid) */
static int decode_rune(const uint8_t *p, size_t n, uint32_t *r) {
    if (n == 0) {
        *r = 0xFFFD;
        return 0;
    }
    uint8_t c = p[0];
    if (c < 0x80) {
        *r = c;
        return 1;
    }
    int need;
    uint32_t cp, min;
    if (c >= 0xC2 && c <= 0xDF) {
        need = 1;
        cp = c & 0x1F;
        min = 0x80;
    } else if (c >= 0xE0 && c <= 0xEF) {
        need = 2;
        cp = c & 0x0F;
        min = 0x800;
    } else if (c >= 0xF0 && c <= 0xF4) {
        need = 3;
        cp = c & 0x07;
        min = 0x10000;
    } else {
        *r = 0xFFFD;
        return 1;
    }
    if (n < (size_t)need + 1) {
        *r = 0xFFFD;
        return 1;
    }
    for (int i = 1; i <= need; i++) {
        if ((p[i] & 0xC0) != 0x80) {
            *r = 0xFFFD;
            return 1;
        }
        cp = (cp << 6) | (p[i] & 0x3F);
    }
    if (cp < min || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) {
        *r = 0xFFFD;
        return 1;
    }
    *r = cp;
    return need + 1;
}

static void trim_func(const uint8_t *buf, size_t *start, size_t *stop) {
    size_t a = *start, b = *stop;
    while (a < b) { /* TrimLeftFunc */
        uint32_t r;
        int w = decode_rune(buf + a, b - a, &r);
        if (!is_unicode_space(r)) break;
        a += (size_t)w;
    }
    while (b > a) { /* TrimRightFunc: utf8.DecodeLastRune */
        uint32_t r = 0xFFFD;
        size_t w = 1;
        if (buf[b - 1] < 0x80) {
            r = buf[b - 1];
        } else {
            size_t lim = b - a < 4 ? b - a : 4;
            for (size_t back = 1; back <= lim; back++) {
                uint8_t c = buf[b - back];
                if ((c & 0xC0) != 0x80) { /* a start byte */
                    uint32_t rr;
                    int ww = decode_rune(buf + b - back, back, &rr);
                    if ((size_t)ww == back) {
                        r = rr;
                        w = back;
                    }
                    break;
                }
            }
        }
        if (!is_unicode_space(r)) break;
        b -= w;
    }
    *start = a;
    *stop = b;
}

void sjo_trim_space(const uint8_t *buf, size_t len, size_t *start_out, size_t *stop_out) {
    size_t start = 0, stop = len;
    for (; start < len; start++) {
        uint8_t c = buf[start];
        if (c >= 0x80) {
            trim_func(buf, &start, &stop);
            goto done;
        }
        if (!is_ascii_space(c)) break;
    }
    for (; stop > start; stop--) {
        uint8_t c = buf[stop - 1];
        if (c >= 0x80) {
            trim_func(buf, &start, &stop);
            goto done;
        }
        if (!is_ascii_space(c)) break;
    }
done:
    *start_out = start;
    *stop_out = stop;
}

typedef struct {
    const uint8_t *msg;
    size_t len;
    const uint32_t *deltas;
    size_t n_idx, next;
    uint64_t *tape;
    size_t tape_cap, tape_len;
    uint8_t *strings;
    size_t str_cap, str_len;
    int copy_strings;
    int overflow;
    uint64_t *scope;
    size_t scope_len, scope_cap;
} machine_t;

static inline void write_tape(machine_t *m, uint64_t val, uint8_t c) {
    if (m->tape_len < m->tape_cap)
        m->tape[m->tape_len] = val | ((uint64_t)c << 56);
    else
        m->overflow = 1;
    m->tape_len++;
}
static inline void annotate(machine_t *m, uint64_t loc, uint64_t val) {
    if (loc < m->tape_cap) m->tape[loc] |= val;
}
static void push_scope(machine_t *m, uint64_t v) {
    if (m->scope_len == m->scope_cap) {
        m->scope_cap = m->scope_cap ? m->scope_cap * 2 : 128;
        m->scope = (uint64_t *)realloc(m->scope, m->scope_cap * sizeof(uint64_t));
    }
    m->scope[m->scope_len++] = v;
}

/* stage2_build_tape_amd64.go:72-113 */
static int machine_string(machine_t *m, uint64_t idx, uint64_t max_string_size) {
    uint64_t src_len = 0, dst_len = 0;
    if (!sjo_parse_string_validate_only(m->msg + idx, m->len - idx, max_string_size, &src_len, &dst_len)) return 0;
    int need_copy = m->copy_strings || src_len != dst_len;
    uint64_t size = dst_len;
    if (!need_copy) {
        write_tape(m, idx + 1, '"');
    } else {
        uint64_t start = m->str_len;
        if (start + dst_len <= m->str_cap) {
            uint64_t dl = 0;
            sjo_parse_string(m->msg + idx, m->len - idx, m->strings + start, &dl);
            size = dl;
        } else {
            m->overflow = 1;
        }
        m->str_len += size;
        write_tape(m, 0x80000000000000ULL + start, '"');
    }
    write_tape(m, size, 0);
    return 1;
}

#define UPDATE_CHAR()                                \
    do {                                             \
        if (m->next >= m->n_idx) goto succeed;       \
        idx += m->deltas[m->next++];                 \
    } while (0)
#define PEEK() (m->next < m->n_idx ? (uint64_t)m->deltas[m->next] : 0)

/* stage2_build_tape_amd64.go:160-446 */
static int unified_machine(machine_t *m) {
    const uint8_t *buf = m->msg;
    uint64_t idx = ~0ULL, offset;
    enum { RET_START = 1, RET_OBJECT = 2, RET_ARRAY = 3 };
    push_scope(m, ((uint64_t)m->tape_len << 2) | RET_START);
    write_tape(m, 0, 'r');
    UPDATE_CHAR();
continue_root:
    switch (buf[idx]) {
    case '{':
        push_scope(m, ((uint64_t)m->tape_len << 2) | RET_START);
        write_tape(m, 0, '{');
        goto object_begin;
    case '[':
        push_scope(m, ((uint64_t)m->tape_len << 2) | RET_START);
        write_tape(m, 0, '[');
        goto array_begin;
    default: goto fail;
    }
start_continue:
    UPDATE_CHAR();
    if (buf[idx] != '\n') goto fail;
    while (buf[idx] == '\n') UPDATE_CHAR();
    offset = m->scope[--m->scope_len];
    annotate(m, offset >> 2, (uint64_t)m->tape_len + 1);
    write_tape(m, offset >> 2, 'r');
    push_scope(m, ((uint64_t)m->tape_len << 2) | RET_START);
    write_tape(m, 0, 'r');
    goto continue_root;

object_begin:
    UPDATE_CHAR();
    switch (buf[idx]) {
    case '"':
        if (!machine_string(m, idx, PEEK())) goto fail;
        goto object_key_state;
    case '}': goto scope_end;
    default: goto fail;
    }
object_key_state:
    UPDATE_CHAR();
    if (buf[idx] != ':') goto fail;
    UPDATE_CHAR();
    switch (buf[idx]) {
    case '"':
        if (!machine_string(m, idx, PEEK())) goto fail;
        break;
    case 't':
        if (!sjo_is_valid_true_atom(buf + idx, m->len - idx)) goto fail;
        write_tape(m, 0, 't');
        break;
    case 'f':
        if (!sjo_is_valid_false_atom(buf + idx, m->len - idx)) goto fail;
        write_tape(m, 0, 'f');
        break;
    case 'n':
        if (!sjo_is_valid_null_atom(buf + idx, m->len - idx)) goto fail;
        write_tape(m, 0, 'n');
        break;
    case '{':
        push_scope(m, ((uint64_t)m->tape_len << 2) | RET_OBJECT);
        write_tape(m, 0, '{');
        goto object_begin;
    case '[':
        push_scope(m, ((uint64_t)m->tape_len << 2) | RET_OBJECT);
        write_tape(m, 0, '[');
        goto array_begin;
    default:
        if (buf[idx] == '-' || (buf[idx] >= '0' && buf[idx] <= '9')) {
            uint64_t val = 0, tag = sjo_parse_number(buf + idx, m->len - idx, &val);
            if (tag == 0) goto fail;
            write_tape(m, tag & 0x00FFFFFFFFFFFFFFULL, (uint8_t)(tag >> 56));
            write_tape(m, val, 0);
            break;
        }
        goto fail;
    }
object_continue:
    UPDATE_CHAR();
    switch (buf[idx]) {
    case ',':
        UPDATE_CHAR();
        if (buf[idx] != '"') goto fail;
        if (!machine_string(m, idx, PEEK())) goto fail;
        goto object_key_state;
    case '}': goto scope_end;
    default: goto fail;
    }
scope_end:
    offset = m->scope[--m->scope_len];
    write_tape(m, offset >> 2, buf[idx]);
    annotate(m, offset >> 2, (uint64_t)m->tape_len);
    switch (offset & 3) {
    case RET_ARRAY: goto array_continue;
    case RET_OBJECT: goto object_continue;
    default: goto start_continue;
    }
array_begin:
    UPDATE_CHAR();
    if (buf[idx] == ']') goto scope_end;
main_array_switch:
    switch (buf[idx]) {
    case '"':
        if (!machine_string(m, idx, PEEK())) goto fail;
        break;
    case 't':
        if (!sjo_is_valid_true_atom(buf + idx, m->len - idx)) goto fail;
        write_tape(m, 0, 't');
        break;
    case 'f':
        if (!sjo_is_valid_false_atom(buf + idx, m->len - idx)) goto fail;
        write_tape(m, 0, 'f');
        break;
    case 'n':
        if (!sjo_is_valid_null_atom(buf + idx, m->len - idx)) goto fail;
        write_tape(m, 0, 'n');
        break;
    case '{':
        push_scope(m, ((uint64_t)m->tape_len << 2) | RET_ARRAY);
        write_tape(m, 0, '{');
        goto object_begin;
    case '[':
        push_scope(m, ((uint64_t)m->tape_len << 2) | RET_ARRAY);
        write_tape(m, 0, '[');
        goto array_begin;
    default:
        if (buf[idx] == '-' || (buf[idx] >= '0' && buf[idx] <= '9')) {
            uint64_t val = 0, tag = sjo_parse_number(buf + idx, m->len - idx, &val);
            if (tag == 0) goto fail;
            write_tape(m, tag & 0x00FFFFFFFFFFFFFFULL, (uint8_t)(tag >> 56));
            write_tape(m, val, 0);
            break;
        }
        goto fail;
    }
array_continue:
    UPDATE_CHAR();
    switch (buf[idx]) {
    case ',':
        UPDATE_CHAR();
        goto main_array_switch;
    case ']': goto scope_end;
    default: goto fail;
    }
succeed:
    offset = m->scope[--m->scope_len];
    if (m->scope_len != 0) return 0;
    annotate(m, offset >> 2, (uint64_t)m->tape_len + 1);
    write_tape(m, offset >> 2, 'r');
    return 1;
fail:
    return 0;
}

/* parse_json_amd64.go:52-127 */
int sjo_parse(const uint8_t *msg, size_t len, uint32_t flags, uint64_t *tape, size_t tape_cap, size_t *tape_len,
              uint8_t *strings, size_t strings_cap, size_t *strings_len, size_t *msg_off, size_t *msg_len) {
    size_t a, b;
    sjo_trim_space(msg, len, &a, &b);
    if (msg_off) *msg_off = a;
    if (msg_len) *msg_len = b - a;
    const uint8_t *m0 = msg + a;
    size_t n = b - a;
    *tape_len = 0;
    *strings_len = 0;

    /* per-thread scratch, kept across calls (the reference reuses its index ring the same way) */
    static __thread uint32_t *tl_deltas = NULL;
    static __thread size_t tl_cap = 0;
    if (tl_cap < n + 64) {
        free(tl_deltas);
        tl_cap = n + 64 + (n >> 3);
        tl_deltas = (uint32_t *)malloc(tl_cap * sizeof(uint32_t) + 64);
    }
    uint32_t *deltas = tl_deltas;
    size_t n_idx = 0;
    int s1 = stage1_driver(m0, n, (flags & SJO_FLAG_NDJSON) != 0, deltas, n + 64, &n_idx);

    machine_t m;
    memset(&m, 0, sizeof m);
    m.msg = m0;
    m.len = n;
    m.deltas = deltas;
    m.n_idx = n_idx;
    m.tape = tape;
    m.tape_cap = tape_cap;
    m.strings = strings;
    m.str_cap = strings_cap;
    m.copy_strings = (flags & SJO_FLAG_COPY_STRINGS) != 0;
    int s2 = unified_machine(&m);
    free(m.scope);
    *tape_len = m.tape_len;
    *strings_len = m.str_len;
    if (s1 != 1) return SJO_ERR_STAGE1;
    if (!s2) return SJO_ERR_STAGE2;
    if (m.overflow) return SJO_ERR_CAPACITY;
    return SJO_OK;
}

/* ======================================================================= */
/* tape consumers of the reference's NDJSON tests / benchmarks              */
/* ======================================================================= */
#define SJO_VALUE_MASK 0x00ffffffffffffffull /* JSONVALUEMASK parsed_json.go:26 */
#define SJO_STRINGBUFBIT 0x80000000000000ull /* parsed_json.go:29 */

/* parsed_json.go:107-120 stringByteAt */
static const uint8_t *tape_string(uint64_t w, const uint8_t *strings, const uint8_t *msg) {
    uint64_t v = w & SJO_VALUE_MASK;
    return (v & SJO_STRINGBUFBIT) ? strings + (v - SJO_STRINGBUFBIT) : msg + v;
}

/* countWhere(key, value, pj) ndjson_test.go:421-459: the roots are chained through their
 * payload (count_raw_tape, ndjson_test.go:410-419); for every root whose element is an
 * object, Object.FindKey (parsed_object.go:97-140) returns the FIRST member named key, and
 * the record counts when that member is a string equal to value.  *roots = countObjects
 * (ndjson_test.go:461-474). */
uint64_t sjo_count_where(const uint64_t *tape, size_t tape_len, const uint8_t *strings, const uint8_t *msg,
                         const uint8_t *key, size_t klen, const uint8_t *value, size_t vlen, uint64_t *roots) {
    uint64_t count = 0, nroots = 0;
    size_t open = 0;
    while (open < tape_len) {
        uint64_t rw = tape[open];
        size_t next_root = (size_t)(rw & SJO_VALUE_MASK);
        if ((rw >> 56) != 'r' || next_root <= open) break;
        nroots++;
        if (open + 2 < tape_len && (tape[open + 1] >> 56) == '{') {
            size_t close = (size_t)(tape[open + 1] & SJO_VALUE_MASK) - 1;
            size_t i = open + 2;
            while (i < close && i + 2 < tape_len) {
                uint64_t kw = tape[i];
                if ((kw >> 56) != '"') break;
                uint64_t kl = tape[i + 1];
                size_t vi = i + 2;
                uint64_t vw = tape[vi];
                uint8_t vt = (uint8_t)(vw >> 56);
                if (kl == klen && memcmp(tape_string(kw, strings, msg), key, klen) == 0) {
                    if (vt == '"' && tape[vi + 1] == vlen && memcmp(tape_string(vw, strings, msg), value, vlen) == 0)
                        count++;
                    break;
                }
                size_t nx;
                if (vt == '"' || vt == 'l' || vt == 'u' || vt == 'd')
                    nx = vi + 2;
                else if (vt == '{' || vt == '[')
                    nx = (size_t)(vw & SJO_VALUE_MASK);
                else
                    nx = vi + 1;
                if (nx <= i) break;
                i = nx;
            }
        }
        open = next_root;
    }
    if (roots) *roots = nroots;
    return count;
}
