// number.cuh -- parseNumber on the device (parse_number.go:65-135).
//
// The reference delegates the arithmetic to Go's strconv (ParseInt / ParseUint /
// ParseFloat, parse_number.go:105,114,130).  Restated here:
//   * integers: exact base-10 accumulation with ErrRange / ErrSyntax distinction
//   * floats:   Go's readFloat syntax, then a correctly rounded (round-half-even) binary64:
//               Clinger exact fast path -> Eisel-Lemire with the 128-bit 10^q table ->
//               exact multi-precision decimal shifting (the classic "decimal" algorithm).
//     Every stage either returns THE correctly rounded value or defers to the next, so
//     the result is bit-identical to strconv.ParseFloat / glibc strtod.
#pragma once
#include "common.cuh"

namespace sj {

__device__ const uint64_t POW10_128[696][2] = {
#include "pow10_table.inc"
};
constexpr int POW10_MIN = -348, POW10_MAX = 347;

enum : uint32_t { NF_PART = 1, NF_FLOAT_ONLY = 2, NF_MINUS = 4, NF_EOV = 8, NF_DIGIT = 16, NF_MUST_DIGIT = 32 };

// parse_number.go:36-60
__device__ __forceinline__ uint32_t number_rune(uint32_t c) {
    if (c - '0' <= 9u) return NF_PART | NF_DIGIT;
    switch (c) {
    case '.': return NF_PART | NF_FLOAT_ONLY | NF_MUST_DIGIT;
    case '+': return NF_PART;
    case '-': return NF_PART | NF_MINUS | NF_MUST_DIGIT;
    case 'e':
    case 'E': return NF_PART | NF_FLOAT_ONLY;
    case ',':
    case '}':
    case ']':
    case ' ':
    case '\t':
    case '\r':
    case '\n':
    case ':': return NF_EOV;
    default: return 0;
    }
}

// ---- exact slow path: arbitrary-precision decimal, shifted by powers of two ------------
struct BigDec {
    uint8_t d[800];
    int nd, dp;
    bool trunc;
};

__device__ __noinline__ void bd_trim(BigDec& a) {
    while (a.nd > 0 && a.d[a.nd - 1] == 0) a.nd--;
    if (a.nd == 0) a.dp = 0;
}

__device__ __noinline__ void bd_rshift(BigDec& a, int k) {  // a /= 2^k, 1 <= k <= 60
    int r = 0, w = 0;
    uint64_t n = 0;
    for (; (n >> k) == 0; r++) {
        if (r >= a.nd) {
            if (n == 0) {
                a.nd = 0;
                return;
            }
            while ((n >> k) == 0) {
                n *= 10;
                r++;
            }
            break;
        }
        n = n * 10 + a.d[r];
    }
    a.dp -= r - 1;
    const uint64_t mask = (1ull << k) - 1;
    for (; r < a.nd; r++) {
        uint64_t dig = n >> k;
        n &= mask;
        a.d[w++] = (uint8_t)dig;
        n = n * 10 + a.d[r];
    }
    while (n > 0) {
        uint64_t dig = n >> k;
        n &= mask;
        if (w < 800)
            a.d[w++] = (uint8_t)dig;
        else if (dig > 0)
            a.trunc = true;
        n *= 10;
    }
    a.nd = w;
    bd_trim(a);
}

__device__ __noinline__ void bd_lshift(BigDec& a, int k) {  // a *= 2^k, 1 <= k <= 60
    if (a.nd == 0) return;
    const int maxd = ((k * 1233) >> 12) + 1;  // >= number of decimal digits 2^k can add
    int w = a.nd + maxd - 1;
    uint64_t n = 0;
    for (int r = a.nd - 1; r >= 0; r--) {
        n += (uint64_t)a.d[r] << k;
        uint64_t q = n / 10;
        uint32_t rem = (uint32_t)(n - 10 * q);
        if (w < 800)
            a.d[w] = (uint8_t)rem;
        else if (rem)
            a.trunc = true;
        w--;
        n = q;
    }
    while (n > 0) {
        uint64_t q = n / 10;
        uint32_t rem = (uint32_t)(n - 10 * q);
        if (w < 800)
            a.d[w] = (uint8_t)rem;
        else if (rem)
            a.trunc = true;
        w--;
        n = q;
    }
    const int lead = w + 1;  // unused leading positions
    int end = a.nd + maxd;
    if (end > 800) end = 800;
    const int newnd = end - lead;
    for (int i = 0; i < newnd; i++) a.d[i] = a.d[i + lead];
    a.nd = newnd;
    a.dp += maxd - lead;
    bd_trim(a);
}

__device__ __noinline__ void bd_shift(BigDec& a, int k) {
    if (a.nd == 0) return;
    while (k > 60) {
        bd_lshift(a, 60);
        k -= 60;
    }
    if (k > 0) bd_lshift(a, k);
    while (k < -60) {
        bd_rshift(a, 60);
        k += 60;
    }
    if (k < 0) bd_rshift(a, -k);
}

__device__ __noinline__ uint64_t bd_rounded_integer(const BigDec& a) {
    if (a.dp > 20) return ~0ull;
    uint64_t n = 0;
    int i = 0;
    for (; i < a.dp && i < a.nd; i++) n = n * 10 + a.d[i];
    for (; i < a.dp; i++) n *= 10;
    // should round up?
    bool up = false;
    if (a.dp >= 0 && a.dp < a.nd) {
        if (a.d[a.dp] == 5 && a.dp + 1 == a.nd) {  // exactly halfway: round to even
            up = a.trunc || (a.dp > 0 && (a.d[a.dp - 1] & 1));
        } else {
            up = a.d[a.dp] >= 5;
        }
    }
    return n + (up ? 1 : 0);
}

// text[0..n): sign already consumed by the caller; digits / '.' / exponent validated.
// Returns IEEE bits of |value| or ~0ull on overflow.
__device__ __noinline__ uint64_t slow_decimal_to_double(const uint8_t* text, int n) {
    BigDec a;
    a.nd = 0;
    a.dp = 0;
    a.trunc = false;
    bool sawdot = false;
    int i = 0;
    for (; i < n; i++) {
        uint32_t c = text[i];
        if (c == '.') {
            sawdot = true;
            a.dp = a.nd;
            continue;
        }
        if (c - '0' > 9u) break;
        if (c == '0' && a.nd == 0) {  // ignore leading zeros
            a.dp--;
            continue;
        }
        if (a.nd < 800)
            a.d[a.nd++] = (uint8_t)(c - '0');
        else if (c != '0')
            a.trunc = true;
    }
    if (!sawdot) a.dp = a.nd;
    if (i < n && (text[i] == 'e' || text[i] == 'E')) {
        i++;
        int esign = 1;
        if (text[i] == '+')
            i++;
        else if (text[i] == '-') {
            i++;
            esign = -1;
        }
        int e = 0;
        for (; i < n && (uint32_t)(text[i] - '0') <= 9u; i++)
            if (e < 10000) e = e * 10 + (text[i] - '0');
        a.dp += e * esign;
    }
    bd_trim(a);
    if (a.nd == 0) return 0;
    if (a.dp > 310) return ~0ull;
    if (a.dp < -330) return 0;
    const int powtab[9] = {1, 3, 6, 9, 13, 16, 19, 23, 26};
    int exp = 0;
    while (a.dp > 0) {
        int s = a.dp >= 9 ? 27 : powtab[a.dp];
        bd_shift(a, -s);
        exp += s;
    }
    while (a.dp < 0 || (a.dp == 0 && a.d[0] < 5)) {
        int s = -a.dp >= 9 ? 27 : powtab[-a.dp];
        bd_shift(a, s);
        exp -= s;
    }
    exp--;  // [0.5, 1) -> [1, 2)
    const int bias = -1023, mantbits = 52;
    if (exp < bias + 1) {  // denormal
        int s = bias + 1 - exp;
        bd_shift(a, -s);
        exp += s;
    }
    if (exp - bias >= 2047) return ~0ull;
    bd_shift(a, 1 + mantbits);
    uint64_t mant = bd_rounded_integer(a);
    if (mant == (2ull << mantbits)) {
        mant >>= 1;
        exp++;
        if (exp - bias >= 2047) return ~0ull;
    }
    if ((mant & (1ull << mantbits)) == 0) exp = bias;
    return (mant & ((1ull << mantbits) - 1)) | ((uint64_t)((exp - bias) & 2047) << mantbits);
}

// ---- Eisel-Lemire: man * 10^exp10 -> binary64 bits, or false when it cannot decide -------
__device__ __forceinline__ bool eisel_lemire64(uint64_t man, int exp10, uint64_t* bits) {
    if (man == 0) {
        *bits = 0;
        return true;
    }
    if (exp10 < POW10_MIN || exp10 > POW10_MAX) return false;
    int clz = __clzll(man);
    man <<= clz;
    uint64_t ret_exp2 = (uint64_t)(((217706 * exp10) >> 16) + 64 + 1023) - (uint64_t)clz;
    const uint64_t phi = POW10_128[exp10 - POW10_MIN][0], plo = POW10_128[exp10 - POW10_MIN][1];
    uint64_t x_hi = __umul64hi(man, phi), x_lo = man * phi;
    if ((x_hi & 0x1FF) == 0x1FF && x_lo + man < man) {
        uint64_t y_hi = __umul64hi(man, plo), y_lo = man * plo;
        uint64_t m_hi = x_hi, m_lo = x_lo + y_hi;
        if (m_lo < x_lo) m_hi++;
        if ((m_hi & 0x1FF) == 0x1FF && m_lo + 1 == 0 && y_lo + man < man) return false;
        x_hi = m_hi;
        x_lo = m_lo;
    }
    uint64_t msb = x_hi >> 63;
    uint64_t ret_man = x_hi >> (msb + 9);
    ret_exp2 -= 1 ^ msb;
    if (x_lo == 0 && (x_hi & 0x1FF) == 0 && (ret_man & 3) == 1) return false;  // half-way ambiguity
    ret_man += ret_man & 1;
    ret_man >>= 1;
    if (ret_man >> 53) {
        ret_man >>= 1;
        ret_exp2 += 1;
    }
    if (ret_exp2 - 1 >= 0x7FF - 1) return false;  // subnormal / overflow: let the exact path decide
    *bits = (ret_exp2 << 52) | (ret_man & 0x000FFFFFFFFFFFFFull);
    return true;
}

// Go strconv.ParseFloat on text[0..n) (only bytes 0-9 . + - e E can occur).
// Returns 0 ok (bits set), 1 syntax error, 2 range error (+-Inf).
__device__ __forceinline__ int go_parse_float64(const uint8_t* text, int n, uint64_t* bits) {
    int i = 0;
    if (n == 0) return 1;
    bool neg = false;
    if (text[0] == '+' || text[0] == '-') {
        neg = text[0] == '-';
        i = 1;
    }
    const int digits_start = i;
    bool sawdot = false, sawdigits = false, trunc = false;
    uint64_t mant = 0;
    int nd = 0, ndmant = 0, dp = 0;
    for (; i < n; i++) {
        uint32_t c = text[i];
        if (c == '.') {
            if (sawdot) break;
            sawdot = true;
            dp = nd;
            continue;
        }
        if (c - '0' <= 9u) {
            sawdigits = true;
            if (c == '0' && nd == 0) {
                dp--;
                continue;
            }
            nd++;
            if (ndmant < 19) {
                mant = mant * 10 + (c - '0');
                ndmant++;
            } else if (c != '0') {
                trunc = true;
            }
            continue;
        }
        break;
    }
    if (!sawdigits) return 1;
    if (!sawdot) dp = nd;
    int exp10 = 0;
    if (i < n && (text[i] == 'e' || text[i] == 'E')) {
        i++;
        if (i >= n) return 1;
        int esign = 1;
        if (text[i] == '+')
            i++;
        else if (text[i] == '-') {
            i++;
            esign = -1;
        }
        if (i >= n || (uint32_t)(text[i] - '0') > 9u) return 1;
        int e = 0;
        for (; i < n && (uint32_t)(text[i] - '0') <= 9u; i++)
            if (e < 10000) e = e * 10 + (text[i] - '0');
        exp10 = e * esign;
    }
    if (i != n) return 1;
    const uint64_t sign = neg ? 0x8000000000000000ull : 0;
    if (mant == 0) {
        *bits = sign;
        return 0;
    }
    const int e = exp10 + dp - ndmant;  // value = mant * 10^e (mant truncated to 19 digits when trunc)
    // Clinger: exact when the mantissa fits 53 bits and one correctly rounded * or / by an exact 10^|e|
    if (!trunc && mant < (1ull << 53) && e >= -22 && e <= 22) {
        const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        double d = (double)(long long)mant;
        d = e < 0 ? __ddiv_rn(d, P10[-e]) : __dmul_rn(d, P10[e]);
        *bits = (uint64_t)__double_as_longlong(d) | sign;
        return 0;
    }
    uint64_t b;
    if (eisel_lemire64(mant, e, &b)) {
        bool ok = true;
        if (trunc) {  // the true mantissa lies in (mant, mant+1): both ends must agree
            uint64_t b2;
            ok = eisel_lemire64(mant + 1, e, &b2) && b2 == b;
        }
        if (ok) {
            *bits = b | sign;
            return 0;
        }
    }
    b = slow_decimal_to_double(text + digits_start, n - digits_start);
    if (b == ~0ull) return 2;
    *bits = b | sign;
    return 0;
}

// strconv.ParseInt(s, 10, 64): 0 ok, 1 syntax, 2 range
__device__ __forceinline__ int go_parse_int64(const uint8_t* s, int n, uint64_t* out) {
    if (n == 0) return 1;
    bool neg = false;
    int i = 0;
    if (s[0] == '+' || s[0] == '-') {
        neg = s[0] == '-';
        i = 1;
        if (n == 1) return 1;
    }
    uint64_t v = 0;
    for (; i < n; i++) {
        uint32_t d = s[i] - '0';
        if (d > 9u) return 1;
        if (v > (~0ull - d) / 10) return 2;
        v = v * 10 + d;
    }
    if (!neg && v > 0x7FFFFFFFFFFFFFFFull) return 2;
    if (neg && v > 0x8000000000000000ull) return 2;
    *out = neg ? (0 - v) : v;
    return 0;
}

__device__ __forceinline__ int go_parse_uint64(const uint8_t* s, int n, uint64_t* out) {
    if (n == 0) return 1;
    uint64_t v = 0;
    for (int i = 0; i < n; i++) {
        uint32_t d = s[i] - '0';
        if (d > 9u) return 1;
        if (v > (~0ull - d) / 10) return 2;
        v = v * 10 + d;
    }
    *out = v;
    return 0;
}

// parse_number.go:65-135.  buf[0..avail) is the rest of the message from the number's first
// byte.  Returns the tape tag word (tag << 56 | flags) or 0; *val = raw 64-bit value.
__device__ __forceinline__ uint64_t parse_number(const uint8_t* buf, uint64_t avail, uint64_t* val) {
    uint32_t found = 0;
    int pos = 0;
    for (uint64_t i = 0; i < avail; i++) {
        uint32_t t = number_rune(buf[i]);
        if (t == 0) return 0;
        if (t == NF_EOV) break;
        if (t & NF_MUST_DIGIT) {
            if (avail < i + 2 || !(number_rune(buf[i + 1]) & NF_DIGIT)) return 0;
        }
        found |= t;
        pos = (int)i + 1;
        if (pos >= 0x7ffffff0) return 0;  // (a single number longer than 2 GiB cannot occur: SJ_MAX_MESSAGE)
    }
    if (pos == 0) return 0;
    uint64_t float_tag = (uint64_t)'d' << 56;
    if (!(found & NF_FLOAT_ONLY) && pos <= 20) {
        if (!(found & NF_MINUS)) {
            if (pos > 1 && buf[0] == '0') return 0;
        } else {
            if (pos > 2 && buf[1] == '0') return 0;
        }
        uint64_t v;
        int r = go_parse_int64(buf, pos, &v);
        if (r == 0) {
            *val = v;
            return (uint64_t)'l' << 56;
        }
        if (r == 2) float_tag |= 1;  // FloatOverflowedInteger, parsed_json.go:40-44
        if (!(found & NF_MINUS)) {
            r = go_parse_uint64(buf, pos, &v);
            if (r == 0) {
                *val = v;
                return (uint64_t)'u' << 56;
            }
            if (r == 2) float_tag |= 1;
        }
    } else if (!(found & NF_FLOAT_ONLY)) {
        float_tag |= 1;
    }
    if (pos > 1 && buf[0] == '0' && !(number_rune(buf[1]) & NF_FLOAT_ONLY)) return 0;
    uint64_t bits;
    if (go_parse_float64(buf, pos, &bits) != 0) return 0;
    *val = bits;
    return float_tag;
}

// The common shapes in one pass:  [-] digits [ . digits ]  followed by an end-of-value byte, at most 18 digits in all,
// no leading zero other than a lone "0" in front of the point, no exponent.  Such a number cannot hit any of the
// special rules of parse_number.go:65-135 (leading zeros, integer overflow into float, must-have-digit, the 20-character
// limit) or of strconv (19-digit truncation), so the result is computed directly: integers as 'l', decimals through
// Clinger's exact case or Eisel-Lemire.  Everything else -- and the rare Eisel-Lemire "cannot decide" -- returns
// PN_SLOW and takes parse_number() above, which stays the definition of the behaviour.
constexpr uint64_t PN_SLOW = ~0ull;
__device__ __forceinline__ uint64_t parse_number_fast(const uint8_t* buf, uint64_t avail, uint64_t* val) {
    if (avail < 26) return PN_SLOW;  // (the loop below never looks past the message)
    uint32_t i = 0;
    const bool neg = buf[0] == '-';
    i = neg ? 1 : 0;
    const uint32_t first = buf[i];
    uint64_t mant = 0;
    uint32_t nint = 0, nfrac = 0;
    bool dot = false;
    uint32_t c;
    for (;;) {
        c = buf[i];
        const uint32_t d = c - '0';
        if (d <= 9u) {
            mant = mant * 10 + d;
            if (dot)
                nfrac++;
            else
                nint++;
        } else if (c == '.' && !dot && nint != 0) {
            dot = true;
        } else {
            break;
        }
        i++;
        if (i >= 24) return PN_SLOW;
    }
    if (nint == 0 || (dot && nfrac == 0) || nint + nfrac > 18) return PN_SLOW;
    if (first == '0' && nint != 1) return PN_SLOW;  // leading zeros: the full rules decide
    if (!(c == ',' || c == '}' || c == ']' || c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == ':')) return PN_SLOW;
    if (!dot) {  // ParseInt succeeds (at most 18 digits): int64, "-0" included
        *val = neg ? (0 - mant) : mant;
        return (uint64_t)'l' << 56;
    }
    const uint64_t sign = neg ? 0x8000000000000000ull : 0;
    if (mant == 0) {
        *val = sign;
        return (uint64_t)'d' << 56;
    }
    const int e = -(int)nfrac;
    if (mant < (1ull << 53) && e >= -22) {
        const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        const double dv = __ddiv_rn((double)(long long)mant, P10[-e]);
        *val = (uint64_t)__double_as_longlong(dv) | sign;
        return (uint64_t)'d' << 56;
    }
    uint64_t b;
    if (!eisel_lemire64(mant, e, &b)) return PN_SLOW;
    *val = b | sign;
    return (uint64_t)'d' << 56;
}

}  // namespace sj
