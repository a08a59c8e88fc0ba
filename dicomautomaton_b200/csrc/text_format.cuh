// dcma_b200 -- double <-> decimal text, correctly rounded, usable from host and device code.
//
// The reference stores a deformation field as text: every double of every slice on its own line, written with
// `os.precision(max_digits10); os << val << '\n'` (src/Alignment_Field.cc:322-357) -- i.e. printf("%.17g\n") -- and
// read back with `is >> val` (src/Alignment_Field.cc:359-436).  For 512 x 512 x 256 that is 2.0e8 numbers / 4.8 GB of
// text, the step right after the registration.  This header holds the per-number arithmetic:
//
//   format_g17(x, out)  -> the characters printf("%.17g", x) produces (glibc, round-to-nearest), 1..24 of them
//   parse_double(s, n)  -> the double strtod() returns for a decimal literal of the form the reference writes
//
// Both are exact.  The fast path multiplies the 64-bit significand by a 128-bit truncated power of ten
// (pow10_table.inc) and keeps the error interval of the 192-bit product; whenever a rounding decision falls inside
// that interval the decision is made with exact integer arithmetic (Big, below) instead.  For 0 <= k <= 55 the table
// entry is exact and so is the product, which is where exact ties (dyadic rationals such as 100001/2^18) live.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define DCMA_TXT_HD __host__ __device__
#else
#define DCMA_TXT_HD
#endif

namespace dcma {
namespace txt {

struct Pow10 {
    uint64_t hi, lo;  // 10^k in [hi:lo, hi:lo + 1) * 2^e2, top bit of hi set
    int32_t e2;
};
constexpr int kPow10Min = -350, kPow10Max = 350;
constexpr int kMaxChars = 24;  // "-1.2345678901234567e-308"

#if defined(__CUDACC__)
static __device__ const Pow10 d_pow10[kPow10Max - kPow10Min + 1] = {
#include "pow10_table.inc"
};
#endif
static const Pow10 h_pow10[kPow10Max - kPow10Min + 1] = {
#include "pow10_table.inc"
};

DCMA_TXT_HD inline Pow10 pow10_entry(int k) {
#if defined(__CUDA_ARCH__)
    return d_pow10[k - kPow10Min];
#else
    return h_pow10[k - kPow10Min];
#endif
}

DCMA_TXT_HD inline uint64_t mul64(uint64_t a, uint64_t b, uint64_t *lo) {  // returns the high word
#if defined(__CUDA_ARCH__)
    *lo = a * b;
    return __umul64hi(a, b);
#else
    const unsigned __int128 p = static_cast<unsigned __int128>(a) * b;
    *lo = static_cast<uint64_t>(p);
    return static_cast<uint64_t>(p >> 64);
#endif
}
DCMA_TXT_HD inline int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll(static_cast<long long>(v));
#else
    return __builtin_clzll(v);
#endif
}
DCMA_TXT_HD inline uint64_t double_bits(double x) {
#if defined(__CUDA_ARCH__)
    return static_cast<uint64_t>(__double_as_longlong(x));
#else
    uint64_t b;
    std::memcpy(&b, &x, 8);
    return b;
#endif
}
DCMA_TXT_HD inline double bits_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(static_cast<long long>(b));
#else
    double x;
    std::memcpy(&x, &b, 8);
    return x;
#endif
}

DCMA_TXT_HD inline int clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __clz(static_cast<int>(v));
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

// ---- exact integers for the rare undecided cases (at most ~1150 bits are ever needed) --------------------------------
struct Big {
    static constexpr int W = 40;
    uint32_t w[W];
    int n;     // used words (no leading zero words), 0 for zero
    bool ovf;  // a result did not fit (only reachable from literals with hundreds of digits)
    DCMA_TXT_HD void set(uint64_t v) {
        n = 0;
        ovf = false;
        if (v) w[n++] = static_cast<uint32_t>(v);
        if (v >> 32) w[n++] = static_cast<uint32_t>(v >> 32);
    }
    DCMA_TXT_HD void mul_small(uint32_t f) {
        uint64_t carry = 0;
        for (int i = 0; i < n; ++i) {
            const uint64_t t = static_cast<uint64_t>(w[i]) * f + carry;
            w[i] = static_cast<uint32_t>(t);
            carry = t >> 32;
        }
        if (carry) {
            if (n < W) w[n++] = static_cast<uint32_t>(carry);
            else ovf = true;
        }
    }
    DCMA_TXT_HD void mul10_add(uint32_t digit) {
        uint64_t carry = digit;
        for (int i = 0; i < n; ++i) {
            const uint64_t t = static_cast<uint64_t>(w[i]) * 10u + carry;
            w[i] = static_cast<uint32_t>(t);
            carry = t >> 32;
        }
        if (carry) {
            if (n < W) w[n++] = static_cast<uint32_t>(carry);
            else ovf = true;
        }
    }
    DCMA_TXT_HD void mul_pow5(int p) {
        while (p >= 13) {
            mul_small(1220703125u);  // 5^13
            p -= 13;
        }
        uint32_t f = 1;
        while (p-- > 0) f *= 5u;
        if (f != 1u) mul_small(f);
    }
    DCMA_TXT_HD void shl(int bits) {
        if (n == 0 || bits <= 0) return;
        const int ws = bits >> 5, bs = bits & 31;
        int nn = n + ws + 1;
        if (nn > W) {
            // bits above W words would be lost: exact only if they are all zero
            const int top_bits = (n - 1) * 32 + (32 - static_cast<int>(clz32(w[n - 1]))) + bits;
            if (top_bits > W * 32) ovf = true;
            nn = W;
        }
        for (int i = nn - 1; i >= 0; --i) {
            const int src = i - ws;
            uint32_t v = 0;
            if (src >= 0 && src < n) v = w[src] << bs;
            if (bs && src - 1 >= 0 && src - 1 < n) v |= w[src - 1] >> (32 - bs);
            w[i] = v;
        }
        n = nn;
        while (n > 0 && w[n - 1] == 0) --n;
    }
    DCMA_TXT_HD int cmp(const Big &o) const {
        if (n != o.n) return n < o.n ? -1 : 1;
        for (int i = n - 1; i >= 0; --i)
            if (w[i] != o.w[i]) return w[i] < o.w[i] ? -1 : 1;
        return 0;
    }
};

// sign of  a * 10^k  -  h * 2^eh, exactly (a is consumed); *ovf is set when the integers did not fit
DCMA_TXT_HD inline int exact_compare(Big &a, int k, uint64_t h, int eh, bool *ovf) {
    Big b;
    b.set(h);
    if (k > 0) a.mul_pow5(k);
    if (k < 0) b.mul_pow5(-k);
    const int sh = k - eh;
    if (sh >= 0)
        a.shl(sh);
    else
        b.shl(-sh);
    if (ovf) *ovf = a.ovf || b.ovf;
    return a.cmp(b);
}
// sign of  m * 2^e * 10^k  -  h / 2   (m, h < 2^63)
DCMA_TXT_HD inline int exact_compare_half(uint64_t m, int e, int k, uint64_t h) {
    Big a;
    a.set(m);
    return exact_compare(a, k, h, -1 - e, nullptr);
}

// ---- double -> "%.17g" -------------------------------------------------------------------------------------------------
// mode 0: normal; mode 1: every rounding decision through the exact path (tests); *used_exact reports the exact path.
DCMA_TXT_HD inline int format_g17(double x, char *out, int mode = 0, int *used_exact = nullptr) {
    const uint64_t bits = double_bits(x);
    int n = 0;
    if (bits >> 63) out[n++] = '-';
    const uint32_t be = static_cast<uint32_t>(bits >> 52) & 0x7ffu;
    const uint64_t frac = bits & ((1ull << 52) - 1);
    if (be == 0x7ffu) {
        if (frac) {
            out[n++] = 'n'; out[n++] = 'a'; out[n++] = 'n';
        } else {
            out[n++] = 'i'; out[n++] = 'n'; out[n++] = 'f';
        }
        return n;
    }
    if (be == 0 && frac == 0) {
        out[n++] = '0';
        return n;
    }
    const uint64_t m = be ? (frac | (1ull << 52)) : frac;
    const int e = be ? static_cast<int>(be) - 1075 : -1074;  // x = m * 2^e
    const int lz = clz64(m);
    const uint64_t mn = m << lz;  // [2^63, 2^64)
    const int en = e - lz;
    const int xlow = ((en + 63) * 315653) >> 20;  // floor(log10(2^floor(log2 x))): exponent of x or one below
    const int k = 16 - xlow;                      // V = x * 10^k in [1e16, 2e17)
    const Pow10 g = pow10_entry(k);
    uint64_t a_lo, b_lo;
    const uint64_t a_hi = mul64(mn, g.lo, &a_lo);
    const uint64_t b_hi = mul64(mn, g.hi, &b_lo);
    const uint64_t p0 = a_lo, p1 = a_hi + b_lo, p2 = b_hi + (p1 < a_hi ? 1u : 0u);  // P = mn * G, V in [P, P + mn) / 2^s
    const int s2 = -(en + g.e2) - 128;                                              // s - 128, in [5, 10]
    const uint64_t d0 = p2 >> s2;
    uint64_t dh, mult, hb2;  // keep 17 digits dh; decide against the boundary hb2 * 2^128 of  F = (V - dh * mult) * 2^s
    int X;                   // decimal exponent of the leading digit
    if (d0 >= 100000000000000000ull) {
        dh = d0 / 10u; mult = 10u; hb2 = 5ull << s2; X = 17 - k;
    } else {
        dh = d0; mult = 1u; hb2 = 1ull << (s2 - 1); X = 16 - k;
    }
    const uint64_t f2 = p2 - ((dh * mult) << s2);
    int up = -1;
    if (mode == 0) {
        if (static_cast<unsigned>(k) <= 55u) {  // exact product
            if (f2 > hb2 || (f2 == hb2 && (p1 | p0))) up = 1;
            else if (f2 == hb2) up = static_cast<int>(dh & 1u);
            else up = 0;
        } else if (f2 >= hb2) {
            up = 1;  // the true value is strictly above P
        } else {
            const uint64_t g0 = p0 + mn;
            const uint64_t g1 = p1 + (g0 < p0 ? 1u : 0u);
            const uint64_t g2 = f2 + (g1 < p1 ? 1u : 0u);
            if (g2 < hb2 || (g2 == hb2 && g1 == 0 && g0 == 0)) up = 0;
        }
    }
    if (up < 0) {
        if (used_exact) *used_exact = 1;
        const uint64_t h = 2u * dh * mult + (mult == 1u ? 1u : 10u);
        const int c = exact_compare_half(m, e, k, h);
        up = c > 0 ? 1 : (c == 0 ? static_cast<int>(dh & 1u) : 0);
    }
    uint64_t dr = dh + static_cast<uint64_t>(up);
    if (dr >= 100000000000000000ull) {
        dr = 10000000000000000ull;
        ++X;
    }
    // 17 digits, most significant first
    char d[17];
    {
        uint32_t hi9 = static_cast<uint32_t>(dr / 100000000u), lo8 = static_cast<uint32_t>(dr % 100000000u);
        for (int i = 16; i >= 9; --i) { d[i] = static_cast<char>('0' + lo8 % 10u); lo8 /= 10u; }
        for (int i = 8; i >= 0; --i) { d[i] = static_cast<char>('0' + hi9 % 10u); hi9 /= 10u; }
    }
    int nd = 17;
    while (nd > 1 && d[nd - 1] == '0') --nd;
    if (X < -4 || X >= 17) {
        out[n++] = d[0];
        if (nd > 1) {
            out[n++] = '.';
            for (int i = 1; i < nd; ++i) out[n++] = d[i];
        }
        out[n++] = 'e';
        int ax = X;
        if (ax < 0) { out[n++] = '-'; ax = -ax; } else out[n++] = '+';
        if (ax >= 100) { out[n++] = static_cast<char>('0' + ax / 100); ax %= 100; out[n++] = static_cast<char>('0' + ax / 10); }
        else out[n++] = static_cast<char>('0' + ax / 10);
        out[n++] = static_cast<char>('0' + ax % 10);
    } else if (X >= 0) {
        for (int i = 0; i <= X; ++i) out[n++] = d[i];
        if (nd > X + 1) {
            out[n++] = '.';
            for (int i = X + 1; i < nd; ++i) out[n++] = d[i];
        }
    } else {
        out[n++] = '0';
        out[n++] = '.';
        for (int i = 0; i < -X - 1; ++i) out[n++] = '0';
        for (int i = 0; i < nd; ++i) out[n++] = d[i];
    }
    return n;
}


// ---- decimal text -> double (what `is >> val` / strtod return) ------------------------------------------------------------
struct U256 {
    uint64_t w[4];
    DCMA_TXT_HD static U256 shifted(uint64_t v, int s) {  // v << s, 0 <= s < 256 (bits beyond 256 are dropped)
        U256 r;
        r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0;
        const int wi = s >> 6, bs = s & 63;
        if (wi < 4) r.w[wi] = v << bs;
        if (bs && wi + 1 < 4) r.w[wi + 1] = v >> (64 - bs);
        return r;
    }
    DCMA_TXT_HD void add64(uint64_t v) {
        for (int i = 0; i < 4 && v; ++i) {
            const uint64_t t = w[i] + v;
            v = t < w[i] ? 1u : 0u;
            w[i] = t;
        }
    }
    DCMA_TXT_HD int cmp(const U256 &o) const {
        for (int i = 3; i >= 0; --i)
            if (w[i] != o.w[i]) return w[i] < o.w[i] ? -1 : 1;
        return 0;
    }
};

enum ParseStatus { kParseOk = 0, kParseMalformed = 1, kParseUndecided = 2, kParseOverflow = 3 };

// One token (no surrounding white space) of the form [+-]digits[.digits][(e|E)[+-]digits]; anything else -- including
// "inf"/"nan", which the reference's stream extraction rejects as well -- is kParseMalformed.  Results that overflow to
// infinity are kParseOverflow (the stream extraction sets failbit for them); underflow to zero/subnormal is fine.
DCMA_TXT_HD inline int parse_double(const char *s, int n, double *out, int mode = 0, int *used_exact = nullptr) {
    int i = 0;
    uint64_t sign = 0;
    if (i < n && (s[i] == '+' || s[i] == '-')) {
        sign = s[i] == '-' ? 1ull << 63 : 0;
        ++i;
    }
    uint64_t w = 0;  // first 19 significant digits
    int nd = 0, dexp = 0;
    bool any_digit = false, after_point = false, sticky = false;
    const int digits_begin = i;
    for (; i < n; ++i) {
        const char c = s[i];
        if (c == '.') {
            if (after_point) return kParseMalformed;
            after_point = true;
            continue;
        }
        if (c < '0' || c > '9') break;
        any_digit = true;
        const uint32_t dgt = static_cast<uint32_t>(c - '0');
        if (nd == 0 && dgt == 0) {  // leading zero
            if (after_point) --dexp;
            continue;
        }
        if (nd < 19) {
            w = w * 10u + dgt;
            ++nd;
            if (after_point) --dexp;
        } else {
            sticky = sticky || dgt != 0;
            if (!after_point) ++dexp;
        }
    }
    const int digits_end = i;
    if (!any_digit) return kParseMalformed;
    int ex = 0;
    if (i < n && (s[i] == 'e' || s[i] == 'E')) {
        ++i;
        bool eneg = false;
        if (i < n && (s[i] == '+' || s[i] == '-')) {
            eneg = s[i] == '-';
            ++i;
        }
        if (i >= n || s[i] < '0' || s[i] > '9') return kParseMalformed;
        for (; i < n && s[i] >= '0' && s[i] <= '9'; ++i)
            if (ex < 100000) ex = ex * 10 + (s[i] - '0');
        if (eneg) ex = -ex;
    }
    if (i != n) return kParseMalformed;
    if (w == 0) {
        *out = bits_double(sign);
        return kParseOk;
    }
    const long long q_ll = static_cast<long long>(dexp) + ex;  // value = (w [+ tail]) * 10^q
    if (q_ll > 310) return kParseOverflow;
    if (q_ll < -345) {  // below 1e19 * 1e-345 < 2^-1075: rounds to zero
        *out = bits_double(sign);
        return kParseOk;
    }
    const int q = static_cast<int>(q_ll);
    const int lz = clz64(w);
    const uint64_t wn = w << lz;
    const Pow10 g = pow10_entry(q);
    uint64_t a_lo, b_lo;
    const uint64_t a_hi = mul64(wn, g.lo, &a_lo);
    const uint64_t b_hi = mul64(wn, g.hi, &b_lo);
    const uint64_t p0 = a_lo, p1 = a_hi + b_lo, p2 = b_hi + (p1 < a_hi ? 1u : 0u);  // V in [P, P + wn) * 2^E
    const int E = g.e2 - lz;
    const int nb = 192 - clz64(p2);  // 191 or 192 significant bits
    int sh = nb - 53, ec = E + sh;   // candidate mantissa mc = P >> sh, value in [mc, mc + 1) * 2^ec
    if (ec < -1074) {                // subnormal range: the exponent is pinned
        ec = -1074;
        sh = -1074 - E;
    }
    if (sh >= 196) {  // V < 2^192 * 2^E <= 2^-1078: far below half the smallest subnormal
        *out = bits_double(sign);
        return kParseOk;
    }
    const uint64_t mc = (sh - 128 >= 64) ? 0 : (p2 >> (sh - 128));
    U256 mid = U256::shifted(2u * mc + 1u, sh - 1);
    U256 lo;
    lo.w[0] = p0; lo.w[1] = p1; lo.w[2] = p2; lo.w[3] = 0;
    int up = -1;
    if (mode == 0 && !sticky) {
        const int c = lo.cmp(mid);
        if (static_cast<unsigned>(q) <= 55u) {
            up = c > 0 ? 1 : (c == 0 ? static_cast<int>(mc & 1u) : 0);
        } else if (c >= 0) {
            up = 1;
        } else {
            U256 hi = lo;
            hi.add64(wn);
            if (hi.cmp(mid) <= 0) up = 0;
        }
    }
    if (up < 0) {
        if (used_exact) *used_exact = 1;
        Big a;
        int qa = q;
        if (!sticky) {
            a.set(w);
        } else {  // all significant digits of the literal
            a.set(0);
            bool started = false;
            int frac_digits = 0;
            bool pt = false;
            for (int j = digits_begin; j < digits_end; ++j) {
                if (s[j] == '.') { pt = true; continue; }
                const uint32_t dgt = static_cast<uint32_t>(s[j] - '0');
                if (pt) ++frac_digits;
                if (!started && dgt == 0) continue;
                started = true;
                a.mul10_add(dgt);
            }
            qa = ex - frac_digits;
        }
        bool ovf = false;
        const int c = exact_compare(a, qa, 2u * mc + 1u, ec - 1, &ovf);
        if (ovf) return kParseUndecided;
        up = c > 0 ? 1 : (c == 0 ? static_cast<int>(mc & 1u) : 0);
    }
    const uint64_t mr = mc + static_cast<uint64_t>(up);
    uint64_t bits;
    if (ec == -1074 && mr < (1ull << 53)) {
        // subnormal, or the smallest normals (mr in [2^52, 2^53) has biased exponent 1, which is what adding gives)
        bits = mr;
    } else {
        uint64_t mm = mr;
        int ee = ec;
        if (mm == (1ull << 53)) { mm >>= 1; ++ee; }
        const int be = ee + 1075;
        if (be >= 2047) return kParseOverflow;
        bits = (static_cast<uint64_t>(be) << 52) | (mm & ((1ull << 52) - 1));
    }
    *out = bits_double(bits | sign);
    return kParseOk;
}

}  // namespace txt
}  // namespace dcma
