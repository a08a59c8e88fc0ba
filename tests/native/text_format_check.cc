// CPU check of dicomautomaton_b200/csrc/text_format.cuh (the same source the kernels compile): format_g17 against the
// C library's printf("%.17g") -- which is what the reference's `os << val` at max_digits10 produces
// (src/Alignment_Field.cc:324-350) -- over random bit patterns, displacement-like values, exact ties, powers of two and
// ten, subnormals and the non-finite values.  usage: text_format_check <n_random> [mode]
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../dicomautomaton_b200/csrc/text_format.cuh"

static long long g_checked = 0, g_bad = 0, g_exact = 0, g_pchecked = 0, g_pbad = 0, g_pexact = 0, g_undecided = 0;

// parse_double against strtod (which is what `is >> val` calls, reference src/Alignment_Field.cc:418-420)
static void check_parse(const char *lit, int mode) {
    char *end = nullptr;
    const double want = std::strtod(lit, &end);
    double got = 0.0;
    int used = 0;
    const int st = dcma::txt::parse_double(lit, static_cast<int>(std::strlen(lit)), &got, mode, &used);
    ++g_pchecked;
    g_pexact += used;
    if (st == dcma::txt::kParseUndecided) { ++g_undecided; return; }
    const bool overflow = std::isinf(want);
    bool ok;
    // strtod also takes "inf", "nan" and hex floats; the stream extraction of the reference does not (failbit), nor do we
    const bool letters = std::strpbrk(lit, "xXnNiI") != nullptr;
    if (*end != 0 || end == lit || letters) ok = st == dcma::txt::kParseMalformed;
    else if (overflow) ok = st == dcma::txt::kParseOverflow;
    else ok = st == dcma::txt::kParseOk && dcma::txt::double_bits(want) == dcma::txt::double_bits(got);
    if (!ok) {
        if (g_pbad < 20) std::printf("PARSE MISMATCH '%s': want %a got %a status %d\n", lit, want, got, st);
        ++g_pbad;
    }
}

static void check(double x, int mode) {
    char want[64], got[64];
    std::snprintf(want, sizeof want, "%.17g", x);
    int used = 0;
    const int n = dcma::txt::format_g17(x, got, mode, &used);
    got[n] = 0;
    ++g_checked;
    g_exact += used;
    if (std::strcmp(want, got) != 0 || n > dcma::txt::kMaxChars) {
        if (g_bad < 20) std::printf("MISMATCH %a: want '%s' got '%s'\n", x, want, got);
        ++g_bad;
    }
    if (std::isfinite(x)) {  // and back: the text must read as the same double
        check_parse(want, mode);
        double back = 0.0;
        if (dcma::txt::parse_double(got, n, &back, mode) != 0 || dcma::txt::double_bits(back) != dcma::txt::double_bits(x)) {
            if (g_pbad < 20) std::printf("ROUND TRIP %a -> '%s' -> %a\n", x, got, back);
            ++g_pbad;
        }
    }
}

int main(int argc, char **argv) {
    const long long n = argc > 1 ? std::atoll(argv[1]) : 1000000;
    const int mode = argc > 2 ? std::atoi(argv[2]) : 0;
    std::mt19937_64 rng(12345);
    // 1. random bit patterns (all exponents, both signs, NaN payloads)
    for (long long i = 0; i < n; ++i) check(dcma::txt::bits_double(rng()), mode);
    // 2. displacement-like values
    std::normal_distribution<double> nd(0.0, 3.0);
    for (long long i = 0; i < n; ++i) check(nd(rng), mode);
    for (long long i = 0; i < n / 4; ++i) check(nd(rng) * std::pow(10.0, static_cast<int>(rng() % 60) - 40), mode);
    // 3. short decimals and grid values (0.1 * col, -0.2 * row, 0.05 * slice of the reference's round-trip test)
    for (int i = -2000; i <= 2000; ++i) { check(0.1 * i, mode); check(-0.2 * i, mode); check(0.05 * i, mode); check(i, mode); check(i * 0.25, mode); }
    // 4. exact ties and near ties: odd m / 2^s
    for (int s = 1; s <= 80; ++s)
        for (int j = 0; j < 2000; ++j) {
            const uint64_t m = (rng() >> (11 + rng() % 40)) | 1u;
            check(std::ldexp(static_cast<double>(m), -s), mode);
            check(-std::ldexp(static_cast<double>(m), -s), mode);
        }
    for (uint64_t m = 1; m < 400000; m += 2) check(std::ldexp(static_cast<double>(m), -18), mode);  // 100001/2^18 = 0.381473541259765625
    // 5. powers of two and ten, their neighbours, boundaries of the %g style switch
    for (int e = -1074; e <= 1023; ++e) {
        const double p = std::ldexp(1.0, e);
        check(p, mode); check(std::nextafter(p, 0.0), mode); check(std::nextafter(p, INFINITY), mode); check(-p, mode);
    }
    for (int e = -323; e <= 308; ++e) {
        char buf[32];
        std::snprintf(buf, sizeof buf, "1e%d", e);
        const double p = std::strtod(buf, nullptr);
        check(p, mode); check(std::nextafter(p, 0.0), mode); check(std::nextafter(p, INFINITY), mode);
        std::snprintf(buf, sizeof buf, "9.9999999999999995e%d", e);
        check(std::strtod(buf, nullptr), mode);
        std::snprintf(buf, sizeof buf, "5e%d", e);
        check(std::strtod(buf, nullptr), mode);
    }
    // 6. integers up to 2^63 with many trailing zeros, large exact multiples of powers of five
    for (int j = 0; j < 200000; ++j) {
        uint64_t v = rng() >> (rng() % 64);
        check(static_cast<double>(v), mode);
    }
    {
        double p5 = 1.0;
        for (int j = 0; j <= 22; ++j, p5 *= 5.0)
            for (int e = -60; e <= 200; e += 1) { check(std::ldexp(p5, e), mode); check(std::ldexp(3 * p5, e), mode); check(std::ldexp(p5 * 15, e), mode); }
    }
    // 7. specials
    const double sp[] = {0.0, -0.0, INFINITY, -INFINITY, NAN, -NAN, 4.9406564584124654e-324, 2.2250738585072009e-308,
                         2.2250738585072014e-308, 1.7976931348623157e308, 0.0001, 0.00001, 1e16, 1e17, 123456789012345678.0,
                         99999999999999984.0, 99999999999999992.0, 1e22, 1e23, 9007199254740993.0, 0.3, 2.0 / 3.0};
    for (double v : sp) check(v, mode);
    // 8. decimal literals of every length, with and without exponents, and the halfway literals between adjacent doubles
    for (long long i = 0; i < n; ++i) {
        char lit[400];
        int p = 0;
        if (rng() & 1) lit[p++] = (rng() & 1) ? '-' : '+';
        const int nint = rng() % 4 == 0 ? 0 : 1 + static_cast<int>(rng() % 22), nfr = static_cast<int>(rng() % 24);
        for (int j = 0; j < nint; ++j) lit[p++] = static_cast<char>('0' + rng() % 10);
        if (nfr || rng() % 3 == 0) lit[p++] = '.';
        for (int j = 0; j < nfr; ++j) lit[p++] = static_cast<char>('0' + rng() % 10);
        if (nint + nfr == 0) lit[p++] = '7';
        if (rng() % 2) p += std::snprintf(lit + p, 16, "%c%d", (rng() & 1) ? 'e' : 'E', static_cast<int>(rng() % 700) - 350);
        lit[p] = 0;
        check_parse(lit, mode);
    }
    for (long long i = 0; i < n / 8; ++i) {  // exact midpoints (ties) and their neighbours, written out in full
        uint64_t b = rng();
        b = (b & ~(0x7ffull << 52)) | (static_cast<uint64_t>(1023 - 60 + rng() % 120) << 52);
        const double x = dcma::txt::bits_double(b & ~(1ull << 63));
        const long double midp = (static_cast<long double>(x) + static_cast<long double>(std::nextafter(x, INFINITY))) / 2;  // exact in 64-bit
        char lit[400];
        std::snprintf(lit, sizeof lit, "%.120Lf", midp);
        check_parse(lit, mode);
        const int L = static_cast<int>(std::strlen(lit));
        lit[L - 1] = '1';  // just above the tie
        check_parse(lit, mode);
    }
    const char *lits[] = {"0", "-0", "0.0", ".5", "5.", "1e0", "1E+2", "1e-2", "0e999", "1e309", "1.7976931348623157e308", "1.7976931348623159e308",
                          "1.8e308", "2.4703282292062327e-324", "2.4703282292062328e-324", "4.9406564584124654e-324", "2.2250738585072011e-308",
                          "2.2250738585072014e-308", "1e-400", "123456789012345678901234567890", "0.000000000000000000000000000001",
                          "9007199254740993", "9007199254740992.5", "9007199254740993.0000000000000000001", "1e23", "8.5e-324", "", "-", ".", "e5",
                          "1e", "1e+", "1.2.3", "12abc", "inf", "nan", "-inf", "0x10", "1e5.5", "+.e1", "00012.5000e-00", "1e400", "-1e400",
                          "0.1e310", "100e-326", "17976931348623157e292", "179769313486231580793e289"};
    for (const char *l : lits) check_parse(l, mode);
    std::printf("parse checked %lld mismatches %lld exact_path %lld undecided %lld\n", g_pchecked, g_pbad, g_pexact, g_undecided);
    std::printf("checked %lld mismatches %lld exact_path %lld mode %d\n", g_checked, g_bad, g_exact, mode);
    return (g_bad || g_pbad || g_undecided) ? 1 : 0;
}
