/* GPU check of dcma_b200_format_doubles / dcma_b200_parse_doubles through the C ABI, without Python: the text must be
 * byte-for-byte what printf("%.17g\n") writes (= `os << val << '\n'` at max_digits10, reference
 * src/Alignment_Field.cc:324-348) and must read back bit-for-bit (= `is >> val`, :418-420).  Sizes cross the library's
 * internal chunking in both directions.  usage: text_io_gpu_check [n_values]   (default 12 million) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "dcma_b200/demons_c_api.h"

static uint64_t s_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) {
    s_state ^= s_state << 13; s_state ^= s_state >> 7; s_state ^= s_state << 17;
    return s_state;
}
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 12000000;
    char err[512] = {0};
    if (dcma_b200_device_count() < 1) { printf("SKIP no device\n"); return 77; }
    double *v = (double *)malloc(sizeof(double) * n), *back = (double *)malloc(sizeof(double) * n);
    for (int64_t i = 0; i < n; ++i) {
        const unsigned kind = (unsigned)(rnd() % 8);
        double x;
        if (kind == 0) { uint64_t b = rnd(); if (((b >> 52) & 0x7ff) == 0x7ff) b &= ~(1ull << 62); memcpy(&x, &b, 8); } /* any finite bits */
        else if (kind == 1) x = ldexp((double)((rnd() >> 20) | 1), -(int)(rnd() % 80));                               /* dyadic: exact ties */
        else if (kind == 2) x = 0.05 * (double)((int64_t)(rnd() % 4001) - 2000);                                        /* grid values */
        else if (kind == 3) x = (rnd() % 16 == 0) ? 0.0 : (double)((int64_t)(rnd() % 2000001) - 1000000);
        else x = ((double)(int64_t)rnd() / 9.2e18) * 8.0 * pow(10.0, -(double)(rnd() % 12));                            /* displacements */
        v[i] = x;
    }
    char *text = (char *)malloc((size_t)n * 25), *want = (char *)malloc((size_t)n * 25 + 32);
    int64_t len = 0, wlen = 0;
    double t0 = now();
    int rc = dcma_b200_format_doubles(v, n, text, n * 25, &len, -1, err, sizeof err);
    double t1 = now();
    if (rc != DCMA_B200_OK) { printf("FAIL format rc=%d %s\n", rc, err); return 1; }
    for (int64_t i = 0; i < n; ++i) wlen += sprintf(want + wlen, "%.17g\n", v[i]);
    double t2 = now();
    if (len != wlen || memcmp(text, want, (size_t)len) != 0) {
        int64_t k = 0; while (k < len && k < wlen && text[k] == want[k]) ++k;
        printf("FAIL format: len %lld want %lld first difference at byte %lld: got '%.30s' want '%.30s'\n", (long long)len, (long long)wlen,
               (long long)k, text + (k > 10 ? k - 10 : 0), want + (k > 10 ? k - 10 : 0));
        return 1;
    }
    printf("format ok: %lld numbers, %lld bytes, gpu call %.3f s, printf %.3f s\n", (long long)n, (long long)len, t1 - t0, t2 - t1);
    int64_t used = 0;
    t0 = now();
    rc = dcma_b200_parse_doubles(text, len, n, back, &used, -1, err, sizeof err);
    t1 = now();
    if (rc != DCMA_B200_OK) { printf("FAIL parse rc=%d %s\n", rc, err); return 1; }
    if (used != len - 1) { printf("FAIL parse consumed %lld want %lld\n", (long long)used, (long long)(len - 1)); return 1; }
    if (memcmp(v, back, sizeof(double) * n) != 0) {
        int64_t k = 0; while (k < n && memcmp(v + k, back + k, 8) == 0) ++k;
        printf("FAIL parse: value %lld %a read back as %a\n", (long long)k, v[k], back[k]);
        return 1;
    }
    printf("parse ok: %lld numbers bit-identical, gpu call %.3f s\n", (long long)n, t1 - t0);
    /* a prefix only, odd white space, and the offset reported */
    const char *mixed = "  1.5\t-2e3\r\n\n 7   0.1 trailing";
    double m[4];
    rc = dcma_b200_parse_doubles(mixed, (int64_t)strlen(mixed), 4, m, &used, -1, err, sizeof err);
    if (rc != DCMA_B200_OK || m[0] != 1.5 || m[1] != -2000.0 || m[2] != 7.0 || m[3] != 0.1 || used != 21) {
        printf("FAIL mixed rc=%d used=%lld %g %g %g %g %s\n", rc, (long long)used, m[0], m[1], m[2], m[3], err); return 1; }
    /* errors: too few numbers, a token that is not a number, inf, overflow, too small a buffer */
    if (dcma_b200_parse_doubles("1 2 3", 5, 4, m, &used, -1, err, sizeof err) != DCMA_B200_EINVAL) { printf("FAIL short text accepted\n"); return 1; }
    if (dcma_b200_parse_doubles("1 2x 3", 6, 3, m, &used, -1, err, sizeof err) != DCMA_B200_EINVAL) { printf("FAIL bad token accepted\n"); return 1; }
    if (dcma_b200_parse_doubles("1 inf 3", 7, 3, m, &used, -1, err, sizeof err) != DCMA_B200_EINVAL) { printf("FAIL inf accepted\n"); return 1; }
    if (dcma_b200_parse_doubles("1e999", 5, 1, m, &used, -1, err, sizeof err) != DCMA_B200_EINVAL) { printf("FAIL overflow accepted\n"); return 1; }
    if (dcma_b200_format_doubles(v, 1000, text, 100, &len, -1, err, sizeof err) != DCMA_B200_EINVAL) { printf("FAIL small buffer accepted\n"); return 1; }
    if (dcma_b200_format_doubles(v, 0, text, 0, &len, -1, err, sizeof err) != DCMA_B200_OK || len != 0) { printf("FAIL empty input\n"); return 1; }
    printf("ALL OK\n");
    return 0;
}
