/* TEST stand-in for the three C-ABI entry points host/Field_Payload.cc calls, implemented with this C library's printf and
 * strtod -- which is what the reference's `os << val` / `is >> val` execute.  Linked INSTEAD of libdcma_b200.so by
 * tests/test_field_text.py so that the codec's host logic (bounded read-ahead, seeking back, failure states) and the
 * patched reference writer/reader can be exercised on a machine without a GPU.  Never part of the product. */
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FAKE_EINVAL (-1) /* DCMA_B200_EINVAL, include/dcma_b200/demons_c_api.h */

int dcma_b200_device_count(void) { return 1; }

int dcma_b200_format_doubles(const double *values, int64_t n, char *text, int64_t capacity, int64_t *text_len, int device,
                             char *errbuf, size_t errbuf_len) {
    (void)device; (void)errbuf; (void)errbuf_len;
    int64_t at = 0;
    for (int64_t i = 0; i < n; ++i) {
        char line[64];
        const int len = snprintf(line, sizeof line, "%.17g\n", values[i]);
        if (at + len > capacity) return FAKE_EINVAL;
        memcpy(text + at, line, (size_t)len);
        at += len;
    }
    *text_len = at;
    return 0;
}

int dcma_b200_parse_doubles(const char *text, int64_t text_len, int64_t n, double *values, int64_t *consumed, int device,
                            char *errbuf, size_t errbuf_len) {
    (void)device;
    int64_t at = 0;
    for (int64_t i = 0; i < n; ++i) {
        while (at < text_len && isspace((unsigned char)text[at])) ++at;
        int64_t end = at;
        while (end < text_len && !isspace((unsigned char)text[end])) ++end;
        char tok[512];
        if (end == at || end - at >= (int64_t)sizeof tok) { snprintf(errbuf, errbuf_len, "Pixel data could not be read"); return FAKE_EINVAL; }
        memcpy(tok, text + at, (size_t)(end - at));
        tok[end - at] = 0;
        char *stop = NULL;
        values[i] = strtod(tok, &stop);
        if (stop == tok || *stop != 0 || !(values[i] - values[i] == 0.0) || tok[0] == 'i' || tok[0] == 'n') {
            snprintf(errbuf, errbuf_len, "Pixel data could not be read (token '%s')", tok);
            return FAKE_EINVAL;
        }
        at = end;
    }
    *consumed = at;
    return 0;
}
