/*
 * hts_oracle.c -- CPU restatement of the htscodecs static-rANS hot path (see hts_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: never linked into or called from the product path.
 *
 * Written from the reference's behaviour, not copied from it: one N-way coder (N = 4 reproduces
 * the reference byte-for-byte, N = 32 is the X_32 generalisation), plain division instead of the
 * reciprocal trick, one careful loop instead of unrolled fast/slow pairs.  Every function names
 * the reference lines (under /root/reference/htscodecs/) whose results it must reproduce.
 *
 * Parity: N=4 and 4x8 decode PINNED (golden streams + libref.so); X_32 PARITY UNPINNED.
 */
#include "hts_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>

#define L16   (1u << 15)   /* 4x16 lower state bound, rANS_word.h:63 */
#define L8    (1u << 23)   /* 4x8  lower state bound, rANS_byte.h:62 */
#define TF12  12
#define MAXWAY 32

/* ------------------------------------------------------------------ varints (varint.h:85-160) */
int ho_var_put_u32(uint8_t *p, uint32_t v) {
    int n = 1;
    for (uint32_t t = v >> 7; t; t >>= 7) n++;
    for (int k = n - 1; k >= 0; k--)
        *p++ = (uint8_t)(((v >> (7 * k)) & 0x7f) | (k ? 0x80 : 0));
    return n;
}

int ho_var_get_u32(const uint8_t *p, const uint8_t *end, uint32_t *v) {
    uint32_t x = 0;
    int n = 0;
    *v = 0;
    if (p >= end) return 0;
    for (;;) {
        uint8_t c = p[n++];
        x = (x << 7) | (c & 0x7f);
        if (!(c & 0x80) || p + n >= end) break;
    }
    *v = x;
    return n;
}

/* ------------------------------------------------------------------ small helpers */
/* round2, rANS_static4x16pr.c:105-114 (0 -> 0, powers of two map to themselves) */
static uint32_t pow2_ceil(uint32_t v) {
    if (v == 0) return 0;
    uint32_t p = 1;
    while (p < v && p) p <<= 1;
    return p;
}

/* normalise_freq, rANS_static4x16pr.c:116-163.  Scales the non-zero counts of F (sum `total`) so
 * that they sum to `target`; never lets a present symbol drop to zero. */
static int scale_freqs(uint32_t F[256], uint32_t total, uint32_t target) {
    if (!total) return 0;
    int sum_in = (int)total;
    int big_at = 0;
    for (int attempt = 0;; attempt++) {
        uint64_t mul = ((uint64_t)target << 31) / (uint64_t)(int64_t)sum_in + (uint64_t)((1 << 30) / sum_in);
        uint32_t big = 0;
        int sum = 0;
        big_at = 0;
        for (int j = 0; j < 256; j++) {
            if (!F[j]) continue;
            if (big < F[j]) { big = F[j]; big_at = j; }     /* first index holding the maximum */
            uint32_t s = (uint32_t)(((uint64_t)F[j] * mul) >> 31);
            F[j] = s ? s : 1;
            sum += (int)F[j];
        }
        int slack = (int)(target - (uint32_t)sum);
        if (slack > 0) { F[big_at] += (uint32_t)slack; break; }
        if (slack == 0) break;
        uint32_t need = (uint32_t)(-slack);
        if (F[big_at] > need && (attempt == 1 || F[big_at] / 2 >= need)) {
            F[big_at] -= need;
            break;
        }
        if (attempt < 1) { sum_in = sum; continue; }        /* one retry from the rescaled counts */
        /* spread the deficit over every symbol that can spare it, in index order */
        slack += (int)F[big_at] - 1;
        F[big_at] = 1;
        for (int j = 0; slack && j < 256; j++) {
            if (F[j] < 2) continue;
            int take = (F[j] > (uint32_t)(-slack)) ? slack : 1 - (int)F[j];
            F[j] = (uint32_t)((int)F[j] + take);
            slack -= take;
        }
        break;
    }
    return F[big_at] > 0 ? 0 : -1;
}

/* normalise_freq_shift, rANS_static4x16pr.c:168-179 */
static void shift_freqs(uint32_t F[256], uint32_t sum, uint32_t target) {
    if (sum == 0 || sum == target) return;
    int sh = 0;
    while (sum < target) { sum *= 2; sh++; }
    for (int j = 0; j < 256; j++) F[j] <<= sh;
}

/* encode_alphabet, rANS_static4x16pr.c:182-206 */
static int put_alphabet(uint8_t *p, const uint32_t present[256]) {
    uint8_t *p0 = p;
    int implied = 0;
    for (int j = 0; j < 256; j++) {
        if (!present[j]) continue;
        if (implied) { implied--; continue; }
        *p++ = (uint8_t)j;
        if (j && present[j - 1]) {
            int e = j + 1;
            while (e < 256 && present[e]) e++;
            implied = e - (j + 1);
            *p++ = (uint8_t)implied;
        }
    }
    *p++ = 0;
    return (int)(p - p0);
}

/* decode_alphabet, rANS_static4x16pr.c:208-255 (the bounds-checked branch) */
static int get_alphabet(const uint8_t *p, const uint8_t *end, uint32_t present[256]) {
    if (p >= end) return 0;
    const uint8_t *p0 = p;
    int implied = 0;
    int j = *p++;
    do {
        present[j] = 1;
        if (p >= end) return 0;
        if (!implied && j + 1 == *p) {
            if (p + 1 >= end) return 0;
            j = *p++;
            implied = *p++;
        } else if (implied) {
            implied--;
            if (++j > 255) return 0;
        } else {
            j = *p++;
        }
    } while (j && p < end);
    return (int)(p - p0);
}

/* encode_freq, rANS_static4x16pr.c:257-269 */
static int put_freqs_o0(uint8_t *p, const uint32_t F[256]) {
    uint8_t *p0 = p;
    p += put_alphabet(p, F);
    for (int j = 0; j < 256; j++)
        if (F[j]) p += ho_var_put_u32(p, F[j]);
    return (int)(p - p0);
}

/* decode_freq, rANS_static4x16pr.c:271-289 */
static int get_freqs_o0(const uint8_t *p, const uint8_t *end, uint32_t F[256], uint32_t *sum) {
    if (p >= end) return 0;
    const uint8_t *p0 = p;
    p += get_alphabet(p, end, F);
    uint32_t tot = 0;
    for (int j = 0; j < 256; j++) {
        if (!F[j]) continue;
        p += ho_var_get_u32(p, end, &F[j]);
        tot += F[j];
    }
    *sum = tot;
    return (int)(p - p0);
}

/* encode_freq_d, rANS_static4x16pr.c:295-325: one order-1 row over the order-0 alphabet;
 * a run of k zeros is written as 00 (k-1). */
static int put_freqs_row(uint8_t *p, const uint32_t A[256], const uint32_t F[256]) {
    uint8_t *p0 = p;
    int zrun = 0;
    for (int j = 0; j < 256; j++) {
        if (!A[j]) continue;
        if (F[j]) {
            if (zrun) { *p++ = 0; *p++ = (uint8_t)(zrun - 1); zrun = 0; }
            p += ho_var_put_u32(p, F[j]);
        } else {
            zrun++;
        }
    }
    if (zrun) { *p++ = 0; *p++ = (uint8_t)(zrun - 1); }
    return (int)(p - p0);
}

/* decode_freq_d, rANS_static4x16pr.c:327-358 */
static int get_freqs_row(const uint8_t *p, const uint8_t *end, const uint32_t A[256],
                         uint32_t F[256], uint32_t *total) {
    if (p >= end) return 0;
    const uint8_t *p0 = p;
    uint32_t T = 0;
    int zrun = 0;
    for (int j = 0; j < 256 && p < end; j++) {
        if (!A[j]) continue;
        uint32_t f = 0;
        if (zrun) {
            zrun--;
        } else {
            p += ho_var_get_u32(p, end, &f);
            if (f == 0) {
                if (p >= end) return 0;
                zrun = *p++;
            }
        }
        F[j] = f;
        T += f;
    }
    *total = T;
    return (int)(p - p0);
}

/* ------------------------------------------------------------------ bound (…4x16pr.c:360-372) */
unsigned int ho_compress_bound(unsigned int size, int order) {
    int N = order >> 8;
    if (!N) N = 4;
    order &= 0xff;
    double d = 1.05 * size;
    d += (order == 0) ? (257 * 3 + 4) : (257 * 257 * 3 + 4 + 257 * 3 + 4);
    d += (order & HO_PACK) ? 1 : 0;
    d += (order & HO_RLE) ? (1 + 257 * 3 + 4) : 0;
    d += 20;
    d += (order & HO_STRIPE) ? (1 + 5 * N) : 0;
    int sz = (int)d;
    return (unsigned int)(sz + (sz & 1) + 2);
}

/* ------------------------------------------------------------------ N-way encoder core */
/* A write-backwards cursor: the encoder fills the scratch area from its end (rANS_word.h:281-321
 * emits 16-bit little-endian words below the previous one; :104-116 flushes 4 bytes the same way). */
typedef struct { uint8_t *p; } backw;

static inline void bw_u16(backw *b, uint32_t x) { b->p -= 2; b->p[0] = (uint8_t)x; b->p[1] = (uint8_t)(x >> 8); }
static inline void bw_u32(backw *b, uint32_t x) { b->p -= 4; b->p[0] = (uint8_t)x; b->p[1] = (uint8_t)(x >> 8); b->p[2] = (uint8_t)(x >> 16); b->p[3] = (uint8_t)(x >> 24); }

/* RansEncPutSymbol == RansEncPut (rANS_word.h:93-100, 281-321): renormalise then x = C(s,x). */
static inline uint32_t enc_step(uint32_t x, backw *b, uint32_t start, uint32_t freq, int bits) {
    uint32_t x_max = ((L16 >> bits) << 16) * freq;
    if (x >= x_max) { bw_u16(b, x & 0xffff); x >>= 16; }
    return ((x / freq) << bits) + (x % freq) + start;
}

/* rans_compress_O0_4x16, rANS_static4x16pr.c:379-494, with 4 -> nway. */
int ho_enc_o0(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int nway) {
    if (nway != 4 && nway != 32) return -1;
    if (n == 0) { *out_size = 0; return 0; }               /* :405-406: nothing at all is written */

    uint32_t F[256] = {0};
    for (uint32_t i = 0; i < n; i++) F[in[i]]++;            /* hist8, utils.h:81-102 */

    uint32_t target = pow2_ceil(n);
    if (target > (1u << TF12)) target = 1u << TF12;
    if (scale_freqs(F, n, target) < 0) return -1;           /* :412-419 */
    int tab = put_freqs_o0(out, F);                         /* :422 */
    if (scale_freqs(F, target, 1u << TF12) < 0) return -1;  /* :426 */

    uint32_t C[256];
    for (uint32_t j = 0, x = 0; j < 256; j++) { C[j] = x; x += F[j]; }

    uint32_t cap = ho_compress_bound(n, 0) + 4 * MAXWAY;
    uint8_t *scratch = malloc(cap);
    if (!scratch) return -1;
    backw b = { scratch + cap };
    uint32_t R[MAXWAY];
    for (int z = 0; z < nway; z++) R[z] = L16;

    /* symbol i belongs to state i % nway; encode from the last symbol to the first (:442-480) */
    for (uint32_t i = n; i-- > 0;) {
        int z = (int)(i % (uint32_t)nway);
        R[z] = enc_step(R[z], &b, C[in[i]], F[in[i]], TF12);
    }
    for (int z = nway - 1; z >= 0; z--) bw_u32(&b, R[z]);   /* :482-485 */

    uint32_t body = (uint32_t)(scratch + cap - b.p);
    memcpy(out + tab, b.p, body);
    *out_size = (uint32_t)tab + body;
    free(scratch);
    return 0;
}

/* fast_log, rANS_static4x16pr.c:620-623 */
static double approx_log(double a) {
    union { double d; long long x; } u = { a };
    return (double)(u.x - 4606921278410026770LL) * 1.539095918623324e-16;
}

/* compute_shift, rANS_static4x16pr.c:629-691.  Same doubles, same accumulation order. */
static int choose_shift(const uint32_t A[256], uint32_t (*F)[256], const uint32_t T[256], int S[256]) {
    double e10 = 0, e12 = 0;
    int max_tot = 0;
    for (int i = 0; i < 256; i++) {
        if (!A[i]) continue;
        int max_val = (int)pow2_ceil(T[i]);
        int ns = 0, sm10 = 0, sm12 = 0;
        for (int j = 0; j < 256; j++) {
            if (F[i][j] && (uint32_t)max_val / F[i][j] > 1024) sm10++;
            if (F[i][j] && (uint32_t)max_val / F[i][j] > 4096) sm12++;
        }
        double l10 = log((double)(1024 + sm10));
        double l12 = log((double)(4096 + sm12));
        for (int j = 0; j < 256; j++) {
            if (!F[i][j]) continue;
            ns++;
            int x = (int)((double)1024 * F[i][j] / T[i]);
            e10 -= F[i][j] * (approx_log(x > 1 ? x : 1) - l10);
            x = (int)((double)4096 * F[i][j] / T[i]);
            e12 -= F[i][j] * (approx_log(x > 1 ? x : 1) - l12);
            e10 += 4;
            e12 += 6;
        }
        if (ns < 64 && max_val > 128) max_val /= 2;
        if (max_val > 1024) max_val /= 2;
        if (max_val > 4096) max_val = 4096;
        S[i] = max_val;
        if (max_tot < max_val) max_tot = max_val;
    }
    return (e10 / e12 < 1.01 || max_tot <= 1024) ? 10 : 12;
}

/* rans_compress_O1_4x16, rANS_static4x16pr.c:695-847, with 4 -> nway. */
int ho_enc_o1(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int nway) {
    if (nway != 4 && nway != 32) return -1;
    if (n < (uint32_t)nway) return -1;                      /* callers route such inputs to order 0 */

    uint32_t (*F)[256] = calloc(256, sizeof(*F));
    uint16_t (*Cm)[256] = calloc(256, sizeof(*Cm));          /* cumulative starts per context */
    uint32_t T[256] = {0}, A[256] = {0};
    int S[256] = {0};
    uint8_t *scratch = NULL;
    int rc = -1;
    if (!F || !Cm) goto done;

    /* hist1_4, utils.h:137-202: every adjacent pair of the whole buffer, first context 0 */
    {
        uint8_t prev = 0;
        for (uint32_t i = 0; i < n; i++) { F[prev][in[i]]++; T[prev]++; prev = in[i]; }
    }
    uint32_t seg = n / (uint32_t)nway;
    for (int k = 1; k < nway; k++) F[0][in[(uint32_t)k * seg]]++;   /* :720-723 */
    T[0] += (uint32_t)nway - 1;

    for (uint32_t i = 0; i < n; i++) A[in[i]] = 1;          /* present8, utils.h:109-131 */
    A[0] = 1;                                               /* :731 */

    uint8_t *cp = out;
    *cp++ = 0;
    cp += put_alphabet(cp, A);                              /* :732 */
    int shift = choose_shift(A, F, T, S);                   /* :737 */

    for (int i = 0; i < 256; i++) {                         /* :740-764 */
        if (!A[i]) continue;
        uint32_t target = (uint32_t)S[i];
        if (shift == 10 && target > 1024) target = 1024;
        if (scale_freqs(F[i], T[i], target) < 0) goto done;
        cp += put_freqs_row(cp, A, F[i]);
        shift_freqs(F[i], target, 1u << shift);
        uint32_t x = 0;
        for (int j = 0; j < 256; j++) { Cm[i][j] = (uint16_t)x; x += F[i][j]; }
    }
    out[0] = (uint8_t)(shift << 4);

    if (cp - out > 1000) {                                  /* :767-780: try order-0 on the table */
        uint32_t usz = (uint32_t)(cp - (out + 1)), csz = 0;
        uint8_t *ctab = malloc(ho_compress_bound(usz, 0));
        if (ctab && ho_enc_o0(out + 1, usz, ctab, &csz, 4) == 0 && csz + 6 < (uint32_t)(cp - out)) {
            uint8_t *op = out;
            *op++ |= 1;
            op += ho_var_put_u32(op, usz);
            op += ho_var_put_u32(op, csz);
            memcpy(op, ctab, csz);
            cp = op + csz;
        }
        free(ctab);
    }
    uint32_t tab = (uint32_t)(cp - out);

    uint32_t cap = ho_compress_bound(n, 1) + 4 * MAXWAY;
    scratch = malloc(cap);
    if (!scratch) goto done;
    backw b = { scratch + cap };
    uint32_t R[MAXWAY];
    for (int z = 0; z < nway; z++) R[z] = L16;

    /* State k owns in[k*seg .. (k+1)*seg); the last state also owns the tail.  Symbols are coded
     * last-to-first with the preceding byte as context (0 at the start of a segment). :794-834 */
    int last = nway - 1;
    for (uint32_t i = n - 1; i >= (uint32_t)nway * seg; i--) {      /* tail on the last state */
        uint8_t ctx = in[i - 1], s = in[i];
        R[last] = enc_step(R[last], &b, Cm[ctx][s], F[ctx][s], shift);
    }
    for (uint32_t t = seg; t-- > 1;) {
        for (int k = last; k >= 0; k--) {
            uint32_t pos = (uint32_t)k * seg + t;
            uint8_t ctx = in[pos - 1], s = in[pos];
            R[k] = enc_step(R[k], &b, Cm[ctx][s], F[ctx][s], shift);
        }
    }
    for (int k = last; k >= 0; k--) {
        uint8_t s = in[(uint32_t)k * seg];
        R[k] = enc_step(R[k], &b, Cm[0][s], F[0][s], shift);
    }
    for (int k = last; k >= 0; k--) bw_u32(&b, R[k]);

    uint32_t body = (uint32_t)(scratch + cap - b.p);
    memcpy(out + tab, b.p, body);
    *out_size = tab + body;
    rc = 0;
done:
    free(scratch);
    free(F);
    free(Cm);
    return rc;
}

/* ------------------------------------------------------------------ N-way decoder core */
/* RansDecRenorm / RansDecRenormSafe, rANS_word.h:356-410: pull one LE u16 if the state fell
 * below 2^15 and at least two bytes remain. */
static inline uint32_t dec_renorm(uint32_t x, const uint8_t **pp, const uint8_t *end) {
    if (x < L16 && *pp + 1 < end) {
        x = (x << 16) | (uint32_t)((*pp)[0] | ((*pp)[1] << 8));
        *pp += 2;
    }
    return x;
}

static inline uint32_t rd_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* rans_uncompress_O0_4x16, rANS_static4x16pr.c:501-616, with 4 -> nway. */
int ho_dec_o0(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t n, int nway) {
    if (nway != 4 && nway != 32) return -1;
    if (in_size < 4u * (uint32_t)nway) return -1;           /* :503 */
    if (n >= INT_MAX) return -1;                            /* :506 */
    const uint8_t *cp = in, *end = in + in_size;

    uint32_t F[256] = {0}, sum = 0;
    int used = get_freqs_o0(cp, end - 8, F, &sum);          /* :516,530 (cp_end = in+in_size-8) */
    if (!used) return -1;
    cp += used;
    shift_freqs(F, sum, 1u << TF12);                        /* :535 */

    static const uint32_t M = 1u << TF12;
    uint8_t  sym[1 << TF12];
    uint16_t frq[1 << TF12], off[1 << TF12];
    uint32_t x = 0;
    for (int j = 0; j < 256; j++) {                         /* :538-549 */
        if (!F[j]) continue;
        if (F[j] > M - x) return -1;
        for (uint32_t y = 0; y < F[j]; y++) { sym[x + y] = (uint8_t)j; frq[x + y] = (uint16_t)F[j]; off[x + y] = (uint16_t)y; }
        x += F[j];
    }
    if (x != M) return -1;                                  /* :551 */
    if (cp + 4 * nway > end) return -1;                     /* :554 */

    uint32_t R[MAXWAY];
    for (int z = 0; z < nway; z++) { R[z] = rd_u32(cp); cp += 4; if (R[z] < L16) return -1; }

    for (uint32_t i = 0; i < n; i++) {                      /* :574-607 */
        int z = (int)(i % (uint32_t)nway);
        uint32_t m = R[z] & (M - 1);
        out[i] = sym[m];
        R[z] = frq[m] * (R[z] >> TF12) + off[m];
        R[z] = dec_renorm(R[z], &cp, end);
    }
    return 0;
}

/* rans_uncompress_O1_4x16, rANS_static4x16pr.c:870-1130, with 4 -> nway. */
int ho_dec_o1(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t n, int nway) {
    if (nway != 4 && nway != 32) return -1;
    if (in_size < 4u * (uint32_t)nway) return -1;           /* :872 */
    if (n >= INT_MAX) return -1;
    const uint8_t *cp = in, *end = in + in_size;
    int rc = -1;
    uint8_t *tab_raw = NULL;
    const uint8_t *after_tab = NULL, *tab_end = end;

    uint32_t shift = *cp >> 4;                              /* :943 */
    /* The reference's loops are specialised for 12 and 10 only (:1027,:1071); nothing else
     * occurs in valid streams. */
    if (shift != 10 && shift != 12) return -1;
    uint32_t M = 1u << shift;
    uint8_t  (*sym)[1 << 12] = malloc(256 * sizeof(*sym));
    uint16_t (*frq)[256] = calloc(256, sizeof(*frq));
    uint16_t (*cum)[256] = calloc(256, sizeof(*cum));
    if (!sym || !frq || !cum) goto done;
    memset(sym, 0, 256 * sizeof(*sym));

    if (*cp++ & 1) {                                        /* :944-955 compressed table */
        uint32_t usz, csz;
        cp += ho_var_get_u32(cp, end, &usz);
        cp += ho_var_get_u32(cp, end, &csz);
        if ((int64_t)csz >= (int64_t)(end - cp) - 16) goto done;        /* :948 */
        after_tab = cp + csz;
        tab_raw = malloc(usz ? usz : 1);
        if (!tab_raw || ho_dec_o0(cp, csz, tab_raw, usz, 4) < 0) goto done;
        cp = tab_raw;
        tab_end = tab_raw + usz;
    }

    uint32_t A[256] = {0};
    int used = get_alphabet(cp, tab_end, A);                /* :959 */
    if (!used) goto done;
    cp += used;
    if (cp >= tab_end) goto done;

    for (int i = 0; i < 256; i++) {                         /* :967-998 */
        if (!A[i]) continue;
        uint32_t F[256] = {0}, T = 0;
        used = get_freqs_row(cp, tab_end, A, F, &T);
        if (!used) goto done;
        cp += used;
        if (!T) continue;
        shift_freqs(F, T, M);
        uint32_t x = 0;
        for (int j = 0; j < 256; j++) {
            if (!F[j]) continue;
            if (F[j] > M - x) goto done;
            memset(&sym[i][x], j, F[j]);
            frq[i][j] = (uint16_t)F[j];
            cum[i][j] = (uint16_t)x;
            x += F[j];
        }
        if (x != M) goto done;
    }
    if (after_tab) cp = after_tab;
    if (cp + 4 * nway > end) goto done;                     /* :1005 */

    uint32_t R[MAXWAY], pos[MAXWAY];
    uint8_t ctx[MAXWAY] = {0};
    uint32_t seg = n / (uint32_t)nway;
    for (int z = 0; z < nway; z++) {
        R[z] = rd_u32(cp); cp += 4;
        if (R[z] < L16) goto done;
        pos[z] = (uint32_t)z * seg;
    }
    for (uint32_t t = 0; t < seg; t++) {                    /* :1031-1060 */
        for (int z = 0; z < nway; z++) {
            uint32_t m = R[z] & (M - 1);
            uint8_t c = sym[ctx[z]][m];
            R[z] = frq[ctx[z]][c] * (R[z] >> shift) + m - cum[ctx[z]][c];
            out[pos[z]++] = ctx[z] = c;
        }
        for (int z = 0; z < nway; z++) R[z] = dec_renorm(R[z], &cp, end);
    }
    int z = nway - 1;                                       /* :1063-1070 tail on the last state */
    for (; pos[z] < n; pos[z]++) {
        uint32_t m = R[z] & (M - 1);
        uint8_t c = sym[ctx[z]][m];
        out[pos[z]] = c;
        R[z] = frq[ctx[z]][c] * (R[z] >> shift) + m - cum[ctx[z]][c];
        R[z] = dec_renorm(R[z], &cp, end);
        ctx[z] = c;
    }
    rc = 0;
done:
    free(sym); free(frq); free(cum); free(tab_raw);
    return rc;
}

/* ------------------------------------------------------------------ PACK (pack.c) */
/* hts_pack, pack.c:56-151.  meta = [nsym][symbols...]; codes LSB-first. Returns 0, or -1 when the
 * alphabet has more than 16 symbols (the caller then drops X_PACK, except for the 256-symbol wrap). */
int ho_pack(const uint8_t *in, uint32_t n, uint8_t *meta, int *meta_len, uint8_t *out, uint32_t *out_len) {
    int code[256], nsym = 0;
    uint8_t seen[256] = {0};
    for (uint32_t i = 0; i < n; i++) seen[in[i]] = 1;
    for (int s = 0; s < 256; s++)
        if (seen[s]) { code[s] = nsym++; meta[nsym] = (uint8_t)s; }
    meta[0] = (uint8_t)nsym;                                /* 256 wraps to 0 (pack.c:73) */
    if (nsym > 16) {                                        /* pack.c:77-84 */
        *meta_len = 1;
        memcpy(out, in, n);
        *out_len = n;
        return 0;
    }
    *meta_len = nsym + 1;
    int per = nsym > 4 ? 2 : nsym > 2 ? 4 : nsym > 1 ? 8 : 0;
    if (!per) { *out_len = 0; return 0; }
    int bits = 8 / per;
    uint32_t o = 0;
    for (uint32_t i = 0; i < n; i += (uint32_t)per) {
        unsigned v = 0;
        for (int k = 0; k < per && i + (uint32_t)k < n; k++) v |= (unsigned)code[in[i + (uint32_t)k]] << (k * bits);
        out[o++] = (uint8_t)v;
    }
    *out_len = o;
    return 0;
}

/* hts_unpack_meta, pack.c:165-198.  Returns bytes consumed (0 = failure). */
static int unpack_meta(const uint8_t *d, uint32_t len, uint8_t map[16], int *per) {
    if (!len) return 0;
    unsigned ns = d[0] ? d[0] : 256;
    if (ns <= 1) *per = 0; else if (ns <= 2) *per = 8; else if (ns <= 4) *per = 4; else if (ns <= 16) *per = 2;
    else { *per = 1; return 1; }
    if (len <= 1) return 0;
    unsigned c = 0, j = 1;
    do { map[c++] = d[j++]; } while (c < ns && j < len);
    return c < ns ? 0 : (int)j;
}

/* hts_unpack, pack.c:211-348 */
static int unpack(const uint8_t *d, uint64_t len, uint8_t *out, uint64_t out_len, int per, const uint8_t map[16]) {
    if (per == 1) { memcpy(out, d, len); return 0; }
    if (per == 0) { memset(out, map[0], out_len); return 0; }
    if (per != 2 && per != 4 && per != 8) return -1;
    if ((out_len + (uint64_t)per - 1) / (uint64_t)per > len) return -1;
    int bits = 8 / per;
    unsigned mask = (1u << bits) - 1;
    for (uint64_t i = 0; i < out_len; i++)
        out[i] = map[(d[i / (uint64_t)per] >> ((i % (uint64_t)per) * (uint64_t)bits)) & mask];
    return 0;
}

/* ------------------------------------------------------------------ RLE (rle.c) */
/* rle_find_syms + rle_encode, rle.c:48-138.  A symbol gets run-length treatment when it repeats
 * its predecessor more often than not.  lits gets one byte per run (RLE symbols) or per byte
 * (others); runs gets varint(run_length-1) per RLE-symbol literal. */
int ho_rle_encode(const uint8_t *in, uint32_t n, uint8_t *runs, uint32_t *runs_len,
                  uint8_t *syms, int *nsyms, uint8_t *lits, uint32_t *lits_len) {
    int64_t score[256] = {0};
    int prev = -1;
    for (uint32_t i = 0; i < n; i++) { score[in[i]] += (in[i] == prev) ? 1 : -1; prev = in[i]; }
    int ns = 0;
    for (int s = 0; s < 256; s++) if (score[s] > 0) syms[ns++] = (uint8_t)s;
    *nsyms = ns;

    uint32_t k = 0, r = 0;
    for (uint32_t i = 0; i < n;) {
        uint8_t s = in[i];
        lits[k++] = s;
        if (score[s] > 0) {
            uint32_t e = i + 1;
            while (e < n && in[e] == s) e++;
            r += (uint32_t)ho_var_put_u32(runs + r, e - i - 1);
            i = e;
        } else {
            i++;
        }
    }
    *runs_len = r;
    *lits_len = k;
    return 0;
}

/* rle_decode, rle.c:142-187 */
static int rle_expand(const uint8_t *lit, uint64_t lit_len, const uint8_t *run, uint64_t run_len,
                      const uint8_t *syms, int nsyms, uint8_t *out, uint64_t *out_len) {
    uint8_t is_rle[256] = {0};
    for (int j = 0; j < nsyms; j++) is_rle[syms[j]] = 1;
    const uint8_t *run_end = run + run_len;
    uint8_t *o = out, *o_end = out + *out_len;
    for (uint64_t i = 0; i < lit_len; i++) {
        if (o >= o_end) return -1;
        uint8_t b = lit[i];
        if (!is_rle[b]) { *o++ = b; continue; }
        uint32_t extra;
        run += ho_var_get_u32(run, run_end, &extra);
        if (extra) {
            if (o + extra >= o_end) return -1;              /* rle.c:172 */
            memset(o, b, (size_t)extra + 1);
            o += (size_t)extra + 1;
        } else {
            *o++ = b;
        }
    }
    *out_len = (uint64_t)(o - out);
    return 0;
}

/* ------------------------------------------------------------------ container */
static int nway_of(int flags) { return (flags & HO_X32) ? 32 : 4; }

/* rans_compress_to_4x16, rANS_static4x16pr.c:1138-1345 */
int ho_compress(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int order) {
    if (n <= 20) order &= ~HO_STRIPE;                       /* :1151 */
    uint32_t hdr;

    if (order & HO_STRIPE) {                                /* :1154-1216 */
        int N = order >> 8;
        if (N == 0) N = 4;
        if (N > 255) return -1;
        uint8_t *tr = malloc(n);
        uint32_t cap = ho_compress_bound(n, order) + 64;
        uint8_t *best = malloc(cap), *trial = malloc(cap), *payload = malloc(cap + 5u * 256);
        if (!tr || !best || !trial || !payload) { free(tr); free(best); free(trial); free(payload); return -1; }
        uint32_t len[256], at[256];
        for (int j = 0; j < N; j++) {
            len[j] = n / (uint32_t)N + ((n % (uint32_t)N) > (uint32_t)j);
            at[j] = j ? at[j - 1] + len[j - 1] : 0;
        }
        for (uint32_t i = 0; i < n; i++) tr[at[i % (uint32_t)N] + i / (uint32_t)N] = in[i];   /* :1168-1180 */

        out[0] = (uint8_t)(order & ~HO_NOSZ);               /* :1185 */
        hdr = 1;
        hdr += (uint32_t)ho_var_put_u32(out + hdr, n);
        out[hdr++] = (uint8_t)N;
        uint32_t pay = 0;
        static const int cand[4] = { 1, 64, 128, 0 };       /* :1192 */
        int rc = 0;
        for (int j = 0; j < N && !rc; j++) {
            uint32_t best_sz = n + 10, got = 0;
            int have = 0;
            for (int c = 0; c < 4; c++) {
                if ((order & cand[c]) != cand[c]) continue;
                if (ho_compress(tr + at[j], len[j], trial, &got, cand[c] | HO_NOSZ | (order & HO_X32)) < 0) { rc = -1; break; }
                if (best_sz > got) { best_sz = got; memcpy(best, trial, got); have = 1; }   /* strict '<' wins */
            }
            if (!have) { rc = -1; break; }
            memcpy(payload + pay, best, best_sz);
            pay += best_sz;
            hdr += (uint32_t)ho_var_put_u32(out + hdr, best_sz);
        }
        if (!rc) { memcpy(out + hdr, payload, pay); *out_size = hdr + pay; }
        free(tr); free(best); free(trial); free(payload);
        return rc;
    }

    if (order & HO_CAT) {                                   /* :1218-1225 */
        out[0] = HO_CAT;
        hdr = 1 + (uint32_t)ho_var_put_u32(out + 1, n);
        memcpy(out + hdr, in, n);
        *out_size = hdr + n;
        return 0;
    }

    int do_pack = order & HO_PACK, do_rle = order & HO_RLE, nosz = order & HO_NOSZ;
    int nway = nway_of(order);
    out[0] = (uint8_t)order;                                /* :1231 */
    hdr = 1;
    if (!nosz) hdr += (uint32_t)ho_var_put_u32(out + 1, n);
    int o1 = order & HO_ORDER1;

    uint8_t *packed = NULL, *lits = NULL;
    const uint8_t *cur = in;
    uint32_t cur_n = n;
    int rc = -1;

    if (do_pack && cur_n) {                                 /* :1244-1267 */
        int mlen;
        uint32_t plen;
        packed = malloc((size_t)cur_n + 1);
        if (!packed) goto done;
        ho_pack(cur, cur_n, out + hdr, &mlen, packed, &plen);
        if (mlen == 1 && out[hdr] > 16) {                   /* 17..255 symbols: give up on PACK */
            out[0] &= ~HO_PACK;
        } else {                                            /* also taken for the 256-symbol wrap */
            cur = packed; cur_n = plen;
            hdr += (uint32_t)mlen;
            hdr += (uint32_t)ho_var_put_u32(out + hdr, cur_n);
        }
    } else if (do_pack) {
        out[0] &= ~HO_PACK;
    }

    if (do_rle && cur_n) {                                  /* :1269-1319 */
        uint8_t *meta = malloc((size_t)cur_n * 5 + 600), rsyms[256];
        lits = malloc((size_t)cur_n + 1);
        if (!meta || !lits) { free(meta); goto done; }
        int nrs;
        uint32_t runs_len, lit_len;
        ho_rle_encode(cur, cur_n, meta + 257, &runs_len, rsyms, &nrs, lits, &lit_len);
        uint32_t meta_len = 1 + (uint32_t)nrs + runs_len;
        uint8_t *m = meta + 257 - 1 - nrs;                  /* [nsyms][syms...][runs...] */
        m[0] = (uint8_t)nrs;
        memcpy(m + 1, rsyms, (size_t)nrs);
        if ((double)((uint64_t)lit_len + meta_len) >= .99 * cur_n) {    /* :1287 */
            out[0] &= ~HO_RLE;
        } else {
            uint32_t csz = 0;
            uint8_t *cm = malloc(ho_compress_bound(meta_len, 0) + 4 * MAXWAY);
            if (!cm || ho_enc_o0(m, meta_len, cm, &csz, nway) < 0) { free(cm); free(meta); goto done; }
            if (csz < meta_len) {                           /* :1299-1301 */
                hdr += (uint32_t)ho_var_put_u32(out + hdr, meta_len * 2);
                hdr += (uint32_t)ho_var_put_u32(out + hdr, lit_len);
                hdr += (uint32_t)ho_var_put_u32(out + hdr, csz);
                memcpy(out + hdr, cm, csz);
                hdr += csz;
            } else {                                        /* :1302-1308 raw meta */
                hdr += (uint32_t)ho_var_put_u32(out + hdr, meta_len * 2 + 1);
                hdr += (uint32_t)ho_var_put_u32(out + hdr, lit_len);
                memcpy(out + hdr, m, meta_len);
                hdr += meta_len;
            }
            free(cm);
            cur = lits; cur_n = lit_len;
        }
        free(meta);
    } else if (do_rle) {
        out[0] &= ~HO_RLE;
    }

    /* :1322-1325 (and its N-way analogue: a segment must hold at least one symbol) */
    if (o1 && (cur_n < 8 || cur_n < (uint32_t)nway)) { out[0] &= ~1; o1 = 0; }

    uint32_t body = 0;
    if ((o1 ? ho_enc_o1(cur, cur_n, out + hdr, &body, nway) : ho_enc_o0(cur, cur_n, out + hdr, &body, nway)) < 0)
        goto done;
    if (body >= cur_n) {                                    /* :1332-1337 */
        out[0] &= ~3;
        out[0] |= HO_CAT | nosz;
        memcpy(out + hdr, cur, cur_n);
        body = cur_n;
    }
    *out_size = hdr + body;
    rc = 0;
done:
    free(packed);
    free(lits);
    return rc;
}

int ho_peek_size(const uint8_t *in, uint32_t in_size, uint32_t *ulen) {
    if (in_size < 2 || (in[0] & HO_NOSZ)) return -1;
    return ho_var_get_u32(in + 1, in + in_size, ulen) ? 0 : -1;
}

/* rans_uncompress_to_4x16, rANS_static4x16pr.c:1352-1636 */
int ho_uncompress(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size) {
    const uint8_t *end = in + in_size;
    if (in_size == 0) return -1;

    if (in[0] & HO_STRIPE) {                                /* :1360-1433 */
        uint32_t ulen, pos = 1;
        pos += (uint32_t)ho_var_get_u32(in + pos, end, &ulen);
        if (pos >= in_size) return -1;
        uint32_t N = in[pos++];
        if (N == 0) return -1;
        if (ulen != *out_size) return -1;                   /* :1379 exact size required */
        uint32_t clen[256], ul[256], at[256];
        uint64_t ctot = 0;
        for (uint32_t j = 0; j < N; j++) {
            ul[j] = ulen / N + ((ulen % N) > j);
            at[j] = j ? at[j - 1] + ul[j - 1] : 0;
            pos += (uint32_t)ho_var_get_u32(in + pos, end, &clen[j]);
            ctot += clen[j];
            if (pos > in_size || clen[j] > in_size || clen[j] < 1) return -1;
        }
        if (pos + ctot > in_size) return -1;
        in_size = (uint32_t)(pos + ctot);
        uint8_t *parts = malloc(ulen ? ulen : 1);
        if (!parts) return -1;
        for (uint32_t j = 0; j < N; j++) {
            uint32_t got = ul[j];
            if (in_size < pos || ho_uncompress(in + pos, in_size - pos, parts + at[j], &got) < 0 || got != ul[j]) {
                free(parts);
                return -1;
            }
            pos += clen[j];
        }
        for (uint32_t i = 0; i < ulen; i++) out[i] = parts[at[i % N] + i / N];  /* unstripe, utils.h:41-73 */
        free(parts);
        *out_size = ulen;
        return 0;
    }

    int flags = *in++; in_size--;
    int do_pack = flags & HO_PACK, do_rle = flags & HO_RLE, do_cat = flags & HO_CAT, nosz = flags & HO_NOSZ;
    int nway = nway_of(flags), o1 = flags & HO_ORDER1;
    uint32_t osz;
    if (!nosz) {
        int s = ho_var_get_u32(in, end, &osz);
        in += s; in_size -= (uint32_t)s;
    } else {
        osz = *out_size;
    }
    if (*out_size < osz) return -1;                         /* :1464 */
    *out_size = osz;

    uint8_t *tmp = NULL, *meta_buf = NULL;
    const uint8_t *meta = NULL;
    int rc = -1;
    uint8_t *t1 = out, *t2 = out, *t3 = out;                /* rANS -> t1, unRLE -> t2, unpack -> t3 */
    uint32_t t1_size = osz;
    if (do_pack || do_rle) {                                /* :1498-1520 */
        tmp = malloc(osz ? osz : 1);
        if (!tmp) return -1;
        if (do_pack && do_rle) { t1 = out; t2 = tmp; t3 = out; }
        else if (do_pack)      { t1 = tmp; t2 = tmp; t3 = out; }
        else                   { t1 = tmp; t2 = out; t3 = out; }
    }

    uint8_t map[16] = {0};
    int per = 0;
    uint64_t unpacked = 0;
    if (do_pack) {                                          /* :1527-1545 */
        int c = unpack_meta(in, in_size, map, &per);
        if (!c) goto done;
        unpacked = osz;
        in += c; in_size -= (uint32_t)c;
        uint32_t psz;
        int s = ho_var_get_u32(in, end, &psz);
        in += s; in_size -= (uint32_t)s;
        if (psz > t1_size) goto done;
        t1_size = psz;
    }

    uint32_t u_meta = 0;
    if (do_rle) {                                           /* :1549-1572 */
        uint32_t c_meta, rle_len, s;
        s = (uint32_t)ho_var_get_u32(in, end, &u_meta);
        s += (uint32_t)ho_var_get_u32(in + s, end, &rle_len);
        if (rle_len > t1_size) goto done;
        if (u_meta & 1) {
            meta = in + s;
            u_meta = (u_meta / 2 > (uint64_t)(end - meta)) ? (uint32_t)(end - meta) : u_meta / 2;
            c_meta = u_meta;
        } else {
            s += (uint32_t)ho_var_get_u32(in + s, end, &c_meta);
            u_meta /= 2;
            if (s > in_size) goto done;
            meta_buf = malloc(u_meta ? u_meta : 1);
            if (!meta_buf || ho_dec_o0(in + s, in_size - s, meta_buf, u_meta, nway) < 0) goto done;
            meta = meta_buf;
        }
        if ((uint64_t)c_meta + s > in_size) goto done;
        in += c_meta + s; in_size -= c_meta + s;
        t1_size = rle_len;
    }

    if (in_size) {                                          /* :1577-1595 */
        if (do_cat) {
            if (t1_size > in_size || t1_size > *out_size) goto done;
            memcpy(t1, in, t1_size);
        } else if ((o1 ? ho_dec_o1(in, in_size, t1, t1_size, nway) : ho_dec_o0(in, in_size, t1, t1_size, nway)) < 0) {
            goto done;
        }
    } else {
        t1_size = 0;
    }
    uint64_t t2_size = t1_size, t3_size = t1_size;

    if (do_rle) {                                           /* :1598-1613 */
        if (u_meta == 0) goto done;
        int nrs = meta[0] ? meta[0] : 256;
        if (u_meta < 1u + (uint32_t)nrs) goto done;
        uint64_t un = *out_size;
        if (rle_expand(t1, t1_size, meta + 1 + nrs, u_meta - (1u + (uint32_t)nrs), meta + 1, nrs, t2, &un) < 0) goto done;
        t3_size = t2_size = un;
    }
    if (do_pack) {                                          /* :1614-1623 */
        if (per == 1) unpacked = t2_size;
        if (unpack(t2, t2_size, t3, unpacked, per, map) < 0) goto done;
        t3_size = unpacked;
    }
    (void)t3;
    *out_size = (uint32_t)t3_size;
    rc = 0;
done:
    free(tmp);
    free(meta_buf);
    return rc;
}

/* ------------------------------------------------------------------ legacy rANS 4x8 decode */
/* RansDecRenorm / RansDecRenormSafe, rANS_byte.h:435-551: up to two single-byte refills. */
static inline uint32_t dec_renorm8(uint32_t x, const uint8_t **pp, const uint8_t *end) {
    for (int k = 0; k < 2 && x < L8 && *pp < end; k++) x = (x << 8) | *(*pp)++;
    return x;
}

/* One "sym [run] freq" table of the 4x8 format (rANS_static.c:271-303 / :748-813); fills
 * F[sym] in stream order and returns the frequency sum, or -1. */
static int get_table_4x8(const uint8_t **pp, const uint8_t *end, uint32_t F[256], int zero_is_4096) {
    const uint8_t *cp = *pp;
    int run = 0, x = 0;
    int j = *cp++;
    do {
        if (cp > end - 16) return -1;
        int f = *cp++;
        if (f >= 128) f = ((f & 127) << 8) | *cp++;
        if (!f && zero_is_4096) f = 4096;
        if (x + f > 4096) return -1;
        F[j] = (uint32_t)f;
        x += f;
        if (!run && j + 1 == *cp) { j = *cp++; run = *cp++; }
        else if (run) { run--; if (++j > 255) return -1; }
        else j = *cp++;
    } while (j);
    *pp = cp;
    return x;
}

/* rans_uncompress_O0, rANS_static.c:225-363 */
static int dec4x8_o0(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size) {
    if (in_size < 26 || in[0] != 0) return -1;
    uint32_t clen = rd_u32(in + 1), n = rd_u32(in + 5);
    if (clen != in_size - 9 || n >= INT_MAX || n > *out_size) return -1;
    const uint8_t *cp = in + 9, *end = in + in_size;
    uint32_t F[256] = {0};
    int tot = get_table_4x8(&cp, end, F, 0);
    if (tot < 4095 || tot > 4096) return -1;                /* :305 */
    uint8_t sym[4096];
    uint16_t frq[4096], off[4096];
    /* cumulative order is ascending symbol order because the table is written ascending */
    uint32_t x = 0;
    memset(sym, 0, sizeof(sym)); memset(frq, 0, sizeof(frq)); memset(off, 0, sizeof(off));
    for (int j = 0; j < 256; j++)
        for (uint32_t y = 0; y < F[j]; y++, x++) { sym[x] = (uint8_t)j; frq[x] = (uint16_t)F[j]; off[x] = (uint16_t)y; }
    if (cp > end - 16) return -1;
    uint32_t R[4];
    for (int z = 0; z < 4; z++) { R[z] = rd_u32(cp); cp += 4; if (R[z] < L8) return -1; }
    uint32_t n4 = n & ~3u;
    for (uint32_t i = 0; i < n4; i += 4) {                  /* :318-344 */
        for (int z = 0; z < 4; z++) {
            uint32_t m = R[z] & 4095;
            out[i + (uint32_t)z] = sym[m];
            R[z] = frq[m] * (R[z] >> 12) + off[m];
        }
        for (int z = 0; z < 4; z++) R[z] = dec_renorm8(R[z], &cp, end);
    }
    for (uint32_t z = 0; z < (n & 3); z++) out[n4 + z] = sym[R[z] & 4095];   /* :346-355 peek only */
    *out_size = n;
    return 0;
}

/* rans_uncompress_O1, rANS_static.c:676-922 */
static int dec4x8_o1(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size) {
    if (in_size < 27 || in[0] != 1) return -1;
    uint32_t clen = rd_u32(in + 1), n = rd_u32(in + 5);
    if (clen != in_size - 9 || n >= INT_MAX || n > *out_size) return -1;
    const uint8_t *cp = in + 9, *end = in + in_size;
    int rc = -1;
    uint8_t  (*sym)[4096] = calloc(256, sizeof(*sym));
    uint16_t (*frq)[256] = calloc(256, sizeof(*frq));
    uint16_t (*cum)[256] = calloc(256, sizeof(*cum));
    if (!sym || !frq || !cum) goto done;

    int run = 0, i = *cp++;
    do {                                                    /* :748-813 */
        uint32_t F[256] = {0};
        int tot = get_table_4x8(&cp, end, F, 1);
        if (tot < 4095 || tot > 4096) goto done;
        /* inner table is written in ascending symbol order, so cumulative == ascending */
        uint32_t x = 0;
        for (int j = 0; j < 256; j++) {
            if (!F[j]) continue;
            frq[i][j] = (uint16_t)F[j];
            cum[i][j] = (uint16_t)x;
            memset(&sym[i][x], j, F[j]);
            x += F[j];
        }
        if (x < 4096) sym[i][x] = sym[i][x - 1];            /* :799-800 (valid streams only) */
        if (!run && i + 1 == *cp) { i = *cp++; run = *cp++; }
        else if (run) { run--; if (++i > 255) goto done; }
        else i = *cp++;
    } while (i);

    if (cp > end - 16) goto done;
    uint32_t R[4], pos[4], seg = n >> 2;
    uint8_t ctx[4] = {0, 0, 0, 0};
    for (int z = 0; z < 4; z++) { R[z] = rd_u32(cp); cp += 4; if (R[z] < L8) goto done; pos[z] = (uint32_t)z * seg; }
    for (uint32_t t = 0; t < seg; t++) {                    /* :850-898 */
        for (int z = 0; z < 4; z++) {
            uint32_t m = R[z] & 4095;
            uint8_t c = sym[ctx[z]][m];
            out[pos[z]++] = c;
            R[z] = frq[ctx[z]][c] * (R[z] >> 12) + m - cum[ctx[z]][c];
            ctx[z] = c;
        }
        for (int z = 0; z < 4; z++) R[z] = dec_renorm8(R[z], &cp, end);
    }
    for (; pos[3] < n; pos[3]++) {                          /* :901-909 */
        uint32_t m = R[3] & 4095;
        uint8_t c = sym[ctx[3]][m];
        out[pos[3]] = c;
        R[3] = frq[ctx[3]][c] * (R[3] >> 12) + m - cum[ctx[3]][c];
        R[3] = dec_renorm8(R[3], &cp, end);
        ctx[3] = c;
    }
    *out_size = n;
    rc = 0;
done:
    free(sym); free(frq); free(cum);
    return rc;
}

/* rans_uncompress, rANS_static.c:934-943 */
int ho_uncompress_4x8(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size) {
    if (in_size < 9) return -1;
    return in[0] ? dec4x8_o1(in, in_size, out, out_size) : dec4x8_o0(in, in_size, out, out_size);
}

/* ------------------------------------------------------------------ rANS 4x8 encode (legacy) */
/* the CRAM 3.0 encoder, restated.  PINNED: reproduces the 8
 * tests/dat/r4x8 golden streams and libref.so's rans_compress byte for byte (tests/test_oracle.py). */

/* RansEncPutSymbol, rANS_byte.h:281-315: up to two renormalisation bytes (low byte first, each
 * written below the previous one), then x = C(s,x) with M = 4096. */
static inline uint32_t enc_step8(uint32_t x, backw *b, uint32_t start, uint32_t freq) {
    uint32_t x_max = ((L8 >> TF12) << 8) * freq;
    if (x >= x_max) { *--b->p = (uint8_t)x; x >>= 8; }
    if (x >= x_max) { *--b->p = (uint8_t)x; x >>= 8; }
    return ((x / freq) << TF12) + (x % freq) + start;
}

/* the "sym [run]" list shared by both table levels, rANS_static.c:139-153 / :491-503 */
static inline uint8_t *put_sym_rle(uint8_t *cp, int j, int *rle, const int present[256]) {
    if (*rle) { (*rle)--; return cp; }
    *cp++ = (uint8_t)j;
    if (j && present[j - 1]) {
        int e = j + 1;
        while (e < 256 && present[e]) e++;
        *rle = e - (j + 1);
        *cp++ = (uint8_t)*rle;
    }
    return cp;
}

static inline uint8_t *put_freq8(uint8_t *cp, int f) {         /* :155-161 */
    if (f < 128) *cp++ = (uint8_t)f;
    else { *cp++ = (uint8_t)(128 | (f >> 8)); *cp++ = (uint8_t)(f & 0xff); }
    return cp;
}

static void put_hdr8(uint8_t *out, int order, uint32_t total, uint32_t n) {   /* :201-213 */
    out[0] = (uint8_t)order;
    for (int k = 0; k < 4; k++) { out[1 + k] = (uint8_t)((total - 9) >> (8 * k)); out[5 + k] = (uint8_t)(n >> (8 * k)); }
}

unsigned int ho_compress_bound_4x8(unsigned int n) {            /* malloc size, :87 / :449 */
    return (unsigned int)(1.05 * n + 257 * 257 * 3 + 9);
}

/* rans_compress_O0, rANS_static.c:85-218 */
static int enc4x8_o0(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size) {
    if (n == 0) return -1;                                  /* the reference divides by in_size (:105) */
    int F[256] = {0}, present[256];
    for (uint32_t i = 0; i < n; i++) F[in[i]]++;
    uint64_t tr = ((uint64_t)4096 << 31) / n + (1 << 30) / n;
    for (;;) {                                              /* :107-131 */
        int fsum = 0, m = 0, M = 0;
        for (int j = 0; j < 256; j++) {
            if (!F[j]) continue;
            if (m < F[j]) { m = F[j]; M = j; }
            if ((F[j] = (int)(((uint64_t)F[j] * tr) >> 31)) == 0) F[j] = 1;
            fsum += F[j];
        }
        fsum++;
        if (fsum < 4096) { F[M] += 4096 - fsum; break; }
        if (fsum - 4096 > F[M] / 2) { tr = 2104533975; continue; }
        F[M] -= fsum - 4096;
        break;
    }
    for (int j = 0; j < 256; j++) present[j] = F[j] != 0;
    uint32_t C[256];
    uint8_t *cp = out + 9;
    int rle = 0;
    for (uint32_t j = 0, x = 0; j < 256; j++) {             /* :136-166 */
        if (!F[j]) continue;
        cp = put_sym_rle(cp, (int)j, &rle, present);
        cp = put_freq8(cp, F[j]);
        C[j] = x; x += (uint32_t)F[j];
    }
    *cp++ = 0;
    uint32_t tab = (uint32_t)(cp - out);

    uint32_t cap = ho_compress_bound_4x8(n);
    uint8_t *scratch = malloc(cap);
    if (!scratch) return -1;
    backw b = { scratch + cap };
    uint32_t R[4] = { L8, L8, L8, L8 };
    for (uint32_t i = n; i-- > 0;)                          /* :176-194: symbol i <-> state i & 3, last first */
        R[i & 3] = enc_step8(R[i & 3], &b, C[in[i]], (uint32_t)F[in[i]]);
    for (int z = 3; z >= 0; z--) bw_u32(&b, R[z]);
    uint32_t body = (uint32_t)(scratch + cap - b.p);
    memcpy(out + tab, b.p, body);
    *out_size = tab + body;
    put_hdr8(out, 0, *out_size, n);
    free(scratch);
    return 0;
}

/* rans_compress_O1, rANS_static.c:409-631 */
static int enc4x8_o1(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size) {
    if (n < 4) return enc4x8_o0(in, n, out, out_size);      /* :438-439 */
    int (*F)[256] = calloc(256, sizeof(*F));
    uint32_t (*Cm)[256] = calloc(256, sizeof(*Cm));
    uint8_t *scratch = NULL;
    int T[256] = {0}, tp[256], present[256];
    int rc = -1;
    if (!F || !Cm) goto done;
    {
        uint8_t prev = 0;
        for (uint32_t i = 0; i < n; i++) { F[prev][in[i]]++; T[prev]++; prev = in[i]; }   /* hist1_4 */
    }
    uint32_t q4 = n >> 2;
    for (int k = 1; k < 4; k++) F[0][in[(uint32_t)k * q4]]++;  /* :455-458 */
    T[0] += 3;
    for (int i = 0; i < 256; i++) tp[i] = T[i] != 0;

    uint8_t *cp = out + 9;
    int rle_i = 0;
    for (int i = 0; i < 256; i++) {
        if (!T[i]) continue;
        double p = ((double)4096) / T[i];                    /* :469 */
        for (;;) {                                           /* :470-492 */
            int t2 = 0, m = 0, M = 0;
            for (int j = 0; j < 256; j++) {
                if (!F[i][j]) continue;
                if (m < F[i][j]) { m = F[i][j]; M = j; }
                if ((F[i][j] = (int)(F[i][j] * p)) == 0) F[i][j] = 1;
                t2 += F[i][j];
            }
            t2++;
            if (t2 < 4096) { F[i][M] += 4096 - t2; break; }
            if (t2 - 4096 >= F[i][M] / 2) { p = .98; continue; }
            F[i][M] -= t2 - 4096;
            break;
        }
        cp = put_sym_rle(cp, i, &rle_i, tp);                 /* :495-508 */
        for (int j = 0; j < 256; j++) present[j] = F[i][j] != 0;
        int rle_j = 0;
        for (uint32_t j = 0, x = 0; j < 256; j++) {          /* :510-540 */
            if (!F[i][j]) continue;
            cp = put_sym_rle(cp, (int)j, &rle_j, present);
            cp = put_freq8(cp, F[i][j]);
            Cm[i][j] = x; x += (uint32_t)F[i][j];
        }
        *cp++ = 0;
    }
    *cp++ = 0;
    uint32_t tab = (uint32_t)(cp - out);

    uint32_t cap = ho_compress_bound_4x8(n);
    scratch = malloc(cap);
    if (!scratch) goto done;
    backw b = { scratch + cap };
    uint32_t R[4] = { L8, L8, L8, L8 };
    /* state k owns quarter k; state 3 also the remainder; coded last to first with the
     * preceding byte as context and context 0 at the start of each quarter (:557-600) */
    for (uint32_t i = n - 1; i >= 4 * q4; i--)
        R[3] = enc_step8(R[3], &b, Cm[in[i - 1]][in[i]], (uint32_t)F[in[i - 1]][in[i]]);
    for (uint32_t t = q4; t-- > 1;)
        for (int k = 3; k >= 0; k--) {
            uint32_t pos = (uint32_t)k * q4 + t;
            R[k] = enc_step8(R[k], &b, Cm[in[pos - 1]][in[pos]], (uint32_t)F[in[pos - 1]][in[pos]]);
        }
    for (int k = 3; k >= 0; k--) {
        uint8_t s = in[(uint32_t)k * q4];
        R[k] = enc_step8(R[k], &b, Cm[0][s], (uint32_t)F[0][s]);
    }
    for (int k = 3; k >= 0; k--) bw_u32(&b, R[k]);
    uint32_t body = (uint32_t)(scratch + cap - b.p);
    memcpy(out + tab, b.p, body);
    *out_size = tab + body;
    put_hdr8(out, 1, *out_size, n);
    rc = 0;
done:
    free(scratch); free(F); free(Cm);
    return rc;
}

/* rans_compress, rANS_static.c:927-932.  out must hold ho_compress_bound_4x8(n) bytes. */
int ho_compress_4x8(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int order) {
    return order ? enc4x8_o1(in, n, out, out_size) : enc4x8_o0(in, n, out, out_size);
}
