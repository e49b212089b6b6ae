/*
 * ref_mt.c -- multi-threaded timing driver around a block codec's C entry points.
 * TEST / BASELINE INFRASTRUCTURE ONLY (built into oracle/_ref/libref.so next to the unmodified
 * reference sources, and into libhtsoracle.so next to our restatement).
 *
 * This is how the reference is meant to be parallelised: "the caller runs one block per thread"
 * (reference rANS_static4x16pr.c:853-858 keeps only per-thread TLS).  Threads pull block indices
 * from a shared atomic counter (dynamic schedule), each with its own pre-touched output buffer.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef unsigned char *(*dec_fn)(unsigned char *, unsigned int, unsigned char *, unsigned int *);
typedef unsigned char *(*enc_fn)(unsigned char *, unsigned int, unsigned char *, unsigned int *, int);

#ifdef REF_MT_REFERENCE
unsigned char *rans_uncompress_to_4x16(unsigned char *, unsigned int, unsigned char *, unsigned int *);
unsigned char *rans_compress_to_4x16(unsigned char *, unsigned int, unsigned char *, unsigned int *, int);
unsigned int rans_compress_bound_4x16(unsigned int, int);
unsigned char *rans_uncompress(unsigned char *, unsigned int, unsigned int *);
#define DEC rans_uncompress_to_4x16
#define ENC rans_compress_to_4x16
#define BOUND rans_compress_bound_4x16
#else
#include "hts_oracle.h"
static unsigned char *dec_wrap(unsigned char *in, unsigned int n, unsigned char *out, unsigned int *osz) {
    return ho_uncompress(in, n, out, osz) == 0 ? out : NULL;
}
static unsigned char *enc_wrap(unsigned char *in, unsigned int n, unsigned char *out, unsigned int *osz, int order) {
    return ho_compress(in, n, out, osz, order) == 0 ? out : NULL;
}
#define DEC dec_wrap
#define ENC enc_wrap
#define BOUND ho_compress_bound
#endif

typedef struct {
    const uint8_t *base;
    const uint64_t *off;
    const uint32_t *ilen;
    const uint32_t *olen;   /* decode: uncompressed size; encode: ignored */
    const int32_t *order;   /* encode only */
    int nblk, reps, encode, method;
    volatile long next;
    volatile long errors;
    uint64_t out_bytes;     /* sum of produced bytes (checksum-ish, defeats dead-code elimination) */
    pthread_mutex_t mu;
} job_t;

static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    uint32_t cap = 0;
    for (int i = 0; i < J->nblk; i++) {
        uint32_t c = J->encode ? BOUND(J->ilen[i], J->order[i]) : J->olen[i];
        if (c > cap) cap = c;
    }
    unsigned char *buf = malloc((size_t)cap + 64);
    if (!buf) { __sync_fetch_and_add(&J->errors, 1); return NULL; }
    memset(buf, 1, (size_t)cap + 64);                       /* pre-touch */
    uint64_t produced = 0;
    long total = (long)J->nblk * J->reps;
    for (;;) {
        long t = __sync_fetch_and_add(&J->next, 1);
        if (t >= total) break;
        int i = (int)(t % J->nblk);
        unsigned char *in = (unsigned char *)(J->base + J->off[i]);
        unsigned int osz;
        unsigned char *r;
        if (J->encode) {
            osz = cap;
            r = ENC(in, J->ilen[i], buf, &osz, J->order[i]);
#ifdef REF_MT_REFERENCE
        } else if (J->method == 1) {
            r = rans_uncompress(in, J->ilen[i], &osz);
            free(r);
#endif
        } else {
            osz = J->olen[i];
            r = DEC(in, J->ilen[i], buf, &osz);
        }
        if (!r) __sync_fetch_and_add(&J->errors, 1);
        produced += osz;
    }
    pthread_mutex_lock(&J->mu);
    J->out_bytes += produced;
    pthread_mutex_unlock(&J->mu);
    free(buf);
    return NULL;
}

/* Runs reps passes over the nblk blocks on nthreads threads; returns wall seconds (<0 on error). */
double ref_mt_run(const uint8_t *base, const uint64_t *off, const uint32_t *ilen, const uint32_t *olen,
                  const int32_t *order, int nblk, int nthreads, int reps, int encode, int method,
                  uint64_t *out_bytes) {
    job_t J;
    memset(&J, 0, sizeof(J));
    J.base = base; J.off = off; J.ilen = ilen; J.olen = olen; J.order = order;
    J.nblk = nblk; J.reps = reps; J.encode = encode; J.method = method;
    pthread_mutex_init(&J.mu, NULL);
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < nthreads; k++) pthread_create(&th[k], NULL, worker, &J);
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    if (out_bytes) *out_bytes = J.out_bytes;
    if (J.errors) return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
