/*
 * htscodecs_b200.h -- C ABI of libhtscodecs_b200.so: the B200 (sm_100a) implementation of
 * htscodecs' static-rANS hot path.  Plain C, pointers and sizes only.
 *
 * Section 1 re-exports the reference's own entry points unchanged (a drop-in for
 * htscodecs/rANS_static4x16.h:40-50 and rANS_static.h:40-43 of the reference): same names, same
 * argument meaning, same NULL-on-error behaviour, malloc'ed results the caller free()s.
 * Section 2 is new: batched entry points (many independent CRAM blocks per call), which is
 * what a GPU needs to be fast.  Every byte of codec work happens in CUDA kernels; there is no
 * CPU fallback -- the calls fail (NULL / negative status) if no sm_100 device is usable.
 */
#ifndef HTSCODECS_B200_H
#define HTSCODECS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * 1. Drop-in API (replaces the reference symbols of the same name)
 * ---------------------------------------------------------------------------------------- */

/* order argument / first stream byte (reference rANS_static4x16pr.c:39-43).  Bits 8-15 of
 * `order` carry the stripe count N (0 means 4).  RANS_ORDER_X32 (0x04) selects the 32-way
 * interleave; it is not in the v1.1 reference and is defined as its N=32 generalisation. */
#define RANS_ORDER_1      0x01
#define RANS_ORDER_X32    0x04
#define RANS_ORDER_STRIPE 0x08
#define RANS_ORDER_NOSZ   0x10
#define RANS_ORDER_CAT    0x20
#define RANS_ORDER_RLE    0x40
#define RANS_ORDER_PACK   0x80

/* replaces rans_compress_bound_4x16, reference rANS_static4x16.h:40 (.c:360-372) */
unsigned int rans_compress_bound_4x16(unsigned int size, int order);

/* replaces rans_compress_to_4x16, reference rANS_static4x16.h:41-43 (.c:1138-1345).
 * out == NULL: a bound-sized buffer is malloc'ed.  out != NULL: *out_size is the capacity on
 * entry (must be >= rans_compress_bound_4x16(in_size, order)) and the stream length on return. */
unsigned char *rans_compress_to_4x16(unsigned char *in, unsigned int in_size,
                                     unsigned char *out, unsigned int *out_size, int order);

/* replaces rans_compress_4x16, reference rANS_static4x16.h:44-45 (.c:1347-1350) */
unsigned char *rans_compress_4x16(unsigned char *in, unsigned int in_size,
                                  unsigned int *out_size, int order);

/* replaces rans_uncompress_to_4x16, reference rANS_static4x16.h:46-47 (.c:1352-1636).
 * out == NULL: the result is malloc'ed.  out != NULL: *out_size is the capacity on entry
 * (exactly the stored size for X_STRIPE streams, the expected size for X_NOSZ streams). */
unsigned char *rans_uncompress_to_4x16(unsigned char *in, unsigned int in_size,
                                       unsigned char *out, unsigned int *out_size);

/* replaces rans_uncompress_4x16, reference rANS_static4x16.h:48-49 (.c:1638-1641) */
unsigned char *rans_uncompress_4x16(unsigned char *in, unsigned int in_size,
                                    unsigned int *out_size);

/* replaces rans_uncompress (legacy CRAM 3.0 rANS 4x8), reference rANS_static.h:42-43
 * (rANS_static.c:934-943). */
unsigned char *rans_uncompress(unsigned char *in, unsigned int in_size, unsigned int *out_size);

/* replaces rans_compress (legacy rANS 4x8 encoder), reference rANS_static.h:40-41
 * (rANS_static.c:927-932; order 0 :85-218, order 1 :409-631).  The result is malloc'ed;
 * in_size == 0 returns NULL (the reference divides by in_size there). */
unsigned char *rans_compress(unsigned char *in, unsigned int in_size, unsigned int *out_size, int order);

/* ------------------------------------------------------------------------------------------
 * 2. Batched API (new; no reference equivalent -- the reference's callers loop over blocks,
 *    e.g. tests/rANS_static4x16pr_test.c:190-215, one call per block per thread)
 * ---------------------------------------------------------------------------------------- */

typedef struct hts_b200_ctx hts_b200_ctx;

/* OR this into order[i] of the batched encoders to code block i with the legacy rANS 4x8 codec
 * (order 0 / 1 in the low bits); its capacity bound is hts_b200_compress_bound_4x8(). */
#define HTS_B200_ORDER_RANS4x8 0x40000000
/* the reference encoder's own buffer size for n input bytes, rANS_static.c:87 */
unsigned int hts_b200_compress_bound_4x8(unsigned int size);

/* per-block status codes */
#define HTS_B200_OK            0
#define HTS_B200_ERR_FORMAT   (-1)   /* malformed stream (the reference returns NULL) */
#define HTS_B200_ERR_SIZE     (-2)   /* output capacity too small / size mismatch */
#define HTS_B200_ERR_SCRATCH  (-3)   /* device scratch (arena / work lists) too small: only reported by the
                                        asynchronous device-resident calls (sync == 0), which cannot retry;
                                        the next call on the context starts with the grown scratch */
#define HTS_B200_ERR_NESTED   (-4)   /* X_STRIPE inside X_STRIPE: never written by the encoder */
#define HTS_B200_ERR_INTERNAL (-5)

/* block codec selector for the batched decoders */
#define HTS_B200_RANS4x16 0          /* CRAM 3.1 rANS Nx16 (block method RANSPR = 5) */
#define HTS_B200_RANS4x8  1          /* CRAM 3.0 rANS 4x8  (block method RANS   = 4) */

/* One context per (host thread, device).  Owns a CUDA stream, work lists and scratch arenas that
 * grow on demand and are reused across calls.  Not thread-safe: use one per thread (the drop-in
 * calls above keep a thread-local one).  device < 0 means the current CUDA device. */
hts_b200_ctx *hts_b200_create(int device);
void hts_b200_destroy(hts_b200_ctx *ctx);
const char *hts_b200_last_error(const hts_b200_ctx *ctx);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
unsigned long long hts_b200_launch_count(const hts_b200_ctx *ctx);
/* device memory the context currently holds for work lists, scratch arenas and staging (diagnostics) */
size_t hts_b200_scratch_bytes(const hts_b200_ctx *ctx);
/* the context's CUDA stream (a cudaStream_t), so callers can order their own work or events.
 * It is a NON-BLOCKING stream: it does not wait for the legacy default stream, so device inputs
 * (data, offsets, lengths, capacities) produced on another stream must be complete -- or ordered
 * before this stream with an event -- when a *_dev call is made. */
void *hts_b200_stream(const hts_b200_ctx *ctx);

/*
 * Device-resident batched decode.  Every pointer argument is a DEVICE pointer on the context's
 * device.  Block i reads in_base[in_off[i] .. +in_len[i]) and writes out_base[out_off[i] ..).
 * out_len[i] is the capacity on entry (and the expected size for X_NOSZ streams) and the decoded
 * size on return; status[i] receives HTS_B200_OK or an error.  method[i] picks the codec
 * (NULL = all RANS4x16).  Work is enqueued on the context's stream; the call returns after the
 * stream has drained (sync != 0) or immediately (sync == 0: results are valid once the stream
 * has been synchronised; a rare scratch-arena overflow is then reported per block as
 * HTS_B200_ERR_INTERNAL instead of being retried).
 * Returns 0, or a negative value if the batch could not be run at all.
 */
int hts_b200_uncompress_batch_dev(hts_b200_ctx *ctx, int nblk,
                                  const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                                  uint8_t *out_base, const uint64_t *out_off, uint32_t *out_len,
                                  int32_t *status, const uint8_t *method, int sync);

/*
 * Host-resident batched decode: same contract with HOST pointers (pinned or pageable).  Inputs
 * are copied to the device, decoded and copied back in overlapping chunks on the context's
 * streams.  This is the end-to-end path bench.py times ("e2e").
 */
int hts_b200_uncompress_batch_host(hts_b200_ctx *ctx, int nblk,
                                   const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                                   uint8_t *out_base, const uint64_t *out_off, uint32_t *out_len,
                                   int32_t *status, const uint8_t *method);

/*
 * Device-resident batched encode.  order[i] is the reference's `order` argument for block i (OR in
 * HTS_B200_ORDER_RANS4x8 for the legacy codec).  out_len[i]: capacity on entry -- a block whose capacity is below
 * rans_compress_bound_4x16(in_len[i], order[i]) (hts_b200_compress_bound_4x8 for the legacy codec) is refused
 * with HTS_B200_ERR_SIZE and nothing is written to it, like the reference's coders (rANS_static4x16pr.c:396-397,
 * :706-707) -- stream length on return.  A plain block (no X_PACK / X_RLE / X_STRIPE) is coded inside its own
 * output region, so in and out must not overlap.  sync != 0: the call returns when the results are there (and has
 * retried by itself if order-1 alphabets beyond 16 symbols needed more scratch than the context held); sync == 0:
 * such blocks report HTS_B200_ERR_SCRATCH and the next call starts with enough.
 */
int hts_b200_compress_batch_dev(hts_b200_ctx *ctx, int nblk,
                                const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                                uint8_t *out_base, const uint64_t *out_off, uint32_t *out_len,
                                int32_t *status, const int32_t *order, int sync);

/* The same, always asynchronous: host_in_len / host_order are HOST copies of in_len / order (the encoder lays out
 * its work lists on the host from them), so nothing is read back and the call returns as soon as the kernels are
 * enqueued on hts_b200_stream(ctx).  hts_b200_compress_batch_dev(..., sync = 0) without them fetches the two arrays
 * first (one small copy and a stream synchronisation before the kernels are enqueued).  A block that needs more
 * scratch than the context holds reports HTS_B200_ERR_SCRATCH; the next call starts with enough. */
int hts_b200_compress_batch_dev_async(hts_b200_ctx *ctx, int nblk,
                                      const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                                      uint8_t *out_base, const uint64_t *out_off, uint32_t *out_len,
                                      int32_t *status, const int32_t *order,
                                      const uint32_t *host_in_len, const int32_t *host_order);

int hts_b200_compress_batch_host(hts_b200_ctx *ctx, int nblk,
                                 const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                                 uint8_t *out_base, const uint64_t *out_off, uint32_t *out_len,
                                 int32_t *status, const int32_t *order);

/* Pointer-array convenience forms (what a caller holding one malloc'ed buffer per CRAM block
 * uses); thin wrappers that stage through pinned memory and call the *_host functions. */
int rans4x16_uncompress_batch(hts_b200_ctx *ctx, int nblk,
                              const unsigned char *const *in, const unsigned int *in_size,
                              unsigned char *const *out, unsigned int *out_size, int *status);
int rans4x16_compress_batch(hts_b200_ctx *ctx, int nblk,
                            const unsigned char *const *in, const unsigned int *in_size,
                            unsigned char *const *out, unsigned int *out_size,
                            const int *order, int *status);

/*
 * Try-all-methods encode: the selection loop of the reference's name-tokeniser back end
 * (tokenise_name3.c:1246-1299 `compress()`, whose level table :1254-1260 lists up to nine orders per
 * byte column).  Block i is coded with every order in methods[0 .. nmethods) in ONE batched device
 * pass -- orders carrying X_STRIPE are skipped for sizes that are not a multiple of 4 (:1269) -- and
 * the first strictly smallest stream wins (:1280): it is written to out[i], its order to best[i]
 * (may be NULL).  out_size[i]: capacity on entry (>= the largest rans_compress_bound_4x16(in_size[i], m)
 * over the methods), stream length on return.  Only the winners leave the device.
 */
int rans4x16_compress_best_batch(hts_b200_ctx *ctx, int nblk,
                                 const unsigned char *const *in, const unsigned int *in_size,
                                 unsigned char *const *out, unsigned int *out_size,
                                 const int *methods, int nmethods, int *best, int *status);

/* Stored uncompressed size of a 4x16 (method 0) or 4x8 (method 1) stream held in HOST memory;
 * returns 0 and sets *ulen, or -1 (X_NOSZ stream / truncated header). */
int hts_b200_peek_size(const uint8_t *in, uint32_t in_len, int method, uint32_t *ulen);

/* Host-buffer calls overlap their host->device and device->host copies (full duplex, the default).
 * full = 0 sends all inputs of a call first and fetches the results afterwards (staging memory =
 * the whole batch) for hosts that handle mixed-direction traffic badly.  The environment variable
 * HTSCODECS_B200_COPY_DUPLEX=half|full sets the initial value. */
void hts_b200_set_copy_duplex(hts_b200_ctx *ctx, int full);

/* Diagnostics: the chunk boundaries hts_b200_{un,}compress_batch_host would use for this batch (same arguments,
 * HOST pointers; no device work).  Writes up to max_cuts boundaries -- blocks [cuts[k], cuts[k+1]) form chunk k --
 * and returns how many there are (chunks + 1), or -1. */
int hts_b200_plan_chunks(int enc, int nblk, const uint8_t *in_base, const uint64_t *in_off,
                         const uint32_t *in_len, const uint32_t *out_len, const uint8_t *method,
                         const int32_t *order, int *cuts, int max_cuts);

/* ------------------------------------------------------------------------------------------
 * 3. Multi-device host-buffer calls (blocks are independent, so a batch is cut into
 *    contiguous ranges balanced on uncompressed bytes, one host thread + context per device, and
 *    the per-block sizes / status land in the caller's arrays -- no collective).  The reference's
 *    own loop over blocks, tests/rANS_static4x16pr_test.c:190-215, becomes one call.
 * ---------------------------------------------------------------------------------------- */

/* Same contract as hts_b200_{un,}compress_batch_host, spread over devices[0 .. ndev).  The library keeps
 * one context per device (created on first use, reused, destroyed at exit); calls are serialised.
 * Default policy: every device runs its own full-duplex chunk pipeline (host->device, kernels, device->host
 * overlapped) and takes chunks of the WHOLE batch from a shared cursor whenever one of its pipeline stages is free, so
 * devices behind a slower host link end up with fewer blocks (on the 8 x B200 host of round 2, GPUs 0-3 moved
 * 10.7 GB/s each and GPUs 4-7 17 GB/s each).  hts_b200_multi_set_phased(1) / HTSCODECS_B200_MULTI_PHASED=1 selects the
 * "phased" policy instead -- contiguous ranges balanced on uncompressed bytes (hts_b200_partition), every device sends
 * its inputs first, the threads meet at a host-side barrier, then the results travel back -- for hosts whose
 * device->host rate collapses while host->device copies are in flight. */
int hts_b200_uncompress_batch_host_multi(int ndev, const int *devices, int nblk,
                                         const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                                         uint8_t *out_base, const uint64_t *out_off, uint32_t *out_len,
                                         int32_t *status, const uint8_t *method);
int hts_b200_compress_batch_host_multi(int ndev, const int *devices, int nblk,
                                       const uint8_t *in_base, const uint64_t *in_off, const uint32_t *in_len,
                                       uint8_t *out_base, const uint64_t *out_off, uint32_t *out_len,
                                       int32_t *status, const int32_t *order);
void hts_b200_multi_set_phased(int phased);   /* 1 phased, 0 full duplex + shared chunk queue, -1 environment / default (0) */

/* Per-device breakdown of the last multi-device call (busy times from CUDA events on the copy and
 * compute streams, wall times from the host clock). */
typedef struct hts_b200_dev_stats {
    int device, first_blk, nblk, pad;  /* first_blk: phased policy only (-1 when chunks were taken from the shared queue) */
    uint64_t in_bytes, out_bytes;      /* bytes sent to / fetched from the device (payload) */
    double h2d_ms, kernel_ms, d2h_ms;  /* summed busy time of the chunks' copies and kernels */
    double h2d_phase_ms, wait_ms, d2h_phase_ms;   /* phased mode: send phase, barrier wait, fetch phase (host clock) */
    double wall_ms;                    /* the device thread's whole call */
} hts_b200_dev_stats;
int hts_b200_multi_last_stats(hts_b200_dev_stats *out, int max);   /* returns the number of devices */
unsigned long long hts_b200_multi_launch_count(void);              /* kernels launched by the multi-device contexts */
const char *hts_b200_multi_last_error(void);

/* The partition the multi-device calls use: cuts[0 .. nparts] with blocks [cuts[k], cuts[k+1]) going to part k,
 * contiguous, balanced on the cumulative weights (uncompressed bytes).  Pure host arithmetic. */
int hts_b200_partition(int nblk, const uint32_t *weight, int nparts, int *cuts);

/* ------------------------------------------------------------------------------------------
 * 4. Container glue: the `[u32 clen][stream]...` framing written by the reference's test programs
 *    (tests/rANS_static4x16pr_test.c:261-296, rANS_static_test.c; native-endian 32-bit length, then
 *    the stream) and consumed by htslib-style callers that hold many CRAM blocks in one buffer.
 * ---------------------------------------------------------------------------------------- */

/* Walks a framed buffer and fills the batch arrays of the host-buffer decode calls: in_off[i] / in_len[i]
 * locate stream i inside `buf`, out_len[i] is its stored uncompressed size (hts_b200_peek_size), out_off[i]
 * the running sum of the sizes rounded up to `out_align` (0 or 1: packed).  method: HTS_B200_RANS4x16 / 4x8.
 * Returns the number of frames (which may exceed max_blk: only the first max_blk are written; pass
 * max_blk = 0 to count), or -1 for a truncated frame / a stream without a stored size.  *out_total (may be
 * NULL) receives the output bytes needed. */
long hts_b200_frames_scan(const uint8_t *buf, size_t len, int method, long max_blk,
                          uint64_t *in_off, uint32_t *in_len, uint64_t *out_off, uint32_t *out_len,
                          uint64_t *out_total, uint32_t out_align);
/* Writer twin: lays nblk encoded streams (src_base + src_off[i], src_len[i]; status[i] != 0 frames are
 * skipped when status is not NULL) out as `[u32 clen][stream]...` into dst.  Returns the bytes written, or
 * (size_t)-1 when dst_cap is too small.  dst == NULL: returns the size needed. */
size_t hts_b200_frames_write(uint8_t *dst, size_t dst_cap, long nblk, const uint8_t *src_base,
                             const uint64_t *src_off, const uint32_t *src_len, const int32_t *status);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost) for callers without a CUDA runtime. */
void *hts_b200_host_alloc(size_t bytes);
void hts_b200_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* HTSCODECS_B200_H */
