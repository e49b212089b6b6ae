/*
 * hts_oracle.h -- CPU restatement of the htscodecs static-rANS hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (htscodecs_b200/, include/) links,
 * imports or executes this.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may use it, and only as the checker.
 *
 * Parity status: the N=4 paths are PINNED -- tests/test_oracle.py checks them byte-for-byte
 * against the reference's 24 r4x16 + 8 r4x8 golden streams (tests/golden/) and, in the build
 * container, against oracle/_ref/libref.so (the unmodified reference C) on randomised inputs.
 * The X_32 (flag 0x04, 32 interleaved states) paths are "PARITY UNPINNED": the mounted reference
 * (v1.1) has no 32-way codec, so they are the N=32 generalisation of the pinned N=4 code
 *  and are validated only by N=4 == reference plus round trips.
 */
#ifndef HTS_ORACLE_H
#define HTS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* flag byte, reference rANS_static4x16pr.c:39-43 (0x04 = X_32, absent from the v1.1 reference) */
#define HO_ORDER1 0x01
#define HO_X32    0x04
#define HO_STRIPE 0x08
#define HO_NOSZ   0x10
#define HO_CAT    0x20
#define HO_RLE    0x40
#define HO_PACK   0x80

/* rans_compress_bound_4x16, reference rANS_static4x16pr.c:360-372 */
unsigned int ho_compress_bound(unsigned int size, int order);

/* rans_compress_to_4x16, reference rANS_static4x16pr.c:1138-1345.  out must hold
 * ho_compress_bound(n, order) bytes; *out_size is set to the bytes written.  0 ok, -1 error. */
int ho_compress(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int order);

/* rans_uncompress_to_4x16, reference rANS_static4x16pr.c:1352-1636.  *out_size: in = capacity
 * (and the expected size for X_NOSZ streams), out = bytes produced.  0 ok, -1 error. */
int ho_uncompress(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size);

/* Peek the uncompressed size stored in a 4x16 stream (0 ok, -1 if X_NOSZ / malformed). */
int ho_peek_size(const uint8_t *in, uint32_t in_size, uint32_t *ulen);

/* rans_uncompress (legacy 4x8), reference rANS_static.c:934-943.  *out_size: in = capacity. */
int ho_uncompress_4x8(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size);

/* rans_compress (legacy 4x8 encoder), reference rANS_static.c:927-932 (O0 :85-218, O1 :409-631).
 * out must hold ho_compress_bound_4x8(n) bytes.  0 ok, -1 error (n == 0 is undefined in the reference). */
unsigned int ho_compress_bound_4x8(unsigned int n);
int ho_compress_4x8(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int order);

/* Bare entropy coders (no container), exposed so tests can pin them separately.
 * nway is 4 or 32.  out capacity must be >= ho_compress_bound(n, order)-20. */
int ho_enc_o0(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int nway);
int ho_enc_o1(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t *out_size, int nway);
int ho_dec_o0(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t n, int nway);
int ho_dec_o1(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t n, int nway);

/* transforms (pack.c, rle.c), exposed for unit tests */
int ho_pack(const uint8_t *in, uint32_t n, uint8_t *meta, int *meta_len, uint8_t *out, uint32_t *out_len);
int ho_rle_encode(const uint8_t *in, uint32_t n, uint8_t *runs, uint32_t *runs_len,
                  uint8_t *syms, int *nsyms, uint8_t *lits, uint32_t *lits_len);

/* 7-bit big-endian varints, reference varint.h:85-104 / :131-160 */
int ho_var_put_u32(uint8_t *p, uint32_t v);
int ho_var_get_u32(const uint8_t *p, const uint8_t *end, uint32_t *v);

#ifdef __cplusplus
}
#endif
#endif
