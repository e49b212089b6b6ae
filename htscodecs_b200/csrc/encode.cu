// encode.cu -- sm_100a kernels for batched rANS Nx16 encode (order-0/1, 4-way and X_32) with the
// PACK / RLE / STRIPE / CAT / NOSZ container logic of rans_compress_to_4x16
// (reference rANS_static4x16pr.c:1138-1345).
//
// The host expands each block into *leaves* (one non-striped container each: a whole block, or
// one candidate method for one stripe sub-array, :1190-1211) and *streams* (one entropy-coded
// byte string each: a leaf's body, or its RLE run-length meta data).  All "try it and keep the
// smallest" decisions of the reference are taken on the device from the candidate sizes.
//
//   enc_stripe_kernel     byte transpose of X_STRIPE blocks                       (:1161-1180)
//   enc_transform_kernel  PACK (pack.c:56-151) and RLE (rle.c:48-138) per leaf, RLE keep rule (:1287)
//   enc_hist_kernel       order-0 histogram; order-1 pair counts over the compacted alphabet
//                         (utils.h:81-202, + the segment-start counts of :720-723)
//   enc_table_kernel      normalise_freq (:116-163), compute_shift (:629-691, doubles), table
//                         serialisation (:182-325), encoder symbol tables (rANS_word.h:190-266),
//                         optional order-0 compression of the order-1 table (:767-780)
//   enc_rans_kernel       the reverse rANS loop, lane = state; renormalisation words are placed by
//                         __ballot_sync / __popc suffix offsets                  (:442-485, :805-839)
//   enc_finish_kernel     container assembly, CAT fallback (:1332-1337), RLE-meta raw/rANS choice (:1298-1308)
//   enc_block_kernel      stripe candidate selection (:1192-1208) and block output
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "encode.h"

namespace hb {

// ------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------
struct EncSym {            // RansEncSymbol, rANS_word.h:170-181 (16 bytes)
    uint32_t x_max, rcp_freq, bias, cmpl_shift;   // cmpl_freq << 16 | (rcp_shift - 32)
};

struct EncStream {
    const uint8_t* src;    // bytes to code (dynamic for transformed leaves)
    uint32_t n;
    uint32_t order;        // requested order (0/1)
    uint32_t nway;         // 4 or 32
    uint32_t leaf;
    uint8_t* out;          // [table ... free ... payload written backwards from out + cap]
    uint32_t cap;          // even
    uint32_t order_eff;    // order actually used (0 when n is too small, :1322)
    uint32_t tab_len;      // table bytes at out[0..tab_len)
    uint32_t pay_off;      // payload = out[pay_off .. cap)
    uint32_t size;         // tab_len + payload length; 0xffffffff = failed
    uint32_t ns;           // alphabet size (order-1: including the forced symbol 0)
    uint32_t shift;        // order-1 table bits (10 / 12)
    uint32_t* F0;          // 256 x u32: order-0 histogram / presence
    uint32_t* F1;          // order-1: ns x ns pair counts, then normalised freqs (compact by rank)
    EncSym* syms;          // order-0: 256 entries; order-1: ns x ns entries (compact)
    uint8_t* ctab;         // scratch for the order-0 compressed order-1 table
    uint32_t codec;        // 0: rANS Nx16 (CRAM 3.1), 1: legacy rANS 4x8 (CRAM 3.0; byte-wise renormalisation)
    uint32_t hdr;          // != 0: the stream is built IN the block's own output region, `hdr` bytes in (a plain leaf:
                           // [flags][varint n] or the 9-byte 4x8 header precede the table); out / cap are patched by
                           // enc_fix_kernel and enc_finish_kernel only moves the payload down behind the table
};

struct EncLeaf {
    const uint8_t* src;    // leaf input
    uint8_t* out;          // assembled container
    uint8_t* packed;       // PACK scratch (n + 16)
    uint8_t* lits;         // RLE literal scratch (n + 16)
    uint8_t* rmeta;        // RLE meta scratch: header grows down from rmeta + 272, runs up from there
    uint32_t n;
    uint32_t flags;        // requested flag byte
    uint32_t blk;
    uint32_t body;         // stream index
    uint32_t meta;         // stream index of the RLE meta (or 0xffffffff)
    uint32_t out_flags;    // flag byte as emitted
    uint32_t pmeta_len;    // PACK meta bytes ([nsym][symbols])
    uint32_t packed_len;
    uint32_t rmeta_len;    // RLE meta length (0 = RLE not used)
    uint32_t rmeta_off;    // meta starts at rmeta + rmeta_off
    uint32_t lit_len;
    uint32_t cur_n;        // bytes handed to the entropy coder
    uint32_t out_size;     // container size (device result)
    int32_t status;
    uint8_t pmeta[20];
};

struct EncBlock {
    uint8_t* out;          // block output
    const uint8_t* in;
    uint8_t* tr;           // transposed copy (stripe)
    uint32_t n;
    uint32_t order;        // as passed by the caller
    uint32_t cap;
    uint32_t mode;         // 0 plain leaf, 1 stripe, 2 X_CAT requested, 3 error, 4 legacy rANS 4x8 leaf
    uint32_t leaf0;        // first leaf
    uint32_t N, ncand;     // stripe: leaves are [j * ncand + c]
    uint32_t pad;
};

struct EncWork {
    EncLeaf* leaves; EncStream* streams; EncBlock* blocks;
    uint32_t nleaves, nstreams, nblocks;
    uint32_t next_misc[16];        // persistent-kernel cursors, one per enc_rans_kernel launch
    // Streams bucketed by coder variant once their tables are known (enc_bucket_kernel): vlist[v * nstreams + k] is
    // the k-th stream of variant v (= its cursor index), so every warp of a variant's launch holds a full group.
    uint32_t vcount[16];
    uint32_t* vlist;
    uint32_t o0_lo, o1_lo;         // large batches: 4-way alphabets up to these sizes go to the compact-table variants
    // Order-1 scratch that depends on the alphabet found by the histogram pass: pair counts and encoder symbols of
    // alphabets beyond 16 symbols (ns x ns x 20 bytes) and the staging area of a table worth compressing come from this
    // arena (bump allocation; `overflow` makes the host retry with a larger one).  Small alphabets never touch it.
    uint8_t* arena;
    unsigned long long arena_cap, arena_used;
    uint32_t overflow;
};

__device__ __forceinline__ uint8_t* enc_arena_alloc(EncWork* W, uint64_t bytes) {
    bytes = (bytes + 255) & ~255ull;
    unsigned long long at = atomicAdd(&W->arena_used, (unsigned long long)bytes);
    if (at + bytes > W->arena_cap) { W->overflow = 1; return nullptr; }
    return W->arena + at;
}

__constant__ double c_log10[257];   // log(1024 + k)  (host libm, see encode_init)
__constant__ double c_log12[257];   // log(4096 + k)

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pow2_ceil(uint32_t v) {      // round2, :105-114
    if (v == 0) return 0;
    uint32_t p = 1;
    while (p < v && p) p <<= 1;
    return p;
}

// normalise_freq, rANS_static4x16pr.c:116-163, on `cnt` entries (absent symbols carry 0 and are
// skipped exactly like the reference's `if (!F[j]) continue`).  Serial, one thread.
__device__ int scale_freqs(uint32_t* F, uint32_t cnt, uint32_t total, uint32_t target) {
    if (!total) return 0;
    int sum_in = (int)total;
    uint32_t big_at = 0;
    for (int attempt = 0;; attempt++) {
        unsigned long long mul = ((unsigned long long)target << 31) / (unsigned long long)(long long)sum_in +
                                 (unsigned long long)((1 << 30) / sum_in);
        uint32_t big = 0;
        int sum = 0;
        big_at = 0;
        for (uint32_t j = 0; j < cnt; j++) {
            uint32_t f = F[j];
            if (!f) continue;
            if (big < f) { big = f; big_at = j; }
            uint32_t s = (uint32_t)(((unsigned long long)f * mul) >> 31);
            s = s ? s : 1;
            F[j] = s;
            sum += (int)s;
        }
        int slack = (int)(target - (uint32_t)sum);
        if (slack > 0) { F[big_at] += (uint32_t)slack; break; }
        if (slack == 0) break;
        uint32_t need = (uint32_t)(-slack);
        if (F[big_at] > need && (attempt == 1 || F[big_at] / 2 >= need)) { F[big_at] -= need; break; }
        if (attempt < 1) { sum_in = sum; continue; }
        slack += (int)F[big_at] - 1;
        F[big_at] = 1;
        for (uint32_t j = 0; slack && j < cnt; j++) {
            if (F[j] < 2) continue;
            int take = (F[j] > (uint32_t)(-slack)) ? slack : 1 - (int)F[j];
            F[j] = (uint32_t)((int)F[j] + take);
            slack -= take;
        }
        break;
    }
    return F[big_at] > 0 ? 0 : -1;
}

// RansEncSymbolInit, rANS_word.h:190-266
// lbits: log2 of the coder's lower bound (15 for Nx16 with 16-bit words; 23 for 4x8 with bytes,
// rANS_byte.h:195-266), wbits: renormalisation unit in bits
__device__ __forceinline__ EncSym make_sym(uint32_t start, uint32_t freq, uint32_t bits, uint32_t lbits = 15,
                                           uint32_t wbits = 16) {
    EncSym s;
    s.x_max = (((1u << lbits) >> bits) << wbits) * freq;
    uint32_t cmpl = ((1u << bits) - freq) & 0xffffu;
    if (freq < 2) {
        s.rcp_freq = ~0u;
        s.bias = start + (1u << bits) - 1;
        s.cmpl_shift = cmpl << 16;                      // shift 0
    } else {
        uint32_t sh = 0;
        while (freq > (1u << sh)) sh++;
        s.rcp_freq = (uint32_t)(((1ull << (sh + 31)) + freq - 1) / freq);
        s.bias = start;
        s.cmpl_shift = (cmpl << 16) | (sh - 1);
    }
    return s;
}

// encode_alphabet, :182-206.  `present(j)` tells whether symbol j is in the alphabet.
template <typename P>
__device__ uint32_t put_alphabet(uint8_t* p, P present) {
    uint8_t* p0 = p;
    int implied = 0;
    for (int j = 0; j < 256; j++) {
        if (!present(j)) continue;
        if (implied) { implied--; continue; }
        *p++ = (uint8_t)j;
        if (j && present(j - 1)) {
            int e = j + 1;
            while (e < 256 && present(e)) e++;
            implied = e - (j + 1);
            *p++ = (uint8_t)implied;
        }
    }
    *p++ = 0;
    return (uint32_t)(p - p0);
}

// ------------------------------------------------------------------------------------------
// enc_stripe_kernel: transposed[at[j] + x] = in[x*N + j]   (:1168-1180)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) enc_stripe_kernel(EncWork* W) {
    __shared__ uint32_t at[256];
    for (uint32_t b = blockIdx.x; b < W->nblocks; b += gridDim.x) {
        const EncBlock B = W->blocks[b];
        if (B.mode != 1) continue;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t a = 0;
            for (uint32_t j = 0; j < B.N; j++) { at[j] = a; a += B.n / B.N + ((B.n % B.N) > j); }
        }
        __syncthreads();
        const bool fast4 = B.N == 4 && (B.n & 15) == 0 && (reinterpret_cast<uintptr_t>(B.in) & 15) == 0 &&
                           (reinterpret_cast<uintptr_t>(B.tr) & 3) == 0;
        if (fast4) {
            // 16 input bytes = 4 elements x 4 streams: one 32-bit word per stream, 128-byte lines per warp
            const uint4* src = reinterpret_cast<const uint4*>(B.in);
            const uint32_t q = B.n / 4;                      // bytes per stream (multiple of 4)
            for (uint32_t t = threadIdx.x; t < B.n / 16; t += blockDim.x) {
                const uint4 v = __ldg(src + t);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t sel = (uint32_t)j | ((4u + j) << 4);          // byte j of each of two words
                    const uint32_t lo = __byte_perm(v.x, v.y, sel), hi = __byte_perm(v.z, v.w, sel);
                    *reinterpret_cast<uint32_t*>(B.tr + (size_t)j * q + 4 * (size_t)t) = __byte_perm(lo, hi, 0x5410);
                }
            }
        } else {
            for (uint32_t i = threadIdx.x; i < B.n; i += blockDim.x) B.tr[at[i % B.N] + i / B.N] = B.in[i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// enc_transform_kernel: PACK then RLE for one leaf per CTA
// ------------------------------------------------------------------------------------------
constexpr int TT = 256;

template <typename T>
__device__ __forceinline__ T cta_exscan(T v, T* warp_tot, T* total) {
    uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { T y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += y; }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    T pre = 0, tot = 0;
    for (int k = 0; k < TT / 32; k++) { T t = warp_tot[k]; if (k < (int)w) pre += t; tot += t; }
    __syncthreads();
    *total = tot;
    return pre + x - v;
}
// inclusive max-scan of int32 over the CTA
__device__ __forceinline__ int cta_incl_maxscan(int v, int* warp_tot) {
    uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x = max(x, y); }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    int pre = INT_MIN;
    for (int k = 0; k < (int)w; k++) pre = max(pre, warp_tot[k]);
    __syncthreads();
    return max(pre, x);
}

__global__ void __launch_bounds__(TT) enc_transform_kernel(EncWork* W) {
    __shared__ uint32_t seen[256];
    __shared__ int code[256];
    __shared__ int score[256];
    __shared__ uint32_t wtot[TT / 32];
    __shared__ int wmax[TT / 32];
    __shared__ int s_sp[TT];
    __shared__ uint32_t s_n, s_cnt;
    const uint32_t tid = threadIdx.x;
    for (uint32_t li = blockIdx.x; li < W->nleaves; li += gridDim.x) {
        EncLeaf& L = W->leaves[li];
        if (!(L.flags & (F_PACK | F_RLE))) continue;
        __syncthreads();
        const uint8_t* cur = L.src;
        uint32_t cur_n = L.n;
        uint32_t out_flags = L.flags;

        // ---------------- PACK, pack.c:56-151 + :1244-1267
        if (L.flags & F_PACK) {
            if (cur_n == 0) out_flags &= ~F_PACK;                    // :1265-1267
            else {
                seen[tid] = 0;
                __syncthreads();
                {                                                    // which byte values occur: 16 bytes per load
                    const uint32_t head = min(cur_n, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(cur) & 15)) & 15));
                    const uint32_t nv = (cur_n - head) / 16;
                    const uint4* v = reinterpret_cast<const uint4*>(cur + head);
                    for (uint32_t i = tid; i < head; i += TT) seen[cur[i]] = 1;
                    for (uint32_t i = tid; i < nv; i += TT) {
                        const uint4 q = __ldg(v + i);
                        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
#pragma unroll
                            for (int bb = 0; bb < 4; bb++) seen[(w[k] >> (8 * bb)) & 0xffu] = 1;
                        }
                    }
                    for (uint32_t i = head + nv * 16 + tid; i < cur_n; i += TT) seen[cur[i]] = 1;
                }
                __syncthreads();
                if (tid == 0) {
                    uint32_t ns = 0;
                    for (int s = 0; s < 256; s++) if (seen[s]) { code[s] = (int)ns; if (ns < 16) L.pmeta[1 + ns] = (uint8_t)s; ns++; }
                    L.pmeta[0] = (uint8_t)ns;                        // 256 wraps to 0
                    s_cnt = ns;
                }
                __syncthreads();
                const uint32_t ns = s_cnt;
                if (ns > 16 && ns != 256) {
                    out_flags &= ~F_PACK;                            // :1249-1253
                } else if (ns == 256) {                              // the wrap quirk: PACK kept, data copied, 1-byte meta
                    for (uint32_t i = tid; i < cur_n; i += TT) L.packed[i] = cur[i];
                    if (tid == 0) { L.pmeta_len = 1; L.packed_len = cur_n; }
                    cur = L.packed;
                } else {
                    const uint32_t per = ns > 4 ? 2 : ns > 2 ? 4 : ns > 1 ? 8 : 0;
                    uint32_t plen = 0;
                    if (per) {
                        const uint32_t bits = 8 / per;
                        plen = (cur_n + per - 1) / per;
                        // four output bytes (one aligned word) per thread from 4 * per = 8 / 16 / 32 input bytes
                        uint32_t done = 0;
                        if ((reinterpret_cast<uintptr_t>(cur) & 15) == 0 && (reinterpret_cast<uintptr_t>(L.packed) & 3) == 0) {
                            const uint32_t full = cur_n / (4 * per);
                            const uint2* src8 = reinterpret_cast<const uint2*>(cur);
                            uint32_t* dst = reinterpret_cast<uint32_t*>(L.packed);
                            for (uint32_t w = tid; w < full; w += TT) {
                                uint32_t word = 0;
                                const uint32_t n8 = per / 2;         // 8-byte pieces per output word
                                for (uint32_t h = 0; h < n8; h++) {
                                    const uint2 q = __ldg(src8 + (size_t)w * n8 + h);
                                    const uint32_t in2[2] = {q.x, q.y};
                                    uint32_t v = 0;                  // 8 codes -> 8 * bits bits
#pragma unroll
                                    for (int k = 0; k < 8; k++) v |= (uint32_t)code[(in2[k >> 2] >> (8 * (k & 3))) & 0xffu] << (k * bits);
                                    word |= v << (8 * bits * h);
                                }
                                dst[w] = word;
                            }
                            done = full * 4;
                        }
                        for (uint32_t o = done + tid; o < plen; o += TT) {
                            uint32_t v = 0;
                            for (uint32_t k = 0; k < per && o * per + k < cur_n; k++) v |= (uint32_t)code[cur[o * per + k]] << (k * bits);
                            L.packed[o] = (uint8_t)v;
                        }
                    }
                    if (tid == 0) { L.pmeta_len = ns + 1; L.packed_len = plen; }
                    cur = L.packed; cur_n = plen;
                }
                __syncthreads();
            }
        }

        // ---------------- RLE, rle.c:48-138 + :1269-1319
        uint32_t rmeta_len = 0;
        if (L.flags & F_RLE) {
            if (cur_n == 0) out_flags &= ~F_RLE;                     // :1317-1319
            else {
                score[tid] = 0;
                __syncthreads();
                // rle_find_syms: +1 when a byte repeats its predecessor, -1 otherwise
                // (each thread scores 16 consecutive bytes and adds one total per stretch of equal
                // bytes: run-heavy data would otherwise serialise on same-address atomics)
                for (uint32_t t0 = 0; t0 < cur_n; t0 += TT * 16) {
                    const uint32_t base = t0 + tid * 16;
                    if (base >= cur_n) continue;
                    const uint32_t cnt = min(16u, cur_n - base);
                    int prev = base ? (int)cur[base - 1] : -1, sym = -1, acc = 0;
                    for (uint32_t k = 0; k < cnt; k++) {
                        const int b = cur[base + k];
                        if (b != sym) { if (sym >= 0) atomicAdd(&score[sym], acc); sym = b; acc = 0; }
                        acc += (prev == b) ? 1 : -1;
                        prev = b;
                    }
                    atomicAdd(&score[sym], acc);
                }
                __syncthreads();
                if (tid == 0) {
                    uint32_t nrs = 0;
                    uint8_t* hdr = L.rmeta + 272;
                    for (int s = 0; s < 256; s++) if (score[s] > 0) nrs++;
                    uint8_t* m = hdr - 1 - nrs;
                    m[0] = (uint8_t)nrs;
                    uint32_t k = 0;
                    for (int s = 0; s < 256; s++) if (score[s] > 0) m[1 + k++] = (uint8_t)s;
                    s_cnt = nrs;
                }
                __syncthreads();
                const uint32_t nrs = s_cnt;
                uint8_t* runs = L.rmeta + 272;
                // A byte starts a literal unless it continues a run of an RLE symbol.  A run's varint
                // is emitted at the run's LAST byte (varints appear in literal order either way).
                // Each thread owns RE consecutive bytes of a TT*RE-byte tile: three register-resident
                // walks over them, separated by one CTA scan each (literal count; position of the most
                // recent literal start, a max-scan; varint bytes).
                constexpr int RE = 16;
                uint32_t lit_base = 0, run_base = 0;
                int start_carry = -1;                                // last literal start before this tile
                for (uint32_t t0 = 0; t0 < cur_n; t0 += TT * RE) {
                    const uint32_t base = t0 + tid * RE;
                    const uint32_t cnt = base < cur_n ? min((uint32_t)RE, cur_n - base) : 0u;
                    uint8_t bv[RE];
                    uint32_t isr_mask = 0;
#pragma unroll
                    for (int k = 0; k < RE; k++) {
                        bv[k] = (uint32_t)k < cnt ? cur[base + k] : 0;
                        if ((uint32_t)k < cnt && score[bv[k]] > 0) isr_mask |= 1u << k;
                    }
                    const int prevb = (cnt && base > 0) ? (int)cur[base - 1] : -1;
                    const int nextb = (cnt && base + cnt < cur_n) ? (int)cur[base + cnt] : -1;
                    // walk 1: which bytes start a literal
                    uint32_t start_mask = 0;
                    int last = -1;
                    {
                        int pb = prevb;
#pragma unroll
                        for (int k = 0; k < RE; k++) {
                            if ((uint32_t)k < cnt) {
                                const bool st = !(((isr_mask >> k) & 1u) && pb == (int)bv[k]);
                                if (st) { start_mask |= 1u << k; last = (int)(base + k); }
                                pb = bv[k];
                            }
                        }
                    }
                    uint32_t tot;
                    const uint32_t lidx = cta_exscan<uint32_t>(__popc(start_mask), wtot, &tot);
                    int sp = cta_incl_maxscan(last, wmax);           // inclusive over threads ...
                    s_sp[tid] = sp;
                    __syncthreads();
                    sp = tid ? max(s_sp[tid - 1], start_carry) : start_carry;   // ... made exclusive + tile carry
                    const int tile_last = max(s_sp[TT - 1], start_carry);
                    __syncthreads();
                    // walk 2: literals out, run lengths and their varint sizes
                    uint32_t rl[RE], vlen_tot = 0, end_mask = 0;
                    {
                        uint32_t li = lit_base + lidx;
                        int cs = sp;
#pragma unroll
                        for (int k = 0; k < RE; k++) {
                            rl[k] = 0;
                            if ((uint32_t)k < cnt) {
                                if ((start_mask >> k) & 1u) { L.lits[li++] = bv[k]; cs = (int)(base + k); }
                                const int nb = (k + 1 < (int)cnt) ? (int)bv[k + 1 < RE ? k + 1 : RE - 1] : nextb;
                                if (((isr_mask >> k) & 1u) && nb != (int)bv[k]) {
                                    end_mask |= 1u << k;
                                    rl[k] = base + k - (uint32_t)cs;
                                    vlen_tot += (uint32_t)var_len_u32(rl[k]);
                                }
                            }
                        }
                    }
                    uint32_t vtot;
                    uint32_t voff = cta_exscan<uint32_t>(vlen_tot, wtot, &vtot);
                    // walk 3: the varints
#pragma unroll
                    for (int k = 0; k < RE; k++)
                        if ((end_mask >> k) & 1u) voff += (uint32_t)var_put_u32(runs + run_base + voff, rl[k]);
                    lit_base += tot;
                    run_base += vtot;
                    start_carry = tile_last;
                }
                rmeta_len = 1 + nrs + run_base;
                const uint32_t lit_len = lit_base;
                if ((double)((unsigned long long)lit_len + rmeta_len) >= .99 * (double)cur_n) {   // :1287
                    out_flags &= ~F_RLE;
                    rmeta_len = 0;
                } else {
                    if (tid == 0) { L.rmeta_off = 272 - 1 - nrs; L.lit_len = lit_len; }
                    cur = L.lits; cur_n = lit_len;
                }
            }
        }
        if (tid == 0) {
            L.out_flags = out_flags;
            L.rmeta_len = rmeta_len;
            L.cur_n = cur_n;
            EncStream& S = W->streams[L.body];
            S.src = cur; S.n = cur_n;
            if (L.meta != 0xffffffffu) {
                EncStream& M = W->streams[L.meta];
                M.src = L.rmeta + (rmeta_len ? L.rmeta_off : 0u);
                M.n = rmeta_len;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// enc_hist_kernel: one CTA (4 warps) per stream
// ------------------------------------------------------------------------------------------
// Counting is done without atomics: every lane owns a private column of 16-bit counters -- word
// `lane` of a 128-byte row holds the lane's counters for slots 2r (low half) and 2r+1 (high
// half), so lane l only ever touches bank l: no conflicts, no lost updates -- 16 KB per warp for
// 256 slots, folded into 32-bit totals after at most HT_TILE bytes per lane.  Slots are
// byte values for the order-0 histogram (utils.h:81-102) and rank pairs ctx * ns + sym for
// order-1 statistics over alphabets of up to 16 symbols (utils.h:137-202); larger alphabets use
// shared-memory atomics on a 96 x 96 table, or global atomics beyond that.
constexpr int HT = 128;
constexpr int HT_WARPS = HT / 32;
constexpr uint32_t HT_TILE = 32768;                     // bytes per lane between folds (< 65536)
constexpr uint32_t O1_PRIV_NS = 16;                     // lane-private pair counters up to this alphabet size
constexpr uint32_t O1_SMEM_NS = 96;                     // shared-memory atomics up to this alphabet size
constexpr int HIST_SMEM = HT_WARPS * 256 * 64;          // 64 KB of private counters (reused as the 96 x 96 table)

// Address of lane `lane`'s counter for `slot` inside one warp's 16 KB area.
__device__ __forceinline__ uint16_t* hist_ctr(uint16_t* warp_area, uint32_t slot, uint32_t lane) {
    return warp_area + (slot >> 1) * 64 + lane * 2 + (slot & 1);
}

// Fold the private counters of `nslots` slots into tot[] (u32, shared) and clear them.  One
// thread per slot pair (one 128-byte row per warp area).
__device__ __forceinline__ void hist_fold(uint16_t* cnt, uint32_t* tot, uint32_t nslots) {
    __syncthreads();
    const uint32_t nrows = (nslots + 1) / 2;
    for (uint32_t r = threadIdx.x; r < nrows; r += HT) {
        uint32_t lo = 0, hi = 0;
        for (int w = 0; w < HT_WARPS; w++) {
            uint32_t* row = reinterpret_cast<uint32_t*>(cnt + (size_t)w * 256 * 32) + r * 32;
            for (uint32_t j = 0; j < 32; j++) {
                const uint32_t jj = (j + r) & 31;     // rotate: neighbouring rows start in different banks
                const uint32_t v = row[jj];
                lo += v & 0xffffu; hi += v >> 16;
                row[jj] = 0;
            }
        }
        tot[2 * r] += lo;
        if (2 * r + 1 < nslots) tot[2 * r + 1] += hi;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(HT) enc_hist_kernel(EncWork* W) {
    extern __shared__ __align__(16) uint8_t hsm[];
    __shared__ uint32_t h[256];
    __shared__ uint32_t ptot[O1_PRIV_NS * O1_PRIV_NS];
    __shared__ uint8_t rank[256];
    __shared__ uint32_t s_ns;
    uint16_t* cnt = reinterpret_cast<uint16_t*>(hsm);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint16_t* myarea = cnt + (size_t)warp * 256 * 32;                // this warp's 16 KB of counters

    for (uint32_t k = tid; k < HIST_SMEM / 16; k += HT) reinterpret_cast<uint4*>(hsm)[k] = make_uint4(0, 0, 0, 0);
    for (uint32_t si = blockIdx.x; si < W->nstreams; si += gridDim.x) {
        EncStream& S = W->streams[si];
        __syncthreads();
        const uint8_t* in = S.src;
        const uint32_t n = S.n, nway = S.nway;
        uint32_t order = S.order;
        if (S.codec == 0) { if (order && (n < 8 || n < nway)) order = 0; }   // :1322-1325 (+ N-way analogue)
        else if (order && n < 4) order = 0;                          // rANS_static.c:438-439
        for (uint32_t k = tid; k < 256; k += HT) h[k] = 0;
        __syncthreads();
        // ---- hist8, utils.h:81-102: aligned 16-byte chunks, one per thread per step
        const uintptr_t a = reinterpret_cast<uintptr_t>(in);
        const uint32_t head = min(n, (uint32_t)((16 - (a & 15)) & 15));
        const uint32_t nv = (n - head) / 16;
        const uint4* v = reinterpret_cast<const uint4*>(in + head);
        {
            for (uint32_t i = tid; i < head; i += HT) atomicAdd(&h[in[i]], 1u);
            for (uint32_t i = head + nv * 16 + tid; i < n; i += HT) atomicAdd(&h[in[i]], 1u);
            // PF chunks per thread are kept in flight: with 12 warps per SM a single 16-byte load each
            // would leave the kernel bound by memory latency (bytes in flight), not by its counters
            constexpr int PF = 4;
            uint32_t since = 0;
            uint4 nq[PF];
#pragma unroll
            for (int d = 0; d < PF; d++) nq[d] = (tid + d * HT < nv) ? __ldg(v + tid + d * HT) : make_uint4(0, 0, 0, 0);
            for (uint32_t i0 = 0; i0 < nv; i0 += PF * HT) {
#pragma unroll
                for (int d = 0; d < PF; d++) {
                    const uint32_t i = i0 + d * HT + tid;
                    const uint4 q = nq[d];
                    if (i + PF * HT < nv) nq[d] = __ldg(v + i + PF * HT);
                    if (i < nv) {
                        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
#pragma unroll
                            for (int bb = 0; bb < 4; bb++) (*hist_ctr(myarea, (w[k] >> (8 * bb)) & 0xffu, lane))++;
                        }
                    }
                }
                since += 16 * PF;
                if (since >= HT_TILE) { hist_fold(cnt, h, 256); since = 0; }          // block-uniform
            }
            hist_fold(cnt, h, 256);
        }
        for (uint32_t k = tid; k < 256; k += HT) S.F0[k] = h[k];
        if (tid == 0) { S.order_eff = order; S.size = 0; S.tab_len = 0; S.pay_off = S.cap; }
        if (!order || n == 0) continue;

        // ---- order 1: alphabet = present symbols + symbol 0 (:729-731); ranks; pair counts
        if (tid == 0) {
            uint32_t ns = 0;
            for (int s = 0; s < 256; s++) {
                bool p = h[s] != 0 || s == 0;
                rank[s] = (uint8_t)ns;
                if (p) ns++;
            }
            s_ns = ns;
            S.ns = ns;
        }
        __syncthreads();
        const uint32_t ns = s_ns;
        const uint32_t seg = n / nway;
        if (ns > 16) {                                               // beyond the stream's own 256-entry areas: from the arena
            __shared__ int s_noarena;
            if (tid == 0) {
                uint8_t* a = enc_arena_alloc(W, (uint64_t)ns * ns * 20);
                s_noarena = a == nullptr;
                if (a) { S.F1 = reinterpret_cast<uint32_t*>(a + (size_t)ns * ns * 16); S.syms = reinterpret_cast<EncSym*>(a); }
                else S.size = 0xffffffffu;
            }
            __syncthreads();
            if (s_noarena) continue;
        }
        if (ns <= O1_PRIV_NS) {
            // hist1_4, utils.h:137-202: every adjacent pair of the whole buffer, first context 0
            for (uint32_t k = tid; k < ns * ns; k += HT) ptot[k] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < head; i += HT) atomicAdd(&ptot[rank[i ? in[i - 1] : 0u] * ns + rank[in[i]]], 1u);
            for (uint32_t i = head + nv * 16 + tid; i < n; i += HT) atomicAdd(&ptot[rank[i ? in[i - 1] : 0u] * ns + rank[in[i]]], 1u);
            // the segment starts are coded in context 0 (:720-723, with 4 -> nway)
            for (uint32_t k = 1 + tid; k < nway; k += HT) atomicAdd(&ptot[rank[0] * ns + rank[in[k * seg]]], 1u);
            constexpr int PF = 4;                                    // chunks per thread in flight (see the order-0 pass)
            uint32_t since = 0;
            uint4 nq[PF];
            uint32_t nc[PF];                                         // byte preceding each prefetched chunk
#pragma unroll
            for (int d = 0; d < PF; d++) {
                const uint32_t i = tid + d * HT;
                nq[d] = make_uint4(0, 0, 0, 0); nc[d] = 0;
                if (i < nv) { nq[d] = __ldg(v + i); const uint32_t at = head + i * 16; nc[d] = at ? in[at - 1] : 0u; }
            }
            for (uint32_t i0 = 0; i0 < nv; i0 += PF * HT) {
#pragma unroll
                for (int d = 0; d < PF; d++) {
                    const uint32_t i = i0 + d * HT + tid;
                    const uint4 q = nq[d];
                    const uint32_t c0 = nc[d];
                    if (i + PF * HT < nv) { nq[d] = __ldg(v + i + PF * HT); nc[d] = in[head + (i + PF * HT) * 16 - 1]; }
                    if (i < nv) {
                        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                        // all 17 rank look-ups first (the counter stores below could alias the rank table as far
                        // as the compiler knows, which would serialise look-up and update per byte)
                        uint32_t rk[16];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
#pragma unroll
                            for (int bb = 0; bb < 4; bb++) rk[4 * k + bb] = rank[(w[k] >> (8 * bb)) & 0xffu];
                        }
                        uint32_t rc = rank[c0] * ns;                 // context of the chunk's first byte
#pragma unroll
                        for (int t = 0; t < 16; t++) {
                            (*hist_ctr(myarea, rc + rk[t], lane))++;
                            rc = rk[t] * ns;
                        }
                    }
                }
                since += 16 * PF;
                if (since >= HT_TILE) { hist_fold(cnt, ptot, ns * ns); since = 0; }
            }
            hist_fold(cnt, ptot, ns * ns);
            for (uint32_t k = tid; k < ns * ns; k += HT) S.F1[k] = ptot[k];
        } else {
            // the private-counter area doubles as the 96 x 96 atomic table (it is all zeros here)
            const bool in_smem = ns <= O1_SMEM_NS;
            uint32_t* P = in_smem ? reinterpret_cast<uint32_t*>(hsm) : S.F1;
            if (!in_smem) for (uint32_t k = tid; k < ns * ns; k += HT) P[k] = 0;
            __syncthreads();
            // (16 bytes per load, the rank look-ups before the atomics; byte-at-a-time this pass took 38 ms for 4096 x 1 MiB)
            auto pair = [&](uint32_t i) { atomicAdd(&P[rank[i ? in[i - 1] : 0u] * ns + rank[in[i]]], 1u); };
            for (uint32_t i = tid; i < head; i += HT) pair(i);
            for (uint32_t i = head + nv * 16 + tid; i < n; i += HT) pair(i);
            for (uint32_t i = tid; i < nv; i += HT) {
                const uint4 q = __ldg(v + i);
                const uint32_t at = head + i * 16;
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
                uint32_t rk[16];
#pragma unroll
                for (int k = 0; k < 4; k++) {
#pragma unroll
                    for (int bb = 0; bb < 4; bb++) rk[4 * k + bb] = rank[(w[k] >> (8 * bb)) & 0xffu];
                }
                uint32_t rc = rank[at ? in[at - 1] : 0u] * ns;
#pragma unroll
                for (int t = 0; t < 16; t++) {
                    atomicAdd(&P[rc + rk[t]], 1u);
                    rc = rk[t] * ns;
                }
            }
            for (uint32_t k = 1 + tid; k < nway; k += HT) atomicAdd(&P[rank[0] * ns + rank[in[k * seg]]], 1u);
            __syncthreads();
            if (in_smem) for (uint32_t k = tid; k < ns * ns; k += HT) { S.F1[k] = P[k]; P[k] = 0; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// lane groups and the encoder step (shared by enc_table_kernel's nested coder and enc_rans_kernel)
// ------------------------------------------------------------------------------------------
template <int NWAY> struct EGrp {
    static constexpr int G = 32 / NWAY;
    static constexpr uint32_t GM = (NWAY == 32) ? 0xffffffffu : ((1u << NWAY) - 1u);
    uint32_t g, glane, gshift, gmask;
    __device__ __forceinline__ EGrp() {
        uint32_t lane = lane_id();
        g = lane / NWAY; glane = lane % NWAY; gshift = g * NWAY; gmask = GM << gshift;
    }
};

// RansEncPutSymbol (rANS_word.h:281-321) for all lanes of the warp at once.  `wp` is the group's
// write pointer (moves down); emitting lanes store their low 16 bits in descending lane order.
// Branch-free: the 16-bit store is predicated and the state shift is a select, so the coder state never waits for a
// divergent branch to reconverge (the branchy form measured 160 cycles per 4-way step, two thirds of them the
// emit path's serialisation).
__device__ __forceinline__ void st_u16_if(uint8_t* p, uint32_t v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b16 h;\n\tsetp.ne.u32 q, %2, 0;\n\tcvt.u16.u32 h, %1;\n\t@q st.global.u16 [%0], h;\n\t}"
                 :: "l"(p), "r"(v), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void st_u8_if(uint8_t* p, uint32_t v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u8 [%0], %1;\n\t}"
                 :: "l"(p), "r"(v), "r"((uint32_t)pred) : "memory");
}

template <int NWAY>
__device__ __forceinline__ uint32_t enc_put(uint32_t x, bool act, const EncSym s, uint8_t*& wp, const EGrp<NWAY>& G) {
    const bool emit = act && x >= s.x_max;
    const uint32_t m = (__ballot_sync(0xffffffffu, emit) >> G.gshift) & EGrp<NWAY>::GM;
    const uint32_t above = __popc((m >> G.glane) >> 1);              // emitting lanes with a higher index write first
    // the stored half-word gets a register of its own: with the state register as the store's source, the state's
    // shift had to wait for the store to issue -- i.e. for the whole ballot / popc address chain (ncu: 11 of 129 cycles)
    st_u16_if(wp - 2 * (above + 1), __byte_perm(x, 0, 0x4410), emit);
    x = emit ? (x >> 16) : x;
    wp -= 2 * __popc(m);
    const uint32_t q = __umulhi(x, s.rcp_freq) >> (s.cmpl_shift & 31u);
    const uint32_t xn = x + s.bias + q * (s.cmpl_shift >> 16);
    return act ? xn : x;
}

// The X_32 Nx16 step in PTX (every lane active): 15 instructions, one predicate feeding the vote,
// the compacted 16-bit store and the state shift.  `wpo` = write offset from `base` (moves down).
__device__ __forceinline__ uint32_t enc_put_x32(uint32_t x, const EncSym s, uint32_t& wpo, const uint8_t* base, uint32_t gt) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 b, k, t, q;\n\t"
        ".reg .b16 h;\n\t"
        ".reg .b64 a;\n\t"
        "setp.ge.u32 p, %0, %2;\n\t"
        "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
        "and.b32 k, b, %7;\n\t"
        "popc.b32 k, k;\n\t"
        "mad.lo.u32 t, k, 0xfffffffe, %1;\n\t"
        "mad.wide.u32 a, t, 1, %6;\n\t"
        "cvt.u16.u32 h, %0;\n\t"
        "@p st.global.u16 [a+-2], h;\n\t"
        "@p shr.u32 %0, %0, 16;\n\t"
        "popc.b32 k, b;\n\t"
        "mad.lo.u32 %1, k, 0xfffffffe, %1;\n\t"
        "add.u32 t, %0, %4;\n\t"
        "mul.hi.u32 q, %0, %3;\n\t"
        "shf.r.wrap.b32 q, q, 0, %5;\n\t"
        "shr.u32 k, %5, 16;\n\t"
        "mad.lo.u32 %0, q, k, t;\n\t"
        "}"
        : "+r"(x), "+r"(wpo)
        : "r"(s.x_max), "r"(s.rcp_freq), "r"(s.bias), "r"(s.cmpl_shift), "l"(base), "r"(gt)
        : "memory");
    return x;
}

// The 4x8 form (rANS_byte.h:281-315): up to two renormalisation BYTES per state per step, low byte
// first, each written below the previous one; states emit in descending lane order.
template <int NWAY>
__device__ __forceinline__ uint32_t enc_put8(uint32_t x, bool act, const EncSym s, uint8_t*& wp, const EGrp<NWAY>& G) {
    const bool e1 = act && x >= s.x_max;
    const uint32_t x1 = e1 ? (x >> 8) : x;
    const bool e2 = e1 && x1 >= s.x_max;
    const uint32_t m1 = (__ballot_sync(0xffffffffu, e1) >> G.gshift) & EGrp<NWAY>::GM;
    const uint32_t m2 = (__ballot_sync(0xffffffffu, e2) >> G.gshift) & EGrp<NWAY>::GM;
    const uint32_t above = __popc((m1 >> G.glane) >> 1) + __popc((m2 >> G.glane) >> 1);
    uint8_t* p = wp - 1 - above;
    st_u8_if(p, __byte_perm(x, 0, 0x4440), e1);                        // (registers of their own, as in enc_put)
    st_u8_if(p - 1, __byte_perm(x, 0, 0x4441), e2);
    x = e2 ? (x1 >> 8) : x1;
    wp -= __popc(m1) + __popc(m2);
    const uint32_t q = __umulhi(x, s.rcp_freq) >> (s.cmpl_shift & 31u);
    const uint32_t xn = x + s.bias + q * (s.cmpl_shift >> 16);
    return act ? xn : x;
}
template <int NWAY, bool BYTE>
__device__ __forceinline__ uint32_t enc_step(uint32_t x, bool act, const EncSym s, uint8_t*& wp, const EGrp<NWAY>& G) {
    return BYTE ? enc_put8<NWAY>(x, act, s, wp, G) : enc_put<NWAY>(x, act, s, wp, G);
}

// ------------------------------------------------------------------------------------------
// enc_table_kernel: one CTA (128 threads) per stream
// ------------------------------------------------------------------------------------------
constexpr int KT = 128;

// rans_compress_O0_4x16's table half (:408-435) for `n` symbols with histogram F (256 entries,
// modified in place).  Writes the table to `tab`, the encoder symbols to `syms`; one thread.
__device__ int build_o0_tables_enc(uint32_t* F, uint32_t n, uint8_t* tab, EncSym* syms, uint32_t* tab_len) {
    uint32_t target = pow2_ceil(n);
    if (target > 4096) target = 4096;
    if (scale_freqs(F, 256, n, target) < 0) return -1;
    uint8_t* p = tab;
    p += put_alphabet(p, [&](int j) { return F[j] != 0; });
    for (int j = 0; j < 256; j++) if (F[j]) p += var_put_u32(p, F[j]);
    *tab_len = (uint32_t)(p - tab);
    if (scale_freqs(F, 256, target, 4096) < 0) return -1;
    uint32_t x = 0;
    for (int j = 0; j < 256; j++) {
        if (F[j]) { syms[j] = make_sym(x, F[j], 12); x += F[j]; }
    }
    return 0;
}

// ---- legacy rANS 4x8 tables ------------------------------------------------------------------
// One "sym [run]" list step of the 4x8 table format (rANS_static.c:139-153 / :495-508 / :518-529):
// `present(r)` tells whether rank r is listed, `sym(r)` its byte value.  Writes 0, 1 or 2 bytes.
template <typename P, typename Y>
__device__ __forceinline__ uint32_t put_sym_run(uint8_t* cp, uint32_t r, uint32_t nr, uint32_t& run, P present, Y sym) {
    if (run) { run--; return 0; }
    cp[0] = (uint8_t)sym(r);
    if (sym(r) && r && present(r - 1) && sym(r - 1) + 1 == sym(r)) {
        uint32_t e = r + 1;
        while (e < nr && present(e) && sym(e) == sym(e - 1) + 1) e++;
        run = e - (r + 1);
        cp[1] = (uint8_t)run;
        return 2;
    }
    return 1;
}
__device__ __forceinline__ uint32_t put_freq8(uint8_t* cp, uint32_t f) {          // :155-161
    if (f < 128) { cp[0] = (uint8_t)f; return 1; }
    cp[0] = (uint8_t)(128 | (f >> 8)); cp[1] = (uint8_t)(f & 0xff);
    return 2;
}

// rans_compress_O0's table half, rANS_static.c:100-166: one thread.  F: 256 counts (modified).
__device__ void table_4x8_o0(int* F, uint32_t n, uint8_t* tab, EncSym* syms, uint32_t* tab_len) {
    unsigned long long tr = (((unsigned long long)4096 << 31) / n) + (unsigned long long)((1 << 30) / (int)n);
    for (;;) {
        int fsum = 0, m = 0, M = 0;
        for (int j = 0; j < 256; j++) {
            if (!F[j]) continue;
            if (m < F[j]) { m = F[j]; M = j; }
            if ((F[j] = (int)(((unsigned long long)F[j] * tr) >> 31)) == 0) F[j] = 1;
            fsum += F[j];
        }
        fsum++;
        if (fsum < 4096) { F[M] += 4096 - fsum; break; }
        if (fsum - 4096 > F[M] / 2) { tr = 2104533975ull; continue; }
        F[M] -= fsum - 4096;
        break;
    }
    uint8_t* cp = tab;
    uint32_t run = 0, x = 0;
    for (uint32_t j = 0; j < 256; j++) {
        if (!F[j]) continue;
        cp += put_sym_run(cp, j, 256, run, [&](uint32_t r) { return F[r] != 0; }, [&](uint32_t r) { return r; });
        cp += put_freq8(cp, (uint32_t)F[j]);
        syms[j] = make_sym(x, (uint32_t)F[j], 12, 23, 8);
        x += (uint32_t)F[j];
    }
    *cp++ = 0;
    *tab_len = (uint32_t)(cp - tab);
}

// rANS 4x16 order-0 encode of the serialised order-1 table (:767-780) by one CTA: histogram by all
// threads, table by thread 0, the 4-way coder by lanes 0-3 of warp 0 (lane = state).  hist/syms
// are shared scratch (256 entries each).  Returns the stream length written to out (0 = failed);
// the value is CTA-uniform.
__device__ uint32_t nested_o0_encode(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, uint32_t* hist,
                                     EncSym* syms, uint32_t* s_res) {
    const uint32_t tid = threadIdx.x;
    for (uint32_t k = tid; k < 256; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += blockDim.x) atomicAdd(&hist[in[i]], 1u);
    __syncthreads();
    if (tid == 0) {
        uint32_t tab = 0;
        s_res[0] = build_o0_tables_enc(hist, n, out, syms, &tab) < 0 ? 0xffffffffu : tab;
    }
    __syncthreads();
    const uint32_t tab = s_res[0];
    __syncthreads();
    if (tab == 0xffffffffu) return 0;
    if (tid < 32) {
        const EGrp<4> G;
        const bool mine = tid < 4;
        uint8_t* const end = out + (cap & ~1u);
        uint8_t* wp = end;
        uint32_t x = 1u << 15;
        const uint32_t rows = (n + 3) / 4;
        for (uint32_t k = 0; k < rows; k++) {
            const uint32_t pos = (rows - 1 - k) * 4 + G.glane;
            const bool act = mine && pos < n;
            const EncSym sy = syms[act ? in[pos] : 0];
            x = enc_put<4>(x, act, sy, wp, G);
        }
        if (mine) {
            uint8_t* p = wp - 4 * (4 - tid);
            p[0] = (uint8_t)x; p[1] = (uint8_t)(x >> 8); p[2] = (uint8_t)(x >> 16); p[3] = (uint8_t)(x >> 24);
        }
        wp = reinterpret_cast<uint8_t*>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)wp, 0)) - 16;
        __syncwarp();
        const uint32_t body = (uint32_t)(end - wp);
        // move the payload down behind the table; ascending 32-byte steps never overtake unread bytes
        for (uint32_t i0 = 0; i0 < body; i0 += 32) {
            uint8_t v = 0;
            if (i0 + tid < body) v = wp[i0 + tid];
            __syncwarp();
            if (i0 + tid < body) out[tab + i0 + tid] = v;
            __syncwarp();
        }
        if (tid == 0) s_res[0] = tab + body;
    }
    __syncthreads();
    const uint32_t r = s_res[0];
    __syncthreads();
    return r;
}

__device__ __forceinline__ double approx_log(int x) {              // fast_log, :620-623
    double a = (double)x;
    long long bits = __double_as_longlong(a);
    return __dmul_rn(__ll2double_rn(bits - 4606921278410026770LL), 1.539095918623324e-16);
}

__global__ void __launch_bounds__(KT) enc_table_kernel(EncWork* W) {
    __shared__ uint32_t Fs[256];
    __shared__ uint8_t unrank[256];
    __shared__ int Sv[256];
    __shared__ uint32_t Tv[256];
    __shared__ uint32_t rowlen[256], rowoff[256];
    __shared__ uint32_t s_shift, s_fail, s_alpha_len, s_res[2];
    __shared__ int s_maxtot;
    __shared__ double rl10[256], rl12[256];
    __shared__ __align__(16) EncSym nsyms[256];
    const uint32_t tid = threadIdx.x;
    for (uint32_t si = blockIdx.x; si < W->nstreams; si += gridDim.x) {
        EncStream& S = W->streams[si];
        __syncthreads();
        const uint32_t n = S.n;
        const bool legacy = S.codec != 0;
        if (n == 0) { if (tid == 0) { S.size = legacy ? 0xffffffffu : 0u; S.tab_len = 0; } continue; }   // :405-406
        if (S.size == 0xffffffffu) continue;                         // no table space (see enc_hist_kernel)
        if (S.order_eff == 0) {
            if (tid == 0) {
                uint32_t tab = 0;
                for (int j = 0; j < 256; j++) Fs[j] = S.F0[j];
                if (legacy) table_4x8_o0(reinterpret_cast<int*>(Fs), n, S.out, S.syms, &tab);
                else if (build_o0_tables_enc(Fs, n, S.out, S.syms, &tab) < 0) S.size = 0xffffffffu;
                S.tab_len = tab;
                uint32_t present = 0;
                for (int j = 0; j < 256; j++) present += S.F0[j] != 0;
                S.ns = present;                              // picks the order-0 coder variant (compact symbol table or not)
            }
            continue;
        }
        // ================= order 1, :719-780 =================
        const uint32_t ns = S.ns;
        uint32_t* F1 = S.F1;
        if (tid == 0) {
            uint32_t r = 0;
            for (int s = 0; s < 256; s++) if (S.F0[s] != 0 || s == 0) unrank[r++] = (uint8_t)s;
            s_fail = 0;
        }
        __syncthreads();
        // row totals T[i] (utils.h T0[] + the nway-1 segment starts)
        for (uint32_t i = tid; i < ns; i += KT) {
            uint32_t t = 0;
            for (uint32_t j = 0; j < ns; j++) t += F1[i * ns + j];
            Tv[i] = t;
        }
        __syncthreads();
        if (legacy) {
            // ---- rans_compress_O1's table half, rANS_static.c:460-545, over ranks (ascending symbols)
            auto symof = [&](uint32_t r) { return (uint32_t)unrank[r]; };
            for (uint32_t i = tid; i < ns; i += KT) {        // normalise each context to 4096, size its table
                int* row = reinterpret_cast<int*>(F1 + i * ns);
                const int T = (int)Tv[i];
                uint32_t len = 0;
                if (T) {
                    double p = __ddiv_rn(4096.0, (double)T);                         // :469
                    for (;;) {
                        int t2 = 0, m = 0, M = 0;
                        for (uint32_t j = 0; j < ns; j++) {
                            if (!row[j]) continue;
                            if (m < row[j]) { m = row[j]; M = (int)j; }
                            if ((row[j] = __double2int_rz(__dmul_rn((double)row[j], p))) == 0) row[j] = 1;
                            t2 += row[j];
                        }
                        t2++;
                        if (t2 < 4096) { row[M] += 4096 - t2; break; }
                        if (t2 - 4096 >= row[M] / 2) { p = .98; continue; }
                        row[M] -= t2 - 4096;
                        break;
                    }
                    uint32_t run = 0;
                    uint8_t tmp[2];
                    for (uint32_t j = 0; j < ns; j++) {
                        if (!row[j]) continue;
                        len += put_sym_run(tmp, j, ns, run, [&](uint32_t r) { return row[r] != 0; }, symof);
                        len += row[j] < 128 ? 1u : 2u;
                    }
                    len++;                                                           // the row's terminating 0
                }
                rowlen[i] = len;
            }
            __syncthreads();
            uint8_t* tab = S.out;
            if (tid == 0) {                                  // context list (:495-508) and row offsets
                uint32_t o = 0, run = 0;
                for (uint32_t i = 0; i < ns; i++) {
                    if (!Tv[i]) continue;
                    o += put_sym_run(tab + o, i, ns, run, [&](uint32_t r) { return Tv[r] != 0; }, symof);
                    rowoff[i] = o;
                    o += rowlen[i];
                }
                tab[o++] = 0;
                S.tab_len = o;
                S.shift = 12;
            }
            __syncthreads();
            for (uint32_t i = tid; i < ns; i += KT) {        // write the rows, build the encoder symbols
                if (!Tv[i]) continue;
                const int* row = reinterpret_cast<const int*>(F1 + i * ns);
                uint8_t* cp = tab + rowoff[i];
                EncSym* srow = S.syms + (size_t)i * ns;
                uint32_t run = 0, x = 0;
                for (uint32_t j = 0; j < ns; j++) {
                    if (!row[j]) continue;
                    cp += put_sym_run(cp, j, ns, run, [&](uint32_t r) { return row[r] != 0; }, symof);
                    cp += put_freq8(cp, (uint32_t)row[j]);
                    srow[j] = make_sym(x, (uint32_t)row[j], 12, 23, 8);
                    x += (uint32_t)row[j];
                }
                *cp = 0;
            }
            continue;
        }
        // compute_shift, :629-691.  The per-pair terms are computed in parallel; the two running sums
        // are then accumulated by ONE thread in the reference's index order (double addition does
        // not commute with reordering), with no fused multiply-add anywhere.
        if (tid == 0) s_maxtot = 0;
        __syncthreads();
        for (uint32_t i = tid; i < ns; i += KT) {
            const uint32_t T = Tv[i];
            int max_val = (int)pow2_ceil(T);
            int nsym = 0, sm10 = 0, sm12 = 0;
            for (uint32_t j = 0; j < ns; j++) {
                const uint32_t f = F1[i * ns + j];
                if (!f) continue;
                nsym++;
                const uint32_t q = (uint32_t)max_val / f;
                if (q > 1024) sm10++;
                if (q > 4096) sm12++;
            }
            rl10[i] = c_log10[sm10]; rl12[i] = c_log12[sm12];
            if (nsym < 64 && max_val > 128) max_val /= 2;
            if (max_val > 1024) max_val /= 2;
            if (max_val > 4096) max_val = 4096;
            Sv[i] = max_val;
            atomicMax(&s_maxtot, max_val);
        }
        __syncthreads();
        double2* terms = reinterpret_cast<double2*>(S.syms);        // ns x ns x 16 B: not yet in use
        for (uint32_t k = tid; k < ns * ns; k += KT) {
            const uint32_t i = k / ns, f = F1[k];
            double2 t;
            t.x = t.y = __longlong_as_double(0x7ff8000000000000LL);  // NaN marks an absent pair
            if (f) {
                const double fd = (double)f, Td = (double)Tv[i];
                int x = __double2int_rz(__ddiv_rn(__dmul_rn(1024.0, fd), Td));
                t.x = __dmul_rn(fd, __dsub_rn(approx_log(x > 1 ? x : 1), rl10[i]));
                x = __double2int_rz(__ddiv_rn(__dmul_rn(4096.0, fd), Td));
                t.y = __dmul_rn(fd, __dsub_rn(approx_log(x > 1 ? x : 1), rl12[i]));
            }
            terms[k] = t;
        }
        __syncthreads();
        if (tid == 0) {
            double e10 = 0, e12 = 0;
            const uint32_t np = ns * ns;
#pragma unroll 8
            for (uint32_t k = 0; k < np; k++) {
                const double2 t = terms[k];
                if (t.x == t.x) {
                    e10 = __dadd_rn(__dsub_rn(e10, t.x), 4.0);
                    e12 = __dadd_rn(__dsub_rn(e12, t.y), 6.0);
                }
            }
            s_shift = (__ddiv_rn(e10, e12) < 1.01 || s_maxtot <= 1024) ? 10u : 12u;
        }
        __syncthreads();
        const uint32_t shift = s_shift;
        // normalise each row (:740-752), measure its serialised length (encode_freq_d :295-325)
        for (uint32_t i = tid; i < ns; i += KT) {
            uint32_t target = (uint32_t)Sv[i];
            if (shift == 10 && target > 1024) target = 1024;
            uint32_t* row = F1 + i * ns;
            if (scale_freqs(row, ns, Tv[i], target) < 0) s_fail = 1;
            Tv[i] = target;
            uint32_t len = 0, z = 0;
            for (uint32_t j = 0; j < ns; j++) {
                if (row[j]) { if (z) { len += 2; z = 0; } len += (uint32_t)var_len_u32(row[j]); }
                else z++;
            }
            if (z) len += 2;
            rowlen[i] = len;
        }
        __syncthreads();
        uint8_t* tab = S.out;
        if (tid == 0) {
            tab[0] = (uint8_t)(shift << 4);
            uint32_t al = put_alphabet(tab + 1, [&](int j) { return S.F0[j] != 0 || j == 0; });
            s_alpha_len = al;
            uint32_t o = 1 + al;
            for (uint32_t i = 0; i < ns; i++) { rowoff[i] = o; o += rowlen[i]; }
            S.tab_len = o;
            S.shift = shift;
        }
        __syncthreads();
        // write the rows, shift them up to 1<<shift (:756), build the encoder symbols (:758-762)
        for (uint32_t i = tid; i < ns; i += KT) {
            uint32_t* row = F1 + i * ns;
            uint8_t* p = tab + rowoff[i];
            uint32_t z = 0;
            for (uint32_t j = 0; j < ns; j++) {
                if (row[j]) {
                    if (z) { *p++ = 0; *p++ = (uint8_t)(z - 1); z = 0; }
                    p += var_put_u32(p, row[j]);
                } else z++;
            }
            if (z) { *p++ = 0; *p++ = (uint8_t)(z - 1); }
            uint32_t sh = 0, t = Tv[i];
            if (t != 0 && t != (1u << shift)) while ((t << sh) < (1u << shift)) sh++;
            uint32_t x = 0;
            EncSym* srow = S.syms + (size_t)i * ns;
            for (uint32_t j = 0; j < ns; j++) {
                uint32_t f = row[j] << sh;
                srow[j] = make_sym(x, f, shift);
                x += f;
            }
        }
        __syncthreads();
        if (tid == 0) {
            if (s_fail) S.size = 0xffffffffu;
            s_res[1] = (!s_fail && S.tab_len > 1000) ? S.tab_len : 0u;               // :767-780
        }
        __syncthreads();
        const uint32_t tlen = s_res[1];
        if (tlen) {                                                  // CTA-uniform
            const uint32_t usz = tlen - 1;
            const uint32_t ccap = ((usz + usz / 16 + 1024) & ~15u);
            if (tid == 0) {
                S.ctab = enc_arena_alloc(W, (uint64_t)ccap + 1024 + 4096 + 64);
                if (!S.ctab) S.size = 0xffffffffu;                   // the choice below decides output bytes: fail, never guess
                s_res[0] = S.ctab != nullptr;
            }
            __syncthreads();
            const bool have = s_res[0] != 0;
            __syncthreads();
            if (!have) continue;
            const uint32_t csz = nested_o0_encode(tab + 1, usz, S.ctab, ccap, Fs, nsyms, s_res);
            if (csz && csz + 6 < tlen) {
                uint32_t hdr = 0;
                if (tid == 0) {
                    uint8_t* op = tab;
                    *op++ |= 1;
                    op += var_put_u32(op, usz);
                    op += var_put_u32(op, csz);
                    hdr = (uint32_t)(op - tab);
                    S.tab_len = hdr + csz;
                    s_res[0] = hdr;
                }
                __syncthreads();
                hdr = s_res[0];
                for (uint32_t k = tid; k < csz; k += KT) tab[hdr + k] = S.ctab[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// enc_rans_kernel: persistent, one warp per CTA, a group of NWAY lanes per stream
// ------------------------------------------------------------------------------------------
constexpr int ENC_O0_SMEM_PER_GROUP = 256 * 16;                      // EncSym[256]
// Order 1, per group: [NSCAP x NSCAP 16-byte encoder symbols][256 B byte -> rank] and, for the small-alphabet variants,
// the same symbols once more as 8-byte packed entries {rcp_freq, bias | freq << 13 | rcp_shift << 26} for the hot loop:
// a 128-bit look-up per lane is four shared-memory wavefronts per warp (plus replays), and ncu shows the X_32 order-1
// coder bound by exactly those (LSU data pipe 93 % busy at 33 % issue); x_max and cmpl_freq are one shift and one
// subtraction away from freq (rANS_word.h:190-266).
__host__ __device__ constexpr bool enc_o1_packed(int nscap) { return nscap <= 16; }
__host__ __device__ constexpr int enc_o1_smem(int nscap) {
    return nscap * nscap * 16 + 256 + (enc_o1_packed(nscap) ? ((nscap * nscap * 8 + 15) & ~15) : 0);
}

// Per-lane backward byte reader of the order-1 loop.  A lane walks its own segment from the end,
// so a plain byte load per symbol would cost one memory transaction per lane per step; instead
// a lane pulls the aligned 16-byte line it is in once and shifts bytes out of registers.
struct ByteSrc {
    const uint8_t* line;         // aligned line currently held
    const uint8_t* lo_line;      // aligned line of the lane's first byte: nothing below it is touched
    uint32_t w0, w1, w2, w3;     // its bytes; the next byte to hand out is the top byte of w3
    uint4 pre;                   // the line below, already on its way
    uint32_t k;                  // bytes left in the window
    __device__ __forceinline__ uint4 fetch(const uint8_t* l) const {
        return (l >= lo_line) ? __ldg(reinterpret_cast<const uint4*>(l)) : make_uint4(0, 0, 0, 0);
    }
    // [begin, end): the lane's bytes, handed out from end - 1 down to begin
    __device__ __forceinline__ void init(const uint8_t* begin, const uint8_t* end) {
        lo_line = begin - (reinterpret_cast<uintptr_t>(begin) & 15);
        const uint32_t lo = (uint32_t)(reinterpret_cast<uintptr_t>(end) & 15);
        w0 = w1 = w2 = w3 = 0; k = 0;
        pre = make_uint4(0, 0, 0, 0);
        if (begin == end) { line = end; return; }
        if (lo == 0) { line = end; pre = fetch(line - 16); return; }
        line = end - lo;
        const uint4 v = fetch(line);                          // 16-byte aligned: never leaves the page of end[-1]
        w0 = v.x; w1 = v.y; w2 = v.z; w3 = v.w;
        pre = fetch(line - 16);
        for (uint32_t q = lo; q < 16; q++) {                  // drop the bytes at and above `end`
            w3 = __funnelshift_l(w2, w3, 8); w2 = __funnelshift_l(w1, w2, 8); w1 = __funnelshift_l(w0, w1, 8); w0 <<= 8;
        }
        k = lo;
    }
    // (A branch-free refill -- selects and one predicated load -- was measured SLOWER: order-1 encode 110 -> 95 GB/s
    // 4-way, 300 -> 263 X_32; the rare divergent refill costs less than eight extra instructions per byte.)
    __device__ __forceinline__ uint32_t get() {
        if (k == 0) {
            w0 = pre.x; w1 = pre.y; w2 = pre.z; w3 = pre.w;
            line -= 16;
            pre = fetch(line - 16);
            k = 16;
        }
        const uint32_t b = w3 >> 24;
        w3 = __funnelshift_l(w2, w3, 8); w2 = __funnelshift_l(w1, w2, 8); w1 = __funnelshift_l(w0, w1, 8); w0 <<= 8;
        k--;
        return b;
    }
};

// Which enc_rans_kernel variant codes a stream (the cursor indices of encode_run); 0xff: nothing to code.
__device__ __forceinline__ uint32_t enc_variant(const EncWork* W, const EncStream& S) {
    if (S.n == 0 || S.size == 0xffffffffu) return 0xffu;
    const bool x32 = S.nway == 32, legacy = S.codec != 0;
    if (S.order_eff == 0) {
        if (x32) return 2;
        if (S.ns <= W->o0_lo) return legacy ? 9 : 7;
        return legacy ? 5 : 0;
    }
    if (x32) return S.ns <= 16 ? 3 : 4;
    if (S.ns <= W->o1_lo) return legacy ? 10 : 8;
    if (S.ns <= 16) return legacy ? 6 : 1;
    return legacy ? 12 : 11;
}

// One CTA per variant: a stable partition of the stream list, so every variant codes its streams in batch order.
__global__ void __launch_bounds__(256) enc_bucket_kernel(EncWork* W) {
    const uint32_t v = blockIdx.x, w = threadIdx.x >> 5, l = threadIdx.x & 31, nstreams = W->nstreams;
    __shared__ uint32_t s_warp[8], s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (uint32_t s0 = 0; s0 < nstreams; s0 += 256) {
        const uint32_t si = s0 + threadIdx.x;
        const bool mine = si < nstreams && enc_variant(W, W->streams[si]) == v;
        const uint32_t bal = __ballot_sync(0xffffffffu, mine);
        if (l == 0) s_warp[w] = __popc(bal);
        __syncthreads();
        uint32_t off = s_base;
        for (uint32_t k = 0; k < w; k++) off += s_warp[k];
        if (mine) W->vlist[(size_t)v * nstreams + off + __popc(bal & ((1u << l) - 1u))] = si;
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t t = 0; for (int k = 0; k < 8; k++) t += s_warp[k]; s_base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) W->vcount[v] = s_base;
}

// Order-1 symbol tables live in shared memory when the alphabet has at most NSCAP symbols; two
// kernel variants (NSCAP 16: 4 KB per group, many resident warps; NSCAP 48: 36 KB) split the
// streams between them by alphabet size, larger alphabets read the table from global memory.
template <int NWAY, int ORDER, int NSCAP, bool BYTE = false>
__global__ void __launch_bounds__(32) enc_rans_kernel(EncWork* W, uint32_t cursor_id) {
    using EG = EGrp<NWAY>;
    extern __shared__ __align__(16) uint8_t esm[];
    const EG G;
    // order 0, NSCAP 48: symbol table compacted over the alphabet ([256 B byte -> rank][NSCAP x 16 B]), 1 KB per
    // stream instead of 4 KB -- the 4-way variant for batches too large for one wave of the 4 KB kernel
    constexpr bool O0C = ORDER == 0 && NSCAP == 48;
    constexpr uint32_t PER_GROUP = ORDER ? (uint32_t)enc_o1_smem(NSCAP) : (O0C ? (256 + NSCAP * 16) : ENC_O0_SMEM_PER_GROUP);
    constexpr bool PK = ORDER == 1 && enc_o1_packed(NSCAP);
    uint8_t* gsm = esm + G.g * PER_GROUP;
    EncSym* ssym = reinterpret_cast<EncSym*>(O0C ? gsm + 256 : gsm);
    uint8_t* srank = O0C ? gsm : gsm + NSCAP * NSCAP * 16;          // byte -> rank (order 1, compact order 0)
    uint32_t* cursor = &W->next_misc[cursor_id];
    const uint32_t nstreams = W->vcount[cursor_id];
    const uint32_t* vlist = W->vlist + (size_t)cursor_id * W->nstreams;
    uint32_t gt_mask;
    asm("mov.u32 %0, %%lanemask_gt;" : "=r"(gt_mask));

    // this variant's streams (enc_bucket_kernel) are claimed EG::G at a time from an atomic cursor (the launch is
    // shaped so that every SM holds the same number of CTAs)
    for (;;) {
        uint32_t s0 = 0;
        if (lane_id() == 0) s0 = atomicAdd(cursor, (uint32_t)EG::G);
        s0 = __shfl_sync(0xffffffffu, s0, 0);
        if (s0 >= nstreams) break;
        // (the groups of a last, partly filled warp code the list's last stream again -- identical bytes to identical
        //  addresses -- instead of idling: an idle group would keep the whole warp in the ragged loops at half speed)
        const bool act_s = true;
        EncStream* S = &W->streams[vlist[min(s0 + G.g, nstreams - 1u)]];

        const uint8_t* in = act_s ? S->src : nullptr;
        const uint32_t n = act_s ? S->n : 0;
        uint32_t ns = 1;
        const EncSym* syms = ssym;
        if (ORDER == 0) {
            if (act_s && !O0C) for (uint32_t k = G.glane; k < 256; k += NWAY) ssym[k] = S->syms[k];
            if (act_s && O0C) {                                      // ranks over the present symbols, their entries packed
                uint32_t r = 0;
                if (G.glane == 0) for (int sy = 0; sy < 256; sy++) if (S->F0[sy] != 0) { srank[sy] = (uint8_t)r; ssym[r] = S->syms[sy]; r++; }
            }
        } else if (act_s) {
            ns = S->ns;
            // symbol -> rank
            uint32_t r = 0;
            if (G.glane == 0) for (int s = 0; s < 256; s++) { srank[s] = (uint8_t)r; if (S->F0[s] != 0 || s == 0) r++; }
            if (ns <= NSCAP) {
                const uint32_t Mo1 = 1u << S->shift;
                for (uint32_t k = G.glane; k < ns * ns; k += NWAY) {
                    const EncSym e = S->syms[k];
                    ssym[k] = e;
                    if (PK) {
                        const uint32_t fq = Mo1 - (e.cmpl_shift >> 16);
                        reinterpret_cast<uint2*>(gsm + NSCAP * NSCAP * 16 + 256)[k] =
                            make_uint2(e.rcp_freq, (e.bias & 0x1fffu) | (fq << 13) | ((e.cmpl_shift & 31u) << 26));
                    }
                }
            } else syms = S->syms;
        }
        __syncwarp();

        uint8_t* wp = act_s ? S->out + S->cap : nullptr;             // payload grows down from the end
        uint32_t x = BYTE ? (1u << 23) : (1u << 15);                 // RansEncInit, rANS_word.h:69-72 / rANS_byte.h:68-71
        auto symidx = [&](uint32_t byte) -> uint32_t { return O0C ? (uint32_t)srank[byte] : byte; };
        if (ORDER == 0) {
            // symbol i belongs to state i % NWAY; rows are coded from the last to the first (:442-480).
            // The symbol bytes do not depend on the coder state, so they are fetched a batch of rows
            // ahead (non-coherent loads, free to move above the payload stores) while the serial
            // state updates of the previous batch run.
            const uint32_t rows = (n + NWAY - 1) / NWAY;
            const uint32_t maxrows = (NWAY == 32) ? rows : __reduce_max_sync(0xffffffffu, rows);
            // rows of this group that are complete and that every other group of the warp also has
            const uint32_t full = __reduce_min_sync(0xffffffffu, act_s ? n / NWAY : 0u);
            uint32_t k = 0;
            for (; k < maxrows - full; k++) {                        // ragged part (partial last row, shorter groups)
                const bool in_rows = k + rows >= maxrows;            // groups with fewer rows idle first
                const uint32_t kk = k - (maxrows - rows);
                const uint32_t pos = in_rows ? (rows - 1 - kk) * NWAY + G.glane : 0;
                const bool act = in_rows && pos < n;
                EncSym s = ssym[act ? symidx(__ldg(in + pos)) : 0];
                x = enc_step<NWAY, BYTE>(x, act, s, wp, G);
            }
            // from here on every lane of the warp codes rows full-1 .. 0 of its stream
            // 4-way: a lone warp per scheduler, so the fetch runs ~1300 cycles ahead.  (32 rows, and two rounds in flight
            // in the order-1 word source, were measured SLOWER -- 307 -> 294 and 213 -> 192 GB/s, presumably the twice
            // as long unrolled bodies missing the instruction cache (not profiled) -- although ncu still shows a quarter of
            // the samples on a batch's first byte.)
            constexpr int B = (NWAY == 32) ? 8 : 16;
            const uint8_t* ip = in + G.glane;
            uint8_t* const obase = act_s ? S->out : nullptr;         // (kept in registers: the asm steps clobber memory)
            uint32_t r = full;                                       // rows left
            uint32_t nb[B];
#pragma unroll
            for (int u = 0; u < B; u++) nb[u] = (r > (uint32_t)u) ? __ldg(ip + (size_t)(r - 1 - u) * NWAY) : 0u;
            const uint8_t* pp = ip + (size_t)r * NWAY;               // one row past the first row of the current batch
            while (r >= B) {
                uint32_t b[B];
#pragma unroll
                for (int u = 0; u < B; u++) b[u] = nb[u];
                r -= B;
                pp -= B * NWAY;                                      // rows r-1-u sit at pp - (u+1)*NWAY
                if (r >= B) {
#pragma unroll
                    for (int u = 0; u < B; u++) nb[u] = __ldg(pp - (u + 1) * NWAY);
                } else {
#pragma unroll
                    for (int u = 0; u < B; u++) nb[u] = (r > (uint32_t)u) ? __ldg(pp - (u + 1) * NWAY) : 0u;
                }
                EncSym sy[B];
#pragma unroll
                for (int u = 0; u < B; u++) sy[u] = ssym[symidx(b[u])];
                if (NWAY == 32 && !BYTE) {
                    uint32_t wpo = (uint32_t)(wp - obase);
#pragma unroll
                    for (int u = 0; u < B; u++) x = enc_put_x32(x, sy[u], wpo, obase, gt_mask);
                    wp = obase + wpo;
                } else {
#pragma unroll
                    for (int u = 0; u < B; u++) x = enc_step<NWAY, BYTE>(x, true, sy[u], wp, G);
                }
            }
            for (uint32_t u = 0; u < r; u++) {                       // fewer than B rows left (already fetched)
                uint32_t bsel = nb[0];
#pragma unroll
                for (int q = 1; q < B; q++) if (u == (uint32_t)q) bsel = nb[q];
                x = enc_step<NWAY, BYTE>(x, true, ssym[symidx(bsel)], wp, G);
            }
        } else {
            // state z owns in[z*seg, (z+1)*seg), the last state also the tail; coded last-to-first
            // with the preceding byte as context, 0 at a segment start (:794-834).  Lanes with fewer
            // symbols idle first, so every state finishes on the last step.
            const uint32_t seg = n / NWAY, tail = n - seg * NWAY;
            const uint32_t my_n = act_s ? seg + ((G.glane == NWAY - 1) ? tail : 0u) : 0u;
            const uint32_t maxsteps = __reduce_max_sync(0xffffffffu, my_n);
            const uint32_t minsteps = __reduce_min_sync(0xffffffffu, my_n);
            const uint32_t idle = maxsteps - my_n;
            ByteSrc src;
            src.init(in + (size_t)G.glane * seg, in + (size_t)G.glane * seg + my_n);
            const bool sm_syms = (syms == ssym);                     // this group's symbols are in shared memory
            uint8_t* const obase1 = act_s ? S->out : nullptr;
            uint32_t rs = 0;                                         // rank of the symbol to code next
            if (my_n) rs = srank[src.get()];
            uint32_t left = my_n;                                    // symbols this lane still has to code
            uint32_t k = 0;
            EncSym idle_sym;
            idle_sym.x_max = 0xffffffffu; idle_sym.rcp_freq = 0; idle_sym.bias = 0; idle_sym.cmpl_shift = 0;
            for (; k < maxsteps - minsteps; k++) {                   // ragged start: some lanes idle
                const bool act = k >= idle;
                EncSym s = idle_sym;
                if (act) {
                    const uint32_t rc = (left > 1) ? (uint32_t)srank[src.get()] : (uint32_t)srank[0];
                    s = syms[rc * ns + rs];
                    rs = rc; left--;
                }
                x = enc_step<NWAY, BYTE>(x, act, s, wp, G);
            }
            // every lane active, a context byte exists.  Table look-ups do not depend on the coder
            // state: the symbols of the next four steps are gathered (shared memory, or global memory
            // for large alphabets) while the four serial state updates of the current ones run.
            auto gather = [&](EncSym (&sy)[4]) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t rc = srank[src.get()];
                    const EncSym* sp = syms + rc * ns + rs;
                    if (sm_syms) sy[u] = *sp;
                    else {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(sp));
                        sy[u].x_max = v.x; sy[u].rcp_freq = v.y; sy[u].bias = v.z; sy[u].cmpl_shift = v.w;
                    }
                    rs = rc;
                }
            };
            auto run4 = [&](const EncSym (&sy)[4]) {
                if (NWAY == 32 && !BYTE) {
                    uint32_t wpo = (uint32_t)(wp - obase1);
#pragma unroll
                    for (int u = 0; u < 4; u++) x = enc_put_x32(x, sy[u], wpo, obase1, gt_mask);
                    wp = obase1 + wpo;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; u++) x = enc_step<NWAY, BYTE>(x, true, sy[u], wp, G);
                }
            };
            if (__all_sync(0xffffffffu, sm_syms)) {                  // shared-memory tables: nothing to hide
                // Rounds of 16 steps with the context bytes taken from aligned 32-bit words that every lane fetches
                // at the same step, a round ahead (ByteSrc's per-lane refill is a divergent branch per byte, and the
                // compiler's copies of its in-flight line stalled every step: ncu, 275 K of 2.2 M samples).
                if (k + 17 <= maxsteps) {
                    const uint8_t* lane_begin = in + (size_t)G.glane * seg;
                    const uint8_t* e = src.line + src.k;             // one past the next byte to hand out
                    const uint32_t s8 = ((uint32_t)reinterpret_cast<uintptr_t>(e) & 3u) * 8u;
                    const uint8_t* wa = e - (s8 >> 3);
                    const uint8_t* lo_word = lane_begin - (reinterpret_cast<uintptr_t>(lane_begin) & 3);
                    auto word = [&](const uint8_t* p) -> uint32_t {  // nothing below the word of the lane's first byte is touched
                        return (p >= lo_word) ? __ldg(reinterpret_cast<const uint32_t*>(p)) : 0u;
                    };
                    uint32_t hi = s8 ? word(wa) : 0u;                // the word holding e[-1] when e is not word aligned
                    uint32_t nq[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) nq[j] = word(wa - 4 * (j + 1));
                    wa -= 16;
                    const uint32_t ssym_a = smem_addr(ssym), srank_a = smem_addr(srank);
                    const uint32_t spk_a = smem_addr(gsm + NSCAP * NSCAP * 16 + 256);
                    const uint32_t tbits = act_s ? S->shift : 12u, Mo1 = 1u << tbits;
                    const uint32_t xs = (BYTE ? 23u + 8u : 15u + 16u) - tbits;      // x_max = freq << xs (make_sym)
                    for (; k + 17 <= maxsteps; k += 16) {
                        uint32_t c[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) c[j] = nq[j];
#pragma unroll
                        for (int j = 0; j < 4; j++) nq[j] = word(wa - 4 * (j + 1));
                        wa -= 16;
                        e -= 16;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t cur = __funnelshift_r(c[j], hi, s8);   // bytes e-4j-4 .. e-4j-1, the last one on top
                            hi = c[j];
                            EncSym sy[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const uint32_t rc = lds_u8(srank_a + __byte_perm(cur, 0, 0x4443 - u));
                                if (PK) {
                                    const uint2 v = lds_v2(spk_a + ((rc * ns + rs) << 3));
                                    const uint32_t fq = (v.y >> 13) & 0x1fffu;
                                    sy[u].rcp_freq = v.x;
                                    sy[u].bias = v.y & 0x1fffu;
                                    sy[u].x_max = fq << xs;
                                    sy[u].cmpl_shift = __byte_perm(v.y >> 26, Mo1 - fq, 0x5410);
                                } else {
                                    const uint4 v = lds_v4(ssym_a + ((rc * ns + rs) << 4));
                                    sy[u].x_max = v.x; sy[u].rcp_freq = v.y; sy[u].bias = v.z; sy[u].cmpl_shift = v.w;
                                }
                                rs = rc;
                            }
                            run4(sy);
                        }
                    }
                    src.init(lane_begin, e);                         // the remaining (< 16 + 4) steps go through the byte source
                }
                for (; k + 5 <= maxsteps; k += 4) {
                    EncSym sy[4];
                    gather(sy);
                    run4(sy);
                }
            } else if (k + 9 <= maxsteps) {                          // global tables: eight look-ups in flight (an L2 round trip
                EncSym ca[4], cb[4], na[4], nb[4];                   // is ~250 cycles, a state update ~40)
                gather(ca); gather(cb);
                for (; k + 17 <= maxsteps; k += 8) {
                    gather(na); gather(nb);
                    run4(ca); run4(cb);
#pragma unroll
                    for (int u = 0; u < 4; u++) { ca[u] = na[u]; cb[u] = nb[u]; }
                }
                run4(ca); run4(cb);
                k += 8;
            }
            for (; k + 1 < maxsteps; k++) {
                const uint32_t rc = srank[src.get()];
                const EncSym s = syms[rc * ns + rs];
                rs = rc;
                x = enc_step<NWAY, BYTE>(x, true, s, wp, G);
            }
            if (k < maxsteps && minsteps) {                          // segment starts: context 0
                const EncSym s = syms[(uint32_t)srank[0] * ns + rs];
                x = enc_step<NWAY, BYTE>(x, true, s, wp, G);
            }
        }
        // RansEncFlush (rANS_word.h:104-116): states NWAY-1 .. 0, so state 0 ends lowest
        if (act_s) {
            uint8_t* p = wp - 4 * (NWAY - G.glane);
            p[0] = (uint8_t)x; p[1] = (uint8_t)(x >> 8); p[2] = (uint8_t)(x >> 16); p[3] = (uint8_t)(x >> 24);
            if (G.glane == 0) {
                uint32_t pay_off = (uint32_t)(wp - 4 * NWAY - S->out);
                S->pay_off = pay_off;
                S->size = S->tab_len + (S->cap - pay_off);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// enc_finish_kernel: one CTA per leaf -- assemble the container (:1231-1342)
// ------------------------------------------------------------------------------------------
// CTA-wide copy with arbitrary mutual alignment: 16-byte aligned stores, the source read as
// aligned 32-bit words and realigned with funnel shifts.  src and dst must not overlap.
__device__ __forceinline__ void cta_copy(uint8_t* dst, const uint8_t* src, uint32_t n) {
    const uint32_t head = min(n, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
    const uint32_t nv = (n - head) / 16;
    const uint8_t* s0 = src + head;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(s0) & 3) * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s0 - (sh >> 3));
    uint4* dv = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) {
        const uint32_t* w = sw + 4 * (size_t)i;
        uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3];
        if (sh) {                                            // block-uniform
            const uint32_t a4 = w[4];                        // inside the source: the chunk ends sh/8 bytes into it
            a0 = __funnelshift_r(a0, a1, sh); a1 = __funnelshift_r(a1, a2, sh);
            a2 = __funnelshift_r(a2, a3, sh); a3 = __funnelshift_r(a3, a4, sh);
        }
        dv[i] = make_uint4(a0, a1, a2, a3);
    }
    for (uint32_t i = head + nv * 16 + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// The same for a destination BELOW an overlapping source (the payload of a stream built in place moves down behind its
// table): tile by tile, every thread reads its 16 bytes (aligned 32-bit source words, realigned with funnel shifts)
// before any thread writes the tile's 16-byte aligned stores.
__device__ __forceinline__ void cta_move_down(uint8_t* dst, const uint8_t* src, uint32_t n) {
    if (dst == src || n == 0) return;                       // (CTA-uniform)
    const uint32_t head = min(n, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    {                                                       // the bytes up to dst's first 16-byte boundary
        uint8_t v = 0;
        if (threadIdx.x < head) v = src[threadIdx.x];
        __syncthreads();
        if (threadIdx.x < head) dst[threadIdx.x] = v;
        __syncthreads();
    }
    const uint32_t nv = (n - head) / 16;
    const uint8_t* s0 = src + head;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(s0) & 3) * 8;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s0 - (sh >> 3));
    uint4* dv = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t i0 = 0; i0 < nv; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (i < nv) {
            const uint32_t* w = sw + 4 * (size_t)i;
            a0 = w[0]; a1 = w[1]; a2 = w[2]; a3 = w[3];
            if (sh) {                                       // block-uniform; w[4] is inside the source (the chunk ends sh/8 bytes into it)
                const uint32_t a4 = w[4];
                a0 = __funnelshift_r(a0, a1, sh); a1 = __funnelshift_r(a1, a2, sh);
                a2 = __funnelshift_r(a2, a3, sh); a3 = __funnelshift_r(a3, a4, sh);
            }
        }
        __syncthreads();
        if (i < nv) dv[i] = make_uint4(a0, a1, a2, a3);
        __syncthreads();
    }
    {                                                       // tail
        const uint32_t t0 = head + nv * 16;
        uint8_t v = 0;
        if (t0 + threadIdx.x < n) v = src[t0 + threadIdx.x];
        __syncthreads();
        if (t0 + threadIdx.x < n) dst[t0 + threadIdx.x] = v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) enc_finish_kernel(EncWork* W) {
    __shared__ uint32_t s_hdr, s_fail, s_cat;
    for (uint32_t li = blockIdx.x; li < W->nleaves; li += gridDim.x) {
        EncLeaf& L = W->leaves[li];
        __syncthreads();
        const EncStream& B = W->streams[L.body];
        uint8_t* out = L.out;
        if (!out) {                                                  // the block's region is too small (enc_fix_kernel)
            if (threadIdx.x == 0) { L.out_size = 0; L.status = ST_SIZE; }
            continue;
        }
        if (B.codec != 0) {
            // legacy rANS 4x8 block (rANS_static.c:197-215): [order][u32 size - 9][u32 n][table][payload]
            const bool bad = B.size == 0xffffffffu;
            const uint32_t pay = B.cap - B.pay_off, total = 9 + B.tab_len + pay;
            if (threadIdx.x == 0 && !bad) {
                out[0] = (uint8_t)B.order_eff;
                for (int k = 0; k < 4; k++) { out[1 + k] = (uint8_t)((total - 9) >> (8 * k)); out[5 + k] = (uint8_t)(L.n >> (8 * k)); }
            }
            if (!bad && B.hdr) cta_move_down(out + 9 + B.tab_len, B.out + B.pay_off, pay);   // table already in place
            else if (!bad) {
                cta_copy(out + 9, B.out, B.tab_len);
                cta_copy(out + 9 + B.tab_len, B.out + B.pay_off, pay);
            }
            if (threadIdx.x == 0) { L.out_size = bad ? 0u : total; L.status = bad ? ST_FORMAT : ST_OK; }
            continue;
        }
        if (threadIdx.x == 0) {
            uint32_t flags = L.out_flags;
            uint32_t hdr = 1;
            s_fail = 0;
            if (!(L.flags & F_NOSZ)) hdr += var_put_u32(out + hdr, L.n);                     // :1234
            if (flags & F_PACK) {                                                              // :1257-1262
                for (uint32_t k = 0; k < L.pmeta_len; k++) out[hdr + k] = L.pmeta[k];
                hdr += L.pmeta_len;
                hdr += var_put_u32(out + hdr, L.packed_len);
            }
            if (flags & F_RLE) {                                                               // :1295-1310
                const EncStream& M = W->streams[L.meta];
                if (M.size == 0xffffffffu) s_fail = 1;
                if (M.size < L.rmeta_len) {
                    hdr += var_put_u32(out + hdr, L.rmeta_len * 2);
                    hdr += var_put_u32(out + hdr, L.lit_len);
                    hdr += var_put_u32(out + hdr, M.size);
                } else {
                    hdr += var_put_u32(out + hdr, L.rmeta_len * 2 + 1);
                    hdr += var_put_u32(out + hdr, L.lit_len);
                }
            }
            if (B.order_eff == 0) flags &= ~1u;                                                // :1322-1325
            if (B.size == 0xffffffffu) s_fail = 1;
            const uint32_t cur_n = L.cur_n;
            s_cat = B.size >= cur_n;                                                           // :1332
            if (s_cat) flags = (flags & ~3u) | F_CAT | (L.flags & F_NOSZ);
            out[0] = (uint8_t)flags;
            s_hdr = hdr;
        }
        __syncthreads();
        uint32_t hdr = s_hdr;
        const uint32_t flags = out[0];
        if (flags & F_RLE) {
            const EncStream& M = W->streams[L.meta];
            if (M.size < L.rmeta_len) {
                cta_copy(out + hdr, M.out, M.tab_len);
                cta_copy(out + hdr + M.tab_len, M.out + M.pay_off, M.cap - M.pay_off);
                hdr += M.size;
            } else {
                cta_copy(out + hdr, L.rmeta + L.rmeta_off, L.rmeta_len);
                hdr += L.rmeta_len;
            }
        }
        const uint32_t cur_n = L.cur_n;
        uint32_t body;
        if (s_cat) { cta_copy(out + hdr, B.src, cur_n); body = cur_n; }
        else if (B.hdr) {                                            // built in place (hdr == B.hdr): the table is where it belongs
            cta_move_down(out + hdr + B.tab_len, B.out + B.pay_off, B.cap - B.pay_off);
            body = B.size;
        } else {
            cta_copy(out + hdr, B.out, B.tab_len);
            cta_copy(out + hdr + B.tab_len, B.out + B.pay_off, B.cap - B.pay_off);
            body = B.size;
        }
        if (threadIdx.x == 0) { L.out_size = hdr + body; L.status = s_fail ? ST_INTERNAL : ST_OK; }
    }
}

// ------------------------------------------------------------------------------------------
// enc_block_kernel: one CTA per block -- stripe winner selection / X_CAT / result reporting
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) enc_block_kernel(EncWork* W, uint32_t* out_len, int32_t* status) {
    __shared__ uint32_t win[256], woff[256];
    __shared__ uint32_t s_hdr;
    __shared__ int s_st;
    for (uint32_t b = blockIdx.x; b < W->nblocks; b += gridDim.x) {
        const EncBlock B = W->blocks[b];
        __syncthreads();
        if (B.mode == 3) { if (threadIdx.x == 0) { status[b] = ST_SIZE; out_len[b] = 0; } continue; }
        if (B.mode == 2) {                                                                     // :1218-1225
            if (threadIdx.x == 0) { B.out[0] = F_CAT; s_hdr = 1 + var_put_u32(B.out + 1, B.n); }
            __syncthreads();
            cta_copy(B.out + s_hdr, B.in, B.n);
            if (threadIdx.x == 0) { out_len[b] = s_hdr + B.n; status[b] = ST_OK; }
            continue;
        }
        if (B.mode == 0 || B.mode == 4) {
            if (threadIdx.x == 0) {
                const EncLeaf& L = W->leaves[B.leaf0];
                out_len[b] = L.out_size;
                status[b] = (L.status == ST_INTERNAL && W->overflow) ? ST_ARENA : L.status;
            }
            continue;
        }
        // stripe, :1182-1215
        if (threadIdx.x == 0) {
            uint32_t hdr = 1;
            int st = ST_OK;
            B.out[0] = (uint8_t)(B.order & ~F_NOSZ);
            hdr += var_put_u32(B.out + hdr, B.n);
            B.out[hdr++] = (uint8_t)B.N;
            uint32_t pay = 0;
            for (uint32_t j = 0; j < B.N; j++) {
                uint32_t best = B.n + 10, bi = 0;                    // strict '<' : the first smallest wins
                for (uint32_t c = 0; c < B.ncand; c++) {
                    const EncLeaf& L = W->leaves[B.leaf0 + j * B.ncand + c];
                    if (L.status != ST_OK) st = L.status;
                    if (best > L.out_size) { best = L.out_size; bi = c; }
                }
                win[j] = B.leaf0 + j * B.ncand + bi;
                woff[j] = pay;
                pay += best;
                hdr += var_put_u32(B.out + hdr, best);
            }
            s_hdr = hdr; s_st = st;
            out_len[b] = hdr + pay;
            status[b] = (st == ST_INTERNAL && W->overflow) ? ST_ARENA : st;
        }
        __syncthreads();
        for (uint32_t j = 0; j < B.N; j++) {
            const EncLeaf& L = W->leaves[win[j]];
            cta_copy(B.out + s_hdr + woff[j], L.out, L.out_size);
        }
    }
}

// Patches the pointers that depend on the caller's device-resident offset arrays.
// Also the capacity rule: like the reference's coders (rANS_static4x16pr.c:396-397, :706-707) a block whose output
// region is smaller than rans_compress_bound_4x16(n, order) is refused (ST_SIZE) and nothing is written to it.
__global__ void enc_fix_kernel(EncWork* W, const uint8_t* in_base, const uint64_t* in_off, uint8_t* out_base,
                               const uint64_t* out_off, const uint32_t* out_cap) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= W->nblocks) return;
    EncBlock& B = W->blocks[b];
    B.in = in_base + in_off[b];
    B.out = out_base + out_off[b];
    const bool fits = out_cap[b] >= B.cap;
    if (B.mode == 0 || B.mode == 4) {
        EncLeaf& L = W->leaves[B.leaf0];
        L.src = B.in; L.out = fits ? B.out : nullptr;
        EncStream& S = W->streams[L.body];
        S.src = B.in;
        if (S.hdr) {                                             // the stream is built in the block's own region
            if (!fits) S.n = 0;                                  // nothing is coded, nothing is written
            else {
                S.out = B.out + S.hdr;
                // the payload grows down from the last even ADDRESS of the region (16-bit stores)
                S.cap = (uint32_t)(((reinterpret_cast<uintptr_t>(B.out) + B.cap) & ~uintptr_t(1)) - reinterpret_cast<uintptr_t>(S.out));
            }
        }
    }
    if (!fits) B.mode = 3;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
namespace {

struct EncImpl {
    uint8_t* d_scratch = nullptr; size_t scratch_cap = 0;
    uint8_t* d_desc = nullptr; size_t desc_cap = 0;
    uint8_t* h_desc = nullptr; size_t h_desc_cap = 0;          // pinned
    cudaEvent_t uploaded = nullptr;
    bool pending = false;
    std::vector<uint32_t> h_len;
    std::vector<int32_t> h_order;
    SideStreams side;                                          // the rANS kernel variants run side by side
    // streams per variant of the previous batch (read back asynchronously): a batch that used a single variant
    // predicts another one, which is launched on the caller's stream alone (13 % faster than from a side stream)
    uint32_t* h_vcount = nullptr; cudaEvent_t vc_ready = nullptr; bool vc_pending = false; int mixed = -1;
    // arena for alphabet-dependent order-1 scratch (EncWork::arena) and the read-back of how much a batch wanted
    uint8_t* d_arena = nullptr; size_t arena_cap = 0;
    unsigned long long* h_ret = nullptr;                        // pinned: [arena_used, overflow]
    cudaEvent_t ret_ready = nullptr; bool ret_pending = false;
};

int g_sms_enc = 0;
// the largest arena any batch of the process asked for, in bytes (the chunk stages of a host-buffer call are separate
// slots); a batch never gets more than it could need (1.6 MB per order-1 stream)
std::atomic<size_t> g_arena_hint{0};
void raise_arena_hint(size_t v) {
    size_t cur = g_arena_hint.load();
    while (cur < v && !g_arena_hint.compare_exchange_weak(cur, v)) {}
}
int g_grid_enc[2][2];

unsigned int host_bound(unsigned int size, int order) {        // rans_compress_bound_4x16, :360-372
    int N = order >> 8;
    if (!N) N = 4;
    order &= 0xff;
    double d = 1.05 * size;
    d += (order == 0) ? (257 * 3 + 4) : (257 * 257 * 3 + 4 + 257 * 3 + 4);
    d += (order & F_PACK) ? 1 : 0;
    d += (order & F_RLE) ? (1 + 257 * 3 + 4) : 0;
    d += 20;
    d += (order & F_STRIPE) ? (1 + 5 * N) : 0;
    int sz = (int)d;
    return (unsigned int)(sz + (sz & 1) + 2);
}

unsigned int host_bound_4x8(unsigned int size) {               // the reference's malloc size, rANS_static.c:87
    return (unsigned int)(1.05 * size + 257 * 257 * 3 + 9);
}
constexpr int ORDER_LEGACY_4x8 = 0x40000000;                    // HTS_B200_ORDER_RANS4x8

size_t up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

__global__ void __launch_bounds__(256) enc_upload_kernel(uint4* dst, const uint4* src, uint32_t n16) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

template <typename K> int occ_grid(K kernel, int smem, int sms) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);   // launches may pad (shaping)
    int per = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, 32, smem);
    return std::max(per, 1) * sms;
}

constexpr int SM_O0_32 = ENC_O0_SMEM_PER_GROUP, SM_O0_4 = ENC_O0_SMEM_PER_GROUP * 8;
constexpr int SM_O1_32_S = enc_o1_smem(16), SM_O1_32_L = enc_o1_smem(48);           // small / large alphabet variants
constexpr int SM_O1_4_S = enc_o1_smem(16) * 8;                                       // 4-way: small only (larger: global)
constexpr int SM_O0_4_C = (256 + 48 * 16) * 8;                                       // 4-way order 0, compact symbol tables
constexpr int SM_O1_4_T = enc_o1_smem(9) * 8;                                        // 4-way order 1, <= 9 symbols
int g_grid_o0_4_c = 0, g_grid_o1_4_t = 0, g_grid_o0_8_c = 0, g_grid_o1_8_t = 0;
int g_grid_o1_32_s = 0, g_grid_o1_32_l = 0, g_grid_o1_4_s = 0;

}  // namespace

void EncSlot::release() {
    EncImpl* I = static_cast<EncImpl*>(impl);
    if (!I) return;
    if (I->d_scratch) cudaFree(I->d_scratch);
    if (I->d_desc) cudaFree(I->d_desc);
    if (I->h_desc) cudaFreeHost(I->h_desc);
    if (I->uploaded) cudaEventDestroy(I->uploaded);
    I->side.release();
    if (I->h_vcount) cudaFreeHost(I->h_vcount);
    if (I->vc_ready) cudaEventDestroy(I->vc_ready);
    if (I->d_arena) cudaFree(I->d_arena);
    if (I->h_ret) cudaFreeHost(I->h_ret);
    if (I->ret_ready) cudaEventDestroy(I->ret_ready);
    delete I;
    impl = nullptr;
}

// After the stream the batch ran on has drained: did it run out of arena?  If so the next encode_run gets a larger one
// and the caller should run the batch again (blocks that wanted the missing space carry ST_ARENA).
bool encode_needs_retry(EncSlot& slot) {
    EncImpl* I = static_cast<EncImpl*>(slot.impl);
    if (!I || !I->ret_pending) return false;
    if (cudaEventQuery(I->ret_ready) != cudaSuccess) { cudaGetLastError(); return false; }
    I->ret_pending = false;
    if (!(I->h_ret[1] & 0xffffffffull)) return false;
    raise_arena_hint((size_t)(I->h_ret[0] + I->h_ret[0] / 4 + (16 << 20)));
    return true;
}

size_t encode_device_bytes(const EncSlot& slot) {
    const EncImpl* I = static_cast<const EncImpl*>(slot.impl);
    return I ? I->scratch_cap + I->desc_cap + I->arena_cap : 0;
}

int encode_init(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    g_sms_enc = prop.multiProcessorCount;
    double l10[257], l12[257];
    for (int k = 0; k <= 256; k++) { l10[k] = log((double)(1024 + k)); l12[k] = log((double)(4096 + k)); }
    if (cudaMemcpyToSymbol(c_log10, l10, sizeof(l10)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_log12, l12, sizeof(l12)) != cudaSuccess) return -1;
    cudaFuncSetAttribute(enc_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HIST_SMEM);
    g_grid_enc[0][0] = occ_grid(enc_rans_kernel<4, 0, 16>, SM_O0_4, g_sms_enc);
    g_grid_enc[1][0] = occ_grid(enc_rans_kernel<32, 0, 16>, SM_O0_32, g_sms_enc);
    g_grid_o1_4_s = occ_grid(enc_rans_kernel<4, 1, 16>, SM_O1_4_S, g_sms_enc);
    occ_grid(enc_rans_kernel<4, 0, 16, true>, SM_O0_4, g_sms_enc);
    g_grid_o0_4_c = occ_grid(enc_rans_kernel<4, 0, 48>, SM_O0_4_C, g_sms_enc);
    g_grid_o1_4_t = occ_grid(enc_rans_kernel<4, 1, 9>, SM_O1_4_T, g_sms_enc);
    g_grid_o0_8_c = occ_grid(enc_rans_kernel<4, 0, 48, true>, SM_O0_4_C, g_sms_enc);
    g_grid_o1_8_t = occ_grid(enc_rans_kernel<4, 1, 9, true>, SM_O1_4_T, g_sms_enc);
    occ_grid(enc_rans_kernel<4, 1, 16, true>, SM_O1_4_S, g_sms_enc);
    g_grid_o1_32_s = occ_grid(enc_rans_kernel<32, 1, 16>, SM_O1_32_S, g_sms_enc);
    g_grid_o1_32_l = occ_grid(enc_rans_kernel<32, 1, 48>, SM_O1_32_L, g_sms_enc);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int encode_run(EncSlot& slot, const EncodeBatch& b, const uint32_t* h_in_len, const int32_t* h_order,
               cudaStream_t st, char* err, size_t errlen) {
#define ECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, errlen, "encode.cu:%d %s", __LINE__, cudaGetErrorString(e_)); return -1; } } while (0)
    if (!slot.impl) slot.impl = new EncImpl();
    EncImpl* I = static_cast<EncImpl*>(slot.impl);
    if (!I->uploaded) ECK(cudaEventCreateWithFlags(&I->uploaded, cudaEventDisableTiming));
    const int nblk = b.nblk;
    if (!h_in_len || !h_order) {                                  // device-resident API: fetch the two small arrays
        I->h_len.resize(nblk); I->h_order.resize(nblk);
        ECK(cudaMemcpyAsync(I->h_len.data(), b.in_len, 4 * (size_t)nblk, cudaMemcpyDeviceToHost, st));
        ECK(cudaMemcpyAsync(I->h_order.data(), b.order, 4 * (size_t)nblk, cudaMemcpyDeviceToHost, st));
        ECK(cudaStreamSynchronize(st));
        h_in_len = I->h_len.data(); h_order = I->h_order.data();
    }

    // ---- expand blocks into leaves and streams; lay out the scratch arena (offsets first)
    std::vector<EncBlock> blocks(nblk);
    std::vector<EncLeaf> leaves;
    std::vector<EncStream> streams;
    std::vector<uint64_t> in_off_needed;                           // (unused: offsets are read on the device)
    size_t scratch = 0;
    auto take = [&](size_t bytes) { size_t at = scratch; scratch += up(bytes + 32); return at; };
    // Device pointers are formed as (uint8_t*)offset + tag; fixed up in a tiny kernel-free way below:
    // leaves/streams store offsets relative to scratch base (tag 1) or to the block's in/out (tag 2).
    struct Fix { uint32_t kind; uint32_t idx; uint32_t field; uint32_t blk; };
    (void)in_off_needed;

    // hdr != 0: the stream is built in the block's own output region (see EncStream::hdr), no output scratch
    size_t n_o1 = 0;
    auto add_stream = [&](uint32_t leaf, uint32_t n_max, uint32_t order, uint32_t nway, uint32_t hdr, uint32_t codec = 0) {
        EncStream S;
        memset(&S, 0, sizeof(S));
        S.n = n_max; S.order = order; S.nway = nway; S.leaf = leaf; S.codec = codec; S.hdr = hdr;
        if (!hdr) {
            uint32_t cap = ((codec ? host_bound_4x8(n_max) : host_bound(n_max, order)) + 4 * nway + 64) & ~1u;
            S.cap = cap;
            S.out = reinterpret_cast<uint8_t*>(take(cap));
        }
        S.F0 = reinterpret_cast<uint32_t*>(take(1024));
        // 256 encoder symbols / pair counts: order 0, and order-1 alphabets of up to 16 symbols (16 x 16 entries);
        // larger alphabets get theirs from the arena once the histogram pass knows the size (enc_hist_kernel)
        S.syms = reinterpret_cast<EncSym*>(take(256 * 16));
        if (order) { S.F1 = reinterpret_cast<uint32_t*>(take(256 * 4)); n_o1++; }
        streams.push_back(S);
        return (uint32_t)streams.size() - 1;
    };
    // src/out of leaves: encoded as offset | tag in the low bits is fragile; keep explicit side tables
    std::vector<uint8_t> leaf_src_is_scratch, leaf_out_is_scratch;
    std::vector<uint64_t> leaf_src_off, leaf_out_off;

    auto add_leaf = [&](uint32_t blk, uint32_t n, uint32_t flags, bool src_scratch, uint64_t src_off, bool out_scratch,
                        uint64_t out_off, uint32_t codec = 0) {
        EncLeaf L;
        memset(&L, 0, sizeof(L));
        L.n = n; L.flags = flags; L.blk = blk; L.out_flags = flags; L.cur_n = n;
        if (flags & F_PACK) L.packed = reinterpret_cast<uint8_t*>(take((size_t)n + 16));
        if (flags & F_RLE) {
            L.lits = reinterpret_cast<uint8_t*>(take((size_t)n + 16));
            L.rmeta = reinterpret_cast<uint8_t*>(take((size_t)n + 272 + 64));
        }
        const uint32_t li = (uint32_t)leaves.size();
        const uint32_t nway = (flags & F_X32) ? 32 : 4;
        // a plain leaf written straight to the caller's block: [flags][varint n] (or the 9-byte 4x8 header), then the stream
        uint32_t hdr = 0;
        if (!src_scratch && !out_scratch && !(flags & (F_PACK | F_RLE))) {
            hdr = codec ? 9u : 1u;
            if (!codec && !(flags & F_NOSZ)) for (uint32_t t = n;; t >>= 7) { hdr++; if (t < 128) break; }
        }
        L.body = add_stream(li, n, flags & 1, nway, hdr, codec);
        L.meta = (flags & F_RLE) ? add_stream(li, n + 272, 0, nway, 0) : 0xffffffffu;
        leaves.push_back(L);
        leaf_src_is_scratch.push_back(src_scratch); leaf_src_off.push_back(src_off);
        leaf_out_is_scratch.push_back(out_scratch); leaf_out_off.push_back(out_off);
        return li;
    };

    std::vector<uint64_t> blk_tr_off(nblk, 0);
    for (int i = 0; i < nblk; i++) {
        EncBlock& B = blocks[i];
        memset(&B, 0, sizeof(B));
        uint32_t n = h_in_len[i];
        int order = h_order[i];
        B.n = n; B.order = (uint32_t)order;
        if (order & ORDER_LEGACY_4x8) {                            // legacy CRAM 3.0 codec: order 0 / 1 only
            B.cap = host_bound_4x8(n);
            B.mode = 4;
            B.leaf0 = add_leaf(i, n, (uint32_t)(order & 0xff) ? 1u : 0u, false, 0, false, 0, 1);
            continue;
        }
        B.cap = host_bound(n, order);
        if (n <= 20) order &= ~(int)F_STRIPE;                      // :1151
        if (order & F_STRIPE) {
            int N = order >> 8;
            if (N == 0) N = 4;
            if (N > 255) { B.mode = 3; continue; }                 // :1158
            B.mode = 1; B.N = (uint32_t)N;
            blk_tr_off[i] = take((size_t)n + 16);
            static const int cand[4] = {1, 64, 128, 0};            // :1192
            uint32_t ncand = 0;
            int use[4];
            for (int c = 0; c < 4; c++) if ((order & cand[c]) == cand[c]) use[ncand++] = cand[c];
            B.ncand = ncand;
            B.leaf0 = (uint32_t)leaves.size();
            uint64_t at = 0;
            for (int j = 0; j < N; j++) {
                uint32_t len = n / N + ((n % N) > (uint32_t)j);
                for (uint32_t c = 0; c < ncand; c++) {
                    uint32_t fl = (uint32_t)use[c] | F_NOSZ | ((uint32_t)order & F_X32);
                    size_t o = take(host_bound(len, (int)fl) + 64);
                    add_leaf(i, len, fl, true, blk_tr_off[i] + at, true, o);
                }
                at += len;
            }
        } else if (order & F_CAT) {
            B.mode = 2;
        } else {
            B.mode = 0;
            B.leaf0 = add_leaf(i, n, (uint32_t)order & 0xff, false, 0, false, 0);
        }
    }

    // ---- device memory
    const size_t desc_bytes = up(sizeof(EncWork)) + up(sizeof(EncBlock) * blocks.size()) + up(sizeof(EncLeaf) * leaves.size()) +
                              up(sizeof(EncStream) * streams.size()) + up(4 * 16 * streams.size()) + 256;
    if (I->pending) { ECK(cudaEventSynchronize(I->uploaded)); I->pending = false; }
    // ---- arena for alphabet-dependent order-1 scratch.  What a finished batch wanted sizes the next one (an
    // asynchronous caller cannot be retried; encode_needs_retry does it for the synchronous ones).
    if (I->ret_pending && cudaEventQuery(I->ret_ready) == cudaSuccess) {
        I->ret_pending = false;
        if (I->h_ret[1] & 0xffffffffull) raise_arena_hint((size_t)(I->h_ret[0] + I->h_ret[0] / 4 + (16 << 20)));
    }
    cudaGetLastError();
    {
        // default: 64 KB per order-1 stream (an alphabet of up to 56 symbols) + 64 MB (forty 256-symbol streams)
        const size_t want = std::max(n_o1 * (size_t)(64 << 10) + (64 << 20),
                                     std::min(g_arena_hint.load(), n_o1 * (size_t)(1600 << 10) + (64 << 20)));
        if (want > I->arena_cap) {
            if (I->d_arena) cudaFree(I->d_arena);
            I->d_arena = nullptr; I->arena_cap = 0;
            if (cudaMalloc(&I->d_arena, want) != cudaSuccess) { cudaGetLastError(); snprintf(err, errlen, "out of device memory for the encode arena (%zu B)", want); return -1; }
            I->arena_cap = want;
        }
        if (!I->h_ret) {
            if (cudaHostAlloc(&I->h_ret, 16, cudaHostAllocDefault) != cudaSuccess || cudaEventCreateWithFlags(&I->ret_ready, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError(); snprintf(err, errlen, "out of memory for the encode read-back"); return -1;
            }
            I->h_ret[0] = I->h_ret[1] = 0;
        }
    }
    if (scratch + 256 > I->scratch_cap) {
        if (I->d_scratch) cudaFree(I->d_scratch);
        I->d_scratch = nullptr; I->scratch_cap = 0;
        size_t want = scratch + scratch / 8 + (1 << 20);
        if (cudaMalloc(&I->d_scratch, want) != cudaSuccess) { cudaGetLastError(); snprintf(err, errlen, "out of device memory for encode scratch (%zu B)", want); return -1; }
        I->scratch_cap = want;
    }
    if (desc_bytes > I->desc_cap) {
        if (I->d_desc) cudaFree(I->d_desc);
        if (I->h_desc) cudaFreeHost(I->h_desc);
        I->d_desc = nullptr; I->h_desc = nullptr; I->desc_cap = 0;
        size_t want = desc_bytes + desc_bytes / 4;
        if (cudaMalloc(&I->d_desc, want) != cudaSuccess || cudaHostAlloc(&I->h_desc, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError(); snprintf(err, errlen, "out of memory for encode descriptors"); return -1;
        }
        I->desc_cap = want;
    }
    // ---- resolve offsets into device pointers.  Block in/out pointers need in_off/out_off, which
    // live on the device: a fix-up kernel patches them (enc_fix_kernel below).
    uint8_t* sb = I->d_scratch;
    for (auto& S : streams) {
        if (!S.hdr) S.out = sb + (size_t)S.out;
        S.F0 = reinterpret_cast<uint32_t*>(sb + (size_t)S.F0);
        S.syms = reinterpret_cast<EncSym*>(sb + (size_t)S.syms);
        if (S.order) S.F1 = reinterpret_cast<uint32_t*>(sb + (size_t)S.F1);
    }
    for (size_t k = 0; k < leaves.size(); k++) {
        EncLeaf& L = leaves[k];
        if (L.flags & F_PACK) L.packed = sb + (size_t)L.packed;
        if (L.flags & F_RLE) { L.lits = sb + (size_t)L.lits; L.rmeta = sb + (size_t)L.rmeta; }
        L.src = leaf_src_is_scratch[k] ? sb + leaf_src_off[k] : nullptr;      // nullptr: patched from the block
        L.out = leaf_out_is_scratch[k] ? sb + leaf_out_off[k] : nullptr;
        streams[L.body].src = L.src;
    }
    for (int i = 0; i < nblk; i++) if (blocks[i].mode == 1) blocks[i].tr = sb + blk_tr_off[i];

    size_t o = 0;
    EncWork hw;
    memset(&hw, 0, sizeof(hw));
    const size_t o_work = o; o += up(sizeof(EncWork));
    const size_t o_blocks = o; o += up(sizeof(EncBlock) * blocks.size());
    const size_t o_leaves = o; o += up(sizeof(EncLeaf) * leaves.size());
    const size_t o_streams = o; o += up(sizeof(EncStream) * streams.size());
    const size_t o_vlist = o;                                    // device-only, not uploaded
    hw.blocks = reinterpret_cast<EncBlock*>(I->d_desc + o_blocks);
    hw.leaves = reinterpret_cast<EncLeaf*>(I->d_desc + o_leaves);
    hw.streams = reinterpret_cast<EncStream*>(I->d_desc + o_streams);
    hw.vlist = reinterpret_cast<uint32_t*>(I->d_desc + o_vlist);
    hw.nblocks = (uint32_t)blocks.size(); hw.nleaves = (uint32_t)leaves.size(); hw.nstreams = (uint32_t)streams.size();
    // 4-way streams of a batch too large for one wave of the 4 KB-table kernels (48 streams per SM) go to
    // variants with compacted tables by alphabet size: order 0 with <= 48 symbols (1 KB per stream),
    // order 1 with <= 9 symbols (1.5 KB)
    const bool big4 = (streams.size() + 7) / 8 > (size_t)g_grid_enc[0][0];
    hw.o0_lo = big4 ? 48u : 0u; hw.o1_lo = big4 ? 9u : 0u;
    hw.arena = I->d_arena; hw.arena_cap = I->arena_cap; hw.arena_used = 0; hw.overflow = 0;
    memcpy(I->h_desc + o_work, &hw, sizeof(hw));
    memcpy(I->h_desc + o_blocks, blocks.data(), sizeof(EncBlock) * blocks.size());
    if (!leaves.empty()) memcpy(I->h_desc + o_leaves, leaves.data(), sizeof(EncLeaf) * leaves.size());
    if (!streams.empty()) memcpy(I->h_desc + o_streams, streams.data(), sizeof(EncStream) * streams.size());
    // The descriptors go up by a KERNEL reading the pinned host copy (unified addressing), not through the copy engine:
    // in the host-buffer pipeline that engine is busy with the following chunks' input for as long as the host keeps it
    // fed, and a small copy on the compute stream queued behind all of them -- the chunk trace showed every chunk's
    // kernels starting only once the fourth chunk's input had landed (38 -> see DESIGN.md 4.1).
    enc_upload_kernel<<<64, 256, 0, st>>>(reinterpret_cast<uint4*>(I->d_desc), reinterpret_cast<const uint4*>(I->h_desc), (uint32_t)((o + 15) / 16));
    ECK(cudaEventRecord(I->uploaded, st));
    I->pending = true;

    EncWork* dW = reinterpret_cast<EncWork*>(I->d_desc + o_work);
    int launches = 1;                                            // enc_upload_kernel
    const int g = g_sms_enc * 4;
    enc_fix_kernel<<<(nblk + 127) / 128, 128, 0, st>>>(dW, b.in_base, b.in_off, b.out_base, b.out_off, b.out_len); launches++;
    bool any_stripe = false, any_tr = false, any32[2] = {false, false}, any4[2] = {false, false}, any8[2] = {false, false};
    for (auto& B : blocks) any_stripe |= B.mode == 1;
    for (auto& L : leaves) any_tr |= (L.flags & (F_PACK | F_RLE)) != 0;
    for (auto& S : streams) {
        if (S.codec) { any8[0] = true; any8[1] |= S.order != 0; }
        else if (S.nway == 32) { any32[0] = true; any32[1] |= S.order != 0; }
        else { any4[0] = true; any4[1] |= S.order != 0; }
    }
    if (any_stripe) { enc_stripe_kernel<<<g, 256, 0, st>>>(dW); launches++; }
    if (any_tr) { enc_transform_kernel<<<g, TT, 0, st>>>(dW); launches++; }
    if (!streams.empty()) {
        enc_hist_kernel<<<g_sms_enc * 3, HT, HIST_SMEM, st>>>(dW); launches++;
        enc_table_kernel<<<g * 2, KT, 0, st>>>(dW); launches++;
        enc_bucket_kernel<<<16, 256, 0, st>>>(dW); launches++;
        // an order-1 request can fall back to order 0 on the device, so the order-0 kernels always run
        // cursors: next_misc[0..4]; order-1 streams go to the small-alphabet variant (ns <= 16, tables in
        // shared memory at high occupancy) or the large one (ns > 16: shared up to 48 symbols, else global)
        // Shaped launches: a persistent kernel gives a warp to EG::G streams and the CTA scheduler fills
        // one SM before the next, so the dynamic shared-memory request is padded until exactly
        // c = ceil(groups / SMs) CTAs fit per SM and the grid is SMs x c (every SM holds the same load).
        const uint32_t ngroups32 = (uint32_t)streams.size(), ngroups4 = ((uint32_t)streams.size() + 7) / 8;
        auto shaped = [&](int cap, int smem, uint32_t groups, int* grid) {
            int c = (int)((groups + g_sms_enc - 1) / g_sms_enc);
            c = std::max(1, std::min(c, cap));
            *grid = g_sms_enc * c;
            static const int min_c = getenv("HTSCODECS_B200_MIN_C") ? atoi(getenv("HTSCODECS_B200_MIN_C")) : 4;   // (see shaped_launch)
            if (c < min_c) { c = std::min(min_c, cap); *grid = std::min(g_sms_enc * c, (int)std::max(groups, 1u)); }
            return c < cap ? std::max(smem, std::min(232448, (233472 / c - 1024) & ~127)) : smem;
        };
        int grid = 0, sm = 0;
        static const bool side_on = !(getenv("HTSCODECS_B200_SIDE") && atoi(getenv("HTSCODECS_B200_SIDE")) == 0);
        if (I->vc_pending && cudaEventQuery(I->vc_ready) == cudaSuccess) {
            int used = 0;
            for (int v = 0; v < 16; v++) used += I->h_vcount[v] != 0;
            I->mixed = used > 1;
            I->vc_pending = false;
        }
        SideStreams* side = (side_on && I->mixed != 0 && I->side.init() == 0) ? &I->side : nullptr;
        int nside = 0;
        cudaStream_t ks = st;
        if (side) cudaEventRecord(side->fork, st);
#define LAUNCH_ENC(CAP, SMEM, GROUPS, KERNEL, CURSOR)                                            \
        {                                                                                          \
            if (side) { ks = side->s[nside]; cudaStreamWaitEvent(ks, side->fork, 0); }             \
            sm = shaped((CAP) / g_sms_enc, SMEM, GROUPS, &grid);                                   \
            KERNEL<<<grid, 32, sm, ks>>>(dW, CURSOR); launches++;                                  \
            if (side) { cudaEventRecord(side->join[nside], ks); nside++; }                         \
        }
#define K(...) enc_rans_kernel<__VA_ARGS__>
        // the long-latency (4-way, order-1) variants first
        // (alphabets beyond 16 symbols read their tables from global memory at twice the step time: a launch of
        // their own, cursors 11 / 12, so that no warp mixes the two speeds)
        // variant (cursor): alphabet class, as bucketed by enc_bucket_kernel
        if (any4[1]) { LAUNCH_ENC(g_grid_o1_4_s, SM_O1_4_S, ngroups4, K(4, 1, 16), 11)                     // > 16 symbols
                       LAUNCH_ENC(g_grid_o1_4_s, SM_O1_4_S, ngroups4, K(4, 1, 16), 1)                      // <= 16
                       if (big4) LAUNCH_ENC(g_grid_o1_4_t, SM_O1_4_T, ngroups4, K(4, 1, 9), 8) }           // <= 9, large batch
        if (any8[1]) { LAUNCH_ENC(g_grid_o1_4_s, SM_O1_4_S, ngroups4, K(4, 1, 16, true), 12)
                       LAUNCH_ENC(g_grid_o1_4_s, SM_O1_4_S, ngroups4, K(4, 1, 16, true), 6)
                       if (big4) LAUNCH_ENC(g_grid_o1_8_t, SM_O1_4_T, ngroups4, K(4, 1, 9, true), 10) }
        if (any4[0]) { LAUNCH_ENC(g_grid_enc[0][0], SM_O0_4, ngroups4, K(4, 0, 16), 0)
                       if (big4) LAUNCH_ENC(g_grid_o0_4_c, SM_O0_4_C, ngroups4, K(4, 0, 48), 7) }          // <= 48, large batch
        if (any8[0]) { LAUNCH_ENC(g_grid_enc[0][0], SM_O0_4, ngroups4, K(4, 0, 16, true), 5)
                       if (big4) LAUNCH_ENC(g_grid_o0_8_c, SM_O0_4_C, ngroups4, K(4, 0, 48, true), 9) }
        if (any32[1]) { LAUNCH_ENC(g_grid_o1_32_l, SM_O1_32_L, ngroups32, K(32, 1, 48), 4)                 // > 16 symbols
                        LAUNCH_ENC(g_grid_o1_32_s, SM_O1_32_S, ngroups32, K(32, 1, 16), 3) }
        if (any32[0]) {                                              // the throughput-bound variant stays on the caller's stream
            sm = shaped(g_grid_enc[1][0] / g_sms_enc, SM_O0_32, ngroups32, &grid);
            enc_rans_kernel<32, 0, 16><<<grid, 32, sm, st>>>(dW, 2); launches++;
        }
#undef K
#undef LAUNCH_ENC
        for (int i = 0; i < nside; i++) cudaStreamWaitEvent(st, side->join[i], 0);
        enc_finish_kernel<<<g, 256, 0, st>>>(dW); launches++;
        if (!I->h_vcount) { cudaHostAlloc(&I->h_vcount, 64, cudaHostAllocDefault); cudaEventCreateWithFlags(&I->vc_ready, cudaEventDisableTiming); }
        if (I->h_vcount && I->vc_ready) {
            cudaMemcpyAsync(I->h_vcount, dW->vcount, 64, cudaMemcpyDeviceToHost, st);
            cudaEventRecord(I->vc_ready, st);
            I->vc_pending = true;
        }
    }
    enc_block_kernel<<<g, 256, 0, st>>>(dW, b.out_len, b.status); launches++;
    ECK(cudaGetLastError());
    // how much arena the batch wanted, and whether it ran out (read by encode_needs_retry / the next call)
    ECK(cudaMemcpyAsync(&I->h_ret[0], &dW->arena_used, 8, cudaMemcpyDeviceToHost, st));
    ECK(cudaMemcpyAsync(&I->h_ret[1], &dW->overflow, 4, cudaMemcpyDeviceToHost, st));
    ECK(cudaEventRecord(I->ret_ready, st));
    I->ret_pending = true;
    return launches;
#undef ECK
}

}  // namespace hb
