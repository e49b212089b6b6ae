// decode.cu -- sm_100a kernels for batched rANS Nx16 / 4x8 decode and the inverse transforms.
//
// Pipeline for one batch (one stream, no host round trip in between):
//   plan_kernel      1 thread / block: parses the container (rANS_static4x16pr.c:1352-1572),
//                    emits jobs, chains and stripe ops, bump-allocates temporaries from the arena
//   dec_o0_kernel    persistent, 1 warp / CTA: order-0 entropy decode  (…4x16pr.c:501-616, rANS_static.c:225-363)
//   dec_o1_kernel    persistent, 1 warp / CTA: order-1 entropy decode  (…4x16pr.c:870-1130, rANS_static.c:676-922)
//   copy_kernel      X_CAT bodies                                      (…4x16pr.c:1578-1584)
//   rle_kernel       un-RLE  (rle.c:142-187)
//   unpack_kernel    un-PACK (pack.c:211-348)
//   unstripe_kernel  byte de-interleave (utils.h:41-73)
//
// Lane mapping of the entropy kernels: a *group* of NWAY lanes owns one job and lane z holds
// rANS state z.  NWAY = 32 (X_32): one job per warp.  NWAY = 4: eight independent jobs per
// warp.  Every step each lane decodes one symbol through shared-memory tables, then the lanes
// whose state fell below the lower bound fetch their renormalisation words from a
// shared-memory ring: __ballot_sync gives the set of hungry lanes and __popc of the lower lanes
// gives each lane its word index (the format stores the words in state order,
// rANS_word.h:356-410).  The ring is refilled by coalesced 128-bit loads issued half a ring
// ahead.  Table set-up runs per group (group-masked warp syncs); the decode loop runs converged.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "decode.h"

namespace hb {

// ------------------------------------------------------------------------------------------
// arena + list helpers (device)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t* arena_alloc(DecWork* W, uint64_t bytes) {
    bytes = (bytes + 255) & ~255ull;
    unsigned long long at = atomicAdd(&W->arena_used, (unsigned long long)bytes);
    if (at + bytes > W->arena_cap) return nullptr;
    return W->arena + at;
}

__device__ __forceinline__ bool push_job(DecWork* W, uint32_t kind, const DecJob& j) {
    // the host launches only the kinds its hint names: a job outside the hint would never run, so its block fails
    // (status ST_ARENA without the overflow flag: no retry) rather than report success over unwritten output
    if (!((W->kinds >> kind) & 1u)) return false;
    uint32_t at = atomicAdd(&W->njobs[kind], 1u);
    if (at >= W->job_cap) { W->overflow = 1; return false; }
    W->jobs[kind][at] = j;
    return true;
}

__device__ __forceinline__ DecJob make_job(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len,
                                           uint32_t blk) {
    DecJob j;
    j.in = in; j.in_len = in_len; j.out = out; j.out_len = out_len; j.blk = blk; j.fuse = 0; j.aux = nullptr;
    j.fin_len = 0; j.pad = 0;
    for (int k = 0; k < 16; k++) j.map[k] = 0;
    return j;
}

// ------------------------------------------------------------------------------------------
// plan_kernel
// ------------------------------------------------------------------------------------------
struct PlanArgs {
    DecWork* W;
    const uint8_t* in_base;
    const uint64_t* in_off;
    const uint32_t* in_len;
    uint8_t* out_base;
    const uint64_t* out_off;
    uint32_t* out_len;
    int32_t* status;
    const uint8_t* method;
    int nblk;
};

// hts_unpack_meta, pack.c:165-198.  Returns bytes consumed (0 = failure).
__device__ int unpack_meta(const uint8_t* d, uint32_t len, uint8_t* map, uint32_t* per) {
    if (!len) return 0;
    uint32_t ns = d[0] ? d[0] : 256;
    if (ns <= 1) *per = 0; else if (ns <= 2) *per = 8; else if (ns <= 4) *per = 4; else if (ns <= 16) *per = 2;
    else { *per = 1; return 1; }
    if (len <= 1) return 0;
    uint32_t c = 0, j = 1;
    do { map[c++] = d[j++]; } while (c < ns && j < len);
    return c < ns ? 0 : (int)j;
}

// compact order-1 tables (see the order-1 section): bytes needed for `ns` symbols, and the table
// budget of the small-alphabet X_32 kernel variant
__host__ __device__ constexpr uint32_t o1_compact_bytes(uint32_t ns) { return ns * 64 + ns * 4 * (ns + 3); }
constexpr uint32_t O1_SMALL_TAB = 4608;
constexpr uint32_t O1_TAB4 = 3072, O1_TAB4M = 12544;   // compact-table bytes per 4-way stream: regular / medium variant

// bounded reader over global-memory bytes
struct GRd {
    const uint8_t* p; const uint8_t* end;
    __device__ __forceinline__ bool more() const { return p < end; }
    __device__ __forceinline__ uint32_t peek() const { return p < end ? *p : 0u; }
    __device__ __forceinline__ uint32_t get() { uint32_t v = peek(); p++; return v; }
    __device__ __forceinline__ uint32_t varint() {
        uint32_t x = 0;
        if (p >= end) return 0;
        for (;;) {
            uint32_t c = get();
            x = (x << 7) | (c & 0x7f);
            if (!(c & 0x80) || p >= end) break;
        }
        return x;
    }
};

// One 4x8 "sym [run] freq ... 0" table (rANS_static.c:271-303; zero_is_4096 for the order-1
// inner tables, :775-776).  Returns false when malformed; *sum = frequency total.
template <typename RD, typename ST>
__device__ bool parse_table_4x8(RD& r, ST store, uint32_t* sum, bool zero_is_4096) {
    uint32_t run = 0, x = 0, j = r.get();
    do {
        uint32_t f = r.get();
        if (f >= 128) f = ((f & 127) << 8) | r.get();
        if (!f && zero_is_4096) f = 4096;
        if (x + f > 4096) return false;
        store(j, f);
        x += f;
        if (!run && j + 1 == r.peek()) { r.get(); j++; run = r.get(); }
        else if (run) { run--; if (++j > 255) return false; }
        else j = r.get();
        if (!r.more()) return false;
    } while (j);
    *sum = x;
    return true;
}
template <typename RD>
__device__ bool parse_table_4x8(RD& r, uint32_t F, uint32_t* sum, bool zero_is_4096) {
    return parse_table_4x8(r, [&](uint32_t j, uint32_t f) { sts_u32(F + 4 * j, f); }, sum, zero_is_4096);
}

// decode_alphabet (rANS_static4x16pr.c:208-255) as a counter: number of symbols listed (an upper
// bound on distinct symbols; exact for streams whose list is strictly increasing, as written by
// the encoder).  Returns false when the bytes run out.
template <typename E>
__device__ bool list_alphabet(GRd& r, uint32_t* ns, E emit) {
    if (!r.more()) return false;
    uint32_t run = 0, j = r.get(), n = 0;
    do {
        emit(j);
        n++;
        if (!r.more()) return false;
        if (!run && j + 1 == r.peek()) {
            r.get();
            if (!r.more()) return false;
            j++;
            run = r.get();
        } else if (run) {
            run--;
            if (++j > 255) return false;
        } else {
            j = r.get();
        }
    } while (j && r.more() && n < 512);
    *ns = n;
    return true;
}
__device__ bool count_alphabet(GRd& r, uint32_t* ns) { return list_alphabet(r, ns, [](uint32_t) {}); }
constexpr uint32_t O0C_MAX_NS = 64;     // alphabet limit of the compact order-0 kernels
constexpr uint32_t O1_SENTINEL = 0xfff00fffu;   // compact-table row terminator: last slot 0xfff, rank 0, F 4096

// One non-striped container, rANS_static4x16pr.c:1435-1629, turned into a plan.  `cap` is the
// caller's capacity (the exact size for X_NOSZ); expect != 0xffffffff marks a stripe sub-stream
// that must produce exactly that many bytes.  `ci` is the chain slot reserved by the caller.
__device__ int32_t plan_chain(DecWork* W, const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t cap,
                              uint32_t expect, uint32_t blk, uint32_t* out_len_slot, uint32_t ci) {
    Chain c;
    c.t1 = c.t2 = c.t3 = out; c.meta = nullptr;
    c.blk = blk; c.flags = 0; c.u_meta = 0; c.t1_size = 0; c.osz = 0; c.t2_size = 0;
    c.final_size = 0xfffffffeu; c.expect = expect; c.per = 0;
    for (int k = 0; k < 16; k++) c.map[k] = 0;
    W->chains[ci] = c;                                               // a failed plan leaves a harmless chain

    if (in_len == 0) return ST_FORMAT;
    const uint8_t* end = in + in_len;
    uint32_t flags = *in++; in_len--;
    if (flags & F_STRIPE) return ST_NESTED;
    uint32_t osz;
    if (!(flags & F_NOSZ)) {
        int s = var_get_u32(in, end, &osz);
        in += s; in_len -= s;
    } else {
        osz = cap;
    }
    if (cap < osz) return ST_SIZE;                                   // :1464
    if (osz >= 0x7fffffffu) return ST_FORMAT;                        // :506
    c.flags = flags & (F_RLE | F_PACK);
    c.osz = osz;

    uint32_t t1_size = osz;

    if (flags & F_PACK) {                                            // :1527-1545
        int m = unpack_meta(in, in_len, c.map, &c.per);
        if (!m) return ST_FORMAT;
        in += m; in_len -= m;
        uint32_t psz;
        int s = var_get_u32(in, end, &psz);
        if ((uint32_t)s > in_len) return ST_FORMAT;
        in += s; in_len -= s;
        if (psz > t1_size) return ST_FORMAT;
        t1_size = psz;
    }
    // X_PACK alone over an order-0 stream: the decoder expands every byte as it produces it (no intermediate buffer,
    // no un-PACK pass)
    const bool fused = (flags & F_PACK) && !(flags & (F_RLE | F_CAT | F_ORDER1)) && c.per >= 2 && in_len > 0;
    if (fused && ((uint64_t)osz + c.per - 1) / c.per > t1_size) return ST_FORMAT;      // pack.c:238,279,314
    uint8_t* tmp = nullptr;
    if ((flags & (F_PACK | F_RLE)) && !fused) {                      // :1498-1513
        tmp = arena_alloc(W, (uint64_t)osz + 16);
        if (!tmp) return ST_ARENA;
    }
    if ((flags & F_PACK) && (flags & F_RLE)) { c.t1 = out; c.t2 = tmp; c.t3 = out; }
    else if (fused)                          { c.t1 = out; c.t2 = out; c.t3 = out; }
    else if (flags & F_PACK)                 { c.t1 = tmp; c.t2 = tmp; c.t3 = out; }
    else if (flags & F_RLE)                  { c.t1 = tmp; c.t2 = out; c.t3 = out; }

    const bool x32 = (flags & F_X32) != 0;
    if (flags & F_RLE) {                                             // :1549-1572
        uint32_t u_meta, rle_len, c_meta;
        uint32_t s = var_get_u32(in, end, &u_meta);
        if (s > in_len) return ST_FORMAT;
        s += var_get_u32(in + s, end, &rle_len);
        if (s > in_len) return ST_FORMAT;
        if (rle_len > t1_size) return ST_FORMAT;
        if (u_meta & 1) {
            c.meta = in + s;
            uint32_t avail = (uint32_t)(end - c.meta);
            u_meta = (u_meta / 2 > avail) ? avail : u_meta / 2;
            c_meta = u_meta;
        } else {
            s += var_get_u32(in + s, end, &c_meta);
            u_meta /= 2;
            if (s > in_len) return ST_FORMAT;
            // A valid encoder keeps lits+meta < .99*size (:1287), so the meta never exceeds osz.
            if (u_meta > osz + 1024u) return ST_FORMAT;
            uint8_t* mb = arena_alloc(W, (uint64_t)u_meta + 16);
            if (!mb) return ST_ARENA;
            if (!push_job(W, x32 ? JK_O0_32 : JK_O0_4, make_job(in + s, in_len - s, mb, u_meta, blk))) return ST_ARENA;
            c.meta = mb;
        }
        if ((uint64_t)c_meta + s > in_len) return ST_FORMAT;
        in += c_meta + s; in_len -= c_meta + s;
        c.u_meta = u_meta;
        t1_size = rle_len;
        if (u_meta == 0) return ST_FORMAT;                           // :1600
    }

    if (in_len) {                                                    // :1577-1595
        DecJob j = make_job(in, in_len, c.t1, t1_size, blk);
        if (flags & F_CAT) {
            if (t1_size > in_len || t1_size > osz) return ST_FORMAT;
            if (!push_job(W, JK_COPY, j)) return ST_ARENA;
        } else if (flags & F_ORDER1) {
            if (in_len >= 4 * (x32 ? 32u : 4u) && (in[0] & 1)) {     // O0-compressed table (:944-955): expanded by an
                uint32_t usz, csz;                                   // order-0 job (always 4-way) ahead of the order-1 kernel
                const uint8_t* p = in + 1;
                p += var_get_u32(p, end, &usz);
                p += var_get_u32(p, end, &csz);
                if (usz > 257u * 257u * 3u + 1024u) return ST_FORMAT;
                if ((int64_t)csz >= (int64_t)(end - p) - 16) return ST_FORMAT;   // :948 (quirk kept)
                j.aux = arena_alloc(W, (uint64_t)usz + 16);
                if (!j.aux) return ST_ARENA;
                if (!push_job(W, JK_TAB, make_job(p, csz, j.aux, usz, blk))) return ST_ARENA;
            }
            // a table worth compressing (> 1000 bytes, :767) belongs to an alphabet beyond the regular 4-way variant
            uint32_t kind = x32 ? JK_O1_32 : (j.aux ? JK_O1_4M : JK_O1_4);
            if (!j.aux && in_len > 1) {                              // uncompressed table: count the alphabet to pick
                GRd ar{in + 1, end};                                 // the kernel variant (small alphabets: rows in
                uint32_t ns = 0;                                     // registers, or two to three times the occupancy)
                if (count_alphabet(ar, &ns)) {
                    if (x32 && o1_compact_bytes(ns) <= O1_SMALL_TAB) kind = JK_O1_32S;
                    if (!x32 && ns <= 8) kind = JK_O1_4R8;
                    else if (!x32 && ns <= 16) kind = JK_O1_4R16;
                    else if (!x32 && o1_compact_bytes(ns) > O1_TAB4) kind = JK_O1_4M;
                }
            }
            if (!push_job(W, kind, j)) return ST_ARENA;
        } else if (fused) {
            j.fuse = c.per; j.fin_len = osz;
            for (int k = 0; k < 16; k++) j.map[k] = c.map[k];
            if (!push_job(W, x32 ? JK_O0_32P : JK_O0_4P, j)) return ST_ARENA;
        } else {
            uint32_t kind = x32 ? JK_O0_32 : JK_O0_4;
            if (!x32) {                                              // small alphabets: tables in registers; large batch:
                GRd ar{in, end};                                     // compact tables at six times the occupancy
                uint32_t ns = 0;
                if (count_alphabet(ar, &ns)) {
                    if (ns <= 8) kind = JK_O0_4R8;
                    else if (ns <= 16) kind = JK_O0_4R16;
                    else if (W->big_batch && ns <= O0C_MAX_NS) kind = JK_O0_4C;
                }
            }
            if (!push_job(W, kind, j)) return ST_ARENA;
        }
    } else {
        t1_size = 0;
    }
    c.t1_size = t1_size;
    c.t2_size = t1_size;
    if (!(flags & (F_RLE | F_PACK))) c.final_size = t1_size;
    if (fused) c.final_size = osz;

    if (flags & F_RLE) {
        uint32_t at = atomicAdd(&W->nrle, 1u);
        if (at >= W->chain_cap) { W->overflow = 1; return ST_ARENA; }
        W->rle_list[at] = ci;
    }
    if ((flags & F_PACK) && !fused) {
        uint32_t at = atomicAdd(&W->nunpack, 1u);
        if (at >= W->chain_cap) { W->overflow = 1; return ST_ARENA; }
        W->unpack_list[at] = ci;
    }
    W->chains[ci] = c;
    if (out_len_slot && !(flags & (F_RLE | F_PACK))) *out_len_slot = t1_size;
    if (out_len_slot && fused) *out_len_slot = osz;
    return ST_OK;
}

__device__ __forceinline__ bool reserve_chains(DecWork* W, uint32_t n, uint32_t* first) {
    uint32_t at = atomicAdd(&W->nchains, n);
    if ((uint64_t)at + n > W->chain_cap) { W->overflow = 1; return false; }
    *first = at;
    return true;
}

__device__ void plan_block(const PlanArgs& A, int blk) {
    DecWork* W = A.W;
    const uint8_t* in = A.in_base + A.in_off[blk];
    uint32_t in_len = A.in_len[blk];
    uint8_t* out = A.out_base + A.out_off[blk];
    uint32_t cap = A.out_len[blk];
    int32_t st = ST_OK;

    if (A.method && A.method[blk] == 1) {                            // rANS 4x8, rANS_static.c:934-943
        if (in_len < 9 || in[0] > 1 || in_len < (in[0] ? 27u : 26u)) st = ST_FORMAT;   // :241 / :708
        else {
            uint32_t clen = ld_u32_le(in + 1), n = ld_u32_le(in + 5);
            if (clen != in_len - 9 || n >= 0x7fffffffu) st = ST_FORMAT;
            else if (n > cap) st = ST_SIZE;
            else {
                uint32_t kind = in[0] ? JK_R8_O1 : JK_R8_O0;
                if (in[0]) {
                    // count the byte values named by the order-1 tables (contexts and symbols, plus 0) to pick
                    // the kernel variant; the scan stops once the regular variant's limit is exceeded
                    GRd ar{in + 9, in + in_len};
                    uint32_t seen[8] = {1u, 0, 0, 0, 0, 0, 0, 0}, ns = 1, run_i = 0, c = ar.get(), dummy;
                    bool ok = in_len >= 27;
                    auto mark = [&](uint32_t j) { if (!((seen[j >> 5] >> (j & 31)) & 1u)) { seen[j >> 5] |= 1u << (j & 31); ns++; } };
                    while (ok) {
                        mark(c);
                        if (!parse_table_4x8(ar, [&](uint32_t j, uint32_t) { mark(j); }, &dummy, true)) { ok = false; break; }
                        if (o1_compact_bytes(ns) > O1_TAB4) break;
                        if (!run_i && c + 1 == ar.peek()) { ar.get(); c++; run_i = ar.get(); }
                        else if (run_i) { run_i--; if (++c > 255) { ok = false; break; } }
                        else c = ar.get();
                        if (!c) break;
                    }
                    if (ok && ns <= 8) kind = JK_R8_O1R8;
                    else if (ok && ns <= 16) kind = JK_R8_O1R16;
                    else if (ok && o1_compact_bytes(ns) > O1_TAB4) kind = JK_R8_O1M;
                }
                if (!in[0]) {
                    GRd ar{in + 9, in + in_len};
                    uint32_t ns = 0, sum = 0;
                    if (parse_table_4x8(ar, [&](uint32_t, uint32_t) { ns++; }, &sum, false)) {
                        if (ns <= 8) kind = JK_R8_O0R8;
                        else if (ns <= 16) kind = JK_R8_O0R16;
                        else if (W->big_batch && ns <= O0C_MAX_NS) kind = JK_R8_O0C;
                    }
                }
                if (!push_job(W, kind, make_job(in, in_len, out, n, blk))) st = ST_ARENA;
                A.out_len[blk] = n;
            }
        }
    } else if (in_len == 0) {
        st = ST_FORMAT;                                              // :1357
    } else if (in[0] & F_STRIPE) {                                   // :1360-1433
        const uint8_t* end = in + in_len;
        uint32_t ulen, pos = 1;
        pos += var_get_u32(in + pos, end, &ulen);
        if (pos >= in_len) st = ST_FORMAT;
        else {
            uint32_t N = in[pos++];
            uint64_t ctot = 0;
            uint32_t p2 = pos;
            if (N == 0 || ulen >= 0x7fffffffu) st = ST_FORMAT;
            else if (ulen > cap) st = ST_SIZE;
            for (uint32_t j = 0; st == ST_OK && j < N; j++) {
                uint32_t cl;
                p2 += var_get_u32(in + p2, end, &cl);
                ctot += cl;
                if (p2 > in_len || cl > in_len || cl < 1) st = ST_FORMAT;
            }
            if (st == ST_OK && (uint64_t)p2 + ctot > in_len) st = ST_FORMAT;
            uint32_t chain0 = 0, so = 0;
            uint8_t* parts = nullptr;
            if (st == ST_OK) {
                parts = arena_alloc(W, (uint64_t)ulen + 16);
                if (!parts || !reserve_chains(W, N, &chain0)) st = ST_ARENA;
            }
            if (st == ST_OK) {
                so = atomicAdd(&W->nstripe, 1u);
                if (so >= W->stripe_cap) { W->overflow = 1; st = ST_ARENA; }
            }
            if (st == ST_OK) {
                const uint32_t tot_len = (uint32_t)(p2 + ctot);
                uint32_t at = 0, data = p2;
                for (uint32_t j = 0; j < N; j++) {
                    uint32_t cl;
                    pos += var_get_u32(in + pos, end, &cl);
                    uint32_t ul = ulen / N + ((ulen % N) > j);
                    int32_t s2 = plan_chain(W, in + data, tot_len - data, parts + at, ul, ul, blk, nullptr, chain0 + j);
                    if (s2 != ST_OK && st == ST_OK) st = s2;
                    at += ul;
                    data += cl;
                }
                StripeOp op;
                op.parts = parts; op.out = out; op.blk = blk; op.ulen = ulen; op.N = N; op.chain0 = chain0;
                W->stripes[so] = op;
                A.out_len[blk] = ulen;
            }
        }
    } else {
        uint32_t ci;
        if (!reserve_chains(W, 1, &ci)) st = ST_ARENA;
        else st = plan_chain(W, in, in_len, out, cap, 0xffffffffu, blk, &A.out_len[blk], ci);
    }
    A.status[blk] = st;
}

__global__ void plan_kernel(PlanArgs A) {
    for (int blk = blockIdx.x * blockDim.x + threadIdx.x; blk < A.nblk; blk += gridDim.x * blockDim.x) plan_block(A, blk);
}

// ------------------------------------------------------------------------------------------
// entropy decode: shared pieces
// ------------------------------------------------------------------------------------------
template <int NWAY> struct GroupCfg {
    static constexpr int G = 32 / NWAY;                 // jobs per warp
    static constexpr int U = (NWAY == 32) ? 1 : 2;      // 16-byte chunks per lane per half ring
    static constexpr int HALF = NWAY * 16 * U;          // bytes
    static constexpr int RING = 2 * HALF;
    static constexpr uint32_t GM = (NWAY == 32) ? 0xffffffffu : ((1u << NWAY) - 1u);
};

// Lane coordinates + group-scoped warp primitives.
template <int NWAY> struct Grp {
    uint32_t g, glane, gshift, gmask;
    __device__ __forceinline__ Grp() {
        uint32_t lane = lane_id();
        g = lane / NWAY; glane = lane % NWAY; gshift = g * NWAY;
        gmask = GroupCfg<NWAY>::GM << gshift;
    }
    __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
    template <typename T> __device__ __forceinline__ T bcast(T v) const { return __shfl_sync(gmask, v, 0, NWAY); }
    __device__ __forceinline__ const uint8_t* bcast_ptr(const uint8_t* p) const {
        return reinterpret_cast<const uint8_t*>(__shfl_sync(gmask, (unsigned long long)(uintptr_t)p, 0, NWAY));
    }
    __device__ __forceinline__ bool all(bool p) const { return __all_sync(gmask, p) != 0; }
    // exclusive prefix sum over the group; *total = group sum
    __device__ __forceinline__ uint32_t exscan(uint32_t v, uint32_t* total) const {
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < NWAY; d <<= 1) {
            uint32_t y = __shfl_up_sync(gmask, x, d, NWAY);
            if (glane >= (uint32_t)d) x += y;
        }
        *total = __shfl_sync(gmask, x, NWAY - 1, NWAY);
        return x - v;
    }
};

// Streaming view of the compressed words of one job.  The ring holds the bytes
// [filled - RING, filled) counted from `abase` (the 16-byte aligned address at or below the first
// word); `pre` holds the next half ring, already on its way from global memory.
template <int NWAY> struct WordRing {
    using C = GroupCfg<NWAY>;
    const uint8_t* abase;
    const uint8_t* in_end;
    uint32_t ring;          // shared-memory address of this group's ring
    uint32_t head;          // byte offset (from abase) of the next unread byte
    uint32_t filled;
    uint4 pre[C::U];
    bool mirror;            // the first MIRROR bytes of the ring are repeated behind its end, so a read that
                            // starts inside the ring may run up to MIRROR bytes past it without wrapping
    static constexpr uint32_t MIRROR = 64;
    __device__ __forceinline__ void put(uint32_t off, uint4 v) const {       // off < RING
        sts_v4(ring + off, v);
        if (mirror && off < MIRROR) sts_v4(ring + C::RING + off, v);
    }

    __device__ __forceinline__ uint4 fetch(uint32_t byte_off) const {
        const uint8_t* gp = abase + byte_off;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gp < in_end) v = __ldcs(reinterpret_cast<const uint4*>(gp));
        return v;
    }
    // called by the lanes of one group (group-uniform `active`)
    __device__ __forceinline__ void init(const uint8_t* first, const uint8_t* end, uint32_t ring_addr,
                                         const Grp<NWAY>& G, bool active, bool with_mirror = false) {
        mirror = with_mirror;
        abase = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(first) & ~uintptr_t(15));
        in_end = end;
        ring = ring_addr;
        head = active ? (uint32_t)(first - abase) : 0u;
        filled = C::RING;
#pragma unroll
        for (int u = 0; u < C::U; u++) pre[u] = make_uint4(0, 0, 0, 0);
        if (active) {
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int u = 0; u < C::U; u++) {
                    uint32_t off = h * C::HALF + (u * NWAY + G.glane) * 16;
                    put(off, fetch(off));
                }
#pragma unroll
            for (int u = 0; u < C::U; u++) pre[u] = fetch(filled + (u * NWAY + G.glane) * 16);
        }
        G.sync();
    }
    // End of every step, called converged by the whole warp.  Groups whose head moved past the
    // older half overwrite it with `pre` and start fetching the half after that.
    __device__ __forceinline__ void advance(uint32_t glane, bool active) {
        bool need = active && head >= filled - C::HALF;
        if (__any_sync(0xffffffffu, need)) {
            if (need) {
#pragma unroll
                for (int u = 0; u < C::U; u++)
                    put((filled + (u * NWAY + glane) * 16) & (C::RING - 1), pre[u]);
                filled += C::HALF;
#pragma unroll
                for (int u = 0; u < C::U; u++) pre[u] = fetch(filled + (u * NWAY + glane) * 16);
            }
            __syncwarp();
        }
    }
    __device__ __forceinline__ uint32_t byte_at(uint32_t off) const { return lds_u8(ring + (off & (C::RING - 1))); }
    template <bool ALIGNED> __device__ __forceinline__ uint32_t word_at(uint32_t off) const {
        if (ALIGNED) return lds_u16(ring + (off & (C::RING - 1)));
        return byte_at(off) | (byte_at(off + 1) << 8);
    }
};

// The renormalisation step shared by the order-0 and order-1 loops (whole warp, converged).
// BYTE: rANS_byte.h:435-551, up to two single bytes, L = 2^23.  Else rANS_word.h:356-410, at
// most one little-endian u16, L = 2^15.
template <int NWAY, bool BYTE, bool ALIGNED>
__device__ __forceinline__ uint32_t renorm_step(uint32_t R, bool p, WordRing<NWAY>& ring, uint32_t lt, uint32_t gshift) {
    constexpr uint32_t GM = GroupCfg<NWAY>::GM;
    // p: this lane's state is below the lower bound (and the lane is active)
    if (BYTE) {
        const bool p1 = p, p2 = p && R < (1u << 15);
        uint32_t m1 = (__ballot_sync(0xffffffffu, p1) >> gshift) & GM;
        uint32_t m2 = (__ballot_sync(0xffffffffu, p2) >> gshift) & GM;
        uint32_t off = ring.head + __popc(m1 & lt) + __popc(m2 & lt);
        if (p1) {
            R = (R << 8) | ring.byte_at(off);
            if (p2) R = (R << 8) | ring.byte_at(off + 1);
        }
        ring.head += __popc(m1) + __popc(m2);
    } else {
        uint32_t m = (__ballot_sync(0xffffffffu, p) >> gshift) & GM;
        if (p) R = (R << 16) | ring.template word_at<ALIGNED>(ring.head + 2 * __popc(m & lt));
        ring.head += 2 * __popc(m);
    }
    return R;
}

// Copy `nchunks` 16-byte chunks starting at the aligned address `gp` into shared memory at `dst`
// (chunks at or beyond `end` become zeros).  Called by the NWAY lanes of a group.
template <int NWAY>
__device__ __forceinline__ void stage_chunks(uint32_t dst, const uint8_t* gp, const uint8_t* end, int nchunks,
                                             uint32_t glane) {
    for (int c = glane; c < nchunks; c += NWAY) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gp + 16 * c < end) v = __ldg(reinterpret_cast<const uint4*>(gp + 16 * c));
        sts_v4(dst + 16 * c, v);
    }
}

// bounded reader over shared-memory bytes
struct SRd {
    uint32_t a, end;
    __device__ __forceinline__ bool more() const { return a < end; }
    __device__ __forceinline__ uint32_t peek() const { return a < end ? lds_u8(a) : 0u; }
    __device__ __forceinline__ uint32_t get() { uint32_t v = peek(); a++; return v; }
    __device__ __forceinline__ uint32_t varint() {           // varint.h:131-160
        uint32_t x = 0;
        if (a >= end) return 0;
        for (;;) {
            uint32_t c = get();
            x = (x << 7) | (c & 0x7f);
            if (!(c & 0x80) || a >= end) break;
        }
        return x;
    }
};

// decode_alphabet (rANS_static4x16pr.c:208-255): marks present symbols by storing `mark` into
// the byte table at shared address `tab`.  Returns false when the bytes run out.
template <typename RD>
__device__ bool read_alphabet(RD& r, uint32_t tab, uint32_t mark) {
    if (!r.more()) return false;
    uint32_t run = 0, j = r.get();
    do {
        sts_u8(tab + j, mark);
        if (!r.more()) return false;
        if (!run && j + 1 == r.peek()) {
            r.get();
            if (!r.more()) return false;
            j++;
            run = r.get();
        } else if (run) {
            run--;
            if (++j > 255) return false;
        } else {
            j = r.get();
        }
    } while (j && r.more());
    return true;
}

// decode_freq, rANS_static4x16pr.c:271-289, read from shared memory by ONE lane.  F is a zeroed
// 256-entry u32 array in shared memory, `pres` a zeroed 256-byte table.  Returns the table length
// in bytes (0 = malformed) and the frequency sum.
__device__ uint32_t parse_o0_table_4x16(uint32_t src, uint32_t lim, uint32_t F, uint32_t pres, uint32_t* sum) {
    SRd r{src, src + lim};
    if (!read_alphabet(r, pres, 1u)) return 0;
    uint32_t tot = 0;
    for (uint32_t s = 0; s < 256; s++) {
        if (!lds_u8(pres + s)) continue;
        uint32_t f = r.varint();
        sts_u32(F + 4 * s, f);
        tot += f;
    }
    *sum = tot;
    return r.a - src;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}

// ---- renormalisation bytes of a 4-lane group as a register window (4-way / 4x8 kernels) ----------------------
// The 8 bytes at the group's read position (a step consumes at most 4 words, or 4 x 2 bytes for 4x8) are fetched at
// the START of a step, and after the ballot each lane picks its word with two byte permutes -- the selector comes from
// a permute table indexed by the ballot bits of the lower lanes -- so neither a popc nor a shared-memory load sits
// between the ballot and the next state.  Needs the ring's mirror (its first 64 bytes repeated behind its end).
struct Win { uint32_t lo, hi; };

// prmt.b32 without __byte_perm's selector masking: every selector nibble used here is <= 7 (bit 3 = sign replication)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// The 8 bytes at byte offset `head` of a 4-way group's ring (RING = 256 bytes + mirror of its first 64).
__device__ __forceinline__ Win win_load(uint32_t ring, uint32_t head) {
    const uint32_t a = ring + (head & 252u);
    const uint32_t w0 = lds_u32(a), w1 = lds_u32(a + 4), w2 = lds_u32(a + 8);
    const uint32_t sh = (head & 3u) * 8u;
    Win w;
    w.lo = __funnelshift_r(w0, w1, sh);
    w.hi = __funnelshift_r(w1, w2, sh);
    return w;
}

// Renormalise every lane of the warp (rANS_word.h:356-410 / rANS_byte.h:435-551) from its group's window.
// p: this lane's state `x` is below the lower bound (and the lane is active).  Advances `head`.
template <bool BYTE>
__device__ __forceinline__ uint32_t win_renorm(uint32_t x, bool p, const Win w, uint32_t& head, uint32_t lt4, uint32_t gshift) {
    if (BYTE) {
        const bool p2 = p && x < (1u << 15);
        const uint32_t b1 = __ballot_sync(0xffffffffu, p) >> gshift, b2 = __ballot_sync(0xffffffffu, p2) >> gshift;
        // bytes taken by the lower lanes = popc(b1 & lt4) + popc(b2 & lt4), each from a permute table (v in 0..7)
        const uint32_t o = prmt(0x02010100u, 0x03020201u, b1 & lt4) + prmt(0x02010100u, 0x03020201u, b2 & lt4);   // 0 .. 6
        const uint32_t ww = prmt(w.lo, w.hi, o * 0x11u + 0x10u);           // bytes o, o + 1
        if (p2) x = prmt(ww, x, 0x5401);                                   // x << 16 | first << 8 | second
        else if (p) x = prmt(ww, x, 0x6540);                               // x << 8 | first
        head += __popc(b1 & 15u) + __popc(b2 & 15u);
    } else {
        const uint32_t b = __ballot_sync(0xffffffffu, p) >> gshift;
        // word index k = popc(b & lt4) -> selector bytes (2k, 2k+1), by table: v = b & lt4 in 0..7
        const uint32_t selk = prmt(0x54323210u, 0x76545432u, b & lt4);
        const uint32_t ww = prmt(w.lo, w.hi, selk);
        if (p) x = prmt(ww, x, 0x5410);                                    // x << 16 | word
        head += 2u * __popc(b & 15u);
    }
    return x;
}


// ------------------------------------------------------------------------------------------
// order-0
// ------------------------------------------------------------------------------------------
// Shared memory of dec_o0_kernel: G symbol LUTs (4096 B each; during set-up a LUT doubles as
// header staging [0,1040), presence bytes [1280,1536) and frequency scratch [2048,3072)),
// then G fc tables (256 x {F, L + C}, L = the renormalisation bound), then G word rings.  8 KB per
// X_32 warp with the per-CTA reserve: 28 resident warps per SM.
template <int NWAY, bool FUSE = false> struct O0Smem {
    static constexpr int G = GroupCfg<NWAY>::G;
    static constexpr int LUT = 0, FC = G * 4096, RINGO = G * 6144;
    static constexpr int RINGSZ = GroupCfg<NWAY>::RING + 64;            // ring + mirror of its first 64 bytes
    static constexpr int EXPO = G * (6144 + RINGSZ);                    // FUSE: 16 x 4-byte un-PACK nibble expansions per group
    static constexpr int TOTAL = G * (6144 + RINGSZ + (FUSE ? 64 : 0));
};
constexpr int HDR_STAGE = 1040;     // bytes of stream head staged for the table parser

// Turn 256 frequencies (shared u32 array F) into the decode tables of one group:
//   fc[s]  = { F[s], L + C[s] }    so that   X = F*(x>>12) + m,  x' = X + L - fc.y  and the
//                                  renormalisation test x' < L is X < fc.y, one level earlier
//   lut[m] = s                     for C[s] <= m < C[s]+F[s]
// Group-synchronous.  Returns false unless the frequencies sum to `want` (or `want_alt`).
template <int NWAY>
__device__ bool build_o0_tables(const Grp<NWAY>& G, uint32_t F, uint32_t fc, uint32_t lut, uint32_t want,
                                uint32_t want_alt, uint32_t L) {
    constexpr int K = 256 / NWAY;
    uint32_t mine = 0;
    bool bad = false;
    for (int k = 0; k < K; k++) {
        uint32_t f = lds_u32(F + 4 * (G.glane * K + k));
        if (f > 4096) bad = true;                                    // :540
        mine += f;
    }
    uint32_t total;
    uint32_t c = G.exscan(bad ? 8192u : mine, &total);
    if (total != want && total != want_alt) return false;            // :551 (group-uniform)
    for (int k = 0; k < K; k++) {
        uint32_t s = G.glane * K + k;
        uint32_t f = lds_u32(F + 4 * s);
        sts_v2(fc + 8 * s, make_uint2(f, L + c));
        c += f;
    }
    G.sync();
    for (uint32_t s = 0; s < 256; s++) {                             // :538-549, the group fills one symbol at a time
        uint2 e = lds_v2(fc + 8 * s);
        uint32_t f = e.x, cs = e.y - L;
        for (uint32_t k = G.glane; k < f; k += NWAY) sts_u8(lut + cs + k, s);
    }
    G.sync();
    return true;
}

// Per-group set-up: stage the stream head, parse the table, load the states, build the tables.
// Returns group-uniform success; *R = this lane's state, *first = offset of the first word.
template <int NWAY, bool BYTE>
__device__ bool o0_setup(const Grp<NWAY>& G, const DecJob& job, uint32_t lut, uint32_t fc, uint32_t* R,
                         uint32_t* first_word) {
    const uint8_t* in_end = job.in + job.in_len;
    const uint8_t* a0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(job.in) & ~uintptr_t(15));
    const uint32_t skew = (uint32_t)(job.in - a0);
    const uint32_t pres = lut + 1280, Ftmp = lut + 2048;
    stage_chunks<NWAY>(lut, a0, in_end, HDR_STAGE / 16, G.glane);
    for (uint32_t k = G.glane; k < 256; k += NWAY) { sts_u32(Ftmp + 4 * k, 0u); sts_u8(pres + k, 0u); }
    G.sync();
    const uint32_t hdr0 = BYTE ? 9u : 0u;                // 4x8: [order][clen][ulen] precede the table
    uint32_t tab = 0, sum = 0;
    if (G.glane == 0 && job.in_len >= hdr0 + 4 * NWAY) {                         // :503
        uint32_t lim = min(job.in_len, (uint32_t)HDR_STAGE - 16u);
        if (BYTE) {
            SRd r{lut + skew + hdr0, lut + skew + lim};
            if (parse_table_4x8(r, Ftmp, &sum, false)) tab = r.a - (lut + skew + hdr0);
        } else {
            tab = parse_o0_table_4x16(lut + skew, lim, Ftmp, pres, &sum);
        }
    }
    tab = G.bcast(tab);
    sum = G.bcast(sum);
    const uint32_t first = hdr0 + tab;                   // offset of the NWAY initial states
    if (tab == 0 || first + 4 * NWAY > job.in_len || first + 4 * NWAY > HDR_STAGE - 16) return false;
    // normalise_freq_shift (…4x16pr.c:168-179): stored sums are powers of two <= 4096
    if (!BYTE && sum != 0 && sum < 4096) {
        uint32_t sh = 0;
        while ((sum << sh) < 4096) sh++;
        for (uint32_t k = G.glane; k < 256; k += NWAY) sts_u32(Ftmp + 4 * k, lds_u32(Ftmp + 4 * k) << sh);
    }
    uint32_t p = lut + skew + first + 4 * G.glane;
    uint32_t r0 = lds_u8(p) | (lds_u8(p + 1) << 8) | (lds_u8(p + 2) << 16) | (lds_u8(p + 3) << 24);
    G.sync();
    if (!G.all(r0 >= (BYTE ? (1u << 23) : (1u << 15)))) return false;            // :557-561
    *R = r0;
    *first_word = first + 4 * NWAY;
    return build_o0_tables<NWAY>(G, Ftmp, fc, lut, 4096u, BYTE ? 4095u : 4096u,   // 4x8 tables may sum to 4095 (:305)
                                 BYTE ? (1u << 23) : (1u << 15));
}

// Four steps of a 4-lane group produce 16 contiguous output bytes: lane z holds the bytes of
// positions 4k + z (k = 0..3, byte k of `w`).  A 4 x 4 byte transpose inside the group (two
// shuffle/permute rounds) gives lane z positions 4z .. 4z + 3, so the group writes one 32-bit word
// per lane instead of four scattered bytes per lane: a quarter of the (partial) sector writes.
__device__ __forceinline__ uint32_t transpose4x4(uint32_t w, uint32_t glane) {
    uint32_t t = __shfl_xor_sync(0xffffffffu, w, 1);
    w = __byte_perm(w, t, (glane & 1) ? 0x3715 : 0x6240);
    t = __shfl_xor_sync(0xffffffffu, w, 2);
    return __byte_perm(w, t, (glane & 2) ? 0x3276 : 0x5410);
}

// One decode step (rANS_static4x16pr.c:576-597 / rANS_static.c:318-344) for every lane.
template <int NWAY, bool BYTE, bool ALIGNED, bool ALLACT>
__device__ __forceinline__ uint32_t o0_step(uint32_t R, bool act, WordRing<NWAY>& ring, uint32_t lut, uint32_t fc,
                                            uint8_t* op, uint32_t lt, uint32_t gshift, uint32_t* sym_out = nullptr) {
    constexpr uint32_t L = BYTE ? (1u << 23) : (1u << 15);
    Win win;
    if (NWAY == 4) win = win_load(ring.ring, ring.head);
    const uint32_t m = R & 0xfffu;
    const uint32_t s = lds_u8(lut + m);
    const uint2 e = lds_v2(fc + s * 8);
    const uint32_t X = e.x * (R >> 12) + m;
    const bool p = (ALLACT || act) && X < e.y;               // x' < L
    if (ALLACT || act) { R = X + L - e.y; if (sym_out) *sym_out = s; else *op = (uint8_t)s; }
    if (NWAY == 4) return win_renorm<BYTE>(R, p, win, ring.head, lt, gshift);
    return renorm_step<NWAY, BYTE, ALIGNED>(R, p, ring, lt, gshift);
}

// The X_32 hot step in PTX, so that the dependent chain is exactly
//   LUT -> (F, L+C) -> mad -> setp -> vote -> and/popc -> mad -> word -> and -> LUT ...
// Everything else (state merge, head bookkeeping, the output store) hangs off it.  `m` = R & 0xfff
// is carried between steps (after a renormalisation it comes straight from the fetched word);
// `hp` = shared address of the next unread word, wrapped inside the ring (1 KB, 1 KB-aligned, with a
// 64-byte mirror so that hp + 2k needs no wrap); `head` = the same position, unwrapped, for refills.
template <int OFF>
__device__ __forceinline__ void o0_step_x32(uint32_t& R, uint32_t& m, uint32_t& hp, uint32_t& head, uint32_t lut,
                                            uint32_t fc, uint32_t ringbase, uint32_t lt, uint8_t* op) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 s, sa, F, Y, q, X, b, k, a, w, Rs, kk;\n\t"
        "add.u32 sa, %1, %4;\n\t"
        "ld.shared.u8 s, [sa];\n\t"
        "shr.u32 q, %0, 12;\n\t"
        "shl.b32 sa, s, 3;\n\t"
        "add.u32 sa, sa, %5;\n\t"
        "ld.shared.v2.u32 {F, Y}, [sa];\n\t"
        "st.global.u8 [%8+%9], s;\n\t"
        "mad.lo.u32 X, F, q, %1;\n\t"
        "setp.lt.u32 p, X, Y;\n\t"
        "sub.u32 %0, X, Y;\n\t"
        "add.u32 %0, %0, 32768;\n\t"
        "vote.sync.ballot.b32 b, p, 0xffffffff;\n\t"
        "and.b32 %1, %0, 4095;\n\t"
        "shl.b32 Rs, %0, 16;\n\t"
        "and.b32 k, b, %7;\n\t"
        "popc.b32 k, k;\n\t"
        "mad.lo.u32 a, k, 2, %2;\n\t"
        "@p ld.shared.u16 w, [a];\n\t"
        "@p and.b32 %1, w, 4095;\n\t"
        "@p or.b32 %0, Rs, w;\n\t"
        "popc.b32 kk, b;\n\t"
        "shl.b32 kk, kk, 1;\n\t"
        "add.u32 %3, %3, kk;\n\t"
        "add.u32 %2, %2, kk;\n\t"
        "and.b32 %2, %2, 1023;\n\t"
        "or.b32 %2, %2, %6;\n\t"
        "}"
        : "+r"(R), "+r"(m), "+r"(hp), "+r"(head)
        : "r"(lut), "r"(fc), "r"(ringbase), "r"(lt), "l"(op), "n"(OFF)
        : "memory");
}

// `minit` (warp-uniform) = steps for which every lane of the warp is active: they run four to a
// ring check (four steps consume at most half of a half ring).
template <int NWAY, bool BYTE, bool ALIGNED>
__device__ __forceinline__ void o0_loop(uint32_t R, WordRing<NWAY>& ring, uint32_t lut, uint32_t fc, uint8_t* out,
                                        uint32_t iters, uint32_t rem, uint32_t minit, uint32_t maxit, const Grp<NWAY>& G) {
    const uint32_t lt = (NWAY == 32) ? lanemask_lt() : ((1u << G.glane) - 1u);
    uint8_t* op = out + G.glane;
    uint32_t i = 0;
    if (NWAY == 32 && !BYTE && ALIGNED && (ring.ring & 1023u) == 0) {
        uint32_t m = R & 0xfffu, hp = ring.ring + (ring.head & 1023u);
        for (; i + 4 <= minit; i += 4) {
            o0_step_x32<0>(R, m, hp, ring.head, lut, fc, ring.ring, lt, op);
            o0_step_x32<32>(R, m, hp, ring.head, lut, fc, ring.ring, lt, op);
            o0_step_x32<64>(R, m, hp, ring.head, lut, fc, ring.ring, lt, op);
            o0_step_x32<96>(R, m, hp, ring.head, lut, fc, ring.ring, lt, op);
            op += 128;
            ring.advance(G.glane, true);
        }
    } else if (NWAY == 4 && __all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(out) & 3) == 0)) {
        auto four = [&]() -> uint32_t {
            uint32_t w = 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                uint32_t sy;
                R = o0_step<NWAY, BYTE, ALIGNED, true>(R, true, ring, lut, fc, nullptr, lt, G.gshift, &sy);
                w |= sy << (8 * u);
            }
            return w;
        };
        if (minit >= 12) {                                   // rounds of eight steps, as in dec_o0r_kernel
            uint32_t pend = four();
            ring.advance(G.glane, true);
            for (i = 4; i + 8 <= minit; i += 8) {
                *reinterpret_cast<uint32_t*>(op + 3 * G.glane) = transpose4x4(pend, G.glane);
                const uint32_t w0 = four();
                *reinterpret_cast<uint32_t*>(op + 16 + 3 * G.glane) = transpose4x4(w0, G.glane);
                pend = four();
                op += 32;
                ring.advance(G.glane, true);
            }
            *reinterpret_cast<uint32_t*>(op + 3 * G.glane) = transpose4x4(pend, G.glane);
            op += 16;
        }
        for (; i + 4 <= minit; i += 4) {                     // word stores after a 4 x 4 transpose
            *reinterpret_cast<uint32_t*>(op + 3 * G.glane) = transpose4x4(four(), G.glane);
            op += 4 * NWAY;
            ring.advance(G.glane, true);
        }
    } else {
        for (; i + 4 <= minit; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++)
                R = o0_step<NWAY, BYTE, ALIGNED, true>(R, true, ring, lut, fc, op + u * NWAY, lt, G.gshift);
            op += 4 * NWAY;
            ring.advance(G.glane, true);
        }
    }
    for (; i < maxit; i++) {
        const bool act = i < iters;
        R = o0_step<NWAY, BYTE, ALIGNED, false>(R, act, ring, lut, fc, op, lt, G.gshift);
        if (act) op += NWAY;
        ring.advance(G.glane, act);
    }
    // the last n % NWAY symbols: peek only (rANS_static.c:346-355; for Nx16 nothing follows them)
    if (G.glane < rem) *op = (uint8_t)lds_u8(lut + (R & 0xfffu));
}

// un-PACK as the decoder's sink (pack.c:211-348): decoded byte number idx expands to `per` symbols and lands at
// out[idx * per ..); the last byte of a stream may carry fewer symbols than `per`.  The expansion goes through the
// group's 16-entry nibble table (built at set-up from the container's symbol map): a nibble holds per / 2 codes, so
// T[lo] and T[hi] are the byte's first and second half (4 + 4, 2 + 2 or 1 + 1 symbols, LSB-first).
__device__ __forceinline__ void emit_packed(uint8_t* out, uint32_t idx, uint32_t s, uint32_t expt, uint32_t per, uint32_t fin_len, bool aligned) {
    const uint64_t at = (uint64_t)idx * per;
    if (at >= fin_len) return;
    const uint32_t lo = lds_u32(expt + 4 * (s & 15u)), hi = lds_u32(expt + 4 * (s >> 4));
    uint2 e;
    if (per == 8) e = make_uint2(lo, hi);
    else e = make_uint2(lo | (hi << (4 * per)), 0u);         // per 4: 16-bit halves; per 2: 8-bit halves
    uint8_t* p = out + at;
    if (aligned && at + per <= fin_len) {
        if (per == 4) *reinterpret_cast<uint32_t*>(p) = e.x;
        else if (per == 8) *reinterpret_cast<uint2*>(p) = e;
        else *reinterpret_cast<uint16_t*>(p) = (uint16_t)e.x;
    } else {
        const uint32_t n = (uint32_t)min((uint64_t)per, fin_len - at);
        for (uint32_t k = 0; k < n; k++) p[k] = (uint8_t)((k < 4 ? e.x >> (8 * k) : e.y >> (8 * (k - 4))));
    }
}

template <int NWAY>
__device__ __forceinline__ void o0_loop_packed(uint32_t R, WordRing<NWAY>& ring, uint32_t lut, uint32_t fc, uint32_t expt, const DecJob& job,
                                               uint32_t iters, uint32_t rem, uint32_t minit, uint32_t maxit, const Grp<NWAY>& G) {
    const uint32_t lt = (NWAY == 32) ? lanemask_lt() : ((1u << G.glane) - 1u);
    const uint32_t per = job.fuse ? job.fuse : 1u;
    const bool aligned = (reinterpret_cast<uintptr_t>(job.out) & (per - 1)) == 0;
    uint32_t idx = G.glane, i = 0;
    for (; i + 4 <= minit; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint32_t sy;
            R = o0_step<NWAY, false, false, true>(R, true, ring, lut, fc, nullptr, lt, G.gshift, &sy);
            emit_packed(job.out, idx, sy, expt, per, job.fin_len, aligned);
            idx += NWAY;
        }
        ring.advance(G.glane, true);
    }
    for (; i < maxit; i++) {
        const bool act = i < iters;
        uint32_t sy = 0;
        R = o0_step<NWAY, false, false, false>(R, act, ring, lut, fc, nullptr, lt, G.gshift, &sy);
        if (act) { emit_packed(job.out, idx, sy, expt, per, job.fin_len, aligned); idx += NWAY; }
        ring.advance(G.glane, act);
    }
    if (G.glane < rem) emit_packed(job.out, idx, lds_u8(lut + (R & 0xfffu)), expt, per, job.fin_len, aligned);
}

template <int NWAY, bool BYTE, bool FUSE = false>
__global__ void __launch_bounds__(32) dec_o0_kernel(DecWork* W, int32_t* status, uint32_t kind) {
    using C = GroupCfg<NWAY>;
    using S = O0Smem<NWAY, FUSE>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const Grp<NWAY> G;
    uint32_t base = smem_addr(smem_raw);
    asm volatile("" : "+r"(base));                      // keep the window base in a register (no re-derivation per step)
    const uint32_t lut = base + S::LUT + G.g * 4096, fc = base + S::FC + G.g * 2048;
    const uint32_t ringa = base + S::RINGO + G.g * S::RINGSZ;
    const uint32_t njobs = W->njobs[kind];
    const DecJob* jobs = W->jobs[kind];

    // Jobs are claimed from an atomic cursor.  The host shapes the launch so that every SM holds the
    // same number of CTAs (shaped_launch below): a batch that fits in one wave is then spread evenly.
    for (;;) {
        uint32_t j0 = 0;
        if (lane_id() == 0) j0 = atomicAdd(&W->next[kind], (uint32_t)C::G);
        j0 = __shfl_sync(0xffffffffu, j0, 0);
        if (j0 >= njobs) break;
        // (the groups of a last, partly filled warp repeat the list's last job -- identical bytes to identical
        //  addresses -- instead of idling: an idle group would keep the whole warp out of the all-lanes-active loops,
        //  i.e. at half speed, and the kernel ends with its slowest warp)
        const uint32_t ji = min(j0 + G.g, njobs - 1u);
        const bool active = true;
        DecJob job = make_job(nullptr, 0, nullptr, 0, 0);
        uint32_t R = 0, first_word = 0;
        bool ok = false;
        WordRing<NWAY> ring;
        if (active) {
            job = jobs[ji];
            ok = o0_setup<NWAY, BYTE>(G, job, lut, fc, &R, &first_word);
            if (!ok && G.glane == 0) set_status(status, job.blk, ST_FORMAT);
        }
        ring.init(job.in + first_word, job.in + job.in_len, ringa, G, ok, true);
        __syncwarp();
        const uint32_t iters = ok ? job.out_len / NWAY : 0, rem = ok ? job.out_len % NWAY : 0;
        const uint32_t maxit = __reduce_max_sync(0xffffffffu, iters), minit = __reduce_min_sync(0xffffffffu, iters);
        if (FUSE) {
            // the group's un-PACK nibble expansions: 4 bits -> per / 2 symbols (pack.c:211-348, LSB-first)
            const uint32_t expt = base + S::EXPO + G.g * 64;
            if (ok) {
                const uint32_t per = job.fuse, bits = 8 / (per ? per : 1u), cmask = (1u << bits) - 1u, half = per / 2;
                for (uint32_t v = G.glane; v < 16; v += NWAY) {
                    uint32_t t = 0;
                    for (uint32_t k = 0; k < half; k++) t |= (uint32_t)job.map[(v >> (k * bits)) & cmask & 15u] << (8 * k);
                    sts_u32(expt + 4 * v, t);
                }
            }
            __syncwarp();
            o0_loop_packed<NWAY>(R, ring, lut, fc, expt, job, iters, rem, minit, maxit, G);
        } else {
            const bool aligned = !BYTE && __all_sync(0xffffffffu, (ring.head & 1u) == 0);
            if (aligned) o0_loop<NWAY, BYTE, true >(R, ring, lut, fc, job.out, iters, rem, minit, maxit, G);
            else         o0_loop<NWAY, BYTE, false>(R, ring, lut, fc, job.out, iters, rem, minit, maxit, G);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// order-0 on compact tables (4-way and 4x8 streams of large batches)
// ------------------------------------------------------------------------------------------
// The 4 KB symbol LUT limits the LUT kernels to 40 resident 4-way streams per SM, and a 4-way
// stream is four lanes of strictly serial work, so a batch beyond one wave (5920 streams) gains
// nothing.  Here a stream's table is the compact form of section "order-1" with a single context --
// a 64-bucket coarse index and one packed entry (C+F-1)<<20 | sym<<12 | (F-1) per symbol -- 592
// bytes per stream with its word ring: 256 resident streams per SM.  Alphabets of up to 64 symbols.
struct O0CSmem {
    static constexpr int COARSE = 0, ENT = 64, RINGO = 64 + 4 * (O0C_MAX_NS + 4), STRIDE = RINGO + 256 + 64;   // ring + mirror
    static constexpr int TOTAL = 8 * STRIDE;
};

template <bool BYTE>
__device__ bool o0c_setup(const Grp<4>& G, const DecJob& job, uint32_t gsm, uint32_t* R, uint32_t* first_word) {
    constexpr uint32_t L = BYTE ? (1u << 23) : (1u << 15);
    const uint32_t coarse = gsm + O0CSmem::COARSE, ent = gsm + O0CSmem::ENT, syms = gsm + O0CSmem::RINGO;   // ring: free for now
    const uint8_t* in_end = job.in + job.in_len;
    const uint32_t hdr0 = BYTE ? 9u : 0u;
    uint32_t tab = 0;
    if (G.glane == 0 && job.in_len >= hdr0 + 16) {
        GRd rd{job.in + hdr0, in_end};
        uint32_t ns = 0, sum = 0;
        bool ok;
        if (BYTE) {
            ok = parse_table_4x8(rd, [&](uint32_t j, uint32_t f) {
                if (ns < O0C_MAX_NS) { sts_u8(syms + ns, j); sts_u32(ent + 4 * ns, f); }
                ns++; }, &sum, false);
        } else {
            uint32_t cnt = 0;
            ok = list_alphabet(rd, &ns, [&](uint32_t j) { if (cnt < O0C_MAX_NS) sts_u8(syms + cnt, j); cnt++; });
            if (ok && ns <= O0C_MAX_NS)
                for (uint32_t i = 0; i < ns; i++) { const uint32_t f = rd.varint(); sts_u32(ent + 4 * i, f); sum += f; }
        }
        ok = ok && ns >= 1 && ns <= O0C_MAX_NS;
        uint32_t sh = 0;
        if (ok && !BYTE && sum != 0 && sum < 4096) while ((sum << sh) < 4096) sh++;      // normalise_freq_shift
        uint32_t c = 0, idx = 0;
        for (uint32_t k = 0; k < 16; k++) sts_u32(coarse + 4 * k, 0u);
        for (uint32_t i = 0; ok && i < ns; i++) {
            const uint32_t f = lds_u32(ent + 4 * i) << sh;
            if (!f) continue;
            if (f > 4096 - c) { ok = false; break; }
            const uint32_t last = c + f - 1;
            sts_u32(ent + 4 * idx, (last << 20) | (lds_u8(syms + i) << 12) | (f - 1));
            for (uint32_t q = (c + 63) >> 6; q <= (last >> 6); q++) sts_u8(coarse + q, idx);
            idx++;
            c += f;
        }
        if (ok && c != 4096 && !(BYTE && c == 4095)) ok = false;                         // :551 / rANS_static.c:305
        for (uint32_t k = 0; k < 3; k++) sts_u32(ent + 4 * (idx + k), O1_SENTINEL);
        if (ok) tab = (uint32_t)(rd.p - job.in);
    }
    tab = G.bcast(tab);
    if (tab == 0 || tab + 16 > job.in_len) return false;
    const uint32_t r0 = ld_u32_le(job.in + tab + 4 * G.glane);
    if (!G.all(r0 >= L)) return false;
    *R = r0;
    *first_word = tab + 16;
    G.sync();
    return true;
}

template <bool BYTE, bool ALIGNED, bool ALLACT>
__device__ __forceinline__ uint32_t o0c_step(uint32_t R, bool act, WordRing<4>& ring, uint32_t coarse, uint32_t ent,
                                             uint8_t* op, uint32_t lt, uint32_t gshift, uint32_t* sym_out = nullptr) {
    constexpr uint32_t L = BYTE ? (1u << 23) : (1u << 15);
    const Win win = win_load(ring.ring, ring.head);
    const uint32_t m = R & 0xfffu, mk = m << 20, q = R >> 12, qm = q + m;
    uint32_t ea = ent + 4 * lds_u8(coarse + (m >> 6));
    const uint32_t e0 = lds_u32(ea), e1 = lds_u32(ea + 4), e2 = lds_u32(ea + 8);
    uint32_t e = (mk > e1) ? e2 : ((mk > e0) ? e1 : e0);
    if (mk > e) {                                            // >= 4 symbols share the bucket (sentinel-bounded scan)
        ea += 12;
        do { e = lds_u32(ea); ea += 4; } while (mk > e);
    }
    const uint32_t Rn = (e & 0xfffu) * (q + 1u) + (qm - (e >> 20));
    const bool p = (ALLACT || act) && Rn < L;
    if (ALLACT || act) { R = Rn; if (sym_out) *sym_out = (e >> 12) & 0xffu; else *op = (uint8_t)(e >> 12); }
    return win_renorm<BYTE>(R, p, win, ring.head, lt, gshift);
}

template <bool BYTE, bool ALIGNED>
__device__ __forceinline__ void o0c_loop(uint32_t R, WordRing<4>& ring, uint32_t coarse, uint32_t ent, uint8_t* out,
                                         uint32_t iters, uint32_t rem, uint32_t minit, uint32_t maxit, const Grp<4>& G) {
    const uint32_t lt = (1u << G.glane) - 1u;
    uint8_t* op = out + G.glane;
    uint32_t i = 0;
    if (__all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(out) & 3) == 0)) {
        auto four = [&]() -> uint32_t {
            uint32_t w = 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                uint32_t sy;
                R = o0c_step<BYTE, ALIGNED, true>(R, true, ring, coarse, ent, nullptr, lt, G.gshift, &sy);
                w |= sy << (8 * u);
            }
            return w;
        };
        if (minit >= 12) {                                   // rounds of eight steps, as in dec_o0r_kernel
            uint32_t pend = four();
            ring.advance(G.glane, true);
            for (i = 4; i + 8 <= minit; i += 8) {
                *reinterpret_cast<uint32_t*>(op + 3 * G.glane) = transpose4x4(pend, G.glane);
                const uint32_t w0 = four();
                *reinterpret_cast<uint32_t*>(op + 16 + 3 * G.glane) = transpose4x4(w0, G.glane);
                pend = four();
                op += 32;
                ring.advance(G.glane, true);
            }
            *reinterpret_cast<uint32_t*>(op + 3 * G.glane) = transpose4x4(pend, G.glane);
            op += 16;
        }
        for (; i + 4 <= minit; i += 4) {                     // word stores after a 4 x 4 transpose
            *reinterpret_cast<uint32_t*>(op + 3 * G.glane) = transpose4x4(four(), G.glane);
            op += 16;
            ring.advance(G.glane, true);
        }
    }
    for (; i + 4 <= minit; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++)
            R = o0c_step<BYTE, ALIGNED, true>(R, true, ring, coarse, ent, op + u * 4, lt, G.gshift);
        op += 16;
        ring.advance(G.glane, true);
    }
    for (; i < maxit; i++) {
        const bool act = i < iters;
        R = o0c_step<BYTE, ALIGNED, false>(R, act, ring, coarse, ent, op, lt, G.gshift);
        if (act) op += 4;
        ring.advance(G.glane, act);
    }
    if (G.glane < rem) {                                     // rANS_static.c:346-355: peek only
        const uint32_t m = R & 0xfffu, mk = m << 20;
        uint32_t ea = ent + 4 * lds_u8(coarse + (m >> 6)), e;
        do { e = lds_u32(ea); ea += 4; } while (mk > e);
        *op = (uint8_t)(e >> 12);
    }
}

template <bool BYTE>
__global__ void __launch_bounds__(32, 32) dec_o0c_kernel(DecWork* W, int32_t* status, uint32_t kind) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const Grp<4> G;
    uint32_t base = smem_addr(smem_raw) + G.g * O0CSmem::STRIDE;
    asm volatile("" : "+r"(base));
    const uint32_t njobs = W->njobs[kind];
    const DecJob* jobs = W->jobs[kind];
    for (;;) {
        uint32_t j0 = 0;
        if (lane_id() == 0) j0 = atomicAdd(&W->next[kind], 8u);
        j0 = __shfl_sync(0xffffffffu, j0, 0);
        if (j0 >= njobs) break;
        // (the groups of a last, partly filled warp repeat the list's last job -- identical bytes to identical
        //  addresses -- instead of idling: an idle group would keep the whole warp out of the all-lanes-active loops,
        //  i.e. at half speed, and the kernel ends with its slowest warp)
        const uint32_t ji = min(j0 + G.g, njobs - 1u);
        const bool active = true;
        DecJob job = make_job(nullptr, 0, nullptr, 0, 0);
        uint32_t R = 0, first_word = 0;
        bool ok = false;
        if (active) {
            job = jobs[ji];
            ok = o0c_setup<BYTE>(G, job, base, &R, &first_word);
            if (!ok && G.glane == 0) set_status(status, job.blk, ST_FORMAT);
        }
        WordRing<4> ring;
        ring.init(job.in + first_word, job.in + job.in_len, base + O0CSmem::RINGO, G, ok, true);
        __syncwarp();
        const uint32_t iters = ok ? job.out_len / 4 : 0, rem = ok ? job.out_len % 4 : 0;
        const uint32_t maxit = __reduce_max_sync(0xffffffffu, iters), minit = __reduce_min_sync(0xffffffffu, iters);
        const bool aligned = !BYTE && __all_sync(0xffffffffu, (ring.head & 1u) == 0);
        const uint32_t coarse = base + O0CSmem::COARSE, ent = base + O0CSmem::ENT;
        if (aligned) o0c_loop<BYTE, true >(R, ring, coarse, ent, job.out, iters, rem, minit, maxit, G);
        else         o0c_loop<BYTE, false>(R, ring, coarse, ent, job.out, iters, rem, minit, maxit, G);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// order-1
// ------------------------------------------------------------------------------------------
// Symbols are ranked over the stream's alphabet (rank r = index of a symbol among the present
// ones).  Two table forms:
//
//  * compact (shared memory; Nx16 streams whose alphabet fits).  Per context a row of packed
//    entries, one per symbol of non-zero frequency, in cumulative order,
//        entry = (C + F - 1) << 20 | rank << 12 | (F - 1)
//    preceded by a 64-bucket coarse index: coarse[m >> (shift - 6)] = index of the entry holding
//    the bucket's first slot.  A lookup reads the coarse byte and three consecutive entries and
//    steps forward while m lies beyond an entry's last slot (a fourth or later entry inside one
//    bucket is reached by a short scan; rows end in sentinels whose last slot is 0xfff).  This
//    takes ns * (64 + 4 * (ns + 3)) bytes whatever the table precision: 1 KB for binned
//    qualities, 12 KB for 46 symbols -- where byte-per-slot rows need 9-184 KB.
//  * the same compact layout in global memory (from the arena, L1/L2 cached) for alphabets that
//    do not fit the shared-memory budget: 12 KB for 46 symbols, 282 KB for all 256 -- against
//    0.2-1.3 MB for the reference's byte-per-slot rows (rANS_static4x16pr.c:922-930).
//
// Shared memory of one group: [0,256) rank -> symbol, [256,512) symbol -> rank, frequency
// scratch, the word ring, then TAB bytes of compact tables.
template <int NWAY, int SZ = 0> struct O1Smem {               // SZ: 0 regular, 1 small, 2 medium (4-way only),
    static constexpr bool SMALL = SZ == 1;                    //     3 / 4: rows of 8 / 16 entries for the register-table kernels
    static constexpr bool REG = SZ >= 3;
    static constexpr int REG_NS = SZ == 3 ? 8 : 16;
    // a register-table row: REG_NS entries, then its entry count, padded to REG_NS * 4 + 16 bytes -- rows then start in
    // all eight 16-byte bank groups (48 / 80-byte stride) instead of two or four, which is what the 128-bit row
    // loads of lanes sitting in different contexts collide on
    static constexpr int REG_ROW = REG_NS * 4 + 16;
    static constexpr int UNRANK = 0, RANK = 256;
    // the frequency scratch is only live during set-up; X_32 (1 KB ring) and the small 4-way variants
    // (alphabets of <= 16 symbols, 256-byte ring) let it share the word ring's memory
    static constexpr bool OVERLAY = (NWAY == 32) || SMALL || REG;
    static constexpr int FTMP = 512, RINGO = OVERLAY ? 512 : 1536;
    static constexpr int TABO = RINGO + GroupCfg<NWAY>::RING + (NWAY == 4 ? 64 : 0);   // 4-way: + mirror of the ring's first 64 bytes
    // X_32: 12992 B (<= 48 symbols, 15 warps / SM) or, SMALL, 4608 B (<= 25 symbols, 28 warps / SM);
    // 4-way: 3072 B per group (<= 19 symbols, 40 groups / SM), SMALL 1024 B (<= 9 symbols, 120 groups / SM),
    // medium 12544 B (<= 47 symbols, 16 groups / SM), REG ns x ns entries (256 B / 1 KB)
    static constexpr int TAB = (NWAY == 32) ? (SMALL ? (int)O1_SMALL_TAB : 12992)
                               : REG ? REG_NS * REG_ROW
                                     : (SMALL ? 1024 : SZ == 2 ? (int)O1_TAB4M : (int)O1_TAB4);
    static constexpr int FTMP_ENTRIES = OVERLAY ? GroupCfg<NWAY>::RING / 4 : 256;
    static constexpr int STRIDE = TABO + TAB;                        // multiple of 16
    static constexpr int TOTAL = STRIDE * GroupCfg<NWAY>::G;
};
struct O1Tables {
    uint32_t compact;       // 1: compact form in shared memory
    uint32_t tabs;          // compact: shared address of context 0's block [64 B coarse | 4 * (ns + 3) B entries]
    uint32_t bstride;       // compact: bytes per context block
    uint8_t* g_tabs;        // the same layout in global memory when it does not fit (compact == 0)
    uint32_t ns, shift;
    uint32_t row_off;       // entries start this far into a context block (64: behind the coarse index; 0: no index)
    uint32_t cshift;        // position of an entry's last-slot field (20; register-table kernels: 32 - shift)
    uint32_t cnt_off;       // register-table rows: the row's entry count is stored at this offset (0: not stored)
};

// Group-cooperative: frequencies of one context (shared u32 array F, indexed by rank, `ns`
// entries used) -> compact row + coarse index.  Returns false if they do not sum to 1 << shift
// after the power-of-two shift (…4x16pr.c:982-997).
template <int NWAY>
__device__ bool build_o1_row_compact(const Grp<NWAY>& G, uint32_t F, const O1Tables& T, uint32_t ctx, uint32_t Tsum,
                                     bool allow_4095 = false) {
    const uint32_t M = 1u << T.shift, ns = T.ns, bs = T.shift - 6;
    uint32_t sh = 0;
    if (Tsum < M) while ((Tsum << sh) < M) sh++;
    const uint32_t K = (ns + NWAY - 1) / NWAY, r0 = G.glane * K;     // each lane owns K consecutive ranks
    uint32_t mine = 0, nz = 0;
    bool bad = false;
    for (uint32_t k = 0; k < K; k++) {
        const uint32_t r = r0 + k;
        const uint32_t f = (r < ns) ? (lds_u32(F + 4 * r) << sh) : 0u;
        if (f > M) bad = true;
        mine += f;
        nz += (f != 0);
    }
    uint32_t total, nzt;
    uint32_t c = G.exscan(bad ? 2 * M : mine, &total);
    uint32_t idx = G.exscan(nz, &nzt);
    // 4x8 tables written by the reference sum to M - 1 (rANS_static.c:122-130); slot M - 1 then
    // belongs to no symbol and resolves to the row's sentinel (never reached by a valid stream)
    if (total != M && !(allow_4095 && total == M - 1)) return false;
    const uint32_t crs = T.tabs + ctx * T.bstride, row = crs + T.row_off;   // shared form
    uint8_t* gcrs = T.g_tabs + (size_t)ctx * T.bstride;             // global form
    uint32_t* grow = reinterpret_cast<uint32_t*>(gcrs + 64);
    const bool coarse = T.row_off != 0;
    if (T.cnt_off && G.glane == 0) sts_u32(crs + T.cnt_off, nzt);
    for (uint32_t k = 0; k < K; k++) {
        const uint32_t r = r0 + k;
        if (r >= ns) break;
        const uint32_t f = lds_u32(F + 4 * r) << sh;
        if (!f) continue;
        const uint32_t last = c + f - 1;
        const uint32_t e = (last << T.cshift) | (r << 12) | (f - 1);
        if (T.compact) {
            sts_u32(row + 4 * idx, e);
            if (coarse) for (uint32_t q = (c + (1u << bs) - 1) >> bs; q <= (last >> bs); q++) sts_u8(crs + q, idx);
        } else {
            grow[idx] = e;
            for (uint32_t q = (c + (1u << bs) - 1) >> bs; q <= (last >> bs); q++) gcrs[q] = (uint8_t)idx;
        }
        idx++;
        c += f;
    }
    G.sync();
    return true;
}

// Per-group order-1 set-up.  Returns 0 ok, ST_FORMAT or ST_ARENA (group-uniform).
template <int NWAY, bool BYTE, int SZ>
__device__ int32_t o1_setup(const Grp<NWAY>& G, DecWork* W, const DecJob& job, uint8_t* gsm, uint32_t base,
                            O1Tables* Tout, uint32_t* R, const uint8_t** first_word, uint32_t* ctx0) {
    using S = O1Smem<NWAY, SZ>;
    const uint32_t unrank = base + S::UNRANK, rank = base + S::RANK, Ftmp = base + S::FTMP, tabs = base + S::TABO;
    const uint8_t* in_end = job.in + job.in_len;

    // ---- phase 1 (one lane): locate the table bytes, read the alphabet
    for (uint32_t k = G.glane; k < 256; k += NWAY) { sts_u8(rank + k, 0xffu); sts_u8(unrank + k, 0u); }
    G.sync();
    uint32_t err = 0, shift = 12, ns = 0u;
    GRd rd{nullptr, nullptr};
    const uint8_t* body = nullptr;
    if (G.glane == 0) {
        if (BYTE) {                                          // rANS_static.c:676-: tables follow the 9-byte header
            if (job.in_len < 27) err = 1;
            rd.p = job.in + 9; rd.end = in_end;
            // The 4x8 format names no alphabet up front: a first pass over the tables marks every
            // byte value that occurs as a context or as a symbol (plus 0, the segment-start context).
            if (!err) {
                GRd sc = rd;
                uint32_t run_i = 0, c = sc.get(), dummy;
                sts_u8(rank + 0, 0u);
                for (;;) {
                    sts_u8(rank + c, 0u);
                    if (sc.p + 16 > sc.end ||
                        !parse_table_4x8(sc, [&](uint32_t j, uint32_t) { sts_u8(rank + j, 0u); }, &dummy, true)) { err = 1; break; }
                    if (!run_i && c + 1 == sc.peek()) { sc.get(); c++; run_i = sc.get(); }
                    else if (run_i) { run_i--; if (++c > 255) { err = 1; break; } }
                    else c = sc.get();
                    if (!c) break;
                }
            }
        } else if (job.in_len < 4 * NWAY) {                  // :872
            err = 1;
        } else {
            uint32_t h = job.in[0];
            shift = h >> 4;
            if (shift != 10 && shift != 12) err = 1;         // the reference has loops for 12 and 10 only
            if (h & 1) {                                     // :944-955: the planner had the table expanded into aux
                uint32_t usz, csz;                           // by an order-0 job (and applied the :948 check)
                const uint8_t* p = job.in + 1;
                p += var_get_u32(p, in_end, &usz);
                p += var_get_u32(p, in_end, &csz);
                if (!job.aux || (int64_t)csz >= (int64_t)(in_end - p) - 16) err = 1;
                rd.p = job.aux; rd.end = job.aux + usz;
                body = p + csz;
            } else {
                rd.p = job.in + 1; rd.end = in_end;
            }
            if (!err && (!read_alphabet(rd, rank, 0u) || !rd.more())) err = 1;   // :959-965
        }
        {
            if (!err) {
                for (uint32_t s = 0; s < 256 && !err; s++)
                    if (lds_u8(rank + s) == 0) { sts_u8(unrank + ns, s); sts_u8(rank + s, 0xfeu); ns++; }
                // second pass: 0xfe marks -> ranks (a rank can legitimately be 0xfe/0xff when ns > 254)
                uint32_t r2 = 0;
                for (uint32_t s = 0; s < 256 && !err; s++)
                    if (lds_u8(rank + s) == 0xfeu) sts_u8(rank + s, r2++);
                if (!err && (ns == 0 || lds_u8(unrank + 0) != 0)) err = 1;       // symbol 0 = context of segment starts
            }
        }
    }
    G.sync();
    err = G.bcast(err); ns = G.bcast(ns); shift = G.bcast(shift);
    if (err || ns > (uint32_t)S::FTMP_ENTRIES) return ST_FORMAT;     // (the planner only routes fitting alphabets)
    const uint32_t M = 1u << shift;

    // ---- table storage
    O1Tables T;
    T.ns = ns; T.shift = shift;
    T.g_tabs = nullptr;
    T.bstride = S::REG ? (uint32_t)S::REG_ROW : 64 + 4 * (ns + 3);
    T.row_off = S::REG ? 0u : 64u;
    T.cshift = S::REG ? 32u - shift : 20u;
    T.cnt_off = S::REG ? 4u * S::REG_NS : 0u;
    T.tabs = tabs;
    T.compact = S::REG ? 1u : ((o1_compact_bytes(ns) <= (uint32_t)S::TAB) ? 1u : 0u);
    if (S::REG && ns > (uint32_t)S::REG_NS) return ST_INTERNAL;          // (the planner counts the alphabet before routing here)
    const uint32_t nwords = ns * (T.bstride / 4);
    if (T.compact) {
        for (uint32_t k = G.glane; k < nwords; k += NWAY)                    // coarse: 0, entries: sentinels
            sts_u32(tabs + 4 * k, (S::REG ? (k % (T.bstride / 4)) >= (uint32_t)S::REG_NS : (k % (T.bstride / 4)) < 16) ? 0u : O1_SENTINEL);
    } else {
        uint8_t* a = nullptr;
        if (G.glane == 0) a = arena_alloc(W, (uint64_t)nwords * 4);
        T.g_tabs = const_cast<uint8_t*>(G.bcast_ptr(a));
        if (!T.g_tabs) return ST_ARENA;
        uint32_t* gw = reinterpret_cast<uint32_t*>(T.g_tabs);
        for (uint32_t k = G.glane; k < nwords; k += NWAY) gw[k] = (k % (T.bstride / 4)) < 16 ? 0u : O1_SENTINEL;
    }
    G.sync();

    // ---- phase 2: one row per context; lane 0 parses, the group fills
    if (!BYTE) {
        for (uint32_t ci = 0; ci < ns; ci++) {               // :967-998 (ascending symbol == ascending rank)
            for (uint32_t k = G.glane; k < ns; k += NWAY) sts_u32(Ftmp + 4 * k, 0u);
            G.sync();
            uint32_t Tsum = 0;
            if (G.glane == 0) {                              // decode_freq_d :327-358
                if (!rd.more()) err = 1;
                uint32_t zrun = 0;
                for (uint32_t sj = 0; sj < ns && rd.more() && !err; sj++) {
                    uint32_t f = 0;
                    if (zrun) zrun--;
                    else {
                        f = rd.varint();
                        if (f == 0) { if (!rd.more()) { err = 1; break; } zrun = rd.get(); }
                    }
                    sts_u32(Ftmp + 4 * sj, f);
                    Tsum += f;
                }
            }
            G.sync();
            err = G.bcast(err); Tsum = G.bcast(Tsum);
            if (err) return ST_FORMAT;
            if (!Tsum) continue;                             // :977-980
            if (!build_o1_row_compact<NWAY>(G, Ftmp, T, ci, Tsum)) return ST_FORMAT;
        }
    } else {
        // rANS_static.c:748-813: outer "sym [run]" list of contexts, one 4x8 table each (second pass:
        // frequencies land at their symbol's rank)
        uint32_t run_i = 0, ctx = 0;
        if (G.glane == 0) ctx = rd.get();
        for (;;) {
            for (uint32_t k = G.glane; k < ns; k += NWAY) sts_u32(Ftmp + 4 * k, 0u);
            G.sync();
            uint32_t x = 0;
            if (G.glane == 0) {
                if (rd.p + 16 > rd.end ||
                    !parse_table_4x8(rd, [&](uint32_t j, uint32_t f) { sts_u32(Ftmp + 4 * lds_u8(rank + j), f); }, &x, true)) err = 1;
                else if (x < 4095 || x > 4096) err = 1;      // :797
            }
            G.sync();
            err = G.bcast(err); ctx = G.bcast(ctx); x = G.bcast(x);
            if (err) return ST_FORMAT;
            const uint32_t rctx = lds_u8(rank + ctx);
            if (!build_o1_row_compact<NWAY>(G, Ftmp, T, rctx, 4096u, true)) return ST_FORMAT;
            uint32_t more = 0;
            if (G.glane == 0) {
                if (!run_i && ctx + 1 == rd.peek()) { rd.get(); ctx++; run_i = rd.get(); }
                else if (run_i) { run_i--; if (++ctx > 255) err = 1; }
                else ctx = rd.get();
                more = (!err && ctx != 0) ? 1u : 0u;
            }
            more = G.bcast(more); err = G.bcast(err);
            if (err) return ST_FORMAT;
            if (!more) break;
        }
    }
    __threadfence_block();
    G.sync();
    if (S::REG && S::REG_NS == 16) {
        // an entry also tells whether the row of ITS symbol (the next context) holds more than 8 entries: only then
        // does the decode loop fetch the row's second half (bit 16; ranks are < 16 here)
        for (uint32_t k = G.glane; k < ns * 16u; k += NWAY) {
            const uint32_t a = tabs + (k >> 4) * T.bstride + 4 * (k & 15u);
            const uint32_t e = lds_u32(a);
            if (lds_u32(tabs + ((e >> 12) & 15u) * T.bstride + T.cnt_off) > 8u) sts_u32(a, e | 0x10000u);
        }
        G.sync();
    }

    // ---- states
    const uint8_t* sp = nullptr;
    if (G.glane == 0) sp = body ? body : rd.p;
    sp = G.bcast_ptr(sp);
    if (sp + 4 * NWAY > in_end) return ST_FORMAT;            // :1005
    uint32_t r0 = ld_u32_le(sp + 4 * G.glane);
    if (!G.all(r0 >= (BYTE ? (1u << 23) : (1u << 15)))) return ST_FORMAT;
    *R = r0;
    *first_word = sp + 4 * NWAY;
    *ctx0 = lds_u8(rank + 0);
    *Tout = T;
    return ST_OK;
}

// Per-lane output sink of the order-1 loop.  A lane writes its own contiguous segment, so a
// plain byte store per symbol would cost one memory transaction per lane per step; instead the
// newest 16 bytes ride in registers and leave as one 128-bit store each time a 16-byte line is
// complete.  The partial lines at either end of a segment are written bytewise (neighbouring
// lanes' segments share those lines).
__device__ __noinline__ void sink_drain(uint8_t* line, uint32_t k, uint32_t skip, uint32_t w0, uint32_t w1, uint32_t w2,
                                        uint32_t w3) {
    // the newest (k - skip) bytes sit at the top of the window; they belong at line[skip .. k)
    for (uint32_t q = k; q > skip; q--) {
        line[q - 1] = (uint8_t)(w3 >> 24);
        w3 = __funnelshift_l(w2, w3, 8); w2 = __funnelshift_l(w1, w2, 8); w1 = __funnelshift_l(w0, w1, 8); w0 <<= 8;
    }
}

struct ByteSink {
    uint8_t* line;               // 16-byte aligned address of the line being gathered
    uint32_t w0, w1, w2, w3;     // the newest 16 bytes, newest in the top byte of w3
    uint32_t k;                  // bytes of this line accounted for (including `skip`)
    uint32_t skip;               // leading bytes of the first line that belong to the previous segment
    __device__ __forceinline__ void init(uint8_t* q) {
        const uint32_t lo = (uint32_t)(reinterpret_cast<uintptr_t>(q) & 15);
        line = q - lo; k = skip = lo; w0 = w1 = w2 = w3 = 0;
    }
    __device__ __forceinline__ void put(uint32_t s) {
        w0 = __funnelshift_r(w0, w1, 8); w1 = __funnelshift_r(w1, w2, 8); w2 = __funnelshift_r(w2, w3, 8);
        w3 = (w3 >> 8) | (s << 24);
        if (++k == 16) {
            if (skip == 0) *reinterpret_cast<uint4*>(line) = make_uint4(w0, w1, w2, w3);
            else { sink_drain(line, 16, skip, w0, w1, w2, w3); skip = 0; }
            line += 16; k = 0;
        }
    }
    // four bytes at once (oldest in the low byte); requires k % 4 == 0
    __device__ __forceinline__ void put4(uint32_t v) {
        w0 = w1; w1 = w2; w2 = w3; w3 = v;
        k += 4;
        if (k == 16) {
            if (skip == 0) *reinterpret_cast<uint4*>(line) = make_uint4(w0, w1, w2, w3);
            else { sink_drain(line, 16, skip, w0, w1, w2, w3); skip = 0; }
            line += 16; k = 0;
        }
    }
    __device__ __forceinline__ void finish() { if (k > skip) sink_drain(line, k, skip, w0, w1, w2, w3); }
};

// One decode step of a lane (rANS_static4x16pr.c:1033-1047 / rANS_static.c:850-878): updates R
// and the lane's context `cs` (COMPACT: shared address of the context's table block; LUT form:
// the rank) and returns the decoded rank.  COMPACT: the warp holds compact tables only.
template <bool COMPACT>
__device__ __forceinline__ uint32_t o1_symbol(uint32_t& R, uint32_t& cs, const O1Tables& T, uint32_t mask) {
    const uint32_t m = R & mask;
    if (COMPACT || T.compact) {
        const uint32_t blk = COMPACT ? cs : T.tabs + cs * T.bstride;
        const uint32_t ci = lds_u8(blk + (m >> (T.shift - 6)));
        uint32_t ea = blk + 64 + 4 * ci;
        const uint32_t e0 = lds_u32(ea), e1 = lds_u32(ea + 4), e2 = lds_u32(ea + 8);
        const uint32_t mk = m << 20;                         // mk > e  <=>  m > last slot of e (top 12 bits)
        uint32_t e = (mk > e1) ? e2 : ((mk > e0) ? e1 : e0);
        const uint32_t q = R >> T.shift, qm = q + m;
        // x' = F * q + m - C with F - 1 = e[11:0] and C + F - 1 = e[31:20]
        R = (e & 0xfffu) * (q + 1u) + (qm - (e >> 20));
        if (mk > e) {                                        // >= 4 symbols share the bucket: scan on (sentinel-bounded)
            ea += 12;
            do { e = lds_u32(ea); ea += 4; } while (mk > e);
            R = (e & 0xfffu) * (q + 1u) + (qm - (e >> 20));
        }
        const uint32_t r = (e >> 12) & 0xffu;
        cs = COMPACT ? T.tabs + r * T.bstride : r;
        return r;
    }
    // the same look-up in global memory (cs = rank of the context)
    const uint8_t* blk = T.g_tabs + (size_t)cs * T.bstride;
    const uint32_t ci = blk[m >> (T.shift - 6)];
    const uint32_t* ep = reinterpret_cast<const uint32_t*>(blk + 64) + ci;
    const uint32_t e0 = ep[0], e1 = ep[1], e2 = ep[2];
    const uint32_t mk = m << 20;
    uint32_t e = (mk > e1) ? e2 : ((mk > e0) ? e1 : e0);
    if (mk > e) {
        ep += 3;
        do { e = *ep++; } while (mk > e);
    }
    const uint32_t q = R >> T.shift;
    R = (e & 0xfffu) * (q + 1u) + (q + m - (e >> 20));
    cs = (e >> 12) & 0xffu;
    return cs;
}

// `minit` (warp-uniform) = steps for which every lane of the warp is active; they run four to a
// ring check and, when every lane's segment starts 4-byte aligned (AL4, warp-uniform), four
// symbols to one sink operation.
template <int NWAY, bool BYTE, bool ALIGNED, bool COMPACT>
__device__ __forceinline__ void o1_loop(uint32_t R, WordRing<NWAY>& ring, const O1Tables T, uint32_t unrank,
                                        uint32_t ctx0, uint8_t* out, uint32_t seg, uint32_t tail, uint32_t minit,
                                        uint32_t maxit, const Grp<NWAY>& G) {
    const uint32_t lt = (NWAY == 32) ? lanemask_lt() : ((1u << G.glane) - 1u);
    const uint32_t mask = (1u << T.shift) - 1u;
    const uint32_t mine = seg + ((G.glane == NWAY - 1) ? tail : 0u);   // symbols this lane decodes
    const uint32_t group_steps = seg + tail;
    constexpr uint32_t LB = BYTE ? (1u << 23) : (1u << 15);            // renormalisation bound
    ByteSink sink;
    uint8_t* const op0 = out + (size_t)G.glane * seg;
    sink.init(op0);
    uint32_t cs = COMPACT ? T.tabs + ctx0 * T.bstride : ctx0;
    uint32_t i = 0;
    const bool al4 = __all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(op0) & 3) == 0);
    if (al4) {
        // (flushing the sink with a predicated store instead of a branch, once every lane's first partial line is out,
        //  was measured 1-2 % slower)
        auto four = [&]() {
            uint32_t pack = 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                Win win;
                if (NWAY == 4) win = win_load(ring.ring, ring.head);
                const uint32_t r = o1_symbol<COMPACT>(R, cs, T, mask);
                pack |= lds_u8(unrank + r) << (8 * u);
                if (NWAY == 4) R = win_renorm<BYTE>(R, R < LB, win, ring.head, lt, G.gshift);
                else R = renorm_step<NWAY, BYTE, ALIGNED>(R, R < LB, ring, lt, G.gshift);
            }
            sink.put4(pack);
        };
        if (NWAY == 4)                                       // 4-way: eight steps use <= 64 of the 128 bytes kept ahead
            for (; i + 8 <= minit; i += 8) {
                four(); four();
                ring.advance(G.glane, true);
            }
        for (; i + 4 <= minit; i += 4) {
            four();
            ring.advance(G.glane, true);
        }
    } else {
        for (; i + 4 <= minit; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                Win win;
                if (NWAY == 4) win = win_load(ring.ring, ring.head);
                const uint32_t r = o1_symbol<COMPACT>(R, cs, T, mask);
                sink.put(lds_u8(unrank + r));
                if (NWAY == 4) R = win_renorm<BYTE>(R, R < LB, win, ring.head, lt, G.gshift);
                else R = renorm_step<NWAY, BYTE, ALIGNED>(R, R < LB, ring, lt, G.gshift);
            }
            ring.advance(G.glane, true);
        }
    }
    for (; i < maxit; i++) {
        const bool act = i < mine;
        Win win;
        if (NWAY == 4) win = win_load(ring.ring, ring.head);
        if (act) {
            const uint32_t r = o1_symbol<COMPACT>(R, cs, T, mask);
            sink.put(lds_u8(unrank + r));
        }
        if (NWAY == 4) R = win_renorm<BYTE>(R, act && R < LB, win, ring.head, lt, G.gshift);
        else R = renorm_step<NWAY, BYTE, ALIGNED>(R, act && R < LB, ring, lt, G.gshift);
        ring.advance(G.glane, i < group_steps);
    }
    sink.finish();
}

template <int NWAY, bool BYTE, int SZ>
__global__ void __launch_bounds__(32, (SZ == 1 && NWAY == 32) ? 28 : 1) dec_o1_kernel(DecWork* W, int32_t* status, uint32_t kind) {
    using C = GroupCfg<NWAY>;
    using S = O1Smem<NWAY, SZ>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const Grp<NWAY> G;
    uint8_t* gsm = smem_raw + G.g * S::STRIDE;
    uint32_t base = smem_addr(gsm);
    asm volatile("" : "+r"(base));                      // keep the window base in a register
    const uint32_t njobs = W->njobs[kind];
    const DecJob* jobs = W->jobs[kind];

    // Jobs are claimed from an atomic cursor.  The host shapes the launch so that every SM holds the
    // same number of CTAs (shaped_launch below): a batch that fits in one wave is then spread evenly.
    for (;;) {
        uint32_t j0 = 0;
        if (lane_id() == 0) j0 = atomicAdd(&W->next[kind], (uint32_t)C::G);
        j0 = __shfl_sync(0xffffffffu, j0, 0);
        if (j0 >= njobs) break;
        // (the groups of a last, partly filled warp repeat the list's last job -- identical bytes to identical
        //  addresses -- instead of idling: an idle group would keep the whole warp out of the all-lanes-active loops,
        //  i.e. at half speed, and the kernel ends with its slowest warp)
        const uint32_t ji = min(j0 + G.g, njobs - 1u);
        const bool active = true;
        DecJob job = make_job(nullptr, 0, nullptr, 0, 0);
        O1Tables T;
        T.compact = 1; T.tabs = base + S::TABO; T.bstride = 0;
        T.g_tabs = nullptr; T.ns = 1; T.shift = 12; T.row_off = 64; T.cshift = 20; T.cnt_off = 0;
        uint32_t R = 0, ctx0 = 0;
        const uint8_t* first_word = nullptr;
        bool ok = false;
        if (active) {
            job = jobs[ji];
            // an order-0 job may have been expanding this stream's table: skip if that (or anything else) failed
            int32_t st = (job.aux && status[job.blk] != ST_OK) ? ST_FORMAT
                                                               : o1_setup<NWAY, BYTE, SZ>(G, W, job, gsm, base, &T, &R, &first_word, &ctx0);
            ok = st == ST_OK;
            if (!ok && G.glane == 0) set_status(status, job.blk, st);
        }
        WordRing<NWAY> ring;
        ring.init(first_word, job.in + job.in_len, base + S::RINGO, G, ok, NWAY == 4);
        __syncwarp();
        const uint32_t seg = ok ? job.out_len / NWAY : 0, tail = ok ? job.out_len - seg * NWAY : 0;
        const uint32_t maxit = __reduce_max_sync(0xffffffffu, seg + tail), minit = __reduce_min_sync(0xffffffffu, seg);
        const bool aligned = !BYTE && __all_sync(0xffffffffu, (ring.head & 1u) == 0);
        const bool compact = __all_sync(0xffffffffu, T.compact != 0);
        const uint32_t unrank = base + S::UNRANK;
        if (compact && aligned)      o1_loop<NWAY, BYTE, true, true>(R, ring, T, unrank, ctx0, job.out, seg, tail, minit, maxit, G);
        else if (compact)            o1_loop<NWAY, BYTE, false, true>(R, ring, T, unrank, ctx0, job.out, seg, tail, minit, maxit, G);
        else                         o1_loop<NWAY, BYTE, false, false>(R, ring, T, unrank, ctx0, job.out, seg, tail, minit, maxit, G);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// 4-way / 4x8 streams with small alphabets: tables in registers, renormalisation bytes as a window
// ------------------------------------------------------------------------------------------
// A 4-way stream is four lanes of strictly serial work, so at the batch sizes of BASELINE.json (4096 blocks:
// 3.5 warps per SM) its speed is the latency of one step, not throughput.  The look-up-table kernels above spend
// three shared-memory round trips (~30 cycles each) per step: symbol, (F, C), renormalisation word.  Here
//   * the stream's table (order 0) or the current context's row (order 1) sits in NS registers per lane as packed
//     entries (C+F-1) << cshift | sym-or-rank << 12 | (F-1), and the symbol is found by a branch-free binary search
//     (log2 NS compare/select levels, ~9 cycles each); an order-1 lane loads its NEXT row (two or four 128-bit
//     shared loads) as soon as the symbol is known, behind the renormalisation chain;
//   * the 8 bytes at the group's read position (a step consumes at most 4 words, or 4 x 2 bytes for 4x8) are fetched
//     at the START of the step, and after the ballot each lane picks its word with two byte permutes -- the
//     selector comes from a permute-table indexed by the ballot bits of the lower lanes, so neither a popc nor a
//     shared-memory load sits between the ballot and the next state.
// Results are identical to the table kernels by construction (same F, same C); alphabets of up to 16 symbols.
// First entry whose last slot is >= the probe (entries ascending, padded with sentinels): M > e  <=>  slot beyond e.
template <int NS> __device__ __forceinline__ uint32_t reg_search(const uint32_t (&e)[NS], uint32_t M);
template <> __device__ __forceinline__ uint32_t reg_search<8>(const uint32_t (&e)[8], uint32_t M) {
    const bool p2 = M > e[3];
    const uint32_t a0 = p2 ? e[4] : e[0], a1 = p2 ? e[5] : e[1], a2 = p2 ? e[6] : e[2], a3 = p2 ? e[7] : e[3];
    const bool p1 = M > a1;
    const uint32_t b0 = p1 ? a2 : a0, b1 = p1 ? a3 : a1;
    return (M > b0) ? b1 : b0;
}
template <> __device__ __forceinline__ uint32_t reg_search<16>(const uint32_t (&e)[16], uint32_t M) {
    const bool p3 = M > e[7];
    uint32_t h[8];
#pragma unroll
    for (int k = 0; k < 8; k++) h[k] = p3 ? e[8 + k] : e[k];
    return reg_search<8>(h, M);
}

// Entries of one row into registers.  NS = 16: the second half only when `lng` (the row holds more than 8 entries),
// else sentinels -- binned qualities have 9 contexts (symbol 0 is forced into the alphabet) of at most 8 symbols.
template <int NS> __device__ __forceinline__ void load_row(uint32_t (&e)[NS], uint32_t addr, uint32_t lng = 1u) {
#pragma unroll
    for (int k = 0; k < NS; k += 4) {
        uint4 v = make_uint4(O1_SENTINEL, O1_SENTINEL, O1_SENTINEL, O1_SENTINEL);
        if (k < 8)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr + 4 * k));
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                         : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "r"(addr + 4 * k), "r"(lng));
        e[k] = v.x; e[k + 1] = v.y; e[k + 2] = v.z; e[k + 3] = v.w;
    }
}

// One decode step from a register row.  Returns the entry (symbol / rank in bits 12.., long-row flag in bit 16 for
// order 1); x becomes the new state before renormalisation, *p whether it must be renormalised.  shift / sh32 =
// 32 - shift are per-lane values.
template <int NS, bool BYTE>
__device__ __forceinline__ uint32_t reg_symbol(uint32_t& x, const uint32_t (&E)[NS], uint32_t shift, uint32_t sh32, uint32_t mask, bool* p) {
    constexpr uint32_t L = BYTE ? (1u << 23) : (1u << 15);
    const uint32_t e = reg_search<NS>(E, x << sh32);
    const uint32_t q = x >> shift, m = x & mask;
    // x' = F q + m - C with F - 1 = e[11:0] and C + F - 1 = e >> sh32:  X = (F-1)(q+1) + q + m = x' + (C+F-1)
    const uint32_t last = e >> sh32;
    const uint32_t X = (e & 0xfffu) * (q + 1u) + (q + m);
    *p = X < last + L;
    x = X - last;
    return e;
}

struct RegSmem0 {            // order 0: per group [entries NS x 4 | ring 256 + 64 mirror]
    static constexpr int ENT = 0, RINGO = 64, STRIDE = 64 + 256 + 64, TOTAL = 8 * STRIDE;
};

// Order-0 set-up (one lane parses straight from global memory): entries (C+F-1) << 20 | sym << 12 | (F-1) for the
// symbols of non-zero frequency, in cumulative order, padded with sentinels.
template <int NS, bool BYTE>
__device__ bool o0r_setup(const Grp<4>& G, const DecJob& job, uint32_t gsm, uint32_t* R, uint32_t* first_word) {
    constexpr uint32_t L = BYTE ? (1u << 23) : (1u << 15);
    const uint32_t ent = gsm + RegSmem0::ENT, syms = gsm + RegSmem0::RINGO, ftmp = syms + 64;   // ring: free for now
    const uint8_t* in_end = job.in + job.in_len;
    const uint32_t hdr0 = BYTE ? 9u : 0u;
    uint32_t tab = 0;
    if (G.glane == 0 && job.in_len >= hdr0 + 16) {
        GRd rd{job.in + hdr0, in_end};
        uint32_t ns = 0, sum = 0;
        bool ok;
        if (BYTE) {
            ok = parse_table_4x8(rd, [&](uint32_t j, uint32_t f) {
                if (ns < (uint32_t)NS) { sts_u8(syms + ns, j); sts_u32(ftmp + 4 * ns, f); }
                ns++; }, &sum, false);
        } else {
            uint32_t cnt = 0;
            ok = list_alphabet(rd, &ns, [&](uint32_t j) { if (cnt < (uint32_t)NS) sts_u8(syms + cnt, j); cnt++; });
            if (ok && ns <= (uint32_t)NS)
                for (uint32_t i = 0; i < ns; i++) { const uint32_t f = rd.varint(); sts_u32(ftmp + 4 * i, f); sum += f; }
        }
        ok = ok && ns >= 1 && ns <= (uint32_t)NS;
        uint32_t sh = 0;
        if (ok && !BYTE && sum != 0 && sum < 4096) while ((sum << sh) < 4096) sh++;      // normalise_freq_shift
        uint32_t c = 0, idx = 0;
        for (uint32_t i = 0; ok && i < ns; i++) {
            const uint32_t f = lds_u32(ftmp + 4 * i) << sh;
            if (!f) continue;
            if (f > 4096 - c) { ok = false; break; }
            sts_u32(ent + 4 * idx, ((c + f - 1) << 20) | (lds_u8(syms + i) << 12) | (f - 1));
            idx++;
            c += f;
        }
        if (ok && c != 4096 && !(BYTE && c == 4095)) ok = false;                         // :551 / rANS_static.c:305
        for (uint32_t k = idx; k < (uint32_t)NS; k++) sts_u32(ent + 4 * k, O1_SENTINEL);
        if (ok) tab = (uint32_t)(rd.p - job.in);
    }
    tab = G.bcast(tab);
    if (tab == 0 || tab + 16 > job.in_len) return false;
    const uint32_t r0 = ld_u32_le(job.in + tab + 4 * G.glane);
    if (!G.all(r0 >= L)) return false;
    *R = r0;
    *first_word = tab + 16;
    G.sync();
    return true;
}

template <int NS, bool BYTE>
__global__ void __launch_bounds__(32) dec_o0r_kernel(DecWork* W, int32_t* status, uint32_t kind) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const Grp<4> G;
    uint32_t base = smem_addr(smem_raw) + G.g * RegSmem0::STRIDE;
    asm volatile("" : "+r"(base));
    const uint32_t njobs = W->njobs[kind];
    const DecJob* jobs = W->jobs[kind];
    const uint32_t lt4 = (1u << G.glane) - 1u;
    for (;;) {
        uint32_t j0 = 0;
        if (lane_id() == 0) j0 = atomicAdd(&W->next[kind], 8u);
        j0 = __shfl_sync(0xffffffffu, j0, 0);
        if (j0 >= njobs) break;
        // (the groups of a last, partly filled warp repeat the list's last job -- identical bytes to identical
        //  addresses -- instead of idling: an idle group would keep the whole warp out of the all-lanes-active loops,
        //  i.e. at half speed, and the kernel ends with its slowest warp)
        const uint32_t ji = min(j0 + G.g, njobs - 1u);
        const bool active = true;
        DecJob job = make_job(nullptr, 0, nullptr, 0, 0);
        uint32_t R = 0, first_word = 0;
        bool ok = false;
        if (active) {
            job = jobs[ji];
            ok = o0r_setup<NS, BYTE>(G, job, base, &R, &first_word);
            if (!ok && G.glane == 0) set_status(status, job.blk, ST_FORMAT);
        }
        uint32_t E[NS];
        load_row<NS>(E, base + RegSmem0::ENT);
        if (!ok) {
#pragma unroll
            for (int k = 0; k < NS; k++) E[k] = O1_SENTINEL;
        }
        __syncwarp();
        WordRing<4> ring;
        ring.init(job.in + first_word, job.in + job.in_len, base + RegSmem0::RINGO, G, ok, true);
        __syncwarp();
        const uint32_t iters = ok ? job.out_len / 4 : 0, rem = ok ? job.out_len % 4 : 0;
        const uint32_t maxit = __reduce_max_sync(0xffffffffu, iters), minit = __reduce_min_sync(0xffffffffu, iters);
        uint8_t* op = job.out + G.glane;
        uint32_t i = 0;
        const bool al4 = __all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(job.out) & 3) == 0);
        if (al4 && minit >= 12) {
            // Eight steps to a ring check (a step consumes at most 8 bytes, the ring keeps 128 ahead), and the
            // lane-transposed output word of steps 4..7 is exchanged and stored during the NEXT round's steps:
            // the two shuffles' latency sat exposed at the end of every four steps (ncu: 7 % of the samples, and
            // as much again for the ring check).  Measured 216 -> 271 GB/s at 4096 blocks (4x8: 142 -> 160).
            auto four = [&]() -> uint32_t {
                uint32_t w = 0;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const Win win = win_load(ring.ring, ring.head);
                    bool p;
                    const uint32_t sy = (reg_symbol<NS, BYTE>(R, E, 12u, 20u, 0xfffu, &p) >> 12) & 0xffu;
                    w |= sy << (8 * u);
                    R = win_renorm<BYTE>(R, p, win, ring.head, lt4, G.gshift);
                }
                return w;
            };
            uint32_t pend = four();                          // symbols i .. i+3 of every lane, not yet written
            ring.advance(G.glane, true);
            for (i = 4; i + 8 <= minit; i += 8) {
                *reinterpret_cast<uint32_t*>(op + 3 * G.glane) = transpose4x4(pend, G.glane);
                const uint32_t w0 = four();
                *reinterpret_cast<uint32_t*>(op + 16 + 3 * G.glane) = transpose4x4(w0, G.glane);
                pend = four();
                op += 32;
                ring.advance(G.glane, true);
            }
            *reinterpret_cast<uint32_t*>(op + 3 * G.glane) = transpose4x4(pend, G.glane);
            op += 16;
        }
        for (; i + 4 <= minit; i += 4) {                     // every lane of the warp active: four steps to a ring check
            uint32_t w = 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const Win win = win_load(ring.ring, ring.head);
                bool p;
                const uint32_t sy = (reg_symbol<NS, BYTE>(R, E, 12u, 20u, 0xfffu, &p) >> 12) & 0xffu;
                w |= sy << (8 * u);
                R = win_renorm<BYTE>(R, p, win, ring.head, lt4, G.gshift);
            }
            if (al4) *reinterpret_cast<uint32_t*>(op + 3 * G.glane) = transpose4x4(w, G.glane);   // 16 contiguous bytes per group
            else {
#pragma unroll
                for (int u = 0; u < 4; u++) op[4 * u] = (uint8_t)(w >> (8 * u));
            }
            op += 16;
            ring.advance(G.glane, true);
        }
        for (; i < maxit; i++) {
            const bool act = i < iters;
            const Win win = win_load(ring.ring, ring.head);
            uint32_t x = R;
            bool p;
            const uint32_t sy = (reg_symbol<NS, BYTE>(x, E, 12u, 20u, 0xfffu, &p) >> 12) & 0xffu;
            if (act) { R = x; *op = (uint8_t)sy; op += 4; }
            R = win_renorm<BYTE>(R, act && p, win, ring.head, lt4, G.gshift);
            ring.advance(G.glane, act);
        }
        if (G.glane < rem) {                                 // the last n % 4 symbols: peek only (rANS_static.c:346-355)
            const uint32_t e = reg_search<NS>(E, R << 20);
            *op = (uint8_t)(e >> 12);
        }
        __syncwarp();
    }
}

// Order 1: the row of the lane's current context in registers; the next row is fetched the moment the symbol is known.
template <int NS, bool BYTE>
__global__ void __launch_bounds__(32) dec_o1r_kernel(DecWork* W, int32_t* status, uint32_t kind) {
    constexpr int SZ = NS == 8 ? 3 : 4;
    using S = O1Smem<4, SZ>;
    constexpr uint32_t RS = S::REG_ROW;                   // bytes per row
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const Grp<4> G;
    uint8_t* gsm = smem_raw + G.g * S::STRIDE;
    uint32_t base = smem_addr(gsm);
    asm volatile("" : "+r"(base));
    const uint32_t njobs = W->njobs[kind];
    const DecJob* jobs = W->jobs[kind];
    const uint32_t lt4 = (1u << G.glane) - 1u;
    const uint32_t rows = base + S::TABO, unrank = base + S::UNRANK;
    for (;;) {
        uint32_t j0 = 0;
        if (lane_id() == 0) j0 = atomicAdd(&W->next[kind], 8u);
        j0 = __shfl_sync(0xffffffffu, j0, 0);
        if (j0 >= njobs) break;
        // (the groups of a last, partly filled warp repeat the list's last job -- identical bytes to identical
        //  addresses -- instead of idling: an idle group would keep the whole warp out of the all-lanes-active loops,
        //  i.e. at half speed, and the kernel ends with its slowest warp)
        const uint32_t ji = min(j0 + G.g, njobs - 1u);
        const bool active = true;
        DecJob job = make_job(nullptr, 0, nullptr, 0, 0);
        O1Tables T;
        T.compact = 1; T.tabs = rows; T.bstride = RS; T.g_tabs = nullptr; T.ns = 1; T.shift = 12; T.row_off = 0; T.cshift = 20; T.cnt_off = 4 * NS;
        uint32_t R = 0, ctx0 = 0;
        const uint8_t* first_word = nullptr;
        bool ok = false;
        if (active) {
            job = jobs[ji];
            int32_t st = job.aux ? ST_INTERNAL : o1_setup<4, BYTE, SZ>(G, W, job, gsm, base, &T, &R, &first_word, &ctx0);
            ok = st == ST_OK;
            if (!ok && G.glane == 0) set_status(status, job.blk, st);
        }
        if (!ok) for (uint32_t k = G.glane; k <= (uint32_t)NS; k += 4) sts_u32(rows + 4 * k, k < (uint32_t)NS ? O1_SENTINEL : 0u);   // a harmless row 0
        __syncwarp();
        WordRing<4> ring;
        ring.init(first_word, job.in + job.in_len, base + S::RINGO, G, ok, true);
        __syncwarp();
        const uint32_t seg = ok ? job.out_len / 4 : 0, tail = ok ? job.out_len - seg * 4 : 0;
        const uint32_t maxit = __reduce_max_sync(0xffffffffu, seg + tail), minit = __reduce_min_sync(0xffffffffu, seg);
        const uint32_t shift = ok ? T.shift : 12u, sh32 = 32u - shift, mask = (1u << shift) - 1u;
        const uint32_t mine = seg + ((G.glane == 3) ? tail : 0u), group_steps = seg + tail;
        // rank -> symbol in registers (NS = 8: one permute; NS = 16: two and a select)
        uint32_t U[NS / 4];
#pragma unroll
        for (int k = 0; k < NS / 4; k++) U[k] = lds_u32(unrank + 4 * k);
        auto unrk = [&](uint32_t r) -> uint32_t {
            if (NS == 8) return prmt(U[0], U[1], r) & 0xffu;
            const uint32_t lo = prmt(U[0], U[1], r & 7u), hi = prmt(U[NS / 4 - 2], U[NS / 4 - 1], r & 7u);
            return ((r & 8u) ? hi : lo) & 0xffu;
        };
        uint32_t E[NS];
        {
            const uint32_t r0a = rows + (ok ? ctx0 : 0u) * RS;
            load_row<NS>(E, r0a, NS == 16 ? (uint32_t)(lds_u32(r0a + 4 * NS) > 8u) : 1u);
        }
        // next row: rank in bits 12-15 of the entry, "more than 8 entries" in bit 16
        auto fetch_row = [&](uint32_t e) { load_row<NS>(E, rows + ((e >> 12) & 15u) * RS, e & 0x10000u); };
        ByteSink sink;
        uint8_t* const op0 = job.out + (size_t)G.glane * seg;
        sink.init(op0);
        uint32_t i = 0;
        const bool al4 = __all_sync(0xffffffffu, (reinterpret_cast<uintptr_t>(op0) & 3) == 0);
        if constexpr (NS == 16) if (al4 && minit >= 8) {
            // No row of any of the warp's streams holds more than 8 entries (binned qualities: 9 contexts of at most 8
            // symbols): search 8 registers instead of 16 -- the first compare / select level, its register moves and the
            // predicated second-half loads were 18 of the 78 instructions of a step, and the kernel is bound by its
            // instruction count (ncu: the lone warp issues on 49 % of the cycles).
            bool lng = false;
            if (ok) for (uint32_t k = G.glane; k < T.ns; k += 4) lng |= lds_u32(rows + k * RS + 4 * NS) > 8u;
            if (!__any_sync(0xffffffffu, lng)) {
                uint32_t E8[8];
#pragma unroll
                for (int k = 0; k < 8; k++) E8[k] = E[k];
                auto four8 = [&]() {
                    uint32_t pack = 0;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const Win win = win_load(ring.ring, ring.head);
                        bool p;
                        const uint32_t e = reg_symbol<8, BYTE>(R, E8, shift, sh32, mask, &p);
                        load_row<8>(E8, rows + ((e >> 12) & 15u) * RS);
                        pack |= unrk((e >> 12) & 15u) << (8 * u);
                        R = win_renorm<BYTE>(R, p, win, ring.head, lt4, G.gshift);
                    }
                    sink.put4(pack);
                };
                for (; i + 8 <= minit; i += 8) {
                    four8(); four8();
                    ring.advance(G.glane, true);
                }
#pragma unroll
                for (int k = 0; k < 8; k++) { E[k] = E8[k]; E[8 + k] = O1_SENTINEL; }
            }
        }
        if (al4) {
            auto four = [&]() {
                uint32_t pack = 0;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const Win win = win_load(ring.ring, ring.head);
                    bool p;
                    const uint32_t e = reg_symbol<NS, BYTE>(R, E, shift, sh32, mask, &p);
                    fetch_row(e);
                    pack |= unrk((e >> 12) & 15u) << (8 * u);
                    R = win_renorm<BYTE>(R, p, win, ring.head, lt4, G.gshift);
                }
                sink.put4(pack);
            };
            for (; i + 8 <= minit; i += 8) {                 // eight steps to a ring check (<= 64 of the 128 bytes kept ahead)
                four(); four();
                ring.advance(G.glane, true);
            }
            for (; i + 4 <= minit; i += 4) {
                four();
                ring.advance(G.glane, true);
            }
        } else {
            for (; i + 4 <= minit; i += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const Win win = win_load(ring.ring, ring.head);
                    bool p;
                    const uint32_t e = reg_symbol<NS, BYTE>(R, E, shift, sh32, mask, &p);
                    fetch_row(e);
                    sink.put(unrk((e >> 12) & 15u));
                    R = win_renorm<BYTE>(R, p, win, ring.head, lt4, G.gshift);
                }
                ring.advance(G.glane, true);
            }
        }
        for (; i < maxit; i++) {
            const bool act = i < mine;
            const Win win = win_load(ring.ring, ring.head);
            bool p = false;
            if (act) {
                const uint32_t e = reg_symbol<NS, BYTE>(R, E, shift, sh32, mask, &p);
                fetch_row(e);
                sink.put(unrk((e >> 12) & 15u));
            }
            R = win_renorm<BYTE>(R, act && p, win, ring.head, lt4, G.gshift);
            ring.advance(G.glane, i < group_steps);
        }
        sink.finish();
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// X_CAT copy: one CTA per job
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) copy_kernel(DecWork* W) {
    const uint32_t njobs = W->njobs[JK_COPY];
    for (uint32_t ji = blockIdx.x; ji < njobs; ji += gridDim.x) {
        DecJob job = W->jobs[JK_COPY][ji];
        const uint8_t* s = job.in;
        uint8_t* d = job.out;
        uint32_t n = job.out_len;
        if ((((uintptr_t)s ^ (uintptr_t)d) & 15) == 0 && n >= 64) {
            uint32_t head = (uint32_t)((16 - ((uintptr_t)d & 15)) & 15);
            for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) d[i] = s[i];
            uint32_t nv = (n - head) / 16;
            const uint4* sv = reinterpret_cast<const uint4*>(s + head);
            uint4* dv = reinterpret_cast<uint4*>(d + head);
            for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) dv[i] = sv[i];
            for (uint32_t i = head + nv * 16 + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
        } else {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// un-RLE, rle.c:142-187.  One 256-thread CTA per chain.
//   pass A: run-length varints -> values (scratch u32 per varint, carried across tiles)
//   pass B: literals -> output offsets (block scan with carry) -> fill
// ------------------------------------------------------------------------------------------
constexpr int RLE_T = 256;

template <typename T>
__device__ __forceinline__ T block_exscan(T v, T* warp_tot, T* total) {
    uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { T y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += y; }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    T pre = 0, tot = 0;
    for (int k = 0; k < RLE_T / 32; k++) { T t = warp_tot[k]; if (k < (int)w) pre += t; tot += t; }
    __syncthreads();
    *total = tot;
    return pre + x - v;
}

__global__ void __launch_bounds__(RLE_T) rle_kernel(DecWork* W, int32_t* status, uint32_t* out_len) {
    __shared__ uint8_t is_rle[256];
    __shared__ uint32_t wtot[RLE_T / 32];
    __shared__ unsigned long long wtot64[RLE_T / 32];
    __shared__ uint32_t* runval_s;
    __shared__ int skip_s;
    __shared__ uint32_t s_off[RLE_T + 1];
    __shared__ uint8_t s_lit[RLE_T];
    const uint32_t tid = threadIdx.x;
    for (uint32_t li = blockIdx.x; li < W->nrle; li += gridDim.x) {
        const uint32_t ci = W->rle_list[li];
        const Chain c = W->chains[ci];
        __syncthreads();
        if (tid == 0) {
            int skip = status[c.blk] != ST_OK;
            uint32_t nrs0 = 0;
            runval_s = nullptr;
            if (!skip) {
                nrs0 = c.meta[0] ? c.meta[0] : 256u;
                if (c.u_meta < 1 + nrs0) { set_status(status, c.blk, ST_FORMAT); skip = 1; }     // …4x16pr.c:1604
            }
            if (!skip) {
                runval_s = reinterpret_cast<uint32_t*>(arena_alloc(W, 4ull * (c.u_meta + 1)));
                if (!runval_s) { set_status(status, c.blk, ST_ARENA); skip = 1; }
            }
            skip_s = skip;
        }
        is_rle[tid] = 0;
        __syncthreads();
        if (skip_s) continue;
        const uint8_t* meta = c.meta;
        const uint32_t nrs = meta[0] ? meta[0] : 256u;
        if (tid < nrs) is_rle[meta[1 + tid]] = 1;
        uint32_t* runval = runval_s;
        __syncthreads();
        const uint8_t* run = meta + 1 + nrs;
        const uint32_t run_len = c.u_meta - (1 + nrs);

        // pass A: a varint ends at a byte without the continuation bit (or at the last byte)
        uint32_t nvar = 0;
        for (uint32_t t0 = 0; t0 < run_len; t0 += RLE_T) {
            uint32_t i = t0 + tid;
            uint32_t term = 0;
            if (i < run_len) term = (!(run[i] & 0x80) || i + 1 == run_len) ? 1u : 0u;
            uint32_t tot, idx = block_exscan<uint32_t>(term, wtot, &tot);
            if (term) {                                      // walk back to the start of this varint
                uint32_t s = i;
                while (s > 0 && (run[s - 1] & 0x80)) s--;
                uint32_t v = 0;
                for (uint32_t k = s; k <= i; k++) v = (v << 7) | (run[k] & 0x7f);
                runval[nvar + idx] = v;
            }
            nvar += tot;
        }
        __syncthreads();

        // pass B: per tile of RLE_T literals, output offsets by a CTA scan, then the tile's output
        // is produced 16 aligned bytes per thread (binary search for the chunk's first literal,
        // then a walk), so the stores are full 128-bit lines instead of one byte loop per run.
        const uint8_t* lit = c.t1;
        const uint32_t nlit = c.t1_size;
        uint8_t* out = c.t2;
        unsigned long long opos = 0;
        uint32_t kbase = 0;
        int overflow = 0;
        for (uint32_t t0 = 0; t0 < nlit; t0 += RLE_T) {
            uint32_t i = t0 + tid;
            uint32_t b = 0, isr = 0;
            if (i < nlit) { b = lit[i]; isr = is_rle[b]; }
            uint32_t tot, k = block_exscan<uint32_t>(isr, wtot, &tot);
            unsigned long long len = 0;
            if (i < nlit) {
                len = 1;
                if (isr) { uint32_t kk = kbase + k; len += (kk < nvar) ? runval[kk] : 0u; }
            }
            unsigned long long ltot, off = block_exscan<unsigned long long>(len, wtot64, &ltot);
            if (opos + ltot > c.osz) { overflow = 1; break; }            // rle.c:161,172 (CTA-uniform)
            s_off[tid] = (uint32_t)off;                                  // tile-relative; ltot <= osz < 2^31
            s_lit[tid] = (uint8_t)b;
            if (tid == 0) s_off[RLE_T] = (uint32_t)ltot;
            __syncthreads();
            const uint32_t nt = min((uint32_t)RLE_T, nlit - t0);         // literals in this tile
            uint8_t* tout = out + opos;
            const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(tout) & 15);
            // chunk m covers tile-relative bytes [16 m - skew, 16 m - skew + 16)
            const uint32_t nchunk = (uint32_t)((ltot + skew + 15) / 16);
            for (uint32_t m = tid; m < nchunk; m += RLE_T) {
                const int64_t c0 = (int64_t)16 * m - skew;
                const uint32_t lo = c0 < 0 ? 0u : (uint32_t)c0;
                const uint32_t hi = (uint32_t)min((unsigned long long)(c0 + 16), ltot);
                // last literal whose offset is <= lo
                uint32_t a = 0, z = nt;
                while (z - a > 1) { const uint32_t mid = (a + z) >> 1; if (s_off[mid] <= lo) a = mid; else z = mid; }
                uint32_t j = a, nxt = (j + 1 < nt) ? s_off[j + 1] : (uint32_t)ltot;
                unsigned long long wl = 0, wh = 0;                       // the chunk's 16 bytes
                uint32_t cur = s_lit[j];
                auto ones = [](uint32_t k) { return k >= 8 ? ~0ull : ((1ull << (8 * k)) - 1ull); };   // k low bytes set
                for (uint32_t p = lo; p < hi;) {
                    while (p >= nxt) { j++; cur = s_lit[j]; nxt = (j + 1 < nt) ? s_off[j + 1] : (uint32_t)ltot; }
                    const uint32_t e = min(nxt, hi);
                    const uint32_t q0 = (uint32_t)(p - c0), q1 = (uint32_t)(e - c0);     // 0 <= q0 < q1 <= 16
                    const unsigned long long v = cur * 0x0101010101010101ull;
                    wl |= v & (ones(min(q1, 8u)) & ~ones(min(q0, 8u)));
                    wh |= v & (ones(max(q1, 8u) - 8) & ~ones(max(q0, 8u) - 8));
                    p = e;
                }
                if (hi - lo == 16) {
                    *reinterpret_cast<uint4*>(tout + c0) = make_uint4((uint32_t)wl, (uint32_t)(wl >> 32), (uint32_t)wh, (uint32_t)(wh >> 32));
                } else {
                    for (uint32_t p = lo; p < hi; p++) {
                        const uint32_t q = (uint32_t)(p - c0);
                        tout[p] = (uint8_t)(q < 8 ? (wl >> (8 * q)) : (wh >> (8 * (q - 8))));
                    }
                }
            }
            __syncthreads();
            opos += ltot;
            kbase += tot;
        }
        overflow = __syncthreads_or(overflow);
        if (tid == 0) {
            if (overflow || opos > c.osz) set_status(status, c.blk, ST_FORMAT);
            else {
                W->chains[ci].t2_size = (uint32_t)opos;
                if (!(c.flags & F_PACK)) {
                    W->chains[ci].final_size = (uint32_t)opos;
                    if (c.expect == 0xffffffffu) out_len[c.blk] = (uint32_t)opos;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// un-PACK, pack.c:211-348.  One CTA per chain.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpack_kernel(DecWork* W, int32_t* status, uint32_t* out_len) {
    __shared__ uint8_t map[16];
    __shared__ int skip_s;
    for (uint32_t li = blockIdx.x; li < W->nunpack; li += gridDim.x) {
        const uint32_t ci = W->unpack_list[li];
        __syncthreads();
        const Chain c = W->chains[ci];
        if (threadIdx.x == 0) skip_s = status[c.blk] != ST_OK;
        if (threadIdx.x < 16) map[threadIdx.x] = c.map[threadIdx.x];
        __syncthreads();
        if (skip_s) continue;
        const uint8_t* in = c.t2;
        const uint32_t len = c.t2_size;
        uint8_t* out = c.t3;
        const uint32_t olen = (c.per == 1) ? len : c.osz;    // rANS_static4x16pr.c:1616-1617
        bool bad = false;
        if (c.per == 1) {
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) out[i] = in[i];
        } else if (c.per == 0) {
            for (uint32_t i = threadIdx.x; i < olen; i += blockDim.x) out[i] = map[0];
        } else {
            const uint32_t per = c.per, bits = 8 / per, mask = (1u << bits) - 1u;
            if (((uint64_t)olen + per - 1) / per > len) bad = true;              // pack.c:238,279,314
            else {
                const uint32_t nin = (olen + per - 1) / per; // each thread expands whole input bytes
                for (uint32_t j = threadIdx.x; j < nin; j += blockDim.x) {
                    uint32_t v = in[j];
                    uint32_t o = j * per;
                    for (uint32_t k = 0; k < per && o + k < olen; k++) out[o + k] = map[(v >> (k * bits)) & mask];
                }
            }
        }
        if (threadIdx.x == 0) {
            if (bad) set_status(status, c.blk, ST_FORMAT);
            else {
                W->chains[ci].final_size = olen;
                if (c.expect == 0xffffffffu) out_len[c.blk] = olen;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// unstripe, utils.h:41-73.  One CTA per striped block.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unstripe_kernel(DecWork* W, int32_t* status) {
    __shared__ uint32_t at[256];
    __shared__ int bad;
    for (uint32_t si = blockIdx.x; si < W->nstripe; si += gridDim.x) {
        const StripeOp op = W->stripes[si];
        __syncthreads();
        if (threadIdx.x == 0) {
            bad = status[op.blk] != ST_OK ? 2 : 0;
            uint32_t a = 0;
            for (uint32_t j = 0; j < op.N; j++) { at[j] = a; a += op.ulen / op.N + ((op.ulen % op.N) > j); }
        }
        __syncthreads();
        if (threadIdx.x < op.N && bad == 0) {
            const Chain& c = W->chains[op.chain0 + threadIdx.x];
            if (c.final_size != c.expect) bad = 1;           // rANS_static4x16pr.c:1419-1420
        }
        __syncthreads();
        if (bad) { if (threadIdx.x == 0 && bad == 1) set_status(status, op.blk, ST_FORMAT); continue; }
        const uint32_t N = op.N, ulen = op.ulen;
        const bool fast4 = N == 4 && (ulen & 15) == 0 && (reinterpret_cast<uintptr_t>(op.out) & 15) == 0 &&
                           (reinterpret_cast<uintptr_t>(op.parts) & 3) == 0;
        if (fast4) {
            // one 32-bit word from each of the four sub-streams -> 16 interleaved output bytes
            const uint32_t q = ulen / 4;
            for (uint32_t t = threadIdx.x; t < ulen / 16; t += blockDim.x) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = *reinterpret_cast<const uint32_t*>(op.parts + (size_t)j * q + 4 * (size_t)t);
                uint4 v;                                     // output word e = bytes e of w[0..3]
                v.x = __byte_perm(__byte_perm(w[0], w[1], 0x0040), __byte_perm(w[2], w[3], 0x0040), 0x5410);
                v.y = __byte_perm(__byte_perm(w[0], w[1], 0x0051), __byte_perm(w[2], w[3], 0x0051), 0x5410);
                v.z = __byte_perm(__byte_perm(w[0], w[1], 0x0062), __byte_perm(w[2], w[3], 0x0062), 0x5410);
                v.w = __byte_perm(__byte_perm(w[0], w[1], 0x0073), __byte_perm(w[2], w[3], 0x0073), 0x5410);
                reinterpret_cast<uint4*>(op.out)[t] = v;
            }
        } else {
            for (uint32_t i = threadIdx.x; i < ulen; i += blockDim.x) op.out[i] = op.parts[at[i % N] + i / N];
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int g_cap[JK_NKINDS];        // resident CTAs per SM of each persistent kernel at its own shared-memory size
static int g_smem[JK_NKINDS];
static int g_sms = 0;
constexpr int SM_SMEM = 233472, CTA_RESERVE = 1024, MAX_DYN = 232448;

template <typename K>
static void persistent_setup(uint32_t kind, K kernel, int smem, int threads) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    g_cap[kind] = per_sm < 1 ? 1 : per_sm;
    g_smem[kind] = smem;
}

// The per-batch header travels as a kernel argument (copied at launch), so the host may reuse
// its copy immediately even when several batches are in flight on the stream.
__global__ void work_init_kernel(DecWork* dst, DecWork hdr) { *dst = hdr; }

int SideStreams::init() {
    if (fork) return 0;
    if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess) return -1;
    for (int i = 0; i < N; i++)
        if (cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming) != cudaSuccess) return -1;
    return 0;
}
void SideStreams::release() {
    for (int i = 0; i < N; i++) {
        if (s[i]) cudaStreamDestroy(s[i]);
        if (join[i]) cudaEventDestroy(join[i]);
        s[i] = nullptr; join[i] = nullptr;
    }
    if (fork) cudaEventDestroy(fork);
    fork = nullptr;
}

int decode_init(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    g_sms = prop.multiProcessorCount;
    persistent_setup(JK_O0_32, dec_o0_kernel<32, false>, O0Smem<32>::TOTAL, 32);
    persistent_setup(JK_O0_4,  dec_o0_kernel<4, false>,  O0Smem<4>::TOTAL, 32);
    persistent_setup(JK_R8_O0, dec_o0_kernel<4, true>,   O0Smem<4>::TOTAL, 32);
    persistent_setup(JK_O0_4C,  dec_o0c_kernel<false>, O0CSmem::TOTAL, 32);
    persistent_setup(JK_R8_O0C, dec_o0c_kernel<true>,  O0CSmem::TOTAL, 32);
    persistent_setup(JK_O1_32,  dec_o1_kernel<32, false, 0>, O1Smem<32>::TOTAL, 32);
    persistent_setup(JK_O1_32S, dec_o1_kernel<32, false, 1>,  O1Smem<32, 1>::TOTAL, 32);
    persistent_setup(JK_O1_4,   dec_o1_kernel<4, false, 0>,  O1Smem<4>::TOTAL, 32);
    persistent_setup(JK_O0_4R8,   dec_o0r_kernel<8, false>,  RegSmem0::TOTAL, 32);
    persistent_setup(JK_O0_4R16,  dec_o0r_kernel<16, false>, RegSmem0::TOTAL, 32);
    persistent_setup(JK_R8_O0R8,  dec_o0r_kernel<8, true>,   RegSmem0::TOTAL, 32);
    persistent_setup(JK_R8_O0R16, dec_o0r_kernel<16, true>,  RegSmem0::TOTAL, 32);
    persistent_setup(JK_O1_4R8,   dec_o1r_kernel<8, false>,  O1Smem<4, 3>::TOTAL, 32);
    persistent_setup(JK_O1_4R16,  dec_o1r_kernel<16, false>, O1Smem<4, 4>::TOTAL, 32);
    persistent_setup(JK_R8_O1R8,  dec_o1r_kernel<8, true>,   O1Smem<4, 3>::TOTAL, 32);
    persistent_setup(JK_R8_O1R16, dec_o1r_kernel<16, true>,  O1Smem<4, 4>::TOTAL, 32);
    persistent_setup(JK_R8_O1,  dec_o1_kernel<4, true, 0>,   O1Smem<4>::TOTAL, 32);
    persistent_setup(JK_O1_4M,  dec_o1_kernel<4, false, 2>,  O1Smem<4, 2>::TOTAL, 32);
    persistent_setup(JK_R8_O1M, dec_o1_kernel<4, true, 2>,   O1Smem<4, 2>::TOTAL, 32);
    persistent_setup(JK_TAB,    dec_o0_kernel<4, false>,  O0Smem<4>::TOTAL, 32);
    persistent_setup(JK_O0_4P,  dec_o0_kernel<4, false, true>,  O0Smem<4, true>::TOTAL, 32);
    persistent_setup(JK_O0_32P, dec_o0_kernel<32, false, true>, O0Smem<32, true>::TOTAL, 32);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// A persistent kernel gives one warp (CTA) to a group of G streams, and the CTA scheduler fills one
// SM to capacity before it moves to the next: a batch smaller than one full wave would crowd a few
// SMs and leave the rest idle.  So the launch is shaped: with `groups` work items expected, the
// dynamic shared-memory request is padded until exactly c = ceil(groups / SMs) CTAs fit per SM and
// the grid is SMs x c -- every SM then holds the same number of streams.
// (Spreading by gating instead -- every CTA slot launched, CTAs beyond an SM's fair share exit at once, kernels keep
// their natural footprint so that chunks and kinds could share SMs -- was built and measured in round 2: 4-way decode
// at 4096 blocks 177 vs 211 GB/s order 0, 146 vs 204 order 1, and the host-buffer pipeline no faster.  Padding stays.)
struct Shape { int grid, smem; };
static Shape shaped_launch(uint32_t kind, uint32_t groups) {
    int c = (int)((groups + g_sms - 1) / g_sms);
    c = std::max(1, std::min(c, g_cap[kind]));
    // experiments: HTSCODECS_B200_CAP_O0_32=<n> limits the X_32 order-0 kernel to n resident warps per SM
    static const int cap32 = getenv("HTSCODECS_B200_CAP_O0_32") ? atoi(getenv("HTSCODECS_B200_CAP_O0_32")) : 0;
    if (kind == JK_O0_32 && cap32 > 0) c = std::min(c, cap32);
    // Never pad below 4 CTAs per SM (HTSCODECS_B200_MIN_C overrides): four one-warp CTAs still have a scheduler each, and a
    // small launch then leaves whole SMs to the kernels of the host pipeline's other chunks (measured: end-to-end encode
    // +2-3 GB/s on every stream type, decode unchanged)
    static const int min_c = getenv("HTSCODECS_B200_MIN_C") ? atoi(getenv("HTSCODECS_B200_MIN_C")) : 4;
    int grid = g_sms * c;
    if (c < min_c) { c = std::min(min_c, g_cap[kind]); grid = std::min(g_sms * c, (int)std::max(groups, 1u)); }
    int smem = g_smem[kind];
    if (c < g_cap[kind]) smem = std::max(smem, std::min(MAX_DYN, (SM_SMEM / c - CTA_RESERVE) & ~127));
    return Shape{grid, smem};
}


// Enqueue the whole decode pipeline for one batch.  Returns the number of kernels launched.
int decode_launch(const DecodeBatch& b, cudaStream_t st) {
    PlanArgs A;
    A.W = b.work; A.in_base = b.in_base; A.in_off = b.in_off; A.in_len = b.in_len;
    A.out_base = b.out_base; A.out_off = b.out_off; A.out_len = b.out_len;
    A.status = b.status; A.method = b.method; A.nblk = b.nblk;
    int launches = 0;
    auto want = [&](uint32_t k) { return (b.kinds >> k) & 1u; };
    // expected work items per kind: the host only knows the block count (the planner decides kinds on
    // the device), which is exact for the common single-kind batch and an upper bound otherwise
    auto shape = [&](uint32_t k, uint32_t per_cta) { return shaped_launch(k, (b.nblk + per_cta - 1) / per_cta); };
    work_init_kernel<<<1, 1, 0, st>>>(b.work, *b.hdr); launches++;
    plan_kernel<<<(b.nblk + 127) / 128, 128, 0, st>>>(A); launches++;
    Shape sh;
    // compressed order-1 tables first (tiny jobs), then one stream per kind
    if (want(JK_TAB)) { sh = shape(JK_TAB, 8); dec_o0_kernel<4, false><<<sh.grid, 32, sh.smem, st>>>(b.work, b.status, JK_TAB); launches++; }
    static const bool side_on = !(getenv("HTSCODECS_B200_SIDE") && atoi(getenv("HTSCODECS_B200_SIDE")) == 0);
    SideStreams* side = side_on ? b.side : nullptr;
    int nside = 0;
    if (side) cudaEventRecord(side->fork, st);
    cudaStream_t ks = st;
    // b.hot: the kinds the context's previous batch had jobs for (all bits set: not known).  Kinds outside it are
    // launched all the same -- the planner decides on the device -- but on ONE shared "cold" stream beside the caller's,
    // so that a single-kind batch does not pay two dozen empty launches one after the other; a single hot kind runs on
    // the caller's stream (measured 13 % faster there than from a side stream), several get a side stream each.
    const uint32_t hot = b.hot & b.kinds;
    const uint32_t hot_e = hot & ~((1u << JK_COPY) | (1u << JK_TAB));          // entropy kinds among them
    const bool one_hot = hot_e != 0 && (hot_e & (hot_e - 1)) == 0 && b.hot != ~0u;
    bool cold_used = false;
    cudaStream_t cold = side ? side->s[SideStreams::N - 1] : st;
#define LAUNCH_DEC(K, KERNEL, PER, ON_MAIN)                                                    \
    if (want(K)) {                                                                             \
        const bool is_hot = (hot >> K) & 1u;                                                   \
        const bool on_main = !side || (is_hot && (ON_MAIN || one_hot));                        \
        const bool on_cold = !on_main && !is_hot;                                              \
        if (on_main) ks = st;                                                                  \
        else if (on_cold) { ks = cold; if (!cold_used) { cudaStreamWaitEvent(cold, side->fork, 0); cold_used = true; } } \
        else { ks = side->s[nside]; cudaStreamWaitEvent(ks, side->fork, 0); }                  \
        sh = shape(K, PER); KERNEL<<<sh.grid, 32, sh.smem, ks>>>(b.work, b.status, K); launches++; \
        if (!on_main && !on_cold) { cudaEventRecord(side->join[nside], ks); nside++; }         \
    }
    // the long-latency kinds go first so that they start on an empty machine
    LAUNCH_DEC(JK_O1_4M, (dec_o1_kernel<4, false, 2>), 8, false)
    LAUNCH_DEC(JK_R8_O1M, (dec_o1_kernel<4, true, 2>), 8, false)
    LAUNCH_DEC(JK_O1_4, (dec_o1_kernel<4, false, 0>), 8, false)
    LAUNCH_DEC(JK_R8_O1, (dec_o1_kernel<4, true, 0>), 8, false)
    LAUNCH_DEC(JK_O0_4, (dec_o0_kernel<4, false>), 8, false)
    LAUNCH_DEC(JK_R8_O0, (dec_o0_kernel<4, true>), 8, false)
    // (fetching the renormalisation bytes a step AHEAD of their use was tried and measured no faster at 4096 blocks
    //  -- 211 vs 215 GB/s order 0, 187 vs 205 order 1 -- and 15 % slower at 16384: the window load is not what bounds a step)
    LAUNCH_DEC(JK_O1_4R16, (dec_o1r_kernel<16, false>), 8, false)
    LAUNCH_DEC(JK_R8_O1R16, (dec_o1r_kernel<16, true>), 8, false)
    LAUNCH_DEC(JK_O1_4R8, (dec_o1r_kernel<8, false>), 8, false)
    LAUNCH_DEC(JK_R8_O1R8, (dec_o1r_kernel<8, true>), 8, false)
    LAUNCH_DEC(JK_O0_4R16, (dec_o0r_kernel<16, false>), 8, false)
    LAUNCH_DEC(JK_R8_O0R16, (dec_o0r_kernel<16, true>), 8, false)
    LAUNCH_DEC(JK_O0_4R8, (dec_o0r_kernel<8, false>), 8, false)
    LAUNCH_DEC(JK_R8_O0R8, (dec_o0r_kernel<8, true>), 8, false)
    LAUNCH_DEC(JK_O0_4P, (dec_o0_kernel<4, false, true>), 8, false)
    LAUNCH_DEC(JK_O0_32P, (dec_o0_kernel<32, false, true>), 1, false)
    LAUNCH_DEC(JK_O0_4C, (dec_o0c_kernel<false>), 8, false)
    LAUNCH_DEC(JK_R8_O0C, (dec_o0c_kernel<true>), 8, false)
    LAUNCH_DEC(JK_O1_32, (dec_o1_kernel<32, false, 0>), 1, false)
    LAUNCH_DEC(JK_O1_32S, (dec_o1_kernel<32, false, 1>), 1, false)
    // the throughput-bound kind stays on the caller's stream: measured 13 % slower from a side stream (13 streams
    // share 8 hardware queues; bench.py's headline batch is all of this kind)
    LAUNCH_DEC(JK_O0_32, (dec_o0_kernel<32, false>), 1, true)
#undef LAUNCH_DEC
    if (cold_used) { cudaEventRecord(side->join[SideStreams::N - 1], cold); cudaStreamWaitEvent(st, side->join[SideStreams::N - 1], 0); }
    for (int i = 0; i < nside; i++) cudaStreamWaitEvent(st, side->join[i], 0);
    if (want(JK_COPY))  { copy_kernel<<<g_sms * 4, 256, 0, st>>>(b.work); launches++; }
    if (b.post & 1u) { rle_kernel<<<g_sms * 4, RLE_T, 0, st>>>(b.work, b.status, b.out_len); launches++; }
    if (b.post & 2u) { unpack_kernel<<<g_sms * 4, 256, 0, st>>>(b.work, b.status, b.out_len); launches++; }
    if (b.post & 4u) { unstripe_kernel<<<g_sms * 4, 256, 0, st>>>(b.work, b.status); launches++; }
    return launches;
}

}  // namespace hb
