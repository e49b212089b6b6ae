// common.cuh -- descriptors and small device helpers shared by the decode/encode kernels.
//
// Vocabulary: a *block* is one CRAM block handed to the batch API.  A *chain* is one
// non-striped rANS Nx16 container (a whole block, or one of the N sub-streams of an X_STRIPE
// block): entropy stage -> un-RLE -> un-PACK.  A *job* is one leaf entropy stream handed to a
// decode kernel (a warp, or a group of 4 lanes, owns one job).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

// container flag byte, reference rANS_static4x16pr.c:39-43 (+ X_32)
enum : uint32_t {
    F_ORDER1 = 0x01, F_X32 = 0x04, F_STRIPE = 0x08, F_NOSZ = 0x10,
    F_CAT = 0x20, F_RLE = 0x40, F_PACK = 0x80,
};

enum : int32_t {
    ST_OK = 0, ST_FORMAT = -1, ST_SIZE = -2, ST_ARENA = -3, ST_NESTED = -4, ST_INTERNAL = -5,
};

// decode job kinds: one persistent kernel instantiation per kind
enum JobKind : uint32_t {
    JK_O0_4 = 0, JK_O0_32, JK_O1_4, JK_O1_32, JK_R8_O0, JK_R8_O1, JK_COPY,
    JK_O1_32S,             // X_32 order-1 streams with a small alphabet: the high-occupancy kernel variant
    JK_O0_4C, JK_R8_O0C,   // 4-way / 4x8 order-0 streams on compact tables: 256 resident streams per SM, used
                           // for batches too large for one wave of the 4 KB-LUT kernels (17-64 symbols)
    JK_O1_4M, JK_R8_O1M,   // 4-way / 4x8 order-1 streams of 20-47 symbols (q40-style qualities): 12.5 KB of compact
                           // tables per stream in shared memory, 16 resident streams per SM
    JK_TAB,                // order-0 (4-way) jobs that expand an order-1 stream's compressed table: they run before
                           // every other kind, which then run side by side
    // 4-way / 4x8 streams of up to 8 / 16 symbols: table (order 0) or context row (order 1) in registers, the
    // renormalisation bytes prefetched as a window (dec_o0r_kernel / dec_o1r_kernel) -- the short-latency variants
    JK_O0_4R8, JK_O0_4R16, JK_O1_4R8, JK_O1_4R16, JK_R8_O0R8, JK_R8_O0R16, JK_R8_O1R8, JK_R8_O1R16,
    JK_O0_4P, JK_O0_32P,   // order-0 streams of an X_PACK container (no X_RLE): un-PACK is the decoder's sink (pack.c:211-348)
    JK_NKINDS
};

struct DecJob {
    const uint8_t* in;
    uint8_t* out;
    uint32_t in_len;
    uint32_t out_len;
    uint32_t blk;
    uint32_t fuse;         // != 0: un-PACK fused into the decoder's sink -- every decoded byte expands to `fuse` (2, 4, 8)
                           // symbols through `map`, written straight to `out` (fin_len bytes in all)
    uint8_t* aux;          // order-1 with an O0-compressed table: scratch for the expanded table
    uint32_t fin_len;
    uint32_t pad;
    uint8_t map[16];
};

struct Chain {
    uint8_t* t1;           // entropy-stage output (literals / packed bytes / final)
    uint8_t* t2;           // un-RLE output
    uint8_t* t3;           // un-PACK output (always the caller's buffer)
    const uint8_t* meta;   // RLE meta: [nsyms][syms..][varint run lengths..]
    uint32_t blk;
    uint32_t flags;        // F_RLE / F_PACK bits of this chain
    uint32_t u_meta;       // RLE meta size
    uint32_t t1_size;
    uint32_t osz;          // declared uncompressed size = capacity of t2 / t3
    uint32_t t2_size;      // set by the planner, overwritten by the RLE kernel
    uint32_t final_size;   // set by whichever stage runs last
    uint32_t expect;       // stripe sub-stream: required final size; 0xffffffff for a top-level chain
    uint32_t per;          // PACK: symbols per byte (0,1,2,4,8)
    uint8_t  map[16];      // PACK: code -> symbol
};

struct StripeOp {
    const uint8_t* parts;  // N decoded sub-streams, concatenated
    uint8_t* out;
    uint32_t blk;
    uint32_t ulen;
    uint32_t N;
    uint32_t chain0;       // first of the N chains
};

// Per-batch device work area.  The header (everything up to `arena`) is zeroed and then
// patched by the host before each batch; lists live in one device allocation.
struct DecWork {
    // counters (device-written)
    uint32_t njobs[JK_NKINDS];
    uint32_t next[JK_NKINDS];
    uint32_t nchains, nrle, nunpack, nstripe;
    uint32_t next_rle, next_unpack, next_stripe;
    uint32_t overflow;                 // a list capacity was exceeded
    unsigned long long arena_used;     // bytes requested from the arena (may exceed capacity)
    // capacities and pointers (host-written)
    uint32_t job_cap, chain_cap, stripe_cap;
    uint32_t big_batch;                // host hint: more 4-way streams than the LUT kernels hold in one wave
    uint32_t kinds;                    // job kinds whose kernels this batch launches: a job of any other kind is refused
    unsigned long long arena_cap;
    DecJob* jobs[JK_NKINDS];
    Chain* chains;
    uint32_t* rle_list;
    uint32_t* unpack_list;
    StripeOp* stripes;
    uint8_t* arena;
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
    uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t a, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 7-bit big-endian varint, reference varint.h:131-160 (bounded form).  Works on any address space.
__device__ __forceinline__ int var_get_u32(const uint8_t* p, const uint8_t* end, uint32_t* v) {
    uint32_t x = 0;
    int n = 0;
    *v = 0;
    if (p >= end) return 0;
    for (;;) {
        uint32_t c = p[n++];
        x = (x << 7) | (c & 0x7f);
        if (!(c & 0x80) || p + n >= end) break;
    }
    *v = x;
    return n;
}

// reference varint.h:85-104
__device__ __forceinline__ int var_put_u32(uint8_t* p, uint32_t v) {
    int n = 1;
    for (uint32_t t = v >> 7; t; t >>= 7) n++;
    for (int k = n - 1; k >= 0; k--) *p++ = (uint8_t)(((v >> (7 * k)) & 0x7f) | (k ? 0x80 : 0));
    return n;
}
__device__ __forceinline__ int var_len_u32(uint32_t v) {
    int n = 1;
    for (uint32_t t = v >> 7; t; t >>= 7) n++;
    return n;
}

__device__ __forceinline__ uint32_t ld_u32_le(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Record the first error of a block (0 = OK is the largest value, errors are negative).
__device__ __forceinline__ void set_status(int32_t* status, uint32_t blk, int32_t code) {
    atomicMin(&status[blk], code);
}

// One side stream per job kind (decode) or kernel variant (encode): the entropy kernels of a batch that mixes
// codecs run side by side (each is latency-bound -- one warp per stream or per eight 4-way streams -- so a mixed
// batch then costs its slowest kind, not the sum over kinds).  Every kind's launch is shaped for the whole batch
// (c CTAs per SM), so the kinds settle on disjoint groups of SMs, c CTAs each.  Owned by the caller, reused across batches.
struct SideStreams {
    static constexpr int N = 24;
    cudaStream_t s[N] = {};
    cudaEvent_t fork = nullptr, join[N] = {};
    int init();
    void release();
};

}  // namespace hb
