// api.cu -- the C ABI of libhtscodecs_b200.so (see include/htscodecs_b200.h).
//
// Host code does no codec work: it sizes work lists and arenas, moves bytes between host and
// device, enqueues the kernels of decode.cu / encode.cu and gathers per-block sizes and status.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/htscodecs_b200.h"
#include "decode.h"
#include "encode.h"

using namespace hb;

#define CK(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s", __FILE__, __LINE__,           \
                     cudaGetErrorString(e_));                                              \
            return -1;                                                                     \
        }                                                                                  \
    } while (0)

namespace {

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    int ensure(size_t n) {
        if (n <= cap) return 0;
        size_t want = std::max(n, cap + cap / 2);
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <typename T> struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        size_t want = std::max(n, cap + cap / 2);
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        if (cudaHostAlloc(&p, want * sizeof(T), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Device-side lists + arena for one in-flight decode batch.
struct DecSlot {
    DevBuf<uint8_t> work;       // DecWork header
    DevBuf<uint8_t> lists;      // jobs, chains, index lists, stripe ops
    DevBuf<uint8_t> arena;
    DevBuf<uint32_t> cap_save;  // copy of the caller's capacities, for retries
    PinBuf<DecWork> h_work;     // header as sent / as read back
    uint32_t job_cap = 0, chain_cap = 0, stripe_cap = 0;
    // Did the previous batch use more than one entropy kernel kind?  (Read from its header once that has arrived.)
    // A single-kind batch predicts another one, which is launched on the caller's stream alone -- measured 13 %
    // faster than from a side stream; -1 = not known yet.
    cudaEvent_t hdr_ready = nullptr; bool hdr_pending = false; int mixed = -1; uint32_t hot = ~0u;
    void release() { work.release(); lists.release(); arena.release(); cap_save.release(); h_work.release(); if (hdr_ready) cudaEventDestroy(hdr_ready); hdr_ready = nullptr; }
};

// Staging for the host-resident API: one pipeline stage.
struct Stage {
    DevBuf<uint8_t> d_in, d_out;
    DevBuf<uint64_t> d_off;     // [in_off | out_off]
    DevBuf<uint32_t> d_u32;     // [in_len | out_len | status | order]
    DevBuf<uint8_t> d_method;
    PinBuf<uint64_t> h_off;
    PinBuf<uint32_t> h_u32;
    PinBuf<uint8_t> h_method;
    DecSlot dec;
    EncSlot enc;
    SideStreams side;           // the chunk's kernels of different kinds run side by side
    cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
    // timing twins (busy time per stage of the pipeline, reported by the multi-device call)
    cudaEvent_t t_h2d0 = nullptr, t_h2d1 = nullptr, t_k0 = nullptr, t_k1 = nullptr, t_d2h0 = nullptr, t_d2h1 = nullptr;
    cudaStream_t s_compute = nullptr;   // one per stage: a block is decoded by ONE warp, so a chunk's kernels
                                        // last as long as its slowest block; chunks must overlap on the SMs
    // what the device->host half of a chunk needs (it may be enqueued later than the first half)
    int c_a = 0, c_n = 0; int c_lay = 0; bool c_equal = false; uint64_t c_lo = 0, c_bytes = 0, c_stride = 0;
    uint64_t c_w = 0;           // encode: leading bytes of every block already fetched (0 = whole regions)
    int init() {
        if (s_compute) return 0;
        if (cudaEventCreateWithFlags(&h2d_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&compute_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&d2h_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreate(&t_h2d0) != cudaSuccess || cudaEventCreate(&t_h2d1) != cudaSuccess ||
            cudaEventCreate(&t_k0) != cudaSuccess || cudaEventCreate(&t_k1) != cudaSuccess ||
            cudaEventCreate(&t_d2h0) != cudaSuccess || cudaEventCreate(&t_d2h1) != cudaSuccess ||
            cudaStreamCreateWithFlags(&s_compute, cudaStreamNonBlocking) != cudaSuccess) return -1;
        return 0;
    }
    void release() {
        if (s_compute) cudaStreamDestroy(s_compute);
        d_in.release(); d_out.release(); d_off.release(); d_u32.release(); d_method.release();
        h_off.release(); h_u32.release(); h_method.release(); dec.release(); enc.release(); side.release();
        if (h2d_done) cudaEventDestroy(h2d_done);
        if (compute_done) cudaEventDestroy(compute_done);
        if (d2h_done) cudaEventDestroy(d2h_done);
        for (cudaEvent_t e : {t_h2d0, t_h2d1, t_k0, t_k1, t_d2h0, t_d2h1}) if (e) cudaEventDestroy(e);
    }
};

constexpr int NSTAGE = 4;   // chunks in flight: a chunk's copy-in waits for the chunk NSTAGE before it to be collected

// job-kind sets for the host's "which kernels can be needed" hint
constexpr uint32_t K_R8_O0 = (1u << JK_R8_O0) | (1u << JK_R8_O0R8) | (1u << JK_R8_O0R16);
constexpr uint32_t K_R8_O1 = (1u << JK_R8_O1) | (1u << JK_R8_O1M) | (1u << JK_R8_O1R8) | (1u << JK_R8_O1R16);
constexpr uint32_t K_R8 = K_R8_O0 | K_R8_O1 | (1u << JK_R8_O0C);
constexpr uint32_t K_BIG = (1u << JK_O0_4C) | (1u << JK_R8_O0C);   // large batches only
constexpr uint32_t K_O0_4 = (1u << JK_O0_4) | (1u << JK_O0_4R8) | (1u << JK_O0_4R16);
constexpr uint32_t K_O1_4 = (1u << JK_O1_4) | (1u << JK_O1_4M) | (1u << JK_O1_4R8) | (1u << JK_O1_4R16);

}  // namespace

struct hts_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // compute (and the whole device-resident API)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    DecSlot dec;                        // device-resident API
    EncSlot enc;
    SideStreams side_dec;
    std::vector<Stage> stage;           // NSTAGE slots reused round-robin (full duplex) or one per chunk (half duplex)
    cudaEvent_t all_h2d = nullptr;
    bool full_duplex = true;            // overlap host->device with device->host copies (see hts_b200_set_copy_duplex)
    PinBuf<uint8_t> pin_in, pin_out;    // pointer-array wrappers
    DevBuf<uint8_t> best_in, best_cand, best_out;   // rans4x16_compress_best_batch: inputs, candidates, winners
    DevBuf<uint64_t> best_off;          // [in_off | cand_off] per candidate, then [src_off | dst_off] per winner
    DevBuf<uint32_t> best_u32;          // [in_len | cand_len | status | order] per candidate
    PinBuf<uint64_t> best_hoff;
    PinBuf<uint32_t> best_hu32;
    size_t arena_hint = 0;
    double enc_frac = 0;                // encode: largest stream / capacity ratio seen lately (0 = none yet), see launch_out
    unsigned long long launches = 0;
    char err[256] = {0};
};

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
namespace {
struct DeviceGuard {            // creating a context on another device must not move the caller's current device
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" hts_b200_ctx* hts_b200_create(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return nullptr; }
    DeviceGuard guard;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return nullptr;
    if (device >= ndev || cudaSetDevice(device) != cudaSuccess) return nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return nullptr;
    if (prop.major < 10) {
        fprintf(stderr, "htscodecs_b200: device %d is sm_%d%d; this library is built for sm_100a only\n",
                device, prop.major, prop.minor);
        return nullptr;
    }
    hts_b200_ctx* ctx = new hts_b200_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) != cudaSuccess ||
        decode_init(device) != 0 || encode_init(device) != 0) {
        fprintf(stderr, "htscodecs_b200: initialisation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return nullptr;
    }
    ctx->stage.resize(NSTAGE);
    for (int s = 0; s < NSTAGE; s++)
        if (ctx->stage[s].init()) { hts_b200_destroy(ctx); return nullptr; }
    cudaEventCreateWithFlags(&ctx->all_h2d, cudaEventDisableTiming);
    if (const char* e = getenv("HTSCODECS_B200_COPY_DUPLEX")) ctx->full_duplex = strcmp(e, "half") != 0;
    return ctx;
}

extern "C" void hts_b200_destroy(hts_b200_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    ctx->dec.release();
    ctx->enc.release();
    ctx->side_dec.release();
    for (auto& st : ctx->stage) st.release();
    if (ctx->all_h2d) cudaEventDestroy(ctx->all_h2d);
    ctx->pin_in.release(); ctx->pin_out.release();
    ctx->best_in.release(); ctx->best_cand.release(); ctx->best_out.release(); ctx->best_off.release();
    ctx->best_u32.release(); ctx->best_hoff.release(); ctx->best_hu32.release();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    delete ctx;
}

extern "C" const char* hts_b200_last_error(const hts_b200_ctx* ctx) { return ctx ? ctx->err : "no context"; }
extern "C" unsigned long long hts_b200_launch_count(const hts_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* hts_b200_stream(const hts_b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" void hts_b200_set_copy_duplex(hts_b200_ctx* ctx, int full) { if (ctx) ctx->full_duplex = full != 0; }
extern "C" size_t hts_b200_scratch_bytes(const hts_b200_ctx* ctx) {
    if (!ctx) return 0;
    auto dec = [](const DecSlot& d) { return d.work.cap + d.lists.cap + d.arena.cap + 4 * d.cap_save.cap; };
    size_t n = dec(ctx->dec) + encode_device_bytes(ctx->enc);
    for (const auto& S : ctx->stage) n += dec(S.dec) + encode_device_bytes(S.enc) + S.d_in.cap + S.d_out.cap + 8 * S.d_off.cap + 4 * S.d_u32.cap + S.d_method.cap;
    return n + ctx->best_in.cap + ctx->best_cand.cap + ctx->best_out.cap + 8 * ctx->best_off.cap + 4 * ctx->best_u32.cap;
}
extern "C" void* hts_b200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }   // pinned for every device
    return p;
}
extern "C" void hts_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------------
// decode: device-resident
// ------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Lay the lists out inside slot.lists and fill the host copy of the header.
static int dec_prepare(hts_b200_ctx* ctx, DecSlot& s, int nblk, size_t arena_bytes) {
    uint32_t job_cap = std::max<uint32_t>(s.job_cap, 2u * nblk + 1024u);
    uint32_t chain_cap = std::max<uint32_t>(s.chain_cap, (uint32_t)nblk + 1024u);
    uint32_t stripe_cap = std::max<uint32_t>(s.stripe_cap, (uint32_t)nblk + 16u);
    size_t off = 0, o_jobs[JK_NKINDS];
    for (int k = 0; k < JK_NKINDS; k++) { o_jobs[k] = off; off += align_up(sizeof(DecJob) * job_cap, 256); }
    size_t o_chains = off; off += align_up(sizeof(Chain) * chain_cap, 256);
    size_t o_rle = off; off += align_up(4 * (size_t)chain_cap, 256);
    size_t o_unp = off; off += align_up(4 * (size_t)chain_cap, 256);
    size_t o_str = off; off += align_up(sizeof(StripeOp) * stripe_cap, 256);
    if (s.lists.ensure(off) || s.work.ensure(sizeof(DecWork)) || s.h_work.ensure(2) ||
        s.arena.ensure(std::max<size_t>(arena_bytes, 1 << 20)) || s.cap_save.ensure(nblk)) {
        snprintf(ctx->err, sizeof(ctx->err), "out of device memory (lists %zu B, arena %zu B)", off, arena_bytes);
        return -1;
    }
    s.job_cap = job_cap; s.chain_cap = chain_cap; s.stripe_cap = stripe_cap;
    DecWork& h = s.h_work.p[0];
    memset(&h, 0, sizeof(h));
    h.job_cap = job_cap; h.chain_cap = chain_cap; h.stripe_cap = stripe_cap;
    h.big_batch = 0;
    h.arena_cap = s.arena.cap;
    for (int k = 0; k < JK_NKINDS; k++) h.jobs[k] = reinterpret_cast<DecJob*>(s.lists.p + o_jobs[k]);
    h.chains = reinterpret_cast<Chain*>(s.lists.p + o_chains);
    h.rle_list = reinterpret_cast<uint32_t*>(s.lists.p + o_rle);
    h.unpack_list = reinterpret_cast<uint32_t*>(s.lists.p + o_unp);
    h.stripes = reinterpret_cast<StripeOp*>(s.lists.p + o_str);
    h.arena = s.arena.p;
    return 0;
}

// Enqueue one decode attempt on `st` (header upload, kernels, header download into h_work[1]).
static int dec_enqueue(hts_b200_ctx* ctx, DecSlot& s, const DecodeBatch& b, cudaStream_t st) {
    DecodeBatch bb = b;
    if (s.hdr_pending && cudaEventQuery(s.hdr_ready) == cudaSuccess) {
        const DecWork& r = s.h_work.p[1];
        int used = 0;
        for (int k = 0; k < JK_NKINDS; k++) used += (k != JK_COPY && k != JK_TAB && r.njobs[k] != 0);
        s.mixed = used > 1;
        uint32_t hot = 0;
        for (int k = 0; k < JK_NKINDS; k++) if (r.njobs[k] != 0) hot |= 1u << k;
        s.hot = hot;
        s.hdr_pending = false;
    }
    bb.hot = s.hot;
    bb.work = reinterpret_cast<DecWork*>(s.work.p);
    s.h_work.p[0].big_batch = b.big_batch ? 1u : 0u;
    s.h_work.p[0].kinds = b.kinds;
    bb.hdr = &s.h_work.p[0];
    ctx->launches += decode_launch(bb, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&s.h_work.p[1], s.work.p, sizeof(DecWork), cudaMemcpyDeviceToHost, st));
    if (!s.hdr_ready) CK(cudaEventCreateWithFlags(&s.hdr_ready, cudaEventDisableTiming));
    CK(cudaEventRecord(s.hdr_ready, st));                        // (a query succeeds only once the LATEST copy has landed)
    s.hdr_pending = true;
    return 0;
}

// After the stream drained: did the attempt run out of list or arena space?  If so grow.
static bool dec_needs_retry(DecSlot& s, size_t* arena_bytes) {
    const DecWork& r = s.h_work.p[1];
    bool retry = false;
    // blocks that found the arena full stopped asking, so the reported need is a lower bound: over-provision
    if (r.arena_used > r.arena_cap) { *arena_bytes = (size_t)(2 * r.arena_used + (16 << 20)); retry = true; }
    if (r.overflow) {
        uint32_t mj = 0;
        for (int k = 0; k < JK_NKINDS; k++) mj = std::max(mj, r.njobs[k]);
        s.job_cap = std::max(s.job_cap, mj + mj / 4 + 1024);
        s.chain_cap = std::max(s.chain_cap, std::max(r.nchains, std::max(r.nrle, r.nunpack)) * 5 / 4 + 1024);
        s.stripe_cap = std::max(s.stripe_cap, r.nstripe * 5 / 4 + 16);
        retry = true;
    }
    return retry;
}

extern "C" int hts_b200_uncompress_batch_dev(hts_b200_ctx* ctx, int nblk, const uint8_t* in_base,
                                             const uint64_t* in_off, const uint32_t* in_len, uint8_t* out_base,
                                             const uint64_t* out_off, uint32_t* out_len, int32_t* status,
                                             const uint8_t* method, int sync) {
    if (!ctx || nblk < 0) return -1;
    if (nblk == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    DecSlot& s = ctx->dec;
    // an asynchronous call cannot retry: what its batch header reported (arena need, list overflow) sizes this one
    if (s.hdr_pending && cudaEventQuery(s.hdr_ready) == cudaSuccess) {
        size_t ab = 0;
        if (dec_needs_retry(s, &ab)) ctx->arena_hint = std::max(ctx->arena_hint, ab);
        else ctx->arena_hint = std::max(ctx->arena_hint, (size_t)s.h_work.p[1].arena_used);
    }
    cudaGetLastError();
    size_t arena_bytes = std::max<size_t>(ctx->arena_hint, 64u << 20);
    DecodeBatch b;
    b.work = nullptr; b.hdr = nullptr; b.in_base = in_base; b.in_off = in_off; b.in_len = in_len;
    b.out_base = out_base; b.out_off = out_off; b.out_len = out_len; b.status = status; b.method = method;
    b.nblk = nblk; b.kinds = method ? ~0u : ~K_R8; b.post = 7u;
    // more 4-way streams than the LUT kernels keep resident (40 per SM): let the planner route
    // small-alphabet order-0 streams to the compact-table kernels (256 per SM)
    static const int compact_min = getenv("HTSCODECS_B200_COMPACT_MIN") ? atoi(getenv("HTSCODECS_B200_COMPACT_MIN")) : 5000;
    const bool big = nblk > compact_min;
    if (!big) b.kinds &= ~K_BIG;
    if (ctx->side_dec.init() == 0) b.side = &ctx->side_dec;
    b.big_batch = big;
    for (int attempt = 0; attempt < 8; attempt++) {
        if (dec_prepare(ctx, s, nblk, arena_bytes)) return -1;
        if (sync) {
            if (attempt == 0) CK(cudaMemcpyAsync(s.cap_save.p, out_len, 4 * (size_t)nblk, cudaMemcpyDeviceToDevice, ctx->stream));
            else CK(cudaMemcpyAsync(out_len, s.cap_save.p, 4 * (size_t)nblk, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (dec_enqueue(ctx, s, b, ctx->stream)) return -1;
        if (!sync) return 0;
        CK(cudaStreamSynchronize(ctx->stream));
        if (!dec_needs_retry(s, &arena_bytes)) { ctx->arena_hint = std::max(ctx->arena_hint, (size_t)s.h_work.p[1].arena_used); return 0; }
        ctx->arena_hint = arena_bytes;
    }
    snprintf(ctx->err, sizeof(ctx->err), "decode work area kept overflowing");
    return -1;
}

// ------------------------------------------------------------------------------------------
// host-resident batches: chunked H2D -> kernels -> D2H pipeline over NSTAGE staging slots
// ------------------------------------------------------------------------------------------
namespace {

// How the blocks [a,b) of one direction sit in the caller's buffer.
//   TILED    ascending and contiguous (off[i+1] == off[i] + len[i]): the span holds nothing but the blocks
//   STRIDED  ascending, equal lengths, evenly spaced with stride >= length (padding between the blocks)
//   LOOSE    neither, but the gaps are small (<= 64 B per block on average): fine to READ as one span
//   SCATTER  anything else
enum Layout { TILED, STRIDED, LOOSE, SCATTER };
struct Range { uint64_t lo, hi, stride; Layout lay; };

Range span_of(const uint64_t* off, const uint32_t* len, int a, int b) {
    Range r{UINT64_MAX, 0, 0, TILED};
    if (a >= b) { r.lo = r.hi = 0; return r; }
    uint64_t sum = 0;
    bool tiled = true, strided = b - a > 1;
    const uint64_t stride = b - a > 1 ? off[a + 1] - off[a] : 0;
    for (int i = a; i < b; i++) {
        r.lo = std::min(r.lo, off[i]);
        r.hi = std::max(r.hi, off[i] + len[i]);
        sum += len[i];
        if (i > a) {
            if (off[i] != off[i - 1] + len[i - 1]) tiled = false;
            if (off[i] <= off[i - 1] || off[i] - off[i - 1] != stride || len[i] != len[a]) strided = false;
        }
    }
    if (strided && (stride < len[a] || stride > 0x7fffffffull)) strided = false;
    r.stride = stride;
    if (tiled && r.lo == off[a]) r.lay = TILED;
    else if (strided) r.lay = STRIDED;
    else if (r.hi - r.lo <= sum + 64ull * (b - a)) r.lay = LOOSE;
    else r.lay = SCATTER;
    return r;
}

// A rendezvous of the per-device threads of a multi-device call (see hts_b200_*_batch_host_multi): every device
// finishes its host->device copies before any device starts copying results back.
struct PhaseSync {
    std::mutex mu;
    std::condition_variable cv;
    int want = 0, arrived = 0;
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(mu);
        if (++arrived >= want) cv.notify_all();
        else cv.wait(lk, [&] { return arrived >= want; });
    }
};

struct HostRun {                 // optional extras of one host-buffer call
    PhaseSync* phase = nullptr;  // non-null: phased copies with a cross-device rendezvous
    bool arrived = false;
    hts_b200_dev_stats* stats = nullptr;
    // multi-device calls, full-duplex policy: the chunks of the WHOLE batch and a cursor shared by the device threads --
    // a device takes the next chunk whenever one of its pipeline stages is free, so devices behind a slower host link
    // (or sharing one) simply end up with fewer chunks
    const std::vector<int>* cuts = nullptr;
    std::atomic<int>* cursor = nullptr;
    const uint32_t* caps = nullptr;      // the batch's capacities (a copy of out_len taken before any thread writes sizes)
    std::function<void(int, int)> on_chunk;   // called with (first block, count) once a chunk's sizes, status and bytes have landed
};

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// Chunk boundaries of a host-buffer call: blocks [cuts[k], cuts[k+1]) travel and run together.  Pure host logic
// (exported as hts_b200_plan_chunks for the CPU tests).
static std::vector<int> plan_chunks(bool enc, int nblk, const uint8_t* in_base, const uint64_t* in_off,
                                    const uint32_t* in_len, const uint32_t* out_len, const uint8_t* method,
                                    const int32_t* order, bool* latency_bound = nullptr) {
    // ---- chunking.  The entropy kernels give one warp (or 4 lanes) to a block, so a chunk's kernel time is
    // its slowest block's time however few blocks it holds: chunks must be big enough to fill the SMs
    // (hundreds of blocks) yet numerous enough (>= ~6) for the copies of one to hide behind the next.
    uint64_t total_bytes = 0;
    for (int i = 0; i < nblk; i++) total_bytes += (uint64_t)in_len[i] + out_len[i];
    uint64_t target = std::max<uint64_t>(32ull << 20, std::min<uint64_t>(384ull << 20, total_bytes / 8));
    {
        // A chunk's kernels cannot finish before its slowest stream does: n / lanes serial steps of ~130 ns
        // (1 MiB: 4 ms X_32, 34 ms 4-way).  Three chunks overlap on the device (NSTAGE), so a chunk should carry
        // at least that much copy time (~55 GB/s per direction), or the copy engines wait for kernels:
        // 4096 x 1 MiB 4-way blocks went from 20.6 to 38.0 GB/s end to end (order 1: 13.4 to 27.0).
        uint64_t steps = 0;
        for (int i = 0; i < nblk; i++) {
            const uint32_t u = enc ? in_len[i] : out_len[i];
            uint32_t lanes = 4, parts = 1;
            if (enc) {
                if (!(order[i] & HTS_B200_ORDER_RANS4x8)) { if (order[i] & RANS_ORDER_X32) lanes = 32; if (order[i] & RANS_ORDER_STRIPE) parts = 4; }
            } else if (!(method && method[i] == 1) && in_len[i]) {
                const uint8_t f = in_base[in_off[i]];
                if (f & F_X32) lanes = 32;
                if (f & F_STRIPE) parts = 4;
            }
            steps = std::max<uint64_t>(steps, u / lanes / parts);
        }
        static const double step_ns = getenv("HTSCODECS_B200_STEP_NS") ? atof(getenv("HTSCODECS_B200_STEP_NS")) : 130.0;   // (experiment knob)
        const uint64_t by_floor = (uint64_t)(steps * step_ns * 1e-9 * 55e9);
        if (latency_bound) *latency_bound = by_floor > target;       // chunks sized by the kernels' latency, not by the copies
        target = std::max(target, std::min<uint64_t>(std::min<uint64_t>(by_floor, 2048ull << 20), total_bytes / 3));
        if (enc && total_bytes > (64ull << 20)) {
            // Encode is bound by the host->device stream, which starts at once, and every chunk costs one round of
            // latency-bound kernels K (histogram, tables, coder, assembly; rounds of consecutive chunks do not share
            // SMs).  With n equal chunks the call takes  H/n + max((n-1) H/n, (n-1) K) + K + D/n  (H, D: copy times
            // at ~55 GB/s; compressed output ~ a third of its capacity): take the n that minimises it.
            uint64_t in_bytes = 0, cap_bytes = 0;
            for (int i = 0; i < nblk; i++) { in_bytes += in_len[i]; cap_bytes += out_len[i]; }
            const double H = in_bytes / 55e9, D = 0.35 * cap_bytes / 55e9, K = steps * 80e-9 + 3e-3;
            int best_n = 1;
            double best_t = 1e30;
            for (int n = 1; n <= 16; n++) {
                const double t = H / n + std::max((n - 1) * H / n, (n - 1) * K) + K + D / n;
                if (t < best_t - 1e-9) { best_t = t; best_n = n; }
            }
            target = std::max<uint64_t>(32ull << 20, (total_bytes + best_n - 1) / best_n);
        }
    }
    std::vector<int> cuts{0};
    {
        // decode: the first chunks are smaller (1/8, 1/4, 1/2 of the target) so that the device->host stream,
        // its bottleneck, starts early instead of waiting for a full-size chunk.  Encode is bound by the
        // host->device stream, which starts at once: equal chunks, as few as the floor above allows (every chunk
        // costs one latency-bound round of kernels, and those of consecutive chunks do not share SMs).
        uint64_t acc = 0;
        int ramp = (!enc && total_bytes > 4 * target) ? 3 : 0;
        for (int i = 0; i < nblk; i++) {
            uint64_t w = (uint64_t)in_len[i] + out_len[i];
            const uint64_t lim = std::max<uint64_t>(32ull << 20, target >> ramp);
            if (acc && acc + w > lim) { cuts.push_back(i); acc = 0; if (ramp) ramp--; }
            acc += w;
        }
        // a small remainder would still cost a whole kernel round (its slowest stream): it joins the previous chunk
        if (cuts.size() > 1 && acc < target / 4) cuts.pop_back();
        (void)0;
        cuts.push_back(nblk);
    }
    return cuts;
}

// One direction-agnostic driver: `enc` selects compress (order != NULL) or uncompress.
static int run_host_batch_body(hts_b200_ctx* ctx, bool enc, int nblk, const uint8_t* in_base, const uint64_t* in_off,
                               const uint32_t* in_len, uint8_t* out_base, const uint64_t* out_off, uint32_t* out_len,
                               int32_t* status, const uint8_t* method, const int32_t* order, HostRun* hr) {
    CK(cudaSetDevice(ctx->device));
    const bool shared_queue = hr && hr->cuts && hr->cursor;
    const std::vector<int> own_cuts = shared_queue ? std::vector<int>() : plan_chunks(enc, nblk, in_base, in_off, in_len, out_len, method, order);
    const std::vector<int>& cuts = shared_queue ? *hr->cuts : own_cuts;
    const int nchunk = (int)cuts.size() - 1;
    std::vector<int> redo;
    // capacities (out_len is overwritten with sizes as chunks complete); shared by the threads of a multi-device call
    const std::vector<uint32_t> own_caps = (hr && hr->caps) ? std::vector<uint32_t>() : std::vector<uint32_t>(out_len, out_len + nblk);
    const uint32_t* const caps = (hr && hr->caps) ? hr->caps : own_caps.data();
    hts_b200_dev_stats* stats = hr ? hr->stats : nullptr;

    auto launch_out = [&](Stage& S) -> int {                         // device->host half of a chunk
        const int a = S.c_a, n = S.c_n;
        const uint64_t* h_out_off = S.h_off.p + n;
        CK(cudaStreamWaitEvent(ctx->s_out, S.compute_done, 0));
        CK(cudaEventRecord(S.t_d2h0, ctx->s_out));
        CK(cudaMemcpyAsync(S.h_u32.p + n, S.d_u32.p + n, 8 * (size_t)n, cudaMemcpyDeviceToHost, ctx->s_out));   // out_len + status
        // A block may only write its own region out_base[out_off[i] .. + capacity): a single copy of the span is
        // used only when the regions tile it exactly; evenly spaced regions travel as one pitched copy of exactly
        // `capacity` bytes per row; anything else block by block.
        // Encode: a block's region is its capacity (~1.05 x the input) but its stream a fraction of that.  With equal,
        // evenly spaced regions the pitched copy fetches the leading W bytes of every block, W from the ratios the
        // previous chunks ended with; a stream that turns out longer gets its tail in collect_chunk.
        S.c_w = 0;
        const bool rows = (S.c_lay == STRIDED || (S.c_lay == TILED && S.c_equal)) && n > 1;
        if (enc && rows && ctx->enc_frac > 0) {
            const uint64_t w = std::min<uint64_t>(caps[a], align_up((uint64_t)(caps[a] * std::min(1.0, ctx->enc_frac * 1.15 + 0.02)), 256));
            if (w < caps[a]) S.c_w = w;
        }
        if (S.c_w) CK(cudaMemcpy2DAsync(out_base + S.c_lo, S.c_stride, S.d_out.p, S.c_stride, S.c_w, n, cudaMemcpyDeviceToHost, ctx->s_out));
        else if (S.c_lay == TILED) CK(cudaMemcpyAsync(out_base + S.c_lo, S.d_out.p, S.c_bytes, cudaMemcpyDeviceToHost, ctx->s_out));
        else if (S.c_lay == STRIDED) CK(cudaMemcpy2DAsync(out_base + S.c_lo, S.c_stride, S.d_out.p, S.c_stride, caps[a], n, cudaMemcpyDeviceToHost, ctx->s_out));
        else for (int i = 0; i < n; i++)
            if (caps[a + i]) CK(cudaMemcpyAsync(out_base + out_off[a + i], S.d_out.p + h_out_off[i], caps[a + i], cudaMemcpyDeviceToHost, ctx->s_out));
        CK(cudaEventRecord(S.t_d2h1, ctx->s_out));
        CK(cudaEventRecord(S.d2h_done, ctx->s_out));
        return 0;
    };
    auto launch_chunk = [&](int k, Stage& S, bool defer_out) -> int {
        const int a = cuts[k], b = cuts[k + 1], n = b - a;
        Range ri = span_of(in_off, in_len, a, b);
        Range ro = span_of(out_off, caps, a, b);
        const bool in_mirror = ri.lay != SCATTER;                    // reading a few padding bytes along is harmless
        const bool out_mirror = ro.lay == TILED || ro.lay == STRIDED;
        // staging layout mirrors the host layout where one (pitched) copy moves it, else blocks are packed 256-byte aligned
        uint64_t in_bytes = 0, out_bytes = 0;
        if (S.h_off.ensure(2 * (size_t)n) || S.h_u32.ensure(4 * (size_t)n) || S.h_method.ensure(n) ||
            S.d_off.ensure(2 * (size_t)n) || S.d_u32.ensure(4 * (size_t)n) || S.d_method.ensure(n)) return -1;
        uint64_t* h_in_off = S.h_off.p; uint64_t* h_out_off = S.h_off.p + n;
        uint32_t* h_in_len = S.h_u32.p; uint32_t* h_out_len = S.h_u32.p + n;
        uint32_t* h_order = S.h_u32.p + 3 * (size_t)n;
        bool equal = true;
        for (int i = 0; i < n; i++) {
            h_in_len[i] = in_len[a + i];
            h_out_len[i] = caps[a + i];
            if (caps[a + i] != caps[a]) equal = false;
            if (in_mirror) h_in_off[i] = in_off[a + i] - ri.lo; else { h_in_off[i] = in_bytes; in_bytes += align_up(in_len[a + i], 256); }
            if (out_mirror) h_out_off[i] = out_off[a + i] - ro.lo; else { h_out_off[i] = out_bytes; out_bytes += align_up(caps[a + i], 256); }
            if (method) S.h_method.p[i] = method[a + i];
            if (order) h_order[i] = (uint32_t)order[a + i];
        }
        if (in_mirror) in_bytes = ri.hi - ri.lo;
        if (out_mirror) out_bytes = ro.hi - ro.lo;
        if (S.d_in.ensure(in_bytes + 256) || S.d_out.ensure(out_bytes + 256)) {
            snprintf(ctx->err, sizeof(ctx->err), "out of device memory for staging");
            return -1;
        }
        // ---- H2D
        CK(cudaStreamWaitEvent(ctx->s_in, S.d2h_done, 0));           // previous user of this stage is done
        CK(cudaEventRecord(S.t_h2d0, ctx->s_in));
        if (in_mirror) CK(cudaMemcpyAsync(S.d_in.p, in_base + ri.lo, in_bytes, cudaMemcpyHostToDevice, ctx->s_in));
        else for (int i = 0; i < n; i++)
            if (in_len[a + i]) CK(cudaMemcpyAsync(S.d_in.p + h_in_off[i], in_base + in_off[a + i], in_len[a + i], cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaMemcpyAsync(S.d_off.p, S.h_off.p, 16 * (size_t)n, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaMemcpyAsync(S.d_u32.p, S.h_u32.p, 16 * (size_t)n, cudaMemcpyHostToDevice, ctx->s_in));
        if (method) CK(cudaMemcpyAsync(S.d_method.p, S.h_method.p, n, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaEventRecord(S.t_h2d1, ctx->s_in));
        CK(cudaEventRecord(S.h2d_done, ctx->s_in));
        // ---- kernels
        cudaStream_t cs = S.s_compute;
        CK(cudaStreamWaitEvent(cs, S.h2d_done, 0));
        CK(cudaEventRecord(S.t_k0, cs));
        uint32_t* d_status = S.d_u32.p + 2 * (size_t)n;
        if (!enc) {
            if (dec_prepare(ctx, S.dec, n, std::max<size_t>(ctx->arena_hint, 64u << 20))) return -1;
            DecodeBatch db;
            db.work = nullptr; db.hdr = nullptr; db.in_base = S.d_in.p; db.in_off = S.d_off.p; db.in_len = S.d_u32.p;
            db.out_base = S.d_out.p; db.out_off = S.d_off.p + n; db.out_len = S.d_u32.p + n;
            db.status = reinterpret_cast<int32_t*>(d_status); db.method = method ? S.d_method.p : nullptr; db.nblk = n;
            // host hint: which kernels can be needed (first byte of each stream; stripes hide their sub-streams)
            uint32_t kinds = 0, post = 0;
            for (int i = 0; i < n; i++) {
                if (!in_len[a + i]) continue;
                uint8_t f = in_base[in_off[a + i]];
                if (method && method[a + i] == 1) { kinds |= f ? K_R8_O1 : K_R8_O0; continue; }
                if (f & F_STRIPE) { kinds |= ~(K_R8 | K_BIG); post |= 7u; continue; }
                bool x32 = f & F_X32;
                if (f & F_CAT) kinds |= 1u << JK_COPY;
                else if (f & F_ORDER1) kinds |= (x32 ? (1u << JK_O1_32) | (1u << JK_O1_32S) : K_O1_4) | (1u << JK_TAB);   // TAB: compressed tables
                else kinds |= x32 ? (1u << JK_O0_32) : K_O0_4;
                if (f & F_RLE) { post |= 1u; kinds |= 1u << (x32 ? JK_O0_32 : JK_O0_4); }
                if (f & F_PACK) { post |= 2u; kinds |= 1u << (x32 ? JK_O0_32P : JK_O0_4P); }
            }
            db.kinds = kinds; db.post = post;
            if (S.side.init() == 0) db.side = &S.side;
            if (dec_enqueue(ctx, S.dec, db, cs)) return -1;
        } else {
            EncodeBatch eb;
            eb.in_base = S.d_in.p; eb.in_off = S.d_off.p; eb.in_len = S.d_u32.p;
            eb.out_base = S.d_out.p; eb.out_off = S.d_off.p + n; eb.out_len = S.d_u32.p + n;
            eb.status = reinterpret_cast<int32_t*>(d_status);
            eb.order = reinterpret_cast<const int32_t*>(S.d_u32.p + 3 * (size_t)n);
            eb.nblk = n;
            int l = encode_run(S.enc, eb, h_in_len, reinterpret_cast<const int32_t*>(h_order), cs, ctx->err, sizeof(ctx->err));
            if (l < 0) return -1;
            ctx->launches += l;
        }
        CK(cudaEventRecord(S.t_k1, cs));
        CK(cudaEventRecord(S.compute_done, cs));
        S.c_a = a; S.c_n = n; S.c_lay = out_mirror ? ro.lay : SCATTER; S.c_equal = equal; S.c_stride = ro.lay == STRIDED ? ro.stride : caps[a];
        S.c_lo = out_mirror ? ro.lo : 0; S.c_bytes = out_bytes;
        if (stats) { for (int i = a; i < b; i++) { stats->in_bytes += in_len[i]; } }
        return defer_out ? 0 : launch_out(S);
    };
    // HTSCODECS_B200_TRACE=1: one stderr line per chunk with its H2D / kernel / D2H intervals (ms since the call began)
    static const bool trace = getenv("HTSCODECS_B200_TRACE") && atoi(getenv("HTSCODECS_B200_TRACE")) != 0;
    cudaEvent_t t_base = nullptr;
    if (trace) { cudaEventCreate(&t_base); cudaEventRecord(t_base, ctx->s_in); }
    auto collect_chunk = [&](int k, Stage& S) -> int {               // after S.d2h_done
        const int a = cuts[k], n = cuts[k + 1] - a;
        if (trace && t_base) {
            float t[6] = {0, 0, 0, 0, 0, 0};
            cudaEvent_t ev[6] = {S.t_h2d0, S.t_h2d1, S.t_k0, S.t_k1, S.t_d2h0, S.t_d2h1};
            for (int q = 0; q < 6; q++) cudaEventElapsedTime(&t[q], t_base, ev[q]);
            fprintf(stderr, "[hts_b200 trace] dev %d chunk %d blocks %d: h2d %.2f-%.2f  kernels %.2f-%.2f  d2h %.2f-%.2f  host %.2f\n",
                    ctx->device, k, n, t[0], t[1], t[2], t[3], t[4], t[5], now_ms());
            cudaGetLastError();
        }
        if (stats) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, S.t_h2d0, S.t_h2d1) == cudaSuccess) stats->h2d_ms += ms;
            if (cudaEventElapsedTime(&ms, S.t_k0, S.t_k1) == cudaSuccess) stats->kernel_ms += ms;
            if (cudaEventElapsedTime(&ms, S.t_d2h0, S.t_d2h1) == cudaSuccess) stats->d2h_ms += ms;
            cudaGetLastError();
        }
        bool retry = false;
        if (!enc) {
            size_t ab = 0;
            if (dec_needs_retry(S.dec, &ab)) { ctx->arena_hint = std::max(ctx->arena_hint, ab); retry = true; }
            else ctx->arena_hint = std::max(ctx->arena_hint, (size_t)S.dec.h_work.p[1].arena_used);
        }
        if (enc && encode_needs_retry(S.enc)) retry = true;
        if (retry) { redo.push_back(k); return 0; }
        if (enc) {
            const uint32_t* len = S.h_u32.p + n;
            const int32_t* stt = reinterpret_cast<const int32_t*>(S.h_u32.p + 2 * (size_t)n);
            const uint64_t* h_out_off = S.h_off.p + n;
            double frac = 0;
            bool tails = false;
            for (int i = 0; i < n; i++) {
                if (stt[i] != 0 || !caps[a + i]) continue;
                frac = std::max(frac, (double)len[i] / caps[a + i]);
                if (S.c_w && len[i] > S.c_w) {                       // longer than predicted: fetch the rest now
                    CK(cudaMemcpyAsync(out_base + out_off[a + i] + S.c_w, S.d_out.p + h_out_off[i] + S.c_w, len[i] - S.c_w,
                                       cudaMemcpyDeviceToHost, ctx->s_out));
                    tails = true;
                }
            }
            if (tails) CK(cudaStreamSynchronize(ctx->s_out));
            ctx->enc_frac = ctx->enc_frac > 0 ? std::max(frac, 0.5 * (ctx->enc_frac + frac)) : frac;
        }
        memcpy(out_len + a, S.h_u32.p + n, 4 * (size_t)n);
        memcpy(status + a, S.h_u32.p + 2 * (size_t)n, 4 * (size_t)n);
        if (stats) for (int i = 0; i < n; i++) if (status[a + i] == 0) stats->out_bytes += out_len[a + i];
        if (hr && hr->on_chunk) hr->on_chunk(a, n);
        return 0;
    };

    const bool phased = hr && hr->phase;
    if (!phased && (ctx->full_duplex || nchunk == 1 || shared_queue)) {
        // ---- software pipeline: the j-th chunk this context takes uses stage j % NSTAGE; the chunk that used the
        // stage before is collected first.  Chunks come in order, or -- multi-device call -- from the shared cursor.
        int inflight[NSTAGE];
        int taken = 0, local_next = 0;
        for (;;) {
            const int k = shared_queue ? hr->cursor->fetch_add(1) : local_next++;
            if (k >= nchunk) break;
            Stage& S = ctx->stage[taken % NSTAGE];
            if (taken >= NSTAGE) {
                CK(cudaEventSynchronize(S.d2h_done));
                if (collect_chunk(inflight[taken % NSTAGE], S)) return -1;
            }
            if (launch_chunk(k, S, false)) return -1;
            inflight[taken % NSTAGE] = k;
            taken++;
            if (stats && shared_queue) stats->nblk += cuts[k + 1] - cuts[k];
        }
        for (int j = std::max(0, taken - NSTAGE); j < taken; j++) {
            Stage& S = ctx->stage[j % NSTAGE];
            CK(cudaEventSynchronize(S.d2h_done));
            if (collect_chunk(inflight[j % NSTAGE], S)) return -1;
        }
    } else {
        // ---- half duplex: every chunk has its own stage; all host->device copies (and the kernels behind
        // them) are enqueued first, the device->host copies start once the last input has landed -- on every
        // device of a multi-device call (PhaseSync).  Some hosts lose most of their device->host rate while
        // any host->device traffic is in flight.
        if ((int)ctx->stage.size() < nchunk) ctx->stage.resize(nchunk);
        const double t0 = now_ms();
        for (int k = 0; k < nchunk; k++) {
            if (ctx->stage[k].init()) { snprintf(ctx->err, sizeof(ctx->err), "cannot create a staging slot"); return -1; }
            if (launch_chunk(k, ctx->stage[k], true)) return -1;
        }
        CK(cudaEventRecord(ctx->all_h2d, ctx->s_in));
        if (phased) {
            CK(cudaEventSynchronize(ctx->all_h2d));
            const double t1 = now_ms();
            hr->arrived = true;
            hr->phase->arrive_and_wait();
            if (stats) { stats->h2d_phase_ms = t1 - t0; stats->wait_ms = now_ms() - t1; }
        } else {
            CK(cudaStreamWaitEvent(ctx->s_out, ctx->all_h2d, 0));
        }
        const double t2 = now_ms();
        for (int k = 0; k < nchunk; k++) if (launch_out(ctx->stage[k])) return -1;
        for (int k = 0; k < nchunk; k++) {
            CK(cudaEventSynchronize(ctx->stage[k].d2h_done));
            if (collect_chunk(k, ctx->stage[k])) return -1;
        }
        if (stats) stats->d2h_phase_ms = now_ms() - t2;
    }
    // ---- rare: chunks whose scratch overflowed are redone one at a time with the grown arena
    for (int attempt = 0; !redo.empty() && attempt < 8; attempt++) {
        std::vector<int> again;
        again.swap(redo);
        for (int k : again) {
            Stage& S = ctx->stage[0];
            if (launch_chunk(k, S, false)) return -1;
            CK(cudaEventSynchronize(S.d2h_done));
            if (collect_chunk(k, S)) return -1;
        }
    }
    if (!redo.empty()) { snprintf(ctx->err, sizeof(ctx->err), "work area kept overflowing"); return -1; }
    return 0;
}

static int run_host_batch(hts_b200_ctx* ctx, bool enc, int nblk, const uint8_t* in_base, const uint64_t* in_off,
                          const uint32_t* in_len, uint8_t* out_base, const uint64_t* out_off, uint32_t* out_len,
                          int32_t* status, const uint8_t* method, const int32_t* order, HostRun* hr = nullptr) {
    int rc = -1;
    if (ctx && nblk >= 0) {
        const double t0 = now_ms();
        rc = nblk == 0 ? 0 : run_host_batch_body(ctx, enc, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, method, order, hr);
        if (rc != 0) {
            // no copy into the caller's buffers may still be in flight when the call returns
            cudaSetDevice(ctx->device);
            cudaStreamSynchronize(ctx->s_in);
            for (auto& S : ctx->stage) if (S.s_compute) cudaStreamSynchronize(S.s_compute);
            cudaStreamSynchronize(ctx->s_out);
            cudaGetLastError();
        }
        if (hr && hr->stats) hr->stats->wall_ms = now_ms() - t0;
    }
    if (hr && hr->phase && !hr->arrived) { hr->arrived = true; hr->phase->arrive_and_wait(); }   // never strand the other devices
    return rc;
}

extern "C" int hts_b200_plan_chunks(int enc, int nblk, const uint8_t* in_base, const uint64_t* in_off,
                                    const uint32_t* in_len, const uint32_t* out_len, const uint8_t* method,
                                    const int32_t* order, int* cuts, int max_cuts) {
    if (nblk <= 0 || (enc && !order) || (!enc && !in_base)) return -1;
    const std::vector<int> c = plan_chunks(enc != 0, nblk, in_base, in_off, in_len, out_len, method, order);
    for (size_t k = 0; k < c.size() && (int)k < max_cuts; k++) cuts[k] = c[k];
    return (int)c.size();
}

extern "C" int hts_b200_uncompress_batch_host(hts_b200_ctx* ctx, int nblk, const uint8_t* in_base,
                                              const uint64_t* in_off, const uint32_t* in_len, uint8_t* out_base,
                                              const uint64_t* out_off, uint32_t* out_len, int32_t* status,
                                              const uint8_t* method) {
    return run_host_batch(ctx, false, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, method, nullptr);
}

extern "C" int hts_b200_compress_batch_host(hts_b200_ctx* ctx, int nblk, const uint8_t* in_base,
                                            const uint64_t* in_off, const uint32_t* in_len, uint8_t* out_base,
                                            const uint64_t* out_off, uint32_t* out_len, int32_t* status,
                                            const int32_t* order) {
    if (!order) return -1;
    return run_host_batch(ctx, true, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, nullptr, order);
}

// ------------------------------------------------------------------------------------------
// multi-device host-buffer calls: one thread + context per device, copy phases coordinated across devices
// ------------------------------------------------------------------------------------------
extern "C" int hts_b200_partition(int nblk, const uint32_t* weight, int nparts, int* cuts) {
    if (nblk < 0 || nparts <= 0 || !cuts || (nblk && !weight)) return -1;
    typedef unsigned __int128 u128;
    uint64_t total = 0;
    for (int i = 0; i < nblk; i++) total += weight[i];
    // part k ends at the block boundary nearest to k / nparts of the total weight (all comparisons scaled by nparts)
    cuts[0] = 0;
    uint64_t before = 0;         // weight of blocks [0, i)
    int i = 0;
    for (int k = 1; k < nparts; k++) {
        int c;
        if (total == 0) c = (int)((long long)nblk * k / nparts);
        else {
            const u128 tgt = (u128)total * k;
            while (i < nblk && (u128)(before + weight[i]) * nparts < tgt) before += weight[i++];
            c = i;               // first block whose inclusive prefix reaches the target
            if (c < nblk && (u128)(before + weight[c]) * nparts - tgt <= tgt - (u128)before * nparts) c++;
        }
        cuts[k] = std::min(std::max(c, cuts[k - 1]), nblk);
    }
    cuts[nparts] = nblk;
    return 0;
}

namespace {
std::mutex g_multi_mu;
std::map<int, hts_b200_ctx*> g_multi_ctx;
std::vector<hts_b200_dev_stats> g_multi_stats;
char g_multi_err[320] = {0};
int g_multi_phased = -1;     // -1: from the environment (default 0)
struct MultiReaper { ~MultiReaper() { for (auto& kv : g_multi_ctx) hts_b200_destroy(kv.second); g_multi_ctx.clear(); } } g_multi_reaper;

int run_multi(bool enc, int ndev, const int* devices, int nblk, const uint8_t* in_base, const uint64_t* in_off,
              const uint32_t* in_len, uint8_t* out_base, const uint64_t* out_off, uint32_t* out_len,
              int32_t* status, const uint8_t* method, const int32_t* order) {
    if (ndev <= 0 || !devices || nblk < 0 || (enc && !order)) return -1;
    std::lock_guard<std::mutex> lk(g_multi_mu);
    g_multi_err[0] = 0;
    g_multi_stats.assign(ndev, hts_b200_dev_stats{});
    if (nblk == 0) return 0;
    for (int d = 0; d < ndev; d++) {
        for (int e = 0; e < d; e++) if (devices[e] == devices[d]) { snprintf(g_multi_err, sizeof(g_multi_err), "device %d listed twice", devices[d]); return -1; }
        if (!g_multi_ctx.count(devices[d])) {
            hts_b200_ctx* c = hts_b200_create(devices[d]);
            if (!c) { snprintf(g_multi_err, sizeof(g_multi_err), "cannot create a context on device %d (no CPU fallback)", devices[d]); return -1; }
            g_multi_ctx[devices[d]] = c;
        }
    }
    // Copy policy.  Default: every device runs its own full-duplex chunk pipeline and takes chunks of the whole batch
    // from a shared cursor.  Phased (all devices send, meet at a barrier, then fetch; static partition on uncompressed
    // bytes) is there for hosts that lose device->host rate while host->device copies are in flight; on the boxes
    // measured in round 2 it was never faster (2 x B200: 86 vs 98 GB/s; 8 x B200: 82 vs 85 before balancing).
    int phased = g_multi_phased;
    if (phased < 0) { const char* e = getenv("HTSCODECS_B200_MULTI_PHASED"); phased = e ? atoi(e) != 0 : 0; }
    if (ndev == 1) phased = 0;
    const std::vector<uint32_t> caps(out_len, out_len + nblk);
    std::vector<int> cuts(ndev + 1);
    std::vector<int> chunks;
    std::atomic<int> cursor{0};
    if (phased) hts_b200_partition(nblk, enc ? in_len : out_len, ndev, cuts.data());      // uncompressed bytes
    else {
        bool latency_bound = false;
        chunks = plan_chunks(enc, nblk, in_base, in_off, in_len, out_len, method, order, &latency_bound);
        // finer chunks towards the end (halves, then quarters): the devices then finish within a fraction of a chunk of
        // each other instead of a whole one (a slow device's chunk takes ~35 ms on the 8 x B200 host).  Not when the
        // chunks are as small as the kernels' latency allows (4-way streams): every extra chunk costs a kernel round.
        for (int pass = 0; pass < 2 && ndev > 1 && !latency_bound; pass++) {
            const int nch = (int)chunks.size() - 1, first = std::max(0, nch - 2 * ndev);
            std::vector<int> fine(chunks.begin(), chunks.begin() + first + 1);
            for (int k = first; k < nch; k++) {
                const int a = chunks[k], b = chunks[k + 1];
                if (b - a >= 2) fine.push_back(a + (b - a) / 2);
                fine.push_back(b);
            }
            if (nch > ndev) chunks.swap(fine);
        }
    }
    PhaseSync ps;
    ps.want = ndev;
    std::vector<int> rcs(ndev, 0);
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; d++) {
        th.emplace_back([&, d] {
            hts_b200_ctx* c = g_multi_ctx[devices[d]];
            hts_b200_dev_stats& st = g_multi_stats[d];
            st.device = devices[d];
            HostRun hr;
            hr.stats = &st;
            if (phased) {
                const int a = cuts[d], n = cuts[d + 1] - a;
                st.first_blk = a; st.nblk = n;
                hr.phase = &ps;
                rcs[d] = run_host_batch(c, enc, n, in_base, in_off + a, in_len + a, out_base, out_off + a, out_len + a,
                                        status + a, method ? method + a : nullptr, order ? order + a : nullptr, &hr);
            } else {
                st.first_blk = -1; st.nblk = 0;                      // counted as chunks are taken
                hr.cuts = &chunks; hr.cursor = &cursor; hr.caps = caps.data();
                rcs[d] = run_host_batch(c, enc, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, method, order, &hr);
            }
        });
    }
    for (auto& t : th) t.join();
    for (int d = 0; d < ndev; d++)
        if (rcs[d] != 0) {
            snprintf(g_multi_err, sizeof(g_multi_err), "device %d: %.250s", devices[d], g_multi_ctx[devices[d]]->err);
            return -1;
        }
    return 0;
}
}  // namespace

extern "C" int hts_b200_uncompress_batch_host_multi(int ndev, const int* devices, int nblk, const uint8_t* in_base,
                                                    const uint64_t* in_off, const uint32_t* in_len, uint8_t* out_base,
                                                    const uint64_t* out_off, uint32_t* out_len, int32_t* status,
                                                    const uint8_t* method) {
    return run_multi(false, ndev, devices, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, method, nullptr);
}
extern "C" int hts_b200_compress_batch_host_multi(int ndev, const int* devices, int nblk, const uint8_t* in_base,
                                                  const uint64_t* in_off, const uint32_t* in_len, uint8_t* out_base,
                                                  const uint64_t* out_off, uint32_t* out_len, int32_t* status,
                                                  const int32_t* order) {
    return run_multi(true, ndev, devices, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, nullptr, order);
}
extern "C" void hts_b200_multi_set_phased(int phased) { std::lock_guard<std::mutex> lk(g_multi_mu); g_multi_phased = phased < 0 ? -1 : (phased ? 1 : 0); }
extern "C" int hts_b200_multi_last_stats(hts_b200_dev_stats* out, int max) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    for (int i = 0; i < (int)g_multi_stats.size() && i < max; i++) out[i] = g_multi_stats[i];
    return (int)g_multi_stats.size();
}
extern "C" unsigned long long hts_b200_multi_launch_count(void) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    unsigned long long n = 0;
    for (auto& kv : g_multi_ctx) n += kv.second->launches;
    return n;
}
extern "C" const char* hts_b200_multi_last_error(void) { return g_multi_err; }

static int compress_dev(hts_b200_ctx* ctx, int nblk, const uint8_t* in_base, const uint64_t* in_off, const uint32_t* in_len,
                        uint8_t* out_base, const uint64_t* out_off, uint32_t* out_len, int32_t* status, const int32_t* order,
                        const uint32_t* h_in_len, const int32_t* h_order, int sync) {
    if (!ctx || nblk < 0 || !order) return -1;
    if (nblk == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    EncodeBatch eb;
    eb.in_base = in_base; eb.in_off = in_off; eb.in_len = in_len; eb.out_base = out_base; eb.out_off = out_off;
    eb.out_len = out_len; eb.status = status; eb.order = order; eb.nblk = nblk;
    for (int attempt = 0; attempt < 6; attempt++) {
        // out_len is capacity in, size out: a retry needs the capacities back
        if (sync) {
            if (ctx->dec.cap_save.ensure(nblk)) { snprintf(ctx->err, sizeof(ctx->err), "out of device memory"); return -1; }
            if (attempt == 0) CK(cudaMemcpyAsync(ctx->dec.cap_save.p, out_len, 4 * (size_t)nblk, cudaMemcpyDeviceToDevice, ctx->stream));
            else CK(cudaMemcpyAsync(out_len, ctx->dec.cap_save.p, 4 * (size_t)nblk, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        int l = encode_run(ctx->enc, eb, h_in_len, h_order, ctx->stream, ctx->err, sizeof(ctx->err));
        if (l < 0) return -1;
        ctx->launches += l;
        if (!sync) return 0;
        CK(cudaStreamSynchronize(ctx->stream));
        if (!encode_needs_retry(ctx->enc)) return 0;                 // (rare: an order-1 alphabet larger than the arena was sized for)
    }
    snprintf(ctx->err, sizeof(ctx->err), "encode arena kept overflowing");
    return -1;
}

extern "C" int hts_b200_compress_batch_dev(hts_b200_ctx* ctx, int nblk, const uint8_t* in_base,
                                           const uint64_t* in_off, const uint32_t* in_len, uint8_t* out_base,
                                           const uint64_t* out_off, uint32_t* out_len, int32_t* status,
                                           const int32_t* order, int sync) {
    return compress_dev(ctx, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, order, nullptr, nullptr, sync);
}

extern "C" int hts_b200_compress_batch_dev_async(hts_b200_ctx* ctx, int nblk, const uint8_t* in_base,
                                                 const uint64_t* in_off, const uint32_t* in_len, uint8_t* out_base,
                                                 const uint64_t* out_off, uint32_t* out_len, int32_t* status,
                                                 const int32_t* order, const uint32_t* host_in_len,
                                                 const int32_t* host_order) {
    if (!host_in_len || !host_order) return -1;
    return compress_dev(ctx, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, order, host_in_len, host_order, 0);
}

// ------------------------------------------------------------------------------------------
// pointer-array wrappers
// ------------------------------------------------------------------------------------------
// Gather / scatter between the caller's per-block buffers and the pinned staging arena.  A single thread copies at
// ~10 GB/s, which for a large batch costs more than the GPU work: split the blocks over a few host threads.
template <typename F> static void for_blocks_parallel(int nblk, uint64_t bytes, F f, int max_threads = 8) {
    const int nt = (int)std::min<uint64_t>(max_threads, std::max<uint64_t>(1, bytes >> 25));   // one thread per 32 MiB, at most 8
    if (nt <= 1) { f(0, nblk); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([=] { f((int)((long long)nblk * t / nt), (int)((long long)nblk * (t + 1) / nt)); });
    for (auto& x : th) x.join();
}

static int run_ptr_batch(hts_b200_ctx* ctx, bool enc, int nblk, const unsigned char* const* in,
                         const unsigned int* in_size, unsigned char* const* out, unsigned int* out_size,
                         const int* order, int* status, const uint8_t* method) {
    if (!ctx || nblk < 0) return -1;
    if (nblk == 0) return 0;
    std::vector<uint64_t> off(2 * (size_t)nblk);
    uint64_t ib = 0, ob = 0;
    for (int i = 0; i < nblk; i++) {
        off[i] = ib; ib += align_up(in_size[i], 16);
        off[nblk + i] = ob; ob += align_up(out_size[i], 16);
    }
    if (ctx->pin_in.ensure(ib + 16) || ctx->pin_out.ensure(ob + 16)) { snprintf(ctx->err, sizeof(ctx->err), "out of pinned memory"); return -1; }
    uint8_t* const pin_in = ctx->pin_in.p;
    const uint64_t* const offp = off.data();
    for_blocks_parallel(nblk, ib, [=](int a, int b) { for (int i = a; i < b; i++) memcpy(pin_in + offp[i], in[i], in_size[i]); });
    std::vector<int32_t> st(nblk), ord;
    if (order) ord.assign(order, order + nblk);
    // The results of a chunk are scattered to the caller's buffers as soon as the chunk has landed, by threads of their
    // own, while the pipeline moves the following chunks (scattering after the call cost as much as the call itself).
    const uint8_t* const pin_out = ctx->pin_out.p;
    const int32_t* const stp = st.data();
    std::vector<std::thread> scatter;
    HostRun hr;
    hr.on_chunk = [&](int a, int n) {
        uint64_t bytes = 0;
        for (int i = a; i < a + n; i++) if (stp[i] == 0) bytes += out_size[i];
        scatter.emplace_back([=] {
            for_blocks_parallel(n, bytes, [=](int x, int y) {
                for (int i = a + x; i < a + y; i++) if (stp[i] == 0) memcpy(out[i], pin_out + offp[nblk + i], out_size[i]);
            }, 4);
        });
    };
    int rc = run_host_batch(ctx, enc, nblk, ctx->pin_in.p, off.data(), in_size, ctx->pin_out.p, off.data() + nblk,
                            out_size, st.data(), method, order ? ord.data() : nullptr, &hr);
    for (auto& t : scatter) t.join();
    if (rc) return rc;
    for (int i = 0; i < nblk; i++) status[i] = st[i];
    return 0;
}

extern "C" int rans4x16_uncompress_batch(hts_b200_ctx* ctx, int nblk, const unsigned char* const* in,
                                         const unsigned int* in_size, unsigned char* const* out,
                                         unsigned int* out_size, int* status) {
    return run_ptr_batch(ctx, false, nblk, in, in_size, out, out_size, nullptr, status, nullptr);
}

extern "C" int rans4x16_compress_batch(hts_b200_ctx* ctx, int nblk, const unsigned char* const* in,
                                       const unsigned int* in_size, unsigned char* const* out,
                                       unsigned int* out_size, const int* order, int* status) {
    if (!order) return -1;
    return run_ptr_batch(ctx, true, nblk, in, in_size, out, out_size, order, status, nullptr);
}

// ------------------------------------------------------------------------------------------
// try-all-methods encode (tokenise_name3.c:1246-1299)
// ------------------------------------------------------------------------------------------
// Winner i: len[i] bytes from src + src_off[i] to dst + dst_off[i]; offsets and slot sizes are multiples of 16.
__global__ void gather_kernel(const uint8_t* src, const uint64_t* src_off, const uint32_t* len, uint8_t* dst,
                              const uint64_t* dst_off, int n) {
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint4* s = reinterpret_cast<const uint4*>(src + src_off[i]);
        uint4* d = reinterpret_cast<uint4*>(dst + dst_off[i]);
        const uint32_t nv = (len[i] + 15) / 16;
        for (uint32_t t = threadIdx.x; t < nv; t += blockDim.x) d[t] = s[t];
    }
}

extern "C" int rans4x16_compress_best_batch(hts_b200_ctx* ctx, int nblk, const unsigned char* const* in,
                                            const unsigned int* in_size, unsigned char* const* out,
                                            unsigned int* out_size, const int* methods, int nmethods, int* best,
                                            int* status) {
    if (!ctx || nblk < 0 || nmethods <= 0 || !methods || !status) return -1;
    if (nblk == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    const size_t cand_budget = (size_t)1 << 30;                   // candidate bytes per pass
    int a = 0;
    while (a < nblk) {
        // ---- blocks [a, b) of this pass and their candidates
        size_t in_bytes = 0, cand_bytes = 0, ncand = 0;
        int b = a;
        for (; b < nblk; b++) {
            size_t cb = 0, nc = 0;
            for (int m = 0; m < nmethods; m++) {
                if ((methods[m] & RANS_ORDER_STRIPE) && in_size[b] % 4) continue;        // :1269
                cb += align_up(rans_compress_bound_4x16(in_size[b], methods[m]), 16); nc++;
            }
            if (b > a && cand_bytes + cb > cand_budget) break;
            cand_bytes += cb; ncand += nc; in_bytes += align_up(in_size[b], 16);
        }
        const int nb = b - a;
        if (ctx->pin_in.ensure(in_bytes + 16) || ctx->best_in.ensure(in_bytes + 16) || ctx->best_cand.ensure(cand_bytes + 16) ||
            ctx->best_off.ensure(2 * ncand + 2 * (size_t)nb + 2) || ctx->best_u32.ensure(4 * ncand + nb + 4) ||
            ctx->best_hoff.ensure(2 * ncand + 2 * (size_t)nb + 2) || ctx->best_hu32.ensure(4 * ncand + nb + 4)) {
            snprintf(ctx->err, sizeof(ctx->err), "out of memory for the try-all-methods pass");
            return -1;
        }
        uint64_t* h_in_off = ctx->best_hoff.p; uint64_t* h_c_off = h_in_off + ncand;
        uint32_t* h_in_len = ctx->best_hu32.p; uint32_t* h_c_len = h_in_len + ncand;
        uint32_t* h_st = h_c_len + ncand; uint32_t* h_ord = h_st + ncand;
        size_t io = 0, co = 0, k = 0;
        for (int i = a; i < b; i++) {
            if (in_size[i]) memcpy(ctx->pin_in.p + io, in[i], in_size[i]);
            for (int m = 0; m < nmethods; m++) {
                if ((methods[m] & RANS_ORDER_STRIPE) && in_size[i] % 4) continue;
                const uint32_t bound = rans_compress_bound_4x16(in_size[i], methods[m]);
                h_in_off[k] = io; h_in_len[k] = in_size[i]; h_c_off[k] = co; h_c_len[k] = bound;
                h_st[k] = 0; h_ord[k] = (uint32_t)methods[m];
                co += align_up(bound, 16); k++;
            }
            io += align_up(in_size[i], 16);
        }
        cudaStream_t st = ctx->stream;
        CK(cudaMemcpyAsync(ctx->best_in.p, ctx->pin_in.p, in_bytes, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->best_off.p, ctx->best_hoff.p, 16 * ncand, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->best_u32.p, ctx->best_hu32.p, 16 * ncand, cudaMemcpyHostToDevice, st));
        uint32_t* d_in_len = ctx->best_u32.p; uint32_t* d_c_len = d_in_len + ncand;
        uint32_t* d_st = d_c_len + ncand; uint32_t* d_ord = d_st + ncand;
        // ---- every candidate of every block in one encode batch
        if (ncand == 0) { for (int i = a; i < b; i++) status[i] = HTS_B200_ERR_INTERNAL; a = b; continue; }
        EncodeBatch eb;
        eb.in_base = ctx->best_in.p; eb.in_off = ctx->best_off.p; eb.in_len = d_in_len;
        eb.out_base = ctx->best_cand.p; eb.out_off = ctx->best_off.p + ncand; eb.out_len = d_c_len;
        eb.status = reinterpret_cast<int32_t*>(d_st); eb.order = reinterpret_cast<const int32_t*>(d_ord);
        eb.nblk = (int)ncand;
        const std::vector<uint32_t> bounds(h_c_len, h_c_len + ncand);
        for (int attempt = 0;; attempt++) {
            if (attempt) {                                               // the arena was too small: capacities back, run again
                memcpy(h_c_len, bounds.data(), 4 * ncand);
                memset(h_st, 0, 4 * ncand);
                CK(cudaMemcpyAsync(ctx->best_u32.p, ctx->best_hu32.p, 16 * ncand, cudaMemcpyHostToDevice, st));
            }
            int l = encode_run(ctx->enc, eb, h_in_len, reinterpret_cast<const int32_t*>(h_ord), st, ctx->err, sizeof(ctx->err));
            if (l < 0) return -1;
            ctx->launches += l;
            CK(cudaMemcpyAsync(h_c_len, d_c_len, 8 * ncand, cudaMemcpyDeviceToHost, st));     // lengths + status
            CK(cudaStreamSynchronize(st));
            if (!encode_needs_retry(ctx->enc) || attempt >= 5) break;
        }
        // ---- pick the winners (first strictly smallest, :1280), gather them densely, fetch
        uint64_t* h_src = ctx->best_hoff.p + 2 * ncand; uint64_t* h_dst = h_src + nb;
        std::vector<uint32_t> wlen(nb);
        size_t dense = 0;
        k = 0;
        for (int i = a; i < b; i++) {
            uint64_t best_sz = UINT64_MAX; int bm = -1; size_t bk = 0; int fail = 0;
            for (int m = 0; m < nmethods; m++) {
                if ((methods[m] & RANS_ORDER_STRIPE) && in_size[i] % 4) continue;
                if ((int32_t)h_st[k] != 0) fail = (int32_t)h_st[k];                      // :1273 any failure fails the column
                else if (h_c_len[k] < best_sz) { best_sz = h_c_len[k]; bm = m; bk = k; }
                k++;
            }
            const int j = i - a;
            if (fail || bm < 0) { status[i] = fail ? fail : HTS_B200_ERR_INTERNAL; wlen[j] = 0; h_src[j] = 0; h_dst[j] = dense; continue; }
            if (best_sz > out_size[i]) { status[i] = HTS_B200_ERR_SIZE; wlen[j] = 0; h_src[j] = 0; h_dst[j] = dense; continue; }
            status[i] = HTS_B200_OK;
            if (best) best[i] = methods[bm];
            out_size[i] = (unsigned int)best_sz;
            wlen[j] = (uint32_t)best_sz; h_src[j] = h_c_off[bk]; h_dst[j] = dense;
            dense += align_up(best_sz, 16);
        }
        if (dense) {
            if (ctx->best_out.ensure(dense + 16) || ctx->pin_out.ensure(dense + 16)) { snprintf(ctx->err, sizeof(ctx->err), "out of memory for the winners"); return -1; }
            uint32_t* h_wlen = ctx->best_hu32.p;                     // (candidate arrays are no longer needed)
            memcpy(h_wlen, wlen.data(), 4 * (size_t)nb);
            CK(cudaMemcpyAsync(ctx->best_off.p + 2 * ncand, h_src, 16 * (size_t)nb, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(ctx->best_u32.p, h_wlen, 4 * (size_t)nb, cudaMemcpyHostToDevice, st));
            gather_kernel<<<std::min(nb, 148 * 8), 256, 0, st>>>(ctx->best_cand.p, ctx->best_off.p + 2 * ncand, ctx->best_u32.p,
                                                                ctx->best_out.p, ctx->best_off.p + 2 * ncand + nb, nb);
            ctx->launches++;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(ctx->pin_out.p, ctx->best_out.p, dense, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (int j = 0; j < nb; j++) if (wlen[j]) memcpy(out[a + j], ctx->pin_out.p + h_dst[j], wlen[j]);
        }
        a = b;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// drop-in single-block API
// ------------------------------------------------------------------------------------------
static hts_b200_ctx* tls_ctx() {
    struct Holder { hts_b200_ctx* c = nullptr; ~Holder() { if (c) hts_b200_destroy(c); } };
    static thread_local Holder h;
    if (!h.c) h.c = hts_b200_create(-1);
    return h.c;
}

static int host_var_get_u32(const uint8_t* p, const uint8_t* end, uint32_t* v) {   // varint.h:131-160
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

extern "C" int hts_b200_peek_size(const uint8_t* in, uint32_t in_len, int method, uint32_t* ulen) {
    if (method == HTS_B200_RANS4x8) {
        if (in_len < 9) return -1;
        *ulen = (uint32_t)in[5] | ((uint32_t)in[6] << 8) | ((uint32_t)in[7] << 16) | ((uint32_t)in[8] << 24);
        return 0;
    }
    if (in_len < 2 || ((in[0] & RANS_ORDER_NOSZ) && !(in[0] & RANS_ORDER_STRIPE))) return -1;
    return host_var_get_u32(in + 1, in + in_len, ulen) ? 0 : -1;
}

// ------------------------------------------------------------------------------------------
// container glue: [u32 clen][stream]... (tests/rANS_static4x16pr_test.c:261-296)
// ------------------------------------------------------------------------------------------
extern "C" long hts_b200_frames_scan(const uint8_t* buf, size_t len, int method, long max_blk, uint64_t* in_off,
                                     uint32_t* in_len, uint64_t* out_off, uint32_t* out_len, uint64_t* out_total,
                                     uint32_t out_align) {
    if (!buf && len) return -1;
    if (out_align == 0) out_align = 1;
    size_t pos = 0;
    long n = 0;
    uint64_t out_pos = 0;
    while (pos < len) {
        if (len - pos < 4) return -1;
        uint32_t clen;
        memcpy(&clen, buf + pos, 4);                                 // native endian, like the reference's fwrite
        pos += 4;
        if (clen > len - pos) return -1;
        uint32_t ulen = 0;
        if (hts_b200_peek_size(buf + pos, clen, method, &ulen) != 0) return -1;
        if (n < max_blk) {
            if (in_off) in_off[n] = pos;
            if (in_len) in_len[n] = clen;
            if (out_off) out_off[n] = out_pos;
            if (out_len) out_len[n] = ulen;
        }
        out_pos += ((uint64_t)ulen + out_align - 1) / out_align * out_align;
        pos += clen;
        n++;
    }
    if (out_total) *out_total = out_pos;
    return n;
}

extern "C" size_t hts_b200_frames_write(uint8_t* dst, size_t dst_cap, long nblk, const uint8_t* src_base,
                                        const uint64_t* src_off, const uint32_t* src_len, const int32_t* status) {
    size_t need = 0;
    for (long i = 0; i < nblk; i++) if (!status || status[i] == 0) need += 4 + (size_t)src_len[i];
    if (!dst) return need;
    if (need > dst_cap) return (size_t)-1;
    size_t pos = 0;
    for (long i = 0; i < nblk; i++) {
        if (status && status[i] != 0) continue;
        memcpy(dst + pos, &src_len[i], 4);
        memcpy(dst + pos + 4, src_base + src_off[i], src_len[i]);
        pos += 4 + (size_t)src_len[i];
    }
    return pos;
}

// rans_compress_bound_4x16, reference rANS_static4x16pr.c:360-372 (double arithmetic on purpose)
extern "C" unsigned int rans_compress_bound_4x16(unsigned int size, int order) {
    int N = order >> 8;
    if (!N) N = 4;
    order &= 0xff;
    double d = 1.05 * size;
    d += (order == 0) ? (257 * 3 + 4) : (257 * 257 * 3 + 4 + 257 * 3 + 4);
    d += (order & RANS_ORDER_PACK) ? 1 : 0;
    d += (order & RANS_ORDER_RLE) ? (1 + 257 * 3 + 4) : 0;
    d += 20;
    d += (order & RANS_ORDER_STRIPE) ? (1 + 5 * N) : 0;
    int sz = (int)d;
    return (unsigned int)(sz + (sz & 1) + 2);
}

static unsigned char* uncompress_one(unsigned char* in, unsigned int in_size, unsigned char* out,
                                     unsigned int* out_size, int method) {
    if (!in || !out_size) return nullptr;
    uint32_t ulen = 0;
    bool have = hts_b200_peek_size(in, in_size, method, &ulen) == 0;
    if (method == HTS_B200_RANS4x16) {
        if (in_size == 0) return nullptr;                            // :1357
        if (in[0] & RANS_ORDER_STRIPE) {
            if (!have) return nullptr;
            if (out && ulen != *out_size) return nullptr;            // :1379 exact size required
        } else if (in[0] & RANS_ORDER_NOSZ) {
            if (!out) return nullptr;                                // :1456
            ulen = *out_size; have = true;
        } else if (!have) return nullptr;
        if (out && *out_size < ulen) return nullptr;                 // :1464
    } else if (!have) return nullptr;
    if (ulen >= 0x7fffffffu) return nullptr;
    hts_b200_ctx* ctx = tls_ctx();
    if (!ctx) { fprintf(stderr, "htscodecs_b200: no usable sm_100 device (there is no CPU fallback)\n"); return nullptr; }
    unsigned char* dst = out ? out : (unsigned char*)malloc(ulen ? ulen : 1);
    if (!dst) return nullptr;
    const unsigned char* ins[1] = {in};
    unsigned char* outs[1] = {dst};
    unsigned int isz[1] = {in_size}, osz[1] = {out ? *out_size : ulen};
    if (method == HTS_B200_RANS4x16 && (in[0] & RANS_ORDER_STRIPE)) osz[0] = ulen;
    int st[1] = {0};
    uint8_t m[1] = {(uint8_t)method};
    int rc = run_ptr_batch(ctx, false, 1, ins, isz, outs, osz, nullptr, st, method ? m : nullptr);
    if (rc != 0 || st[0] != 0) {
        if (!out) free(dst);
        return nullptr;
    }
    *out_size = osz[0];
    return dst;
}

extern "C" unsigned char* rans_uncompress_to_4x16(unsigned char* in, unsigned int in_size, unsigned char* out,
                                                  unsigned int* out_size) {
    return uncompress_one(in, in_size, out, out_size, HTS_B200_RANS4x16);
}
extern "C" unsigned char* rans_uncompress_4x16(unsigned char* in, unsigned int in_size, unsigned int* out_size) {
    return uncompress_one(in, in_size, nullptr, out_size, HTS_B200_RANS4x16);
}
extern "C" unsigned char* rans_uncompress(unsigned char* in, unsigned int in_size, unsigned int* out_size) {
    if (in_size < 9) return nullptr;                                 // rANS_static.c:938
    return uncompress_one(in, in_size, nullptr, out_size, HTS_B200_RANS4x8);
}

extern "C" unsigned char* rans_compress_to_4x16(unsigned char* in, unsigned int in_size, unsigned char* out,
                                                unsigned int* out_size, int order) {
    if (!out_size || (!in && in_size)) return nullptr;
    unsigned int bound = rans_compress_bound_4x16(in_size, order);
    if (out && *out_size < bound) return nullptr;                    // the reference's sub-encoders refuse too (:396)
    hts_b200_ctx* ctx = tls_ctx();
    if (!ctx) { fprintf(stderr, "htscodecs_b200: no usable sm_100 device (there is no CPU fallback)\n"); return nullptr; }
    unsigned char* dst = out ? out : (unsigned char*)malloc(bound);
    if (!dst) return nullptr;
    static unsigned char dummy = 0;
    const unsigned char* ins[1] = {in ? in : &dummy};
    unsigned char* outs[1] = {dst};
    unsigned int isz[1] = {in_size}, osz[1] = {bound};
    int st[1] = {0}, ord[1] = {order};
    int rc = run_ptr_batch(ctx, true, 1, ins, isz, outs, osz, ord, st, nullptr);
    if (rc != 0 || st[0] != 0) {
        if (!out) free(dst);
        return nullptr;
    }
    *out_size = osz[0];
    return dst;
}
extern "C" unsigned char* rans_compress_4x16(unsigned char* in, unsigned int in_size, unsigned int* out_size, int order) {
    return rans_compress_to_4x16(in, in_size, nullptr, out_size, order);
}

// the reference's own malloc size, rANS_static.c:87 (double arithmetic on purpose)
extern "C" unsigned int hts_b200_compress_bound_4x8(unsigned int size) {
    return (unsigned int)(1.05 * size + 257 * 257 * 3 + 9);
}

extern "C" unsigned char* rans_compress(unsigned char* in, unsigned int in_size, unsigned int* out_size, int order) {
    if (!out_size || !in || in_size == 0) return nullptr;
    hts_b200_ctx* ctx = tls_ctx();
    if (!ctx) { fprintf(stderr, "htscodecs_b200: no usable sm_100 device (there is no CPU fallback)\n"); return nullptr; }
    const unsigned int bound = hts_b200_compress_bound_4x8(in_size);
    unsigned char* dst = (unsigned char*)malloc(bound);
    if (!dst) return nullptr;
    const unsigned char* ins[1] = {in};
    unsigned char* outs[1] = {dst};
    unsigned int isz[1] = {in_size}, osz[1] = {bound};
    int st[1] = {0}, ord[1] = {(order ? 1 : 0) | HTS_B200_ORDER_RANS4x8};
    int rc = run_ptr_batch(ctx, true, 1, ins, isz, outs, osz, ord, st, nullptr);
    if (rc != 0 || st[0] != 0) { free(dst); return nullptr; }
    *out_size = osz[0];
    return dst;
}
