// decode.h -- host-side entry points of decode.cu
#pragma once
#include "common.cuh"

namespace hb {

struct DecodeBatch {
    DecWork* work;              // device
    const DecWork* hdr;         // host: initial header (counters zero, capacities, list pointers)
    const uint8_t* in_base;     // device
    const uint64_t* in_off;     // device
    const uint32_t* in_len;     // device
    uint8_t* out_base;          // device
    const uint64_t* out_off;    // device
    uint32_t* out_len;          // device (in: capacity, out: decoded size)
    int32_t* status;            // device
    const uint8_t* method;      // device or nullptr
    int nblk;
    uint32_t kinds;             // bit k set: job kind k may occur (host hint; ~0u = unknown)
    uint32_t post;              // bit0 RLE, bit1 PACK, bit2 STRIPE may occur
    bool big_batch = false;     // route small-alphabet 4-way order-0 streams to the compact-table kernels
    SideStreams* side = nullptr;   // nullptr: every kernel on the caller's stream, one after the other
    uint32_t hot = ~0u;            // kinds the previous batch on this context had jobs for (~0u: not known)
};

int decode_init(int device);
int decode_launch(const DecodeBatch& b, cudaStream_t st);   // returns kernels launched

}  // namespace hb
