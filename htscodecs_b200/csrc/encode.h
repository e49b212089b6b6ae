// encode.h -- host-side entry points of encode.cu
#pragma once
#include <stddef.h>
#include "common.cuh"

namespace hb {

struct EncodeBatch {
    const uint8_t* in_base;     // device
    const uint64_t* in_off;     // device
    const uint32_t* in_len;     // device
    uint8_t* out_base;          // device
    const uint64_t* out_off;    // device
    uint32_t* out_len;          // device (in: capacity, out: stream length)
    int32_t* status;            // device
    const int32_t* order;       // device
    int nblk;
};

// Device work area of one in-flight encode batch; grows on demand, reused across calls.
struct EncSlot {
    void* impl = nullptr;
    void release();
};

int encode_init(int device);
// Enqueue the encode pipeline on `st`.  h_in_len / h_order are host mirrors of the device arrays
// (nullptr: they are fetched with a small synchronous copy).  Returns kernels launched, or -1.
int encode_run(EncSlot& slot, const EncodeBatch& b, const uint32_t* h_in_len, const int32_t* h_order,
               cudaStream_t st, char* err, size_t errlen);
// Once the batch's stream has drained: true if it ran out of scratch arena (blocks then carry ST_ARENA); the slot's
// next encode_run allocates what was missing, so running the batch again succeeds.
bool encode_needs_retry(EncSlot& slot);
size_t encode_device_bytes(const EncSlot& slot);   // device memory the slot holds (scratch + descriptors + arena)

}  // namespace hb
