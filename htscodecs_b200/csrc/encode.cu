// encode.cu -- placeholder until the encode kernels land (next milestone).
#include <stdio.h>
#include "encode.h"
namespace hb {
void EncSlot::release() {}
int encode_init(int) { return 0; }
int encode_run(EncSlot&, const EncodeBatch&, const uint32_t*, const int32_t*, cudaStream_t, char* err, size_t errlen) {
    snprintf(err, errlen, "encode kernels not built yet");
    return -1;
}
}  // namespace hb
