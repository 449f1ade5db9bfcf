// gather.cuh -- the step behind the tokenizer: batches of the reference's DataCollection
// (genz_tokenize/models/bert/dataset.py:28-55: shuffle over the whole collection, then batch) cut out of planes that stay on
// the GPU.  ONE launch gathers the rows index[0..n_index) of every field: a warp per output row walks the fields and copies
// each row with 16-byte vectors (rows of the encoder's planes are whole vectors: 1 KB of ids, 256 B of mask / token types at
// max_len 256), bytewise where a field's rows are not.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gzt {

static const int GATHER_MAX_FIELDS = 8;
struct GatherArgs {
    const uint8_t* src[GATHER_MAX_FIELDS];
    uint8_t* dst[GATHER_MAX_FIELDS];
    uint32_t row_bytes[GATHER_MAX_FIELDS];
    int32_t n_fields;
    int64_t n_rows;           // rows of every field
    const int64_t* index;     // n_index row numbers; a number outside [0, n_rows) gives a row of zero bytes
    int64_t n_index;
};

__global__ void __launch_bounds__(256) k_gather_rows(GatherArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < A.n_index; i += nwarps) {
        const int64_t r = A.index[i];
        const bool ok = r >= 0 && r < A.n_rows;
        for (int f = 0; f < A.n_fields; f++) {
            const uint32_t rb = A.row_bytes[f];
            const uint8_t* s = A.src[f] + (ok ? r : 0) * (int64_t)rb;
            uint8_t* d = A.dst[f] + i * (int64_t)rb;
            if (((rb | (uint32_t)reinterpret_cast<uintptr_t>(A.src[f]) | (uint32_t)reinterpret_cast<uintptr_t>(A.dst[f])) & 15u) == 0) {
                for (uint32_t k = 16u * lane; k < rb; k += 512u)
                    __stcs(reinterpret_cast<uint4*>(d + k), ok ? __ldcs(reinterpret_cast<const uint4*>(s + k)) : make_uint4(0, 0, 0, 0));
            } else {
                for (uint32_t k = lane; k < rb; k += 32u) d[k] = ok ? s[k] : (uint8_t)0;
            }
        }
    }
}

}  // namespace gzt
