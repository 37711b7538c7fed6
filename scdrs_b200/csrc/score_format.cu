// Device side of the score-file writer (bin/scdrs:276-288 of the reference; format docs/file_format.rst:81-108):
// the n_cell x n_ctrl block of normalised control scores of a `.full_score.gz` file -- 1.1e8 numbers per trait at
// the default flags -- is turned into pandas' text ON the GPU, where the matrix already lies, instead of being
// downloaded as floats and printed by host threads (~0.1 us each).  Every value becomes one 16-byte slot,
//     bytes 0 .. len-1 : the shortest round-trip text of the float32 (score_format.cuh; at most 15 bytes, e.g.
//                        "-0.000123456784" or "-1.23456789e-10")
//     byte 15          : len
// written with one 128-bit store (consecutive threads, consecutive slots).  Fixed-size slots need no prefix sum
// of the text lengths on the device; the host side (score_writer.cu) copies 16 bytes per value and advances by
// len while it assembles the row blocks it deflates.
#include "common.cuh"
#include "score_format.cuh"

namespace scdrs {

__global__ void __launch_bounds__(256)
format_slots_kernel(int64_t n_val, const float* __restrict__ vals, uint4* __restrict__ slots) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_val; i += (int64_t)gridDim.x * blockDim.x) {
        union {
            uint4 q;
            char c[16];
        } s;
        s.q = make_uint4(0u, 0u, 0u, 0u);
        const int len = fmt_f32_shortest(vals[i], s.c);
        s.c[15] = (char)len;
        slots[i] = s.q;
    }
}

int format_text_slots(scdrs_ctx* c, int64_t n_val, const float* dev_vals, void* dev_slots) {
    if (n_val <= 0) return 0;
    int64_t grid = (n_val + 255) / 256;
    if (grid > kNumSM * 16) grid = kNumSM * 16;
    format_slots_kernel<<<(unsigned)grid, 256, 0, c->stream>>>(n_val, dev_vals, reinterpret_cast<uint4*>(dev_slots));
    SCDRS_LAUNCH_CHECK(c);
    return 0;
}

}  // namespace scdrs
