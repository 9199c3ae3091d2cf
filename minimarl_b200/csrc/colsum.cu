// Column sums of a tall row-major matrix: out[c] = sum_r x[r][c]  (bias gradients of the Linear / GRU layers,
// rows = batch x time up to 10^7, C <= 256).  Flat grid-stride reads with a stride that is a multiple of C, so a thread
// always sees the same column; four independent accumulators per thread keep enough loads in flight for HBM.
#include "common.cuh"

namespace {

constexpr int kMaxThreads = 256;

__global__ void __launch_bounds__(kMaxThreads) colsum_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                             long long n, int C, int tb) {
    __shared__ float sm[kMaxThreads];
    const int tid = threadIdx.x;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (tid < tb) {
        const long long stride = (long long)tb * gridDim.x;
        long long e = (long long)blockIdx.x * tb + tid;
        for (; e + 3 * stride < n; e += 4 * stride) {
            a0 += mm::ld_stream(x + e);
            a1 += mm::ld_stream(x + e + stride);
            a2 += mm::ld_stream(x + e + 2 * stride);
            a3 += mm::ld_stream(x + e + 3 * stride);
        }
        for (; e < n; e += stride) a0 += mm::ld_stream(x + e);
    }
    sm[tid] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (tid < C) {
        float s = 0.f;
        for (int i = tid; i < tb; i += C) s += sm[i];
        atomicAdd(out + tid, s);
    }
}

}  // namespace

// out [C] must be zeroed by the caller (the kernel accumulates).
extern "C" int mm_colsum(const float* x, float* out, int64_t rows, int32_t C, void* stream) {
    if (!x || !out || rows <= 0 || C <= 0) return MM_ERR_BAD_ARG;
    if (C > kMaxThreads) return MM_ERR_UNSUPPORTED;
    const int tb = C * (kMaxThreads / C);
    const long long n = rows * C;
    long long blocks = (n + (long long)tb * 16 - 1) / ((long long)tb * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    colsum_kernel<<<(unsigned)blocks, kMaxThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, out, n, C, tb);
    return mm::check_launch();
}
