// Self-test of the tcgen05 building blocks: D[128 x N] = A[128 x K] * B[N x K]^T with kind::tf32, one CTA,
// operands staged by plain loads into the K-major no-swizzle layout, accumulator in TMEM, read back with
// tcgen05.ld.  Exists so that descriptor / layout mistakes show up in a 2-second GPU test
// (tests/test_tc_gpu.py) and not inside the fused network kernels that reuse tc_common.cuh.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* sA = smem;                         // 128 x K tf32
    uint8_t* sB = smem + 128 * K * 4;           // N x K tf32
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;

    if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) tc::mbar_init(&bar, 1);
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i - r * K;
        *reinterpret_cast<float*>(sA + tc::kmajor_offset(r, k, K)) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i - r * K;
        *reinterpret_cast<float*>(sB + tc::kmajor_offset(r, k, K)) = B[i];
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N);
        const uint32_t sbo = (uint32_t)(K >> 2) * 128u;
        for (int kk = 0; kk < K / 8; ++kk) {
            const uint64_t da = tc::make_desc(tc::smem_u32(sA) + kk * 256, 128, sbo);
            const uint64_t db = tc::make_desc(tc::smem_u32(sB) + kk * 256, 128, sbo);
            tc::mma_tf32(tmem, da, db, idesc, kk > 0);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c + j < N) D[row * N + c + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

// Raw probe: shared-memory images and descriptor fields given by the caller (layout exploration / regression of the
// descriptor conventions documented in tc_common.cuh).  Finding recorded in tests/test_tc_gpu.py: with kind::tf32 and the
// no-swizzle layout the a_major / b_major ("MN-major") bits of the instruction descriptor make the MMA read zeros --
// transposed tf32 operands need the 128B_BASE32B swizzled layout -- so the backward kernels stage explicit transposed
// weight tiles instead.
__global__ void __launch_bounds__(128, 1)
tc_raw_selftest_kernel(const float* __restrict__ imgA, const float* __restrict__ imgB, float* __restrict__ D, int N, int steps,
                       int a_step, int a_lbo, int a_sbo, int b_step, int b_lbo, int b_sbo, int ta, int tb) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = sA + 128 * 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) tc::mbar_init(&bar, 1);
    for (int i = tid; i < 128 * 128; i += 128) { sA[i] = imgA[i]; sB[i] = imgB[i]; }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(128, N) | ((uint32_t)ta << 15) | ((uint32_t)tb << 16);
        for (int kk = 0; kk < steps; ++kk)
            tc::mma_tf32(tmem, tc::make_desc(tc::smem_u32(sA) + kk * a_step, a_lbo, a_sbo),
                         tc::make_desc(tc::smem_u32(sB) + kk * b_step, b_lbo, b_sbo), idesc, kk > 0);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c + j < N) D[row * N + c + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

// Lane probe: D = A[M x K] * B[N x K]^T with the instruction descriptor's M = 64 or 128, every TMEM lane pre-filled with
// a sentinel, then all 128 lanes x N columns dumped.  Records which lanes a 64-row accumulator occupies
// (tests/test_tc_gpu.py): rows 16q..16q+15 sit in lanes 32q..32q+15, the upper 16 lanes of each quarter are untouched.
__global__ void __launch_bounds__(128, 1)
tc_lane_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int M, int N, int K) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* sA = smem;                         // 128 x K tf32 (rows >= M zero)
    uint8_t* sB = smem + 128 * K * 4;           // N x K tf32
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) tc::mbar_init(&bar, 1);
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i - r * K;
        *reinterpret_cast<float*>(sA + tc::kmajor_offset(r, k, K)) = r < M ? A[i] : 0.f;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i - r * K;
        *reinterpret_cast<float*>(sB + tc::kmajor_offset(r, k, K)) = B[i];
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    float sentinel[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sentinel[j] = -12345.0f;
    for (int c = 0; c < N; c += 8) tc::tmem_st8(lane_addr + c, sentinel);
    tc::tmem_wait_st();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32(M, N);
        const uint32_t sbo = (uint32_t)(K >> 2) * 128u;
        for (int kk = 0; kk < K / 8; ++kk)
            tc::mma_tf32(tmem, tc::make_desc(tc::smem_u32(sA) + kk * 256, 128, sbo),
                         tc::make_desc(tc::smem_u32(sB) + kk * 256, 128, sbo), idesc, kk > 0);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tc::tmem_ld16(lane_addr + c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c + j < N) D[row * N + c + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

// Raw probe for kind::f16 (fp16 operands): 64 KB images, descriptor fields and the two major bits given by the caller; M = 64
// or 128.  All 128 TMEM lanes x N columns are dumped.  tests/test_tc_gpu.py records with it that ONE [rows][features] tile in
// the no-swizzle core-matrix layout is read K-major (features contracted) and MN-major (rows contracted) by swapping the
// leading / stride byte offsets -- what ppo_trunk's in-kernel weight gradients rest on.
__global__ void __launch_bounds__(128, 1)
tc_raw_f16_selftest_kernel(const uint32_t* __restrict__ imgA, const uint32_t* __restrict__ imgB, float* __restrict__ D, int M,
                           int N, int steps, int a_step, int a_lbo, int a_sbo, int b_step, int b_lbo, int b_sbo, int ta, int tb) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint32_t* sA = reinterpret_cast<uint32_t*>(smem);
    uint32_t* sB = sA + 16384;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) tc::mbar_init(&bar, 1);
    for (int i = tid; i < 16384; i += 128) { sA[i] = imgA[i]; sB[i] = imgB[i]; }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    float sentinel[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sentinel[j] = -12345.0f;
    for (int c = 0; c < N; c += 8) tc::tmem_st8(lane_addr + c, sentinel);
    tc::tmem_wait_st();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_f16(M, N, ta != 0, tb != 0);
        for (int kk = 0; kk < steps; ++kk)
            tc::mma_f16(tmem, tc::make_desc(tc::smem_u32(sA) + kk * a_step, a_lbo, a_sbo),
                        tc::make_desc(tc::smem_u32(sB) + kk * b_step, b_lbo, b_sbo), idesc, kk > 0);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tc::tmem_ld16(lane_addr + c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c + j < N) D[row * N + c + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

}  // namespace

extern "C" int mm_tc_raw_f16_selftest(const void* imgA, const void* imgB, float* D, int32_t M, int32_t N, int32_t steps,
                                      int32_t a_step, int32_t a_lbo, int32_t a_sbo, int32_t b_step, int32_t b_lbo, int32_t b_sbo,
                                      int32_t trans_a, int32_t trans_b, void* stream) {
    if (!imgA || !imgB || !D || (M != 64 && M != 128) || N < 16 || N > 256 || N % 16 || steps < 1 || steps > 16)
        return MM_ERR_BAD_ARG;
    const size_t smem = (size_t)2 * 65536;
    cudaError_t e = cudaFuncSetAttribute(tc_raw_f16_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    tc_raw_f16_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint32_t*>(imgA), static_cast<const uint32_t*>(imgB), D, M, N, steps, a_step, a_lbo, a_sbo, b_step,
        b_lbo, b_sbo, trans_a, trans_b);
    return mm::check_launch();
}

extern "C" int mm_tc_gemm_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, void* stream) {
    if (!A || !B || !D || N < 16 || N > 256 || N % 16 || K < 8 || K % 8 || K > 128) return MM_ERR_BAD_ARG;
    const size_t smem = (size_t)(128 + N) * K * 4;
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    tc_gemm_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, B, D, N, K);
    return mm::check_launch();
}

extern "C" int mm_tc_raw_selftest(const float* imgA, const float* imgB, float* D, int32_t N, int32_t steps, int32_t a_step,
                                  int32_t a_lbo, int32_t a_sbo, int32_t b_step, int32_t b_lbo, int32_t b_sbo, int32_t trans_a,
                                  int32_t trans_b, void* stream) {
    if (!imgA || !imgB || !D || N < 16 || N > 256 || N % 16 || steps < 1 || steps > 16) return MM_ERR_BAD_ARG;
    const size_t smem = (size_t)2 * 128 * 128 * 4;
    cudaError_t e = cudaFuncSetAttribute(tc_raw_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    tc_raw_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(imgA, imgB, D, N, steps, a_step, a_lbo, a_sbo,
                                                                                b_step, b_lbo, b_sbo, trans_a, trans_b);
    return mm::check_launch();
}

extern "C" int mm_tc_lane_probe(const float* A, const float* B, float* D, int32_t M, int32_t N, int32_t K, void* stream) {
    if (!A || !B || !D || (M != 64 && M != 128) || N < 16 || N > 256 || N % 16 || K < 8 || K % 8 || K > 128) return MM_ERR_BAD_ARG;
    const size_t smem = (size_t)(128 + N) * K * 4;
    cudaError_t e = cudaFuncSetAttribute(tc_lane_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    tc_lane_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(A, B, D, M, N, K);
    return mm::check_launch();
}
