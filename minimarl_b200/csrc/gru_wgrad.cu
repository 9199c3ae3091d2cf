// Recurrent weight gradients of the GRU: dW_hh = [dgi_r | dgi_z | dgh_n]^T h_prev and db_hh = column sums of the same gradients,
// contracted over all (step, row) pairs on tcgen05.
//
// Reference: the weight_hh_l0 / bias_hh_l0 gradients autograd produces for nn.GRU inside RNNLayer
// (offpolicy/utils/algorithm_utils/rnn.py:4-47) when QMix.train_policy_on_batch (offpolicy/algorithms/qmix/qmix.py:196) runs
// loss.backward().  The BPTT kernel (gru_tc.cu) emits dgi [T R, 3H] (the r and z columns are also dgh's) and dgh_n [T R, H]; the
// TF32 no-swizzle tiles it works on cannot be read MN-major, so the contraction over rows used to be two slab-split cuBLAS GEMMs
// plus two column-sum launches on a materialised shifted copy of h_seq.  Here: persistent CTAs stream 128-row tiles -- gradients
// and h_{t-1} rows (h_seq shifted by one step; h0 or zero for t = 0) -- convert them to fp16 operand tiles in the no-swizzle core-
// matrix layout (gradients scaled by a power of two), and one MN-major MMA group per tile accumulates
//     acc[lane = hidden unit k | lane H = ones][col = gate g] += [h_prev | 1]^T [dgi_r | dgi_z | dgh_n]
// in TMEM over all tiles of the CTA.  Two tile buffers: the MMA of tile i runs while tile i + 1 is loaded.  HBM-bound by design:
// 4 (3H + H) bytes per row in, nothing out but 148 accumulators.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128, kThreads = 512, H = 64, G3 = 3 * H;
constexpr int CH = kRows * 16, CHP = CH + 16;                 // chunk strides: h_prev tile / gradient tile (written 8 B at a time)
constexpr int HP_BYTES = 16 * CH, G_BYTES = 24 * CHP, STAGE = HP_BYTES + ((G_BYTES + 127) & ~127), S_BYTES = 2 * STAGE;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, int kk, uint32_t chunk_bytes) {
    return tc::make_desc(addr + (uint32_t)kk * 256u, 128u, chunk_bytes);
}

struct Args {
    const float* dgi;       // [T R][3H]
    const float* dghn;      // [T R][H]
    const float* h_seq;     // [T R][H]
    const float* h0;        // [R][H] or null (zero initial state)
    const float* scale;     // device float: power-of-two gradient scale
    float* acc;             // [grid][128][3H]
    long long rows, R, n_tiles;
};

__global__ void __launch_bounds__(kThreads, 1) gru_wgrad_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
    if (tid == 0) {
        tc::mbar_init(&bar[0], 1);
        tc::mbar_init(&bar[1], 1);
    }
    // chunk 8 of each h_prev tile: feature 0 = 1 for every row (bias gradients); chunks 9..15: zero
    for (int s = 0; s < 2; ++s)
        for (int i = tid; i < 8 * CH / 16; i += kThreads)
            reinterpret_cast<uint4*>(smem + s * STAGE + 8 * CH)[i] = (i < kRows) ? make_uint4(0x00003c00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = tc::make_idesc_f16(128, G3, true, true);
    const float gscale = a.scale[0];
    uint32_t par[2] = {0, 0};
    int committed[2] = {0, 0}, waited[2] = {0, 0};
    bool started = false;
    int it = 0;
    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        uint8_t* sHP = smem + b * STAGE;
        uint8_t* sG = sHP + HP_BYTES;
        const long long row0 = tile * kRows;
        const int n_valid = (int)min((long long)kRows, a.rows - row0);
        // ---- all 16 coalesced 16-byte loads of this thread first (gradients [128][192] = dgi[:, :2H] | dgh_n: 12; h_{t-1}: 4) ----------
        float4 v[12], w[4];
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            const int idx = u * kThreads + tid, r = idx / (G3 / 4), p = idx - r * (G3 / 4);
            const float* src = p < 2 * H / 4 ? a.dgi + (row0 + r) * G3 + 4 * p : a.dghn + (row0 + r) * H + 4 * (p - 2 * H / 4);
            v[u] = r < n_valid ? __ldcs(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {                    // global row i - R of h_seq, or h0 / zero for the first step
            const int idx = u * kThreads + tid, r = idx >> 4, p = idx & 15;
            const long long i = row0 + r;
            w[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < n_valid) {
                if (i >= a.R) w[u] = __ldcs(reinterpret_cast<const float4*>(a.h_seq + (i - a.R) * H + 4 * p));
                else if (a.h0) w[u] = __ldcs(reinterpret_cast<const float4*>(a.h0 + i * H + 4 * p));
            }
        }
        if (it >= 2) {                                   // the MMA that read this buffer two tiles ago has completed
            tc::mbar_wait(&bar[b], par[b]);
            par[b] ^= 1;
            ++waited[b];
        }
#pragma unroll
        for (int u = 0; u < 12; ++u) {
            const int idx = u * kThreads + tid, r = idx / (G3 / 4), p = idx - r * (G3 / 4);
            *reinterpret_cast<uint2*>(sG + (p >> 1) * CHP + r * 16 + (p & 1) * 8) =
                make_uint2(pack2(v[u].x * gscale, v[u].y * gscale), pack2(v[u].z * gscale, v[u].w * gscale));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = u * kThreads + tid, r = idx >> 4, p = idx & 15;
            *reinterpret_cast<uint2*>(sHP + (p >> 1) * CH + r * 16 + (p & 1) * 8) = make_uint2(pack2(w[u].x, w[u].y), pack2(w[u].z, w[u].w));
        }
        if (tid < kRows) *reinterpret_cast<uint32_t*>(sHP + 8 * CH + tid * 16) = tid < n_valid ? 0x00003c00u : 0u;   // ones column
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (tid == 0) {
            const uint32_t hp = tc::smem_u32(sHP), g = tc::smem_u32(sG);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) tc::mma_f16(tmem, desc_mn(hp, kk, CH), desc_mn(g, kk, CHP), idesc, started || kk > 0);
            tc::mma_commit(&bar[b]);
        }
        ++committed[b];
        started = true;
    }
    // drain: the MMA groups still outstanding on either buffer
    for (int b = 0; b < 2; ++b)
        for (; waited[b] < committed[b]; ++waited[b]) {
            tc::mbar_wait(&bar[b], par[b]);
            par[b] ^= 1;
        }
    tc::fence_after_sync();
    {
        const int row = (warp & 3) * 32 + (tid & 31), part = warp >> 2;          // 16 warps: lane quarter x column quarter
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        float* dst = a.acc + ((size_t)blockIdx.x * kRows + row) * G3 + part * (G3 / 4);
#pragma unroll 1
        for (int c = 0; c < G3 / 4; c += 8) {
            float v[8];
            tc::tmem_ld8(tl + part * (G3 / 4) + c, v);
            if (!started) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
            *reinterpret_cast<float4*>(dst + c) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// one thread per gradient element: dw_hh [3H][H] then db_hh [3H]
__global__ void __launch_bounds__(256) gru_wgrad_unfold_kernel(const float* __restrict__ acc, int n_ctas, const float* __restrict__ scale,
                                                                float* __restrict__ dw_hh, float* __restrict__ db_hh, int accumulate) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= G3 * H + G3) return;
    int lane, col;
    if (i < G3 * H) { col = i / H; lane = i - col * H; }
    else            { col = i - G3 * H; lane = H; }
    const float* p = acc + (size_t)lane * G3 + col;
    const size_t stride = (size_t)kRows * G3;
    float s[16];                                         // 16 independent loads in flight
#pragma unroll
    for (int u = 0; u < 16; ++u) s[u] = 0.f;
    int c = 0;
    for (; c + 15 < n_ctas; c += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u) s[u] += p[(c + u) * stride];
    }
    for (; c < n_ctas; ++c) s[0] += p[c * stride];
#pragma unroll
    for (int u = 8; u > 0; u >>= 1)
#pragma unroll
        for (int w = 0; w < u; ++w) s[w] += s[w + u];
    const float v = s[0] / scale[0];
    float* dst = i < G3 * H ? dw_hh + i : db_hh + col;
    *dst = accumulate ? *dst + v : v;
}

int n_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

extern "C" int mm_gru_wgrad_ctas(int64_t rows) {
    const long long tiles = (rows + kRows - 1) / kRows;
    return (int)(tiles < n_sms() ? tiles : n_sms());
}

// acc: scratch [n_acc_ctas >= mm_gru_wgrad_ctas(T R)][128][3H] floats.  scale: device float (mm_ppo_trunk2_scale of dgi).
// accumulate != 0: the gradients are ADDED to dw_hh / db_hh (the parameters' .grad storage) instead of written.
extern "C" int mm_gru_wgrad(const float* dgi, const float* dghn, const float* h_seq, const float* h0, const float* scale, float* acc,
                            int32_t n_acc_ctas, float* dw_hh, float* db_hh, int32_t accumulate, int32_t T, int64_t R, int32_t Hh,
                            void* stream) {
    if (!dgi || !dghn || !h_seq || !scale || !acc || !dw_hh || !db_hh || T <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (Hh != H) return MM_ERR_UNSUPPORTED;
    if (!mm::aligned(dgi, 16) || !mm::aligned(dghn, 16) || !mm::aligned(h_seq, 16) || (h0 && !mm::aligned(h0, 16)) || !mm::aligned(acc, 16))
        return MM_ERR_BAD_ARG;
    Args a{dgi, dghn, h_seq, h0, scale, acc, (long long)T * R, R, ((long long)T * R + kRows - 1) / kRows};
    const int grid = mm_gru_wgrad_ctas(a.rows);
    if (n_acc_ctas < grid) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaFuncSetAttribute(gru_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_BYTES);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    gru_wgrad_kernel<<<grid, kThreads, S_BYTES, st>>>(a);
    gru_wgrad_unfold_kernel<<<(G3 * H + G3 + 255) / 256, 256, 0, st>>>(acc, grid, scale, dw_hh, db_hh, accumulate);
    return mm::check_launch();
}
