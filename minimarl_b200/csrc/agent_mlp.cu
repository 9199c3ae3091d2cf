// The row-parallel half of the recurrent agent network's DIFFERENTIATED pass (the live network of a QMIX / VDN update):
//     LayerNorm(D) -> [Linear(D,64) ReLU LayerNorm] -> [Linear(64,64) ReLU LayerNorm] -> GRU input projection (64 -> 192)
// forward and backward as one kernel each, every contraction on tcgen05 (kind::f16: fp16 operands carry tf32's 11-bit
// significand, fp32 accumulate), every weight gradient contracted inside the backward kernel.
//
// Reference: MLPBase / MLPLayer (offpolicy/utils/algorithm_utils/mlp.py:7-89) and the input half of nn.GRU inside RNNLayer
// (rnn.py:4-47) as AgentQFunction.forward runs them (offpolicy/algorithms/qmix/algorithm/agent_q_function.py:34-66) in
// QMix.train_policy_on_batch (offpolicy/algorithms/qmix/qmix.py:80-210).  Rows are (t, episode, agent) entries, T x B x A of
// them, independent of each other -- the recurrence itself stays in csrc/gru_tc.cu, which takes this kernel's output gi
// [rows, 192] and hands back dgi.
//
// Same building blocks as csrc/ppo_trunk2.cu: thread = row (a thread reads its accumulator row with tcgen05.ld and does bias /
// ReLU / LayerNorm in registers), fp16 tiles in the no-swizzle core-matrix layout that serve as K-major operands when features
// are contracted and as MN-major operands when the 128 ROWS of the tile are (dW = G^T X, accumulated in TMEM over all tiles of
// the CTA, bias gradients through a ones chunk), LayerNorm affines folded into the following Linear by the caller.  Unlike
// the PPO trunk there is no time loop: a tile passes through once, so the kernels are plain persistent loops over tiles with
// one thread issuing the MMAs between block barriers; the forward runs two CTAs per SM to overlap them.
// The LayerNorm backward reads the normalised activations as the fp32 rows the forward saved (not the fp16 operand values):
// its projection cancels the component of dx along them, and rounding them would be amplified by |dx| / |dx_perp|.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;
constexpr int H = 64;
constexpr int G3 = 3 * H;                  // gate columns
constexpr int K1 = 48;
constexpr int CH = kRows * 16;             // bytes of one 8-feature chunk of a 128-row fp16 tile
constexpr int CHP = CH + 16;               // padded chunk stride of the gradient tile (written 8 bytes at a time by row-major lanes)
constexpr float kEps = 1e-5f;

// packed parameters (fp32, device, LayerNorm affines folded): w1 [H][K1] | w2 [H][H] | wih [3H][H] | b1 [H] | b2 [H] | bih [3H]
constexpr int P_W1 = 0, P_W2 = P_W1 + H * K1, P_WIH = P_W2 + H * H, P_B1 = P_WIH + G3 * H, P_B2 = P_B1 + H, P_BIH = P_B2 + H,
              P_TOTAL = P_BIH + G3;
// saved per row by the forward: xh1 [H] | xh2 [H] fp32 | 1/sigma1, 1/sigma2, relu1 (2 words), relu2 (2 words), 2 spare
constexpr int SV_ROW_FLOATS = 2 * H + 8;
// accumulator read-out of the backward: [128 lanes][ACC_COLS]: gi columns (192) | fc2 outputs (64) | fc1 outputs (64)
constexpr int ACC_COLS = G3 + H + H;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void store_chunk(uint8_t* tile, int chunk, int row, const float* v) {
    *reinterpret_cast<uint4*>(tile + chunk * CH + row * 16) =
        make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
__device__ __forceinline__ void stage_weight(uint8_t* dst, const float* src, int N, int K, int tid, int nthreads) {
    for (int i = tid; i < N * K; i += nthreads) {
        const int n = i / K, k = i - n * K;
        *reinterpret_cast<__half*>(dst + (k >> 3) * (N * 16) + n * 16 + (k & 7) * 2) = __float2half_rn(src[i]);
    }
}
__device__ __forceinline__ uint64_t desc_k(uint32_t addr, int kk, uint32_t chunk_bytes) {
    return tc::make_desc(addr + (uint32_t)kk * 2u * chunk_bytes, chunk_bytes, 128u);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, int kk, uint32_t chunk_bytes) {
    return tc::make_desc(addr + (uint32_t)kk * 256u, 128u, chunk_bytes);
}

template <int N>
__device__ __forceinline__ float sum_n(const float (&v)[N]) {
    float s0 = v[0], s1 = v[1], s2 = v[2], s3 = v[3];
#pragma unroll
    for (int j = 4; j < N; j += 4) { s0 += v[j]; s1 += v[j + 1]; s2 += v[j + 2]; s3 += v[j + 3]; }
    return (s0 + s1) + (s2 + s3);
}
template <int N>
__device__ __forceinline__ float dot_n(const float (&a)[N], const float (&b)[N]) {
    float s0 = a[0] * b[0], s1 = a[1] * b[1], s2 = a[2] * b[2], s3 = a[3] * b[3];
#pragma unroll
    for (int j = 4; j < N; j += 4) {
        s0 = fmaf(a[j], b[j], s0); s1 = fmaf(a[j + 1], b[j + 1], s1); s2 = fmaf(a[j + 2], b[j + 2], s2); s3 = fmaf(a[j + 3], b[j + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
// v -> LayerNorm(ReLU(v + b)) without affine, in place; relu[0..1] mark the positive pre-activations; returns 1 / sigma
__device__ __forceinline__ float bias_relu_norm(float (&v)[H], const float* b, uint32_t (&relu)[2]) {
    relu[0] = relu[1] = 0;
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int j = 0; j < H / 4; ++j) {
        const float4 q = b4[j];
        const float bb[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = 4 * j + u;
            const float x = v[k] + bb[u];
            const bool pos = x > 0.f;
            relu[k >> 5] |= (uint32_t)pos << (k & 31);
            v[k] = pos ? x : 0.f;
        }
    }
    const float mean = sum_n<H>(v) * (1.0f / H);
#pragma unroll
    for (int j = 0; j < H; ++j) v[j] -= mean;
    const float rstd = rsqrtf(dot_n<H>(v, v) * (1.0f / H) + kEps);
#pragma unroll
    for (int j = 0; j < H; ++j) v[j] *= rstd;
    return rstd;
}

struct Args {
    const float* params;
    const float* obs;       // [rows][D]
    float* gi;              // [rows][3H]                                         (forward)
    float* saved_rows;      // [rows][SV_ROW_FLOATS] or null (inference)          (forward writes, backward reads)
    __half* saved_x0;       // [n_tiles][6 chunks] fp16 images of the normalised observations, or null
    const float* dgi;       // [rows][3H]                                         (backward)
    const float* scale;     // device float: power-of-two gradient scale          (backward)
    float* acc;             // [grid][128][ACC_COLS]                              (backward)
    unsigned int* ticket;   // tile counter (zero at launch; the last CTA to leave resets it): tiles are handed out dynamically,
                            // so a CTA that starts late because another kernel holds its SM does not stretch the launch
    int D;
    long long rows, n_tiles;
};

// next tile of this CTA, the same value in every thread (one atomic per tile; the barrier also separates the tiles)
__device__ __forceinline__ long long next_tile(unsigned int* ticket, unsigned int* slot) {
    __syncthreads();
    if (threadIdx.x == 0) *slot = atomicAdd(ticket, 1u);
    __syncthreads();
    return (long long)*slot;
}
// the last CTA to run out of tiles puts the counter back to zero for the next launch
__device__ __forceinline__ void retire(unsigned int* ticket, unsigned int n_tiles) {
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == n_tiles + 2u * gridDim.x - 1u) *ticket = 0u;
    }
}

__device__ __forceinline__ void block_publish() {
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
}

// Thread = row makes every per-thread global access of a warp touch 32 different 128-byte lines (16 bytes of each): the L1 / LSU
// transaction rate, not HBM, then bounds the kernel (ncu: lg_throttle, long_scoreboard).  Rows therefore travel through a
// warp-private staging slab [32 rows][32 floats] (row stride 36 floats: conflict-free for 16-byte accesses both ways): the
// global side of a slab is 8 instructions of 4 rows x 128 contiguous bytes each.
constexpr int SLAB_STRIDE = 36, SLAB_FLOATS = 32 * SLAB_STRIDE;
// v: 32 consecutive columns of this lane's row  ->  global rows gdst + r * row_stride (r < n_rows), 32 floats each
__device__ __forceinline__ void slab_store(float* slab, const float* v, float* gdst, long long row_stride, int n_rows, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(slab + lane * SLAB_STRIDE + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int rr = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + rr;
        if (r < n_rows) mm::st_stream(reinterpret_cast<float4*>(gdst + r * row_stride + c4), *reinterpret_cast<const float4*>(slab + r * SLAB_STRIDE + c4));
    }
    __syncwarp();
}
// ======================================================= forward ==========================================================
// shared memory: W1 | W2 | Wih | biases | X0 tile (6 chunks) | XH1 tile (8) | XH2 tile (8)
constexpr int F_W1 = 0, F_W2 = F_W1 + H * K1 * 2, F_WIH = F_W2 + H * H * 2, F_BIAS = F_WIH + G3 * H * 2,
              F_X0 = F_BIAS + (2 * H + G3) * 4, F_XH1 = F_X0 + 6 * CH, F_XH2 = F_XH1 + 8 * CH, F_SLAB = F_XH2 + 8 * CH,
              F_BYTES = F_SLAB + 4 * SLAB_FLOATS * 4;

__global__ void __launch_bounds__(128, 2) agent_mlp_fwd_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot, tile_slot;
    const int tid = threadIdx.x, warp = tid >> 5, row = tid;
    float* sB = reinterpret_cast<float*>(smem + F_BIAS);
    uint8_t* sX0 = smem + F_X0;
    uint8_t* sXH1 = smem + F_XH1;
    uint8_t* sXH2 = smem + F_XH2;
    float* slab = reinterpret_cast<float*>(smem + F_SLAB) + warp * SLAB_FLOATS;
    const int lane = tid & 31;

    stage_weight(smem + F_W1, a.params + P_W1, H, K1, tid, 128);
    stage_weight(smem + F_W2, a.params + P_W2, H, H, tid, 128);
    stage_weight(smem + F_WIH, a.params + P_WIH, G3, H, tid, 128);
    for (int i = tid; i < 2 * H + G3; i += 128) sB[i] = a.params[P_B1 + i];
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
    if (tid == 0) tc::mbar_init(&bar, 1);
    block_publish();
    const uint32_t tmem = tmem_slot;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t w1 = tc::smem_u32(smem + F_W1), w2 = tc::smem_u32(smem + F_W2), wih = tc::smem_u32(smem + F_WIH);
    const uint32_t x0 = tc::smem_u32(sX0), xh1 = tc::smem_u32(sXH1), xh2 = tc::smem_u32(sXH2);
    const uint32_t id64 = tc::make_idesc_f16(128, H), id192 = tc::make_idesc_f16(128, G3);
    uint32_t par = 0;
    auto wait_mma = [&]() {
        tc::mbar_wait(&bar, par);
        par ^= 1;
        tc::fence_after_sync();
    };

    for (long long tile = next_tile(a.ticket, &tile_slot); tile < a.n_tiles; tile = next_tile(a.ticket, &tile_slot)) {
        const long long row0 = tile * kRows;
        const bool valid = row0 + row < a.rows;
        const long long grow = valid ? row0 + row : a.rows - 1;          // rows past the end shadow the last row
        const long long wrow0 = row0 + 32 * warp;                         // first row of this warp's slab
        const int wn = (int)max(0LL, min(32LL, a.rows - wrow0));          // valid rows in it
        // ---- LayerNorm(D) of the observation row (no affine) -> X0 tile ----------------------------------------------------
        {
            float x[K1];
            const float* src = a.obs + grow * a.D;
#pragma unroll
            for (int k = 0; k < K1; ++k) x[k] = k < a.D ? src[k] : 0.f;
            const float mean = sum_n<K1>(x) / a.D;
#pragma unroll
            for (int k = 0; k < K1; ++k) x[k] = k < a.D ? x[k] - mean : 0.f;
            const float rstd = rsqrtf(dot_n<K1>(x, x) / a.D + kEps);
#pragma unroll
            for (int k = 0; k < K1; ++k) x[k] *= rstd;
#pragma unroll
            for (int c = 0; c < K1 / 8; ++c) store_chunk(sX0, c, row, x + 8 * c);
        }
        block_publish();                        // also: the previous tile's bulk copy out of X0 has been waited for (below)
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) tc::mma_f16(tmem, desc_k(x0, kk, CH), desc_k(w1, kk, H * 16), id64, kk > 0);
            tc::mma_commit(&bar);
            if (a.saved_x0) {                   // the image of the normalised observations leaves as it is (for dW1)
                tc::bulk_s2g(reinterpret_cast<uint8_t*>(a.saved_x0) + (size_t)tile * 6 * CH, sX0, 6 * CH);
                tc::bulk_commit();
            }
        }
        wait_mma();
        // ---- fc1 epilogue -> xh1 ---------------------------------------------------------------------------------------------
        float rstd1, rstd2;
        uint32_t relu1[2], relu2[2];
        float* srow = a.saved_rows ? a.saved_rows + grow * SV_ROW_FLOATS : nullptr;
        {
            float v[H];
            tc::tmem_ld32_issue(tl, *reinterpret_cast<float(*)[32]>(v));
            tc::tmem_ld32_issue(tl + 32, *reinterpret_cast<float(*)[32]>(v + 32));
            tc::tmem_ld_fence<H>(v);
            rstd1 = bias_relu_norm(v, sB, relu1);
#pragma unroll
            for (int c = 0; c < H / 8; ++c) store_chunk(sXH1, c, row, v + 8 * c);
            if (a.saved_rows) {
                float* g0 = a.saved_rows + wrow0 * SV_ROW_FLOATS;
                slab_store(slab, v, g0, SV_ROW_FLOATS, wn, lane);
                slab_store(slab, v + 32, g0 + 32, SV_ROW_FLOATS, wn, lane);
            }
        }
        block_publish();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc::mma_f16(tmem + 64, desc_k(xh1, kk, CH), desc_k(w2, kk, H * 16), id64, kk > 0);
            tc::mma_commit(&bar);
        }
        wait_mma();
        // ---- fc2 epilogue -> xh2 ---------------------------------------------------------------------------------------------
        {
            float v[H];
            tc::tmem_ld32_issue(tl + 64, *reinterpret_cast<float(*)[32]>(v));
            tc::tmem_ld32_issue(tl + 96, *reinterpret_cast<float(*)[32]>(v + 32));
            tc::tmem_ld_fence<H>(v);
            rstd2 = bias_relu_norm(v, sB + H, relu2);
#pragma unroll
            for (int c = 0; c < H / 8; ++c) store_chunk(sXH2, c, row, v + 8 * c);
            if (a.saved_rows) {
                float* g0 = a.saved_rows + wrow0 * SV_ROW_FLOATS + H;
                slab_store(slab, v, g0, SV_ROW_FLOATS, wn, lane);
                slab_store(slab, v + 32, g0 + 32, SV_ROW_FLOATS, wn, lane);
            }
            if (srow && valid) {
                mm::st_stream(reinterpret_cast<float4*>(srow + 2 * H),
                              make_float4(rstd1, rstd2, __uint_as_float(relu1[0]), __uint_as_float(relu1[1])));
                mm::st_stream(reinterpret_cast<float4*>(srow + 2 * H) + 1,
                              make_float4(__uint_as_float(relu2[0]), __uint_as_float(relu2[1]), 0.f, 0.f));
            }
        }
        if (a.saved_x0 && tid == 0) tc::bulk_wait_read_all();          // X0 may be rewritten once every thread is past the barrier
        block_publish();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc::mma_f16(tmem, desc_k(xh2, kk, CH), desc_k(wih, kk, G3 * 16), id192, kk > 0);
            tc::mma_commit(&bar);
        }
        wait_mma();
        // ---- input projection of the GRU: + bias, rows out --------------------------------------------------------------------
        {
            float* out = a.gi + wrow0 * G3;
#pragma unroll 1
            for (int c = 0; c < G3 / 32; ++c) {
                float v[32];
                tc::tmem_ld32_issue(tl + 32 * c, v);
                tc::tmem_ld_fence<32>(v);
                const float4* b4 = reinterpret_cast<const float4*>(sB + 2 * H + 32 * c);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = b4[j];
                    v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                }
                slab_store(slab, v, out + 32 * c, G3, wn, lane);
            }
        }
        tc::fence_before_sync();                                         // orders this tile's TMEM reads before the next tile's MMAs
    }
    if (a.saved_x0 && tid == 0) tc::bulk_wait_all();
    retire(a.ticket, (unsigned int)a.n_tiles);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// ======================================================= backward =========================================================
// shared memory: Wih | W2 | ones chunk x3 interleaved with the activation tiles:
//   [XH2 8 chunks | ones | XH1 8 chunks | ones | X0 6 chunks | ones | pad] -- each activation tile followed by a chunk whose feature 0
//   is 1 for every row, so that the M = 128 (M = 64 for X0) MN-major operand starting at the tile also yields the bias gradient;
//   G tile: 24 chunks at the padded stride; dA2 tile, dA1 tile (8 chunks each).
constexpr int B_WIH = 0, B_W2 = B_WIH + G3 * H * 2, B_XH2 = B_W2 + H * H * 2, B_XH1 = B_XH2 + 9 * CH, B_X0 = B_XH1 + 9 * CH,
              B_G = B_X0 + 8 * CH, B_DA2 = B_G + 24 * CHP, B_DA1 = B_DA2 + 8 * CH, B_BYTES = B_DA1 + 8 * CH;
constexpr int T_WORK = 0, T_ACC_IH = 64, T_ACC_2 = T_ACC_IH + G3, T_ACC_1 = T_ACC_2 + H;      // TMEM columns

__global__ void __launch_bounds__(128, 1) agent_mlp_bwd_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar_x0;
    __shared__ uint32_t tmem_slot, tile_slot;
    const int tid = threadIdx.x, warp = tid >> 5, row = tid;
    uint8_t* sXH2 = smem + B_XH2;
    uint8_t* sXH1 = smem + B_XH1;
    uint8_t* sX0 = smem + B_X0;
    uint8_t* sG = smem + B_G;
    uint8_t* sDA2 = smem + B_DA2;
    uint8_t* sDA1 = smem + B_DA1;

    stage_weight(smem + B_WIH, a.params + P_WIH, G3, H, tid, 128);
    stage_weight(smem + B_W2, a.params + P_W2, H, H, tid, 128);
    {   // the three ones chunks (feature 0 = 1) and the pad chunk behind the last
        const uint4 one = make_uint4(0x00003c00u, 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(sXH2 + 8 * CH + row * 16) = one;
        *reinterpret_cast<uint4*>(sXH1 + 8 * CH + row * 16) = one;
        *reinterpret_cast<uint4*>(sX0 + 6 * CH + row * 16) = one;
        *reinterpret_cast<uint4*>(sX0 + 7 * CH + row * 16) = zero;
    }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init(&bar_x0, 1);
    }
    block_publish();
    const uint32_t tmem = tmem_slot;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t wih = tc::smem_u32(smem + B_WIH), w2 = tc::smem_u32(smem + B_W2);
    const uint32_t xh2 = tc::smem_u32(sXH2), xh1 = tc::smem_u32(sXH1), x0 = tc::smem_u32(sX0), g = tc::smem_u32(sG),
                   da2 = tc::smem_u32(sDA2), da1 = tc::smem_u32(sDA1);
    const uint32_t id_dx = tc::make_idesc_f16(128, H, false, true);            // A K-major, B = W read MN-major (W^T)
    const uint32_t idw192 = tc::make_idesc_f16(128, G3, true, true), idw64 = tc::make_idesc_f16(128, H, true, true),
                   idw64_m64 = tc::make_idesc_f16(64, H, true, true);          // rows contracted
    const float gscale = a.scale[0];
    uint32_t par = 0, par_x0 = 0;
    bool started = false;
    auto wait_mma = [&]() {
        tc::mbar_wait(&bar, par);
        par ^= 1;
        tc::fence_after_sync();
    };

    for (long long tile = next_tile(a.ticket, &tile_slot); tile < a.n_tiles; tile = next_tile(a.ticket, &tile_slot)) {
        const long long row0 = tile * kRows;
        const int n_valid = (int)min((long long)kRows, a.rows - row0);
        const bool valid = row < n_valid;
        const long long grow = row0 + min(row, n_valid - 1);
        // every MMA of the previous tile has completed (its last commit was waited for): the X0 image of this tile can land
        if (tid == 0) {
            tc::mbar_expect_tx(&bar_x0, 6 * CH);
            tc::bulk_g2s(sX0, reinterpret_cast<const uint8_t*>(a.saved_x0) + (size_t)tile * 6 * CH, 6 * CH, &bar_x0);
        }
        // ---- gradient tile: dgi rows (fp32, contiguous) -> scaled fp16 operand tile, 16 bytes of a row per lane ----------------
        {
            const float4* src = reinterpret_cast<const float4*>(a.dgi + row0 * G3);
            const int n_pieces = n_valid * (G3 / 4);
#pragma unroll 1
            for (int i0 = 0; i0 < kRows * (G3 / 4) / 128; i0 += 16) {          // 16 loads in flight per thread, then their 16 stores
                float4 v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int idx = (i0 + u) * 128 + tid;
                    v[u] = idx < n_pieces ? __ldcs(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int idx = (i0 + u) * 128 + tid, r = idx / (G3 / 4), p = idx - r * (G3 / 4);   // row r, features 4p .. 4p + 3
                    *reinterpret_cast<uint2*>(sG + (p >> 1) * CHP + r * 16 + (p & 1) * 8) =
                        make_uint2(pack2(v[u].x * gscale, v[u].y * gscale), pack2(v[u].z * gscale, v[u].w * gscale));
                }
            }
        }
        // ---- this thread's saved rows: xh2 -> operand tile (and kept for the LayerNorm backward) ----------------------------
        const float* srow = a.saved_rows + grow * SV_ROW_FLOATS;
        // (per-thread row reads: 34 independent 16-byte loads in flight; staging them through shared memory like the forward's
        // stores serialises four load -> store -> read round trips and was measured slower, 0.49 vs 0.41 ms)
        float xh[H];
#pragma unroll
        for (int j = 0; j < H / 4; ++j) {
            const float4 q = reinterpret_cast<const float4*>(srow + H)[j];
            xh[4 * j] = q.x; xh[4 * j + 1] = q.y; xh[4 * j + 2] = q.z; xh[4 * j + 3] = q.w;
        }
        const float4 e0 = reinterpret_cast<const float4*>(srow + 2 * H)[0], e1 = reinterpret_cast<const float4*>(srow + 2 * H)[1];
#pragma unroll
        for (int c = 0; c < H / 8; ++c) store_chunk(sXH2, c, row, xh + 8 * c);
        block_publish();
        if (tid == 0) {
            // dxh2 = dgi Wih' (K = 192: the weights read MN-major) ; dWih'^T += [xh2 | 1]^T dgi over the rows of the tile
#pragma unroll
            for (int kk = 0; kk < G3 / 16; ++kk) tc::mma_f16(tmem + T_WORK, desc_k(g, kk, CHP), desc_mn(wih, kk, G3 * 16), id_dx, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) tc::mma_f16(tmem + T_ACC_IH, desc_mn(xh2, kk, CH), desc_mn(g, kk, CHP), idw192, started || kk > 0);
            tc::mma_commit(&bar);
        }
        // meanwhile: xh1 of this row (registers until its LayerNorm backward)
        float xg[H];
#pragma unroll
        for (int j = 0; j < H / 4; ++j) {
            const float4 q = reinterpret_cast<const float4*>(srow)[j];
            xg[4 * j] = q.x; xg[4 * j + 1] = q.y; xg[4 * j + 2] = q.z; xg[4 * j + 3] = q.w;
        }
        wait_mma();
        // ---- dxh2 -> LayerNorm / ReLU backward -> dA2 --------------------------------------------------------------------------
        {
            float dx[H];
            tc::tmem_ld32_issue(tl + T_WORK, *reinterpret_cast<float(*)[32]>(dx));
            tc::tmem_ld32_issue(tl + T_WORK + 32, *reinterpret_cast<float(*)[32]>(dx + 32));
            tc::tmem_ld_fence<H>(dx);
            const float sg = sum_n<H>(dx) * (1.0f / H), sq = dot_n<H>(dx, xh) * (1.0f / H);
            const uint32_t m0 = __float_as_uint(e1.x), m1 = __float_as_uint(e1.y);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float v = e0.y * (dx[j] - sg - xh[j] * sq);
                dx[j] = (((j < 32 ? m0 : m1) >> (j & 31)) & 1u) && valid ? v : 0.f;
            }
#pragma unroll
            for (int c = 0; c < H / 8; ++c) store_chunk(sDA2, c, row, dx + 8 * c);
        }
#pragma unroll
        for (int c = 0; c < H / 8; ++c) store_chunk(sXH1, c, row, xg + 8 * c);
        block_publish();
        if (tid == 0) {
            // dxh1 = dA2 W2' ; dW2'^T += [xh1 | 1]^T dA2
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc::mma_f16(tmem + T_WORK, desc_k(da2, kk, CH), desc_mn(w2, kk, H * 16), id_dx, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) tc::mma_f16(tmem + T_ACC_2, desc_mn(xh1, kk, CH), desc_mn(da2, kk, CH), idw64, started || kk > 0);
            tc::mma_commit(&bar);
        }
        wait_mma();
        // ---- dxh1 -> LayerNorm / ReLU backward -> dA1 --------------------------------------------------------------------------
        {
            float dx[H];
            tc::tmem_ld32_issue(tl + T_WORK, *reinterpret_cast<float(*)[32]>(dx));
            tc::tmem_ld32_issue(tl + T_WORK + 32, *reinterpret_cast<float(*)[32]>(dx + 32));
            tc::tmem_ld_fence<H>(dx);
            const float sg = sum_n<H>(dx) * (1.0f / H), sq = dot_n<H>(dx, xg) * (1.0f / H);
            const uint32_t m0 = __float_as_uint(e0.z), m1 = __float_as_uint(e0.w);
#pragma unroll
            for (int j = 0; j < H; ++j) {
                const float v = e0.x * (dx[j] - sg - xg[j] * sq);
                dx[j] = (((j < 32 ? m0 : m1) >> (j & 31)) & 1u) && valid ? v : 0.f;
            }
#pragma unroll
            for (int c = 0; c < H / 8; ++c) store_chunk(sDA1, c, row, dx + 8 * c);
        }
        tc::mbar_wait(&bar_x0, par_x0);             // the X0 image has landed (async proxy -> async proxy: the MMA may read it)
        par_x0 ^= 1;
        block_publish();
        if (tid == 0) {
            // dW1'^T += [x0 | 1]^T dA1 (M = 64: 48 observation features, the ones chunk, one zero chunk)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) tc::mma_f16(tmem + T_ACC_1, desc_mn(x0, kk, CH), desc_mn(da1, kk, CH), idw64_m64, started || kk > 0);
            tc::mma_commit(&bar);
        }
        started = true;
        wait_mma();                                 // nothing of this tile is read by the tensor core any more
    }
    retire(a.ticket, (unsigned int)a.n_tiles);
    // ---- the weight-gradient accumulators leave TMEM: [128 lanes][ACC_COLS] per CTA -----------------------------------------------
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    {
        float* dst = a.acc + ((size_t)blockIdx.x * kRows + row) * ACC_COLS;
#pragma unroll 1
        for (int c = 0; c < ACC_COLS; c += 32) {
            float v[32];
            tc::tmem_ld32_issue(tl + T_ACC_IH + c, v);
            tc::tmem_ld_fence<32>(v);
            if (!started) {                       // this CTA never received a tile: its accumulators were never written
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(dst + c + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- the algebra around the two kernels, one launch each: folding the LayerNorm affines into the packed parameters, and summing
// the per-CTA accumulators + unfolding them into the gradients of the module's own tensors --------------------------------------
enum { Q_LN0W, Q_LN0B, Q_W1, Q_B1, Q_LN1W, Q_LN1B, Q_W2, Q_B2, Q_LN2W, Q_LN2B, Q_WIH, Q_BIH, Q_N };
struct ParamPtrs { const float* p[Q_N]; };

__device__ __forceinline__ float dot_row(const float* w, const float* v, int n) {
    float acc = 0.f;
    for (int k = 0; k < n; ++k) acc = fmaf(w[k], v[k], acc);
    return acc;
}

__global__ void __launch_bounds__(256) agent_mlp_pack_kernel(const ParamPtrs q, float* __restrict__ out, int D) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P_TOTAL) return;
    float v;
    if (i < P_W2) {
        const int n = i / K1, k = i - n * K1;
        v = k < D ? q.p[Q_W1][n * D + k] * q.p[Q_LN0W][k] : 0.f;
    } else if (i < P_WIH) {
        const int j = i - P_W2;
        v = q.p[Q_W2][j] * q.p[Q_LN1W][j & (H - 1)];
    } else if (i < P_B1) {
        const int j = i - P_WIH;
        v = q.p[Q_WIH][j] * q.p[Q_LN2W][j & (H - 1)];
    } else if (i < P_B2) {
        const int n = i - P_B1;
        v = q.p[Q_B1][n] + dot_row(q.p[Q_W1] + n * D, q.p[Q_LN0B], D);
    } else if (i < P_BIH) {
        const int n = i - P_B2;
        v = q.p[Q_B2][n] + dot_row(q.p[Q_W2] + n * H, q.p[Q_LN1B], H);
    } else {
        const int n = i - P_BIH;
        v = q.p[Q_BIH][n] + dot_row(q.p[Q_WIH] + n * H, q.p[Q_LN2B], H);
    }
    out[i] = v;
}

struct GradLayout { int off[Q_N + 1]; };
__host__ __device__ inline GradLayout grad_layout(int D) {
    const int n[Q_N] = {D, D, H * D, H, H, H, H * H, H, H, H, G3 * H, G3};
    GradLayout g;
    g.off[0] = 0;
    for (int i = 0; i < Q_N; ++i) g.off[i + 1] = g.off[i] + n[i];
    return g;
}


// every CTA sums a slice of the [n_ctas][128][ACC_COLS] accumulators into `red`; the last one turns them into gradients:
// W' = W diag(g), b' = b + W beta  =>  dW = dW' diag(g) + db' beta^T, dg = sum_n dW'[n,:] W[n,:], dbeta = W^T db'
struct GradPtrs { float* p[Q_N]; };

// Two launches: (1) every thread sums one accumulator slot over the CTAs (16 loads in flight); (2) one thread per GRADIENT ELEMENT
// turns the sums into the module's gradients.  (One kernel whose last CTA did step 2 alone spent 120 of its 140 us there.)
__global__ void __launch_bounds__(256) agent_mlp_reduce_kernel(const float* __restrict__ acc, int n_ctas, const float* __restrict__ scale,
                                                                float* __restrict__ red) {
    const float inv = 1.0f / scale[0];
    constexpr int kN = kRows * ACC_COLS;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < kN; i += gridDim.x * 256) {
        float s[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) s[u] = 0.f;
        int c = 0;
        for (; c + 15 < n_ctas; c += 16) {
#pragma unroll
            for (int u = 0; u < 16; ++u) s[u] += __ldcs(acc + (size_t)(c + u) * kN + i);
        }
        for (; c < n_ctas; ++c) s[0] += __ldcs(acc + (size_t)c * kN + i);
#pragma unroll
        for (int u = 8; u > 0; u >>= 1)
#pragma unroll
            for (int v = 0; v < u; ++v) s[v] += s[v + u];
        red[i] = s[0] * inv;
    }
}

// W' = W diag(g), b' = b + W beta  =>  dW = dW' diag(g) + db' beta^T, dg = sum_n dW'[n,:] W[n,:], dbeta = W^T db'
__global__ void __launch_bounds__(256) agent_mlp_unfold_kernel(const float* __restrict__ r, const ParamPtrs q, const GradPtrs gp,
                                                                int accumulate, int D) {
    const GradLayout L = grad_layout(D);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= L.off[Q_N]) return;
    int t = 0;
    while (i >= L.off[t + 1]) ++t;
    const int j = i - L.off[t];
    // T1: lanes xh2 feature | 1 [64] x gate column;  T2: lanes xh1 feature | 1 [64] x fc2 output;  T3: 64-row accumulator (row r in
    // lane 32 (r / 16) + r % 16): rows x0 feature | 1 [48] x fc1 output
    auto T1 = [&](int lane, int col) { return r[lane * ACC_COLS + col]; };
    auto T2 = [&](int lane, int col) { return r[lane * ACC_COLS + G3 + col]; };
    auto T3 = [&](int rr, int col) { return r[(32 * (rr >> 4) + (rr & 15)) * ACC_COLS + G3 + H + col]; };
    float v = 0.f;
    switch (t) {
        case Q_W1: { const int n = j / D, k = j - n * D; v = T3(k, n) * q.p[Q_LN0W][k] + T3(K1, n) * q.p[Q_LN0B][k]; break; }
        case Q_LN0W: for (int n = 0; n < H; ++n) v = fmaf(T3(j, n), q.p[Q_W1][n * D + j], v); break;
        case Q_LN0B: for (int n = 0; n < H; ++n) v = fmaf(T3(K1, n), q.p[Q_W1][n * D + j], v); break;
        case Q_B1: v = T3(K1, j); break;
        case Q_W2: { const int n = j / H, k = j & (H - 1); v = T2(k, n) * q.p[Q_LN1W][k] + T2(H, n) * q.p[Q_LN1B][k]; break; }
        case Q_B2: v = T2(H, j); break;
        case Q_LN1W: for (int n = 0; n < H; ++n) v = fmaf(T2(j, n), q.p[Q_W2][n * H + j], v); break;
        case Q_LN1B: for (int n = 0; n < H; ++n) v = fmaf(T2(H, n), q.p[Q_W2][n * H + j], v); break;
        case Q_WIH: { const int g = j / H, k = j & (H - 1); v = T1(k, g) * q.p[Q_LN2W][k] + T1(H, g) * q.p[Q_LN2B][k]; break; }
        case Q_BIH: v = T1(H, j); break;
        case Q_LN2W: for (int g = 0; g < G3; ++g) v = fmaf(T1(j, g), q.p[Q_WIH][g * H + j], v); break;
        default: for (int g = 0; g < G3; ++g) v = fmaf(T1(H, g), q.p[Q_WIH][g * H + j], v); break;      // Q_LN2B
    }
    // one destination per parameter tensor: a flat scratch buffer (=) or the parameters' own .grad storage (+=)
    gp.p[t][j] = accumulate ? gp.p[t][j] + v : v;
}

__device__ unsigned int g_ticket_fwd = 0, g_ticket_bwd = 0;

unsigned int* ticket_ptr(bool bwd) {
    void* p = nullptr;
    if (bwd) cudaGetSymbolAddress(&p, g_ticket_bwd);
    else     cudaGetSymbolAddress(&p, g_ticket_fwd);
    return static_cast<unsigned int*>(p);
}

int n_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

// dims[8]: packed parameter count, saved floats per row, accumulator columns, tile rows, K1, H, saved x0 halves per tile, 0
extern "C" int mm_agent_mlp_dims(int32_t* d) {
    if (!d) return MM_ERR_BAD_ARG;
    const int v[8] = {P_TOTAL, SV_ROW_FLOATS, ACC_COLS, kRows, K1, H, 6 * CH / 2, 0};
    for (int i = 0; i < 8; ++i) d[i] = v[i];
    return MM_OK;
}

static bool take_ptrs(const void* const* params12, ParamPtrs& q) {
    if (!params12) return false;
    for (int i = 0; i < Q_N; ++i) {
        if (!params12[i]) return false;
        q.p[i] = static_cast<const float*>(params12[i]);
    }
    return true;
}

// params12: HOST array of the 12 device pointers ln0_w, ln0_b, w1 [64][D], b1, ln1_w, ln1_b, w2, b2, ln2_w, ln2_b, w_ih [192][64], b_ih
extern "C" int mm_agent_mlp_pack(const void* const* params12, float* packed, int32_t D, void* stream) {
    ParamPtrs q;
    if (!take_ptrs(params12, q) || !packed || D < 1 || D > K1) return MM_ERR_BAD_ARG;
    agent_mlp_pack_kernel<<<(P_TOTAL + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, packed, D);
    return mm::check_launch();
}

extern "C" int mm_agent_mlp_grad_layout(int32_t D, int32_t* offsets13) {
    if (!offsets13 || D < 1 || D > K1) return MM_ERR_BAD_ARG;
    const GradLayout g = grad_layout(D);
    for (int i = 0; i <= Q_N; ++i) offsets13[i] = g.off[i];
    return MM_OK;
}

static int unfold_launch(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params12,
                         const GradPtrs& gp, int accumulate, int32_t D, void* stream) {
    ParamPtrs q;
    if (!take_ptrs(params12, q) || !acc || !scale || !red || n_ctas < 1 || D < 1 || D > K1) return MM_ERR_BAD_ARG;
    for (int i = 0; i < Q_N; ++i)
        if (!gp.p[i]) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    agent_mlp_reduce_kernel<<<(kRows * ACC_COLS + 255) / 256, 256, 0, st>>>(acc, n_ctas, scale, red);
    agent_mlp_unfold_kernel<<<(grad_layout(D).off[Q_N] + 255) / 256, 256, 0, st>>>(red, q, gp, accumulate, D);
    return mm::check_launch();
}

extern "C" int mm_agent_mlp_unfold(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params12,
                                   float* grads, int32_t D, void* stream) {
    if (!grads || D < 1 || D > K1) return MM_ERR_BAD_ARG;
    const GradLayout L = grad_layout(D);
    GradPtrs gp;
    for (int i = 0; i < Q_N; ++i) gp.p[i] = grads + L.off[i];
    return unfold_launch(acc, n_ctas, scale, red, params12, gp, 0, D, stream);
}

// the same, every gradient ADDED to its own destination tensor (grads12: HOST array of 12 device pointers, order of params12)
extern "C" int mm_agent_mlp_unfold_into(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params12,
                                        void* const* grads12, int32_t D, void* stream) {
    if (!grads12) return MM_ERR_BAD_ARG;
    GradPtrs gp;
    for (int i = 0; i < Q_N; ++i) gp.p[i] = static_cast<float*>(grads12[i]);
    return unfold_launch(acc, n_ctas, scale, red, params12, gp, 1, D, stream);
}

extern "C" int mm_agent_mlp_bwd_ctas(int64_t rows) {
    const long long tiles = (rows + kRows - 1) / kRows;
    return (int)(tiles < n_sms() ? tiles : n_sms());
}

extern "C" int mm_agent_mlp_fwd(const float* params, const float* obs, float* gi, float* saved_rows, void* saved_x0, int64_t rows,
                                int32_t D, void* stream) {
    if (!params || !obs || !gi || rows <= 0 || D < 1 || D > K1) return MM_ERR_BAD_ARG;
    if (!mm::aligned(gi, 16) || (saved_rows && !mm::aligned(saved_rows, 16)) || (saved_x0 && !mm::aligned(saved_x0, 16)) ||
        ((saved_rows == nullptr) != (saved_x0 == nullptr)))
        return MM_ERR_BAD_ARG;
    Args a{};
    a.params = params; a.obs = obs; a.gi = gi; a.saved_rows = saved_rows; a.saved_x0 = static_cast<__half*>(saved_x0);
    a.D = D; a.rows = rows; a.n_tiles = (rows + kRows - 1) / kRows;
    a.ticket = ticket_ptr(false);
    cudaError_t e = cudaFuncSetAttribute(agent_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_BYTES);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    const long long want = a.n_tiles, cap = 2LL * n_sms();
    agent_mlp_fwd_kernel<<<(unsigned)(want < cap ? want : cap), 128, F_BYTES, static_cast<cudaStream_t>(stream)>>>(a);
    return mm::check_launch();
}

extern "C" int mm_agent_mlp_bwd(const float* params, const float* saved_rows, const void* saved_x0, const float* dgi,
                                const float* scale, float* acc, int32_t n_acc_ctas, int64_t rows, void* stream) {
    if (!params || !saved_rows || !saved_x0 || !dgi || !scale || !acc || rows <= 0) return MM_ERR_BAD_ARG;
    if (!mm::aligned(saved_rows, 16) || !mm::aligned(saved_x0, 16) || !mm::aligned(dgi, 16) || !mm::aligned(acc, 16)) return MM_ERR_BAD_ARG;
    Args a{};
    a.params = params; a.saved_rows = const_cast<float*>(saved_rows);
    a.saved_x0 = static_cast<__half*>(const_cast<void*>(saved_x0)); a.dgi = dgi; a.scale = scale; a.acc = acc;
    a.rows = rows; a.n_tiles = (rows + kRows - 1) / kRows;
    a.ticket = ticket_ptr(true);
    const int grid = mm_agent_mlp_bwd_ctas(rows);
    if (n_acc_ctas < grid) return MM_ERR_BAD_ARG;
    cudaError_t e = cudaFuncSetAttribute(agent_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B_BYTES);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    agent_mlp_bwd_kernel<<<grid, 128, B_BYTES, static_cast<cudaStream_t>(stream)>>>(a);
    return mm::check_launch();
}
