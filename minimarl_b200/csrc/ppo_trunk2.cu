// Fused recurrent trunk of the PPO update, second generation: forward AND backward of
//     LayerNorm(D) -> [Linear ReLU LayerNorm] x2 -> GRU(32) (mask-reset) -> LayerNorm -> Linear head
// over data chunks (R_Actor.evaluate_actions / R_Critic.forward inside R_MAPPO.ppo_update:
// mappo/algorithms/r_actor_critic.py:95-146,189-208, mappo/utils/algorithm_utils/mlp.py, rnn.py:24-80,
// mappo/algorithms/ramppo_network.py:103-209), with EVERY weight gradient contracted inside the backward kernel.
//
// Design (what changed against ppo_trunk.cu, which spent its time in barriers around 8-column thread slices):
//   * thread = row.  A tile is 128 chunks and belongs to one warpgroup (128 threads); a thread reads the whole 32-column
//     accumulator row of its chunk with one tcgen05.ld and does bias / ReLU / LayerNorm / GRU gates / head in registers --
//     no cross-warp exchange, no block barrier in the time loop.
//   * several tiles in flight per SM: NSLOT warpgroups run independent tiles (own TMEM columns, own shared-memory tiles, own
//     mbarriers) and hide each other's MMA and memory latencies; ONE extra warp issues every tcgen05.mma of the CTA (it polls
//     the slots' "operands ready" mbarriers), so shared TMEM accumulators see a single in-order instruction stream.
//   * operands are fp16 (kind::f16; 11-bit significand = tf32's, fp32 accumulate; gradients are scaled by a power of two
//     chosen from max |dout| so that they sit in fp16's normal range).  With 16-bit operands ONE [rows][features] tile in
//     the no-swizzle core-matrix layout is a K-major operand when the features are contracted (activations x weights) and an
//     MN-major operand when the ROWS are contracted (profiles/probe_f16_layouts.py, tests/test_tc_gpu.py): the weight
//     gradients dW = G^T X are tcgen05.mma over the same tiles the data path uses, accumulated in TMEM over all tiles of the
//     CTA.  X and G never go to HBM; the transposed weights of the input-gradient products are MN-major reads as well.
//   * every LayerNorm affine is folded into the following Linear on the host (W' = W diag(g), b' = b + W beta), so the kernels
//     see plain normalised activations and the gains' gradients fall out of the weight-gradient blocks.
//   * LayerNorm(D) of the observations is computed ONCE per update into an fp16 row table (mm_ppo_obs_table) that both
//     networks and all epochs gather from with 16-byte cp.async straight into the operand layout.
//   * the forward saves, per (tile, step), the operand IMAGES the backward needs (xh1 | xh2 | hp tiles, 1/sigma and ReLU
//     bits) with one bulk shared->global copy; the backward brings them back with bulk global->shared copies.
//
// Tile layout ("chunk-major"): element (row r, feature f) of a 128-row fp16 tile sits at (f / 8) * 2048 + r * 16 + (f % 8) * 2.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;
constexpr int H = 32;
constexpr int K1 = 48;
constexpr int NOMAX = 8;
constexpr int CH = kRows * 16;            // bytes of one 8-feature chunk of a 128-row tile
constexpr float kEps = 1e-5f;

// device parameter buffer (fp32): w1 [H][K1] | w2 [H][H] | wih [3H][H] | whh [3H][H] | b1 [H] | b2 [H] | bih [3H] | bhh [3H] |
// wo [NOMAX][H] | bo [NOMAX]; every LayerNorm affine already folded in by the caller
constexpr int P_W1 = 0, P_W2 = P_W1 + H * K1, P_WIH = P_W2 + H * H, P_WHH = P_WIH + 3 * H * H, P_B1 = P_WHH + 3 * H * H,
              P_B2 = P_B1 + H, P_BIH = P_B2 + H, P_BHH = P_BIH + 3 * H, P_WO = P_BHH + 3 * H, P_BO = P_WO + NOMAX * H,
              P_TOTAL = P_BO + NOMAX;
// small parameters in shared memory (floats): b1 | b2 | br | bz | bin | bhn | wo | bo
constexpr int S_B1 = 0, S_B2 = 32, S_BR = 64, S_BZ = 96, S_BIN = 128, S_BHN = 160, S_WO = 192, S_BO = S_WO + NOMAX * H,
              S_TOTAL = S_BO + NOMAX;
constexpr int kSmallBytes = 2048;
static_assert(S_TOTAL * 4 <= kSmallBytes, "small parameters");

// saved block of one (tile, step): xh1 tile | [xh2 | hp] tile | per-row extras {1/sigma1, 1/sigma2, relu1 bits, relu2 bits}
constexpr int SV_XH1 = 0, SV_IN = 4 * CH, SV_EX = SV_IN + 8 * CH, SV_BYTES = SV_EX + kRows * 16;     // 26 624 B

// accumulator read-out of the backward: per CTA [128 lanes][ACC_COLS]
constexpr int ACC_A_COLS = 80, ACC_B_COLS = 96, ACC_COLS = ACC_A_COLS + ACC_B_COLS;

__device__ __forceinline__ float tanh_(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_(float x) { return fmaf(tanh_(0.5f * x), 0.5f, 0.5f); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bar_slot(int slot) { asm volatile("bar.sync %0, 128;" ::"r"(1 + slot) : "memory"); }

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
// eight consecutive features of this thread's row -> one 16-byte store into chunk `chunk` of a tile
__device__ __forceinline__ void store_chunk(uint8_t* tile, int chunk, int row, const float* v) {
    *reinterpret_cast<uint4*>(tile + chunk * CH + row * 16) =
        make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
__device__ __forceinline__ void load_chunk(const uint8_t* tile, int chunk, int row, float* v) {
    const uint4 q = *reinterpret_cast<const uint4*>(tile + chunk * CH + row * 16);
    const float2 a = unpack2(q.x), b = unpack2(q.y), c = unpack2(q.z), d = unpack2(q.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

// Row reductions as four independent chains: a 32-long dependent chain of FADD / FFMA is 128 cycles of pure latency for a warp
// that has one or two neighbours on its scheduler, and nvcc does not reassociate floating-point sums.
__device__ __forceinline__ float sum32(const float (&v)[H]) {
    float s0 = v[0], s1 = v[1], s2 = v[2], s3 = v[3];
#pragma unroll
    for (int j = 4; j < H; j += 4) { s0 += v[j]; s1 += v[j + 1]; s2 += v[j + 2]; s3 += v[j + 3]; }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float dot32(const float (&a)[H], const float (&b)[H]) {
    float s0 = a[0] * b[0], s1 = a[1] * b[1], s2 = a[2] * b[2], s3 = a[3] * b[3];
#pragma unroll
    for (int j = 4; j < H; j += 4) {
        s0 = fmaf(a[j], b[j], s0); s1 = fmaf(a[j + 1], b[j + 1], s1); s2 = fmaf(a[j + 2], b[j + 2], s2); s3 = fmaf(a[j + 3], b[j + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}
// v -> (v - mean) * rstd in place, returns rstd (biased variance, eps inside the root: torch.nn.LayerNorm)
__device__ __forceinline__ float normalise32(float (&v)[H]) {
    const float mean = sum32(v) * (1.0f / H);
#pragma unroll
    for (int j = 0; j < H; ++j) v[j] -= mean;
    const float rstd = rsqrtf(dot32(v, v) * (1.0f / H) + kEps);
#pragma unroll
    for (int j = 0; j < H; ++j) v[j] *= rstd;
    return rstd;
}
// v -> LayerNorm(ReLU(v + b)) without affine; the bits of `relu` mark the positive pre-activations
__device__ __forceinline__ float bias_relu_norm32(float (&v)[H], const float4* b4, uint32_t& relu) {
    relu = 0;
#pragma unroll
    for (int j = 0; j < H / 4; ++j) {
        const float4 b = b4[j];
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float x = v[4 * j + u] + bb[u];
            const bool pos = x > 0.f;
            relu |= (uint32_t)pos << (4 * j + u);
            v[4 * j + u] = pos ? x : 0.f;
        }
    }
    return normalise32(v);
}
// backward of y = normalise(a) followed by ReLU's mask: dx (gradient wrt y) -> gradient wrt the pre-activation, in place
__device__ __forceinline__ void norm_relu_bwd32(float (&dx)[H], const float (&xh)[H], float rstd, uint32_t relu) {
    const float sg = sum32(dx) * (1.0f / H), sq = dot32(dx, xh) * (1.0f / H);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const float v = rstd * (dx[j] - sg - xh[j] * sq);
        dx[j] = (relu >> j & 1u) ? v : 0.f;
    }
}
// head: out[o] = bo[o] + <xh3, wo[o]> for o < N (rows of wo past the real head width are zero), four chains per output
template <int N>
__device__ __forceinline__ void head32(const float (&xh3)[H], const float* sP, float* orow, int NO, bool valid) {
    const float4* w4 = reinterpret_cast<const float4*>(sP + S_WO);
    float acc[N][4];
#pragma unroll
    for (int o = 0; o < N; ++o) { acc[o][0] = sP[S_BO + o]; acc[o][1] = acc[o][2] = acc[o][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < H / 4; ++j)
#pragma unroll
        for (int o = 0; o < N; ++o) {
            const float4 w = w4[o * (H / 4) + j];
            acc[o][0] = fmaf(xh3[4 * j], w.x, acc[o][0]);
            acc[o][1] = fmaf(xh3[4 * j + 1], w.y, acc[o][1]);
            acc[o][2] = fmaf(xh3[4 * j + 2], w.z, acc[o][2]);
            acc[o][3] = fmaf(xh3[4 * j + 3], w.w, acc[o][3]);
        }
    if (valid) {
#pragma unroll
        for (int o = 0; o < N; ++o)
            if (o < NO) orow[o] = (acc[o][0] + acc[o][1]) + (acc[o][2] + acc[o][3]);
    }
}

// weight [N][K] fp32 (device) -> fp16 tile, chunk-major with N rows: (k / 8) * N * 16 + n * 16 + (k % 8) * 2
__device__ __forceinline__ void stage_weight(uint8_t* dst, const float* src, int N, int K, int tid, int nthreads) {
    for (int i = tid; i < N * K; i += nthreads) {
        const int n = i / K, k = i - n * K;
        *reinterpret_cast<__half*>(dst + (k >> 3) * (N * 16) + n * 16 + (k & 7) * 2) = __float2half_rn(src[i]);
    }
}
__device__ __forceinline__ void stage_small(float* sP, const float* p, int tid, int nthreads) {
    for (int i = tid; i < H; i += nthreads) {
        sP[S_B1 + i] = p[P_B1 + i];
        sP[S_B2 + i] = p[P_B2 + i];
        sP[S_BR + i] = p[P_BIH + i] + p[P_BHH + i];
        sP[S_BZ + i] = p[P_BIH + H + i] + p[P_BHH + H + i];
        sP[S_BIN + i] = p[P_BIH + 2 * H + i];
        sP[S_BHN + i] = p[P_BHH + 2 * H + i];
    }
    for (int i = tid; i < NOMAX * H; i += nthreads) sP[S_WO + i] = p[P_WO + i];
    for (int i = tid; i < NOMAX; i += nthreads) sP[S_BO + i] = p[P_BO + i];
}

// K-major reading of a tile whose feature chunks are `chunk_bytes` apart: k-step kk covers chunks 2 kk, 2 kk + 1
__device__ __forceinline__ uint64_t desc_k(uint32_t addr, int kk, uint32_t chunk_bytes) {
    return tc::make_desc(addr + (uint32_t)kk * 2u * chunk_bytes, chunk_bytes, 128u);
}
// MN-major reading (rows contracted): k-step kk covers rows 16 kk .. 16 kk + 15
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, int kk, uint32_t chunk_bytes) {
    return tc::make_desc(addr + (uint32_t)kk * 256u, 128u, chunk_bytes);
}

struct Args {
    const float* params;
    const __half* xhat;        // [rows][K1] normalised observations (mm_ppo_obs_table)
    const long long* first;    // [R] table row of every chunk's first step
    long long step_rows;       // table rows between consecutive steps of a chunk
    const float* h0;           // table [rows][H], row first[r]                                  (forward)
    const float* masks;        // table [rows], entry first[r] + t * step_rows, or null
    float* out;                // [L][R][NO]                                                     (forward)
    float* h_final;            // [R][H] or null                                                 (forward)
    uint8_t* saved;            // [n_tiles][L][SV_BYTES] or null (forward: inference)
    const float* dout;         // [L][R][NO]                                                     (backward)
    const float* scale;        // device float: power-of-two gradient scale                      (backward)
    float* acc;                // [grid][128][ACC_COLS]                                          (backward)
    int L, NO;
    int mask_in_table;         // masks ride in column K1 - 1 of xhat (D < K1): no separate 4-byte gather per row and step
    long long R, n_tiles;
};

// number of tiles slot `gs` (of `total_slots`) works on: gs, gs + total_slots, ...
__device__ __forceinline__ long long tiles_of_slot(long long n_tiles, long long gs, long long total_slots) {
    return n_tiles > gs ? (n_tiles - gs - 1) / total_slots + 1 : 0;
}

constexpr uint32_t kSpinLimit = 1u << 26;

// Debug record: a wait that runs into its bound writes {code, block, thread, 1} to a host-mapped int32[4] (when one is set with
// mm_ppo_trunk2_set_debug) before it traps, so that a wrong barrier protocol can be located without a debugger.
__device__ int* g_dbg = nullptr;
__device__ __noinline__ void dbg_trap(int code) {
    if (g_dbg) {
        g_dbg[0] = code; g_dbg[1] = blockIdx.x; g_dbg[2] = threadIdx.x; g_dbg[3] = 1;
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity, int code) {
    const uint32_t addr = tc::smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    dbg_trap(code);
}

// Register budget.  A CTA of NSLOT tile warpgroups + the issuer's warpgroup is launched with 65 536 / threads registers per thread
// (what __launch_bounds__ lets ptxas use; rounded down to 8).  That allocation is the CTA's pool: the issuer's warpgroup hands its
// share back (setmaxnreg.dec to 32) and the tile warpgroups grow into it (setmaxnreg.inc) -- asking for more than the pool holds
// would block forever.
constexpr int kIssuerRegs = 32;
constexpr int launch_regs(int nslot) {
    const int r = 65536 / ((nslot + 1) * 128) / 8 * 8;
    return r > 248 ? 248 : r;
}
constexpr int tile_regs(int nslot) {
    const int r = ((nslot + 1) * launch_regs(nslot) - kIssuerRegs) / nslot / 8 * 8;
    return r > 248 ? 248 : r;
}
static_assert(tile_regs(4) == 112 && tile_regs(3) == 160 && tile_regs(2) == 232 && tile_regs(1) == 248, "register pool");
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ======================================================= forward ==========================================================
// shared memory: W1 | W2 | Wih | Whh | small | per slot { X0 tile (6 chunks) | saved block | first-row table }
constexpr int F_W1 = 0, F_W2 = F_W1 + H * K1 * 2, F_WIH = F_W2 + H * H * 2, F_WHH = F_WIH + 3 * H * H * 2,
              F_SMALL = F_WHH + 3 * H * H * 2, F_SLOTS = F_SMALL + kSmallBytes;
constexpr int FS_X0 = 0, FS_SV = 6 * CH, FS_ROW = FS_SV + SV_BYTES, FS_BYTES = FS_ROW + kRows * 8;

template <int NSLOT>
__global__ void __launch_bounds__((NSLOT + 1) * 128, 1) trunk2_fwd_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_ready[NSLOT], bar_done[NSLOT];
    __shared__ uint32_t tmem_slot;
    constexpr int kThreads = (NSLOT + 1) * 128;
    constexpr uint32_t kTmemCols = NSLOT * 128 <= 128 ? 128 : NSLOT * 128 <= 256 ? 256 : 512;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* sP = reinterpret_cast<float*>(smem + F_SMALL);

    stage_weight(smem + F_W1, a.params + P_W1, H, K1, tid, kThreads);
    stage_weight(smem + F_W2, a.params + P_W2, H, H, tid, kThreads);
    stage_weight(smem + F_WIH, a.params + P_WIH, 3 * H, H, tid, kThreads);
    stage_weight(smem + F_WHH, a.params + P_WHH, 3 * H, H, tid, kThreads);
    stage_small(sP, a.params, tid, kThreads);
    if (warp == 0) tc::tmem_alloc(&tmem_slot, kTmemCols);
    if (tid == 0) {
        for (int s = 0; s < NSLOT; ++s) {
            tc::mbar_init(&bar_ready[s], 128);
            tc::mbar_init(&bar_done[s], 1);
        }
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const long long total_slots = (long long)gridDim.x * NSLOT;

    if (warp >= NSLOT * 4) {
        // ------------------------------------------------- MMA issuer ----------------------------------------------------
        reg_dec<kIssuerRegs>();
        // one issuer warp per slot (the warpgroup has four): it sleeps on its slot's "operands ready" barrier and answers with the
        // phase's MMAs and a commit -- no accumulator is shared between slots in the forward, so no common instruction stream
        const int s = warp - NSLOT * 4;
        if (s < NSLOT && lane == 0) {
            const uint32_t w1 = tc::smem_u32(smem + F_W1), w2 = tc::smem_u32(smem + F_W2), wih = tc::smem_u32(smem + F_WIH),
                           whh = tc::smem_u32(smem + F_WHH);
            const uint32_t id32 = tc::make_idesc_f16(128, 32), id64 = tc::make_idesc_f16(128, 64);
            const uint32_t base = tc::smem_u32(smem + F_SLOTS + s * FS_BYTES);
            const uint32_t x0 = base + FS_X0, xh1 = base + FS_SV + SV_XH1, xin = base + FS_SV + SV_IN;
            const uint32_t d = tmem + s * 128;
            const long long steps = tiles_of_slot(a.n_tiles, (long long)blockIdx.x * NSLOT + s, total_slots) * a.L;
            uint32_t par = 0;
            for (long long i = 0; i < steps; ++i) {
                wait_bar(&bar_ready[s], par, 110 + s);              // fc1: x0 [128 x 48] W1^T -> [0,32)
                par ^= 1;
                tc::fence_after_sync();
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) tc::mma_f16(d, desc_k(x0, kk, CH), desc_k(w1, kk, H * 16), id32, kk > 0);
                tc::mma_commit(&bar_done[s]);
                wait_bar(&bar_ready[s], par, 120 + s);              // fc2: xh1 W2^T -> [32,64)
                par ^= 1;
                tc::fence_after_sync();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d + 32, desc_k(xh1, kk, CH), desc_k(w2, kk, H * 16), id32, kk > 0);
                tc::mma_commit(&bar_done[s]);
                wait_bar(&bar_ready[s], par, 130 + s);              // gates: r|z -> [64,128), i_n -> [0,32), h_n -> [32,64)
                par ^= 1;
                tc::fence_after_sync();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d + 64, desc_k(xin, kk, CH), desc_k(wih, kk, 3 * H * 16), id64, kk > 0);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d + 64, desc_k(xin + 4 * CH, kk, CH), desc_k(whh, kk, 3 * H * 16), id64, true);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d, desc_k(xin, kk, CH), desc_k(wih + 2 * H * 16, kk, 3 * H * 16), id32, kk > 0);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d + 32, desc_k(xin + 4 * CH, kk, CH), desc_k(whh + 2 * H * 16, kk, 3 * H * 16), id32, kk > 0);
                tc::mma_commit(&bar_done[s]);
            }
        }
    } else {
        // ------------------------------------------------- tile warpgroups ----------------------------------------------
        reg_inc<tile_regs(NSLOT)>();
        const int s = warp >> 2, row = (warp & 3) * 32 + lane;
        uint8_t* slot = smem + F_SLOTS + s * FS_BYTES;
        uint8_t* sX0 = slot + FS_X0;
        uint8_t* sXH1 = slot + FS_SV + SV_XH1;
        uint8_t* sIN = slot + FS_SV + SV_IN;
        uint4* sEX = reinterpret_cast<uint4*>(slot + FS_SV + SV_EX);
        long long* sRow = reinterpret_cast<long long*>(slot + FS_ROW);
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + s * 128;
        uint32_t dpar = 0;
        auto ready = [&]() {
            tc::fence_smem_to_async();
            tc::fence_before_sync();
            tc::mbar_arrive(&bar_ready[s]);
        };
        auto wait_done = [&]() {
            wait_bar(&bar_done[s], dpar, 100 + s);
            dpar ^= 1;
            tc::fence_after_sync();
        };
        const float4* sP4 = reinterpret_cast<const float4*>(sP);

        for (long long tile = (long long)blockIdx.x * NSLOT + s; tile < a.n_tiles; tile += total_slots) {
            const long long row0 = tile * kRows;
            const int n_valid = (int)min((long long)kRows, a.R - row0);
            const bool valid = row < n_valid;
            const long long grow = row0 + min(row, n_valid - 1);         // rows past the end shadow the last row
            bar_slot(s);                                                 // previous tile's readers of sRow are done
            const long long my_first = a.first[grow];
            sRow[row] = my_first;
            bar_slot(s);
            auto fetch_x0 = [&](int t) {
                const long long toff = (long long)t * a.step_rows;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = i * 128 + row, rr = idx >> 3, c = idx & 7;
                    if (c < 6) cp_async16(sX0 + c * CH + rr * 16, a.xhat + (sRow[rr] + toff) * K1 + c * 8);
                }
                cp_async_commit();
            };
            fetch_x0(0);
            float h[H];
            {
                const float4* hp4 = reinterpret_cast<const float4*>(a.h0 + my_first * H);
#pragma unroll
                for (int j = 0; j < H / 4; ++j) {
                    const float4 q = hp4[j];
                    h[4 * j] = q.x; h[4 * j + 1] = q.y; h[4 * j + 2] = q.z; h[4 * j + 3] = q.w;
                }
            }
            auto load_mask = [&](int t) {
                return a.mask_in_table ? __half2float(a.xhat[(my_first + (long long)t * a.step_rows) * K1 + K1 - 1])
                                       : a.masks ? a.masks[my_first + (long long)t * a.step_rows] : 1.0f;
            };
            float mask_next = load_mask(0);
            for (int t = 0; t < a.L; ++t) {
                const float mask = mask_next;
                if (t + 1 < a.L) mask_next = load_mask(t + 1);            // one step ahead: its latency hides behind this step
                // ---- S0: masked input state -> [xh2 | hp] tile (chunks 4..7); observations landed -> fc1 -------------------
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        w[u] = pack2(h[8 * c + 2 * u] * mask, h[8 * c + 2 * u + 1] * mask);
                        const float2 f = unpack2(w[u]);                  // the state enters the cell as the fp16 value the MMA sees
                        h[8 * c + 2 * u] = f.x;
                        h[8 * c + 2 * u + 1] = f.y;
                    }
                    *reinterpret_cast<uint4*>(sIN + (4 + c) * CH + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                cp_async_wait_all();
                ready();
                wait_done();
                if (t + 1 < a.L) fetch_x0(t + 1);                         // fc1 has consumed the X0 tile
                // ---- S1: fc1 epilogue: bias, ReLU, LayerNorm -> xh1 tile -> fc2 ------------------------------------------
                float rstd1, rstd2;
                uint32_t relu1, relu2;
                {
                    float v[H];
                    tc::tmem_ld32_issue(tl, v);
                    tc::tmem_ld_fence<H>(v);
                    rstd1 = bias_relu_norm32(v, sP4 + S_B1 / 4, relu1);
#pragma unroll
                    for (int c = 0; c < 4; ++c) store_chunk(sXH1, c, row, v + 8 * c);
                }
                ready();
                wait_done();
                // ---- S2: fc2 epilogue -> xh2 (chunks 0..3 of the [xh2 | hp] tile) -> gate MMAs ---------------------------
                {
                    float v[H];
                    tc::tmem_ld32_issue(tl + 32, v);
                    tc::tmem_ld_fence<H>(v);
                    rstd2 = bias_relu_norm32(v, sP4 + S_B2 / 4, relu2);
#pragma unroll
                    for (int c = 0; c < 4; ++c) store_chunk(sIN, c, row, v + 8 * c);
                    sEX[row] = make_uint4(__float_as_uint(rstd1), __float_as_uint(rstd2), relu1, relu2);
                }
                ready();
                wait_done();
                // ---- S3: the step's operand images leave with one bulk copy; GRU cell, LayerNorm, head --------------------
                if (a.saved && row == 0) {
                    tc::bulk_s2g(a.saved + ((size_t)tile * a.L + t) * SV_BYTES, slot + FS_SV, SV_BYTES);
                    tc::bulk_commit();
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float gr[8], gz[8], gi[8], gh[8];
                    tc::tmem_ld8_issue(tl + 64 + 8 * c, gr);
                    tc::tmem_ld8_issue(tl + 96 + 8 * c, gz);
                    tc::tmem_ld8_issue(tl + 8 * c, gi);
                    tc::tmem_ld8_issue(tl + 32 + 8 * c, gh);
                    tc::tmem_ld_fence<8>(gr, gz, gi, gh);
                    float br[8], bz[8], bi[8], bh[8];
                    *reinterpret_cast<float4*>(br) = sP4[S_BR / 4 + 2 * c]; *reinterpret_cast<float4*>(br + 4) = sP4[S_BR / 4 + 2 * c + 1];
                    *reinterpret_cast<float4*>(bz) = sP4[S_BZ / 4 + 2 * c]; *reinterpret_cast<float4*>(bz + 4) = sP4[S_BZ / 4 + 2 * c + 1];
                    *reinterpret_cast<float4*>(bi) = sP4[S_BIN / 4 + 2 * c]; *reinterpret_cast<float4*>(bi + 4) = sP4[S_BIN / 4 + 2 * c + 1];
                    *reinterpret_cast<float4*>(bh) = sP4[S_BHN / 4 + 2 * c]; *reinterpret_cast<float4*>(bh + 4) = sP4[S_BHN / 4 + 2 * c + 1];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float r = sigmoid_(gr[j] + br[j]);
                        const float z = sigmoid_(gz[j] + bz[j]);
                        const float n = tanh_(gi[j] + bi[j] + r * (gh[j] + bh[j]));
                        h[8 * c + j] = fmaf(z, h[8 * c + j] - n, n);
                    }
                }
                {
                    float xh3[H];
#pragma unroll
                    for (int j = 0; j < H; ++j) xh3[j] = h[j];
                    normalise32(xh3);
                    float* orow = a.out + ((long long)t * a.R + grow) * a.NO;
                    if (a.NO == 1) head32<1>(xh3, sP, orow, a.NO, valid);
                    else if (a.NO <= 5) head32<5>(xh3, sP, orow, a.NO, valid);
                    else head32<NOMAX>(xh3, sP, orow, a.NO, valid);
                }
                if (a.saved) {                                          // the images were read before the next step rewrites them
                    if (row == 0) tc::bulk_wait_read_all();
                    bar_slot(s);
                }
            }
            if (a.h_final && valid) {
                float4* dst = reinterpret_cast<float4*>(a.h_final + grow * H);
#pragma unroll
                for (int j = 0; j < H / 4; ++j) dst[j] = make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
            }
        }
        if (a.saved && row == 0) tc::bulk_wait_all();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kTmemCols);
}

// ======================================================= backward =========================================================
// shared memory: Wih | Whh | W2 | ones tile (16 features, feature 0 = 1) | small | per slot
//   { BX = [x0 6 chunks | xh1 4 | dout 2] | [xh2 | hp] 8 chunks | extras | A2 = [dA1 4 | dA2 4 | xh3 4 | ones 1] | G4 16 chunks |
//     first-row table }.  A2 is read as a 16-chunk MN-major operand: its chunks 13..15 fall into G4 (finite or not, they only
//     reach accumulator lanes 104..127, which nobody reads).
constexpr int B_WIH = 0, B_WHH = B_WIH + 3 * H * H * 2, B_W2 = B_WHH + 3 * H * H * 2, B_ONES = B_W2 + H * H * 2,
              B_SMALL = B_ONES + 2 * CH, B_SLOTS = B_SMALL + kSmallBytes;
constexpr int BS_BX = 0, BS_XH1 = 6 * CH, BS_DOUT = 10 * CH, BS_IN = 12 * CH, BS_EX = BS_IN + 8 * CH, BS_A2 = BS_EX + kRows * 16,
              BS_G4 = BS_A2 + 13 * CH, BS_ROW = BS_G4 + 16 * CH, BS_BYTES = BS_ROW + kRows * 8;
constexpr int T_ACC_A = 256, T_ACC_B = T_ACC_A + ACC_A_COLS;

template <int NSLOT>
__global__ void __launch_bounds__((NSLOT + 1) * 128, 1) trunk2_bwd_kernel(const __grid_constant__ Args a) {
    static_assert(NSLOT <= 2, "TMEM: two working sets of 128 columns beside the weight-gradient accumulators");
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_ready[NSLOT][2], bar_done[NSLOT], bar_in[NSLOT], bar_x1[NSLOT], bar_final;
    __shared__ int4 ist[NSLOT];                  // issuer state per slot: {ready parity, phase, events left}
    __shared__ uint32_t tmem_slot;
    constexpr int kThreads = (NSLOT + 1) * 128;
    constexpr uint32_t kTmemCols = 512;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* sP = reinterpret_cast<float*>(smem + B_SMALL);

    stage_weight(smem + B_WIH, a.params + P_WIH, 3 * H, H, tid, kThreads);
    stage_weight(smem + B_WHH, a.params + P_WHH, 3 * H, H, tid, kThreads);
    stage_weight(smem + B_W2, a.params + P_W2, H, H, tid, kThreads);
    stage_small(sP, a.params, tid, kThreads);
    for (int i = tid; i < 2 * CH / 16; i += kThreads)                     // ones tile: feature 0 of every row
        *reinterpret_cast<uint4*>(smem + B_ONES + i * 16) = make_uint4(i < kRows ? 0x00003c00u : 0u, 0u, 0u, 0u);
    for (int i = tid; i < NSLOT * kRows; i += kThreads) {
        uint8_t* slot = smem + B_SLOTS + (i / kRows) * BS_BYTES;
        const int r = i % kRows;
        *reinterpret_cast<uint4*>(slot + BS_DOUT + CH + r * 16) = make_uint4(0u, 0u, 0u, 0x3c000000u);   // feature 15 = 1
        *reinterpret_cast<uint4*>(slot + BS_A2 + 12 * CH + r * 16) = make_uint4(0x00003c00u, 0u, 0u, 0u); // feature 96 = 1
    }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, kTmemCols);
    if (tid == 0) {
        for (int s = 0; s < NSLOT; ++s) {
            tc::mbar_init(&bar_ready[s][0], 128);
            tc::mbar_init(&bar_ready[s][1], 128);
            tc::mbar_init(&bar_done[s], 1);
            tc::mbar_init(&bar_in[s], 1);
            tc::mbar_init(&bar_x1[s], 1);
        }
        tc::mbar_init(&bar_final, 1);
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const long long total_slots = (long long)gridDim.x * NSLOT;

    if (warp >= NSLOT * 4) {
        // ------------------------------------------------- MMA issuer ----------------------------------------------------
        reg_dec<kIssuerRegs>();
        if (warp == NSLOT * 4 && lane == 0) {
            const uint32_t wih = tc::smem_u32(smem + B_WIH), whh = tc::smem_u32(smem + B_WHH), w2 = tc::smem_u32(smem + B_W2),
                           ones = tc::smem_u32(smem + B_ONES);
            const uint32_t id32 = tc::make_idesc_f16(128, 32), id64 = tc::make_idesc_f16(128, 64);
            const uint32_t id32_bt = tc::make_idesc_f16(128, 32, false, true);                 // B read MN-major (W^T)
            const uint32_t idw64 = tc::make_idesc_f16(128, 64, true, true), idw16 = tc::make_idesc_f16(128, 16, true, true),
                           idw96 = tc::make_idesc_f16(128, 96, true, true);                    // rows contracted
            int live = 0;
            bool acc_a = false, acc_b = false;
            for (int s = 0; s < NSLOT; ++s) {
                const int n = (int)tiles_of_slot(a.n_tiles, (long long)blockIdx.x * NSLOT + s, total_slots) * a.L * 4;
                ist[s] = make_int4(0, 0, n, 0);
                live += n > 0;
            }
            uint32_t spin = 0;
            while (live) {
#pragma unroll 1
                for (int s = 0; s < NSLOT; ++s) {
                    const int4 st = ist[s];
                    // the four phases of a step alternate between two "operands ready" barriers: the last phase of a step has
                    // no completion wait on the tile side, so the first phase of the next step can complete before this thread
                    // has seen it -- on one barrier that would flip the parity twice and both events would be lost
                    if (st.z > 0 && tc::mbar_test(&bar_ready[s][st.y & 1], (uint32_t)st.y >> 1)) {
                        tc::fence_after_sync();
                        const uint32_t base = tc::smem_u32(smem + B_SLOTS + s * BS_BYTES);
                        const uint32_t bx = base + BS_BX, xin = base + BS_IN, a2 = base + BS_A2, g4 = base + BS_G4;
                        const uint32_t d = tmem + s * 128;
                        if (st.y == 0) {            // gates from xh2 and hp: r|z -> [0,64), i_n -> [64,96), h_n -> [96,128)
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d, desc_k(xin, kk, CH), desc_k(wih, kk, 3 * H * 16), id64, kk > 0);
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d, desc_k(xin + 4 * CH, kk, CH), desc_k(whh, kk, 3 * H * 16), id64, true);
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d + 64, desc_k(xin, kk, CH), desc_k(wih + 2 * H * 16, kk, 3 * H * 16), id32, kk > 0);
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d + 96, desc_k(xin + 4 * CH, kk, CH), desc_k(whh + 2 * H * 16, kk, 3 * H * 16), id32, kk > 0);
                            tc::mma_commit(&bar_done[s]);
                        } else if (st.y == 1) {
                            // dxh2 = [dr dz dn] Wih -> [0,32);  dhp = [dr dz dn*r] Whh -> [32,64): the weights read MN-major
#pragma unroll
                            for (int kk = 0; kk < 6; ++kk) tc::mma_f16(d, desc_k(g4, kk, CH), desc_mn(wih, kk, 3 * H * 16), id32_bt, kk > 0);
#pragma unroll
                            for (int kk = 0; kk < 6; ++kk)
                                tc::mma_f16(d + 32, desc_k(g4, kk < 4 ? kk : kk + 2, CH), desc_mn(whh, kk, 3 * H * 16), id32_bt, kk > 0);
                            // weight gradients of the cell: G4^T [xh2 | hp] and G4^T 1 over the 128 rows of the tile
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk) {
                                tc::mma_f16(tmem + T_ACC_A, desc_mn(g4, kk, CH), desc_mn(xin, kk, CH), idw64, acc_a || kk > 0);
                                tc::mma_f16(tmem + T_ACC_A + 64, desc_mn(g4, kk, CH), desc_mn(ones, kk, CH), idw16, acc_a || kk > 0);
                            }
                            acc_a = true;
                            tc::mma_commit(&bar_done[s]);
                        } else if (st.y == 2) {     // dxh1 = dA2 W2 -> [64,96)
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) tc::mma_f16(d + 64, desc_k(a2 + 4 * CH, kk, CH), desc_mn(w2, kk, H * 16), id32_bt, kk > 0);
                            tc::mma_commit(&bar_done[s]);
                        } else {                    // [dA1 | dA2 | xh3 | 1]^T [x0 | xh1 | dout, 1]; the next commit covers it
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                tc::mma_f16(tmem + T_ACC_B, desc_mn(a2, kk, CH), desc_mn(bx, kk, CH), idw96, acc_b || kk > 0);
                            acc_b = true;
                        }
                        ist[s] = make_int4(0, (st.y + 1) & 3, st.z - 1, 0);
                        live -= st.z == 1;
                        spin = 0;
                    }
                }
                if (++spin > kSpinLimit) dbg_trap(999);
            }
            tc::mma_commit(&bar_final);
        }
    } else {
        // ------------------------------------------------- tile warpgroups ----------------------------------------------
        reg_inc<tile_regs(NSLOT)>();
        const int s = warp >> 2, row = (warp & 3) * 32 + lane;
        uint8_t* slot = smem + B_SLOTS + s * BS_BYTES;
        uint8_t* sX0 = slot + BS_BX;
        uint8_t* sXH1 = slot + BS_XH1;
        uint8_t* sDOUT = slot + BS_DOUT;
        uint8_t* sIN = slot + BS_IN;
        const uint4* sEX = reinterpret_cast<const uint4*>(slot + BS_EX);
        uint8_t* sA2 = slot + BS_A2;
        uint8_t* sG4 = slot + BS_G4;
        long long* sRow = reinterpret_cast<long long*>(slot + BS_ROW);
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + s * 128;
        const float4* sP4 = reinterpret_cast<const float4*>(sP);
        const float gscale = a.scale[0];
        uint32_t dpar = 0, in_par = 0, x1_par = 0;
        auto ready = [&](int ph) {
            tc::fence_smem_to_async();
            tc::fence_before_sync();
            tc::mbar_arrive(&bar_ready[s][ph & 1]);
        };
        auto wait_done = [&]() {
            wait_bar(&bar_done[s], dpar, 200 + s);
            dpar ^= 1;
            tc::fence_after_sync();
        };
        auto load_in = [&](long long tile, int t) {                      // [xh2 | hp] tile and the extras of (tile, t)
            tc::mbar_expect_tx(&bar_in[s], 8 * CH + kRows * 16);
            tc::bulk_g2s(sIN, a.saved + ((size_t)tile * a.L + t) * SV_BYTES + SV_IN, 8 * CH + kRows * 16, &bar_in[s]);
        };
        const long long tile0 = (long long)blockIdx.x * NSLOT + s;
        if (row == 0 && tile0 < a.n_tiles) load_in(tile0, a.L - 1);

        for (long long tile = tile0; tile < a.n_tiles; tile += total_slots) {
            const long long row0 = tile * kRows;
            const int n_valid = (int)min((long long)kRows, a.R - row0);
            const bool valid = row < n_valid;
            const long long grow = row0 + min(row, n_valid - 1);
            bar_slot(s);
            const long long my_first = a.first[grow];
            sRow[row] = my_first;
            bar_slot(s);
            float carry[H];                                               // gradient reaching h[t-1] from the steps after it
#pragma unroll
            for (int j = 0; j < H; ++j) carry[j] = 0.f;

            auto load_mask = [&](int t) {
                return a.mask_in_table ? __half2float(a.xhat[(my_first + (long long)t * a.step_rows) * K1 + K1 - 1])
                                       : a.masks ? a.masks[my_first + (long long)t * a.step_rows] : 1.0f;
            };
            auto load_dout = [&](int t, float (&d)[NOMAX]) {
                const float* drow = a.dout + ((long long)t * a.R + grow) * a.NO;
#pragma unroll
                for (int o = 0; o < NOMAX; ++o) d[o] = (valid && o < a.NO) ? drow[o] : 0.f;
            };
            float mask_next = load_mask(a.L - 1), dO_next[NOMAX];
            load_dout(a.L - 1, dO_next);

            for (int t = a.L - 1; t >= 0; --t) {
                const float mask = mask_next;
                float dO[NOMAX];
#pragma unroll
                for (int o = 0; o < NOMAX; ++o) dO[o] = dO_next[o] * gscale;
                if (t > 0) {                                              // one step ahead: the latency hides behind this step
                    mask_next = load_mask(t - 1);
                    load_dout(t - 1, dO_next);
                }
                // ---- B0: the step's [xh2 | hp] tile has landed -> recompute the gate pre-activations ---------------------
                wait_bar(&bar_in[s], in_par, 210 + s);
                in_par ^= 1;
                uint4 xh2p[4], hpp[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    xh2p[c] = *reinterpret_cast<const uint4*>(sIN + c * CH + row * 16);
                    hpp[c] = *reinterpret_cast<const uint4*>(sIN + (4 + c) * CH + row * 16);
                }
                const uint4 ex = sEX[row];
                ready(0);
                wait_done();          // also: the previous step's (or tile's) last weight-gradient MMAs have finished
                {                     // xh1 and the observation rows of this step; consumed at the end of the step
                    if (row == 0) {
                        tc::mbar_expect_tx(&bar_x1[s], 4 * CH);
                        tc::bulk_g2s(sXH1, a.saved + ((size_t)tile * a.L + t) * SV_BYTES + SV_XH1, 4 * CH, &bar_x1[s]);
                    }
                    const long long toff = (long long)t * a.step_rows;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int idx = i * 128 + row, rr = idx >> 3, c = idx & 7;
                        if (c < 6) cp_async16(sX0 + c * CH + rr * 16, a.xhat + (sRow[rr] + toff) * K1 + c * 8);
                    }
                    cp_async_commit();
                }
                // ---- B1: GRU cell forward (again), LayerNorm, head backward, cell backward -> G4 tile ---------------------
                uint32_t rz_[H], nh_[H];                                   // (r, z) and (n, gh_n + b) per column, fp16 pairs
                float hrow[H];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float gr[8], gz[8], gi[8], gh[8];
                    tc::tmem_ld8_issue(tl + 8 * c, gr);
                    tc::tmem_ld8_issue(tl + 32 + 8 * c, gz);
                    tc::tmem_ld8_issue(tl + 64 + 8 * c, gi);
                    tc::tmem_ld8_issue(tl + 96 + 8 * c, gh);
                    tc::tmem_ld_fence<8>(gr, gz, gi, gh);
                    float br[8], bz[8], bi[8], bh[8], hp[8];
                    *reinterpret_cast<float4*>(br) = sP4[S_BR / 4 + 2 * c]; *reinterpret_cast<float4*>(br + 4) = sP4[S_BR / 4 + 2 * c + 1];
                    *reinterpret_cast<float4*>(bz) = sP4[S_BZ / 4 + 2 * c]; *reinterpret_cast<float4*>(bz + 4) = sP4[S_BZ / 4 + 2 * c + 1];
                    *reinterpret_cast<float4*>(bi) = sP4[S_BIN / 4 + 2 * c]; *reinterpret_cast<float4*>(bi + 4) = sP4[S_BIN / 4 + 2 * c + 1];
                    *reinterpret_cast<float4*>(bh) = sP4[S_BHN / 4 + 2 * c]; *reinterpret_cast<float4*>(bh + 4) = sP4[S_BHN / 4 + 2 * c + 1];
                    {
                        const float2 p0 = unpack2(hpp[c].x), p1 = unpack2(hpp[c].y), p2 = unpack2(hpp[c].z), p3 = unpack2(hpp[c].w);
                        hp[0] = p0.x; hp[1] = p0.y; hp[2] = p1.x; hp[3] = p1.y; hp[4] = p2.x; hp[5] = p2.y; hp[6] = p3.x; hp[7] = p3.y;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float r = sigmoid_(gr[j] + br[j]);
                        const float z = sigmoid_(gz[j] + bz[j]);
                        const float hnb = gh[j] + bh[j];
                        const float n = tanh_(gi[j] + bi[j] + r * hnb);
                        hrow[8 * c + j] = fmaf(z, hp[j] - n, n);
                        rz_[8 * c + j] = pack2(r, z);
                        nh_[8 * c + j] = pack2(n, hnb);
                    }
                }
                const float rstd3 = normalise32(hrow);                    // hrow = xh3 from here on
#pragma unroll
                for (int c = 0; c < 4; ++c) store_chunk(sA2, 8 + c, row, hrow + 8 * c);
                *reinterpret_cast<uint4*>(sDOUT + row * 16) =
                    make_uint4(pack2(dO[0], dO[1]), pack2(dO[2], dO[3]), pack2(dO[4], dO[5]), pack2(dO[6], dO[7]));
                float g[H];                                               // d loss / d xh3 (the head's weights carry ln3's gain)
#pragma unroll
                for (int j = 0; j < H; ++j) g[j] = 0.f;
#pragma unroll
                for (int o = 0; o < NOMAX; ++o) {
                    if (o < a.NO) {
#pragma unroll
                        for (int j = 0; j < H / 4; ++j) {
                            const float4 w = sP4[S_WO / 4 + o * (H / 4) + j];
                            g[4 * j] = fmaf(dO[o], w.x, g[4 * j]);
                            g[4 * j + 1] = fmaf(dO[o], w.y, g[4 * j + 1]);
                            g[4 * j + 2] = fmaf(dO[o], w.z, g[4 * j + 2]);
                            g[4 * j + 3] = fmaf(dO[o], w.w, g[4 * j + 3]);
                        }
                    }
                }
                {
                    const float sg = sum32(g) * (1.0f / H), sq = dot32(g, hrow) * (1.0f / H);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float dr[8], dz[8], dn[8], dnr[8], hp[8];
                        {
                            const float2 p0 = unpack2(hpp[c].x), p1 = unpack2(hpp[c].y), p2 = unpack2(hpp[c].z), p3 = unpack2(hpp[c].w);
                            hp[0] = p0.x; hp[1] = p0.y; hp[2] = p1.x; hp[3] = p1.y; hp[4] = p2.x; hp[5] = p2.y; hp[6] = p3.x; hp[7] = p3.y;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int k = 8 * c + j;
                            const float dh = fmaf(rstd3, g[k] - sg - hrow[k] * sq, carry[k]);
                            const float2 rz = unpack2(rz_[k]), nh = unpack2(nh_[k]);
                            const float r = rz.x, z = rz.y, n = nh.x, hnb = nh.y;
                            dn[j] = dh * (1.f - z) * (1.f - n * n);
                            dz[j] = dh * (hp[j] - n) * z * (1.f - z);
                            dr[j] = dn[j] * hnb * r * (1.f - r);
                            dnr[j] = dn[j] * r;
                            carry[k] = dh * z;                            // direct path to hp; the matrix path is added in B2
                        }
                        store_chunk(sG4, c, row, dr);
                        store_chunk(sG4, 4 + c, row, dz);
                        store_chunk(sG4, 8 + c, row, dn);
                        store_chunk(sG4, 12 + c, row, dnr);
                    }
                }
                ready(1);
                wait_done();
                // the cell's MMAs are done with the [xh2 | hp] tile: bring in the one of the next step (or tile)
                if (row == 0) {
                    if (t > 0) load_in(tile, t - 1);
                    else if (tile + total_slots < a.n_tiles) load_in(tile + total_slots, a.L - 1);
                }
                // ---- B2: dxh2, dhp -> LayerNorm / ReLU backward -> dA2 ---------------------------------------------------
                {
                    float dx[H], dhp[H];
                    tc::tmem_ld32_issue(tl, dx);
                    tc::tmem_ld32_issue(tl + 32, dhp);
                    tc::tmem_ld_fence<H>(dx, dhp);
#pragma unroll
                    for (int j = 0; j < H; ++j) carry[j] = (carry[j] + dhp[j]) * mask;
                    float xh2[H];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float2 p0 = unpack2(xh2p[c].x), p1 = unpack2(xh2p[c].y), p2 = unpack2(xh2p[c].z), p3 = unpack2(xh2p[c].w);
                        xh2[8 * c] = p0.x; xh2[8 * c + 1] = p0.y; xh2[8 * c + 2] = p1.x; xh2[8 * c + 3] = p1.y;
                        xh2[8 * c + 4] = p2.x; xh2[8 * c + 5] = p2.y; xh2[8 * c + 6] = p3.x; xh2[8 * c + 7] = p3.y;
                    }
                    norm_relu_bwd32(dx, xh2, __uint_as_float(ex.y), ex.w);
#pragma unroll
                    for (int c = 0; c < 4; ++c) store_chunk(sA2, 4 + c, row, dx + 8 * c);
                }
                ready(2);
                wait_done();
                // ---- B3: dxh1 -> LayerNorm / ReLU backward -> dA1; then the step's weight-gradient MMAs -------------------
                wait_bar(&bar_x1[s], x1_par, 220 + s);
                x1_par ^= 1;
                {
                    float dx[H], xh1[H];
                    tc::tmem_ld32_issue(tl + 64, dx);
#pragma unroll
                    for (int c = 0; c < 4; ++c) load_chunk(sXH1, c, row, xh1 + 8 * c);
                    tc::tmem_ld_fence<H>(dx);
                    norm_relu_bwd32(dx, xh1, __uint_as_float(ex.x), ex.z);
#pragma unroll
                    for (int c = 0; c < 4; ++c) store_chunk(sA2, c, row, dx + 8 * c);
                }
                cp_async_wait_all();                                      // this thread's pieces of the observation tile
                ready(3);
            }
        }
    }
    // ---- every MMA of the CTA has finished: the weight-gradient accumulators leave TMEM -------------------------------------
    tc::fence_before_sync();
    __syncthreads();
    wait_bar(&bar_final, 0, 230);
    tc::fence_after_sync();
    if (warp < 4) {
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        float* dst = a.acc + ((size_t)blockIdx.x * kRows + warp * 32 + lane) * ACC_COLS;
#pragma unroll 1
        for (int c = 0; c < ACC_COLS; c += 16) {
            float v[16];
            tc::tmem_ld16(tl + T_ACC_A + c, v);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(dst + c + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kTmemCols);
}

// LayerNorm(D) without affine of every observation row, once per update: fp32 [rows][D] -> fp16 [rows][K1], zero padded
__global__ void __launch_bounds__(256) obs_table_kernel(const float* __restrict__ obs, const float* __restrict__ masks,
                                                        __half* __restrict__ out, long long rows, int D) {
    constexpr int RPW = 4;                               // rows per warp: eight loads in flight per lane instead of two
    const long long r0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW;
    const int lane = threadIdx.x & 31;
    if (r0 >= rows) return;
    float x0[RPW], x1[RPW], mk[RPW];
#pragma unroll
    for (int u = 0; u < RPW; ++u) {
        const long long r = min(r0 + u, rows - 1);
        const float* src = obs + r * D;
        x0[u] = lane < D ? __ldcs(src + lane) : 0.f;
        x1[u] = lane + 32 < D ? __ldcs(src + lane + 32) : 0.f;
        mk[u] = (masks && lane == K1 - 1 - 32) ? masks[r] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < RPW; ++u) {
        const long long r = r0 + u;
        float s = x0[u] + x1[u];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(MM_FULL_MASK, s, o);
        const float mean = s / D;
        const float d0 = lane < D ? x0[u] - mean : 0.f, d1 = lane + 32 < D ? x1[u] - mean : 0.f;
        float q = d0 * d0 + d1 * d1;
#pragma unroll
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(MM_FULL_MASK, q, o);
        const float rstd = rsqrtf(q / D + kEps);
        if (r < rows) {
            __half* dst = out + r * K1;
            dst[lane] = __float2half_rn(d0 * rstd);
            if (lane + 32 < K1) dst[lane + 32] = (masks && lane + 32 == K1 - 1) ? __float2half_rn(mk[u]) : __float2half_rn(d1 * rstd);
        }
    }
}

// ---- host-side algebra as two launches: folding the LayerNorm affines into the kernel parameters, and (after the backward)
// summing the per-CTA accumulators and unfolding them into the gradients of the module's own parameters ----------------------
// parameter tensors in the order of nn/ppo_trunk.py PARAM_ORDER
enum { Q_LN0W, Q_LN0B, Q_W1, Q_B1, Q_LN1W, Q_LN1B, Q_W2, Q_B2, Q_LN2W, Q_LN2B, Q_WIH, Q_WHH, Q_BIH, Q_BHH, Q_LN3W, Q_LN3B, Q_WO, Q_BO, Q_N };
struct ParamPtrs { const float* p[Q_N]; };

__device__ __forceinline__ float dot_row(const float* w, const float* v, int n) {
    float acc = 0.f;
    for (int k = 0; k < n; ++k) acc = fmaf(w[k], v[k], acc);
    return acc;
}

__global__ void __launch_bounds__(256) trunk2_pack_kernel(const ParamPtrs q, float* __restrict__ out, int D, int NO) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P_TOTAL) return;
    float v;
    if (i < P_W2) {
        const int n = i / K1, k = i - n * K1;
        v = k < D ? q.p[Q_W1][n * D + k] * q.p[Q_LN0W][k] : 0.f;
    } else if (i < P_WIH) {
        const int j = i - P_W2;
        v = q.p[Q_W2][j] * q.p[Q_LN1W][j & (H - 1)];
    } else if (i < P_WHH) {
        const int j = i - P_WIH;
        v = q.p[Q_WIH][j] * q.p[Q_LN2W][j & (H - 1)];
    } else if (i < P_B1) {
        v = q.p[Q_WHH][i - P_WHH];
    } else if (i < P_B2) {
        const int n = i - P_B1;
        v = q.p[Q_B1][n] + dot_row(q.p[Q_W1] + n * D, q.p[Q_LN0B], D);
    } else if (i < P_BIH) {
        const int n = i - P_B2;
        v = q.p[Q_B2][n] + dot_row(q.p[Q_W2] + n * H, q.p[Q_LN1B], H);
    } else if (i < P_BHH) {
        const int n = i - P_BIH;
        v = q.p[Q_BIH][n] + dot_row(q.p[Q_WIH] + n * H, q.p[Q_LN2B], H);
    } else if (i < P_WO) {
        v = q.p[Q_BHH][i - P_BHH];
    } else if (i < P_BO) {
        const int j = i - P_WO, o = j / H;
        v = o < NO ? q.p[Q_WO][j] * q.p[Q_LN3W][j & (H - 1)] : 0.f;
    } else {
        const int o = i - P_BO;
        v = o < NO ? q.p[Q_BO][o] + dot_row(q.p[Q_WO] + o * H, q.p[Q_LN3B], H) : 0.f;
    }
    out[i] = v;
}

// gradient buffer: the eighteen tensors of PARAM_ORDER back to back (sizes depend on D and NO)
struct GradLayout { int off[Q_N + 1]; };
__host__ __device__ inline GradLayout grad_layout(int D, int NO) {
    const int n[Q_N] = {D, D, H * D, H, H, H, H * H, H, H, H, 3 * H * H, 3 * H * H, 3 * H, 3 * H, H, H, NO * H, NO};
    GradLayout g;
    g.off[0] = 0;
    for (int i = 0; i < Q_N; ++i) g.off[i + 1] = g.off[i] + n[i];
    return g;
}


// grid x 256 threads: every CTA sums a slice of the [n_ctas][128][ACC_COLS] accumulators into `red`; the CTA that finishes
// last (ticket counter) turns them into the parameter gradients (1 / scale applied, LayerNorm affines unfolded:
// W' = W diag(g), b' = b + W beta  =>  dW = dW' diag(g) + db' beta^T, dg = sum_n dW'[n,:] W[n,:], dbeta = W^T db')
struct GradPtrs { float* p[Q_N]; };

// Two launches: (1) every thread sums one accumulator slot over the CTAs (16 loads in flight); (2) one thread per GRADIENT ELEMENT
// turns the sums into the module's gradients (a single kernel whose last CTA did step 2 alone spent most of its 40 us there).
__global__ void __launch_bounds__(256) trunk2_reduce_kernel(const float* __restrict__ acc, int n_ctas, const float* __restrict__ scale,
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

__global__ void __launch_bounds__(256) trunk2_unfold_kernel(const float* __restrict__ r, const ParamPtrs q, const GradPtrs gp,
                                                             int accumulate, int D, int NO) {
    const GradLayout L = grad_layout(D, NO);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= L.off[Q_N]) return;
    int t = 0;
    while (i >= L.off[t + 1]) ++t;
    const int j = i - L.off[t];
    auto A = [&](int lane, int col) { return r[lane * ACC_COLS + col]; };
    auto B = [&](int lane, int col) { return r[lane * ACC_COLS + ACC_A_COLS + col]; };
    auto whh_lane = [](int g) { return g < 2 * H ? g : g + H; };      // dr | dz | (dn_i) | dn_h
    float v = 0.f;
    switch (t) {
        case Q_W1: { const int n = j / D, k = j - n * D; v = B(n, k) * q.p[Q_LN0W][k] + B(n, 95) * q.p[Q_LN0B][k]; break; }
        case Q_LN0W: for (int n = 0; n < H; ++n) v = fmaf(B(n, j), q.p[Q_W1][n * D + j], v); break;
        case Q_LN0B: for (int n = 0; n < H; ++n) v = fmaf(B(n, 95), q.p[Q_W1][n * D + j], v); break;
        case Q_B1: v = B(j, 95); break;
        case Q_W2: { const int n = j / H, k = j & (H - 1); v = B(H + n, 48 + k) * q.p[Q_LN1W][k] + B(H + n, 95) * q.p[Q_LN1B][k]; break; }
        case Q_B2: v = B(H + j, 95); break;
        case Q_LN1W: for (int n = 0; n < H; ++n) v = fmaf(B(H + n, 48 + j), q.p[Q_W2][n * H + j], v); break;
        case Q_LN1B: for (int n = 0; n < H; ++n) v = fmaf(B(H + n, 95), q.p[Q_W2][n * H + j], v); break;
        case Q_WIH: { const int g = j / H, k = j & (H - 1); v = A(g, k) * q.p[Q_LN2W][k] + A(g, 64) * q.p[Q_LN2B][k]; break; }
        case Q_WHH: { const int g = j / H, k = j & (H - 1); v = A(whh_lane(g), H + k); break; }
        case Q_BIH: v = A(j, 64); break;
        case Q_BHH: v = A(whh_lane(j), 64); break;
        case Q_LN2W: for (int g = 0; g < 3 * H; ++g) v = fmaf(A(g, j), q.p[Q_WIH][g * H + j], v); break;
        case Q_LN2B: for (int g = 0; g < 3 * H; ++g) v = fmaf(A(g, 64), q.p[Q_WIH][g * H + j], v); break;
        case Q_LN3W: for (int o = 0; o < NO; ++o) v = fmaf(B(2 * H + j, 80 + o), q.p[Q_WO][o * H + j], v); break;
        case Q_LN3B: for (int o = 0; o < NO; ++o) v = fmaf(B(96, 80 + o), q.p[Q_WO][o * H + j], v); break;
        case Q_WO: { const int o = j / H, k = j & (H - 1); v = B(2 * H + k, 80 + o) * q.p[Q_LN3W][k] + B(96, 80 + o) * q.p[Q_LN3B][k]; break; }
        default: v = B(96, 80 + j); break;                                                                // Q_BO
    }
    // one destination tensor per parameter: a flat scratch buffer (=), or the parameters' own .grad storage (+=: what autograd's
    // AccumulateGrad would do with the values one launch per tensor later)
    gp.p[t][j] = accumulate ? gp.p[t][j] + v : v;
}

// power-of-two gradient scale from max |dout|: scale2[0] = 2^floor(log2(64 / max)), scale2[1] = bits of the maximum (zeroed by the
// caller's memset); one pass over dout, the CTA that finishes last writes the scale
__device__ unsigned int g_scale_ticket = 0;
__global__ void __launch_bounds__(256) trunk2_scale_kernel(const float* __restrict__ dout, long long n, float* __restrict__ scale2) {
    __shared__ bool last;
    unsigned int m = 0;
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = reinterpret_cast<const float4*>(dout)[i];
        m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
    }
    if (blockIdx.x == 0)
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += 256) m = max(m, __float_as_uint(fabsf(dout[i])));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(MM_FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(reinterpret_cast<unsigned int*>(scale2) + 1, m);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&g_scale_ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last || threadIdx.x) return;
    __threadfence();
    g_scale_ticket = 0;
    const float amax = __uint_as_float(*(reinterpret_cast<volatile unsigned int*>(scale2) + 1));
    int e = (amax > 0.f && amax < 3.0e38f) ? ilogbf(64.0f / fmaxf(amax, 1e-30f)) : 0;     // non-finite gradients stay non-finite
    e = e < -60 ? -60 : e > 60 ? 60 : e;
    scale2[0] = ldexpf(1.0f, e);
}

int n_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

int g_fwd_slots = 4, g_bwd_slots = 2;

template <int NSLOT>
int launch_fwd(const Args& a, cudaStream_t st) {
    const size_t smem = F_SLOTS + (size_t)NSLOT * FS_BYTES;
    cudaError_t e = cudaFuncSetAttribute(trunk2_fwd_kernel<NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    const long long want = (a.n_tiles + NSLOT - 1) / NSLOT;
    const int grid = (int)(want < n_sms() ? want : n_sms());
    trunk2_fwd_kernel<NSLOT><<<grid, (NSLOT + 1) * 128, smem, st>>>(a);
    return mm::check_launch();
}
template <int NSLOT>
int launch_bwd(const Args& a, int grid, cudaStream_t st) {
    const size_t smem = B_SLOTS + (size_t)NSLOT * BS_BYTES;
    cudaError_t e = cudaFuncSetAttribute(trunk2_bwd_kernel<NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    trunk2_bwd_kernel<NSLOT><<<grid, (NSLOT + 1) * 128, smem, st>>>(a);
    return mm::check_launch();
}

}  // namespace

// dims[8]: parameter count, saved bytes per (tile, step), accumulator columns, A columns, tile rows, K1, NOMAX, H
extern "C" int mm_ppo_trunk2_dims(int32_t* d) {
    if (!d) return MM_ERR_BAD_ARG;
    const int v[8] = {P_TOTAL, SV_BYTES, ACC_COLS, ACC_A_COLS, kRows, K1, NOMAX, H};
    for (int i = 0; i < 8; ++i) d[i] = v[i];
    return MM_OK;
}

// tuning: tiles in flight per SM (forward 1..4, backward 1..2); 0 keeps the current value
extern "C" int mm_ppo_trunk2_set_slots(int32_t fwd, int32_t bwd) {
    if (fwd < 0 || fwd > 4 || bwd < 0 || bwd > 2) return MM_ERR_BAD_ARG;
    if (fwd) g_fwd_slots = fwd;
    if (bwd) g_bwd_slots = bwd;
    return MM_OK;
}

// debug: host-mapped (pinned) int32[4] that a timed-out wait fills before trapping; NULL = off
extern "C" int mm_ppo_trunk2_set_debug(void* host_mapped_int4) {
    int* p = static_cast<int*>(host_mapped_int4);
    cudaError_t e = cudaMemcpyToSymbol(g_dbg, &p, sizeof(p));
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    return MM_OK;
}

extern "C" int mm_ppo_trunk2_bwd_ctas(int64_t R) {
    const long long tiles = (R + kRows - 1) / kRows, want = (tiles + g_bwd_slots - 1) / g_bwd_slots;
    return (int)(want < n_sms() ? want : n_sms());
}

// params18: HOST array of the 18 device pointers of nn/ppo_trunk.py PARAM_ORDER (ln0_w, ln0_b, w1, b1, ln1_w, ln1_b, w2, b2, ln2_w,
// ln2_b, wih, whh, bih, bhh, ln3_w, ln3_b, wo, bo; w1 is [32][D], wo [NO][32]) -> packed kernel parameters (dims[0] floats)
extern "C" int mm_ppo_trunk2_pack(const void* const* params18, float* packed, int32_t D, int32_t NO, void* stream) {
    if (!params18 || !packed || D < 1 || D > K1 || NO < 1 || NO > NOMAX) return MM_ERR_BAD_ARG;
    ParamPtrs q;
    for (int i = 0; i < Q_N; ++i) {
        if (!params18[i]) return MM_ERR_BAD_ARG;
        q.p[i] = static_cast<const float*>(params18[i]);
    }
    trunk2_pack_kernel<<<(P_TOTAL + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, packed, D, NO);
    return mm::check_launch();
}

// scale2: device float[2]; [0] receives the power of two that puts max |dout| at 64 (what mm_ppo_trunk2_bwd takes as `scale`)
extern "C" int mm_ppo_trunk2_scale(const float* dout, int64_t n, float* scale2, void* stream) {
    if (!dout || !scale2 || n <= 0 || !mm::aligned(dout, 16)) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(scale2, 0, 8, st);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    const long long want = (n / 4 + 255) / 256;
    trunk2_scale_kernel<<<(unsigned)(want < 2 * n_sms() ? (want ? want : 1) : 2 * n_sms()), 256, 0, st>>>(dout, n, scale2);
    return mm::check_launch();
}

// number of floats of the gradient buffer mm_ppo_trunk2_unfold fills, and the 19 offsets of its 18 tensors
extern "C" int mm_ppo_trunk2_grad_layout(int32_t D, int32_t NO, int32_t* offsets19) {
    if (!offsets19 || D < 1 || D > K1 || NO < 1 || NO > NOMAX) return MM_ERR_BAD_ARG;
    const GradLayout g = grad_layout(D, NO);
    for (int i = 0; i <= Q_N; ++i) offsets19[i] = g.off[i];
    return MM_OK;
}

// acc [n_ctas][128][dims[2]] from mm_ppo_trunk2_bwd -> gradients of the module parameters (PARAM_ORDER, back to back in grads);
// red: scratch of 128 * dims[2] floats
static int unfold_launch(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params18,
                         const GradPtrs& gp, int accumulate, int32_t D, int32_t NO, void* stream) {
    if (!acc || !scale || !red || !params18 || n_ctas < 1 || D < 1 || D > K1 || NO < 1 || NO > NOMAX) return MM_ERR_BAD_ARG;
    ParamPtrs q;
    for (int i = 0; i < Q_N; ++i) {
        if (!params18[i] || !gp.p[i]) return MM_ERR_BAD_ARG;
        q.p[i] = static_cast<const float*>(params18[i]);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    trunk2_reduce_kernel<<<(kRows * ACC_COLS + 255) / 256, 256, 0, st>>>(acc, n_ctas, scale, red);
    trunk2_unfold_kernel<<<(grad_layout(D, NO).off[Q_N] + 255) / 256, 256, 0, st>>>(red, q, gp, accumulate, D, NO);
    return mm::check_launch();
}

extern "C" int mm_ppo_trunk2_unfold(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params18,
                                    float* grads, int32_t D, int32_t NO, void* stream) {
    if (!grads || D < 1 || D > K1 || NO < 1 || NO > NOMAX) return MM_ERR_BAD_ARG;
    const GradLayout g = grad_layout(D, NO);
    GradPtrs gp;
    for (int i = 0; i < Q_N; ++i) gp.p[i] = grads + g.off[i];
    return unfold_launch(acc, n_ctas, scale, red, params18, gp, 0, D, NO, stream);
}

// the same, every gradient ADDED to its own destination tensor (grads18: HOST array of 18 device pointers, PARAM_ORDER): the
// parameters' .grad storage, so that no per-tensor accumulation launch follows
extern "C" int mm_ppo_trunk2_unfold_into(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params18,
                                         void* const* grads18, int32_t D, int32_t NO, void* stream) {
    if (!grads18) return MM_ERR_BAD_ARG;
    GradPtrs gp;
    for (int i = 0; i < Q_N; ++i) gp.p[i] = static_cast<float*>(grads18[i]);
    return unfold_launch(acc, n_ctas, scale, red, params18, gp, 1, D, NO, stream);
}

extern "C" int mm_ppo_obs_table(const float* obs, const float* masks, void* xhat, int64_t rows, int32_t D, void* stream) {
    if (!obs || !xhat || rows <= 0 || D < 1 || D > K1) return MM_ERR_BAD_ARG;
    if (!mm::aligned(xhat, 16) || (masks && D >= K1)) return MM_ERR_BAD_ARG;
    obs_table_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(obs, masks, static_cast<__half*>(xhat),
                                                                                                  rows, D);
    return mm::check_launch();
}

extern "C" int mm_ppo_trunk2_fwd(const float* params, const void* xhat, const int64_t* first, int64_t step_rows, const float* h0,
                                 const float* masks, int32_t mask_in_table, float* out, float* h_final, void* saved, int32_t L,
                                 int64_t R, int32_t NO, void* stream) {
    if (!params || !xhat || !first || !h0 || !out || L <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (NO < 1 || NO > NOMAX) return MM_ERR_UNSUPPORTED;
    if (!mm::aligned(xhat, 16) || !mm::aligned(h0, 16) || (saved && !mm::aligned(saved, 16)) ||
        (h_final && !mm::aligned(h_final, 16)))
        return MM_ERR_BAD_ARG;
    Args a{};
    a.params = params; a.xhat = static_cast<const __half*>(xhat); a.first = reinterpret_cast<const long long*>(first);
    a.step_rows = step_rows; a.h0 = h0; a.masks = masks; a.mask_in_table = mask_in_table; a.out = out; a.h_final = h_final;
    a.saved = static_cast<uint8_t*>(saved);
    a.L = L; a.NO = NO; a.R = R; a.n_tiles = (R + kRows - 1) / kRows;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (g_fwd_slots) {
        case 1: return launch_fwd<1>(a, st);
        case 2: return launch_fwd<2>(a, st);
        case 3: return launch_fwd<3>(a, st);
        default: return launch_fwd<4>(a, st);
    }
}

extern "C" int mm_ppo_trunk2_bwd(const float* params, const void* xhat, const int64_t* first, int64_t step_rows,
                                 const float* masks, int32_t mask_in_table, const void* saved, const float* dout,
                                 const float* scale, float* acc, int32_t n_acc_ctas, int32_t L, int64_t R, int32_t NO,
                                 void* stream) {
    if (!params || !xhat || !first || !saved || !dout || !scale || !acc || L <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (NO < 1 || NO > NOMAX) return MM_ERR_UNSUPPORTED;
    if (!mm::aligned(xhat, 16) || !mm::aligned(saved, 16) || !mm::aligned(acc, 16)) return MM_ERR_BAD_ARG;
    Args a{};
    a.params = params; a.xhat = static_cast<const __half*>(xhat); a.first = reinterpret_cast<const long long*>(first);
    a.step_rows = step_rows; a.masks = masks; a.mask_in_table = mask_in_table;
    a.saved = static_cast<uint8_t*>(const_cast<void*>(saved)); a.dout = dout;
    a.scale = scale; a.acc = acc;
    a.L = L; a.NO = NO; a.R = R; a.n_tiles = (R + kRows - 1) / kRows;
    const int grid = mm_ppo_trunk2_bwd_ctas(R);
    if (n_acc_ctas < grid) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return g_bwd_slots == 1 ? launch_bwd<1>(a, grid, st) : launch_bwd<2>(a, grid, st);
}
