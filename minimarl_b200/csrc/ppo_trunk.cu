// Fused recurrent trunk for the PPO update: forward AND backward of
//     LayerNorm(D) -> [Linear ReLU LayerNorm] x2 -> GRU(32) (mask-reset) -> LayerNorm -> Linear head
// over data chunks, every contraction of the data path on tcgen05 (kind::tf32, fp32 accumulators in TMEM).
//
// This is what R_Actor.evaluate_actions / R_Critic.forward run inside R_MAPPO.ppo_update
// (mappo/algorithms/r_actor_critic.py:95-146,189-208; mlp.py; rnn.py:24-80 -- the reference runs the GRU over the
// maximal mask-free runs of steps, which is the same function as multiplying the carried state by masks[t] at
// every step, because every step at which some row has mask 0 starts a new run).  One row = one data chunk
// (shared_buffer.py:318-427 recurrent_generator), one tile = 128 rows, 512 threads: thread (row, part) owns 8 of the
// 32 hidden columns of its row, the four parts of a row meet through shared memory for the LayerNorm statistics.
//
// Forward (training mode) also stores, per (step, row), the operands the weight gradients need: the normalised
// observation x0 [48], the NORMALISED (pre-affine) MLP activations xh1, xh2 [32] -- the host multiplies the matching
// weight-gradient blocks by the LayerNorm gains -- the masked input state hp [32], and in the spare columns of the row
// 1 / sigma of both LayerNorms and the two ReLU masks.
// Backward walks the steps in reverse; for each step it reloads xh1 / xh2 / hp, RE-COMPUTES only the GRU cell (one MMA
// phase: the gates from x2 = xh2 g2 + b2 and hp), then propagates
//     d head -> LayerNorm -> GRU cell -> (dx2 = dGi Wih, dhp = dGh Whh on tcgen05) -> LayerNorm/ReLU -> (dx1 = dA2 W2)
//     -> LayerNorm/ReLU
// and emits the pre-activation gradients dA1, dA2 [32], dGi [96], dGh_n [32] per (step, row).  The weight gradients
// are then four large-K GEMMs of those against x0 / x1 / x2 / hp (host side, TF32 cuBLAS); everything that is a
// column sum -- bias and LayerNorm gains/offsets, the head weight -- is reduced here (warp transpose-reduction,
// shared-memory accumulators, one partial vector per CTA).
//
// LayerNorm(D)'s affine is folded into the first Linear on the host (W1' = W1 diag(g0), b1' = b1 + W1 b0); the host
// unfolds the gradients (dW1 = dW1' diag(g0), dg0 = sum_n dW1'[n,:] * W1[n,:], db0 = W1^T db1').
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;
constexpr int kThreads = 512;
constexpr int H = 32;
constexpr int K1 = 48;
constexpr int CPT = 8;
constexpr int NOMAX = 8;
constexpr float kEps = 1e-5f;

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// MUFU.TANH (one special-function op; abs error ~5e-4, below the tf32 operand rounding); sigmoid through the same unit
__device__ __forceinline__ float tanh_(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_(float x) { return fmaf(tanh_(0.5f * x), 0.5f, 0.5f); }

// Parameter buffer (floats).  Matrices: w1 [H][K1] (LayerNorm(D) folded, zero padded), w2 [H][H], wih [3H][H],
// whh [3H][H].  Small: b1 (folded), ln1_w, ln1_b, b2, ln2_w, ln2_b [H], bih, bhh [3H], ln3_w, ln3_b [H], wo [8][H], bo [8].
struct Layout {
    int w1, w2, wih, whh, small, b1, ln1_w, ln1_b, b2, ln2_w, ln2_b, bih, bhh, ln3_w, ln3_b, wo, bo, total;
};
__host__ __device__ inline Layout make_layout() {
    Layout L;
    int o = 0;
    auto take = [&](int n) { int at = o; o += n; return at; };
    L.w1 = take(H * K1); L.w2 = take(H * H); L.wih = take(3 * H * H); L.whh = take(3 * H * H);
    L.small = o;
    L.b1 = take(H); L.ln1_w = take(H); L.ln1_b = take(H);
    L.b2 = take(H); L.ln2_w = take(H); L.ln2_b = take(H);
    L.bih = take(3 * H); L.bhh = take(3 * H);
    L.ln3_w = take(H); L.ln3_b = take(H);
    L.wo = take(NOMAX * H); L.bo = take(NOMAX);
    L.total = o;
    return L;
}

// row stride of the saved-operand matrix X: x0 [0,48) xh1 [48,80) xh2 [80,112) hp [112,144) | 1 [144] 1/sigma1 [145] 1/sigma2 [146]
// 0 [147] | per column part p: relu1 bits + 256 * relu2 bits as a float [148 + p].  Columns 145..151 are read by the backward
// kernel only; the weight-gradient GEMM produces (ignored) output columns for them.
constexpr int XS = 152;
constexpr int GS = 192;   // row stride of the gradient matrix G: dA1 [0,32) dA2 [32,64) dGi [64,160) dGh_n [160,192)
// per-CTA column-sum accumulators: S[o][j] = sum_r dout[r][o] * xhat3[r][j] (NOMAX x H), dbo (NOMAX), dgamma2 (H), dgamma1 (H)
constexpr int kTmemS = 5;   // head outputs whose S rows are accumulated in TMEM (the rest, if any, by warp shuffles per step)
constexpr int ACC_S = 0, ACC_BO = NOMAX * H, ACC_G2 = ACC_BO + NOMAX, ACC_G1 = ACC_G2 + H, ACC_N = ACC_G1 + H;

struct Args {
    const float* params;
    const float* obs;      // [L][R][D], or (obs_row != null) a row table [*][D] addressed through obs_row
    const long long* obs_row;   // [R] first-step row of every chunk in the tables, or null
    long long obs_step;    // rows between consecutive steps of a chunk in the tables
    const float* h0;       // [R][H]; with obs_row: table [*][H], row obs_row[r]             (forward)
    const float* masks;    // [L][R] or null; with obs_row: table [*], entry obs_row[r] + t * obs_step
    float* out;            // [L][R][NO]                  (forward)
    float* h_final;        // [R][H] or null              (forward)
    float* X;              // [L*R][XS]  forward writes columns [0,144) (may be null = inference); backward reads hp
    const float* dout;     // [L][R][NO]                  (backward)
    float* G;              // [L*R][GS]                   (backward)
    float* gacc;           // [grid][ACC_N]               (backward)
    int L, D, NO;
    long long R, n_tiles;
    long long* prof;       // optional: CTA 0 accumulates clock64() deltas per phase (debug / profiles)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// column sums of v[0..7] over the 32 lanes of the warp (= 32 rows), added to acc[0..7] -- a shared-memory slot that only
// this warp updates, so the read-modify-write needs no atomic
__device__ __forceinline__ void colsum8(const float (&v)[CPT], float* acc, int lane) {
    float a[4], b[2], c;
    const bool hi4 = lane & 16, hi3 = lane & 8, hi2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = hi4 ? v[4 + i] : v[i], send = hi4 ? v[i] : v[4 + i];
        a[i] = keep + __shfl_xor_sync(MM_FULL_MASK, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = hi3 ? a[2 + i] : a[i], send = hi3 ? a[i] : a[2 + i];
        b[i] = keep + __shfl_xor_sync(MM_FULL_MASK, send, 8);
    }
    {
        const float keep = hi2 ? b[1] : b[0], send = hi2 ? b[0] : b[1];
        c = keep + __shfl_xor_sync(MM_FULL_MASK, send, 4);
    }
    c += __shfl_xor_sync(MM_FULL_MASK, c, 2);
    c += __shfl_xor_sync(MM_FULL_MASK, c, 1);
    if ((lane & 3) == 0) acc[((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] += c;
}

__device__ __forceinline__ void store8_kmajor(uint8_t* tile, int r, int k0, int K, const float (&v)[CPT]) {
    *reinterpret_cast<float4*>(tile + tc::kmajor_offset(r, k0, K)) =
        make_float4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
    *reinterpret_cast<float4*>(tile + tc::kmajor_offset(r, k0 + 4, K)) =
        make_float4(to_tf32(v[4]), to_tf32(v[5]), to_tf32(v[6]), to_tf32(v[7]));
}
__device__ __forceinline__ void store8_kmajor_raw(uint8_t* tile, int r, int k0, int K, const float (&v)[CPT]) {
    *reinterpret_cast<float4*>(tile + tc::kmajor_offset(r, k0, K)) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(tile + tc::kmajor_offset(r, k0 + 4, K)) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void load8_kmajor(const uint8_t* tile, int r, int k0, int K, float (&v)[CPT]) {
    const float4 q0 = *reinterpret_cast<const float4*>(tile + tc::kmajor_offset(r, k0, K));
    const float4 q1 = *reinterpret_cast<const float4*>(tile + tc::kmajor_offset(r, k0 + 4, K));
    v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
}
__device__ __forceinline__ void store8_global(float* dst, const float (&v)[CPT]) {
    mm::st_stream(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
    mm::st_stream(reinterpret_cast<float4*>(dst) + 1, make_float4(v[4], v[5], v[6], v[7]));
}
__device__ __forceinline__ void load8_global(const float* src, float (&v)[CPT]) {
    const float4 q0 = *reinterpret_cast<const float4*>(src), q1 = *reinterpret_cast<const float4*>(src + 4);
    v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
}

// Copies between a [128 x F] K-major tile and F consecutive columns of global rows (row stride gs floats).  A warp moves
// 8 rows x 4 sixteen-byte chunks per instruction: conflict-free on the tile side, 64 contiguous bytes per row on the
// global side.  Thread (warp w, lane) always handles row 8 (w & 15) + (lane & 7) and chunks (lane >> 3) + 4 i, so every
// address is a per-thread base plus a compile-time offset.
template <int F>
__device__ __forceinline__ void tile_to_global(const uint8_t* tile, float* g, long long gs, int n_valid, int tid) {
    static_assert(kThreads == 512 && F % 16 == 0, "one pass of 16 warps covers all 128 rows");
    const int r = ((tid >> 5) & 15) * 8 + (tid & 7), c = (tid & 31) >> 3;
    if (r >= n_valid) return;
    const uint8_t* src = tile + tc::kmajor_offset(r, 4 * c, F);
    float4* dst = reinterpret_cast<float4*>(g + r * gs + 4 * c);
#pragma unroll
    for (int i = 0; i < F / 16; ++i) mm::st_stream(dst + 4 * i, *reinterpret_cast<const float4*>(src + 512 * i));
}
template <int F>
__device__ __forceinline__ void global_to_tile_async(uint8_t* tile, const float* g, long long gs, int n_valid, int tid) {
    static_assert(kThreads == 512 && F % 16 == 0, "one pass of 16 warps covers all 128 rows");
    const int r = ((tid >> 5) & 15) * 8 + (tid & 7), c = (tid & 31) >> 3;
    uint8_t* dst = tile + tc::kmajor_offset(r, 4 * c, F);
    const float* src = g + min(r, n_valid - 1) * gs + 4 * c;
#pragma unroll
    for (int i = 0; i < F / 16; ++i) cp_async16(dst + 512 * i, src + 16 * i);
}

#define PROF(i)                                                      \
    do {                                                             \
        if (a.prof && blockIdx.x == 0 && tid == 0) {                 \
            const long long now_ = clock64();                        \
            a.prof[i] += now_ - prof_t;                              \
            prof_t = now_;                                           \
        }                                                            \
    } while (0)

template <bool BWD>
__global__ void __launch_bounds__(kThreads, 1) ppo_trunk_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const Layout L = make_layout();
    // ---- shared memory ---------------------------------------------------------------------------------------------
    uint8_t* sW1 = smem;                                  // [H x K1]
    uint8_t* sW2 = sW1 + H * K1 * 4;                      // [H x H]
    uint8_t* sWih = sW2 + H * H * 4;                      // [3H x H]
    uint8_t* sWhh = sWih + 3 * H * H * 4;                 // [3H x H]
    uint8_t* sX0 = sWhh + 3 * H * H * 4;                  // [128 x K1]
    uint8_t* sX1 = sX0 + kRows * K1 * 4;                  // [128 x H]
    uint8_t* sX2 = sX1 + kRows * H * 4;
    uint8_t* sHp = sX2 + kRows * H * 4;
    float2* red = reinterpret_cast<float2*>(sHp + kRows * H * 4);          // [2][128][4], alternating
    float* sP = reinterpret_cast<float*>(red + 2 * kRows * 4) - L.small;   // small params, biased by L.small
    uint8_t* tail = reinterpret_cast<uint8_t*>(sP + L.total);
    tail = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tail) + 127) & ~uintptr_t(127));
    long long* sRowBase = reinterpret_cast<long long*>(tail);              // [128] first-step table row of the tile's chunks
    tail += kRows * 8;
    // backward only
    uint8_t* sW2T = tail;                                 // [H(in) x H(out)]
    uint8_t* sWihT = sW2T + H * H * 4;                    // [H(in) x 3H(out)]
    uint8_t* sWhhT = sWihT + 3 * H * H * 4;               // [H(in) x 3H(out)]
    uint8_t* sdG = sWhhT + 3 * H * H * 4;                 // [128 x 4H]: dr | dz | dn | dn*r ; dA2 re-uses its first 16 KB
    float* sAcc = reinterpret_cast<float*>(sdG + kRows * 4 * H * 4);       // [4 row quarters][ACC_N]
    float* sDout = sAcc + 4 * ACC_N;                                       // [128][NO] head gradients of the current step
    float* sTail = sDout + kRows * NOMAX;                                  // [128][8] columns 144..151 of the step's X rows
    // regions that are idle while a step starts double as the raw observation tile (filled by cp.async one step ahead)
    // and, in the forward, as the head's partial sums
    float* sObs = reinterpret_cast<float*>(sX1);                           // [128][D]  (<= 24 KB, forward only)
    uint8_t* sStage = tail;                               // [128 x H] un-rounded xh1 / xh2 on their way to X (forward only; aliases sW2T..)
    float* sTailF = reinterpret_cast<float*>(sStage + kRows * H * 4);      // [128][8] columns 144..151 of the step's X rows (forward)
    float* sHead = reinterpret_cast<float*>(sX0);                          // [NOMAX][4][128]  (forward only)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane, part = warp >> 2, j0 = part * CPT;

    constexpr uint32_t kTmemCols = BWD ? 512 : 256;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, kTmemCols);
    if (tid == 0) tc::mbar_init(&bar, 1);
    auto stage = [&](uint8_t* dst, int off, int rows, int K) {
        for (int i = tid; i < rows * K; i += kThreads) {
            const int r = i / K, k = i - r * K;
            *reinterpret_cast<float*>(dst + tc::kmajor_offset(r, k, K)) = to_tf32(a.params[off + i]);
        }
    };
    auto stage_t = [&](uint8_t* dst, int off, int rows, int K) {       // dst[k][r] = src[r][k]: tile [K x rows]
        for (int i = tid; i < rows * K; i += kThreads) {
            const int r = i / K, k = i - r * K;
            *reinterpret_cast<float*>(dst + tc::kmajor_offset(k, r, rows)) = to_tf32(a.params[off + i]);
        }
    };
    stage(sW1, L.w1, H, K1);
    stage(sW2, L.w2, H, H);
    stage(sWih, L.wih, 3 * H, H);
    stage(sWhh, L.whh, 3 * H, H);
    for (int i = L.small + tid; i < L.total; i += kThreads) sP[i] = a.params[i];
    if (BWD) {
        stage_t(sW2T, L.w2, H, H);
        stage_t(sWihT, L.wih, 3 * H, H);
        stage_t(sWhhT, L.whh, 3 * H, H);
        for (int i = tid; i < 4 * ACC_N; i += kThreads) sAcc[i] = 0.f;
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t phase = 0, red_sel = 0;
    // backward: running per-thread column partials live in the upper half of TMEM (64 columns per part: S[0..4] | dgamma2 |
    // dgamma1 | dbo) and are reduced across rows once, at the end of the kernel
    const uint32_t acc_addr = lane_addr + 256 + part * 64;
    auto tmem_accumulate = [&](uint32_t col, const float (&v)[CPT]) {
        float t8[CPT];
        tc::tmem_ld8(acc_addr + col, t8);
#pragma unroll
        for (int j = 0; j < CPT; ++j) t8[j] += v[j];
        tc::tmem_st8(acc_addr + col, t8);
    };
    if (BWD) {
        const float zero8[CPT] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 64; c += 8) tc::tmem_st8(acc_addr + c, zero8);
        tc::tmem_wait_st();
    }
    long long prof_t = clock64();
    float* acc = sAcc + (warp & 3) * ACC_N;               // this warp's row quarter

    auto desc = [&](const uint8_t* base, int row_off, int K) {
        return tc::make_desc(tc::smem_u32(base) + (uint32_t)(row_off >> 3) * (uint32_t)(K >> 2) * 128u, 128u,
                             (uint32_t)(K >> 2) * 128u);
    };
    const uint64_t dX0 = desc(sX0, 0, K1), dW1 = desc(sW1, 0, K1), dX1 = desc(sX1, 0, H), dW2 = desc(sW2, 0, H);
    const uint64_t dX2 = desc(sX2, 0, H), dHp = desc(sHp, 0, H);
    const uint64_t dWih_rz = desc(sWih, 0, H), dWih_n = desc(sWih, 2 * H, H);
    const uint64_t dWhh_rz = desc(sWhh, 0, H), dWhh_n = desc(sWhh, 2 * H, H);
    const uint64_t dG = desc(sdG, 0, 4 * H), dA2 = desc(sdG, 0, H);
    const uint64_t dW2T = desc(sW2T, 0, H), dWihT = desc(sWihT, 0, 3 * H), dWhhT = desc(sWhhT, 0, 3 * H);
    const uint32_t id64 = tc::make_idesc_tf32(128, 64), id32 = tc::make_idesc_tf32(128, 32);
    // TMEM columns: fc1 [0,32) fc2 [32,64) rz [64,128) i_n [128,160) h_n [160,192); backward: dx2 [192,224) dhp [224,256),
    // dx1 re-uses [0,32)
    auto publish = [&]() {
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    };
    auto wait_mma = [&]() {
        if (warp == 0) tc::mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        tc::fence_after_sync();
    };
    // sums of (s, q) over the four parts of a row; the exchange buffer alternates so that one barrier is enough
    auto row_sums = [&](float s, float q, float& S, float& Q) {
        float2* r2 = red + red_sel * (kRows * 4);
        red_sel ^= 1;
        r2[part * kRows + row] = make_float2(s, q);                              // [part][row]: conflict-free both ways
        asm volatile("bar.sync %0, 128;" ::"r"(1 + (warp & 3)) : "memory");      // the four warps that share these 32 rows
        const float2 p0 = r2[row], p1 = r2[kRows + row], p2 = r2[2 * kRows + row], p3 = r2[3 * kRows + row];
        S = (p0.x + p1.x) + (p2.x + p3.x);
        Q = (p0.y + p1.y) + (p2.y + p3.y);
    };
    auto row_stats = [&](float s, float q, float& mean, float& rstd, int n) {
        float S, Q;
        row_sums(s, q, S, Q);
        mean = S / n;
        rstd = rsqrtf(fmaxf(Q / n - mean * mean, 0.f) + kEps);
    };
    // raw observation tile of (tile row0, step t) -> sObs, asynchronously; rows past the end of the batch are zero
    auto fetch_obs = [&](long long row0, int n_valid, int t) {
        if (a.obs_row) {
            // chunk rows gathered straight from the rollout buffer: row r of the tile is table row sRowBase[r] + t * obs_step
            const long long toff = (long long)t * a.obs_step;
#pragma unroll
            for (int j = 0; j < kRows / (kThreads / 32); ++j) {
                const int r = warp + (kThreads / 32) * j;
                if (r < n_valid) {
                    const float* src = a.obs + (sRowBase[r] + toff) * a.D;
                    float* dst = sObs + r * a.D;
                    if (lane < a.D) cp_async4(dst + lane, src + lane);
                    if (lane + 32 < a.D) cp_async4(dst + lane + 32, src + lane + 32);
                }
            }
            for (int i = n_valid * a.D + tid; i < kRows * a.D; i += kThreads) sObs[i] = 0.f;
            cp_async_commit();
            return;
        }
        const float* src = a.obs + ((long long)t * a.R + row0) * a.D;
        const int n = n_valid * a.D;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            for (int i = tid; i < (n >> 2); i += kThreads) cp_async16(sObs + 4 * i, src + 4 * i);
            for (int i = (n & ~3) + tid; i < n; i += kThreads) cp_async4(sObs + i, src + i);
        } else {
            for (int i = tid; i < n; i += kThreads) cp_async4(sObs + i, src + i);
        }
        for (int i = n + tid; i < kRows * a.D; i += kThreads) sObs[i] = 0.f;
        cp_async_commit();
    };
    // head gradients of (tile row0, step t) -> sDout, asynchronously (contiguous n_valid * NO floats)
    auto fetch_dout = [&](long long row0, int n_valid, int t) {
        const float* src = a.dout + ((long long)t * a.R + row0) * a.NO;
        const int n = n_valid * a.NO;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            for (int i = tid; i < (n >> 2); i += kThreads) cp_async16(sDout + 4 * i, src + 4 * i);
            for (int i = (n & ~3) + tid; i < n; i += kThreads) cp_async4(sDout + i, src + i);
        } else {
            for (int i = tid; i < n; i += kThreads) cp_async4(sDout + i, src + i);
        }
        for (int i = n + tid; i < kRows * a.NO; i += kThreads) sDout[i] = 0.f;
    };

    // columns 144..151 of the X rows of (tile, step t): LayerNorm scales and ReLU masks the forward left there
    auto fetch_tail = [&](long long row0, int n_valid, int t) {
        if (tid < 2 * kRows) {
            const int r = tid >> 1, half = tid & 1;
            const float* src = a.X + ((long long)t * a.R + row0 + min(r, n_valid - 1)) * XS + 144 + 4 * half;
            cp_async16(sTail + r * 8 + 4 * half, src);
        }
    };

    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const long long row0 = tile * kRows;
        const int n_valid = (int)min((long long)kRows, a.R - row0);
        const bool valid = row < n_valid;
        const long long grow = valid ? row0 + row : a.R - 1;        // out-of-range rows shadow the last row, stores guarded

        float h[CPT];          // forward: carried hidden state; backward: gradient carried to the previous step
#pragma unroll
        for (int j = 0; j < CPT; ++j) h[j] = 0.f;
        if (a.obs_row) {
            if (tid < kRows) sRowBase[tid] = a.obs_row[row0 + min(tid, n_valid - 1)];
            __syncthreads();
        }
        if (!BWD) {
            load8_global(a.obs_row ? a.h0 + sRowBase[row] * H + j0 : a.h0 + grow * H + j0, h);
            fetch_obs(row0, n_valid, 0);
        } else {
            // backward: the normalised activations and the input state of a step come straight from X into the tiles
            const float* xrow = a.X + ((long long)(a.L - 1) * a.R + row0) * XS;
            global_to_tile_async<H>(sX0, xrow + 48, XS, n_valid, tid);
            global_to_tile_async<H>(sX2, xrow + 80, XS, n_valid, tid);
            global_to_tile_async<H>(sHp, xrow + 112, XS, n_valid, tid);
            fetch_tail(row0, n_valid, a.L - 1);
            fetch_dout(row0, n_valid, a.L - 1);
            cp_async_commit();
        }

        for (int step = 0; step < a.L; ++step) {
            const int t = BWD ? a.L - 1 - step : step;
            const long long tr = (long long)t * a.R + grow;
            const float mask = !a.masks ? 1.0f : a.obs_row ? a.masks[sRowBase[row] + (long long)t * a.obs_step] : a.masks[tr];
            float hp[CPT];
            // ---- F1. plain LayerNorm of the observation row -> X0 tile; masked input state -> Hp tile ----------------
            cp_async_wait_all();
            __syncthreads();
            float xh1[CPT], rstd1, xh2[CPT], rstd2;
            uint32_t relu1 = 0, relu2 = 0;
            float* const xtile = a.X ? a.X + ((long long)t * a.R + row0) * XS : nullptr;      // this tile's rows of X at step t
            if (BWD) {
                // backward: xh1 / xh2 / hp of this step are in the tiles, the LayerNorm scales and ReLU masks in the row's tail
                load8_kmajor(sHp, row, j0, H, hp);
                load8_kmajor(sX0, row, j0, H, xh1);
                load8_kmajor(sX2, row, j0, H, xh2);
                const float* tailp = sTail + row * 8;
                rstd1 = tailp[1];
                rstd2 = tailp[2];
                const uint32_t bits = (uint32_t)tailp[4 + part];
                relu1 = bits & 0xffu;
                relu2 = bits >> 8;
                float y[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) y[j] = fmaf(xh2[j], sP[L.ln2_w + j0 + j], sP[L.ln2_b + j0 + j]);
                store8_kmajor(sX2, row, j0, H, y);              // in place: the operand of the gate MMAs
            } else {
                float x0[12];
                const float* o = sObs + row * a.D;
                float s = 0.f, q = 0.f;
#pragma unroll
                for (int gi = 0; gi < 3; ++gi)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = 4 * (part + 4 * gi) + u;
                        const float x = k < a.D ? o[k] : 0.f;
                        x0[4 * gi + u] = x;
                        s += x; q = fmaf(x, x, q);
                    }
                float mean, rstd;
                row_stats(s, q, mean, rstd, a.D);
#pragma unroll
                for (int gi = 0; gi < 3; ++gi) {
                    const int g = part + 4 * gi;
                    float4 v;
                    v.x = 4 * g + 0 < a.D ? to_tf32((x0[4 * gi + 0] - mean) * rstd) : 0.f;
                    v.y = 4 * g + 1 < a.D ? to_tf32((x0[4 * gi + 1] - mean) * rstd) : 0.f;
                    v.z = 4 * g + 2 < a.D ? to_tf32((x0[4 * gi + 2] - mean) * rstd) : 0.f;
                    v.w = 4 * g + 3 < a.D ? to_tf32((x0[4 * gi + 3] - mean) * rstd) : 0.f;
                    *reinterpret_cast<float4*>(sX0 + tc::kmajor_offset(row, 4 * g, K1)) = v;
                }
                // the state enters the cell rounded to tf32 -- the same value the tensor core multiplies and the backward re-reads
#pragma unroll
                for (int j = 0; j < CPT; ++j) hp[j] = to_tf32(h[j] * mask);
                store8_kmajor(sHp, row, j0, H, hp);
                PROF(0);
                publish();
                PROF(1);
                if (tid == 0) {
#pragma unroll
                    for (int kk = 0; kk < K1 / 8; ++kk) tc::mma_tf32(tmem, dX0 + 16 * kk, dW1 + 16 * kk, id32, kk > 0);
                    tc::mma_commit(&bar);
                }
                if (xtile) {
                    tile_to_global<K1>(sX0, xtile, XS, n_valid, tid);
                    tile_to_global<H>(sHp, xtile + 112, XS, n_valid, tid);
                }
                wait_mma();
                PROF(2);
                // ---- F2. fc1 epilogue --------------------------------------------------------------------------------
                {
                    float v[CPT];
                    tc::tmem_ld8(lane_addr + j0, v);
                    float s1 = 0.f, q1 = 0.f;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        v[j] += sP[L.b1 + j0 + j];
                        if (v[j] > 0.f) relu1 |= 1u << j; else v[j] = 0.f;
                        s1 += v[j]; q1 = fmaf(v[j], v[j], q1);
                    }
                    float mean1;
                    row_stats(s1, q1, mean1, rstd1, H);
                    float y[CPT];
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        xh1[j] = (v[j] - mean1) * rstd1;
                        y[j] = fmaf(xh1[j], sP[L.ln1_w + j0 + j], sP[L.ln1_b + j0 + j]);
                    }
                    store8_kmajor(sX1, row, j0, H, y);
                    if (xtile) store8_kmajor_raw(sStage, row, j0, H, xh1);
                }
                publish();
                if (tid == 0) {
#pragma unroll
                    for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 32, dX1 + 16 * kk, dW2 + 16 * kk, id32, kk > 0);
                    tc::mma_commit(&bar);
                }
                if (xtile) tile_to_global<H>(sStage, xtile + 48, XS, n_valid, tid);
                PROF(3);
                wait_mma();
                PROF(4);
                // ---- F3. fc2 epilogue --------------------------------------------------------------------------------
                {
                    float v[CPT];
                    tc::tmem_ld8(lane_addr + 32 + j0, v);
                    float s2 = 0.f, q2 = 0.f;
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        v[j] += sP[L.b2 + j0 + j];
                        if (v[j] > 0.f) relu2 |= 1u << j; else v[j] = 0.f;
                        s2 += v[j]; q2 = fmaf(v[j], v[j], q2);
                    }
                    float mean2;
                    row_stats(s2, q2, mean2, rstd2, H);
                    float y[CPT];
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        xh2[j] = (v[j] - mean2) * rstd2;
                        y[j] = fmaf(xh2[j], sP[L.ln2_w + j0 + j], sP[L.ln2_b + j0 + j]);
                    }
                    store8_kmajor(sX2, row, j0, H, y);
                    if (xtile) {
                        store8_kmajor_raw(sStage, row, j0, H, xh2);     // the xh1 copy out of sStage finished before the last barrier
                        sTailF[row * 8 + part] = part == 0 ? 1.0f : part == 1 ? rstd1 : part == 2 ? rstd2 : 0.f;
                        sTailF[row * 8 + 4 + part] = (float)(relu1 | (relu2 << 8));
                    }
                }
            }
            publish();
            if (tid == 0) {
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 64, dX2 + 16 * kk, dWih_rz + 16 * kk, id64, kk > 0);
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 64, dHp + 16 * kk, dWhh_rz + 16 * kk, id64, true);
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 128, dX2 + 16 * kk, dWih_n + 16 * kk, id32, kk > 0);
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 160, dHp + 16 * kk, dWhh_n + 16 * kk, id32, kk > 0);
                tc::mma_commit(&bar);
            }
            if (!BWD && xtile) {
                tile_to_global<H>(sStage, xtile + 80, XS, n_valid, tid);
                if (tid < 2 * kRows && (tid >> 1) < n_valid)
                    mm::st_stream(reinterpret_cast<float4*>(xtile + (tid >> 1) * XS + 144 + 4 * (tid & 1)),
                                  *reinterpret_cast<const float4*>(sTailF + (tid >> 1) * 8 + 4 * (tid & 1)));
            }
            PROF(5);
            wait_mma();
            PROF(6);
            if (step + 1 < a.L) {
                if (!BWD) {
                    fetch_obs(row0, n_valid, t + 1);     // the X1 / X2 tiles are consumed, the next observations land there
                } else {
                    const float* xrow = a.X + ((long long)(t - 1) * a.R + row0) * XS;   // every thread holds its xh1 / xh2 in registers
                    global_to_tile_async<H>(sX0, xrow + 48, XS, n_valid, tid);
                    global_to_tile_async<H>(sX2, xrow + 80, XS, n_valid, tid);
                    global_to_tile_async<H>(sHp, xrow + 112, XS, n_valid, tid);
                    fetch_tail(row0, n_valid, t - 1);
                    cp_async_commit();
                }
            }
            // ---- F4. GRU cell, LayerNorm, head -----------------------------------------------------------------------
            float r[CPT], z[CPT], n[CPT], hnb[CPT], xh3[CPT], rstd3;
            {
                float gr[CPT], gz[CPT], gin[CPT], ghn[CPT];
                tc::tmem_ld8_issue(lane_addr + 64 + j0, gr);       // four loads in flight, one wait
                tc::tmem_ld8_issue(lane_addr + 96 + j0, gz);
                tc::tmem_ld8_issue(lane_addr + 128 + j0, gin);
                tc::tmem_ld8_issue(lane_addr + 160 + j0, ghn);
                tc::tmem_ld_fence<CPT>(gr, gz, gin, ghn);
                const float* bi = sP + L.bih + j0;
                const float* bh = sP + L.bhh + j0;
                float hn[CPT], s = 0.f, q = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    r[j] = sigmoid_(gr[j] + bi[j] + bh[j]);
                    z[j] = sigmoid_(gz[j] + bi[H + j] + bh[H + j]);
                    hnb[j] = ghn[j] + bh[2 * H + j];
                    n[j] = tanh_(gin[j] + bi[2 * H + j] + r[j] * hnb[j]);
                    hn[j] = fmaf(z[j], hp[j] - n[j], n[j]);
                    s += hn[j]; q = fmaf(hn[j], hn[j], q);
                }
                float mean;
                row_stats(s, q, mean, rstd3, H);
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    xh3[j] = (hn[j] - mean) * rstd3;
                    if (!BWD) h[j] = hn[j];
                }
            }
            if (!BWD) {
                for (int o = 0; o < a.NO; ++o) {
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < CPT; ++j)
                        sum = fmaf(fmaf(xh3[j], sP[L.ln3_w + j0 + j], sP[L.ln3_b + j0 + j]), sP[L.wo + o * H + j0 + j], sum);
                    sHead[(o * 4 + part) * kRows + row] = sum;
                }
                __syncthreads();
                if (part == 0 && valid)
                    for (int o = 0; o < a.NO; ++o)
                        a.out[tr * a.NO + o] = sP[L.bo + o] + ((sHead[(o * 4) * kRows + row] + sHead[(o * 4 + 1) * kRows + row]) +
                                                               (sHead[(o * 4 + 2) * kRows + row] + sHead[(o * 4 + 3) * kRows + row]));
                PROF(7);
                continue;           // the next step's first barrier orders these sHead reads before the X0 tile is rewritten
            }
            PROF(7);
            // ================================================ backward ================================================
            float dy[CPT];
            {
                float dO[NOMAX];
#pragma unroll
                for (int o = 0; o < NOMAX; ++o) dO[o] = (o < a.NO) ? sDout[row * a.NO + o] : 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) dy[j] = 0.f;
#pragma unroll
                for (int o = 0; o < NOMAX; ++o)
                    if (o < a.NO) {
#pragma unroll
                        for (int j = 0; j < CPT; ++j) dy[j] = fmaf(dO[o], sP[L.wo + o * H + j0 + j], dy[j]);
                    }
                // S[o][:] += dout[o] * xhat3 for the first kTmemS outputs and dbo += dout: running sums in TMEM, the loads of a
                // group in flight together (one wait per group instead of one per output)
                {
                    float t0[CPT], t1[CPT], t2[CPT];
                    tc::tmem_ld8_issue(acc_addr + 0, t0);
                    tc::tmem_ld8_issue(acc_addr + 8, t1);
                    tc::tmem_ld8_issue(acc_addr + 16, t2);
                    tc::tmem_ld_fence<CPT>(t0, t1, t2);
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        t0[j] = fmaf(dO[0], xh3[j], t0[j]); t1[j] = fmaf(dO[1], xh3[j], t1[j]); t2[j] = fmaf(dO[2], xh3[j], t2[j]);
                    }
                    tc::tmem_st8(acc_addr + 0, t0);
                    tc::tmem_st8(acc_addr + 8, t1);
                    tc::tmem_st8(acc_addr + 16, t2);
                }
                {
                    float t3[CPT], t4[CPT], t5[CPT];
                    tc::tmem_ld8_issue(acc_addr + 24, t3);
                    tc::tmem_ld8_issue(acc_addr + 32, t4);
                    tc::tmem_ld8_issue(acc_addr + 56, t5);
                    tc::tmem_ld_fence<CPT>(t3, t4, t5);
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        t3[j] = fmaf(dO[3], xh3[j], t3[j]); t4[j] = fmaf(dO[4], xh3[j], t4[j]);
                        if (part == 0) t5[j] += dO[j];
                    }
                    tc::tmem_st8(acc_addr + 24, t3);
                    tc::tmem_st8(acc_addr + 32, t4);
                    tc::tmem_st8(acc_addr + 56, t5);
                }
                static_assert(kTmemS == 5, "the grouped TMEM accumulation above covers outputs 0..4");
#pragma unroll
                for (int o = kTmemS; o < NOMAX; ++o)
                    if (o < a.NO) {
                        float p[CPT];
#pragma unroll
                        for (int j = 0; j < CPT; ++j) p[j] = dO[o] * xh3[j];
                        colsum8(p, acc + ACC_S + o * H + j0, lane);
                    }
            }
            PROF(13);
            // LayerNorm(h) backward
            float dh[CPT];
            {
                float g[CPT], s = 0.f, q = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    g[j] = dy[j] * sP[L.ln3_w + j0 + j];
                    s += g[j]; q = fmaf(g[j], xh3[j], q);
                }
                float S, Q;
                row_sums(s, q, S, Q);
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                    dh[j] = rstd3 * (g[j] - S * (1.0f / H) - xh3[j] * Q * (1.0f / H)) + h[j];
            }
            PROF(14);
            // GRU cell backward: h = n + z (hp - n)
            {
                float dr[CPT], dz[CPT], dn[CPT], dnr[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const float dn_ = dh[j] * (1.f - z[j]);
                    dn[j] = dn_ * (1.f - n[j] * n[j]);
                    dz[j] = dh[j] * (hp[j] - n[j]) * z[j] * (1.f - z[j]);
                    dr[j] = dn[j] * hnb[j] * r[j] * (1.f - r[j]);
                    dnr[j] = dn[j] * r[j];
                    h[j] = dh[j] * z[j];                              // direct path to hp; the matrix path is added below
                }
                store8_kmajor(sdG, row, j0, 4 * H, dr);
                store8_kmajor(sdG, row, H + j0, 4 * H, dz);
                store8_kmajor(sdG, row, 2 * H + j0, 4 * H, dn);
                store8_kmajor(sdG, row, 3 * H + j0, 4 * H, dnr);
            }
            PROF(15);
            publish();
            if (tid == 0) {
                // dx2 = [dr dz dn] Wih  (K = 96);  dhp = [dr dz dn*r] Whh  (K = 96, the last 32 from tile columns 96..127)
#pragma unroll
                for (int kk = 0; kk < 12; ++kk) tc::mma_tf32(tmem + 192, dG + 16 * kk, dWihT + 16 * kk, id32, kk > 0);
#pragma unroll
                for (int kk = 0; kk < 12; ++kk)
                    tc::mma_tf32(tmem + 224, dG + 16 * (kk < 8 ? kk : kk + 4), dWhhT + 16 * kk, id32, kk > 0);
                tc::mma_commit(&bar);
            }
            float* const gtile = a.G + ((long long)t * a.R + row0) * GS;
            tile_to_global<4 * H>(sdG, gtile + 64, GS, n_valid, tid);
            PROF(8);
            wait_mma();
            PROF(9);
            if (step + 1 < a.L) {                        // every thread is past its sDout reads: bring in the previous step's
                fetch_dout(row0, n_valid, t - 1);
                cp_async_commit();
            }
            // ---- dx2 -> LayerNorm / ReLU backward -> dA2 ---------------------------------------------------------------
            {
                float dhp[CPT], dx[CPT];
                tc::tmem_ld8_issue(lane_addr + 224 + j0, dhp);
                tc::tmem_ld8_issue(lane_addr + 192 + j0, dx);
                tc::tmem_ld_fence<CPT>(dhp, dx);
#pragma unroll
                for (int j = 0; j < CPT; ++j) h[j] = (h[j] + dhp[j]) * mask;
                float g[CPT], p[CPT], s = 0.f, q = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    g[j] = dx[j] * sP[L.ln2_w + j0 + j];
                    p[j] = dx[j] * xh2[j];
                    s += g[j]; q = fmaf(g[j], xh2[j], q);
                }
                tmem_accumulate(40, p);
                float S, Q;
                row_sums(s, q, S, Q);
                float da[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const float v = rstd2 * (g[j] - S * (1.0f / H) - xh2[j] * Q * (1.0f / H));
                    da[j] = (relu2 >> j & 1u) ? v : 0.f;
                }
                store8_kmajor(sdG, row, j0, H, da);                  // dA2 tile over the (consumed) head of the dG tile
            }
            publish();
            if (tid == 0) {
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem, dA2 + 16 * kk, dW2T + 16 * kk, id32, kk > 0);
                tc::mma_commit(&bar);
            }
            tile_to_global<H>(sdG, gtile + 32, GS, n_valid, tid);
            PROF(10);
            wait_mma();
            PROF(11);
            // ---- dx1 -> LayerNorm / ReLU backward -> dA1 ---------------------------------------------------------------
            {
                float dx[CPT];
                tc::tmem_ld8(lane_addr + j0, dx);
                float g[CPT], p[CPT], s = 0.f, q = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    g[j] = dx[j] * sP[L.ln1_w + j0 + j];
                    p[j] = dx[j] * xh1[j];
                    s += g[j]; q = fmaf(g[j], xh1[j], q);
                }
                tmem_accumulate(48, p);
                float S, Q;
                row_sums(s, q, S, Q);
                float da[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const float v = rstd1 * (g[j] - S * (1.0f / H) - xh1[j] * Q * (1.0f / H));
                    da[j] = (relu1 >> j & 1u) ? v : 0.f;
                }
                store8_kmajor(sX1, row, j0, H, da);                  // staged through the (idle) X1 tile for a coalesced store
            }
            tc::tmem_wait_st();
            __syncthreads();
            tile_to_global<H>(sX1, gtile, GS, n_valid, tid);
            PROF(12);
        }
        if (!BWD && a.h_final && valid) store8_global(a.h_final + grow * H + j0, h);
        __syncthreads();           // the next tile's first fetch overwrites the staging region
    }
    if (BWD) {
        float t8[CPT];
        for (int o = 0; o < kTmemS && o < a.NO; ++o) {
            tc::tmem_ld8(acc_addr + 8 * o, t8);
            colsum8(t8, acc + ACC_S + o * H + j0, lane);
        }
        tc::tmem_ld8(acc_addr + 40, t8);
        colsum8(t8, acc + ACC_G2 + j0, lane);
        tc::tmem_ld8(acc_addr + 48, t8);
        colsum8(t8, acc + ACC_G1 + j0, lane);
        if (part == 0) {
            tc::tmem_ld8(acc_addr + 56, t8);
            colsum8(t8, acc + ACC_BO, lane);
        }
        __syncthreads();
        float* dst = a.gacc + (size_t)blockIdx.x * ACC_N;
        for (int i = tid; i < ACC_N; i += kThreads)
            dst[i] = (sAcc[i] + sAcc[ACC_N + i]) + (sAcc[2 * ACC_N + i] + sAcc[3 * ACC_N + i]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kTmemCols);
}

size_t trunk_smem(bool bwd) {
    const Layout L = make_layout();
    size_t s = (size_t)4 * (H * K1 + H * H + 6 * H * H + kRows * K1 + 3 * kRows * H + 2 * kRows * 8 + (L.total - L.small)) + 256 + kRows * 8;
    if (bwd) s += (size_t)4 * (H * H + 6 * H * H + kRows * 4 * H + 4 * ACC_N + 2 * kRows * NOMAX);
    else     s += (size_t)4 * (kRows * H + kRows * 8);   // staging tile, row tails
    return s;
}

int check_common(const float* params, const float* obs, int32_t L, int64_t R, int32_t D, int32_t NO) {
    if (!params || !obs || L <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (D < 1 || D > K1 || NO < 1 || NO > NOMAX) return MM_ERR_UNSUPPORTED;
    if (!mm::aligned(obs, 4)) return MM_ERR_BAD_ARG;
    return MM_OK;
}

long long* g_prof = nullptr;

int n_ctas(long long tiles) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)(tiles < sms ? tiles : sms);
}

}  // namespace

extern "C" int mm_ppo_trunk_param_count(void) { return make_layout().total; }

// offsets[18]: w1, w2, wih, whh, small, b1, ln1_w, ln1_b, b2, ln2_w, ln2_b, bih, bhh, ln3_w, ln3_b, wo, bo, total
extern "C" int mm_ppo_trunk_param_offsets(int32_t* o) {
    if (!o) return MM_ERR_BAD_ARG;
    const Layout L = make_layout();
    const int v[18] = {L.w1, L.w2, L.wih, L.whh, L.small, L.b1, L.ln1_w, L.ln1_b, L.b2, L.ln2_w, L.ln2_b, L.bih, L.bhh,
                       L.ln3_w, L.ln3_b, L.wo, L.bo, L.total};
    for (int i = 0; i < 18; ++i) o[i] = v[i];
    return MM_OK;
}

// dims[8]: X row stride, G row stride, accumulator count, offsets of S, dbo, dgamma2, dgamma1 in the accumulator, NOMAX
extern "C" int mm_ppo_trunk_dims(int32_t* d) {
    if (!d) return MM_ERR_BAD_ARG;
    const int v[8] = {XS, GS, ACC_N, ACC_S, ACC_BO, ACC_G2, ACC_G1, NOMAX};
    for (int i = 0; i < 8; ++i) d[i] = v[i];
    return MM_OK;
}

// Debug: device buffer of 16 int64 into which CTA 0 of later launches accumulates per-phase clock64() deltas (NULL = off).
extern "C" int mm_ppo_trunk_set_prof(long long* dev_buf16) { g_prof = dev_buf16; return MM_OK; }

extern "C" int mm_ppo_trunk_max_ctas(void) { return n_ctas(1 << 30); }

extern "C" int mm_ppo_trunk_fwd(const float* params, const float* obs, const int64_t* obs_row, int64_t obs_step_rows,
                                const float* h0, const float* masks, float* out, float* h_final, float* X, int32_t L, int64_t R,
                                int32_t D, int32_t NO, void* stream) {
    int rc = check_common(params, obs, L, R, D, NO);
    if (rc != MM_OK) return rc;
    if (!h0 || !out || !mm::aligned(h0, 16)) return MM_ERR_BAD_ARG;
    if ((X && !mm::aligned(X, 16)) || (h_final && !mm::aligned(h_final, 16))) return MM_ERR_BAD_ARG;
    Args a{};
    a.params = params; a.obs = obs; a.h0 = h0; a.masks = masks; a.out = out; a.h_final = h_final; a.X = X;
    a.obs_row = reinterpret_cast<const long long*>(obs_row); a.obs_step = obs_step_rows;
    a.prof = g_prof;
    a.L = L; a.D = D; a.NO = NO; a.R = R; a.n_tiles = (R + kRows - 1) / kRows;
    const size_t smem = trunk_smem(false);
    cudaError_t e = cudaFuncSetAttribute(ppo_trunk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    ppo_trunk_kernel<false><<<n_ctas(a.n_tiles), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
    return mm::check_launch();
}

extern "C" int mm_ppo_trunk_bwd(const float* params, const int64_t* obs_row, int64_t obs_step_rows, const float* masks,
                                const float* X, const float* dout,
                                float* G, float* acc_partials, int32_t n_partials, int32_t L, int64_t R, int32_t D,
                                int32_t NO, void* stream) {
    int rc = check_common(params, X, L, R, D, NO);
    if (rc != MM_OK) return rc;
    if (!X || !dout || !G || !acc_partials || !mm::aligned(X, 16) || !mm::aligned(G, 16)) return MM_ERR_BAD_ARG;
    Args a{};
    a.params = params; a.masks = masks; a.X = const_cast<float*>(X); a.dout = dout; a.G = G;
    a.obs_row = reinterpret_cast<const long long*>(obs_row); a.obs_step = obs_step_rows;
    a.gacc = acc_partials; a.prof = g_prof;
    a.L = L; a.D = D; a.NO = NO; a.R = R; a.n_tiles = (R + kRows - 1) / kRows;
    const int ctas = n_ctas(a.n_tiles);
    if (n_partials < ctas) return MM_ERR_BAD_ARG;
    const size_t smem = trunk_smem(true);
    if (smem > 227 * 1024) return MM_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(ppo_trunk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    ppo_trunk_kernel<true><<<ctas, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
    return mm::check_launch();
}
