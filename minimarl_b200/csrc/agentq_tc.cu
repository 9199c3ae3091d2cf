// Tensor-core (tcgen05, TF32 in / fp32 accumulate in TMEM) fused recurrent agent network forward.
//
// Same computation as agentq.cu (AgentQFunction.forward,
// offpolicy/algorithms/qmix/algorithm/agent_q_function.py:34-66) with every dense contraction on the
// 5th-generation tensor cores:
//     fc1   [128 x K1] x [K1 x H]      -> TMEM cols [0, H)
//     fc2   [128 x H ] x [H  x H]      -> TMEM cols [0, H)            (fc1's accumulator is consumed by then)
//     gates [128 x H ] x [H x 2H] (x)  + [128 x H] x [H x 2H] (h)  -> TMEM cols [H, 3H)   r | z, accumulated
//           [128 x H ] x [H x H ] (x)  -> TMEM cols [0, H)            i_n
//           [128 x H ] x [H x H ] (h)  -> TMEM cols [3H, 4H)          h_n
// One CTA = 128 rows, 512 threads: the four warps that may address the same 32 TMEM lanes split the
// H columns of those rows (H/4 each), so bias, ReLU and the GRU gate math are thread-local on the
// accumulators read back with tcgen05.ld, LayerNorm needs one 4-way exchange through shared memory, and the
// activations go back to shared memory directly in the K-major operand layout of the next MMA.  Gates use
// MUFU tanh (sigmoid(x) = 0.5 tanh(x/2) + 0.5): 3 MUFU per hidden unit instead of ~75 FP32 instructions.  Weights (W1, W2, W_ih, W_hh = 124 KB as tf32 at H = 64) stay in shared memory for the whole
// sequence, the hidden state lives in shared memory in operand layout, and the obs tile of step t+1 is
// prefetched with cp.async while step t computes.  MMAs are issued by one thread and tracked with one
// mbarrier (tcgen05.commit); waits are bounded (trap instead of hang).
// Precision: TF32 operands (round-to-nearest on every activation store and on the weights), fp32
// accumulation and fp32 epilogues -> ~1e-3 relative on the Q-values; tests state the tolerance.
#include "checkers_core.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;
constexpr float kEps = 1e-5f;

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// plain parameter buffer (fp32, PyTorch layouts): offsets in floats
struct TcLayout {
    int D, K1, H, NA;
    int ln0_w, ln0_b, w1, b1, ln1_w, ln1_b, w2, b2, ln2_w, ln2_b, wih, whh, bih, bhh, ln3_w, ln3_b, wq, bq, total;
    int small;   // everything from here on (vectors + the head) is copied to shared memory as is
};
__host__ __device__ inline TcLayout make_tc_layout(int D, int H, int NA) {
    TcLayout L;
    L.D = D; L.H = H; L.NA = NA; L.K1 = (D + 7) & ~7;
    int o = 0;
    auto take = [&](int n) { int at = o; o += (n + 3) & ~3; return at; };
    // the four operand matrices first (PyTorch [out][in] layouts; w1 zero-padded to K1 columns) ...
    L.w1 = take(H * L.K1); L.w2 = take(H * H); L.wih = take(3 * H * H); L.whh = take(3 * H * H);
    // ... then the small vectors and the head
    L.small = o;
    L.ln0_w = take(D); L.ln0_b = take(D);
    L.b1 = take(H); L.ln1_w = take(H); L.ln1_b = take(H);
    L.b2 = take(H); L.ln2_w = take(H); L.ln2_b = take(H);
    L.bih = take(3 * H); L.bhh = take(3 * H);
    L.ln3_w = take(H); L.ln3_b = take(H);
    L.wq = take(NA * H); L.bq = take(NA);
    L.total = o;
    return L;
}

// MUFU-based activations: tanh.approx has ~2^-11 relative error, the same order as the TF32 operands
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// 512 threads: thread (row = 32*(warp%4) + lane, part = warp/4) owns CPT = H/4 consecutive columns of its row.
// (A warp may only touch TMEM lanes 32*(warp%4)..+31, so the four warps that share `warp%4` split the columns.)
template <int H>
struct Geo {
    static constexpr int CPT = H / 4;
};

// LayerNorm over a row whose H values are spread over 4 threads: partial (sum, sum of squares) through
// shared memory, one barrier.  var = E[x^2] - mean^2 in fp32 (inputs are O(1) post-ReLU / gate outputs).
template <int H>
__device__ __forceinline__ void layer_norm_split(float (&v)[H / 4], float2* red, int row, int part,
                                                 const float* __restrict__ w, const float* __restrict__ b) {
    constexpr int CPT = H / 4;
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) { s += v[j]; q = fmaf(v[j], v[j], q); }
    red[row * 4 + part] = make_float2(s, q);
    __syncthreads();
    const float4 a = *reinterpret_cast<const float4*>(red + row * 4);
    const float4 c = *reinterpret_cast<const float4*>(red + row * 4 + 2);
    const float mean = (a.x + a.z + c.x + c.z) * (1.0f / H);
    const float var = fmaxf((a.y + a.w + c.y + c.w) * (1.0f / H) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + kEps);
#pragma unroll
    for (int j = 0; j < CPT; ++j) v[j] = fmaf((v[j] - mean) * rstd, w[part * CPT + j], b[part * CPT + j]);
}

template <int N>
__device__ __forceinline__ void store_cols_kmajor(uint8_t* tile, int r, int k0, int K, const float (&v)[N], bool round) {
#pragma unroll
    for (int k = 0; k < N; k += 4)
        *reinterpret_cast<float4*>(tile + tc::kmajor_offset(r, k0 + k, K)) =
            round ? make_float4(to_tf32(v[k]), to_tf32(v[k + 1]), to_tf32(v[k + 2]), to_tf32(v[k + 3]))
                  : make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
}

constexpr int kThreadsTc = 512;
constexpr int kMaxNaEnv = 8;

// Device-side copy of mm_rollout_io plus the env configuration (only read by the ENV instantiation).
struct RolloutDev {
    mm_rollout_io io;
    mm_checkers_cfg cfg;
    float common[3][3];     // float32(python_double(reward_0 + reward_1)) by (food kind of agent 0, of agent 1)
    uint2 seed;
};

// ENV == false: the sequence of observations is given (obs [T1, R, D]), pure network forward.
// ENV == true : fused policy + env rollout.  Rows are (env, agent) pairs, row = 2*env + agent, so a CTA owns 64
//   Checkers boards; after the head of step t the first four warps pick eps-greedy actions (Philox), step their
//   boards with the shuffle-based transition of checkers_core.cuh and expand the next observations straight into
//   the shared obs tile, which is both streamed to obs[t+1] in HBM and consumed by step t+1's LayerNorm.
template <int H, bool ENV>
__global__ void __launch_bounds__(kThreadsTc, 1)
agentq_tc_kernel(const float* __restrict__ params, const float* __restrict__ obs, const float* __restrict__ h0,
                 float* __restrict__ q_out, float* __restrict__ h_final, float* __restrict__ rnn_out, int T1,
                 long long R, int D, int NA, const __grid_constant__ RolloutDev ro) {
    constexpr int CPT = Geo<H>::CPT;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const TcLayout L = make_tc_layout(D, H, NA);
    const int K1 = L.K1;
    // ---- shared memory carve-up ------------------------------------------------------------------
    uint8_t* sW1 = smem;                                  // [H  x K1] tf32, K-major
    uint8_t* sW2 = sW1 + H * K1 * 4;                      // [H  x H ]
    uint8_t* sWih = sW2 + H * H * 4;                      // [3H x H ]
    uint8_t* sWhh = sWih + 3 * H * H * 4;                 // [3H x H ]
    const int act_k = K1 > H ? K1 : H;
    uint8_t* sA = sWhh + 3 * H * H * 4;                   // activation operand tile [128 x max(K1, H)]; also sYt
    uint8_t* sH = sA + kRows * act_k * 4;                 // hidden state [128 x H], operand layout, fp32
    float* sObs = reinterpret_cast<float*>(sH + kRows * H * 4);   // staged obs tile [128][D]
    float2* red = reinterpret_cast<float2*>(sObs + ((kRows * D + 3) & ~3));   // [128][4] LayerNorm partials
    // small params: sP is biased so that sP[L.x] addresses the copy of params[L.x] for every vector / head offset
    float* sP = reinterpret_cast<float*>(red + kRows * 4) - L.small;
    float* sYt = reinterpret_cast<float*>(sA);            // [NA][4][128] partial head sums (aliases the activation tile)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane, part = warp >> 2;
    const int c0 = part * CPT;
    const long long row0 = (long long)blockIdx.x * kRows;
    const int n_valid = (int)min((long long)kRows, R - row0);
    const bool valid = row < n_valid;
    constexpr uint32_t kCols = 4 * H <= 32 ? 32 : (4 * H <= 64 ? 64 : (4 * H <= 128 ? 128 : 256));

    if (warp == 0) tc::tmem_alloc(&tmem_slot, kCols);
    if (tid == 0) tc::mbar_init(&bar, 1);

    // weights -> operand layout (tf32-rounded once), small vectors -> sP (same offsets as `params`)
    auto stage_matrix = [&](uint8_t* dst, int off, int rows, int K) {
        for (int i = tid; i < rows * K; i += kThreadsTc) {
            const int r = i / K, k = i - r * K;
            *reinterpret_cast<float*>(dst + tc::kmajor_offset(r, k, K)) = to_tf32(params[off + i]);
        }
    };
    stage_matrix(sW1, L.w1, H, K1);
    stage_matrix(sW2, L.w2, H, H);
    stage_matrix(sWih, L.wih, 3 * H, H);
    stage_matrix(sWhh, L.whh, 3 * H, H);
    for (int i = L.small + tid; i < L.total; i += kThreadsTc) sP[i] = params[i];
    // hidden state: kept unrounded (the tensor core only truncates it as an operand; the z*h term sees fp32)
#pragma unroll
    for (int k = 0; k < CPT; k += 4) {
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h0 != nullptr && valid) h = *reinterpret_cast<const float4*>(h0 + (row0 + row) * H + c0 + k);
        *reinterpret_cast<float4*>(sH + tc::kmajor_offset(row, c0 + k, H)) = h;
    }
    auto prefetch_obs = [&](int t) {
        const float* src = obs + ((long long)t * R + row0) * D;
        const int n = n_valid * D;
        for (int i = tid; i < n; i += kThreadsTc)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(sObs + i)), "l"(src + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    ck::EnvRegs env_regs;
    env_regs.food = env_regs.cells = env_regs.step = 0;
    const long long env_id = (row0 + row) >> 1;
    const long long n_envs = R >> 1;
    if (ENV) {
        if (part == 0) {
            const long long e = valid ? env_id : n_envs - 1;
            if (ro.io.do_reset) ck::reset_regs(env_regs);
            else {
                env_regs.food = ro.io.env_state[e];
                env_regs.cells = ro.io.env_state[n_envs + e];
                env_regs.step = ro.io.env_state[2 * n_envs + e];
            }
        }
        if (ro.io.do_reset) {
            if (part == 0) ck::fill_tile(sObs, row, row & 1, env_regs, ro.cfg);
            __syncthreads();
            float* dst = ro.io.obs + row0 * D;           // slab 0
            for (int i = tid; i < n_valid * D; i += kThreadsTc) dst[i] = sObs[i];
        } else {
            prefetch_obs(0);
        }
    } else {
        prefetch_obs(0);
    }

    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t phase = 0;

    const uint32_t idesc_h = tc::make_idesc_tf32(128, H);
    const uint32_t idesc_2h = tc::make_idesc_tf32(128, 2 * H);
    auto desc = [&](const uint8_t* base, int row_off, int K) {
        // tile with K columns; operand starts at row `row_off` (multiple of 8); k-step kk adds kk * 256 B, i.e.
        // kk * 16 to the (address >> 4) field in the low bits of the descriptor
        return tc::make_desc(tc::smem_u32(base) + (uint32_t)(row_off >> 3) * (uint32_t)(K >> 2) * 128u, 128u,
                             (uint32_t)(K >> 2) * 128u);
    };
    // loop-invariant operand descriptors (only thread 0 uses them)
    const uint64_t dA1 = desc(sA, 0, K1), dW1 = desc(sW1, 0, K1);
    const uint64_t dA = desc(sA, 0, H), dHid = desc(sH, 0, H), dW2 = desc(sW2, 0, H);
    const uint64_t dWih_rz = desc(sWih, 0, H), dWih_n = desc(sWih, 2 * H, H);
    const uint64_t dWhh_rz = desc(sWhh, 0, H), dWhh_n = desc(sWhh, 2 * H, H);
    auto publish = [&]() {            // activation tile written -> visible to the tensor core, all threads done
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    };
    auto wait_mma = [&]() {           // warp 0 polls the mbarrier, the other 15 warps sleep in the block barrier
        if (warp == 0) tc::mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        tc::fence_after_sync();
    };
    // K1 columns of the obs row are split over the 4 parts in chunks of 4 (k = 4*(part + 4*i) ...)
    const int kq = K1 >> 2;           // float4 groups per row

    for (int t = 0; t < T1; ++t) {
        // ---- A. LayerNorm(D) of the obs rows -> sA [128 x K1] ----------------------------------------------
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                   // obs tile landed; previous step's head is done with sYt
        {
            const float* o = sObs + row * D;
            float s = 0.f, qv = 0.f;
            for (int g = part; g < kq; g += 4)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = 4 * g + u;
                    const float x = (valid && k < D) ? o[k] : 0.f;
                    s += x; qv = fmaf(x, x, qv);
                }
            red[row * 4 + part] = make_float2(s, qv);
            __syncthreads();
            const float4 a = *reinterpret_cast<const float4*>(red + row * 4);
            const float4 c = *reinterpret_cast<const float4*>(red + row * 4 + 2);
            const float mean = (a.x + a.z + c.x + c.z) / (float)D;
            const float var = fmaxf((a.y + a.w + c.y + c.w) / (float)D - mean * mean, 0.f);
            const float rstd = rsqrtf(var + kEps);
            for (int g = part; g < kq; g += 4) {
                float x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = 4 * g + u;
                    x[u] = k < D ? to_tf32(fmaf(((valid ? o[k] : 0.f) - mean) * rstd, sP[L.ln0_w + k], sP[L.ln0_b + k])) : 0.f;
                }
                *reinterpret_cast<float4*>(sA + tc::kmajor_offset(row, 4 * g, K1)) = make_float4(x[0], x[1], x[2], x[3]);
            }
        }
        publish();                                         // also: every thread is done reading sObs and `red`
        if (!ENV && t + 1 < T1) prefetch_obs(t + 1);
        if (tid == 0) {
            for (int kk = 0; kk < K1 / 8; ++kk)
                tc::mma_tf32(tmem, dA1 + 16 * kk, dW1 + 16 * kk, idesc_h, kk > 0);
            tc::mma_commit(&bar);
            // the hidden-state halves of the gates only need h_{t-1}: issued now, behind fc1 and outside its commit, they run under
            // the fc1 / fc2 epilogues instead of in front of the gate phase (columns [H, 4H) are not touched until then)
#pragma unroll
            for (int kk = 0; kk < H / 8; ++kk)
                tc::mma_tf32(tmem + H, dHid + 16 * kk, dWhh_rz + 16 * kk, idesc_2h, kk > 0);
#pragma unroll
            for (int kk = 0; kk < H / 8; ++kk)
                tc::mma_tf32(tmem + 3 * H, dHid + 16 * kk, dWhh_n + 16 * kk, idesc_h, kk > 0);
        }
        wait_mma();

        // ---- B. fc1 epilogue: bias, ReLU, LayerNorm -> sA [128 x H] -----------------------------------------
        float v[CPT];
        tc::tmem_ld<CPT>(lane_addr + c0, v);
#pragma unroll
        for (int j = 0; j < CPT; ++j) v[j] = fmaxf(v[j] + sP[L.b1 + c0 + j], 0.f);
        layer_norm_split<H>(v, red, row, part, sP + L.ln1_w, sP + L.ln1_b);
        store_cols_kmajor<CPT>(sA, row, c0, H, v, true);
        publish();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < H / 8; ++kk)
                tc::mma_tf32(tmem, dA + 16 * kk, dW2 + 16 * kk, idesc_h, kk > 0);
            tc::mma_commit(&bar);
        }
        wait_mma();

        // ---- C. fc2 epilogue -> sA; then all gate GEMMs ------------------------------------------------------
        tc::tmem_ld<CPT>(lane_addr + c0, v);
#pragma unroll
        for (int j = 0; j < CPT; ++j) v[j] = fmaxf(v[j] + sP[L.b2 + c0 + j], 0.f);
        layer_norm_split<H>(v, red, row, part, sP + L.ln2_w, sP + L.ln2_b);
        store_cols_kmajor<CPT>(sA, row, c0, H, v, true);
        publish();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < H / 8; ++kk)         // r | z : x part on top of the h part issued at the start of the step
                tc::mma_tf32(tmem + H, dA + 16 * kk, dWih_rz + 16 * kk, idesc_2h, true);
#pragma unroll
            for (int kk = 0; kk < H / 8; ++kk)         // i_n  (h_n: issued at the start of the step)
                tc::mma_tf32(tmem, dA + 16 * kk, dWih_n + 16 * kk, idesc_h, kk > 0);
            tc::mma_commit(&bar);
        }
        wait_mma();

        // ---- D. GRU gates (PyTorch order r, z, n), new hidden -> sH, LayerNorm, head -------------------------
        {
            float gr[CPT], gz[CPT], gin[CPT], ghn[CPT];
            tc::tmem_ld_issue<CPT>(lane_addr + H + c0, gr);          // four loads in flight, one wait
            tc::tmem_ld_issue<CPT>(lane_addr + 2 * H + c0, gz);
            tc::tmem_ld_issue<CPT>(lane_addr + c0, gin);
            tc::tmem_ld_issue<CPT>(lane_addr + 3 * H + c0, ghn);
            tc::tmem_ld_fence<CPT>(gr, gz, gin, ghn);
#pragma unroll
            for (int j4 = 0; j4 < CPT; j4 += 4) {
                const float4 hp = *reinterpret_cast<const float4*>(sH + tc::kmajor_offset(row, c0 + j4, H));
                const float hprev[4] = {hp.x, hp.y, hp.z, hp.w};
                const int j = c0 + j4;
                const float4 bir = *reinterpret_cast<const float4*>(sP + L.bih + j), bhr = *reinterpret_cast<const float4*>(sP + L.bhh + j);
                const float4 biz = *reinterpret_cast<const float4*>(sP + L.bih + H + j), bhz = *reinterpret_cast<const float4*>(sP + L.bhh + H + j);
                const float4 bin = *reinterpret_cast<const float4*>(sP + L.bih + 2 * H + j), bhn = *reinterpret_cast<const float4*>(sP + L.bhh + 2 * H + j);
                const float br[4] = {bir.x + bhr.x, bir.y + bhr.y, bir.z + bhr.z, bir.w + bhr.w};
                const float bz[4] = {biz.x + bhz.x, biz.y + bhz.y, biz.z + bhz.z, biz.w + bhz.w};
                const float bi[4] = {bin.x, bin.y, bin.z, bin.w}, bh[4] = {bhn.x, bhn.y, bhn.z, bhn.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float rr = sigmoid_fast(gr[j4 + u] + br[u]);
                    const float zz = sigmoid_fast(gz[j4 + u] + bz[u]);
                    const float nn = tanh_fast(gin[j4 + u] + bi[u] + rr * (ghn[j4 + u] + bh[u]));
                    v[j4 + u] = fmaf(zz, hprev[u] - nn, nn);          // (1 - z) n + z h
                }
            }
        }
        // the MMAs that read sH have completed (we waited on their commit), so it is overwritten in place
        store_cols_kmajor<CPT>(sH, row, c0, H, v, false);
        if (t == T1 - 1 && h_final != nullptr && valid) {
#pragma unroll
            for (int k = 0; k < CPT; k += 4)
                *reinterpret_cast<float4*>(h_final + (row0 + row) * H + c0 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
        }
        layer_norm_split<H>(v, red, row, part, sP + L.ln3_w, sP + L.ln3_b);
        if (rnn_out != nullptr && valid) {
            float* dst = rnn_out + ((long long)t * R + row0 + row) * H + c0;
#pragma unroll
            for (int k = 0; k < CPT; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
        }
        if (ENV || q_out != nullptr) {
            // head: every thread dots its CPT columns with all NA rows of W_q; the 4 partials of a row meet in
            // the (now free) activation tile and part 0 adds them up
            for (int a = 0; a < NA; ++a) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; j += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(sP + L.wq + a * H + c0 + j);
                    acc = fmaf(v[j], w.x, fmaf(v[j + 1], w.y, fmaf(v[j + 2], w.z, fmaf(v[j + 3], w.w, acc))));
                }
                sYt[(a * 4 + part) * kRows + row] = acc;
            }
            __syncthreads();
            if (part == 0) {
                float qv[ENV ? kMaxNaEnv : 1];
                int greedy = 0;
                float best = 0.f;
                for (int a = 0; a < NA; ++a) {
                    const float qa = sP[L.bq + a] + ((sYt[(a * 4 + 0) * kRows + row] + sYt[(a * 4 + 1) * kRows + row]) +
                                                     (sYt[(a * 4 + 2) * kRows + row] + sYt[(a * 4 + 3) * kRows + row]));
                    if (valid && q_out != nullptr) q_out[((long long)t * R + row0 + row) * NA + a] = qa;
                    if (ENV) {
                        if (a < kMaxNaEnv) qv[a] = qa;
                        if (a == 0 || qa > best) { best = qa; greedy = a; }      // first maximum, like torch.max
                    }
                }
                if (ENV) {
                    // ---- eps-greedy action (QMixPolicy.actions_from_q: one uniform per agent row) ----------------
                    const long long grow = row0 + row;
                    const unsigned long long ctr = ro.io.step_offset + (unsigned long long)t;
                    const long long key_row = ro.io.eps_per_env ? (grow >> 1) : grow;     // vdn/_network.py: one draw per env
                    const uint4 rnd = mm::philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)key_row,
                                                                   (uint32_t)(key_row >> 32)), ro.seed);
                    const uint4 rnd2 = mm::philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)grow,
                                                                    0x80000000u | (uint32_t)(grow >> 32)), ro.seed);
                    const bool take_random = (float)(rnd.x >> 8) * (1.0f / 16777216.0f) < ro.io.eps;
                    const int act = take_random ? (int)__umulhi(rnd2.x, (uint32_t)NA) : greedy;
                    // ---- env step (frozen once the episode has finished, like collect_rollout's `if not all(dones)`) ----
                    const bool frozen = ro.io.freeze_after_done && (env_regs.food & ck::kDoneLatch);
                    ck::EnvRegs next = env_regs;
                    float reward;
                    uint32_t done;
                    bool bad;
                    const int agent = row & 1;
                    // food kind eaten, recovered from the reward LUT index: recompute via before/after food bits
                    const uint32_t food_before = env_regs.food & ck::kFoodCells;
                    const bool latched = (env_regs.food & ck::kDoneLatch) != 0;   // steps_beyond_done: rewards are [0, 0]
                    ck::transition(next, act, agent, lane, ro.cfg, reward, done, bad);
                    if (frozen) { reward = 0.f; done = 1u; } else { env_regs = next; }
                    if (ro.io.common_reward) {
                        // kinds: which cell this agent ate (0 none, 1 lemon, 2 apple) -> LUT of exactly rounded sums
                        const uint32_t mine_cell = (env_regs.cells >> (8 * agent)) & 0xffu;
                        const uint32_t ate = frozen ? 0u : (food_before & ~env_regs.food & (1u << mine_cell));
                        const int kind = ate ? ((ate & ck::kAppleCells) ? 2 : 1) : 0;
                        const int other = __shfl_xor_sync(MM_FULL_MASK, kind, 1);
                        reward = (latched || frozen) ? 0.f : (agent == 0 ? ro.common[kind][other] : ro.common[other][kind]);
                    }
                    if (valid) {
                        const long long ti = (long long)t;
                        if (ro.io.act_idx) ro.io.act_idx[ti * ro.io.act_t_stride + grow] = act;
                        if (ro.io.act_onehot) {
                            float* oh = ro.io.act_onehot + ti * ro.io.onehot_t_stride + grow * NA;
                            for (int a = 0; a < NA; ++a) oh[a] = (a == act) ? 1.0f : 0.0f;
                        }
                        if (ro.io.rewards) ro.io.rewards[ti * ro.io.rew_t_stride + grow] = reward;
                        if (ro.io.dones) ro.io.dones[ti * ro.io.done_t_stride + grow] = done ? 1.0f : 0.0f;
                        if (ro.io.dones_env && agent == 0) ro.io.dones_env[ti * ro.io.denv_t_stride + (grow >> 1)] = done ? 1.0f : 0.0f;
                    }
                    // next observation straight into the shared obs tile (nobody reads it until the next barrier)
                    ck::fill_tile(sObs, row, agent, env_regs, ro.cfg);
                }
            }
            if (ENV) {
                __syncthreads();
                float* dst = ro.io.obs + (long long)(t + 1) * ro.io.obs_t_stride + row0 * D;
                const int n = n_valid * D;
                if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                    for (int i = tid; i < (n >> 2); i += kThreadsTc)
                        mm::st_stream(reinterpret_cast<float4*>(dst) + i, reinterpret_cast<const float4*>(sObs)[i]);
                } else {
                    for (int i = tid; i < n; i += kThreadsTc) dst[i] = sObs[i];
                }
            }
        }
        // stage A of the next step starts with a barrier before it overwrites sA / sYt
    }

    if (ENV && part == 0 && valid && (row & 1) == 0) {
        ro.io.env_state[env_id] = env_regs.food;
        ro.io.env_state[n_envs + env_id] = env_regs.cells;
        ro.io.env_state[2 * n_envs + env_id] = env_regs.step;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kCols);
}

template <int H>
size_t tc_smem_bytes(int D, int NA) {
    const TcLayout L = make_tc_layout(D, H, NA);
    const int act_k = L.K1 > H ? L.K1 : H;
    return (size_t)4 * (H * L.K1 + H * H + 6 * H * H + kRows * act_k + kRows * H + ((kRows * D + 3) & ~3) +
                        kRows * 8 + (L.total - L.small)) + 128;
}

template <int H>
int launch_tc(const float* params, const float* obs, const float* h0, float* q, float* h_final, float* rnn_out,
              int T1, long long R, int D, int NA, cudaStream_t st) {
    const size_t smem = tc_smem_bytes<H>(D, NA);
    if (smem > 227 * 1024) return MM_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(agentq_tc_kernel<H, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    const long long tiles = (R + kRows - 1) / kRows;
    RolloutDev none{};
    agentq_tc_kernel<H, false><<<(unsigned)tiles, kThreadsTc, smem, st>>>(params, obs, h0, q, h_final, rnn_out, T1, R, D, NA, none);
    return mm::check_launch();
}

template <int H>
int launch_rollout(const float* params, const float* h0, float* h_final, float* q, int T, long long R, int NA,
                   const RolloutDev& ro, cudaStream_t st) {
    const int D = ck::kObs;
    const size_t smem = tc_smem_bytes<H>(D, NA);
    if (smem > 227 * 1024) return MM_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(agentq_tc_kernel<H, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    const long long tiles = (R + kRows - 1) / kRows;
    agentq_tc_kernel<H, true><<<(unsigned)tiles, kThreadsTc, smem, st>>>(params, ro.io.obs, h0, q, h_final, nullptr, T, R, D, NA, ro);
    return mm::check_launch();
}

}  // namespace

extern "C" int mm_agentq_tc_param_count(int32_t D, int32_t H, int32_t NA) {
    if (D <= 0 || D > 96 || (H != 32 && H != 64) || NA <= 0 || NA > 16) return MM_ERR_UNSUPPORTED;
    return make_tc_layout(D, H, NA).total;
}

// offsets[18]: ln0_w, ln0_b, w1, b1, ln1_w, ln1_b, w2, b2, ln2_w, ln2_b, wih, whh, bih, bhh, ln3_w, ln3_b, wq, bq
extern "C" int mm_agentq_tc_param_offsets(int32_t D, int32_t H, int32_t NA, int32_t* offsets_host) {
    if (!offsets_host || mm_agentq_tc_param_count(D, H, NA) < 0) return MM_ERR_UNSUPPORTED;
    const TcLayout L = make_tc_layout(D, H, NA);
    const int o[18] = {L.ln0_w, L.ln0_b, L.w1, L.b1, L.ln1_w, L.ln1_b, L.w2, L.b2, L.ln2_w, L.ln2_b,
                       L.wih, L.whh, L.bih, L.bhh, L.ln3_w, L.ln3_b, L.wq, L.bq};
    for (int i = 0; i < 18; ++i) offsets_host[i] = o[i];
    return MM_OK;
}

extern "C" int mm_agentq_tc_forward(const float* params, const float* obs, const float* h0, float* q,
                                    float* h_final, float* rnn_out, int32_t T1, int64_t R, int32_t D, int32_t H,
                                    int32_t NA, void* stream) {
    if (!params || !obs || T1 <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (mm_agentq_tc_param_count(D, H, NA) < 0) return MM_ERR_UNSUPPORTED;
    if (rnn_out && !mm::aligned(rnn_out, 16)) return MM_ERR_BAD_ARG;
    if (h0 && !mm::aligned(h0, 16)) return MM_ERR_BAD_ARG;
    if (h_final && !mm::aligned(h_final, 16)) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (H == 64) return launch_tc<64>(params, obs, h0, q, h_final, rnn_out, T1, R, D, NA, st);
    return launch_tc<32>(params, obs, h0, q, h_final, rnn_out, T1, R, D, NA, st);
}

extern "C" int mm_qmix_rollout_tc(const float* params, const float* h0, float* h_final, float* q_out,
                                  const mm_rollout_io* io, const mm_checkers_cfg* cfg,
                                  const float* common_reward_lut_host, int32_t T, int64_t n_envs, int32_t H,
                                  int32_t NA, uint64_t seed, void* stream) {
    if (!params || !io || !cfg || !io->env_state || !io->obs || T <= 0 || n_envs <= 0) return MM_ERR_BAD_ARG;
    if (cfg->full_observable || NA > kMaxNaEnv) return MM_ERR_UNSUPPORTED;
    if (io->common_reward && !common_reward_lut_host) return MM_ERR_BAD_ARG;
    if (mm_agentq_tc_param_count(ck::kObs, H, NA) < 0) return MM_ERR_UNSUPPORTED;
    if (h0 && !mm::aligned(h0, 16)) return MM_ERR_BAD_ARG;
    if (h_final && !mm::aligned(h_final, 16)) return MM_ERR_BAD_ARG;
    RolloutDev ro;
    ro.io = *io;
    ro.cfg = *cfg;
    ro.cfg.n_envs = (int32_t)n_envs;
    for (int i = 0; i < 9; ++i) (&ro.common[0][0])[i] = common_reward_lut_host ? common_reward_lut_host[i] : 0.f;
    ro.seed = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long R = 2 * (long long)n_envs;
    if (H == 64) return launch_rollout<64>(params, h0, h_final, q_out, T, R, NA, ro, st);
    return launch_rollout<32>(params, h0, h_final, q_out, T, R, NA, ro, st);
}
