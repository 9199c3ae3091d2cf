// QMIX hypernetwork mixer forward on tcgen05 (TF32 operands, fp32 accumulators in TMEM), A = 2 agents.
//
// Same function as mixer.cu (QMixer.forward, offpolicy/algorithms/qmix/algorithm/q_mixer.py:68-94, hypernet_layers == 2,
// hypernet_hidden_dim HH = 64, mixer_hidden_dim E = 32; offpolicy/config.py:158-165):
//   w1 = |Linear(HH, A E)(relu(Linear(SD, HH)(s)))|, b1 = Linear(SD, E)(s), w2 = |Linear(HH, E)(relu(Linear(SD, HH)(s)))|,
//   b2 = Linear(HH, 1)(relu(Linear(SD, HH)(s))),   Q_tot = elu(q @ w1 + b1) @ w2 + b2
// Persistent CTAs, one per SM, all hypernet weights (108 KB as tf32) resident in shared memory in operand layout; a tile is
// 128 (t, episode) rows, 512 threads = (row, column part):
//   MMA 1   S [128 x 96] x Wcat^T [96 x 224]  -> TMEM cols [0,224):  w1a | w2a | b2a | b1      (the four first layers)
//   epi 1   bias + ReLU, w1a / w2a activations back to shared memory as the next operands, b2 = relu(b2a) . wb2 reduced
//   MMA 2   Ha [128 x 64] x W1b^T -> cols [0,64) (w1), Hb [128 x 64] x W2b^T -> cols [64,96) (w2); b1 stays in cols [192,224)
//   epi 2   per row: h_e = elu(q0 |w1[0][e]| + q1 |w1[1][e]| + b1[e]), Q = sum_e h_e |w2[e]| + b2   (8 e per thread, 4-way reduce)
// The state tile of the next 128 rows is fetched with cp.async (8-byte pieces: rows are 376 B) while the current one computes.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kThreads = 512, kRows = 128;
constexpr int HH = 64, E = 32, A = 2, AE = A * E, NCAT = 3 * HH + E;      // 224
constexpr int K1 = 96;

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// parameter buffer (floats): wcat [224][96] (rows w1a | w2a | b2a | b1, K zero-padded), w1b [64][64], w2b [32][64],
// bcat [224], b1b [64], b2b [32], wb2 [64], bb2 [4]
struct Layout { int wcat, w1b, w2b, bcat, b1b, b2b, wb2, bb2, total; };
__host__ __device__ inline Layout make_layout() {
    Layout L; int o = 0;
    auto take = [&](int n) { int at = o; o += n; return at; };
    L.wcat = take(NCAT * K1); L.w1b = take(AE * HH); L.w2b = take(E * HH);
    L.bcat = take(NCAT); L.b1b = take(AE); L.b2b = take(E); L.wb2 = take(HH); L.bb2 = take(4);
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(kThreads, 1)
qmixer_tc_kernel(const float* __restrict__ params, const float* __restrict__ agent_q, const float* __restrict__ states,
                 float* __restrict__ q_tot, long long M, int SD) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const Layout L = make_layout();
    uint8_t* sWcat = smem;
    uint8_t* sW1b = sWcat + NCAT * K1 * 4;
    uint8_t* sW2b = sW1b + AE * HH * 4;
    uint8_t* sS = sW2b + E * HH * 4;
    uint8_t* sHa = sS + kRows * K1 * 4;
    uint8_t* sHb = sHa + kRows * HH * 4;
    float* sV = reinterpret_cast<float*>(sHb + kRows * HH * 4);          // bcat | b1b | b2b | wb2 | bb2
    float* sPart = sV + (L.total - L.bcat);                              // [8][128]: 4 partial b2 sums, 4 partial Q sums
    const float *vBcat = sV, *vB1b = sV + (L.b1b - L.bcat), *vB2b = sV + (L.b2b - L.bcat), *vWb2 = sV + (L.wb2 - L.bcat),
                *vBb2 = sV + (L.bb2 - L.bcat);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane, part = warp >> 2;

    if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
    if (tid == 0) tc::mbar_init(&bar, 1);
    for (int i = tid; i < NCAT * K1; i += kThreads) {
        const int r = i / K1, k = i - r * K1;
        *reinterpret_cast<float*>(sWcat + tc::kmajor_offset(r, k, K1)) = to_tf32(params[L.wcat + i]);
    }
    for (int i = tid; i < AE * HH; i += kThreads)
        *reinterpret_cast<float*>(sW1b + tc::kmajor_offset(i / HH, i % HH, HH)) = to_tf32(params[L.w1b + i]);
    for (int i = tid; i < E * HH; i += kThreads)
        *reinterpret_cast<float*>(sW2b + tc::kmajor_offset(i / HH, i % HH, HH)) = to_tf32(params[L.w2b + i]);
    for (int i = tid; i < L.total - L.bcat; i += kThreads) sV[i] = params[L.bcat + i];
    for (int i = tid; i < kRows * K1; i += kThreads) reinterpret_cast<float*>(sS)[i] = 0.f;   // K padding stays zero
    __syncthreads();

    const long long tiles = (M + kRows - 1) / kRows;
    const int pairs = SD >> 1;                                            // 8-byte pieces per state row (SD even)
    auto prefetch = [&](long long tile) {
        const long long r0 = tile * kRows;
        const int n_valid = (int)min((long long)kRows, M - r0);
        const float* src = states + r0 * SD;
        for (int u = tid; u < kRows * pairs; u += kThreads) {
            const int r = u / pairs, k = 2 * (u - r * pairs);
            cp_async8(tc::smem_u32(sS + tc::kmajor_offset(r, k, K1)), src + (r < n_valid ? (long long)r * SD + k : k));
        }
        cp_async_commit();
    };
    if (blockIdx.x < tiles) prefetch(blockIdx.x);

    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sbo1 = (uint32_t)(K1 >> 2) * 128u, sbo2 = (uint32_t)(HH >> 2) * 128u;
    const uint64_t dS = tc::make_desc(tc::smem_u32(sS), 128u, sbo1), dWcat = tc::make_desc(tc::smem_u32(sWcat), 128u, sbo1);
    const uint64_t dHa = tc::make_desc(tc::smem_u32(sHa), 128u, sbo2), dHb = tc::make_desc(tc::smem_u32(sHb), 128u, sbo2);
    const uint64_t dW1b = tc::make_desc(tc::smem_u32(sW1b), 128u, sbo2), dW2b = tc::make_desc(tc::smem_u32(sW2b), 128u, sbo2);
    const uint32_t id1 = tc::make_idesc_tf32(128, NCAT), id2a = tc::make_idesc_tf32(128, AE), id2b = tc::make_idesc_tf32(128, E);
    uint32_t phase = 0;

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * kRows;
        const bool valid = row0 + row < M;
        cp_async_wait_all();
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();                                                  // state tile landed; previous tile fully consumed
        tc::fence_after_sync();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < K1 / 8; ++kk) tc::mma_tf32(tmem, dS + 16 * kk, dWcat + 16 * kk, id1, kk > 0);
            tc::mma_commit(&bar);
        }
        float2 q = make_float2(0.f, 0.f);
        if (valid) q = *reinterpret_cast<const float2*>(agent_q + (row0 + row) * A);
        if (warp == 0) tc::mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        tc::fence_after_sync();
        if (tile + gridDim.x < tiles) prefetch(tile + gridDim.x);         // the state tile is free: fetch the next one
        // ---- epilogue 1: 16 columns of each first layer per thread ------------------------------------------------------
        {
            const int c0 = part * 16;
            float ha[16], hb[16], b2a[16];
            tc::tmem_ld_issue<16>(lane_addr + c0, ha);
            tc::tmem_ld_issue<16>(lane_addr + HH + c0, hb);
            tc::tmem_ld_issue<16>(lane_addr + 2 * HH + c0, b2a);
            tc::tmem_ld_fence<16>(ha, hb, b2a);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                ha[j] = to_tf32(fmaxf(ha[j] + vBcat[c0 + j], 0.f));
                hb[j] = to_tf32(fmaxf(hb[j] + vBcat[HH + c0 + j], 0.f));
                s = fmaf(fmaxf(b2a[j] + vBcat[2 * HH + c0 + j], 0.f), vWb2[c0 + j], s);
            }
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
                *reinterpret_cast<float4*>(sHa + tc::kmajor_offset(row, c0 + k, HH)) = make_float4(ha[k], ha[k + 1], ha[k + 2], ha[k + 3]);
                *reinterpret_cast<float4*>(sHb + tc::kmajor_offset(row, c0 + k, HH)) = make_float4(hb[k], hb[k + 1], hb[k + 2], hb[k + 3]);
            }
            sPart[part * kRows + row] = s;
        }
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < HH / 8; ++kk) tc::mma_tf32(tmem, dHa + 16 * kk, dW1b + 16 * kk, id2a, kk > 0);
#pragma unroll
            for (int kk = 0; kk < HH / 8; ++kk) tc::mma_tf32(tmem + AE, dHb + 16 * kk, dW2b + 16 * kk, id2b, kk > 0);
            tc::mma_commit(&bar);
        }
        if (warp == 0) tc::mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        tc::fence_after_sync();
        // ---- epilogue 2: mixer hidden units e in [8 part, 8 part + 8) -------------------------------------------------------
        {
            const int e0 = part * 8;
            float w10[8], w11[8], w2[8], b1[8];
            tc::tmem_ld_issue<8>(lane_addr + e0, w10);
            tc::tmem_ld_issue<8>(lane_addr + E + e0, w11);
            tc::tmem_ld_issue<8>(lane_addr + AE + e0, w2);
            tc::tmem_ld8(lane_addr + 3 * HH + e0, b1);
            tc::tmem_ld_fence<8>(w10, w11, w2);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float h = b1[j] + vBcat[3 * HH + e0 + j];
                h = fmaf(q.x, fabsf(w10[j] + vB1b[e0 + j]), h);
                h = fmaf(q.y, fabsf(w11[j] + vB1b[E + e0 + j]), h);
                h = h > 0.f ? h : expm1f(h);                               // ELU, alpha = 1
                s = fmaf(h, fabsf(w2[j] + vB2b[e0 + j]), s);
            }
            sPart[(4 + part) * kRows + row] = s;
        }
        tc::fence_before_sync();
        __syncthreads();
        if (part == 0 && valid) {
            float s = vBb2[0];
#pragma unroll
            for (int p = 0; p < 8; ++p) s += sPart[p * kRows + row];
            q_tot[row0 + row] = s;
        }
    }
    cp_async_wait_all();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// module tensors (QMixer.parameters() order: hyper_w1.0, hyper_w1.2, hyper_w2.0, hyper_w2.2, hyper_b1, hyper_b2.0, hyper_b2.2; weight
// then bias) -> the packed buffer above, one launch (the live mixer repacks after every optimiser step)
struct ModulePtrs { const float* p[14]; };
__global__ void __launch_bounds__(256) qmixer_tc_pack_kernel(const ModulePtrs q, float* __restrict__ out, int SD) {
    const Layout L = make_layout();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= L.total) return;
    float v = 0.f;
    if (i < L.w1b) {
        const int n = i / K1, k = i - n * K1;
        if (k < SD) {
            if (n < HH) v = q.p[0][n * SD + k];
            else if (n < 2 * HH) v = q.p[4][(n - HH) * SD + k];
            else if (n < 3 * HH) v = q.p[10][(n - 2 * HH) * SD + k];
            else v = q.p[8][(n - 3 * HH) * SD + k];
        }
    } else if (i < L.w2b) v = q.p[2][i - L.w1b];
    else if (i < L.bcat) v = q.p[6][i - L.w2b];
    else if (i < L.b1b) {
        const int n = i - L.bcat;
        v = n < HH ? q.p[1][n] : n < 2 * HH ? q.p[5][n - HH] : n < 3 * HH ? q.p[11][n - 2 * HH] : q.p[9][n - 3 * HH];
    } else if (i < L.b2b) v = q.p[3][i - L.b1b];
    else if (i < L.wb2) v = q.p[7][i - L.b2b];
    else if (i < L.bb2) v = q.p[12][i - L.wb2];
    else if (i == L.bb2) v = q.p[13][0];
    out[i] = v;
}

}  // namespace

extern "C" int mm_qmixer_tc_pack(const void* const* params14, float* packed, int32_t SD, int32_t n_agents, void* stream) {
    if (!params14 || !packed) return MM_ERR_BAD_ARG;
    if (SD <= 0 || SD > K1 || (SD & 1) || n_agents != A) return MM_ERR_UNSUPPORTED;
    ModulePtrs q;
    for (int i = 0; i < 14; ++i) {
        if (!params14[i]) return MM_ERR_BAD_ARG;
        q.p[i] = static_cast<const float*>(params14[i]);
    }
    qmixer_tc_pack_kernel<<<(make_layout().total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, packed, SD);
    return mm::check_launch();
}

extern "C" int mm_qmixer_tc_param_count(int32_t SD, int32_t n_agents) {
    if (SD <= 0 || SD > K1 || (SD & 1) || n_agents != A) return MM_ERR_UNSUPPORTED;
    return make_layout().total;
}

// offsets[8]: wcat, w1b, w2b, bcat, b1b, b2b, wb2, bb2
extern "C" int mm_qmixer_tc_param_offsets(int32_t SD, int32_t n_agents, int32_t* offsets_host) {
    if (!offsets_host || mm_qmixer_tc_param_count(SD, n_agents) < 0) return MM_ERR_UNSUPPORTED;
    const Layout L = make_layout();
    const int o[8] = {L.wcat, L.w1b, L.w2b, L.bcat, L.b1b, L.b2b, L.wb2, L.bb2};
    for (int i = 0; i < 8; ++i) offsets_host[i] = o[i];
    return MM_OK;
}

extern "C" int mm_qmixer_tc_forward(const float* params, const float* agent_q, const float* states, float* q_tot, int64_t M,
                                    int32_t SD, int32_t n_agents, void* stream) {
    if (!params || !agent_q || !states || !q_tot || M <= 0) return MM_ERR_BAD_ARG;
    if (mm_qmixer_tc_param_count(SD, n_agents) < 0) return MM_ERR_UNSUPPORTED;
    if (!mm::aligned(states, 8) || !mm::aligned(agent_q, 8)) return MM_ERR_BAD_ARG;
    const Layout L = make_layout();
    const size_t smem = (size_t)4 * (NCAT * K1 + AE * HH + E * HH + kRows * K1 + 2 * kRows * HH + (L.total - L.bcat) + 8 * kRows) + 128;
    cudaError_t e = cudaFuncSetAttribute(qmixer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (M + kRows - 1) / kRows;
    qmixer_tc_kernel<<<(unsigned)(tiles < sms ? tiles : sms), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        params, agent_q, states, q_tot, M, SD);
    return mm::check_launch();
}
