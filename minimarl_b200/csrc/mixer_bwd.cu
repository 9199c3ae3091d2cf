// Backward of the QMIX hypernetwork mixer (the live mixer of QMix.train_policy_on_batch) as ONE kernel on tcgen05, with every
// weight gradient contracted inside it.
//
// Reference: QMixer.forward (offpolicy/algorithms/qmix/algorithm/q_mixer.py:68-94, hypernet_layers == 2, HH = 64, E = 32, A = 2)
// differentiated by loss.backward() in offpolicy/algorithms/qmix/qmix.py:196-208.  Per (t, episode) row with state s, agent
// values q and upstream gradient g = dL/dQ_tot:
//   h1a = W1a s + c1a, a1 = relu(h1a), w1raw = W1b a1 + c1b,  w1 = |w1raw|        (hyper_w1)
//   h2a = W2a s + c2a, a2 = relu(h2a), w2raw = W2b a2 + c2b,  w2 = |w2raw|        (hyper_w2)
//   hba = Wba s + cba, ab = relu(hba), b2 = wb2 . ab + cb2                          (hyper_b2)        b1 = Wb1 s + cb1  (hyper_b1)
//   pre_e = q0 w1[0][e] + q1 w1[1][e] + b1[e],  h_e = elu(pre_e),  Q_tot = sum_e h_e w2[e] + b2
// The forward pass is mm_qmixer_tc_forward (mixer_tc.cu) and keeps nothing: this kernel RECOMPUTES the activations of a 128-row tile
// (two MMA phases) and then runs the chain backwards (two more), so the only per-row traffic is s, q, g in and dq out -- PyTorch's
// autograd path for the same thing is 7 Linear layers through cuBLAS (forward, dX, dW as separate GEMMs with materialised
// activations), 7 bias column sums and ~30 element-wise launches.
//
// 512 threads = (row, column part) as in mixer_tc.cu; fp16 operand tiles in the no-swizzle core-matrix layout ([8-feature chunk]
// [128 rows][16 B]), which serve K-major when features are contracted and MN-major when the tile's 128 ROWS are (weight gradients),
// see agent_mlp.cu.  TMEM (all 512 columns):
//   work  [0,224)    phase 1: h1a | h2a | hba | b1;  phase 2: w1raw [0,64) w2raw [64,96);  phase 3: da1 [0,64) da2 [64,128)
//   accC0 [224,320)  += [dh1a | dh2a]^T [s | 1]     lanes = first-layer outputs, columns = state features (94 = bias)
//   accC1 [320,416)  += [dhba | db1 | -]^T [s | 1]
//   acc2  [416,512)  += [a1 | a2]^T [dw1raw | dw2raw]: lanes 0..63 x cols 0..63 = dW1b^T, lanes 64..127 x cols 64..95 = dW2b^T
// The second-layer bias gradients, d wb2 and d cb2 are summed in registers over all tiles of the CTA (41 per thread).
// Gradients are scaled by a power of two (scale[0] / 8, from mm_ppo_trunk2_scale on g) so that fp16 tiles neither underflow nor
// overflow; mm_qmixer_bwd_unfold sums the per-CTA accumulators, removes the scale and writes the 14 gradient tensors.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kThreads = 512, kRows = 128;
constexpr int HH = 64, E = 32, AE = 64, NCAT = 224, K1 = 96, SDMAX = 94, ONE_COL = 94;
constexpr int CH = kRows * 16;             // bytes of one 8-feature chunk of a 128-row fp16 tile
constexpr int CHS = CH + 16;               // chunk stride of the state tile (written 4 bytes at a time along the features)
constexpr int ACC_COLS = 288, VEC = 192;
constexpr int T_WORK = 0, T_C0 = 224, T_C1 = 320, T_2 = 416;
constexpr float kScaleMul = 0.125f;

// parameter tensors in module order (QMixer.parameters()): hyper_w1.0, hyper_w1.2, hyper_w2.0, hyper_w2.2, hyper_b1, hyper_b2.0, hyper_b2.2
enum { P_W1A, P_C1A, P_W1B, P_C1B, P_W2A, P_C2A, P_W2B, P_C2B, P_WB1, P_CB1, P_WBA, P_CBA, P_WB2, P_CB2, P_N };
struct ParamPtrs { const float* p[P_N]; };

// shared memory (bytes)
constexpr int S_WCAT = 0, S_W1B = S_WCAT + 12 * NCAT * 16, S_W2B = S_W1B + 8 * AE * 16, S_VEC = S_W2B + 8 * E * 16,
              S_S = S_VEC + 384 * 4, S_STAGE = S_S + 24832, S_A12 = S_STAGE + kRows * SDMAX * 4, S_TB = S_A12 + 16 * CH,
              S_G = S_TB + 16 * CH, S_DQ = S_G + 12 * CH, S_RED = S_DQ + 4 * kRows * 8, S_BYTES = S_RED + VEC * 4;
static_assert(12 * CHS <= 24832 && S_BYTES <= 227 * 1024, "shared memory budget");
static_assert(S_S % 128 == 0 && S_STAGE % 128 == 0 && S_A12 % 128 == 0, "tile alignment");

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void store_chunk(uint8_t* tile, int chunk, int row, const float* v) {
    *reinterpret_cast<uint4*>(tile + chunk * CH + row * 16) =
        make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
// fp32 [N][K] row-major (row stride ld) -> fp16 chunk layout [K/8][N_total rows][8], rows n0 .. n0 + N
__device__ __forceinline__ void stage_weight(uint8_t* dst, const float* src, int N, int K, int ld, int n0, int n_total, int tid) {
    for (int i = tid; i < N * K; i += kThreads) {
        const int n = i / K, k = i - n * K;
        *reinterpret_cast<__half*>(dst + (k >> 3) * (n_total * 16) + (n0 + n) * 16 + (k & 7) * 2) = __float2half_rn(src[n * ld + k]);
    }
}
__device__ __forceinline__ uint64_t desc_k(uint32_t addr, int kk, uint32_t chunk_bytes) {
    return tc::make_desc(addr + (uint32_t)kk * 2u * chunk_bytes, chunk_bytes, 128u);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, int kk, uint32_t chunk_bytes) {
    return tc::make_desc(addr + (uint32_t)kk * 256u, 128u, chunk_bytes);
}
__device__ __forceinline__ void block_publish() {
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
}
__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

struct Args {
    ParamPtrs q;
    const float* states;    // [M][SD]
    const float* agent_q;   // [M][2]
    const float* dout;      // [M]
    const float* scale;     // device float (mm_ppo_trunk2_scale)
    float* dq;              // [M][2]
    float* acc;             // [grid][128][ACC_COLS]
    float* vec;             // [grid][VEC]
    long long M, n_tiles;
    int SD;
};

__global__ void __launch_bounds__(kThreads, 1) qmixer_bwd_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar_s;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane, part = warp >> 2;
    const int SD = a.SD, pairs = SD >> 1;
    uint8_t* sS = smem + S_S;
    const float* sStage = reinterpret_cast<const float*>(smem + S_STAGE);
    uint8_t* sA1 = smem + S_A12;
    uint8_t* sA2 = sA1 + 8 * CH;
    uint8_t* sDhb = smem + S_TB;               // dhba 8 chunks | db1 4 chunks | 4 spare (zero)
    uint8_t* sDb1 = sDhb + 8 * CH;
    uint8_t* sG = smem + S_G;                  // dw1raw 8 chunks | dw2raw 4 chunks
    float* sV = reinterpret_cast<float*>(smem + S_VEC);          // bcat 224 | c1b 64 | c2b 32 | wb2 64
    const float *vBcat = sV, *vC1b = sV + NCAT, *vC2b = vC1b + AE, *vWb2 = vC2b + E;
    float2* sDq = reinterpret_cast<float2*>(smem + S_DQ);
    float* sRed = reinterpret_cast<float*>(smem + S_RED);

    const uint32_t tile_bytes = (uint32_t)kRows * SD * 4;
    const long long first = blockIdx.x;
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init(&bar_s, 1);
    }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    __syncthreads();
    if (tid == 0 && first < a.n_tiles && (first + 1) * kRows <= a.M) {     // the first state tile travels while the weights are staged
        tc::mbar_expect_tx(&bar_s, tile_bytes);
        tc::bulk_g2s(smem + S_STAGE, a.states + first * kRows * SD, tile_bytes, &bar_s);
    }
    stage_weight(smem + S_WCAT, a.q.p[P_W1A], HH, SD, SD, 0, NCAT, tid);
    stage_weight(smem + S_WCAT, a.q.p[P_W2A], HH, SD, SD, HH, NCAT, tid);
    stage_weight(smem + S_WCAT, a.q.p[P_WBA], HH, SD, SD, 2 * HH, NCAT, tid);
    stage_weight(smem + S_WCAT, a.q.p[P_WB1], E, SD, SD, 3 * HH, NCAT, tid);
    for (int i = tid; i < NCAT * (K1 - SD); i += kThreads) {               // K padding (and the ones column) multiplies zero weights
        const int n = i / (K1 - SD), k = SD + i - n * (K1 - SD);
        *reinterpret_cast<__half*>(smem + S_WCAT + (k >> 3) * (NCAT * 16) + n * 16 + (k & 7) * 2) = __float2half_rn(0.f);
    }
    stage_weight(smem + S_W1B, a.q.p[P_W1B], AE, HH, HH, 0, AE, tid);
    stage_weight(smem + S_W2B, a.q.p[P_W2B], E, HH, HH, 0, E, tid);
    for (int i = tid; i < 384; i += kThreads) {
        float v;
        if (i < HH) v = a.q.p[P_C1A][i];
        else if (i < 2 * HH) v = a.q.p[P_C2A][i - HH];
        else if (i < 3 * HH) v = a.q.p[P_CBA][i - 2 * HH];
        else if (i < NCAT) v = a.q.p[P_CB1][i - 3 * HH];
        else if (i < NCAT + AE) v = a.q.p[P_C1B][i - NCAT];
        else if (i < NCAT + AE + E) v = a.q.p[P_C2B][i - NCAT - AE];
        else v = a.q.p[P_WB2][i - NCAT - AE - E];
        sV[i] = v;
    }
    for (int i = tid; i < 24832 / 16; i += kThreads) reinterpret_cast<uint4*>(sS)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 4 * CH / 16; i += kThreads) reinterpret_cast<uint4*>(sDhb + 12 * CH)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < VEC; i += kThreads) sRed[i] = 0.f;
    block_publish();

    const uint32_t tmem = tmem_slot;
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t wcat = tc::smem_u32(smem + S_WCAT), w1b = tc::smem_u32(smem + S_W1B), w2b = tc::smem_u32(smem + S_W2B);
    const uint32_t s_ = tc::smem_u32(sS), a1 = tc::smem_u32(sA1), a2 = tc::smem_u32(sA2), dhb = tc::smem_u32(sDhb), g_ = tc::smem_u32(sG);
    const uint32_t id_cat = tc::make_idesc_f16(128, NCAT), id_w1 = tc::make_idesc_f16(128, AE), id_w2 = tc::make_idesc_f16(128, E);
    const uint32_t id_da = tc::make_idesc_f16(128, HH, false, true);           // A K-major, B = W read MN-major (W^T)
    const uint32_t id_acc = tc::make_idesc_f16(128, 96, true, true);           // rows contracted
    const float gscale = a.scale[0] * kScaleMul, inv_gscale = 1.0f / gscale;
    uint32_t par = 0, par_s = 0;
    bool started = false;
    auto wait_mma = [&]() {
        tc::mbar_wait(&bar, par);
        par ^= 1;
        tc::fence_after_sync();
    };
    float acc_c1b[16], acc_c2b[8], acc_wb2[16], acc_cb2 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc_c1b[j] = acc_wb2[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_c2b[j] = 0.f;

    for (long long tile = first; tile < a.n_tiles; tile += gridDim.x) {
        const long long row0 = tile * kRows;
        const int n_valid = (int)min((long long)kRows, a.M - row0);
        const bool valid = row < n_valid;
        float2 q = make_float2(0.f, 0.f);
        float g = 0.f;
        if (valid) {
            q = *reinterpret_cast<const float2*>(a.agent_q + (row0 + row) * 2);
            g = a.dout[row0 + row] * gscale;
        }
        // ---- state tile: fp32 rows -> fp16 operand tile [s | 1 | 0] ------------------------------------------------------------
        if (n_valid == kRows) {
            tc::mbar_wait(&bar_s, par_s);
            par_s ^= 1;
            for (int idx = tid; idx < kRows * pairs; idx += kThreads) {
                const int r = idx / pairs, p = idx - r * pairs;
                const float2 v = reinterpret_cast<const float2*>(sStage)[idx];
                *reinterpret_cast<uint32_t*>(sS + (p >> 2) * CHS + r * 16 + (p & 3) * 4) = pack2(v.x, v.y);
            }
        } else {
            const float2* src = reinterpret_cast<const float2*>(a.states + row0 * SD);
            for (int idx = tid; idx < kRows * pairs; idx += kThreads) {
                const int r = idx / pairs, p = idx - r * pairs;
                const float2 v = r < n_valid ? src[idx] : make_float2(0.f, 0.f);
                *reinterpret_cast<uint32_t*>(sS + (p >> 2) * CHS + r * 16 + (p & 3) * 4) = pack2(v.x, v.y);
            }
        }
        if (part == 0)
            *reinterpret_cast<uint32_t*>(sS + (ONE_COL >> 3) * CHS + row * 16 + (ONE_COL & 7) * 2) = valid ? 0x00003c00u : 0u;
        block_publish();
        if (tid == 0) {
            const long long next = tile + gridDim.x;
            if (next < a.n_tiles && (next + 1) * kRows <= a.M) {                // the staging buffer is free: fetch the next tile
                tc::mbar_expect_tx(&bar_s, tile_bytes);
                tc::bulk_g2s(smem + S_STAGE, a.states + next * kRows * SD, tile_bytes, &bar_s);
            }
#pragma unroll
            for (int kk = 0; kk < K1 / 16; ++kk) tc::mma_f16(tmem + T_WORK, desc_k(s_, kk, CHS), desc_k(wcat, kk, NCAT * 16), id_cat, kk > 0);
            tc::mma_commit(&bar);
        }
        wait_mma();
        // ---- phase 1 epilogue: ReLU of the three first layers, b1 kept in registers -----------------------------------------------
        float b1v[8];
        uint32_t mask = 0;
        {
            const int c0 = 16 * part;
            float h1[16], h2[16], hb[16];
            tc::tmem_ld16_issue(tl + T_WORK + c0, h1);
            tc::tmem_ld16_issue(tl + T_WORK + HH + c0, h2);
            tc::tmem_ld16_issue(tl + T_WORK + 2 * HH + c0, hb);
            tc::tmem_ld8_issue(tl + T_WORK + 3 * HH + 8 * part, b1v);
            tc::tmem_ld_fence<16>(h1, h2, hb);
            tc::tmem_ld_fence<8>(b1v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float x1 = h1[j] + vBcat[c0 + j], x2 = h2[j] + vBcat[HH + c0 + j], xb = hb[j] + vBcat[2 * HH + c0 + j];
                mask |= (uint32_t)(x1 > 0.f) << j | (uint32_t)(x2 > 0.f) << (16 + j);
                h1[j] = fmaxf(x1, 0.f);
                h2[j] = fmaxf(x2, 0.f);
                acc_wb2[j] = fmaf(g, fmaxf(xb, 0.f), acc_wb2[j]);                 // d wb2 += g ab
                hb[j] = xb > 0.f ? g * vWb2[c0 + j] : 0.f;                        // d hba
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) b1v[j] += vBcat[3 * HH + 8 * part + j];
            store_chunk(sA1, 2 * part, row, h1);
            store_chunk(sA1, 2 * part + 1, row, h1 + 8);
            store_chunk(sA2, 2 * part, row, h2);
            store_chunk(sA2, 2 * part + 1, row, h2 + 8);
            store_chunk(sDhb, 2 * part, row, hb);
            store_chunk(sDhb, 2 * part + 1, row, hb + 8);
            if (part == 0) acc_cb2 += g;
        }
        block_publish();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc::mma_f16(tmem + T_WORK, desc_k(a1, kk, CH), desc_k(w1b, kk, AE * 16), id_w1, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc::mma_f16(tmem + T_WORK + AE, desc_k(a2, kk, CH), desc_k(w2b, kk, E * 16), id_w2, kk > 0);
            tc::mma_commit(&bar);
        }
        wait_mma();
        // ---- phase 2 epilogue: the mixing step forward and backward for hidden units e0 .. e0 + 7 ------------------------------------
        {
            const int e0 = 8 * part;
            float w10[8], w11[8], w2[8];
            tc::tmem_ld8_issue(tl + T_WORK + e0, w10);
            tc::tmem_ld8_issue(tl + T_WORK + E + e0, w11);
            tc::tmem_ld8_issue(tl + T_WORK + AE + e0, w2);
            tc::tmem_ld_fence<8>(w10, w11, w2);
            float dq0 = 0.f, dq1 = 0.f, db1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float r10 = w10[j] + vC1b[e0 + j], r11 = w11[j] + vC1b[E + e0 + j], r2 = w2[j] + vC2b[e0 + j];
                const float pre = fmaf(q.x, fabsf(r10), fmaf(q.y, fabsf(r11), b1v[j]));
                const float h = pre > 0.f ? pre : expm1f(pre);
                const float dpre = g * fabsf(r2) * (pre > 0.f ? 1.0f : h + 1.0f);      // elu'(x) = elu(x) + 1 for x <= 0
                w2[j] = sgn(r2) * g * h;                                               // d w2raw
                w10[j] = sgn(r10) * dpre * q.x;                                        // d w1raw
                w11[j] = sgn(r11) * dpre * q.y;
                db1[j] = dpre;
                dq0 = fmaf(dpre, fabsf(r10), dq0);
                dq1 = fmaf(dpre, fabsf(r11), dq1);
                acc_c1b[j] += w10[j];
                acc_c1b[8 + j] += w11[j];
                acc_c2b[j] += w2[j];
            }
            store_chunk(sG, part, row, w10);
            store_chunk(sG, 4 + part, row, w11);
            store_chunk(sG, 8 + part, row, w2);
            store_chunk(sDb1, part, row, db1);
            sDq[part * kRows + row] = make_float2(dq0, dq1);
        }
        block_publish();
        if (tid == 0) {
            // da1 = dw1raw W1b, da2 = dw2raw W2b (weights read MN-major);  acc2 += [a1 | a2]^T [dw1raw | dw2raw]
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc::mma_f16(tmem + T_WORK, desc_k(g_, kk, CH), desc_mn(w1b, kk, AE * 16), id_da, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) tc::mma_f16(tmem + T_WORK + HH, desc_k(g_ + 8 * CH, kk, CH), desc_mn(w2b, kk, E * 16), id_da, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) tc::mma_f16(tmem + T_2, desc_mn(a1, kk, CH), desc_mn(g_, kk, CH), id_acc, started || kk > 0);
            tc::mma_commit(&bar);
        }
        if (part == 0 && valid) {
            float2 s = sDq[row];
#pragma unroll
            for (int p = 1; p < 4; ++p) {
                const float2 t = sDq[p * kRows + row];
                s.x += t.x;
                s.y += t.y;
            }
            *reinterpret_cast<float2*>(a.dq + (row0 + row) * 2) = make_float2(s.x * inv_gscale, s.y * inv_gscale);
        }
        wait_mma();
        // ---- phase 3 epilogue: ReLU backward of hyper_w1 / hyper_w2's first layers, in place over a1 / a2 ----------------------------
        {
            const int c0 = 16 * part;
            float d1[16], d2[16];
            tc::tmem_ld16_issue(tl + T_WORK + c0, d1);
            tc::tmem_ld16_issue(tl + T_WORK + HH + c0, d2);
            tc::tmem_ld_fence<16>(d1, d2);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                d1[j] = (mask >> j) & 1u ? d1[j] : 0.f;
                d2[j] = (mask >> (16 + j)) & 1u ? d2[j] : 0.f;
            }
            store_chunk(sA1, 2 * part, row, d1);
            store_chunk(sA1, 2 * part + 1, row, d1 + 8);
            store_chunk(sA2, 2 * part, row, d2);
            store_chunk(sA2, 2 * part + 1, row, d2 + 8);
        }
        block_publish();
        if (tid == 0) {
            // first-layer weight gradients: [dh1a | dh2a]^T [s | 1] and [dhba | db1 | 0]^T [s | 1]
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) tc::mma_f16(tmem + T_C0, desc_mn(a1, kk, CH), desc_mn(s_, kk, CHS), id_acc, started || kk > 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) tc::mma_f16(tmem + T_C1, desc_mn(dhb, kk, CH), desc_mn(s_, kk, CHS), id_acc, started || kk > 0);
            tc::mma_commit(&bar);
        }
        started = true;
        wait_mma();                                  // nothing of this tile is read by the tensor core any more
    }
    // ---- accumulators out: TMEM [128 lanes][288] and the register sums --------------------------------------------------------------
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    {
        float* dst = a.acc + ((size_t)blockIdx.x * kRows + row) * ACC_COLS + part * (ACC_COLS / 4);
#pragma unroll 1
        for (int c = 0; c < ACC_COLS / 4; c += 8) {
            float v[8];
            tc::tmem_ld8(tl + T_C0 + part * (ACC_COLS / 4) + c, v);
            if (!started) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.f;
            }
            *reinterpret_cast<float4*>(dst + c) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    auto warp_add = [&](float v, int idx) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(MM_FULL_MASK, v, o);
        if (lane == 0) atomicAdd(sRed + idx, v);
    };
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        warp_add(acc_c1b[j], 8 * part + j);
        warp_add(acc_c1b[8 + j], E + 8 * part + j);
        warp_add(acc_c2b[j], AE + 8 * part + j);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) warp_add(acc_wb2[j], AE + E + 16 * part + j);
    if (part == 0) warp_add(acc_cb2, AE + E + HH);
    tc::fence_before_sync();
    __syncthreads();
    if (tid < VEC) a.vec[(size_t)blockIdx.x * VEC + tid] = sRed[tid];
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// gradient buffer: the 14 tensors back to back in module order
struct GradLayout { int off[P_N + 1]; };
__host__ __device__ inline GradLayout grad_layout(int SD) {
    const int n[P_N] = {HH * SD, HH, AE * HH, AE, HH * SD, HH, E * HH, E, E * SD, E, HH * SD, HH, HH, 1};
    GradLayout L;
    L.off[0] = 0;
    for (int i = 0; i < P_N; ++i) L.off[i + 1] = L.off[i] + n[i];
    return L;
}

// one thread per gradient element: sum of its accumulator slot over the CTAs, scale removed
struct GradPtrs { float* p[P_N]; };

__global__ void __launch_bounds__(256) qmixer_bwd_unfold_kernel(const float* __restrict__ acc, const float* __restrict__ vec, int n_ctas,
                                                                 const float* __restrict__ scale, const GradPtrs gp, int accumulate, int SD) {
    const GradLayout L = grad_layout(SD);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= L.off[P_N]) return;
    int t = 0;
    while (i >= L.off[t + 1]) ++t;
    const int j = i - L.off[t];
    int lane = -1, col = 0, vi = 0;                  // accumulator slot (lane, col) or vector slot vi
    switch (t) {
        case P_W1A: lane = j / SD;       col = j % SD; break;
        case P_C1A: lane = j;            col = ONE_COL; break;
        case P_W2A: lane = HH + j / SD;  col = j % SD; break;
        case P_C2A: lane = HH + j;       col = ONE_COL; break;
        case P_WBA: lane = j / SD;       col = 96 + j % SD; break;
        case P_CBA: lane = j;            col = 96 + ONE_COL; break;
        case P_WB1: lane = HH + j / SD;  col = 96 + j % SD; break;
        case P_CB1: lane = HH + j;       col = 96 + ONE_COL; break;
        case P_W1B: lane = j % HH;       col = 192 + j / HH; break;          // dW1b[out][in] = acc2[in][out]
        case P_W2B: lane = HH + j % HH;  col = 192 + AE + j / HH; break;
        case P_C1B: vi = j; break;
        case P_C2B: vi = AE + j; break;
        case P_WB2: vi = AE + E + j; break;
        default:    vi = AE + E + HH; break;
    }
    // 16 independent loads in flight (the sum over the CTAs is a chain of L2 round trips otherwise)
    const float* p = lane >= 0 ? acc + (size_t)lane * ACC_COLS + col : vec + vi;
    const size_t stride = lane >= 0 ? (size_t)kRows * ACC_COLS : (size_t)VEC;
    float s[16];
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
    const float v = s[0] / (scale[0] * kScaleMul);
    gp.p[t][j] = accumulate ? gp.p[t][j] + v : v;
}

int n_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

bool take_ptrs(const void* const* params14, ParamPtrs& q) {
    if (!params14) return false;
    for (int i = 0; i < P_N; ++i) {
        if (!params14[i]) return false;
        q.p[i] = static_cast<const float*>(params14[i]);
    }
    return true;
}

}  // namespace

// dims[4]: accumulator columns, vector floats per CTA, tile rows, gradient floats (all 14 tensors)
extern "C" int mm_qmixer_bwd_dims(int32_t SD, int32_t* d) {
    if (!d || SD < 2 || SD > SDMAX || (SD & 1)) return MM_ERR_UNSUPPORTED;
    d[0] = ACC_COLS; d[1] = VEC; d[2] = kRows; d[3] = grad_layout(SD).off[P_N];
    return MM_OK;
}

extern "C" int mm_qmixer_bwd_ctas(int64_t M) {
    const long long tiles = (M + kRows - 1) / kRows;
    return (int)(tiles < n_sms() ? tiles : n_sms());
}

extern "C" int mm_qmixer_bwd_grad_layout(int32_t SD, int32_t* offsets15) {
    if (!offsets15 || SD < 2 || SD > SDMAX || (SD & 1)) return MM_ERR_UNSUPPORTED;
    const GradLayout L = grad_layout(SD);
    for (int i = 0; i <= P_N; ++i) offsets15[i] = L.off[i];
    return MM_OK;
}

extern "C" int mm_qmixer_bwd(const void* const* params14, const float* agent_q, const float* states, const float* dout,
                             const float* scale, float* dq, float* acc, float* vec, int32_t n_acc_ctas, int64_t M, int32_t SD,
                             int32_t n_agents, void* stream) {
    Args a{};
    if (!take_ptrs(params14, a.q) || !agent_q || !states || !dout || !scale || !dq || !acc || !vec || M <= 0) return MM_ERR_BAD_ARG;
    if (SD < 2 || SD > SDMAX || (SD & 1) || n_agents != 2) return MM_ERR_UNSUPPORTED;
    if (!mm::aligned(states, 16) || !mm::aligned(agent_q, 8) || !mm::aligned(dq, 8) || !mm::aligned(acc, 16)) return MM_ERR_BAD_ARG;
    const int grid = mm_qmixer_bwd_ctas(M);
    if (n_acc_ctas < grid) return MM_ERR_BAD_ARG;
    a.states = states; a.agent_q = agent_q; a.dout = dout; a.scale = scale; a.dq = dq; a.acc = acc; a.vec = vec;
    a.M = M; a.n_tiles = (M + kRows - 1) / kRows; a.SD = SD;
    cudaError_t e = cudaFuncSetAttribute(qmixer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_BYTES);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    qmixer_bwd_kernel<<<grid, kThreads, S_BYTES, static_cast<cudaStream_t>(stream)>>>(a);
    return mm::check_launch();
}

extern "C" int mm_qmixer_bwd_unfold(const float* acc, const float* vec, int32_t n_ctas, const float* scale, float* grads, int32_t SD,
                                    void* stream) {
    if (!acc || !vec || !scale || !grads || n_ctas < 1) return MM_ERR_BAD_ARG;
    if (SD < 2 || SD > SDMAX || (SD & 1)) return MM_ERR_UNSUPPORTED;
    const GradLayout L = grad_layout(SD);
    GradPtrs gp;
    for (int i = 0; i < P_N; ++i) gp.p[i] = grads + L.off[i];
    qmixer_bwd_unfold_kernel<<<(L.off[P_N] + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(acc, vec, n_ctas, scale, gp, 0, SD);
    return mm::check_launch();
}

// the same, every gradient ADDED to its own destination tensor (grads14: HOST array of 14 device pointers, module order)
extern "C" int mm_qmixer_bwd_unfold_into(const float* acc, const float* vec, int32_t n_ctas, const float* scale, void* const* grads14,
                                         int32_t SD, void* stream) {
    if (!acc || !vec || !scale || !grads14 || n_ctas < 1) return MM_ERR_BAD_ARG;
    if (SD < 2 || SD > SDMAX || (SD & 1)) return MM_ERR_UNSUPPORTED;
    GradPtrs gp;
    for (int i = 0; i < P_N; ++i) {
        if (!grads14[i]) return MM_ERR_BAD_ARG;
        gp.p[i] = static_cast<float*>(grads14[i]);
    }
    qmixer_bwd_unfold_kernel<<<(grad_layout(SD).off[P_N] + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(acc, vec, n_ctas, scale,
                                                                                                                    gp, 1, SD);
    return mm::check_launch();
}
