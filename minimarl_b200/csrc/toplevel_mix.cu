// Fused kernels for the recurrent hypernetwork mixer of the top-level qmix/ package.
//
//   Mix_Net.forward (qmix/_network.py:199-217), per batch row and chunk step:
//       mh   = GRUCell(S, 32)(concat_a obs_a, mh)                       S = A * D, A = 2
//       w1   = |Linear(32, 2*32)(mh)| viewed [32][2];  b1 = Linear(32, 32)(mh)
//       w2   = |Linear(32, 32)(mh)|;                   b2 = Linear(32, 1)(ReLU(Linear(32, 32)(mh)))
//       Qtot = sum_h w2[h] * ReLU(w1[h][0] q_0 + w1[h][1] q_1 + b1[h]) + b2
//   as used by Train_dqn.train (qmix/_train.py:19-121): q_a is the behaviour Q of the taken action (live side, with
//   gradient) or max_o Q_tgt (target side); the mixer state is reset to zero after a terminal step.
//
// * mixnet_seq_kernel: forward over the L steps of the sampled chunks, 32 rows per CTA, 128 threads -- thread (warp g,
//   lane j) owns hidden unit j of rows 8g..8g+7, the sums over units are warp reductions.  SAVE keeps r, z, n, gh_n and
//   mh per (step, row) for the backward.
// * mixnet_bwd_kernel: BPTT; the hypernetwork layers are recomputed from the saved mh, every weight gradient is
//   accumulated in registers across rows and steps (one partial per CTA, reduced in tile order), and the gradient of the
//   taken-action Q values is written straight into the dq tensor the Q_Net backward reads.
// * qmix_chunk_loss_kernel: targets, squared errors, dQtot and the last-step TD error.
// fp32 FFMA throughout, like toplevel_qnet.cu.
#include "common.cuh"

namespace {

constexpr int HH = 32, RT = 32, BT = 128, HS = 36, NAMAX = 8;
constexpr int NHY = 160;       // hypernetwork outputs per row: u1 (a = 0) | u1 (a = 1) | b1 | u2 | b2 hidden
constexpr int MSV = 160;       // saved floats per (step, row): r | z | n | gh_n | mh
constexpr int kMaxS = 96;

struct MixP {                  // reference layout
    const float *wih, *whh, *bih, *bhh;          // gru: [96][S], [96][32], [96], [96]
    const float *w1, *c1;                        // hyper_net_weight_1: [64][32], [64]
    const float *w2, *c2;                        // hyper_net_weight_2: [32][32], [32]
    const float *wb1, *cb1;                      // hyper_net_bias_1:   [32][32], [32]
    const float *wb2a, *cb2a, *wb2, *cb2;        // hyper_net_bias_2:   [32][32], [32], [1][32], [1]
};
constexpr int kMixTensors = 14;

__host__ __device__ inline int pad8(int d) { return (d + 7) & ~7; }
__host__ __device__ inline int mix_param_count(int S) { return 96 * S + 96 * 32 + 96 + 96 + 64 * 32 + 64 + 3 * (32 * 32 + 32) + 32 + 1; }

__device__ __forceinline__ float sigmoid_(float x) { return __frcp_rn(1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { return 1.0f - 2.0f * __frcp_rn(1.0f + __expf(2.0f * x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MM_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ void cp_async4(void* smem_ptr, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_ptr)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// hypernetwork weights as [k][which][j] (forward: lane = unit reads consecutive words); which: 0 u1 a=0, 1 u1 a=1, 2 b1, 3 u2, 4 b2a
__device__ void stage_hyper(float* Wh, float* bh, float* wb2, const MixP& w, int bt) {
    for (int i = bt; i < HH * NHY; i += BT) {
        const int k = i / NHY, rem = i - k * NHY, which = rem >> 5, j = rem & 31;
        float v;
        if (which == 0) v = w.w1[(2 * j) * HH + k];
        else if (which == 1) v = w.w1[(2 * j + 1) * HH + k];
        else if (which == 2) v = w.wb1[j * HH + k];
        else if (which == 3) v = w.w2[j * HH + k];
        else v = w.wb2a[j * HH + k];
        Wh[i] = v;
    }
    for (int i = bt; i < NHY; i += BT) {
        const int which = i >> 5, j = i & 31;
        bh[i] = which == 0 ? w.c1[2 * j] : which == 1 ? w.c1[2 * j + 1] : which == 2 ? w.cb1[j] : which == 3 ? w.c2[j] : w.cb2a[j];
    }
    for (int i = bt; i < HH; i += BT) wb2[i] = w.wb2[i];
}

// hypernetwork pre-activations of unit j for rows r0..r0+7 from the mixer state tile
__device__ __forceinline__ void hyper_forward(const float* Wh, const float* bh, const float* mh, int r0, int j, float (&u1e)[8],
                                              float (&u1o)[8], float (&ub)[8], float (&u2)[8], float (&ua)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { u1e[i] = bh[j]; u1o[i] = bh[32 + j]; ub[i] = bh[64 + j]; u2[i] = bh[96 + j]; ua[i] = bh[128 + j]; }
#pragma unroll 2
    for (int k = 0; k < HH; k += 4) {
        float w0[4], w1[4], w2[4], w3[4], w4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float* wk = Wh + (k + u) * NHY + j;
            w0[u] = wk[0]; w1[u] = wk[32]; w2[u] = wk[64]; w3[u] = wk[96]; w4[u] = wk[128];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 h = *reinterpret_cast<const float4*>(mh + (r0 + i) * HS + k);
            const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                u1e[i] = fmaf(hv[u], w0[u], u1e[i]); u1o[i] = fmaf(hv[u], w1[u], u1o[i]); ub[i] = fmaf(hv[u], w2[u], ub[i]);
                u2[i] = fmaf(hv[u], w3[u], u2[i]); ua[i] = fmaf(hv[u], w4[u], ua[i]);
            }
        }
    }
}

struct MixSeqArgs {
    MixP w;
    const float* obs;            // row (idx(b), l): S contiguous floats at (idx(b) * L + l) * S
    const long long* row_index;  // [B] or null
    const float* dones;          // [rows][L] through row_index (state reset after a terminal step)
    const float* q_all;          // [B][L][2][NA]
    const float* actions;        // [rows][L][2] through row_index (float); null = take max_o (target side)
    float* qtot;                 // [B][L]
    float* save;                 // [L][B][MSV] or null
    int B, L, S, NA;
};

// q fed to the mixer for (row, agent): the taken action's value, or the maximum
__device__ __forceinline__ float pick_q(const float* q, const float* act, int NA, int* which) {
    if (act) {
        const int o = min(max((int)*act, 0), NA - 1);
        if (which) *which = o;
        return q[o];
    }
    float best = q[0];
    for (int o = 1; o < NA; ++o) best = fmaxf(best, q[o]);
    return best;
}

template <bool SAVE>
__global__ void __launch_bounds__(BT, 1) mixnet_seq_kernel(const __grid_constant__ MixSeqArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ long long rows[RT];
    __shared__ float qin[RT][2];
    const int bt = threadIdx.x, j = bt & 31, g = bt >> 5, r0 = g * 8;
    const int Sp = pad8(a.S);
    float* Wgi = smem;                        // [Sp][3][32]
    float* Wgh = Wgi + Sp * 96;               // [32][3][32]
    float* bg = Wgh + HH * 96;                // [6][32]
    float* Wh = bg + 192;                     // [32][5][32]
    float* bh = Wh + HH * NHY;                // [5][32]
    float* wb2 = bh + NHY;                    // [32]
    float* xs = wb2 + HH;                     // [2][RT][Sp]
    float* mh = xs + 2 * RT * Sp;             // [2][RT][HS]
    for (int i = bt; i < Sp * 96; i += BT) {
        const int k = i / 96, rem = i - k * 96;
        Wgi[i] = k < a.S ? a.w.wih[rem * a.S + k] : 0.f;
    }
    for (int i = bt; i < HH * 96; i += BT) {
        const int k = i / 96, rem = i - k * 96;
        Wgh[i] = a.w.whh[rem * HH + k];
    }
    for (int i = bt; i < 192; i += BT) bg[i] = i < 96 ? a.w.bih[i] : a.w.bhh[i - 96];
    stage_hyper(Wh, bh, wb2, a.w, bt);
    const float cb2 = a.w.cb2[0];
    for (int i = bt; i < 2 * RT * Sp + 2 * RT * HS; i += BT) xs[i] = 0.f;
    const int b0 = blockIdx.x * RT;
    const int n_valid = min(RT, a.B - b0);
    if (bt < RT) rows[bt] = bt < n_valid ? (a.row_index ? a.row_index[b0 + bt] : (long long)(b0 + bt)) : 0;
    __syncthreads();
    auto fetch = [&](int l, int buf) {
        float* dst = xs + buf * RT * Sp;
#pragma unroll 1
        for (int rr = 0; rr < 8; ++rr) {
            const int row = g * 8 + rr;
            if (row >= n_valid) break;
            const float* src = a.obs + (rows[row] * a.L + l) * a.S;
            for (int d = j; d < a.S; d += 32) cp_async4(dst + row * Sp + d, src + d);
        }
        cp_async_commit();
    };
    fetch(0, 0);
    int cur = 0;
    for (int l = 0; l < a.L; ++l) {
        if (bt < 2 * RT) {
            const int row = bt >> 1, ag = bt & 1;
            float v = 0.f;
            if (row < n_valid)
                v = pick_q(a.q_all + (((long long)(b0 + row) * a.L + l) * 2 + ag) * a.NA,
                           a.actions ? a.actions + (rows[row] * a.L + l) * 2 + ag : nullptr, a.NA, nullptr);
            qin[row][ag] = v;
        }
        cp_async_wait_all();
        __syncthreads();
        if (l + 1 < a.L) fetch(l + 1, (l + 1) & 1);
        const float* x = xs + (l & 1) * RT * Sp;
        const float* hc = mh + cur * RT * HS;
        float* hn = mh + (cur ^ 1) * RT * HS;
        // ---- GRUCell ---------------------------------------------------------------------------------------------
        {
            float ar[8], az[8], ai[8], ah[8];
            const float cr = bg[j] + bg[96 + j], cz = bg[32 + j] + bg[128 + j], ci = bg[64 + j], ch = bg[160 + j];
#pragma unroll
            for (int i = 0; i < 8; ++i) { ar[i] = cr; az[i] = cz; ai[i] = ci; ah[i] = ch; }
            for (int k = 0; k < Sp; k += 4) {
                float wr[4], wz[4], wn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float* wk = Wgi + (k + u) * 96 + j;
                    wr[u] = wk[0]; wz[u] = wk[32]; wn[u] = wk[64];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(x + (r0 + i) * Sp + k);
                    const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        ar[i] = fmaf(xv[u], wr[u], ar[i]); az[i] = fmaf(xv[u], wz[u], az[i]); ai[i] = fmaf(xv[u], wn[u], ai[i]);
                    }
                }
            }
#pragma unroll 2
            for (int k = 0; k < HH; k += 4) {
                float wr[4], wz[4], wn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float* wk = Wgh + (k + u) * 96 + j;
                    wr[u] = wk[0]; wz[u] = wk[32]; wn[u] = wk[64];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(hc + (r0 + i) * HS + k);
                    const float hv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        ar[i] = fmaf(hv[u], wr[u], ar[i]); az[i] = fmaf(hv[u], wz[u], az[i]); ah[i] = fmaf(hv[u], wn[u], ah[i]);
                    }
                }
            }
            float* sv = SAVE ? a.save + ((long long)l * a.B + b0) * MSV : nullptr;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float r = sigmoid_(ar[i]), z = sigmoid_(az[i]);
                const float n = tanh_(fmaf(r, ah[i], ai[i]));
                const float hp = hc[(r0 + i) * HS + j];
                const float hv = fmaf(z, hp - n, n);
                hn[(r0 + i) * HS + j] = hv;
                if (SAVE && r0 + i < n_valid) {
                    float* o = sv + (r0 + i) * MSV + j;
                    o[0] = r; o[32] = z; o[64] = n; o[96] = ah[i]; o[128] = hv;
                }
            }
        }
        __syncthreads();
        // ---- hypernetworks + mixing ----------------------------------------------------------------------------------
        {
            float u1e[8], u1o[8], ub[8], u2[8], ua[8];
            hyper_forward(Wh, bh, hn, r0, j, u1e, u1o, ub, u2, ua);
            const float wj = wb2[j];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + i;
                const float pre = fmaf(fabsf(u1e[i]), qin[row][0], fmaf(fabsf(u1o[i]), qin[row][1], ub[i]));
                float part = fmaf(fabsf(u2[i]), fmaxf(pre, 0.f), fmaxf(ua[i], 0.f) * wj);
                part = warp_sum(part);
                if (j == 0 && row < n_valid) a.qtot[(long long)(b0 + row) * a.L + l] = part + cb2;
            }
        }
        cur ^= 1;
        // state reset after a terminal step (the next step's GRU reads mh[cur] only after the barrier at the loop top)
        {
            const int row = bt >> 2;
            if (row < n_valid) {
                const float d = a.dones[rows[row] * a.L + l];
                if (d != 0.f) {
                    const float keep = 1.0f - d;
                    float* hcn = mh + cur * RT * HS;
#pragma unroll
                    for (int k = 0; k < 8; ++k) hcn[row * HS + (bt & 3) * 8 + k] *= keep;
                }
            }
        }
    }
}

struct MixBwdArgs {
    MixP w;
    const float* obs;
    const long long* row_index;
    const float* dones;
    const float* q_all;          // [B][L][2][NA]
    const float* actions;        // [rows][L][2] through row_index
    const float* save;           // [L][B][MSV]
    const float* dqtot;          // [B][L]
    float* dq;                   // [B][L][2][NA]: gradient of the taken-action values, zero elsewhere
    float* partials;             // [n_tiles][P]
    int B, L, S, NA;
};

__global__ void __launch_bounds__(BT, 1) mixnet_bwd_kernel(const __grid_constant__ MixBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ long long rows[RT];
    __shared__ float qin[RT][2];
    __shared__ int qact[RT][2];
    __shared__ float dout[RT];
    const int bt = threadIdx.x, j = bt & 31, g = bt >> 5, r0 = g * 8;
    const int Sp = pad8(a.S);
    float* Whh = smem;                        // [96][32] reference layout (dmh_prev = GH Whh)
    float* Wh = Whh + 96 * HH;                // [32][5][32] forward layout
    float* WhT = Wh + HH * NHY;               // [160][32]: row m = (which, unit), column = input k (dmh = G WhT)
    float* bh = WhT + NHY * HH;               // [5][32]
    float* wb2 = bh + NHY;                    // [32]
    float* X = wb2 + HH;                      // [RT][Sp]
    float* MH = X + RT * Sp;                  // [RT][HS]  output state of this step
    float* MP = MH + RT * HS;                 // [RT][HS]  (masked) input state of this step
    float* G = MP + RT * HS;                  // [RT][160] du1e | du1o | dub | du2 | dua
    float* GI = G + RT * NHY;                 // [RT][96]
    float* GH = GI + RT * 96;                 // [RT][96]
    for (int i = bt; i < 96 * HH; i += BT) Whh[i] = a.w.whh[i];
    stage_hyper(Wh, bh, wb2, a.w, bt);
    for (int i = bt; i < RT * (Sp + 2 * HS); i += BT) X[i] = 0.f;
    __syncthreads();
    for (int i = bt; i < NHY * HH; i += BT) {
        const int m = i >> 5, k = i & 31;
        WhT[i] = Wh[k * NHY + m];
    }
    const int b0 = blockIdx.x * RT;
    const int n_valid = min(RT, a.B - b0);
    if (bt < RT) rows[bt] = bt < n_valid ? (a.row_index ? a.row_index[b0 + bt] : (long long)(b0 + bt)) : 0;
    __syncthreads();

    // register accumulators.  dWih [96][Sp]: 6 m x (Sp / 8) k per thread (m = 6 (bt & 15) + u, k = (Sp / 8) (bt >> 4) + v);
    // dWhh [96][32]: 6 m x 4 k (k = 4 (bt >> 4) + v); hypernetworks [160][32]: 10 m x 4 k (m = 10 (bt & 15) + u, k = 4 (bt >> 4) + v)
    constexpr int KW = kMaxS / 8;
    float gWih[6][KW], gWhh[6][4], gWhy[10][4];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
#pragma unroll
        for (int v = 0; v < KW; ++v) gWih[u][v] = 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) gWhh[u][v] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < 10; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) gWhy[u][v] = 0.f;
    float gbih = 0.f, gbhh = 0.f, gbhy0 = 0.f, gbhy1 = 0.f, gwb2 = 0.f, gcb2 = 0.f;
    const int kw = Sp / 8;
    float carry[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) carry[i] = 0.f;

    for (int l = a.L - 1; l >= 0; --l) {
        const float* sv = a.save + ((long long)l * a.B + b0) * MSV;
        const float* svp = l > 0 ? a.save + ((long long)(l - 1) * a.B + b0) * MSV : nullptr;
        for (int i = bt; i < n_valid * 8; i += BT) {
            const int row = i >> 3, q4 = i & 7;
            *reinterpret_cast<float4*>(MH + row * HS + 4 * q4) = *reinterpret_cast<const float4*>(sv + row * MSV + 128 + 4 * q4);
            float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (svp) {
                hp = *reinterpret_cast<const float4*>(svp + row * MSV + 128 + 4 * q4);
                const float keep = 1.0f - a.dones[rows[row] * a.L + l - 1];
                hp.x *= keep; hp.y *= keep; hp.z *= keep; hp.w *= keep;
            }
            *reinterpret_cast<float4*>(MP + row * HS + 4 * q4) = hp;
        }
#pragma unroll 1
        for (int rr = 0; rr < 8; ++rr) {
            const int row = g * 8 + rr;
            if (row >= n_valid) break;
            const float* src = a.obs + (rows[row] * a.L + l) * a.S;
            for (int d = j; d < a.S; d += 32) X[row * Sp + d] = src[d];
        }
        if (bt < 2 * RT) {
            const int row = bt >> 1, ag = bt & 1;
            float v = 0.f;
            int which = 0;
            if (row < n_valid)
                v = pick_q(a.q_all + (((long long)(b0 + row) * a.L + l) * 2 + ag) * a.NA,
                           a.actions + (rows[row] * a.L + l) * 2 + ag, a.NA, &which);
            qin[row][ag] = v;
            qact[row][ag] = which;
        }
        if (bt < RT) dout[bt] = bt < n_valid ? a.dqtot[(long long)(b0 + bt) * a.L + l] : 0.f;
        __syncthreads();
        // ---- M1: hypernetworks recomputed, mixing backward -----------------------------------------------------------------
        {
            float u1e[8], u1o[8], ub[8], u2[8], ua[8];
            hyper_forward(Wh, bh, MH, r0, j, u1e, u1o, ub, u2, ua);
            const float wj = wb2[j];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + i;
                const float q0 = qin[row][0], q1 = qin[row][1], d = dout[row];
                const float pre = fmaf(fabsf(u1e[i]), q0, fmaf(fabsf(u1o[i]), q1, ub[i]));
                const float x = fmaxf(pre, 0.f);
                const float dpre = pre > 0.f ? d * fabsf(u2[i]) : 0.f;
                const float sgn_e = (u1e[i] > 0.f) - (u1e[i] < 0.f), sgn_o = (u1o[i] > 0.f) - (u1o[i] < 0.f);
                const float sgn_2 = (u2[i] > 0.f) - (u2[i] < 0.f);
                float* gr = G + row * NHY + j;
                gr[0] = dpre * q0 * sgn_e;
                gr[32] = dpre * q1 * sgn_o;
                gr[64] = dpre;
                gr[96] = d * x * sgn_2;
                gr[128] = ua[i] > 0.f ? d * wj : 0.f;
                gwb2 = fmaf(d, fmaxf(ua[i], 0.f), gwb2);
                const float dq0 = warp_sum(dpre * fabsf(u1e[i])), dq1 = warp_sum(dpre * fabsf(u1o[i]));
                if (row < n_valid && j < 2 * a.NA) {
                    const int ag = j / a.NA, o = j - ag * a.NA;
                    a.dq[(((long long)(b0 + row) * a.L + l) * 2 + ag) * a.NA + o] = o == qact[row][ag] ? (ag ? dq1 : dq0) : 0.f;
                }
                if (j == 0) gcb2 += d;
            }
        }
        __syncthreads();
        // ---- M2: dmh = carry + G WhT; GRU cell backward --------------------------------------------------------------------
        float dhz[8];
        {
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = carry[i];
#pragma unroll 2
            for (int m = 0; m < NHY; m += 4) {
                const float w0 = WhT[m * HH + j], w1 = WhT[(m + 1) * HH + j], w2 = WhT[(m + 2) * HH + j], w3 = WhT[(m + 3) * HH + j];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 gq = *reinterpret_cast<const float4*>(G + (r0 + i) * NHY + m);
                    acc[i] = fmaf(gq.x, w0, acc[i]); acc[i] = fmaf(gq.y, w1, acc[i]); acc[i] = fmaf(gq.z, w2, acc[i]); acc[i] = fmaf(gq.w, w3, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + i;
                float r = 0.f, z = 0.f, n = 0.f, ghn = 0.f;
                if (row < n_valid) {
                    const float* o = sv + row * MSV + j;
                    r = o[0]; z = o[32]; n = o[64]; ghn = o[96];
                }
                const float dh = acc[i], hp = MP[row * HS + j];
                const float dpn = dh * (1.f - z) * (1.f - n * n);
                const float dpz = dh * (hp - n) * z * (1.f - z);
                const float dpr = dpn * ghn * r * (1.f - r);
                dhz[i] = dh * z;
                GI[row * 96 + j] = dpr; GI[row * 96 + 32 + j] = dpz; GI[row * 96 + 64 + j] = dpn;
                GH[row * 96 + j] = dpr; GH[row * 96 + 32 + j] = dpz; GH[row * 96 + 64 + j] = dpn * r;
            }
        }
        __syncthreads();
        // ---- M3: carry = (GH Whh + dh z) * keep ---------------------------------------------------------------------------------
        {
            float ah[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ah[i] = dhz[i];
#pragma unroll 2
            for (int m = 0; m < 96; m += 4) {
                const float w0 = Whh[m * HH + j], w1 = Whh[(m + 1) * HH + j], w2 = Whh[(m + 2) * HH + j], w3 = Whh[(m + 3) * HH + j];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 gh = *reinterpret_cast<const float4*>(GH + (r0 + i) * 96 + m);
                    ah[i] = fmaf(gh.x, w0, ah[i]); ah[i] = fmaf(gh.y, w1, ah[i]); ah[i] = fmaf(gh.z, w2, ah[i]); ah[i] = fmaf(gh.w, w3, ah[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + i;
                float keep = 1.f;
                if (l > 0 && row < n_valid) keep = 1.0f - a.dones[rows[row] * a.L + l - 1];
                carry[i] = ah[i] * keep;
            }
        }
        // ---- M4: weight gradients ----------------------------------------------------------------------------------------------
        {
            const int m1 = 6 * (bt & 15), k1 = kw * (bt >> 4), k3 = 4 * (bt >> 4), m4 = 10 * (bt & 15);
#pragma unroll 2
            for (int row = 0; row < RT; ++row) {
                const float2 i0 = *reinterpret_cast<const float2*>(GI + row * 96 + m1), i1 = *reinterpret_cast<const float2*>(GI + row * 96 + m1 + 2),
                             i2 = *reinterpret_cast<const float2*>(GI + row * 96 + m1 + 4);
                const float2 h0 = *reinterpret_cast<const float2*>(GH + row * 96 + m1), h1 = *reinterpret_cast<const float2*>(GH + row * 96 + m1 + 2),
                             h2 = *reinterpret_cast<const float2*>(GH + row * 96 + m1 + 4);
                const float gi[6] = {i0.x, i0.y, i1.x, i1.y, i2.x, i2.y}, gh[6] = {h0.x, h0.y, h1.x, h1.y, h2.x, h2.y};
#pragma unroll
                for (int v = 0; v < KW; ++v) {
                    if (v < kw) {
                        const float x = X[row * Sp + k1 + v];
#pragma unroll
                        for (int u = 0; u < 6; ++u) gWih[u][v] = fmaf(gi[u], x, gWih[u][v]);
                    }
                }
                const float4 hq = *reinterpret_cast<const float4*>(MP + row * HS + k3);
                const float4 mq = *reinterpret_cast<const float4*>(MH + row * HS + k3);
                const float hv[4] = {hq.x, hq.y, hq.z, hq.w}, mv[4] = {mq.x, mq.y, mq.z, mq.w};
#pragma unroll
                for (int u = 0; u < 6; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) gWhh[u][v] = fmaf(gh[u], hv[v], gWhh[u][v]);
                float gy[10];
#pragma unroll
                for (int u = 0; u < 10; u += 2) {
                    const float2 t = *reinterpret_cast<const float2*>(G + row * NHY + m4 + u);
                    gy[u] = t.x; gy[u + 1] = t.y;
                }
#pragma unroll
                for (int u = 0; u < 10; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) gWhy[u][v] = fmaf(gy[u], mv[v], gWhy[u][v]);
                if (bt < 96) { gbih += GI[row * 96 + bt]; gbhh += GH[row * 96 + bt]; }
                gbhy0 += G[row * NHY + bt];
                if (bt < NHY - BT) gbhy1 += G[row * NHY + BT + bt];
            }
        }
        __syncthreads();
    }
    // ---- partial gradients, reference layout: wih whh bih bhh w1 c1 w2 c2 wb1 cb1 wb2a cb2a wb2 cb2 ----------------------------
    const int P = mix_param_count(a.S);
    float* out = a.partials + (long long)blockIdx.x * P;
    const int oWih = 0, oWhh = oWih + 96 * a.S, obih = oWhh + 96 * 32, obhh = obih + 96, oW1 = obhh + 96, oc1 = oW1 + 64 * 32,
              oW2 = oc1 + 64, oc2 = oW2 + 32 * 32, oWb1 = oc2 + 32, ocb1 = oWb1 + 32 * 32, oWb2a = ocb1 + 32, ocb2a = oWb2a + 32 * 32,
              owb2 = ocb2a + 32, ocb2 = owb2 + 32;
    // row m of the hypernetwork block (which = m / 32, unit = m % 32) -> (weight offset of that output row, bias offset)
    auto hy_row = [&](int m, int& wrow, int& bidx) {
        const int which = m >> 5, u = m & 31;
        if (which == 0) { wrow = oW1 + (2 * u) * HH; bidx = oc1 + 2 * u; }
        else if (which == 1) { wrow = oW1 + (2 * u + 1) * HH; bidx = oc1 + 2 * u + 1; }
        else if (which == 2) { wrow = oWb1 + u * HH; bidx = ocb1 + u; }
        else if (which == 3) { wrow = oW2 + u * HH; bidx = oc2 + u; }
        else { wrow = oWb2a + u * HH; bidx = ocb2a + u; }
    };
    {
        const int m1 = 6 * (bt & 15), k1 = kw * (bt >> 4), k3 = 4 * (bt >> 4), m4 = 10 * (bt & 15);
#pragma unroll
        for (int u = 0; u < 6; ++u) {
#pragma unroll
            for (int v = 0; v < KW; ++v)
                if (v < kw && k1 + v < a.S) out[oWih + (m1 + u) * a.S + k1 + v] = gWih[u][v];
#pragma unroll
            for (int v = 0; v < 4; ++v) out[oWhh + (m1 + u) * HH + k3 + v] = gWhh[u][v];
        }
#pragma unroll
        for (int u = 0; u < 10; ++u) {
            int wrow, bidx;
            hy_row(m4 + u, wrow, bidx);
#pragma unroll
            for (int v = 0; v < 4; ++v) out[wrow + k3 + v] = gWhy[u][v];
        }
        if (bt < 96) { out[obih + bt] = gbih; out[obhh + bt] = gbhh; }
        {
            int wrow, bidx;
            hy_row(bt, wrow, bidx);
            out[bidx] = gbhy0;
            if (bt < NHY - BT) { hy_row(BT + bt, wrow, bidx); out[bidx] = gbhy1; }
        }
        // dwb2[j]: every warp holds a partial over its 8 rows; dcb2: lane 0 of every warp
        __shared__ float red[4][33];
        red[g][j] = gwb2;
        if (j == 0) red[g][32] = gcb2;
        __syncthreads();
        if (bt < HH) out[owb2 + bt] = (red[0][bt] + red[1][bt]) + (red[2][bt] + red[3][bt]);
        if (bt == 0) out[ocb2] = (red[0][32] + red[1][32]) + (red[2][32] + red[3][32]);
    }
}

struct MixReduceArgs {
    const float* partials;
    float* grad[kMixTensors];
    int sizes[kMixTensors];
    int n_tiles, P, accumulate;
};
__global__ void mix_grad_reduce_kernel(const __grid_constant__ MixReduceArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.P) return;
    float sum = 0.f;
#pragma unroll 16                                     // same summation order, 16 loads in flight instead of 4
    for (int t = 0; t < a.n_tiles; ++t) sum += a.partials[(long long)t * a.P + i];
    int off = i;
#pragma unroll
    for (int k = 0; k < kMixTensors; ++k) {
        if (off < a.sizes[k]) {
            float* dst = a.grad[k] + off;
            *dst = a.accumulate ? *dst + sum : sum;
            return;
        }
        off -= a.sizes[k];
    }
}

// y = w_b * sum_a (r_a + gamma (1 - done) Qtot_tgt)  (qmix/_train.py:77-81; the bootstrap counted once per agent, quirk Q7)
__global__ void qmix_chunk_loss_kernel(const float* __restrict__ qtot, const float* __restrict__ qtot_tgt,
                                       const float* __restrict__ rewards, const float* __restrict__ dones,
                                       const long long* __restrict__ row_index, const float* __restrict__ is_weight,
                                       float* __restrict__ se, float* __restrict__ dqtot, float* __restrict__ td_last, int B, int L,
                                       float gamma) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * L) return;
    const int b = (int)(i / L), l = (int)(i - (long long)b * L);
    const long long srow = (row_index ? row_index[b] : (long long)b) * L + l;
    const float boot = gamma * (1.0f - dones[srow]) * qtot_tgt[i];
    const float y = is_weight[b] * ((rewards[srow * 2] + boot) + (rewards[srow * 2 + 1] + boot));
    const float diff = y - qtot[i];
    se[i] = diff * diff;
    dqtot[i] = -2.0f * diff / (float)B;
    if (l == L - 1) td_last[b] = fabsf(diff);
}

MixP to_mix(const float* const* p) {
    return MixP{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13]};
}
size_t seq_smem(int Sp) { return sizeof(float) * (size_t)(Sp * 96 + HH * 96 + 192 + HH * NHY + NHY + HH + 2 * RT * Sp + 2 * RT * HS); }
size_t bwd_smem(int Sp) {
    return sizeof(float) * (size_t)(96 * HH + 2 * HH * NHY + NHY + HH + RT * Sp + 2 * RT * HS + RT * NHY + 2 * RT * 96);
}

}  // namespace

extern "C" int mm_toplevel_mix_param_count(int32_t S) { return mix_param_count(S); }
extern "C" int mm_toplevel_mix_save_stride(void) { return MSV; }

// params: 14 device pointers, reference layout: gru.weight_ih, gru.weight_hh, gru.bias_ih, gru.bias_hh, hyper_net_weight_1.{weight,
// bias}, hyper_net_weight_2.{weight, bias}, hyper_net_bias_1.{weight, bias}, hyper_net_bias_2.0.{weight, bias}, hyper_net_bias_2.2.{weight, bias}
extern "C" int mm_toplevel_mix_seq(const float* const* params_host, const float* obs, const int64_t* row_index, const float* dones,
                                   const float* q_all, const float* actions, float* qtot, float* save, int32_t B, int32_t L,
                                   int32_t S, int32_t NA, void* stream) {
    if (!params_host || !obs || !dones || !q_all || !qtot || B <= 0 || L <= 0) return MM_ERR_BAD_ARG;
    if (S < 1 || S > kMaxS || NA < 1 || NA > NAMAX) return MM_ERR_UNSUPPORTED;
    MixSeqArgs a{};
    a.w = to_mix(params_host);
    a.obs = obs; a.row_index = reinterpret_cast<const long long*>(row_index); a.dones = dones; a.q_all = q_all; a.actions = actions;
    a.qtot = qtot; a.save = save; a.B = B; a.L = L; a.S = S; a.NA = NA;
    const size_t smem = seq_smem(pad8(S));
    const int grid = (B + RT - 1) / RT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (save) {
        e = cudaFuncSetAttribute(mixnet_seq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
        mixnet_seq_kernel<true><<<grid, BT, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(mixnet_seq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
        mixnet_seq_kernel<false><<<grid, BT, smem, st>>>(a);
    }
    return mm::check_launch();
}

extern "C" int mm_toplevel_mix_bwd(const float* const* params_host, float* const* grads_host, const float* obs,
                                   const int64_t* row_index, const float* dones, const float* q_all, const float* actions,
                                   const float* save, const float* dqtot, float* dq, float* partials, int32_t accumulate, int32_t B,
                                   int32_t L, int32_t S, int32_t NA, void* stream) {
    if (!params_host || !grads_host || !obs || !dones || !q_all || !actions || !save || !dqtot || !dq || !partials || B <= 0 || L <= 0)
        return MM_ERR_BAD_ARG;
    if (S < 1 || S > kMaxS || NA < 1 || NA > NAMAX) return MM_ERR_UNSUPPORTED;
    MixBwdArgs a{};
    a.w = to_mix(params_host);
    a.obs = obs; a.row_index = reinterpret_cast<const long long*>(row_index); a.dones = dones; a.q_all = q_all; a.actions = actions;
    a.save = save; a.dqtot = dqtot; a.dq = dq; a.partials = partials; a.B = B; a.L = L; a.S = S; a.NA = NA;
    const size_t smem = bwd_smem(pad8(S));
    const int grid = (B + RT - 1) / RT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaFuncSetAttribute(mixnet_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    mixnet_bwd_kernel<<<grid, BT, smem, st>>>(a);
    if (int rc = mm::check_launch()) return rc;
    MixReduceArgs r{};
    r.partials = partials; r.n_tiles = grid; r.P = mix_param_count(S); r.accumulate = accumulate;
    const int sizes[kMixTensors] = {96 * S, 96 * 32, 96, 96, 64 * 32, 64, 32 * 32, 32, 32 * 32, 32, 32 * 32, 32, 32, 1};
    for (int k = 0; k < kMixTensors; ++k) { r.grad[k] = grads_host[k]; r.sizes[k] = sizes[k]; }
    mix_grad_reduce_kernel<<<(r.P + 255) / 256, 256, 0, st>>>(r);
    return mm::check_launch();
}

extern "C" int mm_qmix_chunk_loss(const float* qtot, const float* qtot_tgt, const float* rewards, const float* dones,
                                  const int64_t* row_index, const float* is_weight, float* se, float* dqtot, float* td_last,
                                  int32_t B, int32_t L, float gamma, void* stream) {
    if (!qtot || !qtot_tgt || !rewards || !dones || !is_weight || !se || !dqtot || !td_last || B <= 0 || L <= 0) return MM_ERR_BAD_ARG;
    const long long n = (long long)B * L;
    qmix_chunk_loss_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        qtot, qtot_tgt, rewards, dones, reinterpret_cast<const long long*>(row_index), is_weight, se, dqtot, td_last, B, L, gamma);
    return mm::check_launch();
}
