// Fused kernels for the per-agent Q_Net of the top-level vdn/ and qmix/ packages (8f row 3).
//
//   Q_Net (vdn/_network.py:61-91, qmix/_network.py:5-77), one set of weights PER AGENT:
//       Linear(D,64) ReLU Linear(64,32) ReLU -> GRUCell(32,32) -> Linear(32,nA)
//
// * toplevel_rollout_kernel: the whole rollout body of vdn/main.py:139-177 / qmix/main.py:183-240 for 32 lockstep envs
//   per CTA and ALL steps of an episode in one launch: behaviour Q_Net of both agents -> epsilon_greedy with ONE uniform
//   per env (vdn/_network.py:52-58) -> Checkers / Switch2 step (device functions of the env kernels) -> next observation
//   expanded straight into the operand tile of the next step; states / actions / rewards / next_states / dones go out in
//   the [chunk][env][step] layout the chunk PER stores.
// * qnet_seq_kernel: the same network over recorded sequences (target network over next_states for cal_td_error,
//   vdn/_utils.py:44-52; behaviour / target pass of Target_Dqn.train / Train_dqn.train over sampled chunks,
//   vdn/_train.py:66-86), optionally keeping the activations the backward needs.
// * qnet_bwd_kernel: BPTT through the chunk (GRUCell, two ReLU layers) with every weight gradient accumulated in
//   registers across rows and steps; one partial per CTA, reduced in a fixed order by qnet_grad_reduce_kernel.
//
// Everything is fp32 FFMA (the networks are 14 k MACs per row and step; parity with the PyTorch modules is 1e-5), operands
// in shared memory: weights transposed so that lane = output unit reads consecutive words, activations read as broadcast
// float4.  A block of 128 threads ("agent block") owns 32 rows of ONE agent: thread (warp g, lane j) computes output units
// j (and j + 32) of rows 8g .. 8g+7.
#include "checkers_core.cuh"
#include "common.cuh"
#include "switch_core.cuh"

namespace {

constexpr int F1 = 64, F2 = 32, HH = 32, NAMAX = 8;
constexpr int RT = 32;        // rows (envs) per CTA at most
constexpr int RPT = 4;        // rows per warp in the forward kernels: a warp owns its rows through every layer and step
constexpr int BT = 32 * (RT / RPT);   // threads per CTA at most (forward)
constexpr int BTB = 128;      // threads per CTA of the backward kernel (8 rows per thread: its registers hold the gradients)
constexpr int HS = 36;        // row stride of the hidden tiles (conflict-free float4 reads by row)
constexpr int SV = 256;       // saved floats per (agent, step, row): x1 64 | x2 32 | r | z | n | gh_n | h (32 each)
constexpr int kMaxD = 96;

struct QNetP {                // reference layout (nn.Linear / nn.GRUCell: weight [out][in])
    const float *w1, *b1, *w2, *b2, *wih, *whh, *bih, *bhh, *wo, *bo;
};

__host__ __device__ inline int pad4(int d) { return (d + 3) & ~3; }
__host__ __device__ inline int pad8(int d) { return (d + 7) & ~7; }

// shared-memory floats: one agent's weights; one warp's activation tiles (one agent, n_x0 observation buffers)
__host__ __device__ inline int weight_floats(int Dp) { return Dp * F1 + F1 + F1 * F2 + F2 + HH * 6 * HH + 6 * HH + HH * NAMAX + NAMAX; }
__host__ __device__ inline int tile_floats(int Dp, int n_x0, int R = RPT) { return R * (n_x0 * Dp + F1 + F2 + 2 * HS + NAMAX); }

// Weight matrices are stored as [k / 4][unit][4]: the four consecutive-k weights of one output unit are one 16-byte word,
// lane = unit reads consecutive words (one conflict-free LDS.128 per unit and four k); the gate weights as
// [k / 4][gate][unit][4] with gates 0..2 = W_ih (r, z, n), 3..5 = W_hh.
struct Wts {
    const float *W1q, *b1, *W2q, *b2, *Wgq, *bg, *Wot, *bo;
};
struct Tile {                 // activation tiles of RPT rows, private to one warp
    float *x0, *x1, *x2, *h, *q;
};

__device__ __forceinline__ Wts carve_weights(float* p, int Dp) {
    Wts w;
    w.W1q = p; p += Dp * F1;
    w.b1 = p; p += F1;
    w.W2q = p; p += F1 * F2;
    w.b2 = p; p += F2;
    w.Wgq = p; p += HH * 6 * HH;
    w.bg = p; p += 6 * HH;
    w.Wot = p; p += HH * NAMAX;
    w.bo = p;
    return w;
}
__device__ __forceinline__ Tile carve_tile(float* p, int Dp, int n_x0, int R = RPT) {
    Tile t;
    t.x0 = p; p += R * n_x0 * Dp;
    t.x1 = p; p += R * F1;
    t.x2 = p; p += R * F2;
    t.h = p; p += 2 * R * HS;
    t.q = p;
    return t;
}

// stages one agent's weights (reference layout in global memory, read with unit stride) into the shared layout
__device__ void stage_weights(float* base, const QNetP& w, int D, int Dp, int NA, int tid, int nthreads) {
    float* W1q = base;
    float* b1 = W1q + Dp * F1;
    float* W2q = b1 + F1;
    float* b2 = W2q + F1 * F2;
    float* Wgq = b2 + F2;
    float* bg = Wgq + HH * 6 * HH;
    float* Wot = bg + 6 * HH;
    float* bo = Wot + HH * NAMAX;
    for (int i = tid; i < F1 * Dp; i += nthreads) {             // i = n * Dp + k
        const int n = i / Dp, k = i - n * Dp;
        W1q[((k >> 2) * F1 + n) * 4 + (k & 3)] = k < D ? w.w1[n * D + k] : 0.f;
    }
    for (int i = tid; i < F1; i += nthreads) b1[i] = w.b1[i];
    for (int i = tid; i < F2 * F1; i += nthreads) {             // i = n * 64 + k
        const int n = i >> 6, k = i & 63;
        W2q[((k >> 2) * F2 + n) * 4 + (k & 3)] = w.w2[i];
    }
    for (int i = tid; i < F2; i += nthreads) b2[i] = w.b2[i];
    for (int i = tid; i < 3 * HH * HH; i += nthreads) {         // i = (gate * 32 + j) * 32 + k
        const int gj = i >> 5, k = i & 31, gate = gj >> 5, j = gj & 31;
        Wgq[(((k >> 2) * 6 + gate) * HH + j) * 4 + (k & 3)] = w.wih[i];
        Wgq[(((k >> 2) * 6 + gate + 3) * HH + j) * 4 + (k & 3)] = w.whh[i];
    }
    for (int i = tid; i < 6 * HH; i += nthreads) bg[i] = i < 3 * HH ? w.bih[i] : w.bhh[i - 3 * HH];
    for (int i = tid; i < HH * NAMAX; i += nthreads) {
        const int k = i / NAMAX, o = i - k * NAMAX;
        Wot[i] = o < NA ? w.wo[o * HH + k] : 0.f;
    }
    for (int i = tid; i < NAMAX; i += nthreads) bo[i] = i < NA ? w.bo[i] : 0.f;
}

__device__ __forceinline__ float sigmoid_(float x) { return __frcp_rn(1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { return 1.0f - 2.0f * __frcp_rn(1.0f + __expf(2.0f * x)); }
__device__ __forceinline__ float dot4(const float4& x, const float4& w, float acc) {
    acc = fmaf(x.x, w.x, acc); acc = fmaf(x.y, w.y, acc); acc = fmaf(x.z, w.z, acc);
    return fmaf(x.w, w.w, acc);
}

// One forward step of one warp's RPT rows of one agent: x0 tile -> q tile; hidden state h[cur] -> h[cur ^ 1].  Lane = output
// unit.  Every layer reads only what this warp wrote, so the only synchronisation is __syncwarp: the warps of a CTA share
// the weights and nothing else, and run through the time loop independently of each other.
// `sv` (SAVE): this warp's rows of the activation record of this step, row stride SV; rows >= n_valid are not stored.
template <bool SAVE, int R>
__device__ __forceinline__ void qnet_step_warp(const Wts& w, const Tile& t, const float* x0, int Dp, int cur, int lane, float* sv,
                                               int n_valid) {
    const int j = lane;
    // ---- fc1: [RPT x Dp] x [Dp x 64] -------------------------------------------------------------------------------
    {
        float a0[R], a1[R];
        const float c0 = w.b1[j], c1 = w.b1[j + 32];
#pragma unroll
        for (int i = 0; i < R; ++i) { a0[i] = c0; a1[i] = c1; }
        const float4* wq = reinterpret_cast<const float4*>(w.W1q) + j;
#pragma unroll 2
        for (int k4 = 0; k4 < (Dp >> 2); ++k4) {
            const float4 wa = wq[k4 * F1], wb = wq[k4 * F1 + 32];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float4 x = *reinterpret_cast<const float4*>(x0 + i * Dp + 4 * k4);
                a0[i] = dot4(x, wa, a0[i]);
                a1[i] = dot4(x, wb, a1[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const float v0 = fmaxf(a0[i], 0.f), v1 = fmaxf(a1[i], 0.f);
            t.x1[i * F1 + j] = v0;
            t.x1[i * F1 + j + 32] = v1;
            if (SAVE && i < n_valid) { sv[i * SV + j] = v0; sv[i * SV + j + 32] = v1; }
        }
    }
    __syncwarp();
    // ---- fc2: [RPT x 64] x [64 x 32] --------------------------------------------------------------------------------
    {
        float a[R];
        const float c = w.b2[j];
#pragma unroll
        for (int i = 0; i < R; ++i) a[i] = c;
        const float4* wq = reinterpret_cast<const float4*>(w.W2q) + j;
#pragma unroll 4
        for (int k4 = 0; k4 < F1 / 4; ++k4) {
            const float4 ww = wq[k4 * F2];
#pragma unroll
            for (int i = 0; i < R; ++i) a[i] = dot4(*reinterpret_cast<const float4*>(t.x1 + i * F1 + 4 * k4), ww, a[i]);
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const float v = fmaxf(a[i], 0.f);
            t.x2[i * F2 + j] = v;
            if (SAVE && i < n_valid) sv[i * SV + 64 + j] = v;
        }
    }
    __syncwarp();
    // ---- GRUCell: gi = W_ih x2 + b_ih, gh = W_hh h + b_hh; r, z, n; h' = (1 - z) n + z h ------------------------------
    const float* hc = t.h + cur * R * HS;
    float* hn = t.h + (cur ^ 1) * R * HS;
    {
        float ar[R], az[R], ai[R], ah[R];
        const float cr = w.bg[j] + w.bg[3 * HH + j], cz = w.bg[HH + j] + w.bg[4 * HH + j];
        const float ci = w.bg[2 * HH + j], ch = w.bg[5 * HH + j];
#pragma unroll
        for (int i = 0; i < R; ++i) { ar[i] = cr; az[i] = cz; ai[i] = ci; ah[i] = ch; }
        const float4* wq = reinterpret_cast<const float4*>(w.Wgq) + j;
#pragma unroll 2
        for (int k4 = 0; k4 < HH / 4; ++k4) {
            const float4 wir = wq[(k4 * 6 + 0) * HH], wiz = wq[(k4 * 6 + 1) * HH], win = wq[(k4 * 6 + 2) * HH];
            const float4 whr = wq[(k4 * 6 + 3) * HH], whz = wq[(k4 * 6 + 4) * HH], whn = wq[(k4 * 6 + 5) * HH];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const float4 x = *reinterpret_cast<const float4*>(t.x2 + i * F2 + 4 * k4);
                const float4 h = *reinterpret_cast<const float4*>(hc + i * HS + 4 * k4);
                ar[i] = dot4(h, whr, dot4(x, wir, ar[i]));
                az[i] = dot4(h, whz, dot4(x, wiz, az[i]));
                ai[i] = dot4(x, win, ai[i]);
                ah[i] = dot4(h, whn, ah[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const float r = sigmoid_(ar[i]), z = sigmoid_(az[i]);
            const float n = tanh_(fmaf(r, ah[i], ai[i]));
            const float hp = hc[i * HS + j];
            const float hv = fmaf(z, hp - n, n);
            hn[i * HS + j] = hv;
            if (SAVE && i < n_valid) {
                float* o = sv + i * SV + 96 + j;
                o[0] = r; o[32] = z; o[64] = n; o[96] = ah[i]; o[128] = hv;
            }
        }
    }
    __syncwarp();
    // ---- head: lane = (row, output): RPT * NAMAX = 32 ----------------------------------------------------------------------
    {
        static_assert(R * NAMAX <= 32, "head mapping: one (row, output) per lane");
        const int r = (lane >> 3) % R, o = lane & 7;               // lanes beyond R * NAMAX repeat rows (same values written)
        float q = w.bo[o];
#pragma unroll
        for (int k = 0; k < HH; k += 4) {
            const float4 h = *reinterpret_cast<const float4*>(hn + r * HS + k);
            const float* wk = w.Wot + k * NAMAX + o;
            q = fmaf(h.x, wk[0], q); q = fmaf(h.y, wk[NAMAX], q); q = fmaf(h.z, wk[2 * NAMAX], q); q = fmaf(h.w, wk[3 * NAMAX], q);
        }
        t.q[r * NAMAX + o] = q;
    }
    __syncwarp();
}

// ====================================================================================================================
// Rollout: behaviour network + epsilon-greedy + env step, all steps of an episode in one launch
// ====================================================================================================================
enum { ENV_CHECKERS = 1, ENV_SWITCH = 2 };

struct RollArgs {
    QNetP w[2];
    float* states;        // [C][N][L][A][D]
    float* next_states;   // [C][N][L][A][D]
    float* actions;       // [C][N][L][A]  (float, like the reference's action tensors)
    float* rewards;       // [C][N][L][A]
    float* dones;         // [C][N][L]     all(done) as float
    float* behv;          // [T][N]        sum_a Q_beh[a][action_a]
    uint8_t* alive;       // [T][N]        the env had not finished before this step
    float* score;         // [N]           sum of rewards of all agents over the steps the env was alive
    uint32_t* env_state;  // optional: final env registers, 3 planes of N
    const int32_t* forced_actions;   // [T][N][A] or null (tests: replay a given action sequence)
    int N, T, L, D, NA, full_obs;
    float eps;
    uint2 seed;
    unsigned long long step_offset;
};

template <int ENV> struct EnvCfg;
template <> struct EnvCfg<ENV_CHECKERS> { using type = mm_checkers_cfg; };
template <> struct EnvCfg<ENV_SWITCH> { using type = mm_switch_cfg; };

// One warp = RE envs with BOTH agents: the warp runs agent 0's network on its RE rows, then agent 1's, then lanes
// 0 .. 2 RE - 1 (lane = 2 env + agent) pick the actions and step the envs, and all lanes expand the next observations into
// the warp's own operand tiles.  No block-level synchronisation inside the time loop.
constexpr int RE = 2;         // envs per warp in the rollout kernel (the sequential env phase wants many warps per SM)
constexpr int BTR = 512;      // its CTA size at most: 16 warps = 32 envs

template <int ENV>
__global__ void __launch_bounds__(BTR, 1)
toplevel_rollout_kernel(const __grid_constant__ RollArgs a, const __grid_constant__ typename EnvCfg<ENV>::type cfg) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(16) uint32_t desc_all[BTR / 32][RE][2][4];     // per (env, agent): what the observation is expanded from
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Dp = pad4(a.D);
    const int wf = weight_floats(Dp), tf_ = tile_floats(Dp, 1, RE);
    stage_weights(smem, a.w[0], a.D, Dp, a.NA, tid, blockDim.x);
    stage_weights(smem + wf, a.w[1], a.D, Dp, a.NA, tid, blockDim.x);
    const Wts W0 = carve_weights(smem, Dp), W1 = carve_weights(smem + wf, Dp);
    float* tiles = smem + 2 * wf + warp * 2 * tf_;
    const Tile tl0 = carve_tile(tiles, Dp, 1, RE), tl1 = carve_tile(tiles + tf_, Dp, 1, RE);     // agent 1's tiles sit tf_ floats after agent 0's
    for (int i = lane; i < 2 * tf_; i += 32) tiles[i] = 0.f;
    uint32_t (*desc)[2][4] = desc_all[warp];
    const int env0 = (blockIdx.x * (blockDim.x >> 5) + warp) * RE;
    const int n_valid = max(0, min(RE, a.N - env0));
    __syncthreads();

    // env registers: lane = 2 * env + agent for the first 2 * RPT lanes
    const int el = (lane >> 1) & (RE - 1), agent = lane & 1;
    const bool env_lane = lane < 2 * RE;
    ck::EnvRegs cs;
    ck::reset_regs(cs);
    uint32_t sw_cells = sw::init_cells(2), sw_done = 0, sw_step = 0;
    float score = 0.f;

    auto describe = [&]() {                       // this lane's agent observation, in compact form
        if constexpr (ENV == ENV_CHECKERS) {
            const mm_checkers_cfg& c = cfg;
            const uint32_t mine = (cs.cells >> (8 * agent)) & 0xffu;
            const int r = mine >> 3, col = mine & 7;
            const uint64_t m = ck::neighbourhood_mask(cs.food, cs.cells, r, col);
            *reinterpret_cast<uint4*>(desc[el][agent]) =
                make_uint4((uint32_t)m, (uint32_t)(m >> 32), __float_as_uint(c.pos_row[r]), __float_as_uint(c.pos_col[col]));
        } else {
            const mm_switch_cfg& c = cfg;
            const uint32_t cell = (sw_cells >> (8 * agent)) & 0xffu;
            *reinterpret_cast<uint4*>(desc[el][agent]) =
                make_uint4(__float_as_uint(c.pos_row[cell >> 3]), __float_as_uint(c.pos_col[cell & 7]),
                           __float_as_uint(sw::clock_value(sw_step, c)), 0u);
        }
    };
    auto obs_value = [&](int row, int tile_agent, int d) -> float {
        int src = tile_agent, dd = d;
        if constexpr (ENV == ENV_CHECKERS) {
            if (a.full_obs) { src = d >= ck::kObs; dd = d - src * ck::kObs; }
            const uint32_t* q = desc[row][src];
            if (dd < 2) return __uint_as_float(q[2 + dd]);
            const int b = dd - 2;
            return ((b < 32 ? q[0] >> b : q[1] >> (b - 32)) & 1u) ? 1.0f : 0.0f;
        } else {
            return __uint_as_float(desc[row][src][dd]);
        }
    };
    // expands the observation of every (row, agent) of this warp into the operand tiles and streams it to up to two global slots
    auto expand = [&](float* g0, float* g1, int l0, int l1) {
#pragma unroll 1
        for (int ra = 0; ra < 2 * RE; ++ra) {
            const int row = ra >> 1, ag = ra & 1;
            if (row >= n_valid) break;
            const long long e = env0 + row;
            float* x0 = tl0.x0 + ag * tf_ + row * Dp;
            for (int d = lane; d < a.D; d += 32) {
                const float v = obs_value(row, ag, d);
                x0[d] = v;
                if (g0) g0[((e * a.L + l0) * 2 + ag) * a.D + d] = v;
                if (g1) g1[((e * a.L + l1) * 2 + ag) * a.D + d] = v;
            }
        }
    };

    if (env_lane) describe();
    __syncwarp();
    expand(a.states, nullptr, 0, 0);
    __syncwarp();

    int cur = 0;
    for (int t = 0; t < a.T; ++t) {
        const int c = t / a.L, l = t - c * a.L;
        qnet_step_warp<false, RE>(W0, tl0, tl0.x0, Dp, cur, lane, nullptr, n_valid);
        qnet_step_warp<false, RE>(W1, tl1, tl1.x0, Dp, cur, lane, nullptr, n_valid);
        cur ^= 1;
        // ---- epsilon-greedy + env step (lanes 0 .. 2 RPT - 1 hold the envs; every lane runs the shuffles) ------------------
        {
            const int e = env0 + el;
            const bool valid = env_lane && el < n_valid;
            const float* q = tl0.q + agent * tf_ + el * NAMAX;
            int act = 0;
            float best = q[0];
            for (int o = 1; o < a.NA; ++o)
                if (q[o] > best) { best = q[o]; act = o; }
            if (a.forced_actions) {
                if (valid) act = a.forced_actions[((long long)t * a.N + e) * 2 + agent];
            } else {
                const unsigned long long ctr = a.step_offset + (unsigned long long)t;
                const uint4 rnd = mm::philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)e, 0u), a.seed);
                const float u = (float)(rnd.x >> 8) * (1.0f / 16777216.0f);
                if (u <= a.eps && a.eps > 0.f) act = (int)__umulhi(agent ? rnd.z : rnd.y, (uint32_t)a.NA);
            }
            float qa = q[min(max(act, 0), a.NA - 1)];
            qa += __shfl_xor_sync(MM_FULL_MASK, qa, 1);
            float reward;
            uint32_t done_all, was_alive;
            if constexpr (ENV == ENV_CHECKERS) {
                was_alive = (cs.food & ck::kDoneLatch) ? 0u : 1u;
                uint32_t done;
                bool bad;
                ck::transition(cs, act, agent, lane, cfg, reward, done, bad);
                done_all = done;                       // Checkers: all agents finish together
            } else {
                was_alive = sw_done == 3u ? 0u : 1u;
                sw::transition<2>(sw_cells, sw_done, sw_step, act, agent, lane, cfg, reward);
                done_all = sw_done == 3u;
            }
            const float rsum = reward + __shfl_xor_sync(MM_FULL_MASK, reward, 1);
            if (valid) {
                const long long ce = ((long long)c * a.N + e) * a.L + l;
                a.actions[ce * 2 + agent] = (float)act;
                a.rewards[ce * 2 + agent] = reward;
                if (was_alive) score += rsum;
                if (agent == 0) {
                    a.dones[ce] = done_all ? 1.0f : 0.0f;
                    a.behv[(long long)t * a.N + e] = qa;
                    a.alive[(long long)t * a.N + e] = (uint8_t)was_alive;
                }
            }
            if (env_lane) describe();
        }
        __syncwarp();
        {
            float* ns = a.next_states + (long long)c * a.N * a.L * 2 * a.D;
            const int t1 = t + 1, c1 = t1 / a.L, l1 = t1 - c1 * a.L;
            float* st = t1 < a.T ? a.states + (long long)c1 * a.N * a.L * 2 * a.D : nullptr;
            expand(ns, st, l, l1);
        }
        __syncwarp();
    }
    if (env_lane && el < n_valid && agent == 0) {
        a.score[env0 + el] = score;
        if (a.env_state) {
            const int e = env0 + el;
            if constexpr (ENV == ENV_CHECKERS) {
                a.env_state[e] = cs.food; a.env_state[a.N + e] = cs.cells; a.env_state[2 * a.N + e] = cs.step;
            } else {
                a.env_state[e] = sw_cells; a.env_state[a.N + e] = sw_done; a.env_state[2 * a.N + e] = sw_step;
            }
        }
    }
}

// ====================================================================================================================
// Sequence forward over recorded observations
// ====================================================================================================================
struct SeqArgs {
    QNetP w[2];
    const float* obs;              // element (c, row, l, agent, d) at c * chunk_stride + ((idx(row) * L + l) * A + agent) * D + d
    const long long* row_index;    // [B] storage row of batch row b (PER slots), or null = identity
    long long chunk_stride;
    const float* dones;            // (c, row, l) at c * done_chunk_stride + idx(row) * L + l; used when reset_on_done
    long long done_chunk_stride;
    float* q_out;                  // [C][B][L][A][NA]
    float* save;                   // [A][C*L][B][SV] or null
    int B, L, C, D, NA, reset_on_done;
};

__device__ __forceinline__ void cp_async4(void* smem_ptr, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_ptr)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// grid (row tiles, agent); one warp = RPT rows of that agent, independent of the other warps after the weights are staged
template <bool SAVE>
__global__ void __launch_bounds__(BT, 2) qnet_seq_kernel(const __grid_constant__ SeqArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, agent = blockIdx.y;
    const int Dp = pad4(a.D);
    const int wf = weight_floats(Dp), tf_ = tile_floats(Dp, 2);
    stage_weights(smem, a.w[agent], a.D, Dp, a.NA, tid, blockDim.x);
    const Wts W = carve_weights(smem, Dp);
    float* tiles = smem + wf + warp * tf_;
    const Tile tl = carve_tile(tiles, Dp, 2);
    for (int i = lane; i < tf_; i += 32) tiles[i] = 0.f;
    const int b0 = (blockIdx.x * (blockDim.x >> 5) + warp) * RPT;
    const int n_valid = max(0, min(RPT, a.B - b0));
    long long rows[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) rows[i] = i < n_valid ? (a.row_index ? a.row_index[b0 + i] : (long long)(b0 + i)) : 0;
    __syncthreads();
    if (n_valid == 0) return;
    const int T = a.C * a.L;
    auto fetch = [&](int t, int buf) {
        const int c = t / a.L, l = t - c * a.L;
        const float* base = a.obs + (long long)c * a.chunk_stride;
        float* dst = tl.x0 + buf * RPT * Dp;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            if (i < n_valid) {
                const float* src = base + ((rows[i] * a.L + l) * 2 + agent) * a.D;
                for (int d = lane; d < a.D; d += 32) cp_async4(dst + i * Dp + d, src + d);
            }
        }
        cp_async_commit();
    };
    fetch(0, 0);
    int cur = 0;
    for (int t = 0; t < T; ++t) {
        const int c = t / a.L, l = t - c * a.L;
        cp_async_wait_all();
        __syncwarp();
        if (t + 1 < T) fetch(t + 1, (t + 1) & 1);
        float* sv = SAVE ? a.save + (((long long)agent * T + t) * a.B + b0) * SV : nullptr;
        qnet_step_warp<SAVE, RPT>(W, tl, tl.x0 + (t & 1) * RPT * Dp, Dp, cur, lane, sv, n_valid);
        cur ^= 1;
        // q tile -> global (lane = (row, output)); hidden reset where the recorded step was terminal
        {
            const int row = lane >> 3, o = lane & 7;
            if (row < n_valid) {
                long long srow = rows[0];
#pragma unroll
                for (int i = 1; i < RPT; ++i) srow = row == i ? rows[i] : srow;
                if (o < a.NA) a.q_out[((((long long)c * a.B + b0 + row) * a.L + l) * 2 + agent) * a.NA + o] = tl.q[row * NAMAX + o];
                if (a.reset_on_done) {
                    const float d = a.dones[(long long)c * a.done_chunk_stride + srow * a.L + l];
                    if (d != 0.f) {
                        float* hc = tl.h + cur * RPT * HS + row * HS + o * 4;
                        const float keep = 1.0f - d;
                        hc[0] *= keep; hc[1] *= keep; hc[2] *= keep; hc[3] *= keep;
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ====================================================================================================================
// Backward through a chunk
// ====================================================================================================================
struct BwdArgs {
    QNetP w[2];
    const float* obs;              // as SeqArgs (C = 1)
    const long long* row_index;
    const float* dones;            // [rows][L] through row_index (hidden resets), or null
    const float* save;             // [A][L][B][SV]
    const float* dq;               // [B][L][A][NA]
    float* partials;               // [A][n_tiles][P]  P = param_count(D, NA), reference layout order
    int B, L, D, NA, reset_on_done;
};

__host__ __device__ inline int param_count(int D, int NA) {
    return F1 * D + F1 + F2 * F1 + F2 + 3 * HH * F2 + 3 * HH * HH + 3 * HH + 3 * HH + NA * HH + NA;
}

__global__ void __launch_bounds__(BTB, 2) qnet_bwd_kernel(const __grid_constant__ BwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ long long rows[RT];
    const int bt = threadIdx.x, lane = bt & 31, warp = bt >> 5, agent = blockIdx.y;
    const int j = lane, g = warp, r0 = g * 8;
    const int Dp = pad8(a.D);                // the dW1 ownership splits the columns into 8 groups
    // shared: weights in reference layout (no W1: inputs need no gradient), then tiles
    float* W2 = smem;                       // [32][64]
    float* Wih = W2 + F2 * F1;              // [96][32]
    float* Whh = Wih + 3 * HH * F2;         // [96][32]
    float* Wo = Whh + 3 * HH * HH;          // [NAMAX][32]
    float* X0 = Wo + NAMAX * HH;            // [RT][Dp]
    float* X1 = X0 + RT * Dp;               // [RT][64]
    float* X2 = X1 + RT * F1;               // [RT][32]
    float* HP = X2 + RT * F2;               // [RT][HS]   masked input state of this step
    float* GI = HP + RT * HS;               // [RT][96]   d(pre r) | d(pre z) | d(pre n)
    float* GH = GI + RT * 96;               // [RT][96]   d(pre r) | d(pre z) | d(pre n) * r
    float* A2 = GH + RT * 96;               // [RT][32]
    float* A1 = A2 + RT * F2;               // [RT][64]
    float* DQ = A1 + RT * F1;               // [RT][NAMAX]
    float* HN = DQ + RT * NAMAX;            // [RT][HS]   output state of this step (head input)
    const QNetP& w = a.w[agent];
    for (int i = bt; i < F2 * F1; i += BTB) W2[i] = w.w2[i];
    for (int i = bt; i < 3 * HH * F2; i += BTB) { Wih[i] = w.wih[i]; Whh[i] = w.whh[i]; }
    for (int i = bt; i < NAMAX * HH; i += BTB) Wo[i] = i < a.NA * HH ? w.wo[i] : 0.f;
    for (int i = bt; i < RT * (Dp + F1 + F2 + HS); i += BTB) X0[i] = 0.f;       // X0, X1, X2, HP: rows past the batch stay zero
    for (int i = bt; i < RT * HS; i += BTB) HN[i] = 0.f;
    const int b0 = blockIdx.x * RT;
    const int n_valid = min(RT, a.B - b0);
    if (bt < RT) rows[bt] = bt < n_valid ? (a.row_index ? a.row_index[b0 + bt] : (long long)(b0 + bt)) : 0;
    __syncthreads();

    // weight-gradient accumulators (registers), ownership chosen so that the tile reads are vector loads:
    //   dW1 [64][Dp]: 4 n x (Dp / 8) k per thread  -> n = 4 (bt & 15) + u, k = (Dp / 8) (bt >> 4) + v
    //   dW2 [32][64]: 4 n x 4 k                     -> n = 4 (bt & 7) + u,  k = 4 (bt >> 3) + v
    //   dWih, dWhh [96][32]: 6 m x 4 k              -> m = 6 (bt & 15) + u, k = 4 (bt >> 4) + v
    //   dWo [NA][32]: o = g and o = 4 + g, k = j
    constexpr int KW1 = kMaxD / 8;           // 12 columns of dW1 per thread at most
    float gW1[4][KW1], gW2[4][4], gWih[6][4], gWhh[6][4], gWoA = 0.f, gWoB = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int v = 0; v < KW1; ++v) gW1[u][v] = 0.f;
#pragma unroll
        for (int v = 0; v < 4; ++v) gW2[u][v] = 0.f;
    }
#pragma unroll
    for (int u = 0; u < 6; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) { gWih[u][v] = 0.f; gWhh[u][v] = 0.f; }
    float gb = 0.f, gb2 = 0.f, gbo = 0.f;    // bias gradients: b1 / b2 (bt < 96), bih / bhh (bt < 96), bo (bt < NA)
    float gbih = 0.f, gbhh = 0.f;
    const int kw = Dp / 8;                   // dW1 columns per thread

    float carry[8];                          // d loss / d h_out of the current step coming from the step after it
#pragma unroll
    for (int i = 0; i < 8; ++i) carry[i] = 0.f;

    for (int l = a.L - 1; l >= 0; --l) {
        const float* sv = a.save + (((long long)agent * a.L + l) * a.B + b0) * SV;
        const float* svp = l > 0 ? a.save + (((long long)agent * a.L + l - 1) * a.B + b0) * SV : nullptr;
        // ---- tiles of this step -------------------------------------------------------------------------------------
        for (int i = bt; i < n_valid * (F1 / 4); i += BTB) {
            const int row = i >> 4, q4 = i & 15;
            *reinterpret_cast<float4*>(X1 + row * F1 + 4 * q4) = *reinterpret_cast<const float4*>(sv + row * SV + 4 * q4);
        }
        for (int i = bt; i < n_valid * (F2 / 4); i += BTB) {
            const int row = i >> 3, q4 = i & 7;
            *reinterpret_cast<float4*>(X2 + row * F2 + 4 * q4) = *reinterpret_cast<const float4*>(sv + row * SV + 64 + 4 * q4);
            *reinterpret_cast<float4*>(HN + row * HS + 4 * q4) = *reinterpret_cast<const float4*>(sv + row * SV + 224 + 4 * q4);
            float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (svp) {
                hp = *reinterpret_cast<const float4*>(svp + row * SV + 224 + 4 * q4);
                if (a.reset_on_done) {
                    const float keep = 1.0f - a.dones[rows[row] * a.L + l - 1];
                    hp.x *= keep; hp.y *= keep; hp.z *= keep; hp.w *= keep;
                }
            }
            *reinterpret_cast<float4*>(HP + row * HS + 4 * q4) = hp;
        }
#pragma unroll 1
        for (int rr = 0; rr < 8; ++rr) {
            const int row = warp * 8 + rr;
            if (row >= n_valid) break;
            const float* src = a.obs + ((rows[row] * a.L + l) * 2 + agent) * a.D;
            for (int d = lane; d < a.D; d += 32) X0[row * Dp + d] = src[d];
        }
        for (int i = bt; i < RT * NAMAX; i += BTB) {
            const int row = i >> 3, o = i & 7;
            DQ[i] = (row < n_valid && o < a.NA) ? a.dq[(((long long)(b0 + row) * a.L + l) * 2 + agent) * a.NA + o] : 0.f;
        }
        __syncthreads();
        // ---- B1: dh = carry + Wo^T dq; GRU cell backward (thread owns unit j of rows r0..r0+7) ---------------------------
        float dhz[8];                         // direct path dh * z to the input state
        {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + i;
                float dh = carry[i];
#pragma unroll
                for (int o = 0; o < NAMAX; ++o) dh = fmaf(DQ[row * NAMAX + o], Wo[o * HH + j], dh);
                float r = 0.f, z = 0.f, n = 0.f, ghn = 0.f;
                if (row < n_valid) {
                    const float* o = sv + row * SV + 96 + j;
                    r = o[0]; z = o[32]; n = o[64]; ghn = o[96];
                }
                const float hp = HP[row * HS + j];
                const float dn = dh * (1.f - z);
                const float dpn = dn * (1.f - n * n);
                const float dpz = dh * (hp - n) * z * (1.f - z);
                const float dpr = dpn * ghn * r * (1.f - r);
                dhz[i] = dh * z;
                GI[row * 96 + j] = dpr; GI[row * 96 + 32 + j] = dpz; GI[row * 96 + 64 + j] = dpn;
                GH[row * 96 + j] = dpr; GH[row * 96 + 32 + j] = dpz; GH[row * 96 + 64 + j] = dpn * r;
            }
        }
        __syncthreads();
        // ---- B2: dx2 = GI Wih, dhp = GH Whh + dh z ------------------------------------------------------------------------
        {
            float ax[8], ah[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { ax[i] = 0.f; ah[i] = dhz[i]; }
#pragma unroll 2
            for (int m = 0; m < 96; m += 4) {
                const float wi0 = Wih[m * F2 + j], wi1 = Wih[(m + 1) * F2 + j], wi2 = Wih[(m + 2) * F2 + j], wi3 = Wih[(m + 3) * F2 + j];
                const float wh0 = Whh[m * HH + j], wh1 = Whh[(m + 1) * HH + j], wh2 = Whh[(m + 2) * HH + j], wh3 = Whh[(m + 3) * HH + j];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 gi = *reinterpret_cast<const float4*>(GI + (r0 + i) * 96 + m);
                    const float4 gh = *reinterpret_cast<const float4*>(GH + (r0 + i) * 96 + m);
                    ax[i] = fmaf(gi.x, wi0, ax[i]); ax[i] = fmaf(gi.y, wi1, ax[i]); ax[i] = fmaf(gi.z, wi2, ax[i]); ax[i] = fmaf(gi.w, wi3, ax[i]);
                    ah[i] = fmaf(gh.x, wh0, ah[i]); ah[i] = fmaf(gh.y, wh1, ah[i]); ah[i] = fmaf(gh.z, wh2, ah[i]); ah[i] = fmaf(gh.w, wh3, ah[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + i;
                A2[row * F2 + j] = X2[row * F2 + j] > 0.f ? ax[i] : 0.f;
                float keep = 1.f;
                if (a.reset_on_done && l > 0 && row < n_valid) keep = 1.0f - a.dones[rows[row] * a.L + l - 1];
                carry[i] = ah[i] * keep;
            }
        }
        __syncthreads();
        // ---- B3: dx1 = A2 W2 (columns j, j + 32) ---------------------------------------------------------------------------
        {
            float a0[8], a1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
#pragma unroll 2
            for (int n = 0; n < F2; n += 4) {
                const float w00 = W2[n * F1 + j], w01 = W2[(n + 1) * F1 + j], w02 = W2[(n + 2) * F1 + j], w03 = W2[(n + 3) * F1 + j];
                const float w10 = W2[n * F1 + j + 32], w11 = W2[(n + 1) * F1 + j + 32], w12 = W2[(n + 2) * F1 + j + 32], w13 = W2[(n + 3) * F1 + j + 32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 d = *reinterpret_cast<const float4*>(A2 + (r0 + i) * F2 + n);
                    a0[i] = fmaf(d.x, w00, a0[i]); a0[i] = fmaf(d.y, w01, a0[i]); a0[i] = fmaf(d.z, w02, a0[i]); a0[i] = fmaf(d.w, w03, a0[i]);
                    a1[i] = fmaf(d.x, w10, a1[i]); a1[i] = fmaf(d.y, w11, a1[i]); a1[i] = fmaf(d.z, w12, a1[i]); a1[i] = fmaf(d.w, w13, a1[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + i;
                A1[row * F1 + j] = X1[row * F1 + j] > 0.f ? a0[i] : 0.f;
                A1[row * F1 + j + 32] = X1[row * F1 + j + 32] > 0.f ? a1[i] : 0.f;
            }
        }
        __syncthreads();
        // ---- B4: weight gradients, rank-1 updates over the rows of the tile --------------------------------------------------
        {
            const int n1 = 4 * (bt & 15), k1 = kw * (bt >> 4);
            const int n2 = 4 * (bt & 7), k2 = 4 * (bt >> 3);
            const int m3 = 6 * (bt & 15), k3 = 4 * (bt >> 4);
#pragma unroll 2
            for (int row = 0; row < RT; ++row) {
                {
                    const float4 gq = *reinterpret_cast<const float4*>(A1 + row * F1 + n1);
                    const float gv[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
                    for (int v = 0; v < KW1; ++v) {
                        if (v < kw) {
                            const float x = X0[row * Dp + k1 + v];
#pragma unroll
                            for (int u = 0; u < 4; ++u) gW1[u][v] = fmaf(gv[u], x, gW1[u][v]);
                        }
                    }
                }
                {
                    const float4 gq = *reinterpret_cast<const float4*>(A2 + row * F2 + n2);
                    const float4 xq = *reinterpret_cast<const float4*>(X1 + row * F1 + k2);
                    const float gv[4] = {gq.x, gq.y, gq.z, gq.w}, xv[4] = {xq.x, xq.y, xq.z, xq.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) gW2[u][v] = fmaf(gv[u], xv[v], gW2[u][v]);
                }
                {
                    const float2 i0 = *reinterpret_cast<const float2*>(GI + row * 96 + m3), i1 = *reinterpret_cast<const float2*>(GI + row * 96 + m3 + 2),
                                 i2 = *reinterpret_cast<const float2*>(GI + row * 96 + m3 + 4);
                    const float2 h0 = *reinterpret_cast<const float2*>(GH + row * 96 + m3), h1 = *reinterpret_cast<const float2*>(GH + row * 96 + m3 + 2),
                                 h2 = *reinterpret_cast<const float2*>(GH + row * 96 + m3 + 4);
                    const float4 xq = *reinterpret_cast<const float4*>(X2 + row * F2 + k3);
                    const float4 hq = *reinterpret_cast<const float4*>(HP + row * HS + k3);
                    const float gi[6] = {i0.x, i0.y, i1.x, i1.y, i2.x, i2.y}, gh[6] = {h0.x, h0.y, h1.x, h1.y, h2.x, h2.y};
                    const float xv[4] = {xq.x, xq.y, xq.z, xq.w}, hv[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
                    for (int u = 0; u < 6; ++u)
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            gWih[u][v] = fmaf(gi[u], xv[v], gWih[u][v]);
                            gWhh[u][v] = fmaf(gh[u], hv[v], gWhh[u][v]);
                        }
                }
                {
                    const float hn = HN[row * HS + j];
                    gWoA = fmaf(DQ[row * NAMAX + g], hn, gWoA);
                    gWoB = fmaf(DQ[row * NAMAX + 4 + g], hn, gWoB);
                }
                if (bt < F1) gb += A1[row * F1 + bt];
                else if (bt < F1 + F2) gb2 += A2[row * F2 + bt - F1];
                if (bt < 96) { gbih += GI[row * 96 + bt]; gbhh += GH[row * 96 + bt]; }
                if (bt < a.NA) gbo += DQ[row * NAMAX + bt];
            }
        }
        __syncthreads();
    }
    // ---- partial gradients of this CTA, reference layout ---------------------------------------------------------------------
    const int P = param_count(a.D, a.NA);
    float* out = a.partials + ((long long)agent * gridDim.x + blockIdx.x) * P;
    const int oW1 = 0, ob1 = oW1 + F1 * a.D, oW2 = ob1 + F1, ob2 = oW2 + F2 * F1, oWih = ob2 + F2, oWhh = oWih + 3 * HH * F2,
              obih = oWhh + 3 * HH * HH, obhh = obih + 3 * HH, oWo = obhh + 3 * HH, obo = oWo + a.NA * HH;
    {
        const int n1 = 4 * (bt & 15), k1 = kw * (bt >> 4);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < KW1; ++v)
                if (v < kw && k1 + v < a.D) out[oW1 + (n1 + u) * a.D + k1 + v] = gW1[u][v];
        const int n2 = 4 * (bt & 7), k2 = 4 * (bt >> 3);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) out[oW2 + (n2 + u) * F1 + k2 + v] = gW2[u][v];
        const int m3 = 6 * (bt & 15), k3 = 4 * (bt >> 4);
#pragma unroll
        for (int u = 0; u < 6; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                out[oWih + (m3 + u) * F2 + k3 + v] = gWih[u][v];
                out[oWhh + (m3 + u) * HH + k3 + v] = gWhh[u][v];
            }
        if (g < a.NA) out[oWo + g * HH + j] = gWoA;
        if (4 + g < a.NA) out[oWo + (4 + g) * HH + j] = gWoB;
        if (bt < F1) out[ob1 + bt] = gb;
        else if (bt < F1 + F2) out[ob2 + bt - F1] = gb2;
        if (bt < 96) { out[obih + bt] = gbih; out[obhh + bt] = gbhh; }
        if (bt < a.NA) out[obo + bt] = gbo;
    }
}

// sums the per-CTA partials in tile order and writes each block into the gradient tensor it belongs to
struct ReduceArgs {
    const float* partials;   // [A][n_tiles][P]
    float* grad[2][10];      // reference layout tensors, same order as QNetP
    int n_tiles, P, D, NA, accumulate;
};
__global__ void qnet_grad_reduce_kernel(const __grid_constant__ ReduceArgs a) {
    const int agent = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.P) return;
    const float* p = a.partials + (long long)agent * a.n_tiles * a.P + i;
    float sum = 0.f;
#pragma unroll 16                                     // same summation order, 16 loads in flight instead of 4
    for (int t = 0; t < a.n_tiles; ++t) sum += p[(long long)t * a.P];
    const int sizes[10] = {F1 * a.D, F1, F2 * F1, F2, 3 * HH * F2, 3 * HH * HH, 3 * HH, 3 * HH, a.NA * HH, a.NA};
    // QNetP order is w1 b1 w2 b2 wih whh bih bhh wo bo: the same as the partial layout
    int off = i;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        if (off < sizes[k]) {
            float* dst = a.grad[agent][k] + off;
            *dst = a.accumulate ? *dst + sum : sum;
            return;
        }
        off -= sizes[k];
    }
}

// ====================================================================================================================
// Element-wise pieces: cal_td_error over a rollout, the chunk loss of Target_Dqn.train
// ====================================================================================================================
// td_sum[c][e] = sum over the steps l of chunk c at which env e was still alive of
//     | sum_a r + (1 - done) gamma sum_a max_o Q_tgt[a][o] - sum_a Q_beh[a][action_a] |      (vdn/_utils.py:44-52)
// chunk_valid[c][e] = env e was alive at the first step of chunk c
__global__ void toplevel_td_kernel(const float* __restrict__ q_tgt, const float* __restrict__ rewards,
                                   const float* __restrict__ dones, const float* __restrict__ behv,
                                   const uint8_t* __restrict__ alive, float* __restrict__ td_sum,
                                   uint8_t* __restrict__ chunk_valid, int N, int L, int C, int NA, float gamma) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)C * N) return;
    const int c = (int)(i / N), e = (int)(i - (long long)c * N);
    float sum = 0.f;
    for (int l = 0; l < L; ++l) {
        const long long ce = ((long long)c * N + e) * L + l;
        const long long t = (long long)c * L + l;
        float mx = 0.f;
        for (int ag = 0; ag < 2; ++ag) {
            const float* q = q_tgt + (ce * 2 + ag) * NA;
            float best = q[0];
            for (int o = 1; o < NA; ++o) best = fmaxf(best, q[o]);
            mx += best;
        }
        const float r = rewards[ce * 2] + rewards[ce * 2 + 1];
        const float td = fabsf(r + (1.0f - dones[ce]) * gamma * mx - behv[t * N + e]);
        if (alive[t * N + e]) sum += td;
    }
    td_sum[i] = sum;
    chunk_valid[i] = alive[(long long)c * L * N + e];
}

// VDN chunk loss (vdn/_train.py:66-99) for batch row b: per step
//     sum_q = sum_a Q_beh[a][action_a],  y = w_b * sum_a (r_a + gamma (1 - done) sum_a' max_o Q_tgt[a'][o])   (quirk Q7)
//     se[b][l] = (y - sum_q)^2;  dq[b][l][a][action_a] = 2 (sum_q - y) / B;  td_last[b] = |y - sum_q| at l = L - 1
__global__ void vdn_chunk_loss_kernel(const float* __restrict__ q_beh, const float* __restrict__ q_tgt,
                                      const float* __restrict__ actions, const float* __restrict__ rewards,
                                      const float* __restrict__ dones, const long long* __restrict__ row_index,
                                      const float* __restrict__ is_weight, float* __restrict__ se, float* __restrict__ dq,
                                      float* __restrict__ td_last, int B, int L, int NA, float gamma) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * L) return;
    const int b = (int)(i / L), l = (int)(i - (long long)b * L);
    const long long srow = (row_index ? row_index[b] : (long long)b) * L + l;
    float sum_q = 0.f, mx = 0.f;
    int act[2];
    for (int ag = 0; ag < 2; ++ag) {
        const float* qb = q_beh + (i * 2 + ag) * NA;
        const float* qt = q_tgt + (i * 2 + ag) * NA;
        act[ag] = min(max((int)actions[srow * 2 + ag], 0), NA - 1);
        sum_q += qb[act[ag]];
        float best = qt[0];
        for (int o = 1; o < NA; ++o) best = fmaxf(best, qt[o]);
        mx += best;
    }
    const float boot = gamma * (1.0f - dones[srow]) * mx;
    const float y = is_weight[b] * ((rewards[srow * 2] + boot) + (rewards[srow * 2 + 1] + boot));
    const float diff = y - sum_q;
    se[i] = diff * diff;
    const float gq = -2.0f * diff / (float)B;
    for (int ag = 0; ag < 2; ++ag)
        for (int o = 0; o < NA; ++o) dq[(i * 2 + ag) * NA + o] = o == act[ag] ? gq : 0.f;
    if (l == L - 1) td_last[b] = fabsf(diff);
}

// warps per CTA (RPT rows each): the full 8 when there is more than one wave of work, otherwise the smallest CTA that still
// gives every SM slot one -- the kernels are sequential over time, so an idle SM is lost for the whole launch
int warps_per_cta(int n_rows, int ctas_per_sm, int grid_y, int rows_per_warp = RPT, int max_warps = BT / 32) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int slots = sms * ctas_per_sm / grid_y;
    if (slots < 1) slots = 1;
    const int rows = (n_rows + slots - 1) / slots;
    const int w = (rows + rows_per_warp - 1) / rows_per_warp;
    return w < 1 ? 1 : (w > max_warps ? max_warps : w);
}

int check_dims(int D, int NA) {
    if (D < 1 || D > kMaxD || NA < 1 || NA > NAMAX) return MM_ERR_UNSUPPORTED;
    return MM_OK;
}

QNetP to_p(const float* const* p) { return QNetP{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9]}; }

int bwd_smem_floats(int Dp) {
    return F2 * F1 + 2 * 3 * HH * F2 + NAMAX * HH + RT * Dp + RT * F1 + RT * F2 + RT * HS + 2 * RT * 96 + RT * F2 + RT * F1 +
           RT * NAMAX + RT * HS;
}

}  // namespace

extern "C" int mm_toplevel_qnet_param_count(int32_t D, int32_t NA) { return param_count(D, NA); }
extern "C" int mm_toplevel_qnet_save_stride(void) { return SV; }

// params: 2 x 10 device pointers (agent 0: w1 b1 w2 b2 wih whh bih bhh wo bo, then agent 1), reference layout
extern "C" int mm_toplevel_rollout(const float* const* params_host, int32_t env_kind, const void* env_cfg_host, float* states,
                                   float* next_states, float* actions, float* rewards, float* dones, float* behv, uint8_t* alive,
                                   float* score, uint32_t* env_state, const int32_t* forced_actions, int32_t N, int32_t T,
                                   int32_t L, int32_t D, int32_t NA, float eps, uint64_t seed, uint64_t step_offset,
                                   void* stream) {
    if (!params_host || !env_cfg_host || !states || !next_states || !actions || !rewards || !dones || !behv || !alive || !score)
        return MM_ERR_BAD_ARG;
    if (N <= 0 || T <= 0 || L <= 0 || T % L != 0) return MM_ERR_BAD_ARG;
    if (int rc = check_dims(D, NA)) return rc;
    RollArgs a{};
    a.w[0] = to_p(params_host); a.w[1] = to_p(params_host + 10);
    a.states = states; a.next_states = next_states; a.actions = actions; a.rewards = rewards; a.dones = dones; a.behv = behv;
    a.alive = alive; a.score = score; a.env_state = env_state; a.forced_actions = forced_actions;
    a.N = N; a.T = T; a.L = L; a.D = D; a.NA = NA; a.eps = eps;
    a.seed = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)); a.step_offset = step_offset;
    const int Dp = pad4(D);
    const int warps = warps_per_cta(N, 1, 1, RE, BTR / 32);
    const size_t smem = sizeof(float) * (size_t)(2 * weight_floats(Dp) + warps * 2 * tile_floats(Dp, 1, RE));
    if (smem > 227 * 1024) return MM_ERR_UNSUPPORTED;
    const int grid = (N + warps * RE - 1) / (warps * RE);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (env_kind == ENV_CHECKERS) {
        const mm_checkers_cfg* cfg = static_cast<const mm_checkers_cfg*>(env_cfg_host);
        a.full_obs = cfg->full_observable;
        if (D != (a.full_obs ? 2 : 1) * MM_CHECKERS_OBS_DIM || cfg->auto_reset) return MM_ERR_UNSUPPORTED;
        e = cudaFuncSetAttribute(toplevel_rollout_kernel<ENV_CHECKERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
        toplevel_rollout_kernel<ENV_CHECKERS><<<grid, 32 * warps, smem, st>>>(a, *cfg);
    } else if (env_kind == ENV_SWITCH) {
        const mm_switch_cfg* cfg = static_cast<const mm_switch_cfg*>(env_cfg_host);
        if (cfg->n_agents != 2 || cfg->full_observable || D != 2 + (cfg->clock ? 1 : 0)) return MM_ERR_UNSUPPORTED;
        e = cudaFuncSetAttribute(toplevel_rollout_kernel<ENV_SWITCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
        toplevel_rollout_kernel<ENV_SWITCH><<<grid, 32 * warps, smem, st>>>(a, *cfg);
    } else {
        return MM_ERR_BAD_ARG;
    }
    return mm::check_launch();
}

extern "C" int mm_toplevel_qnet_seq(const float* const* params_host, const float* obs, const int64_t* row_index,
                                    int64_t chunk_stride, const float* dones, int64_t done_chunk_stride, int32_t reset_on_done,
                                    float* q_out, float* save, int32_t B, int32_t L, int32_t C, int32_t D, int32_t NA,
                                    void* stream) {
    if (!params_host || !obs || !q_out || B <= 0 || L <= 0 || C <= 0) return MM_ERR_BAD_ARG;
    if (reset_on_done && !dones) return MM_ERR_BAD_ARG;
    if (int rc = check_dims(D, NA)) return rc;
    SeqArgs a{};
    a.w[0] = to_p(params_host); a.w[1] = to_p(params_host + 10);
    a.obs = obs; a.row_index = reinterpret_cast<const long long*>(row_index); a.chunk_stride = chunk_stride;
    a.dones = dones; a.done_chunk_stride = done_chunk_stride; a.reset_on_done = reset_on_done;
    a.q_out = q_out; a.save = save; a.B = B; a.L = L; a.C = C; a.D = D; a.NA = NA;
    const int Dp = pad4(D);
    const int per_sm = (sizeof(float) * (size_t)(weight_floats(Dp) + (BT / 32) * tile_floats(Dp, 2))) > 113 * 1024 ? 1 : 2;
    const int warps = warps_per_cta(B, per_sm, 2);
    const size_t smem = sizeof(float) * (size_t)(weight_floats(Dp) + warps * tile_floats(Dp, 2));
    const dim3 grid((B + warps * RPT - 1) / (warps * RPT), 2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (save) {
        e = cudaFuncSetAttribute(qnet_seq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
        qnet_seq_kernel<true><<<grid, 32 * warps, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(qnet_seq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
        qnet_seq_kernel<false><<<grid, 32 * warps, smem, st>>>(a);
    }
    return mm::check_launch();
}

// grads: 2 x 10 device pointers in the order of params; partials: float[2 * ceil(B / 32) * param_count] workspace
extern "C" int mm_toplevel_qnet_bwd(const float* const* params_host, float* const* grads_host, const float* obs,
                                    const int64_t* row_index, const float* dones, int32_t reset_on_done, const float* save,
                                    const float* dq, float* partials, int32_t accumulate, int32_t B, int32_t L, int32_t D,
                                    int32_t NA, void* stream) {
    if (!params_host || !grads_host || !obs || !save || !dq || !partials || B <= 0 || L <= 0) return MM_ERR_BAD_ARG;
    if (reset_on_done && !dones) return MM_ERR_BAD_ARG;
    if (int rc = check_dims(D, NA)) return rc;
    BwdArgs a{};
    a.w[0] = to_p(params_host); a.w[1] = to_p(params_host + 10);
    a.obs = obs; a.row_index = reinterpret_cast<const long long*>(row_index); a.dones = dones; a.reset_on_done = reset_on_done;
    a.save = save; a.dq = dq; a.partials = partials; a.B = B; a.L = L; a.D = D; a.NA = NA;
    const size_t smem = (size_t)bwd_smem_floats(pad8(D)) * sizeof(float);
    const dim3 grid((B + RT - 1) / RT, 2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaFuncSetAttribute(qnet_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    qnet_bwd_kernel<<<grid, BTB, smem, st>>>(a);
    if (int rc = mm::check_launch()) return rc;
    ReduceArgs r{};
    r.partials = partials; r.n_tiles = grid.x; r.P = param_count(D, NA); r.D = D; r.NA = NA; r.accumulate = accumulate;
    for (int ag = 0; ag < 2; ++ag)
        for (int k = 0; k < 10; ++k) r.grad[ag][k] = grads_host[ag * 10 + k];
    qnet_grad_reduce_kernel<<<dim3((r.P + 255) / 256, 2), 256, 0, st>>>(r);
    return mm::check_launch();
}

extern "C" int mm_toplevel_td(const float* q_tgt, const float* rewards, const float* dones, const float* behv,
                              const uint8_t* alive, float* td_sum, uint8_t* chunk_valid, int32_t N, int32_t L, int32_t C,
                              int32_t NA, float gamma, void* stream) {
    if (!q_tgt || !rewards || !dones || !behv || !alive || !td_sum || !chunk_valid || N <= 0 || L <= 0 || C <= 0)
        return MM_ERR_BAD_ARG;
    const long long n = (long long)C * N;
    toplevel_td_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        q_tgt, rewards, dones, behv, alive, td_sum, chunk_valid, N, L, C, NA, gamma);
    return mm::check_launch();
}

extern "C" int mm_vdn_chunk_loss(const float* q_beh, const float* q_tgt, const float* actions, const float* rewards,
                                 const float* dones, const int64_t* row_index, const float* is_weight, float* se, float* dq,
                                 float* td_last, int32_t B, int32_t L, int32_t NA, float gamma, void* stream) {
    if (!q_beh || !q_tgt || !actions || !rewards || !dones || !is_weight || !se || !dq || !td_last || B <= 0 || L <= 0)
        return MM_ERR_BAD_ARG;
    const long long n = (long long)B * L;
    vdn_chunk_loss_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        q_beh, q_tgt, actions, rewards, dones, reinterpret_cast<const long long*>(row_index), is_weight, se, dq, td_last, B, L,
        NA, gamma);
    return mm::check_launch();
}
