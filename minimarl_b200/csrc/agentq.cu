// Fused recurrent agent network forward over (env x agent) rows for sm_100a, fp32.
//
// Replaces the eager layer chain of AgentQFunction.forward
// (offpolicy/algorithms/qmix/algorithm/agent_q_function.py:34-66): RNNBase.forward
// (offpolicy/utils/algorithm_utils/rnn.py:33-47) = LayerNorm(D) -> [Linear(D,H) ReLU LayerNorm] ->
// [Linear(H,H) ReLU LayerNorm] (mlp.py:7-29, layer_N = 1) -> nn.GRU(H,H) over the sequence ->
// LayerNorm(H) -> ACTLayer Linear(H,nA) (act.py:21-37).  The same chain with H = 32 is MAPPO's
// actor / critic trunk (mappo/utils/algorithm_utils/mlp.py:6-55, rnn.py:7-80 without mask resets).
//
// One CTA owns a tile of RT consecutive rows and walks the whole sequence t = 0..T1-1 with the hidden
// state resident in shared memory; all weight matrices (37 k floats at H = 64) are staged into shared
// memory once per CTA, so per step the only HBM traffic is the obs tile in and the q tile out.
// Every layer is a register-tiled GEMM (4 rows x 4 cols per thread, activations kept transposed in
// shared memory so both operands are 128-bit loads that broadcast within the warp); bias, ReLU,
// LayerNorm and the GRU gate math run in the accumulator registers, LayerNorm statistics are reduced
// with shuffles across the H/4 lanes that share a row.  fp32 FFMA throughout: this is the
// tolerance-1e-5 reference path on the device; the tensor-core path reuses its layout.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr float kLnEps = 1e-5f;

struct AgentQLayout {   // offsets (in floats) into the packed parameter buffer; see pack_agentq() in Python
    int D, K1, H, NA, NAp;
    int ln0_w, ln0_b, w1, b1, ln1_w, ln1_b, w2, b2, ln2_w, ln2_b, wih, whh, bih, bhh, ln3_w, ln3_b, wq, bq, total;
};

__host__ __device__ inline AgentQLayout make_layout(int D, int H, int NA) {
    AgentQLayout L;
    L.D = D; L.H = H; L.NA = NA;
    L.K1 = (D + 3) & ~3;
    L.NAp = (NA + 3) & ~3;
    int o = 0;
    auto take = [&](int n) { int at = o; o += (n + 3) & ~3; return at; };   // keep every block 16-byte aligned
    L.ln0_w = take(D); L.ln0_b = take(D);
    L.w1 = take(L.K1 * H); L.b1 = take(H); L.ln1_w = take(H); L.ln1_b = take(H);
    L.w2 = take(H * H); L.b2 = take(H); L.ln2_w = take(H); L.ln2_b = take(H);
    L.wih = take(H * 3 * H); L.whh = take(H * 3 * H); L.bih = take(3 * H); L.bhh = take(3 * H);
    L.ln3_w = take(H); L.ln3_b = take(H);
    L.wq = take(H * L.NAp); L.bq = take(L.NAp);
    L.total = o;
    return L;
}

template <int H>
struct Tile {
    static constexpr int CG = H / 4;              // column groups (4 cols each) = lanes sharing a row
    static constexpr int RG = kThreads / CG;      // row groups (4 rows each)
    static constexpr int RT = RG * 4;             // rows per CTA: 64 at H = 64, 128 at H = 32
    static constexpr int LD = RT + 4;             // padded leading dimension of the transposed tiles
};

// acc[g][i][j] += sum_k Xt[k][r0+i] * Wt[k][g*H + c0+j]   (NG column blocks of width H)
template <int H, int NG>
__device__ __forceinline__ void gemm_acc(float (&acc)[NG][4][4], const float* __restrict__ Xt, int K,
                                         const float* __restrict__ Wt, int r0, int c0) {
    constexpr int LD = Tile<H>::LD;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(Xt + k * LD + r0);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float4 w = *reinterpret_cast<const float4*>(Wt + k * (NG * H) + g * H + c0);
            const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[g][i][j] = fmaf(xs[i], ws[j], acc[g][i][j]);
        }
    }
}

// LayerNorm over the H columns of each of this thread's 4 rows (the row is spread over CG lanes).
template <int H>
__device__ __forceinline__ void layer_norm_rows(float (&v)[4][4], const float* __restrict__ w,
                                                const float* __restrict__ b, int c0) {
    constexpr int CG = Tile<H>::CG;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s = v[i][0] + v[i][1] + v[i][2] + v[i][3];
#pragma unroll
        for (int m = 1; m < CG; m <<= 1) s += __shfl_xor_sync(MM_FULL_MASK, s, m);
        const float mean = s * (1.0f / H);
        float d[4], q = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { d[j] = v[i][j] - mean; q = fmaf(d[j], d[j], q); }
#pragma unroll
        for (int m = 1; m < CG; m <<= 1) q += __shfl_xor_sync(MM_FULL_MASK, q, m);
        const float rstd = rsqrtf(q * (1.0f / H) + kLnEps);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] = fmaf(d[j] * rstd, w[c0 + j], b[c0 + j]);
    }
}

template <int H>
__device__ __forceinline__ void store_transposed(float* __restrict__ Yt, const float (&v)[4][4], int r0, int c0) {
    constexpr int LD = Tile<H>::LD;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(Yt + (c0 + j) * LD + r0) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int H>
__global__ void __launch_bounds__(kThreads, 1)
agentq_forward_kernel(const float* __restrict__ params, const float* __restrict__ obs,
                      const float* __restrict__ h0, float* __restrict__ q_out, float* __restrict__ h_final,
                      float* __restrict__ rnn_out, int T1, long long R, int D, int NA) {
    using TL = Tile<H>;
    constexpr int RT = TL::RT, LD = TL::LD, CG = TL::CG;
    extern __shared__ __align__(16) float smem[];
    const AgentQLayout L = make_layout(D, H, NA);
    float* P = smem;                                    // packed parameters
    const int xa_rows = L.K1 > H ? L.K1 : H;
    float* Xa = P + L.total;                            // [xa_rows][LD]
    float* Xb = Xa + xa_rows * LD;                      // [H][LD]
    float* Ht = Xb + H * LD;                            // [H][LD] hidden state, transposed

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row0 = (long long)blockIdx.x * RT;
    const int n_valid = (int)min((long long)RT, R - row0);
    const int c0 = (tid % CG) * 4, r0 = (tid / CG) * 4;

    for (int i = tid * 4; i < L.total; i += kThreads * 4)
        *reinterpret_cast<float4*>(P + i) = *reinterpret_cast<const float4*>(params + i);
    for (int i = tid; i < H * RT; i += kThreads) {      // Ht[j][r] = h0[row0 + r][j]
        const int r = i / H, j = i - r * H;
        Ht[j * LD + r] = (h0 != nullptr && r < n_valid) ? h0[(row0 + r) * H + j] : 0.f;
    }
    __syncthreads();

    for (int t = 0; t < T1; ++t) {
        // ---- A. obs tile -> LayerNorm(D) -> Xa (transposed); one warp per row --------------------------
        const float* obs_t = obs + ((long long)t * R + row0) * D;
        for (int r = warp; r < RT; r += kThreads / 32) {
            float x[3] = {0.f, 0.f, 0.f};
            if (r < n_valid) {
#pragma unroll
                for (int u = 0; u < 3; ++u)
                    if (lane + 32 * u < D) x[u] = mm::ld_stream(obs_t + (long long)r * D + lane + 32 * u);
            }
            float s = x[0] + x[1] + x[2];
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(MM_FULL_MASK, s, m);
            const float mean = s / (float)D;
            float qv = 0.f, dlt[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                dlt[u] = (lane + 32 * u < D) ? x[u] - mean : 0.f;
                qv = fmaf(dlt[u], dlt[u], qv);
            }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) qv += __shfl_xor_sync(MM_FULL_MASK, qv, m);
            const float rstd = rsqrtf(qv / (float)D + kLnEps);
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int k = lane + 32 * u;
                if (k < D) Xa[k * LD + r] = fmaf(dlt[u] * rstd, P[L.ln0_w + k], P[L.ln0_b + k]);
                else if (k < L.K1) Xa[k * LD + r] = 0.f;
            }
        }
        __syncthreads();

        // ---- B. fc1: Linear(D,H) + ReLU + LayerNorm -> Xb -------------------------------------------
        {
            float acc[1][4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[0][i][j] = P[L.b1 + c0 + j];
            gemm_acc<H, 1>(acc, Xa, L.K1, P + L.w1, r0, c0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[0][i][j] = fmaxf(acc[0][i][j], 0.f);
            layer_norm_rows<H>(acc[0], P + L.ln1_w, P + L.ln1_b, c0);
            store_transposed<H>(Xb, acc[0], r0, c0);
        }
        __syncthreads();

        // ---- C. fc2: Linear(H,H) + ReLU + LayerNorm -> Xa -------------------------------------------
        {
            float acc[1][4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[0][i][j] = P[L.b2 + c0 + j];
            gemm_acc<H, 1>(acc, Xb, H, P + L.w2, r0, c0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[0][i][j] = fmaxf(acc[0][i][j], 0.f);
            layer_norm_rows<H>(acc[0], P + L.ln2_w, P + L.ln2_b, c0);
            store_transposed<H>(Xa, acc[0], r0, c0);
        }
        __syncthreads();

        // ---- D. GRU cell (PyTorch gate order r, z, n) -> Ht, LayerNorm -> Xb ----------------------------
        {
            float gi[3][4][4], gh[3][4][4];
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        gi[g][i][j] = P[L.bih + g * H + c0 + j];
                        gh[g][i][j] = P[L.bhh + g * H + c0 + j];
                    }
            gemm_acc<H, 3>(gi, Xa, H, P + L.wih, r0, c0);
            gemm_acc<H, 3>(gh, Ht, H, P + L.whh, r0, c0);
            float hn[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 ho = *reinterpret_cast<const float4*>(Ht + (c0 + j) * LD + r0);
                const float hold[4] = {ho.x, ho.y, ho.z, ho.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float rr = sigmoidf_(gi[0][i][j] + gh[0][i][j]);
                    const float zz = sigmoidf_(gi[1][i][j] + gh[1][i][j]);
                    const float nn = tanhf(gi[2][i][j] + rr * gh[2][i][j]);
                    hn[i][j] = (1.0f - zz) * nn + zz * hold[i];
                }
            }
            __syncthreads();                      // every thread is done reading the old hidden tile
            store_transposed<H>(Ht, hn, r0, c0);
            layer_norm_rows<H>(hn, P + L.ln3_w, P + L.ln3_b, c0);
            store_transposed<H>(Xb, hn, r0, c0);
            if (rnn_out != nullptr) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (r0 + i < n_valid)
                        *reinterpret_cast<float4*>(rnn_out + ((long long)t * R + row0 + r0 + i) * H + c0) =
                            make_float4(hn[i][0], hn[i][1], hn[i][2], hn[i][3]);
            }
        }
        __syncthreads();

        // ---- E. head: Linear(H, NA) -> q[t, row, :] ---------------------------------------------------
        if (q_out != nullptr) {
            float* q_t = q_out + ((long long)t * R + row0) * NA;
            for (int o = tid; o < n_valid * NA; o += kThreads) {
                const int r = o / NA, a = o - r * NA;
                float acc = P[L.bq + a];
#pragma unroll 8
                for (int k = 0; k < H; ++k) acc = fmaf(Xb[k * LD + r], P[L.wq + k * L.NAp + a], acc);
                q_t[o] = acc;
            }
        }
        // the next write to Xb (step B of t+1) is behind the barrier after step A, so no barrier here
    }

    if (h_final != nullptr) {
        __syncthreads();
        for (int i = tid; i < H * RT; i += kThreads) {
            const int r = i / H, j = i - r * H;
            if (r < n_valid) h_final[(row0 + r) * H + j] = Ht[j * LD + r];
        }
    }
}

template <int H>
size_t smem_bytes(int D, int NA) {
    const AgentQLayout L = make_layout(D, H, NA);
    const int xa_rows = L.K1 > H ? L.K1 : H;
    return sizeof(float) * (size_t)(L.total + (xa_rows + 2 * H) * Tile<H>::LD);
}

template <int H>
int launch(const float* params, const float* obs, const float* h0, float* q, float* h_final, float* rnn_out,
           int T1, long long R, int D, int NA, cudaStream_t st) {
    const size_t smem = smem_bytes<H>(D, NA);
    if (smem > 227 * 1024) return MM_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(agentq_forward_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    const long long tiles = (R + Tile<H>::RT - 1) / Tile<H>::RT;
    agentq_forward_kernel<H><<<(unsigned)tiles, kThreads, smem, st>>>(params, obs, h0, q, h_final, rnn_out, T1, R, D, NA);
    return mm::check_launch();
}

}  // namespace

extern "C" int mm_agentq_param_count(int32_t D, int32_t H, int32_t NA) {
    if (D <= 0 || D > 96 || (H != 32 && H != 64) || NA <= 0) return MM_ERR_UNSUPPORTED;
    return make_layout(D, H, NA).total;
}

// offsets[18]: ln0_w, ln0_b, w1, b1, ln1_w, ln1_b, w2, b2, ln2_w, ln2_b, wih, whh, bih, bhh, ln3_w, ln3_b, wq, bq
extern "C" int mm_agentq_param_offsets(int32_t D, int32_t H, int32_t NA, int32_t* offsets_host) {
    if (!offsets_host || mm_agentq_param_count(D, H, NA) < 0) return MM_ERR_UNSUPPORTED;
    const AgentQLayout L = make_layout(D, H, NA);
    const int o[18] = {L.ln0_w, L.ln0_b, L.w1, L.b1, L.ln1_w, L.ln1_b, L.w2, L.b2, L.ln2_w, L.ln2_b,
                       L.wih, L.whh, L.bih, L.bhh, L.ln3_w, L.ln3_b, L.wq, L.bq};
    for (int i = 0; i < 18; ++i) offsets_host[i] = o[i];
    return MM_OK;
}

extern "C" int mm_agentq_forward(const float* params, const float* obs, const float* h0, float* q,
                                 float* h_final, float* rnn_out, int32_t T1, int64_t R, int32_t D, int32_t H,
                                 int32_t NA, void* stream) {
    if (!params || !obs || T1 <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (!mm::aligned(params, 16)) return MM_ERR_BAD_ARG;
    if (mm_agentq_param_count(D, H, NA) < 0) return MM_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (H == 64) return launch<64>(params, obs, h0, q, h_final, rnn_out, T1, R, D, NA, st);
    return launch<32>(params, obs, h0, q, h_final, rnn_out, T1, R, D, NA, st);
}
