// QMIX hypernetwork mixer forward, fused, fp32, for sm_100a.
//
// Replaces QMixer.forward (offpolicy/algorithms/qmix/algorithm/q_mixer.py:68-94, hypernet_layers == 2):
//   w1 = |Linear(HH, A*E)(relu(Linear(SD, HH)(s)))|      viewed [A, E]
//   b1 = Linear(SD, E)(s)
//   w2 = |Linear(HH, E)(relu(Linear(SD, HH)(s)))|
//   b2 = Linear(HH, 1)(relu(Linear(SD, HH)(s)))
//   Q_tot = elu(q[1 x A] @ w1 + b1) @ w2 + b2
// with HH = hypernet_hidden_dim = 64 and E = mixer_hidden_dim = 32 (offpolicy/config.py:158-165).
//
// Persistent CTAs (one per SM) keep all hypernet weights (33 k floats) in shared memory and loop over
// tiles of 64 (t, episode) rows.  The four first layers are one concatenated GEMM [64 x SD] x [SD x 256],
// the second layers two [64 x 64] x [64 x 64] GEMMs, all register-tiled 4x4 per thread like agentq.cu;
// bias / ReLU / |.| run in the accumulators, b2's 64 -> 1 projection is reduced with shuffles straight
// from the accumulators, and the per-row mixing (A x E bilinear form, ELU, E-dot) is one warp per 8 rows.
// The VDN mixer (vdn_mixer.py:28-39) is the A-way sum and needs no kernel of its own (mm_vdn_mix).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int HH = 64;      // hypernet hidden
constexpr int E = 32;       // mixer hidden
constexpr int RT = 64;      // rows per tile
constexpr int LD = RT + 4;
constexpr int NCAT = 256;   // w1a | w2a | b2a | b1 (32) | pad (32)

struct MixLayout {
    int SD, K1, A, AE;
    int wcat, bcat, w1b, b1b, w2b, b2b, wb2, bb2, total;
};

__host__ __device__ inline MixLayout make_mix_layout(int SD, int A) {
    MixLayout L;
    L.SD = SD; L.A = A; L.AE = A * E;
    L.K1 = (SD + 3) & ~3;
    int o = 0;
    auto take = [&](int n) { int at = o; o += (n + 3) & ~3; return at; };
    L.wcat = take(L.K1 * NCAT); L.bcat = take(NCAT);
    L.w1b = take(HH * L.AE); L.b1b = take(L.AE);
    L.w2b = take(HH * 64); L.b2b = take(64);          // E = 32 columns zero-padded to 64
    L.wb2 = take(HH); L.bb2 = take(4);
    L.total = o;
    return L;
}

// acc[g][i][j] += sum_k Xt[k][r0+i] * Wt[k][g*64 + c0+j], Wt row stride `ldw`
template <int NG>
__device__ __forceinline__ void gemm64(float (&acc)[NG][4][4], const float* __restrict__ Xt, int K,
                                       const float* __restrict__ Wt, int ldw, int r0, int c0) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 x = *reinterpret_cast<const float4*>(Xt + k * LD + r0);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float4 w = *reinterpret_cast<const float4*>(Wt + k * ldw + g * 64 + c0);
            const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[g][i][j] = fmaf(xs[i], ws[j], acc[g][i][j]);
        }
    }
}

__device__ __forceinline__ void store_t(float* __restrict__ Yt, const float (&v)[4][4], int r0, int c0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(Yt + (c0 + j) * LD + r0) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
}

template <int A>
__global__ void __launch_bounds__(kThreads, 1)
qmix_mixer_kernel(const float* __restrict__ params, const float* __restrict__ agent_q,
                  const float* __restrict__ states, float* __restrict__ q_tot, long long M, int SD) {
    constexpr int AE = A * E;
    constexpr int NG1 = AE / 64;                       // column blocks of the w1 second layer
    extern __shared__ __align__(16) float smem[];
    const MixLayout L = make_mix_layout(SD, A);
    float* P = smem;
    float* St = P + L.total;                           // [max(K1, AE)][LD]: states, later |w1|
    const int st_rows = L.K1 > AE ? L.K1 : AE;
    float* Ha = St + st_rows * LD;                     // [64][LD]: relu(w1a), later |w2|
    float* Hb = Ha + HH * LD;                          // [64][LD]: relu(w2a)
    float* B1t = Hb + HH * LD;                         // [E][LD]
    float* B2 = B1t + E * LD;                          // [RT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = (tid & 15) * 4, r0 = (tid >> 4) * 4;
    for (int i = tid * 4; i < L.total; i += kThreads * 4)
        *reinterpret_cast<float4*>(P + i) = *reinterpret_cast<const float4*>(params + i);
    __syncthreads();

    const long long tiles = (M + RT - 1) / RT;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row0 = tile * RT;
        const int n_valid = (int)min((long long)RT, M - row0);
        // ---- states tile, transposed ----------------------------------------------------------------
        for (int i = tid; i < RT * L.K1; i += kThreads) {
            const int r = i / L.K1, k = i - r * L.K1;
            St[k * LD + r] = (r < n_valid && k < SD) ? mm::ld_stream(states + (row0 + r) * SD + k) : 0.f;
        }
        __syncthreads();
        // ---- first layers: [64 x K1] x [K1 x 256] -----------------------------------------------------
        {
            float acc[4][4][4];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[g][i][j] = P[L.bcat + g * 64 + c0 + j];
            gemm64<4>(acc, St, L.K1, P + L.wcat, NCAT, r0, c0);
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[g][i][j] = fmaxf(acc[g][i][j], 0.f);
            store_t(Ha, acc[0], r0, c0);
            store_t(Hb, acc[1], r0, c0);
            if (c0 < E) store_t(B1t, acc[3], r0, c0);
            // b2 = relu(b2a) . wb2 + bb2, reduced over the 16 lanes that share these rows
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) s = fmaf(acc[2][i][j], P[L.wb2 + c0 + j], s);
#pragma unroll
                for (int m = 1; m < 16; m <<= 1) s += __shfl_xor_sync(MM_FULL_MASK, s, m);
                if ((tid & 15) == 0) B2[r0 + i] = s + P[L.bb2];
            }
        }
        __syncthreads();
        // ---- second layers ---------------------------------------------------------------------------
        {
            float w1[NG1][4][4], w2[1][4][4];
#pragma unroll
            for (int g = 0; g < NG1; ++g)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) w1[g][i][j] = P[L.b1b + g * 64 + c0 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) w2[0][i][j] = P[L.b2b + c0 + j];
            gemm64<NG1>(w1, Ha, HH, P + L.w1b, AE, r0, c0);
            gemm64<1>(w2, Hb, HH, P + L.w2b, 64, r0, c0);
            __syncthreads();                            // St / Ha are free now: reuse them for |w1| / |w2|
#pragma unroll
            for (int g = 0; g < NG1; ++g) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) w1[g][i][j] = fabsf(w1[g][i][j]);
                store_t(St + g * 64 * LD, w1[g], r0, c0);
            }
            if (c0 < E) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) w2[0][i][j] = fabsf(w2[0][i][j]);
                store_t(Ha, w2[0], r0, c0);
            }
        }
        __syncthreads();
        // ---- mixing: one warp per 8 rows, lane = mixer hidden unit ------------------------------------------
        for (int r = warp * 8; r < warp * 8 + 8; ++r) {
            if (r >= n_valid) break;
            float h = B1t[lane * LD + r];
#pragma unroll
            for (int a = 0; a < A; ++a) h = fmaf(agent_q[(row0 + r) * A + a], St[(a * E + lane) * LD + r], h);
            h = h > 0.f ? h : expm1f(h);               // ELU, alpha = 1
            float s = h * Ha[lane * LD + r];
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(MM_FULL_MASK, s, m);
            if (lane == 0) q_tot[row0 + r] = s + B2[r];
        }
        __syncthreads();
    }
}

__global__ void vdn_mix_kernel(const float* __restrict__ agent_q, float* __restrict__ q_tot, long long M, int A) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += agent_q[i * A + a];     // torch sums left to right over the last dim
    q_tot[i] = s;
}

template <int A>
size_t mixer_smem(int SD) {
    const MixLayout L = make_mix_layout(SD, A);
    const int st_rows = L.K1 > A * E ? L.K1 : A * E;
    return sizeof(float) * (size_t)(L.total + (st_rows + 2 * HH + E) * LD + RT);
}

template <int A>
int launch_mixer(const float* params, const float* agent_q, const float* states, float* q_tot, long long M,
                 int SD, cudaStream_t st) {
    const size_t smem = mixer_smem<A>(SD);
    if (smem > 227 * 1024) return MM_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(qmix_mixer_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    const long long tiles = (M + RT - 1) / RT;
    const unsigned grid = (unsigned)(tiles < 148 ? tiles : 148);
    qmix_mixer_kernel<A><<<grid, kThreads, smem, st>>>(params, agent_q, states, q_tot, M, SD);
    return mm::check_launch();
}

}  // namespace

extern "C" int mm_qmixer_param_count(int32_t SD, int32_t A) {
    if (SD <= 0 || (A != 2 && A != 4)) return MM_ERR_UNSUPPORTED;
    return make_mix_layout(SD, A).total;
}

// offsets[8]: wcat, bcat, w1b, b1b, w2b, b2b, wb2, bb2
extern "C" int mm_qmixer_param_offsets(int32_t SD, int32_t A, int32_t* offsets_host) {
    if (!offsets_host || mm_qmixer_param_count(SD, A) < 0) return MM_ERR_UNSUPPORTED;
    const MixLayout L = make_mix_layout(SD, A);
    const int o[8] = {L.wcat, L.bcat, L.w1b, L.b1b, L.w2b, L.b2b, L.wb2, L.bb2};
    for (int i = 0; i < 8; ++i) offsets_host[i] = o[i];
    return MM_OK;
}

extern "C" int mm_qmixer_forward(const float* params, const float* agent_q, const float* states, float* q_tot,
                                 int64_t M, int32_t SD, int32_t A, void* stream) {
    if (!params || !agent_q || !states || !q_tot || M <= 0) return MM_ERR_BAD_ARG;
    if (!mm::aligned(params, 16)) return MM_ERR_BAD_ARG;
    if (mm_qmixer_param_count(SD, A) < 0) return MM_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (A == 2) return launch_mixer<2>(params, agent_q, states, q_tot, M, SD, st);
    return launch_mixer<4>(params, agent_q, states, q_tot, M, SD, st);
}

extern "C" int mm_vdn_mix(const float* agent_q, float* q_tot, int64_t M, int32_t A, void* stream) {
    if (!agent_q || !q_tot || M <= 0 || A <= 0) return MM_ERR_BAD_ARG;
    vdn_mix_kernel<<<(unsigned)((M + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(agent_q, q_tot, M, A);
    return mm::check_launch();
}
