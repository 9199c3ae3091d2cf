// LayerNorm(H) followed by the Q head Linear(H, NA) of the recurrent agent network, forward and backward as one kernel each.
//
// Reference: RNNLayer.norm (offpolicy/utils/algorithm_utils/rnn.py:18,45) and ACTLayer.action_out
// (offpolicy/utils/algorithm_utils/act.py) as AgentQFunction.forward applies them to the GRU outputs
// (offpolicy/algorithms/qmix/algorithm/agent_q_function.py:60-66), differentiated by QMix.train_policy_on_batch.  With NA = 5
// outputs the Linear is a bandwidth problem, not a GEMM: autograd ran LayerNorm forward / backward kernels, three padded cuBLAS
// GEMMs (forward, dX, dW) and a bias column sum over 827 392 rows (~0.4 ms); here a row is read once per direction:
//   forward : h [rows, H] -> q [rows, NA]  (+ mean, 1 / sigma per row)                           264 + 20 + 8 B per row
//   backward: dq, h, mean, 1 / sigma -> dh [rows, H]; dW_q, db_q, d gamma, d beta accumulated     284 + 256 B per row
// LPR = H / 4 lanes own a row (one float4 each), a warp covers 32 / LPR rows per access and UN such groups per iteration; row
// statistics and the NA dot products are butterfly reductions inside the LPR-lane group.  Parameter gradients are accumulated per
// lane in registers over all rows of the warp, reduced once per CTA through shared memory and added atomically to fp32 totals.
#include "common.cuh"

namespace {

constexpr int kNaLimit = 8;
constexpr float kEps = 1e-5f;

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int m = LPR / 2; m > 0; m >>= 1) v += __shfl_xor_sync(MM_FULL_MASK, v, m);
    return v;
}

template <int LPR, int UN, int kMaxNA>
__global__ void __launch_bounds__(256, 4)
ln_head_fwd_kernel(const float* __restrict__ h, const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ wq,
                   const float* __restrict__ bq, float* __restrict__ q, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                   long long M, int NA) {
    constexpr int N = 4 * LPR, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, c = lane % LPR;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float4 gv = *reinterpret_cast<const float4*>(g + 4 * c), bv = *reinterpret_cast<const float4*>(b + 4 * c);
    float4 w[kMaxNA];
#pragma unroll
    for (int a = 0; a < kMaxNA; ++a) w[a] = a < NA ? *reinterpret_cast<const float4*>(wq + a * N + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float my_bias = c < NA ? bq[c] : 0.f;
    constexpr float inv_n = 1.0f / (float)N;
    for (long long r0 = warp0 * (RPW * UN); r0 < M; r0 += n_warps * (RPW * UN)) {
        float4 v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long r = r0 + u * RPW + sub;
            v[u] = r < M ? *reinterpret_cast<const float4*>(h + r * N + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long r = r0 + u * RPW + sub;
            const float mean = group_sum<LPR>((v[u].x + v[u].y) + (v[u].z + v[u].w)) * inv_n;
            const float d0 = v[u].x - mean, d1 = v[u].y - mean, d2 = v[u].z - mean, d3 = v[u].w - mean;
            const float rstd = rsqrtf(group_sum<LPR>(fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)))) * inv_n + kEps);
            const float y0 = fmaf(d0 * rstd, gv.x, bv.x), y1 = fmaf(d1 * rstd, gv.y, bv.y), y2 = fmaf(d2 * rstd, gv.z, bv.z),
                        y3 = fmaf(d3 * rstd, gv.w, bv.w);
            float mine = 0.f;                         // lane c < NA keeps output c of its row
#pragma unroll
            for (int a = 0; a < kMaxNA; ++a) {
                if (a < NA) {
                    const float s = group_sum<LPR>(fmaf(y0, w[a].x, fmaf(y1, w[a].y, fmaf(y2, w[a].z, y3 * w[a].w))));
                    if (c == a) mine = s;
                }
            }
            if (r < M) {
                if (c < NA) q[r * NA + c] = mine + my_bias;
                if (c == LPR - 1) { mean_out[r] = mean; rstd_out[r] = rstd; }
            }
        }
    }
}

template <int LPR, int UN, int kMaxNA>
__global__ void __launch_bounds__(256, 2)
ln_head_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ h, const float* __restrict__ mean_in,
                   const float* __restrict__ rstd_in, const float* __restrict__ g, const float* __restrict__ b,
                   const float* __restrict__ wq, float* __restrict__ dh, float* __restrict__ dwq, float* __restrict__ dbq,
                   float* __restrict__ dg, float* __restrict__ db, long long M, int NA) {
    constexpr int N = 4 * LPR, RPW = 32 / LPR;
    __shared__ float s_acc[8][kMaxNA * 4 + 8 + 1][33];        // per warp: the lane's accumulators, padded against bank conflicts
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / LPR, c = lane % LPR;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float4 gv = *reinterpret_cast<const float4*>(g + 4 * c), bv = *reinterpret_cast<const float4*>(b + 4 * c);
    float4 w[kMaxNA];
#pragma unroll
    for (int a = 0; a < kMaxNA; ++a) w[a] = a < NA ? *reinterpret_cast<const float4*>(wq + a * N + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float aw[kMaxNA][4], ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f}, abq = 0.f;
#pragma unroll
    for (int a = 0; a < kMaxNA; ++a) aw[a][0] = aw[a][1] = aw[a][2] = aw[a][3] = 0.f;
    constexpr float inv_n = 1.0f / (float)N;
    for (long long r0 = warp0 * (RPW * UN); r0 < M; r0 += n_warps * (RPW * UN)) {
        float4 x4[UN];
        float mu[UN], rs[UN], gq[UN][kMaxNA];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long r = r0 + u * RPW + sub;
            const bool ok = r < M;
            x4[u] = ok ? *reinterpret_cast<const float4*>(h + r * N + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            mu[u] = ok ? mean_in[r] : 0.f;
            rs[u] = ok ? rstd_in[r] : 0.f;
#pragma unroll
            for (int a = 0; a < kMaxNA; ++a) gq[u][a] = (ok && a < NA) ? dq[r * NA + a] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long r = r0 + u * RPW + sub;
            const float xh[4] = {(x4[u].x - mu[u]) * rs[u], (x4[u].y - mu[u]) * rs[u], (x4[u].z - mu[u]) * rs[u], (x4[u].w - mu[u]) * rs[u]};
            const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
            float dy[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int a = 0; a < kMaxNA; ++a) {
                const float ga = gq[u][a];
                const float wa[4] = {w[a].x, w[a].y, w[a].z, w[a].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dy[j] = fmaf(ga, wa[j], dy[j]);
                    aw[a][j] = fmaf(ga, fmaf(xh[j], gg[j], bb[j]), aw[a][j]);       // dW_q[a][k] += dq_a y_k
                }
                if (c == a) abq += ga;                                               // lane a of the group sums db_q[a]
            }
            float dxh[4], c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ab[j] += dy[j];
                ag[j] = fmaf(dy[j], xh[j], ag[j]);
                dxh[j] = dy[j] * gg[j];
                c1 += dxh[j];
                c2 = fmaf(dxh[j], xh[j], c2);
            }
            c1 = group_sum<LPR>(c1) * inv_n;
            c2 = group_sum<LPR>(c2) * inv_n;
            if (r < M)
                *reinterpret_cast<float4*>(dh + r * N + 4 * c) =
                    make_float4(rs[u] * (dxh[0] - c1 - xh[0] * c2), rs[u] * (dxh[1] - c1 - xh[1] * c2),
                                rs[u] * (dxh[2] - c1 - xh[2] * c2), rs[u] * (dxh[3] - c1 - xh[3] * c2));
        }
    }
    // ---- CTA reduction: 8 warps x RPW row groups share the same columns ---------------------------------------------------------
    constexpr int kVals = kMaxNA * 4 + 8 + 1;
#pragma unroll
    for (int a = 0; a < kMaxNA; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j) s_acc[warp][a * 4 + j][lane] = aw[a][j];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s_acc[warp][kMaxNA * 4 + j][lane] = ag[j];
        s_acc[warp][kMaxNA * 4 + 4 + j][lane] = ab[j];
    }
    s_acc[warp][kMaxNA * 4 + 8][lane] = abq;
    __syncthreads();
    for (int i = threadIdx.x; i < kVals * LPR; i += blockDim.x) {
        const int val = i / LPR, cc = i - val * LPR;
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv)
#pragma unroll
            for (int sb = 0; sb < RPW; ++sb) s += s_acc[wv][val][sb * LPR + cc];
        if (val < kMaxNA * 4) {
            const int a = val >> 2, j = val & 3;
            if (a < NA) atomicAdd(dwq + a * N + 4 * cc + j, s);
        } else if (val < kMaxNA * 4 + 4) {
            atomicAdd(dg + 4 * cc + (val - kMaxNA * 4), s);
        } else if (val < kMaxNA * 4 + 8) {
            atomicAdd(db + 4 * cc + (val - kMaxNA * 4 - 4), s);
        } else if (cc < NA) {
            atomicAdd(dbq + cc, s);
        }
    }
}

unsigned ln_grid(long long warps_needed) {
    long long blocks = (warps_needed + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" int mm_ln_head_fwd(const float* h, const float* ln_w, const float* ln_b, const float* wq, const float* bq, float* q,
                              float* mean, float* rstd, int64_t M, int32_t H, int32_t NA, void* stream) {
    if (!h || !ln_w || !ln_b || !wq || !bq || !q || !mean || !rstd || M <= 0) return MM_ERR_BAD_ARG;
    if ((H != 64 && H != 32) || NA < 1 || NA > kNaLimit) return MM_ERR_UNSUPPORTED;
    if (!mm::aligned(h, 16) || !mm::aligned(ln_w, 16) || !mm::aligned(ln_b, 16) || !mm::aligned(wq, 16)) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // (accumulator / weight registers are sized by the head width: 5 actions in both ma-gym envs, 8 is the general case)
    if (H == 64 && NA <= 5)      ln_head_fwd_kernel<16, 2, 5><<<ln_grid((M + 3) / 4), 256, 0, st>>>(h, ln_w, ln_b, wq, bq, q, mean, rstd, M, NA);
    else if (H == 64)            ln_head_fwd_kernel<16, 2, 8><<<ln_grid((M + 3) / 4), 256, 0, st>>>(h, ln_w, ln_b, wq, bq, q, mean, rstd, M, NA);
    else if (NA <= 5)            ln_head_fwd_kernel<8, 2, 5><<<ln_grid((M + 7) / 8), 256, 0, st>>>(h, ln_w, ln_b, wq, bq, q, mean, rstd, M, NA);
    else                         ln_head_fwd_kernel<8, 2, 8><<<ln_grid((M + 7) / 8), 256, 0, st>>>(h, ln_w, ln_b, wq, bq, q, mean, rstd, M, NA);
    return mm::check_launch();
}

// dwq [NA, H], dbq [NA], dln_w [H], dln_b [H] are ACCUMULATED into (the caller zeroes them)
extern "C" int mm_ln_head_bwd(const float* dq, const float* h, const float* mean, const float* rstd, const float* ln_w,
                              const float* ln_b, const float* wq, float* dh, float* dwq, float* dbq, float* dln_w, float* dln_b,
                              int64_t M, int32_t H, int32_t NA, void* stream) {
    if (!dq || !h || !mean || !rstd || !ln_w || !ln_b || !wq || !dh || !dwq || !dbq || !dln_w || !dln_b || M <= 0) return MM_ERR_BAD_ARG;
    if ((H != 64 && H != 32) || NA < 1 || NA > kNaLimit) return MM_ERR_UNSUPPORTED;
    if (!mm::aligned(h, 16) || !mm::aligned(dh, 16) || !mm::aligned(ln_w, 16) || !mm::aligned(ln_b, 16) || !mm::aligned(wq, 16))
        return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MM_LN_HEAD_BWD(LPR, MAXNA, ROWS) \
    ln_head_bwd_kernel<LPR, 2, MAXNA><<<ln_grid((M + ROWS - 1) / ROWS), 256, 0, st>>>(dq, h, mean, rstd, ln_w, ln_b, wq, dh, dwq, dbq, dln_w, dln_b, M, NA)
    if (H == 64 && NA <= 5) MM_LN_HEAD_BWD(16, 5, 4);
    else if (H == 64)       MM_LN_HEAD_BWD(16, 8, 4);
    else if (NA <= 5)       MM_LN_HEAD_BWD(8, 5, 8);
    else                    MM_LN_HEAD_BWD(8, 8, 8);
#undef MM_LN_HEAD_BWD
    return mm::check_launch();
}
