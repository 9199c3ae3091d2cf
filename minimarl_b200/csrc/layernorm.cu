// LayerNorm forward / backward over short rows (N <= 128) for the differentiated network pass.
//
// The agent / actor / critic chains apply nn.LayerNorm to [T*rows, N] activations with N in {47, 32, 64}
// (offpolicy/utils/algorithm_utils/mlp.py:17-29, rnn.py:18,28; mappo/utils/algorithm_utils/mlp.py:17-30).
// PyTorch's LayerNorm launches one CTA per row; at 827 k rows of 64 floats that is ~1.5 ms per call
// (0.3 TB/s) and the N = 47 case falls off its vectorised path (4 ms).  Here one warp owns one row per
// iteration of a grid-stride loop (coalesced 4-byte accesses, statistics by shuffles), persistent grid, HBM-bound:
// forward 8 B/elem (+8 B/row of saved statistics), backward 12 B/elem.  dgamma / dbeta are accumulated per
// lane in registers over all rows of the warp, reduced once per CTA and added atomically to fp32 totals.
#include "common.cuh"

namespace {

constexpr int kMaxPerLane = 4;      // N <= 128
constexpr float kEps = 1e-5f;

__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
              float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long M, int N) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    float wv[kMaxPerLane], bv[kMaxPerLane];
#pragma unroll
    for (int u = 0; u < kMaxPerLane; ++u) {
        const int k = lane + 32 * u;
        wv[u] = k < N ? w[k] : 0.f;
        bv[u] = k < N ? b[k] : 0.f;
    }
    const float inv_n = 1.0f / (float)N;
    for (long long r = warp0; r < M; r += n_warps) {
        const float* xr = x + r * N;
        float v[kMaxPerLane], s = 0.f;
#pragma unroll
        for (int u = 0; u < kMaxPerLane; ++u) {
            const int k = lane + 32 * u;
            v[u] = k < N ? xr[k] : 0.f;
            s += v[u];
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(MM_FULL_MASK, s, m);
        const float mean = s * inv_n;
        float q = 0.f;
#pragma unroll
        for (int u = 0; u < kMaxPerLane; ++u) {
            const int k = lane + 32 * u;
            const float d = k < N ? v[u] - mean : 0.f;
            q = fmaf(d, d, q);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) q += __shfl_xor_sync(MM_FULL_MASK, q, m);
        const float rstd = rsqrtf(q * inv_n + kEps);
        float* yr = y + r * N;
#pragma unroll
        for (int u = 0; u < kMaxPerLane; ++u) {
            const int k = lane + 32 * u;
            if (k < N) yr[k] = fmaf((v[u] - mean) * rstd, wv[u], bv[u]);
        }
        if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
    }
}

__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dx,
              float* __restrict__ dw, float* __restrict__ db, long long M, int N) {
    __shared__ float s_dw[8][128], s_db[8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    float wv[kMaxPerLane], aw[kMaxPerLane], ab[kMaxPerLane];
#pragma unroll
    for (int u = 0; u < kMaxPerLane; ++u) {
        const int k = lane + 32 * u;
        wv[u] = k < N ? w[k] : 0.f;
        aw[u] = ab[u] = 0.f;
    }
    const float inv_n = 1.0f / (float)N;
    for (long long r = warp0; r < M; r += n_warps) {
        const float mean = mean_in[r], rstd = rstd_in[r];
        float g[kMaxPerLane], xh[kMaxPerLane], c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int u = 0; u < kMaxPerLane; ++u) {
            const int k = lane + 32 * u;
            const float d = k < N ? dy[r * N + k] : 0.f;
            xh[u] = k < N ? (x[r * N + k] - mean) * rstd : 0.f;
            ab[u] += d;
            aw[u] = fmaf(d, xh[u], aw[u]);
            g[u] = d * wv[u];
            c1 += g[u];
            c2 = fmaf(g[u], xh[u], c2);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            c1 += __shfl_xor_sync(MM_FULL_MASK, c1, m);
            c2 += __shfl_xor_sync(MM_FULL_MASK, c2, m);
        }
        c1 *= inv_n; c2 *= inv_n;
#pragma unroll
        for (int u = 0; u < kMaxPerLane; ++u) {
            const int k = lane + 32 * u;
            if (k < N) dx[r * N + k] = rstd * (g[u] - c1 - xh[u] * c2);
        }
    }
#pragma unroll
    for (int u = 0; u < kMaxPerLane; ++u) { s_dw[warp][lane + 32 * u] = aw[u]; s_db[warp][lane + 32 * u] = ab[u]; }
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        float tw = 0.f, tb = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { tw += s_dw[i][k]; tb += s_db[i][k]; }
        atomicAdd(dw + k, tw);
        atomicAdd(db + k, tb);
    }
}

// Vector path for N % 4 == 0 (N = 64, 32, 128...): LPR = N / 4 lanes own one row as float4s, so a warp covers 32 / LPR rows
// per access and UN such groups per loop iteration: 4 x more bytes in flight per warp than the row-per-warp loop above
// (which left the 8 warps of a CTA at ~16 KB per SM in flight -- about half the bandwidth-delay product).
template <int LPR, int UN>
__global__ void __launch_bounds__(256)
ln_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                  float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long long M) {
    constexpr int N = 4 * LPR, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, c = lane % LPR;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float4 wv = *reinterpret_cast<const float4*>(w + 4 * c), bv = *reinterpret_cast<const float4*>(b + 4 * c);
    constexpr float inv_n = 1.0f / (float)N;
    for (long long r0 = warp0 * (RPW * UN); r0 < M; r0 += n_warps * (RPW * UN)) {
        float4 v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long r = r0 + u * RPW + sub;
            v[u] = r < M ? *reinterpret_cast<const float4*>(x + r * N + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long r = r0 + u * RPW + sub;
            float s = (v[u].x + v[u].y) + (v[u].z + v[u].w);
#pragma unroll
            for (int m = LPR / 2; m > 0; m >>= 1) s += __shfl_xor_sync(MM_FULL_MASK, s, m);
            const float mean = s * inv_n;
            const float d0 = v[u].x - mean, d1 = v[u].y - mean, d2 = v[u].z - mean, d3 = v[u].w - mean;
            float q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
#pragma unroll
            for (int m = LPR / 2; m > 0; m >>= 1) q += __shfl_xor_sync(MM_FULL_MASK, q, m);
            const float rstd = rsqrtf(q * inv_n + kEps);
            if (r < M) {
                *reinterpret_cast<float4*>(y + r * N + 4 * c) = make_float4(fmaf(d0 * rstd, wv.x, bv.x), fmaf(d1 * rstd, wv.y, bv.y),
                                                                            fmaf(d2 * rstd, wv.z, bv.z), fmaf(d3 * rstd, wv.w, bv.w));
                if (c == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
            }
        }
    }
}

template <int LPR, int UN>
__global__ void __launch_bounds__(256)
ln_bwd_vec_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dx,
                  float* __restrict__ dw, float* __restrict__ db, long long M) {
    constexpr int N = 4 * LPR, RPW = 32 / LPR;
    __shared__ float s_dw[8][32 * 4], s_db[8][32 * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / LPR, c = lane % LPR;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float4 wv = *reinterpret_cast<const float4*>(w + 4 * c);
    float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr float inv_n = 1.0f / (float)N;
    for (long long r0 = warp0 * (RPW * UN); r0 < M; r0 += n_warps * (RPW * UN)) {
        float4 g4[UN], x4[UN];
        float mu[UN], rs[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long r = r0 + u * RPW + sub;
            const bool ok = r < M;
            g4[u] = ok ? *reinterpret_cast<const float4*>(dy + r * N + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            x4[u] = ok ? *reinterpret_cast<const float4*>(x + r * N + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            mu[u] = ok ? mean_in[r] : 0.f;
            rs[u] = ok ? rstd_in[r] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long r = r0 + u * RPW + sub;
            const float d[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
            const float xh[4] = {(x4[u].x - mu[u]) * rs[u], (x4[u].y - mu[u]) * rs[u], (x4[u].z - mu[u]) * rs[u],
                                 (x4[u].w - mu[u]) * rs[u]};
            const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
            float g[4], c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ab[j] += d[j];
                aw[j] = fmaf(d[j], xh[j], aw[j]);
                g[j] = d[j] * ww[j];
                c1 += g[j];
                c2 = fmaf(g[j], xh[j], c2);
            }
#pragma unroll
            for (int m = LPR / 2; m > 0; m >>= 1) {
                c1 += __shfl_xor_sync(MM_FULL_MASK, c1, m);
                c2 += __shfl_xor_sync(MM_FULL_MASK, c2, m);
            }
            c1 *= inv_n; c2 *= inv_n;
            if (r < M)
                *reinterpret_cast<float4*>(dx + r * N + 4 * c) =
                    make_float4(rs[u] * (g[0] - c1 - xh[0] * c2), rs[u] * (g[1] - c1 - xh[1] * c2),
                                rs[u] * (g[2] - c1 - xh[2] * c2), rs[u] * (g[3] - c1 - xh[3] * c2));
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_dw[warp][lane * 4 + j] = aw[j]; s_db[warp][lane * 4 + j] = ab[j]; }
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {      // column k lives in lane slot (sub, k / 4), element k % 4
        float tw = 0.f, tb = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int sb = 0; sb < RPW; ++sb) {
                tw += s_dw[i][(sb * LPR + k / 4) * 4 + (k & 3)];
                tb += s_db[i][(sb * LPR + k / 4) * 4 + (k & 3)];
            }
        atomicAdd(dw + k, tw);
        atomicAdd(db + k, tb);
    }
}

inline unsigned ln_grid(long long M) {
    const long long want = (M + 7) / 8;                 // 8 warps per CTA
    const long long cap = 148 * 8;                      // persistent: 8 CTAs of 256 threads per SM
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace

extern "C" int mm_layernorm_fwd(const float* x, const float* weight, const float* bias, float* y, float* mean,
                                float* rstd, int64_t M, int32_t N, void* stream) {
    if (!x || !weight || !bias || !y || !mean || !rstd || M <= 0 || N <= 0) return MM_ERR_BAD_ARG;
    if (N > 32 * kMaxPerLane) return MM_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = mm::aligned(x, 16) && mm::aligned(y, 16) && mm::aligned(weight, 16) && mm::aligned(bias, 16);
    if (vec && N == 64)       ln_fwd_vec_kernel<16, 4><<<ln_grid((M + 7) / 8), 256, 0, st>>>(x, weight, bias, y, mean, rstd, M);
    else if (vec && N == 32)  ln_fwd_vec_kernel<8, 4><<<ln_grid((M + 15) / 16), 256, 0, st>>>(x, weight, bias, y, mean, rstd, M);
    else if (vec && N == 128) ln_fwd_vec_kernel<32, 4><<<ln_grid((M + 3) / 4), 256, 0, st>>>(x, weight, bias, y, mean, rstd, M);
    else ln_fwd_kernel<<<ln_grid(M), 256, 0, st>>>(x, weight, bias, y, mean, rstd, M, N);
    return mm::check_launch();
}

extern "C" int mm_layernorm_bwd(const float* dy, const float* x, const float* weight, const float* mean,
                                const float* rstd, float* dx, float* dweight_accum, float* dbias_accum, int64_t M,
                                int32_t N, void* stream) {
    if (!dy || !x || !weight || !mean || !rstd || !dx || !dweight_accum || !dbias_accum || M <= 0 || N <= 0) return MM_ERR_BAD_ARG;
    if (N > 32 * kMaxPerLane) return MM_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = mm::aligned(dy, 16) && mm::aligned(x, 16) && mm::aligned(dx, 16) && mm::aligned(weight, 16);
    if (vec && N == 64)
        ln_bwd_vec_kernel<16, 4><<<ln_grid((M + 7) / 8), 256, 0, st>>>(dy, x, weight, mean, rstd, dx, dweight_accum, dbias_accum, M);
    else if (vec && N == 32)
        ln_bwd_vec_kernel<8, 4><<<ln_grid((M + 15) / 16), 256, 0, st>>>(dy, x, weight, mean, rstd, dx, dweight_accum, dbias_accum, M);
    else if (vec && N == 128)
        ln_bwd_vec_kernel<32, 4><<<ln_grid((M + 3) / 4), 256, 0, st>>>(dy, x, weight, mean, rstd, dx, dweight_accum, dbias_accum, M);
    else
        ln_bwd_kernel<<<ln_grid(M), 256, 0, st>>>(dy, x, weight, mean, rstd, dx, dweight_accum, dbias_accum, M, N);
    return mm::check_launch();
}
