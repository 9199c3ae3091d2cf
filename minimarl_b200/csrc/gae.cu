// GAE / discounted-return reverse scan over T for sm_100a.
//
// Replaces SharedReplayBuffer.compute_returns (mappo/runner/shared/shared_buffer.py:131-157) with
// ValueNorm.denormalize (mappo/utils/valuenorm.py:67-78) folded in, and the advantage line of
// R_MAPPO.train (mappo/algorithms/ramppo_network.py:225-230).
//
// One thread owns one (env, agent) row and walks t = T-1 .. 0; consecutive threads own consecutive
// rows, so every load/store of a warp is one contiguous 128-byte line.  HBM-bound: 16 B per
// (t, row) (+4 B with advantages).  The loads of U steps are issued together into one of two register
// batches while the dependent float64 chain consumes the other, so each warp keeps up to 6*U lines
// in flight and never waits for a full memory round trip between batches.
//
// Arithmetic mirrors the reference's mixed precision operation by operation (numpy float32 arrays,
// Python-float scalars, a float64 `gae` accumulator), with contraction into FMA forbidden, so the
// result is bit-identical to the reference and not merely within tolerance:
//   Vhat(x) = fl32(fl32(x * scale) + shift)                          torch float32
//   delta   = fl32(fl32(r + fl32(fl32(g32 * Vhat[t+1]) * m[t+1])) - Vhat[t])
//   gae     = fl64(delta + fl64(fl32(gl32 * m[t+1]) * gae))          gae is float64
//   ret[t]  = fl32(fl64(gae + Vhat[t]))
#include "common.cuh"

namespace {

constexpr int kUnroll = 16;

struct GaeParams {
    const float* __restrict__ rewards;
    const float* __restrict__ value_preds;
    const float* __restrict__ masks;
    float* __restrict__ returns;
    float* __restrict__ advantages;
    int T;
    long long R;        // row stride of every [t] slice
    long long n_rows;   // rows this launch scans (<= R)
    float g32;       // float32(gamma)
    float gl32;      // float32(gamma * gae_lambda), the product taken in double first
    float scale, shift;
};

template <bool DENORM>
__device__ __forceinline__ float vhat(float x, float scale, float shift) {
    return DENORM ? __fadd_rn(__fmul_rn(x, scale), shift) : x;
}

struct Carry {
    float m_next, vh_next, ret_next;
    double gae;
};

template <bool USE_GAE>
struct Batch {
    float rw[kUnroll], vv[kUnroll], mk[kUnroll];
    // issue the 3*U loads of steps t_hi-1 .. t_hi-U (those >= 0); nothing here depends on the scan
    __device__ __forceinline__ void load(const float* rew, const float* vp, const float* msk, int t_hi,
                                         long long R) {
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) {
            const int t = t_hi - 1 - i;
            if (t >= 0) {
                rw[i] = mm::ld_stream(rew + (long long)t * R);
                if (USE_GAE) vv[i] = mm::ld_stream(vp + (long long)t * R);
                mk[i] = (t > 0) ? mm::ld_stream(msk + (long long)t * R) : 0.f;   // masks[0] is never used
            }
        }
    }
};

// U steps of the scan, latest step first.  Three phases so that a single in-order warp still overlaps
// latencies: (1) everything that does not depend on the accumulator (denormalise, delta, coefficient,
// the float32->float64 conversions) for all U steps at once, (2) the loop-carried float64 chain, two
// dependent instructions per step, (3) rounding, advantages and the stores.  Steps with t < 0 (only in
// the last chunk, and after every valid step in chain order) compute on stale data and store nothing.
template <bool USE_GAE, bool DENORM, bool ADV, int U>
__device__ __forceinline__ void scan_chunk(const float (&rw)[U], const float (&vv)[U], const float (&mk)[U],
                                           Carry& c, const GaeParams& p, float* ret, float* adv, int t_hi,
                                           long long R) {
    if (USE_GAE) {
        float vh[U];
        double delta[U], coef[U], acc[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
            vh[i] = vhat<DENORM>(vv[i], p.scale, p.shift);
            const float vh_next = (i == 0) ? c.vh_next : vh[i - 1];
            const float m_next = (i == 0) ? c.m_next : mk[i - 1];
            const float boot = __fmul_rn(__fmul_rn(p.g32, vh_next), m_next);
            delta[i] = (double)__fsub_rn(__fadd_rn(rw[i], boot), vh[i]);
            coef[i] = (double)__fmul_rn(p.gl32, m_next);
        }
        double g = c.gae;
#pragma unroll
        for (int i = 0; i < U; ++i) {
            g = __dadd_rn(delta[i], __dmul_rn(coef[i], g));
            acc[i] = g;
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const int t = t_hi - 1 - i;
            const float out = (float)__dadd_rn(acc[i], (double)vh[i]);
            if (t >= 0) {
                ret[(long long)t * R] = out;
                if (ADV) adv[(long long)t * R] = __fsub_rn(out, vh[i]);
            }
        }
        const int last = min(U, t_hi) - 1;          // carry = the oldest VALID step of this chunk
#pragma unroll
        for (int i = 0; i < U; ++i)
            if (i == last) { c.gae = acc[i]; c.vh_next = vh[i]; c.m_next = mk[i]; }
    } else {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const int t = t_hi - 1 - i;
            if (t >= 0) {
                const float out = __fadd_rn(rw[i], __fmul_rn(__fmul_rn(c.ret_next, p.g32), c.m_next));
                c.ret_next = out;
                ret[(long long)t * R] = out;
                c.m_next = mk[i];
            }
        }
    }
}

template <bool USE_GAE, bool DENORM, bool ADV>
__device__ __forceinline__ void consume(const Batch<USE_GAE>& b, Carry& c, const GaeParams& p, float* ret,
                                        float* adv, int t_hi, long long R) {
    if (t_hi <= 0) return;
    scan_chunk<USE_GAE, DENORM, ADV, kUnroll>(b.rw, b.vv, b.mk, c, p, ret, adv, t_hi, R);
}

template <bool USE_GAE, bool DENORM, bool ADV>
__global__ void __launch_bounds__(128) gae_kernel(const GaeParams p) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_rows) return;
    const long long R = p.R;
    const float* rew = p.rewards + r;
    const float* vp = p.value_preds + r;
    const float* mk = p.masks + r;
    float* ret = p.returns + r;
    float* adv = ADV ? p.advantages + r : nullptr;

    // two register batches: the loads of batch k+1 are in flight while the float64 chain walks batch k
    Batch<USE_GAE> a, b;
    a.load(rew, vp, mk, p.T, R);
    Carry c;
    c.m_next = mm::ld_stream(mk + (long long)p.T * R);
    c.vh_next = 0.f; c.ret_next = 0.f; c.gae = 0.0;
    if (USE_GAE) c.vh_next = vhat<DENORM>(mm::ld_stream(vp + (long long)p.T * R), p.scale, p.shift);
    else         c.ret_next = ret[(long long)p.T * R];

    for (int t_hi = p.T; t_hi > 0; t_hi -= 2 * kUnroll) {
        b.load(rew, vp, mk, t_hi - kUnroll, R);
        consume<USE_GAE, DENORM, ADV>(a, c, p, ret, adv, t_hi, R);
        a.load(rew, vp, mk, t_hi - 2 * kUnroll, R);
        consume<USE_GAE, DENORM, ADV>(b, c, p, ret, adv, t_hi - kUnroll, R);
    }
}

// ---------------------------------------------------------------------------------------------
// Main path: one warp per tile of 32 consecutive rows.  The [T x 32] slices of rewards / value_preds /
// masks are staged through shared memory with 16-byte cp.async (LDGSTS) in chunks of kChunk steps, two
// chunks in flight per warp.  LDGSTS has no per-warp depth cap (plain LDG stops at ~55 outstanding
// requests per warp), so with only ~7 warps per SM at the BASELINE size (32 768 rows) the bytes in
// flight per SM go from ~49 KB to ~170 KB, which is what it takes to cover HBM latency at full bandwidth.
// The scan itself reads shared memory row-contiguously (conflict free) and streams results out with
// plain coalesced stores.
// ---------------------------------------------------------------------------------------------
constexpr int kChunk = 16;
constexpr int kStages = 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool USE_GAE, bool DENORM, bool ADV>
__global__ void __launch_bounds__(32) gae_tile_kernel(const GaeParams p) {
    __shared__ __align__(16) float sm[kStages][3][kChunk][32];
    const int lane = threadIdx.x;
    const long long R = p.R;
    const long long row0 = (long long)blockIdx.x * 32;
    const long long r = row0 + lane;
    const int seg = (lane & 7) * 4, sub = lane >> 3;     // 8 lanes x 16 B cover one 128-byte row slice

    auto issue = [&](int stage, int t_hi) {
#pragma unroll
        for (int j = 0; j < kChunk / 4; ++j) {
            const int i = 4 * j + sub;               // position inside the chunk, 0 = latest step
            const int t = t_hi - 1 - i;
            if (t >= 0) {
                const long long off = (long long)t * R + row0 + seg;
                cp_async16(&sm[stage][0][i][seg], p.rewards + off);
                if (USE_GAE) cp_async16(&sm[stage][1][i][seg], p.value_preds + off);
                cp_async16(&sm[stage][2][i][seg], p.masks + off);
            }
        }
        cp_async_commit();
    };

    const int n_chunks = (p.T + kChunk - 1) / kChunk;
#pragma unroll
    for (int st = 0; st < kStages; ++st) issue(st, p.T - st * kChunk);

    Carry c;
    c.m_next = mm::ld_stream(p.masks + (long long)p.T * R + r);
    c.vh_next = 0.f; c.ret_next = 0.f; c.gae = 0.0;
    if (USE_GAE) c.vh_next = vhat<DENORM>(mm::ld_stream(p.value_preds + (long long)p.T * R + r), p.scale, p.shift);
    else         c.ret_next = p.returns[(long long)p.T * R + r];
    float* ret = p.returns + r;
    float* adv = ADV ? p.advantages + r : nullptr;

    for (int ch = 0; ch < n_chunks; ++ch) {
        const int stage = ch % kStages;
        const int t_hi = p.T - ch * kChunk;
        cp_async_wait<kStages - 1>();                  // everything but the newest kStages-1 groups has landed
        __syncwarp();                                  // ... and is visible to the other lanes
        float rw[kChunk], vv[kChunk], mk[kChunk];
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
            rw[i] = sm[stage][0][i][lane];
            vv[i] = USE_GAE ? sm[stage][1][i][lane] : 0.f;
            mk[i] = sm[stage][2][i][lane];
        }
        scan_chunk<USE_GAE, DENORM, ADV, kChunk>(rw, vv, mk, c, p, ret, adv, t_hi, R);
        __syncwarp();                                  // every lane is done with this stage: refill it
        issue(stage, t_hi - kStages * kChunk);
    }
}

}  // namespace

extern "C" int mm_gae_returns(const float* rewards, const float* value_preds, const float* masks,
                              float* returns, float* advantages, int32_t T, int64_t R, double gamma,
                              double gae_lambda, int32_t use_gae, int32_t use_denorm, float denorm_scale,
                              float denorm_shift, void* stream) {
    if (!rewards || !masks || !returns || T <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (use_gae && !value_preds) return MM_ERR_BAD_ARG;
    if (advantages && !use_gae) return MM_ERR_BAD_ARG;
    GaeParams p;
    p.rewards = rewards; p.value_preds = value_preds; p.masks = masks;
    p.returns = returns; p.advantages = advantages;
    p.T = T; p.R = R; p.n_rows = R;
    p.g32 = (float)gamma;
    p.gl32 = (float)(gamma * gae_lambda);
    p.scale = denorm_scale; p.shift = denorm_shift;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool adv = advantages != nullptr;
#define MM_GAE_LAUNCH(KERNEL, GRID, THREADS, PARAMS)                                                   \
    do {                                                                                               \
        if (use_gae) {                                                                                 \
            if (use_denorm) { if (adv) KERNEL<true, true, true><<<GRID, THREADS, 0, st>>>(PARAMS);     \
                              else     KERNEL<true, true, false><<<GRID, THREADS, 0, st>>>(PARAMS); }  \
            else            { if (adv) KERNEL<true, false, true><<<GRID, THREADS, 0, st>>>(PARAMS);    \
                              else     KERNEL<true, false, false><<<GRID, THREADS, 0, st>>>(PARAMS); } \
        } else {                                                                                       \
            KERNEL<false, false, false><<<GRID, THREADS, 0, st>>>(PARAMS);                             \
        }                                                                                              \
    } while (0)
    // cp.async tiles need 16-byte aligned 32-row slices: R % 4 == 0 and 16-byte aligned bases
    const bool tiled_ok = (R % 4 == 0) && mm::aligned(rewards, 16) && mm::aligned(masks, 16) &&
                          (!use_gae || mm::aligned(value_preds, 16));
    long long done_rows = 0;
    if (tiled_ok && R >= 32) {
        const long long tiles = R / 32;
        MM_GAE_LAUNCH(gae_tile_kernel, (unsigned)tiles, 32, p);
        done_rows = tiles * 32;
    }
    if (done_rows < R) {
        // remainder rows (or unaligned inputs): register-batched kernel on the tail
        GaeParams q = p;
        q.rewards += done_rows; q.masks += done_rows; q.returns += done_rows;
        if (q.value_preds) q.value_preds += done_rows;
        if (q.advantages) q.advantages += done_rows;
        q.R = R;                       // row stride stays R ...
        q.n_rows = R - done_rows;      // ... but only the tail rows are scanned
        const int threads = (q.n_rows <= 148LL * 32 * 16) ? 32 : 128;
        const dim3 grid((unsigned)((q.n_rows + threads - 1) / threads));
        MM_GAE_LAUNCH(gae_kernel, grid, threads, q);
    }
#undef MM_GAE_LAUNCH
    return mm::check_launch();
}
