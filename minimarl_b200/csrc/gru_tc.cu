// Recurrent half of nn.GRU (forward and BPTT backward) on tcgen05 for the DIFFERENTIATED network pass.
//
// Replaces cuDNN's GRU in RNNLayer.forward (offpolicy/utils/algorithm_utils/rnn.py:26-29;
// mappo/utils/algorithm_utils/rnn.py:24-80) for hidden size H in {32, 64}, one layer.  The input projection
// gi = x W_ih^T + b_ih (and in the backward dx, dW_ih, dW_hh) are large row-parallel GEMMs that stay in cuBLAS;
// what is sequential -- gh = h_{t-1} W_hh^T, the gates, and in the backward dh_{t-1} = dgh W_hh -- runs here:
// one CTA per 64 rows (tcgen05 M = 64: row r in TMEM lane 32 (r / 16) + r % 16, see mm_tc_lane_probe) walks the
// sequence with W_hh (forward) or W_hh^T (backward) and the hidden / gate-gradient tile resident in shared memory
// in tcgen05 operand layout, accumulators in TMEM; 512 threads = 4 quarters x 4 column parts x 2 half-warps, the
// upper half-warp receiving its 8 accumulator columns by shuffle.  cuDNN's backward for this shape (101 steps x
// 8 192 rows x 64) takes 15 ms, 11.5 ms of it in one element-wise kernel; this pair takes 0.44 + 0.61 ms.
//   forward : r = s(gi_r + gh_r + b_hr), z = s(gi_z + gh_z + b_hz), ghn = gh_n + b_hn, n = tanh(gi_n + r ghn),
//             h_t = (1 - z) n + z h_{t-1};           saves r, z, n, ghn for the backward
//   backward: dn = dh (1 - z), dz = dh (h_{t-1} - n), a = dn (1 - n^2), dgi_n = a, dgh_n = a r, dr = a ghn,
//             dgi_r = dgh_r = dr r (1 - r), dgi_z = dgh_z = dz z (1 - z), dh_{t-1} = dh z + dgh W_hh
// Operands are TF32 (like cuDNN's default allow_tf32), accumulation and all gate math fp32
// (ex2.approx-based sigmoid / tanh, relative error ~1e-6: this pass produces the gradients).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kThreads = 512;

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// ex2.approx-based exponentials (max relative error ~2^-22 on the exponent range used here): ~10x fewer
// instructions than expf / tanhf, and far below the TF32 operand error of the recurrence
__device__ __forceinline__ float sigmoid_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) {
    const float e = __expf(-2.0f * fabsf(x));                 // in (0, 1]: no overflow
    return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}

template <int N>
__device__ __forceinline__ void ld_row(const float* __restrict__ p, float (&v)[N], bool ok) {
#pragma unroll
    for (int k = 0; k < N; k += 4) {
        const float4 q = ok ? *reinterpret_cast<const float4*>(p + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
    }
}
template <int N>
__device__ __forceinline__ void st_row(float* __restrict__ p, const float (&v)[N], bool ok) {
    if (!ok) return;
#pragma unroll
    for (int k = 0; k < N; k += 4) *reinterpret_cast<float4*>(p + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ok) {
    const uint32_t n = ok ? 16u : 0u;                                      // 0 source bytes: zero-fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct GruFwdArgs {
    const float *gi, *h0, *whh, *bhh;          // gi [T,R,3H]; h0 [R,H] or null; whh [3H,H]; bhh [3H]
    float *h_seq, *r, *z, *n, *ghn;            // [T,R,H] each (r/z/n/ghn may be null: inference)
    int T;
    long long R;
};

// ---- 64-row forward with coalesced tile I/O -----------------------------------------------------------------------------
// The thread = row binding that TMEM forces makes every per-thread global access touch 16 different 128-byte lines per warp
// instruction; ncu showed the L1 tag stage (lg_throttle / mio_throttle) busy for about half of every step.  Here global
// memory is touched only by whole-warp contiguous accesses: the 64 x 3H gi tile of the NEXT step is fetched one step ahead
// with cp.async, 16 bytes per thread and 512 contiguous bytes per warp instruction, into padded shared rows (row stride
// + 16 B: conflict-free when lane = row reads a float4 column); r, z, n, ghn, h are written to padded staging rows and leave
// at the top of the next step as 512-contiguous-byte stores by all 512 threads, overlapping that step's MMA.
// (Per-row cp.async.bulk copies were tried first: UBLKCP takes uniform operands, so 32 lanes with 32 addresses serialise
// into 32 issues of ~90 cycles each -- 2 900 cycles per step on the critical path.)
template <int H>
__global__ void __launch_bounds__(kThreads, 1) gru_fwd_tile_kernel(const GruFwdArgs a) {
    constexpr int ROWS = 64, CPT = H / 4;
    constexpr int GI_PIECES = 3 * H / 4, OUT_PIECES = H / 4;               // 16-byte pieces per row
    constexpr int GI_ROW = 3 * H * 4 + 16, OUT_ROW = H * 4 + 16;           // padded row strides (bytes)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* sW = smem;                                                   // W_hh [3H x H] K-major
    uint8_t* sH = sW + 3 * H * H * 4;                                     // h [64 x H] K-major
    float* sB = reinterpret_cast<float*>(sH + ROWS * H * 4);              // b_hh [3H]
    uint8_t* sGi = reinterpret_cast<uint8_t*>(sB + 3 * H);                // gi rows of the coming step
    uint8_t* sOut = sGi + ROWS * GI_ROW;                                  // [5][64] rows: h | r | z | n | ghn
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // M = 64 keeps rows in the lower 16 lanes of every TMEM quarter: lanes 0..15 read the 16 accumulator columns of their
    // row and hand the upper 8 to lane + 16 by shuffle, so that all 32 lanes do gate math (8 columns each)
    constexpr int CPH = CPT / 2;
    const int half = lane >> 4;
    const int row = (warp & 3) * 16 + (lane & 15), part = warp >> 2, c0 = part * CPT, cc = c0 + half * CPH;
    const long long row0 = (long long)blockIdx.x * ROWS;
    const int n_valid = (int)min((long long)ROWS, a.R - row0);
    const bool valid = row < n_valid;
    const bool train = a.r != nullptr;
    constexpr uint32_t kCols = 3 * H <= 128 ? 128 : 256;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, kCols);
    if (tid == 0) tc::mbar_init(&bar, 1);
    for (int i = tid; i < 3 * H * H; i += kThreads) {
        const int rr = i / H, k = i - rr * H;
        *reinterpret_cast<float*>(sW + tc::kmajor_offset(rr, k, H)) = to_tf32(a.whh[i]);
    }
    for (int i = tid; i < 3 * H; i += kThreads) sB[i] = a.bhh[i];
    // the gi tile of one step is 64 x 3H floats contiguous in global memory: piece u of it goes to padded row u / GI_PIECES
    auto prefetch = [&](int t) {
        const float* g = a.gi + ((long long)t * a.R + row0) * 3 * H;
#pragma unroll
        for (int u = tid; u < ROWS * GI_PIECES; u += kThreads) {
            const int rr = u / GI_PIECES, c = u - rr * GI_PIECES;
            cp_async16(tc::smem_u32(sGi + rr * GI_ROW + c * 16), g + (rr < n_valid ? u : 0) * 4, rr < n_valid);
        }
        cp_async_commit();
    };
    prefetch(0);
    float h[CPH];
    ld_row<CPH>(a.h0 ? a.h0 + (row0 + row) * H + cc : nullptr, h, valid && a.h0 != nullptr);
#pragma unroll
    for (int k = 0; k < CPH; k += 4)
        *reinterpret_cast<float4*>(sH + tc::kmajor_offset(row, cc + k, H)) = make_float4(h[k], h[k + 1], h[k + 2], h[k + 3]);
    cp_async_wait_all();
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sbo = (uint32_t)(H >> 2) * 128u;
    const uint64_t dH = tc::make_desc(tc::smem_u32(sH), 128u, sbo);
    const uint64_t dW_rz = tc::make_desc(tc::smem_u32(sW), 128u, sbo);
    const uint64_t dW_n = tc::make_desc(tc::smem_u32(sW) + (uint32_t)(2 * H >> 3) * sbo, 128u, sbo);
    const uint32_t idesc_2h = tc::make_idesc_tf32(ROWS, 2 * H), idesc_h = tc::make_idesc_tf32(ROWS, H);
    uint32_t phase = 0;
    const int n_out = train ? 5 : 1;
    // copy-out of one step: piece u of the [n_out][64 x H] staging goes to tensor u / (64 OUT_PIECES), 512 B per warp instruction
    auto copy_out = [&](int t) {
        for (int u = tid; u < n_out * ROWS * OUT_PIECES; u += kThreads) {
            const int s = u / (ROWS * OUT_PIECES), v = u - s * (ROWS * OUT_PIECES), rr = v / OUT_PIECES, c = v - rr * OUT_PIECES;
            if (rr < n_valid) {
                float* dst = (s == 0 ? a.h_seq : s == 1 ? a.r : s == 2 ? a.z : s == 3 ? a.n : a.ghn) +
                             ((long long)t * a.R + row0) * H + v * 4;
                *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(sOut + (s * ROWS + rr) * OUT_ROW + c * 16);
            }
        }
    };

    for (int t = 0; t < a.T; ++t) {
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem, dH + 16 * kk, dW_rz + 16 * kk, idesc_2h, kk > 0);
#pragma unroll
            for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 2 * H, dH + 16 * kk, dW_n + 16 * kk, idesc_h, kk > 0);
            tc::mma_commit(&bar);
        }
        if (t > 0) copy_out(t - 1);                                       // while the tensor core works
        float gir[CPH], giz[CPH], gin[CPH];
        {
            const uint8_t* g = sGi + row * GI_ROW + cc * 4;
#pragma unroll
            for (int k = 0; k < CPH; k += 4) {
                const float4 q0 = *reinterpret_cast<const float4*>(g + k * 4);
                const float4 q1 = *reinterpret_cast<const float4*>(g + (H + k) * 4);
                const float4 q2 = *reinterpret_cast<const float4*>(g + (2 * H + k) * 4);
                gir[k] = q0.x; gir[k + 1] = q0.y; gir[k + 2] = q0.z; gir[k + 3] = q0.w;
                giz[k] = q1.x; giz[k + 1] = q1.y; giz[k + 2] = q1.z; giz[k + 3] = q1.w;
                gin[k] = q2.x; gin[k + 1] = q2.y; gin[k + 2] = q2.z; gin[k + 3] = q2.w;
            }
        }
        if (warp == 0) tc::mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();                                                  // gi rows and output staging are free again
        tc::fence_after_sync();
        if (t + 1 < a.T) prefetch(t + 1);
        float ar[CPT], az[CPT], an[CPT];
        tc::tmem_ld_issue<CPT>(lane_addr + c0, ar);
        tc::tmem_ld_issue<CPT>(lane_addr + H + c0, az);
        tc::tmem_ld_issue<CPT>(lane_addr + 2 * H + c0, an);
        tc::tmem_ld_fence<CPT>(ar, az, an);
        float gr[CPH], gz[CPH], gn[CPH];
#pragma unroll
        for (int j = 0; j < CPH; ++j) {                                   // upper 8 columns travel to lane + 16
            const float ur = __shfl_sync(MM_FULL_MASK, ar[CPH + j], lane & 15);
            const float uz = __shfl_sync(MM_FULL_MASK, az[CPH + j], lane & 15);
            const float un = __shfl_sync(MM_FULL_MASK, an[CPH + j], lane & 15);
            gr[j] = half ? ur : ar[j];
            gz[j] = half ? uz : az[j];
            gn[j] = half ? un : an[j];
        }
        {
            float rr[CPH], zz[CPH], nn[CPH];
#pragma unroll
            for (int j = 0; j < CPH; ++j) {
                rr[j] = sigmoid_(gir[j] + gr[j] + sB[cc + j]);
                zz[j] = sigmoid_(giz[j] + gz[j] + sB[H + cc + j]);
                gn[j] += sB[2 * H + cc + j];
                nn[j] = tanh_(gin[j] + rr[j] * gn[j]);
                h[j] = (1.0f - zz[j]) * nn[j] + zz[j] * h[j];
            }
            uint8_t* o = sOut + row * OUT_ROW + cc * 4;
#pragma unroll
            for (int k = 0; k < CPH; k += 4) {
                *reinterpret_cast<float4*>(o + k * 4) = make_float4(h[k], h[k + 1], h[k + 2], h[k + 3]);
                if (train) {
                    *reinterpret_cast<float4*>(o + 1 * ROWS * OUT_ROW + k * 4) = make_float4(rr[k], rr[k + 1], rr[k + 2], rr[k + 3]);
                    *reinterpret_cast<float4*>(o + 2 * ROWS * OUT_ROW + k * 4) = make_float4(zz[k], zz[k + 1], zz[k + 2], zz[k + 3]);
                    *reinterpret_cast<float4*>(o + 3 * ROWS * OUT_ROW + k * 4) = make_float4(nn[k], nn[k + 1], nn[k + 2], nn[k + 3]);
                    *reinterpret_cast<float4*>(o + 4 * ROWS * OUT_ROW + k * 4) = make_float4(gn[k], gn[k + 1], gn[k + 2], gn[k + 3]);
                }
                *reinterpret_cast<float4*>(sH + tc::kmajor_offset(row, cc + k, H)) = make_float4(h[k], h[k + 1], h[k + 2], h[k + 3]);
            }
        }
        cp_async_wait_all();                                              // this thread's share of the next gi tile has landed
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    copy_out(a.T - 1);
    if (warp == 0) tc::tmem_dealloc(tmem, kCols);
}

struct GruBwdArgs {
    const float *dh_seq, *r, *z, *n, *ghn, *h_seq, *h0, *whh;   // [T,R,H] x6; h0 [R,H] or null; whh [3H,H]
    float *dgi, *dghn, *dh0;                                    // [T,R,3H]; [T,R,H]; [R,H]
    int T;
    long long R;
    float* scale2;                                              // optional: scale2[1] <- max |dgi| as float bits (atomicMax)
};

// ---- 64-row backward with coalesced tile loads --------------------------------------------------------------------------
// Same idea as gru_fwd_tile_kernel: the six [64 x H] input tiles of the NEXT (earlier) step -- dh, r, z, n, ghn, h_prev, each
// contiguous in global memory -- are fetched with 512-contiguous-byte cp.async instructions into padded shared rows while
// this step computes; both half-warps do gate-gradient math (8 columns each).  The gradient stores stay per-thread
// (two 32-byte runs per row and tensor): a staging area for them does not fit beside W_hh^T, the dgh operand tile and the inputs.
template <int H>
__global__ void __launch_bounds__(kThreads, 1) gru_bwd_tile_kernel(const GruBwdArgs a) {
    constexpr int ROWS = 64, CPT = H / 4, CPH = CPT / 2, K3 = 3 * H;
    constexpr int PIECES = H / 4, IN_ROW = H * 4 + 16;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* sWt = smem;                       // W_hh^T [H x 3H] K-major (B operand: N = H rows, K = 3H)
    uint8_t* sG = sWt + H * K3 * 4;            // dgh [64 x 3H] K-major (A operand)
    uint8_t* sIn = sG + ROWS * K3 * 4;         // [6][64] padded rows: dh | r | z | n | ghn | h_prev
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, half = lane >> 4;
    const int row = (warp & 3) * 16 + (lane & 15), part = warp >> 2, c0 = part * CPT, cc = c0 + half * CPH;
    const long long row0 = (long long)blockIdx.x * ROWS;
    const int n_valid = (int)min((long long)ROWS, a.R - row0);
    const bool valid = row < n_valid;
    constexpr uint32_t kCols = H <= 32 ? 32 : 64;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, kCols);
    if (tid == 0) tc::mbar_init(&bar, 1);
    for (int i = tid; i < K3 * H; i += kThreads) {   // whh[g][j] (g in 3H, j in H) -> sWt row j, column g
        const int g = i / H, j = i - g * H;
        *reinterpret_cast<float*>(sWt + tc::kmajor_offset(j, g, K3)) = to_tf32(a.whh[i]);
    }
    auto prefetch = [&](int t) {
        const long long base = ((long long)t * a.R + row0) * H;
#pragma unroll
        for (int u = tid; u < 6 * ROWS * PIECES; u += kThreads) {
            const int s = u / (ROWS * PIECES), v = u - s * (ROWS * PIECES), rr = v / PIECES, c = v - rr * PIECES;
            const float* src;
            bool ok = rr < n_valid;
            if (s < 5) src = (s == 0 ? a.dh_seq : s == 1 ? a.r : s == 2 ? a.z : s == 3 ? a.n : a.ghn) + base;
            else if (t > 0) src = a.h_seq + base - a.R * H;
            else { src = a.h0 ? a.h0 + row0 * H : a.dh_seq; ok = ok && a.h0 != nullptr; }
            cp_async16(tc::smem_u32(sIn + (s * ROWS + rr) * IN_ROW + c * 16), src + (ok ? v : 0) * 4, ok);
        }
        cp_async_commit();
    };
    prefetch(a.T - 1);
    cp_async_wait_all();
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sbo = (uint32_t)(K3 >> 2) * 128u;
    const uint64_t dG = tc::make_desc(tc::smem_u32(sG), 128u, sbo), dW = tc::make_desc(tc::smem_u32(sWt), 128u, sbo);
    const uint32_t idesc = tc::make_idesc_tf32(ROWS, H);
    uint32_t phase = 0;
    float dh_rec[CPH];
#pragma unroll
    for (int j = 0; j < CPH; ++j) dh_rec[j] = 0.f;
    float amax = 0.f;                                           // max |dgi| seen by this thread (for the fp16 consumers' scale)

    for (int t = a.T - 1; t >= 0; --t) {
        float dh[CPH], rr[CPH], zz[CPH], nn[CPH], gn[CPH], hp[CPH];
        {
            const uint8_t* in = sIn + row * IN_ROW + cc * 4;
#pragma unroll
            for (int k = 0; k < CPH; k += 4) {
                const float4 q0 = *reinterpret_cast<const float4*>(in + k * 4);
                const float4 q1 = *reinterpret_cast<const float4*>(in + 1 * ROWS * IN_ROW + k * 4);
                const float4 q2 = *reinterpret_cast<const float4*>(in + 2 * ROWS * IN_ROW + k * 4);
                const float4 q3 = *reinterpret_cast<const float4*>(in + 3 * ROWS * IN_ROW + k * 4);
                const float4 q4 = *reinterpret_cast<const float4*>(in + 4 * ROWS * IN_ROW + k * 4);
                const float4 q5 = *reinterpret_cast<const float4*>(in + 5 * ROWS * IN_ROW + k * 4);
                dh[k] = q0.x; dh[k + 1] = q0.y; dh[k + 2] = q0.z; dh[k + 3] = q0.w;
                rr[k] = q1.x; rr[k + 1] = q1.y; rr[k + 2] = q1.z; rr[k + 3] = q1.w;
                zz[k] = q2.x; zz[k + 1] = q2.y; zz[k + 2] = q2.z; zz[k + 3] = q2.w;
                nn[k] = q3.x; nn[k + 1] = q3.y; nn[k + 2] = q3.z; nn[k + 3] = q3.w;
                gn[k] = q4.x; gn[k + 1] = q4.y; gn[k + 2] = q4.z; gn[k + 3] = q4.w;
                hp[k] = q5.x; hp[k + 1] = q5.y; hp[k + 2] = q5.z; hp[k + 3] = q5.w;
            }
        }
        __syncthreads();                           // every thread holds its inputs: the rows can take the next step's
        if (t > 0) prefetch(t - 1);                // in flight during the rest of this step
        float d_r[CPH], d_z[CPH], d_n[CPH], d_hn[CPH];
#pragma unroll
        for (int j = 0; j < CPH; ++j) {
            const float g = dh[j] + dh_rec[j];
            const float an = g * (1.0f - zz[j]) * (1.0f - nn[j] * nn[j]);
            d_n[j] = an;
            d_hn[j] = an * rr[j];
            d_r[j] = an * gn[j] * rr[j] * (1.0f - rr[j]);
            d_z[j] = g * (hp[j] - nn[j]) * zz[j] * (1.0f - zz[j]);
            dh_rec[j] = g * zz[j];                                   // direct path; the W_hh path is added after the MMA
            amax = fmaxf(amax, fmaxf(fabsf(d_n[j]), fmaxf(fabsf(d_r[j]), fabsf(d_z[j]))));
        }
        const long long o = ((long long)t * a.R + row0 + row) * H + cc;
        float* dgi = a.dgi + ((long long)t * a.R + row0 + row) * K3 + cc;
        st_row<CPH>(dgi, d_r, valid);
        st_row<CPH>(dgi + H, d_z, valid);
        st_row<CPH>(dgi + 2 * H, d_n, valid);
        st_row<CPH>(a.dghn + o, d_hn, valid);
#pragma unroll
        for (int k = 0; k < CPH; k += 4) {             // dgh = [d_r | d_z | d_hn] as the A operand
            *reinterpret_cast<float4*>(sG + tc::kmajor_offset(row, cc + k, K3)) =
                make_float4(to_tf32(d_r[k]), to_tf32(d_r[k + 1]), to_tf32(d_r[k + 2]), to_tf32(d_r[k + 3]));
            *reinterpret_cast<float4*>(sG + tc::kmajor_offset(row, H + cc + k, K3)) =
                make_float4(to_tf32(d_z[k]), to_tf32(d_z[k + 1]), to_tf32(d_z[k + 2]), to_tf32(d_z[k + 3]));
            *reinterpret_cast<float4*>(sG + tc::kmajor_offset(row, 2 * H + cc + k, K3)) =
                make_float4(to_tf32(d_hn[k]), to_tf32(d_hn[k + 1]), to_tf32(d_hn[k + 2]), to_tf32(d_hn[k + 3]));
        }
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < K3 / 8; ++kk) tc::mma_tf32(tmem, dG + 16 * kk, dW + 16 * kk, idesc, kk > 0);
            tc::mma_commit(&bar);
        }
        cp_async_wait_all();                       // this thread's share of the next inputs has landed
        if (warp == 0) tc::mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        tc::fence_after_sync();
        float mm_[CPT];
        tc::tmem_ld<CPT>(lane_addr + c0, mm_);
#pragma unroll
        for (int j = 0; j < CPH; ++j) {            // upper 8 columns travel to lane + 16
            const float u = __shfl_sync(MM_FULL_MASK, mm_[CPH + j], lane & 15);
            dh_rec[j] += half ? u : mm_[j];
        }
        tc::fence_before_sync();          // the next step's MMA overwrites this accumulator: order the loads before it
    }
    if (a.dh0 != nullptr) st_row<CPH>(a.dh0 + (row0 + row) * H + cc, dh_rec, valid);
    if (a.scale2 != nullptr) {
        unsigned int m = valid ? __float_as_uint(amax) : 0u;        // non-negative floats order like their bit patterns
#pragma unroll
        for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(MM_FULL_MASK, m, o));
        if (lane == 0 && m) atomicMax(reinterpret_cast<unsigned int*>(a.scale2) + 1, m);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, kCols);
}

// scale2[0] = the power of two that puts max |dgi| (scale2[1]) at 64..128: what mm_ppo_trunk2_scale computes in a pass of its own
__global__ void gru_scale_finish_kernel(float* __restrict__ scale2) {
    const float amax = scale2[1];
    int e = (amax > 0.f && amax < 3.0e38f) ? ilogbf(64.0f / fmaxf(amax, 1e-30f)) : 0;     // non-finite gradients stay non-finite
    e = e < -60 ? -60 : e > 60 ? 60 : e;
    scale2[0] = ldexpf(1.0f, e);
}

template <int H>
int launch_fwd_tile(const GruFwdArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)4 * (3 * H * H + 64 * H + 3 * H) + (size_t)64 * (3 * H * 4 + 16) + (size_t)5 * 64 * (H * 4 + 16) + 128;
    cudaError_t e = cudaFuncSetAttribute(gru_fwd_tile_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    gru_fwd_tile_kernel<H><<<(unsigned)((a.R + 63) / 64), kThreads, smem, st>>>(a);
    return mm::check_launch();
}
template <int H>
int launch_bwd_tile(const GruBwdArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)4 * (3 * H * H + 64 * 3 * H) + (size_t)6 * 64 * (H * 4 + 16) + 128;
    cudaError_t e = cudaFuncSetAttribute(gru_bwd_tile_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    gru_bwd_tile_kernel<H><<<(unsigned)((a.R + 63) / 64), kThreads, smem, st>>>(a);
    return mm::check_launch();
}
bool aligned16(const void* p) { return p == nullptr || mm::aligned(p, 16); }

}  // namespace

extern "C" int mm_gru_seq_fwd(const float* gi, const float* h0, const float* w_hh, const float* b_hh, float* h_seq,
                              float* r, float* z, float* n, float* ghn, int32_t T, int64_t R, int32_t H, void* stream) {
    if (!gi || !w_hh || !b_hh || !h_seq || T <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (H != 32 && H != 64) return MM_ERR_UNSUPPORTED;
    if ((r || z || n || ghn) && !(r && z && n && ghn)) return MM_ERR_BAD_ARG;
    if (!aligned16(gi) || !aligned16(h0) || !aligned16(h_seq) || !aligned16(r) || !aligned16(z) || !aligned16(n) || !aligned16(ghn))
        return MM_ERR_BAD_ARG;
    GruFwdArgs a{gi, h0, w_hh, b_hh, h_seq, r, z, n, ghn, T, R};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return H == 64 ? launch_fwd_tile<64>(a, st) : launch_fwd_tile<32>(a, st);
}

// scale2 (optional, device float[2]): [1] <- max |dgi|, [0] <- the power-of-two scale mm_ppo_trunk2_scale(dgi) would return
extern "C" int mm_gru_seq_bwd_scaled(const float* dh_seq, const float* r, const float* z, const float* n, const float* ghn,
                                     const float* h_seq, const float* h0, const float* w_hh, float* dgi, float* dghn,
                                     float* dh0, float* scale2, int32_t T, int64_t R, int32_t H, void* stream) {
    if (!dh_seq || !r || !z || !n || !ghn || !h_seq || !w_hh || !dgi || !dghn || T <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (H != 32 && H != 64) return MM_ERR_UNSUPPORTED;
    if (!aligned16(dh_seq) || !aligned16(r) || !aligned16(z) || !aligned16(n) || !aligned16(ghn) || !aligned16(h_seq) ||
        !aligned16(h0) || !aligned16(dgi) || !aligned16(dghn) || !aligned16(dh0))
        return MM_ERR_BAD_ARG;
    GruBwdArgs a{dh_seq, r, z, n, ghn, h_seq, h0, w_hh, dgi, dghn, dh0, T, R, scale2};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (scale2) {
        cudaError_t e = cudaMemsetAsync(scale2, 0, 8, st);
        if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    }
    const int rc = H == 64 ? launch_bwd_tile<64>(a, st) : launch_bwd_tile<32>(a, st);
    if (rc != MM_OK || !scale2) return rc;
    gru_scale_finish_kernel<<<1, 1, 0, st>>>(scale2);
    return mm::check_launch();
}

extern "C" int mm_gru_seq_bwd(const float* dh_seq, const float* r, const float* z, const float* n, const float* ghn,
                              const float* h_seq, const float* h0, const float* w_hh, float* dgi, float* dghn,
                              float* dh0, int32_t T, int64_t R, int32_t H, void* stream) {
    return mm_gru_seq_bwd_scaled(dh_seq, r, z, n, ghn, h_seq, h0, w_hh, dgi, dghn, dh0, nullptr, T, R, H, stream);
}
