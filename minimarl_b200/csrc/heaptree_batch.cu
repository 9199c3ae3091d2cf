// Chunk-PER maintenance for N lockstep envs (top-level vdn/ and qmix/ packages), parallel formulations.
//
// The reference buffer (vdn/replay_buffer/buffer.py:29-39, sumtree.py:38-58) is fed one chunk at a time by one env.  With
// N lockstep envs every chunk boundary delivers N chunks AT ONCE and every training iteration writes back batch_size
// priorities at once.  heaptree.cu replays those as the reference's sequential adds / updates (bit-identical trees; the
// parity path).  The kernels here treat a batch as a SET, which is what "N envs finish a chunk at the same step" is:
//
//   mm_heaptree_add_set   the valid items of the batch first fill the free slots in order; the rest replace the
//                         lowest-priority slots that were occupied BEFORE the batch (ties: lowest slot), the r-th
//                         lowest going to the r-th remaining item.  Items of one batch never evict each other.  Ranks come
//                         from a brute-force count over (priority, slot) keys -- capacity^2 / 2 integer compares spread
//                         over the whole GPU (8 192 slots: 3.4e7 compares, a few microseconds) instead of a serial heap.
//   mm_heaptree_set_leaves  batch priority write-back: leaf = priority[i] for the LAST i naming that leaf (the reference's
//                         sequential order makes the last write win), then the inner nodes are recomputed bottom-up
//                         from the leaves (sums of exactly the current leaves -- no accumulated rounding).
//   mm_rows_scatter       payload rows of up to 8 tensors copied to their slots in one launch.
//
// A rebuilt inner node differs from the reference's incrementally propagated one by float64 rounding only, so sampled
// indices can differ only where a draw falls within 1e-16 (relative) of a leaf boundary.
#include "common.cuh"

namespace {

constexpr int kScanThreads = 1024;

struct SetHeader { int fill, evict, n_old, pad; };

// one CTA: exclusive scan of the valid flags -> vrank, header, n_data update
__global__ void __launch_bounds__(kScanThreads) set_scan_kernel(const uint8_t* __restrict__ valid, int k, int capacity,
                                                                 int* __restrict__ n_data, int* __restrict__ vrank,
                                                                 SetHeader* __restrict__ hdr) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < k; base += kScanThreads) {
        const int i = base + tid;
        const int v = (i < k && (valid == nullptr || valid[i])) ? 1 : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(MM_FULL_MASK, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(MM_FULL_MASK, t, o);
                if (lane >= o) t += y;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + x - v;
        if (i < k) vrank[i] = v ? before : -1;
        __syncthreads();
        if (tid == 0) carry += warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) {
        const int n_old = *n_data, k_eff = carry;
        const int fill = min(k_eff, capacity - n_old);
        hdr->fill = fill;
        hdr->evict = min(k_eff - fill, n_old);
        hdr->n_old = n_old;
        *n_data = n_old + fill;
    }
}

// rank of every old leaf among the old leaves by (priority, slot); the `evict` lowest publish their slot
__global__ void __launch_bounds__(256) set_rank_kernel(const double* __restrict__ tree, int capacity,
                                                       const SetHeader* __restrict__ hdr, int* __restrict__ evict_slot) {
    __shared__ unsigned long long keys[1024];
    const int n_old = hdr->n_old, evict = hdr->evict;
    if (evict == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long* leaf = reinterpret_cast<const unsigned long long*>(tree + capacity - 1);
    // priorities are non-negative doubles: their bit patterns order like the values
    const unsigned long long mine = i < n_old ? leaf[i] : 0ull;
    int rank = 0;
    for (int base = 0; base < n_old; base += 1024) {
        const int m = min(1024, n_old - base);
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += blockDim.x) keys[t] = leaf[base + t];
        __syncthreads();
        if (i < n_old) {
            for (int t = 0; t < m; ++t) {
                const unsigned long long kt = keys[t];
                rank += (kt < mine) || (kt == mine && base + t < i);
            }
        }
    }
    if (i < n_old && rank < evict) evict_slot[rank] = i;
}

__global__ void set_assign_kernel(double* __restrict__ tree, int capacity, const SetHeader* __restrict__ hdr,
                                  const int* __restrict__ vrank, const int* __restrict__ evict_slot,
                                  const double* __restrict__ priority, int k, int64_t* __restrict__ slot_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const int r = vrank[i];
    int slot = -1;
    if (r >= 0) {
        if (r < hdr->fill) slot = hdr->n_old + r;
        else if (r - hdr->fill < hdr->evict) slot = evict_slot[r - hdr->fill];
    }
    slot_out[i] = slot;
    if (slot >= 0) tree[capacity - 1 + slot] = priority[i];
}

// inner nodes from the leaves, deepest level first (one CTA; a level's nodes only read the level below)
__global__ void __launch_bounds__(1024) rebuild_kernel(double* __restrict__ tree, int capacity) {
    const int n_inner = capacity - 1;
    if (n_inner <= 0) return;
    int top = 31 - __clz(n_inner);                 // depth of the deepest inner node (node index n_inner - 1 -> 1-based n_inner)
    for (int d = top; d >= 0; --d) {
        const int lo = (1 << d) - 1, hi = min((1 << (d + 1)) - 1, n_inner);
        for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) tree[i] = tree[2 * i + 1] + tree[2 * i + 2];
        __syncthreads();
    }
}

__global__ void leaves_owner_kernel(const int64_t* __restrict__ node_idx, int n, int capacity, int* __restrict__ owner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long leaf = node_idx[i] - (capacity - 1);
    if (leaf >= 0 && leaf < capacity) atomicMax(owner + leaf, i);
}
__global__ void leaves_write_kernel(double* __restrict__ tree, const int64_t* __restrict__ node_idx,
                                    const double* __restrict__ priority, int n, int capacity, const int* __restrict__ owner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long leaf = node_idx[i] - (capacity - 1);
    if (leaf >= 0 && leaf < capacity && owner[leaf] == i) tree[node_idx[i]] = priority[i];
}

struct ScatterArgs {
    const float* src[8];
    float* dst[8];
    int row_floats[8];
    const int64_t* slot;
    int k, n;
};
__global__ void __launch_bounds__(256) rows_scatter_kernel(const __grid_constant__ ScatterArgs a) {
    const int t = blockIdx.y;
    const long long slot = a.slot[blockIdx.x];
    if (slot < 0) return;
    const int rf = a.row_floats[t];
    const float* s = a.src[t] + (long long)blockIdx.x * rf;
    float* d = a.dst[t] + slot * rf;
    if ((rf & 3) == 0 && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
        for (int i = threadIdx.x; i < (rf >> 2); i += blockDim.x)
            reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i];
    } else {
        for (int i = threadIdx.x; i < rf; i += blockDim.x) d[i] = s[i];
    }
}

}  // namespace

// workspace: int32[k + capacity + 4]
extern "C" int mm_heaptree_add_set(double* tree, int32_t capacity, int32_t* n_data_dev, const double* priority,
                                   const uint8_t* valid, int32_t k, int64_t* slot_out, int32_t* workspace, void* stream) {
    if (!tree || !n_data_dev || !priority || !slot_out || !workspace || capacity <= 0 || k <= 0) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SetHeader* hdr = reinterpret_cast<SetHeader*>(workspace);
    int* vrank = workspace + 4;
    int* evict_slot = vrank + k;
    set_scan_kernel<<<1, kScanThreads, 0, st>>>(valid, k, capacity, n_data_dev, vrank, hdr);
    set_rank_kernel<<<(capacity + 255) / 256, 256, 0, st>>>(tree, capacity, hdr, evict_slot);
    set_assign_kernel<<<(k + 255) / 256, 256, 0, st>>>(tree, capacity, hdr, vrank, evict_slot, priority, k, slot_out);
    rebuild_kernel<<<1, 1024, 0, st>>>(tree, capacity);
    return mm::check_launch();
}

// workspace: int32[capacity]
extern "C" int mm_heaptree_set_leaves(double* tree, int32_t capacity, const int64_t* node_idx, const double* priority, int32_t n,
                                      int32_t* workspace, void* stream) {
    if (!tree || !node_idx || !priority || !workspace || capacity <= 0 || n <= 0) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(workspace, 0xff, sizeof(int) * (size_t)capacity, st);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    leaves_owner_kernel<<<(n + 255) / 256, 256, 0, st>>>(node_idx, n, capacity, workspace);
    leaves_write_kernel<<<(n + 255) / 256, 256, 0, st>>>(tree, node_idx, priority, n, capacity, workspace);
    rebuild_kernel<<<1, 1024, 0, st>>>(tree, capacity);
    return mm::check_launch();
}

extern "C" int mm_rows_scatter(const float* const* src_host, float* const* dst_host, const int32_t* row_floats_host, int32_t n,
                               const int64_t* slot, int32_t k, void* stream) {
    if (!src_host || !dst_host || !row_floats_host || !slot || n <= 0 || n > 8 || k <= 0) return MM_ERR_BAD_ARG;
    ScatterArgs a{};
    for (int t = 0; t < n; ++t) {
        if (!src_host[t] || !dst_host[t] || row_floats_host[t] <= 0) return MM_ERR_BAD_ARG;
        a.src[t] = src_host[t]; a.dst[t] = dst_host[t]; a.row_floats[t] = row_floats_host[t];
    }
    a.slot = slot; a.k = k; a.n = n;
    rows_scatter_kernel<<<dim3(k, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return mm::check_launch();
}
