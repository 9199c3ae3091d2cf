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
                                                                 SetHeader* __restrict__ hdr, int* __restrict__ rank) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    for (int i = tid; i < capacity; i += kScanThreads) rank[i] = 0;
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

// rank of every old leaf among the old leaves by (priority, slot): block (x, y) counts, for 256 leaves, how many of the 1 024
// leaves of slice y sort before each of them, and adds that to rank[]
__global__ void __launch_bounds__(256) set_rank_kernel(const double* __restrict__ tree, int capacity,
                                                       const SetHeader* __restrict__ hdr, int* __restrict__ rank) {
    __shared__ unsigned long long keys[1024];
    const int n_old = hdr->n_old;
    if (hdr->evict == 0) return;
    const int base = blockIdx.y * 1024;
    if (base >= n_old || blockIdx.x * blockDim.x >= n_old) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long* leaf = reinterpret_cast<const unsigned long long*>(tree + capacity - 1);
    // priorities are non-negative doubles: their bit patterns order like the values
    const unsigned long long mine = i < n_old ? leaf[i] : 0ull;
    const int m = min(1024, n_old - base);
    for (int t = threadIdx.x; t < m; t += blockDim.x) keys[t] = leaf[base + t];
    __syncthreads();
    if (i >= n_old) return;
    int r = 0;
#pragma unroll 4
    for (int t = 0; t < m; ++t) {
        const unsigned long long kt = keys[t];
        r += (kt < mine) || (kt == mine && base + t < i);
    }
    if (r) atomicAdd(rank + i, r);
}
__global__ void set_publish_kernel(const SetHeader* __restrict__ hdr, const int* __restrict__ rank, int* __restrict__ evict_slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hdr->n_old && rank[i] < hdr->evict) evict_slot[rank[i]] = i;
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

// workspace: int32[k + 2 * capacity + 4]
extern "C" int mm_heaptree_add_set(double* tree, int32_t capacity, int32_t* n_data_dev, const double* priority,
                                   const uint8_t* valid, int32_t k, int64_t* slot_out, int32_t* workspace, int32_t rebuild,
                                   void* stream) {
    if (!tree || !n_data_dev || !priority || !slot_out || !workspace || capacity <= 0 || k <= 0) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SetHeader* hdr = reinterpret_cast<SetHeader*>(workspace);
    int* vrank = workspace + 4;
    int* evict_slot = vrank + k;
    int* rank = evict_slot + capacity;
    set_scan_kernel<<<1, kScanThreads, 0, st>>>(valid, k, capacity, n_data_dev, vrank, hdr, rank);
    set_rank_kernel<<<dim3((capacity + 255) / 256, (capacity + 1023) / 1024), 256, 0, st>>>(tree, capacity, hdr, rank);
    set_publish_kernel<<<(capacity + 255) / 256, 256, 0, st>>>(hdr, rank, evict_slot);
    set_assign_kernel<<<(k + 255) / 256, 256, 0, st>>>(tree, capacity, hdr, vrank, evict_slot, priority, k, slot_out);
    if (rebuild) rebuild_kernel<<<1, 1024, 0, st>>>(tree, capacity);
    return mm::check_launch();
}

extern "C" int mm_heaptree_rebuild(double* tree, int32_t capacity, void* stream) {
    if (!tree || capacity <= 0) return MM_ERR_BAD_ARG;
    rebuild_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(tree, capacity);
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

// A whole episode of N lockstep envs: C consecutive mm_heaptree_add_set calls (chunk c of every env, c = 0..C-1, in the
// order the reference adds them) and the payload scatter of each, enqueued from ONE host call; inner nodes recomputed once at
// the end.  priority / valid / slot_out: [C][N]; src tensors are chunk-major [C][N][row_floats].  workspace: int32[N + 2 * capacity + 4].
extern "C" int mm_chunk_per_insert_episode(double* tree, int32_t capacity, int32_t* n_data_dev, const double* priority,
                                           const uint8_t* valid, int32_t C, int32_t N, int64_t* slot_out, int32_t* workspace,
                                           const float* const* src_host, float* const* dst_host, const int32_t* row_floats_host,
                                           int32_t n_tensors, void* stream) {
    if (!tree || !n_data_dev || !priority || !slot_out || !workspace || capacity <= 0 || C <= 0 || N <= 0) return MM_ERR_BAD_ARG;
    if (!src_host || !dst_host || !row_floats_host || n_tensors <= 0 || n_tensors > 8) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SetHeader* hdr = reinterpret_cast<SetHeader*>(workspace);
    int* vrank = workspace + 4;
    int* evict_slot = vrank + N;
    int* rank = evict_slot + capacity;
    ScatterArgs sc{};
    for (int t = 0; t < n_tensors; ++t) {
        if (!src_host[t] || !dst_host[t] || row_floats_host[t] <= 0) return MM_ERR_BAD_ARG;
        sc.dst[t] = dst_host[t]; sc.row_floats[t] = row_floats_host[t];
    }
    sc.k = N; sc.n = n_tensors;
    for (int c = 0; c < C; ++c) {
        const double* pr = priority + (size_t)c * N;
        const uint8_t* va = valid ? valid + (size_t)c * N : nullptr;
        int64_t* so = slot_out + (size_t)c * N;
        set_scan_kernel<<<1, kScanThreads, 0, st>>>(va, N, capacity, n_data_dev, vrank, hdr, rank);
        set_rank_kernel<<<dim3((capacity + 255) / 256, (capacity + 1023) / 1024), 256, 0, st>>>(tree, capacity, hdr, rank);
        set_publish_kernel<<<(capacity + 255) / 256, 256, 0, st>>>(hdr, rank, evict_slot);
        set_assign_kernel<<<(N + 255) / 256, 256, 0, st>>>(tree, capacity, hdr, vrank, evict_slot, pr, N, so);
        for (int t = 0; t < n_tensors; ++t) sc.src[t] = src_host[t] + (size_t)c * N * row_floats_host[t];
        sc.slot = so;
        rows_scatter_kernel<<<dim3(N, n_tensors), 256, 0, st>>>(sc);
    }
    rebuild_kernel<<<1, 1024, 0, st>>>(tree, capacity);
    return mm::check_launch();
}
