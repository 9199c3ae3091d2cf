// Chunk-PER sum tree of the top-level vdn/ and qmix/ packages on the device.
//
// Replaces SumTree (vdn/replay_buffer/sumtree.py:8-66; qmix/replay_buffer/sumtree.py:8-72): a heap-ordered
// float64 array of 2*capacity-1 nodes, root at 0, leaf i at capacity-1+i, capacity NOT required to be a power
// of two.  `update` propagates the CHANGE of a leaf to its ancestors one update after the other (the
// reference's floating-point order, so sums stay bit-identical), `sample` is the stratified draw of
// Prioritized_Experience_Replay.sample (vdn/replay_buffer/buffer.py:42-66) with the root-to-leaf walk of
// `_retrieve_max`, `argmin_leaf` is the eviction rule of `add` once the buffer is full.  All latency-bound.
#include "common.cuh"

namespace {

__global__ void heaptree_update_kernel(double* __restrict__ tree, const int64_t* __restrict__ node_idx,
                                       const double* __restrict__ priority, int n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = 0; i < n; ++i) {                 // strictly in order: later updates see earlier ones
        int64_t idx = node_idx[i];
        const double change = priority[i] - tree[idx];
        tree[idx] = priority[i];
        while (idx != 0) {
            idx = (idx - 1) / 2;
            tree[idx] += change;
        }
    }
}

// n leaf updates applied strictly in order, in shared memory.  Phase 1 (thread 0): change_i = p_i - leaf, leaf = p_i -- a
// chain only through repeated leaves.  Phase 2: thread d owns the tree nodes of depth d (they are disjoint from every other
// depth) and adds change_i to the depth-d ancestor of leaf i, for i = 0..n-1 in order: every node receives exactly the
// additions of the reference's _propagate, in the reference's order, so the float64 sums are bit-identical -- but the ~14
// levels proceed in parallel and each read-modify-write costs a shared-memory round trip instead of an L2 one
// (4 096 updates: 27 ms with the one-thread global-memory walk, ~0.2 ms here).
__device__ __forceinline__ void propagate_by_depth(double* s_tree, const int* s_node, const double* s_change, int n, int tid) {
    // s_node[i]: 1-based heap index of the leaf of update i; depth(x) = floor(log2(x))
    if (tid >= 31) return;
    const int depth = tid;
    for (int i = 0; i < n; ++i) {
        const int x = s_node[i], d = 31 - __clz(x);
        if (depth < d) s_tree[(x >> (d - depth)) - 1] += s_change[i];
    }
}

constexpr int kSubBatch = 1024;      // updates staged per pass: 12 KB of shared memory beside the tree

__global__ void heaptree_update_smem_kernel(double* __restrict__ tree, int n_nodes, const int64_t* __restrict__ node_idx,
                                            const double* __restrict__ priority, int n) {
    extern __shared__ __align__(8) unsigned char smem_raw[];
    double* s_tree = reinterpret_cast<double*>(smem_raw);                 // [n_nodes]
    double* s_change = s_tree + n_nodes;                                  // [kSubBatch]
    int* s_node = reinterpret_cast<int*>(s_change + kSubBatch);           // [kSubBatch]
    for (int i = threadIdx.x; i < n_nodes; i += blockDim.x) s_tree[i] = tree[i];
    for (int base = 0; base < n; base += kSubBatch) {
        const int m = min(kSubBatch, n - base);
        for (int i = threadIdx.x; i < m; i += blockDim.x) s_node[i] = (int)node_idx[base + i] + 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < m; ++i) {
                const int leaf = s_node[i] - 1;
                const double p = priority[base + i];
                s_change[i] = p - s_tree[leaf];
                s_tree[leaf] = p;
            }
        }
        __syncthreads();
        propagate_by_depth(s_tree, s_node, s_change, m, threadIdx.x);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n_nodes; i += blockDim.x) tree[i] = s_tree[i];
}

__global__ void heaptree_sample_kernel(const double* __restrict__ tree, int n_nodes, const double* __restrict__ u,
                                       int n, int64_t* __restrict__ node_out, double* __restrict__ prio_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double segment = tree[0] / (double)n;
    const double left_edge = segment * (double)i, right_edge = segment * (double)(i + 1);
    double s = left_edge + (right_edge - left_edge) * u[i];      // random.uniform(a, b) = a + (b - a) * random()
    int64_t idx = 0;
    for (;;) {
        const int64_t left = 2 * idx + 1;
        if (left >= n_nodes) break;
        const double lv = tree[left];
        if (s <= lv) idx = left;
        else { idx = left + 1; s -= lv; }
    }
    node_out[i] = idx;
    prio_out[i] = tree[idx];
}

__global__ void heaptree_argmin_kernel(const double* __restrict__ tree, int capacity, int64_t* __restrict__ out) {
    // one block; first index of the minimum over the leaves
    __shared__ double s_val[256];
    __shared__ int s_idx[256];
    double best = INFINITY;
    int best_i = 0x7fffffff;
    for (int i = threadIdx.x; i < capacity; i += blockDim.x) {
        const double v = tree[capacity - 1 + i];
        if (v < best) { best = v; best_i = i; }
    }
    s_val[threadIdx.x] = best; s_idx[threadIdx.x] = best_i;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            const double v = s_val[threadIdx.x + off];
            const int j = s_idx[threadIdx.x + off];
            if (v < s_val[threadIdx.x] || (v == s_val[threadIdx.x] && j < s_idx[threadIdx.x])) {
                s_val[threadIdx.x] = v; s_idx[threadIdx.x] = j;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s_idx[0];
}

// k SumTree.add calls in one launch.  The sum tree lives in shared memory; warp 0 replays the adds strictly in order:
// slot = n_data while the buffer fills, else the current arg-min leaf (lowest slot on ties, like heaptree_argmin_kernel).
// The arg-min comes from a 32-ary tournament over the leaves -- group minima of 32 slots, of 32 groups, then the root --
// so that repairing the path of a replaced leaf is three warp-wide reductions instead of a 14-level pointer walk by one
// thread: doubles are mapped to order-preserving 64-bit keys and reduced with redux.sync (hi word, lo word, slot).
// The leaf CHANGES are recorded and pushed through the float64 sums by propagate_by_depth, in the reference's order.
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
// warp-wide arg-min over (key, slot): smallest key, lowest slot on ties; lanes without an entry pass key = ~0, slot = ~0
__device__ __forceinline__ void warp_argmin(unsigned long long key, unsigned slot, unsigned long long& key_out, unsigned& slot_out) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(MM_FULL_MASK, hi);
    const unsigned mlo = __reduce_min_sync(MM_FULL_MASK, hi == mhi ? lo : 0xffffffffu);
    const bool is_min = hi == mhi && lo == mlo;
    slot_out = __reduce_min_sync(MM_FULL_MASK, is_min ? slot : 0xffffffffu);
    key_out = ((unsigned long long)mhi << 32) | mlo;
}

__global__ void heaptree_add_batch_kernel(double* __restrict__ tree, int capacity, int n_data,
                                          const double* __restrict__ priority, int k, int64_t* __restrict__ slot_out) {
    extern __shared__ __align__(8) unsigned char smem_raw[];
    const int n_nodes = 2 * capacity - 1;
    const int G1 = (capacity + 31) / 32, G2 = (G1 + 31) / 32;                // G2 <= 32 for capacity <= 32 768
    double* s_tree = reinterpret_cast<double*>(smem_raw);                   // [2*capacity-1] the sum tree; leaves = priorities
    double* s_change = s_tree + n_nodes;                                    // [kSubBatch]
    unsigned long long* g1k = reinterpret_cast<unsigned long long*>(s_change + kSubBatch);   // [G1] min key of slots 32g..32g+31
    unsigned long long* g2k = g1k + G1;                                     // [G2] min key of groups 32h..32h+31
    int* s_node = reinterpret_cast<int*>(g2k + G2);                         // [kSubBatch] 1-based leaf node of item i
    unsigned* g1s = reinterpret_cast<unsigned*>(s_node + kSubBatch);        // [G1] arg-min slot of the group
    unsigned* g2s = g1s + G1;                                               // [G2]
    const double* vals = s_tree + (capacity - 1);                           // leaf priorities by slot
    const bool evicts = n_data + k > capacity;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < n_nodes; i += blockDim.x) s_tree[i] = tree[i];
    __syncthreads();
    if (evicts) {
        for (int g = warp; g < G1; g += n_warps) {
            const int sl = 32 * g + lane;
            unsigned long long key; unsigned slot;
            warp_argmin(sl < capacity ? order_key(vals[sl]) : ~0ull, sl < capacity ? (unsigned)sl : 0xffffffffu, key, slot);
            if (lane == 0) { g1k[g] = key; g1s[g] = slot; }
        }
        __syncthreads();
        for (int h = warp; h < G2; h += n_warps) {
            const int g = 32 * h + lane;
            unsigned long long key; unsigned slot;
            warp_argmin(g < G1 ? g1k[g] : ~0ull, g < G1 ? g1s[g] : 0xffffffffu, key, slot);
            if (lane == 0) { g2k[h] = key; g2s[h] = slot; }
        }
        __syncthreads();
    }
    for (int base = 0; base < k; base += kSubBatch) {
        const int m = min(kSubBatch, k - base);
        if (warp == 0) {
            for (int i = 0; i < m; ++i) {
                int slot;
                if (n_data < capacity) {
                    slot = n_data++;
                } else {
                    unsigned long long key; unsigned sl;
                    warp_argmin(lane < G2 ? g2k[lane] : ~0ull, lane < G2 ? g2s[lane] : 0xffffffffu, key, sl);
                    slot = (int)sl;
                }
                const double p = priority[base + i];
                const int leaf = capacity - 1 + slot;
                if (lane == 0) {
                    s_change[i] = p - s_tree[leaf];
                    s_tree[leaf] = p;                                       // vals[slot]
                    s_node[i] = leaf + 1;
                    slot_out[base + i] = slot;
                }
                __syncwarp();
                if (evicts) {                                               // repair the two group levels above the slot
                    const int g = slot >> 5, h = g >> 5;
                    unsigned long long key; unsigned sl;
                    const int s1 = 32 * g + lane;
                    warp_argmin(s1 < capacity ? order_key(vals[s1]) : ~0ull, s1 < capacity ? (unsigned)s1 : 0xffffffffu, key, sl);
                    if (lane == 0) { g1k[g] = key; g1s[g] = sl; }
                    __syncwarp();
                    const int s2 = 32 * h + lane;
                    warp_argmin(s2 < G1 ? g1k[s2] : ~0ull, s2 < G1 ? g1s[s2] : 0xffffffffu, key, sl);
                    if (lane == 0) { g2k[h] = key; g2s[h] = sl; }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        propagate_by_depth(s_tree, s_node, s_change, m, threadIdx.x);       // the float64 sums, in the reference's order
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n_nodes; i += blockDim.x) tree[i] = s_tree[i];
}

}  // namespace

extern "C" int mm_heaptree_update(double* tree, int32_t capacity, const int64_t* node_idx, const double* priority,
                                  int32_t n, void* stream) {
    if (!tree || !node_idx || !priority || capacity <= 0 || n <= 0) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)(2 * capacity - 1) * sizeof(double) + (size_t)kSubBatch * (sizeof(double) + sizeof(int)) + 16;
    if (n >= 8 && smem <= 227 * 1024) {       // batched updates with the tree in shared memory, levels in parallel
        cudaError_t e = cudaFuncSetAttribute(heaptree_update_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
        heaptree_update_smem_kernel<<<1, 1024, smem, st>>>(tree, 2 * capacity - 1, node_idx, priority, n);
    } else {
        heaptree_update_kernel<<<1, 32, 0, st>>>(tree, node_idx, priority, n);
    }
    return mm::check_launch();
}

extern "C" int mm_heaptree_sample(const double* tree, int32_t capacity, const double* uniforms, int32_t n,
                                  int64_t* node_idx_out, double* priority_out, void* stream) {
    if (!tree || !uniforms || !node_idx_out || !priority_out || capacity <= 0 || n <= 0) return MM_ERR_BAD_ARG;
    heaptree_sample_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(tree, 2 * capacity - 1, uniforms, n,
                                                                                          node_idx_out, priority_out);
    return mm::check_launch();
}

extern "C" int mm_heaptree_argmin_leaf(const double* tree, int32_t capacity, int64_t* leaf_out, void* stream) {
    if (!tree || !leaf_out || capacity <= 0) return MM_ERR_BAD_ARG;
    heaptree_argmin_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(tree, capacity, leaf_out);
    return mm::check_launch();
}

extern "C" int mm_heaptree_add_batch(double* tree, int32_t capacity, int32_t n_data, const double* priority, int32_t k,
                                     int64_t* slot_out, void* stream) {
    if (!tree || !priority || !slot_out || capacity <= 0 || k <= 0 || n_data < 0 || n_data > capacity) return MM_ERR_BAD_ARG;
    const size_t g1 = (capacity + 31) / 32, g2 = (g1 + 31) / 32;
    const size_t smem = (size_t)(2 * capacity - 1) * sizeof(double) + (size_t)kSubBatch * (sizeof(double) + sizeof(int)) +
                        (g1 + g2) * (sizeof(unsigned long long) + sizeof(unsigned)) + 16;
    if (smem > 227 * 1024 || g2 > 32) return MM_ERR_UNSUPPORTED;   // the tree does not fit (capacity > ~14 000): insert one by one
    cudaError_t e = cudaFuncSetAttribute(heaptree_add_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    heaptree_add_batch_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(tree, capacity, n_data, priority, k, slot_out);
    return mm::check_launch();
}
