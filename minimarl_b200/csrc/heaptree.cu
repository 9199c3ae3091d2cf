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

// k SumTree.add calls in one launch.  An arg-min index tree with the shape of the sum tree (2*capacity-1 nodes) and a
// copy of the leaf values live in shared memory; thread 0 then replays the adds strictly in order: slot = n_data while
// the buffer fills, else the current arg-min leaf (lowest slot on ties, like heaptree_argmin_kernel); the leaf CHANGE is
// propagated through the float64 sums in the reference's order; the arg-min path of the slot is repaired.
__device__ __forceinline__ int argmin_pick(const double* vals, int a, int b) {
    // a, b: slot indices; lower value wins, lower slot on ties
    const double va = vals[a], vb = vals[b];
    return (vb < va || (vb == va && b < a)) ? b : a;
}

__global__ void heaptree_add_batch_kernel(double* __restrict__ tree, int capacity, int n_data,
                                          const double* __restrict__ priority, int k, int64_t* __restrict__ slot_out) {
    extern __shared__ __align__(8) unsigned char smem_raw[];
    double* vals = reinterpret_cast<double*>(smem_raw);                   // [capacity] leaf priorities
    int* arg = reinterpret_cast<int*>(vals + capacity);                     // [2*capacity-1] arg-min slot per node
    const bool evicts = n_data + k > capacity;
    if (evicts) {
        for (int i = threadIdx.x; i < capacity; i += blockDim.x) {
            vals[i] = tree[capacity - 1 + i];
            arg[capacity - 1 + i] = i;
        }
        __syncthreads();
        // internal nodes bottom-up: node j has children 2j+1, 2j+2 > j, so levels are processed deepest first
        int hi = capacity - 2;                                              // last internal node
        while (hi >= 0) {
            const int lo = (hi + 1) / 2;                                    // nodes [lo, hi] have all children > hi
            for (int j = lo + threadIdx.x; j <= hi; j += blockDim.x) arg[j] = argmin_pick(vals, arg[2 * j + 1], arg[2 * j + 2]);
            __syncthreads();
            hi = lo - 1;
        }
    }
    if (threadIdx.x != 0) return;
    for (int i = 0; i < k; ++i) {
        int slot;
        if (n_data < capacity) slot = n_data++;
        else slot = arg[0];
        const double p = priority[i];
        int idx = capacity - 1 + slot;
        const double change = p - tree[idx];
        tree[idx] = p;
        while (idx != 0) {
            idx = (idx - 1) / 2;
            tree[idx] += change;
        }
        slot_out[i] = slot;
        if (evicts) {
            vals[slot] = p;
            int j = capacity - 1 + slot;
            while (j != 0) {
                j = (j - 1) / 2;
                arg[j] = argmin_pick(vals, arg[2 * j + 1], arg[2 * j + 2]);
            }
        }
    }
}

}  // namespace

extern "C" int mm_heaptree_update(double* tree, int32_t capacity, const int64_t* node_idx, const double* priority,
                                  int32_t n, void* stream) {
    if (!tree || !node_idx || !priority || capacity <= 0 || n <= 0) return MM_ERR_BAD_ARG;
    heaptree_update_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(tree, node_idx, priority, n);
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
    const size_t smem = (size_t)capacity * sizeof(double) + (size_t)(2 * capacity - 1) * sizeof(int);
    if (smem > 227 * 1024) return MM_ERR_UNSUPPORTED;           // capacity > ~14 500: insert one by one
    cudaError_t e = cudaFuncSetAttribute(heaptree_add_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    heaptree_add_batch_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(tree, capacity, n_data, priority, k, slot_out);
    return mm::check_launch();
}
