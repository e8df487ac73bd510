// topk.cu -- PhpMath::topk (PhpMath.php:933-1058): per row of input[m,n] the k largest values and their indices,
// found with the reference's min-heap so that ties, NaN handling and the output order (heap order when unsorted,
// heap-sort order when sorted) are the reference's, bit for bit.
//
// One warp per row, heap (k values + k indices) in shared memory.  The heap is sequential by nature, the scan is not:
// the 32 lanes load 32 consecutive elements (coalesced) and vote which ones exceed the current heap minimum; only
// those candidates are offered to the heap, by lane 0, in index order, re-tested against the minimum as it rises.
// An element that fails the vote would fail the sequential test too (the minimum never falls), so the sequence of
// heap updates is exactly the reference's.  After the first few hundred elements almost every vote is empty and the
// row streams at load speed.
#include "common.cuh"

namespace {

constexpr int TK_WARPS = 4;
constexpr int TK_UNROLL = 8;

template <typename T>
__device__ __forceinline__ void tk_heapify(int size, T* heap, int* idx, int parent)      // topkMinHeapify, :940-983
{
    int left = 2 * parent + 1, right = 2 * parent + 2;
    while (left < size) {
        const int smallest = (right < size && heap[right] < heap[left]) ? right : left;
        if (heap[parent] <= heap[smallest]) break;
        const T tv = heap[parent]; heap[parent] = heap[smallest]; heap[smallest] = tv;
        const int ti = idx[parent]; idx[parent] = idx[smallest]; idx[smallest] = ti;
        parent = smallest;
        left = 2 * parent + 1;
        right = 2 * parent + 2;
    }
}

template <typename T, typename I>
__global__ void __launch_bounds__(TK_WARPS * 32)
topk_kernel(int64_t m, int n, int k, int sorted, const T* __restrict__ input, T* __restrict__ values, I* __restrict__ indices,
            unsigned char* gheap)
{
    extern __shared__ __align__(16) unsigned char tk_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // heaps in shared memory; for k beyond the shared-memory budget (the reference accepts any k <= n) the same
    // algorithm runs on per-warp heaps in the global workspace (L2-resident for the sizes that get here)
    T* heap;
    int* idx;
    if (gheap == nullptr) {
        heap = reinterpret_cast<T*>(tk_smem) + (size_t)warp * k;
        idx = reinterpret_cast<int*>(reinterpret_cast<T*>(tk_smem) + (size_t)TK_WARPS * k) + (size_t)warp * k;
    } else {
        const size_t gw = (size_t)blockIdx.x * TK_WARPS + warp, nw = (size_t)gridDim.x * TK_WARPS;
        heap = reinterpret_cast<T*>(gheap) + gw * k;
        idx = reinterpret_cast<int*>(reinterpret_cast<T*>(gheap) + nw * k) + gw * k;
    }
    for (int64_t row = (int64_t)blockIdx.x * TK_WARPS + warp; row < m; row += (int64_t)gridDim.x * TK_WARPS) {
        const T* arr = input + row * n;
        for (int i = lane; i < k; i += 32) { heap[i] = arr[i]; idx[i] = i; }
        __syncwarp();
        if (lane == 0) {
            for (int i = k / 2 - 1; i >= 0; --i) tk_heapify<T>(k, heap, idx, i);
        }
        __syncwarp();
        for (int base = k; base < n; base += 32 * TK_UNROLL) {
            T v[TK_UNROLL];                                                    // TK_UNROLL coalesced loads in flight per lane
#pragma unroll
            for (int u = 0; u < TK_UNROLL; u++) {
                const int i = base + u * 32 + lane;
                v[u] = i < n ? arr[i] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < TK_UNROLL; u++) {                              // groups offered in index order
                const int i = base + u * 32 + lane;
                const T top = heap[0];
                unsigned vote = __ballot_sync(0xffffffffu, i < n && v[u] > top);
                while (vote) {                                                 // rare after the heap has warmed up
                    const int src = __ffs(vote) - 1;
                    vote &= vote - 1;
                    const T cv = __shfl_sync(0xffffffffu, v[u], src);
                    if (lane == 0 && cv > heap[0]) {
                        heap[0] = cv;
                        idx[0] = base + u * 32 + src;
                        tk_heapify<T>(k, heap, idx, 0);
                    }
                    __syncwarp();
                }
            }
        }
        if (sorted && lane == 0) {                                         // :1015-1022
            for (int i = k - 1; i > 0; --i) {
                const T tv = heap[0]; heap[0] = heap[i]; heap[i] = tv;
                const int ti = idx[0]; idx[0] = idx[i]; idx[i] = ti;
                tk_heapify<T>(i, heap, idx, 0);
            }
        }
        __syncwarp();
        for (int i = lane; i < k; i += 32) { values[row * k + i] = heap[i]; indices[row * k + i] = (I)idx[i]; }
        __syncwarp();
    }
}

template <typename T, typename I>
int topk_launch(int64_t m, int64_t n, int64_t k, int sorted, const void* input, void* values, void* indices)
{
    rb200::Context& c = rb200::ctx();
    size_t smem = (size_t)TK_WARPS * k * (sizeof(T) + sizeof(int));
    int64_t grid = rb_ceil_div(m, TK_WARPS);
    const int64_t cap = (int64_t)c.sm_count * 16;
    if (grid > cap) grid = cap;
    unsigned char* gheap = nullptr;
    if (smem > 200 * 1024) {
        // heaps in the global workspace: as many warps as 1 GiB of heaps allows
        const int64_t per_cta = (int64_t)smem;
        int64_t fit = ((int64_t)1 << 30) / per_cta;
        if (fit < 1) fit = 1;
        if (grid > fit) grid = fit;
        int rc = rb200::ensure_workspace((size_t)(grid * per_cta));
        if (rc != RB200_OK) return rc;
        gheap = reinterpret_cast<unsigned char*>(c.workspace);
        smem = 0;
    } else if (smem > 48 * 1024) {
        RB_CUDA(cudaFuncSetAttribute(topk_kernel<T, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    topk_kernel<T, I><<<(unsigned)grid, TK_WARPS * 32, smem, c.stream>>>(m, (int)n, (int)k, sorted, reinterpret_cast<const T*>(input),
                                                                        reinterpret_cast<T*>(values), reinterpret_cast<I*>(indices), gheap);
    RB_LAUNCH_CHECK("topk_kernel");
    return RB200_OK;
}

template <typename T>
int topk_index(int index_dtype, int64_t m, int64_t n, int64_t k, int sorted, const void* input, void* values, void* indices)
{
    switch (index_dtype) {
        case RB200_INT32: case RB200_UINT32: return topk_launch<T, int32_t>(m, n, k, sorted, input, values, indices);
        case RB200_INT64: case RB200_UINT64: return topk_launch<T, long long>(m, n, k, sorted, input, values, indices);
        default: RB_UNSUPPORTED("topk: index dtype %d is not a 32- or 64-bit integer", index_dtype);
    }
}

}  // namespace

extern "C" int rb200_topk(int dtype, int index_dtype, int64_t m, int64_t n, const void* input, int64_t offInput, int64_t k,
                          int sorted, void* values, int64_t offValues, void* indices, int64_t offIndices)
{
    RB_REQUIRE_INIT();
    if (!input || !values || !indices) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1 || k < 1 || offInput < 0 || offValues < 0 || offIndices < 0) RB_INVALID("topk: bad shape / offset");
    if (n > 0x7fffffffLL) RB_UNSUPPORTED("topk: n exceeds 2^31-1");
    if (k > n) return RB200_OK;                                            // the reference returns without writing (:1045)
    const int es = rb_dtype_size(dtype), is = rb_dtype_size(index_dtype);
    const char* in = reinterpret_cast<const char*>(input) + offInput * es;
    char* v = reinterpret_cast<char*>(values) + offValues * es;
    char* ix = reinterpret_cast<char*>(indices) + offIndices * is;
    switch (dtype) {
        case RB200_FLOAT32: return topk_index<float>(index_dtype, m, n, k, sorted, in, v, ix);
        case RB200_FLOAT64: return topk_index<double>(index_dtype, m, n, k, sorted, in, v, ix);
        case RB200_INT32:   return topk_index<int32_t>(index_dtype, m, n, k, sorted, in, v, ix);
        case RB200_INT64:   return topk_index<long long>(index_dtype, m, n, k, sorted, in, v, ix);
        default: RB_UNSUPPORTED("topk: unsupported dtype %d", dtype);
    }
}
