// Fisher-Yates swap sequence applied in parallel, as plain grid kernels.
//
// numpy's Generator.permutation (_shuffle_raw) runs, for i = n-1 .. 1, swap(a[i], a[j_i]).  The swap targets j_i come
// from qa_perm_resolve (qa_greedy_par.cu).  Position i is final after step i, and what step i puts there is the
// content of position j_i just before the step - which an earlier step (larger i') with the same target wrote, taking
// it from its own position i', and so on: every output is found by following a chain of "who wrote this position
// last" links.  The links come from a counting sort of the steps by target position:
//   count -> exclusive scan -> fill buckets -> sort + link (succ / parent) -> follow.
// Unlike the resolve (a fixed point over one RNG stream, run by one cluster) these phases are embarrassingly parallel,
// so they run on the whole GPU; used for the permutations the greedy draws ahead of time on side streams.  The
// cluster kernel keeps its own in-cluster version (perm_apply) for the permutations it must draw mid-run.
#include "qa_common.cuh"

namespace qa {

constexpr int PA_T = 256;
constexpr int PA_PER = 8;
constexpr int PA_CHUNK = PA_T * PA_PER;       // elements per block in the scan kernels

__global__ void __launch_bounds__(PA_T) pa_count_kernel(const int32_t* __restrict__ jarr, int m, int32_t* cursor) {
    const int i = 1 + blockIdx.x * PA_T + threadIdx.x;
    if (i < m) atomicAdd(&cursor[jarr[i]], 1);
}

__device__ __forceinline__ int pa_block_excl_scan(int v, int& total) {
    __shared__ int wsum[PA_T / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int ws = lane < PA_T / 32 ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < PA_T / 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, ws, o);
        if (lane >= o) ws += y;
    }
    total = __shfl_sync(0xFFFFFFFFu, ws, PA_T / 32 - 1);
    const int before = __shfl_sync(0xFFFFFFFFu, ws, (w + 31) & 31);
    __syncthreads();
    return (w ? before : 0) + inc - v;
}

__global__ void __launch_bounds__(PA_T) pa_block_sums_kernel(const int32_t* __restrict__ cursor, int m, int32_t* bsum) {
    const int base = blockIdx.x * PA_CHUNK;
    int s = 0;
#pragma unroll
    for (int k = 0; k < PA_PER; ++k) {
        const int p = base + k * PA_T + threadIdx.x;
        if (p < m) s += cursor[p];
    }
    int total;
    pa_block_excl_scan(s, total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// in-place exclusive scan of the block sums by one block (nb = ceil(m / 2048): a few hundred entries)
__global__ void __launch_bounds__(PA_T) pa_scan_top_kernel(int32_t* bsum, int nb) {
    int carry = 0;
    for (int base = 0; base < nb; base += PA_T) {
        const int p = base + threadIdx.x;
        const int v = p < nb ? bsum[p] : 0;
        int total;
        const int ex = pa_block_excl_scan(v, total);
        if (p < nb) bsum[p] = carry + ex;
        carry += total;
    }
}

// off[p] = cursor[p] = exclusive prefix of the counts; off[m] = m - 1 (every step sits in exactly one bucket)
__global__ void __launch_bounds__(PA_T) pa_scan_apply_kernel(int32_t* cursor, int m, const int32_t* __restrict__ bsum, int32_t* off) {
    const int first = blockIdx.x * PA_CHUNK + threadIdx.x * PA_PER;       // this thread's PA_PER consecutive positions
    int v[PA_PER], s = 0;
#pragma unroll
    for (int k = 0; k < PA_PER; ++k) {
        v[k] = first + k < m ? cursor[first + k] : 0;
        s += v[k];
    }
    int total;
    int run = bsum[blockIdx.x] + pa_block_excl_scan(s, total);
#pragma unroll
    for (int k = 0; k < PA_PER; ++k) {
        if (first + k < m) { off[first + k] = run; cursor[first + k] = run; }
        run += v[k];
    }
    if (first <= m - 1 && m - 1 < first + PA_PER) off[m] = run;
}

__global__ void __launch_bounds__(PA_T) pa_fill_kernel(const int32_t* __restrict__ jarr, int m, int32_t* cursor, int32_t* bucket) {
    const int i = 1 + blockIdx.x * PA_T + threadIdx.x;
    if (i < m) bucket[atomicAdd(&cursor[jarr[i]], 1)] = i;
}

__global__ void __launch_bounds__(PA_T) pa_link_kernel(int m, const int32_t* __restrict__ off, int32_t* bucket, int32_t* succ,
                                                       int32_t* parent) {
    const int p = blockIdx.x * PA_T + threadIdx.x;
    if (p >= m) return;
    const int b = off[p], e = off[p + 1];
    for (int a = b + 1; a < e; ++a) {          // insertion sort (buckets hold ~1 entry)
        const int key = bucket[a];
        int q = a - 1;
        while (q >= b && bucket[q] > key) { bucket[q + 1] = bucket[q]; --q; }
        bucket[q + 1] = key;
    }
    int par = -1;
    for (int a = b; a < e; ++a) {
        const int st = bucket[a];
        succ[st] = a + 1 < e ? bucket[a + 1] : -1;
        if (par < 0 && st != p) par = st;
    }
    parent[p] = par;       // first later-executed step that writes position p
}

__global__ void __launch_bounds__(PA_T) pa_follow_kernel(int m, const int32_t* __restrict__ jarr, const int32_t* __restrict__ succ,
                                                         const int32_t* __restrict__ parent, const int32_t* __restrict__ cand,
                                                         int32_t* out) {
    const int i = blockIdx.x * PA_T + threadIdx.x;
    if (i >= m) return;
    // out[0] is the content of position 0 after all steps; out[i] (i >= 1) is what step i read from position j_i
    const int start = i == 0 ? parent[0] : succ[i];
    int val;
    if (start < 0) val = i == 0 ? 0 : jarr[i];          // nobody wrote that position: initial content
    else {
        int cur = start;
        for (int pa = parent[cur]; pa >= 0; pa = parent[cur]) cur = pa;
        val = cur;
    }
    out[i] = cand ? cand[val] : val;
}

__global__ void pa_single_kernel(const int32_t* cand, int32_t* out) { out[0] = cand ? cand[0] : 0; }

static inline int64_t al256(int64_t v) { return (v + 255) / 256 * 256; }

}  // namespace qa

using namespace qa;

extern "C" int64_t qa_perm_apply_work_bytes(int64_t n) {
    return 5 * al256(4 * (n + 1)) + al256(4 * (cdiv(n > 0 ? n : 1, PA_CHUNK) + 1)) + 256;
}

extern "C" int qa_perm_apply(const int32_t* jarr, int64_t n, const int32_t* cand, int32_t* out, void* work, qa_stream_t stream) {
    if (n < 0 || n > 0x3FFFFFFF || (n > 0 && !out) || (n > 1 && (!jarr || !work))) { set_error("qa_perm_apply: bad args"); return 1; }
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 1) {
        pa_single_kernel<<<1, 1, 0, s>>>(cand, out);
        return check_launch("qa_perm_apply");
    }
    const int m = (int)n;
    char* p = reinterpret_cast<char*>(work);
    auto take = [&](int64_t bytes) { char* r = p; p += al256(bytes); return reinterpret_cast<int32_t*>(r); };
    int32_t* cursor = take(4 * (n + 1));
    int32_t* off = take(4 * (n + 1));
    int32_t* bucket = take(4 * (n + 1));
    int32_t* succ = take(4 * (n + 1));
    int32_t* parent = take(4 * (n + 1));
    const int nb = (int)cdiv(n, PA_CHUNK);
    int32_t* bsum = take(4 * (nb + 1));
    const unsigned g_el = (unsigned)cdiv(n, PA_T);
    if (cudaMemsetAsync(cursor, 0, 4 * (size_t)n, s) != cudaSuccess) return check_launch("qa_perm_apply (memset)");
    pa_count_kernel<<<g_el, PA_T, 0, s>>>(jarr, m, cursor);
    pa_block_sums_kernel<<<nb, PA_T, 0, s>>>(cursor, m, bsum);
    pa_scan_top_kernel<<<1, PA_T, 0, s>>>(bsum, nb);
    pa_scan_apply_kernel<<<nb, PA_T, 0, s>>>(cursor, m, bsum, off);
    pa_fill_kernel<<<g_el, PA_T, 0, s>>>(jarr, m, cursor, bucket);
    pa_link_kernel<<<g_el, PA_T, 0, s>>>(m, off, bucket, succ, parent);
    pa_follow_kernel<<<g_el, PA_T, 0, s>>>(m, jarr, succ, parent, cand, out);
    return check_launch("qa_perm_apply");
}
