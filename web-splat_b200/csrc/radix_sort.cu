// radix_sort.cu -- stage 2: onesweep LSD radix sort of (u32 key, u32 value) pairs on sm_100a.
//
// Replaces the reference's Fuchsia-derived WGSL sort (radix_sort.wgsl:48-512, driven by
// gpu_rs.rs:764-863): same contract -- stable, ascending, 8-bit digits, result back in the
// input buffers after an even number of passes -- but one kernel per digit pass:
//   * persistent CTAs take 4096-pair partitions from an atomic ticket (forward progress
//     for the decoupled look-back does not depend on the hardware's CTA scheduling order,
//     which radix_sort.wgsl:365-387 silently relies on);
//   * ranking uses the warp MATCH instruction (one __match_any_sync per key) instead of
//     the reference's per-key shared-memory loop emulation (radix_sort.wgsl:283-302);
//   * 256 threads each run the decoupled look-back of one digit bin, on single 32-bit
//     status words (2-bit flag | 30-bit count) so relaxed accesses are sufficient;
//   * keys/values are reordered in shared memory and leave as per-bin contiguous runs.
// The digit histograms are produced by the kernels that generate the keys (preprocess /
// binning), so a pass reads each pair exactly once: 16 B of HBM traffic per pair per pass.
#include "ws_device.cuh"
#include "ws_kernels.h"

namespace ws {

namespace {

constexpr int WARPS = SORT_THREADS / 32;

__global__ void __launch_bounds__(SORT_THREADS, 3)
onesweep_pass_kernel(SortPassArgs a)
{
    __shared__ uint32_t s_keys[SORT_PART];
    __shared__ uint32_t s_vals[SORT_PART];
    __shared__ uint32_t s_whist[WARPS][256];
    __shared__ uint32_t s_binstart[256];      // first local position of each bin in the partition
    __shared__ uint32_t s_gbase[256];         // global position of local position 0 of each bin, minus binstart
    __shared__ uint32_t s_scan[WARPS];
    __shared__ uint32_t s_part;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t n = *a.n_ptr;
    if (n > a.n_cap) n = a.n_cap;
    const uint32_t nparts = (n + SORT_PART - 1u) / SORT_PART;
    const uint32_t shift = a.shift;

    // exclusive scan of the global digit histogram: thread tid owns bin tid
    uint32_t g_excl;
    {
        const uint32_t c = a.hist[tid];
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) if (w < (int)warp) woff += s_scan[w];
        g_excl = woff + incl - c;
        __syncthreads();
    }

    for (;;) {
        if (tid == 0) s_part = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t part = s_part;
        if (part >= nparts) break;
        const uint32_t base = part * SORT_PART;
        const uint32_t nvalid = (n - base < (uint32_t)SORT_PART) ? (n - base) : (uint32_t)SORT_PART;

        // ---- load, warp-striped: warp w owns [base + w*512, +512), item i of lane l = i*32 + l
        uint32_t key[SORT_ITEMS], val[SORT_ITEMS];
        const uint32_t wbase = warp * (32u * SORT_ITEMS);
        if (nvalid == (uint32_t)SORT_PART) {
#pragma unroll
            for (int i = 0; i < SORT_ITEMS; i++) key[i] = a.keys_in[base + wbase + i * 32u + lane];
#pragma unroll
            for (int i = 0; i < SORT_ITEMS; i++) val[i] = a.vals_in[base + wbase + i * 32u + lane];
        } else {
#pragma unroll
            for (int i = 0; i < SORT_ITEMS; i++) {
                const uint32_t li = wbase + i * 32u + lane;
                key[i] = (li < nvalid) ? a.keys_in[base + li] : 0xffffffffu;   // pads rank last (radix_sort.wgsl:79)
                val[i] = (li < nvalid) ? a.vals_in[base + li] : 0u;
            }
        }
#pragma unroll
        for (int i = 0; i < WARPS; i++) s_whist[i][tid] = 0u;
        __syncthreads();

        // ---- rank inside the warp: match peers with the same digit, leader bumps the warp counter
        uint32_t rank[SORT_ITEMS];
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
            const uint32_t d = (key[i] >> shift) & 255u;
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if ((int)lane == leader) old = atomicAdd(&s_whist[warp][d], (uint32_t)__popc(peers));
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[i] = old + __popc(peers & lanemask_lt());
        }
        __syncthreads();

        // ---- per bin (thread tid = bin): scan over warps, publish, look back
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            const uint32_t c = s_whist[w][tid];
            s_whist[w][tid] = total;
            total += c;
        }
        uint32_t *st = a.status + (size_t)part * 256u + tid;
        if (part == 0u) st_relaxed(st, LB_PREFIX | total);
        else st_relaxed(st, LB_AGGREGATE | total);

        // block exclusive scan of `total` over bins -> first local position of each bin
        uint32_t binstart;
        {
            uint32_t incl = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += t;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            uint32_t woff = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) if (w < (int)warp) woff += s_scan[w];
            binstart = woff + incl - total;
        }

        uint32_t excl = 0;
        if (part > 0u) {
            int64_t p = (int64_t)part - 1;
            uint32_t spins = 0;
            for (;;) {
                const uint32_t s = ld_relaxed(a.status + (size_t)p * 256u + tid);
                const uint32_t f = s >> LB_FLAG_SHIFT;
                if (f == 0u) {                              // predecessor not published yet: spin
                    if (++spins > SPIN_LIMIT) { if (a.err) atomicOr(a.err, 1u); break; }
                    continue;
                }
                excl += s & LB_VALUE_MASK;
                if (f == 2u) break;
                --p;
            }
            st_relaxed(st, LB_PREFIX | (excl + total));
        }
        s_binstart[tid] = binstart;
        s_gbase[tid] = g_excl + excl - binstart;            // wraps mod 2^32 by design
        __syncthreads();

        // ---- reorder in shared memory
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; i++) {
            const uint32_t d = (key[i] >> shift) & 255u;
            const uint32_t pos = s_binstart[d] + s_whist[warp][d] + rank[i];
            s_keys[pos] = key[i];
            s_vals[pos] = val[i];
        }
        __syncthreads();

        // ---- write out: consecutive threads write consecutive addresses within a bin run
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; k++) {
            const uint32_t i = tid + k * SORT_THREADS;
            if (i < nvalid) {
                const uint32_t kk = s_keys[i];
                const uint32_t g = s_gbase[(kk >> shift) & 255u] + i;
                a.keys_out[g] = kk;
                a.vals_out[g] = s_vals[i];
            }
        }
        __syncthreads();
    }
}

// Standalone digit histograms (only for the public ws_sort_pairs_u32 entry point; the frame
// path gets its histograms from preprocess / binning).
__global__ void __launch_bounds__(256)
sort_histogram_kernel(const uint32_t *__restrict__ keys, const uint32_t *n_ptr, uint32_t n_cap,
                      uint32_t *hist, int passes)
{
    __shared__ uint32_t s_hist[4 * 256];
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    for (unsigned i = tid; i < 4u * 256u; i += 256u) s_hist[i] = 0u;
    __syncthreads();
    uint32_t n = *n_ptr;
    if (n > n_cap) n = n_cap;
    const uint32_t nround = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * 256u + tid; i < nround; i += gridDim.x * 256u) {
        const bool ok = i < n;
        const uint32_t k = ok ? keys[i] : 0u;
        for (int d = 0; d < passes; d++) {
            const uint32_t dig = (k >> (8 * d)) & 255u;
            const unsigned peers = __match_any_sync(0xffffffffu, ok ? dig : 0xffffffffu);
            if (ok && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&s_hist[d * 256 + dig], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (unsigned i = tid; i < (unsigned)passes * 256u; i += 256u) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(hist + i, c);
    }
}

}  // namespace

cudaError_t launch_sort_pass(const SortPassArgs &a, int grid, cudaStream_t stream)
{
    onesweep_pass_kernel<<<grid, SORT_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

int sort_pass_blocks_per_sm()
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, onesweep_pass_kernel, SORT_THREADS, 0);
    return nb > 0 ? nb : 1;
}

cudaError_t launch_sort_histogram(const uint32_t *keys, const uint32_t *n_ptr, uint32_t n_cap,
                                  uint32_t *hist, int passes, int grid, cudaStream_t stream)
{
    sort_histogram_kernel<<<grid, 256, 0, stream>>>(keys, n_ptr, n_cap, hist, passes);
    return cudaGetLastError();
}

}  // namespace ws
