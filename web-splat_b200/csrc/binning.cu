// binning.cu -- tile binning between the depth sort and the tile sort (new design; the
// reference has no tiles: it draws every splat as one instanced quad, gaussian.wgsl:30-57).
//
// The logical sort key of the tile-binned renderer is 64-bit (tile | depth).  An LSD radix
// sort of that key is "sort by the depth digits, then stably by the tile digits"; the depth
// is a per-SPLAT property, so its four digit passes run on the V visible splats BEFORE they
// are expanded into P >= V (tile, splat) pairs.  This kernel does the expansion: it walks
// the splats in depth order, turns each tile rectangle into its (tile id, slot) pairs at
// offsets given by a single-pass decoupled look-back scan, and counts the tile-id digits
// for the two onesweep passes that follow.  The result after those passes is bit-identical
// to sorting P 64-bit (tile|depth) keys with (tile, depth, slot) order, at ~1/3 of the traffic.
#include "ws_device.cuh"
#include "ws_kernels.h"

namespace ws {

namespace {

constexpr int BIN_THREADS = 256;
constexpr int BIN_WARPS = BIN_THREADS / 32;

__global__ void __launch_bounds__(BIN_THREADS, 4)
binning_kernel(BinningArgs a)
{
    __shared__ uint32_t s_incl[BIN_THREADS];      // inclusive scan of pair counts in the partition
    __shared__ uint32_t s_xy[BIN_THREADS];        // x0 | y0 << 16
    __shared__ uint32_t s_w[BIN_THREADS];         // rect width
    __shared__ uint32_t s_slot[BIN_THREADS];
    __shared__ uint32_t s_hist[4 * 256];
    __shared__ uint32_t s_scan[BIN_WARPS];
    __shared__ uint32_t s_part, s_base;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t V = a.counters->num_visible;
    const uint32_t nparts = (V + BIN_THREADS - 1u) / BIN_THREADS;
    const uint32_t tiles_x = a.uniforms->tiles_x;
    const uint32_t cap = a.uniforms->pair_capacity;
    const uint32_t ntiles = tiles_x * a.uniforms->tiles_y;
    const int ndig = (ntiles > 65536u) ? 3 : ((ntiles > 256u) ? 2 : 1);

    for (unsigned i = tid; i < 4u * 256u; i += BIN_THREADS) s_hist[i] = 0u;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_part = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t part = s_part;
        if (part >= nparts) break;

        const uint32_t i = part * BIN_THREADS + tid;
        uint32_t cnt = 0, slot = 0, xy = 0, w = 0;
        if (i < V) {
            slot = a.sorted_slots[i];
            const uint2 r = a.rects[slot];
            xy = r.x;
            w = r.y & 0xffffu;
            cnt = w * (r.y >> 16);
        }
        // block inclusive scan of cnt
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int k = 0; k < BIN_WARPS; k++) {
            const uint32_t c = s_scan[k];
            if (k < (int)warp) woff += c;
            total += c;
        }
        incl += woff;
        s_incl[tid] = incl; s_xy[tid] = xy; s_w[tid] = w; s_slot[tid] = slot;

        if (warp == 0) {
            // partition totals can exceed 30 bits only if P does; P is capped far below
            const uint32_t tot30 = total & LB_VALUE_MASK;
            if (lane == 0 && part > 0u) st_relaxed(a.scan_status + part, LB_AGGREGATE | tot30);
            uint32_t excl = (part > 0u) ? lookback_warp(a.scan_status, part, &a.counters->error_flags) : 0u;
            if (lane == 0) {
                st_relaxed(a.scan_status + part, LB_PREFIX | ((excl + tot30) & LB_VALUE_MASK));
                s_base = excl;
                if (part == nparts - 1u) {
                    const uint32_t P = excl + total;
                    a.counters->num_pairs = P;
                    a.counters->pair_overflow = (P > cap) ? 1u : 0u;
                }
            }
        }
        __syncthreads();
        const uint32_t base = s_base;

        // ---- load-balanced expansion: pair q of the partition belongs to the splat `owner`
        //      with s_incl[owner-1] <= q < s_incl[owner]
        const uint32_t total_r = (total + 31u) & ~31u;       // keep warps converged for match_any
        for (uint32_t q = tid; q < total_r; q += BIN_THREADS) {
            const bool ok = q < total;
            uint32_t tile = 0;
            if (ok) {
                int lo = 0, hi = BIN_THREADS - 1;            // first index with s_incl > q
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_incl[mid] > q) hi = mid; else lo = mid + 1;
                }
                const uint32_t ow = s_w[lo];
                const uint32_t start = (lo > 0) ? s_incl[lo - 1] : 0u;   // first pair of the owner's run
                const uint32_t t = q - start;
                const uint32_t ty = t / ow, tx = t - ty * ow;
                const uint32_t oxy = s_xy[lo];
                tile = ((oxy >> 16) + ty) * tiles_x + (oxy & 0xffffu) + tx;
                const uint64_t g = (uint64_t)base + q;
                if (g < cap) {
                    a.pair_tiles[g] = tile;
                    a.pair_slots[g] = s_slot[lo];
                }
            }
            const bool counted = ok && ((uint64_t)base + q < cap);
            for (int d = 0; d < ndig; d++) {
                const uint32_t dig = (tile >> (8 * d)) & 255u;
                const unsigned peers = __match_any_sync(0xffffffffu, counted ? dig : 0xffffffffu);
                if (counted && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&s_hist[d * 256 + dig], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
    }

    for (unsigned i = tid; i < 4u * 256u; i += BIN_THREADS) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(a.hist + i, c);
    }
}

__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint32_t *__restrict__ pair_tiles, const FrameCounters *counters, uint32_t pair_cap,
                   uint2 *ranges)
{
    uint32_t P = counters->num_pairs;
    if (P > pair_cap) P = pair_cap;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < P; i += gridDim.x * 256u) {
        const uint32_t t = pair_tiles[i];
        if (i == 0u || pair_tiles[i - 1u] != t) ranges[t].x = i;
        if (i == P - 1u || pair_tiles[i + 1u] != t) ranges[t].y = i + 1u;
    }
}

}  // namespace

cudaError_t launch_binning(const BinningArgs &a, int grid, cudaStream_t stream)
{
    binning_kernel<<<grid, BIN_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

int binning_blocks_per_sm()
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, binning_kernel, BIN_THREADS, 0);
    return nb > 0 ? nb : 1;
}

cudaError_t launch_tile_ranges(const uint32_t *pair_tiles, const FrameCounters *counters, uint32_t pair_cap,
                               uint2 *ranges, int grid, cudaStream_t stream)
{
    tile_ranges_kernel<<<grid, 256, 0, stream>>>(pair_tiles, counters, pair_cap, ranges);
    return cudaGetLastError();
}

}  // namespace ws
