// binning.cu -- tile binning between the depth sort and the tile sort (new design; the
// reference has no tiles: it draws every splat as one instanced quad, gaussian.wgsl:30-57).
//
// The logical sort key of the tile-binned renderer is 64-bit (tile | depth).  An LSD radix
// sort of that key is "sort by the depth digits, then stably by the tile digits"; the depth
// is a per-SPLAT property, so its four digit passes run on the V visible splats BEFORE they
// are expanded into P >= V (tile, splat) pairs.  This kernel does the expansion: it walks
// the splats in depth order, turns each tile rectangle into its (tile id, slot) pairs at
// offsets given by a scan of the per-partition pair counts, and counts the tile-id digits
// for the onesweep passes that follow.  The result after those passes is bit-identical
// to sorting P 64-bit (tile|depth) keys with (tile, depth, slot) order, at ~1/3 of the traffic.
//
// Three launches, none with an inter-CTA dependency (a single-pass chained scan was measured
// first, profiles/r01*: at ~900 pairs per partition every partition paid ~11 us of serialised
// ticket / gather / look-back round trips and the kernel ran at 8 % of the HBM roofline):
//   COUNT   pairs per 1024-splat partition (slot load + rectangle gather, 4 per thread in flight);
//   SCAN    one CTA: exclusive scan of the partition totals, P, overflow flag;
//   EXPAND  load-balanced over PAIRS, not splats: every splat with at least one tile drops a
//           marker (its index) at the block-local offset of its first pair, a prefix-max over
//           the positions (warp shuffles, 8 positions per thread) tells every output position
//           which splat owns it, and the thread derives (tile x, tile y) from the position with
//           one multiply-high (magic reciprocal of the rectangle width).  Threads emit
//           consecutive pairs straight to global memory, perfectly coalesced, however uneven
//           the rectangles are (a screen-filling splat simply owns many consecutive positions).
//           (A round-2 rewrite -- owner thread walks its rectangle, the block walks the large ones,
//           pairs staged in shared memory -- is kept as bin_expand_kernel / WS_BIN_EXPAND=2: it
//           executes a quarter of the instructions and is 13 % slower, profiles/r02d_*.)
// The digit histogram uses plain shared-memory atomicAdd (no return value): measured at
// 118-135 G warp-ops/s on B200 whatever the address spread, 25x a MATCH.ANY-aggregated update
// (profiles/microbench/rank_primitives.cu).
#include "ws_device.cuh"
#include "ws_kernels.h"

#include <stdlib.h>

namespace ws {

namespace {

constexpr int BIN_THREADS = 256;
constexpr int BIN_WARPS = BIN_THREADS / 32;
constexpr int BIN_SPT = 4;                            // splats per thread
constexpr int BIN_PART = BIN_THREADS * BIN_SPT;       // 1024 splats per partition
constexpr int BIN_NDIG = 3;

struct SplatRects { uint32_t slot[BIN_SPT], xy[BIN_SPT], w[BIN_SPT], h[BIN_SPT], cnt[BIN_SPT]; };

// the range of depth-sorted positions this launch expands (ascending key = far -> near, so the near slab is the upper one);
// the split point is a multiple of 4 so that the 16-byte slot loads stay aligned
__device__ __forceinline__ void slab_bounds(const BinningArgs &a, uint32_t V, uint32_t &lo, uint32_t &hi)
{
    // the nearest near_pct % of the depth-sorted splats form the near slab (0 = the default, a quarter).  Measured on B200
    // (profiles/r02i_*, frames/s at cfg3 | cfg4): 15 %: 916 | 460, 20 %: 930 | 477, 25 %: 927 | 473, 35 %: 901 | 473,
    // 50 % (round 1): 880 | 458, 65 %: 844 | 442 -- a smaller near slab sorts fewer pairs that the far slab's cull would
    // have dropped; below ~20 % too few tiles are saturated when the far slab is binned.
    const uint32_t near = a.near_pct ? (uint32_t)(((uint64_t)V * a.near_pct) / 100u) : (V >> 2);
    const uint32_t split = (V - near) & ~3u;
    lo = (a.slab == 1u) ? split : 0u;
    hi = (a.slab == 2u) ? split : V;
}

// slots + rectangles of the BIN_SPT consecutive depth-sorted splats of one thread (positions first .. first+3, below V).
// keep4: one byte per splat, 0 = emits no pair (far slab, decided by the count kernel); 0x01010101 when there is no filter.
__device__ __forceinline__ void load_rects(const BinningArgs &a, uint32_t first, uint32_t V, SplatRects &r, uint32_t keep4 = 0x01010101u)
{
    if (first + BIN_SPT <= V) {
        const uint4 s4 = *reinterpret_cast<const uint4 *>(a.sorted_slots + first);
        r.slot[0] = s4.x; r.slot[1] = s4.y; r.slot[2] = s4.z; r.slot[3] = s4.w;
    } else {
#pragma unroll
        for (int j = 0; j < BIN_SPT; j++) r.slot[j] = (first + j < V) ? a.sorted_slots[first + j] : 0xffffffffu;
    }
    uint2 rc[BIN_SPT];
#pragma unroll
    for (int j = 0; j < BIN_SPT; j++)
        rc[j] = (r.slot[j] != 0xffffffffu && ((keep4 >> (8 * j)) & 0xffu)) ? __ldg(a.rects + r.slot[j]) : make_uint2(0u, 0u);
#pragma unroll
    for (int j = 0; j < BIN_SPT; j++) {
        r.xy[j] = rc[j].x;
        r.w[j] = rc[j].y & 0xffffu;
        r.h[j] = rc[j].y >> 16;
        r.cnt[j] = r.w[j] * r.h[j];
    }
}

// far slab: a splat all of whose tiles were saturated by the near slab cannot change a pixel.  Only small rectangles are
// tested (they are almost all of them); a pair emitted for a saturated tile is harmless, the compositor skips that tile.
// Returns the keep bytes for load_rects (the expand kernel does not repeat the test, nor fetch the dropped rectangles).
// `done` is the per-tile byte map (global memory, or its copy in shared memory); `bits`, when not NULL, is the same map as
// one BIT per tile in shared memory: a row of up to 16 tiles is then one funnel-shifted compare instead of a byte per tile.
__device__ __forceinline__ uint32_t drop_saturated(const BinningArgs &a, SplatRects &r, const uint8_t *done, const uint32_t *bits)
{
    const uint32_t tiles_x = a.uniforms->tiles_x;
    uint32_t keep4 = 0u;
#pragma unroll
    for (int j = 0; j < BIN_SPT; j++) {
        if (r.cnt[j] > 0u && r.cnt[j] <= 16u) {
            const uint32_t x0 = r.xy[j] & 0xffffu, y0 = r.xy[j] >> 16;
            bool all = true;
            if (bits) {
                const uint32_t mask = (1u << r.w[j]) - 1u;             // w <= 16
                uint32_t b0 = y0 * tiles_x + x0;
                for (uint32_t yy = 0; yy < r.h[j]; yy++, b0 += tiles_x) {
                    const uint32_t w0 = b0 >> 5;
                    const uint32_t row = __funnelshift_r(bits[w0], bits[w0 + 1u], b0 & 31u);   // bit i = tile b0 + i
                    all = all && ((row & mask) == mask);
                }
            } else {
                for (uint32_t yy = 0; yy < r.h[j]; yy++)
                    for (uint32_t xx = 0; xx < r.w[j]; xx++) all = all && (done[(y0 + yy) * tiles_x + x0 + xx] != 0);
            }
            if (all) r.cnt[j] = 0u;
        }
        if (r.cnt[j] > 0u) keep4 |= 1u << (8 * j);
    }
    return keep4;
}

// ---- (1) COUNT ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(BIN_THREADS)
bin_count_kernel(BinningArgs a)
{
    __shared__ uint32_t s_red[BIN_WARPS];
    extern __shared__ __align__(16) uint8_t s_done[];   // far slab: the per-tile "saturated" bytes (T of them, when they fit)
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t V = a.counters->num_visible;
    uint32_t lo, hi;
    slab_bounds(a, V, lo, hi);
    const uint32_t nparts = (hi - lo + BIN_PART - 1u) / BIN_PART;
    // The saturation test reads the "saturated" flag of every tile of every small rectangle: random byte gathers from
    // global memory in round 1 (r02a: 43 us for the far half of cfg3, long_scoreboard 10 warps per issue).  The whole map is
    // 8160 tiles at 1080p, 32 400 at 4K: each CTA packs it into ONE BIT per tile in shared memory (1-4 KB), and a rectangle
    // row becomes two shared-memory words and a funnel shift.
    const uint8_t *done = a.tile_done;
    const uint32_t *bits = nullptr;
    if (a.tile_done && a.done_in_smem) {
        uint32_t *s_bits = reinterpret_cast<uint32_t *>(s_done);
        const uint32_t T = a.uniforms->tiles_x * a.uniforms->tiles_y;
        const uint32_t nwords = (T + 31u) / 32u;
        for (uint32_t w = tid; w <= nwords; w += BIN_THREADS) {         // one word past the end stays 0: the funnel shift reads w0 + 1
            uint32_t v = 0u;
            if (w < nwords) {
                const uint4 *src = reinterpret_cast<const uint4 *>(a.tile_done + (size_t)w * 32u);   // 32 flag bytes (allocation padded)
                const uint4 q0 = src[0], q1 = src[1];
                const uint32_t b[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                for (int k = 0; k < 8; k++) {                          // byte i of the 32 -> bit i (flags are 0 / 1)
                    v |= ((b[k] & 0x1u) | ((b[k] >> 7) & 0x2u) | ((b[k] >> 14) & 0x4u) | ((b[k] >> 21) & 0x8u)) << (4 * k);
                }
                const uint32_t valid = T - w * 32u;                     // tiles beyond T (padding bytes) never count as done... nor matter
                if (valid < 32u) v &= (1u << valid) - 1u;
            }
            s_bits[w] = v;
        }
        __syncthreads();
        bits = s_bits;
    }
    for (uint32_t part = blockIdx.x; part < nparts; part += gridDim.x) {
        SplatRects r;
        load_rects(a, lo + part * BIN_PART + tid * BIN_SPT, hi, r);
        if (a.tile_done) a.keep4[part * BIN_THREADS + tid] = drop_saturated(a, r, done, bits);
        uint32_t c = r.cnt[0] + r.cnt[1] + r.cnt[2] + r.cnt[3];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) s_red[warp] = c;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int k = 0; k < BIN_WARPS; k++) t += s_red[k];
            a.part_counts[part] = t;
        }
        __syncthreads();
    }
}

// ---- (2) SCAN: exclusive scan of up to 2^20 partition totals by one CTA --------------------------
__global__ void __launch_bounds__(1024)
bin_scan_kernel(BinningArgs a)
{
    const uint32_t V = a.counters->num_visible;
    uint32_t lo, hi;
    slab_bounds(a, V, lo, hi);
    const uint32_t nparts = (hi - lo + BIN_PART - 1u) / BIN_PART;
    const uint64_t total_out = block_exclusive_scan_1024<uint64_t>(a.part_counts, a.part_bases, nparts);
    if (threadIdx.x == 0) {
        const uint64_t P = (nparts > 0u) ? total_out : 0ull;           // 64-bit: screen-filling splats can push P past 2^32
        *a.num_pairs_out = P > 0xffffffffull ? 0xffffffffu : (uint32_t)P;
        if (P > (uint64_t)a.pair_cap) a.counters->pair_overflow = 1u;  // zeroed per frame; either slab may set it
    }
}

// ---- (3'') EXPAND, round-2 rewrite (WS_BIN_EXPAND=2; NOT the default: measured slower, see launch_binning) -------------
// ncu on the round-1 kernel (profiles/r02a: 139 lane-instructions per emitted pair, 1168 SASS
// instructions, the marker / prefix-max machinery executed for every output position) showed the general
// load-balancing scheme costing 3x what the data needs: a splat touches 3.5 tiles on average, so the owner thread
// simply walks its own rectangle.  Only rectangles above EXP_SMALL tiles (close-ups, gaussian_scaling > 1) are expanded
// cooperatively by the whole block, one 256-wide stride per rectangle.  Pairs are staged in shared memory at their
// block-local output position (scan of the per-splat counts), so the stores to global memory stay perfectly
// coalesced and the output order -- splat (depth order), then row-major tile -- is exactly that of round 1.
constexpr int EXP_CAP = 4096;                         // staged pairs per chunk (2 x 16 KB)
constexpr uint32_t EXP_SMALL = 32;                    // rectangles up to this many tiles are walked by their owner thread
constexpr int EXP_BIG = 512;                          // cooperative list entries per partition (beyond: owner thread walks)

__global__ void __launch_bounds__(BIN_THREADS, 4)
bin_expand_kernel(BinningArgs a)
{
    __shared__ uint32_t s_tile[EXP_CAP];
    __shared__ uint32_t s_slot[EXP_CAP];
    __shared__ uint4 s_big[EXP_BIG];                  // {first pair (block-local), x0 | y0<<16, slot, w | h<<16}
    __shared__ uint32_t s_hist[BIN_NDIG][256];
    __shared__ uint32_t s_scan[BIN_WARPS];
    __shared__ uint32_t s_nbig;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t V = a.counters->num_visible;
    uint32_t lo, hi;
    slab_bounds(a, V, lo, hi);
    const uint32_t nparts = (hi - lo + BIN_PART - 1u) / BIN_PART;
    const uint32_t tiles_x = a.uniforms->tiles_x;
    const uint32_t cap = a.pair_cap;
    const uint32_t ntiles = tiles_x * a.uniforms->tiles_y;
    const int ndig = (ntiles > 65536u) ? 3 : ((ntiles > 256u) ? 2 : 1);

    for (unsigned i = tid; i < (unsigned)(BIN_NDIG * 256); i += BIN_THREADS) (&s_hist[0][0])[i] = 0u;
    if (tid == 0) s_nbig = 0u;
    __syncthreads();

    for (uint32_t part = blockIdx.x; part < nparts; part += gridDim.x) {
        SplatRects r;
        load_rects(a, lo + part * BIN_PART + tid * BIN_SPT, hi, r, a.tile_done ? a.keep4[part * BIN_THREADS + tid] : 0x01010101u);
        const uint32_t base = __ldg(a.part_bases + part);
        if (base >= cap) continue;                        // beyond the pair capacity (overflow is already flagged): nothing of this partition is stored
        const uint32_t mine = r.cnt[0] + r.cnt[1] + r.cnt[2] + r.cnt[3];
        // block scan of the per-thread totals -> block-local position of every splat's first pair
        uint32_t incl = mine;
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
        uint32_t excl[BIN_SPT];
        excl[0] = incl + woff - mine;
#pragma unroll
        for (int j = 1; j < BIN_SPT; j++) excl[j] = excl[j - 1] + r.cnt[j - 1];
        // large rectangles go to the cooperative list (any order: every pair's position is fixed by excl)
        bool own[BIN_SPT];
#pragma unroll
        for (int j = 0; j < BIN_SPT; j++) {
            own[j] = r.cnt[j] > 0u;
            if (r.cnt[j] > EXP_SMALL) {
                const uint32_t e = atomicAdd(&s_nbig, 1u);
                if (e < (uint32_t)EXP_BIG) { s_big[e] = make_uint4(excl[j], r.xy[j], r.slot[j], r.w[j] | (r.h[j] << 16)); own[j] = false; }
            }
        }
        __syncthreads();
        const uint32_t nbig = s_nbig < (uint32_t)EXP_BIG ? s_nbig : (uint32_t)EXP_BIG;

        for (uint32_t c0 = 0; c0 < total; c0 += EXP_CAP) {
            const uint32_t m = (total - c0 < (uint32_t)EXP_CAP) ? total - c0 : (uint32_t)EXP_CAP;
            // 1. owner threads walk their (small) rectangles, row-major
#pragma unroll
            for (int j = 0; j < BIN_SPT; j++) {
                if (own[j] && excl[j] < c0 + m && excl[j] + r.cnt[j] > c0) {
                    const uint32_t x0 = r.xy[j] & 0xffffu, y0 = r.xy[j] >> 16;
                    uint32_t o = excl[j] - c0;                               // may wrap below 0 for a rectangle continuing from the previous chunk: tested per pair
                    uint32_t row = y0 * tiles_x + x0;
                    for (uint32_t yy = 0; yy < r.h[j]; yy++, row += tiles_x)
                        for (uint32_t xx = 0; xx < r.w[j]; xx++, o++)
                            if (o < m) { s_tile[o] = row + xx; s_slot[o] = r.slot[j]; }
                }
            }
            // 2. the block walks the large rectangles together
            for (uint32_t b = 0; b < nbig; b++) {
                const uint4 inf = s_big[b];
                const uint32_t w = inf.w & 0xffffu, cnt = w * (inf.w >> 16);
                if (inf.x >= c0 + m || inf.x + cnt <= c0) continue;          // block-uniform
                const uint32_t q0 = (inf.x < c0) ? c0 - inf.x : 0u;          // first pair of the rectangle inside this chunk
                const uint32_t q1 = (inf.x + cnt > c0 + m) ? c0 + m - inf.x : cnt;
                const uint32_t origin = (inf.y >> 16) * tiles_x + (inf.y & 0xffffu);
                for (uint32_t q = q0 + tid; q < q1; q += BIN_THREADS) {
                    const uint32_t ty = q / w, tx = q - ty * w;
                    const uint32_t o = inf.x + q - c0;
                    s_tile[o] = origin + ty * tiles_x + tx; s_slot[o] = inf.z;
                }
            }
            __syncthreads();
            // 3. coalesced copy-out + tile-id digit histograms
            for (uint32_t i = tid; i < m; i += BIN_THREADS) {
                const uint64_t g = (uint64_t)base + c0 + i;
                if (g < cap) {
                    const uint32_t tile = s_tile[i];
                    a.pair_tiles[g] = tile;
                    a.pair_slots[g] = s_slot[i];
                    for (int d = 0; d < ndig; d++) atomicAdd(&s_hist[d][(tile >> (8 * d)) & 255u], 1u);
                }
            }
            __syncthreads();
        }
        if (tid == 0) s_nbig = 0u;
        __syncthreads();
    }

    for (unsigned i = tid; i < (unsigned)(ndig * 256); i += BIN_THREADS) {
        const uint32_t c = s_hist[i >> 8][i & 255u];
        if (c) atomicAdd(a.hist + i, c);
    }
}

constexpr int BIN_ROUNDS = 8;                         // (v1) output positions per thread and chunk
constexpr int BIN_CAP = BIN_THREADS * BIN_ROUNDS;     // (v1) 2048 pairs per chunk
// ---- (3) EXPAND (default): load-balanced over PAIRS with markers + prefix-max -----------------------------------------
__global__ void __launch_bounds__(BIN_THREADS, 4)
bin_expand_v1_kernel(BinningArgs a)
{
    __shared__ uint32_t s_owner[BIN_CAP];             // marker = owner index + 1 at the owner's first position
    __shared__ uint4 s_info[BIN_PART];                // per splat: {first pair (block-local), x0 | y0<<16, slot, width}
    __shared__ uint32_t s_magic[BIN_PART];            // ceil(2^32 / width)
    __shared__ uint32_t s_hist[BIN_NDIG][256];
    __shared__ uint32_t s_scan[BIN_WARPS];
    __shared__ uint32_t s_wcarry[BIN_WARPS];

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t V = a.counters->num_visible;
    uint32_t lo, hi;
    slab_bounds(a, V, lo, hi);
    const uint32_t nparts = (hi - lo + BIN_PART - 1u) / BIN_PART;
    const uint32_t tiles_x = a.uniforms->tiles_x;
    const uint32_t cap = a.pair_cap;
    const uint32_t ntiles = tiles_x * a.uniforms->tiles_y;
    const int ndig = (ntiles > 65536u) ? 3 : ((ntiles > 256u) ? 2 : 1);

    for (unsigned i = tid; i < (unsigned)(BIN_NDIG * 256); i += BIN_THREADS) (&s_hist[0][0])[i] = 0u;
    __syncthreads();

    for (uint32_t part = blockIdx.x; part < nparts; part += gridDim.x) {
        SplatRects r;
        load_rects(a, lo + part * BIN_PART + tid * BIN_SPT, hi, r, a.tile_done ? a.keep4[part * BIN_THREADS + tid] : 0x01010101u);
        const uint32_t base = __ldg(a.part_bases + part);
        if (base >= cap) continue;                        // beyond the pair capacity (overflow is already flagged): nothing of this partition is stored
        const uint32_t mine = r.cnt[0] + r.cnt[1] + r.cnt[2] + r.cnt[3];
        // block scan of the per-thread totals
        uint32_t incl = mine;
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
        uint32_t excl[BIN_SPT];
        excl[0] = incl + woff - mine;
#pragma unroll
        for (int j = 1; j < BIN_SPT; j++) excl[j] = excl[j - 1] + r.cnt[j - 1];
#pragma unroll
        for (int j = 0; j < BIN_SPT; j++) {
            s_info[tid * BIN_SPT + j] = make_uint4(excl[j], r.xy[j], r.slot[j], r.w[j]);
            s_magic[tid * BIN_SPT + j] = (r.w[j] > 1u) ? __float2uint_ru(__fdiv_ru(4294967296.f, (float)r.w[j])) : 0u;
            // m >= 2^32/w with m*w - 2^32 <= w + 512: floor(t/w) == umulhi(t, m) for every t < 2^20, w <= 1024 (viewport <= 16384)
        }

        for (uint32_t c0 = 0; c0 < total; c0 += BIN_CAP) {
            const uint32_t m = (total - c0 < (uint32_t)BIN_CAP) ? total - c0 : (uint32_t)BIN_CAP;
            const uint32_t rounds = (m + BIN_THREADS - 1u) / BIN_THREADS;       // 32-position rounds per warp
            const uint32_t span = rounds * 32u;                                 // consecutive positions owned by a warp
            // 1. clear, 2. markers
            for (uint32_t q = tid; q < m; q += BIN_THREADS) s_owner[q] = 0u;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < BIN_SPT; j++) {
                if (r.cnt[j] > 0u) {
                    if (excl[j] >= c0 && excl[j] < c0 + m) s_owner[excl[j] - c0] = tid * BIN_SPT + j + 1u;
                    else if (excl[j] < c0 && excl[j] + r.cnt[j] > c0) s_owner[0] = tid * BIN_SPT + j + 1u;   // continues from the previous chunk
                }
            }
            __syncthreads();
            // 3. prefix-max of the markers over this warp's span
            uint32_t own[BIN_ROUNDS];
            uint32_t carry = 0;
#pragma unroll
            for (int k = 0; k < BIN_ROUNDS; k++) {
                own[k] = 0u;
                if ((uint32_t)k < rounds) {
                    const uint32_t q = warp * span + (uint32_t)k * 32u + lane;
                    uint32_t v = (q < m) ? s_owner[q] : 0u;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
                        if ((int)lane >= o) v = v > t ? v : t;
                    }
                    v = v > carry ? v : carry;
                    carry = __shfl_sync(0xffffffffu, v, 31);
                    own[k] = v;
                }
            }
            if (lane == 0) s_wcarry[warp] = carry;
            __syncthreads();
            uint32_t prev = 0;
#pragma unroll
            for (int k = 0; k < BIN_WARPS; k++) if (k < (int)warp) { const uint32_t c = s_wcarry[k]; prev = prev > c ? prev : c; }
            // 4. emit: position -> (tile, slot), coalesced, plus the tile-id digit histogram
#pragma unroll
            for (int k = 0; k < BIN_ROUNDS; k++) {
                if ((uint32_t)k < rounds) {
                    const uint32_t q = warp * span + (uint32_t)k * 32u + lane;
                    const uint64_t g = (uint64_t)base + c0 + q;
                    if (q < m && g < cap) {
                        const uint32_t o = (own[k] > prev ? own[k] : prev) - 1u;
                        const uint4 inf = s_info[o];
                        const uint32_t t = c0 + q - inf.x;
                        const uint32_t ty = (inf.w > 1u) ? __umulhi(t, s_magic[o]) : t;
                        const uint32_t tx = t - ty * inf.w;
                        const uint32_t tile = ((inf.y >> 16) + ty) * tiles_x + (inf.y & 0xffffu) + tx;
                        a.pair_tiles[g] = tile;
                        a.pair_slots[g] = inf.z;
                        for (int d = 0; d < ndig; d++) atomicAdd(&s_hist[d][(tile >> (8 * d)) & 255u], 1u);
                    }
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }

    for (unsigned i = tid; i < (unsigned)(ndig * 256); i += BIN_THREADS) {
        const uint32_t c = s_hist[i >> 8][i & 255u];
        if (c) atomicAdd(a.hist + i, c);
    }
}

}  // namespace

cudaError_t launch_binning(const BinningArgs &a, int grid_count, int grid_expand, cudaStream_t stream)
{
    // far slab: the tile_done map travels in dynamic shared memory when it fits under the 48 KB default limit
    BinningArgs ac = a;
    // (one bit per tile + one spare word; up to 16384 x 16384 pixels = 1 Mi tiles = 128 KB would not fit the default 48 KB: above
    // 256 Ki tiles the kernel reads the byte map from global memory)
    const size_t done_bytes = (a.tile_done && a.num_tiles_hint && a.num_tiles_hint <= 262144u) ? (((size_t)a.num_tiles_hint + 31u) / 32u + 1u) * 4u : 0u;
    ac.done_in_smem = done_bytes ? 1u : 0u;
    bin_count_kernel<<<grid_count, BIN_THREADS, done_bytes, stream>>>(ac);
    bin_scan_kernel<<<1, 1024, 0, stream>>>(a);
    // default: the round-1 kernel (markers + prefix-max).  Measured on B200 (profiles/r02d_*): 80.0 vs 90.7 us for the near
    // slab of cfg3, binning 0.309 vs 0.399 ms at cfg4 -- the round-2 rewrite executes a quarter of the instructions but its
    // owner-thread loops serialise on the longest rectangle of each warp and its scattered staging stores conflict.
    static const int variant = [] { const char *e = getenv("WS_BIN_EXPAND"); return (e && atoi(e) == 2) ? 2 : 1; }();
    if (variant == 1) bin_expand_v1_kernel<<<grid_expand, BIN_THREADS, 0, stream>>>(a);
    else bin_expand_kernel<<<grid_expand, BIN_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

int binning_blocks_per_sm()
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bin_expand_v1_kernel, BIN_THREADS, 0);
    return nb > 0 ? nb : 1;
}

}  // namespace ws
