// shard.cu -- multi-GPU exchange step of the splat render path.
//
// The reference is single-GPU.  The sharded path splits one frame over G GPUs of a box:
//   stage 1              by Gaussian index: rank r preprocesses Gaussians [r*N/G, (r+1)*N/G);
//   stages 2 and 3       by tile-row band: rank d depth-sorts, bins and composites only the
//                        splats that touch its band of 16-pixel tile rows.
// Alpha compositing is order dependent over ALL splats of a pixel, so partial images cannot be
// merged (no sort-last); what is exchanged between the two phases are the visible splats
// themselves: 32 B each (20-B Splat, 4-B depth key, 8-B tile rectangle clipped to the band).
//
// The exchange is one kernel that STORES DIRECTLY INTO THE DESTINATION GPU'S BUFFERS through
// peer-mapped pointers (cudaIpc handles, NVLink 5 / NVSwitch): routing (which band), compaction
// (ballot ranks + scanned per-partition bases) and the transfer are fused; no send staging, no
// NCCL all-to-all with host-known sizes.  Destination offsets come from a G x G matrix of counts
// that the ranks all-gather (G*4 B each, NCCL) between the count pass and the scatter pass.
// Records arrive ordered by (source rank, source slot) = global Gaussian index order, so the
// stable depth sort that follows breaks ties exactly like the single-GPU path and the G-GPU
// frame is bit-identical to the 1-GPU frame.
#include "ws_device.cuh"
#include "ws_kernels.h"

namespace ws {

namespace {

constexpr int RT_THREADS = 256;
constexpr int RT_WARPS = RT_THREADS / 32;
constexpr int RT_SPT = 4;                          // slots per thread
constexpr int RT_PART = RT_THREADS * RT_SPT;       // 1024 local slots per partition

__device__ __forceinline__ bool touches_band(uint2 rect, uint32_t b0, uint32_t b1)
{
    const uint32_t y0 = rect.x >> 16, h = rect.y >> 16, w = rect.y & 0xffffu;
    return (w != 0u) && (h != 0u) && (y0 < b1) && (y0 + h > b0);
}

__device__ __forceinline__ void load_rects4(const RouteArgs &a, uint32_t first, uint32_t V, uint2 rc[RT_SPT])
{
    if (first + RT_SPT <= V) {
        const uint4 *p = reinterpret_cast<const uint4 *>(a.l_rects + first);       // first is a multiple of 4: 32-B aligned
        const uint4 u = p[0], v = p[1];
        rc[0] = make_uint2(u.x, u.y); rc[1] = make_uint2(u.z, u.w); rc[2] = make_uint2(v.x, v.y); rc[3] = make_uint2(v.z, v.w);
    } else {
#pragma unroll
        for (int j = 0; j < RT_SPT; j++) rc[j] = (first + j < V) ? a.l_rects[first + j] : make_uint2(0u, 0u);
    }
}

// ---- pass 1: per 1024-slot partition, how many local splats go to each band
__global__ void __launch_bounds__(RT_THREADS)
route_count_kernel(RouteArgs a)
{
    __shared__ uint32_t s_cnt[8];
    const unsigned tid = threadIdx.x;
    const uint32_t V = a.counters->num_visible;
    const uint32_t nparts = (V + RT_PART - 1u) / RT_PART;
    for (uint32_t part = blockIdx.x; part < nparts; part += gridDim.x) {
        if (tid < 8) s_cnt[tid] = 0u;
        __syncthreads();
        uint2 rc[RT_SPT];
        load_rects4(a, part * RT_PART + tid * RT_SPT, V, rc);
        for (uint32_t d = 0; d < a.world; d++) {
            uint32_t c = 0;
#pragma unroll
            for (int j = 0; j < RT_SPT; j++) c += touches_band(rc[j], a.band_y0[d], a.band_y0[d + 1]) ? 1u : 0u;
            if (c) atomicAdd(&s_cnt[d], c);           // counts only: order does not matter here
        }
        __syncthreads();
        if (tid < a.world) a.part_band_counts[(size_t)part * a.world + tid] = s_cnt[tid];
        __syncthreads();
    }
}

// ---- pass 2 (one CTA): per band, exclusive scan over the partitions; totals = this rank's matrix row
__global__ void __launch_bounds__(1024)
route_scan_kernel(RouteArgs a)
{
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_tot[8];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t V = a.counters->num_visible;
    const uint32_t nparts = (V + RT_PART - 1u) / RT_PART;
    const uint32_t per = (nparts + 1023u) / 1024u;
    const uint32_t lo = tid * per, hi = (lo + per < nparts) ? lo + per : nparts;
    for (uint32_t d = 0; d < a.world; d++) {
        uint32_t sum = 0;
        for (uint32_t i = lo; i < hi; i++) sum += a.part_band_counts[(size_t)i * a.world + d];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t v = s_w[lane];
            uint32_t vi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
                if ((int)lane >= o) vi += t;
            }
            s_w[lane] = vi - v;
            if (lane == 31) s_total = vi;
        }
        __syncthreads();
        uint32_t run = s_w[warp] + incl - sum;
        for (uint32_t i = lo; i < hi; i++) {
            const uint32_t c = a.part_band_counts[(size_t)i * a.world + d];
            a.part_band_bases[(size_t)i * a.world + d] = run;
            run += c;
        }
        if (tid == 0) { const uint32_t t = (nparts > 0u) ? s_total : 0u; if (a.totals) a.totals[d] = t; s_tot[d] = t; }
        __syncthreads();
    }
    if (a.peer_mail[a.rank]) {
        // publish this rank's row into EVERY rank's mailbox, then raise the row flag (release, system scope)
        const uint32_t epoch = *a.epoch_ptr;          // frame number in device memory: the frame replays as a CUDA graph
        const uint32_t par = epoch & 1u;
        if (tid < a.world * a.world) {
            const uint32_t p = tid / a.world, d = tid - p * a.world;
            a.peer_mail[p]->matrix[par][a.rank * a.world + d] = s_tot[d];
        }
        __threadfence_system();
        __syncthreads();
        if (tid < a.world) st_release_sys(&a.peer_mail[tid]->flag_rows[a.rank], epoch);
    }
}

// ---- pass 3: routing + compaction + transfer, fused: store records into the owners' buffers.
// Each WARP owns 128 consecutive slots of the partition and moves its own records: one block barrier per
// partition (to turn the per-warp, per-band counts into offsets) instead of three per band.
__global__ void __launch_bounds__(RT_THREADS)
route_scatter_kernel(RouteArgs a)
{
    __shared__ uint32_t s_list[RT_WARPS][32 * RT_SPT];   // per warp: local slots bound for the current band, in slot order
    __shared__ uint32_t s_wb[RT_WARPS][8];               // per warp and band: number of records
    __shared__ uint32_t s_recv_off[8];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t V = a.counters->num_visible;
    const uint32_t nparts = (V + RT_PART - 1u) / RT_PART;
    const uint32_t *matrix = a.matrix;
    const uint32_t epoch = a.peer_mail[a.rank] ? *a.epoch_ptr : 0u;
    if (a.peer_mail[a.rank]) {                    // mailbox mode: wait until every rank's count row of this frame has arrived
        if (!a.gated && tid < a.world) wait_epoch(&a.peer_mail[a.rank]->flag_rows[tid], epoch, a.err);
        __syncthreads();
        matrix = a.peer_mail[a.rank]->matrix[epoch & 1u];
    }
    if (tid < a.world) {                          // where this rank's records start in each destination
        uint32_t off = 0;
        for (uint32_t s = 0; s < a.rank; s++) off += matrix[s * a.world + tid];
        s_recv_off[tid] = off;
    }
    __syncthreads();
    for (uint32_t part = blockIdx.x; part < nparts; part += gridDim.x) {
        const uint32_t first = part * RT_PART + tid * RT_SPT;
        uint2 rc[RT_SPT];
        load_rects4(a, first, V, rc);
        // per-warp counts for every band
        for (uint32_t d = 0; d < a.world; d++) {
            uint32_t c = 0;
#pragma unroll
            for (int j = 0; j < RT_SPT; j++) c += touches_band(rc[j], a.band_y0[d], a.band_y0[d + 1]) ? 1u : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) s_wb[warp][d] = c;
        }
        __syncthreads();
        for (uint32_t d = 0; d < a.world; d++) {
            const uint32_t b0 = a.band_y0[d], b1 = a.band_y0[d + 1];
            uint32_t wbase = 0;
            for (uint32_t w = 0; w < warp; w++) wbase += s_wb[w][d];
            const uint32_t cnt = s_wb[warp][d];
            if (cnt == 0u) continue;                                           // warp-uniform
            bool go[RT_SPT];
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < RT_SPT; j++) { go[j] = touches_band(rc[j], b0, b1); mine += go[j] ? 1u : 0u; }
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += t;
            }
            uint32_t pos = incl - mine;
            uint32_t *lst = s_list[warp];
#pragma unroll
            for (int j = 0; j < RT_SPT; j++) if (go[j]) lst[pos++] = first + j;
            __syncwarp();
            const uint32_t dst0 = s_recv_off[d] + a.part_band_bases[(size_t)part * a.world + d] + wbase;
            if (dst0 + cnt > a.recv_cap) {
                if (lane == 0) atomicOr(a.err, 2u);                             // receiver capacity exceeded: nothing is written
            } else {
                uint32_t *ds = a.peer_splats[d] + (size_t)dst0 * 5u;
                const uint32_t nw = cnt * 5u;
                uint32_t i = lane;
                for (; i + 96u < nw; i += 128u) {                               // four independent gathers in flight per lane
                    uint32_t v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t ii = i + (uint32_t)u * 32u, e = ii / 5u, k = ii - e * 5u;
                        v[u] = a.l_splats[(size_t)lst[e] * 5u + k];
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) ds[i + (uint32_t)u * 32u] = v[u];
                }
                for (; i < nw; i += 32u) {
                    const uint32_t e = i / 5u, k = i - e * 5u;
                    ds[i] = a.l_splats[(size_t)lst[e] * 5u + k];
                }
                for (uint32_t e = lane; e < cnt; e += 32u) {
                    const uint32_t src = lst[e];
                    a.peer_keys[d][dst0 + e] = a.l_keys[src];
                    const uint2 r = a.l_rects[src];                            // clip the rectangle to the band's tile rows
                    const uint32_t y0 = r.x >> 16, h = r.y >> 16;
                    const uint32_t ny0 = y0 > b0 ? y0 : b0;
                    const uint32_t ny1 = (y0 + h < b1) ? y0 + h : b1;
                    a.peer_rects[d][dst0 + e] = make_uint2((r.x & 0xffffu) | (ny0 << 16), (r.y & 0xffffu) | ((ny1 - ny0) << 16));
                }
            }
            __syncwarp();
        }
        __syncthreads();
    }
    if (a.peer_mail[a.rank]) {
        // the last CTA to finish tells every rank that this rank's records have landed
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            const uint32_t prev = atomicAdd(a.done_counter, 1u);
            if (prev == gridDim.x - 1u) {
                __threadfence_system();
                for (uint32_t p = 0; p < a.world; p++) st_release_sys(&a.peer_mail[p]->flag_xchg[a.rank], epoch);
            }
        }
    }
}

// ---- after the exchange: V' = records received; payload iota for the depth sort
__global__ void __launch_bounds__(256)
shard_finish_kernel(const uint32_t *matrix, uint32_t world, uint32_t rank, uint32_t recv_cap,
                    FrameCounters *counters, uint32_t *vals)
{
    uint32_t v = 0;
    for (uint32_t s = 0; s < world; s++) v += matrix[s * world + rank];
    if (v > recv_cap) v = recv_cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) { counters->num_local_visible = counters->num_visible; }
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < v; i += gridDim.x * 256u) vals[i] = i;
}
__global__ void shard_set_visible_kernel(const uint32_t *matrix, uint32_t world, uint32_t rank, uint32_t recv_cap, FrameCounters *counters)
{
    uint32_t v = 0;
    for (uint32_t s = 0; s < world; s++) v += matrix[s * world + rank];
    if (v > recv_cap) { v = recv_cap; atomicOr(&counters->error_flags, 2u); }
    counters->num_visible = v;
}

// peer mode: wait for every rank's exchange flag, then V', payload iota and the depth-key digit histograms in one kernel
__global__ void __launch_bounds__(256)
shard_finish_peer_kernel(RouteArgs a, uint32_t *vals, const uint32_t *__restrict__ keys, uint32_t *hist, int passes, FrameCounters *counters)
{
    __shared__ uint32_t s_hist[4 * 256];
    const unsigned tid = threadIdx.x;
    for (unsigned i = tid; i < 4u * 256u; i += 256u) s_hist[i] = 0u;
    const ShardMailbox *mail = a.peer_mail[a.rank];
    const uint32_t epoch = *a.epoch_ptr;
    if (!a.gated && tid < a.world) wait_epoch(&mail->flag_xchg[tid], epoch, a.err);
    __syncthreads();
    const uint32_t *matrix = mail->matrix[epoch & 1u];
    uint32_t v = 0;
    for (uint32_t s = 0; s < a.world; s++) v += matrix[s * a.world + a.rank];
    if (v > a.recv_cap) { v = a.recv_cap; if (blockIdx.x == 0 && tid == 0) atomicOr(a.err, 2u); }
    for (uint32_t i = blockIdx.x * 256u + tid; i < v; i += gridDim.x * 256u) {
        vals[i] = i;
        const uint32_t k = keys[i];
        for (int d = 0; d < passes; d++) atomicAdd(&s_hist[d * 256 + ((k >> (8 * d)) & 255u)], 1u);
    }
    __syncthreads();
    for (unsigned i = tid; i < (unsigned)passes * 256u; i += 256u) { const uint32_t c = s_hist[i]; if (c) atomicAdd(hist + i, c); }
    if (blockIdx.x == 0 && tid == 0) { counters->num_local_visible = counters->num_visible; counters->num_visible = v; }
}

// root only: the assembled frame is complete once every rank's band flag has reached the epoch
__global__ void wait_bands_kernel(const ShardMailbox *mail, uint32_t world, const uint32_t *epoch_ptr, uint32_t *err)
{
    if (threadIdx.x < world) wait_epoch(&mail->flag_band[threadIdx.x], *epoch_ptr, err);
}

// frame number in device memory, advanced by the frame itself: nothing in a sharded frame's launch parameters changes
// from frame to frame (apart from the frame-buffer parity: two graphs), so the whole frame replays as a CUDA graph
__global__ void epoch_advance_kernel(uint32_t *epoch) { *epoch += 1u; }

// Gate: ONE warp spins on a set of epoch flags; the kernel behind it in the stream starts once all have arrived.
// Used instead of the in-kernel waits when several sharded frames are in flight on one GPU: a grid-wide spin
// could fill every SM of this GPU while the peer it waits for is itself blocked behind this GPU's other frame
// (a cross-GPU resource cycle); a one-warp gate cannot starve anything.
__global__ void gate_kernel(const uint32_t *flags, uint32_t world, const uint32_t *epoch_ptr, uint32_t *err)
{
    if (threadIdx.x < world) wait_epoch(flags + threadIdx.x, *epoch_ptr, err);
}

}  // namespace

cudaError_t launch_shard_finish_peer(const RouteArgs &a, uint32_t *vals, const uint32_t *keys, uint32_t *hist, int passes,
                                     FrameCounters *counters, int grid, cudaStream_t stream)
{
    // NOTE: counters->num_visible is rewritten by block 0 while other blocks of this kernel never read it
    if (a.gated) gate_kernel<<<1, 32, 0, stream>>>(a.peer_mail[a.rank]->flag_xchg, a.world, a.epoch_ptr, a.err);
    shard_finish_peer_kernel<<<grid, 256, 0, stream>>>(a, vals, keys, hist, passes, counters);
    return cudaGetLastError();
}

cudaError_t launch_wait_bands(const ShardMailbox *mail, uint32_t world, const uint32_t *epoch_ptr, uint32_t *err, cudaStream_t stream)
{
    wait_bands_kernel<<<1, 32, 0, stream>>>(mail, world, epoch_ptr, err);
    return cudaGetLastError();
}

cudaError_t launch_epoch_advance(uint32_t *epoch, cudaStream_t stream)
{
    epoch_advance_kernel<<<1, 1, 0, stream>>>(epoch);
    return cudaGetLastError();
}

cudaError_t launch_route_count(const RouteArgs &a, int grid, cudaStream_t stream)
{
    route_count_kernel<<<grid, RT_THREADS, 0, stream>>>(a);
    route_scan_kernel<<<1, 1024, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_route_scatter(const RouteArgs &a, int grid, cudaStream_t stream)
{
    if (a.gated && a.peer_mail[a.rank]) gate_kernel<<<1, 32, 0, stream>>>(a.peer_mail[a.rank]->flag_rows, a.world, a.epoch_ptr, a.err);
    route_scatter_kernel<<<grid, RT_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_shard_finish(const uint32_t *matrix, uint32_t world, uint32_t rank, uint32_t recv_cap,
                                FrameCounters *counters, uint32_t *vals, int grid, cudaStream_t stream)
{
    shard_finish_kernel<<<grid, 256, 0, stream>>>(matrix, world, rank, recv_cap, counters, vals);
    shard_set_visible_kernel<<<1, 1, 0, stream>>>(matrix, world, rank, recv_cap, counters);
    return cudaGetLastError();
}

}  // namespace ws
