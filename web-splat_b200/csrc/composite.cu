// composite.cu -- stage 3: tile-binned front-to-back alpha compositing on sm_100a.
//
// Replaces the reference's instanced-quad draw + fixed-function blending: vs_main/fs_main
// (gaussian.wgsl:30-66) and PREMULTIPLIED_ALPHA_BLENDING (renderer.rs:63-67).  Per pixel the
// reference computes, over all splats in ascending key order (far -> near),
//     a = p.p, p = (2[v1 v2])^-1 (ndc - c);  discard if a > 2*CUTOFF;
//     b = min(0.99, exp(-a) * alpha);        dst = (rgb*b, b) + dst * (1 - b).
// Here one CTA owns one 16x16 tile and walks the tile's slice of the sorted pair list from
// its END (nearest splat) to its begin, accumulating C += rgb*b*T, T *= (1-b): the same sum,
// associated front-to-back, which allows the early-out once T < 2^-16.
//
//  * each warp owns an 8x4 pixel block; splats are staged 256 at a time into shared memory
//    (decoded from the 20-B f16 record once per tile, not once per pixel);
//  * per 32 staged splats every lane tests one splat's bounding box against the warp's
//    pixel block and a ballot compacts the survivors: a warp only evaluates splats that
//    can touch its 32 pixels (most of a tile's list does not), and skips everything once
//    all of its pixels are saturated (warp-level early-out);
//  * the centre is expressed relative to the tile origin with exact f16 x integer products,
//    so `a` carries ~1e-5 absolute error at 4K instead of ulp(3840).
#include "ws_device.cuh"
#include "ws_kernels.h"

namespace ws {

namespace {

constexpr int CB_THREADS = 256;
constexpr int CB_BATCH = 256;
constexpr float T_EPS = 1.52587890625e-5f;   // 2^-16: early-out once the remaining transmittance cannot move the result by more
constexpr float LOG2E = 1.4426950408889634f;
constexpr float SQRT_LOG2E = 1.2011224087864498f;

__device__ __forceinline__ float hlo(uint32_t w) { __half_raw r; r.x = (unsigned short)(w & 0xffffu); return __half2float(__half(r)); }
__device__ __forceinline__ float hhi(uint32_t w) { __half_raw r; r.x = (unsigned short)(w >> 16); return __half2float(__half(r)); }
// MUFU.EX2: the argument is in [-6.8, 0], far from the denormal range, so .ftz is exact enough
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int FORMAT>
__global__ void __launch_bounds__(CB_THREADS)
composite_kernel(CompositeArgs a)
{
    // staged splats, SoA as three float4 planes + bbox plane
    __shared__ float4 s_a[CB_BATCH];   // cx, cy (relative to tile origin, pixel units), i00, i01
    __shared__ float4 s_b[CB_BATCH];   // i10, i11, alpha, r
    __shared__ float4 s_c[CB_BATCH];   // g, b, ex, ey

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t W = a.uniforms->width, H = a.uniforms->height;
    const uint32_t tile_x = blockIdx.x, tile_y = blockIdx.y + a.tile_y0;
    const uint32_t tile = tile_y * gridDim.x + tile_x;
    uint2 range = a.ranges[tile];
    range.y = ~range.y;                                  // stored complemented (atomicMin in the sort's last pass)
    if (range.y <= range.x) range.x = range.y = 0u;      // untouched tile
    const float fw = (float)W, fh = (float)H;
    const float hw = 0.5f * fw, hh = 0.5f * fh;
    const float ox = hw - (float)(tile_x * TILE), oy = hh - (float)(tile_y * TILE);   // exact

    // warp's 8x4 pixel block inside the tile
    const uint32_t bx = (warp & 1u) * 8u, by = (warp >> 1) * 4u;
    const uint32_t lx = bx + (lane & 7u), ly = by + (lane >> 3);
    const uint32_t px = tile_x * TILE + lx, py = tile_y * TILE + ly;
    const bool inside = (px < W) && (py < H);
    const float fx = (float)lx + 0.5f, fy = (float)ly + 0.5f;
    // block bounds in pixel-centre coordinates relative to the tile origin
    const float blo_x = (float)bx + 0.5f, bhi_x = (float)bx + 7.5f;
    const float blo_y = (float)by + 0.5f, bhi_y = (float)by + 3.5f;

    float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f;
    bool done = !inside;

    int32_t remaining = (int32_t)(range.y - range.x);
    uint32_t cursor = range.y;           // walk from the end: nearest first
    while (remaining > 0) {
        const bool warp_done = __all_sync(0xffffffffu, done);
        if (__syncthreads_and(warp_done ? 1 : 0)) break;

        const int cnt = remaining < CB_BATCH ? remaining : CB_BATCH;
        if ((int)tid < cnt) {
            const uint32_t slot = a.pair_slots[cursor - 1u - tid];
            const uint32_t *sp = a.splats + (size_t)slot * 5u;
            const uint32_t w0 = __ldg(sp), w1 = __ldg(sp + 1), w2 = __ldg(sp + 2), w3 = __ldg(sp + 3), w4 = __ldg(sp + 4);
            const float v1x = hlo(w0), v1y = hhi(w0), v2x = hlo(w1), v2y = hhi(w1);
            const float ccx = hlo(w2), ccy = hhi(w2);
            // pixel-space map  [dx;dy] = B p,  B = [[W v1x, W v2x],[-H v1y, -H v2y]]  (y down)
            const float b00 = fw * v1x, b01 = fw * v2x, b10 = -(fh * v1y), b11 = -(fh * v2y);   // exact products
            const float det = b00 * b11 - b01 * b10;
            const float inv = SQRT_LOG2E / det;          // fold log2(e) into p so that exp(-a) = exp2(-a')
            float4 A, B, C;
            A.x = ccx * hw + ox;                         // centre relative to the tile origin (exact product + 1 rounding)
            A.y = oy - ccy * hh;
            A.z = b11 * inv;  A.w = -b01 * inv;
            B.x = -b10 * inv; B.y = b00 * inv;
            B.z = hhi(w4);                               // alpha
            B.w = hlo(w3);                               // r
            C.x = hhi(w3); C.y = hlo(w4);                // g, b
            C.z = FOOTPRINT_R * (fw * sqrtf(v1x * v1x + v2x * v2x)) + RECT_PAD;
            C.w = FOOTPRINT_R * (fh * sqrtf(v1y * v1y + v2y * v2y)) + RECT_PAD;
            s_a[tid] = A; s_b[tid] = B; s_c[tid] = C;
        }
        __syncthreads();

        if (!warp_done) {
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                const int k = c0 + (int)lane;
                bool hit = false;
                if (k < cnt) {
                    const float4 A = s_a[k];
                    const float4 C = s_c[k];
                    hit = (A.x + C.z >= blo_x) && (A.x - C.z <= bhi_x) && (A.y + C.w >= blo_y) && (A.y - C.w <= bhi_y);
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int j = c0 + (__ffs(m) - 1);
                    m &= m - 1u;
                    const float4 A = s_a[j];
                    const float4 B = s_b[j];
                    const float dx = fx - A.x, dy = fy - A.y;
                    const float p0 = A.z * dx + A.w * dy;
                    const float p1 = B.x * dx + B.y * dy;
                    const float aa = p0 * p0 + p1 * p1;                 // = a * log2(e)
                    if (!done && aa <= TWO_CUTOFF * LOG2E) {
                        const float4 C = s_c[j];
                        float wgt = ex2_approx(-aa) * B.z;
                        wgt = fminf(0.99f, wgt);
                        const float wt = wgt * T;
                        cr += B.w * wt; cg += C.x * wt; cb += C.y * wt;
                        T *= (1.f - wgt);
                        if (T < T_EPS) done = true;
                    }
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
        remaining -= cnt;
        cursor -= (uint32_t)cnt;
        // next iteration's __syncthreads_and orders these smem reads before the restaging
    }

    if (inside) {
        const float r = cr + a.clear[0] * T, g = cg + a.clear[1] * T, b = cb + a.clear[2] * T;
        const float al = (1.f - T) + a.clear[3] * T;
        uint8_t *row = reinterpret_cast<uint8_t *>(a.dst) + (size_t)(py - a.tile_y0 * TILE) * a.row_pitch;
        if (FORMAT == 2) {
            reinterpret_cast<float4 *>(row)[px] = make_float4(r, g, b, al);
        } else if (FORMAT == 1) {
            __half_raw h0 = __float2half_rn(r), h1 = __float2half_rn(g), h2 = __float2half_rn(b), h3 = __float2half_rn(al);
            reinterpret_cast<uint2 *>(row)[px] = make_uint2((uint32_t)h0.x | ((uint32_t)h1.x << 16), (uint32_t)h2.x | ((uint32_t)h3.x << 16));
        } else {
            const uint32_t r8 = __float2uint_rn(__saturatef(r) * 255.f), g8 = __float2uint_rn(__saturatef(g) * 255.f);
            const uint32_t b8 = __float2uint_rn(__saturatef(b) * 255.f), a8 = __float2uint_rn(__saturatef(al) * 255.f);
            reinterpret_cast<uint32_t *>(row)[px] = r8 | (g8 << 8) | (b8 << 16) | (a8 << 24);
        }
    }
}

}  // namespace

cudaError_t launch_composite(const CompositeArgs &a, uint32_t tiles_x, uint32_t tiles_y, cudaStream_t stream)
{
    dim3 grid(tiles_x, tiles_y);
    switch (a.format) {
    case 0: composite_kernel<0><<<grid, CB_THREADS, 0, stream>>>(a); break;
    case 1: composite_kernel<1><<<grid, CB_THREADS, 0, stream>>>(a); break;
    default: composite_kernel<2><<<grid, CB_THREADS, 0, stream>>>(a); break;
    }
    return cudaGetLastError();
}

}  // namespace ws
