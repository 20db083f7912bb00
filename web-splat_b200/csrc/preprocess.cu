// preprocess.cu -- stage 1 of the splat render path on sm_100a.
//
// Replaces preprocess.wgsl:163-280 / preprocess_compressed.wgsl:206-331 of the reference
// (one thread per Gaussian: cull, 3D->2D covariance, eigen axes, SH colour, f16 pack,
// depth key) and adds what the tile-binned design needs: the per-splat 16x16-tile
// rectangle and the digit histograms of the depth keys for the onesweep that follows.
//
// B200 design notes
//  * persistent CTAs take 256-Gaussian partitions from an atomic ticket; the 28-B (24-B)
//    AoS records of a partition are staged with ONE cp.async.bulk (TMA engine, UBLKCP)
//    into shared memory and read from there at a conflict-free 7-word (6-word: 2-way)
//    stride, so HBM sees only full, coalesced 7168-B (6144-B) bursts;
//  * SH (96 B per survivor) is fetched with three 256-bit loads per lane: one full
//    32-B sector per request, no sector is requested twice;
//  * compaction is deterministic: ballot + block scan + single-pass decoupled look-back
//    over the ticket-ordered partitions (the reference uses one contended global atomic,
//    preprocess.wgsl:262, which also makes its slot order nondeterministic);
//  * outputs are staged through shared memory and leave as coalesced streams;
//  * this file is compiled with -fmad=false: every f32 operation rounds once, in the
//    order written, which makes stage 1 bit-identical to the CPU oracle (raw layout).
#include "ws_device.cuh"
#include "ws_kernels.h"

namespace ws {

namespace {

constexpr int PP_THREADS = 256;
constexpr int PP_WARPS = PP_THREADS / 32;

__device__ __forceinline__ float half_lo(uint32_t w) { __half_raw r; r.x = (unsigned short)(w & 0xffffu); return __half2float(__half(r)); }
__device__ __forceinline__ float half_hi(uint32_t w) { __half_raw r; r.x = (unsigned short)(w >> 16); return __half2float(__half(r)); }
__device__ __forceinline__ uint32_t pack2h(float a, float b)
{
    __half_raw ra = __float2half_rn(a), rb = __float2half_rn(b);
    return (uint32_t)ra.x | ((uint32_t)rb.x << 16);
}

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3s(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 v3add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 v3sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }

// SH basis constants, preprocess.wgsl:4-23
#define WS_SH_C0 0.28209479177387814f
#define WS_SH_C1 0.4886025119029199f
#define WS_SH_C2_0 1.0925484305920792f
#define WS_SH_C2_1 (-1.0925484305920792f)
#define WS_SH_C2_2 0.31539156525252005f
#define WS_SH_C2_3 (-1.0925484305920792f)
#define WS_SH_C2_4 0.5462742152960396f
#define WS_SH_C3_0 (-0.5900435899266435f)
#define WS_SH_C3_1 2.890611442640554f
#define WS_SH_C3_2 (-0.4570457994644658f)
#define WS_SH_C3_3 0.3731763325901154f
#define WS_SH_C3_4 (-0.4570457994644658f)
#define WS_SH_C3_5 1.445305721320277f
#define WS_SH_C3_6 (-0.5900435899266435f)

// Coefficient fetchers: k-th RGB triple as f32 (preprocess.wgsl:114-121 / compressed :147-171)
struct ShRaw {
    uint32_t w[24];                 // [[f16;3];16]: half index = k*3 + ch
    __device__ __forceinline__ float h(int i) const { return (i & 1) ? half_hi(w[i >> 1]) : half_lo(w[i >> 1]); }
    __device__ __forceinline__ V3 coef(int k) const { return V3{h(k * 3), h(k * 3 + 1), h(k * 3 + 2)}; }
};
struct ShQuant {
    const int8_t *p;                // entry base: (sh_idx * ncoef) * 3 bytes
    Quant dc, rest;
    __device__ __forceinline__ float dq(int8_t b, const Quant &q) const
    {
        float sn = (float)b / 127.f;            // unpack4x8snorm: max(i/127, -1)
        if (sn < -1.f) sn = -1.f;
        float v = sn * 127.f;
        return (v - (float)q.zero_point) * q.scale;   // dequantizef4
    }
    __device__ __forceinline__ V3 coef(int k) const
    {
        const Quant &q = (k == 0) ? dc : rest;
        return V3{dq(p[k * 3], q), dq(p[k * 3 + 1], q), dq(p[k * 3 + 2], q)};
    }
};

// evaluate_sh, preprocess.wgsl:124-154: same operation order as oracle/ws_oracle.c
template <class SH>
__device__ __forceinline__ V3 evaluate_sh(float x, float y, float z, const SH &sh, uint32_t deg)
{
    V3 result = v3s(WS_SH_C0, sh.coef(0));
    if (deg > 0u) {
        V3 t = v3s((-WS_SH_C1) * y, sh.coef(1));
        t = v3add(t, v3s(WS_SH_C1 * z, sh.coef(2)));
        t = v3sub(t, v3s(WS_SH_C1 * x, sh.coef(3)));
        result = v3add(result, t);
        if (deg > 1u) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            V3 u = v3s(WS_SH_C2_0 * xy, sh.coef(4));
            u = v3add(u, v3s(WS_SH_C2_1 * yz, sh.coef(5)));
            u = v3add(u, v3s(WS_SH_C2_2 * (2.0f * zz - xx - yy), sh.coef(6)));
            u = v3add(u, v3s(WS_SH_C2_3 * xz, sh.coef(7)));
            u = v3add(u, v3s(WS_SH_C2_4 * (xx - yy), sh.coef(8)));
            result = v3add(result, u);
            if (deg > 2u) {
                V3 w = v3s(WS_SH_C3_0 * y * (3.0f * xx - yy), sh.coef(9));
                w = v3add(w, v3s(WS_SH_C3_1 * xy * z, sh.coef(10)));
                w = v3add(w, v3s(WS_SH_C3_2 * y * (4.0f * zz - xx - yy), sh.coef(11)));
                w = v3add(w, v3s(WS_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy), sh.coef(12)));
                w = v3add(w, v3s(WS_SH_C3_4 * x * (4.0f * zz - xx - yy), sh.coef(13)));
                w = v3add(w, v3s(WS_SH_C3_5 * z * (xx - yy), sh.coef(14)));
                w = v3add(w, v3s(WS_SH_C3_6 * x * (xx - 3.0f * yy), sh.coef(15)));
                result = v3add(result, w);
            }
        }
    }
    result.x = result.x + 0.5f; result.y = result.y + 0.5f; result.z = result.z + 0.5f;
    return result;
}

__device__ __forceinline__ void ldg256(const void *p, uint32_t *r)
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}

struct Stage1 {
    uint32_t splat[5];      // v_0, v_1, pos, color_0, color_1 (pointcloud.rs:352-358)
    uint32_t key;
    uint32_t rect_xy;       // x0 | y0 << 16
    uint32_t rect_wh;       // w  | h  << 16  (w == 0: touches no tile)
};

// Everything after the cull: preprocess.wgsl:194-279 / compressed :235-330.
template <bool COMPRESSED, class SH>
__device__ __forceinline__ void project_tail(const FrameUniforms &U, float x, float y, float z,
                                             float cs0, float cs1, float cs2,
                                             float p0, float p1, float p2, float p3,
                                             const float cov6[6], float opacity, const SH &sh, Stage1 &o)
{
    const float *view = U.cam.view;
    const float fx = U.cam.focal[0], fy = U.cam.focal[1];

    // scale_mod (:196-201)
    float scale_mod = 0.f;
    {
        float ddx = U.rs.center[0] - x, ddy = U.rs.center[1] - y, ddz = U.rs.center[2] - z;
        float dist = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
        float dd = 5.f * dist / U.rs.scene_extend;
        if (U.rs.walltime > dd) {
            float t = U.rs.walltime - dd;
            t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
            scale_mod = t * t * (3.f - 2.f * t);
        }
    }
    const float scaling = U.rs.gaussian_scaling * scale_mod;
    const float c0 = cov6[0] * scaling * scaling, c1 = cov6[1] * scaling * scaling, c2 = cov6[2] * scaling * scaling;
    const float c3 = cov6[3] * scaling * scaling, c4 = cov6[4] * scaling * scaling, c5 = cov6[5] * scaling * scaling;
    const float Vm[3][3] = {{c0, c1, c2}, {c1, c3, c4}, {c2, c4, c5}};

    const float j00 = fx / cs2;
    const float j20 = -(fx * cs0) / (cs2 * cs2);
    const float j11 = -fy / cs2;
    const float j21 = (fy * cs1) / (cs2 * cs2);

    float T0[3], T1[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float w0 = view[i * 4 + 0], w1 = view[i * 4 + 1], w2 = view[i * 4 + 2];
        T0[i] = w0 * j00 + w2 * j20;
        T1[i] = w1 * j11 + w2 * j21;
    }
    float A0[3], A1[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        float a = T0[0] * Vm[0][j]; a = a + T0[1] * Vm[1][j]; a = a + T0[2] * Vm[2][j]; A0[j] = a;
        float b = T1[0] * Vm[0][j]; b = b + T1[1] * Vm[1][j]; b = b + T1[2] * Vm[2][j]; A1[j] = b;
    }
    float cov00 = A0[0] * T0[0]; cov00 = cov00 + A0[1] * T0[1]; cov00 = cov00 + A0[2] * T0[2];
    float cov01 = A1[0] * T0[0]; cov01 = cov01 + A1[1] * T0[1]; cov01 = cov01 + A1[2] * T0[2];
    float cov11 = A1[0] * T1[0]; cov11 = cov11 + A1[1] * T1[1]; cov11 = cov11 + A1[2] * T1[2];

    const float ks = U.rs.kernel_size;
    if (U.rs.mip_splatting) {                      // :226-236
        float det_0 = cov00 * cov11 - cov01 * cov01;
        det_0 = (det_0 > 1e-6f) ? det_0 : 1e-6f;
        float det_1 = (cov00 + ks) * (cov11 + ks) - cov01 * cov01;
        det_1 = (det_1 > 1e-6f) ? det_1 : 1e-6f;
        float coef = sqrtf(det_0 / (det_1 + 1e-6f) + 1e-6f);
        if (det_0 <= 1e-6f || det_1 <= 1e-6f) coef = 0.0f;
        opacity = opacity * coef;
    }

    const float diagonal1 = cov00 + ks, offDiagonal = cov01, diagonal2 = cov11 + ks;
    const float mid = 0.5f * (diagonal1 + diagonal2);
    const float hx = (diagonal1 - diagonal2) / 2.0f;
    const float radius = sqrtf(hx * hx + offDiagonal * offDiagonal);
    float lambda1, lambda2;
    if (!COMPRESSED) {
        lambda1 = mid + radius;
        float l2 = mid - radius;
        lambda2 = (l2 > 0.1f) ? l2 : 0.1f;
    } else {
        float rr = (radius > 0.1f) ? radius : 0.1f;
        lambda1 = mid + rr;
        lambda2 = mid - rr;
    }
    float dvx = offDiagonal, dvy = lambda1 - diagonal1;
    const float dl = sqrtf(dvx * dvx + dvy * dvy);
    dvx = dvx / dl; dvy = dvy / dl;
    const float s1 = sqrtf(2.0f * lambda1), s2 = sqrtf(2.0f * lambda2);
    const float v1x = s1 * dvx, v1y = s1 * dvy, v2x = s2 * dvy, v2y = s2 * (-dvx);
    const float vcx = p0 / p3, vcy = p1 / p3;

    const float dx = x - U.cam.view_inv[12], dy = y - U.cam.view_inv[13], dz = z - U.cam.view_inv[14];
    const float dlen = sqrtf(dx * dx + dy * dy + dz * dz);
    V3 col = evaluate_sh(dx / dlen, dy / dlen, dz / dlen, sh, U.rs.max_sh_deg);
    col.x = (col.x > 0.f) ? col.x : 0.f;
    col.y = (col.y > 0.f) ? col.y : 0.f;
    col.z = (col.z > 0.f) ? col.z : 0.f;

    const float vw = U.cam.viewport[0], vh = U.cam.viewport[1];
    o.splat[0] = pack2h(v1x / vw, v1y / vh);
    o.splat[1] = pack2h(v2x / vw, v2y / vh);
    o.splat[2] = pack2h(vcx, vcy);
    o.splat[3] = pack2h(col.x, col.y);
    o.splat[4] = pack2h(col.z, opacity);

    const float znear = -U.cam.proj[3 * 4 + 2] / U.cam.proj[2 * 4 + 2];
    const float zfar = -U.cam.proj[3 * 4 + 2] / (U.cam.proj[2 * 4 + 2] - 1.f);
    if (!COMPRESSED) {
        o.key = __float_as_uint(zfar - p2);
    } else {
        float kf = 16777215.f - (p2 - znear) / (zfar - znear) * 16777215.f;
        uint32_t k;
        if (!(kf > 0.f)) k = 0u; else if (kf >= 4294967296.f) k = 0xffffffffu; else k = (uint32_t)kf;
        o.key = k;
    }

    // ---- tile rectangle of the STORED (f16-rounded) splat; must equal oracle wso_tile_rects ----
    {
        const float hv1x = half_lo(o.splat[0]), hv1y = half_hi(o.splat[0]);
        const float hv2x = half_lo(o.splat[1]), hv2y = half_hi(o.splat[1]);
        const float hcx = half_lo(o.splat[2]), hcy = half_hi(o.splat[2]);
        const float fw = (float)U.width, fh = (float)U.height;
        const float ex = FOOTPRINT_R * (fw * sqrtf(hv1x * hv1x + hv2x * hv2x));
        const float ey = FOOTPRINT_R * (fh * sqrtf(hv1y * hv1y + hv2y * hv2y));
        const float pcx = (hcx + 1.f) * 0.5f * fw;
        const float pcy = (1.f - hcy) * 0.5f * fh;
        const float fx0 = floorf((pcx - ex - 0.5f - RECT_PAD) * 0.0625f), fx1 = floorf((pcx + ex - 0.5f + RECT_PAD) * 0.0625f);
        const float fy0 = floorf((pcy - ey - 0.5f - RECT_PAD) * 0.0625f), fy1 = floorf((pcy + ey - 0.5f + RECT_PAD) * 0.0625f);
        o.rect_xy = 0u; o.rect_wh = 0u;
        if (fx0 == fx0 && fx1 == fx1 && fy0 == fy0 && fy1 == fy1) {
            const float mx = (float)(U.tiles_x - 1u), my = (float)(U.tiles_y - 1u);
            const float lx = fx0 < 0.f ? 0.f : fx0, ly = fy0 < 0.f ? 0.f : fy0;
            const float hx1 = fx1 > mx ? mx : fx1, hy1 = fy1 > my ? my : fy1;
            if (!(hx1 < lx || hy1 < ly)) {
                uint32_t x0 = (uint32_t)lx, y0 = (uint32_t)ly, x1 = (uint32_t)hx1, y1 = (uint32_t)hy1;
                o.rect_xy = x0 | (y0 << 16);
                o.rect_wh = (x1 - x0 + 1u) | ((y1 - y0 + 1u) << 16);
            }
        }
    }
}

template <bool COMPRESSED>
__global__ void __launch_bounds__(PP_THREADS, 2)
preprocess_kernel(PreprocessArgs a)
{
    constexpr uint32_t REC = COMPRESSED ? 24u : 28u;
    constexpr uint32_t REC_WORDS = REC / 4u;
    __shared__ __align__(128) uint32_t s_rec[PP_THREADS * REC_WORDS];
    __shared__ __align__(16) uint32_t s_splat[PP_THREADS * 5];
    __shared__ uint32_t s_key[PP_THREADS];
    __shared__ uint2 s_rect[PP_THREADS];
    __shared__ uint32_t s_hist[4 * 256];
    __shared__ FrameUniforms s_u;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_warp_cnt[PP_WARPS];
    __shared__ uint32_t s_part, s_base;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    {   // uniforms -> smem
        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.uniforms);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_u);
        for (unsigned i = tid; i < sizeof(FrameUniforms) / 4u; i += PP_THREADS) dst[i] = src[i];
    }
    for (unsigned i = tid; i < 4u * 256u; i += PP_THREADS) s_hist[i] = 0u;
    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    __syncthreads();
    const FrameUniforms &U = s_u;
    const uint32_t n = U.num_points;
    const uint32_t nparts = (n + PP_THREADS - 1u) / PP_THREADS;
    uint32_t parity = 0;

    for (;;) {
        if (tid == 0) s_part = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t part = s_part;
        if (part >= nparts) break;

        // ---- stage the partition's AoS records: one bulk copy (buffer is padded to a full partition)
        if (tid == 0) {
            fence_proxy_async();   // order prior generic-proxy reads of s_rec before the async-proxy write
            mbar_arrive_expect_tx(&s_bar, PP_THREADS * REC);
            bulk_g2s(s_rec, a.gaussians + (size_t)part * (PP_THREADS * REC), PP_THREADS * REC, &s_bar);
        }
        mbar_wait(&s_bar, parity);
        parity ^= 1u;

        const uint32_t idx = part * PP_THREADS + tid;
        bool vis = false;
        Stage1 o;
        o.key = 0u; o.rect_xy = 0u; o.rect_wh = 0u;
        o.splat[0] = o.splat[1] = o.splat[2] = o.splat[3] = o.splat[4] = 0u;
        if (idx < n) {
            const uint32_t *rec = s_rec + tid * REC_WORDS;
            const float x = __uint_as_float(rec[0]), y = __uint_as_float(rec[1]), z = __uint_as_float(rec[2]);
            // clip box (:177) -- any(xyz < min) || any(xyz > max)
            bool keep = !(x < U.rs.clip_min[0] || y < U.rs.clip_min[1] || z < U.rs.clip_min[2] ||
                          x > U.rs.clip_max[0] || y > U.rs.clip_max[1] || z > U.rs.clip_max[2]);
            float cs[4], pp[4];
            if (keep) {
                const float *view = U.cam.view, *proj = U.cam.proj;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    float acc = view[0 * 4 + r] * x;
                    acc = acc + view[1 * 4 + r] * y;
                    acc = acc + view[2 * 4 + r] * z;
                    acc = acc + view[3 * 4 + r] * 1.f;
                    cs[r] = acc;
                }
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    float acc = proj[0 * 4 + r] * cs[0];
                    acc = acc + proj[1 * 4 + r] * cs[1];
                    acc = acc + proj[2 * 4 + r] * cs[2];
                    acc = acc + proj[3 * 4 + r] * cs[3];
                    pp[r] = acc;
                }
                const float bounds = 1.2f * pp[3];
                const float zz = pp[2] / pp[3];
                if (!COMPRESSED) {
                    if (zz <= 0.f || zz >= 1.f || pp[0] < -bounds || pp[0] > bounds || pp[1] < -bounds || pp[1] > bounds) keep = false;
                } else {
                    if (zz < 0.f || zz > 1.f || pp[0] < -bounds || pp[0] > bounds || pp[1] < -bounds || pp[1] > bounds) keep = false;
                }
            }
            if (keep) {
                vis = true;
                if (!COMPRESSED) {
                    const float opacity = half_lo(rec[3]);
                    const float cov6[6] = {half_lo(rec[4]), half_hi(rec[4]), half_lo(rec[5]),
                                           half_hi(rec[5]), half_lo(rec[6]), half_hi(rec[6])};
                    ShRaw sh;
                    const uint8_t *sp = a.sh_coefs + (size_t)idx * 96u;
                    const uint32_t deg = U.rs.max_sh_deg;
                    ldg256(sp, sh.w);
                    if (deg > 1u) ldg256(sp + 32, sh.w + 8); else {
#pragma unroll
                        for (int i = 8; i < 16; i++) sh.w[i] = 0u;
                    }
                    if (deg > 2u) ldg256(sp + 64, sh.w + 16); else {
#pragma unroll
                        for (int i = 16; i < 24; i++) sh.w[i] = 0u;
                    }
                    project_tail<false>(U, x, y, z, cs[0], cs[1], cs[2], pp[0], pp[1], pp[2], pp[3], cov6, opacity, sh, o);
                } else {
                    const uint32_t os = rec[3];
                    const int8_t q_op = (int8_t)(os & 0xffu), q_sf = (int8_t)((os >> 8) & 0xffu);
                    const uint32_t geo_idx = rec[4], sh_idx = rec[5];
                    const float opacity = ((float)q_op - (float)U.quant.opacity.zero_point) * U.quant.opacity.scale;
                    const float sfac = expf(((float)q_sf - (float)U.quant.scaling_factor.zero_point) * U.quant.scaling_factor.scale);
                    const float s2 = sfac * sfac;
                    const uint32_t *cw = reinterpret_cast<const uint32_t *>(a.covars + (size_t)geo_idx * 12u);
                    const uint32_t w0 = __ldg(cw), w1 = __ldg(cw + 1), w2 = __ldg(cw + 2);
                    const float cov6[6] = {half_lo(w0) * s2, half_hi(w0) * s2, half_lo(w1) * s2,
                                           half_hi(w1) * s2, half_lo(w2) * s2, half_hi(w2) * s2};
                    const uint32_t ncoef = (U.file_sh_deg + 1u) * (U.file_sh_deg + 1u);
                    ShQuant sh;
                    sh.p = reinterpret_cast<const int8_t *>(a.sh_coefs) + (size_t)sh_idx * ncoef * 3u;
                    sh.dc = U.quant.color_dc; sh.rest = U.quant.color_rest;
                    project_tail<true>(U, x, y, z, cs[0], cs[1], cs[2], pp[0], pp[1], pp[2], pp[3], cov6, opacity, sh, o);
                }
            }
        }

        // ---- deterministic compaction: ballot, block scan, decoupled look-back ----
        const unsigned bal = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t warp_off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < PP_WARPS; w++) {
            uint32_t c = s_warp_cnt[w];
            if (w < (int)warp) warp_off += c;
            total += c;
        }
        if (warp == 0) {
            if (lane == 0 && part > 0u) st_relaxed(a.scan_status + part, LB_AGGREGATE | total);
            uint32_t excl = (part > 0u) ? lookback_warp(a.scan_status, part, &a.counters->error_flags) : 0u;
            if (lane == 0) {
                st_relaxed(a.scan_status + part, LB_PREFIX | (excl + total));
                s_base = excl;
                if (part == nparts - 1u) a.counters->num_visible = excl + total;
            }
        }
        const uint32_t local = warp_off + __popc(bal & lanemask_lt());
        if (vis) {
#pragma unroll
            for (int k = 0; k < 5; k++) s_splat[local * 5u + k] = o.splat[k];
            s_key[local] = o.key;
            s_rect[local] = make_uint2(o.rect_xy, o.rect_wh);
        }
        // depth-key digit histograms for the onesweep (warp-aggregated: the upper digits of
        // a float key take only a handful of values, plain smem atomics would serialise)
        {
#pragma unroll
            for (int d = 0; d < 4; d++) {
                const uint32_t dig = (o.key >> (8 * d)) & 255u;
                const unsigned peers = __match_any_sync(0xffffffffu, vis ? dig : 0xffffffffu);
                if (vis && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&s_hist[d * 256 + dig], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
        const uint32_t base = s_base;
        // ---- coalesced output streams ----
        for (uint32_t i = tid; i < total * 5u; i += PP_THREADS) a.splats[(size_t)base * 5u + i] = s_splat[i];
        if (tid < total) {
            a.depth_keys[base + tid] = s_key[tid];
            a.slot_vals[base + tid] = base + tid;           // payload = slot, preprocess.wgsl:274
            a.rects[base + tid] = s_rect[tid];
        }
        __syncthreads();
    }

    // flush the CTA's digit histograms
    for (unsigned i = tid; i < 4u * 256u; i += PP_THREADS) {
        uint32_t c = s_hist[i];
        if (c) atomicAdd(a.hist + i, c);
    }
}

}  // namespace

cudaError_t launch_preprocess(const PreprocessArgs &a, bool compressed, int grid, cudaStream_t stream)
{
    if (compressed) preprocess_kernel<true><<<grid, PP_THREADS, 0, stream>>>(a);
    else preprocess_kernel<false><<<grid, PP_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

int preprocess_blocks_per_sm(bool compressed)
{
    int nb = 0;
    if (compressed) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, preprocess_kernel<true>, PP_THREADS, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, preprocess_kernel<false>, PP_THREADS, 0);
    return nb > 0 ? nb : 1;
}

}  // namespace ws
