// ingest.cu -- file formats -> GPU layouts (SURVEY.md section 8(f) N1 .ply, N2 .npz).
//
// .ply:
// Replaces PlyReader::read / read_line (io/ply.rs:50-100,165-195) and GenericGaussianPointCloud::new
// (io/mod.rs:63-105): the reference converts every vertex on ONE host thread (sigmoid, exp,
// quaternion normalise, build_cov, f16 rounding, SH transpose) and then uploads; here the raw
// vertex block is uploaded as it is in the file and one kernel converts it in place into the
// 28-B Gaussian records, the 96-B SH records and the xyz plane the render path consumes, plus the
// bounding box, centroid and second moments (plane fit) by block reductions.
// Compiled with -fmad=false: same operation order as the CPU oracle (oracle/ws_oracle.c: wso_ply_convert);
// expf differs from glibc by an ulp, which the f16 rounding hides except for rare 1-ulp cases.
//
// .npz (compressed 3DGS): replaces the array post-processing of NpzReader::read (io/npz.rs:58-225) and
// GenericGaussianPointCloud::new_compressed (io/mod.rs:107-150).  The zip/npy container is decoded by
// the caller (host I/O); the arrays arrive as plain pointers and three kernels assemble the 24-B records,
// interleave the i8 SH codebook and build the f16 covariance codebook from the quantised rotation / scaling.
#include "ws_device.cuh"
#include "ws_kernels.h"

namespace ws {

namespace {

__device__ __forceinline__ float load_f32(const uint8_t *p, bool big_endian)
{
    uint32_t u = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    if (big_endian) u = __byte_perm(u, 0, 0x0123);
    return __uint_as_float(u);
}
__device__ __forceinline__ unsigned short f2h(float v) { __half_raw r = __float2half_rn(v); return r.x; }

// Quaternion::normalize (q * (1 / |q|), |q|^2 = s*s + v.v, cgmath 0.18) then build_cov (utils.rs:194-204):
// R = Matrix3::from(q) (column-major R[c][r]), L = R * diag(s), M = L L^T; out = m00 m01 m02 m11 m12 m22
__device__ __forceinline__ void build_cov6(float qw, float qx, float qy, float qz, float s0, float s1, float s2, float out[6])
{
    {
        const float mag = sqrtf(qw * qw + (qx * qx + qy * qy + qz * qz));
        const float inv = 1.f / mag;
        qw = qw * inv; qx = qx * inv; qy = qy * inv; qz = qz * inv;
    }
    const float x2 = qx + qx, y2 = qy + qy, z2 = qz + qz;
    const float xx2 = x2 * qx, xy2 = x2 * qy, xz2 = x2 * qz, yy2 = y2 * qy, yz2 = y2 * qz, zz2 = z2 * qz;
    const float sy2 = y2 * qw, sz2 = z2 * qw, sx2 = x2 * qw;
    float R[3][3];
    R[0][0] = 1.f - yy2 - zz2; R[0][1] = xy2 + sz2;       R[0][2] = xz2 - sy2;
    R[1][0] = xy2 - sz2;       R[1][1] = 1.f - xx2 - zz2; R[1][2] = yz2 + sx2;
    R[2][0] = xz2 + sy2;       R[2][1] = yz2 - sx2;       R[2][2] = 1.f - xx2 - yy2;
    const float sc[3] = {s0, s1, s2};
    float L[3][3];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) L[c][r] = R[c][r] * sc[c];
    auto M = [&](int c, int r) { float m = L[0][r] * L[0][c]; m = m + L[1][r] * L[1][c]; m = m + L[2][r] * L[2][c]; return m; };
    out[0] = M(0, 0); out[1] = M(0, 1); out[2] = M(0, 2); out[3] = M(1, 1); out[4] = M(1, 2); out[5] = M(2, 2);
}

// ordered-int encoding so that atomicMin/atomicMax on u32 order like floats
__device__ __forceinline__ uint32_t f2ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }

// per-thread bounding box + first/second moments, reduced through shared memory into the global accumulators
struct CloudStats {
    double acc[9];
    uint32_t mn[3], mx[3];
    __device__ CloudStats()
    {
#pragma unroll
        for (int q = 0; q < 9; q++) acc[q] = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) { mn[d] = 0xffffffffu; mx[d] = 0u; }
    }
    __device__ void add(float x, float y, float z)
    {
        const float p[3] = {x, y, z};
#pragma unroll
        for (int d = 0; d < 3; d++) { const uint32_t o = f2ord(p[d]); mn[d] = o < mn[d] ? o : mn[d]; mx[d] = o > mx[d] ? o : mx[d]; acc[d] += (double)p[d]; }
        acc[3] += (double)x * x; acc[4] += (double)x * y; acc[5] += (double)x * z;
        acc[6] += (double)y * y; acc[7] += (double)y * z; acc[8] += (double)z * z;
    }
    // all threads of the CTA must call this
    __device__ void flush(double *sums, uint32_t *minmax)
    {
        __shared__ double s_sum[9];
        __shared__ uint32_t s_min[3], s_max[3];
        const unsigned tid = threadIdx.x;
        if (tid < 9) s_sum[tid] = 0.0;
        if (tid < 3) { s_min[tid] = 0xffffffffu; s_max[tid] = 0u; }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 9; q++) atomicAdd(&s_sum[q], acc[q]);
#pragma unroll
        for (int d = 0; d < 3; d++) { atomicMin(&s_min[d], mn[d]); atomicMax(&s_max[d], mx[d]); }
        __syncthreads();
        if (tid < 9) atomicAdd(&sums[tid], s_sum[tid]);
        if (tid < 3) { atomicMin(&minmax[tid], s_min[tid]); atomicMax(&minmax[3 + tid], s_max[tid]); }
    }
};

__global__ void __launch_bounds__(256)
ply_convert_kernel(PlyConvertArgs a)
{
    const unsigned tid = threadIdx.x;
    const uint32_t ncoef = (a.sh_deg + 1u) * (a.sh_deg + 1u);
    CloudStats st;
    for (uint32_t i = blockIdx.x * 256u + tid; i < a.n; i += gridDim.x * 256u) {
        const uint8_t *v = a.vertices + (size_t)i * a.stride_bytes;
        auto F = [&](uint32_t k) { return load_f32(v + 4u * k, a.big_endian != 0); };
        const float x = F(0), y = F(1), z = F(2);                 // 3..5 = normals, skipped (io/ply.rs:57-61)
        // SH: dc, then rest stored channel-major [3][C-1] in the file (io/ply.rs:63-75)
        unsigned short sh[48];
#pragma unroll
        for (int q = 0; q < 48; q++) sh[q] = 0;
        for (int ch = 0; ch < 3; ch++) sh[ch] = f2h(F(6u + ch));
        for (uint32_t c = 0; c + 1u < ncoef; c++)
            for (uint32_t ch = 0; ch < 3u; ch++) sh[(c + 1u) * 3u + ch] = f2h(F(9u + ch * (ncoef - 1u) + c));
        uint32_t k = 9u + (ncoef - 1u) * 3u;
        const float op_raw = F(k);
        float opacity;                                             // utils.rs:206-212 numerically stable sigmoid
        if (op_raw >= 0.f) opacity = 1.f / (1.f + expf(-op_raw));
        else { const float e = expf(op_raw); opacity = e / (1.f + e); }
        const float s0 = expf(F(k + 1)), s1 = expf(F(k + 2)), s2 = expf(F(k + 3));
        float m6[6];
        build_cov6(F(k + 4), F(k + 5), F(k + 6), F(k + 7), s0, s1, s2, m6);
        // ---- write the GPU layouts
        uint32_t *rec = reinterpret_cast<uint32_t *>(a.gaussians + (size_t)i * 28u);
        rec[0] = __float_as_uint(x); rec[1] = __float_as_uint(y); rec[2] = __float_as_uint(z);
        rec[3] = (uint32_t)f2h(opacity);
        rec[4] = (uint32_t)f2h(m6[0]) | ((uint32_t)f2h(m6[1]) << 16);
        rec[5] = (uint32_t)f2h(m6[2]) | ((uint32_t)f2h(m6[3]) << 16);
        rec[6] = (uint32_t)f2h(m6[4]) | ((uint32_t)f2h(m6[5]) << 16);
        uint32_t *shw = reinterpret_cast<uint32_t *>(a.sh_coefs + (size_t)i * 96u);
#pragma unroll
        for (int q = 0; q < 24; q++) shw[q] = (uint32_t)sh[2 * q] | ((uint32_t)sh[2 * q + 1] << 16);
        a.xyz[(size_t)i * 3u] = x; a.xyz[(size_t)i * 3u + 1] = y; a.xyz[(size_t)i * 3u + 2] = z;
        st.add(x, y, z);
    }
    st.flush(a.sums, a.minmax);
}

// ---- .npz (compressed) ---------------------------------------------------------------------------
// GaussianCompressed records (io/npz.rs:170-189): xyz f16 -> f32, opacity i8, scale_factor i8 (0 without
// scaling_factor), geometry_idx / sh_idx = the index arrays or the identity.
__global__ void __launch_bounds__(256)
c3dgs_records_kernel(C3dgsArgs a)
{
    CloudStats st;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < a.n; i += gridDim.x * 256u) {
        __half_raw hx, hy, hz;
        hx.x = a.xyz_f16[(size_t)i * 3u]; hy.x = a.xyz_f16[(size_t)i * 3u + 1]; hz.x = a.xyz_f16[(size_t)i * 3u + 2];
        const float x = __half2float(__half(hx)), y = __half2float(__half(hy)), z = __half2float(__half(hz));
        uint32_t *rec = reinterpret_cast<uint32_t *>(a.gaussians + (size_t)i * 24u);
        rec[0] = __float_as_uint(x); rec[1] = __float_as_uint(y); rec[2] = __float_as_uint(z);
        const uint32_t op = (uint8_t)a.opacity[i], sf = a.scaling_factor ? (uint8_t)a.scaling_factor[i] : 0u;
        rec[3] = op | (sf << 8);
        rec[4] = a.gaussian_indices ? (uint32_t)a.gaussian_indices[i] : i;
        rec[5] = a.feature_indices ? (uint32_t)a.feature_indices[i] : i;
        a.xyz[(size_t)i * 3u] = x; a.xyz[(size_t)i * 3u + 1] = y; a.xyz[(size_t)i * 3u + 2] = z;
        st.add(x, y, z);
    }
    st.flush(a.sums, a.minmax);
}

// SH codebook: entry = dc[3] then rest[3C-3] (io/npz.rs:193-205)
__global__ void __launch_bounds__(256)
c3dgs_sh_kernel(C3dgsArgs a)
{
    const uint32_t per = (a.sh_deg + 1u) * (a.sh_deg + 1u) * 3u, rest = per - 3u;
    const uint64_t total = (uint64_t)a.num_features * per;
    for (uint64_t t = (uint64_t)blockIdx.x * 256u + threadIdx.x; t < total; t += (uint64_t)gridDim.x * 256u) {
        const uint64_t e = t / per; const uint32_t j = (uint32_t)(t - e * per);
        a.sh_out[t] = j < 3u ? a.features_dc[e * 3u + j] : a.features_rest[e * rest + (j - 3u)];
    }
}

// covariance codebook (io/npz.rs:99-131,206-211): scaling = exp(deq) without a scaling factor, else
// normalize(max(deq, 0)); rotation = normalize(deq); build_cov -> 6 x f16
__global__ void __launch_bounds__(256)
c3dgs_covars_kernel(C3dgsArgs a)
{
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < a.num_covars; i += gridDim.x * 256u) {
        float s[3], q[4];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const float v = ((float)a.scaling[(size_t)i * 3u + d] - a.scaling_zero_point) * a.scaling_scale;
            s[d] = a.scaling_factor ? fmaxf(v, 0.f) : expf(v);
        }
        if (a.scaling_factor) {   // Vector3::normalize: v * (1 / |v|)
            const float inv = 1.f / sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
            s[0] = s[0] * inv; s[1] = s[1] * inv; s[2] = s[2] * inv;
        }
#pragma unroll
        for (int d = 0; d < 4; d++) q[d] = ((float)a.rotation[(size_t)i * 4u + d] - a.rotation_zero_point) * a.rotation_scale;
        float m6[6];
        build_cov6(q[0], q[1], q[2], q[3], s[0], s[1], s[2], m6);
        uint32_t *o = reinterpret_cast<uint32_t *>(a.covars + (size_t)i * 12u);
        o[0] = (uint32_t)f2h(m6[0]) | ((uint32_t)f2h(m6[1]) << 16);
        o[1] = (uint32_t)f2h(m6[2]) | ((uint32_t)f2h(m6[3]) << 16);
        o[2] = (uint32_t)f2h(m6[4]) | ((uint32_t)f2h(m6[5]) << 16);
    }
}

// Index validation of compressed (24-B) records: geometry_idx / sh_idx come from an untrusted file or caller.  wgpu's
// storage buffers are bounds-checked (an out-of-range index reads zeros); a CUDA gather is not -- one bad index would
// fault the context -- so clouds are checked ONCE at load time and rejected with WS_ERR_INVALID_ARGUMENT.
__global__ void __launch_bounds__(256)
validate_compressed_kernel(const uint8_t *__restrict__ gaussians, uint32_t n, uint32_t num_covars, uint32_t num_features, uint32_t *flag)
{
    bool bad = false;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
        const uint32_t *rec = reinterpret_cast<const uint32_t *>(gaussians + (size_t)i * 24u);
        bad = bad || (rec[4] >= num_covars) || (rec[5] >= num_features);
    }
    if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0) atomicOr(flag, 1u);
}

}  // namespace

cudaError_t launch_validate_compressed(const uint8_t *gaussians, uint32_t n, uint32_t num_covars, uint32_t num_features,
                                       uint32_t *flag, int max_grid, cudaStream_t stream)
{
    if (!n) return cudaSuccess;
    const uint64_t w = ((uint64_t)n + 255u) / 256u;
    validate_compressed_kernel<<<(unsigned)(w < (uint64_t)max_grid ? w : (uint64_t)max_grid), 256, 0, stream>>>(gaussians, n, num_covars, num_features, flag);
    return cudaGetLastError();
}

cudaError_t launch_ply_convert(const PlyConvertArgs &a, int grid, cudaStream_t stream)
{
    ply_convert_kernel<<<grid, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_c3dgs_convert(const C3dgsArgs &a, int max_grid, cudaStream_t stream)
{
    auto grid = [&](uint64_t items) { const uint64_t w = (items + 255u) / 256u; return (unsigned)(w < 1 ? 1 : (w < (uint64_t)max_grid ? w : (uint64_t)max_grid)); };
    if (a.n) c3dgs_records_kernel<<<grid(a.n), 256, 0, stream>>>(a);
    if (a.num_features) c3dgs_sh_kernel<<<grid((uint64_t)a.num_features * (a.sh_deg + 1u) * (a.sh_deg + 1u) * 3u), 256, 0, stream>>>(a);
    if (a.num_covars) c3dgs_covars_kernel<<<grid(a.num_covars), 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace ws
