// 2D-feature -> 3D-Gaussian fusion.  Contract: PointCloudToImageMapper.compute_mapping
// (reference dataset/fusion_utils.py:30-78) and the per-view gather / accumulate / final divide of
// fusion.py:127-148.  The reference does this per view on the CPU in numpy — GPU->CPU copies of
// xyz, the view matrix and the rendered depth, a fancy-index gather of a (C,h,w) map for ALL P
// points, then a P x C host->device copy.  Here one view is: projection + visibility test on the
// device (float64 like numpy), one tiled transpose of the feature map to pixel-major so that a
// Gaussian's C-vector is one contiguous row, and a gather-accumulate into the fp32 sums.
//
// Numerics: projection in float64 with numpy's promotion rules (float32 inputs widened), pixel =
// round-half-to-even; the per-Gaussian sums add the views in call order in fp32 exactly like
// `_features_semantic[mask] += features_mapping[mask]`, so they are bit-identical to the reference.
#include <cuda_fp16.h>
#include "common.cuh"

namespace sgb {

namespace {

constexpr double kSurfaceInit = 999999.0;  // fusion_utils.py:58

struct Proj {
    double z;
    long long u, v;  // pi[0], pi[1]
    bool inside;
};

__device__ __forceinline__ long long round_to_ll(double r) {
    // np.round(...).astype(int): non-finite / out-of-range doubles become INT64_MIN on x86-64
    if (!(fabs(r) < 9.2e18)) return (long long)0x8000000000000000ull;
    return __double2ll_rn(r);  // round-half-to-even like np.round
}

__device__ __forceinline__ Proj project(const sgb_fusion_view& v, const float* __restrict__ w2c_t, int i) {
    // p = W2C @ [x,y,z,1] with W2C = world_to_camera.T (fusion_utils.py:42-45): p[k] = sum_j w2c_t[j][k]*c[j]
    const double x = (double)v.xyz[3 * (size_t)i], y = (double)v.xyz[3 * (size_t)i + 1],
                 z = (double)v.xyz[3 * (size_t)i + 2];
    double p[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
        p[k] = (double)w2c_t[0 * 4 + k] * x + (double)w2c_t[1 * 4 + k] * y + (double)w2c_t[2 * 4 + k] * z +
               (double)w2c_t[3 * 4 + k];
    Proj r;
    r.z = p[2];
    const double pu = (p[0] * v.fx) / p[2] + v.cx;  // fusion_utils.py:46-47
    const double pv = (p[1] * v.fy) / p[2] + v.cy;
    r.u = round_to_ll(rint(pu));
    r.v = round_to_ll(rint(pv));
    r.inside = (r.u >= v.cut_bound) && (r.v >= v.cut_bound) && (r.u < v.w - v.cut_bound) &&
               (r.v < v.h - v.cut_bound);  // fusion_utils.py:50-55
    return r;
}

__global__ void surface_init_kernel(int n, double* zbuf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) zbuf[i] = kSurfaceInit;
}

// fusion_utils.py:57-61: z-buffer of the points themselves (min over p.z > 0.2 inside the image)
__global__ void surface_zbuf_kernel(sgb_fusion_view v, double* zbuf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.P) return;
    Proj p = project(v, v.world_to_camera, i);
    if (p.z > 0.2 && p.inside) {
        // positive doubles order like their bit patterns
        atomicMin(reinterpret_cast<unsigned long long*>(zbuf + p.v * v.w + p.u),
                  (unsigned long long)__double_as_longlong(p.z));
    }
}

// Visibility of Gaussian i in this view -> linear pixel index v*w+u, or -1.
__device__ __forceinline__ int visible_pixel(const sgb_fusion_view& v, const double* zbuf, int i, Proj& p) {
    p = project(v, v.world_to_camera, i);
    bool vis = p.inside;
    if (vis) {
        if (v.depth_mode == SGB_DEPTH_NONE) {
            vis = p.z > 0;  // fusion_utils.py:70-72
        } else {
            double dcur, thr;
            const size_t o = (size_t)p.v * v.w + p.u;
            if (v.depth_mode == SGB_DEPTH_F32) {
                const float df = static_cast<const float*>(v.depth)[o];
                dcur = (double)df;
                thr = (double)((float)v.vis_thres * df);  // python float * float32 array -> float32
            } else {
                dcur = v.depth_mode == SGB_DEPTH_F64 ? static_cast<const double*>(v.depth)[o] : zbuf[o];
                thr = v.vis_thres * dcur;
            }
            vis = fabs(dcur - p.z) <= thr;  // fusion_utils.py:63-69
        }
    }
    return vis ? (int)(p.v * v.w + p.u) : -1;
}

__global__ void fusion_map_kernel(sgb_fusion_view v, const double* zbuf, long long* __restrict__ mapping) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.P) return;
    Proj p;
    const int pix = visible_pixel(v, zbuf, i, p);
    mapping[3 * (size_t)i + 0] = pix >= 0 ? p.v : 0;  // fusion_utils.py:73-75
    mapping[3 * (size_t)i + 1] = pix >= 0 ? p.u : 0;
    mapping[3 * (size_t)i + 2] = pix >= 0 ? 1 : 0;
}

__global__ void fusion_pix_kernel(sgb_fusion_view v, const double* zbuf, int* __restrict__ pix_of,
                                  float* __restrict__ count, int* __restrict__ n_visible) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int pix = -1;
    if (i < v.P) {
        Proj p;
        pix = visible_pixel(v, zbuf, i, p);
        pix_of[i] = pix;
        if (pix >= 0) count[i] += 1.0f;  // gaussians._times[mask] += 1, fusion.py:143
    }
    if (n_visible) {
        unsigned m = __ballot_sync(0xffffffffu, pix >= 0);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_visible, __popc(m));
    }
}

// (C, npix) -> (npix, C) tiled transpose, 32x32 tiles through shared memory.
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int C, int npix) {
    __shared__ T tile[32][33];
    const int px0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, px = px0 + threadIdx.x;
        if (c < C && px < npix) tile[r][threadIdx.x] = in[(size_t)c * npix + px];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int px = px0 + r, c = c0 + threadIdx.x;
        if (c < C && px < npix) out[(size_t)px * C + c] = tile[threadIdx.x][r];
    }
}

__device__ __forceinline__ float to_f32(__half h) { return __half2float(h); }
__device__ __forceinline__ float to_f32(float f) { return f; }

// One warp per Gaussian (grid-stride): feat_sum[g, :] += featT[pix(g), :] for visible g.
// fp16 maps: 64 channels x 64 pixels per CTA, every global access is a 4-byte pair (128 B per warp row instead
// of the 64 B of the element-wise kernel above).  Shared cell (r, j) = the pixel pair (2j, 2j+1) of channel r.
// Requires even npix and C and 4-byte aligned bases.
__global__ void __launch_bounds__(256) transpose_half2_kernel(const __half* __restrict__ in, __half* __restrict__ out,
                                                              int C, int npix) {
    __shared__ uint32_t tile[64][33];
    const int px0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 64; r += 8) {
        const int c = c0 + r, px = px0 + 2 * tx;
        uint32_t v = 0u;
        if (c < C && px < npix) v = __ldg(reinterpret_cast<const uint32_t*>(in + (size_t)c * npix + px));
        tile[r][tx] = v;
    }
    __syncthreads();
    const int c = c0 + 2 * tx;
    if (c >= C) return;
    for (int r = ty; r < 64; r += 8) {
        const int px = px0 + r;
        if (px >= npix) break;
        const uint32_t a = tile[2 * tx][r >> 1], b = tile[2 * tx + 1][r >> 1];
        const uint32_t lo = (r & 1) ? (a >> 16) : (a & 0xFFFFu), hi = (r & 1) ? (b >> 16) : (b & 0xFFFFu);
        *reinterpret_cast<uint32_t*>(out + (size_t)px * C + c) = lo | (hi << 16);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) fusion_gather_kernel(int P, int C, const int* __restrict__ pix_of,
                                                            const T* __restrict__ featT,
                                                            float* __restrict__ feat_sum) {
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < P; g += warps_total) {
        const int pix = pix_of[g];
        if (pix < 0) continue;
        const T* src = featT + (size_t)pix * C;
        float* dst = feat_sum + (size_t)g * C;
        if ((C & 3) == 0) {
            for (int k = lane * 4; k < C; k += 128) {
                float4 acc = *reinterpret_cast<const float4*>(dst + k);
                acc.x += to_f32(src[k]);
                acc.y += to_f32(src[k + 1]);
                acc.z += to_f32(src[k + 2]);
                acc.w += to_f32(src[k + 3]);
                *reinterpret_cast<float4*>(dst + k) = acc;
            }
        } else {
            for (int k = lane; k < C; k += 32) dst[k] += to_f32(src[k]);
        }
    }
}

// fusion.py:146-147
__global__ void fusion_normalize_kernel(int P, int C, float* __restrict__ feat_sum, float* __restrict__ count) {
    const size_t n = (size_t)P * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const size_t g = e / C;
        float t = count[g];
        if (t == 0.f) t = 1e-5f;
        feat_sum[e] = feat_sum[e] / t;
    }
}
__global__ void fusion_fix_count_kernel(int P, float* __restrict__ count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P && count[i] == 0.f) count[i] = 1e-5f;
}

int prepare_zbuf(sgb_ctx* ctx, const sgb_fusion_view& v, size_t extra, double** zbuf, char** extra_ptr,
                 cudaStream_t s) {
    const size_t npix = (size_t)v.w * v.h;
    const size_t zb = (v.depth_mode == SGB_DEPTH_SURFACE) ? align_up(npix * sizeof(double)) : 0;
    int rc = ctx->misc.ensure(zb + extra + 256);
    if (rc) return rc;
    *zbuf = zb ? (double*)ctx->misc.p : nullptr;
    *extra_ptr = (char*)ctx->misc.p + zb;
    if (zb) {
        surface_init_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>((int)npix, *zbuf);
        surface_zbuf_kernel<<<(v.P + 255) / 256, 256, 0, s>>>(v, *zbuf);
        SGB_LAUNCH_CHECK("surface_zbuf_kernel", 0, s);
    }
    return SGB_OK;
}

int check_view(const sgb_fusion_view* v) {
    if (!v || v->P < 0 || (v->P > 0 && !v->xyz) || !v->world_to_camera || v->w <= 0 || v->h <= 0) {
        set_error("fusion view: null pointer or non-positive size");
        return SGB_E_INVALID;
    }
    if ((v->depth_mode == SGB_DEPTH_F32 || v->depth_mode == SGB_DEPTH_F64) && !v->depth) {
        set_error("fusion view: depth_mode needs a depth map");
        return SGB_E_INVALID;
    }
    if (v->depth_mode < 0 || v->depth_mode > 3) {
        set_error("fusion view: unknown depth_mode %d", v->depth_mode);
        return SGB_E_INVALID;
    }
    return SGB_OK;
}

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" int sgb_fusion_map(sgb_ctx* ctx, const sgb_fusion_view* v, int64_t* mapping, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_view(v);
    if (rc) return rc;
    if (v->P == 0) return SGB_OK;
    if (!ctx || !mapping) { set_error("sgb_fusion_map: null ctx/mapping"); return SGB_E_INVALID; }
    double* zbuf; char* extra;
    rc = prepare_zbuf(ctx, *v, 0, &zbuf, &extra, s);
    if (rc) return rc;
    fusion_map_kernel<<<(v->P + 255) / 256, 256, 0, s>>>(*v, zbuf, (long long*)mapping);
    SGB_LAUNCH_CHECK("fusion_map_kernel", 0, s);
    return SGB_OK;
}

extern "C" int sgb_fusion_accumulate(sgb_ctx* ctx, const sgb_fusion_view* v, const void* features, int32_t C,
                                     int32_t feat_dtype, float* feat_sum, float* count, int32_t* n_visible_dev,
                                     void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_view(v);
    if (rc) return rc;
    if (!ctx || !features || !feat_sum || !count || C <= 0 || (feat_dtype != SGB_FEAT_F16 && feat_dtype != SGB_FEAT_F32)) {
        set_error("sgb_fusion_accumulate: bad arguments");
        return SGB_E_INVALID;
    }
    if (n_visible_dev) SGB_CUDA(cudaMemsetAsync(n_visible_dev, 0, sizeof(int32_t), s));
    if (v->P == 0) return SGB_OK;
    const size_t npix = (size_t)v->w * v->h;
    const size_t esz = feat_dtype == SGB_FEAT_F16 ? 2 : 4;
    const size_t tbytes = align_up(npix * C * esz);
    const size_t pbytes = align_up(sizeof(int) * (size_t)v->P);
    double* zbuf; char* extra;
    rc = prepare_zbuf(ctx, *v, tbytes + pbytes, &zbuf, &extra, s);
    if (rc) return rc;
    void* featT = extra;
    int* pix_of = (int*)(extra + tbytes);

    {
        StageTimer t(ctx, ST_FUSION_PROJECT, s);
        fusion_pix_kernel<<<(v->P + 255) / 256, 256, 0, s>>>(*v, zbuf, pix_of, count, n_visible_dev);
        SGB_LAUNCH_CHECK("fusion_pix_kernel", 0, s);
    }
    dim3 tgrid((unsigned)((npix + 31) / 32), (unsigned)((C + 31) / 32)), tblock(32, 8);
    const int gblocks = 148 * 8;
    {
        StageTimer t(ctx, ST_FUSION_TRANSPOSE, s);
        const bool pairs = feat_dtype == SGB_FEAT_F16 && (npix % 2 == 0) && (C % 2 == 0) &&
                           ((reinterpret_cast<uintptr_t>(features) & 3) == 0) && ((reinterpret_cast<uintptr_t>(featT) & 3) == 0);
        if (pairs)
            transpose_half2_kernel<<<dim3((unsigned)((npix + 63) / 64), (unsigned)((C + 63) / 64)), tblock, 0, s>>>(
                (const __half*)features, (__half*)featT, C, (int)npix);
        else if (feat_dtype == SGB_FEAT_F16)
            transpose_kernel<__half><<<tgrid, tblock, 0, s>>>((const __half*)features, (__half*)featT, C, (int)npix);
        else
            transpose_kernel<float><<<tgrid, tblock, 0, s>>>((const float*)features, (float*)featT, C, (int)npix);
    }
    {
        StageTimer t(ctx, ST_FUSION_GATHER, s);
        if (feat_dtype == SGB_FEAT_F16)
            fusion_gather_kernel<__half><<<gblocks, 256, 0, s>>>(v->P, C, pix_of, (const __half*)featT, feat_sum);
        else
            fusion_gather_kernel<float><<<gblocks, 256, 0, s>>>(v->P, C, pix_of, (const float*)featT, feat_sum);
    }
    SGB_LAUNCH_CHECK("fusion_gather_kernel", 0, s);
    ctx->launches += 3;
    return SGB_OK;
}

extern "C" int sgb_fusion_normalize(int32_t P, int32_t C, float* feat_sum, float* count, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || C <= 0 || !feat_sum || !count) { set_error("sgb_fusion_normalize: bad arguments"); return SGB_E_INVALID; }
    if (P == 0) return SGB_OK;
    fusion_normalize_kernel<<<148 * 8, 256, 0, s>>>(P, C, feat_sum, count);
    fusion_fix_count_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, count);
    SGB_LAUNCH_CHECK("fusion_normalize_kernel", 0, s);
    return SGB_OK;
}
