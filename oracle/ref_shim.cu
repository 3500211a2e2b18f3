// TEST INFRASTRUCTURE — not part of the product.
//
// C-ABI shim around the UNMODIFIED reference CUDA rasterizer library
// (/root/reference/submodules/{channel,rgbd}-rasterization/cuda_rasterizer/*.cu), compiled
// where the sources lie by oracle/build.py into oracle/_ref/libref_*.so.  It replaces the
// reference's torch glue (rasterize_points.cu:38-223) with plain pointers so tests and
// bench.py can drive CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (cuda_rasterizer/rasterizer.h:23-84) without libtorch, and exposes the opaque
// geometry/binning/image state (rasterizer_impl.cu:155-194) for stage-by-stage bit-exact
// comparison.  Only tests/, __graft_entry__.smoke() and bench.py's reference leg load it.
//
// Build variants (oracle/build.py):  -DREF_RGBD selects the rgbd signature (out_depth, no
// debug / num_channels arguments).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <cuda_runtime.h>
#include "rasterizer.h"
#include "rasterizer_impl.h"
#include "config.h"

namespace {
struct Buf {
    char* p = nullptr;
    size_t cap = 0;
    char* grow(size_t n) {
        if (n > cap) {
            if (p) cudaFree(p);
            size_t want = n + (n >> 2) + 256;
            if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cap = 0; return nullptr; }
            cap = want;
        }
        return p;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct State {
    Buf geom, bin, img;
    int P = 0, R = 0, W = 0, H = 0;
};
char g_err[512] = "";
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err; }
int ref_num_channels_bwd() { return NUM_CHANNELS; }
int ref_is_rgbd() {
#ifdef REF_RGBD
    return 1;
#else
    return 0;
#endif
}

void* ref_state_new() { return new State(); }
void ref_state_free(void* s_) {
    State* s = (State*)s_;
    if (!s) return;
    s->geom.release(); s->bin.release(); s->img.release();
    delete s;
}

// Returns num_rendered (>=0) or -1 on error.  All array arguments are DEVICE pointers;
// absent optional inputs are nullptr (rasterizer_impl.cu:243,324).
int ref_forward(void* s_, int P, int D, int M, const float* bg, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                int prefiltered, int num_channels, float* out_color, float* out_depth,
                int* radii, int debug) {
    State* s = (State*)s_;
    s->P = P; s->W = W; s->H = H;
    try {
        auto gf = [s](size_t n) { return s->geom.grow(n); };
        auto bf = [s](size_t n) { return s->bin.grow(n); };
        auto imf = [s](size_t n) { return s->img.grow(n); };
#ifdef REF_RGBD
        (void)num_channels; (void)debug;
        s->R = CudaRasterizer::Rasterizer::forward(gf, bf, imf, P, D, M, bg, W, H, means3D, shs,
            colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp,
            viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, prefiltered != 0,
            out_color, out_depth, radii);
#else
        (void)out_depth;
        s->R = CudaRasterizer::Rasterizer::forward(gf, bf, imf, P, D, M, bg, W, H, means3D, shs,
            colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp,
            viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, prefiltered != 0,
            num_channels, out_color, radii, debug != 0);
#endif
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s", cudaGetErrorString(e)); return -1; }
        return s->R;
    } catch (const std::exception& ex) {
        snprintf(g_err, sizeof g_err, "%s", ex.what());
        return -1;
    }
}

// Gradient buffers must be zero-filled by the caller (rasterize_points.cu:157-165).
int ref_backward(void* s_, int P, int D, int M, int R, const float* bg, int W, int H,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* scales, float scale_modifier, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                 const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                 const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                 float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                 float* dL_dscale, float* dL_drot, int debug) {
    State* s = (State*)s_;
    try {
#ifdef REF_RGBD
        (void)debug;
        CudaRasterizer::Rasterizer::backward(P, D, M, R, bg, W, H, means3D, shs, colors_precomp,
            scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, radii, s->geom.p, s->bin.p, s->img.p, dL_dpix, dL_dmean2D,
            dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot);
#else
        CudaRasterizer::Rasterizer::backward(P, D, M, R, bg, W, H, means3D, shs, colors_precomp,
            scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, radii, s->geom.p, s->bin.p, s->img.p, dL_dpix, dL_dmean2D,
            dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot,
            debug != 0);
#endif
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { snprintf(g_err, sizeof g_err, "%s", cudaGetErrorString(e)); return -1; }
        return 0;
    } catch (const std::exception& ex) {
        snprintf(g_err, sizeof g_err, "%s", ex.what());
        return -1;
    }
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Copy one named field of the opaque state of the last forward() into dst (device memory,
// caller-sized).  Layout follows GeometryState/BinningState/ImageState::fromChunk.
// Returns the number of bytes copied or -1.
long long ref_state_field(void* s_, const char* name, void* dst) {
    State* s = (State*)s_;
    using namespace CudaRasterizer;
    char* gp = s->geom.p; char* bp = s->bin.p; char* ip = s->img.p;
    if (!gp || !ip) return -1;
    GeometryState g = GeometryState::fromChunk(gp, s->P);
    ImageState im = ImageState::fromChunk(ip, (size_t)s->W * s->H);
    const void* src = nullptr; size_t n = 0;
    size_t P = s->P, R = s->R, N = (size_t)s->W * s->H;
    size_t tiles = (size_t)((s->W + BLOCK_X - 1) / BLOCK_X) * ((s->H + BLOCK_Y - 1) / BLOCK_Y);
    if (!strcmp(name, "depths")) { src = g.depths; n = P * 4; }
    else if (!strcmp(name, "clamped")) { src = g.clamped; n = P * 3; }
    else if (!strcmp(name, "means2D")) { src = g.means2D; n = P * 8; }
    else if (!strcmp(name, "cov3D")) { src = g.cov3D; n = P * 24; }
    else if (!strcmp(name, "conic_opacity")) { src = g.conic_opacity; n = P * 16; }
    else if (!strcmp(name, "rgb")) { src = g.rgb; n = P * 12; }
    else if (!strcmp(name, "tiles_touched")) { src = g.tiles_touched; n = P * 4; }
    else if (!strcmp(name, "point_offsets")) { src = g.point_offsets; n = P * 4; }
    else if (!strcmp(name, "accum_alpha")) { src = im.accum_alpha; n = N * 4; }
    else if (!strcmp(name, "n_contrib")) { src = im.n_contrib; n = N * 4; }
    else if (!strcmp(name, "ranges")) { src = im.ranges; n = tiles * 8; }
    else {
        if (!bp || R == 0) return (!strcmp(name, "point_list") || !strcmp(name, "point_list_keys")) ? 0 : -1;
        BinningState b = BinningState::fromChunk(bp, R);
        if (!strcmp(name, "point_list")) { src = b.point_list; n = R * 4; }
        else if (!strcmp(name, "point_list_keys")) { src = b.point_list_keys; n = R * 8; }
        else return -1;
    }
    if (n && cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice) != cudaSuccess) return -1;
    return (long long)n;
}

}  // extern "C"
