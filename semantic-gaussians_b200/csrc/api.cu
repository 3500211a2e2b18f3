// C-ABI entry points (include/sgb200.h): argument validation, state carving, stage sequencing.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace sgb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return SGB_E_CUDA;
}

namespace {

// Same argument rules as GaussianRasterizer.forward (channel_rasterization/__init__.py:258-264)
// and Rasterizer::forward (rasterizer_impl.cu:243-246).
int check_inputs(const sgb_view_inputs* in) {
    if (!in) { set_error("null sgb_view_inputs"); return SGB_E_INVALID; }
    if (in->P < 0 || in->W <= 0 || in->H <= 0 || in->C <= 0) {
        set_error("invalid sizes P=%d W=%d H=%d C=%d", in->P, in->W, in->H, in->C);
        return SGB_E_INVALID;
    }
    if ((in->shs == nullptr) == (in->colors_precomp == nullptr)) {
        set_error("Please provide excatly one of either SHs or precomputed colors!");
        return SGB_E_INVALID;
    }
    const bool sr = in->scales != nullptr && in->rotations != nullptr;
    const bool any_sr = in->scales != nullptr || in->rotations != nullptr;
    if ((!sr && in->cov3D_precomp == nullptr) || (any_sr && in->cov3D_precomp != nullptr)) {
        set_error("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        return SGB_E_INVALID;
    }
    if (in->C != 3 && in->colors_precomp == nullptr) {
        set_error("For non-RGB, provide precomputed Gaussian colors!");
        return SGB_E_INVALID;
    }
    if (in->shs && (in->M <= 0 || in->M > SGB_MAX_SH_COEFFS || in->D < 0 || (in->D + 1) * (in->D + 1) > in->M)) {
        set_error("SH degree %d needs %d coefficients, got M=%d", in->D, (in->D + 1) * (in->D + 1), in->M);
        return SGB_E_INVALID;
    }
    if (!in->background || !in->means3D || !in->opacities || !in->viewmatrix || !in->projmatrix || !in->campos) {
        set_error("null required input pointer");
        return SGB_E_INVALID;
    }
    return SGB_OK;
}

__global__ void extract_rec_kernel(int P, const SplatRec* __restrict__ rec, int what, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const SplatRec r = rec[i];
    if (what == 0) dst[i] = r.depth;
    else if (what == 1) { dst[2 * i] = r.mx; dst[2 * i + 1] = r.my; }
    else { dst[4 * i] = r.cx; dst[4 * i + 1] = r.cy; dst[4 * i + 2] = r.cz; dst[4 * i + 3] = r.op; }
}

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" {

const char* sgb_last_error(void) { return g_err; }
const char* sgb_version(void) { return "sgb200 0.1.0 (sm_100a)"; }

int sgb_ctx_create(sgb_ctx** out, int device) {
    if (!out) { set_error("null out"); return SGB_E_INVALID; }
    SGB_CUDA(cudaSetDevice(device));
    sgb_ctx* c = new sgb_ctx();
    c->device = device;
    cudaError_t e = cudaMallocHost(&c->pinned, 64);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaMallocHost"); }
    *out = c;
    return SGB_OK;
}

void sgb_ctx_destroy(sgb_ctx* c) {
    if (!c) return;
    if (c->prof.created)
        for (int st = 0; st < ST_COUNT; st++)
            for (int i = 0; i < kProfRing; i++) {
                cudaEventDestroy(c->prof.ev[st][i][0]);
                cudaEventDestroy(c->prof.ev[st][i][1]);
            }
    if (c->geom.p) cudaFree(c->geom.p);
    if (c->bin.p) cudaFree(c->bin.p);
    if (c->misc.p) cudaFree(c->misc.p);
    if (c->pool.p) cudaFree(c->pool.p);
    if (c->pinned) cudaFreeHost(c->pinned);
    delete c;
}

int sgb_profile_enable(sgb_ctx* c, int on) {
    if (!c) { set_error("null ctx"); return SGB_E_INVALID; }
    if (on && !c->prof.created) {
        for (int st = 0; st < ST_COUNT; st++)
            for (int i = 0; i < kProfRing; i++) {
                SGB_CUDA(cudaEventCreate(&c->prof.ev[st][i][0]));
                SGB_CUDA(cudaEventCreate(&c->prof.ev[st][i][1]));
            }
        c->prof.created = true;
    }
    for (int st = 0; st < ST_COUNT; st++) c->prof.n[st] = 0;
    c->prof.on = on != 0;
    return SGB_OK;
}

int sgb_profile_read(sgb_ctx* c, float* ms_sum, int32_t* count) {
    if (!c || !ms_sum || !count) { set_error("null argument"); return SGB_E_INVALID; }
    for (int st = 0; st < ST_COUNT; st++) {
        float acc = 0.f;
        for (int i = 0; i < c->prof.n[st]; i++) {
            SGB_CUDA(cudaEventSynchronize(c->prof.ev[st][i][1]));
            float ms = 0.f;
            SGB_CUDA(cudaEventElapsedTime(&ms, c->prof.ev[st][i][0], c->prof.ev[st][i][1]));
            acc += ms;
        }
        ms_sum[st] = acc;
        count[st] = c->prof.n[st];
        c->prof.n[st] = 0;
    }
    return SGB_OK;
}

int sgb_profile_num_stages(void) { return ST_COUNT; }
const char* sgb_profile_stage_name(int st) {
    static const char* names[ST_COUNT] = {"preprocess", "depth_sort", "scan", "emit", "tile_sort", "ranges",
                                          "blend_fwd", "blend_bwd", "geom_bwd", "fusion_project",
                                          "fusion_transpose", "fusion_gather", "alpha_pass", "dfeature"};
    return (st >= 0 && st < ST_COUNT) ? names[st] : "";
}
uint64_t sgb_ctx_launch_count(const sgb_ctx* c, int library_calls) {
    return c ? (library_calls ? c->lib_launches : c->launches) : 0;
}

size_t sgb_ctx_scratch_bytes(const sgb_ctx* c) { return c ? c->geom.cap + c->bin.cap + c->misc.cap + c->pool.cap : 0; }

size_t sgb_geometry_bytes(int32_t P) { return GeomView::carve(nullptr, P > 0 ? P : 1).bytes; }
size_t sgb_binning_bytes(int64_t R) { return BinView::carve(nullptr, R).bytes; }
size_t sgb_image_bytes(int32_t W, int32_t H) { return ImgView::carve(nullptr, W, H).bytes; }

int sgb_forward_geometry(sgb_ctx* ctx, const sgb_view_inputs* in, void* geometry_state, int32_t* radii,
                         int64_t* num_rendered_host, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_inputs(in);
    if (rc) return rc;
    if (!ctx || !num_rendered_host || (in->P > 0 && (!geometry_state || !radii))) {
        set_error("sgb_forward_geometry: null ctx/state/radii/num_rendered");
        return SGB_E_INVALID;
    }
    *num_rendered_host = 0;
    ctx->last_P = 0;
    ctx->pool_valid = false;  // a new view starts: cached weight rows belong to the previous one
    if (in->P == 0) return SGB_OK;  // rasterize_points.cu:84: nothing to do for an empty scene
    GeomView g = GeomView::carve(geometry_state, in->P);
    rc = run_depth_order_and_scan(ctx, *in, g, radii, num_rendered_host, s);
    if (rc) return rc;
    if (*num_rendered_host > 0x7fffffffLL) {
        set_error("num_rendered %lld exceeds int32", (long long)*num_rendered_host);
        return SGB_E_OVERFLOW;
    }
    return SGB_OK;
}

int sgb_forward_render(sgb_ctx* ctx, const sgb_view_inputs* in, int64_t num_rendered, void* geometry_state,
                       void* binning_state, void* image_state, const int32_t* radii, float* out_color,
                       float* out_depth, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_inputs(in);
    if (rc) return rc;
    if (!ctx || !image_state || !out_color || (in->P > 0 && (!geometry_state || !radii)) ||
        (num_rendered > 0 && !binning_state)) {
        set_error("sgb_forward_render: null ctx/state/output");
        return SGB_E_INVALID;
    }
    GeomView g = GeomView::carve(geometry_state, in->P > 0 ? in->P : 1);
    BinView b = BinView::carve(binning_state, num_rendered);
    ImgView im = ImgView::carve(image_state, in->W, in->H);
    if (in->P == 0) num_rendered = 0;
    rc = run_binning(ctx, *in, num_rendered, g, b, im, radii, s);
    if (rc) return rc;
    const float* colors = in->colors_precomp ? in->colors_precomp : g.rgb;  // rasterizer_impl.cu:324
    if (in->C > 4) {
        if (out_depth) {
            set_error("out_depth is only produced by the 3-channel RGB-D path (C <= 4)");
            return SGB_E_INVALID;
        }
        return blend_forward_v3(ctx, *in, num_rendered, g, b, im, colors, out_color, s);
    }
    StageTimer t(ctx, ST_BLEND_FWD, s);
    ctx->launches += 1;
    return launch_blend_forward(*in, g, b, im, colors, out_color, out_depth, s);
}

int sgb_backward(sgb_ctx* ctx, const sgb_view_inputs* in, int64_t num_rendered, const int32_t* radii,
                 const void* geometry_state, const void* binning_state, const void* image_state,
                 const float* dL_dpix, const sgb_view_grads* gr, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = check_inputs(in);
    if (rc) return rc;
    if (!ctx || !gr || !dL_dpix || !image_state) { set_error("sgb_backward: null argument"); return SGB_E_INVALID; }
    if (in->P == 0) {
        if (ctx->feature_grad_event) SGB_CUDA(cudaEventRecord(ctx->feature_grad_event, s));
        return SGB_OK;
    }
    if (!geometry_state || !radii || !gr->dL_dmeans2D || !gr->dL_dconic || !gr->dL_dopacity || !gr->dL_dcolors ||
        !gr->dL_dmeans3D || !gr->dL_dcov3D || (in->shs && !gr->dL_dsh) || (in->scales && (!gr->dL_dscales || !gr->dL_drotations))) {
        set_error("sgb_backward: null state or gradient buffer");
        return SGB_E_INVALID;
    }
    GeomView g = GeomView::carve(const_cast<void*>(geometry_state), in->P);
    BinView b = BinView::carve(const_cast<void*>(binning_state), num_rendered);
    ImgView im = ImgView::carve(const_cast<void*>(image_state), in->W, in->H);
    const float* colors = in->colors_precomp ? in->colors_precomp : g.rgb;  // rasterizer_impl.cu:394
    if (num_rendered > 0 && in->C > 4) {
        rc = blend_backward_v3(ctx, *in, num_rendered, g, b, im, colors, dL_dpix, gr->dL_dmeans2D, gr->dL_dconic,
                               gr->dL_dopacity, gr->dL_dcolors, s);
        if (rc) return rc;
    } else if (num_rendered > 0) {
        StageTimer t(ctx, ST_BLEND_BWD, s);
        ctx->launches += 1;
        rc = launch_blend_backward(*in, g, b, im, colors, dL_dpix, gr->dL_dmeans2D, gr->dL_dconic, gr->dL_dopacity,
                                   gr->dL_dcolors, s);
        if (rc) return rc;
    }
    // C > 4 records the event inside blend_backward_v3, right after the dL/dfeature kernel
    if (ctx->feature_grad_event && !(num_rendered > 0 && in->C > 4))
        SGB_CUDA(cudaEventRecord(ctx->feature_grad_event, s));
    const float* cov3D = in->cov3D_precomp ? in->cov3D_precomp : g.cov3D;  // rasterizer_impl.cu:417
    StageTimer t(ctx, ST_GEOM_BWD, s);
    ctx->launches += 1;
    return launch_geom_backward(*in, g, radii, cov3D, gr->dL_dcolors, *gr, s);
}

int64_t sgb_ctx_view_stat(const sgb_ctx* ctx, int which) {
    if (!ctx) return -1;
    return which == 0 ? ctx->stat_blended_pairs : which == 1 ? ctx->stat_pool_chunks : -1;
}

int sgb_ctx_set_feature_grad_event(sgb_ctx* ctx, void* cuda_event) {
    if (!ctx) { set_error("sgb_ctx_set_feature_grad_event: null ctx"); return SGB_E_INVALID; }
    ctx->feature_grad_event = (cudaEvent_t)cuda_event;
    return SGB_OK;
}

int sgb_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
    (void)projmatrix;  // the reference computes p_proj but only tests view-space z (auxiliary.h:149-154)
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
        set_error("sgb_mark_visible: bad arguments");
        return SGB_E_INVALID;
    }
    if (P == 0) return SGB_OK;
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

int64_t sgb_state_field(const char* name, int32_t P, int64_t R, int32_t W, int32_t H, const void* geometry_state,
                        const void* binning_state, const void* image_state, void* dst, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!name || !dst) { set_error("sgb_state_field: null"); return SGB_E_INVALID; }
    const size_t N = (size_t)W * H;
    const size_t tiles = (size_t)((W + SGB_TILE - 1) / SGB_TILE) * ((H + SGB_TILE - 1) / SGB_TILE);
    const void* src = nullptr;
    size_t n = 0;
    if (geometry_state && P > 0) {
        GeomView g = GeomView::carve(const_cast<void*>(geometry_state), P);
        int what = !strcmp(name, "depths") ? 0 : !strcmp(name, "means2D") ? 1 : !strcmp(name, "conic_opacity") ? 2 : -1;
        if (what >= 0) {
            extract_rec_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.rec, what, (float*)dst);
            SGB_LAUNCH_CHECK("extract_rec_kernel", 0, s);
            return (int64_t)P * (what == 0 ? 4 : what == 1 ? 8 : 16);
        }
        if (!strcmp(name, "cov3D")) { src = g.cov3D; n = (size_t)P * 24; }
        else if (!strcmp(name, "rgb")) { src = g.rgb; n = (size_t)P * 12; }
        else if (!strcmp(name, "clamped")) { src = g.clamped; n = (size_t)P * 3; }
        else if (!strcmp(name, "tiles_touched")) { src = g.tiles_touched; n = (size_t)P * 4; }
    }
    if (!src && image_state) {
        ImgView im = ImgView::carve(const_cast<void*>(image_state), W, H);
        if (!strcmp(name, "final_T")) { src = im.final_T; n = N * 4; }
        else if (!strcmp(name, "n_contrib")) { src = im.n_contrib; n = N * 4; }
        else if (!strcmp(name, "ranges")) { src = im.ranges; n = tiles * 8; }
        else if (!strcmp(name, "tile_last")) { src = im.tile_last; n = tiles * 4; }
    }
    if (!src && !strcmp(name, "point_list")) {
        if (R == 0) return 0;
        if (!binning_state) { set_error("no binning state"); return SGB_E_INVALID; }
        src = BinView::carve(const_cast<void*>(binning_state), R).point_list;
        n = (size_t)R * 4;
    }
    if (!src) { set_error("unknown or unavailable state field '%s'", name); return SGB_E_INVALID; }
    SGB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s));
    return (int64_t)n;
}

}  // extern "C"
