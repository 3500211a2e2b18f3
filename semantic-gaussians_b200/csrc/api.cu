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
const char* sgb_version(void) { return "sgb200 0.2.0 (sm_100a)"; }
#ifndef SGB_BUILD_ID
#define SGB_BUILD_ID "unknown"
#endif
const char* sgb_build_id(void) { return "sgb200 0.2.0 src:" SGB_BUILD_ID; }

int sgb_ctx_create(sgb_ctx** out, int device) {
    if (!out) { set_error("null out"); return SGB_E_INVALID; }
    // the caller's current device is left as it was (a host framework tracks it; changing it behind its back makes
    // later launches land on streams of a non-current device)
    int prev = -1;
    SGB_CUDA(cudaGetDevice(&prev));
    SGB_CUDA(cudaSetDevice(device));
    sgb_ctx* c = new sgb_ctx();
    c->device = device;
    cudaError_t e = cudaMallocHost(&c->pinned, 1024);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
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
    for (PoolSlot& sl : c->pools)
        if (sl.mem.p) cudaFree(sl.mem.p);
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
                                          "fusion_sort", "fusion_gather", "alpha_pass", "dfeature"};
    return (st >= 0 && st < ST_COUNT) ? names[st] : "";
}
uint64_t sgb_ctx_launch_count(const sgb_ctx* c, int library_calls) {
    return c ? (library_calls ? c->lib_launches : c->launches) : 0;
}

size_t sgb_ctx_scratch_bytes(const sgb_ctx* c) {
    if (!c) return 0;
    size_t n = c->geom.cap + c->bin.cap + c->misc.cap;
    for (const PoolSlot& sl : c->pools) n += sl.mem.cap;
    return n;
}

size_t sgb_geometry_bytes(int32_t P) { return GeomView::carve(nullptr, P > 0 ? P : 1).bytes; }
size_t sgb_binning_bytes(int64_t R) { return BinView::carve(nullptr, R).bytes; }
size_t sgb_image_bytes(int32_t W, int32_t H) { return ImgView::carve(nullptr, W, H).bytes; }

// ---- forward / backward, single view and batched (the single-view entry points are the V = 1 case) -----------
static sgb_view_inputs with_camera(const sgb_view_inputs& in, const sgb_camera* cams, int v) {
    sgb_view_inputs o = in;
    if (cams) {
        o.viewmatrix = cams[v].viewmatrix;
        o.projmatrix = cams[v].projmatrix;
        o.campos = cams[v].campos;
        o.tan_fovx = cams[v].tan_fovx;
        o.tan_fovy = cams[v].tan_fovy;
    }
    return o;
}

static int check_batch(const sgb_view_inputs* in, int32_t V, const sgb_camera* cams, sgb_view_inputs* probe) {
    if (!in) { set_error("null sgb_view_inputs"); return SGB_E_INVALID; }
    if (V < 1 || V > SGB_MAX_BATCH) { set_error("batch of %d views: need 1 <= V <= %d", V, SGB_MAX_BATCH); return SGB_E_INVALID; }
    if (!cams) { set_error("null camera array"); return SGB_E_INVALID; }
    for (int v = 0; v < V; v++) {
        *probe = with_camera(*in, cams, v);
        int rc = check_inputs(probe);
        if (rc) return rc;
    }
    return SGB_OK;
}

static int forward_geometry_impl(sgb_ctx* ctx, const sgb_view_inputs& in, int V, const sgb_camera* cams,
                                 void* const* geometry_states, int32_t* const* radii, int64_t* num_rendered_host,
                                 cudaStream_t s) {
    for (int v = 0; v < V; v++) num_rendered_host[v] = 0;
    ctx->last_P = 0;
    ctx->last_V = 0;
    if (in.P == 0) return SGB_OK;  // rasterize_points.cu:84: nothing to do for an empty scene
    int rc = run_depth_order_and_scan(ctx, in, V, cams, geometry_states, radii, num_rendered_host, s);
    if (rc) return rc;
    for (int v = 0; v < V; v++)
        if (num_rendered_host[v] > 0x7fffffffLL) {
            set_error("num_rendered %lld exceeds int32", (long long)num_rendered_host[v]);
            return SGB_E_OVERFLOW;
        }
    return SGB_OK;
}

static int forward_render_impl(sgb_ctx* ctx, const sgb_view_inputs& in_common, int V, const sgb_camera* cams,
                               const int64_t* num_rendered, void* const* geometry_states, void* const* binning_states,
                               void* const* image_states, const int32_t* const* radii, float* const* out_colors,
                               float* const* out_depths, cudaStream_t s) {
    int64_t maxR = 0;
    for (int v = 0; v < V; v++) maxR = num_rendered[v] > maxR ? num_rendered[v] : maxR;
    int rc = reserve_binning(ctx, in_common, in_common.P > 0 ? maxR : 0, s);
    if (rc) return rc;
    struct View { sgb_view_inputs in; int64_t R; GeomView g; BinView b; ImgView im; const float* colors; };
    View vw[SGB_MAX_BATCH];
    const bool wide = in_common.C > 4;
    for (int v = 0; v < V; v++) {
        View& w = vw[v];
        w.in = with_camera(in_common, cams, v);
        w.R = w.in.P > 0 ? num_rendered[v] : 0;
        w.g = GeomView::carve(geometry_states[v], w.in.P > 0 ? w.in.P : 1);
        w.b = BinView::carve(binning_states[v], w.R);
        w.im = ImgView::carve(image_states[v], w.in.W, w.in.H);
        w.colors = w.in.colors_precomp ? w.in.colors_precomp : w.g.rgb;  // rasterizer_impl.cu:324
        rc = run_binning(ctx, w.in, v, w.R, w.g, w.b, w.im, radii[v], s);
        if (rc) return rc;
        if (wide) {
            rc = blend_forward_v3_alpha(ctx, v, w.in, w.R, w.g, w.b, w.im, s);
            if (rc) return rc;
        } else {
            StageTimer t(ctx, ST_BLEND_FWD, s);
            ctx->launches += 1;
            rc = launch_blend_forward(w.in, w.g, w.b, w.im, w.colors, out_colors[v], out_depths ? out_depths[v] : nullptr, s);
            if (rc) return rc;
        }
    }
    if (!wide) return SGB_OK;
    // ONE sync validates the weight pools of all views (the host waits for the alpha passes only); a view whose
    // pool overflowed (first view of a new scene) repeats its alpha pass.  Then the forward GEMMs are enqueued.
    bool pending[SGB_MAX_BATCH];
    for (int v = 0; v < V; v++) pending[v] = true;
    for (int attempt = 0;; attempt++) {
        SGB_CUDA(cudaStreamSynchronize(s));
        bool again = false;
        for (int v = 0; v < V; v++) {
            if (!pending[v]) continue;
            View& w = vw[v];
            const int f = blend_forward_v3_finish(ctx, v, w.in, w.R, w.b);
            if (f < 0) return f;
            if (f == 0) { pending[v] = false; continue; }
            if (attempt >= 3) { set_error("weight pool kept overflowing"); return SGB_E_NOMEM; }
            rc = blend_forward_v3_alpha(ctx, v, w.in, w.R, w.g, w.b, w.im, s);
            if (rc) return rc;
            again = true;
        }
        if (!again) break;
    }
    for (int v = 0; v < V; v++) {
        View& w = vw[v];
        rc = blend_forward_v3_gemm(ctx, w.in, w.R, w.b, w.im, w.colors, out_colors[v], s);
        if (rc) return rc;
    }
    return SGB_OK;
}

static int backward_impl(sgb_ctx* ctx, const sgb_view_inputs& in_common, int V, const sgb_camera* cams,
                         const int64_t* num_rendered, const int32_t* const* radii, const void* const* geometry_states,
                         const void* const* binning_states, const void* const* image_states,
                         const float* const* dL_dpix, const sgb_view_grads* grads, cudaStream_t s) {
    if (in_common.P == 0) {
        if (ctx->feature_grad_event) SGB_CUDA(cudaEventRecord(ctx->feature_grad_event, s));
        return SGB_OK;
    }
    struct View { sgb_view_inputs in; int64_t R; GeomView g; BinView b; ImgView im; const float* colors; };
    View vw[SGB_MAX_BATCH];
    for (int v = 0; v < V; v++) {
        const sgb_view_grads& gr = grads[v];
        if (!geometry_states[v] || !radii[v] || !image_states[v] || !dL_dpix[v] || !gr.dL_dmeans2D || !gr.dL_dconic ||
            !gr.dL_dopacity || !gr.dL_dcolors || !gr.dL_dmeans3D || !gr.dL_dcov3D || (in_common.shs && !gr.dL_dsh) ||
            (in_common.scales && (!gr.dL_dscales || !gr.dL_drotations))) {
            set_error("sgb_backward: null state or gradient buffer");
            return SGB_E_INVALID;
        }
        vw[v].in = with_camera(in_common, cams, v);
        vw[v].R = num_rendered[v];
        vw[v].g = GeomView::carve(const_cast<void*>(geometry_states[v]), in_common.P);
        vw[v].b = BinView::carve(const_cast<void*>(binning_states[v]), num_rendered[v]);
        vw[v].im = ImgView::carve(const_cast<void*>(image_states[v]), in_common.W, in_common.H);
        vw[v].colors = in_common.colors_precomp ? in_common.colors_precomp : vw[v].g.rgb;  // rasterizer_impl.cu:394
    }
    const bool wide = in_common.C > 4;
    int rc;
    // dL/dfeature of every view first: it needs only the weight rows and dL/dout, it is the one large gradient and
    // it accumulates across the views of a batch, so a data-parallel caller can start exchanging it while the
    // chain / geometry kernels of the whole batch run (sgb200.h)
    if (wide)
        for (int v = 0; v < V; v++) {
            if (vw[v].R <= 0) continue;
            rc = blend_backward_v3_dfeature(ctx, vw[v].in, vw[v].R, vw[v].g, vw[v].b, vw[v].im, dL_dpix[v],
                                            grads[v].dL_dcolors, s);
            if (rc) return rc;
        }
    if (wide && ctx->feature_grad_event) SGB_CUDA(cudaEventRecord(ctx->feature_grad_event, s));
    for (int v = 0; v < V; v++) {
        const sgb_view_grads& gr = grads[v];
        if (vw[v].R > 0 && wide) {
            rc = blend_backward_v3_chain(ctx, vw[v].in, vw[v].R, vw[v].g, vw[v].b, vw[v].im, vw[v].colors, dL_dpix[v],
                                         gr.dL_dmeans2D, gr.dL_dconic, gr.dL_dopacity, s);
            if (rc) return rc;
        } else if (vw[v].R > 0) {
            StageTimer t(ctx, ST_BLEND_BWD, s);
            ctx->launches += 1;
            rc = launch_blend_backward(vw[v].in, vw[v].g, vw[v].b, vw[v].im, vw[v].colors, dL_dpix[v], gr.dL_dmeans2D,
                                       gr.dL_dconic, gr.dL_dopacity, gr.dL_dcolors, s);
            if (rc) return rc;
        }
        if (!wide && v == V - 1 && ctx->feature_grad_event) SGB_CUDA(cudaEventRecord(ctx->feature_grad_event, s));
        const float* cov3D = in_common.cov3D_precomp ? in_common.cov3D_precomp : vw[v].g.cov3D;  // rasterizer_impl.cu:417
        StageTimer t(ctx, ST_GEOM_BWD, s);
        ctx->launches += 1;
        rc = launch_geom_backward(vw[v].in, vw[v].g, radii[v], cov3D, gr.dL_dcolors, gr, s);
        if (rc) return rc;
    }
    return SGB_OK;
}

int sgb_forward_geometry(sgb_ctx* ctx, const sgb_view_inputs* in, void* geometry_state, int32_t* radii,
                         int64_t* num_rendered_host, void* stream) {
    int rc = check_inputs(in);
    if (rc) return rc;
    if (!ctx || !num_rendered_host || (in->P > 0 && (!geometry_state || !radii))) {
        set_error("sgb_forward_geometry: null ctx/state/radii/num_rendered");
        return SGB_E_INVALID;
    }
    void* gs[1] = {geometry_state};
    int32_t* rd[1] = {radii};
    return forward_geometry_impl(ctx, *in, 1, nullptr, gs, rd, num_rendered_host, (cudaStream_t)stream);
}

int sgb_forward_geometry_batch(sgb_ctx* ctx, const sgb_view_inputs* in, int32_t V, const sgb_camera* cams,
                               void* const* geometry_states, int32_t* const* radii, int64_t* num_rendered_host,
                               void* stream) {
    sgb_view_inputs probe;
    int rc = check_batch(in, V, cams, &probe);
    if (rc) return rc;
    if (!ctx || !num_rendered_host || !geometry_states || !radii) {
        set_error("sgb_forward_geometry_batch: null ctx/state/radii/num_rendered");
        return SGB_E_INVALID;
    }
    if (in->P > 0)
        for (int v = 0; v < V; v++)
            if (!geometry_states[v] || !radii[v]) { set_error("sgb_forward_geometry_batch: null state of view %d", v); return SGB_E_INVALID; }
    return forward_geometry_impl(ctx, *in, V, cams, geometry_states, radii, num_rendered_host, (cudaStream_t)stream);
}

int sgb_forward_render(sgb_ctx* ctx, const sgb_view_inputs* in, int64_t num_rendered, void* geometry_state,
                       void* binning_state, void* image_state, const int32_t* radii, float* out_color,
                       float* out_depth, void* stream) {
    int rc = check_inputs(in);
    if (rc) return rc;
    if (!ctx || !image_state || !out_color || (in->P > 0 && (!geometry_state || !radii)) ||
        (num_rendered > 0 && !binning_state)) {
        set_error("sgb_forward_render: null ctx/state/output");
        return SGB_E_INVALID;
    }
    if (in->C > 4 && out_depth) {
        set_error("out_depth is only produced by the 3-channel RGB-D path (C <= 4)");
        return SGB_E_INVALID;
    }
    void* gs[1] = {geometry_state};
    void* bs[1] = {binning_state};
    void* is[1] = {image_state};
    const int32_t* rd[1] = {radii};
    float* oc[1] = {out_color};
    float* od[1] = {out_depth};
    return forward_render_impl(ctx, *in, 1, nullptr, &num_rendered, gs, bs, is, rd, oc, od, (cudaStream_t)stream);
}

int sgb_forward_render_batch(sgb_ctx* ctx, const sgb_view_inputs* in, int32_t V, const sgb_camera* cams,
                             const int64_t* num_rendered, void* const* geometry_states, void* const* binning_states,
                             void* const* image_states, const int32_t* const* radii, float* const* out_colors,
                             float* const* out_depths, void* stream) {
    sgb_view_inputs probe;
    int rc = check_batch(in, V, cams, &probe);
    if (rc) return rc;
    if (!ctx || !num_rendered || !geometry_states || !binning_states || !image_states || !radii || !out_colors) {
        set_error("sgb_forward_render_batch: null argument");
        return SGB_E_INVALID;
    }
    if (in->C > 4 && out_depths) {
        set_error("out_depth is only produced by the 3-channel RGB-D path (C <= 4)");
        return SGB_E_INVALID;
    }
    for (int v = 0; v < V; v++)
        if (!image_states[v] || !out_colors[v] || (in->P > 0 && (!geometry_states[v] || !radii[v])) ||
            (num_rendered[v] > 0 && !binning_states[v])) {
            set_error("sgb_forward_render_batch: null state/output of view %d", v);
            return SGB_E_INVALID;
        }
    return forward_render_impl(ctx, *in, V, cams, num_rendered, geometry_states, binning_states, image_states, radii,
                               out_colors, out_depths, (cudaStream_t)stream);
}

int sgb_backward(sgb_ctx* ctx, const sgb_view_inputs* in, int64_t num_rendered, const int32_t* radii,
                 const void* geometry_state, const void* binning_state, const void* image_state,
                 const float* dL_dpix, const sgb_view_grads* gr, void* stream) {
    int rc = check_inputs(in);
    if (rc) return rc;
    if (!ctx || !gr || !dL_dpix || !image_state) { set_error("sgb_backward: null argument"); return SGB_E_INVALID; }
    const int32_t* rd[1] = {radii};
    const void* gs[1] = {geometry_state};
    const void* bs[1] = {binning_state};
    const void* is[1] = {image_state};
    const float* dl[1] = {dL_dpix};
    return backward_impl(ctx, *in, 1, nullptr, &num_rendered, rd, gs, bs, is, dl, gr, (cudaStream_t)stream);
}

int sgb_backward_batch(sgb_ctx* ctx, const sgb_view_inputs* in, int32_t V, const sgb_camera* cams,
                       const int64_t* num_rendered, const int32_t* const* radii, const void* const* geometry_states,
                       const void* const* binning_states, const void* const* image_states,
                       const float* const* dL_dpix, const sgb_view_grads* grads, void* stream) {
    sgb_view_inputs probe;
    int rc = check_batch(in, V, cams, &probe);
    if (rc) return rc;
    if (!ctx || !num_rendered || !radii || !geometry_states || !binning_states || !image_states || !dL_dpix || !grads) {
        set_error("sgb_backward_batch: null argument");
        return SGB_E_INVALID;
    }
    return backward_impl(ctx, *in, V, cams, num_rendered, radii, geometry_states, binning_states, image_states, dL_dpix,
                         grads, (cudaStream_t)stream);
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
