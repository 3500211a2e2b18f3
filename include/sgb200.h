/*
 * sgb200 — B200-native Gaussian-splatting rasterizer + 2D->3D fusion kernels: C ABI.
 *
 * This is the drop-in boundary for the reference's native rasterizer library.  Each entry
 * point names the reference interface it replaces (paths relative to
 * /root/reference/submodules/channel-rasterization unless prefixed):
 *
 *   sgb_forward_geometry + sgb_forward_render
 *        == CudaRasterizer::Rasterizer::forward           cuda_rasterizer/rasterizer.h:30-53,
 *           rgbd variant (out_depth)                      rgbd/cuda_rasterizer/rasterizer.h:31-53
 *   sgb_backward
 *        == CudaRasterizer::Rasterizer::backward          cuda_rasterizer/rasterizer.h:55-84
 *   sgb_mark_visible
 *        == CudaRasterizer::Rasterizer::markVisible       cuda_rasterizer/rasterizer.h:23-28
 *   sgb_fusion_map / sgb_fusion_accumulate / sgb_fusion_normalize
 *        == PointCloudToImageMapper.compute_mapping       dataset/fusion_utils.py:30-78
 *           + the per-view gather/accumulate and final divide   fusion.py:127-148
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.  Every array argument is a
 *     DEVICE pointer unless its name ends in _host.  An absent optional input is NULL
 *     (the reference encodes it the same way: rasterizer_impl.cu:243,324,394,417).
 *   - Every call takes the CUDA stream to launch on (cudaStream_t passed as void*); the
 *     reference launches on the legacy default stream.  Calls are asynchronous except
 *     sgb_forward_geometry, which returns num_rendered and therefore synchronises the stream
 *     once — the reference does the same with a blocking cudaMemcpy (rasterizer_impl.cu:283).
 *   - Return value: SGB_OK (0) or a negative error code; sgb_last_error() gives the message of
 *     the last failure on the calling thread.  (Reference: C++ exceptions,
 *     rasterizer_impl.cu:245, auxiliary.h:166-173.)
 *   - Ownership: the caller owns every buffer.  Forward fills three opaque state buffers
 *     (geometry / binning / image) that backward consumes — same contract as the reference's
 *     geomBuffer/binningBuffer/imgBuffer (rasterize_points.cu:28-36,73-80), but sized up
 *     front through sgb_*_bytes() instead of std::function resize callbacks.  Scratch that does
 *     not outlive a call (sort double-buffers, CUB temp storage) lives in the sgb_ctx.
 *   - A ctx is bound to one device and must not be used by two streams concurrently.
 */
#ifndef SGB200_H_INCLUDED
#define SGB200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGB_OK 0
#define SGB_E_INVALID (-1)   /* bad argument combination (reference: "provide exactly one of ...") */
#define SGB_E_CUDA (-2)      /* CUDA runtime / launch error */
#define SGB_E_NOMEM (-3)     /* scratch allocation failed */
#define SGB_E_OVERFLOW (-4)  /* num_rendered does not fit in int32 */

#define SGB_TILE 16          /* BLOCK_X = BLOCK_Y = 16, cuda_rasterizer/config.h:16-17 */
#define SGB_MAX_SH_COEFFS 16

typedef struct sgb_ctx sgb_ctx;

/* Inputs of one view.  Field-for-field the argument list of Rasterizer::forward. */
typedef struct sgb_view_inputs {
    int32_t P;                   /* number of Gaussians */
    int32_t D;                   /* active SH degree (0..3) */
    int32_t M;                   /* SH coefficients per colour channel in `shs` (0 if shs NULL) */
    int32_t W, H;                /* image size in pixels */
    int32_t C;                   /* channels: 3 for RGB; run-time C for the feature raster */
    const float* background;     /* [C] */
    const float* means3D;        /* [P,3] */
    const float* shs;            /* [P,M,3] or NULL (requires C == 3) */
    const float* colors_precomp; /* [P,C] or NULL */
    const float* opacities;      /* [P] */
    const float* scales;         /* [P,3] or NULL */
    float scale_modifier;
    const float* rotations;      /* [P,4] (w,x,y,z), used as given, or NULL */
    const float* cov3D_precomp;  /* [P,6] or NULL */
    const float* viewmatrix;     /* [16]  W2C transposed (scene/camera.py:87) */
    const float* projmatrix;     /* [16]  (P*W2C) transposed (scene/camera.py:91-93) */
    const float* campos;         /* [3] */
    float tan_fovx, tan_fovy;
    int32_t prefiltered;         /* trap if a Gaussian fails the near cull (auxiliary.h:156-160) */
    int32_t debug;               /* synchronise + check after every stage (auxiliary.h:166-173) */
} sgb_view_inputs;

/* Gradients of one view (all caller-zero-filled, as rasterize_points.cu:157-165 does). */
typedef struct sgb_view_grads {
    float* dL_dmeans2D;   /* [P,3]   .xy in the reference's units (backward.cu:455-456,540-541) */
    float* dL_dconic;     /* [P,4]   (x, y, _, w) as backward.cu:544-546 */
    float* dL_dopacity;   /* [P] */
    float* dL_dcolors;    /* [P,C]   any C (the reference ships C==3 only) */
    float* dL_dmeans3D;   /* [P,3] */
    float* dL_dcov3D;     /* [P,6] */
    float* dL_dsh;        /* [P,M,3] or NULL when shs is NULL */
    float* dL_dscales;    /* [P,3]  or NULL when scales is NULL */
    float* dL_drotations; /* [P,4]  or NULL when rotations is NULL */
} sgb_view_grads;

const char* sgb_last_error(void);
const char* sgb_version(void);

int sgb_ctx_create(sgb_ctx** out, int device);
void sgb_ctx_destroy(sgb_ctx* ctx);
/* Bytes of device scratch the ctx currently holds (diagnostics). */
size_t sgb_ctx_scratch_bytes(const sgb_ctx* ctx);

/* Stage tracing: when enabled, every stage of the calls made through this ctx is bracketed by a
 * CUDA event pair on the caller's stream.  sgb_profile_read synchronises those events and returns,
 * per stage, the summed milliseconds and the number of intervals since the previous read
 * (arrays of sgb_profile_num_stages() entries).  At most 128 intervals per stage are kept. */
int sgb_profile_enable(sgb_ctx* ctx, int on);
int sgb_profile_read(sgb_ctx* ctx, float* ms_sum, int32_t* count);
int sgb_profile_num_stages(void);
const char* sgb_profile_stage_name(int stage);
/* Kernels of this library launched through ctx so far (library_calls = 0), or the number of
 * CUB device-wide calls (library_calls = 1; each is several kernels). */
uint64_t sgb_ctx_launch_count(const sgb_ctx* ctx, int library_calls);
/* Statistics of the most recent C > 4 view rendered through ctx (no synchronisation: the values were read
 * back with the weight-pool header).  which = 0: blended (pixel, Gaussian) pairs, i.e. n-bar * W * H, the
 * quantity the blend's algorithmic flops scale with; 1: 16-entry weight-row chunks in use.  -1 if unknown. */
int64_t sgb_ctx_view_stat(const sgb_ctx* ctx, int which);
/* Optional cudaEvent_t that sgb_backward records on its stream as soon as dL_dcolors (the (P, C)
 * feature gradient — the only large per-Gaussian gradient) is final, i.e. BEFORE the chain and
 * geometry gradient kernels.  A data-parallel caller lets its communication stream wait on this
 * event so that the all-reduce of the feature gradient overlaps the rest of the backward pass
 * (the reference is single-GPU and has no counterpart).  NULL (default) disables it. */
int sgb_ctx_set_feature_grad_event(sgb_ctx* ctx, void* cuda_event);

/* Sizes of the three caller-owned state buffers (multiples of 256 B). */
size_t sgb_geometry_bytes(int32_t P);
size_t sgb_binning_bytes(int64_t num_rendered);
size_t sgb_image_bytes(int32_t W, int32_t H);

/*
 * Stage 1 of forward: per-Gaussian projection / covariance / conic / radius / tile rect / SH
 * colour (forward.cu:155-256), depth ordering and the scan over tiles_touched
 * (rasterizer_impl.cu:279).  Writes `radii` [P] int32 and *num_rendered_host (instances R).
 */
int sgb_forward_geometry(sgb_ctx* ctx, const sgb_view_inputs* in, void* geometry_state,
                         int32_t* radii, int64_t* num_rendered_host, void* stream);

/*
 * Stage 2 of forward: instance emission, tile sort, tile ranges (rasterizer_impl.cu:291-321) and
 * the per-tile front-to-back blend (forward.cu:262-375; rgbd/forward.cu:261-393 when out_depth
 * is non-NULL).  out_color [C,H,W]; out_depth [1,H,W] or NULL.  Every pixel is written, so the
 * outputs need no zero-fill.
 */
int sgb_forward_render(sgb_ctx* ctx, const sgb_view_inputs* in, int64_t num_rendered,
                       void* geometry_state, void* binning_state, void* image_state,
                       const int32_t* radii, float* out_color, float* out_depth, void* stream);

/* backward.cu:394-552 (blend), :141-271 (cov2D), :341-391 (projection / SH / scale+rot). */
int sgb_backward(sgb_ctx* ctx, const sgb_view_inputs* in, int64_t num_rendered,
                 const int32_t* radii, const void* geometry_state, const void* binning_state,
                 const void* image_state, const float* dL_dpix /* [C,H,W] */,
                 const sgb_view_grads* grads, void* stream);

/* ---- batched views (SURVEY.md §8 row n2 / BASELINE config K4): V views of the SAME Gaussians in one call.
 *
 * The reference renders one view per call in a Python loop (eval_segmentation.py:146-157, fusion.py:58-64,106-144);
 * a view-sharded training / evaluation step renders a batch of views per GPU (K4: 32 views over 8 GPUs = 4 each).
 * `in` carries everything the views share (sizes, Gaussian arrays, background, flags); its camera fields are
 * ignored and replaced by cams[v].  Per view the results are IDENTICAL to V single-view calls; what the batch
 * removes is per-view overhead:
 *   - sgb_forward_geometry_batch enqueues the V projection / depth-order / scan sequences back to back and
 *     synchronises the stream ONCE for all V instance counts (the single-view call synchronises per view);
 *   - sgb_forward_render_batch enqueues binning, alpha pass and blend of all V views and synchronises ONCE for
 *     the V weight-pool checks; every view keeps its own weight-pool slot inside the ctx (V <= SGB_MAX_BATCH),
 *     so sgb_backward_batch reuses the rows of all V forwards;
 *   - sgb_backward_batch accumulates: grads[v].dL_dcolors may be THE SAME (P, C) buffer for every view — the
 *     per-Gaussian feature gradient is summed over the local views in place (one zero-fill, no V x (P, C)
 *     temporaries, no V-way add), which is what a data-parallel step exchanges.  All V dL/dfeature kernels run
 *     first, then the feature-gradient event (sgb_ctx_set_feature_grad_event) is recorded, then the V chain /
 *     geometry kernels: the exchange of the big gradient overlaps the rest of the whole batch.  The small
 *     per-Gaussian gradients (means2D, conic, opacity, means3D, cov3D, scales, rotations, sh) are per view
 *     (distinct buffers per grads[v]; viewspace gradients feed per-view densification statistics).
 */
#define SGB_MAX_BATCH 8

typedef struct sgb_camera {
    const float* viewmatrix;     /* [16] device */
    const float* projmatrix;     /* [16] device */
    const float* campos;         /* [3]  device */
    float tan_fovx, tan_fovy;
} sgb_camera;

int sgb_forward_geometry_batch(sgb_ctx* ctx, const sgb_view_inputs* in, int32_t V, const sgb_camera* cams,
                               void* const* geometry_states, int32_t* const* radii,
                               int64_t* num_rendered_host /* [V] */, void* stream);
int sgb_forward_render_batch(sgb_ctx* ctx, const sgb_view_inputs* in, int32_t V, const sgb_camera* cams,
                             const int64_t* num_rendered /* [V] host */, void* const* geometry_states,
                             void* const* binning_states, void* const* image_states,
                             const int32_t* const* radii, float* const* out_colors,
                             float* const* out_depths /* NULL or [V] (C <= 4 only) */, void* stream);
int sgb_backward_batch(sgb_ctx* ctx, const sgb_view_inputs* in, int32_t V, const sgb_camera* cams,
                       const int64_t* num_rendered, const int32_t* const* radii,
                       const void* const* geometry_states, const void* const* binning_states,
                       const void* const* image_states, const float* const* dL_dpix,
                       const sgb_view_grads* grads /* [V] */, void* stream);

/* Identity of the build: "<version> src:<sha256 prefix of csrc/ + include/>" (set by build.py; bench.py prints
 * it so that a stale prebuilt library cannot be mistaken for the sources next to it). */
const char* sgb_build_id(void);

/* rasterizer_impl.cu:54-66,141-153: present[i] = (view-space z > 0.2). present is uint8 [P]. */
int sgb_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);

/* Named read-only views into the opaque state, for tests and diagnostics.  Copies the field
 * into dst (device) and returns the byte count, or a negative error.  Fields: "depths",
 * "means2D", "conic_opacity", "cov3D", "rgb", "clamped", "tiles_touched", "point_list",
 * "ranges", "n_contrib", "final_T". */
int64_t sgb_state_field(const char* name, int32_t P, int64_t num_rendered, int32_t W, int32_t H,
                        const void* geometry_state, const void* binning_state,
                        const void* image_state, void* dst, void* stream);

/* ------------------------------------------------------------------ fusion (2D -> 3D) ---- */

#define SGB_DEPTH_NONE 0     /* fusion.depth == None   : keep points with z > 0          */
#define SGB_DEPTH_F32 1      /* depth map float32 [h,w] ("render", fusion.py:106-120)     */
#define SGB_DEPTH_F64 2      /* depth map float64 [h,w] ("image": png / depth_scale)      */
#define SGB_DEPTH_SURFACE 3  /* z-buffer of the points themselves (fusion_utils.py:57-61) */

#define SGB_FEAT_F16 0
#define SGB_FEAT_F32 1

typedef struct sgb_fusion_view {
    int32_t P;
    const float* xyz;            /* [P,3] */
    const float* world_to_camera;/* [16] = view.world_view_transform (W2C transposed), f32 */
    double fx, fy, cx, cy;       /* intrinsics AFTER the img_dim rescale of fusion_utils.py:23-28 */
    int32_t w, h;                /* image_dim = [w, h] */
    int32_t cut_bound;
    double vis_thres;
    int32_t depth_mode;
    const void* depth;           /* [h,w] f32 or f64 per depth_mode, or NULL */
} sgb_fusion_view;

/* compute_mapping: mapping int64 [P,3] = (v, u, mask) exactly as fusion_utils.py:73-78. */
int sgb_fusion_map(sgb_ctx* ctx, const sgb_fusion_view* v, int64_t* mapping, void* stream);

/*
 * One fused view of fusion.py:127-144 without the host round trip: project, test, gather the
 * C-channel pixel feature from `features` ([C,h,w], f16 or f32) and add it into feat_sum [P,C]
 * f32; count [P] f32 += 1 for every visible Gaussian.  n_visible_dev (int32, device, may be
 * NULL) receives the number of visible Gaussians of this view.
 */
int sgb_fusion_accumulate(sgb_ctx* ctx, const sgb_fusion_view* v, const void* features,
                          int32_t C, int32_t feat_dtype, float* feat_sum, float* count,
                          int32_t* n_visible_dev, void* stream);

/* fusion.py:146-147: count[count==0] = 1e-5; feat_sum /= count (in place). */
int sgb_fusion_normalize(int32_t P, int32_t C, float* feat_sum, float* count, void* stream);

/* ---- semantic head: what every render_chn caller runs on the rendered feature image.
 *
 * sgb_semantic_head: render (C, N) planar fp32 with N = H*W, text (K, C) row-major:
 *     sim[k][p]  = sum_c text[k][c] * render[c][p] / (||render[:, p]||_2 + 1e-8)   eval_segmentation.py:155-156
 *     label[p]   = argmax_{k >= first_class} sim[k][p] - first_class                eval_segmentation.py:157
 * in ONE pass over the image.  sim (K, N) and label (N, int64) are optional (NULL to skip; a label-only
 * call never writes the K planes).  ctx is needed only for K > 32 with a label map (scratch), else may be NULL.
 *
 * sgb_feature_logits: out[p][k] = sum_c features[p][c] * text[k][c]  (einsum "cq,dq->dc",
 * eval_segmentation.py:132, view_viser.py:185) with row pitch Kpad >= K, columns K..Kpad-1 zeroed — blending is
 * linear, so rendering these Kpad channels gives the un-normalised similarities (and hence the label map)
 * without ever materialising the (C, H, W) feature image.
 *
 * sgb_label_argmax: label[p] = argmax_{first_class <= k < K} planes[k][p] - first_class
 * (rendering[1:].argmax(dim=0), eval_segmentation.py:144). */
/* Distillation loss of a rendered feature image against per-pixel class embeddings and its gradient, one pass:
 *     loss = -(1 / (C Nv)) sum_p <render[:, p], class_emb[label(p)]>,   dL_drender[c][p] = -class_emb[label(p)][c] / (C Nv)
 * render / dL_drender (C, N) planar fp32, class_emb (K, C), labels (N) int32 or int64.  A label outside [0, K)
 * (e.g. -1 / 255 "unannotated") marks an IGNORED pixel: zero gradient, no loss term, not counted in the normaliser
 * Nv = number of valid pixels (Nv = N when every label is in range).
 * loss: TWO doubles on the device (zeroed by the call): [0] the loss, [1] Nv. */
int sgb_distill_loss(int32_t C, int32_t K, int64_t N, const float* render, const float* class_emb, const void* labels,
                     int32_t labels_are_int64, float* dL_drender, double* loss, void* stream);
int sgb_semantic_head(sgb_ctx* ctx, int32_t C, int32_t K, int64_t N, const float* render, const float* text,
                      int32_t first_class, float* sim, int64_t* label, void* stream);
int sgb_feature_logits(int32_t P, int32_t C, int32_t K, int32_t Kpad, const float* features, const float* text,
                       float* out, void* stream);
int sgb_label_argmax(int32_t K, int32_t first_class, int64_t N, const float* planes, int64_t* label, void* stream);

/* ---- 3-nearest-neighbour mean squared distance: `distCUDA2` of the reference's
 * simple-knn extension (submodules/simple-knn/simple_knn.cu:185-220, spatial.cu), used by
 * GaussianModel.create_from_pcd (model/gaussian_model.py:150-186).  points (P,3) fp32 device, mean_dist2 (P) fp32
 * device: (d1+d2+d3)/3 of the three smallest squared distances to OTHER points (exact, bit-identical to the
 * reference).  Fewer than 4 points leave FLT_MAX terms, as in the reference. */
int sgb_knn_mean_dist2(sgb_ctx* ctx, int32_t P, const float* points, float* mean_dist2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGB200_H_INCLUDED */
