// sgb200 internal header: state layouts, error handling, small device helpers.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/sgb200.h"

#define SGB_TILE_PIX (SGB_TILE * SGB_TILE)

namespace sgb {

// ------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define SGB_CUDA(call)                                            \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return sgb::cuda_fail(e__, #call); \
    } while (0)
#define SGB_LAUNCH_CHECK(what, dbg, stream)                                   \
    do {                                                                      \
        cudaError_t e__ = cudaGetLastError();                                 \
        if (e__ == cudaSuccess && (dbg)) e__ = cudaStreamSynchronize(stream); \
        if (e__ != cudaSuccess) return sgb::cuda_fail(e__, what);             \
    } while (0)

// ------------------------------------------------------------------ opaque state layouts
// One 32-byte record per Gaussian: everything the blend needs, one DRAM sector per gather
// (the reference gathers means2D 8 B + conic_opacity 16 B [+ depth 4 B] from three arrays,
// forward.cu:318-321).
struct __align__(16) SplatRec {
    float mx, my;          // pixel-space mean            (GeometryState::means2D)
    float depth;           // view-space z                (GeometryState::depths)
    float pad;
    float cx, cy, cz, op;  // conic (a,b,c) and opacity   (GeometryState::conic_opacity)
};
static_assert(sizeof(SplatRec) == 32, "SplatRec must be 32 bytes");

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct GeomView {  // carved out of geometry_state (caller-owned, P-sized)
    SplatRec* rec;
    float* cov3D;            // [P,6]
    float* rgb;              // [P,3]   SH path only
    uint8_t* clamped;        // [P,3]
    uint32_t* tiles_touched; // [P]
    size_t bytes;
    static GeomView carve(void* base, int64_t P) {
        GeomView g;
        char* p = (char*)base;
        size_t off = 0;
        g.rec = (SplatRec*)(p + off); off += align_up(sizeof(SplatRec) * P);
        g.cov3D = (float*)(p + off); off += align_up(sizeof(float) * 6 * P);
        g.rgb = (float*)(p + off); off += align_up(sizeof(float) * 3 * P);
        g.clamped = (uint8_t*)(p + off); off += align_up(3 * (size_t)P);
        g.tiles_touched = (uint32_t*)(p + off); off += align_up(sizeof(uint32_t) * P);
        g.bytes = off;
        return g;
    }
};

struct BinView {  // carved out of binning_state (R-sized)
    uint32_t* point_list;  // [R] Gaussian ids, sorted by (tile, depth bits, id)
    size_t bytes;
    static BinView carve(void* base, int64_t R) {
        BinView b;
        b.point_list = (uint32_t*)base;
        b.bytes = align_up(sizeof(uint32_t) * (size_t)(R > 0 ? R : 1));
        return b;
    }
};

struct ImgView {  // carved out of image_state
    float* final_T;        // [H*W]   ImageState::accum_alpha
    uint32_t* n_contrib;   // [H*W]
    uint2* ranges;         // [tiles]
    uint32_t* tile_last;   // [tiles] max n_contrib inside the tile (backward start point)
    size_t bytes;
    static ImgView carve(void* base, int W, int H) {
        ImgView v;
        char* p = (char*)base;
        size_t N = (size_t)W * H;
        size_t tiles = (size_t)((W + SGB_TILE - 1) / SGB_TILE) * ((H + SGB_TILE - 1) / SGB_TILE);
        size_t off = 0;
        v.final_T = (float*)(p + off); off += align_up(4 * N);
        v.n_contrib = (uint32_t*)(p + off); off += align_up(4 * N);
        v.ranges = (uint2*)(p + off); off += align_up(8 * tiles);
        v.tile_last = (uint32_t*)(p + off); off += align_up(4 * tiles);
        v.bytes = off;
        return v;
    }
};

// ------------------------------------------------------------------ scratch context
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t n);  // grows (cudaFree + cudaMalloc); returns SGB_OK / SGB_E_NOMEM
};

}  // namespace sgb

namespace sgb {
// Built-in stage tracing (the reference has none; its only timing hook is train.py's per-iteration
// event pair).  When enabled, every stage is bracketed by a CUDA event pair recorded on the
// caller's stream; sgb_profile_read() sums the pairs recorded since the last read.
enum Stage {
    ST_PREPROCESS = 0, ST_DEPTH_SORT, ST_SCAN, ST_EMIT, ST_TILE_SORT, ST_RANGES, ST_BLEND_FWD,
    ST_BLEND_BWD, ST_GEOM_BWD, ST_FUSION_PROJECT, ST_FUSION_SORT, ST_FUSION_GATHER, ST_ALPHA, ST_DFEATURE,
    ST_COUNT
};
constexpr int kProfRing = 320;
struct Profiler {
    bool on = false;
    cudaEvent_t ev[ST_COUNT][kProfRing][2];
    int n[ST_COUNT];
    bool created = false;
};
}  // namespace sgb

namespace sgb {
constexpr int kMaxBatch = 8;  // views per batched call == weight-pool slots kept per ctx

// One weight pool (blend_v3.cu) = the per-tile alpha*T rows of ONE view.  A ctx keeps up to kMaxBatch of them so
// that the backward of each view of a batch (or of a forward-forward-...-backward-backward sequence) finds the
// rows its forward built.  A slot is identified by the view's binning-state pointer: a new forward through the
// same pointer necessarily overwrites that slot, so a slot can never describe a different view's instance list.
struct PoolSlot {
    Scratch mem;
    bool valid = false;
    const void* key_bin = nullptr;
    int64_t key_R = 0;
    int key_W = 0, key_H = 0, key_P = 0;
    uint32_t chunks = 0;   // capacity the slot was carved with
    uint64_t stamp = 0;    // LRU clock
};
}  // namespace sgb

struct sgb_ctx {
    int device = 0;
    sgb::Profiler prof;
    uint64_t launches = 0;       // kernels of this library launched through this ctx
    uint64_t lib_launches = 0;   // CUB device-wide calls (each several kernels)
    sgb::Scratch geom;     // depth-sort keys/values, offsets, CUB temp (one slice per view of a batch)
    sgb::Scratch bin;      // unsorted / sorted tile keys, unsorted values, CUB temp
    sgb::Scratch misc;     // fusion: pixel-sorted visible list, z-buffer
    sgb::PoolSlot pools[sgb::kMaxBatch];  // per-tile weight rows of the C-channel blend (blend_v3.cu)
    uint64_t pool_clock = 0;
    uint64_t pool_chunks_hint = 0;  // high-water mark of the pool demand (chunks)
    int64_t* pinned = nullptr;  // host-pinned readback slots (1 KB)
    int64_t stat_blended_pairs = 0;  // last alpha pass: blended (pixel, Gaussian) pairs
    int64_t stat_pool_chunks = 0;    // last alpha pass: 16-entry weight-row chunks in use
    cudaEvent_t feature_grad_event = nullptr;  // caller-owned; recorded when dL_dcolors is final (sgb200.h)
    // cached layout of the last sgb_forward_geometry[_batch] call (consumed by sgb_forward_render[_batch])
    int64_t last_P = 0;
    int last_V = 0;
    uint32_t* d_perm[sgb::kMaxBatch] = {};     // [P] Gaussian ids in (depth bits, id) order, per view of the batch
    uint32_t* d_offsets[sgb::kMaxBatch] = {};  // [P] inclusive scan of tiles_touched in that order
};

namespace sgb {

struct StageTimer {  // RAII event pair around one stage
    sgb_ctx* c; int st; cudaStream_t s; int slot;
    StageTimer(sgb_ctx* ctx, int stage, cudaStream_t stream) : c(ctx), st(stage), s(stream), slot(-1) {
        if (c && c->prof.on && c->prof.n[st] < kProfRing) {
            slot = c->prof.n[st]++;
            cudaEventRecord(c->prof.ev[st][slot][0], s);
        }
    }
    ~StageTimer() {
        if (slot >= 0) cudaEventRecord(c->prof.ev[st][slot][1], s);
    }
};

// cudaFuncSetAttribute is per device: a process-wide "done" flag would leave the second device of a
// single-process multi-GPU caller without its opt-in shared-memory size.  Returns true the first time it is
// called with this flag array on the current device.
struct DeviceOnce {
    bool done[64] = {};
    bool first_use_on_device() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return true;
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};

// ------------------------------------------------------------------ stage launchers
int launch_preprocess(const sgb_view_inputs& in, GeomView g, int32_t* radii, uint32_t* depth_keys,
                      cudaStream_t s);
int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t s);
// Depth order + scan of V views of the same Gaussians (cams[v] replaces the camera fields of `in`; V = 1 with
// cams = nullptr is the single-view call): everything is enqueued back to back, ONE stream sync reads all R.
int run_depth_order_and_scan(sgb_ctx* ctx, const sgb_view_inputs& in, int V, const sgb_camera* cams,
                             void* const* geometry_states, int32_t* const* radii, int64_t* R_host, cudaStream_t s);
int reserve_binning(sgb_ctx* ctx, const sgb_view_inputs& in, int64_t R, cudaStream_t s);
int run_binning(sgb_ctx* ctx, const sgb_view_inputs& in, int view_slot, int64_t R, GeomView g, BinView b, ImgView im,
                const int32_t* radii, cudaStream_t s);
int launch_blend_forward(const sgb_view_inputs& in, GeomView g, BinView b, ImgView im, const float* colors,
                         float* out_color, float* out_depth, cudaStream_t s);
int launch_blend_backward(const sgb_view_inputs& in, GeomView g, BinView b, ImgView im, const float* colors,
                          const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                          float* dL_dcolors, cudaStream_t s);
// C > 4 forward blend of one view, split so that a batch can enqueue the alpha passes of all its views before the
// one stream sync that validates their weight pools:
//   alpha    alpha pass into a weight-pool slot (no sync; pool header -> pinned slot `view_slot`)
//   finish   after the stream was synchronised: 0 = slot valid, 1 = the pool overflowed (slot grown: alpha again)
//   gemm     forward GEMM from the validated slot
int blend_forward_v3_alpha(sgb_ctx* ctx, int view_slot, const sgb_view_inputs& in, int64_t R, GeomView g, BinView b,
                           ImgView im, cudaStream_t s);
int blend_forward_v3_finish(sgb_ctx* ctx, int view_slot, const sgb_view_inputs& in, int64_t R, BinView b);
int blend_forward_v3_gemm(sgb_ctx* ctx, const sgb_view_inputs& in, int64_t R, BinView b, ImgView im,
                          const float* colors, float* out_color, cudaStream_t s);
// backward of one view in two halves (a batch runs all dL/dfeature kernels first, records the feature-gradient
// event, then the chain kernels): `prepare` finds or rebuilds the view's weight rows.
int blend_backward_v3_dfeature(sgb_ctx* ctx, const sgb_view_inputs& in, int64_t R, GeomView g, BinView b, ImgView im,
                               const float* dL_dpix, float* dL_dcolors, cudaStream_t s);
int blend_backward_v3_chain(sgb_ctx* ctx, const sgb_view_inputs& in, int64_t R, GeomView g, BinView b, ImgView im,
                            const float* colors, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                            float* dL_dopacity, cudaStream_t s);
int launch_geom_backward(const sgb_view_inputs& in, GeomView g, const int32_t* radii, const float* cov3D,
                         const float* dL_dcolor_rgb, const sgb_view_grads& gr, cudaStream_t s);

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waits for the phase with the given parity.  try_wait sleeps in hardware between polls; a bounded
// spin turns a protocol bug into a trap (launch failure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spins > (1u << 24)) __trap();
    }
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP); dst/src 16-B aligned,
// bytes a multiple of 16; completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// packed fp32 FMA (sm_100+): d.xy = a.xy * b.xy + c.xy, one issue slot for two FMAs.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
// auxiliary.h:46-56 (getRect).  max_radius is an int parameter: the float radius is converted at
// the call.  Used by preprocess and re-derived by the instance emitter exactly like
// duplicateWithKeys does (rasterizer_impl.cu:91).
__device__ __forceinline__ void get_rect(const float2 p, int max_radius, uint2& rect_min, uint2& rect_max,
                                         dim3 grid) {
    rect_min = {min(grid.x, max((int)0, (int)((p.x - max_radius) / SGB_TILE))),
                min(grid.y, max((int)0, (int)((p.y - max_radius) / SGB_TILE)))};
    rect_max = {min(grid.x, max((int)0, (int)((p.x + max_radius + SGB_TILE - 1) / SGB_TILE))),
                min(grid.y, max((int)0, (int)((p.y + max_radius + SGB_TILE - 1) / SGB_TILE)))};
}
__device__ __forceinline__ float4 ldg_nc_f4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
#endif

}  // namespace sgb
