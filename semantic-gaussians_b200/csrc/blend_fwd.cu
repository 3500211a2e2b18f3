// Per-tile front-to-back alpha compositing (forward) for the RGB / RGB-D path (C <= 4).
// Contract: reference forward.cu:262-375 and rgbd/cuda_rasterizer/forward.cu:261-393 (median depth).
//
// One CTA per 16x16 tile, thread = pixel, Gaussians staged in shared memory per batch like the
// reference.  The per-pixel alpha / transmittance chain and the accumulation `C[ch] += f*alpha*T`
// are the reference's statement sequence verbatim, so pixels, depth, final_T and n_contrib are
// bit-identical.  The kernel is templated on a channel-chunk width and a bulk-copy staging mode that
// the wide (C > 4) path used in its first generation; that path now lives in blend_v3.cu.
#include <cstdlib>
#include "common.cuh"

namespace sgb {

namespace {

constexpr int kThreads = SGB_TILE_PIX;  // 256, one per pixel
constexpr int kBatch = 64;              // Gaussians per pipeline stage

template <int CH>
struct __align__(16) FwdStage {
    float4 recA[kBatch];     // mx, my, depth, pad
    float4 recB[kBatch];     // conic.x, conic.y, conic.z, opacity
    float feat[kBatch][CH];  // feature slice [ch0, ch0+CH) of each staged Gaussian
};

// CH   : channels per CTA (multiple of 4)
// BULK : feature rows are 16-byte aligned slices (C % 4 == 0) -> TMA bulk copies + mbarrier;
//        otherwise cooperative scalar loads + __syncthreads (C = 3 and other odd widths).
// EXACT: accumulate as the reference spells it, C[ch] += f * alpha * T (bit-identical colours;
//        used for the <= 4-channel RGB / RGB-D path).
template <int CH, bool BULK, bool EXACT, bool DEPTH>
__global__ void __launch_bounds__(kThreads) blend_forward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int C,
    const SplatRec* __restrict__ rec, const float* __restrict__ features, const float* __restrict__ bg_color,
    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ tile_last,
    float* __restrict__ out_color, float* __restrict__ out_depth) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdStage<CH>* stage = reinterpret_cast<FwdStage<CH>*>(smem_raw);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t s_last;

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int tile = blockIdx.x;
    const int ch0 = blockIdx.y * CH;
    const int nch = min(CH, C - ch0);
    const int tid = threadIdx.x;
    const uint32_t tx = tid & (SGB_TILE - 1), ty = tid >> 4;
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};
    const uint2 pix = {pix_min.x + tx, pix_min.y + ty};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    bool done = !inside;

    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    const int nbatches = (total + kBatch - 1) / kBatch;

    if (tid == 0) {
        if (BULK) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            mbar_fence_init();
        }
        s_last = 0;
    }
    __syncthreads();

    auto issue = [&](int b) {
        FwdStage<CH>& st = stage[b & 1];
        const int base = b * kBatch;
        const int cnt = min(kBatch, total - base);
        if (BULK) {
            if (tid == 0) mbar_arrive_expect_tx(&bar[b & 1], (uint32_t)cnt * (32u + (uint32_t)nch * 4u));
            if (tid < cnt) {
                const uint32_t id = point_list[range.x + base + tid];
                bulk_g2s(&st.recA[tid], reinterpret_cast<const float4*>(rec + id), 16, &bar[b & 1]);
                bulk_g2s(&st.recB[tid], reinterpret_cast<const float4*>(rec + id) + 1, 16, &bar[b & 1]);
                bulk_g2s(&st.feat[tid][0], features + (size_t)id * C + ch0, (uint32_t)nch * 4u, &bar[b & 1]);
            }
        } else {
            if (tid < cnt) {
                const uint32_t id = point_list[range.x + base + tid];
                const float4* rp = reinterpret_cast<const float4*>(rec + id);
                st.recA[tid] = __ldg(rp);
                st.recB[tid] = __ldg(rp + 1);
            }
            for (int e = tid; e < cnt * nch; e += kThreads) {
                const int j = e / nch, k = e - j * nch;
                const uint32_t id = point_list[range.x + base + j];
                st.feat[j][k] = __ldg(features + (size_t)id * C + ch0 + k);
            }
        }
    };

    float T = 1.0f;
    uint32_t contributor = 0;
    uint32_t last_contributor = 0;
    float D = 15.0f;  // rgbd forward.cu:308: default median depth
    float acc[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) acc[k] = 0.f;

    int issued = 0, consumed = 0;
    if (nbatches > 0) { issue(0); issued = 1; }

    for (int b = 0; b < nbatches; b++) {
        // Block-wide early out (forward.cu:310-312).  Also the WAR fence of the stage that the
        // prefetch below overwrites: every thread has finished reading batch b-1.
        const int num_done = __syncthreads_count(done);
        if (num_done == kThreads) break;
        if (b + 1 < nbatches) { issue(b + 1); issued = b + 2; }
        if (BULK) mbar_wait(&bar[b & 1], (uint32_t)((b >> 1) & 1));
        else __syncthreads();
        consumed = b + 1;

        const FwdStage<CH>& st = stage[b & 1];
        const int cnt = min(kBatch, total - b * kBatch);
        for (int j = 0; !done && j < cnt; j++) {
            contributor++;
            // forward.cu:333-352, verbatim arithmetic
            const float4 a = st.recA[j];
            const float2 xy = {a.x, a.y};
            const float2 d = {xy.x - pixf.x, xy.y - pixf.y};
            const float4 con_o = st.recB[j];
            const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
            if (power > 0.0f) continue;
            const float alpha = min(0.99f, con_o.w * exp(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) {
                done = true;
                continue;
            }
            if (EXACT) {
#pragma unroll
                for (int k = 0; k < CH; k++)
                    if (k < nch) acc[k] += st.feat[j][k] * alpha * T;  // forward.cu:355-356
            } else {
                const float w = alpha * T;
                const float2 w2 = {w, w};
#pragma unroll
                for (int k = 0; k < CH; k += 4) {
                    const float4 f = *reinterpret_cast<const float4*>(&st.feat[j][k]);
                    float2 r0 = ffma2(make_float2(f.x, f.y), w2, make_float2(acc[k], acc[k + 1]));
                    float2 r1 = ffma2(make_float2(f.z, f.w), w2, make_float2(acc[k + 2], acc[k + 3]));
                    acc[k] = r0.x; acc[k + 1] = r0.y; acc[k + 2] = r1.x; acc[k + 3] = r1.y;
                }
            }
            if (DEPTH) {
                if (T > 0.5f && test_T < 0.5) D = a.z;  // rgbd forward.cu:368-372: median depth
            }
            T = test_T;
            last_contributor = contributor;
        }
    }
    // never leave with a bulk copy in flight into this CTA's shared memory
    if (BULK && issued > consumed) mbar_wait(&bar[consumed & 1], (uint32_t)((consumed >> 1) & 1));

    if (blockIdx.y == 0) {
        if (inside) {
            final_T[pix_id] = T;
            n_contrib[pix_id] = last_contributor;
            if (DEPTH) out_depth[pix_id] = D;
            atomicMax(&s_last, last_contributor);
        }
        __syncthreads();
        if (tid == 0) tile_last[tile] = s_last;
    }
    if (inside) {
        const size_t plane = (size_t)H * W;
#pragma unroll
        for (int k = 0; k < CH; k++)
            if (k < nch) out_color[(size_t)(ch0 + k) * plane + pix_id] = acc[k] + T * bg_color[ch0 + k];
    }
}

template <int CH, bool BULK, bool EXACT, bool DEPTH>
int launch_one(const sgb_view_inputs& in, GeomView g, BinView b, ImgView im, const float* colors,
               float* out_color, float* out_depth, cudaStream_t s) {
    const int tiles = ((in.W + SGB_TILE - 1) / SGB_TILE) * ((in.H + SGB_TILE - 1) / SGB_TILE);
    const int chunks = (in.C + CH - 1) / CH;
    const size_t smem = 2 * sizeof(FwdStage<CH>);
    auto kern = blend_forward_kernel<CH, BULK, EXACT, DEPTH>;
    static DeviceOnce attr_set;
    if (attr_set.first_use_on_device()) {
        SGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    kern<<<dim3(tiles, chunks), kThreads, smem, s>>>(im.ranges, b.point_list, in.W, in.H, in.C, g.rec, colors,
                                                    in.background, im.final_T, im.n_contrib, im.tile_last,
                                                    out_color, out_depth);
    SGB_LAUNCH_CHECK("blend_forward_kernel", in.debug, s);
    return SGB_OK;
}

}  // namespace

int launch_blend_forward(const sgb_view_inputs& in, GeomView g, BinView b, ImgView im, const float* colors,
                         float* out_color, float* out_depth, cudaStream_t s) {
    // RGB / RGB-D path (C <= 4): the reference's accumulation order, bit for bit.  Wider rasters go
    // through the weights-once pipeline in blend_v3.cu.
    if (in.C > 4) {
        set_error("launch_blend_forward handles C <= 4 only");
        return SGB_E_INVALID;
    }
    if (out_depth) return launch_one<4, false, true, true>(in, g, b, im, colors, out_color, out_depth, s);
    return launch_one<4, false, true, false>(in, g, b, im, colors, out_color, nullptr, s);
}

}  // namespace sgb
