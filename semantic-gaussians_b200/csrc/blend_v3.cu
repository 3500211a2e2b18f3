// C-channel blend, "weights once" pipeline (C > 4).
//
// Measured on the K3 scene (1 M Gaussians, 1080p): a tile walks ~280 list entries before all its
// pixels saturate, but only ~115 of them touch any pixel of the tile (the reference bins by the
// 3-sigma square of the major axis, rasterizer_impl.cu:91 / forward.cu:229-235), and the scalar
// alpha / transmittance chain costs about as many issue slots as a 64-channel accumulation.  A
// kernel that fuses chain and accumulation per channel chunk (blend_fwd.cu / blend_bwd.cu v2)
// therefore spends most of its instructions re-deriving the same weights in every chunk, forward
// and backward.  Here the chain runs ONCE per view:
//
//   alpha_pass        one CTA per tile, thread = pixel: the reference's chain verbatim
//                     (forward.cu:326-363) -> final_T, n_contrib, and for every Gaussian that
//                     touches the tile a 1 KB row of weights w[pixel] = alpha * T (0 where the
//                     pixel skips it) appended to a per-tile linked list of 16-entry chunks.
//   blend_forward_v3  CTA = (tile, 64-channel chunk), barrier-free: each warp streams the tile's
//                     weight rows and feature slices straight from L2 and accumulates an
//                     8 px x 8 ch register micro-tile per lane (outer product, packed FMA).
//   chain_backward_v3 CTA = tile: s = <feature, dL/dout> per (pixel, Gaussian) for all channels
//                     (register micro-tiles + transposed shuffle reduce), then the reference's
//                     back-to-front chain (backward.cu:477-550) in dot-product form -> dL/dmean2D,
//                     dL/dconic, dL/dopacity.
//   dfeature_v3       CTA = (tile, 64-channel chunk): dL/dfeature[g][ch] = sum_px w * dL/dout, one
//                     warp per 8-channel slice over all 256 pixels, one 32-byte reduction per
//                     (Gaussian, tile, slice).
//
// Results are unchanged: the integer outputs come from the verbatim chain; every accumulator still
// adds its Gaussians in depth order.
#include <cstdlib>
#include <cstring>
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include "common.cuh"
#include "blend_pool.cuh"

namespace sgb {

namespace {

constexpr int kThreads = SGB_TILE_PIX;

// Lane -> operand-group mapping of the register-tiled GEMM loops (round 2, measured with tools/lds_probe.cu under ncu
// on B200): a warp-wide LDS.128 costs 2 shared-memory wavefronts when every aligned group of 4 lanes reads at most 2
// distinct 16-byte chunks and each half-warp at most 8 (conflict-free) chunks, and 4 wavefronts otherwise.  With the
// natural split (one operand indexed by lane & 7, the other by lane >> 3) the lane & 7 operand pays 4 per load and the
// L1 data pipe — not the FMA pipe — bounded all three contraction kernels (ncu r02: 68 / 74 / 83 % of peak).  Giving
// each operand exactly one of the two low lane bits makes every operand load a 2-wavefront load.
__device__ __forceinline__ int lane_group8(int lane) { return (lane & 1) | (((lane >> 2) & 3) << 1); }  // bits 0, 2, 3
__device__ __forceinline__ int lane_group4(int lane) { return ((lane >> 1) & 1) | ((lane >> 4) << 1); } // bits 1, 4

template <int N>
__device__ __forceinline__ void xreduce_step(float (&v)[8], int lane, int step) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        const float send = upper ? v[i] : v[i + N / 2];
        const float keep = upper ? v[i + N / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
}

// ------------------------------------------------------------------------------------ alpha pass
constexpr int kAB = 32;  // list entries per staging round

struct __align__(16) AlphaSmem {
    float4 recA[kAB];
    float4 recB[kAB];
    uint32_t ids[kAB];
    float wbuf[kAB][SGB_TILE_PIX];
    uint32_t wmask[8];
    uint32_t slot_chunk[kAB];
    uint32_t cur_chunk;
    uint32_t s_last;
};

template <bool DEPTH>
__global__ void __launch_bounds__(kThreads) alpha_pass_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const SplatRec* __restrict__ rec, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    uint32_t* __restrict__ tile_last, float* __restrict__ out_depth, PoolView pool) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    AlphaSmem& sm = *reinterpret_cast<AlphaSmem*>(smem_raw);

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tx = tid & (SGB_TILE - 1), ty = tid >> 4;
    const uint2 pix = {(uint32_t)(tile % tiles_x) * SGB_TILE + tx, (uint32_t)(tile / tiles_x) * SGB_TILE + ty};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    bool done = !inside;

    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    const int nbatches = (total + kAB - 1) / kAB;
    const uint32_t dbase = range.x / kChunkEntries + (uint32_t)tile;
    if (tid == 0) { sm.cur_chunk = kNone; sm.s_last = 0; }

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float D = 15.0f;
    uint32_t n_tile = 0;  // entries appended so far (uniform)
    uint32_t n_blend = 0; // Gaussians blended into this pixel

    // Staging of the (id, splat record) batches is software-pipelined in warp 0's registers: the ids run two
    // batches ahead, the records (a dependent gather through the id) one batch ahead, so neither round trip
    // sits between two batches of the chain (it used to: two dependent L2/DRAM latencies per 32 entries).
    auto load_id = [&](int bb) -> uint32_t {
        const int i = bb * kAB + tid;
        return (tid < kAB && bb < nbatches && i < total) ? __ldg(point_list + range.x + i) : 0u;
    };
    uint32_t id_cur = load_id(0), id_nxt = load_id(1);
    float4 rA = make_float4(0.f, 0.f, 0.f, 0.f), rB = rA;
    if (tid < kAB && tid < total) {
        const float4* rp = reinterpret_cast<const float4*>(rec + id_cur);
        rA = __ldg(rp);
        rB = __ldg(rp + 1);
    }
    for (int b = 0; b < nbatches; b++) {
        const int num_done = __syncthreads_count(done);  // forward.cu:310-312
        if (num_done == kThreads) break;
        const int base = b * kAB;
        const int cnt = min(kAB, total - base);
        if (tid < cnt) {
            sm.ids[tid] = id_cur;
            sm.recA[tid] = rA;
            sm.recB[tid] = rB;
        }
        __syncthreads();
        if (tid < kAB) {  // records of batch b+1 (its ids are already here), ids of batch b+2
            id_cur = id_nxt;
            if (base + kAB + tid < total) {
                const float4* rp = reinterpret_cast<const float4*>(rec + id_cur);
                rA = __ldg(rp);
                rB = __ldg(rp + 1);
            }
            id_nxt = load_id(b + 2);
        }
        uint32_t my_mask = 0;
        for (int j = 0; j < cnt; j++) {
            float w = 0.f;
            if (!done) {
                // forward.cu:333-362 verbatim
                const float4 a = sm.recA[j];
                const float2 xy = {a.x, a.y};
                const float2 d = {xy.x - pixf.x, xy.y - pixf.y};
                const float4 con_o = sm.recB[j];
                const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
                if (!(power > 0.0f)) {
                    const float alpha = min(0.99f, con_o.w * exp(power));
                    if (!(alpha < 1.0f / 255.0f)) {
                        const float test_T = T * (1 - alpha);
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
                            w = alpha * T;
                            if (DEPTH) {
                                if (T > 0.5f && test_T < 0.5) D = a.z;
                            }
                            T = test_T;
                            last_contributor = (uint32_t)(base + j + 1);
                        }
                    }
                }
            }
            sm.wbuf[j][tid] = w;
            n_blend += (w != 0.f);
            if (__ballot_sync(0xffffffffu, w != 0.f)) my_mask |= 1u << j;
        }
        if (lane == 0) sm.wmask[warp] = my_mask;
        __syncthreads();
        uint32_t tm = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) tm |= sm.wmask[q];
        const int n_act = __popc(tm);
        if (n_act) {
            if (tid == 0) {
                uint32_t e = n_tile, cur = sm.cur_chunk;
                for (int k = 0; k < n_act; k++, e++) {
                    if ((e & (kChunkEntries - 1)) == 0) {
                        uint32_t nw = atomicAdd(&pool.hdr->counter, 1u);
                        if (nw >= pool.capacity) {
                            pool.hdr->overflow = 1;
                            nw = pool.capacity - 1;
                        }
                        pool.dir[dbase + e / kChunkEntries] = nw;
                        cur = nw;
                    }
                    sm.slot_chunk[k] = cur;
                }
                sm.cur_chunk = cur;
            }
            __syncthreads();
            int k = 0;
            for (uint32_t m = tm; m; m &= m - 1, k++) {
                const int j = __ffs(m) - 1;
                const uint32_t e = n_tile + k;
                WChunk& ck = pool.chunks[sm.slot_chunk[k]];
                const int s = e & (kChunkEntries - 1);
                ck.w[s][tid] = sm.wbuf[j][tid];
                if (tid == 0) {
                    uint32_t strips = 0;
#pragma unroll
                    for (int q = 0; q < 8; q++) strips |= ((sm.wmask[q] >> j) & 1u) << q;
                    ck.meta[s] = make_uint2(sm.ids[j], strips);
                }
            }
            n_tile += n_act;
        }
    }
    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        if (DEPTH) out_depth[pix_id] = D;
        atomicMax(&sm.s_last, last_contributor);
    }
    n_blend = __reduce_add_sync(0xffffffffu, n_blend);
    if (lane == 0 && n_blend) atomicAdd(&pool.hdr->blended, (unsigned long long)n_blend);
    __syncthreads();
    if (tid == 0) {
        tile_last[tile] = sm.s_last;
        pool.count[tile] = n_tile;
        pool.dirbase[tile] = dbase;
    }
}

// ------------------------------------------------------------------------------------ forward
// Non-finite features.  The GEMM-shaped kernels multiply every (pixel, entry) pair of a tile, zero weights
// included, and 0 * inf = NaN: one non-finite feature row would poison every pixel of every tile its Gaussian is
// binned to, where the reference only touches the pixels that actually blend it (forward.cu:340-356 `continue`s
// before the accumulation).  A pair is blended exactly when its weight alpha * T is non-zero (alpha >= 1/255 and
// T >= 1e-4 on that path), so the exact semantics are "accumulate only where w != 0".  Guarding every FMA would
// double the inner loop; instead the epilogue tests the accumulators (acc * 0 summed: NaN iff any accumulator is
// non-finite, 32 packed FMAs per lane) and only a warp that sees a non-finite value recomputes its 32 px x CH
// slice with the guarded loop below, straight from the weight rows and feature rows in global memory.
template <int MCH>
__device__ __forceinline__ bool acc_nonfinite(const float2 (&acc)[8][MCH / 2]) {
    float2 z = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < MCH / 2; k++) z = ffma2(acc[i][k], make_float2(0.f, 0.f), z);
    const float t = z.x + z.y;
    return __any_sync(0xffffffffu, t != t);
}

// RING = true: lane cg owns channels {4cg..4cg+3} U {32+4cg..} of the slice (blend_forward_tma_kernel);
// false: channels cg*MCH .. cg*MCH+MCH-1 (blend_forward_v3_kernel).  Self-contained (own accumulators, own
// stores) so that the fast path's accumulators never have their address taken.
template <int MCH, bool RING>
__device__ __noinline__ void forward_redo_guarded(const PoolView& pool, uint32_t n, uint32_t dbase,
                                                  const float* __restrict__ features, int C, int ch0, int nch,
                                                  int warp, int woff, int cg, const float* __restrict__ bg_color,
                                                  const float* __restrict__ final_T, int W, int H, uint32_t row,
                                                  uint32_t col0, float* __restrict__ out_color) {
    float acc[8][MCH];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < MCH; k++) acc[i][k] = 0.f;
    for (uint32_t e = 0; e < n; e++) {
        const WChunk* ck = pool.chunks + chunk_of(pool, dbase, (int)(e / kChunkEntries));
        const int s = (int)(e & (kChunkEntries - 1));
        const uint2 meta = ck->meta[s];
        if (!((meta.y >> warp) & 1u)) continue;
        const float* fr = features + (size_t)meta.x * C + ch0;
        float f[MCH];
#pragma unroll
        for (int k = 0; k < MCH; k++) {
            const int chl = RING ? ((k >> 2) * 32 + cg * 4 + (k & 3)) : (cg * MCH + k);
            f[k] = chl < nch ? __ldg(fr + chl) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float w = ck->w[s][woff + i];
            if (w != 0.f) {
#pragma unroll
                for (int k = 0; k < MCH; k++) acc[i][k] = fmaf(f[k], w, acc[i][k]);
            }
        }
    }
    if (row >= (uint32_t)H) return;
    const size_t plane = (size_t)H * W;
#pragma unroll
    for (int k = 0; k < MCH; k++) {
        const int chl = RING ? ((k >> 2) * 32 + cg * 4 + (k & 3)) : (cg * MCH + k);
        if (chl >= nch) continue;
        const float bgc = bg_color[ch0 + chl];
        float* dst = out_color + (size_t)(ch0 + chl) * plane + (size_t)W * row + col0;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (col0 + i < (uint32_t)W) dst[i] = acc[i][k] + final_T[(size_t)W * row + col0 + i] * bgc;
    }
}

template <int CH, bool VEC>
__global__ void __launch_bounds__(kThreads, 2) blend_forward_v3_kernel(
    int W, int H, int C, const float* __restrict__ features, const float* __restrict__ bg_color,
    const float* __restrict__ final_T, PoolView pool, float* __restrict__ out_color) {
    constexpr int MCH = CH / 8;
    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int nchunksC = (C + CH - 1) / CH;
    const int tile = blockIdx.x / nchunksC;           // chunk index fastest: the CTAs of one tile are
    const int ch0 = (blockIdx.x % nchunksC) * CH;     // co-scheduled and share its weight rows in L2
    const int nch = min(CH, C - ch0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pg = lane >> 3, cg = lane & 7;
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};

    float2 acc[8][MCH / 2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < MCH / 2; k++) acc[i][k] = make_float2(0.f, 0.f);

    const uint32_t n = pool.count[tile];
    const uint32_t dbase = pool.dirbase[tile];
    const int woff = warp * 32 + pg * 8;
    const int foff = ch0 + cg * MCH;
    for (uint32_t e = 0; e < n;) {
        const WChunk* ck = pool.chunks + chunk_of(pool, dbase, (int)(e / kChunkEntries));
        const int m = (int)min((uint32_t)kChunkEntries, n - e);
        for (int s = 0; s < m; s++) {
            const uint2 meta = ck->meta[s];
            if (!((meta.y >> warp) & 1u)) continue;
            const float4* wp = reinterpret_cast<const float4*>(&ck->w[s][woff]);
            const float4 w0 = wp[0], w1 = wp[1];
            const float* fr = features + (size_t)meta.x * C + foff;
            float2 f[MCH / 2];
            if (VEC) {
#pragma unroll
                for (int q = 0; q < MCH / 4; q++) {
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cg * MCH + 4 * q < nch) t = __ldg(reinterpret_cast<const float4*>(fr) + q);
                    f[2 * q] = make_float2(t.x, t.y);
                    f[2 * q + 1] = make_float2(t.z, t.w);
                }
            } else {
#pragma unroll
                for (int k = 0; k < MCH / 2; k++) {
                    f[k].x = (cg * MCH + 2 * k < nch) ? __ldg(fr + 2 * k) : 0.f;
                    f[k].y = (cg * MCH + 2 * k + 1 < nch) ? __ldg(fr + 2 * k + 1) : 0.f;
                }
            }
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float2 w2 = make_float2(wv[i], wv[i]);
#pragma unroll
                for (int k = 0; k < MCH / 2; k++) acc[i][k] = ffma2(f[k], w2, acc[i][k]);
            }
        }
        e += m;
    }
    // out = acc + T * bg (forward.cu:372-373)
    const uint32_t row = pix_min.y + 2 * warp + (pg >> 1);
    const uint32_t col0 = pix_min.x + (pg & 1) * 8;
    if (acc_nonfinite<MCH>(acc)) {
        forward_redo_guarded<MCH, false>(pool, n, dbase, features, C, ch0, nch, warp, woff, cg, bg_color, final_T, W, H,
                                         row, col0, out_color);
        return;
    }
    if (row < (uint32_t)H) {
        const size_t plane = (size_t)H * W;
        const bool vec = ((W & 3) == 0) && (col0 + 8 <= (uint32_t)W);
        float Tv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) Tv[i] = (col0 + i < (uint32_t)W) ? final_T[(size_t)W * row + col0 + i] : 0.f;
#pragma unroll
        for (int k = 0; k < MCH; k++) {
            const int chl = cg * MCH + k;
            if (chl >= nch) continue;
            const float bgc = bg_color[ch0 + chl];
            float* dst = out_color + (size_t)(ch0 + chl) * plane + (size_t)W * row + col0;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = ((k & 1) ? acc[i][k / 2].y : acc[i][k / 2].x) + Tv[i] * bgc;
            if (vec) {
                reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (col0 + i < (uint32_t)W) dst[i] = o[i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------ async-copy helpers
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// 3-D tensor tile global -> shared through the TMA engine (SASS: UTMALDG): box corner (x, y, z) in elements, out-of-range
// elements arrive as zeros; completion is signalled on `bar` as the box's bytes.  dst 128-byte aligned.
__device__ __forceinline__ void tma_tile3d_g2s(void* dst_smem, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
// Orders this thread's earlier generic-proxy shared-memory accesses before later async-proxy (TMA) writes.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}


constexpr int kSeg = 64;  // entries per backward segment (4 chunks); S[8 warps][kSeg][32 lanes] = 64 KB

// ------------------------------------------------------------------------------------ GEMM-shaped kernels
// With the weights materialised per tile, the three C-wide contractions are small dense GEMMs over the
// tile's touching Gaussians (G ~ 115 on K3), done in fp32 on the CUDA cores (north_star: no tensor cores;
// the 1e-4 fp32 bar rules out TF32 anyway):
//     forward   out[256 px][64 ch]  = W^T[256 px][G]  . F[G][64 ch]      K = G      lane tile 8 px x 8 ch
//     s-pass    S[256 px][G]        = dL[256 px][C]   . F^T[C][G]        K = C      lane tile 8 px x 8 entries
//     dfeature  dF[G][64 ch]        = W[G][256 px]    . dL[256 px][64]   K = 256 px lane tile 4 entries x 8 ch
// Register tiles make every product 64-128 FMAs per 4-6 shared-memory loads and need no cross-lane
// reductions (an earlier shuffle-reduce formulation spent ~40 % of its issue slots on SHFL/FSEL/FADD).
// Forward GEMM with a TMA-fed ring.  With plain loads of the weight rows 67 % of the instructions were
// packed FMAs but only 28 % of the issue slots were used — every warp waited an L2/DRAM round trip per
// entry (ncu long_scoreboard 9.2 stalls per issue).  Here one ring stage = one 16-entry chunk: per staged
// Gaussian two 1-D bulk copies (cp.async.bulk: the 1 KB weight row and the 256 B feature slice), NS stages
// guarded by full/empty mbarriers.  Warp 0 is producer AND consumer, so nothing it does for production may
// block its math: the chunk indices come from the tile's directory (copied to shared memory up front, no
// pointer chasing), the (id, mask) records of the NEXT batch are fetched into registers one whole batch of
// math before they are needed, and the stage it refills is the one everybody left TWO batches ago, so the
// empty-barrier wait is already satisfied (refilling the stage of the previous batch coupled warp 0 to
// the slowest warp on every batch: ncu showed the FMA pipe 50 % idle with long-scoreboard stalls on top).
template <int CH, int NS>
__global__ void __launch_bounds__(kThreads, 2) blend_forward_tma_kernel(
    int W, int H, int C, const float* __restrict__ features, const float* __restrict__ bg_color,
    const float* __restrict__ final_T, PoolView pool, float* __restrict__ out_color) {
    constexpr int MCH = CH / 8;
    constexpr int ES = kChunkEntries;  // entries per stage
    constexpr int LA = NS - 2;         // batches in flight ahead of the one being consumed
    constexpr int kDirCap = 192;       // directory entries cached in shared memory (3072 active Gaussians / tile)
    struct Stage {
        float w[ES][SGB_TILE_PIX];
        float f[ES][CH];
    };
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage* stg = reinterpret_cast<Stage*>(smem_raw);
    __shared__ uint64_t full_bar[NS], empty_bar[NS];
    __shared__ uint32_t Cdir[kDirCap];
    __shared__ __align__(16) float Tsm[SGB_TILE_PIX];  // final_T of the tile (the epilogue used to stall on these loads)
    __shared__ float bgS[CH];

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int nchunksC = (C + CH - 1) / CH;
    const int tile = blockIdx.x / nchunksC;
    const int ch0 = (blockIdx.x % nchunksC) * CH;
    const int nch = min(CH, C - ch0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pg = lane_group4(lane), cg = lane_group8(lane);
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};

    const uint32_t n = pool.count[tile];
    const int nb = (int)((n + ES - 1) / ES);
    const uint32_t dbase = pool.dirbase[tile];
    if (tid == 0) {
        for (int i = 0; i < NS; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], kThreads / 32);
        }
        mbar_fence_init();
    }
    for (int k = tid; k < min(nb, kDirCap); k += kThreads) Cdir[k] = chunk_of(pool, dbase, k);
    {
        const uint32_t x = pix_min.x + (tid & (SGB_TILE - 1)), y = pix_min.y + (tid >> 4);
        Tsm[tid] = (x < (uint32_t)W && y < (uint32_t)H) ? final_T[(size_t)W * y + x] : 0.f;
        if (tid < nch) bgS[tid] = bg_color[ch0 + tid];
    }
    if (nch < CH)  // zero the never-written tail of every feature row once
        for (int e = tid; e < NS * ES * CH; e += kThreads) {
            const int k = e % CH;
            if (k >= nch) stg[e / (ES * CH)].f[(e / CH) % ES][k] = 0.f;
        }
    __syncthreads();

    // ---- producer (warp 0, lanes 0..15 = entry slots of a chunk)
    auto chunk_ptr = [&](int bi) {
        return pool.chunks + (bi < kDirCap ? Cdir[bi] : chunk_of(pool, dbase, bi));
    };
    auto load_meta = [&](int bi) {  // (Gaussian id, strip mask) of this lane's entry of batch bi
        uint2 m = make_uint2(0u, 0u);
        if (bi < nb && lane < min(ES, (int)n - bi * ES)) m = __ldg(&chunk_ptr(bi)->meta[lane]);
        return m;
    };
    auto issue = [&](int bi, uint2 meta) {  // warp 0, converged; bi < nb
        const int st = bi % NS;
        const int cnt = min(ES, (int)n - bi * ES);
        if (bi >= NS) mbar_wait(&empty_bar[st], (uint32_t)(((bi / NS) - 1) & 1));  // batch bi-NS released by all warps
        if (lane < cnt) {
            const WChunk* ck = chunk_ptr(bi);
            bulk_g2s(&stg[st].w[lane][0], &ck->w[lane][0], SGB_TILE_PIX * 4u, &full_bar[st]);
            bulk_g2s(&stg[st].f[lane][0], features + (size_t)meta.x * C + ch0, (uint32_t)nch * 4u, &full_bar[st]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&full_bar[st], (uint32_t)cnt * (SGB_TILE_PIX * 4u + (uint32_t)nch * 4u));
    };
    uint2 meta_next = make_uint2(0u, 0u);  // record of batch `pb`, the next one to issue
    int pb = 0;
    if (warp == 0) {
        uint2 m[LA];
#pragma unroll
        for (int i = 0; i < LA; i++) m[i] = load_meta(i);  // independent loads, one round trip
#pragma unroll
        for (int i = 0; i < LA; i++)
            if (i < nb) issue(i, m[i]);
        pb = min(LA, nb);
        meta_next = load_meta(pb);
    }

    float2 acc[8][MCH / 2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < MCH / 2; k++) acc[i][k] = make_float2(0.f, 0.f);

    const int woff = warp * 32 + pg * 8;
    // Lane cg accumulates channels {4cg..4cg+3} and {32+4cg..32+4cg+3} of the slice: each of its two LDS.128 then
    // reads 8 x 16 B that are contiguous across the 8 channel lanes (one 128-byte wavefront); the natural
    // assignment 8cg..8cg+7 spread them over 256 B = two wavefronts per load.
    auto entry = [&](const Stage& sg, int e) {
        const float4 w0 = *reinterpret_cast<const float4*>(&sg.w[e][woff]);
        const float4 w1 = *reinterpret_cast<const float4*>(&sg.w[e][woff + 4]);
        float2 f[MCH / 2];
#pragma unroll
        for (int q = 0; q < MCH / 4; q++) {
            const float4 t = *reinterpret_cast<const float4*>(&sg.f[e][q * 32 + cg * 4]);
            f[2 * q] = make_float2(t.x, t.y);
            f[2 * q + 1] = make_float2(t.z, t.w);
        }
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float2 w2 = make_float2(wv[i], wv[i]);
#pragma unroll
            for (int k = 0; k < MCH / 2; k++) acc[i][k] = ffma2(f[k], w2, acc[i][k]);
        }
    };
    for (int b = 0; b < nb; b++) {
        const int st = b % NS;
        const int cnt = min(ES, (int)n - b * ES);
        if (warp == 0 && pb < nb) {  // refill the stage of batch b-2 with batch b+LA
            issue(pb, meta_next);
            pb++;
            meta_next = load_meta(pb);  // lands while this batch is being consumed
        }
        mbar_wait(&full_bar[st], (uint32_t)((b / NS) & 1));
        // dense on purpose: skipping strips whose 32 weights are all zero (about 15 % of the entries)
        // breaks the unrolled load/FMA software pipeline and measured 10 % slower on K3
        if (cnt == ES) {
#pragma unroll
            for (int e = 0; e < ES; e++) entry(stg[st], e);
        } else {
            for (int e = 0; e < cnt; e++) entry(stg[st], e);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[st]);
    }
    const uint32_t row = pix_min.y + 2 * warp + (pg >> 1);
    const uint32_t col0 = pix_min.x + (pg & 1) * 8;
    if (acc_nonfinite<MCH>(acc)) {
        forward_redo_guarded<MCH, true>(pool, n, dbase, features, C, ch0, nch, warp, woff, cg, bg_color, final_T, W, H,
                                        row, col0, out_color);
        return;
    }
    if (row < (uint32_t)H) {
        const size_t plane = (size_t)H * W;
        const bool vec = ((W & 3) == 0) && (col0 + 8 <= (uint32_t)W);
        const float4 t0 = *reinterpret_cast<const float4*>(&Tsm[(2 * warp + (pg >> 1)) * SGB_TILE + (pg & 1) * 8]);
        const float4 t1 = *reinterpret_cast<const float4*>(&Tsm[(2 * warp + (pg >> 1)) * SGB_TILE + (pg & 1) * 8 + 4]);
        const float Tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int k = 0; k < MCH; k++) {
            const int chl = (k >> 2) * 32 + cg * 4 + (k & 3);  // see `entry`: lane cg owns channels cg*4.. and 32+cg*4..
            if (chl >= nch) continue;
            const float bgc = bgS[chl];
            float* dst = out_color + (size_t)(ch0 + chl) * plane + (size_t)W * row + col0;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = ((k & 1) ? acc[i][k / 2].y : acc[i][k / 2].x) + Tv[i] * bgc;
            if (vec) {
                reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (col0 + i < (uint32_t)W) dst[i] = o[i];
            }
        }
    }
}

// dF[entry][ch] = sum over the tile's 256 pixels of w[entry][px] * dL/dout[px][ch]   (K = pixels).
// CTA = (tile, 64-channel chunk).  The dL tile is copied once into shared memory in its native
// [channel][pixel] order by cp.async row pieces (no transposing stores); warp w owns entries
// 16w..16w+15 of each 128-entry pass and streams their weight rows through a private double-buffered
// slab [16][32 px].  Lane = (eg, cg) accumulates entries {eg + 4j} x channels {4cg..4cg+3} U {32+4cg..32+4cg+3}; one K
// step covers 4 pixels with LDS.128 of both operands, and the packed FMAs pair (even, odd) pixels — no register
// duplication, the two halves are added at the end.
// Round 2 (ncu + tools/lds_probe.cu): the kernel was bound by the LSU, not the FMA pipe — 32 scalar red.global per
// lane and pass (~30-40 LSU cycles each) on top of 4-wavefront operand loads.  Now
//   * a lane owns two blocks of 4 CONSECUTIVE channels, so a Gaussian's sums leave as two red.global.add.v4.f32
//     (8 reductions per lane and pass instead of 32);
//   * the 16-byte pixel quads of channel row r sit at quad ^ ((r >> 2) & 7): the 8 channel groups of one load then hit 8
//     different bank groups although their rows are 4 apart (pitch 256 floats, no padding);
//   * eg / cg come from lane_group4 / lane_group8: every operand load is a 2-wavefront LDS.128.
template <int CH>
__global__ void __launch_bounds__(kThreads, 2) dfeature_gemm_kernel(int W, int H, int C,
                                                                   const float* __restrict__ dL_dpixels,
                                                                   PoolView pool, float* __restrict__ dL_dcolors) {
    static_assert(CH == 64, "64-channel chunks");
    constexpr int DP = SGB_TILE_PIX;  // pitch of a channel row of the dL tile (quads XOR-swizzled, see above)
    constexpr int WP = 36;            // pitch of the per-warp weight slab rows (32 px + pad)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float(*dLs)[DP] = reinterpret_cast<float(*)[DP]>(smem_raw);
    float* wslab = reinterpret_cast<float*>(smem_raw + sizeof(float) * CH * DP);
    __shared__ const float* Wrow[128];
    __shared__ uint32_t Gid[128];

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int nchunksC = (C + CH - 1) / CH;
    const int tile = blockIdx.x / nchunksC;
    const int ch0 = (blockIdx.x % nchunksC) * CH;
    const int nch = min(CH, C - ch0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = pool.count[tile];
    if (n == 0) return;  // (empty tiles are common: returning before the 64 KB dL copy matters)
    const uint32_t dbase = __ldg(pool.dirbase + tile);
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};
    const size_t plane = (size_t)H * W;
    const bool rows16 = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(dL_dpixels) & 15) == 0;
    // dL tile -> smem [ch][quad ^ swizzle][4 px]: 16-byte pieces (4 pixels of one tile row of one channel)
    for (int idx = tid; idx < CH * SGB_TILE * 4; idx += kThreads) {
        const int pc = idx & 3, r = (idx >> 2) & (SGB_TILE - 1), c = idx >> 6;
        const uint32_t y = pix_min.y + r, x = pix_min.x + pc * 4;
        float* dst = &dLs[c][((r * 4 + pc) ^ ((c >> 2) & 7)) * 4];
        const float* src = dL_dpixels + (size_t)(ch0 + c) * plane + (size_t)W * y + x;
        const bool rowok = c < nch && y < (uint32_t)H;
        if (rows16) {
            const bool ok = rowok && x + 4 <= (uint32_t)W;
            cp_async16(dst, ok ? src : dL_dpixels, ok ? 16 : 0);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) dst[i] = (rowok && x + i < (uint32_t)W) ? __ldg(src + i) : 0.f;
        }
    }
    cp_async_commit();

    const int eg = lane_group4(lane), cg = lane_group8(lane);
    const bool red16 = ((C & 3) == 0) && ((reinterpret_cast<uintptr_t>(dL_dcolors) & 15) == 0);
    for (uint32_t base = 0; base < n; base += 128) {
        const int cnt = (int)min(128u, n - base);
        __syncthreads();  // previous pass done with Wrow / Gid
        if (tid < cnt) {
            const WChunk* ck = pool.chunks + chunk_of(pool, dbase, (int)((base + tid) / kChunkEntries));
            const int s = (base + tid) & (kChunkEntries - 1);
            Wrow[tid] = &ck->w[s][0];
            Gid[tid] = ck->meta[s].x;
        }
        cp_async_wait<0>();
        __syncthreads();  // Wrow / Gid visible; dL tile landed (first pass)
        if (warp * 16 < cnt) {
            float(*wsl)[16][WP] = reinterpret_cast<float(*)[16][WP]>(wslab + (size_t)warp * 2 * 16 * WP);
            const float* lrow[4];
#pragma unroll
            for (int i = 0; i < 4; i++) lrow[i] = Wrow[min(warp * 16 + (lane >> 3) + 4 * i, cnt - 1)] + (lane & 7) * 4;
            auto issue = [&](int sl, int buf) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    cp_async16(&wsl[buf][(lane >> 3) + 4 * i][(lane & 7) * 4], lrow[i] + sl * 32, 16);
                cp_async_commit();
            };
            float2 acc[4][8];  // [entry eg+4j][channel (k >> 2) * 32 + 4 cg + (k & 3)], .x even pixels, .y odd pixels
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int k = 0; k < 8; k++) acc[j][k] = make_float2(0.f, 0.f);
            issue(0, 0);
            for (int sl = 0; sl < SGB_TILE_PIX / 32; sl++) {
                const int buf = sl & 1;
                if (sl + 1 < SGB_TILE_PIX / 32) { issue(sl + 1, buf ^ 1); cp_async_wait<1>(); }
                else cp_async_wait<0>();
                __syncwarp();
                // Fully unrolled and software-pipelined by hand: ptxas otherwise issues every LDS right in
                // front of its first consumer (ncu: half of all stall samples were short-scoreboard waits on
                // those FFMA2s).  dL rows are fetched two K-steps ahead, the next pixel quad's weights while
                // the current quad is being consumed.
                const float* drow = &dLs[4 * cg][0];
                const float* wbase = &wsl[buf][eg][0];
                auto ld_d = [&](int step) {  // step = p4 * 8 + k; row (k >> 2) * 32 + 4 cg + (k & 3), quad sl * 8 + p4
                    const int k = step & 7, p4 = step >> 3;
                    return *reinterpret_cast<const float4*>(drow + ((k >> 2) * 32 + (k & 3)) * DP +
                                                            (((sl * 8 + p4) ^ cg) << 2));
                };
                float4 wq[4], wn[4];
#pragma unroll
                for (int j = 0; j < 4; j++) wq[j] = *reinterpret_cast<const float4*>(wbase + 4 * j * WP);
                float4 d0 = ld_d(0), d1 = ld_d(1);
#pragma unroll
                for (int p4 = 0; p4 < 8; p4++) {
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const int step = p4 * 8 + k;
                        float4 d2 = d1;
                        if (step + 2 < 64) d2 = ld_d(step + 2);
                        if (k == 2 && p4 + 1 < 8) {
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                wn[j] = *reinterpret_cast<const float4*>(wbase + 4 * j * WP + (p4 + 1) * 4);
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            acc[j][k] = ffma2(make_float2(wq[j].x, wq[j].y), make_float2(d0.x, d0.y), acc[j][k]);
                            acc[j][k] = ffma2(make_float2(wq[j].z, wq[j].w), make_float2(d0.z, d0.w), acc[j][k]);
                        }
                        d0 = d1;
                        d1 = d2;
                    }
                    if (p4 + 1 < 8) {
#pragma unroll
                        for (int j = 0; j < 4; j++) wq[j] = wn[j];
                    }
                }
                __syncwarp();  // slab `buf` may be refilled by the next issue
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int e = warp * 16 + eg + 4 * j;
                if (e < cnt) {
                    float* dst = dL_dcolors + (size_t)Gid[e] * C + ch0 + 4 * cg;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int chl = h * 32 + 4 * cg;
                        const float4 v = make_float4(acc[j][4 * h + 0].x + acc[j][4 * h + 0].y, acc[j][4 * h + 1].x + acc[j][4 * h + 1].y,
                                                     acc[j][4 * h + 2].x + acc[j][4 * h + 2].y, acc[j][4 * h + 3].x + acc[j][4 * h + 3].y);
                        if (red16 && chl + 4 <= nch) {
                            red_add_v4_f32(dst + h * 32, v);
                        } else {
                            if (chl + 0 < nch) red_add_f32(dst + h * 32 + 0, v.x);
                            if (chl + 1 < nch) red_add_f32(dst + h * 32 + 1, v.y);
                            if (chl + 2 < nch) red_add_f32(dst + h * 32 + 2, v.z);
                            if (chl + 3 < nch) red_add_f32(dst + h * 32 + 3, v.w);
                        }
                    }
                }
            }
        }
    }
}

// Backward chain with the s-pass as a GEMM over channels.  CTA = tile; segments of 64 entries walked
// from the back of the list; per segment S[32 px][64 entries] per warp accumulates in registers over
// all channels (lane tile 8 px x 8 entries), is parked in shared memory, and lane = pixel then runs
// the reference's back-to-front chain over the segment.
template <bool VEC>
__global__ void __launch_bounds__(kThreads, 2) chain_backward_gemm_kernel(
    int W, int H, int C, const float* __restrict__ bg_color, const SplatRec* __restrict__ rec,
    const float* __restrict__ features, const float* __restrict__ final_Ts, const float* __restrict__ dL_dpixels,
    PoolView pool, float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic2D, float* __restrict__ dL_dopacity) {
    constexpr int CK = 16;         // channels per staged slab
    constexpr int FP = kSeg + 4;   // pitch of the transposed feature slab [ch][entry]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float(*S)[kSeg][32] = reinterpret_cast<float(*)[kSeg][32]>(smem_raw);                       // 64 KB
    float(*FT)[CK][FP] = reinterpret_cast<float(*)[CK][FP]>(smem_raw + sizeof(float) * 8 * kSeg * 32);  // 2 slabs
    float(*DS)[2][CK][32] = reinterpret_cast<float(*)[2][CK][32]>(
        smem_raw + sizeof(float) * (8 * kSeg * 32 + 2 * CK * FP));  // per-warp dL slabs, 2 x 2 KB each
    __shared__ const float* Wrow[kSeg];
    __shared__ uint2 Meta[kSeg];
    __shared__ float4 RecA[kSeg], RecB[kSeg];

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pg = lane >> 3, eg = lane & 7;
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};
    const uint32_t tx = tid & (SGB_TILE - 1), ty = tid >> 4;
    const uint2 pix = {pix_min.x + tx, pix_min.y + ty};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    const uint32_t n = pool.count[tile];
    if (n == 0) return;
    const uint32_t dbase = pool.dirbase[tile];

    const size_t plane = (size_t)H * W;
    const bool rows16 = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(dL_dpixels) & 15) == 0);
    const int woff = warp * 32 + lane;

    // background term of the own pixel over all channels (backward.cu:527-529).  With an all-zero background
    // (the usual case) the term vanishes; skipping it saves a second full read of dL/dout (ncu: 4.6 GB read
    // by this kernel against 2.1 GB of dL/dout).
    int bg_nonzero = 0;
    for (int ch = tid; ch < C; ch += kThreads) bg_nonzero |= (bg_color[ch] != 0.f);
    bg_nonzero = __syncthreads_or(bg_nonzero);
    float bgdot = 0.f;
    if (inside && bg_nonzero)
        for (int ch = 0; ch < C; ch++) bgdot += bg_color[ch] * __ldg(dL_dpixels + (size_t)ch * plane + pix_id);

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    float T = T_final;
    float last_alpha = 0.f, s_last = 0.f, A = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    // segments are aligned to 64 entries from the FRONT so that a segment is exactly 4 whole chunks
    // (the last one may be partial); they are visited back to front.
    const int nseg = (int)((n + kSeg - 1) / kSeg);
    for (int sg = nseg - 1; sg >= 0; sg--) {
        const int base = sg * kSeg;
        const int cnt = min(kSeg, (int)n - base);
        __syncthreads();  // previous segment done with S / Wrow / Meta / FT
        if (tid < kSeg) {
            if (tid < cnt) {
                const WChunk* ck = pool.chunks + chunk_of(pool, dbase, (base + tid) / kChunkEntries);
                const int s = (base + tid) & (kChunkEntries - 1);
                Wrow[tid] = &ck->w[s][0];
                const uint2 mt = ck->meta[s];
                Meta[tid] = mt;
                const float4* rp = reinterpret_cast<const float4*>(rec + mt.x);
                RecA[tid] = __ldg(rp);
                RecB[tid] = __ldg(rp + 1);
            } else {
                Meta[tid] = make_uint2(0u, 0u);
            }
        }
        __syncthreads();

        float2 acc[8][4];  // [px][entry pair]
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);

        const int nslab = (C + CK - 1) / CK;
        // feature slab: thread t -> entry t>>2, 4 channels (t&3)*4.. of the slab; loaded into registers
        // one slab ahead, stored transposed [ch][entry] after the current slab's math
        float4 fpre;
        auto fload = [&](int sl) {
            const int e = tid >> 2, q = tid & 3;
            const int chb = sl * CK + q * 4;
            fpre = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < cnt) {
                const float* src = features + (size_t)Meta[e].x * C + chb;
                if (VEC && chb + 4 <= C) fpre = __ldg(reinterpret_cast<const float4*>(src));
                else {
                    if (chb < C) fpre.x = __ldg(src);
                    if (chb + 1 < C) fpre.y = __ldg(src + 1);
                    if (chb + 2 < C) fpre.z = __ldg(src + 2);
                    if (chb + 3 < C) fpre.w = __ldg(src + 3);
                }
            }
        };
        // An FT row holds the 64 entries of the segment PERMUTED: entries 8g..8g+3 of every entry group g first
        // (32 floats), then entries 8g+4..8g+7 — so that each of a lane's two LDS.128 reads 8 x 16 B contiguous
        // across the 8 entry-group lanes (one wavefront instead of two).
        auto fstore = [&](int buf) {
            const int e = tid >> 2, q = tid & 3;
            const int pos = ((e >> 3) << 2) | (e & 3) | (((e >> 2) & 1) << 5);
            FT[buf][q * 4 + 0][pos] = fpre.x;
            FT[buf][q * 4 + 1][pos] = fpre.y;
            FT[buf][q * 4 + 2][pos] = fpre.z;
            FT[buf][q * 4 + 3][pos] = fpre.w;
        };
        // dL slab [CK ch][32 px of this warp]: 4 x cp.async(16 B) per lane, private to the warp
        auto dissue = [&](int sl, int buf) {
#pragma unroll
            for (int i = 0; i < CK / 4; i++) {
                const int chl = (lane >> 3) + 4 * i;
                const int ch = sl * CK + chl;
                const int pc = lane & 7;  // 16-byte piece: tile row pc>>2 of the strip, columns (pc&3)*4..
                const uint32_t y = pix_min.y + 2 * warp + (pc >> 2);
                const uint32_t x = pix_min.x + (pc & 3) * 4;
                const bool rowin = ch < C && y < (uint32_t)H;
                const float* src = dL_dpixels + (size_t)ch * plane + (size_t)W * y + x;
                if (rows16) {
                    const bool ok = rowin && x + 4 <= (uint32_t)W;
                    cp_async16(&DS[warp][buf][chl][pc * 4], ok ? src : dL_dpixels, ok ? 16 : 0);
                } else {  // image rows not 16-byte aligned: plain loads, ordered by the slab barrier
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        DS[warp][buf][chl][pc * 4 + u] = (rowin && x + u < (uint32_t)W) ? __ldg(src + u) : 0.f;
                }
            }
            cp_async_commit();
        };
        fload(0);
        dissue(0, 0);
        fstore(0);
        if (nslab > 1) fload(1);
        for (int sl = 0; sl < nslab; sl++) {
            const int buf = sl & 1;
            if (sl + 1 < nslab) { dissue(sl + 1, buf ^ 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();  // FT[buf] stored by everyone, this warp's dL slab landed; slab sl-1 consumed
#pragma unroll 4
            for (int k = 0; k < CK; k++) {
                const float4 d0 = *reinterpret_cast<const float4*>(&DS[warp][buf][k][pg * 8]);
                const float4 d1 = *reinterpret_cast<const float4*>(&DS[warp][buf][k][pg * 8 + 4]);
                const float4 f0 = *reinterpret_cast<const float4*>(&FT[buf][k][eg * 4]);        // entries 8eg..8eg+3
                const float4 f1 = *reinterpret_cast<const float4*>(&FT[buf][k][32 + eg * 4]);   // entries 8eg+4..8eg+7
                const float2 ff[4] = {make_float2(f0.x, f0.y), make_float2(f0.z, f0.w), make_float2(f1.x, f1.y),
                                      make_float2(f1.z, f1.w)};
                const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float2 d2 = make_float2(d[i], d[i]);
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = ffma2(ff[j], d2, acc[i][j]);
                }
            }
            if (sl + 1 < nslab) {
                fstore(buf ^ 1);  // readers of FT[buf^1] (slab sl-1) all passed this iteration's barrier
                if (sl + 2 < nslab) fload(sl + 2);
            }
            __syncwarp();  // every lane is done with DS[warp][buf] before the next cp.async refills it
        }
        // park S: S[warp][entry][px]
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float* dst = &S[warp][eg * 8 + j][pg * 8];
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = (j & 1) ? acc[i][j / 2].y : acc[i][j / 2].x;
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncwarp();

        // back-to-front chain over the segment (backward.cu:477-550 in dot-product form)
        // own-pixel weights are prefetched two entries ahead (the only global load left in this loop)
        float wn0 = __ldg(Wrow[cnt - 1] + woff);
        float wn1 = cnt > 1 ? __ldg(Wrow[cnt - 2] + woff) : 0.f;
        for (int li = cnt - 1; li >= 0; li--) {
            const float w = wn0;
            wn0 = wn1;
            if (li >= 2) wn1 = __ldg(Wrow[li - 2] + woff);
            const uint2 meta = Meta[li];
            if (!((meta.y >> warp) & 1u)) continue;
            const float sdot = S[warp][li][lane];
            const float4 a = RecA[li], con_o = RecB[li];
            float gv[8];
#pragma unroll
            for (int v = 0; v < 8; v++) gv[v] = 0.f;
            if (w != 0.f) {
                const float2 d = {a.x - pixf.x, a.y - pixf.y};
                const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
                const float G = exp(power);
                const float alpha = min(0.99f, con_o.w * G);
                T = T / (1.f - alpha);
                A = last_alpha * s_last + (1.f - last_alpha) * A;
                s_last = sdot;
                float dL_dalpha = (sdot - A) * T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
                const float dL_dG = con_o.w * dL_dalpha;
                const float gdx = G * d.x, gdy = G * d.y;
                const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
                const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;
                gv[0] = dL_dG * dG_ddelx * ddelx_dx;
                gv[1] = dL_dG * dG_ddely * ddely_dy;
                gv[2] = -0.5f * gdx * d.x * dL_dG;
                gv[3] = -0.5f * gdx * d.y * dL_dG;
                gv[4] = -0.5f * gdy * d.y * dL_dG;
                gv[5] = G * dL_dalpha;
            }
            xreduce_step<8>(gv, lane, 4);
            xreduce_step<4>(gv, lane, 2);
            xreduce_step<2>(gv, lane, 1);
            float gq = gv[0];
            gq += __shfl_xor_sync(0xffffffffu, gq, 8);
            gq += __shfl_xor_sync(0xffffffffu, gq, 16);
            if (lane < 6) {
                const size_t id = meta.x;
                float* dst = lane < 2 ? dL_dmean2D + id * 3 + lane
                           : lane < 5 ? dL_dconic2D + id * 4 + (lane == 4 ? 3 : lane - 2)
                                      : dL_dopacity + id;
                red_add_f32(dst, gq);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ warp-autonomous chain backward
// Second generation of the chain backward (round 2).  ncu on the CTA-synchronous kernel above (K3): FMA pipe 50 %,
// issue 49 %, 16 warps/SM; executed FMAs = 2.0 x the algorithmic ones — every warp multiplied all 64 slots of every
// segment of the TILE list although its 32-pixel strip is touched by ~85 % of the tile's entries, and segments of 64
// padded the list by another ~22 %; one CTA-wide barrier per 16-channel slab.  Here every warp owns its strip end to
// end and nothing is CTA-synchronous after the prologue:
//   * the warp walks the tile list from the back and COMPACTS it on the fly to the entries whose strip-mask bit is
//     set (ballot + popc ranks), 32 entries per segment: no zero-strip work, padding
//     <= 31 slots per strip instead of <= 63 per tile;
//   * s-pass per segment: S[32 px][32 entries] over all channels, lane tile 8 px x 4 entries (16 packed FMAs per 3
//     LDS.128); the warp stages its own operands — the dL/dout slab [16 ch][32 px] by cp.async, the feature slab by
//     4 x LDG.128 per lane (lane = entry) one slab ahead in registers, stored transposed [ch][entry];
//   * S is parked in the warp's dL slab region (XOR-swizzled 16-byte chunks: conflict-free both ways) and lane = pixel
//     runs the reference's back-to-front chain (backward.cu:477-550, dot-product form) over the 32 entries.
// Shared memory 9.3 KB per warp, 128 registers, 2 CTAs/SM (a 80-register / 3-CTA build measured 13 % SLOWER on K3:
// 5.22 vs 4.62 ms — the extra warps thrash the 28 KB of L1 that three CTAs leave for the gathered rows); the warps of
// a tile share their feature rows through L1/L2 only.
constexpr int kChainRG = 5;  // entries whose six gradient terms are summed over the strip per flush (30 of 32 lanes busy)
struct __align__(16) ChainWarpSmem {
    float DS[2][16][32];   // dL/dout slabs [buf][ch][px of the strip]; S[32 entries][32 px] aliases it after the s-pass
    union {
        float FT[2][16][36];           // s-pass: feature slabs [buf][ch][entry]
        float RB[kChainRG * 6][32];    // chain phase: partial gradient terms, one row per (entry of the group, term)
    };
    float4 RecA[32], RecB[32];
    const float* Wrow[32];
    uint32_t Gid[32];
};
constexpr int kMetaCap = 512;  // tile-list entries whose (id, mask) records are cached in shared memory

template <bool VEC>
__global__ void __launch_bounds__(kThreads, 2) chain_backward_warp_kernel(
    int W, int H, int C, const float* __restrict__ bg_color, const SplatRec* __restrict__ rec,
    const float* __restrict__ features, const float* __restrict__ final_Ts, const float* __restrict__ dL_dpixels,
    PoolView pool, float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic2D, float* __restrict__ dL_dopacity,
    const __grid_constant__ CUtensorMap dl_map, const int use_tma) {
    constexpr int CK = 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint2 MetaS[kMetaCap];
    __shared__ uint64_t dbar[kThreads / 32][2];  // per warp, per dL slab buffer: TMA completion
    __shared__ uint32_t Cdir[kMetaCap / kChunkEntries];

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pg = lane_group4(lane), eg = lane_group8(lane);
    ChainWarpSmem& ws = reinterpret_cast<ChainWarpSmem*>(smem_raw)[warp];
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};
    const uint32_t tx = tid & (SGB_TILE - 1), ty = tid >> 4;
    const uint2 pix = {pix_min.x + tx, pix_min.y + ty};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    const uint32_t n = pool.count[tile];
    if (n == 0) return;
    const uint32_t dbase = pool.dirbase[tile];
    const size_t plane = (size_t)H * W;
    const bool rows16 = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(dL_dpixels) & 15) == 0);
    const int woff = warp * 32 + lane;

    // ---- CTA prologue: directory + (id, mask) records of the tile list -> shared memory; background flag
    const uint32_t ncache = min(n, (uint32_t)kMetaCap);
    for (uint32_t k = tid; k * kChunkEntries < ncache; k += kThreads) Cdir[k] = chunk_of(pool, dbase, (int)k);
    if (lane == 0) {
        mbar_init(&dbar[warp][0], 1);
        mbar_init(&dbar[warp][1], 1);
        mbar_fence_init();
    }
    int bg_nonzero = 0;
    for (int ch = tid; ch < C; ch += kThreads) bg_nonzero |= (bg_color[ch] != 0.f);
    bg_nonzero = __syncthreads_or(bg_nonzero);   // also orders the Cdir stores and the barrier inits
    for (uint32_t e = tid; e < ncache; e += kThreads)
        MetaS[e] = __ldg(&pool.chunks[Cdir[e / kChunkEntries]].meta[e & (kChunkEntries - 1)]);
    __syncthreads();
    auto chunk_ptr = [&](uint32_t e) -> const WChunk* {
        return pool.chunks + (e < ncache ? Cdir[e / kChunkEntries] : chunk_of(pool, dbase, (int)(e / kChunkEntries)));
    };
    auto meta_of = [&](uint32_t e) -> uint2 {
        return e < ncache ? MetaS[e] : __ldg(&chunk_ptr(e)->meta[e & (kChunkEntries - 1)]);
    };

    // background term of the own pixel over all channels (backward.cu:527-529); zero background: term vanishes
    float bgdot = 0.f;
    if (inside && bg_nonzero)
        for (int ch = 0; ch < C; ch++) bgdot += bg_color[ch] * __ldg(dL_dpixels + (size_t)ch * plane + pix_id);

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    float T = T_final;
    float last_alpha = 0.f, s_last = 0.f, A = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const int nslab = (C + CK - 1) / CK;
    float (*S)[32] = reinterpret_cast<float (*)[32]>(&ws.DS[0][0][0]);

    // dL slab [CK ch][32 px of this strip]: 4 x cp.async(16 B) per lane; lane -> (channel row (lane >> 3) + 4 i of the
    // slab, 16-byte piece lane & 7: tile row pc >> 2 of the strip, columns (pc & 3) * 4 ..)
    const int d_pc = lane & 7;
    const uint32_t d_y = pix_min.y + 2 * warp + (d_pc >> 2), d_x = pix_min.x + (d_pc & 3) * 4;
    const bool d_rowin = d_y < (uint32_t)H;
    const bool d_vec_ok = d_rowin && d_x + 4 <= (uint32_t)W;
    const float* d_src0 = dL_dpixels + (size_t)(lane >> 3) * plane + (size_t)W * d_y + d_x;
    // Fast path (16-byte aligned image rows, full slab): the lane keeps a running source pointer; a lane whose piece is
    // outside the image points at the image base with zero strides and copies 0 bytes (cp.async zero-fills).
    const char* const d_base = reinterpret_cast<const char*>(d_vec_ok ? d_src0 : dL_dpixels);
    const size_t d_step4 = d_vec_ok ? 4 * plane * sizeof(float) : 0;       // 4 channel rows further
    const size_t d_stepslab = d_vec_ok ? CK * plane * sizeof(float) : 0;   // next slab
    const int d_bytes = d_vec_ok ? 16 : 0;
    const char* d_run = d_base;
    // TMA path (image rows 16-byte aligned; the map is encoded per launch by the host): ONE instruction of one lane
    // fetches the whole [16 ch][2 rows][16 px] box — rows below the image, columns right of it and channels >= C arrive
    // as zeros — and none of it passes through the LSU data pipe (the four LDGSTS per lane it replaces were 64 of the
    // ~210 L1 wavefronts per slab, and the L1 data pipe bounds this kernel).
    uint32_t dphase = 0;  // bit b: parity the next wait on buffer b expects
    auto dissue = [&](int sl, int buf) {
        if (use_tma) {
            if (lane == 0) {
                fence_proxy_async_smem();  // S of the previous segment was written to this memory by generic stores
                mbar_arrive_expect_tx(&dbar[warp][buf], CK * 32 * sizeof(float));
                tma_tile3d_g2s(&ws.DS[buf][0][0], &dl_map, (int)pix_min.x, (int)pix_min.y + 2 * warp, sl * CK,
                               &dbar[warp][buf]);
            }
            return;
        }
        if (rows16 && (sl + 1) * CK <= C) {
            float* dst = &ws.DS[buf][lane >> 3][d_pc * 4];
            const char* src = d_run;
#pragma unroll
            for (int i = 0; i < CK / 4; i++) {
                cp_async16(dst + i * 4 * 32, src, d_bytes);
                src += d_step4;
            }
        } else {
            const float* srcs = d_src0 + (size_t)sl * CK * plane;
#pragma unroll
            for (int i = 0; i < CK / 4; i++) {
                const int chl = (lane >> 3) + 4 * i;
                const bool chin = sl * CK + chl < C;
                const float* src = srcs + (size_t)(4 * i) * plane;
                if (rows16) {
                    const bool ok = chin && d_vec_ok;
                    cp_async16(&ws.DS[buf][chl][d_pc * 4], ok ? src : dL_dpixels, ok ? 16 : 0);
                } else {  // image rows not 16-byte aligned: plain loads, ordered by the warp barrier of the slab loop
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        ws.DS[buf][chl][d_pc * 4 + u] = (chin && d_rowin && d_x + u < (uint32_t)W) ? __ldg(src + u) : 0.f;
                }
            }
        }
        d_run += d_stepslab;
        cp_async_commit();
    };

    uint32_t cursor = n;  // tile-list entries [0, cursor) are still to be visited (back to front)
    while (cursor > 0) {
        // ---- gather the next <= 32 entries of THIS strip, walking the tile list backwards.  Slot 0 = furthest back.
        int cnt = 0;
        while (cnt < 32 && cursor > 0) {
            const bool valid = (uint32_t)lane < cursor;
            const uint32_t e = valid ? cursor - 1 - (uint32_t)lane : 0u;
            const uint2 mt = valid ? meta_of(e) : make_uint2(0u, 0u);
            const bool bit = valid && ((mt.y >> warp) & 1u);
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            const int room = 32 - cnt;
            const int nset = __popc(bal);
            const int rank = __popc(bal & ((1u << lane) - 1u));
            if (bit && rank < room) {
                const int slot = cnt + rank;
                const WChunk* ck = chunk_ptr(e);
                ws.Wrow[slot] = &ck->w[e & (kChunkEntries - 1)][0];
                ws.Gid[slot] = mt.x;
                const float4* rp = reinterpret_cast<const float4*>(rec + mt.x);
                ws.RecA[slot] = __ldg(rp);
                ws.RecB[slot] = __ldg(rp + 1);
            }
            if (nset <= room) {
                cursor -= min(32u, cursor);
                cnt += nset;
            } else {  // segment full: resume right after the last entry taken
                const int last_lane = __ffs(__ballot_sync(0xffffffffu, bit && rank == room - 1)) - 1;
                cursor -= (uint32_t)(last_lane + 1);
                cnt = 32;
            }
        }
        __syncwarp();
        if (cnt == 0) break;

        // ---- s-pass: S[px][entry] = sum_ch dL[px][ch] * F[entry][ch]
        float2 acc[8][2];  // [px][entry pair]
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
        // Feature slab [32 entries][16 ch] -> FT[ch][entry].  Fast path (16-byte aligned rows, full slab): lane
        // (r8 = lane >> 2, c4 = lane & 3) loads channels 4c4..4c4+3 of entries r8 + 8q, so one LDG.128 covers 8 rows x
        // 64 contiguous bytes (8 L1 tag lookups; lane = entry touched ~27 lines per instruction and the feature gather
        // alone was a third of the kernel's L1 wavefronts, ncu r02).  FT rows of channels 8..15 hold their 8-entry
        // blocks swapped pairwise (block b at b ^ 1) — with the 36-float pitch that makes the transposing stores of
        // this mapping conflict-free; the s-pass reads entry group eg of channel k at chunk eg ^ ((k >> 3) << 1).
        float4 fpre[CK / 4];
        const int f_r8 = lane >> 2, f_c4 = lane & 3;
        const float* frow = features + (size_t)ws.Gid[min(lane, cnt - 1)] * C;          // general path: lane = entry
        uint32_t f_gid[CK / 4];                                                            // fast path: 4 rows per lane
#pragma unroll
        for (int q = 0; q < CK / 4; q++) f_gid[q] = ws.Gid[min(f_r8 + 8 * q, cnt - 1)];    // rows >= cnt: duplicates
        auto slab_fast = [&](int sl) { return VEC && (sl + 1) * CK <= C; };
        auto fload = [&](int sl) {
            if (slab_fast(sl)) {
                const int choff = sl * CK + f_c4 * 4;
#pragma unroll
                for (int q = 0; q < CK / 4; q++)
                    fpre[q] = __ldg(reinterpret_cast<const float4*>(features + (size_t)f_gid[q] * C + choff));
                return;
            }
#pragma unroll
            for (int q = 0; q < CK / 4; q++) {
                const int chb = sl * CK + q * 4;
                fpre[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < cnt) {
                    if (VEC && chb + 4 <= C) fpre[q] = __ldg(reinterpret_cast<const float4*>(frow + chb));
                    else {
                        if (chb < C) fpre[q].x = __ldg(frow + chb);
                        if (chb + 1 < C) fpre[q].y = __ldg(frow + chb + 1);
                        if (chb + 2 < C) fpre[q].z = __ldg(frow + chb + 2);
                        if (chb + 3 < C) fpre[q].w = __ldg(frow + chb + 3);
                    }
                }
            }
        };
        const int f_swap = (f_c4 >> 1) * 8;  // channels 8..15: 8-entry blocks swapped pairwise
        auto fstore = [&](int buf, int sl) {
            if (slab_fast(sl)) {
#pragma unroll
                for (int q = 0; q < CK / 4; q++) {
                    float* col = &ws.FT[buf][f_c4 * 4][8 * q + f_r8 + ((q & 1) ? -f_swap : f_swap)];
                    col[0 * 36] = fpre[q].x;
                    col[1 * 36] = fpre[q].y;
                    col[2 * 36] = fpre[q].z;
                    col[3 * 36] = fpre[q].w;
                }
                return;
            }
#pragma unroll
            for (int q = 0; q < CK / 4; q++) {  // lane = entry, channels 4q..4q+3
                const int pos = lane ^ ((q >> 1) << 3);
                ws.FT[buf][q * 4 + 0][pos] = fpre[q].x;
                ws.FT[buf][q * 4 + 1][pos] = fpre[q].y;
                ws.FT[buf][q * 4 + 2][pos] = fpre[q].z;
                ws.FT[buf][q * 4 + 3][pos] = fpre[q].w;
            }
        };
        // One warp barrier per slab: at the top of iteration sl every lane has finished the math of slab sl-1, so the
        // other buffers (dL by cp.async, features from the registers loaded one slab earlier) can be refilled BEFORE
        // the math of slab sl and their latency hides behind it.
        fload(0);
        d_run = d_base;
        dissue(0, 0);
        fstore(0, 0);
        if (nslab > 1) fload(1);
        for (int sl = 0; sl < nslab; sl++) {
            const int buf = sl & 1;
            if (use_tma) {
                mbar_wait(&dbar[warp][buf], (dphase >> buf) & 1u);
                dphase ^= 1u << buf;
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();  // DS[buf] landed, FT[buf] stored by every lane; DS/FT[buf ^ 1] are free
            if (sl + 1 < nslab) {
                dissue(sl + 1, buf ^ 1);
                fstore(buf ^ 1, sl + 1);
                if (sl + 2 < nslab) fload(sl + 2);
            }
#pragma unroll 8
            for (int k = 0; k < CK; k++) {
                const float4 d0 = *reinterpret_cast<const float4*>(&ws.DS[buf][k][pg * 8]);
                const float4 d1 = *reinterpret_cast<const float4*>(&ws.DS[buf][k][pg * 8 + 4]);
                const float4 f0 = *reinterpret_cast<const float4*>(&ws.FT[buf][k][(eg ^ ((k >> 3) << 1)) * 4]);
                const float2 fa = make_float2(f0.x, f0.y), fb = make_float2(f0.z, f0.w);
                const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float2 d2 = make_float2(d[i], d[i]);
                    acc[i][0] = ffma2(fa, d2, acc[i][0]);
                    acc[i][1] = ffma2(fb, d2, acc[i][1]);
                }
            }
        }
        __syncwarp();  // every lane is done with DS / FT
        // ---- park S[entry][px] in the (now free) dL slab region; 16-byte chunk c of row r sits at chunk c ^ (r >> 2)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = eg * 4 + j;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
            *reinterpret_cast<float4*>(&S[r][((2 * pg) ^ eg) * 4]) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(&S[r][((2 * pg + 1) ^ eg) * 4]) = make_float4(v[4], v[5], v[6], v[7]);
        }
        __syncwarp();

        // ---- back-to-front chain over the segment (backward.cu:477-550 in dot-product form); slot 0 is the
        // furthest-back entry.  Entries go in groups of kChainRG, fully unrolled (every shared-memory address of the
        // group is a constant plus a lane term — ncu r02: 40 of ~170 instructions per entry were address arithmetic
        // of the runtime-indexed version); the own-pixel weights of the next group are in flight during the current
        // one.  The six per-Gaussian sums over the strip's 32 pixels go through shared memory instead of a shuffle
        // butterfly: every lane parks its terms as rows of RB, then lane r adds up row r with 8 x LDS.128 and issues
        // that row's one red.global.  16-byte chunk c of row r sits at c ^ (r & 7): conflict-free both ways.
        // The transmittance in front of an entry is recovered as T_behind + w (w = alpha T_front is what the forward
        // stored): one add instead of the reference's T / (1 - alpha), same value to an ulp and no error build-up.
        constexpr int RG = kChainRG;
        float wc[RG], wn[RG];
#pragma unroll
        for (int u = 0; u < RG; u++) wc[u] = u < cnt ? __ldg(ws.Wrow[u] + woff) : 0.f;
        for (int base = 0; base < cnt; base += RG) {
#pragma unroll
            for (int u = 0; u < RG; u++) wn[u] = base + RG + u < cnt ? __ldg(ws.Wrow[base + RG + u] + woff) : 0.f;
#pragma unroll
            for (int u = 0; u < RG; u++) {
                const int li = base + u;
                if (li < cnt) {  // warp-uniform
                    // Branch-free per lane (selects instead of `if (w != 0)`): the five entries of a group then sit in
                    // one basic block and the scheduler overlaps their LDS -> exp -> product latencies.
                    const float w = wc[u];
                    const bool on = w != 0.f;
                    const float sdot = S[li][(((lane >> 2) ^ (li >> 2)) << 2) | (lane & 3)];
                    const float4 a = ws.RecA[li], con_o = ws.RecB[li];
                    const float2 d = {a.x - pixf.x, a.y - pixf.y};
                    const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
                    const float G = __expf(power);
                    const float alpha = fminf(0.99f, con_o.w * G);
                    T += w;
                    const float A_new = last_alpha * s_last + (1.f - last_alpha) * A;
                    A = on ? A_new : A;
                    s_last = on ? sdot : s_last;
                    last_alpha = on ? alpha : last_alpha;
                    float dL_dalpha = (sdot - A) * T;
                    if (bg_nonzero) dL_dalpha -= T_final / (1.f - alpha) * bgdot;
                    const float dL_dG = con_o.w * dL_dalpha;
                    const float gdx = G * d.x, gdy = G * d.y;
                    const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
                    const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;
                    float gv[6];
                    gv[0] = dL_dG * dG_ddelx * ddelx_dx;
                    gv[1] = dL_dG * dG_ddely * ddely_dy;
                    gv[2] = -0.5f * gdx * d.x * dL_dG;
                    gv[3] = -0.5f * gdx * d.y * dL_dG;
                    gv[4] = -0.5f * gdy * d.y * dL_dG;
                    gv[5] = G * dL_dalpha;
#pragma unroll
                    for (int v = 0; v < 6; v++) {
                        const int r = u * 6 + v;  // compile-time
                        ws.RB[r][(((lane >> 2) ^ (r & 7)) << 2) | (lane & 3)] = on ? gv[v] : 0.f;
                    }
                }
            }
            __syncwarp();
            const int nvalid = min(RG, cnt - base);
            if (lane < nvalid * 6) {
                const float4* row = reinterpret_cast<const float4*>(&ws.RB[lane][0]);
                float4 t = row[lane & 7];  // chunk 0 of row `lane`
#pragma unroll
                for (int q = 1; q < 8; q++) {
                    const float4 uu = row[q ^ (lane & 7)];
                    t.x += uu.x; t.y += uu.y; t.z += uu.z; t.w += uu.w;
                }
                const float tot = (t.x + t.y) + (t.z + t.w);
                const int slot = lane / 6, comp = lane - slot * 6;
                const size_t id = ws.Gid[base + slot];
                float* dst = comp < 2 ? dL_dmean2D + id * 3 + comp
                           : comp < 5 ? dL_dconic2D + id * 4 + (comp == 4 ? 3 : comp - 2)
                                      : dL_dopacity + id;
                red_add_f32(dst, tot);
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < RG; u++) wc[u] = wn[u];
        }
        __syncwarp();  // the next segment's gather / dL slab overwrite Gid, Wrow, Rec and S
    }
}

// ------------------------------------------------------------------------------------ warp-autonomous forward
// Forward counterpart of chain_backward_warp_kernel (round 2).  The TMA-ring kernel above multiplies every entry of the
// TILE list into every strip (ncu r01: 57 % FMA pipe, 13 % of the stalls on the per-CTA prologue, 35 % of the FMAs on
// zero weights).  Here a warp owns (strip, 32-channel slice): it compacts the tile list to the entries whose
// strip-mask bit is set, streams for each of them 128 B of weights (its strip only — the ring staged the whole 1 KB
// row for every 64-channel slice) and 128 B of features through a private cp.async double buffer (8 entries per
// stage) and accumulates an 8 px x 4 ch register tile per lane (16 packed FMAs per 3 LDS.128).  No CTA barrier after
// the prologue; 64 registers, 4.4 KB of shared memory per warp.  CTA = (tile, SL consecutive slices handled one after
// the other by every warp).  Accumulation order = list order, like forward.cu:355-356.
struct __align__(16) FwdWarpSmem {
    float Wst[2][8][32];   // [stage][entry][pixel of the strip]
    float Fst[2][8][32];   // [stage][entry][channel of the slice]
    const float* Wrow[32]; // queue of gathered entries: weight rows (already offset to this strip)
    uint32_t Gid[32];
};

template <bool VEC>
__global__ void __launch_bounds__(kThreads, 3) blend_forward_warp_kernel(
    int W, int H, int C, int SL, const float* __restrict__ features, const float* __restrict__ bg_color,
    const float* __restrict__ final_T, PoolView pool, float* __restrict__ out_color) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint2 MetaS[kMetaCap];
    __shared__ uint32_t Cdir[kMetaCap / kChunkEntries];

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int nslices = (C + 31) / 32;
    const int groups = (nslices + SL - 1) / SL;   // CTAs per tile
    const int tile = blockIdx.x / groups;
    const int slice0 = (blockIdx.x % groups) * SL;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pg = lane >> 3, cg = lane & 7;
    FwdWarpSmem& ws = reinterpret_cast<FwdWarpSmem*>(smem_raw)[warp];
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};
    const uint32_t n = pool.count[tile];
    const uint32_t dbase = pool.dirbase[tile];
    const size_t plane = (size_t)H * W;

    const uint32_t ncache = min(n, (uint32_t)kMetaCap);
    for (uint32_t k = tid; k * kChunkEntries < ncache; k += kThreads) Cdir[k] = chunk_of(pool, dbase, (int)k);
    __syncthreads();
    for (uint32_t e = tid; e < ncache; e += kThreads)
        MetaS[e] = __ldg(&pool.chunks[Cdir[e / kChunkEntries]].meta[e & (kChunkEntries - 1)]);
    __syncthreads();
    auto chunk_ptr = [&](uint32_t e) -> const WChunk* {
        return pool.chunks + (e < ncache ? Cdir[e / kChunkEntries] : chunk_of(pool, dbase, (int)(e / kChunkEntries)));
    };
    auto meta_of = [&](uint32_t e) -> uint2 {
        return e < ncache ? MetaS[e] : __ldg(&chunk_ptr(e)->meta[e & (kChunkEntries - 1)]);
    };

    // output pixels of this lane: 8 consecutive pixels of tile row 2*warp + (pg >> 1), columns (pg & 1) * 8 ..
    const uint32_t row = pix_min.y + 2 * warp + (pg >> 1);
    const uint32_t col0 = pix_min.x + (pg & 1) * 8;
    float Tv[8];
#pragma unroll
    for (int i = 0; i < 8; i++)
        Tv[i] = (row < (uint32_t)H && col0 + i < (uint32_t)W) ? final_T[(size_t)W * row + col0 + i] : 0.f;
    const int sp = lane & 7, se = lane >> 3;  // staging role: 16-byte piece sp of entries se and se + 4 of a stage

    for (int sli = 0; sli < SL && slice0 + sli < nslices; sli++) {
        const int ch0 = (slice0 + sli) * 32;
        const int nch = min(32, C - ch0);
        float2 acc[8][2];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);

        uint32_t cursor = 0;  // tile-list entries [cursor, n) are still to be visited (front to back)
        while (cursor < n) {
            // ---- gather the next <= 32 entries of this strip
            int cnt = 0;
            while (cnt < 32 && cursor < n) {
                const uint32_t e = cursor + (uint32_t)lane;
                const bool valid = e < n;
                const uint2 mt = valid ? meta_of(e) : make_uint2(0u, 0u);
                const bool bit = valid && ((mt.y >> warp) & 1u);
                const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                const int room = 32 - cnt;
                const int nset = __popc(bal);
                const int rank = __popc(bal & ((1u << lane) - 1u));
                if (bit && rank < room) {
                    const int slot = cnt + rank;
                    ws.Wrow[slot] = &chunk_ptr(e)->w[e & (kChunkEntries - 1)][warp * 32];
                    ws.Gid[slot] = mt.x;
                }
                if (nset <= room) {
                    cursor += min(32u, n - cursor);
                    cnt += nset;
                } else {
                    const int last_lane = __ffs(__ballot_sync(0xffffffffu, bit && rank == room - 1)) - 1;
                    cursor += (uint32_t)(last_lane + 1);
                    cnt = 32;
                }
            }
            __syncwarp();
            if (cnt == 0) break;
            const int nst = (cnt + 7) >> 3;
            auto issue = [&](int st) {  // stage st of the queue -> buffer st & 1 (missing entries / channels: zeros)
                const int buf = st & 1;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int el = se + 4 * h, q = st * 8 + el;
                    const bool ok = q < cnt;
                    const float* wsrc = ok ? ws.Wrow[q] + sp * 4 : pool.chunks->w[0];
                    cp_async16(&ws.Wst[buf][el][sp * 4], wsrc, ok ? 16 : 0);
                    if (VEC) {
                        const bool fok = ok && sp * 4 < nch;
                        const float* fsrc = fok ? features + (size_t)ws.Gid[q] * C + ch0 + sp * 4 : features;
                        cp_async16(&ws.Fst[buf][el][sp * 4], fsrc, fok ? 16 : 0);
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int c = sp * 4 + u;
                            ws.Fst[buf][el][c] = (ok && c < nch) ? __ldg(features + (size_t)ws.Gid[q] * C + ch0 + c) : 0.f;
                        }
                    }
                }
                cp_async_commit();
            };
            issue(0);
            for (int st = 0; st < nst; st++) {
                const int buf = st & 1;
                if (st + 1 < nst) { issue(st + 1); cp_async_wait<1>(); }
                else cp_async_wait<0>();
                __syncwarp();
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&ws.Wst[buf][e][pg * 8]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&ws.Wst[buf][e][pg * 8 + 4]);
                    const float4 f0 = *reinterpret_cast<const float4*>(&ws.Fst[buf][e][cg * 4]);
                    const float2 fa = make_float2(f0.x, f0.y), fb = make_float2(f0.z, f0.w);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float2 w2 = make_float2(wv[i], wv[i]);
                        acc[i][0] = ffma2(fa, w2, acc[i][0]);
                        acc[i][1] = ffma2(fb, w2, acc[i][1]);
                    }
                }
                __syncwarp();  // buffer `buf` may be refilled
            }
        }

        // ---- non-finite features (see acc_nonfinite above): a warp that sees a non-finite accumulator recomputes its
        // 32 px x 32 ch with the guarded loop (accumulate only where the weight is non-zero)
        {
            float2 z = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                z = ffma2(acc[i][0], make_float2(0.f, 0.f), z);
                z = ffma2(acc[i][1], make_float2(0.f, 0.f), z);
            }
            const float t = z.x + z.y;
            if (__any_sync(0xffffffffu, t != t)) {
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
                for (uint32_t e = 0; e < n; e++) {
                    const uint2 mt = meta_of(e);
                    if (!((mt.y >> warp) & 1u)) continue;
                    const float* wr = &chunk_ptr(e)->w[e & (kChunkEntries - 1)][warp * 32 + pg * 8];
                    const float* fr = features + (size_t)mt.x * C + ch0 + cg * 4;
                    float f[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) f[k] = (cg * 4 + k < nch) ? __ldg(fr + k) : 0.f;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float w = wr[i];
                        if (w != 0.f) {
                            acc[i][0].x = fmaf(f[0], w, acc[i][0].x);
                            acc[i][0].y = fmaf(f[1], w, acc[i][0].y);
                            acc[i][1].x = fmaf(f[2], w, acc[i][1].x);
                            acc[i][1].y = fmaf(f[3], w, acc[i][1].y);
                        }
                    }
                }
            }
        }

        // ---- out = acc + T * bg (forward.cu:372-373)
        if (row < (uint32_t)H) {
            const bool vec = ((W & 3) == 0) && (col0 + 8 <= (uint32_t)W);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int chl = cg * 4 + k;
                if (chl >= nch) continue;
                const float bgc = bg_color[ch0 + chl];
                float* dst = out_color + (size_t)(ch0 + chl) * plane + (size_t)W * row + col0;
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; i++) o[i] = ((k & 1) ? acc[i][k >> 1].y : acc[i][k >> 1].x) + Tv[i] * bgc;
                if (vec) {
                    reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
                    reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        if (col0 + i < (uint32_t)W) dst[i] = o[i];
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------ host side
size_t pool_bytes(int tiles, uint32_t chunks, int64_t R, PoolView* v, void* base) {
    size_t off = 0;
    char* p = (char*)base;
    auto take = [&](size_t n) { size_t o = off; off += align_up(n); return p ? p + o : nullptr; };
    void* hdr = take(sizeof(PoolHdr));
    void* dbase = take(4 * (size_t)tiles);
    void* cnt = take(4 * (size_t)tiles);
    void* dir = take(4 * ((size_t)(R / kChunkEntries) + (size_t)tiles + 1));
    void* ch = take(sizeof(WChunk) * (size_t)chunks);
    if (v) {
        v->hdr = (PoolHdr*)hdr; v->dirbase = (uint32_t*)dbase; v->count = (uint32_t*)cnt; v->dir = (uint32_t*)dir;
        v->chunks = (WChunk*)ch; v->capacity = chunks;
    }
    return off;
}

}  // namespace

// ---- weight-pool slots ---------------------------------------------------------------------------------------
static inline int num_tiles(const sgb_view_inputs& in) {
    return ((in.W + SGB_TILE - 1) / SGB_TILE) * ((in.H + SGB_TILE - 1) / SGB_TILE);
}

static PoolSlot* pool_find(sgb_ctx* ctx, const sgb_view_inputs& in, int64_t R, BinView b) {
    for (PoolSlot& sl : ctx->pools)
        if (sl.valid && sl.key_bin == (const void*)b.point_list && sl.key_R == R && sl.key_W == in.W && sl.key_H == in.H &&
            sl.key_P == in.P) {
            sl.stamp = ++ctx->pool_clock;
            return &sl;
        }
    return nullptr;
}

// Slot for a view that is about to be (re)built: the one already keyed by this binning state (a new forward
// through the same pointer replaces it), else an empty one, else the least recently used.
static PoolSlot* pool_acquire(sgb_ctx* ctx, BinView b) {
    PoolSlot* pick = nullptr;
    for (PoolSlot& sl : ctx->pools)
        if (sl.key_bin == (const void*)b.point_list) { pick = &sl; break; }
    if (!pick)
        for (PoolSlot& sl : ctx->pools)
            if (!sl.valid && !sl.key_bin) { pick = &sl; break; }
    if (!pick) {
        pick = &ctx->pools[0];
        for (PoolSlot& sl : ctx->pools)
            if (sl.stamp < pick->stamp) pick = &sl;
    }
    pick->valid = false;
    pick->key_bin = (const void*)b.point_list;
    pick->stamp = ++ctx->pool_clock;
    return pick;
}

static uint64_t pool_first_guess(sgb_ctx* ctx, int tiles, int64_t R) {
    // ~8 chunks (128 touching Gaussians) per tile, bounded by the instance count, at least the high-water mark
    uint64_t guess = (uint64_t)tiles * 8;
    const uint64_t by_R = (uint64_t)(R / kChunkEntries) + (uint64_t)tiles;
    if (guess > by_R) guess = by_R;
    if (guess < ctx->pool_chunks_hint) guess = ctx->pool_chunks_hint;
    if (guess < 16) guess = 16;
    return guess;
}

static int launch_alpha_pass(sgb_ctx* ctx, const sgb_view_inputs& in, GeomView g, BinView b, ImgView im,
                             float* out_depth, const PoolView& pv, cudaStream_t s) {
    const int tiles = num_tiles(in);
    const size_t smem = sizeof(AlphaSmem);
    static DeviceOnce attr_set;
    if (attr_set.first_use_on_device()) {
        SGB_CUDA(cudaFuncSetAttribute(alpha_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SGB_CUDA(cudaFuncSetAttribute(alpha_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    SGB_CUDA(cudaMemsetAsync(pv.hdr, 0, sizeof(PoolHdr), s));
    StageTimer t(ctx, ST_ALPHA, s);
    if (out_depth)
        alpha_pass_kernel<true><<<tiles, kThreads, smem, s>>>(im.ranges, b.point_list, in.W, in.H, g.rec, im.final_T,
                                                            im.n_contrib, im.tile_last, out_depth, pv);
    else
        alpha_pass_kernel<false><<<tiles, kThreads, smem, s>>>(im.ranges, b.point_list, in.W, in.H, g.rec, im.final_T,
                                                             im.n_contrib, im.tile_last, nullptr, pv);
    SGB_LAUNCH_CHECK("alpha_pass_kernel", in.debug, s);
    ctx->launches += 1;
    return SGB_OK;
}

static int launch_forward_gemm(sgb_ctx* ctx, const sgb_view_inputs& in, ImgView im, const float* colors,
                               float* out_color, const PoolView& pv, cudaStream_t s) {
    const int tiles = num_tiles(in);
    const int chunks = (in.C + 63) / 64;
    const bool vec = (in.C % 4 == 0) && ((reinterpret_cast<uintptr_t>(colors) & 15) == 0);
    if (vec && blend_mma_enabled()) return launch_forward_mma(ctx, in, im, colors, out_color, pv, s);  // opt-in experiment
    // SGB_FWD_WARP=1 selects the warp-autonomous kernel (per-strip compacted lists; measured SLOWER than the TMA ring on
    // K3: 2.98 vs 2.56 ms — kept for A/B measurements and for scenes with sparser strips)
    static const bool use_warp = [] { const char* e = getenv("SGB_FWD_WARP"); return e && e[0] == '1'; }();
    StageTimer t(ctx, ST_BLEND_FWD, s);
    ctx->launches += 1;
    if (use_warp) {
        const int nslices = (in.C + 31) / 32;
        const int SL = nslices >= 4 ? 4 : nslices;  // 128 channels per CTA: prologue amortised, CTAs still short
        const int groups = (nslices + SL - 1) / SL;
        const size_t smem_w = sizeof(FwdWarpSmem) * (kThreads / 32);
        if (vec)
            blend_forward_warp_kernel<true><<<tiles * groups, kThreads, smem_w, s>>>(in.W, in.H, in.C, SL, colors, in.background,
                                                                                    im.final_T, pv, out_color);
        else
            blend_forward_warp_kernel<false><<<tiles * groups, kThreads, smem_w, s>>>(in.W, in.H, in.C, SL, colors, in.background,
                                                                                     im.final_T, pv, out_color);
        SGB_LAUNCH_CHECK("blend_forward_warp_kernel", in.debug, s);
        return SGB_OK;
    }
    if (vec) {
        constexpr int NS = 5;
        const size_t smem_f = (size_t)NS * kChunkEntries * (SGB_TILE_PIX + 64) * sizeof(float);
        static DeviceOnce fattr;
        if (fattr.first_use_on_device()) {
            SGB_CUDA(cudaFuncSetAttribute(blend_forward_tma_kernel<64, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem_f));
        }
        blend_forward_tma_kernel<64, NS><<<tiles * chunks, kThreads, smem_f, s>>>(in.W, in.H, in.C, colors, in.background,
                                                                                  im.final_T, pv, out_color);
    } else {
        // feature rows that are not 16-byte aligned slices (C % 4 != 0) cannot be bulk-copied: plain loads
        blend_forward_v3_kernel<64, false><<<tiles * chunks, kThreads, 0, s>>>(in.W, in.H, in.C, colors, in.background,
                                                                                   im.final_T, pv, out_color);
    }
    SGB_LAUNCH_CHECK("blend_forward kernel", in.debug, s);
    return SGB_OK;
}

// pinned readback layout: [0, 8 * kMaxBatch) the R values of a geometry batch; then one PoolHdr (16 B) per view slot
static inline PoolHdr* pinned_hdr(sgb_ctx* ctx, int view_slot) {
    return reinterpret_cast<PoolHdr*>(reinterpret_cast<char*>(ctx->pinned) + 8 * kMaxBatch) + view_slot;
}

// Alpha pass of one view into a pool slot; no stream sync (the pool header is copied to pinned slot `view_slot`).
// The forward GEMM is launched by blend_forward_v3_gemm once blend_forward_v3_finish has validated the slot: the
// host blocks only for the alpha pass (not for the GEMM), so the caller keeps enqueueing the rest of its step
// while the GEMM runs.
int blend_forward_v3_alpha(sgb_ctx* ctx, int view_slot, const sgb_view_inputs& in, int64_t R, GeomView g, BinView b,
                           ImgView im, cudaStream_t s) {
    const int tiles = num_tiles(in);
    PoolSlot* sl = pool_acquire(ctx, b);
    uint64_t want = pool_first_guess(ctx, tiles, R);
    if (sl->chunks > want && sl->mem.cap >= pool_bytes(tiles, sl->chunks, R, nullptr, nullptr)) want = sl->chunks;
    const uint32_t chunks = (uint32_t)want;
    int rc = sl->mem.ensure(pool_bytes(tiles, chunks, R, nullptr, nullptr));
    if (rc) return rc;
    sl->chunks = chunks;
    PoolView pv;
    pool_bytes(tiles, chunks, R, &pv, sl->mem.p);
    rc = launch_alpha_pass(ctx, in, g, b, im, nullptr, pv, s);
    if (rc) return rc;
    SGB_CUDA(cudaMemcpyAsync(pinned_hdr(ctx, view_slot), pv.hdr, sizeof(PoolHdr), cudaMemcpyDeviceToHost, s));
    return SGB_OK;
}

int blend_forward_v3_gemm(sgb_ctx* ctx, const sgb_view_inputs& in, int64_t R, BinView b, ImgView im,
                          const float* colors, float* out_color, cudaStream_t s) {
    PoolSlot* sl = pool_find(ctx, in, R, b);
    if (!sl) { set_error("forward GEMM without validated weight rows"); return SGB_E_INVALID; }
    PoolView pv;
    pool_bytes(num_tiles(in), sl->chunks, R, &pv, sl->mem.p);
    return launch_forward_gemm(ctx, in, im, colors, out_color, pv, s);
}

// After the stream sync: 0 = the view is done (slot validated), 1 = the pool overflowed — the slot was grown to the
// real demand (the counter keeps counting past capacity) and the caller runs the alpha pass again; < 0 error.
int blend_forward_v3_finish(sgb_ctx* ctx, int view_slot, const sgb_view_inputs& in, int64_t R, BinView b) {
    PoolSlot* sl = nullptr;
    for (PoolSlot& c : ctx->pools)
        if (c.key_bin == (const void*)b.point_list) { sl = &c; break; }
    if (!sl) { set_error("weight-pool slot of the view vanished"); return SGB_E_INVALID; }
    const PoolHdr h = *pinned_hdr(ctx, view_slot);
    if (h.overflow) {
        const uint64_t need = (uint64_t)h.counter + h.counter / 8 + 64;
        if (need > ctx->pool_chunks_hint) ctx->pool_chunks_hint = need;
        sl->chunks = 0;  // re-carve with the new hint
        return 1;
    }
    ctx->stat_blended_pairs = (int64_t)h.blended;
    ctx->stat_pool_chunks = h.counter;
    if (h.counter > ctx->pool_chunks_hint) ctx->pool_chunks_hint = (uint64_t)h.counter + h.counter / 16 + 16;
    sl->valid = true;
    sl->key_R = R;
    sl->key_W = in.W;
    sl->key_H = in.H;
    sl->key_P = in.P;
    return 0;
}

// Weight rows of a view for its backward: the slot its forward filled, else rebuilt here (one more alpha pass and a
// stream sync for the pool check — only when the forward ran through another ctx or the slot was recycled).
static int pool_for_backward(sgb_ctx* ctx, const sgb_view_inputs& in, int64_t R, GeomView g, BinView b, ImgView im,
                             PoolView* pv, cudaStream_t s) {
    const int tiles = num_tiles(in);
    if (PoolSlot* hit = pool_find(ctx, in, R, b)) {
        pool_bytes(tiles, hit->chunks, R, pv, hit->mem.p);
        return SGB_OK;
    }
    for (int attempt = 0; attempt < 4; attempt++) {
        PoolSlot* sl = pool_acquire(ctx, b);
        const uint32_t chunks = (uint32_t)pool_first_guess(ctx, tiles, R);
        int rc = sl->mem.ensure(pool_bytes(tiles, chunks, R, nullptr, nullptr));
        if (rc) return rc;
        sl->chunks = chunks;
        pool_bytes(tiles, chunks, R, pv, sl->mem.p);
        rc = launch_alpha_pass(ctx, in, g, b, im, nullptr, *pv, s);
        if (rc) return rc;
        SGB_CUDA(cudaMemcpyAsync(pinned_hdr(ctx, 0), pv->hdr, sizeof(PoolHdr), cudaMemcpyDeviceToHost, s));
        SGB_CUDA(cudaStreamSynchronize(s));
        rc = blend_forward_v3_finish(ctx, 0, in, R, b);
        if (rc <= 0) return rc;
    }
    set_error("weight pool kept overflowing");
    return SGB_E_NOMEM;
}

int blend_backward_v3_dfeature(sgb_ctx* ctx, const sgb_view_inputs& in, int64_t R, GeomView g, BinView b, ImgView im,
                               const float* dL_dpix, float* dL_dcolors, cudaStream_t s) {
    PoolView pv;
    int rc = pool_for_backward(ctx, in, R, g, b, im, &pv, s);
    if (rc) return rc;
    if (blend_mma_enabled() && in.C % 4 == 0 && (reinterpret_cast<uintptr_t>(dL_dcolors) & 15) == 0)
        return launch_dfeature_mma(ctx, in, dL_dpix, dL_dcolors, pv, s);  // opt-in experiment
    const int tiles = num_tiles(in);
    const int chunks = (in.C + 63) / 64;
    const size_t smem_d = sizeof(float) * (64 * SGB_TILE_PIX + 8 * 2 * 16 * 36);
    static DeviceOnce attr_set;
    if (attr_set.first_use_on_device())
        SGB_CUDA(cudaFuncSetAttribute(dfeature_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
    StageTimer t(ctx, ST_DFEATURE, s);
    ctx->launches += 1;
    dfeature_gemm_kernel<64><<<tiles * chunks, kThreads, smem_d, s>>>(in.W, in.H, in.C, dL_dpix, pv, dL_dcolors);
    SGB_LAUNCH_CHECK("dfeature_gemm_kernel", in.debug, s);
    return SGB_OK;
}

// Tensor map of dL/dout (C, H, W) fp32 with a [16 ch][2 rows][16 px] box for the chain kernel's slab loads.  Returns
// false (the kernel then stages the slabs with cp.async) when the layout does not meet the TMA rules (base and row
// pitch multiples of 16 bytes) or the driver entry point is not available.  SGB_CHAIN_TMA=0 forces the cp.async path.
static bool encode_dl_map(CUtensorMap* map, const float* dL_dpix, int W, int H, int C) {
    using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (EncodeFn)fn;
    }();
    memset(map, 0, sizeof(*map));
    const char* off = getenv("SGB_CHAIN_TMA");
    if (off && off[0] == '0') return false;
    if (!encode || (W & 3) != 0 || (reinterpret_cast<uintptr_t>(dL_dpix) & 15) != 0) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C};
    const cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)W * H * sizeof(float)};
    const cuuint32_t box[3] = {SGB_TILE, 2, 16};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(dL_dpix), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int blend_backward_v3_chain(sgb_ctx* ctx, const sgb_view_inputs& in, int64_t R, GeomView g, BinView b, ImgView im,
                            const float* colors, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                            float* dL_dopacity, cudaStream_t s) {
    PoolView pv;
    int rc = pool_for_backward(ctx, in, R, g, b, im, &pv, s);
    if (rc) return rc;
    const int tiles = num_tiles(in);
    const bool vec = (in.C % 4 == 0) && ((reinterpret_cast<uintptr_t>(colors) & 15) == 0);
    if (vec && blend_mma_enabled())   // opt-in experiment
        return launch_chain_mma(ctx, in, g, im, colors, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, pv, s);
    // SGB_CHAIN_V3=1 selects the first-generation CTA-synchronous kernel (kept for A/B measurements)
    static const bool use_v3 = [] { const char* e = getenv("SGB_CHAIN_V3"); return e && e[0] == '1'; }();
    const size_t smem_g = sizeof(float) * (8 * kSeg * 32 + 2 * 16 * (kSeg + 4) + 8 * 2 * 16 * 32);
    const size_t smem_w = sizeof(ChainWarpSmem) * (kThreads / 32);
    static DeviceOnce attr_set;
    if (attr_set.first_use_on_device()) {
        SGB_CUDA(cudaFuncSetAttribute(chain_backward_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        SGB_CUDA(cudaFuncSetAttribute(chain_backward_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        SGB_CUDA(cudaFuncSetAttribute(chain_backward_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
        SGB_CUDA(cudaFuncSetAttribute(chain_backward_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
    }
    StageTimer t(ctx, ST_BLEND_BWD, s);
    ctx->launches += 1;
#define SGB_CHAIN_ARGS in.W, in.H, in.C, in.background, g.rec, colors, im.final_T, dL_dpix, pv, dL_dmean2D, dL_dconic, dL_dopacity
    if (use_v3) {
        if (vec) chain_backward_gemm_kernel<true><<<tiles, kThreads, smem_g, s>>>(SGB_CHAIN_ARGS);
        else chain_backward_gemm_kernel<false><<<tiles, kThreads, smem_g, s>>>(SGB_CHAIN_ARGS);
    } else {
        CUtensorMap dl_map;
        const int use_tma = encode_dl_map(&dl_map, dL_dpix, in.W, in.H, in.C) ? 1 : 0;
        if (vec) chain_backward_warp_kernel<true><<<tiles, kThreads, smem_w, s>>>(SGB_CHAIN_ARGS, dl_map, use_tma);
        else chain_backward_warp_kernel<false><<<tiles, kThreads, smem_w, s>>>(SGB_CHAIN_ARGS, dl_map, use_tma);
    }
#undef SGB_CHAIN_ARGS
    SGB_LAUNCH_CHECK("chain backward kernel", in.debug, s);
    return SGB_OK;
}

}  // namespace sgb
