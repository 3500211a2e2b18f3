// OPT-IN tensor-core forward blend (SGB_BLEND_MMA=1) — an experiment, not the default path.
//
// north_star rules tensor cores out of this path on the premise that it is a memory-bound gather/blend.  The
// round-1 profiles say otherwise for C >= 64: with the per-tile weights materialised once (blend_v3.cu) the forward
// is a dense contraction  out[256 px][C] = W^T[256 px][n] . F[n][C]  that the CUDA cores run at 27 TFLOP/s against
// a 70 TFLOP/s fp32 roof, i.e. FMA-issue bound at ~0.2 of the HBM roofline.  This file measures what the mandate
// costs: the same contraction on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) with fp32
// accuracy recovered by the error-compensated 3 x TF32 split
//     x = hi + lo,  hi = tf32(x) (round to nearest),  lo = tf32(x - hi)       (|lo| <= 2^-12 |x|)
//     w * f  ~=  hi_w * hi_f + hi_w * lo_f + lo_w * hi_f                      (lo_w * lo_f <= 2^-24 |w f| dropped)
// three kind::tf32 MMAs per K block accumulating in fp32.  The legacy mma.sync path was measured first
// (tools/microbench.cu): 277 TFLOP/s dense TF32 -> 67 TFLOP/s fp32-equivalent for the 3x split, no better than the
// CUDA cores; only tcgen05 (1.1 PFLOP/s dense TF32) changes the picture.
//
// Kernel: CTA = (tile, 128-channel slice), 256 threads, 2 CTAs/SM (256 TMEM columns each: D[256 px][128 ch] as two
// M = 128 halves).  Per batch of 16 list entries every thread moves 6 x 16 B of raw fp32 operands global ->
// registers (one batch ahead) -> hi / lo copies in shared memory, laid out as the UMMA canonical K-major no-swizzle
// operand (16-byte chunk = 4 consecutive entries k of one row, 8 rows = one 128-byte core matrix; layout and
// descriptor fields validated on hardware by tools/tc_probe.cu — both operands are naturally MN-major here, but the
// MN-major descriptor forms produced no output in the probe, so the staging threads transpose).  One elected thread issues the 12 MMAs of a batch
// (2 K blocks x 2 pixel halves x 3 products) and commits them to the stage's mbarrier; the epilogue reads TMEM with
// tcgen05.ld (TMEM lane = pixel), adds T * bg and stores the planar image.
// Reference semantics: forward.cu:355-356, 372-373 (accumulation order differs: fp32 tree inside the tensor core).
#include <cstdlib>
#include "common.cuh"
#include "blend_pool.cuh"

namespace sgb {

namespace {

constexpr int kMmaThreads = 256;
constexpr int kNch = 128;              // channels per CTA (MMA N)
constexpr int kBatch = 16;             // list entries per pipeline stage = 2 K blocks of 8
// Operand blocks (one K block of 8 entries x 128 rows).  Operands written by TRANSPOSING scalar stores use a padded
// geometry — leading offset (second K half) 144 B, row-group stride 288 B instead of 128 / 256 — which the UMMA
// descriptor expresses directly (LBO / SBO are free parameters) and which makes the 32 scalar stores of a warp hit 32
// distinct banks (ncu on the unpadded layout: 75 % of the shared-store wavefronts were bank conflicts and the tensor
// pipe was 16 % busy).  Operands stored with 16-byte pieces in their natural order keep the dense 128 / 256 geometry.
constexpr uint32_t kLboPad = 144, kSboPad = 288;
constexpr uint32_t kABlk = 16 * kSboPad;   // bytes of one (K block, pixel half) operand block: 16 row groups
constexpr uint32_t kBBlk = (kNch / 8) * kSboPad;
constexpr uint32_t kStageA = 2 * 2 * kABlk;   // [kb][mh]
constexpr uint32_t kStageB = 2 * kBBlk;       // [kb]
constexpr uint32_t kStageBytes = 2 * kStageA + 2 * kStageB;   // hi + lo of both operands = 48 KB
constexpr int kStages = 2;

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lead_bytes, uint32_t stride_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lead_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((stride_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version
    return d;                // no swizzle
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {  // round to nearest TF32 (low 13 mantissa bits zero)
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ float4 tf32_hi(float4 x) {
    return make_float4(tf32_rna(x.x), tf32_rna(x.y), tf32_rna(x.z), tf32_rna(x.w));
}
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

__global__ void __launch_bounds__(kMmaThreads, 2) blend_forward_mma_kernel(
    int W, int H, int C, const float* __restrict__ features, const float* __restrict__ bg_color,
    const float* __restrict__ final_T, PoolView pool, float* __restrict__ out_color) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t empty_bar[kStages], done_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t Cdir[192];
    __shared__ float bgS[kNch];

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int nslices = (C + kNch - 1) / kNch;
    const int tile = blockIdx.x / nslices;
    const int ch0 = (blockIdx.x % nslices) * kNch;
    const int nch = min(kNch, C - ch0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};
    const size_t plane = (size_t)H * W;
    const uint32_t n = pool.count[tile];
    const uint32_t dbase = pool.dirbase[tile];

    // pixel of this thread in the epilogue: TMEM lane l of pixel half mh <-> tile pixel mh*128 + l
    const int mh_e = warp >> 2;
    const int p_tile = mh_e * 128 + (warp & 3) * 32 + lane;
    const uint32_t px = pix_min.x + (p_tile & 15), py = pix_min.y + (p_tile >> 4);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const float Tfin = inside ? final_T[(size_t)W * py + px] : 0.f;
    if (tid < nch) bgS[tid] = bg_color[ch0 + tid];

    if (n == 0) {  // nothing blended: out = T * bg (forward.cu:372-373)
        __syncthreads();
        if (inside)
            for (int c = 0; c < nch; c++) out_color[(size_t)(ch0 + c) * plane + (size_t)W * py + px] = Tfin * bgS[c];
        return;
    }
    const int nb = (int)((n + kBatch - 1) / kBatch);
    for (int k = tid; k < min(nb, 192); k += kMmaThreads) Cdir[k] = chunk_of(pool, dbase, k);
    if (tid == 0) {
        for (int i = 0; i < kStages; i++) mbar_init(&empty_bar[i], 1);
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    auto chunk_ptr = [&](int bi) { return pool.chunks + (bi < 192 ? Cdir[bi] : chunk_of(pool, dbase, bi)); };

    // ---- staging roles.  lane -> (qq = lane >> 3: one of 4 adjacent 16-byte pieces, e8 = lane & 7: entry of the K
    // block): 8 lanes fill one 128-byte core matrix, a warp 512 contiguous bytes; per row 64 contiguous global bytes.
    const int qq = lane >> 3, e8 = lane & 7;
    // raw operands of the next TWO batches live in two explicit register sets (global latency >> one batch of MMAs)
    float4 wregA[4], fregA[2], wregB[4], fregB[2];
    auto load_batch = [&](int b, float4 (&wreg)[4], float4 (&freg)[2]) {
        const WChunk* ck = chunk_ptr(b);
        const int left = (int)n - b * kBatch;  // entries of this batch that exist
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int c = it * 8 + warp, kb = c >> 4, qblk = c & 15;
            const int e = kb * 8 + e8;
            wreg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < left) wreg[it] = __ldg(reinterpret_cast<const float4*>(&ck->w[e][(qblk * 4 + qq) * 4]));
        }
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int e = it * 8 + e8;           // kb = it
            const int chl = (warp * 4 + qq) * 4; // channel of this 16-byte piece inside the slice
            freg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < left && chl < nch) {
                const uint32_t gid = __ldg(&ck->meta[e].x);
                freg[it] = __ldg(reinterpret_cast<const float4*>(features + (size_t)gid * C + ch0 + chl));
            }
        }
    };
    // K-major no-swizzle operand block (validated by tools/tc_probe.cu; the MN-major forms produced no output there):
    // element (row r of the 128-row block, entry k of the 8-entry K block) at
    //     (r / 8) * SBO + (k / 4) * LBO + (r % 8) * 16 + (k % 4) * 4          (LBO = 144 B, SBO = 288 B, see above).
    // A thread holds 4 consecutive rows of ONE entry (a 16-byte piece of a weight / feature row), i.e. 4 scalar stores.
    auto put4 = [&](unsigned char* blk, int r0, float4 v) {
        const uint32_t col = (uint32_t)(e8 >> 2) * kLboPad + (uint32_t)(e8 & 3) * 4u;
        const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = r0 + j;
            *reinterpret_cast<float*>(blk + (uint32_t)(r >> 3) * kSboPad + (uint32_t)(r & 7) * 16u + col) = x[j];
        }
    };
    auto store_batch = [&](int st, float4 (&wreg)[4], float4 (&freg)[2]) {
        unsigned char* base = smem_raw + (size_t)st * kStageBytes;
        unsigned char* a_hi = base;
        unsigned char* a_lo = base + kStageA;
        unsigned char* b_hi = base + 2 * kStageA;
        unsigned char* b_lo = base + 2 * kStageA + kStageB;
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int c = it * 8 + warp, kb = c >> 4, qblk = c & 15;
            const int q = qblk * 4 + qq, mh = q >> 5, ql = q & 31;
            const uint32_t blk = (uint32_t)(kb * 2 + mh) * kABlk;
            const float4 hi = tf32_hi(wreg[it]);
            put4(a_hi + blk, ql * 4, hi);
            put4(a_lo + blk, ql * 4, tf32_hi(sub4(wreg[it], hi)));
        }
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int q = warp * 4 + qq;
            const float4 hi = tf32_hi(freg[it]);
            put4(b_hi + (uint32_t)it * kBBlk, q * 4, hi);
            put4(b_lo + (uint32_t)it * kBBlk, q * 4, tf32_hi(sub4(freg[it], hi)));
        }
    };
    // instruction descriptor: D f32 | A, B tf32 | A, B K-major | N >> 3 | M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNch >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    auto step = [&](int b, float4 (&wreg)[4], float4 (&freg)[2]) {
        const int st = b % kStages;
        if (b >= kStages) mbar_wait(&empty_bar[st], (uint32_t)(((b / kStages) - 1) & 1));  // MMAs of batch b-2 retired
        store_batch(st, wreg, freg);
        if (b + 2 < nb) load_batch(b + 2, wreg, freg);  // lands while the tensor core works on this batch and the next
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sbase = smem_u32(smem_raw) + (uint32_t)st * kStageBytes;
            const int nkb = ((int)n - b * kBatch > 8) ? 2 : 1;  // a batch whose second K block is empty skips it
            for (int kb = 0; kb < nkb; kb++) {
                const uint32_t bh = sbase + 2 * kStageA + kb * kBBlk, bl = bh + kStageB;
                const uint64_t dbh = umma_desc(bh, kLboPad, kSboPad), dbl = umma_desc(bl, kLboPad, kSboPad);
                for (int mh = 0; mh < 2; mh++) {
                    const uint32_t ah = sbase + (uint32_t)(kb * 2 + mh) * kABlk, al = ah + kStageA;
                    const uint64_t dah = umma_desc(ah, kLboPad, kSboPad), dal = umma_desc(al, kLboPad, kSboPad);
                    const uint32_t d = tmem + (uint32_t)mh * kNch;
                    const uint32_t first = (b == 0 && kb == 0) ? 0u : 1u;
                    umma_tf32(d, dal, dbh, idesc, first);  // small terms first
                    umma_tf32(d, dah, dbl, idesc, 1u);
                    umma_tf32(d, dah, dbh, idesc, 1u);
                }
            }
            umma_commit(b + 1 < nb ? &empty_bar[st] : &done_bar);
        }
    };
    load_batch(0, wregA, fregA);
    if (nb > 1) load_batch(1, wregB, fregB);
    for (int b = 0; b < nb; b += 2) {
        step(b, wregA, fregA);
        if (b + 1 < nb) step(b + 1, wregB, fregB);
    }
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: out = acc + T * bg; TMEM lane = pixel, 32 channels per tcgen05.ld
    bool poisoned = false;
    for (int c0 = 0; c0 < kNch; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mh_e * kNch + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (inside) {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int c = c0 + j;
                if (c < nch) {
                    const float v = __uint_as_float(r[j]);
                    poisoned |= !(fabsf(v) <= 3.0e38f);
                    out_color[(size_t)(ch0 + c) * plane + (size_t)W * py + px] = v + Tfin * bgS[c];
                }
            }
        }
    }
    // Non-finite features: 0 * inf = NaN inside the dense product, where the reference only touches the pixels that
    // blend the Gaussian (see blend_v3.cu).  A pixel that saw a non-finite value recomputes its channels with the
    // guarded scalar loop straight from global memory (rare path).
    if (poisoned) {
        const int woff_p = p_tile;
        for (int c0 = 0; c0 < nch; c0 += 16) {
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; j++) acc[j] = 0.f;
            for (uint32_t e = 0; e < n; e++) {
                const WChunk* ck = pool.chunks + chunk_of(pool, dbase, (int)(e / kChunkEntries));
                const int s = (int)(e & (kChunkEntries - 1));
                const float w = ck->w[s][woff_p];
                if (w == 0.f) continue;
                const float* fr = features + (size_t)ck->meta[s].x * C + ch0 + c0;
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (c0 + j < nch) acc[j] = fmaf(__ldg(fr + j), w, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; j++)
                if (c0 + j < nch) out_color[(size_t)(ch0 + c0 + j) * plane + (size_t)W * py + px] = acc[j] + Tfin * bgS[c0 + j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------ dL/dfeature on tcgen05
// dF[entry][ch] = sum over the tile's 256 pixels of w[entry][px] * dL/dout[px][ch]     (backward.cu:519 summed per tile)
// as  D[M = 128 entries][N = 128 channels] += A[M][K = 8 pixels] . B[N][K]^T,  32 K blocks per tile.  Both operands are
// K-major as they lie in memory (a weight row is contiguous in pixels, a dL/dout channel plane row too), so staging is
// a straight 16-byte copy plus the hi / lo split.  CTA = (tile, 128-channel slice); one pipeline stage = one tile row
// (16 pixels = 2 K blocks): 128 x 64 B of weights and 128 x 64 B of dL/dout; 3 stages, operands loaded two stages
// ahead into registers.  Epilogue: TMEM lane = entry; each thread adds its entry's channels to dL_dcolors with 16-byte
// vector reductions (red.global.add.v4.f32).  Lists longer than 128 entries take further passes.
constexpr int kDfStages = 3;
constexpr uint32_t kDfBlk = 16 * 256;                 // one K block of a 128-row operand
constexpr uint32_t kDfStageBytes = 4 * 2 * kDfBlk;    // {A hi, A lo, B hi, B lo} x 2 K blocks = 32 KB

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kMmaThreads, 2) dfeature_mma_kernel(int W, int H, int C,
                                                                      const float* __restrict__ dL_dpixels,
                                                                      PoolView pool, float* __restrict__ dL_dcolors) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t empty_bar[kDfStages], done_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ const float* Wrow[128];
    __shared__ uint32_t Gid[128];

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int nslices = (C + kNch - 1) / kNch;
    const int tile = blockIdx.x / nslices;
    const int ch0 = (blockIdx.x % nslices) * kNch;
    const int nch = min(kNch, C - ch0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = pool.count[tile];
    if (n == 0) return;
    const uint32_t dbase = __ldg(pool.dirbase + tile);
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};
    const size_t plane = (size_t)H * W;
    const bool rows16 = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(dL_dpixels) & 15) == 0;

    if (tid == 0) {
        for (int i = 0; i < kDfStages; i++) mbar_init(&empty_bar[i], 1);
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNch >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    // staging role: quarter-warp = 8 consecutive rows of one 16-byte piece (conflict-free 128-byte stores), the four
    // quarter-warps = the four pieces of a 64-byte row segment (8 rows x 64 contiguous global bytes per instruction)
    const int r8 = lane & 7, quad = lane >> 3;
    const uint32_t soff = (uint32_t)(quad >> 1) * kDfBlk + (uint32_t)(quad & 1) * 128u + (uint32_t)r8 * 16u;
    int done_phase = 0;
    int gstage = 0;  // stages issued so far over all passes (stage buffer = gstage % kDfStages)

    for (uint32_t base = 0; base < n; base += 128) {
        const int cnt = (int)min(128u, n - base);
        __syncthreads();  // previous pass finished with Wrow / Gid and with the accumulator
        if (tid < 128) {
            if (tid < cnt) {
                const WChunk* ck = pool.chunks + chunk_of(pool, dbase, (int)((base + tid) / kChunkEntries));
                const int sidx = (base + tid) & (kChunkEntries - 1);
                Wrow[tid] = &ck->w[sidx][0];
                Gid[tid] = ck->meta[sidx].x;
            } else {
                Wrow[tid] = nullptr;
            }
        }
        __syncthreads();
        float4 aA[2], bA[2], aB[2], bB[2];
        auto load_stage = [&](int ty, float4 (&a)[2], float4 (&b)[2]) {  // tile row ty: pixels 16 ty .. 16 ty + 15
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int row = (i * 8 + warp) * 8 + r8;   // entry (A) / channel (B) row of this thread
                const float* wr = Wrow[row];
                a[i] = wr ? __ldg(reinterpret_cast<const float4*>(wr + ty * 16 + quad * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const uint32_t y = pix_min.y + ty, x = pix_min.x + quad * 4;
                b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < nch && y < (uint32_t)H) {
                    const float* src = dL_dpixels + (size_t)(ch0 + row) * plane + (size_t)W * y + x;
                    if (rows16 && x + 4 <= (uint32_t)W) b[i] = __ldg(reinterpret_cast<const float4*>(src));
                    else {
                        if (x < (uint32_t)W) b[i].x = __ldg(src);
                        if (x + 1 < (uint32_t)W) b[i].y = __ldg(src + 1);
                        if (x + 2 < (uint32_t)W) b[i].z = __ldg(src + 2);
                        if (x + 3 < (uint32_t)W) b[i].w = __ldg(src + 3);
                    }
                }
            }
        };
        auto step = [&](int ty, float4 (&a)[2], float4 (&b)[2]) {
            const int st = gstage % kDfStages;
            if (gstage >= kDfStages) mbar_wait(&empty_bar[st], (uint32_t)(((gstage / kDfStages) - 1) & 1));
            unsigned char* sb = smem_raw + (size_t)st * kDfStageBytes;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const uint32_t off = soff + (uint32_t)(i * 8 + warp) * 256u;
                const float4 ah = tf32_hi(a[i]), bh = tf32_hi(b[i]);
                *reinterpret_cast<float4*>(sb + off) = ah;
                *reinterpret_cast<float4*>(sb + 2 * kDfBlk + off) = tf32_hi(sub4(a[i], ah));
                *reinterpret_cast<float4*>(sb + 4 * kDfBlk + off) = bh;
                *reinterpret_cast<float4*>(sb + 6 * kDfBlk + off) = tf32_hi(sub4(b[i], bh));
            }
            if (ty + 2 < SGB_TILE) load_stage(ty + 2, a, b);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t s0 = smem_u32(smem_raw) + (uint32_t)st * kDfStageBytes;
#pragma unroll
                for (int kb = 0; kb < 2; kb++) {
                    const uint64_t dah = umma_desc(s0 + kb * kDfBlk, 128, 256), dal = umma_desc(s0 + (2 + kb) * kDfBlk, 128, 256);
                    const uint64_t dbh = umma_desc(s0 + (4 + kb) * kDfBlk, 128, 256), dbl = umma_desc(s0 + (6 + kb) * kDfBlk, 128, 256);
                    umma_tf32(tmem, dal, dbh, idesc, (ty == 0 && kb == 0) ? 0u : 1u);
                    umma_tf32(tmem, dah, dbl, idesc, 1u);
                    umma_tf32(tmem, dah, dbh, idesc, 1u);
                }
                umma_commit(ty + 1 < SGB_TILE ? &empty_bar[st] : &done_bar);
            }
            gstage++;
        };
        load_stage(0, aA, bA);
        load_stage(1, aB, bB);
#pragma unroll 1
        for (int ty = 0; ty < SGB_TILE; ty += 2) {
            step(ty, aA, bA);
            step(ty + 1, aB, bB);
        }
        // NOTE on the empty barriers: the last stage of a pass commits to done_bar instead of its empty barrier, so the
        // stage-buffer parity bookkeeping only counts commits that went to empty_bar: keep them in step by one more
        // commit to that stage's empty barrier (covers no new MMAs, completes immediately after the previous ones).
        if (tid == 0) umma_commit(&empty_bar[(gstage - 1) % kDfStages]);
        mbar_wait(&done_bar, (uint32_t)(done_phase & 1));
        done_phase++;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue: TMEM lane = entry; warps 0-3 take channels 0..63 of the slice, warps 4-7 channels 64..127
        const int e = (warp & 3) * 32 + lane;
        const int cbase = (warp >> 2) * 64;
#pragma unroll 1
        for (int c0 = cbase; c0 < cbase + 64; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (e < cnt) {
                float* dst = dL_dcolors + (size_t)Gid[e] * C + ch0 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (c0 + j < nch)   // C % 4 == 0 on this path: a 4-channel group is in range as a whole
                        red_add_v4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------ chain backward on tcgen05
// s-pass  S[px][entry] = sum_ch dL/dout[px][ch] * F[entry][ch]  as  D[M = 128 px][N = 128 entries] (two pixel halves),
// K = channels (8 per K block), then the reference's back-to-front chain (backward.cu:477-550 in dot-product form, the
// code of chain_backward_warp_kernel) with thread = pixel = TMEM lane.  CTA = tile; the list is walked in passes of 128
// entries from the back.  A = dL/dout is MN-major in memory (pixels contiguous): the staging threads transpose it into
// the K-major operand with scalar stores; B = features is K-major as it lies (channels of a row contiguous).  Stage =
// 16 channels; 2 stages of 48 KB; during the chain phase the stage memory holds the parked S columns and the
// reduction rows of every warp.
constexpr uint32_t kChBBlk = 16 * 256;   // features: dense geometry (LBO 128, SBO 256)
constexpr uint32_t kChStageA = 2 * 2 * kABlk, kChStageB = 2 * kChBBlk;
constexpr uint32_t kChStageBytes = 2 * kChStageA + 2 * kChStageB;   // 48 KB

template <int N>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kMmaThreads, 2) chain_backward_mma_kernel(
    int W, int H, int C, const float* __restrict__ bg_color, const SplatRec* __restrict__ rec,
    const float* __restrict__ features, const float* __restrict__ final_Ts, const float* __restrict__ dL_dpixels,
    PoolView pool, float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic2D, float* __restrict__ dL_dopacity) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t empty_bar[2], done_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ const float* Wrow[128];
    __shared__ uint32_t Gid[128], Mask[128];
    __shared__ float4 RecA[128], RecB[128];
    __shared__ int BufCol[8][4];

    const int tiles_x = (W + SGB_TILE - 1) / SGB_TILE;
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = pool.count[tile];
    if (n == 0) return;
    const uint32_t dbase = pool.dirbase[tile];
    const uint2 pix_min = {(uint32_t)(tile % tiles_x) * SGB_TILE, (uint32_t)(tile / tiles_x) * SGB_TILE};
    const size_t plane = (size_t)H * W;
    const bool rows16 = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(dL_dpixels) & 15) == 0);

    // own pixel (chain phase): TMEM lane l of pixel half mh <-> tile pixel mh * 128 + l; warp = strip, as in the pool masks
    const int mh_e = warp >> 2;
    const int p_tile = mh_e * 128 + (warp & 3) * 32 + lane;
    const uint2 pix = {pix_min.x + (uint32_t)(p_tile & 15), pix_min.y + (uint32_t)(p_tile >> 4)};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;

    int bg_nonzero = 0;
    for (int ch = tid; ch < C; ch += kMmaThreads) bg_nonzero |= (bg_color[ch] != 0.f);
    if (tid == 0) {
        mbar_init(&empty_bar[0], 1);
        mbar_init(&empty_bar[1], 1);
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    bg_nonzero = __syncthreads_or(bg_nonzero);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    float bgdot = 0.f;  // backward.cu:527-529; vanishes for an all-zero background
    if (inside && bg_nonzero)
        for (int ch = 0; ch < C; ch++) bgdot += bg_color[ch] * __ldg(dL_dpixels + (size_t)ch * plane + pix_id);
    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    float T = T_final;
    float last_alpha = 0.f, s_last = 0.f, A = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    const int nst = (C + 15) / 16;          // stages (16 channels) per pass
    const int k8 = lane & 7, qq = lane >> 3;  // dL staging: channel k8 of the K block, 16-byte piece qq of a tile row
    int gstage = 0, done_phase = 0;
    // chain-phase views of the (idle) stage memory: per warp S[32 entries][32 px] (4 KB) and 18 reduction rows (2.3 KB)
    float* Sw = reinterpret_cast<float*>(smem_raw) + warp * 1024;
    float* RB = reinterpret_cast<float*>(smem_raw + 32768) + warp * 576;

    const int npass = (int)((n + 127) / 128);
    for (int ps = npass - 1; ps >= 0; ps--) {
        const uint32_t base = (uint32_t)ps * 128u;
        const int cnt = (int)min(128u, n - base);
        __syncthreads();  // previous pass: chain phase done with S / RB / entry records, accumulator free
        if (tid < 128) {
            if (tid < cnt) {
                const WChunk* ck = pool.chunks + chunk_of(pool, dbase, (int)((base + tid) / kChunkEntries));
                const int sidx = (base + tid) & (kChunkEntries - 1);
                const uint2 mt = ck->meta[sidx];
                Wrow[tid] = &ck->w[sidx][0];
                Gid[tid] = mt.x;
                Mask[tid] = mt.y;
                const float4* rp = reinterpret_cast<const float4*>(rec + mt.x);
                RecA[tid] = __ldg(rp);
                RecB[tid] = __ldg(rp + 1);
            } else {
                Mask[tid] = 0u;
            }
        }
        __syncthreads();
        // ---- s-pass on the tensor core
        float4 dA[4], fA[2], dB[4], fB[2];
        auto load_stage = [&](int sl, float4 (&d)[4], float4 (&f)[2]) {
#pragma unroll
            for (int it = 0; it < 4; it++) {   // dL/dout: channel (kb, k8), tile row ty, columns qq*4..
                const int c = it * 8 + warp, kb = c >> 4, ty = c & 15;
                const int ch = sl * 16 + kb * 8 + k8;
                const uint32_t y = pix_min.y + ty, x = pix_min.x + qq * 4;
                d[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ch < C && y < (uint32_t)H) {
                    const float* src = dL_dpixels + (size_t)ch * plane + (size_t)W * y + x;
                    if (rows16 && x + 4 <= (uint32_t)W) d[it] = __ldg(reinterpret_cast<const float4*>(src));
                    else {
                        if (x < (uint32_t)W) d[it].x = __ldg(src);
                        if (x + 1 < (uint32_t)W) d[it].y = __ldg(src + 1);
                        if (x + 2 < (uint32_t)W) d[it].z = __ldg(src + 2);
                        if (x + 3 < (uint32_t)W) d[it].w = __ldg(src + 3);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {      // features: entry row (i*8+warp)*8 + k8, channels sl*16 + qq*4 ..
                const int row = (i * 8 + warp) * 8 + k8;
                const int chb = sl * 16 + qq * 4;
                f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < cnt && chb < C) f[i] = __ldg(reinterpret_cast<const float4*>(features + (size_t)Gid[row] * C + chb));
            }
        };
        auto step = [&](int sl, float4 (&d)[4], float4 (&f)[2]) {
            const int st = gstage & 1;
            if (gstage >= 2) mbar_wait(&empty_bar[st], (uint32_t)(((gstage >> 1) - 1) & 1));
            unsigned char* sb = smem_raw + (size_t)st * kChStageBytes;
            unsigned char* a_hi = sb;
            unsigned char* a_lo = sb + kChStageA;
            unsigned char* b_hi = sb + 2 * kChStageA;
            unsigned char* b_lo = sb + 2 * kChStageA + kChStageB;
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const int c = it * 8 + warp, kb = c >> 4, ty = c & 15;
                const int p0 = ty * 16 + qq * 4, mh = p0 >> 7, r0 = p0 & 127;
                const uint32_t blk = (uint32_t)(kb * 2 + mh) * kABlk;
                const uint32_t col = (uint32_t)(k8 >> 2) * kLboPad + (uint32_t)(k8 & 3) * 4u;
                const float4 hi = tf32_hi(d[it]);
                const float4 lo = tf32_hi(sub4(d[it], hi));
                const float xh[4] = {hi.x, hi.y, hi.z, hi.w}, xl[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int r = r0 + j;
                    const uint32_t off = blk + (uint32_t)(r >> 3) * kSboPad + (uint32_t)(r & 7) * 16u + col;
                    *reinterpret_cast<float*>(a_hi + off) = xh[j];
                    *reinterpret_cast<float*>(a_lo + off) = xl[j];
                }
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const uint32_t off = (uint32_t)(qq >> 1) * kChBBlk + (uint32_t)(i * 8 + warp) * 256u + (uint32_t)(qq & 1) * 128u +
                                     (uint32_t)k8 * 16u;
                const float4 hi = tf32_hi(f[i]);
                *reinterpret_cast<float4*>(b_hi + off) = hi;
                *reinterpret_cast<float4*>(b_lo + off) = tf32_hi(sub4(f[i], hi));
            }
            if (sl + 2 < nst) load_stage(sl + 2, d, f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t s0 = smem_u32(smem_raw) + (uint32_t)st * kChStageBytes;
                const int nkb = (C - sl * 16 > 8) ? 2 : 1;
                for (int kb = 0; kb < nkb; kb++) {
                    const uint32_t bh = s0 + 2 * kChStageA + kb * kChBBlk, bl = bh + kChStageB;
                    const uint64_t dbh = umma_desc(bh, 128, 256), dbl = umma_desc(bl, 128, 256);
                    for (int mh = 0; mh < 2; mh++) {
                        const uint32_t ah = s0 + (uint32_t)(kb * 2 + mh) * kABlk, al = ah + kChStageA;
                        const uint64_t dah = umma_desc(ah, kLboPad, kSboPad), dal = umma_desc(al, kLboPad, kSboPad);
                        const uint32_t dcol = tmem + (uint32_t)mh * 128u;
                        umma_tf32(dcol, dal, dbh, idesc, (sl == 0 && kb == 0) ? 0u : 1u);
                        umma_tf32(dcol, dah, dbl, idesc, 1u);
                        umma_tf32(dcol, dah, dbh, idesc, 1u);
                    }
                }
                if (sl + 1 < nst) umma_commit(&empty_bar[st]);
                else { umma_commit(&empty_bar[st]); umma_commit(&done_bar); }
            }
            gstage++;
        };
        load_stage(0, dA, fA);
        if (nst > 1) load_stage(1, dB, fB);
#pragma unroll 1
        for (int sl = 0; sl < nst; sl += 2) {
            step(sl, dA, fA);
            if (sl + 1 < nst) step(sl + 1, dB, fB);
        }
        mbar_wait(&done_bar, (uint32_t)(done_phase & 1));
        done_phase++;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- chain phase: 32-entry column chunks from the back; the warp parks its 32 x 32 block of S in shared memory
        // (stage memory is idle: every MMA of this pass has completed) and runs the chain of chain_backward_warp_kernel
        for (int cc = (cnt - 1) >> 5; cc >= 0; cc--) {
            uint32_t r[32];
            tmem_ld32<0>(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mh_e * 128 + cc * 32), r);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j++) Sw[j * 32 + lane] = __uint_as_float(r[j]);
            __syncwarp();
            const int jhi = min(31, cnt - 1 - cc * 32);
            constexpr int RG = 3;
            int nbuf = 0;
            for (int j = jhi; j >= 0; j--) {
                const int col = cc * 32 + j;
                const bool mine = (Mask[col] >> warp) & 1u;   // uniform per warp (warp = strip)
                float gv[6];
#pragma unroll
                for (int v = 0; v < 6; v++) gv[v] = 0.f;
                if (mine) {
                    const float w = __ldg(Wrow[col] + p_tile);
                    if (w != 0.f) {
                        const float sdot = Sw[j * 32 + lane];
                        const float4 a = RecA[col], con_o = RecB[col];
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
#pragma unroll
                    for (int v = 0; v < 6; v++) {
                        const int rr = nbuf * 6 + v;
                        RB[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))] = gv[v];
                    }
                    if (lane == 0) BufCol[warp][nbuf] = col;   // which list entry each buffered slot belongs to
                    nbuf++;
                }
                // flush when the buffer is full or the chunk ends
                if (nbuf == RG || (j == 0 && nbuf > 0)) {
                    __syncwarp();
                    if (lane < nbuf * 6) {
                        const float4* row = reinterpret_cast<const float4*>(RB + lane * 32);
                        float4 t = row[lane & 7];
#pragma unroll
                        for (int q = 1; q < 8; q++) {
                            const float4 u = row[q ^ (lane & 7)];
                            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                        }
                        const float tot = (t.x + t.y) + (t.z + t.w);
                        const int slot = lane / 6, comp = lane - slot * 6;
                        const size_t id = Gid[BufCol[warp][slot]];
                        float* dst = comp < 2 ? dL_dmean2D + id * 3 + comp
                                   : comp < 5 ? dL_dconic2D + id * 4 + (comp == 4 ? 3 : comp - 2)
                                              : dL_dopacity + id;
                        red_add_f32(dst, tot);
                    }
                    __syncwarp();
                    nbuf = 0;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

}  // namespace

bool blend_mma_enabled() {
    static const bool on = [] { const char* e = getenv("SGB_BLEND_MMA"); return e && e[0] == '1'; }();
    return on;
}

int launch_forward_mma(sgb_ctx* ctx, const sgb_view_inputs& in, ImgView im, const float* colors, float* out_color,
                       const PoolView& pv, cudaStream_t s) {
    const int tiles = ((in.W + SGB_TILE - 1) / SGB_TILE) * ((in.H + SGB_TILE - 1) / SGB_TILE);
    const int slices = (in.C + kNch - 1) / kNch;
    const size_t smem = (size_t)kStages * kStageBytes + 128;
    static DeviceOnce attr;
    if (attr.first_use_on_device())
        SGB_CUDA(cudaFuncSetAttribute(blend_forward_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageTimer t(ctx, ST_BLEND_FWD, s);
    ctx->launches += 1;
    blend_forward_mma_kernel<<<tiles * slices, kMmaThreads, smem, s>>>(in.W, in.H, in.C, colors, in.background, im.final_T,
                                                                       pv, out_color);
    SGB_LAUNCH_CHECK("blend_forward_mma_kernel", in.debug, s);
    return SGB_OK;
}

int launch_dfeature_mma(sgb_ctx* ctx, const sgb_view_inputs& in, const float* dL_dpix, float* dL_dcolors,
                        const PoolView& pv, cudaStream_t s) {
    const int tiles = ((in.W + SGB_TILE - 1) / SGB_TILE) * ((in.H + SGB_TILE - 1) / SGB_TILE);
    const int slices = (in.C + kNch - 1) / kNch;
    const size_t smem = (size_t)kDfStages * kDfStageBytes + 1024;
    static DeviceOnce attr;
    if (attr.first_use_on_device())
        SGB_CUDA(cudaFuncSetAttribute(dfeature_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageTimer t(ctx, ST_DFEATURE, s);
    ctx->launches += 1;
    dfeature_mma_kernel<<<tiles * slices, kMmaThreads, smem, s>>>(in.W, in.H, in.C, dL_dpix, pv, dL_dcolors);
    SGB_LAUNCH_CHECK("dfeature_mma_kernel", in.debug, s);
    return SGB_OK;
}

int launch_chain_mma(sgb_ctx* ctx, const sgb_view_inputs& in, GeomView g, ImgView im, const float* colors,
                     const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, const PoolView& pv,
                     cudaStream_t s) {
    const int tiles = ((in.W + SGB_TILE - 1) / SGB_TILE) * ((in.H + SGB_TILE - 1) / SGB_TILE);
    const size_t smem = 2 * (size_t)kChStageBytes + 128;
    static DeviceOnce attr;
    if (attr.first_use_on_device())
        SGB_CUDA(cudaFuncSetAttribute(chain_backward_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageTimer t(ctx, ST_BLEND_BWD, s);
    ctx->launches += 1;
    chain_backward_mma_kernel<<<tiles, kMmaThreads, smem, s>>>(in.W, in.H, in.C, in.background, g.rec, colors, im.final_T,
                                                               dL_dpix, pv, dL_dmean2D, dL_dconic, dL_dopacity);
    SGB_LAUNCH_CHECK("chain_backward_mma_kernel", in.debug, s);
    return SGB_OK;
}

}  // namespace sgb
