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
constexpr uint32_t kABlk = 16 * 256;   // bytes of one (K block, pixel half) operand block: 16 row groups x 256 B
constexpr uint32_t kBBlk = (kNch / 8) * 256;
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
    extern __shared__ __align__(1024) unsigned char smem_raw[];
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
    //     (r / 8) * 256 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4      -> descriptor LBO = 128 B, SBO = 256 B.
    // A thread holds 4 consecutive rows of ONE entry (a 16-byte piece of a weight / feature row), i.e. 4 scalar stores.
    auto put4 = [&](unsigned char* blk, int r0, float4 v) {
        const uint32_t col = (uint32_t)(e8 >> 2) * 128u + (uint32_t)(e8 & 3) * 4u;
        const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = r0 + j;
            *reinterpret_cast<float*>(blk + (uint32_t)(r >> 3) * 256u + (uint32_t)(r & 7) * 16u + col) = x[j];
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
                const uint64_t dbh = umma_desc(bh, 128, 256), dbl = umma_desc(bl, 128, 256);
                for (int mh = 0; mh < 2; mh++) {
                    const uint32_t ah = sbase + (uint32_t)(kb * 2 + mh) * kABlk, al = ah + kStageA;
                    const uint64_t dah = umma_desc(ah, 128, 256), dal = umma_desc(al, 128, 256);
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

}  // namespace

bool blend_mma_enabled() {
    static const bool on = [] { const char* e = getenv("SGB_BLEND_MMA"); return e && e[0] == '1'; }();
    return on;
}

int launch_forward_mma(sgb_ctx* ctx, const sgb_view_inputs& in, ImgView im, const float* colors, float* out_color,
                       const PoolView& pv, cudaStream_t s) {
    const int tiles = ((in.W + SGB_TILE - 1) / SGB_TILE) * ((in.H + SGB_TILE - 1) / SGB_TILE);
    const int slices = (in.C + kNch - 1) / kNch;
    const size_t smem = (size_t)kStages * kStageBytes + 1024;
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

}  // namespace sgb
