// fp32 FMA issue-rate microbenchmark for sm_100a: scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
// Prints achieved TFLOP/s so that the blend kernels' FMA roof is a measured number.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
                 "mov.b64 rc, {%6, %7};\n\tfma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
                 : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

template <int ILP>
__global__ void k_ffma(float* out, float a, float b, int iters) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3f + i;
    float x = a + threadIdx.x * 1e-6f, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fmaf(acc[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma2(float* out, float a, float b, int iters) {
    float2 acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 x = make_float2(a + threadIdx.x * 1e-6f, a), y = make_float2(b, b * 0.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = ffma2(acc[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// blend-like inner loop: accumulators += broadcast-LDS.128 feature * per-thread weight
template <int CH>
__global__ void k_blendlike(float* out, int iters) {
    __shared__ __align__(16) float feat[64][CH];
    for (int e = threadIdx.x; e < 64 * CH; e += blockDim.x) (&feat[0][0])[e] = e * 1e-4f;
    __syncthreads();
    float2 acc[CH / 2];
#pragma unroll
    for (int i = 0; i < CH / 2; i++) acc[i] = make_float2(0.f, 0.f);
    float w = threadIdx.x * 1e-3f;
    for (int it = 0; it < iters; it++) {
        const int j = it & 63;
        const float2 w2 = make_float2(w, w);
#pragma unroll
        for (int k = 0; k < CH; k += 4) {
            float4 f = *reinterpret_cast<const float4*>(&feat[j][k]);
            acc[k / 2] = ffma2(make_float2(f.x, f.y), w2, acc[k / 2]);
            acc[k / 2 + 1] = ffma2(make_float2(f.z, f.w), w2, acc[k / 2 + 1]);
        }
        w += 1e-6f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH / 2; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// register-tiled outer product like blend_forward_tma_kernel's inner loop: per step 2+2 LDS.128,
// 8 px x 8 ch accumulators, 32 FFMA2
__global__ void __launch_bounds__(256, 2) k_tile8x8(float* out, int iters) {
    __shared__ __align__(16) float wsm[16][256];
    __shared__ __align__(16) float fsm[16][64];
    for (int e = threadIdx.x; e < 16 * 256; e += blockDim.x) (&wsm[0][0])[e] = (e % 97) * 1e-3f;
    for (int e = threadIdx.x; e < 16 * 64; e += blockDim.x) (&fsm[0][0])[e] = (e % 31) * 1e-2f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pg = lane >> 3, cg = lane & 7;
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[i][k] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
        const int e = it & 15;
        const float4 w0 = *reinterpret_cast<const float4*>(&wsm[e][warp * 32 + pg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&wsm[e][warp * 32 + pg * 8 + 4]);
        const float4 f0 = *reinterpret_cast<const float4*>(&fsm[e][cg * 8]);
        const float4 f1 = *reinterpret_cast<const float4*>(&fsm[e][cg * 8 + 4]);
        const float2 f[4] = {make_float2(f0.x, f0.y), make_float2(f0.z, f0.w), make_float2(f1.x, f1.y), make_float2(f1.z, f1.w)};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float2 w2 = make_float2(wv[i], wv[i]);
#pragma unroll
            for (int k = 0; k < 4; k++) acc[i][k] = ffma2(f[k], w2, acc[i][k]);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < 4; k++) s += acc[i][k].x + acc[i][k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// same with scalar FFMA (no packing, no duplication moves)
__global__ void __launch_bounds__(256, 2) k_tile8x8_scalar(float* out, int iters) {
    __shared__ __align__(16) float wsm[16][256];
    __shared__ __align__(16) float fsm[16][64];
    for (int e = threadIdx.x; e < 16 * 256; e += blockDim.x) (&wsm[0][0])[e] = (e % 97) * 1e-3f;
    for (int e = threadIdx.x; e < 16 * 64; e += blockDim.x) (&fsm[0][0])[e] = (e % 31) * 1e-2f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pg = lane >> 3, cg = lane & 7;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc[i][k] = 0.f;
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
        const int e = it & 15;
        const float4 w0 = *reinterpret_cast<const float4*>(&wsm[e][warp * 32 + pg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&wsm[e][warp * 32 + pg * 8 + 4]);
        const float4 f0 = *reinterpret_cast<const float4*>(&fsm[e][cg * 8]);
        const float4 f1 = *reinterpret_cast<const float4*>(&fsm[e][cg * 8 + 4]);
        const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int k = 0; k < 8; k++) acc[i][k] = fmaf(f[k], wv[i], acc[i][k]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < 8; k++) s += acc[i][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- legacy tensor-core path (mma.sync, compiles for sm_100a; SASS HMMA): what an error-compensated 3xTF32
// contraction could reach without tcgen05/TMEM.  Register-resident fragments, NT independent accumulator tiles.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NT, bool BF16>
__global__ void __launch_bounds__(256) k_mma(float* out, int iters) {
    float d[NT][4];
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
        for (int i = 0; i < 4; i++) d[t][i] = 0.f;
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i) & 0xffffe000u;
    b[0] = __float_as_uint(0.5f) & 0xffffe000u;
    b[1] = __float_as_uint(0.25f) & 0xffffe000u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int t = 0; t < NT; t++) {
            if (BF16) mma_bf16(d[t], a, b);
            else mma_tf32(d[t], a, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int t = 0; t < NT; t++) s += d[t][0] + d[t][1] + d[t][2] + d[t][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3xTF32 warp tile 32 px x 64 ch, K = 8 entries per step, operands from shared memory (fp32), hi/lo split in
// registers: the shape of the forward blend contraction per warp (2 m-tiles x 8 n-tiles x 3 products = 48 HMMA / step)
__global__ void __launch_bounds__(256, 2) k_tf32x3_tile(float* out, int iters) {
    __shared__ __align__(16) float wsm[16][264];   // [entry][px]  (pitch = 8 mod 32: conflict-free fragment loads)
    __shared__ __align__(16) float fsm[16][72];    // [entry][ch]
    for (int e = threadIdx.x; e < 16 * 264; e += blockDim.x) (&wsm[0][0])[e] = (e % 97) * 1e-3f;
    for (int e = threadIdx.x; e < 16 * 72; e += blockDim.x) (&fsm[0][0])[e] = (e % 31) * 1e-2f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gq = lane >> 2, tq = lane & 3;
    float d[2][8][4];
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
        for (int n = 0; n < 8; n++)
#pragma unroll
            for (int i = 0; i < 4; i++) d[m][n][i] = 0.f;
    auto split = [](float x, uint32_t& hi, uint32_t& lo) {
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
        const float r = x - __uint_as_float(hi);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
    };
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const int k0 = (it & 1) * 8;
        uint32_t ah[2][4], al[2][4], bh[8][2], bl[8][2];
#pragma unroll
        for (int m = 0; m < 2; m++) {   // A = W^T [px][entry]: a0 (row g, col t) a1 (row g+8) a2 (col t+4) a3
            const int px = warp * 32 + m * 16 + gq;
            split(wsm[k0 + tq][px], ah[m][0], al[m][0]);
            split(wsm[k0 + tq][px + 8], ah[m][1], al[m][1]);
            split(wsm[k0 + tq + 4][px], ah[m][2], al[m][2]);
            split(wsm[k0 + tq + 4][px + 8], ah[m][3], al[m][3]);
        }
#pragma unroll
        for (int n = 0; n < 8; n++) {   // B = F [entry][ch]: b0 (k t, n g) b1 (k t+4)
            split(fsm[k0 + tq][n * 8 + gq], bh[n][0], bl[n][0]);
            split(fsm[k0 + tq + 4][n * 8 + gq], bh[n][1], bl[n][1]);
        }
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int n = 0; n < 8; n++) {
                mma_tf32(d[m][n], al[m], bh[n]);
                mma_tf32(d[m][n], ah[m], bl[n]);
                mma_tf32(d[m][n], ah[m], bh[n]);
            }
    }
    float s = 0;
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
        for (int n = 0; n < 8; n++) s += d[m][n][0] + d[m][n][1] + d[m][n][2] + d[m][n][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 256, blocks = sms * 8, iters = 4096;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    printf("device %s, %d SMs\n", p.name, sms);
    {
        constexpr int ILP = 16;
        float ms = time_ms([&] { k_ffma<ILP><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
        double fl = 2.0 * ILP * (double)iters * blocks * threads;
        printf("FFMA   (scalar, ILP %d): %.3f ms  %.1f TFLOP/s\n", ILP, ms, fl / ms * 1e-9);
    }
    {
        constexpr int ILP = 16;
        float ms = time_ms([&] { k_ffma2<ILP><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
        double fl = 4.0 * ILP * (double)iters * blocks * threads;
        printf("FFMA2  (packed, ILP %d): %.3f ms  %.1f TFLOP/s\n", ILP, ms, fl / ms * 1e-9);
    }
    {
        constexpr int CH = 64;
        float ms = time_ms([&] { k_blendlike<CH><<<blocks, threads>>>(out, iters); });
        double fl = 2.0 * CH * (double)iters * blocks * threads;
        printf("blend-like (LDS.128 broadcast + FFMA2, CH %d): %.3f ms  %.1f TFLOP/s\n", CH, ms, fl / ms * 1e-9);
    }
    {
        constexpr int CH = 32;
        float ms = time_ms([&] { k_blendlike<CH><<<blocks, threads>>>(out, iters); });
        double fl = 2.0 * CH * (double)iters * blocks * threads;
        printf("blend-like (LDS.128 broadcast + FFMA2, CH %d): %.3f ms  %.1f TFLOP/s\n", CH, ms, fl / ms * 1e-9);
    }
    {
        const int it2 = 8192;
        float ms = time_ms([&] { k_tile8x8<<<sms * 2, 256>>>(out, it2); });
        double fl = 2.0 * 64 * (double)it2 * sms * 2 * 256;
        printf("8x8 register tile, FFMA2, 2 CTA/SM x 8 warps: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_tile8x8_scalar<<<sms * 2, 256>>>(out, it2); });
        printf("8x8 register tile, FFMA,  2 CTA/SM x 8 warps: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    {
        const int it3 = 4096;
        float ms = time_ms([&] { k_mma<8, false><<<sms * 4, 256>>>(out, it3); });
        double fl = 2.0 * 16 * 8 * 8 * 8.0 * (double)it3 * sms * 4 * 8;
        printf("mma.sync m16n8k8 tf32 (8 indep. tiles/warp, 4 CTA/SM x 8 warps): %.3f ms  %.1f TFLOP/s dense\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { k_mma<8, true><<<sms * 4, 256>>>(out, it3); });
        fl = 2.0 * 16 * 8 * 16 * 8.0 * (double)it3 * sms * 4 * 8;
        printf("mma.sync m16n8k16 bf16 (8 indep. tiles/warp, 4 CTA/SM x 8 warps): %.3f ms  %.1f TFLOP/s dense\n", ms, fl / ms * 1e-9);
        const int it4 = 8192;
        ms = time_ms([&] { k_tf32x3_tile<<<sms * 2, 256>>>(out, it4); });
        fl = 2.0 * 32 * 64 * 8 * (double)it4 * sms * 2 * 8;   // fp32-equivalent flops (one product per element pair)
        printf("3xTF32 warp tile 32x64xK8 from smem (split in registers), 2 CTA/SM x 8 warps: %.3f ms  %.1f TFLOP/s fp32-equivalent\n", ms, fl / ms * 1e-9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
