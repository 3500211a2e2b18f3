// fp32 FMA issue-rate microbenchmark for sm_100a: scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
// Prints achieved TFLOP/s so that the blend kernels' FMA roof is a measured number.
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
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
