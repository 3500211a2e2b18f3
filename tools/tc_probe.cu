// tcgen05 probe (sm_100a): validates, on real hardware, the shared-memory operand layouts / descriptor fields that the
// opt-in tensor-core forward (SGB_BLEND_MMA=1) relies on, against a CPU product of exactly representable integers.
//   kind::tf32, cta_group::1, M = 128, N = 64, K = 8 per instruction, no swizzle.
// Layout hypotheses tried in one run (each prints its max abs error; 0 = the hypothesis holds):
//   A MN-major: 16-byte chunk (4 consecutive M of one k) at  base + (m/4)*SBO + (k%8)*16            [+ (k/8)*LBO]
//   B MN-major: same with N;   B K-major: chunk (4 consecutive k of one n) at base + (n/8)*SBO + (n%8)*16 + (k/4)*LBO
// and for each the two possible assignments of (SBO, LBO) to the descriptor's stride / leading fields.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe.bin tools/tc_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lead_bytes, uint32_t stride_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lead_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((stride_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

struct Params {
    int a_mn_major, b_mn_major;  // 1 = MN-major, 0 = K-major
    int swap_fields;             // 0: (lead = LBO, stride = SBO); 1: swapped
    int ksteps;                  // number of K = 8 MMAs accumulated
};

constexpr int M = 128, N = 64, KMAX = 16;

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A /*[M][KMAX]*/,
                                                    const float* __restrict__ B /*[N][KMAX]*/, float* __restrict__ D,
                                                    Params p) {
    __shared__ __align__(1024) float sA[M * KMAX];
    __shared__ __align__(1024) float sB[N * KMAX];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = 8 * p.ksteps;
    // ---- place the operands.  One "K block" = 8 consecutive k.
    constexpr uint32_t A_SBO = 128, B_SBO_MN = 128;            // adjacent 128-byte core matrices along M / N
    const uint32_t A_LBO = (M / 4) * 128;                      // next K block of A: after all M groups
    const uint32_t B_LBO_MN = (N / 4) * 128;
    const uint32_t B_SBO_K = 256, B_LBO_K = 128;               // K-major B: [n/8][k/4][n%8][k%4]
    for (int e = tid; e < M * K; e += 128) {
        const int m = e / K, k = e % K;
        const float v = A[m * KMAX + k];
        uint32_t off;
        if (p.a_mn_major) off = (m / 4) * A_SBO + (k / 8) * A_LBO + (k % 8) * 16 + (m % 4) * 4;
        else off = (m / 8) * 256 + ((k % 8) / 4) * 128 + (m % 8) * 16 + (k % 4) * 4 + (k / 8) * (M / 8) * 256;
        *reinterpret_cast<float*>(reinterpret_cast<char*>(sA) + off) = v;
    }
    for (int e = tid; e < N * K; e += 128) {
        const int n = e / K, k = e % K;
        const float v = B[n * KMAX + k];
        uint32_t off;
        if (p.b_mn_major) off = (n / 4) * B_SBO_MN + (k / 8) * B_LBO_MN + (k % 8) * 16 + (n % 4) * 4;
        else off = (n / 8) * B_SBO_K + ((k % 8) / 4) * B_LBO_K + (n % 8) * 16 + (k % 4) * 4 + (k / 8) * (N / 8) * 256;
        *reinterpret_cast<float*>(reinterpret_cast<char*>(sB) + off) = v;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    // instruction descriptor: D f32, A/B tf32, majors, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn_major << 15) |
                           ((uint32_t)p.b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) {
        for (int ks = 0; ks < p.ksteps; ks++) {
            uint32_t a_lead, a_stride, b_lead, b_stride, a_off, b_off;
            if (p.a_mn_major) { a_lead = A_LBO; a_stride = A_SBO; a_off = ks * A_LBO; }
            else { a_lead = 128; a_stride = 256; a_off = ks * (M / 8) * 256; }
            if (p.b_mn_major) { b_lead = B_LBO_MN; b_stride = B_SBO_MN; b_off = ks * B_LBO_MN; }
            else { b_lead = B_LBO_K; b_stride = B_SBO_K; b_off = ks * (N / 8) * 256; }
            if (p.swap_fields) {
                uint32_t t = a_lead; a_lead = a_stride; a_stride = t;
                t = b_lead; b_lead = b_stride; b_stride = t;
            }
            const uint64_t da = make_desc(smem_u32(sA) + a_off, a_lead, a_stride);
            const uint64_t db = make_desc(smem_u32(sB) + b_off, b_lead, b_stride);
            const uint32_t accum = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everybody waits for the MMAs
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: thread t of warp w owns TMEM lane 32 w + t = row m of D
    uint32_t r[32];
    for (int c0 = 0; c0 < N; c0 += 32) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    static float hA[M * KMAX], hB[N * KMAX], hD[M * N], ref[M * N];
    for (int m = 0; m < M; m++)
        for (int k = 0; k < KMAX; k++) hA[m * KMAX + k] = (float)(((m * 3 + k * 7) % 13) - 6);
    for (int n = 0; n < N; n++)
        for (int k = 0; k < KMAX; k++) hB[n * KMAX + k] = (float)(((n * 5 + k * 11) % 9) - 4);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    int nok = 0;
    for (int ks = 1; ks <= 2; ks++)
        for (int amn = 1; amn >= 0; amn--)
            for (int bmn = 1; bmn >= 0; bmn--)
                for (int sw = 0; sw < 2; sw++) {
                    Params p{amn, bmn, sw, ks};
                    for (int m = 0; m < M; m++)
                        for (int n = 0; n < N; n++) {
                            float s = 0;
                            for (int k = 0; k < 8 * ks; k++) s += hA[m * KMAX + k] * hB[n * KMAX + k];
                            ref[m * N + n] = s;
                        }
                    cudaMemset(dD, 0xFF, sizeof hD);
                    probe_kernel<<<1, 128>>>(dA, dB, dD, p);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("ksteps %d A %s B %s swap %d: CUDA error %s\n", ks, amn ? "MN" : "K ", bmn ? "MN" : "K ", sw,
                               cudaGetErrorString(e));
                        return 1;
                    }
                    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
                    double err = 0;
                    for (int i = 0; i < M * N; i++) {
                        double d = (double)hD[i] - ref[i];
                        if (d != d) d = 1e30;
                        if (d < 0) d = -d;
                        if (d > err) err = d;
                    }
                    printf("ksteps %d  A %s-major  B %s-major  fields %s : max abs err %g %s\n", ks, amn ? "MN" : "K ",
                           bmn ? "MN" : "K ", sw ? "swapped" : "as derived", err, err == 0 ? "OK" : "");
                    nok += err == 0;
                }
    printf("%d configurations exact\n", nok);
    return 0;
}
