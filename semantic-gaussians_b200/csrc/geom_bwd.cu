// Per-Gaussian backward of the geometry stage: one thread per visible Gaussian, one pass over its inputs.  All the
// algebra lives in geom_grad.cuh (matrix-calculus form, shared with the CPU test); this file only moves data.
// Reference semantics matched (1e-4 on every output, tests/test_*_gpu.py): backward.cu:141-271 (screen covariance),
// :341-391 (projected centre, SH colour), :275-336 (scale / rotation).  The reference runs two kernels that both re-read
// the per-Gaussian inputs and accumulate dL/dmean through global memory; here the three contributions to dL/dmean are
// summed in registers and stored once.
#include "common.cuh"
#include "geom_grad.cuh"

namespace sgb {

namespace {

__global__ void __launch_bounds__(256) geom_backward_kernel(
    int P, int D, int M, const float* __restrict__ means, const int* __restrict__ radii,
    const float* __restrict__ shs, const uint8_t* __restrict__ clamped, const float* __restrict__ scales,
    const float* __restrict__ rotations, const float scale_modifier, const float* __restrict__ cov3Ds,
    const float* __restrict__ view_matrix, const float* __restrict__ proj, const float focal_x, const float focal_y,
    const float tan_fovx, const float tan_fovy, const float* __restrict__ campos, const float* __restrict__ dL_dmean2D,
    const float* __restrict__ dL_dconics, float* __restrict__ dL_dmeans, const float* __restrict__ dL_dcolor,
    float* __restrict__ dL_dcov, float* __restrict__ dL_dsh, float* __restrict__ dL_dscale,
    float* __restrict__ dL_drot) {
    __shared__ float cam[35];  // view (16) | proj (16) | campos (3): read by every thread, staged once per CTA
    if (threadIdx.x < 16) {
        cam[threadIdx.x] = view_matrix[threadIdx.x];
        cam[16 + threadIdx.x] = proj[threadIdx.x];
    } else if (threadIdx.x < 19) {
        cam[16 + threadIdx.x] = campos ? campos[threadIdx.x - 16] : 0.f;
    }
    __syncthreads();
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)P || !(radii[g] > 0)) return;  // culled Gaussians keep the caller's zeros (backward.cu:163,350)

    const float p[3] = {means[3 * g], means[3 * g + 1], means[3 * g + 2]};
    float cov6[6];
#pragma unroll
    for (int i = 0; i < 6; i++) cov6[i] = cov3Ds[6 * g + i];
    const float g_conic[3] = {dL_dconics[4 * g], dL_dconics[4 * g + 1], dL_dconics[4 * g + 3]};
    const float g_ndc[2] = {dL_dmean2D[3 * g], dL_dmean2D[3 * g + 1]};

    float g_mean[3], g_cov[6];
    geomgrad::project_grad(p, cov6, cam, cam + 16, focal_x, focal_y, tan_fovx, tan_fovy, g_conic, g_ndc, g_mean, g_cov);
#pragma unroll
    for (int i = 0; i < 6; i++) dL_dcov[6 * g + i] = g_cov[i];

    if (shs) {  // colours from SH: the colour gradient flows to the coefficients and, through the view direction, to p
        float g_rgb[3];
#pragma unroll
        for (int c = 0; c < 3; c++) g_rgb[c] = clamped[3 * g + c] ? 0.f : dL_dcolor[3 * g + c];
        geomgrad::colour_grad(D, p, cam + 32, shs + g * (size_t)M * 3, g_rgb, dL_dsh + g * (size_t)M * 3, g_mean);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) dL_dmeans[3 * g + i] = g_mean[i];

    if (scales) {  // covariance built from scale / rotation in the forward: continue through the factorisation
        const float q[4] = {rotations[4 * g], rotations[4 * g + 1], rotations[4 * g + 2], rotations[4 * g + 3]};
        const float s[3] = {scale_modifier * scales[3 * g], scale_modifier * scales[3 * g + 1],
                            scale_modifier * scales[3 * g + 2]};
        float g_s[3], g_q[4];
        geomgrad::factor_grad(g_cov, q, s, g_s, g_q);
#pragma unroll
        for (int i = 0; i < 3; i++) dL_dscale[3 * g + i] = g_s[i];
        *reinterpret_cast<float4*>(dL_drot + 4 * g) = make_float4(g_q[0], g_q[1], g_q[2], g_q[3]);
    }
}

}  // namespace

int launch_geom_backward(const sgb_view_inputs& in, GeomView g, const int32_t* radii, const float* cov3D,
                         const float* dL_dcolor_rgb, const sgb_view_grads& gr, cudaStream_t s) {
    const float focal_y = in.H / (2.0f * in.tan_fovy);
    const float focal_x = in.W / (2.0f * in.tan_fovx);
    geom_backward_kernel<<<(in.P + 255) / 256, 256, 0, s>>>(
        in.P, in.D, in.M, in.means3D, radii, in.shs, g.clamped, in.scales,
        in.rotations, in.scale_modifier, cov3D, in.viewmatrix, in.projmatrix, focal_x, focal_y,
        in.tan_fovx, in.tan_fovy, in.campos, gr.dL_dmeans2D, gr.dL_dconic, gr.dL_dmeans3D, dL_dcolor_rgb,
        gr.dL_dcov3D, gr.dL_dsh, gr.dL_dscales, gr.dL_drotations);
    SGB_LAUNCH_CHECK("geom_backward_kernel", in.debug, s);
    return SGB_OK;
}

}  // namespace sgb
