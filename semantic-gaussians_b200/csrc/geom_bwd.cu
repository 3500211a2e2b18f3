// Per-Gaussian backward: conic -> cov2D -> cov3D / mean (reference backward.cu:141-271), mean2D ->
// mean3D through the projection (:341-383), SH colour backward (:20-136) and cov3D -> scale /
// rotation (:275-336).  The reference runs two kernels (computeCov2DCUDA, preprocessCUDA<C>) that
// both re-read the same per-Gaussian inputs; here it is one pass, one thread per Gaussian.
#include "common.cuh"
#include "linalg.cuh"

namespace sgb {

namespace {

// auxiliary.h:107-117
__device__ __forceinline__ float3 dnormvdv(float3 v, float3 dv) {
    float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 operator*(V3 v, float s) { return {v.x * s, v.y * s, v.z * s}; }

// backward.cu:20-136
__device__ void sh_backward(int idx, int deg, int max_coeffs, const float3 mean, const float* campos,
                            const float* shs, const uint8_t* clamped, const float* dL_dcolor,
                            float* dL_dmeans, float* dL_dshs) {
    V3 dir_orig = {mean.x - campos[0], mean.y - campos[1], mean.z - campos[2]};
    float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    V3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};
    const V3* sh = reinterpret_cast<const V3*>(shs) + (size_t)idx * max_coeffs;
    V3 dL_dRGB = {dL_dcolor[3 * idx + 0], dL_dcolor[3 * idx + 1], dL_dcolor[3 * idx + 2]};
    dL_dRGB.x *= clamped[3 * idx + 0] ? 0 : 1;
    dL_dRGB.y *= clamped[3 * idx + 1] ? 0 : 1;
    dL_dRGB.z *= clamped[3 * idx + 2] ? 0 : 1;
    V3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
    float x = dir.x, y = dir.y, z = dir.z;
    V3* dL_dsh = reinterpret_cast<V3*>(dL_dshs) + (size_t)idx * max_coeffs;
    float dRGBdsh0 = SH_C0;
    dL_dsh[0] = dRGBdsh0 * dL_dRGB;
    if (deg > 0) {
        float dRGBdsh1 = -SH_C1 * y;
        float dRGBdsh2 = SH_C1 * z;
        float dRGBdsh3 = -SH_C1 * x;
        dL_dsh[1] = dRGBdsh1 * dL_dRGB;
        dL_dsh[2] = dRGBdsh2 * dL_dRGB;
        dL_dsh[3] = dRGBdsh3 * dL_dRGB;
        dRGBdx = -SH_C1 * sh[3];
        dRGBdy = -SH_C1 * sh[1];
        dRGBdz = SH_C1 * sh[2];
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, yz = y * z, xz = x * z;
            float dRGBdsh4 = SH_C2[0] * xy;
            float dRGBdsh5 = SH_C2[1] * yz;
            float dRGBdsh6 = SH_C2[2] * (2.f * zz - xx - yy);
            float dRGBdsh7 = SH_C2[3] * xz;
            float dRGBdsh8 = SH_C2[4] * (xx - yy);
            dL_dsh[4] = dRGBdsh4 * dL_dRGB;
            dL_dsh[5] = dRGBdsh5 * dL_dRGB;
            dL_dsh[6] = dRGBdsh6 * dL_dRGB;
            dL_dsh[7] = dRGBdsh7 * dL_dRGB;
            dL_dsh[8] = dRGBdsh8 * dL_dRGB;
            dRGBdx = dRGBdx + (SH_C2[0] * y * sh[4] + SH_C2[2] * 2.f * -x * sh[6] + SH_C2[3] * z * sh[7] +
                               SH_C2[4] * 2.f * x * sh[8]);
            dRGBdy = dRGBdy + (SH_C2[0] * x * sh[4] + SH_C2[1] * z * sh[5] + SH_C2[2] * 2.f * -y * sh[6] +
                               SH_C2[4] * 2.f * -y * sh[8]);
            dRGBdz = dRGBdz + (SH_C2[1] * y * sh[5] + SH_C2[2] * 2.f * 2.f * z * sh[6] + SH_C2[3] * x * sh[7]);
            if (deg > 2) {
                float dRGBdsh9 = SH_C3[0] * y * (3.f * xx - yy);
                float dRGBdsh10 = SH_C3[1] * xy * z;
                float dRGBdsh11 = SH_C3[2] * y * (4.f * zz - xx - yy);
                float dRGBdsh12 = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                float dRGBdsh13 = SH_C3[4] * x * (4.f * zz - xx - yy);
                float dRGBdsh14 = SH_C3[5] * z * (xx - yy);
                float dRGBdsh15 = SH_C3[6] * x * (xx - 3.f * yy);
                dL_dsh[9] = dRGBdsh9 * dL_dRGB;
                dL_dsh[10] = dRGBdsh10 * dL_dRGB;
                dL_dsh[11] = dRGBdsh11 * dL_dRGB;
                dL_dsh[12] = dRGBdsh12 * dL_dRGB;
                dL_dsh[13] = dRGBdsh13 * dL_dRGB;
                dL_dsh[14] = dRGBdsh14 * dL_dRGB;
                dL_dsh[15] = dRGBdsh15 * dL_dRGB;
                dRGBdx = dRGBdx + (sh[9] * (SH_C3[0] * 3.f * 2.f * xy) + sh[10] * (SH_C3[1] * yz) +
                                   sh[11] * (SH_C3[2] * -2.f * xy) + sh[12] * (SH_C3[3] * -3.f * 2.f * xz) +
                                   sh[13] * (SH_C3[4] * (-3.f * xx + 4.f * zz - yy)) + sh[14] * (SH_C3[5] * 2.f * xz) +
                                   sh[15] * (SH_C3[6] * 3.f * (xx - yy)));
                dRGBdy = dRGBdy + (sh[9] * (SH_C3[0] * 3.f * (xx - yy)) + sh[10] * (SH_C3[1] * xz) +
                                   sh[11] * (SH_C3[2] * (-3.f * yy + 4.f * zz - xx)) +
                                   sh[12] * (SH_C3[3] * -3.f * 2.f * yz) + sh[13] * (SH_C3[4] * -2.f * xy) +
                                   sh[14] * (SH_C3[5] * -2.f * yz) + sh[15] * (SH_C3[6] * -3.f * 2.f * xy));
                dRGBdz = dRGBdz + (sh[10] * (SH_C3[1] * xy) + sh[11] * (SH_C3[2] * 4.f * 2.f * yz) +
                                   sh[12] * (SH_C3[3] * 3.f * (2.f * zz - xx - yy)) +
                                   sh[13] * (SH_C3[4] * 4.f * 2.f * xz) + sh[14] * (SH_C3[5] * (xx - yy)));
            }
        }
    }
    float3 dL_ddir = {dot3(dRGBdx, dL_dRGB), dot3(dRGBdy, dL_dRGB), dot3(dRGBdz, dL_dRGB)};
    float3 dL_dmean = dnormvdv(float3{dir_orig.x, dir_orig.y, dir_orig.z}, dL_ddir);
    dL_dmeans[0] += dL_dmean.x;
    dL_dmeans[1] += dL_dmean.y;
    dL_dmeans[2] += dL_dmean.z;
}

__global__ void __launch_bounds__(256) geom_backward_kernel(
    int P, int D, int M, const float3* __restrict__ means, const int* __restrict__ radii,
    const float* __restrict__ shs, const uint8_t* __restrict__ clamped, const float3* __restrict__ scales,
    const float4* __restrict__ rotations, const float scale_modifier, const float* __restrict__ cov3Ds,
    const float* __restrict__ view_matrix, const float* __restrict__ proj, const float h_x, float h_y,
    const float tan_fovx, float tan_fovy, const float* __restrict__ campos, const float* __restrict__ dL_dmean2D,
    const float* __restrict__ dL_dconics, float* __restrict__ dL_dmeans, const float* __restrict__ dL_dcolor,
    float* __restrict__ dL_dcov, float* __restrict__ dL_dsh, float* __restrict__ dL_dscale,
    float* __restrict__ dL_drot) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P || !(radii[idx] > 0)) return;

    // ---------------- backward.cu:156-270 (computeCov2DCUDA)
    const float* cov3D = cov3Ds + 6 * (size_t)idx;
    float3 mean = means[idx];
    float3 dL_dconic = {dL_dconics[4 * (size_t)idx], dL_dconics[4 * (size_t)idx + 1], dL_dconics[4 * (size_t)idx + 3]};
    float3 t = transformPoint4x3(mean, view_matrix);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = min(limx, max(-limx, txtz)) * t.z;
    t.y = min(limy, max(-limy, tytz)) * t.z;
    const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
    const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;

    M3 J = cols(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z), 0.0f, h_y / t.z, -(h_y * t.y) / (t.z * t.z), 0, 0, 0);
    M3 W = cols(view_matrix[0], view_matrix[4], view_matrix[8], view_matrix[1], view_matrix[5], view_matrix[9],
                view_matrix[2], view_matrix[6], view_matrix[10]);
    M3 Vrk = cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    M3 T = W * J;
    M3 cov2D = transpose(T) * transpose(Vrk) * T;
    float a = cov2D.m[0][0] += 0.3f;
    float b = cov2D.m[0][1];
    float c = cov2D.m[1][1] += 0.3f;
    float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dcv[6];
#define TT(i, j) T.m[i][j]
#define VV(i, j) Vrk.m[i][j]
    if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
        dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
        dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);
        dcv[0] = (TT(0, 0) * TT(0, 0) * dL_da + TT(0, 0) * TT(1, 0) * dL_db + TT(1, 0) * TT(1, 0) * dL_dc);
        dcv[3] = (TT(0, 1) * TT(0, 1) * dL_da + TT(0, 1) * TT(1, 1) * dL_db + TT(1, 1) * TT(1, 1) * dL_dc);
        dcv[5] = (TT(0, 2) * TT(0, 2) * dL_da + TT(0, 2) * TT(1, 2) * dL_db + TT(1, 2) * TT(1, 2) * dL_dc);
        dcv[1] = 2 * TT(0, 0) * TT(0, 1) * dL_da + (TT(0, 0) * TT(1, 1) + TT(0, 1) * TT(1, 0)) * dL_db +
                 2 * TT(1, 0) * TT(1, 1) * dL_dc;
        dcv[2] = 2 * TT(0, 0) * TT(0, 2) * dL_da + (TT(0, 0) * TT(1, 2) + TT(0, 2) * TT(1, 0)) * dL_db +
                 2 * TT(1, 0) * TT(1, 2) * dL_dc;
        dcv[4] = 2 * TT(0, 2) * TT(0, 1) * dL_da + (TT(0, 1) * TT(1, 2) + TT(0, 2) * TT(1, 1)) * dL_db +
                 2 * TT(1, 1) * TT(1, 2) * dL_dc;
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++) dcv[i] = 0;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) dL_dcov[6 * (size_t)idx + i] = dcv[i];

    float dL_dT00 = 2 * (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_da +
                    (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_db;
    float dL_dT01 = 2 * (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_da +
                    (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_db;
    float dL_dT02 = 2 * (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_da +
                    (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_db;
    float dL_dT10 = 2 * (TT(1, 0) * VV(0, 0) + TT(1, 1) * VV(0, 1) + TT(1, 2) * VV(0, 2)) * dL_dc +
                    (TT(0, 0) * VV(0, 0) + TT(0, 1) * VV(0, 1) + TT(0, 2) * VV(0, 2)) * dL_db;
    float dL_dT11 = 2 * (TT(1, 0) * VV(1, 0) + TT(1, 1) * VV(1, 1) + TT(1, 2) * VV(1, 2)) * dL_dc +
                    (TT(0, 0) * VV(1, 0) + TT(0, 1) * VV(1, 1) + TT(0, 2) * VV(1, 2)) * dL_db;
    float dL_dT12 = 2 * (TT(1, 0) * VV(2, 0) + TT(1, 1) * VV(2, 1) + TT(1, 2) * VV(2, 2)) * dL_dc +
                    (TT(0, 0) * VV(2, 0) + TT(0, 1) * VV(2, 1) + TT(0, 2) * VV(2, 2)) * dL_db;
#undef TT
#undef VV
    float dL_dJ00 = W.m[0][0] * dL_dT00 + W.m[0][1] * dL_dT01 + W.m[0][2] * dL_dT02;
    float dL_dJ02 = W.m[2][0] * dL_dT00 + W.m[2][1] * dL_dT01 + W.m[2][2] * dL_dT02;
    float dL_dJ11 = W.m[1][0] * dL_dT10 + W.m[1][1] * dL_dT11 + W.m[1][2] * dL_dT12;
    float dL_dJ12 = W.m[2][0] * dL_dT10 + W.m[2][1] * dL_dT11 + W.m[2][2] * dL_dT12;
    float tz = 1.f / t.z;
    float tz2 = tz * tz;
    float tz3 = tz2 * tz;
    float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
    float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
    float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                   (2 * h_y * t.y) * tz3 * dL_dJ12;
    // transformVec4x3Transpose (auxiliary.h:89-97); this part overwrites (backward.cu:270)
    float dmean[3] = {view_matrix[0] * dL_dtx + view_matrix[1] * dL_dty + view_matrix[2] * dL_dtz,
                      view_matrix[4] * dL_dtx + view_matrix[5] * dL_dty + view_matrix[6] * dL_dtz,
                      view_matrix[8] * dL_dtx + view_matrix[9] * dL_dty + view_matrix[10] * dL_dtz};

    // ---------------- backward.cu:365-382 (mean2D -> mean3D)
    float3 m = mean;
    float4 m_hom = transformPoint4x4(m, proj);
    float m_w = 1.0f / (m_hom.w + 0.0000001f);
    float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
    float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
    const float g2x = dL_dmean2D[3 * (size_t)idx], g2y = dL_dmean2D[3 * (size_t)idx + 1];
    dmean[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
    dmean[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
    dmean[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;

    if (shs) sh_backward(idx, D, M, mean, campos, shs, clamped, dL_dcolor, dmean, dL_dsh);
    dL_dmeans[3 * (size_t)idx + 0] = dmean[0];
    dL_dmeans[3 * (size_t)idx + 1] = dmean[1];
    dL_dmeans[3 * (size_t)idx + 2] = dmean[2];

    // ---------------- backward.cu:275-336 (cov3D -> scale, rotation)
    if (scales) {
        float4 q = rotations[idx];
        float r = q.x, x = q.y, y = q.z, z = q.w;
        M3 R = cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                    2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                    2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
        M3 S = cols(1.0f, 0.f, 0.f, 0.f, 1.0f, 0.f, 0.f, 0.f, 1.0f);
        float3 sc = scales[idx];
        float s[3] = {scale_modifier * sc.x, scale_modifier * sc.y, scale_modifier * sc.z};
        S.m[0][0] = s[0];
        S.m[1][1] = s[1];
        S.m[2][2] = s[2];
        M3 Mm = S * R;
        M3 dL_dSigma = cols(dcv[0], 0.5f * dcv[1], 0.5f * dcv[2], 0.5f * dcv[1], dcv[3], 0.5f * dcv[4],
                            0.5f * dcv[2], 0.5f * dcv[4], dcv[5]);
        M3 twoM;
#pragma unroll
        for (int cc = 0; cc < 3; cc++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) twoM.m[cc][rr] = 2.0f * Mm.m[cc][rr];
        M3 dL_dM = twoM * dL_dSigma;
        M3 Rt = transpose(R);
        M3 dL_dMt = transpose(dL_dM);
        float ds[3];
#pragma unroll
        for (int k = 0; k < 3; k++)
            ds[k] = Rt.m[k][0] * dL_dMt.m[k][0] + Rt.m[k][1] * dL_dMt.m[k][1] + Rt.m[k][2] * dL_dMt.m[k][2];
        dL_dscale[3 * (size_t)idx + 0] = ds[0];
        dL_dscale[3 * (size_t)idx + 1] = ds[1];
        dL_dscale[3 * (size_t)idx + 2] = ds[2];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int rr = 0; rr < 3; rr++) dL_dMt.m[k][rr] *= s[k];
#define Q(i, j) dL_dMt.m[i][j]
        float4 dq;
        dq.x = 2 * z * (Q(0, 1) - Q(1, 0)) + 2 * y * (Q(2, 0) - Q(0, 2)) + 2 * x * (Q(1, 2) - Q(2, 1));
        dq.y = 2 * y * (Q(1, 0) + Q(0, 1)) + 2 * z * (Q(2, 0) + Q(0, 2)) + 2 * r * (Q(1, 2) - Q(2, 1)) -
               4 * x * (Q(2, 2) + Q(1, 1));
        dq.z = 2 * x * (Q(1, 0) + Q(0, 1)) + 2 * r * (Q(2, 0) - Q(0, 2)) + 2 * z * (Q(1, 2) + Q(2, 1)) -
               4 * y * (Q(2, 2) + Q(0, 0));
        dq.w = 2 * r * (Q(0, 1) - Q(1, 0)) + 2 * x * (Q(2, 0) + Q(0, 2)) + 2 * y * (Q(1, 2) + Q(2, 1)) -
               4 * z * (Q(1, 1) + Q(0, 0));
#undef Q
        *reinterpret_cast<float4*>(dL_drot + 4 * (size_t)idx) = dq;
    }
}

}  // namespace

int launch_geom_backward(const sgb_view_inputs& in, GeomView g, const int32_t* radii, const float* cov3D,
                         const float* dL_dcolor_rgb, const sgb_view_grads& gr, cudaStream_t s) {
    const float focal_y = in.H / (2.0f * in.tan_fovy);
    const float focal_x = in.W / (2.0f * in.tan_fovx);
    geom_backward_kernel<<<(in.P + 255) / 256, 256, 0, s>>>(
        in.P, in.D, in.M, (const float3*)in.means3D, radii, in.shs, g.clamped, (const float3*)in.scales,
        (const float4*)in.rotations, in.scale_modifier, cov3D, in.viewmatrix, in.projmatrix, focal_x, focal_y,
        in.tan_fovx, in.tan_fovy, in.campos, gr.dL_dmeans2D, gr.dL_dconic, gr.dL_dmeans3D, dL_dcolor_rgb,
        gr.dL_dcov3D, gr.dL_dsh, gr.dL_dscales, gr.dL_drotations);
    SGB_LAUNCH_CHECK("geom_backward_kernel", in.debug, s);
    return SGB_OK;
}

}  // namespace sgb
