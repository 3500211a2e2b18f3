// Small fp32 vector / 3x3 helpers shared by the per-Gaussian kernels.  The 3x3 product is spelled
// exactly like GLM's column-major operator* (third_party/glm/glm/detail/type_mat3x3.inl:486-519)
// so that nvcc contracts it like the reference's code.
#pragma once
#include "common.cuh"

namespace sgb {
namespace {

// auxiliary.h:22-39
__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                  -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                  0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                  -0.5900435899266435f};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 operator*(float s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }

// Column-major 3x3 (m[col][row]) whose product is spelled exactly like GLM's
// (type_mat3x3.inl operator*): Result[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2].
struct M3 {
    float m[3][3];
};
__device__ __forceinline__ M3 cols(float a, float b, float c, float d, float e, float f, float g, float h,
                                   float i) {
    M3 r;
    r.m[0][0] = a; r.m[0][1] = b; r.m[0][2] = c;
    r.m[1][0] = d; r.m[1][1] = e; r.m[1][2] = f;
    r.m[2][0] = g; r.m[2][1] = h; r.m[2][2] = i;
    return r;
}
__device__ __forceinline__ M3 operator*(const M3& A, const M3& B) {
    M3 R;
    R.m[0][0] = A.m[0][0] * B.m[0][0] + A.m[1][0] * B.m[0][1] + A.m[2][0] * B.m[0][2];
    R.m[0][1] = A.m[0][1] * B.m[0][0] + A.m[1][1] * B.m[0][1] + A.m[2][1] * B.m[0][2];
    R.m[0][2] = A.m[0][2] * B.m[0][0] + A.m[1][2] * B.m[0][1] + A.m[2][2] * B.m[0][2];
    R.m[1][0] = A.m[0][0] * B.m[1][0] + A.m[1][0] * B.m[1][1] + A.m[2][0] * B.m[1][2];
    R.m[1][1] = A.m[0][1] * B.m[1][0] + A.m[1][1] * B.m[1][1] + A.m[2][1] * B.m[1][2];
    R.m[1][2] = A.m[0][2] * B.m[1][0] + A.m[1][2] * B.m[1][1] + A.m[2][2] * B.m[1][2];
    R.m[2][0] = A.m[0][0] * B.m[2][0] + A.m[1][0] * B.m[2][1] + A.m[2][0] * B.m[2][2];
    R.m[2][1] = A.m[0][1] * B.m[2][0] + A.m[1][1] * B.m[2][1] + A.m[2][1] * B.m[2][2];
    R.m[2][2] = A.m[0][2] * B.m[2][0] + A.m[1][2] * B.m[2][1] + A.m[2][2] * B.m[2][2];
    return R;
}
__device__ __forceinline__ M3 transpose(const M3& A) {
    return cols(A.m[0][0], A.m[1][0], A.m[2][0], A.m[0][1], A.m[1][1], A.m[2][1], A.m[0][2], A.m[1][2],
                A.m[2][2]);
}

__device__ __forceinline__ float3 transformPoint4x3(const float3& p, const float* matrix) {
    float3 transformed = {
        matrix[0] * p.x + matrix[4] * p.y + matrix[8] * p.z + matrix[12],
        matrix[1] * p.x + matrix[5] * p.y + matrix[9] * p.z + matrix[13],
        matrix[2] * p.x + matrix[6] * p.y + matrix[10] * p.z + matrix[14],
    };
    return transformed;
}
__device__ __forceinline__ float4 transformPoint4x4(const float3& p, const float* matrix) {
    float4 transformed = {matrix[0] * p.x + matrix[4] * p.y + matrix[8] * p.z + matrix[12],
                          matrix[1] * p.x + matrix[5] * p.y + matrix[9] * p.z + matrix[13],
                          matrix[2] * p.x + matrix[6] * p.y + matrix[10] * p.z + matrix[14],
                          matrix[3] * p.x + matrix[7] * p.y + matrix[11] * p.z + matrix[15]};
    return transformed;
}
// auxiliary.h:41-44: double arithmetic, narrowed on return.
__device__ __forceinline__ float ndc2Pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }


}  // namespace
}  // namespace sgb
