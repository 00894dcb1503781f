// Device math for the world -> screen projection of one Beta primitive and its VJP.
//
// Semantics follow the reference kernels fully_fused_projection_fwd.cu:43-177 / _bwd.cu:69-201 and the helpers
// in utils.cuh:252-502 (persp_proj, pos/covar_world_to_cam, inverse, add_blur).  The reference expresses the
// 3x3 products through glm (column-major, `tmp = A0*b.x; tmp += A1*b.y; tmp += A2*b.z`); here everything is
// plain row-major float arithmetic written in the same association order so that the same compiler, with the
// same --use_fast_math contraction rules, rounds the integer-deciding quantities (depth, radius, mean2d)
// identically.  No glm, no local-memory matrices: all values live in registers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ubs {

struct Cam {
    float r[9];  // row-major world->camera rotation
    float t[3];
    float fx, fy, cx, cy;
};

__device__ __forceinline__ Cam load_cam(const float *__restrict__ viewmat, const float *__restrict__ K) {
    Cam c;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) c.r[i * 3 + j] = viewmat[i * 4 + j];
        c.t[i] = viewmat[i * 4 + 3];
    }
    c.fx = K[0];
    c.cx = K[2];
    c.fy = K[4];
    c.cy = K[5];
    return c;
}

// a0*b0 + a1*b1 + a2*b2 in the reference's left-to-right association.
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    float tmp = a0 * b0;
    tmp += a1 * b1;
    tmp += a2 * b2;
    return tmp;
}

// p_c = R p + t   (utils.cuh:374-383)
__device__ __forceinline__ void world_to_cam_point(const Cam &c, const float p[3], float pc[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) pc[i] = dot3(c.r[i * 3 + 0], p[0], c.r[i * 3 + 1], p[1], c.r[i * 3 + 2], p[2]) + c.t[i];
}

// S is the symmetric world covariance as 6 floats (xx,xy,xz,yy,yz,zz); Sc = R S R^T full 3x3 row-major
// (utils.cuh:402-411).
__device__ __forceinline__ void world_to_cam_covar(const Cam &c, const float s6[6], float Sc[9]) {
    const float S[9] = {s6[0], s6[1], s6[2], s6[1], s6[3], s6[4], s6[2], s6[4], s6[5]};
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            M[i * 3 + j] = dot3(c.r[i * 3 + 0], S[0 * 3 + j], c.r[i * 3 + 1], S[1 * 3 + j], c.r[i * 3 + 2], S[2 * 3 + j]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Sc[i * 3 + j] = dot3(M[i * 3 + 0], c.r[j * 3 + 0], M[i * 3 + 1], c.r[j * 3 + 1], M[i * 3 + 2], c.r[j * 3 + 2]);
}

struct PerspJ {
    float j00, j11, j02, j12;  // J = [[j00, 0, j02], [0, j11, j12]]
    float rz, rz2, tx, ty;
    bool x_in, y_in;  // fov clamp inactive (needed by the VJP)
};

__device__ __forceinline__ PerspJ persp_jacobian(const Cam &c, const float pc[3], uint32_t width, uint32_t height) {
    PerspJ o;
    const float x = pc[0], y = pc[1], z = pc[2];
    const float tan_fovx = 0.5f * width / c.fx;
    const float tan_fovy = 0.5f * height / c.fy;
    const float lim_x_pos = (width - c.cx) / c.fx + 0.3f * tan_fovx;
    const float lim_x_neg = c.cx / c.fx + 0.3f * tan_fovx;
    const float lim_y_pos = (height - c.cy) / c.fy + 0.3f * tan_fovy;
    const float lim_y_neg = c.cy / c.fy + 0.3f * tan_fovy;
    const float rz = 1.f / z;
    const float rz2 = rz * rz;
    const float xr = x * rz, yr = y * rz;
    const float tx = z * fminf(lim_x_pos, fmaxf(-lim_x_neg, xr));
    const float ty = z * fminf(lim_y_pos, fmaxf(-lim_y_neg, yr));
    o.x_in = (xr <= lim_x_pos) && (xr >= -lim_x_neg);
    o.y_in = (yr <= lim_y_pos) && (yr >= -lim_y_neg);
    o.j00 = c.fx * rz;
    o.j11 = c.fy * rz;
    o.j02 = -c.fx * tx * rz2;
    o.j12 = -c.fy * ty * rz2;
    o.rz = rz;
    o.rz2 = rz2;
    o.tx = tx;
    o.ty = ty;
    return o;
}

// cov2d = J Sc J^T (2x2, stored c00,c01,c10,c11), mean2d = (fx x/z + cx, fy y/z + cy)   (utils.cuh:252-292)
__device__ __forceinline__ void persp_project(const Cam &c, const float pc[3], const float Sc[9], const PerspJ &J,
                                              float cov2d[4], float mean2d[2]) {
    float A[6];  // A = J Sc   (2x3)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        A[0 * 3 + k] = dot3(J.j00, Sc[0 * 3 + k], 0.f, Sc[1 * 3 + k], J.j02, Sc[2 * 3 + k]);
        A[1 * 3 + k] = dot3(0.f, Sc[0 * 3 + k], J.j11, Sc[1 * 3 + k], J.j12, Sc[2 * 3 + k]);
    }
    cov2d[0] = dot3(A[0], J.j00, A[1], 0.f, A[2], J.j02);
    cov2d[2] = dot3(A[3], J.j00, A[4], 0.f, A[5], J.j02);  // row 1, col 0
    cov2d[1] = dot3(A[0], 0.f, A[1], J.j11, A[2], J.j12);  // row 0, col 1
    cov2d[3] = dot3(A[3], 0.f, A[4], J.j11, A[5], J.j12);
    mean2d[0] = c.fx * pc[0] * J.rz + c.cx;
    mean2d[1] = c.fy * pc[1] * J.rz + c.cy;
}

struct Splat2D {
    float mean2d[2];
    float depth;
    float conic[3];
    float compensation;
    int32_t radius;  // 0 => culled
};

// Full per-(camera, primitive) forward: returns radius 0 when culled (fully_fused_projection_fwd.cu:70-164).
__device__ __forceinline__ Splat2D project_splat(const Cam &c, const float p[3], const float s6[6], uint32_t width,
                                                 uint32_t height, float eps2d, float near_plane, float far_plane,
                                                 float radius_clip) {
    Splat2D o;
    o.radius = 0;
    o.mean2d[0] = o.mean2d[1] = o.depth = 0.f;
    o.conic[0] = o.conic[1] = o.conic[2] = 0.f;
    o.compensation = 0.f;

    float pc[3];
    world_to_cam_point(c, p, pc);
    if (pc[2] < near_plane || pc[2] > far_plane) return o;

    float Sc[9];
    world_to_cam_covar(c, s6, Sc);
    const PerspJ J = persp_jacobian(c, pc, width, height);
    float cov2d[4], mean2d[2];
    persp_project(c, pc, Sc, J, cov2d, mean2d);

    // add_blur (utils.cuh:458-466)
    const float det_orig = cov2d[0] * cov2d[3] - cov2d[1] * cov2d[2];
    cov2d[0] += eps2d;
    cov2d[3] += eps2d;
    const float det = cov2d[0] * cov2d[3] - cov2d[1] * cov2d[2];
    const float compensation = sqrtf(fmaxf(0.f, det_orig / det));
    if (det <= 0.f) return o;

    // inverse (utils.cuh:437-449)
    const float det2 = cov2d[0] * cov2d[3] - cov2d[1] * cov2d[2];
    const float inv_det = 1.f / det2;
    const float i00 = cov2d[3] * inv_det;
    const float i01 = -cov2d[1] * inv_det;
    const float i11 = cov2d[0] * inv_det;

    // one-sigma radius: the Beta kernel has compact support sigma < 1 (fully_fused_projection_fwd.cu:147-150)
    const float b = 0.5f * (cov2d[0] + cov2d[3]);
    const float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
    const float radius = ceilf(sqrtf(v1));
    if (radius <= radius_clip) return o;
    if (mean2d[0] + radius <= 0 || mean2d[0] - radius >= width || mean2d[1] + radius <= 0 ||
        mean2d[1] - radius >= height)
        return o;

    o.radius = (int32_t)radius;
    o.mean2d[0] = mean2d[0];
    o.mean2d[1] = mean2d[1];
    o.depth = pc[2];
    o.conic[0] = i00;
    o.conic[1] = i01;
    o.conic[2] = i11;
    o.compensation = compensation;
    return o;
}

// VJP of project_splat for one visible (camera, primitive): accumulates into v_p[3], v_s6[6] (symmetrised as
// the reference does: xx, xy+yx, xz+zx, yy, yz+zy, zz) and optionally v_R[9] (row-major), v_t[3].
// (fully_fused_projection_bwd.cu:69-201, utils.cuh:294-372,385-435,451-502)
__device__ __forceinline__ void project_splat_vjp(const Cam &c, const float p[3], const float s6[6], uint32_t width,
                                                  uint32_t height, float eps2d, const float conic[3],
                                                  const float *compensation, const float v_mean2d[2], float v_depth,
                                                  const float v_conic[3], const float *v_compensation, float v_p[3],
                                                  float v_s6[6], float *v_R, float *v_t) {
    // inverse_vjp: v_cov2d = -P vP P,  P = conic (symmetric), vP = [[vA, vB/2],[vB/2, vC]]
    const float P00 = conic[0], P01 = conic[1], P11 = conic[2];
    const float G00 = v_conic[0], G01 = 0.5f * v_conic[1], G11 = v_conic[2];
    // T = P * G
    const float T00 = P00 * G00 + P01 * G01, T01 = P00 * G01 + P01 * G11;
    const float T10 = P01 * G00 + P11 * G01, T11 = P01 * G01 + P11 * G11;
    float V00 = -(T00 * P00 + T01 * P01), V01 = -(T00 * P01 + T01 * P11);
    float V10 = -(T10 * P00 + T11 * P01), V11 = -(T10 * P01 + T11 * P11);
    if (v_compensation != nullptr) {
        const float comp = *compensation, v_comp = *v_compensation;
        const float det_conic = P00 * P11 - P01 * P01;
        const float v_sqr = v_comp * 0.5f / (comp + 1e-6f);
        const float om = 1.f - comp * comp;
        V00 += v_sqr * (om * P00 - eps2d * det_conic);
        V01 += v_sqr * (om * P01);
        V10 += v_sqr * (om * P01);
        V11 += v_sqr * (om * P11 - eps2d * det_conic);
    }

    float pc[3];
    world_to_cam_point(c, p, pc);
    float Sc[9];
    world_to_cam_covar(c, s6, Sc);
    const PerspJ J = persp_jacobian(c, pc, width, height);
    const float x = pc[0], y = pc[1];
    const float rz = J.rz, rz2 = J.rz2, rz3 = rz2 * rz;

    // v_Sc = J^T V J  (3x3)
    const float Jm[6] = {J.j00, 0.f, J.j02, 0.f, J.j11, J.j12};
    float VJ[6];  // V J (2x3)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        VJ[k] = V00 * Jm[k] + V01 * Jm[3 + k];
        VJ[3 + k] = V10 * Jm[k] + V11 * Jm[3 + k];
    }
    float vSc[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) vSc[i * 3 + k] = Jm[i] * VJ[k] + Jm[3 + i] * VJ[3 + k];

    float vpc[3];
    vpc[0] = c.fx * rz * v_mean2d[0];
    vpc[1] = c.fy * rz * v_mean2d[1];
    vpc[2] = -(c.fx * x * v_mean2d[0] + c.fy * y * v_mean2d[1]) * rz2;

    // v_J = V J Sc^T + V^T J Sc   (2x3)
    float JS[6], JSt[6];  // J Sc and J Sc^T
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        JS[k] = Jm[0] * Sc[0 * 3 + k] + Jm[2] * Sc[2 * 3 + k];
        JS[3 + k] = Jm[4] * Sc[1 * 3 + k] + Jm[5] * Sc[2 * 3 + k];
        JSt[k] = Jm[0] * Sc[k * 3 + 0] + Jm[2] * Sc[k * 3 + 2];
        JSt[3 + k] = Jm[4] * Sc[k * 3 + 1] + Jm[5] * Sc[k * 3 + 2];
    }
    float vJ[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        vJ[k] = V00 * JSt[k] + V01 * JSt[3 + k] + V00 * JS[k] + V10 * JS[3 + k];
        vJ[3 + k] = V10 * JSt[k] + V11 * JSt[3 + k] + V01 * JS[k] + V11 * JS[3 + k];
    }
    if (J.x_in) vpc[0] += -c.fx * rz2 * vJ[2];
    else vpc[2] += -c.fx * rz3 * vJ[2] * J.tx;
    if (J.y_in) vpc[1] += -c.fy * rz2 * vJ[5];
    else vpc[2] += -c.fy * rz3 * vJ[5] * J.ty;
    vpc[2] += -c.fx * rz2 * vJ[0] - c.fy * rz2 * vJ[4] + 2.f * c.fx * J.tx * rz3 * vJ[2] +
              2.f * c.fy * J.ty * rz3 * vJ[5];
    vpc[2] += v_depth;

    // world <- camera
#pragma unroll
    for (int j = 0; j < 3; ++j) v_p[j] += c.r[0 * 3 + j] * vpc[0] + c.r[1 * 3 + j] * vpc[1] + c.r[2 * 3 + j] * vpc[2];
    // v_S = R^T vSc R
    float RtV[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            RtV[i * 3 + k] = c.r[0 * 3 + i] * vSc[0 * 3 + k] + c.r[1 * 3 + i] * vSc[1 * 3 + k] + c.r[2 * 3 + i] * vSc[2 * 3 + k];
    float vS[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            vS[i * 3 + j] = RtV[i * 3 + 0] * c.r[0 * 3 + j] + RtV[i * 3 + 1] * c.r[1 * 3 + j] + RtV[i * 3 + 2] * c.r[2 * 3 + j];
    v_s6[0] += vS[0];
    v_s6[1] += vS[1] + vS[3];
    v_s6[2] += vS[2] + vS[6];
    v_s6[3] += vS[4];
    v_s6[4] += vS[5] + vS[7];
    v_s6[5] += vS[8];

    if (v_R != nullptr) {
        // v_R = vpc (x) p + vSc R S^T + vSc^T R S   with S symmetric -> (vSc + vSc^T) R S
        const float S[9] = {s6[0], s6[1], s6[2], s6[1], s6[3], s6[4], s6[2], s6[4], s6[5]};
        float RS[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                RS[i * 3 + j] = c.r[i * 3 + 0] * S[0 * 3 + j] + c.r[i * 3 + 1] * S[1 * 3 + j] + c.r[i * 3 + 2] * S[2 * 3 + j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float acc = vpc[i] * p[j];
#pragma unroll
                for (int k = 0; k < 3; ++k) acc += (vSc[i * 3 + k] + vSc[k * 3 + i]) * RS[k * 3 + j];
                v_R[i * 3 + j] += acc;
            }
#pragma unroll
        for (int i = 0; i < 3; ++i) v_t[i] += vpc[i];
    }
}

}  // namespace ubs
