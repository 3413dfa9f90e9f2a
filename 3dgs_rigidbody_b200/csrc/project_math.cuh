// Per-Gaussian math of the rigid transform + EWA projection (forward) and its VJP building blocks.
// Operation order follows the reference kernels so that, compiled with the same flags (-O3 -use_fast_math), the
// integer outputs (radii -> tile lists) agree with the reference's:
//   rigid transform     main.py:173-228, gsplat/utils.py:109-134
//   quat/scale -> covar gsplat/cuda/include/Utils.cuh:142-164, 191-205
//   world -> camera     Utils.cuh:18-57
//   pinhole / ortho EWA  Utils.cuh:428-452, 498-537; fisheye: own formulation (rs_fisheye_ray), same mathematics as 618-655
//   blur + conic + radius + culls  csrc/ProjectionEWA3DGSFused.cu:69-212, Utils.cuh:380-388
// Matrices are row-major float[9]: m[3*r + c].  Sums run k = 0,1,2 left to right (glm's order).
#pragma once
#include "common.cuh"

struct RsBody { // one row of the pose table held in shared memory (20 floats)
    float R[9];
    float c[3];
    float t[3];
    float q[4];
    float pad;
};
#define RS_BODY_FLOATS 20
#define RS_MAX_SMEM_BODIES 512

// q/|q| then gsplat/utils.py:109-134.  Every operation is individually rounded (no FMA contraction, IEEE sqrt/div)
// because the reference evaluates this in torch, one rounded op per tensor op.
__device__ __forceinline__ void rs_make_body(const rs_rigid_t &rg, int k, float *out /*20 floats*/) {
    float w = rg.body_quats[4 * k + 0], x = rg.body_quats[4 * k + 1], y = rg.body_quats[4 * k + 2],
          z = rg.body_quats[4 * k + 3];
    // main.py:207 torch.linalg.norm of a 4-vector on CUDA: four threads square one component each and combine by a
    // shuffle tree, i.e. sqrt((w^2 + y^2) + (x^2 + z^2)) with rounded squares -- measured bit-exact on 512 random
    // quaternions by tools/rigid_association_experiment.py (profiles/r02_rigid_association.json); the sequential order
    // ((w^2 + x^2) + y^2) + z^2 matches only 85 % of them
    float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(w, w), __fmul_rn(y, y)), __fadd_rn(__fmul_rn(x, x), __fmul_rn(z, z))));
    w = __fdiv_rn(w, n);
    x = __fdiv_rn(x, n);
    y = __fdiv_rn(y, n);
    z = __fdiv_rn(z, n);
    float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    float xy = __fmul_rn(x, y), xz = __fmul_rn(x, z), yz = __fmul_rn(y, z);
    float wx = __fmul_rn(w, x), wy = __fmul_rn(w, y), wz = __fmul_rn(w, z);
    out[0] = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(yy, zz)));
    out[1] = __fmul_rn(2.f, __fsub_rn(xy, wz));
    out[2] = __fmul_rn(2.f, __fadd_rn(xz, wy));
    out[3] = __fmul_rn(2.f, __fadd_rn(xy, wz));
    out[4] = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(xx, zz)));
    out[5] = __fmul_rn(2.f, __fsub_rn(yz, wx));
    out[6] = __fmul_rn(2.f, __fsub_rn(xz, wy));
    out[7] = __fmul_rn(2.f, __fadd_rn(yz, wx));
    out[8] = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(xx, yy)));
    if (rg.body_centers != nullptr) {
        out[9] = rg.body_centers[3 * k + 0];
        out[10] = rg.body_centers[3 * k + 1];
        out[11] = rg.body_centers[3 * k + 2];
    } else {
        out[9] = out[10] = out[11] = 0.f;
    }
    out[12] = rg.body_trans[3 * k + 0];
    out[13] = rg.body_trans[3 * k + 1];
    out[14] = rg.body_trans[3 * k + 2];
    out[15] = w;
    out[16] = x;
    out[17] = y;
    out[18] = z;
    out[19] = 0.f;
}

// Fill the shared-memory pose table (whole CTA).  Caller syncs afterwards.
__device__ __forceinline__ void rs_load_pose_table(const rs_rigid_t &rg, float *table) {
    if (rg.cluster_ids == nullptr || rg.K > RS_MAX_SMEM_BODIES)
        return;
    for (int k = threadIdx.x; k < rg.K; k += blockDim.x)
        rs_make_body(rg, k, table + RS_BODY_FLOATS * k);
}

// mean' = R (mean - c) + c + t ; quat' = q_body (x) quat  (main.py:210-222, 173-181).
// The quaternion product is evaluated one rounded op at a time, left to right, like the torch expression.
__device__ __forceinline__ void rs_apply_body(const float *b, float m[3], float q[4], bool has_quat) {
    float dx = __fsub_rn(m[0], b[9]), dy = __fsub_rn(m[1], b[10]), dz = __fsub_rn(m[2], b[11]);
    float rx = __fmaf_rn(dz, b[2], __fmaf_rn(dy, b[1], __fmul_rn(dx, b[0])));
    float ry = __fmaf_rn(dz, b[5], __fmaf_rn(dy, b[4], __fmul_rn(dx, b[3])));
    float rz = __fmaf_rn(dz, b[8], __fmaf_rn(dy, b[7], __fmul_rn(dx, b[6])));
    m[0] = __fadd_rn(__fadd_rn(rx, b[9]), b[12]);
    m[1] = __fadd_rn(__fadd_rn(ry, b[10]), b[13]);
    m[2] = __fadd_rn(__fadd_rn(rz, b[11]), b[14]);
    if (has_quat) {
        float w1 = b[15], x1 = b[16], y1 = b[17], z1 = b[18];
        float w2 = q[0], x2 = q[1], y2 = q[2], z2 = q[3];
        q[0] = __fsub_rn(__fsub_rn(__fsub_rn(__fmul_rn(w1, w2), __fmul_rn(x1, x2)), __fmul_rn(y1, y2)),
                         __fmul_rn(z1, z2));
        q[1] = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, x2), __fmul_rn(x1, w2)), __fmul_rn(y1, z2)),
                         __fmul_rn(z1, y2));
        q[2] = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(w1, y2), __fmul_rn(x1, z2)), __fmul_rn(y1, w2)),
                         __fmul_rn(z1, x2));
        q[3] = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(w1, z2), __fmul_rn(x1, y2)), __fmul_rn(y1, x2)),
                         __fmul_rn(z1, w2));
    }
}

// Fetch the body row for Gaussian `gid` (shared table when it fits, else computed on the fly) and apply it.
// `rot_out` (optional, 9 floats) receives R_k for covariance inputs / the backward pass.
__device__ __forceinline__ int rs_rigid_transform(const rs_rigid_t &rg, const float *table, int gid, float m[3],
                                                  float q[4], bool has_quat, float *body_out /*20 or null*/) {
    if (rg.cluster_ids == nullptr)
        return -1;
    int k = rg.cluster_ids[gid];
    if (k < 0 || k >= rg.K)
        return -1;
    float local[RS_BODY_FLOATS];
    const float *b;
    if (rg.K <= RS_MAX_SMEM_BODIES) {
        b = table + RS_BODY_FLOATS * k;
    } else {
        rs_make_body(rg, k, local);
        b = local;
    }
    rs_apply_body(b, m, q, has_quat);
    if (body_out != nullptr) {
#pragma unroll
        for (int i = 0; i < RS_BODY_FLOATS; ++i)
            body_out[i] = b[i];
    }
    return k;
}

// Utils.cuh:142-164 (rsqrt-normalised, wxyz) -> row-major R
__device__ __forceinline__ void rs_quat_to_rotmat(const float q[4], float R[9]) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    float inv_norm = rsqrtf(x * x + y * y + z * z + w * w);
    x *= inv_norm;
    y *= inv_norm;
    z *= inv_norm;
    w *= inv_norm;
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    float wx = w * x, wy = w * y, wz = w * z;
    R[0] = (1.f - 2.f * (y2 + z2));
    R[3] = (2.f * (xy + wz));
    R[6] = (2.f * (xz - wy));
    R[1] = (2.f * (xy - wz));
    R[4] = (1.f - 2.f * (x2 + z2));
    R[7] = (2.f * (yz + wx));
    R[2] = (2.f * (xz + wy));
    R[5] = (2.f * (yz - wx));
    R[8] = (1.f - 2.f * (x2 + y2));
}

// C = A * B, row-major, k ascending
__device__ __forceinline__ void rs_mm3(const float A[9], const float B[9], float C[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i + 0] * B[0 + j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// C = A * B^T
__device__ __forceinline__ void rs_mm3_nt(const float A[9], const float B[9], float C[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i + 0] * B[3 * j + 0] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
// C = A^T * B
__device__ __forceinline__ void rs_mm3_tn(const float A[9], const float B[9], float C[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[0 + i] * B[0 + j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}

// Utils.cuh:191-205: covar = (R S)(R S)^T
__device__ __forceinline__ void rs_quat_scale_to_covar(const float q[4], const float s[3], float covar[9],
                                                       float *Rq_out /*9 or null*/) {
    float Rq[9];
    rs_quat_to_rotmat(q, Rq);
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        M[3 * i + 0] = Rq[3 * i + 0] * s[0];
        M[3 * i + 1] = Rq[3 * i + 1] * s[1];
        M[3 * i + 2] = Rq[3 * i + 2] * s[2];
    }
    rs_mm3_nt(M, M, covar);
    if (Rq_out != nullptr) {
#pragma unroll
        for (int i = 0; i < 9; ++i)
            Rq_out[i] = Rq[i];
    }
}

struct RsCam {
    float R[9]; // world->camera rotation, row-major
    float t[3];
    float fx, fy, cx, cy;
};
__device__ __forceinline__ void rs_load_cam(const float *__restrict__ viewmat, const float *__restrict__ K, RsCam &c) {
    c.R[0] = viewmat[0];
    c.R[1] = viewmat[1];
    c.R[2] = viewmat[2];
    c.R[3] = viewmat[4];
    c.R[4] = viewmat[5];
    c.R[5] = viewmat[6];
    c.R[6] = viewmat[8];
    c.R[7] = viewmat[9];
    c.R[8] = viewmat[10];
    c.t[0] = viewmat[3];
    c.t[1] = viewmat[7];
    c.t[2] = viewmat[11];
    c.fx = K[0];
    c.fy = K[4];
    c.cx = K[2];
    c.cy = K[5];
}

// The 2x3 projection Jacobian J = [[j00, j01, j02], [j10, j11, j12]] and the 2D mean.
struct RsProj {
    float J[6];
    float mx, my;
};

// Utils.cuh:498-537
__device__ __forceinline__ void rs_persp_J(const float p[3], const RsCam &c, uint32_t width, uint32_t height,
                                           RsProj &o) {
    float x = p[0], y = p[1], z = p[2];
    float tan_fovx = 0.5f * width / c.fx;
    float tan_fovy = 0.5f * height / c.fy;
    float lim_x_pos = (width - c.cx) / c.fx + 0.3f * tan_fovx;
    float lim_x_neg = c.cx / c.fx + 0.3f * tan_fovx;
    float lim_y_pos = (height - c.cy) / c.fy + 0.3f * tan_fovy;
    float lim_y_neg = c.cy / c.fy + 0.3f * tan_fovy;
    float rz = 1.f / z;
    float rz2 = rz * rz;
    float tx = z * min(lim_x_pos, max(-lim_x_neg, x * rz));
    float ty = z * min(lim_y_pos, max(-lim_y_neg, y * rz));
    o.J[0] = c.fx * rz;
    o.J[1] = 0.f;
    o.J[2] = -c.fx * tx * rz2;
    o.J[3] = 0.f;
    o.J[4] = c.fy * rz;
    o.J[5] = -c.fy * ty * rz2;
    o.mx = c.fx * x * rz + c.cx;
    o.my = c.fy * y * rz + c.cy;
}
// Utils.cuh:428-452
__device__ __forceinline__ void rs_ortho_J(const float p[3], const RsCam &c, RsProj &o) {
    o.J[0] = c.fx;
    o.J[1] = 0.f;
    o.J[2] = 0.f;
    o.J[3] = 0.f;
    o.J[4] = c.fy;
    o.J[5] = 0.f;
    o.mx = c.fx * p[0] + c.cx;
    o.my = c.fy * p[1] + c.cy;
}
// Equidistant fisheye (replaces Utils.cuh:618-655): (u, v) = (fx s x + cx, fy s y + cy) with r = |(x, y)|, theta = atan2(r, z)
// and s = theta / r.  Everything the Jacobian and its derivatives need is expressed through three scalars of the ray:
//     s = theta / r,    w = 1 / (r^2 + z^2),    k = (z w - s) / r^2        (ds/dx = x k, ds/dy = y k, ds/dz = -w)
// so that J = [[fx (s + x^2 k), fx x y k, -fx x w], [fy x y k, fy (s + y^2 k), -fy y w]].  Near the optical axis z w - s is a
// difference of two numbers that agree to O(r^2): there the Taylor expansions in (r/z)^2 are used instead (the reference
// keeps the cancellation and adds 1e-7 to r and x^2).
struct RsFisheyeRay {
    float s, w, k, r2;
};
__device__ __forceinline__ RsFisheyeRay rs_fisheye_ray(float x, float y, float z) {
    RsFisheyeRay f;
    f.r2 = x * x + y * y;
    const float rho2 = f.r2 + z * z;
    f.w = 1.f / rho2;
    if (z > 0.f && f.r2 < 1e-4f * z * z) { // |r/z| < 0.01: series, relative error below 1e-9
        const float iz = 1.f / z, q = f.r2 * iz * iz;
        f.s = iz * (1.f - q * (1.f / 3.f - q * 0.2f));
        f.k = -iz * iz * iz * (2.f / 3.f - q * 0.8f);
    } else {
        const float r = sqrtf(f.r2);
        f.s = atan2f(r, z) / r;
        f.k = (z * f.w - f.s) / f.r2;
    }
    return f;
}
__device__ __forceinline__ void rs_fisheye_J(const float p[3], const RsCam &c, RsProj &o) {
    const float x = p[0], y = p[1];
    const RsFisheyeRay f = rs_fisheye_ray(x, y, p[2]);
    o.mx = c.fx * f.s * x + c.cx;
    o.my = c.fy * f.s * y + c.cy;
    const float xyk = x * y * f.k;
    o.J[0] = c.fx * (f.s + x * x * f.k);
    o.J[1] = c.fx * xyk;
    o.J[2] = -c.fx * x * f.w;
    o.J[3] = c.fy * xyk;
    o.J[4] = c.fy * (f.s + y * y * f.k);
    o.J[5] = -c.fy * y * f.w;
}

// cov2d = J * cov3d * J^T, all four entries computed separately like glm (mat3x2 * mat3 * mat2x3).
// out = {c00, c01, c10, c11} (row, col).
__device__ __forceinline__ void rs_project_cov(const float J[6], const float S[9], float out[4]) {
    float T[6];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            T[3 * i + j] = J[3 * i + 0] * S[0 + j] + J[3 * i + 1] * S[3 + j] + J[3 * i + 2] * S[6 + j];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            out[2 * i + j] = T[3 * i + 0] * J[3 * j + 0] + T[3 * i + 1] * J[3 * j + 1] + T[3 * i + 2] * J[3 * j + 2];
}

struct RsProjected {
    int32_t rx, ry; // 0,0 when culled
    float mx, my, depth;
    float ca, cb, cc; // conic
    float comp;
};

// csrc/ProjectionEWA3DGSFused.cu:69-212.  `mean` / `covar` are in world space AFTER the rigid transform.
__device__ __forceinline__ bool rs_project_gaussian(const float mean[3], const float covar[9], const RsCam &cam,
                                                    int camera_model, uint32_t width, uint32_t height, float eps2d,
                                                    float near_plane, float far_plane, float radius_clip,
                                                    const float *opacity /*null = none*/, bool use_comp,
                                                    RsProjected &o) {
    o.rx = 0;
    o.ry = 0;
    float p[3];
    p[0] = cam.R[0] * mean[0] + cam.R[1] * mean[1] + cam.R[2] * mean[2] + cam.t[0];
    p[1] = cam.R[3] * mean[0] + cam.R[4] * mean[1] + cam.R[5] * mean[2] + cam.t[1];
    p[2] = cam.R[6] * mean[0] + cam.R[7] * mean[1] + cam.R[8] * mean[2] + cam.t[2];
    if (p[2] < near_plane || p[2] > far_plane)
        return false;

    float A[9], covar_c[9];
    rs_mm3(cam.R, covar, A);
    rs_mm3_nt(A, cam.R, covar_c);

    RsProj pj;
    if (camera_model == RS_PINHOLE)
        rs_persp_J(p, cam, width, height, pj);
    else if (camera_model == RS_ORTHO)
        rs_ortho_J(p, cam, pj);
    else
        rs_fisheye_J(p, cam, pj);
    float c2[4];
    rs_project_cov(pj.J, covar_c, c2);

    // add_blur, Utils.cuh:380-388
    float det_orig = c2[0] * c2[3] - c2[2] * c2[1];
    c2[0] += eps2d;
    c2[3] += eps2d;
    float det_blur = c2[0] * c2[3] - c2[2] * c2[1];
    float compensation = sqrtf(max(0.f, det_orig / det_blur));
    if (det_blur <= 0.f)
        return false;

    // glm::inverse(mat2)
    float ood = 1.f / (c2[0] * c2[3] - c2[1] * c2[2]);
    float inv00 = c2[3] * ood;  // conic.x
    float inv01 = -c2[2] * ood; // conic.y  (glm covar2d_inv[0][1])
    float inv11 = c2[0] * ood;  // conic.z

    float extend = 3.33f;
    if (opacity != nullptr) {
        float op = *opacity;
        if (use_comp)
            op *= compensation;
        if (op < RS_ALPHA_THRESHOLD)
            return false;
        extend = min(extend, sqrtf(2.0f * __logf(op / RS_ALPHA_THRESHOLD)));
    }
    float radius_x = ceilf(extend * sqrtf(c2[0]));
    float radius_y = ceilf(extend * sqrtf(c2[3]));
    if (radius_x <= radius_clip && radius_y <= radius_clip)
        return false;
    if (pj.mx + radius_x <= 0 || pj.mx - radius_x >= width || pj.my + radius_y <= 0 || pj.my - radius_y >= height)
        return false;

    o.rx = (int32_t)radius_x;
    o.ry = (int32_t)radius_y;
    o.mx = pj.mx;
    o.my = pj.my;
    o.depth = p[2];
    o.ca = inv00;
    o.cb = inv01;
    o.cc = inv11;
    o.comp = compensation;
    return true;
}
