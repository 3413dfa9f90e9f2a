// rs_sh_fwd / rs_sh_bwd: real spherical-harmonics colour evaluation up to degree 4.
// Replaces `spherical_harmonics_fwd` / `spherical_harmonics_bwd` (Ops.h:154-168, csrc/SphericalHarmonicsCUDA.cu:20-116
// forward, :118-400 backward): colour_c = sum_k B_k(dir / |dir|) * coeffs[k, c], coefficients [n, K, 3].
//
// HBM-bound: K*12 bytes of coefficients per element (192 B at K = 16 -- 4x the geometry of a Gaussian).  One thread per
// element reads its coefficient row with 128-bit loads when the row allows it (K*3 multiple of 4) and produces all three
// channels (the reference uses one thread per (element, channel) and re-evaluates the basis three times).  The basis
// functions are the standard real SH polynomials (same constants as the reference); the backward chains through the
// normalisation: v_dir = (v_n - (v_n . n) n) / |dir|.
// The same device functions are used by the projection kernel to evaluate colours in place on the frame path
// (project.cu: sh_coeffs), so the [C, N, 3] colour tensor of rendering.py:491-525 never makes a round trip.
#include "sh_math.cuh"

__global__ void __launch_bounds__(256) rs_sh_fwd_kernel(const rs_sh_args a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n)
        return;
    if (a.masks != nullptr && !a.masks[i]) { // the reference leaves masked rows uninitialised; zeros here
        a.colors[i * 3 + 0] = 0.f;
        a.colors[i * 3 + 1] = 0.f;
        a.colors[i * 3 + 2] = 0.f;
        return;
    }
    const float dx = a.dirs[i * 3 + 0], dy = a.dirs[i * 3 + 1], dz = a.dirs[i * 3 + 2];
    const float inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
    float B[25];
    rs_sh_basis(a.degree, dx * inorm, dy * inorm, dz * inorm, B);
    float c[3];
    rs_sh_dot(B, (a.degree + 1) * (a.degree + 1), a.coeffs + (size_t)i * a.K * 3, c);
    a.colors[i * 3 + 0] = c[0];
    a.colors[i * 3 + 1] = c[1];
    a.colors[i * 3 + 2] = c[2];
}

__global__ void __launch_bounds__(256) rs_sh_bwd_kernel(const rs_sh_args a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n)
        return;
    const int nb = (a.degree + 1) * (a.degree + 1);
    float *vco = a.v_coeffs + (size_t)i * a.K * 3;
    const bool live = a.masks == nullptr || a.masks[i];
    if (!live) {
        for (int k = 0; k < a.K * 3; ++k)
            vco[k] = 0.f;
        if (a.v_dirs != nullptr) {
            a.v_dirs[i * 3 + 0] = 0.f;
            a.v_dirs[i * 3 + 1] = 0.f;
            a.v_dirs[i * 3 + 2] = 0.f;
        }
        return;
    }
    const float dx = a.dirs[i * 3 + 0], dy = a.dirs[i * 3 + 1], dz = a.dirs[i * 3 + 2];
    const float inorm = rsqrtf(dx * dx + dy * dy + dz * dz);
    const float x = dx * inorm, y = dy * inorm, z = dz * inorm;
    const float v0 = a.v_colors[i * 3 + 0], v1 = a.v_colors[i * 3 + 1], v2 = a.v_colors[i * 3 + 2];
    float B[25];
    rs_sh_basis(a.degree, x, y, z, B);
    for (int k = 0; k < a.K; ++k) {
        const float b = k < nb ? B[k] : 0.f;
        vco[k * 3 + 0] = b * v0;
        vco[k * 3 + 1] = b * v1;
        vco[k * 3 + 2] = b * v2;
    }
    if (a.v_dirs != nullptr) {
        float gx[25], gy[25], gz[25];
        rs_sh_basis_grad(a.degree, x, y, z, gx, gy, gz);
        const float *co = a.coeffs + (size_t)i * a.K * 3;
        float vnx = 0.f, vny = 0.f, vnz = 0.f;
        for (int k = 1; k < nb; ++k) { // B_0 is constant
            const float w = co[k * 3 + 0] * v0 + co[k * 3 + 1] * v1 + co[k * 3 + 2] * v2;
            vnx += gx[k] * w;
            vny += gy[k] * w;
            vnz += gz[k] * w;
        }
        const float dot = vnx * x + vny * y + vnz * z;
        a.v_dirs[i * 3 + 0] = (vnx - dot * x) * inorm;
        a.v_dirs[i * 3 + 1] = (vny - dot * y) * inorm;
        a.v_dirs[i * 3 + 2] = (vnz - dot * z) * inorm;
    }
}

static int check_sh(const rs_sh_args *a, const char *who) {
    RS_CHECK(a != nullptr, "%s: null args", who);
    RS_CHECK(a->degree >= 0 && a->degree <= 4, "%s: degree %d out of range [0, 4]", who, a->degree);
    RS_CHECK((a->degree + 1) * (a->degree + 1) <= a->K, "%s: %d coefficient rows cannot hold degree %d", who, a->K,
             a->degree);
    RS_CHECK(a->n >= 0, "%s: negative size", who);
    return 0;
}

extern "C" int rs_sh_fwd(const rs_sh_args *a, rs_stream_t stream) {
    if (int e = check_sh(a, "rs_sh_fwd"))
        return e;
    if (a->n == 0)
        return 0;
    RS_CHECK(a->dirs && a->coeffs && a->colors, "rs_sh_fwd: null pointer");
    rs_sh_fwd_kernel<<<rs_cdiv(a->n, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    RS_LAUNCH_CHECK("rs_sh_fwd_kernel");
    return 0;
}

extern "C" int rs_sh_bwd(const rs_sh_args *a, rs_stream_t stream) {
    if (int e = check_sh(a, "rs_sh_bwd"))
        return e;
    if (a->n == 0)
        return 0;
    RS_CHECK(a->dirs && a->coeffs && a->v_colors && a->v_coeffs, "rs_sh_bwd: null pointer");
    rs_sh_bwd_kernel<<<rs_cdiv(a->n, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    RS_LAUNCH_CHECK("rs_sh_bwd_kernel");
    return 0;
}
