// Real spherical-harmonics basis (degree <= 4) and its gradient, shared by sh.cu and project.cu.
// Constants and polynomial forms: the standard real SH used by the reference (csrc/SphericalHarmonicsCUDA.cu:20-116).
#pragma once
#include "common.cuh"

#define RS_SH_C0 0.28209479177387814f
#define RS_SH_C1 0.4886025119029199f

// B[0 .. (degree+1)^2) at the UNIT direction (x, y, z)
__device__ __forceinline__ void rs_sh_basis(int degree, float x, float y, float z, float B[25]) {
    B[0] = RS_SH_C0;
    if (degree < 1)
        return;
    B[1] = -RS_SH_C1 * y;
    B[2] = RS_SH_C1 * z;
    B[3] = -RS_SH_C1 * x;
    if (degree < 2)
        return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    B[4] = 1.0925484305920792f * xy;
    B[5] = -1.0925484305920792f * yz;
    B[6] = 0.31539156525252005f * (2.f * zz - xx - yy);
    B[7] = -1.0925484305920792f * xz;
    B[8] = 0.5462742152960396f * (xx - yy);
    if (degree < 3)
        return;
    B[9] = -0.5900435899266435f * y * (3.f * xx - yy);
    B[10] = 2.890611442640554f * xy * z;
    B[11] = -0.4570457994644658f * y * (4.f * zz - xx - yy);
    B[12] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
    B[13] = -0.4570457994644658f * x * (4.f * zz - xx - yy);
    B[14] = 1.445305721320277f * z * (xx - yy);
    B[15] = -0.5900435899266435f * x * (xx - 3.f * yy);
    if (degree < 4)
        return;
    B[16] = 2.5033429417967046f * xy * (xx - yy);
    B[17] = -1.7701307697799304f * yz * (3.f * xx - yy);
    B[18] = 0.9461746957575601f * xy * (7.f * zz - 1.f);
    B[19] = -0.6690465435572892f * yz * (7.f * zz - 3.f);
    B[20] = 0.10578554691520431f * (zz * (35.f * zz - 30.f) + 3.f);
    B[21] = -0.6690465435572892f * xz * (7.f * zz - 3.f);
    B[22] = 0.47308734787878004f * (xx - yy) * (7.f * zz - 1.f);
    B[23] = -1.7701307697799304f * xz * (xx - 3.f * yy);
    B[24] = 0.6258357354491761f * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
}

// partial derivatives of the polynomials above with x, y, z taken as independent variables (the radial component is
// removed afterwards by the normalisation VJP, so any polynomial form that agrees on the sphere gives the same result)
__device__ __forceinline__ void rs_sh_basis_grad(int degree, float x, float y, float z, float gx[25], float gy[25],
                                                 float gz[25]) {
    gx[0] = gy[0] = gz[0] = 0.f;
    if (degree < 1)
        return;
    gx[1] = 0.f, gy[1] = -RS_SH_C1, gz[1] = 0.f;
    gx[2] = 0.f, gy[2] = 0.f, gz[2] = RS_SH_C1;
    gx[3] = -RS_SH_C1, gy[3] = 0.f, gz[3] = 0.f;
    if (degree < 2)
        return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    {
        const float a0 = 1.0925484305920792f, a1 = -1.0925484305920792f, a2 = 0.31539156525252005f,
                    a3 = -1.0925484305920792f, a4 = 0.5462742152960396f;
        gx[4] = a0 * y, gy[4] = a0 * x, gz[4] = 0.f;
        gx[5] = 0.f, gy[5] = a1 * z, gz[5] = a1 * y;
        gx[6] = -2.f * a2 * x, gy[6] = -2.f * a2 * y, gz[6] = 4.f * a2 * z;
        gx[7] = a3 * z, gy[7] = 0.f, gz[7] = a3 * x;
        gx[8] = 2.f * a4 * x, gy[8] = -2.f * a4 * y, gz[8] = 0.f;
    }
    if (degree < 3)
        return;
    {
        const float b0 = -0.5900435899266435f, b1 = 2.890611442640554f, b2 = -0.4570457994644658f,
                    b3 = 0.3731763325901154f, b4 = -0.4570457994644658f, b5 = 1.445305721320277f,
                    b6 = -0.5900435899266435f;
        gx[9] = 6.f * b0 * xy, gy[9] = b0 * (3.f * xx - 3.f * yy), gz[9] = 0.f;
        gx[10] = b1 * yz, gy[10] = b1 * xz, gz[10] = b1 * xy;
        gx[11] = -2.f * b2 * xy, gy[11] = b2 * (4.f * zz - xx - 3.f * yy), gz[11] = 8.f * b2 * yz;
        gx[12] = -6.f * b3 * xz, gy[12] = -6.f * b3 * yz, gz[12] = b3 * (6.f * zz - 3.f * xx - 3.f * yy);
        gx[13] = b4 * (4.f * zz - 3.f * xx - yy), gy[13] = -2.f * b4 * xy, gz[13] = 8.f * b4 * xz;
        gx[14] = 2.f * b5 * xz, gy[14] = -2.f * b5 * yz, gz[14] = b5 * (xx - yy);
        gx[15] = b6 * (3.f * xx - 3.f * yy), gy[15] = -6.f * b6 * xy, gz[15] = 0.f;
    }
    if (degree < 4)
        return;
    {
        const float c0 = 2.5033429417967046f, c1 = -1.7701307697799304f, c2 = 0.9461746957575601f,
                    c3 = -0.6690465435572892f, c4 = 0.10578554691520431f, c5 = -0.6690465435572892f,
                    c6 = 0.47308734787878004f, c7 = -1.7701307697799304f, c8 = 0.6258357354491761f;
        gx[16] = c0 * y * (3.f * xx - yy), gy[16] = c0 * x * (xx - 3.f * yy), gz[16] = 0.f;
        gx[17] = 6.f * c1 * xy * z, gy[17] = c1 * z * (3.f * xx - 3.f * yy), gz[17] = c1 * y * (3.f * xx - yy);
        gx[18] = c2 * y * (7.f * zz - 1.f), gy[18] = c2 * x * (7.f * zz - 1.f), gz[18] = 14.f * c2 * xy * z;
        gx[19] = 0.f, gy[19] = c3 * z * (7.f * zz - 3.f), gz[19] = c3 * y * (21.f * zz - 3.f);
        gx[20] = 0.f, gy[20] = 0.f, gz[20] = c4 * (140.f * zz * z - 60.f * z);
        gx[21] = c5 * z * (7.f * zz - 3.f), gy[21] = 0.f, gz[21] = c5 * x * (21.f * zz - 3.f);
        gx[22] = 2.f * c6 * x * (7.f * zz - 1.f), gy[22] = -2.f * c6 * y * (7.f * zz - 1.f),
        gz[22] = 14.f * c6 * z * (xx - yy);
        gx[23] = c7 * z * (3.f * xx - 3.f * yy), gy[23] = -6.f * c7 * xy * z, gz[23] = c7 * x * (xx - 3.f * yy);
        gx[24] = c8 * (4.f * xx * x - 12.f * x * yy), gy[24] = c8 * (4.f * yy * y - 12.f * xx * y), gz[24] = 0.f;
    }
}

// c[0..3) = sum_{k < nb} B[k] * row[k*3 + c]; 128-bit loads when the row is 16-byte aligned
__device__ __forceinline__ void rs_sh_dot(const float B[25], int nb, const float *__restrict__ row, float c[3]) {
    c[0] = c[1] = c[2] = 0.f;
    if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        // flat index f = k*3 + ch; walk it four floats at a time
        const int nf = nb * 3;
        int f = 0;
        for (; f + 4 <= nf; f += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(row + f);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ff = f + j;
                const int k = ff / 3, ch = ff - 3 * k;
                const float t = B[k] * vv[j];
                if (ch == 0)
                    c[0] += t;
                else if (ch == 1)
                    c[1] += t;
                else
                    c[2] += t;
            }
        }
        for (; f < nf; ++f) {
            const int k = f / 3, ch = f - 3 * k;
            c[ch] += B[k] * row[f];
        }
    } else {
        for (int k = 0; k < nb; ++k) {
            c[0] += B[k] * row[k * 3 + 0];
            c[1] += B[k] * row[k * 3 + 1];
            c[2] += B[k] * row[k * 3 + 2];
        }
    }
}
