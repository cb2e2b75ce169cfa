// thb_math.cuh - scalar building blocks shared by the kernels (host+device inline).
//
// Conventions restated from the reference (citations relative to the THUNDER tree):
//   * rotate3D(quat)              src/Geometry/Euler.cpp:181-189    R = I + 2 q0 K + 2 K^2
//   * Hermitian fold              include/Image/Volume.h:135-147    x < 0 -> negate (x,y,z), conjugate
//   * trilinear weights           include/Functions/Interpolation.h:163-200  w[k][j][i] = v0[i]*v1[j]*v2[k]
//   * half-complex index + wrap   include/Image/Volume.h:567-575    negative y/z stored at +n
//   * translation phase ramp      src/Image/ImageFunctions.cpp:233-252
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define THB_HD __host__ __device__ __forceinline__
#else
#define THB_HD inline
#endif

namespace thb {

// first two columns of the rotation matrix (the third multiplies the zero z-coordinate of a slice)
struct Rot2 {
    double c0[3];  // R(:,0)
    double c1[3];  // R(:,1)
};

// Full 3x3, column-major m[c*3+r], exactly the expression of Euler.cpp:181-189
THB_HD void quat_to_mat(const double q[4], double m[9])
{
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    // K = [0 -z y; z 0 -x; -y x 0];  K^2 = [-(y^2+z^2) xy xz; xy -(x^2+z^2) yz; xz yz -(x^2+y^2)]
    m[0] = 1.0 + 2.0 * (-(y * y + z * z));
    m[1] = 2.0 * w * z + 2.0 * (x * y);
    m[2] = 2.0 * w * (-y) + 2.0 * (x * z);
    m[3] = 2.0 * w * (-z) + 2.0 * (x * y);
    m[4] = 1.0 + 2.0 * (-(x * x + z * z));
    m[5] = 2.0 * w * x + 2.0 * (y * z);
    m[6] = 2.0 * w * y + 2.0 * (x * z);
    m[7] = 2.0 * w * (-x) + 2.0 * (y * z);
    m[8] = 1.0 + 2.0 * (-(x * x + y * y));
}

THB_HD Rot2 quat_to_rot2(const double q[4])
{
    double m[9];
    quat_to_mat(q, m);
    Rot2 r;
    r.c0[0] = m[0]; r.c0[1] = m[1]; r.c0[2] = m[2];
    r.c1[0] = m[3]; r.c1[1] = m[4]; r.c1[2] = m[5];
    return r;
}

// MODE_2D (src/Geometry/Euler.cpp:125-131, rotate2D(dmat22&, dvec2)): the in-plane rotation [[c, -s], [s, c]] of the unit
// vector (c, s) = (quat[0], quat[1]), written as the first two columns of a 3x3 whose z row is zero: the slice coordinate is
// then (c a - s b, s a + c b, 0) in the same double arithmetic as the reference's dmat22 * dvec2, and every 3D kernel works on
// a reference image stored as plane 0 of a two-plane volume (zd = 0: the bilinear weights times exactly 1 and 0)
THB_HD Rot2 make_rot2(const double q[4], int mode2D)
{
    if (!mode2D) return quat_to_rot2(q);
    Rot2 r;
    r.c0[0] = q[0]; r.c0[1] = q[1]; r.c0[2] = 0.0;
    r.c1[0] = -q[1]; r.c1[1] = q[0]; r.c1[2] = 0.0;
    return r;
}

// One trilinear cell in a half-complex volume of dimension n (nColFT = n/2+1):
// element offsets of the 8 corners (order [k][j][i], i fastest, as the reference's w[2][2][2])
// and their weights.  conj = the value (gather) / the inserted value (scatter) must be conjugated.
struct Cell {
    float w[8];
    int64_t idx[8];
    bool conj;
};

// the 3 coordinates -> fold, floor, fractional parts; returns conj flag.
THB_HD bool fold_floor(float& x, float& y, float& z, int& x0, int& y0, int& z0, float& xd, float& yd, float& zd)
{
    bool conj = false;
    if (!(x >= 0.0f)) {  // reference: if (iCol >= 0) return false;
        x = -x; y = -y; z = -z;
        conj = true;
    }
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    xd = x - fx; yd = y - fy; zd = z - fz;
    x0 = (int)fx; y0 = (int)fy; z0 = (int)fz;
    return conj;
}

#if defined(__CUDACC__)
// Same results as fold_floor, without the conversion pipe: for |v| < 2^22, (v + 1.5*2^23) rounded towards
// -inf is 1.5*2^23 + floor(v) exactly, its bit pattern is 0x4B400000 + floor(v), and v - floor(v) is exact.
// The integer outputs are BIASED by THB_FLOOR_BIAS (callers fold the bias into their origin).
#define THB_FLOOR_BIAS 0x4B400000
__device__ __forceinline__ bool fold_floor_fast(float& x, float& y, float& z, int& xb, int& yb, int& zb, float& xd, float& yd,
                                                float& zd)
{
    bool conj = false;
    if (!(x >= 0.0f)) {
        x = -x; y = -y; z = -z;
        conj = true;
    }
    const float M = 12582912.0f;
    const float tx = __fadd_rd(x, M), ty = __fadd_rd(y, M), tz = __fadd_rd(z, M);
    xd = x - (tx - M); yd = y - (ty - M); zd = z - (tz - M);
    xb = __float_as_int(tx); yb = __float_as_int(ty); zb = __float_as_int(tz);
    return conj;
}
#endif

THB_HD void tri_weights(float xd, float yd, float zd, float w[8])
{
    const float vx0 = 1.0f - xd, vx1 = xd;
    const float vy0 = 1.0f - yd, vy1 = yd;
    const float vz0 = 1.0f - zd, vz1 = zd;
    const float a00 = vx0 * vy0, a01 = vx1 * vy0, a10 = vx0 * vy1, a11 = vx1 * vy1;
    w[0] = a00 * vz0; w[1] = a01 * vz0; w[2] = a10 * vz0; w[3] = a11 * vz0;
    w[4] = a00 * vz1; w[5] = a01 * vz1; w[6] = a10 * vz1; w[7] = a11 * vz1;
}

THB_HD int wrap_idx(int j, int n) { return j >= 0 ? j : j + n; }

// row offsets (in elements) of the four (y,z) corner rows; x0 and x0+1 are added by the caller
THB_HD void row_offsets(int y0, int z0, int n, int nColFT, int64_t off[4])
{
    const int ya = wrap_idx(y0, n), yb = wrap_idx(y0 + 1, n);
    const int za = wrap_idx(z0, n), zb = wrap_idx(z0 + 1, n);
    const int64_t plane = (int64_t)nColFT * n;
    off[0] = za * plane + (int64_t)ya * nColFT;
    off[1] = za * plane + (int64_t)yb * nColFT;
    off[2] = zb * plane + (int64_t)ya * nColFT;
    off[3] = zb * plane + (int64_t)yb * nColFT;
}

// slice coordinate of packed pixel (a = pf*iCol, b = pf*iRow) under rotation r: double matvec,
// then rounded to float exactly like the RFLOAT arguments of getByInterpolationFT / addFT
THB_HD void slice_coord(const Rot2& r, double a, double b, float& x, float& y, float& z)
{
    x = (float)(r.c0[0] * a + r.c1[0] * b);
    y = (float)(r.c0[1] * a + r.c1[1] * b);
    z = (float)(r.c0[2] * a + r.c1[2] * b);
}

// phase of translate(): RFLOAT rCol = tx / nCol ; phase = M_2X_PI * (iCol*rCol + iRow*rRow)
THB_HD float translate_phase(int iCol, int iRow, float rCol, float rRow)
{
    const float s = (float)iCol * rCol + (float)iRow * rRow;
    return (float)(6.28318530717959 * (double)s);
}

}  // namespace thb
