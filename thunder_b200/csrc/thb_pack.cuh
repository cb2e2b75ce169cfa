// thb_pack.cuh - packing kernel shared by thb_api.cu (thb_pack_stack) and thb_reco.cu (thb_remask_pack).
#pragma once
#include <cuda_runtime.h>
#include "thb_math.cuh"

namespace thb {

// a2: Optimiser::allocPreCal (src/Optimiser.cpp:8043-8171), image-major, OPTIMISER_CTF_ON_THE_FLY: the packed arrays of
// one image from its full half-complex FT, the per-group sigma table and its CTF attributes (CTF(): src/CTF.cpp:118-151,
// same float / double mix).  Output in the resident blocked order: d*[l][i] describes the caller's pixel perm[i].
struct CtfAttr7 { float voltage, defocusU, defocusV, theta, Cs, ac, phaseShift; };

// CTF() of src/CTF.cpp:118-151 for one pixel (iCol, iRow), split into what does not depend on the defocus values (the CTF search of
// the insert evaluates it once per draw with defocusU d, defocusV d) and the rest; same float / double mix and unfused order
struct CtfPixel { float u2, u4, c2; };     // |k|^2, |k|^4 in physical units, cos(2 (angle - theta))
// PRECISE: cos / sin through the double-precision library, rounded once - what glibc's cosf / sinf (correctly rounded in all but
// rare cases) give the reference; CUDA's cosf / sinf are 1 - 2 ulp off, and one ulp of cos(2 angle) moves a phase of hundreds of
// radians by 1e-5.  The packing kernel (once per image and iteration) uses it; the per-sample CTF of the insert does not.
template <bool PRECISE = false>
__device__ __forceinline__ CtfPixel ctf_pixel(int iCol, int iRow, float pixelSize, int N, float theta)
{
    const float u = (float)hypot((double)((float)iCol / (pixelSize * (float)N)), (double)((float)iRow / (pixelSize * (float)N)));
    const float angle = (float)(atan2((double)iRow, (double)iCol) - (double)theta);
    const double u2d = (double)u * u;
    CtfPixel c;
    c.u2 = (float)u2d; c.u4 = (float)(u2d * u2d);
    c.c2 = PRECISE ? (float)cos((double)__fmul_rn(2.0f, angle)) : cosf(__fmul_rn(2.0f, angle));
    return c;
}
struct CtfConst { float K1, K2, w1, w2, phaseShift; };
__device__ __forceinline__ CtfConst ctf_const(float voltage, float Cs, float ac, float phaseShift)
{
    const float lambda = (float)(12.2643247 / sqrt((double)voltage * (1 + (double)voltage * 0.978466e-6)));
    CtfConst k;
    k.w1 = sqrtf(1 - (float)((double)ac * (double)ac));
    k.w2 = ac;
    k.K1 = (float)(3.14159265358979323846 * lambda);
    k.K2 = (float)(1.57079632679489661923 * Cs * (float)((double)lambda * lambda * lambda));
    k.phaseShift = phaseShift;
    return k;
}
template <bool PRECISE = false>
__device__ __forceinline__ float ctf_eval(const CtfPixel& c, const CtfConst& k, float dU, float dV)
{
    const float defocus = __fmul_rn(-__fadd_rn(__fadd_rn(dU, dV), __fmul_rn(__fadd_rn(dU, -dV), c.c2)), 0.5f);
    const float ki = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(k.K1, defocus), c.u2), __fmul_rn(k.K2, c.u4)), -k.phaseShift);
    const float sn = PRECISE ? (float)sin((double)ki) : sinf(ki), cs = PRECISE ? (float)cos((double)ki) : cosf(ki);
    return __fadd_rn(__fmul_rn(-k.w1, sn), __fmul_rn(k.w2, cs));
}

static __global__ void pack_stack_kernel(const float2* __restrict__ imgFT, size_t imgStride, const int4* __restrict__ pix,
                                  const int* __restrict__ perm, const int* __restrict__ iPxl, const int* __restrict__ iSig, int P,
                                  const float* __restrict__ sigRcpTab, int nRing, const int* __restrict__ group,
                                  const CtfAttr7* __restrict__ attr, float pixelSize, int N, float2* __restrict__ ddat,
                                  float* __restrict__ dctf, float* __restrict__ dsig)
{
    const int l = blockIdx.y;
    const CtfAttr7 a = attr[l];
    const CtfConst k = ctf_const(a.voltage, a.Cs, a.ac, a.phaseShift);
    const int g = group ? group[l] : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const int s = perm[i];
        const int4 c = pix[i];
        const size_t d = (size_t)l * P + i;
        ddat[d] = imgFT[(size_t)l * imgStride + iPxl[s]];
        if (dsig) dsig[d] = sigRcpTab[(size_t)g * nRing + iSig[s]];
        // the phase reaches hundreds of radians: the reference's unfused operation order is kept (ctf_eval), one ulp of ki is
        // already 3e-5 in the CTF value - hence the correctly rounded cos / sin here
        dctf[d] = ctf_eval<true>(ctf_pixel<true>(c.z, c.w, pixelSize, N, a.theta), k, a.defocusU, a.defocusV);
    }
}

}  // namespace thb
