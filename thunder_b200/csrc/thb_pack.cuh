// thb_pack.cuh - packing kernel shared by thb_api.cu (thb_pack_stack) and thb_reco.cu (thb_remask_pack).
#pragma once
#include <cuda_runtime.h>
#include "thb_math.cuh"

namespace thb {

// a2: Optimiser::allocPreCal (src/Optimiser.cpp:8043-8171), image-major, OPTIMISER_CTF_ON_THE_FLY: the packed arrays of
// one image from its full half-complex FT, the per-group sigma table and its CTF attributes (CTF(): src/CTF.cpp:118-151,
// same float / double mix).  Output in the resident blocked order: d*[l][i] describes the caller's pixel perm[i].
struct CtfAttr7 { float voltage, defocusU, defocusV, theta, Cs, ac, phaseShift; };

static __global__ void pack_stack_kernel(const float2* __restrict__ imgFT, size_t imgStride, const int4* __restrict__ pix,
                                  const int* __restrict__ perm, const int* __restrict__ iPxl, const int* __restrict__ iSig, int P,
                                  const float* __restrict__ sigRcpTab, int nRing, const int* __restrict__ group,
                                  const CtfAttr7* __restrict__ attr, float pixelSize, int N, float2* __restrict__ ddat,
                                  float* __restrict__ dctf, float* __restrict__ dsig)
{
    const int l = blockIdx.y;
    const CtfAttr7 a = attr[l];
    const float lambda = (float)(12.2643247 / sqrt((double)a.voltage * (1 + (double)a.voltage * 0.978466e-6)));
    const float w1 = sqrtf(1 - (float)((double)a.ac * (double)a.ac));
    const float w2 = a.ac;
    const float K1 = (float)(3.14159265358979323846 * lambda);
    const float K2 = (float)(1.57079632679489661923 * a.Cs * (float)((double)lambda * lambda * lambda));
    const int g = group ? group[l] : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const int s = perm[i];
        const int4 c = pix[i];
        const size_t d = (size_t)l * P + i;
        ddat[d] = imgFT[(size_t)l * imgStride + iPxl[s]];
        if (dsig) dsig[d] = sigRcpTab[(size_t)g * nRing + iSig[s]];
        const float u = (float)hypot((double)((float)c.z / (pixelSize * (float)N)), (double)((float)c.w / (pixelSize * (float)N)));
        const float angle = (float)(atan2((double)c.w, (double)c.z) - (double)a.theta);
        // the phase reaches hundreds of radians: keep the reference's unfused operation order (no FMA contraction),
        // one ulp of ki is already 3e-5 in the CTF value
        const float defocus = __fmul_rn(-__fadd_rn(__fadd_rn(a.defocusU, a.defocusV), __fmul_rn(__fadd_rn(a.defocusU, -a.defocusV), cosf(__fmul_rn(2.0f, angle)))), 0.5f);
        const double u2d = (double)u * u;
        const float u2 = (float)u2d, u4 = (float)(u2d * u2d);
        const float ki = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(K1, defocus), u2), __fmul_rn(K2, u4)), -a.phaseShift);
        dctf[d] = __fadd_rn(__fmul_rn(-w1, sinf(ki)), __fmul_rn(w2, cosf(ki)));
    }
}

}  // namespace thb
