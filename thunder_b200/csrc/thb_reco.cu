// thb_reco.cu - SURVEY.md section 8(f) row 1: Reconstructor::reconstruct and Projector::setProjectee on the device,
// so that the loop  E-step -> insert -> all-reduce -> reconstruct -> new projector volume  never leaves HBM.
// cuFFT only for the 3D transforms (the north star's "once-per-round 3D inverse"); everything else is elementwise
// kernels over the half-complex grids.
//
// reference (MODE_3D, default Config.h switches; paths relative to the THUNDER tree):
//   reconstruct      src/Reconstructor.cpp:1129-1831   convoluteC :2595-2660   checkC :2522-2593 (CHECK_C_MAX)
//   prepareTF        src/Reconstructor.cpp:1056-1091   (normalisation by 1 / Re T[0]; C1: no symmetrisation)
//   kernel table     src/Reconstructor.cpp:55-90 (TabFunction of MKB_RL_R2 on [0,1], 1e5 bins), src/TabFunction.cpp:26-45
//   MKB_RL_R2        src/Functions/Functions.cpp:181-213 (FUNCTIONS_MKB_ORDER_0)      TIK_RL :236-239
//   setProjectee     src/Projector.cpp:123-148, gridCorrection :573-583
//   FFT conventions  src/FFT.cpp:346-376 (c2r unnormalised then scaled by 1/N; r2c unnormalised)
#include <cufft.h>
#include <cmath>
#include <cstring>
#include <climits>
#include <vector>
#include "thb_context.h"
#include "thb_pack.cuh"

namespace thb {

#define THB_FFT(ctx, call)                                                                   \
    do {                                                                                     \
        cufftResult r__ = (call);                                                            \
        if (r__ != CUFFT_SUCCESS) return set_error(ctx, THB_E_CUDA, "cuFFT error %d in %s", (int)r__, #call); \
    } while (0)

constexpr int RECO_TAB_N = 100000;

__device__ __forceinline__ int signed_coord(int mem, int n) { return mem >= n / 2 ? mem - n : mem; }

// (i, j, k) of element idx of a half-complex grid [m][m][m/2+1]
__device__ __forceinline__ void ft_coords(size_t idx, int m, int& i, int& j, int& k)
{
    const int nc = m / 2 + 1;
    i = (int)(idx % nc);
    const size_t row = idx / nc;
    j = signed_coord((int)(row % m), m);
    k = signed_coord((int)(row / m), m);
}

__global__ void reco_read_sf_kernel(const float4* acc, float* sf) { *sf = 1.0f / acc[0].z; }

__global__ void reco_scale_kernel(float4* acc, size_t nVox, const float* sf)
{
    const float s = *sf;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nVox; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = acc[i];
        v.x *= s; v.y *= s; v.z *= s;
        acc[i] = v;
    }
}

// MAP: T /= FSC' on WIENER_FACTOR_MIN_R*pf <= |k| < maxRadius*pf ; then W = [|k| < maxRadius*pf], T = max(T, 1e-25)
__global__ void reco_init_kernel(float4* acc, float* W, size_t nVox, int m, int pf, int maxRadius, const float* fsc, int nFsc, int joinHalf)
{
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nVox; idx += (size_t)gridDim.x * blockDim.x) {
        int i, j, k;
        ft_coords(idx, m, i, j, k);
        const long long r2 = (long long)i * i + (long long)j * j + (long long)k * k;
        const long long rmax2 = (long long)maxRadius * pf * maxRadius * pf;
        float4 v = acc[idx];
        if (fsc && r2 >= 25LL * pf * pf && r2 < rmax2) {
            const int u = (int)rint(sqrt((double)r2)) / pf;
            float f = u >= nFsc ? 0.0f : fsc[u];
            f = fmaxf(1e-3f, fminf(1.0f - 1e-3f, f));
            if (joinHalf) f = sqrtf(2 * f / (1 + f));
            v.z = v.z / f;
        }
        v.z = fmaxf(v.z, 1e-25f);
        acc[idx] = v;
        W[idx] = r2 < rmax2 ? 1.0f : 0.0f;
    }
}

// The x = 0 and x = m/2 planes of a half-complex array hold both members of each Hermitian pair, and the scatter does
// not keep them conjugate.  FFTW's c2r (the reference) transforms y,z first and then drops the imaginary part of
// those two planes, which equals replacing P(0,j,k) by (P(0,j,k) + conj P(0,-j,-k)) / 2.  cuFFT may take another
// route for some sizes, so the planes are made consistent explicitly: identical to FFTW for any input.
__global__ void reco_hermitian_planes_kernel(float2* C, int m, int nz)      // nz = m, or 1 for an image (MODE_2D)
{
    const int nc = m / 2 + 1;
    const size_t plane = (size_t)m * nz;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * plane; t += (size_t)gridDim.x * blockDim.x) {
        const int i = t < plane ? 0 : m / 2;
        const size_t r = t < plane ? t : t - plane;
        const int jm = (int)(r % m), km = (int)(r / m);
        const int jp = (m - jm) % m, kp = (m - km) % m;
        const size_t a = ((size_t)km * m + jm) * nc + i, b = ((size_t)kp * m + jp) * nc + i;
        if (a > b) continue;                       // one thread per pair (a == b: self-conjugate entries)
        const float2 va = C[a], vb = C[b];
        const float2 na = make_float2(0.5f * (va.x + vb.x), 0.5f * (va.y - vb.y));
        C[a] = na;
        C[b] = make_float2(na.x, -na.y);
    }
}

__global__ void reco_make_c_kernel(const float4* acc, const float* W, float2* C, size_t nVox)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nVox; i += (size_t)gridDim.x * blockDim.x)
        C[i] = make_float2(acc[i].z * W[i], 0.0f);
}

// real space: C(x) *= kernelRL(|x|^2 / (N pf)^2) / nf, with the 1/m^3 of the backward transform
__global__ void reco_kernel_rl_kernel(float* c, int m, int nz, const float* tab, float tabStep, float nf, float M2, float invVol)
{
    const size_t total = (size_t)m * m * nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = signed_coord((int)(idx % m), m);
        const size_t row = idx / m;
        const int j = signed_coord((int)(row % m), m), k = signed_coord((int)(row / m), m);
        const float x = (float)((double)((long long)i * i + (long long)j * j + (long long)k * k) / (double)M2);
        const int t = (int)rint((double)(x / tabStep));
        c[idx] = (c[idx] * invVol) * tab[min(t, RECO_TAB_N)] / nf;
    }
}

// W /= max(|C|, 1e-6) inside the radius; diff = max | |C| - 1 | there (non-negative floats order like their bit patterns)
__global__ void reco_update_w_kernel(float* W, const float2* C, size_t nVox, int m, int pf, int maxRadius, int* diffBits)
{
    float local = 0.0f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nVox; idx += (size_t)gridDim.x * blockDim.x) {
        int i, j, k;
        ft_coords(idx, m, i, j, k);
        const long long r2 = (long long)i * i + (long long)j * j + (long long)k * k;
        if (r2 < (long long)maxRadius * pf * maxRadius * pf) {
            const float2 c = C[idx];
            const float a = (float)hypot((double)c.x, (double)c.y);
            W[idx] = W[idx] / fmaxf(a, 1e-6f);
            local = fmaxf(local, fabsf(a - 1.0f));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local = fmaxf(local, __shfl_xor_sync(0xffffffffu, local, o));
    if ((threadIdx.x & 31) == 0) atomicMax(diffBits, __float_as_int(local));
}

__global__ void reco_w_nogrid_kernel(float* W, const float4* acc, size_t nVox, int m, int pf, int maxRadius)
{
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nVox; idx += (size_t)gridDim.x * blockDim.x) {
        int i, j, k;
        ft_coords(idx, m, i, j, k);
        const long long r2 = (long long)i * i + (long long)j * j + (long long)k * k;
        if (r2 < (long long)maxRadius * pf * maxRadius * pf) W[idx] = 1.0f / fmaxf(fabsf(acc[idx].z), 1e-6f);
    }
}

// padDst (M grid, zeroed) = F * W inside the radius, at the same signed frequencies
__global__ void reco_pad_ft_kernel(const float4* acc, const float* W, size_t nVox, int m, int M, int pf, int maxRadius, float2* pad)
{
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nVox; idx += (size_t)gridDim.x * blockDim.x) {
        int i, j, k;
        ft_coords(idx, m, i, j, k);
        const long long r2 = (long long)i * i + (long long)j * j + (long long)k * k;
        if (r2 < (long long)maxRadius * pf * maxRadius * pf) {
            const float4 v = acc[idx];
            const float w = W[idx];
            const size_t o = ((size_t)(k < 0 ? k + M : k) * M + (size_t)(j < 0 ? j + M : j)) * (M / 2 + 1) + i;
            pad[o] = make_float2(v.x * w, v.y * w);
        }
    }
}

// dst (N^3) = central part of padReal (M^3, scaled by 1/M^3) / TIK_RL(|x| / (pf N))
__global__ void reco_extract_kernel(const float* padReal, int M, int N, int nz, int pf, float invVol, float* dst)
{
    const size_t total = (size_t)N * N * nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = signed_coord((int)(idx % N), N);
        const size_t row = idx / N;
        const int j = signed_coord((int)(row % N), N), k = signed_coord((int)(row / N), N);
        const size_t src = ((size_t)(k < 0 ? k + M : k) * M + (size_t)(j < 0 ? j + M : j)) * M + (size_t)(i < 0 ? i + M : i);
        const double r = sqrt((double)((long long)i * i + (long long)j * j + (long long)k * k)) / (double)(pf * N);
        const double x = 3.14159265358979323846 * r;
        const double j0 = x == 0.0 ? 1.0 : sin(x) / x;
        dst[idx] = (padReal[src] * invVol) / (float)(j0 * j0);
    }
}

// setProjectee: pad (n^3, zeroed) <- vol (N^3) at the same signed coordinates, / TIK_RL(|x| / (pf n))
__global__ void proj_pad_rl_kernel(const float* vol, int N, int nz, int n, int pf, float* pad)
{
    const size_t total = (size_t)N * N * nz;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = signed_coord((int)(idx % N), N);
        const size_t row = idx / N;
        const int j = signed_coord((int)(row % N), N), k = signed_coord((int)(row / N), N);
        const size_t dst = ((size_t)(k < 0 ? k + n : k) * n + (size_t)(j < 0 ? j + n : j)) * n + (size_t)(i < 0 ? i + n : i);
        const double r = sqrt((double)((long long)i * i + (long long)j * j + (long long)k * k)) / ((double)pf * n);
        const double x = 3.14159265358979323846 * r;
        const double j0 = x == 0.0 ? 1.0 : sin(x) / x;
        pad[dst] = vol[idx] / (float)(j0 * j0);
    }
}

// ---- section 8(f) row 2: reCentreImg + reMaskImg on a batch of full half-complex image FTs [nImg][N][N/2+1]
// translate(Image&, const Image&, ...) (src/Image/ImageFunctions.cpp:269-284): img = ori * polar(-2 pi (i rCol + j rRow))
__global__ void img_translate_kernel(float2* img, int N, int nImg, const float* rColRow)
{
    const int nc = N / 2 + 1;
    const size_t per = (size_t)nc * N, total = per * nImg;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(idx / per);
        const size_t r = idx - (size_t)l * per;
        const int i = (int)(r % nc), j = signed_coord((int)(r / nc), N);
        const float ph = translate_phase(i, j, rColRow[2 * l], rColRow[2 * l + 1]);
        float sn, cs;
        sincosf(ph, &sn, &cs);
        const float2 v = img[idx];
        // v * (cos(-ph), sin(-ph))
        img[idx] = make_float2(v.x * cs + v.y * sn, v.y * cs - v.x * sn);
    }
}

// x = 0 and x = N/2 columns of every image made Hermitian-consistent (what FFTW's c2r does implicitly, see above)
__global__ void img_hermitian_cols_kernel(float2* img, int N, int nImg)
{
    const int nc = N / 2 + 1;
    const size_t total = (size_t)2 * N * nImg;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(t / (2 * N));
        const int r = (int)(t % (2 * N));
        const int i = r < N ? 0 : N / 2, jm = r < N ? r : r - N, jp = (N - jm) % N;
        if (jm > jp) continue;
        float2* base = img + (size_t)l * nc * N;
        const float2 va = base[(size_t)jm * nc + i], vb = base[(size_t)jp * nc + i];
        const float2 na = make_float2(0.5f * (va.x + vb.x), 0.5f * (va.y - vb.y));
        base[(size_t)jm * nc + i] = na;
        base[(size_t)jp * nc + i] = make_float2(na.x, -na.y);
    }
}

// real space: x softMask(r, ew) (src/Functions/Mask.cpp:333-350) with the 1/N^2 of the backward transform
__global__ void img_mask_kernel(float* rl, int N, int nImg, float r, float ew)
{
    const size_t per = (size_t)N * N, total = per * nImg;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t q = idx % per;
        const int i = signed_coord((int)(q % N), N), j = signed_coord((int)(q / N), N);
        const float u = (float)hypot((double)i, (double)j);
        float w;
        if (u > r + ew) w = 0.0f;
        else if (u >= r) w = (float)(0.5 + 0.5 * cos((double)((u - r) / ew) * 3.14159265358979323846));
        else w = 1.0f;
        rl[idx] = (rl[idx] * (1.0f / (float)per)) * w;
    }
}

// ---- section 8(f) row 3: per-image part of Optimiser::allReduceSigma (src/Optimiser.cpp:6428-6600) on the resident stacks.
// One CTA per image.  E-stack pixel i (ring uE): model = ctf * polar(-phase(t)) * slice ; M-stack pixel (ring uM < rSig,
// |k|^2 < rSig^2): the same slice at t - offset.  Ring sums in shared memory, then ring averages (the ring populations
// cntE / cntM are properties of the pixel lists) and the per-group sums the reference all-reduces.
__device__ __forceinline__ float2 gather_lin(const float2* __restrict__ vol, int n, int pitch, float x, float y, float z)
{
    int x0, y0, z0;
    float xd, yd, zd;
    const bool conj = fold_floor(x, y, z, x0, y0, z0, xd, yd, zd);
    float w[8];
    tri_weights(xd, yd, zd, w);
    int64_t off[4];
    row_offsets(y0, z0, n, pitch, off);
    float re = 0.0f, im = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float2 a = __ldg(vol + off[c] + x0), b = __ldg(vol + off[c] + x0 + 1);
        re += a.x * w[2 * c]; im += a.y * w[2 * c];
        re += b.x * w[2 * c + 1]; im += b.y * w[2 * c + 1];
    }
    return make_float2(re, conj ? -im : im);
}

#define SIGMA_DUP (1 << 30)
struct SigmaArgs {
    const float2* vols[THB_MAX_SLOTS];
    int vdim, pitch, N, rSig, nGroup, mode2D;
    const float2* datE; const float* ctfE; const int* slotE; const int4* pixE; const int* ringE; int PE;
    const float2* datM; const float* ctfM; const int4* pixM; const int* ringM; int PM;   // ringM < 0: pixel not in the sigma set
    const int* imgIdx; const double* quat; const double* tran; const double* offS; const int* group;
    const float* cntE; const float* cntM;
    double* sigM; double* sigN; double* svd;      // [nGroup][rSig + 1]
};

__global__ void __launch_bounds__(256) sigma_kernel(const SigmaArgs A)
{
    extern __shared__ float sRing[];               // [4][rSig]: vSigM, sSVD, dSVD, vSigN
    const int l = blockIdx.x, tid = threadIdx.x, rS = A.rSig;
    const int img = A.imgIdx ? A.imgIdx[l] : l;
    for (int i = tid; i < 4 * rS; i += blockDim.x) sRing[i] = 0.0f;
    __syncthreads();
    const float2* __restrict__ vol = A.vols[A.slotE ? A.slotE[img] : 0];
    double q[4] = {1.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < (A.mode2D ? 2 : 4); ++c) q[c] = A.quat[(A.mode2D ? 2 : 4) * l + c];
    const Rot2 rot = make_rot2(q, A.mode2D);
    const float tx = (float)A.tran[2 * l], ty = (float)A.tran[2 * l + 1];
    const float ox = (float)(A.tran[2 * l] - A.offS[2 * l]), oy = (float)(A.tran[2 * l + 1] - A.offS[2 * l + 1]);
    for (int i = tid; i < A.PE; i += blockDim.x) {
        int u = A.ringE[i];
        if (u < 0) continue;
        const float wgt = (u & SIGMA_DUP) ? 2.0f : 1.0f;
        u &= ~SIGMA_DUP;
        const int4 c = A.pixE[i];
        float x, y, z;
        slice_coord(rot, (double)c.x, (double)c.y, x, y, z);
        const float2 p = gather_lin(vol, A.vdim, A.pitch, x, y, z);
        const float ph = translate_phase(c.z, c.w, tx / (float)A.N, ty / (float)A.N);
        float sn, cs;
        sincosf(ph, &sn, &cs);
        // p * polar(-ph), then * ctf
        const float cf = A.ctfE[(size_t)img * A.PE + i];
        const float mx = (p.x * cs + p.y * sn) * cf, my = (p.y * cs - p.x * sn) * cf;
        const float2 d = A.datE[(size_t)img * A.PE + i];
        const float rx = d.x - mx, ry = d.y - my;
        atomicAdd(&sRing[u], wgt * (rx * rx + ry * ry));
        atomicAdd(&sRing[rS + u], wgt * (mx * mx + my * my));
        atomicAdd(&sRing[2 * rS + u], wgt * (d.x * d.x + d.y * d.y));
    }
    for (int i = tid; i < A.PM; i += blockDim.x) {
        int u = A.ringM[i];
        if (u < 0) continue;
        const float wgt = (u & SIGMA_DUP) ? 2.0f : 1.0f;
        u &= ~SIGMA_DUP;
        const int4 c = A.pixM[i];
        float x, y, z;
        slice_coord(rot, (double)c.x, (double)c.y, x, y, z);
        const float2 p = gather_lin(vol, A.vdim, A.pitch, x, y, z);
        const float ph = translate_phase(c.z, c.w, ox / (float)A.N, oy / (float)A.N);
        float sn, cs;
        sincosf(ph, &sn, &cs);
        const float cf = A.ctfM[(size_t)img * A.PM + i];
        const float mx = (p.x * cs + p.y * sn) * cf, my = (p.y * cs - p.x * sn) * cf;
        const float2 d = A.datM[(size_t)img * A.PM + i];
        const float rx = d.x - mx, ry = d.y - my;
        atomicAdd(&sRing[3 * rS + u], wgt * (rx * rx + ry * ry));
    }
    __syncthreads();
    const int g = A.group ? A.group[l] : 0;
    const size_t row = (size_t)g * (rS + 1);
    for (int u = tid; u < rS; u += blockDim.x) {
        const float cE = A.cntE[u], cM = A.cntM[u];
        if (cE > 0.0f) {
            const float vM = sRing[u] / cE, sS = sRing[rS + u] / cE, dS = sRing[2 * rS + u] / cE;
            atomicAdd(&A.sigM[row + u], (double)(vM / 2));
            atomicAdd(&A.svd[row + u], (double)sqrtf(sS / dS));
        }
        if (cM > 0.0f) atomicAdd(&A.sigN[row + u], (double)((sRing[3 * rS + u] / cM) / 2));
    }
    if (tid == 0) {
        atomicAdd(&A.sigM[row + rS], 1.0);
        atomicAdd(&A.sigN[row + rS], 1.0);
        atomicAdd(&A.svd[row + rS], 1.0);
    }
}

// Reconstructor::symmetrizeF / symmetrizeT (src/Reconstructor.cpp:2676-2690) = SYMMETRIZE_FT (include/Geometry/
// Transformation.h:105-131, 170-194): dst(v) = src(v) + sum_e src interpolated at R_e v, for every voxel v of the half volume
// whose rotated position lies inside radius r.  F (complex, conjugated on the Hermitian fold) and T (real) of a voxel travel
// together in the float4 accumulator, so one gather serves both.
__global__ void __launch_bounds__(256) symmetrize_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int n, int nElem,
                                                         const double* __restrict__ R, double r2)
{
    const int nColFT = n / 2 + 1;
    const size_t total = (size_t)nColFT * n * n;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % nColFT);
        const size_t row = idx / nColFT;
        const int jm = (int)(row % n), km = (int)(row / n);
        const double a = (double)i, b = (double)(jm < n / 2 ? jm : jm - n), c = (double)(km < n / 2 ? km : km - n);
        float4 v = src[idx];
        for (int e = 0; e < nElem; ++e) {
            const double* m = R + 9 * e;                               // column-major, as Eigen's dmat33
            const double ox = m[0] * a + m[3] * b + m[6] * c, oy = m[1] * a + m[4] * b + m[7] * c, oz = m[2] * a + m[5] * b + m[8] * c;
            if (!(ox * ox + oy * oy + oz * oz < r2)) continue;
            float x = (float)ox, y = (float)oy, z = (float)oz;
            int x0, y0, z0;
            float xd, yd, zd;
            const bool conj = fold_floor(x, y, z, x0, y0, z0, xd, yd, zd);
            float w[8];
            tri_weights(xd, yd, zd, w);
            int64_t off[4];
            row_offsets(y0, z0, n, nColFT, off);
            float re = 0.0f, im = 0.0f, t = 0.0f;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float4 p = __ldg(src + off[cc] + x0), q = __ldg(src + off[cc] + x0 + 1);
                re += p.x * w[2 * cc]; im += p.y * w[2 * cc]; t += p.z * w[2 * cc];
                re += q.x * w[2 * cc + 1]; im += q.y * w[2 * cc + 1]; t += q.z * w[2 * cc + 1];
            }
            v.x += re; v.y += conj ? -im : im; v.z += t;
        }
        dst[idx] = v;
    }
}

// Optimiser::normCorrection (src/Optimiser.cpp:6201-6393, OPTIMISER_NORM_MASK): norm[l] = sum over the half plane
// {rL^2 <= |k|^2 < rNorm^2} of |masked image - ctf * translated slice at the image's best orientation|^2; the mirror pixel of
// the i = 0 column counted twice, as in sigma_kernel
__global__ void __launch_bounds__(256) norm_kernel(const SigmaArgs A, float rL2, float rN2, double* __restrict__ norm)
{
    __shared__ double red[8];
    const int l = blockIdx.x, tid = threadIdx.x;
    const int img = A.imgIdx ? A.imgIdx[l] : l;
    const float2* __restrict__ vol = A.vols[A.slotE ? A.slotE[img] : 0];
    double q[4] = {1.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < (A.mode2D ? 2 : 4); ++c) q[c] = A.quat[(A.mode2D ? 2 : 4) * l + c];
    const Rot2 rot = make_rot2(q, A.mode2D);
    const float tx = (float)A.tran[2 * l], ty = (float)A.tran[2 * l + 1];
    float s = 0.0f;
    for (int i = tid; i < A.PE; i += blockDim.x) {
        const int4 c = A.pixE[i];
        const float u = (float)(c.z * c.z + c.w * c.w);
        if (!(u >= rL2 && u < rN2)) continue;
        float x, y, z;
        slice_coord(rot, (double)c.x, (double)c.y, x, y, z);
        const float2 p = gather_lin(vol, A.vdim, A.pitch, x, y, z);
        const float ph = translate_phase(c.z, c.w, tx / (float)A.N, ty / (float)A.N);
        float sn, cs;
        sincosf(ph, &sn, &cs);
        const float cf = A.ctfE[(size_t)img * A.PE + i];
        const float mx = (p.x * cs + p.y * sn) * cf, my = (p.y * cs - p.x * sn) * cf;
        const float2 d = A.datE[(size_t)img * A.PE + i];
        const float rx = d.x - mx, ry = d.y - my;
        s += ((c.z == 0 && c.w > 0) ? 2.0f : 1.0f) * (rx * rx + ry * ry);
    }
    double v = (double)s;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        norm[l] = t;
    }
}

// _img[l] *= s, _imgOri[l] *= s (src/Optimiser.cpp:6381-6391) on the resident stacks
__global__ void scale_images_kernel(float2* __restrict__ datE, int PE, float2* __restrict__ datM, int PM, const int* __restrict__ imgIdx,
                                    const float* __restrict__ scale)
{
    const int l = blockIdx.y;
    const int img = imgIdx ? imgIdx[l] : l;
    const float s = scale[l];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(PE, PM); i += gridDim.x * blockDim.x) {
        if (datE && i < PE) { float2 v = datE[(size_t)img * PE + i]; datE[(size_t)img * PE + i] = make_float2(v.x * s, v.y * s); }
        if (datM && i < PM) { float2 v = datM[(size_t)img * PM + i]; datM[(size_t)img * PM + i] = make_float2(v.x * s, v.y * s); }
    }
}

__global__ void reco_upload_kernel(const float2* F, const float* T, size_t nVox, float4* acc)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nVox; i += (size_t)gridDim.x * blockDim.x)
        acc[i] = make_float4(F[i].x, F[i].y, T[i], 0.0f);
}

// ---- host: the kernel table.  MKB_RL_R2 of order 0: (2 pi)^1.5 a^3 / I0(alpha) / v^1.5 * (I_1.5(v) | J_1.5(v)),
// with the closed forms of the half-integer Bessel functions.
static double bessel_i0(double x)
{
    double s = 1.0, t = 1.0;
    for (int k = 1; k < 200; ++k) {
        t *= (x / (2.0 * k)) * (x / (2.0 * k));
        s += t;
        if (t < 1e-17 * s) break;
    }
    return s;
}

static double mkb_rl_r2(double r2, double a, double alpha)
{
    const double u2 = (2 * M_PI * a) * (2 * M_PI * a) * r2;
    const bool inside = u2 <= alpha * alpha;
    const double v = std::sqrt(inside ? alpha * alpha - u2 : u2 - alpha * alpha);
    const double pre = std::pow(2 * M_PI, 1.5) * a * a * a / bessel_i0(alpha);
    // B(v) / v^1.5 with B = I_1.5 or J_1.5 ; both tend to sqrt(2/pi)/3 as v -> 0
    double q;
    if (v < 1e-4)
        q = std::sqrt(2.0 / M_PI) / 3.0;
    else if (inside)
        q = std::sqrt(2.0 / (M_PI * v)) * (std::cosh(v) - std::sinh(v) / v) / std::pow(v, 1.5);
    else
        q = std::sqrt(2.0 / (M_PI * v)) * (std::sin(v) / v - std::cos(v)) / std::pow(v, 1.5);
    return pre * q;
}

struct RecoState {
    cufftHandle planC2R = 0, planR2C = 0;
    int planDim = 0, planRank = 0;
    cufftHandle planProj = 0;
    int projDim = 0, projPitch = 0, projRank = 0;
    float* dTab = nullptr;
    double tabA = -1, tabAlpha = -1;
    float nf = 1.0f;
    float* dVol = nullptr;      // last reconstruction, N^3 real
    cufftHandle planImgC2R = 0, planImgR2C = 0;
    int imgN = 0, imgBatch = 0;
    int volN = 0, volRank = 3;
};

static RecoState* reco_state(thb_ctx* ctx)
{
    if (!ctx->reco) ctx->reco = new RecoState();
    return static_cast<RecoState*>(ctx->reco);
}

void reco_free(thb_ctx* ctx)
{
    if (!ctx->reco) return;
    RecoState* s = static_cast<RecoState*>(ctx->reco);
    if (s->planC2R) cufftDestroy(s->planC2R);
    if (s->planR2C) cufftDestroy(s->planR2C);
    if (s->planProj) cufftDestroy(s->planProj);
    if (s->planImgC2R) cufftDestroy(s->planImgC2R);
    if (s->planImgR2C) cufftDestroy(s->planImgR2C);
    cudaFree(s->dTab);
    cudaFree(s->dVol);
    delete s;
    ctx->reco = nullptr;
}

static int ensure_plans(thb_ctx* ctx, RecoState* s, int m, int rank)
{
    if (s->planDim == m && s->planRank == rank) return THB_OK;
    if (s->planC2R) { cufftDestroy(s->planC2R); s->planC2R = 0; }
    if (s->planR2C) { cufftDestroy(s->planR2C); s->planR2C = 0; }
    s->planDim = 0;
    if (rank == 3) {
        THB_FFT(ctx, cufftPlan3d(&s->planC2R, m, m, m, CUFFT_C2R));
        THB_FFT(ctx, cufftPlan3d(&s->planR2C, m, m, m, CUFFT_R2C));
    } else {
        THB_FFT(ctx, cufftPlan2d(&s->planC2R, m, m, CUFFT_C2R));
        THB_FFT(ctx, cufftPlan2d(&s->planR2C, m, m, CUFFT_R2C));
    }
    s->planRank = rank;
    THB_FFT(ctx, cufftSetStream(s->planC2R, ctx->stream));
    THB_FFT(ctx, cufftSetStream(s->planR2C, ctx->stream));
    s->planDim = m;
    return THB_OK;
}

static int ensure_table(thb_ctx* ctx, RecoState* s, double a, double alpha)
{
    if (s->dTab && s->tabA == a && s->tabAlpha == alpha) return THB_OK;
    std::vector<float> tab(RECO_TAB_N + 1);
    const float step = 1.0f / (float)RECO_TAB_N;
    for (int i = 0; i <= RECO_TAB_N; ++i) tab[i] = (float)mkb_rl_r2((double)((float)i * step), a, alpha);
    if (!s->dTab) THB_CUDA(ctx, cudaMalloc(&s->dTab, sizeof(float) * (RECO_TAB_N + 1)));
    THB_CUDA(ctx, cudaMemcpy(s->dTab, tab.data(), sizeof(float) * (RECO_TAB_N + 1), cudaMemcpyHostToDevice));
    s->nf = (float)mkb_rl_r2(0.0, a, alpha);
    s->tabA = a;
    s->tabAlpha = alpha;
    return THB_OK;
}

}  // namespace thb

using namespace thb;

extern "C" {

int thb_reco_upload(thb_ctx* ctx, int slot, const float* F, const float* T)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || !ctx->accs[slot].d) return set_error(ctx, THB_E_STATE, "reco_upload: slot %d not allocated", slot);
    if (!F || !T) return set_error(ctx, THB_E_ARG, "reco_upload: NULL arrays");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const Accum& a0 = ctx->accs[slot];
    struct { float4* d; size_t nVox; } a = {a0.d, ctx->mode2D ? a0.nVox / 2 : a0.nVox};     // MODE_2D: F, T are images = plane 0
    if (ctx->mode2D) THB_CUDA(ctx, cudaMemsetAsync(a0.d, 0, a0.nVox * sizeof(float4), ctx->stream));
    float2* dF = (float2*)scratch(ctx, 2, a.nVox * sizeof(float2));
    float* dT = (float*)scratch(ctx, 3, a.nVox * sizeof(float));
    if (!dF || !dT) return THB_E_CUDA;
    THB_CUDA(ctx, cudaMemcpyAsync(dF, F, a.nVox * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dT, T, a.nVox * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    reco_upload_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(dF, dT, a.nVox, a.d);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_reconstruct(thb_ctx* ctx, int slot, int N, int pf, double a, double alpha, int gridCorr, int joinHalf, const float* fsc,
                    int nFsc, int normalise, float* dstReal, int* nIterOut)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || !ctx->accs[slot].d) return set_error(ctx, THB_E_STATE, "reconstruct: slot %d not allocated", slot);
    // MODE_2D: the same steps on images (Reconstructor::reconstruct, the MODE_2D branches of src/Reconstructor.cpp:1129-1831):
    // plane 0 of the two-plane accumulator, 2D transforms, dstReal = the N x N class average
    const int rank = ctx->mode2D ? 2 : 3;
    const Accum& acc = ctx->accs[slot];
    const int m = acc.vdim;
    if (N <= 0 || (N & 1) || pf <= 0 || m % pf || m / pf > N) return set_error(ctx, THB_E_ARG, "reconstruct: accumulator dimension %d does not fit N = %d, pf = %d", m, N, pf);
    if (fsc && nFsc <= 0) return set_error(ctx, THB_E_ARG, "reconstruct: empty FSC");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    RecoState* s = reco_state(ctx);
    const int size = m / pf, M = N * pf;
    const int maxRadius = size / 2 - (int)std::ceil(a);       // Reconstructor::init, src/Reconstructor.cpp:88
    if (maxRadius <= 0) return set_error(ctx, THB_E_ARG, "reconstruct: volume too small for the kernel radius");
    const int mz = rank == 3 ? m : 1, Mz = rank == 3 ? M : 1, Nz = rank == 3 ? N : 1;
    const size_t nVox = rank == 3 ? acc.nVox : acc.nVox / 2, nReal = (size_t)m * m * mz;
    const size_t nVoxM = (size_t)(M / 2 + 1) * M * Mz, nRealM = (size_t)M * M * Mz;
    int rc;
    if ((rc = ensure_table(ctx, s, a, alpha))) return rc;
    float* W = (float*)scratch(ctx, 1, nVox * sizeof(float));
    float2* C = (float2*)scratch(ctx, 2, std::max(nVox, nVoxM) * sizeof(float2));
    float* R = (float*)scratch(ctx, 3, std::max(nReal, nRealM) * sizeof(float));
    float* dSmall = (float*)scratch(ctx, 0, sizeof(float) * (size_t)(std::max(nFsc, 1) + 8));
    if (!W || !C || !R || !dSmall) return THB_E_CUDA;
    float* dSf = dSmall; int* dDiff = (int*)(dSmall + 1); float* dFsc = dSmall + 8;
    if (fsc) THB_CUDA(ctx, cudaMemcpyAsync(dFsc, fsc, sizeof(float) * nFsc, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = ctx->smCount * 8;
    span_begin(ctx, KF_PACK);
    if (normalise) {
        reco_read_sf_kernel<<<1, 1, 0, ctx->stream>>>(acc.d, dSf);
        reco_scale_kernel<<<grid, 256, 0, ctx->stream>>>(acc.d, nVox, dSf);
        ctx->launches += 2;
    }
    reco_init_kernel<<<grid, 256, 0, ctx->stream>>>(acc.d, W, nVox, m, pf, maxRadius, fsc ? dFsc : nullptr, nFsc, joinHalf);
    ctx->launches++;
    int nIter = 0;
    if (gridCorr) {
        if ((rc = ensure_plans(ctx, s, m, rank))) return rc;
        float diffC = 3.40282347e38f, diffPrev;
        int noDecrease = 0;
        for (int it = 0; it < 30; ++it) {                     // MAX_N_ITER_BALANCE
            reco_make_c_kernel<<<grid, 256, 0, ctx->stream>>>(acc.d, W, C, nVox);
            reco_hermitian_planes_kernel<<<grid, 256, 0, ctx->stream>>>(C, m, mz);
            THB_FFT(ctx, cufftExecC2R(s->planC2R, reinterpret_cast<cufftComplex*>(C), R));
            reco_kernel_rl_kernel<<<grid, 256, 0, ctx->stream>>>(R, m, mz, s->dTab, 1.0f / (float)RECO_TAB_N, s->nf, (float)M * (float)M,
                                                                 (float)(1.0 / (double)nReal));
            THB_FFT(ctx, cufftExecR2C(s->planR2C, R, reinterpret_cast<cufftComplex*>(C)));
            THB_CUDA(ctx, cudaMemsetAsync(dDiff, 0, sizeof(int), ctx->stream));
            reco_update_w_kernel<<<grid, 256, 0, ctx->stream>>>(W, C, nVox, m, pf, maxRadius, dDiff);
            ctx->launches += 4;
            int bits = 0;
            THB_CUDA(ctx, cudaMemcpyAsync(&bits, dDiff, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            diffPrev = diffC;
            memcpy(&diffC, &bits, sizeof(float));
            nIter = it + 1;
            noDecrease = diffC > diffPrev * 0.95f ? noDecrease + 1 : 0;               // DIFF_C_DECREASE_THRES
            if (diffC < 1e-2f || (it >= 10 && noDecrease == 2)) break;                 // DIFF_C_THRES, MIN_N_ITER_BALANCE, N_DIFF_C_NO_DECREASE
        }
    } else {
        reco_w_nogrid_kernel<<<grid, 256, 0, ctx->stream>>>(W, acc.d, nVox, m, pf, maxRadius);
        ctx->launches++;
    }
    // padDst = F * W on the (N pf)^3 grid, inverse transform, crop, sinc^2 correction
    if ((rc = ensure_plans(ctx, s, M, rank))) return rc;
    THB_CUDA(ctx, cudaMemsetAsync(C, 0, nVoxM * sizeof(float2), ctx->stream));
    reco_pad_ft_kernel<<<grid, 256, 0, ctx->stream>>>(acc.d, W, nVox, m, M, pf, maxRadius, C);
    reco_hermitian_planes_kernel<<<grid, 256, 0, ctx->stream>>>(C, M, Mz);
    THB_FFT(ctx, cufftExecC2R(s->planC2R, reinterpret_cast<cufftComplex*>(C), R));
    if (s->volN != N || s->volRank != rank) {
        cudaFree(s->dVol);
        s->dVol = nullptr;
        s->volN = 0;
        THB_CUDA(ctx, cudaMalloc(&s->dVol, sizeof(float) * (size_t)N * N * Nz));
        s->volN = N;
        s->volRank = rank;
    }
    reco_extract_kernel<<<grid, 256, 0, ctx->stream>>>(R, M, N, Nz, pf, (float)(1.0 / (double)nRealM), s->dVol);
    ctx->launches += 3;
    span_end(ctx);
    THB_CUDA(ctx, cudaGetLastError());
    if (dstReal) THB_CUDA(ctx, cudaMemcpyAsync(dstReal, s->dVol, sizeof(float) * (size_t)N * N * Nz, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (nIterOut) *nIterOut = nIter;
    return THB_OK;
}

// section 8(f) row 2: E-stack images [base, base+nImg) rebuilt from the ORIGINAL image FTs: translate by the running
// offset (reCentreImg), soft mask in real space (reMaskImg, batched 2D cuFFT), then the packing of thb_pack_stack.
int thb_remask_pack(thb_ctx* ctx, int base, int nImg, const float* imgOriFT, const double* offset, float maskRadiusPx, int zeroMask,
                    const int* iPxl, const int* iSig, const float* sigRcpTab, int nGroup, int nRing, const int* groupOfImg,
                    const float* ctfAttr, float pixelSize, const int* slotOfImg, float* imgOutFT)
{
    if (!ctx) return THB_E_ARG;
    const int P = ctx->nPxlE, N = ctx->N;
    Stack& st = ctx->stackE;
    if (P <= 0) return set_error(ctx, THB_E_STATE, "remask_pack: E pixel list not set");
    if (!st.dat) return set_error(ctx, THB_E_STATE, "remask_pack: E stack not reserved (thb_stack_reserve)");
    if (nImg <= 0 || !imgOriFT || !offset || !iPxl || !iSig || !sigRcpTab || !ctfAttr || nGroup <= 0 || nRing <= 0 || pixelSize <= 0)
        return set_error(ctx, THB_E_ARG, "remask_pack: bad arguments");
    if (base < 0 || base + nImg > st.nImg) return set_error(ctx, THB_E_ARG, "remask_pack: images [%d,%d) exceed the capacity %d", base, base + nImg, st.nImg);
    const size_t imgElems = (size_t)(N / 2 + 1) * N;
    for (int i = 0; i < P; ++i)
        if (iPxl[i] < 0 || (size_t)iPxl[i] >= imgElems || iSig[i] < 0 || iSig[i] >= nRing) return set_error(ctx, THB_E_ARG, "remask_pack: iPxl / iSig[%d] out of range", i);
    for (int l = 0; l < nImg; ++l) {
        if (groupOfImg && (groupOfImg[l] < 0 || groupOfImg[l] >= nGroup)) return set_error(ctx, THB_E_ARG, "remask_pack: groupOfImg[%d] out of range", l);
        if (slotOfImg && (slotOfImg[l] < 0 || slotOfImg[l] >= THB_MAX_SLOTS)) return set_error(ctx, THB_E_ARG, "remask_pack: slotOfImg[%d] out of range", l);
    }
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    RecoState* s = reco_state(ctx);
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nImg, ((size_t)256 << 20) / (imgElems * sizeof(float2))));
    if (zeroMask && (s->imgN != N || s->imgBatch != chunk)) {
        if (s->planImgC2R) { cufftDestroy(s->planImgC2R); s->planImgC2R = 0; }
        if (s->planImgR2C) { cufftDestroy(s->planImgR2C); s->planImgR2C = 0; }
        s->imgN = 0;
        int dims[2] = {N, N};
        THB_FFT(ctx, cufftPlanMany(&s->planImgC2R, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, chunk));
        THB_FFT(ctx, cufftPlanMany(&s->planImgR2C, 2, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, chunk));
        THB_FFT(ctx, cufftSetStream(s->planImgC2R, ctx->stream));
        THB_FFT(ctx, cufftSetStream(s->planImgR2C, ctx->stream));
        s->imgN = N;
        s->imgBatch = chunk;
    }
    int* dIdx = (int*)scratch(ctx, 0, sizeof(int) * (2 * (size_t)P + nImg) + sizeof(float) * ((size_t)nGroup * nRing + 9 * (size_t)nImg));
    float2* dImg = (float2*)scratch(ctx, 4, (size_t)chunk * imgElems * sizeof(float2));
    float* dRl = (float*)scratch(ctx, 5, (size_t)chunk * N * N * sizeof(float));
    if (!dIdx || !dImg || !dRl) return THB_E_CUDA;
    int* dPxl = dIdx; int* dSig = dPxl + P; int* dGrp = dSig + P;
    float* dTab = (float*)(dGrp + nImg); float* dAttr = dTab + (size_t)nGroup * nRing; float* dOff = dAttr + 7 * (size_t)nImg;
    std::vector<float> rcr(2 * (size_t)nImg);
    for (int l = 0; l < nImg; ++l) {   // RFLOAT rCol = nTransCol / nColRL
        rcr[2 * l] = (float)offset[2 * l] / (float)N;
        rcr[2 * l + 1] = (float)offset[2 * l + 1] / (float)N;
    }
    THB_CUDA(ctx, cudaMemcpyAsync(dPxl, iPxl, sizeof(int) * P, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dSig, iSig, sizeof(int) * P, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dTab, sigRcpTab, sizeof(float) * (size_t)nGroup * nRing, cudaMemcpyHostToDevice, ctx->stream));
    if (groupOfImg) THB_CUDA(ctx, cudaMemcpyAsync(dGrp, groupOfImg, sizeof(int) * nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dAttr, ctfAttr, sizeof(float) * 7 * (size_t)nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dOff, rcr.data(), sizeof(float) * 2 * (size_t)nImg, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = ctx->smCount * 8;
    for (int i0 = 0; i0 < nImg; i0 += chunk) {
        const int c = std::min(chunk, nImg - i0);
        THB_CUDA(ctx, cudaMemcpyAsync(dImg, imgOriFT + 2 * (size_t)i0 * imgElems, (size_t)c * imgElems * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
        span_begin(ctx, KF_PACK);
        img_translate_kernel<<<grid, 256, 0, ctx->stream>>>(dImg, N, c, dOff + 2 * i0);
        ctx->launches++;
        if (zeroMask) {
            if (c < chunk) THB_CUDA(ctx, cudaMemsetAsync(dImg + (size_t)c * imgElems, 0, (size_t)(chunk - c) * imgElems * sizeof(float2), ctx->stream));
            img_hermitian_cols_kernel<<<grid, 256, 0, ctx->stream>>>(dImg, N, c);
            THB_FFT(ctx, cufftExecC2R(s->planImgC2R, reinterpret_cast<cufftComplex*>(dImg), dRl));
            img_mask_kernel<<<grid, 256, 0, ctx->stream>>>(dRl, N, c, maskRadiusPx, 6.0f /* EDGE_WIDTH_RL, include/Macro.h:99 */);
            THB_FFT(ctx, cufftExecR2C(s->planImgR2C, dRl, reinterpret_cast<cufftComplex*>(dImg)));
            ctx->launches += 2;
        }
        if (imgOutFT)
            THB_CUDA(ctx, cudaMemcpyAsync(imgOutFT + 2 * (size_t)i0 * imgElems, dImg, (size_t)c * imgElems * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
        const size_t doff = (size_t)(base + i0) * P;
        dim3 pgrid(std::min((P + 255) / 256, 64), c);
        pack_stack_kernel<<<pgrid, 256, 0, ctx->stream>>>(dImg, imgElems, ctx->pixE, ctx->permE, dPxl, dSig, P, dTab, nRing, groupOfImg ? dGrp + i0 : nullptr,
                                                          reinterpret_cast<const CtfAttr7*>(dAttr) + i0, pixelSize, N, st.dat + doff, st.ctf + doff, st.sig + doff);
        span_end(ctx);
        ctx->launches++;
        THB_CUDA(ctx, cudaGetLastError());
    }
    if (slotOfImg)
        THB_CUDA(ctx, cudaMemcpyAsync(st.slot + base, slotOfImg, (size_t)nImg * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    else
        THB_CUDA(ctx, cudaMemsetAsync(st.slot + base, 0, (size_t)nImg * sizeof(int), ctx->stream));
    if (st.hslot.size() >= (size_t)base + nImg)          // host copy of the slots (slot validation, launch order of the lockstep E kernel)
        for (int i = 0; i < nImg; ++i) st.hslot[(size_t)base + i] = slotOfImg ? slotOfImg[i] : 0;
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

// section 8(f) row 3: the image loop of Optimiser::allReduceSigma on the resident stacks
int thb_sigma_accumulate(thb_ctx* ctx, int nImg, const int* imgIdx, const double* quat, const double* tran, const double* offS,
                         const int* groupOfImg, int nGroup, int rSig, const int* iSigE, const int* iSigM, double* sigM, double* sigN,
                         double* svd)
{
    if (!ctx) return THB_E_ARG;
    if (!ctx->pixE || !ctx->pixM || !ctx->stackE.dat || !ctx->stackM.dat) return set_error(ctx, THB_E_STATE, "sigma_accumulate: pixel lists / stacks missing");
    if (nImg <= 0 || !quat || !tran || !iSigE || !iSigM || !sigM || !sigN || !svd || nGroup <= 0 || rSig <= 0)
        return set_error(ctx, THB_E_ARG, "sigma_accumulate: bad arguments");
    if (!imgIdx && (nImg > ctx->stackE.nImg || nImg > ctx->stackM.nImg)) return set_error(ctx, THB_E_ARG, "sigma_accumulate: nImg exceeds the stacks");
    int vdim = 0;
    for (int i = 0; i < THB_MAX_SLOTS; ++i)
        if (ctx->vols[i].d) vdim = ctx->vols[i].vdim;
    if (!vdim) return set_error(ctx, THB_E_STATE, "sigma_accumulate: no projector volume");
    for (int l = 0; l < nImg; ++l) {
        if (imgIdx && (imgIdx[l] < 0 || imgIdx[l] >= ctx->stackE.nImg || imgIdx[l] >= ctx->stackM.nImg)) return set_error(ctx, THB_E_ARG, "sigma_accumulate: imgIdx[%d] outside the stacks", l);
        if (groupOfImg && (groupOfImg[l] < 0 || groupOfImg[l] >= nGroup)) return set_error(ctx, THB_E_ARG, "sigma_accumulate: groupOfImg[%d] out of range", l);
    }
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int PE = ctx->nPxlE, PM = ctx->nPxlM, pfM = ctx->pfM;
    // rings in the device (blocked) pixel order; ring populations of the sigma pixel set {|k|^2 < rSig^2, rint|k| < rSig}
    std::vector<int> permE(PE), permM(PM), ringE(PE), ringM(PM);
    std::vector<int4> pixM(PM);
    THB_CUDA(ctx, cudaMemcpy(permE.data(), ctx->permE, sizeof(int) * PE, cudaMemcpyDeviceToHost));
    THB_CUDA(ctx, cudaMemcpy(permM.data(), ctx->permM, sizeof(int) * PM, cudaMemcpyDeviceToHost));
    THB_CUDA(ctx, cudaMemcpy(pixM.data(), ctx->pixM, sizeof(int4) * PM, cudaMemcpyDeviceToHost));
    // powerSpectrum() walks the whole half plane, the packed lists leave out (i = 0, j < 0) (allocPreCalIdx,
    // src/Optimiser.cpp:8015): the mirror pixel (0, j > 0) has the same modulus for the Hermitian images on this path and
    // is counted twice (SIGMA_DUP flag)
    std::vector<int4> pixE(PE);
    THB_CUDA(ctx, cudaMemcpy(pixE.data(), ctx->pixE, sizeof(int4) * PE, cudaMemcpyDeviceToHost));
    std::vector<float> cntE(rSig, 0.f), cntM(rSig, 0.f);
    auto ring_of = [&](const int4& c, int u, std::vector<float>& cnt) {
        const long long a = c.z, b = c.w;                  // unpadded iCol, iRow
        if (!(a * a + b * b < (long long)rSig * rSig && u >= 0 && u < rSig)) return -1;
        const bool dup = a == 0 && b > 0;
        cnt[u] += dup ? 2.f : 1.f;
        return dup ? (u | SIGMA_DUP) : u;
    };
    for (int i = 0; i < PE; ++i) ringE[i] = ring_of(pixE[i], iSigE[permE[i]], cntE);
    for (int i = 0; i < PM; ++i) ringM[i] = ring_of(pixM[i], iSigM[permM[i]], cntM);
    (void)pfM;
    const size_t nOut = (size_t)nGroup * (rSig + 1);
    unsigned char* buf = (unsigned char*)scratch(ctx, 0, sizeof(double) * (3 * nOut + 8 * (size_t)nImg) + sizeof(int) * ((size_t)PE + PM + 2 * (size_t)nImg) + sizeof(float) * 2 * (size_t)rSig + 64);
    if (!buf) return THB_E_CUDA;
    double* dOut = (double*)buf;
    double* dQ = dOut + 3 * nOut; double* dT = dQ + 4 * (size_t)nImg; double* dO = dT + 2 * (size_t)nImg;
    int* dRingE = (int*)(dO + 2 * (size_t)nImg); int* dRingM = dRingE + PE; int* dIdx = dRingM + PM; int* dGrp = dIdx + nImg;
    float* dCntE = (float*)(dGrp + nImg); float* dCntM = dCntE + rSig;
    std::vector<double> zeroOff(2 * (size_t)nImg, 0.0);
    THB_CUDA(ctx, cudaMemsetAsync(dOut, 0, sizeof(double) * 3 * nOut, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dQ, quat, sizeof(double) * (ctx->mode2D ? 2 : 4) * nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dT, tran, sizeof(double) * 2 * nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dO, offS ? offS : zeroOff.data(), sizeof(double) * 2 * nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dRingE, ringE.data(), sizeof(int) * PE, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dRingM, ringM.data(), sizeof(int) * PM, cudaMemcpyHostToDevice, ctx->stream));
    if (imgIdx) THB_CUDA(ctx, cudaMemcpyAsync(dIdx, imgIdx, sizeof(int) * nImg, cudaMemcpyHostToDevice, ctx->stream));
    if (groupOfImg) THB_CUDA(ctx, cudaMemcpyAsync(dGrp, groupOfImg, sizeof(int) * nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dCntE, cntE.data(), sizeof(float) * rSig, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dCntM, cntM.data(), sizeof(float) * rSig, cudaMemcpyHostToDevice, ctx->stream));
    SigmaArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < THB_MAX_SLOTS; ++i) a.vols[i] = ctx->vols[i].d;
    a.vdim = vdim; a.pitch = (vdim / 2 + 2 + 3) & ~3; a.N = ctx->N; a.rSig = rSig; a.nGroup = nGroup;
    a.datE = ctx->stackE.dat; a.ctfE = ctx->stackE.ctf; a.slotE = ctx->stackE.slot; a.pixE = ctx->pixE; a.ringE = dRingE; a.PE = PE;
    a.datM = ctx->stackM.dat; a.ctfM = ctx->stackM.ctf; a.pixM = ctx->pixM; a.ringM = dRingM; a.PM = PM;
    a.imgIdx = imgIdx ? dIdx : nullptr; a.quat = dQ; a.tran = dT; a.offS = dO; a.group = groupOfImg ? dGrp : nullptr;
    a.mode2D = ctx->mode2D;
    a.cntE = dCntE; a.cntM = dCntM; a.sigM = dOut; a.sigN = dOut + nOut; a.svd = dOut + 2 * nOut;
    span_begin(ctx, KF_PACK);
    sigma_kernel<<<nImg, 256, sizeof(float) * 4 * rSig, ctx->stream>>>(a);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    THB_CUDA(ctx, cudaMemcpyAsync(sigM, dOut, sizeof(double) * nOut, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(sigN, dOut + nOut, sizeof(double) * nOut, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(svd, dOut + 2 * nOut, sizeof(double) * nOut, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_symmetrize(thb_ctx* ctx, int slot, int nElem, const double* R, double radius)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || !ctx->accs[slot].d) return set_error(ctx, THB_E_STATE, "symmetrize: slot %d not allocated", slot);
    if (ctx->mode2D) return set_error(ctx, THB_E_STATE, "symmetrize: MODE_3D only");
    if (nElem < 0 || (nElem > 0 && !R) || !(radius > 0)) return set_error(ctx, THB_E_ARG, "symmetrize: bad arguments");
    if (nElem == 0) return THB_OK;                                   // C1
    Accum& a = ctx->accs[slot];
    if (radius > a.vdim / 2 - 1) return set_error(ctx, THB_E_ARG, "symmetrize: radius %g reaches outside the accumulator (%d)", radius, a.vdim);
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    float4* out = nullptr;
    THB_CUDA(ctx, cudaMalloc(&out, a.nVox * sizeof(float4)));
    struct Guard { float4* p; ~Guard() { cudaFree(p); } } guard{out};     // the new buffer on an error return, the old one on success
    double* dR = (double*)scratch(ctx, 0, sizeof(double) * 9 * (size_t)nElem);
    if (!dR) return THB_E_CUDA;
    THB_CUDA(ctx, cudaMemcpyAsync(dR, R, sizeof(double) * 9 * (size_t)nElem, cudaMemcpyHostToDevice, ctx->stream));
    span_begin(ctx, KF_PACK);
    symmetrize_kernel<<<ctx->smCount * 16, 256, 0, ctx->stream>>>(a.d, out, a.vdim, nElem, dR, radius * radius);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    // symmetrizeO (src/Reconstructor.cpp:2692-2716): O += sum_e R_e O, counter *= 1 + nElem
    double O[3];
    int cnt = 0;
    THB_CUDA(ctx, cudaMemcpyAsync(O, ctx->dO + 3 * slot, sizeof(O), cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(&cnt, ctx->dCounter + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double res[3] = {O[0], O[1], O[2]};
    for (int e = 0; e < nElem; ++e) {
        const double* m = R + 9 * e;
        for (int k = 0; k < 3; ++k) res[k] += m[k] * O[0] + m[3 + k] * O[1] + m[6 + k] * O[2];
    }
    cnt *= 1 + nElem;
    THB_CUDA(ctx, cudaMemcpyAsync(ctx->dO + 3 * slot, res, sizeof(res), cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(ctx->dCounter + slot, &cnt, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    guard.p = a.d;
    a.d = out;
    return THB_OK;
}

int thb_norm_residual(thb_ctx* ctx, int nImg, const int* imgIdx, const double* quat, const double* tran, float rL, float rNorm,
                      double* norm)
{
    if (!ctx) return THB_E_ARG;
    if (!ctx->pixE || !ctx->stackE.dat) return set_error(ctx, THB_E_STATE, "norm_residual: E pixel list / stack missing");
    if (nImg <= 0 || !quat || !tran || !norm) return set_error(ctx, THB_E_ARG, "norm_residual: bad arguments");
    if (!imgIdx && nImg > ctx->stackE.nImg) return set_error(ctx, THB_E_ARG, "norm_residual: nImg exceeds the stack");
    if (imgIdx)
        for (int l = 0; l < nImg; ++l)
            if (imgIdx[l] < 0 || imgIdx[l] >= ctx->stackE.nImg) return set_error(ctx, THB_E_ARG, "norm_residual: imgIdx[%d] outside the stack", l);
    int vdim = 0;
    for (int i = 0; i < THB_MAX_SLOTS; ++i)
        if (ctx->vols[i].d) vdim = ctx->vols[i].vdim;
    if (!vdim) return set_error(ctx, THB_E_STATE, "norm_residual: no projector volume");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int qc = ctx->mode2D ? 2 : 4;
    double* buf = (double*)scratch(ctx, 0, sizeof(double) * (size_t)nImg * (qc + 3) + sizeof(int) * (size_t)nImg);
    if (!buf) return THB_E_CUDA;
    double* dQ = buf; double* dT = dQ + (size_t)qc * nImg; double* dN = dT + 2 * (size_t)nImg;
    int* dIdx = (int*)(dN + nImg);
    THB_CUDA(ctx, cudaMemcpyAsync(dQ, quat, sizeof(double) * qc * nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dT, tran, sizeof(double) * 2 * nImg, cudaMemcpyHostToDevice, ctx->stream));
    if (imgIdx) THB_CUDA(ctx, cudaMemcpyAsync(dIdx, imgIdx, sizeof(int) * nImg, cudaMemcpyHostToDevice, ctx->stream));
    SigmaArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < THB_MAX_SLOTS; ++i) a.vols[i] = ctx->vols[i].d;
    a.vdim = vdim; a.pitch = (vdim / 2 + 2 + 3) & ~3; a.N = ctx->N; a.mode2D = ctx->mode2D;
    a.datE = ctx->stackE.dat; a.ctfE = ctx->stackE.ctf; a.slotE = ctx->stackE.slot; a.pixE = ctx->pixE; a.PE = ctx->nPxlE;
    a.imgIdx = imgIdx ? dIdx : nullptr; a.quat = dQ; a.tran = dT;
    span_begin(ctx, KF_PACK);
    norm_kernel<<<nImg, 256, 0, ctx->stream>>>(a, rL * rL, rNorm * rNorm, dN);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    THB_CUDA(ctx, cudaMemcpyAsync(norm, dN, sizeof(double) * nImg, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_scale_images(thb_ctx* ctx, int nImg, const int* imgIdx, const float* scale)
{
    if (!ctx) return THB_E_ARG;
    if (nImg <= 0 || !scale) return set_error(ctx, THB_E_ARG, "scale_images: bad arguments");
    if (!ctx->stackE.dat && !ctx->stackM.dat) return set_error(ctx, THB_E_STATE, "scale_images: no resident stack");
    const int cap = std::min(ctx->stackE.dat ? ctx->stackE.nImg : INT_MAX, ctx->stackM.dat ? ctx->stackM.nImg : INT_MAX);
    if (!imgIdx && nImg > cap) return set_error(ctx, THB_E_ARG, "scale_images: nImg exceeds the stacks");
    if (imgIdx)
        for (int l = 0; l < nImg; ++l)
            if (imgIdx[l] < 0 || imgIdx[l] >= cap) return set_error(ctx, THB_E_ARG, "scale_images: imgIdx[%d] outside the stacks", l);
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    float* dS = (float*)scratch(ctx, 0, (sizeof(float) + sizeof(int)) * (size_t)nImg);
    if (!dS) return THB_E_CUDA;
    int* dIdx = (int*)(dS + nImg);
    THB_CUDA(ctx, cudaMemcpyAsync(dS, scale, sizeof(float) * nImg, cudaMemcpyHostToDevice, ctx->stream));
    if (imgIdx) THB_CUDA(ctx, cudaMemcpyAsync(dIdx, imgIdx, sizeof(int) * nImg, cudaMemcpyHostToDevice, ctx->stream));
    const int P = std::max(ctx->stackE.dat ? ctx->nPxlE : 0, ctx->stackM.dat ? ctx->nPxlM : 0);
    dim3 grid(std::min((P + 255) / 256, 64), nImg);
    span_begin(ctx, KF_PACK);
    scale_images_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->stackE.dat, ctx->stackE.dat ? ctx->nPxlE : 0, ctx->stackM.dat,
                                                       ctx->stackM.dat ? ctx->nPxlM : 0, imgIdx ? dIdx : nullptr, dS);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_set_projectee(thb_ctx* ctx, int slot, const float* volReal, int N, int pf)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || N <= 0 || (N & 1) || pf <= 0) return set_error(ctx, THB_E_ARG, "set_projectee: bad arguments");
    // MODE_2D: Projector::setProjectee(Image) (src/Projector.cpp:97-121): volReal = the N x N class average, the result is
    // plane 0 of the two-plane reference
    const int rank = ctx->mode2D ? 2 : 3;
    const int Nz = rank == 3 ? N : 1;
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    RecoState* s = reco_state(ctx);
    const size_t nIn = (size_t)N * N * Nz;
    const float* dIn = nullptr;
    if (volReal) {
        float* tmp = (float*)scratch(ctx, 1, nIn * sizeof(float));
        if (!tmp) return THB_E_CUDA;
        THB_CUDA(ctx, cudaMemcpyAsync(tmp, volReal, nIn * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        dIn = tmp;
    } else {
        if (!s->dVol || s->volN != N || s->volRank != rank) return set_error(ctx, THB_E_STATE, "set_projectee: no reconstruction of edge %d on the device", N);
        dIn = s->dVol;
    }
    const int n = N * pf;
    const int pitch = (n / 2 + 2 + 3) & ~3;
    Volume3& v = ctx->vols[slot];
    if (v.quad) {
        THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(v.quad);
        v.quad = nullptr;
    }
    const int nz = rank == 3 ? n : 1;
    const size_t rows = rank == 3 ? (size_t)n * n : 2 * (size_t)n;          // MODE_2D: two planes, the second stays zero
    if (v.vdim != n) {
        cudaFree(v.d);
        v.d = nullptr;
        v.vdim = 0;
        THB_CUDA(ctx, cudaMalloc(&v.d, rows * pitch * sizeof(float2)));
        v.vdim = n;
        v.pitch = pitch;
    }
    THB_CUDA(ctx, cudaMemsetAsync(v.d, 0, rows * pitch * sizeof(float2), ctx->stream));
    float* R = (float*)scratch(ctx, 3, (size_t)n * n * nz * sizeof(float));
    if (!R) return THB_E_CUDA;
    THB_CUDA(ctx, cudaMemsetAsync(R, 0, (size_t)n * n * nz * sizeof(float), ctx->stream));
    span_begin(ctx, KF_PACK);
    proj_pad_rl_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(dIn, N, Nz, n, pf, R);
    ctx->launches++;
    if (s->projDim != n || s->projPitch != pitch || s->projRank != rank) {
        if (s->planProj) { cufftDestroy(s->planProj); s->planProj = 0; }
        s->projDim = 0;
        if (rank == 3) {
            int dims[3] = {n, n, n}, inembed[3] = {n, n, n}, onembed[3] = {n, n, pitch};
            THB_FFT(ctx, cufftPlanMany(&s->planProj, 3, dims, inembed, 1, n * n * n, onembed, 1, n * n * pitch, CUFFT_R2C, 1));
        } else {
            int dims[2] = {n, n}, inembed[2] = {n, n}, onembed[2] = {n, pitch};
            THB_FFT(ctx, cufftPlanMany(&s->planProj, 2, dims, inembed, 1, n * n, onembed, 1, n * pitch, CUFFT_R2C, 1));
        }
        THB_FFT(ctx, cufftSetStream(s->planProj, ctx->stream));
        s->projDim = n;
        s->projPitch = pitch;
        s->projRank = rank;
    }
    THB_FFT(ctx, cufftExecR2C(s->planProj, R, reinterpret_cast<cufftComplex*>(v.d)));
    span_end(ctx);
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

}  // extern "C"
