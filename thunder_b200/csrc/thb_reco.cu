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
#include <vector>
#include "thb_context.h"

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
__global__ void reco_hermitian_planes_kernel(float2* C, int m)
{
    const int nc = m / 2 + 1;
    const size_t plane = (size_t)m * m;
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
__global__ void reco_kernel_rl_kernel(float* c, int m, const float* tab, float tabStep, float nf, float M2, float invVol)
{
    const size_t total = (size_t)m * m * m;
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
__global__ void reco_extract_kernel(const float* padReal, int M, int N, int pf, float invVol, float* dst)
{
    const size_t total = (size_t)N * N * N;
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
__global__ void proj_pad_rl_kernel(const float* vol, int N, int n, int pf, float* pad)
{
    const size_t total = (size_t)N * N * N;
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
    int planDim = 0;
    cufftHandle planProj = 0;
    int projDim = 0, projPitch = 0;
    float* dTab = nullptr;
    double tabA = -1, tabAlpha = -1;
    float nf = 1.0f;
    float* dVol = nullptr;      // last reconstruction, N^3 real
    int volN = 0;
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
    cudaFree(s->dTab);
    cudaFree(s->dVol);
    delete s;
    ctx->reco = nullptr;
}

static int ensure_plans(thb_ctx* ctx, RecoState* s, int m)
{
    if (s->planDim == m) return THB_OK;
    if (s->planC2R) { cufftDestroy(s->planC2R); s->planC2R = 0; }
    if (s->planR2C) { cufftDestroy(s->planR2C); s->planR2C = 0; }
    s->planDim = 0;
    THB_FFT(ctx, cufftPlan3d(&s->planC2R, m, m, m, CUFFT_C2R));
    THB_FFT(ctx, cufftPlan3d(&s->planR2C, m, m, m, CUFFT_R2C));
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
    const Accum& a = ctx->accs[slot];
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
    const Accum& acc = ctx->accs[slot];
    const int m = acc.vdim;
    if (N <= 0 || (N & 1) || pf <= 0 || m % pf || m / pf > N) return set_error(ctx, THB_E_ARG, "reconstruct: accumulator dimension %d does not fit N = %d, pf = %d", m, N, pf);
    if (fsc && nFsc <= 0) return set_error(ctx, THB_E_ARG, "reconstruct: empty FSC");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    RecoState* s = reco_state(ctx);
    const int size = m / pf, M = N * pf;
    const int maxRadius = size / 2 - (int)std::ceil(a);       // Reconstructor::init, src/Reconstructor.cpp:88
    if (maxRadius <= 0) return set_error(ctx, THB_E_ARG, "reconstruct: volume too small for the kernel radius");
    const size_t nVox = acc.nVox, nReal = (size_t)m * m * m;
    const size_t nVoxM = (size_t)(M / 2 + 1) * M * M, nRealM = (size_t)M * M * M;
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
        if ((rc = ensure_plans(ctx, s, m))) return rc;
        float diffC = 3.40282347e38f, diffPrev;
        int noDecrease = 0;
        for (int it = 0; it < 30; ++it) {                     // MAX_N_ITER_BALANCE
            reco_make_c_kernel<<<grid, 256, 0, ctx->stream>>>(acc.d, W, C, nVox);
            reco_hermitian_planes_kernel<<<grid, 256, 0, ctx->stream>>>(C, m);
            THB_FFT(ctx, cufftExecC2R(s->planC2R, reinterpret_cast<cufftComplex*>(C), R));
            reco_kernel_rl_kernel<<<grid, 256, 0, ctx->stream>>>(R, m, s->dTab, 1.0f / (float)RECO_TAB_N, s->nf, (float)M * (float)M,
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
    if ((rc = ensure_plans(ctx, s, M))) return rc;
    THB_CUDA(ctx, cudaMemsetAsync(C, 0, nVoxM * sizeof(float2), ctx->stream));
    reco_pad_ft_kernel<<<grid, 256, 0, ctx->stream>>>(acc.d, W, nVox, m, M, pf, maxRadius, C);
    reco_hermitian_planes_kernel<<<grid, 256, 0, ctx->stream>>>(C, M);
    THB_FFT(ctx, cufftExecC2R(s->planC2R, reinterpret_cast<cufftComplex*>(C), R));
    if (s->volN != N) {
        cudaFree(s->dVol);
        s->dVol = nullptr;
        s->volN = 0;
        THB_CUDA(ctx, cudaMalloc(&s->dVol, sizeof(float) * (size_t)N * N * N));
        s->volN = N;
    }
    reco_extract_kernel<<<grid, 256, 0, ctx->stream>>>(R, M, N, pf, (float)(1.0 / (double)nRealM), s->dVol);
    ctx->launches += 3;
    span_end(ctx);
    THB_CUDA(ctx, cudaGetLastError());
    if (dstReal) THB_CUDA(ctx, cudaMemcpyAsync(dstReal, s->dVol, sizeof(float) * (size_t)N * N * N, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (nIterOut) *nIterOut = nIter;
    return THB_OK;
}

int thb_set_projectee(thb_ctx* ctx, int slot, const float* volReal, int N, int pf)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || N <= 0 || (N & 1) || pf <= 0) return set_error(ctx, THB_E_ARG, "set_projectee: bad arguments");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    RecoState* s = reco_state(ctx);
    const size_t nIn = (size_t)N * N * N;
    const float* dIn = nullptr;
    if (volReal) {
        float* tmp = (float*)scratch(ctx, 1, nIn * sizeof(float));
        if (!tmp) return THB_E_CUDA;
        THB_CUDA(ctx, cudaMemcpyAsync(tmp, volReal, nIn * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        dIn = tmp;
    } else {
        if (!s->dVol || s->volN != N) return set_error(ctx, THB_E_STATE, "set_projectee: no reconstruction of edge %d on the device", N);
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
    const size_t rows = (size_t)n * n;
    if (v.vdim != n) {
        cudaFree(v.d);
        v.d = nullptr;
        v.vdim = 0;
        THB_CUDA(ctx, cudaMalloc(&v.d, rows * pitch * sizeof(float2)));
        v.vdim = n;
        v.pitch = pitch;
    }
    THB_CUDA(ctx, cudaMemsetAsync(v.d, 0, rows * pitch * sizeof(float2), ctx->stream));
    float* R = (float*)scratch(ctx, 3, (size_t)n * n * n * sizeof(float));
    if (!R) return THB_E_CUDA;
    THB_CUDA(ctx, cudaMemsetAsync(R, 0, (size_t)n * n * n * sizeof(float), ctx->stream));
    span_begin(ctx, KF_PACK);
    proj_pad_rl_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(dIn, N, n, pf, R);
    ctx->launches++;
    if (s->projDim != n || s->projPitch != pitch) {
        if (s->planProj) { cufftDestroy(s->planProj); s->planProj = 0; }
        s->projDim = 0;
        int dims[3] = {n, n, n}, inembed[3] = {n, n, n}, onembed[3] = {n, n, pitch};
        THB_FFT(ctx, cufftPlanMany(&s->planProj, 3, dims, inembed, 1, n * n * n, onembed, 1, n * n * pitch, CUFFT_R2C, 1));
        THB_FFT(ctx, cufftSetStream(s->planProj, ctx->stream));
        s->projDim = n;
        s->projPitch = pitch;
    }
    THB_FFT(ctx, cufftExecR2C(s->planProj, R, reinterpret_cast<cufftComplex*>(v.d)));
    span_end(ctx);
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

}  // extern "C"
