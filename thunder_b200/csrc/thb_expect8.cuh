// thb_expect8.cuh - global scan with SHARED TEMPLATES (src/Optimiser.cpp:756-914, logDataVSPrior_m_n_huabin :9931-9973).
//
// In the scan every image is compared with the SAME rotation set: the reference projects each rotation once
// (Projector::project, :770-786; translations and logDataVSPrior_m_n over all images, :789-830) and evaluates all images against that slice.  The fused local-search kernel, used for the scan
// until now, re-gathers every shared rotation for every image; ncu showed the scan bound by instruction issue, 60 % of the
// instructions being the coordinate / weight / gather arithmetic of those repeated projections.  Here:
//
//   scan_project_kernel    one slice per rotation, ONCE per launch, into a template table  tmpl[pixel][rotation]  (rotation
//                          fastest, so that the lanes of a warp - one rotation each - read consecutive addresses)
//   scan_contract_kernel   one CTA per image: the expanded likelihood
//                              logL(r,t) = sum_i sig_i |dat_i|^2 + sum_i ( u_ti . pri_ri + g_i |pri_ri|^2 )
//                          as a register-tiled contraction over the pixels - 4 rotations per lane x TC translations per pass,
//                          2 FMAs per (rotation, translation, pixel), the pixel records (g_i, u_ti) broadcast from shared
//                          memory - no coordinates, no weights, no gather: the fp32 FMA pipe is the roofline of this kernel
//   scan_epilogue_kernel   baseline and marginal weights from the [image][nR][nT] table (src/Optimiser.cpp:834-894)
//
// Same values as the fused kernel: the slice arithmetic is gather_ft (identical operation order), the records are the same
// expressions; only the order of the sum over pixels differs (partial sums per pixel part, combined in double).
#pragma once
#include <cuda_runtime.h>
#include "thb_kernels.cuh"
#include "thb_expect3.cuh"

namespace thb {

constexpr int E8_THREADS = 256;
constexpr int E8_RPL = 4;                 // rotations per lane
constexpr int E8_WROT = 32 * E8_RPL;      // rotations per warp
constexpr int E8_TILE = 128;

// one record per pixel; its size in 16-byte words is ODD (TC = 15: 9 words, 144 bytes; TC = 9: 5 words, 80 bytes) so that the records of
// consecutive pixels start in different shared-memory banks: with 128-byte records every store of the record build was a
// 16-way bank conflict (ncu: 45 % of all shared-memory wavefronts of the kernel)
template <int TC>
struct __align__(16) ScanRec {
    float g, pad;       // sig * ctf^2
    float2 u[TC];       // -2 sig ctf dat conj(tra_t)
    float4 skew[((8 + 8 * TC) / 16) % 2 == 0 ? 1 : 0];
};
static_assert(sizeof(ScanRec<15>) == 144 && sizeof(ScanRec<9>) == 80, "record sizes");

// templates of rotations [r0, r0 + nRc) of one reference: tmpl[i * nRpad + colBase + (r - r0)], pixel i in the resident (permuted) order
__global__ void scan_project_kernel(const float2* __restrict__ vol, int n, int pitch, const int4* __restrict__ pix, int P,
                                    View3 quat, int r0, int nRc, int nRpad, int mode2D, float2* __restrict__ tmpl, int colBase = 0)
{
    const size_t total = (size_t)P * nRc;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / nRc), rr = (int)(idx % nRc);
        double q[4] = {1.0, 0.0, 0.0, 0.0};
        for (int c = 0; c < (mode2D ? 2 : 4); ++c) q[c] = quat.at(0, r0 + rr, c);
        const Rot2 rot = make_rot2(q, mode2D);
        const int4 px = pix[i];
        float x, y, z;
        slice_coord(rot, (double)px.x, (double)px.y, x, y, z);
        tmpl[(size_t)i * nRpad + colBase + rr] = gather_ft(vol, n, pitch, x, y, z);
    }
}

template <int TC>
constexpr size_t e8_smem_bytes()
{
    // record tile; the parking area of the partial sums ((parts - 1) x 1024 rotation slots x (TC + 1) floats at most
    // 7 x 128 x (TC + 1)) reuses it
    const size_t tile = E8_TILE * sizeof(ScanRec<TC>);
    const size_t park = (size_t)7 * E8_WROT * (TC + 1) * sizeof(float);
    return tile > park ? tile : park;
}

// table[p][r][t] for r in [r0, r0 + nRc): grid = images of the launch, 256 threads
template <int TC>
__global__ void __launch_bounds__(E8_THREADS, 2) scan_contract_kernel(const ExpectArgs A, const float2* __restrict__ tmpl, int r0,
                                                                    int nRc, int nRpad, float* __restrict__ table)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ScanRec<TC>* tile = reinterpret_cast<ScanRec<TC>*>(smem_raw);
    __shared__ float sRC[TC], sRR[TC];
    __shared__ double redd[E8_THREADS / 32];
    __shared__ double sK0;

    const int p = blockIdx.x;
    if (A.active && !A.active[p]) return;
    const int img = A.imgIdx ? A.imgIdx[p] : p + A.imgBase;
    const int P = A.P;
    const float2* __restrict__ dat = A.dat + (size_t)img * P;
    const float* __restrict__ ctf = A.ctf + (size_t)img * P;
    const float* __restrict__ sig = A.sig + (size_t)img * P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* __restrict__ tab = table + (size_t)p * A.nR * A.nT;
    // the 8 warps: G rotation groups of 128 x 8 / G pixel parts
    const int groupsNeeded = (nRc + E8_WROT - 1) / E8_WROT;
    const int G = groupsNeeded >= 8 ? 8 : groupsNeeded >= 4 ? 4 : groupsNeeded >= 2 ? 2 : 1;
    const int nParts = (E8_THREADS / 32) / G;
    const int g = warp % G, ph = warp / G;
    double k0sum = 0.0;
    bool first = true;

    for (int rbase = 0; rbase < nRc; rbase += E8_WROT * G) {
        const int rw = rbase + g * E8_WROT;              // first rotation (within the chunk) of this warp
        for (int tbase = 0; tbase < A.nT; tbase += TC) {
            __syncthreads();
            if (tid < TC) {
                const int t = tbase + tid;
                float tx = 0.0f, ty = 0.0f;
                if (t < A.nT) {
                    tx = (float)A.tran.at(p, t, 0);
                    ty = (float)A.tran.at(p, t, 1);
                }
                sRC[tid] = tx / (float)A.N;
                sRR[tid] = ty / (float)A.N;
            }
            float acc[E8_RPL][TC], nrm[E8_RPL];
#pragma unroll
            for (int j = 0; j < E8_RPL; ++j) {
                nrm[j] = 0.0f;
#pragma unroll
                for (int t = 0; t < TC; ++t) acc[j][t] = 0.0f;
            }
            for (int tile0 = 0; tile0 < P; tile0 += E8_TILE) {
                const int cnt = min(E8_TILE, P - tile0);
                __syncthreads();
                {
                    const int k = tid >> 1, sub = tid & 1;
                    if (k < cnt) {
                        const int i = tile0 + k;
                        const int4 c = A.pix[i];
                        const float2 d = dat[i];
                        const float cf = ctf[i];
                        const float sg = sig[i];
                        const float m2 = -2.0f * sg * cf;
                        ScanRec<TC>& rec = tile[k];
                        if (sub == 0) {
                            rec.g = sg * cf * cf;
                            rec.pad = 0.0f;
                            if (first) k0sum += (double)(sg * (d.x * d.x + d.y * d.y));
                        }
#pragma unroll
                        for (int t = 0; t < TC; ++t) {
                            if ((t & 1) != sub) continue;
                            const float phs = translate_phase(c.z, c.w, sRC[t], sRR[t]);
                            float s, co;
                            sincosf(phs, &s, &co);
                            rec.u[t] = make_float2(m2 * (d.x * co - d.y * s), m2 * (d.x * s + d.y * co));
                        }
                    }
                }
                __syncthreads();
                if (rw < nRc) {
                    const float2* __restrict__ trow = tmpl + (size_t)tile0 * nRpad + rw + lane;
                    // the template values of the NEXT pixel are in flight while this one is contracted (the loads are L2 hits:
                    // without the prefetch the kernel waits for them every iteration)
                    float2 pv[E8_RPL], pn[E8_RPL];
                    if (ph < cnt) {
#pragma unroll
                        for (int j = 0; j < E8_RPL; ++j) pv[j] = __ldg(trow + (size_t)ph * nRpad + 32 * j);     // padded: always inside the table
                    }
#pragma unroll 1
                    for (int k = ph; k < cnt; k += nParts) {
                        const int kn = min(k + nParts, cnt - 1);
                        const float2* tp = trow + (size_t)kn * nRpad;
#pragma unroll
                        for (int j = 0; j < E8_RPL; ++j) pn[j] = __ldg(tp + 32 * j);
                        const ScanRec<TC>& rec = tile[k];
                        const float gk = rec.g;
#pragma unroll
                        for (int j = 0; j < E8_RPL; ++j) nrm[j] = fmaf(gk, fmaf(pv[j].x, pv[j].x, pv[j].y * pv[j].y), nrm[j]);
#pragma unroll
                        for (int t = 0; t < TC; ++t) {
                            const float2 u = rec.u[t];
#pragma unroll
                            for (int j = 0; j < E8_RPL; ++j) acc[j][t] = fmaf(u.x, pv[j].x, fmaf(u.y, pv[j].y, acc[j][t]));
                        }
#pragma unroll
                        for (int j = 0; j < E8_RPL; ++j) pv[j] = pn[j];
                    }
                }
            }
            // ---- end of the pass: constant term (first pass), partial sums of the pixel parts
            __syncthreads();
            if (first) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) k0sum += __shfl_xor_sync(0xffffffffu, k0sum, o);
                if (lane == 0) redd[warp] = k0sum;
                __syncthreads();
                if (tid == 0) {
                    double s = 0.0;
                    for (int w2 = 0; w2 < E8_THREADS / 32; ++w2) s += redd[w2];
                    sK0 = s;
                }
                first = false;
                __syncthreads();
            }
            const double k0 = sK0;
            float* park = reinterpret_cast<float*>(smem_raw);      // [part - 1][G * 128 rotation slots][TC + 1]
            if (ph > 0) {
#pragma unroll
                for (int j = 0; j < E8_RPL; ++j) {
                    float* pk = park + ((size_t)(ph - 1) * (G * E8_WROT) + g * E8_WROT + j * 32 + lane) * (TC + 1);
#pragma unroll
                    for (int t = 0; t < TC; ++t) pk[t] = acc[j][t];
                    pk[TC] = nrm[j];
                }
            }
            __syncthreads();
            if (ph == 0) {
                const size_t pstride = (size_t)(G * E8_WROT) * (TC + 1);
#pragma unroll
                for (int j = 0; j < E8_RPL; ++j) {
                    const int rc = rw + j * 32 + lane;                // rotation within the chunk
                    if (rc >= nRc) continue;
                    const float* pk0 = park + (size_t)(g * E8_WROT + j * 32 + lane) * (TC + 1);
                    double nn = (double)nrm[j];
                    for (int q = 1; q < nParts; ++q) nn += (double)pk0[(q - 1) * pstride + TC];
#pragma unroll
                    for (int t = 0; t < TC; ++t) {
                        if (tbase + t >= A.nT) continue;
                        double tot = (double)acc[j][t];
                        for (int q = 1; q < nParts; ++q) tot += (double)pk0[(q - 1) * pstride + t];
                        tab[(size_t)(r0 + rc) * A.nT + tbase + t] = (float)(k0 + nn + tot);
                    }
                }
            }
        }
    }
}

// baseline, marginal weights and the optional copy of the table: grid = images, 256 threads
__global__ void __launch_bounds__(256) scan_epilogue_kernel(const ExpectArgs A, float* __restrict__ table)
{
    __shared__ float redf[8];
    __shared__ double redd[8];
    const int p = blockIdx.x;
    if (A.active && !A.active[p]) return;
    expect_epilogue<256>(A, p, table + (size_t)p * A.nR * A.nT, redf, redd);
}

// MODE_2D classification scan of ALL classes in one table [image][nK * nR][nT] (template index f = class * nR + rotation: the
// contraction kernel does not know about classes, 20 classes x 100 rotations fill 2 000 of 2 048 rotation slots instead of 100 of
// 128 twenty times, and the pixel records are built once for all classes).  Epilogue as ExpectGlobal2D returns it
// (gpu/interface/Interface.h:176-198, src/Optimiser.cpp:834-894): ONE baseline per image across the classes,
//   wC[img][k] = sum_rt w pR pT,  wR[k][img][r] = sum_t w pT,  wT[k][img][t] = sum_r w pR,   w = exp(logL - baseline)
__global__ void __launch_bounds__(256) scan_classes_epilogue_kernel(const float* __restrict__ table, int nAct, int nK, int nR, int nT,
                                                                  const double* __restrict__ pR, const double* __restrict__ pT,
                                                                  float* __restrict__ wC, float* __restrict__ wR, float* __restrict__ wT,
                                                                  float* __restrict__ base)
{
    __shared__ float redf[8];
    __shared__ double redd[8];
    const int p = blockIdx.x, tid = threadIdx.x;
    const size_t nF = (size_t)nK * nR;
    const float* L = table + (size_t)p * nF * nT;
    float m = -INFINITY;
    for (size_t i = tid; i < nF * nT; i += 256) m = fmaxf(m, L[i]);
    m = block_reduce_max(m, redf);
    if (tid == 0) base[p] = m;
    for (int k = 0; k < nK; ++k) {
        const float* Lk = L + (size_t)k * nR * nT;
        double uc = 0.0;
        for (int r = tid; r < nR; r += 256) {
            float s = 0.0f;
            for (int t = 0; t < nT; ++t) s = (float)((double)s + (double)expf(Lk[r * nT + t] - m) * pT[t]);
            wR[((size_t)k * nAct + p) * nR + r] = s;
            uc += (double)s * pR[r];
        }
        for (int t = tid; t < nT; t += 256) {
            float s = 0.0f;
            for (int r = 0; r < nR; ++r) s = (float)((double)s + (double)expf(Lk[r * nT + t] - m) * pR[r]);
            wT[((size_t)k * nAct + p) * nT + t] = s;
        }
        uc = block_reduce_sum(uc, redd);
        if (tid == 0) wC[(size_t)p * nK + k] = (float)uc;
        __syncthreads();
    }
}

}  // namespace thb
