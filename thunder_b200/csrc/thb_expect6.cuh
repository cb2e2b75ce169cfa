// thb_expect6.cuh - the fused E kernel spread over the whole chip for a HANDFUL of images.
//
// expect_direct_kernel (thb_expect3.cuh) gives every image one CTA: right for thousands of images per launch, but one
// image-phase is 3.1 M samples (125 rotations x 25 134 pixels) and would keep a single SM busy for 11 ms.  The reference's
// own accelerator seam is exactly that case: its local search drives ONE image at a time through ExpectLocalRTD /
// ExpectLocalPreI3D / ExpectLocalM under a per-device lock (src/Optimiser.cpp:2484-2560; gpu/interface/Interface.h:31-164),
// and the tail of an adaptive E-step has a few unfinished particles left.  Here the grid is
//     (pixel chunk, rotation group of 32, image)
// - 4 x 37 CTAs for one image of the benchmark shape - every CTA walks its tiles with the same records, gather and expanded
// likelihood as the direct kernel (8 warps = 8 pixel parts of one rotation group) and ADDS its partial sums, constant term
// included, to a double-precision table [image][rotation][translation]; a second, tiny kernel turns the table into the
// baseline and the marginal weights (src/Optimiser.cpp:1383-1402).  Same arithmetic per sample as the direct kernel; only the
// order of the sum over pixels differs (double atomics: the table is exact to 1e-12).
#pragma once
#include <cuda_runtime.h>
#include "thb_expect3.cuh"

namespace thb {

constexpr int E6_THREADS = 256;

template <bool OCT, bool M2D>
__global__ void __launch_bounds__(E6_THREADS, 2) expect_spread_kernel(const ExpectArgs A, double* __restrict__ table, int nChunk)
{
    constexpr int TC = E_TC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PixelRecT<TC>* tile = reinterpret_cast<PixelRecT<TC>*>(smem_raw);
    __shared__ float sRC[TC], sRR[TC];
    __shared__ double redd[E6_THREADS / 32];

    const int pos = blockIdx.z;                             // launch position -> particle (A.order: compacted list of active particles)
    const int p = A.order ? A.order[pos] : pos;
    if (A.active && !A.active[p]) return;
    const int img = A.imgIdx ? A.imgIdx[p] : p + A.imgBase;
    const int slot = (M2D && A.slotAll >= 0) ? A.slotAll : (A.slotOfImg ? A.slotOfImg[img] : 0);
    const Quad* __restrict__ vol = reinterpret_cast<const Quad*>(A.quads.p[slot]);
    const int n = A.vdim;
    const int P = A.P;
    const float2* __restrict__ dat = A.dat + (size_t)img * P;
    const float* __restrict__ ctf = A.ctf + (size_t)img * P;
    const float* __restrict__ sig = A.sig + (size_t)img * P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int LB = A.quadBrick;
    const int rbase = blockIdx.y * 32;
    const int r = rbase + lane;                 // one rotation per lane, the 8 warps take every 8th pixel of a tile
    const bool rvalid = r < A.nR;
    constexpr int nParts = E6_THREADS / 32;
    const int ph = warp;
    double* __restrict__ tab = table + (size_t)pos * A.nR * A.nT;

    Rot2 rot;
    {
        double q[4] = {1.0, 0.0, 0.0, 0.0};
        if (rvalid)
            for (int c = 0; c < (M2D ? 2 : 4); ++c) q[c] = A.quat.at(p, r, c);
        rot = make_rot2(q, M2D);
    }
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
        float acc[TC];
#pragma unroll
        for (int t = 0; t < TC; ++t) acc[t] = 0.0f;
        float nrm = 0.0f;
        double k0 = 0.0;
        for (int tile0 = blockIdx.x * E3_TILE; tile0 < P; tile0 += nChunk * E3_TILE) {
            const int cnt = min(E3_TILE, P - tile0);
            __syncthreads();
            {
                const int k = tid >> 1, sub = tid & 1;
                if (k < cnt) {
                    const int i = tile0 + k;
                    const int4 c = A.pix[i];
                    const float2 d = dat[i];
                    const float cf = ctf[i], sg = sig[i];
                    const float m2 = -2.0f * sg * cf;
                    PixelRecT<TC>& rec = tile[k];
                    if (sub == 0) {
                        rec.a = (double)c.x;
                        rec.b = (double)c.y;
                        rec.g = sg * cf * cf;
                        rec.pad = 0.0f;
                        k0 += (double)(sg * (d.x * d.x + d.y * d.y));
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
            if (rvalid) {
#pragma unroll 2
                for (int k = ph; k < cnt; k += nParts) {
                    const PixelRecT<TC>& rec = tile[k];
                    float x, y, z;
                    slice_coord(rot, rec.a, rec.b, x, y, z);
                    int xb, yb, zb;
                    float xd, yd, zd;
                    const bool conj = fold_floor_fast(x, y, z, xb, yb, zb, xd, yd, zd);
                    const int x0 = xb - THB_FLOOR_BIAS, y0 = yb - THB_FLOOR_BIAS, z0 = zb - THB_FLOOR_BIAS;
                    const int ym = y0 < 0 ? y0 + n : y0;
                    const int zm = z0 < 0 ? z0 + n : z0;
                    const int zm1 = (z0 + 1 < 0) ? z0 + 1 + n : z0 + 1;
                    const Quad* q0 = OCT ? vol + 2 * quad_index(x0, ym, zm, n, LB) : vol + quad_index(x0, ym, zm, n, LB);
                    const Quad* q1 = OCT ? q0 + 1 : vol + quad_index(x0, ym, zm1, n, LB);
                    const Quad a = ldg_quad(q0), b = M2D ? Quad{} : ldg_quad(q1);
                    float w[8];
                    tri_weights(xd, yd, zd, w);
                    float re = a.v00.x * w[0], im = a.v00.y * w[0];
                    re = fmaf(a.v10.x, w[1], re); im = fmaf(a.v10.y, w[1], im);
                    re = fmaf(a.v01.x, w[2], re); im = fmaf(a.v01.y, w[2], im);
                    re = fmaf(a.v11.x, w[3], re); im = fmaf(a.v11.y, w[3], im);
                    re = fmaf(b.v00.x, w[4], re); im = fmaf(b.v00.y, w[4], im);
                    re = fmaf(b.v10.x, w[5], re); im = fmaf(b.v10.y, w[5], im);
                    re = fmaf(b.v01.x, w[6], re); im = fmaf(b.v01.y, w[6], im);
                    re = fmaf(b.v11.x, w[7], re); im = fmaf(b.v11.y, w[7], im);
                    if (conj) im = -im;
                    nrm = fmaf(rec.g, fmaf(re, re, im * im), nrm);
#pragma unroll
                    for (int t = 0; t < TC; ++t) acc[t] = fmaf(rec.u[t].x, re, fmaf(rec.u[t].y, im, acc[t]));
                }
            }
        }
        // ---- this CTA's share of the constant term (every rotation group walks all pixels of its chunks: each adds its own)
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) k0 += __shfl_xor_sync(0xffffffffu, k0, o);
        if (lane == 0) redd[warp] = k0;
        __syncthreads();
        double k0cta = 0.0;
        for (int w2 = 0; w2 < E6_THREADS / 32; ++w2) k0cta += redd[w2];
        // ---- partial sums of the pixel parts, parked in the record area, then one atomic per (rotation, translation)
        float* park = reinterpret_cast<float*>(tile);
        if (ph > 0) {
            float* pk = park + ((size_t)(ph - 1) * 32 + lane) * (TC + 1);
#pragma unroll
            for (int t = 0; t < TC; ++t) pk[t] = acc[t];
            pk[TC] = nrm;
        }
        __syncthreads();
        if (ph == 0 && rvalid) {
            const float* pk0 = park + (size_t)lane * (TC + 1);
            const size_t pstride = (size_t)32 * (TC + 1);
            double nn = (double)nrm;
            for (int j = 1; j < nParts; ++j) nn += (double)pk0[(j - 1) * pstride + TC];
#pragma unroll
            for (int t = 0; t < TC; ++t) {
                if (tbase + t >= A.nT) continue;
                double tot = (double)acc[t];
                for (int j = 1; j < nParts; ++j) tot += (double)pk0[(j - 1) * pstride + t];
                atomicAdd(&tab[(size_t)r * A.nT + tbase + t], k0cta + nn + tot);
            }
        }
    }
}

// table [image][nR][nT] (double) -> baseline, marginal weights, optional raw log-likelihoods.  grid = nAct, block = 256
__global__ void __launch_bounds__(256) expect_table_epilogue_kernel(const ExpectArgs A, const double* __restrict__ table, float* __restrict__ work)
{
    __shared__ float redf[8];
    __shared__ double redd[8];
    const int pos = blockIdx.x;
    const int p = A.order ? A.order[pos] : pos;
    if (A.active && !A.active[p]) return;
    const int nRT = A.nR * A.nT;
    float* sL = work + (size_t)pos * nRT;
    for (int i = threadIdx.x; i < nRT; i += 256) sL[i] = (float)table[(size_t)pos * nRT + i];
    __syncthreads();
    expect_epilogue<256>(A, p, sL, redf, redd);
}

}  // namespace thb
