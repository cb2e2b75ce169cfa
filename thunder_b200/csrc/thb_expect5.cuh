// thb_expect5.cuh - fused E kernel, local-search shape, PIXELS ON THE LANES.
//
// thb_expect3.cuh puts the 32 rotations of a cloud on the lanes of a warp: one warp-wide load fetches the cells of ONE pixel
// under 32 rotations.  For the wide clouds of the benchmark regime those cells are scattered over +-20 voxels: 32 different
// bricks, 32 different DRAM pages per load.  Here the lanes of a warp hold 32 consecutive pixels of the blocked pixel order
// (an 8 x 4 patch of the image) under ONE rotation: their cells lie 2 voxels apart on a plane, i.e. in a handful of 4x4x4
// bricks (tools/gpu/drambench.cu measures what that locality is worth to the DRAM).  Cost: the sums over pixels now run
// across lanes - one butterfly of 10 values per (rotation, 128-pixel tile) - and the per-pixel records are read per lane
// (SoA in shared memory) instead of by broadcast.
//
// One CTA per image, 256 threads = 8 warps; warp w owns rotations r = w, w + 8, ... of a 128-rotation pass (so the running
// sums of a rotation, kept in shared memory, are only ever touched by one warp).  Records, expanded likelihood, passes,
// epilogue: as thb_expect3.cuh.
#pragma once
#include <cuda_runtime.h>
#include "thb_math.cuh"
#include "thb_types.cuh"
#include "thb_expect3.cuh"

namespace thb {

constexpr int E5_THREADS = 256;
constexpr int E5_ROTS = 128;
constexpr int E5_TILE = 128;
constexpr int E5_NV = E_TC + 1;                       // 9 translation sums + the norm term
struct E5Smem {
    double a[E5_TILE], b[E5_TILE];                    // pf*iCol, pf*iRow
    float2 u[E_TC][E5_TILE];                          // -2 sig ctf dat conj(tra_t)
    float g[E5_TILE];                                 // sig ctf^2
    Rot2 rot[E5_ROTS];
    float acc[E5_ROTS][E5_NV];
};
constexpr size_t E5_SMEM_BYTES = sizeof(E5Smem);      // + the [nR][nT] table for single-pass shapes

template <bool OCT, bool M2D>
__global__ void __launch_bounds__(E5_THREADS, 2) expect_pix_kernel(const ExpectArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    E5Smem& S = *reinterpret_cast<E5Smem*>(smem_raw);
    __shared__ float sRC[E_TC], sRR[E_TC];
    __shared__ float redf[E5_THREADS / 32];
    __shared__ double redd[E5_THREADS / 32];

    const int p = blockIdx.x;
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
    const int nRT = A.nR * A.nT;
    const bool single = A.nR <= E5_ROTS && A.nT <= E_TC;
    float* sL = single ? reinterpret_cast<float*>(smem_raw + E5_SMEM_BYTES) : A.work + (size_t)p * nRT;
    double k0sum = 0.0;          // sum_i sig_i |dat_i|^2
    const int LB = A.quadBrick;

    for (int rbase = 0; rbase < A.nR; rbase += E5_ROTS) {
        const int nRc = min(E5_ROTS, A.nR - rbase);
        __syncthreads();
        if (tid < nRc) {
            double q[4] = {1.0, 0.0, 0.0, 0.0};
            for (int c = 0; c < (M2D ? 2 : 4); ++c) q[c] = A.quat.at(p, rbase + tid, c);
            S.rot[tid] = make_rot2(q, M2D);
        }
        for (int tbase = 0; tbase < A.nT; tbase += E_TC) {
            __syncthreads();
            if (tid < E_TC) {
                const int t = tbase + tid;
                float tx = 0.0f, ty = 0.0f;
                if (t < A.nT) {
                    tx = (float)A.tran.at(p, t, 0);
                    ty = (float)A.tran.at(p, t, 1);
                }
                sRC[tid] = tx / (float)A.N;
                sRR[tid] = ty / (float)A.N;
            }
            for (int i = tid; i < E5_ROTS * E5_NV; i += E5_THREADS) (&S.acc[0][0])[i] = 0.0f;
            const bool firstPass = (rbase == 0 && tbase == 0);

            for (int tile0 = 0; tile0 < P; tile0 += E5_TILE) {
                const int cnt = min(E5_TILE, P - tile0);
                __syncthreads();   // previous tile consumed (also orders the sRC / sRR / acc writes)
                {
                    // pixel records: 2 threads per pixel, translations split between them
                    const int k = tid >> 1, sb = tid & 1;
                    if (k < cnt) {
                        const int i = tile0 + k;
                        const int4 c = A.pix[i];
                        const float2 d = dat[i];
                        const float cf = ctf[i], sg = sig[i];
                        const float m2 = -2.0f * sg * cf;
                        if (sb == 0) {
                            S.a[k] = (double)c.x;
                            S.b[k] = (double)c.y;
                            S.g[k] = sg * cf * cf;
                            if (firstPass) k0sum += (double)(sg * (d.x * d.x + d.y * d.y));
                        }
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) {
                            if ((t & 1) != sb) continue;
                            const float phs = translate_phase(c.z, c.w, sRC[t], sRR[t]);
                            float s, co;
                            sincosf(phs, &s, &co);
                            S.u[t][k] = make_float2(m2 * (d.x * co - d.y * s), m2 * (d.x * s + d.y * co));
                        }
                    } else if (k < E5_TILE && sb == 0) {
                        // pad of the last tile: a pixel at the origin with zero weight (valid address, contributes nothing)
                        S.a[k] = 0.0; S.b[k] = 0.0; S.g[k] = 0.0f;
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) S.u[t][k] = make_float2(0.f, 0.f);
                    }
                }
                __syncthreads();
                for (int r = warp; r < nRc; r += E5_THREADS / 32) {
                    const Rot2 rot = S.rot[r];
                    float acc[E5_NV];
#pragma unroll
                    for (int t = 0; t < E5_NV; ++t) acc[t] = 0.0f;
#pragma unroll
                    for (int j = 0; j < E5_TILE / 32; ++j) {
                        const int k = j * 32 + lane;
                        float x, y, z;
                        slice_coord(rot, S.a[k], S.b[k], x, y, z);
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
                        acc[E_TC] = fmaf(S.g[k], fmaf(re, re, im * im), acc[E_TC]);
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) {
                            const float2 u = S.u[t][k];
                            acc[t] = fmaf(u.x, re, fmaf(u.y, im, acc[t]));
                        }
                    }
                    // sums over the 128 pixels of the tile: butterfly, then lane t adds value t to the rotation's running sums
#pragma unroll
                    for (int t = 0; t < E5_NV; ++t) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
                    }
                    float mine = acc[0];
#pragma unroll
                    for (int t = 1; t < E5_NV; ++t) mine = lane == t ? acc[t] : mine;
                    if (lane < E5_NV) S.acc[r][lane] += mine;
                }
            }
            // ---- end of the pass: constant term, table
            __syncthreads();
            if (firstPass) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) k0sum += __shfl_xor_sync(0xffffffffu, k0sum, o);
                if (lane == 0) redd[warp] = k0sum;
                __syncthreads();
                double s = 0.0;
                for (int w2 = 0; w2 < E5_THREADS / 32; ++w2) s += redd[w2];
                k0sum = s;
                __syncthreads();
            }
            for (int i = tid; i < nRc * E_TC; i += E5_THREADS) {
                const int r = i / E_TC, t = i - r * E_TC;
                if (tbase + t < A.nT)
                    sL[(size_t)(rbase + r) * A.nT + tbase + t] = (float)(k0sum + (double)S.acc[r][E_TC] + (double)S.acc[r][t]);
            }
        }
    }
    __syncthreads();

    expect_epilogue<E5_THREADS>(A, p, sL, redf, redd);
}

}  // namespace thb
