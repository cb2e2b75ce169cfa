// thb_expect7.cuh - fused E kernel, local-search shape, SEVERAL rotations per lane.
//
// ncu on the default kernel (thb_expect3.cuh, profiles/r02_ncu_expect_direct_b5000_details.txt) shows the L1/TEX pipe as the
// busiest unit (89 %), ahead of DRAM (65 %).  Two things load it per (rotation, pixel) sample: the two scattered 256-bit
// gathers (one cycle per 32-byte sector) and the BROADCAST of the pixel record from shared memory - 96 bytes to every lane,
// 24 data-return cycles per warp-wide sample against 64 for the gathers.  The record is the same for every rotation, so a lane
// that carries RPL rotations amortises one record read over RPL samples: 24 / RPL cycles.  The record also shrinks to 80 bytes
// (pixel coordinates as two int16 instead of two doubles; the int -> double conversion is one integer op + one DADD, no
// conversion-pipe instruction).
//
// Everything else is the default kernel's: one CTA per image, tiles of 128 pixels, the cell ("oct") or quad layout, the expanded
// likelihood, coordinates / fold / floor / weights as the reference (src/Projector.cpp:356-374, src/Image/Volume.cpp:314-338,
// include/Functions/Interpolation.h:187-200), the epilogue of src/Optimiser.cpp:1383-1402.  Shapes: nR <= 128, nT <= 9, MODE_3D,
// no CTF search (the launcher falls back to expect_direct_kernel otherwise).
#pragma once
#include <cuda_runtime.h>
#include "thb_expect3.cuh"

namespace thb {

struct __align__(16) PixelRec7 {
    int ab;             // (pf*iRow) << 16 | (pf*iCol & 0xffff)
    float g;            // sig * ctf^2
    float2 u[E_TC];     // -2 sig ctf dat conj(tra_t)
};
static_assert(sizeof(PixelRec7) == 80, "PixelRec7 is five 16-byte words");

// exact int -> double without the conversion pipe: 2^52 + 2^31 + i has the bit pattern {0x43300000, 0x80000000 ^ i}
__device__ __forceinline__ double int_to_double(int i)
{
    return __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)i)) - 4503601774854144.0;
}

template <int RPL>
constexpr size_t e7_smem_bytes(int nR, int nT)
{
    // record tile and the parking area of the partial sums share one region; the [nR][nT] table follows
    const size_t tile = E3_TILE * sizeof(PixelRec7);
    const int G = (E3_ROTS / 32) / RPL, nParts = (E3_THREADS / 32) / G;
    const size_t park = (size_t)(nParts - 1) * E3_ROTS * (E_TC + 1) * sizeof(float);
    return (tile > park ? tile : park) + sizeof(float) * (size_t)nR * nT;
}

template <int RPL, bool OCT>
__global__ void __launch_bounds__(E3_THREADS, RPL >= 4 ? 1 : 2) expect_multi_kernel(const ExpectArgs A)
{
    constexpr int G = (E3_ROTS / 32) / RPL;            // rotation groups of 32 * RPL
    constexpr int NPARTS = (E3_THREADS / 32) / G;      // pixel parts
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PixelRec7* tile = reinterpret_cast<PixelRec7*>(smem_raw);
    __shared__ float sRC[E_TC], sRR[E_TC];
    __shared__ float redf[E3_THREADS / 32];
    __shared__ double redd[E3_THREADS / 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = A.vdim;
    const int P = A.P;
    const int nR = A.nR, nT = A.nT;
    constexpr size_t tileBytes = E3_TILE * sizeof(PixelRec7);
    constexpr size_t parkBytes = (size_t)(NPARTS - 1) * E3_ROTS * (E_TC + 1) * sizeof(float);
    float* sL = reinterpret_cast<float*>(smem_raw + (tileBytes > parkBytes ? tileBytes : parkBytes));
    const int LB = A.quadBrick;
    const int g = warp % G, ph = warp / G;
    // lockstep: barrier j of wave w has its own counter lockCtr[w * K + j]; it is complete when every CTA of wave w has arrived
    const int nTilesImg = (P + E3_TILE - 1) / E3_TILE;
    const int lockTiles = A.lockCtr ? max(1, A.lockTiles) : 0;
    const unsigned K = lockTiles ? (unsigned)((nTilesImg + lockTiles - 1) / lockTiles) : 0u;
    bool lockOn = A.lockCtr != nullptr;

    for (int it = blockIdx.x, wave = 0; it < A.nAct; it += gridDim.x, ++wave) {
    const int p = A.order ? A.order[it] : it;
    __syncthreads();       // the previous image's epilogue is done with the shared arrays
    if (A.active && !A.active[p]) {
        if (A.lockCtr)                                           // arrive at all barriers of this wave at once
            for (unsigned j = tid; j < K; j += E3_THREADS) atomicAdd(A.lockCtr + (unsigned)wave * K + j, 1u);
        continue;
    }
    const int img = A.imgIdx ? A.imgIdx[p] : p + A.imgBase;
    const int slot = A.slotOfImg ? A.slotOfImg[img] : 0;
    const Quad* __restrict__ vol = reinterpret_cast<const Quad*>(A.quads.p[slot]);
    const float2* __restrict__ dat = A.dat + (size_t)img * P;
    const float* __restrict__ ctf = A.ctf + (size_t)img * P;
    const float* __restrict__ sig = A.sig + (size_t)img * P;

    // rotation j of this lane: slot g*32*RPL + j*32 + lane (consecutive lanes hold consecutive rotations)
    Rot2 rot[RPL];
    bool rvalid[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
        const int r = (g * RPL + j) * 32 + lane;
        rvalid[j] = r < nR;
        double q[4] = {1.0, 0.0, 0.0, 0.0};
        if (rvalid[j])
            for (int c = 0; c < 4; ++c) q[c] = A.quat.at(p, r, c);
        rot[j] = make_rot2(q, 0);
    }
    if (tid < E_TC) {
        float tx = 0.0f, ty = 0.0f;
        if (tid < nT) {
            tx = (float)A.tran.at(p, tid, 0);
            ty = (float)A.tran.at(p, tid, 1);
        }
        sRC[tid] = tx / (float)A.N;
        sRR[tid] = ty / (float)A.N;
    }
    float acc[RPL][E_TC], nrm[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
        nrm[j] = 0.0f;
#pragma unroll
        for (int t = 0; t < E_TC; ++t) acc[j][t] = 0.0f;
    }
    double k0sum = 0.0;          // sum_i sig_i |dat_i|^2

    for (int tile0 = 0; tile0 < P; tile0 += E3_TILE) {
        const int cnt = min(E3_TILE, P - tile0);
        __syncthreads();   // previous tile consumed (also orders the sRC / sRR writes)
        const bool lockHere = lockTiles && (tile0 / E3_TILE) % lockTiles == 0;
        const unsigned lockJ = lockHere ? (unsigned)((tile0 / E3_TILE) / lockTiles) : 0u;
        if (lockHere && tid == 0) atomicAdd(A.lockCtr + (unsigned)wave * K + lockJ, 1u);     // done with everything before barrier lockJ
        {
            // pixel records: 2 threads per pixel, translations split between them
            const int k = tid >> 1, sub = tid & 1;
            if (k < cnt) {
                const int i = tile0 + k;
                const int4 c = A.pix[i];
                const float2 d = dat[i];
                const float cf = ctf[i];
                const float sg = sig[i];
                const float m2 = -2.0f * sg * cf;
                PixelRec7& rec = tile[k];
                if (sub == 0) {
                    rec.ab = (c.y << 16) | (c.x & 0xffff);
                    rec.g = sg * cf * cf;
                    k0sum += (double)(sg * (d.x * d.x + d.y * d.y));
                }
#pragma unroll
                for (int t = 0; t < E_TC; ++t) {
                    if ((t & 1) != sub) continue;
                    const float phs = translate_phase(c.z, c.w, sRC[t], sRR[t]);
                    float s, co;
                    sincosf(phs, &s, &co);
                    rec.u[t] = make_float2(m2 * (d.x * co - d.y * s), m2 * (d.x * s + d.y * co));
                }
            }
        }
        if (lockHere && lockOn && tid == 0) {
            // wait (behind the record build) until every CTA has arrived at barrier lockJ - lockWindow; the spin is bounded:
            // lockstep is a matter of speed, not of correctness
            // (one counter per barrier, indexed across waves: every wave but the last has gridDim.x participants)
            const int Jg = wave * (int)K + (int)lockJ - A.lockWindow;
            if (Jg >= 0) {
                const int wv = Jg / (int)K;
                const unsigned nwv = (unsigned)min((int)gridDim.x, A.nAct - wv * (int)gridDim.x);
                const volatile unsigned int* c = A.lockCtr + Jg;
                int spins = 0;
                while (*c < nwv) {
                    __nanosleep(200);
                    if (++spins > 2000000) { lockOn = false; break; }
                }
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int k = ph; k < cnt; k += NPARTS) {
            const PixelRec7& rec = tile[k];
            const int ab = rec.ab;
            const double a = int_to_double((ab << 16) >> 16), b = int_to_double(ab >> 16);
            Quad qa[RPL], qb[RPL];
            float w[RPL][8];
            bool cj[RPL];
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                float x, y, z;
                slice_coord(rot[j], a, b, x, y, z);
                int xb, yb, zb;
                float xd, yd, zd;
                cj[j] = fold_floor_fast(x, y, z, xb, yb, zb, xd, yd, zd);
                const int x0 = xb - THB_FLOOR_BIAS, y0 = yb - THB_FLOOR_BIAS, z0 = zb - THB_FLOOR_BIAS;
                const int ym = y0 < 0 ? y0 + n : y0;
                const int zm = z0 < 0 ? z0 + n : z0;
                const int zm1 = (z0 + 1 < 0) ? z0 + 1 + n : z0 + 1;
                const Quad* q0 = OCT ? vol + 2 * quad_index(x0, ym, zm, n, LB) : vol + quad_index(x0, ym, zm, n, LB);
                const Quad* q1 = OCT ? q0 + 1 : vol + quad_index(x0, ym, zm1, n, LB);
                qa[j] = ldg_quad(q0);
                qb[j] = ldg_quad(q1);
                tri_weights(xd, yd, zd, w[j]);
            }
            const float gk = rec.g;
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                const Quad& qA = qa[j];
                const Quad& qB = qb[j];
                float re = qA.v00.x * w[j][0], im = qA.v00.y * w[j][0];
                re = fmaf(qA.v10.x, w[j][1], re); im = fmaf(qA.v10.y, w[j][1], im);
                re = fmaf(qA.v01.x, w[j][2], re); im = fmaf(qA.v01.y, w[j][2], im);
                re = fmaf(qA.v11.x, w[j][3], re); im = fmaf(qA.v11.y, w[j][3], im);
                re = fmaf(qB.v00.x, w[j][4], re); im = fmaf(qB.v00.y, w[j][4], im);
                re = fmaf(qB.v10.x, w[j][5], re); im = fmaf(qB.v10.y, w[j][5], im);
                re = fmaf(qB.v01.x, w[j][6], re); im = fmaf(qB.v01.y, w[j][6], im);
                re = fmaf(qB.v11.x, w[j][7], re); im = fmaf(qB.v11.y, w[j][7], im);
                if (cj[j]) im = -im;
                nrm[j] = fmaf(gk, fmaf(re, re, im * im), nrm[j]);
#pragma unroll
                for (int t = 0; t < E_TC; ++t) acc[j][t] = fmaf(rec.u[t].x, re, fmaf(rec.u[t].y, im, acc[j][t]));
            }
        }
    }
    // ---- constant term, partial sums of the pixel parts
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k0sum += __shfl_xor_sync(0xffffffffu, k0sum, o);
    if (lane == 0) redd[warp] = k0sum;
    __syncthreads();
    {
        double s = 0.0;
        for (int w2 = 0; w2 < E3_THREADS / 32; ++w2) s += redd[w2];
        k0sum = s;
    }
    __syncthreads();
    float* park = reinterpret_cast<float*>(smem_raw);      // [part - 1][rotation slot][E_TC + 1]
    if (ph > 0) {
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            float* pk = park + ((size_t)(ph - 1) * E3_ROTS + (g * RPL + j) * 32 + lane) * (E_TC + 1);
#pragma unroll
            for (int t = 0; t < E_TC; ++t) pk[t] = acc[j][t];
            pk[E_TC] = nrm[j];
        }
    }
    __syncthreads();
    if (ph == 0) {
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            if (!rvalid[j]) continue;
            const int r = (g * RPL + j) * 32 + lane;
            const float* pk0 = park + (size_t)r * (E_TC + 1);
            constexpr size_t pstride = (size_t)E3_ROTS * (E_TC + 1);
            double nn = (double)nrm[j];
            for (int q = 1; q < NPARTS; ++q) nn += (double)pk0[(q - 1) * pstride + E_TC];
#pragma unroll
            for (int t = 0; t < E_TC; ++t) {
                if (t >= nT) continue;
                double tot = (double)acc[j][t];
                for (int q = 1; q < NPARTS; ++q) tot += (double)pk0[(q - 1) * pstride + t];
                sL[(size_t)r * nT + t] = (float)(k0sum + nn + tot);
            }
        }
    }
    __syncthreads();
    expect_epilogue<E3_THREADS>(A, p, sL, redf, redd);
    }
}

}  // namespace thb
