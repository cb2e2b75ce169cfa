// thb_expect3.cuh - fused E kernel, local-search shape, direct gather from a "quad" volume layout.
//
// One CTA per image, 256 threads = 8 warps: warp w serves rotation group (w & 3) - one rotation sample
// per lane - and pixel half (w >> 2).  The image is walked in tiles of 128 pixels whose records (pixel
// coordinates and the image turned by the conjugate phase ramp of each translation, pre-multiplied for the
// expanded likelihood) are built once in shared memory and broadcast to all rotations.
//
// HBM layout of the projector volume for this kernel ("quad"): element (x, y, z) holds the four taps
//   { V(x,y,z), V(x+1,y,z), V(x,y+1,z), V(x+1,y+1,z) }      32 bytes, y+1 wrapped as the reference wraps it
// so that one 8-tap trilinear cell is TWO 256-bit loads (LDG.256: z and z+1) instead of eight 64-bit
// ones.  The gather of a cloud of orientations is bound by the number of distinct L1 lines a warp-wide load
// touches, not by bytes: fewer, wider loads move the same lines with a quarter of the instructions
// (measured: tools/gpu/gatherbench.cu).  4x the volume bytes (2.2 GB at box 256) buys that.
//
// Likelihood in the expanded form (see thb_expect2.cuh): 2 FMAs per (sample, translation).
// Coordinates, fold, floor, weights follow the reference exactly (src/Projector.cpp:356-374,
// src/Image/Volume.cpp:314-338, include/Functions/Interpolation.h:187-200).
#pragma once
#include <cuda_runtime.h>
#include "thb_math.cuh"
#include "thb_types.cuh"
#include "thb_expect2.cuh"   // PixelRec

namespace thb {

struct __align__(32) Quad { float2 v00, v10, v01, v11; };   // (x,y) (x+1,y) (x,y+1) (x+1,y+1)

__device__ __forceinline__ Quad ldg_quad(const Quad* p)
{
    Quad q;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(q.v00.x), "=f"(q.v00.y), "=f"(q.v10.x), "=f"(q.v10.y), "=f"(q.v01.x), "=f"(q.v01.y), "=f"(q.v11.x), "=f"(q.v11.y)
                 : "l"(p));
    return q;
}

// element index of quad (x, ym, zm) (memory coordinates, x in [0, n/2)): x fastest inside bricks of (2^LB)^3 quads,
// bricks x fastest.  LB = 0 is the plain [z][y][x] order.  Bricks keep the cells of a pixel tile within a few DRAM
// pages / one TLB entry instead of one page per (y,z) row.
__device__ __forceinline__ size_t quad_index(int x, int ym, int zm, int n, int LB)
{
    const int half = n >> 1, m = (1 << LB) - 1;
    const size_t brick = ((size_t)(zm >> LB) * (n >> LB) + (ym >> LB)) * (half >> LB) + (x >> LB);
    return (brick << (3 * LB)) | (size_t)((((zm & m) << LB) | (ym & m)) << LB | (x & m));
}

// "oct" variant of the layout: element (x, y, z) = { quad(x, y, z), quad(x, y, z+1) }, 64 bytes, z+1 wrapped as the
// reference wraps it: the whole trilinear cell is one 64-byte aligned chunk (two LDG.256 to one line, one DRAM burst
// of 64 bytes instead of two scattered 32-byte sectors).  8x the volume bytes (4.3 GB at box 256).
__global__ void build_oct_kernel(const float2* __restrict__ vol, int n, int pitch, int LB, Quad* __restrict__ out)
{
    const int half = n / 2;
    const size_t total = (size_t)n * n * half;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % half);
        const size_t row = i / half;
        const int y = (int)(row % n), z = (int)(row / n);
        const int y1 = (y + 1 == n) ? 0 : y + 1, z1 = (z + 1 == n) ? 0 : z + 1;
        const float2* r00 = vol + ((size_t)z * n + y) * pitch + x;
        const float2* r01 = vol + ((size_t)z * n + y1) * pitch + x;
        const float2* r10 = vol + ((size_t)z1 * n + y) * pitch + x;
        const float2* r11 = vol + ((size_t)z1 * n + y1) * pitch + x;
        Quad a, b;
        a.v00 = r00[0]; a.v10 = r00[1]; a.v01 = r01[0]; a.v11 = r01[1];
        b.v00 = r10[0]; b.v10 = r10[1]; b.v01 = r11[0]; b.v11 = r11[1];
        const size_t o = 2 * quad_index(x, y, z, n, LB);
        out[o] = a;
        out[o + 1] = b;
    }
}

// linear pitched float2 volume -> quad layout, x in [0, half)
__global__ void build_quad_kernel(const float2* __restrict__ vol, int n, int nz, int pitch, int LB, Quad* __restrict__ out)
{
    const int half = n / 2;
    const size_t total = (size_t)nz * n * half;          // nz = n, or 1 for a MODE_2D class average (LB = 0)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % half);
        const size_t row = i / half;
        const int y = (int)(row % n), z = (int)(row / n);
        const int y1 = (y + 1 == n) ? 0 : y + 1;     // memory row of (y + 1): the wrap of Volume.h:567-575
        const float2* r0 = vol + ((size_t)z * n + y) * pitch + x;
        const float2* r1 = vol + ((size_t)z * n + y1) * pitch + x;
        Quad q;
        q.v00 = r0[0]; q.v10 = r0[1]; q.v01 = r1[0]; q.v11 = r1[1];
        out[quad_index(x, y, z, n, LB)] = q;
    }
}

// Epilogue shared by the direct-gather E kernels: baseline (the reference's running baseline ends at the maximum), exp, and the
// prior-weighted marginals uR, uT, uC exactly as src/Optimiser.cpp:1383-1402.  sL = the [nR][nT] log-likelihood table.
template <int THREADS>
__device__ __forceinline__ void expect_epilogue(const ExpectArgs& A, int p, float* sL, float* redf, double* redd)
{
    const int tid = threadIdx.x, nRT = A.nR * A.nT;
    float m = -INFINITY;
    for (int i = tid; i < nRT; i += THREADS) m = fmaxf(m, sL[i]);
    m = block_reduce_max(m, redf);
    if (A.logL)
        for (int i = tid; i < nRT; i += THREADS) A.logL[(size_t)p * nRT + i] = sL[i];
    __syncthreads();
    for (int i = tid; i < nRT; i += THREADS) sL[i] = expf(sL[i] - m);
    __syncthreads();
    double uc = 0.0;
    for (int r = tid; r < A.nR; r += THREADS) {
        float s = 0.0f;
        for (int t = 0; t < A.nT; ++t) s = (float)((double)s + (double)sL[r * A.nT + t] * A.wT.at(p, t, 0));
        if (A.uR) A.uR[(size_t)p * A.nR + r] = s;
        uc += (double)s * A.wR.at(p, r, 0);
    }
    for (int t = tid; t < A.nT; t += THREADS) {
        float s = 0.0f;
        for (int r = 0; r < A.nR; ++r) s = (float)((double)s + (double)sL[r * A.nT + t] * A.wR.at(p, r, 0));
        if (A.uT) A.uT[(size_t)p * A.nT + t] = s;
    }
    uc = block_reduce_sum(uc, redd);
    if (tid == 0) {
        if (A.uC) A.uC[p] = (float)uc;
        if (A.base) A.base[p] = m;
    }
}

// CTF of one pixel for the defocus factor d, exactly the mixed float / double expression of src/Optimiser.cpp:1253-1268:
//   RFLOAT ki = K1 * defocusP * d * TSGSL_pow_2(f) + K2 * TSGSL_pow_4(f) - phaseShift      (d is a double: the first product is
//   a float, everything up to the assignment a double; TSGSL_pow_n return (float)(double power), src/Precision.cpp:263-276)
//   ctf = -TS_SQRT(1 - TSGSL_pow_2(ac)) * TS_SIN(ki) + ac * TS_COS(ki)
__device__ __forceinline__ float ctf_search_value(float defP, float f, double d, float K1, float K2, float phaseShift, float w1, float ac)
{
    const double f2d = (double)f * (double)f;
    const float pow2f = (float)f2d, pow4f = (float)(f2d * f2d);
    const float p1 = __fmul_rn(K1, defP), p2 = __fmul_rn(K2, pow4f);
    const float ki = (float)(__dsub_rn(__dadd_rn(__dmul_rn(__dmul_rn((double)p1, d), (double)pow2f), (double)p2), (double)phaseShift));
    return __fadd_rn(__fmul_rn(-w1, sinf(ki)), __fmul_rn(ac, cosf(ki)));
}

// Epilogue with the defocus dimension: sL = the [nR][nT][nD] table; marginals carry the prior weights of the OTHER dimensions
// (src/Optimiser.cpp:1383-1402): uR += s wT wD, uT += s wR wD, uD += s wR wT, uC += s wR wT wD
template <int THREADS>
__device__ __forceinline__ void expect_epilogue_ctf(const ExpectArgs& A, int p, float* sL, float* redf, double* redd)
{
    const int tid = threadIdx.x, nD = A.nD, nTD = A.nT * nD, nRT = A.nR * nTD;
    float m = -INFINITY;
    for (int i = tid; i < nRT; i += THREADS) m = fmaxf(m, sL[i]);
    m = block_reduce_max(m, redf);
    if (A.logL)
        for (int i = tid; i < nRT; i += THREADS) A.logL[(size_t)p * nRT + i] = sL[i];
    __syncthreads();
    for (int i = tid; i < nRT; i += THREADS) sL[i] = expf(sL[i] - m);
    __syncthreads();
    double uc = 0.0;
    for (int r = tid; r < A.nR; r += THREADS) {
        double s = 0.0;
        for (int t = 0; t < A.nT; ++t)
            for (int d = 0; d < nD; ++d) s += (double)sL[(r * A.nT + t) * nD + d] * A.wT.at(p, t, 0) * A.wD.at(p, d, 0);
        if (A.uR) A.uR[(size_t)p * A.nR + r] = (float)s;
        uc += s * A.wR.at(p, r, 0);
    }
    for (int t = tid; t < A.nT; t += THREADS) {
        double s = 0.0;
        for (int r = 0; r < A.nR; ++r)
            for (int d = 0; d < nD; ++d) s += (double)sL[(r * A.nT + t) * nD + d] * A.wR.at(p, r, 0) * A.wD.at(p, d, 0);
        if (A.uT) A.uT[(size_t)p * A.nT + t] = (float)s;
    }
    for (int d = tid; d < nD; d += THREADS) {
        double s = 0.0;
        for (int r = 0; r < A.nR; ++r)
            for (int t = 0; t < A.nT; ++t) s += (double)sL[(r * A.nT + t) * nD + d] * A.wR.at(p, r, 0) * A.wT.at(p, t, 0);
        if (A.uD) A.uD[(size_t)p * nD + d] = (float)s;
    }
    uc = block_reduce_sum(uc, redd);
    if (tid == 0) {
        if (A.uC) A.uC[p] = (float)uc;
        if (A.base) A.base[p] = m;
    }
}

constexpr int E3_THREADS = 256;
constexpr int E3_ROTS = 128;
constexpr int E3_TILE = 128;
// pixel record with TC translations per pass: TC = 9 = mLT of the local search (the default); the classification / global
// scans carry 15 so that 30 translations take two passes over the gather instead of four
template <int TC>
struct __align__(16) PixelRecT {
    double a, b;        // pf*iCol, pf*iRow
    float g, pad;       // sig * ctf^2
    float2 u[TC];       // -2 sig ctf dat conj(tra_t)
};
static_assert(sizeof(PixelRecT<E_TC>) == sizeof(PixelRec), "PixelRecT<9> is PixelRec");
constexpr int E3_TC_SCAN = 15;
constexpr size_t E3_SMEM_BYTES = E3_TILE * sizeof(PixelRec);     // + the [nR][nT] table for single-pass shapes

template <int MINB, bool OCT, bool M2D = false, int TC = E_TC, bool CTFS = false>
__global__ void __launch_bounds__(E3_THREADS, MINB) expect_direct_kernel(const ExpectArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PixelRecT<TC>* tile = reinterpret_cast<PixelRecT<TC>*>(smem_raw);
    __shared__ float sRC[TC], sRR[TC];
    __shared__ float redf[E3_THREADS / 32];
    __shared__ double redd[E3_THREADS / 32];

    const int pos = blockIdx.x;                             // launch position; A.order (compacted list of active particles) maps it to the particle
    const int p = A.order ? A.order[pos] : pos;
    if (A.active && !A.active[p]) return;
    const int img = A.imgIdx ? A.imgIdx[p] : p + A.imgBase;
    const int slot = (M2D && A.slotAll >= 0) ? A.slotAll : (A.slotOfImg ? A.slotOfImg[img] : 0);
    const Quad* __restrict__ vol = reinterpret_cast<const Quad*>(A.quads.p[slot]);
    const int n = A.vdim, half = n / 2;
    const int P = A.P;
    const float2* __restrict__ dat = A.dat + (size_t)img * P;
    const float* __restrict__ ctf = A.ctf + (size_t)img * P;
    const float* __restrict__ sig = A.sig + (size_t)img * P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nD = CTFS ? A.nD : 1;
    const int nRT = A.nR * A.nT * nD;
    const bool single = !CTFS && A.nR <= E3_ROTS && A.nT <= TC;
    float* sL = single ? reinterpret_cast<float*>(smem_raw + E3_TILE * sizeof(PixelRecT<TC>)) : A.work + (size_t)pos * nRT;
    const float* __restrict__ defP = CTFS ? A.defP + (size_t)img * P : nullptr;
    float cK1 = 0.f, cK2 = 0.f, cPs = 0.f, cAc = 0.f, cW1 = 0.f;
    if (CTFS) {
        cK1 = A.ctfK[4 * p]; cK2 = A.ctfK[4 * p + 1]; cPs = A.ctfK[4 * p + 2]; cAc = A.ctfK[4 * p + 3];
        cW1 = sqrtf(1.0f - (float)((double)cAc * (double)cAc));
    }
    double k0sum = 0.0;          // sum_i sig_i |dat_i|^2
    const int LB = A.quadBrick;

    for (int rbase = 0; rbase < A.nR; rbase += E3_ROTS) {
        const int nRc = min(E3_ROTS, A.nR - rbase);
        // the 8 warps are split into G rotation groups (32 rotations each) x 8 / G pixel parts: a full cloud of 125 uses
        // 4 x 2, the small supports of the scans' chunks, of demo_3D.json's mLR = 25 or of MODE_2D's mLR = 9 use 1 x 8
        const int G = nRc > 64 ? 4 : nRc > 32 ? 2 : 1;
        const int nParts = (E3_THREADS / 32) / G;
        const int rloc = (warp % G) * 32 + lane;   // rotation slot of this thread within the pass
        const int ph = warp / G;                   // pixel part
        const bool rvalid = rloc < nRc;
        // Rotation slots are handed out in the order of the cloud along its widest axis, so that the 32 lanes of a warp
        // hold a compact sub-cloud: a warp-wide load then touches fewer distinct lines.  rsrc = rotation of this slot.
        int rsrc = rloc;
        if (A.sortRot && nRc > 32) {
            __syncthreads();
            float4* sQ = reinterpret_cast<float4*>(tile);          // keys live in the record area until the first tile
            float key[3] = {0.f, 0.f, 0.f};
            if (ph == 0 && rvalid) {
                // vector part of conj(q_0) * q_r (hemisphere-aligned): the small rotation that takes rotation 0 to r
                double q0[4], q[4];
                for (int c = 0; c < 4; ++c) { q0[c] = A.quat.at(p, rbase, c); q[c] = A.quat.at(p, rbase + rloc, c); }
                const double w = q0[0] * q[0] + q0[1] * q[1] + q0[2] * q[2] + q0[3] * q[3];
                const double sg = w < 0 ? -1.0 : 1.0;
                key[0] = (float)(sg * (q0[0] * q[1] - q0[1] * q[0] - q0[2] * q[3] + q0[3] * q[2]));
                key[1] = (float)(sg * (q0[0] * q[2] + q0[1] * q[3] - q0[2] * q[0] - q0[3] * q[1]));
                key[2] = (float)(sg * (q0[0] * q[3] - q0[1] * q[2] + q0[2] * q[1] - q0[3] * q[0]));
            }
            int* sSlot = reinterpret_cast<int*>(sQ + E3_ROTS);
            if (ph == 0) {
                sQ[rloc] = make_float4(key[0], key[1], key[2], rvalid ? 1.f : 0.f);
                sSlot[rloc] = rloc;                                // stays a valid index even for NaN keys
            }
            __syncthreads();
            // axis of the largest variance (every thread computes it: 128 broadcast reads)
            float s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
            for (int j = 0; j < nRc; ++j) {
                const float4 o = sQ[j];
                s1[0] += o.x; s1[1] += o.y; s1[2] += o.z;
                s2[0] += o.x * o.x; s2[1] += o.y * o.y; s2[2] += o.z * o.z;
            }
            float best = -1.f;
            int ax = 0;
            for (int c = 0; c < 3; ++c) {
                const float var = s2[c] - s1[c] * s1[c] / (float)nRc;
                if (var > best) { best = var; ax = c; }
            }
            // slot -> rotation: the slot of rotation r is its rank along that axis
            if (ph == 0 && rvalid) {
                const float mine = ax == 0 ? key[0] : ax == 1 ? key[1] : key[2];
                int rank = 0;
                for (int j = 0; j < nRc; ++j) {
                    const float4 o = sQ[j];
                    const float kj = ax == 0 ? o.x : ax == 1 ? o.y : o.z;
                    rank += (kj < mine) || (kj == mine && j < rloc);
                }
                sSlot[rank] = rloc;
            }
            __syncthreads();
            if (rvalid) rsrc = min(max(sSlot[rloc], 0), nRc - 1);
            __syncthreads();
        }
        Rot2 rot;
        {
            double q[4] = {1.0, 0.0, 0.0, 0.0};
            if (rvalid)
                for (int c = 0; c < (M2D ? 2 : 4); ++c) q[c] = A.quat.at(p, rbase + rsrc, c);
            rot = make_rot2(q, M2D);
        }
        for (int iD = 0; iD < nD; ++iD)
        for (int tbase = 0; tbase < A.nT; tbase += TC) {
            const double dfac = CTFS ? A.dpar.at(p, iD, 0) : 1.0;
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
            const bool firstPass = (rbase == 0 && tbase == 0 && iD == 0);

            for (int tile0 = 0; tile0 < P; tile0 += E3_TILE) {
                const int cnt = min(E3_TILE, P - tile0);
                __syncthreads();   // previous tile consumed (also orders the sRC / sRR writes)
                {
                    // pixel records: 2 threads per pixel, translations split between them
                    const int k = tid >> 1, sub = tid & 1;
                    if (k < cnt) {
                        const int i = tile0 + k;
                        const int4 c = A.pix[i];
                        const float2 d = dat[i];
                        const float cf = CTFS ? ctf_search_value(defP[i], A.freq[i], dfac, cK1, cK2, cPs, cW1, cAc) : ctf[i];
                        const float sg = sig[i];
                        const float m2 = -2.0f * sg * cf;
                        PixelRecT<TC>& rec = tile[k];
                        if (sub == 0) {
                            rec.a = (double)c.x;
                            rec.b = (double)c.y;
                            rec.g = sg * cf * cf;
                            rec.pad = 0.0f;
                            if (firstPass) k0sum += (double)(sg * (d.x * d.x + d.y * d.y));
                        }
#pragma unroll
                        for (int t = 0; t < TC; ++t) {
                            if ((t & 1) != sub) continue;
                            const float phs = translate_phase(c.z, c.w, sRC[t], sRR[t]);
                            float s, co;
                            sincosf(phs, &s, &co);
                            // tra = (cos(-ph), sin(-ph)); dat * conj(tra) = dat * (co + i s)
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
                        // MODE_2D: z = 0 exactly, the second plane carries weight 0 and does not exist
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
            // ---- end of the pass: constant term, halves
            __syncthreads();
            if (firstPass) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) k0sum += __shfl_xor_sync(0xffffffffu, k0sum, o);
                if (lane == 0) redd[warp] = k0sum;
                __syncthreads();
                double s = 0.0;
                for (int w2 = 0; w2 < E3_THREADS / 32; ++w2) s += redd[w2];
                k0sum = s;
                __syncthreads();
            }
            // partial sums of pixel parts 1 .. nParts-1 are parked in the record area ((nParts-1) x 32 G x (TC+1) floats <= 14 KB)
            float* park = reinterpret_cast<float*>(tile);
            if (ph > 0) {
                float* pk = park + ((size_t)(ph - 1) * (32 * G) + rloc) * (TC + 1);
#pragma unroll
                for (int t = 0; t < TC; ++t) pk[t] = acc[t];
                pk[TC] = nrm;
            }
            __syncthreads();
            if (ph == 0 && rvalid) {
                const float* pk0 = park + (size_t)rloc * (TC + 1);
                const size_t pstride = (size_t)(32 * G) * (TC + 1);
                double nn = (double)nrm;
                for (int j = 1; j < nParts; ++j) nn += (double)pk0[(j - 1) * pstride + TC];
#pragma unroll
                for (int t = 0; t < TC; ++t) {
                    if (tbase + t >= A.nT) continue;
                    double tot = (double)acc[t];
                    for (int j = 1; j < nParts; ++j) tot += (double)pk0[(j - 1) * pstride + t];
                    sL[((size_t)(rbase + rsrc) * A.nT + tbase + t) * nD + iD] = (float)(k0sum + nn + tot);
                }
            }
        }
    }
    __syncthreads();

    if (CTFS)
        expect_epilogue_ctf<E3_THREADS>(A, p, sL, redf, redd);
    else
        expect_epilogue<E3_THREADS>(A, p, sL, redf, redd);
}

}  // namespace thb
