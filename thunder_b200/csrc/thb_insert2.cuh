// thb_insert2.cuh - fused M kernel, ordered by z-slab of the accumulator ("slab insert"), MODE_3D.
//
// Why.  The scatter of Reconstructor::insertP (src/Reconstructor.cpp:782-863 -> Volume::addFT, src/Image/Volume.cpp:340-375,
// 565-712) is eight 16-byte reductions per sample into a 1.08 GB accumulator (box 256).  Against a footprint that large every
// reduction is a DRAM read-modify-write of a random 32-byte sector: measured 3.9 - 4.6 G samples/s (tools/gpu/l2bench.cu,
// profiles/r02_l2bench_gather_red_vs_footprint.log), and the image-ordered kernel of round 1 sat at 7.3 G samples/s with an L2
// hit rate of 53 %.  The same reductions into a footprint that FITS the L2 (<= 64 MB) run at 22 - 26 G samples/s, resolved by the
// L2's own atomic units without touching DRAM.  So the work is re-ordered: the grid is (image, slab) with the image index
// fastest, slab s = the accumulator planes z0 in [zlo, zhi); a CTA scatters exactly the samples of its image whose cell base
// falls into its slab.  CTAs are dispatched in grid order, so at any time the whole chip works on one slab (~45 MB) of one
// half map, which stays L2-resident, is read from DRAM once and written back once.  Sums are the reference's up to fp32 order.
//
// How a CTA finds its samples without testing every pixel: the M pixel list is kept in row-major runs ("segments": one row
// j, consecutive columns i).  Along a segment the slice coordinate is linear in i - z = R20 pf i + R21 pf j, x likewise - so the
// pixels whose z0 = floor(+-z) lies in the slab form at most two index intervals (one per side of the Hermitian fold x = 0)
// plus a few pixels around the fold.  One thread computes the intervals of one (rotation group, segment) CONSERVATIVELY (in
// double, margins far above the fp32 rounding of the exact coordinate), a block-wide prefix sum flattens them, and every
// element then repeats the EXACT coordinate arithmetic of the scatter (slice_coord -> fold -> floor) and drops out unless its
// z0 is inside the slab: slabs are disjoint and cover all z0, so every sample is scattered exactly once.
//
// Draws of one image that share a bit-identical rotation are merged as in the round-1 kernel (thb_kernels.cuh); the grouping,
// the phase-ramp slopes of translate() (src/Image/ImageFunctions.cpp:471-492) and insertDir (src/Reconstructor.cpp:407-422)
// are done once per image by insert_prep_kernel.
#pragma once
#include <cuda_runtime.h>
#include "thb_math.cuh"
#include "thb_types.cuh"
#include "thb_kernels.cuh"
#include "thb_slab.cuh"

namespace thb {

constexpr int M2_THREADS = 256;
constexpr int M2_MAXD = 256;        // draws (and rotation groups) per image the slab kernel stages in shared memory
constexpr int M2_MAXSEG = 512;      // row segments of the M pixel list (box 512: 510)
constexpr int M2_ENT = 1024;        // (group, segment) entries flattened per round

// per-image output of insert_prep_kernel, `stride` bytes apart:
//   PrepHdr | Rot2 rot[maxD] | int grpEnd[maxD] | float4 ramp[maxD]        (draws sorted by group; grpEnd = end position;
//   ramp = phase-ramp slopes of the draw's translation and, with CTF search, its defocusU d, defocusV d)
struct PrepHdr { int nGrp, nDraw; float wgt; int slot; };
__host__ __device__ inline size_t prep_stride(int maxD) { return sizeof(PrepHdr) + (size_t)maxD * (sizeof(Rot2) + sizeof(int) + sizeof(float4)); }

struct InsertSlabArgs {
    InsertArgs a;
    const Seg* seg;           // row-major runs of the pixel list
    int nSeg;
    const int* order;         // grid x -> image position l (images of one slot adjacent), or null
    unsigned char* prep;
    int maxD;                 // capacity of the per-image arrays (>= total draws of an image)
    int pf;                   // padding factor: pixel (i, j) of a segment sits at (pf i, pf j) in the volume
    float rMaxPad;            // largest |(pf i, pf j)| of the pixel list
    int zMin, th;             // slab s covers z0 in [zMin + s th, zMin + (s + 1) th)
};

// ---------------------------------------------------------------------------------------------------------------------
// once per image: group the draws by rotation, phase-ramp slopes, insertDir.  grid = nImg, block = 128
// ---------------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) insert_prep_kernel(const InsertSlabArgs S)
{
    const InsertArgs& A = S.a;
    __shared__ double sQ[128][4];
    __shared__ unsigned short sRep[128];
    __shared__ double redd[4];
    const int l = blockIdx.x, tid = threadIdx.x;
    const int img = A.imgIdx ? A.imgIdx[l] : l + A.imgBase;
    const int slot = A.slotOfImg ? A.slotOfImg[img] : 0;
    const double ox = A.offS ? A.offS[2 * l] : 0.0, oy = A.offS ? A.offS[2 * l + 1] : 0.0;
    unsigned char* base = S.prep + (size_t)l * prep_stride(S.maxD);
    PrepHdr* hdr = reinterpret_cast<PrepHdr*>(base);
    Rot2* gRot = reinterpret_cast<Rot2*>(base + sizeof(PrepHdr));
    int* gEnd = reinterpret_cast<int*>(gRot + S.maxD);
    float4* ramp = reinterpret_cast<float4*>(gEnd + S.maxD);

    const int mTot = A.drawCount ? max(0, min(A.mReco, A.drawCount[l])) : A.mReco;
    int nGrpTot = 0;
    for (int mbase = 0; mbase < mTot; mbase += 128) {
        const int mcnt = min(128, mTot - mbase);
        __syncthreads();
        double dx = 0.0, dy = 0.0, dz = 0.0;
        Rot2 rot;
        float rc = 0.f, rr = 0.f, dUs = 0.f, dVs = 0.f;
        if (tid < mcnt) {
            const int m = mbase + tid;
            const long long sr = A.drawR ? A.drawR[(size_t)l * A.mReco + m] : m;
            const long long st = A.drawT ? A.drawT[(size_t)l * A.mReco + m] : m;
            double q[4];
            for (int c = 0; c < 4; ++c) { q[c] = A.nr.at(l, sr, c); sQ[tid][c] = q[c]; }
            rot = quat_to_rot2(q);
            const double tx = A.nt.at(l, st, 0) - ox, ty = A.nt.at(l, st, 1) - oy;
            rc = (float)(-tx) / (float)A.N;      // translate(dst, src, -(tran - offset)(0), ...): RFLOAT arguments
            rr = (float)(-ty) / (float)A.N;
            if (A.nd.p) {     // CTF(ctf, pixelSize, voltage, defocusU * d, defocusV * d, ...): RFLOAT arguments (src/Optimiser.cpp:7173-7187)
                const double dfac = A.nd.at(l, A.drawD ? A.drawD[(size_t)l * A.mReco + m] : m, 0);
                dUs = (float)((double)A.ctfAttr[7 * l + 1] * dfac);
                dVs = (float)((double)A.ctfAttr[7 * l + 2] * dfac);
            }
            dx = -(rot.c0[0] * tx + rot.c1[0] * ty);   // insertDir(-rot3D * (tran - offset, 0))
            dy = -(rot.c0[1] * tx + rot.c1[1] * ty);
            dz = -(rot.c0[2] * tx + rot.c1[2] * ty);
        }
        dx = block_reduce_sum(dx, redd);
        dy = block_reduce_sum(dy, redd);
        dz = block_reduce_sum(dz, redd);
        if (tid == 0) {
            atomicAdd(&A.acc.O[3 * slot + 0], dx);
            atomicAdd(&A.acc.O[3 * slot + 1], dy);
            atomicAdd(&A.acc.O[3 * slot + 2], dz);
            atomicAdd(&A.acc.counter[slot], mcnt);
        }
        __syncthreads();
        if (tid < mcnt) {
            int rep = tid;
            if (MODE == 0)
                for (int j = 0; j < tid; ++j)
                    if (sQ[j][0] == sQ[tid][0] && sQ[j][1] == sQ[tid][1] && sQ[j][2] == sQ[tid][2] && sQ[j][3] == sQ[tid][3]) {
                        rep = j;
                        break;
                    }
            sRep[tid] = (unsigned short)rep;
        }
        __syncthreads();
        bool leader = false;
        if (tid < mcnt) {
            const int rep = sRep[tid];
            int pos = 0, size = 0, gidx = 0;
            for (int j = 0; j < mcnt; ++j) {
                const int rj = sRep[j];
                pos += (rj < rep) || (rj == rep && j < tid);
                size += rj == tid;
                gidx += (rj == j) && (j < tid);
            }
            ramp[mbase + pos] = make_float4(rc, rr, dUs, dVs);
            leader = rep == tid;
            if (leader) {
                gRot[nGrpTot + gidx] = rot;
                gEnd[nGrpTot + gidx] = mbase + pos + size;
            }
        }
        nGrpTot += __syncthreads_count(leader);
    }
    if (tid == 0) {
        hdr->nGrp = nGrpTot;
        hdr->nDraw = mTot;
        hdr->wgt = A.w ? A.w[l] : A.wAll;
        hdr->slot = slot;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// grid = (nImg, nSlab), block = 256
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(M2_THREADS, 2) insert_slab_kernel(const InsertSlabArgs S)
{
    const InsertArgs& A = S.a;
    __shared__ Rot2 sRot[M2_MAXD];
    __shared__ int sEnd[M2_MAXD];
    __shared__ float4 sRamp[M2_MAXD];
    __shared__ Seg sSeg[M2_MAXSEG];
    __shared__ SlabRec sRec[M2_ENT];
    __shared__ int sPre[M2_ENT + 1];
    __shared__ int sWarp[M2_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = S.order ? S.order[blockIdx.x] : blockIdx.x;
    const int zlo = S.zMin + (int)blockIdx.y * S.th, zhi = zlo + S.th;
    const unsigned char* base = S.prep + (size_t)l * prep_stride(S.maxD);
    const PrepHdr hdr = *reinterpret_cast<const PrepHdr*>(base);
    const Rot2* gRot = reinterpret_cast<const Rot2*>(base + sizeof(PrepHdr));
    const int* gEnd = reinterpret_cast<const int*>(gRot + S.maxD);
    const float4* ramp = reinterpret_cast<const float4*>(gEnd + S.maxD);
    const int nGrp = hdr.nGrp;
    if (nGrp == 0) return;

    // ---- does any slice of this image reach the slab at all?  max |z| over the disc of radius rmax is rmax |(R20, R21)|
    bool reach = false;
    for (int g = tid; g < nGrp; g += M2_THREADS) {
        const Rot2 r = gRot[g];
        sRot[g] = r;
        const double zmax = (double)S.rMaxPad * sqrt(r.c0[2] * r.c0[2] + r.c1[2] * r.c1[2]) + 2.0;
        reach |= (double)zlo <= zmax && (double)zhi >= -zmax;
    }
    if (!__syncthreads_or(reach)) return;
    for (int g = tid; g < nGrp; g += M2_THREADS) sEnd[g] = gEnd[g];
    for (int d = tid; d < hdr.nDraw; d += M2_THREADS) sRamp[d] = ramp[d];
    for (int k = tid; k < S.nSeg; k += M2_THREADS) sSeg[k] = S.seg[k];
    __syncthreads();

    const int img = A.imgIdx ? A.imgIdx[l] : l + A.imgBase;
    float4* __restrict__ acc = A.acc.p[hdr.slot];
    const int n = A.vdim, nColFT = n / 2 + 1;
    const int P = A.P;
    const float wgt = hdr.wgt;
    const float2* __restrict__ dat = A.dat + (size_t)img * P;
    const float* __restrict__ ctf = A.ctf + (size_t)img * P;
    const bool cSearch = A.nd.p != nullptr;
    CtfConst ck{};
    float cTheta = 0.f;
    if (cSearch) {
        const float* at = A.ctfAttr + 7 * (size_t)l;
        ck = ctf_const(at[0], at[4], at[5], at[6]);
        cTheta = at[3];
    }

    const int GB = max(1, min(nGrp, M2_ENT / S.nSeg));     // rotation groups per round
    for (int g0 = 0; g0 < nGrp; g0 += GB) {
        const int gb = min(GB, nGrp - g0);
        const int nEnt = gb * S.nSeg;
        // ---- intervals: thread t owns entries 4t .. 4t+3
        int cnt[4], tot = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = 4 * tid + u;
            cnt[u] = 0;
            if (e < nEnt) {
                const int gi = e / S.nSeg, k = e - gi * S.nSeg;
                SlabRec rec;
                cnt[u] = seg_intervals(sRot[g0 + gi], S.pf, sSeg[k], zlo, zhi, rec);
                sRec[e] = rec;
            }
            tot += cnt[u];
        }
        // ---- block-wide exclusive prefix sum of the counts
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) sWarp[warp] = inc;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w2 = 0; w2 < M2_THREADS / 32; ++w2) {
            const int v = sWarp[w2];
            if (w2 < warp) wbase += v;
            total += v;
        }
        int run = wbase + inc - tot;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = 4 * tid + u;
            if (e < nEnt) sPre[e] = run;
            run += cnt[u];
        }
        if (tid == 0) sPre[nEnt] = total;
        __syncthreads();

        // ---- the elements of this round
        for (int t = tid; t < total; t += M2_THREADS) {
            int lo = 0, hi = nEnt;                 // largest e with sPre[e] <= t
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sPre[mid] <= t) lo = mid; else hi = mid;
            }
            const int e = lo;
            const int p = slab_element(sRec[e], t - sPre[e]);
            const int gi = e / S.nSeg;
            const int g = g0 + gi;
            const Rot2 rot = sRot[g];
            const int4 c = A.pix[p];
            float x, y, z;
            slice_coord(rot, (double)c.x, (double)c.y, x, y, z);
            int x0, y0, z0;
            float xd, yd, zd;
            const bool conj = fold_floor(x, y, z, x0, y0, z0, xd, yd, zd);
            if (z0 < zlo || z0 >= zhi) continue;
            const float2 d = dat[p];
            float cf = ctf[p];
            float fx = 0.f, fy = 0.f, tv = 0.f;
            const int start = g ? sEnd[g - 1] : 0, end = sEnd[g];
            CtfPixel cp{};
            if (cSearch) cp = ctf_pixel(c.z, c.w, A.pixelSize, A.N, cTheta);
            for (int q = start; q < end; ++q) {
                const float4 rp = sRamp[q];
                const float ph = translate_phase(c.z, c.w, rp.x, rp.y);
                float s, co;
                sincosf(ph, &s, &co);
                // src * COMPLEX_POLAR(-ph) = d * (co - i s)
                const float vx = d.x * co + d.y * s;
                const float vy = d.y * co - d.x * s;
                if (cSearch) {
                    cf = ctf_eval(cp, ck, rp.z, rp.w);        // this draw's own CTF
                    tv += (cf * cf) * wgt;
                }
                fx += (vx * cf) * wgt;
                fy += (vy * cf) * wgt;
            }
            if (!cSearch) tv = (cf * cf) * wgt * (float)(end - start);
            if (conj) fy = -fy;
            float w8[8];
            tri_weights(xd, yd, zd, w8);
            int64_t off[4];
            row_offsets(y0, z0, n, nColFT, off);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float4* row = acc + off[cc] + x0;
                red_add_v4(row, fx * w8[2 * cc], fy * w8[2 * cc], tv * w8[2 * cc]);
                red_add_v4(row + 1, fx * w8[2 * cc + 1], fy * w8[2 * cc + 1], tv * w8[2 * cc + 1]);
            }
        }
        __syncthreads();
    }
}

}  // namespace thb
