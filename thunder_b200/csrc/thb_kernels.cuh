// thb_kernels.cuh - sm_100a kernels of the Optimiser hot path (expectation / insertion).
//
// Layouts in HBM
//   projector volume   float2 [(n/2+1) x n x n]   FFTW half-complex, x fastest (as uploaded)
//   E stack            float2 dat[nImg][P], float ctf[nImg][P], float sigRcp[nImg][P]
//   M stack            float2 dat[nImg][P], float ctf[nImg][P]
//   pixel lists        int4 {a = pf*iCol, b = pf*iRow, iCol, iRow}
//   accumulator        float4 {F.re, F.im, T, 0} [(m/2+1) x m x m]  - one 16-byte vector red per corner
#pragma once
#include <cuda_runtime.h>
#include "thb_math.cuh"
#include "thb_types.cuh"
#include "thb_pack.cuh"

namespace thb {




// ------------------------------------------------------------------------------------------------
// trilinear gather of one complex sample (reference Volume::getByInterpolationFT, Volume.cpp:314-338)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 gather_ft(const float2* __restrict__ vol, int n, int nColFT /* row pitch */, float x, float y,
                                            float z)
{
    int x0, y0, z0;
    float xd, yd, zd;
    const bool conj = fold_floor(x, y, z, x0, y0, z0, xd, yd, zd);
    float w[8];
    tri_weights(xd, yd, zd, w);
    int64_t off[4];
    row_offsets(y0, z0, n, nColFT, off);
    float2 v[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float2* row = vol + off[c] + x0;
        v[2 * c] = __ldg(row);
        v[2 * c + 1] = __ldg(row + 1);
    }
    float re = 0.0f, im = 0.0f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        re += v[c].x * w[c];
        im += v[c].y * w[c];
    }
    return make_float2(re, conj ? -im : im);
}

// ------------------------------------------------------------------------------------------------
// Projector::project for many rotations (src/Projector.cpp:356-374); lanes = pixels
// ------------------------------------------------------------------------------------------------
__global__ void project_kernel(const float2* __restrict__ vol, int n, int pitch, const int4* __restrict__ pix,
                               const int* __restrict__ perm, int P, const double* __restrict__ quat,
                               float2* __restrict__ dst, int mode2D)
{
    const int r = blockIdx.y;
    double q[4] = {1.0, 0.0, 0.0, 0.0};
    const int nc = mode2D ? 2 : 4;          // MODE_2D rotation arrays are [..][2] = (cos, sin)
    for (int c = 0; c < nc; ++c) q[c] = quat[nc * r + c];
    const Rot2 rot = make_rot2(q, mode2D);
    const int nColFT = pitch;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const int4 px = pix[i];
        float x, y, z;
        slice_coord(rot, (double)px.x, (double)px.y, x, y, z);
        dst[(size_t)r * P + perm[i]] = gather_ft(vol, n, nColFT, x, y, z);   // caller's pixel order
    }
}

// ------------------------------------------------------------------------------------------------
// Fused E kernel, local-search shape.  One CTA per image, one rotation sample per thread,
// TC translations in registers; the image tile (already multiplied by the conjugate phase ramp of
// each translation) is staged in shared memory and broadcast to all rotations.
//
//   logL(r,t) = sum_i | dat_i - ctf_i * tra_{t,i} * pri_{r,i} |^2 * sigRcp_i
//             = sum_i | dat_i * conj(tra_{t,i}) - ctf_i * pri_{r,i} |^2 * sigRcp_i      (|tra| = 1)
//
// reference: translate (ImageFunctions.cpp:233-252), Projector::project (Projector.cpp:356-374),
// logDataVSPrior_m_huabin (Optimiser.cpp:9187-9213), weight accumulation (Optimiser.cpp:1383-1402).
// ------------------------------------------------------------------------------------------------


__device__ __forceinline__ float block_reduce_max(float v, float* red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
    return r;
}

__device__ __forceinline__ double block_reduce_sum(double v, double* red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = red[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) r += red[i];
    return r;
}

__global__ void __launch_bounds__(E_THREADS, 4) expect_local_kernel(const ExpectArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PixelE* tile = reinterpret_cast<PixelE*>(smem_raw);
    float* sL = reinterpret_cast<float*>(smem_raw + sizeof(PixelE) * E_TILE);   // [nR][nT]
    __shared__ float sRC[E_TC], sRR[E_TC];
    __shared__ float redf[E_THREADS / 32];
    __shared__ double redd[E_THREADS / 32];

    const int p = blockIdx.x;
    if (A.active && !A.active[p]) return;
    const int img = A.imgIdx ? A.imgIdx[p] : p + A.imgBase;
    const int slot = A.slotAll >= 0 ? A.slotAll : (A.slotOfImg ? A.slotOfImg[img] : 0);
    const float2* __restrict__ vol = A.vols.p[slot];
    const int n = A.vdim, nColFT = A.pitch;
    const int P = A.P;
    const float2* __restrict__ dat = A.dat + (size_t)img * P;
    const float* __restrict__ ctf = A.ctf + (size_t)img * P;
    const float* __restrict__ sig = A.sig + (size_t)img * P;
    const int tid = threadIdx.x;

    for (int rbase = 0; rbase < A.nR; rbase += E_THREADS) {
        const int r = rbase + tid;
        const bool rvalid = r < A.nR;
        Rot2 rot;
        {
            double q[4] = {1.0, 0.0, 0.0, 0.0};
            if (rvalid)
                for (int c = 0; c < (A.mode2D ? 2 : 4); ++c) q[c] = A.quat.at(p, r, c);
            rot = make_rot2(q, A.mode2D);
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
            float acc[E_TC];
#pragma unroll
            for (int t = 0; t < E_TC; ++t) acc[t] = 0.0f;

            for (int tile0 = 0; tile0 < P; tile0 += E_TILE) {
                __syncthreads();   // previous tile fully consumed (also orders sRC/sRR writes)
                {
                    const int i = tile0 + tid;
                    PixelE px;
                    if (i < P) {
                        const int4 c = A.pix[i];
                        const float2 d = dat[i];
                        px.a = (double)c.x;
                        px.b = (double)c.y;
                        px.ctf = ctf[i];
                        px.sig = sig[i];
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) {
                            const float ph = translate_phase(c.z, c.w, sRC[t], sRR[t]);
                            float s, co;
                            sincosf(ph, &s, &co);
                            // tra = (cos(-ph), sin(-ph)); dat * conj(tra) = dat * (co + i s)
                            px.d[t] = make_float2(d.x * co - d.y * s, d.x * s + d.y * co);
                        }
                    } else {
                        px.a = 0.0; px.b = 0.0; px.ctf = 0.0f; px.sig = 0.0f;
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) px.d[t] = make_float2(0.0f, 0.0f);
                    }
                    tile[tid] = px;
                }
                __syncthreads();
                if (rvalid) {
                    const int cnt = min(E_TILE, P - tile0);
#pragma unroll 2
                    for (int k = 0; k < cnt; ++k) {
                        const PixelE& px = tile[k];
                        float x, y, z;
                        slice_coord(rot, px.a, px.b, x, y, z);
                        const float2 pr = gather_ft(vol, n, nColFT, x, y, z);
                        const float c = px.ctf, sg = px.sig;
                        const float qx = c * pr.x, qy = c * pr.y;
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) {
                            const float ex = px.d[t].x - qx;
                            const float ey = px.d[t].y - qy;
                            acc[t] += (ex * ex + ey * ey) * sg;
                        }
                    }
                }
            }
            if (rvalid) {
#pragma unroll
                for (int t = 0; t < E_TC; ++t)
                    if (tbase + t < A.nT) sL[(size_t)r * A.nT + tbase + t] = acc[t];
            }
        }
    }
    __syncthreads();

    // ---------------- epilogue: baseline, weights, marginals (Optimiser.cpp:1383-1402) ----------
    const int nRT = A.nR * A.nT;
    float m = -INFINITY;
    for (int i = tid; i < nRT; i += E_THREADS) m = fmaxf(m, sL[i]);
    m = block_reduce_max(m, redf);
    if (A.logL)
        for (int i = tid; i < nRT; i += E_THREADS) A.logL[(size_t)p * nRT + i] = sL[i];
    __syncthreads();
    for (int i = tid; i < nRT; i += E_THREADS) sL[i] = expf(sL[i] - m);
    __syncthreads();
    double uc = 0.0;
    for (int r = tid; r < A.nR; r += E_THREADS) {
        float s = 0.0f;
        for (int t = 0; t < A.nT; ++t) s = (float)((double)s + (double)sL[r * A.nT + t] * A.wT.at(p, t, 0));
        if (A.uR) A.uR[(size_t)p * A.nR + r] = s;
        uc += (double)s * A.wR.at(p, r, 0);
    }
    for (int t = tid; t < A.nT; t += E_THREADS) {
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

// ------------------------------------------------------------------------------------------------
// Fused M kernel: translate + CTF + weight + trilinear scatter of F and T, + insertDir.
// grid = (nImg, nSplit); lanes = pixels; the mReco draws of the image loop inside.
// reference: Optimiser::reconstructRef insert loop (Optimiser.cpp:7036-7241), translate
// (ImageFunctions.cpp:471-492), Reconstructor::insertP (Reconstructor.cpp:782-863),
// Volume::addFT (Volume.cpp:340-375, 565-712), insertDir (Reconstructor.cpp:407-422).
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void red_add_v4(float4* p, float a, float b, float c)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(0.0f)
                 : "memory");
}


// MODE 0 (default): draws of one image that share the SAME rotation (the support of a resampled particle
// filter holds many exact duplicates) are merged before the scatter: the translated image values of the
// group are summed and scattered once, T once with the group's multiplicity.  Same sums as draw-by-draw
// insertion up to fp32 rounding order.  MODE 2: draw-by-draw (A/B measurement).
template <int MODE>
__global__ void __launch_bounds__(M_THREADS) insert_kernel(const InsertArgs A)
{
    __shared__ Rot2 sRot[M_MAXRECO];
    __shared__ float sRC[M_MAXRECO], sRR[M_MAXRECO];
    __shared__ double sQ[M_MAXRECO][4];
    __shared__ unsigned short sRep[M_MAXRECO], sOrder[M_MAXRECO], sGrpEnd[M_MAXRECO], sCls[M_MAXRECO];
    __shared__ double redd[M_THREADS / 32];

    const int l = blockIdx.x;
    const int img = A.imgIdx ? A.imgIdx[l] : l + A.imgBase;
    const int slot = A.slotOfImg ? A.slotOfImg[img] : 0;
    float4* __restrict__ acc = A.acc.p[slot];
    const int n = A.vdim, nColFT = n / 2 + 1;
    const int P = A.P;
    const float wgt = A.w ? A.w[l] : A.wAll;
    const double ox = A.offS ? A.offS[2 * l] : 0.0, oy = A.offS ? A.offS[2 * l + 1] : 0.0;
    const int tid = threadIdx.x;

    // 3D classification (src/Optimiser.cpp:6862-6950): the caller passes max-count draws per image and the number of them
    // that belong to this class; the rest of the image's rows are not read
    const int mTot = A.drawCount ? max(0, min(A.mReco, A.drawCount[l])) : A.mReco;
    for (int mbase = 0; mbase < mTot; mbase += M_MAXRECO) {
        const int mcnt = min(M_MAXRECO, mTot - mbase);
        __syncthreads();
        double dx = 0.0, dy = 0.0, dz = 0.0;
        if (tid < mcnt) {
            const int m = mbase + tid;
            const long long sr = A.drawR ? A.drawR[(size_t)l * A.mReco + m] : m;
            const long long st = A.drawT ? A.drawT[(size_t)l * A.mReco + m] : m;
            double q[4] = {1.0, 0.0, 0.0, 0.0};
            for (int c = 0; c < (A.mode2D ? 2 : 4); ++c) q[c] = A.nr.at(l, sr, c);
            // MODE_2D with several classes: the draw's class (InsertI2D's nC) picks the accumulator and keeps draws of
            // different classes in different merge groups (the class rides in the unused third component of the key)
            const int cls = A.drawC ? A.drawC[(size_t)l * A.mReco + m] : slot;
            if (A.drawC) q[2] = (double)cls;
            sCls[tid] = (unsigned short)cls;
            for (int c = 0; c < 4; ++c) sQ[tid][c] = q[c];
            if (A.drawC) q[2] = 0.0;
            const Rot2 rot = make_rot2(q, A.mode2D);
            sRot[tid] = rot;
            const double tx = A.nt.at(l, st, 0) - ox, ty = A.nt.at(l, st, 1) - oy;
            // translate(dst, src, -(tran - offset)(0), -(tran - offset)(1), ...): RFLOAT arguments
            sRC[tid] = (float)(-tx) / (float)A.N;
            sRR[tid] = (float)(-ty) / (float)A.N;
            // insertDir(-rot3D * (tran - offset, 0))
            dx = -(rot.c0[0] * tx + rot.c1[0] * ty);
            dy = -(rot.c0[1] * tx + rot.c1[1] * ty);
            dz = -(rot.c0[2] * tx + rot.c1[2] * ty);
        }
        if (blockIdx.y == 0 && A.drawC) {
            if (tid < mcnt) {
                const int cls = sCls[tid];
                atomicAdd(&A.acc.O[3 * cls + 0], dx);
                atomicAdd(&A.acc.O[3 * cls + 1], dy);
                atomicAdd(&A.acc.O[3 * cls + 2], dz);
                atomicAdd(&A.acc.counter[cls], 1);
            }
        } else if (blockIdx.y == 0) {
            dx = block_reduce_sum(dx, redd);
            dy = block_reduce_sum(dy, redd);
            dz = block_reduce_sum(dz, redd);
            if (tid == 0) {
                atomicAdd(&A.acc.O[3 * slot + 0], dx);
                atomicAdd(&A.acc.O[3 * slot + 1], dy);
                atomicAdd(&A.acc.O[3 * slot + 2], dz);
                atomicAdd(&A.acc.counter[slot], mcnt);
            }
        }
        __syncthreads();
        // ---- group the draws by rotation: representative = first draw with bit-identical quaternion
        if (tid < mcnt) {
            int rep = tid;
            if (MODE == 0) {
                for (int j = 0; j < tid; ++j)
                    if (sQ[j][0] == sQ[tid][0] && sQ[j][1] == sQ[tid][1] && sQ[j][2] == sQ[tid][2] && sQ[j][3] == sQ[tid][3]) {
                        rep = j;
                        break;
                    }
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
            sOrder[pos] = (unsigned short)tid;
            leader = rep == tid;
            if (leader) sGrpEnd[gidx] = (unsigned short)(pos + size);
        }
        const int nGrp = __syncthreads_count(leader);

        // M_KP pixels per thread in registers; per rotation group the rotation and the draws' phase ramps are read from
        // shared memory ONCE for all of them (the reduction traffic owns the memory-instruction queue of the SM)
        for (int i0 = blockIdx.y * M_THREADS * M_KP; i0 < P; i0 += gridDim.y * M_THREADS * M_KP) {
            int4 c[M_KP];
            float2 d[M_KP];
            float cf[M_KP];
#pragma unroll
            for (int k = 0; k < M_KP; ++k) {
                const int i = i0 + k * M_THREADS + tid;
                const bool ok = i < P;
                c[k] = ok ? A.pix[i] : make_int4(0, 0, 0, 0);
                d[k] = ok ? A.dat[(size_t)img * P + i] : make_float2(0.f, 0.f);
                cf[k] = ok ? A.ctf[(size_t)img * P + i] : 0.f;
            }
            int start = 0;
            for (int g = 0; g < nGrp; ++g) {
                const int end = sGrpEnd[g];
                float fx[M_KP], fy[M_KP];
#pragma unroll
                for (int k = 0; k < M_KP; ++k) fx[k] = fy[k] = 0.0f;
                for (int q = start; q < end; ++q) {
                    const int m = sOrder[q];
                    const float rc = sRC[m], rr = sRR[m];
#pragma unroll
                    for (int k = 0; k < M_KP; ++k) {
                        const float ph = translate_phase(c[k].z, c[k].w, rc, rr);
                        float s, co;
                        sincosf(ph, &s, &co);
                        // src * COMPLEX_POLAR(-ph) = d * (co - i s)
                        const float vx = d[k].x * co + d[k].y * s;
                        const float vy = d[k].y * co - d[k].x * s;
                        fx[k] += (vx * cf[k]) * wgt;
                        fy[k] += (vy * cf[k]) * wgt;
                    }
                }
                const float mult = (float)(end - start);
                const Rot2 rot = sRot[sOrder[start]];
                float4* __restrict__ accg = A.drawC ? A.acc.p[sCls[sOrder[start]]] : acc;
                start = end;
#pragma unroll
                for (int k = 0; k < M_KP; ++k) {
                    if (i0 + k * M_THREADS + tid >= P) continue;
                    const float tv = (cf[k] * cf[k]) * wgt * mult;
                    float x, y, z;
                    slice_coord(rot, (double)c[k].x, (double)c[k].y, x, y, z);
                    int x0, y0, z0;
                    float xd, yd, zd;
                    float gy = fy[k];
                    if (fold_floor(x, y, z, x0, y0, z0, xd, yd, zd)) gy = -gy;
                    float w8[8];
                    tri_weights(xd, yd, zd, w8);
                    int64_t off[4];
                    row_offsets(y0, z0, n, nColFT, off);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (A.mode2D && cc == 2) break;          // z0 = 0, zd = 0: plane 1 only ever receives zeros
                        float4* row = accg + off[cc] + x0;
                        red_add_v4(row, fx[k] * w8[2 * cc], gy * w8[2 * cc], tv * w8[2 * cc]);
                        red_add_v4(row + 1, fx[k] * w8[2 * cc + 1], gy * w8[2 * cc + 1], tv * w8[2 * cc + 1]);
                    }
                }
            }
        }
    }
}

// de-interleave the accumulator: F complex64, T real fp32, optional 1/T[0] normalisation
__global__ void unpack_acc_kernel(const float4* __restrict__ acc, size_t nVox, float2* __restrict__ F,
                                  float* __restrict__ T, int normalise)
{
    const float sf = normalise ? 1.0f / acc[0].z : 1.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nVox; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = acc[i];
        if (F) F[i] = make_float2(v.x * sf, v.y * sf);
        if (T) T[i] = v.z * sf;
    }
}

// pixel list in the device (blocked) order: out[i] describes the caller's pixel perm[i]
__global__ void make_pix_kernel(const int* __restrict__ a, const int* __restrict__ b, const int* __restrict__ perm, int P,
                                int pf, int padded, int4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int s = perm[i];
    if (padded)
        out[i] = make_int4(a[s], b[s], a[s] / pf, b[s] / pf);
    else
        out[i] = make_int4(a[s] * pf, b[s] * pf, a[s], b[s]);
}

// resident stack = caller's image-major packed arrays with the pixels of each image permuted into
// the blocked order: dst[l][i] = src[l][perm[i]]
__global__ void permute_stack_kernel(const float2* __restrict__ sdat, const float* __restrict__ sctf,
                                     const float* __restrict__ ssig, const int* __restrict__ perm, int P, int nImg,
                                     float2* __restrict__ ddat, float* __restrict__ dctf, float* __restrict__ dsig)
{
    const int l = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const size_t s = (size_t)l * P + perm[i], d = (size_t)l * P + i;
        ddat[d] = sdat[s];
        dctf[d] = sctf[s];
        if (ssig) dsig[d] = ssig[s];
    }
}

// read a resident stack back in the caller's pixel order: dst[l][perm[i]] = src[l][i]
__global__ void unpermute_stack_kernel(const float2* __restrict__ sdat, const float* __restrict__ sctf,
                                       const float* __restrict__ ssig, const int* __restrict__ perm, int P,
                                       float2* __restrict__ ddat, float* __restrict__ dctf, float* __restrict__ dsig)
{
    const int l = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const size_t s = (size_t)l * P + i, d = (size_t)l * P + perm[i];
        ddat[d] = sdat[s];
        dctf[d] = sctf[s];
        if (ssig && dsig) dsig[d] = ssig[s];
    }
}

}  // namespace thb
