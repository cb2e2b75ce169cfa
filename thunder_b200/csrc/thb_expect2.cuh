// thb_expect2.cuh - fused E kernel, local-search shape, with the projector volume staged through
// shared memory by the TMA engine (cp.async.bulk, one bulk copy per volume row of the box).
//
// One CTA per image.  256 threads = 8 warps: warp w serves rotation group (w & 3) - one rotation
// sample per lane - and pixel half (w >> 2).  The image is walked in 8x8-pixel tiles (the blocked
// pixel order of thb_api.cu keeps a tile contiguous).  Per tile:
//   (a) warps 4-7 build the pixel records of the tile in shared memory: the image pixel turned by the
//       conjugate phase ramp of each translation and pre-multiplied for the expanded likelihood,
//   (b) warps 0-3 classify every rotation against the tile: the cells its slice touches are bounded
//       by the image of the tile rectangle under the rotation; rotations whose cells lie within a
//       margin of the cloud's medoid rotation are "core", the union of their bounds is the box,
//   (c) every thread issues TMA bulk copies HBM/L2 -> shared memory, one per (y,z) row of the box,
//       completing on an mbarrier,
//   (d) while the box is in flight, the non-core rotations (wide-cloud tails, tiles that straddle the
//       Hermitian fold) are evaluated with pixels on the lanes and the 8-tap gather going to L1/L2,
//   (e) the core rotations gather from the shared-memory box: bank conflicts instead of L1 wavefronts.
//
// Likelihood (reference logDataVSPrior_m_huabin, src/Optimiser.cpp:9187-9213, with priAllP = traP * priRotP):
//   logL(r,t) = sum_i sig_i | dat_i - ctf_i tra_ti pri_ri |^2 ,   sig_i = -0.5 / sigma_i^2
// With D_ti = dat_i conj(tra_ti) (|tra| = 1) this is expanded as
//   logL(r,t) = sum_i sig_i |dat_i|^2  +  sum_i ( u_ti . pri_ri  +  g_i |pri_ri|^2 ),
//   u_ti = -2 sig_i ctf_i D_ti ,  g_i = sig_i ctf_i^2 ,
// which needs 2 FMAs per (sample, translation) instead of 5 operations and keeps the running sums
// small (the large constant first term is summed once per image, in double).
//
// Everything else follows the reference exactly: double-precision R*(pf*i, pf*j, 0) rounded to float
// (src/Projector.cpp:356-374), Hermitian fold before floor (include/Image/Volume.h:135-147), trilinear
// weights w[k][j][i] (include/Functions/Interpolation.h:187-200), wrap of negative y/z (Volume.h:567-575),
// weight accumulation against the maximum (src/Optimiser.cpp:1383-1402).
#pragma once
#include <cuda_runtime.h>
#include "thb_math.cuh"
#include "thb_types.cuh"

namespace thb {

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy wrappers (PTX ISA 8.x, sm_90+)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// global -> shared bulk copy by the TMA engine; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// gather from the pitched volume in HBM (reference Volume::getByInterpolationFT, Volume.cpp:314-338)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 gather_ft_pitched(const float2* __restrict__ vol, int n, int pitch, float x, float y, float z)
{
    int x0, y0, z0;
    float xd, yd, zd;
    const bool conj = fold_floor(x, y, z, x0, y0, z0, xd, yd, zd);
    float w[8];
    tri_weights(xd, yd, zd, w);
    int64_t off[4];
    row_offsets(y0, z0, n, pitch, off);
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

struct __align__(16) PixelRec {
    double a, b;        // pf*iCol, pf*iRow
    float g, pad;       // sig * ctf^2
    float2 u[E_TC];     // -2 sig ctf dat conj(tra_t)
};
static_assert(sizeof(PixelRec) == 96, "PixelRec must be 96 bytes");

struct BoxRange { int lo[3], hi[3]; };   // inclusive cell ranges (x, y, z) in post-fold signed coordinates

// cell range touched by the image of a tile rectangle (centre ca,cb ; half extents ha,hb ; padded units)
// under a rotation given by float copies of its first two columns, in the frame sgn (+1: unfolded,
// -1: folded = negated coordinates).  eps covers float rounding of this estimate and of the samples.
__device__ __forceinline__ void tile_cells(const float c0[3], const float c1[3], float ca, float cb, float ha, float hb,
                                           float sgn, BoxRange& r)
{
    const float eps = 0.02f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float c = sgn * (c0[k] * ca + c1[k] * cb);
        const float h = fabsf(c0[k]) * ha + fabsf(c1[k]) * hb + eps;
        r.lo[k] = (int)floorf(c - h);
        r.hi[k] = (int)floorf(c + h);
    }
}

constexpr int E2_THREADS = 256;
constexpr int E2_ROTS = 128;
constexpr int E2_TILE = 64;
constexpr int E2_BOX_ELEMS = 11776;             // float2 elements (92 KB)
constexpr int E2_HM_MAX = 10;
constexpr size_t E2_SMEM_BYTES = (size_t)E2_BOX_ELEMS * 8 + E2_TILE * sizeof(PixelRec) + E2_ROTS * sizeof(Rot2) +
                                 E2_ROTS * E_TC * sizeof(float);

__global__ void __launch_bounds__(E2_THREADS, 2) expect_local_tma_kernel(const ExpectArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* box = reinterpret_cast<float2*>(smem_raw);
    PixelRec* tile = reinterpret_cast<PixelRec*>(smem_raw + (size_t)E2_BOX_ELEMS * 8);
    Rot2* sRot = reinterpret_cast<Rot2*>(smem_raw + (size_t)E2_BOX_ELEMS * 8 + E2_TILE * sizeof(PixelRec));
    float* sAcc = reinterpret_cast<float*>(smem_raw + (size_t)E2_BOX_ELEMS * 8 + E2_TILE * sizeof(PixelRec) + E2_ROTS * sizeof(Rot2));
    // log-likelihood table [nR][nT]: aliases the box when the shape is a single pass (it is filled after the last
    // tile), else lives in the caller's scratch (passes re-stage the box)
    __shared__ __align__(8) uint64_t sBar;
    __shared__ float sRC[E_TC], sRR[E_TC];
    __shared__ int sRed[4][6];
    __shared__ unsigned char sCore[E2_ROTS];
    __shared__ int sOut[E2_ROTS];
    __shared__ int sNOut[2];
    __shared__ int sCentral;
    __shared__ float redf[E2_THREADS / 32];
    __shared__ double redd[E2_THREADS / 32];
    __shared__ float sMedD[4];
    __shared__ int sMedI[4];

    const int p = blockIdx.x;
    if (A.active && !A.active[p]) return;
    const int img = A.imgIdx ? A.imgIdx[p] : p + A.imgBase;
    const int slot = A.slotOfImg ? A.slotOfImg[img] : 0;
    const float2* __restrict__ vol = A.vols.p[slot];
    const int n = A.vdim, pitch = A.pitch, half = n / 2;
    const int P = A.P;
    const float2* __restrict__ dat = A.dat + (size_t)img * P;
    const float* __restrict__ ctf = A.ctf + (size_t)img * P;
    const float* __restrict__ sig = A.sig + (size_t)img * P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rloc = (warp & 3) * 32 + lane;   // rotation slot of this thread within a pass
    const int ph = warp >> 2;                  // pixel half
    const int nRT = A.nR * A.nT;
    float* sL = (A.nR <= E2_ROTS && A.nT <= E_TC) ? reinterpret_cast<float*>(smem_raw) : A.work + (size_t)p * nRT;

    if (tid == 0) {
        mbar_init(&sBar, 1);
        sNOut[0] = 0;
        sNOut[1] = 0;
    }
    uint32_t barParity = 0;
    int tileSeq = 0;
    double k0sum = 0.0;          // sum_i sig_i |dat_i|^2, accumulated by the record builders of the first pass
    __syncthreads();

    for (int rbase = 0; rbase < A.nR; rbase += E2_ROTS) {
        const int nRc = min(E2_ROTS, A.nR - rbase);
        const bool rvalid = rloc < nRc;
        // ---- rotations of this pass: matrices to shared memory, float quaternions for the medoid
        __syncthreads();
        float4* sQ = reinterpret_cast<float4*>(tile);     // 128 x 16 B = 2 KB, inside the record area (6 KB)
        if (ph == 0) {
            double q[4] = {1.0, 0.0, 0.0, 0.0};
            if (rvalid)
                for (int c = 0; c < 4; ++c) q[c] = A.quat.at(p, rbase + rloc, c);
            sRot[rloc] = quat_to_rot2(q);
            sQ[rloc] = make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]);
        }
        __syncthreads();
        const Rot2 rot = sRot[rloc];
        // medoid of the cloud: the member with the smallest summed chordal distance to the others
        if (ph == 0) {
            float s = 3.0e38f;
            if (rvalid) {
                const float4 me = sQ[rloc];
                s = 0.0f;
                for (int j = 0; j < nRc; ++j) {
                    const float4 o = sQ[j];
                    s += 1.0f - fabsf(me.x * o.x + me.y * o.y + me.z * o.z + me.w * o.w);
                }
            }
            int bi = rloc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
                const int b2 = __shfl_xor_sync(0xffffffffu, bi, o);
                if (s2 < s || (s2 == s && b2 < bi)) { s = s2; bi = b2; }
            }
            if (lane == 0) { sMedD[warp] = s; sMedI[warp] = bi; }
        }
        __syncthreads();
        if (tid == 0) {
            float s = sMedD[0];
            int bi = sMedI[0];
            for (int w2 = 1; w2 < 4; ++w2)
                if (sMedD[w2] < s) { s = sMedD[w2]; bi = sMedI[w2]; }
            sCentral = bi;
        }
        __syncthreads();
        float cc0[3], cc1[3], rc0[3], rc1[3];   // float copies: central rotation, own rotation
        {
            const Rot2 cr = sRot[sCentral];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                cc0[k] = (float)cr.c0[k]; cc1[k] = (float)cr.c1[k];
                rc0[k] = (float)rot.c0[k]; rc1[k] = (float)rot.c1[k];
            }
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
            for (int i = tid; i < E2_ROTS * E_TC; i += E2_THREADS) sAcc[i] = 0.0f;
            float acc[E_TC];
#pragma unroll
            for (int t = 0; t < E_TC; ++t) acc[t] = 0.0f;
            float nrm = 0.0f;
            const bool firstPass = (rbase == 0 && tbase == 0);

            for (int ti = 0; ti < A.nTiles; ++ti, ++tileSeq) {
                const TileDesc td = A.tiles[ti];
                const int cur = tileSeq & 1;
                __syncthreads();   // (A) previous tile finished: box, records, out-list are free
                if (ph == 1) {
                    // ---------------- (a) pixel records: 2 threads per pixel, translations split between them
                    const int k = (tid - 128) >> 1, sub = tid & 1;
                    if (k < td.count) {
                        const int i = td.start + k;
                        const int4 c = A.pix[i];
                        const float2 d = dat[i];
                        const float cf = ctf[i], sg = sig[i];
                        const float m2 = -2.0f * sg * cf;
                        PixelRec& rec = tile[k];
                        if (sub == 0) {
                            rec.a = (double)c.x;
                            rec.b = (double)c.y;
                            rec.g = sg * cf * cf;
                            rec.pad = 0.0f;
                            if (firstPass) k0sum += (double)(sg * (d.x * d.x + d.y * d.y));
                        }
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) {
                            if ((t & 1) != sub) continue;
                            const float phs = translate_phase(c.z, c.w, sRC[t], sRR[t]);
                            float s, co;
                            sincosf(phs, &s, &co);
                            // tra = (cos(-ph), sin(-ph)); dat * conj(tra) = dat * (co + i s)
                            rec.u[t] = make_float2(m2 * (d.x * co - d.y * s), m2 * (d.x * s + d.y * co));
                        }
                    }
                } else {
                    // ---------------- (b) classify the rotations against this tile
                    BoxRange cen, own;
                    const float cx = cc0[0] * td.ca + cc1[0] * td.cb;
                    const float sgn = cx >= 0.0f ? 1.0f : -1.0f;
                    tile_cells(cc0, cc1, td.ca, td.cb, td.ha, td.hb, sgn, cen);
                    tile_cells(rc0, rc1, td.ca, td.cb, td.ha, td.hb, sgn, own);
                    // largest margin (in cells) such that the central range grown by it still fits the box
                    int hm = -1;
                    if (cen.lo[0] >= 0) {
                        const int ex = cen.hi[0] - cen.lo[0] + 3, ey = cen.hi[1] - cen.lo[1] + 2, ez = cen.hi[2] - cen.lo[2] + 2;
                        for (int h = E2_HM_MAX; h >= 0; --h)
                            if ((ex + 2 * h) * (ey + 2 * h) * (ez + 2 * h) <= E2_BOX_ELEMS) { hm = h; break; }
                    }
                    bool core = rvalid && hm >= 0 && own.lo[0] >= 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        core = core && own.lo[k] >= cen.lo[k] - hm && own.hi[k] <= cen.hi[k] + hm;
                    // the cells must exist in the volume (garbage quaternions do not reach shared memory)
                    core = core && own.hi[0] + 1 <= half && own.lo[1] >= -half && own.hi[1] + 1 <= half && own.lo[2] >= -half &&
                           own.hi[2] + 1 <= half;
                    sCore[rloc] = core ? 1 : 0;
                    if (rvalid && !core) sOut[atomicAdd(&sNOut[cur], 1)] = rloc;
                    const int big = 1 << 28;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int lo = __reduce_min_sync(0xffffffffu, core ? own.lo[k] : big);
                        const int hi = __reduce_max_sync(0xffffffffu, core ? own.hi[k] : -big);
                        if (lane == 0) { sRed[warp][k] = lo; sRed[warp][3 + k] = hi; }
                    }
                }
                __syncthreads();   // (B) records, core flags, out-list, partial box bounds visible
                if (tid == 0) sNOut[cur ^ 1] = 0;
                int lo[3], hi[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    lo[k] = min(min(sRed[0][k], sRed[1][k]), min(sRed[2][k], sRed[3][k]));
                    hi[k] = max(max(sRed[0][3 + k], sRed[1][3 + k]), max(sRed[2][3 + k], sRed[3][3 + k]));
                }
                const bool haveBox = lo[0] <= hi[0];
                const int xloE = lo[0] & ~1;
                const int Lx = ((hi[0] + 2 - xloE) + 1) & ~1;     // cells lo..hi need taps lo..hi+1
                const int ny = hi[1] - lo[1] + 2, nz = hi[2] - lo[2] + 2;
                if (haveBox) {
                    // ---------------- (c) stage the box: one TMA bulk copy per (y,z) row
                    fence_proxy_async();
                    const int rows = ny * nz;
                    if (tid == 0) mbar_arrive_expect_tx(&sBar, (uint32_t)rows * (uint32_t)Lx * 8u);
                    for (int r = tid; r < rows; r += E2_THREADS) {
                        const int bz = r / ny, by = r - bz * ny;
                        const int ym = wrap_idx(lo[1] + by, n), zm = wrap_idx(lo[2] + bz, n);
                        tma_bulk_g2s(box + (size_t)r * Lx, vol + ((size_t)zm * n + ym) * pitch + xloE, (uint32_t)Lx * 8u, &sBar);
                    }
                }
                // ---------------- (d) non-core rotations: pixels on the lanes, gather from L1/L2
                {
                    const int nOut = sNOut[cur];
                    for (int it = warp; it < nOut; it += E2_THREADS / 32) {
                        const int rl = sOut[it];
                        const Rot2 ro = sRot[rl];
                        float v[E_TC];
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) v[t] = 0.0f;
                        for (int k = lane; k < td.count; k += 32) {
                            const PixelRec& rec = tile[k];
                            float x, y, z;
                            slice_coord(ro, rec.a, rec.b, x, y, z);
                            const float2 pr = gather_ft_pitched(vol, n, pitch, x, y, z);
                            const float m = rec.g * (pr.x * pr.x + pr.y * pr.y);
#pragma unroll
                            for (int t = 0; t < E_TC; ++t) v[t] += rec.u[t].x * pr.x + rec.u[t].y * pr.y + m;
                        }
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) {
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v[t] += __shfl_xor_sync(0xffffffffu, v[t], o);
                        }
                        if (lane == 0) {
#pragma unroll
                            for (int t = 0; t < E_TC; ++t) sAcc[rl * E_TC + t] += v[t];   // one warp per (rotation, tile)
                        }
                    }
                }
                // ---------------- (e) core rotations: gather from the staged box
                if (haveBox) {
                    mbar_wait(&sBar, barParity);
                    barParity ^= 1;
                    if (sCore[rloc]) {
                        const int sy = Lx, sz = Lx * ny;
#pragma unroll 2
                        for (int k = ph; k < td.count; k += 2) {
                            const PixelRec& rec = tile[k];
                            float x, y, z;
                            slice_coord(rot, rec.a, rec.b, x, y, z);
                            int x0, y0, z0;
                            float xd, yd, zd;
                            const bool conj = fold_floor(x, y, z, x0, y0, z0, xd, yd, zd);
                            float w[8];
                            tri_weights(xd, yd, zd, w);
                            const float2* b0 = box + ((z0 - lo[2]) * ny + (y0 - lo[1])) * Lx + (x0 - xloE);
                            const float2 v0 = b0[0], v1 = b0[1], v2 = b0[sy], v3 = b0[sy + 1];
                            const float2 v4 = b0[sz], v5 = b0[sz + 1], v6 = b0[sz + sy], v7 = b0[sz + sy + 1];
                            float re = 0.0f, im = 0.0f;
                            re += v0.x * w[0]; im += v0.y * w[0];
                            re += v1.x * w[1]; im += v1.y * w[1];
                            re += v2.x * w[2]; im += v2.y * w[2];
                            re += v3.x * w[3]; im += v3.y * w[3];
                            re += v4.x * w[4]; im += v4.y * w[4];
                            re += v5.x * w[5]; im += v5.y * w[5];
                            re += v6.x * w[6]; im += v6.y * w[6];
                            re += v7.x * w[7]; im += v7.y * w[7];
                            if (conj) im = -im;
                            nrm += rec.g * (re * re + im * im);
#pragma unroll
                            for (int t = 0; t < E_TC; ++t) acc[t] += rec.u[t].x * re + rec.u[t].y * im;
                        }
                    }
                }
            }
            // ---- end of the pass over the tiles: combine halves, fallback sums and the constant term
            __syncthreads();
            if (firstPass) {
                // block sum of k0sum (double)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) k0sum += __shfl_xor_sync(0xffffffffu, k0sum, o);
                if (lane == 0) redd[warp] = k0sum;
                __syncthreads();
                double s = 0.0;
                for (int w2 = 0; w2 < E2_THREADS / 32; ++w2) s += redd[w2];
                k0sum = s;
                __syncthreads();
            }
            // halves: ph == 1 parks its sums in the record area, ph == 0 adds everything up
            float* park = reinterpret_cast<float*>(tile);     // 128 x 10 floats = 5 KB <= 6 KB
            if (ph == 1) {
#pragma unroll
                for (int t = 0; t < E_TC; ++t) park[rloc * (E_TC + 1) + t] = acc[t];
                park[rloc * (E_TC + 1) + E_TC] = nrm;
            }
            __syncthreads();
            if (ph == 0 && rvalid) {
                const double nn = (double)nrm + (double)park[rloc * (E_TC + 1) + E_TC];
#pragma unroll
                for (int t = 0; t < E_TC; ++t)
                    if (tbase + t < A.nT)
                        sL[(size_t)(rbase + rloc) * A.nT + tbase + t] =
                            (float)(k0sum + nn + (double)acc[t] + (double)park[rloc * (E_TC + 1) + t] + (double)sAcc[rloc * E_TC + t]);
            }
        }
    }
    __syncthreads();

    // ---------------- epilogue: baseline, weights, marginals (Optimiser.cpp:1383-1402) ----------
    float m = -INFINITY;
    for (int i = tid; i < nRT; i += E2_THREADS) m = fmaxf(m, sL[i]);
    m = block_reduce_max(m, redf);
    if (A.logL)
        for (int i = tid; i < nRT; i += E2_THREADS) A.logL[(size_t)p * nRT + i] = sL[i];
    __syncthreads();
    for (int i = tid; i < nRT; i += E2_THREADS) sL[i] = expf(sL[i] - m);
    __syncthreads();
    double uc = 0.0;
    for (int r = tid; r < A.nR; r += E2_THREADS) {
        float s = 0.0f;
        for (int t = 0; t < A.nT; ++t) s = (float)((double)s + (double)sL[r * A.nT + t] * A.wT.at(p, t, 0));
        if (A.uR) A.uR[(size_t)p * A.nR + r] = s;
        uc += (double)s * A.wR.at(p, r, 0);
    }
    for (int t = tid; t < A.nT; t += E2_THREADS) {
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

}  // namespace thb
