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
//       margin of the cloud's medoid rotation are "core".  The margin is the largest one whose staged
//       region fits the box: every candidate margin is evaluated in parallel (one warp each).
//   (c) the staged region is the bounding box of the core cells CUT BY THE SLAB around the medoid slice
//       plane (the slices of a cloud fill a thin oblique slab, not its bounding box): each (y,z) row
//       of the box keeps only the x-interval inside the slab, rows are packed back to back (block
//       prefix sum) and copied HBM/L2 -> shared memory by the TMA engine, one bulk copy per row,
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

#ifndef THB_E2_TIMERS
#define THB_E2_TIMERS 1
#endif
#define E2_TICK(slot) do { if (THB_E2_TIMERS && A.stats && tid == 0) { const long long now_ = clock64(); tacc[slot] += now_ - tlast; tlast = now_; } } while (0)

constexpr int E2_THREADS = 256;
constexpr int E2_ROTS = 128;
constexpr int E2_TILE = 128;                    // max pixels per tile
constexpr int E2_BOX_ELEMS = 9984;              // float2 elements (78 KB)
constexpr int E2_MAXROWS = 1536;                // (y,z) rows of a box
constexpr int E2_HM_MAX = 10;                   // candidate margins 0..E2_HM_MAX
constexpr int E2_NCAND = E2_HM_MAX + 1;

struct __align__(16) RotClass {                 // per rotation, per tile
    int lo[3], hi[3];
    int need;                                   // margin this rotation needs around the medoid's cell range
    float dist;                                 // its largest distance from the medoid slice plane
};
struct __align__(16) Cand {                     // per candidate margin
    int lo[3], hi[3];
    float dist;
    float vol;                                  // estimated staged elements
};

constexpr size_t E2_OFF_TILE = (size_t)E2_BOX_ELEMS * 8;
constexpr size_t E2_OFF_ROT = E2_OFF_TILE + E2_TILE * sizeof(PixelRec);
constexpr size_t E2_OFF_ACC = E2_OFF_ROT + E2_ROTS * sizeof(Rot2);
constexpr size_t E2_OFF_BIAS = E2_OFF_ACC + E2_ROTS * E_TC * sizeof(float);
constexpr size_t E2_OFF_CLS = E2_OFF_BIAS + E2_MAXROWS * sizeof(int);
constexpr size_t E2_SMEM_BYTES = E2_OFF_CLS + E2_ROTS * sizeof(RotClass);

// sum of 9 values over the 32 lanes of a warp: on return lane l holds the total of value (l >> 2) & 7 in
// `out` (values 0..7) and every lane holds the total of value 8 in `out8`
__device__ __forceinline__ void warp_sum9(const float v[E_TC], int lane, float& out, float& out8)
{
    float w[4], x[2];
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b16 ? v[i] : v[i + 4];
        const float keep = b16 ? v[i + 4] : v[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b8 ? w[i] : w[i + 2];
        const float keep = b8 ? w[i + 2] : w[i];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const float send = b4 ? x[0] : x[1];
        const float keep = b4 ? x[1] : x[0];
        out = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    out += __shfl_xor_sync(0xffffffffu, out, 2);
    out += __shfl_xor_sync(0xffffffffu, out, 1);
    float e = v[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    out8 = e;
}

__global__ void __launch_bounds__(E2_THREADS, 2) expect_local_tma_kernel(const ExpectArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* box = reinterpret_cast<float2*>(smem_raw);
    PixelRec* tile = reinterpret_cast<PixelRec*>(smem_raw + E2_OFF_TILE);
    Rot2* sRot = reinterpret_cast<Rot2*>(smem_raw + E2_OFF_ROT);
    float* sAcc = reinterpret_cast<float*>(smem_raw + E2_OFF_ACC);
    int* sBias = reinterpret_cast<int*>(smem_raw + E2_OFF_BIAS);
    RotClass* sCls = reinterpret_cast<RotClass*>(smem_raw + E2_OFF_CLS);
    __shared__ __align__(8) uint64_t sBar;
    __shared__ float sRC[E_TC], sRR[E_TC];
    __shared__ Cand sCand[E2_NCAND];
    __shared__ int sWarpTot[E2_THREADS / 32];
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
    // log-likelihood table [nR][nT]: aliases the box when the shape is a single pass (it is filled after the last
    // tile), else lives in the caller's scratch (passes re-stage the box)
    float* sL = (A.nR <= E2_ROTS && A.nT <= E_TC) ? reinterpret_cast<float*>(smem_raw) : A.work + (size_t)p * nRT;

    if (tid == 0) {
        mbar_init(&sBar, 1);
        sNOut[0] = 0;
        sNOut[1] = 0;
    }
    uint32_t barParity = 0;
    int tileSeq = 0;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();   // per-phase cycles of this CTA (thread 0)
    double k0sum = 0.0;          // sum_i sig_i |dat_i|^2, accumulated by the record builders of the first pass
    __syncthreads();

    for (int rbase = 0; rbase < A.nR; rbase += E2_ROTS) {
        const int nRc = min(E2_ROTS, A.nR - rbase);
        const bool rvalid = rloc < nRc;
        // ---- rotations of this pass: matrices to shared memory, float quaternions for the medoid
        __syncthreads();
        float4* sQ = reinterpret_cast<float4*>(tile);     // 128 x 16 B = 2 KB, inside the record area
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
        // float copies: medoid ("central") rotation, its slice-plane normal, own rotation and the
        // coefficients of its distance from the medoid plane: n . (c0 a + c1 b) = al a + be b
        float cc0[3], cc1[3], rc0[3], rc1[3], nc[3], al, be;
        {
            const Rot2 cr = sRot[sCentral];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                cc0[k] = (float)cr.c0[k]; cc1[k] = (float)cr.c1[k];
                rc0[k] = (float)rot.c0[k]; rc1[k] = (float)rot.c1[k];
            }
            nc[0] = cc0[1] * cc1[2] - cc0[2] * cc1[1];
            nc[1] = cc0[2] * cc1[0] - cc0[0] * cc1[2];
            nc[2] = cc0[0] * cc1[1] - cc0[1] * cc1[0];
            al = nc[0] * rc0[0] + nc[1] * rc0[1] + nc[2] * rc0[2];
            be = nc[0] * rc1[0] + nc[1] * rc1[1] + nc[2] * rc1[2];
        }
        const float kappa = fabsf(nc[0]) + fabsf(nc[1]) + fabsf(nc[2]);

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
                __syncthreads();   // (A) previous tile finished: box, records, tables are free
                E2_TICK(0);
                {
                    // ---------------- (a) pixel records: 2 threads per pixel, translations split between them
                    const int k = tid >> 1, sub = tid & 1;
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
                }
                // frame of the tile: the side of the Hermitian fold the medoid puts its centre on
                const float sgn = (cc0[0] * td.ca + cc1[0] * td.cb) >= 0.0f ? 1.0f : -1.0f;
                BoxRange cen;
                tile_cells(cc0, cc1, td.ca, td.cb, td.ha, td.hb, sgn, cen);
                if (ph == 0) {
                    // ---------------- (b) classify the rotations against this tile
                    BoxRange own;
                    tile_cells(rc0, rc1, td.ca, td.cb, td.ha, td.hb, sgn, own);
                    int need = 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) need = max(need, max(cen.lo[k] - own.lo[k], own.hi[k] - cen.hi[k]));
                    // the slice must stay on one side of the fold, and its cells must exist in the volume
                    // (garbage quaternions do not reach shared memory)
                    const bool ok = rvalid && cen.lo[0] >= 0 && own.lo[0] >= 0 && own.hi[0] + 1 <= half && own.lo[1] >= -half &&
                                    own.hi[1] + 1 <= half && own.lo[2] >= -half && own.hi[2] + 1 <= half;
                    RotClass rc;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { rc.lo[k] = own.lo[k]; rc.hi[k] = own.hi[k]; }
                    rc.need = ok ? need : (1 << 20);
                    rc.dist = fabsf(al * td.ca + be * td.cb) + fabsf(al) * td.ha + fabsf(be) * td.hb;
                    sCls[rloc] = rc;
                }
                __syncthreads();   // (B) records and per-rotation classes visible
                E2_TICK(1);
                // ---------------- candidate margins, one warp each: union of the cell ranges and slab thickness of the
                // rotations that need at most that margin, and the estimated size of the staged region
                for (int h = warp; h < E2_NCAND; h += E2_THREADS / 32) {
                    const int big = 1 << 28;
                    int lo[3] = {big, big, big}, hi[3] = {-big, -big, -big};
                    float dist = -1.0f;
                    for (int r = lane; r < E2_ROTS; r += 32) {
                        const RotClass rc = sCls[r];
                        if (rc.need <= h) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], rc.lo[k]); hi[k] = max(hi[k], rc.hi[k]); }
                            dist = fmaxf(dist, rc.dist);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
                        hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) dist = fmaxf(dist, __shfl_xor_sync(0xffffffffu, dist, o));
                    if (lane == 0) {
                        Cand c;
                        float vol = 3.0e38f;
                        if (lo[0] <= hi[0]) {
                            const float ex = (float)(hi[0] - lo[0] + 4), ey = (float)(hi[1] - lo[1] + 2), ez = (float)(hi[2] - lo[2] + 2);
                            // elements of (bounding box ^ slab): the slab passes through the middle of the box, where the
                            // cross-section is largest; its projection on the face normal to axis k cannot exceed that face,
                            // so section <= face_k / |n_k| for every k.  + about 2 elements of alignment / rounding per row.
                            const float thick = 2.0f * (dist + kappa + 0.05f);
                            float section = 3.0e38f;
                            if (fabsf(nc[0]) > 1e-3f) section = fminf(section, ey * ez / fabsf(nc[0]));
                            if (fabsf(nc[1]) > 1e-3f) section = fminf(section, ex * ez / fabsf(nc[1]));
                            if (fabsf(nc[2]) > 1e-3f) section = fminf(section, ex * ey / fabsf(nc[2]));
                            float rowsNE = ey * ez;       // rows the slab reaches
                            const float reach = thick + fabsf(nc[0]) * ex;
                            if (fabsf(nc[1]) > 1e-3f) rowsNE = fminf(rowsNE, ez * (reach / fabsf(nc[1]) + 2.0f));
                            if (fabsf(nc[2]) > 1e-3f) rowsNE = fminf(rowsNE, ey * (reach / fabsf(nc[2]) + 2.0f));
                            vol = fminf(ex * ey * ez, section * thick) + 2.0f * rowsNE;
                            if (ey * ez > (float)E2_MAXROWS) vol = 3.0e38f;
                        }
#pragma unroll
                        for (int k = 0; k < 3; ++k) { c.lo[k] = lo[k]; c.hi[k] = hi[k]; }
                        c.dist = dist;
                        c.vol = vol;
                        sCand[h] = c;
                    }
                }
                __syncthreads();   // (B2) candidates visible
                E2_TICK(2);
                if (tid == 0) sNOut[cur ^ 1] = 0;
                int hm = -1;
                for (int h = E2_HM_MAX; h >= 0; --h)
                    if (sCand[h].vol * 1.05f <= (float)E2_BOX_ELEMS) { hm = h; break; }
                // ---------------- (c) rows of the box: x-interval of each (y,z) row inside the slab, packed by a prefix sum.
                // The size above is an estimate; when the exact size does not fit, the margin is reduced and the rows redone.
                int lo[3] = {0, 0, 0}, hi[3] = {-1, -1, -1};
                int xloE = 0, ny = 1, rows = 0, rpt = 0, total = 0, base = 0;
                int rowXs[E2_MAXROWS / E2_THREADS], rowLen[E2_MAXROWS / E2_THREADS];
                bool haveBox = false;
                int retries = 0;
                while (hm >= 0) {
                    const Cand c = sCand[hm];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { lo[k] = c.lo[k]; hi[k] = c.hi[k]; }
                    const float slabD = c.dist + kappa + 0.05f;
                    xloE = lo[0] & ~1;
                    const int xhiT = hi[0] + 1;
                    ny = hi[1] - lo[1] + 2;
                    rows = ny * (hi[2] - lo[2] + 2);
                    rpt = (rows + E2_THREADS - 1) / E2_THREADS;     // rows per thread (<= 6)
                    int mine = 0;
#pragma unroll
                    for (int j = 0; j < E2_MAXROWS / E2_THREADS; ++j) {
                        rowXs[j] = 0;
                        rowLen[j] = 0;
                        const int r = tid * rpt + j;
                        if (j < rpt && r < rows) {
                            const int bz = r / ny, by = r - bz * ny;
                            const float cyz = nc[1] * (float)(lo[1] + by) + nc[2] * (float)(lo[2] + bz);
                            int xs = xloE, xe = xhiT;
                            if (fabsf(nc[0]) > 1e-3f) {
                                const float inv = 1.0f / nc[0];
                                const float x1 = (-slabD - cyz) * inv, x2 = (slabD - cyz) * inv;
                                const float xa = fminf(fmaxf(fminf(x1, x2), -1.0e6f), 1.0e6f), xb = fminf(fmaxf(fmaxf(x1, x2), -1.0e6f), 1.0e6f);
                                // taps are integers: those inside [xa, xb] are ceil(xa) .. floor(xb)
                                xs = max(xs, ((int)ceilf(xa)) & ~1);
                                xe = min(xe, (int)floorf(xb));
                            } else if (fabsf(cyz) > slabD) {
                                xe = xs - 1;
                            }
                            if (xe >= xs) {
                                rowXs[j] = xs;
                                rowLen[j] = (xe - xs + 2) & ~1;
                            }
                            mine += rowLen[j];
                        }
                    }
                    int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    if (lane == 31) sWarpTot[warp] = incl;
                    __syncthreads();   // (C) warp totals visible
                    total = 0;
                    base = incl - mine;
#pragma unroll
                    for (int w2 = 0; w2 < E2_THREADS / 32; ++w2) {
                        const int v = sWarpTot[w2];
                        if (w2 < warp) base += v;
                        total += v;
                    }
                    if (total > 0 && total <= E2_BOX_ELEMS) { haveBox = true; break; }
                    if (++retries >= 3) break;
                    hm -= retries;          // 1, then 2 less
                    __syncthreads();       // sWarpTot is rewritten
                }
                if (!haveBox) hm = -1;
                // rotations beyond the margin -> list for path (d); no box at all -> every rotation takes path (d)
                const bool core = haveBox && sCls[rloc].need <= hm;
                if (ph == 0 && rvalid && haveBox && !core) sOut[atomicAdd(&sNOut[cur], 1)] = rloc;
                if (A.stats && tid == 0) {
                    atomicAdd(&A.stats[0], 1ull);                                   // tiles
                    atomicAdd(&A.stats[1], haveBox ? 1ull : 0ull);                  // tiles with a staged box
                    atomicAdd(&A.stats[2], (unsigned long long)(hm >= 0 ? hm : 0)); // sum of margins
                    atomicAdd(&A.stats[3], (unsigned long long)(haveBox ? total : 0));          // staged elements
                    atomicAdd(&A.stats[5], (unsigned long long)nRc);                // (rotation, tile) pairs
                    atomicAdd(&A.stats[6], (unsigned long long)retries);            // margin reductions (estimate too optimistic)
                    atomicAdd(&A.stats[7], (unsigned long long)rows);
                }
                if (haveBox) {
                    fence_proxy_async();
                    if (tid == 0) mbar_arrive_expect_tx(&sBar, (uint32_t)total * 8u);
#pragma unroll
                    for (int j = 0; j < E2_MAXROWS / E2_THREADS; ++j) {
                        const int r = tid * rpt + j;
                        if (j < rpt && r < rows) {
                            sBias[r] = base - rowXs[j] - THB_FLOOR_BIAS;
                            if (rowLen[j] > 0) {
                                const int bz = r / ny, by = r - bz * ny;
                                const int ym = wrap_idx(lo[1] + by, n), zm = wrap_idx(lo[2] + bz, n);
                                tma_bulk_g2s(box + base, vol + ((size_t)zm * n + ym) * pitch + rowXs[j], (uint32_t)rowLen[j] * 8u, &sBar);
                            }
                            base += rowLen[j];
                        }
                    }
                }
                __syncthreads();   // (D) out-list and row table visible
                E2_TICK(3);
                if (A.stats && tid == 0) atomicAdd(&A.stats[4], (unsigned long long)(haveBox ? sNOut[cur] : nRc));   // pairs on path (d)
                // ---------------- (d) non-core rotations: pixels on the lanes, gather from L1/L2
                {
                    const int nOut = haveBox ? sNOut[cur] : nRc;
                    for (int it = warp; it < nOut; it += E2_THREADS / 32) {
                        const int rl = haveBox ? sOut[it] : it;
                        const Rot2 ro = sRot[rl];
                        float v[E_TC];
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) v[t] = 0.0f;
#pragma unroll 2
                        for (int k = lane; k < td.count; k += 32) {
                            const PixelRec& rec = tile[k];
                            float x, y, z;
                            slice_coord(ro, rec.a, rec.b, x, y, z);
                            const float2 pr = gather_ft_pitched(vol, n, pitch, x, y, z);
                            const float m = rec.g * fmaf(pr.x, pr.x, pr.y * pr.y);
#pragma unroll
                            for (int t = 0; t < E_TC; ++t) v[t] += fmaf(rec.u[t].x, pr.x, fmaf(rec.u[t].y, pr.y, m));
                        }
                        float s07, s8;
                        warp_sum9(v, lane, s07, s8);
                        if ((lane & 3) == 0) sAcc[rl * E_TC + ((lane >> 2) & 7)] += s07;   // one warp per (rotation, tile)
                        if (lane == 1) sAcc[rl * E_TC + 8] += s8;
                    }
                }
                // ---------------- (e) core rotations: gather from the staged box
                if (haveBox) {
                    E2_TICK(4);
                    mbar_wait(&sBar, barParity);
                    E2_TICK(5);
                    barParity ^= 1;
                    if (core) {
                        // origin of the row table in biased cell coordinates (see fold_floor_fast)
                        const int oy = lo[1] + THB_FLOOR_BIAS, oz = lo[2] + THB_FLOOR_BIAS;
#pragma unroll 2
                        for (int k = ph; k < td.count; k += 2) {
                            const PixelRec& rec = tile[k];
                            float x, y, z;
                            slice_coord(rot, rec.a, rec.b, x, y, z);
                            int xb, yb, zb;
                            float xd, yd, zd;
                            const bool conj = fold_floor_fast(x, y, z, xb, yb, zb, xd, yd, zd);
                            float w[8];
                            tri_weights(xd, yd, zd, w);
                            const int* bi = sBias + (zb - oz) * ny + (yb - oy);
                            const float2* r0 = box + (bi[0] + xb);
                            const float2* r1 = box + (bi[1] + xb);
                            const float2* r2 = box + (bi[ny] + xb);
                            const float2* r3 = box + (bi[ny + 1] + xb);
                            const float2 v0 = r0[0], v1 = r0[1], v2 = r1[0], v3 = r1[1];
                            const float2 v4 = r2[0], v5 = r2[1], v6 = r3[0], v7 = r3[1];
                            float re = v0.x * w[0], im = v0.y * w[0];
                            re = fmaf(v1.x, w[1], re); im = fmaf(v1.y, w[1], im);
                            re = fmaf(v2.x, w[2], re); im = fmaf(v2.y, w[2], im);
                            re = fmaf(v3.x, w[3], re); im = fmaf(v3.y, w[3], im);
                            re = fmaf(v4.x, w[4], re); im = fmaf(v4.y, w[4], im);
                            re = fmaf(v5.x, w[5], re); im = fmaf(v5.y, w[5], im);
                            re = fmaf(v6.x, w[6], re); im = fmaf(v6.y, w[6], im);
                            re = fmaf(v7.x, w[7], re); im = fmaf(v7.y, w[7], im);
                            if (conj) im = -im;
                            nrm = fmaf(rec.g, fmaf(re, re, im * im), nrm);
#pragma unroll
                            for (int t = 0; t < E_TC; ++t) acc[t] = fmaf(rec.u[t].x, re, fmaf(rec.u[t].y, im, acc[t]));
                        }
                    }
                }
            }
            // ---- end of the pass over the tiles: combine halves, fallback sums and the constant term
            __syncthreads();
            E2_TICK(6);
            if (firstPass) {
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
            float* park = reinterpret_cast<float*>(tile);     // 128 x 10 floats = 5 KB
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
        if (THB_E2_TIMERS && A.stats)
            for (int i = 0; i < 8; ++i) atomicAdd(&A.stats[8 + i], (unsigned long long)tacc[i]);
    }
}

}  // namespace thb
