// thb_expect4.cuh - fused E kernel, local-search shape: the direct gather of thb_expect3.cuh with TWO LANES PER SAMPLE.
//
// Why: in the benchmark regime (clouds of degrees) every sample is its own 64-byte cell from HBM and a warp-wide LDG.256 of
// thb_expect3.cuh touches 32 distinct 128-byte lines; the L1/TEX tag stage takes one cycle per distinct line, two such loads
// per sample = 2 cycles/sample/SM (tools/gpu/gatherbench.cu), and ncu shows L1/TEX 82 % busy next to DRAM at 73 %.
// Here lanes 2p and 2p+1 serve the SAME (rotation, pixel) sample and load the two 32-byte halves of its cell: one warp-wide
// load = 16 samples = 16 distinct lines, half the tag-stage cycles per sample.  The price is arithmetic: both lanes compute
// the coordinates, each interpolates its own z plane (4 taps), one shuffle pair joins the halves, and the translations are
// split 5 / 4 (+ the |p|^2 term) between the lanes - about 1.5x the issue slots of a kernel that used a third of them.
//
// 256 threads = 8 warps, warp w serves rotations [16 w, 16 w + 16) of a 128-rotation pass and walks every pixel of the
// 128-pixel tile.  Records, expanded likelihood, passes, epilogue: as thb_expect3.cuh.
#pragma once
#include <cuda_runtime.h>
#include "thb_math.cuh"
#include "thb_types.cuh"
#include "thb_expect3.cuh"

namespace thb {

constexpr int E4_THREADS = 256;
constexpr int E4_ROTS = 128;
constexpr int E4_TILE = 128;
constexpr int E4_TH = 5;         // translations per lane of a pair (lane 0: 0..4, lane 1: 5..8 and the norm term)
constexpr size_t E4_SMEM_BYTES = E4_TILE * sizeof(PixelRec);
static_assert(E_TC == 9, "the 5 / 4 split below is written for 9 translations per pass");

template <bool OCT, bool M2D>
__global__ void __launch_bounds__(E4_THREADS, 2) expect_pair_kernel(const ExpectArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PixelRec* tile = reinterpret_cast<PixelRec*>(smem_raw);
    __shared__ float sRC[E_TC], sRR[E_TC];
    __shared__ float redf[E4_THREADS / 32];
    __shared__ double redd[E4_THREADS / 32];

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
    const int sub = lane & 1;                       // which half of the cell / of the translations this lane serves
    const int rloc = warp * 16 + (lane >> 1);       // rotation slot of the pair within a pass
    const int nRT = A.nR * A.nT;
    const bool single = A.nR <= E4_ROTS && A.nT <= E_TC;
    float* sL = single ? reinterpret_cast<float*>(smem_raw + E4_SMEM_BYTES) : A.work + (size_t)p * nRT;
    double k0sum = 0.0;          // sum_i sig_i |dat_i|^2
    const int LB = A.quadBrick;

    for (int rbase = 0; rbase < A.nR; rbase += E4_ROTS) {
        const int nRc = min(E4_ROTS, A.nR - rbase);
        const bool rvalid = rloc < nRc;
        Rot2 rot;
        {
            double q[4] = {1.0, 0.0, 0.0, 0.0};
            if (rvalid)
                for (int c = 0; c < (M2D ? 2 : 4); ++c) q[c] = A.quat.at(p, rbase + rloc, c);
            rot = make_rot2(q, M2D);
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
            float acc[E4_TH];    // lane 0: translations 0..4; lane 1: translations 5..8, acc[4] = the norm term
#pragma unroll
            for (int t = 0; t < E4_TH; ++t) acc[t] = 0.0f;
            const bool firstPass = (rbase == 0 && tbase == 0);

            for (int tile0 = 0; tile0 < P; tile0 += E4_TILE) {
                const int cnt = min(E4_TILE, P - tile0);
                __syncthreads();   // previous tile consumed (also orders the sRC / sRR writes)
                {
                    // pixel records: 2 threads per pixel, translations split between them
                    const int k = tid >> 1, sb = tid & 1;
                    if (k < cnt) {
                        const int i = tile0 + k;
                        const int4 c = A.pix[i];
                        const float2 d = dat[i];
                        const float cf = ctf[i], sg = sig[i];
                        const float m2 = -2.0f * sg * cf;
                        PixelRec& rec = tile[k];
                        if (sb == 0) {
                            rec.a = (double)c.x;
                            rec.b = (double)c.y;
                            rec.g = sg * cf * cf;
                            rec.pad = 0.0f;
                            if (firstPass) k0sum += (double)(sg * (d.x * d.x + d.y * d.y));
                        }
#pragma unroll
                        for (int t = 0; t < E_TC; ++t) {
                            if ((t & 1) != sb) continue;
                            const float phs = translate_phase(c.z, c.w, sRC[t], sRR[t]);
                            float s, co;
                            sincosf(phs, &s, &co);
                            rec.u[t] = make_float2(m2 * (d.x * co - d.y * s), m2 * (d.x * s + d.y * co));
                        }
                    }
                }
                __syncthreads();
                // every lane of a warp that holds at least one valid pair runs the loop (the shuffles need the whole warp);
                // invalid pairs carry the identity rotation and are not written out
                if (warp * 16 < nRc) {
#pragma unroll 4
                    for (int k = 0; k < cnt; ++k) {
                        const PixelRec& rec = tile[k];
                        float x, y, z;
                        slice_coord(rot, rec.a, rec.b, x, y, z);
                        int xb, yb, zb;
                        float xd, yd, zd;
                        const bool conj = fold_floor_fast(x, y, z, xb, yb, zb, xd, yd, zd);
                        const int x0 = xb - THB_FLOOR_BIAS, y0 = yb - THB_FLOOR_BIAS, z0 = zb - THB_FLOOR_BIAS;
                        const int ym = y0 < 0 ? y0 + n : y0;
                        const int zm = z0 < 0 ? z0 + n : z0;
                        const Quad* q;
                        if (OCT) {
                            q = vol + 2 * quad_index(x0, ym, zm, n, LB) + sub;
                        } else {
                            const int zm1 = (z0 + 1 < 0) ? z0 + 1 + n : z0 + 1;
                            q = vol + quad_index(x0, ym, sub ? zm1 : zm, n, LB);
                        }
                        Quad a = Quad{};
                        if (!M2D || sub == 0) a = ldg_quad(q);
                        // this lane's z plane: weights (1 - xd | xd)(1 - yd | yd) * (1 - zd | zd), same products as tri_weights
                        const float vx0 = 1.0f - xd, vy0 = 1.0f - yd;
                        const float vz = sub ? zd : 1.0f - zd;
                        const float w0 = (vx0 * vy0) * vz, w1 = (xd * vy0) * vz, w2 = (vx0 * yd) * vz, w3 = (xd * yd) * vz;
                        float re = a.v00.x * w0, im = a.v00.y * w0;
                        re = fmaf(a.v10.x, w1, re); im = fmaf(a.v10.y, w1, im);
                        re = fmaf(a.v01.x, w2, re); im = fmaf(a.v01.y, w2, im);
                        re = fmaf(a.v11.x, w3, re); im = fmaf(a.v11.y, w3, im);
                        re += __shfl_xor_sync(0xffffffffu, re, 1);
                        im += __shfl_xor_sync(0xffffffffu, im, 1);
                        if (conj) im = -im;
                        const float2* u = rec.u + sub * E4_TH;
#pragma unroll
                        for (int t = 0; t < E4_TH - 1; ++t) acc[t] = fmaf(u[t].x, re, fmaf(u[t].y, im, acc[t]));
                        // fifth slot: translation 4 on lane 0, g |p|^2 on lane 1 (u[9] does not exist: never dereferenced)
                        const float ux = sub ? rec.g * re : rec.u[E4_TH - 1].x;
                        const float uy = sub ? rec.g * im : rec.u[E4_TH - 1].y;
                        acc[E4_TH - 1] = fmaf(ux, re, fmaf(uy, im, acc[E4_TH - 1]));
                    }
                }
            }
            // ---- end of the pass: constant term, the norm term of the pair
            __syncthreads();
            if (firstPass) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) k0sum += __shfl_xor_sync(0xffffffffu, k0sum, o);
                if (lane == 0) redd[warp] = k0sum;
                __syncthreads();
                double s = 0.0;
                for (int w2 = 0; w2 < E4_THREADS / 32; ++w2) s += redd[w2];
                k0sum = s;
                __syncthreads();
            }
            const float nrmPair = __shfl_sync(0xffffffffu, acc[E4_TH - 1], lane | 1);     // lane 1 of the pair holds it
            if (rvalid) {
                const double base = k0sum + (double)nrmPair;
#pragma unroll
                for (int t = 0; t < E4_TH; ++t) {
                    const int tt = sub * E4_TH + t;
                    if (tt < E_TC && tbase + tt < A.nT) sL[(size_t)(rbase + rloc) * A.nT + tbase + tt] = (float)(base + (double)acc[t]);
                }
            }
        }
    }
    __syncthreads();

    expect_epilogue<E4_THREADS>(A, p, sL, redf, redd);
}

}  // namespace thb
