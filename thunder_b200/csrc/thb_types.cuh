// thb_types.cuh - argument structs shared by the kernels and the host-side launch code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace thb {

constexpr int THB_MAX_SLOTS = 32;   // volume / accumulator slots: half maps in 3D, classes in 2D (demo_2D.json: k = 20)
constexpr int E_THREADS = 128;   // one rotation sample per thread
constexpr int E_TILE = 128;      // pixels staged in shared memory per step
constexpr int E_TC = 9;          // translations carried in registers per pass (mLT default = 9)
constexpr int M_THREADS = 256;
constexpr int M_MAXRECO = 128;
constexpr int M_KP = 1;          // pixels per thread held in registers by the insert kernel

struct VolTable { const float2* p[THB_MAX_SLOTS]; };
struct AccTable { float4* p[THB_MAX_SLOTS]; double* O; int* counter; };

// strided view of a [particle][sample][component] array (API arrays and the SoA particle state
// use different strides)
struct View3 {
    const double* p;
    long long sP, sS, sC;
#ifdef __CUDACC__
    __device__ __forceinline__ double at(long long ip, long long is, long long ic) const
    {
        return p[ip * sP + is * sS + ic * sC];
    }
#endif
};

struct __align__(16) PixelE {
    double a, b;        // pf*iCol, pf*iRow
    float ctf, sig;
    float2 d[E_TC];     // dat * conj(tra_t)
};
static_assert(sizeof(PixelE) == 96, "PixelE must be 96 bytes");

// 8x8-pixel tile of the blocked pixel order: pixels [start, start+count) and the rectangle they span
// in padded units (centre ca,cb ; half extents ha,hb)
struct TileDesc {
    int start, count;
    float ca, cb, ha, hb;
    int pad0, pad1;
};

struct ExpectArgs {
    VolTable vols;
    VolTable quads;         // the same volumes in the quad layout (thb_expect3.cuh), or null
    int vdim;
    int pitch;              // row pitch of the volumes in elements (>= vdim/2 + 2, multiple of 4)
    const TileDesc* tiles;
    int nTiles;
    float* work;            // [nAct][nR*nT] scratch for multi-pass shapes (nR > 128 or nT > 9), else null
    int quadBrick;          // log2 of the brick edge of the quad layout (0 = plain rows)
    int sortRot;            // hand rotation slots out in the order of the cloud along its widest axis
    unsigned long long* stats;   // optional [8] counters of the staging decisions (null = off)
    const float2* dat;
    const float* ctf;
    const float* sig;
    const int* slotOfImg;
    const int4* pix;
    int P, N;
    int mode2D;             // MODE_2D: quat = (cos, sin, -, -) of the in-plane rotation, references are 2-plane volumes
    int slotAll;            // >= 0: every image against this slot (2D classification scan), else slotOfImg
    int nAct;
    const int* imgIdx;      // may be null: image = particle index + imgBase
    int imgBase;
    const unsigned char* active;   // may be null; skip particle when active[p] == 0
    int nR, nT;
    View3 quat, tran, wR, wT;
    float* uR;      // [nAct][nR]
    float* uT;      // [nAct][nT]
    float* uC;      // [nAct]
    float* base;    // [nAct]
    float* logL;    // [nAct][nR][nT] or null  ([nAct][nR][nT][nD] with CTF search)
    // ---- CTF search (SEARCH_TYPE_CTF, src/Optimiser.cpp:1248-1273, 1383-1402): nD defocus factors per image, the CTF of every one
    // computed on the fly from the per-pixel defocus and frequency (allocPreCal, :8125-8168); nD == 0: off
    int nD;
    const float* defP;      // [nImg][P]   _defocusP, resident with the E stack (blocked pixel order)
    const float* freq;      // [P]         _frequency
    View3 dpar, wD;         // [nAct][nD]  defocus factors and their prior weights
    const float* ctfK;      // [nAct][4]   K1, K2, phaseShift, amplitudeContrast of the image
    float* uD;              // [nAct][nD]
    // ---- lockstep launch (thb_expect7.cuh): a persistent grid walks the images wave by wave, all CTAs on the same pixel tiles at
    // the same time, so that the cells the chip reads at any moment are one thin spherical shell of the volume (pixels in radial order)
    const int* order;       // [nAct] launch position -> particle (images of one slot adjacent), or null
    unsigned int* lockCtr;  // arrival counters of the tile barriers, one per (wave, barrier), zeroed before the launch; null = free-running
    int lockTiles;          // one barrier every lockTiles tiles
    int lockWindow;         // a CTA may run this many barriers ahead of the slowest one
    int scanSlot1;          // global scan: 1 + the ONE reference (slot) every image of the launch is compared with, rotations
                            // shared by all images (quat.sP == 0): eligible for the shared-template path (thb_expect8.cuh); 0 = no
};

struct InsertArgs {
    AccTable acc;
    int vdim;                 // padded accumulator dimension
    const float2* dat;
    const float* ctf;
    const int* slotOfImg;
    const int4* pix;
    int P, N;
    int mode2D;
    int nImg;
    const int* imgIdx;        // may be null
    int imgBase;
    int mReco;
    const float* w;           // [nImg] or null (then wAll)
    float wAll;
    const double* offS;       // [nImg][2] indexed by image, or null
    View3 nr, nt;             // [nImg][mReco][4], [nImg][mReco][2]
    const int* drawR;         // optional [nImg][mReco] indices into nr/nt sample axis (particle filter draws)
    const int* drawT;
    const int* drawC;         // optional [nImg][mReco] accumulator slot of every draw (MODE_2D classes), else slotOfImg
    const int* drawCount;     // optional [nImg]: only the first drawCount[l] <= mReco draws of image l are inserted
    // ---- CTF search (cSearch of InsertFT, src/Optimiser.cpp:7171-7215): the CTF of every draw from its own defocus factor
    View3 nd;                 // [nImg][mReco][1] defocus factors, p == null: off
    const int* drawD;         // optional [nImg][mReco] indices into the sample axis of nd (particle filter draws)
    const float* ctfAttr;     // [nImg][7] voltage, defocusU, defocusV, defocusTheta, Cs, amplitudeContrast, phaseShift
    float pixelSize;
};

}  // namespace thb
