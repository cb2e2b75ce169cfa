// thb_context.h - the opaque context behind the C ABI (host side, C++).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/thunder_b200.h"
#include "thb_types.cuh"

namespace thb {

enum KernelFamily { KF_EXPECT = 0, KF_INSERT = 1, KF_PF = 2, KF_PACK = 3, KF_COMM = 4, KF_COUNT = 5 };

struct TimedSpan {
    cudaEvent_t a, b;
    int family;
};

struct Stack {
    int nImg = 0;                // capacity (images)
    float2* dat = nullptr;
    float* ctf = nullptr;
    float* sig = nullptr;
    float* def = nullptr;        // E stack, CTF search only: per-pixel defocus (_defocusP)
    int* slot = nullptr;
    std::vector<int> hslot;      // host copy of slot[] (validation of the slots a launch will touch)
};

// projector volume in HBM: FFTW half-complex rows of vdim/2+1 elements, stored with a row pitch of
// `pitch` elements (multiple of 4 = 32 bytes, >= vdim/2+2) so that every row is 16-byte aligned for the
// TMA bulk copies and a box row may over-read one element past the Nyquist column
struct Volume3 {
    float2* d = nullptr;
    int vdim = 0;
    int pitch = 0;
    int quadBrick = -1;          // brick setting the quad copy was built with
    int quadOct = -1;            // 0: quad (32 B / voxel), 1: oct (64 B / voxel) copy
    void* quad = nullptr;        // quad layout for the direct-gather kernel: n*n*(n/2) x 32 bytes (built on upload)
};

struct Accum {
    float4* d = nullptr;
    int vdim = 0;
    size_t nVox = 0;
};

// device-resident particle-filter state, SoA with the particle index fastest:
//   r[(c*mLR + i)*nPar + p], t[(c*mLT + i)*nPar + p], wR[i*nPar + p], wT[i*nPar + p]
struct PFState {
    int nPar = 0;
    thb_pf_params prm{};
    double* r = nullptr;
    double* t = nullptr;
    double* wR = nullptr;
    double* wT = nullptr;
    double* scal = nullptr;      // [20][nPar]
    double* dbl = nullptr;       // double scratch: uR, uT (as double), resampling buffers
    int imgBase = 0;             // particle p <-> image imgBase + p of the resident stacks
    uint64_t streamBase = 0;     // offset of the per-particle RNG stream (rank / batch offset)
    float* uR = nullptr;         // [nPar][mLR]
    float* uT = nullptr;         // [nPar][mLT]
    float* uC = nullptr;
    float* base = nullptr;
    unsigned char* active = nullptr;
    int* order = nullptr;        // [nPar] compacted list of the particles still active (adaptive E-step)
    int* nPhase = nullptr;
    double* vari = nullptr;      // [3][nPar] best variR, variT, and no-decrease counter
    int* drawR = nullptr;        // [nPar][mReco]
    int* drawT = nullptr;
    int* drawD = nullptr;
    // CTF search (prm.mLD > 0)
    double* d = nullptr;         // [(mLD + 1)][nPar]: defocus factors, last row = the most likely one
    double* wD = nullptr;        // [mLD][nPar]
    double* uDd = nullptr;       // [mLD][nPar] likelihood weights as doubles (resampling)
    float* uD = nullptr;         // [nPar][mLD] E-kernel output
    float* ctfK = nullptr;       // [nPar][4]
    float* ctfAttr = nullptr;    // [nPar][7]
    float pixelSize = 0.f;
    bool ctfSet = false;
    int mode2D = 0;              // the support was built in MODE_2D (thb_pf_from_scan): von Mises operators in the phase loop
    int drawCap = 0;
    uint64_t epoch = 0;          // advances the counter-based RNG stream between calls
    float* traceR = nullptr;     // option "pf_trace": the marginal weights of every phase, [traceCap][nPar][mLR] / [..][mLT]
    float* traceT = nullptr;
    float* traceB = nullptr;     // [traceCap][nPar]: the baselines (largest log-likelihood of the phase)
    double* traceSt = nullptr;   // and the support: [traceCap][2: after the perturbation, after the resampling][nPar][4 mLR + 2 mLT]
    int traceCap = 0, traceN = 0, traceWant = 0;
};

}  // namespace thb

#define THB_N_SCRATCH 16
#define THB_DEFAULT_EXPECT_IMPL 7
struct thb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int smCount = 148;

    // pixel sets
    int N = 0, pf = 0, nPxlE = 0, nPxlM = 0, NM = 0, pfM = 0;
    int4* pixE = nullptr;        // device pixel lists, in the blocked order
    int4* pixM = nullptr;
    int* permE = nullptr;        // blocked position -> caller's pixel index
    int* permM = nullptr;
    float* freqE = nullptr;      // CTF search: |k| / (N pixelSize) of the E pixel list (_frequency), blocked order
    void* segM = nullptr;        // thb::Seg[]: row-major runs {j, iFirst, count, start} of the M pixel list (slab insert, thb_insert2.cuh)
    int nSegM = 0;
    float rMaxPadM = 0.f;        // largest |(pf i, pf j)| of the M pixel list
    int insertSlabMB = 48;       // option "insert_slab_mb": accumulator bytes one slab of the slab insert may span
    int insertSlabPlanes = 0;    // option "insert_slab_planes": slab thickness override (tests), 0 = from insert_slab_mb
    thb::TileDesc* tilesE = nullptr;   // 8x8-pixel tiles of the E pixel list (blocked order)
    int nTilesE = 0;
    int mode2D = 0;              // thb_set_mode: references are images, rotations in-plane (MODE_2D of the reference)
    int insertImpl = 0;
    unsigned long long* dStats = nullptr;   // [8] staging counters of the E kernel (option "stats")
    int statsOn = 0;
    int tileW = 8, tileH = 8;   // pixel tile of the E pixel list (tileW * tileH <= 128)
    int expectImpl = THB_DEFAULT_EXPECT_IMPL;   // 7: several rotations per lane, lockstep launch (default); 3: direct gather, one rotation per lane;
                                 // 2: TMA-staged box; 1: direct gather, linear layout; 4 / 5: paired-lane / pixels-on-lanes variants
    int expectRpl = 2;           // option "expect_rpl": rotations per lane of expect_impl 7 (2 or 4)
    int expectOrder = 1;         // option "expect_order": pixel order of the E stack, 0 = 8x8 blocks, 1 = radial (rings)
    int expectOrderBuilt = 1;    // the order the current E pixel list (and the resident E stack) was built with
    int expectLock = 1;          // option "expect_lock": lockstep launch of expect_impl 7
    int expectLockTiles = 1;     // option "expect_lock_tiles": one barrier every so many tiles of 128 pixels
    int expectLockWindow = 1;    // option "expect_lock_window": barriers a CTA may run ahead of the slowest one
    std::vector<int> expectOrderHost;   // launch order of the lockstep launch (images of one slot adjacent)
    int pfCompact = 1;           // option "pf_compact": adaptive E-step launches only the unfinished particles (compacted list)
    int pfStage = 1;             // option "pf_stage": the particle filter stages the state of a particle in shared memory
    int scanTemplates = 1;       // option "scan_templates": scans project each shared rotation once per launch (thb_expect8.cuh)
    int quadBrick = 2;           // log2 brick edge of the quad layout (option "quad_brick"; 4x4x4 bricks measured best)
    int quadOct = 1;             // option "quad_oct": whole trilinear cell in one 64-byte element (8x volume bytes; default,
                                 // falls back to the 32-byte quad when HBM is short)
    int sortRot = 0;             // option "sort_rot"
    int expectSpread = -1;       // option "expect_spread": spread each image over many CTAs (-1: when a launch has few images)
    int expectMinBlocks = 2;     // CTAs per SM the quad kernel is compiled for (2 or 3)

    thb::Volume3 vols[thb::THB_MAX_SLOTS];
    thb::Accum accs[thb::THB_MAX_SLOTS];
    double* dO = nullptr;        // [THB_MAX_SLOTS][3]
    int* dCounter = nullptr;     // [THB_MAX_SLOTS]

    thb::Stack stackE, stackM;
    thb::PFState pf_;

    void* reco = nullptr;        // thb::RecoState (cuFFT plans, kernel table, last reconstruction), thb_reco.cu

    // NCCL
    void* ncclComm = nullptr;
    int nRanks = 1, rank = 0;
    void* commBuf = nullptr;     // wire buffer of the all-reduce: 3 floats per voxel, all slots back to back
    size_t commBufBytes = 0, commBytesLast = 0;

    // accounting
    cudaEvent_t tA = nullptr, tB = nullptr;
    int64_t launches = 0;
    bool timing = false;
    std::vector<thb::TimedSpan> spans;
    double famMs[thb::KF_COUNT] = {0, 0, 0, 0, 0};
    int64_t famN[thb::KF_COUNT] = {0, 0, 0, 0, 0};

    // scratch device buffers (grown on demand)
    void* scratch[THB_N_SCRATCH] = {};
    size_t scratchCap[THB_N_SCRATCH] = {};
    // second stream for host->device image uploads that overlap the kernels of the previous batch
    cudaStream_t copyStream = nullptr;
    cudaEvent_t copyDone = nullptr;
    bool copyPending = false;
};

namespace thb {

int set_error(thb_ctx* ctx, int code, const char* fmt, ...);
int cuda_fail(thb_ctx* ctx, cudaError_t e, const char* what);
void* scratch(thb_ctx* ctx, int which, size_t bytes);   // nullptr on failure (error set)
void span_begin(thb_ctx* ctx, int family);
void span_end(thb_ctx* ctx);
void resolve_spans(thb_ctx* ctx);

#define THB_CUDA(ctx, call)                                         \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return thb::cuda_fail(ctx, e__, #call); \
    } while (0)

// launches (thb_launch.cu)
int launch_expect_local(thb_ctx* ctx, const ExpectArgs& a);
int launch_insert(thb_ctx* ctx, const InsertArgs& a, const int* hImgIdx /* host copy of a.imgIdx, or null */);
int check_expect_state(thb_ctx* ctx, const char* who);   // returns the common volume edge, or a negative error
int check_insert_slots(thb_ctx* ctx, int nImg, const int* imgIdx, int imgBase, const char* who);
VolTable vol_table(const thb_ctx* ctx);
AccTable acc_table(const thb_ctx* ctx);

// NCCL (thb_comm.cpp)
int comm_allreduce(thb_ctx* ctx);
void comm_destroy(thb_ctx* ctx);

// particle filter (thb_pf.cu)
void pf_free(thb_ctx* ctx);

// reconstruct / setProjectee (thb_reco.cu)
void reco_free(thb_ctx* ctx);

}  // namespace thb
