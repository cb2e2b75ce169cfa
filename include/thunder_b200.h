/*
 * thunder_b200.h - C ABI of the B200-native Optimiser hot path for THUNDER.
 *
 * One shared library (libthunder_b200.so), one opaque context per GPU / per process.
 * Plain pointers and sizes only: no THUNDER, Eigen, MPI or torch types.  Every entry point
 * returns 0 on success and a negative THB_E_* code on failure (thb_last_error() gives the
 * message); nothing in here ever calls exit()/abort(), unlike the seam it replaces
 * (reference gpu/config/Device.cuh.in:27-61).
 *
 * The entry points are what a binding of the reference's GPU seam for this path needs.
 * Reference interface each one replaces (paths relative to the THUNDER tree):
 *
 *   thb_pixel_list            Optimiser::allocPreCalIdx              src/Optimiser.cpp:7991-8041
 *                             (+ ExpectPreidx                        gpu/interface/Interface.h:18)
 *   thb_set_expect_pixels     ExpectPreidx / ExpectPrefre            gpu/interface/Interface.h:18-29
 *   thb_set_insert_pixels     Reconstructor::setPreCal               src/Reconstructor.cpp (setPreCal), InsertFT args iCol/iRow
 *   thb_set_volume            ExpectLocalV3D / ManagedArrayTexture   gpu/interface/Interface.h:31-164
 *                             (Projector::projectee3D(), src/Projector.cpp:123-148)
 *   thb_pack_stack            Optimiser::allocPreCal + CTF()         src/Optimiser.cpp:8043-8171, src/CTF.cpp:118-151
 *   thb_upload_stack          ExpectLocalP / the datP,ctfP,sigRcpP   src/Optimiser.cpp:8043-8171 (allocPreCal)
 *                             arguments of ExpectGlobal3D, InsertFT
 *   thb_project               Projector::project(Complex*,...)       src/Projector.cpp:356-374
 *   thb_expect_local          ExpectLocalRTD + ExpectLocalPreI3D +   gpu/interface/Interface.h:31-164
 *                             ExpectLocalM (one call, many images)   (CPU loop: src/Optimiser.cpp:1162-1402)
 *   thb_set_frequency,        ExpectPrefre, defO of ExpectLocalP,    gpu/interface/Interface.h:24-164
 *   thb_upload_stack_defocus, k1/k2/.. of ExpectLocalPreI3D, wD of   (CPU loop with SEARCH_TYPE_CTF: src/Optimiser.cpp:1248-1273,
 *   thb_expect_local_ctf      ExpectLocalM                            1328-1402; allocPreCal :8125-8168)
 *   thb_insert_ctf            InsertFT with cSearch                  gpu/interface/Interface.h:267-318, src/Optimiser.cpp:7171-7215
 *   thb_expect_scan           ExpectRotran + ExpectProject +         gpu/interface/Interface.h:199-221
 *                             ExpectGlobal3D                         (CPU loop: src/Optimiser.cpp:633-914)
 *   thb_reco_alloc/reset      Reconstructor::allocSpace / reset      src/Reconstructor.cpp:92-143
 *   thb_insert                InsertFT                               gpu/interface/Interface.h:267-318
 *                             (CPU loop: src/Optimiser.cpp:7036-7241; insertP src/Reconstructor.cpp:782-863;
 *                              insertDir :407-422)
 *   thb_comm_* / thb_allreduce Reconstructor::allReduceF/T/O         src/Reconstructor.cpp:2350-2520
 *                             (NCCL twin: gpu/src/cuthunder.cu:5294-5324, 5903-5985)
 *   thb_reco_download         the F3D/T3D/O3D/counter out-arguments  gpu/interface/Interface.cpp:581-619
 *                             of InsertFT, + prepareTF normalisation src/Reconstructor.cpp:1056-1091, 2458-2483
 *   thb_reconstruct           Reconstructor::reconstruct             src/Reconstructor.cpp:1129-1831 (GPU twin reconstructG :1835-2346)
 *   thb_set_projectee         Projector::setProjectee                src/Projector.cpp:123-148
 *   thb_remask_pack           Optimiser::reCentreImg / reMaskImg     src/Optimiser.cpp:6065-6151 (GPU twin reMaskImgG)
 *   thb_sigma_accumulate      Optimiser::allReduceSigma (image loop) src/Optimiser.cpp:6397-6709
 *   thb_pf_* / thb_expectation Particle::perturb/resample/calVari/.. src/Particle.cpp:1004-1478, 1964-2002, 2309-2495
 *                             + the phase loop of                    src/Optimiser.cpp:1162-1660
 *   thb_reconstruct_insert    the insert loop of reconstructRef      src/Optimiser.cpp:7036-7241
 *   thb_insert_counts         InsertFT with nC (3D classification)   gpu/interface/Interface.h:267-291, src/Optimiser.cpp:6862-6950
 *   thb_set_mode,             MODE_2D: ExpectGlobal2D, InsertI2D     gpu/interface/Interface.h:176-198, 239-265
 *   thb_insert_classes        (Projector / Reconstructor 2D twins)   src/Projector.cpp:337-354, src/Reconstructor.cpp:708-780
 *   thb_expect_scan_classes   ExpectGlobal2D (all classes at once)   gpu/interface/Interface.h:176-198, CPU loop src/Optimiser.cpp:756-914
 *   thb_pf_from_scan          the post-scan Particle logic           src/Optimiser.cpp:921-1075
 *   thb_symmetrize            Reconstructor::symmetrizeF/T/O         src/Reconstructor.cpp:2676-2716, include/Geometry/Transformation.h:105-194
 *   thb_norm_residual,        Optimiser::normCorrection              src/Optimiser.cpp:6201-6393
 *   thb_scale_images
 *   thb_upload_stack_at(_async), the per-call host pinning + H2D     gpu/src/cuthunder.cu:5370-5412
 *   thb_upload_wait, thb_download_stack / thb_reco_upload            (test / staging helpers of the same arrays)
 *   thb_create / thb_destroy  device selection and per-call set-up   gpu/interface/Interface.h:16 (getAviDevice), gpu/src/cuthunder.cu:5294-5324
 *   thb_device_count
 *   thb_last_error, thb_timer, thb_enable_timing, thb_kernel_ms, thb_launch_count, thb_set_option, thb_expect_stats,
 *   thb_synchronize, thb_version: no counterpart in the reference (error channel, accounting and tuning of this library)
 *
 * Data conventions (identical to the reference seam):
 *   complex = float[2] (re, im); packed image arrays are image-major [img][nPxl];
 *   Fourier volumes are FFTW r2c half-complex, x fastest, (vdim/2+1) x vdim x vdim,
 *   origin at index 0, negative y/z stored at +vdim; quaternions double[4] (w,x,y,z);
 *   translations double[2] in pixels; all host pointers are caller-owned.
 */
#ifndef THUNDER_B200_H
#define THUNDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct thb_ctx thb_ctx;

enum {
    THB_OK = 0,
    THB_E_ARG = -1,      /* invalid argument / call order */
    THB_E_CUDA = -2,     /* CUDA runtime error (message in thb_last_error) */
    THB_E_NOGPU = -3,    /* no usable sm_100 device: the product path refuses to run */
    THB_E_NCCL = -4,     /* NCCL missing or failed */
    THB_E_STATE = -5     /* object not initialised (volume / stack / pixels missing) */
};

/* stack kinds for thb_upload_stack */
enum { THB_STACK_EXPECT = 0 /* masked images, E pixel set */, THB_STACK_INSERT = 1 /* unmasked, M pixel set */ };

#define THB_UNIQUE_ID_BYTES 128

/* ---------------------------------------------------------------- library / context */
int thb_version(void);
int thb_device_count(void);                       /* CUDA devices visible; 0 if none */
int thb_create(thb_ctx** out, int device);        /* fails with THB_E_NOGPU without an sm_100 GPU */
void thb_destroy(thb_ctx* ctx);
const char* thb_last_error(const thb_ctx* ctx);   /* ctx may be NULL: last error of thb_create */
int thb_synchronize(thb_ctx* ctx);
/* counters for bench.py: kernels launched by this library since the last reset */
int64_t thb_launch_count(thb_ctx* ctx, int reset);
/* CUDA-event time (ms) accumulated per kernel family since the last reset:
 * which = 0 expect, 1 insert, 2 particle filter, 3 pack/unpack, 4 allreduce ; launches returned in *n */
double thb_kernel_ms(thb_ctx* ctx, int which, int64_t* n, int reset);
int thb_enable_timing(thb_ctx* ctx, int on);
/* tuning switches (development / A-B measurement).  key "expect_impl": 7 = several rotations per lane, lockstep launch on the
 * radial pixel order (default; 0 selects the default), 3 = direct gather from the cell layout, one rotation per lane,
 * 2 = TMA-staged box, 1 = direct gather from the linear layout, 4 / 5 = paired-lane / pixels-on-lanes variants;
 * "expect_rpl" (2 | 4 rotations per lane), "expect_order" (0 = 8x8 blocks, 1 = radial; takes effect at the next
 * thb_set_expect_pixels and drops the resident E stack), "expect_lock" / "expect_lock_tiles" / "expect_lock_window" (tile
 * barriers of the lockstep launch), "scan_templates" (scans project each shared rotation once per launch), "pf_stage" (particle
 * filter state staged in shared memory), "quad_oct", "quad_brick", "sort_rot", "expect_minb", "expect_spread", "insert_impl",
 * "insert_slab_mb", "stats": see DESIGN.md.  Unknown key -> THB_E_ARG. */
int thb_set_option(thb_ctx* ctx, const char* key, int value);
/* staging counters of the E kernel since the last reset (enable with option "stats" = 1): tiles, tiles with a
 * shared-memory box, sum of margins, staged elements, (rotation,tile) pairs on the L1/L2 path, all pairs,
 * over-capacity tiles, box rows */
int thb_expect_stats(thb_ctx* ctx, uint64_t out[16], int reset);   /* [8..15]: cycles per kernel phase, summed over CTAs */
/* CUDA-event stopwatch on the library's launch stream: stop == 0 records the start, stop != 0 records the
 * end, waits for it and returns the elapsed device time in *ms */
int thb_timer(thb_ctx* ctx, int stop, float* ms);

/* ---------------------------------------------------------------- a1: pixel list (host integer math) */
/* Returns nPxl (>= 0) or a negative error.  Any output pointer may be NULL.  Arrays must hold
 * (N/2+1)*N ints.  Order and rounding follow the reference loop exactly. */
int thb_pixel_list(int N, int pf, float rU, float rL, int* iCol, int* iRow, int* iPxl, int* iSig,
                   int* iColPad, int* iRowPad);

/* ---------------------------------------------------------------- geometry of the problem */
/* E pixel set: unpadded iCol/iRow (the library multiplies by pf, as Projector::project does). */
int thb_set_expect_pixels(thb_ctx* ctx, int N, int pf, int nPxl, const int* iCol, const int* iRow);
/* M pixel set: padded iColPad/iRowPad (already multiplied by pf, as Reconstructor::insertP expects)
 * and N (image size, for the translation phase ramp). */
int thb_set_insert_pixels(thb_ctx* ctx, int N, int pf, int nPxl, const int* iColPad, const int* iRowPad);

/* Projector volume for slot (class x half-set), half-complex (vdim/2+1) x vdim x vdim complex64. */
int thb_set_volume(thb_ctx* ctx, int slot, const float* volFT, int vdim);
/* MODE_2D (2D classification, demo_2D.json): the references are padded class-average FTs [vdim][vdim/2+1] (thb_set_volume with
 * the same arguments; slot = class), the rotations of every entry point are in-plane and passed as [..][2] = (cos, sin) - the
 * first two components of the reference's Particle rotation, rotate2D(dmat22&, dvec2) src/Geometry/Euler.cpp:125-131, and the
 * layout ExpectGlobal2D / InsertI2D receive (gpu/interface/Interface.h:176-198, 239-264) - the accumulators are images
 * (thb_reco_download returns F, T as [vdim][vdim/2+1], O as (ox, oy, 0)), thb_expect_scan compares EVERY image with class
 * `slot`, and thb_insert_classes scatters each draw into the accumulator of its own class.  Same kernels as MODE_3D:
 * bilinear gather / scatter = the trilinear cell of a two-plane volume at z = 0 (Projector::project src/Projector.cpp:337-354,
 * Image::getByInterpolationFT src/Image/Image.cpp:345-368, Reconstructor::insertP src/Reconstructor.cpp:708-780).
 * thb_reconstruct / thb_set_projectee work on images in this mode (the MODE_2D branches of Reconstructor::reconstruct,
 * Projector::setProjectee(Image) src/Projector.cpp:97-121: dstReal / volReal are N x N).  The device particle filter and
 * thb_symmetrize are MODE_3D only.  Switching the mode drops volumes and accumulators. */
enum { THB_MODE_3D = 0, THB_MODE_2D = 1 };
int thb_set_mode(thb_ctx* ctx, int mode);
int thb_get_mode(const thb_ctx* ctx);
int thb_get_volume(thb_ctx* ctx, int slot, float* volFT);      /* round trip of the device layout */

/* Resident packed stack, image-major [nImg][nPxl].  sigRcp may be NULL for THB_STACK_INSERT.
 * slotOfImg[nImg] (may be NULL = all 0) says which volume slot / accumulator each image uses. */
int thb_upload_stack(thb_ctx* ctx, int kind, int nImg, const float* dat, const float* ctf,
                     const float* sigRcp, const int* slotOfImg);
/* The same in pieces: reserve HBM for `capacity` images, then fill images [base, base+nImg) - lets a
 * caller stream a stack that is larger than its host buffers (100k x 256^2 = 70 GB) batch by batch. */
int thb_stack_reserve(thb_ctx* ctx, int kind, int capacity);
int thb_upload_stack_at(thb_ctx* ctx, int kind, int base, int nImg, const float* dat, const float* ctf,
                        const float* sigRcp, const int* slotOfImg);
/* The same on a second stream, returning at once, so that the upload of the next batch overlaps the kernels of
 * the current one (the reference pins the caller's arrays and streams them per call, gpu/src/cuthunder.cu:5370-5412).
 * Contract: page-locked host arrays that stay valid, and images [base, base+nImg) untouched by any other call,
 * until thb_upload_wait() has returned; thb_upload_wait also orders all later work on the context after the upload. */
int thb_upload_stack_at_async(thb_ctx* ctx, int kind, int base, int nImg, const float* dat, const float* ctf,
                              const float* sigRcp, const int* slotOfImg);
int thb_upload_wait(thb_ctx* ctx);

/* a2 on the device - Optimiser::allocPreCal (src/Optimiser.cpp:8043-8171, image-major, OPTIMISER_CTF_ON_THE_FLY):
 * fill images [base, base+nImg) of a reserved stack from their full half-complex FTs (_img for the E stack, _imgOri
 * for the M stack; imgFT[nImg][(N/2+1)*N] complex64, FFTW layout):  dat = img[iPxl], sigRcp = sigRcpTab[group][iSig]
 * (E stack only; the reference's _sigRcp = -0.5 / sigma^2, nGroup x nRing), ctf = CTF() of src/CTF.cpp:118-151 from
 * ctfAttr[nImg][7] = {voltage, defocusU, defocusV, defocusTheta, Cs, amplitudeContrast, phaseShift}.
 * iPxl / iSig come from thb_pixel_list; groupOfImg (0-based) may be NULL = group 0. */
int thb_pack_stack(thb_ctx* ctx, int kind, int base, int nImg, const float* imgFT, const int* iPxl, const int* iSig,
                   const float* sigRcpTab, int nGroup, int nRing, const int* groupOfImg, const float* ctfAttr,
                   float pixelSize, const int* slotOfImg);
/* a resident stack back in the caller's pixel order: dat[nImg][nPxl][2], ctf, sigRcp (any may be NULL) */
int thb_download_stack(thb_ctx* ctx, int kind, int base, int nImg, float* dat, float* ctf, float* sigRcp);

/* ---------------------------------------------------------------- a4/a5: slice extraction */
/* dst[nRot][nPxl] complex64 (host) = Projector::project for each rotation (quat[nRot][4]) */
int thb_project(thb_ctx* ctx, int slot, int nRot, const double* quat, float* dst);

/* ---------------------------------------------------------------- a3-a8: fused E kernel, local-search shape */
/* For each of nAct images (imgIdx into the THB_STACK_EXPECT stack): nR rotations x nT translations.
 *   quat[nAct][nR][4], tran[nAct][nT][2], wR[nAct][nR], wT[nAct][nT]   (prior weights, double)
 * outputs (host, any may be NULL):
 *   uR[nAct][nR], uT[nAct][nT], uC[nAct]   marginal weights relative to the per-image maximum
 *   base[nAct]                             the maximum log-likelihood (the "baseline")
 *   logL[nAct][nR][nT]                     raw log-likelihoods                                        */
int thb_expect_local(thb_ctx* ctx, int nAct, const int* imgIdx, int nR, int nT, const double* quat,
                     const double* tran, const double* wR, const double* wT, float* uR, float* uT,
                     float* uC, float* base, float* logL);

/* ---------------------------------------------------------------- a6/a8 with the defocus dimension: CTF search */
/* SEARCH_TYPE_CTF (src/Optimiser.cpp:1159-1273, 1328-1402; the reference's seam: ExpectPrefre, the defO of ExpectLocalP, the
 * k1 / k2 / phaseShift / conT of ExpectLocalPreI3D, the dpara / oldD of ExpectLocalRTD, the wD of ExpectLocalM,
 * gpu/interface/Interface.h:24-164).  The CTF of every defocus factor d is computed on the fly,
 *   ki = K1 defocusP d f^2 + K2 f^4 - phaseShift ,  ctf = -sqrt(1 - ac^2) sin ki + ac cos ki      (src/Optimiser.cpp:1253-1268)
 * from what allocPreCal prepares (:8125-8168): freQ[nPxl] = |k| / (N pixelSize) (thb_set_frequency, E pixel order of the caller),
 * defP[nImg][nPxl] = the per-pixel defocus of every image (thb_upload_stack_defocus, images [base, base + nImg) of the E stack),
 * ctfK[nAct][4] = {K1, K2, phaseShift, amplitudeContrast}.  dpar[nAct][nD] defocus factors, wD[nAct][nD] their prior weights;
 * uR / uT / uD / uC carry the prior weights of the other dimensions as the reference's wR / wT / wD / wC do (:1383-1402);
 * logL[nAct][nR][nT][nD] optional. */
int thb_set_frequency(thb_ctx* ctx, const float* freQ);
int thb_upload_stack_defocus(thb_ctx* ctx, int base, int nImg, const float* defP);
int thb_expect_local_ctf(thb_ctx* ctx, int nAct, const int* imgIdx, int nR, int nT, int nD, const double* quat, const double* tran,
                         const double* dpar, const double* wR, const double* wT, const double* wD, const float* ctfK, float* uR,
                         float* uT, float* uD, float* uC, float* base, float* logL);

/* ---------------------------------------------------------------- a7: global scan shape */
/* One shared set of nR rotations x nT translations against every image of the E stack that uses
 * `slot`.  pR[nR], pT[nT] prior weights.  Outputs (host): wC[nImg], wR[nImg][nR], wT[nImg][nT],
 * base[nImg]; logL[nImg][nR][nT] optional. Images of other slots get zeros. */
int thb_expect_scan(thb_ctx* ctx, int slot, int nR, int nT, const double* quat, const double* tran,
                    const double* pR, const double* pT, float* wC, float* wR, float* wT, float* base,
                    float* logL);
/* the same for the images [imgBase, imgBase + nImg) of the E stack (a batch of a larger resident stack); output rows are
 * relative to imgBase.  The whole rotation set goes through ONE launch per chunk of images (passes of 128 rotations inside
 * the kernel, baseline over the whole table): no merging of partial results on the host. */
int thb_expect_scan_range(thb_ctx* ctx, int slot, int imgBase, int nImg, int nR, int nT, const double* quat, const double* tran,
                          const double* pR, const double* pT, float* wC, float* wR, float* wT, float* base, float* logL);
/* MODE_2D classification scan of ALL classes (slots 0 .. nK-1) at once - the shape of ExpectGlobal2D (gpu/interface/Interface.h:176-198,
 * caller src/Optimiser.cpp:1873-1920; CPU loop :756-914): quat[nR][2] = (cos, sin) of the shared in-plane rotations, tran[nT][2]; every
 * image of [imgBase, imgBase + nImg) against every class, ONE baseline per image across the classes.  Outputs (host): wC[nImg][nK],
 * wR[nK][nImg][nR], wT[nK][nImg][nT], base[nImg] - the layouts thb_pf_from_scan takes.  The templates of all (class, rotation) pairs
 * are projected once per call, the pixel records of an image are built once for all classes. */
int thb_expect_scan_classes(thb_ctx* ctx, int nK, int imgBase, int nImg, int nR, int nT, const double* quat, const double* tran,
                            const double* pR, const double* pT, float* wC, float* wR, float* wT, float* base);

/* ---------------------------------------------------------------- a11-a14: fused M kernel */
int thb_reco_alloc(thb_ctx* ctx, int slot, int vdimPad);   /* accumulators F,T: (vdimPad/2+1) x vdimPad^2 */
int thb_reco_reset(thb_ctx* ctx, int slot);
/* Insert nImg images of the THB_STACK_INSERT stack (imgIdx may be NULL = 0..nImg-1):
 *   w[nImg] (already divided by mReco), offS[nImg][2] (may be NULL = 0),
 *   nr[nImg][mReco][4] quaternions, nt[nImg][mReco][2] translations.
 * Each image goes to the accumulator of its slot. */
int thb_insert(thb_ctx* ctx, int nImg, const int* imgIdx, int mReco, const float* w, const double* offS,
               const double* nr, const double* nt);
/* F[(vdimPad/2+1)*vdimPad^2][2], T[...] real part, O[3], counter.  normalise != 0 applies
 * sf = 1/T[0] to T and F (RECONSTRUCTOR_NORMALISE_T_F). Any pointer may be NULL. */
/* 3D classification (k > 1; the nC of InsertFT, gpu/interface/Interface.h:267-291, as reconstructRef fills it,
 * src/Optimiser.cpp:6862-6950): nr / nt hold mReco = the largest per-image count of draws of this class, nDraw[nImg] says how
 * many of them each image really has (0 .. mReco); only those are inserted and counted */
int thb_insert_counts(thb_ctx* ctx, int nImg, const int* imgIdx, int mReco, const float* w, const double* offS,
                      const int* nDraw, const double* nr, const double* nt);
/* CTF search (cSearch of InsertFT, src/Optimiser.cpp:7171-7215): nd[nImg][mReco] defocus factor of every draw; its CTF is
 * CTF(pixelSize, voltage, defocusU d, defocusV d, theta, Cs, amplitudeContrast, phaseShift) (src/CTF.cpp:118-151) computed on the
 * fly from ctfAttr[nImg][7] (the layout of thb_pack_stack) instead of the resident ctf array.  MODE_3D, mReco <= 256. */
int thb_insert_ctf(thb_ctx* ctx, int nImg, const int* imgIdx, int mReco, const float* w, const double* offS, const double* nr,
                   const double* nt, const double* nd, const float* ctfAttr, float pixelSize);
/* MODE_2D: the same with the class of every draw, nc[nImg][mReco] (InsertI2D's nC): the accumulator slot of the draw */
int thb_insert_classes(thb_ctx* ctx, int nImg, const int* imgIdx, int mReco, const float* w, const double* offS,
                       const int* nc, const double* nr, const double* nt);
int thb_reco_download(thb_ctx* ctx, int slot, float* F, float* T, double* O, int* counter, int normalise);

/* ---------------------------------------------------------------- a15: half-map allreduce */
int thb_comm_unique_id(char id[THB_UNIQUE_ID_BYTES]);
int thb_comm_init(thb_ctx* ctx, int nRanks, int rank, const char id[THB_UNIQUE_ID_BYTES]);
/* One sum-allreduce of every allocated accumulator (F|T interleaved, all slots) + O, counter. */
int thb_allreduce(thb_ctx* ctx);

/* ---------------------------------------------------------------- f1: reconstruct + setProjectee */
/* set the accumulators of a slot from host arrays (restart / tests): F[(m/2+1)*m*m][2], T[(m/2+1)*m*m] */
int thb_reco_upload(thb_ctx* ctx, int slot, const float* F, const float* T);
/* Reconstructor::reconstruct (src/Reconstructor.cpp:1129-1831, MODE_3D, C1, default Config.h) on the accumulators of
 * `slot` (after thb_insert / thb_allreduce), edge m = pf * size, size <= N:
 *   normalise != 0: the prepareTF normalisation first (sf = 1 / T[0], src/Reconstructor.cpp:1056-1091);
 *   fsc != NULL: MAP weighting of T by the half-map FSC (fsc[nFsc] per shell of the unpadded grid), joinHalf as setJoinHalf;
 *   gridCorr: the iterative gridding correction (at most 30 pairs of 3D FFTs, cuFFT), else W = 1 / T;
 *   then F * W -> inverse 3D FFT on the (pf N)^3 grid -> central N^3 -> / sinc^2.
 * The accumulators are consumed (scaled / weighted in place, as the reference's _T3D / _F3D are).  The result stays on
 * the device for thb_set_projectee and is copied to dstReal[N][N][N] (origin at index 0) when dstReal != NULL. */
int thb_reconstruct(thb_ctx* ctx, int slot, int N, int pf, double a, double alpha, int gridCorr, int joinHalf,
                    const float* fsc, int nFsc, int normalise, float* dstReal, int* nIterOut);
/* Projector::setProjectee (src/Projector.cpp:123-148, gridCorrection :573-583): real volume N^3 -> zero-padded
 * (pf N)^3 -> / sinc^2 -> forward 3D FFT into projector slot `slot`.  volReal == NULL: the result of the last
 * thb_reconstruct (no host round trip). */
int thb_set_projectee(thb_ctx* ctx, int slot, const float* volReal, int N, int pf);

/* ---------------------------------------------------------------- f2: re-centre + re-mask + pack */
/* Optimiser::reCentreImg + reMaskImg (src/Optimiser.cpp:6065-6151) + allocPreCal on the device: E-stack images
 * [base, base+nImg) rebuilt from the ORIGINAL image FTs imgOriFT[nImg][(N/2+1)*N]:
 *   translate by offset[nImg][2] (the running _offset, already updated by the caller)  ->  2D c2r (batched cuFFT)  ->
 *   x soft mask (radius maskRadiusPx = maskRadius / pixelSize, EDGE_WIDTH_RL = 6; skipped when zeroMask == 0)  ->  2D r2c  ->
 *   the packing of thb_pack_stack.  imgOutFT (may be NULL) receives the masked image FTs (the reference's _img). */
int thb_remask_pack(thb_ctx* ctx, int base, int nImg, const float* imgOriFT, const double* offset, float maskRadiusPx,
                    int zeroMask, const int* iPxl, const int* iSig, const float* sigRcpTab, int nGroup, int nRing,
                    const int* groupOfImg, const float* ctfAttr, float pixelSize, const int* slotOfImg, float* imgOutFT);

/* ---------------------------------------------------------------- f3: sigma^2 refresh */
/* The image loop of Optimiser::allReduceSigma (src/Optimiser.cpp:6428-6600; OPTIMISER_SIGMA_RANK1ST, MODE_3D, no CTF search) on
 * the resident stacks: for image l (imgIdx[l] or l) with its best orientation quat[l][4], translation tran[l][2] and running
 * offset offS[l][2] (NULL = 0), the ring-averaged power of  masked image - ctf * translated slice  (E stack, translation
 * tran), of  original image - ctf * translated slice  (M stack, translation tran - offS) and the signal / data spectra of the
 * SVD ratio, summed per group.  iSigE / iSigM: the ring of every pixel of the two pixel lists (thb_pixel_list).  The sigma
 * pixel set is {|k|^2 < rSig^2, rint|k| < rSig}; rings the E list does not reach (below its rL) stay zero.  The packed lists
 * hold one of each Hermitian pair on the i = 0 column (src/Optimiser.cpp:8015) while powerSpectrum() walks the whole half
 * plane: those pixels are counted twice, which is exact for the FTs of real images.
 * Outputs [nGroup][rSig + 1] doubles, last column = weight sum: exactly the matrices the reference all-reduces over the
 * hemisphere; the final normalisation and mixing (:6651-6709) are a few hundred flops on the caller's side. */
int thb_sigma_accumulate(thb_ctx* ctx, int nImg, const int* imgIdx, const double* quat, const double* tran, const double* offS,
                         const int* groupOfImg, int nGroup, int rSig, const int* iSigE, const int* iSigM, double* sigM,
                         double* sigN, double* svd);

/* Point-group symmetrisation of the accumulators after the all-reduce, as prepareTF does (Reconstructor::symmetrizeF / T / O,
 * src/Reconstructor.cpp:2676-2716 = SYMMETRIZE_FT, include/Geometry/Transformation.h:105-194): F, T += their trilinear
 * interpolation at R_e v for each of the nElem symmetry elements (R[nElem][9], column-major dmat33 as Symmetry::get(L, R, i)
 * returns them; C1: nElem = 0), for rotated positions inside `radius` (the reference passes maxRadius * pf + 1);
 * O += sum_e R_e O; counter *= 1 + nElem.  MODE_3D. */
int thb_symmetrize(thb_ctx* ctx, int slot, int nElem, const double* R, double radius);

/* Optimiser::normCorrection (src/Optimiser.cpp:6201-6393, OPTIMISER_NORM_MASK as in include/Config.h:171), image loop:
 * norm[l] = sum over the half plane {rL^2 <= |k|^2 < rNorm^2} of |masked image - ctf * slice at quat[l] translated by tran[l]|^2
 * on the resident E stack (its pixel list must reach rNorm).  The median over ALL particles (an MPI all-reduce in the reference,
 * :6352-6369) and s_l = sqrt(median / norm_l) stay with the caller; thb_scale_images applies s_l to the images of both
 * resident stacks (_img[l] *= s, _imgOri[l] *= s, :6381-6391). */
int thb_norm_residual(thb_ctx* ctx, int nImg, const int* imgIdx, const double* quat, const double* tran, float rL, float rNorm,
                      double* norm);
int thb_scale_images(thb_ctx* ctx, int nImg, const int* imgIdx, const float* scale);

/* ---------------------------------------------------------------- a9: device-resident particle filter */
typedef struct thb_pf_params {
    int mLR, mLT;               /* support sizes (rotation, translation) */
    double transS, transQ;      /* translation prior sigma and the re-centre quantile */
    double perturbFactorL, perturbFactorS;
    int minPhase, maxPhase;     /* MIN_N_PHASE_PER_ITER_LOCAL, MAX_N_PHASE_PER_ITER */
    int fixedPhases;            /* > 0: run exactly this many phases (benchmark mode) */
    double decreaseFactor;      /* PARTICLE_FILTER_DECREASE_FACTOR */
    int noDecreaseLimit;        /* N_PHASE_WITH_NO_VARI_DECREASE */
    uint64_t seed;
    /* CTF search (SEARCH_TYPE_CTF): mLD > 0 adds the defocus dimension - Particle::initD(mLD, ctfRefineS) after the first
     * perturbation, perturb(perturbFactorSCTF, PAR_D) after the later ones, calRank1st / calVari / resample of PAR_D after those of
     * the rotations and translations (src/Optimiser.cpp:1193-1215, 1483-1488) - needs thb_pf_set_ctf, thb_set_frequency and
     * thb_upload_stack_defocus; mLD <= max(mLR, mLT).  0: no CTF search. */
    int mLD;
    double ctfRefineS, perturbFactorSCTF;
} thb_pf_params;

/* Initialise nPar particles from (quat[nPar][4], k1,k2,k3[nPar], tran[nPar][2], s0,s1[nPar]) -
 * Particle::load semantics (src/Particle.cpp:401-556): ACG cloud about quat, Gaussian cloud about tran. */
int thb_pf_load(thb_ctx* ctx, int nPar, const thb_pf_params* p, const double* quat, const double* k123,
                const double* tran, const double* s01);
/* From the global scan to the support of the local phases - the per-image logic that follows the scan in Optimiser::expectation
 * (src/Optimiser.cpp:921-1075), on the device: choice of the class (setUC, keepHalfHeightPeak with PEAK_FACTOR_C, resample(nK, PAR_C),
 * rand), likelihood weights of THAT class on the shared scan grid quat[nR][4] (MODE_2D: [nR][2] = (cos, sin)) / tran[nT][2],
 * setPeakFactor + keepHalfHeightPeak(PAR_R), resample(mLR, PAR_R), resample(mLT, PAR_T), calVari, floors kFloor on k1 (.. k3) and
 * sFloor on s0, s1 (OPTIMISER_SCAN_SET_MIN_STD_WITH_PERTURB).  Inputs as thb_expect_scan returns them, class by class:
 * wC[nPar][nK] (relative to ONE baseline per image: scale the per-class results to the largest baseline), wR[nK][nPar][nR],
 * wT[nK][nPar][nT].  The chosen class becomes the slot of the image in both resident stacks when nK > 1 (clsOut[nPar], may be NULL; with nK = 1 - a
 * refinement whose slots are the half sets - the slots are left alone), the
 * particles are the loaded ones (as after thb_pf_load; MODE_2D allowed: the phase loop then runs the von Mises operators).
 * A global-search iteration then continues with thb_expectation with perturbFactorL = perturbFactorSGlobal (its phases start
 * at 1 in the reference: there is no large first perturbation). */
int thb_pf_from_scan(thb_ctx* ctx, int nPar, const thb_pf_params* p, int nK, int nR, int nT, const double* quat, const double* tran,
                     const float* wC, const float* wR, const float* wT, double kFloor, double sFloor, int* clsOut);
/* CTF search: per particle the constants of the on-the-fly CTF of the E-step, ctfK[nPar][4] = {K1, K2, phaseShift,
 * amplitudeContrast} (allocPreCal, src/Optimiser.cpp:8163-8167), and the attributes the M-step computes the CTF of every draw
 * from, ctfAttr[nPar][7] (the layout of thb_pack_stack), with the pixel size.  Call after thb_pf_load. */
int thb_pf_set_ctf(thb_ctx* ctx, const float* ctfK, const float* ctfAttr, float pixelSize);
/* defocus factors of the loaded particles: d[nPar][mLD + 1] (the last one = the most likely, _topD), wD[nPar][mLD], sD[nPar] */
int thb_pf_get_d(thb_ctx* ctx, double* d, double* wD, double* sD);
/* Read back particle state.  Any pointer may be NULL.
 *   r[nPar][mLR][4], t[nPar][mLT][2], wR[nPar][mLR], wT[nPar][mLT],
 *   scal[nPar][20] = k1,k2,k3,s0,s1,rho,topR[4],topT[2],score,nPhase,variR,variT,peakFactorR,
 *                    noDecreaseCount,variD,spare */
int thb_pf_get(thb_ctx* ctx, double* r, double* t, double* wR, double* wT, double* scal);
int thb_pf_set(thb_ctx* ctx, const double* r, const double* t, const double* wR, const double* wT,
               const double* scal);
/* particle p <-> image imgBase + p of the resident stacks; streamBase offsets the per-particle random
 * stream (use the global index of the first particle so that ranks / batches draw different numbers) */
int thb_pf_set_image_base(thb_ctx* ctx, int imgBase, uint64_t streamBase);
/* Reproducibility.  The random numbers of particle p in a call are GSL 2.4's distributions (the ones the reference calls,
 * src/Particle.cpp, src/Geometry/DirectionalStat.cpp:39-62) over a Philox4x32-10 bit stream keyed by (seed, streamBase + p,
 * epoch): thb_pf_load, thb_expectation, thb_reconstruct_insert and thb_pf_op each advance the context's epoch counter e by one
 * and use the key (e << 20) (+ phase + 1 for the operators after the likelihoods of a phase).  thb_pf_set_epoch sets e, so that
 * a run - or the reference's Particle class with the same bit generator plugged in (as the parity tests do) - can replay it. */
int thb_pf_set_epoch(thb_ctx* ctx, uint64_t epoch);
/* keep the marginal weights uR / uT of the first nPhases phases of the following thb_expectation calls (0 = off);
 * thb_pf_get_trace copies them out: uR[nPhases][nPar][mLR], uT[nPhases][nPar][mLT], base[nPhases][nPar] (the largest
 * log-likelihood of the phase, which the weights are relative to); any pointer may be NULL */
int thb_pf_trace(thb_ctx* ctx, int nPhases);
int thb_pf_get_trace(thb_ctx* ctx, int nPhases, float* uR, float* uT, float* base);
/* and the traced supports: st[nPhases][2][nPar][4 mLR + 2 mLT] = (r[mLR][4], t[mLT][2]) after the perturbation ([..][0]) and after
 * the resampling ([..][1]) of every traced phase */
int thb_pf_get_trace_states(thb_ctx* ctx, int nPhases, double* st);
/* the support indices drawn by the last thb_reconstruct_insert: drawR/drawT [nPar][mReco] */
int thb_pf_get_draws(thb_ctx* ctx, int mReco, int* drawR, int* drawT);
int thb_pf_get_draws_d(thb_ctx* ctx, int mReco, int* drawD);      /* CTF search: the defocus-factor index of every draw */
/* E-step of one iteration over all loaded particles (particle p <-> image imgBase + p of the E stack):
 * phase loop of Optimiser::expectation with the particle filter on the device. */
int thb_expectation(thb_ctx* ctx, int* nPhaseOut /* [nPar] or NULL */);
/* M-step insert loop of reconstructRef: mReco uniform draws per particle from its support
 * (Particle::rand), w = 1/mReco (or compressR/mReco when parGra != 0). */
int thb_reconstruct_insert(thb_ctx* ctx, int mReco, int parGra, const double* offS);

/* single particle-filter operators on the device state, for parity tests against Particle::* */
int thb_pf_op(thb_ctx* ctx, int op, double arg, const float* uR, const float* uT);
enum {
    THB_PF_PERTURB_R = 1, THB_PF_PERTURB_T = 2, THB_PF_SET_U_KEEP_PEAK = 3, THB_PF_RANK1ST = 4,
    THB_PF_CALVARI = 5, THB_PF_RESAMPLE = 6, THB_PF_BALANCE_R = 7, THB_PF_BALANCE_T = 8
};

#ifdef __cplusplus
}
#endif
#endif /* THUNDER_B200_H */
