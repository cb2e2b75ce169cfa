// oracle/ref_harness.cpp - extern "C" driver around the REFERENCE's own CPU classes.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/build_ref.sh together with the reference
// sources (where they lie under /root/reference) into oracle/_ref/libthunder_ref.so.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load that library.  Nothing here is product code and nothing here re-implements the
// reference's arithmetic: every numeric result comes out of a reference function
// (Projector::project, translate, CTF, logDataVSPrior_*, Reconstructor::insertP/insertDir/
// prepareTF, Particle::*, rotate3D).  The two "loop" entry points at the bottom reproduce
// only the *driver loops* of Optimiser::expectation (reference src/Optimiser.cpp:1162-1660)
// and Optimiser::reconstructRef (src/Optimiser.cpp:7036-7241), because the Optimiser object
// itself cannot run without >=3 MPI ranks, a .thu database and MRC stacks.

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <map>
#include <omp.h>
#include <pthread.h>

#define private public
#define protected public
#include "Projector.h"
#include "Reconstructor.h"
#include "Particle.h"
#include "Optimiser.h"
#undef private
#undef protected

#include "CTF.h"
#include "FFT.h"
#include "Spectrum.h"
#include "Symmetry.h"
#include "ImageFunctions.h"
#include "Euler.h"
#include "Random.h"
#include "DirectionalStat.h"

INITIALIZE_EASYLOGGINGPP

// Free functions defined in the reference's src/Optimiser.cpp and forward-declared only there
// (src/Optimiser.cpp:13-22); re-declared so the harness links against the reference objects.
RFLOAT logDataVSPrior_m_huabin(const Complex* dat, const Complex* pri, const RFLOAT* ctf,
                               const RFLOAT* sigRcp, const int m);
RFLOAT logDataVSPrior_m_huabin_SIMD256(Complex* dat, const Complex* pri, const RFLOAT* ctf,
                                       const RFLOAT* sigRcp, const int m);
RFLOAT* logDataVSPrior_m_n_huabin(const Complex* dat, const Complex* pri, const RFLOAT* ctf,
                                  const RFLOAT* sigRcp, const int n, const int m, RFLOAT* result);
RFLOAT* logDataVSPrior_m_n_huabin_SIMD256(Complex* dat, const Complex* pri, const RFLOAT* ctf,
                                          const RFLOAT* sigRcp, const int n, const int m, RFLOAT* result);

// ------------------------------------------------------------------------------------------
// Deterministic replacement for the reference's src/Functions/Random.cpp (urandom-seeded,
// :51-100).  Same engine type (gsl_rng_mt19937), one engine per thread, seed = base + tid.
// ------------------------------------------------------------------------------------------
static unsigned long g_seed_base = 20240229UL;
static unsigned long g_seed_epoch = 1;
struct TlsRng { gsl_rng* eng; unsigned long epoch; };
static __thread TlsRng t_rng = {NULL, 0};

// ------------------------------------------------------------------------------------------
// Replay bit generator.  To compare the device particle filter with the reference's Particle class draw by draw, the
// reference's engine can be swapped for a counter-based Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random
// numbers: as easy as 1, 2, 3", SC'11) keyed by (seed, stream = particle, epoch = call / phase) - the bit generator the
// CUDA library uses - wrapped as a gsl_rng_type with the range of mt19937 (0 .. 2^32 - 1, get_double = get / 2^32), so that
// every GSL distribution the reference calls (gsl_ran_gaussian, gsl_ran_bivariate_gaussian, gsl_ran_flat, gsl_ran_shuffle,
// gsl_rng_uniform_int) runs unchanged on top of it.  Restated here from the published algorithm, independent of the product.
// ------------------------------------------------------------------------------------------
struct PhiloxState { uint32_t k0, k1, c[4], o[4]; int have; };

static void philox_block(PhiloxState* p)
{
    uint32_t x0 = p->c[0], x1 = p->c[1], x2 = p->c[2], x3 = p->c[3], a = p->k0, b = p->k1;
    for (int r = 0; r < 10; r++)
    {
        const uint64_t m0 = (uint64_t)0xD2511F53u * x0, m1 = (uint64_t)0xCD9E8D57u * x2;
        const uint32_t y0 = (uint32_t)(m1 >> 32) ^ x1 ^ a, y1 = (uint32_t)m1, y2 = (uint32_t)(m0 >> 32) ^ x3 ^ b, y3 = (uint32_t)m0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    p->o[0] = x0; p->o[1] = x1; p->o[2] = x2; p->o[3] = x3;
    p->c[0]++;
    p->have = 4;
}
static void philox_key(PhiloxState* p, unsigned long long seed, unsigned long long stream, unsigned long long epoch)
{
    p->k0 = (uint32_t)seed; p->k1 = (uint32_t)(seed >> 32);
    p->c[0] = 0; p->c[1] = (uint32_t)epoch; p->c[2] = (uint32_t)stream;
    p->c[3] = (uint32_t)(stream >> 32) ^ (uint32_t)(epoch >> 32);
    p->have = 0;
}
static void philox_set(void* st, unsigned long int seed) { philox_key((PhiloxState*)st, seed, 0, 0); }
static unsigned long int philox_get(void* st)
{
    PhiloxState* p = (PhiloxState*)st;
    if (p->have == 0) philox_block(p);
    return p->o[--p->have];            // words of a block are handed out last to first
}
static double philox_get_double(void* st) { return philox_get(st) / 4294967296.0; }
static const gsl_rng_type philox_type = {"thb_philox4x32_10", 0xffffffffUL, 0, sizeof(PhiloxState), &philox_set, &philox_get,
                                         &philox_get_double};

static int g_replay = 0;
static unsigned long long g_rp_seed = 0, g_rp_stream = 0, g_rp_epoch = 0, g_rp_stride = 1;   // keys of the loops below (stream + stride * image index)
static __thread gsl_rng* t_replay = NULL;

gsl_rng* get_random_engine()
{
    if (g_replay)
    {
        if (!t_replay) t_replay = gsl_rng_alloc(&philox_type);
        return t_replay;
    }
    if (!t_rng.eng) t_rng.eng = gsl_rng_alloc(gsl_rng_mt19937);
    if (t_rng.epoch != g_seed_epoch)
    {
        gsl_rng_set(t_rng.eng, g_seed_base + (unsigned long)omp_get_thread_num());
        t_rng.epoch = g_seed_epoch;
    }
    return t_rng.eng;
}

static inline void replay_key(unsigned long long seed, unsigned long long stream, unsigned long long epoch)
{
    if (!t_replay) t_replay = gsl_rng_alloc(&philox_type);
    philox_key((PhiloxState*)t_replay->state, seed, stream, epoch);
}

static void quiet_loggers()
{
    el::Configurations conf;
    conf.setToDefault();
    conf.set(el::Level::All, el::ConfigurationType::ToFile, "false");
    conf.set(el::Level::All, el::ConfigurationType::ToStandardOutput, "false");
    conf.set(el::Level::All, el::ConfigurationType::Enabled, "false");
    conf.set(el::Level::Fatal, el::ConfigurationType::Enabled, "true");
    conf.set(el::Level::Fatal, el::ConfigurationType::ToStandardOutput, "true");
    el::Loggers::setDefaultConfigurations(conf, true);
    const char* names[] = {"LOGGER_SYS", "LOGGER_INIT", "LOGGER_ROUND", "LOGGER_COMPARE", "LOGGER_RECO",
                           "LOGGER_MPI", "LOGGER_FFT", "LOGGER_GPU", "LOGGER_MEM", "LOGGER"};
    for (size_t i = 0; i < sizeof(names) / sizeof(*names); ++i) el::Loggers::getLogger(names[i]);
    el::Loggers::reconfigureAllLoggers(conf);
}

static inline dmat33 mat_from_colmajor(const double* m)
{
    dmat33 r;
    for (int c = 0; c < 3; c++)
        for (int rr = 0; rr < 3; rr++) r(rr, c) = m[c * 3 + rr];
    return r;
}

struct RefProjector
{
    Projector proj;
};

struct RefReco
{
    Reconstructor reco;
    std::vector<int> iCol, iRow, iPxl, iSig;
};

extern "C" {

int ref_init(int nThreadFFTW)
{
    static bool done = false;
    if (!done)
    {
        quiet_loggers();
        TSFFTW_init_threads();
        done = true;
    }
    (void)nThreadFFTW;
    return 0;
}

void ref_set_seed(unsigned long seed)
{
    g_seed_base = seed;
    g_seed_epoch++;
}

int ref_sizeof_rfloat() { return (int)sizeof(RFLOAT); }

// replay mode of the random engine (see PhiloxState above).  ref_rng_key keys the CALLING thread's engine (class-level calls
// from Python); ref_rng_replay_loop sets the keys the driver loops below use per image: stream + l, epoch
void ref_rng_replay(int on) { g_replay = on; }
void ref_rng_key(unsigned long long seed, unsigned long long stream, unsigned long long epoch) { replay_key(seed, stream, epoch); }
void ref_rng_replay_loop(unsigned long long seed, unsigned long long stream, unsigned long long epoch)
{
    g_rp_seed = seed; g_rp_stream = stream; g_rp_epoch = epoch;
}
void ref_rng_replay_stride(unsigned long long stride) { g_rp_stride = stride; }   // image l of a loop <-> stream + stride * l
// n raw draws of the calling thread's engine through GSL's own entry points (the pin of the product's restatement of them):
// kind 0 gsl_rng_uniform, 1 gsl_ran_gaussian(sigma = a), 2 gsl_rng_uniform_int(n = a), 3 gsl_ran_flat(a, b),
// 4 gsl_ran_bivariate_gaussian(a, b, rho = 0.3): x, y interleaved
void ref_rng_draw(int kind, int n, double a, double b, double* out)
{
    gsl_rng* e = get_random_engine();
    for (int i = 0; i < n; i++)
        switch (kind)
        {
            case 0: out[i] = gsl_rng_uniform(e); break;
            case 1: out[i] = gsl_ran_gaussian(e, a); break;
            case 2: out[i] = (double)gsl_rng_uniform_int(e, (unsigned long)a); break;
            case 3: out[i] = gsl_ran_flat(e, a, b); break;
            case 4: gsl_ran_bivariate_gaussian(e, a, b, 0.3, out + 2 * i, out + 2 * i + 1); break;
        }
}

// ---------------------------------------------------------------- pixel list
// Optimiser::allocPreCalIdx (reference src/Optimiser.cpp:7991-8041), run on a minimal
// Optimiser object (one empty N x N Fourier image, rank 1 of 3 so that IF_MASTER is false).
int ref_alloc_precal_idx(int N, int pf, float rU, float rL, int* iCol, int* iRow, int* iPxl, int* iSig,
                         int* iColPad, int* iRowPad)
{
    Optimiser* opt = new Optimiser();
    opt->setMPIEnv(3, 1, MPI_COMM_SELF, MPI_COMM_SELF);
    opt->_para.pf = pf;
    opt->_para.size = N;
    opt->_imgOri.push_back(Image(N, N, FT_SPACE));
    opt->allocPreCalIdx(rU, rL);
    int n = opt->_nPxl;
    if (iCol) memcpy(iCol, opt->_iCol, n * sizeof(int));
    if (iRow) memcpy(iRow, opt->_iRow, n * sizeof(int));
    if (iPxl) memcpy(iPxl, opt->_iPxl, n * sizeof(int));
    if (iSig) memcpy(iSig, opt->_iSig, n * sizeof(int));
    if (iColPad) memcpy(iColPad, opt->_iColPad, n * sizeof(int));
    if (iRowPad) memcpy(iRowPad, opt->_iRowPad, n * sizeof(int));
    opt->freePreCalIdx();
    // The Optimiser destructor touches members that were never initialised in this minimal
    // object; leak the (small) shell instead of running it.
    opt->_imgOri.clear();
    return n;
}

// ---------------------------------------------------------------- geometry / small helpers
void ref_rotate3D(const double* quat, double* mat9_colmajor)
{
    dmat33 m;
    rotate3D(m, dvec4(quat[0], quat[1], quat[2], quat[3]));
    memcpy(mat9_colmajor, m.data(), 9 * sizeof(double));
}

void ref_translate(float* dst, float tx, float ty, int N, const int* iCol, const int* iRow, int nPxl)
{
    translate((Complex*)dst, tx, ty, N, N, iCol, iRow, nPxl, 1);
}

void ref_translate_src(float* dst, const float* src, float tx, float ty, int N, const int* iCol,
                       const int* iRow, int nPxl)
{
    translate((Complex*)dst, (const Complex*)src, tx, ty, N, N, iCol, iRow, nPxl, 1);
}

void ref_ctf(float* dst, float pixelSize, float voltage, float defocusU, float defocusV, float theta, float Cs,
             float amplitudeContrast, float phaseShift, int N, const int* iCol, const int* iRow, int nPxl)
{
    CTF(dst, pixelSize, voltage, defocusU, defocusV, theta, Cs, amplitudeContrast, phaseShift, N, N, iCol, iRow,
        nPxl, 1);
}

// variant 0: scalar (src/Optimiser.cpp:9187-9213); 1: AVX256 (:9410-9471, default build)
float ref_logDataVSPrior(const float* dat, const float* pri, const float* ctf, const float* sigRcp, int m,
                         int variant)
{
    if (variant == 0)
        return logDataVSPrior_m_huabin((const Complex*)dat, (const Complex*)pri, ctf, sigRcp, m);
    return logDataVSPrior_m_huabin_SIMD256((Complex*)dat, (const Complex*)pri, ctf, sigRcp, m);
}

// pixel-major, n images against one template (src/Optimiser.cpp:9931-9973 / :9222-9306)
void ref_logDataVSPrior_m_n(const float* dat, const float* pri, const float* ctf, const float* sigRcp, int n,
                            int m, float* result, int variant)
{
    if (variant == 0)
        logDataVSPrior_m_n_huabin((const Complex*)dat, (const Complex*)pri, ctf, sigRcp, n, m, result);
    else
        logDataVSPrior_m_n_huabin_SIMD256((Complex*)dat, (const Complex*)pri, ctf, sigRcp, n, m, result);
}

// ---------------------------------------------------------------- Projector
void* ref_projector_create(int pf)
{
    RefProjector* p = new RefProjector();
    p->proj.setMode(MODE_3D);
    p->proj.setInterp(LINEAR_INTERP);
    p->proj.setPf(pf);
    return p;
}

void ref_projector_destroy(void* h) { delete (RefProjector*)h; }

// real-space N^3 volume -> FFT -> Projector::setProjectee (pad, grid correction, FFT;
// reference src/Projector.cpp:123-148)
void ref_projector_set_from_real(void* h, const float* volRL, int N, int nThread)
{
    RefProjector* p = (RefProjector*)h;
    Volume v(N, N, N, RL_SPACE);
    for (size_t i = 0; i < v.sizeRL(); i++) v(i) = volRL[i];
    FFT fft;
    fft.fw(v, nThread);
    v.clearRL();
    p->proj.setProjectee(v.copyVolume(), nThread);
}

// load an already padded half-complex Fourier volume (pfN/2+1) x pfN x pfN verbatim
void ref_projector_set_padded_ft(void* h, const float* volFT, int pfN)
{
    RefProjector* p = (RefProjector*)h;
    p->proj._projectee3D.alloc(pfN, pfN, pfN, FT_SPACE);
    memcpy(&p->proj._projectee3D[0], volFT, p->proj._projectee3D.sizeFT() * sizeof(Complex));
    p->proj._maxRadius = pfN / p->proj._pf / 2 - 1;
}

int ref_projector_padded_dim(void* h) { return (int)((RefProjector*)h)->proj._projectee3D.nSlcFT(); }

void ref_projector_get_padded_ft(void* h, float* out)
{
    RefProjector* p = (RefProjector*)h;
    memcpy(out, &p->proj._projectee3D[0], p->proj._projectee3D.sizeFT() * sizeof(Complex));
}

void ref_projector_set_max_radius(void* h, int r) { ((RefProjector*)h)->proj.setMaxRadius(r); }

// Projector::project(Complex*, const dmat33&, iCol, iRow, nPxl, nThread), src/Projector.cpp:356-374
void ref_projector_project(void* h, float* dst, const double* mat9_colmajor, const int* iCol, const int* iRow,
                           int nPxl)
{
    RefProjector* p = (RefProjector*)h;
    p->proj.project((Complex*)dst, mat_from_colmajor(mat9_colmajor), iCol, iRow, nPxl, 1);
}

// Projector::project(Image&, const dmat33&, const dvec2&, nThread), src/Projector.cpp:452-464: the whole-image form that
// Optimiser::allReduceSigma uses; out = the half-complex image [N][N/2+1]
void ref_projector_project_image(void* h, int N, const double* quat, const double* tran, float* out)
{
    RefProjector* p = (RefProjector*)h;
    Image img(N, N, FT_SPACE);
    SET_0_FT(img);
    dmat33 rot;
    rotate3D(rot, dvec4(quat[0], quat[1], quat[2], quat[3]));
    p->proj.project(img, rot, dvec2(tran[0], tran[1]), 1);
    memcpy(out, &img[0], (size_t)(N / 2 + 1) * N * sizeof(Complex));
}

// ---------------------------------------------------------------- Reconstructor
void* ref_reco_create(int size, int N, int pf, int nThread)
{
    RefReco* r = new RefReco();
    r->reco.setMPIEnv(3, 1, MPI_COMM_SELF, MPI_COMM_SELF);
    r->reco.init(MODE_3D, size, N, pf, NULL, 1.9, 15);
    r->reco.allocSpace(nThread);
    return r;
}

void ref_reco_destroy(void* h)
{
    RefReco* r = (RefReco*)h;
    r->reco.freeSpace();
    delete r;
}

void ref_reco_reset(void* h, int nThread) { ((RefReco*)h)->reco.reset(nThread); }

void ref_reco_set_precal(void* h, int nPxl, const int* iColPad, const int* iRowPad, const int* iPxl,
                         const int* iSig)
{
    RefReco* r = (RefReco*)h;
    r->iCol.assign(iColPad, iColPad + nPxl);
    r->iRow.assign(iRowPad, iRowPad + nPxl);
    r->iPxl.assign(iPxl, iPxl + nPxl);
    r->iSig.assign(iSig, iSig + nPxl);
    r->reco.setPreCal(nPxl, r->iCol.data(), r->iRow.data(), r->iPxl.data(), r->iSig.data());
}

// Reconstructor::insertP(const Complex*, const RFLOAT*, const dmat33&, RFLOAT, NULL)
// src/Reconstructor.cpp:782-863
void ref_reco_insertP(void* h, const float* src, const float* ctf, const double* mat9_colmajor, float w)
{
    ((RefReco*)h)->reco.insertP((const Complex*)src, ctf, mat_from_colmajor(mat9_colmajor), w, NULL);
}

void ref_reco_insertDir(void* h, double ox, double oy, double oz) { ((RefReco*)h)->reco.insertDir(ox, oy, oz); }

int ref_reco_pad_size(void* h) { return (int)((RefReco*)h)->reco._F3D.nSlcFT(); }

// F as complex64 half-volume, T as the real part of the reference's complex T volume
void ref_reco_get(void* h, float* F, float* T, double* O3, int* counter)
{
    RefReco* r = (RefReco*)h;
    size_t n = r->reco._F3D.sizeFT();
    if (F) memcpy(F, &r->reco._F3D[0], n * sizeof(Complex));
    if (T)
        for (size_t i = 0; i < n; i++) T[i] = REAL(r->reco._T3D[i]);
    if (O3)
    {
        O3[0] = r->reco._ox;
        O3[1] = r->reco._oy;
        O3[2] = r->reco._oz;
    }
    if (counter) *counter = r->reco._counter;
}

// Reconstructor::prepareTF (src/Reconstructor.cpp:1056-1091): allreduce (identity at one rank),
// normalise by 1/Re T[0], symmetrise (C1: identity)
void ref_reco_prepareTF(void* h, int nThread) { ((RefReco*)h)->reco.prepareTF(nThread); }

// set the accumulators directly (parity tests of reconstruct): F complex64, T real part
void ref_reco_set(void* h, const float* F, const float* T)
{
    RefReco* r = (RefReco*)h;
    size_t n = r->reco._F3D.sizeFT();
    memcpy(&r->reco._F3D[0], F, n * sizeof(Complex));
    for (size_t i = 0; i < n; i++) r->reco._T3D[i] = COMPLEX(T[i], 0);
}

// Reconstructor::reconstruct(Volume&, nThread) (src/Reconstructor.cpp:1129-1831) -> real N^3 volume (origin at index 0)
// fsc == NULL: MAP off.  Returns the edge of the result.
int ref_reco_reconstruct(void* h, float* dst, int gridCorr, int joinHalf, const float* fsc, int nFsc, int nThread)
{
    RefReco* r = (RefReco*)h;
    r->reco.setGridCorr(gridCorr != 0);
    r->reco.setJoinHalf(joinHalf != 0);
    r->reco.setMAP(fsc != NULL);
    if (fsc)
    {
        vec f(nFsc);
        for (int i = 0; i < nFsc; i++) f(i) = fsc[i];
        r->reco.setFSC(f);
    }
    Volume v;
    r->reco.reconstruct(v, nThread);
    if (dst) memcpy(dst, &v(0), v.sizeRL() * sizeof(RFLOAT));
    return (int)v.nColRL();
}

int ref_reco_max_radius(void* h) { return ((RefReco*)h)->reco.maxRadius(); }

// ---------------------------------------------------------------- re-centre + re-mask of one image
// Optimiser::reCentreImg (src/Optimiser.cpp:6065-6091): _img = translate(_imgOri, offset) ; Optimiser::reMaskImg
// (:6093-6151, zeroMask): 2D c2r, x softMask(maskRadius / pixelSize, EDGE_WIDTH_RL), 2D r2c.
// imgOriFT / imgFT: half-complex [N][N/2+1] complex64 (FFTW layout).
void ref_recentre_remask(float* imgFT, const float* imgOriFT, int N, double offx, double offy, float maskRadiusPx, int zeroMask)
{
    Image ori(N, N, FT_SPACE), img(N, N, FT_SPACE);
    memcpy(&ori[0], imgOriFT, ori.sizeFT() * sizeof(Complex));
    translate(img, ori, offx, offy, 1);
    if (zeroMask)
    {
        Image mask(N, N, RL_SPACE);
        softMask(mask, maskRadiusPx, EDGE_WIDTH_RL, 1);
        FFT fft;
        fft.bw(img, 1);
        MUL_RL(img, mask);
        fft.fw(img, 1);
        img.clearRL();
    }
    memcpy(imgFT, &img[0], img.sizeFT() * sizeof(Complex));
}

// ---------------------------------------------------------------- sigma^2 refresh, per-image part
// Body of the image loop of Optimiser::allReduceSigma (src/Optimiser.cpp:6428-6600) with the reference's own functions,
// OPTIMISER_SIGMA_RANK1ST (one orientation per image: its rank-1st), MODE_3D, OPTIMISER_CTF_ON_THE_FLY, no CTF search,
// OPTIMISER_RECENTRE_IMAGE_EACH_ITERATION (the original image is compared at tran - offset), w = 1, per-group sums.
// sigM / sigN / svd: [nGroup][rSig + 1] float, last column = weight sum; the caller does the all-reduce and :6651-6709.
void ref_sigma_accumulate(void* projH, int nImg, int N, int rSig, const float* imgFT, const float* imgOriFT, const double* quat,
                          const double* tran, const double* offS, const float* ctfAttr7, float pixelSize, const int* group,
                          int nGroup, float* sigM, float* sigN, float* svd)
{
    RefProjector* P = (RefProjector*)projH;
    const size_t nFT = (size_t)(N / 2 + 1) * N;
    for (int i = 0; i < nGroup * (rSig + 1); i++) sigM[i] = sigN[i] = svd[i] = 0;
    for (int l = 0; l < nImg; l++)
    {
        Image img(N, N, FT_SPACE), imgOri(N, N, FT_SPACE), imgM(N, N, FT_SPACE), imgN(N, N, FT_SPACE), ctf(N, N, FT_SPACE);
        memcpy(&img[0], imgFT + 2 * nFT * l, nFT * sizeof(Complex));
        memcpy(&imgOri[0], imgOriFT + 2 * nFT * l, nFT * sizeof(Complex));
        SET_0_FT(imgM);
        SET_0_FT(imgN);
        SET_0_FT(ctf);
        dmat33 rot3D;
        rotate3D(rot3D, dvec4(quat[4 * l], quat[4 * l + 1], quat[4 * l + 2], quat[4 * l + 3]));
        dvec2 t(tran[2 * l], tran[2 * l + 1]), off(offS[2 * l], offS[2 * l + 1]);
        P->proj.project(imgM, rot3D, t, 1);
        P->proj.project(imgN, rot3D, t - off, 1);
        const float* a = ctfAttr7 + 7 * l;
        CTF(ctf, pixelSize, a[0], a[1], a[2], a[3], a[4], a[5], a[6], CEIL(rSig) + 1, 1);
        FOR_EACH_PIXEL_FT(imgM)
            imgM[i] *= REAL(ctf[i]);
        FOR_EACH_PIXEL_FT(imgN)
            imgN[i] *= REAL(ctf[i]);
        vec vSigM(rSig), vSigN(rSig), sSVD(rSig), dSVD(rSig);
        powerSpectrum(sSVD, imgM, rSig, 1);
        powerSpectrum(dSVD, img, rSig, 1);
        NEG_FT(imgM);
        NEG_FT(imgN);
        ADD_FT(imgM, img);
        ADD_FT(imgN, imgOri);
        powerSpectrum(vSigM, imgM, rSig, 1);
        powerSpectrum(vSigN, imgN, rSig, 1);
        const int g = group ? group[l] : 0;
        for (int i = 0; i < rSig; i++)
        {
            sigM[g * (rSig + 1) + i] += vSigM(i) / 2;
            sigN[g * (rSig + 1) + i] += vSigN(i) / 2;
            svd[g * (rSig + 1) + i] += sqrt(sSVD(i) / dSVD(i));
        }
        sigM[g * (rSig + 1) + rSig] += 1;
        sigN[g * (rSig + 1) + rSig] += 1;
        svd[g * (rSig + 1) + rSig] += 1;
    }
}


// Symmetry elements of a point group (include/Geometry/Symmetry.h): R matrices, column-major; returns their number
int ref_symmetry_elements(const char* name, double* R9, int maxElem)
{
    Symmetry sym(name);
    int n = sym.nSymmetryElement();
    for (int i = 0; i < n && i < maxElem; i++)
    {
        dmat33 L, R;
        sym.get(L, R, i);
        memcpy(R9 + 9 * i, R.data(), 9 * sizeof(double));
    }
    return n;
}

// Reconstructor::symmetrizeF / symmetrizeT / symmetrizeO (src/Reconstructor.cpp:2676-2716) on the accumulators as they are
void ref_reco_symmetrize(void* h, const char* name, int nThread)
{
    RefReco* r = (RefReco*)h;
    Symmetry sym(name);
    r->reco._sym = &sym;
    r->reco.symmetrizeT(nThread);
    r->reco.symmetrizeF(nThread);
    r->reco.symmetrizeO();
    r->reco._sym = NULL;
}

void ref_reco_set_O(void* h, const double* O3, int counter)
{
    RefReco* r = (RefReco*)h;
    r->reco._ox = O3[0]; r->reco._oy = O3[1]; r->reco._oz = O3[2];
    r->reco._counter = counter;
}

// MODE_2D twins of ref_reco_set / ref_reco_reconstruct and of Projector::setProjectee(Image) (src/Projector.cpp:97-121)
void ref_reco2d_set(void* h, const float* F, const float* T)
{
    RefReco* r = (RefReco*)h;
    size_t n = r->reco._F2D.sizeFT();
    memcpy(&r->reco._F2D[0], F, n * sizeof(Complex));
    for (size_t i = 0; i < n; i++) r->reco._T2D[i] = COMPLEX(T[i], 0);
}

int ref_reco2d_reconstruct(void* h, float* dst, int gridCorr, int joinHalf, const float* fsc, int nFsc, int nThread)
{
    RefReco* r = (RefReco*)h;
    r->reco.setGridCorr(gridCorr != 0);
    r->reco.setJoinHalf(joinHalf != 0);
    r->reco.setMAP(fsc != NULL);
    if (fsc)
    {
        vec f(nFsc);
        for (int i = 0; i < nFsc; i++) f(i) = fsc[i];
        r->reco.setFSC(f);
    }
    Volume v;
    r->reco.reconstruct(v, nThread);                    // N x N x 1 in MODE_2D
    if (dst) memcpy(dst, &v(0), v.sizeRL() * sizeof(RFLOAT));
    return (int)v.nColRL();
}

// real-space N x N image -> FFT -> Projector::setProjectee(Image); returns the padded dimension, out = padded FT
int ref_projector2d_set_from_real(void* h, const float* imgRL, int N, float* outFT)
{
    RefProjector* p = (RefProjector*)h;
    Image im(N, N, RL_SPACE);
    for (size_t i = 0; i < im.sizeRL(); i++) im(i) = imgRL[i];
    FFT fft;
    fft.fw(im, 1);
    im.clearRL();
    p->proj.setProjectee(im.copyImage(), 1);
    const int n = (int)p->proj._projectee2D.nRowFT();
    if (outFT) memcpy(outFT, &p->proj._projectee2D[0], p->proj._projectee2D.sizeFT() * sizeof(Complex));
    return n;
}

// image loop of Optimiser::normCorrection (src/Optimiser.cpp:6201-6350, MODE_3D, OPTIMISER_NORM_MASK, no CTF search) with the
// reference's functions: norm[l] = sum_{rL^2 <= |k|^2 < rNorm^2} |_img[l] - CTF * project(rot, tran)|^2 (RFLOAT accumulator)
void ref_norm_residual(void* projH, int nImg, int N, float rL, float rNorm, const float* imgFT, const double* quat,
                       const double* tran, const float* ctfAttr7, float pixelSize, float* norm)
{
    RefProjector* P = (RefProjector*)projH;
    const size_t nFT = (size_t)(N / 2 + 1) * N;
    for (int l = 0; l < nImg; l++)
    {
        Image img(N, N, FT_SPACE), mask(N, N, FT_SPACE), ctf(N, N, FT_SPACE);
        memcpy(&mask[0], imgFT + 2 * nFT * l, nFT * sizeof(Complex));
        SET_0_FT(img);
        SET_0_FT(ctf);
        dmat33 rot3D;
        rotate3D(rot3D, dvec4(quat[4 * l], quat[4 * l + 1], quat[4 * l + 2], quat[4 * l + 3]));
        P->proj.project(img, rot3D, dvec2(tran[2 * l], tran[2 * l + 1]), 1);
        const float* a = ctfAttr7 + 7 * l;
        CTF(ctf, pixelSize, a[0], a[1], a[2], a[3], a[4], a[5], a[6], CEIL(rNorm) + 1, 1);
        FOR_EACH_PIXEL_FT(img)
            img[i] *= REAL(ctf[i]);
        NEG_FT(img);
        ADD_FT(img, mask);
        RFLOAT s = 0;
        IMAGE_FOR_EACH_PIXEL_FT(img)
        {
            if ((QUAD(i, j) >= TSGSL_pow_2(rL)) && (QUAD(i, j) < TSGSL_pow_2(rNorm)))
                s += ABS2(img.getFTHalf(i, j));
        }
        norm[l] = s;
    }
}

// von Mises-like family of the MODE_2D particle filter (src/Geometry/DirectionalStat.cpp:252-384), as is
void ref_sample_vms(double k, int n, double* cs)
{
    dmat4 d(n, 4);                                   // the dmat4 overload is the one Particle::perturb calls (and the one that links)
    sampleVMS(d, dvec4(1, 0, 0, 0), k, n);
    for (int i = 0; i < n; i++) { cs[2 * i] = d(i, 0); cs[2 * i + 1] = d(i, 1); }
}

void ref_infer_vms(int n, const double* cs, double* mu2, double* k)
{
    dmat2 src(n, 2);
    for (int i = 0; i < n; i++) { src(i, 0) = cs[2 * i]; src(i, 1) = cs[2 * i + 1]; }
    dvec2 mu;
    inferVMS(mu, *k, src);
    mu2[0] = mu(0); mu2[1] = mu(1);
}

double ref_pdf_vms(const double* x2, const double* mu2, double k) { return pdfVMS(dvec2(x2[0], x2[1]), dvec2(mu2[0], mu2[1]), k); }

// Particle::balanceWeight(PAR_R) in MODE_2D (src/Particle.cpp:2315-2329) with the reference's own inferVMS / pdfVMS
void ref_balance_r_2d(int n, const double* cs, double* w)
{
    dmat2 src(n, 2);
    for (int i = 0; i < n; i++) { src(i, 0) = cs[2 * i]; src(i, 1) = cs[2 * i + 1]; }
    dvec2 mu;
    double k;
    inferVMS(mu, k, src);
    for (int i = 0; i < n; i++) w[i] = 1.0 / pdfVMS(dvec2(src(i, 0), src(i, 1)), mu, k);
}

// Particle::resample(nOut, PAR_C) of the reference class itself (MODE_2D particle with nIn classes)
int ref_particle_resample_c(int nIn, const int* c, const double* wC, const double* uC, int nOut, int* cOut, double* wOut)
{
    Particle p(MODE_2D, nIn, 1, 1, 1, 2.0, 0.01, NULL);
    for (int i = 0; i < nIn; i++) { p._c(i) = c[i]; p._wC(i) = wC[i]; p._uC(i) = uC[i]; }
    p.resample(nOut, PAR_C);
    for (int j = 0; j < nOut; j++) { cOut[j] = (int)p._c(j); wOut[j] = p._wC(j); }
    return (int)p._topC;
}

// ---------------------------------------------------------------- MODE_2D (2D classification, demo_2D.json)
// Projector in MODE_2D holding an already padded half-complex class average [pfN][pfN/2+1] verbatim
void* ref_projector2d_create(int pf, const float* imgFT, int pfN)
{
    RefProjector* p = new RefProjector();
    p->proj.setMode(MODE_2D);
    p->proj.setInterp(LINEAR_INTERP);
    p->proj.setPf(pf);
    p->proj._projectee2D.alloc(pfN, pfN, FT_SPACE);
    memcpy(&p->proj._projectee2D[0], imgFT, p->proj._projectee2D.sizeFT() * sizeof(Complex));
    p->proj._maxRadius = pfN / pf / 2 - 1;
    return p;
}

// rotate2D(dmat22&, dvec2) (src/Geometry/Euler.cpp:125-131) + Projector::project(Complex*, const dmat22&, iCol, iRow, nPxl,
// nThread) (src/Projector.cpp:337-354)
void ref_projector2d_project(void* h, float* dst, const double* cs, const int* iCol, const int* iRow, int nPxl)
{
    dmat22 rot;
    rotate2D(rot, dvec2(cs[0], cs[1]));
    ((RefProjector*)h)->proj.project((Complex*)dst, rot, iCol, iRow, nPxl, 1);
}

void* ref_reco2d_create(int size, int N, int pf, int nThread)
{
    RefReco* r = new RefReco();
    r->reco.setMPIEnv(3, 1, MPI_COMM_SELF, MPI_COMM_SELF);
    r->reco.init(MODE_2D, size, N, pf, NULL, 1.9, 15);
    r->reco.allocSpace(nThread);
    return r;
}

int ref_reco2d_pad_size(void* h) { return (int)((RefReco*)h)->reco._F2D.nRowFT(); }

// the M-step body of Optimiser::reconstructRef in MODE_2D (src/Optimiser.cpp:7072-7148) for one draw of one image:
// translate(-(t - offset)) of the unmasked packed image, Reconstructor::insertP(src, ctf, dmat22, w, NULL)
// (src/Reconstructor.cpp:708-780), insertDir(-rot2D * (t - offset))
void ref_reco2d_insert_draw(void* h, const float* datP, const float* ctfP, int N, const int* iCol, const int* iRow, int nPxl,
                            const double* cs, const double* tran, const double* off, float w)
{
    RefReco* r = (RefReco*)h;
    std::vector<Complex> tmp(nPxl);
    dvec2 t(tran[0] - (off ? off[0] : 0.0), tran[1] - (off ? off[1] : 0.0));
    translate(tmp.data(), (const Complex*)datP, -t(0), -t(1), N, N, iCol, iRow, nPxl, 1);
    dmat22 rot;
    rotate2D(rot, dvec2(cs[0], cs[1]));
    r->reco.insertP(tmp.data(), ctfP, rot, w, NULL);
    dvec2 dir = -rot * t;
    r->reco.insertDir(dir);
}

void ref_reco2d_get(void* h, float* F, float* T, double* O3, int* counter)
{
    RefReco* r = (RefReco*)h;
    size_t n = r->reco._F2D.sizeFT();
    if (F) memcpy(F, &r->reco._F2D[0], n * sizeof(Complex));
    if (T)
        for (size_t i = 0; i < n; i++) T[i] = REAL(r->reco._T2D[i]);
    if (O3)
    {
        O3[0] = r->reco._ox;
        O3[1] = r->reco._oy;
        O3[2] = r->reco._oz;
    }
    if (counter) *counter = r->reco._counter;
}

// ---------------------------------------------------------------- Particle (reference class, as is)
void* ref_particle_create(int nC, int nR, int nT, int nD, double transS, double transQ)
{
    return new Particle(MODE_3D, nC, nR, nT, nD, transS, transQ, NULL);
}
void ref_particle_destroy(void* h) { delete (Particle*)h; }

void ref_particle_load(void* h, int nR, int nT, int nD, const double* q, double k1, double k2, double k3,
                       const double* t, double s0, double s1, double d, double s, double score)
{
    ((Particle*)h)->load(nR, nT, nD, dvec4(q[0], q[1], q[2], q[3]), k1, k2, k3, dvec2(t[0], t[1]), s0, s1, d, s,
                         score);
}

void ref_particle_get_counts(void* h, int* n4)
{
    Particle* p = (Particle*)h;
    n4[0] = p->nC(); n4[1] = p->nR(); n4[2] = p->nT(); n4[3] = p->nD();
}

// state <-> flat arrays: r[nR][4], t[nT][2], d[nD], wC,wR,wT,wD, uC,uR,uT,uD ; c[nC] as int
void ref_particle_get(void* h, int* c, double* r, double* t, double* d, double* wC, double* wR, double* wT,
                      double* wD, double* uC, double* uR, double* uT, double* uD)
{
    Particle* p = (Particle*)h;
    for (int i = 0; i < p->_nC; i++) { if (c) c[i] = (int)p->_c(i); if (wC) wC[i] = p->_wC(i); if (uC) uC[i] = p->_uC(i); }
    for (int i = 0; i < p->_nR; i++)
    {
        if (r) for (int j = 0; j < 4; j++) r[i * 4 + j] = p->_r(i, j);
        if (wR) wR[i] = p->_wR(i);
        if (uR) uR[i] = p->_uR(i);
    }
    for (int i = 0; i < p->_nT; i++)
    {
        if (t) { t[i * 2] = p->_t(i, 0); t[i * 2 + 1] = p->_t(i, 1); }
        if (wT) wT[i] = p->_wT(i);
        if (uT) uT[i] = p->_uT(i);
    }
    for (int i = 0; i < p->_nD; i++) { if (d) d[i] = p->_d(i); if (wD) wD[i] = p->_wD(i); if (uD) uD[i] = p->_uD(i); }
}

void ref_particle_set(void* h, const int* c, const double* r, const double* t, const double* d, const double* wC,
                      const double* wR, const double* wT, const double* wD)
{
    Particle* p = (Particle*)h;
    for (int i = 0; i < p->_nC; i++) { if (c) p->_c(i) = c[i]; if (wC) p->_wC(i) = wC[i]; }
    for (int i = 0; i < p->_nR; i++)
    {
        if (r) for (int j = 0; j < 4; j++) p->_r(i, j) = r[i * 4 + j];
        if (wR) p->_wR(i) = wR[i];
    }
    for (int i = 0; i < p->_nT; i++)
    {
        if (t) { p->_t(i, 0) = t[i * 2]; p->_t(i, 1) = t[i * 2 + 1]; }
        if (wT) p->_wT(i) = wT[i];
    }
    for (int i = 0; i < p->_nD; i++) { if (d) p->_d(i) = d[i]; if (wD) p->_wD(i) = wD[i]; }
}

// scalars: k1,k2,k3,s0,s1,rho,s,score, topR[4], topT[2], topD, peakFactor C,R,T,D  (19 doubles)
void ref_particle_get_scalars(void* h, double* out)
{
    Particle* p = (Particle*)h;
    out[0] = p->_k1; out[1] = p->_k2; out[2] = p->_k3; out[3] = p->_s0; out[4] = p->_s1; out[5] = p->_rho;
    out[6] = p->_s; out[7] = p->_score;
    for (int j = 0; j < 4; j++) out[8 + j] = p->_topR(j);
    out[12] = p->_topT(0); out[13] = p->_topT(1); out[14] = p->_topD;
    out[15] = p->_peakFactorC; out[16] = p->_peakFactorR; out[17] = p->_peakFactorT; out[18] = p->_peakFactorD;
}
void ref_particle_set_scalars(void* h, const double* in)
{
    Particle* p = (Particle*)h;
    p->_k1 = in[0]; p->_k2 = in[1]; p->_k3 = in[2]; p->_s0 = in[3]; p->_s1 = in[4]; p->_rho = in[5];
    p->_s = in[6]; p->_score = in[7];
    for (int j = 0; j < 4; j++) p->_topR(j) = in[8 + j];
    p->_topT(0) = in[12]; p->_topT(1) = in[13]; p->_topD = in[14];
    p->_peakFactorC = in[15]; p->_peakFactorR = in[16]; p->_peakFactorT = in[17]; p->_peakFactorD = in[18];
}

void ref_particle_set_u(void* h, int which /*0 C,1 R,2 T,3 D*/, const double* u, int n)
{
    Particle* p = (Particle*)h;
    for (int i = 0; i < n; i++)
    {
        if (which == 0) p->setUC(u[i], i);
        else if (which == 1) p->setUR(u[i], i);
        else if (which == 2) p->setUT(u[i], i);
        else p->setUD(u[i], i);
    }
}

// The per-image logic that follows the global scan in Optimiser::expectation (reference src/Optimiser.cpp:921-1075) on a Particle
// that holds the shared scan grid with the uniform weights of Particle::reset: class choice, marginals of that class, peak
// factors, resampling down to (mLR, mLT), variances and their floors (OPTIMISER_SCAN_SET_MIN_STD_WITH_PERTURB: the caller passes
// the floors).  Replay mode: the engine is keyed (seed, stream, epoch) after the grid is in place.
int ref_particle_from_scan(int mode2D, int nK, int nR, int nT, const double* gridR, const double* gridT, const float* wC,
                           const float* wR, const float* wT, int mLR, int mLT, double kFloor, double sFloor, double transS,
                           double transQ, unsigned long long seed, unsigned long long stream, unsigned long long epoch, double* rOut,
                           double* tOut, double* wROut, double* wTOut, double* scal19)
{
    Particle par(mode2D ? MODE_2D : MODE_3D, nK, nR, nT, 1, transS, transQ, NULL);
    {
        uvec c(nK);
        for (int k = 0; k < nK; k++) c(k) = k;
        par.setC(c);
        dmat4 r = dmat4::Zero(nR, 4);
        for (int i = 0; i < nR; i++)
            for (int j = 0; j < (mode2D ? 2 : 4); j++) r(i, j) = gridR[(size_t)i * (mode2D ? 2 : 4) + j];
        par.setR(r);
        dmat2 t(nT, 2);
        for (int i = 0; i < nT; i++) { t(i, 0) = gridT[2 * i]; t(i, 1) = gridT[2 * i + 1]; }
        par.setT(t);
        par.setWC(dvec::Constant(nK, 1.0 / nK)); par.setWR(dvec::Constant(nR, 1.0 / nR)); par.setWT(dvec::Constant(nT, 1.0 / nT));
        par.setUC(dvec::Constant(nK, 1.0 / nK)); par.setUR(dvec::Constant(nR, 1.0 / nR)); par.setUT(dvec::Constant(nT, 1.0 / nT));
    }
    if (g_replay) replay_key(seed, stream, epoch);

    for (int iC = 0; iC < nK; iC++) par.setUC(wC[iC], iC);
    par.setPeakFactor(PAR_C);
    par.keepHalfHeightPeak(PAR_C);
    par.resample(nK, PAR_C);
    size_t cls;
    par.rand(cls);
    par.setNC(1);
    par.setC(uvec::Constant(1, cls));
    par.setWC(dvec::Constant(1, 1));
    par.setUC(dvec::Constant(1, 1));
    for (int iR = 0; iR < nR; iR++) par.setUR(wR[(size_t)cls * nR + iR], iR);
    for (int iT = 0; iT < nT; iT++) par.setUT(wT[(size_t)cls * nT + iT], iT);
    par.setPeakFactor(PAR_R);
    par.keepHalfHeightPeak(PAR_R);
    par.resample(mLR, PAR_R);
    par.resample(mLT, PAR_T);
    par.calVari(PAR_R);
    par.calVari(PAR_T);
    par.setK1(TSGSL_MAX_RFLOAT(kFloor, par.k1()));
    if (!mode2D)
    {
        par.setK2(TSGSL_MAX_RFLOAT(kFloor, par.k2()));
        par.setK3(TSGSL_MAX_RFLOAT(kFloor, par.k3()));
    }
    par.setS0(TSGSL_MAX_RFLOAT(sFloor, par.s0()));
    par.setS1(TSGSL_MAX_RFLOAT(sFloor, par.s1()));

    for (int i = 0; i < mLR; i++) { for (int j = 0; j < 4; j++) rOut[i * 4 + j] = par._r(i, j); wROut[i] = par._wR(i); }
    for (int i = 0; i < mLT; i++) { tOut[2 * i] = par._t(i, 0); tOut[2 * i + 1] = par._t(i, 1); wTOut[i] = par._wT(i); }
    ref_particle_get_scalars(&par, scal19);
    return (int)cls;
}

void ref_particle_initD(void* h, int nD, double sD) { ((Particle*)h)->initD(nD, sD); }
void ref_particle_perturb(void* h, double pf, int pt) { ((Particle*)h)->perturb(pf, (ParticleType)pt); }
void ref_particle_resample(void* h, int n, int pt) { ((Particle*)h)->resample(n, (ParticleType)pt); }
void ref_particle_calVari(void* h, int pt) { ((Particle*)h)->calVari((ParticleType)pt); }
void ref_particle_calRank1st(void* h, int pt) { ((Particle*)h)->calRank1st((ParticleType)pt); }
void ref_particle_keepHalfHeightPeak(void* h, int pt) { ((Particle*)h)->keepHalfHeightPeak((ParticleType)pt); }
void ref_particle_setPeakFactor(void* h, int pt) { ((Particle*)h)->setPeakFactor((ParticleType)pt); }
void ref_particle_resetPeakFactor(void* h) { ((Particle*)h)->resetPeakFactor(); }
void ref_particle_normW(void* h) { ((Particle*)h)->normW(); }
void ref_particle_shuffle(void* h, int pt) { ((Particle*)h)->shuffle((ParticleType)pt); }
void ref_particle_balanceWeight(void* h, int pt) { ((Particle*)h)->balanceWeight((ParticleType)pt); }
void ref_particle_calScore(void* h) { ((Particle*)h)->calScore(); }
double ref_particle_compressR(void* h) { return ((Particle*)h)->compressR(); }
double ref_particle_compressT(void* h) { return ((Particle*)h)->compressT(); }
double ref_particle_variR(void* h) { return ((Particle*)h)->variR(); }
double ref_particle_variT(void* h) { return ((Particle*)h)->variT(); }
double ref_particle_variD(void* h) { return ((Particle*)h)->variD(); }
void ref_particle_rand(void* h, int* cls, double* quat, double* tran, double* d)
{
    size_t c; dvec4 q; dvec2 t; double dd;
    ((Particle*)h)->rand(c, q, t, dd);
    *cls = (int)c; for (int j = 0; j < 4; j++) quat[j] = q(j);
    tran[0] = t(0); tran[1] = t(1); *d = dd;
}
void ref_particle_rank1st(void* h, int* cls, double* quat, double* tran, double* d)
{
    size_t c; dvec4 q; dvec2 t; double dd;
    ((Particle*)h)->rank1st(c, q, t, dd);
    *cls = (int)c; for (int j = 0; j < 4; j++) quat[j] = q(j);
    tran[0] = t(0); tran[1] = t(1); *d = dd;
}

// ---------------------------------------------------------------- DirectionalStat (free functions)
// inferACG(dmat44&, const dmat4&) (src/Geometry/DirectionalStat.cpp:93-145) + the derived
// k1,k2,k3 (:184-222) and mean (:224-250); pdfACG (:19-24)
void ref_inferACG(const double* r, int n, double* A16_rowmajor, double* k123, double* mean4)
{
    dmat4 src(n, 4);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 4; j++) src(i, j) = r[i * 4 + j];
    if (A16_rowmajor)
    {
        dmat44 A;
        inferACG(A, src);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) A16_rowmajor[i * 4 + j] = A(i, j);
    }
    if (k123) inferACG(k123[0], k123[1], k123[2], src);
    if (mean4)
    {
        dvec4 m;
        inferACG(m, src);
        for (int j = 0; j < 4; j++) mean4[j] = m(j);
    }
}

double ref_pdfACG(const double* x4, const double* A16_rowmajor)
{
    dmat44 A;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) A(i, j) = A16_rowmajor[i * 4 + j];
    return pdfACG(dvec4(x4[0], x4[1], x4[2], x4[3]), A);
}

// ------------------------------------------------------------------------------------------
// Driver loop 1: the particle-filter phase loop of Optimiser::expectation
// (reference src/Optimiser.cpp:1162-1660) for SEARCH_TYPE_LOCAL, MODE_3D, k = 1, no CTF search,
// default Config.h switches (OPTIMISER_PEAK_FACTOR_R/T on, OPTIMISER_COMPRESS_CRITERIA on).
//
//   pars        nImg Particle handles (state is advanced in place)
//   datP,ctfP,sigRcpP   image-major packed arrays [nImg][nPxl]
//   minPhase/maxPhase   MIN_N_PHASE_PER_ITER_LOCAL / MAX_N_PHASE_PER_ITER; fixedPhases > 0 runs
//                       exactly that many phases (benchmark mode, stop rule disabled)
//   nPhaseOut   phases actually run per image
//   dvpOut      optional [nImg][mLR*mLT] log-likelihoods of the LAST phase run (for parity)
// ------------------------------------------------------------------------------------------
void ref_expectation_local_trace(void** pars, int nImg, void* projH, const float* datP, const float* ctfP,
                           const float* sigRcpP, const int* iCol, const int* iRow, int nPxl, int N, int mLR,
                           int mLT, double perturbFactorL, double perturbFactorS, int minPhase, int maxPhase,
                           int noDecreaseLimit, double decreaseFactor, int fixedPhases, int simd,
                           int nThread, int* nPhaseOut, float* dvpOut, const float* uRIn, const float* uTIn, float* uROwn,
                           float* uTOwn, int nTrace, double* condOut, const double* rIn, const double* tIn, double* rPert,
                           double* tPert, double* rRes, double* tRes);

void ref_expectation_local(void** pars, int nImg, void* projH, const float* datP, const float* ctfP,
                           const float* sigRcpP, const int* iCol, const int* iRow, int nPxl, int N, int mLR,
                           int mLT, double perturbFactorL, double perturbFactorS, int minPhase, int maxPhase,
                           int noDecreaseLimit, double decreaseFactor, int fixedPhases, int simd,
                           int nThread, int* nPhaseOut, float* dvpOut)
{
    ref_expectation_local_trace(pars, nImg, projH, datP, ctfP, sigRcpP, iCol, iRow, nPxl, N, mLR, mLT, perturbFactorL,
                                perturbFactorS, minPhase, maxPhase, noDecreaseLimit, decreaseFactor, fixedPhases, simd, nThread,
                                nPhaseOut, dvpOut, NULL, NULL, NULL, NULL, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
}

// The same loop with a trace.  Replay mode (ref_rng_replay(1), ref_rng_replay_loop(seed, stream, epoch)): the engine of image
// l is keyed (seed, stream + l, epoch) before the first perturbation and (seed, stream + l, epoch + phase + 1) before the
// post-likelihood operators of every phase - the points at which the CUDA library keys its own generator.
//   uROwn / uTOwn   out, [nTrace][nImg][mLR / mLT]: the marginal weights THIS loop computed in each phase (relative to its
//                   final baseline = the maximum), from the reference's project / translate / logDataVSPrior
//   uRIn / uTIn     in, same shape, optional: weights handed to setUR / setUT INSTEAD of the loop's own (the device's, so that
//                   both filters resample from bit-identical weights and their states can be compared exactly; the loop's
//                   own weights are still computed, from ITS state, and returned for the comparison with the device's)
void ref_expectation_local_trace(void** pars, int nImg, void* projH, const float* datP, const float* ctfP,
                           const float* sigRcpP, const int* iCol, const int* iRow, int nPxl, int N, int mLR,
                           int mLT, double perturbFactorL, double perturbFactorS, int minPhase, int maxPhase,
                           int noDecreaseLimit, double decreaseFactor, int fixedPhases, int simd,
                           int nThread, int* nPhaseOut, float* dvpOut, const float* uRIn, const float* uTIn, float* uROwn,
                           float* uTOwn, int nTrace, double* condOut, const double* rIn, const double* tIn, double* rPert,
                           double* tPert, double* rRes, double* tRes)
{
    RefProjector* P = (RefProjector*)projH;
    if (nThread <= 0) nThread = omp_get_max_threads();

    Complex* poolPriRotP = (Complex*)TSFFTW_malloc((size_t)nPxl * nThread * sizeof(Complex));
    Complex* poolPriAllP = (Complex*)TSFFTW_malloc((size_t)nPxl * nThread * sizeof(Complex));
    Complex* poolTraP = (Complex*)TSFFTW_malloc((size_t)mLT * nPxl * nThread * sizeof(Complex));

    #pragma omp parallel for schedule(dynamic) num_threads(nThread)
    for (int l = 0; l < nImg; l++)
    {
        Particle& par = *(Particle*)pars[l];
        Complex* priRotP = poolPriRotP + (size_t)nPxl * omp_get_thread_num();
        Complex* priAllP = poolPriAllP + (size_t)nPxl * omp_get_thread_num();
        Complex* traP = poolTraP + (size_t)mLT * nPxl * omp_get_thread_num();
        Complex* dat = (Complex*)datP + (size_t)l * nPxl;
        const RFLOAT* ctf = ctfP + (size_t)l * nPxl;
        const RFLOAT* sigRcp = sigRcpP + (size_t)l * nPxl;

        int nPhaseWithNoVariDecrease = 0;
        double variR = DBL_MAX, variT = DBL_MAX, variD = DBL_MAX;
        int phasesRun = 0;
        int phaseMax = fixedPhases > 0 ? fixedPhases : maxPhase;

        for (int phase = 0; phase < phaseMax; phase++)
        {
            if (phase == 0)
            {
                if (g_replay) replay_key(g_rp_seed, g_rp_stream + g_rp_stride * (unsigned long long)l, g_rp_epoch);
                par.perturb(perturbFactorL, PAR_R);
                par.perturb(perturbFactorL, PAR_T);
            }
            else
            {
                par.perturb(perturbFactorS, PAR_R);
                par.perturb(perturbFactorS, PAR_T);
            }

            if (phase < nTrace)
            {
                if (rPert) for (int i = 0; i < mLR; i++) for (int j = 0; j < 4; j++) rPert[(((size_t)phase * nImg + l) * mLR + i) * 4 + j] = par._r(i, j);
                if (tPert) for (int i = 0; i < mLT; i++) for (int j = 0; j < 2; j++) tPert[(((size_t)phase * nImg + l) * mLT + i) * 2 + j] = par._t(i, j);
                if (rIn)
                {
                    for (int i = 0; i < mLR; i++) for (int j = 0; j < 4; j++) par._r(i, j) = rIn[(((size_t)phase * nImg + l) * mLR + i) * 4 + j];
                    par.balanceWeight(PAR_R);
                }
                if (tIn)
                {
                    for (int i = 0; i < mLT; i++) for (int j = 0; j < 2; j++) par._t(i, j) = tIn[(((size_t)phase * nImg + l) * mLT + i) * 2 + j];
                    par.balanceWeight(PAR_T);
                }
            }

            RFLOAT baseLine = GSL_NAN;
            vec wC = vec::Zero(1);
            vec wR = vec::Zero(mLR);
            vec wT = vec::Zero(mLT);
            vec wD = vec::Zero(1);

            dmat33 rot3D;
            dvec2 t;

            FOR_EACH_C(par)
            {
                FOR_EACH_T(par)
                {
                    par.t(t, iT);
                    translate(traP + (size_t)iT * nPxl, t(0), t(1), N, N, iCol, iRow, nPxl, 1);
                }

                FOR_EACH_R(par)
                {
                    par.rot(rot3D, iR);
                    P->proj.project(priRotP, rot3D, iCol, iRow, nPxl, 1);

                    FOR_EACH_T(par)
                    {
                        for (int i = 0; i < nPxl; i++) priAllP[i] = traP[(size_t)nPxl * iT + i] * priRotP[i];

                        FOR_EACH_D(par)
                        {
                            RFLOAT w = simd ? logDataVSPrior_m_huabin_SIMD256(dat, priAllP, ctf, sigRcp, nPxl)
                                            : logDataVSPrior_m_huabin(dat, priAllP, ctf, sigRcp, nPxl);

                            if (dvpOut && (fixedPhases <= 0 || phase == phaseMax - 1))
                                dvpOut[(size_t)l * mLR * mLT + (size_t)iR * mLT + iT] = w;

                            baseLine = TSGSL_isnan(baseLine) ? w : baseLine;
                            if (w > baseLine)
                            {
                                RFLOAT nf = exp(baseLine - w);
                                wC *= nf; wR *= nf; wT *= nf; wD *= nf;
                                baseLine = w;
                            }
                            RFLOAT s = exp(w - baseLine);
                            wC(iC) += s * (par.wR(iR) * par.wT(iT) * par.wD(iD));
                            wR(iR) += s * (par.wC(iC) * par.wT(iT) * par.wD(iD));
                            wT(iT) += s * (par.wC(iC) * par.wR(iR) * par.wD(iD));
                            wD(iD) += s * (par.wC(iC) * par.wR(iR) * par.wT(iT));
                        }
                    }
                }
            }

            if (phase < nTrace)
            {
                if (uROwn) for (int iR = 0; iR < mLR; iR++) uROwn[((size_t)phase * nImg + l) * mLR + iR] = wR(iR);
                if (uTOwn) for (int iT = 0; iT < mLT; iT++) uTOwn[((size_t)phase * nImg + l) * mLT + iT] = wT(iT);
                if (uRIn) for (int iR = 0; iR < mLR; iR++) wR(iR) = uRIn[((size_t)phase * nImg + l) * mLR + iR];
                if (uTIn) for (int iT = 0; iT < mLT; iT++) wT(iT) = uTIn[((size_t)phase * nImg + l) * mLT + iT];
            }
            if (g_replay) replay_key(g_rp_seed, g_rp_stream + g_rp_stride * (unsigned long long)l, g_rp_epoch + (unsigned long long)phase + 1);

            par.setUC(wC(0), 0);
            for (int iR = 0; iR < mLR; iR++) par.setUR(wR(iR), iR);
            par.keepHalfHeightPeak(PAR_R);
            for (int iT = 0; iT < mLT; iT++) par.setUT(wT(iT), iT);
            // OPTIMISER_PEAK_FACTOR_T is off in the reference's Config.h:218 -> no keepHalfHeightPeak(PAR_T)

            par.calRank1st(PAR_R);
            par.calRank1st(PAR_T);
            par.calVari(PAR_R);
            par.calVari(PAR_T);
            par.resample(mLR, PAR_R);
            par.resample(mLT, PAR_T);

            if (phase < nTrace)
            {
                if (rRes) for (int i = 0; i < mLR; i++) for (int j = 0; j < 4; j++) rRes[(((size_t)phase * nImg + l) * mLR + i) * 4 + j] = par._r(i, j);
                if (tRes) for (int i = 0; i < mLT; i++) for (int j = 0; j < 2; j++) tRes[(((size_t)phase * nImg + l) * mLT + i) * 2 + j] = par._t(i, j);
            }
            if (condOut && phase < nTrace)
            {
                dmat44 A;
                inferACG(A, par._r);
                Eigen::SelfAdjointEigenSolver<dmat44> es(A);
                const double lo = es.eigenvalues()(0), hi = es.eigenvalues()(3);
                condOut[(size_t)phase * nImg + l] = (lo > 0) ? hi / lo : 1e300;
            }

            phasesRun = phase + 1;

            if (fixedPhases <= 0 && phase >= minPhase)
            {
                double variRCur = par.variR();
                double variTCur = par.variT();
                double variDCur = par.variD();

                if ((variRCur < variR * decreaseFactor) || (variTCur < variT * decreaseFactor) ||
                    (variDCur < variD * decreaseFactor))
                    nPhaseWithNoVariDecrease = 0;
                else
                    nPhaseWithNoVariDecrease += 1;

                if (variRCur < variR) variR = variRCur;
                if (variTCur < variT) variT = variTCur;
                if (variDCur < variD) variD = variDCur;

                if (nPhaseWithNoVariDecrease == noDecreaseLimit) break;
            }
        }
        if (nPhaseOut) nPhaseOut[l] = phasesRun;
    }

    TSFFTW_free(poolPriRotP);
    TSFFTW_free(poolPriAllP);
    TSFFTW_free(poolTraP);
}

// ------------------------------------------------------------------------------------------
// Driver loop 2: the insert loop of Optimiser::reconstructRef (reference
// src/Optimiser.cpp:7036-7241), MODE_3D, k = 1, no CTF search, OPTIMISER_RECENTRE_IMAGE_EACH_ITERATION.
// Draws come either from Particle::rand (pars != NULL) or from explicit arrays
// nr[nImg][mReco][4], nt[nImg][mReco][2] (the layout the reference GPU seam uses, :6993-7013).
// ------------------------------------------------------------------------------------------
void ref_insert_loop(void* recoH, void** pars, int nImg, const float* datP, const float* ctfP,
                     const float* wImg, const double* offS, const double* nr, const double* nt,
                     const int* iCol, const int* iRow, int nPxl, int N, int mReco, int nThread)
{
    RefReco* R = (RefReco*)recoH;
    if (nThread <= 0) nThread = omp_get_max_threads();
    Complex* poolTransImgP = (Complex*)TSFFTW_malloc((size_t)nPxl * nThread * sizeof(Complex));

    #pragma omp parallel for num_threads(nThread)
    for (int l = 0; l < nImg; l++)
    {
        RFLOAT w = wImg ? wImg[l] : (RFLOAT)1;
        if (!wImg) w /= mReco;   // reference: w = 1 (or compressR) then w /= mReco; wImg is already divided

        Complex* transImgP = poolTransImgP + (size_t)nPxl * omp_get_thread_num();
        const Complex* orignImgP = (const Complex*)datP + (size_t)nPxl * l;
        dvec2 offset(offS ? offS[2 * l] : 0, offS ? offS[2 * l + 1] : 0);

        if (g_replay && pars) replay_key(g_rp_seed, g_rp_stream + g_rp_stride * (unsigned long long)l, g_rp_epoch);

        for (int m = 0; m < mReco; m++)
        {
            size_t cls;
            dvec4 quat;
            dvec2 tran;
            double d;

            if (pars)
                ((Particle*)pars[l])->rand(cls, quat, tran, d);
            else
            {
                const double* q = nr + ((size_t)l * mReco + m) * 4;
                const double* t = nt + ((size_t)l * mReco + m) * 2;
                quat = dvec4(q[0], q[1], q[2], q[3]);
                tran = dvec2(t[0], t[1]);
            }

            dmat33 rot3D;
            rotate3D(rot3D, quat);

            translate(transImgP, orignImgP, -(tran - offset)(0), -(tran - offset)(1), N, N, iCol, iRow, nPxl, 1);

            R->reco.insertP(transImgP, ctfP + (size_t)nPxl * l, rot3D, w, NULL);

            dvec3 dir = -rot3D * dvec3((tran - offset)[0], (tran - offset)[1], 0);
            R->reco.insertDir(dir);
        }
    }
    TSFFTW_free(poolTransImgP);
}


// ------------------------------------------------------------------------------------------
// CTF search (SEARCH_TYPE_CTF).  (1) what Optimiser::allocPreCal prepares for it (src/Optimiser.cpp:8125-8168), one image;
// (2) one phase of the likelihood loop with the defocus dimension (:1236-1402) for explicit supports and prior weights;
// (3) the insert loop of reconstructRef with cSearch (:7067-7241): the CTF of every draw from its own defocus factor.
// ------------------------------------------------------------------------------------------
void ref_precal_ctf(float voltage, float defocusU, float defocusV, float defocusTheta, float Cs, int N, float pixelSize,
                    const int* iCol, const int* iRow, int nPxl, float* frequency, float* defocusP, float* K1K2)
{
    for (int i = 0; i < nPxl; i++)
        frequency[i] = NORM(iCol[i], iRow[i]) / N / pixelSize;
    for (int i = 0; i < nPxl; i++)
    {
        RFLOAT angle = atan2(iRow[i], iCol[i]) - defocusTheta;
        RFLOAT defocus = -(defocusU + defocusV + (defocusU - defocusV) * cos(2 * angle)) / 2;
        defocusP[i] = defocus;
    }
    RFLOAT lambda = 12.2643274 / sqrt(voltage * (1 + voltage * 0.978466e-6));
    K1K2[0] = M_PI * lambda;
    K1K2[1] = M_PI_2 * Cs * TSGSL_pow_3(lambda);
}

void ref_expect_ctf(void* projH, const float* datP, const float* sigRcpP, const float* defocusP, const float* frequency, float K1,
                    float K2, float phaseShift, float amplitudeContrast, const double* quat, const double* tran, const double* dpar,
                    const double* pwR, const double* pwT, const double* pwD, double pwC, const int* iCol, const int* iRow, int nPxl,
                    int N, int nR, int nT, int nD, int simd, float* wCout, float* wRout, float* wTout, float* wDout, float* baseOut,
                    float* logL)
{
    RefProjector* P = (RefProjector*)projH;
    Complex* priRotP = (Complex*)TSFFTW_malloc((size_t)nPxl * sizeof(Complex));
    Complex* priAllP = (Complex*)TSFFTW_malloc((size_t)nPxl * sizeof(Complex));
    Complex* traP = (Complex*)TSFFTW_malloc((size_t)nT * nPxl * sizeof(Complex));
    RFLOAT* ctfP = (RFLOAT*)TSFFTW_malloc((size_t)nD * nPxl * sizeof(RFLOAT));
    Complex* dat = (Complex*)datP;

    for (int iT = 0; iT < nT; iT++)
        translate(traP + (size_t)iT * nPxl, tran[2 * iT], tran[2 * iT + 1], N, N, iCol, iRow, nPxl, 1);

    for (int iD = 0; iD < nD; iD++)
    {
        double d = dpar[iD];
        for (int i = 0; i < nPxl; i++)
        {
            RFLOAT ki = K1 * defocusP[i] * d * TSGSL_pow_2(frequency[i]) + K2 * TSGSL_pow_4(frequency[i]) - phaseShift;
            ctfP[(size_t)nPxl * iD + i] = -TS_SQRT(1 - TSGSL_pow_2(amplitudeContrast)) * TS_SIN(ki) + amplitudeContrast * TS_COS(ki);
        }
    }

    RFLOAT baseLine = GSL_NAN;
    vec wC = vec::Zero(1), wR = vec::Zero(nR), wT = vec::Zero(nT), wD = vec::Zero(nD);
    for (int iR = 0; iR < nR; iR++)
    {
        dmat33 rot3D;
        rotate3D(rot3D, dvec4(quat[4 * iR], quat[4 * iR + 1], quat[4 * iR + 2], quat[4 * iR + 3]));
        P->proj.project(priRotP, rot3D, iCol, iRow, nPxl, 1);
        for (int iT = 0; iT < nT; iT++)
        {
            for (int i = 0; i < nPxl; i++) priAllP[i] = traP[(size_t)nPxl * iT + i] * priRotP[i];
            for (int iD = 0; iD < nD; iD++)
            {
                RFLOAT w = simd ? logDataVSPrior_m_huabin_SIMD256(dat, priAllP, ctfP + (size_t)iD * nPxl, sigRcpP, nPxl)
                                : logDataVSPrior_m_huabin(dat, priAllP, ctfP + (size_t)iD * nPxl, sigRcpP, nPxl);
                if (logL) logL[((size_t)iR * nT + iT) * nD + iD] = w;
                baseLine = TSGSL_isnan(baseLine) ? w : baseLine;
                if (w > baseLine)
                {
                    RFLOAT nf = exp(baseLine - w);
                    wC *= nf; wR *= nf; wT *= nf; wD *= nf;
                    baseLine = w;
                }
                RFLOAT s = exp(w - baseLine);
                wC(0) += s * (pwR[iR] * pwT[iT] * pwD[iD]);
                wR(iR) += s * (pwC * pwT[iT] * pwD[iD]);
                wT(iT) += s * (pwC * pwR[iR] * pwD[iD]);
                wD(iD) += s * (pwC * pwR[iR] * pwT[iT]);
            }
        }
    }
    wCout[0] = wC(0);
    for (int i = 0; i < nR; i++) wRout[i] = wR(i);
    for (int i = 0; i < nT; i++) wTout[i] = wT(i);
    for (int i = 0; i < nD; i++) wDout[i] = wD(i);
    *baseOut = baseLine;
    TSFFTW_free(priRotP); TSFFTW_free(priAllP); TSFFTW_free(traP); TSFFTW_free(ctfP);
}

// ------------------------------------------------------------------------------------------
// Driver loop 3: the initial phase of the global search / 2D classification (reference src/Optimiser.cpp:756-914): one shared set
// of nR rotations x nT translations against ALL images, class by class - projection of one template per (class, rotation),
// product with every translation's phase ramp, logDataVSPrior_m_n over the pixel-major packed arrays, running-baseline weights
// under per-image locks; OpenMP over rotations, as the reference.
//   projH[nK]            Projector handles (MODE_2D: ref_projector2d_create, rot[nR][2] = (cos, sin); MODE_3D: rot[nR][4])
//   datPM, ctfPM, sigPM  pixel-major [nPxl][nImg]
//   out: wC[nImg][nK], wR[nK][nImg][nR], wT[nK][nImg][nT], baseLine[nImg]
// ------------------------------------------------------------------------------------------
void ref_scan(void** projH, int nK, int mode2D, const float* datPM, const float* ctfPM, const float* sigPM, int nImg, int nPxl, int N,
              const int* iCol, const int* iRow, const double* rot, int nR, const double* tran, int nT, const double* pR,
              const double* pT, int simd, int nThread, float* wCout, float* wRout, float* wTout, float* baseOut)
{
    if (nThread <= 0) nThread = omp_get_max_threads();
    Complex* traP = (Complex*)TSFFTW_malloc((size_t)nT * nPxl * sizeof(Complex));
    #pragma omp parallel for schedule(dynamic) num_threads(nThread)
    for (int m = 0; m < nT; m++) translate(traP + (size_t)m * nPxl, tran[2 * m], tran[2 * m + 1], N, N, iCol, iRow, nPxl, 1);
    Complex* poolPriRotP = (Complex*)TSFFTW_malloc((size_t)nPxl * nThread * sizeof(Complex));
    Complex* poolPriAllP = (Complex*)TSFFTW_malloc((size_t)nPxl * nThread * sizeof(Complex));
    RFLOAT* poolSIMDResult = (RFLOAT*)TSFFTW_malloc((size_t)nImg * nThread * sizeof(RFLOAT));
    std::vector<omp_lock_t> mtx(nImg);
    for (int l = 0; l < nImg; l++) omp_init_lock(&mtx[l]);
    std::vector<RFLOAT> baseLine(nImg, GSL_NAN);
    memset(wCout, 0, sizeof(float) * (size_t)nImg * nK);
    memset(wRout, 0, sizeof(float) * (size_t)nK * nImg * nR);
    memset(wTout, 0, sizeof(float) * (size_t)nK * nImg * nT);

    for (int t = 0; t < nK; t++)
    {
        RefProjector* P = (RefProjector*)projH[t];
        #pragma omp parallel for schedule(dynamic) num_threads(nThread)
        for (int m = 0; m < nR; m++)
        {
            Complex* priRotP = poolPriRotP + (size_t)nPxl * omp_get_thread_num();
            Complex* priAllP = poolPriAllP + (size_t)nPxl * omp_get_thread_num();
            RFLOAT* SIMDResult = poolSIMDResult + (size_t)omp_get_thread_num() * nImg;
            if (mode2D)
            {
                dmat22 rot2D;
                rotate2D(rot2D, dvec2(rot[2 * m], rot[2 * m + 1]));
                P->proj.project(priRotP, rot2D, iCol, iRow, nPxl, 1);
            }
            else
            {
                dmat33 rot3D;
                rotate3D(rot3D, dvec4(rot[4 * m], rot[4 * m + 1], rot[4 * m + 2], rot[4 * m + 3]));
                P->proj.project(priRotP, rot3D, iCol, iRow, nPxl, 1);
            }
            for (int n = 0; n < nT; n++)
            {
                for (int i = 0; i < nPxl; i++) priAllP[i] = traP[(size_t)nPxl * n + i] * priRotP[i];
                memset(SIMDResult, '\0', nImg * sizeof(RFLOAT));
                RFLOAT* dvp = simd ? logDataVSPrior_m_n_huabin_SIMD256((Complex*)datPM, priAllP, (RFLOAT*)ctfPM, (RFLOAT*)sigPM, nImg, nPxl, SIMDResult)
                                   : logDataVSPrior_m_n_huabin((Complex*)datPM, priAllP, (RFLOAT*)ctfPM, (RFLOAT*)sigPM, nImg, nPxl, SIMDResult);
                for (int l = 0; l < nImg; l++)
                {
                    omp_set_lock(&mtx[l]);
                    if (TSGSL_isnan(baseLine[l]))
                        baseLine[l] = dvp[l];
                    else if (dvp[l] > baseLine[l])
                    {
                        RFLOAT offset = dvp[l] - baseLine[l];
                        RFLOAT nf = exp(-offset);
                        for (int td = 0; td < nK; td++)
                        {
                            wCout[(size_t)l * nK + td] *= nf;
                            for (int a = 0; a < nR; a++) wRout[((size_t)td * nImg + l) * nR + a] *= nf;
                            for (int a = 0; a < nT; a++) wTout[((size_t)td * nImg + l) * nT + a] *= nf;
                        }
                        baseLine[l] += offset;
                    }
                    RFLOAT w = exp(dvp[l] - baseLine[l]);
                    wCout[(size_t)l * nK + t] += w * (pR[m] * pT[n]);
                    wRout[((size_t)t * nImg + l) * nR + m] += w * pT[n];
                    wTout[((size_t)t * nImg + l) * nT + n] += w * pR[m];
                    omp_unset_lock(&mtx[l]);
                }
            }
        }
    }
    for (int l = 0; l < nImg; l++) { baseOut[l] = baseLine[l]; omp_destroy_lock(&mtx[l]); }
    TSFFTW_free(traP); TSFFTW_free(poolPriRotP); TSFFTW_free(poolPriAllP); TSFFTW_free(poolSIMDResult);
}

// the insert loop of reconstructRef in MODE_2D with several classes (src/Optimiser.cpp:7036-7148): mReco draws per image, each into
// the Reconstructor of its class; OpenMP over images as the reference
void ref_insert_loop_2d(void** recoH, int nImg, const float* datP, const float* ctfP, const float* wImg, const double* offS,
                        const int* nc, const double* nr, const double* nt, const int* iCol, const int* iRow, int nPxl, int N,
                        int mReco, int nThread)
{
    if (nThread <= 0) nThread = omp_get_max_threads();
    Complex* pool = (Complex*)TSFFTW_malloc((size_t)nPxl * nThread * sizeof(Complex));
    #pragma omp parallel for num_threads(nThread)
    for (int l = 0; l < nImg; l++)
    {
        Complex* transImgP = pool + (size_t)nPxl * omp_get_thread_num();
        const Complex* orignImgP = (const Complex*)datP + (size_t)nPxl * l;
        dvec2 offset(offS ? offS[2 * l] : 0, offS ? offS[2 * l + 1] : 0);
        for (int m = 0; m < mReco; m++)
        {
            const size_t o = (size_t)l * mReco + m;
            dvec2 tran(nt[2 * o], nt[2 * o + 1]);
            dmat22 rot2D;
            rotate2D(rot2D, dvec2(nr[2 * o], nr[2 * o + 1]));
            translate(transImgP, orignImgP, -(tran - offset)(0), -(tran - offset)(1), N, N, iCol, iRow, nPxl, 1);
            RefReco* R = (RefReco*)recoH[nc[o]];
            R->reco.insertP(transImgP, ctfP + (size_t)nPxl * l, rot2D, wImg[l], NULL);
            dvec2 dir = -rot2D * (tran - offset);
            R->reco.insertDir(dir);
        }
    }
    TSFFTW_free(pool);
}

void ref_insert_loop_ctf(void* recoH, int nImg, const float* datP, const float* wImg, const double* offS, const double* nr,
                         const double* nt, const double* nd, const float* ctfAttr, float pixelSize, const int* iCol,
                         const int* iRow, int nPxl, int N, int mReco, int nThread)
{
    RefReco* R = (RefReco*)recoH;
    if (nThread <= 0) nThread = omp_get_max_threads();
    Complex* poolTransImgP = (Complex*)TSFFTW_malloc((size_t)nPxl * nThread * sizeof(Complex));
    RFLOAT* poolCtf = (RFLOAT*)TSFFTW_malloc((size_t)nPxl * nThread * sizeof(RFLOAT));

    #pragma omp parallel for num_threads(nThread)
    for (int l = 0; l < nImg; l++)
    {
        RFLOAT w = wImg[l];
        Complex* transImgP = poolTransImgP + (size_t)nPxl * omp_get_thread_num();
        RFLOAT* ctf = poolCtf + (size_t)nPxl * omp_get_thread_num();
        const Complex* orignImgP = (const Complex*)datP + (size_t)nPxl * l;
        dvec2 offset(offS ? offS[2 * l] : 0, offS ? offS[2 * l + 1] : 0);
        const float* a = ctfAttr + 7 * (size_t)l;     // voltage, defocusU, defocusV, defocusTheta, Cs, amplitudeContrast, phaseShift
        for (int m = 0; m < mReco; m++)
        {
            const double* q = nr + ((size_t)l * mReco + m) * 4;
            const double* t = nt + ((size_t)l * mReco + m) * 2;
            double d = nd[(size_t)l * mReco + m];
            dvec4 quat(q[0], q[1], q[2], q[3]);
            dvec2 tran(t[0], t[1]);
            dmat33 rot3D;
            rotate3D(rot3D, quat);
            translate(transImgP, orignImgP, -(tran - offset)(0), -(tran - offset)(1), N, N, iCol, iRow, nPxl, 1);
            CTF(ctf, pixelSize, a[0], a[1] * d, a[2] * d, a[3], a[4], a[5], a[6], N, N, iCol, iRow, nPxl, 1);
            R->reco.insertP(transImgP, ctf, rot3D, w, NULL);
            dvec3 dir = -rot3D * dvec3((tran - offset)[0], (tran - offset)[1], 0);
            R->reco.insertDir(dir);
        }
    }
    TSFFTW_free(poolTransImgP);
    TSFFTW_free(poolCtf);
}

}  // extern "C"
