// Interface.cpp - the reference's accelerator seam (gpu/interface/Interface.{h,cpp}) re-implemented on the
// C ABI of libthunder_b200.so.  Host glue only: every number is computed by the CUDA kernels behind thb_*.
#include "Interface.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace {

std::mutex g_mu;
std::map<int, thb_ctx*> g_ctx;
bool g_fillIntermediates = false;   // THB_FILL_INTERMEDIATES=1: ExpectProject also returns rotP on the host

// the global-scan calls of the reference are split over three functions; what the first two are given is kept here
struct ScanState {
    std::vector<double> trans, rot;     // [nT][2], [nR][4]
    int nR = 0, nT = 0, idim = 0, npxl = 0;
    std::vector<int> iCol, iRow;
    const Complex* volume = nullptr;    // identity of the uploaded projector volume (pointer + vdim)
    int vdim = 0, pf = 0;
    const Complex* stackDat = nullptr;  // identity of the uploaded image stack
    int stackImgs = 0;
    std::vector<float> aC, aR, aT, aBase;   // running (over classes) per-image results
} g_scan;

[[noreturn]] void die(thb_ctx* ctx, const char* where, int rc)
{
    // the reference's seam prints and exits on any device error (gpu/config/Device.cuh.in:27-61)
    fprintf(stderr, "thunder_b200 [%s]: error %d: %s\n", where, rc, thb_last_error(ctx));
    abort();
}

#define CHK(ctx, call)                          \
    do {                                        \
        int rc__ = (call);                      \
        if (rc__ != THB_OK) die(ctx, #call, rc__); \
    } while (0)

}  // namespace

thb_ctx* thbContext(int gpuIdx)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ctx.find(gpuIdx);
    if (it != g_ctx.end()) return it->second;
    thb_ctx* c = nullptr;
    int rc = thb_create(&c, gpuIdx);
    if (rc != THB_OK) die(nullptr, "thb_create", rc);
    if (const char* e = getenv("THB_FILL_INTERMEDIATES")) g_fillIntermediates = atoi(e) != 0;
    g_ctx[gpuIdx] = c;
    return c;
}

void thbShutdown()
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_ctx) thb_destroy(kv.second);
    g_ctx.clear();
    g_scan = ScanState();
}

void getAviDevice(std::vector<int>& gpus)
{
    gpus.clear();
    const int n = thb_device_count();
    for (int i = 0; i < n; ++i) gpus.push_back(i);
}

void ExpectPreidx(int gpuIdx, int** deviCol, int** deviRow, int* iCol, int* iRow, int npxl)
{
    (void)thbContext(gpuIdx);
    g_scan.iCol.assign(iCol, iCol + npxl);
    g_scan.iRow.assign(iRow, iRow + npxl);
    g_scan.npxl = npxl;
    if (deviCol) *deviCol = iCol;   // opaque tokens: the pixel list lives inside the library
    if (deviRow) *deviRow = iRow;
}

void ExpectFreeIdx(int, int** deviCol, int** deviRow)
{
    if (deviCol) *deviCol = nullptr;
    if (deviRow) *deviRow = nullptr;
}

void ExpectRotran(Complex* traP, double* trans, double* rot, double* rotMat, const int* iCol, const int* iRow, int nR, int nT,
                  int idim, int npxl)
{
    (void)traP;   // the phase ramps are computed inside the fused kernel; nothing on the hot path reads this array
    g_scan.trans.assign(trans, trans + (size_t)nT * 2);
    g_scan.rot.assign(rot, rot + (size_t)nR * 4);
    g_scan.nR = nR; g_scan.nT = nT; g_scan.idim = idim; g_scan.npxl = npxl;
    g_scan.iCol.assign(iCol, iCol + npxl);
    g_scan.iRow.assign(iRow, iRow + npxl);
    if (rotMat) {
        // R = I + 2 w K + 2 K^2, column-major, as kernel_getRotMat (gpu/src/Kernel.cu:572-620) / Euler.cpp:181-189
        for (int r = 0; r < nR; ++r) {
            const double w = rot[4 * r], x = rot[4 * r + 1], y = rot[4 * r + 2], z = rot[4 * r + 3];
            double* m = rotMat + 9 * (size_t)r;
            m[0] = 1.0 + 2.0 * (-(y * y + z * z)); m[1] = 2.0 * w * z + 2.0 * (x * y);      m[2] = 2.0 * w * (-y) + 2.0 * (x * z);
            m[3] = 2.0 * w * (-z) + 2.0 * (x * y); m[4] = 1.0 + 2.0 * (-(x * x + z * z));    m[5] = 2.0 * w * x + 2.0 * (y * z);
            m[6] = 2.0 * w * y + 2.0 * (x * z);    m[7] = 2.0 * w * (-x) + 2.0 * (y * z);    m[8] = 1.0 + 2.0 * (-(x * x + y * y));
        }
    }
}

void ExpectProject(Complex* volume, Complex* rotP, double* rotMat, const int* iCol, const int* iRow, int nR, int pf, int interp,
                   int vdim, int npxl)
{
    (void)rotMat; (void)interp; (void)iCol; (void)iRow; (void)nR; (void)npxl;
    thb_ctx* ctx = thbContext(0);
    CHK(ctx, thb_set_mode(ctx, THB_MODE_3D));
    CHK(ctx, thb_set_expect_pixels(ctx, g_scan.idim, pf, g_scan.npxl, g_scan.iCol.data(), g_scan.iRow.data()));
    CHK(ctx, thb_set_volume(ctx, 0, reinterpret_cast<const float*>(volume), vdim));
    g_scan.volume = volume; g_scan.vdim = vdim; g_scan.pf = pf;
    g_scan.stackDat = nullptr;   // a new pixel list invalidates the resident stack
    if (g_fillIntermediates && rotP)
        CHK(ctx, thb_project(ctx, 0, g_scan.nR, g_scan.rot.data(), reinterpret_cast<float*>(rotP)));
}

// classes share one baseline per image: weights of earlier classes are rescaled when a later class raises it
// (kernel_setBaseLine, gpu/src/Kernel.cu:1096-1135; CPU: Optimiser.cpp:846-870)
static void mergeClass(const std::vector<float>& cC, const std::vector<float>& cR, const std::vector<float>& cT, const std::vector<float>& cB,
                       RFLOAT* wC, RFLOAT* wR, RFLOAT* wT, RFLOAT* baseL, int kIdx, int nK, int nR, int nT, int imgNum)
{
    for (int l = 0; l < imgNum; ++l) {
        float scaleOld = 1.f, scaleNew = 1.f;
        if (kIdx == 0) {
            baseL[l] = cB[l];
        } else if (cB[l] > baseL[l]) {
            scaleOld = expf(baseL[l] - cB[l]);
            baseL[l] = cB[l];
        } else {
            scaleNew = expf(cB[l] - baseL[l]);
        }
        if (scaleOld != 1.f)
            for (int k = 0; k < kIdx; ++k) {
                wC[(size_t)l * nK + k] *= scaleOld;
                for (int r = 0; r < nR; ++r) wR[((size_t)l * nK + k) * nR + r] *= scaleOld;
                for (int t = 0; t < nT; ++t) wT[((size_t)l * nK + k) * nT + t] *= scaleOld;
            }
        wC[(size_t)l * nK + kIdx] = cC[l] * scaleNew;
        for (int r = 0; r < nR; ++r) wR[((size_t)l * nK + kIdx) * nR + r] = cR[(size_t)l * nR + r] * scaleNew;
        for (int t = 0; t < nT; ++t) wT[((size_t)l * nK + kIdx) * nT + t] = cT[(size_t)l * nT + t] * scaleNew;
    }
}

void ExpectGlobal3D(Complex* rotP, Complex* traP, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP, RFLOAT* wC, RFLOAT* wR, RFLOAT* wT,
                    double* pR, double* pT, RFLOAT* baseL, int kIdx, int nK, int nR, int nT, int npxl, int imgNum)
{
    (void)rotP; (void)traP; (void)npxl;
    thb_ctx* ctx = thbContext(0);
    if (g_scan.stackDat != datP || g_scan.stackImgs != imgNum) {
        CHK(ctx, thb_upload_stack(ctx, THB_STACK_EXPECT, imgNum, reinterpret_cast<const float*>(datP), ctfP, sigRcpP, nullptr));
        g_scan.stackDat = datP; g_scan.stackImgs = imgNum;
    }
    std::vector<float> cC(imgNum), cR((size_t)imgNum * nR), cT((size_t)imgNum * nT), cB(imgNum);
    CHK(ctx, thb_expect_scan(ctx, 0, nR, nT, g_scan.rot.data(), g_scan.trans.data(), pR, pT, cC.data(), cR.data(), cT.data(),
                             cB.data(), nullptr));
    mergeClass(cC, cR, cT, cB, wC, wR, wT, baseL, kIdx, nK, nR, nT, imgNum);
}

// MODE_2D classification scan (gpu/interface/Interface.h:176-198; call site src/Optimiser.cpp:1873-1920): vol = nK padded
// class averages [vdim][vdim/2+1], rot = nR x (cos, sin), every image against every class, one baseline per image
void ExpectGlobal2D(Complex* vol, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP, double* trans, RFLOAT* wC, RFLOAT* wR, RFLOAT* wT,
                    double* pR, double* pT, double* rot, const int* iCol, const int* iRow, int nK, int nR, int nT, int pf, int interp,
                    int idim, int vdim, int npxl, int imgNum)
{
    (void)interp;
    thb_ctx* ctx = thbContext(0);
    CHK(ctx, thb_set_mode(ctx, THB_MODE_2D));
    CHK(ctx, thb_set_expect_pixels(ctx, idim, pf, npxl, iCol, iRow));
    const size_t sizeModel = (size_t)(vdim / 2 + 1) * vdim;
    for (int k = 0; k < nK; ++k) CHK(ctx, thb_set_volume(ctx, k, reinterpret_cast<const float*>(vol + k * sizeModel), vdim));
    CHK(ctx, thb_upload_stack(ctx, THB_STACK_EXPECT, imgNum, reinterpret_cast<const float*>(datP), ctfP, sigRcpP, nullptr));
    g_scan.stackDat = nullptr;
    std::vector<float> cC(imgNum), cR((size_t)imgNum * nR), cT((size_t)imgNum * nT), cB(imgNum), baseL(imgNum);
    for (int k = 0; k < nK; ++k) {
        CHK(ctx, thb_expect_scan(ctx, k, nR, nT, rot, trans, pR, pT, cC.data(), cR.data(), cT.data(), cB.data(), nullptr));
        mergeClass(cC, cR, cT, cB, wC, wR, wT, baseL.data(), k, nK, nR, nT, imgNum);
    }
}

// MODE_2D M-step (gpu/interface/Interface.h:239-265; call site src/Optimiser.cpp:6770-6850): F2D [nk][vdim][vdim/2+1] complex,
// T2D the same as reals, O2D [nk][2], counter [nk]; nC / nR / nT: class, (cos, sin) and translation of every draw
void InsertI2D(Complex* F2D, RFLOAT* T2D, double* O2D, int* counter, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP, RFLOAT* w, double* offS,
               int* nC, double* nR, double* nT, double* nD, void* ctfaData, const int* iCol, const int* iRow, RFLOAT pixelSize,
               bool cSearch, int nk, int opf, int npxl, int mReco, int idim, int vdim, int imgNum)
{
    (void)sigRcpP; (void)nD; (void)ctfaData; (void)pixelSize;
    thb_ctx* ctx = thbContext(0);
    if (cSearch) {
        fprintf(stderr, "thunder_b200 [InsertI2D]: CTF search (cSearch) is not on the accelerated path of this build\n");
        abort();
    }
    CHK(ctx, thb_set_mode(ctx, THB_MODE_2D));
    CHK(ctx, thb_set_insert_pixels(ctx, idim, opf, npxl, iCol, iRow));
    CHK(ctx, thb_upload_stack(ctx, THB_STACK_INSERT, imgNum, reinterpret_cast<const float*>(datP), ctfP, nullptr, nullptr));
    for (int k = 0; k < nk; ++k) CHK(ctx, thb_reco_alloc(ctx, k, vdim));
    CHK(ctx, thb_insert_classes(ctx, imgNum, nullptr, mReco, w, offS, nC, nR, nT));
    CHK(ctx, thb_allreduce(ctx));
    const size_t n = (size_t)(vdim / 2 + 1) * vdim;
    std::vector<float> F(2 * n), T(n);
    for (int k = 0; k < nk; ++k) {
        double O[3];
        int cnt = 0;
        CHK(ctx, thb_reco_download(ctx, k, F.data(), T.data(), O, &cnt, 0));
        for (size_t i = 0; i < n; ++i) {
            F2D[k * n + i].dat[0] += F[2 * i];
            F2D[k * n + i].dat[1] += F[2 * i + 1];
            T2D[k * n + i] += T[i];
        }
        O2D[2 * k] += O[0];
        O2D[2 * k + 1] += O[1];
        counter[k] += cnt;
    }
}

void ExpectLocalBatch(int gpuIdx, Complex* volume, int vdim, int pf, int idim, const int* iCol, const int* iRow, int npxl, Complex* datP,
                      RFLOAT* ctfP, RFLOAT* sigRcpP, int imgNum, int nR, int nT, const double* quat, const double* tran,
                      const double* wRprior, const double* wTprior, RFLOAT* wC, RFLOAT* wR, RFLOAT* wT, RFLOAT* baseL)
{
    thb_ctx* ctx = thbContext(gpuIdx);
    if (volume) {
        CHK(ctx, thb_set_mode(ctx, THB_MODE_3D));
        CHK(ctx, thb_set_expect_pixels(ctx, idim, pf, npxl, iCol, iRow));
        CHK(ctx, thb_set_volume(ctx, 0, reinterpret_cast<const float*>(volume), vdim));
    }
    if (datP) CHK(ctx, thb_upload_stack(ctx, THB_STACK_EXPECT, imgNum, reinterpret_cast<const float*>(datP), ctfP, sigRcpP, nullptr));
    CHK(ctx, thb_expect_local(ctx, imgNum, nullptr, nR, nT, quat, tran, wRprior, wTprior, wR, wT, wC, baseL, nullptr));
}

void InsertFT(Complex* F3D, Complex* T3D, int vdim, double* O3D, int* counter, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP,
              void* ctfaData, double* offS, RFLOAT* w, double* nR, double* nT, double* nD, int* nC, const int* iCol, const int* iRow,
              RFLOAT pixelSize, bool cSearch, int opf, int npxl, int mReco, int idim, int dimSize, int imgNum)
{
    (void)sigRcpP; (void)ctfaData; (void)nD; (void)pixelSize;
    thb_ctx* ctx = thbContext(0);
    if (cSearch) {
        fprintf(stderr, "thunder_b200 [InsertFT]: CTF search (cSearch) is not on the accelerated path of this build\n");
        abort();
    }
    const size_t nVox = (size_t)(vdim / 2 + 1) * vdim * vdim;
    if ((size_t)dimSize != nVox) {
        fprintf(stderr, "thunder_b200 [InsertFT]: dimSize %d does not match vdim %d\n", dimSize, vdim);
        abort();
    }
    CHK(ctx, thb_set_mode(ctx, THB_MODE_3D));
    CHK(ctx, thb_set_insert_pixels(ctx, idim, opf, npxl, iCol, iRow));   // Reconstructor's _iCol/_iRow are padded (x pf)
    CHK(ctx, thb_upload_stack(ctx, THB_STACK_INSERT, imgNum, reinterpret_cast<const float*>(datP), ctfP, nullptr, nullptr));
    CHK(ctx, thb_reco_alloc(ctx, 0, vdim));
    if (nC)   // 3D classification: nC[l] of the mReco rows of image l belong to this class (Optimiser.cpp:6862-6950)
        CHK(ctx, thb_insert_counts(ctx, imgNum, nullptr, mReco, w, offS, nC, nR, nT));
    else
        CHK(ctx, thb_insert(ctx, imgNum, nullptr, mReco, w, offS, nR, nT));
    CHK(ctx, thb_allreduce(ctx));                                         // hemisphere sum (no-op with one rank)
    std::vector<float> F(2 * nVox), T(nVox);
    double O[3];
    int cnt = 0;
    CHK(ctx, thb_reco_download(ctx, 0, F.data(), T.data(), O, &cnt, 0));
    for (size_t i = 0; i < nVox; ++i) {
        F3D[i].dat[0] += F[2 * i];
        F3D[i].dat[1] += F[2 * i + 1];
        T3D[i].dat[0] += T[i];
    }
    for (int k = 0; k < 3; ++k) O3D[k] += O[k];
    counter[0] += cnt;
}

// C-linkage aliases so that the parity tests can drive the shims through ctypes
extern "C" {
int thbi_device_count()
{
    std::vector<int> g;
    getAviDevice(g);
    return (int)g.size();
}
void thbi_ExpectRotran(float* traP, double* trans, double* rot, double* rotMat, const int* iCol, const int* iRow, int nR, int nT,
                       int idim, int npxl)
{
    ExpectRotran(reinterpret_cast<Complex*>(traP), trans, rot, rotMat, iCol, iRow, nR, nT, idim, npxl);
}
void thbi_ExpectProject(float* volume, float* rotP, double* rotMat, const int* iCol, const int* iRow, int nR, int pf, int interp,
                        int vdim, int npxl)
{
    ExpectProject(reinterpret_cast<Complex*>(volume), reinterpret_cast<Complex*>(rotP), rotMat, iCol, iRow, nR, pf, interp, vdim, npxl);
}
void thbi_ExpectGlobal3D(float* rotP, float* traP, float* datP, float* ctfP, float* sigRcpP, float* wC, float* wR, float* wT, double* pR,
                         double* pT, float* baseL, int kIdx, int nK, int nR, int nT, int npxl, int imgNum)
{
    ExpectGlobal3D(reinterpret_cast<Complex*>(rotP), reinterpret_cast<Complex*>(traP), reinterpret_cast<Complex*>(datP), ctfP, sigRcpP, wC,
                   wR, wT, pR, pT, baseL, kIdx, nK, nR, nT, npxl, imgNum);
}
void thbi_ExpectLocalBatch(int gpuIdx, float* volume, int vdim, int pf, int idim, const int* iCol, const int* iRow, int npxl, float* datP,
                           float* ctfP, float* sigRcpP, int imgNum, int nR, int nT, const double* quat, const double* tran,
                           const double* wRprior, const double* wTprior, float* wC, float* wR, float* wT, float* baseL)
{
    ExpectLocalBatch(gpuIdx, reinterpret_cast<Complex*>(volume), vdim, pf, idim, iCol, iRow, npxl, reinterpret_cast<Complex*>(datP), ctfP,
                     sigRcpP, imgNum, nR, nT, quat, tran, wRprior, wTprior, wC, wR, wT, baseL);
}
void thbi_InsertFT(float* F3D, float* T3D, int vdim, double* O3D, int* counter, float* datP, float* ctfP, double* offS, float* w, double* nR,
                   double* nT, const int* iCol, const int* iRow, int opf, int npxl, int mReco, int idim, int dimSize, int imgNum)
{
    InsertFT(reinterpret_cast<Complex*>(F3D), reinterpret_cast<Complex*>(T3D), vdim, O3D, counter, reinterpret_cast<Complex*>(datP), ctfP,
             nullptr, nullptr, offS, w, nR, nT, nullptr, nullptr, iCol, iRow, 1.32f, false, opf, npxl, mReco, idim, dimSize, imgNum);
}
void thbi_ExpectGlobal2D(float* vol, float* datP, float* ctfP, float* sigRcpP, double* trans, float* wC, float* wR, float* wT, double* pR,
                         double* pT, double* rot, const int* iCol, const int* iRow, int nK, int nR, int nT, int pf, int idim, int vdim,
                         int npxl, int imgNum)
{
    ExpectGlobal2D(reinterpret_cast<Complex*>(vol), reinterpret_cast<Complex*>(datP), ctfP, sigRcpP, trans, wC, wR, wT, pR, pT, rot, iCol,
                   iRow, nK, nR, nT, pf, 1, idim, vdim, npxl, imgNum);
}
void thbi_InsertI2D(float* F2D, float* T2D, double* O2D, int* counter, float* datP, float* ctfP, float* w, double* offS, int* nC, double* nR,
                    double* nT, const int* iCol, const int* iRow, int nk, int opf, int npxl, int mReco, int idim, int vdim, int imgNum)
{
    InsertI2D(reinterpret_cast<Complex*>(F2D), T2D, O2D, counter, reinterpret_cast<Complex*>(datP), ctfP, nullptr, w, offS, nC, nR, nT,
              nullptr, nullptr, iCol, iRow, 1.32f, false, nk, opf, npxl, mReco, idim, vdim, imgNum);
}
void thbi_shutdown() { thbShutdown(); }
}
