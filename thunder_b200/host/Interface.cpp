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

// ------------------------------------------------------------------------------------------------ local search, reference protocol
namespace {
struct LocalState {
    int nPxl = 0, slots = 0, pf = 0, idim = 0;
    bool pixelsSet = false;
    bool cSearch = false;
    std::vector<float> freq;            // ExpectPrefre: _frequency of the pixel list
    bool freqSet = false;
} g_local[64];
LocalState& localOf(int gpuIdx) { return g_local[gpuIdx & 63]; }
int g_token;    // address handed out as the opaque device pointer
}  // namespace

void ExpectPrefre(int gpuIdx, RFLOAT** devfreQ, RFLOAT* freQ, int npxl)
{
    // the frequency table feeds the on-the-fly CTF of the CTF search only (src/Optimiser.cpp:2173-2185); it goes to the device
    // together with the pixel list (local_pixels)
    LocalState& L = localOf(gpuIdx);
    L.freq.assign(freQ, freQ + npxl);
    L.freqSet = false;
    if (devfreQ) *devfreQ = (RFLOAT*)&g_token;
}

void ExpectLocalIn(int gpuIdx, Complex** devdatP, RFLOAT** devctfP, RFLOAT** devdefO, RFLOAT** devsigP, int nPxl, int cpyNumL,
                   int searchType)
{
    thb_ctx* c = thbContext(gpuIdx);
    LocalState& L = localOf(gpuIdx);
    L.cSearch = searchType == 2;           // SEARCH_TYPE_CTF (include/Optimiser.h: SEARCH_TYPE_GLOBAL 0, _LOCAL 1, _CTF 2)
    if (L.cSearch && (int)L.freq.size() != nPxl) { fprintf(stderr, "thunder_b200 [ExpectLocalIn]: CTF search: ExpectPrefre must come first\n"); abort(); }
    if (g_scan.npxl != nPxl || g_scan.iCol.empty()) { fprintf(stderr, "thunder_b200 [ExpectLocalIn]: ExpectPreidx must come first (pixel list of %d pixels)\n", nPxl); abort(); }
    L.nPxl = nPxl; L.slots = cpyNumL > 0 ? cpyNumL : 1; L.pixelsSet = false;
    if (devdatP) *devdatP = (Complex*)&g_token;
    if (devctfP) *devctfP = (RFLOAT*)&g_token;
    if (devdefO) *devdefO = (RFLOAT*)&g_token;
    if (devsigP) *devsigP = (RFLOAT*)&g_token;
    (void)c;
}

void ExpectLocalV3D(int gpuIdx, ManagedArrayTexture* mgr, Complex* volume, int vdim)
{
    thb_ctx* c = thbContext(gpuIdx);
    if (thb_get_mode(c) != THB_MODE_3D) CHK(c, thb_set_mode(c, THB_MODE_3D));
    mgr->slot = 0;
    CHK(c, thb_set_volume(c, 0, (const float*)volume, vdim));
}

void ExpectLocalV2D(int gpuIdx, ManagedArrayTexture* mgr, Complex* volume, int dimSize)
{
    // the caller creates one manager per class in class order (src/Optimiser.cpp:2216-2230): slot = creation order per device
    thb_ctx* c = thbContext(gpuIdx);
    if (thb_get_mode(c) != THB_MODE_2D) CHK(c, thb_set_mode(c, THB_MODE_2D));
    (void)dimSize;
    CHK(c, thb_set_volume(c, mgr->slot, (const float*)volume, mgr->getVdim()));
}

static void local_pixels(thb_ctx* c, LocalState& L, int pf, int idim)
{
    if (L.pixelsSet && L.pf == pf && L.idim == idim) return;
    CHK(c, thb_set_expect_pixels(c, idim, pf, L.nPxl, g_scan.iCol.data(), g_scan.iRow.data()));
    CHK(c, thb_stack_reserve(c, THB_STACK_EXPECT, L.slots));
    if (L.cSearch) CHK(c, thb_set_frequency(c, L.freq.data()));
    L.pf = pf; L.idim = idim; L.pixelsSet = true;
}

// image slots need the pixel list, which needs pf and the image size: the reference passes them only to ExpectLocalPreI*, so the
// images given to ExpectLocalP before the first ExpectLocalPreI* call of a device are kept as host pointers and uploaded there
namespace {
struct PendingImg { const Complex* dat; const RFLOAT* ctf; const RFLOAT* sig; const RFLOAT* def; };
std::map<long long, PendingImg> g_pending;    // key = gpuIdx * 4096 + slot
}

static void upload_slot(thb_ctx* c, int slot, int npxl, const PendingImg& im)
{
    (void)npxl;
    CHK(c, thb_upload_stack_at(c, THB_STACK_EXPECT, slot, 1, (const float*)im.dat, im.ctf, im.sig, nullptr));
    if (im.def) CHK(c, thb_upload_stack_defocus(c, slot, 1, im.def));
}

void ExpectLocalP(int gpuIdx, Complex*, RFLOAT*, RFLOAT*, RFLOAT*, Complex* datP, RFLOAT* ctfP, RFLOAT* defO, RFLOAT* sigP,
                  int threadId, int imgId, int npxl, int cSearch)
{
    thb_ctx* c = thbContext(gpuIdx);
    LocalState& L = localOf(gpuIdx);
    // CTF search: the per-pixel defocus of the image instead of its CTF (the reference passes ctfP all the same; it may be the
    // stale array of the previous search type, and is not read)
    const bool cs = cSearch == 2;
    PendingImg im{datP + (size_t)imgId * npxl, ctfP ? ctfP + (size_t)imgId * npxl : nullptr, sigP + (size_t)imgId * npxl,
                  cs && defO ? defO + (size_t)imgId * npxl : nullptr};
    std::vector<float> zeros;
    if (!im.ctf) { zeros.assign(npxl, 0.f); im.ctf = zeros.data(); }
    std::lock_guard<std::mutex> lk(g_mu);
    if (L.pixelsSet) upload_slot(c, threadId, npxl, im);
    else g_pending[(long long)gpuIdx * 4096 + threadId] = im;
}

void ExpectLocalHostA(int, RFLOAT** wC, RFLOAT** wR, RFLOAT** wT, RFLOAT** wD, double** oldR, double** oldT, double** oldD,
                      double** trans, double** rot, double** dpara, int mR, int mT, int mD, int cSearch)
{
    (void)cSearch;
    const int nD = mD > 0 ? mD : 1;
    *wC = (RFLOAT*)calloc(1, sizeof(RFLOAT)); *wR = (RFLOAT*)calloc(mR, sizeof(RFLOAT)); *wT = (RFLOAT*)calloc(mT, sizeof(RFLOAT));
    *wD = (RFLOAT*)calloc(nD, sizeof(RFLOAT));
    *oldR = (double*)calloc(mR, sizeof(double)); *oldT = (double*)calloc(mT, sizeof(double)); *oldD = (double*)calloc(nD, sizeof(double));
    *trans = (double*)calloc((size_t)mT * 2, sizeof(double)); *rot = (double*)calloc((size_t)mR * 4, sizeof(double));
    *dpara = (double*)calloc(nD, sizeof(double));
}

void ExpectLocalHostF(int, RFLOAT** wC, RFLOAT** wR, RFLOAT** wT, RFLOAT** wD, double** oldR, double** oldT, double** oldD,
                      double** trans, double** rot, double** dpara, int)
{
    free(*wC); free(*wR); free(*wT); free(*wD); free(*oldR); free(*oldT); free(*oldD); free(*trans); free(*rot); free(*dpara);
    *wC = *wR = *wT = *wD = nullptr; *oldR = *oldT = *oldD = *trans = *rot = *dpara = nullptr;
}

void ExpectLocalRTD(int, ManagedCalPoint* mcp, double* oldR, double* oldT, double* oldD, double* trans, double* rot, double* dpara)
{
    mcp->oldR = oldR; mcp->oldT = oldT; mcp->oldD = oldD; mcp->trans = trans; mcp->rot = rot; mcp->dpara = dpara;
}

static void local_prei(int gpuIdx, ManagedArrayTexture* mgr, ManagedCalPoint* mcp, int pf, int idim, int npxl, RFLOAT phaseShift = 0,
                       RFLOAT conT = 0, RFLOAT k1 = 0, RFLOAT k2 = 0)
{
    mcp->ctfK[0] = k1; mcp->ctfK[1] = k2; mcp->ctfK[2] = phaseShift; mcp->ctfK[3] = conT;
    thb_ctx* c = thbContext(gpuIdx);
    LocalState& L = localOf(gpuIdx);
    std::lock_guard<std::mutex> lk(g_mu);
    if (!L.pixelsSet || L.pf != pf || L.idim != idim) {
        local_pixels(c, L, pf, idim);
        for (auto it = g_pending.begin(); it != g_pending.end();) {
            if (it->first / 4096 == gpuIdx) { upload_slot(c, (int)(it->first % 4096), npxl, it->second); it = g_pending.erase(it); }
            else ++it;
        }
    }
    mcp->mgr = mgr;
}

void ExpectLocalPreI3D(int gpuIdx, int, ManagedArrayTexture* mgr, ManagedCalPoint* mcp, RFLOAT*, RFLOAT*, int*, int*, RFLOAT phaseShift,
                       RFLOAT conT, RFLOAT k1, RFLOAT k2, int pf, int idim, int, int npxl, int)
{
    local_prei(gpuIdx, mgr, mcp, pf, idim, npxl, phaseShift, conT, k1, k2);
}

void ExpectLocalPreI2D(int gpuIdx, int, ManagedArrayTexture* mgr, ManagedCalPoint* mcp, RFLOAT*, RFLOAT*, int*, int*, RFLOAT, RFLOAT,
                       RFLOAT, RFLOAT, int pf, int idim, int, int npxl, int)
{
    local_prei(gpuIdx, mgr, mcp, pf, idim, npxl);
}

void ExpectLocalM(int gpuIdx, int datShift, ManagedCalPoint* mcp, Complex*, RFLOAT*, RFLOAT*, RFLOAT* wC, RFLOAT* wR, RFLOAT* wT,
                  RFLOAT* wD, double oldC, int npxl)
{
    thb_ctx* c = thbContext(gpuIdx);
    (void)npxl;
    const int nR = mcp->getNR(), nT = mcp->getNT();
    if (!mcp->rot || !mcp->trans || !mcp->oldR || !mcp->oldT) { fprintf(stderr, "thunder_b200 [ExpectLocalM]: ExpectLocalRTD must come first\n"); abort(); }
    const int img = datShift;
    // MODE_2D: rot[mR][4] holds (cos, sin, 0, 0) per row (Particle::quaternion of a 2D particle); the C ABI takes [mR][2]
    std::vector<double> cs;
    const double* q = mcp->rot;
    if (thb_get_mode(c) == THB_MODE_2D) {
        cs.resize((size_t)nR * 2);
        for (int r = 0; r < nR; ++r) { cs[2 * r] = mcp->rot[4 * r]; cs[2 * r + 1] = mcp->rot[4 * r + 1]; }
        q = cs.data();
    }
    float uC = 0.f;
    if (mcp->getCSearch() == 2) {
        // CTF search: the defocus dimension (mD factors dpara with prior weights oldD), CTF on the fly from k1 / k2 / phaseShift / conT
        const int nD = mcp->getMD();
        if (!mcp->dpara || !mcp->oldD || !wD) { fprintf(stderr, "thunder_b200 [ExpectLocalM]: CTF search needs dpara / oldD / wD\n"); abort(); }
        std::lock_guard<std::mutex> lk(g_mu);
        CHK(c, thb_expect_local_ctf(c, 1, &img, nR, nT, nD, q, mcp->trans, mcp->dpara, mcp->oldR, mcp->oldT, mcp->oldD, mcp->ctfK, wR, wT, wD,
                                    &uC, nullptr, nullptr));
        for (int r = 0; r < nR; ++r) wR[r] = (RFLOAT)(wR[r] * oldC);
        for (int t = 0; t < nT; ++t) wT[t] = (RFLOAT)(wT[t] * oldC);
        for (int d = 0; d < nD; ++d) wD[d] = (RFLOAT)(wD[d] * oldC);
        wC[0] = uC;
        return;
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);      // one launch per device at a time, as the caller's omp lock already guarantees
        CHK(c, thb_expect_local(c, 1, &img, nR, nT, q, mcp->trans, mcp->oldR, mcp->oldT, wR, wT, &uC, nullptr, nullptr));
    }
    // the reference's sums carry the prior weights of the OTHER dimensions (src/Optimiser.cpp:1383-1402): wC += s wR wT wD,
    // wR += s wC wT wD, ...; without CTF search there is one defocus sample of prior weight oldD[0]
    const double pD = mcp->oldD ? mcp->oldD[0] : 1.0;
    for (int r = 0; r < nR; ++r) wR[r] = (RFLOAT)(wR[r] * oldC * pD);
    for (int t = 0; t < nT; ++t) wT[t] = (RFLOAT)(wT[t] * oldC * pD);
    wC[0] = (RFLOAT)(uC * pD);
    if (wD) wD[0] = (RFLOAT)(uC * oldC);
}

void ExpectLocalFin(int gpuIdx, Complex** devdatP, RFLOAT** devctfP, RFLOAT** devdefO, RFLOAT** devfreQ, RFLOAT** devsigP, int)
{
    LocalState& L = localOf(gpuIdx);
    L = LocalState();
    if (devdatP) *devdatP = nullptr;
    if (devctfP) *devctfP = nullptr;
    if (devdefO) *devdefO = nullptr;
    if (devfreQ) *devfreQ = nullptr;
    if (devsigP) *devsigP = nullptr;
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
    // all classes in one launch, one baseline per image across the classes (thb_expect_scan_classes); the reference's layouts are
    // image-major: weightC[l * k + c], weightR[(l * k + c) * nR + r], weightT[(l * k + c) * nT + t]
    std::vector<float> cR((size_t)nK * imgNum * nR), cT((size_t)nK * imgNum * nT), cB(imgNum);
    CHK(ctx, thb_expect_scan_classes(ctx, nK, 0, imgNum, nR, nT, rot, trans, pR, pT, wC, cR.data(), cT.data(), cB.data()));
    for (int k = 0; k < nK; ++k)
        for (int l = 0; l < imgNum; ++l) {
            memcpy(wR + ((size_t)l * nK + k) * nR, cR.data() + ((size_t)k * imgNum + l) * nR, sizeof(float) * nR);
            memcpy(wT + ((size_t)l * nK + k) * nT, cT.data() + ((size_t)k * imgNum + l) * nT, sizeof(float) * nT);
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

void PrepareTF(int gpuIdx, Complex* F3D, Complex* T3D, int vdim, double* symMat, int nSymmetryElement, int maxRadius, int pf)
{
    thb_ctx* ctx = thbContext(gpuIdx);
    const size_t nVox = (size_t)(vdim / 2 + 1) * vdim * vdim;
    // normalisation first, as Reconstructor::prepareTF orders it (src/Reconstructor.cpp:1056-1091): sf = 1 / Re T[0]
    const float sf = 1.0f / T3D[0].dat[0];
    std::vector<float> F(2 * nVox), T(nVox);
    for (size_t i = 0; i < nVox; ++i) {
        F[2 * i] = F3D[i].dat[0] * sf;
        F[2 * i + 1] = F3D[i].dat[1] * sf;
        T[i] = T3D[i].dat[0] * sf;
    }
    if (nSymmetryElement > 0) {
        CHK(ctx, thb_set_mode(ctx, THB_MODE_3D));
        CHK(ctx, thb_reco_alloc(ctx, 0, vdim));
        CHK(ctx, thb_reco_upload(ctx, 0, F.data(), T.data()));
        CHK(ctx, thb_symmetrize(ctx, 0, nSymmetryElement, symMat, (double)(maxRadius * pf + 1)));
        CHK(ctx, thb_reco_download(ctx, 0, F.data(), T.data(), nullptr, nullptr, 0));
    }
    for (size_t i = 0; i < nVox; ++i) {
        F3D[i].dat[0] = F[2 * i];
        F3D[i].dat[1] = F[2 * i + 1];
        T3D[i].dat[0] = T[i];
        T3D[i].dat[1] = 0.0f;
    }
}

void InsertFT(Complex* F3D, Complex* T3D, int vdim, double* O3D, int* counter, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP,
              void* ctfaData, double* offS, RFLOAT* w, double* nR, double* nT, double* nD, int* nC, const int* iCol, const int* iRow,
              RFLOAT pixelSize, bool cSearch, int opf, int npxl, int mReco, int idim, int dimSize, int imgNum)
{
    (void)sigRcpP;
    thb_ctx* ctx = thbContext(0);
    if (cSearch && (!ctfaData || !nD)) {
        fprintf(stderr, "thunder_b200 [InsertFT]: cSearch needs ctfaData and nD\n");
        abort();
    }
    if (cSearch && nC) {
        fprintf(stderr, "thunder_b200 [InsertFT]: CTF search together with per-class draw counts is not implemented\n");
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
    if (cSearch)   // CTFAttr = 7 consecutive RFLOATs (include/Database.h:302-338): the layout thb_insert_ctf takes; one CTF per draw
        CHK(ctx, thb_insert_ctf(ctx, imgNum, nullptr, mReco, w, offS, nR, nT, nD, reinterpret_cast<const float*>(ctfaData), pixelSize));
    else if (nC)   // 3D classification: nC[l] of the mReco rows of image l belong to this class (Optimiser.cpp:6862-6950)
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
// The local search driven through the reference's own call sequence (src/Optimiser.cpp:2169-2293 set-up, :2813-3300 per image
// and phase, :3320-3400 tear-down), one image in flight per thread slot, nPhase calls of RTD / PreI3D / M per image with the
// support quat[img][phase] / tran[img][phase] the caller's Particle objects would hand over.  Results per (image, phase).
void thbi_ExpectLocalProtocol(int gpuIdx, float* volume, int vdim, int pf, int idim, int* iCol, int* iRow, int npxl, float* datP,
                              float* ctfP, float* sigRcpP, int imgNum, int nPhase, int nR, int nT, const double* quat, const double* tran,
                              const double* wRprior, const double* wTprior, double oldC, int nSlots, float* wC, float* wR, float* wT)
{
    int* deviCol = nullptr; int* deviRow = nullptr;
    ExpectPreidx(gpuIdx, &deviCol, &deviRow, iCol, iRow, npxl);
    Complex* devdatP = nullptr; RFLOAT *devctfP = nullptr, *devdefO = nullptr, *devsigP = nullptr, *devfreQ = nullptr;
    ExpectLocalIn(gpuIdx, &devdatP, &devctfP, &devdefO, &devsigP, npxl, nSlots, 1 /* SEARCH_TYPE_LOCAL */);
    ManagedArrayTexture* mgr = new ManagedArrayTexture();
    mgr->Init(1 /* MODE_3D */, vdim, gpuIdx);
    ExpectLocalV3D(gpuIdx, mgr, reinterpret_cast<Complex*>(volume), vdim);
    ManagedCalPoint* mcp = new ManagedCalPoint();
    mcp->Init(1, 1, gpuIdx, nR, nT, 1, npxl);
    RFLOAT *hwC, *hwR, *hwT, *hwD; double *oldR, *oldT, *oldD, *trans, *rot, *dpara;
    ExpectLocalHostA(gpuIdx, &hwC, &hwR, &hwT, &hwD, &oldR, &oldT, &oldD, &trans, &rot, &dpara, nR, nT, 1, 1);
    for (int l = 0; l < imgNum; ++l) {
        const int slot = l % nSlots;                 // the reference: threadId % cpyNum
        ExpectLocalP(gpuIdx, devdatP, devctfP, devdefO, devsigP, reinterpret_cast<Complex*>(datP), ctfP, nullptr, sigRcpP, slot, l, npxl, 1);
        for (int ph = 0; ph < nPhase; ++ph) {
            const size_t o = (size_t)l * nPhase + ph;
            for (int r = 0; r < nR; ++r) oldR[r] = wRprior[o * nR + r];
            for (int t = 0; t < nT; ++t) oldT[t] = wTprior[o * nT + t];
            oldD[0] = 1.0;
            memcpy(trans, tran + o * nT * 2, sizeof(double) * nT * 2);
            memcpy(rot, quat + o * nR * 4, sizeof(double) * nR * 4);
            ExpectLocalRTD(gpuIdx, mcp, oldR, oldT, oldD, trans, rot, dpara);
            ExpectLocalPreI3D(gpuIdx, slot, mgr, mcp, devdefO, devfreQ, deviCol, deviRow, 0.f, 0.1f, 0.f, 0.f, pf, idim, vdim, npxl, 1);
            ExpectLocalM(gpuIdx, slot, mcp, devdatP, devctfP, devsigP, hwC, hwR, hwT, hwD, oldC, npxl);
            wC[o] = hwC[0];
            memcpy(wR + o * nR, hwR, sizeof(float) * nR);
            memcpy(wT + o * nT, hwT, sizeof(float) * nT);
        }
    }
    ExpectLocalHostF(gpuIdx, &hwC, &hwR, &hwT, &hwD, &oldR, &oldT, &oldD, &trans, &rot, &dpara, 1);
    ExpectLocalFin(gpuIdx, &devdatP, &devctfP, &devdefO, &devfreQ, &devsigP, 1);
    delete mcp;
    delete mgr;
    ExpectFreeIdx(gpuIdx, &deviCol, &deviRow);
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
