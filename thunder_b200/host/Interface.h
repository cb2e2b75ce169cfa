// Interface.h - host-side mirror of THUNDER's accelerator seam for the Optimiser hot path.
//
// The reference's Optimiser / Reconstructor call C++ free functions declared in gpu/interface/Interface.h
// (compiled in with -DGPU_VERSION, include/Config.h:16-25).  This header declares the SAME function names
// with the same argument order and meaning for the functions on the hot path, implemented on top of the
// C ABI of libthunder_b200.so (include/thunder_b200.h).  THUNDER class types in the signatures are
// replaced by their raw storage so that the file builds without the THUNDER tree:
//     Volume&  ->  Complex* (FFTW half-complex, x fastest) + its dimension
//     MPI_Comm& hemi / slav  ->  dropped: the half-map reduction is thb_allreduce over the context's
//                                persistent NCCL communicator (thb_comm_init), a no-op with one rank
//     CTFAttr* -> void* (7 consecutive RFLOATs per image, include/Database.h:302-338), read when cSearch is set
// With -DTHB_WITH_THUNDER and the THUNDER include path, the Volume& / MPI_Comm& overloads of the
// reference are provided too (inline, forwarding to the raw versions); see INTEGRATION.md.
//
// Reference declarations mirrored (gpu/interface/Interface.h): getAviDevice :16, ExpectPreidx :18,
// ExpectFreeIdx :163, ExpectRotran :199, ExpectProject :210, ExpectGlobal3D :221, InsertFT :267-318.
// Error behaviour: the reference prints and exit(1)s on any CUDA failure (gpu/config/Device.cuh.in:27-61);
// these functions print the library's message and abort() the same way - the C ABI underneath returns codes.
#pragma once
#include <vector>
#include "../../include/thunder_b200.h"

#ifdef THB_WITH_THUNDER
#include "Precision.h"      // RFLOAT, Complex (include/Precision.h:64-106)
#else
typedef float RFLOAT;                       // default build: SINGLE_PRECISION (CMakeLists.txt:48)
// include/Precision.h:100-106, with the reference's struct tag: a typedef of a NAMED struct mangles by the tag, so the symbols
// this library exports (ExpectRotran(_complex_float_t*, ...)) are the ones THUNDER's objects reference
typedef struct _complex_float_t { float dat[2]; } Complex;
#endif

// devices this process may use (reference: every visible device with compute capability >= 3)
void getAviDevice(std::vector<int>& gpus);

// ---- E-step, pixel list (Optimiser.cpp:1702-1711).  The device arrays stay inside the library: *deviCol /
// *deviRow receive opaque non-null tokens so that caller code which only forwards them keeps working.
void ExpectPreidx(int gpuIdx, int** deviCol, int** deviRow, int* iCol, int* iRow, int npxl);
void ExpectFreeIdx(int gpuIdx, int** deviCol, int** deviRow);

// ---- E-step, global scan (Optimiser.cpp:1815-1863).  ExpectRotran / ExpectProject fill the host arrays
// the reference fills (traP, rotMat, rotP) AND remember trans / rot / volume; ExpectGlobal3D then runs the
// fused kernel on them (it does not re-read rotP / traP: the projections never leave the device).
void ExpectRotran(Complex* traP, double* trans, double* rot, double* rotMat, const int* iCol, const int* iRow, int nR,
                  int nT, int idim, int npxl);
void ExpectProject(Complex* volume, Complex* rotP, double* rotMat, const int* iCol, const int* iRow, int nR, int pf,
                   int interp, int vdim, int npxl);
void ExpectGlobal3D(Complex* rotP, Complex* traP, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP, RFLOAT* wC, RFLOAT* wR,
                    RFLOAT* wT, double* pR, double* pT, RFLOAT* baseL, int kIdx, int nK, int nR, int nT, int npxl,
                    int imgNum);

// ---- E-step, local search, THE REFERENCE'S OWN PROTOCOL (gpu/interface/Interface.h:31-164, callers src/Optimiser.cpp:2169-3300):
// one image at a time per device - ExpectLocalP copies image imgId into the device slot of the calling thread, then for every
// phase ExpectLocalRTD hands over the support (rot[mR][4], trans[mT][2]) and the prior weights (oldR, oldT, oldD),
// ExpectLocalPreI3D names the references, and ExpectLocalM returns the marginal weights wC / wR / wT / wD to the caller's host
// arrays before it returns (the caller resamples from them at once).  Same names, argument order and meaning; what differs is
// inside: the image slots, the pixel list and the projector volume live in the library's context of the device, the
// projections never exist as arrays (ExpectLocalPreI3D only records its arguments), and ExpectLocalM is ONE launch of the fused
// kernel spread over the whole chip for that image (thb_expect6.cuh).  The opaque device pointers handed back through
// Complex** / RFLOAT** are non-null tokens, as with ExpectPreidx.  CTF search (searchType / cSearch == 2): ExpectPrefre's
// frequency table, the defO of ExpectLocalP, the k1 / k2 / phaseShift / conT of ExpectLocalPreI3D and the dpara / oldD of
// ExpectLocalRTD feed the defocus dimension of the fused kernel (thb_expect_local_ctf), ExpectLocalM returns wD as well.
class ManagedArrayTexture {            // gpu/include/ManagedArrayTexture.h:13-30: here, a handle on the context's volume slot
public:
    ~ManagedArrayTexture() {}
    void Init(int mode, int vdim, int gpuIdx) { _mode = mode; _vdim = vdim; _gpu = gpuIdx; }
    int getMode() const { return _mode; }
    int getVdim() const { return _vdim; }
    int getDeviceId() const { return _gpu; }
    int slot = 0;                      // volume slot of the library context (class index in MODE_2D)
private:
    int _mode = 1, _vdim = 0, _gpu = 0;
};
class ManagedCalPoint {                // gpu/include/ManagedCalPoint.h:13-80: here, the support and priors of the image in flight
public:
    ~ManagedCalPoint() {}
    void Init(int mode, int cSearch, int gpuIdx, int nR, int nT, int mD, int npxl)
    {
        _mode = mode; _cSearch = cSearch; _gpu = gpuIdx; _nR = nR; _nT = nT; _mD = mD; _npxl = npxl;
    }
    int getMode() const { return _mode; }
    int getCSearch() const { return _cSearch; }
    int getDeviceId() const { return _gpu; }
    int getNR() const { return _nR; }
    int getNT() const { return _nT; }
    int getMD() const { return _mD; }
    // what ExpectLocalRTD / ExpectLocalPreI3D were given for the image in flight (host arrays of the caller, read by ExpectLocalM)
    const double *oldR = nullptr, *oldT = nullptr, *oldD = nullptr, *trans = nullptr, *rot = nullptr, *dpara = nullptr;
    const ManagedArrayTexture* mgr = nullptr;
    float ctfK[4] = {0, 0, 0, 0};      // k1, k2, phaseShift, conT of ExpectLocalPreI3D (CTF search)
private:
    int _mode = 1, _cSearch = 0, _gpu = 0, _nR = 0, _nT = 0, _mD = 1, _npxl = 0;
};
void ExpectPrefre(int gpuIdx, RFLOAT** devfreQ, RFLOAT* freQ, int npxl);
void ExpectLocalIn(int gpuIdx, Complex** devdatP, RFLOAT** devctfP, RFLOAT** devdefO, RFLOAT** devsigP, int nPxl, int cpyNumL,
                   int searchType);
void ExpectLocalV2D(int gpuIdx, ManagedArrayTexture* mgr, Complex* volume, int dimSize);
void ExpectLocalV3D(int gpuIdx, ManagedArrayTexture* mgr, Complex* volume, int vdim);
void ExpectLocalP(int gpuIdx, Complex* devdatP, RFLOAT* devctfP, RFLOAT* devdefO, RFLOAT* devsigP, Complex* datP, RFLOAT* ctfP,
                  RFLOAT* defO, RFLOAT* sigP, int threadId, int imgId, int npxl, int cSearch);
void ExpectLocalHostA(int gpuIdx, RFLOAT** wC, RFLOAT** wR, RFLOAT** wT, RFLOAT** wD, double** oldR, double** oldT, double** oldD,
                      double** trans, double** rot, double** dpara, int mR, int mT, int mD, int cSearch);
void ExpectLocalRTD(int gpuIdx, ManagedCalPoint* mcp, double* oldR, double* oldT, double* oldD, double* trans, double* rot,
                    double* dpara);
void ExpectLocalPreI2D(int gpuIdx, int datShift, ManagedArrayTexture* mgr, ManagedCalPoint* mcp, RFLOAT* devdefO, RFLOAT* devfreQ,
                       int* deviCol, int* deviRow, RFLOAT phaseShift, RFLOAT conT, RFLOAT k1, RFLOAT k2, int pf, int idim, int vdim,
                       int npxl, int interp);
void ExpectLocalPreI3D(int gpuIdx, int datShift, ManagedArrayTexture* mgr, ManagedCalPoint* mcp, RFLOAT* devdefO, RFLOAT* devfreQ,
                       int* deviCol, int* deviRow, RFLOAT phaseShift, RFLOAT conT, RFLOAT k1, RFLOAT k2, int pf, int idim, int vdim,
                       int npxl, int interp);
void ExpectLocalM(int gpuIdx, int datShift, ManagedCalPoint* mcp, Complex* devdatP, RFLOAT* devctfP, RFLOAT* devsigP, RFLOAT* wC,
                  RFLOAT* wR, RFLOAT* wT, RFLOAT* wD, double oldC, int npxl);
void ExpectLocalHostF(int gpuIdx, RFLOAT** wC, RFLOAT** wR, RFLOAT** wT, RFLOAT** wD, double** oldR, double** oldT, double** oldD,
                      double** trans, double** rot, double** dpara, int cSearch);
void ExpectLocalFin(int gpuIdx, Complex** devdatP, RFLOAT** devctfP, RFLOAT** devdefO, RFLOAT** devfreQ, RFLOAT** devsigP,
                    int cSearch);

// ---- E-step, local search, batched: what a maintainer should call instead where the caller can hand over many images at once
// (the reference crosses PCIe once per image per phase; INTEGRATION.md shows the patch of the loop): one call per phase.
//   quat[nImg][nR][4], tran[nImg][nT][2], wR[nImg][nR], wT[nImg][nT] priors; outputs uC[nImg], uR, uT as the
//   reference's wC / wR / wT of ExpectLocalM (weights relative to the per-image maximum), baseL[nImg]
void ExpectLocalBatch(int gpuIdx, Complex* volume, int vdim, int pf, int idim, const int* iCol, const int* iRow, int npxl,
                      Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP, int imgNum, int nR, int nT, const double* quat,
                      const double* tran, const double* wRprior, const double* wTprior, RFLOAT* wC, RFLOAT* wR, RFLOAT* wT,
                      RFLOAT* baseL);

// ---- M-step (Reconstructor::insertI, Reconstructor.cpp:867-985).  F3D / T3D are accumulated INTO, as the
// reference does (upload, insert, all-reduce over the hemisphere, download); T3D is the complex volume of the
// CPU class whose real part carries T (Interface.cpp:581-619).  vdim = F3D.nSlcFT().
void InsertFT(Complex* F3D, Complex* T3D, int vdim, double* O3D, int* counter, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP,
              void* ctfaData, double* offS, RFLOAT* w, double* nR, double* nT, double* nD, int* nC, const int* iCol,
              const int* iRow, RFLOAT pixelSize, bool cSearch, int opf, int npxl, int mReco, int idim, int dimSize,
              int imgNum);

// ---- MODE_2D (2D classification).  Same argument lists as gpu/interface/Interface.h:176-198 (ExpectGlobal2D) and :239-265
// (InsertI2D, minus the two MPI communicators: see the THB_WITH_THUNDER overload below and thb_comm_init): vol = nK padded
// class averages, rot = nR x (cos, sin); every image is compared with every class and the classes share one baseline per
// image; InsertI2D scatters each draw into the accumulator of its class nC and ADDS the result to F2D / T2D / O2D / counter.
void ExpectGlobal2D(Complex* vol, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP, double* trans, RFLOAT* wC, RFLOAT* wR, RFLOAT* wT,
                    double* pR, double* pT, double* rot, const int* iCol, const int* iRow, int nK, int nR, int nT, int pf, int interp,
                    int idim, int vdim, int npxl, int imgNum);
void InsertI2D(Complex* F2D, RFLOAT* T2D, double* O2D, int* counter, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP, RFLOAT* w, double* offS,
               int* nC, double* nR, double* nT, double* nD, void* ctfaData, const int* iCol, const int* iRow, RFLOAT pixelSize,
               bool cSearch, int nk, int opf, int npxl, int mReco, int idim, int vdim, int imgNum);

// ---- a15 through the seam: PrepareTF (gpu/interface/Interface.h:320-326, caller Reconstructor::prepareTFG,
// src/Reconstructor.cpp:1019-1052): normalise F and T by sf = 1 / Re T[0] and add their symmetry copies
// (symMat[nSymmetryElement][9], column-major dmat33 as Symmetry::get(L, R, i) returns them; rotated positions inside
// maxRadius * pf + 1), in place on the caller's host volumes.  T3D is the reference's COMPLEX volume (real part used).
void PrepareTF(int gpuIdx, Complex* F3D, Complex* T3D, int vdim, double* symMat, int nSymmetryElement, int maxRadius, int pf);

#ifdef THB_WITH_THUNDER
#include "mpi.h"
#include "Volume.h"
#include "Database.h"      // CTFAttr (include/Database.h:302), as gpu/interface/Interface.h:10 includes it
inline void InsertFT(Volume& F3D, Volume& T3D, double* O3D, int* counter, MPI_Comm&, MPI_Comm&, Complex* datP, RFLOAT* ctfP,
                     RFLOAT* sigRcpP, CTFAttr* ctfaData, double* offS, RFLOAT* w, double* nR, double* nT, double* nD,
                     const int* iCol, const int* iRow, RFLOAT pixelSize, bool cSearch, int opf, int npxl, int mReco, int idim,
                     int dimSize, int imgNum)
{
    InsertFT(&F3D[0], &T3D[0], (int)F3D.nSlcFT(), O3D, counter, datP, ctfP, sigRcpP, (void*)ctfaData, offS, w, nR, nT, nD,
             (int*)0, iCol, iRow, pixelSize, cSearch, opf, npxl, mReco, idim, dimSize, imgNum);
}
inline void InsertFT(Volume& F3D, Volume& T3D, double* O3D, int* counter, MPI_Comm&, MPI_Comm&, Complex* datP, RFLOAT* ctfP,
                     RFLOAT* sigRcpP, CTFAttr* ctfaData, double* offS, RFLOAT* w, double* nR, double* nT, double* nD, int* nC,
                     const int* iCol, const int* iRow, RFLOAT pixelSize, bool cSearch, int opf, int npxl, int mReco, int idim,
                     int dimSize, int imgNum)      // 3D classification: nC[l] draws of image l belong to this class
{
    InsertFT(&F3D[0], &T3D[0], (int)F3D.nSlcFT(), O3D, counter, datP, ctfP, sigRcpP, (void*)ctfaData, offS, w, nR, nT, nD,
             nC, iCol, iRow, pixelSize, cSearch, opf, npxl, mReco, idim, dimSize, imgNum);
}
inline void PrepareTF(int gpuIdx, Volume& F3D, Volume& T3D, double* symMat, int nSymmetryElement, int maxRadius, int pf)
{
    PrepareTF(gpuIdx, &F3D[0], &T3D[0], (int)T3D.nSlcFT(), symMat, nSymmetryElement, maxRadius, pf);
}
inline void InsertI2D(Complex* F2D, RFLOAT* T2D, double* O2D, int* counter, MPI_Comm&, MPI_Comm&, Complex* datP, RFLOAT* ctfP,
                      RFLOAT* sigRcpP, RFLOAT* w, double* offS, int* nC, double* nR, double* nT, double* nD, CTFAttr* ctfaData,
                      const int* iCol, const int* iRow, RFLOAT pixelSize, bool cSearch, int nk, int opf, int npxl, int mReco, int idim,
                      int vdim, int imgNum)
{
    InsertI2D(F2D, T2D, O2D, counter, datP, ctfP, sigRcpP, w, offS, nC, nR, nT, nD, (void*)ctfaData, iCol, iRow, pixelSize, cSearch, nk,
              opf, npxl, mReco, idim, vdim, imgNum);
}
#endif

// the library context of device gpuIdx (created on first use; one per device per process)
thb_ctx* thbContext(int gpuIdx);
void thbShutdown();
