// compile-only: the reference-typed overloads of thunder_b200/host/Interface.h against THUNDER's own headers
#include "Interface.h"
#include "Reconstructor.h"
void drive(Volume& F3D, Volume& T3D, MPI_Comm& hemi, MPI_Comm& slav, Complex* datP, RFLOAT* ctfP, RFLOAT* sigP, CTFAttr* ca,
           double* offS, RFLOAT* w, double* nr, double* nt, double* nd, int* nc, const int* iCol, const int* iRow)
{
    double O3D[3] = {0, 0, 0};
    int counter[1] = {0};
    InsertFT(F3D, T3D, O3D, counter, hemi, slav, datP, ctfP, sigP, ca, offS, w, nr, nt, nd, iCol, iRow, 1.32f, false, 2, 100, 10, 64,
             (int)T3D.sizeFT(), 4);
    InsertFT(F3D, T3D, O3D, counter, hemi, slav, datP, ctfP, sigP, ca, offS, w, nr, nt, nd, nc, iCol, iRow, 1.32f, false, 2, 100, 10, 64,
             (int)T3D.sizeFT(), 4);
    RFLOAT T2D[4]; Complex F2D[4]; double O2D[2];
    InsertI2D(F2D, T2D, O2D, counter, hemi, slav, datP, ctfP, sigP, w, offS, nc, nr, nt, nd, ca, iCol, iRow, 1.32f, false, 1, 2, 100, 10,
              64, 128, 4);
}

// the E-step seam as Optimiser::expectationG calls it (src/Optimiser.cpp:1702-1711, 1815-1920)
void drive_expect(Complex* vol, Complex* rotP, Complex* traP, Complex* datP, RFLOAT* ctfP, RFLOAT* sigRcpP, double* trans, double* rot,
                  double* rotMat, RFLOAT* wC, RFLOAT* wR, RFLOAT* wT, double* pR, double* pT, RFLOAT* baseL, int* iCol, int* iRow)
{
    std::vector<int> gpus;
    getAviDevice(gpus);
    int* dCol = 0; int* dRow = 0;
    ExpectPreidx(0, &dCol, &dRow, iCol, iRow, 100);
    ExpectRotran(traP, trans, rot, rotMat, iCol, iRow, 10, 5, 64, 100);
    ExpectProject(vol, rotP, rotMat, iCol, iRow, 10, 2, 1, 128, 100);
    ExpectGlobal3D(rotP, traP, datP, ctfP, sigRcpP, wC, wR, wT, pR, pT, baseL, 0, 1, 10, 5, 100, 4);
    ExpectGlobal2D(vol, datP, ctfP, sigRcpP, trans, wC, wR, wT, pR, pT, rot, iCol, iRow, 2, 10, 5, 2, 1, 64, 128, 100, 4);
    ExpectFreeIdx(0, &dCol, &dRow);
}

// the local-search seam in the order Optimiser::expectationG calls it (src/Optimiser.cpp:2169-2342 set-up, 2484-2700 per phase)
void drive_local(Complex* vol, Complex* datP, RFLOAT* ctfP, RFLOAT* defO, RFLOAT* sigRcpP, RFLOAT* freQ, int* iCol, int* iRow)
{
    int* dCol = 0; int* dRow = 0;
    ExpectPreidx(0, &dCol, &dRow, iCol, iRow, 100);
    RFLOAT* devfreQ = 0;
    ExpectPrefre(0, &devfreQ, freQ, 100);
    ManagedArrayTexture* mgr = new ManagedArrayTexture();
    mgr->Init(1, 128, 0);
    Complex* devdatP = 0; RFLOAT* devctfP = 0; RFLOAT* devdefO = 0; RFLOAT* devsigP = 0;
    ExpectLocalIn(0, &devdatP, &devctfP, &devdefO, &devsigP, 100, 4, 1);
    ManagedCalPoint* mcp = new ManagedCalPoint();
    mcp->Init(1, 1, 0, 125, 9, 1, 100);
    RFLOAT *wC, *wR, *wT, *wD; double *oldR, *oldT, *oldD, *trans, *rot, *dpara;
    ExpectLocalHostA(0, &wC, &wR, &wT, &wD, &oldR, &oldT, &oldD, &trans, &rot, &dpara, 125, 9, 1, 1);
    ExpectLocalV3D(0, mgr, vol, 128);
    ExpectLocalV2D(0, mgr, vol, 128 * 65);
    ExpectLocalP(0, devdatP, devctfP, devdefO, devsigP, datP, ctfP, defO, sigRcpP, 0, 3, 100, 1);
    ExpectLocalRTD(0, mcp, oldR, oldT, oldD, trans, rot, dpara);
    ExpectLocalPreI3D(0, 0, mgr, mcp, devdefO, devfreQ, dCol, dRow, 0.f, 0.1f, 0.f, 0.f, 2, 64, 128, 100, 1);
    ExpectLocalPreI2D(0, 0, mgr, mcp, devdefO, devfreQ, dCol, dRow, 0.f, 0.1f, 0.f, 0.f, 2, 64, 128, 100, 1);
    ExpectLocalM(0, 0, mcp, devdatP, devctfP, devsigP, wC, wR, wT, wD, 1.0, 100);
    ExpectLocalHostF(0, &wC, &wR, &wT, &wD, &oldR, &oldT, &oldD, &trans, &rot, &dpara, 1);
    ExpectLocalFin(0, &devdatP, &devctfP, &devdefO, &devfreQ, &devsigP, 1);
    delete mcp;
    delete mgr;
    ExpectFreeIdx(0, &dCol, &dRow);
}
