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
