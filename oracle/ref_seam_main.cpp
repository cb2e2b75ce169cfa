// oracle/ref_seam_main.cpp - the reference's OWN Reconstructor::insertI (src/Reconstructor.cpp:867-985), compiled from the
// reference's source with GPU_INSERT defined and THIS repository's Interface.h in place of gpu/interface/Interface.h, linked
// against libthb_interface.so the way THUNDER links libcuthuem: the proof that the M seam is a drop-in.
//
// TEST INFRASTRUCTURE ONLY (built by oracle/build_ref.sh into oracle/_ref/seam_insertI; run by tests/test_interface_shim.py on the
// GPU box).  Nothing here restates arithmetic: the call chain is Reconstructor::insertI -> InsertFT(Volume&, Volume&, ...,
// MPI_Comm&, MPI_Comm&, ...) -> libthunder_b200; the expected values come from Reconstructor::insertP on the CPU, in the test.
//
// usage: seam_insertI in.bin out.bin [point group]     (with a point group, e.g. C4: Reconstructor::prepareTFG(0) follows the
//                                                      insert - PrepareTF through the seam, src/Reconstructor.cpp:1019-1052)
//   in.bin : int32 {size, N, pf, nPxl, mReco, nImg, withCounts, cSearch}, float32 pixelSize,
//            int32 iColPad[nPxl], iRowPad[nPxl], iPxl[nPxl], iSig[nPxl],
//            float32 datP[nImg][nPxl][2], ctfP[nImg][nPxl], w[nImg], float64 offS[nImg][2], nr[nImg][mReco][4], nt[nImg][mReco][2],
//            [int32 nc[nImg] when withCounts] [float64 nd[nImg][mReco], float32 ctfAttr[nImg][7] when cSearch]
//   out.bin: int32 padSize, float32 F[sizeFT][2], T[sizeFT], float64 O[3], int32 counter
#include <cstdio>
#include <cstdlib>
#include <vector>

#define private public
#define protected public
#include "Reconstructor.h"
#include "Symmetry.h"
#undef private
#undef protected

template <class T>
static std::vector<T> rd(FILE* f, size_t n)
{
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "seam_insertI: short read\n"); exit(2); }
    return v;
}

int main(int argc, char** argv)
{
    if (argc != 3 && argc != 4) { fprintf(stderr, "usage: seam_insertI in.bin out.bin [point group]\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    std::vector<int> h = rd<int>(f, 8);
    const int size = h[0], N = h[1], pf = h[2], nPxl = h[3], mReco = h[4], nImg = h[5], withCounts = h[6], cSearch = h[7];
    const float pixelSize = rd<float>(f, 1)[0];
    std::vector<int> iCol = rd<int>(f, nPxl), iRow = rd<int>(f, nPxl), iPxl = rd<int>(f, nPxl), iSig = rd<int>(f, nPxl);
    std::vector<float> dat = rd<float>(f, (size_t)nImg * nPxl * 2), ctf = rd<float>(f, (size_t)nImg * nPxl), w = rd<float>(f, nImg);
    std::vector<double> offS = rd<double>(f, (size_t)nImg * 2), nr = rd<double>(f, (size_t)nImg * mReco * 4),
                        nt = rd<double>(f, (size_t)nImg * mReco * 2);
    std::vector<int> nc;
    std::vector<double> nd;
    std::vector<CTFAttr> attr;
    if (withCounts) nc = rd<int>(f, nImg);
    if (cSearch) {
        nd = rd<double>(f, (size_t)nImg * mReco);
        std::vector<float> a = rd<float>(f, (size_t)nImg * 7);
        attr.resize(nImg);
        for (int l = 0; l < nImg; l++) {
            attr[l].voltage = a[7 * l]; attr[l].defocusU = a[7 * l + 1]; attr[l].defocusV = a[7 * l + 2];
            attr[l].defocusTheta = a[7 * l + 3]; attr[l].Cs = a[7 * l + 4]; attr[l].amplitudeContrast = a[7 * l + 5];
            attr[l].phaseShift = a[7 * l + 6];
        }
    }
    fclose(f);

    Reconstructor reco;
    reco.setMPIEnv(3, 1, MPI_COMM_SELF, MPI_COMM_SELF);
    reco.init(MODE_3D, size, N, pf, NULL, 1.9, 15);
    reco.allocSpace(1);
    reco.setPreCal(nPxl, iCol.data(), iRow.data(), iPxl.data(), iSig.data());
    // the call expressions of Optimiser::reconstructRef (src/Optimiser.cpp:6944-6950 with nc, :7016-7019 without)
    if (withCounts)
        reco.insertI((Complex*)dat.data(), ctf.data(), NULL, w.data(), offS.data(), nr.data(), nt.data(), cSearch ? nd.data() : NULL,
                     nc.data(), cSearch ? attr.data() : NULL, pixelSize, cSearch != 0, pf, mReco, N, nImg);
    else
        reco.insertI((Complex*)dat.data(), ctf.data(), NULL, w.data(), offS.data(), nr.data(), nt.data(), cSearch ? nd.data() : NULL,
                     cSearch ? attr.data() : NULL, pixelSize, cSearch != 0, pf, mReco, N, nImg);

    if (argc == 4) {
        Symmetry sym(argv[3]);
        reco._sym = &sym;
        reco.prepareTFG(0);
        reco._sym = NULL;
    }

    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    const int pad = (int)reco._F3D.nSlcFT();
    const size_t n = reco._F3D.sizeFT();
    std::vector<float> T(n);
    for (size_t i = 0; i < n; i++) T[i] = REAL(reco._T3D[i]);
    const double O[3] = {reco._ox, reco._oy, reco._oz};
    const int counter = reco._counter;
    fwrite(&pad, sizeof(int), 1, o);
    fwrite(&reco._F3D[0], sizeof(Complex), n, o);
    fwrite(T.data(), sizeof(float), n, o);
    fwrite(O, sizeof(double), 3, o);
    fwrite(&counter, sizeof(int), 1, o);
    fclose(o);
    reco.freeSpace();
    return 0;
}
