// thb_pf.cu - device-resident particle filter and the iteration-level drivers built on it:
// thb_expectation (phase loop of Optimiser::expectation, src/Optimiser.cpp:1162-1660) and
// thb_reconstruct_insert (insert loop of Optimiser::reconstructRef, src/Optimiser.cpp:7036-7241).
// One CUDA thread per particle runs the serial operators of thb_pf.cuh on the SoA state; the fused
// E kernel reads rotations / translations / priors straight from that state and writes the marginal
// weights back, so a whole E-step needs no host round trip per phase (the reference's GPU path
// crosses PCIe per image per phase, src/Optimiser.cpp:2813-3300).
#include <cstring>
#include <vector>
#include "thb_context.h"
#include "thb_pf.cuh"
#include "thb_pf2d.cuh"

namespace thb {

struct PFDev {
    double *r, *t, *wR, *wT, *uR, *uT, *scal, *r2, *t2, *w2, *w3, *w4, *dd, *wD, *uD;
    const float *uRf, *uTf, *uDf;
    unsigned char* active;
    int* nPhase;
    int* activeCount;
    int nPar, mLR, mLT, mLD, mode2D;
    uint64_t seed, streamBase;
};

// the particle-filter kernels run ONE WARP PER PARTICLE: every lane executes the serial operators redundantly on
// the same particle (identical random stream, identical values, uniform control flow), and the one hot loop - the
// ACG fixed-point inference over the support points - is spread over the lanes (pf::infer_acg_warp)
constexpr int PF_BLOCK = 128;   // 4 particles per CTA
__device__ __forceinline__ int pf_particle() { return (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5); }

__device__ __forceinline__ pf::View make_view(const PFDev& d, int p)
{
    pf::View v;
    v.lane = (int)(threadIdx.x & 31);
    v.r = d.r; v.t = d.t; v.wR = d.wR; v.wT = d.wT; v.uR = d.uR; v.uT = d.uT; v.scal = d.scal;
    v.r2 = d.r2; v.t2 = d.t2; v.w2 = d.w2; v.w3 = d.w3; v.w4 = d.w4;
    v.n = d.nPar; v.p = p; v.mLR = d.mLR; v.mLT = d.mLT;
    v.d = d.dd; v.wD = d.wD; v.uD = d.uD; v.mLD = d.mLD; v.mode2D = d.mode2D;
    return v;
}

__global__ void pf_load_kernel(PFDev d, uint64_t epoch, const double* quat, const double* k123, const double* tran,
                               const double* s01)
{
    const int p = pf_particle();
    if (p >= d.nPar) return;
    pf::View v = make_view(d, p);
    pf::Rng g;
    g.init(d.seed, d.streamBase + p, epoch);
    pf::load(v, quat + 4 * p, k123[3 * p], k123[3 * p + 1], k123[3 * p + 2], tran + 2 * p, s01[2 * p], s01[2 * p + 1], g);
    d.active[p] = 1;
    d.nPhase[p] = 0;
}

// begin an E-step: per-iteration stop-rule state
__global__ void pf_begin_kernel(PFDev d)
{
    const int p = pf_particle();
    if (p >= d.nPar) return;
    pf::View v = make_view(d, p);
    v.S(pf::S_VARIR) = 1.79769313486231570e308;
    v.S(pf::S_VARIT) = 1.79769313486231570e308;
    v.S(pf::S_VARID) = 1.79769313486231570e308;
    v.S(pf::S_NODEC) = 0.0;
    v.S(pf::S_NPHASE) = 0.0;
    d.active[p] = 1;
    d.nPhase[p] = 0;
}

struct StepArgs {
    int doPost, doPre;
    int phase;              // phase whose likelihoods were just computed (doPost)
    double prePf;           // perturbation factor of the next perturb (doPre)
    double ctfRefineS, prePfD;   // CTF search: sigma of initD (first perturbation), perturbation factor of PAR_D (later ones)
    double* traceRes;       // optional [nPar][4 mLR + 2 mLT]: the support after the resampling (doPost)
    double* tracePert;      // optional, same shape: the support after the perturbation (doPre)
    double transS, transQ;
    int minPhase, maxPhase, fixedPhases, noDecreaseLimit;
    double decreaseFactor;
    uint64_t epoch;
};

// support of particle v.p -> dst[p][4 mLR + 2 mLT] as r[mLR][4], t[mLT][2] (test trace; every lane writes the same values)
__device__ void pf_trace_store(const pf::View& v, double* dst, long long p)
{
    double* o = dst + (size_t)p * (4 * v.mLR + 2 * v.mLT);
    for (int i = 0; i < v.mLR; ++i)
        for (int c = 0; c < 4; ++c) o[4 * i + c] = v.R(i, c);
    o += 4 * v.mLR;
    for (int i = 0; i < v.mLT; ++i)
        for (int c = 0; c < 2; ++c) o[2 * i + c] = v.T(i, c);
}

// stable compaction of the particles still active into order[0 .. count): one CTA, chunks of its size with a running offset
__global__ void __launch_bounds__(1024) pf_compact_kernel(const unsigned char* __restrict__ active, int nPar, int* __restrict__ order,
                                                         int* __restrict__ count)
{
    __shared__ int sWarp[32];
    __shared__ int sBase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sBase = 0;
    __syncthreads();
    for (int p0 = 0; p0 < nPar; p0 += 1024) {
        const int p = p0 + tid;
        const int a = (p < nPar && active[p]) ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, a);
        if (lane == 0) sWarp[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const int v = sWarp[w];
            if (w < warp) before += v;
            total += v;
        }
        if (a) order[sBase + before + __popc(m & ((1u << lane) - 1u))] = p;
        __syncthreads();
        if (tid == 0) sBase += total;
        __syncthreads();
    }
    if (tid == 0) *count = sBase;
}

// The operators walk the particle's state element by element in one serial chain; in global memory (SoA, particle index
// fastest) every access is a dependent L2 round trip.  pf_step_kernel therefore stages the state of its particle in shared
// memory (one block of doubles per warp), runs the operators there and writes the persistent arrays back.
__host__ __device__ inline int pf_stage_doubles(int mLR, int mLT, int mLD)
{
    const int mx = mLR > mLT ? mLR : mLT;
    return 4 * mLR + 2 * mLT + mLR + mLT + mLR + mLT + pf::S_COUNT + 4 * mLR + 2 * mLT + mLR + 2 * mx + (mLD > 0 ? 3 * mLD + 1 : 0);
}

struct PFStage {
    pf::View g, s;          // the particle in global memory, and its staged copy (n = 1, p = 0)
};

__device__ __forceinline__ void pf_copy_rows(double* dst, long long dn, long long dp, const double* src, long long sn, long long sp,
                                             int rows, int lane)
{
    for (int i = lane; i < rows; i += 32) dst[(long long)i * dn + dp] = src[(long long)i * sn + sp];
}

__device__ __forceinline__ PFStage pf_stage_in(const PFDev& d, int p, double* sm)
{
    PFStage st;
    st.g = make_view(d, p);
    pf::View v = st.g;
    const int mLR = d.mLR, mLT = d.mLT, mLD = d.mLD, mx = mLR > mLT ? mLR : mLT;
    v.n = 1; v.p = 0;
    v.r = sm; sm += 4 * mLR;
    v.t = sm; sm += 2 * mLT;
    v.wR = sm; sm += mLR;
    v.wT = sm; sm += mLT;
    v.uR = sm; sm += mLR;
    v.uT = sm; sm += mLT;
    v.scal = sm; sm += pf::S_COUNT;
    v.r2 = sm; sm += 4 * mLR;
    v.t2 = sm; sm += 2 * mLT;
    v.w2 = sm; sm += mLR;
    v.w3 = sm; sm += mx;
    v.w4 = sm; sm += mx;
    if (mLD > 0) {
        v.d = sm; sm += mLD + 1;
        v.wD = sm; sm += mLD;
        v.uD = sm; sm += mLD;
    }
    st.s = v;
    const pf::View& g = st.g;
    const int lane = v.lane;
    pf_copy_rows(v.r, 1, 0, g.r, g.n, g.p, 4 * mLR, lane);
    pf_copy_rows(v.t, 1, 0, g.t, g.n, g.p, 2 * mLT, lane);
    pf_copy_rows(v.wR, 1, 0, g.wR, g.n, g.p, mLR, lane);
    pf_copy_rows(v.wT, 1, 0, g.wT, g.n, g.p, mLT, lane);
    pf_copy_rows(v.uR, 1, 0, g.uR, g.n, g.p, mLR, lane);
    pf_copy_rows(v.uT, 1, 0, g.uT, g.n, g.p, mLT, lane);
    pf_copy_rows(v.scal, 1, 0, g.scal, g.n, g.p, pf::S_COUNT, lane);
    if (mLD > 0) {
        pf_copy_rows(v.d, 1, 0, g.d, g.n, g.p, mLD + 1, lane);
        pf_copy_rows(v.wD, 1, 0, g.wD, g.n, g.p, mLD, lane);
        pf_copy_rows(v.uD, 1, 0, g.uD, g.n, g.p, mLD, lane);
    }
    __syncwarp();
    return st;
}

__device__ __forceinline__ void pf_stage_out(const PFStage& st)
{
    const pf::View& g = st.g;
    const pf::View& v = st.s;
    const int mLR = v.mLR, mLT = v.mLT, mLD = v.mLD, lane = v.lane;
    __syncwarp();
    pf_copy_rows(g.r, g.n, g.p, v.r, 1, 0, 4 * mLR, lane);
    pf_copy_rows(g.t, g.n, g.p, v.t, 1, 0, 2 * mLT, lane);
    pf_copy_rows(g.wR, g.n, g.p, v.wR, 1, 0, mLR, lane);
    pf_copy_rows(g.wT, g.n, g.p, v.wT, 1, 0, mLT, lane);
    pf_copy_rows(g.uR, g.n, g.p, v.uR, 1, 0, mLR, lane);
    pf_copy_rows(g.uT, g.n, g.p, v.uT, 1, 0, mLT, lane);
    pf_copy_rows(g.scal, g.n, g.p, v.scal, 1, 0, pf::S_COUNT, lane);
    if (mLD > 0) {
        pf_copy_rows(g.d, g.n, g.p, v.d, 1, 0, mLD + 1, lane);
        pf_copy_rows(g.wD, g.n, g.p, v.wD, 1, 0, mLD, lane);
        pf_copy_rows(g.uD, g.n, g.p, v.uD, 1, 0, mLD, lane);
    }
}

__global__ void pf_step_kernel(PFDev d, StepArgs a, int staged)
{
    extern __shared__ double pf_smem[];
    const int p = pf_particle();
    if (p >= d.nPar) return;
    if (!d.active[p]) return;
    PFStage st;
    if (staged)
        st = pf_stage_in(d, p, pf_smem + (size_t)(threadIdx.x >> 5) * pf_stage_doubles(d.mLR, d.mLT, d.mLD));
    else
        st.g = st.s = make_view(d, p);
    pf::View& v = st.s;
    pf::Rng g;
    g.init(d.seed, d.streamBase + p, a.epoch);
    bool cont = true;
    if (a.doPost) {
        pf::set_u_keep_peak(v, d.uRf + (size_t)p * d.mLR, d.uTf + (size_t)p * d.mLT);
        if (d.mLD > 0)
            for (int i = 0; i < d.mLD; ++i) v.UD(i) = (double)d.uDf[(size_t)p * d.mLD + i];
        pf::rank1st(v);
        if (d.mode2D) { pf::cal_vari_R_2d(v); pf::cal_vari_T(v); }
        else pf::cal_vari(v, g);
        pf::resample_R(v, g);
        pf::resample_T(v, g);
        if (d.mLD > 0) {     // SEARCH_TYPE_CTF: calRank1st / calVari / resample of PAR_D after those of R and T (src/Optimiser.cpp:1483-1488)
            pf::rank1st_D(v);
            pf::cal_vari_D(v);
            pf::resample_D(v, g);
        }
        pf::norm_w(v);
        if (a.traceRes) pf_trace_store(v, a.traceRes, p);
        d.nPhase[p] = a.phase + 1;
        v.S(pf::S_NPHASE) = (double)(a.phase + 1);
        bool done;
        if (a.fixedPhases > 0)
            done = a.phase + 1 >= a.fixedPhases;
        else
            done = pf::stop_rule(v, a.phase, a.minPhase, a.decreaseFactor, a.noDecreaseLimit) || a.phase + 1 >= a.maxPhase;
        if (done) {
            d.active[p] = 0;
            v.S(pf::S_SCORE) = pf::compress_R(v);   // calScore() at the start of reconstructRef
            cont = false;
        }
    }
    if (cont && a.doPre) {
        if (d.mode2D) { pf::perturb_R_2d(v, a.prePf, g); pf::balance_R_2d(v); pf::norm_w(v); }
        else pf::perturb_R(v, a.prePf, g);
        pf::perturb_T(v, a.prePf, a.transS, a.transQ, g);
        if (d.mLD > 0) {     // src/Optimiser.cpp:1193-1215
            if (a.phase < 0) pf::init_D(v, a.ctfRefineS, g);
            else pf::perturb_D(v, a.prePfD, g);
        }
        if (a.tracePert) pf_trace_store(v, a.tracePert, p);
        if ((threadIdx.x & 31) == 0) atomicAdd(d.activeCount, 1);
    }
    if (staged) pf_stage_out(st);
}

// Particle::rand(cls, quat, tran, d) x mReco: independent uniform draws of support indices
__global__ void pf_draw_kernel(PFDev d, uint64_t epoch, int mReco, int parGra, int* drawR, int* drawT, int* drawD, float* w)
{
    const int p = pf_particle();
    if (p >= d.nPar) return;
    pf::View v = make_view(d, p);
    pf::Rng g;
    g.init(d.seed, d.streamBase + p, epoch);
    for (int m = 0; m < mReco; ++m) {
        (void)g.uniform_int(1);                                   // class
        drawR[(size_t)p * mReco + m] = (int)g.uniform_int((uint32_t)d.mLR);
        drawT[(size_t)p * mReco + m] = (int)g.uniform_int((uint32_t)d.mLT);
        const int iD = (int)g.uniform_int((uint32_t)(d.mLD > 0 ? d.mLD : 1));   // defocus
        if (drawD) drawD[(size_t)p * mReco + m] = iD;
    }
    const double ww = parGra ? pf::compress_R(v) : 1.0;
    w[p] = (float)ww / (float)mReco;                              // RFLOAT w; w /= mReco
}

__global__ void pf_op_kernel(PFDev d, int op, double arg, double transS, double transQ, uint64_t epoch)
{
    const int p = pf_particle();
    if (p >= d.nPar) return;
    pf::View v = make_view(d, p);
    pf::Rng g;
    g.init(d.seed, d.streamBase + p, epoch);
    switch (op) {
        case THB_PF_PERTURB_R:
            if (d.mode2D) { pf::perturb_R_2d(v, arg, g); pf::balance_R_2d(v); pf::norm_w(v); }
            else pf::perturb_R(v, arg, g);
            break;
        case THB_PF_PERTURB_T: pf::perturb_T(v, arg, transS, transQ, g); break;
        case THB_PF_SET_U_KEEP_PEAK: pf::set_u_keep_peak(v, d.uRf + (size_t)p * d.mLR, d.uTf + (size_t)p * d.mLT); break;
        case THB_PF_RANK1ST: pf::rank1st(v); break;
        case THB_PF_CALVARI:
            if (d.mode2D) { pf::cal_vari_R_2d(v); pf::cal_vari_T(v); }
            else pf::cal_vari(v, g);
            break;
        case THB_PF_RESAMPLE: pf::resample_R(v, g); pf::resample_T(v, g); pf::norm_w(v); break;
        case THB_PF_BALANCE_R:
            if (d.mode2D) { pf::balance_R_2d(v); pf::norm_w(v); }
            else pf::balance_R(v);
            break;
        case THB_PF_BALANCE_T: pf::balance_T(v); break;
        default: break;
    }
}

// one THREAD per particle (serial operators, v.lane = -1): the hand-over from the global scan is once per iteration
struct ScanArgs {
    int nK, nR, nT, qc, nMax, p0, pN;
    const double* gridR; const double* gridT;
    const float* wC; const float* wR; const float* wT;     // [nPar][nK], [nK][nPar][nR], [nK][nPar][nT]
    double kFloor, sFloor;
    double* scratch;        // [chunk][3 nMax]
    int* idx;               // [chunk][mLR + mLT]
    int* cls;               // [nPar]
    uint64_t epoch;
};
__global__ void pf_from_scan_kernel(PFDev d, ScanArgs a)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = a.p0 + q;
    if (q >= a.pN) return;
    pf::View v = make_view(d, p);
    v.lane = -1;
    pf::Rng g;
    g.init(d.seed, d.streamBase + p, a.epoch);
    double* sc = a.scratch + (size_t)q * 3 * a.nMax;
    int* ix = a.idx + (size_t)q * (d.mLR + d.mLT);
    const int cls = pf::from_scan(v, g, d.mode2D, a.nK, a.nR, a.nT, a.gridR, a.qc, a.gridT, a.wC + (size_t)p * a.nK,
                                  a.wR + (size_t)p * a.nR, (size_t)d.nPar * a.nR, a.wT + (size_t)p * a.nT, (size_t)d.nPar * a.nT, a.kFloor,
                                  a.sFloor, sc, sc + a.nMax, sc + 2 * a.nMax, ix, ix + d.mLR);
    a.cls[p] = cls;
    d.active[p] = 1;
    d.nPhase[p] = 0;
}

__global__ void pf_set_slots_kernel(const int* cls, int n, int imgBase, int* slotE, int* slotM)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (slotE) slotE[imgBase + p] = cls[p];
    if (slotM) slotM[imgBase + p] = cls[p];
}

void pf_free(thb_ctx* ctx)
{
    PFState& s = ctx->pf_;
    cudaFree(s.r); cudaFree(s.t); cudaFree(s.wR); cudaFree(s.wT); cudaFree(s.scal);
    cudaFree(s.uR); cudaFree(s.uT); cudaFree(s.uC); cudaFree(s.base); cudaFree(s.active); cudaFree(s.order); cudaFree(s.nPhase);
    cudaFree(s.vari); cudaFree(s.drawR); cudaFree(s.drawT); cudaFree(s.drawD);
    cudaFree(s.d); cudaFree(s.wD); cudaFree(s.uDd); cudaFree(s.uD); cudaFree(s.ctfK); cudaFree(s.ctfAttr);
    cudaFree(s.dbl); cudaFree(s.traceR); cudaFree(s.traceT); cudaFree(s.traceSt); cudaFree(s.traceB);
    s = PFState();
}

static PFDev dev_view(thb_ctx* ctx)
{
    PFState& s = ctx->pf_;
    PFDev d;
    const size_t n = s.nPar;
    d.r = s.r; d.t = s.t; d.wR = s.wR; d.wT = s.wT; d.scal = s.scal;
    // dbl: [uR mLR][uT mLT][r2 4 mLR][t2 2 mLT][w2, w3, w4: max(mLR,mLT) each] x nPar doubles
    double* q = s.dbl;
    d.uR = q; q += n * s.prm.mLR;
    d.uT = q; q += n * s.prm.mLT;
    d.r2 = q; q += n * 4 * s.prm.mLR;
    d.t2 = q; q += n * 2 * s.prm.mLT;
    const size_t mw = s.prm.mLR > s.prm.mLT ? s.prm.mLR : s.prm.mLT;
    d.w2 = q; q += n * mw;
    d.w3 = q; q += n * mw;
    d.w4 = q;
    d.uRf = s.uR; d.uTf = s.uT; d.uDf = s.uD;
    d.dd = s.d; d.wD = s.wD; d.uD = s.uDd; d.mLD = s.prm.mLD; d.mode2D = s.mode2D;
    d.active = s.active; d.nPhase = s.nPhase; d.activeCount = (int*)s.vari;
    d.nPar = s.nPar; d.mLR = s.prm.mLR; d.mLT = s.prm.mLT;
    d.seed = s.prm.seed; d.streamBase = s.streamBase;
    return d;
}

static int pf_alloc(thb_ctx* ctx, int nPar, const thb_pf_params& p)
{
    PFState& s = ctx->pf_;
    if (s.nPar == nPar && s.prm.mLR == p.mLR && s.prm.mLT == p.mLT && s.prm.mLD == p.mLD && s.r) {
        s.prm = p;
        return THB_OK;
    }
    {   // a reallocation must not lose the image pairing and the position in the random stream (thb_pf_set_image_base may
        // legitimately precede the first thb_pf_load)
        const int imgBase = s.imgBase, traceWant = s.traceWant;
        const uint64_t streamBase = s.streamBase, epoch = s.epoch;
        pf_free(ctx);
        s.imgBase = imgBase; s.streamBase = streamBase; s.epoch = epoch; s.traceWant = traceWant;
    }
    const size_t n = nPar;
    const int mw = p.mLR > p.mLT ? p.mLR : p.mLT;
    THB_CUDA(ctx, cudaMalloc(&s.r, sizeof(double) * n * 4 * p.mLR));
    THB_CUDA(ctx, cudaMalloc(&s.t, sizeof(double) * n * 2 * p.mLT));
    THB_CUDA(ctx, cudaMalloc(&s.wR, sizeof(double) * n * p.mLR));
    THB_CUDA(ctx, cudaMalloc(&s.wT, sizeof(double) * n * p.mLT));
    THB_CUDA(ctx, cudaMalloc(&s.scal, sizeof(double) * n * pf::S_COUNT));
    THB_CUDA(ctx, cudaMalloc(&s.dbl, sizeof(double) * n * (size_t)(5 * p.mLR + 3 * p.mLT + 3 * mw)));
    THB_CUDA(ctx, cudaMalloc(&s.uR, sizeof(float) * n * p.mLR));
    THB_CUDA(ctx, cudaMalloc(&s.uT, sizeof(float) * n * p.mLT));
    THB_CUDA(ctx, cudaMalloc(&s.uC, sizeof(float) * n));
    THB_CUDA(ctx, cudaMalloc(&s.base, sizeof(float) * n));
    THB_CUDA(ctx, cudaMalloc(&s.active, n));
    THB_CUDA(ctx, cudaMalloc(&s.order, sizeof(int) * n));
    THB_CUDA(ctx, cudaMalloc(&s.nPhase, sizeof(int) * n));
    THB_CUDA(ctx, cudaMalloc(&s.vari, sizeof(double) * 4));
    THB_CUDA(ctx, cudaMemset(s.scal, 0, sizeof(double) * n * pf::S_COUNT));
    if (p.mLD > 0) {
        THB_CUDA(ctx, cudaMalloc(&s.d, sizeof(double) * n * (p.mLD + 1)));
        THB_CUDA(ctx, cudaMalloc(&s.wD, sizeof(double) * n * p.mLD));
        THB_CUDA(ctx, cudaMalloc(&s.uDd, sizeof(double) * n * p.mLD));
        THB_CUDA(ctx, cudaMalloc(&s.uD, sizeof(float) * n * p.mLD));
        THB_CUDA(ctx, cudaMalloc(&s.ctfK, sizeof(float) * n * 4));
        THB_CUDA(ctx, cudaMalloc(&s.ctfAttr, sizeof(float) * n * 7));
        THB_CUDA(ctx, cudaMemset(s.d, 0, sizeof(double) * n * (p.mLD + 1)));
        THB_CUDA(ctx, cudaMemset(s.wD, 0, sizeof(double) * n * p.mLD));
        THB_CUDA(ctx, cudaMemset(s.uDd, 0, sizeof(double) * n * p.mLD));
    }
    s.nPar = nPar;
    s.prm = p;
    return THB_OK;
}

// [nPar][S][C] (API, host) <-> [(c*S + i)*nPar + p] (device SoA)
static void to_soa(const double* aos, double* soa, size_t nPar, int S, int Cn)
{
#pragma omp parallel for
    for (long long p = 0; p < (long long)nPar; ++p)
        for (int i = 0; i < S; ++i)
            for (int c = 0; c < Cn; ++c) soa[((size_t)c * S + i) * nPar + p] = aos[((size_t)p * S + i) * Cn + c];
}
static void to_aos(const double* soa, double* aos, size_t nPar, int S, int Cn)
{
#pragma omp parallel for
    for (long long p = 0; p < (long long)nPar; ++p)
        for (int i = 0; i < S; ++i)
            for (int c = 0; c < Cn; ++c) aos[((size_t)p * S + i) * Cn + c] = soa[((size_t)c * S + i) * nPar + p];
}

}  // namespace thb

using namespace thb;

static inline int nblk(int n) { return (n * 32 + PF_BLOCK - 1) / PF_BLOCK; }   // one warp per particle

extern "C" {

int thb_pf_load(thb_ctx* ctx, int nPar, const thb_pf_params* p, const double* quat, const double* k123, const double* tran,
                const double* s01)
{
    if (!ctx) return THB_E_ARG;
    if (nPar <= 0 || !p || !quat || !k123 || !tran || !s01 || p->mLR < 2 || p->mLT < 2)
        return set_error(ctx, THB_E_ARG, "pf_load: bad arguments");
    if (p->mLD < 0 || p->mLD > (p->mLR > p->mLT ? p->mLR : p->mLT)) return set_error(ctx, THB_E_ARG, "pf_load: mLD must be in [0, max(mLR, mLT)]");
    if (ctx->mode2D) return set_error(ctx, THB_E_STATE, "pf_load: the device particle filter is MODE_3D only; drive MODE_2D through thb_expect_local / thb_expect_scan");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = pf_alloc(ctx, nPar, *p);
    if (rc) return rc;
    PFState& s = ctx->pf_;
    s.ctfSet = false;                 // the CTF constants belong to the particles of one load
    s.mode2D = 0;
    const size_t n = nPar;
    double* din = (double*)scratch(ctx, 0, sizeof(double) * n * 11);
    if (!din) return THB_E_CUDA;
    THB_CUDA(ctx, cudaMemcpyAsync(din, quat, sizeof(double) * n * 4, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(din + 4 * n, k123, sizeof(double) * n * 3, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(din + 7 * n, tran, sizeof(double) * n * 2, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(din + 9 * n, s01, sizeof(double) * n * 2, cudaMemcpyHostToDevice, ctx->stream));
    s.epoch += 1;
    span_begin(ctx, KF_PF);
    pf_load_kernel<<<nblk(nPar), PF_BLOCK, 0, ctx->stream>>>(dev_view(ctx), s.epoch << 20, din, din + 4 * n, din + 7 * n, din + 9 * n);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_pf_from_scan(thb_ctx* ctx, int nPar, const thb_pf_params* p, int nK, int nR, int nT, const double* quat, const double* tran,
                     const float* wC, const float* wR, const float* wT, double kFloor, double sFloor, int* clsOut)
{
    if (!ctx) return THB_E_ARG;
    if (nPar <= 0 || !p || nK <= 0 || nK > THB_MAX_SLOTS || nR < 2 || nT < 2 || !quat || !tran || !wC || !wR || !wT || p->mLR < 2 || p->mLT < 2)
        return set_error(ctx, THB_E_ARG, "pf_from_scan: bad arguments");
    if (p->mLD != 0) return set_error(ctx, THB_E_ARG, "pf_from_scan: the CTF search starts from thb_pf_load (mLD must be 0 here)");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    {
        const int imgBase = ctx->pf_.imgBase;
        if (imgBase < 0 || imgBase + nPar > ctx->stackE.nImg) return set_error(ctx, THB_E_STATE, "pf_from_scan: particles [%d,%d) exceed the E stack", imgBase, imgBase + nPar);
    }
    int rc = pf_alloc(ctx, nPar, *p);
    if (rc) return rc;
    PFState& s = ctx->pf_;
    s.mode2D = ctx->mode2D;
    s.ctfSet = false;
    const int qc = ctx->mode2D ? 2 : 4;
    const int nMax = std::max(std::max(nR, nT), 4 * nK);
    const size_t n = nPar;
    const size_t inBytes = sizeof(double) * ((size_t)nR * qc + (size_t)nT * 2) + sizeof(float) * (n * nK + (size_t)nK * n * nR + (size_t)nK * n * nT) + sizeof(int) * n;
    unsigned char* din = (unsigned char*)scratch(ctx, 0, inBytes + 64);
    if (!din) return THB_E_CUDA;
    double* dR = (double*)din; double* dT = dR + (size_t)nR * qc;
    float* dwC = (float*)(dT + (size_t)nT * 2); float* dwR = dwC + n * nK; float* dwT = dwR + (size_t)nK * n * nR;
    int* dcls = (int*)(dwT + (size_t)nK * n * nT);
    THB_CUDA(ctx, cudaMemcpyAsync(dR, quat, sizeof(double) * (size_t)nR * qc, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dT, tran, sizeof(double) * (size_t)nT * 2, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwC, wC, sizeof(float) * n * nK, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwR, wR, sizeof(float) * (size_t)nK * n * nR, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwT, wT, sizeof(float) * (size_t)nK * n * nT, cudaMemcpyHostToDevice, ctx->stream));
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>(n, ((size_t)256 << 20) / (sizeof(double) * 3 * nMax + sizeof(int) * (p->mLR + p->mLT))));
    double* dsc = (double*)scratch(ctx, 1, sizeof(double) * 3 * (size_t)nMax * chunk);
    int* dix = (int*)scratch(ctx, 2, sizeof(int) * (size_t)(p->mLR + p->mLT) * chunk);
    if (!dsc || !dix) return THB_E_CUDA;
    s.epoch += 1;
    ScanArgs a;
    memset(&a, 0, sizeof(a));
    a.nK = nK; a.nR = nR; a.nT = nT; a.qc = qc; a.nMax = nMax;
    a.gridR = dR; a.gridT = dT; a.wC = dwC; a.wR = dwR; a.wT = dwT; a.kFloor = kFloor; a.sFloor = sFloor;
    a.scratch = dsc; a.idx = dix; a.cls = dcls; a.epoch = s.epoch << 20;
    span_begin(ctx, KF_PF);
    for (int p0 = 0; p0 < nPar; p0 += chunk) {
        a.p0 = p0; a.pN = std::min(chunk, nPar - p0);
        pf_from_scan_kernel<<<(a.pN + 63) / 64, 64, 0, ctx->stream>>>(dev_view(ctx), a);
        ctx->launches++;
    }
    // the chosen class is the image's reference from here on (projector slot of the E-step, accumulator of the M-step).  With ONE
    // class (k = 1: refinement, the slots are the two half sets) there is nothing to choose and the slots stay what they are.
    if (nK > 1) {
        pf_set_slots_kernel<<<(nPar + 255) / 256, 256, 0, ctx->stream>>>(dcls, nPar, s.imgBase, ctx->stackE.slot,
                                                                        (ctx->stackM.slot && s.imgBase + nPar <= ctx->stackM.nImg) ? ctx->stackM.slot : nullptr);
        ctx->launches++;
    }
    span_end(ctx);
    THB_CUDA(ctx, cudaGetLastError());
    std::vector<int> hc(nPar);
    THB_CUDA(ctx, cudaMemcpyAsync(hc.data(), dcls, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < nPar; ++q) {
        if (nK > 1) {
            ctx->stackE.hslot[(size_t)s.imgBase + q] = hc[q];
            if ((size_t)s.imgBase + q < ctx->stackM.hslot.size()) ctx->stackM.hslot[(size_t)s.imgBase + q] = hc[q];
        }
        if (clsOut) clsOut[q] = hc[q];
    }
    return THB_OK;
}

int thb_pf_set_image_base(thb_ctx* ctx, int imgBase, uint64_t streamBase)
{
    if (!ctx) return THB_E_ARG;
    ctx->pf_.imgBase = imgBase;
    ctx->pf_.streamBase = streamBase;
    return THB_OK;
}

int thb_pf_get(thb_ctx* ctx, double* r, double* t, double* wR, double* wT, double* scal)
{
    if (!ctx) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.r) return set_error(ctx, THB_E_STATE, "pf_get: no particles loaded");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t n = s.nPar;
    std::vector<double> tmp;
    auto fetch = [&](const double* dsrc, double* dst, int S, int Cn) -> int {
        if (!dst) return THB_OK;
        tmp.resize(n * S * Cn);
        THB_CUDA(ctx, cudaMemcpy(tmp.data(), dsrc, sizeof(double) * n * S * Cn, cudaMemcpyDeviceToHost));
        to_aos(tmp.data(), dst, n, S, Cn);
        return THB_OK;
    };
    int rc;
    if ((rc = fetch(s.r, r, s.prm.mLR, 4))) return rc;
    if ((rc = fetch(s.t, t, s.prm.mLT, 2))) return rc;
    if ((rc = fetch(s.wR, wR, s.prm.mLR, 1))) return rc;
    if ((rc = fetch(s.wT, wT, s.prm.mLT, 1))) return rc;
    if ((rc = fetch(s.scal, scal, pf::S_COUNT, 1))) return rc;
    return THB_OK;
}

int thb_pf_set(thb_ctx* ctx, const double* r, const double* t, const double* wR, const double* wT, const double* scal)
{
    if (!ctx) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.r) return set_error(ctx, THB_E_STATE, "pf_set: no particles loaded");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t n = s.nPar;
    std::vector<double> tmp;
    auto put = [&](const double* src, double* ddst, int S, int Cn) -> int {
        if (!src) return THB_OK;
        tmp.resize(n * S * Cn);
        to_soa(src, tmp.data(), n, S, Cn);
        THB_CUDA(ctx, cudaMemcpy(ddst, tmp.data(), sizeof(double) * n * S * Cn, cudaMemcpyHostToDevice));
        return THB_OK;
    };
    int rc;
    if ((rc = put(r, s.r, s.prm.mLR, 4))) return rc;
    if ((rc = put(t, s.t, s.prm.mLT, 2))) return rc;
    if ((rc = put(wR, s.wR, s.prm.mLR, 1))) return rc;
    if ((rc = put(wT, s.wT, s.prm.mLT, 1))) return rc;
    if ((rc = put(scal, s.scal, pf::S_COUNT, 1))) return rc;
    return THB_OK;
}

static int expect_args_from_pf(thb_ctx* ctx, ExpectArgs& a)
{
    PFState& s = ctx->pf_;
    const int vdim = check_expect_state(ctx, "expectation");    // same volume edge in every slot, >= pf * N
    if (vdim < 0) return vdim;
    if (s.imgBase < 0 || s.imgBase + s.nPar > ctx->stackE.nImg)
        return set_error(ctx, THB_E_STATE, "expectation: particles [%d,%d) exceed the E stack (%d images)", s.imgBase,
                         s.imgBase + s.nPar, ctx->stackE.nImg);
    memset(&a, 0, sizeof(a));
    a.vols = vol_table(ctx); a.vdim = vdim; a.pitch = (vdim / 2 + 2 + 3) & ~3;
    a.dat = ctx->stackE.dat; a.ctf = ctx->stackE.ctf; a.sig = ctx->stackE.sig; a.slotOfImg = ctx->stackE.slot;
    a.pix = ctx->pixE; a.P = ctx->nPxlE; a.N = ctx->N;
    a.nAct = s.nPar; a.imgIdx = nullptr; a.imgBase = s.imgBase; a.active = s.active;
    a.slotAll = -1;                  // every image against the reference of ITS slot (MODE_2D: the class thb_pf_from_scan chose)
    a.nR = s.prm.mLR; a.nT = s.prm.mLT;
    const long long n = s.nPar;
    a.quat = View3{s.r, 1, n, n * s.prm.mLR};
    a.tran = View3{s.t, 1, n, n * s.prm.mLT};
    a.wR = View3{s.wR, 1, n, 0};
    a.wT = View3{s.wT, 1, n, 0};
    a.uR = s.uR; a.uT = s.uT; a.uC = s.uC; a.base = s.base; a.logL = nullptr;
    if (s.prm.mLD > 0) {
        if (!s.ctfSet) return set_error(ctx, THB_E_STATE, "expectation: CTF search needs thb_pf_set_ctf");
        if (!ctx->freqE || !ctx->stackE.def) return set_error(ctx, THB_E_STATE, "expectation: CTF search needs thb_set_frequency and thb_upload_stack_defocus");
        a.nD = s.prm.mLD; a.defP = ctx->stackE.def; a.freq = ctx->freqE; a.ctfK = s.ctfK;
        a.dpar = View3{s.d, 1, n, 0};
        a.wD = View3{s.wD, 1, n, 0};
        a.uD = s.uD;
    }
    return THB_OK;
}

int thb_expectation(thb_ctx* ctx, int* nPhaseOut)
{
    if (!ctx) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.r) return set_error(ctx, THB_E_STATE, "expectation: no particles loaded (thb_pf_load)");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    ExpectArgs ea;
    int rc = expect_args_from_pf(ctx, ea);
    if (rc) return rc;
    const thb_pf_params& p = s.prm;
    PFDev d = dev_view(ctx);
    const int nb = nblk(s.nPar);
    s.epoch += 1;
    StepArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.transS = p.transS; sa.transQ = p.transQ;
    sa.minPhase = p.minPhase; sa.maxPhase = p.maxPhase; sa.fixedPhases = p.fixedPhases;
    sa.noDecreaseLimit = p.noDecreaseLimit; sa.decreaseFactor = p.decreaseFactor;
    sa.ctfRefineS = p.ctfRefineS; sa.prePfD = p.perturbFactorSCTF;

    s.traceN = 0;
    if (s.traceWant > 0 && s.traceCap != s.traceWant) {
        cudaFree(s.traceR); cudaFree(s.traceT); cudaFree(s.traceSt); cudaFree(s.traceB);
        s.traceR = s.traceT = s.traceB = nullptr;
        s.traceSt = nullptr;
        s.traceCap = 0;
        THB_CUDA(ctx, cudaMalloc(&s.traceR, sizeof(float) * (size_t)s.traceWant * s.nPar * p.mLR));
        THB_CUDA(ctx, cudaMalloc(&s.traceT, sizeof(float) * (size_t)s.traceWant * s.nPar * p.mLT));
        THB_CUDA(ctx, cudaMalloc(&s.traceB, sizeof(float) * (size_t)s.traceWant * s.nPar));
        THB_CUDA(ctx, cudaMalloc(&s.traceSt, sizeof(double) * (size_t)s.traceWant * 2 * s.nPar * (4 * p.mLR + 2 * p.mLT)));
        s.traceCap = s.traceWant;
    }
    const size_t stSz = (size_t)s.nPar * (4 * p.mLR + 2 * p.mLT);
    const bool tracing = s.traceWant > 0 && s.traceSt;
    if (tracing) THB_CUDA(ctx, cudaMemsetAsync(s.traceSt, 0, sizeof(double) * (size_t)s.traceCap * 2 * stSz, ctx->stream));
    span_begin(ctx, KF_PF);
    pf_begin_kernel<<<nb, PF_BLOCK, 0, ctx->stream>>>(d);
    sa.doPost = 0; sa.doPre = 1; sa.phase = -1; sa.prePf = p.perturbFactorL; sa.epoch = (s.epoch << 20);
    THB_CUDA(ctx, cudaMemsetAsync(d.activeCount, 0, sizeof(int), ctx->stream));
    sa.tracePert = tracing ? s.traceSt : nullptr;
    // the particle's state staged in shared memory (one block per warp) unless the support is too large for it
    const size_t stageBytes = sizeof(double) * (size_t)pf_stage_doubles(p.mLR, p.mLT, p.mLD) * (PF_BLOCK / 32);
    const int staged = (stageBytes <= 200 * 1024 && ctx->pfStage) ? 1 : 0;
    if (staged) THB_CUDA(ctx, cudaFuncSetAttribute(pf_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stageBytes));
    pf_step_kernel<<<nb, PF_BLOCK, staged ? stageBytes : 0, ctx->stream>>>(d, sa, staged);
    span_end(ctx);
    ctx->launches += 2;
    const int phaseMax = p.fixedPhases > 0 ? p.fixedPhases : p.maxPhase;
    for (int phase = 0; phase < phaseMax; ++phase) {
        rc = launch_expect_local(ctx, ea);
        if (rc) return rc;
        if (s.traceWant > 0 && phase < s.traceCap) {     // test option: keep the weights of this phase
            THB_CUDA(ctx, cudaMemcpyAsync(s.traceR + (size_t)phase * s.nPar * p.mLR, s.uR, sizeof(float) * (size_t)s.nPar * p.mLR, cudaMemcpyDeviceToDevice, ctx->stream));
            THB_CUDA(ctx, cudaMemcpyAsync(s.traceT + (size_t)phase * s.nPar * p.mLT, s.uT, sizeof(float) * (size_t)s.nPar * p.mLT, cudaMemcpyDeviceToDevice, ctx->stream));
            THB_CUDA(ctx, cudaMemcpyAsync(s.traceB + (size_t)phase * s.nPar, s.base, sizeof(float) * (size_t)s.nPar, cudaMemcpyDeviceToDevice, ctx->stream));
            s.traceN = phase + 1;
        }
        sa.doPost = 1; sa.doPre = (phase + 1 < phaseMax); sa.phase = phase; sa.prePf = p.perturbFactorS;
        sa.traceRes = (tracing && phase < s.traceCap) ? s.traceSt + ((size_t)phase * 2 + 1) * stSz : nullptr;
        sa.tracePert = (tracing && phase + 1 < s.traceCap) ? s.traceSt + ((size_t)(phase + 1) * 2) * stSz : nullptr;
        sa.epoch = (s.epoch << 20) + (uint64_t)(phase + 1);
        THB_CUDA(ctx, cudaMemsetAsync(d.activeCount, 0, sizeof(int), ctx->stream));
        span_begin(ctx, KF_PF);
        pf_step_kernel<<<nb, PF_BLOCK, staged ? stageBytes : 0, ctx->stream>>>(d, sa, staged);
        span_end(ctx);
        ctx->launches++;
        THB_CUDA(ctx, cudaGetLastError());
        if (p.fixedPhases <= 0) {
            // adaptive E-step (MIN / MAX_N_PHASE_PER_ITER and the variance rule, src/Optimiser.cpp:1490-1560): the next launch takes the
            // compacted list of unfinished particles - a tail of a handful of particles then runs spread over the chip
            // (expect_spread_kernel) instead of one image per SM
            int act = 0;
            const bool compact = ctx->pfCompact && (ctx->expectImpl == 3 || ctx->expectImpl == 7);
            if (compact) pf_compact_kernel<<<1, 1024, 0, ctx->stream>>>(s.active, s.nPar, s.order, d.activeCount);
            THB_CUDA(ctx, cudaMemcpyAsync(&act, d.activeCount, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (act == 0) break;
            if (compact && act < s.nPar) { ea.order = s.order; ea.nAct = act; }
        }
    }
    if (nPhaseOut) THB_CUDA(ctx, cudaMemcpyAsync(nPhaseOut, s.nPhase, sizeof(int) * s.nPar, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_reconstruct_insert(thb_ctx* ctx, int mReco, int parGra, const double* offS)
{
    if (!ctx) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.r) return set_error(ctx, THB_E_STATE, "reconstruct_insert: no particles loaded");
    if (mReco <= 0) return set_error(ctx, THB_E_ARG, "reconstruct_insert: mReco <= 0");
    if (!ctx->pixM || !ctx->stackM.dat) return set_error(ctx, THB_E_STATE, "reconstruct_insert: M pixels / stack missing");
    if (s.imgBase < 0 || s.imgBase + s.nPar > ctx->stackM.nImg) return set_error(ctx, THB_E_STATE, "reconstruct_insert: particles exceed the M stack");
    {
        int rc = check_insert_slots(ctx, s.nPar, nullptr, s.imgBase, "reconstruct_insert");
        if (rc) return rc;
    }
    int vdim = 0;
    for (int i = 0; i < THB_MAX_SLOTS; ++i)
        if (ctx->accs[i].d) vdim = ctx->accs[i].vdim;
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = s.nPar;
    if (s.drawCap < mReco) {
        cudaFree(s.drawR); cudaFree(s.drawT); cudaFree(s.drawD);
        s.drawR = s.drawT = s.drawD = nullptr;
        THB_CUDA(ctx, cudaMalloc(&s.drawR, sizeof(int) * n * mReco));
        THB_CUDA(ctx, cudaMalloc(&s.drawT, sizeof(int) * n * mReco));
        THB_CUDA(ctx, cudaMalloc(&s.drawD, sizeof(int) * n * mReco));
        s.drawCap = mReco;
    }
    float* dw = (float*)scratch(ctx, 0, sizeof(float) * n + sizeof(double) * 2 * n + 64);
    if (!dw) return THB_E_CUDA;
    double* doff = (double*)(dw + ((n + 3) / 4) * 4);
    if (offS) THB_CUDA(ctx, cudaMemcpyAsync(doff, offS, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    s.epoch += 1;
    span_begin(ctx, KF_PF);
    pf_draw_kernel<<<nblk(s.nPar), PF_BLOCK, 0, ctx->stream>>>(dev_view(ctx), s.epoch << 20, mReco, parGra, s.drawR, s.drawT, s.drawD, dw);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());

    InsertArgs a;
    memset(&a, 0, sizeof(a));
    a.acc = acc_table(ctx); a.vdim = vdim;
    a.dat = ctx->stackM.dat; a.ctf = ctx->stackM.ctf; a.slotOfImg = ctx->stackM.slot;
    a.pix = ctx->pixM; a.P = ctx->nPxlM; a.N = ctx->NM;
    a.nImg = s.nPar; a.imgIdx = nullptr; a.imgBase = s.imgBase;
    a.mReco = mReco; a.w = dw; a.offS = offS ? doff : nullptr;
    const long long nn = s.nPar;
    a.nr = View3{s.r, 1, nn, nn * s.prm.mLR};
    a.nt = View3{s.t, 1, nn, nn * s.prm.mLT};
    a.drawR = s.drawR; a.drawT = s.drawT;
    a.mode2D = ctx->mode2D;
    if (s.prm.mLD > 0) {     // cSearch: the CTF of every draw from its own defocus factor (src/Optimiser.cpp:7171-7215)
        if (!s.ctfSet) return set_error(ctx, THB_E_STATE, "reconstruct_insert: CTF search needs thb_pf_set_ctf");
        a.nd = View3{s.d, 1, nn, 0};
        a.drawD = s.drawD;
        a.ctfAttr = s.ctfAttr;
        a.pixelSize = s.pixelSize;
    }
    int rc = launch_insert(ctx, a, nullptr);
    if (rc) return rc;
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_pf_set_ctf(thb_ctx* ctx, const float* ctfK, const float* ctfAttr, float pixelSize)
{
    if (!ctx) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.r || s.prm.mLD <= 0) return set_error(ctx, THB_E_STATE, "pf_set_ctf: load particles with mLD > 0 first");
    if (!ctfK || !ctfAttr || !(pixelSize > 0)) return set_error(ctx, THB_E_ARG, "pf_set_ctf: NULL arrays / pixelSize <= 0");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaMemcpyAsync(s.ctfK, ctfK, sizeof(float) * 4 * (size_t)s.nPar, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(s.ctfAttr, ctfAttr, sizeof(float) * 7 * (size_t)s.nPar, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    s.pixelSize = pixelSize;
    s.ctfSet = true;
    return THB_OK;
}

int thb_pf_get_d(thb_ctx* ctx, double* d, double* wD, double* sD)
{
    if (!ctx) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.d) return set_error(ctx, THB_E_STATE, "pf_get_d: no defocus dimension (mLD == 0)");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t n = s.nPar;
    const int D = s.prm.mLD;
    std::vector<double> tmp(n * (D + 1));
    if (d) {
        THB_CUDA(ctx, cudaMemcpy(tmp.data(), s.d, sizeof(double) * n * (D + 1), cudaMemcpyDeviceToHost));
        to_aos(tmp.data(), d, n, D + 1, 1);
    }
    if (wD) {
        THB_CUDA(ctx, cudaMemcpy(tmp.data(), s.wD, sizeof(double) * n * D, cudaMemcpyDeviceToHost));
        to_aos(tmp.data(), wD, n, D, 1);
    }
    if (sD) THB_CUDA(ctx, cudaMemcpy(sD, s.scal + (size_t)pf::S_SD * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return THB_OK;
}

int thb_pf_set_epoch(thb_ctx* ctx, uint64_t epoch)
{
    if (!ctx) return THB_E_ARG;
    ctx->pf_.epoch = epoch;
    return THB_OK;
}

int thb_pf_trace(thb_ctx* ctx, int nPhases)
{
    if (!ctx || nPhases < 0) return THB_E_ARG;
    ctx->pf_.traceWant = nPhases;
    return THB_OK;
}

int thb_pf_get_trace(thb_ctx* ctx, int nPhases, float* uR, float* uT, float* base)
{
    if (!ctx) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.traceR || nPhases <= 0 || nPhases > s.traceN) return set_error(ctx, THB_E_STATE, "pf_get_trace: %d phases asked, %d traced", nPhases, s.traceN);
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (uR) THB_CUDA(ctx, cudaMemcpy(uR, s.traceR, sizeof(float) * (size_t)nPhases * s.nPar * s.prm.mLR, cudaMemcpyDeviceToHost));
    if (uT) THB_CUDA(ctx, cudaMemcpy(uT, s.traceT, sizeof(float) * (size_t)nPhases * s.nPar * s.prm.mLT, cudaMemcpyDeviceToHost));
    if (base) THB_CUDA(ctx, cudaMemcpy(base, s.traceB, sizeof(float) * (size_t)nPhases * s.nPar, cudaMemcpyDeviceToHost));
    return THB_OK;
}

int thb_pf_get_trace_states(thb_ctx* ctx, int nPhases, double* st)
{
    if (!ctx || !st) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.traceSt || nPhases <= 0 || nPhases > s.traceN) return set_error(ctx, THB_E_STATE, "pf_get_trace_states: %d phases asked, %d traced", nPhases, s.traceN);
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    THB_CUDA(ctx, cudaMemcpy(st, s.traceSt, sizeof(double) * (size_t)nPhases * 2 * s.nPar * (4 * s.prm.mLR + 2 * s.prm.mLT), cudaMemcpyDeviceToHost));
    return THB_OK;
}

int thb_pf_get_draws(thb_ctx* ctx, int mReco, int* drawR, int* drawT)
{
    if (!ctx) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.drawR || mReco > s.drawCap) return set_error(ctx, THB_E_STATE, "pf_get_draws: no draws of that size");
    THB_CUDA(ctx, cudaMemcpy(drawR, s.drawR, sizeof(int) * (size_t)s.nPar * mReco, cudaMemcpyDeviceToHost));
    THB_CUDA(ctx, cudaMemcpy(drawT, s.drawT, sizeof(int) * (size_t)s.nPar * mReco, cudaMemcpyDeviceToHost));
    return THB_OK;
}

int thb_pf_get_draws_d(thb_ctx* ctx, int mReco, int* drawD)
{
    if (!ctx || !drawD) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.drawD || mReco > s.drawCap) return set_error(ctx, THB_E_STATE, "pf_get_draws_d: no draws of that size");
    THB_CUDA(ctx, cudaMemcpy(drawD, s.drawD, sizeof(int) * (size_t)s.nPar * mReco, cudaMemcpyDeviceToHost));
    return THB_OK;
}

int thb_pf_op(thb_ctx* ctx, int op, double arg, const float* uR, const float* uT)
{
    if (!ctx) return THB_E_ARG;
    PFState& s = ctx->pf_;
    if (!s.r) return set_error(ctx, THB_E_STATE, "pf_op: no particles loaded");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = s.nPar;
    if (op == THB_PF_SET_U_KEEP_PEAK) {
        if (!uR || !uT) return set_error(ctx, THB_E_ARG, "pf_op: uR/uT required");
        THB_CUDA(ctx, cudaMemcpyAsync(s.uR, uR, sizeof(float) * n * s.prm.mLR, cudaMemcpyHostToDevice, ctx->stream));
        THB_CUDA(ctx, cudaMemcpyAsync(s.uT, uT, sizeof(float) * n * s.prm.mLT, cudaMemcpyHostToDevice, ctx->stream));
    }
    s.epoch += 1;
    span_begin(ctx, KF_PF);
    pf_op_kernel<<<nblk(s.nPar), PF_BLOCK, 0, ctx->stream>>>(dev_view(ctx), op, arg, s.prm.transS, s.prm.transQ, s.epoch << 20);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

}  // extern "C"
