// thb_api.cu - C ABI entry points: context, geometry, resident data, E / M launches.
// See include/thunder_b200.h for the reference interface each entry point replaces.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <vector>
#include "thb_context.h"
#include "thb_kernels.cuh"
#include "thb_expect2.cuh"
#include "thb_expect3.cuh"
#include "thb_expect4.cuh"
#include "thb_expect5.cuh"
#include "thb_insert2.cuh"
#include "thb_expect6.cuh"
#include "thb_expect7.cuh"
#include "thb_expect8.cuh"
#include <cstdlib>

static thread_local std::string g_create_error;

namespace thb {

int set_error(thb_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

int cuda_fail(thb_ctx* ctx, cudaError_t e, const char* what)
{
    return set_error(ctx, THB_E_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

void* scratch(thb_ctx* ctx, int which, size_t bytes)
{
    if (ctx->scratchCap[which] >= bytes && ctx->scratch[which]) return ctx->scratch[which];
    if (ctx->scratch[which]) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->scratch[which]);
        ctx->scratch[which] = nullptr;
        ctx->scratchCap[which] = 0;
    }
    size_t cap = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&ctx->scratch[which], cap);
    if (e != cudaSuccess) {
        cuda_fail(ctx, e, "cudaMalloc(scratch)");
        return nullptr;
    }
    ctx->scratchCap[which] = cap;
    return ctx->scratch[which];
}

void span_begin(thb_ctx* ctx, int family)
{
    if (!ctx->timing) return;
    TimedSpan s;
    cudaEventCreate(&s.a);
    cudaEventCreate(&s.b);
    s.family = family;
    cudaEventRecord(s.a, ctx->stream);
    ctx->spans.push_back(s);
}

void span_end(thb_ctx* ctx)
{
    if (!ctx->timing || ctx->spans.empty()) return;
    cudaEventRecord(ctx->spans.back().b, ctx->stream);
}

void resolve_spans(thb_ctx* ctx)
{
    if (ctx->spans.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            ctx->famMs[s.family] += ms;
            ctx->famN[s.family] += 1;
        }
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    ctx->spans.clear();
}

VolTable vol_table(const thb_ctx* ctx)
{
    VolTable t;
    for (int i = 0; i < THB_MAX_SLOTS; ++i) t.p[i] = ctx->vols[i].d;
    return t;
}

VolTable quad_table(const thb_ctx* ctx)
{
    VolTable t;
    for (int i = 0; i < THB_MAX_SLOTS; ++i) t.p[i] = reinterpret_cast<const float2*>(ctx->vols[i].quad);
    return t;
}

// quad layout of slot `slot` (built lazily: only the quad kernel needs the 4x copy)
static int ensure_quad(thb_ctx* ctx, int slot)
{
    Volume3& v = ctx->vols[slot];
    if (!v.d) return THB_OK;
    if (ctx->mode2D) {
        // class average: the bilinear cell of every (x, y) in plain rows, one 256-bit load per sample
        if (v.quad) return THB_OK;
        const size_t elems2 = (size_t)v.vdim * (v.vdim / 2);
        THB_CUDA(ctx, cudaMalloc(&v.quad, elems2 * sizeof(Quad)));
        span_begin(ctx, KF_PACK);
        build_quad_kernel<<<ctx->smCount * 2, 256, 0, ctx->stream>>>(v.d, v.vdim, 1, v.pitch, 0, reinterpret_cast<Quad*>(v.quad));
        span_end(ctx);
        ctx->launches++;
        v.quadBrick = 0;
        v.quadOct = 0;
        THB_CUDA(ctx, cudaGetLastError());
        return THB_OK;
    }
    if (v.quad && v.quadBrick == ctx->quadBrick && v.quadOct == ctx->quadOct) return THB_OK;
    if ((v.vdim / 2) % (1 << ctx->quadBrick)) ctx->quadBrick = 0;   // tiny volumes: plain rows
    if (v.quad) {
        THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(v.quad);
        v.quad = nullptr;
    }
    const size_t elems = (size_t)v.vdim * v.vdim * (v.vdim / 2);
    if (ctx->quadOct && cudaMalloc(&v.quad, elems * sizeof(Quad) * 2) != cudaSuccess) {
        cudaGetLastError();          // not enough HBM for the 64-byte layout: use the 32-byte one
        v.quad = nullptr;
        ctx->quadOct = 0;
    }
    if (!ctx->quadOct) THB_CUDA(ctx, cudaMalloc(&v.quad, elems * sizeof(Quad)));
    span_begin(ctx, KF_PACK);
    if (ctx->quadOct)
        build_oct_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(v.d, v.vdim, v.pitch, ctx->quadBrick, reinterpret_cast<Quad*>(v.quad));
    else
        build_quad_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(v.d, v.vdim, v.vdim, v.pitch, ctx->quadBrick, reinterpret_cast<Quad*>(v.quad));
    v.quadBrick = ctx->quadBrick;
    v.quadOct = ctx->quadOct;
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    return THB_OK;
}

static int launch_expect_v3(thb_ctx* ctx, ExpectArgs a)
{
    // the cell layout (64-byte oct or 32-byte quad) is ONE decision for all slots: when the oct copy of a later slot does not
    // fit, ensure_quad falls back to the quad layout and the slots built before it are rebuilt in a second sweep
    for (int sweep = 0; sweep < 2; ++sweep) {
        const int octBefore = ctx->quadOct;
        for (int i = 0; i < THB_MAX_SLOTS; ++i) {
            int rc = ensure_quad(ctx, i);
            if (rc) return rc;
        }
        if (ctx->quadOct == octBefore) break;
    }
    for (int i = 0; i < THB_MAX_SLOTS; ++i)
        if (ctx->vols[i].d && !ctx->mode2D && ctx->vols[i].quadOct != ctx->quadOct)
            return set_error(ctx, THB_E_STATE, "expect: volume slot %d is not in the layout the launch uses", i);
    a.quads = quad_table(ctx);
    if (ctx->expectImpl != 3 && ctx->expectImpl != 7 && a.order)
        return set_error(ctx, THB_E_STATE, "expect: a compacted particle list needs expect_impl 3 or 7");
    a.quadBrick = ctx->mode2D ? 0 : ctx->quadBrick;
    a.sortRot = ctx->mode2D ? 0 : ctx->sortRot;
    a.work = nullptr;
    if (a.nD > 0) {
        // CTF search: the defocus dimension inside the kernel (the CTF of every defocus factor on the fly), table in global scratch
        if (ctx->mode2D) return set_error(ctx, THB_E_STATE, "expect: CTF search is MODE_3D only");
        a.work = (float*)scratch(ctx, 7, sizeof(float) * (size_t)a.nAct * a.nR * a.nT * a.nD);
        if (!a.work) return THB_E_CUDA;
        span_begin(ctx, KF_EXPECT);
        if (ctx->quadOct)
            expect_direct_kernel<2, true, false, E_TC, true><<<a.nAct, E3_THREADS, E3_SMEM_BYTES, ctx->stream>>>(a);
        else
            expect_direct_kernel<2, false, false, E_TC, true><<<a.nAct, E3_THREADS, E3_SMEM_BYTES, ctx->stream>>>(a);
        span_end(ctx);
        ctx->launches++;
        THB_CUDA(ctx, cudaGetLastError());
        return THB_OK;
    }
    // a handful of images (the reference's one-image-at-a-time seam, the tail of an adaptive E-step): one CTA per image would
    // leave the chip idle - spread every image over (pixel chunk, rotation group) CTAs instead (thb_expect6.cuh)
    if ((ctx->expectImpl == 3 || ctx->expectImpl == 7) && (ctx->expectSpread == 1 || (ctx->expectSpread < 0 && a.nAct * 4 <= ctx->smCount))) {
        const int nRT = a.nR * a.nT;
        double* table = (double*)scratch(ctx, 13, sizeof(double) * (size_t)a.nAct * nRT);
        a.work = (float*)scratch(ctx, 7, sizeof(float) * (size_t)a.nAct * nRT);
        if (!table || !a.work) return THB_E_CUDA;
        THB_CUDA(ctx, cudaMemsetAsync(table, 0, sizeof(double) * (size_t)a.nAct * nRT, ctx->stream));
        const int nGroups = (a.nR + 31) / 32, tiles = (a.P + E3_TILE - 1) / E3_TILE;
        const int nChunk = std::max(1, std::min(tiles, (2 * ctx->smCount + nGroups * a.nAct - 1) / (nGroups * a.nAct)));
        const dim3 grid(nChunk, nGroups, a.nAct);
        span_begin(ctx, KF_EXPECT);
        if (ctx->mode2D)
            expect_spread_kernel<false, true><<<grid, E6_THREADS, E3_SMEM_BYTES, ctx->stream>>>(a, table, nChunk);
        else if (ctx->quadOct)
            expect_spread_kernel<true, false><<<grid, E6_THREADS, E3_SMEM_BYTES, ctx->stream>>>(a, table, nChunk);
        else
            expect_spread_kernel<false, false><<<grid, E6_THREADS, E3_SMEM_BYTES, ctx->stream>>>(a, table, nChunk);
        expect_table_epilogue_kernel<<<a.nAct, 256, 0, ctx->stream>>>(a, table, a.work);
        span_end(ctx);
        ctx->launches += 2;
        THB_CUDA(ctx, cudaGetLastError());
        return THB_OK;
    }
    // more translations than one pass of the local-search kernel carries (the scans: nT = 30 in demo_2D.json): the variant
    // with 15 per pass halves the number of passes over the gather
    const bool wideT = (ctx->expectImpl == 3 || ctx->expectImpl == 7) && a.nT > E_TC;
    const int tc = wideT ? E3_TC_SCAN : E_TC;
    const bool single = a.nR <= E3_ROTS && a.nT <= tc;
    if (!single) {
        a.work = (float*)scratch(ctx, 7, sizeof(float) * (size_t)a.nAct * a.nR * a.nT);
        if (!a.work) return THB_E_CUDA;
    }
    const size_t smem = (ctx->expectImpl == 5 ? E5_SMEM_BYTES : wideT ? E3_TILE * sizeof(PixelRecT<E3_TC_SCAN>) : E3_SMEM_BYTES) +
                        (single ? sizeof(float) * (size_t)a.nR * a.nT : 0);
    span_begin(ctx, KF_EXPECT);
    // (supports of <= 64 rotations - mLR = 25 of demo_3D.json - would leave half of its lanes idle: they take the one-rotation-per-lane
    // kernel below, whose warps split 1 x 8 / 2 x 4 between rotations and pixels)
    if (ctx->expectImpl == 7 && single && !wideT && !ctx->mode2D && a.vdim < 65536 && a.nR > 64) {
        // several rotations per lane (thb_expect7.cuh): the record broadcast is amortised over RPL samples
        const int rpl = ctx->expectRpl >= 4 ? 4 : 2;
        const size_t sm7 = rpl == 4 ? e7_smem_bytes<4>(a.nR, a.nT) : e7_smem_bytes<2>(a.nR, a.nT);
        void (*kern)(const ExpectArgs) = rpl == 4 ? (ctx->quadOct ? expect_multi_kernel<4, true> : expect_multi_kernel<4, false>)
                                                  : (ctx->quadOct ? expect_multi_kernel<2, true> : expect_multi_kernel<2, false>);
        THB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm7));
        int grid = a.nAct;
        a.lockCtr = nullptr; a.lockTiles = 0; a.lockWindow = 0;      // (a.order: null, or the caller's compacted list of active particles)
        if (ctx->expectLock) {
            // lockstep launch: a persistent grid of co-resident CTAs walks the images wave by wave with a barrier every few pixel
            // tiles; the images of one slot are adjacent in the launch order so that a wave reads ONE volume
            int occ = 0;
            THB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, E3_THREADS, sm7));
            grid = std::max(1, std::min(a.nAct, occ * ctx->smCount));
            // arrival counters: one per barrier of every wave
            const int lockTiles = std::max(1, ctx->expectLockTiles);
            const size_t nBar = (size_t)((a.nAct + grid - 1) / grid) * (((a.P + E3_TILE - 1) / E3_TILE + lockTiles - 1) / lockTiles);
            int* dOrder = (int*)scratch(ctx, 14, sizeof(int) * ((size_t)a.nAct + 4 + nBar));
            if (!dOrder) return THB_E_CUDA;
            unsigned int* dCtr = reinterpret_cast<unsigned int*>(dOrder + a.nAct + ((4 - (a.nAct & 3)) & 3));
            if (!a.order && !a.imgIdx && a.slotOfImg && (size_t)(a.imgBase + a.nAct) <= ctx->stackE.hslot.size()) {
                std::vector<int>& ord = ctx->expectOrderHost;
                ord.resize(a.nAct);
                for (int i = 0; i < a.nAct; ++i) ord[i] = i;
                const int* hs = ctx->stackE.hslot.data() + a.imgBase;
                std::stable_sort(ord.begin(), ord.end(), [&](int l, int r) { return hs[l] < hs[r]; });
                THB_CUDA(ctx, cudaMemcpyAsync(dOrder, ord.data(), sizeof(int) * (size_t)a.nAct, cudaMemcpyHostToDevice, ctx->stream));
                a.order = dOrder;
            }
            THB_CUDA(ctx, cudaMemsetAsync(dCtr, 0, sizeof(unsigned int) * nBar, ctx->stream));
            a.lockCtr = dCtr; a.lockTiles = lockTiles; a.lockWindow = ctx->expectLockWindow;
        }
        kern<<<grid, E3_THREADS, sm7, ctx->stream>>>(a);
    } else if (ctx->expectImpl == 5) {
        if (ctx->mode2D)
            expect_pix_kernel<false, true><<<a.nAct, E5_THREADS, smem, ctx->stream>>>(a);
        else if (ctx->quadOct)
            expect_pix_kernel<true, false><<<a.nAct, E5_THREADS, smem, ctx->stream>>>(a);
        else
            expect_pix_kernel<false, false><<<a.nAct, E5_THREADS, smem, ctx->stream>>>(a);
    } else if (ctx->expectImpl == 4) {
        if (ctx->mode2D)
            expect_pair_kernel<false, true><<<a.nAct, E4_THREADS, smem, ctx->stream>>>(a);
        else if (ctx->quadOct)
            expect_pair_kernel<true, false><<<a.nAct, E4_THREADS, smem, ctx->stream>>>(a);
        else
            expect_pair_kernel<false, false><<<a.nAct, E4_THREADS, smem, ctx->stream>>>(a);
    } else if (wideT) {
        if (ctx->mode2D)
            expect_direct_kernel<2, false, true, E3_TC_SCAN><<<a.nAct, E3_THREADS, smem, ctx->stream>>>(a);
        else if (ctx->quadOct)
            expect_direct_kernel<2, true, false, E3_TC_SCAN><<<a.nAct, E3_THREADS, smem, ctx->stream>>>(a);
        else
            expect_direct_kernel<2, false, false, E3_TC_SCAN><<<a.nAct, E3_THREADS, smem, ctx->stream>>>(a);
    } else if (ctx->mode2D)
        expect_direct_kernel<2, false, true><<<a.nAct, E3_THREADS, smem, ctx->stream>>>(a);
    else if (ctx->quadOct)
        expect_direct_kernel<2, true><<<a.nAct, E3_THREADS, smem, ctx->stream>>>(a);
    else if (ctx->expectMinBlocks >= 3)
        expect_direct_kernel<3, false><<<a.nAct, E3_THREADS, smem, ctx->stream>>>(a);
    else
        expect_direct_kernel<2, false><<<a.nAct, E3_THREADS, smem, ctx->stream>>>(a);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    return THB_OK;
}

AccTable acc_table(const thb_ctx* ctx)
{
    AccTable t;
    for (int i = 0; i < THB_MAX_SLOTS; ++i) t.p[i] = ctx->accs[i].d;
    t.O = ctx->dO;
    t.counter = ctx->dCounter;
    return t;
}

static int launch_expect_v2(thb_ctx* ctx, ExpectArgs a)
{
    a.tiles = ctx->tilesE;
    a.nTiles = ctx->nTilesE;
    a.work = nullptr;
    a.stats = ctx->statsOn ? ctx->dStats : nullptr;
    if (a.nR > E2_ROTS || a.nT > E_TC) {
        a.work = (float*)scratch(ctx, 7, sizeof(float) * (size_t)a.nAct * a.nR * a.nT);
        if (!a.work) return THB_E_CUDA;
    }
    // the opt-in is per device, and cheap: set it at every launch (a process may hold contexts on several GPUs)
    THB_CUDA(ctx, cudaFuncSetAttribute(expect_local_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E2_SMEM_BYTES));
    span_begin(ctx, KF_EXPECT);
    expect_local_tma_kernel<<<a.nAct, E2_THREADS, E2_SMEM_BYTES, ctx->stream>>>(a);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    return THB_OK;
}

// Global scan with shared templates (thb_expect8.cuh): every rotation of the shared set is projected ONCE per launch into a
// [pixel][rotation] table (in chunks of rotations within 1 GiB), every image is contracted against it, then the epilogue.
static int launch_expect_scan_templates(thb_ctx* ctx, ExpectArgs a)
{
    const int slot = a.scanSlot1 - 1;
    const Volume3& v = ctx->vols[slot];
    const int P = a.P;
    const size_t nRT = (size_t)a.nR * a.nT;
    float* table = (float*)scratch(ctx, 7, sizeof(float) * (size_t)a.nAct * nRT);
    if (!table) return THB_E_CUDA;
    const size_t budget = (size_t)1 << 30;
    const int fit = (int)std::max<size_t>(E8_WROT, budget / ((size_t)P * sizeof(float2)) / E8_WROT * E8_WROT);
    const int chunkR = std::min(a.nR, fit);
    const int nRpad = (chunkR + E8_WROT - 1) / E8_WROT * E8_WROT;
    const size_t tbytes = (size_t)P * nRpad * sizeof(float2);
    const bool fresh = ctx->scratchCap[15] < tbytes || !ctx->scratch[15];
    float2* tmpl = (float2*)scratch(ctx, 15, tbytes);
    if (!tmpl) return THB_E_CUDA;
    if (fresh) THB_CUDA(ctx, cudaMemsetAsync(tmpl, 0, ctx->scratchCap[15], ctx->stream));   // the padding columns are read, never used
    const bool tc15 = a.nT > E_TC;
    const size_t smem = tc15 ? e8_smem_bytes<E3_TC_SCAN>() : e8_smem_bytes<E_TC>();
    if (tc15)
        THB_CUDA(ctx, cudaFuncSetAttribute(scan_contract_kernel<E3_TC_SCAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        THB_CUDA(ctx, cudaFuncSetAttribute(scan_contract_kernel<E_TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    span_begin(ctx, KF_EXPECT);
    for (int r0 = 0; r0 < a.nR; r0 += chunkR) {
        const int nRc = std::min(chunkR, a.nR - r0);
        scan_project_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(v.d, v.vdim, v.pitch, ctx->pixE, P, a.quat, r0, nRc, nRpad, ctx->mode2D, tmpl);
        if (tc15)
            scan_contract_kernel<E3_TC_SCAN><<<a.nAct, E8_THREADS, smem, ctx->stream>>>(a, tmpl, r0, nRc, nRpad, table);
        else
            scan_contract_kernel<E_TC><<<a.nAct, E8_THREADS, smem, ctx->stream>>>(a, tmpl, r0, nRc, nRpad, table);
        ctx->launches += 2;
    }
    scan_epilogue_kernel<<<a.nAct, 256, 0, ctx->stream>>>(a, table);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    return THB_OK;
}

int launch_expect_local(thb_ctx* ctx, const ExpectArgs& a_in)
{
    ExpectArgs a = a_in;
    if (a.nAct <= 0) return THB_OK;
    a.mode2D = ctx->mode2D;
    if (!ctx->mode2D) a.slotAll = -1;
    if (a.scanSlot1 > 0 && ctx->scanTemplates && a.nD == 0 && a.quat.sP == 0 && ctx->expectImpl >= 3 &&
        ctx->vols[a.scanSlot1 - 1].d)
        return launch_expect_scan_templates(ctx, a);
    if (ctx->expectImpl >= 3) return launch_expect_v3(ctx, a);
    if (ctx->expectImpl == 2 && !ctx->mode2D) return launch_expect_v2(ctx, a);
    const size_t smem = sizeof(PixelE) * E_TILE + (size_t)a.nR * a.nT * sizeof(float);
    if (smem > 200 * 1024)
        return set_error(ctx, THB_E_ARG, "expect_local: nR*nT = %d too large for the local-search kernel", a.nR * a.nT);
    THB_CUDA(ctx, cudaFuncSetAttribute(expect_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    span_begin(ctx, KF_EXPECT);
    expect_local_kernel<<<a.nAct, E_THREADS, smem, ctx->stream>>>(a);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    return THB_OK;
}

// round-1 image-ordered kernel: MODE_2D, per-draw classes, shapes beyond the slab kernel's tables, A/B measurements
static int launch_insert_legacy(thb_ctx* ctx, const InsertArgs& a)
{
    // enough CTAs to fill the chip even for a handful of images
    int split = 1;
    const int tiles = (a.P + M_THREADS * M_KP - 1) / (M_THREADS * M_KP);
    while (a.nImg * split < 4 * ctx->smCount && split < tiles) split *= 2;
    split = std::min(split, std::max(tiles, 1));
    dim3 grid(a.nImg, split);
    span_begin(ctx, KF_INSERT);
    if (ctx->insertImpl == 2)
        insert_kernel<2><<<grid, M_THREADS, 0, ctx->stream>>>(a);
    else
        insert_kernel<0><<<grid, M_THREADS, 0, ctx->stream>>>(a);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    return THB_OK;
}

int launch_insert(thb_ctx* ctx, const InsertArgs& a, const int* hImgIdx)
{
    if (a.nImg <= 0) return THB_OK;
    const bool slab = !ctx->mode2D && !a.drawC && (ctx->insertImpl == 0 || ctx->insertImpl == 3) && a.mReco <= M2_MAXD &&
                      ctx->segM && ctx->nSegM <= M2_MAXSEG;
    if (!slab && a.nd.p) return set_error(ctx, THB_E_STATE, "insert: the CTF search (per-draw CTF) needs the slab kernel (MODE_3D, mReco <= %d)", M2_MAXD);
    if (!slab) return launch_insert_legacy(ctx, a);
    // ---- slab insert (thb_insert2.cuh): grid (image, slab), images of one slot adjacent, slab thickness from the L2 budget
    const int n = a.vdim;
    const size_t accBytes = (size_t)(n / 2 + 1) * n * n * sizeof(float4);
    const size_t budget = (size_t)std::max(ctx->insertSlabMB, 1) << 20;
    int nSlab = (int)((accBytes + budget - 1) / budget);
    int th = ctx->insertSlabPlanes > 0 ? ctx->insertSlabPlanes : (n + nSlab - 1) / nSlab;
    th = std::max(1, std::min(th, n));
    nSlab = (n + th - 1) / th;
    const int maxD = (a.mReco + 3) & ~3;
    const size_t stride = prep_stride(maxD);
    const int chunk = 16384;
    std::vector<int> order;
    for (int l0 = 0; l0 < a.nImg; l0 += chunk) {
        const int c = std::min(chunk, a.nImg - l0);
        // images of one slot adjacent in the grid, so that one (slot, slab) block of accumulator is live at a time
        order.resize(c);
        bool mixed = false;
        {
            const std::vector<int>& hs = ctx->stackM.hslot;
            auto slotOf = [&](int l) {
                const size_t img = hImgIdx ? (size_t)hImgIdx[l] : (size_t)a.imgBase + l;
                return img < hs.size() ? hs[img] : 0;
            };
            const int s0 = slotOf(l0);
            for (int l = 0; l < c; ++l) { order[l] = l; mixed |= slotOf(l0 + l) != s0; }
            if (mixed) std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return slotOf(l0 + x) < slotOf(l0 + y); });
        }
        unsigned char* prep = (unsigned char*)scratch(ctx, 11, stride * (size_t)c);
        if (!prep) return THB_E_CUDA;
        int* dOrder = nullptr;
        if (mixed) {
            dOrder = (int*)scratch(ctx, 12, sizeof(int) * (size_t)c);
            if (!dOrder) return THB_E_CUDA;
            THB_CUDA(ctx, cudaMemcpyAsync(dOrder, order.data(), sizeof(int) * (size_t)c, cudaMemcpyHostToDevice, ctx->stream));
        }
        InsertSlabArgs s;
        memset(&s, 0, sizeof(s));
        s.a = a;
        s.a.nImg = c;
        // the chunk's rows of the per-image arrays
        if (a.imgIdx) s.a.imgIdx = a.imgIdx + l0; else s.a.imgBase = a.imgBase + l0;
        if (a.w) s.a.w = a.w + l0;
        if (a.offS) s.a.offS = a.offS + 2 * (size_t)l0;
        s.a.nr.p = a.nr.p + (size_t)l0 * a.nr.sP;
        s.a.nt.p = a.nt.p + (size_t)l0 * a.nt.sP;
        if (a.drawR) s.a.drawR = a.drawR + (size_t)l0 * a.mReco;
        if (a.drawT) s.a.drawT = a.drawT + (size_t)l0 * a.mReco;
        if (a.drawCount) s.a.drawCount = a.drawCount + l0;
        if (a.nd.p) { s.a.nd.p = a.nd.p + (size_t)l0 * a.nd.sP; s.a.ctfAttr = a.ctfAttr + 7 * (size_t)l0; }
        if (a.drawD) s.a.drawD = a.drawD + (size_t)l0 * a.mReco;
        s.seg = (const Seg*)ctx->segM; s.nSeg = ctx->nSegM; s.order = dOrder; s.prep = prep; s.maxD = maxD;
        s.pf = ctx->pfM; s.rMaxPad = ctx->rMaxPadM; s.zMin = -(n / 2); s.th = th;
        span_begin(ctx, KF_INSERT);
        if (ctx->insertImpl == 3)
            insert_prep_kernel<2><<<c, 128, 0, ctx->stream>>>(s);     // no merging of equal rotations (A/B)
        else
            insert_prep_kernel<0><<<c, 128, 0, ctx->stream>>>(s);
        insert_slab_kernel<<<dim3(c, nSlab), M2_THREADS, 0, ctx->stream>>>(s);
        span_end(ctx);
        ctx->launches += 2;
        THB_CUDA(ctx, cudaGetLastError());
        if (mixed && l0 + chunk < a.nImg) THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `order` scratch is reused
    }
    return THB_OK;
}

}  // namespace thb

using namespace thb;

extern "C" {

int thb_version(void) { return 100; }

int thb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int thb_create(thb_ctx** out, int device)
{
    if (!out) return set_error(nullptr, THB_E_ARG, "thb_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(nullptr, THB_E_NOGPU,
                         "thb_create: no CUDA device visible (%s); libthunder_b200 has no CPU path",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= n) return set_error(nullptr, THB_E_ARG, "thb_create: device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return set_error(nullptr, THB_E_NOGPU, "thb_create: device %d is sm_%d%d; this library is built for sm_100a only",
                         device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    thb_ctx* ctx = new thb_ctx();
    ctx->device = device;
    ctx->smCount = prop.multiProcessorCount;
    if (const char* e = getenv("THB_EXPECT_IMPL")) ctx->expectImpl = std::max(1, std::min(7, atoi(e)));
    if (const char* e = getenv("THB_EXPECT_RPL")) ctx->expectRpl = atoi(e) >= 4 ? 4 : 2;
    if (const char* e = getenv("THB_PF_STAGE")) ctx->pfStage = atoi(e) != 0;
    if (const char* e = getenv("THB_PF_COMPACT")) ctx->pfCompact = atoi(e) != 0;
    if (const char* e = getenv("THB_SCAN_TEMPLATES")) ctx->scanTemplates = atoi(e) != 0;
    if (const char* e = getenv("THB_EXPECT_ORDER")) ctx->expectOrder = atoi(e) == 1;
    if (const char* e = getenv("THB_EXPECT_LOCK")) ctx->expectLock = atoi(e) != 0;
    if (const char* e = getenv("THB_EXPECT_LOCK_TILES")) ctx->expectLockTiles = std::max(1, atoi(e));
    if (const char* e = getenv("THB_EXPECT_LOCK_WINDOW")) ctx->expectLockWindow = std::max(0, atoi(e));
    if (const char* e = getenv("THB_QUAD_BRICK")) ctx->quadBrick = std::max(0, std::min(4, atoi(e)));
    if (const char* e = getenv("THB_QUAD_OCT")) ctx->quadOct = atoi(e) != 0;
    if (const char* e = getenv("THB_SORT_ROT")) ctx->sortRot = atoi(e) != 0;
    if (const char* e = getenv("THB_EXPECT_MINB")) ctx->expectMinBlocks = atoi(e) >= 3 ? 3 : 2;
    if (const char* e = getenv("THB_INSERT_IMPL")) ctx->insertImpl = atoi(e);
    if (const char* e = getenv("THB_INSERT_SLAB_MB")) ctx->insertSlabMB = std::max(1, atoi(e));
    if (const char* e = getenv("THB_TILE_W")) ctx->tileW = std::max(1, std::min(16, atoi(e)));
    if (const char* e = getenv("THB_TILE_H")) ctx->tileH = std::max(1, std::min(E2_TILE / ctx->tileW, atoi(e)));
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return cuda_fail(nullptr, e, "cudaStreamCreate");
    }
    if ((e = cudaMalloc(&ctx->dO, sizeof(double) * 3 * THB_MAX_SLOTS)) != cudaSuccess ||
        (e = cudaMalloc(&ctx->dCounter, sizeof(int) * THB_MAX_SLOTS)) != cudaSuccess) {
        delete ctx;
        return cuda_fail(nullptr, e, "cudaMalloc");
    }
    cudaMemset(ctx->dO, 0, sizeof(double) * 3 * THB_MAX_SLOTS);
    cudaMemset(ctx->dCounter, 0, sizeof(int) * THB_MAX_SLOTS);
    *out = ctx;
    return THB_OK;
}

static void free_stack(Stack& s)
{
    cudaFree(s.dat); cudaFree(s.ctf); cudaFree(s.sig); cudaFree(s.def); cudaFree(s.slot);
    s = Stack();
}

void thb_destroy(thb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    resolve_spans(ctx);
    comm_destroy(ctx);
    pf_free(ctx);
    reco_free(ctx);
    for (int i = 0; i < THB_MAX_SLOTS; ++i) {
        cudaFree(ctx->vols[i].d);
        cudaFree(ctx->vols[i].quad);
        cudaFree(ctx->accs[i].d);
    }
    free_stack(ctx->stackE);
    free_stack(ctx->stackM);
    cudaFree(ctx->pixE); cudaFree(ctx->pixM); cudaFree(ctx->permE); cudaFree(ctx->permM); cudaFree(ctx->segM); cudaFree(ctx->freqE); cudaFree(ctx->tilesE); cudaFree(ctx->dStats);
    cudaFree(ctx->dO); cudaFree(ctx->dCounter); cudaFree(ctx->commBuf);
    for (int i = 0; i < THB_N_SCRATCH; ++i) cudaFree(ctx->scratch[i]);
    if (ctx->copyDone) cudaEventDestroy(ctx->copyDone);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* thb_last_error(const thb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int thb_synchronize(thb_ctx* ctx)
{
    if (!ctx) return THB_E_ARG;
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

// CUDA-event stopwatch on the library's launch stream (bench.py times its steps with it)
int thb_timer(thb_ctx* ctx, int stop, float* ms)
{
    if (!ctx) return THB_E_ARG;
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!stop) {
        if (!ctx->tA) { THB_CUDA(ctx, cudaEventCreate(&ctx->tA)); THB_CUDA(ctx, cudaEventCreate(&ctx->tB)); }
        THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        THB_CUDA(ctx, cudaEventRecord(ctx->tA, ctx->stream));
        return THB_OK;
    }
    if (!ctx->tA) return set_error(ctx, THB_E_STATE, "timer: stop without start");
    THB_CUDA(ctx, cudaEventRecord(ctx->tB, ctx->stream));
    THB_CUDA(ctx, cudaEventSynchronize(ctx->tB));
    float v = 0.f;
    THB_CUDA(ctx, cudaEventElapsedTime(&v, ctx->tA, ctx->tB));
    if (ms) *ms = v;
    return THB_OK;
}

int64_t thb_launch_count(thb_ctx* ctx, int reset)
{
    if (!ctx) return 0;
    int64_t v = ctx->launches;
    if (reset) ctx->launches = 0;
    return v;
}

int thb_enable_timing(thb_ctx* ctx, int on)
{
    if (!ctx) return THB_E_ARG;
    resolve_spans(ctx);
    ctx->timing = on != 0;
    return THB_OK;
}

int thb_set_option(thb_ctx* ctx, const char* key, int value)
{
    if (!ctx || !key) return THB_E_ARG;
    if (!strcmp(key, "expect_impl")) {
        if (value < 0 || value > 7 || value == 6) return set_error(ctx, THB_E_ARG, "set_option: expect_impl must be 0 (default), 1 .. 5 or 7");
        ctx->expectImpl = value ? value : THB_DEFAULT_EXPECT_IMPL;
        return THB_OK;
    }
    if (!strcmp(key, "pf_compact")) {     // adaptive E-step: launch the E kernel on the compacted list of unfinished particles
        ctx->pfCompact = value != 0;
        return THB_OK;
    }
    if (!strcmp(key, "pf_stage")) {
        ctx->pfStage = value != 0;
        return THB_OK;
    }
    if (!strcmp(key, "scan_templates")) {  // scans: project each shared rotation once per launch (1, default) or per image (0)
        ctx->scanTemplates = value != 0;
        return THB_OK;
    }
    if (!strcmp(key, "expect_order")) {    // pixel order of the E stack: 0 = 8x8 blocks, 1 = radial; at the next thb_set_expect_pixels
        if (value < 0 || value > 1) return set_error(ctx, THB_E_ARG, "set_option: expect_order must be 0 or 1");
        ctx->expectOrder = value;
        return THB_OK;
    }
    if (!strcmp(key, "expect_lock")) {     // lockstep launch of expect_impl 7 (persistent grid, tile barriers)
        ctx->expectLock = value != 0;
        return THB_OK;
    }
    if (!strcmp(key, "expect_lock_tiles")) {
        if (value < 1) return set_error(ctx, THB_E_ARG, "set_option: expect_lock_tiles must be >= 1");
        ctx->expectLockTiles = value;
        return THB_OK;
    }
    if (!strcmp(key, "expect_lock_window")) {
        if (value < 0) return set_error(ctx, THB_E_ARG, "set_option: expect_lock_window must be >= 0");
        ctx->expectLockWindow = value;
        return THB_OK;
    }
    if (!strcmp(key, "expect_rpl")) {      // rotations per lane of expect_impl 7
        if (value != 2 && value != 4) return set_error(ctx, THB_E_ARG, "set_option: expect_rpl must be 2 or 4");
        ctx->expectRpl = value;
        return THB_OK;
    }
    if (!strcmp(key, "tile_w") || !strcmp(key, "tile_h")) {   // takes effect at the next thb_set_expect_pixels
        if (value < 1 || value > 16) return set_error(ctx, THB_E_ARG, "set_option: tile_w / tile_h must be in [1,16]");
        (key[5] == 'w' ? ctx->tileW : ctx->tileH) = value;
        if (ctx->tileW * ctx->tileH > E2_TILE) return set_error(ctx, THB_E_ARG, "set_option: tile_w * tile_h must be <= %d", E2_TILE);
        return THB_OK;
    }
    if (!strcmp(key, "quad_brick")) {
        if (value < 0 || value > 4) return set_error(ctx, THB_E_ARG, "set_option: quad_brick must be in [0,4]");
        ctx->quadBrick = value;
        return THB_OK;
    }
    if (!strcmp(key, "quad_oct")) {
        ctx->quadOct = value != 0;
        return THB_OK;
    }
    if (!strcmp(key, "sort_rot")) {
        ctx->sortRot = value != 0;
        return THB_OK;
    }
    if (!strcmp(key, "expect_minb")) {
        ctx->expectMinBlocks = value >= 3 ? 3 : 2;
        return THB_OK;
    }
    if (!strcmp(key, "stats")) {
        if (value && !ctx->dStats) {
            THB_CUDA(ctx, cudaMalloc(&ctx->dStats, 16 * sizeof(unsigned long long)));
            THB_CUDA(ctx, cudaMemset(ctx->dStats, 0, 16 * sizeof(unsigned long long)));
        }
        ctx->statsOn = value != 0;
        return THB_OK;
    }
    if (!strcmp(key, "expect_spread")) {   // -1: automatic (few images), 0: never, 1: always
        ctx->expectSpread = value < 0 ? -1 : (value ? 1 : 0);
        return THB_OK;
    }
    if (!strcmp(key, "insert_slab_mb")) {
        if (value < 1) return set_error(ctx, THB_E_ARG, "set_option: insert_slab_mb must be >= 1");
        ctx->insertSlabMB = value;
        return THB_OK;
    }
    if (!strcmp(key, "insert_slab_planes")) {
        if (value < 0) return set_error(ctx, THB_E_ARG, "set_option: insert_slab_planes must be >= 0");
        ctx->insertSlabPlanes = value;
        return THB_OK;
    }
    if (!strcmp(key, "insert_impl")) {
        ctx->insertImpl = value;
        return THB_OK;
    }
    return set_error(ctx, THB_E_ARG, "set_option: unknown key %s", key);
}

int thb_expect_stats(thb_ctx* ctx, uint64_t out[16], int reset)
{
    if (!ctx || !out) return THB_E_ARG;
    if (!ctx->dStats) return set_error(ctx, THB_E_STATE, "expect_stats: enable with thb_set_option(ctx, \"stats\", 1)");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    THB_CUDA(ctx, cudaMemcpy(out, ctx->dStats, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) THB_CUDA(ctx, cudaMemset(ctx->dStats, 0, 16 * sizeof(uint64_t)));
    return THB_OK;
}

double thb_kernel_ms(thb_ctx* ctx, int which, int64_t* n, int reset)
{
    if (!ctx || which < 0 || which >= KF_COUNT) return 0.0;
    resolve_spans(ctx);
    double v = ctx->famMs[which];
    if (n) *n = ctx->famN[which];
    if (reset) {
        ctx->famMs[which] = 0;
        ctx->famN[which] = 0;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// a1: Optimiser::allocPreCalIdx (reference src/Optimiser.cpp:7991-8041), host integer math.
//   IMAGE_FOR_PIXEL_R_FT(rU + 1): j in [-(rU+1), rU+1), i in [0, rU+1]  (include/Image/Image.h:68-70)
//   QUAD(i,j) = i*i + j*j ; NORM = hypot ; AROUND = rint                 (include/Functions/Functions.h)
// ------------------------------------------------------------------------------------------------
int thb_pixel_list(int N, int pf, float rU, float rL, int* iCol, int* iRow, int* iPxl, int* iSig, int* iColPad,
                   int* iRowPad)
{
    if (N <= 0 || (N & 1) || pf <= 0) return THB_E_ARG;
    const float rU2 = rU * rU, rL2 = rL * rL;
    const float R = rU + 1;                // the macro argument (a float expression) is the loop bound
    const int nColFT = N / 2 + 1;
    int n = 0;
    for (long j = (long)(-R); (float)j < R; ++j)
        for (long i = 0; (float)i <= R; ++i) {
            if (i == 0 && j < 0) continue;
            const float u = (float)((double)(i * i) + (double)(j * j));
            if (u < rU2 && u >= rL2) {
                const int v = (int)rint(hypot((double)i, (double)j));
                if ((float)v < rU && (float)v >= rL) {
                    if (n >= nColFT * N) return THB_E_ARG;
                    if (iPxl) iPxl[n] = (int)((j >= 0 ? j : j + N) * nColFT + i);
                    if (iCol) iCol[n] = (int)(i);
                    if (iRow) iRow[n] = (int)(j);
                    if (iSig) iSig[n] = v;
                    if (iColPad) iColPad[n] = (int)(i * pf);
                    if (iRowPad) iRowPad[n] = (int)(j * pf);
                    ++n;
                }
            }
        }
    return n;
}

// Blocked pixel order.  The reference walks the half-plane row by row; consecutive rows of one
// orientation cloud then revisit the same volume neighbourhood only after a whole row of work, far
// beyond L1/L2 reach.  The device keeps every per-pixel array in an order that walks 8x8-pixel
// blocks boustrophedon (rows of blocks alternate direction), so the trilinear cells of consecutive
// pixels stay inside a compact 3D neighbourhood.  Sums over pixels are order-independent up to fp32
// rounding; per-pixel results handed back to the caller (thb_project) are un-permuted.
static void blocked_order(int n, const int* a, const int* b, int unit, std::vector<int>& perm, std::vector<long long>* blockOf = nullptr,
                          int BW = 8, int BH = 8)
{
    std::vector<long long> key(n);
    for (int i = 0; i < n; ++i) {
        const int x = a[i] / unit, y = b[i] / unit + (1 << 20);
        const int by = y / BH, bx = x / BW;
        const int sx = (by & 1) ? (1 << 16) - bx : bx;
        const int iy = y % BH, ix = (iy & 1) ? BW - 1 - x % BW : x % BW;
        key[i] = (((long long)by << 40) | ((long long)sx << 20) | (long long)(iy << 5 | ix));
    }
    perm.resize(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int l, int r) { return key[l] < key[r]; });
    if (blockOf) {
        blockOf->resize(n);
        for (int i = 0; i < n; ++i) (*blockOf)[i] = key[perm[i]] >> 20;   // (by, sx): one value per 8x8 block
    }
}

// Radial order (option "expect_order" = 1): ring by ring (rounded radius), along the ring by angle, alternating direction from
// ring to ring.  The samples of a pixel at radius rho fall on the sphere of radius pf * rho of the volume whatever the
// orientation, so CTAs that walk their images in this order IN STEP (thb_expect7.cuh, lockstep launch) read one thin spherical
// shell of the volume at a time - a working set the L2 holds.
static void radial_order(int n, const int* a, const int* b, int unit, std::vector<int>& perm, std::vector<long long>* blockOf)
{
    std::vector<long long> key(n);
    for (int i = 0; i < n; ++i) {
        const double x = a[i] / unit, y = b[i] / unit;
        const long long ring = llround(sqrt(x * x + y * y) * 2.0);           // half-pixel rings
        const double ang = atan2(y, x);                                     // [-pi, pi]
        long long aq = llround((ang + 3.2) * 1e6);
        if (ring & 1) aq = 7000000 - aq;
        key[i] = (ring << 32) | aq;
    }
    perm.resize(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int l, int r) { return key[l] < key[r]; });
    if (blockOf) {
        blockOf->resize(n);
        for (int i = 0; i < n; ++i) (*blockOf)[i] = i / 64;                  // tiles of the staged kernel: runs of 64 pixels
    }
}

// tiles of the blocked order: maximal runs of pixels of one 8x8 block, with the rectangle they span
static void build_tiles(int n, const int* a, const int* b, int unit, int pf, const std::vector<int>& perm,
                        const std::vector<long long>& blockOf, std::vector<TileDesc>& tiles)
{
    tiles.clear();
    int i = 0;
    while (i < n) {
        int j = i;
        int amin = 1 << 30, amax = -(1 << 30), bmin = 1 << 30, bmax = -(1 << 30);
        while (j < n && blockOf[j] == blockOf[i] && j - i < E2_TILE) {
            const int x = a[perm[j]] / unit * pf, y = b[perm[j]] / unit * pf;   // padded units
            amin = std::min(amin, x); amax = std::max(amax, x);
            bmin = std::min(bmin, y); bmax = std::max(bmax, y);
            ++j;
        }
        TileDesc t;
        t.start = i; t.count = j - i;
        t.ca = 0.5f * (float)(amin + amax); t.cb = 0.5f * (float)(bmin + bmax);
        t.ha = 0.5f * (float)(amax - amin); t.hb = 0.5f * (float)(bmax - bmin);
        t.pad0 = t.pad1 = 0;
        tiles.push_back(t);
        i = j;
    }
}

// Row-major order: rows j ascending, columns i ascending inside a row, cut into runs of consecutive columns
// {j, iFirst, count, startIdx} (for the reference's pixel lists with rL = 0: one run per row, and the order is the caller's own)
static void rowmajor_order(int n, const int* a, const int* b, int unit, std::vector<int>& perm, std::vector<Seg>& segs)
{
    perm.resize(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int l, int r) {
        const int yl = b[l] / unit, yr = b[r] / unit;
        return yl != yr ? yl < yr : a[l] / unit < a[r] / unit;
    });
    segs.clear();
    for (int k = 0; k < n; ++k) {
        const int x = a[perm[k]] / unit, y = b[perm[k]] / unit;
        if (!segs.empty() && segs.back().j == y && segs.back().iFirst + segs.back().count == x && segs.back().count < 32767)
            segs.back().count++;
        else
            segs.push_back(Seg{y, x, 1, k});
    }
}

static int upload_pixels(thb_ctx* ctx, int pf, int nPxl, const int* a, const int* b, int padded, int4** dst, int** dperm,
                         std::vector<TileDesc>* tiles = nullptr, int BW = 8, int BH = 8, std::vector<Seg>* segs = nullptr)
{
    if (nPxl <= 0 || !a || !b) return set_error(ctx, THB_E_ARG, "pixel list is empty or NULL");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int> perm;
    std::vector<long long> blockOf;
    if (segs)
        rowmajor_order(nPxl, a, b, padded ? pf : 1, perm, *segs);
    else if (tiles && ctx->expectOrder == 1)
        radial_order(nPxl, a, b, padded ? pf : 1, perm, &blockOf);
    else
        blocked_order(nPxl, a, b, padded ? pf : 1, perm, &blockOf, BW, BH);
    if (tiles) build_tiles(nPxl, a, b, padded ? pf : 1, pf, perm, blockOf, *tiles);
    int* tmp = (int*)scratch(ctx, 0, sizeof(int) * 2 * (size_t)nPxl);
    if (!tmp) return THB_E_CUDA;
    THB_CUDA(ctx, cudaMemcpyAsync(tmp, a, sizeof(int) * nPxl, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(tmp + nPxl, b, sizeof(int) * nPxl, cudaMemcpyHostToDevice, ctx->stream));
    if (*dst) { cudaFree(*dst); *dst = nullptr; }
    if (*dperm) { cudaFree(*dperm); *dperm = nullptr; }
    THB_CUDA(ctx, cudaMalloc(dst, sizeof(int4) * (size_t)nPxl));
    THB_CUDA(ctx, cudaMalloc(dperm, sizeof(int) * (size_t)nPxl));
    THB_CUDA(ctx, cudaMemcpyAsync(*dperm, perm.data(), sizeof(int) * nPxl, cudaMemcpyHostToDevice, ctx->stream));
    make_pix_kernel<<<(nPxl + 255) / 256, 256, 0, ctx->stream>>>(tmp, tmp + nPxl, *dperm, nPxl, pf, padded, *dst);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

static void free_stack(Stack& s);

int thb_set_expect_pixels(thb_ctx* ctx, int N, int pf, int nPxl, const int* iCol, const int* iRow)
{
    if (!ctx) return THB_E_ARG;
    if (N <= 0 || pf <= 0) return set_error(ctx, THB_E_ARG, "set_expect_pixels: bad N/pf");
    std::vector<TileDesc> tiles;
    int rc = upload_pixels(ctx, pf, nPxl, iCol, iRow, 0, &ctx->pixE, &ctx->permE, &tiles, ctx->tileW, ctx->tileH);
    if (rc) return rc;
    // the staging counters (option "stats") are independent of the pixel list: they are cleared, not released
    cudaFree(ctx->tilesE);
    ctx->tilesE = nullptr;
    if (ctx->dStats) THB_CUDA(ctx, cudaMemset(ctx->dStats, 0, 16 * sizeof(unsigned long long)));
    ctx->nTilesE = (int)tiles.size();
    THB_CUDA(ctx, cudaMalloc(&ctx->tilesE, sizeof(TileDesc) * tiles.size()));
    THB_CUDA(ctx, cudaMemcpy(ctx->tilesE, tiles.data(), sizeof(TileDesc) * tiles.size(), cudaMemcpyHostToDevice));
    // a resident stack belongs to one pixel list AND one pixel order (its arrays are stored in that order)
    if (ctx->nPxlE != nPxl || ctx->expectOrderBuilt != ctx->expectOrder) free_stack(ctx->stackE);
    ctx->expectOrderBuilt = ctx->expectOrder;
    cudaFree(ctx->freqE);                              // ... and so does the frequency table of the CTF search
    ctx->freqE = nullptr;
    ctx->N = N; ctx->pf = pf; ctx->nPxlE = nPxl;
    return THB_OK;
}

int thb_set_insert_pixels(thb_ctx* ctx, int N, int pf, int nPxl, const int* iColPad, const int* iRowPad)
{
    if (!ctx) return THB_E_ARG;
    if (N <= 0 || pf <= 0) return set_error(ctx, THB_E_ARG, "set_insert_pixels: bad N/pf");
    // the M pixel list is kept in row-major runs (the slab insert walks index intervals along rows, thb_insert2.cuh)
    std::vector<Seg> segs;
    int rc = upload_pixels(ctx, pf, nPxl, iColPad, iRowPad, 1, &ctx->pixM, &ctx->permM, nullptr, 8, 8, &segs);
    if (rc) return rc;
    cudaFree(ctx->segM);
    ctx->segM = nullptr;
    ctx->nSegM = (int)segs.size();
    THB_CUDA(ctx, cudaMalloc(&ctx->segM, sizeof(Seg) * segs.size()));
    THB_CUDA(ctx, cudaMemcpy(ctx->segM, segs.data(), sizeof(Seg) * segs.size(), cudaMemcpyHostToDevice));
    double rmax = 0.0;
    for (int i = 0; i < nPxl; ++i) rmax = std::max(rmax, hypot((double)iColPad[i], (double)iRowPad[i]));
    ctx->rMaxPadM = (float)rmax;
    if (ctx->nPxlM != nPxl) free_stack(ctx->stackM);
    ctx->NM = N; ctx->pfM = pf; ctx->nPxlM = nPxl;
    return THB_OK;
}

static size_t vol_elems(int vdim) { return (size_t)(vdim / 2 + 1) * vdim * vdim; }
static int vol_pitch(int vdim) { return (vdim / 2 + 2 + 3) & ~3; }

int thb_set_volume(thb_ctx* ctx, int slot, const float* volFT, int vdim)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || !volFT || vdim <= 0 || (vdim & 1))
        return set_error(ctx, THB_E_ARG, "set_volume: bad slot/vdim");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    Volume3& v = ctx->vols[slot];
    const int pitch = vol_pitch(vdim);
    // MODE_2D: the reference image is plane 0 of a two-plane volume whose plane 1 stays zero (thb_math.cuh, make_rot2)
    const size_t rows = ctx->mode2D ? (size_t)vdim : (size_t)vdim * vdim;
    const size_t rowsAlloc = ctx->mode2D ? 2 * (size_t)vdim : rows;
    if (v.quad) {   // the quad copy belongs to the previous contents
        THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(v.quad);
        v.quad = nullptr;
    }
    if (v.vdim != vdim) {
        cudaFree(v.d);
        v.d = nullptr;
        v.vdim = 0;
        THB_CUDA(ctx, cudaMalloc(&v.d, rowsAlloc * pitch * sizeof(float2)));
        THB_CUDA(ctx, cudaMemsetAsync(v.d, 0, rowsAlloc * pitch * sizeof(float2), ctx->stream));
        v.vdim = vdim;
        v.pitch = pitch;
    }
    const size_t rowBytes = (size_t)(vdim / 2 + 1) * sizeof(float2);
    THB_CUDA(ctx, cudaMemcpy2DAsync(v.d, (size_t)pitch * sizeof(float2), volFT, rowBytes, rowBytes, rows, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_get_volume(thb_ctx* ctx, int slot, float* volFT)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || !volFT || !ctx->vols[slot].d)
        return set_error(ctx, THB_E_STATE, "get_volume: slot %d has no volume", slot);
    const Volume3& v = ctx->vols[slot];
    const size_t rowBytes = (size_t)(v.vdim / 2 + 1) * sizeof(float2);
    THB_CUDA(ctx, cudaMemcpy2DAsync(volFT, rowBytes, v.d, (size_t)v.pitch * sizeof(float2), rowBytes,
                                    ctx->mode2D ? (size_t)v.vdim : (size_t)v.vdim * v.vdim, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_stack_reserve(thb_ctx* ctx, int kind, int capacity)
{
    if (!ctx) return THB_E_ARG;
    if (kind != THB_STACK_EXPECT && kind != THB_STACK_INSERT) return set_error(ctx, THB_E_ARG, "stack_reserve: bad kind");
    const int P = kind == THB_STACK_EXPECT ? ctx->nPxlE : ctx->nPxlM;
    if (P <= 0) return set_error(ctx, THB_E_STATE, "stack_reserve: pixel list for this stack kind not set");
    if (capacity <= 0) return set_error(ctx, THB_E_ARG, "stack_reserve: capacity <= 0");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    Stack& s = kind == THB_STACK_EXPECT ? ctx->stackE : ctx->stackM;
    if (s.nImg == capacity && s.dat) return THB_OK;
    free_stack(s);
    const size_t n = (size_t)capacity * P;
    THB_CUDA(ctx, cudaMalloc(&s.dat, n * sizeof(float2)));
    THB_CUDA(ctx, cudaMalloc(&s.ctf, n * sizeof(float)));
    if (kind == THB_STACK_EXPECT) THB_CUDA(ctx, cudaMalloc(&s.sig, n * sizeof(float)));
    THB_CUDA(ctx, cudaMalloc(&s.slot, (size_t)capacity * sizeof(int)));
    THB_CUDA(ctx, cudaMemsetAsync(s.slot, 0, (size_t)capacity * sizeof(int), ctx->stream));
    s.hslot.assign((size_t)capacity, 0);
    s.nImg = capacity;
    return THB_OK;
}

static int upload_stack_impl(thb_ctx* ctx, int kind, int base, int nImg, const float* dat, const float* ctf, const float* sigRcp,
                             const int* slotOfImg, bool async)
{
    if (!ctx) return THB_E_ARG;
    if (kind != THB_STACK_EXPECT && kind != THB_STACK_INSERT) return set_error(ctx, THB_E_ARG, "upload_stack: bad kind");
    const int P = kind == THB_STACK_EXPECT ? ctx->nPxlE : ctx->nPxlM;
    const int* perm = kind == THB_STACK_EXPECT ? ctx->permE : ctx->permM;
    if (P <= 0) return set_error(ctx, THB_E_STATE, "upload_stack: pixel list for this stack kind not set");
    if (nImg <= 0 || !dat || !ctf) return set_error(ctx, THB_E_ARG, "upload_stack: empty stack / NULL arrays");
    if (kind == THB_STACK_EXPECT && !sigRcp) return set_error(ctx, THB_E_ARG, "upload_stack: sigRcp required for the E stack");
    Stack& s = kind == THB_STACK_EXPECT ? ctx->stackE : ctx->stackM;
    if (!s.dat) return set_error(ctx, THB_E_STATE, "upload_stack: stack not reserved");
    if (base < 0 || base + nImg > s.nImg) return set_error(ctx, THB_E_ARG, "upload_stack: images [%d,%d) exceed the capacity %d", base, base + nImg, s.nImg);
    if (slotOfImg)
        for (int i = 0; i < nImg; ++i)
            if (slotOfImg[i] < 0 || slotOfImg[i] >= THB_MAX_SLOTS)
                return set_error(ctx, THB_E_ARG, "upload_stack: slotOfImg[%d] = %d out of range", i, slotOfImg[i]);
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    // stage chunks of images in HBM, then permute each chunk into the resident blocked layout
    cudaStream_t st = ctx->stream;
    int sb = 4;
    if (async) {
        if (!ctx->copyStream) {
            THB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
            THB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copyDone, cudaEventDisableTiming));
        }
        st = ctx->copyStream;
        sb = 8;
    }
    const size_t perImg = (size_t)P * 16;
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nImg, ((size_t)256 << 20) / perImg));
    float2* sdat = (float2*)scratch(ctx, sb, (size_t)chunk * P * sizeof(float2));
    float* sctf = (float*)scratch(ctx, sb + 1, (size_t)chunk * P * sizeof(float));
    float* ssig = kind == THB_STACK_EXPECT ? (float*)scratch(ctx, sb + 2, (size_t)chunk * P * sizeof(float)) : nullptr;
    if (!sdat || !sctf || (kind == THB_STACK_EXPECT && !ssig)) return THB_E_CUDA;
    for (int i0 = 0; i0 < nImg; i0 += chunk) {
        const int c = std::min(chunk, nImg - i0);
        const size_t off = (size_t)i0 * P, n = (size_t)c * P;
        THB_CUDA(ctx, cudaMemcpyAsync(sdat, dat + 2 * off, n * sizeof(float2), cudaMemcpyHostToDevice, st));
        THB_CUDA(ctx, cudaMemcpyAsync(sctf, ctf + off, n * sizeof(float), cudaMemcpyHostToDevice, st));
        if (ssig) THB_CUDA(ctx, cudaMemcpyAsync(ssig, sigRcp + off, n * sizeof(float), cudaMemcpyHostToDevice, st));
        const size_t doff = (size_t)(base + i0) * P;
        dim3 grid(std::min((P + 255) / 256, 64), c);
        if (!async) span_begin(ctx, KF_PACK);
        permute_stack_kernel<<<grid, 256, 0, st>>>(sdat, sctf, ssig, perm, P, c, s.dat + doff, s.ctf + doff,
                                                   ssig ? s.sig + doff : nullptr);
        if (!async) span_end(ctx);
        ctx->launches++;
        THB_CUDA(ctx, cudaGetLastError());
    }
    if (slotOfImg)
        THB_CUDA(ctx, cudaMemcpyAsync(s.slot + base, slotOfImg, (size_t)nImg * sizeof(int), cudaMemcpyHostToDevice, st));
    else
        THB_CUDA(ctx, cudaMemsetAsync(s.slot + base, 0, (size_t)nImg * sizeof(int), st));
    for (int i = 0; i < nImg; ++i) s.hslot[(size_t)base + i] = slotOfImg ? slotOfImg[i] : 0;
    if (async) {
        THB_CUDA(ctx, cudaEventRecord(ctx->copyDone, st));
        ctx->copyPending = true;
        return THB_OK;
    }
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_upload_stack_at(thb_ctx* ctx, int kind, int base, int nImg, const float* dat, const float* ctf, const float* sigRcp,
                        const int* slotOfImg)
{
    return upload_stack_impl(ctx, kind, base, nImg, dat, ctf, sigRcp, slotOfImg, false);
}

// The same upload on a second stream, returning at once: the copies and the re-layout overlap whatever the compute stream is
// doing (the previous batch's kernels).  Contract: the host arrays stay valid and the images [base, base+nImg) are neither
// read nor written by other calls until thb_upload_wait() returns.
int thb_upload_stack_at_async(thb_ctx* ctx, int kind, int base, int nImg, const float* dat, const float* ctf, const float* sigRcp,
                              const int* slotOfImg)
{
    return upload_stack_impl(ctx, kind, base, nImg, dat, ctf, sigRcp, slotOfImg, true);
}

int thb_upload_wait(thb_ctx* ctx)
{
    if (!ctx) return THB_E_ARG;
    if (!ctx->copyPending) return THB_OK;
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copyDone, 0));
    THB_CUDA(ctx, cudaEventSynchronize(ctx->copyDone));
    ctx->copyPending = false;
    return THB_OK;
}

// a2 on the device: images [base, base+nImg) of a reserved stack packed from full half-complex FTs
int thb_pack_stack(thb_ctx* ctx, int kind, int base, int nImg, const float* imgFT, const int* iPxl, const int* iSig,
                   const float* sigRcpTab, int nGroup, int nRing, const int* groupOfImg, const float* ctfAttr, float pixelSize,
                   const int* slotOfImg)
{
    if (!ctx) return THB_E_ARG;
    if (kind != THB_STACK_EXPECT && kind != THB_STACK_INSERT) return set_error(ctx, THB_E_ARG, "pack_stack: bad kind");
    const bool E = kind == THB_STACK_EXPECT;
    const int P = E ? ctx->nPxlE : ctx->nPxlM;
    const int N = E ? ctx->N : ctx->NM;
    const int4* pix = E ? ctx->pixE : ctx->pixM;
    const int* perm = E ? ctx->permE : ctx->permM;
    Stack& s = E ? ctx->stackE : ctx->stackM;
    if (P <= 0) return set_error(ctx, THB_E_STATE, "pack_stack: pixel list for this stack kind not set");
    if (!s.dat) return set_error(ctx, THB_E_STATE, "pack_stack: stack not reserved (thb_stack_reserve)");
    if (nImg <= 0 || !imgFT || !iPxl || !ctfAttr || pixelSize <= 0) return set_error(ctx, THB_E_ARG, "pack_stack: bad arguments");
    if (E && (!iSig || !sigRcpTab || nGroup <= 0 || nRing <= 0)) return set_error(ctx, THB_E_ARG, "pack_stack: sigma table required for the E stack");
    if (base < 0 || base + nImg > s.nImg) return set_error(ctx, THB_E_ARG, "pack_stack: images [%d,%d) exceed the capacity %d", base, base + nImg, s.nImg);
    const size_t imgElems = (size_t)(N / 2 + 1) * N;
    for (int i = 0; i < P; ++i)
        if (iPxl[i] < 0 || (size_t)iPxl[i] >= imgElems || (E && (iSig[i] < 0 || iSig[i] >= nRing)))
            return set_error(ctx, THB_E_ARG, "pack_stack: iPxl / iSig[%d] out of range", i);
    for (int l = 0; l < nImg; ++l) {
        if (groupOfImg && (groupOfImg[l] < 0 || groupOfImg[l] >= std::max(nGroup, 1))) return set_error(ctx, THB_E_ARG, "pack_stack: groupOfImg[%d] out of range", l);
        if (slotOfImg && (slotOfImg[l] < 0 || slotOfImg[l] >= THB_MAX_SLOTS)) return set_error(ctx, THB_E_ARG, "pack_stack: slotOfImg[%d] out of range", l);
    }
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    int* dIdx = (int*)scratch(ctx, 0, sizeof(int) * (2 * (size_t)P + nImg) + sizeof(float) * ((size_t)std::max(nGroup, 1) * std::max(nRing, 1) + 7 * (size_t)nImg));
    if (!dIdx) return THB_E_CUDA;
    int* dPxl = dIdx; int* dSig = dPxl + P; int* dGrp = dSig + P;
    float* dTab = (float*)(dGrp + nImg); float* dAttr = dTab + (size_t)std::max(nGroup, 1) * std::max(nRing, 1);
    THB_CUDA(ctx, cudaMemcpyAsync(dPxl, iPxl, sizeof(int) * P, cudaMemcpyHostToDevice, ctx->stream));
    if (E) {
        THB_CUDA(ctx, cudaMemcpyAsync(dSig, iSig, sizeof(int) * P, cudaMemcpyHostToDevice, ctx->stream));
        THB_CUDA(ctx, cudaMemcpyAsync(dTab, sigRcpTab, sizeof(float) * (size_t)nGroup * nRing, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (groupOfImg) THB_CUDA(ctx, cudaMemcpyAsync(dGrp, groupOfImg, sizeof(int) * nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dAttr, ctfAttr, sizeof(float) * 7 * (size_t)nImg, cudaMemcpyHostToDevice, ctx->stream));
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nImg, ((size_t)256 << 20) / (imgElems * sizeof(float2))));
    float2* dImg = (float2*)scratch(ctx, 4, (size_t)chunk * imgElems * sizeof(float2));
    if (!dImg) return THB_E_CUDA;
    for (int i0 = 0; i0 < nImg; i0 += chunk) {
        const int c = std::min(chunk, nImg - i0);
        THB_CUDA(ctx, cudaMemcpyAsync(dImg, imgFT + 2 * (size_t)i0 * imgElems, (size_t)c * imgElems * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
        const size_t doff = (size_t)(base + i0) * P;
        dim3 grid(std::min((P + 255) / 256, 64), c);
        span_begin(ctx, KF_PACK);
        pack_stack_kernel<<<grid, 256, 0, ctx->stream>>>(dImg, imgElems, pix, perm, dPxl, E ? dSig : nullptr, P, E ? dTab : nullptr, nRing,
                                                         groupOfImg ? dGrp + i0 : nullptr, reinterpret_cast<const CtfAttr7*>(dAttr) + i0, pixelSize, N,
                                                         s.dat + doff, s.ctf + doff, E ? s.sig + doff : nullptr);
        span_end(ctx);
        ctx->launches++;
        THB_CUDA(ctx, cudaGetLastError());
    }
    if (slotOfImg)
        THB_CUDA(ctx, cudaMemcpyAsync(s.slot + base, slotOfImg, (size_t)nImg * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    else
        THB_CUDA(ctx, cudaMemsetAsync(s.slot + base, 0, (size_t)nImg * sizeof(int), ctx->stream));
    for (int i = 0; i < nImg; ++i) s.hslot[(size_t)base + i] = slotOfImg ? slotOfImg[i] : 0;
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

// resident stack back to the host in the caller's pixel order (any output may be NULL)
int thb_download_stack(thb_ctx* ctx, int kind, int base, int nImg, float* dat, float* ctf, float* sigRcp)
{
    if (!ctx) return THB_E_ARG;
    if (kind != THB_STACK_EXPECT && kind != THB_STACK_INSERT) return set_error(ctx, THB_E_ARG, "download_stack: bad kind");
    const bool E = kind == THB_STACK_EXPECT;
    const int P = E ? ctx->nPxlE : ctx->nPxlM;
    const int* perm = E ? ctx->permE : ctx->permM;
    Stack& s = E ? ctx->stackE : ctx->stackM;
    if (!s.dat) return set_error(ctx, THB_E_STATE, "download_stack: no stack");
    if (nImg <= 0 || base < 0 || base + nImg > s.nImg) return set_error(ctx, THB_E_ARG, "download_stack: bad range");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)nImg * P;
    float2* dd = (float2*)scratch(ctx, 4, n * sizeof(float2));
    float* dc = (float*)scratch(ctx, 5, n * sizeof(float));
    float* ds = (float*)scratch(ctx, 6, n * sizeof(float));
    if (!dd || !dc || !ds) return THB_E_CUDA;
    const size_t off = (size_t)base * P;
    dim3 grid(std::min((P + 255) / 256, 64), nImg);
    unpermute_stack_kernel<<<grid, 256, 0, ctx->stream>>>(s.dat + off, s.ctf + off, s.sig ? s.sig + off : nullptr, perm, P, dd, dc, ds);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    if (dat) THB_CUDA(ctx, cudaMemcpyAsync(dat, dd, n * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctf) THB_CUDA(ctx, cudaMemcpyAsync(ctf, dc, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (sigRcp && s.sig) THB_CUDA(ctx, cudaMemcpyAsync(sigRcp, ds, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_upload_stack(thb_ctx* ctx, int kind, int nImg, const float* dat, const float* ctf, const float* sigRcp,
                     const int* slotOfImg)
{
    int rc = thb_stack_reserve(ctx, kind, nImg);
    if (rc) return rc;
    return thb_upload_stack_at(ctx, kind, 0, nImg, dat, ctf, sigRcp, slotOfImg);
}

int thb_project(thb_ctx* ctx, int slot, int nRot, const double* quat, float* dst)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || !ctx->vols[slot].d) return set_error(ctx, THB_E_STATE, "project: no volume in slot %d", slot);
    if (!ctx->pixE) return set_error(ctx, THB_E_STATE, "project: E pixel list not set");
    if (nRot <= 0 || !quat || !dst) return set_error(ctx, THB_E_ARG, "project: bad arguments");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int P = ctx->nPxlE;
    const int qc = ctx->mode2D ? 2 : 4;
    double* dq = (double*)scratch(ctx, 0, sizeof(double) * qc * (size_t)nRot);
    float2* dd = (float2*)scratch(ctx, 1, sizeof(float2) * (size_t)nRot * P);
    if (!dq || !dd) return THB_E_CUDA;
    THB_CUDA(ctx, cudaMemcpyAsync(dq, quat, sizeof(double) * qc * (size_t)nRot, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((P + 255) / 256, nRot);
    span_begin(ctx, KF_EXPECT);
    project_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->vols[slot].d, ctx->vols[slot].vdim, ctx->vols[slot].pitch, ctx->pixE, ctx->permE, P, dq, dd, ctx->mode2D);
    span_end(ctx);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    THB_CUDA(ctx, cudaMemcpyAsync(dst, dd, sizeof(float2) * (size_t)nRot * P, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

}  // extern "C"
namespace thb {
int check_expect_state(thb_ctx* ctx, const char* who)
{
    if (!ctx->pixE) return set_error(ctx, THB_E_STATE, "%s: E pixel list not set", who);
    if (!ctx->stackE.dat) return set_error(ctx, THB_E_STATE, "%s: E stack not uploaded", who);
    int vdim = 0;
    for (int i = 0; i < THB_MAX_SLOTS; ++i)
        if (ctx->vols[i].d) {
            if (vdim && ctx->vols[i].vdim != vdim) return set_error(ctx, THB_E_STATE, "%s: volumes of different size", who);
            vdim = ctx->vols[i].vdim;
        }
    if (!vdim) return set_error(ctx, THB_E_STATE, "%s: no projector volume uploaded", who);
    if (vdim < ctx->pf * ctx->N) return set_error(ctx, THB_E_STATE, "%s: volume dimension %d < pf*N = %d", who, vdim, ctx->pf * ctx->N);
    return vdim;
}
}  // namespace thb
extern "C" {

int thb_expect_local(thb_ctx* ctx, int nAct, const int* imgIdx, int nR, int nT, const double* quat, const double* tran,
                     const double* wR, const double* wT, float* uR, float* uT, float* uC, float* base, float* logL)
{
    if (!ctx) return THB_E_ARG;
    const int vdim = check_expect_state(ctx, "expect_local");
    if (vdim < 0) return vdim;
    if (nAct <= 0 || nR <= 0 || nT <= 0 || !quat || !tran || !wR || !wT) return set_error(ctx, THB_E_ARG, "expect_local: bad arguments");
    if (imgIdx)
        for (int i = 0; i < nAct; ++i)
            if (imgIdx[i] < 0 || imgIdx[i] >= ctx->stackE.nImg)
                return set_error(ctx, THB_E_ARG, "expect_local: imgIdx[%d] = %d outside the stack", i, imgIdx[i]);
    if (!imgIdx && nAct > ctx->stackE.nImg) return set_error(ctx, THB_E_ARG, "expect_local: nAct exceeds the stack");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));

    const int qc = ctx->mode2D ? 2 : 4;      // MODE_2D: quat[nAct][nR][2] = (cos, sin)
    const size_t nq = (size_t)nAct * nR * qc, nt = (size_t)nAct * nT * 2, nwr = (size_t)nAct * nR, nwt = (size_t)nAct * nT;
    double* din = (double*)scratch(ctx, 0, sizeof(double) * (nq + nt + nwr + nwt) + sizeof(int) * (size_t)nAct);
    const size_t nout = nwr + nwt + 2 * (size_t)nAct + (logL ? nwr * nT : 0);
    float* dout = (float*)scratch(ctx, 1, sizeof(float) * nout);
    if (!din || !dout) return THB_E_CUDA;
    double* dq = din; double* dt = dq + nq; double* dwr = dt + nt; double* dwt = dwr + nwr;
    int* didx = (int*)(dwt + nwt);
    THB_CUDA(ctx, cudaMemcpyAsync(dq, quat, sizeof(double) * nq, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dt, tran, sizeof(double) * nt, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwr, wR, sizeof(double) * nwr, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwt, wT, sizeof(double) * nwt, cudaMemcpyHostToDevice, ctx->stream));
    if (imgIdx) THB_CUDA(ctx, cudaMemcpyAsync(didx, imgIdx, sizeof(int) * (size_t)nAct, cudaMemcpyHostToDevice, ctx->stream));

    ExpectArgs a;
    memset(&a, 0, sizeof(a));
    a.vols = vol_table(ctx);
    a.vdim = vdim; a.pitch = vol_pitch(vdim);
    a.dat = ctx->stackE.dat; a.ctf = ctx->stackE.ctf; a.sig = ctx->stackE.sig; a.slotOfImg = ctx->stackE.slot;
    a.pix = ctx->pixE; a.P = ctx->nPxlE; a.N = ctx->N;
    a.nAct = nAct; a.imgIdx = imgIdx ? didx : nullptr; a.imgBase = 0; a.active = nullptr;
    a.nR = nR; a.nT = nT; a.slotAll = -1;
    a.quat = View3{dq, (long long)nR * qc, qc, 1};
    a.tran = View3{dt, (long long)nT * 2, 2, 1};
    a.wR = View3{dwr, (long long)nR, 1, 0};
    a.wT = View3{dwt, (long long)nT, 1, 0};
    a.uR = dout; a.uT = a.uR + nwr; a.uC = a.uT + nwt; a.base = a.uC + nAct;
    a.logL = logL ? a.base + nAct : nullptr;
    int rc = launch_expect_local(ctx, a);
    if (rc) return rc;
    if (uR) THB_CUDA(ctx, cudaMemcpyAsync(uR, a.uR, sizeof(float) * nwr, cudaMemcpyDeviceToHost, ctx->stream));
    if (uT) THB_CUDA(ctx, cudaMemcpyAsync(uT, a.uT, sizeof(float) * nwt, cudaMemcpyDeviceToHost, ctx->stream));
    if (uC) THB_CUDA(ctx, cudaMemcpyAsync(uC, a.uC, sizeof(float) * nAct, cudaMemcpyDeviceToHost, ctx->stream));
    if (base) THB_CUDA(ctx, cudaMemcpyAsync(base, a.base, sizeof(float) * nAct, cudaMemcpyDeviceToHost, ctx->stream));
    if (logL) THB_CUDA(ctx, cudaMemcpyAsync(logL, a.logL, sizeof(float) * nwr * nT, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

// ------------------------------------------------------------------------------------------------ CTF search (SEARCH_TYPE_CTF)
__global__ void permute_float_kernel(const float* __restrict__ src, const int* __restrict__ perm, int P, int nImg, float* __restrict__ dst)
{
    const int l = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) dst[(size_t)l * P + i] = src[(size_t)l * P + perm[i]];
}

int thb_set_frequency(thb_ctx* ctx, const float* freQ)
{
    if (!ctx) return THB_E_ARG;
    if (!ctx->pixE || !freQ) return set_error(ctx, THB_E_STATE, "set_frequency: E pixel list not set / NULL table");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int P = ctx->nPxlE;
    float* tmp = (float*)scratch(ctx, 0, sizeof(float) * (size_t)P);
    if (!tmp) return THB_E_CUDA;
    if (!ctx->freqE) THB_CUDA(ctx, cudaMalloc(&ctx->freqE, sizeof(float) * (size_t)P));
    THB_CUDA(ctx, cudaMemcpyAsync(tmp, freQ, sizeof(float) * (size_t)P, cudaMemcpyHostToDevice, ctx->stream));
    permute_float_kernel<<<dim3(std::min((P + 255) / 256, 64), 1), 256, 0, ctx->stream>>>(tmp, ctx->permE, P, 1, ctx->freqE);
    ctx->launches++;
    THB_CUDA(ctx, cudaGetLastError());
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_upload_stack_defocus(thb_ctx* ctx, int base, int nImg, const float* defP)
{
    if (!ctx) return THB_E_ARG;
    Stack& s = ctx->stackE;
    if (!s.dat) return set_error(ctx, THB_E_STATE, "upload_stack_defocus: E stack not reserved");
    if (nImg <= 0 || !defP || base < 0 || base + nImg > s.nImg) return set_error(ctx, THB_E_ARG, "upload_stack_defocus: bad range / NULL array");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int P = ctx->nPxlE;
    if (!s.def) THB_CUDA(ctx, cudaMalloc(&s.def, sizeof(float) * (size_t)s.nImg * P));
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nImg, ((size_t)256 << 20) / ((size_t)P * 4)));
    float* tmp = (float*)scratch(ctx, 4, sizeof(float) * (size_t)chunk * P);
    if (!tmp) return THB_E_CUDA;
    for (int i0 = 0; i0 < nImg; i0 += chunk) {
        const int c = std::min(chunk, nImg - i0);
        THB_CUDA(ctx, cudaMemcpyAsync(tmp, defP + (size_t)i0 * P, sizeof(float) * (size_t)c * P, cudaMemcpyHostToDevice, ctx->stream));
        permute_float_kernel<<<dim3(std::min((P + 255) / 256, 64), c), 256, 0, ctx->stream>>>(tmp, ctx->permE, P, c, s.def + (size_t)(base + i0) * P);
        ctx->launches++;
        THB_CUDA(ctx, cudaGetLastError());
        THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return THB_OK;
}

int thb_expect_local_ctf(thb_ctx* ctx, int nAct, const int* imgIdx, int nR, int nT, int nD, const double* quat, const double* tran,
                         const double* dpar, const double* wR, const double* wT, const double* wD, const float* ctfK, float* uR, float* uT,
                         float* uD, float* uC, float* base, float* logL)
{
    if (!ctx) return THB_E_ARG;
    const int vdim = check_expect_state(ctx, "expect_local_ctf");
    if (vdim < 0) return vdim;
    if (ctx->mode2D) return set_error(ctx, THB_E_STATE, "expect_local_ctf: MODE_3D only");
    if (!ctx->freqE || !ctx->stackE.def) return set_error(ctx, THB_E_STATE, "expect_local_ctf: frequency table / per-pixel defocus not uploaded (thb_set_frequency, thb_upload_stack_defocus)");
    if (nAct <= 0 || nR <= 0 || nT <= 0 || nD <= 0 || !quat || !tran || !dpar || !wR || !wT || !wD || !ctfK)
        return set_error(ctx, THB_E_ARG, "expect_local_ctf: bad arguments");
    if (imgIdx)
        for (int i = 0; i < nAct; ++i)
            if (imgIdx[i] < 0 || imgIdx[i] >= ctx->stackE.nImg) return set_error(ctx, THB_E_ARG, "expect_local_ctf: imgIdx[%d] outside the stack", i);
    if (!imgIdx && nAct > ctx->stackE.nImg) return set_error(ctx, THB_E_ARG, "expect_local_ctf: nAct exceeds the stack");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nq = (size_t)nAct * nR * 4, nt = (size_t)nAct * nT * 2, nwr = (size_t)nAct * nR, nwt = (size_t)nAct * nT, nwd = (size_t)nAct * nD;
    double* din = (double*)scratch(ctx, 0, sizeof(double) * (nq + nt + nwr + nwt + 2 * nwd) + sizeof(float) * 4 * (size_t)nAct + sizeof(int) * (size_t)nAct);
    const size_t nL = nwr * nT * nD;
    const size_t nout = nwr + nwt + nwd + 2 * (size_t)nAct + (logL ? nL : 0);
    float* dout = (float*)scratch(ctx, 1, sizeof(float) * nout);
    if (!din || !dout) return THB_E_CUDA;
    double* dq = din; double* dt = dq + nq; double* dwr = dt + nt; double* dwt = dwr + nwr; double* dwd = dwt + nwt; double* dd = dwd + nwd;
    float* dk = (float*)(dd + nwd);
    int* didx = (int*)(dk + 4 * (size_t)nAct);
    THB_CUDA(ctx, cudaMemcpyAsync(dq, quat, sizeof(double) * nq, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dt, tran, sizeof(double) * nt, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwr, wR, sizeof(double) * nwr, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwt, wT, sizeof(double) * nwt, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwd, wD, sizeof(double) * nwd, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dd, dpar, sizeof(double) * nwd, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dk, ctfK, sizeof(float) * 4 * (size_t)nAct, cudaMemcpyHostToDevice, ctx->stream));
    if (imgIdx) THB_CUDA(ctx, cudaMemcpyAsync(didx, imgIdx, sizeof(int) * (size_t)nAct, cudaMemcpyHostToDevice, ctx->stream));
    ExpectArgs a;
    memset(&a, 0, sizeof(a));
    a.vols = vol_table(ctx);
    a.vdim = vdim; a.pitch = vol_pitch(vdim);
    a.dat = ctx->stackE.dat; a.ctf = ctx->stackE.ctf; a.sig = ctx->stackE.sig; a.slotOfImg = ctx->stackE.slot;
    a.pix = ctx->pixE; a.P = ctx->nPxlE; a.N = ctx->N;
    a.nAct = nAct; a.imgIdx = imgIdx ? didx : nullptr; a.nR = nR; a.nT = nT; a.slotAll = -1;
    a.quat = View3{dq, (long long)nR * 4, 4, 1};
    a.tran = View3{dt, (long long)nT * 2, 2, 1};
    a.wR = View3{dwr, (long long)nR, 1, 0};
    a.wT = View3{dwt, (long long)nT, 1, 0};
    a.nD = nD; a.defP = ctx->stackE.def; a.freq = ctx->freqE; a.ctfK = dk;
    a.dpar = View3{dd, (long long)nD, 1, 0};
    a.wD = View3{dwd, (long long)nD, 1, 0};
    a.uR = dout; a.uT = a.uR + nwr; a.uD = a.uT + nwt; a.uC = a.uD + nwd; a.base = a.uC + nAct;
    a.logL = logL ? a.base + nAct : nullptr;
    const int saved = ctx->expectImpl;
    ctx->expectImpl = 3;                         // the CTF search lives in the default (cell-layout) kernel only
    int rc = launch_expect_local(ctx, a);
    ctx->expectImpl = saved;
    if (rc) return rc;
    if (uR) THB_CUDA(ctx, cudaMemcpyAsync(uR, a.uR, sizeof(float) * nwr, cudaMemcpyDeviceToHost, ctx->stream));
    if (uT) THB_CUDA(ctx, cudaMemcpyAsync(uT, a.uT, sizeof(float) * nwt, cudaMemcpyDeviceToHost, ctx->stream));
    if (uD) THB_CUDA(ctx, cudaMemcpyAsync(uD, a.uD, sizeof(float) * nwd, cudaMemcpyDeviceToHost, ctx->stream));
    if (uC) THB_CUDA(ctx, cudaMemcpyAsync(uC, a.uC, sizeof(float) * nAct, cudaMemcpyDeviceToHost, ctx->stream));
    if (base) THB_CUDA(ctx, cudaMemcpyAsync(base, a.base, sizeof(float) * nAct, cudaMemcpyDeviceToHost, ctx->stream));
    if (logL) THB_CUDA(ctx, cudaMemcpyAsync(logL, a.logL, sizeof(float) * nL, cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

// Global-scan shape (src/Optimiser.cpp:633-914): one shared rotation / translation set against every image.  The fused kernel
// takes the whole rotation set in one launch (passes of 128 rotations inside the kernel, log-likelihood table in a global scratch
// buffer, baseline = the maximum over the whole table, marginals in its epilogue); what is chunked is the IMAGE list, so that the
// table fits a scratch budget - no merging of partial results, no host round trip per chunk.
int thb_expect_scan(thb_ctx* ctx, int slot, int nR, int nT, const double* quat, const double* tran, const double* pR,
                    const double* pT, float* wC, float* wR, float* wT, float* base, float* logL)
{
    if (!ctx) return THB_E_ARG;
    return thb_expect_scan_range(ctx, slot, 0, ctx->stackE.nImg, nR, nT, quat, tran, pR, pT, wC, wR, wT, base, logL);
}

// the same for the images [imgBase, imgBase + nImgRange) of the stack only; output rows are relative to imgBase
int thb_expect_scan_range(thb_ctx* ctx, int slot, int imgBase, int nImgRange, int nR, int nT, const double* quat, const double* tran,
                          const double* pR, const double* pT, float* wC, float* wR, float* wT, float* base, float* logL)
{
    if (!ctx) return THB_E_ARG;
    const int vdim = check_expect_state(ctx, "expect_scan");
    if (vdim < 0) return vdim;
    if (nR <= 0 || nT <= 0 || !quat || !tran || !pR || !pT) return set_error(ctx, THB_E_ARG, "expect_scan: bad arguments");
    if (slot < 0 || slot >= THB_MAX_SLOTS || !ctx->vols[slot].d) return set_error(ctx, THB_E_STATE, "expect_scan: no volume in slot %d", slot);
    if (imgBase < 0 || nImgRange <= 0 || imgBase + nImgRange > ctx->stackE.nImg) return set_error(ctx, THB_E_ARG, "expect_scan: images [%d,%d) outside the stack", imgBase, imgBase + nImgRange);
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nImg = nImgRange;
    const std::vector<int>& hslot = ctx->stackE.hslot;
    std::vector<int> idx;           // positions within the range
    // MODE_2D (classification, src/Optimiser.cpp:756-914): every image is compared with every class reference
    for (int i = 0; i < nImg; ++i) if (ctx->mode2D || hslot[imgBase + i] == slot) idx.push_back(i);
    const int nAll = (int)idx.size();
    if (wC) memset(wC, 0, sizeof(float) * nImg);
    if (wR) memset(wR, 0, sizeof(float) * (size_t)nImg * nR);
    if (wT) memset(wT, 0, sizeof(float) * (size_t)nImg * nT);
    if (base) memset(base, 0, sizeof(float) * nImg);
    if (logL) memset(logL, 0, sizeof(float) * (size_t)nImg * nR * nT);
    if (nAll == 0) return THB_OK;

    const int qc = ctx->mode2D ? 2 : 4;
    const size_t nRT = (size_t)nR * nT;
    // images per launch: the table (and the optional copy of it) within 1 GiB, at least one wave of CTAs where the list allows
    const size_t budget = (size_t)1 << 30;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nAll, budget / (nRT * sizeof(float) * (logL ? 2 : 1))));
    double* din = (double*)scratch(ctx, 0, sizeof(double) * ((size_t)nR * 4 + nT * 2 + nR + nT) + sizeof(int) * (size_t)nAll);
    const size_t nout = (size_t)chunk * nR + (size_t)chunk * nT + 2 * (size_t)chunk + (logL ? (size_t)chunk * nRT : 0);
    float* dout = (float*)scratch(ctx, 1, sizeof(float) * nout);
    if (!din || !dout) return THB_E_CUDA;
    double* dq = din; double* dt = dq + (size_t)nR * 4; double* dwr = dt + nT * 2; double* dwt = dwr + nR;
    int* didx = (int*)(dwt + nT);
    THB_CUDA(ctx, cudaMemcpyAsync(dq, quat, sizeof(double) * qc * (size_t)nR, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dt, tran, sizeof(double) * 2 * nT, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwr, pR, sizeof(double) * nR, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwt, pT, sizeof(double) * nT, cudaMemcpyHostToDevice, ctx->stream));
    {
        std::vector<int> abs_idx(nAll);
        for (int i = 0; i < nAll; ++i) abs_idx[i] = imgBase + idx[i];
        THB_CUDA(ctx, cudaMemcpyAsync(didx, abs_idx.data(), sizeof(int) * (size_t)nAll, cudaMemcpyHostToDevice, ctx->stream));
        THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    std::vector<float> hR, hT, hC, hB, hL;
    const bool scatter = nAll != nImg;        // results go to the rows of the selected images
    for (int l0 = 0; l0 < nAll; l0 += chunk) {
        const int nAct = std::min(chunk, nAll - l0);
        ExpectArgs a;
        memset(&a, 0, sizeof(a));
        a.vols = vol_table(ctx); a.vdim = vdim; a.pitch = vol_pitch(vdim);
        a.dat = ctx->stackE.dat; a.ctf = ctx->stackE.ctf; a.sig = ctx->stackE.sig; a.slotOfImg = ctx->stackE.slot;
        a.pix = ctx->pixE; a.P = ctx->nPxlE; a.N = ctx->N;
        a.nAct = nAct; a.imgIdx = didx + l0; a.nR = nR; a.nT = nT;
        a.slotAll = ctx->mode2D ? slot : -1;
        a.scanSlot1 = slot + 1;
        a.quat = View3{dq, 0, qc, 1};
        a.tran = View3{dt, 0, 2, 1};
        a.wR = View3{dwr, 0, 1, 0};
        a.wT = View3{dwt, 0, 1, 0};
        a.uR = dout; a.uT = a.uR + (size_t)nAct * nR; a.uC = a.uT + (size_t)nAct * nT; a.base = a.uC + nAct;
        a.logL = logL ? a.base + nAct : nullptr;
        int rc = launch_expect_local(ctx, a);
        if (rc) return rc;
        if (!scatter) {
            // the selection is the whole stack in order: straight into the caller's arrays
            if (wR) THB_CUDA(ctx, cudaMemcpyAsync(wR + (size_t)l0 * nR, a.uR, sizeof(float) * (size_t)nAct * nR, cudaMemcpyDeviceToHost, ctx->stream));
            if (wT) THB_CUDA(ctx, cudaMemcpyAsync(wT + (size_t)l0 * nT, a.uT, sizeof(float) * (size_t)nAct * nT, cudaMemcpyDeviceToHost, ctx->stream));
            if (wC) THB_CUDA(ctx, cudaMemcpyAsync(wC + l0, a.uC, sizeof(float) * nAct, cudaMemcpyDeviceToHost, ctx->stream));
            if (base) THB_CUDA(ctx, cudaMemcpyAsync(base + l0, a.base, sizeof(float) * nAct, cudaMemcpyDeviceToHost, ctx->stream));
            if (logL) THB_CUDA(ctx, cudaMemcpyAsync(logL + (size_t)l0 * nRT, a.logL, sizeof(float) * (size_t)nAct * nRT, cudaMemcpyDeviceToHost, ctx->stream));
            THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the scratch table is reused by the next chunk
            continue;
        }
        hR.resize((size_t)nAct * nR); hT.resize((size_t)nAct * nT); hC.resize(nAct); hB.resize(nAct);
        THB_CUDA(ctx, cudaMemcpyAsync(hR.data(), a.uR, sizeof(float) * hR.size(), cudaMemcpyDeviceToHost, ctx->stream));
        THB_CUDA(ctx, cudaMemcpyAsync(hT.data(), a.uT, sizeof(float) * hT.size(), cudaMemcpyDeviceToHost, ctx->stream));
        THB_CUDA(ctx, cudaMemcpyAsync(hC.data(), a.uC, sizeof(float) * nAct, cudaMemcpyDeviceToHost, ctx->stream));
        THB_CUDA(ctx, cudaMemcpyAsync(hB.data(), a.base, sizeof(float) * nAct, cudaMemcpyDeviceToHost, ctx->stream));
        if (logL) {
            hL.resize((size_t)nAct * nRT);
            THB_CUDA(ctx, cudaMemcpyAsync(hL.data(), a.logL, sizeof(float) * hL.size(), cudaMemcpyDeviceToHost, ctx->stream));
        }
        THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int l = 0; l < nAct; ++l) {
            const int i = idx[l0 + l];
            if (wC) wC[i] = hC[l];
            if (base) base[i] = hB[l];
            if (wR) memcpy(wR + (size_t)i * nR, hR.data() + (size_t)l * nR, sizeof(float) * nR);
            if (wT) memcpy(wT + (size_t)i * nT, hT.data() + (size_t)l * nT, sizeof(float) * nT);
            if (logL) memcpy(logL + (size_t)i * nRT, hL.data() + (size_t)l * nRT, sizeof(float) * nRT);
        }
    }
    return THB_OK;
}

// MODE_2D classification scan of all classes at once (ExpectGlobal2D's shape, gpu/interface/Interface.h:176-198): the templates
// of every (class, rotation) pair in ONE table, every image contracted against all of them in one launch, one baseline per image
// across the classes.  See scan_classes_epilogue_kernel (thb_expect8.cuh).
int thb_expect_scan_classes(thb_ctx* ctx, int nK, int imgBase, int nImg, int nR, int nT, const double* quat, const double* tran,
                            const double* pR, const double* pT, float* wC, float* wR, float* wT, float* base)
{
    if (!ctx) return THB_E_ARG;
    if (!ctx->mode2D) return set_error(ctx, THB_E_STATE, "expect_scan_classes: MODE_2D only (thb_set_mode)");
    const int vdim = check_expect_state(ctx, "expect_scan_classes");
    if (vdim < 0) return vdim;
    if (nK <= 0 || nK > THB_MAX_SLOTS || nR <= 0 || nT <= 0 || !quat || !tran || !pR || !pT || !wC || !wR || !wT || !base)
        return set_error(ctx, THB_E_ARG, "expect_scan_classes: bad arguments");
    for (int k = 0; k < nK; ++k)
        if (!ctx->vols[k].d) return set_error(ctx, THB_E_STATE, "expect_scan_classes: no class reference in slot %d", k);
    if (imgBase < 0 || nImg <= 0 || imgBase + nImg > ctx->stackE.nImg)
        return set_error(ctx, THB_E_ARG, "expect_scan_classes: images [%d,%d) outside the stack", imgBase, imgBase + nImg);
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int P = ctx->nPxlE;
    const size_t nF = (size_t)nK * nR;
    const int nFpad = (int)((nF + E8_WROT - 1) / E8_WROT * E8_WROT);
    const size_t tbytes = (size_t)P * nFpad * sizeof(float2);
    if (tbytes > ((size_t)2 << 30)) return set_error(ctx, THB_E_ARG, "expect_scan_classes: template table of %zu MB; scan class by class", tbytes >> 20);
    const bool fresh = ctx->scratchCap[15] < tbytes || !ctx->scratch[15];
    float2* tmpl = (float2*)scratch(ctx, 15, tbytes);
    if (!tmpl) return THB_E_CUDA;
    if (fresh) THB_CUDA(ctx, cudaMemsetAsync(tmpl, 0, ctx->scratchCap[15], ctx->stream));
    double* din = (double*)scratch(ctx, 0, sizeof(double) * ((size_t)nR * 2 + nT * 2 + nR + nT));
    if (!din) return THB_E_CUDA;
    double* dq = din; double* dt = dq + (size_t)nR * 2; double* dwr = dt + nT * 2; double* dwt = dwr + nR;
    THB_CUDA(ctx, cudaMemcpyAsync(dq, quat, sizeof(double) * 2 * (size_t)nR, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dt, tran, sizeof(double) * 2 * nT, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwr, pR, sizeof(double) * nR, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dwt, pT, sizeof(double) * nT, cudaMemcpyHostToDevice, ctx->stream));
    const size_t budget = (size_t)1 << 30;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nImg, budget / (nF * nT * sizeof(float))));
    float* table = (float*)scratch(ctx, 7, sizeof(float) * (size_t)chunk * nF * nT);
    const size_t nout = (size_t)chunk * nK + (size_t)nK * chunk * nR + (size_t)nK * chunk * nT + chunk;
    float* dout = (float*)scratch(ctx, 1, sizeof(float) * nout);
    if (!table || !dout) return THB_E_CUDA;
    const bool tc15 = nT > E_TC;
    const size_t smem = tc15 ? e8_smem_bytes<E3_TC_SCAN>() : e8_smem_bytes<E_TC>();
    if (tc15)
        THB_CUDA(ctx, cudaFuncSetAttribute(scan_contract_kernel<E3_TC_SCAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        THB_CUDA(ctx, cudaFuncSetAttribute(scan_contract_kernel<E_TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    span_begin(ctx, KF_EXPECT);
    for (int k = 0; k < nK; ++k) {
        const Volume3& v = ctx->vols[k];
        scan_project_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(v.d, v.vdim, v.pitch, ctx->pixE, P, View3{dq, 0, 2, 1}, 0, nR, nFpad, 1, tmpl, k * nR);
    }
    span_end(ctx);
    ctx->launches += nK;
    for (int l0 = 0; l0 < nImg; l0 += chunk) {
        const int nAct = std::min(chunk, nImg - l0);
        ExpectArgs a;
        memset(&a, 0, sizeof(a));
        a.vdim = vdim; a.pitch = vol_pitch(vdim);
        a.dat = ctx->stackE.dat; a.ctf = ctx->stackE.ctf; a.sig = ctx->stackE.sig; a.slotOfImg = ctx->stackE.slot;
        a.pix = ctx->pixE; a.P = P; a.N = ctx->N; a.mode2D = 1;
        a.nAct = nAct; a.imgIdx = nullptr; a.imgBase = imgBase + l0; a.nR = (int)nF; a.nT = nT;
        a.tran = View3{dt, 0, 2, 1};
        float* dWC = dout; float* dWR = dWC + (size_t)nAct * nK; float* dWT = dWR + (size_t)nK * nAct * nR; float* dB = dWT + (size_t)nK * nAct * nT;
        span_begin(ctx, KF_EXPECT);
        if (tc15)
            scan_contract_kernel<E3_TC_SCAN><<<nAct, E8_THREADS, smem, ctx->stream>>>(a, tmpl, 0, (int)nF, nFpad, table);
        else
            scan_contract_kernel<E_TC><<<nAct, E8_THREADS, smem, ctx->stream>>>(a, tmpl, 0, (int)nF, nFpad, table);
        scan_classes_epilogue_kernel<<<nAct, 256, 0, ctx->stream>>>(table, nAct, nK, nR, nT, dwr, dwt, dWC, dWR, dWT, dB);
        span_end(ctx);
        ctx->launches += 2;
        THB_CUDA(ctx, cudaGetLastError());
        THB_CUDA(ctx, cudaMemcpyAsync(wC + (size_t)l0 * nK, dWC, sizeof(float) * (size_t)nAct * nK, cudaMemcpyDeviceToHost, ctx->stream));
        THB_CUDA(ctx, cudaMemcpyAsync(base + l0, dB, sizeof(float) * nAct, cudaMemcpyDeviceToHost, ctx->stream));
        for (int k = 0; k < nK; ++k) {
            THB_CUDA(ctx, cudaMemcpyAsync(wR + ((size_t)k * nImg + l0) * nR, dWR + (size_t)k * nAct * nR, sizeof(float) * (size_t)nAct * nR, cudaMemcpyDeviceToHost, ctx->stream));
            THB_CUDA(ctx, cudaMemcpyAsync(wT + ((size_t)k * nImg + l0) * nT, dWT + (size_t)k * nAct * nT, sizeof(float) * (size_t)nAct * nT, cudaMemcpyDeviceToHost, ctx->stream));
        }
        THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the scratch table and outputs are reused by the next chunk
    }
    return THB_OK;
}

// ------------------------------------------------------------------------------------------------
int thb_reco_alloc(thb_ctx* ctx, int slot, int vdimPad)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || vdimPad <= 0 || (vdimPad & 1)) return set_error(ctx, THB_E_ARG, "reco_alloc: bad slot/vdim");
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    Accum& a = ctx->accs[slot];
    if (a.vdim != vdimPad) {
        cudaFree(a.d);
        a = Accum();
        const size_t n = ctx->mode2D ? 2 * (size_t)(vdimPad / 2 + 1) * vdimPad : vol_elems(vdimPad);   // MODE_2D: two planes
        THB_CUDA(ctx, cudaMalloc(&a.d, n * sizeof(float4)));
        a.vdim = vdimPad;
        a.nVox = n;
    }
    return thb_reco_reset(ctx, slot);
}

int thb_reco_reset(thb_ctx* ctx, int slot)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || !ctx->accs[slot].d) return set_error(ctx, THB_E_STATE, "reco_reset: slot %d not allocated", slot);
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaMemsetAsync(ctx->accs[slot].d, 0, ctx->accs[slot].nVox * sizeof(float4), ctx->stream));
    THB_CUDA(ctx, cudaMemsetAsync(ctx->dO + 3 * slot, 0, 3 * sizeof(double), ctx->stream));
    THB_CUDA(ctx, cudaMemsetAsync(ctx->dCounter + slot, 0, sizeof(int), ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

}  // extern "C"
namespace thb {
// every image of an insert goes to the accumulator of its slot: all of them must be allocated, with one common edge
// that holds the M pixel list (a null accumulator would be a sticky illegal-address error inside the kernel)
int check_insert_slots(thb_ctx* ctx, int nImg, const int* imgIdx, int imgBase, const char* who)
{
    int vdim = 0;
    for (int i = 0; i < THB_MAX_SLOTS; ++i)
        if (ctx->accs[i].d) {
            if (vdim && ctx->accs[i].vdim != vdim) return set_error(ctx, THB_E_STATE, "%s: accumulators of different size", who);
            vdim = ctx->accs[i].vdim;
        }
    if (!vdim) return set_error(ctx, THB_E_STATE, "%s: no accumulator allocated (thb_reco_alloc)", who);
    const std::vector<int>& hs = ctx->stackM.hslot;
    for (int l = 0; l < nImg; ++l) {
        const size_t img = imgIdx ? (size_t)imgIdx[l] : (size_t)imgBase + l;
        if (img >= hs.size()) return set_error(ctx, THB_E_ARG, "%s: image %zu outside the M stack", who, img);
        const int s = hs[img];
        if (s < 0 || s >= THB_MAX_SLOTS || !ctx->accs[s].d)
            return set_error(ctx, THB_E_STATE, "%s: image %zu uses slot %d, which has no accumulator (thb_reco_alloc)", who, img, s);
    }
    return THB_OK;
}
}  // namespace thb
extern "C" {

static int insert_impl(thb_ctx* ctx, int nImg, const int* imgIdx, int mReco, const float* w, const double* offS, const int* nc,
                       const double* nr, const double* nt, const int* nDraw = nullptr, const double* nd = nullptr,
                       const float* ctfAttr = nullptr, float pixelSize = 0.f)
{
    if (!ctx) return THB_E_ARG;
    if (nImg <= 0 || mReco <= 0 || !w || !nr || !nt) return set_error(ctx, THB_E_ARG, "insert: bad arguments");
    if (nc && !ctx->mode2D) return set_error(ctx, THB_E_STATE, "insert_classes: per-draw classes are a MODE_2D feature (thb_set_mode)");
    if (nc)
        for (size_t i = 0; i < (size_t)nImg * mReco; ++i)
            if (nc[i] < 0 || nc[i] >= THB_MAX_SLOTS || !ctx->accs[nc[i]].d) return set_error(ctx, THB_E_ARG, "insert_classes: nc[%zu] = %d has no accumulator", i, nc[i]);
    if (!ctx->pixM) return set_error(ctx, THB_E_STATE, "insert: M pixel list not set");
    if (!ctx->stackM.dat) return set_error(ctx, THB_E_STATE, "insert: M stack not uploaded");
    if (nDraw)
        for (int i = 0; i < nImg; ++i)
            if (nDraw[i] < 0) return set_error(ctx, THB_E_ARG, "insert_counts: nDraw[%d] = %d is negative", i, nDraw[i]);   // > mReco: clipped
    int vdim = 0;
    for (int i = 0; i < THB_MAX_SLOTS; ++i)
        if (ctx->accs[i].d) {
            if (vdim && ctx->accs[i].vdim != vdim) return set_error(ctx, THB_E_STATE, "insert: accumulators of different size");
            vdim = ctx->accs[i].vdim;
        }
    if (!vdim) return set_error(ctx, THB_E_STATE, "insert: no accumulator allocated (thb_reco_alloc)");
    if (imgIdx)
        for (int i = 0; i < nImg; ++i)
            if (imgIdx[i] < 0 || imgIdx[i] >= ctx->stackM.nImg) return set_error(ctx, THB_E_ARG, "insert: imgIdx[%d] outside the stack", i);
    if (!imgIdx && nImg > ctx->stackM.nImg) return set_error(ctx, THB_E_ARG, "insert: nImg exceeds the stack");
    if (!nc) {
        int rc = check_insert_slots(ctx, nImg, imgIdx, 0, "insert");
        if (rc) return rc;
    }
    THB_CUDA(ctx, cudaSetDevice(ctx->device));

    const int qc = ctx->mode2D ? 2 : 4;      // MODE_2D: nr[nImg][mReco][2] = (cos, sin), as InsertI2D receives it
    const size_t nq = (size_t)nImg * mReco * qc, ntt = (size_t)nImg * mReco * 2;
    const size_t nnd = nd ? (size_t)nImg * mReco : 0;
    double* din = (double*)scratch(ctx, 0, sizeof(double) * (nq + ntt + 2 * (size_t)nImg + nnd) + sizeof(float) * (8 * (size_t)nImg) + sizeof(int) * (2 * (size_t)nImg + (size_t)nImg * mReco));
    if (!din) return THB_E_CUDA;
    double* dq = din; double* dt = dq + nq; double* doff = dt + ntt; double* dnd = doff + 2 * (size_t)nImg;
    float* dattr = (float*)(dnd + nnd);
    float* dw = dattr + 7 * (size_t)nImg;
    if (nd) {
        THB_CUDA(ctx, cudaMemcpyAsync(dnd, nd, sizeof(double) * nnd, cudaMemcpyHostToDevice, ctx->stream));
        THB_CUDA(ctx, cudaMemcpyAsync(dattr, ctfAttr, sizeof(float) * 7 * (size_t)nImg, cudaMemcpyHostToDevice, ctx->stream));
    }
    int* didx = (int*)(dw + nImg);
    int* dnc = didx + nImg;
    int* dcount = dnc + (size_t)nImg * mReco;
    if (nc) THB_CUDA(ctx, cudaMemcpyAsync(dnc, nc, sizeof(int) * (size_t)nImg * mReco, cudaMemcpyHostToDevice, ctx->stream));
    if (nDraw) THB_CUDA(ctx, cudaMemcpyAsync(dcount, nDraw, sizeof(int) * (size_t)nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dq, nr, sizeof(double) * nq, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dt, nt, sizeof(double) * ntt, cudaMemcpyHostToDevice, ctx->stream));
    if (offS) THB_CUDA(ctx, cudaMemcpyAsync(doff, offS, sizeof(double) * 2 * (size_t)nImg, cudaMemcpyHostToDevice, ctx->stream));
    THB_CUDA(ctx, cudaMemcpyAsync(dw, w, sizeof(float) * nImg, cudaMemcpyHostToDevice, ctx->stream));
    if (imgIdx) THB_CUDA(ctx, cudaMemcpyAsync(didx, imgIdx, sizeof(int) * (size_t)nImg, cudaMemcpyHostToDevice, ctx->stream));

    InsertArgs a;
    memset(&a, 0, sizeof(a));
    a.acc = acc_table(ctx);
    a.vdim = vdim;
    a.dat = ctx->stackM.dat; a.ctf = ctx->stackM.ctf; a.slotOfImg = ctx->stackM.slot;
    a.pix = ctx->pixM; a.P = ctx->nPxlM; a.N = ctx->NM;
    a.nImg = nImg; a.imgIdx = imgIdx ? didx : nullptr; a.imgBase = 0;
    a.mReco = mReco; a.w = dw; a.wAll = 0.f; a.offS = offS ? doff : nullptr;
    a.nr = View3{dq, (long long)mReco * qc, qc, 1};
    a.nt = View3{dt, (long long)mReco * 2, 2, 1};
    a.mode2D = ctx->mode2D;
    a.drawC = nc ? dnc : nullptr;
    a.drawCount = nDraw ? dcount : nullptr;
    if (nd) {
        a.nd = View3{dnd, (long long)mReco, 1, 0};
        a.ctfAttr = dattr;
        a.pixelSize = pixelSize;
    }
    int rc = launch_insert(ctx, a, imgIdx);
    if (rc) return rc;
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

int thb_insert(thb_ctx* ctx, int nImg, const int* imgIdx, int mReco, const float* w, const double* offS, const double* nr,
               const double* nt)
{
    return insert_impl(ctx, nImg, imgIdx, mReco, w, offS, nullptr, nr, nt);
}

int thb_insert_counts(thb_ctx* ctx, int nImg, const int* imgIdx, int mReco, const float* w, const double* offS, const int* nDraw,
                      const double* nr, const double* nt)
{
    if (ctx && !nDraw) return set_error(ctx, THB_E_ARG, "insert_counts: nDraw == NULL");
    return insert_impl(ctx, nImg, imgIdx, mReco, w, offS, nullptr, nr, nt, nDraw);
}

// CTF search: every draw carries its own defocus factor, its CTF is computed on the fly (src/Optimiser.cpp:7171-7215)
int thb_insert_ctf(thb_ctx* ctx, int nImg, const int* imgIdx, int mReco, const float* w, const double* offS, const double* nr,
                   const double* nt, const double* nd, const float* ctfAttr, float pixelSize)
{
    if (ctx && (!nd || !ctfAttr || !(pixelSize > 0))) return set_error(ctx, THB_E_ARG, "insert_ctf: nd / ctfAttr NULL or pixelSize <= 0");
    return insert_impl(ctx, nImg, imgIdx, mReco, w, offS, nullptr, nr, nt, nullptr, nd, ctfAttr, pixelSize);
}

int thb_insert_classes(thb_ctx* ctx, int nImg, const int* imgIdx, int mReco, const float* w, const double* offS, const int* nc,
                       const double* nr, const double* nt)
{
    if (ctx && !nc) return set_error(ctx, THB_E_ARG, "insert_classes: nc == NULL");
    return insert_impl(ctx, nImg, imgIdx, mReco, w, offS, nc, nr, nt);
}

// MODE_3D (default) <-> MODE_2D.  The projector references and the accumulators have different shapes in the two modes, so
// switching drops them; pixel lists and resident stacks are mode-independent and stay.
int thb_set_mode(thb_ctx* ctx, int mode)
{
    if (!ctx) return THB_E_ARG;
    if (mode != THB_MODE_3D && mode != THB_MODE_2D) return set_error(ctx, THB_E_ARG, "set_mode: mode %d", mode);
    const int m2 = mode == THB_MODE_2D;
    if (m2 == ctx->mode2D) return THB_OK;
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < THB_MAX_SLOTS; ++i) {
        cudaFree(ctx->vols[i].d);
        cudaFree(ctx->vols[i].quad);
        ctx->vols[i] = Volume3();
        cudaFree(ctx->accs[i].d);
        ctx->accs[i] = Accum();
    }
    ctx->mode2D = m2;
    return THB_OK;
}

int thb_get_mode(const thb_ctx* ctx) { return ctx && ctx->mode2D ? THB_MODE_2D : THB_MODE_3D; }

int thb_reco_download(thb_ctx* ctx, int slot, float* F, float* T, double* O, int* counter, int normalise)
{
    if (!ctx) return THB_E_ARG;
    if (slot < 0 || slot >= THB_MAX_SLOTS || !ctx->accs[slot].d) return set_error(ctx, THB_E_STATE, "reco_download: slot %d not allocated", slot);
    THB_CUDA(ctx, cudaSetDevice(ctx->device));
    const Accum& a0 = ctx->accs[slot];
    struct { float4* d; size_t nVox; } a = {a0.d, ctx->mode2D ? a0.nVox / 2 : a0.nVox};     // MODE_2D: plane 0 is the image
    if (F || T) {
        float2* dF = nullptr; float* dT = nullptr;
        if (F) { dF = (float2*)scratch(ctx, 2, a.nVox * sizeof(float2)); if (!dF) return THB_E_CUDA; }
        if (T) { dT = (float*)scratch(ctx, 3, a.nVox * sizeof(float)); if (!dT) return THB_E_CUDA; }
        span_begin(ctx, KF_PACK);
        unpack_acc_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(a.d, a.nVox, dF, dT, normalise);
        span_end(ctx);
        ctx->launches++;
        THB_CUDA(ctx, cudaGetLastError());
        if (F) THB_CUDA(ctx, cudaMemcpyAsync(F, dF, a.nVox * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
        if (T) THB_CUDA(ctx, cudaMemcpyAsync(T, dT, a.nVox * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (O) THB_CUDA(ctx, cudaMemcpyAsync(O, ctx->dO + 3 * slot, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (counter) THB_CUDA(ctx, cudaMemcpyAsync(counter, ctx->dCounter + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    THB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return THB_OK;
}

}  // extern "C"
