// thb_comm.cpp - half-map allreduce over NCCL (NVLink 5 / NVSwitch), one rank per GPU.
//
// Replaces Reconstructor::allReduceF/T/O (reference src/Reconstructor.cpp:2350-2520; NCCL twin
// gpu/src/cuthunder.cu:5294-5324, 5903-5985, which creates and destroys communicators per call).
// Here: ONE persistent communicator per context.  The accumulators are interleaved {F.re, F.im, T, 0} for the 16-byte
// reductions of the insert; on the wire only the three live floats travel: every allocated slot is packed into ONE contiguous
// buffer of 3 floats per voxel, reduced by a single ncclAllReduce (one large message instead of one per slot, 25 % fewer bytes
// over NVLink than the padded vectors), and unpacked; O (fp64) and the counters (int32) ride in the same NCCL group.
//
// NCCL is resolved at run time with dlopen so that the library loads on machines without it and
// binds to whichever libnccl.so.2 the host process already carries (torch's bundled copy or the
// system one) instead of pulling in a second copy.
#include <dlfcn.h>
#include <cstring>
#include <cstdio>
#include "thb_context.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclInt32 = 2, ncclFloat32 = 7, ncclFloat64 = 8 };

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& api()
{
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.h) break;
    }
    if (!a.h) return a;
#define LOAD(sym) *(void**)(&a.sym) = dlsym(a.h, "nccl" #sym)
    LOAD(GetUniqueId); LOAD(CommInitRank); LOAD(CommDestroy); LOAD(AllReduce); LOAD(GroupStart); LOAD(GroupEnd);
    LOAD(GetErrorString);
#undef LOAD
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd;
    return a;
}

}  // namespace

namespace thb {

// {F.re, F.im, T, 0} x nVox  <->  3 floats x nVox
__global__ void pack_acc3_kernel(const float4* __restrict__ acc, size_t nVox, float* __restrict__ buf)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nVox; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = acc[i];
        buf[3 * i] = v.x; buf[3 * i + 1] = v.y; buf[3 * i + 2] = v.z;
    }
}
__global__ void unpack_acc3_kernel(const float* __restrict__ buf, size_t nVox, float4* __restrict__ acc)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nVox; i += (size_t)gridDim.x * blockDim.x)
        acc[i] = make_float4(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2], 0.f);
}

void comm_destroy(thb_ctx* ctx)
{
    if (ctx->ncclComm && api().ok) api().CommDestroy((ncclComm_t)ctx->ncclComm);
    ctx->ncclComm = nullptr;
}

int comm_allreduce(thb_ctx* ctx)
{
    if (ctx->nRanks <= 1) return THB_OK;   // single rank: the sum over ranks is the identity
    NcclApi& n = api();
    if (!n.ok || !ctx->ncclComm) return set_error(ctx, THB_E_NCCL, "allreduce: communicator not initialised");
    ncclComm_t comm = (ncclComm_t)ctx->ncclComm;
    size_t total = 0;
    for (int s = 0; s < THB_MAX_SLOTS; ++s)
        if (ctx->accs[s].d) total += ctx->accs[s].nVox;
    if (total * 3 * sizeof(float) > ctx->commBufBytes) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->commBuf);
        ctx->commBuf = nullptr;
        ctx->commBufBytes = 0;
        cudaError_t e = cudaMalloc(&ctx->commBuf, total * 3 * sizeof(float));
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(all-reduce wire buffer)");
        ctx->commBufBytes = total * 3 * sizeof(float);
    }
    float* buf = (float*)ctx->commBuf;
    span_begin(ctx, KF_COMM);
    size_t off = 0;
    for (int s = 0; s < THB_MAX_SLOTS; ++s)
        if (ctx->accs[s].d) {
            pack_acc3_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(ctx->accs[s].d, ctx->accs[s].nVox, buf + 3 * off);
            off += ctx->accs[s].nVox;
        }
    ncclResult_t r = n.GroupStart();
    if (r == 0) r = n.AllReduce(buf, buf, total * 3, ncclFloat32, ncclSum, comm, ctx->stream);
    if (r == 0) r = n.AllReduce(ctx->dO, ctx->dO, 3 * THB_MAX_SLOTS, ncclFloat64, ncclSum, comm, ctx->stream);
    if (r == 0) r = n.AllReduce(ctx->dCounter, ctx->dCounter, THB_MAX_SLOTS, ncclInt32, ncclSum, comm, ctx->stream);
    ncclResult_t r2 = n.GroupEnd();
    off = 0;
    for (int s = 0; s < THB_MAX_SLOTS; ++s)
        if (ctx->accs[s].d) {
            unpack_acc3_kernel<<<ctx->smCount * 8, 256, 0, ctx->stream>>>(buf + 3 * off, ctx->accs[s].nVox, ctx->accs[s].d);
            off += ctx->accs[s].nVox;
        }
    span_end(ctx);
    ctx->commBytesLast = total * 3 * sizeof(float);
    if (r == 0) r = r2;
    if (r != 0) return set_error(ctx, THB_E_NCCL, "ncclAllReduce failed: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
    ctx->launches += 3;
    return THB_OK;
}

}  // namespace thb

extern "C" {

int thb_comm_unique_id(char id[THB_UNIQUE_ID_BYTES])
{
    NcclApi& n = api();
    if (!n.ok || !id) return THB_E_NCCL;
    ncclUniqueId u;
    if (n.GetUniqueId(&u) != 0) return THB_E_NCCL;
    static_assert(sizeof(u) == THB_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &u, sizeof(u));
    return THB_OK;
}

int thb_comm_init(thb_ctx* ctx, int nRanks, int rank, const char id[THB_UNIQUE_ID_BYTES])
{
    if (!ctx) return THB_E_ARG;
    if (nRanks < 1 || rank < 0 || rank >= nRanks) return thb::set_error(ctx, THB_E_ARG, "comm_init: bad rank %d of %d", rank, nRanks);
    thb::comm_destroy(ctx);
    ctx->nRanks = nRanks;
    ctx->rank = rank;
    if (nRanks == 1) return THB_OK;
    NcclApi& n = api();
    if (!n.ok) return thb::set_error(ctx, THB_E_NCCL, "comm_init: libnccl.so.2 not found (%s)", dlerror());
    if (!id) return thb::set_error(ctx, THB_E_ARG, "comm_init: unique id is NULL");
    cudaSetDevice(ctx->device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    ncclResult_t r = n.CommInitRank(&comm, nRanks, u, rank);
    if (r != 0) return thb::set_error(ctx, THB_E_NCCL, "ncclCommInitRank failed: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
    ctx->ncclComm = comm;
    return THB_OK;
}

int thb_allreduce(thb_ctx* ctx)
{
    if (!ctx) return THB_E_ARG;
    cudaSetDevice(ctx->device);
    int rc = thb::comm_allreduce(ctx);
    if (rc) return rc;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return thb::cuda_fail(ctx, e, "allreduce sync");
    return THB_OK;
}

}  // extern "C"
