// gatherbench.cu - micro-benchmark of the 8-tap trilinear gather through four paths of a B200 SM:
//   ldg   : __ldg from a linear half-complex volume in HBM (x fastest)
//   tex   : tex3D<float2> point sampling from a 3D cudaArray (8 fetches per sample)
//   lds   : shared-memory box (the addresses of the same pattern folded into a 22^3 box)
// Pattern: one warp = 32 orientation samples of one pixel; the cell of lane l is base + a random offset
// in [-s, s]^3 (the spread of the orientation cloud at that pixel); consecutive steps walk a slice.
// Prints cycles per warp-step per SM and samples/s for each spread.   Development aid, not product.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int N = 512, NC = 260;          // volume n x n x pitch
constexpr int STEPS = 512;
constexpr int THREADS = 128;

__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// cell of (block, warp, lane, step): slices through the volume centre region, lanes scattered by s
__device__ __forceinline__ void cell_of(int blk, int warp, int lane, int step, int s, int& x, int& y, int& z)
{
    const unsigned h = hash32(blk * 131071u + warp * 8191u + 17u);
    // a slice direction per warp; walk 2 voxels per step along a line, wrapping every 64 steps to the next line
    const int bx = 20 + (int)(h % 150), by = 100 + (int)((h >> 8) % 300), bz = 100 + (int)((h >> 17) % 300);
    const int u = step & 63, v = step >> 6;
    const unsigned r = hash32(h + lane * 2654435761u);
    const int ox = s ? (int)(r % (2 * s + 1)) - s : 0, oy = s ? (int)((r >> 8) % (2 * s + 1)) - s : 0,
              oz = s ? (int)((r >> 16) % (2 * s + 1)) - s : 0;
    x = bx + u + ox + (v & 1);
    y = by + (u >> 1) + 2 * v + oy;
    z = bz + (u >> 2) + v + oz;
}

__global__ void k_ldg(const float2* __restrict__ vol, int s, float* out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ax = 0.f, ay = 0.f;
    for (int st = 0; st < STEPS; ++st) {
        int x, y, z;
        cell_of(blockIdx.x, warp, lane, st, s, x, y, z);
        const float2* p = vol + ((size_t)z * N + y) * NC + x;
        const size_t sy = NC, sz = (size_t)N * NC;
        const float2 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + sy), v3 = __ldg(p + sy + 1);
        const float2 v4 = __ldg(p + sz), v5 = __ldg(p + sz + 1), v6 = __ldg(p + sz + sy), v7 = __ldg(p + sz + sy + 1);
        ax += v0.x + v1.x + v2.x + v3.x + v4.x + v5.x + v6.x + v7.x;
        ay += v0.y + v1.y + v2.y + v3.y + v4.y + v5.y + v6.y + v7.y;
    }
    if (ax == 12345.f) out[0] = ay;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = ax + ay;
}

__global__ void k_tex(cudaTextureObject_t tex, int s, float* out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ax = 0.f, ay = 0.f;
    for (int st = 0; st < STEPS; ++st) {
        int x, y, z;
        cell_of(blockIdx.x, warp, lane, st, s, x, y, z);
        const float fx = x + 0.5f, fy = y + 0.5f, fz = z + 0.5f;
        const float2 v0 = tex3D<float2>(tex, fx, fy, fz), v1 = tex3D<float2>(tex, fx + 1, fy, fz);
        const float2 v2 = tex3D<float2>(tex, fx, fy + 1, fz), v3 = tex3D<float2>(tex, fx + 1, fy + 1, fz);
        const float2 v4 = tex3D<float2>(tex, fx, fy, fz + 1), v5 = tex3D<float2>(tex, fx + 1, fy, fz + 1);
        const float2 v6 = tex3D<float2>(tex, fx, fy + 1, fz + 1), v7 = tex3D<float2>(tex, fx + 1, fy + 1, fz + 1);
        ax += v0.x + v1.x + v2.x + v3.x + v4.x + v5.x + v6.x + v7.x;
        ay += v0.y + v1.y + v2.y + v3.y + v4.y + v5.y + v6.y + v7.y;
    }
    if (ax == 12345.f) out[0] = ay;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = ax + ay;
}

constexpr int BX = 24, BY = 22, BZ = 22;
__global__ void k_lds(const float2* __restrict__ vol, int s, float* out)
{
    __shared__ float2 box[BX * BY * BZ];   // 93 KB > 48 KB static limit -> use dynamic below
    (void)vol; (void)s; (void)out;
}

__global__ void k_lds_dyn(const float2* __restrict__ vol, int s, float* out)
{
    extern __shared__ float2 box[];
    for (int i = threadIdx.x; i < BX * BY * BZ; i += blockDim.x) box[i] = vol[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ax = 0.f, ay = 0.f;
    for (int st = 0; st < STEPS; ++st) {
        int x, y, z;
        cell_of(blockIdx.x, warp, lane, st, s, x, y, z);
        // fold the walk into the box, keep the lane scatter
        x = (x & 7) + (x >> 8) % 3 + 6; y = (y & 3) + 8 + (y >> 9); z = (z & 3) + 8 + (z >> 9);
        x = min(max(x, 0), BX - 2); y = min(max(y, 0), BY - 2); z = min(max(z, 0), BZ - 2);
        const float2* p = box + (z * BY + y) * BX + x;
        const int sy = BX, sz = BX * BY;
        const float2 v0 = p[0], v1 = p[1], v2 = p[sy], v3 = p[sy + 1];
        const float2 v4 = p[sz], v5 = p[sz + 1], v6 = p[sz + sy], v7 = p[sz + sy + 1];
        ax += v0.x + v1.x + v2.x + v3.x + v4.x + v5.x + v6.x + v7.x;
        ay += v0.y + v1.y + v2.y + v3.y + v4.y + v5.y + v6.y + v7.y;
    }
    if (ax == 12345.f) out[0] = ay;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = ax + ay;
}

// pairs layout: V4[i] = (A[i], A[i+1]) as float4, 16 B per voxel -> 4 x LDG.128 per sample
__global__ void k_ldg4(const float4* __restrict__ vol, int s, float* out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ax = 0.f, ay = 0.f;
    for (int st = 0; st < STEPS; ++st) {
        int x, y, z;
        cell_of(blockIdx.x, warp, lane, st, s, x, y, z);
        const float4* p = vol + ((size_t)z * N + y) * NC + x;
        const size_t sy = NC, sz = (size_t)N * NC;
        const float4 v0 = __ldg(p), v1 = __ldg(p + sy), v2 = __ldg(p + sz), v3 = __ldg(p + sz + sy);
        ax += v0.x + v0.z + v1.x + v1.z + v2.x + v2.z + v3.x + v3.z;
        ay += v0.y + v0.w + v1.y + v1.w + v2.y + v2.w + v3.y + v3.w;
    }
    if (ax == 12345.f) out[0] = ay;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = ax + ay;
}

struct __align__(32) f8 { float v[8]; };
__device__ __forceinline__ f8 ldg256(const f8* p)
{
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
// quad layout: V8[i] = (A[x,y], A[x+1,y], A[x,y+1], A[x+1,y+1]), 32 B per voxel -> 2 x LDG.256 per sample
__global__ void k_ldg8(const f8* __restrict__ vol, int s, float* out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ax = 0.f, ay = 0.f;
    for (int st = 0; st < STEPS; ++st) {
        int x, y, z;
        cell_of(blockIdx.x, warp, lane, st, s, x, y, z);
        const f8* p = vol + ((size_t)z * N + y) * NC + x;
        const size_t sz = (size_t)N * NC;
        const f8 a = ldg256(p), b = ldg256(p + sz);
        ax += a.v[0] + a.v[2] + a.v[4] + a.v[6] + b.v[0] + b.v[2] + b.v[4] + b.v[6];
        ay += a.v[1] + a.v[3] + a.v[5] + a.v[7] + b.v[1] + b.v[3] + b.v[5] + b.v[7];
    }
    if (ax == 12345.f) out[0] = ay;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = ax + ay;
}
// the same through two 128-bit loads per quad (4 x LDG.128 per sample)
__global__ void k_ldg8b(const float4* __restrict__ vol, int s, float* out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ax = 0.f, ay = 0.f;
    for (int st = 0; st < STEPS; ++st) {
        int x, y, z;
        cell_of(blockIdx.x, warp, lane, st, s, x, y, z);
        const float4* p = vol + 2 * (((size_t)z * N + y) * NC + x);
        const size_t sz = 2 * (size_t)N * NC;
        const float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + sz), v3 = __ldg(p + sz + 1);
        ax += v0.x + v0.z + v1.x + v1.z + v2.x + v2.z + v3.x + v3.z;
        ay += v0.y + v0.w + v1.y + v1.w + v2.y + v2.w + v3.y + v3.w;
    }
    if (ax == 12345.f) out[0] = ay;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = ax + ay;
}
// pairs in the shared-memory box: 4 x LDS.128 per sample (box of 16 B elements, 12 x 22 x 22)
__global__ void k_lds4(const float2* __restrict__ vol, int s, float* out)
{
    extern __shared__ float4 box4[];
    constexpr int BX4 = 12;
    for (int i = threadIdx.x; i < BX4 * BY * BZ; i += blockDim.x) box4[i] = make_float4(vol[i].x, vol[i].y, vol[i + 1].x, vol[i + 1].y);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ax = 0.f, ay = 0.f;
    for (int st = 0; st < STEPS; ++st) {
        int x, y, z;
        cell_of(blockIdx.x, warp, lane, st, s, x, y, z);
        x = (x & 3) + (x >> 8) % 3 + 2; y = (y & 3) + 8 + (y >> 9); z = (z & 3) + 8 + (z >> 9);
        x = min(max(x, 0), BX4 - 1); y = min(max(y, 0), BY - 2); z = min(max(z, 0), BZ - 2);
        const float4* p = box4 + (z * BY + y) * BX4 + x;
        const int sy = BX4, sz = BX4 * BY;
        const float4 v0 = p[0], v1 = p[sy], v2 = p[sz], v3 = p[sz + sy];
        ax += v0.x + v0.z + v1.x + v1.z + v2.x + v2.z + v3.x + v3.z;
        ay += v0.y + v0.w + v1.y + v1.w + v2.y + v2.w + v3.y + v3.w;
    }
    if (ax == 12345.f) out[0] = ay;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = ax + ay;
}

int main()
{
    const size_t elems = (size_t)N * N * NC;
    std::vector<float2> h(elems);
    for (size_t i = 0; i < elems; ++i) h[i] = make_float2((float)(i % 97) * 0.01f, (float)(i % 89) * 0.02f);
    float2* dvol;
    CK(cudaMalloc(&dvol, elems * sizeof(float2)));
    CK(cudaMemcpy(dvol, h.data(), elems * sizeof(float2), cudaMemcpyHostToDevice));
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float2>();
    cudaArray_t arr;
    CK(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(NC, N, N)));
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr(h.data(), NC * sizeof(float2), NC, N);
    cp.dstArray = arr;
    cp.extent = make_cudaExtent(NC, N, N);
    cp.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&cp));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    float* dout;
    CK(cudaMalloc(&dout, 16));
    CK(cudaFuncSetAttribute(k_lds_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, BX * BY * BZ * 8));
    CK(cudaFuncSetAttribute(k_lds4, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * BY * BZ * 16));
    float4* dvol4;   // pairs
    CK(cudaMalloc(&dvol4, elems * sizeof(float4)));
    CK(cudaMemset(dvol4, 0, elems * sizeof(float4)));
    f8* dvol8;       // quads
    CK(cudaMalloc(&dvol8, elems * sizeof(f8)));
    CK(cudaMemset(dvol8, 0, elems * sizeof(f8)));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int spreads[] = {0, 1, 2, 3, 5, 8};
    printf("path spread  ms   Gsamples/s  cyc/sample/SM (at 1.9 GHz)\n");
    for (int s : spreads) {
        for (int path = 0; path < 7; ++path) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(a);
                if (path == 0) k_ldg<<<blocks, THREADS>>>(dvol, s, dout);
                if (path == 1) k_tex<<<blocks, THREADS>>>(tex, s, dout);
                if (path == 2) k_lds_dyn<<<blocks, THREADS, BX * BY * BZ * 8>>>(dvol, s, dout);
                if (path == 3) k_ldg4<<<blocks, THREADS>>>(dvol4, s, dout);
                if (path == 4) k_ldg8<<<blocks, THREADS>>>(dvol8, s, dout);
                if (path == 5) k_ldg8b<<<blocks, THREADS>>>((const float4*)dvol8, s, dout);
                if (path == 6) k_lds4<<<blocks, THREADS, 12 * BY * BZ * 16>>>(dvol, s, dout);
                cudaEventRecord(b);
                CK(cudaEventSynchronize(b));
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                best = ms < best ? ms : best;
            }
            const double samples = (double)blocks * THREADS * STEPS;
            printf("%s  %d  %8.3f  %8.1f  %6.2f\n", (const char*[]){"ldg", "tex", "lds", "ldg4", "ldg8", "ldg8b", "lds4"}[path], s, best, samples / best / 1e6,
                   best * 1e-3 * 1.9e9 * sms / samples);
        }
    }
    return 0;
}
