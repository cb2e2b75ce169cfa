// drambench.cu - how much DRAM bandwidth does a gather of 64-byte cells get, as a function of WHERE the 32 lanes of a warp
// point?  Cells are 64-byte elements of a 4.3 GB "cell volume" (512 x 512 x 256 elements, 4x4x4 bricks, as thb_expect3.cuh).
//   rot  : lanes = 32 rotations of one pixel: cells scattered in a cube of +-S voxels around the warp's walking base
//          (the default E kernel in the wide-cloud regime)
//   pix  : lanes = an 8 x 4 patch of neighbouring pixels of ONE rotation: cells 2 voxels apart on a tilted plane
//          (a "pixels on the lanes" organisation of the same work)
// No cell is touched twice (no cache reuse): both patterns are DRAM-bound; the difference is DRAM page / sector locality.
// Development aid, not product.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o drambench drambench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int N = 512, H = 256, LB = 2, STEPS = 256;
struct __align__(32) Q { float v[8]; };
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ size_t cell_index(int x, int y, int z)
{
    const int m = (1 << LB) - 1;
    const size_t brick = ((size_t)(z >> LB) * (N >> LB) + (y >> LB)) * (H >> LB) + (x >> LB);
    return (brick << (3 * LB)) | (size_t)((((z & m) << LB) | (y & m)) << LB | (x & m));
}
__device__ __forceinline__ Q ldq(const Q* p)
{
    Q q;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(q.v[0]), "=f"(q.v[1]), "=f"(q.v[2]), "=f"(q.v[3]), "=f"(q.v[4]), "=f"(q.v[5]), "=f"(q.v[6]), "=f"(q.v[7]) : "l"(p));
    return q;
}
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(const Q* __restrict__ vol, int S, float* out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned h = hash32(blockIdx.x * 131071u + warp * 8191u + 17u);
    float acc = 0.f;
#pragma unroll 2
    for (int st = 0; st < STEPS; ++st) {
        const unsigned g = hash32(h + st * 40503u);
        int x, y, z;
        if (MODE == 0) {          // rot: random base per step (a pixel of the slice), lanes scattered +-S
            const unsigned r = hash32(g + lane * 2654435761u);
            x = 24 + (int)(g % 200) + (int)(r % (2 * S + 1)) - S; y = 30 + (int)((g >> 8) % 440) + (int)((r >> 8) % (2 * S + 1)) - S;
            z = 30 + (int)((g >> 17) % 440) + (int)((r >> 16) % (2 * S + 1)) - S;
            x = max(x, 0);
        } else {                  // pix: random base per step (a rotation x patch), lanes = 8 x 4 pixels, 2 voxels apart, tilted plane
            const int lx = lane & 7, ly = lane >> 3;
            x = 4 + (int)(g % 230) + 2 * lx; y = 4 + (int)((g >> 8) % 490) + 2 * ly; z = 4 + (int)((g >> 17) % 490) + ((lx + 2 * ly) >> 1);
        }
        const Q* p = vol + 2 * cell_index(x, y, z);
        const Q a = ldq(p), b = ldq(p + 1);
        acc += a.v[0] + a.v[7] + b.v[0] + b.v[7];
    }
    if (acc == 12345.f) out[0] = acc;
}
int main()
{
    const size_t elems = (size_t)N * N * H;
    Q* vol; float* out;
    CK(cudaMalloc(&vol, elems * 64)); CK(cudaMemset(vol, 0, elems * 64)); CK(cudaMalloc(&out, 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 2 * 8;
    for (int mode = 0; mode < 2; ++mode)
        for (int S : {20, 8}) {
            if (mode == 1 && S == 8) continue;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<blocks, 256>>>(vol, S, out); else k<1><<<blocks, 256>>>(vol, S, out);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double samples = (double)blocks * 256 * STEPS;
                if (rep == 2) printf("%s S=%d: %.2f ms  %.1f G cells/s  %.0f GB/s\n", mode ? "pix" : "rot", S, ms, samples / ms / 1e6, samples * 64 / ms / 1e6);
            }
        }
    return 0;
}
