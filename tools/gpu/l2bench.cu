// l2bench.cu - how fast are (a) random 64-byte cell gathers and (b) random 16-byte vector reductions as a function of the
// FOOTPRINT they fall into (L2-resident slab vs the whole volume in HBM), and how do reductions scale with the number of
// CTAs and with the way the lanes of a warp are laid over the cell (one lane per corner pair vs one lane per sample)?
// Decides whether ordering the work of the E / M kernels by z-slab of the volume pays.  Development aid, not product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2bench l2bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
struct __align__(32) Q { float v[8]; };
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ Q ldq(const Q* p)
{
    Q q;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(q.v[0]), "=f"(q.v[1]), "=f"(q.v[2]), "=f"(q.v[3]), "=f"(q.v[4]), "=f"(q.v[5]), "=f"(q.v[6]), "=f"(q.v[7]) : "l"(p));
    return q;
}
__device__ __forceinline__ void red4(float4* p, float a, float b, float c)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(0.0f) : "memory");
}
__device__ __forceinline__ void red2(float2* p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
constexpr int STEPS = 256;

// random 64-byte cells inside a footprint of nCells cells
__global__ void __launch_bounds__(256, 2) k_gather(const Q* __restrict__ vol, unsigned nCells, float* out)
{
    const unsigned h = hash32((blockIdx.x * 256u + threadIdx.x) * 2654435761u + 99u);
    float acc = 0.f;
#pragma unroll 4
    for (int st = 0; st < STEPS; ++st) {
        const unsigned g = hash32(h + st * 40503u);
        const Q* p = vol + 2 * (size_t)(g % nCells);
        const Q a = ldq(p), b = ldq(p + 1);
        acc += a.v[0] + a.v[7] + b.v[0] + b.v[7];
    }
    if (acc == 12345.f) out[0] = acc;
}

// the insert kernel's pattern: a sample adds into the 8 corners of a cell of a float4 volume [nz][ny][nx]: 4 rows x 2 x-adjacent
// MODE 0: one lane per sample, 8 red.v4 in sequence (today's kernel)
// MODE 1: two lanes per sample (x0 and x0+1 of the same row = one 32-byte sector when x0 is even), 4 red.v4 each
// MODE 2: one lane per sample, 3 words as red.v2 + red (12-byte voxels in a float4 slot)
// MODE 3: one lane per sample, 8 red.v4 to fully random float4 (no cell structure)
template <int MODE>
__global__ void __launch_bounds__(256, 2) k_red(float4* __restrict__ acc, int nx, int ny, int nz, float* out)
{
    const unsigned tid = blockIdx.x * 256u + threadIdx.x;
    const unsigned sample = MODE == 1 ? tid >> 1 : tid;
    const unsigned h = hash32(sample * 2654435761u + 7u);
    const size_t sy = nx, sz = (size_t)nx * ny;
    for (int st = 0; st < STEPS / 4; ++st) {
        const unsigned g = hash32(h + st * 40503u), g2 = hash32(g + 17u);
        const int x = g % (nx - 1), y = (g >> 10) % (ny - 1), z = g2 % (nz - 1);
        float4* p = acc + (size_t)z * sz + (size_t)y * sy + x;
        if (MODE == 0) {
            red4(p, 1.f, 2.f, 3.f); red4(p + 1, 1.f, 2.f, 3.f);
            red4(p + sy, 1.f, 2.f, 3.f); red4(p + sy + 1, 1.f, 2.f, 3.f);
            red4(p + sz, 1.f, 2.f, 3.f); red4(p + sz + 1, 1.f, 2.f, 3.f);
            red4(p + sz + sy, 1.f, 2.f, 3.f); red4(p + sz + sy + 1, 1.f, 2.f, 3.f);
        } else if (MODE == 1) {
            float4* q = p + (tid & 1);
            red4(q, 1.f, 2.f, 3.f); red4(q + sy, 1.f, 2.f, 3.f); red4(q + sz, 1.f, 2.f, 3.f); red4(q + sz + sy, 1.f, 2.f, 3.f);
        } else if (MODE == 2) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4* q = p + (c & 1) + ((c >> 1) & 1) * sy + (c >> 2) * sz;
                red2(reinterpret_cast<float2*>(q), 1.f, 2.f);
                atomicAdd(reinterpret_cast<float*>(q) + 2, 3.f);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const unsigned r = hash32(g2 + c * 977u);
                red4(acc + (size_t)(r % (unsigned)(sz * nz)), 1.f, 2.f, 3.f);
            }
        }
    }
    if (out && tid == 0xffffffffu) out[0] = 0.f;
}

// shared-memory variant: 24 fp32 atomic adds per sample into a 48 KB tile, then nothing (cost of the privatised scatter)
__global__ void __launch_bounds__(256, 2) k_smem(float* out)
{
    extern __shared__ float tile[];
    const int W = 12288;   // 48 KB of floats
    for (int i = threadIdx.x; i < W; i += 256) tile[i] = 0.f;
    __syncthreads();
    const unsigned h = hash32((blockIdx.x * 256u + threadIdx.x) * 2654435761u + 3u);
    for (int st = 0; st < STEPS; ++st) {
        const unsigned g = hash32(h + st * 40503u);
        const int base = g % (W - 3 * 600);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int o = base + 3 * ((c & 1) + ((c >> 1) & 1) * 17 + (c >> 2) * 289);
            atomicAdd(&tile[o], 1.f); atomicAdd(&tile[o + 1], 2.f); atomicAdd(&tile[o + 2], 3.f);
        }
    }
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < W; i += 256) s += tile[i];
    if (s == 12345.f) out[0] = s;
}

template <typename F>
static float time_ms(F f, int reps = 3)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best;
    }
    return best;
}

int main()
{
    float* out;
    CK(cudaMalloc(&out, 64));
    const size_t maxBytes = (size_t)4 << 30;
    void* buf;
    CK(cudaMalloc(&buf, maxBytes));
    CK(cudaMemset(buf, 0, maxBytes));
    const int grid = 148 * 8;
    printf("== random 64-byte cell gathers, %d CTAs x 256 threads x %d steps\n", grid, STEPS);
    printf("footprint_MB  ms  Gcells/s  GB/s\n");
    for (int mb : {8, 16, 24, 32, 48, 64, 96, 128, 256, 1024, 4096}) {
        const unsigned nCells = (unsigned)(((size_t)mb << 20) / 64);
        const float ms = time_ms([&] { k_gather<<<grid, 256>>>((const Q*)buf, nCells, out); });
        const double n = (double)grid * 256 * STEPS;
        printf("%5d  %.3f  %.1f  %.0f\n", mb, ms, n / ms / 1e6, n * 64 / ms / 1e6);
    }
    printf("== cell-pattern reductions into a float4 volume (8 corners per sample); Gsamples/s (x8 red.v4 each)\n");
    printf("footprint_MB  grid  mode0_lane/sample  mode1_lane-pair  mode2_v2+scalar  mode3_random\n");
    for (int mb : {16, 32, 64, 128, 1024, 2048}) {
        // volume nx = 257, ny = 512, nz from the footprint
        const int nx = 257, ny = 512;
        const int nz = (int)(((size_t)mb << 20) / ((size_t)nx * ny * 16));
        for (int g : {148, 296, 148 * 8}) {
            float r[4];
            const double n = (double)g * 256 * (STEPS / 4);
            r[0] = time_ms([&] { k_red<0><<<g, 256>>>((float4*)buf, nx, ny, nz, out); });
            r[1] = time_ms([&] { k_red<1><<<2 * g, 256>>>((float4*)buf, nx, ny, nz, out); });
            r[2] = time_ms([&] { k_red<2><<<g, 256>>>((float4*)buf, nx, ny, nz, out); });
            r[3] = time_ms([&] { k_red<3><<<g, 256>>>((float4*)buf, nx, ny, nz, out); });
            printf("%5d  %5d  %.2f  %.2f  %.2f  %.2f\n", mb, g, n / r[0] / 1e6, n / r[1] / 1e6, n / r[2] / 1e6, n / r[3] / 1e6);
        }
    }
    printf("== scaling of mode 0 with the number of CTAs (footprint 32 MB and 2048 MB)\n");
    for (int mb : {32, 2048}) {
        const int nx = 257, ny = 512;
        const int nz = (int)(((size_t)mb << 20) / ((size_t)nx * ny * 16));
        for (int g : {16, 32, 64, 148, 296, 592}) {
            const double n = (double)g * 256 * (STEPS / 4);
            const float ms = time_ms([&] { k_red<0><<<g, 256>>>((float4*)buf, nx, ny, nz, out); });
            printf("%5d MB  grid %4d  %.2f Gsamples/s\n", mb, g, n / ms / 1e6);
        }
    }
    CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
    {
        const float ms = time_ms([&] { k_smem<<<296, 256, 49152>>>(out); });
        const double n = 296.0 * 256 * STEPS;
        printf("== shared-memory scatter (24 fp32 atomics per sample): %.2f Gsamples/s\n", n / ms / 1e6);
    }
    return 0;
}
