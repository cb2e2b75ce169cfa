// tmabench.cu - can the TMA engine deliver scattered 64-byte cells into shared memory faster than the LSU data pipe gathers them?
// The default E kernel is bound by the L1/TEX data pipe: one wavefront per 32-byte sector of a scattered LDG.256, two per 64-byte cell
// (DESIGN.md section 4.1).  cp.async.bulk copies bypass that pipe.  This measures, for an L2-resident footprint, cells/s of
//   mode 0: two LDG.256 per cell into registers (what the kernel does)
//   mode 1: one 64-byte cp.async.bulk per cell and lane into shared memory (mbarrier per warp and stage), then 4 x LDS.128
//   mode 2: four 128-bit texture fetches per cell from a linear texture; mode 3: LDG and texture on alternate cells
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gpu/bin/tmabench tools/gpu/tmabench.cu ; run: ./tools/gpu/bin/tmabench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

struct __align__(32) Quad { float v[8]; };
__device__ __forceinline__ Quad ldg_quad(const Quad* p)
{
    Quad q;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(q.v[0]), "=f"(q.v[1]), "=f"(q.v[2]), "=f"(q.v[3]), "=f"(q.v[4]), "=f"(q.v[5]), "=f"(q.v[6]), "=f"(q.v[7]) : "l"(p));
    return q;
}

constexpr int THREADS = 256, STAGES = 4;

__global__ void __launch_bounds__(THREADS, 2) gather_ldg(const Quad* __restrict__ vol, uint32_t nCells, int steps, float* out)
{
    const uint32_t gid = blockIdx.x * THREADS + threadIdx.x;
    float acc = 0.f;
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
        const uint32_t c = hash32(gid * 9781u + s * 6271u) % nCells;
        const Quad a = ldg_quad(vol + 2 * (size_t)c), b = ldg_quad(vol + 2 * (size_t)c + 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += a.v[k] * 0.5f + b.v[k];
    }
    if (acc == 123.456f) out[gid] = acc;
}

__global__ void __launch_bounds__(THREADS, 2) gather_tma(const Quad* __restrict__ vol, uint32_t nCells, int steps, float* out)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    float4 (*cells)[THREADS][4] = reinterpret_cast<float4 (*)[THREADS][4]>(dyn);     // [STAGES]: 64 bytes per lane and stage
    __shared__ __align__(8) uint64_t bars[STAGES][THREADS / 32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t gid = blockIdx.x * THREADS + tid;
    if (lane == 0)
        for (int st = 0; st < STAGES; ++st) mbar_init(&bars[st][warp], 1);
    __syncwarp();
    float acc = 0.f;
    // prologue: STAGES - 1 copies in flight
    for (int s = 0; s < STAGES - 1 && s < steps; ++s) {
        if (lane == 0) mbar_arrive_expect_tx(&bars[s][warp], 32 * 64);
        __syncwarp();
        const uint32_t c = hash32(gid * 9781u + s * 6271u) % nCells;
        tma_bulk_g2s(&cells[s][tid][0], vol + 2 * (size_t)c, 64, &bars[s][warp]);
    }
    for (int s = 0; s < steps; ++s) {
        const int sn = s + STAGES - 1;
        if (sn < steps) {
            const int st = sn % STAGES;
            if (lane == 0) mbar_arrive_expect_tx(&bars[st][warp], 32 * 64);
            __syncwarp();
            const uint32_t c = hash32(gid * 9781u + sn * 6271u) % nCells;
            tma_bulk_g2s(&cells[st][tid][0], vol + 2 * (size_t)c, 64, &bars[st][warp]);
        }
        const int st = s % STAGES;
        mbar_wait(&bars[st][warp], (s / STAGES) & 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = cells[st][tid][k];
            acc += v.x * 0.5f + v.y + v.z + v.w;
        }
        __syncwarp();
    }
    if (acc == 123.456f) out[gid] = acc;
}

// mode 2: the cell through the TEXTURE data pipe (4 x 128-bit fetches from a linear texture); mode 3: even steps LDG.256 x 2, odd steps
// texture - do the two data pipes of the L1/TEX unit add up?
__global__ void __launch_bounds__(THREADS, 2) gather_tex(cudaTextureObject_t tex, uint32_t nCells, int steps, float* out)
{
    const uint32_t gid = blockIdx.x * THREADS + threadIdx.x;
    float acc = 0.f;
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
        const uint32_t c = hash32(gid * 9781u + s * 6271u) % nCells;
        const float4 a = tex1Dfetch<float4>(tex, (int)(4 * c)), b = tex1Dfetch<float4>(tex, (int)(4 * c + 1));
        const float4 d = tex1Dfetch<float4>(tex, (int)(4 * c + 2)), e = tex1Dfetch<float4>(tex, (int)(4 * c + 3));
        acc += (a.x + a.y + a.z + a.w) * 0.5f + b.x + b.y + b.z + b.w + d.x + d.y + d.z + d.w + e.x + e.y + e.z + e.w;
    }
    if (acc == 123.456f) out[gid] = acc;
}

__global__ void __launch_bounds__(THREADS, 2) gather_mix(const Quad* __restrict__ vol, cudaTextureObject_t tex, uint32_t nCells, int steps, float* out)
{
    const uint32_t gid = blockIdx.x * THREADS + threadIdx.x;
    float acc = 0.f;
    for (int s = 0; s < steps; s += 2) {
        const uint32_t c0 = hash32(gid * 9781u + s * 6271u) % nCells, c1 = hash32(gid * 9781u + (s + 1) * 6271u) % nCells;
        const Quad a = ldg_quad(vol + 2 * (size_t)c0), b = ldg_quad(vol + 2 * (size_t)c0 + 1);
        const float4 t0 = tex1Dfetch<float4>(tex, (int)(4 * c1)), t1 = tex1Dfetch<float4>(tex, (int)(4 * c1 + 1));
        const float4 t2 = tex1Dfetch<float4>(tex, (int)(4 * c1 + 2)), t3 = tex1Dfetch<float4>(tex, (int)(4 * c1 + 3));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += a.v[k] * 0.5f + b.v[k];
        acc += t0.x + t0.y + t0.z + t0.w + t1.x + t1.y + t1.z + t1.w + t2.x + t2.y + t2.z + t2.w + t3.x + t3.y + t3.z + t3.w;
    }
    if (acc == 123.456f) out[gid] = acc;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t maxBytes = (size_t)512 << 20;
    Quad* vol;
    float* out;
    cudaMalloc(&vol, maxBytes);
    cudaMemset(vol, 0, maxBytes);
    cudaMalloc(&out, sizeof(float) * sms * 2 * THREADS);
    const int steps = 2048, grid = sms * 2;
    cudaFuncSetAttribute(gather_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * THREADS * 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = vol;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>();
    rd.res.linear.sizeInBytes = maxBytes;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) { printf("texture object failed\n"); return 1; }
    printf("random 64-byte cells, %d CTAs x %d threads x %d steps\nfootprint_MB  mode  ms  Gcells/s  GB/s\n", grid, THREADS, steps);
    for (int mb : {16, 48, 96, 512}) {
        const uint32_t nCells = (uint32_t)(((size_t)mb << 20) / 64);
        for (int mode = 0; mode < 4; ++mode) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) gather_ldg<<<grid, THREADS>>>(vol, nCells, steps, out);
                else if (mode == 1) gather_tma<<<grid, THREADS, STAGES * THREADS * 64>>>(vol, nCells, steps, out);
                else if (mode == 2) gather_tex<<<grid, THREADS>>>(tex, nCells, steps, out);
                else gather_mix<<<grid, THREADS>>>(vol, tex, nCells, steps, out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            const double cellsN = (double)grid * THREADS * steps;
            printf("%6d  %s  %.3f  %.1f  %.0f\n", mb, (const char*[]){"ldg256x2", "tma64", "tex128x4", "ldg+tex"}[mode], best, cellsN / best / 1e6, cellsN * 64 / best / 1e6);
        }
    }
    return 0;
}
