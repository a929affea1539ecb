// ubench.cu -- B200 micro-measurements that drive the persistent decode kernel design.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench tools/ubench.cu
// 1. grid barrier latency: one shared counter (RED.release + acquire poll) vs per-CTA flags (all-to-all poll)
// 2. cold HBM stream vs L2-prefetched stream of a per-CTA slab (cp.async.bulk.prefetch.L2)
// 3. how long a barrier takes while every SM also has bulk prefetch traffic in flight
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void l2_prefetch(const void *p, unsigned bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory"); }
__device__ __forceinline__ uint4 ldg_stream16(const void *p) { uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }

__device__ __forceinline__ void bar_counter(unsigned *bar, unsigned &gen) {
    gen++;
    __syncthreads();
    if (threadIdx.x == 0) { red_release_add(bar, 1u); unsigned t = gen * gridDim.x; while (ld_acquire(bar) < t) {} }
    __syncthreads();
}
__device__ __forceinline__ void bar_flags(unsigned *flags, unsigned &gen) {
    gen++;
    __syncthreads();
    if (threadIdx.x == 0) st_release(flags + blockIdx.x, gen);
    if (threadIdx.x < gridDim.x) { while (ld_acquire(flags + threadIdx.x) < gen) {} }
    __syncthreads();
}

__global__ void k_barrier(unsigned *bar, unsigned *flags, int iters, int mode, long long *out, const uint8_t *slab, int pf_bytes) {
    unsigned gen = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (pf_bytes && threadIdx.x == 0) l2_prefetch(slab + ((size_t)blockIdx.x * 64 + (i & 63)) * (size_t)pf_bytes, pf_bytes);
        if (mode == 0) bar_counter(bar, gen); else bar_flags(flags, gen);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// each CTA streams `bytes` from its own slab with 512 threads x 16 rows-in-flight style loads
__global__ void k_stream(const uint8_t *slab, size_t bytes, int do_prefetch, int wait_cycles, long long *out, unsigned *sink) {
    const uint8_t *base = slab + (size_t)blockIdx.x * bytes;
    if (do_prefetch && threadIdx.x == 0) l2_prefetch(base, (unsigned)bytes);
    __syncthreads();
    if (wait_cycles) { long long t = clock64(); while (clock64() - t < wait_cycles) {} }
    __syncthreads();
    long long t0 = clock64();
    unsigned acc = 0;
    for (size_t off = (size_t)threadIdx.x * 16; off < bytes; off += (size_t)blockDim.x * 16 * 8) {
        uint4 w[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { size_t o = off + (size_t)j * blockDim.x * 16; w[j] = o < bytes ? ldg_stream16(base + o) : make_uint4(0, 0, 0, 0); }
#pragma unroll
        for (int j = 0; j < 8; j++) acc += w[j].x ^ w[j].y ^ w[j].z ^ w[j].w;
    }
    __syncthreads();
    long long t1 = clock64();
    if (acc == 0x12345) sink[0] = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
    int dev = 0, sms = 0, clk = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    printf("SMs %d, clock %.0f MHz\n", sms, clk / 1e3);
    unsigned *bar, *flags, *sink; long long *out;
    CK(cudaMalloc(&bar, 256)); CK(cudaMalloc(&flags, 4096)); CK(cudaMalloc(&sink, 256)); CK(cudaMalloc(&out, sms * 8));
    uint8_t *slab; size_t slab_bytes = (size_t)3 << 30;
    CK(cudaMalloc(&slab, slab_bytes)); CK(cudaMemset(slab, 1, slab_bytes));
    long long *h = (long long *)malloc(sms * 8);
    auto report = [&](const char *name, double div) {
        CK(cudaMemcpy(h, out, sms * 8, cudaMemcpyDeviceToHost));
        long long mx = 0, mn = 1ll << 62; double avg = 0;
        for (int i = 0; i < sms; i++) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; avg += h[i]; }
        printf("%-58s avg %.0f  min %.0f  max %.0f cycles\n", name, avg / sms / div, mn / div, mx / div);
    };
    const int iters = 2000;
    for (int mode = 0; mode < 2; mode++)
        for (int pf = 0; pf <= 196608; pf += 65536) {
            CK(cudaMemset(bar, 0, 256)); CK(cudaMemset(flags, 0, 4096));
            void *args[] = { &bar, &flags, (void *)&iters, &mode, &out, &slab, &pf };
            CK(cudaLaunchCooperativeKernel((const void *)k_barrier, dim3(sms), dim3(512), args, 0, 0));
            CK(cudaDeviceSynchronize());
            char name[128]; snprintf(name, sizeof name, "barrier %s, %d KB L2 prefetch per CTA per barrier", mode ? "flags  " : "counter", pf / 1024);
            report(name, iters);
        }
    for (size_t kb : { 192, 768 }) {
        size_t bytes = kb * 1024;
        for (int pf = 0; pf < 2; pf++) {
            // flush L2 by touching 512 MB elsewhere
            CK(cudaMemset(slab + ((size_t)2 << 30), 2, (size_t)512 << 20));
            int wait = pf ? 40000 : 0;
            k_stream<<<sms, 512>>>(slab, bytes, pf, wait, out, sink);
            CK(cudaDeviceSynchronize());
            char name[128]; snprintf(name, sizeof name, "stream %zu KB per CTA, %s", kb, pf ? "after L2 prefetch + 40k-cycle wait" : "cold from HBM");
            report(name, 1);
            printf("   -> %.1f B/cycle/SM\n", (double)bytes / ((double)h[0]));
        }
    }
    return 0;
}
