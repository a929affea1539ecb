/*
 * vb_runtime.cu -- device bring-up, HBM allocation, host->device mirror registry.
 *
 * The reference has no accelerator init hook outside USE_METAL (main.c:178-180),
 * so everything here is lazy and driven from vox_load()/the first kernel wrapper.
 * There is deliberately no CPU fallback: without an sm_100 device the engine
 * refuses to load (north_star "no CPU fallback").
 */
#include "vb_engine.h"
#include <string.h>

static VbEngine *g_default_engine = NULL;
static VbEngine  g_bare_engine;          /* used by kernel wrappers when no model is loaded */
static int       g_bare_ready = 0;

extern "C" { __thread jmp_buf *vb_err_jmp = NULL; }

extern "C" void vb_cuda_fail(cudaError_t err, const char *file, int line) {
    fprintf(stderr, "voxtral_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(err), file, line, cudaGetErrorString(err));
    const char *logp = getenv("VOX_CUDA_ERRLOG");       /* test harnesses capture stderr: keep a copy where they cannot lose it */
    if (logp) { FILE *f = fopen(logp, "a"); if (f) { fprintf(f, "CUDA error %s at %s:%d: %s\n", cudaGetErrorName(err), file, line, cudaGetErrorString(err)); fclose(f); } }
    if (vb_err_jmp) longjmp(*vb_err_jmp, 1);
    abort();
}

/* Test hook: VOX_CUDA_FAIL_ALLOC_AFTER=n makes the n-th device allocation from now on (and every later one) fail with
 * cudaErrorMemoryAllocation, so the out-of-memory paths of the API can be exercised (tests/test_gpu_error_paths.py). */
static long long g_alloc_budget = -1;
extern "C" void vox_cuda_debug_fail_alloc_after(long long n) { g_alloc_budget = n; }

extern "C" void vb_require_gpu(const char *what) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fprintf(stderr, "voxtral_b200: %s needs a CUDA device (sm_100a); none is available "
                        "(%s). This engine has no CPU fallback.\n", what, cudaGetErrorString(e));
        abort();
    }
}

extern "C" int vb_device_init(VbEngine *e) {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n <= 0) {
        fprintf(stderr, "voxtral_b200: no CUDA device available (%s); this engine has no CPU fallback\n",
                cudaGetErrorString(err));
        return -1;
    }
    int dev = 0;
    const char *env = getenv("VOX_CUDA_DEVICE");
    if (!env) env = getenv("LOCAL_RANK");          /* one process per GPU under torchrun */
    if (env) dev = atoi(env) % n;
    VB_CUDA_OK(cudaSetDevice(dev));
    cudaDeviceProp p;
    VB_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (p.major < 10) {
        fprintf(stderr, "voxtral_b200: device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; "
                        "the kernels in this library are built for sm_100a only\n",
                dev, p.name, p.major, p.minor);
        return -1;
    }
    e->device = dev; e->sm_count = p.multiProcessorCount; e->cc_major = p.major; e->cc_minor = p.minor;
    VB_CUDA_OK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    VB_CUDA_OK(cudaEventCreate(&e->ev0));
    VB_CUDA_OK(cudaEventCreate(&e->ev1));
    VB_CUDA_OK(cudaEventCreate(&e->ev_user0));
    VB_CUDA_OK(cudaEventCreate(&e->ev_user1));
    if (vox_verbose >= 1)
        fprintf(stderr, "CUDA device %d: %s, %d SMs, %.1f GB\n", dev, p.name, p.multiProcessorCount,
                (double)p.totalGlobalMem / 1e9);
    return 0;
}

extern "C" void vb_device_shutdown(VbEngine *e) {
    if (!e->stream) return;
    cudaStreamSynchronize(e->stream);
    for (int i = 0; i < e->n_owned; i++) cudaFree(e->owned[i]);
    free(e->owned); e->owned = NULL; e->n_owned = 0;
    if (!e->parent) { free(e->mirrors); e->mirrors = NULL; e->n_mirrors = 0; }
    for (int i = 0; i < VB_WS_SLOTS; i++) { cudaFree(e->ws[i]); e->ws[i] = NULL; e->ws_bytes[i] = 0; }
    cudaFree(e->d_dist_adapter); e->d_dist_adapter = NULL; e->dist_adapter_cap = 0;
    if (e->v2.err_host) { cudaFreeHost(e->v2.err_host); e->v2.err_host = NULL; }
    if (e->step_graph_ready) cudaGraphExecDestroy(e->step_graph);
    cudaEventDestroy(e->ev0); cudaEventDestroy(e->ev1);
    cudaEventDestroy(e->ev_user0); cudaEventDestroy(e->ev_user1);
    if (!e->parent) cudaStreamDestroy(e->stream);
    e->stream = NULL;
    if (g_default_engine == e) g_default_engine = NULL;
}

extern "C" void *vb_dev_alloc(size_t bytes) {
    void *p = NULL;
    if (g_alloc_budget >= 0) {
        if (g_alloc_budget == 0) vb_cuda_fail(cudaErrorMemoryAllocation, __FILE__, __LINE__);
        g_alloc_budget--;
    }
    VB_CUDA_OK(cudaMalloc(&p, bytes ? bytes : 16));
    return p;
}

extern "C" void vb_register_mirror(VbEngine *e, const void *host, size_t bytes, void *dev) {
    if (e->n_mirrors == e->cap_mirrors) {
        e->cap_mirrors = e->cap_mirrors ? e->cap_mirrors * 2 : 1024;
        e->mirrors = (VbHostMirror *)realloc(e->mirrors, sizeof(VbHostMirror) * e->cap_mirrors);
    }
    e->mirrors[e->n_mirrors++] = VbHostMirror{ host, bytes, dev };
}

extern "C" void *vb_dev_alloc_owned(VbEngine *e, size_t bytes) {
    void *d = vb_dev_alloc(bytes);
    if (e->n_owned == e->cap_owned) {
        e->cap_owned = e->cap_owned ? e->cap_owned * 2 : 1024;
        e->owned = (void **)realloc(e->owned, sizeof(void *) * e->cap_owned);
    }
    e->owned[e->n_owned++] = d;
    e->weight_bytes += bytes;
    return d;
}

/* Host -> device copy during vox_load: asynchronous on the engine's stream when the source lies inside the checkpoint mapping
 * that vox_load registered as pinned memory (DMA straight from the page cache at PCIe rate), synchronous through the driver's
 * staging buffer otherwise (small malloc'd f32 tensors). */
extern "C" void vb_load_copy(VbEngine *e, void *dev, const void *host, size_t bytes) {
    const uint8_t *h = (const uint8_t *)host;
    if (e->pin_base && h >= e->pin_base && h + bytes <= e->pin_base + e->pin_bytes)
        VB_CUDA_OK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, e->stream));
    else
        VB_CUDA_OK(cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice));
}
extern "C" void vb_load_copy_2d(VbEngine *e, void *dev, size_t dpitch, const void *host, size_t spitch, size_t width, size_t height) {
    const uint8_t *h = (const uint8_t *)host;
    if (e->pin_base && h >= e->pin_base && h + spitch * height <= e->pin_base + e->pin_bytes)
        VB_CUDA_OK(cudaMemcpy2DAsync(dev, dpitch, host, spitch, width, height, cudaMemcpyHostToDevice, e->stream));
    else
        VB_CUDA_OK(cudaMemcpy2D(dev, dpitch, host, spitch, width, height, cudaMemcpyHostToDevice));
}

extern "C" void *vb_dev_upload(VbEngine *e, const void *host, size_t bytes) {
    void *d = vb_dev_alloc_owned(e, bytes);
    vb_load_copy(e, d, host, bytes);
    vb_register_mirror(e, host, bytes, d);
    return d;
}

extern "C" void *vb_find_mirror(VbEngine *e, const void *host) {
    if (!e) return NULL;
    for (int i = 0; i < e->n_mirrors; i++)
        if (e->mirrors[i].host == host) return e->mirrors[i].dev;
    return NULL;
}

extern "C" float *vb_ws(VbEngine *e, int slot, size_t bytes) {
    if (bytes > e->ws_bytes[slot]) {
        VB_CUDA_OK(cudaStreamSynchronize(e->stream));
        cudaFree(e->ws[slot]);
        e->ws[slot] = NULL; e->ws_bytes[slot] = 0;          /* a failed allocation below must not leave a dangling slot */
        size_t want = bytes + bytes / 8 + 256;
        e->ws[slot] = (float *)vb_dev_alloc(want);
        e->ws_bytes[slot] = want;
    }
    return e->ws[slot];
}

extern "C" void vb_set_default_engine(VbEngine *e) { g_default_engine = e; }

extern "C" VbEngine *vb_default_engine(void) {
    if (g_default_engine) return g_default_engine;
    if (!g_bare_ready) {
        vb_require_gpu("the kernel dispatch surface");
        memset(&g_bare_engine, 0, sizeof g_bare_engine);
        if (vb_device_init(&g_bare_engine) != 0) abort();
        g_bare_ready = 1;
    }
    return &g_bare_engine;
}

/* ---- small public helpers (include/voxtral_b200.h section 7) ---- */
extern "C" void *vox_cuda_malloc(vox_ctx_t *ctx, size_t bytes) {
    VB_CUDA_OK(cudaSetDevice(vb_engine(ctx)->device));
    return vb_dev_alloc(bytes);
}
extern "C" void vox_cuda_free(vox_ctx_t *ctx, void *p) {
    VB_CUDA_OK(cudaSetDevice(vb_engine(ctx)->device));
    cudaStreamSynchronize(vb_engine(ctx)->stream);
    cudaFree(p);
}
extern "C" int vox_cuda_memcpy_h2d(vox_ctx_t *ctx, void *d, const void *h, size_t bytes) {
    VbEngine *e = vb_engine(ctx);
    VB_CUDA_OK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    return 0;
}
extern "C" int vox_cuda_memcpy_d2h(vox_ctx_t *ctx, void *h, const void *d, size_t bytes) {
    VbEngine *e = vb_engine(ctx);
    VB_CUDA_OK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, e->stream));
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    return 0;
}
extern "C" void vox_cuda_timer_start(vox_ctx_t *ctx) {
    VbEngine *e = vb_engine(ctx);
    VB_CUDA_OK(cudaStreamSynchronize(e->stream));
    VB_CUDA_OK(cudaEventRecord(e->ev_user0, e->stream));
}
extern "C" double vox_cuda_timer_stop_ms(vox_ctx_t *ctx) {
    VbEngine *e = vb_engine(ctx);
    VB_CUDA_OK(cudaEventRecord(e->ev_user1, e->stream));
    VB_CUDA_OK(cudaEventSynchronize(e->ev_user1));
    float ms = 0.f;
    VB_CUDA_OK(cudaEventElapsedTime(&ms, e->ev_user0, e->ev_user1));
    return (double)ms;
}
