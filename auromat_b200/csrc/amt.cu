// auromat_b200: CUDA kernels (sm_100a) + C ABI for the georeference + regrid hot path.
// See include/auromat_b200.h for the boundary and DESIGN.md for the kernel inventory.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared
//        -Xcompiler -fPIC  (auromat_b200/csrc/build.py)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <new>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>
#include "amt_math.cuh"

using namespace amt;

// ============================================================================ context
// Device-side temporaries are owned per (context, stream): calls issued on different streams of
// one context (the sequence pipeline uses five) never share a scratch arena or a statistics key
// block.  A context is driven by one host thread at a time.
struct Workspace {
    cudaStream_t stream;
    void* scratch;
    size_t scratch_bytes;
    void* stat_keys;          // device StatKeys block, self-cleaning (see k_outline_eval)
    void* stat_queue;         // outline node queue (kStatQueue flat corner indices)
};

struct amt_ctx {
    int device;
    int sm_count;
    unsigned long long launches;
    std::vector<Workspace> ws;
};

static Workspace& workspace(amt_ctx* ctx, cudaStream_t st) {
    for (auto& w : ctx->ws)
        if (w.stream == st) return w;
    ctx->ws.push_back(Workspace{st, nullptr, 0, nullptr, nullptr});
    return ctx->ws.back();
}

static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return set_err(AMT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                             \
    } while (0)

#define CHECK_ARG(cond, msg)                                                \
    do {                                                                    \
        if (!(cond)) return set_err(AMT_ERR_INVALID_ARGUMENT, "%s", msg);   \
    } while (0)

#define LAUNCH_CHECK(ctx)                                                                   \
    do {                                                                                    \
        (ctx)->launches++;                                                                  \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess)                                                              \
            return set_err(AMT_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    } while (0)

static int ensure_scratch(amt_ctx* ctx, cudaStream_t st, size_t bytes, void** ptr) {
    Workspace& w = workspace(ctx, st);
    if (w.scratch_bytes < bytes) {
        if (w.scratch) CUDA_TRY(cudaFree(w.scratch));        // synchronises the device: safe, and rare
        w.scratch = nullptr;
        w.scratch_bytes = 0;
        size_t want = bytes + bytes / 4;
        CUDA_TRY(cudaMalloc(&w.scratch, want));
        w.scratch_bytes = want;
    }
    *ptr = w.scratch;
    return AMT_OK;
}

extern "C" const char* amt_last_error(void) { return g_err; }
extern "C" int amt_abi_version(void) { return AMT_ABI_VERSION; }

extern "C" int amt_ctx_create(int device, amt_ctx** out) {
    CHECK_ARG(out != nullptr, "amt_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_err(AMT_ERR_NO_DEVICE, "no CUDA device available (%s)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return set_err(AMT_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_err(AMT_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                       device, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(device));
    amt_ctx* c = new (std::nothrow) amt_ctx();
    if (!c) return set_err(AMT_ERR_CUDA, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->launches = 0;
    *out = c;
    return AMT_OK;
}

extern "C" int amt_ctx_destroy(amt_ctx* ctx) {
    if (!ctx) return AMT_OK;
    cudaSetDevice(ctx->device);
    for (auto& w : ctx->ws) {
        if (w.scratch) cudaFree(w.scratch);
        if (w.stat_keys) cudaFree(w.stat_keys);
        if (w.stat_queue) cudaFree(w.stat_queue);
    }
    delete ctx;
    return AMT_OK;
}

extern "C" int amt_ctx_device(const amt_ctx* ctx, int* device) {
    CHECK_ARG(ctx && device, "amt_ctx_device: NULL argument");
    *device = ctx->device;
    return AMT_OK;
}

extern "C" int amt_ctx_launch_count(const amt_ctx* ctx, uint64_t* count) {
    CHECK_ARG(ctx && count, "amt_ctx_launch_count: NULL argument");
    *count = ctx->launches;
    return AMT_OK;
}

#define ENTER(ctx)                                               \
    CHECK_ARG((ctx) != nullptr, "context is NULL");              \
    CUDA_TRY(cudaSetDevice((ctx)->device))

extern "C" int amt_alloc_device(amt_ctx* ctx, size_t bytes, void** d_ptr) {
    ENTER(ctx);
    CHECK_ARG(d_ptr, "amt_alloc_device: d_ptr is NULL");
    CUDA_TRY(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return AMT_OK;
}
extern "C" int amt_free_device(amt_ctx* ctx, void* d_ptr) {
    ENTER(ctx);
    CUDA_TRY(cudaFree(d_ptr));
    return AMT_OK;
}
extern "C" int amt_alloc_pinned(amt_ctx* ctx, size_t bytes, void** h_ptr) {
    ENTER(ctx);
    CHECK_ARG(h_ptr, "amt_alloc_pinned: h_ptr is NULL");
    CUDA_TRY(cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return AMT_OK;
}
extern "C" int amt_free_pinned(amt_ctx* ctx, void* h_ptr) {
    ENTER(ctx);
    CUDA_TRY(cudaFreeHost(h_ptr));
    return AMT_OK;
}
extern "C" int amt_copy_h2d(amt_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, void* stream) {
    ENTER(ctx);
    CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return AMT_OK;
}
extern "C" int amt_copy_h2d_2d(amt_ctx* ctx, void* d_dst, size_t d_pitch, const void* h_src, size_t h_pitch,
                               size_t width_bytes, size_t rows, void* stream) {
    ENTER(ctx);
    CHECK_ARG(width_bytes <= d_pitch && width_bytes <= h_pitch, "amt_copy_h2d_2d: width exceeds a pitch");
    if (rows == 0 || width_bytes == 0) return AMT_OK;
    CUDA_TRY(cudaMemcpy2DAsync(d_dst, d_pitch, h_src, h_pitch, width_bytes, rows, cudaMemcpyHostToDevice,
                               (cudaStream_t)stream));
    return AMT_OK;
}
extern "C" int amt_copy_d2h(amt_ctx* ctx, void* h_dst, const void* d_src, size_t bytes, void* stream) {
    ENTER(ctx);
    CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return AMT_OK;
}
extern "C" int amt_memset_device(amt_ctx* ctx, void* d_ptr, int value, size_t bytes, void* stream) {
    ENTER(ctx);
    CUDA_TRY(cudaMemsetAsync(d_ptr, value, bytes, (cudaStream_t)stream));
    return AMT_OK;
}
extern "C" int amt_stream_synchronize(amt_ctx* ctx, void* stream) {
    ENTER(ctx);
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return AMT_OK;
}

// ------------------------------------------------------------------ FP64 peak probe
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, double b, double c0, int iters) {
    double a[8];
    const double t = threadIdx.x * 1e-9;
    const double c = c0 + t;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 1.0 + j * 0.125 + t;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fma(a[j], b, c);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int amt_measure_fp64_peak(amt_ctx* ctx, double* dfma_per_second) {
    ENTER(ctx);
    CHECK_ARG(dfma_per_second, "amt_measure_fp64_peak: NULL argument");
    const int blocks = ctx->sm_count * 8, iters = 4096;
    void* scratch = nullptr;
    int rc = ensure_scratch(ctx, (cudaStream_t)0, (size_t)blocks * 256 * sizeof(double), &scratch);
    if (rc) return rc;
    double* out = (double*)scratch;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    k_fp64_peak<<<blocks, 256>>>(out, 1.0000001, 1e-9, 64);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CUDA_TRY(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, 256>>>(out, 1.0000001, 1e-9, iters);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->launches += 6;
    *dfma_per_second = (double)blocks * 256 * iters * 8 / (best * 1e-3);
    return AMT_OK;
}

// ------------------------------------------------------------------ L2 atomic peak probe
// Throughput ceiling of the scatter: u64 atomicAdd (RED) to pseudo-random cells of a grid of `cells`
// words, 32 distinct addresses per warp instruction -- the pattern of the run tails of the binning.
__global__ void __launch_bounds__(256) k_atomic_peak(unsigned long long* grid, unsigned int cells, int iters) {
    unsigned int h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; ++i) {
        h = h * 1664525u + 1013904223u;
        atomicAdd(&grid[(h >> 4) % cells], 1ULL);
    }
}

extern "C" int amt_measure_atomic_peak(amt_ctx* ctx, size_t cells, double* atomics_per_second) {
    ENTER(ctx);
    CHECK_ARG(atomics_per_second && cells > 0 && cells < (1ULL << 31), "amt_measure_atomic_peak: bad arguments");
    void* scratch = nullptr;
    int rc = ensure_scratch(ctx, (cudaStream_t)0, cells * 8, &scratch);
    if (rc) return rc;
    CUDA_TRY(cudaMemset(scratch, 0, cells * 8));
    const int blocks = ctx->sm_count * 8, iters = 256;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    k_atomic_peak<<<blocks, 256>>>((unsigned long long*)scratch, (unsigned int)cells, 16);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CUDA_TRY(cudaEventRecord(e0));
        k_atomic_peak<<<blocks, 256>>>((unsigned long long*)scratch, (unsigned int)cells, iters);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->launches += 6;
    *atomics_per_second = (double)blocks * 256 * iters / (best * 1e-3);
    return AMT_OK;
}

// ===================================================================== georeference
struct GeorefParams {
    FrameC f;
    double sip_a[AMT_SIP_MAX_COEF];
    double sip_b[AMT_SIP_MAX_COEF];
    amt_georef_out o;
    unsigned long long* ill;     // n_ill_conditioned counter (nullable)
    // Row permutation of the point kernels: grid row r works on image row (r * row_stride) % rows,
    // row_stride coprime to rows and close to rows / golden ratio.  An ISS limb frame is ~40 % sky at
    // the top: in image order the kernel would first run a store-only phase (NaN rows, HBM bound) and
    // then an FP64-bound phase; permuted, every window of consecutive CTAs holds the frame's mix of
    // both, so the NaN stores of some CTAs overlap the arithmetic of their neighbours on the same SM.
    unsigned int row_stride;
};

static unsigned int gcd_u(unsigned int a, unsigned int b) { return b ? gcd_u(b, a % b) : a; }
static unsigned int golden_stride(unsigned int rows) {
    if (rows < 3 || getenv("AMT_NO_ROW_PERMUTATION")) return 1;
    unsigned int s = (unsigned int)(rows * 0.6180339887498949);
    if (s < 1) s = 1;
    while (gcd_u(s, rows) != 1) ++s;
    return s % rows ? s % rows : 1;
}

// The 817 MB of coordinate planes a frame writes are never read again by the kernel that writes them (the passes that follow read
// them from DRAM anyway: 868 MB against 126 MB of L2): streaming stores
// (st.global.cs, evict-first in L2) keep them from displacing the L2-resident working set of everything that
// runs beside the kernel -- the grid accumulators of its own atomics (8 MB), the bitmaps and node queue of the
// stage-A kernels of other frames.  Kernel alone 237.5 -> 236.1 us, frame period in the engine 0.261 -> 0.257 ms.
#ifndef AMT_PLANE_STCS
#define AMT_PLANE_STCS 1
#endif
#if AMT_PLANE_STCS
#define PLANE_ST(ptr, v) __stcs((ptr), (v))
#else
#define PLANE_ST(ptr, v) (*(ptr) = (v))
#endif

// Writes NaN to every requested plane of one point (a ray that misses the ellipsoid).
__device__ __forceinline__ void emit_nan(size_t i, double* __restrict__ a, double* __restrict__ b,
                                         double* __restrict__ c, double* __restrict__ d) {
    const double nan = qnan();
    if (a) PLANE_ST(&a[i], nan);
    if (b) PLANE_ST(&b[i], nan);
    if (c) PLANE_ST(&c[i], nan);
    if (d) PLANE_ST(&d[i], nan);
}

// One intersection point -> all requested outputs at flat index i.
// Returns |P|^2 (needed by the elevation).
__device__ __forceinline__ double emit_point(const GeorefParams& p, const double P[3], size_t i,
                                             double* __restrict__ lat_o, double* __restrict__ lon_o,
                                             double* __restrict__ mlat_o, double* __restrict__ mlt_o) {
    double r2;
    if (lat_o || lon_o) {
        double lat, lon;
        point_to_geo(p.f, P, lat, lon, r2);
        if (lat_o) PLANE_ST(&lat_o[i], lat);
        if (lon_o) PLANE_ST(&lon_o[i], lon);
    } else {
        r2 = fma(P[2], P[2], fma(P[1], P[1], P[0] * P[0]));
    }
    if (mlat_o || mlt_o) {
        double mlat, mlt;
        point_to_mag(p.f, P, mlat, mlt);
        if (mlat_o) PLANE_ST(&mlat_o[i], mlat);
        if (mlt_o) PLANE_ST(&mlt_o[i], mlt);
    }
    return r2;
}

__device__ __forceinline__ void count_grazing(unsigned long long* counter, bool graze, unsigned lane) {
    if (counter) {
        const unsigned m = __ballot_sync(0xffffffffu, graze);
        if (m && lane == 0) atomicAdd(counter, (unsigned long long)__popc(m));
    }
}

// fastCenterCalculation == False: corners and centres are independent rays
// (mapping/astrometry.py:49-64,86-106).  One thread per pixel evaluates BOTH the corner
// (x-0.5, y-0.5) and the centre (x, y): two independent FP64 dependency chains per thread
// (ILP for the Newton / polynomial sequences) that share the per-frame constants.  Launch:
// (W+1) x (H+1) threads, a warp covers 32 consecutive x of one row, so its two hit ballots are
// exactly one word each of the row-padded validity bitmaps.
// Rays that miss the ellipsoid (reference: NaN rows that propagate through every later pass)
// cost only the direction + discriminant: a warp whose rays all miss leaves right there.
// FULL: all nine planes and both bitmaps are requested (the sequence pipeline with MLat/MLT): no
// null-pointer tests, 32-bit indices.
// PLAIN: pure TAN header and WCS camera model (no SIP polynomial, no fisheye branch in the code).
template <bool WANT_K, bool WANT_C, bool FULL, bool PLAIN = false>
#ifndef AMT_GEOREF_MINBLOCKS
#define AMT_GEOREF_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(256, AMT_GEOREF_MINBLOCKS) k_georef_points(const __grid_constant__ GeorefParams p) {
    const int W = p.f.W, H = p.f.H;
    const int y = (int)((blockIdx.y * p.row_stride) % gridDim.y);
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_k = WANT_K && x <= W;
    const bool in_c = WANT_C && x < W && y < H;
    // per-frame SIP coefficients: constant bank -> shared memory once per CTA (uniform branch)
    __shared__ double s_sip[PLAIN ? 1 : 2 * AMT_SIP_MAX_COEF];
    if (!PLAIN && (p.f.sip_oa | p.f.sip_ob)) {
        if (threadIdx.x < 2 * AMT_SIP_MAX_COEF)
            s_sip[threadIdx.x] = threadIdx.x < AMT_SIP_MAX_COEF ? p.sip_a[threadIdx.x]
                                                                : p.sip_b[threadIdx.x - AMT_SIP_MAX_COEF];
        __syncthreads();
    }
    const double* sip_a = s_sip;
    const double* sip_b = PLAIN ? s_sip : s_sip + AMT_SIP_MAX_COEF;
    bool graze_k = false, graze_c = false, hit_k = false, hit_c = false;
    double dk[3], Pk[3], dc[3], Pc[3], cam_el = 0.0;
    const double nan = qnan();
    if (in_k | in_c) {
        // wcs.py:41-44: corner grids start at -0.5
        if (!PLAIN && p.f.model == AMT_MODEL_ALLSKY) {
            const double fx = (double)x, fy = (double)y;
            if (WANT_K) pix2dir_allsky(p.f, fx - 0.5, fy - 0.5, dk);
            if (WANT_C) cam_el = pix2dir_allsky(p.f, fx, fy, dc);
        } else {
            dirs_kc<!PLAIN>(p.f, sip_a, sip_b, x, y, dk, dc);
        }
        if (WANT_K) hit_k = intersect(p.f, dk, Pk, graze_k) && in_k;
        if (WANT_C) hit_c = intersect(p.f, dc, Pc, graze_c) && in_c;
        graze_k &= in_k;
        graze_c &= in_c;
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned mk = __ballot_sync(0xffffffffu, hit_k), mc = __ballot_sync(0xffffffffu, hit_c);
    if (lane == 0) {
        const int wk = (W + 1 + 31) >> 5, wc = (W + 31) >> 5;
        const unsigned xw = (unsigned)x >> 5;
        if (WANT_K && (FULL || p.o.d_valid_k) && xw < (unsigned)wk) p.o.d_valid_k[(unsigned)y * wk + xw] = mk;
        if (WANT_C && (FULL || p.o.d_valid_c) && y < H && xw < (unsigned)wc) p.o.d_valid_c[(unsigned)y * wc + xw] = mc;
    }
    count_grazing(p.ill, graze_k | graze_c, lane);
    // fill_frame guarantees (W+1)*(H+1) < 2^31: 32-bit flat indices
    const unsigned ik = (unsigned)y * (unsigned)(W + 1) + (unsigned)x, ic = (unsigned)y * (unsigned)W + (unsigned)x;
    if ((mk | mc) == 0) {                    // the whole warp looks at space
        if (FULL) {
            if (in_k) { PLANE_ST(&p.o.d_lat_k[ik], nan); PLANE_ST(&p.o.d_lon_k[ik], nan); PLANE_ST(&p.o.d_mlat_k[ik], nan); PLANE_ST(&p.o.d_mlt_k[ik], nan); }
            if (in_c) {
                PLANE_ST(&p.o.d_lat_c[ic], nan); PLANE_ST(&p.o.d_lon_c[ic], nan); PLANE_ST(&p.o.d_mlat_c[ic], nan); PLANE_ST(&p.o.d_mlt_c[ic], nan);
                PLANE_ST(&p.o.d_elev_c[ic], nan);
            }
            return;
        }
        if (in_k) emit_nan(ik, p.o.d_lat_k, p.o.d_lon_k, p.o.d_mlat_k, p.o.d_mlt_k);
        if (in_c) {
            emit_nan(ic, p.o.d_lat_c, p.o.d_lon_c, p.o.d_mlat_c, p.o.d_mlt_c);
            if (p.o.d_elev_c) PLANE_ST(&p.o.d_elev_c[ic], nan);
        }
        return;
    }
    // Both chains run unconditionally from here (a missed ray carries NaN through the
    // arithmetic, exactly like the reference's NaN rows); only the stores are predicated.
    if (!hit_k) Pk[0] = Pk[1] = Pk[2] = nan;
    if (!hit_c) Pc[0] = Pc[1] = Pc[2] = nan;
    const bool geo = FULL || p.o.d_lat_k || p.o.d_lon_k || p.o.d_lat_c || p.o.d_lon_c;
    const bool mag = FULL || p.o.d_mlat_k || p.o.d_mlt_k || p.o.d_mlat_c || p.o.d_mlt_c;
    double r2_c = 0.0;
    if (geo) {
        double la_k, lo_k, la_c, lo_c;
        if (WANT_K) point_to_geo(p.f, Pk, la_k, lo_k);
        if (WANT_C) point_to_geo(p.f, Pc, la_c, lo_c, r2_c);
        if (WANT_K && in_k) {
            if (FULL || p.o.d_lat_k) PLANE_ST(&p.o.d_lat_k[ik], la_k);
            if (FULL || p.o.d_lon_k) PLANE_ST(&p.o.d_lon_k[ik], lo_k);
        }
        if (WANT_C && in_c) {
            if (FULL || p.o.d_lat_c) PLANE_ST(&p.o.d_lat_c[ic], la_c);
            if (FULL || p.o.d_lon_c) PLANE_ST(&p.o.d_lon_c[ic], lo_c);
        }
    }
    if (mag) {
        double ml_k, mt_k, ml_c, mt_c;
        if (WANT_K) point_to_mag(p.f, Pk, ml_k, mt_k);
        if (WANT_C) point_to_mag(p.f, Pc, ml_c, mt_c);
        if (WANT_K && in_k) {
            if (FULL || p.o.d_mlat_k) PLANE_ST(&p.o.d_mlat_k[ik], ml_k);
            if (FULL || p.o.d_mlt_k) PLANE_ST(&p.o.d_mlt_k[ik], mt_k);
        }
        if (WANT_C && in_c) {
            if (FULL || p.o.d_mlat_c) PLANE_ST(&p.o.d_mlat_c[ic], ml_c);
            if (FULL || p.o.d_mlt_c) PLANE_ST(&p.o.d_mlt_c[ic], mt_c);
        }
    }
    if (WANT_C && in_c && (FULL || p.o.d_elev_c)) {
        if (!geo) r2_c = fma(Pc[2], Pc[2], fma(Pc[1], Pc[1], Pc[0] * Pc[0]));
        double e = (!PLAIN && p.f.model == AMT_MODEL_ALLSKY) ? cam_el : elevation_deg<false>(dc, Pc, r2_c);
        PLANE_ST(&p.o.d_elev_c[ic], hit_c ? e : nan);
    }
}

// Validity bitmaps only (plane-free resampling, `intersectsEarth`): both rays of every pixel up
// to the discriminant -- no square root, no division, no coordinate plane.
__global__ void __launch_bounds__(256) k_hit_bits(const __grid_constant__ GeorefParams p) {
    const int W = p.f.W, H = p.f.H;
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_k = x <= W;
    const bool in_c = x < W && y < H;
    __shared__ double s_sip[2 * AMT_SIP_MAX_COEF];
    if (p.f.sip_oa | p.f.sip_ob) {
        if (threadIdx.x < 2 * AMT_SIP_MAX_COEF)
            s_sip[threadIdx.x] = threadIdx.x < AMT_SIP_MAX_COEF ? p.sip_a[threadIdx.x]
                                                                : p.sip_b[threadIdx.x - AMT_SIP_MAX_COEF];
        __syncthreads();
    }
    bool graze_k = false, graze_c = false, hit_k = false, hit_c = false;
    if (in_k) {
        double dk[3], dc[3];
        if (p.f.model == AMT_MODEL_ALLSKY) {
            const double fx = (double)x, fy = (double)y;
            pix2dir_allsky(p.f, fx - 0.5, fy - 0.5, dk);
            pix2dir_allsky(p.f, fx, fy, dc);
        } else {
            dirs_kc<true>(p.f, s_sip, s_sip + AMT_SIP_MAX_COEF, x, y, dk, dc);
        }
        hit_k = intersect_hit(p.f, dk, graze_k);
        hit_c = intersect_hit(p.f, dc, graze_c) && in_c;
        graze_c &= in_c;
    }
    const unsigned lane = threadIdx.x & 31;
    const unsigned mk = __ballot_sync(0xffffffffu, hit_k), mc = __ballot_sync(0xffffffffu, hit_c);
    if (lane == 0) {
        const int wk = (W + 1 + 31) >> 5, wc = (W + 31) >> 5;
        const unsigned xw = (unsigned)x >> 5;
        if (p.o.d_valid_k && xw < (unsigned)wk) p.o.d_valid_k[(unsigned)y * wk + xw] = mk;
        if (p.o.d_valid_c && y < H && xw < (unsigned)wc) p.o.d_valid_c[(unsigned)y * wc + xw] = mc;
    }
    count_grazing(p.ill, graze_k | graze_c, lane);
}

// ------------------------------------------------------------------ limb solver
// Hit bitmaps of a pure-TAN frame in O(H) ray evaluations instead of O(W H).  With the affine ray
// model the discriminant of the intersection is, along an image row, a QUADRATIC in the pixel
// index x:  D(x) = D0 + x D1  =>  dDO = d0 + d1 x,  dDD = e0 + e1 x + e2 x^2,
//     rt(x) = dDO^2 + (1 - oDO) dDD = qa x^2 + qb x + qc.
// The hit predicate (rt >= 0 and, for a camera outside the ellipsoid, dDO >= 0) can therefore only
// change at the (at most two) real roots of rt -- dDO changes sign where rt = (1 - oDO) dDD < 0,
// inside a miss interval.  One warp per row: the roots are located analytically (both branches of
// an uncertainty band 1e-11 relative on the discriminant of the quadratic, which is 4 orders above
// the rounding of its coefficients), every 32-pixel bitmap word that lies within 2 pixels of a
// root interval is evaluated pixel by pixel with the SAME predicate the per-pixel kernels use
// (dirs_kc + intersect_hit: bit-identical decisions at the limb), every other word is constant
// and takes the predicate of its first pixel.  Rows whose quadratic degenerates (NaN, no usable
// pivot) are evaluated completely.  Grazing rays (rt/dDD < 1e-10) sit within 1e-2 pixels of a root,
// i.e. inside the evaluated words: the n_ill_conditioned count is complete.
__global__ void __launch_bounds__(128) k_limb_bits(const __grid_constant__ GeorefParams p,
                                                   uint32_t* __restrict__ valid_k, uint32_t* __restrict__ valid_c) {
    const int W = p.f.W, H = p.f.H;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y > H) return;                                   // warp-uniform
    const int lane = threadIdx.x & 31;
    const int wk = (W + 1 + 31) >> 5, wc = (W + 31) >> 5;
    const FrameC& f = p.f;
    // root intervals [lo, hi] (pixel index) of the corner row (kind 0) and the centre row (kind 1)
    double r_lo[2][2], r_hi[2][2];
    bool full_row = false;
#pragma unroll
    for (int kind = 0; kind < 2; ++kind) {
        double Da[3], Db[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double a0 = fma((double)y, f.aff_c[k], f.aff_a[k]);
            if (kind) a0 += f.aff_h[k];
            Da[k] = a0 * f.rad[k];
            Db[k] = f.aff_b[k] * f.rad[k];
        }
        const double d0 = Da[0] * f.otr[0] + Da[1] * f.otr[1] + Da[2] * f.otr[2];
        const double d1 = Db[0] * f.otr[0] + Db[1] * f.otr[1] + Db[2] * f.otr[2];
        const double e0 = Da[0] * Da[0] + Da[1] * Da[1] + Da[2] * Da[2];
        const double e1 = 2.0 * (Da[0] * Db[0] + Da[1] * Db[1] + Da[2] * Db[2]);
        const double e2 = Db[0] * Db[0] + Db[1] * Db[1] + Db[2] * Db[2];
        const double k1 = 1.0 - f.oDO;
        const double qa = d1 * d1 + k1 * e2, qb = 2.0 * d0 * d1 + k1 * e1, qc = d0 * d0 + k1 * e0;
        const double disc = qb * qb - 4.0 * qa * qc;
        const double band = 1e-11 * (qb * qb + fabs(4.0 * qa * qc));
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        r_lo[kind][0] = r_lo[kind][1] = inf;             // empty intervals
        r_hi[kind][0] = r_hi[kind][1] = -inf;
        if (!(disc == disc) || !(band == band) || isinf(disc)) {
            full_row = true;
        } else if (disc >= -band) {
            const double sq[2] = {sqrt(fmax(disc - band, 0.0)), sqrt(disc + band)};
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                // stable quadratic roots: q = -(qb + sign(qb) sq)/2, r1 = q/qa, r2 = qc/q
                const double q = -0.5 * (qb + copysign(sq[v], qb));
                const double r1 = q / qa, r2 = qc / q;
                if (q == 0.0 || !(r2 == r2)) full_row = true;
                if (r1 == r1) { r_lo[kind][0] = fmin(r_lo[kind][0], r1); r_hi[kind][0] = fmax(r_hi[kind][0], r1); }
                if (r2 == r2) { r_lo[kind][1] = fmin(r_lo[kind][1], r2); r_hi[kind][1] = fmax(r_hi[kind][1], r2); }
            }
        }
    }
    unsigned n_graze = 0;
    for (int base = 0; base < wk; base += 32) {           // warp-uniform
        const int i = base + lane;
        // pixels of word i: 32 i .. 32 i + 31; uncertain iff a root interval (+- 2 px) touches them
        const double x_lo = 32.0 * i - 2.0, x_hi = 32.0 * i + 33.0;
        bool unc = full_row;
#pragma unroll
        for (int kind = 0; kind < 2; ++kind)
#pragma unroll
            for (int r = 0; r < 2; ++r) unc |= r_lo[kind][r] <= x_hi && r_hi[kind][r] >= x_lo;
        unc &= i < wk;
        unsigned word_k = 0, word_c = 0;
        if (i < wk && !unc) {                             // constant word: predicate of its first pixel
            double dk[3], dc[3];
            bool gz;
            dirs_kc<false>(f, nullptr, nullptr, 32 * i, y, dk, dc);
            const int nk = min(32, W + 1 - 32 * i), nc = min(32, W - 32 * i);
            if (intersect_hit(f, dk, gz)) word_k = nk >= 32 ? 0xffffffffu : ((1u << nk) - 1u);
            if (nc > 0 && intersect_hit(f, dc, gz)) word_c = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
        }
        unsigned todo = __ballot_sync(0xffffffffu, unc);
        while (todo) {                                    // pixel-by-pixel words, the warp on the 32 pixels
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int x = 32 * (base + src) + lane;
            bool hk = false, hc = false, gk = false, gc = false;
            if (x <= W) {
                double dk[3], dc[3];
                dirs_kc<false>(f, nullptr, nullptr, x, y, dk, dc);
                hk = intersect_hit(f, dk, gk);
                hc = intersect_hit(f, dc, gc) && x < W;
                gc &= x < W && y < H;
            }
            const unsigned bk = __ballot_sync(0xffffffffu, hk), bc = __ballot_sync(0xffffffffu, hc);
            n_graze += __popc(__ballot_sync(0xffffffffu, gk | gc));
            if (lane == src) { word_k = bk; word_c = bc; }
        }
        if (i < wk) {
            valid_k[(size_t)y * wk + i] = word_k;
            if (y < H && i < wc) valid_c[(size_t)y * wc + i] = word_c;
        }
    }
    if (p.ill && n_graze && lane == 0) atomicAdd(p.ill, (unsigned long long)n_graze);
}

// The same for TAN-SIP frames.  Pixel (x, y) looks where the pure TAN model looks at (x + fx, y + fy) with
// |fx| <= sip_dx, |fy| <= sip_dy (rigorous bounds of the polynomial over the frame, fill_sip_bounds), and the
// discriminant of the pure TAN model is a quadratic FORM Q(x', y') of the (real-valued) pixel position.  For a
// bitmap word (32 pixels of a row, corner and centre rays) all positions lie in the box
//     [32 i - 1 - sip_dx, 32 i + 32 + sip_dx] x [y - 1 - sip_dy, y + 1 + sip_dy];
// Q at the box centre, its exact gradient and Hessian bound the variation over the box:
//     |Q(c + d) - Q(c)| <= |Qx| hx + |Qy| hy + (|Qxx| hx^2 + 2 |Qxy| hx hy + |Qyy| hy^2) / 2.
// If |Q(c)| exceeds the bound (with a 1e-6 relative reserve for the rounding of either evaluation), every
// ray of the word is on the same side of the limb: the word is constant and takes the EXACT predicate
// (SIP polynomial, dirs_kc + intersect_hit) of its first pixel -- where Q > 0 throughout, dDO cannot change
// sign either (it vanishes only where rt < 0).  Every other word is evaluated pixel by pixel with that
// predicate.  Same bitmaps as k_hit_bits, bit for bit (tests), at ~1/20 of its ray evaluations.
__global__ void __launch_bounds__(128) k_limb_bits_sip(const __grid_constant__ GeorefParams p,
                                                       uint32_t* __restrict__ valid_k,
                                                       uint32_t* __restrict__ valid_c) {
    const int W = p.f.W, H = p.f.H;
    __shared__ double s_sip[2 * AMT_SIP_MAX_COEF];
    for (int i = threadIdx.x; i < 2 * AMT_SIP_MAX_COEF; i += blockDim.x)
        s_sip[i] = i < AMT_SIP_MAX_COEF ? p.sip_a[i] : p.sip_b[i - AMT_SIP_MAX_COEF];
    __syncthreads();
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y > H) return;                                   // warp-uniform
    const int lane = threadIdx.x & 31;
    const int wk = (W + 1 + 31) >> 5, wc = (W + 31) >> 5;
    const FrameC& f = p.f;
    const double* sa = s_sip;
    const double* sb = s_sip + AMT_SIP_MAX_COEF;
    double Dx[3], Dy[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { Dx[k] = f.aff_b[k] * f.rad[k]; Dy[k] = f.aff_c[k] * f.rad[k]; }
    const double k1 = 1.0 - f.oDO;
    const double ox = Dx[0] * f.otr[0] + Dx[1] * f.otr[1] + Dx[2] * f.otr[2];
    const double oy = Dy[0] * f.otr[0] + Dy[1] * f.otr[1] + Dy[2] * f.otr[2];
    const double q_xx = 2.0 * ox * ox + k1 * 2.0 * (Dx[0] * Dx[0] + Dx[1] * Dx[1] + Dx[2] * Dx[2]);
    const double q_xy = 2.0 * ox * oy + k1 * 2.0 * (Dx[0] * Dy[0] + Dx[1] * Dy[1] + Dx[2] * Dy[2]);
    const double q_yy = 2.0 * oy * oy + k1 * 2.0 * (Dy[0] * Dy[0] + Dy[1] * Dy[1] + Dy[2] * Dy[2]);
    const double hx = 16.5 + f.sip_dx, hy = 1.0 + f.sip_dy;
    const double curv = 0.5 * (fabs(q_xx) * hx * hx + 2.0 * fabs(q_xy) * hx * hy + fabs(q_yy) * hy * hy);
    unsigned n_graze = 0;
    for (int base = 0; base < wk; base += 32) {           // warp-uniform
        const int i = base + lane;
        bool unc = true;
        if (i < wk) {
            const double cx = 32.0 * i + 15.5, cy = (double)y;          // corner-model coordinates of the box centre
            double Dc[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) Dc[k] = (f.aff_a[k] + f.aff_b[k] * cx + f.aff_c[k] * cy) * f.rad[k];
            const double dDO = Dc[0] * f.otr[0] + Dc[1] * f.otr[1] + Dc[2] * f.otr[2];
            const double dDD = Dc[0] * Dc[0] + Dc[1] * Dc[1] + Dc[2] * Dc[2];
            const double Q = dDO * dDO + k1 * dDD;
            const double q_x = 2.0 * dDO * ox + k1 * 2.0 * (Dc[0] * Dx[0] + Dc[1] * Dx[1] + Dc[2] * Dx[2]);
            const double q_y = 2.0 * dDO * oy + k1 * 2.0 * (Dc[0] * Dy[0] + Dc[1] * Dy[1] + Dc[2] * Dy[2]);
            const double bound = fabs(q_x) * hx + fabs(q_y) * hy + curv;
            // reserve: 1e-6 of the bound and of the cancelling terms of Q itself
            const double reserve = 1e-6 * (bound + dDO * dDO + fabs(k1 * dDD));
            unc = !(fabs(Q) > bound + reserve);              // NaN -> uncertain
        }
        unc &= i < wk;
        unsigned word_k = 0, word_c = 0;
        if (i < wk && !unc) {                             // constant word: exact predicate of its first pixel
            double dk[3], dc[3];
            bool gz;
            dirs_kc<true>(f, sa, sb, 32 * i, y, dk, dc);
            const int nk = min(32, W + 1 - 32 * i), nc = min(32, W - 32 * i);
            if (intersect_hit(f, dk, gz)) word_k = nk >= 32 ? 0xffffffffu : ((1u << nk) - 1u);
            if (nc > 0 && intersect_hit(f, dc, gz)) word_c = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);
        }
        unsigned todo = __ballot_sync(0xffffffffu, unc);
        while (todo) {                                    // pixel-by-pixel words, the warp on the 32 pixels
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int x = 32 * (base + src) + lane;
            bool hk = false, hc = false, gk = false, gc = false;
            if (x <= W) {
                double dk[3], dc[3];
                dirs_kc<true>(f, sa, sb, x, y, dk, dc);
                hk = intersect_hit(f, dk, gk);
                hc = intersect_hit(f, dc, gc) && x < W;
                gc &= x < W && y < H;
            }
            const unsigned bk = __ballot_sync(0xffffffffu, hk), bc = __ballot_sync(0xffffffffu, hc);
            n_graze += __popc(__ballot_sync(0xffffffffu, gk | gc));
            if (lane == src) { word_k = bk; word_c = bc; }
        }
        if (i < wk) {
            valid_k[(size_t)y * wk + i] = word_k;
            if (y < H && i < wc) valid_c[(size_t)y * wc + i] = word_c;
        }
    }
    if (p.ill && n_graze && lane == 0) atomicAdd(p.ill, (unsigned long long)n_graze);
}

// fastCenterCalculation == True: a CTA evaluates a (TH+1)x(TW+1) patch of corner rays into
// shared memory, then derives each centre from the mean of its 4 corner intersection points
// and (un-normalised) directions: mapping/astrometry.py:154-160 (`_calcCenters`).
constexpr int TW = 32, TH = 8;

// one corner ray of the tile: direction + intersection into shared memory, outputs if owned
__device__ __forceinline__ bool tile_corner(const GeorefParams& p, const double* sip_a, const double* sip_b,
                                            double (*sP)[TH + 1][TW + 1],
                                            double (*sD)[TH + 1][TW + 1], int x0, int y0, int cx, int cy,
                                            bool& graze) {
    const int W = p.f.W, H = p.f.H;
    const int x = x0 + cx, y = y0 + cy;
    bool hit = false;
    if (x <= W && y <= H) {
        double dir[3], P[3];
        bool g;
        pix2dir<true>(p.f, sip_a, sip_b, (double)x - 0.5, (double)y - 0.5, dir);
        hit = intersect(p.f, dir, P, g);
#pragma unroll
        for (int k = 0; k < 3; ++k) { sP[k][cy][cx] = hit ? P[k] : qnan(); sD[k][cy][cx] = dir[k]; }
        // each corner is owned by exactly one tile for the global outputs / counters
        const bool own = (cx < TW || x == W) && (cy < TH || y == H);
        if (own) {
            graze |= g;
            const size_t i = (size_t)y * (W + 1) + x;
            if (hit) emit_point(p, P, i, p.o.d_lat_k, p.o.d_lon_k, p.o.d_mlat_k, p.o.d_mlt_k);
            else emit_nan(i, p.o.d_lat_k, p.o.d_lon_k, p.o.d_mlat_k, p.o.d_mlt_k);
        } else {
            hit = false;     // not ours: do not publish the bit
        }
    }
    return hit;
}

__global__ void __launch_bounds__(TW* TH) k_georef_tiles(const __grid_constant__ GeorefParams p) {
    __shared__ double sP[3][TH + 1][TW + 1];
    __shared__ double sD[3][TH + 1][TW + 1];
    __shared__ double s_sip[2 * AMT_SIP_MAX_COEF];
    const int W = p.f.W, H = p.f.H;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * TW + tx;
    if (p.f.sip_oa | p.f.sip_ob) {
        if (tid < 2 * AMT_SIP_MAX_COEF)
            s_sip[tid] = tid < AMT_SIP_MAX_COEF ? p.sip_a[tid] : p.sip_b[tid - AMT_SIP_MAX_COEF];
        __syncthreads();
    }
    const double* sip_a = s_sip;
    const double* sip_b = s_sip + AMT_SIP_MAX_COEF;
    const int wpr_k = (W + 1 + 31) >> 5, wpr_c = (W + 31) >> 5;
    bool graze = false;
    // main 32x8 block of corners: one warp per row == one bitmap word
    {
        const bool hit = tile_corner(p, sip_a, sip_b, sP, sD, x0, y0, tx, ty, graze);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (p.o.d_valid_k && tx == 0 && y0 + ty <= H && (x0 >> 5) < wpr_k)
            p.o.d_valid_k[(size_t)(y0 + ty) * wpr_k + (x0 >> 5)] = m;
    }
    // halo row cy == TH (warp 0) and halo column cx == TW (warp 1, lanes 0..TH)
    if (ty == 0) {
        const bool hit = tile_corner(p, sip_a, sip_b, sP, sD, x0, y0, tx, TH, graze);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        // owned only when this is the last tile row (y0 + TH == H)
        if (p.o.d_valid_k && tx == 0 && y0 + TH == H && (x0 >> 5) < wpr_k)
            p.o.d_valid_k[(size_t)H * wpr_k + (x0 >> 5)] = m;
    } else if (ty == 1) {
        bool hit = false;
        if (tx <= TH) hit = tile_corner(p, sip_a, sip_b, sP, sD, x0, y0, TW, tx, graze);
        // owned only when x0 + TW == W: then bit 0 of a word of its own (W % 32 == 0)
        if (p.o.d_valid_k && tx <= TH && x0 + TW == W && y0 + tx <= H && (tx < TH || y0 + TH == H))
            p.o.d_valid_k[(size_t)(y0 + tx) * wpr_k + (W >> 5)] = hit ? 1u : 0u;
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    bool chit = false;
    if (x < W && y < H) {
        double P[3], dir[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // corners[:-1,:-1] + corners[:-1,1:]; += corners[1:,1:]; += corners[1:,:-1]; /= 4
            double s = sP[k][ty][tx] + sP[k][ty][tx + 1];
            s = s + sP[k][ty + 1][tx + 1];
            s = s + sP[k][ty + 1][tx];
            P[k] = s * 0.25;
            double d = sD[k][ty][tx] + sD[k][ty][tx + 1];
            d = d + sD[k][ty + 1][tx + 1];
            d = d + sD[k][ty + 1][tx];
            dir[k] = d * 0.25;
        }
        const size_t i = (size_t)y * W + x;
        chit = P[0] == P[0];                        // all four corner rays hit
        if (chit) {
            const double r2 = emit_point(p, P, i, p.o.d_lat_c, p.o.d_lon_c, p.o.d_mlat_c, p.o.d_mlt_c);
            if (p.o.d_elev_c) PLANE_ST(&p.o.d_elev_c[i], elevation_deg<true>(dir, P, r2));
        } else {
            emit_nan(i, p.o.d_lat_c, p.o.d_lon_c, p.o.d_mlat_c, p.o.d_mlt_c);
            if (p.o.d_elev_c) PLANE_ST(&p.o.d_elev_c[i], qnan());
        }
    }
    __syncwarp();
    const unsigned mc = __ballot_sync(0xffffffffu, chit);
    if (p.o.d_valid_c && tx == 0 && y < H && (x0 >> 5) < wpr_c) p.o.d_valid_c[(size_t)y * wpr_c + (x0 >> 5)] = mc;
    count_grazing(p.ill, graze, tid & 31);
}

static int fill_frame(const amt_frame* fr, GeorefParams& p) {
    CHECK_ARG(fr->width > 0 && fr->height > 0, "amt_frame: width/height must be positive");
    CHECK_ARG((long long)(fr->width + 1) * (fr->height + 1) < (1LL << 31), "amt_frame: frame too large");
    CHECK_ARG(fr->sip_order_a >= 0 && fr->sip_order_a <= AMT_SIP_MAX_ORDER &&
              fr->sip_order_b >= 0 && fr->sip_order_b <= AMT_SIP_MAX_ORDER, "amt_frame: SIP order out of range");
    FrameC& f = p.f;
    f.W = fr->width; f.H = fr->height;
    f.fast_center = fr->fast_center; f.origin_inside = fr->origin_inside;
    f.crpix0 = fr->crpix[0]; f.crpix1 = fr->crpix[1];
    memcpy(f.cd, fr->cd, sizeof f.cd);
    memcpy(f.rot, fr->rot, sizeof f.rot);
    memcpy(f.cam, fr->cam, sizeof f.cam);
    memcpy(f.rad, fr->inv_axes, sizeof f.rad);
    // intersection.py:63,68,74: origin = -lineOrigin; originTimesRadius; originDotOrigin
    volatile double o0 = (-fr->cam[0]) * fr->inv_axes[0];
    volatile double o1 = (-fr->cam[1]) * fr->inv_axes[1];
    volatile double o2 = (-fr->cam[2]) * fr->inv_axes[2];
    f.otr[0] = o0; f.otr[1] = o1; f.otr[2] = o2;
    volatile double s0 = o0 * o0, s1 = o1 * o1, s2 = o2 * o2;
    volatile double s01 = s0 + s1;
    f.oDO = s01 + s2;
    memcpy(f.m_geo, fr->m_geo, sizeof f.m_geo);
    memcpy(f.m_sm, fr->m_sm, sizeof f.m_sm);
    // transform.py:254-255,290: e2 = (a*a-b*b)/(a*a); d = (a*a-b*b)/b; e2*a
    const double a = fr->wgs_a, b = fr->wgs_b;
    volatile double aa = a * a, bb = b * b;
    volatile double num = aa - bb;
    volatile double e2 = num / aa;
    f.a = a; f.b = b;
    f.b_over_a = 2.0 * (b / a);       // 2 b/a: the kernels hold 0.5/p (an exact doubling folded into the constant)
    f.e2a = e2 * a;
    f.d = num / b;
    f.sip_oa = fr->sip_order_a; f.sip_ob = fr->sip_order_b;
    CHECK_ARG(fr->model == AMT_MODEL_WCS || fr->model == AMT_MODEL_ALLSKY, "amt_frame: unknown camera model");
    CHECK_ARG(fr->model == AMT_MODEL_WCS || (fr->allsky_k > 0 && !fr->fast_center),
              "amt_frame: all-sky model needs k > 0 and fast_center == 0");
    f.model = fr->model;
    f.as_xc = fr->allsky_xc; f.as_yc = fr->allsky_yc; f.as_k = fr->allsky_k; f.as_rot = fr->allsky_rotation;
    memcpy(p.sip_a, fr->sip_a, sizeof p.sip_a);
    memcpy(p.sip_b, fr->sip_b, sizeof p.sip_b);
    fill_affine(f);
    f.sip_dx = f.sip_dy = 0.0;
    if (f.sip_oa | f.sip_ob) fill_sip_bounds(f, p.sip_a, p.sip_b);
    return AMT_OK;
}

extern "C" int amt_georef(amt_ctx* ctx, const amt_frame* frame, const amt_georef_out* out,
                          amt_stats* d_stats, void* stream) {
    ENTER(ctx);
    CHECK_ARG(frame && out, "amt_georef: NULL argument");
    GeorefParams p;
    memset(&p, 0, sizeof p);
    int rc = fill_frame(frame, p);
    if (rc) return rc;
    p.o = *out;
    p.ill = d_stats ? (unsigned long long*)&d_stats->n_ill_conditioned : nullptr;
    const int W = frame->width, H = frame->height;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_stats) CUDA_TRY(cudaMemsetAsync(&d_stats->n_ill_conditioned, 0, sizeof(uint64_t), st));
    if (frame->fast_center) {
        dim3 grid((W + TW) / TW, (H + TH) / TH);   // covers x<=W, y<=H
        k_georef_tiles<<<grid, dim3(TW, TH), 0, st>>>(p);
    } else {
        const bool want_k = out->d_lat_k || out->d_lon_k || out->d_mlat_k || out->d_mlt_k || out->d_valid_k;
        const bool want_c = out->d_lat_c || out->d_lon_c || out->d_mlat_c || out->d_mlt_c || out->d_elev_c ||
                            out->d_valid_c;
        if (!want_k && !want_c) return AMT_OK;
        dim3 grid((W + 1 + 255) / 256, want_k ? H + 1 : H);
        p.row_stride = golden_stride(grid.y);
        const bool any_plane = out->d_lat_k || out->d_lon_k || out->d_mlat_k || out->d_mlt_k || out->d_lat_c ||
                               out->d_lon_c || out->d_mlat_c || out->d_mlt_c || out->d_elev_c;
        if (!any_plane) {                          // validity bitmaps only
            const bool affine = frame->model == AMT_MODEL_WCS && frame->sip_order_a == 0 && frame->sip_order_b == 0;
            const bool solver = frame->model == AMT_MODEL_WCS && out->d_valid_k && out->d_valid_c &&
                                !getenv("AMT_NO_LIMB_SOLVER");
            if (solver && affine) {
                k_limb_bits<<<(H + 1 + 3) / 4, 128, 0, st>>>(p, out->d_valid_k, out->d_valid_c);
            } else if (solver) {
                k_limb_bits_sip<<<(H + 1 + 3) / 4, 128, 0, st>>>(p, out->d_valid_k, out->d_valid_c);
            } else {
                dim3 gh((W + 1 + 255) / 256, H + 1);
                k_hit_bits<<<gh, 256, 0, st>>>(p);
            }
            LAUNCH_CHECK(ctx);
            return AMT_OK;
        }
        const bool full = out->d_lat_k && out->d_lon_k && out->d_mlat_k && out->d_mlt_k && out->d_valid_k &&
                          out->d_lat_c && out->d_lon_c && out->d_mlat_c && out->d_mlt_c && out->d_elev_c &&
                          out->d_valid_c;
        const bool plain = frame->model == AMT_MODEL_WCS && frame->sip_order_a == 0 && frame->sip_order_b == 0;
        if (full && plain) k_georef_points<true, true, true, true><<<grid, 256, 0, st>>>(p);
        else if (full) k_georef_points<true, true, true><<<grid, 256, 0, st>>>(p);
        else if (want_k && want_c) k_georef_points<true, true, false><<<grid, 256, 0, st>>>(p);
        else if (want_k) k_georef_points<true, false, false><<<grid, 256, 0, st>>>(p);
        else k_georef_points<false, true, false><<<grid, 256, 0, st>>>(p);
    }
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// ============================================== target grid / pre-rotation (shared)
struct GridC {
    int nx, ny, prerotate;
    double lo_x, hi_x, step_x, inv_step_x, round_x, eps_x;
    double lo_y, hi_y, step_y, inv_step_y, round_y, eps_y;
    double altitude, a, b, e2, e2a, d;
    double rot[9];
    double side_scale;        // > 0: the side channel accumulates llrint(value * side_scale) in int64
};

static int fill_grid(const amt_grid* g, GridC& c, bool pre_only = false) {
    if (!pre_only) {
        CHECK_ARG(g->nx > 0 && g->ny > 0, "amt_grid: nx, ny must be positive");
        CHECK_ARG(g->hi_x > g->lo_x && g->hi_y > g->lo_y, "amt_grid: bin edges must increase");
        CHECK_ARG(g->step_x > 0 && g->step_y > 0, "amt_grid: steps must be positive");
    }
    CHECK_ARG(g->prerotate >= 0 && g->prerotate <= 2, "amt_grid: bad prerotate mode");
    c.nx = g->nx; c.ny = g->ny; c.prerotate = g->prerotate;
    c.lo_x = g->lo_x; c.hi_x = g->hi_x; c.step_x = g->step_x; c.inv_step_x = 1.0 / g->step_x; c.round_x = g->round_x;
    c.lo_y = g->lo_y; c.hi_y = g->hi_y; c.step_y = g->step_y; c.inv_step_y = 1.0 / g->step_y; c.round_y = g->round_y;
    // safety margin (in cells) of the floor shortcut of bin_index: 1e-9 plus 32x the worst-case
    // rounding of q and of the numpy edges; 1 (= shortcut off) when that is not << 1
    auto margin = [](double lo, double hi, double inv_step, int n) {
        const double u = 2.220446049250313e-16;
        const double e = 1e-9 + 32.0 * u * (fmax(fabs(lo), fabs(hi)) * inv_step + (double)n);
        return e < 1e-3 ? e : 1.0;
    };
    c.eps_x = pre_only ? 1.0 : margin(c.lo_x, c.hi_x, c.inv_step_x, c.nx);
    c.eps_y = pre_only ? 1.0 : margin(c.lo_y, c.hi_y, c.inv_step_y, c.ny);
    c.altitude = g->altitude; c.a = g->wgs_a; c.b = g->wgs_b;
    c.side_scale = g->side_scale > 0 ? g->side_scale : 0.0;
    volatile double aa = c.a * c.a, bb = c.b * c.b;
    volatile double num = aa - bb;
    volatile double e2 = c.a != 0 ? num / aa : 0;
    c.e2 = e2;
    c.e2a = e2 * c.a;
    c.d = c.b != 0 ? num / c.b : 0;
    memcpy(c.rot, g->rot, sizeof c.rot);
    return AMT_OK;
}

// resample.py:176-218: coordinates are rotated out of the pole / date line before binning.
__device__ __forceinline__ void prerotate(const GridC& g, double& la, double& lo) {
    if (g.prerotate == AMT_PRE_WRAP180) {
        lo = wrap_at_180(lo + 180.0);
    } else if (g.prerotate == AMT_PRE_POLE) {
        // transform.py:301-322 rotatePole: geodetic2Ecef -> R -> ecef2Geodetic, rad<->deg outside
        double G[3], R[3], l2, o2;
        geodetic2ecef(g.a, g.e2, la * kDeg2Rad, lo * kDeg2Rad, g.altitude, G[0], G[1], G[2]);
        mat3(g.rot, G, R);
        bowring_ref(g.a, g.b, g.e2a, g.d, R[0], R[1], R[2], l2, o2);
        la = l2 * kRad2Deg;
        lo = o2 * kRad2Deg;
    }
}

// ============================================================== sanitize + bbox stats
// Validity bitmaps: one bit per corner / centre, rows padded to whole 32-bit words
// (words per row: wpr_k = ceil((W+1)/32), wpr_c = ceil(W/32)); padding bits are 0.  The
// georeference kernels emit them for free (warp ballots); all mask logic of
// mapping/mapping.py:1063-1125 and the outline/bounding-box reductions of :655-743 then run
// on these ~1.5 MB bitmaps (L2 resident) instead of re-reading the 96 MB coordinate planes.
__device__ __forceinline__ unsigned long long dkey(double d) {       // order-preserving key
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)b);
}

struct StatKeys {
    unsigned long long lat_min, lat_max, lon_min, lon_max, lon_min_pos, lon_max_neg;
    unsigned long long n_valid_k, n_boundary, n_valid_c;
    unsigned int pole_flags;
    unsigned int blocks_done;
    unsigned int row_min, row_max, col_min, col_max;    // pixel box of the valid centres
    unsigned int queue_n;                               // outline nodes in the stream's node queue
};

__device__ __forceinline__ void stats_reset(StatKeys* s) {
    s->lat_min = s->lon_min = s->lon_min_pos = ~0ULL;
    s->lat_max = s->lon_max = s->lon_max_neg = 0ULL;
    s->n_valid_k = s->n_boundary = s->n_valid_c = 0ULL;
    s->pole_flags = 0u;
    s->blocks_done = 0u;
    s->queue_n = 0u;
    s->row_min = s->col_min = 0xffffffffu;
    s->row_max = s->col_max = 0u;
}

__global__ void k_stats_init(StatKeys* s) { stats_reset(s); }

// keys -> amt_stats, then the key block is reset: it is clean for the next call (no init launch)
__device__ __forceinline__ void stats_finish(StatKeys* s, amt_stats* out) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    volatile StatKeys* v = s;
    const unsigned long long a = v->lat_min, b = v->lon_min, c = v->lon_min_pos;
    const unsigned long long d = v->lat_max, e = v->lon_max, f = v->lon_max_neg;
    out->lat_min = a == ~0ULL ? inf : dunkey(a);
    out->lon_min = b == ~0ULL ? inf : dunkey(b);
    out->lon_min_pos = c == ~0ULL ? inf : dunkey(c);
    out->lat_max = d == 0ULL ? -inf : dunkey(d);
    out->lon_max = e == 0ULL ? -inf : dunkey(e);
    out->lon_max_neg = f == 0ULL ? -inf : dunkey(f);
    out->n_valid_corners = v->n_valid_k;
    out->n_boundary_corners = v->n_boundary;
    out->n_valid_centers = v->n_valid_c;
    out->pole_flags = v->pole_flags;
    const unsigned r0 = v->row_min, r1 = v->row_max, c0 = v->col_min, c1 = v->col_max;
    const bool any = r0 != 0xffffffffu;
    out->row_min_c = any ? (int)r0 : 0;
    out->row_max_c = any ? (int)r1 : -1;
    out->col_min_c = any ? (int)c0 : 0;
    out->col_max_c = any ? (int)c1 : -1;
    stats_reset(s);
}

__device__ __forceinline__ bool isnan_d(double v) { return !(v == v); }

struct Bits {                 // a row-padded bitmap
    const unsigned* w;
    int wpr, rows;
    __device__ __forceinline__ unsigned at(int y, int i) const {
        return (y < 0 || y >= rows || i < 0 || i >= wpr) ? 0u : w[(size_t)y * wpr + i];
    }
    // bit x of the result = bit (x-1) / (x+1) of row y
    __device__ __forceinline__ unsigned from_left(int y, int i) const { return (at(y, i) << 1) | (at(y, i - 1) >> 31); }
    __device__ __forceinline__ unsigned from_right(int y, int i) const { return (at(y, i) >> 1) | (at(y, i + 1) << 31); }
};

// bitmaps from NaN-marked planes (generic mappings, after masking)
__global__ void __launch_bounds__(256) k_valid_bits(int W, int H, const double* __restrict__ lat_k,
                                                    const double* __restrict__ lat_c, unsigned* __restrict__ bk,
                                                    unsigned* __restrict__ bc, int wpr_k, int wpr_c) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const bool corner = (int)blockIdx.y <= H;
    const int y = corner ? blockIdx.y : blockIdx.y - (H + 1);
    const int rowlen = corner ? W + 1 : W;
    const double* src = corner ? lat_k : lat_c;
    const bool v = x < rowlen && !isnan_d(src[(size_t)y * rowlen + x]);
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < (corner ? wpr_k : wpr_c))
        (corner ? bk : bc)[(size_t)y * (corner ? wpr_k : wpr_c) + (x >> 5)] = m;
}

__device__ __forceinline__ void nan_centers(const amt_georef_out& o, size_t base, unsigned bits) {
    const double nan = qnan();
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const size_t i = base + b;
        if (o.d_lat_c) o.d_lat_c[i] = nan;
        if (o.d_lon_c) o.d_lon_c[i] = nan;
        if (o.d_mlat_c) o.d_mlat_c[i] = nan;
        if (o.d_mlt_c) o.d_mlt_c[i] = nan;
        if (o.d_elev_c) o.d_elev_c[i] = nan;
    }
}
__device__ __forceinline__ void nan_corners(const amt_georef_out& o, size_t base, unsigned bits) {
    const double nan = qnan();
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const size_t i = base + b;
        if (o.d_lat_k) o.d_lat_k[i] = nan;
        if (o.d_lon_k) o.d_lon_k[i] = nan;
        if (o.d_mlat_k) o.d_mlat_k[i] = nan;
        if (o.d_mlt_k) o.d_mlt_k[i] = nan;
    }
}

// Corner rule alone (after masking, mapping.py:1082-1093 with afterMasking=True): a corner stays
// valid only if one of its neighbouring centres is; corners that lost validity get NaN.
__global__ void k_bits_corner_final(int W, int H, Bits K1, Bits C1, unsigned* __restrict__ k0, amt_georef_out o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (i >= K1.wpr || y > H) return;
    const unsigned anyc = C1.from_left(y - 1, i) | C1.at(y - 1, i) | C1.from_left(y, i) | C1.at(y, i);
    const unsigned k2 = K1.at(y, i) & anyc;
    const unsigned old = k0[(size_t)y * K1.wpr + i];
    if (k2 != old) {
        k0[(size_t)y * K1.wpr + i] = k2;
        nan_corners(o, (size_t)y * (W + 1) + 32 * i, old & ~k2);
    }
}

// All three steps of _doSanitize (mapping.py:1082-1117) in one launch:
//   1. a corner stays valid only if at least one of its (<=4) neighbouring centres is valid,
//   2. a centre stays valid only if all 4 corners are (still) valid,
//   3. rule 1 once more with the updated centres;
// elements that lose validity get NaN in every plane.  K0 / C0 are COPIES of the input bitmaps (the
// results overwrite the originals while neighbouring words are still being read); every thread
// owns one (row, word) position of both bitmaps and recomputes the intermediate words it needs
// from the 4 x 4 (C0) / 3 x 3 (K0) neighbourhood -- a few hundred bit operations on an
// L2-resident 1.5 MB working set instead of two more launches.
struct SanitizeWords {
    Bits K0, C0;
    // step 1: corner valid and one of its neighbouring centres valid
    __device__ __forceinline__ unsigned k1(int y, int i) const {
        const unsigned anyc = C0.from_left(y - 1, i) | C0.at(y - 1, i) | C0.from_left(y, i) | C0.at(y, i);
        return K0.at(y, i) & anyc;
    }
};

// Launch shape: grid-stride over the (row, word) positions with a few hundred CTAs.  The work is a few
// hundred bit operations per word on an L2-resident working set, i.e. the kernel is latency bound, and in
// the sequence engine it runs BESIDE the long fused kernel of other frames: what it costs there is the
// residency of its CTAs (each one displaces a CTA of the fused kernel for its lifetime), so it uses few
// CTAs that each do several positions -- 5666 CTAs of 128 threads cost the fused kernel 8 us per frame
// (AMT_SEQ_SKIP timeline, profiles/r02_stage_a_interference.txt).
// WRITE_ALL: every word of the result is stored (kout / cout are different buffers than K0 / C0: no copy
// of the input needed); otherwise only the words that change (in place on a copy of the input).
template <bool WRITE_ALL>
__global__ void __launch_bounds__(256) k_sanitize_fused(int W, int H, Bits K0, Bits C0, unsigned* __restrict__ kout,
                                                        unsigned* __restrict__ cout, amt_georef_out o) {
    const SanitizeWords sw{K0, C0};
    const int total = K0.wpr * (H + 1);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int y = t / K0.wpr, i = t - y * K0.wpr;
        // k1 on rows y-1 .. y+1, words i-1 .. i+1
        unsigned k1[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) k1[r][c] = sw.k1(y - 1 + r, i - 1 + c);
        // step 2: c1 on rows y-1 .. y, words i-1 .. i  (all four corners valid in k1)
        unsigned c1[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const unsigned a = k1[r][c], a_r = (a >> 1) | (k1[r][c + 1] << 31);
                const unsigned b = k1[r + 1][c], b_r = (b >> 1) | (k1[r + 1][c + 1] << 31);
                c1[r][c] = C0.at(y - 1 + r, i - 1 + c) & a & a_r & b & b_r;
            }
        // step 3: corners once more against the updated centres
        const unsigned up_l = (c1[0][1] << 1) | (c1[0][0] >> 31), dn_l = (c1[1][1] << 1) | (c1[1][0] >> 31);
        const unsigned k2 = k1[1][1] & (up_l | c1[0][1] | dn_l | c1[1][1]);
        const unsigned k_old = K0.at(y, i);
        if (WRITE_ALL || k2 != k_old) {
            kout[(size_t)y * K0.wpr + i] = k2;
            if (k2 != k_old) nan_corners(o, (size_t)y * (W + 1) + 32 * i, k_old & ~k2);
        }
        if (y < H && i < C0.wpr) {
            const unsigned c_old = C0.at(y, i), c_new = c1[1][1];
            if (WRITE_ALL || c_new != c_old) {
                cout[(size_t)y * C0.wpr + i] = c_new;
                if (c_new != c_old) nan_centers(o, (size_t)y * W + 32 * i, c_old & ~c_new);
            }
        }
    }
}

// The same three steps with the intermediate words kept in shared memory: a CTA owns kSanRows full rows, forms
// k1 for its rows + 1 halo row either side, c1 for its rows + 1 above, then the results -- each intermediate
// word is computed ~1.1 times instead of 9 (k1) / 4 (c1) times as in the register-only version above:
// 6.5 M -> 3.7 M warp instructions per 12-Mpix frame (profiles/r02_stagea_ncu.txt), which is what the kernel
// takes away from the fused kernel it runs beside in the sequence engine.  Dynamic shared memory:
// (2 kSanRows + 3) (wpr + 2) words; frames too wide for 48 KB use the register-only kernel.
constexpr int kSanRows = 14;
template <bool WRITE_ALL>
__global__ void __launch_bounds__(256) k_sanitize_tile(int W, int H, Bits K0, Bits C0, unsigned* __restrict__ kout,
                                                       unsigned* __restrict__ cout, amt_georef_out o) {
    extern __shared__ unsigned s_san[];
    const SanitizeWords sw{K0, C0};
    const int wk = K0.wpr, pitch = wk + 2;
    unsigned* s_k1 = s_san;                               // rows y0-1 .. y0+kSanRows, words -1 .. wk
    unsigned* s_c1 = s_san + (kSanRows + 2) * pitch;      // rows y0-1 .. y0+kSanRows-1, words -1 .. wk
    const int y0 = blockIdx.x * kSanRows;
    for (int idx = threadIdx.x; idx < (kSanRows + 2) * pitch; idx += 256) {
        const int r = idx / pitch, c = idx - r * pitch;
        s_k1[idx] = sw.k1(y0 - 1 + r, c - 1);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < (kSanRows + 1) * pitch; idx += 256) {
        const int r = idx / pitch, c = idx - r * pitch;
        unsigned v = 0u;
        if (c + 1 < pitch) {                              // word wk has no right neighbour in the tile and is never used
            const unsigned a = s_k1[r * pitch + c], a_r = (a >> 1) | (s_k1[r * pitch + c + 1] << 31);
            const unsigned b = s_k1[(r + 1) * pitch + c], b_r = (b >> 1) | (s_k1[(r + 1) * pitch + c + 1] << 31);
            v = C0.at(y0 - 1 + r, c - 1) & a & a_r & b & b_r;
        }
        s_c1[idx] = v;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kSanRows * wk; idx += 256) {
        const int r = idx / wk, i = idx - r * wk, y = y0 + r;
        if (y > H) break;
        // c1 rows: r -> y-1, r+1 -> y; columns: i -> word i-1, i+1 -> word i
        const unsigned c_up = s_c1[r * pitch + i + 1], c_up_l = s_c1[r * pitch + i];
        const unsigned c_dn = s_c1[(r + 1) * pitch + i + 1], c_dn_l = s_c1[(r + 1) * pitch + i];
        const unsigned up_l = (c_up << 1) | (c_up_l >> 31), dn_l = (c_dn << 1) | (c_dn_l >> 31);
        const unsigned k2 = s_k1[(r + 1) * pitch + i + 1] & (up_l | c_up | dn_l | c_dn);
        const unsigned k_old = K0.at(y, i);
        if (WRITE_ALL || k2 != k_old) {
            kout[(size_t)y * wk + i] = k2;
            if (k2 != k_old) nan_corners(o, (size_t)y * (W + 1) + 32 * i, k_old & ~k2);
        }
        if (y < H && i < C0.wpr) {
            const unsigned c_old = C0.at(y, i);
            if (WRITE_ALL || c_dn != c_old) {
                cout[(size_t)y * C0.wpr + i] = c_dn;
                if (c_dn != c_old) nan_centers(o, (size_t)y * W + 32 * i, c_old & ~c_dn);
            }
        }
    }
}

template <bool WRITE_ALL>
static void launch_sanitize(const amt_ctx* ctx, cudaStream_t st, int W, int H, const Bits& K0, const Bits& C0,
                            unsigned* kout, unsigned* cout, const amt_georef_out& o);

static inline int sanitize_blocks(const amt_ctx* ctx, size_t words) {
    static const int per_sm = getenv("AMT_SANITIZE_BLOCKS_PER_SM") ? atoi(getenv("AMT_SANITIZE_BLOCKS_PER_SM")) : 2;
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->sm_count * per_sm, (words + 255) / 256));
}

template <bool WRITE_ALL>
static void launch_sanitize(const amt_ctx* ctx, cudaStream_t st, int W, int H, const Bits& K0, const Bits& C0,
                            unsigned* kout, unsigned* cout, const amt_georef_out& o) {
    const size_t smem = (size_t)(2 * kSanRows + 3) * (K0.wpr + 2) * sizeof(unsigned);
    if (smem <= 48 * 1024 && !getenv("AMT_SANITIZE_REGISTERS")) {
        k_sanitize_tile<WRITE_ALL><<<(H + 1 + kSanRows - 1) / kSanRows, 256, smem, st>>>(W, H, K0, C0, kout, cout, o);
    } else {
        const size_t nk = (size_t)K0.wpr * (H + 1);
        k_sanitize_fused<WRITE_ALL><<<sanitize_blocks(ctx, nk), 256, 0, st>>>(W, H, K0, C0, kout, cout, o);
    }
}

__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// Bounding box over the outline (valid corners with an invalid / out-of-array 4-neighbour,
// mapping.py:672-703,729-737) + valid counts, from the bitmaps.
// Outline coordinates either from the planes (lat_k != NULL) or recomputed from the frame model for the few
// outline nodes -- the plane-free (fused) resampling path and the sequence engine, where this kernel runs
// BEFORE the planes exist.
//
// Two launches.  k_outline_collect (<= 1 CTA per SM; four independent bitmap words per thread and step,
// 20 loads in flight) classifies the words, counts, and COMPACTS the outline nodes into a queue in global
// memory (one atomic per warp that found any).  k_outline_eval evaluates the queue densely, one node per
// thread, and the last of its CTAs converts the keys.  An ISS frame has ~14 k outline nodes, 10 k of them
// one per bitmap word (left / right frame edge) and 4 k in one row (bottom edge): evaluated word by word
// (32 lanes on the 32 bits of a word: the first version) that was ~6000 warp-level passes of the
// ~400-instruction FP64 chain, most with one active lane, in 592 long-lived CTAs, which cost the fused
// kernel they ran beside 10 us per frame (AMT_SEQ_SKIP timeline, profiles/r02_stage_a_interference.txt);
// compacted per CTA it is unbalanced (the bottom row lands in one CTA); from the global queue it is ~55
// CTAs with one node per thread.
constexpr unsigned kStatQueue = 1u << 20;        // nodes (4 MB per stream); the excess is evaluated in place

struct OutlineAcc {
    unsigned long long mn_la = ~0ULL, mx_la = 0ULL, mn_lo = ~0ULL, mx_lo = 0ULL, mn_pos = ~0ULL, mx_neg = 0ULL;
    __device__ __forceinline__ void add(const GridC& g, double a_, double o_) {
        if (g.prerotate != AMT_PRE_NONE) prerotate(g, a_, o_);
        const unsigned long long kla = dkey(a_), klo = dkey(o_);
        mn_la = umin64(mn_la, kla); mx_la = umax64(mx_la, kla);
        mn_lo = umin64(mn_lo, klo); mx_lo = umax64(mx_lo, klo);
        if (o_ > 0.0) mn_pos = umin64(mn_pos, klo); else mx_neg = umax64(mx_neg, klo);
    }
    __device__ __forceinline__ void flush(StatKeys* s) const {
        if (mn_la == ~0ULL) return;
        atomicMin(&s->lat_min, mn_la); atomicMax(&s->lat_max, mx_la);
        atomicMin(&s->lon_min, mn_lo); atomicMax(&s->lon_max, mx_lo);
        if (mn_pos != ~0ULL) atomicMin(&s->lon_min_pos, mn_pos);
        if (mx_neg != 0ULL) atomicMax(&s->lon_max_neg, mx_neg);
    }
};

// coordinates of corner node `idx` (flat index into the (H+1) x (W+1) corner array)
template <bool FRAME>
__device__ __noinline__ void outline_coord(const GeorefParams& frame, const double* __restrict__ lat_k,
                                           const double* __restrict__ lon_k, int W, unsigned idx, double& la,
                                           double& lo) {
    if (FRAME) {
        const int y = (int)(idx / (unsigned)(W + 1)), x = (int)(idx - (unsigned)y * (unsigned)(W + 1));
        double dir[3], dcen[3], P[3];
        bool gz;
        dirs_kc<true>(frame.f, frame.sip_a, frame.sip_b, x, y, dir, dcen);
        intersect(frame.f, dir, P, gz);
        point_to_geo(frame.f, P, la, lo);
    } else {
        la = lat_k[idx];
        lo = lon_k[idx];
    }
}

template <bool FRAME>
__global__ void __launch_bounds__(256, 4) k_outline_collect(int W, int H, Bits K, Bits C,
                                                            const double* __restrict__ lat_k,
                                                            const double* __restrict__ lon_k,
                                                            const __grid_constant__ GridC g, StatKeys* s,
                                                            const __grid_constant__ GeorefParams frame,
                                                            unsigned* __restrict__ queue) {
    OutlineAcc acc;                                      // nodes beyond the queue capacity only
    unsigned nvk = 0, nb = 0, nvc = 0;
    const int lane = threadIdx.x & 31;
    constexpr int U = 4;
    const int nwk = K.wpr * (H + 1);
    for (int base = blockIdx.x * (256 * U); base < nwk; base += gridDim.x * (256 * U)) {
        unsigned v[U], lf[U], rt[U], up[U], dn[U];
        int yy[U], ii[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {                    // all loads of the step first
            const int t = base + u * 256 + threadIdx.x;
            const bool in = t < nwk;
            const int y = in ? t / K.wpr : 0, i = in ? t - y * K.wpr : 0;
            yy[u] = y; ii[u] = i;
            v[u] = in ? K.w[t] : 0u;
            lf[u] = in ? K.at(y, i - 1) : 0u;
            rt[u] = in ? K.at(y, i + 1) : 0u;
            up[u] = in ? K.at(y - 1, i) : 0u;
            dn[u] = in ? K.at(y + 1, i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            unsigned b = 0;
            if (v[u]) {
                nvk += __popc(v[u]);
                const unsigned interior = v[u] & ((v[u] << 1) | (lf[u] >> 31)) & ((v[u] >> 1) | (rt[u] << 31)) & up[u] & dn[u];
                b = v[u] & ~interior;
                nb += __popc(b);
            }
            if (__ballot_sync(0xffffffffu, b != 0) == 0) continue;           // warp-uniform
            // warp-aggregated append: one atomic per warp
            const unsigned n = __popc(b);
            unsigned incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t_ = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t_;
            }
            unsigned wbase = 0;
            if (lane == 31) wbase = atomicAdd(&s->queue_n, incl);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            unsigned off = wbase + incl - n;
            const unsigned first = (unsigned)yy[u] * (unsigned)(W + 1) + 32u * (unsigned)ii[u];
            while (b) {
                const unsigned idx = first + (unsigned)(__ffs(b) - 1);
                b &= b - 1;
                if (off < kStatQueue) {
                    queue[off] = idx;
                } else {                                 // queue full (pathological masks): evaluate in place
                    double la, lo;
                    outline_coord<FRAME>(frame, lat_k, lon_k, W, idx, la, lo);
                    acc.add(g, la, lo);
                }
                ++off;
            }
        }
    }
    acc.flush(s);
    const int nwc = C.wpr * H;
    unsigned r_lo = 0xffffffffu, r_hi = 0u, c_lo = 0xffffffffu, c_hi = 0u;
    for (int base = blockIdx.x * (256 * U); base < nwc; base += gridDim.x * (256 * U)) {
        unsigned v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = base + u * 256 + threadIdx.x;
            v[u] = t < nwc ? C.w[t] : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!v[u]) continue;
            const int t = base + u * 256 + threadIdx.x;
            nvc += __popc(v[u]);
            const unsigned y = t / C.wpr, i = t - y * C.wpr;
            r_lo = min(r_lo, y); r_hi = max(r_hi, y);
            c_lo = min(c_lo, 32u * i + (unsigned)(__ffs(v[u]) - 1));
            c_hi = max(c_hi, 32u * i + (unsigned)(31 - __clz(v[u])));
        }
    }
    r_lo = __reduce_min_sync(0xffffffffu, r_lo); r_hi = __reduce_max_sync(0xffffffffu, r_hi);
    c_lo = __reduce_min_sync(0xffffffffu, c_lo); c_hi = __reduce_max_sync(0xffffffffu, c_hi);
    nvk = __reduce_add_sync(0xffffffffu, nvk);
    nb = __reduce_add_sync(0xffffffffu, nb);
    nvc = __reduce_add_sync(0xffffffffu, nvc);
    __shared__ unsigned shc[3][8];
    __shared__ unsigned shr[4][8];
    const int warp = threadIdx.x >> 5;
    if (lane == 0) {
        shc[0][warp] = nvk; shc[1][warp] = nb; shc[2][warp] = nvc;
        shr[0][warp] = r_lo; shr[1][warp] = r_hi; shr[2][warp] = c_lo; shr[3][warp] = c_hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            shc[0][0] += shc[0][w]; shc[1][0] += shc[1][w]; shc[2][0] += shc[2][w];
            shr[0][0] = min(shr[0][0], shr[0][w]); shr[1][0] = max(shr[1][0], shr[1][w]);
            shr[2][0] = min(shr[2][0], shr[2][w]); shr[3][0] = max(shr[3][0], shr[3][w]);
        }
        if (shr[0][0] != 0xffffffffu) {               // one set of atomics per block (same-address atomics are slow)
            atomicMin(&s->row_min, shr[0][0]); atomicMax(&s->row_max, shr[1][0]);
            atomicMin(&s->col_min, shr[2][0]); atomicMax(&s->col_max, shr[3][0]);
        }
        if (shc[0][0]) atomicAdd(&s->n_valid_k, (unsigned long long)shc[0][0]);
        if (shc[2][0]) atomicAdd(&s->n_valid_c, (unsigned long long)shc[2][0]);
        if (shc[1][0]) atomicAdd(&s->n_boundary, (unsigned long long)shc[1][0]);
    }
}

// One queued outline node per thread; CTAs beyond the queue's length leave at once.  The last CTA to
// arrive converts the keys and leaves the key block (and the queue counter) clean for the next call.
template <bool FRAME>
__global__ void __launch_bounds__(256, 4) k_outline_eval(int W, const double* __restrict__ lat_k,
                                                         const double* __restrict__ lon_k,
                                                         const __grid_constant__ GridC g, StatKeys* s,
                                                         const __grid_constant__ GeorefParams frame,
                                                         const unsigned* __restrict__ queue, amt_stats* out) {
    const unsigned n = min(*(volatile unsigned*)&s->queue_n, kStatQueue);
    OutlineAcc acc;
    for (unsigned j = blockIdx.x * 256 + threadIdx.x; j < n; j += gridDim.x * 256) {
        double la, lo;
        outline_coord<FRAME>(frame, lat_k, lon_k, W, queue[j], la, lo);
        acc.add(g, la, lo);
    }
    unsigned long long mn_la = acc.mn_la, mx_la = acc.mx_la, mn_lo = acc.mn_lo, mx_lo = acc.mx_lo, mn_pos = acc.mn_pos,
                       mx_neg = acc.mx_neg;
    __shared__ unsigned long long sh[6][8];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        mn_la = umin64(mn_la, __shfl_xor_sync(0xffffffffu, mn_la, o));
        mx_la = umax64(mx_la, __shfl_xor_sync(0xffffffffu, mx_la, o));
        mn_lo = umin64(mn_lo, __shfl_xor_sync(0xffffffffu, mn_lo, o));
        mx_lo = umax64(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, o));
        mn_pos = umin64(mn_pos, __shfl_xor_sync(0xffffffffu, mn_pos, o));
        mx_neg = umax64(mx_neg, __shfl_xor_sync(0xffffffffu, mx_neg, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sh[0][warp] = mn_la; sh[1][warp] = mx_la; sh[2][warp] = mn_lo; sh[3][warp] = mx_lo;
        sh[4][warp] = mn_pos; sh[5][warp] = mx_neg;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        OutlineAcc t;
        for (int w = 0; w < 8; ++w) {
            t.mn_la = umin64(t.mn_la, sh[0][w]); t.mx_la = umax64(t.mx_la, sh[1][w]);
            t.mn_lo = umin64(t.mn_lo, sh[2][w]); t.mx_lo = umax64(t.mx_lo, sh[3][w]);
            t.mn_pos = umin64(t.mn_pos, sh[4][w]); t.mx_neg = umax64(t.mx_neg, sh[5][w]);
        }
        t.flush(s);
        __threadfence();
        if (atomicAdd(&s->blocks_done, 1u) == gridDim.x - 1) {
            __threadfence();
            stats_finish(s, out);
        }
    }
}

static inline int stats_blocks(const amt_ctx* ctx, int words) {
    static const int per_sm = getenv("AMT_STATS_BLOCKS_PER_SM") ? atoi(getenv("AMT_STATS_BLOCKS_PER_SM")) : 1;
    return std::max(1, std::min(ctx->sm_count * per_sm, (words + 1023) / 1024));
}
constexpr int kOutlineEvalBlocks = 64;           // 16 k nodes with one node per thread; longer queues loop

// Pole test for mappings without a camera model (GenericMapping): the longitudes of the 4
// corners of a valid pixel wind once around (+-360 deg) iff the quad encloses a pole.
// Replaces the outline / azimuth-sum test of mapping.py:705-718, geodesic.py:112-202.
__global__ void __launch_bounds__(256) k_pole_test(int W, int H, Bits C, const double* __restrict__ lat_k,
                                                   const double* __restrict__ lon_k, StatKeys* s) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    if (!((C.at(y, x >> 5) >> (x & 31)) & 1u)) return;
    const size_t k = (size_t)y * (W + 1) + x;
    const double l0 = lon_k[k], l1 = lon_k[k + 1], l2 = lon_k[k + W + 2], l3 = lon_k[k + W + 1];
    const double dl[4] = {l1 - l0, l2 - l1, l3 - l2, l0 - l3};
    double w = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double d = dl[q];
        if (d > 180.0) d -= 360.0;
        if (d <= -180.0) d += 360.0;
        w += d;
    }
    if (fabs(w) > 180.0) atomicOr(&s->pole_flags, lat_k[k] > 0.0 ? 1u : 2u);
}

static inline int wpr_of(int n) { return (n + 31) / 32; }

extern "C" int amt_valid_bits(amt_ctx* ctx, int32_t W, int32_t H, const double* d_lat_k, const double* d_lat_c,
                              uint32_t* d_valid_k, uint32_t* d_valid_c, void* stream) {
    ENTER(ctx);
    CHECK_ARG(W > 0 && H > 0 && d_lat_k && d_lat_c && d_valid_k && d_valid_c, "amt_valid_bits: bad arguments");
    dim3 grid((W + 1 + 255) / 256, 2 * H + 1);
    k_valid_bits<<<grid, 256, 0, (cudaStream_t)stream>>>(W, H, d_lat_k, d_lat_c, d_valid_k, d_valid_c,
                                                          wpr_of(W + 1), wpr_of(W));
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

extern "C" int amt_sanitize(amt_ctx* ctx, int32_t W, int32_t H, const amt_georef_out* planes, void* stream) {
    ENTER(ctx);
    CHECK_ARG(planes && W > 0 && H > 0, "amt_sanitize: bad arguments");
    CHECK_ARG(planes->d_valid_k && planes->d_valid_c, "amt_sanitize: validity bitmaps are required");
    cudaStream_t st = (cudaStream_t)stream;
    const int wk = wpr_of(W + 1), wc = wpr_of(W);
    const size_t nk = (size_t)wk * (H + 1), nc = (size_t)wc * H;
    void* scratch = nullptr;
    int rc = ensure_scratch(ctx, st, (nk + nc) * 4, &scratch);
    if (rc) return rc;
    // the kernel reads copies of the input bitmaps and overwrites the originals
    unsigned* k0 = (unsigned*)scratch;
    unsigned* c0 = k0 + nk;
    if (planes->d_valid_c == planes->d_valid_k + nk) {           // allocated back to back: one copy
        CUDA_TRY(cudaMemcpyAsync(k0, planes->d_valid_k, (nk + nc) * 4, cudaMemcpyDeviceToDevice, st));
    } else {
        CUDA_TRY(cudaMemcpyAsync(k0, planes->d_valid_k, nk * 4, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c0, planes->d_valid_c, nc * 4, cudaMemcpyDeviceToDevice, st));
    }
    Bits K0{k0, wk, H + 1}, C0{c0, wc, H};
    launch_sanitize<false>(ctx, st, W, H, K0, C0, planes->d_valid_k, planes->d_valid_c, *planes);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// Final validity bitmaps of a WCS frame straight from its header (stage A of the sequence engine): the raw
// hit bitmaps go to the stream's scratch buffer, the sanitisation stencils read them there and write
// every word of the final bitmaps -- no copy of the input, two launches.
static int hit_bits_sanitized(amt_ctx* ctx, const amt_frame* frame, uint32_t* d_valid_k, uint32_t* d_valid_c,
                              amt_stats* d_stats, cudaStream_t st) {
    const int W = frame->width, H = frame->height;
    const int wk = wpr_of(W + 1), wc = wpr_of(W);
    const size_t nk = (size_t)wk * (H + 1), nc = (size_t)wc * H;
    void* scratch = nullptr;
    int rc = ensure_scratch(ctx, st, (nk + nc) * 4, &scratch);
    if (rc) return rc;
    amt_georef_out raw;
    memset(&raw, 0, sizeof raw);
    raw.d_valid_k = (uint32_t*)scratch;
    raw.d_valid_c = raw.d_valid_k + nk;
    rc = amt_georef(ctx, frame, &raw, d_stats, st);
    if (rc) return rc;
    Bits K0{raw.d_valid_k, wk, H + 1}, C0{raw.d_valid_c, wc, H};
    amt_georef_out none;
    memset(&none, 0, sizeof none);
    launch_sanitize<true>(ctx, st, W, H, K0, C0, d_valid_k, d_valid_c, none);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

static int get_stat_keys(amt_ctx* ctx, cudaStream_t st, StatKeys** keys) {
    Workspace& w = workspace(ctx, st);
    if (!w.stat_keys) {
        CUDA_TRY(cudaMalloc(&w.stat_queue, (size_t)kStatQueue * sizeof(unsigned)));
        CUDA_TRY(cudaMalloc(&w.stat_keys, sizeof(StatKeys)));
        k_stats_init<<<1, 1, 0, st>>>((StatKeys*)w.stat_keys);         // once per stream; afterwards self-cleaning
        LAUNCH_CHECK(ctx);
    }
    *keys = (StatKeys*)w.stat_keys;
    return AMT_OK;
}

extern "C" int amt_bbox_stats_frame(amt_ctx* ctx, const amt_frame* frame, const uint32_t* d_valid_k,
                                    const uint32_t* d_valid_c, const amt_grid* pre, amt_stats* d_stats,
                                    void* stream) {
    ENTER(ctx);
    CHECK_ARG(frame && d_valid_k && d_valid_c && d_stats, "amt_bbox_stats_frame: bad arguments");
    CHECK_ARG(frame->model == AMT_MODEL_WCS, "amt_bbox_stats_frame: WCS frames only");
    cudaStream_t st = (cudaStream_t)stream;
    GeorefParams p;
    memset(&p, 0, sizeof p);
    int rc = fill_frame(frame, p);
    if (rc) return rc;
    GridC g;
    memset(&g, 0, sizeof g);
    if (pre) {
        rc = fill_grid(pre, g, true);
        if (rc) return rc;
    }
    const int W = frame->width, H = frame->height;
    const int wk = wpr_of(W + 1), wc = wpr_of(W);
    StatKeys* keys;
    rc = get_stat_keys(ctx, st, &keys);
    if (rc) return rc;
    Bits K{d_valid_k, wk, H + 1}, C{d_valid_c, wc, H};
    const int words = wk * (H + 1);
    // the frame constants travel as a kernel parameter (constant bank), like in the georeference kernels
    unsigned* queue = (unsigned*)workspace(ctx, st).stat_queue;
    k_outline_collect<true><<<stats_blocks(ctx, words), 256, 0, st>>>(W, H, K, C, nullptr, nullptr, g, keys, p, queue);
    LAUNCH_CHECK(ctx);
    k_outline_eval<true><<<kOutlineEvalBlocks, 256, 0, st>>>(W, nullptr, nullptr, g, keys, p, queue, d_stats);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

extern "C" int amt_bbox_stats(amt_ctx* ctx, int32_t W, int32_t H, const double* d_lat_k, const double* d_lon_k,
                              const uint32_t* d_valid_k, const uint32_t* d_valid_c, int32_t pole_test,
                              const amt_grid* pre, amt_stats* d_stats, void* stream) {
    ENTER(ctx);
    CHECK_ARG(W > 0 && H > 0 && d_lat_k && d_lon_k && d_valid_k && d_valid_c && d_stats, "amt_bbox_stats: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    GridC g;
    memset(&g, 0, sizeof g);
    if (pre) {
        int rc = fill_grid(pre, g, true);
        if (rc) return rc;
    }
    const int wk = wpr_of(W + 1), wc = wpr_of(W);
    StatKeys* keys;
    int rc = get_stat_keys(ctx, st, &keys);
    if (rc) return rc;
    Bits K{d_valid_k, wk, H + 1}, C{d_valid_c, wc, H};
    if (pole_test) {                     // ORs into the key block; converted and reset by k_outline_eval
        dim3 grid((W + 255) / 256, H);
        k_pole_test<<<grid, 256, 0, st>>>(W, H, C, d_lat_k, d_lon_k, keys);
        LAUNCH_CHECK(ctx);
    }
    const int words = wk * (H + 1);
    static const GeorefParams no_frame = {};
    unsigned* queue = (unsigned*)workspace(ctx, st).stat_queue;
    k_outline_collect<false><<<stats_blocks(ctx, words), 256, 0, st>>>(W, H, K, C, d_lat_k, d_lon_k, g, keys, no_frame, queue);
    LAUNCH_CHECK(ctx);
    k_outline_eval<false><<<kOutlineEvalBlocks, 256, 0, st>>>(W, d_lat_k, d_lon_k, g, keys, no_frame, queue, d_stats);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// createMasked / maskedByElevation (mapping.py:845-864,1171-1231) on the device planes.
__global__ void __launch_bounds__(256) k_apply_center_mask(int W, int H, const unsigned char* __restrict__ mask,
                                                           double min_elev, unsigned* __restrict__ cbits, int wpr_c,
                                                           amt_georef_out o) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    bool keep = false;
    if (x < W) {
        const size_t i = (size_t)y * W + x;
        keep = (cbits[(size_t)y * wpr_c + (x >> 5)] >> (x & 31)) & 1u;
        bool m = mask && mask[i];
        if (keep && min_elev == min_elev) m |= !(o.d_elev_c[i] >= min_elev);   // (elevation < min).filled(True)
        if (keep && m) {
            keep = false;
            const double nan = qnan();
            if (o.d_lat_c) o.d_lat_c[i] = nan;
            if (o.d_lon_c) o.d_lon_c[i] = nan;
            if (o.d_mlat_c) o.d_mlat_c[i] = nan;
            if (o.d_mlt_c) o.d_mlt_c[i] = nan;
            if (o.d_elev_c) o.d_elev_c[i] = nan;
        }
    }
    __syncwarp();
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < wpr_c) cbits[(size_t)y * wpr_c + (x >> 5)] = b;
}

extern "C" int amt_apply_center_mask(amt_ctx* ctx, int32_t W, int32_t H, const uint8_t* d_mask,
                                     double min_elevation, const amt_georef_out* planes, void* stream) {
    ENTER(ctx);
    CHECK_ARG(planes && W > 0 && H > 0, "amt_apply_center_mask: bad arguments");
    CHECK_ARG(planes->d_valid_k && planes->d_valid_c, "amt_apply_center_mask: validity bitmaps are required");
    CHECK_ARG(!(min_elevation == min_elevation) || planes->d_elev_c, "amt_apply_center_mask: elevation plane required");
    cudaStream_t st = (cudaStream_t)stream;
    const int wk = wpr_of(W + 1), wc = wpr_of(W);
    dim3 grid((W + 255) / 256, H);
    k_apply_center_mask<<<grid, 256, 0, st>>>(W, H, d_mask, min_elevation, planes->d_valid_c, wc, *planes);
    LAUNCH_CHECK(ctx);
    // afterMasking=True: only "corner needs a valid neighbouring centre" (mapping.py:1082-1093)
    Bits K{planes->d_valid_k, wk, H + 1}, C{planes->d_valid_c, wc, H};
    dim3 gk((wk + 127) / 128, H + 1);
    k_bits_corner_final<<<gk, 128, 0, st>>>(W, H, K, C, planes->d_valid_k, *planes);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

__global__ void k_rotate_coords(double* __restrict__ lat, double* __restrict__ lon, size_t n,
                                const __grid_constant__ GridC g) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double la = lat[i], lo = lon[i];
    prerotate(g, la, lo);
    lat[i] = la;
    lon[i] = lo;
}

extern "C" int amt_rotate_coords(amt_ctx* ctx, double* d_lat, double* d_lon, size_t n, const amt_grid* pre,
                                 void* stream) {
    ENTER(ctx);
    CHECK_ARG(d_lat && d_lon && pre, "amt_rotate_coords: NULL argument");
    GridC g;
    memset(&g, 0, sizeof g);
    int rc = fill_grid(pre, g, true);
    if (rc) return rc;
    if (n == 0 || g.prerotate == AMT_PRE_NONE) return AMT_OK;
    k_rotate_coords<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_lat, d_lon, n, g);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// maskedByPolygon (mapping.py:866-917): crossing-number test of every corner against an ordered
// polygon (utils.py:58-74; x = latitude, y = longitude as in the reference's call), then a
// centre is masked unless its four corners are all defined and inside.
constexpr int kPolyChunk = 1024;

__global__ void __launch_bounds__(256) k_polygon_corner_bits(int W, int H, const double* __restrict__ lat_k,
                                                             const double* __restrict__ lon_k,
                                                             const double* __restrict__ poly, int n_poly,
                                                             const __grid_constant__ GridC g,
                                                             unsigned* __restrict__ inside_bits, int wpr_k,
                                                             unsigned long long* __restrict__ n_inside) {
    __shared__ double s_x[kPolyChunk + 1], s_y[kPolyChunk + 1];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    double tx = qnan(), ty = qnan();
    if (x <= W) {
        const size_t i = (size_t)y * (W + 1) + x;
        tx = lat_k[i];
        ty = lon_k[i];
        if (tx == tx && g.prerotate != AMT_PRE_NONE) prerotate(g, tx, ty);
    }
    bool inside = false;
    // edge j runs from vertex j-1 (cyclically) to vertex j; chunks overlap by one vertex
    for (int base = 0; base < n_poly; base += kPolyChunk) {
        const int cnt = min(kPolyChunk, n_poly - base);
        __syncthreads();
        for (int j = threadIdx.x; j <= cnt; j += blockDim.x) {
            const int v = (base + j - 1 + n_poly) % n_poly;
            s_x[j] = poly[2 * v];
            s_y[j] = poly[2 * v + 1];
        }
        __syncthreads();
        double x0 = s_x[0], y0 = s_y[0];
        bool f0 = y0 >= ty;
        for (int j = 1; j <= cnt; ++j) {
            const double x1 = s_x[j], y1 = s_y[j];
            const bool f1 = y1 >= ty;
            if (f0 != f1) {
                const bool hit = ((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == f1;
                inside ^= hit;
            }
            x0 = x1; y0 = y1; f0 = f1;
        }
    }
    const unsigned b = __ballot_sync(0xffffffffu, inside && tx == tx && ty == ty);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < wpr_k) {
        inside_bits[(size_t)y * wpr_k + (x >> 5)] = b;
        if (n_inside && b) atomicAdd(n_inside, (unsigned long long)__popc(b));
    }
}

__global__ void __launch_bounds__(256) k_polygon_center_mask(int W, int H, const unsigned* __restrict__ inside_bits,
                                                             int wpr_k, unsigned char* __restrict__ mask) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    auto bit = [&](int yy, int xx) { return (inside_bits[(size_t)yy * wpr_k + (xx >> 5)] >> (xx & 31)) & 1u; };
    const unsigned all4 = bit(y, x) & bit(y, x + 1) & bit(y + 1, x) & bit(y + 1, x + 1);
    mask[(size_t)y * W + x] = all4 ? 0 : 1;
}

extern "C" int amt_polygon_center_mask(amt_ctx* ctx, int32_t W, int32_t H, const double* d_lat_k,
                                       const double* d_lon_k, const double* d_polygon, int32_t n_polygon,
                                       const amt_grid* pre, uint8_t* d_center_mask, uint64_t* d_n_inside,
                                       void* stream) {
    ENTER(ctx);
    CHECK_ARG(W > 0 && H > 0 && d_lat_k && d_lon_k && d_polygon && d_center_mask, "amt_polygon_center_mask: bad arguments");
    CHECK_ARG(n_polygon >= 3, "amt_polygon_center_mask: a polygon needs at least 3 points");
    GridC g;
    memset(&g, 0, sizeof g);
    if (pre) {
        int rc = fill_grid(pre, g, true);
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int wk = wpr_of(W + 1);
    void* scratch = nullptr;
    int rc = ensure_scratch(ctx, st, (size_t)wk * (H + 1) * 4, &scratch);
    if (rc) return rc;
    unsigned* bits = (unsigned*)scratch;
    if (d_n_inside) CUDA_TRY(cudaMemsetAsync(d_n_inside, 0, sizeof(uint64_t), st));
    dim3 gk((wk * 32 + 255) / 256, H + 1);
    k_polygon_corner_bits<<<gk, 256, 0, st>>>(W, H, d_lat_k, d_lon_k, d_polygon, n_polygon, g, bits, wk,
                                              (unsigned long long*)d_n_inside);
    LAUNCH_CHECK(ctx);
    dim3 gc((W + 255) / 256, H);
    k_polygon_center_mask<<<gc, 256, 0, st>>>(W, H, bits, wk, d_center_mask);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// numpy.linspace(start, stop, num)[i] = fl(fl(i*step) + start), last element == stop
__device__ __forceinline__ double linspace_at(double start, double stop, double step, int num, int i) {
    return i == num - 1 ? stop : __dadd_rn(__dmul_rn((double)i, step), start);
}

// resample.py:229-241: node coordinates -> corner (node + step/2) and centre grids.
__global__ void k_plate_carree(int nx, int ny, double lat_hi, double lat_lo, double lat_step,
                               double lon_lo, double lon_hi, double lon_step,
                               double* __restrict__ lat_k, double* __restrict__ lon_k,
                               double* __restrict__ lat_c, double* __restrict__ lon_c) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c <= nx && r <= ny) {
        const size_t i = (size_t)r * (nx + 1) + c;
        if (lat_k) lat_k[i] = linspace_at(lat_hi, lat_lo, lat_step, ny + 2, r) + lat_step / 2;
        if (lon_k) lon_k[i] = linspace_at(lon_lo, lon_hi, lon_step, nx + 2, c) + lon_step / 2;
    }
    if (c < nx && r < ny) {
        const size_t i = (size_t)r * nx + c;
        if (lat_c) lat_c[i] = linspace_at(lat_hi, lat_lo, lat_step, ny + 2, r + 1);
        if (lon_c) lon_c[i] = linspace_at(lon_lo, lon_hi, lon_step, nx + 2, c + 1);
    }
}

extern "C" int amt_plate_carree_coords(amt_ctx* ctx, int32_t nx, int32_t ny, double lat_hi, double lat_lo,
                                       double lon_lo, double lon_hi, double* d_lat_k, double* d_lon_k,
                                       double* d_lat_c, double* d_lon_c, void* stream) {
    ENTER(ctx);
    CHECK_ARG(nx > 0 && ny > 0, "amt_plate_carree_coords: empty grid");
    // np.linspace(..., retstep=True): step = (stop - start) / (num - 1)
    const double lat_step = (lat_lo - lat_hi) / (double)(ny + 1);
    const double lon_step = (lon_hi - lon_lo) / (double)(nx + 1);
    dim3 block(32, 8), grid((nx + 1 + 31) / 32, (ny + 1 + 7) / 8);
    k_plate_carree<<<grid, block, 0, (cudaStream_t)stream>>>(nx, ny, lat_hi, lat_lo, lat_step, lon_lo, lon_hi,
                                                            lon_step, d_lat_k, d_lon_k, d_lat_c, d_lon_c);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// ================================================= generic lat/lon -> MLat/MLT route
__global__ void k_latlon_to_mlatmlt(const double* __restrict__ lat, const double* __restrict__ lon, size_t n,
                                    double alt, double a, double e2, FrameC fm /* m_sm = geo->sm */,
                                    double* __restrict__ mlat, double* __restrict__ mlt) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // mapping.py:540-550: deg2rad, geodetic2Ecef(lat, lon, altitude), geoToMLatMLT
    const double la = lat[i] * kDeg2Rad, lo = lon[i] * kDeg2Rad;
    double G[3], S[3];
    geodetic2ecef(a, e2, la, lo, alt, G[0], G[1], G[2]);
    mat3(fm.m_sm, G, S);
    double ml, mt;
    sm_to_mlat_mlt(S, ml, mt);
    mlat[i] = ml;
    mlt[i] = mt;
}

extern "C" int amt_latlon_to_mlatmlt(amt_ctx* ctx, const double* d_lat, const double* d_lon, size_t n,
                                     double altitude, double wgs_a, double wgs_b, const double m_geo_sm[9],
                                     double* d_mlat, double* d_mlt, void* stream) {
    ENTER(ctx);
    CHECK_ARG(d_lat && d_lon && d_mlat && d_mlt && m_geo_sm, "amt_latlon_to_mlatmlt: NULL argument");
    if (n == 0) return AMT_OK;
    FrameC fm;
    memset(&fm, 0, sizeof fm);
    memcpy(fm.m_sm, m_geo_sm, sizeof fm.m_sm);
    volatile double aa = wgs_a * wgs_a, bb = wgs_b * wgs_b;
    volatile double num = aa - bb;
    const double e2 = num / aa;
    k_latlon_to_mlatmlt<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_lat, d_lon, n, altitude, wgs_a, e2, fm, d_mlat, d_mlt);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// smToLatLon (transform.py:461-485): spherical_to_cartesian(1, lat, lon) -> sm_to_geo (the
// transpose of mat_geo_to_sm) -> ecef2Geodetic -> degrees.  Reference order, libm (cold path).
__global__ void k_sm_to_latlon(double* __restrict__ lat, double* __restrict__ lon, size_t n, FrameC fm,
                               double a, double b, double e2a, double d) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double la = lat[i] * kDeg2Rad, lo = lon[i] * kDeg2Rad;
    double sl, cl, so, co;
    sincos(la, &sl, &cl);
    sincos(lo, &so, &co);
    const double S[3] = {(cl * 1.0) * co, (cl * 1.0) * so, sl * 1.0};
    // reverse transform: mat.T . v
    const double* M = fm.m_sm;
    const double G0 = (M[0] * S[0] + M[3] * S[1]) + M[6] * S[2];
    const double G1 = (M[1] * S[0] + M[4] * S[1]) + M[7] * S[2];
    const double G2 = (M[2] * S[0] + M[5] * S[1]) + M[8] * S[2];
    double l2, o2;
    bowring_ref(a, b, e2a, d, G0, G1, G2, l2, o2);
    lat[i] = l2 * kRad2Deg;
    lon[i] = o2 * kRad2Deg;
}

extern "C" int amt_sm_to_latlon(amt_ctx* ctx, double* d_lat, double* d_lon, size_t n, const double m_geo_sm[9],
                                double wgs_a, double wgs_b, void* stream) {
    ENTER(ctx);
    CHECK_ARG(d_lat && d_lon && m_geo_sm, "amt_sm_to_latlon: NULL argument");
    if (n == 0) return AMT_OK;
    FrameC fm;
    memset(&fm, 0, sizeof fm);
    memcpy(fm.m_sm, m_geo_sm, sizeof fm.m_sm);
    volatile double aa = wgs_a * wgs_a, bb = wgs_b * wgs_b;
    volatile double num = aa - bb;
    volatile double e2 = num / aa;
    k_sm_to_latlon<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_lat, d_lon, n, fm, wgs_a, wgs_b,
                                                                               e2 * wgs_a, num / wgs_b);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// ============================================ ground-based imagers: altitude reprojection
// themis.py:224-253 `reproject`: corner coordinates calibrated for one emission height are
// moved to another: geodetic2Ecef(lat, lon, heightRef) - station -> viewing direction ->
// intersection with the ellipsoid inflated by heightNew -> ecef2Geodetic -> degrees.
__global__ void __launch_bounds__(256) k_reproject(const double* __restrict__ lat_ref,
                                                   const double* __restrict__ lon_ref, size_t n,
                                                   const __grid_constant__ FrameC f, double e2, double h_ref,
                                                   double* __restrict__ lat_out, double* __restrict__ lon_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double la = lat_ref[i], lo = lon_ref[i];
    double ol = qnan(), oo = qnan();
    if (la == la && lo == lo) {
        double G[3], P[3];
        geodetic2ecef(f.a, e2, la * kDeg2Rad, lo * kDeg2Rad, h_ref, G[0], G[1], G[2]);
        const double dir[3] = {G[0] - f.cam[0], G[1] - f.cam[1], G[2] - f.cam[2]};
        bool graze;
        if (intersect(f, dir, P, graze)) {
            double r2;
            bowring(f.b_over_a, f.e2a, f.d, P[0], P[1], P[2], ol, oo, r2);     // degrees
        }
    }
    lat_out[i] = ol;
    lon_out[i] = oo;
}

extern "C" int amt_reproject(amt_ctx* ctx, const double* d_lat_ref, const double* d_lon_ref, size_t n,
                             const double station_ecef[3], double height_ref, double height_new, double wgs_a,
                             double wgs_b, double* d_lat_out, double* d_lon_out, void* stream) {
    ENTER(ctx);
    CHECK_ARG(d_lat_ref && d_lon_ref && station_ecef && d_lat_out && d_lon_out, "amt_reproject: NULL argument");
    CHECK_ARG(wgs_a > 0 && wgs_b > 0 && wgs_a + height_new > 0 && wgs_b + height_new > 0, "amt_reproject: bad ellipsoid");
    if (n == 0) return AMT_OK;
    amt_frame fr;
    memset(&fr, 0, sizeof fr);
    fr.width = fr.height = 1;
    const double a = wgs_a + height_new, b = wgs_b + height_new;
    memcpy(fr.cam, station_ecef, sizeof fr.cam);
    fr.inv_axes[0] = fr.inv_axes[1] = 1.0 / a;
    fr.inv_axes[2] = 1.0 / b;
    // intersection.py:239-241 _isInsideEllipsoid
    const double q0 = station_ecef[0] / a, q1 = station_ecef[1] / a, q2 = station_ecef[2] / b;
    fr.origin_inside = (q0 * q0 + q1 * q1 + q2 * q2) < 1.0;
    fr.wgs_a = wgs_a;
    fr.wgs_b = wgs_b;
    fr.model = AMT_MODEL_WCS;
    GeorefParams p;
    memset(&p, 0, sizeof p);
    int rc = fill_frame(&fr, p);
    if (rc) return rc;
    volatile double aa = wgs_a * wgs_a, bb = wgs_b * wgs_b;
    volatile double num = aa - bb;
    const double e2 = num / aa;
    k_reproject<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_lat_ref, d_lon_ref, n, p.f, e2,
                                                                              height_ref, d_lat_out, d_lon_out);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// themis.py:425-426 / astrometry.py:154-160: centre = mean of the four corners,
// ((a + b) + c + d) / 4 in the reference's summation order; NaN if any corner is.
__global__ void __launch_bounds__(256) k_corner_means(int W, int H, const double* __restrict__ lat_k,
                                                      const double* __restrict__ lon_k,
                                                      double* __restrict__ lat_c, double* __restrict__ lon_c) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t k = (size_t)y * (W + 1) + x, i = (size_t)y * W + x;
    // lats[:-1,:-1] + lats[1:,:-1] + lats[:-1,1:] + lats[1:,1:]
    lat_c[i] = (((lat_k[k] + lat_k[k + W + 1]) + lat_k[k + 1]) + lat_k[k + W + 2]) / 4.0;
    lon_c[i] = (((lon_k[k] + lon_k[k + W + 1]) + lon_k[k + 1]) + lon_k[k + W + 2]) / 4.0;
}

extern "C" int amt_corner_means(amt_ctx* ctx, int32_t W, int32_t H, const double* d_lat_k, const double* d_lon_k,
                                double* d_lat_c, double* d_lon_c, void* stream) {
    ENTER(ctx);
    CHECK_ARG(W > 0 && H > 0 && d_lat_k && d_lon_k && d_lat_c && d_lon_c, "amt_corner_means: bad arguments");
    dim3 grid((W + 255) / 256, H);
    k_corner_means<<<grid, 256, 0, (cudaStream_t)stream>>>(W, H, d_lat_k, d_lon_k, d_lat_c, d_lon_c);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// ================================================================== stage 3: binning
// Flat cell (row 0 = northernmost) of a centre coordinate, or -1.
template <bool NEAR>
__device__ __forceinline__ int cell_of(const GridC& g, double la, double lo, int& ix, int& iy, bool& near) {
    near = false;
    ix = iy = -1;
    if (!(la == la)) return -1;                       // resample.py:316: dropped iff lat is NaN
    prerotate(g, la, lo);
    bool nx_, ny_;
    ix = bin_index<NEAR>(lo, g.lo_x, g.hi_x, g.step_x, g.inv_step_x, g.nx, g.round_x, g.eps_x, nx_);
    iy = bin_index<NEAR>(la, g.lo_y, g.hi_y, g.step_y, g.inv_step_y, g.ny, g.round_y, g.eps_y, ny_);
    near = nx_ || ny_;
    if (ix < 0 || iy < 0) return -1;
    return (g.ny - 1 - iy) * g.nx + ix;               // flipud, resample.py:349
}

// Accumulate one warp's samples.  Consecutive lanes that hit the same cell form a run
// (neighbouring pixels land in the same ~100"/px cell); run sums come from one segmented warp
// scan over the channels packed into 64-bit words, and only the last lane of a run issues
// the atomics.
template <typename T, int C>
struct Packed {
    static constexpr int BITS = sizeof(T) == 1 ? 16 : 21;      // 32*255 < 2^16, 32*65535 < 2^21
    static constexpr int PER = 64 / BITS;
    static constexpr int NW = (C + PER - 1) / PER;
    static constexpr unsigned long long MASK = (1ULL << BITS) - 1;
};

// Side channel (elevation) modes: kSideNone; kSideF64 = f64 atomics (any value, NaN included;
// the sum depends on the order of the atomics in its last bits); kSideFixed = the value is scaled
// by a power of two chosen by the caller so that the grand total fits 62 bits (amt_grid.side_scale)
// and accumulated as int64: exact, order independent, run-to-run and rank-to-rank identical.
enum { kSideNone = 0, kSideF64 = 1, kSideFixed = 2 };

template <typename T, int C, int SIDE>
__device__ __forceinline__ void warp_accumulate(int cell, const unsigned (&val)[C], double side,
                                                unsigned long long sfx,
                                                unsigned long long* __restrict__ count,
                                                unsigned long long* __restrict__ sums,
                                                double* __restrict__ fsum, size_t plane) {
    using P = Packed<T, C>;
    const unsigned lane = threadIdx.x & 31;
    const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
    const bool head = lane == 0 || prev != cell;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned le = 0xffffffffu >> (31 - lane);
    const int start = 31 - __clz(heads & le);         // first lane of my run
    const int next = __shfl_down_sync(0xffffffffu, cell, 1);
    const bool tail = lane == 31 || next != cell;
    unsigned long long w[P::NW];
#pragma unroll
    for (int k = 0; k < P::NW; ++k) w[k] = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) w[c / P::PER] |= (unsigned long long)val[c] << ((c % P::PER) * P::BITS);
    // a NaN side value (possible only for hand-made mappings) must stay inside its own cell:
    // then the side channel of this warp falls back to one atomic per sample
    if (SIDE == kSideF64 && __ballot_sync(0xffffffffu, cell >= 0 && side != side) != 0) {
        if (cell >= 0) atomicAdd(&fsum[cell], side);
        side = 0.0;                                   // the scan below then carries zeros for this warp
    }
    // Segmented inclusive scan (Kogge-Stone on the position inside the run): after the step of
    // distance d a lane holds the sum of the last min(2d, off+1) samples of its run.  Runs are
    // short (a 100"/px cell holds ~3 consecutive 34" pixels), so the loop stops as soon as d
    // exceeds the longest run of the warp -- typically after 2-3 of the 5 steps.
    const unsigned off = lane - (unsigned)start;
    const unsigned maxoff = __reduce_max_sync(0xffffffffu, off);
    double s = side;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        if ((unsigned)d > maxoff) break;              // warp-uniform
#pragma unroll
        for (int k = 0; k < P::NW; ++k) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, w[k], d);
            if (off >= (unsigned)d) w[k] += t;
        }
        if (SIDE == kSideF64) {
            const double t = __shfl_up_sync(0xffffffffu, s, d);
            if (off >= (unsigned)d) s += t;
        }
        if (SIDE == kSideFixed) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, sfx, d);
            if (off >= (unsigned)d) sfx += t;
        }
    }
    if (tail && cell >= 0) {
        atomicAdd(&count[cell], (unsigned long long)(off + 1));
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const unsigned long long run = (w[c / P::PER] >> ((c % P::PER) * P::BITS)) & P::MASK;
            atomicAdd(&sums[(size_t)c * plane + cell], run);
        }
        if (SIDE == kSideF64) atomicAdd(&fsum[cell], s);
        if (SIDE == kSideFixed) atomicAdd((unsigned long long*)fsum + cell, sfx);
    }
}

// The same accumulation with hardware warp reductions instead of the scan (integer channels, side channel
// absent or fixed point with side_scale <= 2^32, the library's own range): `match.any` hands every lane the mask
// of the lanes that share its cell, and `redux.sync` adds a 32-bit word over exactly those lanes -- every group
// of the warp in the same instruction.  The samples are packed so that no field overflows into its neighbour
// over 32 lanes: u8 channels two per word (13-bit sums at a 16-bit pitch), the fixed-point side value split
// into its low 14 bits (19-bit sum, sharing a word with an odd last u8 channel at bit 13) and the signed
// rest (|value| < 2^39 -> |rest| < 2^25 -> |sum| < 2^30).  The lowest lane of a group issues its atomics.
// ~45 instructions per warp instead of ~170, and non-adjacent samples of the same cell aggregate too.
template <typename T, int C>
struct ReduxPack {
    static constexpr bool U8 = sizeof(T) == 1;
    static constexpr int CH_PER = U8 ? 2 : 1;
    static constexpr int NCW = (C + CH_PER - 1) / CH_PER;           // channel words
    static constexpr bool SHARE = U8 && (C & 1);                     // side-lo in the last channel word
};
constexpr double kReduxMaxScale = 4294967296.0;                      // 2^32
// Measured (BASELINE configs[1], 100"/px: 3.8 groups per warp): fused kernel 247.3 us against 237.8 us with the
// scan; at 10"/px (every lane its own cell) 366 against 317 us -- MATCH.ANY and REDUX with per-group masks are
// processed group by group.  Off by default; kept for the record and for coarser grids.
#ifndef AMT_ACC_REDUX
#define AMT_ACC_REDUX 0
#endif

template <typename T, int C, int SIDE>
__device__ __forceinline__ void warp_accumulate_redux(int cell, const unsigned (&val)[C], unsigned long long sfx,
                                                      unsigned long long* __restrict__ count,
                                                      unsigned long long* __restrict__ sums,
                                                      double* __restrict__ fsum, size_t plane) {
    using P = ReduxPack<T, C>;
    static_assert(SIDE == kSideNone || SIDE == kSideFixed, "integer sums only");
    const unsigned lane = threadIdx.x & 31;
    const unsigned grp = __match_any_sync(0xffffffffu, cell);
    unsigned w[P::NCW];
#pragma unroll
    for (int k = 0; k < P::NCW; ++k) w[k] = 0u;
#pragma unroll
    for (int c = 0; c < C; ++c) w[c / P::CH_PER] |= val[c] << ((c % P::CH_PER) * 16);
    const unsigned lo = (unsigned)sfx & 0x3fffu;
    const int hi = (int)((long long)sfx >> 14);
    if (SIDE == kSideFixed && P::SHARE) w[P::NCW - 1] |= lo << 13;
#pragma unroll
    for (int k = 0; k < P::NCW; ++k) w[k] = __reduce_add_sync(grp, w[k]);
    unsigned slo = 0u;
    int shi = 0;
    if (SIDE == kSideFixed) {
        if (!P::SHARE) slo = __reduce_add_sync(grp, lo);
        shi = __reduce_add_sync(grp, hi);
    }
    if (cell >= 0 && lane == (unsigned)(__ffs(grp) - 1)) {
        atomicAdd(&count[cell], (unsigned long long)__popc(grp));
#pragma unroll
        for (int c = 0; c < C; ++c) {
            unsigned v = w[c / P::CH_PER] >> ((c % P::CH_PER) * 16);
            if (P::U8) v &= (P::SHARE && c == C - 1) ? 0x1fffu : 0xffffu;
            atomicAdd(&sums[(size_t)c * plane + cell], (unsigned long long)v);
        }
        if (SIDE == kSideFixed) {
            if (P::SHARE) slo = w[P::NCW - 1] >> 13;
            const long long tot = (long long)shi * 16384LL + (long long)slo;
            atomicAdd((unsigned long long*)fsum + cell, (unsigned long long)tot);
        }
    }
}

// value -> fixed point; a NaN (hand-made mappings only) cannot be represented and poisons nothing:
// the caller falls back to kSideF64 for such mappings (amt_grid.side_scale == 0)
__device__ __forceinline__ unsigned long long to_fixed(double v, double scale) {
    return (unsigned long long)__double2ll_rn(v * scale);
}

// Each warp bins kBinSeg consecutive segments of 32 pixels; all global loads of the segments
// are issued before any dependent work (memory-level parallelism: the kernel is bound by load
// latency / HBM, not by arithmetic).
#ifndef AMT_BIN_SEG
#define AMT_BIN_SEG 2
#endif
constexpr int kBinSeg = AMT_BIN_SEG;

template <typename T, int C, bool NEAR, int SIDE>
__global__ void __launch_bounds__(256) k_bin(const double* __restrict__ lat, const double* __restrict__ lon,
                                             const double* __restrict__ side, const T* __restrict__ img, size_t n,
                                             const __grid_constant__ GridC g, unsigned long long* __restrict__ count,
                                             unsigned long long* __restrict__ sums, double* __restrict__ fsum,
                                             unsigned long long* __restrict__ near_counter) {
    const unsigned lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t base = warp * (32 * kBinSeg) + lane;
    double la[kBinSeg], lo[kBinSeg], sd[kBinSeg];
    unsigned val[kBinSeg][C];
#pragma unroll
    for (int k = 0; k < kBinSeg; ++k) {
        const size_t i = base + 32 * k;
        const bool in = i < n;
        // read once: streaming loads (evict-first) leave the L2 to the grid accumulators
        la[k] = in ? __ldcs(&lat[i]) : qnan();
        lo[k] = in ? __ldcs(&lon[i]) : 0.0;
        sd[k] = (SIDE != kSideNone && in) ? __ldcs(&side[i]) : 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) val[k][c] = in ? (unsigned)__ldcs(&img[i * C + c]) : 0u;
    }
    bool near_any = false;
#pragma unroll
    for (int k = 0; k < kBinSeg; ++k) {
        int cell = -1, ix, iy;
        bool near = false;
        if (la[k] == la[k]) cell = cell_of<NEAR>(g, la[k], lo[k], ix, iy, near);
        near_any |= near;
        if (NEAR) {
            const unsigned m = __ballot_sync(0xffffffffu, near);
            if (m && lane == 0) atomicAdd(near_counter, (unsigned long long)__popc(m));
        }
        if (__ballot_sync(0xffffffffu, cell >= 0) == 0) continue;   // nothing of this segment lands in the grid
        if (cell < 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) val[k][c] = 0;
            sd[k] = 0.0;
        }
        if (SIDE != kSideF64 && AMT_ACC_REDUX && (SIDE == kSideNone || g.side_scale <= kReduxMaxScale))
            warp_accumulate_redux<T, C, SIDE == kSideF64 ? kSideNone : SIDE>(
                cell, val[k], SIDE == kSideFixed ? to_fixed(sd[k], g.side_scale) : 0ULL, count, sums, fsum,
                (size_t)g.nx * g.ny);
        else
            warp_accumulate<T, C, SIDE>(cell, val[k], sd[k], SIDE == kSideFixed ? to_fixed(sd[k], g.side_scale) : 0ULL,
                                        count, sums, fsum, (size_t)g.nx * g.ny);
    }
    (void)near_any;
}

template <typename T, bool NEAR, int SIDE>
static void launch_bin_c(int channels, unsigned blocks, cudaStream_t st, const double* lat, const double* lon,
                         const double* side, const void* img, size_t n, const GridC& g, unsigned long long* count,
                         unsigned long long* sums, double* fsum, unsigned long long* near) {
    const T* im = (const T*)img;
    switch (channels) {
        case 1: k_bin<T, 1, NEAR, SIDE><<<blocks, 256, 0, st>>>(lat, lon, side, im, n, g, count, sums, fsum, near); break;
        case 2: k_bin<T, 2, NEAR, SIDE><<<blocks, 256, 0, st>>>(lat, lon, side, im, n, g, count, sums, fsum, near); break;
        case 3: k_bin<T, 3, NEAR, SIDE><<<blocks, 256, 0, st>>>(lat, lon, side, im, n, g, count, sums, fsum, near); break;
        case 4: k_bin<T, 4, NEAR, SIDE><<<blocks, 256, 0, st>>>(lat, lon, side, im, n, g, count, sums, fsum, near); break;
    }
}

template <typename T, bool NEAR>
static void launch_bin(int channels, unsigned blocks, cudaStream_t st, const double* lat, const double* lon,
                       const double* side, const void* img, size_t n, const GridC& g, unsigned long long* count,
                       unsigned long long* sums, double* fsum, unsigned long long* near) {
    if (side == nullptr) launch_bin_c<T, NEAR, kSideNone>(channels, blocks, st, lat, lon, side, img, n, g, count, sums, fsum, near);
    else if (g.side_scale > 0.0) launch_bin_c<T, NEAR, kSideFixed>(channels, blocks, st, lat, lon, side, img, n, g, count, sums, fsum, near);
    else launch_bin_c<T, NEAR, kSideF64>(channels, blocks, st, lat, lon, side, img, n, g, count, sums, fsum, near);
}

extern "C" int amt_bin_accumulate(amt_ctx* ctx, const double* d_lat_c, const double* d_lon_c,
                                  const double* d_side, const void* d_img, int32_t dtype, int32_t channels,
                                  size_t n_pixels, const amt_grid* grid, uint64_t* d_count, uint64_t* d_sums,
                                  double* d_fsum, uint64_t* d_near_edge, void* stream) {
    ENTER(ctx);
    CHECK_ARG(d_lat_c && d_lon_c && d_img && grid && d_count && d_sums, "amt_bin_accumulate: NULL argument");
    CHECK_ARG(channels >= 1 && channels <= 4, "amt_bin_accumulate: channels must be 1..4");
    CHECK_ARG((d_side == nullptr) == (d_fsum == nullptr), "amt_bin_accumulate: d_side and d_fsum go together");
    if (dtype != AMT_U8 && dtype != AMT_U16)
        return set_err(AMT_ERR_UNSUPPORTED, "amt_bin_accumulate: image dtype must be uint8 or uint16 (mapping.py:1005)");
    GridC g;
    int rc = fill_grid(grid, g);
    if (rc) return rc;
    if (n_pixels == 0) return AMT_OK;
    const size_t per_block = 256 * (size_t)kBinSeg;
    const unsigned blocks = (unsigned)((n_pixels + per_block - 1) / per_block);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* cnt = (unsigned long long*)d_count;
    unsigned long long* sm = (unsigned long long*)d_sums;
    unsigned long long* ne = (unsigned long long*)d_near_edge;
    if (dtype == AMT_U8) {
        if (ne) launch_bin<unsigned char, true>(channels, blocks, st, d_lat_c, d_lon_c, d_side, d_img, n_pixels, g, cnt, sm, d_fsum, ne);
        else launch_bin<unsigned char, false>(channels, blocks, st, d_lat_c, d_lon_c, d_side, d_img, n_pixels, g, cnt, sm, d_fsum, ne);
    } else {
        if (ne) launch_bin<unsigned short, true>(channels, blocks, st, d_lat_c, d_lon_c, d_side, d_img, n_pixels, g, cnt, sm, d_fsum, ne);
        else launch_bin<unsigned short, false>(channels, blocks, st, d_lat_c, d_lon_c, d_side, d_img, n_pixels, g, cnt, sm, d_fsum, ne);
    }
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

__global__ void k_cell_indices(const double* __restrict__ lat, const double* __restrict__ lon, size_t n,
                               const __grid_constant__ GridC g, int* __restrict__ oix, int* __restrict__ oiy) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ix, iy;
    bool near;
    cell_of<false>(g, lat[i], lon[i], ix, iy, near);
    oix[i] = ix;
    oiy[i] = iy;
}

extern "C" int amt_cell_indices(amt_ctx* ctx, const double* d_lat_c, const double* d_lon_c, size_t n_pixels,
                                const amt_grid* grid, int32_t* d_ix, int32_t* d_iy, void* stream) {
    ENTER(ctx);
    CHECK_ARG(d_lat_c && d_lon_c && grid && d_ix && d_iy, "amt_cell_indices: NULL argument");
    GridC g;
    int rc = fill_grid(grid, g);
    if (rc) return rc;
    if (n_pixels == 0) return AMT_OK;
    k_cell_indices<<<(unsigned)((n_pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_lat_c, d_lon_c, n_pixels, g, d_ix, d_iy);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// resample.py:339-351 + :128-136: mean = sum/count in float64, NaN where count == 0,
// np.round (half-even) + cast for integer images.
// rint(sum / n) of the reference (float64 division, round half to even) for integer sums: when sum and n fit
// 32 bits the quotient is formed in integers -- exact, and equal to the float64 result (sum/n = q + r/n with
// |r/n - 1/2| >= 1/(2n) >= 2^-33, far more than an ulp of a quotient < 2^32, so the rounded division cannot
// create or remove a tie); an IEEE double division costs ~40 instructions, and at 10"/px a frame has 20 M cells.
__device__ __forceinline__ unsigned long long mean_rint(unsigned long long sum, unsigned long long n) {
    if ((sum | n) >> 32) return (unsigned long long)rint((double)sum / (double)n);
    const unsigned s32 = (unsigned)sum, n32 = (unsigned)n;
    const unsigned q = s32 / n32, r = s32 - q * n32;
    const unsigned twice = 2u * r;                      // r < n < 2^32: the carry is the comparison
    const bool above = twice < r || twice > n32;        // 2r > n
    return q + ((above || (twice == n32 && (q & 1u))) ? 1u : 0u);
}

template <typename T>
__global__ void __launch_bounds__(256) k_normalise(size_t cells, int C, const unsigned long long* __restrict__ count,
                            const unsigned long long* __restrict__ sums, const double* __restrict__ fsum,
                            T* __restrict__ out_img, unsigned char* __restrict__ out_mask, double* __restrict__ out_side,
                            double side_scale, double inv_scale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const unsigned long long n = count[i];
    // every load of the cell is issued before the first use (the sums used to wait for the count)
    unsigned long long sm[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < C) sm[c] = sums[(size_t)c * cells + i];
    double tot = 0.0;
    if (out_side) {
        tot = fsum[i];
        if (side_scale > 0.0) tot = (double)((const long long*)fsum)[i] * inv_scale;      // exact: power of two
    }
    if (out_mask) out_mask[i] = n == 0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < C) out_img[i * C + c] = n ? (T)mean_rint(sm[c], n) : (T)0;
    if (out_side) out_side[i] = n ? div_fast(tot, (double)n) : qnan();       // <= 1 ulp (amt_fastmath.cuh)
}

extern "C" int amt_normalise(amt_ctx* ctx, const amt_grid* grid, int32_t dtype, int32_t channels,
                             const uint64_t* d_count, const uint64_t* d_sums, const double* d_fsum,
                             void* d_out_img, uint8_t* d_out_mask, double* d_out_side, void* stream) {
    ENTER(ctx);
    CHECK_ARG(grid && d_count && d_sums && d_out_img, "amt_normalise: NULL argument");
    CHECK_ARG(grid->nx > 0 && grid->ny > 0, "amt_normalise: empty grid");
    CHECK_ARG(channels >= 1 && channels <= 4, "amt_normalise: channels must be 1..4");
    CHECK_ARG(!d_out_side || d_fsum, "amt_normalise: d_out_side needs d_fsum");
    const size_t cells = (size_t)grid->nx * grid->ny;
    const unsigned blocks = (unsigned)((cells + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AMT_U8)
        k_normalise<unsigned char><<<blocks, 256, 0, st>>>(cells, channels, (const unsigned long long*)d_count,
                                                          (const unsigned long long*)d_sums, d_fsum,
                                                          (unsigned char*)d_out_img, d_out_mask, d_out_side,
                                                          grid->side_scale > 0 ? grid->side_scale : 0.0,
                                                          grid->side_scale > 0 ? 1.0 / grid->side_scale : 0.0);
    else if (dtype == AMT_U16)
        k_normalise<unsigned short><<<blocks, 256, 0, st>>>(cells, channels, (const unsigned long long*)d_count,
                                                           (const unsigned long long*)d_sums, d_fsum,
                                                           (unsigned short*)d_out_img, d_out_mask, d_out_side,
                                                           grid->side_scale > 0 ? grid->side_scale : 0.0,
                                                           grid->side_scale > 0 ? 1.0 / grid->side_scale : 0.0);
    else
        return set_err(AMT_ERR_UNSUPPORTED, "amt_normalise: image dtype must be uint8 or uint16");
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// ================================================= fused georeference (+ binning) kernel
// The frame's FINAL validity bitmaps are known before this kernel runs (hit test by the limb
// solver / k_hit_bits, then the sanitisation stencils on the bitmaps), and with them the outline
// statistics and therefore the target grid.  One pass then does everything per pixel: corner ray
// and centre ray -> intersection -> lat/lon, MLat/MLT, elevation -> (PLANES) the nine coordinate
// planes, NaN where the bitmaps say so -- no later NaN patching -- and (BIN) the cell of the centre
// on the target grid and the accumulation of its image sample, with the coordinates still in
// registers: the binning pass never re-reads 27 B/pixel of planes.
//   PLANES = false, BIN = true is the plane-free resampling (3 B/pixel in, the grids out).
//
// Launch shape: a CTA is a TILE of kFusedCols x kFusedRows pixels (32 x 4: each warp one 32-pixel
// row segment, i.e. one word of each bitmap and 256-byte coalesced plane stores).  The scatter is
// PRIVATISED IN SHARED MEMORY per tile (resample.py:330-338 / histogram.py:244-250 `bincount`): the
// tile's samples land in a small window of the target grid (~12 x 7 cells at 100"/px), whose
// bounds come from one warp min/max reduction per row segment; samples are added with 32-bit
// shared-memory atomics (count and u8 channel sums packed two per word, the fixed-point elevation
// split 24 | 40 bits), and each touched cell of the window is flushed ONCE with global 64-bit
// reductions.  Per pixel that is 4 ATOMS instead of the ~100-instruction segmented warp scan of
// warp_accumulate, and ~2x fewer global atomics than the run-aggregated strips.  A tile whose
// window exceeds the shared-memory capacity (grid much finer than the pixels: every sample its own
// cell, nothing to privatise) or that needs f64 side sums uses warp_accumulate directly.
#ifndef AMT_FUSED_ROWS
#define AMT_FUSED_ROWS 4
#endif
#ifndef AMT_FUSED_PRIV
#define AMT_FUSED_PRIV 0
#endif
#ifndef AMT_FUSED_ITER
#define AMT_FUSED_ITER 4
#endif
constexpr int kFusedRows = AMT_FUSED_ROWS;       // rows of a tile = warps of the CTA
constexpr int kFusedCols = 32;
constexpr int kFusedThreads = kFusedCols * kFusedRows;
// Resident CTAs per SM the kernel is compiled for.  128-thread CTAs at 8 per SM are the 64-register budget of
// the point kernels; the pure TAN instantiation fits 56 registers without spilling (9 CTAs' worth of
// registers), which leaves room for the CTAs of the small stage-A kernels of other frames to run BESIDE
// the fused kernel's instead of displacing them (in the engine: 0.276 -> 0.262 ms per frame together with
// the smaller CTAs; the SIP instantiation needs its 64 registers).
#ifndef AMT_FUSED_MINBLOCKS
#define AMT_FUSED_MINBLOCKS(SIP) ((AMT_GEOREF_MINBLOCKS * 256 / (32 * AMT_FUSED_ROWS)) + ((SIP) ? 0 : 1))
#endif
constexpr int kFusedIter = AMT_FUSED_ITER;       // 32 x 8 tiles per CTA (<= 32: one lane per iteration holds its bitmap words)
static_assert(kFusedIter >= 1 && kFusedIter <= 32, "one lane per iteration");
constexpr int kTileWords = 3072;                 // 12 KB of shared memory per CTA
static_assert(AMT_FUSED_PRIV == 0 || kFusedThreads == 256, "the privatised scatter is written for 256-thread tiles");

template <typename T, int C>
struct TileAcc {
    static constexpr int NF = 1 + C;                       // count + channels
    static constexpr int PER = sizeof(T) == 1 ? 2 : 1;     // u8: 16-bit fields (256 * 255 < 2^16, count <= 256)
    static constexpr int SHIFT = 32 / PER;
    static constexpr int NW = (NF + PER - 1) / PER;
    static constexpr int WORDS = NW + 2;                   // + fixed-point side channel: low 24 bits | rest
    static constexpr int CAP = kTileWords / WORDS;         // cells of the window
    static constexpr unsigned FMASK = PER == 2 ? 0xffffu : 0xffffffffu;
};

// Tile-privatised accumulation (all 256 threads of the CTA call it; s_acc zeroed, s_win =
// {INT_MAX, INT_MAX, -1, -1} and a barrier passed since).  (ix, fy) = column and output row of the
// sample's cell, cell < 0 = no sample.  Returns false if the tile must take the global path.
template <typename T, int C>
__device__ __forceinline__ bool tile_accumulate(int cell, int ix, int fy, const unsigned (&val)[C],
                                                unsigned long long sfx, bool has_side, unsigned* s_acc, int* s_win,
                                                int nx, size_t plane, unsigned long long* __restrict__ count,
                                                unsigned long long* __restrict__ sums,
                                                unsigned long long* __restrict__ fsum) {
    using A = TileAcc<T, C>;
    const unsigned lane = threadIdx.x & 31;
    const bool ok = cell >= 0;
    const int mnx = __reduce_min_sync(0xffffffffu, ok ? ix : 0x7fffffff);
    const int mny = __reduce_min_sync(0xffffffffu, ok ? fy : 0x7fffffff);
    const int mxx = __reduce_max_sync(0xffffffffu, ok ? ix : -1);
    const int mxy = __reduce_max_sync(0xffffffffu, ok ? fy : -1);
    if (lane == 0 && mxx >= 0) {
        atomicMin(&s_win[0], mnx); atomicMin(&s_win[1], mny);
        atomicMax(&s_win[2], mxx); atomicMax(&s_win[3], mxy);
    }
    __syncthreads();
    const int x0 = s_win[0], y0 = s_win[1], x1 = s_win[2], y1 = s_win[3];
    if (x1 < 0) return true;                              // no sample of this tile lands in the grid
    const int wx = x1 - x0 + 1, wy = y1 - y0 + 1;
    if (wy > A::CAP || wx > A::CAP || wx * wy > A::CAP) return false;      // CTA-uniform
    if (ok) {
        const int loc = (fy - y0) * wx + (ix - x0);
        unsigned w[A::NW];
#pragma unroll
        for (int k = 0; k < A::NW; ++k) w[k] = 0u;
        w[0] = 1u;
#pragma unroll
        for (int c = 0; c < C; ++c) w[(c + 1) / A::PER] |= val[c] << (((c + 1) % A::PER) * A::SHIFT);
#pragma unroll
        for (int k = 0; k < A::NW; ++k) atomicAdd(&s_acc[k * A::CAP + loc], w[k]);
        if (has_side) {
            atomicAdd(&s_acc[A::NW * A::CAP + loc], (unsigned)(sfx & 0xffffffu));
            atomicAdd(&s_acc[(A::NW + 1) * A::CAP + loc], (unsigned)(int)((long long)sfx >> 24));
        }
    }
    __syncthreads();
    const int wcells = wx * wy;
    for (int i = threadIdx.x; i < wcells; i += 256) {
        unsigned w[A::NW];
#pragma unroll
        for (int k = 0; k < A::NW; ++k) w[k] = s_acc[k * A::CAP + i];
        const unsigned n = w[0] & A::FMASK;
        if (n == 0) continue;
        const int ly = i / wx, lx = i - ly * wx;
        const size_t gc = (size_t)(y0 + ly) * nx + (x0 + lx);
        atomicAdd(&count[gc], (unsigned long long)n);
#pragma unroll
        for (int c = 0; c < C; ++c)
            atomicAdd(&sums[(size_t)c * plane + gc],
                      (unsigned long long)((w[(c + 1) / A::PER] >> (((c + 1) % A::PER) * A::SHIFT)) & A::FMASK));
        if (has_side) {
            const long long hi = (long long)(int)s_acc[(A::NW + 1) * A::CAP + i];
            const long long tot = hi * 16777216LL + (long long)s_acc[A::NW * A::CAP + i];
            atomicAdd(&fsum[gc], (unsigned long long)tot);
        }
    }
    return true;
}

// One 32-pixel row segment of the tile: everything the fused kernel does for pixel (x, y), given the
// words mk / mc of the frame's final bitmaps that cover the warp's 32 corners / centres.
template <typename T, int C, bool PLANES, bool MAG, bool BIN, bool SIP, bool PRIV>
__device__ __forceinline__ void fused_row(const GeorefParams& p, const double* s_sip, unsigned* s_acc, int* s_win,
                                          const int x, const int y, const unsigned mk, const unsigned mc,
                                          const T* __restrict__ img, const GridC& g,
                                          unsigned long long* __restrict__ count,
                                          unsigned long long* __restrict__ sums, double* __restrict__ fsum) {
    const int W = p.f.W, H = p.f.H;
    const unsigned lane = threadIdx.x & 31;
    const bool in_k = PLANES && x <= W && y <= H;
    const bool in_c = x < W && y < H;
    // fill_frame guarantees (W+1)*(H+1) < 2^31: 32-bit flat indices
    const unsigned ik = (unsigned)y * (unsigned)(W + 1) + (unsigned)x, ic = (unsigned)y * (unsigned)W + (unsigned)x;
    const double nan = qnan();
    int cell = -1, cx = -1, cy = -1;
    unsigned val[C];
#pragma unroll
    for (int c = 0; c < C; ++c) val[c] = 0u;
    double elev = 0.0;
    if ((mk | mc) == 0) {                    // nothing defined in this warp's 32 pixels
        if (PLANES) {
            if (in_k) {
                PLANE_ST(&p.o.d_lat_k[ik], nan); PLANE_ST(&p.o.d_lon_k[ik], nan);
                if (MAG) { PLANE_ST(&p.o.d_mlat_k[ik], nan); PLANE_ST(&p.o.d_mlt_k[ik], nan); }
            }
            if (in_c) {
                PLANE_ST(&p.o.d_lat_c[ic], nan); PLANE_ST(&p.o.d_lon_c[ic], nan); PLANE_ST(&p.o.d_elev_c[ic], nan);
                if (MAG) { PLANE_ST(&p.o.d_mlat_c[ic], nan); PLANE_ST(&p.o.d_mlt_c[ic], nan); }
            }
        }
        if (!PRIV) return;
    } else {
        const bool vk = (mk >> lane) & 1u, vc = (mc >> lane) & 1u;
        // the image sample travels with the thread from the start: its latency hides under the FP64 chain
        if (BIN && vc) {
#pragma unroll
            for (int c = 0; c < C; ++c) val[c] = (unsigned)img[(size_t)ic * C + c];
        }
        double dk[3], dc[3], Pk[3], Pc[3];
        dirs_kc<SIP>(p.f, s_sip, s_sip + (SIP ? AMT_SIP_MAX_COEF : 0), x, y, dk, dc);
        // valid elements hit by construction (same arithmetic as the hit test).  An undefined element gets
        // a NaN into the first coordinate of its point: every dot product downstream then is NaN, i.e. all
        // nine outputs come out NaN without a select per plane (the reference's NaN rows, for free).
        if (PLANES) intersect_valid(p.f, dk, Pk);
        intersect_valid(p.f, dc, Pc);
        if (PLANES && !vk) Pk[0] = nan;
        if (!vc) Pc[0] = nan;
        double la_c, lo_c, r2_c;
        {
            double la_k, lo_k;
            if (PLANES) point_to_geo(p.f, Pk, la_k, lo_k);
            point_to_geo(p.f, Pc, la_c, lo_c, r2_c);
            if (PLANES) {
                if (in_k) { PLANE_ST(&p.o.d_lat_k[ik], la_k); PLANE_ST(&p.o.d_lon_k[ik], lo_k); }
                if (in_c) { PLANE_ST(&p.o.d_lat_c[ic], la_c); PLANE_ST(&p.o.d_lon_c[ic], lo_c); }
            }
        }
        if (PLANES && MAG) {
            double ml_k, mt_k, ml_c, mt_c;
            point_to_mag(p.f, Pk, ml_k, mt_k);
            point_to_mag(p.f, Pc, ml_c, mt_c);
            if (in_k) { PLANE_ST(&p.o.d_mlat_k[ik], ml_k); PLANE_ST(&p.o.d_mlt_k[ik], mt_k); }
            if (in_c) { PLANE_ST(&p.o.d_mlat_c[ic], ml_c); PLANE_ST(&p.o.d_mlt_c[ic], mt_c); }
        }
        if (PLANES || (BIN && fsum != nullptr)) {
            elev = elevation_deg<false>(dc, Pc, r2_c);
            if (PLANES && in_c) PLANE_ST(&p.o.d_elev_c[ic], elev);
        }
        if (BIN && vc) {
            bool near;
            cell = cell_of<false>(g, la_c, lo_c, cx, cy, near);
        }
        if (BIN && !PRIV) {
            if (__ballot_sync(0xffffffffu, cell >= 0) == 0) return;       // warp-uniform
        }
    }
    if (BIN) {
        if (cell < 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) val[c] = 0u;
        }
        const size_t plane = (size_t)g.nx * g.ny;
        const bool fixed = g.side_scale > 0.0;
        if (PRIV && (fsum == nullptr || fixed)) {
            const unsigned long long sfx = (fsum != nullptr && cell >= 0) ? to_fixed(elev, g.side_scale) : 0ULL;
            if (tile_accumulate<T, C>(cell, cx, g.ny - 1 - cy, val, sfx, fsum != nullptr, s_acc, s_win, g.nx, plane,
                                      count, sums, (unsigned long long*)fsum))
                return;
        }
        if (__ballot_sync(0xffffffffu, cell >= 0) == 0) return;
        if (AMT_ACC_REDUX && fsum == nullptr) warp_accumulate_redux<T, C, kSideNone>(cell, val, 0ULL, count, sums, fsum, plane);
        else if (AMT_ACC_REDUX && fixed && g.side_scale <= kReduxMaxScale)
            warp_accumulate_redux<T, C, kSideFixed>(cell, val, cell >= 0 ? to_fixed(elev, g.side_scale) : 0ULL, count,
                                                    sums, fsum, plane);
        else if (fsum == nullptr) warp_accumulate<T, C, kSideNone>(cell, val, 0.0, 0ULL, count, sums, fsum, plane);
        else if (fixed)
            warp_accumulate<T, C, kSideFixed>(cell, val, 0.0, cell >= 0 ? to_fixed(elev, g.side_scale) : 0ULL, count,
                                              sums, fsum, plane);
        else warp_accumulate<T, C, kSideF64>(cell, val, cell >= 0 ? elev : 0.0, 0ULL, count, sums, fsum, plane);
    }
}

// A CTA works on kFusedIter vertically adjacent 32 x 8 tiles, one after the other (warp w: row w of each).
// The bitmap words of ALL its row segments are fetched by one load per bitmap at the start (lane r holds the
// words of iteration r): the ~1 us of launch + L2 latency that precedes the first FP64 instruction of a warp
// -- 16 % of all warp residency in the one-row-per-warp version (profiles/r02_fused_stalls.txt) -- is paid
// once per kFusedIter rows.
template <typename T, int C, bool PLANES, bool MAG, bool BIN, bool SIP>
__global__ void __launch_bounds__(kFusedThreads, AMT_FUSED_MINBLOCKS(SIP))
k_georef_fused(const __grid_constant__ GeorefParams p, const uint32_t* __restrict__ valid_k,
               const uint32_t* __restrict__ valid_c, const T* __restrict__ img, const __grid_constant__ GridC g,
               unsigned long long* __restrict__ count, unsigned long long* __restrict__ sums,
               double* __restrict__ fsum) {
    constexpr bool PRIV = BIN && kFusedRows > 1 && AMT_FUSED_PRIV != 0;
    const int W = p.f.W, H = p.f.H;
    const int ty = (int)((blockIdx.y * p.row_stride) % gridDim.y);
    const int y0 = ty * (kFusedRows * kFusedIter) + (int)(threadIdx.x / kFusedCols);
    const int x = blockIdx.x * kFusedCols + (int)(threadIdx.x % kFusedCols);
    __shared__ double s_sip[SIP ? 2 * AMT_SIP_MAX_COEF : 1];
    __shared__ unsigned s_acc[PRIV ? kTileWords : 1];
    __shared__ int s_win[4];
    if (SIP) {
        for (int i = threadIdx.x; i < 2 * AMT_SIP_MAX_COEF; i += kFusedThreads)
            s_sip[i] = i < AMT_SIP_MAX_COEF ? p.sip_a[i] : p.sip_b[i - AMT_SIP_MAX_COEF];
        __syncthreads();
    }
    const unsigned lane = threadIdx.x & 31, xw = (unsigned)x >> 5;
    const int wk = (W + 1 + 31) >> 5, wc = (W + 31) >> 5;
    const int yl = y0 + (int)lane * kFusedRows;               // lane r < kFusedIter: the row of iteration r
    const bool mine = lane < (unsigned)kFusedIter;
    const unsigned mkv = (PLANES && mine && yl <= H && xw < (unsigned)wk) ? valid_k[(unsigned)yl * wk + xw] : 0u;
    const unsigned mcv = (mine && yl < H && xw < (unsigned)wc) ? valid_c[(unsigned)yl * wc + xw] : 0u;
#pragma unroll 1
    for (int r = 0; r < kFusedIter; ++r) {
        const int y = y0 + r * kFusedRows;
        if (!PRIV && y > H) break;                            // warp-uniform
        const unsigned mk = __shfl_sync(0xffffffffu, mkv, r), mc = __shfl_sync(0xffffffffu, mcv, r);
        if (PRIV) {
            if (r) __syncthreads();
#pragma unroll
            for (int i = 0; i < kTileWords / 256; ++i) s_acc[i * 256 + threadIdx.x] = 0u;
            if (threadIdx.x < 4) s_win[threadIdx.x] = threadIdx.x < 2 ? 0x7fffffff : -1;
            __syncthreads();
        }
        fused_row<T, C, PLANES, MAG, BIN, SIP, PRIV>(p, s_sip, s_acc, s_win, x, y, mk, mc, img, g, count, sums, fsum);
    }
}

template <typename T, int C, bool PLANES, bool MAG, bool BIN>
static void launch_fused_sip(bool sip, dim3 grid, cudaStream_t st, const GeorefParams& p, const uint32_t* vk,
                             const uint32_t* vc, const T* img, const GridC& g, unsigned long long* count,
                             unsigned long long* sums, double* fsum) {
    if (sip) k_georef_fused<T, C, PLANES, MAG, BIN, true><<<grid, kFusedThreads, 0, st>>>(p, vk, vc, img, g, count, sums, fsum);
    else k_georef_fused<T, C, PLANES, MAG, BIN, false><<<grid, kFusedThreads, 0, st>>>(p, vk, vc, img, g, count, sums, fsum);
}

template <typename T, int C>
static void launch_fused_c(bool planes, bool mag, bool bin, bool sip, dim3 grid, cudaStream_t st,
                           const GeorefParams& p, const uint32_t* vk, const uint32_t* vc, const void* img,
                           const GridC& g, unsigned long long* count, unsigned long long* sums, double* fsum) {
    const T* im = (const T*)img;
    if (planes && mag && bin) launch_fused_sip<T, C, true, true, true>(sip, grid, st, p, vk, vc, im, g, count, sums, fsum);
    else if (planes && bin) launch_fused_sip<T, C, true, false, true>(sip, grid, st, p, vk, vc, im, g, count, sums, fsum);
    else if (bin) launch_fused_sip<T, C, false, false, true>(sip, grid, st, p, vk, vc, im, g, count, sums, fsum);
    else if (mag) launch_fused_sip<T, 1, true, true, false>(sip, grid, st, p, vk, vc, (const T*)nullptr, g, count, sums, fsum);
    else launch_fused_sip<T, 1, true, false, false>(sip, grid, st, p, vk, vc, (const T*)nullptr, g, count, sums, fsum);
}

template <typename T>
static void launch_fused(int channels, bool planes, bool mag, bool bin, bool sip, dim3 grid, cudaStream_t st,
                         const GeorefParams& p, const uint32_t* vk, const uint32_t* vc, const void* img,
                         const GridC& g, unsigned long long* count, unsigned long long* sums, double* fsum) {
    if (!bin) channels = 1;
    switch (channels) {
        case 1: launch_fused_c<T, 1>(planes, mag, bin, sip, grid, st, p, vk, vc, img, g, count, sums, fsum); break;
        case 2: launch_fused_c<T, 2>(planes, mag, bin, sip, grid, st, p, vk, vc, img, g, count, sums, fsum); break;
        case 3: launch_fused_c<T, 3>(planes, mag, bin, sip, grid, st, p, vk, vc, img, g, count, sums, fsum); break;
        case 4: launch_fused_c<T, 4>(planes, mag, bin, sip, grid, st, p, vk, vc, img, g, count, sums, fsum); break;
    }
}

// Planes and / or binning of one WCS frame from its final validity bitmaps.
//   out == NULL        no coordinate plane is written (plane-free resampling);
//   out != NULL        lat/lon corner + centre planes and the elevation plane are required, the four
//                      MLat/MLT planes are written iff all four are given; the bitmaps in `out` are ignored;
//   grid == NULL       no binning (d_img, d_count, d_sums, d_fsum unused).
static int georef_fused(amt_ctx* ctx, const amt_frame* frame, const amt_georef_out* out, const uint32_t* d_valid_k,
                        const uint32_t* d_valid_c, const void* d_img, int32_t dtype, int32_t channels,
                        const amt_grid* grid, uint64_t* d_count, uint64_t* d_sums, double* d_fsum, cudaStream_t st) {
    CHECK_ARG(frame && d_valid_c, "amt_georef_fused: NULL argument");
    CHECK_ARG(out || grid, "amt_georef_fused: neither planes nor a grid requested");
    if (frame->model != AMT_MODEL_WCS || frame->fast_center)
        return set_err(AMT_ERR_UNSUPPORTED, "amt_georef_fused: WCS frames with fast_center == 0 only "
                                            "(fast centres need the corner intersection points)");
    const bool planes = out != nullptr, bin = grid != nullptr;
    bool mag = false;
    if (planes) {
        CHECK_ARG(d_valid_k, "amt_georef_fused: the corner bitmap is required with planes");
        CHECK_ARG(out->d_lat_k && out->d_lon_k && out->d_lat_c && out->d_lon_c && out->d_elev_c,
                  "amt_georef_fused: lat/lon corner + centre planes and the elevation plane are required");
        const int nm = (out->d_mlat_k != nullptr) + (out->d_mlt_k != nullptr) + (out->d_mlat_c != nullptr) +
                       (out->d_mlt_c != nullptr);
        CHECK_ARG(nm == 0 || nm == 4, "amt_georef_fused: MLat/MLT planes come as a set of four");
        mag = nm == 4;
    }
    if (bin) {
        CHECK_ARG(d_img && d_count && d_sums, "amt_georef_fused: image and accumulators are required with a grid");
        CHECK_ARG(channels >= 1 && channels <= 4, "amt_georef_fused: channels must be 1..4");
        if (dtype != AMT_U8 && dtype != AMT_U16)
            return set_err(AMT_ERR_UNSUPPORTED, "amt_georef_fused: image dtype must be uint8 or uint16");
    }
    GeorefParams p;
    memset(&p, 0, sizeof p);
    int rc = fill_frame(frame, p);
    if (rc) return rc;
    if (planes) p.o = *out;
    GridC g;
    memset(&g, 0, sizeof g);
    if (bin) {
        rc = fill_grid(grid, g);
        if (rc) return rc;
    }
    const bool sip = frame->sip_order_a != 0 || frame->sip_order_b != 0;
    const int W = frame->width, H = frame->height;
    const int tile_rows = kFusedRows * kFusedIter;
    dim3 lg(((planes ? W + 1 : W) + kFusedCols - 1) / kFusedCols, ((planes ? H + 1 : H) + tile_rows - 1) / tile_rows);
    p.row_stride = golden_stride(lg.y);
    unsigned long long* cnt = (unsigned long long*)d_count;
    unsigned long long* sm = (unsigned long long*)d_sums;
    if (bin && dtype == AMT_U16)
        launch_fused<unsigned short>(channels, planes, mag, bin, sip, lg, st, p, d_valid_k, d_valid_c, d_img, g, cnt, sm, d_fsum);
    else
        launch_fused<unsigned char>(channels, planes, mag, bin, sip, lg, st, p, d_valid_k, d_valid_c, d_img, g, cnt, sm, d_fsum);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

extern "C" int amt_georef_fused(amt_ctx* ctx, const amt_frame* frame, const amt_georef_out* out,
                                const uint32_t* d_valid_k, const uint32_t* d_valid_c, const void* d_img,
                                int32_t dtype, int32_t channels, const amt_grid* grid, uint64_t* d_count,
                                uint64_t* d_sums, double* d_fsum, void* stream) {
    ENTER(ctx);
    return georef_fused(ctx, frame, out, d_valid_k, d_valid_c, d_img, dtype, channels, grid, d_count, d_sums, d_fsum,
                        (cudaStream_t)stream);
}

extern "C" int amt_georef_bin_fused(amt_ctx* ctx, const amt_frame* frame, const uint32_t* d_valid_c,
                                    const void* d_img, int32_t dtype, int32_t channels, const amt_grid* grid,
                                    uint64_t* d_count, uint64_t* d_sums, double* d_fsum, void* stream) {
    ENTER(ctx);
    CHECK_ARG(grid, "amt_georef_bin_fused: NULL argument");
    return georef_fused(ctx, frame, nullptr, nullptr, d_valid_c, d_img, dtype, channels, grid, d_count, d_sums, d_fsum,
                        (cudaStream_t)stream);
}

// ------------------------------------------------------------------ SIP known-answer access
// The FITS-SIP forward polynomial of a frame evaluated for n (u, v) pixel offsets:
// (u', v') = (u + sum A_pq u^p v^q, v + sum B_pq u^p v^q) -- exactly the device function the
// georeference kernels call, exposed so that tests can pin it against exact rational arithmetic.
__global__ void k_sip_distort(const __grid_constant__ GeorefParams p, const double* __restrict__ u,
                              const double* __restrict__ v, size_t n, double* __restrict__ uo, double* __restrict__ vo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = u[i], b = v[i];
    if (p.f.sip_oa | p.f.sip_ob) sip_distort(p.sip_a, p.f.sip_oa, p.sip_b, p.f.sip_ob, a, b);
    uo[i] = a;
    vo[i] = b;
}

extern "C" int amt_sip_distort(amt_ctx* ctx, const amt_frame* frame, const double* d_u, const double* d_v, size_t n,
                               double* d_u_out, double* d_v_out, void* stream) {
    ENTER(ctx);
    CHECK_ARG(frame && d_u && d_v && d_u_out && d_v_out, "amt_sip_distort: NULL argument");
    if (n == 0) return AMT_OK;
    GeorefParams p;
    memset(&p, 0, sizeof p);
    int rc = fill_frame(frame, p);
    if (rc) return rc;
    k_sip_distort<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, d_u, d_v, n, d_u_out, d_v_out);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// ============================================== host arithmetic (grid derivation) + sequence engine
#include "amt_host.cuh"
#include "amt_seq.cuh"
