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
#include "amt_math.cuh"

using namespace amt;

// ============================================================================ context
struct amt_ctx {
    int device;
    int sm_count;
    unsigned long long launches;
    void* scratch;
    size_t scratch_bytes;
};

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

static int ensure_scratch(amt_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return AMT_OK;
    if (ctx->scratch) CUDA_TRY(cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes + bytes / 4;
    CUDA_TRY(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
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
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    *out = c;
    return AMT_OK;
}

extern "C" int amt_ctx_destroy(amt_ctx* ctx) {
    if (!ctx) return AMT_OK;
    cudaSetDevice(ctx->device);
    if (ctx->scratch) cudaFree(ctx->scratch);
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

// ===================================================================== georeference
struct GeorefParams {
    FrameC f;
    double sip_a[AMT_SIP_MAX_COEF];
    double sip_b[AMT_SIP_MAX_COEF];
    amt_georef_out o;
    unsigned long long* ill;     // n_ill_conditioned counter (nullable)
    unsigned int nbk;            // blocks that handle corners (flat kernel)
};

// A ray "grazes" when its normalised discriminant q = rootTerm/dDD = 1 - (perpendicular miss
// distance / ellipsoid radius)^2 is below this: 1 ulp of direction noise then moves the
// intersection by ~3e-15/sqrt(q) degrees, i.e. more than 1e-9 deg for q < 1e-11.
constexpr double kIllThreshold = 1e-10;

// One intersection point -> all requested outputs at flat index i.
__device__ __forceinline__ void emit_point(const GeorefParams& p, const double P[3], size_t i,
                                           double* __restrict__ lat_o, double* __restrict__ lon_o,
                                           double* __restrict__ mlat_o, double* __restrict__ mlt_o) {
    if (lat_o || lon_o) {
        double lat, lon;
        point_to_geo(p.f, P, lat, lon);
        if (lat_o) lat_o[i] = lat;
        if (lon_o) lon_o[i] = lon;
    }
    if (mlat_o || mlt_o) {
        double mlat, mlt;
        point_to_mag(p.f, P, mlat, mlt);
        if (mlat_o) mlat_o[i] = mlat;
        if (mlt_o) mlt_o[i] = mlt;
    }
}

// fastCenterCalculation == False: corners and centres are independent points
// (mapping/astrometry.py:49-64,86-106).  Blocks [0, nbk) process corners, the rest centres.
__global__ void __launch_bounds__(256) k_georef_points(const __grid_constant__ GeorefParams p) {
    const bool corner = blockIdx.x < p.nbk;
    const int W = p.f.W, H = p.f.H;
    const int rowlen = corner ? W + 1 : W;
    const size_t n = corner ? (size_t)(W + 1) * (H + 1) : (size_t)W * H;
    const size_t i = (size_t)(corner ? blockIdx.x : blockIdx.x - p.nbk) * blockDim.x + threadIdx.x;
    bool ill = false;
    if (i < n) {
        const int y = (int)(i / rowlen);
        const int x = (int)(i - (size_t)y * rowlen);
        // wcs.py:41-44: corner grids start at -0.5
        const double px = corner ? (double)x - 0.5 : (double)x;
        const double py = corner ? (double)y - 0.5 : (double)y;
        double dir[3], P[3];
        pix2dir(p.f, p.sip_a, p.sip_b, px, py, dir);
        const double q = intersect(p.f, dir, P);
        ill = q >= 0.0 && q < kIllThreshold;
        if (corner) {
            emit_point(p, P, i, p.o.d_lat_k, p.o.d_lon_k, p.o.d_mlat_k, p.o.d_mlt_k);
        } else {
            emit_point(p, P, i, p.o.d_lat_c, p.o.d_lon_c, p.o.d_mlat_c, p.o.d_mlt_c);
            if (p.o.d_elev_c) p.o.d_elev_c[i] = elevation_deg(dir, P);
        }
    }
    if (p.ill) {
        const unsigned m = __ballot_sync(0xffffffffu, ill);
        if (m && (threadIdx.x & 31) == 0) atomicAdd(p.ill, (unsigned long long)__popc(m));
    }
}

// fastCenterCalculation == True: a CTA evaluates a (TH+1)x(TW+1) patch of corner rays into
// shared memory, then derives each centre from the mean of its 4 corner intersection points
// and (un-normalised) directions: mapping/astrometry.py:154-160 (`_calcCenters`).
constexpr int TW = 32, TH = 8;
__global__ void __launch_bounds__(TW* TH) k_georef_tiles(const __grid_constant__ GeorefParams p) {
    __shared__ double sP[3][TH + 1][TW + 1];
    __shared__ double sD[3][TH + 1][TW + 1];
    const int W = p.f.W, H = p.f.H;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.y * TW + threadIdx.x;
    bool ill = false;
    for (int c = tid; c < (TW + 1) * (TH + 1); c += TW * TH) {
        const int cy = c / (TW + 1), cx = c - cy * (TW + 1);
        const int x = x0 + cx, y = y0 + cy;
        if (x <= W && y <= H) {
            double dir[3], P[3];
            pix2dir(p.f, p.sip_a, p.sip_b, (double)x - 0.5, (double)y - 0.5, dir);
            const double q = intersect(p.f, dir, P);
#pragma unroll
            for (int k = 0; k < 3; ++k) { sP[k][cy][cx] = P[k]; sD[k][cy][cx] = dir[k]; }
            // each corner is owned by exactly one tile for the global outputs / counters
            const bool own = (cx < TW || x == W) && (cy < TH || y == H);
            if (own) {
                ill |= q >= 0.0 && q < kIllThreshold;
                const size_t i = (size_t)y * (W + 1) + x;
                emit_point(p, P, i, p.o.d_lat_k, p.o.d_lon_k, p.o.d_mlat_k, p.o.d_mlt_k);
            }
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < W && y < H) {
        const int cx = threadIdx.x, cy = threadIdx.y;
        double P[3], dir[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // corners[:-1,:-1] + corners[:-1,1:]; += corners[1:,1:]; += corners[1:,:-1]; /= 4
            double s = sP[k][cy][cx] + sP[k][cy][cx + 1];
            s = s + sP[k][cy + 1][cx + 1];
            s = s + sP[k][cy + 1][cx];
            P[k] = s / 4.0;
            double d = sD[k][cy][cx] + sD[k][cy][cx + 1];
            d = d + sD[k][cy + 1][cx + 1];
            d = d + sD[k][cy + 1][cx];
            dir[k] = d / 4.0;
        }
        const size_t i = (size_t)y * W + x;
        emit_point(p, P, i, p.o.d_lat_c, p.o.d_lon_c, p.o.d_mlat_c, p.o.d_mlt_c);
        if (p.o.d_elev_c) p.o.d_elev_c[i] = elevation_deg(dir, P);
    }
    if (p.ill) {
        const unsigned m = __ballot_sync(0xffffffffu, ill);
        if (m && (tid & 31) == 0) atomicAdd(p.ill, (unsigned long long)__popc(m));
    }
}

static int fill_frame(const amt_frame* fr, GeorefParams& p) {
    CHECK_ARG(fr->width > 0 && fr->height > 0, "amt_frame: width/height must be positive");
    CHECK_ARG((long long)(fr->width + 1) * (fr->height + 1) < (1LL << 40), "amt_frame: frame too large");
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
    f.e2a = e2 * a;
    f.d = num / b;
    f.sip_oa = fr->sip_order_a; f.sip_ob = fr->sip_order_b;
    memcpy(p.sip_a, fr->sip_a, sizeof p.sip_a);
    memcpy(p.sip_b, fr->sip_b, sizeof p.sip_b);
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
    if (frame->fast_center) {
        dim3 grid((W + TW) / TW, (H + TH) / TH);   // covers x<=W, y<=H
        k_georef_tiles<<<grid, dim3(TW, TH), 0, st>>>(p);
    } else {
        const size_t nk = (size_t)(W + 1) * (H + 1), nc = (size_t)W * H;
        const bool want_k = out->d_lat_k || out->d_lon_k || out->d_mlat_k || out->d_mlt_k;
        const bool want_c = out->d_lat_c || out->d_lon_c || out->d_mlat_c || out->d_mlt_c || out->d_elev_c;
        const unsigned nbk = want_k ? (unsigned)((nk + 255) / 256) : 0;
        const unsigned nbc = want_c ? (unsigned)((nc + 255) / 256) : 0;
        p.nbk = nbk;
        if (nbk + nbc == 0) return AMT_OK;
        k_georef_points<<<nbk + nbc, 256, 0, st>>>(p);
    }
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// ============================================== target grid / pre-rotation (shared)
struct GridC {
    int nx, ny, prerotate;
    double lo_x, hi_x, step_x, inv_step_x, round_x;
    double lo_y, hi_y, step_y, inv_step_y, round_y;
    double altitude, a, b, e2, e2a, d;
    double rot[9];
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
    c.altitude = g->altitude; c.a = g->wgs_a; c.b = g->wgs_b;
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
        bowring(g.a, g.b, g.e2a, g.d, R[0], R[1], R[2], l2, o2);
        la = l2 * kRad2Deg;
        lo = o2 * kRad2Deg;
    }
}

// ============================================================== sanitize + bbox stats
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
};

__global__ void k_stats_init(StatKeys* s) {
    s->lat_min = s->lon_min = s->lon_min_pos = ~0ULL;
    s->lat_max = s->lon_max = s->lon_max_neg = 0ULL;
    s->n_valid_k = s->n_boundary = s->n_valid_c = 0ULL;
    s->pole_flags = 0u;
}

__global__ void k_stats_final(const StatKeys* s, amt_stats* out) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    out->lat_min = s->lat_min == ~0ULL ? inf : dunkey(s->lat_min);
    out->lon_min = s->lon_min == ~0ULL ? inf : dunkey(s->lon_min);
    out->lon_min_pos = s->lon_min_pos == ~0ULL ? inf : dunkey(s->lon_min_pos);
    out->lat_max = s->lat_max == 0ULL ? -inf : dunkey(s->lat_max);
    out->lon_max = s->lon_max == 0ULL ? -inf : dunkey(s->lon_max);
    out->lon_max_neg = s->lon_max_neg == 0ULL ? -inf : dunkey(s->lon_max_neg);
    out->n_valid_corners = s->n_valid_k;
    out->n_boundary_corners = s->n_boundary;
    out->n_valid_centers = s->n_valid_c;
    out->pole_flags = s->pole_flags;
}

__device__ __forceinline__ bool isnan_d(double v) { return !(v == v); }

// _doSanitize step 1 (mapping.py:1082-1093): corner masked if itself NaN or all (<=4)
// neighbouring centres are missing.  Writes a byte mask.
__global__ void k_sanitize_corner_mask(int W, int H, const double* __restrict__ lat_k,
                                       const double* __restrict__ lat_c, unsigned char* __restrict__ mk) {
    const size_t n = (size_t)(W + 1) * (H + 1);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int y = (int)(i / (W + 1)), x = (int)(i - (size_t)y * (W + 1));
    bool all_missing = true;
#pragma unroll
    for (int dy = -1; dy <= 0; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 0; ++dx) {
            const int cy = y + dy, cx = x + dx;
            if (cy >= 0 && cy < H && cx >= 0 && cx < W) all_missing &= isnan_d(lat_c[(size_t)cy * W + cx]);
        }
    mk[i] = (isnan_d(lat_k[i]) || all_missing) ? 1 : 0;
}

// step 2 (mapping.py:1095-1104): centre masked if any of its 4 corners is masked.
__global__ void k_sanitize_centers(int W, int H, const unsigned char* __restrict__ mk, amt_georef_out o) {
    const size_t n = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
    const size_t k = (size_t)y * (W + 1) + x;
    const bool any = mk[k] | mk[k + 1] | mk[k + W + 1] | mk[k + W + 2];
    if (any) {
        const double nan = qnan();
        if (o.d_lat_c) o.d_lat_c[i] = nan;
        if (o.d_lon_c) o.d_lon_c[i] = nan;
        if (o.d_mlat_c) o.d_mlat_c[i] = nan;
        if (o.d_mlt_c) o.d_mlt_c[i] = nan;
        if (o.d_elev_c) o.d_elev_c[i] = nan;
    }
}

// step 3 (mapping.py:1106-1117): corners again, using the updated centre mask.
__global__ void k_sanitize_corners(int W, int H, const unsigned char* __restrict__ mk,
                                   const double* __restrict__ lat_c, amt_georef_out o) {
    const size_t n = (size_t)(W + 1) * (H + 1);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int y = (int)(i / (W + 1)), x = (int)(i - (size_t)y * (W + 1));
    bool all_missing = true;
#pragma unroll
    for (int dy = -1; dy <= 0; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 0; ++dx) {
            const int cy = y + dy, cx = x + dx;
            if (cy >= 0 && cy < H && cx >= 0 && cx < W) all_missing &= isnan_d(lat_c[(size_t)cy * W + cx]);
        }
    if ((mk && mk[i]) || all_missing) {
        const double nan = qnan();
        if (o.d_lat_k) o.d_lat_k[i] = nan;
        if (o.d_lon_k) o.d_lon_k[i] = nan;
        if (o.d_mlat_k) o.d_mlat_k[i] = nan;
        if (o.d_mlt_k) o.d_mlt_k[i] = nan;
    }
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}

// bbox min/max over the boundary corners + valid counts (mapping.py:694-703,729-737).
// Blocks [0,nbk) scan corners, the rest count valid centres.
__global__ void __launch_bounds__(256) k_stats(int W, int H, unsigned nbk, const double* __restrict__ lat_k,
                                               const double* __restrict__ lon_k,
                                               const double* __restrict__ lat_c, const __grid_constant__ GridC g,
                                               StatKeys* s) {
    const int lane = threadIdx.x & 31;
    if (blockIdx.x >= nbk) {
        const size_t n = (size_t)W * H;
        const size_t i = (size_t)(blockIdx.x - nbk) * blockDim.x + threadIdx.x;
        const bool v = i < n && !isnan_d(lat_c[i]);
        const unsigned m = __ballot_sync(0xffffffffu, v);
        if (m && lane == 0) atomicAdd(&s->n_valid_c, (unsigned long long)__popc(m));
        if (v) {
            // Pole test: the longitudes of the 4 corners of a valid pixel wind once around
            // (+-360 deg) iff the quad encloses a pole.
            const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
            const size_t k = (size_t)y * (W + 1) + x;
            const double l0 = lon_k[k], l1 = lon_k[k + 1], l2 = lon_k[k + W + 2], l3 = lon_k[k + W + 1];
            double w = 0.0;
            const double dl[4] = {l1 - l0, l2 - l1, l3 - l2, l0 - l3};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double d = dl[q];
                if (d > 180.0) d -= 360.0;
                if (d <= -180.0) d += 360.0;
                w += d;
            }
            if (fabs(w) > 180.0) atomicOr(&s->pole_flags, lat_k[k] > 0.0 ? 1u : 2u);
        }
        return;
    }
    const size_t n = (size_t)(W + 1) * (H + 1);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false, boundary = false;
    double la = 0, lo = 0;
    if (i < n) {
        la = lat_k[i];
        valid = !isnan_d(la);
        if (valid) {
            lo = lon_k[i];
            const int y = (int)(i / (W + 1)), x = (int)(i - (size_t)y * (W + 1));
            boundary = x == 0 || y == 0 || x == W || y == H;
            if (!boundary)
                boundary = isnan_d(lat_k[i - 1]) || isnan_d(lat_k[i + 1]) ||
                           isnan_d(lat_k[i - (W + 1)]) || isnan_d(lat_k[i + (W + 1)]);
        }
    }
    const unsigned mv = __ballot_sync(0xffffffffu, valid);
    const unsigned mb = __ballot_sync(0xffffffffu, boundary);
    if (mv && lane == 0) atomicAdd(&s->n_valid_k, (unsigned long long)__popc(mv));
    if (!mb) return;
    if (boundary && g.prerotate != AMT_PRE_NONE) prerotate(g, la, lo);
    unsigned long long kmin_la = boundary ? dkey(la) : ~0ULL, kmax_la = boundary ? dkey(la) : 0ULL;
    unsigned long long kmin_lo = boundary ? dkey(lo) : ~0ULL, kmax_lo = boundary ? dkey(lo) : 0ULL;
    unsigned long long kmin_pos = (boundary && lo > 0.0) ? dkey(lo) : ~0ULL;
    unsigned long long kmax_neg = (boundary && !(lo > 0.0)) ? dkey(lo) : 0ULL;
    kmin_la = warp_min_u64(kmin_la); kmax_la = warp_max_u64(kmax_la);
    kmin_lo = warp_min_u64(kmin_lo); kmax_lo = warp_max_u64(kmax_lo);
    kmin_pos = warp_min_u64(kmin_pos); kmax_neg = warp_max_u64(kmax_neg);
    if (lane == 0) {
        atomicAdd(&s->n_boundary, (unsigned long long)__popc(mb));
        atomicMin(&s->lat_min, kmin_la); atomicMax(&s->lat_max, kmax_la);
        atomicMin(&s->lon_min, kmin_lo); atomicMax(&s->lon_max, kmax_lo);
        if (kmin_pos != ~0ULL) atomicMin(&s->lon_min_pos, kmin_pos);
        if (kmax_neg != 0ULL) atomicMax(&s->lon_max_neg, kmax_neg);
    }
}

extern "C" int amt_sanitize(amt_ctx* ctx, int32_t W, int32_t H, const amt_georef_out* planes, void* stream) {
    ENTER(ctx);
    CHECK_ARG(planes && W > 0 && H > 0, "amt_sanitize: bad arguments");
    CHECK_ARG(planes->d_lat_k && planes->d_lat_c, "amt_sanitize: lat_k and lat_c planes are required");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nk = (size_t)(W + 1) * (H + 1), nc = (size_t)W * H;
    int rc = ensure_scratch(ctx, nk);
    if (rc) return rc;
    unsigned char* mk = (unsigned char*)ctx->scratch;
    const unsigned nbk = (unsigned)((nk + 255) / 256), nbc = (unsigned)((nc + 255) / 256);
    k_sanitize_corner_mask<<<nbk, 256, 0, st>>>(W, H, planes->d_lat_k, planes->d_lat_c, mk);
    LAUNCH_CHECK(ctx);
    k_sanitize_centers<<<nbc, 256, 0, st>>>(W, H, mk, *planes);
    LAUNCH_CHECK(ctx);
    k_sanitize_corners<<<nbk, 256, 0, st>>>(W, H, mk, planes->d_lat_c, *planes);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

extern "C" int amt_bbox_stats(amt_ctx* ctx, int32_t W, int32_t H, const double* d_lat_k, const double* d_lon_k,
                              const double* d_lat_c, const amt_grid* pre, amt_stats* d_stats, void* stream) {
    ENTER(ctx);
    CHECK_ARG(W > 0 && H > 0 && d_lat_k && d_lon_k && d_stats, "amt_bbox_stats: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    GridC g;
    memset(&g, 0, sizeof g);
    if (pre) {
        int rc = fill_grid(pre, g, true);
        if (rc) return rc;
    }
    const size_t nk = (size_t)(W + 1) * (H + 1), nc = (size_t)W * H;
    // the key block lives behind the sanitize mask in the scratch arena
    const size_t off = (nk + 255) / 256 * 256;
    int rc = ensure_scratch(ctx, off + sizeof(StatKeys));
    if (rc) return rc;
    StatKeys* keys = (StatKeys*)((unsigned char*)ctx->scratch + off);
    const unsigned nbk = (unsigned)((nk + 255) / 256), nbc = d_lat_c ? (unsigned)((nc + 255) / 256) : 0;
    k_stats_init<<<1, 1, 0, st>>>(keys);
    LAUNCH_CHECK(ctx);
    k_stats<<<nbk + nbc, 256, 0, st>>>(W, H, nbk, d_lat_k, d_lon_k, d_lat_c, g, keys);
    LAUNCH_CHECK(ctx);
    k_stats_final<<<1, 1, 0, st>>>(keys, d_stats);
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

// createMasked / maskedByElevation (mapping.py:845-864,1171-1231) on the device planes.
__global__ void k_apply_center_mask(size_t n, const unsigned char* __restrict__ mask, double min_elev,
                                    amt_georef_out o) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool m = mask && mask[i];
    if (min_elev == min_elev) m |= !(o.d_elev_c[i] >= min_elev);   // (elevation < min).filled(True)
    if (m) {
        const double nan = qnan();
        if (o.d_lat_c) o.d_lat_c[i] = nan;
        if (o.d_lon_c) o.d_lon_c[i] = nan;
        if (o.d_mlat_c) o.d_mlat_c[i] = nan;
        if (o.d_mlt_c) o.d_mlt_c[i] = nan;
        if (o.d_elev_c) o.d_elev_c[i] = nan;
    }
}

extern "C" int amt_apply_center_mask(amt_ctx* ctx, int32_t W, int32_t H, const uint8_t* d_mask,
                                     double min_elevation, const amt_georef_out* planes, void* stream) {
    ENTER(ctx);
    CHECK_ARG(planes && W > 0 && H > 0, "amt_apply_center_mask: bad arguments");
    CHECK_ARG(planes->d_lat_k && planes->d_lat_c, "amt_apply_center_mask: lat_k and lat_c planes are required");
    CHECK_ARG(!(min_elevation == min_elevation) || planes->d_elev_c, "amt_apply_center_mask: elevation plane required");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nk = (size_t)(W + 1) * (H + 1), nc = (size_t)W * H;
    k_apply_center_mask<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(nc, d_mask, min_elevation, *planes);
    LAUNCH_CHECK(ctx);
    k_sanitize_corners<<<(unsigned)((nk + 255) / 256), 256, 0, st>>>(W, H, nullptr, planes->d_lat_c, *planes);
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

// ================================================================== stage 3: binning
// Flat cell (row 0 = northernmost) of a centre coordinate, or -1.
template <bool NEAR>
__device__ __forceinline__ int cell_of(const GridC& g, double la, double lo, int& ix, int& iy, bool& near) {
    near = false;
    ix = iy = -1;
    if (!(la == la)) return -1;                       // resample.py:316: dropped iff lat is NaN
    prerotate(g, la, lo);
    bool nx_, ny_;
    ix = bin_index<NEAR>(lo, g.lo_x, g.hi_x, g.step_x, g.inv_step_x, g.nx, g.round_x, nx_);
    iy = bin_index<NEAR>(la, g.lo_y, g.hi_y, g.step_y, g.inv_step_y, g.ny, g.round_y, ny_);
    near = nx_ || ny_;
    if (ix < 0 || iy < 0) return -1;
    return (g.ny - 1 - iy) * g.nx + ix;               // flipud, resample.py:349
}

// Accumulate one warp's samples: consecutive lanes that hit the same cell form a run; a
// segmented warp scan sums each run and only its last lane issues the atomics.
template <int C>
__device__ __forceinline__ void warp_accumulate(int cell, const unsigned (&val)[C], double side, bool has_side,
                                                unsigned long long* __restrict__ count,
                                                unsigned long long* __restrict__ sums,
                                                double* __restrict__ fsum, size_t plane) {
    const unsigned lane = threadIdx.x & 31;
    const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
    const bool head = lane == 0 || prev != cell;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned le = 0xffffffffu >> (31 - lane);
    const int start = 31 - __clz(heads & le);         // first lane of my run
    const int next = __shfl_down_sync(0xffffffffu, cell, 1);
    const bool tail = lane == 31 || next != cell;
    unsigned v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = val[c];
    double s = side;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const bool take = (int)lane - d >= start;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const unsigned t = __shfl_up_sync(0xffffffffu, v[c], d);
            if (take) v[c] += t;
        }
        if (has_side) {
            const double t = __shfl_up_sync(0xffffffffu, s, d);
            if (take) s += t;
        }
    }
    if (tail && cell >= 0) {
        atomicAdd(&count[cell], (unsigned long long)(lane - start + 1));
#pragma unroll
        for (int c = 0; c < C; ++c) atomicAdd(&sums[(size_t)c * plane + cell], (unsigned long long)v[c]);
        if (has_side) atomicAdd(&fsum[cell], s);
    }
}

template <typename T, int C, bool NEAR>
__global__ void __launch_bounds__(256) k_bin(const double* __restrict__ lat, const double* __restrict__ lon,
                                             const double* __restrict__ side, const T* __restrict__ img, size_t n,
                                             const __grid_constant__ GridC g, unsigned long long* __restrict__ count,
                                             unsigned long long* __restrict__ sums, double* __restrict__ fsum,
                                             unsigned long long* __restrict__ near_counter) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int cell = -1, ix, iy;
    bool near = false;
    unsigned val[C];
#pragma unroll
    for (int c = 0; c < C; ++c) val[c] = 0;
    double s = 0.0;
    if (i < n) {
        cell = cell_of<NEAR>(g, lat[i], lon[i], ix, iy, near);
        if (cell >= 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) val[c] = img[i * C + c];
            if (side) s = side[i];
        }
    }
    warp_accumulate<C>(cell, val, s, side != nullptr, count, sums, fsum, (size_t)g.nx * g.ny);
    if (NEAR) {
        const unsigned m = __ballot_sync(0xffffffffu, near);
        if (m && (threadIdx.x & 31) == 0) atomicAdd(near_counter, (unsigned long long)__popc(m));
    }
}

template <typename T, bool NEAR>
static void launch_bin(int channels, unsigned blocks, cudaStream_t st, const double* lat, const double* lon,
                       const double* side, const void* img, size_t n, const GridC& g, unsigned long long* count,
                       unsigned long long* sums, double* fsum, unsigned long long* near) {
    const T* im = (const T*)img;
    switch (channels) {
        case 1: k_bin<T, 1, NEAR><<<blocks, 256, 0, st>>>(lat, lon, side, im, n, g, count, sums, fsum, near); break;
        case 2: k_bin<T, 2, NEAR><<<blocks, 256, 0, st>>>(lat, lon, side, im, n, g, count, sums, fsum, near); break;
        case 3: k_bin<T, 3, NEAR><<<blocks, 256, 0, st>>>(lat, lon, side, im, n, g, count, sums, fsum, near); break;
        case 4: k_bin<T, 4, NEAR><<<blocks, 256, 0, st>>>(lat, lon, side, im, n, g, count, sums, fsum, near); break;
    }
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
    const unsigned blocks = (unsigned)((n_pixels + 255) / 256);
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
template <typename T>
__global__ void k_normalise(size_t cells, int C, const unsigned long long* __restrict__ count,
                            const unsigned long long* __restrict__ sums, const double* __restrict__ fsum,
                            T* __restrict__ out_img, unsigned char* __restrict__ out_mask, double* __restrict__ out_side) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const unsigned long long n = count[i];
    if (out_mask) out_mask[i] = n == 0;
    const double dn = (double)n;
    for (int c = 0; c < C; ++c) {
        T v = 0;
        if (n) v = (T)rint((double)sums[(size_t)c * cells + i] / dn);
        out_img[i * C + c] = v;
    }
    if (out_side) out_side[i] = n ? fsum[i] / dn : qnan();
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
                                                          (unsigned char*)d_out_img, d_out_mask, d_out_side);
    else if (dtype == AMT_U16)
        k_normalise<unsigned short><<<blocks, 256, 0, st>>>(cells, channels, (const unsigned long long*)d_count,
                                                           (const unsigned long long*)d_sums, d_fsum,
                                                           (unsigned short*)d_out_img, d_out_mask, d_out_side);
    else
        return set_err(AMT_ERR_UNSUPPORTED, "amt_normalise: image dtype must be uint8 or uint16");
    LAUNCH_CHECK(ctx);
    return AMT_OK;
}

extern "C" int amt_georef_bin_fused(amt_ctx* ctx, const amt_frame* frame, const void* d_img, int32_t dtype,
                                    int32_t channels, const amt_grid* grid, uint64_t* d_count,
                                    uint64_t* d_sums, double* d_fsum, void* stream) {
    (void)ctx; (void)frame; (void)d_img; (void)dtype; (void)channels; (void)grid; (void)d_count; (void)d_sums;
    (void)d_fsum; (void)stream;
    return set_err(AMT_ERR_UNSUPPORTED, "amt_georef_bin_fused: not built yet");
}
