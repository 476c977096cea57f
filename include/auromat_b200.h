/*
 * auromat_b200 -- C ABI of the B200 (sm_100a) georeference + regrid hot path.
 *
 * The reference (esa/auromat v1.0.8) is pure Python and has no FFI; its boundary for this
 * path is the Python API `auromat.mapping.spacecraft.getMapping` + `auromat.resample.resample`.
 * This header declares the flat C entry points a maintainer would bind (ctypes) to replace
 * the numpy array passes underneath that API.  Each entry point cites the reference
 * interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - every function returns an `amt_status` (0 == AMT_OK); `amt_last_error()` returns a
 *     thread-local, human readable message for the last non-zero status;
 *   - all pointers named `d_*` are DEVICE pointers, all pointers named `h_*` are HOST
 *     pointers; the library never returns memory it owns except through amt_alloc_* ;
 *   - all kernels are enqueued on the `stream` argument (a `cudaStream_t` passed as void*,
 *     NULL == legacy default stream) and the functions do NOT synchronise unless stated;
 *   - arrays are C-contiguous, row-major, the layout numpy gives the reference's arrays;
 *   - misses (no ray/ellipsoid intersection) are NaN, exactly as in the reference
 *     (`numpy.ma.masked_invalid` is applied by the Python wrapper).
 */
#ifndef AUROMAT_B200_H
#define AUROMAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMT_ABI_VERSION 3

typedef enum amt_status {
    AMT_OK = 0,
    AMT_ERR_INVALID_ARGUMENT = 1, /* -> ValueError / AssertionError in the wrapper      */
    AMT_ERR_UNSUPPORTED = 2,      /* -> NotImplementedError (projection, method, dtype)  */
    AMT_ERR_CUDA = 3,             /* -> RuntimeError with the CUDA error string          */
    AMT_ERR_NO_DEVICE = 4         /* -> RuntimeError: no CUDA device / wrong arch        */
} amt_status;

/* ----------------------------------------------------------------------------------------
 * Per-frame constants.  Everything in here is computed ON THE HOST by the wrapper exactly
 * as the reference computes it (scalar Python/numpy), once per frame:
 *   crpix, cd          coordinates/wcs.py:83-90    header CRPIX1/2, CD1_1..CD2_2
 *   rot                coordinates/wcs.py:135-139  euler_matrix(..., 'rzxz')[:3,:3]
 *   sip_*              FITS-SIP forward polynomial (reference: astropy fallback, wcs.py:54-56)
 *   cam                mapping/mapping.py:318      cameraPosGCRS [km]
 *   inv_axes           coordinates/intersection.py:66  (1/a, 1/a, 1/b) of the inflated ellipsoid,
 *                      a = wgs84A + altitude, b = wgs84B + altitude (mapping/mapping.py:1497-1500)
 *   origin_inside      coordinates/intersection.py:239-241
 *   m_geo              coordinates/transform.py:683-686  mat_j2000_to_geo(et)
 *   m_sm               coordinates/transform.py:688-691  mat_j2000_to_sm(et)
 *   wgs_a, wgs_b       coordinates/geodesic.py:20-21  (un-inflated, used by Bowring)
 * -------------------------------------------------------------------------------------- */
#define AMT_SIP_MAX_ORDER 9
#define AMT_SIP_MAX_COEF 55 /* (9+1)(9+2)/2 */

typedef struct amt_frame {
    int32_t width, height;      /* IMAGEW, IMAGEH                                         */
    int32_t fast_center;        /* mapping/astrometry.py:24-40 fastCenterCalculation      */
    int32_t origin_inside;
    double crpix[2];
    double cd[4];               /* row-major 2x2                                          */
    double rot[9];              /* row-major 3x3, native -> celestial                     */
    double cam[3];
    double inv_axes[3];
    double m_geo[9];            /* row-major 3x3                                          */
    double m_sm[9];
    double wgs_a, wgs_b;
    int32_t sip_order_a;        /* 0 == no SIP                                            */
    int32_t sip_order_b;
    /* packed triangular: index(p,q) = p*(order+1) - p*(p-1)/2 + q  for p+q <= order      */
    double sip_a[AMT_SIP_MAX_COEF];
    double sip_b[AMT_SIP_MAX_COEF];
    /* Camera model.  AMT_MODEL_WCS: everything above.  AMT_MODEL_ALLSKY: ground-based fisheye
     * all-sky imager (mapping/miracle.py:240-258,314-347): zenith at (row, col) = (xc, yc) in
     * pixels, zenith distance z = d/k [rad], image rotation [rad]; `rot` then is
     * R_lon * R_lat (local -> ECEF, :249-252), `cam` the station in ECEF [km], `m_geo` the
     * identity, and the elevation plane receives the camera elevation angle 90 - z.         */
    int32_t model;
    int32_t reserved;
    double allsky_xc, allsky_yc, allsky_k, allsky_rotation;
} amt_frame;

enum { AMT_MODEL_WCS = 0, AMT_MODEL_ALLSKY = 1 };

/* Outputs of the georeference pass; any pointer may be NULL (that plane is not written).
 * Corner planes have (height+1)*(width+1) doubles, centre planes height*width.
 * Replaces the lazy array properties of mapping/astrometry.py:108-212:
 *   lat_k, lon_k   -> lats, lons                 (:138-144)   degrees
 *   mlat_k, mlt_k  -> mLatMlt                    (:170-183)   degrees / hours
 *   lat_c, lon_c   -> latsCenter, lonsCenter     (:146-152)
 *   mlat_c, mlt_c  -> mLatMltCenter              (:185-198)
 *   elev_c         -> elevation                  (:200-212)   degrees                    */
typedef struct amt_georef_out {
    double* d_lat_k;
    double* d_lon_k;
    double* d_mlat_k;
    double* d_mlt_k;
    double* d_lat_c;
    double* d_lon_c;
    double* d_mlat_c;
    double* d_mlt_c;
    double* d_elev_c;
    /* validity bitmaps, 1 bit per corner / centre (1 == defined), rows padded to 32-bit words:
     * (height+1) x ceil((width+1)/32) and height x ceil(width/32) words; padding bits are 0.
     * Written by amt_georef (ballots of the ray hit test), consumed and updated by
     * amt_sanitize, amt_apply_center_mask; read by amt_bbox_stats.                        */
    uint32_t* d_valid_k;
    uint32_t* d_valid_c;
} amt_georef_out;

/* Reductions over the corner planes (mapping/mapping.py:694-743 boundingBox min/max part;
 * "boundary" = valid corner with an invalid or out-of-array 4-neighbour, i.e. the nodes the
 * reference's outline (mapping.py:672-681, utils.py:97-139) runs through).               */
typedef struct amt_stats {
    double lat_min, lat_max;         /* over boundary corners                             */
    double lon_min, lon_max;
    double lon_min_pos, lon_max_neg; /* min over lon>0, max over lon<=0 (discontinuity)    */
    uint64_t n_valid_corners;
    uint64_t n_boundary_corners;
    uint64_t n_valid_centers;
    uint64_t n_ill_conditioned;      /* rays with tiny discriminant        */
    uint64_t pole_flags;             /* bit0: a valid pixel quad encloses the north pole,
                                        bit1: the south pole (replaces the outline/azimuth
                                        test of mapping.py:705-718, geodesic.py:183-202)   */
    int32_t row_min_c, row_max_c;    /* pixel box of the valid centres (rows / columns that  */
    int32_t col_min_c, col_max_c;    /* hold at least one defined pixel); max < min if none:
                                        only these image rows are ever read by the binning   */
} amt_stats;

/* Plate-carree target grid as seen by the binning kernel.  The wrapper derives these on the
 * host with the reference's own arithmetic (resample.py:220-241,281-299,330-335):
 *   x axis = longitude, y axis = latitude (util/histogram.py call at resample.py:337)
 *   edges_x[i] = fl(fl(i*step_x) + lo_x), edges_x[nx] = hi_x   (numpy.linspace)           */
typedef enum amt_prerotate {
    AMT_PRE_NONE = 0,
    AMT_PRE_WRAP180 = 1,   /* resample.py:203-218  lon := wrap_at_180(lon + 180)           */
    AMT_PRE_POLE = 2       /* resample.py:176-201  rotatePole(+90 deg about X)             */
} amt_prerotate;

typedef struct amt_grid {
    int32_t nx, ny;        /* number of core bins: (nLon-2), (nLat-2)                      */
    int32_t prerotate;     /* amt_prerotate                                                */
    int32_t reserved;
    double lo_x, hi_x, step_x;
    double lo_y, hi_y, step_y;
    double round_x, round_y;   /* 10**decimal of util/histogram.py:218-219                 */
    double altitude;           /* for AMT_PRE_POLE (transform.py:301-322)                  */
    double wgs_a, wgs_b;
    double rot[9];             /* rotation_matrix(+90deg, X)[:3,:3] for AMT_PRE_POLE       */
    /* Side channel (elevation sums, the float weight of the histogram2d call at resample.py:119-120,
     * 337).  0: accumulated with f64 atomics (any value incl. NaN; last bits depend on the order).
     * > 0 (a power of two): accumulated as int64 of llrint(value * side_scale) -- exact, order
     * independent, identical from run to run and across ranks; the caller picks the scale so that
     * |value| * side_scale * (number of samples) < 2^62 and guarantees finite values; `d_fsum`
     * then holds int64.                                                                          */
    double side_scale;
} amt_grid;

typedef enum amt_dtype { AMT_U8 = 0, AMT_U16 = 1 } amt_dtype;

/* ------------------------------------------------------------------------ context ---- */
typedef struct amt_ctx amt_ctx;

const char* amt_last_error(void);
int amt_abi_version(void);

/* Creates a context bound to CUDA device `device` (must be compute capability 10.x).     */
int amt_ctx_create(int device, amt_ctx** out);
int amt_ctx_destroy(amt_ctx* ctx);
int amt_ctx_device(const amt_ctx* ctx, int* device);
/* Number of kernels this context has launched so far (bench.py `gpu_launches`).           */
int amt_ctx_launch_count(const amt_ctx* ctx, uint64_t* count);

/* Measures the FP64 pipe peak of the device with a register-resident DFMA stream (8 independent
 * chains per thread, 8 CTAs of 256 threads per SM): *dfma_per_second = warp-lane DFMA/s; the FP64
 * roofline denominator of bench.py (FLOP/s = 2 x that).  Synchronises.                      */
int amt_measure_fp64_peak(amt_ctx* ctx, double* dfma_per_second);

/* Measures the L2 atomic ceiling of the scatter: u64 atomicAdd to pseudo-random words of a grid of
 * `cells` accumulators (32 distinct addresses per warp instruction, the pattern of the binning's run
 * tails): *atomics_per_second.  The denominator against which the binning's atomic rate is judged
 * (why the scatter is not privatised in shared memory).  Synchronises.                 */
int amt_measure_atomic_peak(amt_ctx* ctx, size_t cells, double* atomics_per_second);

/* Plain memory helpers so that a C caller does not need the CUDA runtime API.            */
int amt_alloc_device(amt_ctx* ctx, size_t bytes, void** d_ptr);
int amt_free_device(amt_ctx* ctx, void* d_ptr);
int amt_alloc_pinned(amt_ctx* ctx, size_t bytes, void** h_ptr);
int amt_free_pinned(amt_ctx* ctx, void* h_ptr);
int amt_copy_h2d(amt_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, void* stream);
/* Rectangle of a row-major host image -> the same rectangle of its device copy
 * (cudaMemcpy2DAsync): the sequence pipeline uploads only the pixel box that holds
 * georeferenced pixels (amt_stats.row_min_c .. col_max_c).                                 */
int amt_copy_h2d_2d(amt_ctx* ctx, void* d_dst, size_t d_pitch, const void* h_src, size_t h_pitch,
                    size_t width_bytes, size_t rows, void* stream);
int amt_copy_d2h(amt_ctx* ctx, void* h_dst, const void* d_src, size_t bytes, void* stream);
int amt_memset_device(amt_ctx* ctx, void* d_ptr, int value, size_t bytes, void* stream);
int amt_stream_synchronize(amt_ctx* ctx, void* stream);

/* ------------------------------------------------------------------- stage 1 + 2 ---- */
/* Fused pixel -> WCS(TAN[+SIP]) -> ray -> inflated-ellipsoid intersection -> geodetic
 * lat/lon, MLat/MLT, elevation, for all corners and all centres of one frame.
 * Replaces: coordinates/wcs.py:18-157, mapping/mapping.py:1474-1510,
 * coordinates/intersection.py:58-104, coordinates/transform.py:252-297,324-343,403-430,
 * mapping/astrometry.py:49-64,86-106,138-212, utils.py:28-46.                             */
int amt_georef(amt_ctx* ctx, const amt_frame* frame, const amt_georef_out* out,
               amt_stats* d_stats /* nullable: n_ill_conditioned is zeroed, then counted; other
                                      fields untouched */, void* stream);

/* Validity bitmaps from NaN-marked latitude planes (mappings that were not produced by
 * amt_georef: uploaded arrays, resampled grids).                                          */
int amt_valid_bits(amt_ctx* ctx, int32_t width, int32_t height, const double* d_lat_k,
                   const double* d_lat_c, uint32_t* d_valid_k, uint32_t* d_valid_c, void* stream);

/* Mask sanitisation, in place: planes get NaN where the reference would mask, the bitmaps
 * are updated.  Replaces mapping/mapping.py:1063-1125 (`_doSanitize`, afterMasking=False, no
 * image mask): corner masked if all of its (<=4) neighbouring centres are missing; centre
 * masked if any of its 4 corners is masked; corners once more.  The stencils run on the
 * bitmaps; only newly masked elements of the planes are touched.                          */
int amt_sanitize(amt_ctx* ctx, int32_t width, int32_t height, const amt_georef_out* planes,
                 void* stream);

/* Bounding-box reductions over the outline = valid corners with an invalid or out-of-array
 * 4-neighbour (min/max part of mapping/mapping.py:694-743) plus valid counts.  When `pre` is
 * non-NULL and pre->prerotate != AMT_PRE_NONE the outline coordinates are first rotated
 * exactly as resample.py:176-218 rotates the outline (only prerotate, altitude, wgs_a/b and
 * rot of `pre` are read).  pole_test != 0 additionally runs the per-pixel longitude winding
 * test (one pass over lon_k).  `d_stats` is a DEVICE amt_stats; n_ill_conditioned is left
 * untouched.                                                                              */
int amt_bbox_stats(amt_ctx* ctx, int32_t width, int32_t height, const double* d_lat_k,
                   const double* d_lon_k, const uint32_t* d_valid_k, const uint32_t* d_valid_c,
                   int32_t pole_test, const amt_grid* pre, amt_stats* d_stats, void* stream);

/* Apply a centre mask (mapping/mapping.py:845-864 maskedByElevation, :1171-1231 createMasked
 * followed by `_doSanitize(afterMasking=True)`), in place: a centre becomes NaN if
 * d_mask[i] != 0 (d_mask nullable) or if !(d_elev_c[i] >= min_elevation) (skipped when
 * min_elevation is NaN); then every corner whose (<=4) neighbouring centres are all missing
 * becomes NaN.                                                                             */
int amt_apply_center_mask(amt_ctx* ctx, int32_t width, int32_t height, const uint8_t* d_mask,
                          double min_elevation, const amt_georef_out* planes, void* stream);

/* rotatePole (coordinates/transform.py:301-322) or the 180-degree longitude wrap
 * (resample.py:213,218,276-277) applied in place to `n` (lat, lon) pairs in degrees; only
 * prerotate, altitude, wgs_a/b and rot of `pre` are read.                                  */
int amt_rotate_coords(amt_ctx* ctx, double* d_lat, double* d_lon, size_t n, const amt_grid* pre,
                      void* stream);

/* Altitude reprojection of a ground-based imager's calibrated corner coordinates
 * (mapping/themis.py:224-253 `reproject`): every (lat, lon) [deg] valid for emission height
 * `height_ref` [km] is turned into the viewing direction from the station (ECEF position
 * `station_ecef` [km], = geodetic2EcefZero of the station) and intersected with the WGS84
 * ellipsoid inflated by `height_new`; result in degrees, NaN for NaN input or a missed ray.   */
int amt_reproject(amt_ctx* ctx, const double* d_lat_ref, const double* d_lon_ref, size_t n,
                  const double station_ecef[3], double height_ref, double height_new, double wgs_a,
                  double wgs_b, double* d_lat_out, double* d_lon_out, void* stream);

/* Centre coordinates as the mean of the four surrounding corners (mapping/themis.py:425-426,
 * same summation order; mapping/astrometry.py:154-160): corner planes (height+1)x(width+1) ->
 * centre planes height x width.                                                              */
int amt_corner_means(amt_ctx* ctx, int32_t width, int32_t height, const double* d_lat_k,
                     const double* d_lon_k, double* d_lat_c, double* d_lon_c, void* stream);

/* maskedByPolygon (mapping/mapping.py:866-917) with the inside test of utils.py:58-74
 * (matplotlib `Path.contains_points` in the reference; here the crossing-number test with
 * half-open edges, x = latitude, y = longitude): d_center_mask[y*width+x] = 0 if the four corners
 * of pixel (x,y) are all defined (not NaN) and inside the ordered, unclosed polygon
 * d_polygon[n_polygon][2] = (lat, lon) in degrees (DEVICE memory), else 1.  `pre` (nullable)
 * applies the 180-degree wrap / pole rotation of amt_rotate_coords to the corner coordinates
 * first; the caller rotates the polygon the same way.  *d_n_inside (nullable, device) receives
 * the number of corners inside (0 => "the given mask would mask all pixels", :906-907).      */
int amt_polygon_center_mask(amt_ctx* ctx, int32_t width, int32_t height, const double* d_lat_k,
                            const double* d_lon_k, const double* d_polygon, int32_t n_polygon,
                            const amt_grid* pre, uint8_t* d_center_mask, uint64_t* d_n_inside,
                            void* stream);

/* Coordinates of the plate-carree target grid (resample.py:229-241): writes the 2-D corner
 * planes (ny+1)x(nx+1) and centre planes ny x nx of the resampled mapping, row 0 = north,
 * from the snapped node ranges lat: linspace(lat_hi, lat_lo, ny+2), lon: linspace(lon_lo,
 * lon_hi, nx+2) -- bit-identical to numpy.linspace / meshgrid.  Any output may be NULL.   */
int amt_plate_carree_coords(amt_ctx* ctx, int32_t nx, int32_t ny, double lat_hi, double lat_lo,
                            double lon_lo, double lon_hi, double* d_lat_k, double* d_lon_k,
                            double* d_lat_c, double* d_lon_c, void* stream);

/* Generic (non-astrometry) geodetic -> MLat/MLT route of mapping/mapping.py:540-550 +
 * coordinates/transform.py:156-178,432-459 for `n` points; lat/lon in degrees.           */
int amt_latlon_to_mlatmlt(amt_ctx* ctx, const double* d_lat, const double* d_lon, size_t n,
                          double altitude, double wgs_a, double wgs_b, const double m_geo_sm[9],
                          double* d_mlat, double* d_mlt, void* stream);

/* Solar-magnetic (lat, lon) in degrees -> geodetic (lat, lon) in degrees, in place:
 * coordinates/transform.py:461-485 `smToLatLon` (unit vector -> M_geo_sm^T -> Bowring), used by
 * `convertSMMappingToGeo` (mapping/mapping.py:1549-1559) after `resampleMLatMLT`.            */
int amt_sm_to_latlon(amt_ctx* ctx, double* d_lat, double* d_lon, size_t n, const double m_geo_sm[9],
                     double wgs_a, double wgs_b, void* stream);

/* ----------------------------------------------------------------------- stage 3 ---- */
/* Accumulators: `d_count` and `d_sums` are u64 planes of ny*nx cells, row 0 = NORTHERNMOST
 * latitude row (the reference's flipud, resample.py:349); `d_sums` holds `channels` planes;
 * `d_fsum` is one f64 plane for the float side channel (elevation) or NULL.
 * The caller zeroes them (amt_memset_device) -- or keeps accumulating several frames into
 * the same grid (mosaic).
 * Replaces resample.py:176-218 (pre-rotation), :315-338 + util/histogram.py:205-262.
 * `d_img`: height*width*channels interleaved (numpy HWC), dtype u8 or u16.
 * A centre pixel participates iff its latitude is not NaN (resample.py:316).
 * `d_near_edge`: optional u64 counter of samples within 1 ulp of a bin edge.              */
int amt_bin_accumulate(amt_ctx* ctx, const double* d_lat_c, const double* d_lon_c,
                       const double* d_side, const void* d_img, int32_t dtype, int32_t channels,
                       size_t n_pixels, const amt_grid* grid, uint64_t* d_count, uint64_t* d_sums,
                       double* d_fsum, uint64_t* d_near_edge, void* stream);

/* Per-sample core-bin indices (ix, iy), -1 for outliers / NaN: the index computation of
 * util/histogram.py:205-224 exposed for the bit-exactness tests.                          */
int amt_cell_indices(amt_ctx* ctx, const double* d_lat_c, const double* d_lon_c, size_t n_pixels,
                     const amt_grid* grid, int32_t* d_ix, int32_t* d_iy, void* stream);

/* sum / count -> mean; NaN (mask=1) where count == 0; integer images are rounded half-even
 * and cast back to `dtype` (resample.py:128-136,339-351).
 * out_img: ny*nx*channels interleaved; out_mask: ny*nx bytes; out_side: ny*nx doubles.   */
int amt_normalise(amt_ctx* ctx, const amt_grid* grid, int32_t dtype, int32_t channels,
                  const uint64_t* d_count, const uint64_t* d_sums, const double* d_fsum,
                  void* d_out_img, uint8_t* d_out_mask, double* d_out_side, void* stream);

/* Plane-free resampling, three calls:
 *   1. amt_georef with only d_valid_k / d_valid_c set: hit ballots of all corner and centre
 *      rays (direction + discriminant only), then amt_sanitize on the bitmaps;
 *   2. amt_bbox_stats_frame: outline min/max with the outline coordinates recomputed from the
 *      frame model for the few outline nodes (no planes) -> the host derives the grid;
 *   3. amt_georef_bin_fused: pixel -> ray -> intersection -> lat/lon (+elevation) -> cell ->
 *      accumulate for every centre whose bit is set; 3 B/pixel of image in, the grids out.
 * fast_center frames are not supported (their centres need the corner intersection points). */
int amt_bbox_stats_frame(amt_ctx* ctx, const amt_frame* frame, const uint32_t* d_valid_k,
                         const uint32_t* d_valid_c, const amt_grid* pre, amt_stats* d_stats,
                         void* stream);
/* The fused pass of the sequence pipeline: from the frame's FINAL validity bitmaps (hit test +
 * amt_sanitize, both on bitmaps; grid derived from amt_bbox_stats_frame) one kernel writes the
 * coordinate planes -- NaN exactly where the bitmaps say so, nothing is patched afterwards -- and
 * bins every defined centre into the target grid while its coordinates are still in registers.
 *   out  != NULL: d_lat_k, d_lon_k, d_lat_c, d_lon_c, d_elev_c are required, the four MLat/MLT
 *                 planes are written iff all four are given (the bitmap members are ignored);
 *   out  == NULL: plane-free resampling (== amt_georef_bin_fused);
 *   grid != NULL: binning into d_count / d_sums / d_fsum as amt_bin_accumulate does (same cells,
 *                 same sums: the kernel bins the very values it stores);
 *   grid == NULL: planes only.
 * Replaces, for WCS frames with fast_center == 0: amt_georef + the plane part of amt_sanitize +
 * amt_bin_accumulate, i.e. the reference lines cited there.                                     */
int amt_georef_fused(amt_ctx* ctx, const amt_frame* frame, const amt_georef_out* out,
                     const uint32_t* d_valid_k, const uint32_t* d_valid_c, const void* d_img,
                     int32_t dtype, int32_t channels, const amt_grid* grid, uint64_t* d_count,
                     uint64_t* d_sums, double* d_fsum, void* stream);
int amt_georef_bin_fused(amt_ctx* ctx, const amt_frame* frame, const uint32_t* d_valid_c,
                         const void* d_img, int32_t dtype, int32_t channels, const amt_grid* grid,
                         uint64_t* d_count, uint64_t* d_sums, double* d_fsum, void* stream);

/* The FITS-SIP forward distortion of `frame` for n pixel offsets (u, v) = (x - CRPIX1, y - CRPIX2)
 * (1-based x, y):  u' = u + sum_{p+q<=A_ORDER} A_p_q u^p v^q,  v' = v + sum B_p_q u^p v^q  -- the device
 * function of the georeference kernels, exposed for known-answer tests (the reference reaches SIP only
 * through astropy.wcs, coordinates/wcs.py:54-56; tests pin this against exact rational arithmetic).  */
int amt_sip_distort(amt_ctx* ctx, const amt_frame* frame, const double* d_u, const double* d_v, size_t n,
                    double* d_u_out, double* d_v_out, void* stream);

/* --------------------------------------------------------- host arithmetic of the grid ---- */
/* Bit-identical C ports of the scalar host code around resample() (no CUDA call; usable without a
 * GPU): they let the sequence engine derive a frame's target grid without returning to Python.    */
typedef struct amt_grid_info {
    int32_t n_lat, n_lon;                        /* nodes of the snapped box (resample.py:297-298)  */
    double lat_min_in_grid, lat_max_in_grid;     /* fixedGrid, resample.py:293-296                  */
    double lon_min_in_grid, lon_max_in_grid;
    double lat_step, lon_step;                   /* np.linspace(..., retstep=True), :222-227        */
    double lat_px_per_deg, lon_px_per_deg;
} amt_grid_info;

/* resample.py:36-61 plateCarreeResolution (arc length on the auxiliary sphere by Vincenty's inverse
 * iteration in place of geographiclib's a12).                                                      */
int amt_plate_carree_resolution(double lat_south, double lon_west, double lat_north, double lon_east,
                                double arcsec_per_px, double* lat_px_per_deg, double* lon_px_per_deg);
/* resample.py:281-299 fixedGrid + :220-241,330-335 and util/histogram.py:185-186,215-219: the
 * amt_grid of a bounding box without pre-rotation.  *fallback = 1: the decimal of histogram.py:218
 * needs the materialised bin edges (step within 1e-9 of a power of ten) -- derive it in Python.     */
int amt_target_grid(double lat_px_per_deg, double lon_px_per_deg, double lat_min, double lat_max,
                    double lon_min, double lon_max, amt_grid* grid, amt_grid_info* info, int32_t* fallback);
/* amt_grid.side_scale for n_samples values of magnitude < 128 (elevations).                        */
double amt_side_scale(uint64_t n_samples);
/* Pixels of a WCS frame that show the geographic north / south pole at the mapping altitude (inverse
 * WCS projection of the pole point); in_frame[i] == 0 when pole i is not in the field of view.
 * Replaces the outline / azimuth-sum test of mapping/mapping.py:705-718, geodesic.py:183-202: the
 * mapping encloses the pole iff that pixel is defined.                                             */
int amt_pole_pixels(const amt_frame* frame, int32_t ix[2], int32_t iy[2], int32_t in_frame[2]);
/* Rigorous bounds, in pixels, of the FITS-SIP displacement (u, v) -> (u + f(u,v), v + g(u,v)) over the pixel
 * array of a frame (coordinates/wcs.py:54-56: the reference hands -SIP headers to astropy/wcslib):
 * sum |A_pq| U^p V^q with U, V the largest |u|, |v| of the frame.  Host arithmetic, no device needed; 0, 0 for
 * a header without SIP.  The hit-bitmap solver of TAN-SIP frames (amt_georef with only the bitmaps requested)
 * inflates the box of every bitmap word by these bounds.                                                 */
int amt_sip_displacement_bound(const amt_frame* frame, double* dx, double* dy);

/* ----------------------------------------------------------------- sequence engine ---- */
/* The pipelined composition of getMappingSequence (mapping/spacecraft.py:308-332) and
 * ResampleProvider (resample.py:370-394) for WCS frames with fast_center == 0: two calls per frame.
 *   amt_seq_stage_a   hit bitmaps (limb solver / per-pixel hit test) -> amt_sanitize on the bitmaps
 *                     -> amt_bbox_stats_frame -> asynchronous copy of the statistics to pinned memory;
 *   amt_seq_wait_stats blocks until those statistics are on the host; the caller derives the
 *                     bounding box and the target grid (reference arithmetic);
 *   amt_seq_stage_b   upload of the pixel box of the host image that holds defined pixels -> zeroed
 *                     accumulators -> amt_georef_fused (planes + binning) -> amt_normalise ->
 *                     asynchronous copy of the results to pinned memory;
 *   amt_seq_wait_result blocks until the results of the slot are complete.
 * The engine orders four caller-provided streams with events (long kernels on `main_stream`, the
 * microsecond kernels of stage A on `aux_stream`, copies on their own streams) and owns NO memory:
 * ring-slot buffers come through amt_seq_set_slot, per-frame buffers through amt_seq_job.  A slot
 * may be re-submitted as soon as the caller no longer needs its planes / bitmaps; the engine makes
 * the new frame's kernels wait for the slot's previous fused kernel.  One host thread per engine;
 * NVTX ranges mark both stages.                                                                  */
typedef struct amt_seq amt_seq;

typedef struct amt_seq_slot {
    amt_georef_out planes;   /* d_valid_k / d_valid_c required; coordinate planes: all of lat/lon corner +
                                centre + elevation (and optionally the four MLat/MLT planes), or none
                                (plane-free resampling)                                               */
    amt_stats* d_stats;      /* device statistics block                                              */
    amt_stats* h_stats;      /* pinned host copy                                                     */
    void* d_img;             /* device image buffer height*width*channels (nullable when every job
                                brings its own d_img)                                                */
} amt_seq_slot;

typedef struct amt_seq_job {
    const amt_grid* grid;
    const void* h_img;       /* host image (pinned for an asynchronous copy), complete frame; NULL: the
                                image is already on the device                                       */
    const void* d_img;       /* device image to read instead of the slot's buffer (nullable)        */
    int32_t row0, row1, col0, col1;  /* pixel box to upload, inclusive (amt_stats.row_min_c ...)     */
    uint64_t* d_acc;         /* (2 + channels) * nx*ny words: count | sums[channels] | side sums      */
    void* d_out;             /* image | mask | side, see amt_seq_output_layout                       */
    void* h_out;             /* pinned host copy of d_out (nullable: results stay on the device)     */
    size_t out_bytes;        /* capacity of d_out (and h_out)                                        */
} amt_seq_job;

/* Byte offsets of the mask and of the side (elevation) plane inside the output buffer of a grid,
 * and its total size: image nx*ny*channels at 0, mask nx*ny bytes and side nx*ny doubles at the
 * next 64-byte boundaries.  Host only (no CUDA call).                                             */
int amt_seq_output_layout(int32_t nx, int32_t ny, int32_t channels, int32_t dtype, size_t* off_mask,
                          size_t* off_side, size_t* total);
int amt_seq_create(amt_ctx* ctx, int32_t width, int32_t height, int32_t channels, int32_t dtype,
                   int32_t n_slots, void* main_stream, void* aux_stream, void* copy_stream,
                   void* out_stream, amt_seq** out);
int amt_seq_destroy(amt_seq* seq);
int amt_seq_set_slot(amt_seq* seq, int32_t slot, const amt_seq_slot* buffers);
int amt_seq_stage_a(amt_seq* seq, int32_t slot, const amt_frame* frame);
int amt_seq_wait_stats(amt_seq* seq, int32_t slot, amt_stats* out /* nullable */);
/* Statistics -> pole flags -> bounding box (mapping/mapping.py:694-743) -> target grid, in C.  One of
 * arcsec_per_px (> 0) or the pxPerDeg pair selects the resolution (resample.py:95-117).             */
enum { AMT_PLAN_OK = 0, AMT_PLAN_HOST = 1, AMT_PLAN_EMPTY = 2 };
int amt_seq_plan(amt_seq* seq, int32_t slot, double arcsec_per_px, double lat_px_per_deg,
                 double lon_px_per_deg, amt_stats* stats_out, amt_grid* grid, amt_grid_info* info,
                 int32_t* outcome);
int amt_seq_stage_b(amt_seq* seq, int32_t slot, const amt_seq_job* job);
int amt_seq_wait_result(amt_seq* seq, int32_t slot);
/* Image bytes copied host -> device so far (bench.py h2d_bytes_per_step).                         */
int amt_seq_h2d_bytes(const amt_seq* seq, uint64_t* bytes);
/* Tracing (the reference logs wall-clock durations per stage, e.g. resample.py:139-141):
 * with AMT_SEQ_TRACE=1 in the environment when the engine is created, every stream event of the engine
 * carries a time stamp; out_ms7 = device times in ms since engine creation of {stage A start, stage A
 * end, upload start, upload end, fused kernel start, fused kernel end, results complete} of the slot's
 * last frame.  Call after amt_seq_wait_result, before the slot is reused. */
int amt_seq_trace(amt_seq* seq, int32_t slot, double* out_ms7);

#ifdef __cplusplus
}
#endif
#endif /* AUROMAT_B200_H */
