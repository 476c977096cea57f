// Sequence engine: the per-frame launch sequences of the pipelined getMapping + resample path
// (reference: getMappingSequence, mapping/spacecraft.py:308-332, composed with ResampleProvider,
// resample.py:370-394) as TWO C calls per frame instead of ~25 Python-level operations.
//
//   stage A (frame header only):   hit bitmaps (limb solver for pure TAN frames, per-pixel hit test
//       otherwise) -> sanitisation stencils on the bitmaps -> outline statistics from the frame
//       model -> asynchronous copy of the 104-byte statistics block to pinned host memory.
//       All on the auxiliary high-priority stream: microsecond kernels that overlap the long
//       kernel of the preceding frames.
//   host:                          bounding box -> target grid (reference arithmetic, Python).
//   stage B (grid + image):        zeroed accumulators (auxiliary stream) + upload of the pixel box that holds defined
//       pixels (copy stream: DMA only) -> ONE fused kernel: coordinate planes + binning (main stream, nothing
//       else ever runs there: the long kernels follow each other back to back) -> normalise ->
//       results to pinned host memory (output stream).
//
// The engine owns streams' ORDER (events), never memory: every buffer is provided by the caller
// (ring slots: planes, bitmaps, statistics block, device image; per frame: accumulators, outputs).
// One engine per amt_ctx, driven by one host thread.  Included at the end of amt.cu.
#pragma once
#include <nvtx3/nvToolsExt.h>

struct SeqSlot {
    amt_seq_slot buf;
    bool is_set;
    amt_frame frame;
    cudaEvent_t ev_a, ev_up, ev_b, ev_out, ev_zero, ev_fork;
    cudaEvent_t tr_a0, tr_up0, tr_b0;      // trace mode only: starts of stage A, of the upload, of the fused kernel
    bool a_rec, b_rec, out_rec;
};

// Development builds only (-DAMT_DEV_SKIP): AMT_SEQ_SKIP is a bit mask of launches the engine leaves out once every
// slot has seen a frame (1 statistics, 4 hit bitmaps + sanitise, 8 memset, 16 normalise, 32 download).  With a
// sequence of IDENTICAL frames the stale buffers hold the right values, so the timeline shows what each launch
// costs the fused kernel it overlaps (scripts/seq_trace.py).  Not compiled into the product library.
#ifdef AMT_DEV_SKIP
static int seq_skip_mask() { static int m = getenv("AMT_SEQ_SKIP") ? atoi(getenv("AMT_SEQ_SKIP")) : 0; return m; }
#define SEQ_SKIP(seq, bit) ((seq_skip_mask() & (bit)) && (seq)->frames_a > 2 * (seq)->slots.size())
#else
#define SEQ_SKIP(seq, bit) false
#endif

struct amt_seq {
    amt_ctx* ctx;
    size_t frames_a = 0;
    int W, H, C, dtype;
    cudaStream_t s_main, s_aux, s_copy, s_out;
    // Second stream for the long kernels (same priority as s_main), used by every other frame: the fused
    // kernels of consecutive frames are independent (own ring slot, own accumulators), and on two streams the
    // next one starts dispatching as soon as the current one has dispatched its last CTA -- the head of frame
    // i+1 fills the SMs that the tail of frame i leaves idle.  On one stream it waited for the last CTA to finish.
    cudaStream_t s_main2;
    size_t frames_b = 0;
    std::vector<SeqSlot> slots;
    unsigned long long h2d_bytes;
    bool trace;                            // AMT_SEQ_TRACE: timing events, see amt_seq_trace
    cudaEvent_t ev_base;
};

static size_t seq_align64(size_t n) { return (n + 63) / 64 * 64; }

extern "C" int amt_seq_output_layout(int32_t nx, int32_t ny, int32_t channels, int32_t dtype, size_t* off_mask,
                                     size_t* off_side, size_t* total) {
    CHECK_ARG(nx > 0 && ny > 0 && channels >= 1 && channels <= 4, "amt_seq_output_layout: bad arguments");
    CHECK_ARG(dtype == AMT_U8 || dtype == AMT_U16, "amt_seq_output_layout: dtype must be uint8 or uint16");
    const size_t cells = (size_t)nx * ny, item = dtype == AMT_U8 ? 1 : 2;
    const size_t om = seq_align64(cells * channels * item), os = seq_align64(om + cells);
    if (off_mask) *off_mask = om;
    if (off_side) *off_side = os;
    if (total) *total = os + cells * 8;
    return AMT_OK;
}

extern "C" int amt_seq_create(amt_ctx* ctx, int32_t width, int32_t height, int32_t channels, int32_t dtype,
                              int32_t n_slots, void* main_stream, void* aux_stream, void* copy_stream,
                              void* out_stream, amt_seq** out) {
    ENTER(ctx);
    CHECK_ARG(out, "amt_seq_create: out is NULL");
    CHECK_ARG(width > 0 && height > 0 && channels >= 1 && channels <= 4, "amt_seq_create: bad frame shape");
    CHECK_ARG(dtype == AMT_U8 || dtype == AMT_U16, "amt_seq_create: image dtype must be uint8 or uint16");
    CHECK_ARG(n_slots >= 1 && n_slots <= 64, "amt_seq_create: 1..64 slots");
    amt_seq* s = new (std::nothrow) amt_seq();
    if (!s) return set_err(AMT_ERR_CUDA, "out of host memory");
    s->ctx = ctx;
    s->W = width; s->H = height; s->C = channels; s->dtype = dtype;
    s->s_main = (cudaStream_t)main_stream; s->s_aux = (cudaStream_t)aux_stream;
    s->s_copy = (cudaStream_t)copy_stream; s->s_out = (cudaStream_t)out_stream;
    s->h2d_bytes = 0;
    s->s_main2 = nullptr;
    if (!getenv("AMT_SEQ_ONE_MAIN_STREAM")) {
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&s->s_main2, cudaStreamNonBlocking, lo));
    }
    const char* tr = getenv("AMT_SEQ_TRACE");
    s->trace = tr && tr[0] && tr[0] != '0';
    s->ev_base = nullptr;
    if (s->trace) {
        CUDA_TRY(cudaEventCreate(&s->ev_base));
        CUDA_TRY(cudaEventRecord(s->ev_base, s->s_main));
    }
    s->slots.resize(n_slots);
    for (auto& sl : s->slots) {
        memset(&sl.buf, 0, sizeof sl.buf);
        sl.is_set = sl.a_rec = sl.b_rec = sl.out_rec = false;
        sl.tr_a0 = sl.tr_up0 = sl.tr_b0 = nullptr;
        cudaEvent_t* evs[9] = {&sl.ev_a, &sl.ev_up, &sl.ev_b, &sl.ev_out, &sl.ev_zero, &sl.ev_fork, &sl.tr_a0, &sl.tr_up0,
                               &sl.tr_b0};
        for (int k = 0; k < (s->trace ? 9 : 6); ++k) {
            cudaEvent_t* e = evs[k];
            cudaError_t err = cudaEventCreateWithFlags(e, s->trace ? cudaEventDefault : cudaEventDisableTiming);
            if (err != cudaSuccess) {
                delete s;
                return set_err(AMT_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(err));
            }
        }
    }
    *out = s;
    return AMT_OK;
}

extern "C" int amt_seq_destroy(amt_seq* seq) {
    if (!seq) return AMT_OK;
    cudaSetDevice(seq->ctx->device);
    for (auto& sl : seq->slots) {
        cudaEventDestroy(sl.ev_a); cudaEventDestroy(sl.ev_up); cudaEventDestroy(sl.ev_b); cudaEventDestroy(sl.ev_out);
        cudaEventDestroy(sl.ev_zero);
        cudaEventDestroy(sl.ev_fork);
        if (seq->trace) { cudaEventDestroy(sl.tr_a0); cudaEventDestroy(sl.tr_up0); cudaEventDestroy(sl.tr_b0); }
    }
    if (seq->ev_base) cudaEventDestroy(seq->ev_base);
    if (seq->s_main2) cudaStreamDestroy(seq->s_main2);
    delete seq;
    return AMT_OK;
}

#define SEQ_SLOT(seq, slot)                                                                     \
    CHECK_ARG((seq) != nullptr, "sequence engine is NULL");                                     \
    CHECK_ARG((slot) >= 0 && (slot) < (int)(seq)->slots.size(), "slot index out of range");     \
    SeqSlot& sl = (seq)->slots[slot]

extern "C" int amt_seq_set_slot(amt_seq* seq, int32_t slot, const amt_seq_slot* buffers) {
    SEQ_SLOT(seq, slot);
    CHECK_ARG(buffers, "amt_seq_set_slot: buffers is NULL");
    CHECK_ARG(buffers->planes.d_valid_k && buffers->planes.d_valid_c, "amt_seq_set_slot: the validity bitmaps are required");
    CHECK_ARG(buffers->d_stats && buffers->h_stats, "amt_seq_set_slot: device and pinned statistics blocks are required");
    sl.buf = *buffers;
    sl.is_set = true;
    return AMT_OK;
}

extern "C" int amt_seq_stage_a(amt_seq* seq, int32_t slot, const amt_frame* frame) {
    SEQ_SLOT(seq, slot);
    amt_ctx* ctx = seq->ctx;
    ENTER(ctx);
    CHECK_ARG(frame && sl.is_set, "amt_seq_stage_a: frame is NULL or the slot has no buffers");
    CHECK_ARG(frame->width == seq->W && frame->height == seq->H, "amt_seq_stage_a: frame shape differs from the engine's");
    if (frame->model != AMT_MODEL_WCS || frame->fast_center)
        return set_err(AMT_ERR_UNSUPPORTED, "amt_seq_stage_a: WCS frames with fast_center == 0 only");
    nvtxRangePushA("amt_seq stage A: hit bitmaps, sanitise, outline statistics");
    sl.frame = *frame;
    cudaStream_t st = seq->s_aux;
    // the bitmaps of this slot are read by the fused kernel of the slot's previous frame
    if (sl.b_rec) CUDA_TRY(cudaStreamWaitEvent(st, sl.ev_b, 0));
    if (seq->trace) CUDA_TRY(cudaEventRecord(sl.tr_a0, st));
    amt_georef_out bits;
    memset(&bits, 0, sizeof bits);
    bits.d_valid_k = sl.buf.planes.d_valid_k;
    bits.d_valid_c = sl.buf.planes.d_valid_c;
    ++seq->frames_a;
    int rc = SEQ_SKIP(seq, 4) ? 0 : hit_bits_sanitized(ctx, frame, bits.d_valid_k, bits.d_valid_c, sl.buf.d_stats, st);
    if (!rc && !SEQ_SKIP(seq, 1))
        rc = amt_bbox_stats_frame(ctx, frame, bits.d_valid_k, bits.d_valid_c, nullptr, sl.buf.d_stats, st);
    if (rc) { nvtxRangePop(); return rc; }
    CUDA_TRY(cudaMemcpyAsync(sl.buf.h_stats, sl.buf.d_stats, sizeof(amt_stats), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(sl.ev_a, st));
    sl.a_rec = true;
    nvtxRangePop();
    return AMT_OK;
}

extern "C" int amt_seq_wait_stats(amt_seq* seq, int32_t slot, amt_stats* out) {
    SEQ_SLOT(seq, slot);
    CHECK_ARG(sl.a_rec, "amt_seq_wait_stats: stage A has not been submitted for this slot");
    CUDA_TRY(cudaEventSynchronize(sl.ev_a));
    if (out) *out = *sl.buf.h_stats;
    return AMT_OK;
}

// Statistics of the slot's frame -> bounding box -> target grid, entirely in C (amt_host.cuh): what
// BaseMapping.boundingBox + resample.deriveGrid do in Python.  *outcome: AMT_PLAN_OK (grid, info valid),
// AMT_PLAN_HOST (the frame encloses a pole or straddles the date line, or the bin-edge bookkeeping needs
// the materialised edges: the caller derives the grid with the Python path -- rare), AMT_PLAN_EMPTY (no
// defined corner at all).  `stats_out` always receives the statistics including the pole flags.
extern "C" int amt_seq_plan(amt_seq* seq, int32_t slot, double arcsec_per_px, double lat_px_per_deg,
                            double lon_px_per_deg, amt_stats* stats_out, amt_grid* grid, amt_grid_info* info,
                            int32_t* outcome) {
    SEQ_SLOT(seq, slot);
    CHECK_ARG(stats_out && grid && info && outcome, "amt_seq_plan: NULL argument");
    CHECK_ARG(sl.a_rec, "amt_seq_plan: stage A has not been submitted for this slot");
    CUDA_TRY(cudaEventSynchronize(sl.ev_a));
    amt_stats st = *sl.buf.h_stats;
    // pole containment: is the pixel that shows a pole defined?  (one bitmap word, only when a pole
    // projects into the pixel array at all)
    int32_t ix[2], iy[2], in_frame[2];
    int rc = amt_pole_pixels(&sl.frame, ix, iy, in_frame);
    if (rc) return rc;
    st.pole_flags = 0;
    for (int i = 0; i < 2; ++i) {
        if (!in_frame[i]) continue;
        CUDA_TRY(cudaSetDevice(seq->ctx->device));
        const int wc = (seq->W + 31) / 32;
        uint32_t word = 0;
        CUDA_TRY(cudaMemcpyAsync(&word, sl.buf.planes.d_valid_c + (size_t)iy[i] * wc + ix[i] / 32, 4,
                                 cudaMemcpyDeviceToHost, seq->s_aux));
        CUDA_TRY(cudaStreamSynchronize(seq->s_aux));
        if ((word >> (ix[i] % 32)) & 1u) st.pole_flags |= (uint64_t)(1u << i);
    }
    *stats_out = st;
    if (st.n_boundary_corners == 0) { *outcome = AMT_PLAN_EMPTY; return AMT_OK; }
    // BaseMapping.boundingBox (reference mapping/mapping.py:694-743)
    double lat_s = st.lat_min, lat_n = st.lat_max, lon_w = st.lon_min, lon_e = st.lon_max;
    if (st.pole_flags || st.lon_max - st.lon_min > 180) { *outcome = AMT_PLAN_HOST; return AMT_OK; }
    double px_lat = lat_px_per_deg, px_lon = lon_px_per_deg;
    if (arcsec_per_px > 0) {
        rc = amt_plate_carree_resolution(lat_s, lon_w, lat_n, lon_e, arcsec_per_px, &px_lat, &px_lon);
        if (rc) return rc;
    }
    int32_t fb = 0;
    rc = amt_target_grid(px_lat, px_lon, lat_s, lat_n, lon_w, lon_e, grid, info, &fb);
    if (rc) return rc;
    grid->altitude = 0.0;
    grid->side_scale = amt_side_scale((uint64_t)seq->W * (uint64_t)seq->H);
    *outcome = fb ? AMT_PLAN_HOST : AMT_PLAN_OK;
    return AMT_OK;
}

extern "C" int amt_seq_stage_b(amt_seq* seq, int32_t slot, const amt_seq_job* job) {
    SEQ_SLOT(seq, slot);
    amt_ctx* ctx = seq->ctx;
    ENTER(ctx);
    CHECK_ARG(job && job->grid && job->d_acc && job->d_out, "amt_seq_stage_b: NULL argument");
    CHECK_ARG(sl.a_rec, "amt_seq_stage_b: stage A has not been submitted for this slot");
    const amt_grid* g = job->grid;
    size_t om, os, total;
    int rc = amt_seq_output_layout(g->nx, g->ny, seq->C, seq->dtype, &om, &os, &total);
    if (rc) return rc;
    CHECK_ARG(job->out_bytes >= total, "amt_seq_stage_b: output buffer too small for the grid");
    nvtxRangePushA("amt_seq stage B: upload, fused georeference + binning, normalise, download");
    const size_t cells = (size_t)g->nx * g->ny;
    const size_t item = seq->dtype == AMT_U8 ? 1 : 2, px = item * seq->C, row_bytes = px * seq->W;
    const void* d_img = job->d_img ? job->d_img : sl.buf.d_img;
    if (!d_img) { nvtxRangePop(); return set_err(AMT_ERR_INVALID_ARGUMENT, "amt_seq_stage_b: no device image buffer"); }
    // zeroed accumulators of this frame: a memset is a KERNEL, and a kernel of an ordinary-priority stream
    // gets its first CTA only when the fused kernel of the preceding frame has dispatched all of its own
    // (47 k CTAs): on the copy stream it held back the upload behind it until that kernel's tail, i.e.
    // upload and kernel ran one after the other (0.72 ms/frame instead of max(0.41, 0.31); AMT_SEQ_TRACE
    // timeline, profiles/r02_seq_trace.txt).  The auxiliary stream has high priority: its CTAs go first.
    if (!SEQ_SKIP(seq, 8)) CUDA_TRY(cudaMemsetAsync(job->d_acc, 0, (size_t)(2 + seq->C) * cells * 8, seq->s_aux));
    CUDA_TRY(cudaEventRecord(sl.ev_zero, seq->s_aux));
    // copy stream: nothing but DMA -- the pixel box of the host image
    if (seq->trace) CUDA_TRY(cudaEventRecord(sl.tr_up0, seq->s_copy));
    if (job->h_img && !job->d_img) {
        // the slot's image buffer is read by the fused kernel of the slot's previous frame
        if (sl.b_rec) CUDA_TRY(cudaStreamWaitEvent(seq->s_copy, sl.ev_b, 0));
        const int r0 = job->row0 < 0 ? 0 : job->row0, r1 = job->row1 >= seq->H ? seq->H - 1 : job->row1;
        const int c0 = job->col0 < 0 ? 0 : job->col0, c1 = job->col1 >= seq->W ? seq->W - 1 : job->col1;
        if (r1 >= r0 && c1 >= c0) {
            const size_t nrows = (size_t)(r1 - r0 + 1), ncols = (size_t)(c1 - c0 + 1);
            unsigned char* dst = (unsigned char*)sl.buf.d_img;
            const unsigned char* src = (const unsigned char*)job->h_img;
            if (ncols * 10 < (size_t)seq->W * 9) {
                // the defined pixels sit in a column band (limb roughly vertical): copy the box
                const size_t off = (size_t)r0 * row_bytes + (size_t)c0 * px;
                CUDA_TRY(cudaMemcpy2DAsync(dst + off, row_bytes, src + off, row_bytes, ncols * px, nrows,
                                           cudaMemcpyHostToDevice, seq->s_copy));
                seq->h2d_bytes += ncols * px * nrows;
            } else {
                const size_t off = (size_t)r0 * row_bytes;
                CUDA_TRY(cudaMemcpyAsync(dst + off, src + off, nrows * row_bytes, cudaMemcpyHostToDevice, seq->s_copy));
                seq->h2d_bytes += nrows * row_bytes;
            }
        }
    }
    CUDA_TRY(cudaEventRecord(sl.ev_up, seq->s_copy));
    // main stream: nothing but the long fused kernels, back to back from frame to frame
    cudaStream_t st = (seq->s_main2 && (seq->frames_b++ & 1)) ? seq->s_main2 : seq->s_main;
    if (st != seq->s_main) {
        // whatever the caller has enqueued on its stream so far (e.g. the kernel that produced a device image)
        CUDA_TRY(cudaEventRecord(sl.ev_fork, seq->s_main));
        CUDA_TRY(cudaStreamWaitEvent(st, sl.ev_fork, 0));
    }
    CUDA_TRY(cudaStreamWaitEvent(st, sl.ev_a, 0));             // final bitmaps of this frame (auxiliary stream)
    CUDA_TRY(cudaStreamWaitEvent(st, sl.ev_up, 0));            // image on the device
    CUDA_TRY(cudaStreamWaitEvent(st, sl.ev_zero, 0));          // accumulators zeroed
    uint64_t* count = job->d_acc;
    uint64_t* sums = count + cells;
    double* fsum = (double*)(count + (size_t)(1 + seq->C) * cells);
    const amt_georef_out* planes = sl.buf.planes.d_lat_k ? &sl.buf.planes : nullptr;
    if (seq->trace) CUDA_TRY(cudaEventRecord(sl.tr_b0, st));
    rc = georef_fused(ctx, &sl.frame, planes, sl.buf.planes.d_valid_k, sl.buf.planes.d_valid_c, d_img, seq->dtype,
                      seq->C, g, count, sums, fsum, st);
    if (rc) { nvtxRangePop(); return rc; }
    CUDA_TRY(cudaEventRecord(sl.ev_b, st));
    sl.b_rec = true;
    // output stream: normalise, then the results leave for pinned host memory -- neither delays the
    // next frame's fused kernel
    CUDA_TRY(cudaStreamWaitEvent(seq->s_out, sl.ev_b, 0));
    unsigned char* o = (unsigned char*)job->d_out;
    if (!SEQ_SKIP(seq, 16))
        rc = amt_normalise(ctx, g, seq->dtype, seq->C, count, sums, fsum, o, o + om, (double*)(o + os), seq->s_out);
    if (rc) { nvtxRangePop(); return rc; }
    if (job->h_out && !SEQ_SKIP(seq, 32)) CUDA_TRY(cudaMemcpyAsync(job->h_out, job->d_out, total, cudaMemcpyDeviceToHost, seq->s_out));
    CUDA_TRY(cudaEventRecord(sl.ev_out, seq->s_out));
    sl.out_rec = true;
    nvtxRangePop();
    return AMT_OK;
}

extern "C" int amt_seq_wait_result(amt_seq* seq, int32_t slot) {
    SEQ_SLOT(seq, slot);
    CHECK_ARG(sl.b_rec, "amt_seq_wait_result: stage B has not been submitted for this slot");
    // results are produced on the output stream: a host-level wait makes them visible to every stream
    CUDA_TRY(cudaEventSynchronize(sl.out_rec ? sl.ev_out : sl.ev_b));
    return AMT_OK;
}

// Trace mode (the engine was created with AMT_SEQ_TRACE=1 in the environment): device timeline of the
// slot's last frame in ms since the engine was created: {stage A start, stage A end (statistics on the
// host), upload start, upload end, fused kernel start, fused kernel end, results complete}.  Call after
// amt_seq_wait_result and before the slot is reused.
extern "C" int amt_seq_trace(amt_seq* seq, int32_t slot, double* out_ms7) {
    SEQ_SLOT(seq, slot);
    CHECK_ARG(out_ms7, "amt_seq_trace: NULL argument");
    if (!seq->trace) return set_err(AMT_ERR_UNSUPPORTED, "amt_seq_trace: engine created without AMT_SEQ_TRACE=1");
    CHECK_ARG(sl.a_rec && sl.b_rec && sl.out_rec, "amt_seq_trace: the slot has no complete frame");
    CUDA_TRY(cudaEventSynchronize(sl.ev_out));
    cudaEvent_t evs[7] = {sl.tr_a0, sl.ev_a, sl.tr_up0, sl.ev_up, sl.tr_b0, sl.ev_b, sl.ev_out};
    for (int k = 0; k < 7; ++k) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, seq->ev_base, evs[k]));
        out_ms7[k] = ms;
    }
    return AMT_OK;
}

extern "C" int amt_seq_h2d_bytes(const amt_seq* seq, uint64_t* bytes) {
    CHECK_ARG(seq && bytes, "amt_seq_h2d_bytes: NULL argument");
    *bytes = seq->h2d_bytes;
    return AMT_OK;
}
