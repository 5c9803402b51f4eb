// pcfusion.cu -- context management and the C ABI of libpcfusion.so (see include/pcfusion.h).
// The host side here replaces the reference's L2 pipeline (three std::threads + two mutex-protected deques,
// node.cpp:130-143,218-325) with pinned staging + two CUDA streams, and OccupancyGrid's containers
// (OG.hpp:99-136) with the flat HBM arrays described at the top of pcf_kernels.cuh.
// There is deliberately no CPU fallback: without a CUDA device pcf_create fails with PCF_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pcfusion.h"
#include "pcf_kernels.cuh"
#include "pcf_stager.hpp"

using namespace pcf;

namespace {

thread_local std::string g_create_error;

struct DevBuf {   // grow-only device allocation
    void* p = nullptr;
    size_t cap = 0;
};

constexpr int kRing = 4;           // staging slots for host frames
constexpr int kTickets = 16;       // upload-completion events kept (newer uploads imply older ones)
constexpr uint32_t kMaxChunks = 1u << 24;   // slot index must fit 32 bits: 2^24 chunks * 256

}  // namespace

struct pcf_ctx {
    pcf_config cfg{};
    GridParams g{};
    int device = 0;
    int sm_count = 148;
    int ctas_per_sm = kBulkMinBlocks;     // persistent CTAs per SM of the bulk kernel
    bool trace = false;                   // PCF_TRACE=1
    bool use_bulk = true;                 // PCF_INGEST=generic forces the plain-load kernel (A/B measurements)
    bool ingest_bits = true;              // PCF_INGEST_BITS=0: the ingest kernels leave the occupancy bitmap alone and it is rebuilt by a grid sweep (A/B)
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    std::string err;
    bool started = false;

    // persistent grid state
    uint32_t* first_frame = nullptr;
    uint32_t* holder = nullptr;           // lazily allocated
    uint32_t* nrm_bits = nullptr;
    uint32_t* occ_bits = nullptr;
    uint32_t* occ_rank = nullptr;
    uint64_t n_words = 0;
    bool occ_dirty = true;                // occ_rank / n_vox are stale (the bitmap itself is kept current by the ingest kernels)
    bool occ_from_grid = false;           // a caller reduced the dense grid externally (pcf_grid_buffer): rebuild the bitmap from it
    bool log_installed = false;           // the log came from pcf_install_records / pcf_log_replace: chunk_frame no longer describes it
    uint32_t n_vox = 0;                   // occupied cells at the last bitmap build
    float4* vp_table = nullptr;
    // point log
    float4* log = nullptr;
    uint32_t* chunk_count = nullptr;
    uint32_t* chunk_frame = nullptr;      // frame_idx of every log chunk (the exchange ships it with the records)
    uint32_t cap_chunks = 0, n_chunks = 0;
    uint32_t merged_chunks = 0;           // interleaved schedules across ranks: log chunks [0, merged_chunks) hold the records of ALL
                                          // ranks (installed by pcf_round_install), [merged_chunks, n_chunks) this rank's current round
    // normals (append-only records)
    DevBuf n_cell, n_nrm, n_mark;
    DevBuf upd_cell, upd_nrm;             // normal records of the pass being computed (pcf_update_local), not yet committed
    uint32_t upd_n = 0, upd_mark = 0;
    bool upd_open = false;
    uint32_t n_normals = 0;
    std::vector<uint32_t> marks;                       // log slot cursor at each update pass
    std::vector<std::pair<uint32_t, uint32_t>> pending_holder;   // [begin,end) normal records per pass not yet registered
    int64_t last_frame_idx = -1;
    int32_t slab_lo = 0, slab_hi = -1;               // x-range owned by this context; hi < 0 = whole grid
    DevBuf dense_log;
    // frame-sharded exchange: the routing plan, the plane histogram and the row of totals live in device memory
    DevBuf ex_hist, ex_plan, ex_row;
    uint32_t ex_ranks = 0;
    int32_t ex_self = -1;                 // this context's rank in the exchange (-1: unknown, host-driven API)
    bool plan_valid = false;
    void* recv_buf = nullptr;             // receive buffer of the exchange (plain cudaMalloc so that it can be IPC-exported)
    size_t recv_cap = 0;
    std::vector<void*> ipc_opened;

    // staging for host frames
    float* stage[kRing] = {};
    size_t stage_cap[kRing] = {};
    cudaEvent_t ev_copied[kRing] = {}, ev_free[kRing] = {};
    int ring_pos = 0;
    cudaEvent_t ev_upload[kTickets] = {};
    uint64_t uploads = 0;                 // ticket of the most recent host push
    // host staging pool (pcf_submit_*): clip-and-pack threads + pinned slots, see pcf_stager.hpp
    std::unique_ptr<pcf::Stager> stager;
    std::vector<cudaEvent_t> slot_ev;     // per pinned slot: its last upload has left the slot
    std::vector<cudaEvent_t> raw_ev;      // per raw lane: its last unstaged upload has left the caller's buffer
    uint64_t raw_seen = 0;                // raw uploads already waited for by a drain
    uint64_t staged_dropped = 0;          // clouds dropped by pcf_reset before a staging thread took them

    // scratch
    DevBuf scan1, scan2, tmpA, tmpB, tmpC, tmpD, hist, sort_tab, keysA, keysB, valsA, valsB, sorted, uv_cell, uv_off, nidx,
        sc_a, sc_b, sc_c, flags, slots, cand, res_dev, total_dev, sc_keys, sc_ids, sc_order, sc_okeys, sc_tab;
    int score_unroll = 1;                 // PCF_SCORE_UNR: cylinder tests evaluated back to back in k_score (1, 2 or 4; measured: no gain)
    bool score_balance = true;            // PCF_SCORE_BALANCE=0 keeps the x-major voxel -> lane assignment
    int coop_slots = 8;                   // PCF_COOP_SLOTS: 8 or 16 points per voxel and round in k_score_coop
    int score_coop = -1;                  // PCF_SCORE_COOP: 1 always / 0 never use k_score_coop on canonical schedules (-1: by point density)
    uint32_t* total_host = nullptr;       // pinned, 4 words
    // host results (pinned)
    void* res_host = nullptr;
    size_t res_host_cap = 0;
    void* st_host = nullptr;
    size_t st_host_cap = 0;

    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
    float t_update = 0, t_extract_dev = 0, t_extract_d2h = 0;
    pcf_stats stats{};
    // derived state valid after prepare_sorted()
    uint64_t n_points = 0;
    bool sorted_valid = false;
};

namespace {

int fail(pcf_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(c, PCF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// PCF_TRACE=1: synchronise after every launch and print its name and duration (debugging / per-kernel tables)
#define TRACE_BEGIN(c) std::chrono::steady_clock::time_point t0_; if ((c)->trace) { cudaStreamSynchronize((c)->stream); t0_ = std::chrono::steady_clock::now(); }
#define TRACE_END(c, name)                                                                                         \
    if ((c)->trace) {                                                                                              \
        cudaError_t e_ = cudaStreamSynchronize((c)->stream);                                                       \
        fprintf(stderr, "[pcf] %-48s %10.3f ms %s\n", name,                                                         \
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_).count(),          \
                e_ == cudaSuccess ? "" : cudaGetErrorString(e_));                                                  \
    }
#define LAUNCH(c, kernel, grid, block, ...)                      \
    do {                                                         \
        TRACE_BEGIN(c)                                           \
        kernel<<<(grid), (block), 0, (c)->stream>>>(__VA_ARGS__); \
        (c)->stats.kernel_launches++;                            \
        TRACE_END(c, #kernel)                                    \
    } while (0)

#define LAUNCH_SMEM(c, kernel, grid, block, smem, ...)              \
    do {                                                             \
        TRACE_BEGIN(c)                                               \
        kernel<<<(grid), (block), (smem), (c)->stream>>>(__VA_ARGS__); \
        (c)->stats.kernel_launches++;                                \
        TRACE_END(c, #kernel)                                        \
    } while (0)

inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

int reserve(pcf_ctx* c, DevBuf& b, size_t bytes, bool keep = false) {
    if (bytes <= b.cap) return PCF_OK;
    size_t want = keep ? std::max(bytes, b.cap * 2) : bytes + bytes / 8;
    void* np = nullptr;
    CU(cudaMalloc(&np, want));
    if (keep && b.p && b.cap) {
        CU(cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    if (b.p) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(b.p));
    }
    b.p = np;
    b.cap = want;
    return PCF_OK;
}

float thr_hi(double v) {   // smallest float f with double(f) >= v  :  (double)x >= v  <=>  x >= f
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, INFINITY);
    return f;
}
float thr_lo(double v) {   // largest float f with double(f) <= v   :  (double)x <= v  <=>  x <= f
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -INFINITY);
    return f;
}

int build_grid_params(pcf_ctx* c) {
    const pcf_config& cfg = c->cfg;
    GridParams& g = c->g;
    if (cfg.k_neighbourhood != 2) return fail(c, PCF_ERR_INVALID, "only k_neighbourhood=2 is supported (OG.hpp:334 hard-codes 125 probes)");
    if (cfg.walk_k < 0 || cfg.walk_k > 3) return fail(c, PCF_ERR_INVALID, "walk_k must be in [0,3]");
    for (int a = 0; a < 3; a++) {
        double lo = cfg.box[2 * a], hi = cfg.box[2 * a + 1];
        if (!(cfg.res[a] > 0.f) || !(hi > lo)) return fail(c, PCF_ERR_INVALID, "bad box/resolution on axis %d", a);
        g.min[a] = lo;
        g.res[a] = (double)cfg.res[a];
        g.inv_res[a] = 1.0 / g.res[a];
        g.half_res[a] = g.res[a] / 2.0;
        g.lo[a] = thr_lo(lo);
        g.hi[a] = thr_hi(hi);
        double d = (hi - lo) / g.res[a];
        if (d >= 1048575.0) return fail(c, PCF_ERR_INVALID, "grid dimension exceeds the 20-bit hash field (OG.hpp:151-165)");
        g.dim[a] = (int)d;                       // truncation, OG.hpp:623-625
        g.n1[a] = (uint32_t)g.dim[a] + 1;
    }
    g.nzp = (g.n1[2] + 31u) & ~31u;
    uint64_t plane = (uint64_t)g.n1[1] * g.nzp;
    g.cells = (uint64_t)g.n1[0] * plane;
    for (int a = 0; a < 3; a++) g.nb[a] = (g.n1[a] + 63u) / 64u;
    g.phys_cells = ((uint64_t)g.nb[0] * g.nb[1] * g.nb[2]) << 18;
    if (g.cells >= 0xFFFFFFFFull || g.phys_cells >= 0xFFFFFFFFull || plane >= 0xFFFFFFFFull)
        return fail(c, PCF_ERR_INVALID, "grid has %llu cells; the cell index is 32 bit", (unsigned long long)std::max(g.cells, g.phys_cells));
    g.plane_cells = (uint32_t)plane;
    g.clip_lo = thr_lo(cfg.clip_zmin);
    g.clip_hi = thr_hi(cfg.clip_zmax);
    g.walk_k = cfg.walk_k;
    for (int s = 0; s < 7; s++) g.walk_step[s] = 0.f;
    for (int i = -cfg.walk_k; i <= cfg.walk_k; i++) g.walk_step[i + cfg.walk_k] = (float)((double)i * g.res[0]);   // OG.hpp:405 uses xres_
    g.min_neighbours = cfg.min_neighbours;
    g.ball_radius_f = (float)cfg.ball_radius;
    g.cylinder_radius = cfg.cylinder_radius;
    g.cylinder_thr = thr_hi(cfg.cylinder_radius);     // smallest float f with double(f) >= radius
    return PCF_OK;
}

// ---- device-wide exclusive scan (in == out allowed); optional total to total_dev[slot] --------------------
// POPC: the input words are bitmaps and their population counts are scanned (occupancy bitmap -> rank)
template <bool POPC = false>
int scan_u32(pcf_ctx* c, const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* total_dev) {
    if (n == 0) {
        if (total_dev) CU(cudaMemsetAsync(total_dev, 0, 4, c->stream));
        return PCF_OK;
    }
    uint32_t nb1 = div_up(n, kChunk);
    if (nb1 == 1) {
        LAUNCH(c, k_scan_tiles<POPC>, 1, kBlock, in, out, n, (const uint32_t*)nullptr, total_dev);
        return PCF_OK;
    }
    int rc = reserve(c, c->scan1, (size_t)nb1 * 4);
    if (rc) return rc;
    uint32_t* s1 = (uint32_t*)c->scan1.p;
    LAUNCH(c, k_block_sums<POPC>, nb1, kBlock, in, n, s1);
    uint32_t nb2 = div_up(nb1, kChunk);
    if (nb2 == 1) {
        LAUNCH(c, k_scan_tiles<false>, 1, kBlock, s1, s1, (uint64_t)nb1, (const uint32_t*)nullptr, (uint32_t*)nullptr);
    } else {
        rc = reserve(c, c->scan2, (size_t)nb2 * 4);
        if (rc) return rc;
        uint32_t* s2 = (uint32_t*)c->scan2.p;
        LAUNCH(c, k_block_sums<false>, nb2, kBlock, s1, (uint64_t)nb1, s2);
        LAUNCH(c, k_scan_tiles<false>, 1, kBlock, s2, s2, (uint64_t)nb2, (const uint32_t*)nullptr, (uint32_t*)nullptr);   // nb2 <= 2048
        LAUNCH(c, k_scan_tiles<false>, nb2, kBlock, s1, s1, (uint64_t)nb1, s2, (uint32_t*)nullptr);
    }
    LAUNCH(c, k_scan_tiles<POPC>, nb1, kBlock, in, out, n, s1, total_dev);
    return PCF_OK;
}

int read_total(pcf_ctx* c, const uint32_t* total_dev, uint32_t* out) {
    CU(cudaMemcpyAsync(c->total_host, total_dev, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->stats.d2h_bytes += 4;
    *out = c->total_host[0];
    return PCF_OK;
}

int ensure_log(pcf_ctx* c, uint32_t need_chunks) {
    if (need_chunks <= c->cap_chunks) return PCF_OK;
    if (need_chunks > kMaxChunks) return fail(c, PCF_ERR_CAPACITY, "point log limit reached (%u chunks of %d input points)", kMaxChunks, kWChunk);
    uint32_t cap = std::max<uint32_t>(need_chunks, std::min<uint64_t>((uint64_t)c->cap_chunks * 2, kMaxChunks));
    float4* nl = nullptr;
    uint32_t *nc = nullptr, *nfr = nullptr;
    CU(cudaMalloc(&nl, (size_t)cap * kWChunk * sizeof(float4)));
    CU(cudaMalloc(&nc, (size_t)cap * 4));
    CU(cudaMalloc(&nfr, (size_t)cap * 4));
    if (c->n_chunks) {
        CU(cudaMemcpyAsync(nl, c->log, (size_t)c->n_chunks * kWChunk * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(nc, c->chunk_count, (size_t)c->n_chunks * 4, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(nfr, c->chunk_frame, (size_t)c->n_chunks * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    if (c->log) CU(cudaFree(c->log));
    if (c->chunk_count) CU(cudaFree(c->chunk_count));
    if (c->chunk_frame) CU(cudaFree(c->chunk_frame));
    c->log = nl;
    c->chunk_count = nc;
    c->chunk_frame = nfr;
    c->cap_chunks = cap;
    return PCF_OK;
}

// occupancy rank (cell -> compact voxel id) from the bitmap (shared by update and extraction).  The bitmap is maintained
// by the ingest kernels, so this reads cells/8 bytes twice instead of sweeping the 4-byte-per-cell grid.
int build_occupancy(pcf_ctx* c) {
    if (!c->occ_dirty) return PCF_OK;
    if (c->occ_from_grid) {
        LAUNCH(c, k_cells_to_bits, div_up(c->n_words, kBlock), kBlock, c->first_frame, c->g, c->occ_bits, c->n_words);
        c->occ_from_grid = false;
    }
    uint32_t* tot = (uint32_t*)c->total_dev.p;
    int rc = scan_u32<true>(c, c->occ_bits, c->occ_rank, c->n_words, tot);
    if (rc) return rc;
    rc = read_total(c, tot, &c->n_vox);
    if (rc) return rc;
    c->occ_dirty = false;
    return PCF_OK;
}

// holder registration of passes that have not been followed by a frame yet (OG.hpp:443-449)
int flush_holders(pcf_ctx* c) {
    if (c->pending_holder.empty()) return PCF_OK;
    if (!c->holder) {
        CU(cudaMalloc(&c->holder, c->g.cells * 4));
        CU(cudaMemsetAsync(c->holder, 0, c->g.cells * 4, c->stream));
    }
    // occupancy is unchanged since those passes (no frame was pushed in between), so occ_bits is current
    for (auto& pr : c->pending_holder) {
        uint32_t n = pr.second - pr.first;
        if (!n) continue;
        LAUNCH(c, k_holder<true>, div_up(n, kBlock), kBlock, (const uint32_t*)c->n_cell.p, (const float4*)c->n_nrm.p, pr.first,
               pr.second, c->g, c->occ_bits, c->holder);
        LAUNCH(c, k_holder<false>, div_up(n, kBlock), kBlock, (const uint32_t*)c->n_cell.p, (const float4*)c->n_nrm.p, pr.first,
               pr.second, c->g, c->occ_bits, c->holder);
    }
    c->pending_holder.clear();
    return PCF_OK;
}

// one launch over `nf` equally sized clouds resident in device memory.  Batch = IngestBatch (up to 256 frames, 24 KB of
// kernel parameters) or IngestBatch1 (one frame, 144 bytes: the per-frame host path launches 200 times per step).
inline uint32_t* ingest_occ(pcf_ctx* c) {
    if (!c->ingest_bits) { c->occ_from_grid = true; return nullptr; }
    return c->occ_bits;
}
// stride: floats per point; 0 = organized PointCloud2 layout described by `rows` (generic kernel only)
template <class Batch>
int launch_ingest_t(pcf_ctx* c, const float* pts_dev, uint64_t frame_stride, uint32_t n, uint32_t nf, uint32_t stride,
                    const double* poses, uint32_t first_frame_idx, const float* explicit_vp, const RowLayout* rows) {
    static thread_local Batch b;
    const GridParams& g = c->g;
    uint32_t chunks = div_up(n, kWChunk);
    b.pts = pts_dev;
    b.frame_stride = frame_stride;
    b.n = n;
    b.n_frames = nf;
    b.first_frame_idx = first_frame_idx;
    b.chunk_base = c->n_chunks;
    b.chunks_per_frame = chunks;
    b.explicit_vp = explicit_vp ? 1u : 0u;
    for (int i = 0; i < 3; i++) b.vp[i] = explicit_vp ? explicit_vp[i] : 0.f;
    b.vp[3] = 1.f;
    for (uint32_t f = 0; f < nf; f++)
        for (int i = 0; i < 12; i++) b.T[f][i] = poses[(size_t)f * 16 + i];
    RowLayout rl{n ? n : 1u, stride, 0};
    if (rows) rl = *rows;
    if (explicit_vp) {       // pcf_add_points: cloud already in the fusion frame
        dim3 grid(div_up(chunks, kWarps), nf, 1);
        LAUNCH(c, (k_ingest<0, true, Batch>), grid, kBlock, b, rl, g, c->first_frame, ingest_occ(c), c->log, c->chunk_count, c->chunk_frame, c->vp_table);
        CU(cudaGetLastError());
        c->n_chunks += chunks * nf;
        return PCF_OK;
    }
    // B200 path: bulk-async ring (needs 16-byte aligned chunks); anything else takes the generic kernel
    const bool aligned = ((uintptr_t)pts_dev % 16 == 0) && ((frame_stride * 4) % 16 == 0 || nf == 1);
    const uint32_t total = chunks * nf;
    const uint32_t grid_bulk = std::min<uint32_t>((uint32_t)(c->sm_count * c->ctas_per_sm), div_up(total, kWarps));
    if (c->use_bulk && aligned && stride == 4 && !rows) {
        LAUNCH_SMEM(c, (k_ingest_bulk<16, kBulkMinBlocks, kBulkRounds, 1, Batch>), grid_bulk, kBlock, kWarps * kWChunk * 16, b, g, c->first_frame, ingest_occ(c), c->log, c->chunk_count, c->chunk_frame, c->vp_table);
    } else if (c->use_bulk && aligned && stride == 3 && n % 4 == 0 && !rows) {
        LAUNCH_SMEM(c, (k_ingest_bulk<12, kBulkMinBlocks, kBulkRounds, 1, Batch>), grid_bulk, kBlock, kWarps * kWChunk * 12, b, g, c->first_frame, ingest_occ(c), c->log, c->chunk_count, c->chunk_frame, c->vp_table);
    } else {
        dim3 grid(div_up(chunks, kWarps), nf, 1);
        if (stride == 4 && !rows) LAUNCH(c, (k_ingest<4, false, Batch>), grid, kBlock, b, rl, g, c->first_frame, ingest_occ(c), c->log, c->chunk_count, c->chunk_frame, c->vp_table);
        else if (stride == 3 && !rows) LAUNCH(c, (k_ingest<3, false, Batch>), grid, kBlock, b, rl, g, c->first_frame, ingest_occ(c), c->log, c->chunk_count, c->chunk_frame, c->vp_table);
        else LAUNCH(c, (k_ingest<0, false, Batch>), grid, kBlock, b, rl, g, c->first_frame, ingest_occ(c), c->log, c->chunk_count, c->chunk_frame, c->vp_table);
    }
    CU(cudaGetLastError());
    c->n_chunks += chunks * nf;
    return PCF_OK;
}
int launch_ingest(pcf_ctx* c, const float* pts_dev, uint64_t frame_stride, uint32_t n, uint32_t nf, uint32_t stride,
                  const double* poses, uint32_t first_frame_idx, const float* explicit_vp = nullptr, const RowLayout* rows = nullptr) {
    if (nf == 1) return launch_ingest_t<IngestBatch1>(c, pts_dev, frame_stride, n, nf, stride, poses, first_frame_idx, explicit_vp, rows);
    return launch_ingest_t<IngestBatch>(c, pts_dev, frame_stride, n, nf, stride, poses, first_frame_idx, explicit_vp, rows);
}

int check_frame_idx(pcf_ctx* c, uint32_t first, uint32_t count) {
    if ((int64_t)first <= c->last_frame_idx) return fail(c, PCF_ERR_INVALID, "frame_idx %u does not increase (last %lld)", first, (long long)c->last_frame_idx);
    if ((uint64_t)first + count > c->cfg.max_frames || (uint64_t)first + count >= kEmpty) return fail(c, PCF_ERR_CAPACITY, "frame_idx %u exceeds max_frames %u", first + count - 1, c->cfg.max_frames);
    return PCF_OK;
}

// (cell, slot) pairs sorted by cell, stable in slot order; then the sorted point stream and per-voxel CSR
int prepare_sorted(pcf_ctx* c) {
    if (c->sorted_valid) return PCF_OK;
    // (pending holder registrations are not needed here: they only matter once a later frame lands, and
    //  pcf_push_* flushes them before that frame is integrated)
    int rc = build_occupancy(c);
    if (rc) return rc;
    uint32_t* tot = (uint32_t*)c->total_dev.p;
    // number of kept points
    uint32_t P = 0;
    if (c->n_chunks) {
        rc = reserve(c, c->tmpB, (size_t)c->n_chunks * 4);
        if (rc) return rc;
        rc = scan_u32(c, c->chunk_count, (uint32_t*)c->tmpB.p, c->n_chunks, tot);
        if (rc) return rc;
        rc = read_total(c, tot, &P);
        if (rc) return rc;
    }
    c->n_points = P;
    c->stats.points_kept = P;
    c->stats.occupied_voxels = c->n_vox;
    rc = reserve(c, c->uv_off, ((size_t)c->n_vox + 1) * 4);
    if (rc) return rc;
    rc = reserve(c, c->uv_cell, ((size_t)c->n_vox + 1) * 4);
    if (rc) return rc;
    rc = reserve(c, c->nidx, ((size_t)c->n_vox + 1) * 4);
    if (rc) return rc;
    if (P == 0) {
        CU(cudaMemsetAsync(c->uv_off.p, 0, ((size_t)c->n_vox + 1) * 4, c->stream));
        c->sorted_valid = true;
        return PCF_OK;
    }
    if ((rc = reserve(c, c->keysA, (size_t)P * 4))) return rc;
    if ((rc = reserve(c, c->keysB, (size_t)P * 4))) return rc;
    if ((rc = reserve(c, c->valsA, (size_t)P * 4))) return rc;
    if ((rc = reserve(c, c->valsB, (size_t)P * 4))) return rc;
    if ((rc = reserve(c, c->sorted, (size_t)P * sizeof(float4)))) return rc;

    int bits = 1;
    while ((1ull << bits) < c->g.cells) bits++;
    // pass M: the top digit of the CELL, straight from the log, writing the voxel's compact id as the key; then local LSD
    // passes over (id - first id of the bucket) inside each bucket: as many bits as the fullest bucket has voxels
    const int msd_bits = std::min(8, bits), rem = bits - msd_bits;
    const uint32_t tiles_m = div_up(c->n_chunks, kWarps);
    const uint32_t tiles_max = div_up(P, kChunk) + 256;              // every bucket may end with a partial tile
    if ((rc = reserve(c, c->hist, (size_t)256 * std::max(tiles_m, tiles_max) * 4))) return rc;
    if ((rc = reserve(c, c->sort_tab, (size_t)tiles_max * sizeof(SortTile) + (257 + 258 + 257) * 4))) return rc;
    uint32_t* hist = (uint32_t*)c->hist.p;
    SortTile* tab = (SortTile*)c->sort_tab.p;
    uint32_t* bucket_start = (uint32_t*)(tab + tiles_max);
    uint32_t* tile_base = bucket_start + 257;
    uint32_t* bucket_rank0 = tile_base + 258;
    uint32_t *kin = nullptr, *vin = nullptr, *kout = (uint32_t*)c->keysA.p, *vout = (uint32_t*)c->valsA.p;
    {
        SortSrc src{};
        src.log = c->log; src.chunk_count = c->chunk_count; src.n_chunks = c->n_chunks; src.n_tiles = tiles_m;
        src.occ_bits = c->occ_bits; src.occ_rank = c->occ_rank;
        const uint32_t mask = (1u << msd_bits) - 1;
        LAUNCH(c, k_sort_hist<true>, tiles_m, kBlock, src, (uint32_t)rem, mask, hist);
        if ((rc = scan_u32(c, hist, hist, (uint64_t)256 * tiles_m, nullptr))) return rc;
        LAUNCH(c, k_sort_scatter<true>, tiles_m, kBlock, src, (uint32_t)rem, mask, hist, kout, vout);
        kin = kout; vin = vout;
        kout = (uint32_t*)c->keysB.p; vout = (uint32_t*)c->valsB.p;
    }
    LAUNCH(c, k_sort_bucket_tiles, 1, 256, (const uint32_t*)hist, tiles_m, 1u << msd_bits, P, bucket_start, tile_base, c->occ_bits, c->occ_rank,
           (uint32_t)rem, c->g.cells, c->n_vox, bucket_rank0);
    CU(cudaMemcpyAsync(c->total_host, tile_base + 256, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->stats.d2h_bytes += 8;
    const uint32_t n_tiles = c->total_host[0], max_bucket_vox = c->total_host[1];
    int kbits = 0;
    while ((1ull << kbits) < max_bucket_vox) kbits++;
    const int lp = (kbits + 7) / 8, per = lp ? (kbits + lp - 1) / lp : 0;
    if (lp > 0) {
        LAUNCH(c, k_sort_tile_table, div_up(std::max<uint32_t>(n_tiles, 1), kBlock), kBlock, (const uint32_t*)bucket_start, (const uint32_t*)tile_base,
               (const uint32_t*)bucket_rank0, tab);
        for (int p = 0; p < lp; p++) {
            SortSrc src{};
            src.keys = kin; src.vals = vin; src.tab = tab; src.n_tiles_dev = tile_base + 256;
            const uint32_t shift = (uint32_t)(p * per), mask = (1u << per) - 1;
            LAUNCH(c, k_sort_hist<false>, n_tiles, kBlock, src, shift, mask, hist);
            if ((rc = scan_u32(c, hist, hist, (uint64_t)256 * n_tiles, nullptr))) return rc;
            LAUNCH(c, k_sort_scatter<false>, n_tiles, kBlock, src, shift, mask, hist, kout, vout);
            kin = kout; vin = vout;
            kout = (kin == (uint32_t*)c->keysA.p) ? (uint32_t*)c->keysB.p : (uint32_t*)c->keysA.p;
            vout = (vin == (uint32_t*)c->valsA.p) ? (uint32_t*)c->valsB.p : (uint32_t*)c->valsA.p;
        }
    }
    LAUNCH(c, k_gather_points, div_up(P, kBlock), kBlock, c->log, kin, vin, (uint64_t)P, (float4*)c->sorted.p, (uint32_t*)c->uv_cell.p,
           (uint32_t*)c->uv_off.p);
    CU(cudaGetLastError());
    c->sorted_valid = true;
    return PCF_OK;
}

// logical cell range [lo, hi) of the x-slab this context works on (whole grid when no slab is set)
void slab_cells(const pcf_ctx* c, uint32_t& lo, uint32_t& hi) {
    const uint64_t plane = (uint64_t)c->g.plane_cells;
    lo = c->slab_hi < 0 ? 0u : (uint32_t)((uint64_t)c->slab_lo * plane);
    hi = c->slab_hi < 0 ? 0xFFFFFFFFu : (uint32_t)std::min<uint64_t>(c->g.cells, (uint64_t)c->slab_hi * plane);
}

// scoring of every voxel that has a normal -> sc_a (centroid,count) sc_b (sd,mean_dist) sc_c (sd_dist); nidx map
int run_scoring(pcf_ctx* c) {
    int rc = prepare_sorted(c);
    if (rc) return rc;
    if (c->n_vox) CU(cudaMemsetAsync(c->nidx.p, 0xFF, (size_t)c->n_vox * 4, c->stream));
    uint32_t nn = c->n_normals;
    c->stats.normals_found = nn;
    if (!nn) return PCF_OK;
    if ((rc = reserve(c, c->sc_a, (size_t)nn * 16))) return rc;
    if ((rc = reserve(c, c->sc_b, (size_t)nn * 16))) return rc;
    if ((rc = reserve(c, c->sc_c, (size_t)nn * 4))) return rc;
    LAUNCH(c, k_map_normals, div_up(nn, kBlock), kBlock, (const uint32_t*)c->n_cell.p, nn, c->occ_bits, c->occ_rank, (uint32_t*)c->nidx.p);
    ScoreOut so{(float4*)c->sc_a.p, (float4*)c->sc_b.p, (float*)c->sc_c.p};
    uint32_t cell_lo, cell_hi;
    slab_cells(c, cell_lo, cell_hi);
    uint32_t* fault = (uint32_t*)c->total_dev.p + 8;
    CU(cudaMemsetAsync(fault, 0, 32, c->stream));
    // work-balanced voxel -> lane assignment: stable one-pass counting sort of the voxel ids by an 8-bit work key
    const uint32_t* order = nullptr;
    if (c->score_balance && nn > 4096) {
        const uint32_t nt = div_up(nn, kChunk);
        if ((rc = reserve(c, c->sc_keys, (size_t)nn * 4))) return rc;
        if ((rc = reserve(c, c->sc_ids, (size_t)nn * 4))) return rc;
        if ((rc = reserve(c, c->sc_okeys, (size_t)nn * 4))) return rc;
        if ((rc = reserve(c, c->sc_order, (size_t)nn * 4))) return rc;
        if ((rc = reserve(c, c->hist, (size_t)256 * nt * 4))) return rc;
        if ((rc = reserve(c, c->sc_tab, (size_t)nt * sizeof(SortTile) + 16))) return rc;
        SortTile* tab = (SortTile*)c->sc_tab.p;
        uint32_t* nt_dev = (uint32_t*)(tab + nt);
        LAUNCH(c, k_score_work, div_up(nn, kBlock), kBlock, (const uint32_t*)c->n_cell.p, (const float4*)c->n_nrm.p, nn, c->g, c->occ_bits,
               c->occ_rank, (const uint32_t*)c->uv_off.p, (uint32_t*)c->sc_keys.p, (uint32_t*)c->sc_ids.p, cell_lo, cell_hi);
        LAUNCH(c, k_sort_flat_tiles, div_up(nt, kBlock), kBlock, nn, tab, nt_dev);
        SortSrc src{};
        src.keys = (const uint32_t*)c->sc_keys.p; src.vals = (const uint32_t*)c->sc_ids.p; src.tab = tab; src.n_tiles_dev = nt_dev;
        uint32_t* hist = (uint32_t*)c->hist.p;
        LAUNCH(c, k_sort_hist<false>, nt, kBlock, src, 0u, 255u, hist);
        if ((rc = scan_u32(c, hist, hist, (uint64_t)256 * nt, nullptr))) return rc;
        LAUNCH(c, k_sort_scatter<false>, nt, kBlock, src, 0u, 255u, hist, (uint32_t*)c->sc_okeys.p, (uint32_t*)c->sc_order.p);
        order = (const uint32_t*)c->sc_order.p;
    }
    // canonical schedule: one update pass that saw every point currently in the log
    const bool simple = c->marks.size() == 1 && c->marks[0] == c->n_chunks * (uint32_t)kWChunk && !c->holder;
#define SCORE_ARGS order, (const uint32_t*)c->n_cell.p, (const float4*)c->n_nrm.p, (const uint32_t*)c->n_mark.p, nn, c->g, c->occ_bits, c->occ_rank, \
                   (const uint32_t*)c->uv_off.p, (const uint32_t*)c->nidx.p, (const float4*)c->sorted.p, (const uint32_t*)c->holder, so,  \
                   (uint32_t)c->n_points, (const uint32_t*)c->uv_cell.p, fault, cell_lo, cell_hi
    // dense buffers (the scans of real surfaces): cooperative kernel; a handful of points per voxel (C5's synthetic sheets):
    // one thread per voxel wastes less
    const bool coop = simple && c->score_coop != 0 && (c->score_coop > 0 || c->n_points >= 8ull * std::max<uint32_t>(c->n_vox, 1u));
#define COOP_ARGS order, (const uint32_t*)c->n_cell.p, (const float4*)c->n_nrm.p, nn, c->g, c->occ_bits, c->occ_rank, (const uint32_t*)c->uv_off.p, \
                  (const float4*)c->sorted.p, so, (uint32_t)c->n_points, (const uint32_t*)c->uv_cell.p, fault, cell_lo, cell_hi
    if (coop && c->coop_slots == 8) LAUNCH(c, k_score_coop<8>, div_up(nn, kCoopWarps * 32), kCoopWarps * 32, COOP_ARGS);
    else if (coop) LAUNCH(c, k_score_coop<16>, div_up(nn, kCoopWarps * 32), kCoopWarps * 32, COOP_ARGS);
#undef COOP_ARGS
    else if (!simple) LAUNCH(c, (k_score<false, 1>), div_up(nn, 128), 128, SCORE_ARGS);
    else if (c->score_unroll >= 4) LAUNCH(c, (k_score<true, 4>), div_up(nn, 128), 128, SCORE_ARGS);
    else if (c->score_unroll >= 2) LAUNCH(c, (k_score<true, 2>), div_up(nn, 128), 128, SCORE_ARGS);
    else LAUNCH(c, (k_score<true, 1>), div_up(nn, 128), 128, SCORE_ARGS);
#undef SCORE_ARGS
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->total_host + 8, fault, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->total_host[8]) {
        const uint32_t* f = c->total_host + 8;
        return fail(c, PCF_ERR_INTERNAL, "k_score consistency fault %u: voxel record %u, words %u %u %u %u %u", f[0], f[1], f[2], f[3], f[4], f[5], f[6]);
    }
    return PCF_OK;
}

int ensure_pinned(pcf_ctx* c, void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return PCF_OK;
    if (*p) CU(cudaFreeHost(*p));
    *p = nullptr; *cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CU(cudaMallocHost(p, want));
    *cap = want;
    return PCF_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int extract_impl(pcf_ctx* c, int32_t min_count, pcf_result* out) {
    if (!out) return fail(c, PCF_ERR_INVALID, "null result");
    memset(out, 0, sizeof *out);
    CU(cudaEventRecord(c->ev_a, c->stream));
    int rc = run_scoring(c);
    if (rc) return rc;
    uint32_t n_out = 0;
    uint32_t nv = c->n_vox;
    if (nv && c->n_normals) {
        if ((rc = reserve(c, c->flags, (size_t)nv * 4))) return rc;
        if ((rc = reserve(c, c->slots, (size_t)nv * 4))) return rc;
        uint32_t cell_lo, cell_hi;
        slab_cells(c, cell_lo, cell_hi);
        LAUNCH(c, k_extract_flags, div_up(nv, kBlock), kBlock, (const uint32_t*)c->uv_cell.p, nv, (const uint32_t*)c->nidx.p,
               (const float4*)c->sc_a.p, c->g, min_count, (uint32_t*)c->flags.p, cell_lo, cell_hi);
        uint32_t* tot = (uint32_t*)c->total_dev.p;
        if ((rc = scan_u32(c, (uint32_t*)c->flags.p, (uint32_t*)c->slots.p, nv, tot))) return rc;
        if ((rc = read_total(c, tot, &n_out))) return rc;
    }
    // device + host SoA blocks: hash(8) centroid(12) normal(12) sd(12) mean(4) sdd(4) count(4)
    size_t n = n_out;
    size_t o_hash = 0, o_cen = align256(o_hash + n * 8), o_nrm = align256(o_cen + n * 12), o_sd = align256(o_nrm + n * 12),
           o_md = align256(o_sd + n * 12), o_sdd = align256(o_md + n * 4), o_cnt = align256(o_sdd + n * 4), total = align256(o_cnt + n * 4);
    if ((rc = reserve(c, c->res_dev, total))) return rc;
    if ((rc = ensure_pinned(c, &c->res_host, &c->res_host_cap, total))) return rc;
    char* d = (char*)c->res_dev.p;
    if (n) {
        ResultDev r{(uint64_t*)(d + o_hash), (float*)(d + o_cen), (float*)(d + o_nrm), (float*)(d + o_sd), (float*)(d + o_md),
                    (float*)(d + o_sdd), (int32_t*)(d + o_cnt)};
        LAUNCH(c, k_extract_gather, div_up(nv, kBlock), kBlock, (const uint32_t*)c->uv_cell.p, nv, (const uint32_t*)c->nidx.p,
               (const uint32_t*)c->flags.p, (const uint32_t*)c->slots.p, (const float4*)c->n_nrm.p, (const float4*)c->sc_a.p,
               (const float4*)c->sc_b.p, (const float*)c->sc_c.p, c->g, r);
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(c->ev_b, c->stream));
    if (n) {
        CU(cudaMemcpyAsync(c->res_host, d, total, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += total;
    }
    CU(cudaEventRecord(c->ev_c, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventElapsedTime(&c->t_extract_dev, c->ev_a, c->ev_b));
    CU(cudaEventElapsedTime(&c->t_extract_d2h, c->ev_b, c->ev_c));
    char* h = (char*)c->res_host;
    out->n = n;
    out->hash = (const uint64_t*)(h + o_hash);
    out->centroid = (const float*)(h + o_cen);
    out->normal = (const float*)(h + o_nrm);
    out->sd = (const float*)(h + o_sd);
    out->mean_dist = (const float*)(h + o_md);
    out->sd_dist = (const float*)(h + o_sdd);
    out->count = (const int32_t*)(h + o_cnt);
    return PCF_OK;
}

void destroy_impl(pcf_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stager) { c->stager->drain(); c->stager.reset(); }      // joins the staging threads, frees the pinned slots
    for (cudaEvent_t e : c->slot_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->raw_ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    DevBuf* bufs[] = {&c->n_cell, &c->n_nrm, &c->n_mark, &c->scan1, &c->scan2, &c->tmpA, &c->tmpB, &c->tmpC, &c->tmpD, &c->hist, &c->sort_tab,
                      &c->keysA, &c->keysB, &c->valsA, &c->valsB, &c->sorted, &c->uv_cell, &c->uv_off, &c->nidx, &c->sc_a, &c->sc_b,
                      &c->sc_c, &c->flags, &c->slots, &c->cand, &c->res_dev, &c->total_dev, &c->dense_log, &c->sc_keys, &c->sc_ids,
                      &c->sc_order, &c->sc_okeys, &c->sc_tab, &c->ex_hist, &c->ex_plan, &c->ex_row, &c->upd_cell, &c->upd_nrm};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    void* raw[] = {c->first_frame, c->holder, c->nrm_bits, c->occ_bits, c->occ_rank, c->vp_table, c->log, c->chunk_count, c->chunk_frame, c->recv_buf};
    for (void* p : raw) if (p) cudaFree(p);
    for (int i = 0; i < kRing; i++) {
        if (c->stage[i]) cudaFree(c->stage[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    }
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    for (int i = 0; i < kTickets; i++) if (c->ev_upload[i]) cudaEventDestroy(c->ev_upload[i]);
    if (c->total_host) cudaFreeHost(c->total_host);
    if (c->res_host) cudaFreeHost(c->res_host);
    if (c->st_host) cudaFreeHost(c->st_host);
    cudaEvent_t evs[] = {c->ev_a, c->ev_b, c->ev_c};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int reset_grid_state(pcf_ctx* c) {
    LAUNCH(c, k_fill_u32, 148 * 8, 512, c->first_frame, c->g.phys_cells, kEmpty);
    CU(cudaMemsetAsync(c->nrm_bits, 0, (c->n_words + 2) * 4, c->stream));
    CU(cudaMemsetAsync(c->occ_bits, 0, (c->n_words + 2) * 4, c->stream));
    CU(cudaMemsetAsync(c->vp_table, 0, (size_t)c->cfg.max_frames * sizeof(float4), c->stream));
    if (c->holder) CU(cudaMemsetAsync(c->holder, 0, c->g.cells * 4, c->stream));
    c->n_chunks = 0;
    c->merged_chunks = 0;
    c->upd_open = false;
    c->n_normals = 0;
    c->n_vox = 0;
    c->n_points = 0;
    c->marks.clear();
    c->pending_holder.clear();
    c->occ_dirty = true;
    c->occ_from_grid = false;
    c->log_installed = false;
    c->sorted_valid = false;
    c->last_frame_idx = -1;
    c->slab_lo = 0;
    c->slab_hi = -1;
    return PCF_OK;
}

// Every entry point that touches the grid first lets the staging pool hand over what was submitted before it:
// submission order == integration order, whatever mix of pcf_submit_* and direct calls the caller uses.
int drain_staged(pcf_ctx* c) {
    if (!c->stager) return PCF_OK;
    int rc = c->stager->drain();
    if (rc < 0 && c->err.empty()) c->err = "a staged frame failed to integrate";
    const uint64_t raw = c->stager->raw_pushed();
    if (raw != c->raw_seen) {            // clouds uploaded unstaged are read by the copy engine until their copy completes
        cudaStreamSynchronize(c->copy_stream);
        c->raw_seen = raw;
    }
    return rc < 0 ? rc : PCF_OK;
}
#define ENTER(c)                                  \
    do {                                          \
        CU(cudaSetDevice((c)->device));           \
        int rc_ = drain_staged(c);                \
        if (rc_) return rc_;                      \
    } while (0)

}  // namespace

// ============================================ C ABI ===================================================
extern "C" {

void pcf_default_config(pcf_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    const double box[6] = {-0.8, 1.8, -1.5, 1.5, 0.0, 1.0};   // launch:8
    memcpy(cfg->box, box, sizeof box);
    cfg->res[0] = cfg->res[1] = cfg->res[2] = (float)0.005;   // node.cpp:91,161
    cfg->clip_zmin = 0.28;                                    // node.cpp:92
    cfg->clip_zmax = 0.6;                                     // node.cpp:93
    cfg->k_neighbourhood = 2;                                 // node.cpp:163
    cfg->walk_k = 3;                                          // node.cpp:311
    cfg->min_neighbours = 20;                                 // OG.hpp:352
    cfg->cylinder_radius = 0.001;                             // OG.hpp:36
    cfg->ball_radius = 0.015;                                 // OG.hpp:35
    cfg->device = 0;
    cfg->max_frames = 1u << 16;
    cfg->log_capacity_hint = 0;
    cfg->stage_threads = 0;                                   // auto
    cfg->stage_raw_lanes = 0;                                 // auto
}

int pcf_create(const pcf_config* cfg, pcf_ctx** out) {
    pcf_ctx* c = nullptr;
    if (!cfg || !out) return fail(c, PCF_ERR_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(c, PCF_ERR_NO_DEVICE, "no CUDA device: libpcfusion has no CPU path");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(c, PCF_ERR_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
    pcf_ctx* ctx = new pcf_ctx();
    ctx->cfg = *cfg;
    if (ctx->cfg.max_frames == 0) ctx->cfg.max_frames = 1u << 16;
    ctx->device = cfg->device;
    int rc = build_grid_params(ctx);
    if (rc) { g_create_error = ctx->err; delete ctx; return rc; }
    c = ctx;
    auto bail = [&](int code) { g_create_error = ctx->err; destroy_impl(ctx); return code; };
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, PCF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(PCF_ERR_CUDA); } } while (0)
    CUC(cudaSetDevice(c->device));
    CUC(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    CUC(cudaFuncSetAttribute(k_ingest_bulk<16, kBulkMinBlocks, kBulkRounds, 1, IngestBatch>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * kWChunk * 16));
    CUC(cudaFuncSetAttribute(k_ingest_bulk<12, kBulkMinBlocks, kBulkRounds, 1, IngestBatch>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * kWChunk * 12));
    CUC(cudaFuncSetAttribute(k_ingest_bulk<16, kBulkMinBlocks, kBulkRounds, 1, IngestBatch1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * kWChunk * 16));
    CUC(cudaFuncSetAttribute(k_ingest_bulk<12, kBulkMinBlocks, kBulkRounds, 1, IngestBatch1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * kWChunk * 12));
    {
        const char* t = getenv("PCF_TRACE");
        c->trace = t && atoi(t) > 0;
        const char* e = getenv("PCF_INGEST");
        c->use_bulk = !(e && strcmp(e, "generic") == 0);
        const char* ib = getenv("PCF_INGEST_BITS");
        if (ib) c->ingest_bits = atoi(ib) != 0;
        const char* u = getenv("PCF_SCORE_UNR");
        if (u && atoi(u) > 0) c->score_unroll = atoi(u);
        const char* bl = getenv("PCF_SCORE_BALANCE");
        if (bl) c->score_balance = atoi(bl) != 0;
        const char* cs = getenv("PCF_COOP_SLOTS");
        if (cs && (atoi(cs) == 8 || atoi(cs) == 16)) c->coop_slots = atoi(cs);
        const char* co = getenv("PCF_SCORE_COOP");
        if (co) c->score_coop = atoi(co) != 0 ? 1 : 0;
    }
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    c->n_words = (c->g.cells + 31) / 32;
    CUC(cudaMalloc(&c->first_frame, c->g.phys_cells * 4));
    CUC(cudaMalloc(&c->nrm_bits, (c->n_words + 2) * 4));
    CUC(cudaMalloc(&c->occ_bits, (c->n_words + 2) * 4));
    CUC(cudaMalloc(&c->occ_rank, (c->n_words + 2) * 4));
    CUC(cudaMalloc(&c->vp_table, (size_t)c->cfg.max_frames * sizeof(float4)));
    CUC(cudaMallocHost(&c->total_host, 64));
    for (int i = 0; i < kTickets; i++) CUC(cudaEventCreateWithFlags(&c->ev_upload[i], cudaEventDisableTiming));
    for (int i = 0; i < kRing; i++) {
        CUC(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
    }
    CUC(cudaEventCreate(&c->ev_a));
    CUC(cudaEventCreate(&c->ev_b));
    CUC(cudaEventCreate(&c->ev_c));
    if (reserve(c, c->total_dev, 64)) return bail(PCF_ERR_CUDA);
    if (reset_grid_state(c)) return bail(PCF_ERR_CUDA);
    uint64_t hint = cfg->log_capacity_hint ? cfg->log_capacity_hint : (uint64_t)64 * 307200;
    if (ensure_log(c, std::min<uint64_t>(div_up(hint, kWChunk) + 8, kMaxChunks))) return bail(PCF_ERR_CUDA);
    CUC(cudaStreamSynchronize(c->stream));
#undef CUC
    *out = c;
    return PCF_OK;
}

void pcf_destroy(pcf_ctx* ctx) { destroy_impl(ctx); }

const char* pcf_last_error(const pcf_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int pcf_dims(const pcf_ctx* ctx, int32_t dims[3]) {
    if (!ctx || !dims) return PCF_ERR_INVALID;
    for (int a = 0; a < 3; a++) dims[a] = ctx->g.dim[a];
    return PCF_OK;
}

int pcf_start(pcf_ctx* c) { if (!c) return PCF_ERR_INVALID; c->started = true; return PCF_OK; }
int pcf_stop(pcf_ctx* c) { if (!c) return PCF_ERR_INVALID; c->started = false; return PCF_OK; }
int pcf_reset(pcf_ctx* c) {
    // node.cpp:351-359: start_ = false and clouds_.clear() -- the clouds no staging thread has taken yet are dropped,
    // those already taken (the reference's clouds_processed_ deque) still integrate, the grid is left alone.
    if (!c) return PCF_ERR_INVALID;
    c->started = false;
    if (c->stager) c->staged_dropped += c->stager->drop_queued();
    return PCF_OK;
}

// H2D copy of one host cloud into the device staging ring + the integration launch behind it.  `slot_done`, when given,
// is recorded on the copy stream after the copy (the pinned source may be refilled once it has fired).
static int push_host_cloud_impl(pcf_ctx* c, const void* src_host, size_t bytes, size_t first_float, uint32_t n, uint32_t stride,
                                const RowLayout* rows, const double pose[16], const float* explicit_vp, uint32_t frame_idx,
                                uint32_t n_offered, cudaEvent_t slot_done) {
    int rc = check_frame_idx(c, frame_idx, 1);
    if (rc) return rc;
    if ((rc = flush_holders(c))) return rc;
    uint32_t chunks = div_up(n, kWChunk);
    if ((uint64_t)c->n_chunks + chunks > kMaxChunks) return fail(c, PCF_ERR_CAPACITY, "point log limit reached");
    if ((rc = ensure_log(c, c->n_chunks + chunks))) return rc;
    int s = c->ring_pos;
    c->ring_pos = (c->ring_pos + 1) % kRing;
    if (bytes > c->stage_cap[s]) {
        CU(cudaEventSynchronize(c->ev_free[s]));
        if (c->stage[s]) CU(cudaFree(c->stage[s]));
        c->stage[s] = nullptr;
        CU(cudaMalloc(&c->stage[s], bytes + bytes / 4));
        c->stage_cap[s] = bytes + bytes / 4;
    }
    // copy stream: wait until the kernel that last read this slot is done, then upload
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_free[s], 0));
    if (bytes) CU(cudaMemcpyAsync(c->stage[s], src_host, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(c->ev_copied[s], c->copy_stream));
    if (slot_done) {
        CU(cudaEventRecord(slot_done, c->copy_stream));
    } else {
        c->uploads++;
        CU(cudaEventRecord(c->ev_upload[c->uploads % kTickets], c->copy_stream));
    }
    CU(cudaStreamWaitEvent(c->stream, c->ev_copied[s], 0));
    if (chunks) {
        rc = launch_ingest(c, c->stage[s] + first_float, 0, n, 1, stride, pose, frame_idx, explicit_vp, rows);
        if (rc) return rc;
    }
    CU(cudaEventRecord(c->ev_free[s], c->stream));
    c->last_frame_idx = frame_idx;
    c->occ_dirty = true;
    c->sorted_valid = false;
    c->stats.frames_pushed++;
    c->stats.points_offered += n_offered;
    c->stats.h2d_bytes += bytes;
    return PCF_OK;
}

static int push_host_cloud(pcf_ctx* c, const float* pts_host, uint32_t n, uint32_t stride, const double pose[16],
                           const float* explicit_vp, uint32_t frame_idx) {
    if (!c) return PCF_ERR_INVALID;
    if (!c->started) return PCF_DROPPED;
    if ((!pts_host && n) || !pose || stride < 3) return fail(c, PCF_ERR_INVALID, "bad frame arguments");
    ENTER(c);
    const size_t bytes = (size_t)n * stride * 4;          // pts_host is an [n, stride] float array
    return push_host_cloud_impl(c, pts_host, bytes, 0, n, stride, nullptr, pose, explicit_vp, frame_idx, n, nullptr);
}

int pcf_push_frame(pcf_ctx* c, const float* pts_host, uint32_t n, uint32_t stride, const double pose[16], uint32_t frame_idx) {
    return push_host_cloud(c, pts_host, n, stride, pose, nullptr, frame_idx);
}

static int check_pointcloud2(pcf_ctx* c, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step, uint32_t row_step,
                             uint32_t x_offset, uint32_t y_offset, uint32_t z_offset, const double* pose) {
    if (!data || !pose) return fail(c, PCF_ERR_INVALID, "null cloud / pose");
    if (point_step % 4 || x_offset % 4 || y_offset != x_offset + 4 || z_offset != x_offset + 8 || point_step < x_offset + 12)
        return fail(c, PCF_ERR_INVALID, "PointCloud2 layout not supported: x, y, z must be consecutive 4-byte aligned float32 fields");
    if (row_step < (uint64_t)width * point_step || row_step % 4) return fail(c, PCF_ERR_INVALID, "row_step smaller than width * point_step (or not a multiple of 4)");
    if ((uint64_t)width * height > 0xFFFFFFFFull) return fail(c, PCF_ERR_INVALID, "cloud too large");
    return PCF_OK;
}

// sensor_msgs/PointCloud2 front end (node.cpp:182-216): x, y, z are consecutive float32 fields `x_offset` bytes into
// every `point_step`-byte point, rows are `row_step` bytes apart.  Unlike the reference, which walks only the first
// `row_step` bytes (D6), every row of an organized cloud is integrated.  The message bytes are uploaded as they are --
// [data, last z of the last point], never a byte more -- and the kernel strides over points and rows.
int pcf_push_pointcloud2(pcf_ctx* c, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step, uint32_t row_step,
                         uint32_t x_offset, uint32_t y_offset, uint32_t z_offset, const double pose[16], uint32_t frame_idx) {
    if (!c) return PCF_ERR_INVALID;
    int rc = check_pointcloud2(c, data, width, height, point_step, row_step, x_offset, y_offset, z_offset, pose);
    if (rc) return rc;
    if (!c->started) return PCF_DROPPED;
    ENTER(c);
    const uint32_t n = width * height;
    const bool dense_rows = row_step == (uint64_t)width * point_step || height <= 1;
    const size_t bytes = n ? (size_t)(height - 1) * row_step + (size_t)(width - 1) * point_step + x_offset + 12 : 0;
    if (dense_rows && x_offset == 0)      // plain [n, point_step] array: the fast paths (float4 / packed xyz bulk kernels) apply
        return push_host_cloud_impl(c, data, (size_t)n * point_step, 0, n, point_step / 4, nullptr, pose, nullptr, frame_idx, n, nullptr);
    RowLayout rl{dense_rows ? (n ? n : 1u) : width, point_step / 4, row_step / 4};
    return push_host_cloud_impl(c, data, bytes, x_offset / 4, n, 0, &rl, pose, nullptr, frame_idx, n, nullptr);
}

int pcf_add_points(pcf_ctx* c, const float* pts_host, uint32_t n, uint32_t stride, const float viewpoint[3], uint32_t frame_idx) {
    static const double identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (!viewpoint) return c ? fail(c, PCF_ERR_INVALID, "null viewpoint") : PCF_ERR_INVALID;
    return push_host_cloud(c, pts_host, n, stride, identity, viewpoint, frame_idx);
}

// ---- host staging (node.cpp:218-263): clip-and-pack on the host, so that only the clipped cloud crosses PCIe ----------
int pcf_stage_frame(pcf_ctx* c, const float* pts_host, uint32_t n, uint32_t stride, float* staged_xyz, uint32_t* n_staged) {
    if (!c || !staged_xyz || !n_staged || (!pts_host && n) || stride < 3) return c ? fail(c, PCF_ERR_INVALID, "bad staging arguments") : PCF_ERR_INVALID;
    StageJob j;
    j.data = reinterpret_cast<const uint8_t*>(pts_host);
    j.rows = 1; j.cols = n; j.row_step = 0; j.point_step = stride * 4; j.x_offset = 0;
    *n_staged = clip_pack(j, c->g.clip_lo, c->g.clip_hi, staged_xyz);
    return PCF_OK;
}

static int ensure_stager(pcf_ctx* c) {
    if (c->stager) return PCF_OK;
    int threads = c->cfg.stage_threads;
    if (const char* e = getenv("PCF_STAGE_THREADS")) if (atoi(e) > 0) threads = atoi(e);
    // packing is bound by memory bandwidth, not by cores: 3/4 of the hardware threads saturate it, and a full house makes the
    // in-order hand-over wait for descheduled stragglers (measured on a 16-vCPU host: 12 threads 6.7 G points/s, 16 threads 5.4 G)
    if (threads <= 0) threads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency() * 3u / 4u));
    Stager::Hooks h;
    const int dev = c->device;
    h.alloc_pinned = [](size_t bytes) { return pcf_host_alloc(bytes); };
    h.free_pinned = [](void* p) { pcf_host_free(p); };
    h.thread_init = [dev](int) { cudaSetDevice(dev); };
    h.slot_wait = [c](int s) { cudaEventSynchronize(c->slot_ev[s]); };
    h.push = [c](int s, const float* xyz, uint32_t n_staged, uint32_t n_offered, const double* pose, uint32_t frame_idx) {
        return push_host_cloud_impl(c, xyz, (size_t)n_staged * 12, 0, n_staged, 3, nullptr, pose, nullptr, frame_idx, n_offered, c->slot_ev[s]);
    };
    // raw lanes: measured harmful where the packers saturate the host's memory bandwidth (16 vCPUs, 12 packers: 6.0 -> 5.4 G
    // points/s) and useful where cores are scarce (8 GPUs on 32 vCPUs: 3 packers 8.5 G, 3 packers + 8 lanes 9.9 G, 1 packer +
    // 12 lanes 11.6 G points/s -- there the link, not the packers, should carry the clouds).  Auto: none with 8 or more
    // packers, else "upload mode" = 1 packer + 12 lanes (every lane holds one cloud in flight and waits for its own copy, so
    // the link is never oversubscribed; clouds that are not pinned are packed by the lanes as well).
    int raw_lanes = std::max(c->cfg.stage_raw_lanes, 0);
    if (c->cfg.stage_raw_lanes == 0 && !getenv("PCF_RAW_LANES")) {
        raw_lanes = threads >= 8 ? 0 : 12;
        if (threads < 8 && !getenv("PCF_STAGE_THREADS")) threads = 1;
    }
    if (const char* e = getenv("PCF_RAW_LANES")) raw_lanes = std::max(atoi(e), 0);
    h.raw_ok = [](const StageJob& j) {
        if (j.x_offset != 0 || (j.point_step != 16 && j.point_step != 12) || ((uintptr_t)j.data & 15u)) return false;
        if (j.rows > 1 && j.row_step != (uint64_t)j.cols * j.point_step) return false;
        if (j.point_step == 12 && ((uint64_t)j.rows * j.cols) % 4) return false;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, j.data) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;                   // page-locked: the copy engine can read it asynchronously
    };
    h.push_raw = [c](int lane, const StageJob& j) {
        const uint32_t n = j.rows * j.cols;
        return push_host_cloud_impl(c, j.data, (size_t)n * j.point_step, 0, n, j.point_step / 4, nullptr, j.pose, nullptr, j.frame_idx, n, c->raw_ev[lane]);
    };
    h.raw_wait = [c](int lane) { cudaEventSynchronize(c->raw_ev[lane]); };
    c->slot_ev.assign((size_t)(threads + raw_lanes) * 2, nullptr);
    for (cudaEvent_t& e : c->slot_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->raw_ev.assign((size_t)raw_lanes, nullptr);
    for (cudaEvent_t& e : c->raw_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->stager.reset(new Stager(threads, raw_lanes, c->g.clip_lo, c->g.clip_hi, std::move(h)));
    return PCF_OK;
}

static int submit_job(pcf_ctx* c, StageJob& j, const double pose[16], uint32_t frame_idx) {
    if (!c->started) return PCF_DROPPED;                     // node.cpp:329-331
    CU(cudaSetDevice(c->device));
    int rc = ensure_stager(c);
    if (rc) return rc;
    memcpy(j.pose, pose, sizeof j.pose);
    j.frame_idx = frame_idx;
    c->stager->submit(j);                                    // node.cpp:345-347
    return PCF_OK;
}

int pcf_submit_frame(pcf_ctx* c, const float* pts_host, uint32_t n, uint32_t stride, const double pose[16], uint32_t frame_idx) {
    if (!c) return PCF_ERR_INVALID;
    if ((!pts_host && n) || !pose || stride < 3) return fail(c, PCF_ERR_INVALID, "bad frame arguments");
    StageJob j;
    j.data = reinterpret_cast<const uint8_t*>(pts_host);
    j.rows = 1; j.cols = n; j.row_step = 0; j.point_step = stride * 4; j.x_offset = 0;
    return submit_job(c, j, pose, frame_idx);
}

int pcf_submit_pointcloud2(pcf_ctx* c, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step, uint32_t row_step,
                           uint32_t x_offset, uint32_t y_offset, uint32_t z_offset, const double pose[16], uint32_t frame_idx) {
    if (!c) return PCF_ERR_INVALID;
    int rc = check_pointcloud2(c, data, width, height, point_step, row_step, x_offset, y_offset, z_offset, pose);
    if (rc) return rc;
    StageJob j;
    j.data = data;
    j.rows = height; j.cols = width; j.row_step = row_step; j.point_step = point_step; j.x_offset = x_offset;
    return submit_job(c, j, pose, frame_idx);
}

int pcf_drain(pcf_ctx* c) {
    if (!c) return PCF_ERR_INVALID;
    return drain_staged(c);
}
int pcf_staged_count(pcf_ctx* c, uint64_t* n) {
    if (!c || !n) return PCF_ERR_INVALID;
    *n = c->stager ? c->stager->staged() : 0;
    return PCF_OK;
}
int pcf_wait_staged(pcf_ctx* c, uint64_t n) {
    if (!c) return PCF_ERR_INVALID;
    if (c->stager) {
        c->stager->wait_staged(n);
        const uint64_t raw = c->stager->raw_pushed();
        if (raw != c->raw_seen) { CU(cudaSetDevice(c->device)); CU(cudaStreamSynchronize(c->copy_stream)); c->raw_seen = raw; }
    }
    return PCF_OK;
}

int pcf_upload_ticket(pcf_ctx* c, uint64_t* ticket) {
    if (!c || !ticket) return PCF_ERR_INVALID;
    *ticket = c->uploads;
    return PCF_OK;
}
int pcf_wait_upload(pcf_ctx* c, uint64_t ticket) {
    if (!c) return PCF_ERR_INVALID;
    if (ticket == 0 || ticket > c->uploads) return PCF_OK;
    // copies complete in order on the copy stream: if the ticket's event has been recycled by a newer upload,
    // waiting for that newer one is a (stronger) valid wait
    CU(cudaEventSynchronize(c->ev_upload[ticket % kTickets]));
    return PCF_OK;
}

void* pcf_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void pcf_host_free(void* p) { if (p) cudaFreeHost(p); }

int pcf_push_frames_device(pcf_ctx* c, const float* pts_dev, uint32_t n_frames, uint32_t n_per_frame, uint32_t stride,
                           const double* poses, uint32_t first_frame_idx) {
    if (!c) return PCF_ERR_INVALID;
    if (!c->started) return PCF_DROPPED;
    if (!n_frames) return PCF_OK;
    if (!pts_dev || !poses || stride < 3) return fail(c, PCF_ERR_INVALID, "bad frame arguments");
    ENTER(c);
    int rc = check_frame_idx(c, first_frame_idx, n_frames);
    if (rc) return rc;
    if ((rc = flush_holders(c))) return rc;
    uint32_t chunks = div_up(n_per_frame, kWChunk);
    if ((uint64_t)c->n_chunks + (uint64_t)chunks * n_frames > kMaxChunks) return fail(c, PCF_ERR_CAPACITY, "point log limit reached");
    if ((rc = ensure_log(c, c->n_chunks + chunks * n_frames))) return rc;
    if (chunks) {
        for (uint32_t f0 = 0; f0 < n_frames; f0 += kMaxBatch) {
            uint32_t nf = std::min<uint32_t>(kMaxBatch, n_frames - f0);
            rc = launch_ingest(c, pts_dev + (size_t)f0 * n_per_frame * stride, (uint64_t)n_per_frame * stride, n_per_frame, nf, stride,
                               poses + (size_t)f0 * 16, first_frame_idx + f0);
            if (rc) return rc;
        }
    }
    c->last_frame_idx = (int64_t)first_frame_idx + n_frames - 1;
    c->occ_dirty = true;
    c->sorted_valid = false;
    c->stats.frames_pushed += n_frames;
    c->stats.points_offered += (uint64_t)n_frames * n_per_frame;
    return PCF_OK;
}

int pcf_sync(pcf_ctx* c) {
    if (!c) return PCF_ERR_INVALID;
    ENTER(c);
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    return PCF_OK;
}

int pcf_count_kept(pcf_ctx* c, uint64_t* kept) {
    if (!c || !kept) return PCF_ERR_INVALID;
    ENTER(c);
    CU(cudaStreamSynchronize(c->copy_stream));
    uint32_t P = 0;
    if (c->n_chunks) {
        int rc = reserve(c, c->tmpB, (size_t)c->n_chunks * 4);
        if (rc) return rc;
        uint32_t* tot = (uint32_t*)c->total_dev.p;
        if ((rc = scan_u32(c, c->chunk_count, (uint32_t*)c->tmpB.p, c->n_chunks, tot))) return rc;
        if ((rc = read_total(c, tot, &P))) return rc;
    } else {
        CU(cudaStreamSynchronize(c->stream));
    }
    c->stats.points_kept = P;
    *kept = P;
    return PCF_OK;
}

// One update pass in two halves.  update_local: candidates (occupied, no normal yet, inside this context's x-slab) -> 125-probe
// scan + PCA normal -> the new records compacted into upd_cell / upd_nrm (x-major).  update_commit: append records (the own
// ones, or the ones gathered from every rank in slab order = x-major order) as this pass and mark the voxels.
static int update_local_impl(pcf_ctx* c) {
    int rc = flush_holders(c);
    if (rc) return rc;
    if ((rc = build_occupancy(c))) return rc;
    uint32_t mark = c->n_chunks * (uint32_t)kWChunk;   // n_chunks <= 2^24 -> fits (2^32 wraps only at the hard limit)
    if (c->n_chunks >= kMaxChunks) mark = 0xFFFFFFFFu;
    c->upd_mark = mark;
    c->upd_n = 0;
    c->upd_open = true;
    uint32_t n_cand = 0;
    uint32_t* tot = (uint32_t*)c->total_dev.p;
    if (!c->n_vox) return PCF_OK;
    if ((rc = reserve(c, c->tmpA, (size_t)c->n_words * 4))) return rc;
    if ((rc = reserve(c, c->tmpB, (size_t)c->n_words * 4))) return rc;
    uint32_t* cnt = (uint32_t*)c->tmpA.p;
    uint32_t* off = (uint32_t*)c->tmpB.p;
    const uint64_t plane = (uint64_t)c->g.plane_cells;
    const uint64_t cell_lo = c->slab_hi < 0 ? 0 : (uint64_t)c->slab_lo * plane;
    const uint64_t cell_hi = c->slab_hi < 0 ? c->g.cells : std::min<uint64_t>(c->g.cells, (uint64_t)c->slab_hi * plane);
    LAUNCH(c, k_cand_count, div_up(c->n_words, kBlock), kBlock, c->occ_bits, c->nrm_bits, c->n_words, cell_lo, cell_hi, cnt);
    if ((rc = scan_u32(c, cnt, off, c->n_words, tot))) return rc;
    if ((rc = read_total(c, tot, &n_cand))) return rc;
    if (!n_cand) return PCF_OK;
    if ((rc = reserve(c, c->cand, (size_t)n_cand * 4))) return rc;
    if ((rc = reserve(c, c->tmpC, (size_t)n_cand * 16))) return rc;
    if ((rc = reserve(c, c->tmpD, (size_t)n_cand * 8))) return rc;
    uint32_t* cand = (uint32_t*)c->cand.p;
    float4* tnrm = (float4*)c->tmpC.p;
    uint32_t* flag = (uint32_t*)c->tmpD.p;
    uint32_t* foff = flag + n_cand;
    LAUNCH(c, k_cand_list, div_up(c->n_words, kBlock), kBlock, c->occ_bits, c->nrm_bits, c->n_words, cell_lo, cell_hi, off, cand);
    LAUNCH(c, k_normals, div_up(n_cand, kBlock), kBlock, cand, n_cand, c->g, c->occ_bits, c->first_frame, c->vp_table, tnrm, flag);
    if ((rc = scan_u32(c, flag, foff, n_cand, tot))) return rc;
    uint32_t n_new = 0;
    if ((rc = read_total(c, tot, &n_new))) return rc;
    if (n_new) {
        if ((rc = reserve(c, c->upd_cell, (size_t)n_new * 4))) return rc;
        if ((rc = reserve(c, c->upd_nrm, (size_t)n_new * 16))) return rc;
        LAUNCH(c, k_compact_normals, div_up(n_cand, kBlock), kBlock, cand, n_cand, tnrm, flag, foff, (uint32_t*)c->upd_cell.p, (float4*)c->upd_nrm.p);
    }
    c->upd_n = n_new;
    CU(cudaGetLastError());
    return PCF_OK;
}

static int update_commit_impl(pcf_ctx* c, const uint32_t* cells_dev, const float4* nrm_dev, uint32_t n) {
    if (!c->upd_open) return fail(c, PCF_ERR_INVALID, "pcf_update_commit without pcf_update_local");
    c->upd_open = false;
    c->marks.push_back(c->upd_mark);
    c->stats.update_passes++;
    if (n) {
        int rc;
        size_t need = (size_t)c->n_normals + n;
        if ((rc = reserve(c, c->n_cell, need * 4, true))) return rc;
        if ((rc = reserve(c, c->n_nrm, need * 16, true))) return rc;
        if ((rc = reserve(c, c->n_mark, need * 4, true))) return rc;
        LAUNCH(c, k_commit_normals, div_up(n, kBlock), kBlock, cells_dev, nrm_dev, n, c->n_normals, c->upd_mark, (uint32_t*)c->n_cell.p,
               (float4*)c->n_nrm.p, (uint32_t*)c->n_mark.p, c->nrm_bits);
        CU(cudaGetLastError());
        c->pending_holder.emplace_back(c->n_normals, c->n_normals + n);
        c->n_normals += n;
    }
    c->stats.normals_found = c->n_normals;
    c->stats.occupied_voxels = c->n_vox;
    return PCF_OK;
}

int pcf_update(pcf_ctx* c) {
    if (!c) return PCF_ERR_INVALID;
    ENTER(c);
    CU(cudaEventRecord(c->ev_a, c->stream));
    int rc = update_local_impl(c);
    if (rc) return rc;
    if ((rc = update_commit_impl(c, (const uint32_t*)c->upd_cell.p, (const float4*)c->upd_nrm.p, c->upd_n))) return rc;
    CU(cudaEventRecord(c->ev_b, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventElapsedTime(&c->t_update, c->ev_a, c->ev_b));
    return PCF_OK;
}

int pcf_update_local(pcf_ctx* c, void** cells_dev, void** normals_dev, uint32_t* n_new) {
    if (!c || !cells_dev || !normals_dev || !n_new) return PCF_ERR_INVALID;
    ENTER(c);
    if (c->n_chunks != c->merged_chunks && c->merged_chunks)
        return fail(c, PCF_ERR_INVALID, "pcf_update_local: this rank's round has not been exchanged (pcf_round_export / pcf_round_install first)");
    int rc = update_local_impl(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    *cells_dev = c->upd_cell.p;
    *normals_dev = c->upd_nrm.p;
    *n_new = c->upd_n;
    return PCF_OK;
}

int pcf_update_commit(pcf_ctx* c, const void* cells_dev, const void* normals_dev, uint32_t n) {
    if (!c || (n && (!cells_dev || !normals_dev))) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    int rc = update_commit_impl(c, (const uint32_t*)cells_dev, (const float4*)normals_dev, n);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return PCF_OK;
}

// ---- interleaved schedules across ranks: replicated grid state ---------------------------------------------------------------
// Between two update passes the frames are split over the ranks; at the update point every rank contributes the records of
// its frames (pcf_round_export), the caller all-gathers them in rank order (= frame order = arrival order) and every rank
// installs the SAME merged records (pcf_round_install): log, first-frame grid and occupancy bitmap are then identical on
// every rank and identical, record for record, to what one GPU fed every frame would hold.  Normal estimation, scoring and
// extraction are sharded by x-slab (pcf_set_slab + pcf_update_local / pcf_update_commit + pcf_extract).
int pcf_round_export(pcf_ctx* c, void** records_dev, uint64_t* n_records) {
    if (!c || !records_dev || !n_records) return PCF_ERR_INVALID;
    ENTER(c);
    CU(cudaStreamSynchronize(c->copy_stream));
    *records_dev = nullptr;
    *n_records = 0;
    const uint32_t first = c->merged_chunks, nloc = c->n_chunks - first;
    uint32_t P = 0;
    int rc;
    if (nloc) {
        if ((rc = reserve(c, c->tmpB, (size_t)nloc * 4))) return rc;
        uint32_t* off = (uint32_t*)c->tmpB.p;
        uint32_t* tot = (uint32_t*)c->total_dev.p;
        if ((rc = scan_u32(c, c->chunk_count + first, off, nloc, tot))) return rc;
        if ((rc = read_total(c, tot, &P))) return rc;
        if ((rc = reserve(c, c->dense_log, std::max<size_t>((size_t)P, 1) * sizeof(float4)))) return rc;
        LAUNCH(c, k_round_export, div_up(nloc, kWarps), kBlock, c->log, c->chunk_count, c->chunk_frame, off, first, c->n_chunks, (float4*)c->dense_log.p);
        CU(cudaGetLastError());
    } else if ((rc = reserve(c, c->dense_log, sizeof(float4)))) {
        return rc;
    }
    CU(cudaStreamSynchronize(c->stream));
    *records_dev = c->dense_log.p;
    *n_records = P;
    return PCF_OK;
}

int pcf_round_install(pcf_ctx* c, const void* records_dev, uint64_t n) {
    if (!c || (!records_dev && n)) return PCF_ERR_INVALID;
    ENTER(c);
    const uint32_t chunks = div_up(n, kWChunk);
    if ((uint64_t)c->merged_chunks + chunks > kMaxChunks) return fail(c, PCF_ERR_CAPACITY, "point log limit reached");
    // holders of the previous pass are registered against the occupancy AS OF that pass (OG.hpp:443-449): a rank that pushed no
    // frame of its own in this round has not done it yet, and the records installed below change the occupancy
    int rc = flush_holders(c);
    if (rc) return rc;
    c->n_chunks = c->merged_chunks;          // this rank's own chunks of the round are superseded by the merged records (own ones included)
    if ((rc = ensure_log(c, std::max<uint32_t>(c->merged_chunks + chunks, 1)))) return rc;
    if (n) {
        LAUNCH(c, k_install_records, div_up(n, kBlock), kBlock, (const float4*)records_dev, n, c->g, c->first_frame, c->occ_bits,
               c->log + (size_t)c->merged_chunks * kWChunk);
        LAUNCH(c, k_chunk_counts_dense, div_up(chunks, kBlock), kBlock, c->chunk_count + c->merged_chunks, chunks, n);
        CU(cudaGetLastError());
    }
    c->merged_chunks += chunks;
    c->n_chunks = c->merged_chunks;
    c->log_installed = true;
    c->occ_dirty = true;
    c->sorted_valid = false;
    CU(cudaStreamSynchronize(c->stream));
    return PCF_OK;
}

int pcf_extract(pcf_ctx* c, pcf_result* out) {
    if (!c) return PCF_ERR_INVALID;
    ENTER(c);
    return extract_impl(c, 0, out);
}

int pcf_extract_hq(pcf_ctx* c, double threshold, pcf_result* out) {
    if (!c) return PCF_ERR_INVALID;
    ENTER(c);
    // downloadHQ skips `count < threshold` (OG.hpp:561): for an int count that is count >= ceil(threshold)
    return extract_impl(c, (int32_t)std::ceil(threshold), out);
}

int pcf_clear(pcf_ctx* c) {
    if (!c) return PCF_ERR_INVALID;
    ENTER(c);
    CU(cudaStreamSynchronize(c->copy_stream));
    int rc = reset_grid_state(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return PCF_OK;
}

int pcf_process(pcf_ctx* c, const char* cloud_path, const char* meta_path) {
    if (!c) return PCF_ERR_INVALID;
    int rc = pcf_sync(c);                 // node.cpp:380-394: wait until both queues are drained
    if (rc) return rc;
    pcf_result r;
    if ((rc = pcf_extract(c, &r))) return rc;
    if ((rc = pcf_write_result(&r, cloud_path, meta_path))) return fail(c, rc, "could not write %s / %s", cloud_path ? cloud_path : "-", meta_path ? meta_path : "-");
    return pcf_clear(c);                  // node.cpp:438
}

// Pre-size the scratch of pcf_update / pcf_extract for up to `max_points` kept points and `max_voxels` occupied voxels, so that
// the first process() of a scan does not pay for cudaMalloc (a live node calls process() once per scan: 50-500 ms of
// allocation on the first call otherwise).  Buffers still grow on demand beyond these sizes.
int pcf_reserve_process(pcf_ctx* c, uint64_t max_points, uint64_t max_voxels) {
    if (!c) return PCF_ERR_INVALID;
    ENTER(c);
    if (max_points >= 0xFFFFFFFFull || max_voxels >= 0xFFFFFFFFull) return fail(c, PCF_ERR_INVALID, "pcf_reserve_process: sizes must fit 32 bits");
    const size_t P = (size_t)max_points, V = (size_t)max_voxels;
    const size_t tiles = P / kChunk + 257, chunks = std::max<size_t>(c->cap_chunks, P / 64 + 1);
    struct { DevBuf* b; size_t bytes; } want[] = {
        {&c->keysA, P * 4}, {&c->keysB, P * 4}, {&c->valsA, P * 4}, {&c->valsB, P * 4}, {&c->sorted, P * sizeof(float4)},
        {&c->hist, 256 * std::max(tiles, chunks / kWarps + 1) * 4}, {&c->sort_tab, tiles * sizeof(SortTile) + (257 + 258 + 257) * 4},
        {&c->tmpA, (size_t)c->n_words * 4}, {&c->tmpB, std::max((size_t)c->n_words, chunks) * 4}, {&c->tmpC, V * 16}, {&c->tmpD, V * 8},
        {&c->cand, V * 4}, {&c->upd_cell, V * 4}, {&c->upd_nrm, V * 16}, {&c->uv_off, (V + 1) * 4}, {&c->uv_cell, (V + 1) * 4}, {&c->nidx, (V + 1) * 4},
        {&c->n_cell, V * 4}, {&c->n_nrm, V * 16}, {&c->n_mark, V * 4}, {&c->sc_a, V * 16}, {&c->sc_b, V * 16}, {&c->sc_c, V * 4},
        {&c->sc_keys, V * 4}, {&c->sc_ids, V * 4}, {&c->sc_okeys, V * 4}, {&c->sc_order, V * 4}, {&c->sc_tab, (V / kChunk + 1) * sizeof(SortTile) + 16},
        {&c->flags, V * 4}, {&c->slots, V * 4}, {&c->res_dev, V * 64 + 8 * 256}, {&c->scan1, (std::max((size_t)c->n_words, P) / kChunk + 1) * 256 * 4},
    };
    for (auto& w : want) {
        const bool keep = w.b == &c->n_cell || w.b == &c->n_nrm || w.b == &c->n_mark;     // these hold state across calls
        int rc = reserve(c, *w.b, w.bytes, keep);
        if (rc) return rc;
    }
    int rc = ensure_pinned(c, &c->res_host, &c->res_host_cap, V * 64 + 8 * 256);
    return rc;
}

int pcf_dump_state(pcf_ctx* c, pcf_state* out) {
    if (!c || !out) return PCF_ERR_INVALID;
    memset(out, 0, sizeof *out);
    if (c->slab_hi >= 0) return fail(c, PCF_ERR_INVALID, "pcf_dump_state needs the whole grid (an x-slab is set)");
    ENTER(c);
    int rc = run_scoring(c);
    if (rc) return rc;
    size_t n = c->n_vox;
    size_t o_hash = 0, o_len = align256(n * 8), o_cnt = align256(o_len + n * 4), o_nrm = align256(o_cnt + n * 4),
           o_vp = align256(o_nrm + n * 12), o_nf = align256(o_vp + n * 12), total = align256(o_nf + n);
    if ((rc = reserve(c, c->res_dev, total))) return rc;
    if ((rc = ensure_pinned(c, &c->st_host, &c->st_host_cap, total))) return rc;
    char* d = (char*)c->res_dev.p;
    if (n) {
        StateDev s{(uint64_t*)(d + o_hash), (int32_t*)(d + o_len), (uint8_t*)(d + o_nf), (int32_t*)(d + o_cnt), (float*)(d + o_nrm), (float*)(d + o_vp)};
        LAUNCH(c, k_dump_state, div_up(n, kBlock), kBlock, (const uint32_t*)c->uv_cell.p, (const uint32_t*)c->uv_off.p, (uint32_t)n,
               (const uint32_t*)c->nidx.p, (const uint32_t*)c->n_mark.p, (const float4*)c->n_nrm.p, (const float4*)c->sc_a.p,
               (const float4*)c->sorted.p, c->first_frame, c->vp_table, c->g, s);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(c->st_host, d, total, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += total;
    }
    CU(cudaStreamSynchronize(c->stream));
    char* h = (char*)c->st_host;
    out->n = n;
    out->hash = (const uint64_t*)(h + o_hash);
    out->buffer_len = (const int32_t*)(h + o_len);
    out->normal_found = (const uint8_t*)(h + o_nf);
    out->count = (const int32_t*)(h + o_cnt);
    out->normal = (const float*)(h + o_nrm);
    out->viewpoint = (const float*)(h + o_vp);
    return PCF_OK;
}

int pcf_get_stats(pcf_ctx* c, pcf_stats* out) {
    if (!c || !out) return PCF_ERR_INVALID;
    c->stats.staged_dropped = c->staged_dropped;
    *out = c->stats;
    return PCF_OK;
}
int pcf_reset_stats(pcf_ctx* c) {
    if (!c) return PCF_ERR_INVALID;
    c->stats = pcf_stats{};
    c->staged_dropped = 0;
    return PCF_OK;
}
int pcf_last_timings(pcf_ctx* c, float* update_ms, float* extract_device_ms, float* extract_d2h_ms) {
    if (!c) return PCF_ERR_INVALID;
    if (update_ms) *update_ms = c->t_update;
    if (extract_device_ms) *extract_device_ms = c->t_extract_dev;
    if (extract_d2h_ms) *extract_d2h_ms = c->t_extract_d2h;
    return PCF_OK;
}
void* pcf_stream(pcf_ctx* c) { return c ? (void*)c->stream : nullptr; }

// ---- multi-GPU hooks ---------------------------------------------------------------------------------
int pcf_grid_buffer(pcf_ctx* c, void** first_frame_dev, uint64_t* n_cells) {
    if (!c || !first_frame_dev || !n_cells) return PCF_ERR_INVALID;
    ENTER(c);
    *first_frame_dev = c->first_frame;
    *n_cells = c->g.phys_cells;      // bricked physical layout: identical on every rank, so an elementwise reduce is still valid
    c->occ_dirty = true;        // the caller is about to reduce into it: the occupancy bitmap must be rebuilt from the grid
    c->occ_from_grid = true;
    c->sorted_valid = false;
    return PCF_OK;
}
int pcf_viewpoint_table(pcf_ctx* c, void** vp_dev, uint32_t* max_frames) {
    if (!c || !vp_dev || !max_frames) return PCF_ERR_INVALID;
    *vp_dev = c->vp_table;
    *max_frames = c->cfg.max_frames;
    return PCF_OK;
}
int pcf_log_compact(pcf_ctx* c, void** log_dev, uint64_t* n_points) {
    if (!c || !log_dev || !n_points) return PCF_ERR_INVALID;
    ENTER(c);
    CU(cudaStreamSynchronize(c->copy_stream));
    *log_dev = nullptr;
    *n_points = 0;
    uint32_t P = 0;
    if (c->n_chunks) {
        int rc = reserve(c, c->tmpB, (size_t)c->n_chunks * 4);
        if (rc) return rc;
        uint32_t* off = (uint32_t*)c->tmpB.p;
        uint32_t* tot = (uint32_t*)c->total_dev.p;
        if ((rc = scan_u32(c, c->chunk_count, off, c->n_chunks, tot))) return rc;
        if ((rc = read_total(c, tot, &P))) return rc;
        if ((rc = reserve(c, c->dense_log, std::max<size_t>((size_t)P, 1) * sizeof(float4)))) return rc;
        LAUNCH(c, k_log_compact, div_up(c->n_chunks, kWarps), kBlock, c->log, c->chunk_count, off, c->n_chunks, (float4*)c->dense_log.p);
        CU(cudaGetLastError());
    } else {
        int rc = reserve(c, c->dense_log, sizeof(float4));
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(c->stream));
    *log_dev = c->dense_log.p;
    *n_points = P;
    return PCF_OK;
}

int pcf_set_slab(pcf_ctx* c, int32_t x_lo, int32_t x_hi) {
    if (!c) return PCF_ERR_INVALID;
    if (x_hi >= 0 && (x_lo < 0 || x_lo > x_hi || (uint32_t)x_hi > c->g.n1[0])) return fail(c, PCF_ERR_INVALID, "bad slab [%d,%d)", x_lo, x_hi);
    c->slab_lo = x_lo;
    c->slab_hi = x_hi;
    return PCF_OK;
}

int pcf_log_replace(pcf_ctx* c, const void* log_dev, uint64_t n_points) {
    if (!c || (!log_dev && n_points)) return PCF_ERR_INVALID;
    if (n_points >= 0xFFFFFFFFull) return fail(c, PCF_ERR_CAPACITY, "merged log too large");
    if (c->n_normals) return fail(c, PCF_ERR_INVALID, "pcf_log_replace after pcf_update: sharded merge of interleaved schedules is not supported");
    ENTER(c);
    c->log_installed = true;
    CU(cudaStreamSynchronize(c->copy_stream));
    const float4* in = (const float4*)log_dev;
    // keep the records of this context's x-slab plus the reach of the +-K walk (OG.hpp:403-405): walk_k cells
    const uint64_t plane = (uint64_t)c->g.plane_cells;
    uint64_t cell_lo = 0, cell_hi = c->g.cells;
    if (c->slab_hi >= 0) {
        int64_t lo = (int64_t)c->slab_lo - c->g.walk_k, hi = (int64_t)c->slab_hi + c->g.walk_k;
        cell_lo = (uint64_t)std::max<int64_t>(lo, 0) * plane;
        cell_hi = std::min<uint64_t>(c->g.cells, (uint64_t)std::max<int64_t>(hi, 0) * plane);
    }
    uint32_t kept = 0;
    int rc;
    c->n_chunks = 0;
    if (n_points) {
        if ((rc = reserve(c, c->tmpC, (size_t)n_points * 4))) return rc;
        if ((rc = reserve(c, c->tmpD, (size_t)n_points * 4))) return rc;
        uint32_t* flag = (uint32_t*)c->tmpC.p;
        uint32_t* pos = (uint32_t*)c->tmpD.p;
        uint32_t* tot = (uint32_t*)c->total_dev.p;
        LAUNCH(c, k_log_filter_flags, div_up(n_points, kBlock), kBlock, in, n_points, cell_lo, cell_hi, flag);
        if ((rc = scan_u32(c, flag, pos, n_points, tot))) return rc;
        if ((rc = read_total(c, tot, &kept))) return rc;
        uint32_t chunks = div_up(kept, kWChunk);
        if (chunks > kMaxChunks) return fail(c, PCF_ERR_CAPACITY, "point log limit reached");
        if ((rc = ensure_log(c, std::max<uint32_t>(chunks, 1)))) return rc;
        if (kept) {
            LAUNCH(c, k_log_install, div_up(n_points, kBlock), kBlock, in, n_points, flag, pos, c->log);
            LAUNCH(c, k_chunk_counts_dense, div_up(chunks, kBlock), kBlock, c->chunk_count, chunks, (uint64_t)kept);
        }
        CU(cudaGetLastError());
        c->n_chunks = chunks;
    }
    CU(cudaStreamSynchronize(c->stream));
    c->occ_dirty = true;
    c->sorted_valid = false;
    return PCF_OK;
}

int pcf_plane_counts(pcf_ctx* c, uint32_t* counts_host) {
    if (!c || !counts_host) return PCF_ERR_INVALID;
    ENTER(c);
    int rc = build_occupancy(c);
    if (rc) return rc;
    uint32_t np = c->g.n1[0];
    if ((rc = reserve(c, c->tmpA, ((size_t)np + 1) * 4))) return rc;
    LAUNCH(c, k_plane_counts, div_up(np + 1, kBlock), kBlock, c->occ_bits, c->occ_rank, np, (uint64_t)c->g.plane_cells, c->n_vox,
           (uint32_t*)c->tmpA.p);
    CU(cudaMemcpyAsync(counts_host, c->tmpA.p, ((size_t)np + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PCF_OK;
}

// ---- exchange v2: slab-routed records, compaction fused with the (peer) write ---------------------------------
// Two ways to drive it.  Device-resident (one process per GPU, sharded.py::merge_and_extract_v3): pcf_exchange_hist ->
// [all-reduce of the histogram in place] -> pcf_exchange_plan -> [all-gather of the row] -> pcf_exchange_scatter_async ->
// [stream-ordered barrier] -> pcf_install_records; nothing in it waits for the host except the one read-back of the
// gathered rows that sizes the receive buffers.  Host-driven (several contexts in one process: the C++ replay driver,
// the emulated-rank tests): pcf_plane_point_counts -> pcf_exchange_counts(bounds) -> pcf_exchange_scatter -> pcf_install_records.
static int launch_plane_hist(pcf_ctx* c) {
    const uint32_t np = c->g.n1[0];
    if ((size_t)np * 4 > 200 * 1024) return fail(c, PCF_ERR_INVALID, "plane histogram: %u x-planes exceed the shared-memory histogram (51200)", np);
    if ((size_t)np * 4 > 48 * 1024) CU(cudaFuncSetAttribute(k_plane_point_counts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)np * 4));
    int rc = reserve(c, c->ex_hist, (size_t)np * 8);
    if (rc) return rc;
    CU(cudaMemsetAsync(c->ex_hist.p, 0, (size_t)np * 8, c->stream));
    if (c->n_chunks) {
        uint32_t grid = std::min<uint32_t>(div_up(c->n_chunks, kWarps), (uint32_t)c->sm_count * 4);
        k_plane_point_counts<<<grid, kBlock, (size_t)np * 4, c->stream>>>(c->log, c->chunk_count, c->n_chunks, c->g.plane_cells, np,
                                                                           (unsigned long long*)c->ex_hist.p);
        c->stats.kernel_launches++;
        CU(cudaGetLastError());
    }
    return PCF_OK;
}

int pcf_plane_point_counts(pcf_ctx* c, uint32_t* counts_host) {
    if (!c || !counts_host) return PCF_ERR_INVALID;
    ENTER(c);
    CU(cudaStreamSynchronize(c->copy_stream));
    int rc = launch_plane_hist(c);
    if (rc) return rc;
    const uint32_t np = c->g.n1[0];
    std::vector<unsigned long long> h(np);
    CU(cudaMemcpyAsync(h.data(), c->ex_hist.p, (size_t)np * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < np; i++) counts_host[i] = (uint32_t)h[i];
    return PCF_OK;
}

int pcf_exchange_hist(pcf_ctx* c, void** hist_dev, uint32_t* n_planes) {
    if (!c || !hist_dev || !n_planes) return PCF_ERR_INVALID;
    ENTER(c);
    if (c->n_normals) return fail(c, PCF_ERR_INVALID, "exchange after pcf_update: sharded merge of interleaved schedules is not supported");
    if (c->log_installed) return fail(c, PCF_ERR_INVALID, "exchange after pcf_install_records / pcf_log_replace: the installed log has lost its per-chunk frame indices; pcf_clear first");
    int rc = launch_plane_hist(c);
    if (rc) return rc;
    *hist_dev = c->ex_hist.p;
    *n_planes = c->g.n1[0];
    return PCF_OK;
}

// counts of this rank's records per destination, from the plan in c->ex_plan -> c->ex_row (device): [R totals | R + 1 bounds]
static int exchange_count_row(pcf_ctx* c, int32_t n_ranks) {
    int rc;
    if ((rc = reserve(c, c->ex_row, (size_t)(2 * kMaxRanks + 1) * 8))) return rc;
    const size_t n = (size_t)n_ranks * c->n_chunks;
    if ((rc = reserve(c, c->tmpC, (n + 1) * 4))) return rc;
    if ((rc = reserve(c, c->tmpD, (n + 1) * 4))) return rc;
    uint32_t* cnt = (uint32_t*)c->tmpC.p;
    uint32_t* off = (uint32_t*)c->tmpD.p;
    if (c->n_chunks) {
        LAUNCH(c, k_exchange_count, div_up(c->n_chunks, kWarps), kBlock, c->log, c->chunk_count, c->n_chunks, (const ExchangePlan*)c->ex_plan.p, cnt);
        CU(cudaMemsetAsync(cnt + n, 0, 4, c->stream));
        if ((rc = scan_u32(c, cnt, off, n + 1, nullptr))) return rc;       // off[n] = grand total
    }
    LAUNCH(c, k_exchange_row, 1, 32, (const uint32_t*)off, c->n_chunks, (const ExchangePlan*)c->ex_plan.p, (long long*)c->ex_row.p);
    CU(cudaGetLastError());
    c->ex_ranks = (uint32_t)n_ranks;
    c->plan_valid = true;
    return PCF_OK;
}

int pcf_exchange_plan(pcf_ctx* c, int32_t n_ranks, int32_t self, void** row_dev) {
    if (!c || !row_dev || n_ranks < 1 || n_ranks > kMaxRanks || self < 0 || self >= n_ranks)
        return c ? fail(c, PCF_ERR_INVALID, "bad exchange arguments (1..%d ranks)", kMaxRanks) : PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    if (!c->ex_hist.p) return fail(c, PCF_ERR_INVALID, "pcf_exchange_plan without pcf_exchange_hist");
    int rc = reserve(c, c->ex_plan, sizeof(ExchangePlan));
    if (rc) return rc;
    const int32_t halo = std::max(c->g.walk_k, 2);      // +-K walk (OG.hpp:403-405) and the 5x5x5 scan (OG.hpp:334)
    LAUNCH(c, k_slab_bounds, 1, 1024, (const unsigned long long*)c->ex_hist.p, c->g.n1[0], (uint32_t)n_ranks, (uint32_t)halo, c->g.plane_cells,
           (ExchangePlan*)c->ex_plan.p);
    if ((rc = exchange_count_row(c, n_ranks))) return rc;
    c->ex_self = self;
    *row_dev = c->ex_row.p;
    return PCF_OK;
}

int pcf_exchange_counts(pcf_ctx* c, const int32_t* bounds, int32_t n_ranks, uint64_t* counts_host) {
    if (!c || !bounds || !counts_host || n_ranks < 1 || n_ranks > kMaxRanks) return c ? fail(c, PCF_ERR_INVALID, "bad exchange arguments (1..%d ranks)", kMaxRanks) : PCF_ERR_INVALID;
    ENTER(c);
    CU(cudaStreamSynchronize(c->copy_stream));
    if (c->n_normals) return fail(c, PCF_ERR_INVALID, "exchange after pcf_update: sharded merge of interleaved schedules is not supported");
    if (c->log_installed) return fail(c, PCF_ERR_INVALID, "exchange after pcf_install_records / pcf_log_replace: the installed log has lost its per-chunk frame indices; pcf_clear first");
    const int32_t halo = std::max(c->g.walk_k, 2);      // +-K walk (OG.hpp:403-405) and the 5x5x5 scan (OG.hpp:334)
    ExchangePlan p{};
    p.n_ranks = (uint32_t)n_ranks;
    p.plane_cells = c->g.plane_cells;
    for (int d = 0; d < n_ranks; d++) {
        if (bounds[d] < 0 || bounds[d] > bounds[d + 1] || (uint32_t)bounds[d + 1] > c->g.n1[0]) return fail(c, PCF_ERR_INVALID, "bad slab bounds");
        bool empty = bounds[d] == bounds[d + 1];
        p.lo[d] = empty ? 0u : (uint32_t)std::max<int32_t>(bounds[d] - halo, 0);
        p.hi[d] = empty ? 0u : (uint32_t)std::min<int64_t>((int64_t)bounds[d + 1] + halo, c->g.n1[0]);
        p.bounds[d] = bounds[d];
    }
    p.bounds[n_ranks] = bounds[n_ranks];
    int rc = reserve(c, c->ex_plan, sizeof(ExchangePlan));
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->ex_plan.p, &p, sizeof p, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));               // `p` is a stack object
    if ((rc = exchange_count_row(c, n_ranks))) return rc;
    c->ex_self = -1;
    long long row[2 * kMaxRanks + 1];
    CU(cudaMemcpyAsync(row, c->ex_row.p, (size_t)(2 * n_ranks + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int d = 0; d < n_ranks; d++) counts_host[d] = (uint64_t)row[d];
    c->stats.d2h_bytes += 8 * (2 * n_ranks + 1);
    return PCF_OK;
}

static int exchange_scatter_impl(pcf_ctx* c, void* const* dst_bufs, const uint64_t* dst_offsets, bool sync) {
    if (!c || !dst_bufs || !dst_offsets) return PCF_ERR_INVALID;
    if (!c->plan_valid) return fail(c, PCF_ERR_INVALID, "pcf_exchange_scatter without pcf_exchange_counts / pcf_exchange_plan");
    CU(cudaSetDevice(c->device));
    ExchangeDst to{};
    for (uint32_t d = 0; d < c->ex_ranks; d++) to.dst[d] = (float4*)dst_bufs[d] + dst_offsets[d];
    if (c->n_chunks) {
        LAUNCH(c, k_exchange_scatter, div_up(c->n_chunks, kWarps), kBlock, c->log, c->chunk_count, c->chunk_frame, c->n_chunks,
               (const ExchangePlan*)c->ex_plan.p, to, (const uint32_t*)c->tmpD.p);
        CU(cudaGetLastError());
    }
    if (sync) CU(cudaStreamSynchronize(c->stream));     // host-driven callers: the barrier across ranks comes next
    c->plan_valid = false;
    return PCF_OK;
}
int pcf_exchange_scatter(pcf_ctx* c, void* const* dst_bufs, const uint64_t* dst_offsets) { return exchange_scatter_impl(c, dst_bufs, dst_offsets, true); }
int pcf_exchange_scatter_async(pcf_ctx* c, void* const* dst_bufs, const uint64_t* dst_offsets) { return exchange_scatter_impl(c, dst_bufs, dst_offsets, false); }

int pcf_recv_buffer(pcf_ctx* c, uint64_t n_records, void** dev_ptr) {
    if (!c || !dev_ptr) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    size_t bytes = std::max<size_t>((size_t)n_records, 1) * sizeof(float4);
    if (bytes > c->recv_cap) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->recv_buf) CU(cudaFree(c->recv_buf));
        c->recv_buf = nullptr; c->recv_cap = 0;
        size_t want = bytes + bytes / 8;
        CU(cudaMalloc(&c->recv_buf, want));
        c->recv_cap = want;
    }
    *dev_ptr = c->recv_buf;
    return PCF_OK;
}

int pcf_ipc_export(pcf_ctx* c, void* handle64) {
    if (!c || !handle64 || !c->recv_buf) return c ? fail(c, PCF_ERR_INVALID, "no receive buffer to export") : PCF_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->recv_buf));
    memcpy(handle64, &h, 64);
    return PCF_OK;
}
int pcf_ipc_open(pcf_ctx* c, const void* handle64, void** peer_ptr) {
    if (!c || !handle64 || !peer_ptr) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->ipc_opened.push_back(p);
    *peer_ptr = p;
    return PCF_OK;
}
int pcf_ipc_close_all(pcf_ctx* c) {
    if (!c) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    c->ipc_opened.clear();
    return PCF_OK;
}

int pcf_get_viewpoints(pcf_ctx* c, float* host4, uint32_t first, uint32_t count) {
    if (!c || !host4 || (uint64_t)first + count > c->cfg.max_frames) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(host4, c->vp_table + first, (size_t)count * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PCF_OK;
}
int pcf_set_viewpoints(pcf_ctx* c, const float* host4, uint32_t first, uint32_t count) {
    if (!c || !host4 || (uint64_t)first + count > c->cfg.max_frames) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->vp_table + first, host4, (size_t)count * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PCF_OK;
}
int pcf_enable_peer_access(pcf_ctx* c, int32_t peer_device) {
    if (!c) return PCF_ERR_INVALID;
    if (peer_device == c->device) return PCF_OK;
    CU(cudaSetDevice(c->device));
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, c->device, peer_device));
    if (!can) return fail(c, PCF_ERR_INVALID, "device %d cannot access device %d directly", c->device, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    if (e != cudaSuccess) return fail(c, PCF_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", peer_device, cudaGetErrorString(e));
    return PCF_OK;
}

int pcf_install_records(pcf_ctx* c, const void* records_dev, uint64_t n) {
    if (!c || (!records_dev && n)) return PCF_ERR_INVALID;
    if (n >= 0xFFFFFFFFull) return fail(c, PCF_ERR_CAPACITY, "merged log too large");
    if (c->n_normals) return fail(c, PCF_ERR_INVALID, "pcf_install_records after pcf_update: sharded merge of interleaved schedules is not supported");
    ENTER(c);
    c->log_installed = true;
    uint32_t chunks = div_up(n, kWChunk);
    if (chunks > kMaxChunks) return fail(c, PCF_ERR_CAPACITY, "point log limit reached");
    const bool keep_region = c->ex_self >= 0 && c->ex_plan.p;     // device-resident exchange: this rank's region is known
    if (keep_region) {
        // the own cells inside slab + halo stay (the routed records include them), the own cells outside are emptied
        if (c->n_chunks)
            LAUNCH(c, k_unmark_outside, div_up(c->n_chunks, kWarps), kBlock, c->log, c->chunk_count, c->n_chunks, c->g, (const ExchangePlan*)c->ex_plan.p,
                   (uint32_t)c->ex_self, c->first_frame, c->occ_bits);
    } else {
        CU(cudaStreamSynchronize(c->copy_stream));
        LAUNCH(c, k_fill_u32, 148 * 8, 512, c->first_frame, c->g.phys_cells, kEmpty);
        CU(cudaMemsetAsync(c->occ_bits, 0, (c->n_words + 2) * 4, c->stream));
    }
    c->n_chunks = 0;                       // the local log is superseded by the routed records (own ones included)
    int rc = ensure_log(c, std::max<uint32_t>(chunks, 1));
    if (rc) return rc;
    if (n) {
        LAUNCH(c, k_install_records, div_up(n, kBlock), kBlock, (const float4*)records_dev, n, c->g, c->first_frame, c->occ_bits, c->log);
        LAUNCH(c, k_chunk_counts_dense, div_up(chunks, kBlock), kBlock, c->chunk_count, chunks, n);
        CU(cudaGetLastError());
    }
    c->n_chunks = chunks;
    if (!keep_region) CU(cudaStreamSynchronize(c->stream));
    c->ex_self = -1;
    c->occ_dirty = true;
    c->sorted_valid = false;
    return PCF_OK;
}

// ---- known-answer hooks -------------------------------------------------------------------------------
int pcf_kat_transform_voxel(pcf_ctx* c, const float* pts_host, uint32_t n, uint32_t stride, const double pose[16],
                            float* world_xyz, int32_t* ijk, uint8_t* kept) {
    if (!c || !pts_host || !pose || stride < 3) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    float* d_in; float* d_w; int32_t* d_ijk; uint8_t* d_k;
    CU(cudaMalloc(&d_in, (size_t)n * stride * 4));
    CU(cudaMalloc(&d_w, (size_t)n * 12));
    CU(cudaMalloc(&d_ijk, (size_t)n * 12));
    CU(cudaMalloc(&d_k, (size_t)n));
    CU(cudaMemcpyAsync(d_in, pts_host, (size_t)n * stride * 4, cudaMemcpyHostToDevice, c->stream));
    PoseParam fd;
    for (int i = 0; i < 12; i++) fd.T[i] = pose[i];
    LAUNCH(c, k_kat_transform_voxel, div_up(n, 256), 256, d_in, n, stride, fd, c->g, d_w, d_ijk, d_k);
    CU(cudaMemcpyAsync(world_xyz, d_w, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(ijk, d_ijk, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(kept, d_k, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_in); cudaFree(d_w); cudaFree(d_ijk); cudaFree(d_k);
    return PCF_OK;
}
// the staging pool's clip-and-pack on its own (host code, no context and no GPU needed): `isa` 0 scalar, 1 AVX2, 2 AVX-512,
// -1 the one the pool uses on this CPU.  Returns the implementation that ran, or a negative status.
int pcf_kat_clip_pack(const uint8_t* data, uint32_t rows, uint32_t cols, uint32_t point_step, uint64_t row_step, uint32_t x_offset,
                      float clip_lo, float clip_hi, int32_t isa, float* out_xyz, uint32_t* n_out) {
    if (!data || !out_xyz || !n_out || point_step % 4 || x_offset % 4 || point_step < x_offset + 12) return PCF_ERR_INVALID;
    StageJob j;
    j.data = data; j.rows = rows; j.cols = cols; j.row_step = row_step; j.point_step = point_step; j.x_offset = x_offset;
    const int use = isa < 0 ? clip_pack_isa() : std::min<int>(isa, clip_pack_isa());
    *n_out = clip_pack_with(use, j, clip_lo, clip_hi, out_xyz);
    return use;
}
// x[i] / c[i] through the scoring kernel's shared-reciprocal division vs the compiler's div.rn.f32: number of results that
// differ in any bit (two NaNs count as equal) and the index of one of them
int pcf_kat_div(pcf_ctx* c, const float* x_host, const float* c_host, uint32_t n, uint32_t* mismatches, uint32_t* first_bad) {
    if (!c || !x_host || !c_host || !mismatches || !first_bad) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    float *dx, *dc; uint32_t* dm;
    CU(cudaMalloc(&dx, (size_t)n * 4 + 4));
    CU(cudaMalloc(&dc, (size_t)n * 4 + 4));
    CU(cudaMalloc(&dm, 8));
    CU(cudaMemsetAsync(dm, 0, 8, c->stream));
    CU(cudaMemcpyAsync(dx, x_host, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dc, c_host, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    if (n) LAUNCH(c, k_kat_div, div_up(n, 256), 256, dx, dc, n, dm, dm + 1);
    uint32_t h[2];
    CU(cudaMemcpyAsync(h, dm, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *mismatches = h[0]; *first_bad = h[1];
    cudaFree(dx); cudaFree(dc); cudaFree(dm);
    return PCF_OK;
}
int pcf_kat_normal(pcf_ctx* c, const float* xyz_host, uint32_t n_points, float* normal3) {
    if (!c || !xyz_host || !normal3) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    float *d_in, *d_out;
    CU(cudaMalloc(&d_in, (size_t)n_points * 12));
    CU(cudaMalloc(&d_out, 12));
    CU(cudaMemcpyAsync(d_in, xyz_host, (size_t)n_points * 12, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_kat_normal, 1, 32, d_in, n_points, d_out);
    CU(cudaMemcpyAsync(normal3, d_out, 12, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_in); cudaFree(d_out);
    return PCF_OK;
}
int pcf_kat_score(pcf_ctx* c, const float* xyz_host, uint32_t n_points, const float axis_pt[3], const float normal[3],
                  float* centroid3, float* sd3, float* mean_dist, float* sd_dist, int32_t* count) {
    if (!c || !xyz_host) return PCF_ERR_INVALID;
    CU(cudaSetDevice(c->device));
    float *d_in, *d_out; int32_t* d_cnt;
    CU(cudaMalloc(&d_in, (size_t)n_points * 12 + 4));
    CU(cudaMalloc(&d_out, 64));
    CU(cudaMalloc(&d_cnt, 4));
    CU(cudaMemcpyAsync(d_in, xyz_host, (size_t)n_points * 12, cudaMemcpyHostToDevice, c->stream));
    V3 a{axis_pt[0], axis_pt[1], axis_pt[2]}, nn{normal[0], normal[1], normal[2]};
    LAUNCH(c, k_kat_score, 1, 32, d_in, n_points, a, nn, c->g, d_out, d_cnt);
    float h[9];
    CU(cudaMemcpyAsync(h, d_out, 32, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(count, d_cnt, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(centroid3, h, 12); memcpy(sd3, h + 3, 12); *mean_dist = h[6]; *sd_dist = h[7];
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_cnt);
    return PCF_OK;
}

}  // extern "C"
