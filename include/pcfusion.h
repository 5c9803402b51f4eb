/* pcfusion.h -- C ABI of libpcfusion.so, the B200 (sm_100a) drop-in for the frame-integration and
 * process() hot path of the `pointcloud_fusion` node (REXJJ/high-fidelity-pointcloud-fusion).
 *
 * The reference has no FFI layer: its boundary is the public surface of `class OccupancyGrid`
 * (OG.hpp = pointcloud_fusion/pointcloud_fusion/include/utilities/OccupancyGrid.hpp:99-136) as driven by
 * `PointcloudFusion` (node.cpp = pointcloud_fusion/pointcloud_fusion/src/pointcloud_fusion_and_filter.cpp).
 * Every entry point below names the reference interface it replaces.  Plain pointers and sizes only.
 *
 * Threading: like the reference grid (serialised by grid_mtx_, node.cpp:291-296,305-321) one context must be
 * driven from one thread at a time.  pcf_push_* are asynchronous on the context's CUDA stream; pcf_sync,
 * pcf_update (results), pcf_extract, pcf_dump_state, pcf_process and pcf_clear synchronise.
 * Errors: every call returns PCF_OK (0) or a negative pcf_status; nothing throws across this boundary.
 */
#ifndef PCFUSION_H
#define PCFUSION_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcf_ctx pcf_ctx;

typedef enum {
    PCF_OK = 0,
    PCF_DROPPED = 1,          /* frame offered while stopped: ignored, like node.cpp:329-331 */
    PCF_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
    PCF_ERR_CUDA = -2,        /* CUDA runtime error (text in pcf_last_error) */
    PCF_ERR_CAPACITY = -3,    /* point log / frame table limit reached */
    PCF_ERR_IO = -4,          /* cloud / metadata file could not be written */
    PCF_ERR_NO_DEVICE = -5,   /* no usable CUDA device: the library has no CPU fallback */
    PCF_ERR_INTERNAL = -6     /* a device-side consistency check failed (text in pcf_last_error) */
} pcf_status;

/* Grid + stage parameters.  Replaces the compile-time constants and ctor calls of the node:
 *   box      setDimensions(xmin,xmax,ymin,ymax,zmin,zmax)  OG.hpp:604-612, node.cpp:162, launch:8
 *   res      setResolution(float,float,float)              OG.hpp:614-619, node.cpp:91,161 (float on purpose:
 *            dims = (int)((max-min)/double(res)) keeps the reference's truncation, OG.hpp:621-626)
 *   clip_*   kZmin / kZmax camera-frame depth clip          node.cpp:92-93,251
 *   k_neighbourhood  setK(2): 5x5x5 probe block             OG.hpp:138-149,334 (only 2 is supported, D9)
 *   walk_k   K of updateThicknessVectors<N,K>               node.cpp:311, OG.hpp:403
 *   min_neighbours   `total>20` gate                        OG.hpp:352
 *   cylinder_radius / ball_radius  kCylinderRadius / kBballRadius   OG.hpp:35-36
 */
typedef struct {
    double box[6];
    float res[3];
    double clip_zmin, clip_zmax;
    int32_t k_neighbourhood;
    int32_t walk_k;
    int32_t min_neighbours;
    double cylinder_radius;
    double ball_radius;
    int32_t device;              /* CUDA device ordinal */
    uint32_t max_frames;         /* size of the frame/viewpoint table (frame_idx < max_frames) */
    uint64_t log_capacity_hint;  /* initial point-log capacity in INPUT points; grows on demand */
    int32_t stage_threads;       /* host staging threads of pcf_submit_* (the reference has ONE addPoints thread,
                                    node.cpp:166,218); 0 = min(16, 3/4 of the hardware threads); env PCF_STAGE_THREADS overrides */
    int32_t stage_raw_lanes;     /* extra threads that upload pinned float4 / xyz clouds UNSTAGED while clouds pile up behind the
                                    packers, using the PCIe time the staged copies leave free.  0 = auto: none when there are 8 or
                                    more packers (they saturate the host's memory bandwidth; raw DMA traffic then costs more than
                                    it brings); with fewer, "upload mode": 1 packer + 12 lanes (core-starved hosts: the link
                                    carries the clouds).  < 0 = none; env PCF_RAW_LANES overrides */
} pcf_config;

/* Extraction output, structure of arrays, x-major voxel order = the reference's scan order
 * (downloadData, OG.hpp:463-480).  Host memory owned by the library, valid until the next
 * pcf_extract / pcf_process / pcf_clear / pcf_destroy on the same context. */
typedef struct {
    uint64_t n;
    const uint64_t* hash;     /* getHashId(x,y,z), OG.hpp:151-156 */
    const float* centroid;    /* n x 3, VoxelInfo::centroid (mean of projected in-cylinder points) */
    const float* normal;      /* n x 3 */
    const float* sd;          /* n x 3 */
    const float* mean_dist;   /* n */
    const float* sd_dist;     /* n */
    const int32_t* count;     /* n, points inside the 1 mm normal cylinder */
} pcf_result;

/* Per-voxel state for parity tests: every occupied cell (pad cells index==dim included), x-major. */
typedef struct {
    uint64_t n;
    const uint64_t* hash;
    const int32_t* buffer_len;    /* VoxelInfo::buffer.size(), OG.hpp:70,211 */
    const uint8_t* normal_found;  /* OG.hpp:72 */
    const int32_t* count;         /* OG.hpp:73 */
    const float* normal;          /* n x 3 (zeros where !normal_found) */
    const float* viewpoint;       /* n x 3, first inserting frame's viewpoint, OG.hpp:229 */
} pcf_state;

typedef struct {
    uint64_t frames_pushed;
    uint64_t points_offered;      /* input points incl. clipped / cropped ones */
    uint64_t points_kept;         /* passed z clip and box test (valid after pcf_sync) */
    uint64_t occupied_voxels;     /* valid after pcf_extract / pcf_dump_state */
    uint64_t normals_found;
    uint64_t kernel_launches;     /* library kernels launched since create / last pcf_reset_stats */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t update_passes;
    uint64_t staged_dropped;      /* clouds dropped by pcf_reset before a staging thread took them (node.cpp:356) */
} pcf_stats;

void pcf_default_config(pcf_config* cfg);       /* launch-file defaults: launch:4-8, node.cpp:91-93 */
int pcf_create(const pcf_config* cfg, pcf_ctx** out);   /* ctor + setResolution/setDimensions/setK/construct, node.cpp:146-169 */
void pcf_destroy(pcf_ctx* ctx);
const char* pcf_last_error(const pcf_ctx* ctx);  /* ctx may be NULL: error of the last failed pcf_create */
int pcf_dims(const pcf_ctx* ctx, int32_t dims[3]);      /* xdim_,ydim_,zdim_  OG.hpp:621-625 */

/* Service semantics (std_srvs/Trigger handlers). */
int pcf_start(pcf_ctx* ctx);    /* node.cpp:361-367 */
int pcf_stop(pcf_ctx* ctx);     /* node.cpp:369-375: queued frames still integrate */
int pcf_reset(pcf_ctx* ctx);    /* node.cpp:351-359: start_ = false + clouds_.clear(): closes the gate and drops the clouds submitted
                                   with pcf_submit_* that no staging thread has taken yet; clouds already being staged
                                   (the reference's clouds_processed_) still integrate; the grid is kept */

/* Frame integration = onReceivedPointCloud -> addPoints thread (z clip) -> updateStates thread
 * (transformPointCloud + OccupancyGrid::addPoints): node.cpp:327-349, 248-255, 288-296, OG.hpp:185-280.
 * pts: n points, `stride_floats` floats apart (>=3; 4 = float4 fast path), camera frame, host memory
 * (pinned for full speed).  pose: row-major 4x4 fusion<-camera (Eigen::Affine3d of node.cpp:338).
 * frame_idx must increase from call to call on one context.  Returns PCF_DROPPED while stopped. */
int pcf_push_frame(pcf_ctx* ctx, const float* pts_host, uint32_t n, uint32_t stride_floats,
                   const double pose[16], uint32_t frame_idx);
/* sensor_msgs/PointCloud2 front end = pointCloud2ToPclXYZRGBOMP (node.cpp:182-216) + the rest of pcf_push_frame: `data` is
 * msg.data, x/y/z_offset are fields[0..2].offset (consecutive float32).  All rows of an organized cloud are used (D6);
 * rgb is ignored, as it is downstream in the reference (OG.hpp:471-477 never reads it). */
int pcf_push_pointcloud2(pcf_ctx* ctx, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step,
                         uint32_t row_step, uint32_t x_offset, uint32_t y_offset, uint32_t z_offset, const double pose[16],
                         uint32_t frame_idx);
/* OccupancyGrid::addPoints<N>(cloud, viewpoint), OG.hpp:185-280, verbatim: the cloud is ALREADY in the fusion frame
 * (the caller did node.cpp:248-255,288-290 itself), no depth clip, no transform; `viewpoint` is the Eigen::Vector3f
 * argument.  This is the entry point the C++ drop-in class (include/pcfusion/OccupancyGrid.hpp) binds; pcf_push_frame
 * is the faster route that also moves the clip and the transform onto the GPU. */
int pcf_add_points(pcf_ctx* ctx, const float* pts_host, uint32_t n, uint32_t stride_floats, const float viewpoint[3],
                   uint32_t frame_idx);
/* ---- host staging = the reference's addPoints() thread (node.cpp:218-263: decode + camera-frame depth clip) -----------
 * pcf_submit_frame / pcf_submit_pointcloud2 are the asynchronous counterparts of pcf_push_frame / pcf_push_pointcloud2 and
 * the entry points a live bridge should use: the call only queues the cloud (node.cpp:345-347; it blocks while 4 x
 * stage_threads clouds are waiting).  A pool of staging threads walks each cloud once, drops the points outside
 * clip_zmin < z < clip_zmax (node.cpp:251; the same float thresholds the kernel uses, exactly equivalent to the
 * reference's double compares) and packs the survivors, in point order, as 12-byte xyz into pinned slots; slots are
 * uploaded and integrated strictly in submission order, so results are bit-identical to pcf_push_frame while only the
 * clipped cloud crosses PCIe.  `pts_host` / `data` may be pageable memory (a ROS message) and must stay valid until the
 * cloud has been staged: until pcf_drain, or any call that drains (pcf_sync, pcf_count_kept, pcf_update, pcf_extract,
 * pcf_process, pcf_clear, direct pcf_push_*).  Returns PCF_DROPPED while stopped.  Errors of the deferred integration
 * surface at the next draining call. */
int pcf_submit_frame(pcf_ctx* ctx, const float* pts_host, uint32_t n, uint32_t stride_floats, const double pose[16],
                     uint32_t frame_idx);
int pcf_submit_pointcloud2(pcf_ctx* ctx, const uint8_t* data, uint32_t width, uint32_t height, uint32_t point_step,
                           uint32_t row_step, uint32_t x_offset, uint32_t y_offset, uint32_t z_offset, const double pose[16],
                           uint32_t frame_idx);
int pcf_drain(pcf_ctx* ctx);    /* wait until every submitted cloud has been staged and handed to the GPU (not for the GPU) */
/* Source-buffer bookkeeping for callers that recycle their cloud memory: pcf_staged_count = clouds handed to the GPU by the
 * pool so far (in submission order; clouds dropped by pcf_reset are not counted); pcf_wait_staged(n) blocks until that count
 * reaches n or nothing is pending. */
int pcf_staged_count(pcf_ctx* ctx, uint64_t* n);
int pcf_wait_staged(pcf_ctx* ctx, uint64_t n);
/* The clip-and-pack step on its own, on the calling thread, for hosts that run their own staging threads:
 * staged_xyz (>= 3 * (n + 4) floats, pinned for full upload speed) receives the clipped points as packed xyz, padded
 * with NaN points to a multiple of 4; pass it to pcf_push_frame(ctx, staged_xyz, *n_staged, 3, pose, frame_idx). */
int pcf_stage_frame(pcf_ctx* ctx, const float* pts_host, uint32_t n, uint32_t stride_floats, float* staged_xyz,
                    uint32_t* n_staged);
/* Pinned (page-locked) host memory for staging clouds: what the reference's deques hold (node.cpp:130-143) lives
 * here so that pcf_push_frame's H2D copy runs at full PCIe speed and asynchronously. */
void* pcf_host_alloc(size_t bytes);
void pcf_host_free(void* p);
/* Upload tickets: pcf_push_frame / pcf_add_points return before the H2D copy of a pinned cloud has finished.
 * pcf_upload_ticket gives the ticket of the most recent push; pcf_wait_upload blocks until that push's copy (and
 * every earlier one) is done, i.e. until the staging buffer may be refilled. */
int pcf_upload_ticket(pcf_ctx* ctx, uint64_t* ticket);
int pcf_wait_upload(pcf_ctx* ctx, uint64_t ticket);
/* Same for clouds already resident in device memory: n_frames clouds of n_per_frame points each, back to
 * back, integrated by ONE launch (up to 256 frames per launch; longer batches are split).  poses: n_frames x 16
 * doubles (host).  Frames get indices first_frame_idx .. first_frame_idx+n_frames-1.  The kernel runs on the context's
 * own stream (pcf_stream): the clouds must be complete when this is called (synchronise the stream that produced them,
 * or make pcf_stream wait on it), and must stay valid until pcf_sync. */
int pcf_push_frames_device(pcf_ctx* ctx, const float* pts_dev, uint32_t n_frames, uint32_t n_per_frame,
                           uint32_t stride_floats, const double* poses, uint32_t first_frame_idx);
int pcf_sync(pcf_ctx* ctx);     /* wait for all queued integration work */
/* Drain, then read back how many points passed the clip + box test so far (the integration result summary:
 * what the node prints as "Pointcloud N states updated..", node.cpp:297). */
int pcf_count_kept(pcf_ctx* ctx, uint64_t* kept);

/* updateThicknessVectors<N,K>(): neighbour scan, PCA normal, +-K walk registration.  OG.hpp:311-454,
 * called by the cleanGrid thread every 5 s (node.cpp:301-325); here the schedule is explicit (D4). */
int pcf_update(pcf_ctx* ctx);

/* Scoring + x-major stream-compacted extraction; does not modify the grid.  OG.hpp:456-488 (scan part). */
int pcf_extract(pcf_ctx* ctx, pcf_result* out);
/* getFusedCloud(): drain, downloadData(cloud_path, meta_path) as ASCII PCD + CSV, clearVoxels().
 * node.cpp:377-440, OG.hpp:456-488, 167-183.  Paths may be NULL to skip a file. */
int pcf_process(pcf_ctx* ctx, const char* cloud_path, const char* meta_path);
int pcf_write_result(const pcf_result* res, const char* cloud_path, const char* meta_path);
/* Dead-code variants kept by the reference (OG.hpp:514-575, node.cpp:399-437): same scan with a predicate.
 * Returns in `out` only voxels with count >= threshold (downloadHQ). */
int pcf_extract_hq(pcf_ctx* ctx, double threshold, pcf_result* out);
int pcf_clear(pcf_ctx* ctx);    /* clearVoxels() + work lists, OG.hpp:167-183 (D5: full reset) */

/* Optional: pre-size the scratch of pcf_update / pcf_extract for up to max_points kept points and max_voxels occupied voxels,
 * so that the first process() of a scan does not pay for device allocations (everything still grows on demand). */
int pcf_reserve_process(pcf_ctx* ctx, uint64_t max_points, uint64_t max_voxels);
int pcf_dump_state(pcf_ctx* ctx, pcf_state* out);
int pcf_get_stats(pcf_ctx* ctx, pcf_stats* out);
int pcf_reset_stats(pcf_ctx* ctx);
/* Milliseconds spent on the device by the last pcf_update / pcf_extract (CUDA events on the ctx stream). */
int pcf_last_timings(pcf_ctx* ctx, float* update_ms, float* extract_device_ms, float* extract_d2h_ms);
/* CUDA stream of the context as a cudaStream_t (for callers that time with their own events). */
void* pcf_stream(pcf_ctx* ctx);

/* ---- multi-GPU merge hooks (one context per GPU; the exchange itself is the caller's, e.g. NCCL) --------
 * Frames are sharded over ranks in contiguous frame_idx blocks.  At process() the dense first-frame grid is
 * min-reduced across ranks, viewpoints are all-gathered, and each rank's point log is exchanged. */
int pcf_grid_buffer(pcf_ctx* ctx, void** first_frame_dev, uint64_t* n_cells);    /* uint32 per cell, 0x7FFFFFFF = empty; bricked physical order, the same on every rank with the same config: reduce it elementwise */
int pcf_viewpoint_table(pcf_ctx* ctx, void** vp_dev, uint32_t* max_frames);       /* float4 per frame_idx, w=1 when set */
int pcf_log_compact(pcf_ctx* ctx, void** log_dev, uint64_t* n_points);            /* float4 (x,y,z,cell) in arrival order */
/* x-slab [x_lo, x_hi) of voxels this context normal-estimates, scores and extracts (x_hi < 0: whole grid).  The
 * concatenation of the ranks' extractions in slab order is the reference's x-major order (OG.hpp:463-465). */
int pcf_set_slab(pcf_ctx* ctx, int32_t x_lo, int32_t x_hi);
/* Install a merged log (records of all ranks in arrival order); only the slab +- walk_k planes are kept. */
int pcf_log_replace(pcf_ctx* ctx, const void* log_dev, uint64_t n_points);
/* counts_host[x] = occupied voxels with plane index < x, x = 0 .. xdim+1 (for balanced slab boundaries). */
int pcf_plane_counts(pcf_ctx* ctx, uint32_t* counts_host);

/* ---- exchange v2: records routed to the owner of their x-slab, no dense grid on the wire -------------------------
 * 1. pcf_plane_point_counts on every rank, summed over ranks by the caller -> slab bounds b[0..R] balanced by points.
 * 2. pcf_exchange_counts(bounds): how many of this rank's records go to each destination (slab +- halo planes);
 *    the R x R count matrix (all-gathered by the caller) gives every (source, destination) offset.
 * 3. pcf_recv_buffer on every rank (sized by its column sum); peers map it with pcf_ipc_export / pcf_ipc_open (one
 *    process per GPU) or simply pass the pointer (several contexts in one process).
 * 4. pcf_exchange_scatter(dst_bufs, dst_offsets): ONE kernel compacts this rank's log per destination and writes
 *    the (x, y, z, frame_idx) records straight into the destinations' receive buffers (NVLink peer stores).
 * 5. barrier across ranks, then pcf_install_records(own receive buffer): recompute cells, rebuild the first-frame grid,
 *    install the log.  pcf_set_slab + pcf_update + pcf_extract follow as before. */
int pcf_plane_point_counts(pcf_ctx* ctx, uint32_t* counts_host /* xdim+1 entries */);
int pcf_exchange_counts(pcf_ctx* ctx, const int32_t* bounds /* n_ranks+1 */, int32_t n_ranks, uint64_t* counts_host /* n_ranks */);
int pcf_exchange_scatter(pcf_ctx* ctx, void* const* dst_bufs /* n_ranks device pointers */, const uint64_t* dst_offsets /* records */);
/* Device-resident variant (one process per GPU): nothing below waits for the host.
 *   pcf_exchange_hist   per-plane record histogram of this rank, uint64[n_planes] in device memory, computed on pcf_stream;
 *                       the caller all-reduces (SUM) it IN PLACE across ranks, stream-ordered after pcf_stream.
 *   pcf_exchange_plan   slab bounds from the reduced histogram (k_slab_bounds), routing counts, and the row this rank
 *                       contributes to the all-gather: int64[2 * n_ranks + 1] = [records to every destination | bounds]
 *                       in device memory.  The gathered n_ranks x n_ranks totals give every (source, destination) offset.
 *   pcf_exchange_scatter_async   as pcf_exchange_scatter without the trailing stream synchronisation; the caller orders the
 *                       peers' install after it with a stream-ordered collective.
 *   pcf_install_records after pcf_exchange_plan keeps the own cells inside slab + halo and empties the own cells outside
 *                       (one pass over the own log) instead of refilling the dense grid, and does not synchronise. */
int pcf_exchange_hist(pcf_ctx* ctx, void** hist_dev, uint32_t* n_planes);
int pcf_exchange_plan(pcf_ctx* ctx, int32_t n_ranks, int32_t self, void** row_dev);
int pcf_exchange_scatter_async(pcf_ctx* ctx, void* const* dst_bufs, const uint64_t* dst_offsets);
int pcf_recv_buffer(pcf_ctx* ctx, uint64_t n_records, void** dev_ptr);
int pcf_ipc_export(pcf_ctx* ctx, void* handle64);                          /* cudaIpcMemHandle_t of the receive buffer */
int pcf_ipc_open(pcf_ctx* ctx, const void* handle64, void** peer_ptr);
int pcf_ipc_close_all(pcf_ctx* ctx);
int pcf_install_records(pcf_ctx* ctx, const void* records_dev, uint64_t n_records);
/* Helpers for a host that drives several contexts from ONE process (host/pcf_replay --gpus N): rows [first, first+count)
 * of the viewpoint table as 4 floats per frame (x, y, z, 1 when set), and peer access from this context's device to
 * `peer_device` so that pcf_exchange_scatter may store into another context's receive buffer directly. */
int pcf_get_viewpoints(pcf_ctx* ctx, float* host4, uint32_t first, uint32_t count);
int pcf_set_viewpoints(pcf_ctx* ctx, const float* host4, uint32_t first, uint32_t count);
int pcf_enable_peer_access(pcf_ctx* ctx, int32_t peer_device);

/* ---- interleaved update schedules across ranks: replicated grid state, sharded ingest / normals / scoring / extraction ---------
 * (the live node's cleanGrid pass every 5 s, node.cpp:301-325, combined with frame sharding).  Between two update passes the
 * frames are split over the ranks in contiguous sub-blocks.  At an update point:
 *   pcf_round_export    this rank's records since the last round, (x, y, z, frame_idx) in arrival order, device memory;
 *   [the caller all-gathers the records in rank order = frame order = arrival order, and sums the viewpoint rows of the round]
 *   pcf_round_install   every rank appends the SAME gathered records to its log (and applies them to the first-frame grid and the
 *                       occupancy bitmap): the grid state is now identical on all ranks and, record for record, the one a single
 *                       GPU fed every frame would hold (same arrival order, same cursor at the pass);
 *   pcf_set_slab + pcf_update_local   neighbour scan + PCA normals of the candidates in this rank's x-slab (any partition works,
 *                       it may change from pass to pass: the state is replicated), records NOT committed;
 *   [the caller all-gathers (cell, normal) in slab order = x-major order]
 *   pcf_update_commit   every rank appends the same records as this pass; holders / dependants (OG.hpp:417,443-449) follow from
 *                       the replicated records exactly as on one GPU.
 * pcf_extract with a slab set scores and extracts only that slab; the slabs concatenated in rank order are the x-major scan.
 * pcf_update == pcf_update_local + pcf_update_commit of the own records. */
int pcf_round_export(pcf_ctx* ctx, void** records_dev, uint64_t* n_records);
int pcf_round_install(pcf_ctx* ctx, const void* records_dev, uint64_t n_records);
int pcf_update_local(pcf_ctx* ctx, void** cells_dev /* uint32 */, void** normals_dev /* float4 */, uint32_t* n_new);
int pcf_update_commit(pcf_ctx* ctx, const void* cells_dev, const void* normals_dev, uint32_t n);

/* ---- known-answer hooks: run ONE device function over an array (tests bit-compare with the oracle) ---- */
int pcf_kat_transform_voxel(pcf_ctx* ctx, const float* pts_host, uint32_t n, uint32_t stride_floats,
                            const double pose[16], float* world_xyz, int32_t* ijk, uint8_t* kept);
/* clip-and-pack of the staging pool (host code; runs without a context or a GPU).  isa: 0 scalar, 1 AVX2, 2 AVX-512, -1 = what
 * the pool uses on this CPU; out_xyz needs 3 * rows * cols + 16 floats.  Returns the implementation that ran (>= 0). */
int pcf_kat_clip_pack(const uint8_t* data, uint32_t rows, uint32_t cols, uint32_t point_step, uint64_t row_step, uint32_t x_offset,
                      float clip_lo, float clip_hi, int32_t isa, float* out_xyz, uint32_t* n_out);
/* the cooperative scoring kernel's division (shared reciprocal of float(count)) against the compiler's IEEE division */
int pcf_kat_div(pcf_ctx* ctx, const float* x_host, const float* c_host, uint32_t n, uint32_t* mismatches, uint32_t* first_bad);
int pcf_kat_normal(pcf_ctx* ctx, const float* xyz_host, uint32_t n_points, float* normal3);
int pcf_kat_format_float(float v, int precision, char* out32);   /* the writer's float formatting (6 = CSV %g, 8 = PCD %.8g) */
int pcf_kat_score(pcf_ctx* ctx, const float* xyz_host, uint32_t n_points, const float axis_pt[3],
                  const float normal[3], float* centroid3, float* sd3, float* mean_dist, float* sd_dist,
                  int32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* PCFUSION_H */
