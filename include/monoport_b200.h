/*
 * monoport_b200 -- C ABI of the B200-native occupancy-field hot path of MonoPort.
 *
 * The reference (Project-Splinter/MonoPort) is pure Python: it has no FFI.  Its boundary for this path is
 * a Python call protocol.  This header is the C-ABI a binding for that protocol
 * would call; every entry point names the reference interface it replaces (paths relative to the
 * reference tree).  Plain pointers and sizes only -- no torch types.  INTEGRATION.md shows the ctypes
 * stub (monoport_b200/_lib.py is that stub).
 *
 * Conventions
 *   - every function returns MP_OK (0) or a negative MP_E_* code; mp_last_error() gives a thread-local
 *     message.  An empty reconstruction is NOT an error (RTL/recon.py:32-33 tolerates None).
 *   - "dev" pointers are CUDA device pointers of the current device, "host" pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are asynchronous
 *     on that stream unless stated otherwise.  No process-global mutable state: all scratch lives in
 *     the handles, so netG and netC queries may run concurrently from different threads
 *     (RTL/dataloader.py:734-751 runs every stage in its own thread).
 *   - volumes are [D,H,W] = [z,y,x] float32, x fastest (RTL/recon.py:35-38).
 */
#ifndef MONOPORT_B200_H_
#define MONOPORT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP_OK 0
#define MP_E_INVALID (-1)   /* bad argument                                  */
#define MP_E_CUDA (-2)      /* a CUDA runtime call failed                    */
#define MP_E_UNSUPPORTED (-3) /* shape / mode not supported by this kernel   */
#define MP_E_CAPACITY (-4)  /* caller-provided output buffer too small       */
#define MP_E_NOMEM (-5)
#define MP_E_RANGE (-6)     /* inputs outside the validated range of the requested fast path; use the exact one */

/* last_op of the head: heads/SurfaceClassifier.py:68-69 (None / nn.Sigmoid / nn.Tanh) */
#define MP_LAST_NONE 0
#define MP_LAST_SIGMOID 1
#define MP_LAST_TANH 2

/* projection: geometry.py:19-34 (orthogonal) / :37-55 (perspective) */
#define MP_PROJ_ORTHOGONAL 0
#define MP_PROJ_PERSPECTIVE 1

/* arithmetic of the fused sample+MLP kernel */
#define MP_MODE_FP32 0   /* CUDA-core fp32 everywhere (|err| ~1e-6 vs the reference)                 */
#define MP_MODE_TC 1     /* tcgen05 fp16 operands / fp32 TMEM accumulators, last layer fp32 (<=1e-4)   */
#define MP_MODE_AUTO 2   /* TC when the head/feature shape is supported by the tcgen05 kernel, else FP32 */
/* 3 was MP_MODE_TC_V2 (the self-contained tensor-core program, removed in round 2: 335 vs 485 Mpoints/s); rejected now */
#define MP_MODE_TC_V3 4  /* alias of MP_MODE_TC: the one tensor-core program (layer 0 hoisted to texels)                     */

const char* mp_last_error(void);
int mp_version(void);
/* sm count / compute capability of the current device */
int mp_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Head weights.  Replaces the parameters of SurfaceClassifier (heads/SurfaceClassifier.py:7-37):
 * `channels` = filter_channels (n_layers+1 entries, e.g. {257,1024,512,256,128,1});
 * weights[l] is filters.l.weight as [Cout_l, Cin_l] row-major fp32 with Cin_l = channels[l] +
 * (l>0 && skip ? channels[0] : 0)  (:24-34), biases[l] is [Cout_l].  skip = !no_residual.
 * Pointers are host (on_device=0) or device (on_device=1) memory; they are only read during the call.
 * The handle owns repacked copies (fp32 K-major for the CUDA-core kernel, fp16 UMMA tiles for tcgen05).
 * ------------------------------------------------------------------------------------------- */
typedef struct mp_mlp mp_mlp_t;
int mp_mlp_create(int n_layers, const int* channels, const float* const* weights,
                  const float* const* biases, int skip, int last_op, int on_device, mp_mlp_t** out);
int mp_mlp_destroy(mp_mlp_t* h);
/* 1 if MP_MODE_TC is available for this head on this device */
int mp_mlp_tc_supported(const mp_mlp_t* h);
/* Validated range of the tensor-core program (no reference counterpart: the reference computes in fp32 throughout,
 * MonoPortNet.py:86-89).  MP_MODE_TC / MP_MODE_AUTO evaluate a frame on the tensor cores only while its largest
 * |feature| stays under `limit` (defaults: 12 geometry head, 8 colour head -- both keep |error| <= 1e-4 on what query()
 * returns); frames above it take the exact fp32 kernel.  The choice is made on the device from the frame itself.
 * +infinity disables the guard. */
int mp_mlp_set_tc_feature_limit(mp_mlp_t* h, float limit);

/* ---------------------------------------------------------------------------------------------
 * Feature volume.  Replaces the tensor index() samples (geometry.py:4-16): one [C,H,W] fp32 NCHW map
 * (the last hourglass stage in eval mode, MonoPortNet.py:63-64).  The handle keeps a channel-last fp32
 * copy so one bilinear tap is one contiguous vector (taps are always read in fp32).
 * ------------------------------------------------------------------------------------------- */
typedef struct mp_feat mp_feat_t;
int mp_feat_create(int C, int H, int W, mp_feat_t** out);
int mp_feat_upload(mp_feat_t* h, const float* nchw, int on_device, void* stream);
/* same for a map that is already channel-last ([H,W,C] fp32, device memory -- a torch.channels_last tensor): one copy, no
 * transposing kernel */
/* Zero-copy variant: the handle reads the caller's channel-last [H,W,C] fp32 device map IN PLACE for this frame (what a
 * channels_last encoder emits, HGFilters.py:158-159 / RTL/main.py:382-387: no repack kernel, no copy).  The memory must stay
 * valid and unchanged until the frame's queries have completed; the next upload / bind replaces the binding. */
int mp_feat_bind_nhwc(mp_feat_t* h, const float* nhwc_dev, void* stream);
int mp_feat_upload_nhwc(mp_feat_t* h, const float* nhwc_dev, void* stream);
int mp_feat_destroy(mp_feat_t* h);

/* calib: 12 host floats = rows 0..2 of the [4,4] / [3,4] calibration (R | t), or NULL for calibs=None
 * (MonoPortNet.py:66-67).  z_scale = DepthNormalizer scale (normalizers/DepthNormalizer.py:32,40). */

/* MonoPortNet.query(feats, points[1,3,N], calibs) -> [Res,N]   (MonoPortNet.py:48-91)
 * points_dev: coordinate a (0..2) of point i lives at points_dev[a*row_stride + i*point_stride]  -- (N,1) for a
 * contiguous [3,N] tensor, (1,3) for the permuted [N,3] view RTL/main.py:176-177 passes.
 * out_dev: [Res, N], row stride ld_out. */
int mp_query_points(mp_mlp_t* mlp, mp_feat_t* feat, const float* points_dev, int64_t n, int64_t row_stride,
                    int64_t point_stride, const float* calib12, int projection, float z_scale, float* out_dev,
                    int64_t ld_out, int mode, void* stream);
/* Same call with HOST buffers (points [3,N], out [Res,N], feature map NCHW fp32 or NULL to reuse the
 * uploaded one): copies in, runs, copies out, synchronises.  This is what bench.py's e2e leg times. */
int mp_query_points_host(mp_mlp_t* mlp, mp_feat_t* feat, const float* feat_nchw_host, const float* points_host,
                         int64_t n, const float* calib12, int projection, float z_scale, float* out_host,
                         int mode, void* stream);

/* Dense grid query: the node centres of planes [z0, z0+nz) of an R^3 grid spanning [b_min,b_max]
 * (world = (idx+0.5)/R*(b_max-b_min)+b_min, the engine's convention, cf. RTL/main.py:204-209) are generated
 * in-kernel; out_dev is the [nz,R,R] slab, channel 0 of the head (Res==1 required).  z-slab sharding
 * across GPUs calls this with a different [z0,nz) per rank. */
int mp_query_grid(mp_mlp_t* mlp, mp_feat_t* feat, int R, int z0, int nz, const float* b_min3, const float* b_max3,
                  const float* calib12, int projection, float z_scale, float* out_dev, int mode, void* stream);
/* The same for any contiguous range [lin0, lin0+n) of the grid's z-major linear node order (a z slab is the range
 * [z0*R*R, (z0+nz)*R*R)): balanced sharding over GPUs whose boundaries need not fall on plane boundaries (257 = 8*32+1:
 * with whole planes one rank carries 33 planes and the others 32).  out_dev receives n values. */
int mp_query_grid_range(mp_mlp_t* mlp, mp_feat_t* feat, int R, int64_t lin0, int64_t n, const float* b_min3,
                        const float* b_max3, const float* calib12, int projection, float z_scale, float* out_dev, int mode,
                        void* stream);

/* Fused slab exchange (instead of the all-gather): the slab [z0, z0+nz) is evaluated like mp_query_grid,
 * but every value is stored straight into the FULL [R,R,R] volumes of all n_peers ranks (peer_vols: host array of
 * device pointers, the caller's own volume included; peers' volumes are peer-memory mappings obtained through
 * mp_ipc_open) while the tiles are computed -- compute and transfer in ONE kernel over NVLink.  The caller synchronises
 * the ranks afterwards (any barrier ordered after this call on `stream`); with two alternating volume sets one barrier
 * per frame suffices.  n_peers <= 8.  mp_query_grid_range_peers: the same for a range of the linear node order. */
int mp_query_grid_peers(mp_mlp_t* mlp, mp_feat_t* feat, int R, int z0, int nz, const float* b_min3, const float* b_max3,
                        const float* calib12, int projection, float z_scale, float* const* peer_vols, int n_peers,
                        int mode, void* stream);
int mp_query_grid_range_peers(mp_mlp_t* mlp, mp_feat_t* feat, int R, int64_t lin0, int64_t n, const float* b_min3,
                              const float* b_max3, const float* calib12, int projection, float z_scale,
                              float* const* peer_vols, int n_peers, int mode, void* stream);
/* Exportable device memory for such volumes: cudaMalloc + the 64-byte cudaIpcMemHandle_t to hand to the other ranks
 * (one process per GPU), which map it with mp_ipc_open (enables peer access) and unmap it with mp_ipc_close. */
int mp_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64);
int mp_ipc_open(const unsigned char* handle64, void** dev_ptr);
int mp_ipc_close(void* dev_ptr);
int mp_ipc_free(void* dev_ptr);

/* mp_query_grid with HOST buffers: uploads the NCHW fp32 feature map (NULL = reuse), evaluates the slab, copies
 * the [nz,R,R] result to out_host, synchronises.  Slabs of 4 M nodes or more are evaluated and read back in four z-chunks,
 * the copy of one chunk (internal copy stream) overlapping the evaluation of the next: pin out_host to get the overlap.
 * bench.py's e2e leg. */
int mp_query_grid_host(mp_mlp_t* mlp, mp_feat_t* feat, const float* feat_nchw_host, int R, int z0, int nz,
                       const float* b_min3, const float* b_max3, const float* calib12, int projection, float z_scale,
                       float* out_host, int mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Coarse-to-fine occupancy engine.  Replaces implicit_seg.functional.Seg3dLossless / Seg3dTopk
 * (third-party, un-vendored; call sites RTL/main.py:28-29,188-195,390-395).
 *   resolutions: n_levels odd cube sizes 2^k+1 (RTL/main.py:187), coarse -> fine.
 *   faster != 0 : RTL/main.py:195 mode (box dilation 9/7/3, last level interpolated only)
 *   faster == 0 : lossless mode (k=3, every level examined, conflict loop)
 *   topk_points : NULL for the boundary engine, else n_levels ints = nodes evaluated per level (Seg3dTopk)
 * Two ways to drive it:
 *   (a) stepping API -- the binding calls back into an arbitrary Python query_func per step
 *       (mp_octree_begin -> loop { mp_octree_next -> query -> mp_octree_commit });
 *   (b) mp_octree_run_fused -- the whole pyramid with the fused sample+MLP kernel, no host round trip per level.
 * ------------------------------------------------------------------------------------------- */
typedef struct mp_octree mp_octree_t;
int mp_octree_create(int n_levels, const int* resolutions, const float* b_min3, const float* b_max3,
                     float balance_value, int faster, const int* topk_points, mp_octree_t** out);
int mp_octree_destroy(mp_octree_t* h);
int mp_octree_begin(mp_octree_t* h, void* stream);
/* Produces the next batch of nodes to evaluate.  *n_out = number of nodes (0 => reconstruction finished);
 * *level_out = level they belong to.  points_dev_out: device pointer to [n,3] world-space points (owned by
 * the handle, valid until the next call), idx_dev_out: their linear indices (z*R*R+y*R+x) at that level.
 * Synchronises the stream (the count has to reach the host). */
int mp_octree_next(mp_octree_t* h, int64_t* n_out, int* level_out, const float** points_dev_out,
                   const int32_t** idx_dev_out, void* stream);
/* Scatter the n values ([n] floats, device) returned by the query for the batch handed out by the last
 * mp_octree_next. */
int mp_octree_commit(mp_octree_t* h, const float* values_dev, void* stream);
/* After the last step: copies the final [R,R,R] volume to out_dev.  *nonempty = 0 when level 0 had no
 * value > balance (the engine then returns None, RTL/recon.py:32-33) -- synchronises. */
int mp_octree_finish(mp_octree_t* h, float* out_dev, int* nonempty, void* stream);
/* Whole pyramid on the device.  stats_host (may be NULL): n_levels int64 = nodes evaluated per level.
 * Synchronises once at the end (to report nonempty/stats). */
int mp_octree_run_fused(mp_octree_t* h, mp_mlp_t* mlp, mp_feat_t* feat, const float* calib12, int projection,
                        float z_scale, int mode, float* out_dev, int* nonempty, int64_t* stats_host, void* stream);
/* Multi-GPU list sharding of mp_octree_run_fused (one process per GPU, one handle per rank, same ctor
 * arguments).  Every rank keeps the whole pyramid and runs the volume passes itself -- identical inputs give identical
 * node lists, the lossless conflict loop included -- and evaluates only its window of each level's ordered node list
 * (balanced to one 128-point tile); the fused kernel stores the values into the value lists of ALL ranks over NVLink peer
 * memory and a flag barrier between the ranks' streams replaces the collective.  Every rank ends up with the full volume.
 *   mp_octree_shard_export : 3 x 64-byte IPC handles (two value lists, the barrier flags) to hand to the other ranks;
 *   mp_octree_shard_set    : rank / world and, per rank, the peer mappings of those three blocks (mp_ipc_open; entry
 *                            [rank] is ignored).  world == 1 switches sharding off.  All ranks must then call
 *                            mp_octree_run_fused the same number of times with the same inputs. */
/* The fused run in two halves for CUDA-graph captured frame steps: _async only enqueues (no host memory is read, no
 * synchronisation: capturable; engines without a conflict loop, i.e. faster != 0 or top-k), mp_octree_fetch reads the
 * non-empty flag and the per-level counts back afterwards (synchronises). */
int mp_octree_run_fused_async(mp_octree_t* h, mp_mlp_t* mlp, mp_feat_t* feat, const float* calib12, int projection,
                              float z_scale, int mode, float* out_dev, void* stream);
int mp_octree_fetch(mp_octree_t* h, int* nonempty, int64_t* stats_host, void* stream);
int mp_octree_shard_export(mp_octree_t* h, unsigned char* handles192);
int mp_octree_shard_set(mp_octree_t* h, int rank, int world, float* const* vals0, float* const* vals1,
                        uint32_t* const* flags);

/* ---------------------------------------------------------------------------------------------
 * Marching cubes (absent from the reference; PIFu-style reconstruction() asked for by the north star).
 * Two calls: count (synchronises, returns sizes) then emit into caller buffers.
 *   verts: [nV,3] float32 (x,y,z) in index space; faces: [nF,3] int32.
 * Deterministic: vertex ids follow (node, axis) order, faces follow (cell, table) order.
 * ------------------------------------------------------------------------------------------- */
typedef struct mp_mcubes mp_mcubes_t;
int mp_mcubes_create(int D, int H, int W, mp_mcubes_t** out);
int mp_mcubes_destroy(mp_mcubes_t* h);
int mp_mcubes_count(mp_mcubes_t* h, const float* vol_dev, float iso, int64_t* n_verts, int64_t* n_faces, void* stream);
int mp_mcubes_emit(mp_mcubes_t* h, const float* vol_dev, float iso, float* verts_dev, int32_t* faces_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Visible-surface extraction.  Replaces forward_vertices (RTL/recon.py:27-89).
 * direction: 0 front, 1 back, 2 left, 3 right.  Outputs (device, capacity R*R): X,Y int64, Z float, norm [n,3].
 * *n_out on the host (synchronises).
 * ------------------------------------------------------------------------------------------- */
int mp_forward_vertices(const float* vol_dev, int R, int direction, int64_t* x_dev, int64_t* y_dev, float* z_dev,
                        float* norm_dev, int64_t* n_out, void* stream);
/* Enqueue-only variant for CUDA-graph captured frame steps (the reference overlaps its stages with one
 * Python thread per processor, RTL/dataloader.py:734-751): no allocation, no host read-back.  scratch_dev: at least
 * mp_forward_vertices_scratch_bytes(R) bytes of device memory; count_dev: one int64 on the device. */
int64_t mp_forward_vertices_scratch_bytes(int R);
int mp_forward_vertices_async(const float* vol_dev, int R, int direction, int64_t* x_dev, int64_t* y_dev, float* z_dev,
                              float* norm_dev, int64_t* count_dev, void* scratch_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Direct rendering of the visible surface with the colour head.  Replaces colorization (RTL/main.py:212-249): vertex i =
 * (X[i], Y[i], R - Z[i]) of mp_forward_vertices -> world space by mat_color (:201-210) -> netC.query -> pred*0.5+0.5 ->
 * canvas[X[i], Y[i], :]  ([R,R,3] float32, prepared by the caller), in ONE launch without intermediate tensors.
 * Needs the tensor-core program of the colour head (MP_E_UNSUPPORTED otherwise) and a frame inside its validated feature
 * range (MP_E_RANGE otherwise, see mp_mlp_set_tc_feature_limit) -- the binding then takes the generic mp_query_points
 * route.
 * ------------------------------------------------------------------------------------------- */
int mp_colorize_surface(mp_mlp_t* mlp, mp_feat_t* feat, const int64_t* x_dev, const int64_t* y_dev, const float* z_dev,
                        int64_t n, int R, const float* b_min3, const float* b_max3, const float* calib12, int projection,
                        float z_scale, float* canvas_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MONOPORT_B200_H_ */
