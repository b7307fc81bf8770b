/*
 * saf_b200.h -- C ABI of libsaf_b200.so: the B200 (sm_100a) implementation of the
 * spatially_aware_AI RGB-D fusion + language-query hot path.
 *
 * The reference has no FFI of its own for this path: its boundary is the Python class surface
 * (SURVEY.md section 8b).  Each entry point below names the reference code it replaces; the
 * Python classes in spatially_aware_ai_b200/ (same names/signatures as the reference's) are
 * thin ctypes callers of these functions.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no torch / C++ types.
 *   - every pointer marked "device" is a CUDA device pointer valid on the current device;
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - all functions are asynchronous w.r.t. the host unless stated otherwise, allocate no
 *     device memory (the caller owns the workspace) and keep no global mutable state, so
 *     one workspace per (volume, stream) is safe.  A volume must not be integrated from two
 *     streams at once (the reference's integrate() is not re-entrant either).
 *   - return value: 0 = ok, >0 = cudaError_t of a failed CUDA call, <0 = SAF_ERR_* argument error.
 *   - layouts follow the reference buffers (clip_seem_fusion.py:640-672): voxel (x,y,z) of the
 *     grid lives at flat index (x*ny + y)*nz + z; an x-slab [x_begin,x_end) is the contiguous
 *     flat range starting at x_begin*ny*nz, and slab-local buffers start at that voxel.
 */
#ifndef SAF_B200_H
#define SAF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAF_ABI_VERSION 3
#define SAF_MAX_BATCH 16         /* frames per integrate() call / per window; every reference caller uses 1 */
#define SAF_BLOCK_EDGE 8         /* voxel blocks are 8x8x8 */

/* argument errors */
#define SAF_ERR_NULL        (-1)
#define SAF_ERR_BATCH       (-2)
#define SAF_ERR_GRID        (-3)
#define SAF_ERR_WORKSPACE   (-4)  /* workspace too small for this grid / batch / table */
#define SAF_ERR_SHAPE       (-5)
#define SAF_ERR_DTYPE       (-6)
#define SAF_ERR_ALIGNMENT   (-7)
#define SAF_ERR_UNSUPPORTED (-8)
#define SAF_ERR_DEVICE      (-9)  /* not an sm_100 device */

/* class-map element types accepted for saf_frame.seg (the reference passes an int tensor and
 * converts with .float(), clip_seem_fusion.py:755-760) */
#define SAF_SEG_NONE 0
#define SAF_SEG_U8   1
#define SAF_SEG_I16  2
#define SAF_SEG_I32  3
#define SAF_SEG_I64  4
#define SAF_SEG_F32  5

/* Element types of saf_frame.depth / saf_frame.rgb.  The reference's integrate() takes fp32 tensors; its
 * dataset classes produce them from sensor formats with depth = float(u16 millimetres) / 1000 and
 * rgb = float(u8) / 255 (clipfusion.py:185-188, 245-254, 355-362).  The sensor-format types move 3.25x fewer
 * bytes host -> device and are converted in-kernel with exactly those roundings. */
#define SAF_DEPTH_F32    0   /* metres                                  */
#define SAF_DEPTH_U16_MM 1   /* uint16 millimetres, depth = v / 1000    */
#define SAF_RGB_F32      0   /* [0,1]                                   */
#define SAF_RGB_U8       1   /* uint8, rgb = v / 255                    */

/* Feature source of a frame (saf_frame.table_mode).
 * SAF_TABLE_PATCH_GRID: the reference's - `table` is the tiled-patch CLIP feature image [C,npy,npx], sampled
 *   bilinearly with zeros padding at the voxel's projection (clip_seem_fusion.py:800-805).  The default.
 * SAF_TABLE_SEGMENTS: BASELINE.json north_star's "segment -> CLIP table" (an extension, the
 *   reference has no such mode): `table` holds one embedding per segment / class id, npy = 1, npx = n_segments,
 *   and a voxel's sample is the row of its nearest-sampled class id - the id of saf_frame.seg that also feeds the
 *   label histogram (clip_seem_fusion.py:786-791; id 0 for a pixel outside the image).  Same kernels, a one-tap
 *   generator (row id, weight 1) instead of the four bilinear taps; needs seg.  An id outside [0, npx) samples
 *   zeros and sets SAF_FLAG_BAD_CLASS_ID. */
#define SAF_TABLE_PATCH_GRID 0
#define SAF_TABLE_SEGMENTS   1

/* rgb sampling: clipfusion.py:701-706 (nearest) / clip_seem_fusion.py:793-798 (bilinear) */
#define SAF_RGB_NEAREST  0
#define SAF_RGB_BILINEAR 1

/* bits of saf_stats.error_flags */
#define SAF_FLAG_BAD_CLASS_ID 1u  /* a sampled class id was outside [0,n_classes): torch one_hot raises */

/* Voxel grid geometry.  ClipSeemFusion.__init__ arguments origin / voxel_size / nvox
 * (clip_seem_fusion.py:612-672); x_begin/x_end select the x-slab the buffers hold
 * (0 / nvox[0] for the whole grid).  Voxel centres are fl(fl(i)*voxel_size)+origin with the
 * GLOBAL index i, exactly as clip_seem_fusion.py:664-669 computes xyz_world.
 * Block-cyclic slabs (multi-GPU load balance): with x_span > 0 the buffers hold the stripes
 * [x_begin + k*x_stride, x_begin + k*x_stride + x_span) for k = 0, 1, ... clipped to x_end, concatenated in
 * order (rank r of n: x_begin = r*x_span, x_stride = n*x_span, x_end = nvox[0]).  x_span and x_stride are
 * multiples of SAF_BLOCK_EDGE.  Fusion, label argmax and the query accept such slabs; the mesh and object
 * entry points need contiguous slabs (x_span = 0) and return SAF_ERR_UNSUPPORTED otherwise.
 * Sheared block columns (y_ranks = n > 1; x_begin = 0, x_end = nvox[0], x_span = 0): the grid is cut into columns
 * of 8 x 8 x nz voxels and column (bx, by) belongs to rank (bx + by) mod n, so that a surface lying in ONE x-plane or
 * ONE y-plane (a wall of an axis-aligned room) is spread over all ranks - x-stripes leave it on one.  A rank's
 * buffers are a dense [nvox[0], ny_local, nz] grid, ny_local = 8 * ceil(ceil(ny / 8) / n): local y-block j of
 * x-block bx is global y-block j*n + ((y_rank - bx) mod n); local columns beyond the grid are never touched. */
typedef struct saf_grid_desc {
    float   origin[3];
    float   voxel_size;
    int32_t nvox[3];
    int32_t x_begin;
    int32_t x_end;
    int32_t x_span;    /* 0 = one contiguous slab [x_begin, x_end) */
    int32_t x_stride;
    int32_t y_ranks;   /* > 1: sheared block-column layout over this many ranks (below), else 0 */
    int32_t y_rank;    /* this rank's index in it */
} saf_grid_desc;

/* Device buffers of one slab; names, dtypes and shapes are the reference's registered buffers
 * (clip_seem_fusion.py:640-659).  labels_one_hot may be NULL (ClipFusion has none). */
typedef struct saf_volume {
    float   *tsdf;            /* [N]            */
    int32_t *tsdf_weight;     /* [N]            */
    int32_t *weight;          /* [N]            */
    float   *rgb;             /* [N,3]          */
    float   *clip_feat;       /* [N,feature_dim]*/
    int32_t *labels_one_hot;  /* [N,n_classes] or NULL */
    int32_t  feature_dim;
    int32_t  n_classes;       /* 143 in the reference (133 + 10) */
} saf_volume;

/* One RGB-D frame plus the two producer outputs integrate() pulls in
 * (clip_seem_fusion.py:691-695 feature image, :755 class map). */
typedef struct saf_frame {
    const void  *depth;       /* device [H,W] of depth_dtype (f32 metres), 0 = missing             */
    const void  *rgb;         /* device [H,W,3] of rgb_dtype (f32 in [0,1]; the reference's rgb_imgs[b], HWC) */
    const void  *seg;         /* device [H,W] class ids of type seg_dtype, or NULL                   */
    const float *table;       /* device feature image: element (c, py, px) at                        */
    int64_t      table_stride_c;  /*   table[c*stride_c + (py*npx+px)*stride_r]  (elements)          */
    int64_t      table_stride_r;
    int32_t      npy, npx;    /* patch grid (clipfusion.py:795-796)                                  */
    int32_t      seg_dtype;   /* SAF_SEG_*                                                           */
    int32_t      depth_dtype; /* SAF_DEPTH_*                                                         */
    float        pose[16];    /* camera->world, row-major 4x4 (clipfusion.py:308-312 axes)           */
    float        K[9];        /* intrinsics, row-major 3x3                                           */
    int32_t      rgb_dtype;   /* SAF_RGB_F32 / SAF_RGB_U8                                            */
    const float *pose_device; /* optional device copies of pose[16] / K[9]: when non-NULL the kernels  */
    const float *K_device;    /* read these instead (the reference's callers hand integrate() CUDA     */
                              /* tensors, clip_seem_fusion.py:308-311; no host sync needed this way)   */
    int32_t      table_mode;  /* SAF_TABLE_*                                                         */
    int32_t      reserved;
} saf_frame;

/* Counters kept in the workspace (device) and copied out by saf_read_stats. */
typedef struct saf_stats {
    uint64_t total_frames;       /* frames integrated since saf_workspace_init                       */
    uint64_t total_valid;        /* sum over frames of #voxels with `valid` (feature updates)         */
    uint64_t total_tsdf_valid;   /* sum over frames of #voxels with `tsdf_valid`                      */
    uint64_t total_blocks;       /* sum over calls of visible 8^3 blocks                              */
    uint32_t last_blocks;        /* visible blocks of the most recent call                            */
    uint32_t last_valid[SAF_MAX_BATCH];
    uint32_t last_tsdf_valid[SAF_MAX_BATCH];
    uint32_t error_flags;        /* SAF_FLAG_* (sticky)                                               */
    uint32_t last_processed;     /* visible blocks of the last call that survived K2's depth test     */
    uint32_t depth_cull_on;      /* 1 while the adaptive depth-aware block cull is switched on        */
    uint64_t total_calls;        /* integrate() calls / windows launched (one K0+K1+K2 trio each)      */
    uint64_t total_union;        /* window mode: sum over windows of the voxels valid in >= 1 frame of the window
                                    (= feature rows read and written once per window)                  */
    uint32_t last_union;         /* ... of the most recent window                                      */
    uint32_t reserved;
} saf_stats;

/* Caller-owned device scratch.  `base` is a device allocation of `bytes` (>= saf_workspace_bytes
 * for the same grid / max_batch / max_table_elems), 256-byte aligned. */
typedef struct saf_workspace {
    void    *base;
    uint64_t bytes;
    int32_t  max_batch;         /* largest batch integrate() will be called with (<= SAF_MAX_BATCH) */
    int32_t  reserved;
    int64_t  max_table_elems;   /* largest npy*npx*feature_dim of a feature image                    */
} saf_workspace;

int         saf_abi_version(void);
const char *saf_error_string(int code);

/* ---- workspace -------------------------------------------------------------------------- */

/* Bytes of device scratch needed to integrate batches of up to max_batch frames with feature
 * images of up to max_table_elems (= npy*npx*C) elements into this slab. */
int saf_workspace_bytes(const saf_grid_desc *grid, int32_t max_batch, int64_t max_table_elems,
                        uint64_t *bytes_out);
/* Zero the counters and record the layout; call once after allocating. */
int saf_workspace_init(const saf_workspace *ws, const saf_grid_desc *grid, void *stream);
/* Copy the counters to host memory.  Synchronises `stream`. */
int saf_read_stats(const saf_workspace *ws, saf_stats *out_host, void *stream);

/* ---- fusion: ClipSeemFusion.integrate (clip_seem_fusion.py:676-822) and
 *              ClipFusion.integrate     (clipfusion.py:627-721) ------------------------------ */

/* K1  frame set-up: conservative frustum test of every 8^3 voxel block against the B camera
 * frusta -> compact visible-block list in the workspace; repacks channel-major feature images
 * to [R,C] rows.  Replaces nothing in the reference (it projects all N voxels,
 * clip_seem_fusion.py:698-712); it bounds the work of K2. */
int saf_frustum_cull(const saf_grid_desc *grid, const saf_frame *frames, int32_t batch,
                     int32_t height, int32_t width, float trunc, const saf_workspace *ws, void *stream);

/* K2  projection + nearest depth sample + sdf/masks + TSDF and tsdf_weight running average
 * (clip_seem_fusion.py:698-744) for the voxels of the visible blocks; appends every `valid`
 * voxel (index, gx, gy) to the per-frame list in the workspace.
 * valid_out / tsdf_valid_out: optional device [batch, N] uint8 masks (0/1), must be zeroed by
 * the caller (only visited voxels are written). */
int saf_tsdf_update(const saf_grid_desc *grid, const saf_volume *vol, const saf_frame *frames,
                    int32_t batch, int32_t height, int32_t width, float trunc, const saf_workspace *ws,
                    uint8_t *valid_out, uint8_t *tsdf_valid_out, void *stream);

/* K3  for every listed voxel of frame `frame_index`: nearest class id -> one histogram counter,
 * rgb sample, bilinear feature sample from the [R,C] table (staged in shared memory by TMA),
 * running averages of rgb and clip_feat, weight += 1  (clip_seem_fusion.py:751-822). */
int saf_feature_accumulate(const saf_grid_desc *grid, const saf_volume *vol, const saf_frame *frames,
                           int32_t batch, int32_t frame_index, int32_t height, int32_t width,
                           int32_t rgb_mode, const saf_workspace *ws, void *stream);

/* Window mode (what saf_integrate_sequence issues for every run of up to SAF_MAX_BATCH consecutive frames):
 * the `batch` frames are successive single-frame integrate() calls.  K2 advances each voxel's TSDF frame by
 * frame in registers and emits ONE list of the voxels valid in any frame; K3 reads each listed feature row
 * once, applies the valid frames' updates in order and writes it back once.  Results are those of the
 * single-frame calls.  Call saf_frustum_cull on the same batch first.  2 <= batch <= the workspace's
 * max_batch <= SAF_MAX_BATCH. */
int saf_tsdf_update_window(const saf_grid_desc *grid, const saf_volume *vol, const saf_frame *frames,
                           int32_t batch, int32_t height, int32_t width, float trunc,
                           const saf_workspace *ws, void *stream);
int saf_feature_accumulate_window(const saf_grid_desc *grid, const saf_volume *vol, const saf_frame *frames,
                                  int32_t batch, int32_t height, int32_t width, int32_t rgb_mode,
                                  const saf_workspace *ws, void *stream);

/* The two stages of saf_feature_accumulate_window, separately (what the bench times one by one):
 * SAF_STAGE_TILE_SETUP   repacks the frames' feature images and runs K2T, which prepares the update metadata of the
 *                        window's tiles and applies the window's rgb / weight / label-counter updates
 *                        (clip_seem_fusion.py:786-798, 808-813, 820-822);
 * SAF_STAGE_ACCUMULATE   K3W, the running average of the feature rows (clip_seem_fusion.py:800-814).
 * SAF_STAGE_ACCUMULATE alone is only valid after SAF_STAGE_TILE_SETUP on the same window. */
#define SAF_STAGE_TILE_SETUP 1
#define SAF_STAGE_ACCUMULATE 2
int saf_feature_accumulate_window_stages(const saf_grid_desc *grid, const saf_volume *vol, const saf_frame *frames,
                                         int32_t batch, int32_t height, int32_t width, int32_t rgb_mode,
                                         const saf_workspace *ws, int32_t stages, void *stream);

/* One reference integrate() call: K1, K2, then K3 for each frame of the batch in order. */
int saf_integrate(const saf_grid_desc *grid, const saf_volume *vol, const saf_frame *frames,
                  int32_t batch, int32_t height, int32_t width, float trunc, int32_t rgb_mode,
                  const saf_workspace *ws, void *stream);

/* Host-side, pose-only test: 1 if the camera frustum of a frame (pose: camera->world row-major 4x4, K: row-major
 * 3x3, both HOST pointers) may contain a voxel centre of the slab `grid` describes, 0 if it cannot (conservative),
 * < 0 on bad arguments.  No device work: a multi-GPU feeder calls it to decide whether a frame needs to be copied
 * to this rank at all.  Block-cyclic slabs always answer 1. */
int saf_frame_reaches_slab(const saf_grid_desc *grid, const float *pose, const float *K, int32_t height, int32_t width);

/* n_frames successive single-frame integrate() calls (the reference's frame loop,
 * clip_seem_fusion.py:305-313) issued from one host call.  With a workspace sized for max_batch >= 2 the
 * frames are fused in windows of max_batch frames (see the window-mode calls above) and
 * K1 + K2 of the next window overlap the feature kernel of the current one; a max_batch = 1 workspace runs
 * frame by frame.  Either way the result equals the frame loop's.  For a contiguous sub-slab volume (x_begin > 0 or
 * x_end < nvox[0], x_span = 0) the call first drops the frames that cannot touch the slab: by the host test above
 * for frames whose pose is passed by value, and for >= 16 survivors additionally by a depth-aware pre-pass kernel
 * (this variant synchronises `stream` once before it launches the windows; SAF_REACH_PREPASS=0 disables it). */
int saf_integrate_sequence(const saf_grid_desc *grid, const saf_volume *vol, const saf_frame *frames,
                           int32_t n_frames, int32_t height, int32_t width, float trunc,
                           int32_t rgb_mode, const saf_workspace *ws, void *stream);

/* argmax_with_check_2d_efficient (clip_seem_fusion.py:315-325): per-voxel argmax of the label
 * histogram (first maximum), -1 where the histogram is all zero.  out: device int64 [n]. */
int saf_label_argmax(const int32_t *labels_one_hot, int64_t n, int32_t n_classes, int64_t *out,
                     void *stream);

/* ---- query: Clip.run_query / Clip.clip_feature_surgery (clipfusion.py:899-934) and the
 *             callers' normalisation (clip_seem_fusion.py:507-511) ---------------------------- */

#define SAF_NORM_NONE        0   /* rows already unit length                                          */
#define SAF_NORM_NAN_TO_NUM  1   /* f/|f|, zero rows -> 0        (clip_seem_fusion.py:507-511)        */
#define SAF_NORM_CLAMP_MIN   2   /* f/max(|f|, 0.1)              (hypersim_eval.py:50-51)             */

#define SAF_SCORE_DOT        0   /* S = F X^T                                                         */
#define SAF_SCORE_SOFTMAX100 1   /* softmax(100*S) over texts    (clipfusion.py:902-903)              */
#define SAF_SCORE_SURGERY    2   /* w_t S[m,t] - mean_s(w_s S[m,s])  (clipfusion.py:913-932)          */

/* Scores of M feature rows against T text embeddings.
 *   feats  device [M, ldf] f32 (first C columns used), text device [T,C] f32,
 *   surgery_w device [T] f32: the class weights w (required for SAF_SCORE_SURGERY),
 *   out    device [M,T] f32.
 * precision: 0 = fp32 CUDA-core FMA, 1 = tf32 tensor cores (tcgen05).  (Exact rankings on the tensor cores come
 * from saf_query_topk, which filters with the tf32 scores and rescoring the survivors in fp32.) */
int saf_query_scores(const float *feats, int64_t M, int32_t C, int64_t ldf, const float *text, int32_t T,
                     int32_t norm_mode, int32_t score_mode, const float *surgery_w, int32_t precision,
                     float *out, void *stream);

/* Top-k rows per text without materialising [M,T]: out_scores/out_index device [T,k]
 * (descending score, ties to the lower row); row indices are offset by index_base
 * (the slab's first global voxel).  ws: device scratch of saf_query_topk_workspace_bytes. */
int saf_query_topk_workspace_bytes(int64_t M, int32_t T, int32_t k, uint64_t *bytes_out);
int saf_query_topk(const float *feats, int64_t M, int32_t C, int64_t ldf, const float *text, int32_t T,
                   int32_t norm_mode, int32_t score_mode, const float *surgery_w, int32_t precision,
                   int32_t k, int64_t index_base, float *out_scores, int64_t *out_index, void *ws,
                   uint64_t ws_bytes, void *stream);

/* Per-row consumers of the scores, computed chunk by chunk (64 Ki rows at a time, the chunk's score block stays
 * in L2) so that the [M,T] matrix is never materialised.  ws: device scratch of saf_query_rows_workspace_bytes(T),
 * 256-byte aligned.
 * saf_query_row_labels: segment() of eval_scannet_segmentation.py:546-561 - out_labels[m, j] = the text with the
 *   j-th largest softmax(100 cos) of row m, j < k (the reference returns the full argsort and its callers read
 *   the first 1 / 5 columns; k = T reproduces it), ties to the lower text index; out_probs (optional, [M,k]) the
 *   softmax values.  T <= 2048.
 * saf_query_text_presence: hypersim_eval.py:80-89 - text = [n_background + n_targets, C]; out[i] = max over rows of
 *   softmax(100 [s_bg.., s_target_i])[-1], the number the reference compares with its thresholds. */
int saf_query_rows_workspace_bytes(int32_t T, uint64_t *bytes_out);
int saf_query_row_labels(const float *feats, int64_t M, int32_t C, int64_t ldf, const float *text, int32_t T,
                         int32_t norm_mode, int32_t precision, int32_t k, int64_t *out_labels, float *out_probs,
                         void *ws, uint64_t ws_bytes, void *stream);
int saf_query_text_presence(const float *feats, int64_t M, int32_t C, int64_t ldf, const float *text,
                            int32_t n_background, int32_t n_targets, int32_t norm_mode, int32_t precision,
                            float *out, void *ws, uint64_t ws_bytes, void *stream);

/* ---- mesh: ClipSeemFusion.extract_mesh (clip_seem_fusion.py:824-888) and ClipFusion.extract_mesh
 *            (clipfusion.py:723-763) ------------------------------------------------------------
 * The reference masks unobserved voxels (weight == 0) to NaN, runs skimage.measure.marching_cubes
 * (level 0) on the host, drops faces with a NaN vertex and the vertices no face uses, then samples
 * rgb / clip_feat (trilinear) and voxel_obj_idx / objects_segmentation_color (nearest) at the
 * vertices with torch grid_sample.  The three calls below do the same on the device, without the
 * host round trip; vertices are ordered by (voxel, axis) of their grid edge and faces by cell.
 * x-slabs: halo_tsdf / halo_weight (device [ny*nz], both or neither) and halo_field (device [ny*nz,
 * channels]) are the NEXT slab's first plane (x = x_end); with them the slab also meshes the cells across
 * the cut and emits the vertices lying in that plane, so that the slabs' meshes join without a gap
 * (spatially_aware_ai_b200/slab.py exchanges the planes and welds the duplicated vertices).  NULL = none. */

#define SAF_SAMPLE_TRILINEAR 0   /* grid_sample mode="bilinear" on the 5-D view */
#define SAF_SAMPLE_NEAREST   1

/* Device scratch for the two calls below: 12 bytes per voxel of the slab (+ small per-CTA arrays). */
int saf_mesh_workspace_bytes(const saf_grid_desc *grid, uint64_t *bytes_out);
/* Pass 1: classify cells, count surviving faces and used vertices.  Synchronises `stream` and
 * returns the counts so that the caller can allocate the outputs.  ws: device, 256-byte aligned. */
int saf_mesh_count(const saf_grid_desc *grid, const float *tsdf, const int32_t *weight,
                   const float *halo_tsdf, const int32_t *halo_weight, void *ws, uint64_t ws_bytes,
                   uint64_t *n_verts_out, uint64_t *n_faces_out, void *stream);
/* Pass 2 (same inputs and workspace as the saf_mesh_count call before it): vertices in voxel-index
 * coordinates [V,3] (global x), optionally also verts * voxel_size + origin (clip_seem_fusion.py:880) and
 * each vertex's grid edge id ((x*ny + y)*nz + z)*3 + axis in global numbering [V] int64 (the identity slabs
 * weld by), and faces [F,3] int64. */
int saf_mesh_emit(const saf_grid_desc *grid, const float *tsdf, const int32_t *weight,
                  const float *halo_tsdf, const int32_t *halo_weight, void *ws, uint64_t ws_bytes,
                  float *verts_out, float *verts_world_out, int64_t *edge_ids_out, int64_t *faces_out,
                  void *stream);
/* grid_sample of a per-voxel field [N,channels] (slab-local rows) at `verts` (index coordinates):
 * grid = (verts + 0.5) / nvox * 2 - 1, align_corners=False, zeros padding, mode SAF_SAMPLE_*;
 * clamp01 != 0 applies .clamp(0, 1) (the colour outputs).  out: device [V,channels]. */
int saf_mesh_sample(const saf_grid_desc *grid, const float *verts, int64_t n_verts, const float *field,
                    const float *halo_field, int32_t channels, int32_t mode, int32_t clamp01, float *out,
                    void *stream);

/* ---- object labelling: flood_fill_3d (handy_utils.py:295-480) without its in-situ classifier -----
 * 26-connected components of equal class id over the per-voxel class grid [nx,ny,nz] (int64, as
 * saf_label_argmax writes it; -1 = unobserved and `null_class` = 133 are background); components with
 * fewer than `min_voxels` (3) voxels are rejected; accepted objects are numbered -2, -3, ... in
 * ascending order of their first voxel (the reference's scan order).  out_obj: device int32 [nx*ny*nz],
 * the reference's voxel_obj_ids (-1 elsewhere).  Synchronises `stream`; *n_objects_out (host) = count. */
int saf_label_components_workspace_bytes(int64_t n_voxels, uint64_t *bytes_out);
int saf_label_components(const int64_t *labels, int32_t nx, int32_t ny, int32_t nz, int32_t null_class,
                         int32_t min_voxels, int32_t *out_obj, void *ws, uint64_t ws_bytes,
                         uint32_t *n_objects_out, void *stream);

/* ---- scene bounds: backproject_pcd (clipfusion.py:510-572) -------------------------------------------
 * nu x nv pixel samples (columns us[nu], rows vs[nv]; the reference uses round(linspace) with 7 each) of every
 * frame back-projected to world space: ray = K^-1 (u, v, 1), point = R (ray * depth) + t.
 *   depth device [F,H,W] f32, poses device [F,4,4] camera->world, k_inverse device [F,3,3],
 *   xyz_out device [F, nv*nu, 3] f32, valid_out device [F, nv*nu] u8 (depth not NaN, > 0, < max_depth). */
int saf_backproject_samples(const float *depth, const float *poses, const float *k_inverse, const int32_t *us,
                            const int32_t *vs, int32_t n_frames, int32_t height, int32_t width, int32_t nu,
                            int32_t nv, float max_depth, float *xyz_out, uint8_t *valid_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SAF_B200_H */
