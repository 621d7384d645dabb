/*
 * glenet_geom.h -- C ABI of libglenet_geom.so: the B200-native (sm_100a) rotated-box
 * geometry hot path of GLENet / OpenPCDet (pairwise rotated BEV / 3D IoU, bitmask NMS,
 * points-in-boxes).
 *
 * Every entry point replaces one function of the reference's two pybind11 FFI modules
 * (`iou3d_nms_cuda`, `roiaware_pool3d_cuda`); the reference interface each one stands in
 * for is cited as file:line relative to the reference tree.  Conventions:
 *
 *   - boxes are float32 rows [x, y, z, dx, dy, dz, heading], row-major, contiguous;
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns inputs, outputs and workspaces; the library never allocates,
 *     frees or synchronises (all work is enqueued on `stream`) -- the one exception is the
 *     explicit glenet_symm_* allocator of the multi-GPU exchange windows;
 *   - return value: 0 on success, a negative code on failure (-(cudaError_t) for CUDA
 *     errors, <= -1000 for argument errors); glenet_last_error() gives the text;
 *   - n == 0 is legal everywhere and launches nothing (the reference prints a launch
 *     error instead, callers guard: pcdet/models/model_utils/model_nms_utils.py:26).
 *
 * The reference reports errors with fprintf + exit(-1) (iou3d_nms.cpp:14-38); this
 * library returns codes instead so that the host wrapper can raise.
 */
#ifndef GLENET_GEOM_H
#define GLENET_GEOM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* glenet_stream_t; /* == cudaStream_t */

/* ABI version (bumped on any signature change) and last error text of this thread. */
int glenet_abi_version(void);
const char* glenet_last_error(void);

/* ---------------------------------------------------------------- rotated IoU
 * boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap)   pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:49-68
 * boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou)           pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:70-88
 * out is (na, nb) float32; every element is written (no pre-zeroing needed).
 * GPU dialect: libdevice trig + the reference kernels' FMA contraction. */
int glenet_boxes_overlap_bev_gpu(const float* boxes_a, int na, const float* boxes_b, int nb,
                                 float* ans_overlap, glenet_stream_t stream);
int glenet_boxes_iou_bev_gpu(const float* boxes_a, int na, const float* boxes_b, int nb,
                             float* ans_iou, glenet_stream_t stream);

/* boxes_iou3d_gpu(boxes_a, boxes_b): the reference composes it in Python from
 * boxes_overlap_bev_gpu plus ~10 elementwise torch kernels
 * (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:88-121); here it is one fused kernel with the
 * same per-step rounding. */
int glenet_boxes_iou3d_gpu(const float* boxes_a, int na, const float* boxes_b, int nb,
                           float* ans_iou3d, glenet_stream_t stream);

/* Frame-batched variant: `frames` independent (na, nb) problems in ONE launch, frame f reading
 * boxes_a + f * a_frame_stride and boxes_b + f * b_frame_stride (strides in floats; 0 = the same
 * boxes for every frame) and writing the dense (na, nb) block out + f * na * nb.
 * This is the loop `for k in range(batch_size)` of the target assigner
 * (pcdet/models/dense_heads/target_assigner/axis_aligned_target_assigner.py:60-105: the same anchors
 * against each frame's padded GT boxes) as one grid: the tiles of all frames share the SMs, so the
 * per-launch latency chain of a single frame no longer bounds the throughput.
 * mode 0 = overlap, 1 = BEV IoU, 2 = 3D IoU; same arithmetic as the single-frame entry points. */
int glenet_boxes_iou_frames_gpu(int mode, const float* boxes_a, long long a_frame_stride, int na,
                                const float* boxes_b, long long b_frame_stride, int nb,
                                float* out, int frames, glenet_stream_t stream);

/* Sparse variant of the frame-batched call: the (frames, na, nb) matrix is NOT materialised; every element that
 * is not exactly +0.0 (NaN included) is appended to a coordinate list instead,
 *     idx[k] = f * na * nb + row * nb + col   (int64),   val[k] = the value the dense call would store there,
 * in no particular order.  *count (device, zeroed by the call on `stream`) ends up as the number of such elements
 * and keeps counting past `cap`, so count > cap tells the caller to retry with larger buffers.
 * For the consumers of the reference that only reduce the matrix -- row / column max + argmax of the target
 * assigner (axis_aligned_target_assigner.py:141-165), RoI sampling (proposal_target_layer.py:113-114), recall
 * counting (detector3d_template.py:344-359) -- this removes the 4 B / pair HBM-write bound of the dense call. */
int glenet_boxes_iou_frames_sparse_gpu(int mode, const float* boxes_a, long long a_frame_stride, int na,
                                       const float* boxes_b, long long b_frame_stride, int nb, int frames,
                                       long long* idx, float* val, long long cap, unsigned long long* count,
                                       glenet_stream_t stream);

/* Reduced variant: per frame, the maximum of every row and of every column of the (na, nb) matrix together with the
 * FIRST index attaining it (numpy's argmax rule, which is what the reference's assigner applies), again without
 * materialising the matrix.  row_key: (frames, na) u64, col_key: (frames, nb) u64, both zeroed by the call on `stream`;
 * key = (IEEE bits of the maximum << 32) | (0xffffffff - argmax); key == 0 means "no non-zero element": max 0, argmax 0. */
int glenet_boxes_iou_frames_max_gpu(int mode, const float* boxes_a, long long a_frame_stride, int na,
                                    const float* boxes_b, long long b_frame_stride, int nb, int frames,
                                    unsigned long long* row_key, unsigned long long* col_key, glenet_stream_t stream);

/* The two key arrays of glenet_boxes_iou_frames_max_gpu turned into (max, argmax) vectors by ONE launch: n_row = frames * na,
 * n_col = frames * nb; max float32, argmax int64 (first index among equal maxima; 0 where nothing overlaps).  The key
 * arrays are left zeroed. */
int glenet_iou_keys_decode_gpu(unsigned long long* row_key, long long n_row, unsigned long long* col_key, long long n_col,
                               float* row_max, long long* row_arg, float* col_max, long long* col_arg, glenet_stream_t stream);

/* ---------------------------------------------------------------- multi-GPU: the row-sharded sweep (additive)
 * The reference never shards its geometry ops; its consumer of the big matrices is the target assigner
 * (pcdet/models/dense_heads/target_assigner/axis_aligned_target_assigner.py:132-165), which needs per frame the row maxima of
 * the anchors and the column maxima + first row over ALL anchors.  Here rank r of `world` (one process per GPU of one NVLink /
 * NVSwitch box) computes rows [row_offset, row_offset + na) of the (na_total, nb) matrix of every frame; results move between
 * GPUs inside the IoU kernel -- peer stores / system-scope atomics into "exchange windows" mapped through CUDA IPC -- not
 * through NCCL.
 *
 * Symmetric memory.  glenet_symm_alloc is the ONE place where the library allocates (CUDA IPC exports whole allocations);
 * the block is zeroed.  export/import move the 64-byte IPC handle between the processes (any transport, e.g.
 * torch.distributed); import maps a peer's block into this process.  These five calls synchronise like cudaMalloc does. */
int glenet_symm_alloc(size_t bytes, void** dev_ptr);
int glenet_symm_free(void* dev_ptr);
int glenet_symm_export(const void* dev_ptr, unsigned char* handle64);
int glenet_symm_import(const unsigned char* handle64, void** peer_ptr);
int glenet_symm_unmap(void* peer_ptr);
/* Size of one rank's exchange window for problems of up to `frames` x `nb` column keys and coordinate lists of `list_cap`
 * entries per source rank (0: no gather).  All ranks must use the same three numbers. */
size_t glenet_exchange_window_bytes(int frames, int nb, long long list_cap);
/* Diagnostic (synchronises the device): error bits left in the LOCAL window by the consumer kernels -- 1 = gave up waiting
 * for a peer's flag (~4 s), 2 = a coordinate list overflowed list_cap. */
int glenet_exchange_status(const void* window_local, unsigned int* status_host);

/* assign: this rank's slab of the IoU matrix (out: (frames, na, nb), or NULL to skip the matrix) plus the assigner's
 * reductions -- row_max / row_arg (frames, na): max over the columns and first column attaining it; col_max / col_arg
 * (frames, nb): max over the rows OF ALL RANKS and the first GLOBAL row attaining it (identical on every rank).
 * windows: host array of `world` device pointers, windows[rank] = the local window, the others IPC-mapped.
 * row_key: (frames, na) u64 scratch, zeroed before the first call (the call leaves it zeroed).
 * step: 1, 2, 3, ... -- the same number on every rank for the same collective call.  The call enqueues the IoU kernel
 * (which pushes its column keys into the peers' windows and raises a flag) and a decode kernel that waits for all flags. */
int glenet_boxes_iou_frames_assign_gpu(int mode, const float* boxes_a, long long a_frame_stride, int na,
                                       const float* boxes_b, long long b_frame_stride, int nb, int frames,
                                       float* out, int row_offset, long long na_total, unsigned long long* row_key,
                                       float* row_max, long long* row_arg, float* col_max, long long* col_arg,
                                       int world, int rank, void* const* windows, long long list_cap, unsigned int step,
                                       glenet_stream_t stream);
/* gather: the WHOLE (frames, na_total, nb) matrix on every rank.  Each rank zero-fills its own copy and the kernel ships
 * only the non-zero elements of its slab -- (flat index, value) entries stored into every peer's window from the clip
 * epilogue; a scatter kernel applies all ranks' lists once their flags are up.  An anchor sweep is > 99 % zeros: ~14 MB
 * travel instead of the 1.2 GB of an all-gather.
 * zero_fill: 1 = memset out_full, IoU kernel, scatter, all on `stream`; 0 = the same without the memset (the caller has
 * zeroed out_full); 2 = IoU kernel only and 3 = scatter only -- the two halves of one step (same `step`), so that a caller
 * can run the fill on a second stream under the IoU kernel and make only the scatter wait for it. */
int glenet_boxes_iou_frames_gather_gpu(int mode, const float* boxes_a, long long a_frame_stride, int na,
                                       const float* boxes_b, long long b_frame_stride, int nb, int frames,
                                       float* out_full, int zero_fill, int row_offset, long long na_total,
                                       int world, int rank, void* const* windows, long long list_cap, unsigned int step,
                                       glenet_stream_t stream);

/* Row-aligned variants: out[i] = f(boxes_a[i], boxes_b[i / group]) for i < na, where boxes_b
 * holds ceil(na / group) rows.  Additive API for the CVAE label-uncertainty workload
 * (30 sampled boxes per GT object); mode 0 = overlap, 1 = BEV IoU, 2 = 3D IoU.
 * Same arithmetic as the pairwise entry points. */
int glenet_boxes_iou_aligned_gpu(int mode, const float* boxes_a, int na, const float* boxes_b,
                                 int group, float* out, glenet_stream_t stream);

/* ---------------------------------------------------------------- pcdet/ops/iou3d (row-aligned IoU of the IoU-aware heads)
 * boxes_aligned_iou3d_gpu(boxes_a, boxes_b, box_mode, rect=False, need_bev)   pcdet/ops/iou3d/iou3d_utils.py:332-387
 *   = boxes3d_to_bev_torch (:79-106) + boxes_aligned_overlap_bev_gpu (pcdet/ops/iou3d/src/iou3d.cpp:55-73,
 *     kernel iou3d_kernel.cu:284-293) + ~25 torch elementwise kernels, here one kernel with the same per-step rounding.
 * boxes_a, boxes_b: (n, 7) [x, y, z, d3, d4, d5, ry]; w_index / l_index / h_index (a permutation of 3, 4, 5) say
 * which of d3..d5 is the extent along x, along y and the height ('wlh' -> 3, 4, 5).  Each of the three outputs may
 * be NULL: iou3d (n), iou_bev (n), overlap_bev (n).  This op has its own arithmetic (clockwise rotation,
 * [x1, y1, x2, y2] edges -/+ 1e-5 in check_in_box2d), reproduced from the SASS of the reference kernel. */
int glenet_iou3d_v1_boxes_aligned_gpu(const float* boxes_a, const float* boxes_b, int n, int w_index, int l_index,
                                      int h_index, float* iou3d, float* iou_bev, float* overlap_bev,
                                      glenet_stream_t stream);
/* boxes_aligned_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap) itself: (n, 5) [x1, y1, x2, y2, ry] rows. */
int glenet_iou3d_v1_aligned_overlap_bev_gpu(const float* boxes_a_bev, const float* boxes_b_bev, int n,
                                            float* ans_overlap, glenet_stream_t stream);
/* The same in the CPU dialect of pcdet/ops/iou3d/src/iou3d_cpu.cpp (box_overlap :126-247; no FMA contraction, host
 * libm trig tables as for glenet_boxes_iou_bev_cpu_dialect): row i of boxes_overlap_bev_cpu's diagonal. */
int glenet_iou3d_v1_aligned_overlap_bev_cpu_dialect(const float* boxes_a_bev, const float* trig_a,
                                                    const float* boxes_b_bev, const float* trig_b, int n,
                                                    float* ans_overlap, glenet_stream_t stream);

/* boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou)           pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:232-252
 * The reference runs this single-threaded on the host.  Here it executes on the GPU in the
 * "CPU dialect" (no FMA contraction; cos/sin supplied by the host's libm so that the
 * 0.01 m margin predicate of check_in_box2d, iou3d_cpu.cpp:74-84, is decided bit-identically).
 * trig_a / trig_b: (n, 4) float32 rows {cosf(h), sinf(h), cosf(-h), sinf(-h)}, device pointers. */
int glenet_boxes_iou_bev_cpu_dialect(const float* boxes_a, const float* trig_a, int na,
                                     const float* boxes_b, const float* trig_b, int nb,
                                     float* ans_iou, glenet_stream_t stream);

/* ---------------------------------------------------------------- NMS
 * nms_gpu(boxes, keep, thresh)        pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:90-136  (rotated)
 * nms_normal_gpu(boxes, keep, thresh) pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:139-186 (axis aligned)
 *
 * boxes: (frames, n, 7) already sorted by descending score within each frame.
 * keep:  (frames, n) int64, device (the reference's is a host tensor); the first
 *        num_keep[f] entries of row f are the kept indices in ascending order.
 * num_keep: (frames) int32, device.
 * The 64-bit suppression mask and the greedy sweep both stay on the device: no cudaMalloc,
 * no D2H copy of the mask, no host loop.  workspace must hold
 * glenet_nms_workspace_bytes(frames, n) bytes, 16-byte aligned. */
size_t glenet_nms_workspace_bytes(int frames, int n);
int glenet_nms_gpu(const float* boxes, int frames, int n, float nms_overlap_thresh,
                   int64_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes,
                   glenet_stream_t stream);
int glenet_nms_normal_gpu(const float* boxes, int frames, int n, float nms_overlap_thresh,
                          int64_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes,
                          glenet_stream_t stream);

/* GLENet's variance-voting NMS and soft-NMS: the Python loops of nms_func (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:227-273, the
 * body of new_nms_gpu :200-224, NMS_TYPE of every shipped GLENet config) and softnms (:312-356) as one kernel per call, one
 * CTA per frame.  boxes (frames, n, 7) and scores (frames, n) are updated IN PLACE: a retired box is replaced by the
 * variance-weighted average of the boxes that overlap it by more than iou_threshold (variance: (frames, n, var_cols) or NULL
 * for no voting; var_cols >= 7 for mode 0, >= 6 otherwise), scores are zeroed (mode 0), decayed by exp(-iou^2 / soft_sigma)
 * (mode 1) or by 1 - iou where iou >= soft_sigma (mode 2).  iou: (frames, n, n), element [j][i] = IoU(box j as a, box i as b)
 * of the ORIGINAL boxes (glenet_boxes_iou_bev_cpu_dialect for new_nms_gpu, glenet_boxes_iou_bev_gpu for softnms), resident on
 * the device.  The float32 sums of the vote run in index order, as numpy's do.  n <= 12288. */
int glenet_variance_nms_gpu(float* boxes, float* scores, const float* variance, int var_cols, const float* iou, int frames, int n,
                            float iou_threshold, float score_threshold, int mode, float soft_sigma, glenet_stream_t stream);

/* ---------------------------------------------------------------- points in boxes
 * points_in_boxes_gpu(boxes, pts, box_idx_of_points)  pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:98-118
 * boxes (B, N, 7), pts (B, M, 3), box_idx_of_points (B, M) int32: index of the first box
 * containing the point, -1 if none.  Every element is written (no pre-fill needed).
 * workspace: glenet_points_in_boxes_workspace_bytes(B, N) bytes, 16-byte aligned. */
size_t glenet_points_in_boxes_workspace_bytes(int batch, int boxes_num);
int glenet_points_in_boxes_gpu(const float* boxes, const float* pts, int batch, int boxes_num,
                               int pts_num, int32_t* box_idx_of_points, void* workspace,
                               size_t workspace_bytes, glenet_stream_t stream);

/* points_in_boxes_cpu(boxes, pts, pts_indices)        pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:143-168
 * boxes (N, 7), pts (M, 3), pts_indices (N, M) int32 0/1 (MARGIN = 1e-2, no early exit).
 * Executed on the GPU in the CPU dialect; trig: (N, 2) float32 rows {cosf(-h), sinf(-h)}
 * evaluated by the host's libm. */
int glenet_points_in_boxes_cpu_dialect(const float* boxes, const float* trig, int boxes_num,
                                       const float* pts, int pts_num, int32_t* pts_indices,
                                       glenet_stream_t stream);

/* ---------------------------------------------------------------- GT-database crops (next scope row)
 * The per-object selection + centring loop of the GT-database builders,
 *   pcdet/datasets/kitti/kitti_dataset.py:248-259    gt_points = points[point_indices[i] > 0]; gt_points[:, :3] -= gt_boxes[i, :3]
 *   pcdet/datasets/waymo/waymo_dataset.py:369-380    gt_points = points[box_idxs_of_pts == i];  gt_points[:, :3] -= gt_boxes[i, :3]
 * as stream compaction on the device.  selection: GLENET_CROP_MASK -> (n_boxes, n_points) int32 mask of
 * glenet_points_in_boxes_cpu_dialect (object i takes the points with mask > 0); GLENET_CROP_INDEX -> (n_points) int32 of
 * glenet_points_in_boxes_gpu for one frame (object i takes the points with index == i).  points: (n_points, features) float32,
 * features >= 3, xyz first; centres: (n_boxes, 3) FLOAT64 (numpy subtracts in the boxes' precision and rounds to float32).
 * Output: offsets (n_boxes + 1) int64, crops (capacity, features) float32 -- rows offsets[i] .. offsets[i + 1] are object i's
 * points in ascending point order, xyz relative to its centre: the bytes ndarray.tofile writes to the object's .bin
 * (cvae_uncertainty/dataset.py:313 reads them back).  Rows beyond `capacity` are dropped (offsets stay exact, so the caller
 * can retry); capacity == 0 is a sizing call.  workspace: glenet_gt_crop_workspace_bytes() bytes, 16-byte aligned. */
enum { GLENET_CROP_MASK = 0, GLENET_CROP_INDEX = 1 };
size_t glenet_gt_crop_workspace_bytes(int n_boxes, long long n_points);
int glenet_gt_crop_gpu(int mode, const int32_t* selection, const float* points, long long n_points, int features,
                       const double* centres, int n_boxes, long long capacity, long long* offsets, float* crops,
                       void* workspace, size_t workspace_bytes, glenet_stream_t stream);

/* ---------------------------------------------------------------- KITTI evaluator's rotated IoU (next scope row)
 * rotate_iou_gpu_eval(boxes, query_boxes, criterion)   pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py:263-330
 * (a numba-CUDA kernel in the reference; its callers: kitti_object_eval_python/eval.py:115-151).
 * boxes (n, 5), query_boxes (k, 5): rows [x, y, x_d, y_d, angle], angle clockwise when positive; iou (n, k) float32, every
 * element written: iou[i][j] = devRotateIoUEval(query_boxes[j], boxes[i]): criterion -1 intersection / union,
 * 0 intersection / area(query box), 1 intersection / area(box), anything else the intersection area.
 * Same float32 / float64 typing and the same FMA contraction as the kernel numba compiles for sm_100a. */
int glenet_rotate_iou_eval_gpu(const float* boxes, int n, const float* query_boxes, int k, int criterion, float* iou,
                               glenet_stream_t stream);
/* Block-diagonal form (additive): group g pairs boxes[box_offsets[g] .. box_offsets[g+1]) with
 * query_boxes[query_offsets[g] .. query_offsets[g+1]) and writes its dense (nb_g, nq_g) block at iou + out_offsets[g] --
 * the per-frame blocks calculate_iou_partly (eval.py:383-397) slices out of the dense matrix of one evaluation part.
 * Offsets are device arrays (groups + 1 entries; int32, int32, int64); max_boxes / max_queries bound the group sizes. */
int glenet_rotate_iou_eval_blocks_gpu(const float* boxes, const int* box_offsets, const float* query_boxes, const int* query_offsets,
                                      const long long* out_offsets, int groups, int max_boxes, int max_queries, int criterion,
                                      float* iou, glenet_stream_t stream);

/* iou3d(gboxes, qboxes)                                cvae_uncertainty/eval_utils/eval_utils.py:14-65
 * The recall IoU of the CVAE evaluation (:219-229): row-aligned 3D IoU of n (ground truth, prediction) pairs, boxes
 * [x, y, z, w, l, h, ry] float32, clamped to +-200 as the reference does.  The reference evaluates the BEV overlap with
 * Python loops over numpy float32 scalars on the host (pcdet/utils/loss_utils.py:276-411,551-635); this is one kernel in
 * that dialect (no FMA contraction, descending-angle vertex order, float32 fan sum).  ious: (n) float32. */
int glenet_cvae_iou3d_gpu(const float* gboxes, const float* qboxes, int n, float* ious, glenet_stream_t stream);

/* ---------------------------------------------------------------- host helpers (CPU dialect)
 * Per-box trigonometry evaluated by the HOST's libm, exactly the calls the reference's CPU
 * code makes (iou3d_cpu.cpp:74-84,146-151; roiaware_pool3d.cpp:121-125).  boxes_host: (n, 7)
 * host floats.  trig4 rows: {cosf(h), sinf(h), cosf(-h), sinf(-h)}; trig2 rows: {cosf(-h), sinf(-h)}. */
void glenet_host_trig4(const float* boxes_host, int n, float* out_host);
/* the same table for angles stored with an arbitrary stride (in floats), e.g. column 4 of (n, 5) [x1, y1, x2, y2, ry] rows */
void glenet_host_trig4_strided(const float* angles_host, int stride, int n, float* out_host);
void glenet_host_trig2(const float* boxes_host, int n, float* out_host);

#ifdef __cplusplus
}
#endif
#endif /* GLENET_GEOM_H */
