// Ground-truth database crops for sm_100a: stream compaction behind points-in-boxes.
//
// Replaces the per-object host loops of the GT-database builders,
//   pcdet/datasets/kitti/kitti_dataset.py:248-259   gt_points = points[point_indices[i] > 0]; gt_points[:, :3] -= gt_boxes[i, :3]
//   pcdet/datasets/waymo/waymo_dataset.py:369-380   gt_points = points[box_idxs_of_pts == i];  gt_points[:, :3] -= gt_boxes[i, :3]
// whose output -- float32 rows [x - cx, y - cy, z - cz, features...] per object, written with ndarray.tofile -- is the
// on-disk input format of the CVAE (cvae_uncertainty/dataset.py:313: np.fromfile(..., float32).reshape(-1, 4)).
// The reference ships the whole (boxes x points) int32 mask (KITTI) or the (points,) index vector (Waymo) to the host and
// selects with numpy once per object.  Here the selection never leaves the GPU: one pass counts, one block scans, one pass
// scatters the selected rows, already shifted to the object's centre, into ONE buffer grouped by object (objects ascending,
// points ascending within an object -- numpy's boolean-index order) plus an (objects + 1) offset vector; only that crosses PCIe.
//
// Two selection rules, the two the reference uses:
//   GLENET_CROP_MASK   mask (n_boxes, n_points) int32, row i selects the points with mask > 0 (points_in_boxes_cpu: a point
//                      may belong to several boxes);
//   GLENET_CROP_INDEX  index (n_points,) int32, object i selects the points with index == i (points_in_boxes_gpu: first hit).
// Centre subtraction: numpy evaluates `float32_points -= centres` in the centres' precision (the KITTI infos hold float64
// boxes) and rounds to float32 on the store, so the centres arrive as float64 and the difference is taken in float64.
#include "common.cuh"
#include "../../include/glenet_geom.h"

namespace glenet {

constexpr int CROP_THREADS = 256;
constexpr int CROP_PER_THREAD = 8;
constexpr int CROP_CHUNK = CROP_THREADS * CROP_PER_THREAD;   // points per CTA

template <int MODE>
__device__ __forceinline__ unsigned int crop_flags(const int* __restrict__ sel, int box, long long n_points, long long p0) {
    // bit t of the result = point p0 + t is selected (p0 is a multiple of 8; rows are 16-byte aligned when n_points % 4 == 0)
    const int* row = MODE == 0 ? sel + (size_t)box * n_points : sel;
    unsigned int bits = 0u;
#pragma unroll
    for (int t = 0; t < CROP_PER_THREAD; ++t) {
        const long long p = p0 + t;
        if (p < n_points) {
            const int v = __ldg(row + p);
            if (MODE == 0 ? v > 0 : v == box) bits |= 1u << t;
        }
    }
    return bits;
}

// counts[box * nchunks + chunk] = selected points of `box` in the chunk
template <int MODE>
__global__ void __launch_bounds__(CROP_THREADS)
crop_count_kernel(const int* __restrict__ sel, int n_boxes, long long n_points, int nchunks, int* __restrict__ counts) {
    const int chunk = blockIdx.x, box = blockIdx.y;
    const int tid = threadIdx.x;
    const unsigned int bits = crop_flags<MODE>(sel, box, n_points, (long long)chunk * CROP_CHUNK + tid * CROP_PER_THREAD);
    int c = __popc(bits);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ int warp_sum[CROP_THREADS / 32];
    if ((tid & 31) == 0) warp_sum[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
#pragma unroll
        for (int w = 0; w < CROP_THREADS / 32; ++w) s += warp_sum[w];
        counts[(size_t)box * nchunks + chunk] = s;
    }
}

// exclusive scan of the counts in (box, chunk) order -> chunk_off; offsets[box] = first row of the object, offsets[n_boxes] = total
__global__ void __launch_bounds__(1024)
crop_scan_kernel(const int* __restrict__ counts, int n_boxes, int nchunks, long long* __restrict__ chunk_off, long long* __restrict__ offsets) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long total = (long long)n_boxes * nchunks;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < total; base += 1024) {
        const long long i = base + tid;
        const long long v = i < total ? counts[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            warp_tot[lane] = wi - w;   // exclusive prefix of the warp totals
        }
        __syncthreads();
        const long long excl = carry + warp_tot[warp] + incl - v;
        if (i < total) {
            chunk_off[i] = excl;
            if (i % nchunks == 0) offsets[i / nchunks] = excl;
        }
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) offsets[n_boxes] = carry;
}

// rows of the selected points, shifted by the object's centre, to crops[(chunk_off + rank) * features ...]
template <int MODE>
__global__ void __launch_bounds__(CROP_THREADS)
crop_scatter_kernel(const int* __restrict__ sel, int n_boxes, long long n_points, int nchunks, const long long* __restrict__ chunk_off,
                    const float* __restrict__ points, int features, const double* __restrict__ centres, long long capacity, float* __restrict__ crops) {
    const int chunk = blockIdx.x, box = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long p0 = (long long)chunk * CROP_CHUNK + tid * CROP_PER_THREAD;
    const unsigned int bits = crop_flags<MODE>(sel, box, n_points, p0);
    const int c = __popc(bits);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __shared__ int warp_tot[CROP_THREADS / 32];
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = incl - c;
#pragma unroll
    for (int w = 0; w < CROP_THREADS / 32; ++w) before += w < warp ? warp_tot[w] : 0;
    if (!bits) return;
    const double cx = centres[3 * box], cy = centres[3 * box + 1], cz = centres[3 * box + 2];
    long long row = chunk_off[(size_t)box * nchunks + chunk] + before;
    for (unsigned int m = bits; m; m &= m - 1, ++row) {
        if (row >= capacity) break;   // the caller sees offsets[n_boxes] > capacity and retries with a larger buffer
        const float* src = points + (size_t)(p0 + (__ffs(m) - 1)) * features;
        float* dst = crops + (size_t)row * features;
        dst[0] = (float)((double)src[0] - cx);
        dst[1] = (float)((double)src[1] - cy);
        dst[2] = (float)((double)src[2] - cz);
        for (int f = 3; f < features; ++f) dst[f] = src[f];
    }
}

}  // namespace glenet

using namespace glenet;

extern "C" {

size_t glenet_gt_crop_workspace_bytes(int n_boxes, long long n_points) {
    if (n_boxes <= 0 || n_points <= 0) return 256;
    const long long nchunks = (n_points + CROP_CHUNK - 1) / CROP_CHUNK;
    return align_up((size_t)n_boxes * nchunks * 4, 256) + align_up((size_t)n_boxes * nchunks * 8, 256);
}

int glenet_gt_crop_gpu(int mode, const int32_t* selection, const float* points, long long n_points, int features, const double* centres,
                       int n_boxes, long long capacity, long long* offsets, float* crops, void* workspace, size_t workspace_bytes,
                       glenet_stream_t s) {
    const char* what = "glenet_gt_crop_gpu";
    cudaStream_t st = (cudaStream_t)s;
    if (mode != GLENET_CROP_MASK && mode != GLENET_CROP_INDEX) return fail(GLENET_EINVAL, "%s: unknown selection mode", what);
    if (n_boxes < 0 || n_points < 0 || capacity < 0 || features < 3) return fail(GLENET_EINVAL, "%s: negative count or fewer than 3 features", what);
    if (!offsets) return fail(GLENET_EINVAL, "%s: null offsets", what);
    if (n_boxes == 0 || n_points == 0) {
        cudaError_t e = cudaMemsetAsync(offsets, 0, (size_t)(n_boxes + 1) * sizeof(long long), st);
        return e == cudaSuccess ? GLENET_OK : fail(-(int)e, "%s: cudaMemsetAsync failed", what);
    }
    if (!selection || !points || !centres || (!crops && capacity > 0) || !workspace) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (n_boxes > 65535) return fail(GLENET_EINVAL, "%s: more than 65535 boxes", what);
    if (workspace_bytes < glenet_gt_crop_workspace_bytes(n_boxes, n_points)) return fail(GLENET_EWORKSPACE, "%s: workspace too small", what);
    if ((uintptr_t)workspace & 15) return fail(GLENET_EALIGN, "%s: workspace must be 16-byte aligned", what);
    const long long nchunks = (n_points + CROP_CHUNK - 1) / CROP_CHUNK;
    if (nchunks > 0x7fffffffLL) return fail(GLENET_EINVAL, "%s: too many points", what);
    int* counts = reinterpret_cast<int*>(workspace);
    long long* chunk_off = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(workspace) + align_up((size_t)n_boxes * nchunks * 4, 256));
    const dim3 grid((unsigned)nchunks, (unsigned)n_boxes);
    if (mode == GLENET_CROP_MASK) crop_count_kernel<0><<<grid, CROP_THREADS, 0, st>>>(selection, n_boxes, n_points, (int)nchunks, counts);
    else crop_count_kernel<1><<<grid, CROP_THREADS, 0, st>>>(selection, n_boxes, n_points, (int)nchunks, counts);
    int rc = check_launch(what);
    if (rc) return rc;
    crop_scan_kernel<<<1, 1024, 0, st>>>(counts, n_boxes, (int)nchunks, chunk_off, offsets);
    rc = check_launch(what);
    if (rc) return rc;
    if (capacity == 0) return GLENET_OK;   // sizing call: offsets only
    if (mode == GLENET_CROP_MASK) crop_scatter_kernel<0><<<grid, CROP_THREADS, 0, st>>>(selection, n_boxes, n_points, (int)nchunks, chunk_off, points, features, centres, capacity, crops);
    else crop_scatter_kernel<1><<<grid, CROP_THREADS, 0, st>>>(selection, n_boxes, n_points, (int)nchunks, chunk_off, points, features, centres, capacity, crops);
    return check_launch(what);
}

}  // extern "C"
