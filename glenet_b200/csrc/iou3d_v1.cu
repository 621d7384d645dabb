// Row-aligned 3D IoU of pcdet/ops/iou3d ("V1" dialect) for sm_100a.
//
// Replaces boxes_aligned_iou3d_gpu of pcdet/ops/iou3d/iou3d_utils.py:332-387 -- boxes3d_to_bev_torch (:79-106, ~10 torch
// elementwise kernels per input), boxes_aligned_overlap_bev_gpu (iou3d.cpp:55-73 -> boxes_aligned_overlap_kernel,
// iou3d_kernel.cu:284-293, 16 threads per block) and another ~25 torch elementwise kernels for the BEV / height / volume
// terms -- by ONE kernel with the same per-step rounding.  Callers: the IoU-aware GLENet heads
// (pcdet/models/dense_heads/anchor_head_kl_label.py:428, anchor_head_iou.py:209), every training step.
//
// One lane per pair; every pair is clipped (predictions vs their own regression targets always overlap), so there is
// no culling and no queue.  Both BoxPre records of a pair live in shared memory and the pair goes through the phased clip
// of clip.cuh in its V1 form (A: branch-free edge + corner tests, B: the warp's crossings pooled one per lane, C: packed-key
// sorting network + fan; more than eight vertices: the whole warp on one pair) -- no local-memory vertex lists.
#include "common.cuh"
#include "geom.cuh"
#include "clip.cuh"
#include "../../include/glenet_geom.h"

namespace glenet {

constexpr int V1_THREADS = 128;
constexpr int V1_BPS = 17;     // record stride: the V1 margins sit in the slots up to BP_ZMAX (geom.cuh: BP1_*); odd => conflict-free

struct V1Smem {
    float2 verts[V1_THREADS * CLIP_SLOTS];
    unsigned int wl[V1_THREADS / 32][32 * CLIP_SLOTS];   // per-warp work lists of phase B (then the scratch of clip_warp_slow)
    float arec[V1_THREADS * V1_BPS], brec[V1_THREADS * V1_BPS];
};

// overlap of this lane's pair (records already at sm.arec / sm.brec + tid * V1_BPS); all lanes of the warp must call
template <bool FMA>
__device__ __forceinline__ float v1_overlap_phased(V1Smem& sm, int tid, bool live) {
    const int warp = tid >> 5;
    const float* a = sm.arec + tid * V1_BPS;
    const float* b = sm.brec + tid * V1_BPS;
    __syncwarp();
    float2* slots = sm.verts + tid * CLIP_SLOTS;
    const unsigned int w = clip_pair_tests<FMA, true, false>(a, b, live);
    const unsigned int hits = clip_hits16(w);
    const int cnt = __popc(hits) + __popc(clip_corners8(w));
    const bool fast = cnt >= 3 && cnt <= CLIP_SLOTS;
    if (fast) clip_write_corners(a, b, w, slots);
    clip_warp_points<FMA>(fast ? hits : 0u, (unsigned int)tid, (unsigned int)tid, sm.wl[warp], sm.arec, sm.brec, V1_BPS, sm.verts + (warp * 32) * CLIP_SLOTS);
    const bool slow = cnt > CLIP_SLOTS;
    const float ov_slow = clip_warp_slow<FMA>(slow, w, (unsigned int)tid, (unsigned int)tid, sm.arec, sm.brec, V1_BPS, reinterpret_cast<float2*>(sm.wl[warp]));
    return slow ? ov_slow : (fast ? clip_area8<FMA>(slots, cnt) : 0.f);
}

// boxes3d_to_bev_torch (iou3d_utils.py:95-105, rect = False): every step is one torch elementwise op on float32
__device__ __forceinline__ void v1_to_bev(const float* __restrict__ box, int wi, int li, float& x1, float& y1, float& x2, float& y2) {
    const float half_w = __fmul_rn(box[wi], 0.5f), half_l = __fmul_rn(box[li], 0.5f);   // x / 2. is exact
    x1 = __fsub_rn(box[0], half_w); y1 = __fsub_rn(box[1], half_l);
    x2 = __fadd_rn(box[0], half_w); y2 = __fadd_rn(box[1], half_l);
}

template <bool FMA>
__global__ void __launch_bounds__(V1_THREADS, 4)
iou3d_v1_aligned_kernel(const float* __restrict__ A, const float* __restrict__ B, int n, int wi, int li, int hi,
                        float* __restrict__ iou3d, float* __restrict__ iou_bev, float* __restrict__ overlap_bev) {
    __shared__ V1Smem sm;
    const int tid = threadIdx.x, i = blockIdx.x * V1_THREADS + tid;
    const bool live = i < n;
    const float* ba = A + (size_t)(live ? i : 0) * 7;
    const float* bb = B + (size_t)(live ? i : 0) * 7;
    if (live) {
        float x1, y1, x2, y2;
        v1_to_bev(ba, wi, li, x1, y1, x2, y2);
        box_prepare_v1<FMA>(x1, y1, x2, y2, device_trig_fused(ba[6]), sm.arec + tid * V1_BPS);
        v1_to_bev(bb, wi, li, x1, y1, x2, y2);
        box_prepare_v1<FMA>(x1, y1, x2, y2, device_trig_fused(bb[6]), sm.brec + tid * V1_BPS);
    }
    const float ov = v1_overlap_phased<FMA>(sm, tid, live);
    if (!live) return;
    if (overlap_bev) overlap_bev[i] = ov;
    if (iou_bev) {   // iou3d_utils.py:351-353
        const float area_a = __fmul_rn(ba[wi], ba[li]), area_b = __fmul_rn(bb[wi], bb[li]);
        iou_bev[i] = __fdiv_rn(ov, clamp_min_nan(__fsub_rn(__fadd_rn(area_a, area_b), ov), 1e-7f));
    }
    if (iou3d) {     // iou3d_utils.py:363-382
        const float hha = __fmul_rn(ba[hi], 0.5f), hhb = __fmul_rn(bb[hi], 0.5f);
        const float va = __fmul_rn(__fmul_rn(ba[3], ba[4]), ba[5]), vb = __fmul_rn(__fmul_rn(bb[3], bb[4]), bb[5]);
        iou3d[i] = iou3d_from_terms(__fsub_rn(ba[2], hha), __fadd_rn(ba[2], hha), va, __fsub_rn(bb[2], hhb), __fadd_rn(bb[2], hhb), vb, ov, 1e-7f);
    }
}

// (N, 5) [x1, y1, x2, y2, angle] x (N, 5) -> overlap; the native call of the reference (boxes_aligned_overlap_bev_gpu)
template <bool FMA>
__global__ void __launch_bounds__(V1_THREADS, 4)
iou3d_v1_aligned_overlap_bev_kernel(const float* __restrict__ A, const float* __restrict__ B, const float4* __restrict__ trigA,
                                    const float4* __restrict__ trigB, int n, float* __restrict__ out) {
    __shared__ V1Smem sm;
    const int tid = threadIdx.x, i = blockIdx.x * V1_THREADS + tid;
    const bool live = i < n;
    if (live) {
        const float* ba = A + (size_t)i * 5;
        const float* bb = B + (size_t)i * 5;
        box_prepare_v1<FMA>(ba[0], ba[1], ba[2], ba[3], trigA ? trigA[i] : device_trig_fused(ba[4]), sm.arec + tid * V1_BPS);
        box_prepare_v1<FMA>(bb[0], bb[1], bb[2], bb[3], trigB ? trigB[i] : device_trig_fused(bb[4]), sm.brec + tid * V1_BPS);
    }
    const float ov = v1_overlap_phased<FMA>(sm, tid, live);
    if (live) out[i] = ov;
}

}  // namespace glenet

using namespace glenet;

extern "C" {

int glenet_iou3d_v1_boxes_aligned_gpu(const float* boxes_a, const float* boxes_b, int n, int w_index, int l_index, int h_index,
                                      float* iou3d, float* iou_bev, float* overlap_bev, glenet_stream_t s) {
    const char* what = "glenet_iou3d_v1_boxes_aligned_gpu";
    if (n < 0) return fail(GLENET_EINVAL, "%s: negative box count", what);
    const int idx[3] = {w_index, l_index, h_index};
    for (int k = 0; k < 3; ++k)
        if (idx[k] < 3 || idx[k] > 5) return fail(GLENET_EINVAL, "%s: w/l/h indices must be a permutation of 3, 4, 5", what);
    if (w_index == l_index || w_index == h_index || l_index == h_index) return fail(GLENET_EINVAL, "%s: w/l/h indices must be a permutation of 3, 4, 5", what);
    if (n == 0) return GLENET_OK;
    if (!boxes_a || !boxes_b || (!iou3d && !iou_bev && !overlap_bev)) return fail(GLENET_EINVAL, "%s: null pointer", what);
    iou3d_v1_aligned_kernel<true><<<(n + V1_THREADS - 1) / V1_THREADS, V1_THREADS, 0, (cudaStream_t)s>>>(
        boxes_a, boxes_b, n, w_index, l_index, h_index, iou3d, iou_bev, overlap_bev);
    return check_launch(what);
}

int glenet_iou3d_v1_aligned_overlap_bev_gpu(const float* boxes_a_bev, const float* boxes_b_bev, int n, float* ans_overlap, glenet_stream_t s) {
    const char* what = "glenet_iou3d_v1_aligned_overlap_bev_gpu";
    if (n < 0) return fail(GLENET_EINVAL, "%s: negative box count", what);
    if (n == 0) return GLENET_OK;
    if (!boxes_a_bev || !boxes_b_bev || !ans_overlap) return fail(GLENET_EINVAL, "%s: null pointer", what);
    iou3d_v1_aligned_overlap_bev_kernel<true><<<(n + V1_THREADS - 1) / V1_THREADS, V1_THREADS, 0, (cudaStream_t)s>>>(
        boxes_a_bev, boxes_b_bev, nullptr, nullptr, n, ans_overlap);
    return check_launch(what);
}

int glenet_iou3d_v1_aligned_overlap_bev_cpu_dialect(const float* boxes_a_bev, const float* trig_a, const float* boxes_b_bev, const float* trig_b,
                                                    int n, float* ans_overlap, glenet_stream_t s) {
    const char* what = "glenet_iou3d_v1_aligned_overlap_bev_cpu_dialect";
    if (n < 0) return fail(GLENET_EINVAL, "%s: negative box count", what);
    if (n == 0) return GLENET_OK;
    if (!boxes_a_bev || !boxes_b_bev || !ans_overlap || !trig_a || !trig_b) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (((uintptr_t)trig_a | (uintptr_t)trig_b) & 15) return fail(GLENET_EALIGN, "%s: trig tables must be 16-byte aligned", what);
    iou3d_v1_aligned_overlap_bev_kernel<false><<<(n + V1_THREADS - 1) / V1_THREADS, V1_THREADS, 0, (cudaStream_t)s>>>(
        boxes_a_bev, boxes_b_bev, reinterpret_cast<const float4*>(trig_a), reinterpret_cast<const float4*>(trig_b), n, ans_overlap);
    return check_launch(what);
}

}  // extern "C"
