// GLENet's variance-voting NMS and soft-NMS on the device, for sm_100a.
//
// Replaces the Python loops of pcdet/ops/iou3d_nms/iou3d_nms_utils.py: nms_func (:227-273, the body of new_nms_gpu -- the
// NMS_TYPE of every shipped GLENet config, tools/cfgs/kitti_models/GLENet_VR.yaml:178) and softnms (:312-356).  The
// reference computes an N x N IoU matrix on ONE CPU core (3.4 s for N = 4096) and then iterates in numpy / torch: pick the
// best remaining box, replace it by the variance-weighted average of the boxes that overlap it, decay or zero the scores
// of those boxes.  Every IoU either loop ever looks at is between ORIGINAL boxes (a box is rewritten only in the iteration
// that retires it), so one matrix -- produced by this library's IoU kernel and left in device memory -- serves the whole
// loop, and the loop itself runs here: one CTA per frame, state in shared memory, no host round trip per iteration.
//
// Per iteration (all 1024 threads unless noted):
//   1. arg max of the scores of the boxes still in play (first index among equal scores, numpy / torch argmax);
//   2. column `top` of the IoU matrix -> shared memory (IoU(box_j as a, box_top as b), the reference's ious_all[:, idx]);
//   3. with variances: the boxes with IoU > threshold are compacted IN INDEX ORDER, their weights / variances / coordinates
//      staged in shared memory, and seven lanes (one per box dimension) run the reference's float32 sums SEQUENTIALLY in
//      that order -- numpy's reduction order -- so the voted box differs from the reference's only by expf's last bits;
//   4. score update and retirement, one bit-mask word per warp.
#include "common.cuh"
#include "../../include/glenet_geom.h"
#include <math_constants.h>
#include <atomic>

namespace glenet {

constexpr int VN_THREADS = 1024;
constexpr int VN_STAGE = 512;          // selected boxes staged per pass of the vote
constexpr int VN_MAX_N = 12288;        // scores + IoU column + selection list (12 B / box) + staging must fit 227 KB

enum { VN_HARD = 0, VN_SOFT_GAUSSIAN = 1, VN_SOFT_LINEAR = 2 };

struct VnmsBest { float score; int idx; };
__device__ __forceinline__ VnmsBest vn_better(VnmsBest a, VnmsBest b) {   // higher score; equal scores: the smaller index
    return (b.score > a.score || (b.score == a.score && b.idx < a.idx)) ? b : a;
}

template <int MODE>
__global__ void __launch_bounds__(VN_THREADS, 1)
vnms_kernel(float* __restrict__ boxes_all, float* __restrict__ scores_all, const float* __restrict__ var_all, int var_cols,
            const float* __restrict__ iou_all, int n, float thr, float score_thr, float soft_sigma) {
    extern __shared__ __align__(16) unsigned char vn_smem[];
    float* s_score = reinterpret_cast<float*>(vn_smem);                 // [n]
    float* s_iou = s_score + n;                                          // [n] column `top` of the IoU matrix
    int* s_sel = reinterpret_cast<int*>(s_iou + n);                      // [n] boxes that vote, ascending index
    unsigned int* s_undone = reinterpret_cast<unsigned int*>(s_sel + n); // [ceil(n / 32)]
    float* s_w = reinterpret_cast<float*>(s_undone + (n + 31) / 32);     // [VN_STAGE] exp(-(1 - iou)^2 / 0.05)
    float* s_var = s_w + VN_STAGE;                                       // [VN_STAGE][7]
    float* s_box = s_var + VN_STAGE * 7;                                 // [VN_STAGE][7]
    float* s_term = s_box + VN_STAGE * 7;                                // [VN_STAGE][7] the addends of the current pass
    __shared__ VnmsBest s_best[VN_THREADS / 32];
    __shared__ int s_cnt[VN_THREADS / 32];
    __shared__ int s_top, s_left, s_nsel;
    __shared__ float s_sum[7], s_toph;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x;
    float* boxes = boxes_all + (size_t)frame * n * 7;
    float* scores = scores_all + (size_t)frame * n;
    const float* var = var_all ? var_all + (size_t)frame * n * var_cols : nullptr;
    const float* iou = iou_all + (size_t)frame * n * n;
    constexpr bool SOFT = MODE != VN_HARD;
    constexpr int DIMS = SOFT ? 6 : 7;                                   // softnms votes x, y, z, dx, dy, dz; nms_func also the heading
    const int rounds = (n + VN_THREADS - 1) / VN_THREADS;

    // undone_mask = scores >= score_threshold (:229 / :315)
    for (int k = 0; k < rounds; ++k) {
        const int j = k * VN_THREADS + tid;
        const float sc = j < n ? scores[j] : 0.f;
        if (j < n) s_score[j] = sc;
        const unsigned int word = __ballot_sync(0xffffffffu, j < n && sc >= score_thr);
        if (lane == 0 && k * VN_THREADS + warp * 32 < n) s_undone[(k * VN_THREADS + warp * 32) >> 5] = word;
    }
    __syncthreads();

    for (;;) {
        // ---- 1. the best box still in play, and how many are left
        VnmsBest best = {-CUDART_INF_F, 0x7fffffff};
        int left = 0;
        for (int k = 0; k < rounds; ++k) {
            const int j = k * VN_THREADS + tid;
            if (j < n && ((s_undone[j >> 5] >> (j & 31)) & 1u)) {
                ++left;
                const VnmsBest c = {s_score[j], j};
                best = vn_better(best, c);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            VnmsBest other;
            other.score = __shfl_xor_sync(0xffffffffu, best.score, o);
            other.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
            best = vn_better(best, other);
            left += __shfl_xor_sync(0xffffffffu, left, o);
        }
        if (lane == 0) { s_best[warp] = best; s_cnt[warp] = left; }
        __syncthreads();
        if (warp == 0) {
            best = s_best[lane];
            left = s_cnt[lane];
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                VnmsBest other;
                other.score = __shfl_xor_sync(0xffffffffu, best.score, o);
                other.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
                best = vn_better(best, other);
                left += __shfl_xor_sync(0xffffffffu, left, o);
            }
            if (lane == 0) {
                s_left = left;
                s_top = best.idx;
                // nms_func keeps iterating over boxes whose score has been zeroed (0 < 0 is false, so they never leave the mask);
                // those iterations can neither keep a box nor change a row the caller reads, so the loop ends with the last positive score
                if (!SOFT && score_thr <= 0.f && !(best.score > 0.f)) s_left = 0;
            }
        }
        __syncthreads();
        if (s_left <= (SOFT ? 1 : 0) || s_top >= n) break;   // `while undone_mask.sum() > 0` (:239) / `> 1` (:316)
        const int top = s_top;
        if (SOFT && tid == 0) s_undone[top >> 5] &= ~(1u << (top & 31));   // softnms retires the top box before it gathers (:320)
        if (tid == 0) s_nsel = 0;
        __syncthreads();

        // ---- 2. + 3a. IoU column, and the voters in index order
        int nsel = 0;
        // (a column of the row-major matrix: one sector per element -- all rounds' loads are issued before anything waits on them)
#pragma unroll 4
        for (int k = 0; k < rounds; ++k) {
            const int j = k * VN_THREADS + tid;
            if (j < n && ((s_undone[j >> 5] >> (j & 31)) & 1u)) s_iou[j] = __ldg(iou + (size_t)j * n + top);
        }
        for (int k = 0; k < rounds; ++k) {
            const int j = k * VN_THREADS + tid;
            const bool in_play = j < n && ((s_undone[j >> 5] >> (j & 31)) & 1u);
            const float v = in_play ? s_iou[j] : 0.f;      // (written by this thread above)
            if (var) {
                const bool sel = in_play && v > thr;
                const unsigned int m = __ballot_sync(0xffffffffu, sel);
                if (lane == 0) s_cnt[warp] = __popc(m);
                __syncthreads();
                int before = 0, total = 0;
                for (int w2 = 0; w2 < VN_THREADS / 32; ++w2) { const int c = s_cnt[w2]; total += c; before += w2 < warp ? c : 0; }
                if (sel) s_sel[nsel + before + __popc(m & ((1u << lane) - 1u))] = j;
                nsel += total;
                __syncthreads();
            }
        }
        // ---- 3b. the vote (:246-262 / :327-346): float32 sums in index order, one lane per box dimension
        if (var) {
            const int nvote = SOFT ? nsel + 1 : nsel;              // softnms appends the top box itself (weight 1) as the LAST voter
            if (tid == 0) s_toph = boxes[(size_t)top * 7 + 6];
            for (int pass = 0; pass < 2; ++pass) {                 // pass 0: pi.sum(0); pass 1: (pi / sum * klbox).sum(0)
                float acc = 0.f;
                for (int c0 = 0; c0 < nvote; c0 += VN_STAGE) {
                    const int cn = min(VN_STAGE, nvote - c0);
                    __syncthreads();
                    const bool staged = pass == 1 && nvote <= VN_STAGE;     // one chunk: weights, boxes and variances are still there from pass 0
                    if (!staged)
                    for (int t = tid; t < cn; t += VN_THREADS) {
                        const bool is_top = SOFT && c0 + t == nsel;
                        const int j = is_top ? top : s_sel[c0 + t];
                        const float d = 1.f - s_iou[j];
                        s_w[t] = is_top ? 1.f : expf(__fdiv_rn(-1.f * (d * d), 0.05f));   // std_iou_sigma = 0.05 (:257 / :339)
                    }
                    if (!staged)
                    for (int t = tid; t < cn * 7; t += VN_THREADS) {
                        const int r = t / 7, c = t - r * 7;
                        const int j = (SOFT && c0 + r == nsel) ? top : s_sel[c0 + r];
                        s_box[t] = boxes[(size_t)j * 7 + c];
                        s_var[t] = c < var_cols ? var[(size_t)j * var_cols + c] : 1.f;
                    }
                    __syncthreads();
                    // the terms in parallel (weight / variance, heading rules, normalised product -- each rounded as numpy rounds
                    // it); only the additions have to run in index order
                    {
                        const float top_h = s_toph;
                        for (int t = tid; t < cn * 7; t += VN_THREADS) {
                            const int r = t / 7, c = t - r * 7;
                            if (c >= DIMS) continue;
                            float x = s_box[t];
                            float w = __fdiv_rn(s_w[r], s_var[t]);
                            if (!SOFT && c == 6) {
                                // headings on the far side of the +-pi cut are moved next to the top box (:250-253), and
                                // boxes turned by pi/4 or more do not vote for the heading (:261)
                                if (fabsf(x - top_h) >= 4.712388980384690f) x = top_h > 0.f ? x + 6.283185307179586f : x - 6.283185307179586f;
                                if (fabsf(x - top_h) >= 0.7853981633974483f) w = 0.f;
                            }
                            s_term[t] = pass ? __fmul_rn(__fdiv_rn(w, s_sum[c]), x) : w;
                        }
                    }
                    __syncthreads();
                    if (tid < DIMS) {
                        for (int t = 0; t < cn; ++t) acc = __fadd_rn(acc, s_term[t * 7 + tid]);
                    }
                }
                __syncthreads();
                if (tid < DIMS) {
                    if (pass == 0) s_sum[tid] = acc;
                    else boxes[(size_t)top * 7 + tid] = acc;
                }
                __syncthreads();
            }
        }
        // ---- 4. scores of the boxes in play, retirement (:265-267 / :348-351)
        for (int k = 0; k < rounds; ++k) {
            const int j = k * VN_THREADS + tid;
            bool in_play = j < n && ((s_undone[j >> 5] >> (j & 31)) & 1u);
            if (!SOFT && j == top) in_play = false;                // undone_mask[idx] = False, then the update of the others
            if (in_play) {
                const float v = s_iou[j];
                float scale;
                if (MODE == VN_HARD) scale = v < thr ? 1.f : 0.f;
                else if (MODE == VN_SOFT_GAUSSIAN) scale = expf(__fdiv_rn(-(v * v), soft_sigma));
                else scale = v >= soft_sigma ? 1.f - v : 1.f;
                const float sc = s_score[j] * scale;
                s_score[j] = sc;
                if (sc < score_thr) in_play = false;
            }
            const unsigned int word = __ballot_sync(0xffffffffu, in_play);
            if (lane == 0 && k * VN_THREADS + warp * 32 < n) s_undone[(k * VN_THREADS + warp * 32) >> 5] = word;
        }
        __syncthreads();
    }
    for (int j = tid; j < n; j += VN_THREADS) scores[j] = s_score[j];
}

static size_t vnms_smem_bytes(int n) {
    return sizeof(float) * 2 * (size_t)n + sizeof(int) * (size_t)n + sizeof(unsigned int) * (((size_t)n + 31) / 32) + sizeof(float) * VN_STAGE * 22 + 64;
}

}  // namespace glenet

using namespace glenet;

extern "C" {

int glenet_variance_nms_gpu(float* boxes, float* scores, const float* variance, int var_cols, const float* iou, int frames, int n,
                            float iou_threshold, float score_threshold, int mode, float soft_sigma, glenet_stream_t s) {
    const char* what = "glenet_variance_nms_gpu";
    if (frames < 0 || n < 0 || mode < 0 || mode > 2 || (variance && (var_cols < (mode == VN_HARD ? 7 : 6) || var_cols > 64)))
        return fail(GLENET_EINVAL, "%s: bad argument", what);
    if (frames == 0 || n == 0) return GLENET_OK;
    if (!boxes || !scores || !iou) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (n > VN_MAX_N) return fail(GLENET_EINVAL, "%s: more than 12288 boxes per frame", what);
    const size_t smem = vnms_smem_bytes(n);
    cudaStream_t st = (cudaStream_t)s;
    int rc;
    if (mode == VN_HARD) { auto k = vnms_kernel<VN_HARD>; rc = set_smem(k, smem, what); if (rc) return rc; k<<<frames, VN_THREADS, smem, st>>>(boxes, scores, variance, var_cols, iou, n, iou_threshold, score_threshold, soft_sigma); }
    else if (mode == VN_SOFT_GAUSSIAN) { auto k = vnms_kernel<VN_SOFT_GAUSSIAN>; rc = set_smem(k, smem, what); if (rc) return rc; k<<<frames, VN_THREADS, smem, st>>>(boxes, scores, variance, var_cols, iou, n, iou_threshold, score_threshold, soft_sigma); }
    else { auto k = vnms_kernel<VN_SOFT_LINEAR>; rc = set_smem(k, smem, what); if (rc) return rc; k<<<frames, VN_THREADS, smem, st>>>(boxes, scores, variance, var_cols, iou, n, iou_threshold, score_threshold, soft_sigma); }
    return check_launch(what);
}

}  // extern "C"
