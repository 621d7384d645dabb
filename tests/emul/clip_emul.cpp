// CPU walk-through of the phased clip (csrc/clip.cuh) -- TEST INFRASTRUCTURE ONLY, never part of the product.
//
// The device code's phases are driven here by plain loops: for every pair, phase A for the four quad lanes, the quad
// gather as an OR of the shifted bytes, phase B for the four lanes, phase C; pairs with more than eight vertices go
// through clip_slow_pair.  Trigonometry comes in as a table (the caller passes glibc values, which is what the C
// oracle's GPU dialect uses), so the results can be compared with oracle/geom_oracle.c bit for bit.
#include "cuda_shim.h"
#include "../../glenet_b200/csrc/clip.cuh"

using namespace glenet;

extern "C" {

// a, b: (n, 7) boxes; trig_a, trig_b: (n, 4) {cos h, sin h, cos -h, sin -h}; out_overlap, out_iou: (n); out_cnt: (n) vertex count
void emul_clip_aligned(const float* a, const float* trig_a, const float* b, const float* trig_b, int n,
                       float* out_overlap, float* out_iou, int* out_cnt) {
    for (int p = 0; p < n; ++p) {
        float ra[BP_STRIDE], rb[BP_STRIDE];
        const float4 ta = make_float4(trig_a[4 * p], trig_a[4 * p + 1], trig_a[4 * p + 2], trig_a[4 * p + 3]);
        const float4 tb = make_float4(trig_b[4 * p], trig_b[4 * p + 1], trig_b[4 * p + 2], trig_b[4 * p + 3]);
        box_prepare<true>(a + 7 * p, ta, ra);
        box_prepare<true>(b + 7 * p, tb, rb);
        unsigned int w = 0;
        for (int i = 0; i < 4; ++i) w |= clip_edge_tests<true>(ra, rb, i) << (8 * i);
        const int cnt = __popc(clip_hits16(w)) + __popc(clip_corners8(w));
        float ov = 0.f;
        if (cnt >= 3 && cnt <= CLIP_SLOTS) {
            float2 slots[CLIP_SLOTS];
            for (int k = 0; k < CLIP_SLOTS; ++k) slots[k] = make_float2(__builtin_nanf(""), __builtin_nanf(""));   // unused slots must not matter
            for (int q = 0; q < 4; ++q) clip_write_vertices<true>(ra, rb, q, w, slots);
            ov = clip_area8<true>(slots, cnt);
        } else if (cnt > CLIP_SLOTS) {
            float2 sv[CLIP_SLOW_SLOTS];
            float sk[CLIP_SLOW_SLOTS];
            ov = clip_slow_pair<true>(ra, rb, w, sv, sk);
        }
        out_overlap[p] = ov;
        out_iou[p] = iou_from_overlap(ra[BP_AREA], rb[BP_AREA], ov);
        out_cnt[p] = cnt;
    }
}

// the throughput-oriented variant: one lane does all 24 tests and writes the corners, the crossings are computed per work
// item (here simply in order), phase C as above
void emul_clip_aligned_lane(const float* a, const float* trig_a, const float* b, const float* trig_b, int n, float* out_overlap) {
    for (int p = 0; p < n; ++p) {
        float ra[BP_STRIDE], rb[BP_STRIDE];
        box_prepare<true>(a + 7 * p, make_float4(trig_a[4 * p], trig_a[4 * p + 1], trig_a[4 * p + 2], trig_a[4 * p + 3]), ra);
        box_prepare<true>(b + 7 * p, make_float4(trig_b[4 * p], trig_b[4 * p + 1], trig_b[4 * p + 2], trig_b[4 * p + 3]), rb);
        const unsigned int w = clip_pair_tests<true>(ra, rb);
        const unsigned int hits = clip_hits16(w);
        const int cnt = __popc(hits) + __popc(clip_corners8(w));
        float ov = 0.f;
        if (cnt >= 3 && cnt <= CLIP_SLOTS) {
            float2 slots[CLIP_SLOTS];
            for (int k = 0; k < CLIP_SLOTS; ++k) slots[k] = make_float2(__builtin_nanf(""), __builtin_nanf(""));
            clip_write_corners(ra, rb, w, slots);
            int nth = 0;
            for (unsigned int m = hits; m; m &= m - 1, ++nth) {      // what clip_warp_points does with one crossing per lane
                const int e = __ffs((int)m) - 1, i = e >> 2, j = e & 3, i1 = (i + 1) & 3, j1 = (j + 1) & 3;
                slots[nth] = edge_point<true>(ra[BP_PX + i], ra[BP_PY + i], ra[BP_PX + i1], ra[BP_PY + i1], rb[BP_PX + j], rb[BP_PY + j], rb[BP_PX + j1], rb[BP_PY + j1]);
            }
            ov = clip_area8<true>(slots, cnt);
        } else if (cnt > CLIP_SLOTS) {
            float2 sv[CLIP_SLOW_SLOTS];
            float sk[CLIP_SLOW_SLOTS];
            ov = clip_slow_pair<true>(ra, rb, w, sv, sk);
        }
        out_overlap[p] = ov;
    }
}

// the round-1 single-chain clip of geom.cuh on the same inputs (both structures must agree)
void emul_clip_reference_chain(const float* a, const float* trig_a, const float* b, const float* trig_b, int n, float* out_overlap) {
    for (int p = 0; p < n; ++p) {
        float ra[BP_STRIDE], rb[BP_STRIDE];
        box_prepare<true>(a + 7 * p, make_float4(trig_a[4 * p], trig_a[4 * p + 1], trig_a[4 * p + 2], trig_a[4 * p + 3]), ra);
        box_prepare<true>(b + 7 * p, make_float4(trig_b[4 * p], trig_b[4 * p + 1], trig_b[4 * p + 2], trig_b[4 * p + 3]), rb);
        out_overlap[p] = box_overlap_unrolled<true>(ra, rb);
    }
}

// the approximate true overlap of geom.cuh (NMS decision filter) with its error bands
void emul_overlap_approx(const float* a, const float* trig_a, const float* b, const float* trig_b, int n,
                         float* out_approx, float* out_slack, float* out_band, int* out_usable) {
    for (int p = 0; p < n; ++p) {
        float ra[BP_STRIDE], rb[BP_STRIDE];
        const float4 ta = make_float4(trig_a[4 * p], trig_a[4 * p + 1], trig_a[4 * p + 2], trig_a[4 * p + 3]);
        const float4 tb = make_float4(trig_b[4 * p], trig_b[4 * p + 1], trig_b[4 * p + 2], trig_b[4 * p + 3]);
        box_prepare<true>(a + 7 * p, ta, ra);
        box_prepare<true>(b + 7 * p, tb, rb);
        out_approx[p] = overlap_approx(ra, rb);
        overlap_approx_band(ra, rb, out_slack[p], out_band[p]);
        out_usable[p] = overlap_approx_usable(ra, rb) ? 1 : 0;
    }
}

}  // extern "C"
