// Phased rotated-rectangle clip: box_overlap (iou3d_nms_kernel.cu:104-225) restructured so that the polygon never
// lives in per-thread local memory and the lanes of a warp stay busy.
//
// The reference (and round 1 of this repo) clips one pair per thread in one long, divergent chain: 16 edge tests with
// early-outs, two IEEE divisions per crossing found, a vertex list that is indexed dynamically (=> local memory) and a
// sort.  Here a batch of pairs, staged as BoxPre records in shared memory, goes through three data-parallel phases:
//
//   A  one lane per pair: the 16 edge tests run BRANCH-FREE (same expressions as edge_crosses, the early-outs become
//      ANDs) together with the eight corner-in-box tests -> 24 result bits.  The vertex positions follow from the
//      bits alone (slot = popcount of the result bits below, in the reference's discovery order), so the admitted
//      corners are written to the pair's eight SHARED-MEMORY slots at once.
//      (clip_edge_tests is also usable with four lanes per pair, lane i owning edge i of box a, plus clip_quad_gather
//      and clip_write_vertices -- the latency-oriented variant.)
//   B  the crossings found (usually 2-4 per pair, at most 8 on the fast path) are the expensive part -- two IEEE
//      divisions each -- and the reference computes them with ~5 of 32 lanes active.  Here the 32 pairs of a warp pool
//      their crossings in a warp-local work list (prefix sum of the counts, no atomics) and the warp computes them one
//      crossing per lane, full warps at a time, each point going straight to its pair's slot.
//   C  one lane per pair: centroid, one pseudo-angle key per vertex, keys packed as (monotone key bits | slot) into
//      ONE register each, a 19-comparator network of min/max pairs on those registers (2 instructions per comparator
//      instead of 7 when coordinates travel with the keys), vertices re-read in sorted order, the reference's fan.
//   Pairs with more than eight vertices (corners admitted by the 0.01 m margin next to a crossing: ~1 % of dense
//   workloads, < 1e-4 of an anchor sweep) are set aside and finished by clip_slow_pair, which keeps up to 16 vertices
//   in a shared-memory scratch slot -- still no local memory.
//
// The arithmetic of every value that reaches the result (crossing points, margin predicate, fan terms and their
// summation order) is the pinned arithmetic of geom.cuh; only the ORDER in which work is done changed.  Vertex order
// between directions that differ by a few ulps may differ from the reference (as in round 1: monotone pseudo-angle
// instead of atan2f; now additionally the three lowest key bits carry the slot, which reproduces the reference's
// stable tie-break on discovery order).
#pragma once
#include "geom.cuh"

namespace glenet {

constexpr int CLIP_SLOTS = 8;          // vertex slots of the fast path
constexpr int CLIP_SLOW_SLOTS = 16;    // cross_points[16] of the reference (:155)

// ---------------------------------------------------------------- phase A
// Result bits of quad lane i: bit j (0..3) = edge i of a crosses edge j of b (intersection() would return 1),
// bit 4 = corner i of b is inside a, bit 5 = corner i of a is inside b (check_in_box2d, MARGIN).
// WARP_SKIP: all 32 lanes call with their own pair (live = false for lanes without one) and an edge pair whose bounding
// boxes are disjoint in EVERY lane of the warp -- typically the opposite sides of two similar boxes -- is skipped.
template <bool FMA, bool V1 = false, bool WARP_SKIP = false>
__device__ __forceinline__ unsigned int clip_edge_tests(const float* __restrict__ a, const float* __restrict__ b, int i, bool live = true) {
    const int i1 = (i + 1) & 3;
    const float p0x = a[BP_PX + i], p0y = a[BP_PY + i], p1x = a[BP_PX + i1], p1y = a[BP_PY + i1];
    float bx[4], by[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { bx[k] = b[BP_PX + k]; by[k] = b[BP_PY + k]; }
    const float pminx = fminf(p0x, p1x), pmaxx = fmaxf(p0x, p1x), pminy = fminf(p0y, p1y), pmaxy = fmaxf(p0y, p1y);
    const float pdx = __fsub_rn(p1x, p0x), pdy = __fsub_rn(p1y, p0y);
    unsigned int bits = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int j1 = (j + 1) & 3;
        const float q0x = bx[j], q0y = by[j], q1x = bx[j1], q1y = by[j1];
        // check_rect_cross (:43-48)
        const bool rc = live & (pminx <= fmaxf(q0x, q1x)) & (fminf(q0x, q1x) <= pmaxx) & (pminy <= fmaxf(q0y, q1y)) & (fminf(q0y, q1y) <= pmaxy);
#ifndef GLENET_HOST_EMUL
        if (WARP_SKIP && !__any_sync(0xffffffffu, rc)) continue;
#endif
        const float qdx = __fsub_rn(q1x, q0x), qdy = __fsub_rn(q1y, q0y);
        const float s1 = mul_sub<FMA>(__fsub_rn(q0x, p0x), pdy, pdx, __fsub_rn(q0y, p0y));
        const float s2 = __fsub_rn(__fmul_rn(pdx, __fsub_rn(q1y, p0y)), __fmul_rn(pdy, __fsub_rn(q1x, p0x)));
        const float s3 = mul_sub<FMA>(__fsub_rn(p0x, q0x), qdy, __fsub_rn(p0y, q0y), qdx);   // (p0 - q0 is exactly -(q0 - p0): the compiler reuses s1's differences)
        const float s4 = mul_sub<FMA>(qdx, __fsub_rn(p1y, q0y), qdy, __fsub_rn(p1x, q0x));
        if (rc & (__fmul_rn(s1, s2) > 0.f) & (__fmul_rn(s3, s4) > 0.f)) bits |= 1u << j;
    }
    // corner i of b against a (re-read with the dynamic index from shared memory: bx[i] would force the array out of registers)
    if (live & corner_test<FMA, V1>(a, b[BP_PX + i], b[BP_PY + i])) bits |= 16u;
    if (live & corner_test<FMA, V1>(b, p0x, p0y)) bits |= 32u;
    return bits;
}

// The pair's 24 result bits from the four lanes' 6: byte i of the word = bits of lane i.  All 32 lanes must call this.
__device__ __forceinline__ unsigned int clip_quad_gather(unsigned int bits, int i) {
#ifdef GLENET_HOST_EMUL
    (void)i;
    return bits;   // the emulation assembles the word itself
#else
    unsigned int w = bits << (8 * i);
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    return w;
#endif
}
__device__ __forceinline__ unsigned int clip_hits16(unsigned int w) {   // bit 4 i + j: edge pair (i, j), row-major = discovery order
    return (w & 0xfu) | ((w >> 4) & 0xf0u) | ((w >> 8) & 0xf00u) | ((w >> 12) & 0xf000u);
}
__device__ __forceinline__ unsigned int clip_corners8(unsigned int w) { // bit 2 k: b-corner k in a, bit 2 k + 1: a-corner k in b (:177-194)
    return ((w >> 4) & 0x3u) | ((w >> 10) & 0xcu) | ((w >> 16) & 0x30u) | ((w >> 22) & 0xc0u);
}

// All 24 result bits of a pair by ONE lane (phase A of the throughput-oriented kernels); byte i = clip_edge_tests(a, b, i).
// All 32 lanes of the warp must call (lanes without a pair pass live = false and any valid records).
// WARP_SKIP pays where the boxes of a warp's pairs resemble each other (aligned samples vs their GT: the opposite sides
// never meet, 25 % of the tests go away; measured - 3 %); on mixed pairs (tile / NMS kernels) the votes only cost (+ 10 %).
template <bool FMA, bool V1 = false, bool WARP_SKIP = false>
__device__ __forceinline__ unsigned int clip_pair_tests(const float* __restrict__ a, const float* __restrict__ b, bool live = true) {
    unsigned int w = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) w |= clip_edge_tests<FMA, V1, WARP_SKIP>(a, b, i, live) << (8 * i);
    return w;
}

// The admitted corners of a fast-path pair go to the slots behind its k2 crossings, in discovery order (:177-194).
__device__ __forceinline__ void clip_write_corners(const float* __restrict__ a, const float* __restrict__ b, unsigned int w, float2* __restrict__ slots) {
    const unsigned int corners = clip_corners8(w);
    int pos = __popc(clip_hits16(w));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if ((corners >> (2 * k)) & 1u) { slots[pos & (CLIP_SLOTS - 1)] = make_float2(b[BP_PX + k], b[BP_PY + k]); ++pos; }
        if ((corners >> (2 * k + 1)) & 1u) { slots[pos & (CLIP_SLOTS - 1)] = make_float2(a[BP_PX + k], a[BP_PY + k]); ++pos; }
    }
}

#ifndef GLENET_HOST_EMUL
// ---------------------------------------------------------------- phase B, warp-pooled
// hits: the 16 crossing bits of this lane's pair if it is on the fast path (3..8 vertices), else 0.  a_idx / b_idx: record
// numbers of the pair's boxes in arec / brec (< 2048 / < 512).  wl: this WARP's work list (256 entries).  wslots: the
// slots of the warp's 32 pairs (lane l's pair at wslots + 8 l).  All 32 lanes must call.
template <bool FMA>
__device__ __forceinline__ void clip_warp_points(unsigned int hits, unsigned int a_idx, unsigned int b_idx, unsigned int* __restrict__ wl,
                                                 const float* __restrict__ arec, const float* __restrict__ brec, int stride,
                                                 float2* __restrict__ wslots) {
    const int lane = threadIdx.x & 31;
    const int k2 = __popc(hits);
    int incl = k2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    unsigned int* dst = wl + (incl - k2);
    const unsigned int head = (a_idx << 21) | (b_idx << 12) | ((unsigned int)lane << 7);
#pragma unroll 1
    for (unsigned int m = hits, n = 0u; m; m &= m - 1, ++n) *dst++ = head | (n << 4) | (unsigned int)(__ffs((int)m) - 1);
    __syncwarp();
#pragma unroll 1
    for (int t = lane; t < total; t += 32) {
        const unsigned int ent = wl[t];
        const int e = ent & 15, i = e >> 2, j = e & 3, i1 = (i + 1) & 3, j1 = (j + 1) & 3;
        const float* a = arec + (ent >> 21) * stride;
        const float* b = brec + ((ent >> 12) & 511u) * stride;
        wslots[((ent >> 7) & 31u) * CLIP_SLOTS + ((ent >> 4) & 7u)] =
            edge_point<FMA>(a[BP_PX + i], a[BP_PY + i], a[BP_PX + i1], a[BP_PY + i1], b[BP_PX + j], b[BP_PY + j], b[BP_PX + j1], b[BP_PY + j1]);
    }
    __syncwarp();
}

// ---------------------------------------------------------------- more than eight vertices, by the whole warp
// The pairs of this warp that left the fast path (is_slow; ~1 % of a dense workload) are finished one after the other by
// all 32 lanes: lane l computes vertex l (crossing or admitted corner, discovery order, at most 16 as cross_points[16]),
// the centroid is the reference's sequential sum, every lane ranks its vertex by atan2f (stable, like the reference's
// bubble sort), the sorted vertices go through a 16-slot scratch, lane k computes fan term k and the terms are added in
// the reference's order.  No local memory, no second pass over the batch, no CTA barrier.  All 32 lanes must call; the
// overlap of a slow lane's pair is returned to that lane (0 elsewhere).  scratch: 16 float2 of this warp.
template <bool FMA>
__device__ __forceinline__ float clip_warp_slow(bool is_slow, unsigned int w, unsigned int a_idx, unsigned int b_idx,
                                                const float* __restrict__ arec, const float* __restrict__ brec, int stride,
                                                float2* __restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    unsigned int todo = __ballot_sync(0xffffffffu, is_slow);
    float result = 0.f;
#pragma unroll 1
    while (todo) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1;
        const unsigned int ws = __shfl_sync(0xffffffffu, w, src);
        const float* a = arec + __shfl_sync(0xffffffffu, a_idx, src) * stride;
        const float* b = brec + __shfl_sync(0xffffffffu, b_idx, src) * stride;
        const unsigned int hits = clip_hits16(ws), corners = clip_corners8(ws);
        const int k2 = min(__popc(hits), CLIP_SLOW_SLOTS);
        const int cnt = min(k2 + __popc(corners), CLIP_SLOW_SLOTS);
        float2 v = make_float2(0.f, 0.f);
        if (lane < cnt) {
            unsigned int m = lane < k2 ? hits : corners;
#pragma unroll 1
            for (int t = lane < k2 ? lane : lane - k2; t > 0; --t) m &= m - 1;
            const int e = __ffs((int)m) - 1;
            if (lane < k2) {
                const int i = e >> 2, j = e & 3, i1 = (i + 1) & 3, j1 = (j + 1) & 3;
                v = edge_point<FMA>(a[BP_PX + i], a[BP_PY + i], a[BP_PX + i1], a[BP_PY + i1], b[BP_PX + j], b[BP_PY + j], b[BP_PX + j1], b[BP_PY + j1]);
            } else {
                const float* rec = (e & 1) ? a : b;
                v = make_float2(rec[BP_PX + (e >> 1)], rec[BP_PY + (e >> 1)]);
            }
        }
        float sx = 0.f, sy = 0.f;   // sequential, discovery order (:196-197)
#pragma unroll 1
        for (int k = 0; k < cnt; ++k) { sx += __shfl_sync(0xffffffffu, v.x, k); sy += __shfl_sync(0xffffffffu, v.y, k); }
        const float ccx = __fdiv_rn(sx, (float)cnt), ccy = __fdiv_rn(sy, (float)cnt);
        const float key = lane < cnt ? atan2f(__fsub_rn(v.y, ccy), __fsub_rn(v.x, ccx)) : 3.0e38f;
        int rank = 0;
#pragma unroll 1
        for (int k = 0; k < cnt; ++k) {
            const float kk = __shfl_sync(0xffffffffu, key, k);
            rank += ((kk < key) | ((kk == key) & (k < lane))) ? 1 : 0;
        }
        if (lane < cnt) scratch[rank] = v;
        __syncwarp();
        float term = 0.f;
        if (lane >= 1 && lane < cnt) {
            const float2 v0 = scratch[0], p = scratch[lane - 1], c = scratch[lane];
            const float ux = __fsub_rn(p.x, v0.x), uy = __fsub_rn(p.y, v0.y), wx = __fsub_rn(c.x, v0.x), wy = __fsub_rn(c.y, v0.y);
            term = mul_sub<FMA>(ux, wy, uy, wx);
        }
        float area = 0.f;
#pragma unroll 1
        for (int k = 1; k < cnt; ++k) area = __fadd_rn(area, __shfl_sync(0xffffffffu, term, k));
        if (lane == src) result = __fmul_rn(fabsf(area), 0.5f);
        __syncwarp();   // scratch is reused by the next slow pair
    }
    return result;
}
#endif

// ---------------------------------------------------------------- phase B, four lanes per pair
// Quad lane q of a fast-path pair: crossings number q and q + 4 (discovery order) -> slots q, q + 4; its own admitted
// corners -> slots k2 + (admitted corners before them).
template <bool FMA>
__device__ __forceinline__ void clip_write_vertices(const float* __restrict__ a, const float* __restrict__ b, int q, unsigned int w,
                                                    float2* __restrict__ slots) {
    const unsigned int hits = clip_hits16(w), corners = clip_corners8(w);
    const int k2 = __popc(hits);
    unsigned int m = hits;
    if (q >= 1) m &= m - 1;
    if (q >= 2) m &= m - 1;
    if (q >= 3) m &= m - 1;
#pragma unroll 1
    for (int slot = q; m && slot < CLIP_SLOTS; slot += 4) {
        const int e = __ffs((int)m) - 1;
        const int i = e >> 2, j = e & 3, i1 = (i + 1) & 3, j1 = (j + 1) & 3;
        slots[slot] = edge_point<FMA>(a[BP_PX + i], a[BP_PY + i], a[BP_PX + i1], a[BP_PY + i1], b[BP_PX + j], b[BP_PY + j], b[BP_PX + j1], b[BP_PY + j1]);
        m &= m - 1; m &= m - 1; m &= m - 1; m &= m - 1;   // four crossings further
    }
    const unsigned int mine = (corners >> (2 * q)) & 3u;
    if (mine) {
        int pos = k2 + __popc(corners & ((1u << (2 * q)) - 1u));
        if (mine & 1u) { if (pos < CLIP_SLOTS) slots[pos] = make_float2(b[BP_PX + q], b[BP_PY + q]); ++pos; }
        if (mine & 2u) { if (pos < CLIP_SLOTS) slots[pos] = make_float2(a[BP_PX + q], a[BP_PY + q]); }
    }
}

// ---------------------------------------------------------------- phase C
// Sort key of a vertex direction: pseudo-angle shifted from [-2, 2] to [2, 6].  Positive floats order like their bit
// patterns, and in [2, 8) one ulp is 2.4e-7 .. 4.8e-7 ABSOLUTE, so "keys within CLIP_TIE_ULPS" below means "directions
// within ~4e-6 .. 8e-6 of pseudo-angle", uniformly.  The rounding of pseudo_angle + 4 is below 1e-6 (approximate
// division 3e-7, the two additions 6e-8 + 2.4e-7 .. 4.8e-7), so two keys further apart than that are in the order of their
// true directions -- which is also atan2f's order.
constexpr unsigned int CLIP_TIE_ULPS = 16u;
__device__ __forceinline__ unsigned int clip_sort_key(float dy, float dx) {
    return __float_as_uint(__fadd_rn(pseudo_angle(dy, dx), 4.0f));
}

// Area of the polygon in `slots` (3 <= cnt <= 8 vertices in discovery order): the reference's centroid / angular order /
// fan (:196-225) with the pinned arithmetic of polygon_area (geom.cuh).
template <bool FMA>
__device__ __forceinline__ float clip_area8(const float2* __restrict__ slots, int cnt) {
    float X[CLIP_SLOTS], Y[CLIP_SLOTS];
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int k = 0; k < CLIP_SLOTS; ++k) {
        const float2 p = k < cnt ? slots[k] : make_float2(0.f, 0.f);
        X[k] = p.x; Y[k] = p.y;
        sx += p.x; sy += p.y;
    }
    // the reference's centroid: sequential sum in discovery order, IEEE division (:196-197) -- it decides the order of
    // vertices that lie on one ray from it
    const float ccx = __fdiv_rn(sx, (float)cnt), ccy = __fdiv_rn(sy, (float)cnt);
    unsigned int P[CLIP_SLOTS];
#pragma unroll
    for (int k = 0; k < CLIP_SLOTS; ++k) {
        const unsigned int key = clip_sort_key(__fsub_rn(Y[k], ccy), __fsub_rn(X[k], ccx));
        P[k] = k < cnt ? ((key & ~7u) | (unsigned int)k) : 0xffffffffu;
    }
#define GLENET_CE(i, j) { const unsigned int lo_ = min(P[i], P[j]), hi_ = max(P[i], P[j]); P[i] = lo_; P[j] = hi_; }
    GLENET_CE(0, 1) GLENET_CE(2, 3) GLENET_CE(4, 5) GLENET_CE(6, 7)
    GLENET_CE(0, 2) GLENET_CE(1, 3) GLENET_CE(4, 6) GLENET_CE(5, 7)
    GLENET_CE(1, 2) GLENET_CE(5, 6) GLENET_CE(0, 4) GLENET_CE(3, 7)
    GLENET_CE(1, 5) GLENET_CE(2, 6)
    GLENET_CE(1, 4) GLENET_CE(3, 6)
    GLENET_CE(2, 4) GLENET_CE(3, 5)
    GLENET_CE(3, 4)
#undef GLENET_CE
    // Neighbours whose keys are within CLIP_TIE_ULPS (see clip_sort_key) are NOT interchangeable: a corner admitted by the margin
    // and a crossing 1 cm from it can lie on one ray from the centroid, and their order decides ~1e-3 of IoU.  The packed
    // keys lost three bits to the slot number and the pseudo-angle is only monotone up to its own rounding, so such polygons
    // (~1e-4 of the pairs) are sorted again exactly as the reference does it: atan2f keys, ties to the discovery order
    // (stable sort, :199-209).
    bool tie = false;
#pragma unroll
    for (int k = 0; k + 1 < CLIP_SLOTS; ++k) tie |= (k + 1 < cnt) & (P[k + 1] - P[k] < 8u * CLIP_TIE_ULPS);
    if (tie) {
        float K[CLIP_SLOTS];
#pragma unroll
        for (int k = 0; k < CLIP_SLOTS; ++k) {
            const float2 p = slots[k < cnt ? k : 0];
            K[k] = k < cnt ? atan2f(__fsub_rn(p.y, ccy), __fsub_rn(p.x, ccx)) : 3.0e38f;
            P[k] = (unsigned int)k;
        }
#define GLENET_CE(i, j) { const bool sw_ = (K[i] > K[j]) | ((K[i] == K[j]) & (P[i] > P[j]));                      \
                          const float tk_ = sw_ ? K[j] : K[i]; const unsigned int tp_ = sw_ ? P[j] : P[i];        \
                          K[j] = sw_ ? K[i] : K[j]; P[j] = sw_ ? P[i] : P[j]; K[i] = tk_; P[i] = tp_; }
        GLENET_CE(0, 1) GLENET_CE(2, 3) GLENET_CE(4, 5) GLENET_CE(6, 7)
        GLENET_CE(0, 2) GLENET_CE(1, 3) GLENET_CE(4, 6) GLENET_CE(5, 7)
        GLENET_CE(1, 2) GLENET_CE(5, 6) GLENET_CE(0, 4) GLENET_CE(3, 7)
        GLENET_CE(1, 5) GLENET_CE(2, 6)
        GLENET_CE(1, 4) GLENET_CE(3, 6)
        GLENET_CE(2, 4) GLENET_CE(3, 5)
        GLENET_CE(3, 4)
#undef GLENET_CE
    }
    // fan from the first sorted vertex (:219-222); term k = cross(v[k-1] - v0, v[k] - v0)
    const float2 v0 = slots[P[0] & 7u];
    float area = 0.f, ux = 0.f, uy = 0.f;
#pragma unroll
    for (int k = 1; k < CLIP_SLOTS; ++k) {
        const float2 v = slots[P[k] & 7u];   // slots beyond cnt decode to slot 7: in bounds, never added
        const float wx = __fsub_rn(v.x, v0.x), wy = __fsub_rn(v.y, v0.y);
        const float term = mul_sub<FMA>(ux, wy, uy, wx);
        if (k < cnt) area = __fadd_rn(area, term);
        ux = wx;
        uy = wy;
    }
    return __fmul_rn(fabsf(area), 0.5f);
}

// ---------------------------------------------------------------- more than eight vertices
// One lane, one pair, everything in order: all crossings, all admitted corners (<= 16 kept, as cross_points[16]), a
// stable insertion sort on atan2f about the reference's centroid, the fan.  `sv` / `sk` = this lane's scratch slot in shared memory.
template <bool FMA, bool V1 = false>
__device__ __noinline__ float clip_slow_pair(const float* __restrict__ a, const float* __restrict__ b, unsigned int w,
                                             float2* __restrict__ sv, float* __restrict__ sk) {
    const unsigned int corners = clip_corners8(w);
    int cnt = 0;
    float sx = 0.f, sy = 0.f;
#pragma unroll 1
    for (unsigned int m = clip_hits16(w); m; m &= m - 1) {
        const int e = __ffs((int)m) - 1;
        const int i = e >> 2, j = e & 3, i1 = (i + 1) & 3, j1 = (j + 1) & 3;
        const float2 p = edge_point<FMA>(a[BP_PX + i], a[BP_PY + i], a[BP_PX + i1], a[BP_PY + i1], b[BP_PX + j], b[BP_PY + j], b[BP_PX + j1], b[BP_PY + j1]);
        if (cnt < CLIP_SLOW_SLOTS) { sv[cnt++] = p; sx += p.x; sy += p.y; }
    }
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        if (!((corners >> k) & 1u) || cnt >= CLIP_SLOW_SLOTS) continue;
        const float* src = (k & 1) ? a : b;
        const float2 p = make_float2(src[BP_PX + (k >> 1)], src[BP_PY + (k >> 1)]);
        sv[cnt++] = p; sx += p.x; sy += p.y;
    }
    if (cnt < 3) return 0.f;
    const float ccx = __fdiv_rn(sx, (float)cnt), ccy = __fdiv_rn(sy, (float)cnt);   // as the reference (:196-197)
#pragma unroll 1
    for (int k = 0; k < cnt; ++k) {
        const float2 p = sv[k];
        const float kk = atan2f(__fsub_rn(p.y, ccy), __fsub_rn(p.x, ccx));          // point_cmp (:97-99); this path is rare enough for the real thing
        int m = k;
        while (m > 0 && sk[m - 1] > kk) { sk[m] = sk[m - 1]; sv[m] = sv[m - 1]; --m; }
        sk[m] = kk;
        sv[m] = p;
    }
    const float2 v0 = sv[0];
    float area = 0.f, ux = 0.f, uy = 0.f;
#pragma unroll 1
    for (int k = 1; k < cnt; ++k) {
        const float2 v = sv[k];
        const float wx = __fsub_rn(v.x, v0.x), wy = __fsub_rn(v.y, v0.y);
        area = __fadd_rn(area, mul_sub<FMA>(ux, wy, uy, wx));
        ux = wx;
        uy = wy;
    }
    return __fmul_rn(fabsf(area), 0.5f);
}

}  // namespace glenet
