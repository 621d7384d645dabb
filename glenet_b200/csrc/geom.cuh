// Rotated-rectangle geometry shared by the IoU, NMS and points-in-boxes kernels.
//
// The arithmetic of every rounding-sensitive expression is pinned with explicit
// __f{add,sub,mul,maf}_rn intrinsics so that nvcc can neither contract nor
// re-associate it.  The sequences reproduce, operation for operation, what the
// reference computes:
//
//   FMA = true   "GPU dialect": the reference's CUDA kernels as nvcc 12.9 compiles
//                them for sm_100a (fmad on) -- pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu
//                :35-234 (cross / check_rect_cross / check_in_box2d / intersection /
//                rotate_around_center / point_cmp / box_overlap / iou_bev).  Which product
//                of every a*b +- c*d is fused was read from the reference's SASS.
//   FMA = false  "CPU dialect": the reference's host twin, pcdet/ops/iou3d_nms/src/
//                iou3d_cpu.cpp:59-229, compiled by gcc for x86-64 (no FMA, every
//                operation rounded separately).  Trigonometry then comes from the
//                host's libm through the `trig` argument of the kernels.
//
// The structure is NOT the reference's: per-box work (trig, corners, margin
// thresholds, areas) is hoisted into a BoxPre record computed once per box, the
// polygon is the only per-thread array left, and the angular sort
// evaluates one atan2f per vertex instead of two per comparison.
#pragma once
#ifndef GLENET_HOST_EMUL   // tests/emul compiles this header with g++ against host stand-ins for the built-ins
#include <math_constants.h>
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>

namespace glenet {

// ---------------------------------------------------------------- pinned arithmetic
template <bool FMA>
__device__ __forceinline__ float mul_sub(float a, float b, float c, float d) {
    // a*b - c*d ; GPU dialect: fma(a, b, -(c*d)) (second product rounded first)
    if (FMA) return __fmaf_rn(a, b, -__fmul_rn(c, d));
    return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}
template <bool FMA>
__device__ __forceinline__ float mul_add(float a, float b, float c, float d) {
    // a*b + c*d ; GPU dialect: fma(a, b, (c*d))
    if (FMA) return __fmaf_rn(a, b, __fmul_rn(c, d));
    return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}

// ---------------------------------------------------------------- per-box record
// Layout of one BoxPre record (floats).  Stride 21 keeps consecutive records on
// different shared-memory banks.
enum {
    BP_CX = 0, BP_CY = 1,      // centre
    BP_CN = 2, BP_SN = 3,      // cosf(-heading), sinf(-heading)   (check_in_box2d)
    BP_THX = 4, BP_THY = 5,    // dx/2 + 0.01f, dy/2 + 0.01f        (MARGIN, iou3d_nms_kernel.cu:53)
    BP_AREA = 6,               // dx*dy
    BP_PX = 7,                 // 4 rotated corner x
    BP_PY = 11,                // 4 rotated corner y
    BP_ZMIN = 15, BP_ZMAX = 16, BP_VOL = 17,   // boxes_iou3d_gpu terms (iou3d_nms_utils.py:100-117)
    BP_STRIDE = 21,            // record stride with the z terms (3D IoU)
    BP_STRIDE_BEV = 15         // BEV-only kernels drop them (odd stride: consecutive records start on different banks)
};

// trig4 = {cos(h), sin(h), cos(-h), sin(-h)}
template <bool FMA, bool WITH_Z = true>
__device__ __forceinline__ void box_prepare(const float* __restrict__ box, const float4 trig4,
                                            float* __restrict__ o) {
    const float cx = box[0], cy = box[1], z = box[2], dx = box[3], dy = box[4], dz = box[5];
    // iou3d_nms_kernel.cu:109-113 : half extents are exact, so one rounding each
    const float x1 = __fmaf_rn(dx, -0.5f, cx), x2 = __fmaf_rn(dx, 0.5f, cx);
    const float y1 = __fmaf_rn(dy, -0.5f, cy), y2 = __fmaf_rn(dy, 0.5f, cy);
    // rotate_around_center (:91-95) works on (p - centre)
    const float ex1 = __fsub_rn(x1, cx), ex2 = __fsub_rn(x2, cx);
    const float ey1 = __fsub_rn(y1, cy), ey2 = __fsub_rn(y2, cy);
    const float c = trig4.x, s = trig4.y;
    const float exs[4] = {ex1, ex2, ex2, ex1};
    const float eys[4] = {ey1, ey1, ey2, ey2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // new_x = ex*cos + ey*(-sin) + cx ; new_y = ex*sin + ey*cos + cy
        o[BP_PX + k] = __fadd_rn(mul_sub<FMA>(c, exs[k], s, eys[k]), cx);
        o[BP_PY + k] = __fadd_rn(mul_add<FMA>(s, exs[k], c, eys[k]), cy);
    }
    o[BP_CX] = cx;
    o[BP_CY] = cy;
    o[BP_CN] = trig4.z;
    o[BP_SN] = trig4.w;
    o[BP_THX] = __fmaf_rn(dx, 0.5f, 0.01f);   // == dx/2 + MARGIN, single rounding in both dialects
    o[BP_THY] = __fmaf_rn(dy, 0.5f, 0.01f);
    o[BP_AREA] = __fmul_rn(dx, dy);
    if (WITH_Z) {
        const float hz = __fmul_rn(dz, 0.5f);
        o[BP_ZMIN] = __fsub_rn(z, hz);
        o[BP_ZMAX] = __fadd_rn(z, hz);
        o[BP_VOL] = __fmul_rn(__fmul_rn(dx, dy), dz);
    }
}

__device__ __forceinline__ float4 device_trig(float heading) {
    // exactly the four libdevice calls the reference issues (cos(a), sin(a), cos(-a), sin(-a)).  One sincosf would give the
    // same bits for every float (tools/cuda/trig_symmetry.cu checks all 2^32 inputs) and is what the PIB build uses, but
    // in the IoU tile kernel it only shifts register allocation (more spills, 1-7 % slower), so the four calls stay.
    float4 t;
    t.x = cosf(heading);
    t.y = sinf(heading);
    t.z = cosf(-heading);
    t.w = sinf(-heading);
    return t;
}

// The same four values from ONE sincosf (bit-identical for every float, see above).  Used by the thread-per-pair kernels,
// which evaluate the trig of two boxes per pair and gain ~10 % from it.
__device__ __forceinline__ float4 device_trig_fused(float heading) {
    float sn, cs;
    sincosf(heading, &sn, &cs);
    return make_float4(cs, sn, cs, -sn);
}

// Conservative cull radius: circumscribed circle of the box grown by the 0.01 margin
// (a corner may be admitted up to MARGIN outside, in both axes) plus slack for the
// rounding of absolute corner coordinates.  Two boxes whose centres are farther apart
// than r_a + r_b produce no polygon vertex in the reference => overlap is exactly +0.
__device__ __forceinline__ float cull_radius(float cx, float cy, float dx, float dy) {
    const float r = 0.5f * sqrtf(dx * dx + dy * dy);
    return r * 1.0001f + 0.03f + 2e-6f * (fabsf(cx) + fabsf(cy));
}
__device__ __forceinline__ float cull_radius(const float* __restrict__ box) { return cull_radius(box[0], box[1], box[3], box[4]); }

// ---------------------------------------------------------------- polygon
// cross_points[16] of the reference (:155)
constexpr int MAX_POLY = 16;

// intersection() of iou3d_nms_kernel.cu:57-89 for edge p0->p1 of box a and q0->q1 of box b.
// (A branch-free variant that evaluates all 16 edge pairs unconditionally was measured and lost: the
// straight-line code is 135 KB of SASS and stalls on instruction fetch; see DESIGN.md.)
template <bool FMA>
__device__ __forceinline__ bool edge_intersection(float p0x, float p0y, float p1x, float p1y,
                                                  float q0x, float q0y, float q1x, float q1y,
                                                  float& ox, float& oy) {
    // check_rect_cross (:43-48)
    const bool rc = fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) &&
                    fminf(p0y, p1y) <= fmaxf(q0y, q1y) && fminf(q0y, q1y) <= fmaxf(p0y, p1y);
    if (!rc) return false;
    const float pdx = __fsub_rn(p1x, p0x), pdy = __fsub_rn(p1y, p0y);   // p1 - p0
    const float qdx = __fsub_rn(q1x, q0x), qdy = __fsub_rn(q1y, q0y);   // q1 - q0
    // s1 = cross(q0, p1, p0)
    const float s1 = mul_sub<FMA>(__fsub_rn(q0x, p0x), pdy, pdx, __fsub_rn(q0y, p0y));
    // s2 = cross(p1, q1, p0) and s5 = cross(q1, p1, p0) share both products (each rounded)
    const float m1 = __fmul_rn(pdx, __fsub_rn(q1y, p0y));
    const float m2 = __fmul_rn(pdy, __fsub_rn(q1x, p0x));
    const float s2 = __fsub_rn(m1, m2);
    if (!(__fmul_rn(s1, s2) > 0.f)) return false;
    // s3 = cross(p0, q1, q0)
    const float s3 = mul_sub<FMA>(__fsub_rn(p0x, q0x), qdy, __fsub_rn(p0y, q0y), qdx);
    // s4 = cross(q1, p1, q0)
    const float s4 = mul_sub<FMA>(qdx, __fsub_rn(p1y, q0y), qdy, __fsub_rn(p1x, q0x));
    if (!(__fmul_rn(s3, s4) > 0.f)) return false;
    const float s5 = __fsub_rn(m2, m1);
    const float den = __fsub_rn(s5, s1);
    if (fabsf(den) > 1e-8f) {
        ox = __fdiv_rn(mul_sub<FMA>(q0x, s5, q1x, s1), den);
        oy = __fdiv_rn(mul_sub<FMA>(q0y, s5, q1y, s1), den);
    } else {
        const float a0 = __fsub_rn(p0y, p1y), b0 = pdx, c0 = mul_sub<FMA>(p0x, p1y, p1x, p0y);
        const float a1 = __fsub_rn(q0y, q1y), b1 = qdx, c1 = mul_sub<FMA>(q0x, q1y, q0y, q1x);
        const float D = mul_sub<FMA>(b1, a0, b0, a1);
        ox = __fdiv_rn(mul_sub<FMA>(b0, c1, b1, c0), D);
        oy = __fdiv_rn(mul_sub<FMA>(c0, a1, a0, c1), D);
    }
    return true;
}

// The two halves of intersection() separately, for box_overlap's deferred intersection points: edge_crosses = the
// three early-outs (check_rect_cross, s1*s2 > 0, s3*s4 > 0), edge_point = the point of a pair that passed.  Same
// expressions as edge_intersection (which the unrolled clip keeps using).
template <bool FMA>
__device__ __forceinline__ bool edge_crosses(float p0x, float p0y, float p1x, float p1y,
                                             float q0x, float q0y, float q1x, float q1y) {
    const bool rc = fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) &&
                    fminf(p0y, p1y) <= fmaxf(q0y, q1y) && fminf(q0y, q1y) <= fmaxf(p0y, p1y);
    if (!rc) return false;
    const float pdx = __fsub_rn(p1x, p0x), pdy = __fsub_rn(p1y, p0y);
    const float qdx = __fsub_rn(q1x, q0x), qdy = __fsub_rn(q1y, q0y);
    const float s1 = mul_sub<FMA>(__fsub_rn(q0x, p0x), pdy, pdx, __fsub_rn(q0y, p0y));
    const float s2 = __fsub_rn(__fmul_rn(pdx, __fsub_rn(q1y, p0y)), __fmul_rn(pdy, __fsub_rn(q1x, p0x)));
    if (!(__fmul_rn(s1, s2) > 0.f)) return false;
    const float s3 = mul_sub<FMA>(__fsub_rn(p0x, q0x), qdy, __fsub_rn(p0y, q0y), qdx);
    const float s4 = mul_sub<FMA>(qdx, __fsub_rn(p1y, q0y), qdy, __fsub_rn(p1x, q0x));
    return __fmul_rn(s3, s4) > 0.f;
}
template <bool FMA>
__device__ __forceinline__ float2 edge_point(float p0x, float p0y, float p1x, float p1y,
                                             float q0x, float q0y, float q1x, float q1y) {
    const float pdx = __fsub_rn(p1x, p0x), pdy = __fsub_rn(p1y, p0y);
    const float s1 = mul_sub<FMA>(__fsub_rn(q0x, p0x), pdy, pdx, __fsub_rn(q0y, p0y));
    const float m1 = __fmul_rn(pdx, __fsub_rn(q1y, p0y));
    const float m2 = __fmul_rn(pdy, __fsub_rn(q1x, p0x));
    const float s5 = __fsub_rn(m2, m1);
    const float den = __fsub_rn(s5, s1);
    float2 o;
    if (fabsf(den) > 1e-8f) {
        o.x = __fdiv_rn(mul_sub<FMA>(q0x, s5, q1x, s1), den);
        o.y = __fdiv_rn(mul_sub<FMA>(q0y, s5, q1y, s1), den);
    } else {
        const float a0 = __fsub_rn(p0y, p1y), b0 = pdx, c0 = mul_sub<FMA>(p0x, p1y, p1x, p0y);
        const float a1 = __fsub_rn(q0y, q1y), b1 = __fsub_rn(q1x, q0x), c1 = mul_sub<FMA>(q0x, q1y, q0y, q1x);
        const float D = mul_sub<FMA>(b1, a0, b0, a1);
        o.x = __fdiv_rn(mul_sub<FMA>(b0, c1, b1, c0), D);
        o.y = __fdiv_rn(mul_sub<FMA>(c0, a1, a0, c1), D);
    }
    return o;
}

// check_in_box2d (:50-60) of point (px,py) against a prepared box.
template <bool FMA>
__device__ __forceinline__ bool corner_in_box(const float* __restrict__ b, float px, float py) {
    const float dxp = __fsub_rn(px, b[BP_CX]), dyp = __fsub_rn(py, b[BP_CY]);
    const float cn = b[BP_CN], sn = b[BP_SN];
    const float rx = mul_sub<FMA>(cn, dxp, sn, dyp);   // dxp*cos + dyp*(-sin)
    const float ry = mul_add<FMA>(cn, dyp, sn, dxp);   // dxp*sin + dyp*cos (dyp*cos is the fused product)
    return fabsf(rx) < b[BP_THX] && fabsf(ry) < b[BP_THY];
}

// Separating-axis test with a safety margin.  True only when the two rectangles are further apart along
// one of their four edge normals than every tolerance of the reference's clipping code (MARGIN 0.01 of
// check_in_box2d :53, EPS 1e-8 of intersection() :57-89) plus the rounding of this test itself; the
// reference then finds no corner inside the other box and no edge crossing, i.e. cnt == 0 and the overlap
// is exactly +0.  Any NaN/Inf makes every comparison false => "not separated" => the clip code decides.
// Budget of the 0.05 slack: 0.01 MARGIN, 2 x 0.01 for |dx/2 + 0.01| < |dx|/2 when an extent is negative,
// the rest plus the relative terms for float rounding (coordinates up to 1e6 m).
__device__ __forceinline__ bool sat_separated(const float* __restrict__ a, const float* __restrict__ b) {
    const float acx = a[BP_CX], acy = a[BP_CY], bcx = b[BP_CX], bcy = b[BP_CY];
    const float dx = bcx - acx, dy = bcy - acy;
    const float ca = a[BP_CN], sa = a[BP_SN], cb = b[BP_CN], sb = b[BP_SN];
    const float C = fabsf(ca * cb + sa * sb), S = fabsf(sa * cb - ca * sb);
    const float ax = fabsf(a[BP_THX]), ay = fabsf(a[BP_THY]), bx = fabsf(b[BP_THX]), by = fabsf(b[BP_THY]);
    const float slack = 0.05f + 4e-6f * (fabsf(acx) + fabsf(acy) + fabsf(bcx) + fabsf(bcy));
    const float k = 1.0002f;
    const float pax = fabsf(ca * dx - sa * dy), pay = fabsf(sa * dx + ca * dy);
    const float pbx = fabsf(cb * dx - sb * dy), pby = fabsf(sb * dx + cb * dy);
    return pax > (ax + bx * C + by * S) * k + slack || pay > (ay + bx * S + by * C) * k + slack ||
           pbx > (bx + ax * C + ay * S) * k + slack || pby > (by + ax * S + ay * C) * k + slack;
}

// Upper bound of the overlap area the reference's clipping code can return for two prepared boxes.
// Every vertex of its polygon is an edge crossing or a corner that check_in_box2d admits, so all of them lie
// inside BOTH rectangles grown by MARGIN; that intersection is convex, hence the fan area is at most the area
// of the intersection, which in turn is at most (its extent along a's length axis) x (its extent along a's
// width axis) -- and the same in b's frame.  The extents are those of a grown by MARGIN (+ slack, as in
// sat_separated) clipped against the projection of the grown b.  Used by NMS to skip the clip of pairs whose
// IoU cannot exceed the threshold; NaN/Inf inputs give NaN/Inf, which callers must treat as "cannot skip".
__device__ __forceinline__ float overlap_upper_bound(const float* __restrict__ a, const float* __restrict__ b) {
    const float acx = a[BP_CX], acy = a[BP_CY], bcx = b[BP_CX], bcy = b[BP_CY];
    const float dx = bcx - acx, dy = bcy - acy;
    const float ca = a[BP_CN], sa = a[BP_SN], cb = b[BP_CN], sb = b[BP_SN];
    const float C = fabsf(ca * cb + sa * sb), S = fabsf(sa * cb - ca * sb);
    const float slack = 0.03f + 4e-6f * (fabsf(acx) + fabsf(acy) + fabsf(bcx) + fabsf(bcy));
    const float ax = fabsf(a[BP_THX]) + slack, ay = fabsf(a[BP_THY]) + slack, bx = fabsf(b[BP_THX]) + slack, by = fabsf(b[BP_THY]) + slack;
    const float pax = ca * dx - sa * dy, pay = sa * dx + ca * dy;     // centre of b in a's frame
    const float pbx = -(cb * dx - sb * dy), pby = -(sb * dx + cb * dy);   // centre of a in b's frame
    const float rbx = bx * C + by * S, rby = bx * S + by * C;         // half extents of b along a's axes
    const float rax = ax * C + ay * S, ray = ax * S + ay * C;
    const float oax = fminf(ax, pax + rbx) - fmaxf(-ax, pax - rbx), oay = fminf(ay, pay + rby) - fmaxf(-ay, pay - rby);
    const float obx = fminf(bx, pbx + rax) - fmaxf(-bx, pbx - rax), oby = fminf(by, pby + ray) - fmaxf(-by, pby - ray);
    const float ua = (oax > 0.f && oay > 0.f) ? oax * oay : ((oax != oax || oay != oay) ? CUDART_NAN_F : 0.f);
    const float ub = (obx > 0.f && oby > 0.f) ? obx * oby : ((obx != obx || oby != oby) ? CUDART_NAN_F : 0.f);
    return (ua != ua || ub != ub) ? CUDART_NAN_F : fminf(ua, ub) * 1.001f;
}

// Geometric (true) intersection area of the two rectangles, APPROXIMATELY (float32, approximate reciprocals): Green's
// theorem over the boundary of A n B = (edges of b clipped to a) + (edges of a clipped to b).  Each clip is a Liang-Barsky
// parameter interval against an axis-aligned box -- b's edges in a's frame, a's edges in b's frame -- and a clipped edge
// contributes (t1 - t0) * cross(P0, D) to the shoelace sum (for a's own edges that cross product is 2 hx hy).  Branch-free,
// no vertex list, ~240 instructions -- against ~1800 for the bit-faithful clip.  It is NOT the reference's value: the
// reference's polygon also takes corners up to MARGIN outside the other box, so
//     approx - slack  <=  reference overlap  <=  approx + MARGIN-band + slack        (overlap_approx_band)
// which is what NMS needs to decide  IoU > thresh  for every pair that is not within that band of the threshold
// (measured against the reference restated in C, tests/test_clip_emul.py: reference - true in [-4e-5, +0.22 * band] on 2.5e5 proposal pairs).
// Requires positive extents; NaN / Inf propagate (callers treat a non-finite result as "run the exact clip").
__device__ __forceinline__ float rcp_approx(float x) {   // one MUFU.RCP (__fdividef(1, x) adds a range fix-up that the slab tests do not need)
#ifdef GLENET_HOST_EMUL
    return 1.0f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float lb_weight(float px, float py, float dx, float dy, float hx, float hy) {
    const float ix = rcp_approx(dx), iy = rcp_approx(dy);                      // d == 0: +-inf, the slab tests degenerate correctly (0 * inf = NaN is dropped by fminf / fmaxf)
    const float ta = (-hx - px) * ix, tb = (hx - px) * ix, tc = (-hy - py) * iy, td = (hy - py) * iy;
    const float t0 = fmaxf(fmaxf(fminf(ta, tb), fminf(tc, td)), 0.f);
    const float t1 = fminf(fminf(fmaxf(ta, tb), fmaxf(tc, td)), 1.f);
    return fmaxf(t1 - t0, 0.f);
}
__device__ __forceinline__ float overlap_approx(const float* __restrict__ a, const float* __restrict__ b) {
    const float ahx = a[BP_THX] - 0.01f, ahy = a[BP_THY] - 0.01f, bhx = b[BP_THX] - 0.01f, bhy = b[BP_THY] - 0.01f;
    const float acx = a[BP_CX], acy = a[BP_CY], bcx = b[BP_CX], bcy = b[BP_CY];
    const float ca = a[BP_CN], sa = a[BP_SN], cb = b[BP_CN], sb = b[BP_SN];
    float bx[4], by[4], ax[4], ay[4];     // b's corners in a's frame, a's corners in b's frame
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float dx = b[BP_PX + k] - acx, dy = b[BP_PY + k] - acy;
        bx[k] = ca * dx - sa * dy; by[k] = sa * dx + ca * dy;
        const float ex = a[BP_PX + k] - bcx, ey = a[BP_PY + k] - bcy;
        ax[k] = cb * ex - sb * ey; ay[k] = sb * ex + cb * ey;
    }
    float sum = 0.f, wa = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int k1 = (k + 1) & 3;
        const float dx = bx[k1] - bx[k], dy = by[k1] - by[k];
        sum += lb_weight(bx[k], by[k], dx, dy, ahx, ahy) * (bx[k] * dy - by[k] * dx);
        wa += lb_weight(ax[k], ay[k], ax[k1] - ax[k], ay[k1] - ay[k], bhx, bhy);
    }
    return 0.5f * fabsf(sum + wa * (2.f * ahx * ahy));
}
// The shoelace sum counts a boundary segment shared by both rectangles once only if the two clips (done in different
// frames) agree on which side it lies; for edges that are parallel within rounding noise they need not, and the error is the
// whole edge integral.  So the filter is used only for positive-size boxes whose relative heading is at least 1e-3 rad away
// from every multiple of 90 degrees (random proposals: practically always; grid-aligned boxes: never -- they take the clip).
__device__ __forceinline__ bool overlap_approx_usable(const float* __restrict__ a, const float* __restrict__ b) {
    const float ca = a[BP_CN], sa = a[BP_SN], cb = b[BP_CN], sb = b[BP_SN];
    const float C = fabsf(ca * cb + sa * sb), S = fabsf(sa * cb - ca * sb);
    return a[BP_THX] > 0.011f && a[BP_THY] > 0.011f && b[BP_THX] > 0.011f && b[BP_THY] > 0.011f && fminf(C, S) > 1e-3f;
}
// slack: float32 / approximate-reciprocal error of overlap_approx (measured < 2e-4 (Sa + Sb); 10 x margin);
// band: what the reference's MARGIN-admitted corners can add -- its polygon lies inside both rectangles grown by MARGIN * sqrt 2
__device__ __forceinline__ void overlap_approx_band(const float* __restrict__ a, const float* __restrict__ b, float& slack, float& band) {
    const float pa = a[BP_THX] + a[BP_THY], pb = b[BP_THX] + b[BP_THY];        // half perimeters / 2 (+ 0.02)
    // ... plus the reference's own loss far from the origin: its crossing points are computed from ABSOLUTE coordinates
    // (s5 * q0 - s1 * q1, iou3d_nms_kernel.cu:77-89), error ~1e-7 |coordinate| per point (measured 1.3e-7; 8 x margin)
    slack = 2e-3f * (a[BP_AREA] + b[BP_AREA]) + 1e-3f
          + 1e-6f * (fabsf(a[BP_CX]) + fabsf(a[BP_CY]) + fabsf(b[BP_CX]) + fabsf(b[BP_CY])) * 4.f * fmaxf(pa, pb);
    band = 0.015f * 4.f * fminf(pa, pb);
}

// Monotone stand-in for atan2f(dy, dx) on (-pi, pi]: same ordering of the polygon vertices about the
// centroid as point_cmp (:97-99) except between directions that differ by a few ulps, where the fan
// area is insensitive to the order.  ~8 instructions instead of ~50.
__device__ __forceinline__ float pseudo_angle(float dy, float dx) {
    const float den = fabsf(dx) + fabsf(dy);
    const float q = (den > 0.f) ? __fdividef(dx, den) : 1.f;   // in [-1, 1]
    return copysignf(1.f - q, dy);                             // [0, 2] for dy >= +0, [-2, -0] for dy <= -0
}

// Area of the polygon spanned by `cnt` vertices (any order): angular ordering about the centroid, then
// the reference's fan from the first sorted vertex (:196-225).  Vertices are read through `V(k)`.
//   cnt <= 8 (>98 % of pairs): 8 register slots, branch-free 19-comparator sorting network on a monotone
//            pseudo-angle, predicated fan -- no divergence between lanes with different vertex counts.
//   cnt  > 8 (degenerate / nearly coincident boxes): stable insertion sort in per-thread local memory.
// The centroid only feeds the ordering (not the area), so its summation order and an approximate
// reciprocal do not matter; the fan terms and their summation order are the reference's.
template <bool FMA, typename VertexFn>
__device__ __forceinline__ float polygon_area(int cnt, VertexFn V) {
    if (cnt < 3) return 0.f;   // the fan of 0, 1 or 2 vertices has area exactly +0 in the reference too
    if (cnt > MAX_POLY) cnt = MAX_POLY;
    float area = 0.f;
    if (cnt <= 8) {
        float X[8], Y[8], K[8];
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool v = k < cnt;
            const float2 p = v ? V(k) : make_float2(0.f, 0.f);
            X[k] = p.x; Y[k] = p.y;
            sx += p.x; sy += p.y;
        }
        const float inv_cnt = __frcp_rn((float)cnt);
        const float ccx = sx * inv_cnt, ccy = sy * inv_cnt;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            K[k] = (k < cnt) ? pseudo_angle(__fsub_rn(Y[k], ccy), __fsub_rn(X[k], ccx)) : 3.0e38f;
#define GLENET_CE(i, j)                                                       \
        {                                                                     \
            const bool sw = K[i] > K[j];                                      \
            const float tk = sw ? K[j] : K[i], tx = sw ? X[j] : X[i], ty = sw ? Y[j] : Y[i]; \
            K[j] = sw ? K[i] : K[j]; X[j] = sw ? X[i] : X[j]; Y[j] = sw ? Y[i] : Y[j];       \
            K[i] = tk; X[i] = tx; Y[i] = ty;                                  \
        }
        GLENET_CE(0, 1) GLENET_CE(2, 3) GLENET_CE(4, 5) GLENET_CE(6, 7)
        GLENET_CE(0, 2) GLENET_CE(1, 3) GLENET_CE(4, 6) GLENET_CE(5, 7)
        GLENET_CE(1, 2) GLENET_CE(5, 6) GLENET_CE(0, 4) GLENET_CE(3, 7)
        GLENET_CE(1, 5) GLENET_CE(2, 6)
        GLENET_CE(1, 4) GLENET_CE(3, 6)
        GLENET_CE(2, 4) GLENET_CE(3, 5)
        GLENET_CE(3, 4)
#undef GLENET_CE
        // fan from the first sorted vertex (:219-222); term k = cross(v[k-1] - v0, v[k] - v0)
        float ux = 0.f, uy = 0.f;
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            const float wx = __fsub_rn(X[k], X[0]), wy = __fsub_rn(Y[k], Y[0]);
            const float term = mul_sub<FMA>(ux, wy, uy, wx);
            if (k < cnt) area = __fadd_rn(area, term);
            ux = wx;
            uy = wy;
        }
    } else {
        float vx[MAX_POLY], vy[MAX_POLY], key[MAX_POLY];
        float sx = 0.f, sy = 0.f;
        for (int k = 0; k < cnt; ++k) { const float2 p = V(k); vx[k] = p.x; vy[k] = p.y; sx += p.x; sy += p.y; }
        const float inv_cnt = __frcp_rn((float)cnt);
        const float ccx = sx * inv_cnt, ccy = sy * inv_cnt;
        for (int k = 0; k < cnt; ++k) {
            const float x = vx[k], y = vy[k];
            const float kk = pseudo_angle(__fsub_rn(y, ccy), __fsub_rn(x, ccx));
            int m = k;
            while (m > 0 && key[m - 1] > kk) {
                key[m] = key[m - 1];
                vx[m] = vx[m - 1];
                vy[m] = vy[m - 1];
                --m;
            }
            key[m] = kk;
            vx[m] = x;
            vy[m] = y;
        }
        const float x0 = vx[0], y0 = vy[0];
        float ux = 0.f, uy = 0.f;
        for (int k = 1; k < cnt; ++k) {
            const float wx = __fsub_rn(vx[k], x0), wy = __fsub_rn(vy[k], y0);
            area = __fadd_rn(area, mul_sub<FMA>(ux, wy, uy, wx));
            ux = wx;
            uy = wy;
        }
    }
    return __fmul_rn(fabsf(area), 0.5f);
}

// box_overlap (:104-225): overlap area of prepared boxes a (row) and b (column), one thread per pair.
// The vertex list (<= 16 entries, dynamically indexed) lives in per-thread local memory.
//   box_overlap         : 4 x 4 edge loops ROLLED, corners re-read from the BoxPre records in shared memory.
//                         For kernels where only some warps clip while others stream (tile / NMS kernels):
//                         the unrolled form is ~6000 SASS instructions and thrashes the instruction cache.
//   box_overlap_unrolled: everything inlined and unrolled, corners in registers.  For the aligned kernel,
//                         where every warp runs the same clip code over register-resident records.
template <bool FMA>
__device__ __noinline__ float box_overlap(const float* __restrict__ a, const float* __restrict__ b) {
    float2 v[MAX_POLY];
    int cnt = 0;
    // Pass 1: which of the 16 edge pairs cross (cheap early-outs).  Pass 2: the intersection points -- two IEEE
    // divisions each -- only for the pairs that do, every lane walking ITS OWN hit list: the warp iterates
    // max(hits per lane) ~ 4 times with most lanes busy instead of 16 times with ~5 of 32.
    unsigned int hits = 0u;
    float bx[4], by[4];   // b's corners in registers: the j loop is unrolled, the i loop is not
#pragma unroll
    for (int k = 0; k < 4; ++k) { bx[k] = b[BP_PX + k]; by[k] = b[BP_PY + k]; }
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const float p0x = a[BP_PX + i], p0y = a[BP_PY + i], p1x = a[BP_PX + ((i + 1) & 3)], p1y = a[BP_PY + ((i + 1) & 3)];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (edge_crosses<FMA>(p0x, p0y, p1x, p1y, bx[j], by[j], bx[(j + 1) & 3], by[(j + 1) & 3]))
                hits |= 1u << (i * 4 + j);
    }
#pragma unroll 1
    while (hits) {   // ascending (i, j): the reference's order of discovery
        const int e = __ffs(hits) - 1;
        hits &= hits - 1;
        const int i = e >> 2, j = e & 3, i1 = (i + 1) & 3, j1 = (j + 1) & 3;
        v[cnt++] = edge_point<FMA>(a[BP_PX + i], a[BP_PY + i], a[BP_PX + i1], a[BP_PY + i1], b[BP_PX + j], b[BP_PY + j], b[BP_PX + j1], b[BP_PY + j1]);
    }
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const float bxk = b[BP_PX + k], byk = b[BP_PY + k], axk = a[BP_PX + k], ayk = a[BP_PY + k];
        if (corner_in_box<FMA>(a, bxk, byk) && cnt < MAX_POLY) v[cnt++] = make_float2(bxk, byk);
        if (corner_in_box<FMA>(b, axk, ayk) && cnt < MAX_POLY) v[cnt++] = make_float2(axk, ayk);
    }
    return polygon_area<FMA>(cnt, [&](int k) { return v[k]; });
}

// ---------------------------------------------------------------- pcdet/ops/iou3d dialect ("V1")
// The older op behind boxes_aligned_iou3d_gpu (pcdet/ops/iou3d/src/iou3d_kernel.cu): boxes arrive as
// [x1, y1, x2, y2, angle], corners are rotated CLOCKWISE about the box centre (rotate_around_center :122-126) and
// check_in_box2d (:50-66) compares the back-rotated point against the box edges -/+ MARGIN = 1e-5.  Contraction
// pattern read from the SASS of boxes_aligned_overlap_kernel (nvcc 12.9, sm_100a):
//   ex = p.x - cx with cx = (x1 + x2) / 2 (the halving is exact, so fma(x1 + x2, -0.5, p.x) == p.x - cx)
//   rotate : x' = fma(cos, ex, sin*ey) + cx          y' = fma(cos, ey, -(sin*ex)) + cy
//   in box : rx = fma(cn, dxp, sn*dyp) + cx          ry = fma(cn, dyp, -(sn*dxp)) + cy      (cn, sn = cos/sin(-angle))
//   intersection(), the fan and the final |area| / 2 are the expressions of the iou3d_nms kernels (edge_intersection).
enum { BP1_X1M = BP_THX, BP1_Y1M = BP_THY, BP1_X2P = BP_ZMIN, BP1_Y2P = BP_ZMAX };   // box edges -/+ MARGIN
template <bool FMA>
__device__ __forceinline__ void box_prepare_v1(float x1, float y1, float x2, float y2, const float4 trig4, float* __restrict__ o) {
    const float cx = __fmul_rn(__fadd_rn(x1, x2), 0.5f), cy = __fmul_rn(__fadd_rn(y1, y2), 0.5f);
    const float ex1 = __fsub_rn(x1, cx), ex2 = __fsub_rn(x2, cx), ey1 = __fsub_rn(y1, cy), ey2 = __fsub_rn(y2, cy);
    const float c = trig4.x, s = trig4.y;
    const float exs[4] = {ex1, ex2, ex2, ex1};
    const float eys[4] = {ey1, ey1, ey2, ey2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (FMA) {
            o[BP_PX + k] = __fadd_rn(__fmaf_rn(c, exs[k], __fmul_rn(s, eys[k])), cx);
            o[BP_PY + k] = __fadd_rn(__fmaf_rn(c, eys[k], -__fmul_rn(s, exs[k])), cy);
        } else {   // (ex*cos + ey*sin) + cx ; (-ex*sin + ey*cos) + cy, every operation rounded
            o[BP_PX + k] = __fadd_rn(__fadd_rn(__fmul_rn(exs[k], c), __fmul_rn(eys[k], s)), cx);
            o[BP_PY + k] = __fadd_rn(__fadd_rn(-__fmul_rn(exs[k], s), __fmul_rn(eys[k], c)), cy);
        }
    }
    o[BP_CX] = cx; o[BP_CY] = cy; o[BP_CN] = trig4.z; o[BP_SN] = trig4.w;
    o[BP1_X1M] = __fadd_rn(x1, -1e-5f); o[BP1_Y1M] = __fadd_rn(y1, -1e-5f);
    o[BP1_X2P] = __fadd_rn(x2, 1e-5f);  o[BP1_Y2P] = __fadd_rn(y2, 1e-5f);
    o[BP_AREA] = 0.f;
}
template <bool FMA>
__device__ __forceinline__ bool corner_in_box_v1(const float* __restrict__ b, float px, float py) {
    const float cx = b[BP_CX], cy = b[BP_CY], cn = b[BP_CN], sn = b[BP_SN];
    const float dxp = __fsub_rn(px, cx), dyp = __fsub_rn(py, cy);
    float rx, ry;
    if (FMA) {
        rx = __fadd_rn(__fmaf_rn(cn, dxp, __fmul_rn(sn, dyp)), cx);
        ry = __fadd_rn(__fmaf_rn(cn, dyp, -__fmul_rn(sn, dxp)), cy);
    } else {
        rx = __fadd_rn(__fadd_rn(__fmul_rn(dxp, cn), __fmul_rn(dyp, sn)), cx);
        ry = __fadd_rn(__fadd_rn(-__fmul_rn(dxp, sn), __fmul_rn(dyp, cn)), cy);
    }
    return rx > b[BP1_X1M] && rx < b[BP1_X2P] && ry > b[BP1_Y1M] && ry < b[BP1_Y2P];
}
template <bool FMA, bool V1>
__device__ __forceinline__ bool corner_test(const float* __restrict__ b, float px, float py) {
    return V1 ? corner_in_box_v1<FMA>(b, px, py) : corner_in_box<FMA>(b, px, py);
}

template <bool FMA, bool V1 = false>
__device__ __forceinline__ float box_overlap_unrolled(const float* __restrict__ a, const float* __restrict__ b) {
    float ax[4], ay[4], bx[4], by[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ax[k] = a[BP_PX + k]; ay[k] = a[BP_PY + k];
        bx[k] = b[BP_PX + k]; by[k] = b[BP_PY + k];
    }
    float2 v[MAX_POLY];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float ox, oy;
            if (edge_intersection<FMA>(ax[i], ay[i], ax[(i + 1) & 3], ay[(i + 1) & 3],
                                       bx[j], by[j], bx[(j + 1) & 3], by[(j + 1) & 3], ox, oy)) {
                if (cnt < MAX_POLY) v[cnt++] = make_float2(ox, oy);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (corner_test<FMA, V1>(a, bx[k], by[k]) && cnt < MAX_POLY) v[cnt++] = make_float2(bx[k], by[k]);
        if (corner_test<FMA, V1>(b, ax[k], ay[k]) && cnt < MAX_POLY) v[cnt++] = make_float2(ax[k], ay[k]);
    }
    return polygon_area<FMA>(cnt, [&](int k) { return v[k]; });
}

// iou_bev (:227-234)
__device__ __forceinline__ float iou_from_overlap(float sa, float sb, float ov) {
    return __fdiv_rn(ov, fmaxf(__fsub_rn(__fadd_rn(sa, sb), ov), 1e-8f));
}

// torch.clamp(x, min=lo) propagates NaN (iou3d_nms_utils.py:112,119)
__device__ __forceinline__ float clamp_min_nan(float v, float lo) { return (v != v) ? v : fmaxf(v, lo); }
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? (a + b) : fmaxf(a, b); }
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a || b != b) ? (a + b) : fminf(a, b); }

// boxes_iou3d_gpu (iou3d_nms_utils.py:100-119), every step separately rounded as torch does
__device__ __forceinline__ float iou3d_from_terms(float a_zmin, float a_zmax, float a_vol, float b_zmin, float b_zmax, float b_vol, float ov,
                                                  float eps = 1e-6f) {
    const float max_of_min = max_nan(a_zmin, b_zmin);
    const float min_of_max = min_nan(a_zmax, b_zmax);
    const float oh = clamp_min_nan(__fsub_rn(min_of_max, max_of_min), 0.f);
    const float o3 = __fmul_rn(ov, oh);
    const float den = clamp_min_nan(__fsub_rn(__fadd_rn(a_vol, b_vol), o3), eps);
    return __fdiv_rn(o3, den);
}
__device__ __forceinline__ float iou3d_from_overlap(const float* __restrict__ a, const float* __restrict__ b, float ov) {
    return iou3d_from_terms(a[BP_ZMIN], a[BP_ZMAX], a[BP_VOL], b[BP_ZMIN], b[BP_ZMAX], b[BP_VOL], ov);
}

}  // namespace glenet
