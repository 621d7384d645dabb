// Rotated BEV IoU of the KITTI evaluator for sm_100a.
//
// Replaces rotate_iou_gpu_eval / rotate_iou_kernel_eval of
// pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py:263-330 (a numba-CUDA module, called from
// kitti_object_eval_python/eval.py:117,151 for the BEV and 3D overlaps of every evaluation part).  Box rows are
// [x, y, x_d, y_d, angle] (angle clockwise when positive, :289-292); iou[n][k] = devRotateIoUEval(query[k], boxes[n])
// -- the QUERY box is the first argument (:281-283).
//
// The reference runs the whole clip for every (box, query) pair: 64 threads per block, each looping over 64 query
// boxes, corners / trigonometry recomputed per pair, vertex lists in local memory.  Here a CTA owns a 64 x 64 tile:
//   1. the tile's 128 boxes are prepared once (corners, area, edge vectors, cull circle) into shared memory;
//   2. every pair takes a circle test; culled pairs (the evaluator pairs every ground-truth box of ~75 frames with every
//      detection of those frames, > 99 % are far apart) get their final value at once, written coalesced along k;
//   3. the surviving pairs are queued in shared memory and clipped one per thread, vertex lists in shared memory
//      (thread-strided, bank-conflict free) -- no local-memory stack.
//
// Arithmetic: numba types float32 x float32 as float32 and float32 x Python-float as float64, and its NVVM/ptxas pipeline
// contracts a*b +- c*d into one FMA plus one rounded product.  Which product is fused was read from the SASS of the
// reference kernel compiled by numba 0.65 / CUDA 12.9 for this architecture (recipe: tests/golden/make_golden_rotate_iou.py,
// DESIGN.md section 5.7) and is pinned here with *_rn intrinsics:
//   corners    x = fma(cos, ex, rn(sin*ey)) + cx          y = fma(cos, ey, -rn(sin*ex)) + cy
//   in-quad    abab = fma(ab0, ab0, rn(ab1*ab1))          abap = fma(ab1, ap1, rn(ab0*ap0))      (ad likewise)
//   segments   the four orientation tests compare two rounded products; every x*y - z*w afterwards is fma(x, y, -rn(z*w))
//   ordering   d2 = fma(v0, v0, rn(v1*v1)), sqrt.rn, div.rn, key = v1 < 0 ? -2 - v0 : v0, stable insertion sort
//   area       sum over the fan of |(double)fma(a0-c0, b1-c1, -rn((a1-c1)*(b0-c0))) * 0.5| in float64
//   result     float64 division (criterion -1: inter / ((double)(area1 + area2) - inter)), rounded to float32 on store.
// The reference appends vertices without a bound into 8 slots (rotate_iou.py:235); in exact arithmetic two convex
// quadrilaterals never have more, and a ninth produced by rounding is dropped here instead of written out of bounds.
#include "common.cuh"
#include "../../include/glenet_geom.h"

namespace glenet {

constexpr int RI_TILE = 64;                 // boxes x queries per CTA
constexpr int RI_THREADS = 256;
constexpr int RI_REC = 16;                  // floats per prepared box
enum { RB_C = 0, RB_AREA = 8, RB_CX = 9, RB_CY = 10, RB_RAD = 11, RB_AB0 = 12, RB_AB1 = 13, RB_AD0 = 14, RB_AD1 = 15 };

// rbbox_to_corners (rotate_iou.py:204-228) + the per-box terms of point_in_quadrilateral (:160-176) and devRotateIoUEval (:250-251)
__device__ __forceinline__ void ri_prepare(const float* __restrict__ rb, float* __restrict__ o) {
    const float x = rb[0], y = rb[1], xd = rb[2], yd = rb[3], ang = rb[4];
    const float cs = cosf(ang), sn = sinf(ang);
    const float exn = __fmul_rn(xd, -0.5f), exp_ = __fmul_rn(xd, 0.5f), eyn = __fmul_rn(yd, -0.5f), eyp = __fmul_rn(yd, 0.5f);
    const float ex[4] = {exn, exn, exp_, exp_}, ey[4] = {eyn, eyp, eyp, eyn};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        o[RB_C + 2 * i] = __fadd_rn(__fmaf_rn(cs, ex[i], __fmul_rn(sn, ey[i])), x);
        o[RB_C + 2 * i + 1] = __fadd_rn(__fmaf_rn(cs, ey[i], -__fmul_rn(sn, ex[i])), y);
    }
    o[RB_AREA] = __fmul_rn(xd, yd);
    o[RB_CX] = x; o[RB_CY] = y;
    // cull circle: circumradius with slack for the rounding of the corners (relative to their magnitude).  Two boxes whose
    // circles are disjoint have no corner inside the other and no crossing edges in the reference either.
    o[RB_AB0] = __fsub_rn(o[RB_C + 2], o[RB_C + 0]); o[RB_AB1] = __fsub_rn(o[RB_C + 3], o[RB_C + 1]);
    o[RB_AD0] = __fsub_rn(o[RB_C + 6], o[RB_C + 0]); o[RB_AD1] = __fsub_rn(o[RB_C + 7], o[RB_C + 1]);
    // A box with a zero-length edge passes point_in_quadrilateral for EVERY point (0 >= 0 && 0 >= 0): the reference then
    // reports the other box's whole area as the intersection, however far away it is.  Such boxes (and non-finite ones) are
    // never culled.
    const float abab = __fmaf_rn(o[RB_AB0], o[RB_AB0], __fmul_rn(o[RB_AB1], o[RB_AB1]));
    const float adad = __fmaf_rn(o[RB_AD0], o[RB_AD0], __fmul_rn(o[RB_AD1], o[RB_AD1]));
    const float rad = 0.5f * sqrtf(xd * xd + yd * yd) * 1.0001f + 1e-4f + 1e-6f * (fabsf(x) + fabsf(y));
    o[RB_RAD] = (abab > 0.f && adad > 0.f) ? rad : __int_as_float(0x7f800000);
}

// point_in_quadrilateral(pt, corners of Q) (rotate_iou.py:160-176)
__device__ __forceinline__ bool ri_in_quad(float px, float py, const float* __restrict__ q, float abab, float adad) {
    const float ap0 = __fsub_rn(px, q[RB_C + 0]), ap1 = __fsub_rn(py, q[RB_C + 1]);
    const float abap = __fmaf_rn(q[RB_AB1], ap1, __fmul_rn(q[RB_AB0], ap0));
    const float adap = __fmaf_rn(q[RB_AD1], ap1, __fmul_rn(q[RB_AD0], ap0));
    return abab >= abap && abap >= 0.f && adad >= adap && adap >= 0.f;
}

// line_segment_intersection (rotate_iou.py:74-118): edge A->B of pts1, edge C->D of pts2
__device__ __forceinline__ bool ri_segments(float A0, float A1, float B0, float B1, float C0, float C1, float D0, float D1, float& ox, float& oy) {
    const float BA0 = __fsub_rn(B0, A0), BA1 = __fsub_rn(B1, A1), DA0 = __fsub_rn(D0, A0), CA0 = __fsub_rn(C0, A0);
    const float DA1 = __fsub_rn(D1, A1), CA1 = __fsub_rn(C1, A1);
    const bool acd = __fmul_rn(DA1, CA0) > __fmul_rn(CA1, DA0);
    const bool bcd = __fmul_rn(__fsub_rn(D1, B1), __fsub_rn(C0, B0)) > __fmul_rn(__fsub_rn(C1, B1), __fsub_rn(D0, B0));
    if (acd == bcd) return false;
    const bool abc = __fmul_rn(CA1, BA0) > __fmul_rn(BA1, CA0);
    const bool abd = __fmul_rn(DA1, BA0) > __fmul_rn(BA1, DA0);
    if (abc == abd) return false;
    const float DC0 = __fsub_rn(D0, C0), DC1 = __fsub_rn(D1, C1);
    const float ABBA = __fmaf_rn(A0, B1, -__fmul_rn(B0, A1));
    const float CDDC = __fmaf_rn(C0, D1, -__fmul_rn(D0, C1));
    const float DH = __fmaf_rn(BA1, DC0, -__fmul_rn(BA0, DC1));
    const float Dx = __fmaf_rn(ABBA, DC0, -__fmul_rn(BA0, CDDC));
    const float Dy = __fmaf_rn(ABBA, DC1, -__fmul_rn(BA1, CDDC));
    ox = __fdiv_rn(Dx, DH);
    oy = __fdiv_rn(Dy, DH);
    return true;
}

// inter(rbox1 = query, rbox2 = box) (rotate_iou.py:231-246); vx / vy / vs: this thread's slots, element k at [k * stride]
__device__ double ri_inter(const float* __restrict__ q, const float* __restrict__ b, float* vx, float* vy, float* vs, int stride) {
    float p1[8], p2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { p1[i] = q[RB_C + i]; p2[i] = b[RB_C + i]; }
    const float abab1 = __fmaf_rn(q[RB_AB0], q[RB_AB0], __fmul_rn(q[RB_AB1], q[RB_AB1]));
    const float adad1 = __fmaf_rn(q[RB_AD0], q[RB_AD0], __fmul_rn(q[RB_AD1], q[RB_AD1]));
    const float abab2 = __fmaf_rn(b[RB_AB0], b[RB_AB0], __fmul_rn(b[RB_AB1], b[RB_AB1]));
    const float adad2 = __fmaf_rn(b[RB_AD0], b[RB_AD0], __fmul_rn(b[RB_AD1], b[RB_AD1]));
    int n = 0;
    // quadrilateral_intersection (:179-201): corners first, query corner i then box corner i, then the 16 edge pairs
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (ri_in_quad(p1[2 * i], p1[2 * i + 1], b, abab2, adad2)) { vx[n * stride] = p1[2 * i]; vy[n * stride] = p1[2 * i + 1]; ++n; }
        if (ri_in_quad(p2[2 * i], p2[2 * i + 1], q, abab1, adad1)) { vx[n * stride] = p2[2 * i]; vy[n * stride] = p2[2 * i + 1]; ++n; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float ox, oy;
            if (ri_segments(p1[2 * i], p1[2 * i + 1], p1[2 * ((i + 1) & 3)], p1[2 * ((i + 1) & 3) + 1],
                            p2[2 * j], p2[2 * j + 1], p2[2 * ((j + 1) & 3)], p2[2 * ((j + 1) & 3) + 1], ox, oy) && n < 8) {
                vx[n * stride] = ox; vy[n * stride] = oy; ++n;
            }
        }
    }
    if (n < 3) return 0.0;   // area() sums num_of_inter - 2 triangles; the sort does not change the count
    // sort_vertex_in_convex_polygon (:34-71)
    float c0 = 0.f, c1 = 0.f;
    for (int i = 0; i < n; ++i) { c0 = __fadd_rn(c0, vx[i * stride]); c1 = __fadd_rn(c1, vy[i * stride]); }
    c0 = (float)((double)c0 / (double)n);
    c1 = (float)((double)c1 / (double)n);
    for (int i = 0; i < n; ++i) {
        float v0 = __fsub_rn(vx[i * stride], c0), v1 = __fsub_rn(vy[i * stride], c1);
        const float d = __fsqrt_rn(__fmaf_rn(v0, v0, __fmul_rn(v1, v1)));
        v0 = __fdiv_rn(v0, d);
        v1 = __fdiv_rn(v1, d);
        if (v1 < 0.f) v0 = __fsub_rn(-2.0f, v0);
        vs[i * stride] = v0;
    }
    for (int i = 1; i < n; ++i) {
        if (vs[(i - 1) * stride] > vs[i * stride]) {
            const float temp = vs[i * stride], tx = vx[i * stride], ty = vy[i * stride];
            int j = i;
            while (j > 0 && vs[(j - 1) * stride] > temp) {
                vs[j * stride] = vs[(j - 1) * stride];
                vx[j * stride] = vx[(j - 1) * stride];
                vy[j * stride] = vy[(j - 1) * stride];
                --j;
            }
            vs[j * stride] = temp; vx[j * stride] = tx; vy[j * stride] = ty;
        }
    }
    // area (:23-31) over trangle_area (:17-20)
    const float a0 = vx[0], a1 = vy[0];
    double area = 0.0;
    float bx = vx[stride], by = vy[stride];
    for (int i = 0; i < n - 2; ++i) {
        const float cx = vx[(i + 2) * stride], cy = vy[(i + 2) * stride];
        const float t = __fmaf_rn(__fsub_rn(a0, cx), __fsub_rn(by, cy), -__fmul_rn(__fsub_rn(a1, cy), __fsub_rn(bx, cx)));
        area += fabs((double)t * 0.5);
        bx = cx; by = cy;
    }
    return area;
}

// devRotateIoUEval (rotate_iou.py:249-261)
__device__ __forceinline__ float ri_result(double inter, float area1, float area2, int criterion) {
    if (criterion == -1) return (float)(inter / ((double)__fadd_rn(area1, area2) - inter));
    if (criterion == 0) return (float)(inter / (double)area1);
    if (criterion == 1) return (float)(inter / (double)area2);
    return (float)inter;
}

struct RiSmem {
    float box[RI_TILE * RI_REC];
    float qry[RI_TILE * RI_REC];
    float vx[8 * RI_THREADS], vy[8 * RI_THREADS], vs[8 * RI_THREADS];
    unsigned short queue[RI_TILE * RI_TILE];
    int qcount;
};

// grid = (query tiles, box tiles, blocks): `blocks` > 1 is the blocked form -- independent (boxes, queries) groups, e.g. the
// frames of one evaluation part, each writing its own dense sub-matrix (offsets in box_off / qry_off / out_off).
__global__ void __launch_bounds__(RI_THREADS)
rotate_iou_eval_kernel(const float* __restrict__ boxes, int N, const float* __restrict__ query, int K, int criterion, float* __restrict__ out,
                       const int* __restrict__ box_off, const int* __restrict__ qry_off, const long long* __restrict__ out_off) {
    __shared__ RiSmem sm;
    const int tid = threadIdx.x;
    if (box_off) {   // blocked form: this grid slice works on group blockIdx.z
        const int g = blockIdx.z;
        const int b0 = box_off[g], q0 = qry_off[g];
        boxes += (size_t)b0 * 5; query += (size_t)q0 * 5; out += out_off[g];
        N = box_off[g + 1] - b0; K = qry_off[g + 1] - q0;
    }
    const int n0 = blockIdx.y * RI_TILE, k0 = blockIdx.x * RI_TILE;
    if (n0 >= N || k0 >= K) return;
    const int tn = min(RI_TILE, N - n0), tk = min(RI_TILE, K - k0);
    if (tid == 0) sm.qcount = 0;
    if (tid < tn) ri_prepare(boxes + (size_t)(n0 + tid) * 5, sm.box + tid * RI_REC);
    else if (tid >= RI_TILE && tid - RI_TILE < tk) ri_prepare(query + (size_t)(k0 + tid - RI_TILE) * 5, sm.qry + (tid - RI_TILE) * RI_REC);
    __syncthreads();
    // circle tests; a culled pair has intersection area 0 and still takes the reference's final division (0 / 0 stays NaN)
    const int c = tid & (RI_TILE - 1);
    if (c < tk) {
        const float* q = sm.qry + c * RI_REC;
        const float qx = q[RB_CX], qy = q[RB_CY], qr = q[RB_RAD], qa = q[RB_AREA];
        for (int r = tid >> 6; r < tn; r += RI_THREADS / RI_TILE) {
            const float* b = sm.box + r * RI_REC;
            const float dx = qx - b[RB_CX], dy = qy - b[RB_CY], rr = qr + b[RB_RAD];
            if (dx * dx + dy * dy > rr * rr) out[(size_t)(n0 + r) * K + k0 + c] = ri_result(0.0, qa, b[RB_AREA], criterion);
            else sm.queue[atomicAdd(&sm.qcount, 1)] = (unsigned short)(r * RI_TILE + c);
        }
    }
    __syncthreads();
    const int total = sm.qcount;
    for (int e = tid; e < total; e += RI_THREADS) {
        const int r = sm.queue[e] >> 6, cc = sm.queue[e] & (RI_TILE - 1);
        const float* q = sm.qry + cc * RI_REC;
        const float* b = sm.box + r * RI_REC;
        const double inter = ri_inter(q, b, sm.vx + tid, sm.vy + tid, sm.vs + tid, RI_THREADS);
        out[(size_t)(n0 + r) * K + k0 + cc] = ri_result(inter, q[RB_AREA], b[RB_AREA], criterion);
    }
}


// ---------------------------------------------------------------- CVAE recall IoU (aligned pairs, numpy float32 dialect)
// iou3d(gboxes, qboxes) of cvae_uncertainty/eval_utils/eval_utils.py:14-65: the recall of the CVAE's predicted boxes against
// their ground truth (:219-229).  The reference clamps the boxes to +-200, builds corners with torch elementwise ops
// (pcdet/utils/loss_utils.py:721-759) and then runs the same RRPN overlap as above as PYTHON loops over numpy float32
// scalars on the host (compute_vertex :276-411, sort_vertex :551-593, area_polygon :615-635) -- one pair at a time; its
// authors suggest commenting the call out (:197).  One thread per pair here.  Dialect: every float32 operation rounded
// separately (numpy scalars do not contract), corners of g tested against q first (with the extra abab > 0 && adad > 0),
// then corners of q against g, then the edge pairs with a cap of 8 vertices; vertices ordered by DESCENDING float32 angle
// (atan2 in float64 of the float32 unit vector, + 2 * 3.1415926 when negative) with numpy's argsort over all 8 slots
// (unused slots carry angle 0, NaN sorts last); float32 fan sum.  cos / sin come from this device's libdevice, as they do
// for the reference's CUDA tensors.
__device__ __forceinline__ float np_msub(float x, float y, float z, float w) { return __fsub_rn(__fmul_rn(x, y), __fmul_rn(z, w)); }
__device__ __forceinline__ float np_madd(float x, float y, float z, float w) { return __fadd_rn(__fmul_rn(x, y), __fmul_rn(z, w)); }
__device__ __forceinline__ float torch_clamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }   // NaN stays NaN
__device__ __forceinline__ float torch_min(float a, float b) { return (a < b || a != a) ? a : b; }                       // NaN propagates
__device__ __forceinline__ float torch_max(float a, float b) { return (a > b || a != a) ? a : b; }

// rbbox_to_corners (loss_utils.py:728-759) on [x, y, w, l, angle]
__device__ __forceinline__ void cv_corners(float x, float y, float w, float l, float ang, float* c) {
    const float cs = cosf(ang), sn = sinf(ang);
    const float dxcos = __fmul_rn(__fmul_rn(w, cs), 0.5f), dxsin = __fmul_rn(__fmul_rn(w, sn), 0.5f);
    const float dycos = __fmul_rn(__fmul_rn(l, cs), 0.5f), dysin = __fmul_rn(__fmul_rn(l, sn), 0.5f);
    c[0] = __fadd_rn(__fsub_rn(-dxcos, dysin), x); c[1] = __fadd_rn(__fsub_rn(dxsin, dycos), y);
    c[2] = __fadd_rn(__fadd_rn(-dxcos, dysin), x); c[3] = __fadd_rn(__fadd_rn(dxsin, dycos), y);
    c[4] = __fadd_rn(__fadd_rn(dxcos, dysin), x);  c[5] = __fadd_rn(__fadd_rn(-dxsin, dycos), y);
    c[6] = __fadd_rn(__fsub_rn(dxcos, dysin), x);  c[7] = __fadd_rn(__fsub_rn(-dxsin, dycos), y);
}

// corners of `p` inside quadrilateral `q` (compute_vertex steps 1 and 2)
__device__ __forceinline__ bool cv_in_quad(float px, float py, const float* q, bool need_positive) {
    const float ab0 = __fsub_rn(q[2], q[0]), ab1 = __fsub_rn(q[3], q[1]), ad0 = __fsub_rn(q[6], q[0]), ad1 = __fsub_rn(q[7], q[1]);
    const float ap0 = __fsub_rn(px, q[0]), ap1 = __fsub_rn(py, q[1]);
    const float abab = np_madd(ab0, ab0, ab1, ab1), abap = np_madd(ab0, ap0, ab1, ap1);
    const float adad = np_madd(ad0, ad0, ad1, ad1), adap = np_madd(ad0, ap0, ad1, ap1);
    const bool in = abab >= abap && abap >= 0.f && adad >= adap && adap >= 0.f;
    return need_positive ? (in && adad > 0.f && abab > 0.f) : in;
}

constexpr int CV_THREADS = 128;

__global__ void __launch_bounds__(CV_THREADS)
cvae_iou3d_kernel(const float* __restrict__ gboxes, const float* __restrict__ qboxes, int n, float* __restrict__ ious) {
    __shared__ float s_vx[8 * CV_THREADS], s_vy[8 * CV_THREADS], s_ang[8 * CV_THREADS];
    const int tid = threadIdx.x, i = blockIdx.x * CV_THREADS + tid;
    if (i >= n) return;
    float* vx = s_vx + tid; float* vy = s_vy + tid; float* ang = s_ang + tid;
    constexpr int S = CV_THREADS;
    float g[7], q[7];
#pragma unroll
    for (int f = 0; f < 7; ++f) { g[f] = torch_clamp(gboxes[(size_t)i * 7 + f], -200.f, 200.f); q[f] = torch_clamp(qboxes[(size_t)i * 7 + f], -200.f, 200.f); }
    float cg[8], cq[8];
    cv_corners(g[0], g[1], g[3], g[4], g[6], cg);
    cv_corners(q[0], q[1], q[3], q[4], q[6], cq);
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { vx[k * S] = 0.f; vy[k * S] = 0.f; }       // intersections = np.zeros((N, 16)): unused slots are read by the argsort gather
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (cv_in_quad(cg[2 * k], cg[2 * k + 1], cq, true)) { vx[cnt * S] = cg[2 * k]; vy[cnt * S] = cg[2 * k + 1]; ++cnt; }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (cv_in_quad(cq[2 * k], cq[2 * k + 1], cg, false)) { vx[cnt * S] = cq[2 * k]; vy[cnt * S] = cq[2 * k + 1]; ++cnt; }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float A0 = cg[2 * a], A1 = cg[2 * a + 1], B0 = cg[2 * ((a + 1) & 3)], B1 = cg[2 * ((a + 1) & 3) + 1];
            const float C0 = cq[2 * b], C1 = cq[2 * b + 1], D0 = cq[2 * ((b + 1) & 3)], D1 = cq[2 * ((b + 1) & 3) + 1];
            const float BA0 = __fsub_rn(B0, A0), BA1 = __fsub_rn(B1, A1), CA0 = __fsub_rn(C0, A0), CA1 = __fsub_rn(C1, A1);
            const float DA0 = __fsub_rn(D0, A0), DA1 = __fsub_rn(D1, A1);
            const bool acd = __fmul_rn(DA1, CA0) > __fmul_rn(CA1, DA0);
            const bool bcd = __fmul_rn(__fsub_rn(D1, B1), __fsub_rn(C0, B0)) > __fmul_rn(__fsub_rn(C1, B1), __fsub_rn(D0, B0));
            if (acd == bcd) continue;
            const bool abc = __fmul_rn(CA1, BA0) > __fmul_rn(BA1, CA0);
            const bool abd = __fmul_rn(DA1, BA0) > __fmul_rn(BA1, DA0);
            if (abc == abd) continue;
            if (cnt > 7) continue;                                        // loss_utils.py:383-397
            const float DC0 = __fsub_rn(D0, C0), DC1 = __fsub_rn(D1, C1);
            const float ABBA = np_msub(A0, B1, B0, A1), CDDC = np_msub(C0, D1, D0, C1);
            const float DH = np_msub(BA1, DC0, BA0, DC1);
            vx[cnt * S] = __fdiv_rn(np_msub(ABBA, DC0, BA0, CDDC), DH);
            vy[cnt * S] = __fdiv_rn(np_msub(ABBA, DC1, BA1, CDDC), DH);
            ++cnt;
        }
    }
    float area = 0.f;
    if (cnt > 2) {
        // sort_vertex: float32 centroid, float64 atan2 of the float32 unit vector, angles stored as float32
        float c0 = 0.f, c1 = 0.f;
        for (int k = 0; k < cnt; ++k) { c0 = __fadd_rn(c0, vx[k * S]); c1 = __fadd_rn(c1, vy[k * S]); }
        c0 = __fdiv_rn(c0, (float)cnt); c1 = __fdiv_rn(c1, (float)cnt);
#pragma unroll
        for (int k = 0; k < 8; ++k) ang[k * S] = 0.f;
        for (int k = 0; k < cnt; ++k) {
            float v0 = __fsub_rn(vx[k * S], c0), v1 = __fsub_rn(vy[k * S], c1);
            const float d = (float)sqrt((double)np_madd(v0, v0, v1, v1));   // math.sqrt in float64; numpy 2 divides float32 by the float32 value of it
            v0 = __fdiv_rn(v0, d); v1 = __fdiv_rn(v1, d);
            const double a = atan2((double)v1, (double)v0);
            ang[k * S] = a < 0.0 ? (float)(a + 2 * 3.1415926) : (float)a;
        }
        // np.argsort(-angle) over the 8 slots (stable for 8 elements, NaN last): rank of slot k = slots that sort before it
        float px[8], py[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float ak = ang[k * S];
            int rank = 0;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float am = ang[m * S];
                const bool m_nan = am != am, k_nan = ak != ak;
                const bool before = k_nan ? (!m_nan || m < k) : (!m_nan && (am > ak || (am == ak && m < k)));
                rank += (m != k && before) ? 1 : 0;
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) if (rank == r) { px[r] = vx[k * S]; py[r] = vy[k * S]; }
        }
        // area_polygon: float32 fan from the first sorted vertex
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            if (k < cnt - 2) {
                const float t = np_msub(__fsub_rn(px[0], px[k + 2]), __fsub_rn(py[k + 1], py[k + 2]), __fsub_rn(py[0], py[k + 2]), __fsub_rn(px[k + 1], px[k + 2]));
                area = __fadd_rn(area, fabsf(__fmul_rn(t, 0.5f)));
            }
        }
    }
    // eval_utils.py:48-64
    const float top = torch_min(__fadd_rn(g[2], __fmul_rn(0.5f, g[5])), __fadd_rn(q[2], __fmul_rn(0.5f, q[5])));
    const float bot = torch_max(__fsub_rn(g[2], __fmul_rn(0.5f, g[5])), __fsub_rn(q[2], __fmul_rn(0.5f, q[5])));
    float inter_h = __fsub_rn(top, bot);
    if (inter_h < 0.f) inter_h = 0.f;
    const float vg = __fmul_rn(__fmul_rn(g[3], g[4]), g[5]), vq = __fmul_rn(__fmul_rn(q[3], q[4]), q[5]);
    const float inc = __fmul_rn(inter_h, area);
    ious[i] = __fdiv_rn(inc, __fsub_rn(__fadd_rn(vg, vq), inc));
}

}  // namespace glenet

using namespace glenet;

extern "C" {

int glenet_rotate_iou_eval_gpu(const float* boxes, int n, const float* query_boxes, int k, int criterion, float* iou, glenet_stream_t s) {
    const char* what = "glenet_rotate_iou_eval_gpu";
    if (n < 0 || k < 0) return fail(GLENET_EINVAL, "%s: negative box count", what);
    if (n == 0 || k == 0) return GLENET_OK;
    if (!boxes || !query_boxes || !iou) return fail(GLENET_EINVAL, "%s: null pointer", what);
    const dim3 grid((k + RI_TILE - 1) / RI_TILE, (n + RI_TILE - 1) / RI_TILE, 1);
    if (grid.y > 65535u) return fail(GLENET_EINVAL, "%s: more than 4 194 240 boxes", what);
    rotate_iou_eval_kernel<<<grid, RI_THREADS, 0, (cudaStream_t)s>>>(boxes, n, query_boxes, k, criterion, iou, nullptr, nullptr, nullptr);
    return check_launch(what);
}

int glenet_rotate_iou_eval_blocks_gpu(const float* boxes, const int* box_offsets, const float* query_boxes, const int* query_offsets,
                                      const long long* out_offsets, int groups, int max_boxes, int max_queries, int criterion, float* iou,
                                      glenet_stream_t s) {
    const char* what = "glenet_rotate_iou_eval_blocks_gpu";
    if (groups < 0 || max_boxes < 0 || max_queries < 0) return fail(GLENET_EINVAL, "%s: negative count", what);
    if (groups == 0 || max_boxes == 0 || max_queries == 0) return GLENET_OK;
    if (!boxes || !query_boxes || !iou || !box_offsets || !query_offsets || !out_offsets) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (groups > 65535) return fail(GLENET_EINVAL, "%s: more than 65535 groups", what);
    const dim3 grid((max_queries + RI_TILE - 1) / RI_TILE, (max_boxes + RI_TILE - 1) / RI_TILE, groups);
    if (grid.y > 65535u) return fail(GLENET_EINVAL, "%s: group too large", what);
    rotate_iou_eval_kernel<<<grid, RI_THREADS, 0, (cudaStream_t)s>>>(boxes, 0, query_boxes, 0, criterion, iou, box_offsets, query_offsets, out_offsets);
    return check_launch(what);
}

int glenet_cvae_iou3d_gpu(const float* gboxes, const float* qboxes, int n, float* ious, glenet_stream_t s) {
    const char* what = "glenet_cvae_iou3d_gpu";
    if (n < 0) return fail(GLENET_EINVAL, "%s: negative box count", what);
    if (n == 0) return GLENET_OK;
    if (!gboxes || !qboxes || !ious) return fail(GLENET_EINVAL, "%s: null pointer", what);
    cvae_iou3d_kernel<<<(n + CV_THREADS - 1) / CV_THREADS, CV_THREADS, 0, (cudaStream_t)s>>>(gboxes, qboxes, n, ious);
    return check_launch(what);
}

}  // extern "C"
