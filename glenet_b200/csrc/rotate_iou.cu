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

}  // extern "C"
