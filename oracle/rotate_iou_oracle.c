/*
 * rotate_iou_oracle.c -- CPU restatement of the KITTI evaluator's rotated IoU
 * (pcdet/datasets/kitti/kitti_object_eval_python/rotate_iou.py, a numba-CUDA module; SURVEY.md
 * section 8f rank 4, called from kitti_object_eval_python/eval.py:117,151).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as geom_oracle.c).  The product kernel it checks is
 * glenet_b200/csrc/rotate_iou.cu (glenet_rotate_iou_eval_gpu).
 *
 * Parity status: PINNED WITHIN TOLERANCE.  tests/test_oracle.py compares this restatement with
 * tests/golden/rotate_iou_golden.npz, produced by importing the reference file verbatim under numba's CUDA
 * simulator (tests/golden/make_golden_rotate_iou.py; no GPU in the build container).  The simulator
 * evaluates scalar intermediates with NumPy promotion rules, the real numba-CUDA kernel with numba's typing
 * (float32 x float32 -> float32, float32 x Python-float literal -> float64); this file follows numba's
 * typing.  Two dialects (argument `contract` of oracle_rotate_iou_eval_dialect):
 *   0  every operation rounded separately -- what the simulator computes up to its NumPy typing; 1e-5 absolute vs the
 *      simulator goldens;
 *   1  the FMA contraction of the kernel numba 0.65 / NVVM / ptxas 12.9 emit for sm_100a, read from its SASS (every
 *      x*y - z*w is fma(x, y, -rn(z*w)); the sums of two products fuse the product noted at each site) -- BIT parity vs
 *      tests/golden/rotate_iou_gpu_golden.npz, which the reference kernel itself produced on a B200
 *      (tests/golden/make_golden_rotate_iou.py gpu).
 *
 * Box format (rotate_iou.py:289-292): [x, y, x_d, y_d, angle], angle clockwise when positive.
 */
#include <math.h>
#include <stdint.h>

static int g_contract = 0;   /* dialect of the current call (the library is used single-threaded by the tests) */
/* x*y - z*w and x*y + z*w with the first product fused when contracting */
static float msub(float x, float y, float z, float w) { return g_contract ? fmaf(x, y, -(z * w)) : x * y - z * w; }
static float madd(float x, float y, float z, float w) { return g_contract ? fmaf(x, y, z * w) : x * y + z * w; }

/* rotate_iou.py:17-20 -- float32 expression, then "/ 2.0" in float64 (exact) */
static double trangle_area(const float* a, const float* b, const float* c) {
    const float v = msub(a[0] - c[0], b[1] - c[1], a[1] - c[1], b[0] - c[0]);
    return (double)v / 2.0;
}

/* rotate_iou.py:23-31 */
static double area(const float* int_pts, int num_of_inter) {
    double area_val = 0.0;
    for (int i = 0; i < num_of_inter - 2; ++i)
        area_val += fabs(trangle_area(int_pts, int_pts + 2 * i + 2, int_pts + 2 * i + 4));
    return area_val;
}

/* rotate_iou.py:34-71 -- pseudo-angle about the centroid, insertion sort */
static void sort_vertex_in_convex_polygon(float* int_pts, int num_of_inter) {
    if (num_of_inter <= 0) return;
    float center[2] = {0.f, 0.f};
    for (int i = 0; i < num_of_inter; ++i) { center[0] += int_pts[2 * i]; center[1] += int_pts[2 * i + 1]; }
    center[0] = (float)((double)center[0] / (double)num_of_inter);   /* float32 /= int32 goes through float64 */
    center[1] = (float)((double)center[1] / (double)num_of_inter);
    float v[2], vs[16];
    for (int i = 0; i < num_of_inter; ++i) {
        v[0] = int_pts[2 * i] - center[0];
        v[1] = int_pts[2 * i + 1] - center[1];
        const float d = sqrtf(madd(v[0], v[0], v[1], v[1]));
        v[0] = v[0] / d;
        v[1] = v[1] / d;
        if (v[1] < 0) v[0] = (float)(-2.0 - (double)v[0]);
        vs[i] = v[0];
    }
    for (int i = 1; i < num_of_inter; ++i) {
        if (vs[i - 1] > vs[i]) {
            const float temp = vs[i], tx = int_pts[2 * i], ty = int_pts[2 * i + 1];
            int j = i;
            while (j > 0 && vs[j - 1] > temp) {
                vs[j] = vs[j - 1];
                int_pts[j * 2] = int_pts[j * 2 - 2];
                int_pts[j * 2 + 1] = int_pts[j * 2 - 1];
                --j;
            }
            vs[j] = temp; int_pts[j * 2] = tx; int_pts[j * 2 + 1] = ty;
        }
    }
}

/* rotate_iou.py:74-118 */
static int line_segment_intersection(const float* pts1, const float* pts2, int i, int j, float* temp_pts) {
    const float A0 = pts1[2 * i], A1 = pts1[2 * i + 1];
    const float B0 = pts1[2 * ((i + 1) % 4)], B1 = pts1[2 * ((i + 1) % 4) + 1];
    const float C0 = pts2[2 * j], C1 = pts2[2 * j + 1];
    const float D0 = pts2[2 * ((j + 1) % 4)], D1 = pts2[2 * ((j + 1) % 4) + 1];
    const float BA0 = B0 - A0, BA1 = B1 - A1, DA0 = D0 - A0, CA0 = C0 - A0, DA1 = D1 - A1, CA1 = C1 - A1;
    const int acd = DA1 * CA0 > CA1 * DA0;
    const int bcd = (D1 - B1) * (C0 - B0) > (C1 - B1) * (D0 - B0);
    if (acd != bcd) {
        const int abc = CA1 * BA0 > BA1 * CA0;
        const int abd = DA1 * BA0 > BA1 * DA0;
        if (abc != abd) {
            const float DC0 = D0 - C0, DC1 = D1 - C1;
            const float ABBA = msub(A0, B1, B0, A1), CDDC = msub(C0, D1, D0, C1);
            const float DH = msub(BA1, DC0, BA0, DC1);
            const float Dx = msub(ABBA, DC0, BA0, CDDC), Dy = msub(ABBA, DC1, BA1, CDDC);
            temp_pts[0] = Dx / DH;
            temp_pts[1] = Dy / DH;
            return 1;
        }
    }
    return 0;
}

/* rotate_iou.py:160-176 */
static int point_in_quadrilateral(float pt_x, float pt_y, const float* corners) {
    const float ab0 = corners[2] - corners[0], ab1 = corners[3] - corners[1];
    const float ad0 = corners[6] - corners[0], ad1 = corners[7] - corners[1];
    const float ap0 = pt_x - corners[0], ap1 = pt_y - corners[1];
    /* contracted: abab = fma(ab0, ab0, ab1*ab1) but abap = fma(ab1, ap1, ab0*ap0) -- the SECOND product is the fused one */
    const float abab = madd(ab0, ab0, ab1, ab1), abap = madd(ab1, ap1, ab0, ap0);
    const float adad = madd(ad0, ad0, ad1, ad1), adap = madd(ad1, ap1, ad0, ap0);
    return abab >= abap && abap >= 0 && adad >= adap && adap >= 0;
}

/* rotate_iou.py:179-201 -- corners first (1-in-2, 2-in-1 interleaved), then the 16 edge pairs */
static int quadrilateral_intersection(const float* pts1, const float* pts2, float* int_pts) {
    int n = 0;
    for (int i = 0; i < 4; ++i) {
        if (point_in_quadrilateral(pts1[2 * i], pts1[2 * i + 1], pts2)) { int_pts[n * 2] = pts1[2 * i]; int_pts[n * 2 + 1] = pts1[2 * i + 1]; ++n; }
        if (point_in_quadrilateral(pts2[2 * i], pts2[2 * i + 1], pts1)) { int_pts[n * 2] = pts2[2 * i]; int_pts[n * 2 + 1] = pts2[2 * i + 1]; ++n; }
    }
    float temp_pts[2];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (n < 8 && line_segment_intersection(pts1, pts2, i, j, temp_pts)) {   /* int_pts holds 8 vertices (:235); see note below */
                int_pts[n * 2] = temp_pts[0]; int_pts[n * 2 + 1] = temp_pts[1]; ++n;
            }
    return n;
}
/* Note: the reference's intersection_corners is 16 floats = 8 vertices (rotate_iou.py:235) and it appends without
 * a bound; two convex quadrilaterals have at most 8 intersection vertices in exact arithmetic, and the guard above
 * only keeps the restatement memory-safe where rounding would produce a ninth. */

/* rotate_iou.py:204-228 */
static void rbbox_to_corners(float* corners, const float* rbbox, const float* trig) {
    const float angle = rbbox[4];
    /* trig = {cos, sin} as the device's libdevice computed them (the goldens made on the GPU carry the table); else host libm */
    const float a_cos = trig ? trig[0] : cosf(angle), a_sin = trig ? trig[1] : sinf(angle);
    const float center_x = rbbox[0], center_y = rbbox[1], x_d = rbbox[2], y_d = rbbox[3];
    const float corners_x[4] = {-x_d / 2, -x_d / 2, x_d / 2, x_d / 2};
    const float corners_y[4] = {-y_d / 2, y_d / 2, y_d / 2, -y_d / 2};
    for (int i = 0; i < 4; ++i) {
        corners[2 * i] = madd(a_cos, corners_x[i], a_sin, corners_y[i]) + center_x;
        corners[2 * i + 1] = msub(a_cos, corners_y[i], a_sin, corners_x[i]) + center_y;   /* (-a_sin) * x + a_cos * y */
    }
}

/* rotate_iou.py:231-246 */
static double inter(const float* rbbox1, const float* rbbox2, const float* trig1, const float* trig2) {
    float corners1[8], corners2[8], intersection_corners[16];
    rbbox_to_corners(corners1, rbbox1, trig1);
    rbbox_to_corners(corners2, rbbox2, trig2);
    const int num = quadrilateral_intersection(corners1, corners2, intersection_corners);
    sort_vertex_in_convex_polygon(intersection_corners, num);
    return area(intersection_corners, num);
}

/* rotate_iou.py:249-261 */
static float dev_rotate_iou_eval(const float* rbox1, const float* rbox2, int criterion, const float* trig1, const float* trig2) {
    const float area1 = rbox1[2] * rbox1[3], area2 = rbox2[2] * rbox2[3];
    const double area_inter = inter(rbox1, rbox2, trig1, trig2);
    if (criterion == -1) return (float)(area_inter / ((double)(area1 + area2) - area_inter));
    if (criterion == 0) return (float)(area_inter / (double)area1);
    if (criterion == 1) return (float)(area_inter / (double)area2);
    return (float)area_inter;
}

/* rotate_iou_gpu_eval (rotate_iou.py:263-330): iou[n][k] = devRotateIoUEval(query_boxes[k], boxes[n], criterion)
 * -- note the argument order of the kernel (:281-283): the QUERY box is rbox1. */
void oracle_rotate_iou_eval_dialect(const float* boxes, int N, const float* query_boxes, int K, int criterion, float* iou, int contract,
                                    const float* trig_boxes, const float* trig_query) {
    g_contract = contract;
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k)
            iou[(int64_t)n * K + k] = dev_rotate_iou_eval(query_boxes + (int64_t)k * 5, boxes + (int64_t)n * 5, criterion,
                                                          trig_query ? trig_query + 2 * (int64_t)k : 0, trig_boxes ? trig_boxes + 2 * (int64_t)n : 0);
}

void oracle_rotate_iou_eval(const float* boxes, int N, const float* query_boxes, int K, int criterion, float* iou) {
    oracle_rotate_iou_eval_dialect(boxes, N, query_boxes, K, criterion, iou, 0, 0, 0);
}
