// Host-side helpers of the C ABI (no device code).
//
// The reference's CPU paths (pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:74-84,146-151 and
// pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-125) take cos/sin from the host's libm.
// glibc's float trig is not bit-identical to CUDA's libdevice, and the reference's margin
// predicates are discontinuous in those bits, so when a `_cpu` entry point is executed on the
// GPU ("CPU dialect") the O(N) per-box trigonometry is evaluated here, with the very libm the
// reference would have used on this host, and shipped to the device as a small table.
#include <math.h>
#include <stddef.h>

#include "../../include/glenet_geom.h"

extern "C" {

// out: (n, 4) rows {cos(h), sin(h), cos(-h), sin(-h)}; boxes: (n, 7) host floats
void glenet_host_trig4_strided(const float* angles_host, int stride, int n, float* out_host) {
    for (int i = 0; i < n; ++i) {
        const float h = angles_host[(size_t)i * stride];
        float s, c;
        sincosf(h, &s, &c);
        out_host[4 * (size_t)i + 0] = c;
        out_host[4 * (size_t)i + 1] = s;
        out_host[4 * (size_t)i + 2] = cosf(-h);
        out_host[4 * (size_t)i + 3] = sinf(-h);
    }
}

void glenet_host_trig4(const float* boxes_host, int n, float* out_host) {
    for (int i = 0; i < n; ++i) {
        const float h = boxes_host[(size_t)i * 7 + 6];
        float s, c;
        sincosf(h, &s, &c);               // box_overlap: cos(a_angle), sin(a_angle)
        out_host[4 * (size_t)i + 0] = c;
        out_host[4 * (size_t)i + 1] = s;
        out_host[4 * (size_t)i + 2] = cosf(-h);   // check_in_box2d: cos(-box[6]), sin(-box[6])
        out_host[4 * (size_t)i + 3] = sinf(-h);
    }
}

// out: (n, 2) rows {cos(-h), sin(-h)} (lidar_to_local_coords_cpu)
void glenet_host_trig2(const float* boxes_host, int n, float* out_host) {
    for (int i = 0; i < n; ++i) {
        const float h = boxes_host[(size_t)i * 7 + 6];
        out_host[2 * (size_t)i + 0] = cosf(-h);
        out_host[2 * (size_t)i + 1] = sinf(-h);
    }
}

}  // extern "C"
