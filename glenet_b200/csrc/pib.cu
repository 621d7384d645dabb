// points_in_boxes for sm_100a.
//
// Replaces points_in_boxes_kernel (+ launcher) of
// pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-36,313-359 and, in the CPU
// dialect, points_in_boxes_cpu of pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-168.
//
// The reference walks all N boxes for every point and re-evaluates cosf/sinf per
// (point, box).  Exhaustively that is ~2000 instructions per point for N = 200 -- compute
// bound at < 10 % of what HBM can stream (16 B / point).  Here:
//
//   build kernel (1 CTA / frame): per-box record {cx, cy, cz, cos(-h), sin(-h), tx, ty, tz}
//       with the reference's FP64 comparisons folded into directed-rounded FP32 thresholds
//       (exactly equivalent, see box_thresholds()); a fine map of 128 x 128 cells x 16 z slabs
//       (one bit per cell and slab of the frame's z window that a padded footprint touches); and
//       a coarse G x G grid whose cells hold up to four candidate boxes inline (a crowded cell
//       holds the offset and length of its slice of a candidate list instead).
//   query kernel: points stream through coalesced; a point whose (cell, z slab) bit is clear is
//       background at once; the others look up their coarse cell in shared memory and run the
//       exact reference predicate only against that cell's candidates, keeping the minimum
//       index (= first hit of the reference's ascending loop with `break`).
// Frames whose boxes cannot be binned (non-finite extents, list overflow) fall back, on the
// device and per frame, to the exhaustive loop with the same predicate -- never to the host.
#include "common.cuh"
#include "../../include/glenet_geom.h"
#include <float.h>
#include <math.h>
#include <atomic>

namespace glenet {

constexpr int PIB_G = 64;                       // coarse grid (CSR candidate lists), cells per axis
constexpr int PIB_CELLS = PIB_G * PIB_G;
#ifndef GLENET_PIB_ZSLABS        // 16: fine map = 128 x 128 cells x 16 z-slab bits (a point is hot only if a box covers its cell AND its
#define GLENET_PIB_ZSLABS 16     // z slab); 0: 256 x 256 cells x 1 bit + one z window for the whole frame (the round-1 layout, kept for A/B timing)
#endif
#if GLENET_PIB_ZSLABS
constexpr int PIB_FG = 128;                     // fine occupancy map, cells per axis
constexpr int PIB_ZS = 16;                      // z slabs of the frame's z window = bits per cell
constexpr int PIB_FWORDS = PIB_FG * PIB_FG * PIB_ZS / 32;
static_assert(GLENET_PIB_ZSLABS == 16, "the fine map holds 16-bit cells");
#else
constexpr int PIB_FG = 256;                     // fine occupancy bitmap, cells per axis (1 bit per cell)
constexpr int PIB_FWORDS = PIB_FG * PIB_FG / 32;
#endif
constexpr int PIB_ROUND = 4 * 256;              // points per CTA round of the direct kernel (4 per thread)
constexpr int PIB_WBATCH = 128;                 // points per warp batch (4 per lane)
constexpr int PIB_WQ = PIB_WBATCH + 32;         // warp queue: one batch of hot points + a remainder of < 32
#ifndef GLENET_PIB_CHUNK
#define GLENET_PIB_CHUNK 1024
#endif
#ifndef GLENET_PIB_RUNS          // 1: a CTA's chunks of one frame are streamed as one run (no pipeline refill per chunk)
#define GLENET_PIB_RUNS 1
#endif
#ifndef GLENET_PIB_DBG           // timing experiments only (results invalid): 1 no raster, 2 no coarse cells, 4 no scan, 8 no cell pack, 16 no query launch
#define GLENET_PIB_DBG 0
#endif
#ifndef GLENET_PIB_ZWINDOW       // 1: points outside the z window of all boxes of the frame are cold without a table lookup result
#define GLENET_PIB_ZWINDOW 1
#endif
#ifndef GLENET_PIB_L2PF          // > 0: L2 prefetch of the points this many batches ahead of the register prefetch
#define GLENET_PIB_L2PF 3
#endif
#ifndef GLENET_PIB_CTAS          // resident CTAs per SM the query kernel is compiled and launched for
#define GLENET_PIB_CTAS (GLENET_PIB_ZSLABS ? 2 : 3)
#endif
#ifndef GLENET_PIB_PF            // register prefetch depth of the query's point stream: 1 or 2 batches
#define GLENET_PIB_PF 1
#endif
constexpr int PIB_CHUNK = GLENET_PIB_CHUNK;     // points per work item of the query kernel
#ifndef GLENET_PIB_THREADS
#define GLENET_PIB_THREADS 256
#endif
constexpr int PIB_THREADS = GLENET_PIB_THREADS;
constexpr int PIB_DIRECT_BOXES = 32;            // at most this many boxes => single-launch direct kernel for small calls
constexpr int PIB_DIRECT_PTS = 4096;            // points per CTA of the direct kernel
constexpr long PIB_DIRECT_MAX_POINTS = 1 << 20; // ... when the whole call has at most this many points
#ifndef GLENET_PIB_BUILD_THREADS
#define GLENET_PIB_BUILD_THREADS 1024
#endif
constexpr int PIB_BUILD_THREADS = GLENET_PIB_BUILD_THREADS;
constexpr size_t PIB_BUILD_SMEM = (size_t)PIB_CELLS * (8 + 4 + 4) + (size_t)PIB_FWORDS * 4;   // dynamic shared memory of the build kernel
#ifndef GLENET_PIB_REC_STRIDE     // floats per box record in the query kernel's shared memory: 12 spreads random records over all banks
#define GLENET_PIB_REC_STRIDE 12  // (8 = packed: records k and k + 4 share their banks)
#endif
constexpr int PIB_REC_STRIDE = GLENET_PIB_REC_STRIDE;
constexpr int PIB_SMEM_BOXES = 224;             // box records cached in shared memory by the query kernel

struct PibFrame {          // 48 B header per frame
    float gx0, gy0, inv_x, inv_y;   // coarse mapping: cell = floor((x - gx0) * inv_x)
    int exhaustive;        // 1 => query kernel loops over all boxes
    int list_len;          // < 0 => no box of the frame can contain any point
    float finv_x, finv_y;  // fine bitmap mapping
    float zc, zh;          // z window of all boxes together: |z - zc| > zh => the point is in no box (zh = +inf: no window);
                           // with z slabs: slab = floor(fma(z, zc, zh)) (zc = slabs per metre, 0 = no window: every point in slab 0)
    float foff_x, foff_y;  // fine bitmap mapping, fused form: cell = floor(fma(x, finv_x, foff_x)), foff = -gx0 * finv_x
};

__host__ __device__ inline size_t pib_list_cap(int n) { return (size_t)32 * n + 2 * PIB_CELLS; }

struct PibWorkspace {
    PibFrame* frames;      // [B]
    float* rec;            // [B][N][8]
    unsigned int* list;    // [B][cap]
    unsigned int* bits;    // [B][PIB_FWORDS] fine occupancy bitmap
    unsigned long long* cells;   // [B][PIB_CELLS] packed coarse cell: four candidate ids, 16 bits each (0xffff none, last = 0xfffe: walk the list)
    size_t cap;
    size_t bytes;
};

__host__ __device__ inline PibWorkspace pib_layout(void* base, int B, int N) {
    PibWorkspace w;
    size_t off = 0;
    unsigned char* p = (unsigned char*)base;
    w.frames = (PibFrame*)(p + off); off += ((size_t)B * sizeof(PibFrame) + 15) / 16 * 16;
    w.rec = (float*)(p + off);       off += ((size_t)B * N * 8 * sizeof(float) + 15) / 16 * 16;
    w.cap = pib_list_cap(N);
    w.list = (unsigned int*)(p + off);  off += ((size_t)B * w.cap * sizeof(unsigned int) + 15) / 16 * 16;
    w.bits = (unsigned int*)(p + off);  off += ((size_t)B * PIB_FWORDS * sizeof(unsigned int) + 15) / 16 * 16;
    w.cells = (unsigned long long*)(p + off);  off += ((size_t)B * PIB_CELLS * sizeof(unsigned long long) + 15) / 16 * 16;
    w.bytes = off;
    return w;
}

// The reference predicate (check_pt_in_box3d, roiaware_pool3d_kernel.cu:23-36), as compiled:
//   skip   if (double)|z - cz| >  (double)dz * 0.5
//   inside if (double)|lx|     <  fma((double)dx, 0.5, (double)MARGIN)   (same for y)
// For a float a and a double t:  a > t  <=>  a > round_down_to_float(t)
//                                a < t  <=>  a < round_up_to_float(t)
// so the three thresholds are rounded once per box and the per-point test is pure FP32,
// bit-identical to the FP64 comparisons (NaN thresholds keep the same truth values).
template <bool CPU_DIALECT>
__device__ __forceinline__ void box_thresholds(float dx, float dy, float dz, float& tx, float& ty, float& tz) {
    const double margin = CPU_DIALECT ? (double)1e-2f : (double)1e-5f;   // `const float MARGIN` promoted
    tx = __double2float_ru(fma((double)dx, 0.5, margin));
    ty = __double2float_ru(fma((double)dy, 0.5, margin));
    tz = __double2float_rd((double)dz * 0.5);
}

// lidar_to_local_coords (:16-20) as compiled for the GPU: lx = fma(sx, c, -(sy*s)), ly = fma(sy, c, sx*s)
// The record is read as two 16-byte words (two LDS.128 / LDG.128, not eight scalar loads).
__device__ __forceinline__ bool pt_in_box_gpu(float x, float y, float z, const float* __restrict__ r) {
    const float4 r0 = *reinterpret_cast<const float4*>(r), r1 = *reinterpret_cast<const float4*>(r + 4);
    const float sx = __fsub_rn(x, r0.x), sy = __fsub_rn(y, r0.y);
    const float c = r0.w, s = r1.x;
    const float lx = __fmaf_rn(sx, c, -__fmul_rn(sy, s));
    const float ly = __fmaf_rn(sy, c, __fmul_rn(sx, s));
    return !(fabsf(__fsub_rn(z, r0.z)) > r1.w) & (r1.y > fabsf(lx)) & (r1.z > fabsf(ly));
}

// Footprint of a box in world coordinates: padded AABB of the region where the predicate can hold.
struct Footprint {
    float cx, cy, c, s, tx, ty, pad, x0, x1, y0, y1;
    bool never, bad;
};
__device__ __forceinline__ Footprint footprint(const float* __restrict__ r) {
    Footprint f;
    f.cx = r[0]; f.cy = r[1]; f.c = r[3]; f.s = r[4]; f.tx = r[5]; f.ty = r[6];
    // `tx > |lx|` needs tx > 0; NaN centre / heading / extent => the predicate is never true
    f.never = !(f.tx > 0.f) || !(f.ty > 0.f) || (f.c != f.c) || (f.s != f.s) || !(fabsf(f.cx) <= FLT_MAX) || !(fabsf(f.cy) <= FLT_MAX);
    const float hx = fabsf(f.c) * f.tx + fabsf(f.s) * f.ty, hy = fabsf(f.s) * f.tx + fabsf(f.c) * f.ty;
    f.pad = 2e-3f + 1e-6f * (fabsf(f.cx) + fabsf(f.cy) + hx + hy);
    f.x0 = f.cx - hx - f.pad; f.x1 = f.cx + hx + f.pad; f.y0 = f.cy - hy - f.pad; f.y1 = f.cy + hy + f.pad;
    f.bad = !(fabsf(f.x0) <= FLT_MAX) || !(fabsf(f.x1) <= FLT_MAX) || !(fabsf(f.y0) <= FLT_MAX) || !(fabsf(f.y1) <= FLT_MAX);
    return f;
}

// Visit every cell of a G x G grid (origin gx0/gy0, 1/cell = inv) that the footprint may touch.
// The range uses the very mapping of the query kernel
// (monotone => every point of [x0, x1] lands in [ix0, ix1]); a separating-axis test in the box frame
// then drops the corner cells of rotated boxes.
// Fine-bitmap rasterisation of one footprint, row by row: the x-extent of (padded rotated
// rectangle) n (horizontal strip of the row) comes from clipping the rectangle's four edges to the strip
// -- O(rows) instead of O(cells) work -- and the row's bits are set word-wise.  Everything is padded
// outwards (strip and span), so the bitmap is a superset of the cells any inside point can map to.
__device__ __forceinline__ void raster_fine_rows(const Footprint& f, float gx0, float gy0, float finv_x, float finv_y,
                                                 unsigned int* __restrict__ s_bits, unsigned int zmask, int row_phase = 0, int row_stride = 1) {
    const int iy0 = max(0, min(PIB_FG - 1, (int)floorf((f.y0 - gy0) * finv_y)));
    const int iy1 = max(0, min(PIB_FG - 1, (int)floorf((f.y1 - gy0) * finv_y)));
    const float ch = 1.f / finv_y, cw = 1.f / finv_x;
    const float eps_y = f.pad + 2e-3f * ch, eps_x = f.pad + 2e-3f * cw;
    // corners relative to the centre: local (sx*hx, sy*hy) -> world (lx*c + ly*s, -lx*s + ly*c)
    const float hx = f.tx + f.pad, hy = f.ty + f.pad;
    float px[4], py[4];
    const float sxs[4] = {1.f, -1.f, -1.f, 1.f}, sys[4] = {1.f, 1.f, -1.f, -1.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float lx = sxs[k] * hx, ly = sys[k] * hy;
        px[k] = lx * f.c + ly * f.s;
        py[k] = -lx * f.s + ly * f.c;
    }
    for (int iy = iy0 + row_phase; iy <= iy1; iy += row_stride) {
        const float ylo = gy0 + (float)iy * ch - f.cy - eps_y, yhi = gy0 + (float)(iy + 1) * ch - f.cy + eps_y;
        float xmin = FLT_MAX, xmax = -FLT_MAX;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float ax = px[k], ay = py[k], bx = px[(k + 1) & 3], by = py[(k + 1) & 3];
            if ((ay < ylo && by < ylo) || (ay > yhi && by > yhi)) continue;
            float xa = ax, xb = bx;
            const float dy = by - ay;
            if (dy != 0.f) {
                const float inv = 1.f / dy;
                const float t0 = fminf(fmaxf((ylo - ay) * inv, 0.f), 1.f), t1 = fminf(fmaxf((yhi - ay) * inv, 0.f), 1.f);
                xa = ax + t0 * (bx - ax);
                xb = ax + t1 * (bx - ax);
            }
            xmin = fminf(xmin, fminf(xa, xb));
            xmax = fmaxf(xmax, fmaxf(xa, xb));
        }
        if (!(xmax >= xmin)) continue;
        const int ix0 = max(0, min(PIB_FG - 1, (int)floorf((f.cx + xmin - eps_x - gx0) * finv_x)));
        const int ix1 = max(0, min(PIB_FG - 1, (int)floorf((f.cx + xmax + eps_x - gx0) * finv_x)));
#if GLENET_PIB_ZSLABS
        for (int w = ix0 >> 1; w <= (ix1 >> 1); ++w) {      // two 16-bit cells per word: OR the box's z slabs into the row's span
            const unsigned int mask = (2 * w >= ix0 ? zmask : 0u) | (2 * w + 1 <= ix1 ? zmask << 16 : 0u);
            atomicOr(&s_bits[iy * (PIB_FG / 2) + w], mask);
        }
#else
        (void)zmask;
        for (int w = ix0 >> 5; w <= (ix1 >> 5); ++w) {
            const int lo = max(ix0 - (w << 5), 0), hi = min(ix1 - (w << 5), 31);
            const unsigned int mask = (hi >= 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
            atomicOr(&s_bits[iy * (PIB_FG / 32) + w], mask);
        }
#endif
    }
}

// z slabs a box can accept points in (16-bit mask), by the frame's slab mapping.  The predicate keeps a point when
// |fl(z - cz)| <= tz; the slack (1e-3 + 1e-5 of the extent + 1e-6 of the magnitude) is far above the rounding of that
// subtraction and of cz -+ tz.  tz < 0: the predicate rejects everything => no slab.
__device__ __forceinline__ unsigned int z_slab_mask(float cz, float tz, float zinv, float zoff) {
#if GLENET_PIB_ZSLABS
    if (tz < 0.f) return 0u;
    if (zinv == 0.f) return 1u;                                   // no z window in this frame: every point maps to slab 0
    const float sl = 1e-3f + 1e-5f * tz + 1e-6f * fabsf(cz);
    const int lo = max(0, min(PIB_ZS - 1, __float2int_rd(__fmaf_rn(cz - tz - sl, zinv, zoff))));
    const int hi = max(0, min(PIB_ZS - 1, __float2int_rd(__fmaf_rn(cz + tz + sl, zinv, zoff))));
    return ((2u << hi) - 1u) & ~((1u << lo) - 1u);
#else
    (void)cz; (void)zinv; (void)zoff;
    return tz < 0.f ? 0u : 1u;
#endif
}

template <int G, typename F>
__device__ __forceinline__ void for_cells(const Footprint& f, float gx0, float gy0, float inv_x, float inv_y, int row_phase, int row_stride, F visit) {
    const int ix0 = max(0, min(G - 1, (int)floorf((f.x0 - gx0) * inv_x)));
    const int ix1 = max(0, min(G - 1, (int)floorf((f.x1 - gx0) * inv_x)));
    const int iy0 = max(0, min(G - 1, (int)floorf((f.y0 - gy0) * inv_y)));
    const int iy1 = max(0, min(G - 1, (int)floorf((f.y1 - gy0) * inv_y)));
    const float cwx = 1.f / inv_x, cwy = 1.f / inv_y;
    const float rx_ext = 0.5f * (fabsf(f.c) * cwx + fabsf(f.s) * cwy), ry_ext = 0.5f * (fabsf(f.s) * cwx + fabsf(f.c) * cwy);
    const float slack = f.pad + 1e-3f * (cwx + cwy);
    const int nx = ix1 - ix0 + 1, ncell = nx * (iy1 - iy0 + 1);
    (void)ncell;
    for (int iy = iy0 + row_phase; iy <= iy1; iy += row_stride)
    for (int ix = ix0; ix <= ix1; ++ix) {
        const float mx = gx0 + ((float)ix + 0.5f) * cwx - f.cx, my = gy0 + ((float)iy + 0.5f) * cwy - f.cy;
        const float lx = mx * f.c - my * f.s, ly = mx * f.s + my * f.c;
        if (fabsf(lx) > f.tx + rx_ext + slack || fabsf(ly) > f.ty + ry_ext + slack) continue;
        visit(iy * G + ix);
    }
}

__global__ void __launch_bounds__(PIB_BUILD_THREADS)
pib_build_kernel(const float* __restrict__ boxes_all, int N, PibWorkspace ws) {
    extern __shared__ __align__(16) unsigned char pib_build_smem[];   // PIB_BUILD_SMEM bytes
    unsigned long long* s_inl = reinterpret_cast<unsigned long long*>(pib_build_smem);   // [PIB_CELLS] first four candidates, 16 bits each
    unsigned int* cnt = reinterpret_cast<unsigned int*>(s_inl + PIB_CELLS);              // [PIB_CELLS] count, then fill cursor
    unsigned int* s_start = cnt + PIB_CELLS;                                             // [PIB_CELLS] crowded cells: offset of their slice of list[]
    unsigned int* s_bits = s_start + PIB_CELLS;                                          // [PIB_FWORDS]
    __shared__ float red[6][PIB_BUILD_THREADS / 32];
    __shared__ float s_bounds[6];
    __shared__ int s_bad, s_zwide, s_over;
    __shared__ unsigned int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the query grid be scheduled behind this one
    const float* boxes = boxes_all + (size_t)f * N * 7;
    float* rec = ws.rec + (size_t)f * N * 8;
    unsigned int* list = ws.list + (size_t)f * ws.cap;
    unsigned int* bits = ws.bits + (size_t)f * PIB_FWORDS;

    for (int i = tid; i < PIB_CELLS; i += PIB_BUILD_THREADS) { cnt[i] = 0; s_inl[i] = ~0ull; }
    for (int i = tid; i < PIB_FWORDS; i += PIB_BUILD_THREADS) s_bits[i] = 0;
    if (tid == 0) { s_bad = 0; s_zwide = 0; s_over = 0; }
    // The build is launched as a programmatic dependent of whatever precedes it in the stream: the launch and the clearing of
    // 100 KB of shared memory overlap that kernel's tail.  Nothing global is touched before this wait (the previous call's
    // query may still be reading the workspace, a producer may still be writing the boxes).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();

    // pass 1: records + frame bounds
    float bx0 = FLT_MAX, by0 = FLT_MAX, bx1 = -FLT_MAX, by1 = -FLT_MAX;
    float zlo = FLT_MAX, zhi = -FLT_MAX;   // union of the boxes' z intervals [cz - tz, cz + tz]
    bool bad = false, zwide = false;
    for (int k = tid; k < N; k += PIB_BUILD_THREADS) {
        const float* b = boxes + (size_t)k * 7;
        const float rz = b[6];
        float tx, ty, tz;
        box_thresholds<false>(b[3], b[4], b[5], tx, ty, tz);
        float* r = rec + (size_t)k * 8;
        float sn, cs;
        sincosf(-rz, &sn, &cs);   // == cosf(-rz), sinf(-rz) bit for bit (tools/cuda/trig_symmetry.cu)
        r[0] = b[0]; r[1] = b[1]; r[2] = b[2]; r[3] = cs; r[4] = sn; r[5] = tx; r[6] = ty; r[7] = tz;
        const Footprint fp = footprint(r);
        if (fp.never) continue;
        if (fp.bad) { bad = true; continue; }
        // `|z - cz| > tz` rejects: a NaN on either side never rejects (no window then), tz < 0 always does
        if (!(fabsf(b[2]) <= FLT_MAX) || !(tz <= FLT_MAX)) zwide = true;
        else if (tz >= 0.f) { zlo = fminf(zlo, b[2] - tz); zhi = fmaxf(zhi, b[2] + tz); }
        bx0 = fminf(bx0, fp.x0); bx1 = fmaxf(bx1, fp.x1); by0 = fminf(by0, fp.y0); by1 = fmaxf(by1, fp.y1);
    }
    if (bad) s_bad = 1;
    if (zwide) s_zwide = 1;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        zlo = fminf(zlo, __shfl_xor_sync(0xffffffffu, zlo, o));
        zhi = fmaxf(zhi, __shfl_xor_sync(0xffffffffu, zhi, o));
        bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
        by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o));
        bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
        by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
    }
    if (lane == 0) { red[0][warp] = bx0; red[1][warp] = by0; red[2][warp] = bx1; red[3][warp] = by1; red[4][warp] = zlo; red[5][warp] = zhi; }
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = PIB_BUILD_THREADS / 32;
        bx0 = lane < NW ? red[0][lane] : FLT_MAX;  by0 = lane < NW ? red[1][lane] : FLT_MAX;
        bx1 = lane < NW ? red[2][lane] : -FLT_MAX; by1 = lane < NW ? red[3][lane] : -FLT_MAX;
        zlo = lane < NW ? red[4][lane] : FLT_MAX;  zhi = lane < NW ? red[5][lane] : -FLT_MAX;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o));
            bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o)); by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
            zlo = fminf(zlo, __shfl_xor_sync(0xffffffffu, zlo, o)); zhi = fmaxf(zhi, __shfl_xor_sync(0xffffffffu, zhi, o));
        }
    }
    if (tid == 0) {
        s_bounds[0] = bx0; s_bounds[1] = by0; s_bounds[2] = bx1; s_bounds[3] = by1;
        // conservative window: the slack (1e-3 + 1e-5 of the width + 1e-6 of the magnitude) is far above the rounding
        // of `z - cz`, `cz -+ tz` and `z - zc`; anything non-finite => no window
        float zc = 0.5f * zlo + 0.5f * zhi;
        float zh = (0.5f * zhi - 0.5f * zlo) * 1.00001f + 1e-3f + 1e-6f * fmaxf(fabsf(zlo), fabsf(zhi));
        if (s_zwide || !(zhi >= zlo) || !(zh <= FLT_MAX) || !(fabsf(zc) <= FLT_MAX)) { zc = 0.f; zh = __int_as_float(0x7f800000); }
#if GLENET_PIB_ZSLABS
        // slab = floor(fma(z, zinv, zoff)): PIB_ZS slabs over [zc - zh, zc + zh]; the SAME monotone mapping gives every box its
        // slab range (z_slab_mask), so a point the predicate can accept always finds its slab set.  No window: all in slab 0.
        float zinv = ((float)PIB_ZS - 0.01f) / (2.f * zh), zoff = -(zc - zh) * zinv;
        if (!(zh <= FLT_MAX) || !(zinv <= FLT_MAX) || !(fabsf(zoff) <= FLT_MAX)) { zinv = 0.f; zoff = 0.f; }
        zc = zinv; zh = zoff;
#endif
        s_bounds[4] = zc; s_bounds[5] = zh;
    }
    __syncthreads();   // also publishes the records written above to the whole CTA
    bx0 = s_bounds[0]; by0 = s_bounds[1]; bx1 = s_bounds[2]; by1 = s_bounds[3];
    const bool empty = !(bx1 >= bx0);   // no box can contain anything
    float inv_x = 0.f, inv_y = 0.f, finv_x = 0.f, finv_y = 0.f;
    bool exhaustive = s_bad != 0;
    if (!empty) {
        const float ex = bx1 - bx0, ey = by1 - by0;
        inv_x = ((float)PIB_G - 0.01f) / ex;   inv_y = ((float)PIB_G - 0.01f) / ey;
        finv_x = ((float)PIB_FG - 0.01f) / ex; finv_y = ((float)PIB_FG - 0.01f) / ey;
        if (!(inv_x > 0.f) || !(inv_y > 0.f) || !(finv_x <= FLT_MAX) || !(finv_y <= FLT_MAX)) exhaustive = true;
    }
    if (N > 65534) exhaustive = true;   // 0xffff = no candidate, 0xfffe = more than four

    unsigned int total = 0;
    if (!exhaustive && !empty) {
        // One box per thread (N <= 512 boxes is one round): with a few hundred boxes per frame, box-level
        // parallelism beats splitting one box's ~15 rows / ~15 cells over a warp (measured both ways).
        //   pass 0: fine map + coarse counts / inline candidates      pass 1 (only with crowded cells): their list slices
        // With few boxes the CTA has threads to spare: `parts` threads share one box and take interleaved rows of both
        // grids -- laid out so that a warp holds one part only.
        const int parts = N * 4 <= PIB_BUILD_THREADS ? 4 : (N * 2 <= PIB_BUILD_THREADS ? 2 : 1);
        for (int pass = 0; pass < 2; ++pass) {
            const int items = N * parts;
            for (int it = tid; it < items; it += PIB_BUILD_THREADS) {
                const int part = it / N, k = it - part * N;
                const Footprint fp = footprint(rec + (size_t)k * 8);
                if (fp.never) continue;
                // the `parts` threads of a box take interleaved rows of BOTH grids, in both passes
                const int row_phase = part, row_stride = parts;
                if (!(GLENET_PIB_DBG & 1) && pass == 0) {
                    const unsigned int zmask = z_slab_mask(rec[(size_t)k * 8 + 2], rec[(size_t)k * 8 + 7], s_bounds[4], s_bounds[5]);
                    if (zmask) raster_fine_rows(fp, bx0, by0, finv_x, finv_y, s_bits, zmask, row_phase, row_stride);
                }
                if (GLENET_PIB_DBG & 2) continue;
                for_cells<PIB_G>(fp, bx0, by0, inv_x, inv_y, row_phase, row_stride, [&](int cell) {
                    if (pass == 0) {     // the first four candidates of a cell go inline; a fifth makes it a crowded cell
                        const unsigned int pos = atomicAdd(&cnt[cell], 1u);
                        if (pos < 4u) reinterpret_cast<unsigned short*>(s_inl)[cell * 4 + pos] = (unsigned short)k;
                        else s_over = 1;
                    } else if (cnt[cell] > 4u) {
                        const unsigned int pos = atomicAdd(&reinterpret_cast<unsigned int*>(s_inl)[2 * cell], 1u);
                        list[s_start[cell] + pos] = (unsigned int)k;
                    }
                });
            }
            __syncthreads();
            if (pass == 1 || !s_over || (GLENET_PIB_DBG & 4)) break;   // uniform: usually every cell is complete inline -- no lists, no second pass
            // Crowded cells (a fifth candidate arrived): each gets a contiguous slice of list[] -- reserved in any order with one
            // atomicAdd -- and the second pass re-visits the boxes' cells and fills only those slices.  The cell word then holds
            // (slice offset, count, 0xfffe) instead of four inline ids, so the query needs no separate index.
            if (tid == 0) s_total = 0;
            __syncthreads();
            for (int c = tid; c < PIB_CELLS; c += PIB_BUILD_THREADS) {
                if (cnt[c] > 4u) {
                    s_start[c] = atomicAdd(&s_total, cnt[c]);
                    reinterpret_cast<unsigned int*>(s_inl)[2 * c] = 0u;      // fill cursor (the inline ids are no longer needed)
                }
            }
            __syncthreads();
            total = s_total;
            if (total > ws.cap) { exhaustive = true; break; }   // uniform
        }
        for (int i = tid; i < PIB_FWORDS; i += PIB_BUILD_THREADS) bits[i] = s_bits[i];
        if (!exhaustive && !(GLENET_PIB_DBG & 8)) {
            // packed coarse cells: up to four candidates inline, 16 bits each (0xffff = none); a crowded cell holds
            // {offset into list[] : 32, count : 16, 0xfffe : 16}
            unsigned long long* cells = ws.cells + (size_t)f * PIB_CELLS;
            for (int c = tid; c < PIB_CELLS; c += PIB_BUILD_THREADS) {
                unsigned long long v = s_inl[c];
                if (s_over && cnt[c] > 4u) v = (unsigned long long)s_start[c] | ((unsigned long long)cnt[c] << 32) | (0xfffeull << 48);
                cells[c] = v;
            }
        }
    }
    if (tid == 0) {
        PibFrame h;
        h.gx0 = bx0; h.gy0 = by0; h.inv_x = inv_x; h.inv_y = inv_y;
        h.exhaustive = exhaustive ? 1 : 0;
        h.list_len = empty ? -1 : (int)total;   // -1: nothing can match in this frame
        h.finv_x = finv_x; h.finv_y = finv_y;
        h.zc = s_bounds[4]; h.zh = s_bounds[5]; h.foff_x = -bx0 * finv_x; h.foff_y = -by0 * finv_y;
        ws.frames[f] = h;
    }
}

// 4 consecutive points of one thread (three float4 loads when the frame's rows are 16-byte aligned)
struct Pts4 { float x[4], y[4], z[4]; };
__device__ __forceinline__ void load_pts4(const float* __restrict__ pts, int p0, int p_end, bool vec, Pts4& o) {
    const int nvalid = p_end - p0;
    if (vec && nvalid >= 4) {
        const float4* src = reinterpret_cast<const float4*>(pts + (size_t)p0 * 3);
        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        o.x[0] = a.x; o.y[0] = a.y; o.z[0] = a.z; o.x[1] = a.w; o.y[1] = b.x; o.z[1] = b.y;
        o.x[2] = b.z; o.y[2] = b.w; o.z[2] = c.x; o.x[3] = c.y; o.y[3] = c.z; o.z[3] = c.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < nvalid) {
                o.x[i] = __ldg(pts + (size_t)(p0 + i) * 3); o.y[i] = __ldg(pts + (size_t)(p0 + i) * 3 + 1); o.z[i] = __ldg(pts + (size_t)(p0 + i) * 3 + 2);
            } else { o.x[i] = o.y[i] = o.z[i] = __int_as_float(0x7fc00000); }   // NaN => never inside
        }
    }
}

// Query kernel.  Persistent CTAs (GLENET_PIB_CTAS per SM) walk a contiguous range of (frame, chunk) work
// items -- the chunks of one frame as ONE run of points -- and re-stage the frame tables (fine map 32 KB,
// coarse cells 32 KB, box records) only when the frame changes.  Inside a run every WARP is autonomous -- no
// CTA barrier on the streaming path:
//   1. 4 consecutive points per lane (three float4 loads, one batch prefetched into registers and the
//      lines of the batch GLENET_PIB_L2PF further ahead prefetched into L2);
//   2. fine-map lookup in shared memory (cell -> 16 z-slab bits -> the point's slab), provisional -1 for all
//      four points with one int4 store;
//   3. the ~10 % of points that stay hot go to the warp's private queue (ballot prefix, no atomics), and the
//      warp drains it 32 entries at a time: coarse cell -> up to four inline candidates (a list slice through
//      L2 only for crowded cells) -> exact predicate -> minimum index over the provisional -1.
// The LSU data pipe (shared-memory wavefronts) is the busiest unit of this kernel, so the candidate records are
// read as two LDS.128 from a 48-byte stride (a generic pointer here once made them eight scalar loads with
// 8-way bank conflicts: 40 % of all wavefronts).
template <bool REC_SMEM>   // box records staged in shared memory (N <= PIB_SMEM_BOXES) or read through L1 / L2
__global__ void __launch_bounds__(PIB_THREADS, GLENET_PIB_CTAS)
pib_query_kernel(const float* __restrict__ pts_all, int N, int M, PibWorkspace ws, int* __restrict__ out_all,
                 int chunks_per_frame, int total_chunks) {
    extern __shared__ __align__(16) unsigned char pib_smem[];
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the build kernel (previous in the stream) has completed and flushed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a programmatic dependent (e.g. the next call's build) may be scheduled as SMs free up
    float4* s_q = reinterpret_cast<float4*>(pib_smem);                                        // [warps][PIB_WQ] {x, y, z, index}
    unsigned long long* s_cells = reinterpret_cast<unsigned long long*>(s_q + (PIB_THREADS / 32) * PIB_WQ); // [PIB_CELLS] four ids
    unsigned int* s_bits = reinterpret_cast<unsigned int*>(s_cells + PIB_CELLS);              // [PIB_FWORDS]
    float* s_rec = reinterpret_cast<float*>(s_bits + PIB_FWORDS);                             // [N * PIB_REC_STRIDE] (REC_SMEM)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4* wq = s_q + warp * PIB_WQ;
    const int c_begin = (int)((long)blockIdx.x * total_chunks / gridDim.x);
    const int c_end = (int)((long)(blockIdx.x + 1) * total_chunks / gridDim.x);
    constexpr int RS = REC_SMEM ? PIB_REC_STRIDE : 8;       // floats from one record to the next
    int cur_frame = -1;
    PibFrame h;
    h.list_len = -1; h.exhaustive = 0; h.gx0 = h.gy0 = h.inv_x = h.inv_y = h.finv_x = h.finv_y = h.zc = h.zh = h.foff_x = h.foff_y = 0.f;

    for (int c = c_begin; c < c_end;) {
        const int f = c / chunks_per_frame;
        const int chunk = c - f * chunks_per_frame;
#if GLENET_PIB_RUNS
        // this CTA's chunks of one frame form ONE contiguous run of points: the register prefetch streams through
        // it without a pipeline refill (and an integer division) at every chunk boundary
        const int chunk_last = min(c_end - f * chunks_per_frame, chunks_per_frame);
#else
        const int chunk_last = chunk + 1;
#endif
        c = f * chunks_per_frame + chunk_last;
        const float* pts = pts_all + (size_t)f * M * 3;
        int* out = out_all + (size_t)f * M;
        const float* rec_g = ws.rec + (size_t)f * N * 8;
        const float* rec = REC_SMEM ? s_rec : rec_g;          // resolved at compile time: LDS or LDG, never a generic load
        const int p_begin = chunk * PIB_CHUNK;
        const int p_end = (int)min((long)M, (long)chunk_last * PIB_CHUNK);
        const bool vec = ((((uintptr_t)pts) & 15) == 0) && ((((uintptr_t)out) & 15) == 0);
        if (f != cur_frame) {                                 // uniform over the CTA
            __syncthreads();                                  // everyone is done with the previous frame's tables
            h = ws.frames[f];
            if (h.list_len >= 0) {
                if (REC_SMEM) {
                    const float4* src = reinterpret_cast<const float4*>(rec_g);
                    float4* dst = reinterpret_cast<float4*>(s_rec);
                    for (int i = tid; i < N * 2; i += PIB_THREADS) dst[(i >> 1) * (PIB_REC_STRIDE / 4) + (i & 1)] = __ldg(src + i);
                }
                if (!h.exhaustive) {
                    const uint4* b = reinterpret_cast<const uint4*>(ws.bits + (size_t)f * PIB_FWORDS);
                    for (int i = tid; i < PIB_FWORDS / 4; i += PIB_THREADS) reinterpret_cast<uint4*>(s_bits)[i] = __ldg(b + i);
                    const uint4* cl = reinterpret_cast<const uint4*>(ws.cells + (size_t)f * PIB_CELLS);
                    for (int i = tid; i < PIB_CELLS / 2; i += PIB_THREADS) reinterpret_cast<uint4*>(s_cells)[i] = __ldg(cl + i);
                }
            }
            __syncthreads();
            cur_frame = f;
        }
        if (h.list_len < 0) {   // no box of this frame can contain a point
            for (int p = p_begin + tid; p < p_end; p += PIB_THREADS) out[p] = -1;
            continue;
        }
        if (h.exhaustive) {
            for (int p = p_begin + tid; p < p_end; p += PIB_THREADS) {
                const float x = pts[(size_t)p * 3], y = pts[(size_t)p * 3 + 1], z = pts[(size_t)p * 3 + 2];
                int res = -1;
                for (int k = 0; k < N; ++k) {
                    if (pt_in_box_gpu(x, y, z, rec + (size_t)k * RS)) { res = k; break; }
                }
                out[p] = res;
            }
            continue;
        }

        // ---- warp-autonomous streaming over this chunk: batch j of warp w covers 128 points
        const int NB = (p_end - p_begin + PIB_THREADS * 4 - 1) / (PIB_THREADS * 4);   // batches per warp in this run
        // candidate test of one queued point: coarse cell (two inline candidates in shared memory, longer
        // lists through global memory) -> exact predicate -> minimum index
        auto resolve = [&](const float4 e) {
            const int cx = (int)((e.x - h.gx0) * h.inv_x), cy = (int)((e.y - h.gy0) * h.inv_y);
            const int ci = min(cy, PIB_G - 1) * PIB_G + min(cx, PIB_G - 1);
            const unsigned long long cell = s_cells[ci];
            const unsigned int lo = (unsigned int)cell, hi = (unsigned int)(cell >> 32);
            const int k0 = (int)(lo & 0xffffu), k1 = (int)(lo >> 16), k2 = (int)(hi & 0xffffu), k3 = (int)(hi >> 16);
            int res = 0x7fffffff;
            if (k3 == 0xfffe) {          // crowded cell (rare): {offset, count} of its slice of the frame's list; pointer rebuilt here
                const unsigned int s0 = lo, n = hi & 0xffffu;
                const unsigned int* list = ws.list + (size_t)f * ws.cap;
                for (unsigned int i = 0; i < n; i += 4) {   // four independent list loads per round trip to L2
                    int kk[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) kk[u] = (i + u < n) ? (int)__ldg(list + s0 + i + u) : 0x7fffffff;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kk[u] < res && pt_in_box_gpu(e.x, e.y, e.z, rec + (size_t)kk[u] * RS)) res = kk[u];
                }
            } else {                     // candidates are in no particular order: keep the minimum
                if (k0 != 0xffff && pt_in_box_gpu(e.x, e.y, e.z, rec + (size_t)k0 * RS)) res = k0;
                if (k1 != 0xffff && k1 < res && pt_in_box_gpu(e.x, e.y, e.z, rec + (size_t)k1 * RS)) res = k1;
                if (k2 != 0xffff && k2 < res && pt_in_box_gpu(e.x, e.y, e.z, rec + (size_t)k2 * RS)) res = k2;
                if (k3 != 0xffff && k3 < res && pt_in_box_gpu(e.x, e.y, e.z, rec + (size_t)k3 * RS)) res = k3;
            }
            if (res != 0x7fffffff) out[__float_as_int(e.w)] = res;
        };
        int qn = 0;                                           // warp-uniform fill of the warp's queue
        // one batch (4 points per lane) of this warp: lookup, provisional -1, queue append, drain
        auto batch = [&](const Pts4& cur, const int j) {
            const int p0 = p_begin + (j * (PIB_THREADS / 32) + warp) * PIB_WBATCH + lane * 4;
            const int nvalid = max(0, min(4, p_end - p0));
            // branch-free fine-bitmap lookup: one FMA per axis -> floor -> one unsigned range test for both axes -> one LDS.
            // (The build rasterises with floor((x - gx0) * finv); the fused form differs from it by < 6e-5 cell, the
            // footprints are padded by > 2e-3 cell.)  NaN maps to cell 0 (harmless: the exact predicate rejects it),
            // out-of-grid to "cold".
            unsigned int hot = 0;
            const float finv_x = h.finv_x, finv_y = h.finv_y, foff_x = h.foff_x, foff_y = h.foff_y, zc = h.zc, zh = h.zh;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ix = __float2int_rd(__fmaf_rn(cur.x[i], finv_x, foff_x)), iy = __float2int_rd(__fmaf_rn(cur.y[i], finv_y, foff_y));
#if GLENET_PIB_ZSLABS
                // cell = 16 z-slab bits; slabs outside [0, 16) shift the bits out (PTX shr clamps the amount: negative slab
                // numbers are huge unsigned amounts).  NaN z -> slab 0; corrected below.
                const int zs = __float2int_rd(__fmaf_rn(cur.z[i], zc, zh));
                const unsigned int cell = reinterpret_cast<const unsigned short*>(s_bits)[((iy << 7) + ix) & (PIB_FG * PIB_FG - 1)];
                unsigned int bit;
                asm("shr.u32 %0, %1, %2;" : "=r"(bit) : "r"(cell), "r"(zs));
                hot |= ((unsigned int)(ix | iy) < (unsigned int)PIB_FG) ? ((bit & 1u) << i) : 0u;
#else
                const unsigned int word = s_bits[((iy << 3) + (ix >> 5)) & (PIB_FWORDS - 1)];
#if GLENET_PIB_ZWINDOW
                // also cold: points above / below every box (NaN z is not culled here; the exact predicate rejects it)
                const bool in_grid = ((unsigned int)(ix | iy) < (unsigned int)PIB_FG) & !(fabsf(cur.z[i] - zc) > zh);
#else
                const bool in_grid = (unsigned int)(ix | iy) < (unsigned int)PIB_FG;
#endif
                hot |= in_grid ? (((word >> (ix & 31)) & 1u) << i) : 0u;
#endif
            }
#if GLENET_PIB_ZSLABS
            // A NaN z passes the reference's z test (`|z - cz| > dz/2` is false), so such a point belongs to the first box whose
            // footprint holds it, whatever the slab: one sum test per lane, the fix-up itself practically never runs.
            const float zsum = (cur.z[0] + cur.z[1]) + (cur.z[2] + cur.z[3]);
            if (zsum != zsum) {
#pragma unroll 1
                for (int i = 0; i < 4; ++i) {
                    const float zi = i == 0 ? cur.z[0] : i == 1 ? cur.z[1] : i == 2 ? cur.z[2] : cur.z[3];
                    if (zi == zi) continue;
                    const float xi = i == 0 ? cur.x[0] : i == 1 ? cur.x[1] : i == 2 ? cur.x[2] : cur.x[3];
                    const float yi = i == 0 ? cur.y[0] : i == 1 ? cur.y[1] : i == 2 ? cur.y[2] : cur.y[3];
                    const int ix = __float2int_rd(__fmaf_rn(xi, finv_x, foff_x)), iy = __float2int_rd(__fmaf_rn(yi, finv_y, foff_y));
                    if ((unsigned int)(ix | iy) >= (unsigned int)PIB_FG) continue;
                    const unsigned int any = reinterpret_cast<const unsigned short*>(s_bits)[(iy << 7) + ix];
                    if (any) hot |= 1u << i;
                }
            }
#endif
            hot &= (1u << nvalid) - 1u;                        // never queue a point beyond the chunk
            if (vec && nvalid == 4) *reinterpret_cast<int4*>(out + p0) = make_int4(-1, -1, -1, -1);
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (i < nvalid) out[p0 + i] = -1;
            }
            if (__any_sync(0xffffffffu, hot != 0)) {
                // append the hot points: one ballot per point slot gives every lane its queue position (the order of the
                // queue is irrelevant) -- cheaper than a 5-step prefix sum followed by per-lane sequential writes
                const unsigned int lt = (1u << lane) - 1u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool hi = (hot >> i) & 1u;
                    const unsigned int m = __ballot_sync(0xffffffffu, hi);
                    if (hi) wq[qn + __popc(m & lt)] = make_float4(cur.x[i], cur.y[i], cur.z[i], __int_as_float(p0 + i));
                    qn += __popc(m);
                }
                __syncwarp();   // queue visible; also orders the provisional -1 stores before the results
                // drain whole warps only; the remainder (< 32 entries) waits for the next batch
                int head = 0;
                for (; head + 32 <= qn; head += 32) resolve(wq[head + lane]);
                if (head) {
                    const int rest = qn - head;
                    float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lane < rest) keep = wq[head + lane];
                    __syncwarp();
                    if (lane < rest) wq[lane] = keep;
                    qn = rest;
                    __syncwarp();
                }
            }
        };
        auto fetch = [&](Pts4& dst, const int j) {
            load_pts4(pts, p_begin + (j * (PIB_THREADS / 32) + warp) * PIB_WBATCH + lane * 4, p_end, vec, dst);
        };
        // register prefetch GLENET_PIB_PF batches ahead (2: one batch of cold points is shorter than a DRAM round trip)
        Pts4 cur, nxt;
        fetch(cur, 0);
#if GLENET_PIB_PF >= 2
        if (NB > 1) fetch(nxt, 1);
#endif
#if GLENET_PIB_L2PF
        // this lane's line of the batch GLENET_PIB_L2PF ahead; a batch is 1536 bytes = at most 13 lines, and the stride
        // from batch to batch (12288 bytes) keeps the alignment, so the pointer simply advances
        const char* pf_line;
        {
            const char* b0 = reinterpret_cast<const char*>(pts + (size_t)(p_begin + (GLENET_PIB_L2PF * (PIB_THREADS / 32) + warp) * PIB_WBATCH) * 3);
            pf_line = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(b0) & ~(uintptr_t)127) + lane * 128;
            if (!(pf_line < b0 + PIB_WBATCH * 12)) pf_line = reinterpret_cast<const char*>(~(uintptr_t)0);   // lane has no line
        }
        const char* const pf_end = reinterpret_cast<const char*>(pts + (size_t)p_end * 3);
#endif
#pragma unroll 1
        for (int j = 0; j < NB; ++j) {
#if GLENET_PIB_PF >= 2
            Pts4 nxt2;
            if (j + 2 < NB) fetch(nxt2, j + 2);
#else
            if (j + 1 < NB) fetch(nxt, j + 1);
#endif
#if GLENET_PIB_L2PF
            // The register rotation below (cur = nxt; nxt = nxt2) has to wait for nxt2, so the register prefetch is
            // effectively ONE batch deep.  An L2 prefetch further ahead turns that wait into an L2 hit.
            if (pf_line < pf_end) asm volatile("prefetch.global.L2 [%0];" :: "l"(pf_line));
            pf_line += (pf_line < pf_end) ? (size_t)PIB_THREADS * 4 * 12 : 0;
#endif
            batch(cur, j);
#if GLENET_PIB_PF >= 2
            cur = nxt;
            nxt = nxt2;
#else
            cur = nxt;
#endif
        }
        if (qn) {                                              // leftovers of the run
            if (lane < qn) resolve(wq[lane]);
            __syncwarp();
        }
    }
}

// Small problems (N <= PIB_DIRECT_BOXES, few points): one launch, no grid.  Every CTA rebuilds the
// N box records in shared memory and runs the reference's ascending loop with early exit.
__global__ void __launch_bounds__(PIB_ROUND / 4)
pib_direct_kernel(const float* __restrict__ boxes_all, const float* __restrict__ pts_all, int N, int M,
                  int* __restrict__ out_all, int chunks_per_frame) {
    __shared__ __align__(16) float s_rec[PIB_DIRECT_BOXES * 8];
    const int tid = threadIdx.x;
    const int f = blockIdx.x / chunks_per_frame;
    const int chunk = blockIdx.x - f * chunks_per_frame;
    const float* pts = pts_all + (size_t)f * M * 3;
    int* out = out_all + (size_t)f * M;
    if (tid < N) {
        const float* b = boxes_all + ((size_t)f * N + tid) * 7;
        float tx, ty, tz;
        box_thresholds<false>(b[3], b[4], b[5], tx, ty, tz);
        float* r = s_rec + tid * 8;
        float sn, cs;
        sincosf(-b[6], &sn, &cs);
        r[0] = b[0]; r[1] = b[1]; r[2] = b[2]; r[3] = cs; r[4] = sn; r[5] = tx; r[6] = ty; r[7] = tz;
    }
    const int p_begin = chunk * PIB_DIRECT_PTS, p_end = min(M, p_begin + PIB_DIRECT_PTS);
    const bool vec = ((((uintptr_t)pts) & 15) == 0) && ((((uintptr_t)out) & 15) == 0);
    Pts4 cur;
    load_pts4(pts, p_begin + tid * 4, p_end, vec, cur);
    __syncthreads();
    for (int base = p_begin; base < p_end; base += PIB_ROUND) {
        const int p0 = base + tid * 4;
        const int nvalid = max(0, min(4, p_end - p0));
        Pts4 nxt;
        load_pts4(pts, p0 + PIB_ROUND, p_end, vec, nxt);
        int res[4] = {-1, -1, -1, -1};
        for (int k = N - 1; k >= 0; --k) {      // descending + overwrite == first hit of the ascending loop
            const float* r = s_rec + k * 8;
#pragma unroll
            for (int i = 0; i < 4; ++i) if (pt_in_box_gpu(cur.x[i], cur.y[i], cur.z[i], r)) res[i] = k;
        }
        if (vec && nvalid == 4) *reinterpret_cast<int4*>(out + p0) = make_int4(res[0], res[1], res[2], res[3]);
        else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (i < nvalid) out[p0 + i] = res[i];
        }
        cur = nxt;
    }
}

// points_in_boxes_cpu semantics (roiaware_pool3d.cpp:128-165): out[i][j] = point j in box i, MARGIN 1e-2,
// host libm cos/sin, no FMA:  lx = sx*c + sy*(-s) ; ly = sx*s + sy*c  (each product rounded).
__global__ void __launch_bounds__(256)
pib_mask_cpu_dialect_kernel(const float* __restrict__ boxes, const float* __restrict__ trig, int N,
                            const float* __restrict__ pts, int M, int* __restrict__ out) {
    __shared__ float s_rec[8];
    const int i = blockIdx.y;
    if (threadIdx.x == 0) {
        const float* b = boxes + (size_t)i * 7;
        float tx, ty, tz;
        box_thresholds<true>(b[3], b[4], b[5], tx, ty, tz);
        s_rec[0] = b[0]; s_rec[1] = b[1]; s_rec[2] = b[2];
        s_rec[3] = trig[2 * i]; s_rec[4] = trig[2 * i + 1];
        s_rec[5] = tx; s_rec[6] = ty; s_rec[7] = tz;
    }
    __syncthreads();
    const float cx = s_rec[0], cy = s_rec[1], cz = s_rec[2], c = s_rec[3], s = s_rec[4];
    const float tx = s_rec[5], ty = s_rec[6], tz = s_rec[7];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const float x = pts[(size_t)j * 3], y = pts[(size_t)j * 3 + 1], z = pts[(size_t)j * 3 + 2];
        int in = 0;
        if (!(fabsf(__fsub_rn(z, cz)) > tz)) {
            const float sx = __fsub_rn(x, cx), sy = __fsub_rn(y, cy);
            const float lx = __fsub_rn(__fmul_rn(sx, c), __fmul_rn(sy, s));
            const float ly = __fadd_rn(__fmul_rn(sx, s), __fmul_rn(sy, c));
            in = (tx > fabsf(lx)) & (ty > fabsf(ly));
        }
        out[(size_t)i * M + j] = in;
    }
}

}  // namespace glenet

using namespace glenet;

extern "C" {

size_t glenet_points_in_boxes_workspace_bytes(int batch, int boxes_num) {
    if (batch <= 0 || boxes_num <= 0) return 16;
    return pib_layout(nullptr, batch, boxes_num).bytes;
}

int glenet_points_in_boxes_gpu(const float* boxes, const float* pts, int B, int N, int M, int32_t* out, void* ws,
                               size_t ws_bytes, glenet_stream_t s) {
    const char* what = "glenet_points_in_boxes_gpu";
    cudaStream_t st = (cudaStream_t)s;
    if (B < 0 || N < 0 || M < 0) return fail(GLENET_EINVAL, "%s: negative size", what);
    if (B == 0 || M == 0) return GLENET_OK;
    if (!pts || !out) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (N == 0) {   // the reference's wrapper pre-fills -1 and the kernel loop never runs
        cudaError_t e = cudaMemsetAsync(out, 0xff, sizeof(int32_t) * (size_t)B * M, st);
        return e == cudaSuccess ? GLENET_OK : fail(-(int)e, "%s: memset failed", what);
    }
    if (!boxes) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (!ws || ws_bytes < glenet_points_in_boxes_workspace_bytes(B, N)) return fail(GLENET_EWORKSPACE, "%s: workspace too small", what);
    if ((uintptr_t)ws & 15) return fail(GLENET_EALIGN, "%s: workspace must be 16-byte aligned", what);
    if (N <= PIB_DIRECT_BOXES && (long)B * M <= PIB_DIRECT_MAX_POINTS) {
        const int chunks = (M + PIB_DIRECT_PTS - 1) / PIB_DIRECT_PTS;
        pib_direct_kernel<<<(unsigned)(chunks * B), PIB_ROUND / 4, 0, st>>>(boxes, pts, N, M, out, chunks);
        return check_launch(what);
    }
    PibWorkspace w = pib_layout(ws, B, N);
    // the opt-in shared-memory sizes are per device: remember where they have been set
    static std::atomic<bool> attr_done[GLENET_MAX_DEVICES];
    int dev = 0;
    cudaGetDevice(&dev);
    const bool need_attr = dev < 0 || dev >= GLENET_MAX_DEVICES || !attr_done[dev].load(std::memory_order_acquire);
    int rc = GLENET_OK;
    if (need_attr) {
        rc = set_smem(pib_build_kernel, PIB_BUILD_SMEM, what);
        if (rc) return rc;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    {
        cudaLaunchConfig_t bcfg = {};
        bcfg.gridDim = dim3((unsigned)B); bcfg.blockDim = dim3(PIB_BUILD_THREADS); bcfg.dynamicSmemBytes = PIB_BUILD_SMEM; bcfg.stream = st;
        bcfg.attrs = attr; bcfg.numAttrs = 1;
        cudaError_t be = cudaLaunchKernelEx(&bcfg, pib_build_kernel, boxes, N, w);
        if (be != cudaSuccess) {
            snprintf(last_error_buf(), 512, "%s: launch failed: %s", what, cudaGetErrorString(be));
            return -(int)be;
        }
    }
    rc = check_launch(what);
    if (rc) return rc;
    if (GLENET_PIB_DBG & 16) return GLENET_OK;
    const int chunks = (M + PIB_CHUNK - 1) / PIB_CHUNK;
    const long total = (long)chunks * B;
    if (total > 0x7fffffffL) return fail(GLENET_EINVAL, "%s: too many chunks", what);
    const size_t smem_fixed = sizeof(float4) * (PIB_THREADS / 32) * PIB_WQ + sizeof(unsigned long long) * PIB_CELLS + sizeof(unsigned int) * PIB_FWORDS;
    const bool rec_smem = N <= PIB_SMEM_BOXES;
    const size_t smem = smem_fixed + sizeof(float) * PIB_REC_STRIDE * (size_t)(rec_smem ? N : 0);
    if (need_attr) {
        rc = set_smem(pib_query_kernel<true>, smem_fixed + sizeof(float) * PIB_REC_STRIDE * PIB_SMEM_BOXES, what);
        if (rc) return rc;
        rc = set_smem(pib_query_kernel<false>, smem_fixed, what);
        if (rc) return rc;
        if (dev >= 0 && dev < GLENET_MAX_DEVICES) attr_done[dev].store(true, std::memory_order_release);
    }
    const long resident = (long)GLENET_PIB_CTAS * 148;   // persistent: GLENET_PIB_CTAS CTAs per SM
    const unsigned grid = (unsigned)(total < resident ? total : resident);
    // programmatic dependent launch: the query grid is scheduled while the build kernel drains and waits
    // (griddepcontrol.wait) before it touches the workspace -- hides the launch gap between the two kernels
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PIB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int total_i = (int)total;
    cudaError_t le = rec_smem ? cudaLaunchKernelEx(&cfg, pib_query_kernel<true>, pts, N, M, w, out, chunks, total_i)
                              : cudaLaunchKernelEx(&cfg, pib_query_kernel<false>, pts, N, M, w, out, chunks, total_i);
    if (le != cudaSuccess) {
        snprintf(last_error_buf(), 512, "%s: launch failed: %s", what, cudaGetErrorString(le));
        return -(int)le;
    }
    return check_launch(what);
}

int glenet_points_in_boxes_cpu_dialect(const float* boxes, const float* trig, int N, const float* pts, int M,
                                       int32_t* out, glenet_stream_t s) {
    const char* what = "glenet_points_in_boxes_cpu_dialect";
    if (N < 0 || M < 0) return fail(GLENET_EINVAL, "%s: negative size", what);
    if (N == 0 || M == 0) return GLENET_OK;
    if (!boxes || !trig || !pts || !out) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (N > 65535) return fail(GLENET_EINVAL, "%s: more than 65535 boxes", what);
    const int bx = (M + 255) / 256;
    dim3 grid((unsigned)(bx < 1024 ? bx : 1024), (unsigned)N);
    pib_mask_cpu_dialect_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(boxes, trig, N, pts, M, out);
    return check_launch(what);
}

int glenet_abi_version(void) { return 13; }
const char* glenet_last_error(void) { return last_error_buf(); }

}  // extern "C"
