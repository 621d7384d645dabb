// Pairwise rotated-box overlap / BEV IoU / fused 3D IoU for sm_100a.
//
// Replaces boxes_overlap_kernel / boxes_iou_bev_kernel (+ launchers) of
// pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-265,378-398 and the ~10 torch
// elementwise kernels of boxes_iou3d_gpu (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:88-121).
//
// Design (one CTA = one TR x TC tile of the (na, nb) matrix):
//   1. cull pass   : two-level exact-conservative circle test on box tiles staged in shared
//                    memory.  First one test per COLUMN against the bounding box of the tile's
//                    row centres (an anchor tile sees ~4 of 100 GT boxes), then one test per
//                    (row, active column).  The tile itself is zero-filled with 16-byte streaming
//                    stores -- > 99 % of an anchor sweep is exactly +0.0 and this regime is
//                    HBM-write bound (4 B / pair).
//   2. compaction  : surviving pairs are appended to a shared-memory queue with one
//                    warp-aggregated atomic per warp, and their boxes are flagged.
//   3. lazy prepare: only flagged boxes get their BoxPre record (4 trig calls, corners,
//                    margin thresholds) -- once per box per tile, never per pair.
//   4. clip pass   : the queue is drained with all lanes busy (no divergence between
//                    "far" and "near" pairs).
// The reference instead runs the full clipping code, incl. 20 sinf/cosf evaluations, for
// every pair in a 16x16 thread block with 208 B of local-memory stack per thread.
#include "common.cuh"
#include "geom.cuh"
#include "../../include/glenet_geom.h"
#include <float.h>
#include <math_constants.h>

namespace glenet {

constexpr int IOU_THREADS = 256;
constexpr int IOU_TR_MAX = 256;            // tile rows (boxes_a)
constexpr int IOU_TC_MAX = 128;            // tile cols (boxes_b)
constexpr int IOU_QCAP = 4 * IOU_THREADS;  // queue capacity; drained when 2 more column steps could overflow it
constexpr int IOU_CTAS_PER_SM = 4;         // register budget the kernel is compiled for (5 was measured: spills, no gain)
constexpr int IOU_ZCHUNK = 4 * 32;         // float4 stores per warp and zero-fill chunk

enum { MODE_OVERLAP = 0, MODE_IOU_BEV = 1, MODE_IOU3D = 2 };

#ifdef GLENET_PHASE_TIMING   // developer instrumentation: accumulated clock64() per phase, thread 0 of every CTA
__device__ unsigned long long g_phase_cycles[8];
__device__ int g_dbg_flags;   // bit 0: skip the clip pass, bit 1: skip the zero fill (timing experiments only)
#define PHASE_MARK(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_phase_cycles[k], (unsigned long long)(now_ - t_phase_)); t_phase_ = now_; } } while (0)
#define PHASE_INIT long long t_phase_ = clock64()
#else
#define PHASE_MARK(k) do { } while (0)
#define PHASE_INIT do { } while (0)
#endif

template <int BPS>
struct __align__(16) IouSmem {
    float4 row[IOU_TR_MAX];                // {cx, cy, cull radius, -}
    float ccx[IOU_TC_MAX], ccy[IOU_TC_MAX], crad[IOU_TC_MAX];   // SoA so that 4 consecutive columns are one LDS.128
    float rpre[IOU_TR_MAX * BPS];
    float cpre[IOU_TC_MAX * BPS];
    float qres[IOU_QCAP];                  // clipped results, parked until the tile's zero fill is complete
    unsigned short queue[IOU_QCAP];        // (row << 8) | col  (row < 256, col < 128)
    float red[IOU_THREADS / 32][5];
    unsigned char rflag[IOU_TR_MAX], cflag[IOU_TC_MAX], act[IOU_TC_MAX];
    int qcount, nact, zchunk;
};

template <int MODE>
__device__ __forceinline__ float finish_pair(const float* a, const float* b, float ov) {
    if (MODE == MODE_OVERLAP) return ov;
    if (MODE == MODE_IOU_BEV) return iou_from_overlap(a[BP_AREA], b[BP_AREA], ov);
    return iou3d_from_overlap(a, b, ov);
}

// Zero-fill of the tile, chunked so that any warp can take part whenever it has nothing else to do:
// the stores are fire-and-forget, which is what lets them overlap the clip pass of the other warps.
template <typename SM>
__device__ __forceinline__ void zero_fill_tile(SM& sm, float* __restrict__ out_tile, int tr, int tc, int nb, bool vec) {
    const int lane = threadIdx.x & 31;
    if (vec) {
        const int nq = tc >> 2, nquads = tr * nq;
        const int nchunks = (nquads + IOU_ZCHUNK - 1) / IOU_ZCHUNK;
        const bool contiguous = (nq * 4 == nb);
        for (;;) {
            int ch = 0;
            if (lane == 0) ch = atomicAdd(&sm.zchunk, 1);
            ch = __shfl_sync(0xffffffffu, ch, 0);
            if (ch >= nchunks) break;
#pragma unroll
            for (int k = 0; k < IOU_ZCHUNK / 32; ++k) {
                const int q = ch * IOU_ZCHUNK + k * 32 + lane;
                if (q < nquads) {
                    float4* dst;
                    if (contiguous) dst = reinterpret_cast<float4*>(out_tile) + q;
                    else { const int r = q / nq; dst = reinterpret_cast<float4*>(out_tile + (size_t)r * nb) + (q - r * nq); }
                    *dst = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    } else {
        const int npairs = tr * tc;
        const int nchunks = (npairs + IOU_ZCHUNK - 1) / IOU_ZCHUNK;
        for (;;) {
            int ch = 0;
            if (lane == 0) ch = atomicAdd(&sm.zchunk, 1);
            ch = __shfl_sync(0xffffffffu, ch, 0);
            if (ch >= nchunks) break;
#pragma unroll
            for (int k = 0; k < IOU_ZCHUNK / 32; ++k) {
                const int p = ch * IOU_ZCHUNK + k * 32 + lane;
                if (p < npairs) { const int r = p / tc; out_tile[(size_t)r * nb + (p - r * tc)] = 0.f; }
            }
        }
    }
}

// Clip the queued pairs.  Results are parked until the tile's zero fill is complete (barrier),
// then overwrite the zeros (parked in shared memory meanwhile).  Warps without queued pairs go straight to zero filling, the others join
// when their pairs are done -- streaming stores and clipping overlap inside the CTA.
template <int MODE, bool FMA>
__device__ __forceinline__ void drain_queue(IouSmem<MODE == MODE_IOU3D ? BP_STRIDE : BP_STRIDE_BEV>& sm, const float* __restrict__ A, const float* __restrict__ B,
                                            const float4* __restrict__ trigA, const float4* __restrict__ trigB,
                                            int r0, int c0, int tr, int tc, int nb, float* __restrict__ out, bool vec) {
    constexpr int BPS = (MODE == MODE_IOU3D) ? BP_STRIDE : BP_STRIDE_BEV;
    const int tid = threadIdx.x;
    const int n = sm.qcount;
    PHASE_INIT;
    // lazy per-box preparation of the boxes that take part in at least one queued pair
    for (int i = tid; i < tr + tc; i += IOU_THREADS) {
        const bool is_row = i < tr;
        const int k = is_row ? i : i - tr;
        unsigned char* flag = is_row ? &sm.rflag[k] : &sm.cflag[k];
        if (*flag == 1) {
            const int g = is_row ? r0 + k : c0 + k;
            const float* box = (is_row ? A : B) + (size_t)g * 7;
            const float4* trig = is_row ? trigA : trigB;
            const float4 t4 = trig ? trig[g] : device_trig(box[6]);
            box_prepare<FMA, MODE == MODE_IOU3D>(box, t4, (is_row ? sm.rpre : sm.cpre) + k * BPS);
            *flag = 2;
        }
    }
    __syncthreads();
    PHASE_MARK(3);
#ifdef GLENET_PHASE_TIMING
    const int dbg = g_dbg_flags;
#else
    const int dbg = 0;
#endif
    for (int q = tid; q < ((dbg & 1) ? 0 : n); q += IOU_THREADS) {
        const unsigned int e = sm.queue[q];
        const float* a = sm.rpre + (e >> 8) * BPS;
        const float* b = sm.cpre + (e & 255) * BPS;
        sm.qres[q] = finish_pair<MODE>(a, b, box_overlap<FMA>(a, b));
    }
    PHASE_MARK(7);
    if (!(dbg & 2)) zero_fill_tile(sm, out + (size_t)r0 * nb + c0, tr, tc, nb, vec);   // no-op once every chunk has been taken
    __syncthreads();
    PHASE_MARK(4);
    for (int q = tid; q < n; q += IOU_THREADS) {
        const unsigned int e = sm.queue[q];
        out[(size_t)(r0 + (e >> 8)) * nb + (c0 + (e & 255))] = sm.qres[q];
    }
    __syncthreads();
    if (tid == 0) sm.qcount = 0;
    __syncthreads();
    PHASE_MARK(5);
}

// append the lanes' surviving pairs with one atomic per warp
template <typename SM>
__device__ __forceinline__ void enqueue_heavy(SM& sm, unsigned int heavy, int r, int c, int lane) {
    const unsigned int m = __ballot_sync(0xffffffffu, heavy != 0);
    if (!m) return;
    int qb = 0;
    if (lane == 0) qb = atomicAdd(&sm.qcount, __popc(m));
    qb = __shfl_sync(0xffffffffu, qb, 0);
    if (heavy) {
        sm.queue[qb + __popc(m & ((1u << lane) - 1))] = (unsigned short)((r << 8) | c);
        if (sm.rflag[r] == 0) sm.rflag[r] = 1;   // 0 = unused, 1 = wanted, 2 = prepared
        if (sm.cflag[c] == 0) sm.cflag[c] = 1;
    }
}

template <int MODE, bool FMA>
__global__ void __launch_bounds__(IOU_THREADS, IOU_CTAS_PER_SM)
iou_tile_kernel(const float* __restrict__ A, int na, const float* __restrict__ B, int nb,
                const float4* __restrict__ trigA, const float4* __restrict__ trigB,
                float* __restrict__ out, int TR, int TC, int col_tiles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = IouSmem<MODE == MODE_IOU3D ? BP_STRIDE : BP_STRIDE_BEV>;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_r = blockIdx.x / col_tiles, tile_c = blockIdx.x - tile_r * col_tiles;
    const int r0 = tile_r * TR, c0 = tile_c * TC;
    const int tr = min(TR, na - r0), tc = min(TC, nb - c0);
    PHASE_INIT;

    // ---- stage the tile's boxes (centre + cull radius) and the bounding box of the row centres
    float minx = FLT_MAX, maxx = -FLT_MAX, miny = FLT_MAX, maxy = -FLT_MAX, maxr = 0.f;
    for (int i = tid; i < tr + tc; i += IOU_THREADS) {
        const bool is_row = i < tr;
        const int k = is_row ? i : i - tr;
        const float* box = (is_row ? A + (size_t)(r0 + k) * 7 : B + (size_t)(c0 + k) * 7);
        const float cx = box[0], cy = box[1], rad = cull_radius(box);
        if (is_row) {
            sm.row[k] = make_float4(cx, cy, rad, 0.f); sm.rflag[k] = 0;
            minx = fminf(minx, cx); maxx = fmaxf(maxx, cx); miny = fminf(miny, cy); maxy = fmaxf(maxy, cy);
            maxr = (rad != rad) ? CUDART_INF_F : fmaxf(maxr, rad);   // a NaN radius must not be dropped by fmaxf
        } else { sm.ccx[k] = cx; sm.ccy[k] = cy; sm.crad[k] = rad; sm.cflag[k] = 0; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        minx = fminf(minx, __shfl_xor_sync(0xffffffffu, minx, o)); maxx = fmaxf(maxx, __shfl_xor_sync(0xffffffffu, maxx, o));
        miny = fminf(miny, __shfl_xor_sync(0xffffffffu, miny, o)); maxy = fmaxf(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
        maxr = fmaxf(maxr, __shfl_xor_sync(0xffffffffu, maxr, o));
    }
    if (lane == 0) { sm.red[warp][0] = minx; sm.red[warp][1] = maxx; sm.red[warp][2] = miny; sm.red[warp][3] = maxy; sm.red[warp][4] = maxr; }
    if (tid == 0) { sm.qcount = 0; sm.nact = 0; sm.zchunk = 0; }
    __syncthreads();
    PHASE_MARK(0);
    minx = sm.red[0][0]; maxx = sm.red[0][1]; miny = sm.red[0][2]; maxy = sm.red[0][3]; maxr = sm.red[0][4];
#pragma unroll
    for (int w = 1; w < IOU_THREADS / 32; ++w) {
        minx = fminf(minx, sm.red[w][0]); maxx = fmaxf(maxx, sm.red[w][1]);
        miny = fminf(miny, sm.red[w][2]); maxy = fmaxf(maxy, sm.red[w][3]); maxr = fmaxf(maxr, sm.red[w][4]);
    }

    // ---- active columns: a column whose circle cannot reach the rows' bounding box is culled for the
    //      whole tile with ONE test (rows with a NaN centre produce no polygon vertex in the reference
    //      either, so leaving them out of the bounding box is exact).  NaN in the column => stays active.
    for (int c = tid; c < tc; c += IOU_THREADS) {
        const float cx = sm.ccx[c], cy = sm.ccy[c];
        const float ddx = fmaxf(fmaxf(minx - cx, cx - maxx), 0.f), ddy = fmaxf(fmaxf(miny - cy, cy - maxy), 0.f);
        const float rr = maxr + sm.crad[c];
        const bool far = (cx == cx) && (cy == cy) && (ddx * ddx + ddy * ddy > rr * rr);
        if (!far) sm.act[atomicAdd(&sm.nact, 1)] = (unsigned char)c;
    }

    float* out_tile = out + (size_t)r0 * nb + c0;
    (void)out_tile;
    const bool vec = ((nb & 3) == 0) && ((c0 & 3) == 0) && ((tc & 3) == 0) && ((((uintptr_t)out) & 15) == 0);
    __syncthreads();   // act[] / nact complete
    PHASE_MARK(1);

    // ---- per-pair circle test on the active columns only; survivors go to the queue.
    //      One row per thread (TR <= 256 = block size), uniform loop over the active columns whose
    //      data is a shared-memory broadcast; the queue is checked every 2 columns (<= 512 appends).
    const int nact = sm.nact;
    {
        // thread -> (row, column group): rows padded to a power of two so that small tiles still use all threads
        int trp = 32;
        while (trp < tr) trp <<= 1;
        const int groups = IOU_THREADS / trp, row = tid & (trp - 1), grp = tid / trp;
        const bool has_row = row < tr;
        const float4 rw = has_row ? sm.row[row] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int iters = (nact + groups - 1) / groups;
        for (int it0 = 0; it0 < iters; it0 += 2) {
            const int qc = sm.qcount;
            __syncthreads();   // everyone has read the count before anyone appends again => the branch is uniform
            if (qc > IOU_QCAP - 2 * IOU_THREADS) {   // dense tiles only
                drain_queue<MODE, FMA>(sm, A, B, trigA, trigB, r0, c0, tr, tc, nb, out, vec);
            }
            const int it1 = min(iters, it0 + 2);
            for (int it = it0; it < it1; ++it) {
                const int k = it * groups + grp;
                const bool valid = has_row && k < nact;
                const int c = valid ? sm.act[k] : 0;
                const float dx = rw.x - sm.ccx[c], dy = rw.y - sm.ccy[c], rr = rw.z + sm.crad[c];
                // NaN anywhere => the comparison is false => not culled => the clip pass decides, like the reference
                const unsigned int heavy = (valid && !(dx * dx + dy * dy > rr * rr)) ? 1u : 0u;
                enqueue_heavy(sm, heavy, row, c, lane);
            }
            __syncthreads();
        }
    }
    PHASE_MARK(2);
    drain_queue<MODE, FMA>(sm, A, B, trigA, trigB, r0, c0, tr, tc, nb, out, vec);
    PHASE_MARK(6);
}

// out[i] = f(a[i], b[i / group]) -- every pair is "heavy" by construction (CVAE samples vs their GT)
constexpr int ALIGNED_THREADS = 128;
template <int MODE, bool FMA>
__global__ void __launch_bounds__(ALIGNED_THREADS)
iou_aligned_kernel(const float* __restrict__ A, int na, const float* __restrict__ B, int group,
                   float* __restrict__ out) {
    const int tid = threadIdx.x;
    const int i = blockIdx.x * ALIGNED_THREADS + tid;
    if (i >= na) return;
    float a[BP_STRIDE], b[BP_STRIDE];   // statically indexed => registers
    const float* ba = A + (size_t)i * 7;
    const float* bb = B + (size_t)(i / group) * 7;
    float out_v = 0.f;
    const float ddx = ba[0] - bb[0], ddy = ba[1] - bb[1];
    const float rr = cull_radius(ba) + cull_radius(bb);
    if (!(ddx * ddx + ddy * ddy > rr * rr)) {
        box_prepare<FMA>(ba, device_trig(ba[6]), a);
        box_prepare<FMA>(bb, device_trig(bb[6]), b);
        const float ov = box_overlap_unrolled<FMA>(a, b);
        out_v = finish_pair<MODE>(a, b, ov);
    }
    out[i] = out_v;
}

#ifdef GLENET_PHASE_TIMING
static int g_debug_tile_rows = 0;   // developer override (debug build only)
#endif
static void pick_tiles(int na, int nb, int& TR, int& TC, int& row_tiles, int& col_tiles) {
    col_tiles = (nb + IOU_TC_MAX - 1) / IOU_TC_MAX;
    TC = ((nb + col_tiles - 1) / col_tiles + 3) / 4 * 4;   // multiple of 4 keeps every tile on the 16-byte store path
    // aim for >= 4 CTAs per SM (148 SMs) before growing the row tile
    long want = (long)IOU_CTAS_PER_SM * 148;
    long tr = ((long)na * col_tiles + want - 1) / want;
    tr = (tr + 31) / 32 * 32;
    if (tr < 32) tr = 32;
    if (tr > IOU_TR_MAX) tr = IOU_TR_MAX;
    TR = (int)tr;
#ifdef GLENET_PHASE_TIMING
    if (g_debug_tile_rows) TR = g_debug_tile_rows;
#endif
    row_tiles = (na + TR - 1) / TR;
}

template <int MODE, bool FMA>
static int launch_iou(const float* A, const float* trigA, int na, const float* B, const float* trigB, int nb,
                      float* out, cudaStream_t stream, const char* what) {
    if (na < 0 || nb < 0) return fail(GLENET_EINVAL, "%s: negative box count", what);
    if (na == 0 || nb == 0) return GLENET_OK;
    if (!A || !B || !out) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (!FMA && (!trigA || !trigB)) return fail(GLENET_EINVAL, "%s: CPU dialect needs host-evaluated trig tables", what);
    int TR, TC, row_tiles, col_tiles;
    pick_tiles(na, nb, TR, TC, row_tiles, col_tiles);
    auto kernel = iou_tile_kernel<MODE, FMA>;
    static bool attr_done = false;   // per template instantiation
    if (!attr_done) {
        int rc = set_smem(kernel, sizeof(IouSmem<MODE == MODE_IOU3D ? BP_STRIDE : BP_STRIDE_BEV>), what);
        if (rc) return rc;
        attr_done = true;
    }
    const long tiles = (long)row_tiles * col_tiles;
    if (tiles > 0x7fffffffL) return fail(GLENET_EINVAL, "%s: too many tiles", what);
    kernel<<<(unsigned)tiles, IOU_THREADS, sizeof(IouSmem<MODE == MODE_IOU3D ? BP_STRIDE : BP_STRIDE_BEV>), stream>>>(
        A, na, B, nb, reinterpret_cast<const float4*>(trigA), reinterpret_cast<const float4*>(trigB), out, TR, TC,
        col_tiles);
    return check_launch(what);
}

}  // namespace glenet

using namespace glenet;

extern "C" {

#ifdef GLENET_PHASE_TIMING
void glenet_debug_set_tile_rows(int tr) { g_debug_tile_rows = tr; }
void glenet_debug_set_flags(int f) { cudaMemcpyToSymbol(g_dbg_flags, &f, sizeof(int)); }
// developer-only: read and reset the per-phase cycle accumulators of iou_tile_kernel
int glenet_debug_iou_phase_cycles(unsigned long long* host_out8) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out8, g_phase_cycles, sizeof(unsigned long long) * 8);
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
    return 0;
}
#endif

int glenet_boxes_overlap_bev_gpu(const float* a, int na, const float* b, int nb, float* out, glenet_stream_t s) {
    return launch_iou<MODE_OVERLAP, true>(a, nullptr, na, b, nullptr, nb, out, (cudaStream_t)s, "glenet_boxes_overlap_bev_gpu");
}
int glenet_boxes_iou_bev_gpu(const float* a, int na, const float* b, int nb, float* out, glenet_stream_t s) {
    return launch_iou<MODE_IOU_BEV, true>(a, nullptr, na, b, nullptr, nb, out, (cudaStream_t)s, "glenet_boxes_iou_bev_gpu");
}
int glenet_boxes_iou3d_gpu(const float* a, int na, const float* b, int nb, float* out, glenet_stream_t s) {
    return launch_iou<MODE_IOU3D, true>(a, nullptr, na, b, nullptr, nb, out, (cudaStream_t)s, "glenet_boxes_iou3d_gpu");
}
int glenet_boxes_iou_bev_cpu_dialect(const float* a, const float* trig_a, int na, const float* b, const float* trig_b,
                                     int nb, float* out, glenet_stream_t s) {
    if (((uintptr_t)trig_a | (uintptr_t)trig_b) & 15) return fail(GLENET_EALIGN, "%s: trig tables must be 16-byte aligned", "glenet_boxes_iou_bev_cpu_dialect");
    return launch_iou<MODE_IOU_BEV, false>(a, trig_a, na, b, trig_b, nb, out, (cudaStream_t)s, "glenet_boxes_iou_bev_cpu_dialect");
}

int glenet_boxes_iou_aligned_gpu(int mode, const float* a, int na, const float* b, int group, float* out,
                                 glenet_stream_t s) {
    const char* what = "glenet_boxes_iou_aligned_gpu";
    if (na < 0 || group <= 0 || mode < 0 || mode > 2) return fail(GLENET_EINVAL, "%s: bad argument", what);
    if (na == 0) return GLENET_OK;
    if (!a || !b || !out) return fail(GLENET_EINVAL, "%s: null pointer", what);
    const unsigned grid = (na + ALIGNED_THREADS - 1) / ALIGNED_THREADS;
    cudaStream_t st = (cudaStream_t)s;
    if (mode == 0) iou_aligned_kernel<MODE_OVERLAP, true><<<grid, ALIGNED_THREADS, 0, st>>>(a, na, b, group, out);
    else if (mode == 1) iou_aligned_kernel<MODE_IOU_BEV, true><<<grid, ALIGNED_THREADS, 0, st>>>(a, na, b, group, out);
    else iou_aligned_kernel<MODE_IOU3D, true><<<grid, ALIGNED_THREADS, 0, st>>>(a, na, b, group, out);
    return check_launch(what);
}

}  // extern "C"
