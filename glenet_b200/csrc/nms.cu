// Bitmask NMS (rotated and axis-aligned) entirely on the device, for sm_100a.
//
// Replaces nms_kernel / nms_normal_kernel (pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:267-372)
// and the host side of nms_gpu / nms_normal_gpu (pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:90-186):
// cudaMalloc of the mask, blocking D2H copy of N*ceil(N/64) words, serial host sweep.
//
//   mask kernel : one CTA per 64x64 tile of the UPPER triangle only (the reference launches
//                 the full square and never reads the lower half).  Pairs are circle-culled,
//                 compacted into a shared queue, and the survivors go through the same
//                 clipping code as boxes_iou_bev (row box = box_a, column box = box_b, as in
//                 the reference) -> bit (i, j) = iou(i, j) > thresh, j > i.
//   sweep kernel: one CTA per frame.  For every 64-box chunk, warp 0 resolves the diagonal
//                 tile with a find-first-set loop over the not-yet-suppressed bits (only kept
//                 boxes cost an iteration) from shared-memory copies of the diagonal and
//                 super-diagonal words; the other warps OR the kept rows' remaining mask words
//                 into the running suppression words one step later.  keep[] and the count stay
//                 on the device.
//   spatial path (1024 <= n <= 10240, thresh >= 0; "spatial tiles" below): the pairs are enumerated over
//                 groups of neighbouring boxes instead of score-order tiles -- only groups whose bounding
//                 boxes meet are paired -- and the greedy sweep runs per connected component of that
//                 group graph, one warp per component.  Same mask bits, same keep list.
#include "common.cuh"
#include "geom.cuh"
#include "clip.cuh"
#include "../../include/glenet_geom.h"
#include <atomic>
#include <float.h>
#include <stdlib.h>

#ifndef GLENET_NMS_DEFER      // 1: the exact clips of a tile are appended to a global list and run as one dense kernel
#define GLENET_NMS_DEFER 1
#endif
#ifndef GLENET_NMS_APPROX     // 1: decide pairs far from the threshold from the approximate true overlap (no clip)
#define GLENET_NMS_APPROX 1
#endif

namespace glenet {

#ifdef GLENET_PHASE_TIMING   // developer instrumentation of the sweep (tools/nms_sweep_timing.py)
__device__ unsigned long long g_sweep_cycles[4];   // [0] warp 0 busy, [1] whole loop, [2] job warps busy (warp 1), [3] steps
#endif

#ifndef GLENET_NMS_CTAS   // resident CTAs per SM the mask kernel is compiled for (register budget: 4 -> 64, 5 -> 48 registers; measured 340.6 vs 326.8 us for 8 x 4096 once the fallback clip was out of line)
#define GLENET_NMS_CTAS 5
#endif
constexpr int NMS_THREADS = 256;
constexpr int NMS_TILE = 64;
constexpr int SWEEP_THREADS = 512;

// entries of the deferred-clip list (pairs within ~0.03 IoU of the threshold: ~6 n per frame on dense proposal clusters)
static inline unsigned long long nms_list_cap(int frames, int n) { return (unsigned long long)frames * (unsigned long long)n * 32ull; }
constexpr int NMS_PASS = NMS_THREADS;       // pairs per pass of the phased clip: one per lane
constexpr int NBS = BP_STRIDE_BEV;          // BoxPre stride (no z terms in NMS)
struct NmsSmem {
    // vertex slots of the phased clip (clip.cuh); the circle-test queue aliases their first half: it is dead once queue2 exists
    union { float2 verts[NMS_PASS * CLIP_SLOTS]; unsigned short queue[NMS_TILE * NMS_TILE]; };
    float rcx[NMS_TILE], rcy[NMS_TILE], rrad[NMS_TILE];
    float ccx[NMS_TILE], ccy[NMS_TILE], crad[NMS_TILE];
    float rpre[NMS_TILE * NBS];
    float cpre[NMS_TILE * NBS];            // must follow rpre directly (nms_tile_clip<true> addresses it as records 64.. of rpre)
    union { unsigned long long bits[NMS_TILE]; int idx[2 * NMS_TILE]; };   // mask words of the tile's rows | (spatial tiles) the boxes' score-order indices
    unsigned int wl[NMS_THREADS / 32][32 * CLIP_SLOTS];   // per-warp work lists of the clip's phase B
    unsigned short queue2[NMS_TILE * NMS_TILE];   // pairs whose IoU could exceed the threshold: the ones that are clipped
    unsigned char rflag[NMS_TILE], cflag[NMS_TILE];
    int qcount, q2count;
};
static_assert(sizeof(unsigned short) * NMS_TILE * NMS_TILE <= sizeof(float2) * NMS_PASS * CLIP_SLOTS, "queue aliases verts");

// iou_normal (iou3d_nms_kernel.cu:314-325) as compiled in nms_normal_kernel: a = row box, b = column box,
// Sa + Sb is contracted to fma(b.dx, b.dy, Sa).
__device__ __forceinline__ float iou_normal_pair(const float* __restrict__ a, const float* __restrict__ b) {
    const float left = fmaxf(__fmaf_rn(a[3], -0.5f, a[0]), __fmaf_rn(b[3], -0.5f, b[0]));
    const float right = fminf(__fmaf_rn(a[3], 0.5f, a[0]), __fmaf_rn(b[3], 0.5f, b[0]));
    const float top = fmaxf(__fmaf_rn(a[4], -0.5f, a[1]), __fmaf_rn(b[4], -0.5f, b[1]));
    const float bottom = fminf(__fmaf_rn(a[4], 0.5f, a[1]), __fmaf_rn(b[4], 0.5f, b[1]));
    const float width = fmaxf(__fsub_rn(right, left), 0.f), height = fmaxf(__fsub_rn(bottom, top), 0.f);
    const float inter = __fmul_rn(width, height);
    const float sa = __fmul_rn(a[3], a[4]);
    const float den = fmaxf(__fsub_rn(__fmaf_rn(b[3], b[4], sa), inter), 1e-8f);
    return __fdiv_rn(inter, den);
}

// linear index over the upper triangle (row-major, cb >= rb) -> (rb, cb)
__device__ __forceinline__ void tri_decode(int t, int nblk, int& rb, int& cb) {
    // rows before rb hold sum_{k<rb} (nblk - k) = rb*nblk - rb*(rb-1)/2 tiles
    const float fn = (float)nblk + 0.5f;
    int r = (int)(fn - sqrtf(fmaxf(fn * fn - 2.f * (float)t, 0.f)));
    r = max(0, min(r, nblk - 1));
    while (r > 0 && r * nblk - r * (r - 1) / 2 > t) --r;
    while ((r + 1) * nblk - (r + 1) * r / 2 <= t) ++r;
    rb = r;
    cb = r + (t - (r * nblk - r * (r - 1) / 2));
}

// The in-tile exact clip over queue2 (all pairs when thresh < 0, or when the deferred-clip list is full), NMS_PASS pairs at a time.
// SP (spatial tiles): rows and columns are arbitrary boxes -- entry bit 12 says that the column box has the lower score index
// (it is box_a then), and the bit goes straight to the global mask word of (lower index, higher index).
template <bool SP>
__device__ __noinline__ void nms_tile_clip(NmsSmem& sm, int nq2, float thresh, unsigned long long* __restrict__ mask = nullptr, int col_blocks = 0) {
    const int tid = threadIdx.x, warp = tid >> 5;
    auto set_bit = [&](int p, float ov, const float* a, const float* b) {
        if (iou_from_overlap(a[BP_AREA], b[BP_AREA], ov) > thresh) {
            if (SP) {
                const int ri = sm.idx[(p >> 6) & 63], ci = sm.idx[NMS_TILE + (p & 63)];
                const int i = min(ri, ci), j = max(ri, ci);
                atomicOr(reinterpret_cast<unsigned int*>(mask + (size_t)i * col_blocks + (j >> 6)) + ((j >> 5) & 1), 1u << (j & 31));
            } else {   // 32-bit halves: a native shared-memory atomic instead of a 64-bit CAS loop
                atomicOr(reinterpret_cast<unsigned int*>(&sm.bits[p >> 6]) + ((p >> 5) & 1), 1u << (p & 31));
            }
        }
    };
    for (int base = 0; base < nq2; base += NMS_PASS) {
        // one pair per lane; A (result bits, corners), B (the warp's crossings pooled) and C (sort + fan) are warp-local
        const bool live = base + tid < nq2;
        const int p = live ? sm.queue2[base + tid] : 0;
        const bool swap = SP && ((p >> 12) & 1);
        const int ra = (p >> 6) & 63, cb_ = p & 63;
        const float* a = swap ? sm.cpre + cb_ * NBS : sm.rpre + ra * NBS;
        const float* b = swap ? sm.rpre + ra * NBS : sm.cpre + cb_ * NBS;
        float2* slots = sm.verts + tid * CLIP_SLOTS;
        const unsigned int w = clip_pair_tests<true>(a, b, live);
        const unsigned int hits = clip_hits16(w);
        const int cnt = __popc(hits) + __popc(clip_corners8(w));
        const bool fast = cnt >= 3 && cnt <= CLIP_SLOTS;
        if (fast) clip_write_corners(a, b, w, slots);
        // the pooled phases address the records as (array base, record number): cpre follows rpre in NmsSmem, so with
        // rpre as the base of both operands a column record is number 64 + c and either box can be box_a
        const unsigned int ia = SP ? (unsigned int)(swap ? NMS_TILE + cb_ : ra) : (unsigned int)ra;
        const unsigned int ib = SP ? (unsigned int)(swap ? ra : NMS_TILE + cb_) : (unsigned int)cb_;
        const float* brec = SP ? sm.rpre : sm.cpre;
        clip_warp_points<true>(fast ? hits : 0u, ia, ib, sm.wl[warp], sm.rpre, brec, NBS, sm.verts + (warp * 32) * CLIP_SLOTS);
        // more than eight vertices (corners admitted by the margin next to a crossing): the whole warp, one pair at a time
        const bool slow = cnt > CLIP_SLOTS;
        const float ov_slow = clip_warp_slow<true>(slow, w, ia, ib, sm.rpre, brec, NBS, reinterpret_cast<float2*>(sm.wl[warp]));
        if (live) set_bit(p, slow ? ov_slow : (fast ? clip_area8<true>(slots, cnt) : 0.f), a, b);
    }
}

template <bool NORMAL>
__global__ void __launch_bounds__(NMS_THREADS, GLENET_NMS_CTAS)
nms_mask_kernel(const float* __restrict__ boxes_all, int n, float thresh, unsigned long long* __restrict__ mask_all,
                int col_blocks, int tiles_per_frame, unsigned long long* __restrict__ list_count, unsigned long long* __restrict__ list,
                unsigned long long list_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem& sm = *reinterpret_cast<NmsSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x / tiles_per_frame;
    const int t = blockIdx.x - frame * tiles_per_frame;
    const float* boxes = boxes_all + (size_t)frame * n * 7;
    unsigned long long* mask = mask_all + (size_t)frame * n * col_blocks;
    int rb, cb;
    tri_decode(t, col_blocks, rb, cb);
    const int r0 = rb * NMS_TILE, c0 = cb * NMS_TILE;
    const int tr = min(NMS_TILE, n - r0), tc = min(NMS_TILE, n - c0);

    if (NORMAL) {
        // cheap pair function: no culling or queueing, every thread evaluates its pairs directly
        float* rraw = sm.rpre;   // reuse: 64 x 7 raw floats each
        float* craw = sm.cpre;
        for (int i = tid; i < tr * 7; i += NMS_THREADS) rraw[i] = boxes[(size_t)r0 * 7 + i];
        for (int i = tid; i < tc * 7; i += NMS_THREADS) craw[i] = boxes[(size_t)c0 * 7 + i];
        if (tid < NMS_TILE) sm.bits[tid] = 0ull;
        __syncthreads();
        // thread -> (row = tid / 4, 16 columns starting at (tid % 4) * 16)
        const int r = tid >> 2, cs = (tid & 3) * 16;
        unsigned long long w = 0ull;
        if (r < tr) {
            for (int c = cs; c < min(cs + 16, tc); ++c) {
                if (rb == cb && c <= r) continue;
                if (iou_normal_pair(rraw + r * 7, craw + c * 7) > thresh) w |= 1ull << c;
            }
        }
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        if ((tid & 3) == 0 && r < tr) mask[(size_t)(r0 + r) * col_blocks + cb] = w;
        return;
    }

    for (int i = tid; i < tr + tc; i += NMS_THREADS) {
        const bool is_row = i < tr;
        const int k = is_row ? i : i - tr;
        const float* box = boxes + (size_t)((is_row ? r0 : c0) + k) * 7;
        const float cx = box[0], cy = box[1], rad = cull_radius(box);
        if (is_row) { sm.rcx[k] = cx; sm.rcy[k] = cy; sm.rrad[k] = rad; sm.rflag[k] = 0; }
        else        { sm.ccx[k] = cx; sm.ccy[k] = cy; sm.crad[k] = rad; sm.cflag[k] = 0; }
    }
    if (tid < NMS_TILE) sm.bits[tid] = 0ull;
    if (tid == 0) { sm.qcount = 0; sm.q2count = 0; }
    __syncthreads();

    // cull pass over the 64 x 64 pairs (only j > i on the diagonal tile, iou3d_nms_kernel.cu:300-302)
    const bool diag = rb == cb;
    const bool all_pairs = thresh < 0.f;   // an IoU of exactly 0 exceeds a negative threshold: nothing may be culled
    // thread = one column c and 16 rows (r = 4 k + tid / 64): the circle tests go into a register bitmask, then ONE
    // warp-aggregated reservation appends all survivors (the queue holds the whole tile, so it cannot overflow)
    constexpr int NMS_RPT = NMS_TILE * NMS_TILE / NMS_THREADS;   // 16 rows per thread
    const int c = tid & (NMS_TILE - 1), rq = tid >> 6;
    unsigned int hits = 0u;
    if (c < tc) {
        // rows this thread may pair its column with: r = 4 k + rq below tr, and above the diagonal (r < c) on a diagonal tile --
        // one mask instead of two tests per pair
        const int lim = diag ? min(tr, c) : tr;
        const int kmax = lim > rq ? (lim - rq + (NMS_THREADS / NMS_TILE) - 1) / (NMS_THREADS / NMS_TILE) : 0;
        const unsigned int valid = kmax >= NMS_RPT ? (1u << NMS_RPT) - 1u : (1u << kmax) - 1u;
        if (all_pairs) hits = valid;
        else {
            const float ccx = sm.ccx[c], ccy = sm.ccy[c], ccr = sm.crad[c];
#pragma unroll
            for (int k = 0; k < NMS_RPT; ++k) {
                const int r = k * (NMS_THREADS / NMS_TILE) + rq;
                const float ddx = sm.rcx[r] - ccx, ddy = sm.rcy[r] - ccy, rr = sm.rrad[r] + ccr;
                hits |= (!(ddx * ddx + ddy * ddy > rr * rr) ? 1u : 0u) << k;     // NaN anywhere: not culled
            }
            hits &= valid;
        }
    }
    {
        const int cnt = __popc(hits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
        if (wtotal) {
            int qb = 0;
            if (lane == 31) qb = atomicAdd(&sm.qcount, wtotal);
            qb = __shfl_sync(0xffffffffu, qb, 31) + incl - cnt;
            if (hits && sm.cflag[c] == 0) sm.cflag[c] = 1;
            while (hits) {
                const int k = __ffs(hits) - 1;
                hits &= hits - 1;
                const int r = k * (NMS_THREADS / NMS_TILE) + rq;
                sm.queue[qb++] = (unsigned short)((r << 6) | c);
                if (sm.rflag[r] == 0) sm.rflag[r] = 1;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < tr + tc; i += NMS_THREADS) {
        const bool is_row = i < tr;
        const int k = is_row ? i : i - tr;
        if ((is_row ? sm.rflag[k] : sm.cflag[k]) == 1) {
            const float* box = boxes + (size_t)((is_row ? r0 : c0) + k) * 7;
            box_prepare<true, false>(box, device_trig(box[6]), (is_row ? sm.rpre : sm.cpre) + k * NBS);
        }
    }
    __syncthreads();
    // Bound filter: a pair is clipped only if an upper bound of the IoU the reference can compute for it exceeds
    // the threshold (with a safety margin far above float rounding); the others cannot set a mask bit.  At
    // thresh = 0.7 this drops ~45 % of the circle-test survivors of a proposal cluster.  NaN anywhere => clipped.
    const int nq = sm.qcount;
    const float thr_lo = thresh * (1.f - 1e-3f) - 1e-5f, thr_hi = thresh * (1.f + 1e-3f) + 1e-5f;
    for (int q0 = 0; q0 < nq; q0 += NMS_THREADS) {
        const int q = q0 + tid;
        int p = 0;
        bool need = false;
        if (q < nq) {
            p = sm.queue[q];
            const float* a = sm.rpre + (p >> 6) * NBS;
            const float* b = sm.cpre + (p & 63) * NBS;
            const float ub = overlap_upper_bound(a, b);
            const float iou_ub = ub / fmaxf(a[BP_AREA] + b[BP_AREA] - ub, 1e-8f);
            need = all_pairs || (!(ub <= 0.f) && !(iou_ub <= thr_lo)) || !(a[BP_AREA] + b[BP_AREA] > ub);   // degenerate areas: let the clip decide
#if GLENET_NMS_APPROX
            // Second filter: the true intersection area, approximately (geom.cuh: overlap_approx, ~240 instructions).  The
            // reference's overlap lies in [approx - slack, approx + band + slack]; a pair whose IoU is on one side of the
            // threshold for that whole interval is decided here, only the rest (IoU within ~0.03 of the threshold) is clipped.
            if (need && !all_pairs && overlap_approx_usable(a, b)) {
                const float ova = overlap_approx(a, b);
                float slack, band;
                overlap_approx_band(a, b, slack, band);
                const float s = a[BP_AREA] + b[BP_AREA];
                const float hi = ova + band + slack, lo = ova - slack;
                if (hi / fmaxf(s - hi, 1e-8f) <= thr_lo && s > hi) need = false;                    // cannot reach the threshold
                else if (lo > 0.f && s > lo && lo / fmaxf(s - lo, 1e-8f) > thr_hi) {                // exceeds it for certain
                    need = false;
                    atomicOr(reinterpret_cast<unsigned int*>(&sm.bits[p >> 6]) + ((p >> 5) & 1), 1u << (p & 31));
                }                                                                                   // NaN anywhere: both tests fail, the clip decides
            }
#endif
        }
        const unsigned int m = __ballot_sync(0xffffffffu, need);
        if (m) {
            int qb = 0;
            if (lane == 0) qb = atomicAdd(&sm.q2count, __popc(m));
            qb = __shfl_sync(0xffffffffu, qb, 0);
            if (need) sm.queue2[qb + __popc(m & ((1u << lane) - 1))] = (unsigned short)p;
        }
    }
    __syncthreads();
    const int nq2 = sm.q2count;
#if GLENET_NMS_DEFER
    // ---- the pairs that still need the exact clip leave the tile: they are appended to a global list and clipped by
    //      nms_clip_list_kernel, one pair per lane of full warps.  (A tile keeps ~25 of its 4096 pairs; clipping them here
    //      costs the latency of one whole clip per tile -- half of the tile's phase chain -- on a handful of lanes.)
    //      A full list (or thresh < 0, where every pair is kept) falls back to the clip below.
    if (list && nq2 > 0 && !all_pairs) {
        if (tid == 0) {
            const unsigned long long base = atomicAdd(list_count, (unsigned long long)nq2);
            sm.qcount = (base + (unsigned long long)nq2 <= list_cap) ? 1 : 0;
            reinterpret_cast<unsigned long long*>(sm.wl[0])[0] = base;
        }
        __syncthreads();
        if (sm.qcount) {
            const unsigned long long base = reinterpret_cast<unsigned long long*>(sm.wl[0])[0];
            for (int q = tid; q < nq2; q += NMS_THREADS) {
                const int p = sm.queue2[q];
                list[base + q] = ((unsigned long long)frame << 40) | ((unsigned long long)(r0 + (p >> 6)) << 20) | (unsigned long long)(c0 + (p & 63));
            }
            if (tid < tr) mask[(size_t)(r0 + tid) * col_blocks + cb] = sm.bits[tid];
            return;
        }
        // the list is full: what this tile reserved inside it is marked void (only the first overflowing tile reserves
        // anything below the capacity), and the tile clips its pairs itself
        const unsigned long long base = reinterpret_cast<unsigned long long*>(sm.wl[0])[0];
        for (int q = tid; q < nq2; q += NMS_THREADS) if (base + q < list_cap) list[base + q] = ~0ull;
        __syncthreads();   // wl[0] is about to be reused by the clip
    }
#endif
    // ---- phased clip (clip.cuh) over queue2: only when the pairs could not be deferred (out of line: it must not set the
    //      register budget of the phases every tile runs)
    nms_tile_clip<false>(sm, nq2, thresh);
    __syncthreads();
    if (tid < tr) mask[(size_t)(r0 + tid) * col_blocks + cb] = sm.bits[tid];
}

// ---------------------------------------------------------------- spatial tiles
// Proposals arrive in score order, so every 64 x 64 tile of the (i, j) matrix mixes boxes from the whole scene: each of the
// n^2 / 2 pairs costs a circle test, although a box only overlaps the few hundred boxes of its own object.  The mask does
// not care in which order the pairs are examined -- only that every pair (i < j) whose circles meet is, with box i as
// box_a -- so the pairs are enumerated over SPATIAL groups instead:
//   nms_spatial_kernel (one CTA per frame; centre, cull radius and finiteness of every box cached in shared memory):
//       counting sort of the boxes by the Morton code of their centre's cell (32 x 32 cells over the bounding box of the
//       finite centres; boxes with a non-finite term go last); groups = runs of at most 64 consecutive boxes that never leave
//       their 4 x 4 block of cells (a contiguous range of Morton codes), so a group's bounding box is an object-sized patch;
//       the bounding box of every group's cull circles (infinite for a group with a non-finite box); and the list of work
//       items (P <= Q, row quarter) whose group boxes meet -- the only pairs any circle test can pass for;
//   nms_mask_spatial_kernel (persistent CTAs over that list): the tile phases of nms_mask_kernel on 16 rows of group P against
//       group Q; a pair is oriented by its score indices (lower = row of the mask and box_a) and its bit goes straight to the
//       global mask word, which the launcher zeroed.
constexpr int NMS_SP_G = 32;                       // cells per axis of the sorting grid
constexpr int NMS_SP_CELLS = NMS_SP_G * NMS_SP_G;
constexpr int NMS_SP_SUPER = NMS_SP_CELLS / 16;    // 4 x 4 blocks of cells = 16 consecutive Morton codes
constexpr int NMS_SP_THREADS = 1024;
constexpr int NMS_SP_MAX_GROUPS = 255;             // group numbers travel in 8 bits
constexpr int NMS_SP_MAX_N = 160 * NMS_TILE;      // 160 + 64 groups at most (every block of cells can add one partial group); 184 KB of cached boxes
constexpr int NMS_SP_EDGES = 4096;                 // group pairs kept in shared memory for the component labelling
constexpr int NMS_SP_QROWS = 16;                   // rows of a work item (a quarter of group P)

__device__ __forceinline__ unsigned int morton5(unsigned int x, unsigned int y) {   // 5 + 5 bits interleaved
    unsigned int r = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) r |= (((x >> b) & 1u) << (2 * b)) | (((y >> b) & 1u) << (2 * b + 1));
    return r;
}

static size_t nms_spatial_smem(int n) { return (size_t)n * (sizeof(float4) + sizeof(unsigned short)); }

__global__ void __launch_bounds__(NMS_SP_THREADS)
nms_spatial_kernel(const float* __restrict__ boxes_all, int n, int* __restrict__ perm_all, int2* __restrict__ groups_all, int groups_cap,
                   unsigned int* __restrict__ tiles, unsigned int* __restrict__ tile_count,
                   int* __restrict__ labels_all, unsigned long long* __restrict__ members_all, int col_blocks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_box = reinterpret_cast<float4*>(smem_raw);                    // [n] {cx, cy, cull radius (+pad), 1 = all terms finite}
    unsigned short* s_perm = reinterpret_cast<unsigned short*>(s_box + n);  // [n] sorted position -> box
    __shared__ unsigned int cnt[NMS_SP_CELLS + 1];
    __shared__ unsigned int wsum[NMS_SP_THREADS / 32];
    __shared__ float red[4][NMS_SP_THREADS / 32];
    __shared__ float s_b[4];
    __shared__ float4 s_gbox[NMS_SP_MAX_GROUPS];
    __shared__ int2 s_grp[NMS_SP_MAX_GROUPS];
    __shared__ int s_gstart[NMS_SP_SUPER + 1];
    __shared__ int s_lab[NMS_SP_MAX_GROUPS + 1];
    __shared__ int s_changed;
    __shared__ unsigned int s_edge[NMS_SP_EDGES];      // (P << 8) | Q of the meeting pairs, P < Q
    __shared__ unsigned int s_nedge;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x;
    const float* boxes = boxes_all + (size_t)frame * n * 7;
    int* perm = perm_all + (size_t)frame * n;
    int2* groups = groups_all + (size_t)frame * groups_cap;
    for (int i = tid; i <= NMS_SP_CELLS; i += NMS_SP_THREADS) cnt[i] = 0u;
    // one pass over the boxes: cache what the rest needs, bounding box of the finite centres
    float x0 = FLT_MAX, y0 = FLT_MAX, x1 = -FLT_MAX, y1 = -FLT_MAX;
    for (int i = tid; i < n; i += NMS_SP_THREADS) {
        const float* b = boxes + (size_t)i * 7;
        const float cx = b[0], cy = b[1], rad = cull_radius(cx, cy, b[3], b[4]) + 1e-3f;   // the pad is far above the rounding of cx -+ rad
        const bool fin = fabsf(cx) <= FLT_MAX && fabsf(cy) <= FLT_MAX && rad <= FLT_MAX;
        s_box[i] = make_float4(cx, cy, rad, fin ? 1.f : 0.f);
        if (fin) { x0 = fminf(x0, cx); x1 = fmaxf(x1, cx); y0 = fminf(y0, cy); y1 = fmaxf(y1, cy); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if (lane == 0) { red[0][warp] = x0; red[1][warp] = y0; red[2][warp] = x1; red[3][warp] = y1; }
    __syncthreads();
    if (warp == 0) {
        x0 = red[0][lane]; y0 = red[1][lane]; x1 = red[2][lane]; y1 = red[3][lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o));
            x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
        if (lane == 0) {
            // cell = floor((c - lo) * inv), clamped: any mapping will do (it only decides who shares a group)
            const float ex = x1 - x0, ey = y1 - y0;
            s_b[0] = x0; s_b[1] = y0;
            s_b[2] = (ex > 0.f && ex <= FLT_MAX) ? (float)NMS_SP_G / ex : 0.f;
            s_b[3] = (ey > 0.f && ey <= FLT_MAX) ? (float)NMS_SP_G / ey : 0.f;
        }
    }
    __syncthreads();
    const float lox = s_b[0], loy = s_b[1], invx = s_b[2], invy = s_b[3];
    auto cell_of = [&](const float4 b) -> unsigned int {
        if (b.w == 0.f) return NMS_SP_CELLS - 1;
        const int ix = min(NMS_SP_G - 1, max(0, (int)((b.x - lox) * invx))), iy = min(NMS_SP_G - 1, max(0, (int)((b.y - loy) * invy)));
        return morton5((unsigned int)ix, (unsigned int)iy);
    };
    for (int i = tid; i < n; i += NMS_SP_THREADS) atomicAdd(&cnt[cell_of(s_box[i])], 1u);
    __syncthreads();
    {   // exclusive scan of the 1024 cell counts (one per thread); cnt becomes the fill cursor
        const unsigned int v = cnt[tid];
        unsigned int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned int w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            wsum[lane] = w;
        }
        __syncthreads();
        const unsigned int excl = incl - v + (warp ? wsum[warp - 1] : 0u);
        cnt[tid] = excl;
        // groups: every block of 16 consecutive Morton codes cuts its run of boxes into pieces of at most 64
        if ((tid & 15) == 0) s_gstart[tid >> 4] = (int)excl;      // first sorted position of the block (cursor values move below)
        if (tid == 0) s_gstart[NMS_SP_SUPER] = n;
    }
    __syncthreads();
    for (int i = tid; i < n; i += NMS_SP_THREADS) {
        const unsigned int pos = atomicAdd(&cnt[cell_of(s_box[i])], 1u);
        s_perm[pos] = (unsigned short)i;
        perm[pos] = i;
    }
    if (warp == 0) {   // group table: blocks 2 l and 2 l + 1 per lane, exclusive scan of their group counts
        int start[2], len[2], ngr[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            start[h] = s_gstart[2 * lane + h];
            len[h] = s_gstart[2 * lane + h + 1] - start[h];
            ngr[h] = (len[h] + NMS_TILE - 1) / NMS_TILE;
        }
        int incl = ngr[0] + ngr[1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        int g = incl - ngr[0] - ngr[1];
#pragma unroll
        for (int h = 0; h < 2; ++h)
            for (int k = 0; k < ngr[h]; ++k, ++g)
                if (g < NMS_SP_MAX_GROUPS) s_grp[g] = make_int2(start[h] + k * NMS_TILE, min(NMS_TILE, len[h] - k * NMS_TILE));
        if (lane == 31) s_b[0] = __int_as_float(min(incl, NMS_SP_MAX_GROUPS));   // (n <= NMS_SP_MAX_N: never clipped)
    }
    __syncthreads();
    const int ng = __float_as_int(s_b[0]);
    // group bounding boxes: one warp per group, two boxes per lane
    for (int g = warp; g < ng; g += NMS_SP_THREADS / 32) {
        const int2 gr = s_grp[g];
        float gx0 = FLT_MAX, gy0 = FLT_MAX, gx1 = -FLT_MAX, gy1 = -FLT_MAX;
        bool bad = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = h * 32 + lane;
            if (k < gr.y) {
                const float4 b = s_box[s_perm[gr.x + k]];
                if (b.w == 0.f) bad = true;
                else { gx0 = fminf(gx0, b.x - b.z); gx1 = fmaxf(gx1, b.x + b.z); gy0 = fminf(gy0, b.y - b.z); gy1 = fmaxf(gy1, b.y + b.z); }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            gx0 = fminf(gx0, __shfl_xor_sync(0xffffffffu, gx0, o)); gy0 = fminf(gy0, __shfl_xor_sync(0xffffffffu, gy0, o));
            gx1 = fmaxf(gx1, __shfl_xor_sync(0xffffffffu, gx1, o)); gy1 = fmaxf(gy1, __shfl_xor_sync(0xffffffffu, gy1, o));
        }
        if (__any_sync(0xffffffffu, bad)) { gx0 = gy0 = -CUDART_INF_F; gx1 = gy1 = CUDART_INF_F; }   // NaN anywhere: never culled
        if (lane == 0) { s_gbox[g] = make_float4(gx0, gy0, gx1, gy1); groups[g] = gr; }
    }
    for (int g = tid; g < ng; g += NMS_SP_THREADS) s_lab[g] = g;
    if (tid == 0) s_nedge = 0u;
    __syncthreads();
    // work items: group pairs (P <= Q) whose bounding boxes meet, one item per quarter of P's rows
    const int npairs = ng * (ng + 1) / 2;
    auto meet = [&](int P, int Q) -> bool {
        const float4 a = s_gbox[P], b = s_gbox[Q];
        return !(a.x > b.z || b.x > a.z || a.y > b.w || b.y > a.w);
    };
    for (int t = tid; t < npairs; t += NMS_SP_THREADS) {
        int P, Q;
        tri_decode(t, ng, P, Q);
        if (!meet(P, Q)) continue;
        const int nq = (s_grp[P].y + NMS_SP_QROWS - 1) / NMS_SP_QROWS;
        const unsigned int at = atomicAdd(tile_count, (unsigned int)nq);
        for (int q = 0; q < nq; ++q)
            tiles[at + q] = ((unsigned int)frame << 18) | ((unsigned int)P << 10) | ((unsigned int)Q << 2) | (unsigned int)q;
        if (P != Q) {
            atomicMin(&s_lab[Q], P);
            const unsigned int e = atomicAdd(&s_nedge, 1u);
            if (e < (unsigned int)NMS_SP_EDGES) s_edge[e] = ((unsigned int)P << 8) | (unsigned int)Q;
        }
    }
    // Connected components of the "group boxes meet" graph: a mask bit can only join boxes of one component, so the greedy
    // sweep runs per component, all components in parallel (nms_component_sweep_kernel).  Minimum-label propagation over the
    // edges + pointer jumping until no edge joins two labels; the label of a component is its lowest group number.
    for (;;) {
        __syncthreads();
        if (tid == 0) s_changed = 0;
        for (int g = tid; g < ng; g += NMS_SP_THREADS) { const int l = s_lab[g]; const int ll = s_lab[l]; if (ll < l) s_lab[g] = ll; }
        __syncthreads();
        auto join = [&](int P, int Q) {
            const int a = s_lab[P], b = s_lab[Q];
            if (a != b) { const int m = min(a, b); atomicMin(&s_lab[P], m); atomicMin(&s_lab[Q], m); s_changed = 1; }
        };
        const unsigned int ne = s_nedge;       // (final: written before the barrier above)
        if (ne <= (unsigned int)NMS_SP_EDGES) {
            for (unsigned int e = tid; e < ne; e += NMS_SP_THREADS) join((int)(s_edge[e] >> 8), (int)(s_edge[e] & 255u));
        } else {                               // more meeting pairs than the list holds: walk all pairs again
            for (int t = tid; t < npairs; t += NMS_SP_THREADS) {
                int P, Q;
                tri_decode(t, ng, P, Q);
                if (P != Q && meet(P, Q)) join(P, Q);
            }
        }
        __syncthreads();
        if (!s_changed) break;   // uniform
    }
    // labels (-1 beyond the frame's groups) and one membership bitmap per component, indexed by score position
    int* labels = labels_all + (size_t)frame * (groups_cap + 1);
    unsigned long long* members = members_all + (size_t)frame * groups_cap * col_blocks;   // zeroed by the launcher
    int roots = 0;
    for (int g = tid; g < groups_cap; g += NMS_SP_THREADS) {
        const int l = g < ng ? s_lab[g] : -1;
        labels[g] = l;
        roots += (l == g) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) roots += __shfl_xor_sync(0xffffffffu, roots, o);
    if (lane == 0 && roots) atomicAdd(&labels[groups_cap], roots);          // number of components (slot zeroed by the launcher)
    for (int g = warp; g < ng; g += NMS_SP_THREADS / 32) {
        const int2 gr = s_grp[g];
        unsigned long long* m = members + (size_t)s_lab[g] * col_blocks;
        for (int k = lane; k < gr.y; k += 32) {
            const int box = s_perm[gr.x + k];
            atomicOr(&m[box >> 6], 1ull << (box & 63));
        }
    }
}

// Greedy sweep per component: one warp per component label.  The warp holds the component's membership bits, the suppression
// bits and the kept bits in registers (up to 6 words of 64 boxes per lane); each iteration finds the lowest member that is not
// yet suppressed -- it is kept --, ORs its mask row into the suppression bits, and repeats: one iteration (one L2 round trip)
// per KEPT box of the component.  The last warp of a frame to finish turns the kept bits into the ascending index list.
constexpr int NMS_CS_WARPS = 8;
constexpr int NMS_CS_KMAX = (NMS_SP_MAX_N / NMS_TILE + 31) / 32;   // words per lane: at most 5
template <int NMS_CS_K>   // words of 64 boxes per lane: ceil(n / 2048) -- the loop is one dependent instruction chain per warp, so dead words cost
__global__ void __launch_bounds__(NMS_CS_WARPS * 32)
nms_component_sweep_kernel(const unsigned long long* __restrict__ mask_all, int n, int col_blocks, int frames, int groups_cap,
                           const int* __restrict__ labels_all, const unsigned long long* __restrict__ members_all,
                           unsigned long long* __restrict__ kept_all, unsigned int* __restrict__ done_all,
                           long long* __restrict__ keep_all, int* __restrict__ num_keep_all) {
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * NMS_CS_WARPS + (threadIdx.x >> 5);
    if (wid >= frames * groups_cap) return;
    const int frame = wid / groups_cap, g = wid - frame * groups_cap;
    const int* labels = labels_all + (size_t)frame * (groups_cap + 1);
    if (labels[g] != g) return;                                   // not the label of a component
    const unsigned long long* mask = mask_all + (size_t)frame * n * col_blocks;
    const unsigned long long* members = members_all + ((size_t)frame * groups_cap + g) * col_blocks;
    unsigned long long* keptw = kept_all + (size_t)frame * col_blocks;
    unsigned long long mem[NMS_CS_K], remv[NMS_CS_K], kept[NMS_CS_K];
#pragma unroll
    for (int k = 0; k < NMS_CS_K; ++k) {
        const int w = k * 32 + lane;
        mem[k] = w < col_blocks ? members[w] : 0ull;
        remv[k] = 0ull; kept[k] = 0ull;
    }
    constexpr int NB = 4;   // candidates whose rows travel together: one L2 round trip confirms up to four kept boxes
    for (;;) {
        // the NB lowest members not yet suppressed or kept (words ascend with k, then with the lane); the lowest one is kept for
        // sure, each further one unless a row ORed in before it suppresses it
        unsigned long long tmp[NMS_CS_K];
#pragma unroll
        for (int k = 0; k < NMS_CS_K; ++k) tmp[k] = mem[k] & ~remv[k];
        int ck[NB], cs[NB], cb[NB];
        int nc = 0;
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            int fk = -1;
            unsigned long long cw = 0ull;
            unsigned int bal = 0u;
#pragma unroll
            for (int k = 0; k < NMS_CS_K; ++k) {
                if (fk < 0) {
                    const unsigned int b = __ballot_sync(0xffffffffu, tmp[k] != 0ull);
                    if (b) { fk = k; cw = tmp[k]; bal = b; }
                }
            }
            ck[c] = fk; cs[c] = 0; cb[c] = 0;
            if (fk >= 0) {
                cs[c] = __ffs((int)bal) - 1;
                cb[c] = __ffsll((long long)__shfl_sync(0xffffffffu, cw, cs[c])) - 1;
                nc = c + 1;
#pragma unroll
                for (int k = 0; k < NMS_CS_K; ++k) if (k == fk && lane == cs[c]) tmp[k] &= ~(1ull << cb[c]);
            }
        }
        if (nc == 0) break;
        unsigned long long rows[NB][NMS_CS_K];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            const int i = ((ck[c] * 32 + cs[c]) << 6) + cb[c];
            const unsigned long long* row = mask + (size_t)(c < nc ? i : 0) * col_blocks;
#pragma unroll
            for (int k = 0; k < NMS_CS_K; ++k) {
                const int w = k * 32 + lane;
                rows[c][k] = (c < nc && w < col_blocks) ? __ldcg(row + w) : 0ull;     // written by the mask kernels' atomics: read through L2
            }
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            if (c < nc) {
                // still unsuppressed?  (the owner lane's word, broadcast; candidate 0 always is)
                unsigned long long ow = 0ull;
#pragma unroll
                for (int k = 0; k < NMS_CS_K; ++k) if (k == ck[c]) ow = remv[k];
                ow = __shfl_sync(0xffffffffu, ow, cs[c]);
                if (!((ow >> cb[c]) & 1ull)) {
#pragma unroll
                    for (int k = 0; k < NMS_CS_K; ++k) {
                        unsigned long long r = rows[c][k];
                        if (k == ck[c] && lane == cs[c]) { r |= 1ull << cb[c]; kept[k] |= 1ull << cb[c]; }
                        remv[k] |= r;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NMS_CS_K; ++k) {
        const int w = k * 32 + lane;
        if (w < col_blocks && kept[k]) atomicOr(&keptw[w], kept[k]);
    }
    __threadfence();
    __syncwarp();
    unsigned int last = 0u;
    if (lane == 0) last = (atomicAdd(&done_all[frame], 1u) == (unsigned int)labels[groups_cap] - 1u) ? 1u : 0u;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    // kept bits -> ascending indices
    long long* keep = keep_all + (size_t)frame * n;
    int base = 0;
#pragma unroll
    for (int k = 0; k < NMS_CS_K; ++k) {
        const int w = k * 32 + lane;
        unsigned long long kw = w < col_blocks ? __ldcg(keptw + w) : 0ull;
        const int c = __popcll(kw);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        int pos = base + incl - c;
        while (kw) { const int b = __ffsll((long long)kw) - 1; kw &= kw - 1; keep[pos++] = (long long)(w * 64 + b); }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) num_keep_all[frame] = base;
}

__global__ void __launch_bounds__(NMS_THREADS, GLENET_NMS_CTAS)
nms_mask_spatial_kernel(const float* __restrict__ boxes_all, int n, float thresh, unsigned long long* __restrict__ mask_all, int col_blocks,
                        const int* __restrict__ perm_all, const int2* __restrict__ groups_all, int groups_cap,
                        const unsigned int* __restrict__ tiles, unsigned int* tile_count,   // [0] number of items, [1] next item to hand out
                        unsigned long long* __restrict__ list_count, unsigned long long* __restrict__ list, unsigned long long list_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem& sm = *reinterpret_cast<NmsSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned int ntiles = tile_count[0];
    unsigned int* next_item = tile_count + 1;                                // zeroed by the launcher with the count
    const float thr_lo = thresh * (1.f - 1e-3f) - 1e-5f, thr_hi = thresh * (1.f + 1e-3f) + 1e-5f;
    // Work items differ by two orders of magnitude (an object's own groups vs. two groups whose boxes merely touch), so they are
    // handed out dynamically: thread 0 claims the NEXT item when the current one starts (the atomic's latency hides behind the
    // item) and publishes it at the end.
    __shared__ unsigned int s_next;
    unsigned int nxt = 0u;
    auto fetch_next = [&]() -> unsigned int {
        __syncthreads();                       // everybody is done with the item (and with s_next)
        if (tid == 0) s_next = nxt;
        __syncthreads();
        return s_next;
    };
    for (unsigned int t = blockIdx.x; t < ntiles; t = fetch_next()) {
        if (tid == 0) nxt = gridDim.x + atomicAdd(next_item, 1u);
        const unsigned int tw = tiles[t];
        const int frame = (int)(tw >> 18), P = (int)((tw >> 10) & 255u), Q = (int)((tw >> 2) & 255u), rbase = (int)(tw & 3u) * NMS_SP_QROWS;
        const float* boxes = boxes_all + (size_t)frame * n * 7;
        const int* perm = perm_all + (size_t)frame * n;
        unsigned long long* mask = mask_all + (size_t)frame * n * col_blocks;
        const int2 gp = groups_all[(size_t)frame * groups_cap + P], gq = groups_all[(size_t)frame * groups_cap + Q];
        const int tr = gp.y, tc = gq.y;                                  // boxes in the two groups (<= 64)
        const int rend = min(tr, rbase + NMS_SP_QROWS);                  // this item: rows [rbase, rend) of P against all of Q
        for (int i = tid; i < (rend - rbase) + tc; i += NMS_THREADS) {
            const bool is_row = i < rend - rbase;
            const int k = is_row ? rbase + i : i - (rend - rbase);
            const int id = perm[(is_row ? gp.x : gq.x) + k];
            const float* box = boxes + (size_t)id * 7;
            const float cx = box[0], cy = box[1], rad = cull_radius(box);
            if (is_row) { sm.rcx[k] = cx; sm.rcy[k] = cy; sm.rrad[k] = rad; sm.rflag[k] = 0; sm.idx[k] = id; }
            else        { sm.ccx[k] = cx; sm.ccy[k] = cy; sm.crad[k] = rad; sm.cflag[k] = 0; sm.idx[NMS_TILE + k] = id; }
        }
        if (tid == 0) { sm.qcount = 0; sm.q2count = 0; }
        __syncthreads();
        // cull pass: every unordered pair once (r < c inside a group); thread = one column and 4 of the item's 16 rows
        const bool diag = P == Q;
        constexpr int RPT = NMS_SP_QROWS * NMS_TILE / NMS_THREADS;       // 4
        const int c = tid & (NMS_TILE - 1), rq = tid >> 6;
        unsigned int hits = 0u;
        if (c < tc) {
            const float ccx = sm.ccx[c], ccy = sm.ccy[c], ccr = sm.crad[c];
            const int lim = diag ? min(rend, c) : rend;
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int r = rbase + k * (NMS_THREADS / NMS_TILE) + rq;
                if (r < lim) {
                    const float ddx = sm.rcx[r] - ccx, ddy = sm.rcy[r] - ccy, rr = sm.rrad[r] + ccr;
                    hits |= (!(ddx * ddx + ddy * ddy > rr * rr) ? 1u : 0u) << k;     // NaN anywhere: not culled
                }
            }
        }
        {
            const int cnt = __popc(hits);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
            const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
            if (wtotal) {
                int qb = 0;
                if (lane == 31) qb = atomicAdd(&sm.qcount, wtotal);
                qb = __shfl_sync(0xffffffffu, qb, 31) + incl - cnt;
                if (hits && sm.cflag[c] == 0) sm.cflag[c] = 1;
                while (hits) {
                    const int k = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const int r = rbase + k * (NMS_THREADS / NMS_TILE) + rq;
                    sm.queue[qb++] = (unsigned short)((r << 6) | c);
                    if (sm.rflag[r] == 0) sm.rflag[r] = 1;
                }
            }
        }
        __syncthreads();
        const int nq = sm.qcount;
        if (nq == 0) continue;   // uniform
        for (int i = tid; i < (rend - rbase) + tc; i += NMS_THREADS) {
            const bool is_row = i < rend - rbase;
            const int k = is_row ? rbase + i : i - (rend - rbase);
            if ((is_row ? sm.rflag[k] : sm.cflag[k]) == 1) {
                const float* box = boxes + (size_t)sm.idx[is_row ? k : NMS_TILE + k] * 7;
                box_prepare<true, false>(box, device_trig(box[6]), (is_row ? sm.rpre : sm.cpre) + k * NBS);
            }
        }
        __syncthreads();
        // bound filter + approximate overlap, as in nms_mask_kernel; box_a = the box with the lower score index
        for (int q0 = 0; q0 < nq; q0 += NMS_THREADS) {
            const int q = q0 + tid;
            int p = 0;
            bool need = false;
            if (q < nq) {
                p = sm.queue[q];
                const int ri = sm.idx[p >> 6], ci = sm.idx[NMS_TILE + (p & 63)];
                const bool swap = ri > ci;
                const float* a = swap ? sm.cpre + (p & 63) * NBS : sm.rpre + (p >> 6) * NBS;
                const float* b = swap ? sm.rpre + (p >> 6) * NBS : sm.cpre + (p & 63) * NBS;
                if (swap) p |= 1 << 12;
                const float ub = overlap_upper_bound(a, b);
                const float iou_ub = ub / fmaxf(a[BP_AREA] + b[BP_AREA] - ub, 1e-8f);
                need = (!(ub <= 0.f) && !(iou_ub <= thr_lo)) || !(a[BP_AREA] + b[BP_AREA] > ub);
#if GLENET_NMS_APPROX
                if (need && overlap_approx_usable(a, b)) {
                    const float ova = overlap_approx(a, b);
                    float slack, band;
                    overlap_approx_band(a, b, slack, band);
                    const float s = a[BP_AREA] + b[BP_AREA];
                    const float hi = ova + band + slack, lo = ova - slack;
                    if (hi / fmaxf(s - hi, 1e-8f) <= thr_lo && s > hi) need = false;
                    else if (lo > 0.f && s > lo && lo / fmaxf(s - lo, 1e-8f) > thr_hi) {
                        need = false;
                        const int i = min(ri, ci), j = max(ri, ci);
                        atomicOr(reinterpret_cast<unsigned int*>(mask + (size_t)i * col_blocks + (j >> 6)) + ((j >> 5) & 1), 1u << (j & 31));
                    }
                }
#endif
            }
            const unsigned int m = __ballot_sync(0xffffffffu, need);
            if (m) {
                int qb = 0;
                if (lane == 0) qb = atomicAdd(&sm.q2count, __popc(m));
                qb = __shfl_sync(0xffffffffu, qb, 0);
                if (need) sm.queue2[qb + __popc(m & ((1u << lane) - 1))] = (unsigned short)p;
            }
        }
        __syncthreads();
        const int nq2 = sm.q2count;
        if (nq2 == 0) continue;   // uniform
        if (list) {
            if (tid == 0) {
                const unsigned long long base = atomicAdd(list_count, (unsigned long long)nq2);
                sm.qcount = (base + (unsigned long long)nq2 <= list_cap) ? 1 : 0;
                reinterpret_cast<unsigned long long*>(sm.wl[0])[0] = base;
            }
            __syncthreads();
            const unsigned long long base = reinterpret_cast<unsigned long long*>(sm.wl[0])[0];
            if (sm.qcount) {
                for (int q = tid; q < nq2; q += NMS_THREADS) {
                    const int p = sm.queue2[q];
                    const int ri = sm.idx[(p >> 6) & 63], ci = sm.idx[NMS_TILE + (p & 63)];
                    list[base + q] = ((unsigned long long)frame << 40) | ((unsigned long long)min(ri, ci) << 20) | (unsigned long long)max(ri, ci);
                }
                continue;   // uniform
            }
            for (int q = tid; q < nq2; q += NMS_THREADS) if (base + q < list_cap) list[base + q] = ~0ull;   // refused: void entries
            __syncthreads();
        }
        nms_tile_clip<true>(sm, nq2, thresh, mask, col_blocks);
    }
}

// Exact clip of the pairs the mask kernel could not decide (nms_mask_kernel, GLENET_NMS_DEFER): one pair per lane, both
// BoxPre records in shared memory, the phased clip of clip.cuh, the mask bit set with a 32-bit atomicOr.  Persistent CTAs walk
// the list; its length is read from device memory (no host synchronisation).  Row box = box_a, column box = box_b, prepared
// with the same calls as in the tile (iou3d_nms_kernel.cu:267-311).
constexpr int NCL_THREADS = 128;
struct NclSmem {
    float2 verts[NCL_THREADS * CLIP_SLOTS];
    unsigned int wl[NCL_THREADS / 32][32 * CLIP_SLOTS];
    float arec[NCL_THREADS * NBS], brec[NCL_THREADS * NBS];
};
__global__ void __launch_bounds__(NCL_THREADS, 4)
nms_clip_list_kernel(const float* __restrict__ boxes_all, int n, float thresh, unsigned long long* __restrict__ mask_all, int col_blocks,
                     const unsigned long long* __restrict__ list_count, const unsigned long long* __restrict__ list, unsigned long long list_cap) {
    __shared__ NclSmem sm;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned long long total = *list_count;
    if (total > list_cap) total = list_cap;   // reservations beyond the capacity were refused (their tiles clipped locally); refused slots below it hold ~0
    for (unsigned long long base = (unsigned long long)blockIdx.x * NCL_THREADS; base < total; base += (unsigned long long)gridDim.x * NCL_THREADS) {
        bool live = base + tid < total;
        int frame = 0, i = 0, j = 0;
        if (live) {
            const unsigned long long e = list[base + tid];
            live = e != ~0ull;
            frame = (int)(e >> 40); i = (int)((e >> 20) & 0xfffffu); j = (int)(e & 0xfffffu);
        }
        float* a = sm.arec + tid * NBS;
        float* b = sm.brec + tid * NBS;
        if (live) {
            const float* ba = boxes_all + ((size_t)frame * n + i) * 7;
            const float* bb = boxes_all + ((size_t)frame * n + j) * 7;
            box_prepare<true, false>(ba, device_trig(ba[6]), a);
            box_prepare<true, false>(bb, device_trig(bb[6]), b);
        }
        __syncwarp();
        float2* slots = sm.verts + tid * CLIP_SLOTS;
        const unsigned int w = clip_pair_tests<true>(a, b, live);
        const unsigned int hits = clip_hits16(w);
        const int cnt = __popc(hits) + __popc(clip_corners8(w));
        const bool fast = cnt >= 3 && cnt <= CLIP_SLOTS;
        if (fast) clip_write_corners(a, b, w, slots);
        clip_warp_points<true>(fast ? hits : 0u, (unsigned int)tid, (unsigned int)tid, sm.wl[warp], sm.arec, sm.brec, NBS, sm.verts + (warp * 32) * CLIP_SLOTS);
        const bool slow = cnt > CLIP_SLOTS;
        const float ov_slow = clip_warp_slow<true>(slow, w, (unsigned int)tid, (unsigned int)tid, sm.arec, sm.brec, NBS, reinterpret_cast<float2*>(sm.wl[warp]));
        if (live) {
            const float ov = slow ? ov_slow : (fast ? clip_area8<true>(slots, cnt) : 0.f);
            if (iou_from_overlap(a[BP_AREA], b[BP_AREA], ov) > thresh) {
                unsigned long long* word = mask_all + ((size_t)frame * n + i) * col_blocks + (j >> 6);
                atomicOr(reinterpret_cast<unsigned int*>(word) + ((j >> 5) & 1), 1u << (j & 31));
            }
        }
        __syncwarp();   // the records and slots of this batch are dead before the next one overwrites them
    }
}

// Greedy sweep of iou3d_nms.cpp:116-132 on the device.  One CTA per frame, one barrier per 64-box chunk.
// The serial dependency (chunk c+1 needs the kept set of chunk c) is kept off the memory system:
//   * the diagonal words mask[r][r/64] and the super-diagonal words mask[r][r/64 + 1] of ALL rows are
//     staged in shared memory up front (PRE = true; 16 bytes per box), so resolving a chunk and folding
//     its kept rows into the next chunk's suppression word touches shared memory only;
//   * the kept rows' words for chunks >= c+2 are ORed in by the other warps ONE STEP LATER (they are not
//     needed before), overlapping their L2 latency with warp 0's next resolve.
// For very large n (PRE = false) the two words are read from global memory instead.
// (A barrier-free variant with 3 staged super-diagonals and flag-synchronised job warps was measured and
//  lost to this one: the spinning warps cost more than the barrier.)
template <bool PRE>
__global__ void __launch_bounds__(SWEEP_THREADS)
nms_sweep_kernel(const unsigned long long* __restrict__ mask_all, int n, int col_blocks,
                 long long* __restrict__ keep_all, int* __restrict__ num_keep_all) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* remv = reinterpret_cast<unsigned long long*>(smem_raw);   // [col_blocks]
    unsigned long long* s_diag = remv + col_blocks;                                 // [n]  (PRE)
    unsigned long long* s_sup = s_diag + (PRE ? n : 0);                             // [n]  (PRE)
    __shared__ int s_rows[2][NMS_TILE];
    __shared__ int s_nrows[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x;
    const unsigned long long* mask = mask_all + (size_t)frame * n * col_blocks;
    long long* keep = keep_all + (size_t)frame * n;

    for (int j = tid; j < col_blocks; j += SWEEP_THREADS) remv[j] = 0ull;
    if (PRE) {
        for (int r = tid; r < n; r += SWEEP_THREADS) {
            const int c = r >> 6;
            s_diag[r] = mask[(size_t)r * col_blocks + c];
            s_sup[r] = (c + 1 < col_blocks) ? mask[(size_t)r * col_blocks + c + 1] : 0ull;
        }
    }
    if (tid < 2) s_nrows[tid] = 0;
    int num_keep = 0;   // tracked by warp 0
    __syncthreads();

#ifdef GLENET_PHASE_TIMING
    const long long t_loop0 = clock64();
    long long busy = 0;
#endif
    for (int c = 0; c < col_blocks; ++c) {
        const int rows = min(NMS_TILE, n - c * NMS_TILE);
#ifdef GLENET_PHASE_TIMING
        const long long t_step0 = clock64();
#endif
        if (warp == 0) {
            // resolve the diagonal tile: one iteration per KEPT box (find-first-set over the unsuppressed bits)
            const int r_lo = c * NMS_TILE + lane, r_hi = r_lo + 32;
            const unsigned long long d_lo = (lane < rows) ? (PRE ? s_diag[r_lo] : mask[(size_t)r_lo * col_blocks + c]) : 0ull;
            const unsigned long long d_hi = (lane + 32 < rows) ? (PRE ? s_diag[r_hi] : mask[(size_t)r_hi * col_blocks + c]) : 0ull;
            const bool has_next = c + 1 < col_blocks;
            const unsigned long long u_lo = (has_next && lane < rows) ? (PRE ? s_sup[r_lo] : mask[(size_t)r_lo * col_blocks + c + 1]) : 0ull;
            const unsigned long long u_hi = (has_next && lane + 32 < rows) ? (PRE ? s_sup[r_hi] : mask[(size_t)r_hi * col_blocks + c + 1]) : 0ull;
            const unsigned long long valid = (rows == NMS_TILE) ? ~0ull : ((1ull << rows) - 1ull);
            // The serial chain (one iteration per kept box) runs on 32-bit halves: rows 0..31 first -- a kept row ORs both
            // halves of its diagonal word into the suppression word --, then rows 32..63, whose words only have bits in the
            // upper half (the mask holds j > i only).
            const unsigned long long w0 = remv[c];
            unsigned int wl = (unsigned int)w0, wh = (unsigned int)(w0 >> 32);
            const unsigned int vl = (unsigned int)valid, vh = (unsigned int)(valid >> 32);
            const unsigned int dll = (unsigned int)d_lo, dlh = (unsigned int)(d_lo >> 32), dhh = (unsigned int)(d_hi >> 32);
            unsigned int kl = 0u, kh = 0u;
            for (unsigned int cand = ~wl & vl; cand;) {
                const int i = __ffs((int)cand) - 1;
                kl |= 1u << i;
                wl |= __shfl_sync(0xffffffffu, dll, i);
                wh |= __shfl_sync(0xffffffffu, dlh, i);
                cand = ~wl & vl & ~((2u << i) - 1u);
            }
            for (unsigned int cand = ~wh & vh; cand;) {
                const int i = __ffs((int)cand) - 1;
                kh |= 1u << i;
                wh |= __shfl_sync(0xffffffffu, dhh, i);
                cand = ~wh & vh & ~((2u << i) - 1u);
            }
            const unsigned long long kept = (unsigned long long)kl | ((unsigned long long)kh << 32);
            // kept rows of this chunk -> next chunk's suppression word, straight from registers
            const unsigned long long nxl = (((kept >> lane) & 1ull) ? u_lo : 0ull) | (((kept >> (lane + 32)) & 1ull) ? u_hi : 0ull);
            const unsigned long long nx = (unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned int)nxl) |
                                          ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned int)(nxl >> 32)) << 32);
            if (lane == 0 && has_next && nx) atomicOr(&remv[c + 1], nx);
            // emit kept indices in ascending order; remember the rows for the lagging propagation
            const int slot = c & 1;
            if (kept) for (int i = lane; i < NMS_TILE; i += 32) {
                if ((kept >> i) & 1ull) {
                    const int pos = __popcll(kept & ((1ull << i) - 1ull));
                    keep[num_keep + pos] = (long long)(c * NMS_TILE + i);
                    s_rows[slot][pos] = c * NMS_TILE + i;
                }
            }
            num_keep += __popcll(kept);
            if (lane == 0) s_nrows[slot] = __popcll(kept);
        } else if (c > 0) {
            // lagging propagation: rows kept in chunk c-1 -> suppression words of chunks >= c+1
            const int slot = (c - 1) & 1;
            const int first = c + 1, nrem = col_blocks - first;
            const int items = s_nrows[slot] * nrem;
            for (int it = tid - 32; it < items; it += SWEEP_THREADS - 32) {
                const int ki = it / nrem, j = first + (it - ki * nrem);
                const unsigned long long m = mask[(size_t)s_rows[slot][ki] * col_blocks + j];
                if (m) atomicOr(&remv[j], m);
            }
        }
#ifdef GLENET_PHASE_TIMING
        busy += clock64() - t_step0;
#endif
        __syncthreads();
    }
#ifdef GLENET_PHASE_TIMING
    if (frame == 0 && lane == 0) {
        if (warp == 0) { atomicAdd(&g_sweep_cycles[0], (unsigned long long)busy); atomicAdd(&g_sweep_cycles[1], (unsigned long long)(clock64() - t_loop0)); atomicAdd(&g_sweep_cycles[3], (unsigned long long)col_blocks); }
        if (warp == 1) atomicAdd(&g_sweep_cycles[2], (unsigned long long)busy);
    }
#endif
    if (tid == 0) num_keep_all[frame] = num_keep;
}

// workspace: [mask][16-byte list counter][deferred-clip list][spatial: permutation | group table | work items | item counter]
struct NmsWorkspace {
    size_t mask_bytes, list_off, perm_off, groups_off, tiles_off, tcount_off, zero_off, zero_bytes, labels_off, done_off, kept_off, members_off, bytes;
    unsigned long long list_cap;
    size_t tiles_cap;
    int groups_cap;
};
static NmsWorkspace nms_layout(int frames, int n) {
    NmsWorkspace w;
    const size_t col_blocks = ((size_t)n + NMS_TILE - 1) / NMS_TILE;
    w.mask_bytes = align_up((size_t)frames * n * col_blocks * sizeof(unsigned long long), 16);
    w.list_off = w.mask_bytes;                                  // 16-byte counter, then the entries
    w.list_cap = nms_list_cap(frames, n);
    size_t off = w.list_off + 16 + (size_t)w.list_cap * sizeof(unsigned long long);
    off = align_up(off, 16);
    const bool sp = n <= NMS_SP_MAX_N;                          // larger n never takes the spatial path
    w.groups_cap = sp ? (int)(col_blocks + NMS_SP_SUPER < (size_t)NMS_SP_MAX_GROUPS ? col_blocks + NMS_SP_SUPER : (size_t)NMS_SP_MAX_GROUPS) : 0;
    w.perm_off = off;   off += sp ? align_up((size_t)frames * n * sizeof(int), 16) : 0;
    w.groups_off = off; off += align_up((size_t)frames * w.groups_cap * sizeof(int2), 16);
    w.tiles_cap = (size_t)frames * ((size_t)w.groups_cap * (w.groups_cap + 1) / 2) * (NMS_TILE / NMS_SP_QROWS);
    w.tiles_off = off;  off += align_up(w.tiles_cap * sizeof(unsigned int), 16);
    w.tcount_off = off; off += 16;
    // zeroed per call: component labels (+ count), per-frame done counters, kept bits, membership bitmaps
    w.zero_off = off;
    w.labels_off = off;  off += align_up((size_t)frames * (w.groups_cap + 1) * sizeof(int), 16);
    w.done_off = off;    off += align_up((size_t)frames * sizeof(unsigned int), 16);
    w.kept_off = off;    off += sp ? align_up((size_t)frames * col_blocks * sizeof(unsigned long long), 16) : 0;
    w.members_off = off; off += align_up((size_t)frames * w.groups_cap * col_blocks * sizeof(unsigned long long), 16);
    w.zero_bytes = off - w.zero_off;
    w.bytes = off;
    return w;
}

static int launch_nms(bool normal, const float* boxes, int frames, int n, float thresh, int64_t* keep,
                      int32_t* num_keep, void* ws, size_t ws_bytes, cudaStream_t stream, const char* what) {
    if (frames < 0 || n < 0) return fail(GLENET_EINVAL, "%s: negative size", what);
    if (frames == 0) return GLENET_OK;
    if (!num_keep) return fail(GLENET_EINVAL, "%s: null num_keep", what);
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(num_keep, 0, sizeof(int32_t) * frames, stream);
        return e == cudaSuccess ? GLENET_OK : fail(-(int)e, "%s: memset failed", what);
    }
    if (!boxes || !keep) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (ws_bytes < glenet_nms_workspace_bytes(frames, n) || !ws) return fail(GLENET_EWORKSPACE, "%s: workspace too small", what);
    if ((uintptr_t)ws & 15) return fail(GLENET_EALIGN, "%s: workspace must be 16-byte aligned", what);
    const int col_blocks = (n + NMS_TILE - 1) / NMS_TILE;
    const long tiles = (long)col_blocks * (col_blocks + 1) / 2;
    if (tiles * frames > 0x7fffffffL) return fail(GLENET_EINVAL, "%s: too many tiles", what);
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws);
    // opt-in shared-memory sizes are per-device settings: remember where they have been set
    static std::atomic<unsigned char> attr_done[GLENET_MAX_DEVICES];   // bit 0: mask kernels, bit 1 / 2: sweep kernel <false> / <true>, bit 3: spatial mask kernel
    int dev = 0;
    cudaGetDevice(&dev);
    const bool cacheable = dev >= 0 && dev < GLENET_MAX_DEVICES;
    unsigned char done = cacheable ? attr_done[dev].load(std::memory_order_acquire) : 0;
    if (!(done & 1)) {
        int rc = set_smem(nms_mask_kernel<false>, sizeof(NmsSmem), what);
        if (rc) return rc;
        rc = set_smem(nms_mask_kernel<true>, sizeof(NmsSmem), what);
        if (rc) return rc;
        if (cacheable) attr_done[dev].fetch_or(1, std::memory_order_release);
    }
    const unsigned grid = (unsigned)(tiles * frames);
    const NmsWorkspace w = nms_layout(frames, n);
    // deferred exact clips: [count][list] behind the mask
    unsigned long long* list_count = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(ws) + w.list_off);
    unsigned long long* list = list_count + 2;
    const unsigned long long list_cap = w.list_cap;
    const bool defer = GLENET_NMS_DEFER && !normal && n < (1 << 20) && frames < (1 << 23);
    if (defer) {
        cudaError_t e = cudaMemsetAsync(list_count, 0, 16, stream);
        if (e != cudaSuccess) return fail(-(int)e, "%s: memset failed", what);
    }
    // spatial tiles (see nms_spatial_kernel): worth it once the n^2 / 2 circle tests dominate; a negative threshold keeps
    // every pair (one global atomic each) and stays with the score-order tiles.  GLENET_NMS_SPATIAL=0 turns them off, =2 forces them.
    static const int spatial_mode = [] { const char* e = getenv("GLENET_NMS_SPATIAL"); return e ? atoi(e) : 1; }();   // 0 off, 1 by size, 2 whenever legal (tests)
    const bool spatial_on = spatial_mode != 0;
    const bool spatial = spatial_on && !normal && thresh >= 0.f && n <= NMS_SP_MAX_N && frames < (1 << 14) && (spatial_mode == 2 || n >= 1024);
    if (spatial) {
        unsigned char* base = reinterpret_cast<unsigned char*>(ws);
        int* perm = reinterpret_cast<int*>(base + w.perm_off);
        int2* groups = reinterpret_cast<int2*>(base + w.groups_off);
        unsigned int* tile_list = reinterpret_cast<unsigned int*>(base + w.tiles_off);
        unsigned int* tile_count = reinterpret_cast<unsigned int*>(base + w.tcount_off);
        int* labels = reinterpret_cast<int*>(base + w.labels_off);
        unsigned long long* members = reinterpret_cast<unsigned long long*>(base + w.members_off);
        cudaError_t e = cudaMemsetAsync(tile_count, 0, 16 + w.zero_bytes, stream);      // counter + labels, done counters, kept bits, membership bitmaps
        if (e == cudaSuccess) e = cudaMemsetAsync(mask, 0, (size_t)frames * n * col_blocks * sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return fail(-(int)e, "%s: memset failed", what);
        int rc;
        if (!(done & 8)) {
            rc = set_smem(nms_mask_spatial_kernel, sizeof(NmsSmem), what);
            if (rc) return rc;
            rc = set_smem(nms_spatial_kernel, nms_spatial_smem(NMS_SP_MAX_N), what);
            if (rc) return rc;
            if (cacheable) attr_done[dev].fetch_or(8, std::memory_order_release);
        }
        nms_spatial_kernel<<<frames, NMS_SP_THREADS, nms_spatial_smem(n), stream>>>(boxes, n, perm, groups, w.groups_cap, tile_list, tile_count,
                                                                                    labels, members, col_blocks);
        rc = check_launch(what);
        if (rc) return rc;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t most = w.tiles_cap < (size_t)sms * GLENET_NMS_CTAS ? w.tiles_cap : (size_t)sms * GLENET_NMS_CTAS;
        nms_mask_spatial_kernel<<<(unsigned)most, NMS_THREADS, sizeof(NmsSmem), stream>>>(boxes, n, thresh, mask, col_blocks, perm, groups, w.groups_cap,
                                                                                         tile_list, tile_count,
                                                                                         defer ? list_count : nullptr, defer ? list : nullptr, list_cap);
    } else if (normal)
        nms_mask_kernel<true><<<grid, NMS_THREADS, sizeof(NmsSmem), stream>>>(boxes, n, thresh, mask, col_blocks, (int)tiles, nullptr, nullptr, 0ull);
    else
        nms_mask_kernel<false><<<grid, NMS_THREADS, sizeof(NmsSmem), stream>>>(boxes, n, thresh, mask, col_blocks, (int)tiles,
                                                                              defer ? list_count : nullptr, defer ? list : nullptr, list_cap);
    int rc = check_launch(what);
    if (rc) return rc;
    if (defer) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        nms_clip_list_kernel<<<sms * 4, NCL_THREADS, 0, stream>>>(boxes, n, thresh, mask, col_blocks, list_count, list, list_cap);
        rc = check_launch(what);
        if (rc) return rc;
    }
    if (spatial) {
        const NmsWorkspace w2 = nms_layout(frames, n);
        unsigned char* base = reinterpret_cast<unsigned char*>(ws);
        const long warps = (long)frames * w2.groups_cap;
        const unsigned cs_grid = (unsigned)((warps + NMS_CS_WARPS - 1) / NMS_CS_WARPS);
        const int* labels = reinterpret_cast<const int*>(base + w2.labels_off);
        const unsigned long long* members = reinterpret_cast<const unsigned long long*>(base + w2.members_off);
        unsigned long long* kept_bits = reinterpret_cast<unsigned long long*>(base + w2.kept_off);
        unsigned int* done_cnt = reinterpret_cast<unsigned int*>(base + w2.done_off);
        long long* keep_ll = reinterpret_cast<long long*>(keep);
        static_assert(NMS_CS_KMAX == 5, "instantiations below");
#define GLENET_CS_LAUNCH(K) nms_component_sweep_kernel<K><<<cs_grid, NMS_CS_WARPS * 32, 0, stream>>>(mask, n, col_blocks, frames, w2.groups_cap, labels, members, kept_bits, done_cnt, keep_ll, num_keep)
        switch ((col_blocks + 31) / 32) {
            case 1: GLENET_CS_LAUNCH(1); break;
            case 2: GLENET_CS_LAUNCH(2); break;
            case 3: GLENET_CS_LAUNCH(3); break;
            case 4: GLENET_CS_LAUNCH(4); break;
            default: GLENET_CS_LAUNCH(5); break;
        }
#undef GLENET_CS_LAUNCH
        return check_launch(what);
    }
    const size_t remv_bytes = sizeof(unsigned long long) * col_blocks;
    const size_t pre_bytes = remv_bytes + 2 * sizeof(unsigned long long) * (size_t)n;
    if (remv_bytes > 200 * 1024) return fail(GLENET_EINVAL, "%s: n too large for the on-chip suppression words", what);
    const bool pre = pre_bytes <= 200 * 1024;   // n <= ~12 700 boxes
    const size_t sweep_smem = pre ? pre_bytes : remv_bytes;
    const unsigned char sweep_bit = pre ? 4 : 2;
    if (!(done & sweep_bit)) {
        rc = pre ? set_smem(nms_sweep_kernel<true>, 200 * 1024, what) : set_smem(nms_sweep_kernel<false>, 200 * 1024, what);
        if (rc) return rc;
        if (cacheable) attr_done[dev].fetch_or(sweep_bit, std::memory_order_release);
    }
    if (pre)
        nms_sweep_kernel<true><<<frames, SWEEP_THREADS, sweep_smem, stream>>>(mask, n, col_blocks, reinterpret_cast<long long*>(keep), num_keep);
    else
        nms_sweep_kernel<false><<<frames, SWEEP_THREADS, sweep_smem, stream>>>(mask, n, col_blocks, reinterpret_cast<long long*>(keep), num_keep);
    return check_launch(what);
}

}  // namespace glenet

using namespace glenet;

extern "C" {

size_t glenet_nms_workspace_bytes(int frames, int n) {
    if (frames <= 0 || n <= 0) return 16;
    const size_t col_blocks = ((size_t)n + NMS_TILE - 1) / NMS_TILE;
    (void)col_blocks;
    return nms_layout(frames, n).bytes;
}

#ifdef GLENET_PHASE_TIMING
int glenet_debug_sweep_cycles(unsigned long long* host_out4) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out4, g_sweep_cycles, sizeof(unsigned long long) * 4);
    unsigned long long z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_sweep_cycles, z, sizeof(z));
    return 0;
}
#endif

int glenet_nms_gpu(const float* boxes, int frames, int n, float thresh, int64_t* keep, int32_t* num_keep, void* ws,
                   size_t ws_bytes, glenet_stream_t s) {
    return launch_nms(false, boxes, frames, n, thresh, keep, num_keep, ws, ws_bytes, (cudaStream_t)s, "glenet_nms_gpu");
}
int glenet_nms_normal_gpu(const float* boxes, int frames, int n, float thresh, int64_t* keep, int32_t* num_keep,
                          void* ws, size_t ws_bytes, glenet_stream_t s) {
    return launch_nms(true, boxes, frames, n, thresh, keep, num_keep, ws, ws_bytes, (cudaStream_t)s, "glenet_nms_normal_gpu");
}

}  // extern "C"
