// Pairwise rotated-box overlap / BEV IoU / fused 3D IoU for sm_100a.
//
// Replaces boxes_overlap_kernel / boxes_iou_bev_kernel (+ launchers) of
// pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:236-265,378-398 and the ~10 torch
// elementwise kernels of boxes_iou3d_gpu (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:88-121).
//
// Design (one CTA = one TR x TC tile of the (na, nb) matrix; the grid is sized to ONE wave of resident
// CTAs whenever the problem allows it, because the kernel is bound by the latency of a CTA's phase chain):
//   0. zero fill   : > 99 % of an anchor sweep is exactly +0.0 and that regime is HBM-write bound
//                    (4 B / pair).  The tile is zero-filled by the bulk-copy engine (cp.async.bulk
//                    shared -> global from a 4 KB block of zeros): a handful of instructions per CTA,
//                    fully asynchronous, so the stores drain while the CTA works through 1-4.
//   1. cull pass   : two-level exact-conservative circle test on box tiles staged in shared
//                    memory.  First one test per COLUMN against the bounding box of the tile's
//                    row centres (an anchor tile sees ~4 of 100 GT boxes), then one test per
//                    (row, active column), 32 columns at a time into a register bitmask.
//   2. compaction  : surviving pairs are appended to a shared-memory queue (one atomic per warp and
//                    32 columns), their boxes are flagged.
//   3. lazy prepare: only flagged boxes get their BoxPre record (4 trig calls, corners,
//                    margin thresholds) -- once per box per tile, never per pair, on full warps.
//   4. SAT filter  : a separating-axis test on the prepared records drops the pairs whose overlap is
//                    exactly 0 in the reference too; the rest is compacted again.
//   5. clip pass   : one pair per thread, results parked in shared memory and written once the
//                    zero fill has landed.
// The reference instead runs the full clipping code, incl. 20 sinf/cosf evaluations, for
// every pair in a 16x16 thread block with 208 B of local-memory stack per thread.
#include "common.cuh"
#include "geom.cuh"
#include "exchange.cuh"
#include "clip.cuh"
#include "../../include/glenet_geom.h"
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <math_constants.h>
#include <atomic>

namespace glenet {

// CTA shape (overridable for tuning experiments: -DGLENET_IOU_THREADS=128 -DGLENET_IOU_TR_MAX=192 -DGLENET_IOU_CTAS=7 ...)
#ifndef GLENET_IOU_THREADS
#define GLENET_IOU_THREADS 256
#endif
#ifndef GLENET_IOU_TR_MAX
#define GLENET_IOU_TR_MAX 448
#endif
#ifndef GLENET_IOU_CTAS
#define GLENET_IOU_CTAS 3
#endif
#ifndef GLENET_IOU_CLIP_PAIRS   // pairs per pass of the phased clip: one per lane of the first CLIP_PAIRS / 32 chain warps
#define GLENET_IOU_CLIP_PAIRS (GLENET_IOU_THREADS - 32)
#endif
#ifndef GLENET_IOU_QCAP
#define GLENET_IOU_QCAP 768
#endif
#ifndef GLENET_IOU_ZBYTES
#define GLENET_IOU_ZBYTES 4096
#endif
constexpr int IOU_THREADS = GLENET_IOU_THREADS;   // (threads / 32 - 1) "chain" warps (cull, prepare, clip) + 1 fill warp
constexpr int IOU_CHAIN = IOU_THREADS - 32;
constexpr int IOU_TR_MAX = GLENET_IOU_TR_MAX;     // tile rows (boxes_a)
constexpr int IOU_TC_MAX = 128;                   // tile cols (boxes_b)
constexpr int IOU_RPT = (IOU_TR_MAX + IOU_CHAIN - 1) / IOU_CHAIN;   // rows per thread in the circle tests of tall tiles
constexpr int IOU_QCAP = GLENET_IOU_QCAP;         // circle-test survivors per drain
constexpr int IOU_Q2CAP = IOU_QCAP + IOU_CHAIN;   // + the partial clip pass carried over from the previous drain
constexpr int IOU_CTAS_PER_SM = GLENET_IOU_CTAS;  // register budget the kernel is compiled for
constexpr int IOU_ZBYTES = GLENET_IOU_ZBYTES;     // block of zeros in shared memory = largest bulk store of the zero fill
constexpr int IOU_CLIP_PAIRS = GLENET_IOU_CLIP_PAIRS;
static_assert(IOU_CLIP_PAIRS % 32 == 0 && IOU_CLIP_PAIRS >= 32 && IOU_CLIP_PAIRS <= IOU_CHAIN, "whole warps of the chain clip");
constexpr int BPS = BP_STRIDE_BEV;         // BoxPre stride in shared memory (the z terms of 3D IoU are read per clipped pair)

enum { OUT_DENSE = 0, OUT_REDUCED = 1, OUT_BOTH = 2 };   // kernel output: the (na, nb) matrix; the coordinate list / row-column maxima; matrix + maxima
enum { MODE_OVERLAP = 0, MODE_IOU_BEV = 1, MODE_IOU3D = 2 };

#ifdef GLENET_PHASE_TIMING   // developer instrumentation: accumulated clock64() per phase, thread 0 of every CTA
__device__ unsigned long long g_phase_cycles[12];   // [0..7] phase cycles, [8] queued pairs, [9] pairs clipped, [10] boxes prepared, [11] drains
__device__ unsigned long long g_cta_log[4096 * 4];   // per CTA: start ns, end ns, queued pairs, clipped pairs
__device__ int g_dbg_flags;   // bit 0: skip the clip pass, bit 1: skip the zero fill (timing experiments only)
#define PHASE_MARK(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_phase_cycles[k], (unsigned long long)(now_ - t_phase_)); t_phase_ = now_; } } while (0)
#define PHASE_INIT long long t_phase_ = clock64()
#else
#define PHASE_MARK(k) do { } while (0)
#define PHASE_INIT do { } while (0)
#endif

struct IouFrames {
    long long stride_a, stride_b, stride_out;   // per-frame strides in floats
    int na;
    // multi-GPU row slabs: this launch computes rows [row_offset, row_offset + na) of a matrix with na_total rows; the
    // coordinate list and the column keys carry GLOBAL row indices (both are what crosses NVLink)
    int row_offset; long long na_total;
    // sparse output (sp_count != nullptr): instead of the dense matrix, the non-zero elements are appended as
    // (flat index frame * na * nb + row * nb + col, value) in no particular order; sp_count keeps counting past sp_cap
    long long* sp_idx; float* sp_val; unsigned long long* sp_count; long long sp_cap;
    // reduced output (row_key != nullptr): per row / per column  max over the other axis of
    // (value bits << 32) | (0xffffffff - index), i.e. the maximum and the FIRST index attaining it; 0 = no non-zero element
    unsigned long long* row_key; unsigned long long* col_key;
    // with ONE column tile every row's maximum is final inside its tile: the tile writes the decoded (max, first argmax) itself
    // and no row key ever reaches global memory (row_key is then only a non-null marker)
    float* row_max; long long* row_arg;
    IouPeers ex;   // ex.world <= 1: single GPU, nothing below the struct's first member is read
};
__device__ __forceinline__ bool iou_no_matrix(const IouFrames& fr) { return fr.sp_count != nullptr || fr.row_key != nullptr; }

struct __align__(128) IouSmem {
    float4 zero[IOU_ZBYTES / 16];          // source of the bulk zero fill
    float2 verts[IOU_CLIP_PAIRS * CLIP_SLOTS];              // vertex slots of the phased clip (clip.cuh), one pair per lane
    unsigned int wl[IOU_CLIP_PAIRS / 32][32 * CLIP_SLOTS];  // per-warp work lists of its phase B
    unsigned long long rkey[IOU_TR_MAX], ckey[IOU_TC_MAX];  // the tile's row / column maxima (key output modes), flushed to global memory once
    float rrad[IOU_TR_MAX], crad[IOU_TC_MAX];               // cull radii; the centres are slots 0 / 1 of the records below
    float rpre[IOU_TR_MAX * BPS];          // rows: raw box in the first 7 slots until a pair needs it, then the BoxPre record (same centre slots)
    float cpre[IOU_TC_MAX * BPS];
    float qres[IOU_Q2CAP];                 // clipped results, parked until the zero fill has landed
    unsigned short queue[IOU_QCAP];        // (row << 7) | col  (row < 448, col < 128): survivors of the circle test
    unsigned short queue2[IOU_Q2CAP];      // survivors of the separating-axis test: the pairs that are clipped
    unsigned short plist[IOU_TR_MAX + IOU_TC_MAX];   // boxes to prepare in the current drain
    float red[IOU_THREADS / 32][5];
    unsigned char rflag[IOU_TR_MAX], cflag[IOU_TC_MAX], act[IOU_TC_MAX];   // flags: 0 = unused, 1 = wanted, 2 = prepared
    int qcount, q2count, nact, nprep;
};
static_assert(IOU_TR_MAX * IOU_TC_MAX <= 65536 && IOU_TC_MAX == 128, "queue entries are (row << 7) | col in 16 bits");

#ifndef GLENET_IOU_LDHINT   // 1: box loads of the tile prologue carry an L2 evict-last policy
#define GLENET_IOU_LDHINT 1
#endif
// ---- bulk-copy engine (TMA without a tensor map): shared -> global, tracked by bulk async-groups
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned int bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(__cvta_generic_to_global(gdst)), "r"((unsigned int)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// ---- single-instruction warp reductions of sm_100a (CREDUX.F32); NaN inputs are ignored like fminf / fmaxf do
__device__ __forceinline__ float warp_min(float v) { float r; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float warp_max(float v) { float r; asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
// ---- named barriers: 1 = the chain warps among themselves, 2 = "the tile's zero fill has landed" (fill warp arrives, chain warps wait)
__device__ __forceinline__ void chain_sync() { asm volatile("bar.sync 1, %0;" :: "n"(IOU_CHAIN) : "memory"); }
__device__ __forceinline__ void fill_arrive() { asm volatile("bar.arrive 2, %0;" :: "n"(IOU_THREADS) : "memory"); }
__device__ __forceinline__ void fill_wait() { asm volatile("bar.sync 2, %0;" :: "n"(IOU_THREADS) : "memory"); }

// z terms of boxes_iou3d_gpu (iou3d_nms_utils.py:100-117) for one box, every step separately rounded as torch does
struct ZTerms { float zmin, zmax, vol; };
__device__ __forceinline__ ZTerms z_terms(float z, float dz, float area) {
    const float hz = __fmul_rn(dz, 0.5f);
    ZTerms t;
    t.zmin = __fsub_rn(z, hz); t.zmax = __fadd_rn(z, hz); t.vol = __fmul_rn(area, dz);
    return t;
}
__device__ __forceinline__ bool z_terms_finite(const ZTerms& t) { return fabsf(t.zmin) + fabsf(t.zmax) + fabsf(t.vol) < CUDART_INF_F; }

template <int MODE>
__device__ __forceinline__ float finish_pair(const float* a, const float* b, float ov, const float* __restrict__ boxa, const float* __restrict__ boxb) {
    if (MODE == MODE_OVERLAP) return ov;
    if (MODE == MODE_IOU_BEV) return iou_from_overlap(a[BP_AREA], b[BP_AREA], ov);
    const ZTerms za = z_terms(boxa[2], boxa[5], a[BP_AREA]), zb = z_terms(boxb[2], boxb[5], b[BP_AREA]);
    return iou3d_from_terms(za.zmin, za.zmax, za.vol, zb.zmin, zb.zmax, zb.vol, ov);
}

// Zero fill of the tile, executed by ONE warp while the other seven work on the tile's pairs.  16-byte aligned
// tiles go through the bulk-copy engine (a few dozen 4 KB shared -> global copies from a block of zeros);
// everything else falls back to plain stores.  Either way the warp blocks on back-pressure from the memory
// system -- a wave of tiles is ~85 MB of stores -- and nobody waits for it until the results are due.
__device__ __forceinline__ void zero_fill_tile(IouSmem& sm, float* __restrict__ out_tile, int tr, int tc, int nb, bool vec, int lane) {
    if (vec) {
        if (tc == nb) {   // the tile is one contiguous range of the matrix
            const size_t total = (size_t)tr * nb * sizeof(float);
            char* dst = reinterpret_cast<char*>(out_tile);
            for (size_t off = (size_t)lane * IOU_ZBYTES; off < total; off += (size_t)32 * IOU_ZBYTES) {
                const size_t left = total - off;
                bulk_store(dst + off, sm.zero, (unsigned int)(left < (size_t)IOU_ZBYTES ? left : (size_t)IOU_ZBYTES));
            }
        } else {
            for (int r = lane; r < tr; r += 32) bulk_store(out_tile + (size_t)r * nb, sm.zero, (unsigned int)tc * sizeof(float));
        }
        bulk_commit();
        bulk_wait_all();       // writes of this thread's bulk copies are complete ...
        fence_proxy_async();   // ... and ordered before generic-proxy accesses that follow the barrier
    } else {
        const int npairs = tr * tc;
        for (int p = lane; p < npairs; p += 32) { const int r = p / tc; out_tile[(size_t)r * nb + (p - r * tc)] = 0.f; }
    }
}

// Drain the first n queue entries (survivors of the circle test).  Chain warps only.
//   1. the boxes they touch are compacted into a list and prepared (BoxPre) on full warps,
//   2. a separating-axis test on the prepared records drops the pairs whose overlap is exactly 0 in the
//      reference too (sat_separated); the rest is appended to queue2,
//   3. queue2 is clipped one pair per thread -- in full passes only unless this is the tile's last drain; the
//      remainder stays in queue2 for the next drain -- and the results are parked in shared memory,
//   4. once the fill warp has signalled that the tile's zeros have landed (first drain only) they are written.
template <int MODE, bool FMA, int OUT>
__device__ __forceinline__ void drain_queue(IouSmem& sm, const float* __restrict__ A, const float* __restrict__ B,
                                            const float4* __restrict__ trigA, const float4* __restrict__ trigB,
                                            int r0, int c0, int tr, int tc, int nb, float* __restrict__ out, int n, bool last, bool& fill_pending,
                                            const IouFrames& fr, long long frame_base) {
    const int tid = threadIdx.x, lane = tid & 31;
    PHASE_INIT;
    for (int i0 = 0; i0 < tr + tc; i0 += IOU_CHAIN) {
        const int i = i0 + tid;
        const bool want = i < tr + tc && (i < tr ? sm.rflag[i] : sm.cflag[i - tr]) == 1;
        const unsigned int m = __ballot_sync(0xffffffffu, want);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&sm.nprep, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (want) sm.plist[base + __popc(m & ((1u << lane) - 1))] = (unsigned short)i;
        }
    }
    chain_sync();
    const int nprep = sm.nprep;
    for (int j = tid; j < nprep; j += IOU_CHAIN) {
        const int i = sm.plist[j];
        const bool is_row = i < tr;
        const int k = is_row ? i : i - tr;
        float* rec = (is_row ? sm.rpre : sm.cpre) + k * BPS;
        float raw[7];   // staged in the record's first slots by the tile prologue: no global load behind the zero fill's stores
#pragma unroll
        for (int f = 0; f < 7; ++f) raw[f] = rec[f];
        const float4* trig = is_row ? trigA : trigB;
        const float4 t4 = FMA ? device_trig(raw[6]) : trig[is_row ? r0 + k : c0 + k];   // CPU dialect: host-libm table (launcher checks it is there)
        box_prepare<FMA, false>(raw, t4, rec);
        *(is_row ? &sm.rflag[k] : &sm.cflag[k]) = 2;
    }
    chain_sync();
    PHASE_MARK(3);
    for (int q0 = 0; q0 < n; q0 += IOU_CHAIN) {
        const int q = q0 + tid;
        unsigned int e = 0;
        bool keep = false;
        if (q < n) {
            e = sm.queue[q];
            const float* a = sm.rpre + (e >> 7) * BPS;
            const float* b = sm.cpre + (e & 127) * BPS;
            keep = !sat_separated(a, b);
            if (MODE == MODE_IOU3D && !keep) {   // 0 * NaN: a non-finite z term turns even a zero BEV overlap into NaN (iou3d_nms_utils.py:107-117)
                const float* boxa = A + (size_t)(r0 + (e >> 7)) * 7;
                const float* boxb = B + (size_t)(c0 + (e & 127)) * 7;
                keep = !z_terms_finite(z_terms(boxa[2], boxa[5], a[BP_AREA])) || !z_terms_finite(z_terms(boxb[2], boxb[5], b[BP_AREA]));
            }
        }
        const unsigned int m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&sm.q2count, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) sm.queue2[base + __popc(m & ((1u << lane) - 1))] = (unsigned short)e;
        }
    }
    chain_sync();
    PHASE_MARK(7);
    const int n2 = sm.q2count;
    int nclip = last ? n2 : n2 / IOU_CHAIN * IOU_CHAIN;
#ifdef GLENET_PHASE_TIMING
    if (g_dbg_flags & 1) nclip = 0;
    { const unsigned int cl = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
      if (tid == 0 && cl < 4096) { g_cta_log[cl * 4 + 2] += n; g_cta_log[cl * 4 + 3] += nclip; } }
    if (tid == 0) { atomicAdd(&g_phase_cycles[8], (unsigned long long)n); atomicAdd(&g_phase_cycles[9], (unsigned long long)nclip);
                    atomicAdd(&g_phase_cycles[10], (unsigned long long)nprep); atomicAdd(&g_phase_cycles[11], 1ull); }
#endif
    // ---- the phased clip of clip.cuh over queue2[0, nclip): one pair per lane, A (24 result bits, corners to their slots),
    //      B (the warp's crossings pooled, one per lane) and C (sort + fan) are warp-local -- no CTA barrier in here
    if (tid < IOU_CLIP_PAIRS) {
        const int cw = tid >> 5;
#pragma unroll 1
        for (int q = tid; q - (tid & 31) < nclip; q += IOU_CLIP_PAIRS) {   // warp-uniform trip count
            const bool live = q < nclip;
            const unsigned int e = live ? sm.queue2[q] : 0u;
            const float* a = sm.rpre + (e >> 7) * BPS;
            const float* b = sm.cpre + (e & 127) * BPS;
            float2* slots = sm.verts + tid * CLIP_SLOTS;
            const unsigned int w = clip_pair_tests<FMA>(a, b, live);
            const unsigned int hits = clip_hits16(w);
            const int cnt = __popc(hits) + __popc(clip_corners8(w));
            const bool fast = cnt >= 3 && cnt <= CLIP_SLOTS, slow = cnt > CLIP_SLOTS;
            if (fast) clip_write_corners(a, b, w, slots);
            clip_warp_points<FMA>(fast ? hits : 0u, e >> 7, e & 127u, sm.wl[cw], sm.rpre, sm.cpre, BPS, sm.verts + (cw * 32) * CLIP_SLOTS);
            const float ov_slow = clip_warp_slow<FMA>(slow, w, e >> 7, e & 127u, sm.rpre, sm.cpre, BPS, reinterpret_cast<float2*>(sm.wl[cw]));
            if (live) {
                const float ov = slow ? ov_slow : (fast ? clip_area8<FMA>(slots, cnt) : 0.f);
                sm.qres[q] = finish_pair<MODE>(a, b, ov, A + (size_t)(r0 + (e >> 7)) * 7, B + (size_t)(c0 + (e & 127)) * 7);
            }
        }
    }
    PHASE_MARK(4);
    unsigned short carry = 0;
    const int rem = n2 - nclip;   // < IOU_CHAIN unless the clip pass is disabled for a timing experiment
    if (tid < rem) carry = sm.queue2[nclip + tid];
    if (OUT != OUT_REDUCED && fill_pending && (nclip > 0 || last)) { fill_wait(); fill_pending = false; }   // also a barrier among the chain warps
    else chain_sync();
    PHASE_MARK(6);
    if (OUT != OUT_DENSE && fr.row_key) {
        const int frame = blockIdx.z;
        for (int q = tid; q < nclip; q += IOU_CHAIN) {
            const unsigned int e = sm.queue2[q];
            const float v = sm.qres[q];
            if (!(v == 0.f)) {
                // Folded in SHARED memory and flushed once per tile: global atomics issued here, between the tile's barriers, wait
                // for their acknowledgement from an L2 that is busy with a wave of zero fills (measured: + 80 % on the dense + keys launch).
                const unsigned long long hi = (unsigned long long)__float_as_uint(v) << 32;
                atomicMax(&sm.rkey[e >> 7], hi | (0xffffffffu - (c0 + (e & 127))));
                atomicMax(&sm.ckey[e & 127], hi | (0xffffffffu - (unsigned int)(r0 + (e >> 7) + fr.row_offset)));
            }
        }
        (void)frame;
    }
    if (OUT == OUT_REDUCED && fr.sp_count) {
        for (int q0 = 0; q0 < nclip; q0 += IOU_CHAIN) {
            const int q = q0 + tid;
            unsigned int e = 0;
            float v = 0.f;
            if (q < nclip) { e = sm.queue2[q]; v = sm.qres[q]; }
            const bool nz = !(v == 0.f);   // NaN counts: the dense matrix would hold it too
            const unsigned int m = __ballot_sync(0xffffffffu, nz);
            if (m) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(fr.sp_count, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0) + __popc(m & ((1u << lane) - 1));
                if (nz && (long long)base < fr.sp_cap) {
                    const long long gi = frame_base + (long long)(r0 + (e >> 7)) * nb + (c0 + (e & 127));
                    fr.sp_idx[base] = gi;
                    fr.sp_val[base] = v;
                    // multi-GPU gather: the same entry goes into our segment of every peer's window (plain stores over NVLink)
                    for (int p = 0; p < fr.ex.world; ++p)
                        if (p != fr.ex.rank) { fr.ex.idx[p][base] = gi; fr.ex.val[p][base] = v; }
                }
            }
        }
    }
    if (OUT != OUT_REDUCED) {
        for (int q = tid; q < nclip; q += IOU_CHAIN) {
            const unsigned int e = sm.queue2[q];
            out[(size_t)(r0 + (e >> 7)) * nb + (c0 + (e & 127))] = sm.qres[q];
        }
    }
    chain_sync();   // queue2[0, nclip) has been read by everyone
    if (tid < rem) sm.queue2[tid] = carry;
    if (tid == 0) { sm.qcount = 0; sm.q2count = rem < IOU_CHAIN ? rem : 0; sm.nprep = 0; }
    chain_sync();
    PHASE_MARK(5);
}

// Multi-GPU epilogue of the tile kernel (chain warps; every CTA of the grid reaches it exactly once).
//   keys : every CTA has folded its results into the LOCAL column keys; the last CTA to finish pushes the non-empty keys into
//          every peer's window with system-scope atomic max over NVLink (a few KB) -- the cross-rank "max + first index"
//          all-reduce of the target assigner, fused into the kernel that produced the values;
//   list : every CTA has already stored its entries into every peer's window (drain_queue); the last CTA publishes the length.
// Then it raises this rank's flag in every window (its own included); consumers (exchange.cu) wait on the flags.
__device__ __noinline__ void exchange_epilogue(IouSmem& sm, const IouFrames& fr, int nb) {
    const int tid = threadIdx.x;
    if (fr.sp_count) __threadfence_system();   // this thread's list entries in the peers' windows are visible before the CTA counts as done
    else __threadfence();
    chain_sync();
    if (tid == 0) {
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        sm.nact = (atomicAdd(fr.ex.done, 1u) == total - 1u) ? 1 : 0;
    }
    chain_sync();
    if (!sm.nact) return;
    __threadfence();
    const int world = fr.ex.world, rank = fr.ex.rank;
    if (fr.sp_count) {
        const unsigned long long n = *reinterpret_cast<volatile unsigned long long*>(fr.sp_count);
        if (tid < world) *reinterpret_cast<volatile unsigned long long*>(fr.ex.cnt[tid]) = n;
    } else {
        // the finished local keys go into our slot of every peer's window: plain 16-byte stores (the consumer zeroes the
        // slot after reading it, so zeros need not travel -- but a 16-byte store of two keys is cheaper than a test)
        const int nkeys = (int)gridDim.z * nb;
        const unsigned long long* mine = fr.ex.col_key[rank];
        const int pairs = nkeys >> 1;
        for (int i = tid; i < pairs; i += IOU_CHAIN) {
            const ulonglong2 k = __ldcg(reinterpret_cast<const ulonglong2*>(mine) + i);
            for (int p = 0; p < world; ++p) if (p != rank) reinterpret_cast<ulonglong2*>(fr.ex.col_key[p])[i] = k;
        }
        if ((nkeys & 1) && tid == 0) {
            const unsigned long long k = __ldcg(mine + nkeys - 1);
            for (int p = 0; p < world; ++p) if (p != rank) fr.ex.col_key[p][nkeys - 1] = k;
        }
    }
    __threadfence_system();
    chain_sync();
    if (tid == 0) *fr.ex.done = 0u;   // ready for the next launch
    if (tid < world) *reinterpret_cast<volatile unsigned int*>(fr.ex.flag[tid] + rank) = fr.ex.step;
}

template <int MODE, bool FMA, int OUT>
__global__ void __launch_bounds__(IOU_THREADS, IOU_CTAS_PER_SM)
iou_tile_kernel(const float* __restrict__ A, int na, const float* __restrict__ B, int nb,
                const float4* __restrict__ trigA, const float4* __restrict__ trigB,
                float* __restrict__ out, int TR, int TC, const __grid_constant__ IouFrames fr) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    IouSmem& sm = *reinterpret_cast<IouSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // frames: independent (na, nb) problems of one launch, e.g. the GT sets of a batch against the same anchors
    // grid = (column tiles, row tiles, frames): no integer division to find the tile
    const int frame = blockIdx.z, tile_r = blockIdx.y, tile_c = blockIdx.x;
    A += (size_t)frame * fr.stride_a; B += (size_t)frame * fr.stride_b; out += (size_t)frame * fr.stride_out;
    const long long frame_base = ((long long)frame * fr.na_total + fr.row_offset) * nb;   // flat index of this launch's row 0 in frame `frame`
    // Programmatic dependent launch: this grid may have been made resident while the previous kernel of the
    // stream was still draining; everything below reads or writes global memory, so wait for it here.  The
    // launch latency and the CTA scheduling of back-to-back calls is what gets hidden.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int r0 = tile_r * TR, c0 = tile_c * TC;
    const int tr = min(TR, na - r0), tc = min(TC, nb - c0);
    PHASE_INIT;
#ifdef GLENET_PHASE_TIMING
    const unsigned int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (tid == 0 && cta_lin < 4096) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_cta_log[cta_lin * 4] = t; g_cta_log[cta_lin * 4 + 2] = 0; g_cta_log[cta_lin * 4 + 3] = 0; }
#endif

    // ---- stage the tile's boxes: centre + cull radius for the circle tests, the raw box in the first slots of
    //      its BoxPre record (prepared in place if a pair needs it), and the bounding box of the row centres.
    //      All eight warps take part, so every global load of the tile is back before the zero fill starts.
    float minx = FLT_MAX, maxx = -FLT_MAX, miny = FLT_MAX, maxy = -FLT_MAX, maxr = 0.f;
    for (int i = tid; i < tr + tc; i += IOU_THREADS) {
        const bool is_row = i < tr;
        const int k = is_row ? i : i - tr;
        const float* box = (is_row ? A + (size_t)(r0 + k) * 7 : B + (size_t)(c0 + k) * 7);
        float raw[7];
#if GLENET_IOU_LDHINT
        // the boxes are re-read by every frame / tile row while 1.4 GB of results stream through L2: keep them (evict-last)
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#pragma unroll
        for (int f = 0; f < 7; ++f)
            asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(raw[f]) : "l"(box + f), "l"(pol));
#else
#pragma unroll
        for (int f = 0; f < 7; ++f) raw[f] = box[f];
#endif
        float* rec = (is_row ? sm.rpre : sm.cpre) + k * BPS;
#pragma unroll
        for (int f = 0; f < 7; ++f) rec[f] = raw[f];
        const float cx = raw[0], cy = raw[1];
        float rad = cull_radius(cx, cy, raw[3], raw[4]);
        // 3D IoU multiplies the BEV overlap by the z overlap in torch: 0 * NaN = NaN, so a box with a
        // non-finite z term must reach finish_pair for every pair
        if (MODE == MODE_IOU3D && !z_terms_finite(z_terms(raw[2], raw[5], __fmul_rn(raw[3], raw[4])))) rad = CUDART_INF_F;
        if (is_row) {
            sm.rrad[k] = rad; sm.rflag[k] = 0;
            if (OUT != OUT_DENSE) sm.rkey[k] = 0ull;
            minx = fminf(minx, cx); maxx = fmaxf(maxx, cx); miny = fminf(miny, cy); maxy = fmaxf(maxy, cy);
            maxr = (rad != rad) ? CUDART_INF_F : fmaxf(maxr, rad);   // a NaN radius must not be dropped by fmaxf
        } else { sm.crad[k] = rad; sm.cflag[k] = 0; if (OUT != OUT_DENSE) sm.ckey[k] = 0ull; }
    }
    minx = warp_min(minx); maxx = warp_max(maxx); miny = warp_min(miny); maxy = warp_max(maxy); maxr = warp_max(maxr);
    if (lane == 0) { sm.red[warp][0] = minx; sm.red[warp][1] = maxx; sm.red[warp][2] = miny; sm.red[warp][3] = maxy; sm.red[warp][4] = maxr; }
    if (tid == 0) { sm.qcount = 0; sm.q2count = 0; sm.nact = 0; sm.nprep = 0; }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel of the stream may be scheduled as SMs free up

    if (warp == IOU_CHAIN / 32) {   // ---- the fill warp
        if (OUT == OUT_REDUCED) return;   // no matrix, nothing to fill, and nobody waits for this warp
        const bool vec = ((nb & 3) == 0) && ((c0 & 3) == 0) && ((tc & 3) == 0) && ((((uintptr_t)out) & 15) == 0);
        if (vec) {   // the block of zeros the bulk copies read: written and fenced by the warp that issues them
#pragma unroll
            for (int k = 0; k < IOU_ZBYTES / 16 / 32; ++k) sm.zero[k * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            fence_proxy_async();   // generic-proxy writes -> async-proxy reads
            __syncwarp();
        }
#ifdef GLENET_PHASE_TIMING
        if (!(g_dbg_flags & 2))
#endif
        zero_fill_tile(sm, out + (size_t)r0 * nb + c0, tr, tc, nb, vec, lane);
        __threadfence_block();
        fill_arrive();
        return;
    }
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
    for (int c = tid; c < tc; c += IOU_CHAIN) {
        const float cx = sm.cpre[c * BPS + BP_CX], cy = sm.cpre[c * BPS + BP_CY];
        const float ddx = fmaxf(fmaxf(minx - cx, cx - maxx), 0.f), ddy = fmaxf(fmaxf(miny - cy, cy - maxy), 0.f);
        const float rr = maxr + sm.crad[c];
        const bool far = (cx == cx) && (cy == cy) && (ddx * ddx + ddy * ddy > rr * rr);
        if (!far) sm.act[atomicAdd(&sm.nact, 1)] = (unsigned char)c;
    }
    chain_sync();   // act[] / nact complete
    PHASE_MARK(1);

    // ---- per-pair circle test on the active columns.  Warps are laid out as (row warps) x (column groups):
    //      every thread tests its row(s) against up to 32 columns of its group into a register bitmask, then the
    //      survivors are appended to the queue with one atomic per warp.  Slots are reserved optimistically:
    //      entries that fall beyond the queue's capacity stay in the bitmask, the queue is drained and they are
    //      appended in the next round (dense tiles only).
    const int nact = sm.nact;
    const int rwarps = (tr + 31) >> 5;                                   // <= 12
    const int groups = rwarps <= IOU_CHAIN / 32 ? (IOU_CHAIN / 32) / rwarps : 1;
    const bool tall = rwarps > IOU_CHAIN / 32;                           // two rows per thread, one column group
    const int grp = tall ? 0 : warp / rwarps;
    const bool idle = !tall && grp >= groups;
    int rows[IOU_RPT];
    float3 rw[IOU_RPT];   // {cx, cy, cull radius}
#pragma unroll
    for (int j = 0; j < IOU_RPT; ++j) {
        rows[j] = tall ? tid + j * IOU_CHAIN : (j == 0 ? (warp - grp * rwarps) * 32 + lane : IOU_TR_MAX);
        if (idle || rows[j] >= tr) rows[j] = -1;
        rw[j] = rows[j] >= 0 ? make_float3(sm.rpre[rows[j] * BPS + BP_CX], sm.rpre[rows[j] * BPS + BP_CY], sm.rrad[rows[j]]) : make_float3(0.f, 0.f, 0.f);
    }
    bool fill_pending = true;
    for (int cb = 0; cb < nact;) {
        const int per = min(32, (nact - cb + groups - 1) / groups);     // columns per group in this round
        const int k0 = min(nact, cb + grp * per), k1 = idle ? k0 : min(nact, k0 + per);
        cb += per * groups;
        unsigned int m[IOU_RPT];
#pragma unroll
        for (int j = 0; j < IOU_RPT; ++j) m[j] = 0u;
#pragma unroll 4
        for (int k = k0; k < k1; ++k) {
            const int c = sm.act[k];
            const float cx = sm.cpre[c * BPS + BP_CX], cy = sm.cpre[c * BPS + BP_CY], cr = sm.crad[c];
#pragma unroll
            for (int j = 0; j < IOU_RPT; ++j) {
                const float dx = rw[j].x - cx, dy = rw[j].y - cy, rr = rw[j].z + cr;
                // NaN anywhere => the comparison is false => not culled => the clip pass decides, like the reference
                if (!(dx * dx + dy * dy > rr * rr)) m[j] |= 1u << (k - k0);
            }
        }
#pragma unroll
        for (int j = 0; j < IOU_RPT; ++j) if (rows[j] < 0) m[j] = 0u;
        for (;;) {
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < IOU_RPT; ++j) cnt += __popc(m[j]);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const int wtotal = __shfl_sync(0xffffffffu, incl, 31);
            int base = 0;
            if (wtotal) {
                if (lane == 31) base = atomicAdd(&sm.qcount, wtotal);
                base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
            }
            chain_sync();
            const int total = sm.qcount;   // the same value in every thread: nobody adds again before the next barrier
#pragma unroll
            for (int j = 0; j < IOU_RPT; ++j) {
                while (m[j] && base < IOU_QCAP) {
                    const int k = __ffs(m[j]) - 1;
                    m[j] &= m[j] - 1;
                    const int c = sm.act[k0 + k];
                    sm.queue[base++] = (unsigned short)((rows[j] << 7) | c);
                    if (sm.rflag[rows[j]] == 0) sm.rflag[rows[j]] = 1;
                    if (sm.cflag[c] == 0) sm.cflag[c] = 1;
                }
            }
            chain_sync();
            // ONE call site for the drain (it is ~2000 instructions; a second inlined copy costs instruction-cache space):
            // a full queue is drained and the round repeated; after the last chunk's round the tile's final drain runs.
            const bool overflow = total > IOU_QCAP;
            const bool final_drain = !overflow && cb >= nact;
            if (overflow || final_drain)
                drain_queue<MODE, FMA, OUT>(sm, A, B, trigA, trigB, r0, c0, tr, tc, nb, out, overflow ? IOU_QCAP : total, final_drain, fill_pending, fr, frame_base);
            if (!overflow) break;
        }
    }
    PHASE_MARK(2);
    // (a tile without active columns has nothing to write: its chain warps leave, the fill warp finishes on its own)
    if (OUT != OUT_DENSE && fr.row_key) {   // the tile's maxima -> global keys; nothing in this CTA waits for these atomics
        chain_sync();
        unsigned long long* gcol = fr.col_key + (size_t)frame * nb + c0;
        const size_t row0 = (size_t)frame * fr.na + r0;
        for (int i = tid; i < tr + tc; i += IOU_CHAIN) {
            const unsigned long long k = i < tr ? sm.rkey[i] : sm.ckey[i - tr];
            if (i >= tr) { if (k) atomicMax(gcol + (i - tr), k); }
            else if (fr.row_max) {   // decoded in place (the launcher only passes row_max with a single column tile)
                fr.row_max[row0 + i] = __uint_as_float((unsigned int)(k >> 32));
                fr.row_arg[row0 + i] = k ? (long long)(0xffffffffu - (unsigned int)k) : 0;
            } else if (k) atomicMax(fr.row_key + row0 + i, k);
        }
    }
    if (OUT != OUT_DENSE && fr.ex.world > 1) exchange_epilogue(sm, fr, nb);
#ifdef GLENET_PHASE_TIMING
    if (tid == 0 && cta_lin < 4096) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_cta_log[cta_lin * 4 + 1] = t; }
#endif
}

// out[i] = f(a[i], b[i / group]) -- every pair is "heavy" by construction (CVAE samples vs their GT), so this is the
// FP32-bound workload of the path.  Persistent CTAs walk batches of AL_PAIRS pairs through the phased clip of clip.cuh:
//   prepare (one lane per box, BoxPre records in shared memory) | barrier |
//   A (one lane per pair: 24 result bits, corners to their slots) -> B (the warp's crossings pooled, one per lane)
//   -> C (one lane per pair: sort + fan) -- A, B and C are warp-local, no CTA barrier between them.
// Pairs with more than eight vertices (~1 %) are finished by their warp, all 32 lanes on one pair (clip_warp_slow).
constexpr int AL_THREADS = 256;
constexpr int AL_PAIRS = AL_THREADS;        // one pair per lane and batch
#ifndef GLENET_AL_CTAS
#define GLENET_AL_CTAS 4
#endif
constexpr int AL_CTAS_PER_SM = GLENET_AL_CTAS;
struct AlignedSmem {
    float2 verts[AL_PAIRS * CLIP_SLOTS];
    unsigned int wl[AL_THREADS / 32][32 * CLIP_SLOTS];   // per-warp work lists of phase B (then the scratch of clip_warp_slow)
    float arec[AL_PAIRS * BPS];
    float brec[1];                          // [2][nb_max * (BPS + 1)], sized by the launcher: the b boxes one batch can meet (+ cull radii), double-buffered
};

// n / d for n < 2^31 by multiply-shift (Granlund-Montgomery); the launcher precomputes {m, s} for the runtime `group`
struct FastDiv { unsigned int m, s, d; };
static FastDiv fastdiv_make(unsigned int d) {
    FastDiv f; f.d = d; f.s = 0; f.m = 0;
    if (d <= 1) return f;
    while ((1u << f.s) < d) ++f.s;                                   // s = ceil(log2 d)
    f.m = (unsigned int)((((unsigned long long)1 << 32) * ((1ull << f.s) - d)) / d + 1);
    return f;
}
__device__ __forceinline__ unsigned int fastdiv(unsigned int n, const FastDiv& f) {
    if (f.d <= 1) return n;
    const unsigned int t = __umulhi(f.m, n);
    return (t + ((n - t) >> 1)) >> (f.s - 1);
}

template <int MODE, bool FMA>
__device__ __forceinline__ void aligned_stage_b(const float* __restrict__ B, int b0, int nbx, float* __restrict__ brec, float* __restrict__ brad, int tid) {
    for (int k = tid; k < nbx; k += AL_THREADS) {
        const float* bb = B + (size_t)(b0 + k) * 7;
        float raw[7];
#pragma unroll
        for (int f = 0; f < 7; ++f) raw[f] = bb[f];
        box_prepare<FMA, false>(raw, device_trig_fused(raw[6]), brec + k * BPS);
        float rad = cull_radius(raw);
        // 3D IoU: 0 * NaN = NaN in the reference's torch arithmetic, so a pair with a non-finite z term is never culled
        if (MODE == MODE_IOU3D && !z_terms_finite(z_terms(raw[2], raw[5], __fmul_rn(raw[3], raw[4])))) rad = CUDART_INF_F;
        brad[k] = rad;
    }
}

template <int MODE, bool FMA>
__global__ void __launch_bounds__(AL_THREADS, AL_CTAS_PER_SM)
iou_aligned_kernel(const float* __restrict__ A, int na, const float* __restrict__ B, const FastDiv group, float* __restrict__ out, int nbatches, int nb_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AlignedSmem& sm = *reinterpret_cast<AlignedSmem*>(smem_raw);
    // two sets of b records (+ cull radii; +inf = never cull): the first lanes stage the NEXT batch's while the current one is clipped
    const int bset = nb_max * (BPS + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    auto batch_b = [&](int bt, int& b0, int& nbx) {
        const int p0 = bt * AL_PAIRS, np = min(AL_PAIRS, na - p0);
        b0 = (int)fastdiv((unsigned int)p0, group);
        nbx = (int)fastdiv((unsigned int)(p0 + np - 1), group) - b0 + 1;
    };
    int b0, nbx;
    if ((int)blockIdx.x < nbatches) {
        batch_b(blockIdx.x, b0, nbx);
        aligned_stage_b<MODE, FMA>(B, b0, nbx, sm.brec, sm.brec + nb_max * BPS, tid);
    }
    int cur = 0;
    for (int bt = blockIdx.x; bt < nbatches; bt += gridDim.x, cur ^= 1) {
        const int p0 = bt * AL_PAIRS, np = min(AL_PAIRS, na - p0);
        batch_b(bt, b0, nbx);
        const float* brec = sm.brec + cur * bset;
        const float* brad = brec + nb_max * BPS;
        __syncthreads();   // this batch's b records are staged; everybody is done with the other set
        if (bt + (int)gridDim.x < nbatches) {
            int nb0, nnbx;
            batch_b(bt + gridDim.x, nb0, nnbx);
            float* nrec = sm.brec + (cur ^ 1) * bset;
            aligned_stage_b<MODE, FMA>(B, nb0, nnbx, nrec, nrec + nb_max * BPS, tid);
        }
        // this lane's pair
        const int gi = p0 + tid;
        const int bsel = tid < np ? (int)fastdiv((unsigned int)gi, group) - b0 : 0;
        float raw[7];
        const float* ba = A + (size_t)(tid < np ? gi : p0) * 7;
#pragma unroll
        for (int f = 0; f < 7; ++f) raw[f] = ba[f];
        float* a = sm.arec + tid * BPS;
        float arad = cull_radius(raw);
        if (MODE == MODE_IOU3D && !z_terms_finite(z_terms(raw[2], raw[5], __fmul_rn(raw[3], raw[4])))) arad = CUDART_INF_F;
        const float* b = brec + bsel * BPS;
        bool active = false;
        if (tid < np) {
            const float ddx = raw[0] - b[BP_CX], ddy = raw[1] - b[BP_CY], rr = arad + brad[bsel];
            active = !(ddx * ddx + ddy * ddy > rr * rr);
            if (active) box_prepare<FMA, false>(raw, device_trig_fused(raw[6]), a);
            else out[gi] = 0.f;
        }
        __syncwarp();
        // ---- A: result bits, corners to their slots
        float2* slots = sm.verts + tid * CLIP_SLOTS;
        const unsigned int w = clip_pair_tests<FMA, false, true>(a, b, active);
        const unsigned int hits = clip_hits16(w);
        const int cnt = __popc(hits) + __popc(clip_corners8(w));
        const bool fast = cnt >= 3 && cnt <= CLIP_SLOTS;
        if (fast) clip_write_corners(a, b, w, slots);
        // ---- B: the warp's crossings, one per lane
        clip_warp_points<FMA>(fast ? hits : 0u, (unsigned int)tid, (unsigned int)bsel, sm.wl[warp], sm.arec, brec, BPS, sm.verts + (warp * 32) * CLIP_SLOTS);
        // ---- C: sort + fan (more than eight vertices: the whole warp, one pair at a time)
        const bool slow = cnt > CLIP_SLOTS;
        const float ov_slow = clip_warp_slow<FMA>(slow, w, (unsigned int)tid, (unsigned int)bsel, sm.arec, brec, BPS, reinterpret_cast<float2*>(sm.wl[warp]));
        if (active) {
            const float ov = slow ? ov_slow : (fast ? clip_area8<FMA>(slots, cnt) : 0.f);
            out[gi] = finish_pair<MODE>(a, b, ov, ba, B + (size_t)(b0 + bsel) * 7);
        }
    }
}

#ifdef GLENET_PHASE_TIMING
static int g_debug_tile_rows = 0;   // developer override (debug build only)
#endif
// The kernel is bound by the latency of one CTA's phase chain (~13 us) unless the zero fill of a whole wave
// of tiles takes longer, so the row-tile height is chosen to minimise  waves x max(chain, fill of one wave):
// a problem that fits one wave of resident CTAs gets exactly one, with the smallest tile that achieves it.
static int g_col_split = 1;
template <typename K>
static int resident_ctas(K kernel) {
    int dev = 0, sms = 148, per_sm = IOU_CTAS_PER_SM;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, IOU_THREADS, sizeof(IouSmem)) != cudaSuccess || per_sm < 1) per_sm = 1;
    return sms * per_sm;
}
static void pick_tiles(int na, int nb, int frames, int resident, int& TR, int& TC, int& row_tiles, int& col_tiles, bool allow_col_split = true) {
    col_tiles = (nb + IOU_TC_MAX - 1) / IOU_TC_MAX;
    {   // very small problems (< 1/4 wave): split the columns further -- a small dense matrix
        // is bound by the clip passes of its few tiles, and more CTAs are more clip lanes
        const long rt32 = (long)((na + 31) / 32) * frames;
        const long want = resident / rt32, most = (nb + 31) / 32;
        if (allow_col_split && g_col_split && rt32 * col_tiles * 4 < resident && want > col_tiles) col_tiles = (int)(want < most ? want : most);
    }
    TC = ((nb + col_tiles - 1) / col_tiles + 3) / 4 * 4;   // multiple of 4 keeps every tile on the 16-byte store path
    col_tiles = (nb + TC - 1) / TC;
    // Measured on the anchor sweep (tools/tile_rows_env_sweep.py, 128..448 rows): a wave of resident tiles takes
    // 4.7 us + 20.3 ns per tile row (staging, cull and clip all grow with the rows), unless its zero fill takes longer.
    // Few waves are quantised (the launch ends with its last, partly empty wave); many waves behave like a stream.
    const double slots = resident;
    const double fill_bytes_per_us = 6.0e6;
    double best = 0.0;
    TR = 32;
    for (int tr = 32; tr <= IOU_TR_MAX; tr += 32) {
        const double tiles = (double)((na + tr - 1) / tr) * col_tiles * frames;
        const double w = tiles / slots;
        const double waves = w <= 4.0 ? ceil(w) : w + 0.5;
        const double chain_us = 4.7 + 0.0203 * tr;
        const double fill_us = (double)tr * TC * 4.0 * (tiles < slots ? tiles : slots) / fill_bytes_per_us;
        const double cost = waves * (chain_us > fill_us ? chain_us : fill_us) + 0.002 * tr;   // ties go to the smaller tile
        if (tr == 32 || cost < best) { best = cost; TR = tr; }
    }
#ifdef GLENET_PHASE_TIMING
    if (g_debug_tile_rows) TR = g_debug_tile_rows;
#endif
    {   // tuning aid: GLENET_IOU_TILE_ROWS=<multiple of 32> overrides the choice (read once per process)
        static const int forced_rows = [] { const char* e = getenv("GLENET_IOU_TILE_ROWS"); return e ? atoi(e) : 0; }();
        if (forced_rows >= 32 && forced_rows <= IOU_TR_MAX) TR = forced_rows / 32 * 32;
    }
    row_tiles = (na + TR - 1) / TR;
}

// Everything a launch can ask for beyond the dense single-frame matrix.
struct IouLaunch {
    int frames = 1;
    long long stride_a = 0, stride_b = 0, stride_out = 0;
    long long* sp_idx = nullptr; float* sp_val = nullptr; unsigned long long* sp_count = nullptr; long long sp_cap = 0;
    unsigned long long* row_key = nullptr; unsigned long long* col_key = nullptr;
    float* row_max = nullptr; long long* row_arg = nullptr;   // direct row outputs, honoured when the launch has one column tile
    bool keys_prezeroed = false;   // the caller guarantees zeroed key buffers (the exchange's decode kernel re-zeroes them after use)
    bool dense_and_keys = false;   // write the matrix AND fold the maxima (OUT_BOTH)
    int row_offset = 0; long long na_total = 0;
    IouPeers ex = {};              // ex.world <= 1: single GPU
};

template <int MODE, bool FMA, int OUT>
static int launch_tile(const float* A, const float* trigA, int na, const float* B, const float* trigB, int nb, float* out, cudaStream_t stream,
                       const char* what, const IouLaunch& L) {
    auto kernel = iou_tile_kernel<MODE, FMA, OUT>;
    // the opt-in shared-memory size and the occupancy are per-device settings: cached per (instantiation, device)
    static std::atomic<int> resident_of[GLENET_MAX_DEVICES];
    int dev = 0;
    cudaGetDevice(&dev);
    int resident = (dev >= 0 && dev < GLENET_MAX_DEVICES) ? resident_of[dev].load(std::memory_order_acquire) : 0;
    if (!resident) {
        int rc = set_smem(kernel, sizeof(IouSmem), what, true);
        if (rc) return rc;
        resident = resident_ctas(kernel);
        if (dev >= 0 && dev < GLENET_MAX_DEVICES) resident_of[dev].store(resident, std::memory_order_release);
    }
    const int frames = L.frames;
    int TR, TC, row_tiles, col_tiles;
    pick_tiles(na, nb, frames, resident, TR, TC, row_tiles, col_tiles, L.row_max == nullptr);   // direct row outputs need whole rows in one tile
    const long long tiles = (long long)row_tiles * col_tiles;
    if (tiles * frames > 0x7fffffffLL) return fail(GLENET_EINVAL, "%s: too many tiles", what);
    IouFrames fr;
    fr.stride_a = L.stride_a; fr.stride_b = L.stride_b; fr.stride_out = OUT == OUT_REDUCED ? 0 : L.stride_out; fr.na = na;
    fr.row_offset = L.row_offset; fr.na_total = L.na_total > 0 ? L.na_total : na;
    fr.row_key = L.row_key; fr.col_key = L.col_key;
    fr.row_max = col_tiles == 1 ? L.row_max : nullptr; fr.row_arg = col_tiles == 1 ? L.row_arg : nullptr;
    fr.sp_idx = L.sp_idx; fr.sp_val = L.sp_val; fr.sp_count = L.sp_count; fr.sp_cap = L.sp_cap;
    fr.ex = L.ex;
    cudaLaunchConfig_t cfg = {};
    if (row_tiles > 65535 || frames > 65535) return fail(GLENET_EINVAL, "%s: more than 65535 row tiles or frames", what);
    cfg.gridDim = dim3((unsigned)col_tiles, (unsigned)row_tiles, (unsigned)frames); cfg.blockDim = dim3(IOU_THREADS);
    cfg.dynamicSmemBytes = sizeof(IouSmem); cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, A, na, B, nb, reinterpret_cast<const float4*>(trigA),
                                       reinterpret_cast<const float4*>(trigB), out, TR, TC, fr);
    if (e != cudaSuccess) {
        snprintf(last_error_buf(), 512, "%s: launch failed: %s", what, cudaGetErrorString(e));
        return -(int)e;
    }
    return check_launch(what);
}

template <int MODE, bool FMA>
static int launch_iou(const float* A, const float* trigA, int na, const float* B, const float* trigB, int nb,
                      float* out, cudaStream_t stream, const char* what, const IouLaunch& L = IouLaunch()) {
    const int frames = L.frames;
    if (na < 0 || nb < 0 || frames < 0) return fail(GLENET_EINVAL, "%s: negative count", what);
    if (L.sp_count) {
        if (L.sp_cap < 0 || (L.sp_cap > 0 && (!L.sp_idx || !L.sp_val))) return fail(GLENET_EINVAL, "%s: bad sparse buffers", what);
        cudaError_t e = cudaMemsetAsync(L.sp_count, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return fail(-(int)e, "%s: cudaMemsetAsync failed", what);
    }
    if (L.row_key || L.col_key) {
        if (!L.row_key || !L.col_key) return fail(GLENET_EINVAL, "%s: row and column keys go together", what);
        if (!L.keys_prezeroed) {
            cudaError_t e = cudaMemsetAsync(L.row_key, 0, sizeof(unsigned long long) * (size_t)frames * na, stream);
            if (e == cudaSuccess) e = cudaMemsetAsync(L.col_key, 0, sizeof(unsigned long long) * (size_t)frames * nb, stream);
            if (e != cudaSuccess) return fail(-(int)e, "%s: cudaMemsetAsync failed", what);
        }
    }
    const bool reduced_only = (L.sp_count || L.row_key) && !L.dense_and_keys;
    // An empty slab still has to take part in a multi-GPU exchange (its flag is what the peers wait for): the exchange entry
    // points handle that case themselves; here nothing is launched.
    if (na == 0 || nb == 0 || frames == 0) return GLENET_OK;
    if (!A || !B || (!out && !reduced_only)) return fail(GLENET_EINVAL, "%s: null pointer", what);
    if (!FMA && (!trigA || !trigB)) return fail(GLENET_EINVAL, "%s: CPU dialect needs host-evaluated trig tables", what);
    if (reduced_only) return launch_tile<MODE, FMA, OUT_REDUCED>(A, trigA, na, B, trigB, nb, out, stream, what, L);
    if (L.dense_and_keys) return launch_tile<MODE, FMA, OUT_BOTH>(A, trigA, na, B, trigB, nb, out, stream, what, L);
    return launch_tile<MODE, FMA, OUT_DENSE>(A, trigA, na, B, trigB, nb, out, stream, what, L);
}

template <bool FMA = true>
static int launch_iou_mode(int mode, const float* A, int na, const float* B, int nb, float* out, cudaStream_t stream, const char* what, const IouLaunch& L) {
    if (mode == 0) return launch_iou<MODE_OVERLAP, FMA>(A, nullptr, na, B, nullptr, nb, out, stream, what, L);
    if (mode == 1) return launch_iou<MODE_IOU_BEV, FMA>(A, nullptr, na, B, nullptr, nb, out, stream, what, L);
    if (mode == 2) return launch_iou<MODE_IOU3D, FMA>(A, nullptr, na, B, nullptr, nb, out, stream, what, L);
    return fail(GLENET_EINVAL, "%s: mode must be 0 (overlap), 1 (BEV IoU) or 2 (3D IoU)", what);
}

}  // namespace glenet

using namespace glenet;

extern "C" {

#ifdef GLENET_PHASE_TIMING
void glenet_debug_set_tile_rows(int tr) { g_debug_tile_rows = tr; }
void glenet_debug_set_col_split(int on) { g_col_split = on; }
void glenet_debug_set_flags(int f) { cudaMemcpyToSymbol(g_dbg_flags, &f, sizeof(int)); }
int glenet_debug_iou_cta_log(unsigned long long* host_out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(host_out, g_cta_log, sizeof(unsigned long long) * 4 * n);
}
int glenet_debug_iou_resident_ctas() {
    auto kernel = iou_tile_kernel<MODE_IOU_BEV, true, OUT_DENSE>;
    set_smem(kernel, sizeof(IouSmem), "debug", true);
    return resident_ctas(kernel);
}
// developer-only: read and reset the per-phase cycle accumulators of iou_tile_kernel
int glenet_debug_iou_phase_cycles(unsigned long long* host_out8) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out8, g_phase_cycles, sizeof(unsigned long long) * 12);
    unsigned long long z[12] = {0};
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
int glenet_boxes_iou_frames_gpu(int mode, const float* a, long long a_frame_stride, int na, const float* b, long long b_frame_stride,
                                int nb, float* out, int frames, glenet_stream_t s) {
    const char* what = "glenet_boxes_iou_frames_gpu";
    if (a_frame_stride < 0 || b_frame_stride < 0) return fail(GLENET_EINVAL, "%s: negative stride", what);
    IouLaunch L;
    L.frames = frames; L.stride_a = a_frame_stride; L.stride_b = b_frame_stride; L.stride_out = (long long)na * nb;
    return launch_iou_mode(mode, a, na, b, nb, out, (cudaStream_t)s, what, L);
}
int glenet_boxes_iou_frames_sparse_gpu(int mode, const float* a, long long a_frame_stride, int na, const float* b, long long b_frame_stride,
                                       int nb, int frames, long long* idx, float* val, long long cap, unsigned long long* count,
                                       glenet_stream_t s) {
    const char* what = "glenet_boxes_iou_frames_sparse_gpu";
    if (a_frame_stride < 0 || b_frame_stride < 0) return fail(GLENET_EINVAL, "%s: negative stride", what);
    if (!count) return fail(GLENET_EINVAL, "%s: null count pointer", what);
    IouLaunch L;
    L.frames = frames; L.stride_a = a_frame_stride; L.stride_b = b_frame_stride;
    L.sp_idx = idx; L.sp_val = val; L.sp_cap = cap; L.sp_count = count;
    return launch_iou_mode(mode, a, na, b, nb, nullptr, (cudaStream_t)s, what, L);
}
int glenet_boxes_iou_frames_max_gpu(int mode, const float* a, long long a_frame_stride, int na, const float* b, long long b_frame_stride,
                                    int nb, int frames, unsigned long long* row_key, unsigned long long* col_key, glenet_stream_t s) {
    const char* what = "glenet_boxes_iou_frames_max_gpu";
    if (a_frame_stride < 0 || b_frame_stride < 0) return fail(GLENET_EINVAL, "%s: negative stride", what);
    if (!row_key || !col_key) return fail(GLENET_EINVAL, "%s: null key pointer", what);
    IouLaunch L;
    L.frames = frames; L.stride_a = a_frame_stride; L.stride_b = b_frame_stride; L.row_key = row_key; L.col_key = col_key;
    return launch_iou_mode(mode, a, na, b, nb, nullptr, (cudaStream_t)s, what, L);
}
// ---------------------------------------------------------------- multi-GPU exchange (row-sharded sweep)
size_t glenet_exchange_window_bytes(int frames, int nb, long long list_cap) {
    if (frames < 0 || nb < 0 || list_cap < 0) return 0;
    return exchange_layout(frames, nb, list_cap).bytes;
}
int glenet_symm_alloc(size_t bytes, void** dev_ptr) {
    if (!dev_ptr || bytes == 0) return fail(GLENET_EINVAL, "%s: bad argument", "glenet_symm_alloc");
    cudaError_t e = cudaMalloc(dev_ptr, bytes);   // a whole allocation of its own: CUDA IPC exports allocations, not sub-blocks of a pool
    if (e == cudaSuccess) e = cudaMemset(*dev_ptr, 0, bytes);
    return e == cudaSuccess ? GLENET_OK : fail(-(int)e, "%s: cudaMalloc failed", "glenet_symm_alloc");
}
int glenet_symm_free(void* dev_ptr) {
    cudaError_t e = cudaFree(dev_ptr);
    return e == cudaSuccess ? GLENET_OK : fail(-(int)e, "%s: cudaFree failed", "glenet_symm_free");
}
int glenet_symm_export(const void* dev_ptr, unsigned char* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!dev_ptr || !handle64) return fail(GLENET_EINVAL, "%s: null pointer", "glenet_symm_export");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr));
    if (e != cudaSuccess) { snprintf(last_error_buf(), 512, "glenet_symm_export: %s", cudaGetErrorString(e)); return -(int)e; }
    memcpy(handle64, &h, 64);
    return GLENET_OK;
}
int glenet_symm_import(const unsigned char* handle64, void** peer_ptr) {
    if (!peer_ptr || !handle64) return fail(GLENET_EINVAL, "%s: null pointer", "glenet_symm_import");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { snprintf(last_error_buf(), 512, "glenet_symm_import: %s", cudaGetErrorString(e)); return -(int)e; }
    return GLENET_OK;
}
int glenet_symm_unmap(void* peer_ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
    return e == cudaSuccess ? GLENET_OK : fail(-(int)e, "%s: cudaIpcCloseMemHandle failed", "glenet_symm_unmap");
}

/* diagnostic: the error bits a consumer kernel left in the local window (1 = timed out waiting for a peer's flag,
 * 2 = a coordinate list overflowed its capacity).  Synchronises the device. */
int glenet_exchange_status(const void* window_local, unsigned int* status_host) {
    if (!window_local || !status_host) return fail(GLENET_EINVAL, "%s: null pointer", "glenet_exchange_status");
    const ExchangeLayout lay = exchange_layout(0, 0, 0);
    cudaError_t e = cudaMemcpy(status_host, reinterpret_cast<const unsigned char*>(window_local) + lay.off_status, sizeof(unsigned int), cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? GLENET_OK : fail(-(int)e, "%s: cudaMemcpy failed", "glenet_exchange_status");
}

static int exchange_args_ok(const char* what, int world, int rank, void* const* windows, int frames, int nb, long long na_total, int row_offset, int na) {
    if (world < 1 || world > GLENET_MAX_PEERS || rank < 0 || rank >= world) return fail(GLENET_EINVAL, "%s: bad world / rank", what);
    if (!windows) return fail(GLENET_EINVAL, "%s: null window table", what);
    for (int p = 0; p < world; ++p) if (!windows[p] || ((uintptr_t)windows[p] & 255)) return fail(GLENET_EALIGN, "%s: exchange windows must be non-null and 256-byte aligned", what);
    if (frames < 0 || nb < 0 || na < 0 || row_offset < 0 || na_total < (long long)row_offset + na) return fail(GLENET_EINVAL, "%s: slab outside the matrix", what);
    return GLENET_OK;
}

int glenet_boxes_iou_frames_assign_gpu(int mode, const float* a, long long a_frame_stride, int na, const float* b, long long b_frame_stride,
                                       int nb, int frames, float* out, int row_offset, long long na_total, unsigned long long* row_key,
                                       float* row_max, long long* row_arg, float* col_max, long long* col_arg,
                                       int world, int rank, void* const* windows, long long list_cap, unsigned int step, glenet_stream_t s) {
    const char* what = "glenet_boxes_iou_frames_assign_gpu";
    cudaStream_t st = (cudaStream_t)s;
    int rc = exchange_args_ok(what, world, rank, windows, frames, nb, na_total, row_offset, na);
    if (rc) return rc;
    if (a_frame_stride < 0 || b_frame_stride < 0) return fail(GLENET_EINVAL, "%s: negative stride", what);
    if (!row_max || !row_arg || !col_max || !col_arg || (!row_key && na > 0)) return fail(GLENET_EINVAL, "%s: null output pointer", what);
    const ExchangeLayout lay = exchange_layout(frames, nb, list_cap);
    const int par = (int)(step & 1u);
    auto at = [&](int p, size_t off) { return reinterpret_cast<unsigned char*>(windows[p]) + off; };
    IouLaunch L;
    L.frames = frames; L.stride_a = a_frame_stride; L.stride_b = b_frame_stride; L.stride_out = (long long)na * nb;
    // key slots of this parity: [source rank][frames * nb]; ours (slot `rank`) is the local accumulator of the tile kernel
    L.row_key = row_key; L.col_key = reinterpret_cast<unsigned long long*>(at(rank, lay.off_col_key[par] + (size_t)rank * lay.key_slot));
    L.row_max = row_max; L.row_arg = row_arg;
    const bool rows_direct = nb <= IOU_TC_MAX;   // one column tile (pick_tiles never splits columns of a problem this tall)
    L.keys_prezeroed = true; L.dense_and_keys = out != nullptr;
    L.row_offset = row_offset; L.na_total = na_total;
    L.ex.world = world; L.ex.rank = rank; L.ex.step = step;
    L.ex.done = reinterpret_cast<unsigned int*>(at(rank, lay.off_done));
    for (int p = 0; p < world; ++p) {
        L.ex.col_key[p] = reinterpret_cast<unsigned long long*>(at(p, lay.off_col_key[par] + (size_t)rank * lay.key_slot));
        L.ex.flag[p] = reinterpret_cast<unsigned int*>(at(p, lay.off_flags_assign));
    }
    const bool launches_tiles = na > 0 && nb > 0 && frames > 0;
    L.ex.decode = 1;
    L.ex.slot_stride = (long long)(lay.key_slot / 8);
    L.ex.slots = reinterpret_cast<unsigned long long*>(at(rank, lay.off_col_key[par]));
    L.ex.flags_local = reinterpret_cast<const unsigned int*>(at(rank, lay.off_flags_assign));
    L.ex.status = reinterpret_cast<unsigned int*>(at(rank, lay.off_status));
    L.ex.col_max = col_max; L.ex.col_arg = col_arg;
    const IouPeers peers = L.ex;
    if (launches_tiles) {
        // the tile kernel only folds its maxima into the LOCAL keys (no epilogue, no fence); the exchange kernel behind it does the rest
        L.ex = IouPeers();
        rc = launch_iou_mode(mode, a, na, b, nb, out, st, what, L);
        if (rc) return rc;
    }
    const long long n_col = (long long)frames * nb;
    if (n_col > 0) {
        cudaLaunchConfig_t cfg = {};
        long long ctas = (n_col + 255) / 256;          // one key per thread: the push and the decode are latency chains (NVLink / L2 round trips)
        cfg.gridDim = dim3((unsigned)(ctas > 16 ? 16 : ctas)); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, exchange_assign_kernel, peers, (int)n_col);
        if (e != cudaSuccess) { snprintf(last_error_buf(), 512, "%s: launch failed: %s", what, cudaGetErrorString(e)); return -(int)e; }
        rc = check_launch(what);
        if (rc) return rc;
    }
    // row keys only exist when the columns are split over several tiles (nb > 128): decode them (the column part is done)
    if (rows_direct || !launches_tiles) return GLENET_OK;
    {
        const long long n_row = (long long)frames * na;
        long long blocks = (n_row + 255) / 256;
        if (blocks > 296) blocks = 296;
        exchange_decode_kernel<<<(unsigned)blocks, 256, 0, st>>>(row_key, n_row, nullptr, 0, row_max, row_arg, nullptr, nullptr, nullptr, 1, 0u, nullptr);
        return check_launch(what);
    }
}

int glenet_boxes_iou_frames_gather_gpu(int mode, const float* a, long long a_frame_stride, int na, const float* b, long long b_frame_stride,
                                       int nb, int frames, float* out_full, int zero_fill, int row_offset, long long na_total,
                                       int world, int rank, void* const* windows, long long list_cap, unsigned int step, glenet_stream_t s) {
    const char* what = "glenet_boxes_iou_frames_gather_gpu";
    cudaStream_t st = (cudaStream_t)s;
    int rc = exchange_args_ok(what, world, rank, windows, frames, nb, na_total, row_offset, na);
    if (rc) return rc;
    if (a_frame_stride < 0 || b_frame_stride < 0) return fail(GLENET_EINVAL, "%s: negative stride", what);
    if (list_cap <= 0) return fail(GLENET_EINVAL, "%s: the coordinate lists need a capacity", what);
    const long long out_elems = (long long)frames * na_total * nb;
    if (out_elems == 0) return GLENET_OK;
    if (!out_full) return fail(GLENET_EINVAL, "%s: null output pointer", what);
    const ExchangeLayout lay = exchange_layout(frames, nb, list_cap);
    const int par = (int)(step & 1u);
    auto at = [&](int p, size_t off) { return reinterpret_cast<unsigned char*>(windows[p]) + off; };
    if (zero_fill < 0 || zero_fill > 3) return fail(GLENET_EINVAL, "%s: zero_fill must be 0..3", what);
    const bool phase_kernel = zero_fill != 3, phase_scatter = zero_fill != 2;
    if (zero_fill == 1) {
        cudaError_t e = cudaMemsetAsync(out_full, 0, sizeof(float) * (size_t)out_elems, st);
        if (e != cudaSuccess) return fail(-(int)e, "%s: cudaMemsetAsync failed", what);
    }
    IouLaunch L;
    L.frames = frames; L.stride_a = a_frame_stride; L.stride_b = b_frame_stride;
    L.row_offset = row_offset; L.na_total = na_total;
    L.sp_cap = list_cap;
    L.sp_count = reinterpret_cast<unsigned long long*>(at(rank, lay.off_count));
    L.sp_idx = reinterpret_cast<long long*>(at(rank, lay.off_idx[par])) + (size_t)rank * list_cap;
    L.sp_val = reinterpret_cast<float*>(at(rank, lay.off_val[par])) + (size_t)rank * list_cap;
    L.ex.world = world; L.ex.rank = rank; L.ex.step = step;
    L.ex.done = reinterpret_cast<unsigned int*>(at(rank, lay.off_done)) + 1;
    for (int p = 0; p < world; ++p) {
        L.ex.idx[p] = reinterpret_cast<long long*>(at(p, lay.off_idx[par])) + (size_t)rank * list_cap;
        L.ex.val[p] = reinterpret_cast<float*>(at(p, lay.off_val[par])) + (size_t)rank * list_cap;
        L.ex.cnt[p] = reinterpret_cast<unsigned long long*>(at(p, lay.off_cnt[par])) + rank;
        L.ex.flag[p] = reinterpret_cast<unsigned int*>(at(p, lay.off_flags_gather));
    }
    if (!phase_kernel) {
    } else if (na > 0 && nb > 0 && frames > 0) {
        rc = launch_iou_mode(mode, a, na, b, nb, nullptr, st, what, L);
        if (rc) return rc;
    } else {
        cudaError_t e = cudaMemsetAsync(L.sp_count, 0, sizeof(unsigned long long), st);
        if (e != cudaSuccess) return fail(-(int)e, "%s: cudaMemsetAsync failed", what);
        if (world > 1) {
            exchange_signal_kernel<<<1, 32, 0, st>>>(L.ex, true);
            rc = check_launch(what);
            if (rc) return rc;
        }
    }
    if (!phase_scatter) return GLENET_OK;
    // one rank: the kernel's own counter is the list length; several: the per-source lengths the peers' last CTAs published
    const unsigned long long* cnt = world > 1 ? reinterpret_cast<const unsigned long long*>(at(rank, lay.off_cnt[par])) : L.sp_count;
    const long long* idx = reinterpret_cast<const long long*>(at(rank, lay.off_idx[par])) + (world > 1 ? 0 : (size_t)rank * list_cap);
    const float* val = reinterpret_cast<const float*>(at(rank, lay.off_val[par])) + (world > 1 ? 0 : (size_t)rank * list_cap);
    exchange_scatter_kernel<<<296, 256, 0, st>>>(out_full, idx, val, cnt, list_cap, out_elems,
                                                 reinterpret_cast<const unsigned int*>(at(rank, lay.off_flags_gather)), world, step,
                                                 reinterpret_cast<unsigned int*>(at(rank, lay.off_status)));
    return check_launch(what);
}

// keys of glenet_boxes_iou_frames_max_gpu -> (max, argmax) vectors in one launch (and the key buffers are left zeroed)
int glenet_iou_keys_decode_gpu(unsigned long long* row_key, long long n_row, unsigned long long* col_key, long long n_col,
                               float* row_max, long long* row_arg, float* col_max, long long* col_arg, glenet_stream_t s) {
    const char* what = "glenet_iou_keys_decode_gpu";
    if (n_row < 0 || n_col < 0) return fail(GLENET_EINVAL, "%s: negative count", what);
    if (n_row + n_col == 0) return GLENET_OK;
    if ((n_row && (!row_key || !row_max || !row_arg)) || (n_col && (!col_key || !col_max || !col_arg))) return fail(GLENET_EINVAL, "%s: null pointer", what);
    long long blocks = (n_row + n_col + 255) / 256;
    if (blocks > 296) blocks = 296;
    exchange_decode_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(row_key, n_row, col_key, n_col, row_max, row_arg, col_max, col_arg, nullptr, 1, 0u, nullptr);
    return check_launch(what);
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
    const int nbatches = (na + AL_PAIRS - 1) / AL_PAIRS;
    const int nb_max = (AL_PAIRS - 1) / group + 2;   // b boxes a batch of AL_PAIRS consecutive pairs can meet
    const size_t smem = sizeof(AlignedSmem) + sizeof(float) * 2 * (size_t)nb_max * (BPS + 1);
    const FastDiv gdiv = fastdiv_make((unsigned int)group);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = nbatches < sms * AL_CTAS_PER_SM ? nbatches : sms * AL_CTAS_PER_SM;
    cudaStream_t st = (cudaStream_t)s;
    int rc;
    // (the opt-in shared-memory size is a per-device attribute and cheap to set: no cache)
    if (mode == 0) { auto k = iou_aligned_kernel<MODE_OVERLAP, true>; rc = set_smem(k, smem, what); if (rc) return rc; k<<<grid, AL_THREADS, smem, st>>>(a, na, b, gdiv, out, nbatches, nb_max); }
    else if (mode == 1) { auto k = iou_aligned_kernel<MODE_IOU_BEV, true>; rc = set_smem(k, smem, what); if (rc) return rc; k<<<grid, AL_THREADS, smem, st>>>(a, na, b, gdiv, out, nbatches, nb_max); }
    else { auto k = iou_aligned_kernel<MODE_IOU3D, true>; rc = set_smem(k, smem, what); if (rc) return rc; k<<<grid, AL_THREADS, smem, st>>>(a, na, b, gdiv, out, nbatches, nb_max); }
    return check_launch(what);
}

}  // extern "C"
