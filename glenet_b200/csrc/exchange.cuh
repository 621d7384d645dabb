// Multi-GPU exchange for the row-sharded IoU sweep (one process per GPU, all GPUs of one NVLink / NVSwitch box).
//
// The reference never shards its geometry ops; what it offers the sharded path is the
// consumer: the anchor target assigner needs, per frame, the row maxima of its own rows and the COLUMN maxima + first
// row index over all rows (pcdet/models/dense_heads/target_assigner/axis_aligned_target_assigner.py:141-165).
//
// No NCCL on the data path.  Every rank owns an "exchange window" -- one cudaMalloc'ed block that the other ranks map
// through CUDA IPC (glenet_symm_*) -- and the IoU tile kernel itself writes into the peers' windows over NVLink:
//   * assign : column keys (value bits << 32 | ~row) are max-reduced locally (L2 atomics); the last CTA of each rank's
//              kernel then copies the finished key vector into ITS slot of every peer's window with plain 16-byte stores
//              over NVLink (remote atomics were measured at ~3 ns each, serialised: 90 us for 8 ranks x 12 800 keys), and a
//              small decode kernel waits for every rank's flag, takes the maximum over the `world` slots and turns keys
//              into (max, argmax);
//   * gather : > 99 % of an anchor sweep is exactly +0.0, so replicating the matrix on every rank does not need the
//              1.2 GB all-gather: every rank zero-fills its own copy at HBM speed and only the non-zero elements travel,
//              as (flat index, value) entries stored straight into the peers' windows from the clip epilogue; a scatter
//              kernel applies them once the flags are up.
// Buffers that a peer may still be writing for step s+1 while this rank consumes step s are double-buffered by step
// parity; flags are monotonic step numbers, never reset.  (Why two buffers are enough: a rank launches step s+2 only
// after its own consumer of step s+1 has run, and that consumer waited for every peer's step-s+1 kernel, which the peer
// launched after ITS consumer of step s.)
#pragma once
#include "common.cuh"

namespace glenet {

// The peers' exchange windows as the IoU tile kernel sees them (device pointers mapped through CUDA IPC).
struct IouPeers {
    int world, rank;
    unsigned int step;                                        // value published in the flags when this launch's pushes have landed
    unsigned int* done;                                       // local counter of finished CTAs (the last one pushes and signals)
    unsigned long long* col_key[GLENET_MAX_PEERS];            // OUR (frames, nb) key slot in every rank's window (this step's parity); [rank] = the local accumulator
    unsigned int* flag[GLENET_MAX_PEERS];                     // flag array of every rank; entry [rank] is ours to write
    // assign: the exchange kernel (exchange_assign_kernel) pushes the finished local keys, waits for the peers' flags and turns
    // the key slots of the LOCAL window into (max, first row)
    int decode; long long slot_stride;                         // slot_stride in keys
    unsigned long long* slots;                                 // slot 0 of this parity in the local window
    const unsigned int* flags_local; unsigned int* status;
    float* col_max; long long* col_arg;
    long long* idx[GLENET_MAX_PEERS]; float* val[GLENET_MAX_PEERS];   // our segment of every rank's coordinate-list window
    unsigned long long* cnt[GLENET_MAX_PEERS];                // ... and the slot for its length
};

struct ExchangeLayout {
    size_t off_flags_assign, off_flags_gather, off_status, off_done, off_count, off_col_key[2], off_cnt[2], off_idx[2], off_val[2];
    size_t bytes, key_slot;
    long long cap;       // list entries per (parity, source rank)
    long long nkeys;     // frames * nb
};

__host__ __device__ inline ExchangeLayout exchange_layout(int frames, int nb, long long cap) {
    ExchangeLayout l;
    size_t off = 0;
    l.nkeys = (long long)frames * nb;
    l.cap = cap;
    l.off_flags_assign = off; off += 64;      // [GLENET_MAX_PEERS] u32, padded
    l.off_flags_gather = off; off += 64;
    l.off_status = off; off += 64;            // u32 error bits (1 = timeout waiting for a peer, 2 = list overflow)
    l.off_done = off; off += 64;              // [2] u32: CTAs finished (assign, gather)
    l.off_count = off; off += 64;             // u64 local length counter of the gather kernel
    l.key_slot = ((size_t)l.nkeys * 8 + 255) / 256 * 256;   // one (frames, nb) key vector; [GLENET_MAX_PEERS] of them per parity, indexed by SOURCE rank
    for (int p = 0; p < 2; ++p) { l.off_col_key[p] = off; off += l.key_slot * GLENET_MAX_PEERS; }
    for (int p = 0; p < 2; ++p) { l.off_cnt[p] = off; off += 256; }   // [GLENET_MAX_PEERS] u64
    for (int p = 0; p < 2; ++p) { l.off_idx[p] = off; off += ((size_t)GLENET_MAX_PEERS * cap * 8 + 255) / 256 * 256; }
    for (int p = 0; p < 2; ++p) { l.off_val[p] = off; off += ((size_t)GLENET_MAX_PEERS * cap * 4 + 255) / 256 * 256; }
    l.bytes = off;
    return l;
}

enum { EX_STATUS_TIMEOUT = 1, EX_STATUS_OVERFLOW = 2 };

// Wait until every rank's flag has reached `step` (monotonic counters; wrap-safe compare).  One polling thread per
// CTA; ~4 s of polling then give up and record it (a peer that died must not hang this GPU for good).
__device__ __forceinline__ void exchange_wait_flags(const unsigned int* flags, int world, unsigned int step, unsigned int* status) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int src = 0; src < world; ++src) {
            for (;;) {
                unsigned int v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + src) : "memory");
                if ((int)(v - step) >= 0) break;
                if (clock64() - t0 > 8000000000LL) { atomicOr(status, (unsigned int)EX_STATUS_TIMEOUT); break; }
                __nanosleep(200);
            }
        }
    }
    __syncthreads();
}

// keys -> (max, argmax) for the rows of this rank and for all columns; the key buffers are zeroed for their next use.
// key = (IEEE bits of the maximum << 32) | (0xffffffff - first index); 0 = nothing non-zero: max 0, argmax 0.
__global__ void __launch_bounds__(256)
exchange_decode_kernel(unsigned long long* __restrict__ row_key, long long n_row, unsigned long long* __restrict__ col_key, long long n_col,
                       float* __restrict__ row_max, long long* __restrict__ row_arg, float* __restrict__ col_max, long long* __restrict__ col_arg,
                       const unsigned int* flags, int world, unsigned int step, unsigned int* status, int rank = 0, long long slot_stride = 0) {
    // col_key: slot 0 of this parity; slot r (at + r * slot_stride keys) holds rank r's finished keys, slot `rank` the local ones
    if (world > 1) exchange_wait_flags(flags, world, step, status);
    const long long gsz = (long long)gridDim.x * blockDim.x, g0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = g0; i < n_row; i += gsz) {
        const unsigned long long k = row_key[i];
        row_max[i] = __uint_as_float((unsigned int)(k >> 32));
        row_arg[i] = k ? (long long)(0xffffffffu - (unsigned int)k) : 0;
        if (k) row_key[i] = 0ull;
    }
    for (long long i = g0; i < n_col; i += gsz) {
        unsigned long long k = 0ull;
        for (int r = 0; r < world; ++r) {                  // peers' stores land in L2
            unsigned long long* slot = col_key + (size_t)r * slot_stride + i;
            const unsigned long long kr = __ldcg(slot);
            k = kr > k ? kr : k;
            // every slot goes back to zero: the local one is an atomic-max accumulator, and a later call with another
            // (frames, nb) lays the slots out differently.  (Safe: rank r writes this parity again only after it has seen
            // our flag of the NEXT step, which this stream raises after this kernel.)
            if (kr) *slot = 0ull;
        }
        col_max[i] = __uint_as_float((unsigned int)(k >> 32));
        col_arg[i] = k ? (long long)(0xffffffffu - (unsigned int)k) : 0;
        (void)rank;
    }
}

// The cross-rank half of the assigner step, launched as a programmatic dependent of the tile kernel (its launch latency is
// hidden; griddepcontrol.wait returns when the tile grid has completed and its key atomics are visible):
//   push   : the finished local column keys go into OUR slot of every peer's window -- plain 16-byte stores over NVLink;
//   signal : the last CTA to finish its share raises our flag in every window (monotonic step number);
//   decode : wait for every rank's flag, max over the `world` slots of the local window -> (col_max, col_arg); slots re-zeroed.
// Keeping this out of the tile kernel matters: a "last CTA" protocol inside it needs a __threadfence per CTA, which waits
// behind that CTA's 150 KB of result stores (measured: 8 % of the kernel's stall samples, ~25 us per 16-frame step).
__global__ void __launch_bounds__(256)
exchange_assign_kernel(const IouPeers ex, int nkeys) {
    // dependents first: the next step's tile kernel may take the SM slots the draining tile grid frees (it blocks in its own
    // griddepcontrol.wait until THIS grid has completed), so its CTA launch is off the critical path as in a dense sweep
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x, world = ex.world, rank = ex.rank;
    const long long gsz = (long long)gridDim.x * blockDim.x, g0 = (long long)blockIdx.x * blockDim.x + tid;
    __shared__ int s_last;
    if (world > 1) {
        const unsigned long long* mine = ex.col_key[rank];
        const long long pairs = nkeys >> 1;
        for (long long i = g0; i < pairs; i += gsz) {
            const ulonglong2 k = __ldcg(reinterpret_cast<const ulonglong2*>(mine) + i);
            for (int p = 0; p < world; ++p) if (p != rank) reinterpret_cast<ulonglong2*>(ex.col_key[p])[i] = k;
        }
        if ((nkeys & 1) && g0 == 0) {
            const unsigned long long k = __ldcg(mine + nkeys - 1);
            for (int p = 0; p < world; ++p) if (p != rank) ex.col_key[p][nkeys - 1] = k;
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(ex.done, 1u) == gridDim.x - 1u) ? 1 : 0;
        __syncthreads();
        if (s_last) {
            __threadfence_system();
            if (tid == 0) *ex.done = 0u;   // ready for the next step
            if (tid < world) *reinterpret_cast<volatile unsigned int*>(ex.flag[tid] + rank) = ex.step;
        }
        exchange_wait_flags(ex.flags_local, world, ex.step, ex.status);
    }
    for (long long i = g0; i < nkeys; i += gsz) {
        unsigned long long k = 0ull;
        for (int r = 0; r < world; ++r) {
            unsigned long long* slot = ex.slots + (size_t)r * ex.slot_stride + i;
            const unsigned long long kr = __ldcg(slot);
            k = kr > k ? kr : k;
            if (kr) *slot = 0ull;      // see exchange_decode_kernel: every slot goes back to zero
        }
        ex.col_max[i] = __uint_as_float((unsigned int)(k >> 32));
        ex.col_arg[i] = k ? (long long)(0xffffffffu - (unsigned int)k) : 0;
    }
}

// Apply the coordinate lists of all ranks (ours included) to the local, already zero-filled matrix.
__global__ void __launch_bounds__(256)
exchange_scatter_kernel(float* __restrict__ out, const long long* __restrict__ idx, const float* __restrict__ val,
                        const unsigned long long* cnt, long long cap, long long out_elems,
                        const unsigned int* flags, int world, unsigned int step, unsigned int* status) {
    if (world > 1) exchange_wait_flags(flags, world, step, status);
    const long long gsz = (long long)gridDim.x * blockDim.x, g0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int src = 0; src < world; ++src) {
        unsigned long long n = __ldcg(cnt + src);
        if (n > (unsigned long long)cap) { if (g0 == 0) atomicOr(status, (unsigned int)EX_STATUS_OVERFLOW); n = (unsigned long long)cap; }
        const long long* si = idx + (size_t)src * cap;
        const float* sv = val + (size_t)src * cap;
        for (long long k = g0; k < (long long)n; k += gsz) {
            const long long i = __ldcg(si + k);
            if ((unsigned long long)i < (unsigned long long)out_elems) out[i] = __ldcg(sv + k);
        }
    }
}

// A rank whose slab is empty still has to raise its flags (the peers wait for them).
__global__ void exchange_signal_kernel(IouPeers ex, bool gather) {
    const int t = threadIdx.x;
    // (assign: our key slots in the peers' windows are all zero already -- the decode kernels zero what they read)
    if (gather && t < ex.world) *reinterpret_cast<volatile unsigned long long*>(ex.cnt[t]) = 0ull;
    __threadfence_system();
    __syncthreads();
    if (t < ex.world) *reinterpret_cast<volatile unsigned int*>(ex.flag[t] + ex.rank) = ex.step;
}

}  // namespace glenet
