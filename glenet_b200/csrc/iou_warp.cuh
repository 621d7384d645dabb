// Warp-autonomous variant of the pairwise IoU kernel for large sweeps (included by iou.cu inside namespace glenet).
//
// iou_tile_kernel couples the 7 chain warps of a CTA with ~12 barriers per tile; at 4 CTAs/SM that leaves the SM with
// four phase chains and 44 % issue utilisation.  Here a CTA still owns a tile (512 rows x <= 128 columns) and the column
// side is shared -- staged, culled against the tile's row bounding box and prepared (BoxPre) by the whole CTA --, but
// after that every WARP owns 64 rows and runs its own chain without any CTA barrier: circle tests -> warp-private queue ->
// lazy prepare of its rows -> separating-axis filter -> clip -> results, plus its own slice of the zero fill as
// asynchronous bulk copies.  32 independent chains per SM instead of 4.
//
// Same arithmetic, culls and output modes as iou_tile_kernel (shared device functions); results are bit-identical.
#pragma once

constexpr int WK_THREADS = 256;
constexpr int WK_WARPS = WK_THREADS / 32 - 1;   // worker warps; the last warp of the CTA only feeds the zero fill
constexpr int WK_WORKERS = WK_WARPS * 32;
constexpr int WK_RW = 64;                   // rows per worker warp (two per lane)
constexpr int WK_TR = WK_WARPS * WK_RW;     // rows per CTA (448)
constexpr int WK_Q = 128;                   // circle-test survivors per warp and drain
constexpr int WK_CCH = 32;                  // active columns per chunk = column BoxPre records resident at a time
constexpr int WK_CTAS_PER_SM = 4;

struct WkWarp {
    float rpre[WK_RW * BPS];                // raw box, later BoxPre, of the warp's rows
    float qres[WK_Q];
    unsigned short queue[WK_Q], queue2[WK_Q];
    unsigned char plist[WK_RW];
};
struct __align__(128) WkSmem {
    float4 zero[IOU_ZBYTES / 16];
    float ccx[IOU_TC_MAX], ccy[IOU_TC_MAX], crad[IOU_TC_MAX];
    float cpre[WK_CCH * BPS];
    float red[WK_WARPS][5];
    unsigned char act[IOU_TC_MAX];
    int nact;
    int fill_done;                          // set by the fill warp once the tile's zeros have landed
    WkWarp w[WK_WARPS];
};

__device__ __forceinline__ void wk_sync() { asm volatile("bar.sync 1, %0;" :: "n"(WK_WORKERS) : "memory"); }   // worker warps only
__device__ __forceinline__ void wk_wait_fill(WkSmem& sm) {
    while (*reinterpret_cast<volatile int*>(&sm.fill_done) == 0) __nanosleep(200);
    __threadfence_block();
}

// Everything a warp needs to drain its queue; kept in one struct so that the drain is a plain function.
struct WkCtx {
    const float* A; const float* B; const float4* trigA;
    float* out; int nb; int r0; int c0; int cb;   // r0: first global row of this WARP, cb: first active-column index of the chunk
    long long frame_base; int frame;
};

template <int MODE, bool FMA>
__device__ __forceinline__ void wk_drain(WkSmem& sm, WkWarp& ws, const WkCtx& cx, const IouFrames& fr, int n, unsigned int need,
                                         unsigned int& done, bool& fill_pending, int lane) {
    const unsigned int lt = (1u << lane) - 1u;
    // ---- lazy prepare of the rows that have a queued pair and no record yet (lane owns rows lane and lane + 32)
    const unsigned int todo = need & ~done;
    const unsigned int m0 = __ballot_sync(0xffffffffu, todo & 1u), m1 = __ballot_sync(0xffffffffu, todo & 2u);
    const int n0 = __popc(m0), np = n0 + __popc(m1);
    if (todo & 1u) ws.plist[__popc(m0 & lt)] = (unsigned char)lane;
    if (todo & 2u) ws.plist[n0 + __popc(m1 & lt)] = (unsigned char)(lane + 32);
    done |= need;
    __syncwarp();
    for (int p0 = 0; p0 < np; p0 += 32) {
        if (p0 + lane < np) {
            const int r = ws.plist[p0 + lane];
            float* rec = ws.rpre + r * BPS;
            float raw[7];
#pragma unroll
            for (int f = 0; f < 7; ++f) raw[f] = rec[f];
            const float4 t4 = cx.trigA ? cx.trigA[cx.r0 + r] : device_trig(raw[6]);
            box_prepare<FMA, false>(raw, t4, rec);
        }
    }
    __syncwarp();
    // ---- separating-axis filter
    int n2 = 0;
    for (int q0 = 0; q0 < n; q0 += 32) {
        const int q = q0 + lane;
        unsigned int e = 0;
        bool keep = false;
        if (q < n) {
            e = ws.queue[q];
            const float* a = ws.rpre + (e >> 7) * BPS;
            const float* b = sm.cpre + ((int)(e & 127u) - cx.cb) * BPS;
            keep = !sat_separated(a, b);
            if (MODE == MODE_IOU3D && !keep) {   // 0 * NaN (see iou_tile_kernel)
                const float* boxa = cx.A + (size_t)(cx.r0 + (e >> 7)) * 7;
                const float* boxb = cx.B + (size_t)(cx.c0 + sm.act[e & 127u]) * 7;
                keep = !z_terms_finite(z_terms(boxa[2], boxa[5], a[BP_AREA])) || !z_terms_finite(z_terms(boxb[2], boxb[5], b[BP_AREA]));
            }
        }
        const unsigned int m = __ballot_sync(0xffffffffu, keep);
        if (keep) ws.queue2[n2 + __popc(m & lt)] = (unsigned short)e;
        n2 += __popc(m);
    }
    __syncwarp();
    // ---- clip
    for (int q0 = 0; q0 < n2; q0 += 32) {
        const int q = q0 + lane;
        if (q < n2) {
            const unsigned int e = ws.queue2[q];
            const float* a = ws.rpre + (e >> 7) * BPS;
            const float* b = sm.cpre + ((int)(e & 127u) - cx.cb) * BPS;
            ws.qres[q] = finish_pair<MODE>(a, b, box_overlap<FMA>(a, b), cx.A + (size_t)(cx.r0 + (e >> 7)) * 7,
                                           cx.B + (size_t)(cx.c0 + sm.act[e & 127u]) * 7);
        }
    }
    if (fill_pending) {   // the tile's zero fill has landed (flag set by the fill warp)
        wk_wait_fill(sm);
        fill_pending = false;
    }
    __syncwarp();
    // ---- results
    for (int q0 = 0; q0 < n2; q0 += 32) {
        const int q = q0 + lane;
        unsigned int e = 0;
        float v = 0.f;
        if (q < n2) { e = ws.queue2[q]; v = ws.qres[q]; }
        const unsigned int r = cx.r0 + (e >> 7), c = cx.c0 + sm.act[e & 127u];
        const bool nz = q < n2 && !(v == 0.f);
        if (fr.row_key) {
            if (nz) {
                const unsigned long long hi = (unsigned long long)__float_as_uint(v) << 32;
                atomicMax(fr.row_key + (size_t)cx.frame * fr.na + r, hi | (0xffffffffu - c));
                atomicMax(fr.col_key + (size_t)cx.frame * cx.nb + c, hi | (0xffffffffu - r));
            }
        }
        if (fr.sp_count) {
            const unsigned int m = __ballot_sync(0xffffffffu, nz);
            if (m) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(fr.sp_count, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0) + __popc(m & lt);
                if (nz && (long long)base < fr.sp_cap) { fr.sp_idx[base] = cx.frame_base + (long long)r * cx.nb + c; fr.sp_val[base] = v; }
            }
        } else if (!fr.row_key) {
            if (q < n2) cx.out[(size_t)r * cx.nb + c] = v;
        }
    }
    __syncwarp();
}

template <int MODE, bool FMA>
__global__ void __launch_bounds__(WK_THREADS, WK_CTAS_PER_SM)
iou_warp_kernel(const float* __restrict__ A, int na, const float* __restrict__ B, int nb,
                const float4* __restrict__ trigA, const float4* __restrict__ trigB,
                float* __restrict__ out, int TC, IouFrames fr) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WkSmem& sm = *reinterpret_cast<WkSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.z;
    A += (size_t)frame * fr.stride_a; B += (size_t)frame * fr.stride_b; out += (size_t)frame * fr.stride_out;
    const int r0 = blockIdx.y * WK_TR, c0 = blockIdx.x * TC;
    const int tr = min(WK_TR, na - r0), tc = min(TC, nb - c0);
    WkWarp& ws = sm.w[warp];
    const int wr0 = warp * WK_RW, wrows = max(0, min(WK_RW, tr - wr0));
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool no_matrix = iou_no_matrix(fr);
    if (tid == 0) { sm.fill_done = 0; sm.nact = 0; }
    __syncthreads();   // the only CTA-wide barrier: the flag is clear before anybody can set or poll it

    if (warp == WK_WARPS) {   // ---- the fill warp: zero block, bulk copies for the whole tile, completion flag
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (no_matrix) return;
        const bool vec = ((nb & 3) == 0) && ((c0 & 3) == 0) && ((tc & 3) == 0) && ((((uintptr_t)out) & 15) == 0);
        float* ot = out + (size_t)r0 * nb + c0;
        if (vec) {
#pragma unroll
            for (int k = 0; k < IOU_ZBYTES / 16 / 32; ++k) sm.zero[k * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            fence_proxy_async();
            __syncwarp();
            if (tc == nb) {
                const size_t total = (size_t)tr * nb * sizeof(float);
                char* dst = reinterpret_cast<char*>(ot);
                for (size_t off = (size_t)lane * IOU_ZBYTES; off < total; off += (size_t)32 * IOU_ZBYTES) {
                    const size_t left = total - off;
                    bulk_store(dst + off, sm.zero, (unsigned int)(left < (size_t)IOU_ZBYTES ? left : (size_t)IOU_ZBYTES));
                }
            } else {
                for (int r = lane; r < tr; r += 32) bulk_store(ot + (size_t)r * nb, sm.zero, (unsigned int)tc * sizeof(float));
            }
            bulk_commit();
            bulk_wait_all();
            fence_proxy_async();
        } else {
            const int npairs = tr * tc;
            for (int p = lane; p < npairs; p += 32) { const int r = p / tc; ot[(size_t)r * nb + (p - r * tc)] = 0.f; }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile int*>(&sm.fill_done) = 1;
        return;
    }

    // ---- prologue (worker warps): columns (centre + cull radius) together, each warp its own rows (raw box into the record slots)
    for (int c = tid; c < tc; c += WK_WORKERS) {
        const float* box = B + (size_t)(c0 + c) * 7;
        const float cx = box[0], cy = box[1], dx = box[3], dy = box[4];
        float rad = cull_radius(cx, cy, dx, dy);
        if (MODE == MODE_IOU3D && !z_terms_finite(z_terms(box[2], box[5], __fmul_rn(dx, dy)))) rad = CUDART_INF_F;
        sm.ccx[c] = cx; sm.ccy[c] = cy; sm.crad[c] = rad;
    }
    float4 rw[2];
    float minx = CUDART_INF_F, maxx = -CUDART_INF_F, miny = CUDART_INF_F, maxy = -CUDART_INF_F, maxr = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int r = lane + 32 * j;
        rw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < wrows) {
            const float* box = A + (size_t)(r0 + wr0 + r) * 7;
            float raw[7];
#pragma unroll
            for (int f = 0; f < 7; ++f) raw[f] = box[f];
            float* rec = ws.rpre + r * BPS;
#pragma unroll
            for (int f = 0; f < 7; ++f) rec[f] = raw[f];
            float rad = cull_radius(raw[0], raw[1], raw[3], raw[4]);
            if (MODE == MODE_IOU3D && !z_terms_finite(z_terms(raw[2], raw[5], __fmul_rn(raw[3], raw[4])))) rad = CUDART_INF_F;
            rw[j] = make_float4(raw[0], raw[1], rad, 0.f);
            minx = fminf(minx, raw[0]); maxx = fmaxf(maxx, raw[0]); miny = fminf(miny, raw[1]); maxy = fmaxf(maxy, raw[1]);
            maxr = (rad != rad) ? CUDART_INF_F : fmaxf(maxr, rad);
        }
    }
    minx = warp_min(minx); maxx = warp_max(maxx); miny = warp_min(miny); maxy = warp_max(maxy); maxr = warp_max(maxr);
    if (lane == 0) { sm.red[warp][0] = minx; sm.red[warp][1] = maxx; sm.red[warp][2] = miny; sm.red[warp][3] = maxy; sm.red[warp][4] = maxr; }
    wk_sync();
    bool fill_pending = !no_matrix;

    // ---- active columns of the tile (against the bounding box of all its rows)
    minx = sm.red[0][0]; maxx = sm.red[0][1]; miny = sm.red[0][2]; maxy = sm.red[0][3]; maxr = sm.red[0][4];
#pragma unroll
    for (int w = 1; w < WK_WARPS; ++w) {
        minx = fminf(minx, sm.red[w][0]); maxx = fmaxf(maxx, sm.red[w][1]);
        miny = fminf(miny, sm.red[w][2]); maxy = fmaxf(maxy, sm.red[w][3]); maxr = fmaxf(maxr, sm.red[w][4]);
    }
    for (int c = tid; c < tc; c += WK_WORKERS) {
        const float cx = sm.ccx[c], cy = sm.ccy[c];
        const float ddx = fmaxf(fmaxf(minx - cx, cx - maxx), 0.f), ddy = fmaxf(fmaxf(miny - cy, cy - maxy), 0.f);
        const float rr = maxr + sm.crad[c];
        const bool far = (cx == cx) && (cy == cy) && (ddx * ddx + ddy * ddy > rr * rr);
        if (!far) sm.act[atomicAdd(&sm.nact, 1)] = (unsigned char)c;
    }
    wk_sync();
    const int nact = sm.nact;

    WkCtx cx;
    cx.A = A; cx.B = B; cx.trigA = trigA; cx.out = out; cx.nb = nb; cx.r0 = r0 + wr0; cx.c0 = c0; cx.cb = 0;
    cx.frame_base = (long long)frame * na * nb; cx.frame = frame;
    unsigned int need = 0u, done = 0u;
    int qn = 0;
    const unsigned int lt = (1u << lane) - 1u;
    for (int cb = 0; cb < nact; cb += WK_CCH) {
        const int ncol = min(WK_CCH, nact - cb);
        if (cb > 0) wk_sync();   // every warp has clipped what referred to the previous chunk's column records
        if (tid < ncol) {
            const int c = sm.act[cb + tid];
            const float* box = B + (size_t)(c0 + c) * 7;
            float raw[7];
#pragma unroll
            for (int f = 0; f < 7; ++f) raw[f] = box[f];
            const float4 t4 = trigB ? trigB[c0 + c] : device_trig(raw[6]);
            box_prepare<FMA, false>(raw, t4, sm.cpre + tid * BPS);
        }
        wk_sync();
        if (wrows > 0) {
            cx.cb = cb;
            unsigned int m[2] = {0u, 0u};
#pragma unroll 4
            for (int k = 0; k < ncol; ++k) {
                const int c = sm.act[cb + k];
                const float ccx = sm.ccx[c], ccy = sm.ccy[c], cr = sm.crad[c];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float dx = rw[j].x - ccx, dy = rw[j].y - ccy, rr = rw[j].z + cr;
                    if (!(dx * dx + dy * dy > rr * rr)) m[j] |= 1u << k;   // NaN => not culled
                }
            }
            if (lane >= wrows) m[0] = 0u;
            if (lane + 32 >= wrows) m[1] = 0u;
            for (;;) {   // append; entries beyond the queue's capacity stay in the bitmask for the next round
                const int cnt = __popc(m[0]) + __popc(m[1]);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total == 0) break;
                int base = qn + incl - cnt;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    while (m[j] && base < WK_Q) {
                        const int k = __ffs(m[j]) - 1;
                        m[j] &= m[j] - 1;
                        ws.queue[base++] = (unsigned short)(((lane + 32 * j) << 7) | (cb + k));
                        need |= 1u << j;
                    }
                }
                __syncwarp();
                if (qn + total <= WK_Q) { qn += total; break; }
                wk_drain<MODE, FMA>(sm, ws, cx, fr, WK_Q, need, done, fill_pending, lane);
                qn = 0;
            }
            if (cb + WK_CCH < nact && qn) {   // the column records are about to be replaced
                wk_drain<MODE, FMA>(sm, ws, cx, fr, qn, need, done, fill_pending, lane);
                qn = 0;
            }
        }
    }
    if (wrows > 0 && qn) wk_drain<MODE, FMA>(sm, ws, cx, fr, qn, need, done, fill_pending, lane);
    (void)lt;
}
