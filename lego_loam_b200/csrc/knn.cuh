// knn.cuh — warp-cooperative exact, radius-bounded 5-NN over the uniform-grid index (K3 phase A).
//
// The reference only consumes a 5-NN result when the 5th squared distance is below the gate
// (MO:1101, MO:1183: d2[4] < 1.0), so the five nearest points INSIDE the gate radius are all that
// is needed: if fewer than five candidates pass the gate the query is rejected, exactly as the
// reference rejects it.  One warp per query:
//   1. lanes 0..8 fetch the [begin, end) of the 9 contiguous cell runs around the query,
//   2. the first 32 candidates of every non-empty run are loaded by all lanes at once (coalesced
//      float4, all loads in flight before the first use), empty runs cost nothing (warp-uniform),
//   3. each lane computes flann::L2_Simple for its candidate; candidates inside the gate are
//      compacted into a per-warp shared-memory list with ballot/popc (typically 10-30 survive of
//      ~100),
//   4. five rounds of warp-wide arg-min over ONE 64-bit key per lane (distance bits << 32 |
//      original index: unsigned integer order == the oracle's lexicographic (distance, index) tie
//      rule) pick the result.  Lists longer than 32 (dense corner clusters) take a strided path.
#pragma once
#include "common.cuh"
#include "grid_index.cuh"

namespace llb {

constexpr int KNN_CAP = 128;                 // per-warp candidate list capacity (compressed when full)

// flann::L2_Simple<float>: sequential float sum of squared differences
__device__ __forceinline__ float l2_simple(float qx, float qy, float qz, const float4 &p)
{
    float diff = qx - p.x;
    float d = diff * diff;
    diff = qy - p.y; d += diff * diff;
    diff = qz - p.z; d += diff * diff;
    return d;
}

// strided selection of the 5 smallest keys of wkey[0..cnt) (cnt arbitrary).  Results to the out arrays;
// consumed entries are overwritten with ~0.  All lanes return the same values.
__device__ __forceinline__ void knn_select5_general(unsigned long long *wkey, const int *wpos, int cnt, int lane,
                                                    unsigned long long (&ok)[5], int (&op)[5])
{
#pragma unroll
    for (int r = 0; r < 5; r++) {
        unsigned long long best = ~0ull; int bslot = -1;
        for (int s = lane; s < cnt; s += 32) {
            const unsigned long long k = wkey[s];
            if (k < best) { best = k; bslot = s; }
        }
        const unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
        const unsigned mhi = __reduce_min_sync(FULL, hi);
        const unsigned clo = (hi == mhi) ? lo : 0xffffffffu;
        const unsigned mlo = __reduce_min_sync(FULL, clo);
        const bool win = (hi == mhi) && (lo == mlo) && bslot >= 0;
        const unsigned ball = __ballot_sync(FULL, win);
        const int src = ball ? (__ffs(ball) - 1) : 0;
        const int slot = __shfl_sync(FULL, bslot, src);
        ok[r] = ball ? (((unsigned long long)mhi << 32) | mlo) : ~0ull;
        op[r] = ball ? wpos[slot] : -1;
        if (ball && lane == src) wkey[slot] = ~0ull;
        __syncwarp();
    }
}

// Per-thread step (phase A1): the [begin, end) positions in m.sorted of the 9 contiguous cell runs
// (3 x-neighbours per (y,z) row) around the query; empty / out-of-grid rows give begin == end.
__device__ __forceinline__ void knn_ranges(const MapIndexView &m, float qx, float qy, float qz, int (&rb)[9], int (&re)[9])
{
    const GridDesc *g = m.desc;
    const int dimx = g->dim[0], dimy = g->dim[1], dimz = g->dim[2];
    const float inv = g->inv_cell;
    const int cx = grid_coord(qx, g->org[0], inv), cy = grid_coord(qy, g->org[1], inv), cz = grid_coord(qz, g->org[2], inv);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dimx - 1);
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int y = cy + (r % 3) - 1, z = cz + (r / 3) - 1;
        rb[r] = 0; re[r] = 0;
        if (y >= 0 && y < dimy && z >= 0 && z < dimz && x0 <= x1) {
            // cell_begin only holds valid offsets inside occupied rows; the four loads are issued together (no
            // dependent round trip) and the cell offsets are simply discarded when the row directory says "empty"
            const int ry = z * dimy + y;
            const int row = ry * dimx;
            const int r0 = __ldg(&m.row_begin[ry]), r1 = __ldg(&m.row_begin[ry + 1]);
            const int b = __ldg(&m.cell_begin[row + x0]), e = __ldg(&m.cell_begin[row + x1 + 1]);
            const bool occ = r1 > r0;
            rb[r] = occ ? b : 0;
            re[r] = occ ? e : 0;
        }
    }
}

// Warp step (phase A2): exact 5 nearest map points of (qx,qy,qz) among those with squared distance < max_sq
// inside the 9 runs [rbv[r], rev[r]) (warp-uniform values).
// Returns the number found (5, or < 5 => the reference's gate rejects the query); npos = positions in
// m.sorted, nd = squared distances, ni = original map indices, ascending in (distance, index).
// wkey / wpos: this warp's private shared-memory list (KNN_CAP entries).
__device__ __forceinline__ int knn5_warp(const MapIndexView &m, float qx, float qy, float qz, float max_sq, int lane,
                                         const int (&rbv)[9], const int (&rev)[9],
                                         unsigned long long *wkey, int *wpos, int (&npos)[5], float (&nd)[5], int (&ni)[5])
{
    float4 c[9];
#pragma unroll
    for (int r = 0; r < 9; r++)
        if (rev[r] > rbv[r] && rbv[r] + lane < rev[r]) c[r] = __ldg(&m.sorted[rbv[r] + lane]);
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;                                             // warp-uniform
    unsigned long long ck[5]; int cp[5];
    auto push = [&](bool ok, float d, int oi, int pos) {
        const unsigned mask = __ballot_sync(FULL, ok);
        if (mask) {                                          // warp-uniform
            if (cnt + 32 > KNN_CAP) {                        // list full: keep its best five and go on
                __syncwarp();
                knn_select5_general(wkey, wpos, cnt, lane, ck, cp);
#pragma unroll
                for (int k = 0; k < 5; k++) if (lane == k) { wkey[k] = ck[k]; wpos[k] = cp[k]; }
                cnt = 5;
                __syncwarp();
            }
            if (ok) {
                const int off = cnt + __popc(mask & lt);
                wkey[off] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)oi;
                wpos[off] = pos;
            }
            cnt += __popc(mask);
        }
    };
#pragma unroll
    for (int r = 0; r < 9; r++) {
        if (rev[r] > rbv[r]) {                               // warp-uniform: empty runs cost nothing
            const bool v = rbv[r] + lane < rev[r];
            const float d = v ? l2_simple(qx, qy, qz, c[r]) : 0.f;
            push(v && d < max_sq, d, v ? __float_as_int(c[r].w) : 0, rbv[r] + lane);
            for (int base = rbv[r] + 32; base < rev[r]; base += 32) {     // runs longer than a warp
                const int i = base + lane;
                const bool v2 = i < rev[r];
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v2) p = __ldg(&m.sorted[i]);
                const float d2 = l2_simple(qx, qy, qz, p);
                push(v2 && d2 < max_sq, d2, __float_as_int(p.w), i);
            }
        }
    }
    __syncwarp();
    if (cnt < 5) {
#pragma unroll
        for (int r = 0; r < 5; r++) { npos[r] = -1; ni[r] = -1; nd[r] = __int_as_float(0x7f800000); }
        return cnt;
    }
    if (cnt <= 32) {                                         // one candidate per lane, registers only
        unsigned long long key = lane < cnt ? wkey[lane] : ~0ull;
        const int pos = lane < cnt ? wpos[lane] : -1;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned mhi = __reduce_min_sync(FULL, hi);
            const unsigned clo = (hi == mhi) ? lo : 0xffffffffu;
            const unsigned mlo = __reduce_min_sync(FULL, clo);
            const bool win = (hi == mhi) && (lo == mlo);
            const int src = __ffs(__ballot_sync(FULL, win)) - 1;
            npos[r] = __shfl_sync(FULL, pos, src);
            nd[r] = __uint_as_float(mhi);
            ni[r] = (int)mlo;
            if (win) key = ~0ull;
        }
    } else {
        knn_select5_general(wkey, wpos, cnt, lane, ck, cp);
#pragma unroll
        for (int r = 0; r < 5; r++) { npos[r] = cp[r]; nd[r] = __uint_as_float((unsigned)(ck[r] >> 32)); ni[r] = (int)(unsigned)ck[r]; }
    }
    __syncwarp();
    return 5;
}

// Lean variant for the batched, high-occupancy kNN kernel (batch.cu): the same exact result as knn5_warp, but the
// cell runs are visited one after the other (bounds broadcast from lanes 0..8 on demand) instead of holding all nine
// first tiles in registers: ~half the registers, so 2-3x the warps per SM hide the load latency instead of ILP.
// b0 / e0: lane r < 9 holds the [begin, end) of run r.
__device__ __forceinline__ int knn5_warp_lean(const MapIndexView &m, float qx, float qy, float qz, float max_sq, int lane,
                                              int b0, int e0, unsigned long long *wkey, int *wpos,
                                              int (&npos)[5], float &d5th)
{
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;                                             // warp-uniform
#pragma unroll 1
    for (int r = 0; r < 9; r++) {
        const int rb = __shfl_sync(FULL, b0, r), re = __shfl_sync(FULL, e0, r);
#pragma unroll 1
        for (int base = rb; base < re; base += 32) {
            const int i = base + lane;
            const bool v = i < re;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v) p = __ldg(&m.sorted[i]);
            const float d = l2_simple(qx, qy, qz, p);
            const bool ok = v && d < max_sq;
            const unsigned mask = __ballot_sync(FULL, ok);
            if (mask) {                                      // warp-uniform
                if (cnt + 32 > KNN_CAP) {                    // list full: keep its best five and go on
                    unsigned long long ck[5]; int cp[5];
                    __syncwarp();
                    knn_select5_general(wkey, wpos, cnt, lane, ck, cp);
#pragma unroll
                    for (int k = 0; k < 5; k++) if (lane == k) { wkey[k] = ck[k]; wpos[k] = cp[k]; }
                    cnt = 5;
                    __syncwarp();
                }
                if (ok) {
                    const int off = cnt + __popc(mask & lt);
                    wkey[off] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
                    wpos[off] = i;
                }
                cnt += __popc(mask);
            }
        }
    }
    __syncwarp();
    d5th = -1.f;
    if (cnt < 5) {
#pragma unroll
        for (int r = 0; r < 5; r++) npos[r] = -1;
        return cnt;
    }
    if (cnt <= 32) {                                         // one candidate per lane, registers only
        unsigned long long key = lane < cnt ? wkey[lane] : ~0ull;
        const int pos = lane < cnt ? wpos[lane] : -1;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned mhi = __reduce_min_sync(FULL, hi);
            const unsigned clo = (hi == mhi) ? lo : 0xffffffffu;
            const unsigned mlo = __reduce_min_sync(FULL, clo);
            const bool win = (hi == mhi) && (lo == mlo);
            const int src = __ffs(__ballot_sync(FULL, win)) - 1;
            npos[r] = __shfl_sync(FULL, pos, src);
            if (r == 4) d5th = __uint_as_float(mhi);
            if (win) key = ~0ull;
        }
    } else {
        unsigned long long ck[5]; int cp[5];
        knn_select5_general(wkey, wpos, cnt, lane, ck, cp);
#pragma unroll
        for (int r = 0; r < 5; r++) npos[r] = cp[r];
        d5th = __uint_as_float((unsigned)(ck[4] >> 32));
    }
    __syncwarp();
    return 5;
}

}  // namespace llb
