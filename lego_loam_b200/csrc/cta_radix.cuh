// cta_radix.cuh — stable LSD radix sort of up to 65536 (32-bit key, 16-bit value) pairs held in shared memory by
// ONE CTA: 8 bits per pass over the `nbits` significant key bits; per-warp digit counters + __match_any ranking keep
// equal keys in input order.  Used by the throughput voxel filter (voxel_small.cu: voxel index -> point) and by the
// query binning of the batched kNN (batch.cu: map cell -> query).
#pragma once
#include "common.cuh"

namespace llb {

// kin/vin hold the input; on return kin/vin (passed by reference, swapped per pass) point at the sorted arrays.
// s_wcnt: [THREADS/32][256] ints, s_base: [256] ints, s_scan: [33] ints of shared scratch.
// Must be called by all THREADS threads of the CTA; ends with a __syncthreads().
template <int THREADS, int ITEMS>
__device__ __forceinline__ void cta_radix_sort(unsigned *&kin, unsigned *&kout, unsigned short *&vin, unsigned short *&vout,
                                               int n, int nbits, int (*s_wcnt)[256], int *s_base, int *s_scan)
{
    constexpr int NW = THREADS / 32;
    constexpr int SUB = THREADS * ITEMS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int shift = 0; shift < nbits; shift += 8) {
        if (tid < 256) s_base[tid] = 0;
        __syncthreads();                                     // also orders the previous pass's scatter / the key pass
        for (int i = tid; i < n; i += THREADS) atomicAdd(&s_base[(kin[i] >> shift) & 255u], 1);
        __syncthreads();
        {
            const int v = tid < 256 ? s_base[tid] : 0;
            int total;
            const int ex = block_excl_scan(v, s_scan, total);
            if (tid < 256) s_base[tid] = ex;
        }
        for (int sub = 0; sub < n; sub += SUB) {
            for (int k = tid; k < NW * 256; k += THREADS) (&s_wcnt[0][0])[k] = 0;
            __syncthreads();
            unsigned key[ITEMS]; unsigned short val[ITEMS]; int rk[ITEMS]; unsigned dg[ITEMS];
#pragma unroll
            for (int r = 0; r < ITEMS; r++) {
                const int i = sub + w * (32 * ITEMS) + r * 32 + lane;
                const bool valid = i < n;
                key[r] = valid ? kin[i] : 0u;
                val[r] = valid ? vin[i] : (unsigned short)0;
                dg[r] = valid ? ((key[r] >> shift) & 255u) : (256u + lane);    // invalid lanes never match
            }
#pragma unroll
            for (int r = 0; r < ITEMS; r++) {
                const unsigned m = match_low_bits<9>(dg[r]);
                const int pr = __popc(m & lt);
                int cnt = 0;
                if (dg[r] < 256u) cnt = s_wcnt[w][dg[r]];
                rk[r] = cnt + pr;
                __syncwarp();
                if (dg[r] < 256u && pr == 0) s_wcnt[w][dg[r]] = cnt + __popc(m);
                __syncwarp();
            }
            __syncthreads();
            if (tid < 256) {                                 // digit tid: per-warp counts -> scatter offsets
                int run = s_base[tid];
#pragma unroll
                for (int k = 0; k < NW; k++) { const int c = s_wcnt[k][tid]; s_wcnt[k][tid] = run; run += c; }
                s_base[tid] = run;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < ITEMS; r++)
                if (dg[r] < 256u) {
                    const int pos = s_wcnt[w][dg[r]] + rk[r];
                    kout[pos] = key[r]; vout[pos] = val[r];
                }
            __syncthreads();
        }
        unsigned *tk = kin; kin = kout; kout = tk;
        unsigned short *tv = vin; vin = vout; vout = tv;
    }
    __syncthreads();
}

}  // namespace llb
