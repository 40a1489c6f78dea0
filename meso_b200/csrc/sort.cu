// sort.cu -- stable LSD radix sort of (key, value) pairs with a DEVICE-SIDE element count.
//
// Replaces SortPlan<16,K,V> (UM/sort_meso.h:21-112, 350-429: 4-bit digits, 16 ballots per
// element, 16 single-block scans, 3 launches per 4 bits).  Blackwell version: 8-bit digits,
// warp-level MATCH.ANY ranking (one instruction instead of 16 ballots), each CTA owns a
// contiguous span of the input so the digit histogram is [256][#CTAs] with #CTAs a multiple
// of the SM count, and n is read from device memory so the caller never synchronises.
// Stability (needed for "ascending atom index inside a cell", UM/neighbor_meso.cu:588, and
// for tie order in MesoAtom::sort_local) comes from the warp-blocked element order.
#include "internal.h"
#include <cstdlib>
#include <utility>

namespace meso {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 8;                              // per thread per chunk
constexpr int SORT_CHUNK = SORT_THREADS * SORT_ITEMS;      // 2048

__device__ __forceinline__ void block_span(int n, int &beg, int &end)
{
    // contiguous span of whole chunks per CTA
    int nchunks = (n + SORT_CHUNK - 1) / SORT_CHUNK;
    int per = (nchunks + gridDim.x - 1) / gridDim.x;
    beg = min((int)blockIdx.x * per, nchunks) * SORT_CHUNK;
    end = min(min(((int)blockIdx.x + 1) * per, nchunks) * SORT_CHUNK, n);
    if (beg > n) beg = n;
}

template <typename K>
__global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const K *__restrict__ key, const int *__restrict__ d_n,
                                                             uint32_t *__restrict__ hist, int shift)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int beg, end;
    block_span(*d_n, beg, end);
    for (int i = beg + threadIdx.x; i < end; i += SORT_THREADS) atomicAdd(&h[(uint32_t)(key[i] >> shift) & 255u], 1u);
    __syncthreads();
    hist[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// One CTA per digit: exclusive scan of that digit's row hist[d][0..nblk) in place, row total to totals[d].
// (256 short independent scans instead of one long serial one.)
__global__ void __launch_bounds__(256) k_scan_rows(uint32_t *__restrict__ hist, uint32_t *__restrict__ totals, int nblk)
{
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    uint32_t *row = hist + (size_t)blockIdx.x * nblk;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 256) {
        int i = base + t;
        uint32_t v = i < nblk ? row[i] : 0u, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        uint32_t pre = 0;
#pragma unroll
        for (int ww = 0; ww < 8; ww++) pre += (ww < w) ? warp_sum[ww] : 0u;
        uint32_t excl = carry_s + pre + x - v;
        if (i < nblk) row[i] = excl;
        __syncthreads();
        if (t == 255) carry_s = excl + v;
        __syncthreads();
    }
    if (t == 0) totals[blockIdx.x] = carry_s;
}

template <typename K>
__global__ void __launch_bounds__(SORT_THREADS) k_radix_scatter(const K *__restrict__ key_in, const int *__restrict__ val_in,
                                                                K *__restrict__ key_out, int *__restrict__ val_out,
                                                                const int *__restrict__ d_n, const uint32_t *__restrict__ hist,
                                                                const uint32_t *__restrict__ totals, int shift)
{
    __shared__ uint32_t warp_hist[SORT_WARPS][256];
    __shared__ uint32_t gbase[256], cur_base[256];
    __shared__ uint32_t wtot[SORT_WARPS];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    {   // digit base = exclusive scan of the 256 digit totals (thread t owns digit t) + this CTA's offset inside the digit
        uint32_t v = totals[t], x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wtot[w] = x;
        __syncthreads();
        uint32_t pre = 0;
#pragma unroll
        for (int ww = 0; ww < SORT_WARPS; ww++) pre += (ww < w) ? wtot[ww] : 0u;
        gbase[t] = pre + x - v + hist[t * gridDim.x + blockIdx.x];
    }
    int beg, end;
    block_span(*d_n, beg, end);
    for (int chunk = beg; chunk < end; chunk += SORT_CHUNK) {
#pragma unroll
        for (int ww = 0; ww < SORT_WARPS; ww++) warp_hist[ww][t] = 0;
        __syncthreads();
        K k[SORT_ITEMS];
        int v[SORT_ITEMS];
        uint32_t rank[SORT_ITEMS], dig[SORT_ITEMS];
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; r++) {
            int i = chunk + w * (32 * SORT_ITEMS) + r * 32 + lane;     // warp-blocked: order = (warp, round, lane)
            bool ok = i < end;
            k[r] = ok ? key_in[i] : (K)0;
            v[r] = ok ? val_in[i] : 0;
            dig[r] = ok ? ((uint32_t)(k[r] >> shift) & 255u) : 0xFFFFFFFFu;
            uint32_t peers = __match_any_sync(0xffffffffu, dig[r]);
            int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (lane == leader && ok) {
                old = warp_hist[w][dig[r]];
                warp_hist[w][dig[r]] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[r] = old + __popc(peers & lt);
            __syncwarp();
        }
        __syncthreads();
        {   // thread t owns digit t: exclusive prefix over warps, then advance the CTA's global base
            uint32_t run = 0;
#pragma unroll
            for (int ww = 0; ww < SORT_WARPS; ww++) {
                uint32_t c = warp_hist[ww][t];
                warp_hist[ww][t] = run;
                run += c;
            }
            cur_base[t] = gbase[t];
            gbase[t] += run;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; r++) {
            if (dig[r] != 0xFFFFFFFFu) {
                uint32_t dst = cur_base[dig[r]] + warp_hist[w][dig[r]] + rank[r];
                key_out[dst] = k[r];
                val_out[dst] = v[r];
            }
        }
        __syncthreads();
    }
}

template <typename K>
static int sort_impl(meso_ctx *ctx, K *&key, int *&val, K *&key_alt, int *&val_alt, const int *d_n, size_t cap, int bits)
{
    const int nblk = grid_for(ctx, 4);
    if (!ctx->sort.hist.reserve((size_t)256 * nblk + 256)) { ctx->err = "sort: out of device memory"; return MESO_ECUDA; }
    (void)cap;
    for (int shift = 0; shift < bits; shift += 8) {
        k_radix_hist<K><<<nblk, SORT_THREADS, 0, LS(ctx->stream)>>>(key, d_n, ctx->sort.hist.p, shift);
        uint32_t *totals = ctx->sort.hist.p + (size_t)256 * nblk;
        k_scan_rows<<<256, 256, 0, LS(ctx->stream)>>>(ctx->sort.hist.p, totals, nblk);
        k_radix_scatter<K><<<nblk, SORT_THREADS, 0, LS(ctx->stream)>>>(key, val, key_alt, val_alt, d_n, ctx->sort.hist.p, totals, shift);
        K *tk = key; key = key_alt; key_alt = tk;
        int *tv = val; val = val_alt; val_alt = tv;
    }
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// ------------------------------------------------------------------ bucket sort of the reorder keys
// The reorder key is border bit | Morton(cell) | 12-bit sub-cell code (UM/atom_meso.cu:345-360): everything above the low
// `low_bits` bits names one of ~2 x ncell populated buckets of ~9 atoms.  Histogram, exclusive scan, slot claim and a small
// insertion sort per bucket on (key, original position) give exactly the permutation of the stable LSD radix sort -- one pass
// over the pairs instead of four or five (31-bit keys at 64^3, 37-bit keys at 200^3).
__global__ void __launch_bounds__(256) k_bucket_count(const uint64_t *__restrict__ key, const int *__restrict__ d_n, int *__restrict__ cnt, int low_bits)
{
    const int n = *d_n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(cnt + (size_t)(key[i] >> low_bits), 1);
}

__global__ void __launch_bounds__(256) k_bucket_fill(const uint64_t *__restrict__ key, const int *__restrict__ d_n, int *__restrict__ start1,
                                                     int *__restrict__ slot_idx, int low_bits)
{
    const int n = *d_n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        slot_idx[atomicAdd(start1 + (size_t)(key[i] >> low_bits), 1)] = i;
}

__global__ void __launch_bounds__(128) k_bucket_order(const int *__restrict__ start, int nbucket, int *__restrict__ slot_idx,
                                                      const uint64_t *__restrict__ key_in, const int *__restrict__ val_in,
                                                      uint64_t *__restrict__ key_out, int *__restrict__ val_out)
{
    constexpr int LOCAL = 24;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nbucket; b += gridDim.x * blockDim.x) {
        const int a = start[b], n = start[b + 1] - a;
        if (n == 0) continue;
        if (n <= LOCAL) {
            uint64_t k[LOCAL];
            int p[LOCAL];
            for (int q = 0; q < n; q++) {
                const int pi = slot_idx[a + q];
                const uint64_t ki = key_in[pi];
                int r = q;
                while (r > 0 && (k[r - 1] > ki || (k[r - 1] == ki && p[r - 1] > pi))) { k[r] = k[r - 1]; p[r] = p[r - 1]; r--; }
                k[r] = ki; p[r] = pi;
            }
            for (int q = 0; q < n; q++) { key_out[a + q] = k[q]; val_out[a + q] = val_in[p[q]]; }
        } else {                                            // a crowded bucket: the same insertion sort in place in global memory
            int *p = slot_idx + a;
            for (int q = 1; q < n; q++) {
                const int pi = p[q];
                const uint64_t ki = key_in[pi];
                int r = q;
                while (r > 0 && (key_in[p[r - 1]] > ki || (key_in[p[r - 1]] == ki && p[r - 1] > pi))) { p[r] = p[r - 1]; r--; }
                p[r] = pi;
            }
            for (int q = 0; q < n; q++) { key_out[a + q] = key_in[p[q]]; val_out[a + q] = val_in[p[q]]; }
        }
    }
}

// one pass; the result lands in the alternate buffers (the caller swaps)
static int bucket_sort(meso_ctx *ctx, const uint64_t *key, const int *val, uint64_t *key_out, int *val_out, const int *d_n, size_t cap, int bits,
                       int low_bits)
{
    const size_t nb = (size_t)1 << (bits - low_bits);
    SortScratch &s = ctx->sort;
    if (!s.bucket_cnt.reserve(nb + 8) || !s.bucket_start.reserve(nb + 8) || !s.slot_idx.reserve(cap + 8)) { ctx->err = "sort: out of device memory"; return MESO_ECUDA; }
    cudaStream_t st = ctx->stream;
    MESO_CUDA(cudaMemsetAsync(s.bucket_cnt.p, 0, sizeof(int) * nb, st));
    k_bucket_count<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(key, d_n, s.bucket_cnt.p, low_bits);
    if (int rc = scan_into(ctx, s.bucket_cnt.p, s.bucket_start.p, (int)nb)) return rc;
    k_bucket_fill<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(key, d_n, s.bucket_start.p + 1, s.slot_idx.p, low_bits);
    k_bucket_order<<<grid_for(ctx, 16), 128, 0, LS(st)>>>(s.bucket_start.p, (int)nb, s.slot_idx.p, key, val, key_out, val_out);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// Sorts in place from the caller's point of view: on return `key`/`val` hold the sorted
// pairs (the DevBuf pointers are swapped with the scratch buffers when the pass count is odd).
int sort_pairs_u64(meso_ctx *ctx, DevBuf<uint64_t> &key, DevBuf<int> &val, const int *d_n, size_t cap, int bits, int low_bits)
{
    // size the scratch by the LOGICAL capacity: buffers are swapped below, so sizing by key.cap would ratchet up
    if (!ctx->sort.key_alt.reserve(cap) || !ctx->sort.val_alt.reserve(cap)) { ctx->err = "sort: out of device memory"; return MESO_ECUDA; }
    uint64_t *k = key.p, *ka = ctx->sort.key_alt.p;
    int *v = val.p, *va = ctx->sort.val_alt.p;
    int rc;
    // reorder keys (12 low bits of sub-cell code under a cell code): one bucket pass; MESO_SORT_RADIX=1 keeps the radix passes (A/B)
    static const bool force_radix = getenv("MESO_SORT_RADIX") && getenv("MESO_SORT_RADIX")[0] == '1';
    if (low_bits > 0 && bits > low_bits && bits - low_bits <= 26 && !force_radix) {
        rc = bucket_sort(ctx, k, v, ka, va, d_n, cap, bits, low_bits);
        std::swap(k, ka); std::swap(v, va);
    } else
        rc = sort_impl<uint64_t>(ctx, k, v, ka, va, d_n, cap, bits);
    if (k != key.p) {   // odd number of passes: swap buffer ownership (capacities are compatible by construction)
        std::swap(key.p, ctx->sort.key_alt.p); std::swap(key.cap, ctx->sort.key_alt.cap);
        std::swap(val.p, ctx->sort.val_alt.p); std::swap(val.cap, ctx->sort.val_alt.cap);
    }
    return rc;
}

}  // namespace meso
