// neighbor.cu -- cell binning and the atomics-free build of the tile-transposed full neighbor table.
//
// Reference path:
//   MesoNeighbor::setup_bins          UM/neighbor_meso.cu:858-931   (cell lattice aligned to the sub-domain)
//   gpu_stencil_full_bin_3d           UM/neighbor_meso.cu:772-825   (<=27 cells, by (boundary flag, Morton))
//   binning_meso                      UM/neighbor_meso.cu:535-711   (cell id, radix sort, boundaries, expanded stencils)
//   gpu_build_neighbor_list<32,5>     UM/neigh_build_meso.cu:20-119 (warp per cell, 2 ballots per test, row-major rows)
//   gpu_join_neigh_list / gpu_transpose_neigh_list  UM/neigh_build_meso.cu:166-240 (two more passes over the table)
//
// Blackwell version (round 2).
//   * binning is a counting sort (histogram with hardware reductions -> exclusive scan -> claim a slot -> order each
//     cell's handful of atoms by index) instead of a 3-pass radix sort of (cell id, atom) pairs; the ordering pass also
//     writes the cell-ordered records {x, y, z, atom index} the build streams;
//   * one thread owns one local atom.  Cells are numbered x-fastest, so the 27 stencil cells are NINE contiguous runs of
//     records (three x-neighbors each): every lane walks its 9 runs with 4 independent 16-byte loads in flight, and lanes of
//     the same cell walk the same records in lockstep (one cache line per cell and load, not one per lane);
//   * in-range neighbors are staged in a per-lane queue in shared memory (column layout: bank == lane, no atomics, no
//     ballots; the rare entries past its depth go to a per-warp overflow area in global memory) and written out once,
//     when the row's totals are known.
// Row layout: [owned core][owned skin][other core][other skin], where "owned" marks the entries whose pair the force
// kernel evaluates from this row (ghost j, or (i+j) odd ? i<j : i>j) and core/skin is the reference's split at
// r <= r_n - skin (fp32, at build time).  The SET of every row, its counts and the core/skin split are the reference's;
// the order inside a segment is the walk order (x-rows of the stencil, ascending atom index inside a cell).  The reference's
// order (stencil cells by (boundary flag, Morton), ascending atom index inside a cell, skin entries reversed) is a pure
// function of (cell of j, j), so meso_export_pair_table rebuilds it on demand (k_canonical_rows) for the bit-exact parity
// checks.  Table offsets are 64-bit (the reference overflows int past 13.4 M atoms).
//
// Measured and dropped in this round (profiles/r02_s2_build_fine_lattice_ncu_summary.txt): a half-cell lattice with per-atom
// chord clipping of its rows tests 72 candidates per atom instead of 246, but every lane then reads its own records (16 cache
// lines per load instead of ~5) and the 36 row set-ups cost as much as the tests they save: 576 us against ~400 us.
#include "internal.h"
#include "device_math.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace meso {

// ------------------------------------------------------------------ per-cell stencil code rows
// stencil[cell][s] (s < 27): offset code (i+1) + 3*(j+1) + 9*(k+1) of the s-th neighbor cell; byte 31: count.
// slotrank[cell][code]: the inverse (position of the neighbor cell `code` in the stencil order, 0xff = outside the lattice).
__global__ void k_stencil_codes(unsigned char *__restrict__ stencil, unsigned char *__restrict__ slotrank, Box box)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= box.ncell) return;
    const int bx = cell % box.m[0], by = (cell / box.m[0]) % box.m[1], bz = cell / (box.m[0] * box.m[1]);
    uint32_t key[27];
    unsigned char code[27];
    int n = 0;
    for (int k = -1; k <= 1; k++)
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++) {
                int x = bx + i, y = by + j, z = bz + k;
                if (x < 0 || x >= box.m[0] || y < 0 || y >= box.m[1] || z < 0 || z >= box.m[2]) continue;
                uint32_t kk = morton3(x, y, z);
                if (x == 0 || x == box.m[0] - 1 || y == 0 || y == box.m[1] - 1 || z == 0 || z == box.m[2] - 1) kk += 0x80000000u;
                // insertion sort, ascending key (keys are unique)
                int p = n++;
                while (p > 0 && key[p - 1] > kk) { key[p] = key[p - 1]; code[p] = code[p - 1]; p--; }
                key[p] = kk; code[p] = (unsigned char)((i + 1) + 3 * (j + 1) + 9 * (k + 1));
            }
    unsigned char *row = stencil + (size_t)cell * 32, *inv = slotrank + (size_t)cell * 32;
    for (int s = 0; s < 32; s++) inv[s] = 0xff;
    for (int s = 0; s < 27; s++) { row[s] = s < n ? code[s] : 0; if (s < n) inv[code[s]] = (unsigned char)s; }
    row[31] = (unsigned char)n;
}

// ------------------------------------------------------------------ binning: counting sort into reference cells and fine cells
struct SoA3c { const double *c[3]; };


// cell coordinates of every atom (locals + ghosts) packed 10 bits per dimension, and the histogram of the cells
__global__ void __launch_bounds__(256) k_bin_count(SoA3c x, int *__restrict__ cellc, int *__restrict__ cell_cnt,
                                                   const Counts *__restrict__ cnt, Box box)
{
    const int nlocal = cnt->nlocal, nall = nlocal + cnt->nghost;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
        int b[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double xd = x.c[d][i];
            b[d] = clamp_rz(__fma_rn(xd - box.sublo[d], box.bininv[d], 1.0), 0, box.m[d]);   // UM/neighbor_meso.cu:410-412
            if (i >= nlocal) b[d] = (xd >= box.sublo[d]) ? (xd <= box.subhi[d] ? b[d] : box.m[d] - 1) : 0;   // :413-417
        }
        cellc[i] = b[0] | (b[1] << 10) | (b[2] << 20);
        atomicAdd(cell_cnt + b[0] + box.m[0] * (b[1] + b[2] * box.m[1]), 1);
    }
}

// exclusive scan of in[0..n) into out[1..n] (out[0] = 0): three launches, 4096 elements per CTA
constexpr int SCAN_T = 256, SCAN_I = 16, SCAN_TILE = SCAN_T * SCAN_I;

__global__ void __launch_bounds__(SCAN_T) k_scan_sums(const int *__restrict__ in, int *__restrict__ sums, int n)
{
    __shared__ int ws[SCAN_T / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_I;
    int s = 0;
#pragma unroll
    for (int u = 0; u < SCAN_I; u += 4) {
        if (base + u + 3 < n) { const int4 v = *reinterpret_cast<const int4 *>(in + base + u); s += v.x + v.y + v.z + v.w; }
        else for (int q = 0; q < 4; q++) if (base + u + q < n) s += in[base + u + q];
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < SCAN_T / 32; w++) t += ws[w]; sums[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(1024) k_scan_top(int *__restrict__ sums, int nblk)
{
    __shared__ int ws[32];
    __shared__ int carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + t;
        const int v = i < nblk ? sums[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) ws[w] = x;
        __syncthreads();
        if (w == 0) {
            const int s = ws[lane];
            int z = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
            ws[lane] = z - s;
        }
        __syncthreads();
        const int e = carry_s + ws[w] + x - v;
        if (i < nblk) sums[i] = e;
        __syncthreads();
        if (t == 1023) carry_s = e + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const int *__restrict__ in, int *__restrict__ out, const int *__restrict__ sums, int n)
{
    __shared__ int ws[SCAN_T / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int base = blockIdx.x * SCAN_TILE + t * SCAN_I;
    int v[SCAN_I], s = 0;
#pragma unroll
    for (int u = 0; u < SCAN_I; u += 4) {
        if (base + u + 3 < n) { const int4 q = *reinterpret_cast<const int4 *>(in + base + u); v[u] = q.x; v[u + 1] = q.y; v[u + 2] = q.z; v[u + 3] = q.w; }
        else for (int q = 0; q < 4; q++) v[u + q] = base + u + q < n ? in[base + u + q] : 0;
    }
#pragma unroll
    for (int u = 0; u < SCAN_I; u++) s += v[u];
    int x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[w] = x;
    __syncthreads();
    int pre = sums[blockIdx.x] + x - s;
#pragma unroll
    for (int ww = 0; ww < SCAN_T / 32; ww++) pre += ww < w ? ws[ww] : 0;
    if (blockIdx.x == 0 && t == 0) out[0] = 0;
#pragma unroll
    for (int u = 0; u < SCAN_I; u++) {
        if (base + u < n) out[base + u + 1] = pre;          // out[i + 1] = sum of in[0..i): becomes start[i + 1] once cell i is filled
        pre += v[u];
    }
}

// every atom claims a slot of its cell: start1 = cell_start + 1 holds the exclusive prefix before the kernel and the
// inclusive one (= the final cell_start of the next cell) after it
__global__ void __launch_bounds__(256) k_bin_fill(const int *__restrict__ cellc, int *__restrict__ start1, int *__restrict__ cell_atoms,
                                                  const Counts *__restrict__ cnt, int m0, int m1)
{
    const int nall = cnt->nlocal + cnt->nghost;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
        const int cc = cellc[i];
        const int c = (cc & 1023) + m0 * (((cc >> 10) & 1023) + m1 * (cc >> 20));
        cell_atoms[atomicAdd(start1 + c, 1)] = i;
    }
}

// atoms of a cell in ascending index (the order the reference's stable sort of (cell id, atom) leaves, UM/neighbor_meso.cu:588)
// and the cell-ordered copy of the packed coordinates, {x, y, z, bits(atom index)}, that the build kernel streams
__global__ void __launch_bounds__(128) k_cell_order(const int *__restrict__ cell_start, int *__restrict__ cell_atoms,
                                                    const float4 *__restrict__ coord4, float4 *__restrict__ cell_xyzj, int ncell)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int a = cell_start[c], n = cell_start[c + 1] - a;
    int *p = cell_atoms + a;
    for (int k = 1; k < n; k++) {
        const int v = p[k];
        int q = k;
        while (q > 0 && p[q - 1] > v) { p[q] = p[q - 1]; q--; }
        p[q] = v;
    }
    for (int k = 0; k < n; k++) {
        const int j = p[k];
        float4 v = coord4[j];
        v.w = __int_as_float(j);
        cell_xyzj[a + k] = v;
    }
}

// ------------------------------------------------------------------ build
constexpr uint32_t MJ = (1u << 27) - 1u;  // atom indices fit 27 bits in the hit queue (class bits above them); checked on the host
constexpr int NB_THREADS = 128;
constexpr int NB_BATCH = 4;       // candidates tested per iteration (independent 16-byte loads in flight)
constexpr int NQ = 40;            // per-lane queue slots in shared memory (5 KB per warp + 2.3 KB of run lists: 7 CTAs per SM)
constexpr int NQX = 88;           // per-lane overflow slots in global memory (rows of 41 .. 128 hits: a few entries of ~25 % of the rows at rho = 4)

__device__ __forceinline__ void sts_u32(unsigned addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(unsigned addr) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

// the entries of row i that its own lane evaluates in the pair-once force kernel (pair.cu): ghost partners always, local
// partners by the balanced rule (i+j) odd ? i<j : i>j (the mirror entry in row j is then "other")
__device__ __forceinline__ bool owns(int i, int j, int nlocal)
{
    const unsigned key = (unsigned)(j - i) * 0x80000001u;         // sign bit: d odd ? d > 0 : d < 0
    return (int)(key | (unsigned)(nlocal - 1 - j)) < 0;
}

__device__ __forceinline__ size_t slot(int i, int k, int n_col)
{
    return (size_t)((i & ~31) + (k & 31)) * (size_t)n_col + (size_t)((k >> 5) * 32 + (i & 31));
}

// Persistent grid (9 CTAs per SM).  Hit t of a lane lives in its shared-memory column for t < NQ, else in the warp's overflow
// area `scratch` (global memory, L2-resident).  Rows with more than NQ + NQX hits (or wider than the table) are left to the
// fall-back kernel.
__global__ void __launch_bounds__(NB_THREADS, 7) k_build_rows(const float4 *__restrict__ coord4, const int *__restrict__ cellc,
                                                           const int *__restrict__ cell_start, const float4 *__restrict__ cell_xyzj,
                                                           int *__restrict__ pair_count, int *__restrict__ owned_count,
                                                           int *__restrict__ core_split, int *__restrict__ pair_table,
                                                           Counts *__restrict__ cnt, int *__restrict__ fixup, uint32_t *__restrict__ scratch,
                                                           int n_col, float rc2_core, float rc2_tail, int m0, int m1, int m2)
{
    __shared__ uint32_t s_q[NB_THREADS / 32][NQ][32];
    __shared__ int2 s_runs[NB_THREADS / 32][9][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned qbase = (unsigned)__cvta_generic_to_shared(&s_q[wid][0][lane]);
    uint32_t *const xbase = scratch + ((size_t)blockIdx.x * (NB_THREADS / 32) + wid) * (NQX * 32) + lane;   // overflow entry t at xbase[t * 32]
    const int nlocal = cnt->nlocal;
    int worst = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; (i & ~31) < nlocal; i += gridDim.x * blockDim.x) {
        const bool active = i < nlocal;
        const float4 ci = coord4[active ? i : 0];
        const int cc = cellc[active ? i : 0];
        const int cx = cc & 1023, cy = (cc >> 10) & 1023, cz = cc >> 20;
        // ---- the 9 x-rows of the stencil: cells (cx-1 .. cx+1, y, z) are contiguous in the cell-ordered records
        const int xa = max(cx - 1, 0), xb = min(cx + 1, m0 - 1);
#pragma unroll
        for (int r = 0; r < 9; r++) {
            const int y = cy + r % 3 - 1, z = cz + r / 3 - 1;
            const bool ok = active && y >= 0 && y < m1 && z >= 0 && z < m2;
            const int *cs = cell_start + (ok ? (size_t)m0 * (size_t)(y + m1 * z) : (size_t)0);
            const int a = cs[ok ? xa : 0], b = cs[ok ? xb + 1 : 0];
            s_runs[wid][r][lane] = make_int2(a, b - a);
        }
        __syncwarp();
        // ---- flattened walk: every lane advances through its own concatenated candidate list; lanes of the same cell walk
        //      the same records in lockstep.  Hits are pushed as j | skin << 27 (the atom itself is dropped at write-out).
        int qn = 0, r = 0;
        int2 run = s_runs[wid][0][lane];
        int q = run.x, n = run.y;
        while (true) {
#pragma unroll
            for (int adv = 0; adv < 2; adv++)
                if (n <= 0 && r < 8) { r++; run = s_runs[wid][r][lane]; q = run.x; n = run.y; }
            if (!__any_sync(full, n > 0 || r < 8)) break;       // (three empty runs in a row: a non-periodic face)
            float4 v[NB_BATCH];
#pragma unroll
            for (int u = 0; u < NB_BATCH; u++) v[u] = cell_xyzj[n > u ? q + u : 0];
#pragma unroll
            for (int u = 0; u < NB_BATCH; u++) {
                const float dx = ci.x - v[u].x, dy = ci.y - v[u].y, dz = ci.z - v[u].z;
                const float dr2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));   // UM/neigh_build_meso.cu:86-89
                if (n > u && dr2 <= rc2_tail) {
                    const uint32_t en = (__float_as_uint(v[u].w) & MJ) | (dr2 <= rc2_core ? 0u : 1u << 27);
                    if (qn < NQ) sts_u32(qbase + (qn << 7), en);
                    else if (qn < NQ + NQX) xbase[(qn - NQ) << 5] = en;
                    qn++;
                }
            }
            q += NB_BATCH; n -= NB_BATCH;
        }
        bool bad = qn > NQ + NQX;
        if (bad) qn = 0;
        __syncwarp();
        // ---- classify: drop the atom itself (it is always a hit), decide ownership, count the four classes
        const int qmax = __reduce_max_sync(full, qn);
        uint32_t cls_cnt = 0;                                  // 4 x 8-bit counters: owned core, owned skin, other core, other skin
        for (int t = 0; t < qmax; t++) {
            if (t < qn) {
                const uint32_t en = t < NQ ? lds_u32(qbase + (t << 7)) : xbase[(t - NQ) << 5];
                const int j = (int)(en & MJ);
                uint32_t cls = 4;
                if (j != i) { cls = (owns(i, j, nlocal) ? 0u : 2u) | (en >> 27); cls_cnt += 1u << (8 * cls); }
                const uint32_t out = (uint32_t)j | (cls << 27);
                if (t < NQ) sts_u32(qbase + (t << 7), out); else xbase[(t - NQ) << 5] = out;
            }
        }
        const int n0 = cls_cnt & 255, n1 = (cls_cnt >> 8) & 255, n2 = (cls_cnt >> 16) & 255, n3 = cls_cnt >> 24;
        const int ntot = n0 + n1 + n2 + n3;
        bad = bad || ntot > n_col;
        if (active) {
            if (bad) { pair_count[i] = -1; atomicOr(fixup, 1); }
            else { pair_count[i] = ntot; owned_count[i] = n0 + n1; core_split[i] = n0 | (n2 << 16); }
        }
        worst = max(worst, bad ? 0 : ntot);
        // ---- write-out: the four segments, each in encounter order
        uint32_t pos = (uint32_t)n0 << 8 | (uint32_t)(n0 + n1) << 16 | (uint32_t)(n0 + n1 + n2) << 24;   // running position of each class
        int *row0 = pair_table + (size_t)(i & ~31) * (size_t)n_col + (i & 31);   // slot(i,k) = row0[(k&31)*n_col + (k>>5)*32]
        for (int t = 0; t < qmax; t++) {
            if (t < qn && !bad) {
                const uint32_t en = t < NQ ? lds_u32(qbase + (t << 7)) : xbase[(t - NQ) << 5];
                const uint32_t cls = en >> 27;
                if (cls < 4) {
                    const int sh = cls * 8;
                    const int k = (pos >> sh) & 255;
                    pos += 1u << sh;
                    row0[(k & 31) * n_col + (k >> 5) * 32] = (int)(en & MJ);
                }
            }
        }
        __syncwarp();
    }
    // diagnostics only
    worst = __reduce_max_sync(full, worst);
    if ((threadIdx.x & 31) == 0 && worst > 0) atomicMax(&cnt->max_pair, worst);
}

// Fall-back for the rows the kernel above marked (more than NQ + NQX hits, or wider than the table): plain walk of the same
// 9 runs, two passes (count the classes, then write).  Also the whole build when MESO_NB_SLOW=1 (A/B checks of the kernel above).
__global__ void __launch_bounds__(128) k_build_rows_slow(const float4 *__restrict__ coord4, const int *__restrict__ cellc,
                                                         const int *__restrict__ cell_start, const float4 *__restrict__ cell_xyzj,
                                                         int *__restrict__ pair_count, int *__restrict__ owned_count,
                                                         int *__restrict__ core_split, int *__restrict__ pair_table,
                                                         Counts *__restrict__ cnt, const int *__restrict__ fixup, int all_rows, int n_col,
                                                         float rc2_core, float rc2_tail, int m0, int m1, int m2)
{
    if (!all_rows && *fixup == 0) return;
    const int nlocal = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
        if (!all_rows && pair_count[i] >= 0) continue;
        const float4 ci = coord4[i];
        const int cc = cellc[i];
        const int cx = cc & 1023, cy = (cc >> 10) & 1023, cz = cc >> 20;
        const int xa = max(cx - 1, 0), xb = min(cx + 1, m0 - 1);
        int num[4] = {0, 0, 0, 0}, pos[4] = {0, 0, 0, 0};
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) { pos[0] = 0; pos[1] = num[0]; pos[2] = num[0] + num[1]; pos[3] = num[0] + num[1] + num[2]; }
            for (int z = max(cz - 1, 0); z <= min(cz + 1, m2 - 1); z++)
                for (int y = max(cy - 1, 0); y <= min(cy + 1, m1 - 1); y++) {
                    const int *cs = cell_start + (size_t)m0 * (size_t)(y + m1 * z);
                    for (int p = cs[xa]; p < cs[xb + 1]; p++) {
                        const float4 c2 = cell_xyzj[p];
                        const int j = __float_as_int(c2.w);
                        if (j == i) continue;
                        const float dx = ci.x - c2.x, dy = ci.y - c2.y, dz = ci.z - c2.z;
                        const float dr2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                        if (!(dr2 <= rc2_tail)) continue;
                        const int cls = (owns(i, j, nlocal) ? 0 : 2) | (dr2 <= rc2_core ? 0 : 1);
                        if (pass == 0) num[cls]++;
                        else { const int k = pos[cls]++; if (k < n_col) pair_table[slot(i, k, n_col)] = j; }
                    }
                }
        }
        int tot = num[0] + num[1] + num[2] + num[3];
        if (tot > n_col) {                                   // row wider than the table: flagged, clipped (UM/neigh_build_meso.cu:242-252 only printf's)
            atomicOr(&cnt->err, 2);
            tot = n_col; num[0] = min(num[0], n_col); num[1] = min(num[1], n_col - num[0]); num[2] = min(num[2], n_col - num[0] - num[1]);
        }
        pair_count[i] = tot; owned_count[i] = num[0] + num[1]; core_split[i] = num[0] | (num[2] << 16);
        atomicMax(&cnt->max_pair, tot);
    }
}

// ------------------------------------------------------------------ the reference's row order, on demand (exports)
// core entries in stencil order (neighbor cells by (boundary flag, Morton), ascending atom index inside a cell), then the
// skin entries in REVERSE stencil order (UM/neigh_build_meso.cu:58-117,166-200).  key = slot rank of j's cell << 27 | j.
__global__ void __launch_bounds__(128) k_canonical_rows(const int *__restrict__ cellc, const unsigned char *__restrict__ slotrank,
                                                        const int *__restrict__ pair_count, const int *__restrict__ owned_count,
                                                        const int *__restrict__ core_split, const int *__restrict__ pair_table,
                                                        int *__restrict__ out_table, const Counts *__restrict__ cnt, int n_col, int m0, int m1)
{
    const int nlocal = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
        const int np = pair_count[i], nown = owned_count[i], cs = core_split[i];
        const int n_oc = cs & 0xffff, n_nc = cs >> 16;
        const int cc = cellc[i];
        const int cx = cc & 1023, cy = (cc >> 10) & 1023, cz = cc >> 20;
        const unsigned char *inv = slotrank + (size_t)(cx + m0 * (cy + m1 * cz)) * 32;
        auto key_of = [&](int j) -> uint32_t {
            const int cj = cellc[j];
            const int code = ((cj & 1023) - cx + 1) + 3 * (((cj >> 10) & 1023) - cy + 1) + 9 * ((cj >> 20) - cz + 1);
            return ((uint32_t)inv[code] << 27) | (uint32_t)j;
        };
        // insertion sort of one class straight into the output row: class 0 = core -> [0, ncore) ascending keys,
        // class 1 = skin -> [ncore, np) descending keys
        const int ncore = n_oc + n_nc;
        for (int cls = 0; cls < 2; cls++) {
            const int dst0 = cls ? ncore : 0;
            int cntk = 0;
            for (int seg = 0; seg < 2; seg++) {
                // the two segments of this class in the production row: owned part, then other part
                const int a = cls == 0 ? (seg == 0 ? 0 : nown) : (seg == 0 ? n_oc : nown + n_nc);
                const int b = cls == 0 ? (seg == 0 ? n_oc : nown + n_nc) : (seg == 0 ? nown : np);
                for (int k = a; k < b; k++) {
                    const int j = pair_table[slot(i, k, n_col)];
                    const uint32_t key = key_of(j);
                    int p = cntk++;
                    while (p > 0) {
                        const int jp = out_table[slot(i, dst0 + p - 1, n_col)];
                        const uint32_t kp = key_of(jp);
                        if (cls == 0 ? kp > key : kp < key) { out_table[slot(i, dst0 + p, n_col)] = jp; p--; } else break;
                    }
                    out_table[slot(i, dst0 + p, n_col)] = j;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ host drivers
int launch_setup_bins(meso_ctx *ctx)
{
    Box &box = ctx->box;
    // MesoNeighbor::setup_bins, UM/neighbor_meso.cu:858-916
    double dim[3], vol = 1.0;
    for (int d = 0; d < 3; d++) { dim[d] = box.subhi[d] - box.sublo[d]; vol *= dim[d]; }
    double dens = ctx->nlocal_host / vol;
    if (dens < 3) dens = 3;
    double enc = dens * (4.0 / 3.0 * 3.142 * pow(ctx->cutneighmax, 3.0));
    enc *= 4.0;
    enc = std::max(enc, 32.0);
    ctx->expected_neigh_count = enc;
    ctx->n_col = (((int)enc + 31) / 32) * 32;               // MesoNeighList::grow, UM/neigh_list_meso.cu:38-40
    double inv = 1.0 / (1.0 * ctx->cutneighmax);
    for (int d = 0; d < 3; d++) {
        box.m[d] = std::max((int)(dim[d] * inv), 1) + 2;
        box.binsize[d] = dim[d] / (box.m[d] - 2);
        box.bininv[d] = 1.0 / box.binsize[d];
        if (box.m[d] > 1023) { ctx->err = "setup_bins: more than 1023 cells per dimension"; return MESO_EINVAL; }
    }
    box.ncell = box.m[0] * box.m[1] * box.m[2];
    if (!ctx->stencil.reserve((size_t)box.ncell * 32) || !ctx->slotrank.reserve((size_t)box.ncell * 32) ||
        !ctx->cell_start.reserve((size_t)box.ncell + 2) || !ctx->cell_cnt.reserve((size_t)box.ncell + 8) ||
        !ctx->scan_sums.reserve(((size_t)box.ncell + SCAN_TILE - 1) / SCAN_TILE + 8)) {
        ctx->err = "setup_bins: out of device memory";
        return MESO_ECUDA;
    }
    k_stencil_codes<<<(box.ncell + 127) / 128, 128, 0, LS(ctx->stream)>>>(ctx->stencil.p, ctx->slotrank.p, box);
    MESO_CUDA(cudaGetLastError());
    ctx->bins_ready = true;
    return MESO_OK;
}

static void scan_into(meso_ctx *ctx, const int *in, int *out, int n)
{
    const int nblk = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_sums<<<nblk, SCAN_T, 0, LS(ctx->stream)>>>(in, ctx->scan_sums.p, n);
    k_scan_top<<<1, 1024, 0, LS(ctx->stream)>>>(ctx->scan_sums.p, nblk);
    k_scan_apply<<<nblk, SCAN_T, 0, LS(ctx->stream)>>>(in, out, ctx->scan_sums.p, n);
}

int launch_neighbor_build(meso_ctx *ctx)
{
    const Box &box = ctx->box;
    const int ncell = box.ncell;
    if (ctx->cap + 8 > (size_t)MJ) { ctx->err = "neighbor build: more than 2^27 atoms + ghosts on one GPU"; return MESO_EINVAL; }
    const int build_grid = ctx->sm_count * 7;               // persistent: 7 CTAs per SM (33 KB of queues + run lists each), one wave
    if (!ctx->cell_xyzj.reserve(ctx->cap + 8) || !ctx->owned_count.reserve(ctx->cap) || !ctx->core_split.reserve(ctx->cap) ||
        !ctx->nb_fixup.reserve(1) || !ctx->nb_scratch.reserve((size_t)build_grid * (NB_THREADS / 32) * NQX * 32)) {
        ctx->err = "neighbor: out of device memory";
        return MESO_ECUDA;
    }
    cudaStream_t st = ctx->stream;
    MESO_CUDA(cudaMemsetAsync(ctx->cell_cnt.p, 0, sizeof(int) * (size_t)ncell, st));
    MESO_CUDA(cudaMemsetAsync(ctx->nb_fixup.p, 0, sizeof(int), st));
    SoA3c x; for (int d = 0; d < 3; d++) x.c[d] = ctx->x[d].p;
    k_bin_count<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(x, ctx->cell_of.p, ctx->cell_cnt.p, ctx->d_counts, box);
    scan_into(ctx, ctx->cell_cnt.p, ctx->cell_start.p, ncell);
    k_bin_fill<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(ctx->cell_of.p, ctx->cell_start.p + 1, ctx->cell_atoms.p, ctx->d_counts, box.m[0], box.m[1]);
    k_cell_order<<<(ncell + 127) / 128, 128, 0, LS(st)>>>(ctx->cell_start.p, ctx->cell_atoms.p, ctx->coord4.p, ctx->cell_xyzj.p, ncell);
    const float rc2_core = (float)pow(ctx->cutneighmax - ctx->skin, 2.0);   // UM/neigh_build_meso.cu:296-297
    const float rc2_tail = (float)pow(ctx->cutneighmax, 2.0);
    const int slow_grid = std::max(1, std::min((int)((nlocal_bound(ctx) + 127) / 128) + 1, ctx->sm_count * 64));
    if (!ctx->nb_slow)
        k_build_rows<<<build_grid, NB_THREADS, 0, LS(st)>>>(ctx->coord4.p, ctx->cell_of.p, ctx->cell_start.p, ctx->cell_xyzj.p, ctx->pair_count.p,
                                                           ctx->owned_count.p, ctx->core_split.p, ctx->pair_table.p, ctx->d_counts,
                                                           ctx->nb_fixup.p, ctx->nb_scratch.p, ctx->n_col, rc2_core, rc2_tail, box.m[0], box.m[1],
                                                           box.m[2]);
    k_build_rows_slow<<<slow_grid, 128, 0, LS(st)>>>(ctx->coord4.p, ctx->cell_of.p, ctx->cell_start.p, ctx->cell_xyzj.p, ctx->pair_count.p,
                                                     ctx->owned_count.p, ctx->core_split.p, ctx->pair_table.p, ctx->d_counts, ctx->nb_fixup.p,
                                                     ctx->nb_slow ? 1 : 0, ctx->n_col, rc2_core, rc2_tail, box.m[0], box.m[1], box.m[2]);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// the table in the reference's row order, for meso_export_pair_table
int launch_canonical_rows(meso_ctx *ctx, int *out_table)
{
    k_canonical_rows<<<grid_for(ctx, 8), 128, 0, LS(ctx->stream)>>>(ctx->cell_of.p, ctx->slotrank.p, ctx->pair_count.p, ctx->owned_count.p,
                                                                   ctx->core_split.p, ctx->pair_table.p, out_table, ctx->d_counts, ctx->n_col,
                                                                   ctx->box.m[0], ctx->box.m[1]);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
