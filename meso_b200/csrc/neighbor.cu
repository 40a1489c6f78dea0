// neighbor.cu -- cell binning and the atomics-free build of the tile-transposed full neighbor table.
//
// Reference path:
//   MesoNeighbor::setup_bins          UM/neighbor_meso.cu:858-931   (cell lattice aligned to the sub-domain)
//   gpu_stencil_full_bin_3d           UM/neighbor_meso.cu:772-825   (<=27 cells, by (boundary flag, Morton))
//   binning_meso                      UM/neighbor_meso.cu:535-711   (cell id, radix sort, boundaries, expanded stencils)
//   gpu_build_neighbor_list<32,5>     UM/neigh_build_meso.cu:20-119 (warp per cell, 2 ballots per test, row-major rows)
//   gpu_join_neigh_list / gpu_transpose_neigh_list  UM/neigh_build_meso.cu:166-240 (two more passes over the table)
//
// Blackwell version (round 2).  The reference tests every atom of the 27 stencil cells: 246 distance tests per atom for
// 36.8 stored neighbors at rho = 4, because its cells are as wide as the list cutoff r_n.  Here:
//   * binning is a counting sort (histogram with hardware reductions -> exclusive scan -> claim a slot -> order each
//     cell's handful of atoms by index) instead of a 3-pass radix sort of (cell id, atom) pairs;
//   * the same pass files every atom in a HALF-CELL lattice (2 x 2 x 2 fine cells per reference cell, x-fastest), whose
//     cell-ordered copy holds {x, y, z, atom index} records;
//   * one thread owns one local atom and walks the 6 x 6 fine rows of its 27 stencil cells; each row is clipped to the
//     chord of the cutoff sphere, so one run of CONTIGUOUS records replaces up to 6 fine cells and about 72 candidates are
//     tested instead of 246.  The runs of a warp's 32 atoms are staged in shared memory and walked with a predicated
//     (divergence-free) advance;
//   * in-range neighbors are staged in a per-lane queue in shared memory (column layout: bank == lane, no atomics, no
//     ballots) and written out once, when the row's totals are known.
// Row layout: [owned core][owned skin][other core][other skin], where "owned" marks the entries whose pair the force
// kernel evaluates from this row (ghost j, or (i+j) odd ? i<j : i>j) and core/skin is the reference's split at
// r <= r_n - skin (fp32, at build time).  The SET of every row, its counts and the core/skin split are the reference's;
// the order inside a segment is the traversal order of the fine lattice (deterministic: records of a fine cell are kept
// in ascending atom index).  The reference's order (stencil cells by (boundary flag, Morton), ascending atom index inside
// a cell, skin entries reversed) is a pure function of (cell of j, j), so meso_export_pair_table rebuilds it on demand
// (k_canonical_rows) for the bit-exact parity checks.  Table offsets are 64-bit (the reference overflows int past 13.4 M atoms).
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

constexpr uint32_t MJ = (1u << 27) - 1u;  // atom indices fit 27 bits in the hit queue (3 class bits above them)
constexpr uint32_t MQ = (1u << 26) - 1u;  // record positions fit 26 bits in a staged run (6 count bits above them); checked on the host

// cell coordinates of every atom (locals + ghosts) packed 10 bits per dimension, its fine cell, and the fine histogram
__global__ void __launch_bounds__(256) k_bin_count(SoA3c x, int *__restrict__ cellc, int *__restrict__ fine_of, int *__restrict__ fine_cnt,
                                                   const Counts *__restrict__ cnt, Box box)
{
    const int nlocal = cnt->nlocal, nall = nlocal + cnt->nghost;
    const int fm0 = 2 * box.m[0], fm1 = 2 * box.m[1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
        int b[3], h[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double xd = x.c[d][i];
            b[d] = clamp_rz(__fma_rn(xd - box.sublo[d], box.bininv[d], 1.0), 0, box.m[d]);   // UM/neighbor_meso.cu:410-412
            if (i >= nlocal) b[d] = (xd >= box.sublo[d]) ? (xd <= box.subhi[d] ? b[d] : box.m[d] - 1) : 0;   // :413-417
            // which half of its cell (the fine lattice only prunes: any consistent rule will do)
            h[d] = (xd - (box.sublo[d] + (double)(b[d] - 1) * box.binsize[d])) >= 0.5 * box.binsize[d] ? 1 : 0;
        }
        cellc[i] = b[0] | (b[1] << 10) | (b[2] << 20);
        const int f = (2 * b[0] + h[0]) + fm0 * ((2 * b[1] + h[1]) + fm1 * (2 * b[2] + h[2]));
        fine_of[i] = f;
        atomicAdd(fine_cnt + f, 1);
    }
}

// histogram of the reference cells from the stored cell coordinates (exports only)
__global__ void __launch_bounds__(256) k_cell_hist(const int *__restrict__ cellc, int *__restrict__ cell_cnt, const Counts *__restrict__ cnt,
                                                   int m0, int m1)
{
    const int nall = cnt->nlocal + cnt->nghost;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
        const int cc = cellc[i];
        atomicAdd(cell_cnt + (cc & 1023) + m0 * (((cc >> 10) & 1023) + m1 * (cc >> 20)), 1);
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
__global__ void __launch_bounds__(128) k_cell_order(const int *__restrict__ cell_start, int *__restrict__ cell_atoms, int ncell)
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
}

// cell-ordered records of the fine lattice: {x, y, z, bits(atom index)}
__global__ void __launch_bounds__(256) k_fine_fill(const float4 *__restrict__ coord4, const int *__restrict__ fine_of, int *__restrict__ fstart1,
                                                   float4 *__restrict__ fine_rec, const Counts *__restrict__ cnt)
{
    const int nall = cnt->nlocal + cnt->nghost;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
        float4 c = coord4[i];
        c.w = __int_as_float(i);
        fine_rec[atomicAdd(fstart1 + fine_of[i], 1)] = c;
    }
}

// records of a fine cell in ascending atom index: the traversal order, hence the row order, is the same in every run
__global__ void __launch_bounds__(128) k_fine_order(const int *__restrict__ fine_start, float4 *__restrict__ fine_rec, int nfine)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nfine) return;
    const int a = fine_start[c], n = fine_start[c + 1] - a;
    float4 *p = fine_rec + a;
    for (int k = 1; k < n; k++) {
        const float4 v = p[k];
        int q = k;
        while (q > 0 && __float_as_int(p[q - 1].w) > __float_as_int(v.w)) { p[q] = p[q - 1]; q--; }
        p[q] = v;
    }
}

// ------------------------------------------------------------------ build
// geometry of the fine lattice in the packed (sub-box centred, fp32) frame
struct FineGeom {
    float lat_lo[3];      // lower face of fine cell 0
    float w[3], inv_w[3]; // fine cell width and its inverse
    int m[3];             // reference cells per dimension
    float R2;             // (r_n + margin)^2: rows and chords are clipped with a margin far above the fp32 rounding of the geometry
    int clip;             // 2: every atom clips its own rows; 1: atoms of a cell share the union of their runs (lockstep); 0: no clipping
};

constexpr int NF_THREADS = 128;
constexpr int NROW = 36;          // 6 x 6 fine rows cover the 3 x 3 reference-cell rows of the stencil
constexpr int NQ = 64;            // per-lane hit queue depth; denser rows take the fall-back kernel

__device__ __forceinline__ void sts_u32(unsigned addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(unsigned addr) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void stg_u32_if(bool p, uint32_t *addr, uint32_t v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q st.global.b32 [%0], %1;\n\t}" ::"l"(addr), "r"(v), "r"((int)p) : "memory");
}

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

// Persistent grid (a few CTAs per SM); every warp owns a [NQ][32] hit queue in `scratch` (global memory: it lives in L2 and
// costs no shared memory, so the run lists are the only shared-memory tenant and ~40 warps per SM hide the record latency).
__global__ void __launch_bounds__(NF_THREADS) k_build_rows(const float4 *__restrict__ coord4, const int *__restrict__ cellc,
                                                           const int *__restrict__ fine_start, const float4 *__restrict__ fine_rec,
                                                           int *__restrict__ pair_count, int *__restrict__ owned_count,
                                                           int *__restrict__ core_split, int *__restrict__ pair_table,
                                                           Counts *__restrict__ cnt, int *__restrict__ fixup, uint32_t *__restrict__ scratch,
                                                           int n_col, float rc2_core, float rc2_tail, FineGeom g)
{
    __shared__ uint32_t s_runs[NF_THREADS / 32][NROW][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned rbase = (unsigned)__cvta_generic_to_shared(&s_runs[wid][0][lane]);
    uint32_t *const qbase = scratch + ((size_t)blockIdx.x * (NF_THREADS / 32) + wid) * (NQ * 32) + lane;   // entry t at qbase[t * 32]
    const int nlocal = cnt->nlocal;
    const int fm0 = 2 * g.m[0], fm1 = 2 * g.m[1], fm2 = 2 * g.m[2];
    int worst = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; (i & ~31) < nlocal; i += gridDim.x * blockDim.x) {
        const bool active = i < nlocal;
        const float4 ci = coord4[active ? i : 0];
        const int cc = cellc[active ? i : 0];
        const int cx = cc & 1023, cy = (cc >> 10) & 1023, cz = cc >> 20;
        // position in fine-cell units, clamped onto the lattice (an atom beyond it keeps valid lower bounds)
        const float ux = fminf(fmaxf((ci.x - g.lat_lo[0]) * g.inv_w[0], 0.f), (float)fm0);
        const float uy = fminf(fmaxf((ci.y - g.lat_lo[1]) * g.inv_w[1], 0.f), (float)fm1);
        const float uz = fminf(fmaxf((ci.z - g.lat_lo[2]) * g.inv_w[2], 0.f), (float)fm2);
        const int xmin = max(0, 2 * (cx - 1)), xmax = min(fm0 - 1, 2 * cx + 3);
        bool bad = false;
        const unsigned peers = __match_any_sync(full, active ? cc : -1);
        // ---- the 36 fine rows of the stencil, each clipped to the chord of the (margin-enlarged) cutoff sphere; the
        //      non-empty runs {first record, count} are staged compactly, one column per lane.  The 12 boundary loads of a
        //      plane are independent (clamped addresses, no branch around them) so they are in flight together.
        unsigned rp = rbase;
#pragma unroll 1
        for (int rz = 0; rz < 6; rz++) {
            const int fz = 2 * (cz - 1) + rz;
            const float tz = uz - (float)fz;
            const float dz = (tz < 0.f ? -tz : fmaxf(tz - 1.f, 0.f)) * g.w[2];
            const float remz = g.R2 - dz * dz;
            const bool okz = active && fz >= 0 && fz < fm2 && remz >= 0.f;
            if (g.clip == 2 && !__any_sync(full, okz)) continue;
            int q0[6], q1[6];
#pragma unroll
            for (int ry = 0; ry < 6; ry++) {
                const int fy = 2 * (cy - 1) + ry;
                const float ty = uy - (float)fy;
                const float dy = (ty < 0.f ? -ty : fmaxf(ty - 1.f, 0.f)) * g.w[1];
                const float rem = remz - dy * dy;
                const bool ok = okz && fy >= 0 && fy < fm1 && rem >= 0.f;
                const float ch = sqrt_approx(fmaxf(rem, 0.f)) * g.inv_w[0];      // MUFU.SQRT: the margin in R2 dwarfs its 1-ulp error
                int xlo = max(__float2int_rd(ux - ch), xmin), xhi = min(__float2int_rd(ux + ch), xmax);
                bool take = ok && xlo <= xhi;
                if (g.clip == 1) {
                    // lanes of the same reference cell walk the UNION of their clipped runs: identical runs advance in lockstep,
                    // so a record load of the warp touches one line per cell instead of one per lane
                    const unsigned grp = __ballot_sync(peers, take);
                    if (!take) { xlo = 0x7fffffff; xhi = -1; }
                    xlo = __reduce_min_sync(peers, xlo); xhi = __reduce_max_sync(peers, xhi);
                    take = active && grp != 0u;
                } else if (g.clip == 0) { xlo = xmin; xhi = xmax; take = active && fz >= 0 && fz < fm2 && fy >= 0 && fy < fm1; }
                const int *fs = fine_start + (take ? (size_t)fm0 * (size_t)(fy + fm1 * fz) : (size_t)0);
                if (!take) { xlo = 0; xhi = -1; }
                q0[ry] = fs[xlo];
                q1[ry] = fs[xhi + 1];
            }
#pragma unroll
            for (int ry = 0; ry < 6; ry++) {
                const int n = q1[ry] - q0[ry];
                if (n > 63) bad = true;
                else if (n > 0) { sts_u32(rp, (uint32_t)q0[ry] | ((uint32_t)n << 26)); rp += 128; }
            }
        }
        __syncwarp();
        // ---- walk the runs: every lane advances through its own runs (predicated, no divergence), two candidates per
        //      iteration; hits go to the lane's queue as j | skin << 27 (past NQ hits the last slot is overwritten and the
        //      row is redone by the fall-back kernel)
        unsigned rq = rbase;                                   // next run to fetch
        int q = 0, n = 0, qn = 0;
        while (true) {
            if (n <= 0 && rq < rp) { const uint32_t e = lds_u32(rq); rq += 128; q = (int)(e & MQ); n = (int)(e >> 26); }
            if (!__any_sync(full, n > 0)) break;
            const bool a0 = n > 0, a1 = n > 1;
            const float4 c0 = fine_rec[a0 ? q : 0], c1 = fine_rec[a1 ? q + 1 : 0];
            const float dx0 = ci.x - c0.x, dy0 = ci.y - c0.y, dz0 = ci.z - c0.z;
            const float dx1 = ci.x - c1.x, dy1 = ci.y - c1.y, dz1 = ci.z - c1.z;
            const float r0 = __fmaf_rn(dz0, dz0, __fmaf_rn(dy0, dy0, __fmul_rn(dx0, dx0)));   // UM/neigh_build_meso.cu:86-89
            const float r1 = __fmaf_rn(dz1, dz1, __fmaf_rn(dy1, dy1, __fmul_rn(dx1, dx1)));
            const bool h0 = a0 && r0 <= rc2_tail, h1 = a1 && r1 <= rc2_tail;
            stg_u32_if(h0, qbase + (min(qn, NQ - 1) << 5), (__float_as_uint(c0.w) & MJ) | (r0 <= rc2_core ? 0u : 1u << 27));
            qn += h0 ? 1 : 0;
            stg_u32_if(h1, qbase + (min(qn, NQ - 1) << 5), (__float_as_uint(c1.w) & MJ) | (r1 <= rc2_core ? 0u : 1u << 27));
            qn += h1 ? 1 : 0;
            q += 2; n -= 2;
        }
        bad = bad || qn > NQ;
        if (bad) qn = 0;
        __syncwarp();
        // ---- classify: drop the atom itself (it is always a hit), decide ownership, count the four classes
        const int qmax = __reduce_max_sync(full, qn);
        uint32_t cls_cnt = 0;                                  // 4 x 8-bit counters: owned core, owned skin, other core, other skin
        for (int t = 0; t < qmax; t++) {
            if (t < qn) {
                const uint32_t en = qbase[t << 5];
                const int j = (int)(en & MJ);
                uint32_t cls = 4;
                if (j != i) { cls = (owns(i, j, nlocal) ? 0u : 2u) | (en >> 27); cls_cnt += 1u << (8 * cls); }
                qbase[t << 5] = (uint32_t)j | (cls << 27);
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
                const uint32_t en = qbase[t << 5];
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

// Fall-back for the rows the kernel above marked (a fine row longer than 63 records, more than NQ neighbors): plain walk
// of all 6 x 6 fine rows of the 27 stencil cells, unclipped, two passes (count the classes, then write).  Also the whole build
// when MESO_NB_SLOW=1 (A/B checks of the kernel above).
__global__ void __launch_bounds__(128) k_build_rows_slow(const float4 *__restrict__ coord4, const int *__restrict__ cellc,
                                                         const int *__restrict__ fine_start, const float4 *__restrict__ fine_rec,
                                                         int *__restrict__ pair_count, int *__restrict__ owned_count,
                                                         int *__restrict__ core_split, int *__restrict__ pair_table,
                                                         Counts *__restrict__ cnt, const int *__restrict__ fixup, int all_rows, int n_col,
                                                         float rc2_core, float rc2_tail, int m0, int m1, int m2)
{
    if (!all_rows && *fixup == 0) return;
    const int nlocal = cnt->nlocal;
    const int fm0 = 2 * m0, fm1 = 2 * m1, fm2 = 2 * m2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
        if (!all_rows && pair_count[i] >= 0) continue;
        const float4 ci = coord4[i];
        const int cc = cellc[i];
        const int cx = cc & 1023, cy = (cc >> 10) & 1023, cz = cc >> 20;
        const int xlo = max(0, 2 * (cx - 1)), xhi = min(fm0 - 1, 2 * cx + 3);
        int num[4] = {0, 0, 0, 0}, pos[4] = {0, 0, 0, 0};
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) { pos[0] = 0; pos[1] = num[0]; pos[2] = num[0] + num[1]; pos[3] = num[0] + num[1] + num[2]; }
            for (int fz = max(0, 2 * (cz - 1)); fz <= min(fm2 - 1, 2 * cz + 3); fz++)
                for (int fy = max(0, 2 * (cy - 1)); fy <= min(fm1 - 1, 2 * cy + 3); fy++) {
                    const int *fs = fine_start + (size_t)fm0 * (size_t)(fy + fm1 * fz);
                    for (int p = fs[xlo]; p < fs[xhi + 1]; p++) {
                        const float4 c2 = fine_rec[p];
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
    const size_t nfine = (size_t)box.ncell * 8;
    if (nfine + 2 > (size_t)0x7fffffff) { ctx->err = "setup_bins: fine lattice too large"; return MESO_EINVAL; }
    if (!ctx->stencil.reserve((size_t)box.ncell * 32) || !ctx->slotrank.reserve((size_t)box.ncell * 32) ||
        !ctx->cell_start.reserve((size_t)box.ncell + 2) || !ctx->cell_cnt.reserve(nfine + 8) ||
        !ctx->fine_start.reserve(nfine + 2) || !ctx->scan_sums.reserve((nfine + SCAN_TILE - 1) / SCAN_TILE + 8)) {
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
    const int ncell = box.ncell, nfine = 8 * ncell;
    if (ctx->cap + 8 > (size_t)MQ) { ctx->err = "neighbor build: more than 2^26 atoms + ghosts on one GPU"; return MESO_EINVAL; }
    // persistent grid of the build kernel: the hit queues of its warps live in a global scratch buffer
    const int build_grid = ctx->sm_count * 9;                 // 56 registers x 128 threads: 9 CTAs per SM, one wave
    if (!ctx->fine_rec.reserve(ctx->cap + 8) || !ctx->fine_of.reserve(ctx->cap) || !ctx->owned_count.reserve(ctx->cap) ||
        !ctx->core_split.reserve(ctx->cap) || !ctx->nb_fixup.reserve(1) ||
        !ctx->nb_scratch.reserve((size_t)build_grid * (NF_THREADS / 32) * NQ * 32)) {
        ctx->err = "neighbor: out of device memory";
        return MESO_ECUDA;
    }
    cudaStream_t st = ctx->stream;
    int *fine_cnt = ctx->cell_cnt.p;
    MESO_CUDA(cudaMemsetAsync(fine_cnt, 0, sizeof(int) * (size_t)nfine, st));
    MESO_CUDA(cudaMemsetAsync(ctx->nb_fixup.p, 0, sizeof(int), st));
    SoA3c x; for (int d = 0; d < 3; d++) x.c[d] = ctx->x[d].p;
    k_bin_count<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(x, ctx->cell_of.p, ctx->fine_of.p, fine_cnt, ctx->d_counts, box);
    scan_into(ctx, fine_cnt, ctx->fine_start.p, nfine);
    k_fine_fill<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(ctx->coord4.p, ctx->fine_of.p, ctx->fine_start.p + 1, ctx->fine_rec.p, ctx->d_counts);
    k_fine_order<<<(nfine + 127) / 128, 128, 0, LS(st)>>>(ctx->fine_start.p, ctx->fine_rec.p, nfine);
    ctx->cells_valid = false;                               // the reference cell lists are rebuilt on demand (meso_export_cells)
    const float rc2_core = (float)pow(ctx->cutneighmax - ctx->skin, 2.0);   // UM/neigh_build_meso.cu:296-297
    const float rc2_tail = (float)pow(ctx->cutneighmax, 2.0);
    FineGeom g;
    double ext = 0;
    for (int d = 0; d < 3; d++) {
        g.lat_lo[d] = (float)((box.sublo[d] - box.binsize[d]) - box.centre[d]);
        g.w[d] = (float)(0.5 * box.binsize[d]);
        g.inv_w[d] = (float)(2.0 * box.bininv[d]);
        g.m[d] = box.m[d];
        ext = std::max(ext, box.subhi[d] - box.sublo[d] + 2.0 * box.binsize[d]);
    }
    const double margin = 1.0e-3 + 4.0e-6 * ext;            // fp32 rounding of the fine coordinates is ~1e-7 * extent
    g.R2 = (float)pow(ctx->cutneighmax + margin, 2.0);
    g.clip = ctx->nb_clip;
    const int slow_grid = std::max(1, std::min((int)((nlocal_bound(ctx) + 127) / 128) + 1, ctx->sm_count * 64));
    if (!ctx->nb_slow)
        k_build_rows<<<build_grid, NF_THREADS, 0, LS(st)>>>(ctx->coord4.p, ctx->cell_of.p, ctx->fine_start.p, ctx->fine_rec.p, ctx->pair_count.p,
                                                           ctx->owned_count.p, ctx->core_split.p, ctx->pair_table.p, ctx->d_counts,
                                                           ctx->nb_fixup.p, ctx->nb_scratch.p, ctx->n_col, rc2_core, rc2_tail, g);
    k_build_rows_slow<<<slow_grid, 128, 0, LS(st)>>>(ctx->coord4.p, ctx->cell_of.p, ctx->fine_start.p, ctx->fine_rec.p, ctx->pair_count.p,
                                                     ctx->owned_count.p, ctx->core_split.p, ctx->pair_table.p, ctx->d_counts, ctx->nb_fixup.p,
                                                     ctx->nb_slow ? 1 : 0, ctx->n_col, rc2_core, rc2_tail, box.m[0], box.m[1], box.m[2]);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// atoms in (cell, ascending index) order and the first position of every cell, as binning_meso leaves them
// (UM/neighbor_meso.cu:535-711): built from the cell coordinates of the last rebuild, when an export asks for them
int launch_cell_lists(meso_ctx *ctx)
{
    if (ctx->cells_valid) return MESO_OK;
    const Box &box = ctx->box;
    const int ncell = box.ncell;
    cudaStream_t st = ctx->stream;
    int *cell_cnt = ctx->cell_cnt.p;                         // the fine histogram is dead once the table is built
    MESO_CUDA(cudaMemsetAsync(cell_cnt, 0, sizeof(int) * (size_t)ncell, st));
    k_cell_hist<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(ctx->cell_of.p, cell_cnt, ctx->d_counts, box.m[0], box.m[1]);
    scan_into(ctx, cell_cnt, ctx->cell_start.p, ncell);
    k_bin_fill<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(ctx->cell_of.p, ctx->cell_start.p + 1, ctx->cell_atoms.p, ctx->d_counts, box.m[0], box.m[1]);
    k_cell_order<<<(ncell + 127) / 128, 128, 0, LS(st)>>>(ctx->cell_start.p, ctx->cell_atoms.p, ncell);
    MESO_CUDA(cudaGetLastError());
    ctx->cells_valid = true;
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
