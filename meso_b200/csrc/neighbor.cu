// neighbor.cu -- cell binning and the atomics-free build of the ordered, tile-transposed
// full neighbor list.
//
// Reference path:
//   MesoNeighbor::setup_bins          UM/neighbor_meso.cu:858-931   (cell lattice aligned to the sub-domain)
//   gpu_stencil_full_bin_3d           UM/neighbor_meso.cu:772-825   (<=27 cells, by (boundary flag, Morton))
//   binning_meso                      UM/neighbor_meso.cu:535-711   (cell id, radix sort, boundaries, expanded stencils)
//   gpu_build_neighbor_list<32,5>     UM/neigh_build_meso.cu:20-119 (warp per cell, 2 ballots per test, row-major rows)
//   gpu_join_neigh_list / gpu_transpose_neigh_list  UM/neigh_build_meso.cu:166-240 (two more passes over the table)
// Blackwell version: no materialised per-cell stencil rows (5 KB/cell in the reference) -- a
// 32-byte per-cell code row names the <=27 neighbor cells in order and each thread walks
// their atom runs directly; one thread owns one local atom, so the running core/skin
// counters live in registers (no ballots, no shared counters, no atomics), entries are
// written straight into the tile-transposed table (core forward from slot 0, skin backward
// from slot n_col-1) and the owning thread joins its own row at the end: one pass over the
// table instead of three, with 64-bit table offsets (the reference overflows int past 13.4 M atoms).
#include "internal.h"
#include "device_math.cuh"
#include <algorithm>

namespace meso {

int sort_pairs_u64(meso_ctx *ctx, DevBuf<uint64_t> &key, DevBuf<int> &val, const int *d_n, size_t cap, int bits);

// ------------------------------------------------------------------ per-cell stencil code rows
// byte s (< 27): offset code (i+1) + 3*(j+1) + 9*(k+1) of the s-th neighbor cell; byte 31: count.
__global__ void k_stencil_codes(unsigned char *__restrict__ stencil, Box box)
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
    unsigned char *row = stencil + (size_t)cell * 32;
    for (int s = 0; s < 27; s++) row[s] = s < n ? code[s] : 0;
    row[31] = (unsigned char)n;
}

// ------------------------------------------------------------------ cell id of every atom (locals + ghosts)
struct SoA3c { const double *c[3]; };

__global__ void __launch_bounds__(256) k_cell_id(SoA3c x, uint64_t *__restrict__ cell_key, int *__restrict__ cell_val,
                                                 int *__restrict__ cell_of, const Counts *__restrict__ cnt, Box box)
{
    const int nlocal = cnt->nlocal, nall = nlocal + cnt->nghost;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
        int b[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double xd = x.c[d][i];
            b[d] = clamp_rz(__fma_rn(xd - box.sublo[d], box.bininv[d], 1.0), 0, box.m[d]);   // UM/neighbor_meso.cu:410-412
            if (i >= nlocal) b[d] = (xd >= box.sublo[d]) ? (xd <= box.subhi[d] ? b[d] : box.m[d] - 1) : 0;   // :413-417
        }
        int c = b[0] + box.m[0] * (b[1] + b[2] * box.m[1]);
        cell_key[i] = (uint64_t)c;
        cell_val[i] = i;
        cell_of[i] = c;
    }
}

// cell_start[c] = first sorted position whose cell id >= c  (gpu_find_bin_boundary, UM/neighbor_meso.cu:423-460);
// the same pass writes a cell-ordered copy of the packed coordinates, {x, y, z, bits(atom index)}, so that the
// build kernel streams candidates with ONE contiguous 16-byte load each instead of index load + float4 gather.
__global__ void __launch_bounds__(256) k_cell_bounds(const uint64_t *__restrict__ cell_sorted, const int *__restrict__ cell_atoms,
                                                     const float4 *__restrict__ coord4, int *__restrict__ cell_start,
                                                     float4 *__restrict__ cell_xyzj, const Counts *__restrict__ cnt, int ncell)
{
    const int nall = cnt->nlocal + cnt->nghost;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p <= nall; p += gridDim.x * blockDim.x) {
        int cur = p < nall ? (int)cell_sorted[p] : ncell;
        int prev = p > 0 ? (int)cell_sorted[p - 1] : -1;
        for (int c = prev + 1; c <= cur; c++) cell_start[c] = p;
        if (p < nall) {
            const int j = cell_atoms[p];
            float4 v = coord4[j];
            v.w = __int_as_float(j);
            cell_xyzj[p] = v;
        }
    }
}

// runs[c][s] = {first position, count} of the s-th stencil cell of cell c, in stencil order
__global__ void __launch_bounds__(256) k_cell_runs(const unsigned char *__restrict__ stencil, const int *__restrict__ cell_start,
                                                   int2 *__restrict__ runs, Box box)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = t / 27, s = t - c * 27;
    if (c >= box.ncell) return;
    const unsigned char *row = stencil + (size_t)c * 32;
    int2 r = make_int2(0, 0);
    if (s < row[31]) {
        const int code = row[s];
        const int nc = c + (code % 3 - 1) + box.m[0] * ((code / 3) % 3 - 1 + box.m[1] * (code / 9 - 1));
        const int a = cell_start[nc];
        r = make_int2(a, cell_start[nc + 1] - a);
    }
    runs[t] = r;
}

// ------------------------------------------------------------------ build
__device__ __forceinline__ size_t slot(int i, int k, int n_col)
{
    return (size_t)((i & ~31) + (k & 31)) * (size_t)n_col + (size_t)((k >> 5) * 32 + (i & 31));
}

constexpr int NB_BATCH = 4;     // candidates tested per iteration (independent 16-byte loads in flight)
constexpr int NB_THREADS = 128;
constexpr int NB_DEPTH = 64;    // per-lane staging slots: core hits grow from the front, skin hits from the back

// One thread owns one local atom.  In-range candidates are staged in a per-lane two-ended queue in shared
// memory (column layout [slot][lane]: bank == lane, no conflicts, no atomics), which costs 3 issue slots per
// candidate instead of the ~25 of computing a tile-transposed global address for a predicated store; the
// queue is written out once per atom, when both totals are known, so skin entries go straight to their
// final position (reverse encounter order after the core entries) and the reference's join pass disappears.
// Rows denser than the queue (> 60 hits, never at rho = 4) take the spill path: core entries are flushed
// forward, skin entries backward from slot n_col-1 as in the reference, and joined at the end.
__global__ void __launch_bounds__(NB_THREADS) k_build_neighbors(const float4 *__restrict__ coord4, const int *__restrict__ cell_of,
                                                                const int2 *__restrict__ runs, const float4 *__restrict__ cell_xyzj,
                                                                int *__restrict__ pair_count, int *__restrict__ pair_table,
                                                                Counts *__restrict__ cnt, int n_col, float rc2_core, float rc2_tail)
{
    __shared__ int stage[NB_THREADS / 32][NB_DEPTH][32];
    int(*qq)[32] = stage[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int nlocal = cnt->nlocal;
    int worst = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
        const float4 ci = coord4[i];
        const int2 *my = runs + (size_t)cell_of[i] * 27;
        int *row0 = pair_table + (size_t)(i & ~31) * (size_t)n_col + (i & 31);   // slot(i,k) = row0[(k&31)*n_col + (k>>5)*32]
        int ncq = 0, nsq = 0;          // staged core / skin entries
        int ncw = 0, nsw = 0;          // entries already written to the table by the spill path
        bool overflow = false;
        auto put = [&](int k, int j) {
            if (k < n_col) row0[(k & 31) * n_col + (k >> 5) * 32] = j; else overflow = true;
        };
        // flattened walk over the 27 runs: every lane advances through its own concatenated candidate list, so lanes of
        // different cells do not wait for each other's cell sizes
        int s = 0;
        int2 run = my[0], nrun = my[1];
        int q = run.x, n = run.y;
        while (true) {
            while (n == 0 && s < 26) { s++; run = nrun; q = run.x; n = run.y; nrun = my[min(s + 1, 26)]; }
            if (n == 0) break;
            if (ncq + nsq > NB_DEPTH - NB_BATCH) {                               // spill path (dense rows only)
                for (int t = 0; t < ncq; t++) put(ncw + t, qq[t][lane]);
                for (int t = 0; t < nsq; t++) { if (ncw + ncq + nsw + t < n_col) put(n_col - 1 - (nsw + t), qq[NB_DEPTH - 1 - t][lane]); else overflow = true; }
                ncw += ncq; nsw += nsq; ncq = 0; nsq = 0;
            }
            const int take = min(n, NB_BATCH);
            float4 v[NB_BATCH];
#pragma unroll
            for (int u = 0; u < NB_BATCH; u++) v[u] = cell_xyzj[q + u];       // unconditional (array is padded): 4 loads in flight
#pragma unroll
            for (int u = 0; u < NB_BATCH; u++) {
                const int j = __float_as_int(v[u].w);
                const float dx = ci.x - v[u].x, dy = ci.y - v[u].y, dz = ci.z - v[u].z;
                const float dr2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));   // UM/neigh_build_meso.cu:86-89
                const bool ok = u < take && j != i;
                const bool is_core = ok && dr2 <= rc2_core;
                const bool is_skin = ok && !is_core && dr2 <= rc2_tail;
                if (is_core | is_skin) qq[is_core ? ncq : NB_DEPTH - 1 - nsq][lane] = j;
                ncq += is_core ? 1 : 0;
                nsq += is_skin ? 1 : 0;
            }
            q += take; n -= take;
        }
        int n_core = ncw + ncq, n_skin = nsw + nsq;
        if (nsw == 0) {
            // common path: everything is staged; skin entry s (encounter order) lands at n_core + n_skin - 1 - s
            if (n_core + n_skin > n_col) overflow = true;
            for (int t = 0; t < ncq; t++) put(ncw + t, qq[t][lane]);
            for (int t = 0; t < nsq; t++) put(n_core + n_skin - 1 - t, qq[NB_DEPTH - 1 - t][lane]);
        } else {
            for (int t = 0; t < ncq; t++) put(ncw + t, qq[t][lane]);
            for (int t = 0; t < nsq; t++) { if (n_core + nsw + t < n_col) put(n_col - 1 - (nsw + t), qq[NB_DEPTH - 1 - t][lane]); else overflow = true; }
            // join (UM/neigh_build_meso.cu:166-200): ascending t is safe even when the ranges overlap (dst(t) < src(t))
            if (!overflow)
                for (int t = 0; t < n_skin; t++) pair_table[slot(i, n_core + t, n_col)] = pair_table[slot(i, n_col - n_skin + t, n_col)];
        }
        if (overflow) { n_core = min(n_core, n_col); n_skin = 0; atomicOr(&cnt->err, 2); }
        pair_count[i] = n_core + n_skin;
        worst = max(worst, n_core + n_skin);
    }
    // diagnostics only
#pragma unroll
    for (int o = 16; o; o >>= 1) worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if ((threadIdx.x & 31) == 0 && worst > 0) atomicMax(&cnt->max_pair, worst);
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
    }
    box.ncell = box.m[0] * box.m[1] * box.m[2];
    if (!ctx->stencil.reserve((size_t)box.ncell * 32) || !ctx->cell_start.reserve((size_t)box.ncell + 2)) {
        ctx->err = "setup_bins: out of device memory";
        return MESO_ECUDA;
    }
    k_stencil_codes<<<(box.ncell + 127) / 128, 128, 0, ctx->stream>>>(ctx->stencil.p, box);
    MESO_CUDA(cudaGetLastError());
    ctx->bins_ready = true;
    return MESO_OK;
}

int launch_neighbor_build(meso_ctx *ctx)
{
    const Box &box = ctx->box;
    SoA3c x; for (int d = 0; d < 3; d++) x.c[d] = ctx->x[d].p;
    k_cell_id<<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(x, ctx->cell_key.p, ctx->cell_atoms.p, ctx->cell_of.p, ctx->d_counts, box);
    int bits = 1;
    while ((1 << bits) < box.ncell) bits++;                 // ceil(log2(ncell)), UM/neighbor_meso.cu:541
    int rc = sort_pairs_u64(ctx, ctx->cell_key, ctx->cell_atoms, &ctx->d_counts->nall, ctx->cap, bits);
    if (rc) return rc;
    if (!ctx->cell_xyzj.reserve(ctx->cap + 8) || !ctx->cell_runs.reserve((size_t)box.ncell * 27)) { ctx->err = "neighbor: out of device memory"; return MESO_ECUDA; }
    k_cell_bounds<<<grid_for(ctx, 8), 256, 0, ctx->stream>>>(ctx->cell_key.p, ctx->cell_atoms.p, ctx->coord4.p, ctx->cell_start.p, ctx->cell_xyzj.p,
                                                          ctx->d_counts, box.ncell);
    k_cell_runs<<<(box.ncell * 27 + 255) / 256, 256, 0, ctx->stream>>>(ctx->stencil.p, ctx->cell_start.p, ctx->cell_runs.p, box);
    float rc2_core = (float)pow(ctx->cutneighmax - ctx->skin, 2.0);   // UM/neigh_build_meso.cu:296-297
    float rc2_tail = (float)pow(ctx->cutneighmax, 2.0);
    {
        int grid = std::max(1, std::min((int)((nlocal_bound(ctx) + 127) / 128) + 1, ctx->sm_count * 4096));
        k_build_neighbors<<<grid, 128, 0, ctx->stream>>>(ctx->coord4.p, ctx->cell_of.p, ctx->cell_runs.p, ctx->cell_xyzj.p, ctx->pair_count.p,
                                                      ctx->pair_table.p, ctx->d_counts, ctx->n_col, rc2_core, rc2_tail);
    }
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
