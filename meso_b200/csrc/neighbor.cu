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
#include <cmath>
#include <cstdlib>

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

constexpr int NB_THREADS = 128;
// NB_BATCH: candidates tested per iteration (independent 16-byte loads in flight);
// NB_DEPTH: per-lane staging slots (core hits grow from the front, skin hits from the back)

// One thread owns one local atom.  In-range candidates are staged in a per-lane two-ended queue in shared
// memory (column layout [slot][lane]: bank == lane, no conflicts, no atomics), which costs 3 issue slots per
// candidate instead of the ~25 of computing a tile-transposed global address for a predicated store; the
// queue is written out once per atom, when both totals are known, so skin entries go straight to their
// final position (reverse encounter order after the core entries) and the reference's join pass disappears.
// Rows denser than the queue (> 60 hits, never at rho = 4) take the spill path: core entries are flushed
// forward, skin entries backward from slot n_col-1 as in the reference, and joined at the end.
// SKIP (MESO_NB_SKIP=1, off by default, NOT YET RUN ON HARDWARE): a stencil cell whose box lies farther than r_n from the
// atom is not walked at all -- half of the corner cells and a quarter of the edge cells, ~26 % of the candidate tests.  The
// criterion (fp32, 1e-3 margin on r_n^2) never drops a stored neighbor: tests/test_oracle_world.py::
// test_stencil_cell_skip_criterion_is_conservative checks it on the oracle's lattice.  SkipGeom: packed coordinate of the
// lower face of cell index 1 (= sublo - centre), cell size, cells per dimension, the per-cell stencil code rows.
struct SkipGeom { float lo[3], bs[3]; int m[3]; const unsigned char *stencil; float limit; };

template <int NB_DEPTH, int NB_BATCH, bool SKIP = false>
__global__ void __launch_bounds__(NB_THREADS) k_build_neighbors(const float4 *__restrict__ coord4, const int *__restrict__ cell_of,
                                                                const int2 *__restrict__ runs, const float4 *__restrict__ cell_xyzj,
                                                                int *__restrict__ pair_count, int *__restrict__ pair_table,
                                                                Counts *__restrict__ cnt, int n_col, float rc2_core, float rc2_tail,
                                                                const int *__restrict__ fixup, SkipGeom sg = SkipGeom())
{
    __shared__ int stage[NB_THREADS / 32][NB_DEPTH][32];
    // fix-up mode: only the rows the warp-per-cell kernel marked (pair_count == -1); nothing to do when no cell was marked
    if (fixup && *fixup == 0) return;
    int(*qq)[32] = stage[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int nlocal = cnt->nlocal;
    int worst = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
        if (fixup && pair_count[i] >= 0) continue;
        const float4 ci = coord4[i];
        const int mycell = cell_of[i];
        const int2 *my = runs + (size_t)mycell * 27;
        // SKIP: squared distance from the atom to the -1 / +1 neighbor layer along each axis (0 when the atom sits outside its
        // clamped cell on that side) and this cell's stencil code row
        float d2lo[3] = {0.f, 0.f, 0.f}, d2hi[3] = {0.f, 0.f, 0.f};
        const unsigned char *srow = nullptr;
        if constexpr (SKIP) {
            const int b[3] = {mycell % sg.m[0], (mycell / sg.m[0]) % sg.m[1], mycell / (sg.m[0] * sg.m[1])};
            const float p[3] = {ci.x, ci.y, ci.z};
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const float clo = sg.lo[d] + (float)(b[d] - 1) * sg.bs[d], chi = sg.lo[d] + (float)b[d] * sg.bs[d];
                const float a = fmaxf(p[d] - clo, 0.f), c = fmaxf(chi - p[d], 0.f);
                d2lo[d] = a * a; d2hi[d] = c * c;
            }
            srow = sg.stencil + (size_t)mycell * 32;
        }
        auto run_of = [&](int ss) -> int2 {
            int2 rr = my[ss];
            if constexpr (SKIP) {
                const int code = srow[ss], ox = code % 3, oy = (code / 3) % 3, oz = code / 9;
                const float d2 = (ox == 0 ? d2lo[0] : (ox == 2 ? d2hi[0] : 0.f)) + (oy == 0 ? d2lo[1] : (oy == 2 ? d2hi[1] : 0.f)) +
                                 (oz == 0 ? d2lo[2] : (oz == 2 ? d2hi[2] : 0.f));
                if (d2 > sg.limit) rr.y = 0;
            }
            return rr;
        };
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
        int2 run = run_of(0), nrun = run_of(1);
        int q = run.x, n = run.y;
        while (true) {
            while (n == 0 && s < 26) { s++; run = nrun; q = run.x; n = run.y; nrun = run_of(min(s + 1, 26)); }
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

// ------------------------------------------------------------------ build, one warp per cell (default)
// The thread-per-atom kernel above spends ~38 issue slots and ~10 L1 tag lookups per distance test: every lane
// walks its own cell's candidate runs, so each float4 load touches ~10 different 128-byte lines and every hit is a
// predicated shared-memory push.  Here a warp owns one cell and the roles are swapped, as in the reference
// (UM/neigh_build_meso.cu:58-117), but without its shared counters, sentinel, join and transpose passes:
//   fetch : lanes = 32 consecutive CANDIDATES of the cell's concatenated stencil list (the non-empty runs in stencil
//           order).  The run of position t comes from one warp OR-reduction per chunk: bit (start_k - t0) of M marks the
//           runs starting inside the chunk, so k(t0 + lane) = #runs before t0 + popc(M & lanemask_le) - 1.  One
//           coalesced 16-byte load per candidate; two chunks are held in registers as packed pairs.
//   test  : the cell's own atoms (<= 32 per pass) are broadcast one by one from shared memory; the distance test of two
//           candidates is 6 packed fp32x2 instructions (FADD2/FMUL2/FFMA2: same IEEE operations in the same order as
//           the scalar chain) and the 32 results of a chunk become two ballots (r <= r_n, r <= r_n - skin), stored by
//           lane 0: ~0.3 issue slots per test and no per-hit memory traffic.
//   count : lane m clears its own bit (j != i), prefix-sums its ballots over the chunks and publishes the row totals.
//   emit  : the (atom, chunk) ballots are spread over all 32 lanes (32/pow2(n) lanes per atom); every entry goes straight
//           to its final slot of the tile-transposed table: core entries ascending from slot 0, skin entries in reverse
//           encounter order behind them -- bit-identical rows, counts and layout.
// Cells whose candidate list exceeds the shared-memory window (local density >~ 1.5x the mean) are marked and rebuilt by
// the thread-per-atom kernel in fix-up mode, so capacity never changes results.
constexpr int NC_WARPS = 4;

__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// shared memory per warp: 1408 B + 16 B x t_cap
//   mxy[32]  float4 {x, x, y, y}    mz[32] float2 {z, z}    mi[32] int (atom index of member m)
//   tot[32]  int2 {n_core, n_tot}    cfirst[32], coffs[32] int (non-empty runs)
//   masks[m][pair] uint4 {within A, core A, within B, core B}      (8 B x t_cap)
//   pre[m][chunk]  uint {core entries before the chunk | skin entries before the chunk << 16}   (4 B x t_cap)
//   candj[t] int                                                                               (4 B x t_cap)
__global__ void __launch_bounds__(NC_WARPS * 32) k_build_neighbors_cell(const int *__restrict__ cell_start, const int2 *__restrict__ runs,
                                                                        const float4 *__restrict__ cell_xyzj,
                                                                        int *__restrict__ pair_count, int *__restrict__ pair_table,
                                                                        Counts *__restrict__ cnt, int *__restrict__ fixup, int ncell,
                                                                        int n_col, float rc2_core, float rc2_tail, int t_cap)
{
    extern __shared__ __align__(16) unsigned char nb_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int npair_cap = t_cap >> 6;                                 // chunk pairs per member row
    unsigned char *base = nb_smem + ((size_t)1408 + (size_t)16 * t_cap) * wid;
    float4 *mxy = reinterpret_cast<float4 *>(base);
    float2 *mz = reinterpret_cast<float2 *>(base + 512);
    int *mis = reinterpret_cast<int *>(base + 768);
    int2 *tot = reinterpret_cast<int2 *>(base + 896);
    int *cfirst = reinterpret_cast<int *>(base + 1152);
    int *coffs = reinterpret_cast<int *>(base + 1280);
    uint4 *masks = reinterpret_cast<uint4 *>(base + 1408);
    unsigned *pre = reinterpret_cast<unsigned *>(base + 1408 + (size_t)8 * t_cap);
    int *candj = reinterpret_cast<int *>(base + 1408 + (size_t)12 * t_cap);
    const unsigned full = 0xffffffffu, le = 0xffffffffu >> (31 - lane);
    const int nlocal = cnt->nlocal;
    int worst = 0;
    for (int c = blockIdx.x * NC_WARPS + wid; c < ncell; c += gridDim.x * NC_WARPS) {
        const int cs = cell_start[c], nc = cell_start[c + 1] - cs;
        if (nc == 0) continue;
        // atoms of a cell are in ascending index order and locals precede ghosts: a cell whose first atom is a ghost has no rows
        if (__float_as_int(cell_xyzj[cs].w) >= nlocal) continue;
        // ---- the cell's stencil runs, compacted to the non-empty ones (stencil order kept)
        const int2 run = lane < 27 ? runs[(size_t)c * 27 + lane] : make_int2(0, 0);
        int incl = run.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += y;
        }
        const int T = __shfl_sync(full, incl, 31);
        const unsigned nonempty = __ballot_sync(full, run.y > 0);
        const int nne = __popc(nonempty);
        __syncwarp();
        if (run.y > 0) {
            const int k = __popc(nonempty & (le >> 1));
            cfirst[k] = run.x - (incl - run.y);                       // first position minus list offset
            coffs[k] = incl - run.y;
        }
        __syncwarp();
        // lane k keeps the list offset of the k-th non-empty run (strictly increasing); lanes >= nne: never inside a chunk
        const int coff = lane < nne ? coffs[lane] : 0x3fffffff;
        const unsigned ownb = __ballot_sync(full, run.y > 0 && run.x == cs);
        const int own_off = __shfl_sync(full, incl - run.y, (__ffs(ownb) - 1) & 31);
        const bool too_long = T > t_cap;                              // falls back to the thread-per-atom kernel
        const int nch = too_long ? 0 : (T + 31) >> 5, ncp = (nch + 1) >> 1;
        __syncwarp();
        for (int m0 = 0; m0 < nc; m0 += 32) {
            const int ng = min(32, nc - m0);
            float4 me = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff));
            if (lane < ng) me = cell_xyzj[cs + m0 + lane];
            const int mi = __float_as_int(me.w);
            mxy[lane] = make_float4(me.x, me.x, me.y, me.y);
            mz[lane] = make_float2(me.z, me.z);
            mis[lane] = mi;
            __syncwarp();
            if (too_long) {
                if (lane < ng && mi < nlocal) { pair_count[mi] = -1; *fixup = 1; }
                continue;
            }
            // ---- fetch + test, two chunks (A, B) per pass
            int nb = 0;                                               // runs starting before the current chunk
            for (int cp = 0; cp < ncp; cp++) {
                unsigned long long nx, ny, nz;                        // packed negated candidate coordinates {A, B}
                {
                    float4 cd[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int t0 = (2 * cp + h) * 32, t = t0 + lane;
                        const unsigned d = (unsigned)(coff - t0);
                        const unsigned M = __reduce_or_sync(full, d < 32u ? 1u << d : 0u);
                        const int k = nb + __popc(M & le) - 1;
                        nb += __popc(M);
                        cd[h] = make_float4(3.0e18f, 3.0e18f, 3.0e18f, __int_as_float(-1));
                        if (t < T) cd[h] = cell_xyzj[cfirst[k] + t];
                        candj[t] = __float_as_int(cd[h].w);
                    }
                    nx = pack2(-cd[0].x, -cd[1].x); ny = pack2(-cd[0].y, -cd[1].y); nz = pack2(-cd[0].z, -cd[1].z);
                }
                uint4 *mrow = masks + cp;
#pragma unroll 4
                for (int m = 0; m < ng; m++) {
                    const float4 a = mxy[m];
                    const float2 zz = mz[m];
                    const unsigned long long dx = add2(pack2(a.x, a.y), nx), dy = add2(pack2(a.z, a.w), ny), dz = add2(pack2(zz.x, zz.y), nz);
                    const unsigned long long r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));       // UM/neigh_build_meso.cu:86-89
                    float rA, rB;
                    unpack2(r2, rA, rB);
                    uint4 bb;
                    bb.x = __ballot_sync(full, rA <= rc2_tail); bb.y = __ballot_sync(full, rA <= rc2_core);
                    bb.z = __ballot_sync(full, rB <= rc2_tail); bb.w = __ballot_sync(full, rB <= rc2_core);
                    if (lane == 0) mrow[m * npair_cap] = bb;
                }
            }
            __syncwarp();
            // ---- count: lane = member; clear my own bit, prefix over chunks, totals
            if (lane < ng) {
                const int ts = own_off + m0 + lane;                   // my own position in the candidate list (j != i)
                int ncore = 0, nskin = 0;
                for (int cp = 0; cp < ncp; cp++) {
                    uint4 bb = masks[lane * npair_cap + cp];
                    if ((ts >> 6) == cp) {
                        const unsigned bit = ~(1u << (ts & 31));
                        if (ts & 32) { bb.z &= bit; bb.w &= bit; } else { bb.x &= bit; bb.y &= bit; }
                        masks[lane * npair_cap + cp] = bb;
                    }
                    pre[lane * (2 * npair_cap) + 2 * cp] = (unsigned)ncore | ((unsigned)nskin << 16);
                    ncore += __popc(bb.y); nskin += __popc(bb.x & ~bb.y);
                    pre[lane * (2 * npair_cap) + 2 * cp + 1] = (unsigned)ncore | ((unsigned)nskin << 16);
                    ncore += __popc(bb.w); nskin += __popc(bb.z & ~bb.w);
                }
                int n_tot = ncore + nskin;
                if (mi < nlocal) {
                    if (n_tot > n_col) { atomicOr(&cnt->err, 2); n_tot = -min(ncore, n_col) - 1; }   // overflow: core entries only (as above)
                    pair_count[mi] = n_tot < 0 ? -n_tot - 1 : n_tot;
                    worst = max(worst, n_tot < 0 ? -n_tot - 1 : n_tot);
                }
                tot[lane] = make_int2(ncore, n_tot);
            }
            __syncwarp();
            // ---- emit: the ng x nch (member, chunk) ballots are dealt round-robin to the 32 lanes; a lane pops one entry
            //      per iteration and refills from its next item when its ballot is exhausted, so the warp runs
            //      max_lane(entries) iterations whatever the distribution of hits over chunks
            {
                const int nitems = ng * nch;
                int item = lane;
                int ch = (int)((float)lane * (1.0f / (float)ng) + 1.0e-4f), m = lane - ch * ng;   // lane = ch * ng + m (exact for lane < 32)
                const int dch = 32 / ng, dm = 32 - dch * ng;                                        // advance of (ch, m) per 32 items
                unsigned hits = 0, corem = 0;
                int kc = 0, ks = 0;
                int *row0 = pair_table;
                const int *cj = candj;
                while (true) {
                    if (hits == 0) {
                        if (item >= nitems) break;
                        const int am = mis[m];
                        const int2 tt = tot[m];
                        const uint2 mk = reinterpret_cast<const uint2 *>(masks + m * npair_cap + (ch >> 1))[ch & 1];
                        const unsigned pw = pre[m * (2 * npair_cap) + ch];
                        const bool overflow = tt.y < 0;
                        const int n_tot = overflow ? -tt.y - 1 : tt.y;
                        kc = pw & 0xffffu; ks = n_tot - 1 - (int)(pw >> 16);
                        corem = mk.y;
                        hits = am < nlocal ? (overflow ? mk.y & ((kc < n_col) ? 0xffffffffu : 0u) : mk.x) : 0u;
                        if (overflow && hits) {                                   // keep the first n_col core entries only
                            while (kc + __popc(hits) > n_col) hits &= ~(0x80000000u >> __clz(hits));
                        }
                        row0 = pair_table + (size_t)(am & ~31) * (size_t)n_col + (am & 31);   // slot(i,k) = row0[(k&31)*n_col + (k>>5)*32]
                        cj = candj + ch * 32;
                        item += 32; ch += dch; m += dm;
                        if (m >= ng) { m -= ng; ch++; }
                        continue;
                    }
                    const int b = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const bool is_core = (corem >> b) & 1u;
                    const int k = is_core ? kc : ks;
                    kc += is_core ? 1 : 0; ks -= is_core ? 0 : 1;
                    row0[(k & 31) * n_col + (k & ~31)] = cj[b];
                }
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) worst = max(worst, __shfl_xor_sync(full, worst, o));
    if (lane == 0 && worst > 0) atomicMax(&cnt->max_pair, worst);
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
    k_stencil_codes<<<(box.ncell + 127) / 128, 128, 0, LS(ctx->stream)>>>(ctx->stencil.p, box);
    MESO_CUDA(cudaGetLastError());
    ctx->bins_ready = true;
    return MESO_OK;
}

int launch_neighbor_build(meso_ctx *ctx)
{
    const Box &box = ctx->box;
    SoA3c x; for (int d = 0; d < 3; d++) x.c[d] = ctx->x[d].p;
    k_cell_id<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(x, ctx->cell_key.p, ctx->cell_atoms.p, ctx->cell_of.p, ctx->d_counts, box);
    int bits = 1;
    while ((1 << bits) < box.ncell) bits++;                 // ceil(log2(ncell)), UM/neighbor_meso.cu:541
    int rc = sort_pairs_u64(ctx, ctx->cell_key, ctx->cell_atoms, &ctx->d_counts->nall, ctx->cap, bits);
    if (rc) return rc;
    if (!ctx->cell_xyzj.reserve(ctx->cap + 8) || !ctx->cell_runs.reserve((size_t)box.ncell * 27)) { ctx->err = "neighbor: out of device memory"; return MESO_ECUDA; }
    k_cell_bounds<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(ctx->cell_key.p, ctx->cell_atoms.p, ctx->coord4.p, ctx->cell_start.p, ctx->cell_xyzj.p,
                                                          ctx->d_counts, box.ncell);
    k_cell_runs<<<(box.ncell * 27 + 255) / 256, 256, 0, LS(ctx->stream)>>>(ctx->stencil.p, ctx->cell_start.p, ctx->cell_runs.p, box);
    float rc2_core = (float)pow(ctx->cutneighmax - ctx->skin, 2.0);   // UM/neigh_build_meso.cu:296-297
    float rc2_tail = (float)pow(ctx->cutneighmax, 2.0);
    const bool per_atom = ctx->nb_per_atom;
    int grid_atoms = std::max(1, std::min((int)((nlocal_bound(ctx) + 127) / 128) + 1, ctx->sm_count * 4096));
    if (!per_atom) {
        // candidate window per cell: 1.25x the expected stencil population (+64), whole chunk pairs
        double cellvol = box.binsize[0] * box.binsize[1] * box.binsize[2], vol = 1.0;
        for (int d = 0; d < 3; d++) vol *= box.subhi[d] - box.sublo[d];
        const double t_exp = 27.0 * cellvol * std::max(ctx->nlocal_host / vol, 1.0);
        int t_cap = ((int)(1.25 * t_exp) + 64 + 63) / 64 * 64;
        t_cap = std::max(256, std::min(t_cap, 2048));
        if (const char *e = getenv("MESO_NB_TCAP")) t_cap = std::max(64, atoi(e) / 64 * 64);      // tests: force the fix-up path
        const size_t sh = (size_t)NC_WARPS * (1408 + (size_t)16 * t_cap);
        static size_t sh_set = 0;
        if (sh > sh_set) {
            MESO_CUDA(cudaFuncSetAttribute(k_build_neighbors_cell, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
            sh_set = sh;
        }
        if (!ctx->nb_fixup.reserve(1)) { ctx->err = "neighbor: out of device memory"; return MESO_ECUDA; }
        MESO_CUDA(cudaMemsetAsync(ctx->nb_fixup.p, 0, sizeof(int), ctx->stream));
        const int grid = std::max(1, std::min((box.ncell + NC_WARPS - 1) / NC_WARPS, ctx->sm_count * 64));
        k_build_neighbors_cell<<<grid, NC_WARPS * 32, sh, LS(ctx->stream)>>>(ctx->cell_start.p, ctx->cell_runs.p, ctx->cell_xyzj.p, ctx->pair_count.p,
                                                                        ctx->pair_table.p, ctx->d_counts, ctx->nb_fixup.p, box.ncell, ctx->n_col,
                                                                        rc2_core, rc2_tail, t_cap);
        k_build_neighbors<64, 4><<<grid_atoms, 128, 0, LS(ctx->stream)>>>(ctx->coord4.p, ctx->cell_of.p, ctx->cell_runs.p, ctx->cell_xyzj.p, ctx->pair_count.p,
                                                                   ctx->pair_table.p, ctx->d_counts, ctx->n_col, rc2_core, rc2_tail, ctx->nb_fixup.p);
    } else {
#define MESO_NB_ARGS ctx->coord4.p, ctx->cell_of.p, ctx->cell_runs.p, ctx->cell_xyzj.p, ctx->pair_count.p, ctx->pair_table.p, ctx->d_counts, ctx->n_col, rc2_core, rc2_tail, nullptr
        if (ctx->nb_skip) {
            SkipGeom sg;
            for (int d = 0; d < 3; d++) { sg.lo[d] = (float)(box.sublo[d] - box.centre[d]); sg.bs[d] = (float)box.binsize[d]; sg.m[d] = box.m[d]; }
            sg.stencil = ctx->stencil.p;
            sg.limit = rc2_tail * 1.002f;                           // (1.001 r_n)^2
            k_build_neighbors<48, 4, true><<<grid_atoms, 128, 0, LS(ctx->stream)>>>(MESO_NB_ARGS, sg);
        } else
        k_build_neighbors<48, 4><<<grid_atoms, 128, 0, LS(ctx->stream)>>>(MESO_NB_ARGS);   // 48 slots: 9 CTAs/SM; 40 spills too often, 56+ loses occupancy (measured)
#undef MESO_NB_ARGS
    }
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
