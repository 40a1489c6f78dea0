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
//     writes the cell-ordered records {x, y, z, atom index} the build reads;
//   * the build stages cell tiles in shared memory with TMA bulk copies and walks them with one lane per atom (see "build"
//     below); in-range neighbors are staged in a per-lane queue in shared memory (column layout: bank == lane, no atomics,
//     no ballots) and written out once, when the row's totals are known.
// Row layout: [owned][other], where "owned" marks the entries whose pair the pair-once force kernel evaluates from this row
// (ghost j, or (i+j) odd ? i<j : i>j).  The SET of every row and its count are the reference's; the order inside a
// part is the walk order (x-rows of the stencil, ascending atom index inside a cell).  The reference's order (core entries by
// stencil position, skin entries reversed; stencil cells by (boundary flag, Morton)) is a pure function of (cell of j, j) and
// of the build-time distance, so meso_export_pair_table rebuilds it on demand (k_canonical_rows) for the bit-exact parity
// checks.  Table offsets are 64-bit (the reference overflows int past 13.4 M atoms).
//
// Measured and dropped in this round (profiles/): a half-cell lattice with per-atom chord clipping (72 candidates per atom
// instead of 246, but 16 cache lines per load and 36 row set-ups: 576 us); one lane per atom walking the 9 runs straight from
// global memory through L1 (487 us, 352 M warp instructions, 43 per candidate: run switching and the four-way split of the row).
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

// the cell-ordered copies of the packed coordinates: records {x, y, z, bits(atom index)} for the fall-back build and the
// exports, the same as four arrays x | y | z | index for the tile build, and every atom's position in that order.
// One thread per record: coalesced writes, one 16-byte gather.
__global__ void __launch_bounds__(256) k_cell_records(const int *__restrict__ cell_atoms, const float4 *__restrict__ coord4,
                                                      float4 *__restrict__ cell_xyzj, int *__restrict__ pos_of, float *__restrict__ cell_soa,
                                                      size_t soa_stride, const Counts *__restrict__ cnt)
{
    const int nall = cnt->nlocal + cnt->nghost;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nall; p += gridDim.x * blockDim.x) {
        const int j = cell_atoms[p];
        float4 v = coord4[j];
        v.w = __int_as_float(j);
        cell_xyzj[p] = v;
        cell_soa[p] = v.x; cell_soa[soa_stride + p] = v.y; cell_soa[2 * soa_stride + p] = v.z; cell_soa[3 * soa_stride + p] = v.w;
        pos_of[j] = p;
    }
}

// ------------------------------------------------------------------ build
// Work decomposition.  Locals are sorted by (border bit, Morton(cell), sub-cell), so the atoms of an aligned block of
// 2^lbx x 2^lby x 2^lbz cells are ONE run of consecutive indices (two where the block holds bulk and border atoms).  Such a
// run -- a "segment", found on the fly: a CTA owns the segments that START in its 256 atoms -- is built from one
// shared-memory tile: the block's cells plus one layer around them, i.e. up to 36 x-rows of up to 6 cells, each row one
// contiguous piece of the cell-ordered records, brought in by one cp.async.bulk per row (TMA, mbarrier completion).
// One lane owns one atom (lanes = consecutive atoms: the tile-transposed table is written with coalesced stores) and walks
// its 9 runs of 3 cells in the tile: 4 LDS.128 in flight, 6 FP32 ops and a predicated push per candidate.
// Row layout: [owned][other], each part in walk order; see owns().
constexpr int NB_WARPS = 8, NB_THREADS = NB_WARPS * 32;
constexpr int NB_BATCH = 4;                 // candidates per iteration (independent 16-byte shared-memory loads in flight)
constexpr int TILE_ROWS = 36, TILE_NX = 7;  // (4+2) x (4+2) x-rows, (4+2)+1 cell offsets per row
constexpr int TILE_CAP_MAX = 2560;          // records (40 KB): with the static shared memory in front the tile ends below 64 KB

__device__ __forceinline__ void sts_u32(unsigned addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(unsigned addr) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void sts_u16(unsigned addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16(unsigned addr) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ float4 lds_f4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_init(unsigned bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n"
                 ::"r"(bar), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared::cta, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}

// the entries of row i that its own lane evaluates in the pair-once force kernel (pair.cu): ghost partners always, local
// partners by the balanced rule (i+j) odd ? i<j : i>j (the mirror entry in row j is then "other"): every row owns half of
// ITS entries (17.9 +- 2.6 at rho = 4; a geometric half-space rule gives +- 5.9 and a 20 % longer slowest lane per warp)
__device__ __forceinline__ bool owns(int i, int j, int nlocal)
{
    const unsigned key = (unsigned)(j - i) * 0x80000001u;         // sign bit: d odd ? d > 0 : d < 0
    return (int)(key | (unsigned)(nlocal - 1 - j)) < 0;
}

__device__ __forceinline__ int block_of(int cc, int lbx, int lby, int lbz)
{
    return ((cc & 1023) >> lbx) | ((((cc >> 10) & 1023) >> lby) << 10) | (((cc >> 20) >> lbz) << 20);
}

__device__ __forceinline__ size_t slot(int i, int k, int n_col)
{
    return (size_t)((i & ~31) + (k & 31)) * (size_t)n_col + (size_t)((k >> 5) * 32 + (i & 31));
}

struct TileGeom { int lbx, lby, lbz, tile_cap, nq; };

// packed fp32 pairs (FADD2 / FMUL2 / FFMA2: two candidates per instruction, each half rounded like the scalar operation)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ void lds_2x64(unsigned addr, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}

// The tile is four arrays (x, y, z, atom index) of TILE_CAP_MAX words at fixed strides, so one address register serves all
// three coordinate loads and four candidates cost three LDS.128 (12 shared-memory wavefronts, not 16) and twelve packed
// floating-point instructions.  Rows are copied in whole 16-byte groups: a run starts at the group that holds its first record
// and masks what lies outside [first, last).  (Padding every cell to whole groups removes the masks -- 30 instead of 48
// instructions per four candidates -- but adds 17 % candidates and doubles the cell-ordering pass: 377 + 68 us against
// 361 + 39 us in a first measurement whose rows were still incomplete (parity red), so the variant was dropped there:
// profiles/r02_s3_build_tiles_padded_ncu_summary.txt.  Groups of four records [x0..x3][y0..y3][z0..z3][j0..j3], which
// need one bulk copy per row instead of four (the warps spend 22 % of their time waiting for the 144 small copies of a
// tile), put the quads of different cells on two bank groups instead of eight: 437 + 68 us, parity green,
// profiles/r02_s3_build_tiles_grouped_ncu_summary.txt.  Warps per CTA (MESO_NB_WARPS experiment, binning + build per
// rebuild): 5: 0.522 ms, 6: 0.505, 7: 0.508, 8: 0.506, 9: 0.515, 10: 0.540, 12: 0.547.)
constexpr unsigned TILE_STRIDE = TILE_CAP_MAX * 4;          // bytes between the arrays of the tile

__global__ void __launch_bounds__(NB_THREADS, 3) k_build_tiles(const float4 *__restrict__ coord4, const int *__restrict__ cellc,
                                                              const int *__restrict__ cell_start, const float *__restrict__ cell_soa,
                                                              size_t soa_stride, int *__restrict__ pair_count,
                                                              int *__restrict__ owned_count, int *__restrict__ pair_table,
                                                              Counts *__restrict__ cnt, int *__restrict__ fixup, int n_col, float rc2, int m0,
                                                              int m1, int m2, TileGeom g)
{
    extern __shared__ __align__(128) unsigned char smem[];      // [x | y | z | j : TILE_CAP_MAX words each][per-warp hit queues: nq x 32 x u16]
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_cell[TILE_ROWS * TILE_NX];                 // tile index of the first record of every tile cell (+ row end)
    __shared__ int s_rowsrc[TILE_ROWS], s_rowlen[TILE_ROWS], s_rowoff[TILE_ROWS + 1];   // per row: first global group, words copied, tile offset
    __shared__ int s_heads[NB_THREADS], s_nh, s_red[NB_WARPS];
    const unsigned full = 0xffffffffu;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const unsigned tile_base = (unsigned)__cvta_generic_to_shared(smem);
    // a hit is queued as the 16-bit byte offset of its record inside an array of the tile
    const unsigned qbase = tile_base + 4u * TILE_STRIDE + (unsigned)(wid * g.nq * 64 + lane * 2);
    const unsigned qlim = qbase + (unsigned)(g.nq - 1) * 64u;   // the last slot absorbs what does not fit (row goes to the fall-back)
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
    const int nlocal = cnt->nlocal;
    // ---- segments that start in this CTA's atoms
    if (t == 0) { s_nh = 0; mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    {
        const int i = blockIdx.x * NB_THREADS + t;
        if (i < nlocal && (i == 0 || block_of(cellc[i], g.lbx, g.lby, g.lbz) != block_of(cellc[i - 1], g.lbx, g.lby, g.lbz)))
            s_heads[atomicAdd(&s_nh, 1)] = i;
    }
    __syncthreads();
    const int nh = s_nh;
    unsigned phase = 0;
    int worst = 0;
    for (int hh = 0; hh < nh; hh++) {
        const int a = s_heads[hh];
        const int b0 = block_of(cellc[a], g.lbx, g.lby, g.lbz);
        // ---- end of the segment: first atom behind a that lies in another block
        int e = nlocal;
        for (int base = a + 1; base < nlocal; base += NB_THREADS) {
            const int i = base + t;
            int c = (i < nlocal && block_of(cellc[i], g.lbx, g.lby, g.lbz) != b0) ? i : 0x7fffffff;
            c = __reduce_min_sync(full, c);
            if (lane == 0) s_red[wid] = c;
            __syncthreads();
            c = s_red[0];
#pragma unroll
            for (int w = 1; w < NB_WARPS; w++) c = min(c, s_red[w]);
            __syncthreads();
            if (c != 0x7fffffff) { e = c; break; }
        }
        // ---- tile geometry: the block's cells and one layer around them, clipped to the lattice
        const int X0 = (b0 & 1023) << g.lbx, Y0 = ((b0 >> 10) & 1023) << g.lby, Z0 = (b0 >> 20) << g.lbz;
        const int xlo = max(X0 - 1, 0), xhi = min(X0 + (1 << g.lbx), m0 - 1);
        const int ylo = max(Y0 - 1, 0), yhi = min(Y0 + (1 << g.lby), m1 - 1);
        const int zlo = max(Z0 - 1, 0), zhi = min(Z0 + (1 << g.lbz), m2 - 1);
        const int nxt = xhi - xlo + 1, nyt = yhi - ylo + 1, nrows = nyt * (zhi - zlo + 1);
        if (t < nrows) {
            const int c0 = xlo + m0 * ((ylo + t % nyt) + m1 * (zlo + t / nyt));
            const int g0 = cell_start[c0] & ~3, g1 = (cell_start[c0 + nxt] + 3) & ~3;      // whole 16-byte groups
            s_rowsrc[t] = g0;
            s_rowlen[t] = g1 - g0;
        }
        __syncthreads();
        if (t <= nrows) {
            int off = 0;
            for (int r = 0; r < t; r++) off += s_rowlen[r];
            s_rowoff[t] = off;
        }
        __syncthreads();
        const int total = s_rowoff[nrows];
        const bool fits = total <= g.tile_cap;
        if (fits) {
            for (int u = t; u < nrows * TILE_NX; u += NB_THREADS) {
                const int r = u / TILE_NX, xi = u % TILE_NX;
                if (xi <= nxt) {
                    const int c0 = xlo + m0 * ((ylo + r % nyt) + m1 * (zlo + r / nyt));
                    s_cell[u] = s_rowoff[r] + cell_start[c0 + xi] - s_rowsrc[r];
                }
            }
            if (t == 0) mbar_expect_tx(bar, (unsigned)total * 16u);
            if (t < 4 * nrows && s_rowlen[t >> 2] > 0) {
                const int r = t >> 2, arr = t & 3;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_g2s(tile_base + (unsigned)arr * TILE_STRIDE + (unsigned)s_rowoff[r] * 4u, cell_soa + (size_t)arr * soa_stride + s_rowsrc[r],
                         (unsigned)s_rowlen[r] * 4u, bar);
            }
            __syncthreads();                                    // s_cell complete
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        // ---- rows of the segment: one warp per 32-aligned group of atoms
        for (int T = (a >> 5) + wid; T <= ((e - 1) >> 5); T += NB_WARPS) {
            const int i = T * 32 + lane;
            const bool active = i >= a && i < e;
            if (!fits) {
                if (active) { pair_count[i] = -1; atomicOr(fixup, 1); }
                continue;
            }
            const float4 ci = coord4[active ? i : a];
            const int cc = cellc[active ? i : a];
            const int cx = cc & 1023, cy = (cc >> 10) & 1023, cz = cc >> 20;
            const int xa = max(cx - 1, 0) - xlo, xb = min(cx + 1, m0 - 1) - xlo + 1;
            // x - x_i instead of x_i - x: the squares, hence the distance, are bit-identical (UM/neigh_build_meso.cu:86-89)
            const unsigned long long nx2 = pack2(-ci.x, -ci.x), ny2 = pack2(-ci.y, -ci.y), nz2 = pack2(-ci.z, -ci.z);
            unsigned qp = qbase;
            for (int r = 0; r < 9; r++) {
                const int y = cy + r % 3 - 1, z = cz + r / 3 - 1;
                const bool ok = active && y >= 0 && y < m1 && z >= 0 && z < m2;
                const int row = ok ? ((z - zlo) * nyt + (y - ylo)) * TILE_NX : 0;
                const unsigned s0 = 4u * (unsigned)s_cell[row + xa];                 // byte offsets inside an array of the tile
                const unsigned s1 = ok ? 4u * (unsigned)s_cell[row + xb] : s0;
                unsigned q = s0 & ~15u;
                while (true) {
                    const bool more = q < s1;
                    if (!__any_sync(full, more)) break;
                    unsigned long long x01, x23, y01, y23, z01, z23;
                    lds_2x64(tile_base + q, x01, x23);
                    lds_2x64(tile_base + TILE_STRIDE + q, y01, y23);
                    lds_2x64(tile_base + 2u * TILE_STRIDE + q, z01, z23);
                    x01 = add2(x01, nx2); x23 = add2(x23, nx2);
                    y01 = add2(y01, ny2); y23 = add2(y23, ny2);
                    z01 = add2(z01, nz2); z23 = add2(z23, nz2);
                    const unsigned long long d01 = fma2(z01, z01, fma2(y01, y01, mul2(x01, x01)));
                    const unsigned long long d23 = fma2(z23, z23, fma2(y23, y23, mul2(x23, x23)));
                    float dr2[4];
                    unpack2(d01, dr2[0], dr2[1]);
                    unpack2(d23, dr2[2], dr2[3]);
#pragma unroll
                    for (int u = 0; u < NB_BATCH; u++) {
                        const unsigned c = q + 4u * u;
                        const bool hit = c >= s0 && c < s1 && dr2[u] <= rc2;
                        if (hit) { sts_u16(qp, c); qp = min(qp + 64u, qlim); }
                    }
                    if (more) q += 4u * NB_BATCH;
                }
            }
            __syncwarp();
            // ---- write-out: owned entries straight to the front of the row while the others are compacted in the queue,
            //      then the others behind them (both in walk order; the atom itself, always a hit, is dropped)
            int qn = (int)((qp - qbase) >> 6);
            const bool bad = qp >= qlim || qn - 1 > n_col;
            if (!active || bad) qn = 0;
            const int qmax = __reduce_max_sync(full, qn);
            int *row0 = pair_table + (size_t)(i & ~31) * (size_t)n_col + (i & 31);   // slot(i,k) = row0[(k&31)*n_col + (k>>5)*32]
            const unsigned jbase = tile_base + 3u * TILE_STRIDE;
            int own = 0;
            int po = 0;                                          // element offset of slot `own` from row0
            unsigned qo = qbase;                                 // where the next "other" entry is parked
            const int wrap = 32 - 31 * n_col;
            for (int k = 0; k < qmax; k++) {
                const uint32_t rec = k < qn ? lds_u16(qbase + (unsigned)k * 64u) : 0u;     // (stale slots may hold anything)
                const int j = (int)lds_u32(jbase + rec);
                const bool in = k < qn && j != i;
                const bool mine = in && owns(i, j, nlocal);
                if (mine) {
                    row0[po] = j;
                    po += ((own & 31) == 31) ? wrap : n_col;
                    own++;
                }
                if (in && !mine) { sts_u16(qo, rec); qo += 64u; }
            }
            const int oth = (int)((qo - qbase) >> 6);
            const int omax = __reduce_max_sync(full, oth);
            int d = own;
            for (int k = 0; k < omax; k++) {
                const int j = (int)lds_u32(jbase + (k < oth ? lds_u16(qbase + (unsigned)k * 64u) : 0u));
                if (k < oth) row0[po] = j;
                po += ((d & 31) == 31) ? wrap : n_col;
                d++;
            }
            if (active) {
                if (bad) { pair_count[i] = -1; atomicOr(fixup, 1); }
                else { pair_count[i] = own + oth; owned_count[i] = own; worst = max(worst, own + oth); }
            }
            __syncwarp();
        }
        __syncthreads();                                        // tile and cell offsets are reused by the next segment
    }
    // diagnostics only
    worst = __reduce_max_sync(full, worst);
    if (lane == 0 && worst > 0) atomicMax(&cnt->max_pair, worst);
}

// Fall-back for the rows the kernel above left (more hits than queue slots, a tile larger than its shared memory, a row wider
// than the table): plain walk of the same 9 runs in global memory, two passes (count, then write).  Also the whole build when
// MESO_NB_SLOW=1 (A/B checks of the kernel above).
__global__ void __launch_bounds__(128) k_build_rows_slow(const float4 *__restrict__ coord4, const int *__restrict__ cellc,
                                                         const int *__restrict__ cell_start, const float4 *__restrict__ cell_xyzj,
                                                         int *__restrict__ pair_count, int *__restrict__ owned_count,
                                                         int *__restrict__ pair_table, Counts *__restrict__ cnt,
                                                         const int *__restrict__ fixup, int all_rows, int n_col, float rc2, int m0, int m1, int m2)
{
    if (!all_rows && *fixup == 0) return;
    const int nlocal = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
        if (!all_rows && pair_count[i] >= 0) continue;
        const float4 ci = coord4[i];
        const int cc = cellc[i];
        const int cx = cc & 1023, cy = (cc >> 10) & 1023, cz = cc >> 20;
        const int xa = max(cx - 1, 0), xb = min(cx + 1, m0 - 1);
        int n_own = 0, n_oth = 0, k_own = 0, k_oth = 0;
        for (int pass = 0; pass < 2; pass++) {
            k_oth = n_own;
            for (int z = max(cz - 1, 0); z <= min(cz + 1, m2 - 1); z++)
                for (int y = max(cy - 1, 0); y <= min(cy + 1, m1 - 1); y++) {
                    const int *cs = cell_start + (size_t)m0 * (size_t)(y + m1 * z);
                    for (int p = cs[xa]; p < cs[xb + 1]; p++) {
                        const float4 c2 = cell_xyzj[p];
                        const int j = __float_as_int(c2.w);
                        if (j == i) continue;
                        const float dx = ci.x - c2.x, dy = ci.y - c2.y, dz = ci.z - c2.z;
                        const float dr2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                        if (!(dr2 <= rc2)) continue;
                        const bool mine = owns(i, j, nlocal);
                        if (pass == 0) { if (mine) n_own++; else n_oth++; }
                        else { const int k = mine ? k_own++ : k_oth++; if (k < n_col) pair_table[slot(i, k, n_col)] = j; }
                    }
                }
        }
        int tot = n_own + n_oth;
        if (tot > n_col) {                                   // row wider than the table: flagged, clipped (UM/neigh_build_meso.cu:242-252 only printf's)
            atomicOr(&cnt->err, 2);
            tot = n_col; n_own = min(n_own, n_col);
        }
        pair_count[i] = tot; owned_count[i] = n_own;
        atomicMax(&cnt->max_pair, tot);
    }
}

// ------------------------------------------------------------------ the reference's row order, on demand (exports)
// core entries (r <= r_n - skin at build time) in stencil order (neighbor cells by (boundary flag, Morton), ascending atom index
// inside a cell), then the skin entries in REVERSE stencil order (UM/neigh_build_meso.cu:58-117,166-200).  The distances are
// those of the build: cell_xyzj keeps the coordinates the table was built from.  key = slot rank of j's cell << 27 | j.
__global__ void __launch_bounds__(128) k_canonical_rows(const int *__restrict__ cellc, const unsigned char *__restrict__ slotrank,
                                                        const int *__restrict__ pos_of, const float4 *__restrict__ cell_xyzj,
                                                        const int *__restrict__ pair_count, const int *__restrict__ pair_table,
                                                        int *__restrict__ out_table, const Counts *__restrict__ cnt, int n_col, float rc2_core,
                                                        int m0, int m1)
{
    const int nlocal = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
        const int np = pair_count[i];
        const int cc = cellc[i];
        const int cx = cc & 1023, cy = (cc >> 10) & 1023, cz = cc >> 20;
        const unsigned char *inv = slotrank + (size_t)(cx + m0 * (cy + m1 * cz)) * 32;
        const float4 ci = cell_xyzj[pos_of[i]];
        auto key_of = [&](int j) -> uint32_t {
            const int cj = cellc[j];
            const int code = ((cj & 1023) - cx + 1) + 3 * (((cj >> 10) & 1023) - cy + 1) + 9 * ((cj >> 20) - cz + 1);
            return ((uint32_t)inv[code] << 27) | (uint32_t)j;
        };
        auto is_core = [&](int j) -> bool {
            const float4 c2 = cell_xyzj[pos_of[j]];
            const float dx = ci.x - c2.x, dy = ci.y - c2.y, dz = ci.z - c2.z;
            return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))) <= rc2_core;      // UM/neigh_build_meso.cu:296-297
        };
        int ncore = 0;
        for (int k = 0; k < np; k++) ncore += is_core(pair_table[slot(i, k, n_col)]) ? 1 : 0;
        // insertion sort of each class straight into the output row: core -> [0, ncore) ascending keys, skin -> [ncore, np) descending
        int have[2] = {0, 0};
        for (int k = 0; k < np; k++) {
            const int j = pair_table[slot(i, k, n_col)];
            const int cls = is_core(j) ? 0 : 1;
            const int dst0 = cls ? ncore : 0;
            const uint32_t key = key_of(j);
            int p = have[cls]++;
            while (p > 0) {
                const int jp = out_table[slot(i, dst0 + p - 1, n_col)];
                const uint32_t kp = key_of(jp);
                if (cls == 0 ? kp > key : kp < key) { out_table[slot(i, dst0 + p, n_col)] = jp; p--; } else break;
            }
            out_table[slot(i, dst0 + p, n_col)] = j;
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

// out[0] = 0, out[i + 1] = in[0] + ... + in[i - 1] ... see k_scan_apply; three launches on the context's stream
int scan_into(meso_ctx *ctx, const int *in, int *out, int n)
{
    const int nblk = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (!ctx->scan_sums.reserve((size_t)nblk + 8)) { ctx->err = "scan: out of device memory"; return MESO_ECUDA; }
    k_scan_sums<<<nblk, SCAN_T, 0, LS(ctx->stream)>>>(in, ctx->scan_sums.p, n);
    k_scan_top<<<1, 1024, 0, LS(ctx->stream)>>>(ctx->scan_sums.p, nblk);
    k_scan_apply<<<nblk, SCAN_T, 0, LS(ctx->stream)>>>(in, out, ctx->scan_sums.p, n);
    return MESO_OK;
}

// shape of the build for this density: the largest Morton-aligned block of cells whose tile (block + one layer) fits the
// shared-memory budget with 15 % head room, and a hit queue of 1.4 x the expected row length (three CTAs per SM at rho = 4)
static TileGeom tile_geometry(const meso_ctx *ctx, int *warps_out, size_t *smem_out)
{
    const Box &box = ctx->box;
    const double inner = (double)std::max(box.m[0] - 2, 1) * std::max(box.m[1] - 2, 1) * std::max(box.m[2] - 2, 1);
    // atoms per cell from what the rank holds now (the capacity bound of a decomposed run carries 50-100 % head room, which
    // would halve the block and double the tiles)
    const double per_cell = std::max((double)(ctx->nlocal_host > 0 ? (size_t)ctx->nlocal_host : nlocal_bound(ctx)) / inner, 1.0);
    static const int shapes[7][3] = {{2, 2, 2}, {2, 2, 1}, {2, 1, 1}, {1, 1, 1}, {1, 1, 0}, {1, 0, 0}, {0, 0, 0}};
    TileGeom g{};
    const double rn = ctx->cutneighmax;
    double vol = 1.0;
    for (int d = 0; d < 3; d++) vol *= box.binsize[d];
    const double expect = per_cell / vol * (4.0 / 3.0 * 3.14159265 * rn * rn * rn) + 1.0;
    g.nq = std::min(std::max(((int)(1.4 * expect + 4.0) + 7) / 8 * 8, 32), 1024);
    for (int s = 0; s < 7; s++) {
        g.lbx = shapes[s][0]; g.lby = shapes[s][1]; g.lbz = shapes[s][2];
        const int cells = ((1 << g.lbx) + 2) * ((1 << g.lby) + 2) * ((1 << g.lbz) + 2);
        const int rows = ((1 << g.lby) + 2) * ((1 << g.lbz) + 2);
        g.tile_cap = ((int)(cells * per_cell * 1.15) + 4 * rows + 32 + 63) / 64 * 64;     // rows are copied in whole 16-byte groups
        if (g.tile_cap <= TILE_CAP_MAX) break;
    }
    g.tile_cap = std::min(g.tile_cap, TILE_CAP_MAX);
    *smem_out = (size_t)4 * TILE_STRIDE + (size_t)NB_WARPS * g.nq * 64;
    *warps_out = NB_WARPS;
    return g;
}

int launch_neighbor_build(meso_ctx *ctx)
{
    const Box &box = ctx->box;
    const int ncell = box.ncell;
    if (ctx->cap + 8 > ((size_t)1 << 30)) { ctx->err = "neighbor build: more than 2^30 atoms + ghosts on one GPU"; return MESO_EINVAL; }
    const size_t soa_stride = (ctx->cap + 8 + 63) & ~(size_t)63;     // words per array of the cell-ordered x | y | z | index copy
    if (!ctx->cell_xyzj.reserve(ctx->cap + 8) || !ctx->pos_of.reserve(ctx->cap + 8) || !ctx->cell_soa.reserve(4 * soa_stride + 64) || !ctx->owned_count.reserve(ctx->cap) ||
        !ctx->nb_fixup.reserve(1)) {
        ctx->err = "neighbor: out of device memory";
        return MESO_ECUDA;
    }
    cudaStream_t st = ctx->stream;
    MESO_CUDA(cudaMemsetAsync(ctx->cell_cnt.p, 0, sizeof(int) * (size_t)ncell, st));
    MESO_CUDA(cudaMemsetAsync(ctx->nb_fixup.p, 0, sizeof(int), st));
    SoA3c x; for (int d = 0; d < 3; d++) x.c[d] = ctx->x[d].p;
    k_bin_count<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(x, ctx->cell_of.p, ctx->cell_cnt.p, ctx->d_counts, box);
    if (int rc = scan_into(ctx, ctx->cell_cnt.p, ctx->cell_start.p, ncell)) return rc;
    k_bin_fill<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(ctx->cell_of.p, ctx->cell_start.p + 1, ctx->cell_atoms.p, ctx->d_counts, box.m[0], box.m[1]);
    k_cell_order<<<(ncell + 127) / 128, 128, 0, LS(st)>>>(ctx->cell_start.p, ctx->cell_atoms.p, ncell);
    k_cell_records<<<grid_for(ctx, 8), 256, 0, LS(st)>>>(ctx->cell_atoms.p, ctx->coord4.p, ctx->cell_xyzj.p, ctx->pos_of.p, ctx->cell_soa.p, soa_stride,
                                                        ctx->d_counts);
    const float rc2 = (float)pow(ctx->cutneighmax, 2.0);
    const size_t nbound = nlocal_bound(ctx);
    const int slow_grid = std::max(1, std::min((int)((nbound + 127) / 128) + 1, ctx->sm_count * 64));
    if (!ctx->nb_slow) {
        int warps; size_t smem;
        const TileGeom g = tile_geometry(ctx, &warps, &smem);
        if (!ctx->nb_smem_optin) {                          // per device (a gang holds one context per device in this process)
            MESO_CUDA(cudaFuncSetAttribute(k_build_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ctx->nb_smem_optin = true;
        }
        if (smem > (size_t)200 * 1024) { ctx->err = "neighbor build: hit queues do not fit shared memory (density too high)"; return MESO_EINVAL; }
        const int grid = std::max(1, (int)((nbound + NB_THREADS - 1) / NB_THREADS));
        k_build_tiles<<<grid, NB_THREADS, smem, LS(st)>>>(ctx->coord4.p, ctx->cell_of.p, ctx->cell_start.p, ctx->cell_soa.p, soa_stride,
                                                         ctx->pair_count.p, ctx->owned_count.p, ctx->pair_table.p, ctx->d_counts, ctx->nb_fixup.p, ctx->n_col, rc2,
                                                         box.m[0], box.m[1], box.m[2], g);
    }
    k_build_rows_slow<<<slow_grid, 128, 0, LS(st)>>>(ctx->coord4.p, ctx->cell_of.p, ctx->cell_start.p, ctx->cell_xyzj.p, ctx->pair_count.p,
                                                     ctx->owned_count.p, ctx->pair_table.p, ctx->d_counts, ctx->nb_fixup.p, ctx->nb_slow ? 1 : 0,
                                                     ctx->n_col, rc2, box.m[0], box.m[1], box.m[2]);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// the table in the reference's row order, for meso_export_pair_table
int launch_canonical_rows(meso_ctx *ctx, int *out_table)
{
    const float rc2_core = (float)pow(ctx->cutneighmax - ctx->skin, 2.0);   // UM/neigh_build_meso.cu:296-297
    k_canonical_rows<<<grid_for(ctx, 8), 128, 0, LS(ctx->stream)>>>(ctx->cell_of.p, ctx->slotrank.p, ctx->pos_of.p, ctx->cell_xyzj.p, ctx->pair_count.p,
                                                                   ctx->pair_table.p, out_table, ctx->d_counts, ctx->n_col, rc2_core, ctx->box.m[0],
                                                                   ctx->box.m[1]);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
