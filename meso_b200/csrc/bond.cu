// bond.cu -- bead-spring topology on the device: per-atom bond table, tag -> index map, harmonic bonds,
// 1-2 exclusion filter on the neighbor table.
//
// Reference path:
//   AtomVecDPDBond (per-atom nbond + int2{partner tag, type} table)   UM/atom_vec_dpd_bond_meso.h:15-47
//   MesoAtom::map_set_device / gpu_set_map (tag -> lowest index)      UM/atom_meso.cu:74-150
//   MesoNeighbor::bond_all / gpu_map_bond                             UM/neighbor_meso.cu:86-159
//   gpu_bond_harmonic<evflag>                                         UM/bond_harmonic_meso.cu:46-117
//   gpu_filter_exclusion                                              UM/neigh_build_meso.cu:497-569
// The table is column-major ([slot][atom], coalesced over atoms) and travels with its atom through the
// reorder gather.  Partners are resolved once per rebuild through a direct tag -> index array (atomicMin: a
// local atom wins over its periodic images, the minimum image then undoes the wrap -- exactly the
// reference's rule).  On a decomposition the table rides the migration messages (comm.cu: k_mr_exch_bonds_scatter /
// k_mr_exch_bonds_unpack) and partners across a brick face are found among the ghosts.
#include "internal.h"
#include "device_math.cuh"
#include <utility>

namespace meso {

struct SoA3 { double *c[3]; };

__global__ void __launch_bounds__(256) k_map_set(const int *__restrict__ tag, unsigned *__restrict__ map, const Counts *__restrict__ cnt,
                                                 int map_size)
{
    const int nall = cnt->nlocal + cnt->nghost;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x) {
        const int t = tag[i];
        if (t >= 0 && t < map_size) atomicMin(map + t, (unsigned)i);        // gpu_set_map, UM/atom_meso.cu:74-82
    }
}

__global__ void __launch_bounds__(256) k_map_bonds(const int *__restrict__ nbond, const int2 *__restrict__ bonds,
                                                   int2 *__restrict__ mapped, const unsigned *__restrict__ map, Counts *__restrict__ cnt,
                                                   size_t padding, int map_size)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int nb = nbond[i];
        for (int p = 0; p < nb; p++) {
            int2 b = bonds[i + p * padding];
            const unsigned j = (b.x >= 0 && b.x < map_size) ? map[b.x] : 0xffffffffu;
            if (j == 0xffffffffu) { atomicOr(&cnt->err, 32); b.x = i; }        // "Bond atoms missing": partner neither local nor ghost
            else b.x = (int)j;
            mapped[i + p * padding] = b;
        }
    }
}

// gather of the bond table into the sorted order (rides MesoAtomVec::transfer_post_sort, UM/atom_meso.cu:174-183)
__global__ void __launch_bounds__(256) k_gather_bonds(const int *__restrict__ nbond, const int2 *__restrict__ bonds,
                                                      int *__restrict__ nbond_o, int2 *__restrict__ bonds_o,
                                                      const int *__restrict__ perm_from, const Counts *__restrict__ cnt, size_t padding,
                                                      int bond_per_atom)
{
    const int n = cnt->nlocal;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int o = perm_from[p];
        const int nb = nbond[o];
        nbond_o[p] = nb;
        for (int q = 0; q < bond_per_atom; q++)
            if (q < nb) bonds_o[p + q * padding] = bonds[o + q * padding];
    }
}

__device__ __forceinline__ double minimum_image(double dr, double p)       // UM/math_meso.h:148-152
{
    const double p_half = p * 0.5;
    return dr + (dr > -p_half ? (dr < p_half ? 0.0 : -p) : p);
}

// gpu_bond_harmonic, UM/bond_harmonic_meso.cu:46-117: fp64 on the packed fp32 coordinates, each atom sums its own
// bond entries (newton off: the partner holds the mirrored entry), F = 2 k (r - r0) / r along r_j - r_i.
// ACC = 0: f += F (fp64);  ACC = 1: facc += F (fp32 accumulator of the pair-once loop; the pair kernel has finished)
template <int EV, int ACC>
__global__ void __launch_bounds__(256) k_bond_harmonic(const float4 *__restrict__ coord4, const int *__restrict__ nbond,
                                                       const int2 *__restrict__ mapped, SoA3 f, float4 *__restrict__ facc,
                                                       double *__restrict__ virial, double *__restrict__ e_bond,
                                                       const double *__restrict__ kk, const double *__restrict__ r0,
                                                       const Counts *__restrict__ cnt, size_t padding, size_t cap, double3 period)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int nb = nbond[i];
        if (nb == 0) { if (EV) e_bond[i] = 0.; continue; }
        const float4 c1 = coord4[i];
        double fx = 0., fy = 0., fz = 0., e = 0.;
        for (int p = 0; p < nb; p++) {
            const int2 b = mapped[i + p * padding];
            const float4 c2 = coord4[b.x];
            const double dx = minimum_image((double)__fsub_rn(c2.x, c1.x), period.x);
            const double dy = minimum_image((double)__fsub_rn(c2.y, c1.y), period.y);
            const double dz = minimum_image((double)__fsub_rn(c2.z, c1.z), period.z);
            const double rsq = dx * dx + dy * dy + dz * dz;
            const double rinv = rsqrt(rsq);
            const double r = rinv * rsq;
            const double dr = r - r0[b.y];
            const double fbond = 2.0 * kk[b.y] * dr * rinv;
            fx += dx * fbond; fy += dy * fbond; fz += dz * fbond;
            if (EV) e += kk[b.y] * dr * dr;
        }
        if (ACC) {
            float4 a = facc[i];
            a.x += (float)fx; a.y += (float)fy; a.z += (float)fz;
            facc[i] = a;
        } else {
            f.c[0][i] += fx; f.c[1][i] += fy; f.c[2][i] += fz;
        }
        if (EV) {                                                              // UM/bond_harmonic_meso.cu:107-115
            virial[0 * cap + i] += (double)c1.x * fx; virial[1 * cap + i] += (double)c1.y * fy; virial[2 * cap + i] += (double)c1.z * fz;
            virial[3 * cap + i] += (double)c1.x * fy; virial[4 * cap + i] += (double)c1.x * fz; virial[5 * cap + i] += (double)c1.y * fz;
            e_bond[i] = e * 0.5;
        }
    }
}

// gpu_filter_exclusion (UM/neigh_build_meso.cu:497-544) for special_bonds lj 0 x x: the bonded partners (by tag) leave
// the row, the order of the survivors is kept.  One thread per row of the tile-transposed table; the row's two parts
// [owned][other] are compacted in place and the two count arrays follow.
__global__ void __launch_bounds__(128) k_filter_exclusion(const int *__restrict__ tag, const int *__restrict__ nbond,
                                                          const int2 *__restrict__ bonds, int *__restrict__ pair_count,
                                                          int *__restrict__ owned_count, int *__restrict__ pair_table,
                                                          const Counts *__restrict__ cnt, size_t padding, int n_col)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int nb = nbond[i];
        if (nb == 0) continue;
        int *row0 = pair_table + (size_t)(i & ~31) * (size_t)n_col + (i & 31);
        const int np = pair_count[i], nown = owned_count[i];
        const int end[2] = {nown, np};                                     // ends of the two parts in the unfiltered row
        int kept[2] = {0, 0};
        int keep = 0, k = 0;
        for (int seg = 0; seg < 2; seg++) {
            for (; k < end[seg]; k++) {
                const int j = row0[(k & 31) * n_col + (k & ~31)];
                const int t = tag[j];
                bool ok = true;
                for (int p = 0; p < nb; p++) ok = ok && bonds[i + p * padding].x != t;
                if (ok) { if (keep != k) row0[(keep & 31) * n_col + (keep & ~31)] = j; keep++; kept[seg]++; }
            }
        }
        pair_count[i] = keep;
        owned_count[i] = kept[0];
    }
}

__global__ void __launch_bounds__(256) k_sum_partial(const double *__restrict__ a, const Counts *__restrict__ cnt, double *__restrict__ partial)
{
    __shared__ double sm[8];
    const int n = cnt->nlocal;
    double s = 0.;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += a[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int w = 0; w < 8; w++) t += sm[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) k_sum_final(const double *__restrict__ partial, int nb, double *__restrict__ out)
{
    __shared__ double sm[8];
    double s = 0.;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) s += partial[b];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int w = 0; w < 8; w++) t += sm[w];
        *out = t;
    }
}

// ------------------------------------------------------------------ host drivers
bool bonds_active(const meso_ctx *ctx) { return ctx->bond_per_atom > 0 && ctx->nbondtypes > 0; }

int bonds_reserve(meso_ctx *ctx)
{
    const size_t n = (size_t)ctx->cap * (size_t)ctx->bond_per_atom;
    if (!ctx->nbond.reserve(ctx->cap) || !ctx->nbond_alt.reserve(ctx->cap) || !ctx->bonds.reserve(n) || !ctx->bonds_alt.reserve(n) ||
        !ctx->bonds_mapped.reserve(n) || !ctx->e_bond.reserve(ctx->cap)) {
        ctx->err = "out of device memory (bond table)";
        return MESO_ECUDA;
    }
    return MESO_OK;
}

// after sort_local's permutation is known (perm_from) and before the arrays are used again
int launch_bonds_gather(meso_ctx *ctx)
{
    if (!bonds_active(ctx)) return MESO_OK;
    k_gather_bonds<<<grid_for(ctx, 4), 256, 0, LS(ctx->stream)>>>(ctx->nbond.p, ctx->bonds.p, ctx->nbond_alt.p, ctx->bonds_alt.p, ctx->perm_from.p,
                                                           ctx->d_counts, ctx->cap, ctx->bond_per_atom);
    std::swap(ctx->nbond.p, ctx->nbond_alt.p); std::swap(ctx->nbond.cap, ctx->nbond_alt.cap);
    std::swap(ctx->bonds.p, ctx->bonds_alt.p); std::swap(ctx->bonds.cap, ctx->bonds_alt.cap);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// map_set_device + bond_all after the ghosts exist (UM/mvv_meso.cu:316, UM/neighbor_meso.cu:136-159)
int launch_bonds_map(meso_ctx *ctx)
{
    if (!bonds_active(ctx)) return MESO_OK;
    const int map_size = ctx->map_tag_max + 1;
    if (!ctx->tag_map.reserve((size_t)map_size)) { ctx->err = "out of device memory (tag map)"; return MESO_ECUDA; }
    MESO_CUDA(cudaMemsetAsync(ctx->tag_map.p, 0xff, sizeof(unsigned) * (size_t)map_size, ctx->stream));
    k_map_set<<<grid_for(ctx, 4), 256, 0, LS(ctx->stream)>>>(ctx->tag.p, ctx->tag_map.p, ctx->d_counts, map_size);
    k_map_bonds<<<grid_for(ctx, 4), 256, 0, LS(ctx->stream)>>>(ctx->nbond.p, ctx->bonds.p, ctx->bonds_mapped.p, ctx->tag_map.p, ctx->d_counts,
                                                        ctx->cap, map_size);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_bonds_filter(meso_ctx *ctx)
{
    if (!bonds_active(ctx) || ctx->special_lj12 != 0.0) return MESO_OK;
    k_filter_exclusion<<<grid_for(ctx, 8), 128, 0, LS(ctx->stream)>>>(ctx->tag.p, ctx->nbond.p, ctx->bonds.p, ctx->pair_count.p, ctx->owned_count.p,
                                                               ctx->pair_table.p, ctx->d_counts, ctx->cap, ctx->n_col);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_bond_force(meso_ctx *ctx, int evflag, bool into_facc)
{
    if (!bonds_active(ctx)) return MESO_OK;
    SoA3 f;
    for (int d = 0; d < 3; d++) f.c[d] = ctx->f[d].p;
    const Box &b = ctx->box;
    const double3 period = make_double3(b.periodic[0] ? b.prd[0] : 0., b.periodic[1] ? b.prd[1] : 0., b.periodic[2] ? b.prd[2] : 0.);
#define MESO_BOND_ARGS ctx->coord4.p, ctx->nbond.p, ctx->bonds_mapped.p, f, ctx->facc.p, ctx->virial.p, ctx->e_bond.p, ctx->bond_k_dev.p, \
                       ctx->bond_r0_dev.p, ctx->d_counts, ctx->cap, ctx->cap, period
    const int g = grid_for(ctx, 8);
    if (evflag) k_bond_harmonic<1, 0><<<g, 256, 0, LS(ctx->stream)>>>(MESO_BOND_ARGS);
    else if (into_facc) k_bond_harmonic<0, 1><<<g, 256, 0, LS(ctx->stream)>>>(MESO_BOND_ARGS);
    else k_bond_harmonic<0, 0><<<g, 256, 0, LS(ctx->stream)>>>(MESO_BOND_ARGS);
#undef MESO_BOND_ARGS
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_bond_energy_sum(meso_ctx *ctx, double *e)
{
    *e = 0.;
    if (!bonds_active(ctx)) return MESO_OK;
    const int nb = grid_for(ctx, 2);
    if (!ctx->partial.reserve((size_t)nb + 8)) { ctx->err = "reduce: out of device memory"; return MESO_ECUDA; }
    k_sum_partial<<<nb, 256, 0, LS(ctx->stream)>>>(ctx->e_bond.p, ctx->d_counts, ctx->partial.p);
    k_sum_final<<<1, 256, 0, LS(ctx->stream)>>>(ctx->partial.p, nb, ctx->partial.p + nb);
    MESO_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->partial.p + nb, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    *e = ctx->h_result[0];
    return MESO_OK;
}

}  // namespace meso
