// reorder.cu -- on-device PBC wrap, 2-level Morton reorder, ghost (halo) creation and refresh.
//
// Reference path (all of it host-side or PCIe-staged there):
//   MesoDomain::pbc                 UM/domain_meso.cu:30-145      (OpenMP on host arrays)
//   MesoComm::borderness            UM/comm_meso.cu:188-254       (OpenMP on host arrays)
//   gpu_build_reorder_keypair       UM/atom_meso.cu:268-308       (reads borderness through mapped host memory)
//   MesoAtom::sort_local            UM/atom_meso.cu:343-384
//   transfer_post_sort (permuting copy) UM/atom_meso.cu:174-183, UM/atom_vec_meso.h:11-104
//   MesoComm::borders + pack/unpack_border_vel  UM/comm_meso.cu:41-186, UM/atom_vec_dpd_atomic_meso.cu:61-163
//   Comm::forward_comm + pack/unpack_comm_vel   src/comm.cpp:686-753, UM/atom_vec_dpd_atomic_meso.cu:165-243
// Here everything stays in HBM: one kernel wraps + classifies + keys, the radix sort of
// sort.cu orders the keys, one gather kernel permutes every attribute AND emits the packed
// float4 views, and ghosts of a periodic single-rank box are built by ordered stream
// compaction in exactly the reference's 6-swap order (so atom indices, hence neighbor-list
// order, are bit-identical to the reference's).
#include "internal.h"
#include "device_math.cuh"

namespace meso {

int sort_pairs_u64(meso_ctx *ctx, DevBuf<uint64_t> &key, DevBuf<int> &val, const int *d_n, size_t cap, int bits, int low_bits);

struct SoA3 { double *c[3]; };
struct SoA3c { const double *c[3]; };

// ------------------------------------------------------------------ wrap + key
__global__ void __launch_bounds__(256) k_reorder_key(SoA3 x, int *__restrict__ image, uint64_t *__restrict__ key,
                                                     int *__restrict__ val, const Counts *__restrict__ cnt, Box box,
                                                     int l1_shift, uint64_t border_mask)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int img = image[i];
        uint32_t b[3], s[3];
        bool border = false;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double xd = x.c[d][i];
            if (box.periodic[d]) {                       // UM/domain_meso.cu:56-79
                const int sh = 10 * d;
                if (xd < box.boxlo[d]) {
                    xd += box.prd[d];
                    int idim = (img >> sh) & 1023;
                    img = (img ^ (idim << sh)) | (((idim - 1) & 1023) << sh);
                }
                if (xd >= box.boxhi[d]) {
                    xd -= box.prd[d];
                    xd = fmax(xd, box.boxlo[d]);
                    int idim = (img >> sh) & 1023;
                    img = (img ^ (idim << sh)) | (((idim + 1) & 1023) << sh);
                }
                x.c[d][i] = xd;
            }
            // cell and 4-bit sub-cell, UM/atom_meso.cu:288-293 (FMA as nvcc contracts a*b+c)
            b[d] = (uint32_t)clamp_rz(__fma_rn(xd - box.sublo[d], box.bininv[d], 1.0), 0, box.m[d]);
            double ci = __dmul_rn(16.0, box.bininv[d]);
            s[d] = (uint32_t)clamp_rz(__dmul_rn(__fma_rn(-(double)(uint32_t)(b[d] - 1u), box.binsize[d], xd), ci), 0, 16);
            // borderness: inside any send slab, UM/comm_meso.h:71-74 with slabs of src/comm.cpp:590-625
            border |= (box.sendflag[2 * d] && xd <= box.slab_lo_hi[d]) || (box.sendflag[2 * d + 1] && xd >= box.slab_hi_lo[d]);
        }
        image[i] = img;
        uint64_t k = ((uint64_t)morton3(b[0], b[1], b[2]) << l1_shift) | (uint64_t)morton3(s[0], s[1], s[2]);
        if (border) k |= border_mask;
        key[i] = k;
        val[i] = i;
    }
}

// periodic wrap + image flags only (Domain::pbc before Comm::exchange, UM/mvv_meso.cu:283-290)
__global__ void __launch_bounds__(256) k_pbc(SoA3 x, int *__restrict__ image, const Counts *__restrict__ cnt, Box box)
{
    const int n = cnt->nlocal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int img = image[i];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            if (!box.periodic[d]) continue;
            double xd = x.c[d][i];
            const int sh = 10 * d;
            if (xd < box.boxlo[d]) {
                xd += box.prd[d];
                int idim = (img >> sh) & 1023;
                img = (img ^ (idim << sh)) | (((idim - 1) & 1023) << sh);
            }
            if (xd >= box.boxhi[d]) {
                xd -= box.prd[d];
                xd = fmax(xd, box.boxlo[d]);
                int idim = (img >> sh) & 1023;
                img = (img ^ (idim << sh)) | (((idim + 1) & 1023) << sh);
            }
            x.c[d][i] = xd;
        }
        image[i] = img;
    }
}

// ------------------------------------------------------------------ gather into sorted order (+ pack)
__global__ void __launch_bounds__(256) k_gather(SoA3c x, SoA3c v, const int *__restrict__ tag, const int *__restrict__ type,
                                                const int *__restrict__ mask, const int *__restrict__ image, SoA3 xo, SoA3 vo,
                                                int *__restrict__ tago, int *__restrict__ typeo, int *__restrict__ masko,
                                                int *__restrict__ imageo, float4 *__restrict__ coord4,
                                                float4 *__restrict__ veloc4, const uint64_t *__restrict__ key,
                                                const int *__restrict__ perm_from, Counts *__restrict__ cnt, Box box,
                                                uint64_t border_mask, uint32_t seed_now)
{
    const int n = cnt->nlocal;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int o = perm_from[p];
        double xx[3], vv[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            xx[d] = x.c[d][o]; vv[d] = v.c[d][o];
            xo.c[d][p] = xx[d]; vo.c[d][p] = vv[d];
        }
        const int tg = tag[o], ty = type[o];
        tago[p] = tg; typeo[p] = ty; masko[p] = mask[o]; imageo[p] = image[o];
        // dp2sp_merged, UM/atom_vec_meso.cu:152-166
        float4 c, w;
        c.x = (float)(xx[0] - box.centre[0]); c.y = (float)(xx[1] - box.centre[1]); c.z = (float)(xx[2] - box.centre[2]);
        c.w = __int_as_float(ty - 1);
        w.x = (float)vv[0]; w.y = (float)vv[1]; w.z = (float)vv[2];
        w.w = __uint_as_float(signature(seed_now, tg, w.x, w.y, w.z));
        coord4[p] = c; veloc4[p] = w;
        // bulk|border boundary of the sorted order, UM/atom_meso.cu:316-341
        const bool b = (key[p] & border_mask) != 0;
        const bool bprev = p > 0 ? (key[p - 1] & border_mask) != 0 : false;
        if (b && !bprev) { cnt->n_bulk = p; cnt->n_border = n - p; }
        if (p == n - 1 && !b) { cnt->n_bulk = n; cnt->n_border = 0; }
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) { cnt->n_bulk = 0; cnt->n_border = 0; }
}

// ------------------------------------------------------------------ ghosts: ordered compaction per dimension
// One dimension = two swaps sharing the candidate range [n_bulk, nall_before) (UM/comm_meso.cu:60-81).
constexpr int GH_THREADS = 256;
constexpr int GH_ITEMS = 4;
constexpr int GH_TILE = GH_THREADS * GH_ITEMS;

__device__ __forceinline__ void ghost_flags(double xd, const Box &box, int d, bool &lo, bool &hi)
{
    lo = box.sendflag[2 * d] && xd <= box.slab_lo_hi[d];
    hi = box.sendflag[2 * d + 1] && xd >= box.slab_hi_lo[d];
}

__global__ void __launch_bounds__(GH_THREADS) k_ghost_count(const double *__restrict__ xd, const Counts *__restrict__ cnt,
                                                            int2 *__restrict__ tile_counts, Box box, int d, int ntiles)
{
    __shared__ int slo[GH_THREADS / 32], shi[GH_THREADS / 32];
    const int first = cnt->n_bulk, last = cnt->nlocal + cnt->nghost;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int nlo = 0, nhi = 0;
        const int base = first + tile * GH_TILE;
#pragma unroll
        for (int r = 0; r < GH_ITEMS; r++) {
            int i = base + r * GH_THREADS + threadIdx.x;
            if (i < last) {
                bool lo, hi;
                ghost_flags(xd[i], box, d, lo, hi);
                nlo += lo; nhi += hi;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { nlo += __shfl_xor_sync(0xffffffffu, nlo, o); nhi += __shfl_xor_sync(0xffffffffu, nhi, o); }
        if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = nlo; shi[threadIdx.x >> 5] = nhi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int a = 0, b = 0;
            for (int w = 0; w < GH_THREADS / 32; w++) { a += slo[w]; b += shi[w]; }
            tile_counts[tile] = make_int2(a, b);
        }
        __syncthreads();
    }
}

// single CTA: exclusive scan of the per-tile counts, publishes the two swaps' ghost ranges
__global__ void __launch_bounds__(1024) k_ghost_scan(int2 *__restrict__ tile_counts, Counts *__restrict__ cnt, int d, int ntiles,
                                                     int cap)
{
    __shared__ int2 wsum[32];
    __shared__ int2 carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = make_int2(0, 0);
    __syncthreads();
    const int first = cnt->n_bulk, last = cnt->nlocal + cnt->nghost;
    const int used = (max(last - first, 0) + GH_TILE - 1) / GH_TILE;
    for (int base = 0; base < used; base += 1024) {
        int i = base + t;
        int2 v = (i < used) ? tile_counts[i] : make_int2(0, 0), x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int ya = __shfl_up_sync(0xffffffffu, x.x, o), yb = __shfl_up_sync(0xffffffffu, x.y, o);
            if (lane >= o) { x.x += ya; x.y += yb; }
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            int2 s = wsum[lane], z = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int ya = __shfl_up_sync(0xffffffffu, z.x, o), yb = __shfl_up_sync(0xffffffffu, z.y, o);
                if (lane >= o) { z.x += ya; z.y += yb; }
            }
            wsum[lane] = make_int2(z.x - s.x, z.y - s.y);
        }
        __syncthreads();
        int2 c = carry_s;
        int2 e = make_int2(c.x + wsum[w].x + x.x - v.x, c.y + wsum[w].y + x.y - v.y);
        if (i < used) tile_counts[i] = e;
        __syncthreads();
        if (t == 1023) carry_s = make_int2(e.x + v.x, e.y + v.y);
        __syncthreads();
    }
    if (t == 0) {
        int2 tot = carry_s;
        if (last + tot.x + tot.y > cap) { cnt->err |= 1; tot = make_int2(0, 0); }
        cnt->swap_first[2 * d] = last;          cnt->swap_n[2 * d] = tot.x;
        cnt->swap_first[2 * d + 1] = last + tot.x; cnt->swap_n[2 * d + 1] = tot.y;
        // nghost is advanced by k_ghost_advance after the scatter consumed the old value
    }
}

// shift code: 2 bits per dim, 0 = none, 1 = +prd, 2 = -prd
__device__ __forceinline__ double apply_shift(double xd, int code, int d, const Box &box)
{
    int c = (code >> (2 * d)) & 3;
    return c == 0 ? xd : (c == 1 ? xd + box.prd[d] : xd - box.prd[d]);
}

__global__ void __launch_bounds__(GH_THREADS) k_ghost_scatter(SoA3 x, SoA3 v, int *__restrict__ tag, int *__restrict__ type,
                                                              int *__restrict__ mask, float4 *__restrict__ coord4,
                                                              float4 *__restrict__ veloc4, int *__restrict__ ghost_root,
                                                              int *__restrict__ ghost_shift, const Counts *__restrict__ cnt,
                                                              const int2 *__restrict__ tile_counts, Box box, int d, int ntiles)
{
    __shared__ int2 wsum[GH_THREADS / 32];
    const int nlocal = cnt->nlocal;
    const int first = cnt->n_bulk, last = cnt->swap_first[2 * d];   // == nlocal + nghost before this dim
    const int out_lo = cnt->swap_first[2 * d], out_hi = cnt->swap_first[2 * d + 1];
    if (cnt->swap_n[2 * d] + cnt->swap_n[2 * d + 1] == 0) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int base = first + tile * GH_TILE;
        if (base >= last) break;
        int2 run = tile_counts[tile];
        // blocked order inside the tile: round r covers GH_THREADS consecutive candidates
#pragma unroll
        for (int r = 0; r < GH_ITEMS; r++) {
            int i = base + r * GH_THREADS + threadIdx.x;
            bool lo = false, hi = false;
            double xi[3];
            if (i < last) {
#pragma unroll
                for (int q = 0; q < 3; q++) xi[q] = x.c[q][i];
                ghost_flags(xi[d], box, d, lo, hi);
            }
            uint32_t blo = __ballot_sync(0xffffffffu, lo), bhi = __ballot_sync(0xffffffffu, hi);
            if (lane == 0) wsum[w] = make_int2(__popc(blo), __popc(bhi));
            __syncthreads();
            int2 pre = make_int2(0, 0), tot = make_int2(0, 0);
#pragma unroll
            for (int ww = 0; ww < GH_THREADS / 32; ww++) {
                int2 c = wsum[ww];
                if (ww < w) { pre.x += c.x; pre.y += c.y; }
                tot.x += c.x; tot.y += c.y;
            }
            __syncthreads();
#pragma unroll
            for (int side = 0; side < 2; side++) {
                bool hit = side ? hi : lo;
                if (!hit) continue;
                int g = side ? out_hi + run.y + pre.y + __popc(bhi & lt) : out_lo + run.x + pre.x + __popc(blo & lt);
                int code = (i < nlocal) ? 0 : ghost_shift[i - nlocal];
                int root = (i < nlocal) ? i : ghost_root[i - nlocal];
                int pbc = box.pbc[2 * d + side];
                if (pbc) code |= (pbc > 0 ? 1 : 2) << (2 * d);
                double xg[3] = {xi[0], xi[1], xi[2]};
                if (pbc) xg[d] = pbc > 0 ? xi[d] + box.prd[d] : xi[d] - box.prd[d];   // pack_border_vel: x + pbc*prd
#pragma unroll
                for (int q = 0; q < 3; q++) { x.c[q][g] = xg[q]; v.c[q][g] = v.c[q][i]; }
                int ty = type[i];
                tag[g] = tag[i]; type[g] = ty; mask[g] = mask[i];
                ghost_root[g - nlocal] = root; ghost_shift[g - nlocal] = code;
                float4 c;
                c.x = (float)(xg[0] - box.centre[0]); c.y = (float)(xg[1] - box.centre[1]); c.z = (float)(xg[2] - box.centre[2]);
                c.w = __int_as_float(ty - 1);
                coord4[g] = c;
                veloc4[g] = veloc4[i];
            }
            run.x += tot.x; run.y += tot.y;
        }
    }
}

__global__ void k_ghost_advance(Counts *cnt, int d)
{
    cnt->nghost += cnt->swap_n[2 * d] + cnt->swap_n[2 * d + 1];
    cnt->nall = cnt->nlocal + cnt->nghost;
}

__global__ void k_counts_reset_ghosts(Counts *cnt)
{
    cnt->nghost = 0;
    cnt->nall = cnt->nlocal;
    cnt->max_pair = 0;
    for (int s = 0; s < 6; s++) { cnt->swap_first[s] = cnt->nlocal; cnt->swap_n[s] = 0; }
}

// ------------------------------------------------------------------ per-step ghost refresh (single rank)
// ghost g mirrors local root[g] (A14).  FULL = 0 (fused run): only the packed views are refreshed,
// coord = fl((x_root + shift) - centre), veloc/signature copied.  FULL = 1 (phase API, reference order
// forward_comm -> dp2sp_merged(GHOST)): the fp64 ghost x,v are refreshed and packing is left to k_pack.
template <int FULL>
__global__ void __launch_bounds__(256) k_forward_self(SoA3 x, SoA3 v, float4 *__restrict__ coord4, float4 *__restrict__ veloc4,
                                                      const int *__restrict__ ghost_root, const int *__restrict__ ghost_shift,
                                                      const int *__restrict__ type, const Counts *__restrict__ cnt, Box box)
{
    const int nlocal = cnt->nlocal, ng = cnt->nghost;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += gridDim.x * blockDim.x) {
        const int r = ghost_root[g], code = ghost_shift[g];
        const double gx = apply_shift(x.c[0][r], code, 0, box), gy = apply_shift(x.c[1][r], code, 1, box),
                     gz = apply_shift(x.c[2][r], code, 2, box);
        if (FULL) {
            x.c[0][nlocal + g] = gx; x.c[1][nlocal + g] = gy; x.c[2][nlocal + g] = gz;
            v.c[0][nlocal + g] = v.c[0][r]; v.c[1][nlocal + g] = v.c[1][r]; v.c[2][nlocal + g] = v.c[2][r];
        } else {
            float4 c;
            c.x = (float)(gx - box.centre[0]); c.y = (float)(gy - box.centre[1]); c.z = (float)(gz - box.centre[2]);
            c.w = __int_as_float(type[r] - 1);
            coord4[nlocal + g] = c;
            veloc4[nlocal + g] = veloc4[r];
        }
    }
}

// ------------------------------------------------------------------ host drivers
static inline SoA3 soa(DevBuf<double> *b) { SoA3 s; for (int d = 0; d < 3; d++) s.c[d] = b[d].p; return s; }
static inline SoA3c soac(DevBuf<double> *b) { SoA3c s; for (int d = 0; d < 3; d++) s.c[d] = b[d].p; return s; }

int launch_pbc(meso_ctx *ctx)
{
    k_pbc<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(soa(ctx->x), ctx->image.p, ctx->d_counts, ctx->box);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_reorder(meso_ctx *ctx)
{
    const Box &box = ctx->box;
    // MesoAtom::sort_local key layout, UM/atom_meso.cu:345-360
    int max_bin = std::max(std::max(box.m[0], box.m[1]), box.m[2]);
    int l1_width = 3 * (int)floor(log2(max_bin * 2.0));
    const int l2_width = 12;
    uint64_t border_mask = 1ULL << (l1_width + l2_width);
    int bits = 1 + l1_width + l2_width;
    k_reorder_key<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(soa(ctx->x), ctx->image.p, ctx->key.p, ctx->perm_from.p, ctx->d_counts,
                                                          box, l2_width, border_mask);
    int rc = sort_pairs_u64(ctx, ctx->key, ctx->perm_from, &ctx->d_counts->nlocal, ctx->cap, bits, l2_width);
    if (rc) return rc;
    k_gather<<<grid_for(ctx, 8), 256, 0, LS(ctx->stream)>>>(soac(ctx->x), soac(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->image.p,
                                                     soa(ctx->xa), soa(ctx->va), ctx->taga.p, ctx->typea.p, ctx->maska.p,
                                                     ctx->imagea.p, ctx->coord4.p, ctx->veloc4.p, ctx->key.p, ctx->perm_from.p,
                                                     ctx->d_counts, box, border_mask, seed_now(ctx));
    for (int d = 0; d < 3; d++) { std::swap(ctx->x[d].p, ctx->xa[d].p); std::swap(ctx->v[d].p, ctx->va[d].p); }
    std::swap(ctx->tag.p, ctx->taga.p); std::swap(ctx->type.p, ctx->typea.p);
    std::swap(ctx->mask.p, ctx->maska.p); std::swap(ctx->image.p, ctx->imagea.p);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_borders(meso_ctx *ctx)
{
    const Box &box = ctx->box;
    const int ntiles = (int)((ctx->cap + GH_TILE - 1) / GH_TILE);
    if (!ctx->tile_counts.reserve((size_t)ntiles * 2)) { ctx->err = "borders: out of device memory"; return MESO_ECUDA; }
    int2 *tc = reinterpret_cast<int2 *>(ctx->tile_counts.p);
    k_counts_reset_ghosts<<<1, 1, 0, LS(ctx->stream)>>>(ctx->d_counts);
    for (int d = 0; d < 3; d++) {
        if (!box.sendflag[2 * d] && !box.sendflag[2 * d + 1]) continue;
        k_ghost_count<<<grid_for(ctx, 4), GH_THREADS, 0, LS(ctx->stream)>>>(ctx->x[d].p, ctx->d_counts, tc, box, d, ntiles);
        k_ghost_scan<<<1, 1024, 0, LS(ctx->stream)>>>(tc, ctx->d_counts, d, ntiles, (int)ctx->cap);
        k_ghost_scatter<<<grid_for(ctx, 4), GH_THREADS, 0, LS(ctx->stream)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p,
                                                                       ctx->coord4.p, ctx->veloc4.p, ctx->ghost_root.p,
                                                                       ctx->ghost_shift.p, ctx->d_counts, tc, box, d, ntiles);
        k_ghost_advance<<<1, 1, 0, LS(ctx->stream)>>>(ctx->d_counts, d);
    }
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

int launch_forward(meso_ctx *ctx, bool full)
{
    if (full)
        k_forward_self<1><<<grid_for(ctx, 4), 256, 0, LS(ctx->stream)>>>(soa(ctx->x), soa(ctx->v), ctx->coord4.p, ctx->veloc4.p, ctx->ghost_root.p,
                                                                  ctx->ghost_shift.p, ctx->type.p, ctx->d_counts, ctx->box);
    else
        k_forward_self<0><<<grid_for(ctx, 4), 256, 0, LS(ctx->stream)>>>(soa(ctx->x), soa(ctx->v), ctx->coord4.p, ctx->veloc4.p, ctx->ghost_root.p,
                                                                  ctx->ghost_shift.p, ctx->type.p, ctx->d_counts, ctx->box);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso
