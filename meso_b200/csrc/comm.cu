// comm.cu -- spatial decomposition across GPUs: migration, ghost creation and the per-step halo
// refresh, with NCCL send/recv over NVLink.
//
// Reference path (host MPI through pinned staging buffers, every step):
//   MesoComm::exchange                UM/comm_meso.cu:256-420        (migration, per dimension)
//   MesoComm::borders                 UM/comm_meso.cu:41-186         (ghost creation, 6 swaps)
//   Comm::forward_comm                src/comm.cpp:686-753           (ghost x,v refresh, 6 swaps)
//   pack/unpack_{border,comm}_vel     UM/atom_vec_dpd_atomic_meso.cu:61-243
//   MPI_Allreduce of thermo scalars   UM/compute_temp_meso.cu:97
// The swap structure (3 dimensions x {to-lower, to-upper}, later dimensions forwarding earlier
// ghosts) is kept, because it fixes the ghost ORDER and with it the neighbor-list order.  What
// changes: selection, packing and unpacking are device kernels (ordered stream compaction, no
// atomics); messages are fixed-capacity with the record count in a header, so neither side ever
// needs a host round trip to size a message -- the whole rebuild and every step stay asynchronous;
// a swap whose partner is this rank itself (procgrid[d] == 1) unpacks straight from its own send
// buffer; the per-step refresh runs on a side stream so that it overlaps the bulk force kernel.
#include "internal.h"
#include "device_math.cuh"
#include <nccl.h>
#include <algorithm>
#include <climits>
#include <cstring>
#include <vector>

namespace meso {

#define MESO_NCCL(call)                                                                   \
    do {                                                                                  \
        ncclResult_t r_ = (call);                                                         \
        if (r_ != ncclSuccess) {                                                          \
            ctx->err = std::string(#call) + ": " + ncclGetErrorString(r_);                \
            return MESO_ENCCL;                                                            \
        }                                                                                 \
    } while (0)

struct SoA3 { double *c[3]; };
static inline SoA3 soa(DevBuf<double> *b) { SoA3 s; for (int d = 0; d < 3; d++) s.c[d] = b[d].p; return s; }

constexpr int REC = 8;            // doubles per migration record (64 B); record 0 of every message is the header {count}
constexpr int RECB = 9;           // doubles per border record: + {owner rank | shift code << 24, owner's local index}
constexpr int CT = 256;           // threads
constexpr int CI = 4;             // items per thread
constexpr int CTILE = CT * CI;

int comm_init(meso_ctx *ctx, const void *nccl_id)
{
    if (ctx->nccl) return MESO_OK;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    memcpy(&id, nccl_id, sizeof id);
    MESO_CUDA(cudaSetDevice(ctx->device));
    ncclComm_t comm;
    MESO_NCCL(ncclCommInitRank(&comm, ctx->nranks, id, ctx->rank));
    ctx->nccl = comm;
    return MESO_OK;
}

void comm_destroy(meso_ctx *ctx)
{
    if (ctx->nccl) { ncclCommDestroy((ncclComm_t)ctx->nccl); ctx->nccl = nullptr; }
}

// small host-visible reductions (thermo scalars, global atom count)
int comm_allreduce_sum(meso_ctx *ctx, double *host_vals, int n)
{
    if (ctx->nranks == 1) return MESO_OK;
    if (!ctx->nccl) { ctx->err = "communicator not initialised"; return MESO_ENCCL; }
    if (!ctx->reduce_buf.reserve(std::max(64, n))) { ctx->err = "out of device memory"; return MESO_ECUDA; }
    MESO_CUDA(cudaMemcpyAsync(ctx->reduce_buf.p, host_vals, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    MESO_NCCL(ncclAllReduce(ctx->reduce_buf.p, ctx->reduce_buf.p, n, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
    MESO_CUDA(cudaMemcpyAsync(host_vals, ctx->reduce_buf.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return MESO_OK;
}

// ------------------------------------------------------------------ generic ordered 2-way compaction
// flags a,b per candidate in [first,last); tile counts -> exclusive scan -> ranks.  Used for (lo,hi) slabs
// and for (left,right) leavers.
struct Flags { bool a, b; };

template <typename F>
__device__ __forceinline__ void tile_count(F flag, int first, int last, int2 *tile_counts, int ntiles)
{
    __shared__ int sa[CT / 32], sb[CT / 32];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int na = 0, nb = 0;
        const int base = first + tile * CTILE;
#pragma unroll
        for (int r = 0; r < CI; r++) {
            const int i = base + r * CT + threadIdx.x;
            if (i < last) { Flags f = flag(i); na += f.a; nb += f.b; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { na += __shfl_xor_sync(0xffffffffu, na, o); nb += __shfl_xor_sync(0xffffffffu, nb, o); }
        if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = na; sb[threadIdx.x >> 5] = nb; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int a = 0, b = 0;
            for (int w = 0; w < CT / 32; w++) { a += sa[w]; b += sb[w]; }
            tile_counts[tile] = make_int2(a, b);
        }
        __syncthreads();
    }
}

// single CTA: exclusive scan of `used` tile counts; totals to tot[0..1]
__device__ __forceinline__ int2 scan_tiles(int2 *tile_counts, int used)
{
    __shared__ int2 wsum[32];
    __shared__ int2 carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = make_int2(0, 0);
    __syncthreads();
    for (int base = 0; base < used; base += 1024) {
        const int i = base + t;
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
        const int2 c = carry_s;
        const int2 e = make_int2(c.x + wsum[w].x + x.x - v.x, c.y + wsum[w].y + x.y - v.y);
        if (i < used) tile_counts[i] = e;
        __syncthreads();
        if (t == 1023) carry_s = make_int2(e.x + v.x, e.y + v.y);
        __syncthreads();
    }
    return carry_s;
}

// rank of candidate i inside its tile for both flags (blocked order: round r covers CT consecutive candidates)
struct TileRank {
    int2 run;       // running totals of previous rounds in this tile
};

// ------------------------------------------------------------------ ghosts (borders)
__device__ __forceinline__ Flags slab_flags(double xd, const Box &box, int d)
{
    Flags f;
    f.a = box.sendflag[2 * d] && xd <= box.slab_lo_hi[d];
    f.b = box.sendflag[2 * d + 1] && xd >= box.slab_hi_lo[d];
    return f;
}

__global__ void __launch_bounds__(CT) k_mr_border_count(const double *__restrict__ xd, const Counts *__restrict__ cnt,
                                                        int2 *__restrict__ tile_counts, Box box, int d, int ntiles)
{
    const int first = cnt->n_bulk, last = cnt->nlocal + cnt->nghost;
    tile_count([&](int i) { return slab_flags(xd[i], box, d); }, first, last, tile_counts, ntiles);
}

__global__ void __launch_bounds__(1024) k_mr_border_scan(int2 *__restrict__ tile_counts, Counts *__restrict__ cnt, double *__restrict__ send_lo,
                                                         double *__restrict__ send_hi, int d, int swap_cap)
{
    const int first = cnt->n_bulk, last = cnt->nlocal + cnt->nghost;
    const int used = (max(last - first, 0) + CTILE - 1) / CTILE;
    int2 tot = scan_tiles(tile_counts, used);
    if (threadIdx.x == 0) {
        if (tot.x > swap_cap || tot.y > swap_cap) { cnt->err |= 1; tot.x = min(tot.x, swap_cap); tot.y = min(tot.y, swap_cap); }
        cnt->send_n[2 * d] = tot.x; cnt->send_n[2 * d + 1] = tot.y;
        reinterpret_cast<int *>(send_lo)[0] = tot.x;           // message headers
        reinterpret_cast<int *>(send_hi)[0] = tot.y;
    }
}

// pack_border_vel (UM/atom_vec_dpd_atomic_meso.cu:61-100): record = {x+shift (3), v (3), tag|type, mask|signature}
__global__ void __launch_bounds__(CT) k_mr_border_pack(SoA3 x, SoA3 v, const int *__restrict__ tag, const int *__restrict__ type,
                                                       const int *__restrict__ mask, const float4 *__restrict__ veloc4,
                                                       const Counts *__restrict__ cnt, const int2 *__restrict__ tile_counts,
                                                       double *__restrict__ send_lo, double *__restrict__ send_hi,
                                                       int *__restrict__ list_lo, int *__restrict__ list_hi, Box box, int d, int ntiles,
                                                       int swap_cap, const int2 *__restrict__ ghost_origin, int my_rank)
{
    __shared__ int2 wsum[CT / 32];
    const int nlocal = cnt->nlocal;
    const int first = cnt->n_bulk, last = nlocal + cnt->nghost;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int base = first + tile * CTILE;
        if (base >= last) break;
        int2 run = tile_counts[tile];
#pragma unroll
        for (int r = 0; r < CI; r++) {
            const int i = base + r * CT + threadIdx.x;
            Flags f{false, false};
            double xi[3] = {0, 0, 0};
            if (i < last) {
#pragma unroll
                for (int q = 0; q < 3; q++) xi[q] = x.c[q][i];
                f = slab_flags(xi[d], box, d);
            }
            const uint32_t ba = __ballot_sync(0xffffffffu, f.a), bb = __ballot_sync(0xffffffffu, f.b);
            if (lane == 0) wsum[w] = make_int2(__popc(ba), __popc(bb));
            __syncthreads();
            int2 pre = make_int2(0, 0), tot = make_int2(0, 0);
#pragma unroll
            for (int ww = 0; ww < CT / 32; ww++) {
                const int2 c = wsum[ww];
                if (ww < w) { pre.x += c.x; pre.y += c.y; }
                tot.x += c.x; tot.y += c.y;
            }
            __syncthreads();
#pragma unroll
            for (int side = 0; side < 2; side++) {
                if (!(side ? f.b : f.a)) continue;
                const int k = side ? run.y + pre.y + __popc(bb & lt) : run.x + pre.x + __popc(ba & lt);
                if (k >= swap_cap) continue;
                const int pbc = box.pbc[2 * d + side];
                double xg[3] = {xi[0], xi[1], xi[2]};
                if (pbc) xg[d] = pbc > 0 ? xi[d] + box.prd[d] : xi[d] - box.prd[d];   // x + pbc*prd
                double *rec = (side ? send_hi : send_lo) + (size_t)(k + 1) * RECB;
                // owner of this image: a local atom is its own, a forwarded ghost keeps the owner it arrived with;
                // shift code 2 bits per dimension (1 = +prd, 2 = -prd), accumulated over the swaps it went through
                int2 org = (i < nlocal) ? make_int2(my_rank, i) : ghost_origin[i - nlocal];
                if (pbc) org.x |= (pbc > 0 ? 1 : 2) << (24 + 2 * d);
                reinterpret_cast<int2 *>(rec)[8] = org;
                rec[0] = xg[0]; rec[1] = xg[1]; rec[2] = xg[2];
                rec[3] = v.c[0][i]; rec[4] = v.c[1][i]; rec[5] = v.c[2][i];
                reinterpret_cast<int2 *>(rec)[6] = make_int2(tag[i], type[i]);
                reinterpret_cast<int2 *>(rec)[7] = make_int2(mask[i], __float_as_int(veloc4[i].w));
                (side ? list_hi : list_lo)[k] = i;
            }
            run.x += tot.x; run.y += tot.y;
        }
    }
}

// publishes the ghost ranges of the two swaps of dimension d from the received headers
__global__ void k_mr_border_advance(Counts *cnt, const double *recv_a, const double *recv_b, int d, int cap)
{
    int na = reinterpret_cast<const int *>(recv_a)[0], nb = reinterpret_cast<const int *>(recv_b)[0];
    const int last = cnt->nlocal + cnt->nghost;
    if (last + na + nb > cap) { cnt->err |= 1; na = 0; nb = 0; }
    cnt->swap_first[2 * d] = last;          cnt->swap_n[2 * d] = na;
    cnt->swap_first[2 * d + 1] = last + na; cnt->swap_n[2 * d + 1] = nb;
    cnt->nghost += na + nb;
    cnt->nall = cnt->nlocal + cnt->nghost;
}

// unpack_border_vel (UM/atom_vec_dpd_atomic_meso.cu:137-163) + packing of the new ghosts
__global__ void __launch_bounds__(256) k_mr_border_unpack(SoA3 x, SoA3 v, int *__restrict__ tag, int *__restrict__ type, int *__restrict__ mask,
                                                          float4 *__restrict__ coord4, float4 *__restrict__ veloc4,
                                                          const Counts *__restrict__ cnt, const double *__restrict__ recv_a,
                                                          const double *__restrict__ recv_b, Box box, int d, int2 *__restrict__ ghost_origin)
{
    const int na = cnt->swap_n[2 * d], n = na + cnt->swap_n[2 * d + 1], g0 = cnt->swap_first[2 * d];
    const int nlocal = cnt->nlocal;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double *rec = (k < na) ? recv_a + (size_t)(k + 1) * RECB : recv_b + (size_t)(k - na + 1) * RECB;
        const int g = g0 + k;
        ghost_origin[g - nlocal] = reinterpret_cast<const int2 *>(rec)[8];
        const double xx = rec[0], yy = rec[1], zz = rec[2], vx = rec[3], vy = rec[4], vz = rec[5];
        const int2 tt = reinterpret_cast<const int2 *>(rec)[6], ms = reinterpret_cast<const int2 *>(rec)[7];
        x.c[0][g] = xx; x.c[1][g] = yy; x.c[2][g] = zz;
        v.c[0][g] = vx; v.c[1][g] = vy; v.c[2][g] = vz;
        tag[g] = tt.x; type[g] = tt.y; mask[g] = ms.x;
        float4 c, w;
        c.x = (float)(xx - box.centre[0]); c.y = (float)(yy - box.centre[1]); c.z = (float)(zz - box.centre[2]);
        c.w = __int_as_float(tt.y - 1);
        w.x = (float)vx; w.y = (float)vy; w.z = (float)vz; w.w = __int_as_float(ms.y);
        coord4[g] = c; veloc4[g] = w;
    }
}

// ------------------------------------------------------------------ per-step forward (pack_comm_vel / unpack_comm_vel)
// Forward records are 40 bytes: {x + shift (3 x fp64), veloc4 = fp32 v + this step's signature}.  The force kernel reads
// ghosts only through the packed views, and a ghost's packed velocity is what later dimensions forward, so the fp64 ghost
// velocity is simply the widened fp32 value.  Messages carry exactly the records of the send lists built at the last
// rebuild (sizes known to both sides from that rebuild's counts): no headers, no padding.
constexpr int RECF = 5;

__global__ void __launch_bounds__(256) k_mr_forward_pack(SoA3 x, const float4 *__restrict__ veloc4, const Counts *__restrict__ cnt,
                                                         const int *__restrict__ list_lo, const int *__restrict__ list_hi,
                                                         double *__restrict__ send_lo, double *__restrict__ send_hi, Box box, int d)
{
    const int na = cnt->send_n[2 * d], n = na + cnt->send_n[2 * d + 1];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int side = k >= na;
        const int kk = side ? k - na : k;
        const int i = (side ? list_hi : list_lo)[kk];
        const int pbc = box.pbc[2 * d + side];
        double *rec = (side ? send_hi : send_lo) + (size_t)kk * RECF;
        double xg[3] = {x.c[0][i], x.c[1][i], x.c[2][i]};
        if (pbc) xg[d] = pbc > 0 ? xg[d] + box.prd[d] : xg[d] - box.prd[d];
        rec[0] = xg[0]; rec[1] = xg[1]; rec[2] = xg[2];
        const float4 w = veloc4[i];
        reinterpret_cast<float2 *>(rec)[3] = make_float2(w.x, w.y);
        reinterpret_cast<float2 *>(rec)[4] = make_float2(w.z, w.w);
    }
}

__global__ void __launch_bounds__(256) k_mr_forward_unpack(SoA3 x, SoA3 v, const int *__restrict__ type, float4 *__restrict__ coord4,
                                                           float4 *__restrict__ veloc4, Counts *__restrict__ cnt,
                                                           const double *__restrict__ recv_a, const double *__restrict__ recv_b, Box box, int d)
{
    const int na = cnt->swap_n[2 * d], n = na + cnt->swap_n[2 * d + 1], g0 = cnt->swap_first[2 * d];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double *rec = (k < na) ? recv_a + (size_t)k * RECF : recv_b + (size_t)(k - na) * RECF;
        const int g = g0 + k;
        const double xx = rec[0], yy = rec[1], zz = rec[2];
        const float2 w01 = reinterpret_cast<const float2 *>(rec)[3], w23 = reinterpret_cast<const float2 *>(rec)[4];
        x.c[0][g] = xx; x.c[1][g] = yy; x.c[2][g] = zz;
        v.c[0][g] = (double)w01.x; v.c[1][g] = (double)w01.y; v.c[2][g] = (double)w23.x;
        float4 c;
        c.x = (float)(xx - box.centre[0]); c.y = (float)(yy - box.centre[1]); c.z = (float)(zz - box.centre[2]);
        c.w = __int_as_float(type[g] - 1);
        coord4[g] = c; veloc4[g] = make_float4(w01.x, w01.y, w23.x, w23.y);
    }
}

// ------------------------------------------------------------------ migration (exchange)
// leaving test `x >= hi || x < lo`, direction by the minimum image of x - mid (UM/comm_meso.cu:303-330)
__device__ __forceinline__ Flags leave_flags(double xd, const Box &box, int d)
{
    Flags f{false, false};
    if (xd >= box.subhi[d] || xd < box.sublo[d]) {
        double dist = xd - 0.5 * (box.sublo[d] + box.subhi[d]);
        if (box.periodic[d] && fabs(dist) > 0.5 * box.prd[d]) dist += dist < 0.0 ? box.prd[d] : -box.prd[d];
        f.a = dist < 0;
        f.b = !f.a;
    }
    return f;
}

__global__ void __launch_bounds__(CT) k_mr_exch_count(const double *__restrict__ xd, const Counts *__restrict__ cnt,
                                                      int2 *__restrict__ tile_counts, Box box, int d, int ntiles)
{
    tile_count([&](int i) { return leave_flags(xd[i], box, d); }, 0, cnt->nlocal, tile_counts, ntiles);
}

__global__ void __launch_bounds__(1024) k_mr_exch_scan(int2 *__restrict__ tile_counts, Counts *__restrict__ cnt, double *__restrict__ send_l,
                                                       double *__restrict__ send_r, int exch_cap)
{
    const int used = (cnt->nlocal + CTILE - 1) / CTILE;
    int2 tot = scan_tiles(tile_counts, used);
    if (threadIdx.x == 0) {
        if (tot.x > exch_cap || tot.y > exch_cap) cnt->err |= 8;      // would lose atoms
        cnt->exch_n[0] = min(tot.x, exch_cap); cnt->exch_n[1] = min(tot.y, exch_cap);
        reinterpret_cast<int *>(send_l)[0] = cnt->exch_n[0];
        reinterpret_cast<int *>(send_r)[0] = cnt->exch_n[1];
    }
}

// stayers are compacted in order into the alternate arrays; leavers become records {x(3), v(3), tag|type, mask|image}
__global__ void __launch_bounds__(CT) k_mr_exch_scatter(SoA3 x, SoA3 v, const int *__restrict__ tag, const int *__restrict__ type,
                                                        const int *__restrict__ mask, const int *__restrict__ image, SoA3 xo, SoA3 vo,
                                                        int *__restrict__ tago, int *__restrict__ typeo, int *__restrict__ masko,
                                                        int *__restrict__ imageo, const Counts *__restrict__ cnt,
                                                        const int2 *__restrict__ tile_counts, double *__restrict__ send_l,
                                                        double *__restrict__ send_r, Box box, int d, int ntiles, int exch_cap,
                                                        int *__restrict__ dest)
{
    __shared__ int2 wsum[CT / 32];
    const int last = cnt->nlocal;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int base = tile * CTILE;
        if (base >= last) break;
        int2 run = tile_counts[tile];
#pragma unroll
        for (int r = 0; r < CI; r++) {
            const int i = base + r * CT + threadIdx.x;
            Flags f{false, false};
            if (i < last) f = leave_flags(x.c[d][i], box, d);
            const uint32_t ba = __ballot_sync(0xffffffffu, f.a), bb = __ballot_sync(0xffffffffu, f.b);
            if (lane == 0) wsum[w] = make_int2(__popc(ba), __popc(bb));
            __syncthreads();
            int2 pre = make_int2(0, 0), tot = make_int2(0, 0);
#pragma unroll
            for (int ww = 0; ww < CT / 32; ww++) {
                const int2 c = wsum[ww];
                if (ww < w) { pre.x += c.x; pre.y += c.y; }
                tot.x += c.x; tot.y += c.y;
            }
            __syncthreads();
            if (i < last) {
                const int ka = run.x + pre.x + __popc(ba & lt), kb = run.y + pre.y + __popc(bb & lt);
                if (!f.a && !f.b) {
                    const int p = i - ka - kb;                  // stable: stayers keep their relative order
#pragma unroll
                    for (int q = 0; q < 3; q++) { xo.c[q][p] = x.c[q][i]; vo.c[q][p] = v.c[q][i]; }
                    tago[p] = tag[i]; typeo[p] = type[i]; masko[p] = mask[i]; imageo[p] = image[i];
                    if (dest) dest[i] = p;
                } else {
                    const int k = f.a ? ka : kb;
                    if (dest) dest[i] = k < exch_cap ? -(2 * k + (f.a ? 0 : 1)) - 1 : INT_MIN;   // record k of the left / right message
                    if (k < exch_cap) {
                        double *rec = (f.a ? send_l : send_r) + (size_t)(k + 1) * REC;
#pragma unroll
                        for (int q = 0; q < 3; q++) { rec[q] = x.c[q][i]; rec[3 + q] = v.c[q][i]; }
                        reinterpret_cast<int2 *>(rec)[6] = make_int2(tag[i], type[i]);
                        reinterpret_cast<int2 *>(rec)[7] = make_int2(mask[i], image[i]);
                    }
                }
            }
            run.x += tot.x; run.y += tot.y;
        }
    }
}

__global__ void k_mr_exch_shrink(Counts *cnt) { cnt->nlocal -= cnt->exch_n[0] + cnt->exch_n[1]; }

// arrivals are appended: first the upper neighbor's left-movers, then the lower neighbor's right-movers
// (UM/comm_meso.cu:385-417); an arrival outside [lo,hi) in this dimension is dropped there too ("rejected")
__global__ void __launch_bounds__(256) k_mr_exch_unpack(SoA3 x, SoA3 v, int *__restrict__ tag, int *__restrict__ type, int *__restrict__ mask,
                                                        int *__restrict__ image, Counts *__restrict__ cnt, const double *__restrict__ recv_a,
                                                        const double *__restrict__ recv_b, Box box, int d, int nloc_cap)
{
    const int na = reinterpret_cast<const int *>(recv_a)[0], n = na + reinterpret_cast<const int *>(recv_b)[0];
    const int base = cnt->nlocal;
    if (base + n > nloc_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&cnt->err, 8); return; }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double *rec = (k < na) ? recv_a + (size_t)(k + 1) * REC : recv_b + (size_t)(k - na + 1) * REC;
        const int p = base + k;
#pragma unroll
        for (int q = 0; q < 3; q++) { x.c[q][p] = rec[q]; v.c[q][p] = rec[3 + q]; }
        const int2 tt = reinterpret_cast<const int2 *>(rec)[6], mi = reinterpret_cast<const int2 *>(rec)[7];
        tag[p] = tt.x; type[p] = tt.y; mask[p] = mi.x; image[p] = mi.y;
        if (!(rec[d] >= box.sublo[d] && rec[d] < box.subhi[d])) atomicOr(&cnt->err, 8);
    }
}

__global__ void k_mr_exch_grow(Counts *cnt, const double *recv_a, const double *recv_b)
{
    cnt->nlocal += reinterpret_cast<const int *>(recv_a)[0] + reinterpret_cast<const int *>(recv_b)[0];
    cnt->nall = cnt->nlocal;
}

// bead-spring topology rides the migration (AtomVecDPDBond::pack_exchange / unpack_exchange, UM/atom_vec_dpd_bond_meso.cu):
// a second message per direction, records of (1 + bond_per_atom) int2 = {nbond, 0}, {partner tag, bond type}..., in the
// order of the atom records; stayers are compacted with the destination map the atom scatter wrote
__global__ void __launch_bounds__(256) k_mr_exch_bonds_scatter(const int *__restrict__ dest, const int *__restrict__ nbond,
                                                               const int2 *__restrict__ bonds, int *__restrict__ nbond_o,
                                                               int2 *__restrict__ bonds_o, int2 *__restrict__ send_l, int2 *__restrict__ send_r,
                                                               const Counts *__restrict__ cnt, size_t padding, int bpa)
{
    const int n = cnt->nlocal;                                   // still the count before the leavers were removed
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int dd = dest[i], nb = nbond[i];
        if (dd >= 0) {
            nbond_o[dd] = nb;
            for (int q = 0; q < nb; q++) bonds_o[dd + q * padding] = bonds[i + q * padding];
        } else if (dd != INT_MIN) {
            const int sidx = -(dd + 1), k = sidx >> 1;
            int2 *rec = ((sidx & 1) ? send_r : send_l) + (size_t)k * (1 + bpa);
            rec[0] = make_int2(nb, 0);
            for (int q = 0; q < nb; q++) rec[1 + q] = bonds[i + q * padding];
        }
    }
}

__global__ void __launch_bounds__(256) k_mr_exch_bonds_unpack(int *__restrict__ nbond, int2 *__restrict__ bonds, const Counts *__restrict__ cnt,
                                                              const double *__restrict__ hdr_a, const double *__restrict__ hdr_b,
                                                              const int2 *__restrict__ recv_a, const int2 *__restrict__ recv_b, size_t padding,
                                                              int bpa, int nloc_cap)
{
    const int na = reinterpret_cast<const int *>(hdr_a)[0], n = na + reinterpret_cast<const int *>(hdr_b)[0];
    const int base = cnt->nlocal;                                // arrivals are appended; k_mr_exch_grow runs after this kernel
    if (base + n > nloc_cap) return;                             // flagged by k_mr_exch_unpack
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int2 *rec = (k < na) ? recv_a + (size_t)k * (1 + bpa) : recv_b + (size_t)(k - na) * (1 + bpa);
        const int p = base + k, nb = min(max(rec[0].x, 0), bpa);
        nbond[p] = nb;
        for (int q = 0; q < nb; q++) bonds[p + q * padding] = rec[1 + q];
    }
}

// ------------------------------------------------------------------ host drivers
static int ensure_comm_buffers(meso_ctx *ctx)
{
    // slab of width cutghost over the largest (ghost-extended) face, at the brick's density, with head room
    const Box &b = ctx->box;
    double w[3], vol = 1.0;
    for (int d = 0; d < 3; d++) { w[d] = b.subhi[d] - b.sublo[d]; vol *= w[d]; }
    const double dens = std::max(1.0, (double)ctx->nlocal_host / vol);
    double face = 0;
    for (int d = 0; d < 3; d++) {
        double a = 1.0;
        for (int q = 0; q < 3; q++) if (q != d) a *= w[q] + 2.0 * ctx->cutneighmax;
        face = std::max(face, a);
    }
    // ghosts of one swap: slab of width cutghost over the ghost-extended face (+20 %, density fluctuations are ~1 % at this size);
    // leavers per dimension and rebuild: a layer |v| * every * dt thick, ~0.1 % of the brick -- 1/64 is ample; overflow sets err
    int swap_cap = (int)(face * ctx->cutneighmax * dens * 1.2) + 4096;
    int exch_cap = std::max(8192, ctx->nlocal_host / 64);
    if (ctx->nranks > 1 && !ctx->comm_caps_agreed) {
        // message sizes are part of the protocol: every rank must use the same capacities (ranks own slightly different
        // atom counts, so the local estimates can differ).  One max-reduction at the first rebuild after an upload,
        // which every rank reaches at the same point.
        double caps[2] = {(double)swap_cap, (double)exch_cap};
        if (!ctx->reduce_buf.reserve(64)) { ctx->err = "out of device memory"; return MESO_ECUDA; }
        MESO_CUDA(cudaMemcpyAsync(ctx->reduce_buf.p, caps, sizeof caps, cudaMemcpyHostToDevice, ctx->stream));
        MESO_NCCL(ncclAllReduce(ctx->reduce_buf.p, ctx->reduce_buf.p, 2, ncclDouble, ncclMax, (ncclComm_t)ctx->nccl, ctx->stream));
        MESO_CUDA(cudaMemcpyAsync(caps, ctx->reduce_buf.p, sizeof caps, cudaMemcpyDeviceToHost, ctx->stream));
        MESO_CUDA(cudaStreamSynchronize(ctx->stream));
        swap_cap = (int)caps[0]; exch_cap = (int)caps[1];
        ctx->swap_cap = 0; ctx->exch_cap = 0;                 // adopt the agreed sizes exactly
        ctx->comm_caps_agreed = true;
    } else if (ctx->nranks > 1) return MESO_OK;               // agreed sizes stay until the next upload
    if (swap_cap <= ctx->swap_cap && exch_cap <= ctx->exch_cap) return MESO_OK;
    ctx->swap_cap = std::max(ctx->swap_cap, swap_cap);
    ctx->exch_cap = std::max(ctx->exch_cap, exch_cap);
    const size_t msg = (size_t)(std::max(ctx->swap_cap, ctx->exch_cap) + 1) * RECB;
    bool ok = true;
    for (int s = 0; s < 2; s++) ok = ok && ctx->send_buf[s].reserve(msg) && ctx->recv_buf[s].reserve(msg);
    for (int s = 0; s < 6; s++) ok = ok && ctx->sendlist[s].reserve((size_t)ctx->swap_cap);
    const size_t ntiles = (ctx->cap + CTILE - 1) / CTILE;
    ok = ok && ctx->tile_counts.reserve(ntiles * 2 + 16) && ctx->ghost_origin.reserve(ctx->cap);
    if (!ok) { ctx->err = "out of device memory (halo buffers)"; return MESO_ECUDA; }
    return MESO_OK;
}

// one dimension's pair of messages: mine to lower/upper neighbor, theirs from upper/lower.  Self-partnered
// dimensions alias the receive pointers to the send buffers (no copy, no NCCL).
static int swap_messages4(meso_ctx *ctx, int d, size_t send_lo, size_t send_hi, size_t recv_up, size_t recv_lo, cudaStream_t st,
                          const double *&recv_a, const double *&recv_b)
{
    if (ctx->procgrid[d] == 1) { recv_a = ctx->send_buf[0].p; recv_b = ctx->send_buf[1].p; return MESO_OK; }
    ncclComm_t comm = (ncclComm_t)ctx->nccl;
    const int lower = ctx->procneigh[d][0], upper = ctx->procneigh[d][1];
    MESO_NCCL(ncclGroupStart());
    if (send_lo) MESO_NCCL(ncclSend(ctx->send_buf[0].p, send_lo, ncclDouble, lower, comm, st));
    if (recv_up) MESO_NCCL(ncclRecv(ctx->recv_buf[0].p, recv_up, ncclDouble, upper, comm, st));
    if (send_hi) MESO_NCCL(ncclSend(ctx->send_buf[1].p, send_hi, ncclDouble, upper, comm, st));
    if (recv_lo) MESO_NCCL(ncclRecv(ctx->recv_buf[1].p, recv_lo, ncclDouble, lower, comm, st));
    MESO_NCCL(ncclGroupEnd());
    recv_a = ctx->recv_buf[0].p; recv_b = ctx->recv_buf[1].p;
    return MESO_OK;
}
static int swap_messages(meso_ctx *ctx, int d, size_t ndoubles, cudaStream_t st, const double *&recv_a, const double *&recv_b)
{
    return swap_messages4(ctx, d, ndoubles, ndoubles, ndoubles, ndoubles, st, recv_a, recv_b);
}

int launch_exchange_multi(meso_ctx *ctx)
{
    int rc = ensure_comm_buffers(ctx);
    if (rc) return rc;
    const Box &box = ctx->box;
    const int ntiles = (int)((ctx->cap + CTILE - 1) / CTILE);
    int2 *tc = reinterpret_cast<int2 *>(ctx->tile_counts.p);
    cudaStream_t st = ctx->stream;
    const bool bonded = bonds_active(ctx);
    const int bpa = ctx->bond_per_atom;
    const size_t bond_msg = (size_t)ctx->exch_cap * (size_t)(1 + bpa);      // int2 records per bond message
    if (bonded) {
        bool ok = ctx->exch_dest.reserve(ctx->cap);
        for (int q = 0; q < 2; q++) ok = ok && ctx->bond_send[q].reserve(bond_msg) && ctx->bond_recv[q].reserve(bond_msg);
        if (!ok) { ctx->err = "out of device memory (bond migration buffers)"; return MESO_ECUDA; }
    }
    for (int d = 0; d < 3; d++) {
        if (ctx->procgrid[d] == 1) continue;        // the periodic wrap already put every atom back into the brick
        k_mr_exch_count<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(ctx->x[d].p, ctx->d_counts, tc, box, d, ntiles);
        k_mr_exch_scan<<<1, 1024, 0, LS(st)>>>(tc, ctx->d_counts, ctx->send_buf[0].p, ctx->send_buf[1].p, ctx->exch_cap);
        k_mr_exch_scatter<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->image.p,
                                                         soa(ctx->xa), soa(ctx->va), ctx->taga.p, ctx->typea.p, ctx->maska.p, ctx->imagea.p,
                                                         ctx->d_counts, tc, ctx->send_buf[0].p, ctx->send_buf[1].p, box, d, ntiles, ctx->exch_cap,
                                                         bonded ? ctx->exch_dest.p : nullptr);
        if (bonded) {
            k_mr_exch_bonds_scatter<<<grid_for(ctx, 4), 256, 0, LS(st)>>>(ctx->exch_dest.p, ctx->nbond.p, ctx->bonds.p, ctx->nbond_alt.p, ctx->bonds_alt.p,
                                                                   ctx->bond_send[0].p, ctx->bond_send[1].p, ctx->d_counts, ctx->cap, bpa);
            std::swap(ctx->nbond.p, ctx->nbond_alt.p); std::swap(ctx->nbond.cap, ctx->nbond_alt.cap);
            std::swap(ctx->bonds.p, ctx->bonds_alt.p); std::swap(ctx->bonds.cap, ctx->bonds_alt.cap);
        }
        for (int q = 0; q < 3; q++) { std::swap(ctx->x[q].p, ctx->xa[q].p); std::swap(ctx->v[q].p, ctx->va[q].p); }
        std::swap(ctx->tag.p, ctx->taga.p); std::swap(ctx->type.p, ctx->typea.p);
        std::swap(ctx->mask.p, ctx->maska.p); std::swap(ctx->image.p, ctx->imagea.p);
        k_mr_exch_shrink<<<1, 1, 0, LS(st)>>>(ctx->d_counts);
        const double *ra, *rb;
        rc = swap_messages(ctx, d, (size_t)(ctx->exch_cap + 1) * REC, st, ra, rb);
        if (rc) return rc;
        k_mr_exch_unpack<<<grid_for(ctx, 2), 256, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->image.p,
                                                        ctx->d_counts, ra, rb, box, d, (int)ctx->nloc_cap);
        if (bonded) {
            ncclComm_t comm = (ncclComm_t)ctx->nccl;
            const size_t nint = bond_msg * 2;
            MESO_NCCL(ncclGroupStart());
            MESO_NCCL(ncclSend(ctx->bond_send[0].p, nint, ncclInt, ctx->procneigh[d][0], comm, st));
            MESO_NCCL(ncclRecv(ctx->bond_recv[0].p, nint, ncclInt, ctx->procneigh[d][1], comm, st));
            MESO_NCCL(ncclSend(ctx->bond_send[1].p, nint, ncclInt, ctx->procneigh[d][1], comm, st));
            MESO_NCCL(ncclRecv(ctx->bond_recv[1].p, nint, ncclInt, ctx->procneigh[d][0], comm, st));
            MESO_NCCL(ncclGroupEnd());
            k_mr_exch_bonds_unpack<<<grid_for(ctx, 2), 256, 0, LS(st)>>>(ctx->nbond.p, ctx->bonds.p, ctx->d_counts, ra, rb, ctx->bond_recv[0].p,
                                                                  ctx->bond_recv[1].p, ctx->cap, bpa, (int)ctx->nloc_cap);
        }
        k_mr_exch_grow<<<1, 1, 0, LS(st)>>>(ctx->d_counts, ra, rb);
    }
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// ------------------------------------------------------------------ direct halo routes
// The 3-phase creation above fixes WHICH images a rank holds and in WHAT order; it also forwards ghosts of earlier
// dimensions, which makes the per-step refresh of the reference three dependent messages.  Every ghost record carries its
// owner, so after the creation each rank groups its ghosts by owner rank, sends every owner the list {index, shift} it wants
// refreshed (one message per peer, once per rebuild), and from then on a step's refresh is ONE pack kernel, ONE NCCL group
// (a message per peer, exact size) and ONE unpack kernel.  Values are bit-identical to the forwarded ones: each coordinate is
// shifted by at most one period, and the packed velocity/signature is copied.
struct RouteTable {
    int np, self;
    const int2 *slist[27];      // what peer s asked me to send: [0] = {count, 0}, then {index, shift code}
    const int *dst[27];         // ghost slot (0-based behind nlocal) of the k-th record peer s sends me
    double *sbuf[27];           // staging of the records for peer s
    const double *rbuf[27];     // records received from peer s (self: my own staging buffer)
};

__global__ void __launch_bounds__(256) k_route_build(const int2 *__restrict__ ghost_origin, const int *__restrict__ peer_slot, Counts *cnt,
                                                     RouteTable rt, int2 *const *req, int *const *dst, int route_cap, int nranks)
{
    const int ng = cnt->nghost;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    // whole warps iterate together; the lanes that want the same peer share ONE atomic (a handful of counters would
    // otherwise serialise ~10^5 same-address atomics)
    for (int g0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; g0 < ng; g0 += gridDim.x * blockDim.x) {
        const int g = g0 + lane;
        int s = -1;
        int2 o = make_int2(0, 0);
        if (g < ng) {
            o = ghost_origin[g];
            const int r = o.x & 0xffffff;
            s = (r >= 0 && r < nranks) ? peer_slot[r] : -1;
            if (s < 0) atomicOr(&cnt->err, 1);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, s);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader && s >= 0) base = atomicAdd(&cnt->route_recv_n[s], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (s < 0) continue;
        const int k = base + __popc(peers & lt);
        if (k >= route_cap) { atomicOr(&cnt->err, 1); continue; }
        req[s][k + 1] = make_int2(o.y, (o.x >> 24) & 63);
        dst[s][k] = g;
    }
}

__global__ void k_route_reset(Counts *cnt)
{
    if (threadIdx.x < 27) { cnt->route_recv_n[threadIdx.x] = 0; cnt->route_send_n[threadIdx.x] = 0; }
}

__global__ void k_route_headers(Counts *cnt, int2 *const *req, int np, int route_cap)
{
    const int s = threadIdx.x;
    if (s < np) {
        const int n = min(cnt->route_recv_n[s], route_cap);
        cnt->route_recv_n[s] = n;
        req[s][0] = make_int2(n, 0);
    }
}

__global__ void k_route_adopt(Counts *cnt, RouteTable rt, int route_cap)
{
    const int s = threadIdx.x;
    if (s < rt.np) {
        int n = rt.slist[s][0].x;
        if (n < 0 || n > route_cap) { atomicOr(&cnt->err, 1); n = 0; }
        cnt->route_send_n[s] = n;
    }
}

// records: {x + shift (3 x fp64), veloc4 = fp32 v + this step's signature} = 40 bytes, as in the forwarding refresh
__global__ void __launch_bounds__(256) k_route_pack(SoA3 x, const float4 *__restrict__ veloc4, const Counts *__restrict__ cnt, RouteTable rt,
                                                    Box box)
{
    const int s = blockIdx.y;
    const int n = cnt->route_send_n[s];
    const int2 *list = rt.slist[s] + 1;
    double *buf = rt.sbuf[s];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int2 e = list[k];
        const int i = e.x;
        double xg[3] = {x.c[0][i], x.c[1][i], x.c[2][i]};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int c = (e.y >> (2 * d)) & 3;
            if (c) xg[d] = c == 1 ? xg[d] + box.prd[d] : xg[d] - box.prd[d];
        }
        double *rec = buf + (size_t)k * RECF;
        rec[0] = xg[0]; rec[1] = xg[1]; rec[2] = xg[2];
        const float4 w = veloc4[i];
        reinterpret_cast<float2 *>(rec)[3] = make_float2(w.x, w.y);
        reinterpret_cast<float2 *>(rec)[4] = make_float2(w.z, w.w);
    }
}

__global__ void __launch_bounds__(256) k_route_unpack(SoA3 x, SoA3 v, const int *__restrict__ type, float4 *__restrict__ coord4,
                                                      float4 *__restrict__ veloc4, const Counts *__restrict__ cnt, RouteTable rt, Box box)
{
    const int s = blockIdx.y;
    const int n = cnt->route_recv_n[s], nlocal = cnt->nlocal;
    const int *dst = rt.dst[s];
    const double *buf = rt.rbuf[s];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double *rec = buf + (size_t)k * RECF;
        const int g = nlocal + dst[k];
        const double xx = rec[0], yy = rec[1], zz = rec[2];
        const float2 w01 = reinterpret_cast<const float2 *>(rec)[3], w23 = reinterpret_cast<const float2 *>(rec)[4];
        x.c[0][g] = xx; x.c[1][g] = yy; x.c[2][g] = zz;
        v.c[0][g] = (double)w01.x; v.c[1][g] = (double)w01.y; v.c[2][g] = (double)w23.x;
        float4 c;
        c.x = (float)(xx - box.centre[0]); c.y = (float)(yy - box.centre[1]); c.z = (float)(zz - box.centre[2]);
        c.w = __int_as_float(type[g] - 1);
        coord4[g] = c; veloc4[g] = make_float4(w01.x, w01.y, w23.x, w23.y);
    }
}

// distinct ranks among this brick's 26 neighbors, plus itself (slot peer_self)
int comm_build_peers(meso_ctx *ctx)
{
    const Box &b = ctx->box;
    ctx->npeers = 0;
    std::vector<int> slot((size_t)ctx->nranks, -1);
    auto add = [&](int r) {
        if (slot[r] >= 0) return;
        slot[r] = ctx->npeers;
        ctx->peer_rank[ctx->npeers++] = r;
    };
    add(ctx->rank);
    ctx->peer_self = 0;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int off[3] = {dx, dy, dz}, loc[3];
                bool ok = true;
                for (int d = 0; d < 3; d++) {
                    loc[d] = ctx->myloc[d] + off[d];
                    if (loc[d] < 0 || loc[d] >= ctx->procgrid[d]) {
                        if (!b.periodic[d]) { ok = false; break; }
                        loc[d] = (loc[d] + ctx->procgrid[d]) % ctx->procgrid[d];
                    }
                }
                if (ok) add((loc[0] * ctx->procgrid[1] + loc[1]) * ctx->procgrid[2] + loc[2]);
            }
    if (!ctx->peer_slot.reserve((size_t)ctx->nranks)) { ctx->err = "out of device memory"; return MESO_ECUDA; }
    MESO_CUDA(cudaMemcpyAsync(ctx->peer_slot.p, slot.data(), sizeof(int) * ctx->nranks, cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return MESO_OK;
}

static int route_table(meso_ctx *ctx, RouteTable &rt)
{
    rt.np = ctx->npeers; rt.self = ctx->peer_self;
    for (int s = 0; s < 27; s++) { rt.slist[s] = nullptr; rt.dst[s] = nullptr; rt.sbuf[s] = nullptr; rt.rbuf[s] = nullptr; }
    for (int s = 0; s < ctx->npeers; s++) {
        const bool self = s == ctx->peer_self;
        rt.slist[s] = self ? ctx->route_req[s].p : ctx->route_send_list[s].p;     // my own requests are my own send list
        rt.dst[s] = ctx->route_dst[s].p;
        rt.sbuf[s] = ctx->route_sbuf[s].p;
        rt.rbuf[s] = self ? ctx->route_sbuf[s].p : ctx->route_rbuf[s].p;
    }
    return MESO_OK;
}

// after the ghosts exist: group them by owner, exchange the request lists (once per rebuild)
static int build_routes(meso_ctx *ctx)
{
    if (ctx->npeers == 0) { int rc = comm_build_peers(ctx); if (rc) return rc; }
    const int np = ctx->npeers;
    ctx->route_cap = 6 * ctx->swap_cap;           // a peer (or this rank itself, through periodic images) can own the ghosts of all six swaps
    bool ok = true;
    for (int s = 0; s < np; s++) {
        ok = ok && ctx->route_req[s].reserve((size_t)ctx->route_cap + 1) && ctx->route_dst[s].reserve((size_t)ctx->route_cap) &&
             ctx->route_sbuf[s].reserve((size_t)ctx->route_cap * RECF);
        if (s != ctx->peer_self) ok = ok && ctx->route_send_list[s].reserve((size_t)ctx->route_cap + 1) && ctx->route_rbuf[s].reserve((size_t)ctx->route_cap * RECF);
    }
    // device-side pointer tables for the build kernel
    ok = ok && ctx->route_ptrs.reserve(64);
    if (!ok) { ctx->err = "out of device memory (halo routes)"; return MESO_ECUDA; }
    void *hp[54];
    for (int s = 0; s < 27; s++) { hp[s] = s < np ? (void *)ctx->route_req[s].p : nullptr; hp[27 + s] = s < np ? (void *)ctx->route_dst[s].p : nullptr; }
    cudaStream_t st = ctx->stream;
    MESO_CUDA(cudaMemcpyAsync(ctx->route_ptrs.p, hp, sizeof hp, cudaMemcpyHostToDevice, st));
    RouteTable rt;
    route_table(ctx, rt);
    int2 *const *req = reinterpret_cast<int2 *const *>(ctx->route_ptrs.p);
    int *const *dst = reinterpret_cast<int *const *>(ctx->route_ptrs.p + 27);
    k_route_reset<<<1, 32, 0, LS(st)>>>(ctx->d_counts);
    k_route_build<<<grid_for(ctx, 2), 256, 0, LS(st)>>>(ctx->ghost_origin.p, ctx->peer_slot.p, ctx->d_counts, rt, req, dst, ctx->route_cap, ctx->nranks);
    k_route_headers<<<1, 32, 0, LS(st)>>>(ctx->d_counts, req, np, ctx->route_cap);
    if (np > 1) {
        ncclComm_t comm = (ncclComm_t)ctx->nccl;
        const size_t nint = ((size_t)ctx->route_cap + 1) * 2;
        MESO_NCCL(ncclGroupStart());
        for (int s = 0; s < np; s++) {
            if (s == ctx->peer_self) continue;
            MESO_NCCL(ncclSend(ctx->route_req[s].p, nint, ncclInt, ctx->peer_rank[s], comm, st));
            MESO_NCCL(ncclRecv(ctx->route_send_list[s].p, nint, ncclInt, ctx->peer_rank[s], comm, st));
        }
        MESO_NCCL(ncclGroupEnd());
    }
    k_route_adopt<<<1, 32, 0, LS(st)>>>(ctx->d_counts, rt, ctx->route_cap);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// per-step ghost refresh over the routes, on stream `st`
static int forward_routes(meso_ctx *ctx, cudaStream_t st)
{
    const int np = ctx->npeers;
    RouteTable rt;
    route_table(ctx, rt);
    const dim3 grid(ctx->sm_count, np);
    k_route_pack<<<grid, 256, 0, LS(st)>>>(soa(ctx->x), ctx->veloc4.p, ctx->d_counts, rt, ctx->box);
    if (np > 1) {
        ncclComm_t comm = (ncclComm_t)ctx->nccl;
        MESO_NCCL(ncclGroupStart());
        for (int s = 0; s < np; s++) {
            if (s == ctx->peer_self) continue;
            if (ctx->route_send_n[s]) MESO_NCCL(ncclSend(ctx->route_sbuf[s].p, (size_t)ctx->route_send_n[s] * RECF, ncclDouble, ctx->peer_rank[s], comm, st));
            if (ctx->route_recv_n[s]) MESO_NCCL(ncclRecv(ctx->route_rbuf[s].p, (size_t)ctx->route_recv_n[s] * RECF, ncclDouble, ctx->peer_rank[s], comm, st));
        }
        MESO_NCCL(ncclGroupEnd());
    }
    k_route_unpack<<<grid, 256, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->type.p, ctx->coord4.p, ctx->veloc4.p, ctx->d_counts, rt, ctx->box);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// ------------------------------------------------------------------ one-shot migration (MESO_EXCH_ONESHOT=1, off by default)
// NOT YET RUN ON HARDWARE (written at the end of round 1, when the GPU budget was spent): the validated path is
// launch_exchange_multi above.  Instead of one dependent exchange per communicating dimension (an atom that leaves through
// an edge takes two hops), every leaver is sent straight to the brick that will own it: one selection kernel, ONE NCCL group
// with a fixed-capacity message per peer, one unpack kernel.  Ownership after the exchange is the same as after the three
// hops; arrival order differs, which only matters for ties of identical sort keys (already a documented deviation).
struct OneShot {
    int np, self, bpa;
    int slot_of_code[27];       // destination offset code (ox+1) + 3(oy+1) + 9(oz+1) -> peer slot, -1: no such brick (atom lost)
    int multi[3];               // procgrid[d] > 1
    double *sbuf[27];           // [0] header {count}, then REC doubles per leaver
    const double *rbuf[27];
    int2 *bsend[27];            // bond rows of the leavers, (1 + bpa) int2 per record, same order
    const int2 *brecv[27];
};

__device__ __forceinline__ int oneshot_code(const double xi[3], const Box &box, const int multi[3])
{
    int code = 0, w = 1;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        int o = 0;
        if (multi[d]) { const Flags f = leave_flags(xi[d], box, d); o = f.a ? -1 : (f.b ? 1 : 0); }
        code += (o + 1) * w;
        w *= 3;
    }
    return code;
}

__global__ void __launch_bounds__(CT) k_os_count(SoA3 x, const Counts *__restrict__ cnt, int2 *__restrict__ tile_counts, Box box, OneShot os,
                                                 int ntiles)
{
    tile_count([&](int i) {
        const double xi[3] = {x.c[0][i], x.c[1][i], x.c[2][i]};
        Flags f; f.a = oneshot_code(xi, box, os.multi) != 13; f.b = false; return f; }, 0, cnt->nlocal, tile_counts, ntiles);
}

__global__ void __launch_bounds__(1024) k_os_scan(int2 *__restrict__ tile_counts, Counts *__restrict__ cnt)
{
    const int used = (cnt->nlocal + CTILE - 1) / CTILE;
    const int2 tot = scan_tiles(tile_counts, used);
    if (threadIdx.x == 0) { cnt->exch_n[0] = tot.x; cnt->exch_n[1] = 0; }
    if (threadIdx.x < 27) cnt->route_send_n[threadIdx.x] = 0;      // reused as per-peer leaver counters until the routes are rebuilt
}

__global__ void __launch_bounds__(CT) k_os_scatter(SoA3 x, SoA3 v, const int *__restrict__ tag, const int *__restrict__ type,
                                                   const int *__restrict__ mask, const int *__restrict__ image, SoA3 xo, SoA3 vo,
                                                   int *__restrict__ tago, int *__restrict__ typeo, int *__restrict__ masko,
                                                   int *__restrict__ imageo, Counts *__restrict__ cnt, const int2 *__restrict__ tile_counts,
                                                   Box box, OneShot os, int ntiles, int exch_cap, const int *__restrict__ nbond,
                                                   const int2 *__restrict__ bonds, int *__restrict__ nbond_o, int2 *__restrict__ bonds_o,
                                                   size_t padding)
{
    __shared__ int wsum[CT / 32];
    const int last = cnt->nlocal;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int base = tile * CTILE;
        if (base >= last) break;
        int run = tile_counts[tile].x;
#pragma unroll
        for (int r = 0; r < CI; r++) {
            const int i = base + r * CT + threadIdx.x;
            int code = 13;
            double xi[3] = {0, 0, 0};
            if (i < last) {
#pragma unroll
                for (int q = 0; q < 3; q++) xi[q] = x.c[q][i];
                code = oneshot_code(xi, box, os.multi);
            }
            const uint32_t ba = __ballot_sync(0xffffffffu, code != 13);
            if (lane == 0) wsum[w] = __popc(ba);
            __syncthreads();
            int pre = 0, tot = 0;
#pragma unroll
            for (int ww = 0; ww < CT / 32; ww++) { const int c = wsum[ww]; if (ww < w) pre += c; tot += c; }
            __syncthreads();
            if (i < last) {
                if (code == 13) {
                    const int p = i - (run + pre + __popc(ba & lt));    // stable: stayers keep their relative order
#pragma unroll
                    for (int q = 0; q < 3; q++) { xo.c[q][p] = xi[q]; vo.c[q][p] = v.c[q][i]; }
                    tago[p] = tag[i]; typeo[p] = type[i]; masko[p] = mask[i]; imageo[p] = image[i];
                    if (os.bpa) {
                        const int nb = nbond[i];
                        nbond_o[p] = nb;
                        for (int q = 0; q < nb; q++) bonds_o[p + q * padding] = bonds[i + q * padding];
                    }
                } else {
                    const int s = os.slot_of_code[code];
                    if (s < 0) atomicOr(&cnt->err, 8);               // beyond a non-periodic face of the decomposition: lost
                    else {
                        const int k = atomicAdd(&cnt->route_send_n[s], 1);
                        if (k >= exch_cap) atomicOr(&cnt->err, 8);
                        else {
                            double *rec = os.sbuf[s] + (size_t)(k + 1) * REC;
#pragma unroll
                            for (int q = 0; q < 3; q++) { rec[q] = xi[q]; rec[3 + q] = v.c[q][i]; }
                            reinterpret_cast<int2 *>(rec)[6] = make_int2(tag[i], type[i]);
                            reinterpret_cast<int2 *>(rec)[7] = make_int2(mask[i], image[i]);
                            if (os.bpa) {
                                int2 *br = os.bsend[s] + (size_t)k * (1 + os.bpa);
                                const int nb = nbond[i];
                                br[0] = make_int2(nb, 0);
                                for (int q = 0; q < nb; q++) br[1 + q] = bonds[i + q * padding];
                            }
                        }
                    }
                }
            }
            run += tot;
        }
    }
}

__global__ void k_os_headers(Counts *cnt, OneShot os, int exch_cap)
{
    const int s = threadIdx.x;
    if (s < os.np && s != os.self) reinterpret_cast<int *>(os.sbuf[s])[0] = min(cnt->route_send_n[s], exch_cap);
    if (s == 0) cnt->nlocal -= cnt->exch_n[0];
}

// arrivals are appended peer by peer in slot order
__global__ void __launch_bounds__(256) k_os_unpack(SoA3 x, SoA3 v, int *__restrict__ tag, int *__restrict__ type, int *__restrict__ mask,
                                                   int *__restrict__ image, Counts *__restrict__ cnt, OneShot os, Box box, int nloc_cap,
                                                   int *__restrict__ nbond, int2 *__restrict__ bonds, size_t padding)
{
    const int s = blockIdx.y;
    if (s == os.self) return;
    int base = cnt->nlocal, total = 0;
    for (int q = 0; q < os.np; q++) {
        if (q == os.self) continue;
        const int nq = reinterpret_cast<const int *>(os.rbuf[q])[0];
        if (q < s) base += nq;
        total += nq;
    }
    if (cnt->nlocal + total > nloc_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&cnt->err, 8); return; }
    const int n = reinterpret_cast<const int *>(os.rbuf[s])[0];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double *rec = os.rbuf[s] + (size_t)(k + 1) * REC;
        const int p = base + k;
        bool inside = true;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            x.c[q][p] = rec[q]; v.c[q][p] = rec[3 + q];
            if (os.multi[q]) inside = inside && rec[q] >= box.sublo[q] && rec[q] < box.subhi[q];
        }
        const int2 tt = reinterpret_cast<const int2 *>(rec)[6], mi = reinterpret_cast<const int2 *>(rec)[7];
        tag[p] = tt.x; type[p] = tt.y; mask[p] = mi.x; image[p] = mi.y;
        if (!inside) atomicOr(&cnt->err, 8);
        if (os.bpa) {
            const int2 *br = os.brecv[s] + (size_t)k * (1 + os.bpa);
            const int nb = min(max(br[0].x, 0), os.bpa);
            nbond[p] = nb;
            for (int q = 0; q < nb; q++) bonds[p + q * padding] = br[1 + q];
        }
    }
}

__global__ void k_os_grow(Counts *cnt, OneShot os)
{
    int total = 0;
    for (int q = 0; q < os.np; q++) if (q != os.self) total += reinterpret_cast<const int *>(os.rbuf[q])[0];
    cnt->nlocal += total;
    cnt->nall = cnt->nlocal;
}

int launch_exchange_oneshot(meso_ctx *ctx)
{
    int rc = ensure_comm_buffers(ctx);
    if (rc) return rc;
    if (ctx->npeers == 0) { rc = comm_build_peers(ctx); if (rc) return rc; }
    const Box &box = ctx->box;
    const int ntiles = (int)((ctx->cap + CTILE - 1) / CTILE);
    int2 *tc = reinterpret_cast<int2 *>(ctx->tile_counts.p);
    cudaStream_t st = ctx->stream;
    const bool bonded = bonds_active(ctx);
    OneShot os;
    os.np = ctx->npeers; os.self = ctx->peer_self; os.bpa = bonded ? ctx->bond_per_atom : 0;
    bool any = false;
    for (int d = 0; d < 3; d++) { os.multi[d] = ctx->procgrid[d] > 1; any = any || os.multi[d]; }
    if (!any) return MESO_OK;                     // the periodic wrap already put every atom back into the brick
    for (int c = 0; c < 27; c++) {
        const int off[3] = {c % 3 - 1, (c / 3) % 3 - 1, c / 9 - 1};
        int loc[3], slot = -1;
        bool ok = true;
        for (int d = 0; d < 3; d++) {
            loc[d] = ctx->myloc[d] + off[d];
            if (loc[d] < 0 || loc[d] >= ctx->procgrid[d]) {
                if (!box.periodic[d]) { ok = false; break; }
                loc[d] = (loc[d] + ctx->procgrid[d]) % ctx->procgrid[d];
            }
        }
        if (ok) {
            const int r = (loc[0] * ctx->procgrid[1] + loc[1]) * ctx->procgrid[2] + loc[2];
            for (int s = 0; s < ctx->npeers; s++) if (ctx->peer_rank[s] == r) slot = s;
        }
        os.slot_of_code[c] = slot;
    }
    const size_t msg = (size_t)(ctx->exch_cap + 1) * REC, bmsg = (size_t)ctx->exch_cap * (size_t)(1 + os.bpa);
    bool okm = true;
    for (int s = 0; s < 27; s++) { os.sbuf[s] = nullptr; os.rbuf[s] = nullptr; os.bsend[s] = nullptr; os.brecv[s] = nullptr; }
    for (int s = 0; s < os.np; s++) {
        if (s == os.self) continue;
        okm = okm && ctx->os_sbuf[s].reserve(msg) && ctx->os_rbuf[s].reserve(msg);
        if (bonded) okm = okm && ctx->os_bsend[s].reserve(bmsg) && ctx->os_brecv[s].reserve(bmsg);
        os.sbuf[s] = ctx->os_sbuf[s].p; os.rbuf[s] = ctx->os_rbuf[s].p;
        os.bsend[s] = ctx->os_bsend[s].p; os.brecv[s] = ctx->os_brecv[s].p;
    }
    if (!okm) { ctx->err = "out of device memory (one-shot migration buffers)"; return MESO_ECUDA; }
    k_os_count<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(soa(ctx->x), ctx->d_counts, tc, box, os, ntiles);
    k_os_scan<<<1, 1024, 0, LS(st)>>>(tc, ctx->d_counts);
    k_os_scatter<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->image.p, soa(ctx->xa),
                                                soa(ctx->va), ctx->taga.p, ctx->typea.p, ctx->maska.p, ctx->imagea.p, ctx->d_counts, tc, box, os,
                                                ntiles, ctx->exch_cap, ctx->nbond.p, ctx->bonds.p, ctx->nbond_alt.p, ctx->bonds_alt.p, ctx->cap);
    for (int q = 0; q < 3; q++) { std::swap(ctx->x[q].p, ctx->xa[q].p); std::swap(ctx->v[q].p, ctx->va[q].p); }
    std::swap(ctx->tag.p, ctx->taga.p); std::swap(ctx->type.p, ctx->typea.p);
    std::swap(ctx->mask.p, ctx->maska.p); std::swap(ctx->image.p, ctx->imagea.p);
    if (bonded) {
        std::swap(ctx->nbond.p, ctx->nbond_alt.p); std::swap(ctx->nbond.cap, ctx->nbond_alt.cap);
        std::swap(ctx->bonds.p, ctx->bonds_alt.p); std::swap(ctx->bonds.cap, ctx->bonds_alt.cap);
    }
    k_os_headers<<<1, 32, 0, LS(st)>>>(ctx->d_counts, os, ctx->exch_cap);
    ncclComm_t comm = (ncclComm_t)ctx->nccl;
    MESO_NCCL(ncclGroupStart());
    for (int s = 0; s < os.np; s++) {
        if (s == os.self) continue;
        MESO_NCCL(ncclSend(ctx->os_sbuf[s].p, msg, ncclDouble, ctx->peer_rank[s], comm, st));
        MESO_NCCL(ncclRecv(ctx->os_rbuf[s].p, msg, ncclDouble, ctx->peer_rank[s], comm, st));
        if (bonded) {
            MESO_NCCL(ncclSend(ctx->os_bsend[s].p, bmsg * 2, ncclInt, ctx->peer_rank[s], comm, st));
            MESO_NCCL(ncclRecv(ctx->os_brecv[s].p, bmsg * 2, ncclInt, ctx->peer_rank[s], comm, st));
        }
    }
    MESO_NCCL(ncclGroupEnd());
    const dim3 grid(ctx->sm_count, os.np);
    k_os_unpack<<<grid, 256, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->image.p, ctx->d_counts, os, box,
                                    (int)ctx->nloc_cap, ctx->nbond.p, ctx->bonds.p, ctx->cap);
    k_os_grow<<<1, 1, 0, LS(st)>>>(ctx->d_counts, os);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

__global__ void k_mr_reset_ghosts(Counts *cnt)
{
    cnt->nghost = 0;
    cnt->nall = cnt->nlocal;
    cnt->max_pair = 0;
    for (int s = 0; s < 6; s++) { cnt->swap_first[s] = cnt->nlocal; cnt->swap_n[s] = 0; cnt->send_n[s] = 0; }
}

int launch_borders_multi(meso_ctx *ctx)
{
    int rc = ensure_comm_buffers(ctx);
    if (rc) return rc;
    if (!ctx->ghost_origin.reserve(ctx->cap)) { ctx->err = "out of device memory (ghost owners)"; return MESO_ECUDA; }
    const Box &box = ctx->box;
    const int ntiles = (int)((ctx->cap + CTILE - 1) / CTILE);
    int2 *tc = reinterpret_cast<int2 *>(ctx->tile_counts.p);
    cudaStream_t st = ctx->stream;
    k_mr_reset_ghosts<<<1, 1, 0, LS(st)>>>(ctx->d_counts);
    for (int d = 0; d < 3; d++) {
        if (!box.sendflag[2 * d] && !box.sendflag[2 * d + 1] && ctx->procgrid[d] == 1) continue;
        k_mr_border_count<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(ctx->x[d].p, ctx->d_counts, tc, box, d, ntiles);
        k_mr_border_scan<<<1, 1024, 0, LS(st)>>>(tc, ctx->d_counts, ctx->send_buf[0].p, ctx->send_buf[1].p, d, ctx->swap_cap);
        k_mr_border_pack<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->veloc4.p,
                                                        ctx->d_counts, tc, ctx->send_buf[0].p, ctx->send_buf[1].p, ctx->sendlist[2 * d].p,
                                                        ctx->sendlist[2 * d + 1].p, box, d, ntiles, ctx->swap_cap, ctx->ghost_origin.p, ctx->rank);
        const double *ra, *rb;
        rc = swap_messages(ctx, d, (size_t)(ctx->swap_cap + 1) * RECB, st, ra, rb);
        if (rc) return rc;
        k_mr_border_advance<<<1, 1, 0, LS(st)>>>(ctx->d_counts, ra, rb, d, (int)ctx->cap);
        k_mr_border_unpack<<<grid_for(ctx, 2), 256, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->coord4.p,
                                                          ctx->veloc4.p, ctx->d_counts, ra, rb, box, d, ctx->ghost_origin.p);
    }
    MESO_CUDA(cudaGetLastError());
    return ctx->halo_routes ? build_routes(ctx) : MESO_OK;
}

int comm_share_errors(meso_ctx *ctx)
{
    Counts *c = ctx->d_counts;
    MESO_CUDA(cudaMemcpyAsync(&c->err_any, &c->err, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->nranks > 1)
        MESO_NCCL(ncclAllReduce(&c->err_any, &c->err_any, 1, ncclInt, ncclMax, (ncclComm_t)ctx->nccl, ctx->stream));
    return MESO_OK;
}

// per-step ghost refresh on stream `st` (the side stream when it overlaps the bulk force kernel).  Message sizes come from
// the counts of the rebuild that built the send lists: the host waits once per rebuild for that (tiny) copy; the
// sender's send_n[s] equals the receiver's swap_n[s] by construction (it is the header the receiver unpacked).
int launch_forward_multi(meso_ctx *ctx, cudaStream_t st)
{
    const Box &box = ctx->box;
    if (!ctx->fwd_counts_valid) {
        MESO_CUDA(cudaEventSynchronize(ctx->ev_counts));
        if (ctx->h_counts->err_any) {
            // some rank overflowed a capacity during the rebuild: its ghost counts no longer match its partners' send lists.
            // Every rank sees the same flag and stops here, before a size-mismatched message could block the others.
            ctx->err = "device-side capacity error on some rank during the last rebuild (ghost / migration / pair-table capacity)";
            return MESO_ECAPACITY;
        }
        for (int s = 0; s < 6; s++) { ctx->fwd_send_n[s] = ctx->h_counts->send_n[s]; ctx->fwd_recv_n[s] = ctx->h_counts->swap_n[s]; }
        for (int s = 0; s < 27; s++) { ctx->route_send_n[s] = ctx->h_counts->route_send_n[s]; ctx->route_recv_n[s] = ctx->h_counts->route_recv_n[s]; }
        ctx->fwd_counts_valid = true;
    }
    if (ctx->halo_routes) return forward_routes(ctx, st);
    for (int d = 0; d < 3; d++) {
        if (!box.sendflag[2 * d] && !box.sendflag[2 * d + 1] && ctx->procgrid[d] == 1) continue;
        k_mr_forward_pack<<<grid_for(ctx, 2), 256, 0, LS(st)>>>(soa(ctx->x), ctx->veloc4.p, ctx->d_counts, ctx->sendlist[2 * d].p,
                                                         ctx->sendlist[2 * d + 1].p, ctx->send_buf[0].p, ctx->send_buf[1].p, box, d);
        const double *ra, *rb;
        int rc = swap_messages4(ctx, d, (size_t)ctx->fwd_send_n[2 * d] * RECF, (size_t)ctx->fwd_send_n[2 * d + 1] * RECF,
                                (size_t)ctx->fwd_recv_n[2 * d] * RECF, (size_t)ctx->fwd_recv_n[2 * d + 1] * RECF, st, ra, rb);
        if (rc) return rc;
        k_mr_forward_unpack<<<grid_for(ctx, 2), 256, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->type.p, ctx->coord4.p, ctx->veloc4.p, ctx->d_counts,
                                                           ra, rb, box, d);
    }
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

}  // namespace meso

extern "C" int meso_comm_unique_id(void *id128)
{
    if (!id128) return MESO_EINVAL;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return MESO_ENCCL;
    memcpy(id128, &id, sizeof id);
    return MESO_OK;
}
