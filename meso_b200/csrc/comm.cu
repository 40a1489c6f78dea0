// comm.cu -- spatial decomposition across GPUs: migration, ghost creation and the per-step halo refresh as
// remote stores into peer memory over NVLink / NVSwitch, synchronised by flags.  NCCL only bootstraps (it carries the
// memory handles) and reduces thermo scalars; no NCCL call sits on the data path.
//
// Reference path (host MPI through pinned staging buffers, every step):
//   MesoComm::exchange                UM/comm_meso.cu:256-420        (migration, one dependent hop per dimension)
//   MesoComm::borders                 UM/comm_meso.cu:41-186         (ghost creation, 6 dependent swaps, later
//                                                                     dimensions forward the ghosts of earlier ones)
//   Comm::forward_comm                src/comm.cpp:686-753           (ghost x,v refresh, the same 6 swaps)
//   pack/unpack_{border,comm}_vel     UM/atom_vec_dpd_atomic_meso.cu:61-243
//   MPI_Allreduce of thermo scalars   UM/compute_temp_meso.cu:97
//
// Design.  Every rank owns an ARENA in its HBM with one receive slot per direction code c = (ox+1) + 3(oy+1) + 9(oz+1):
// slot c of a rank holds what its neighbor at -o sent in direction o.  Arenas are shared between the processes of a node
// through CUDA IPC handles (or plain peer pointers inside one process), so a sender's pack kernel stores its records
// straight into the receiver's slot; its last CTA then publishes {count, error bits} and an epoch flag, and the receiver's
// stream continues behind a one-warp wait kernel.  No host round trip, no message sizing, no dependent hops:
//   * migration: every leaver goes straight to the brick that will own it (ONE exchange instead of one per dimension);
//   * ghost creation: an atom inside the send slabs of several dimensions is sent as every image it will have -- face,
//     edge and corner neighbors at once -- instead of being forwarded swap by swap.  The receiver orders the 26 slots
//     exactly as the reference's 6-swap sequence would have delivered them (slot_order below) and records inside a
//     slot keep the owner's index order, so ghost indices -- hence neighbor rows -- are bit-identical to the reference;
//   * per-step refresh: the send lists of the ghost creation are replayed (ONE pack kernel, one wait, one unpack kernel),
//     double-buffered by the parity of the epoch, on the side stream under the bulk force kernel.
// All ordered placements are multi-category stream compactions (27 categories per atom, ballots + a per-tile scan): no
// atomics decide an order, results are the same in every run.
#include "internal.h"
#include "device_math.cuh"
#include <nccl.h>
#include <unistd.h>
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

constexpr int NS = 27;            // direction codes; 13 = the rank itself (no message)
constexpr int RECM = 8;           // doubles per migration record {x(3), v(3), tag|type, mask|image}
constexpr int RECB = 8;           // doubles per ghost-creation record {x+shift(3), v(3), tag|type, mask|signature}
constexpr int RECF = 5;           // doubles per refresh record {x+shift(3), packed velocity + signature}
constexpr int CT = 256;           // threads of the compaction kernels
constexpr int CI = 4;             // candidates per thread and tile
constexpr int CTILE = CT * CI;

// first 4 KB of every arena: written by the senders, indexed by the sender's direction code
struct ArenaHdr {
    int flag_mig[32], flag_halo[2][32];   // epoch of the message in the slot (halo: ghost creation and refresh share one sequence,
                                          // buffer = epoch & 1)
    int cnt_mig[32], cnt_halo[32];        // record counts of the migration / ghost-creation message
    int err[32];                          // sender's error bits
    int ack_mig[32], ack_halo[32];        // written by the RECEIVER of what this rank sent in direction c: last epoch consumed
};
constexpr size_t HDR_BYTES = 4096;

// where my messages go: for direction c, the receiver's arena and the layout of MY slot there
struct Peers {
    unsigned char *arena[NS];
    unsigned long long mig_off[NS], bond_off[NS], gho_off[NS];
    int mcap[NS], gcap[NS];
    int bpa;
};
// my own arena: slot c holds what the neighbor at -o(c) sent
struct Mine {
    unsigned char *arena;
    unsigned long long mig_off[NS], bond_off[NS], gho_off[NS];
    int mcap[NS], gcap[NS];
    unsigned recv_mask;           // bit c: a sender exists for slot c
    int order[26];                // slots in the reference's ghost order
    int bpa;
};

struct CommBlob {                 // what a rank tells the others (1 KB, carried by an NCCL all-gather or by the host)
    unsigned magic;
    int rank, device, bpa;
    long long pid;
    unsigned long long arena_ptr, arena_bytes;
    cudaIpcMemHandle_t handle;
    unsigned long long mig_off[NS], bond_off[NS], gho_off[NS];
    int mcap[NS], gcap[NS];
};
static_assert(sizeof(CommBlob) <= 1024, "comm blob grew past its wire size");
constexpr unsigned BLOB_MAGIC = 0x4d45534fu;

struct CommState {                // host-side bookkeeping, hangs off meso_ctx::comm
    unsigned char *arena = nullptr;
    size_t arena_bytes = 0;
    Mine mine{};
    Peers peers{};
    bool ready = false;           // peers' arenas are mapped
    int epoch_mig = 0, epoch_halo = 0;
    int peer_rank[NS];
    std::vector<void *> opened;   // IPC mappings to close
    DevBuf<int> sendlist[NS];     // local indices of the atoms sent in direction c at the last ghost creation
    DevBuf<int> tile_counts;      // [ntiles][32]
    DevBuf<int> sync;             // last-CTA tickets, category totals
    int bpa = -1;
};

static CommState *state(meso_ctx *ctx)
{
    if (!ctx->comm) ctx->comm = new CommState();
    return static_cast<CommState *>(ctx->comm);
}

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

static void close_peers(CommState *cs)
{
    for (void *p : cs->opened) cudaIpcCloseMemHandle(p);
    cs->opened.clear();
    cs->ready = false;
}

void comm_destroy(meso_ctx *ctx)
{
    if (ctx->comm) {
        CommState *cs = static_cast<CommState *>(ctx->comm);
        close_peers(cs);
        if (cs->arena) cudaFree(cs->arena);
        delete cs;
        ctx->comm = nullptr;
    }
    if (ctx->nccl) { ncclCommDestroy((ncclComm_t)ctx->nccl); ctx->nccl = nullptr; }
}

// small host-visible reductions (thermo scalars, global atom count).  Without a communicator (the host exchanged the memory
// handles itself: several ranks of one process, or of one GPU) the values stay this rank's part and the host sums them.
int comm_allreduce_sum(meso_ctx *ctx, double *host_vals, int n)
{
    if (ctx->nranks == 1 || !ctx->nccl) return MESO_OK;
    if (!ctx->reduce_buf.reserve(std::max(64, n))) { ctx->err = "out of device memory"; return MESO_ECUDA; }
    MESO_CUDA(cudaMemcpyAsync(ctx->reduce_buf.p, host_vals, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    MESO_NCCL(ncclAllReduce(ctx->reduce_buf.p, ctx->reduce_buf.p, n, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
    MESO_CUDA(cudaMemcpyAsync(host_vals, ctx->reduce_buf.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return MESO_OK;
}

// ------------------------------------------------------------------ geometry of the 27 directions
static inline void code_offsets(int c, int o[3]) { o[0] = c % 3 - 1; o[1] = (c / 3) % 3 - 1; o[2] = c / 9 - 1; }

// rank of the brick at myloc + o, or -1 beyond a non-periodic face
static int rank_at(const meso_ctx *ctx, const int o[3])
{
    int loc[3];
    for (int d = 0; d < 3; d++) {
        loc[d] = ctx->myloc[d] + o[d];
        if (loc[d] < 0 || loc[d] >= ctx->procgrid[d]) {
            if (!ctx->box.periodic[d]) return -1;
            loc[d] = (loc[d] + ctx->procgrid[d]) % ctx->procgrid[d];
        }
    }
    return (loc[0] * ctx->procgrid[1] + loc[1]) * ctx->procgrid[2] + loc[2];
}

// The reference creates ghosts by 6 swaps (dimension d: swap 2d sends the lower slab to the lower neighbor, swap 2d+1 the upper
// slab to the upper one), and the candidates of dimension d are a rank's border atoms followed by the ghosts it received in
// the dimensions before (UM/comm_meso.cu:60-81).  An image that hopped in dimensions d1 < d2 < d3 therefore sits, in the
// receiver's ghost array, in the range of the swap of its LAST hop, behind the sender's own atoms and ordered like the
// sender's ghosts -- recursively.  Key of a direction code = (swap of the last hop, of the one before, of the first),
// absent hops = -1; ascending lexicographic order is the reference's ghost order.
static void slot_order(int order[26])
{
    struct K { int k[3], c; };
    K keys[26];
    int n = 0;
    for (int c = 0; c < NS; c++) {
        if (c == 13) continue;
        int o[3];
        code_offsets(c, o);
        K e; e.c = c; e.k[0] = e.k[1] = e.k[2] = -1;
        int p = 0;
        for (int d = 2; d >= 0; d--) if (o[d]) e.k[p++] = 2 * d + (o[d] > 0 ? 1 : 0);
        keys[n++] = e;
    }
    std::sort(keys, keys + n, [](const K &a, const K &b) {
        for (int q = 0; q < 3; q++) if (a.k[q] != b.k[q]) return a.k[q] < b.k[q];
        return false;
    });
    for (int q = 0; q < 26; q++) order[q] = keys[q].c;
}

// ------------------------------------------------------------------ arena
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// capacities of my receive slots from my own brick (a sender checks them before it writes: no agreement is needed)
static int ensure_arena(meso_ctx *ctx)
{
    CommState *cs = state(ctx);
    const Box &b = ctx->box;
    const int bpa = bonds_active(ctx) ? ctx->bond_per_atom : 0;
    double w[3], vol = 1.0;
    for (int d = 0; d < 3; d++) { w[d] = b.subhi[d] - b.sublo[d]; vol *= w[d]; }
    const double dens = std::max(1.0, (double)std::max(ctx->nlocal_host, 1) / vol);
    const double cut = ctx->cutneighmax;
    Mine m{};
    size_t off = HDR_BYTES;
    m.recv_mask = 0;
    m.bpa = bpa;
    for (int c = 0; c < NS; c++) {
        m.mcap[c] = m.gcap[c] = 0;
        m.mig_off[c] = m.bond_off[c] = m.gho_off[c] = 0;
        cs->peer_rank[c] = -1;
        if (c == 13) continue;
        int o[3], back[3];
        code_offsets(c, o);
        for (int d = 0; d < 3; d++) back[d] = -o[d];
        const int sender = rank_at(ctx, back);
        cs->peer_rank[c] = rank_at(ctx, o);
        if (sender < 0) continue;
        m.recv_mask |= 1u << c;
        double gvol = 1.0, face = 1.0;
        int nz = 0;
        for (int d = 0; d < 3; d++) { gvol *= o[d] ? cut : w[d]; face *= o[d] ? 1.0 : w[d]; nz += o[d] != 0; }
        // ghosts: the image region at the brick's density, +30 % and a floor for small / inhomogeneous systems
        m.gcap[c] = (int)(gvol * dens * 1.3) + 1024;
        // leavers per rebuild through a face: a layer ~0.05 thick at the deck's temperature; generous floors elsewhere
        m.mcap[c] = nz == 1 ? std::max(4096, (int)(face * dens * 0.25)) : 1024;
        m.mig_off[c] = off; off = align_up(off + (size_t)m.mcap[c] * RECM * 8, 256);
        m.bond_off[c] = off; off = align_up(off + (size_t)m.mcap[c] * (1 + bpa) * 8, 256);
        m.gho_off[c] = off; off = align_up(off + 2 * (size_t)m.gcap[c] * RECB * 8, 256);
    }
    slot_order(m.order);
    bool fits = cs->arena && off <= cs->arena_bytes && bpa == cs->bpa && m.recv_mask == cs->mine.recv_mask;
    for (int c = 0; c < NS && fits; c++) fits = m.gcap[c] <= cs->mine.gcap[c] && m.mcap[c] <= cs->mine.mcap[c];
    if (fits) return MESO_OK;                                  // the mapped arena still serves: nothing to re-publish
    close_peers(cs);
    if (cs->arena) { cudaFree(cs->arena); cs->arena = nullptr; }
    MESO_CUDA(cudaMalloc(&cs->arena, off));
    MESO_CUDA(cudaMemsetAsync(cs->arena, 0, HDR_BYTES, ctx->stream));
    cs->arena_bytes = off;
    m.arena = cs->arena;
    cs->mine = m;
    cs->bpa = bpa;
    // the epoch counters are NOT reset: every rank advances them in step, whatever happens to one rank's arena
    if (!cs->sync.reserve(128)) { ctx->err = "out of device memory"; return MESO_ECUDA; }
    MESO_CUDA(cudaMemsetAsync(cs->sync.p, 0, 128 * sizeof(int), ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return MESO_OK;
}

static void fill_blob(meso_ctx *ctx, CommBlob &bl)
{
    CommState *cs = state(ctx);
    memset(&bl, 0, sizeof bl);
    bl.magic = BLOB_MAGIC; bl.rank = ctx->rank; bl.device = ctx->device; bl.bpa = cs->mine.bpa;
    bl.pid = (long long)getpid();
    bl.arena_ptr = (unsigned long long)(uintptr_t)cs->arena; bl.arena_bytes = cs->arena_bytes;
    if (cudaIpcGetMemHandle(&bl.handle, cs->arena) != cudaSuccess) cudaGetLastError();   // same-process peers use arena_ptr
    for (int c = 0; c < NS; c++) {
        bl.mig_off[c] = cs->mine.mig_off[c]; bl.bond_off[c] = cs->mine.bond_off[c]; bl.gho_off[c] = cs->mine.gho_off[c];
        bl.mcap[c] = cs->mine.mcap[c]; bl.gcap[c] = cs->mine.gcap[c];
    }
}

// map the arenas of my (at most 26 distinct) neighbors: IPC between processes, plain pointers inside one
static int import_blobs(meso_ctx *ctx, const unsigned char *blobs, int nranks)
{
    CommState *cs = state(ctx);
    close_peers(cs);
    std::vector<unsigned char *> base((size_t)nranks, nullptr);
    Peers p{};
    p.bpa = cs->mine.bpa;
    for (int c = 0; c < NS; c++) {
        p.arena[c] = nullptr;
        p.mig_off[c] = p.bond_off[c] = p.gho_off[c] = 0;
        p.mcap[c] = p.gcap[c] = 0;
        if (c == 13) continue;
        const int r = cs->peer_rank[c];
        if (r < 0) continue;
        CommBlob bl;
        memcpy(&bl, blobs + (size_t)r * 1024, sizeof bl);
        if (bl.magic != BLOB_MAGIC || bl.rank != r) { ctx->err = "comm import: malformed blob"; return MESO_EINVAL; }
        if (bl.bpa != cs->mine.bpa) { ctx->err = "comm import: ranks disagree on bond_per_atom"; return MESO_EINVAL; }
        if (!base[r]) {
            if (r == ctx->rank) base[r] = cs->arena;
            else if (bl.pid == (long long)getpid()) {
                if (bl.device != ctx->device) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(bl.device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ctx->err = "cudaDeviceEnablePeerAccess failed"; return MESO_ECUDA; }
                    cudaGetLastError();
                }
                base[r] = (unsigned char *)(uintptr_t)bl.arena_ptr;
            } else {
                void *q = nullptr;
                MESO_CUDA(cudaIpcOpenMemHandle(&q, bl.handle, cudaIpcMemLazyEnablePeerAccess));
                cs->opened.push_back(q);
                base[r] = (unsigned char *)q;
            }
        }
        p.arena[c] = base[r];
        p.mig_off[c] = bl.mig_off[c]; p.bond_off[c] = bl.bond_off[c]; p.gho_off[c] = bl.gho_off[c];
        p.mcap[c] = bl.mcap[c]; p.gcap[c] = bl.gcap[c];
        if (!cs->sendlist[c].reserve((size_t)std::max(p.gcap[c], 1))) { ctx->err = "out of device memory (send lists)"; return MESO_ECUDA; }
    }
    cs->peers = p;
    cs->ready = true;
    return MESO_OK;
}

// called at the first rebuild after an upload: every rank reaches it at the same point
static int ensure_peers(meso_ctx *ctx)
{
    int rc = ensure_arena(ctx);
    if (rc) return rc;
    CommState *cs = state(ctx);
    if (cs->ready) return MESO_OK;
    if (ctx->nranks == 1) {
        CommBlob bl;
        fill_blob(ctx, bl);
        unsigned char buf[1024] = {0};
        memcpy(buf, &bl, sizeof bl);
        return import_blobs(ctx, buf, 1);
    }
    if (!ctx->nccl) { ctx->err = "decomposition without a communicator: exchange meso_comm_export blobs and call meso_comm_import before meso_setup"; return MESO_EINVAL; }
    // all-gather of the 1 KB blobs through NCCL (bootstrap only)
    DevBuf<unsigned char> dev;
    if (!dev.reserve((size_t)1024 * (ctx->nranks + 1))) { ctx->err = "out of device memory"; return MESO_ECUDA; }
    std::vector<unsigned char> host((size_t)1024 * ctx->nranks, 0);
    CommBlob bl;
    fill_blob(ctx, bl);
    unsigned char mine[1024] = {0};
    memcpy(mine, &bl, sizeof bl);
    unsigned char *sendp = dev.p + (size_t)1024 * ctx->nranks;
    MESO_CUDA(cudaMemcpyAsync(sendp, mine, 1024, cudaMemcpyHostToDevice, ctx->stream));
    MESO_NCCL(ncclAllGather(sendp, dev.p, 1024, ncclChar, (ncclComm_t)ctx->nccl, ctx->stream));
    MESO_CUDA(cudaMemcpyAsync(host.data(), dev.p, host.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return import_blobs(ctx, host.data(), ctx->nranks);
}

// ------------------------------------------------------------------ flags
__device__ __forceinline__ void st_sys(int *p, int v) { asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_sys(const int *p) { int v; asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// A sender may overwrite a slot only after the receiver has consumed what it held: the receiver acknowledges every epoch in
// the SENDER's header (ack_*[c]), and a pack kernel starts by waiting for the acknowledgement of the previous use of its
// target buffer (migration: epoch - 1; halo, double-buffered: epoch - 2).  With double buffering the wait never blocks in
// practice; it is what makes the protocol safe when ranks drift apart (or time-share one GPU).
__device__ __forceinline__ void wait_ack(const Mine &mine, Counts *cnt, const Peers &peers, int c, int need, int kind)
{
    if (need <= 0 || c == 13 || !peers.arena[c]) return;
    const ArenaHdr *h = reinterpret_cast<const ArenaHdr *>(mine.arena);
    const int *a = kind == 0 ? &h->ack_mig[c] : &h->ack_halo[c];
    const long long t0 = clock64();
    while (ld_sys(a) < need) {
        if (clock64() - t0 > 40000000000LL) { atomicOr(&cnt->err, 16); break; }
        __nanosleep(100);
    }
}

// End of a pack kernel: every CTA has fenced its remote stores; the last one publishes the per-direction record counts and
// error bits, then the epoch flags.  kind: 0 = migration, 1 = ghost creation, 2 = refresh.
__device__ __forceinline__ void publish(const Peers &peers, int *ticket, const int *counts, int errbits, int epoch, int kind)
{
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1) == (int)(gridDim.x * gridDim.y) - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    const int c = threadIdx.x;
    if (c < NS && c != 13 && peers.arena[c]) {
        ArenaHdr *h = reinterpret_cast<ArenaHdr *>(peers.arena[c]);
        if (kind == 0) { st_sys(&h->cnt_mig[c], counts[c]); st_sys(&h->err[c], errbits); }
        else if (kind == 1) { st_sys(&h->cnt_halo[c], counts[c]); st_sys(&h->err[c], errbits); }
        __threadfence_system();
        st_sys(kind == 0 ? &h->flag_mig[c] : &h->flag_halo[epoch & 1][c], epoch);
    }
    if (threadIdx.x == 0) *ticket = 0;
}

// End of an unpack kernel: the last CTA tells every sender that its slot is free again (slot c was filled by the neighbor at
// -o(c), which sent in direction c: its arena is peers.arena[26 - c])
__device__ __forceinline__ void acknowledge(const Mine &mine, const Peers &peers, int *ticket, int epoch, int kind)
{
    __syncthreads();
    __shared__ int s_last_ack;
    if (threadIdx.x == 0) { __threadfence(); s_last_ack = atomicAdd(ticket, 1) == (int)(gridDim.x * gridDim.y) - 1; }
    __syncthreads();
    if (!s_last_ack) return;
    const int c = threadIdx.x;
    if (c < NS && c != 13 && ((mine.recv_mask >> c) & 1u) && peers.arena[26 - c]) {
        ArenaHdr *h = reinterpret_cast<ArenaHdr *>(peers.arena[26 - c]);
        st_sys(kind == 0 ? &h->ack_mig[c] : &h->ack_halo[c], epoch);
    }
    if (threadIdx.x == 0) *ticket = 0;
}

// one warp: lane c waits until the neighbor that fills slot c has published `epoch` (bounded: a lost peer sets an error bit
// instead of hanging the GPU); sender-side errors are adopted.  kind 0: migration, otherwise halo.
__global__ void k_wait(Mine mine, Counts *cnt, int epoch, int kind)
{
    const int c = threadIdx.x;
    if (c >= NS || !((mine.recv_mask >> c) & 1u)) return;
    ArenaHdr *h = reinterpret_cast<ArenaHdr *>(mine.arena);
    const int *f = kind == 0 ? &h->flag_mig[c] : &h->flag_halo[epoch & 1][c];
    const long long t0 = clock64();
    while (ld_sys(f) < epoch) {
        if (clock64() - t0 > 40000000000LL) { atomicOr(&cnt->err, 16); break; }      // ~20 s
        __nanosleep(100);
    }
    __threadfence_system();
    if (kind < 2) { const int e = ld_sys(&h->err[c]); if (e) atomicOr(&cnt->err, e); }
}

// ------------------------------------------------------------------ ordered compaction into 27 categories
// Every candidate belongs to a set of categories (bit mask).  Rank of a candidate inside category c = number of earlier
// candidates (by index) in c.  Tiles of CTILE candidates in blocked order (round r covers CT consecutive candidates).
// tile_counts[tile][32]: per-tile totals, exclusive-scanned over the tiles by k_mc_scan.
template <typename F>
__device__ __forceinline__ void mc_count(F mask_of, int first, int last, int *tile_counts, int ntiles)
{
    __shared__ int wsum[CT / 32][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int base = first + tile * CTILE;
        if (base >= last) break;
        int mine = 0;                                          // lane c: this warp's count of category c
#pragma unroll
        for (int r = 0; r < CI; r++) {
            const int i = base + r * CT + threadIdx.x;
            const unsigned m = i < last ? mask_of(i) : 0u;
            const unsigned any = __reduce_or_sync(0xffffffffu, m);
            for (unsigned rest = any; rest; rest &= rest - 1) {
                const int c = __ffs(rest) - 1;
                const unsigned b = __ballot_sync(0xffffffffu, (m >> c) & 1u);
                if (lane == c) mine += __popc(b);
            }
        }
        wsum[w][lane] = mine;
        __syncthreads();
        if (w == 0) {
            int t = 0;
#pragma unroll
            for (int ww = 0; ww < CT / 32; ww++) t += wsum[ww][lane];
            tile_counts[(size_t)tile * 32 + lane] = t;
        }
        __syncthreads();
    }
}

// one warp per category: exclusive scan of its column over the used tiles; totals to tot[c]
// range_kind 0: all locals (migration), 1: the border atoms (ghost creation)
__global__ void __launch_bounds__(1024) k_mc_scan(int *__restrict__ tile_counts, const Counts *__restrict__ cnt, int range_kind, int *__restrict__ tot)
{
    const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
    if (c >= NS) return;
    const int first = range_kind == 0 ? 0 : cnt->n_bulk, last = cnt->nlocal;
    const int used = (max(last - first, 0) + CTILE - 1) / CTILE;
    int carry = 0;
    for (int t0 = 0; t0 < used; t0 += 32) {
        const int t = t0 + lane;
        const int v = t < used ? tile_counts[(size_t)t * 32 + c] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (t < used) tile_counts[(size_t)t * 32 + c] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) tot[c] = carry;
}

// ranks of the candidates of one tile: calls emit(i, c, rank) for every (candidate, category) pair
template <typename F, typename E>
__device__ __forceinline__ void mc_place(F mask_of, E emit, int first, int last, const int *tile_counts, int ntiles)
{
    __shared__ unsigned bal[CI][CT / 32][32];
    __shared__ int pre[CI][CT / 32][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int base = first + tile * CTILE;
        if (base >= last) break;
        unsigned m[CI];
#pragma unroll
        for (int r = 0; r < CI; r++) {
            const int i = base + r * CT + threadIdx.x;
            m[r] = i < last ? mask_of(i) : 0u;
            bal[r][w][lane] = 0u;
            const unsigned any = __reduce_or_sync(0xffffffffu, m[r]);
            __syncwarp();
            for (unsigned rest = any; rest; rest &= rest - 1) {
                const int c = __ffs(rest) - 1;
                const unsigned b = __ballot_sync(0xffffffffu, (m[r] >> c) & 1u);
                if (lane == c) bal[r][w][c] = b;
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {                                // lane = category: prefix over the (round, warp) chunks of the tile
            int run = tile_counts[(size_t)tile * 32 + lane];
#pragma unroll
            for (int r = 0; r < CI; r++)
#pragma unroll
                for (int ww = 0; ww < CT / 32; ww++) { pre[r][ww][lane] = run; run += __popc(bal[r][ww][lane]); }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < CI; r++) {
            const int i = base + r * CT + threadIdx.x;
            for (unsigned rest = m[r]; rest; rest &= rest - 1) {
                const int c = __ffs(rest) - 1;
                emit(i, c, pre[r][w][c] + __popc(bal[r][w][c] & lt));
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ migration
// destination of a local atom after the periodic wrap: `x >= hi || x < lo`, direction by the minimum image of x - mid
// (UM/comm_meso.cu:303-330), all communicating dimensions at once
struct MigGeom { int multi[3]; };

__device__ __forceinline__ int leave_dir(double xd, const Box &box, int d)
{
    if (xd >= box.subhi[d] || xd < box.sublo[d]) {
        double dist = xd - 0.5 * (box.sublo[d] + box.subhi[d]);
        if (box.periodic[d] && fabs(dist) > 0.5 * box.prd[d]) dist += dist < 0.0 ? box.prd[d] : -box.prd[d];
        return dist < 0 ? -1 : 1;
    }
    return 0;
}

__device__ __forceinline__ unsigned mig_mask(const SoA3 &x, int i, const Box &box, const MigGeom &g)
{
    int code = 0, w = 1;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const int o = g.multi[d] ? leave_dir(x.c[d][i], box, d) : 0;
        code += (o + 1) * w;
        w *= 3;
    }
    return 1u << code;
}

__global__ void __launch_bounds__(CT) k_mig_count(SoA3 x, const Counts *__restrict__ cnt, int *__restrict__ tile_counts, Box box, MigGeom g,
                                                  int ntiles)
{
    mc_count([&](int i) { return mig_mask(x, i, box, g); }, 0, cnt->nlocal, tile_counts, ntiles);
}

// stayers (category 13) are compacted in order into the alternate arrays; leavers become records in the slot of their
// destination brick.  tot[c] = number of atoms of category c.
__global__ void __launch_bounds__(CT) k_mig_pack(SoA3 x, SoA3 v, const int *__restrict__ tag, const int *__restrict__ type,
                                                 const int *__restrict__ mask, const int *__restrict__ image, SoA3 xo, SoA3 vo,
                                                 int *__restrict__ tago, int *__restrict__ typeo, int *__restrict__ masko,
                                                 int *__restrict__ imageo, Counts *__restrict__ cnt, const int *__restrict__ tile_counts,
                                                 const int *__restrict__ tot, Box box, MigGeom g, Peers peers, int ntiles,
                                                 const int *__restrict__ nbond, const int2 *__restrict__ bonds, int *__restrict__ nbond_o,
                                                 int2 *__restrict__ bonds_o, size_t padding, int *ticket, int *sendn, int epoch, Mine mine)
{
    const int bpa = peers.bpa;
    if (threadIdx.x < NS) wait_ack(mine, cnt, peers, threadIdx.x, epoch - 1, 0);
    __syncthreads();
    mc_place([&](int i) { return mig_mask(x, i, box, g); },
             [&](int i, int c, int k) {
                 if (c == 13) {
#pragma unroll
                     for (int q = 0; q < 3; q++) { xo.c[q][k] = x.c[q][i]; vo.c[q][k] = v.c[q][i]; }
                     tago[k] = tag[i]; typeo[k] = type[i]; masko[k] = mask[i]; imageo[k] = image[i];
                     if (bpa) {
                         const int nb = nbond[i];
                         nbond_o[k] = nb;
                         for (int q = 0; q < nb; q++) bonds_o[k + q * padding] = bonds[i + q * padding];
                     }
                 } else if (!peers.arena[c]) atomicOr(&cnt->err, 8);                 // beyond a non-periodic face: lost
                 else if (k < peers.mcap[c]) {
                     double *rec = reinterpret_cast<double *>(peers.arena[c] + peers.mig_off[c]) + (size_t)k * RECM;
#pragma unroll
                     for (int q = 0; q < 3; q++) { rec[q] = x.c[q][i]; rec[3 + q] = v.c[q][i]; }
                     reinterpret_cast<int2 *>(rec)[6] = make_int2(tag[i], type[i]);
                     reinterpret_cast<int2 *>(rec)[7] = make_int2(mask[i], image[i]);
                     if (bpa) {
                         int2 *br = reinterpret_cast<int2 *>(peers.arena[c] + peers.bond_off[c]) + (size_t)k * (1 + bpa);
                         const int nb = nbond[i];
                         br[0] = make_int2(nb, 0);
                         for (int q = 0; q < nb; q++) br[1 + q] = bonds[i + q * padding];
                     }
                 }
             },
             0, cnt->nlocal, tile_counts, ntiles);
    // counts for the headers (clamped; an overflow would lose atoms: error bit 8 travels with the message).  Every CTA computes
    // the same values; the last one publishes them.
    if (threadIdx.x < NS) {
        const int c = threadIdx.x;
        int n = tot[c];
        if (c != 13 && peers.arena[c] && n > peers.mcap[c]) { n = peers.mcap[c]; atomicOr(&cnt->err, 8); }
        if (blockIdx.x == 0) sendn[c] = c == 13 ? 0 : n;
    }
    __syncthreads();
    publish(peers, ticket, sendn, cnt->err & 8, epoch, 0);
}

__global__ void k_mig_shrink(Counts *cnt, const int *tot) { cnt->nlocal = tot[13]; cnt->nall = tot[13]; }

// arrivals are appended slot by slot (ascending direction code), inside a slot in the sender's index order
__global__ void __launch_bounds__(256) k_mig_unpack(SoA3 x, SoA3 v, int *__restrict__ tag, int *__restrict__ type, int *__restrict__ mask,
                                                    int *__restrict__ image, Counts *__restrict__ cnt, Mine mine, Box box, MigGeom g,
                                                    int nloc_cap, int *__restrict__ nbond, int2 *__restrict__ bonds, size_t padding,
                                                    Peers peers, int *ticket, int epoch)
{
    const int s = blockIdx.y;
    const ArenaHdr *h = reinterpret_cast<const ArenaHdr *>(mine.arena);
    int base = cnt->nlocal, total = 0;
    for (int q = 0; q < NS; q++) {
        if (!((mine.recv_mask >> q) & 1u)) continue;
        const int nq = min(max(h->cnt_mig[q], 0), mine.mcap[q]);
        if (q < s) base += nq;
        total += nq;
    }
    const bool fits = cnt->nlocal + total <= nloc_cap;
    if (!fits && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicOr(&cnt->err, 8);
    const int n = (fits && ((mine.recv_mask >> s) & 1u)) ? min(max(h->cnt_mig[s], 0), mine.mcap[s]) : 0;
    const double *buf = reinterpret_cast<const double *>(mine.arena + mine.mig_off[s]);
    const int2 *bbuf = reinterpret_cast<const int2 *>(mine.arena + mine.bond_off[s]);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double *rec = buf + (size_t)k * RECM;
        const int p = base + k;
        bool inside = true;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            x.c[q][p] = rec[q]; v.c[q][p] = rec[3 + q];
            if (g.multi[q]) inside = inside && rec[q] >= box.sublo[q] && rec[q] < box.subhi[q];
        }
        const int2 tt = reinterpret_cast<const int2 *>(rec)[6], mi = reinterpret_cast<const int2 *>(rec)[7];
        tag[p] = tt.x; type[p] = tt.y; mask[p] = mi.x; image[p] = mi.y;
        if (!inside) atomicOr(&cnt->err, 8);                     // an arrival outside the brick is dropped by the reference too
        if (mine.bpa) {
            const int2 *br = bbuf + (size_t)k * (1 + mine.bpa);
            const int nb = min(max(br[0].x, 0), mine.bpa);
            nbond[p] = nb;
            for (int q = 0; q < nb; q++) bonds[p + q * padding] = br[1 + q];
        }
    }
    acknowledge(mine, peers, ticket, epoch, 0);
}

__global__ void k_mig_grow(Counts *cnt, Mine mine, int nloc_cap)
{
    const ArenaHdr *h = reinterpret_cast<const ArenaHdr *>(mine.arena);
    int total = 0;
    for (int q = 0; q < NS; q++) if ((mine.recv_mask >> q) & 1u) total += min(max(h->cnt_mig[q], 0), mine.mcap[q]);
    if (cnt->nlocal + total <= nloc_cap) cnt->nlocal += total;
    cnt->nall = cnt->nlocal;
}

static int tiles_for(meso_ctx *ctx) { return (int)((ctx->cap + CTILE - 1) / CTILE); }

int launch_exchange_multi(meso_ctx *ctx)
{
    int rc = ensure_peers(ctx);
    if (rc) return rc;
    CommState *cs = state(ctx);
    MigGeom g;
    bool any = false;
    for (int d = 0; d < 3; d++) { g.multi[d] = ctx->procgrid[d] > 1; any = any || g.multi[d]; }
    if (!any) return MESO_OK;                     // the periodic wrap already put every atom back into the brick
    const Box &box = ctx->box;
    const int ntiles = tiles_for(ctx);
    if (!cs->tile_counts.reserve((size_t)ntiles * 32 + 64)) { ctx->err = "out of device memory (tile counts)"; return MESO_ECUDA; }
    int *tc = cs->tile_counts.p, *tot = cs->sync.p + 8, *sendn = cs->sync.p + 40;
    cudaStream_t st = ctx->stream;
    const bool bonded = bonds_active(ctx);
    const int epoch = ++cs->epoch_mig;
    k_mig_count<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(soa(ctx->x), ctx->d_counts, tc, box, g, ntiles);
    k_mc_scan<<<1, 1024, 0, LS(st)>>>(tc, ctx->d_counts, 0, tot);
    k_mig_pack<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->image.p, soa(ctx->xa),
                                                 soa(ctx->va), ctx->taga.p, ctx->typea.p, ctx->maska.p, ctx->imagea.p, ctx->d_counts, tc, tot, box,
                                                 g, cs->peers, ntiles, ctx->nbond.p, ctx->bonds.p, ctx->nbond_alt.p, ctx->bonds_alt.p, ctx->cap,
                                                 cs->sync.p + 0, sendn, epoch, cs->mine);
    for (int q = 0; q < 3; q++) { std::swap(ctx->x[q].p, ctx->xa[q].p); std::swap(ctx->v[q].p, ctx->va[q].p); }
    std::swap(ctx->tag.p, ctx->taga.p); std::swap(ctx->type.p, ctx->typea.p);
    std::swap(ctx->mask.p, ctx->maska.p); std::swap(ctx->image.p, ctx->imagea.p);
    if (bonded) {
        std::swap(ctx->nbond.p, ctx->nbond_alt.p); std::swap(ctx->nbond.cap, ctx->nbond_alt.cap);
        std::swap(ctx->bonds.p, ctx->bonds_alt.p); std::swap(ctx->bonds.cap, ctx->bonds_alt.cap);
    }
    k_mig_shrink<<<1, 1, 0, LS(st)>>>(ctx->d_counts, tot);
    k_wait<<<1, 32, 0, LS(st)>>>(cs->mine, ctx->d_counts, epoch, 0);
    const dim3 grid(std::max(1, ctx->sm_count / 4), NS);
    k_mig_unpack<<<grid, 256, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->image.p, ctx->d_counts, cs->mine, box,
                                        g, (int)ctx->nloc_cap, ctx->nbond.p, ctx->bonds.p, ctx->cap, cs->peers, cs->sync.p + 3, epoch);
    k_mig_grow<<<1, 1, 0, LS(st)>>>(ctx->d_counts, cs->mine, (int)ctx->nloc_cap);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// ------------------------------------------------------------------ ghost creation
// An atom goes in direction o iff, in every dimension with o_d != 0, it lies in the send slab of that side
// (x <= sublo + cutghost for o_d = -1, x >= subhi - cutghost for o_d = +1; src/comm.cpp:590-625) -- the product of the
// per-dimension choices {stay, lower?, upper?} the 6 swaps would have made one after the other.
__device__ __forceinline__ unsigned ghost_mask(const SoA3 &x, int i, const Box &box)
{
    unsigned m[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double xd = x.c[d][i];
        m[d] = 2u | ((box.sendflag[2 * d] && xd <= box.slab_lo_hi[d]) ? 1u : 0u) | ((box.sendflag[2 * d + 1] && xd >= box.slab_hi_lo[d]) ? 4u : 0u);
    }
    const unsigned row = m[0];
    const unsigned plane = ((m[1] & 1u) ? row : 0u) | ((m[1] & 2u) ? row << 3 : 0u) | ((m[1] & 4u) ? row << 6 : 0u);
    const unsigned all = ((m[2] & 1u) ? plane : 0u) | ((m[2] & 2u) ? plane << 9 : 0u) | ((m[2] & 4u) ? plane << 18 : 0u);
    return all & ~(1u << 13);
}

__global__ void __launch_bounds__(CT) k_gho_count(SoA3 x, const Counts *__restrict__ cnt, int *__restrict__ tile_counts, Box box, int ntiles)
{
    mc_count([&](int i) { return ghost_mask(x, i, box); }, cnt->n_bulk, cnt->nlocal, tile_counts, ntiles);
}

// per-direction periodic shift of the images this rank sends (+prd across the lower box face, -prd across the upper one)
struct Shifts { double s[NS][3]; };

// pack_border_vel (UM/atom_vec_dpd_atomic_meso.cu:61-100): record = {x+shift (3), v (3), tag|type, mask|signature}
__global__ void __launch_bounds__(CT) k_gho_pack(SoA3 x, SoA3 v, const int *__restrict__ tag, const int *__restrict__ type,
                                                 const int *__restrict__ mask, const float4 *__restrict__ veloc4, Counts *__restrict__ cnt,
                                                 const int *__restrict__ tile_counts, const int *__restrict__ tot, Box box, Peers peers,
                                                 Shifts sh, int *const *sendlist, int ntiles, int *ticket, int epoch, Mine mine)
{
    if (threadIdx.x < NS) wait_ack(mine, cnt, peers, threadIdx.x, epoch - 2, 1);
    __syncthreads();
    const size_t parity = (size_t)(epoch & 1);
    mc_place([&](int i) { return ghost_mask(x, i, box); },
             [&](int i, int c, int k) {
                 if (!peers.arena[c] || k >= peers.gcap[c]) return;
                 double *rec = reinterpret_cast<double *>(peers.arena[c] + peers.gho_off[c] + parity * peers.gcap[c] * RECB * 8) + (size_t)k * RECB;
                 rec[0] = x.c[0][i] + sh.s[c][0]; rec[1] = x.c[1][i] + sh.s[c][1]; rec[2] = x.c[2][i] + sh.s[c][2];
                 rec[3] = v.c[0][i]; rec[4] = v.c[1][i]; rec[5] = v.c[2][i];
                 reinterpret_cast<int2 *>(rec)[6] = make_int2(tag[i], type[i]);
                 reinterpret_cast<int2 *>(rec)[7] = make_int2(mask[i], __float_as_int(veloc4[i].w));
                 sendlist[c][k] = i;
             },
             cnt->n_bulk, cnt->nlocal, tile_counts, ntiles);
    if (threadIdx.x < NS) {
        const int c = threadIdx.x;
        int n = (c == 13 || !peers.arena[c]) ? 0 : tot[c];
        if (n > peers.gcap[c]) { n = peers.gcap[c]; atomicOr(&cnt->err, 1); }
        if (blockIdx.x == 0) cnt->route_send_n[c] = n;
    }
    __syncthreads();
    publish(peers, ticket, cnt->route_send_n, cnt->err & 1, epoch, 1);
}

// ghost ranges: the 26 slots in the reference's order; also the 6 swap ranges it would have produced
__global__ void k_gho_bases(Counts *cnt, Mine mine, int cap)
{
    const ArenaHdr *h = reinterpret_cast<const ArenaHdr *>(mine.arena);
    int base = 0;
    for (int s = 0; s < 6; s++) { cnt->swap_first[s] = cnt->nlocal; cnt->swap_n[s] = 0; }
    for (int q = 0; q < NS; q++) { cnt->route_recv_n[q] = 0; cnt->slot_base[q] = 0; }
    int cur_swap = -1;
    for (int q = 0; q < 26; q++) {
        const int c = mine.order[q];
        int n = ((mine.recv_mask >> c) & 1u) ? min(max(h->cnt_halo[c], 0), mine.gcap[c]) : 0;
        if (cnt->nlocal + base + n > cap) { cnt->err |= 1; n = 0; }
        // swap of the last hop: highest dimension with a non-zero offset
        const int o[3] = {c % 3 - 1, (c / 3) % 3 - 1, c / 9 - 1};
        const int d = o[2] ? 2 : (o[1] ? 1 : 0);
        const int swap = 2 * d + (o[d] > 0 ? 1 : 0);
        if (swap != cur_swap) { cnt->swap_first[swap] = cnt->nlocal + base; cur_swap = swap; }
        cnt->swap_n[swap] += n;
        cnt->slot_base[c] = base;
        cnt->route_recv_n[c] = n;
        base += n;
    }
    cnt->nghost = base;
    cnt->nall = cnt->nlocal + base;
    cnt->max_pair = 0;
}

// unpack_border_vel (UM/atom_vec_dpd_atomic_meso.cu:137-163) + packing of the new ghosts
__global__ void __launch_bounds__(256) k_gho_unpack(SoA3 x, SoA3 v, int *__restrict__ tag, int *__restrict__ type, int *__restrict__ mask,
                                                    float4 *__restrict__ coord4, float4 *__restrict__ veloc4,
                                                    const Counts *__restrict__ cnt, Mine mine, Box box, Peers peers, int *ticket, int epoch)
{
    const int s = blockIdx.y;
    const int n = cnt->route_recv_n[s];
    const int g0 = cnt->nlocal + cnt->slot_base[s];
    const double *buf = reinterpret_cast<const double *>(mine.arena + mine.gho_off[s] + (size_t)(epoch & 1) * mine.gcap[s] * RECB * 8);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double *rec = buf + (size_t)k * RECB;
        const int g = g0 + k;
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
    acknowledge(mine, peers, ticket, epoch, 1);
}

static void make_shifts(const meso_ctx *ctx, Shifts &sh)
{
    const Box &b = ctx->box;
    for (int c = 0; c < NS; c++) {
        int o[3];
        code_offsets(c, o);
        for (int d = 0; d < 3; d++) {
            // an image sent to the lower neighbor from the lowest brick appears beyond the upper box face: +prd (and vice versa)
            double s = 0.0;
            if (o[d] < 0 && b.pbc[2 * d]) s = b.pbc[2 * d] * b.prd[d];
            if (o[d] > 0 && b.pbc[2 * d + 1]) s = b.pbc[2 * d + 1] * b.prd[d];
            sh.s[c][d] = s;
        }
    }
}

int launch_borders_multi(meso_ctx *ctx)
{
    int rc = ensure_peers(ctx);
    if (rc) return rc;
    CommState *cs = state(ctx);
    const Box &box = ctx->box;
    const int ntiles = tiles_for(ctx);
    if (!cs->tile_counts.reserve((size_t)ntiles * 32 + 64) || !ctx->route_ptrs.reserve(64)) { ctx->err = "out of device memory (tile counts)"; return MESO_ECUDA; }
    int *tc = cs->tile_counts.p, *tot = cs->sync.p + 8;
    cudaStream_t st = ctx->stream;
    void *hp[NS];
    for (int c = 0; c < NS; c++) hp[c] = cs->sendlist[c].p;
    MESO_CUDA(cudaMemcpyAsync(ctx->route_ptrs.p, hp, sizeof hp, cudaMemcpyHostToDevice, st));
    Shifts sh;
    make_shifts(ctx, sh);
    const int epoch = ++cs->epoch_halo;
    k_gho_count<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(soa(ctx->x), ctx->d_counts, tc, box, ntiles);
    k_mc_scan<<<1, 1024, 0, LS(st)>>>(tc, ctx->d_counts, 1, tot);
    k_gho_pack<<<grid_for(ctx, 4), CT, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->veloc4.p, ctx->d_counts, tc,
                                                 tot, box, cs->peers, sh, reinterpret_cast<int *const *>(ctx->route_ptrs.p), ntiles,
                                                 cs->sync.p + 1, epoch, cs->mine);
    k_wait<<<1, 32, 0, LS(st)>>>(cs->mine, ctx->d_counts, epoch, 1);
    k_gho_bases<<<1, 1, 0, LS(st)>>>(ctx->d_counts, cs->mine, (int)ctx->cap);
    const dim3 grid(std::max(1, ctx->sm_count / 2), NS);
    k_gho_unpack<<<grid, 256, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->tag.p, ctx->type.p, ctx->mask.p, ctx->coord4.p, ctx->veloc4.p,
                                        ctx->d_counts, cs->mine, box, cs->peers, cs->sync.p + 4, epoch);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// ------------------------------------------------------------------ per-step refresh (pack_comm_vel / unpack_comm_vel)
// Records are 40 bytes: {x + shift (3 x fp64), veloc4 = fp32 v + this step's signature}.  The force kernel reads ghosts only
// through the packed views, so the fp64 ghost velocity is simply the widened fp32 value.  The send lists and the slot
// ranges are those of the last ghost creation; sizes live on the device.
__global__ void __launch_bounds__(256) k_fwd_pack(SoA3 x, const float4 *__restrict__ veloc4, const Counts *__restrict__ cnt, Peers peers,
                                                  Shifts sh, int *const *sendlist, int *ticket, int epoch, Mine mine, Counts *cntw)
{
    const int c = blockIdx.y;
    const int parity = epoch & 1;
    if (threadIdx.x == 0) wait_ack(mine, cntw, peers, c, epoch - 2, 1);
    __syncthreads();
    const int n = cnt->route_send_n[c];
    if (n > 0 && peers.arena[c]) {
        const int *list = sendlist[c];
        double *buf = reinterpret_cast<double *>(peers.arena[c] + peers.gho_off[c] + (size_t)parity * peers.gcap[c] * RECB * 8);
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
            const int i = list[k];
            double *rec = buf + (size_t)k * RECF;
            rec[0] = x.c[0][i] + sh.s[c][0]; rec[1] = x.c[1][i] + sh.s[c][1]; rec[2] = x.c[2][i] + sh.s[c][2];
            const float4 w = veloc4[i];
            reinterpret_cast<float2 *>(rec)[3] = make_float2(w.x, w.y);
            reinterpret_cast<float2 *>(rec)[4] = make_float2(w.z, w.w);
        }
    }
    publish(peers, ticket, nullptr, 0, epoch, 2);
}

__global__ void __launch_bounds__(256) k_fwd_unpack(SoA3 x, SoA3 v, const int *__restrict__ type, float4 *__restrict__ coord4,
                                                    float4 *__restrict__ veloc4, const Counts *__restrict__ cnt, Mine mine, Box box, Peers peers,
                                                    int *ticket, int epoch)
{
    const int s = blockIdx.y;
    const int parity = epoch & 1;
    const int n = cnt->route_recv_n[s];
    const int g0 = cnt->nlocal + cnt->slot_base[s];
    const double *buf = reinterpret_cast<const double *>(mine.arena + mine.gho_off[s] + (size_t)parity * mine.gcap[s] * RECB * 8);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const double *rec = buf + (size_t)k * RECF;
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
    acknowledge(mine, peers, ticket, epoch, 1);
}

// per-step ghost refresh on stream `st` (the side stream when it overlaps the bulk force kernel)
int launch_forward_multi(meso_ctx *ctx, cudaStream_t st)
{
    CommState *cs = state(ctx);
    if (!cs->ready) { ctx->err = "halo refresh before the first rebuild"; return MESO_EINVAL; }
    Shifts sh;
    make_shifts(ctx, sh);
    const int epoch = ++cs->epoch_halo;
    const dim3 grid(std::max(1, ctx->sm_count / 4), NS);
    k_fwd_pack<<<grid, 256, 0, LS(st)>>>(soa(ctx->x), ctx->veloc4.p, ctx->d_counts, cs->peers, sh, reinterpret_cast<int *const *>(ctx->route_ptrs.p),
                                      cs->sync.p + 2, epoch, cs->mine, ctx->d_counts);
    k_wait<<<1, 32, 0, LS(st)>>>(cs->mine, ctx->d_counts, epoch, 2);
    k_fwd_unpack<<<grid, 256, 0, LS(st)>>>(soa(ctx->x), soa(ctx->v), ctx->type.p, ctx->coord4.p, ctx->veloc4.p, ctx->d_counts, cs->mine, ctx->box,
                                        cs->peers, cs->sync.p + 5, epoch);
    MESO_CUDA(cudaGetLastError());
    return MESO_OK;
}

// a new set of atoms: every rank maps its neighbors again at the next rebuild (all ranks upload together)
void comm_invalidate(meso_ctx *ctx)
{
    if (ctx->comm && ctx->nccl) static_cast<CommState *>(ctx->comm)->ready = false;
}

// host-driven bootstrap (several ranks in one process, or on one GPU): see include/meso_b200.h
int comm_export_blob(meso_ctx *ctx, void *blob1024)
{
    int rc = ensure_arena(ctx);
    if (rc) return rc;
    CommBlob bl;
    fill_blob(ctx, bl);
    memset(blob1024, 0, 1024);
    memcpy(blob1024, &bl, sizeof bl);
    return MESO_OK;
}

int comm_import_blobs(meso_ctx *ctx, const void *blobs, int nranks)
{
    if (nranks != ctx->nranks) { ctx->err = "meso_comm_import: one blob per rank of the decomposition"; return MESO_EINVAL; }
    int rc = ensure_arena(ctx);
    if (rc) return rc;
    return import_blobs(ctx, static_cast<const unsigned char *>(blobs), nranks);
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
