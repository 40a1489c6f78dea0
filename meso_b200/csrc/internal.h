// internal.h -- context and device-side layout shared by the kernels.
//
// HBM layout (one context = one GPU = one rank):
//   x[3], v[3]  fp64 SoA, capacity = locals + ghosts   (A1: MesoAtomVec fields, UM/atom_vec_meso.h:133-163)
//   f[3]        fp64 SoA, locals
//   tag,type,mask,image int32
//   coord4      float4 {x-cx, y-cy, z-cz, bits(type-1)}  \ the per-step packed view the
//   veloc4      float4 {vx, vy, vz, bits(signature)}     / force kernel gathers from (A7)
//   pair_table  int32, tile-transposed [ceil32(n)][n_col] (A6, UM/neigh_list_meso.cu:97-102)
// All per-rebuild counts live in a device-side Counts struct so the step loop never
// needs a host round trip; a pinned mirror is refreshed after each rebuild.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <functional>
#include <string>
#include <vector>
#include "../../include/meso_b200.h"

namespace meso {

struct Counts {
    int nlocal, nghost, n_bulk, n_border;
    int swap_first[6], swap_n[6];   // ghost index range created by each of the 6 border swaps
    int err;                        // bit 0: ghost capacity, 1: pair table overflow, 2: join overlap, 3: lost atom
    int max_pair;
    int nall;                       // nlocal + nghost
    int send_n[6];                  // atoms this rank sends in each border swap (== swap_n on a self-partnered swap)
    int exch_n[2];                  // migration: leavers to the lower / upper neighbor in the current dimension
    int err_any;                    // max of `err` over all ranks at the last rebuild (multi-rank: every rank fails together)
    // halo (comm.cu), per direction code: records this rank sends in / receives from that direction at every ghost creation
    // and refresh, and where the received block starts in the ghost array
    int route_send_n[27], route_recv_n[27];
    int slot_base[27];
};

struct Box {
    double boxlo[3], boxhi[3], prd[3];
    double sublo[3], subhi[3], centre[3];
    int periodic[3];
    int m[3];                       // cells per dim incl. ghost layer (mbinx..)
    double binsize[3], bininv[3];
    double slab_lo_hi[3];           // send slab: x <= sublo + cutghost   (swap 2d)
    double slab_hi_lo[3];           // send slab: x >= subhi - cutghost   (swap 2d+1)
    int sendflag[6];
    int pbc[6];                     // -1/0/+1 shift of swap s along its own dim
    int ncell;
};

// Growing a buffer must not synchronise the device: cudaMalloc / cudaFree wait for every kernel of the context, and with
// several bricks in one process (gang.cu) one of those kernels may be a neighbor's halo kernel waiting for THIS brick.
// Allocation is stream-ordered (cudaMallocAsync on the calling thread's stream, completed before it is handed out); a buffer
// that has been outgrown is parked and released when a context is destroyed (kernels in flight may still read it).
void retire_device_buffer(void *p);
void release_retired_buffers();

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    // returns false on allocation failure; contents are NOT preserved unless keep
    bool reserve(size_t n, bool keep = false, cudaStream_t s = 0)
    {
        if (n <= cap) return true;
        size_t ncap = n + n / 4 + 256;
        T *q = nullptr;
        if (cudaMallocAsync(&q, ncap * sizeof(T), cudaStreamPerThread) != cudaSuccess ||
            cudaStreamSynchronize(cudaStreamPerThread) != cudaSuccess) { cudaGetLastError(); return false; }
        if (keep && p && cap) { cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s); cudaStreamSynchronize(s); }
        if (p) retire_device_buffer(p);
        p = q; cap = ncap;
        return true;
    }
    size_t bytes() const { return cap * sizeof(T); }
};

struct SortScratch {
    DevBuf<uint64_t> key_alt;
    DevBuf<int> val_alt;
    DevBuf<uint32_t> hist;     // [256][ntiles]
    DevBuf<int> bucket_cnt, bucket_start, slot_idx;   // bucket sort of the reorder keys (sort.cu)
};

// device-resident fixes of the channel decks: wall/meso, solid_bound/meso, addforce/meso, pois/meso.
// The list is small and travels to the kernels by value (constant bank), so one streaming pass applies all of them.
enum { FIX_WALL = 1, FIX_SOLID_BOUND = 2, FIX_ADDFORCE = 3, FIX_POIS = 4, FIX_RDF = 5, MAX_FIX = 8 };
struct FixOp {
    int kind, groupbit;
    int dims;                       // wall / solid_bound: bit d = walls across dimension d;  pois: dim_ortho | dim_force << 2
    int aux;                        // solid_bound: force kernel id (1 = rho5rc1s1);  rdf: group bit of the j atoms
                                    // rdf: dims = nbin, p = {every, rc}
    double p[4];                    // wall: {d, 1/d, f};  addforce: {fx, fy, fz};  pois: {strength, bisect_frac}
};
struct FixList {
    int n, nbounce, nforce, nrdf;
    FixOp op[MAX_FIX];
};

enum { NCOEFF = 7, P_CUT = 0, P_CUTSQ, P_CUTINV, P_EXPW, P_A0, P_GAMMA, P_SIGMA };

}  // namespace meso

struct meso_ctx {
    int device = 0;
    void *gang = nullptr;          // != nullptr: this handle stands for several GPUs (gang.cu) and owns no device state itself
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;
    std::string err;

    // domain / decomposition
    meso::Box box{};
    bool box_set = false, bins_ready = false;
    int rank = 0, nranks = 1, procgrid[3] = {1, 1, 1}, myloc[3] = {0, 0, 0}, procneigh[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    void *nccl = nullptr;          // ncclComm_t

    // settings
    double skin = 0.3, cut_global = 1.0, cutneighmax = 1.3, dt = 0.005;
    double max_pair_cut = 0.0;     // largest per-pair cutoff of the coefficient table (0: not set yet)
    double ftm2v = 1.0;            // force->ftm2v of the host's unit system (1 in lj)
    bool reduce_local = false;     // bead-spring topology (bond.cu): column-major [slot][atom] table of {partner tag, bond type}
    int bond_per_atom = 0, nbondtypes = 0, map_tag_max = 0;
    double special_lj12 = 1.0;                // special_bonds lj weight of 1-2 pairs: 1 keeps them in the list, 0 filters them out
    meso::DevBuf<int> nbond, nbond_alt;
    meso::DevBuf<int2> bonds, bonds_alt, bonds_mapped;
    meso::DevBuf<unsigned> tag_map;
    meso::DevBuf<double> bond_k_dev, bond_r0_dev, e_bond;
    // reductions return this rank's part only (an MPI host sums them itself)
    int every = 5, ago = 0;
    int64_t ntimestep = 0;
    int ntypes = 0;
    std::vector<double> mass, coeff;
    int precision = MESO_SP;
    int seed = 0;
    bool coeff_ready = false;
    int n_col = 0;
    double expected_neigh_count = 0;

    // atom store
    int nlocal_host = 0;           // host knowledge of nlocal (exact on 1 rank; refreshed after migration)
    int64_t natoms_global = 0;
    size_t cap = 0;                // capacity of per-atom arrays (locals + ghosts)
    meso::DevBuf<double> x[3], v[3], f[3], xa[3], va[3];
    meso::DevBuf<int> tag, type, mask, image, taga, typea, maska, imagea;
    meso::DevBuf<float4> coord4, veloc4;
    meso::DevBuf<float4> facc;                // fp32 per-atom force accumulator of the pair-once kernel (zero between uses)
    cudaTextureObject_t tex_coord = 0, tex_veloc = 0;   // linear float4 textures over coord4 / veloc4 (gather-path experiments)
    int pair_tex = 2;                         // which gathers of the pair-once kernel use the texture data pipe (MESO_PAIR_TEX)
    bool nb_slow = false;                     // MESO_NB_SLOW=1: every row by the plain 27-cell walk (A/B check of the tile build)
    bool pair_once = true;                    // meso_run evaluates each local pair once (MESO_PAIR_ONCE=0: two-sided kernel)
    meso::DevBuf<double> virial, e_pair;      // [6][cap] SoA, [cap]
    meso::DevBuf<double> mass_dev;            // [ntypes+1]
    meso::DevBuf<float> coeff_sp;
    meso::DevBuf<double> coeff_dp;
    meso::DevBuf<double> staging;             // upload/download AoS staging
    meso::DevBuf<int> istaging;

    // reorder
    meso::DevBuf<uint64_t> key;
    meso::DevBuf<int> perm_from;
    meso::SortScratch sort;
    // ghosts
    meso::DevBuf<int> ghost_root;             // per ghost g: local source index
    meso::DevBuf<int> ghost_shift;            // per ghost g: packed shift code (2 bits per dim)
    meso::DevBuf<int> tile_counts;            // compaction scratch
    // multi-rank halo (comm.cu): arenas in peer memory, flags, send lists -- all behind an opaque state object
    bool comm_path = false;                   // use the message-based border/forward path (always when nranks > 1)
    void *comm = nullptr;                     // meso::CommState
    meso::DevBuf<void *> route_ptrs;          // device copy of the send-list pointer table
    size_t nloc_cap = 0;                      // capacity for local atoms
    meso::DevBuf<double> reduce_buf;
    cudaEvent_t ev_fwd_begin = nullptr, ev_fwd_end = nullptr;
    cudaEvent_t ev_counts = nullptr;          // the pinned Counts mirror of the last rebuild has landed
    bool counts_pending = false;              // ev_counts was recorded and its error flags have not been looked at yet
    // cells
    meso::DevBuf<int> cell_of;                // per atom: cell coordinates packed 10 bits per dimension (x | y << 10 | z << 20)
    meso::DevBuf<int> cell_atoms, cell_start; // atoms in (cell, ascending index) order; first position of every cell (+ total)
    meso::DevBuf<int> cell_cnt, scan_sums;    // histogram of the counting sort and the block sums of its scan
    meso::DevBuf<float4> cell_xyzj;           // cell-ordered records {x, y, z, bits(atom index)}: fall-back build, exports
    meso::DevBuf<float> cell_soa;             // the same as four arrays x | y | z | index: what the tile build copies (TMA)
    meso::DevBuf<unsigned char> stencil;      // [ncell][32]: stencil codes in the reference's order, byte 31 = count
    meso::DevBuf<unsigned char> slotrank;     // [ncell][32]: stencil code -> position in that order
    meso::DevBuf<int> pos_of;                 // per atom: its position in the cell order (index into cell_xyzj / cell_atoms)
    // neighbor table: row = [owned][other]
    meso::DevBuf<int> pair_count, pair_table;
    meso::DevBuf<int> owned_count;            // entries of the row whose pair this row evaluates (pair-once force kernel)
    meso::DevBuf<int> nb_fixup;               // != 0: some row was left to the fall-back build kernel
    bool nb_smem_optin = false;               // the build kernel may use more than 48 KB of shared memory on this device
    size_t table_rows = 0;

    // reductions
    meso::DevBuf<double> partial;
    double *h_result = nullptr;               // pinned, 16 doubles
    meso::Counts *d_counts = nullptr;
    meso::Counts *h_counts = nullptr;         // pinned mirror
    bool f_cleared = true, v_cleared = true;
    bool setup_done = false;

    // device-resident fixes (fix.cu)
    meso::FixList fixes{};
    meso::DevBuf<unsigned long long> rdf_hist[meso::MAX_FIX];   // rdf/fast/meso: pair counts per radial bin, accumulated over samples
    int64_t rdf_samples[meso::MAX_FIX] = {0};

    // every kernel launch of this library passes its stream through LS(): a real count for bench.py's gpu_launches
    int64_t n_launch = 0;
    // timers
    bool timers_on = false;
    double t_ms[MESO_T_COUNT] = {0};
    int64_t t_calls[MESO_T_COUNT] = {0};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> t_pending;
    std::vector<cudaEvent_t> ev_pool;
};

namespace meso {

#define LS(s) (++ctx->n_launch, (s))      // launch-stream wrapper: k<<<grid, block, smem, LS(stream)>>>(...)

#define MESO_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            return MESO_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

// kernels are persistent grid-stride: one launch shape for any n (counts are device-side)
inline int grid_for(const meso_ctx *ctx, int blocks_per_sm) { return ctx->sm_count * blocks_per_sm; }
// host-side upper bound of the local atom count (exact on one rank; the local capacity when atoms can migrate)
inline size_t nlocal_bound(const meso_ctx *ctx) { return ctx->nranks > 1 ? ctx->nloc_cap : (size_t)ctx->nlocal_host; }

// ---- sort.cu
// ---- reorder.cu
int launch_pbc(meso_ctx *ctx);         // periodic wrap + image flags only (before migration)
int launch_reorder(meso_ctx *ctx);     // pbc + key + sort + gather(+pack)
int launch_borders(meso_ctx *ctx);     // ghost creation (single rank: periodic images)
int launch_forward(meso_ctx *ctx, bool full);     // per-step ghost refresh
// ---- comm.cu
int launch_exchange_multi(meso_ctx *ctx);                   // migration: every leaver straight to its new brick
int launch_borders_multi(meso_ctx *ctx);                    // ghost creation: all 26 images at once, reference order
int launch_forward_multi(meso_ctx *ctx, cudaStream_t st);   // per-step ghost refresh
void comm_invalidate(meso_ctx *ctx);
int comm_export_blob(meso_ctx *ctx, void *blob1024);
int comm_import_blobs(meso_ctx *ctx, const void *blobs, int nranks);
// ---- neighbor.cu
int scan_into(meso_ctx *ctx, const int *in, int *out, int n);   // exclusive scan: out[0] = 0, out[i + 1] = sum of in[0..i]
int launch_setup_bins(meso_ctx *ctx);
int launch_neighbor_build(meso_ctx *ctx);
int launch_canonical_rows(meso_ctx *ctx, int *out_table);   // the table in the reference's row order (exports)
// ---- pair.cu
int launch_pack(meso_ctx *ctx, int range);
int launch_pair(meso_ctx *ctx, int range, int evflag, bool accumulate, bool fuse_final, int groupbit);
int launch_pair_once(meso_ctx *ctx, int range);
// ---- bond.cu
bool bonds_active(const meso_ctx *ctx);
int bonds_reserve(meso_ctx *ctx);
int launch_bonds_gather(meso_ctx *ctx);                 // after launch_reorder
int launch_bonds_map(meso_ctx *ctx);                    // after the ghosts exist
int launch_bonds_filter(meso_ctx *ctx);                 // after the neighbor build
int launch_bond_force(meso_ctx *ctx, int evflag, bool into_facc);
int launch_bond_energy_sum(meso_ctx *ctx, double *e);
// ---- fix.cu
int launch_fix_post_force(meso_ctx *ctx, int handle, bool into_facc);   // handle < 0: every registered fix, in order
int launch_fix_bounce(meso_ctx *ctx, int handle);
int launch_fix_rdf(meso_ctx *ctx, int handle);                          // samples every rdf fix whose cadence hits ctx->ntimestep
int fix_rdf_read(meso_ctx *ctx, int handle, int nbin, double *hist, double *ni, double *nj);
// ---- integrate.cu
int launch_initial_integrate(meso_ctx *ctx, int groupbit, bool pack);
int launch_final_integrate(meso_ctx *ctx, int groupbit);
// fused step boundary: [second half-kick of the previous step] + [first half-kick + drift (+ pack) of this step];
// force source = facc (fp32 accumulator) or f; optionally mirrors the force into f and clears the source
// bounce_after: wall fixes reflect once more after the drift (their pre_exchange hook on a re-neighbouring step)
int launch_step_integrate(meso_ctx *ctx, int groupbit, bool do_final, bool do_initial, bool pack, bool src_acc, bool zero_src, bool write_f,
                          bool bounce_after = false);
int launch_ke(meso_ctx *ctx, int groupbit, double *mv2, double *count);
int launch_virial_sum(meso_ctx *ctx, double out7[7]);
int launch_clear(meso_ctx *ctx, int range, int vflag);

uint32_t seed_now(const meso_ctx *ctx);

}  // namespace meso
