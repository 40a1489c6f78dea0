// api.cu -- the extern "C" boundary (include/meso_b200.h) and the step orchestration.
//
// Orchestration follows ModifiedVerlet::setup / ::run (UM/mvv_meso.cu:139-219, 243-425) but
// never leaves the device: the reference's transfer_pre_exchange / pre_sort / pre_border /
// post_border / pre_comm / post_comm PCIe round trips (UM/atom_meso.cu:152-266) have no
// counterpart here; the host only enqueues kernels.
#include <nvtx3/nvToolsExt.h>
#include "internal.h"
#include "device_math.cuh"
#include <cuda_profiler_api.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <functional>

using namespace meso;

namespace meso {
int launch_deinterleave3(meso_ctx *ctx, const double *aos, DevBuf<double> *dst, int n);
int launch_interleave3(meso_ctx *ctx, DevBuf<double> *src, double *aos, int n);
int launch_fill_defaults(meso_ctx *ctx, int n, int st, int sy, int sm, int si);
int launch_zero3(meso_ctx *ctx, DevBuf<double> *a, int n);
int eval_gaussian(meso_ctx *ctx, int n, const uint32_t *si, const uint32_t *sj, float *osp, double *odp);
int eval_math(meso_ctx *ctx, int fn, int n, const double *a, const double *b, double *out);
int eval_log2u(meso_ctx *ctx, int n, const uint32_t *a, double *out);
int comm_init(meso_ctx *ctx, const void *nccl_id);
void comm_destroy(meso_ctx *ctx);
int comm_allreduce_sum(meso_ctx *ctx, double *host_vals, int n);

// gang.cu: several GPUs behind one handle
int gang_each(meso_ctx *ctx, const std::function<int(meso_ctx *, int)> &fn);
int gang_size(meso_ctx *ctx);
meso_ctx *gang_member(meso_ctx *ctx, int k);
int gang_create(meso_ctx **out, int ndev, const int *devices);
void gang_destroy(meso_ctx *ctx);
int gang_set_box(meso_ctx *ctx, const double boxlo[3], const double boxhi[3], const int periodic[3]);
int gang_atoms_upload(meso_ctx *ctx, int nlocal, const double *x, const double *v, const int *tag, const int *type, const int *mask, const int *image);
int gang_bonds_upload(meso_ctx *ctx, int nlocal, int bond_per_atom, const int *num_bond, const int *bond_type, const int *bond_atom, int tag_max);
int gang_ensure_peers(meso_ctx *ctx);
int gang_atoms_download(meso_ctx *ctx, int nmax, double *x, double *v, double *f, int *tag, int *type, int *mask, int *image);
int gang_counts(meso_ctx *ctx, int *nlocal, int *nghost, int *n_bulk, int *n_border);
int gang_sum(meso_ctx *ctx, int width, double *out, const std::function<int(meso_ctx *, double *)> &fn);
void gang_choose_grid(int n, const double prd[3], int grid[3]);
void gang_deal(const double boxlo[3], const double boxhi[3], const int periodic[3], const int grid[3], int n, const double *x, int *owner);
}

static std::string g_create_err;
namespace meso { std::string &create_error() { return g_create_err; } }
static int comm_setup_public(meso_ctx *ctx);

#define CHECK_CTX() do { if (!ctx) return MESO_EINVAL; } while (0)
#define FAIL(code, msg) do { ctx->err = (msg); return (code); } while (0)
#define TRY(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)
// a gang handle (gang.cu) forwards the call to every member, each on its own host thread (`c` = the member)
#define GANG_ALL(call) do { if (ctx->gang) return gang_each(ctx, [&](meso_ctx *c, int) -> int { return (call); }); } while (0)
#define GANG_STEP(call) do { if (ctx->gang) { TRY(gang_ensure_peers(ctx)); return gang_each(ctx, [&](meso_ctx *c, int) -> int { return (call); }); } } while (0)
#define GANG_FIRST(call) do { if (ctx->gang) { meso_ctx *c = gang_member(ctx, 0); int rc_ = (call); if (rc_ < 0) ctx->err = c->err; return rc_; } } while (0)
#define GANG_NONE(what) do { if (ctx->gang) FAIL(MESO_EINVAL, what ": not available on a multi-GPU handle (ask a member)"); } while (0)

// ---------------------------------------------------------------- timers
namespace {
struct PhaseTimer {
    meso_ctx *ctx; int id; cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get(meso_ctx *ctx)
    {
        if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    // every phase is also an NVTX range (the reference brackets its phases with nvtx_push_range/pop under PROFILE=1,
    // UM/nvtx_meso.h:9-22, UM/mvv_meso.cu:262-406); without a tool attached the calls return at once
    PhaseTimer(meso_ctx *c, int i) : ctx(c), id(i)
    {
        static const char *const names[MESO_T_COUNT] = {"integrate", "forward", "pair", "rebuild", "neigh"};
        nvtxRangePushA(names[i]);
        if (!ctx->timers_on) return;
        a = get(ctx); b = get(ctx);
        cudaEventRecord(a, ctx->stream);
    }
    ~PhaseTimer()
    {
        nvtxRangePop();
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        ctx->t_pending.push_back({id, {a, b}});
    }
};
void drain_timers(meso_ctx *ctx)
{
    for (auto &p : ctx->t_pending) {
        float ms = 0.f;
        cudaEventSynchronize(p.second.second);
        cudaEventElapsedTime(&ms, p.second.first, p.second.second);
        ctx->t_ms[p.first] += ms; ctx->t_calls[p.first]++;
        ctx->ev_pool.push_back(p.second.first); ctx->ev_pool.push_back(p.second.second);
    }
    ctx->t_pending.clear();
}
}  // namespace

// ---------------------------------------------------------------- device runtime
extern "C" int meso_device_count(void)
{
    // A process that is about to drive several GPUs (MESO_DEVICES, lammps/USER-MESO-B200/engine_meso.cpp) gets its kernels
    // loaded up front: with lazy loading the first launch of a kernel waits for the kernels already running in that context,
    // and a brick's halo kernel may be one of them, waiting for a neighbor.  This is the first CUDA call of such a process
    // (src/lammps.cpp:439-441 asks for the device count before it builds MesoDevice), so the setting still takes effect.
    if (getenv("MESO_DEVICES") && !getenv("CUDA_MODULE_LOADING")) setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int meso_create(meso_ctx **out, int device)
{
    if (!out) return MESO_EINVAL;
    *out = nullptr;
    int n = meso_device_count();
    if (n <= 0) { g_create_err = "no CUDA device visible: this library has no CPU fallback"; return MESO_ENODEV; }
    if (device < 0) device = 0;
    device %= n;                                        // local_rank % dev_count, src/lammps.cpp:451
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { g_create_err = "cudaGetDeviceProperties failed"; return MESO_ECUDA; }
    if (prop.major < 10) { g_create_err = std::string("device ") + prop.name + " is not sm_100-class; kernels are built for sm_100a only"; return MESO_ENODEV; }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_err = "cudaSetDevice failed"; return MESO_ECUDA; }
    meso_ctx *ctx = new meso_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    // the halo stream outranks the compute stream: its pack / NCCL / unpack CTAs are placed as soon as SM slots free up
    // instead of queueing behind the ~8000 CTAs of the bulk force kernel they are meant to overlap
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaMalloc(&ctx->d_counts, sizeof(Counts)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_counts, sizeof(Counts)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_result, 16 * sizeof(double)) != cudaSuccess) {
        g_create_err = std::string("context allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return MESO_ECUDA;
    }
    cudaEventCreateWithFlags(&ctx->ev_fwd_begin, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_fwd_end, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_counts, cudaEventDisableTiming);
    // MESO_FORCE_COMM_PATH=1 runs the message-based halo path (pack -> [NCCL] -> unpack, side-stream overlap) even on one rank
    const char *fc = getenv("MESO_FORCE_COMM_PATH");
    ctx->comm_path = fc && fc[0] == '1';
    const char *po = getenv("MESO_PAIR_ONCE");               // 0: two-sided force kernel in meso_run (A/B measurements)
    ctx->pair_once = !(po && po[0] == '0');
    if (const char *e = getenv("MESO_PAIR_TEX")) ctx->pair_tex = atoi(e) & 3;
    if (const char *e = getenv("MESO_NB_SLOW")) ctx->nb_slow = e[0] == '1';
    cudaMemsetAsync(ctx->d_counts, 0, sizeof(Counts), ctx->stream);
    memset(ctx->h_counts, 0, sizeof(Counts));
    *out = ctx;
    return MESO_OK;
}

extern "C" int meso_create_gang(meso_ctx **out, int ndev, const int *devices)
{
    if (!out || ndev < 1 || !devices) return MESO_EINVAL;
    if (ndev == 1) return meso_create(out, devices[0]);
    *out = nullptr;
    return gang_create(out, ndev, devices);
}

extern "C" int meso_gang_size(meso_ctx *ctx) { return ctx ? gang_size(ctx) : 0; }

// host-side rules of the gang, callable without a GPU (tests): the processor grid for ndev bricks of a box, and the brick
// (x-major rank) each position is dealt to
extern "C" int meso_gang_layout(int ndev, const double boxlo[3], const double boxhi[3], const int periodic[3], int grid[3], int n,
                                const double *x, int *owner)
{
    if (ndev < 1 || !boxlo || !boxhi || !periodic || !grid || n < 0 || (n > 0 && (!x || !owner))) return MESO_EINVAL;
    double prd[3];
    for (int d = 0; d < 3; d++) { prd[d] = boxhi[d] - boxlo[d]; if (!(prd[d] > 0)) return MESO_EINVAL; }
    gang_choose_grid(ndev, prd, grid);
    gang_deal(boxlo, boxhi, periodic, grid, n, x, owner);
    return MESO_OK;
}

extern "C" void meso_destroy(meso_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->gang) { gang_destroy(ctx); delete ctx; return; }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    release_retired_buffers();
    comm_destroy(ctx);
    drain_timers(ctx);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->tex_coord) cudaDestroyTextureObject(ctx->tex_coord);
    if (ctx->tex_veloc) cudaDestroyTextureObject(ctx->tex_veloc);
    if (ctx->ev_fwd_begin) cudaEventDestroy(ctx->ev_fwd_begin);
    if (ctx->ev_fwd_end) cudaEventDestroy(ctx->ev_fwd_end);
    if (ctx->ev_counts) cudaEventDestroy(ctx->ev_counts);
    if (ctx->d_counts) cudaFree(ctx->d_counts);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    delete ctx;
}

extern "C" const char *meso_last_error(meso_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

static int check_device_flags(meso_ctx *ctx)
{
    int e = ctx->h_counts->err;
    if (!e) return MESO_OK;
    ctx->err = "device-side capacity error:";
    if (e & 1) ctx->err += " ghost capacity exceeded;";
    if (e & 2) ctx->err += " pair table overflow (local density too high for n_col);";
    if (e & 8) ctx->err += " atom lost in migration (migration message capacity, or an atom beyond a non-periodic face);";
    if (e & 16) ctx->err += " a neighbor rank did not deliver its halo message in time;";
    if (e & 32) ctx->err += " Bond atoms missing (a bond partner is neither local nor ghost);";
    return MESO_ECAPACITY;
}

static int refresh_counts(meso_ctx *ctx)
{
    MESO_CUDA(cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, sizeof(Counts), cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return check_device_flags(ctx);
}

extern "C" int meso_sync(meso_ctx *ctx)
{
    CHECK_CTX();
    GANG_ALL(meso_sync(c));
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    drain_timers(ctx);
    return MESO_OK;
}

extern "C" void *meso_stream(meso_ctx *ctx) { return ctx ? (void *)(ctx->gang ? gang_member(ctx, 0)->stream : ctx->stream) : nullptr; }

extern "C" int meso_profiler(meso_ctx *ctx, int start)
{
    CHECK_CTX();
    GANG_ALL(meso_profiler(c, start));
    if (start) cudaProfilerStart(); else cudaProfilerStop();
    return MESO_OK;
}

extern "C" int meso_memory_usage(meso_ctx *ctx, uint64_t *bytes)
{
    CHECK_CTX();
    if (ctx->gang) {
        double o = 0.;
        TRY(gang_sum(ctx, 1, &o, [&](meso_ctx *c, double *part) -> int { uint64_t b = 0; int rc = meso_memory_usage(c, &b); *part = (double)b; return rc; }));
        if (bytes) *bytes = (uint64_t)o;
        return MESO_OK;
    }
    uint64_t b = 0;
    for (int d = 0; d < 3; d++) b += ctx->x[d].bytes() + ctx->v[d].bytes() + ctx->f[d].bytes() + ctx->xa[d].bytes() + ctx->va[d].bytes();
    b += ctx->tag.bytes() + ctx->type.bytes() + ctx->mask.bytes() + ctx->image.bytes() + ctx->taga.bytes() + ctx->typea.bytes() +
         ctx->maska.bytes() + ctx->imagea.bytes() + ctx->coord4.bytes() + ctx->veloc4.bytes() + ctx->virial.bytes() + ctx->e_pair.bytes() +
         ctx->staging.bytes() + ctx->istaging.bytes() + ctx->key.bytes() + ctx->perm_from.bytes() + ctx->sort.key_alt.bytes() +
         ctx->sort.val_alt.bytes() + ctx->sort.hist.bytes() + ctx->ghost_root.bytes() + ctx->ghost_shift.bytes() + ctx->tile_counts.bytes() +
         ctx->cell_of.bytes() + ctx->cell_atoms.bytes() + ctx->cell_start.bytes() + ctx->stencil.bytes() + ctx->slotrank.bytes() +
         ctx->cell_cnt.bytes() + ctx->cell_xyzj.bytes() + ctx->cell_soa.bytes() + ctx->scan_sums.bytes() + ctx->pos_of.bytes() +
         ctx->pair_count.bytes() + ctx->owned_count.bytes() + ctx->pair_table.bytes() + ctx->partial.bytes() +
         ctx->facc.bytes() + ctx->virial.bytes() + ctx->e_pair.bytes();
    *bytes = b;
    return MESO_OK;
}

// ---------------------------------------------------------------- domain
static void update_subbox(meso_ctx *ctx)
{
    Box &b = ctx->box;
    // Domain::set_local_box (uniform split): sublo = boxlo + prd * (loc/p)
    for (int d = 0; d < 3; d++) {
        int p = ctx->procgrid[d], me = ctx->myloc[d];
        double inv = 1.0 / p;
        b.sublo[d] = b.boxlo[d] + b.prd[d] * (me * inv);
        b.subhi[d] = (me < p - 1) ? b.boxlo[d] + b.prd[d] * ((me + 1) * inv) : b.boxhi[d];
        b.centre[d] = 0.5 * (b.subhi[d] + b.sublo[d]);           // UM/atom_vec_meso.cu:172-176
    }
    ctx->bins_ready = false;
}

extern "C" int meso_set_box(meso_ctx *ctx, const double boxlo[3], const double boxhi[3], const int periodic[3])
{
    CHECK_CTX();
    if (ctx->gang) return gang_set_box(ctx, boxlo, boxhi, periodic);
    for (int d = 0; d < 3; d++) {
        if (!(boxhi[d] > boxlo[d])) FAIL(MESO_EINVAL, "meso_set_box: boxhi must exceed boxlo");
        ctx->box.boxlo[d] = boxlo[d]; ctx->box.boxhi[d] = boxhi[d]; ctx->box.prd[d] = boxhi[d] - boxlo[d];
        ctx->box.periodic[d] = periodic[d] ? 1 : 0;
    }
    ctx->box_set = true;
    update_subbox(ctx);
    return MESO_OK;
}

// host-driven bootstrap of the halo (include/meso_b200.h)
extern "C" int meso_comm_blob_size(void) { return 1024; }
extern "C" int meso_comm_export(meso_ctx *ctx, void *blob)
{
    CHECK_CTX();
    GANG_NONE("meso_comm_export");
    if (!blob) FAIL(MESO_EINVAL, "meso_comm_export: null blob");
    if (!ctx->box_set) FAIL(MESO_EINVAL, "meso_comm_export: call after meso_set_box / meso_set_decomposition / meso_atoms_upload");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(comm_setup_public(ctx));
    return comm_export_blob(ctx, blob);
}
extern "C" int meso_comm_import(meso_ctx *ctx, const void *blobs, int nranks)
{
    CHECK_CTX();
    GANG_NONE("meso_comm_import");
    if (!blobs) FAIL(MESO_EINVAL, "meso_comm_import: null blobs");
    MESO_CUDA(cudaSetDevice(ctx->device));
    return comm_import_blobs(ctx, blobs, nranks);
}

extern "C" int meso_set_decomposition(meso_ctx *ctx, int rank, const int procgrid[3], const void *nccl_id)
{
    CHECK_CTX();
    if (ctx->gang) FAIL(MESO_EINVAL, "meso_set_decomposition: a multi-GPU handle decomposes the box itself");
    int n = procgrid[0] * procgrid[1] * procgrid[2];
    if (n < 1 || rank < 0 || rank >= n) FAIL(MESO_EINVAL, "meso_set_decomposition: bad rank/procgrid");
    ctx->rank = rank; ctx->nranks = n;
    for (int d = 0; d < 3; d++) ctx->procgrid[d] = procgrid[d];
    ctx->myloc[0] = rank / (procgrid[1] * procgrid[2]);
    ctx->myloc[1] = (rank / procgrid[2]) % procgrid[1];
    ctx->myloc[2] = rank % procgrid[2];
    for (int d = 0; d < 3; d++) {
        int l[3] = {ctx->myloc[0], ctx->myloc[1], ctx->myloc[2]}, u[3] = {ctx->myloc[0], ctx->myloc[1], ctx->myloc[2]};
        l[d] = (ctx->myloc[d] - 1 + procgrid[d]) % procgrid[d];
        u[d] = (ctx->myloc[d] + 1) % procgrid[d];
        ctx->procneigh[d][0] = (l[0] * procgrid[1] + l[1]) * procgrid[2] + l[2];
        ctx->procneigh[d][1] = (u[0] * procgrid[1] + u[1]) * procgrid[2] + u[2];
    }
    if (ctx->box_set) update_subbox(ctx);
    if (n > 1) {
        // with an ncclUniqueId the library bootstraps itself (NCCL carries the memory handles and the thermo reductions);
        // without one the host exchanges meso_comm_export blobs (ranks of one process, or of one GPU) and sums reductions itself
        if (nccl_id) TRY(comm_init(ctx, nccl_id));
        ctx->comm_path = true;
    }
    return MESO_OK;
}

// Comm::setup (src/comm.cpp:393-640) for style SINGLE, uniform bricks, maxneed <= 1
static int comm_setup(meso_ctx *ctx)
{
    Box &b = ctx->box;
    const double cutghost = ctx->cutneighmax;
    for (int d = 0; d < 3; d++) {
        int p = ctx->procgrid[d], me = ctx->myloc[d];
        int maxneed = (int)(cutghost * p / b.prd[d]) + 1;
        if (!b.periodic[d]) maxneed = std::min(maxneed, p - 1);
        if (maxneed > 1) FAIL(MESO_EINVAL, "sub-domain thinner than the ghost cutoff (maxneed > 1) is not supported");
        int sendneed[2];
        if (!b.periodic[d]) {
            int left = me - 1; if (left < 0) left = p - 1;
            sendneed[0] = std::min(maxneed, p - left - 1);
            int right = me + 1; if (right == p) right = 0;
            sendneed[1] = std::min(maxneed, right);
        } else sendneed[0] = sendneed[1] = maxneed;
        b.sendflag[2 * d] = (maxneed > 0 && sendneed[0] > 0) ? 1 : 0;
        b.sendflag[2 * d + 1] = (maxneed > 0 && sendneed[1] > 0) ? 1 : 0;
        b.slab_lo_hi[d] = b.sublo[d] + cutghost;
        b.slab_hi_lo[d] = b.subhi[d] - cutghost;
        b.pbc[2 * d] = (me == 0) ? 1 : 0;
        b.pbc[2 * d + 1] = (me == p - 1) ? -1 : 0;
    }
    return MESO_OK;
}

static int comm_setup_public(meso_ctx *ctx) { return comm_setup(ctx); }

// ---------------------------------------------------------------- settings
// Neighbor::init: cutneighmax = largest pair cutoff + skin.  Every setter that touches one of the three inputs goes through
// here, so the result does not depend on the order of the calls (the global cutoff stands in until the coefficients arrive).
static void update_cutneighmax(meso_ctx *ctx)
{
    ctx->cutneighmax = (ctx->max_pair_cut > 0.0 ? ctx->max_pair_cut : ctx->cut_global) + ctx->skin;
    ctx->bins_ready = false;
}

extern "C" int meso_set_neighbor(meso_ctx *ctx, double skin, int every)
{
    CHECK_CTX();
    GANG_ALL(meso_set_neighbor(c, skin, every));
    if (skin < 0 || every < 1) FAIL(MESO_EINVAL, "meso_set_neighbor: skin >= 0 and every >= 1 required");
    ctx->skin = skin; ctx->every = every;
    update_cutneighmax(ctx);
    return MESO_OK;
}

extern "C" int meso_set_types(meso_ctx *ctx, int ntypes, const double *mass)
{
    CHECK_CTX();
    GANG_ALL(meso_set_types(c, ntypes, mass));
    if (ntypes < 1 || !mass) FAIL(MESO_EINVAL, "meso_set_types: ntypes >= 1 and mass[ntypes+1] required");
    MESO_CUDA(cudaSetDevice(ctx->device));
    ctx->ntypes = ntypes;
    ctx->mass.assign(mass, mass + ntypes + 1);
    if (!ctx->mass_dev.reserve(ntypes + 1)) FAIL(MESO_ECUDA, "out of device memory");
    MESO_CUDA(cudaMemcpyAsync(ctx->mass_dev.p, ctx->mass.data(), sizeof(double) * (ntypes + 1), cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->coeff_ready = false;
    return MESO_OK;
}

extern "C" int meso_pair_dpd_settings(meso_ctx *ctx, int precision, double cut_global, int seed)
{
    CHECK_CTX();
    GANG_ALL(meso_pair_dpd_settings(c, precision, cut_global, seed));
    if (precision != MESO_SP && precision != MESO_DP) FAIL(MESO_EINVAL, "Illegal pair_style command");
    ctx->precision = precision; ctx->cut_global = cut_global; ctx->seed = seed;
    ctx->max_pair_cut = 0.0;                                // pair_style resets the coefficients (MesoPairDPD::settings)
    ctx->coeff_ready = false;
    update_cutneighmax(ctx);
    return MESO_OK;
}

extern "C" int meso_pair_dpd_coeff(meso_ctx *ctx, const double *coeff7)
{
    CHECK_CTX();
    GANG_ALL(meso_pair_dpd_coeff(c, coeff7));
    if (ctx->ntypes < 1) FAIL(MESO_EINVAL, "meso_pair_dpd_coeff: call meso_set_types first");
    if (!coeff7) FAIL(MESO_EINVAL, "All pair coeffs are not set");
    MESO_CUDA(cudaSetDevice(ctx->device));
    const int n = ctx->ntypes * ctx->ntypes * NCOEFF;
    ctx->coeff.assign(coeff7, coeff7 + n);
    double cmax = 0;
    for (int t = 0; t < ctx->ntypes * ctx->ntypes; t++) cmax = std::max(cmax, coeff7[t * NCOEFF + P_CUT]);
    ctx->max_pair_cut = std::max(cmax, 0.0);
    update_cutneighmax(ctx);
    std::vector<float> sp(n);
    for (int i = 0; i < n; i++) sp[i] = (float)coeff7[i];
    if (!ctx->coeff_sp.reserve(n) || !ctx->coeff_dp.reserve(n)) FAIL(MESO_ECUDA, "out of device memory");
    MESO_CUDA(cudaMemcpyAsync(ctx->coeff_sp.p, sp.data(), sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaMemcpyAsync(ctx->coeff_dp.p, ctx->coeff.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->coeff_ready = true;
    return MESO_OK;
}

extern "C" int meso_set_timestep_size(meso_ctx *ctx, double dt)
{
    CHECK_CTX();
    GANG_ALL(meso_set_timestep_size(c, dt));
    if (!(dt > 0)) FAIL(MESO_EINVAL, "timestep must be positive");
    ctx->dt = dt;
    return MESO_OK;
}
extern "C" int meso_set_force_units(meso_ctx *ctx, double ftm2v)
{
    CHECK_CTX();
    GANG_ALL(meso_set_force_units(c, ftm2v));
    if (!(ftm2v > 0)) FAIL(MESO_EINVAL, "ftm2v must be positive");
    ctx->ftm2v = ftm2v;
    return MESO_OK;
}
extern "C" int meso_set_reduce_scope(meso_ctx *ctx, int local_only)
{
    CHECK_CTX();
    if (ctx->gang) return MESO_OK;                          // the gang always returns whole-box sums
    ctx->reduce_local = local_only != 0;
    return MESO_OK;
}
// page-lock a host array in place so uploads/downloads run at PCIe rate (Pinned<T>, UM/memory_meso.h:224-243)
extern "C" int meso_host_register(meso_ctx *ctx, void *ptr, uint64_t bytes)
{
    CHECK_CTX();
    if (ctx->gang) return MESO_OK;                          // uploads are dealt out through per-member staging copies
    if (!ptr || !bytes) return MESO_OK;
    MESO_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return MESO_OK; }
    MESO_CUDA(e);
    return MESO_OK;
}
extern "C" int meso_host_unregister(meso_ctx *ctx, void *ptr)
{
    CHECK_CTX();
    if (ctx->gang) return MESO_OK;
    if (!ptr) return MESO_OK;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) cudaGetLastError();      // not registered: nothing to undo
    return MESO_OK;
}
extern "C" int meso_set_ntimestep(meso_ctx *ctx, int64_t t)
{
    CHECK_CTX();
    GANG_ALL(meso_set_ntimestep(c, t));
    ctx->ntimestep = t;
    return MESO_OK;
}
extern "C" int64_t meso_get_ntimestep(meso_ctx *ctx) { return ctx ? (ctx->gang ? gang_member(ctx, 0)->ntimestep : ctx->ntimestep) : -1; }

// ---------------------------------------------------------------- atom store
static int ensure_capacity(meso_ctx *ctx, size_t nlocal)
{
    // ghosts: shell of width cutghost around the brick at the brick's density, with head room
    const Box &b = ctx->box;
    double vin = 1, vout = 1;
    for (int d = 0; d < 3; d++) { double w = b.subhi[d] - b.sublo[d]; vin *= w; vout *= w + 2.0 * ctx->cutneighmax; }
    size_t nghost = (size_t)((double)nlocal * (vout / vin - 1.0) * 1.25) + 4096;
    size_t nloc_cap = (ctx->nranks > 1) ? nlocal + nlocal / 8 + 1024 : nlocal;
    size_t cap = nloc_cap + nghost;
    if (cap <= ctx->cap) {
        // the per-atom arrays are large enough, but the split between locals and ghosts may have moved: the pair table is
        // reserved from table_rows at the next rebuild (bins_ready is cleared by every upload)
        ctx->nloc_cap = std::max(ctx->nloc_cap, nloc_cap);
        ctx->table_rows = std::max(ctx->table_rows, ((ctx->nloc_cap + 31) / 32) * 32);
        return MESO_OK;
    }
    ctx->nloc_cap = std::max(ctx->nloc_cap, nloc_cap);
    bool ok = true;
    for (int d = 0; d < 3; d++)
        ok = ok && ctx->x[d].reserve(cap) && ctx->v[d].reserve(cap) && ctx->f[d].reserve(cap) && ctx->xa[d].reserve(cap) && ctx->va[d].reserve(cap);
    ok = ok && ctx->tag.reserve(cap) && ctx->type.reserve(cap) && ctx->mask.reserve(cap) && ctx->image.reserve(cap) &&
         ctx->taga.reserve(cap) && ctx->typea.reserve(cap) && ctx->maska.reserve(cap) && ctx->imagea.reserve(cap) &&
         ctx->coord4.reserve(cap + 1) && ctx->veloc4.reserve(cap + 1) && ctx->key.reserve(cap) && ctx->perm_from.reserve(cap) &&
         ctx->ghost_root.reserve(cap) && ctx->ghost_shift.reserve(cap) && ctx->cell_of.reserve(cap) &&
         ctx->cell_atoms.reserve(cap) && ctx->e_pair.reserve(cap) && ctx->pair_count.reserve(cap);
    if (!ok) FAIL(MESO_ECUDA, "out of device memory growing the atom store");
    ctx->cap = cap;
    {   // linear float4 textures over the packed views
        for (cudaTextureObject_t *t : {&ctx->tex_coord, &ctx->tex_veloc}) if (*t) { cudaDestroyTextureObject(*t); *t = 0; }
        cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
        cudaTextureDesc td; memset(&td, 0, sizeof td);
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        td.readMode = cudaReadModeElementType;
        rd.res.linear.devPtr = ctx->coord4.p; rd.res.linear.sizeInBytes = (cap + 1) * sizeof(float4);
        MESO_CUDA(cudaCreateTextureObject(&ctx->tex_coord, &rd, &td, nullptr));
        rd.res.linear.devPtr = ctx->veloc4.p;
        MESO_CUDA(cudaCreateTextureObject(&ctx->tex_veloc, &rd, &td, nullptr));
    }
    // the `far` record of the pair-once kernel: coord4[cap] = {3.4e38 (0x7f7f7f7f) x 3, type 0}
    MESO_CUDA(cudaMemsetAsync(ctx->coord4.p + cap, 0x7f, 3 * sizeof(float), ctx->stream));
    MESO_CUDA(cudaMemsetAsync(reinterpret_cast<float *>(ctx->coord4.p + cap) + 3, 0, sizeof(float), ctx->stream));   // .w = type 0
    MESO_CUDA(cudaMemsetAsync(ctx->veloc4.p + cap, 0, sizeof(float4), ctx->stream));
    if (!ctx->facc.reserve(cap)) FAIL(MESO_ECUDA, "out of device memory (force accumulator)");
    MESO_CUDA(cudaMemsetAsync(ctx->facc.p, 0, ctx->facc.bytes(), ctx->stream));
    if (!ctx->virial.reserve(6 * ctx->cap)) FAIL(MESO_ECUDA, "out of device memory (virial)");
    ctx->table_rows = ((ctx->nloc_cap + 31) / 32) * 32;
    return MESO_OK;
}

extern "C" int meso_atoms_upload(meso_ctx *ctx, int nlocal, const double *x, const double *v, const int *tag, const int *type,
                                 const int *mask, const int *image)
{
    CHECK_CTX();
    if (ctx->gang) return gang_atoms_upload(ctx, nlocal, x, v, tag, type, mask, image);
    if (!ctx->box_set) FAIL(MESO_EINVAL, "meso_atoms_upload: call meso_set_box first");
    if (nlocal < 0 || (nlocal > 0 && !x)) FAIL(MESO_EINVAL, "meso_atoms_upload: bad arguments");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(ensure_capacity(ctx, (size_t)nlocal));
    const size_t n = (size_t)nlocal;
    if (!ctx->staging.reserve(3 * n + 8)) FAIL(MESO_ECUDA, "out of device memory (staging)");
    MESO_CUDA(cudaMemcpyAsync(ctx->staging.p, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    TRY(launch_deinterleave3(ctx, ctx->staging.p, ctx->x, nlocal));
    if (v) {
        MESO_CUDA(cudaMemcpyAsync(ctx->staging.p, v, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
        TRY(launch_deinterleave3(ctx, ctx->staging.p, ctx->v, nlocal));
    } else TRY(launch_zero3(ctx, ctx->v, nlocal));
    TRY(launch_zero3(ctx, ctx->f, (int)ctx->cap));       // whole capacity: f doubles as the fp64 pair-once accumulator
    if (tag) MESO_CUDA(cudaMemcpyAsync(ctx->tag.p, tag, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (type) MESO_CUDA(cudaMemcpyAsync(ctx->type.p, type, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (mask) MESO_CUDA(cudaMemcpyAsync(ctx->mask.p, mask, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (image) MESO_CUDA(cudaMemcpyAsync(ctx->image.p, image, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));
    TRY(launch_fill_defaults(ctx, nlocal, !tag, !type, !mask, !image));
    Counts c;
    memset(&c, 0, sizeof c);
    c.nlocal = nlocal; c.n_bulk = nlocal; c.nall = nlocal;
    for (int s = 0; s < 6; s++) c.swap_first[s] = nlocal;
    *ctx->h_counts = c;
    MESO_CUDA(cudaMemcpyAsync(ctx->d_counts, ctx->h_counts, sizeof(Counts), cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->nlocal_host = nlocal;
    double tot = nlocal;
    if (ctx->nranks > 1) TRY(comm_allreduce_sum(ctx, &tot, 1));
    ctx->natoms_global = (int64_t)(tot + 0.5);
    ctx->bins_ready = false;
    ctx->setup_done = false;
    ctx->f_cleared = true;
    comm_invalidate(ctx);
    ctx->bond_per_atom = 0;                    // a bond table describes the atoms of ONE upload: re-send it with meso_bonds_upload
    return MESO_OK;
}

extern "C" int meso_atoms_download(meso_ctx *ctx, int nmax, double *x, double *v, double *f, int *tag, int *type, int *mask, int *image)
{
    CHECK_CTX();
    if (ctx->gang) return gang_atoms_download(ctx, nmax, x, v, f, tag, type, mask, image);
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    const int n = ctx->h_counts->nlocal;
    if (nmax < n) FAIL(MESO_EINVAL, "meso_atoms_download: buffer smaller than nlocal");
    if (!ctx->staging.reserve(3 * (size_t)n + 8)) FAIL(MESO_ECUDA, "out of device memory (staging)");
    DevBuf<double> *src[3] = {ctx->x, ctx->v, ctx->f};
    double *dst[3] = {x, v, f};
    for (int a = 0; a < 3; a++) {
        if (!dst[a]) continue;
        TRY(launch_interleave3(ctx, src[a], ctx->staging.p, n));
        MESO_CUDA(cudaMemcpyAsync(dst[a], ctx->staging.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
        MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (tag) MESO_CUDA(cudaMemcpyAsync(tag, ctx->tag.p, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (type) MESO_CUDA(cudaMemcpyAsync(type, ctx->type.p, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (mask) MESO_CUDA(cudaMemcpyAsync(mask, ctx->mask.p, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (image) MESO_CUDA(cudaMemcpyAsync(image, ctx->image.p, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return MESO_OK;
}

extern "C" int meso_counts(meso_ctx *ctx, int *nlocal, int *nghost, int *n_bulk, int *n_border)
{
    CHECK_CTX();
    if (ctx->gang) return gang_counts(ctx, nlocal, nghost, n_bulk, n_border);
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    if (nlocal) *nlocal = ctx->h_counts->nlocal;
    if (nghost) *nghost = ctx->h_counts->nghost;
    if (n_bulk) *n_bulk = ctx->h_counts->n_bulk;
    if (n_border) *n_border = ctx->h_counts->n_border;
    return MESO_OK;
}

extern "C" int64_t meso_natoms_global(meso_ctx *ctx) { return ctx ? (ctx->gang ? (int64_t)ctx->nlocal_host : ctx->natoms_global) : -1; }

// ---------------------------------------------------------------- phases
static int ready(meso_ctx *ctx)
{
    if (!ctx->box_set) FAIL(MESO_EINVAL, "box not set");
    if (!ctx->coeff_ready) FAIL(MESO_EINVAL, "All pair coeffs are not set");
    if (ctx->ntypes < 1) FAIL(MESO_EINVAL, "atom types not set");
    MESO_CUDA(cudaSetDevice(ctx->device));
    return MESO_OK;
}

extern "C" int meso_initial_integrate(meso_ctx *ctx, int groupbit)
{
    CHECK_CTX();
    GANG_STEP(meso_initial_integrate(c, groupbit));
    TRY(ready(ctx));
    PhaseTimer t(ctx, MESO_T_INTEGRATE);
    return launch_initial_integrate(ctx, groupbit, false);
}

extern "C" int meso_final_integrate(meso_ctx *ctx, int groupbit)
{
    CHECK_CTX();
    GANG_STEP(meso_final_integrate(c, groupbit));
    TRY(ready(ctx));
    PhaseTimer t(ctx, MESO_T_INTEGRATE);
    return launch_final_integrate(ctx, groupbit);
}

extern "C" int meso_neighbor_decide(meso_ctx *ctx)
{
    CHECK_CTX();
    GANG_ALL(meso_neighbor_decide(c));
    ctx->ago++;                                            // Neighbor::decide, src/neighbor.cpp:1216-1231 (delay 0, check no)
    return (ctx->ago % ctx->every == 0) ? 1 : 0;
}

static int rebuild_impl(meso_ctx *ctx)
{
    // a capacity error of an EARLIER rebuild truncates rows / drops ghosts: stop as soon as its Counts mirror has landed
    // (non-blocking query; the multi-rank path additionally waits once per rebuild in launch_forward_multi)
    if (ctx->setup_done && ctx->counts_pending && cudaEventQuery(ctx->ev_counts) == cudaSuccess) {
        ctx->counts_pending = false;
        TRY(check_device_flags(ctx));
    }
    if (!ctx->bins_ready) {
        TRY(comm_setup(ctx));
        TRY(launch_setup_bins(ctx));
        size_t need = ctx->table_rows * (size_t)ctx->n_col;
        if (!ctx->pair_table.reserve(need)) FAIL(MESO_ECUDA, "out of device memory (pair table)");
    }
    {
        PhaseTimer t(ctx, MESO_T_REBUILD);
        if (ctx->comm_path) {
            // Domain::pbc -> Comm::exchange -> sort_local -> Comm::borders (UM/mvv_meso.cu:283-316), all on device
            TRY(launch_pbc(ctx));
            TRY(launch_exchange_multi(ctx));
            TRY(launch_reorder(ctx));
            TRY(launch_bonds_gather(ctx));
            TRY(launch_borders_multi(ctx));
        } else {
            TRY(launch_reorder(ctx));        // the key kernel wraps as it goes
            TRY(launch_bonds_gather(ctx));
            TRY(launch_borders(ctx));
        }
        TRY(launch_bonds_map(ctx));          // map_set_device + bond_all, UM/mvv_meso.cu:316, UM/neighbor_meso.cu:136-159
    }
    {
        PhaseTimer t(ctx, MESO_T_NEIGH);
        TRY(launch_neighbor_build(ctx));
        TRY(launch_bonds_filter(ctx));       // filter_exclusion_meso, UM/neigh_build_meso.cu:546-569
    }
    MESO_CUDA(cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, sizeof(Counts), cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaEventRecord(ctx->ev_counts, ctx->stream));
    ctx->counts_pending = true;
    ctx->ago = 0;
    return MESO_OK;
}

extern "C" int meso_rebuild(meso_ctx *ctx)
{
    CHECK_CTX();
    GANG_STEP(meso_rebuild(c));
    TRY(ready(ctx));
    return rebuild_impl(ctx);
}

extern "C" int meso_forward_comm(meso_ctx *ctx)
{
    CHECK_CTX();
    GANG_STEP(meso_forward_comm(c));
    TRY(ready(ctx));
    PhaseTimer t(ctx, MESO_T_FORWARD);
    // phase API keeps the reference's order: ghosts' fp64 x,v are refreshed, packing happens in meso_pair_compute
    if (ctx->comm_path) {
        // the halo records carry the PACKED velocity + signature of the owner (comm.cu: 40-byte records), and in the phase
        // order initial_integrate -> forward_comm -> pair_compute the locals have not been repacked since the half-kick:
        // repack them first, so a ghost receives this step's velocity and signature (what its owner's pair kernel uses)
        TRY(launch_pack(ctx, MESO_LOCAL));
        return launch_forward_multi(ctx, ctx->stream);
    }
    return launch_forward(ctx, true);
}

extern "C" int meso_force_clear(meso_ctx *ctx, int range, int vflag)
{
    CHECK_CTX();
    GANG_STEP(meso_force_clear(c, range, vflag));
    TRY(ready(ctx));
    return launch_clear(ctx, range, vflag);
}

extern "C" int meso_pair_compute(meso_ctx *ctx, int range, int eflag, int vflag)
{
    CHECK_CTX();
    GANG_STEP(meso_pair_compute(c, range, eflag, vflag));
    TRY(ready(ctx));
    PhaseTimer t(ctx, MESO_T_PAIR);
    // compute_bulk packs LOCAL, compute_border packs GHOST, compute packs ALL (UM/pair_dpd_meso.cu:241-266)
    int pack_range = (range == MESO_BULK) ? MESO_LOCAL : (range == MESO_BORDER ? MESO_GHOST : MESO_ALL);
    TRY(launch_pack(ctx, pack_range));
    return launch_pair(ctx, range, eflag || vflag, true, false, 0);
}

extern "C" int meso_compute_ke(meso_ctx *ctx, int groupbit, double *mv2_sum, double *count)
{
    CHECK_CTX();
    if (ctx->gang) {
        double o[2];
        TRY(gang_sum(ctx, 2, o, [&](meso_ctx *c, double *part) -> int { return meso_compute_ke(c, groupbit, part, part + 1); }));
        if (mv2_sum) *mv2_sum = o[0];
        if (count) *count = o[1];
        return MESO_OK;
    }
    TRY(ready(ctx));
    double vals[2];
    TRY(launch_ke(ctx, groupbit, &vals[0], &vals[1]));
    if (ctx->nranks > 1 && !ctx->reduce_local) TRY(comm_allreduce_sum(ctx, vals, 2));
    if (mv2_sum) *mv2_sum = vals[0];
    if (count) *count = vals[1];
    return MESO_OK;
}

extern "C" int meso_compute_virial(meso_ctx *ctx, double virial6[6], double *e_pair)
{
    CHECK_CTX();
    if (ctx->gang) {
        double o[7];
        TRY(gang_sum(ctx, 7, o, [&](meso_ctx *c, double *part) -> int { return meso_compute_virial(c, part, part + 6); }));
        if (virial6) memcpy(virial6, o, 6 * sizeof(double));
        if (e_pair) *e_pair = o[6];
        return MESO_OK;
    }
    TRY(ready(ctx));
    double out[7];
    TRY(launch_virial_sum(ctx, out));
    if (ctx->nranks > 1 && !ctx->reduce_local) TRY(comm_allreduce_sum(ctx, out, 7));
    if (virial6) memcpy(virial6, out, 6 * sizeof(double));
    if (e_pair) *e_pair = out[6];
    return MESO_OK;
}

// ---------------------------------------------------------------- bonded topology
extern "C" int meso_bond_harmonic_coeff(meso_ctx *ctx, int nbondtypes, const double *k, const double *r0)
{
    CHECK_CTX();
    GANG_ALL(meso_bond_harmonic_coeff(c, nbondtypes, k, r0));
    if (nbondtypes < 1 || !k || !r0) FAIL(MESO_EINVAL, "Incorrect args for bond coefficients");
    MESO_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->bond_k_dev.reserve(nbondtypes + 1) || !ctx->bond_r0_dev.reserve(nbondtypes + 1)) FAIL(MESO_ECUDA, "out of device memory");
    MESO_CUDA(cudaMemcpyAsync(ctx->bond_k_dev.p, k, sizeof(double) * (nbondtypes + 1), cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaMemcpyAsync(ctx->bond_r0_dev.p, r0, sizeof(double) * (nbondtypes + 1), cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->nbondtypes = nbondtypes;
    return MESO_OK;
}

extern "C" int meso_set_special_bonds(meso_ctx *ctx, double lj12)
{
    CHECK_CTX();
    GANG_ALL(meso_set_special_bonds(c, lj12));
    if (lj12 != 0.0 && lj12 != 1.0) FAIL(MESO_EINVAL, "special_bonds: only lj weights 0 (1-2 pairs excluded) and 1 (kept) are supported");
    ctx->special_lj12 = lj12;
    return MESO_OK;
}

extern "C" int meso_bonds_upload(meso_ctx *ctx, int nlocal, int bond_per_atom, const int *num_bond, const int *bond_type,
                                 const int *bond_atom, int tag_max)
{
    CHECK_CTX();
    if (ctx->gang) return gang_bonds_upload(ctx, nlocal, bond_per_atom, num_bond, bond_type, bond_atom, tag_max);
    if (nlocal != ctx->nlocal_host) FAIL(MESO_EINVAL, "meso_bonds_upload: call right after meso_atoms_upload with the same atoms");
    if (bond_per_atom < 0 || (bond_per_atom > 0 && (!num_bond || !bond_type || !bond_atom))) FAIL(MESO_EINVAL, "meso_bonds_upload: bad arguments");
    MESO_CUDA(cudaSetDevice(ctx->device));
    ctx->bond_per_atom = bond_per_atom;
    ctx->map_tag_max = tag_max;
    if (bond_per_atom == 0) return MESO_OK;
    TRY(bonds_reserve(ctx));
    // LAMMPS rows [atom][slot] -> column-major [slot][atom] of {partner tag, type}
    std::vector<int2> col((size_t)ctx->cap * bond_per_atom, make_int2(0, 0));
    for (int i = 0; i < nlocal; i++) {
        if (num_bond[i] < 0 || num_bond[i] > bond_per_atom) FAIL(MESO_EINVAL, "meso_bonds_upload: num_bond exceeds bond_per_atom");
        for (int p = 0; p < num_bond[i]; p++)
            col[(size_t)p * ctx->cap + i] = make_int2(bond_atom[(size_t)i * bond_per_atom + p], bond_type[(size_t)i * bond_per_atom + p]);
    }
    MESO_CUDA(cudaMemcpyAsync(ctx->nbond.p, num_bond, sizeof(int) * nlocal, cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaMemcpyAsync(ctx->bonds.p, col.data(), sizeof(int2) * col.size(), cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->setup_done = false;
    return MESO_OK;
}

// force->bond->compute (UM/mvv_meso.cu:387-389): adds the bonded forces of the local atoms into f
extern "C" int meso_bond_compute(meso_ctx *ctx, int eflag, int vflag)
{
    CHECK_CTX();
    GANG_STEP(meso_bond_compute(c, eflag, vflag));
    TRY(ready(ctx));
    PhaseTimer t(ctx, MESO_T_PAIR);
    return launch_bond_force(ctx, eflag || vflag, false);
}

// sum of the per-atom bond energies of the last evflag evaluation (MesoBondHarmonic: e_bond = e/2 per atom)
extern "C" int meso_compute_bond_energy(meso_ctx *ctx, double *e_bond)
{
    CHECK_CTX();
    if (ctx->gang) {
        double o = 0.;
        TRY(gang_sum(ctx, 1, &o, [&](meso_ctx *c, double *part) -> int { return meso_compute_bond_energy(c, part); }));
        if (e_bond) *e_bond = o;
        return MESO_OK;
    }
    TRY(ready(ctx));
    double e = 0.;
    TRY(launch_bond_energy_sum(ctx, &e));
    if (ctx->nranks > 1 && !ctx->reduce_local) TRY(comm_allreduce_sum(ctx, &e, 1));   // collective: every rank calls
    if (e_bond) *e_bond = e;
    return MESO_OK;
}

// ---------------------------------------------------------------- device-resident fixes
static int fix_register(meso_ctx *ctx, const meso::FixOp &op)
{
    meso::FixList &fl = ctx->fixes;
    if (fl.n >= meso::MAX_FIX) FAIL(MESO_EINVAL, "too many device-resident fixes (limit 8)");
    fl.op[fl.n] = op;
    if (op.kind == meso::FIX_WALL || op.kind == meso::FIX_SOLID_BOUND) fl.nbounce++;
    if (op.kind == meso::FIX_RDF) fl.nrdf++; else fl.nforce++;
    return fl.n++;
}

extern "C" int meso_fix_wall(meso_ctx *ctx, int groupbit, int dims, double d, double f)
{
    CHECK_CTX();
    GANG_ALL(meso_fix_wall(c, groupbit, dims, d, f));
    if ((dims & 7) == 0) FAIL(MESO_EINVAL, "Incomplete fix wall command: insufficient arguments");
    meso::FixOp op{};
    op.kind = meso::FIX_WALL; op.groupbit = groupbit; op.dims = dims & 7;
    op.p[0] = d; op.p[1] = 1.0 / d; op.p[2] = f;                 // MesoFixWall::boundary_force passes d, 1.0 / d, f
    return fix_register(ctx, op);
}

extern "C" int meso_fix_solid_bound(meso_ctx *ctx, int groupbit, int dims, int force_kernel)
{
    CHECK_CTX();
    GANG_ALL(meso_fix_solid_bound(c, groupbit, dims, force_kernel));
    if ((dims & 7) == 0) FAIL(MESO_EINVAL, "Incomplete fix wall command: dimension unspecified");
    if (force_kernel != 1) FAIL(MESO_EINVAL, "Incomplete fix wall command: force kernel unspecified");
    meso::FixOp op{};
    op.kind = meso::FIX_SOLID_BOUND; op.groupbit = groupbit; op.dims = dims & 7; op.aux = force_kernel;
    return fix_register(ctx, op);
}

extern "C" int meso_fix_addforce(meso_ctx *ctx, int groupbit, double fx, double fy, double fz)
{
    CHECK_CTX();
    GANG_ALL(meso_fix_addforce(c, groupbit, fx, fy, fz));
    meso::FixOp op{};
    op.kind = meso::FIX_ADDFORCE; op.groupbit = groupbit;
    op.p[0] = fx; op.p[1] = fy; op.p[2] = fz;
    return fix_register(ctx, op);
}

extern "C" int meso_fix_pois(meso_ctx *ctx, int groupbit, int dim_ortho, int dim_force, double strength, double bisect_frac)
{
    CHECK_CTX();
    GANG_ALL(meso_fix_pois(c, groupbit, dim_ortho, dim_force, strength, bisect_frac));
    if (dim_ortho < 0 || dim_ortho > 2 || dim_force < 0 || dim_force > 2) FAIL(MESO_EINVAL, "Illegal fix CUDAPoiseuille command");
    meso::FixOp op{};
    op.kind = meso::FIX_POIS; op.groupbit = groupbit; op.dims = dim_ortho | (dim_force << 2);
    op.p[0] = strength; op.p[1] = bisect_frac;
    return fix_register(ctx, op);
}

// fix ID group rdf/fast/meso output <file> nbin <n> [every <k>] [other <group>]: samples in the post_force slot
extern "C" int meso_fix_rdf(meso_ctx *ctx, int groupbit, int j_groupbit, int nbin, int every)
{
    CHECK_CTX();
    GANG_ALL(meso_fix_rdf(c, groupbit, j_groupbit, nbin, every));
    if (nbin <= 0 || nbin > 8192) FAIL(MESO_EINVAL, "Incomplete compute rdf command: insufficient arguments");
    if (every < 1) every = 1;
    MESO_CUDA(cudaSetDevice(ctx->device));
    meso::FixOp op{};
    op.kind = meso::FIX_RDF; op.groupbit = groupbit; op.aux = j_groupbit; op.dims = nbin;
    op.p[0] = every; op.p[1] = ctx->cut_global;                  // MesoFixRDFFast::init: rc = force->pair->cutforce
    if (ctx->fixes.n >= meso::MAX_FIX) FAIL(MESO_EINVAL, "too many device-resident fixes (limit 8)");
    const int slot = ctx->fixes.n;                                // the handle fix_register is about to hand out
    if (!ctx->rdf_hist[slot].reserve((size_t)nbin + 2)) FAIL(MESO_ECUDA, "out of device memory (rdf histogram)");
    const int h = fix_register(ctx, op);
    if (h < 0) return h;
    MESO_CUDA(cudaMemsetAsync(ctx->rdf_hist[h].p, 0, sizeof(unsigned long long) * ((size_t)nbin + 2), ctx->stream));
    ctx->rdf_samples[h] = 0;
    return h;
}

// accumulated histogram[nbin] (pair counts, both directions), number of samples, sizes of the i and j groups; summed over
// the ranks unless meso_set_reduce_scope(1).  g(r) follows as in MesoFixRDFFast::dump (UM/fix_rdf_fast_meso.cu:182-219).
extern "C" int meso_fix_rdf_read(meso_ctx *ctx, int handle, int nbin, double *histogram, double *n_samples, double *ni, double *nj)
{
    CHECK_CTX();
    if (ctx->gang) {
        if (nbin <= 0 || !histogram) FAIL(MESO_EINVAL, "meso_fix_rdf_read: bad arguments");
        std::vector<double> o((size_t)nbin + 3);
        TRY(gang_sum(ctx, nbin + 3, o.data(), [&](meso_ctx *c, double *part) -> int {
            return meso_fix_rdf_read(c, handle, nbin, part, part + nbin, part + nbin + 1, part + nbin + 2); }));
        memcpy(histogram, o.data(), sizeof(double) * nbin);
        if (n_samples) *n_samples = o[nbin] / gang_size(ctx);      // every member sampled at the same steps
        if (ni) *ni = o[nbin + 1];
        if (nj) *nj = o[nbin + 2];
        return MESO_OK;
    }
    if (handle < 0 || handle >= ctx->fixes.n || ctx->fixes.op[handle].kind != meso::FIX_RDF || nbin != ctx->fixes.op[handle].dims || !histogram)
        FAIL(MESO_EINVAL, "meso_fix_rdf_read: not an rdf fix / wrong bin count");
    MESO_CUDA(cudaSetDevice(ctx->device));
    std::vector<double> buf((size_t)nbin + 2);
    TRY(fix_rdf_read(ctx, handle, nbin, buf.data(), &buf[nbin], &buf[nbin + 1]));
    if (ctx->nranks > 1 && !ctx->reduce_local) TRY(comm_allreduce_sum(ctx, buf.data(), nbin + 2));
    memcpy(histogram, buf.data(), sizeof(double) * nbin);
    if (n_samples) *n_samples = (double)ctx->rdf_samples[handle];
    if (ni) *ni = buf[nbin];
    if (nj) *nj = buf[nbin + 1];
    return MESO_OK;
}

extern "C" int meso_fix_clear(meso_ctx *ctx)
{
    CHECK_CTX();
    GANG_ALL(meso_fix_clear(c));
    ctx->fixes = meso::FixList{};
    return MESO_OK;
}

// modify->post_force for fix `handle` (< 0: all, in registration order): f += wall / body forces of the local atoms
extern "C" int meso_fix_post_force(meso_ctx *ctx, int handle)
{
    CHECK_CTX();
    GANG_STEP(meso_fix_post_force(c, handle));
    TRY(ready(ctx));
    if (handle >= ctx->fixes.n) FAIL(MESO_EINVAL, "meso_fix_post_force: no such fix");
    PhaseTimer t(ctx, MESO_T_INTEGRATE);
    if (handle >= 0 && ctx->fixes.op[handle].kind == meso::FIX_RDF) return launch_fix_rdf(ctx, handle);
    TRY(launch_fix_post_force(ctx, handle, false));
    return handle < 0 ? launch_fix_rdf(ctx, -1) : MESO_OK;
}

// the pre_exchange / end_of_step hook of wall/meso and solid_bound/meso: bounce-forward at the box faces
extern "C" int meso_fix_bounce(meso_ctx *ctx, int handle)
{
    CHECK_CTX();
    GANG_STEP(meso_fix_bounce(c, handle));
    TRY(ready(ctx));
    if (handle >= ctx->fixes.n) FAIL(MESO_EINVAL, "meso_fix_bounce: no such fix");
    PhaseTimer t(ctx, MESO_T_INTEGRATE);
    return launch_fix_bounce(ctx, handle);
}

// ---------------------------------------------------------------- whole-run drivers
extern "C" int meso_setup(meso_ctx *ctx, int eflag, int vflag)
{
    CHECK_CTX();
    GANG_STEP(meso_setup(c, eflag, vflag));
    TRY(ready(ctx));
    ctx->bins_ready = false;
    TRY(rebuild_impl(ctx));                                  // pbc, sort_local, borders, neighbor build (UM/mvv_meso.cu:150-185)
    {
        PhaseTimer t(ctx, MESO_T_PAIR);                      // force_clear + pair->compute (UM/mvv_meso.cu:191-197)
        TRY(launch_pair(ctx, MESO_LOCAL, eflag || vflag, false, false, 0));
        TRY(launch_bond_force(ctx, eflag || vflag, false));   // force->bond->compute, UM/mvv_meso.cu:198-201
        TRY(launch_fix_post_force(ctx, -1, false));           // modify->setup -> Fix::setup -> post_force, UM/mvv_meso.cu:212
    }
    ctx->setup_done = true;
    return refresh_counts(ctx);
}

// The run loop with every local pair evaluated once (pair.cu:k_dpd_once).  Forces are reduced into an accumulator
// (facc for dpd/fast/meso, f itself for dpd/meso) that is complete only when the whole force phase has finished, so the
// second half-kick moves out of the force kernel's epilogue into the next step's streaming pass (integrate.cu:
// k_step_integrate).  Kernels per ordinary step stay at three; on return f holds the last step's force in fp64 and the
// accumulator is zero again, so the phase entry points, downloads and reductions see the same state as before.
static int run_pair_once(meso_ctx *ctx, int nsteps, int groupbit)
{
    // accumulator: facc for the fp32 pair-once kernel; otherwise f itself (fp64 pair-once kernel, or the two-sided kernel in
    // accumulate mode when bonded forces force this loop with MESO_PAIR_ONCE=0)
    const bool acc_facc = ctx->precision == MESO_SP && ctx->pair_once;
    auto pair = [&](int range) -> int {
        if (ctx->pair_once) return launch_pair_once(ctx, range);
        return launch_pair(ctx, range, 0, true, false, 0);
    };
    bool pending = false;                                    // the previous step's second half-kick is still owed
    for (int s = 0; s < nsteps; s++) {
        ctx->ntimestep++;                                    // UM/mvv_meso.cu:256
        const bool rebuild = meso_neighbor_decide(ctx) != 0;
        {
            PhaseTimer t(ctx, MESO_T_INTEGRATE);
            // facc: the first step reads f (accumulator already zero), later steps read and clear facc;
            // f as accumulator: read and clear it every step
            TRY(launch_step_integrate(ctx, groupbit, pending, true, !rebuild, acc_facc && pending, acc_facc ? pending : true, false, rebuild));
        }
        pending = true;
        if (rebuild) {
            TRY(rebuild_impl(ctx));                          // gather + ghost kernels emit this step's packed views
        } else if (ctx->comm_path) {
            // halo refresh on the side stream, overlapped with the bulk force kernel (UM/mvv_meso.cu:338-376)
            MESO_CUDA(cudaEventRecord(ctx->ev_fwd_begin, ctx->stream));
            MESO_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fwd_begin, 0));
            TRY(launch_forward_multi(ctx, ctx->side));
            MESO_CUDA(cudaEventRecord(ctx->ev_fwd_end, ctx->side));
            {
                PhaseTimer t(ctx, MESO_T_PAIR);
                TRY(pair(MESO_BULK));
            }
            {
                PhaseTimer t(ctx, MESO_T_FORWARD);           // exposed (non-overlapped) part of the halo refresh
                MESO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_fwd_end, 0));
            }
            {
                PhaseTimer t(ctx, MESO_T_PAIR);
                TRY(pair(MESO_BORDER));
                TRY(launch_bond_force(ctx, 0, acc_facc));
                TRY(launch_fix_post_force(ctx, -1, acc_facc));
                TRY(launch_fix_rdf(ctx, -1));
            }
            continue;
        } else {
            PhaseTimer t(ctx, MESO_T_FORWARD);
            TRY(launch_forward(ctx, false));
        }
        PhaseTimer t(ctx, MESO_T_PAIR);
        TRY(pair(MESO_LOCAL));
        TRY(launch_bond_force(ctx, 0, acc_facc));            // bonded forces join the same accumulator (UM/mvv_meso.cu:387-389)
        TRY(launch_fix_post_force(ctx, -1, acc_facc));       // modify->post_force (UM/mvv_meso.cu:396)
        TRY(launch_fix_rdf(ctx, -1));
    }
    if (pending) {
        PhaseTimer t(ctx, MESO_T_INTEGRATE);                 // last step's second half-kick; f <- force, accumulator cleared
        TRY(launch_step_integrate(ctx, groupbit, true, false, false, acc_facc, acc_facc, acc_facc));
    }
    return MESO_OK;
}

extern "C" int meso_run(meso_ctx *ctx, int nsteps, int groupbit)
{
    CHECK_CTX();
    GANG_STEP(meso_run(c, nsteps, groupbit));
    TRY(ready(ctx));
    if (!ctx->setup_done) FAIL(MESO_EINVAL, "meso_run: call meso_setup first");
    // bonded forces and post_force fixes need the unfused second half-kick
    if (ctx->pair_once || bonds_active(ctx) || ctx->fixes.n > 0) return run_pair_once(ctx, nsteps, groupbit);
    for (int s = 0; s < nsteps; s++) {
        ctx->ntimestep++;                                    // UM/mvv_meso.cu:256
        const bool rebuild = meso_neighbor_decide(ctx) != 0;
        {
            PhaseTimer t(ctx, MESO_T_INTEGRATE);
            TRY(launch_initial_integrate(ctx, groupbit, !rebuild));
        }
        if (rebuild) {
            TRY(rebuild_impl(ctx));                          // gather + ghost kernels emit this step's packed views
        } else if (ctx->comm_path) {
            // halo refresh on the side stream, overlapped with the bulk force kernel (UM/mvv_meso.cu:338-376):
            // bulk particles have no ghost neighbors, border particles wait for the refreshed ghosts
            MESO_CUDA(cudaEventRecord(ctx->ev_fwd_begin, ctx->stream));
            MESO_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_fwd_begin, 0));
            TRY(launch_forward_multi(ctx, ctx->side));
            MESO_CUDA(cudaEventRecord(ctx->ev_fwd_end, ctx->side));
            {
                PhaseTimer t(ctx, MESO_T_PAIR);
                TRY(launch_pair(ctx, MESO_BULK, 0, false, true, groupbit));
            }
            {
                PhaseTimer t(ctx, MESO_T_FORWARD);           // exposed (non-overlapped) part of the halo refresh
                MESO_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_fwd_end, 0));
            }
            {
                PhaseTimer t(ctx, MESO_T_PAIR);
                TRY(launch_pair(ctx, MESO_BORDER, 0, false, true, groupbit));
            }
            continue;
        } else {
            PhaseTimer t(ctx, MESO_T_FORWARD);
            TRY(launch_forward(ctx, false));
        }
        PhaseTimer t(ctx, MESO_T_PAIR);                      // fused: clear + bulk + border + final_integrate
        TRY(launch_pair(ctx, MESO_LOCAL, 0, false, true, groupbit));
    }
    return MESO_OK;
}

// ---------------------------------------------------------------- exports
extern "C" int meso_export_bins(meso_ctx *ctx, int m[3], double binsize[3], double bininv[3], int *n_col)
{
    CHECK_CTX();
    GANG_FIRST(meso_export_bins(c, m, binsize, bininv, n_col));
    if (!ctx->bins_ready) FAIL(MESO_EINVAL, "bins not set up");
    for (int d = 0; d < 3; d++) { m[d] = ctx->box.m[d]; binsize[d] = ctx->box.binsize[d]; bininv[d] = ctx->box.bininv[d]; }
    if (n_col) *n_col = ctx->n_col;
    return MESO_OK;
}

template <typename T>
struct Tmp {
    T *p = nullptr;
    ~Tmp() { if (p) cudaFree(p); }
    bool alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * (n ? n : 1)) == cudaSuccess; }
};

template <typename T>
static int d2h(meso_ctx *ctx, T *dst, const T *src, size_t n)
{
    MESO_CUDA(cudaMemcpyAsync(dst, src, sizeof(T) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return MESO_OK;
}

extern "C" int meso_export_reorder(meso_ctx *ctx, int nmax, uint64_t *key_sorted, int *permute_from)
{
    CHECK_CTX();
    GANG_NONE("meso_export_reorder");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    int n = ctx->h_counts->nlocal;
    if (nmax < n) FAIL(MESO_EINVAL, "buffer too small");
    if (key_sorted) TRY(d2h(ctx, key_sorted, ctx->key.p, n));
    if (permute_from) TRY(d2h(ctx, permute_from, ctx->perm_from.p, n));
    return MESO_OK;
}

extern "C" int meso_export_packed(meso_ctx *ctx, int nmax, float *coord4, float *veloc4)
{
    CHECK_CTX();
    GANG_NONE("meso_export_packed");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    int n = ctx->h_counts->nlocal + ctx->h_counts->nghost;
    if (nmax < n) FAIL(MESO_EINVAL, "buffer too small");
    if (coord4) TRY(d2h(ctx, reinterpret_cast<float4 *>(coord4), ctx->coord4.p, n));
    if (veloc4) TRY(d2h(ctx, reinterpret_cast<float4 *>(veloc4), ctx->veloc4.p, n));
    return MESO_OK;
}

extern "C" int meso_export_ghosts(meso_ctx *ctx, int nmax, double *x, double *v, int *tag, int *type)
{
    CHECK_CTX();
    GANG_NONE("meso_export_ghosts");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    const int nl = ctx->h_counts->nlocal, ng = ctx->h_counts->nghost;
    if (nmax < ng) FAIL(MESO_EINVAL, "buffer too small");
    std::vector<double> tmp(ng > 0 ? ng : 1);
    for (int d = 0; d < 3; d++) {
        if (x) { TRY(d2h(ctx, tmp.data(), ctx->x[d].p + nl, ng)); for (int i = 0; i < ng; i++) x[3 * (size_t)i + d] = tmp[i]; }
        if (v) { TRY(d2h(ctx, tmp.data(), ctx->v[d].p + nl, ng)); for (int i = 0; i < ng; i++) v[3 * (size_t)i + d] = tmp[i]; }
    }
    if (tag) TRY(d2h(ctx, tag, ctx->tag.p + nl, ng));
    if (type) TRY(d2h(ctx, type, ctx->type.p + nl, ng));
    return MESO_OK;
}

extern "C" int meso_export_cells(meso_ctx *ctx, int ncell_plus1, int *cell_start, int nmax, int *cell_atoms)
{
    CHECK_CTX();
    GANG_NONE("meso_export_cells");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    int n = ctx->h_counts->nlocal + ctx->h_counts->nghost;
    if (cell_start) {
        if (ncell_plus1 < ctx->box.ncell + 1) FAIL(MESO_EINVAL, "buffer too small");
        TRY(d2h(ctx, cell_start, ctx->cell_start.p, ctx->box.ncell + 1));
    }
    if (cell_atoms) {
        if (nmax < n) FAIL(MESO_EINVAL, "buffer too small");
        TRY(d2h(ctx, cell_atoms, ctx->cell_atoms.p, n));
    }
    return MESO_OK;
}

extern "C" int meso_export_stencil(meso_ctx *ctx, int cell, int out27[27])
{
    CHECK_CTX();
    GANG_NONE("meso_export_stencil");
    MESO_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->bins_ready || cell < 0 || cell >= ctx->box.ncell) FAIL(MESO_EINVAL, "bad cell");
    unsigned char row[32];
    TRY(d2h(ctx, row, ctx->stencil.p + (size_t)cell * 32, 32));
    const Box &b = ctx->box;
    int n = row[31];
    for (int s = 0; s < n; s++) {
        int code = row[s];
        out27[s] = cell + (code % 3 - 1) + b.m[0] * ((code / 3) % 3 - 1 + b.m[1] * (code / 9 - 1));
    }
    return n;
}

extern "C" int meso_export_pair_count(meso_ctx *ctx, int nmax, int *pair_count)
{
    CHECK_CTX();
    GANG_NONE("meso_export_pair_count");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    int n = ctx->h_counts->nlocal;
    if (nmax < n) FAIL(MESO_EINVAL, "buffer too small");
    return d2h(ctx, pair_count, ctx->pair_count.p, n);
}

extern "C" int meso_export_pair_table(meso_ctx *ctx, int64_t nmax, int *pair_table)
{
    CHECK_CTX();
    GANG_NONE("meso_export_pair_table");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    size_t rows = ((size_t)ctx->h_counts->nlocal + 31) / 32 * 32;
    size_t n = rows * (size_t)ctx->n_col;
    if ((size_t)nmax < n) FAIL(MESO_EINVAL, "buffer too small");
    // rows leave in the reference's order (neighbor.cu:k_canonical_rows); the production layout is meso_export_pair_rows
    Tmp<int> scratch;
    if (!scratch.alloc(n)) FAIL(MESO_ECUDA, "out of device memory (canonical table)");
    MESO_CUDA(cudaMemsetAsync(scratch.p, 0, sizeof(int) * n, ctx->stream));
    TRY(launch_canonical_rows(ctx, scratch.p));
    return d2h(ctx, pair_table, scratch.p, n);
}

// the table as the force kernels read it: rows [owned][other] and the length of the owned part
extern "C" int meso_export_pair_rows(meso_ctx *ctx, int64_t nmax, int *pair_table, int *owned_count)
{
    CHECK_CTX();
    GANG_NONE("meso_export_pair_rows");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    const int nl = ctx->h_counts->nlocal;
    size_t n = ((size_t)nl + 31) / 32 * 32 * (size_t)ctx->n_col;
    if (pair_table) {
        if ((size_t)nmax < n) FAIL(MESO_EINVAL, "buffer too small");
        TRY(d2h(ctx, pair_table, ctx->pair_table.p, n));
    }
    if (owned_count) TRY(d2h(ctx, owned_count, ctx->owned_count.p, nl));
    return MESO_OK;
}

extern "C" int meso_export_virial(meso_ctx *ctx, int nmax, double *virial6, double *e_pair)
{
    CHECK_CTX();
    GANG_NONE("meso_export_virial");
    MESO_CUDA(cudaSetDevice(ctx->device));
    TRY(refresh_counts(ctx));
    const int n = ctx->h_counts->nlocal;
    if (nmax < n) FAIL(MESO_EINVAL, "buffer too small");
    if (virial6) {
        std::vector<double> tmp(n > 0 ? n : 1);
        for (int q = 0; q < 6; q++) {
            TRY(d2h(ctx, tmp.data(), ctx->virial.p + (size_t)q * ctx->cap, n));
            for (int i = 0; i < n; i++) virial6[6 * (size_t)i + q] = tmp[i];
        }
    }
    if (e_pair) TRY(d2h(ctx, e_pair, ctx->e_pair.p, n));
    return MESO_OK;
}

extern "C" int meso_eval_gaussian(meso_ctx *ctx, int n, const uint32_t *sig_i, const uint32_t *sig_j, float *out_sp, double *out_dp)
{
    CHECK_CTX();
    GANG_FIRST(meso_eval_gaussian(c, n, sig_i, sig_j, out_sp, out_dp));
    MESO_CUDA(cudaSetDevice(ctx->device));
    Tmp<uint32_t> a, b; Tmp<float> s; Tmp<double> d;
    if (!a.alloc(n) || !b.alloc(n) || !s.alloc(n) || !d.alloc(n)) FAIL(MESO_ECUDA, "out of device memory");
    MESO_CUDA(cudaMemcpyAsync(a.p, sig_i, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    MESO_CUDA(cudaMemcpyAsync(b.p, sig_j, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    TRY(eval_gaussian(ctx, n, a.p, b.p, out_sp ? s.p : nullptr, out_dp ? d.p : nullptr));
    if (out_sp) TRY(d2h(ctx, out_sp, s.p, n));
    if (out_dp) TRY(d2h(ctx, out_dp, d.p, n));
    MESO_CUDA(cudaStreamSynchronize(ctx->stream));
    return MESO_OK;
}

extern "C" int meso_eval_math(meso_ctx *ctx, int fn, int n, const double *a, const double *b, double *out)
{
    CHECK_CTX();
    GANG_FIRST(meso_eval_math(c, fn, n, a, b, out));
    MESO_CUDA(cudaSetDevice(ctx->device));
    if (fn < 0 || fn > 7 || (fn == 7 && !b)) FAIL(MESO_EINVAL, "meso_eval_math: bad function id");
    Tmp<double> da, db, dout;
    if (!da.alloc(n) || !db.alloc(n) || !dout.alloc(n)) FAIL(MESO_ECUDA, "out of device memory");
    MESO_CUDA(cudaMemcpyAsync(da.p, a, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (b) MESO_CUDA(cudaMemcpyAsync(db.p, b, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    TRY(eval_math(ctx, fn, n, da.p, db.p, dout.p));
    return d2h(ctx, out, dout.p, n);
}

extern "C" int meso_eval_log2u(meso_ctx *ctx, int n, const uint32_t *a, double *out)
{
    CHECK_CTX();
    GANG_FIRST(meso_eval_log2u(c, n, a, out));
    MESO_CUDA(cudaSetDevice(ctx->device));
    Tmp<uint32_t> da; Tmp<double> dout;
    if (!da.alloc(n) || !dout.alloc(n)) FAIL(MESO_ECUDA, "out of device memory");
    MESO_CUDA(cudaMemcpyAsync(da.p, a, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    TRY(eval_log2u(ctx, n, da.p, dout.p));
    return d2h(ctx, out, dout.p, n);
}

// ---------------------------------------------------------------- timers
extern "C" int meso_launch_count(meso_ctx *ctx, int64_t *n, int reset)
{
    CHECK_CTX();
    if (ctx->gang) {
        double o = 0.;
        TRY(gang_sum(ctx, 1, &o, [&](meso_ctx *c, double *part) -> int { int64_t v = 0; int rc = meso_launch_count(c, &v, reset); *part = (double)v; return rc; }));
        if (n) *n = (int64_t)o;
        return MESO_OK;
    }
    if (n) *n = ctx->n_launch;
    if (reset) ctx->n_launch = 0;
    return MESO_OK;
}

extern "C" int meso_timers_enable(meso_ctx *ctx, int on)
{
    CHECK_CTX();
    GANG_ALL(meso_timers_enable(c, on));
    ctx->timers_on = on != 0;
    return MESO_OK;
}

extern "C" int meso_timers_read(meso_ctx *ctx, double ms[MESO_T_COUNT], int64_t calls[MESO_T_COUNT], int reset)
{
    CHECK_CTX();
    if (ctx->gang) {                                          // slowest member per phase; call counts of the first
        const int nk = gang_size(ctx);
        std::vector<double> all((size_t)nk * MESO_T_COUNT, 0.0);
        std::vector<int64_t> cl((size_t)nk * MESO_T_COUNT, 0);
        TRY(gang_each(ctx, [&](meso_ctx *c, int k) -> int { return meso_timers_read(c, &all[(size_t)k * MESO_T_COUNT], &cl[(size_t)k * MESO_T_COUNT], reset); }));
        for (int q = 0; q < MESO_T_COUNT; q++) {
            double mx = 0.0;
            for (int k = 0; k < nk; k++) mx = std::max(mx, all[(size_t)k * MESO_T_COUNT + q]);
            if (ms) ms[q] = mx;
            if (calls) calls[q] = cl[q];
        }
        return MESO_OK;
    }
    MESO_CUDA(cudaSetDevice(ctx->device));
    drain_timers(ctx);
    for (int i = 0; i < MESO_T_COUNT; i++) {
        if (ms) ms[i] = ctx->t_ms[i];
        if (calls) calls[i] = ctx->t_calls[i];
        if (reset) { ctx->t_ms[i] = 0; ctx->t_calls[i] = 0; }
    }
    return MESO_OK;
}
