// gang.cu -- several GPUs behind ONE context handle: one process, one host thread and one meso_ctx per GPU.
//
// Replaces, for a host that is a single process (LAMMPS built against src/STUBS, a Python driver), the reference's
// "one MPI rank per GPU" process model (src/lammps.cpp:432-452 picks the device of a rank, src/comm.cpp:393-640 and
// UM/comm_meso.cu:41-254 move atoms between ranks through the host).  A gang handle is a meso_ctx without device state of its
// own: meso_set_box decomposes the box into uniform bricks (Comm::set_procs' surface rule, src/comm.cpp:201-287), every entry
// point of include/meso_b200.h fans out to the members on their own host threads (the multi-rank halo kernels wait for their
// neighbors' flags, so the members of a step must be issued concurrently), uploads are dealt out by brick, downloads are
// concatenated in member order and reductions are summed here.  The members find each other's halo arenas through plain
// device pointers (peer access inside one process, comm.cu:import_blobs); no NCCL communicator is created.
#include "internal.h"
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

namespace meso {

static std::mutex g_retired_mu;
static std::vector<void *> g_retired;
void retire_device_buffer(void *p) { std::lock_guard<std::mutex> lk(g_retired_mu); g_retired.push_back(p); }
void release_retired_buffers()
{
    std::vector<void *> v;
    { std::lock_guard<std::mutex> lk(g_retired_mu); v.swap(g_retired); }
    for (void *p : v) cudaFree(p);
}

struct Gang {
    std::vector<meso_ctx *> kids;
    int procgrid[3] = {1, 1, 1};
    double boxlo[3] = {0, 0, 0}, boxhi[3] = {1, 1, 1};
    int periodic[3] = {1, 1, 1};
    bool box_set = false;
    bool peers_stale = true;              // an upload re-sizes the halo arenas: exchange the members' blobs again before stepping
    std::vector<int> owner;               // member of every atom of the last upload (the bond table follows the same deal)
    std::vector<int> kid_n;               // atoms dealt to every member at the last upload
    // worker pool: thread k runs the current job on member k
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    const std::function<int(meso_ctx *, int)> *job = nullptr;
    uint64_t gen = 0;
    int pending = 0;
    bool quit = false;
    std::vector<int> rc;
};

static Gang *gang_of(meso_ctx *ctx) { return static_cast<Gang *>(ctx->gang); }

static void worker(Gang *g, int k)
{
    cudaSetDevice(g->kids[k]->device);
    uint64_t seen = 0;
    for (;;) {
        const std::function<int(meso_ctx *, int)> *job;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_go.wait(lk, [&] { return g->quit || g->gen != seen; });
            if (g->quit) return;
            seen = g->gen;
            job = g->job;
        }
        const int rc = (*job)(g->kids[k], k);
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->rc[k] = rc;
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// runs fn(member, index) on every member concurrently; first failure wins (its message is copied to the gang handle)
int gang_each(meso_ctx *ctx, const std::function<int(meso_ctx *, int)> &fn)
{
    Gang *g = gang_of(ctx);
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->job = &fn;
        g->pending = (int)g->kids.size();
        g->gen++;
        g->cv_go.notify_all();
        g->cv_done.wait(lk, [&] { return g->pending == 0; });
        g->job = nullptr;
    }
    for (size_t k = 0; k < g->kids.size(); k++)
        if (g->rc[k] < 0) {
            ctx->err = "gpu " + std::to_string(g->kids[k]->device) + " (brick " + std::to_string(k) + "): " + g->kids[k]->err;
            return g->rc[k];
        }
    return g->rc[0];
}

int gang_size(meso_ctx *ctx) { return ctx->gang ? (int)gang_of(ctx)->kids.size() : 1; }
meso_ctx *gang_member(meso_ctx *ctx, int k) { return gang_of(ctx)->kids[k]; }

// Comm::set_procs (src/comm.cpp:201-287): the factorisation with the smallest brick surface, ties to the first found
static void choose_grid(int n, const double prd[3], int grid[3])
{
    double best = 1e300;
    for (int px = 1; px <= n; px++) {
        if (n % px) continue;
        for (int py = 1; py <= n / px; py++) {
            if ((n / px) % py) continue;
            const int pz = n / px / py;
            const double surf = prd[0] * prd[1] / (px * py) + prd[0] * prd[2] / (px * pz) + prd[1] * prd[2] / (py * pz);
            if (surf < best) { best = surf; grid[0] = px; grid[1] = py; grid[2] = pz; }
        }
    }
}

void gang_choose_grid(int n, const double prd[3], int grid[3]) { choose_grid(n, prd, grid); }

// brick (x-major rank) of a position for a given box and grid: what gang_atoms_upload deals by
static int brick_of_box(const double boxlo[3], const double boxhi[3], const int periodic[3], const int grid[3], const double *x);

extern std::string &create_error();

int gang_create(meso_ctx **out, int ndev, const int *devices)
{
    // One brick per device.  Bricks that share a device share one CUDA context, whose kernels are not time-sliced against each
    // other the way the processes of tests/test_multi_gpu.py are: a halo kernel waiting for its neighbor's flag can starve the
    // neighbor (streams beyond the hardware queues alias, a lazily loaded kernel waits for the running ones, cudaFree waits for
    // every stream).  With eager module loading, 32 hardware queues and the stream-ordered DevBuf such gangs run -- mostly:
    // 2-8 bricks on one B200 finished or timed out from run to run (round 2) -- so the mode stays behind a switch.
    int most = 0;
    for (int k = 0; k < ndev; k++) {
        int same = 0;
        for (int q = 0; q < ndev; q++) same += devices[q] == devices[k];
        most = std::max(most, same);
    }
    if (most > 1) {
        const char *sh = getenv("MESO_GANG_SHARE_DEVICE"), *ml = getenv("CUDA_MODULE_LOADING"), *mc = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
        if (!(sh && sh[0] == '1')) {
            create_error() = "meso_create_gang: a device is listed more than once; one brick per device (MESO_GANG_SHARE_DEVICE=1 lifts the "
                             "check for experiments: bricks of one CUDA context can starve each other)";
            return MESO_EINVAL;
        }
        if (!(ml && strcmp(ml, "EAGER") == 0) || (mc ? atoi(mc) : 8) < 2 * most + 2) {
            create_error() = "meso_create_gang: bricks that share a device need CUDA_MODULE_LOADING=EAGER and CUDA_DEVICE_MAX_CONNECTIONS >= " +
                             std::to_string(2 * most + 2) + " (both read at the first CUDA call)";
            return MESO_EINVAL;
        }
    }
    Gang *g = new Gang();
    for (int k = 0; k < ndev; k++) {
        meso_ctx *kid = nullptr;
        const int rc = meso_create(&kid, devices[k]);
        if (rc != MESO_OK) {
            for (meso_ctx *c : g->kids) meso_destroy(c);
            delete g;
            return rc;
        }
        g->kids.push_back(kid);
    }
    g->rc.assign(ndev, 0);
    g->kid_n.assign(ndev, 0);
    meso_ctx *ctx = new meso_ctx();
    ctx->device = devices[0];
    ctx->gang = g;
    for (int k = 0; k < ndev; k++) g->threads.emplace_back(worker, g, k);
    *out = ctx;
    return MESO_OK;
}

void gang_destroy(meso_ctx *ctx)
{
    Gang *g = gang_of(ctx);
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
        g->cv_go.notify_all();
    }
    for (auto &t : g->threads) t.join();
    for (meso_ctx *c : g->kids) meso_destroy(c);
    delete g;
    ctx->gang = nullptr;
}

int gang_set_box(meso_ctx *ctx, const double boxlo[3], const double boxhi[3], const int periodic[3])
{
    Gang *g = gang_of(ctx);
    double prd[3];
    for (int d = 0; d < 3; d++) {
        g->boxlo[d] = boxlo[d]; g->boxhi[d] = boxhi[d]; g->periodic[d] = periodic[d] ? 1 : 0;
        prd[d] = boxhi[d] - boxlo[d];
        if (!(prd[d] > 0)) { ctx->err = "meso_set_box: boxhi must exceed boxlo"; return MESO_EINVAL; }
    }
    choose_grid((int)g->kids.size(), prd, g->procgrid);
    g->box_set = true;
    g->peers_stale = true;
    return gang_each(ctx, [&](meso_ctx *c, int k) -> int {
        int rc = meso_set_box(c, boxlo, boxhi, periodic);
        if (rc) return rc;
        rc = meso_set_decomposition(c, k, g->procgrid, nullptr);
        if (rc) return rc;
        return meso_set_reduce_scope(c, 1);                      // the gang sums the members' parts
    });
}

// brick of a position: Domain::set_local_box's uniform split, the position wrapped into a periodic box first (the member
// wraps and migrates for real at its first rebuild) and clamped into the outer bricks of a non-periodic dimension
static int brick_of_box(const double boxlo[3], const double boxhi[3], const int periodic[3], const int grid[3], const double *x)
{
    int loc[3];
    for (int d = 0; d < 3; d++) {
        const double prd = boxhi[d] - boxlo[d];
        double u = (x[d] - boxlo[d]) / prd;
        if (periodic[d]) u -= floor(u);
        int l = (int)(u * grid[d]);
        // the member's own test is sublo <= x < subhi with sublo = boxlo + prd * (l / p): settle rounding on that definition
        const int p = grid[d];
        l = std::min(std::max(l, 0), p - 1);
        const double xw = periodic[d] ? boxlo[d] + u * prd : x[d];
        while (l > 0 && xw < boxlo[d] + prd * (l * (1.0 / p))) l--;
        while (l < p - 1 && xw >= boxlo[d] + prd * ((l + 1) * (1.0 / p))) l++;
        loc[d] = l;
    }
    return (loc[0] * grid[1] + loc[1]) * grid[2] + loc[2];
}

static int brick_of(const Gang *g, const double *x) { return brick_of_box(g->boxlo, g->boxhi, g->periodic, g->procgrid, x); }

void gang_deal(const double boxlo[3], const double boxhi[3], const int periodic[3], const int grid[3], int n, const double *x, int *owner)
{
    for (int i = 0; i < n; i++) owner[i] = brick_of_box(boxlo, boxhi, periodic, grid, x + 3 * (size_t)i);
}

int gang_atoms_upload(meso_ctx *ctx, int nlocal, const double *x, const double *v, const int *tag, const int *type, const int *mask,
                      const int *image)
{
    Gang *g = gang_of(ctx);
    if (!g->box_set) { ctx->err = "meso_atoms_upload: call meso_set_box first"; return MESO_EINVAL; }
    if (nlocal < 0 || (nlocal > 0 && !x)) { ctx->err = "meso_atoms_upload: bad arguments"; return MESO_EINVAL; }
    const int nk = (int)g->kids.size();
    g->owner.resize((size_t)nlocal);
    std::vector<std::vector<int>> idx((size_t)nk);
    for (int i = 0; i < nlocal; i++) {
        const int k = brick_of(g, x + 3 * (size_t)i);
        g->owner[i] = k;
        idx[k].push_back(i);
    }
    for (int k = 0; k < nk; k++) g->kid_n[k] = (int)idx[k].size();
    g->peers_stale = true;
    ctx->nlocal_host = nlocal;
    return gang_each(ctx, [&](meso_ctx *c, int k) -> int {
        const std::vector<int> &id = idx[k];
        const size_t n = id.size();
        std::vector<double> xs(3 * n + 3), vs(v ? 3 * n + 3 : 0);
        std::vector<int> tg(n + 1), ty(type ? n + 1 : 0), mk(mask ? n + 1 : 0), im(image ? n + 1 : 0);
        for (size_t q = 0; q < n; q++) {
            const size_t i = (size_t)id[q];
            for (int d = 0; d < 3; d++) { xs[3 * q + d] = x[3 * i + d]; if (v) vs[3 * q + d] = v[3 * i + d]; }
            tg[q] = tag ? tag[i] : (int)i + 1;                   // the single-context default (index + 1) is global here
            if (type) ty[q] = type[i];
            if (mask) mk[q] = mask[i];
            if (image) im[q] = image[i];
        }
        return meso_atoms_upload(c, (int)n, xs.data(), v ? vs.data() : nullptr, tg.data(), type ? ty.data() : nullptr,
                                 mask ? mk.data() : nullptr, image ? im.data() : nullptr);
    });
}

int gang_bonds_upload(meso_ctx *ctx, int nlocal, int bond_per_atom, const int *num_bond, const int *bond_type, const int *bond_atom,
                      int tag_max)
{
    Gang *g = gang_of(ctx);
    if (nlocal != (int)g->owner.size()) { ctx->err = "meso_bonds_upload: call right after meso_atoms_upload with the same atoms"; return MESO_EINVAL; }
    return gang_each(ctx, [&](meso_ctx *c, int k) -> int {
        const size_t bpa = (size_t)std::max(bond_per_atom, 0);
        std::vector<int> nb, bt, ba;
        nb.reserve((size_t)g->kid_n[k] + 1); bt.reserve((size_t)g->kid_n[k] * bpa + 1); ba.reserve((size_t)g->kid_n[k] * bpa + 1);
        if (bond_per_atom > 0)
            for (int i = 0; i < nlocal; i++)
                if (g->owner[i] == k) {
                    nb.push_back(num_bond[i]);
                    for (size_t p = 0; p < bpa; p++) { bt.push_back(bond_type[(size_t)i * bpa + p]); ba.push_back(bond_atom[(size_t)i * bpa + p]); }
                }
        nb.push_back(0); bt.push_back(0); ba.push_back(0);
        return meso_bonds_upload(c, g->kid_n[k], bond_per_atom, nb.data(), bt.data(), ba.data(), tag_max);
    });
}

// the members' halo arenas exist once atoms and coefficients are in place: exchange their descriptions (host-driven
// bootstrap of include/meso_b200.h) before the first rebuild after an upload
int gang_ensure_peers(meso_ctx *ctx)
{
    Gang *g = gang_of(ctx);
    if (!g->peers_stale) return MESO_OK;
    const int nk = (int)g->kids.size();
    if (nk > 1) {
        std::vector<unsigned char> blobs((size_t)nk * 1024, 0);
        int rc = gang_each(ctx, [&](meso_ctx *c, int k) -> int { return meso_comm_export(c, blobs.data() + (size_t)k * 1024); });
        if (rc) return rc;
        rc = gang_each(ctx, [&](meso_ctx *c, int) -> int { return meso_comm_import(c, blobs.data(), nk); });
        if (rc) return rc;
    }
    g->peers_stale = false;
    return MESO_OK;
}

int gang_atoms_download(meso_ctx *ctx, int nmax, double *x, double *v, double *f, int *tag, int *type, int *mask, int *image)
{
    Gang *g = gang_of(ctx);
    const int nk = (int)g->kids.size();
    std::vector<int> n((size_t)nk, 0), off((size_t)nk + 1, 0);
    int rc = gang_each(ctx, [&](meso_ctx *c, int k) -> int { return meso_counts(c, &n[k], nullptr, nullptr, nullptr); });
    if (rc) return rc;
    for (int k = 0; k < nk; k++) off[k + 1] = off[k] + n[k];
    if (nmax < off[nk]) { ctx->err = "meso_atoms_download: buffer smaller than nlocal"; return MESO_EINVAL; }
    return gang_each(ctx, [&](meso_ctx *c, int k) -> int {
        const size_t o = (size_t)off[k];
        return meso_atoms_download(c, n[k], x ? x + 3 * o : nullptr, v ? v + 3 * o : nullptr, f ? f + 3 * o : nullptr, tag ? tag + o : nullptr,
                                   type ? type + o : nullptr, mask ? mask + o : nullptr, image ? image + o : nullptr);
    });
}

int gang_counts(meso_ctx *ctx, int *nlocal, int *nghost, int *n_bulk, int *n_border)
{
    const int nk = gang_size(ctx);
    std::vector<int> a((size_t)nk * 4, 0);
    int rc = gang_each(ctx, [&](meso_ctx *c, int k) -> int { return meso_counts(c, &a[4 * k], &a[4 * k + 1], &a[4 * k + 2], &a[4 * k + 3]); });
    if (rc) return rc;
    int s[4] = {0, 0, 0, 0};
    for (int k = 0; k < nk; k++) for (int q = 0; q < 4; q++) s[q] += a[4 * k + q];
    if (nlocal) *nlocal = s[0];
    if (nghost) *nghost = s[1];
    if (n_bulk) *n_bulk = s[2];
    if (n_border) *n_border = s[3];
    return MESO_OK;
}

// fn fills `width` doubles per member; out = sum over the members
int gang_sum(meso_ctx *ctx, int width, double *out, const std::function<int(meso_ctx *, double *)> &fn)
{
    const int nk = gang_size(ctx);
    std::vector<double> part((size_t)nk * width, 0.0);
    int rc = gang_each(ctx, [&](meso_ctx *c, int k) -> int { return fn(c, part.data() + (size_t)k * width); });
    if (rc) return rc;
    for (int q = 0; q < width; q++) {
        double s = 0.0;
        for (int k = 0; k < nk; k++) s += part[(size_t)k * width + q];
        out[q] = s;
    }
    return MESO_OK;
}

}  // namespace meso
