/*
 * sarashost_capi.cc -- a C handle over the C++ host objects (parser -> parallel -> grid -> multigrid_d2/d3 ->
 * plainsf) so that the Python tests and bench.py drive the SAME drop-in class a SARAS solver would:
 *     sarashost_mgsolve  ==  mgSolver.mgSolve(Pp, mgRHS)          (hydro_d3.cc:343)
 * One shared object per compile-time mode (PLANAR x TEST_POISSON), as in the reference's build.
 */
#include <cstring>
#include <iostream>
#include <string>

#include "poisson.h"
#include "sarasmg.h"

extern "C" {

/* same field order as the oracle's parameter block so tests can reuse one ctypes structure */
typedef struct {
    int dim, test_poisson;
    int xInd, yInd, zInd;
    int npX, npY;
    int xPer, yPer, zPer;
    int xGrid, yGrid, zGrid;
    double betaX, betaY, betaZ;
    double Lx, Ly, Lz;
    int vcDepth, vcCount, preSmooth, postSmooth, gsSmooth, resType;
    double mgTolerance;
    int wall_null;
} sarashost_params;

}

namespace {

#ifdef PLANAR
typedef multigrid_d2 mg_t;
#else
typedef multigrid_d3 mg_t;
#endif

struct Host {
    parser prm;
    parallel *mpi = nullptr;
    grid *mesh = nullptr;
    mg_t *mg = nullptr;
    plainsf *Pp = nullptr, *mgRHS = nullptr;
    Host() : prm(parser::no_file_t()) {}
    ~Host() { delete Pp; delete mgRHS; delete mg; delete mesh; delete mpi; }
};

thread_local std::string g_err;

}  // namespace

extern "C" {

int sarashost_mode(void) {
    int m = 0;
#ifdef PLANAR
    m |= 1;
#endif
#ifdef TEST_POISSON
    m |= 2;
#endif
    return m;
}

const char *sarashost_last_error(void) { return g_err.c_str(); }

void sarashost_set_process_rank(int rank, int size) { saras_mpi_set_process_rank(rank, size); }
void sarashost_set_hooks(void (*barrier)(void *), void (*allgather)(void *, const void *, void *, int), void *ctx) {
    saras_mpi_hooks h = {barrier, allgather, ctx};
    saras_mpi_set_hooks(&h);
}

void *sarashost_create(const sarashost_params *p, int print_residual) {
    const int planar = (sarashost_mode() & 1), test = (sarashost_mode() & 2) ? 1 : 0;
    if ((p->dim == 2) != (planar != 0) || (p->test_poisson != 0) != (test != 0)) {
        g_err = "sarashost_create: this library was built for a different PLANAR / TEST_POISSON mode";
        return nullptr;
    }
    Host *h = new Host();
    parser &q = h->prm;
    q.xInd = p->xInd; q.yInd = p->yInd; q.zInd = p->zInd;
    q.npX = p->npX; q.npY = p->npY;
    q.domainType = std::string() + (p->xPer ? 'P' : 'N') + (p->yPer ? 'P' : 'N') + (p->zPer ? 'P' : 'N');
    q.meshType = std::string() + (p->xGrid == 2 ? 'D' : 'U') + (p->yGrid == 2 ? 'D' : 'U') + (p->zGrid == 2 ? 'D' : 'U');
    q.betaX = p->betaX; q.betaY = p->betaY; q.betaZ = p->betaZ;
    q.Lx = p->Lx; q.Ly = p->Ly; q.Lz = p->Lz;
    q.vcDepth = p->vcDepth; q.vcCount = p->vcCount; q.preSmooth = p->preSmooth; q.postSmooth = p->postSmooth;
    q.gsSmooth = p->gsSmooth; q.resType = p->resType; q.mgTolerance = p->mgTolerance;
    q.printResidual = print_residual != 0;
    std::string why;
    if (!q.finalize(&why)) { g_err = why; delete h; return nullptr; }
    h->mpi = new parallel(q);
    h->mesh = new grid(q, *h->mpi);
    h->mg = new mg_t(*h->mesh, q);                      /* hydro_d3.cc:65: before checkPeriodic */
    if (p->wall_null) h->mpi->disablePeriodicAtWalls(q); /* hydro.cc:174-203 */
    h->Pp = new plainsf(*h->mesh);
    h->mgRHS = new plainsf(*h->mesh);
    return h;
}

void sarashost_destroy(void *hv) { delete (Host *)hv; }

int sarashost_extents(void *hv, int F[3]) {
    Host *h = (Host *)hv;
    for (int d = 0; d < 3; d++) F[d] = h->Pp->F.extent(d);
    return 0;
}

/* staggered coordinates of this rank along `dir` (length F[dir], index base -1) */
int sarashost_staggr(void *hv, int dir, double *out) {
    Host *h = (Host *)hv;
    const blitz::Array<real, 1> &a = (dir == 0) ? h->mesh->xStaggr : (dir == 1) ? h->mesh->yStaggr : h->mesh->zStaggr;
    for (int i = a.lbound(0); i <= a.ubound(0); i++) *out++ = a(i);
    return 0;
}

int sarashost_rank_info(void *hv, int out[6]) {
    Host *h = (Host *)hv;
    out[0] = h->mpi->xRank; out[1] = h->mpi->yRank;
    for (int f = 0; f < 4; f++) out[2 + f] = (h->mpi->faceRanks(f) == MPI_PROC_NULL) ? -1 : h->mpi->faceRanks(f);
    return 0;
}

/* mgSolver.mgSolve(Pp, mgRHS) on host arrays in the padded level-0 layout */
int sarashost_mgsolve(void *hv, double *p_inout, const double *rhs, int *cycles, double *hist, int *iters) {
    Host *h = (Host *)hv;
    const size_t n = h->Pp->F.size();
    memcpy(h->Pp->F.dataFirst(), p_inout, n * sizeof(double));
    memcpy(h->mgRHS->F.dataFirst(), rhs, n * sizeof(double));
    h->mg->mgSolve(*h->Pp, *h->mgRHS);
    memcpy(p_inout, h->Pp->F.dataFirst(), n * sizeof(double));
    if (cycles) *cycles = h->mg->lastCycles;
    for (int i = 0; i < h->mg->lastCycles; i++) {
        if (hist) hist[i] = h->mg->lastResiduals[(size_t)i];
        if (iters) iters[i] = h->mg->lastCoarseIters[(size_t)i];
    }
    return 0;
}

/* zero-copy variant: mgSolve directly on the plainsf storage (pointers returned by sarashost_field) */
double *sarashost_field(void *hv, int which) {
    Host *h = (Host *)hv;
    return which == 0 ? h->Pp->F.dataFirst() : h->mgRHS->F.dataFirst();
}
int sarashost_mgsolve_inplace(void *hv, int *cycles) {
    Host *h = (Host *)hv;
    h->mg->mgSolve(*h->Pp, *h->mgRHS);
    if (cycles) *cycles = h->mg->lastCycles;
    return 0;
}

/* host-side setup arrays of the drop-in, for CPU-only checks of the setup logic.
 * which: 0..8 = hx hy hz hx2 hz2 hzhx hxhy hyhz hxhyhz (vcDepth+1 doubles), 10..15 = xixx xix2 etyy ety2 ztzz ztz2 (levels
 * concatenated), 20..22 = xWall yWall zWall.  Returns the number of doubles written (out may be NULL to query it). */
int sarashost_setup_array(void *hv, int which, double *out) {
    Host *h = (Host *)hv;
    struct Peek : public mg_t {
        using mg_t::hx; using mg_t::hy; using mg_t::hz; using mg_t::hx2; using mg_t::hz2; using mg_t::hzhx;
        using mg_t::hxhy; using mg_t::hyhz; using mg_t::hxhyhz;
        using mg_t::xixx; using mg_t::xix2; using mg_t::etyy; using mg_t::ety2; using mg_t::ztzz; using mg_t::ztz2;
        using mg_t::xWall; using mg_t::yWall; using mg_t::zWall; using mg_t::maxCount;
    };
    Peek *q = static_cast<Peek *>(h->mg);
    if (which >= 0 && which <= 8) {
        blitz::Array<real, 1> *a[9] = {&q->hx, &q->hy, &q->hz, &q->hx2, &q->hz2, &q->hzhx, &q->hxhy, &q->hyhz, &q->hxhyhz};
        const int n = (int)a[which]->size();
        if (out) for (int i = 0; i < n; i++) out[i] = (*a[which])(i);
        return n;
    }
    std::vector<double> *v = nullptr;
    switch (which) {
        case 10: v = &q->xixx; break; case 11: v = &q->xix2; break; case 12: v = &q->etyy; break;
        case 13: v = &q->ety2; break; case 14: v = &q->ztzz; break; case 15: v = &q->ztz2; break;
        case 20: v = &q->xWall; break; case 21: v = &q->yWall; break; case 22: v = &q->zWall; break;
        case 30: if (out) out[0] = q->maxCount; return 1;
        default: return -1;
    }
    if (out) memcpy(out, v->data(), v->size() * sizeof(double));
    return (int)v->size();
}

void *sarashost_native(void *hv) { return ((Host *)hv)->mg->nativeHandle(); }
double sarashost_test_prolong(void *hv) { return ((Host *)hv)->mg->testProlong(); }
double sarashost_test_transfer(void *hv) { return ((Host *)hv)->mg->testTransfer(); }
double sarashost_test_periodic(void *hv) { return ((Host *)hv)->mg->testPeriodic(); }

}
