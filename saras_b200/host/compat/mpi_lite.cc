/* mpi_lite.cc -- implementation of compat/mpi.h (see there).  Test/build scaffolding, not a general MPI. */
#include "mpi.h"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int n = 1, count = 0;
    unsigned long gen = 0;
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        unsigned long g = gen;
        if (++count == n) { count = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

struct Comm {
    std::vector<int> members;                 /* world ranks, index = rank in this communicator */
    Barrier bar;
    std::vector<std::vector<char>> slots;     /* allgather staging, one per member */
};

struct DType { int count, blocklen, stride, elem; };   /* `count` blocks of `blocklen` elems, stride in elems */

struct Posted { void *buf; int count; MPI_Datatype type; int source; int tag; MPI_Comm comm; bool done; bool active; };

struct World {
    int size = 1;
    bool threaded = false;
    std::mutex m;
    std::vector<std::unique_ptr<Comm>> comms;
    std::vector<DType> types;
    std::map<int, std::vector<long>> subarrays;   /* MPI_Type_create_subarray: element offsets of the selected block, C order */
    std::vector<std::vector<Posted>> posted;  /* per world rank */
} W;

thread_local int t_rank = 0;
bool g_proc_mode = false;
int g_proc_rank = 0, g_proc_size = 1;
saras_mpi_hooks g_hooks = {nullptr, nullptr, nullptr};
bool g_init = false;

void world_init(int n, bool threaded) {
    W.size = n;
    W.threaded = threaded;
    W.comms.clear();
    W.comms.reserve(1024);
    W.types.clear();
    W.types.reserve(65536);
    W.subarrays.clear();
    W.posted.assign(n, {});
    auto c = std::make_unique<Comm>();
    for (int r = 0; r < n; r++) c->members.push_back(r);
    c->bar.n = n;
    c->slots.resize(n);
    W.comms.push_back(std::move(c));
    /* predefined types occupy handles 0..7 */
    W.types.resize(8);
    W.types[MPI_BYTE] = {1, 1, 1, 1};
    W.types[MPI_INT] = {1, 1, 1, (int)sizeof(int)};
    W.types[MPI_FLOAT] = {1, 1, 1, (int)sizeof(float)};
    W.types[MPI_DOUBLE] = {1, 1, 1, (int)sizeof(double)};
    W.types[MPI_LONG] = {1, 1, 1, (int)sizeof(long)};
    W.types[MPI_UNSIGNED] = {1, 1, 1, (int)sizeof(unsigned)};
    g_init = true;
}

void ensure_init() {
    if (g_init) return;
    const char *r = getenv("SARAS_MPI_RANK"), *s = getenv("SARAS_MPI_SIZE");
    if (!r || !s) { r = getenv("RANK"); s = getenv("WORLD_SIZE"); }
    if (r && s && atoi(s) > 1) { g_proc_mode = true; g_proc_rank = atoi(r); g_proc_size = atoi(s); }
    world_init(1, false);
}

/* registries grow while other rank-threads read them: every access goes through the world mutex */
Comm &comm_of(MPI_Comm c) { std::lock_guard<std::mutex> lk(W.m); return *W.comms[(size_t)c]; }
DType type_of(MPI_Datatype t) { std::lock_guard<std::mutex> lk(W.m); return W.types[(size_t)t]; }
int my_rank_in(Comm &c) {
    for (size_t i = 0; i < c.members.size(); i++) if (c.members[i] == t_rank) return (int)i;
    return -1;
}
size_t type_bytes(MPI_Datatype t) { const DType d = type_of(t); return (size_t)d.count * d.blocklen * d.elem; }

static bool subarray_of(MPI_Datatype t, std::vector<long> &off) {
    std::lock_guard<std::mutex> lk(W.m);
    auto it = W.subarrays.find((int)t);
    if (it == W.subarrays.end()) return false;
    off = it->second;
    return true;
}
void pack(const void *buf, int count, MPI_Datatype t, std::vector<char> &out) {
    const DType d = type_of(t);
    std::vector<long> off;
    if (subarray_of(t, off)) {                      /* count must be 1: the only use (mpidata.cc:282-305) */
        out.resize(off.size() * (size_t)d.elem);
        for (size_t q = 0; q < off.size(); q++) memcpy(out.data() + q * d.elem, (const char *)buf + (size_t)off[q] * d.elem, (size_t)d.elem);
        (void)count;
        return;
    }
    out.resize((size_t)count * type_bytes(t));
    const char *src = (const char *)buf;
    char *dst = out.data();
    const size_t extent = ((size_t)(d.count - 1) * d.stride + d.blocklen) * d.elem;
    for (int c = 0; c < count; c++)
        for (int b = 0; b < d.count; b++) {
            memcpy(dst, src + (size_t)c * extent + (size_t)b * d.stride * d.elem, (size_t)d.blocklen * d.elem);
            dst += (size_t)d.blocklen * d.elem;
        }
}
void unpack(void *buf, int count, MPI_Datatype t, const std::vector<char> &in) {
    const DType d = type_of(t);
    std::vector<long> off;
    if (subarray_of(t, off)) {
        for (size_t q = 0; q < off.size(); q++) memcpy((char *)buf + (size_t)off[q] * d.elem, in.data() + q * d.elem, (size_t)d.elem);
        (void)count;
        return;
    }
    char *dst = (char *)buf;
    const char *src = in.data();
    const size_t extent = ((size_t)(d.count - 1) * d.stride + d.blocklen) * d.elem;
    for (int c = 0; c < count; c++)
        for (int b = 0; b < d.count; b++) {
            memcpy(dst + (size_t)c * extent + (size_t)b * d.stride * d.elem, src, (size_t)d.blocklen * d.elem);
            src += (size_t)d.blocklen * d.elem;
        }
}

/* allgather of equal-sized blobs inside a communicator: the primitive every collective is built on */
void allgather_blob(MPI_Comm comm, const void *mine, size_t nbytes, std::vector<char> &all) {
    ensure_init();
    if (g_proc_mode) {
        if (comm != MPI_COMM_WORLD || !g_hooks.allgather_bytes) {
            fprintf(stderr, "mpi-lite: process-rank mode needs saras_mpi_set_hooks for collectives (world communicator only)\n");
            abort();
        }
        all.resize(nbytes * (size_t)g_proc_size);
        g_hooks.allgather_bytes(g_hooks.ctx, mine, all.data(), (int)nbytes);
        return;
    }
    Comm &c = comm_of(comm);
    const int me = my_rank_in(c);
    c.slots[(size_t)me].assign((const char *)mine, (const char *)mine + nbytes);
    c.bar.wait();
    all.resize(nbytes * c.members.size());
    for (size_t r = 0; r < c.members.size(); r++) memcpy(all.data() + r * nbytes, c.slots[r].data(), nbytes);
    c.bar.wait();
}

template <class T>
void reduce_typed(const std::vector<char> &all, int nranks, int count, MPI_Op op, void *out) {
    const T *a = (const T *)all.data();
    T *o = (T *)out;
    for (int i = 0; i < count; i++) {
        T v = a[i];
        for (int r = 1; r < nranks; r++) {           /* rank order: deterministic */
            const T x = a[(size_t)r * count + i];
            if (op == MPI_SUM) v = v + x; else if (op == MPI_MAX) v = (x > v) ? x : v; else v = (x < v) ? x : v;
        }
        o[i] = v;
    }
}

}  // namespace

extern "C" {

int MPI_Init(int *, char ***) { ensure_init(); return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int code) { fprintf(stderr, "mpi-lite: MPI_Abort(%d)\n", code); exit(code ? code : 1); }
double MPI_Wtime(void) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int MPI_Comm_rank(MPI_Comm comm, int *rank) {
    ensure_init();
    if (g_proc_mode) { *rank = g_proc_rank; return MPI_SUCCESS; }
    *rank = my_rank_in(comm_of(comm));
    return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm comm, int *size) {
    ensure_init();
    if (g_proc_mode) { *size = g_proc_size; return MPI_SUCCESS; }
    *size = (int)comm_of(comm).members.size();
    return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm) {
    ensure_init();
    if (g_proc_mode) { *newcomm = MPI_COMM_NULL; return MPI_SUCCESS; }   /* row/column communicators are not used on the GPU path */
    int mine[3] = {color, key, t_rank};
    std::vector<char> all;
    allgather_blob(comm, mine, sizeof(mine), all);
    Comm &c = comm_of(comm);
    const int n = (int)c.members.size();
    const int *a = (const int *)all.data();
    /* the lowest member of each colour creates the communicator; everyone then looks it up */
    std::vector<std::pair<int, int>> grp;    /* (key, world rank) of my colour */
    int leader = -1;
    for (int r = 0; r < n; r++) if (a[3 * r] == color) { grp.push_back({a[3 * r + 1], a[3 * r + 2]}); if (leader < 0) leader = a[3 * r + 2]; }
    std::stable_sort(grp.begin(), grp.end());
    static std::map<std::pair<int, int>, int> created;      /* (parent, colour) -> id, valid between the two barriers */
    if (t_rank == leader) {
        std::lock_guard<std::mutex> lk(W.m);
        auto nc = std::make_unique<Comm>();
        for (auto &g : grp) nc->members.push_back(g.second);
        nc->bar.n = (int)grp.size();
        nc->slots.resize(grp.size());
        W.comms.push_back(std::move(nc));
        created[{comm, color}] = (int)W.comms.size() - 1;
    }
    c.bar.wait();
    { std::lock_guard<std::mutex> lk(W.m); *newcomm = created[{comm, color}]; }
    c.bar.wait();
    if (t_rank == leader) { std::lock_guard<std::mutex> lk(W.m); created.erase({comm, color}); }
    return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm *comm) { *comm = MPI_COMM_NULL; return MPI_SUCCESS; }

int MPI_Barrier(MPI_Comm comm) {
    ensure_init();
    if (g_proc_mode) { if (g_hooks.barrier) g_hooks.barrier(g_hooks.ctx); return MPI_SUCCESS; }
    comm_of(comm).bar.wait();
    return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype) {
    ensure_init();
    std::lock_guard<std::mutex> lk(W.m);
    const DType o = W.types[(size_t)oldtype];
    W.types.push_back({1, count * o.count * o.blocklen, count * o.count * o.blocklen, o.elem});
    *newtype = (int)W.types.size() - 1;
    return MPI_SUCCESS;
}
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype *newtype) {
    ensure_init();
    std::lock_guard<std::mutex> lk(W.m);
    const DType o = W.types[(size_t)oldtype];
    W.types.push_back({count, blocklength * o.blocklen, stride * o.blocklen, o.elem});
    *newtype = (int)W.types.size() - 1;
    return MPI_SUCCESS;
}
int MPI_Type_create_subarray(int ndims, const int *sizes, const int *subsizes, const int *starts, int order, MPI_Datatype oldtype,
                             MPI_Datatype *newtype) {
    ensure_init();
    if (order != MPI_ORDER_C || ndims < 1 || ndims > 3) { fprintf(stderr, "mpi-lite: MPI_Type_create_subarray: C order, 1-3 dimensions only\n"); abort(); }
    std::lock_guard<std::mutex> lk(W.m);
    const DType o = W.types[(size_t)oldtype];
    int sz[3] = {1, 1, 1}, sub[3] = {1, 1, 1}, st[3] = {0, 0, 0};
    for (int d = 0; d < ndims; d++) { sz[3 - ndims + d] = sizes[d]; sub[3 - ndims + d] = subsizes[d]; st[3 - ndims + d] = starts[d]; }
    std::vector<long> off;
    off.reserve((size_t)sub[0] * sub[1] * sub[2]);
    for (int a = 0; a < sub[0]; a++) for (int b = 0; b < sub[1]; b++) for (int c = 0; c < sub[2]; c++)
        off.push_back(((long)(st[0] + a) * sz[1] + (st[1] + b)) * sz[2] + (st[2] + c));
    W.types.push_back({(int)off.size(), 1, 1, o.elem * o.count * o.blocklen});
    *newtype = (int)W.types.size() - 1;
    W.subarrays[*newtype] = off;
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *) { return MPI_SUCCESS; }

int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *req) {
    ensure_init();
    if (source == MPI_PROC_NULL) { *req = MPI_REQUEST_NULL; return MPI_SUCCESS; }
    std::lock_guard<std::mutex> lk(W.m);
    auto &pl = W.posted[(size_t)t_rank];
    pl.push_back({buf, count, type, source, tag, comm, false, true});
    *req = (int)pl.size() - 1;
    return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm) {
    ensure_init();
    if (g_proc_mode) { fprintf(stderr, "mpi-lite: MPI_Send is not available in process-rank mode\n"); abort(); }
    Comm &c = comm_of(comm);
    c.bar.wait();                                   /* every rank has posted its receives and reached its k-th send */
    if (dest != MPI_PROC_NULL) {
        const int me = my_rank_in(c);
        const int wdest = c.members[(size_t)dest];
        std::vector<char> tmp;
        pack(buf, count, type, tmp);
        Posted *hit = nullptr;
        {
            std::lock_guard<std::mutex> lk(W.m);
            for (auto &p : W.posted[(size_t)wdest])
                if (p.active && !p.done && p.comm == comm && p.source == me && (p.tag == tag || p.tag == MPI_ANY_TAG)) { hit = &p; break; }
        }
        if (!hit) { fprintf(stderr, "mpi-lite: MPI_Send(rank %d -> %d, tag %d) has no posted receive\n", me, dest, tag); abort(); }
        unpack(hit->buf, hit->count, hit->type, tmp);
        hit->done = true;
    }
    c.bar.wait();                                   /* all k-th sends have landed */
    return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *) {
    if (*req == MPI_REQUEST_NULL) return MPI_SUCCESS;
    std::lock_guard<std::mutex> lk(W.m);
    auto &pl = W.posted[(size_t)t_rank];
    Posted &p = pl[(size_t)*req];
    if (!p.done) { fprintf(stderr, "mpi-lite: MPI_Wait on a receive that no lock-step send has matched\n"); abort(); }
    p.active = false;
    *req = MPI_REQUEST_NULL;
    bool any = false;
    for (auto &q : pl) any = any || q.active;
    if (!any) pl.clear();
    return MPI_SUCCESS;
}
int MPI_Waitall(int count, MPI_Request *reqs, MPI_Status *) {
    for (int i = 0; i < count; i++) MPI_Wait(&reqs[i], MPI_STATUS_IGNORE);
    return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
    ensure_init();
    const size_t nb = (size_t)count * type_bytes(type);
    std::vector<char> all;
    allgather_blob(comm, sendbuf, nb, all);
    const int n = (int)(all.size() / nb);
    if (type == MPI_DOUBLE) reduce_typed<double>(all, n, count, op, recvbuf);
    else if (type == MPI_FLOAT) reduce_typed<float>(all, n, count, op, recvbuf);
    else if (type == MPI_INT) reduce_typed<int>(all, n, count, op, recvbuf);
    else if (type == MPI_LONG) reduce_typed<long>(all, n, count, op, recvbuf);
    else { fprintf(stderr, "mpi-lite: MPI_Allreduce: unsupported datatype\n"); abort(); }
    return MPI_SUCCESS;
}
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm) {
    int me = 0;
    MPI_Comm_rank(comm, &me);
    std::vector<char> tmp((size_t)count * type_bytes(type));
    MPI_Allreduce(sendbuf, tmp.data(), count, type, op, comm);
    if (me == root) memcpy(recvbuf, tmp.data(), tmp.size());
    return MPI_SUCCESS;
}
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int, MPI_Datatype, MPI_Comm comm) {
    ensure_init();
    const size_t nb = (size_t)sendcount * type_bytes(sendtype);
    std::vector<char> all;
    allgather_blob(comm, sendbuf, nb, all);
    memcpy(recvbuf, all.data(), all.size());
    return MPI_SUCCESS;
}
int MPI_Allgatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, const int *recvcounts,
                   const int *displs, MPI_Datatype recvtype, MPI_Comm comm) {
    ensure_init();
    int n = 1;
    MPI_Comm_size(comm, &n);
    int maxc = 0;
    for (int r = 0; r < n; r++) maxc = std::max(maxc, recvcounts[r]);
    const size_t eb = type_bytes(recvtype);
    std::vector<char> mine((size_t)maxc * eb, 0), all;
    memcpy(mine.data(), sendbuf, (size_t)sendcount * type_bytes(sendtype));
    allgather_blob(comm, mine.data(), mine.size(), all);
    for (int r = 0; r < n; r++)
        memcpy((char *)recvbuf + (size_t)displs[r] * eb, all.data() + (size_t)r * mine.size(), (size_t)recvcounts[r] * eb);
    return MPI_SUCCESS;
}
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
    ensure_init();
    const size_t nb = (size_t)count * type_bytes(type);
    std::vector<char> all;
    allgather_blob(comm, buf, nb, all);
    memcpy(buf, all.data() + (size_t)root * nb, nb);
    return MPI_SUCCESS;
}

void saras_mpi_run(int nranks, void (*fn)(void *), void *arg) {
    world_init(nranks, true);
    g_proc_mode = false;
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; r++) th.emplace_back([=] { t_rank = r; fn(arg); });
    for (auto &t : th) t.join();
    world_init(1, false);
}
void saras_mpi_set_hooks(const saras_mpi_hooks *hooks) { if (hooks) g_hooks = *hooks; else g_hooks = {nullptr, nullptr, nullptr}; }
void saras_mpi_set_process_rank(int rank, int size) {
    ensure_init();
    g_proc_mode = size > 1;
    g_proc_rank = rank;
    g_proc_size = size;
}

}  /* extern "C" */
