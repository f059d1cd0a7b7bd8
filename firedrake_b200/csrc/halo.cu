// Halo exchange (C1/C2) and global reductions (C3) over NCCL.
//
// Replaces firedrake/halo.py:124-172 (PetscSF bcast = global->local REPLACE,
// PetscSF reduce = local->global SUM, both in place on dat._data) and the
// Iallreduce of pyop2/parloop.py:411-442.  One process per GPU; each halo is a
// list of neighbours with a send list (owned dofs that are ghosts elsewhere)
// and a receive list (my ghost tail).  begin() packs on the compute stream and
// posts ncclSend/ncclRecv on a separate communication stream so the exchange
// overlaps the core-cell kernel exactly as pyop2/parloop.py:250-253 overlaps
// MPI with `_compute(core_part)`; end() makes the compute stream wait and
// unpacks (REPLACE) or accumulates (SUM).
//
// NCCL is resolved at run time with dlopen("libnccl.so.2"): inside a process
// that already imported torch this binds to torch's bundled NCCL (same soname),
// otherwise to the system library -- never two copies in one process.
#include <dlfcn.h>
#include <nccl.h>

#include <vector>

#include "common.cuh"

using namespace fdb;

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_nranks = 1;
cudaStream_t g_comm_stream = nullptr;

int load_nccl()
{
    if (g_nccl.handle) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return 1;
    }
#define SYM(field, name)                                                    \
    *(void **)(&g_nccl.field) = dlsym(h, name);                             \
    if (!g_nccl.field) {                                                    \
        set_error("libnccl: missing symbol %s", name);                      \
        return 1;                                                           \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return 0;
}

#define FDB_NCCL(call)                                                              \
    do {                                                                            \
        ncclResult_t r_ = (call);                                                   \
        if (r_ != ncclSuccess) {                                                    \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                      g_nccl.GetErrorString(r_));                                   \
            return 1;                                                               \
        }                                                                           \
    } while (0)

__global__ void k_pack(const double *__restrict__ dat, int cdim, const fdb_int *__restrict__ idx,
                       long long n, double *__restrict__ buf)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = n * cdim;
    for (; i < tot; i += (long long)gridDim.x * blockDim.x) {
        long long k = i / cdim;
        int c = (int)(i - k * cdim);
        buf[i] = dat[(long long)idx[k] * cdim + c];
    }
}

template <bool ADD>
__global__ void k_unpack(double *__restrict__ dat, int cdim, const fdb_int *__restrict__ idx,
                         long long n, const double *__restrict__ buf)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = n * cdim;
    for (; i < tot; i += (long long)gridDim.x * blockDim.x) {
        long long k = i / cdim;
        int c = (int)(i - k * cdim);
        long long j = (long long)idx[k] * cdim + c;
        // ADD is launched once per NEIGHBOUR (exchange_end): within one neighbour's list an
        // owned dof appears once, so the plain read-modify-write is race free and the order in
        // which neighbours are summed is fixed (deterministic); a dof shared with several
        // neighbours (partition corners) is updated by consecutive launches
        if (ADD) dat[j] += buf[i];
        else dat[j] = buf[i];
    }
}

int grid_for(long long n)
{
    long long b = (n + 255) / 256;
    long long cap = (long long)ctx().sm_count * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

struct fdb_halo_s {
    int nneigh = 0;
    std::vector<int> ranks;
    std::vector<long long> send_off, recv_off;   // prefix offsets (in dofs) per neighbour
    fdb_int *d_send_idx = nullptr, *d_recv_idx = nullptr;
    long long nsend = 0, nrecv = 0;
    int max_cdim = 0;
    double *d_send_buf = nullptr, *d_recv_buf = nullptr;
    cudaEvent_t ev_packed = nullptr, ev_done = nullptr;
    int pending = 0;     // 1: g2l in flight, 2: l2g in flight
};

extern "C" {

int fdb_comm_get_unique_id(char *out128)
{
    if (load_nccl()) return 1;
    ncclUniqueId id;
    FDB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return 0;
}

int fdb_comm_init(int rank, int nranks, const char *id128)
{
    if (require_init()) return 1;
    if (load_nccl()) return 1;
    if (g_comm) {
        set_error("fdb_comm_init: communicator already initialised");
        return 1;
    }
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    FDB_NCCL(g_nccl.CommInitRank(&g_comm, nranks, id, rank));
    g_rank = rank;
    g_nranks = nranks;
    FDB_CUDA(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking));
    return 0;
}

int fdb_comm_finalize(void)
{
    if (g_comm) {
        cudaStreamSynchronize(g_comm_stream);
        g_nccl.CommDestroy(g_comm);
        cudaStreamDestroy(g_comm_stream);
        g_comm = nullptr;
    }
    return 0;
}

int fdb_comm_rank(void) { return g_rank; }
int fdb_comm_size(void) { return g_nranks; }

int fdb_halo_create(int nneigh, const int *ranks, const fdb_int *send_counts,
                    const fdb_int *send_idx, const fdb_int *recv_counts, const fdb_int *recv_idx,
                    int max_cdim, fdb_halo_t *out)
{
    if (require_init()) return 1;
    fdb_halo_s *h = new fdb_halo_s;
    h->nneigh = nneigh;
    h->max_cdim = max_cdim < 1 ? 1 : max_cdim;
    h->send_off.push_back(0);
    h->recv_off.push_back(0);
    for (int i = 0; i < nneigh; i++) {
        h->ranks.push_back(ranks[i]);
        h->send_off.push_back(h->send_off.back() + send_counts[i]);
        h->recv_off.push_back(h->recv_off.back() + recv_counts[i]);
    }
    h->nsend = h->send_off.back();
    h->nrecv = h->recv_off.back();
    cudaStream_t st = ctx().stream;
    FDB_CUDA(cudaMalloc(&h->d_send_idx, sizeof(fdb_int) * (h->nsend + 1)));
    FDB_CUDA(cudaMalloc(&h->d_recv_idx, sizeof(fdb_int) * (h->nrecv + 1)));
    // buffers are sized for the larger direction: l2g sends what g2l receives
    long long nbuf = (h->nsend > h->nrecv ? h->nsend : h->nrecv) * h->max_cdim + 1;
    FDB_CUDA(cudaMalloc(&h->d_send_buf, sizeof(double) * nbuf));
    FDB_CUDA(cudaMalloc(&h->d_recv_buf, sizeof(double) * nbuf));
    FDB_CUDA(cudaMemcpyAsync(h->d_send_idx, send_idx, sizeof(fdb_int) * h->nsend,
                             cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaMemcpyAsync(h->d_recv_idx, recv_idx, sizeof(fdb_int) * h->nrecv,
                             cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    FDB_CUDA(cudaEventCreateWithFlags(&h->ev_packed, cudaEventDisableTiming));
    FDB_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
    *out = h;
    return 0;
}

int fdb_halo_destroy(fdb_halo_t h)
{
    if (!h) return 0;
    if (ctx().ready) {
        cudaStreamSynchronize(ctx().stream);
        if (g_comm_stream) cudaStreamSynchronize(g_comm_stream);
        cudaFree(h->d_send_idx);
        cudaFree(h->d_recv_idx);
        cudaFree(h->d_send_buf);
        cudaFree(h->d_recv_buf);
        cudaEventDestroy(h->ev_packed);
        cudaEventDestroy(h->ev_done);
    }
    delete h;
    return 0;
}

// direction 0: owners -> ghosts (send list -> recv list); 1: ghosts -> owners
static int exchange_begin(fdb_halo_t h, double *dat, int cdim, int reverse)
{
    if (require_init()) return 1;
    if (!g_comm && h->nneigh > 0) {
        set_error("halo exchange without a communicator: call fdb_comm_init");
        return 1;
    }
    if (cdim > h->max_cdim) {
        set_error("halo exchange: cdim %d exceeds the halo's max_cdim %d", cdim, h->max_cdim);
        return 1;
    }
    if (h->pending) {
        set_error("halo exchange already in flight");
        return 1;
    }
    cudaStream_t st = ctx().stream;
    const fdb_int *pack_idx = reverse ? h->d_recv_idx : h->d_send_idx;
    const long long npack = reverse ? h->nrecv : h->nsend;
    if (npack > 0) {
        k_pack<<<grid_for(npack * cdim), 256, 0, st>>>(dat, cdim, pack_idx, npack, h->d_send_buf);
        FDB_LAUNCH_CHECK();
    }
    FDB_CUDA(cudaEventRecord(h->ev_packed, st));
    FDB_CUDA(cudaStreamWaitEvent(g_comm_stream, h->ev_packed, 0));
    if (h->nneigh > 0) {
        const std::vector<long long> &so = reverse ? h->recv_off : h->send_off;
        const std::vector<long long> &ro = reverse ? h->send_off : h->recv_off;
        FDB_NCCL(g_nccl.GroupStart());
        for (int i = 0; i < h->nneigh; i++) {
            long long ns = (so[i + 1] - so[i]) * cdim, nr = (ro[i + 1] - ro[i]) * cdim;
            if (ns > 0)
                FDB_NCCL(g_nccl.Send(h->d_send_buf + so[i] * cdim, ns, ncclDouble, h->ranks[i], g_comm,
                                     g_comm_stream));
            if (nr > 0)
                FDB_NCCL(g_nccl.Recv(h->d_recv_buf + ro[i] * cdim, nr, ncclDouble, h->ranks[i], g_comm,
                                     g_comm_stream));
        }
        FDB_NCCL(g_nccl.GroupEnd());
    }
    FDB_CUDA(cudaEventRecord(h->ev_done, g_comm_stream));
    h->pending = reverse ? 2 : 1;
    return 0;
}

static int exchange_end(fdb_halo_t h, double *dat, int cdim, int reverse)
{
    if (require_init()) return 1;
    if (h->pending != (reverse ? 2 : 1)) {
        set_error("halo end without matching begin");
        return 1;
    }
    cudaStream_t st = ctx().stream;
    FDB_CUDA(cudaStreamWaitEvent(st, h->ev_done, 0));
    const fdb_int *unpack_idx = reverse ? h->d_send_idx : h->d_recv_idx;
    const long long nun = reverse ? h->nsend : h->nrecv;
    if (nun > 0) {
        if (reverse) {
            // local->global SUM (firedrake/halo.py:140-172): one launch per neighbour list
            for (int i = 0; i < h->nneigh; i++) {
                const long long o = h->send_off[i], cnt = h->send_off[i + 1] - o;
                if (cnt <= 0) continue;
                k_unpack<true><<<grid_for(cnt * cdim), 256, 0, st>>>(dat, cdim, unpack_idx + o, cnt,
                                                                     h->d_recv_buf + o * cdim);
                FDB_LAUNCH_CHECK();
            }
            h->pending = 0;
            return 0;
        } else
            k_unpack<false><<<grid_for(nun * cdim), 256, 0, st>>>(dat, cdim, unpack_idx, nun, h->d_recv_buf);
        FDB_LAUNCH_CHECK();
    }
    h->pending = 0;
    return 0;
}

int fdb_halo_global_to_local_begin(fdb_halo_t h, double *dat, int cdim) { return exchange_begin(h, dat, cdim, 0); }
int fdb_halo_global_to_local_end(fdb_halo_t h, double *dat, int cdim) { return exchange_end(h, dat, cdim, 0); }
int fdb_halo_local_to_global_begin(fdb_halo_t h, double *dat, int cdim) { return exchange_begin(h, dat, cdim, 1); }
int fdb_halo_local_to_global_end(fdb_halo_t h, double *dat, int cdim) { return exchange_end(h, dat, cdim, 1); }

int fdb_allreduce(double *dev, int n, int op)
{
    if (require_init()) return 1;
    if (!g_comm) {
        if (g_nranks == 1) return 0;
        set_error("fdb_allreduce without a communicator");
        return 1;
    }
    ncclRedOp_t rop = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
    FDB_NCCL(g_nccl.AllReduce(dev, dev, n, ncclDouble, rop, g_comm, ctx().stream));
    return 0;
}

}  // extern "C"
