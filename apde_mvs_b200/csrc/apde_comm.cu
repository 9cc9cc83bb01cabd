// apde_comm.cu -- multi-GPU jobs: depth-map exchange between passes and map gathers before fusion, NCCL over
// NVLink / NVSwitch.  One context per GPU; the contexts of a job live in one process (a host thread each) or in one process per
// GPU.  The reference has nothing to compare with here (one process per GPU on disjoint scans, run.py:127-153); the contract is
// that N GPUs reproduce the single-GPU Jacobi run bit for bit (tests/test_gpu_multi.py).
//
// Exchange design.  A view's depth map is final once its problem has been handed back to the view store (launch_finish).  At
// that point the owner queues a broadcast of that one row on a separate exchange stream, ordered after the compute stream by an
// event, and goes on with the next view of its block: the transfer of view k overlaps the PatchMatch pass of view k + 1, and
// only the last view's broadcast is exposed.  All ranks issue the same sequence of grouped broadcasts (k-th view of rank 0,
// of rank 1, ...), so no host-side synchronisation is needed.  New maps go to the write half of the double-buffered depth pool,
// which nothing reads during the pass (Jacobi), so receiving needs no ordering against the local kernels.
//
// NCCL is resolved with dlopen at the first apde_comm_* call: libapde.so itself has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "apde_context.h"

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("APDE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !n[0]) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = std::string("NCCL not found (libnccl.so.2): ") + dlerror(); return; }
#define APDE_SYM(field, sym)                                                     \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym));      \
    if (!api.field) { api.error = std::string("NCCL symbol missing: ") + sym; return; }
        APDE_SYM(GetUniqueId, "ncclGetUniqueId")
        APDE_SYM(CommInitRank, "ncclCommInitRank")
        APDE_SYM(CommDestroy, "ncclCommDestroy")
        APDE_SYM(Broadcast, "ncclBroadcast")
        APDE_SYM(GroupStart, "ncclGroupStart")
        APDE_SYM(GroupEnd, "ncclGroupEnd")
        APDE_SYM(GetErrorString, "ncclGetErrorString")
#undef APDE_SYM
    });
    return api;
}

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;  // the exchange stream
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr, ev_w0 = nullptr, ev_w1 = nullptr;
    uint64_t bytes = 0;             // received since the last join
    bool pending_wait = false;      // ev_w0 / ev_w1 recorded and not yet read
};

#define NC(call)                                                                                                      \
    do {                                                                                                              \
        ncclResult_t r_ = (call);                                                                                     \
        if (r_ != ncclSuccess) return fail(APDE_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r_)); \
    } while (0)

int apde_comm_attached(const apde_context *c) { return c && c->comm && c->comm->world > 1; }

int apde_comm_block(const apde_context *c, int *first_view, int *num_views) {
    if (!c || !c->comm) return fail(APDE_ERR_STATE, "comm: context is not part of a multi-GPU job");
    apde_view_block(c->V, c->comm->world, c->comm->rank, first_view, num_views);
    return APDE_OK;
}

int apde_comm_share_depth_row(apde_context *c, int k, int which_pool, int w, int h) {
    Comm *m = c->comm;
    if (!m) return fail(APDE_ERR_STATE, "comm: context is not part of a multi-GPU job");
    NcclApi &api = nccl();
    const size_t Pfull = (size_t)c->W * c->H, count = (size_t)w * h;
    // the broadcast of this rank's k-th view starts when everything queued on the compute stream so far has finished
    CU(cudaEventRecord(m->ev_ready, c->stream));
    CU(cudaStreamWaitEvent(m->stream, m->ev_ready, 0));
    NC(api.GroupStart());
    for (int r = 0; r < m->world; ++r) {
        int first, cnt;
        apde_view_block(c->V, m->world, r, &first, &cnt);
        if (k >= cnt) continue;
        float *row = c->d_depth_pool[which_pool] + (size_t)(first + k) * Pfull;
        NC(api.Broadcast(row, row, count, ncclFloat, r, m->comm, m->stream));
        if (r != m->rank) m->bytes += count * sizeof(float);
    }
    NC(api.GroupEnd());
    return APDE_OK;
}

int apde_comm_join(apde_context *c, double *exposed_ms, uint64_t *bytes) {
    Comm *m = c->comm;
    if (!m) return APDE_OK;
    CU(cudaEventRecord(m->ev_done, m->stream));
    CU(cudaEventRecord(m->ev_w0, c->stream));
    CU(cudaStreamWaitEvent(c->stream, m->ev_done, 0));
    CU(cudaEventRecord(m->ev_w1, c->stream));
    if (exposed_ms) {
        CU(cudaEventSynchronize(m->ev_w1));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, m->ev_w0, m->ev_w1));
        *exposed_ms += ms;
    }
    if (bytes) *bytes += m->bytes;
    m->bytes = 0;
    return APDE_OK;
}

extern "C" {

int apde_comm_create_id(uint8_t id[APDE_COMM_ID_BYTES]) {
    if (!id) return fail(APDE_ERR_ARG, "comm_create_id: id is NULL");
    NcclApi &api = nccl();
    if (!api.error.empty()) return fail(APDE_ERR_STATE, "%s", api.error.c_str());
    static_assert(sizeof(ncclUniqueId) == APDE_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NC(api.GetUniqueId(&u));
    memcpy(id, &u, APDE_COMM_ID_BYTES);
    return APDE_OK;
}

int apde_comm_init(apde_context *c, const uint8_t id[APDE_COMM_ID_BYTES], int rank, int world) {
    if (!c || !id) return fail(APDE_ERR_ARG, "comm_init: null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(APDE_ERR_ARG, "comm_init: rank %d of %d", rank, world);
    if (c->comm) return fail(APDE_ERR_STATE, "comm_init: the context already belongs to a job");
    NcclApi &api = nccl();
    if (!api.error.empty()) return fail(APDE_ERR_STATE, "%s", api.error.c_str());
    CU(cudaSetDevice(c->device));
    Comm *m = new Comm();
    m->rank = rank; m->world = world;
    auto drop = [&](int rc) {  // a half-built job must not leak its stream / events / communicator
        if (m->comm) api.CommDestroy(m->comm);
        if (m->ev_ready) cudaEventDestroy(m->ev_ready);
        if (m->ev_done) cudaEventDestroy(m->ev_done);
        if (m->ev_w0) cudaEventDestroy(m->ev_w0);
        if (m->ev_w1) cudaEventDestroy(m->ev_w1);
        if (m->stream) cudaStreamDestroy(m->stream);
        delete m;
        return rc;
    };
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&m->ev_w0);
    if (e == cudaSuccess) e = cudaEventCreate(&m->ev_w1);
    if (e != cudaSuccess) return drop(fail(APDE_ERR_CUDA, "comm_init: %s", cudaGetErrorString(e)));
    ncclUniqueId u;
    memcpy(&u, id, APDE_COMM_ID_BYTES);
    ncclResult_t r = api.CommInitRank(&m->comm, world, u, rank);
    if (r != ncclSuccess) { m->comm = nullptr; return drop(fail(APDE_ERR_CUDA, "ncclCommInitRank(rank %d of %d): %s", rank, world, api.GetErrorString(r))); }
    c->comm = m;
    return APDE_OK;
}

int apde_comm_block_of(int num_views, int world, int rank, int *first_view, int *count) {
    if (num_views < 0 || world < 1 || rank < 0 || rank >= world || !first_view || !count) return fail(APDE_ERR_ARG, "comm_block_of: bad argument");
    apde_view_block(num_views, world, rank, first_view, count);
    return APDE_OK;
}

int apde_comm_info(apde_context *c, int *rank, int *world, int *first_view, int *num_views) {
    if (!c) return fail(APDE_ERR_ARG, "comm_info: null context");
    const int rk = c->comm ? c->comm->rank : 0, wd = c->comm ? c->comm->world : 1;
    if (rank) *rank = rk;
    if (world) *world = wd;
    int first = 0, cnt = c->V;
    if (c->V > 0) apde_view_block(c->V, wd, rk, &first, &cnt);
    if (first_view) *first_view = first;
    if (num_views) *num_views = cnt;
    return APDE_OK;
}

int apde_exchange(apde_context *c, int which) {
    if (!c || c->V <= 0) return fail(APDE_ERR_STATE, "exchange: no scene");
    Comm *m = c->comm;
    if (!m || m->world == 1) return APDE_OK;
    NcclApi &api = nccl();
    CU(cudaSetDevice(c->device));
    void *ptr = nullptr;
    size_t total = 0, per = 0;
    int rc = apde_map_pool(c, which, &ptr, &total, &per);  // synchronises c->stream: the rows are final
    if (rc) return rc;
    NC(api.GroupStart());
    for (int r = 0; r < m->world; ++r) {
        int first, cnt;
        apde_view_block(c->V, m->world, r, &first, &cnt);
        if (cnt == 0) continue;
        uint8_t *rows = static_cast<uint8_t *>(ptr) + (size_t)first * per;
        NC(api.Broadcast(rows, rows, (size_t)cnt * per, ncclUint8, r, m->comm, m->stream));
    }
    NC(api.GroupEnd());
    CU(cudaStreamSynchronize(m->stream));
    return APDE_OK;
}

int apde_fuse_collective(apde_context *c, int variant, int use_weak_filter, float *xyz, float *bgr, int64_t max_points,
                         int64_t *num_points) {
    if (!c || c->V <= 0 || !num_points) return fail(APDE_ERR_STATE, "fuse_collective: bad argument");
    Comm *m = c->comm;
    if (!m || m->world == 1) return apde_fuse_variant(c, variant, use_weak_filter, xyz, bgr, max_points, num_points);
    int first, cnt, rc;
    apde_view_block(c->V, m->world, m->rank, &first, &cnt);
    int mw = 0, mh = 0;
    if (cnt > 0 && (rc = apde_view_download(c, first, nullptr, nullptr, nullptr, nullptr, &mw, &mh))) return rc;
    if (cnt == 0) { mw = c->W; mh = c->H; }  // (more ranks than views: this rank only receives)
    // depth maps are already everywhere (exchanged pass by pass); normals, states and confidences stayed with their owners
    for (int which : {APDE_POOL_NORMAL, APDE_POOL_WEAK, APDE_POOL_CONFIDENCE})
        if ((rc = apde_exchange(c, which))) return rc;
    if ((rc = apde_views_mark_maps(c, mw, mh))) return rc;
    int filter_mode = 0;
    if (use_weak_filter) {
        // WeakVisFilter's work items are reference views (the reference's thread pool, APD.cpp:1040-1047): each rank filters
        // its own block against the maps of all views, then the skip maps are gathered
        if ((rc = apde_weak_vis_filter_range(c, first, cnt, nullptr))) return rc;
        if ((rc = apde_exchange(c, APDE_POOL_SKIP))) return rc;
        filter_mode = APDE_WEAK_FILTER_KEEP;
    }
    *num_points = 0;
    if (m->rank != 0) return APDE_OK;
    // the greedy claim order runs over all views through masks[] (APD.cpp:1149,1176,1209): one rank
    return apde_fuse_variant(c, variant, filter_mode, xyz, bgr, max_points, num_points);
}

int apde_comm_destroy(apde_context *c) {
    if (!c || !c->comm) return APDE_OK;
    Comm *m = c->comm;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(m->stream);
    if (m->comm) nccl().CommDestroy(m->comm);
    cudaEventDestroy(m->ev_ready); cudaEventDestroy(m->ev_done); cudaEventDestroy(m->ev_w0); cudaEventDestroy(m->ev_w1);
    cudaStreamDestroy(m->stream);
    delete m;
    c->comm = nullptr;
    return APDE_OK;
}

}  // extern "C"
