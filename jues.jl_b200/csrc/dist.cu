// Multi-GPU plumbing: one process per GPU.  NCCL is loaded at run time (dlopen) so the library has
// no link-time dependency on a particular NCCL build; a process that already loaded NCCL (e.g.
// through torch.distributed) gets that same copy.  The collectives of the path:
// all-gather of the half residual and of the new T2 slab, sum-all-reduce of small partials.
#include "dist.h"
#include "api_util.h"

#include <dlfcn.h>
#include <nccl.h>
#include <thread>
#include <functional>

namespace jues {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

void load_nccl() {
    if (g_nccl.handle) return;
    const char* names[] = {getenv("JUES_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    std::string tried;
    for (const char* n : names) {
        if (!n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
        tried += std::string(n) + " (" + dlerror() + "); ";
    }
    if (!h) throw Error(JUES_B200_ENCCL, "cannot load NCCL: " + tried);
    auto sym = [&](const char* s) {
        void* p = dlsym(h, s);
        if (!p) throw Error(JUES_B200_ENCCL, std::string("NCCL symbol missing: ") + s);
        return p;
    };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
    g_nccl.CommAbort = (decltype(g_nccl.CommAbort))sym("ncclCommAbort");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    g_nccl.handle = h;
}

void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess)
        throw Error(JUES_B200_ENCCL, std::string(what) + ": " + g_nccl.GetErrorString(r));
}

}  // namespace

void dist_teardown(jues_ctx* ctx) {
    if (ctx->nccl_comm && g_nccl.handle) {
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
}

void all_gather_inplace(jues_ctx* ctx, double* full, size_t count_per_rank) {
    if (ctx->nranks == 1) return;
    if (ctx->sync_comm) JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    // rank r's slab already sits at full + r*count_per_rank (in-place all-gather)
    nccl_check(g_nccl.AllGather(full + (size_t)ctx->rank * count_per_rank, full, count_per_rank, ncclDouble,
                                (ncclComm_t)ctx->nccl_comm, ctx->stream),
               "ncclAllGather");
    if (ctx->sync_comm) JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.collectives++;
    ctx->stats.collective_bytes += (double)count_per_rank * 8.0 * (ctx->nranks - 1);
}

void all_reduce_sum(jues_ctx* ctx, double* buf, size_t count) {
    if (ctx->nranks == 1) return;
    if (ctx->sync_comm) JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    nccl_check(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream),
               "ncclAllReduce");
    if (ctx->sync_comm) JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.collectives++;
    ctx->stats.collective_bytes += (double)count * 8.0;
}

void all_to_all_v(jues_ctx* ctx, const double* send, const size_t* send_off, const size_t* send_cnt,
                  double* recv, const size_t* recv_off, const size_t* recv_cnt) {
    const int me = ctx->rank;
    // own block: plain device copy
    if (send_cnt[me])
        JUES_CUDA(cudaMemcpyAsync(recv + recv_off[me], send + send_off[me], send_cnt[me] * sizeof(double),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->nranks == 1) return;
    if (ctx->sync_comm) JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
    double got = 0.0;
    for (int k = 1; k < ctx->nranks; ++k) {
        const int to = (me + k) % ctx->nranks, from = (me - k + ctx->nranks) % ctx->nranks;
        if (send_cnt[to])
            nccl_check(g_nccl.Send(send + send_off[to], send_cnt[to], ncclDouble, to, (ncclComm_t)ctx->nccl_comm,
                                   ctx->stream), "ncclSend");
        if (recv_cnt[from])
            nccl_check(g_nccl.Recv(recv + recv_off[from], recv_cnt[from], ncclDouble, from,
                                   (ncclComm_t)ctx->nccl_comm, ctx->stream), "ncclRecv");
        got += (double)recv_cnt[from] * 8.0;
    }
    nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
    if (ctx->sync_comm) JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.collectives++;
    ctx->stats.collective_bytes += got;
}

double all_reduce_scalar(jues_ctx* ctx, double x) {
    if (ctx->nranks == 1) return x;
    double* slot = ctx->red_dev + ctx->red_cap - 2;
    JUES_CUDA(cudaMemcpyAsync(slot, &x, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    all_reduce_sum(ctx, slot, 1);
    double r = 0.0;
    JUES_CUDA(cudaMemcpyAsync(&r, slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    return r;
}

// Run `fn(member, is_leader)` on every member of the leader's group, one host thread per GPU (the leader's
// share on the calling thread).  Returns the first non-zero status; its message lands in the leader's
// last_error.  If a member fails while the others wait in a collective, the communicators are aborted so
// that nobody hangs (the group is unusable afterwards and reports JUES_B200_ENCCL).
int group_run(jues_ctx* lead, const std::function<int(jues_ctx*, bool)>& fn) {
    std::vector<jues_ctx*>& g = *lead->group;
    const int n = (int)g.size();
    std::vector<int> rc((size_t)n, 0);
    for (jues_ctx* m : g) m->in_group_call = true;
    std::vector<std::thread> th;
    auto body = [&](int r) {
        cudaSetDevice(g[r]->device);
        rc[r] = fn(g[r], r == 0);
        if (rc[r] != 0 && g_nccl.CommAbort)
            for (jues_ctx* m : g)
                if (m->nccl_comm) { g_nccl.CommAbort((ncclComm_t)m->nccl_comm); m->nccl_comm = nullptr; }
    };
    for (int r = 1; r < n; ++r) th.emplace_back(body, r);
    body(0);
    for (auto& t : th) t.join();
    for (jues_ctx* m : g) m->in_group_call = false;
    cudaSetDevice(lead->device);
    for (int r = 0; r < n; ++r)
        if (rc[r] != 0) {
            if (r != 0) lead->last_error = "rank " + std::to_string(r) + ": " + g[r]->last_error;
            return rc[r];
        }
    return 0;
}

}  // namespace jues

using namespace jues;

extern "C" int jues_b200_init_multi(jues_ctx** out, int ngpu) {
    if (!out) return JUES_B200_EINVAL;
    *out = nullptr;
    std::vector<jues_ctx*>* members = new std::vector<jues_ctx*>();
    auto fail = [&](int code, const std::string& msg) {
        for (jues_ctx* m : *members) { m->group = nullptr; jues_b200_finalize(m); }
        delete members;
        g_init_error = msg;
        return code;
    };
    try {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return fail(JUES_B200_ECUDA, "no CUDA device available; jues_b200 has no CPU fallback");
        }
        if (ngpu <= 0) ngpu = ndev;
        if (ngpu > ndev) return fail(JUES_B200_EINVAL, "jues_b200_init_multi: more GPUs requested than visible");
        for (int r = 0; r < ngpu; ++r) {
            jues_ctx* m = nullptr;
            const int rc = jues_b200_init(&m, r);
            if (rc != 0) return fail(rc, g_init_error);
            members->push_back(m);
        }
        jues_ctx* lead = (*members)[0];
        if (ngpu > 1) {
            load_nccl();
            std::vector<ncclComm_t> comms((size_t)ngpu);
            std::vector<int> devs((size_t)ngpu);
            for (int r = 0; r < ngpu; ++r) devs[r] = r;
            nccl_check(g_nccl.CommInitAll(comms.data(), ngpu, devs.data()), "ncclCommInitAll");
            for (int r = 0; r < ngpu; ++r) {
                (*members)[r]->nccl_comm = comms[r];
                (*members)[r]->rank = r;
                (*members)[r]->nranks = ngpu;
            }
        }
        for (jues_ctx* m : *members) m->leader = lead;
        lead->group = members;
        if (ngpu > 1) {
            // first collectives set up the NVLink / peer connections: here, not inside a timed call
            const int rc = group_run(lead, [&](jues_ctx* m, bool) {
                try {
                    double* slot = m->red_dev + m->red_cap - 2;
                    JUES_CUDA(cudaMemsetAsync(slot, 0, sizeof(double), m->stream));
                    all_reduce_sum(m, slot, 1);
                    all_gather_inplace(m, m->red_dev, 1);
                    std::vector<size_t> off((size_t)m->nranks), cnt((size_t)m->nranks, 1);
                    for (int d = 0; d < m->nranks; ++d) off[d] = (size_t)d;
                    all_to_all_v(m, m->red_dev, off.data(), cnt.data(), m->red_dev + m->nranks, off.data(), cnt.data());
                    JUES_CUDA(cudaStreamSynchronize(m->stream));
                    return 0;
                } catch (const Error& e) {
                    m->last_error = e.what();
                    return e.code;
                }
            });
            if (rc != 0) {
                const std::string msg = lead->last_error;
                lead->group = nullptr;
                return fail(rc, msg);
            }
        }
        *out = lead;
        return JUES_B200_OK;
    } catch (const Error& e) {
        return fail(e.code, e.what());
    }
}

extern "C" int jues_b200_group_size(jues_ctx* ctx) {
    if (!ctx) return JUES_B200_EINVAL;
    return ctx->group ? (int)ctx->group->size() : 1;
}

extern "C" int jues_b200_nccl_unique_id(unsigned char id_out[128]) {
    if (!id_out) return JUES_B200_EINVAL;
    try {
        load_nccl();
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        ncclUniqueId id;
        nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(id_out, &id, 128);
        return JUES_B200_OK;
    } catch (const Error& e) {
        g_init_error = e.what();
        return e.code;
    }
}

extern "C" int jues_b200_init_dist(jues_ctx* ctx, int rank, int nranks, const unsigned char id[128]) {
    JUES_API_BEGIN(ctx)
    JUES_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    dist_teardown(ctx);
    ctx->rank = 0;
    ctx->nranks = 1;
    if (nranks > 1) {
        JUES_REQUIRE(id != nullptr, "null NCCL unique id");
        load_nccl();
        ncclUniqueId uid;
        memcpy(&uid, id, 128);
        ncclComm_t comm;
        nccl_check(g_nccl.CommInitRank(&comm, nranks, uid, rank), "ncclCommInitRank");
        // rank / nranks are published only once the communicator exists (a failed init leaves a
        // consistent single-rank context behind)
        ctx->nccl_comm = comm;
        ctx->rank = rank;
        ctx->nranks = nranks;
        // first collective sets up the NVLink connections (seconds): do it here, not inside a timed call
        double* slot = ctx->red_dev + ctx->red_cap - 2;
        JUES_CUDA(cudaMemsetAsync(slot, 0, sizeof(double), ctx->stream));
        all_reduce_sum(ctx, slot, 1);
        all_gather_inplace(ctx, ctx->red_dev, 1);
        {   // ... and the point-to-point connections of the transform's exchange
            std::vector<size_t> off((size_t)nranks), cnt((size_t)nranks, 1);
            for (int d = 0; d < nranks; ++d) off[d] = (size_t)d;
            all_to_all_v(ctx, ctx->red_dev, off.data(), cnt.data(), ctx->red_dev + nranks, off.data(), cnt.data());
        }
        JUES_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    JUES_API_END(ctx)
}
