// jues_b200 -- internal common definitions (context, errors, device buffers).
// Not part of the public C ABI (that is include/jues_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <iterator>
#include <algorithm>
#include <stdexcept>
#include <chrono>

#include "../../include/jues_b200.h"

namespace jues {

// ---------------------------------------------------------------------------------------------
// Errors: internal code throws jues::Error; the extern "C" layer catches and converts to a
// negative status + message (no exception ever crosses the ABI).
// ---------------------------------------------------------------------------------------------
struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define JUES_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            char buf__[512];                                                                    \
            snprintf(buf__, sizeof buf__, "CUDA error %s (%d) at %s:%d: %s", cudaGetErrorName(e__), \
                     (int)e__, __FILE__, __LINE__, cudaGetErrorString(e__));                    \
            throw ::jues::Error(e__ == cudaErrorMemoryAllocation ? JUES_B200_ENOMEM             \
                                                                 : JUES_B200_ECUDA, buf__);     \
        }                                                                                       \
    } while (0)

#define JUES_REQUIRE(cond, msg)                                                                 \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            char buf__[512];                                                                    \
            snprintf(buf__, sizeof buf__, "invalid argument: %s  [%s] at %s:%d", msg, #cond,    \
                     __FILE__, __LINE__);                                                       \
            throw ::jues::Error(JUES_B200_EINVAL, buf__);                                       \
        }                                                                                       \
    } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// ---------------------------------------------------------------------------------------------
// Per-phase statistics the benchmark reads back (CUDA-event time on the library's stream,
// floating-point operations actually issued by the GEMM kernels, launch counts).
// ---------------------------------------------------------------------------------------------
struct Stats {
    double gemm_flops = 0;       // 2*M*N*K*batch summed over launches (padded dims)
    long long gemm_launches = 0;
    long long aux_launches = 0;  // permute / elementwise / reduction kernels
    long long graph_launches = 0; // sweeps replayed from a CUDA graph (their kernels are counted above)
    long long collectives = 0;   // NCCL calls
    double collective_bytes = 0; // bytes received per rank in collectives
    void reset() { *this = Stats(); }
};

// One device block that serves the small stream-ordered temporaries of a sweep (~300 per sweep): a
// host-side first-fit free list, no driver call per allocation.  All work of a context is on one stream, so
// a block can be handed out again as soon as its owner released it.  Kernels captured into a CUDA graph
// see arena addresses, which stay valid for as long as the arena lives (it outlives the graphs).
struct Arena {
    double* base = nullptr;
    size_t bytes = 0;
    std::map<size_t, size_t> free_by_off;   // offset -> length of free runs, coalesced on release
    size_t live = 0, peak = 0;
    long long misses = 0;                   // requests that did not fit (served by the pool / block cache)
    bool has(const void* p) const {
        return base && (const char*)p >= (const char*)base && (const char*)p < (const char*)base + bytes;
    }
    void reset(double* b, size_t n) {
        base = b; bytes = n; free_by_off.clear(); live = peak = 0;
        if (n) free_by_off[0] = n;
    }
    double* take(size_t n) {              // n: multiple of 256
        for (auto it = free_by_off.begin(); it != free_by_off.end(); ++it) {
            if (it->second < n) continue;
            const size_t off = it->first, len = it->second;
            free_by_off.erase(it);
            if (len > n) free_by_off[off + n] = len - n;
            live += n;
            if (live > peak) peak = live;
            return (double*)((char*)base + off);
        }
        ++misses;
        return nullptr;
    }
    void give(const void* p, size_t n) {
        size_t off = (size_t)((const char*)p - (const char*)base);
        live -= n;
        auto nx = free_by_off.lower_bound(off);
        if (nx != free_by_off.end() && off + n == nx->first) { n += nx->second; nx = free_by_off.erase(nx); }
        if (nx != free_by_off.begin()) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == off) { pv->second += n; return; }
        }
        free_by_off[off] = n;
    }
};

}  // namespace jues

// The opaque context of the C ABI.
struct jues_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;   // second stream: exchanges / copies that overlap compute
    jues::PFN_encodeTiled encode = nullptr;
    std::string last_error;
    jues::Stats stats;
    // scratch for device-side reductions (partials) and small host mirror
    double* red_dev = nullptr;   // [red_cap]
    double* red_host = nullptr;  // pinned, [red_cap]
    size_t red_cap = 0;
    // timing records: name -> milliseconds (accumulated), filled by Timer
    std::vector<std::pair<std::string, float>> timings;
    struct PendingTimer { size_t slot; cudaEvent_t e0, e1; };
    std::vector<PendingTimer> pending;   // recorded, not yet resolved (see Timer)
    size_t bytes_allocated = 0;
    size_t bytes_peak = 0;
    std::multimap<size_t, double*> big_free;   // cached cudaMalloc blocks (>= 64 MB) by size
    size_t big_cached_bytes = 0;
    double alloc_host_s = 0.0;   // host wall-clock spent inside device allocation calls (diagnostic)
    long long alloc_calls = 0;
    // small (pool-sized) allocations: live / peak bytes since the last reset (sizes the sweep arena), and the
    // arena itself while a coupled-cluster driver has one installed
    size_t small_live = 0, small_peak = 0;
    // every allocation made while `measure` is on (the first sweep), except those made under MeasurePause
    // (operand copies that persist for the whole calculation): what a sweep's temporaries need at once
    bool measure = false;
    size_t temp_live = 0, temp_peak = 0;
    long long big_allocs = 0;     // allocations of big_bytes or more NOT served by the arena (a sweep that makes any is not captured)
    jues::Arena* arena = nullptr;          // where DBuf takes sweep temporaries from right now (nullptr: pool / block cache)
    // every arena that may own live blocks (release looks the owner up by address): the main one and the one
    // of the side branch of a sweep, which is issued on the second stream (the free lists assume that the blocks
    // of one arena are used in ONE stream's order, so each concurrent branch has its own)
    jues::Arena* arena_main = nullptr;
    jues::Arena* arena_side = nullptr;
    bool no_big_cache = false;             // side branch: a miss must not take a cached cudaMalloc block (not stream-ordered)
    // multi-GPU (one process per GPU): rank / world size and an NCCL communicator (opaque here)
    int rank = 0;
    int nranks = 1;
    void* nccl_comm = nullptr;
    void* nccl_lib = nullptr;
    void* perm_cache = nullptr;   // jues::PermCache of the running calculation (contract.h)
    // single-process multi-GPU (jues_b200_init_multi): the leader (rank 0) context lists every member,
    // itself included; an entry point called on the leader runs on all members, one host thread per GPU
    std::vector<jues_ctx*>* group = nullptr;   // non-null on the leader only
    jues_ctx* leader = nullptr;                // non-null on members (the leader points to itself)
    bool in_group_call = false;
    // diagnostics (environment, read once in jues_b200_init): JUES_B200_BIG_MB moves the pool / cudaMalloc
    // threshold of DBuf, JUES_B200_SYNC_COMM=1 drains the stream around every collective
    size_t big_bytes = size_t(64) << 20;
    bool sync_comm = false;
    // 0: coarse phases only; 1: cc.part.* / cc.comm.* / tei.* regions (JUES_B200_TRACE=1 or
    // jues_b200_set_trace); 2: additionally one event pair around EVERY DGEMM launch, recorded as
    // "gemm MxNxKxbatch" (in-situ kernel durations for the roofline; sweeps run eagerly while tracing)
    int trace = 0;
    std::set<const void*> smem_attr_done;   // kernels whose dynamic shared-memory limit was raised on this device
    // per-sweep amplitude capture (tests)
    jues_b200_amp_cb amp_cb = nullptr;
    void* amp_user = nullptr;
};

// device-resident rank-4 tensor handle of the C ABI (DiskFourTensor replacement)
struct jues_t4 {
    jues_ctx* ctx = nullptr;
    int64_t d[4] = {0, 0, 0, 0};   // logical extents
    int64_t dp[4] = {0, 0, 0, 0};  // padded (even) extents of the device allocation
    double* p = nullptr;
    size_t bytes = 0;
    bool virtual_synth = false;      // no storage: slabs generated on demand by the counter-based generator
    unsigned long long seed = 0;
    double scale = 0.0;
};

namespace jues {

// ---------------------------------------------------------------------------------------------
// Device buffer with RAII; all allocations go through the context for accounting.
// ---------------------------------------------------------------------------------------------
struct DBuf {
    jues_ctx* ctx = nullptr;
    double* p = nullptr;
    size_t n = 0;  // elements
    DBuf() {}
    DBuf(jues_ctx* c, size_t n_) { alloc(c, n_); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept { *this = std::move(o); }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) {
            release();
            ctx = o.ctx; p = o.p; n = o.n; cap = o.cap; from_big = o.from_big; from_arena = o.from_arena; counted = o.counted;
            o.p = nullptr; o.n = 0; o.cap = 0; o.from_big = false; o.from_arena = false; o.counted = false;
        }
        return *this;
    }
    ~DBuf() { release(); }
    // Small blocks come from the stream-ordered pool (no device synchronisation).  Blocks of
    // kBigBytes and more come from cudaMalloc and are cached per context by size: growing the
    // stream-ordered pool costs ~5 ms/GB at best and seconds once multi-GB blocks of many different
    // sizes fragment it (measured: 8 s of host time in 107 allocations of one large transform).
    // Re-using a cached block right away is safe because all work of a context is on one stream.
    static constexpr size_t kBigBytes = size_t(64) << 20;
    size_t cap = 0;  // bytes actually held (big blocks may be slightly larger than requested)
    bool from_big = false;   // block came from cudaMalloc / the context's big-block cache
    bool from_arena = false; // block came from the context's sweep arena
    bool counted = false;    // included in ctx->temp_live
    void alloc(jues_ctx* c, size_t n_) {
        release();
        ctx = c;
        n = n_;
        size_t bytes = (n_ ? n_ : 1) * sizeof(double);
        // keep every allocation a multiple of 256 B so that TMA boxes that overhang the logical
        // end of a tensor never leave the allocation's page
        bytes = (bytes + 255) & ~size_t(255);
        const auto t0__ = std::chrono::steady_clock::now();
        cudaError_t e = cudaSuccess;
        cap = bytes;
        from_big = bytes >= c->big_bytes && !c->no_big_cache;
        from_arena = false;
        counted = c->measure;
        if (counted) {
            c->temp_live += bytes;
            if (c->temp_live > c->temp_peak) c->temp_peak = c->temp_live;
        }
        if (c->arena) {
            p = c->arena->take(bytes);
            if (p) {
                from_arena = true;
                from_big = false;
                c->alloc_calls++;
                c->bytes_allocated += cap;
                if (c->bytes_allocated > c->bytes_peak) c->bytes_peak = c->bytes_allocated;
                return;
            }
        }
        if (from_big) {
            c->big_allocs++;
            auto it = c->big_free.lower_bound(bytes);
            if (it != c->big_free.end() && it->first <= bytes + bytes / 8) {
                p = it->second;
                cap = it->first;
                c->big_cached_bytes -= cap;
                c->big_free.erase(it);
            } else {
                e = cudaMalloc((void**)&p, bytes);
                if (e == cudaErrorMemoryAllocation) {   // give the cache back to the driver and retry
                    cudaGetLastError();
                    for (auto& kv : c->big_free) cudaFree(kv.second);
                    c->big_free.clear();
                    c->big_cached_bytes = 0;
                    e = cudaMalloc((void**)&p, bytes);
                }
            }
        } else {
            e = cudaMallocAsync((void**)&p, bytes, c->stream);
        }
        c->alloc_host_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0__).count();
        c->alloc_calls++;
        if (e != cudaSuccess) {
            p = nullptr;
            char buf[256];
            snprintf(buf, sizeof buf, "device allocation of %.3f GB failed (%s); %.3f GB already held",
                     bytes / 1e9, cudaGetErrorString(e), c->bytes_allocated / 1e9);
            cudaGetLastError();
            throw Error(JUES_B200_ENOMEM, buf);
        }
        c->bytes_allocated += cap;
        if (c->bytes_allocated > c->bytes_peak) c->bytes_peak = c->bytes_allocated;
        if (!from_big) {
            c->small_live += cap;
            if (c->small_live > c->small_peak) c->small_peak = c->small_live;
        }
    }
    void release() {
        if (p) {
            if (counted) { ctx->temp_live -= std::min(ctx->temp_live, cap); counted = false; }
            if (from_arena) {
                // back to the arena that owns the address (if it is gone, so is its block)
                if (ctx->arena_main && ctx->arena_main->has(p)) ctx->arena_main->give(p, cap);
                else if (ctx->arena_side && ctx->arena_side->has(p)) ctx->arena_side->give(p, cap);
                else if (ctx->arena && ctx->arena->has(p)) ctx->arena->give(p, cap);
            } else if (from_big) {
                ctx->big_free.emplace(cap, p);
                ctx->big_cached_bytes += cap;
            } else {
                cudaFreeAsync(p, ctx->stream);
                ctx->small_live -= cap;
            }
            ctx->bytes_allocated -= cap;
            p = nullptr;
            n = 0;
            cap = 0;
            from_big = false;
            from_arena = false;
        }
    }
    void zero() { JUES_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), ctx->stream)); }
};

// give every cached large block back to the driver
inline void flush_big_cache(jues_ctx* c) {
    for (auto& kv : c->big_free) cudaFree(kv.second);
    c->big_free.clear();
    c->big_cached_bytes = 0;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// CUDA-event timer writing into ctx->timings.
// CUDA-event timer writing into ctx->timings.  Non-blocking: the events are recorded on the stream and
// resolved (one synchronisation) when the phases are read, so timing a phase never stalls the host
// thread that is feeding the GPU.
struct Timer {
    jues_ctx* ctx;
    std::string name;
    cudaEvent_t e0, e1;
    bool open = true;
    Timer(jues_ctx* c, const std::string& n) : ctx(c), name(n) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, c->stream);
    }
    void stop() {
        if (open) {
            cudaEventRecord(e1, ctx->stream);
            ctx->timings.emplace_back(name, -1.0f);
            ctx->pending.push_back({ctx->timings.size() - 1, e0, e1});
            open = false;
        }
    }
    ~Timer() { stop(); }
};

inline void resolve_timers(jues_ctx* ctx) {
    if (ctx->pending.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& pt : ctx->pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, pt.e0, pt.e1);
        if (pt.slot < ctx->timings.size()) ctx->timings[pt.slot].second = ms;
        cudaEventDestroy(pt.e0);
        cudaEventDestroy(pt.e1);
    }
    ctx->pending.clear();
}

// Run a section on another stream: every helper launches on ctx->stream, so the scope swaps it.
// No stream-ordered allocation may happen inside (DBuf frees are ordered on the stream they see).
struct StreamScope {
    jues_ctx* ctx;
    cudaStream_t saved;
    StreamScope(jues_ctx* c, cudaStream_t s) : ctx(c), saved(c->stream) { c->stream = s; }
    ~StreamScope() { ctx->stream = saved; }
};

// Allocations that persist beyond the sweep being measured do not count as its temporaries.
struct MeasurePause {
    jues_ctx* ctx;
    bool saved;
    explicit MeasurePause(jues_ctx* c) : ctx(c), saved(c->measure) { c->measure = false; }
    ~MeasurePause() { ctx->measure = saved; }
};

// Allocations that outlive a sweep must not come from the sweep arena.
struct ArenaPause {
    jues_ctx* ctx;
    Arena* saved;
    explicit ArenaPause(jues_ctx* c) : ctx(c), saved(c->arena) { c->arena = nullptr; }
    ~ArenaPause() { ctx->arena = saved; }
};

// JUES_B200_TRACE=1: fine-grained CUDA-event timings
struct TraceTimer {
    Timer* t = nullptr;
    TraceTimer(jues_ctx* ctx, const char* name) {
        if (ctx->trace > 0) t = new Timer(ctx, name);
    }
    ~TraceTimer() { delete t; }
};

}  // namespace jues
