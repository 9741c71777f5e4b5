// host_state.hpp — host-side state behind the opaque handles of include/mcphylo_b200.h: contexts
// (single device, multi-device group, or one rank of a multi-process group), resident alignments
// (whole, or one shard per device), and the cached evaluation plans.
// Part of libmcphylo_b200.so; included by mcphylo_b200.cu only.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mcphylo_b200.h"
#include "schedule.hpp"
#include "device_layout.cuh"
#include "smem_layout.cuh"
#include "kernel_api.hpp"
#include "nccl_dyn.hpp"

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct mcp_alignment {
    int K = 0;
    long long S = 0, stride = 0;
    int n_leaves = 0;
    unsigned char* d_codes = nullptr;
    std::vector<int32_t> leaf_nums;
    unsigned long long id = 0;
    mcp_ctx* owner = nullptr;
    // Re-uploads (mcp_alignment_update_codes) run on the context's copy stream so that they overlap
    // evaluations of OTHER alignments; these order them against the evaluations of THIS one.
    cudaEvent_t ev_uploaded = nullptr;
    cudaEvent_t ev_read_done = nullptr;      // recorded after every walk that read d_codes, once `streamed`
    mutable bool upload_pending = false;     // an upload has been enqueued that no evaluation has waited for yet
    mutable bool read_since_upload = false;  // an evaluation reading d_codes was enqueued after the last upload
    mutable bool streamed = false;           // has been re-uploaded at least once: evaluations record ev_read_done
    // Alignment of a multi-device context: one shard per device (contiguous site ranges,
    // mcp_shard_bounds); d_codes stays null and the fields above describe the whole alignment.
    std::vector<mcp_alignment*> shards;
    std::vector<long long> shard_lo;
};

// One cached evaluation plan: everything that depends only on (alignments, topologies, K, R, want_grad)
// and the launch knobs -- the schedules, the launch shape, the tile / accumulator-row assignment and the
// topology block resident on the device.  A context keeps a few (least recently used is recycled), so a
// sampler that alternates between topologies (NNI attempts and their rejection) or between site blocks
// of a streamed alignment rebuilds nothing.
struct Plan {
    bool valid = false;
    bool uploaded = false;          // d_topo holds this plan's topology block
    unsigned long long stamp = 0;   // last use, for recycling
    struct TreeSig {
        unsigned long long aln_id;
        int NN;
        std::vector<int32_t> po, pa;
    };
    std::vector<TreeSig> sig;
    int want_grad = -1, K = 0, R = 0;

    std::vector<mcpdev::TreeDev> trees;
    std::vector<mcp::Schedule> scheds;
    int block = 0, cpt = 1;
    bool level_mode = false, smem_scratch = false, acc_global = false, mma = false;
    bool model_grad = false;        // planned for the runtime-K kernel with the model-gradient moments (any K)
    int n_tiles = 0, grid = 0, n_rows = 0, n_slots = 0, n_stack = 0, max_br = 0, max_rows = 1;
    long long total_out = 0, total_dyn = 0, total_btab = 0, scratch_per_cta = 0, row_stride = 0;
    size_t smem_bytes = 0, topo_bytes = 0, off_trees = 0, off_ops = 0, off_rowbase = 0, off_levels = 0, off_fetch = 0;
    // operand ring of the gradient pass (smem_layout.cuh): usable for this plan (the launch still falls back to the
    // plain kernel for decompositions without a null eigenvalue or batches with several models) and the shared
    // memory that variant needs
    bool ring = false;
    size_t smem_ring = 0;
    DevBuf d_topo;
    PinBuf h_topo;
};

constexpr int MCP_PLAN_SLOTS = 8;    // cached plans per context
constexpr int MCP_STAGE_SLOTS = 4;   // pinned parameter staging buffers in flight per context

// Alignment held in host memory and uploaded block by block during each evaluation (mcp_eval_streamed).
struct StreamSet {
    int K = 0, n_leaves = 0, R = 0, NN = 0, want_grad = -1;
    long long S = 0;
    std::vector<int32_t> leaf_nums;
    // ev: device timeline of the block in the last call -- transfer begin / end (copy stream), evaluation
    // enqueued / walk begin / walk end (evaluation stream); read by mcp_stream_timeline
    struct Block { mcp_alignment* aln; long long lo, hi; cudaEvent_t ev[5] = {}; };
    std::vector<Block> blocks;      // of this device's site range, in evaluation order
    // Fused mode (one walk launch over the whole range, started before the data has arrived; tiles wait
    // for per-site-group ready flags the copy stream sets): blocks holds the single whole-range alignment.
    bool fused = false;
    std::vector<std::pair<long long, long long>> chunks;   // transfer units, sites relative to the range start
    DevBuf d_ctl;                   // [0] tile ticket, [1] error flag, [STREAM_CTL_WORDS ..) ready flag per site group
    PinBuf h_epoch, h_err;          // epoch words the flag copies read; error word read back
    unsigned epoch = 0;
};
constexpr int STREAM_GROUP_SHIFT = 11;   // ready-flag granularity: 2048 sites
constexpr int STREAM_CTL_WORDS = 64;     // ticket / error words, padded to their own 256 bytes

// What eval_impl passes to the walk kernel for a streamed evaluation.
struct StreamFlags {
    const unsigned* flags;
    unsigned* ticket;
    unsigned* error;
    unsigned epoch;
    int shift;
};

// Host threads of a multi-device context: member g > 0 is driven by its own persistent thread (member 0
// by the caller), so the devices' work is enqueued concurrently -- issued from one thread, the last of 8
// GPUs would start ~0.3 ms after the first, 3 % of a cfg4 evaluation at 8 GPUs.
struct MemberPool {
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::vector<std::thread> threads;
    const std::function<int(int)>* fn = nullptr;
    std::vector<int> rc;
    unsigned long long generation = 0;
    int remaining = 0;
    bool stop = false;

    explicit MemberPool(int n_members) : rc(n_members, 0) {
        for (int g = 1; g < n_members; ++g) threads.emplace_back([this, g] { loop(g); });
    }
    ~MemberPool() {
        {
            std::lock_guard<std::mutex> lock(mu);
            stop = true;
        }
        cv_go.notify_all();
        for (auto& t : threads) t.join();
    }
    void loop(int g) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<int(int)>* f;
            {
                std::unique_lock<std::mutex> lock(mu);
                cv_go.wait(lock, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                f = fn;
            }
            const int r = (*f)(g);
            std::lock_guard<std::mutex> lock(mu);
            rc[g] = r;
            if (--remaining == 0) cv_done.notify_one();
        }
    }
    // runs f(g) for every member; the caller's thread takes member 0
    void run(const std::function<int(int)>& f) {
        {
            std::lock_guard<std::mutex> lock(mu);
            fn = &f;
            remaining = (int)threads.size();
            ++generation;
        }
        cv_go.notify_all();
        const int r0 = f(0);
        std::unique_lock<std::mutex> lock(mu);
        rc[0] = r0;
        cv_done.wait(lock, [&] { return remaining == 0; });
    }
};

struct mcp_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;     // alignment re-uploads (overlap with evaluations)
    cudaEvent_t ev_walk_done = nullptr;     // the walk kernel of the last evaluation has finished reading the codes
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_done = nullptr;          // everything enqueued by the last evaluation has finished
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;   // mcp_timer_start / mcp_timer_stop
    cudaEvent_t ev_walk_begin = nullptr, ev_walk_end = nullptr;   // borrowed per-block events (mcp_eval_streamed timeline)
    std::string error;
    bool pending_async = false;
    int opt_block = 0, opt_ctas_per_sm = 0, opt_cpt = 0;
    int opt_levels = -1;         // -1 automatic, 0 never, 1 whenever it fits
    int opt_smem_scratch = -1;   // -1 automatic, 0 off, 1 on when it fits
    int opt_acc_mode = -1;       // gradient accumulator of the walk: -1 automatic, 0 shared memory, 1 global memory (RED)
    int opt_ring = -1;           // gradient pass: operand ring (-1 automatic = on where supported, 0 off, 1 on)
    int opt_mma = -1;            // K > 6: FP64 tensor-core walk (-1 / 1) or the runtime-K fallback kernel (0)
    int opt_cherry = -1;         // gradient pass recomputes cherries from their leaves (-1 / 1) or re-reads stored copies (0)
    int opt_timing = 1;          // per-evaluation CUDA timing events behind mcp_get_stats (0: not recorded, times read 0)
    int opt_dynamic = -1;        // resident single-tree walk: tiles by atomic ticket in site order (1) or static ranges (0)
    unsigned long long next_aln_id = 1, clock = 0;

    DevBuf d_dyn, d_btab, d_scratch, d_rows, d_rows_ll, d_out, d_counter, d_part, d_ticket;
    DevBuf d_mg;                            // model-gradient moments of the last mcp_eval_model_gradient
    PinBuf h_out, h_mg;
    // parameter staging ring: an evaluation fills slot `stage_next`, the copy to the device is
    // asynchronous, and the slot is reused only after its event has passed
    PinBuf h_dyn[MCP_STAGE_SLOTS], h_model[MCP_STAGE_SLOTS];
    cudaEvent_t ev_staged[MCP_STAGE_SLOTS] = {};
    bool staged_pending[MCP_STAGE_SLOTS] = {};
    int stage_next = 0;

    std::vector<std::unique_ptr<Plan>> plans;
    Plan* last_plan = nullptr;
    std::unique_ptr<StreamSet> stream_set;
    const StreamFlags* sf = nullptr;        // set by mcp_eval_streamed around its fused launch

    // ---- more than one device behind this handle (mcp_create_multi) ----
    std::vector<mcp_ctx*> members;          // one single-device context per device; empty for a plain context
    int reduce_mode = MCP_REDUCE_AUTO;      // resolved at creation: NCCL, PEER or HOST
    std::vector<mcpnccl::comm_t> comms;     // NCCL: one communicator per member
    DevBuf d_gather;                        // PEER: [member][block][total_out] on device 0
    std::vector<cudaEvent_t> ev_member;     // PEER: member g has written its part
    std::vector<PinBuf> h_member_out;       // HOST: one pinned result vector per member
    std::unique_ptr<MemberPool> pool;       // one host thread per member beyond the first
    // ---- one rank of a multi-process group (mcp_create_rank) ----
    mcpnccl::comm_t rank_comm = nullptr;
    int n_ranks = 1, rank = 0;

    mcp_stats stats{};
};

namespace {

thread_local std::string g_error;

int fail(mcp_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    else g_error = buf;
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return fail(ctx, MCP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                     \
    } while (0)

int ensure_dev(mcp_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (ctx->pending_async) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->pending_async = false;
    }
    if (b.p) CUDA_TRY(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(ctx, MCP_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return 0;
}
int ensure_pin(mcp_ctx* ctx, PinBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CUDA_TRY(ctx, cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    CUDA_TRY(ctx, cudaMallocHost(&b.p, want));
    b.cap = want;
    return 0;
}
void free_dev(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}
void free_pin(PinBuf& b) {
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
}

}  // namespace
