/*
 * mcphylo_b200.h — C ABI of libmcphylo_b200.so
 *
 * B200 (sm_100a) implementation of the one hot path of MCPhylo.jl: the Felsenstein-pruning
 * log-likelihood of a PhyloDist and its analytic branch-length gradient.  These entry points
 * are what a Julia `ccall` (or any FFI) binds to replace
 *
 *     logpdf(d::PhyloDist, x)        /root/reference/src/distributions/Phylodist.jl:107-122
 *     gradlogpdf(d::PhyloDist, x)    /root/reference/src/distributions/Phylodist.jl:124-138
 *     logpdf / __logpdf(d::MultiplePhyloDist, x)   Phylodist.jl:281-297
 *
 * i.e. everything from `my_repeat`/`parallel_transition_prob` down through
 * `FelsensteinFunction` (src/Likelihood/LikelihoodCalculator_Node.jl:3-114).  The substitution
 * model's eigendecomposition (U, D, Uinv, mu) and the tree traversal stay on the caller's side
 * and arrive here as flat arrays.
 *
 * Conventions
 *   - Plain C types only.  Every pointer is HOST memory owned by the caller and only read (or,
 *     for outputs, written) during the call, unless a parameter is explicitly named d_* (device).
 *   - Matrices are column-major (Julia layout).  Node numbers are 1-based `node.num`; the root
 *     must carry the largest number NN; arrays "indexed by num" use position num-1.
 *   - Every function returns 0 on success and a negative mcp_status on failure; the message is
 *     available from mcp_last_error().  Nothing here calls abort()/exit() or throws across the
 *     boundary.  There is no CPU fallback: without a usable CUDA device mcp_create fails.
 *   - A context is thread-compatible, not thread-safe: one call at a time per context.
 *   - Semantics follow the reference exactly: rate categories are NOT mixed
 *     (logL = sum_r sum_s log L(s | r), VectorizedFunctions.jl:76-87), branch lengths are not
 *     clamped, multifurcations and single-child nodes are allowed.
 */
#ifndef MCPHYLO_B200_H
#define MCPHYLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCP_ABI_VERSION 2

typedef struct mcp_ctx mcp_ctx;             /* one per (process, GPU) -- or per (process, set of GPUs), mcp_create_multi */
typedef struct mcp_alignment mcp_alignment; /* leaf data resident on the GPU(s) of its context */

typedef enum mcp_status {
    MCP_OK = 0,
    MCP_ERR_ARG = -1,         /* bad argument / malformed tree */
    MCP_ERR_CUDA = -2,        /* CUDA runtime error (no device, OOM, launch failure) */
    MCP_ERR_UNSUPPORTED = -3, /* e.g. a state count K without a compiled kernel */
    MCP_ERR_DATA = -4         /* leaf data that is neither one-hot nor all-ones */
} mcp_status;

int mcp_abi_version(void);

/* Message of the last failure on this context (ctx == NULL: last failure of a call that had no
 * context, e.g. mcp_create).  The pointer stays valid until the next call on the context. */
const char *mcp_last_error(const mcp_ctx *ctx);

/* Binds a context to CUDA device `device`, creates its stream and staging buffers. */
int mcp_create(mcp_ctx **out, int device);
int mcp_destroy(mcp_ctx *ctx);

/*
 * Multi-GPU, one host process (what a Julia session is): ONE context that owns n_dev GPUs of the box.
 * Every entry point below takes such a context exactly like a single-device one:
 *   - mcp_alignment_from_codes / _from_dense split the SITE axis: device g of G holds the contiguous
 *     columns [g*ceil(S/G), min(S, (g+1)*ceil(S/G))) of every leaf row (mcp_shard_bounds), for all nodes
 *     and rate categories; tree, branch tables and model are replicated.
 *   - mcp_eval / mcp_eval_posterior / mcp_eval_batch / mcp_eval_streamed run the fused walk on every
 *     device at once (one host thread and one stream per device) and sum the per-device vectors
 *     [logL, grad[1..NN-1]] -- the ONLY exchange of the path -- by
 *       MCP_REDUCE_NCCL  one grouped ncclAllReduce(count = NN, ncclDouble, ncclSum) on the devices'
 *                        evaluation streams (communicators from ncclCommInitAll; NCCL is resolved at run
 *                        time with dlopen("libnccl.so.2"), or the library MCPHYLO_B200_NCCL names),
 *       MCP_REDUCE_PEER  no library: every device's final-reduction kernel stores its vector straight
 *                        into a gather buffer on device 0 over NVLink peer access, and one small kernel
 *                        on device 0 adds the G vectors in device order into pinned host memory
 *                        (bit-reproducible; also works with the same device listed more than once),
 *       MCP_REDUCE_HOST  each device copies its vector to pinned host memory, the host adds them in
 *                        device order (bit-reproducible; the comparison SURVEY.md 8e allows),
 *       MCP_REDUCE_AUTO  NCCL when it can be loaded and the devices are distinct, else PEER, else HOST.
 *     A branch-length prior (mcp_eval_posterior) is added once, not per device.
 *   - the tuning knobs apply to every device; mcp_get_stats reports the slowest device's times and
 *     summed bytes / launches, mcp_get_stats_member one device's.
 *   - mcp_eval_device and mcp_set_stream are single-device calls and fail with MCP_ERR_UNSUPPORTED.
 * This is the entry a reference-side `gradlogpdf(d::PhyloDist, x)`
 * (/root/reference/src/distributions/Phylodist.jl:124-138) binds to reach 2-8 GPUs through `ccall`.
 */
enum mcp_reduce_mode { MCP_REDUCE_AUTO = 0, MCP_REDUCE_NCCL = 1, MCP_REDUCE_PEER = 2, MCP_REDUCE_HOST = 3 };
int mcp_create_multi(mcp_ctx **out, int n_dev, const int *dev_ids, int reduce_mode);
/* Devices behind a context (1 for mcp_create / mcp_create_rank) and the reduction it resolved to. */
int mcp_device_count(const mcp_ctx *ctx);
int mcp_reduce_mode(const mcp_ctx *ctx);
/* The site range [*lo, *hi) shard `shard` of `n_shards` owns (the rule mcp_create_multi applies). */
int mcp_shard_bounds(int64_t S, int n_shards, int shard, int64_t *lo, int64_t *hi);

/*
 * Multi-GPU, one host process PER GPU (MPI-style launchers, torchrun): rank 0 obtains a 128-byte NCCL
 * unique id, the launcher's own channel carries it to the other ranks, and every rank creates a context on
 * its device.  The rank passes ITS shard of the alignment (mcp_shard_bounds) to mcp_alignment_from_codes;
 * mcp_eval / mcp_eval_posterior / mcp_eval_batch then all-reduce [logL, grad] over the ranks on the
 * evaluation stream (ncclAllReduce, communicator from ncclCommInitRank) before the result is read back,
 * so every rank returns the full-alignment result.  Collective: all ranks must make the same calls in
 * the same order.  A branch-length prior is added by rank 0 only.
 */
int mcp_nccl_unique_id(void *id128);
int mcp_create_rank(mcp_ctx **out, int device, int n_ranks, int rank, const void *id128);

/* Optional: run on a caller-provided cudaStream_t (cast to void*) instead of the context's own
 * non-blocking stream.  NULL selects the CUDA default stream (that is what e.g. torch's default
 * stream is).  mcp_use_own_stream goes back to the internal stream. */
int mcp_set_stream(mcp_ctx *ctx, void *cuda_stream);
int mcp_use_own_stream(mcp_ctx *ctx);
/* Blocks until everything enqueued by this context (mcp_eval_device, mcp_alignment_update_codes)
 * has finished. */
int mcp_synchronize(mcp_ctx *ctx);

/*
 * Leaf data.  Replaces the dense one-hot array `x[:, :, leaf.num]` the reference re-expands on
 * every call (my_repeat, VectorizedFunctions.jl:13-29) by 1-byte state codes uploaded once.
 *
 * from_codes: codes is (n_leaves, S) row-major; code k < K is state k (one-hot column), any code
 *             >= K is gap/missing (all-ones column, Parser.jl:72-74).  Row i belongs to the leaf
 *             whose node number is leaf_nums[i].
 * from_dense: x is the reference's own array, (K, S, NN) column-major Float64; for every leaf in
 *             leaf_nums each column must be one-hot or all ones (what datafortree produces),
 *             otherwise MCP_ERR_DATA.
 * An alignment belongs to the context that created it (device memory, stream ordering): passing it to
 * a call on another context fails with MCP_ERR_ARG; destroy it before the context.
 */
int mcp_alignment_from_codes(mcp_ctx *ctx, const uint8_t *codes, int K, int64_t S,
                             const int32_t *leaf_nums, int n_leaves, mcp_alignment **out);
int mcp_alignment_from_dense(mcp_ctx *ctx, const double *x, int K, int64_t S, int NN,
                             const int32_t *leaf_nums, int n_leaves, mcp_alignment **out);
/* Re-upload the codes of an existing alignment (same K, S, leaf_nums) from host memory,
 * asynchronously on the context's COPY stream: the transfer overlaps evaluations of other
 * alignments that are already enqueued (site-block pipelining, dist.PipelinedEvaluator), waits for
 * evaluations of this alignment enqueued earlier, and evaluations of this alignment enqueued
 * afterwards wait for it.  Use pinned host memory; `codes` must stay valid until an evaluation of
 * this alignment has completed or mcp_synchronize has returned.  The cached schedule stays valid. */
int mcp_alignment_update_codes(mcp_ctx *ctx, mcp_alignment *aln, const uint8_t *codes);
int mcp_alignment_destroy(mcp_ctx *ctx, mcp_alignment *aln);

/*
 * One evaluation = logpdf (want_grad == 0) or gradlogpdf (want_grad != 0) of one PhyloDist.
 *
 *   NN             number of tree nodes
 *   postorder_num  NN node numbers in post_order(tree) order (children in stored order before
 *                  their mother, root last)
 *   parent_num     NN, indexed by num: number of the mother, 0 for the root
 *   blv            NN-1, indexed by num: get_branchlength_vector(tree)
 *   U, D, Uinv, mu what d.substitution_model(base_freq, substitution_rates) returns
 *                  (K x K col-major, K, K x K col-major, scalar).  Eigenvalues may come in any
 *                  order (LAPACK's); the library works on a permuted copy with the null eigenvalue
 *                  of a rate matrix last and skips that component in its kernels.  A
 *                  decomposition without a null eigenvalue is accepted too (full-K kernels).
 *   rates, R       d.rates
 *   pi             d.base_freq (K)
 *   ll_out         1 double
 *   grad_out       NN-1 doubles indexed by num (d logL / d blv[num]); may be NULL if !want_grad
 */
int mcp_eval(mcp_ctx *ctx, const mcp_alignment *aln, int NN, const int32_t *postorder_num,
             const int32_t *parent_num, const double *blv, const double *U, const double *D,
             const double *Uinv, double mu, const double *rates, int R, const double *pi,
             int want_grad, double *ll_out, double *grad_out);

/*
 * Likelihood + branch-length prior in one call: the body of logpdfgrad!(::Type{provided}, ...)
 * (/root/reference/src/samplers/sampler.jl:172-190), which adds gradlogpdf of the tree's length
 * prior (/root/reference/src/Likelihood/Prior.jl:1-57, evaluated there through Zygote on every
 * leapfrog) to the PhyloDist gradient.  Here the prior is folded into the final reduction on the
 * device.  prior_kind / prior_params:
 *   MCP_PRIOR_NONE                 (UniformBranchLength, Prior.jl:59-66)  params ignored
 *   MCP_PRIOR_EXPONENTIAL          exponentialBL(scale)                   params = {scale}
 *   MCP_PRIOR_COMPOUND_DIRICHLET   CompoundDirichlet(alpha, a, beta, c)   params = {alpha, a, beta, c}
 *                                  (src/distributions/TreeDistribution.jl:23-39; a branch is
 *                                  "internal" when the node below it has children)
 * lp_out receives logL + log prior; grad_out (NN-1 doubles, or NULL for the value only) receives
 * the summed gradient indexed by num-1.  No support check is made (the reference does that in
 * logpdf_sub / insupport, Prior.jl:85-92, before it gets here): non-positive branch lengths give
 * NaN / -Inf from the logarithms exactly as they would in the reference's formula.
 */
enum mcp_prior_kind { MCP_PRIOR_NONE = 0, MCP_PRIOR_EXPONENTIAL = 1, MCP_PRIOR_COMPOUND_DIRICHLET = 2 };
int mcp_eval_posterior(mcp_ctx *ctx, const mcp_alignment *aln, int NN, const int32_t *postorder_num,
                       const int32_t *parent_num, const double *blv, const double *U, const double *D,
                       const double *Uinv, double mu, const double *rates, int R, const double *pi,
                       int prior_kind, const double *prior_params, double *lp_out, double *grad_out);

/*
 * logL, the branch-length gradient AND the gradient with respect to the rate-category multipliers
 * (rate_grad_out, R doubles: d logL / d rates[r]) in one call.  The reference has no such derivative -- it samples
 * the Gamma shape behind `rates` gradient-free (/root/reference/src/Likelihood/Rates.jl:11-38) -- but it falls
 * out of the quantities this path already computes: categories are not mixed and category r sees branch b as
 * t_b * rates[r], hence d logL / d rates[r] = (1 / rates[r]) * sum_b t_b * d logL_r / d t_b.  Evaluated as R
 * single-category evaluations on one cached plan, i.e. the same columns as one mcp_eval.  A host that wants
 * d logL / d alpha for rates = discrete_gamma_rates(alpha, alpha, R) applies the chain rule with
 * d rates / d alpha (mcphylo.jl_b200/rates.py: discrete_gamma_rates_jacobian).
 */
int mcp_eval_rate_gradient(mcp_ctx *ctx, const mcp_alignment *aln, int NN, const int32_t *postorder_num,
                           const int32_t *parent_num, const double *blv, const double *U, const double *D,
                           const double *Uinv, double mu, const double *rates, int R, const double *pi,
                           double *ll_out, double *grad_out, double *rate_grad_out);

/*
 * logL, the branch-length gradient AND the gradient with respect to the parameters of the substitution model
 * (base frequencies, exchangeabilities / free rates) in one evaluation.  The reference has no such derivative --
 * pi and the substitution rates are sampled gradient-free (e.g. SliceSimplex(:mypi),
 * /root/reference/src/samplers/tree_samplers.jl:49; the models are /root/reference/src/Likelihood/SubstitutionModels.jl:13-102)
 * -- it is the adjacent capability SURVEY.md 8f names: the partials the gradient pass already holds are reused.
 * Per branch b and rate category r the gradient pass accumulates the K x K moment matrix
 *     M[b][r][s][k] = sum over the columns of category r of  q_b[s] L_b[k] / den          (= d logL / d P_{b,r}[s][k])
 * (q_b: outer partial at the top of the branch, L_b: partial below it, den: the column likelihood) and the root vector
 *     W[s] = sum over all columns of  L_root[s] / (pi . L_root)                         (= d logL / d pi[s] at the root);
 * the host contracts them with d P_{b,r} / d theta_p, which follows from the eigen-decomposition the caller passes anyway:
 * with A = mu U diag(D) Uinv (so that P_{b,r} = exp(A t_b rates[r])) the caller supplies
 *   dA   K x K x n_par, column-major per parameter: d A / d theta_p  (derivative of the NORMALISED rate matrix, i.e. including
 *        the dependence of mu on the parameter)
 *   dpi  K x n_par column-major: d pi / d theta_p as seen by the ROOT term (1 in row s for "base frequency s", 0 for
 *        rates); NULL if no parameter enters the root distribution
 * and receives par_grad_out[p] = d logL / d theta_p (n_par doubles) and, if rate_grad_out != NULL, d logL / d rates[r]
 * (R doubles: the same quantity mcp_eval_rate_gradient returns, here from the moments of the one evaluation -- with
 * d rates / d alpha the Gamma-shape gradient; n_par may be 0 for a caller that only wants this).  mcphylo.jl_b200/substitution_models.py
 * (model_derivatives) and julia/MCPhyloB200.jl hold dA / dpi for Restriction, JC, GTR and freeK, and a Richardson-extrapolated
 * difference quotient for user-supplied model functions.
 * moments_out (optional, may be NULL): (NN-1) * R * K * K doubles M[b][r][s * K + k] followed by W[K].
 * Runs on the one-thread-per-column kernel of the large alphabets for every K (compile-time-K instantiations for K <= 6;
 * per op and child a warp forms its 32-column sum of outer products cooperatively in shared memory and adds it to M
 * with one atomic per entry), so it costs about 4 plain evaluations (cfg3: 6.1 ms vs 1.4 ms); on a multi-device
 * context every device evaluates its site shard and the host adds the parts.
 * mcp_model_gradient_contract is the host-only second half (no GPU needed): moments -> parameter gradient, and optionally
 * the branch gradient re-derived from the same moments (grad_check_out, n_branches doubles) as a consistency check.
 */
int mcp_eval_model_gradient(mcp_ctx *ctx, const mcp_alignment *aln, int NN, const int32_t *postorder_num,
                            const int32_t *parent_num, const double *blv, const double *U, const double *D,
                            const double *Uinv, double mu, const double *rates, int R, const double *pi,
                            int n_par, const double *dA, const double *dpi, double *ll_out, double *grad_out,
                            double *par_grad_out, double *rate_grad_out, double *moments_out);
int mcp_model_gradient_contract(int K, int R, int n_branches, const double *blv, const double *U, const double *D,
                                const double *Uinv, double mu, const double *rates, const double *moments,
                                const double *root_w, int n_par, const double *dA, const double *dpi,
                                double *par_grad_out, double *grad_check_out, double *rate_grad_out);

/*
 * Same evaluation, result left on the device: d_out (DEVICE pointer, NN doubles) receives
 * [logL, grad[1..NN-1]] (grad part zero if !want_grad).  The work is enqueued on the context's
 * stream and NOT synchronised, so a site-sharded caller can all-reduce d_out across GPUs
 * (one ncclAllReduce of NN doubles) before reading it.
 */
int mcp_eval_device(mcp_ctx *ctx, const mcp_alignment *aln, int NN, const int32_t *postorder_num,
                    const int32_t *parent_num, const double *blv, const double *U, const double *D,
                    const double *Uinv, double mu, const double *rates, int R, const double *pi,
                    int want_grad, double *d_out);

/*
 * One evaluation of an alignment that lives in HOST memory and is not kept on the device: `codes` is the
 * (n_leaves, S) row-major array mcp_alignment_from_codes takes.  Each device's site range is cut into a few
 * blocks (a small first one, then doubling); block b+1 crosses PCIe on the copy stream while block b is
 * evaluated, the block results are added on the device in block order, reduced over the devices of a
 * multi-device context as in mcp_eval, and read back once.  Device buffers and plans persist between
 * calls with the same (K, S, leaf_nums, NN, R, want_grad), so repeated calls only pay transfer + kernels.
 * Use page-locked memory for `codes` (mcp_host_register pins an existing array in place) -- pageable
 * memory works but is staged by the driver and does not overlap.
 * mcp_stream_blocks reports the blocks of one device after a call: up to `cap` pairs [lo, hi) into lo_hi,
 * returns their number.
 */
int mcp_eval_streamed(mcp_ctx *ctx, const uint8_t *codes, int K, int64_t S, const int32_t *leaf_nums,
                      int n_leaves, int NN, const int32_t *postorder_num, const int32_t *parent_num,
                      const double *blv, const double *U, const double *D, const double *Uinv, double mu,
                      const double *rates, int R, const double *pi, int want_grad, double *ll_out,
                      double *grad_out);
int mcp_stream_blocks(const mcp_ctx *ctx, int member, int64_t *lo_hi, int cap);
/* Device timeline of the last mcp_eval_streamed call on one device, 5 doubles per block (milliseconds since
 * the first transfer began): transfer begin, transfer end, evaluation enqueued, walk kernel begin, walk
 * kernel end.  Returns the number of blocks written (at most cap_blocks). */
int mcp_stream_timeline(const mcp_ctx *ctx, int member, double *ms, int cap_blocks);
int mcp_host_register(void *p, size_t bytes);
int mcp_host_unregister(void *p);

/*
 * T independent evaluations in one launch (MultiplePhyloDist; also proposal/chain batches).
 * Every per-tree argument of mcp_eval becomes an array of T entries.  K and R are shared.
 *   ll_out    T doubles
 *   grad_out  T pointers to NN[t]-1 doubles each (array or entries may be NULL if !want_grad)
 */
int mcp_eval_batch(mcp_ctx *ctx, int T, const mcp_alignment *const *alns, const int32_t *NN,
                   const int32_t *const *postorder_num, const int32_t *const *parent_num,
                   const double *const *blv, const double *const *U, const double *const *D,
                   const double *const *Uinv, const double *mu, const double *const *rates, int R,
                   const double *const *pi, int want_grad, double *ll_out, double *const *grad_out);

/*
 * Measurement hooks (bench.py).  Timings are CUDA-event times on the context's stream for the
 * most recent evaluation: the fused pruning+gradient kernel alone, and the whole device-side
 * sequence (uploads, transition tables, walk, reduction, download).  Valid after a synchronous
 * call (mcp_eval / mcp_eval_batch).
 */
typedef struct mcp_stats {
    double walk_ms;            /* kernel felsenstein_walk */
    double device_ms;          /* first H2D .. last D2H on the stream */
    int64_t h2d_bytes;         /* bytes copied host->device by the last evaluation */
    int64_t d2h_bytes;         /* bytes copied device->host by the last evaluation */
    int32_t kernel_launches;   /* kernels launched by the last evaluation */
    int32_t grid, block;       /* walk kernel launch shape */
    int32_t tiles;             /* column tiles processed */
    int32_t schedule_rebuilt;  /* 1 if the topology differed from the cached one */
    int64_t scratch_bytes;     /* device scratch currently held for partials */
    int32_t columns_per_thread;/* walk kernel: alignment columns per thread (1 or 2) */
    int32_t operand_ring;      /* walk kernel: depth of the gradient pass's operand ring (0 = not used by the last evaluation) */
} mcp_stats;
int mcp_get_stats(const mcp_ctx *ctx, mcp_stats *out);
int mcp_get_stats_member(const mcp_ctx *ctx, int member, mcp_stats *out);
/* Device-side stopwatch over any number of calls: mcp_timer_start records a CUDA event on the evaluation
 * stream of every device of the context, mcp_timer_stop records a second one, waits for it and returns
 * the elapsed milliseconds (the slowest device's).  The context's streams are its own, so a caller's
 * events (torch.cuda.Event, ...) would not see this work. */
int mcp_timer_start(mcp_ctx *ctx);
int mcp_timer_stop(mcp_ctx *ctx, double *ms_out);

/*
 * Columns (sites x rate categories) that ONE full wave of the persistent walk grid covers for a large
 * input with K states and a tree of n_nodes nodes: resident CTAs x columns per tile.  A caller that
 * cuts an alignment into site blocks (to overlap mcp_alignment_update_codes of one block with the
 * evaluation of another) keeps every launch free of a ragged last wave by making each block a
 * multiple of columns / R sites.
 */
int mcp_wave_columns(mcp_ctx *ctx, int K, int n_nodes, int want_grad, int64_t *columns);

/* Tuning knobs: block = threads per CTA (a multiple of 32 up to 256; a tile is block x columns-per-thread
 * alignment columns wide; 0 = automatic: 256, the measured optimum at every input size, narrowed only for
 * alignments of fewer than 256 sites), ctas_per_sm = persistent CTAs per SM (0 = occupancy maximum). */
int mcp_set_launch(mcp_ctx *ctx, int block, int ctas_per_sm);
/* The four CUDA timing events every evaluation records for mcp_get_stats (walk_ms, device_ms): 1 = recorded (default),
 * 0 = not recorded, the two times then read 0.  A latency-bound caller (one MCMC-sized evaluation per leapfrog, cfg2:
 * a 38 us kernel) saves the events' share of the call; MCPHYLO_B200_TIMING=0 sets the same default per process. */
int mcp_set_timing(mcp_ctx *ctx, int on);
/* Alignment columns walked by one thread: 1, 2, or 0 = automatic (2 once the GPU is full). */
int mcp_set_columns_per_thread(mcp_ctx *ctx, int cpt);
/* Where a CTA keeps its partial-likelihood scratch: -1 automatic (shared memory for small inputs
 * whose scratch fits, HBM otherwise), 0 always HBM, 1 shared memory whenever it fits. */
int mcp_set_scratch_mode(mcp_ctx *ctx, int mode);
/* Where a CTA of the walk kernel accumulates its branch-gradient sums: -1 automatic (shared memory,
 * without atomics and bit-reproducible from run to run, for trees of up to 4096 branches; the CTA's
 * row in global memory with RED.ADD.F64 beyond that), 0 always shared memory (fails if the tree does
 * not fit), 1 always global memory. */
int mcp_set_accumulator_mode(mcp_ctx *ctx, int mode);
/* Gradient pass of the depth-first walk (K = 2, 4; one substitution model per call; trees of up to 4096 nodes):
 * -1 / 1 the stored child partials are fetched ahead of their use through a per-warp operand ring in shared memory
 * (cp.async.bulk + mbarrier), 0 every thread loads its own partials when it needs them.  Results are identical. */
int mcp_set_ring_mode(mcp_ctx *ctx, int mode);
/* Environment switches read once per process (measurement / fall-back aids, never needed for correct results):
 *   MCPHYLO_B200_INLINE_PARAMS=0   MCMC-sized evaluations stage their parameter block through device memory (a second
 *                                  launch) instead of carrying it in the arguments of the fused small-tree kernel
 *   MCPHYLO_B200_STREAM_BLOCKS=1   mcp_eval_streamed uses the block-per-launch pipeline instead of one fused launch
 *   MCPHYLO_B200_SERIAL_GROUP=1    a multi-device context enqueues its devices from the calling thread
 *   MCPHYLO_B200_NCCL=<path>       NCCL library to dlopen instead of libnccl.so.2
 *   MCPHYLO_B200_TIMING=0          contexts start with mcp_set_timing(ctx, 0)
 *   MCPHYLO_B200_MG_REPLICAS=<n>   replicas of the moment matrices of mcp_eval_model_gradient (default: as many as 64 MB hold, <= 32) */
/* State counts 6 < K <= 32 (e.g. 20-state protein alphabets): -1 / 1 the tile-cooperative kernel that runs the
 * K x K by K x columns products of every node on the FP64 tensor path (mma.sync.m8n8k4.f64; 8 warps x 16 columns
 * per CTA, bit-reproducible gradient), 0 the runtime-K fallback kernel (one thread per column, CUDA cores). */
int mcp_set_large_alphabet_mode(mcp_ctx *ctx, int mode);
/* Gradient pass, K <= 6: an internal node whose two children are leaves (a "cherry") is recomputed from the
 * two leaf codes instead of being stored by the post pass and re-read (bit-identical; a third of the stored
 * partials of a random binary tree).  -1 automatic (on), 0 off (every partial stored), 1 on. */
int mcp_set_cherry_mode(mcp_ctx *ctx, int mode);
/* How the persistent CTAs of the walk share the column tiles of a RESIDENT single-tree evaluation: 0 a static
 * contiguous range per CTA, rate-major; 1 an atomic ticket per tile, site-major (what mcp_eval_streamed always
 * uses); -1 automatic.  Results are identical up to the order of the per-CTA sums. */
int mcp_set_tile_order(mcp_ctx *ctx, int mode);
/* Level-parallel small-tree kernel (all warps of a CTA share one 32-column tile, one barrier per
 * tree level): -1 automatic (inputs of at most a few tiles per SM whose tree fits in shared memory),
 * 0 never, 1 whenever the tree fits. */
int mcp_set_level_mode(mcp_ctx *ctx, int mode);

/*
 * Host-only: emits the device schedule (the flat "walk program") for a topology, so the
 * scheduler can be checked without a GPU.  leaf_row[num-1] = alignment row of that leaf, or -1
 * for internal nodes.  Ops are 8 int32 each (layouts in csrc/schedule.hpp).  Pass cap_* =
 * capacity of the arrays in ops; returns MCP_ERR_ARG if too small.
 *   info[0]=n_post info[1]=n_pre info[2]=n_slots info[3]=n_stack info[4]=n_dnodes
 *   info[5]=post levels info[6]=pre levels (info must hold 8 ints)
 * want_grad: bit 0 = gradient program wanted, bit 1 = emit the level-ordered variant used by the
 * small-tree kernel instead of the depth-first one, bit 2 = cherries recomputed in the gradient pass
 * (OPK_CHERRY children, csrc/schedule.hpp); info[7] = number of such children.
 */
int mcp_schedule_dump(int NN, const int32_t *postorder_num, const int32_t *parent_num,
                      const int32_t *leaf_row, int want_grad, int32_t *post_ops, int cap_post,
                      int32_t *pre_ops, int cap_pre, int32_t *info);
/* Host only: the fetch list of the gradient pass's operand ring for the same tree -- the post slots of the stored child
 * partials in the order the gradient program reads them (per family: child a if stored, then child b if stored).
 * slots: cap uint16; *n_out receives the number of entries. */
int mcp_schedule_fetch_list(int NN, const int32_t *postorder_num, const int32_t *parent_num, const int32_t *leaf_row,
                            int cherries, uint16_t *slots, int cap, int32_t *n_out);

/*
 * Host-only: the reordering every evaluation applies to the caller's eigen-decomposition before it
 * is uploaded -- the eigenvalue of smallest magnitude moved to the last position (columns of U,
 * entries of D, rows of Uinv permuted alike, so U diag(f(D)) Uinv is unchanged).  Returns 1 in
 * *null_last when that eigenvalue is null (<= 8 eps max|D|), i.e. when the kernels skip it.
 * U, Uinv, U_out, Uinv_out are K x K column-major; D, D_out hold K doubles.
 */
int mcp_model_reorder(int K, const double *U, const double *D, const double *Uinv,
                      double *U_out, double *D_out, double *Uinv_out, int *null_last);

#ifdef __cplusplus
}
#endif
#endif /* MCPHYLO_B200_H */
