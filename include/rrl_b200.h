/*
 * rrl_b200.h -- C ABI of the B200-native intersected-line robust registration loss.
 *
 * The reference (Dengzhi-USTC/A-robust-registration-loss) is pure Python/PyTorch and has no FFI
 * of its own; each entry point below names the reference function it replaces (file:line under
 * /root/reference/code/).  INTEGRATION.md shows the ctypes binding a maintainer of the reference
 * would add to route code/loss.py through this library.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name starts with `h_` (host);
 *   - all tensors are dense, row-major, float32 unless stated; the caller owns every buffer;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is enqueued on it
 *     and nothing synchronises with the host (except the rrl_host_* and rrl_measure_* helpers);
 *   - every function returns an int status: RRL_OK or a negative RRL_ERR_*; nothing ever exits the process
 *     (the reference print()s and exit(0)s on bad rank or NaN, loss.py:69-71,89-91);
 *   - the library keeps no global state besides a kernel-launch counter.
 *
 * Geometry of one call: B independent pairs; cloud c of pair b is `tri_c[b]` = nf_c triplets of 9 floats
 * [p0 | p1 | p2] (a point and its two nearest neighbours, loss.py:481-485); `lines[b]` = nl lines of 6 floats
 * [unit direction u | point x0] (loss.py:411).
 */
#ifndef RRL_B200_H
#define RRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRL_VERSION 100            /* 0.1.0 */
#define RRL_HIT_CAP 5              /* hit slots kept per (line, cloud); a line with >4 hits is never selected */
#define RRL_NSTAT 8                /* int64 diagnostics per pair, see rrl_loss_forward */

enum {
    RRL_OK = 0,
    RRL_ERR_ARG = -1,              /* null pointer / non-positive size / window outside 1..4 */
    RRL_ERR_WORKSPACE = -2,        /* workspace smaller than rrl_workspace_bytes() */
    RRL_ERR_CUDA = -3,             /* a CUDA runtime call or kernel launch failed */
    RRL_ERR_STATE = -4             /* host-buffer context: every slot in flight (submit) / ticket not in flight (wait).
                                      A workspace that holds no forward of the stated geometry cannot be seen from the host
                                      without a synchronisation; the kernels that consume a forward (backward, export_hits,
                                      the rrl_shard_* stages) check its header ON THE DEVICE and return early: gradients
                                      stay zero, counts are 0, hit slots -1. */
};

/* per-pair status bits written by rrl_loss_forward into out_status[b] */
enum {
    RRL_STATUS_EMPTY = 1,          /* no (k,j) combo populated: loss = 0, gradients = 0
                                      (the reference returns the tuple (None, None, None), loss.py:232) */
    RRL_STATUS_NAN = 2,            /* a candidate distance was NaN (reference: "Exit the systerm", loss.py:89-91) */
    RRL_STATUS_NAN_RISK = 4,       /* coordinates so large that |AC|^2 cancellation may exceed 2e-4 */
    RRL_STATUS_COMM = 8            /* line shard: a peer never arrived at the exchange (rrl_shard_tail timed out) */
};

int rrl_version(void);
const char *rrl_error_string(int code);
/* number of kernels this library has launched since it was loaded (bench.py's gpu_launches) */
long long rrl_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * Loss: replaces cal_loss_intersection_batch_whole_median_pts_lines (loss.py:170-232) together with
 * cal_intersection_batch2_points_with_line (loss.py:68-112) and
 * cal_loss_intersection_batch_m_n_median_pts_lines (loss.py:115-167), per pair.
 * ------------------------------------------------------------------------------------------------- */

/* bytes of scratch rrl_loss_forward needs for this geometry (also holds what backward re-reads) */
size_t rrl_workspace_bytes(int B, int nf1, int nf2, int nl);

/*
 * Forward.  (k_lo, j_lo, k_hi, j_hi) are the reference's (s_m, s_n, e_m, e_n): half-open windows of hit
 * counts, 1 <= lo < hi <= 5; every caller in the reference passes (1, 1, 5, 5).
 *   out_loss   [B]          loss of each pair, defined as the reference called on that B=1 slice
 *   out_status [B]          RRL_STATUS_* bits
 *   out_median [B]          lower median of the pair's D entries (loss.py:223-224); may be NULL
 *   out_stats  [B*RRL_NSTAT] int64: {0: #selected lines, 1: #D entries, 2: #non-empty combos C,
 *                            3: #queued (line, node-group) candidates cloud 1, 4: same for cloud 2,
 *                            5: #tests within 1 ulp of the threshold that can decide a label, i.e. whose two sibling points pass
 *                               or sit in the band themselves (the north star's separately reported band),
 *                            6: #NaN distances among candidates, 7: reserved}; may be NULL
 */
int rrl_loss_forward(const float *tri1, const float *tri2, const float *lines,
                     int B, int nf1, int nf2, int nl,
                     int k_lo, int j_lo, int k_hi, int j_hi,
                     void *workspace, size_t workspace_bytes,
                     float *out_loss, int *out_status, float *out_median, long long *out_stats,
                     void *stream);

/*
 * Forward with flags.  RRL_REUSE_ORDER: `workspace` holds a completed forward of the SAME geometry (B, nf1, nf2, nl);
 * the spatial order of both clouds computed then is kept and the sort stage is skipped.  Any order gives the same
 * results (it only decides which triplets share a bounding sphere), so this is safe for clouds that changed -- and
 * pays when they changed little: the steps of a registration loop (the target is fixed, the source moves rigidly:
 * test_demo_optimized_Lie_Algebra.py:46-66) or the iterations of RPM-Net / FMR on one batch.
 * Contract of both flags: between its forwards the caller keeps `workspace` to itself (nobody else writes it).  Besides the
 * order, small state survives there from one forward to the next (for batches of up to 4 pairs of small clouds, the
 * counter by which the prep kernel's CTAs hand each other the pair's line extent).
 */
#define RRL_REUSE_ORDER 1
/* RRL_REUSE_TARGET (implies RRL_REUSE_ORDER; clouds above 4096 triplets, ignored below): additionally, cloud 2 -- the
 * registration TARGET -- is bit-identical to the previous forward's in this workspace (the caller's contract: the target of a
 * registration loop never changes, test_demo_optimized_Lie_Algebra.py:46-66).  Its thresholds, triplet records and
 * bounding-sphere levels are then kept instead of rebuilt -- unless the new lines reach farther out than the slack those
 * records were built for, which is checked on the device (they are rebuilt in that case, for 1.1x the new extent). */
#define RRL_REUSE_TARGET 2
int rrl_loss_forward_ex(const float *tri1, const float *tri2, const float *lines,
                        int B, int nf1, int nf2, int nl,
                        int k_lo, int j_lo, int k_hi, int j_hi,
                        void *workspace, size_t workspace_bytes,
                        float *out_loss, int *out_status, float *out_median, long long *out_stats,
                        int flags, void *stream);

/*
 * Backward of the forward held in `workspace` (autograd of loss.py:170-232; closed form SURVEY 9.1).
 * grad_out [B] is d(total)/d(out_loss[b]).  grad_tri1 (B,nf1,9) / grad_tri2 (B,nf2,9) are overwritten
 * (zero where no selected line touches a triplet); either may be NULL when not needed.
 */
int rrl_loss_backward(const void *workspace, size_t workspace_bytes, const float *grad_out,
                      int B, int nf1, int nf2, int nl,
                      float *grad_tri1, float *grad_tri2, void *stream);

/*
 * Per-line intersection sets of the forward held in `workspace` (the reference's label tensor,
 * loss.py:107-112, in sparse form).  cloud = 1 or 2.  out_counts (B,nl) int32 = number of hit triplets
 * (uncapped); out_hits (B,nl,RRL_HIT_CAP) int32 = ascending triplet indices, -1 padded; when a line has more
 * than RRL_HIT_CAP hits the slots hold an arbitrary RRL_HIT_CAP-subset (such lines are never selected).
 */
int rrl_loss_export_hits(const void *workspace, size_t workspace_bytes, int B, int nf1, int nf2, int nl,
                         int cloud, int *out_counts, int *out_hits, void *stream);

/*
 * Line-sharded evaluation of ONE pair across ranks: each rank calls the three stages on its own
 * slice of the lines (B = 1) and combines the small per-pair arrays between them with its collective library.
 *   stage 1  dense + intersection points; then
 *              rrl_shard_counts() -> 18 int64 {n_kj[16], #records, #D entries}: all-reduce(sum);
 *              rrl_shard_pack_entries() -> the rank's D entries, flat: all-gather;
 *   stage 2  given the global median and the all-reduced counts: Welsch sums -> 32 int64 fixed-point partial
 *            sums: all-reduce(sum);
 *   stage 3  given the all-reduced sums: loss (identical on every rank); backward then uses the global counts,
 *            so every rank's point gradient is its lines' share: all-reduce(sum) of the gradient.
 */
int rrl_shard_stage1(const float *tri1, const float *tri2, const float *lines, int nf1, int nf2, int nl,
                     int k_lo, int j_lo, int k_hi, int j_hi, void *workspace, size_t workspace_bytes, void *stream);
/* stage 1 with the flags of rrl_loss_forward_ex (RRL_REUSE_ORDER) */
int rrl_shard_stage1_ex(const float *tri1, const float *tri2, const float *lines, int nf1, int nf2, int nl,
                        int k_lo, int j_lo, int k_hi, int j_hi, void *workspace, size_t workspace_bytes, int flags,
                        void *stream);
int rrl_shard_counts(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl, long long *counts18,
                     void *stream);
int rrl_shard_pack_entries(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl,
                           float *out_entries, long long capacity, void *stream);
/* lower median of a flat device array of n non-negative floats (n read from the host argument) */
int rrl_select_lower_median(const float *values, long long n, float *out_median, void *stream);
/* Distributed lower median without a host round trip or a variable-length exchange: for round = 0, 1 every rank
 * histograms its D entries (round 0: high 16 bits of the float bit pattern; round 1: low 16 bits inside the bin chosen
 * in round 0) into out_hist65536 (int32), the caller sums the histograms over the ranks (all-reduce), and
 * rrl_shard_select_pick narrows state2 = {prefix, rank} from the summed histogram and the summed counts; after round 1
 * out_median holds the global lower median on every rank (0 when no rank has an entry). */
int rrl_shard_select_hist(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl, int round,
                          const long long *state2, int *out_hist65536, void *stream);
int rrl_shard_select_pick(int round, const int *global_hist65536, const long long *global_counts18, long long *state2,
                          float *out_median, void *stream);
int rrl_shard_stage2(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl,
                     const long long *global_counts18, const float *global_median, long long *sums32, void *stream);
int rrl_shard_stage3(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl,
                     const long long *global_sums32, float *out_loss, int *out_status, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Peer-memory exchange of the line shard ("one exchange step", here without NCCL).  One communicator per
 * rank (= process = GPU) of a box; the ranks publish the CUDA IPC handle of their buffer, the caller carries the
 * 64-byte handles between the processes, every rank maps its peers (NVLink peer access).  The exchange is then done
 * inside kernels: payloads are pushed into the peers' buffers with plain stores, arrival is signalled by a sequence
 * number, nothing involves the host, everything is CUDA-graph capturable.  Every rank of a communicator must issue the
 * same sequence of exchanging calls (rrl_shard_tail, rrl_comm_allreduce_f64).
 *   slot_bytes   capacity of one payload; rrl_shard_tail needs 160 + 64 * nl_local bytes
 * ------------------------------------------------------------------------------------------------- */
typedef struct rrl_comm rrl_comm;
int rrl_comm_create(int rank, int world, size_t slot_bytes, rrl_comm **out);     /* on the current device; world <= 16 */
void rrl_comm_destroy(rrl_comm *comm);
size_t rrl_comm_slot_bytes(const rrl_comm *comm);
void *rrl_comm_local_base(const rrl_comm *comm);
int rrl_comm_ipc_handle(const rrl_comm *comm, void *out_handle64);
int rrl_comm_connect_ipc(rrl_comm *comm, const void *handles64 /* world x 64 bytes, rank-major */);
int rrl_comm_connect_ptrs(rrl_comm *comm, void *const *bases /* world local bases of same-process peers */);
int rrl_comm_error(const rrl_comm *comm);                                         /* 1 after a timed-out exchange; synchronises */
/* in-place all-reduce(sum) of n <= 512 doubles, summed in rank order on every rank (bit-identical results) */
int rrl_comm_allreduce_f64(rrl_comm *comm, double *buf, int n, void *stream);
/*
 * Fused tail of the line-sharded forward: after rrl_shard_stage1(_ex) on the same workspace and stream, ONE launch per
 * rank exchanges the counts and the D entries, selects the global lower median (loss.py:223-224 over the lines of all
 * ranks), runs the Welsch stage on the rank's records, exchanges the 32 fixed-point partial sums and writes the loss
 * (loss.py:225-230) -- identical on every rank.  Replaces rrl_shard_counts / rrl_shard_select_* / rrl_shard_stage2/3 and
 * the four collectives between them.  rrl_loss_backward(B = 1) then yields the gradient share of this rank's lines.
 * out_status additionally carries bit 8 when a peer never arrived (the loss is NaN then).
 */
int rrl_shard_tail(void *workspace, size_t workspace_bytes, int nf1, int nf2, int nl, rrl_comm *comm,
                   float *out_loss, int *out_status, float *out_median, long long *out_stats, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * se(3): replaces Reconstruction_point.forward / Transform (loss.py:455-463) and se3.exp3
 * (LieAlgebra/se3.py:83-106, so3.py:17-27, sinc.py:5-17,91-103,120-132).
 * twist (B,6) = [w | v];  R (B,3,3) row-major, T (B,3);  points (B,n,3);  out = points @ R + T (row vectors).
 * ------------------------------------------------------------------------------------------------- */
int rrl_se3_exp(const float *twist, int B, float *R, float *T, void *stream);
/* FMR's se3.Exp (exps_deep_learning/fmr/se_math/se3.py:60-84,133-165): g (B,4,4) = [R p; 0 0 0 1] with the R and
 * p = V v of exp3.  The backward is the reference's ExpMap.backward, grad_twist[k] = sum_ij grad_g[i][j] (gen_k g)[i][j]
 * (the left-trivialised tangent the IC solver of FMR propagates -- not the analytic derivative of exp). */
int rrl_se3_exp4(const float *twist, int B, float *g, void *stream);
int rrl_se3_expmap_backward(const float *twist, const float *grad_g, int B, float *grad_twist, void *stream);
int rrl_se3_apply(const float *twist, const float *points, int B, int n, float *out, void *stream);
/* grad_out (B,n,3) = d loss / d out  ->  grad_twist (B,6).  scratch: B*12 doubles. */
int rrl_se3_apply_backward(const float *twist, const float *points, const float *grad_out, int B, int n,
                           float *grad_twist, double *scratch, void *stream);
/* Backward straight to pose space for a cloud 1 that is exp(twist) applied to `raw_tri1` (B,nf1,9) (the demo's
 * Reconstruction_point -> loss -> backward chain, test_demo_optimized_Lie_Algebra.py:57-66): contracts the sparse point
 * gradient of the forward held in `workspace` with the raw points, acc12 (B,12) double = {sum p (x) g (9), sum g (3)}
 * (zeroed here), without materialising a dense (B,nf1,9) gradient; with grad_twist != NULL the closed-form derivative of
 * exp3 is applied right away, else the caller reduces acc12 over its line shards first and calls rrl_se3_chain. */
int rrl_loss_backward_twist(const void *workspace, size_t workspace_bytes, const float *grad_out,
                            int B, int nf1, int nf2, int nl, const float *twist, const float *raw_tri1,
                            double *acc12, float *grad_twist, void *stream);
int rrl_se3_chain(const float *twist, const double *acc12, int B, float *grad_twist, void *stream);
/* rigid transform with explicit (R,t) for the DCP / RPM-Net / FMR hooks (utils.py:32-37,
 * rpm/common/math_torch/se3.py:55-82, fmr/se_math/se3.py:110-124): out = points @ R^T + t (column convention
 * R p + t).  Backward: grad_R (B,3,3), grad_t (B,3), and optionally grad_points. */
int rrl_rigid_apply(const float *R, const float *t, const float *points, int B, int n, float *out, void *stream);
int rrl_rigid_apply_backward(const float *R, const float *points, const float *grad_out, int B, int n,
                             float *grad_R, float *grad_t, float *grad_points, double *scratch, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Line sampler: replaces Random_uniform_distribution_lines_batch_efficient_resample (loss.py:415-432) with
 * generate_bbox / generate_mesh_by_bbox / cal_intersection_batch2_rand_lines / generate_lines
 * (loss.py:265-381) and Random_uniform_distribution_lines_batch_efficient (loss.py:384-412).
 *   radius (B), centers (B,3), verts1 (B,n1,3), verts2 (B,n2,3) -> out_lines (B,N,6), out_filled (B) int32.
 * Draws come from a counter-based Philox4x32-10 stream keyed by (seed, offset) -- reproducible and independent
 * of launch geometry -- unless `uniforms` (B,rounds,4,N) is supplied, in which case exactly those draws are used
 * (order alpha1, u1, alpha2, u2 per round, like the reference's four torch.rand calls).  Rows >= filled stay
 * all-zero, exactly like the reference.  rounds = 10 in the reference.
 * ------------------------------------------------------------------------------------------------- */
size_t rrl_sampler_workspace_bytes(int B, int N, int rounds);
int rrl_sample_lines(const float *radius, const float *centers, const float *verts1, const float *verts2,
                     int B, int n1, int n2, int N, int rounds, unsigned long long seed, unsigned long long offset,
                     const float *uniforms, float *out_lines, int *out_filled,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Candidate-sharded form for a line-sharded evaluation.  Candidates are a pure function of (seed,
 * offset, pair, round, index), so rank r of `world` evaluates the chunks of 256 candidates with chunk = r (mod world):
 *   1. rrl_sample_lines_shard_flags   -> out_chunk_counts (B, rrl_sampler_num_chunks(N, rounds)) int32: accepted candidates of
 *                                        the rank's own chunks (0 elsewhere); all `rounds` rounds are evaluated;
 *   2. the caller sums the counts over the ranks (one all-reduce of nchunks ints) and keeps its own copy;
 *   3. rrl_sample_lines_shard_scatter -> out_lines_local (B, N, 6): the accepted candidates of the rank's chunks whose row in the
 *                                        reference's ordered compaction (loss.py:365-381) is below N, followed by the rank's share
 *                                        (row = rank mod world) of the rows that stay unfilled (all-zero lines), zeros beyond;
 *                                        out_counts3 (B,3) int32 = {lines placed, zero rows appended, rows filled globally}.
 * The union over the ranks of the first counts3[0] + counts3[1] local rows is exactly the set of N rows rrl_sample_lines
 * returns (the loss does not depend on their order), so no line ever crosses NVLink.  Same workspace for both calls. */
int rrl_sampler_num_chunks(int N, int rounds);
int rrl_sample_lines_shard_flags(const float *radius, const float *centers, const float *verts1, const float *verts2,
                                 int B, int n1, int n2, int N, int rounds, unsigned long long seed, unsigned long long offset,
                                 const float *uniforms, int shard_rank, int shard_world, int *out_chunk_counts,
                                 void *workspace, size_t workspace_bytes, void *stream);
int rrl_sample_lines_shard_scatter(const float *radius, const float *centers, int B, int N, int rounds,
                                   unsigned long long seed, unsigned long long offset, const float *uniforms,
                                   int shard_rank, int shard_world, int *chunk_counts_global, int *chunk_counts_local,
                                   float *out_lines_local, int *out_counts3, void *workspace, size_t workspace_bytes,
                                   void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Monitoring metric: replaces chamfer_dist (loss.py:236-252).  x (B,M,3), y (B,N,3) -> out (1) = mean over the
 * concatenation of both directed min-squared-distances of all pairs.  scratch: B*(M+N) floats.
 * ------------------------------------------------------------------------------------------------- */
int rrl_chamfer(const float *x, const float *y, int B, int M, int N, float *out, float *scratch, void *stream);
/* The differentiable form (the reference's chamfer_dist is plain autograd; Train_DCP.py:248,297 returns it inside the loss
 * tuple): forward also records the argmin of every directed minimum (first index on ties, like torch.min) in
 * argmin (B*(M+N)) int32 -- x's M minima, then y's N --; backward scatters d out / d x (B,M,3) and d out / d y (B,N,3)
 * (either may be NULL) for the upstream gradient grad_out (1). */
int rrl_chamfer_forward(const float *x, const float *y, int B, int M, int N, float *out, float *scratch, int *argmin,
                        void *stream);
int rrl_chamfer_backward(const float *x, const float *y, const int *argmin, const float *grad_out, int B, int M, int N,
                         float *grad_x, float *grad_y, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Pre-processing: replaces Sample_neighs (loss.py:473-485) = utils.farthest_point_sample (utils.py:275-296) followed by
 * a 3-nearest-neighbour query (sklearn KDTree, loss.py:479-480).  `xyz` is one cloud (N,3), float32 (is_double = 0, what
 * every caller of the reference passes) or float64 (is_double = 1).
 *   rrl_fps  out_idx[npoint] int32: the reference recurrence from centroid `start` (the reference draws it with
 *            torch.randint on the CPU generator; the caller supplies it), first index on equal distances like
 *            torch.max; the running minimum is float32 as in the reference (torch.ones(B,N)*1e10).
 *   rrl_knn  out_idx (M,k) int32: the k <= 8 nearest points (squared distance in double, ties: smaller index) of
 *            every query point xyz[query_idx[q]], nearest first -- the query point itself comes first.
 * ------------------------------------------------------------------------------------------------- */
size_t rrl_fps_workspace_bytes(int N);
int rrl_fps(const void *xyz, int is_double, int N, int npoint, int start, int *out_idx, void *workspace,
            size_t workspace_bytes, void *stream);
int rrl_knn(const void *xyz, int is_double, int N, const int *query_idx, int M, int k, int *out_idx, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Host-buffer convenience path (what an FFI caller without device memory uses; bench.py's e2e leg).
 * The context owns device buffers, pinned staging and a stream for one geometry.
 * ------------------------------------------------------------------------------------------------- */
typedef struct rrl_host_ctx rrl_host_ctx;
int rrl_host_create(int B, int nf1, int nf2, int nl, int device, rrl_host_ctx **out_ctx);
void rrl_host_destroy(rrl_host_ctx *ctx);
/* pinned staging owned by the context (write inputs here to avoid an extra host copy); sizes as the device tensors */
float *rrl_host_pinned_tri1(rrl_host_ctx *ctx);
float *rrl_host_pinned_tri2(rrl_host_ctx *ctx);
float *rrl_host_pinned_lines(rrl_host_ctx *ctx);
/* number of sub-batches (streams) the context pipelines a call over: H2D of one runs under the kernels of the previous */
int rrl_host_subbatches(rrl_host_ctx *ctx);
/* H2D of the three inputs, forward, backward w.r.t. cloud 1 with d(total)/d(loss[b]) = 1, D2H of loss [B], status [B]
 * and (if h_grad_tri1 != NULL) the (B,nf1,9) gradient; returns after the stream has drained.  h_* inputs may be
 * the context's own pinned buffers, any other pinned memory (copied asynchronously) or pageable memory (staged by
 * the CUDA driver). */
int rrl_host_loss_fwd_bwd(rrl_host_ctx *ctx, const float *h_tri1, const float *h_tri2, const float *h_lines,
                          int k_lo, int j_lo, int k_hi, int j_hi,
                          float *h_loss, int *h_status, float *h_grad_tri1);

/* Double-buffered form of the same evaluation (a training loop's input pipeline): rrl_host_submit() queues copies and
 * kernels on the next of rrl_host_slots() buffer sets and returns a ticket without waiting, so the H2D copies of
 * evaluation i+1 run under the kernels of evaluation i; rrl_host_wait() drains that ticket and copies out loss [B],
 * status [B] (may be NULL) and, if requested at submit time, the (B,nf1,9) gradient.  The h_* inputs must stay valid and
 * unchanged until the wait returns.  RRL_ERR_STATE: every slot is in flight (submit) / the ticket is not (wait).
 * rrl_host_loss_fwd_bwd == submit + wait. */
int rrl_host_slots(rrl_host_ctx *ctx);
int rrl_host_submit(rrl_host_ctx *ctx, const float *h_tri1, const float *h_tri2, const float *h_lines,
                    int k_lo, int j_lo, int k_hi, int j_hi, int want_grad_tri1, int *out_ticket);
int rrl_host_wait(rrl_host_ctx *ctx, int ticket, float *h_loss, int *h_status, float *h_grad_tri1);

/* ---------------------------------------------------------------------------------------------------
 * Measurement helpers used by bench.py (not part of the reference's interface).
 * ------------------------------------------------------------------------------------------------- */
/* FP32 FMA peak of the current device in TFLOP/s, measured with a register-resident FFMA loop (mode 0) or the
 * packed FFMA2 form (mode 1); synchronises. */
int rrl_measure_fp32_peak(int mode, double *out_tflops, double *out_ms);
/* time (ms, CUDA events on `stream`) of the dense intersection kernel alone over the forward last staged in
 * `workspace` (re-runs prep + dense `iters` times; synchronises) */
int rrl_measure_dense(const float *tri1, const float *tri2, const float *lines, int B, int nf1, int nf2, int nl,
                      void *workspace, size_t workspace_bytes, int iters, float *out_ms_dense, float *out_ms_prep,
                      void *stream);

/* per-stage device times (ms) of forward + backward to points1, averaged over `iters` hot repetitions:
 * out_ms[10] = {memset+prep, sort, node, dense, select, build, median, welsch+finalize, backward, total}; synchronises */
int rrl_measure_stages(const float *tri1, const float *tri2, const float *lines, int B, int nf1, int nf2, int nl,
                       void *workspace, size_t workspace_bytes, int iters, float *out_ms, void *stream);
/* selects the dense-stage variant: 1 = Morton-sorted bounding-sphere nodes (default), 0 = nodes in input order.
 * Results are identical; measurement / A-B testing only. */
int rrl_debug_set_dense_variant(int variant);
/* launch-geometry knobs for A/B measurements: id 1 = triplets per node (0 auto, 8, 16), 2 = target waves of CTAs,
 * 3 = minimum nodes per CTA chunk, 5 = brute-force cross-check kernel, 6 = lines per thread, 7 = 1: no super-node level,
 * 8 = enclosing-ball steps of the node centres (small clouds), 9 = target waves in super-node mode (0 = by the number of line
 * tiles), 10 = k-d refinement of the Hilbert order (0 off, 1 small clouds, 2 everywhere), 11 = entries in flight per thread of
 * the exact kernel, 12 = lowest key bit the large-cloud radix sort looks at (0 auto, -1 every bit), 13 = CTAs per SM of the
 * exact kernel's grid, 14 = enclosing-ball steps on the large-cloud path, 15 = 1: no line-extent hand-off in the small-cloud
 * prep kernel.  Results never depend on them. */
int rrl_debug_set_param(int id, int value);

#ifdef __cplusplus
}
#endif
#endif /* RRL_B200_H */
