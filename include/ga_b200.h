/* ga_b200.h -- C ABI of the B200-native Chamfer / kNN hot path of geometric_adv.
 *
 * One shared library, libga_b200.so (geometric_adv_b200/csrc/), hand-written
 * CUDA for sm_100a.  Plain pointers and sizes only: no torch / TensorFlow types.
 * Every entry point names the reference interface it replaces (path:line below
 * the reference tree); INTEGRATION.md shows the binding a maintainer adds on
 * the reference side.
 *
 * Conventions (reference conventions kept unless noted):
 *   - fp32 row-major contiguous clouds (B,N,3), int32 indices.
 *   - The caller allocates every output (the reference ops use
 *     allocate_output, tf_nndistance.cpp:191-196); kernels never allocate.
 *     Inputs are read only.  Outputs are fully overwritten (no pre-zeroing
 *     needed, unlike chamfer3D.cu:177-178).
 *   - `*_fwd`, `*_bwd`, ga_knn, ... take DEVICE pointers and a cudaStream_t and
 *     only enqueue work (CUDA-graph capturable).  The `*_host` twins take HOST
 *     pointers, stage through pinned buffers and return when results are in the
 *     caller's memory: they are the drop-in for the reference's CPU-tensor
 *     kernels (REGISTER_KERNEL_BUILDER(...DEVICE_CPU), tf_nndistance.cpp:83,166).
 *   - Return value: GA_OK (0); a negative GA_ERR_* for argument errors; or a
 *     positive cudaError_t.  ga_last_error() gives the message for the calling
 *     thread.  (The reference does no CUDA error checking, chamfer3D.cu:145-151
 *     only printf()s.)
 *   - Re-entrant: entry points may be called concurrently from several threads and on several
 *     streams (tests/test_concurrency_gpu.py).  State: the per-thread error string, last-kernel
 *     name and staging arena / graph cache of the *_host entry points; a per-device completion
 *     ticket array (a slot overwritten by a concurrent call can only delay a ticket, never fake
 *     one); and the ga_set_tuning() knobs, which are process-wide plain ints meant for
 *     benchmarks and tests -- set them while no other thread is inside the library.
 *   - The tensor-core filter window is relative to (max|q_c| + max|t_c|)^2, i.e. to the distance of
 *     the clouds from the ORIGIN, not to their extent: clouds far from the origin relative to their
 *     size (offset / extent above ~10) lose the filter's selectivity and run at the speed of the
 *     exact scan (still bit-exact; timings in DESIGN.md "off-origin").  Centre such data first.
 */
#ifndef GA_B200_H_
#define GA_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* ga_stream_t; /* == cudaStream_t */

enum {
  GA_OK = 0,
  GA_ERR_INVALID_ARGUMENT = -1, /* shape / attr rejected, same conditions as the reference OP_REQUIRES */
  GA_ERR_UNSUPPORTED = -2,      /* valid for the reference but outside what this build handles */
  GA_ERR_NO_DEVICE = -3         /* no sm_100 device / driver */
};

/* Arithmetic of the squared distance (bit-level contract):
 *   GA_MODE_CPU_EXACT  ((x*x)+(y*y))+(z*z), unfused, x = target - query: bit-identical
 *                      to nnsearch, tf_nndistance.cpp:21-43 (g++ -O2).      [default]
 *   GA_MODE_GPU_REF    fma(z,z,fma(x,x,y*y)): bit-identical to the reference's
 *                      NmDistanceKernel (tf_nndistance_g.cu:5-127 == chamfer3D.cu:12-134)
 *                      as nvcc compiles it for sm_100a.
 * Both have the same speed: the scan runs on a cheaper filter and only the
 * surviving candidates are evaluated in the selected arithmetic. */
enum { GA_MODE_CPU_EXACT = 0, GA_MODE_GPU_REF = 1 };

int ga_version(void);
const char* ga_last_error(void);
/* Number of kernels this library has launched from the calling process so far
 * (bench.py's gpu_launches). */
long long ga_launch_count(void);

/* ---- argument checks with the reference's own messages ------------------- */
/* tf_nndistance.cpp:51-58 (NnDistance), :91-104 (NnDistanceGrad; pass the four
 * extra ranks/dims, or rank -1 to skip them).  Returns GA_OK or
 * GA_ERR_INVALID_ARGUMENT with the reference's message in ga_last_error(). */
int ga_check_nn_distance(int rank1, const long long* dims1, int rank2, const long long* dims2);
int ga_check_nn_distance_grad(int rank1, const long long* dims1, int rank2, const long long* dims2,
                              int rank_gd1, const long long* dims_gd1, int rank_idx1, const long long* dims_idx1,
                              int rank_gd2, const long long* dims_gd2, int rank_idx2, const long long* dims_idx2);
/* tf_grouping.cpp:112-118 (SelectionSort: k>0, rank 3), :149-155 (GroupPoint). */
int ga_check_selection_sort(int k, int rank, const long long* dims);
int ga_check_group_point(int rank_points, const long long* dims_points, int rank_idx, const long long* dims_idx);

/* ---- Chamfer: nn_distance and its gradient -------------------------------- */
/* Replaces NmDistanceKernelLauncher (tf_nndistance.cpp:168, tf_nndistance_g.cu:128-131)
 * and chamfer_cuda_forward (chamfer_cuda.cpp:9, chamfer3D.cu:136-154), and, through the
 * _host twin, NnDistanceOp's CPU path (tf_nndistance.cpp:45-83).
 * dist1/idx1 (b,n): for each point of xyz1 the squared distance to / index of its
 * nearest point of xyz2 (lowest index on ties); dist2/idx2 (b,m): the other way. */
int ga_nn_distance_fwd(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                       float* dist2, int* idx2, int mode, ga_stream_t stream);
int ga_nn_distance_fwd_host(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                            float* dist2, int* idx2, int mode);

/* Same result, fewer evaluations: with a caller-provided scratch buffer the clouds are first put
 * in Morton order (one small kernel), and each warp of 128 spatially adjacent queries then visits
 * the 32-point target tiles in ascending box-distance order and stops as soon as no remaining tile
 * can hold a candidate (exact; see nn_distance_sorted.cu).  `workspace` must hold
 * ga_nn_distance_workspace_bytes(b,n,m) bytes of device memory; it carries no state between
 * calls.  Without a workspace, or for clouds outside 256..2048 points, this is ga_nn_distance_fwd. */
size_t ga_nn_distance_workspace_bytes(int b, int n, int m);
int ga_nn_distance_fwd_ws(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                          float* dist2, int* idx2, int mode, void* workspace, size_t workspace_bytes,
                          ga_stream_t stream);

/* Replaces NmDistanceGradKernelLauncher (tf_nndistance.cpp:208, tf_nndistance_g.cu:152-157),
 * chamfer_cuda_backward (chamfer_cuda.cpp:12, chamfer3D.cu:176-195) and
 * NnDistanceGradOp's CPU loops (tf_nndistance.cpp:122-163).  Atomic-free: every
 * output element is accumulated by one thread in exactly the CPU loop order, so
 * the result is bit-identical to the CPU reference and run-to-run reproducible.
 * Stream semantics: launched as a programmatic dependent of the preceding kernel on
 * `stream`.  idx1 / idx2 must be the arrays the matching ga_nn_distance_fwd call wrote,
 * unmodified: when that call was the caller's last forward launch on this stream, the
 * kernel starts on its index rows per batch element as soon as they are final, while
 * the rest of the search is still running; everything else (grad_dist*) is read only
 * behind the grid dependency. */
int ga_nn_distance_bwd(int b, int n, int m, const float* xyz1, const float* xyz2, const float* grad_dist1,
                       const int* idx1, const float* grad_dist2, const int* idx2, float* grad_xyz1,
                       float* grad_xyz2, ga_stream_t stream);
/* Forward + gradient in one call for upstream gradients that are final when the call is made (the
 * Chamfer loss of src/adv_ae.py:105,120-121 has constant d loss / d dist).  Same outputs as
 * ga_nn_distance_fwd followed by ga_nn_distance_bwd; the gradient kernel overlaps the last wave of
 * the search.  grad_dist1 / grad_dist2 must not be written by work still pending on `stream`. */
int ga_nn_distance_fwd_bwd(int b, int n, int m, const float* xyz1, const float* xyz2, const float* grad_dist1,
                           const float* grad_dist2, float* dist1, int* idx1, float* dist2, int* idx2,
                           float* grad_xyz1, float* grad_xyz2, int mode, ga_stream_t stream);
int ga_nn_distance_bwd_host(int b, int n, int m, const float* xyz1, const float* xyz2, const float* grad_dist1,
                            const int* idx1, const float* grad_dist2, const int* idx2, float* grad_xyz1,
                            float* grad_xyz2);

/* Forward + backward in one host call (what one training / attack step does
 * with the op): HOST buffers in, HOST buffers out, one H2D and one D2H leg.  Any of dist1 / idx1 /
 * dist2 / idx2 may be NULL: that output is not copied back (the attack consumes only the gradients
 * and a per-cloud loss on the host). */
int ga_nn_distance_fwd_bwd_host(int b, int n, int m, const float* xyz1, const float* xyz2,
                                const float* grad_dist1, const float* grad_dist2, float* dist1, int* idx1,
                                float* dist2, int* idx2, float* grad_xyz1, float* grad_xyz2, int mode);

/* Per-cloud Chamfer scalar, out[i] = mean(dist1[i,:]) + mean(dist2[i,:]): the reduction
 * the scripts apply to the op's outputs (src/adv_ae.py:120-121,
 * attacker/prepare_indices_for_attack.py:113-114).  Fixed summation tree => reproducible. */
int ga_chamfer_per_cloud(int b, int n, int m, const float* dist1, const float* dist2, float* out,
                         ga_stream_t stream);
/* The loss terms the attack graph builds from one nn_distance call (src/adv_ae.py:120-121,131-133), in one
 * launch: cd[i] = mean(dist1[i,:]) + mean(dist2[i,:]) (same summation tree as ga_chamfer_per_cloud: same bits) and
 * max1[i] = max(dist1[i,:]) (NaN if any entry is; max1 may be NULL). */
int ga_chamfer_loss_terms(int b, int n, int m, const float* dist1, const float* dist2, float* cd, float* max1,
                          ga_stream_t stream);

/* All-pairs driver of attacker/prepare_indices_for_attack.py:104-139: `clouds` (s,n,3);
 * out (rows,s) with out[r, j] = CD(source = clouds[j], target = clouds[row0 + r]) as in
 * :139,151.  Row blocks are the multi-GPU shard unit. */
int ga_chamfer_all_pairs(int s, int n, const float* clouds, int row0, int rows, float* out, int mode,
                         ga_stream_t stream);
/* One direction only: out[r, j] = mean over the points p of clouds[row0 + r] of
 * min_q |p - q|^2, q in clouds[j].  CD = D + D^T, so ranks that each own a row block compute
 * half the work of ga_chamfer_all_pairs, all-gather the blocks and symmetrise
 * (geometric_adv_b200/sharding.py). */
int ga_chamfer_all_pairs_directed(int s, int n, const float* clouds, int row0, int rows, float* out, int mode,
                                  ga_stream_t stream);
/* out (rows,s) = directed[row0 + r, j] + directed[j, row0 + r] for the gathered (s,s) matrix of directed
 * terms: rows [row0, row0+rows) of the symmetric Chamfer matrix (chamfer_dist_mat, :139,151). */
int ga_symmetrize_rows(int s, int row0, int rows, const float* directed, float* out, ga_stream_t stream);
/* sort_dist_mat of attacker/prepare_indices_for_attack.py:167-180 for `rows` rows of the (.,s) Chamfer
 * matrix: per target class c (columns [slice_idx[c], slice_idx[c+1]), slice_idx a DEVICE array of
 * nclass+1 ints, max_class = the largest class) the class-local column indices in ascending order of
 * distance, int16, into the same columns of nn_idx (rows,s) -- the array src/adversary_utils.py:51-63
 * takes the attack targets from.  Stable (ties by ascending index, NaN last): np.argsort(kind="stable");
 * the reference's default argsort leaves the order of exact ties unspecified. */
int ga_sort_dist_mat(int s, int rows, const float* dist_rows, int nclass, const int* slice_idx_dev, int max_class,
                     short* nn_idx, ga_stream_t stream);

/* ---- grouping: knn_point / selection_sort / group_point ------------------- */
/* knn_point(k, xyz1, xyz2) of tf_grouping.py:48-75 in ONE kernel: xyz1 (b,n,3) is the
 * data set, xyz2 (b,m,3) the queries; val/idx (b,m,k): the k smallest squared distances
 * ((dx*dx+dy*dy)+dz*dz, dx = xyz1 - xyz2) in ascending order and their indices into
 * xyz1, with exactly the tie behaviour of the reference's selection sort
 * (tf_grouping_g.cu:83-123).  Deliberate ABI change: no dense (b,m,n) matrix. */
int ga_knn(int b, int n, int m, int k, const float* xyz1, const float* xyz2, float* val, int* idx,
           ga_stream_t stream);
int ga_knn_host(int b, int n, int m, int k, const float* xyz1, const float* xyz2, float* val, int* idx);

/* Legacy dense entry, replaces selectionSortLauncher (tf_grouping.cpp:108,
 * tf_grouping_g.cu:129-132): dist (b,m,n) -> outi, out (b,m,n); the first k columns of
 * every row are the selection-sort result, the remaining columns hold the rest of the
 * row exactly as the reference leaves it. */
int ga_selection_sort(int b, int n, int m, int k, const float* dist, int* outi, float* out,
                      ga_stream_t stream);

/* Replaces groupPointLauncher (tf_grouping.cpp:142, tf_grouping_g.cu:133-136):
 * out[b,j,s,:] = points[b, idx[b,j,s], :]. */
int ga_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out,
                   ga_stream_t stream);

/* defender/get_knn_dists_per_point.py:78-81 fused: knn_point(k+1, pc, pc), drop the
 * first neighbour, gather, subtract the centre, sqrt(sum(delta^2)) -> out (b,n,k). */
int ga_knn_dists(int b, int n, int k, const float* pc, float* out, ga_stream_t stream);
int ga_knn_dists_host(int b, int n, int k, const float* pc, float* out);

/* get_outlier_pc_inlier_pc (src/adversary_utils.py:149-178) for a whole batch: per cloud, the points
 * with score > thresh go (in index order) to outlier_pc / outlier_idx, those with score <= thresh to
 * inlier_pc; a part that is neither empty nor the whole cloud is padded with its last point, empty
 * parts and the tail of outlier_idx are zero; outlier_num (b) counts.  The surface defense feeds it the
 * mean of the first two kNN distances and 0.04 (defender/run_defense_surface.py:187-191). */
int ga_split_by_threshold(int b, int n, const float* pc, const float* score, float thresh, float* outlier_pc,
                          int* outlier_idx, int* outlier_num, float* inlier_pc, ga_stream_t stream);

/* ---- measurement helpers --------------------------------------------------- */
/* Dependent-free FFMA loop on every SM; returns achieved FP32 TFLOP/s (2 flop per
 * FFMA lane) through *tflops.  bench.py uses it as the FP32 roofline denominator,
 * since MEASURED_PEAKS.json holds HBM and BF16 only. */
int ga_probe_fp32_peak(int iters, float* tflops, float* ms, ga_stream_t stream);
/* Tuning hooks for benchmarks and tests (process-wide; defaults in parentheses):
 *    0 forward kernel: 0 auto, 1-15 fp32-filter tile shapes, 20 HMMA grid kernel, 21 persistent HMMA,
 *      22 tcgen05/TMEM, 23 balanced persistent HMMA, 24 warp-specialised persistent HMMA   1 kNN variant            2 host path (0 auto,
 *      1 copies, 2 zero-copy)      3 host chunks (0 auto)     4 pruned-path variant    5 cluster split S (-1 auto)
 *    6 split kernel queries/thread  7 HMMA grid config (0 = 5)  8 persistent grids' CTA count (0 = SMs)
 *    9 gradient CTAs per cloud (-1 auto, 0 one, 1 four, 2 two)  10 host graph replay (0 auto, 1 off, 2 at once)
 *   11 replay chunks (0 auto)      12 tcgen05 grid CTAs (0 = SMs)  13 first gradient kernel: stage partner (1)
 *   14 gradient kernel (0 auto, 1 = stable counting sort, 2 = compacted lists, 3 = buckets + one round trip)   15 dependent launch (1)
 *   16 all-pairs kernel (0 auto = tensor-core scan, 1 = fp32 filter)   17 replay: dist/idx mirrored to the host
 *      (0 auto, 1 never, 2 always)   18 completion tickets (0 never, 1 one-call entry only, 2 always + debug, 3 always)
 *   19 all-pairs source clouds per CTA (0 auto)   20 tcgen05 kernel, development build: bit 0 no refine, bit 1 no
 *      drain, bit 3 no helper, bits 8.. traced CTA (0 = product build)   21 automatic choice of the tcgen05 kernel
 *      between the HMMA kernel's wave steps (1)   22 variant 24: scan warps (0 = 8; 8 or 12 of 16)   23 variant 24: start
 *      offset of every scheduler's second scan warp, ns (0)   24 variant 24, development: bit 0 no refine, bit 1 no
 *      query loads in the scan warps (wrong results; timing only)
 *   25 frame of the HMMA grid kernel's filter (2): 0 = origin, always; 1 = centred on the target cloud where that cloud
 *      keeps away from the origin (nn_fwd_mma_kernel<frame>); 2 = the plain kernel until one of its launches on the
 *      device has met such a cloud (it reports through a word of mapped host memory, read without synchronisation at
 *      the next launch), the frame kernel from then on; setting the key clears the report.
 *   26 streamed ingest of the replayed host step (0 = off): n > 0 = the clouds arrive in n groups of batch elements,
 *      the search starts with the first H2D copy and its CTAs wait, per batch element, for a flag the copy lane raises
 *      behind each group (where the forward launch is the HMMA grid kernel and n, m are multiples of 32).  Bit-exact;
 *      not faster than the chunked pipeline on the measured platform (~4 us per copy node), hence opt-in.
 *   27 pulled ingest of the replayed host step (0 = auto: 8 CTAs from 4 MB of traffic where the forward launch is the
 *      HMMA grid kernel, n and m are multiples of 32 and the clouds are device-readable pinned memory; -1 = off;
 *      n > 0 = that many ingest CTAs): a few CTAs load the clouds over PCIe, one arrival flag per batch element, the
 *      search starts behind them as a programmatic dependent and runs under the transfer.  Bit-exact.
 *   29 test hook for key 27 (0): 1 = the ingest kernel withholds the arrival flag of the last batch element, so the
 *      search's CTAs for it give up after their patience, the step is redone on the direct path (results stay
 *      correct) and ga_debug_host_streamed() reports -1.
 *   28 ga_knn_dists on knn_slab_kernel (1): 0 = never, 1 = batches of at least half a wave of 512-query CTAs, 2 = always
 *      (clouds of 512..2048 points, k <= 10).  Values are the same bits as knn_kernel's.
 * Clouds far from the origin (relative to their size): the filters' windows scale with (max|q_c| + max|t_c|)^2 measured
 * from the origin of the frame they are evaluated in, so a unit cube at offset 10 costs the plain tensor-core forward
 * 569 us instead of 57 (B=50, 2048 points); the frame kernel takes 63.5 us at any offset, and 57 on centred clouds
 * (profiles/r02_tune_offset.txt).  Results are the same bits in every frame.  The first launch that meets such clouds
 * is a plain one; a captured graph keeps the kernel chosen at capture time (set key 25 = 1 before capturing if the
 * clouds are known to sit away from the origin).  The fp32-filter kernels, kNN and all-pairs keep the origin frame. */
int ga_set_tuning(int key, int value);
/* 1 if the calling thread's last ga_nn_distance_fwd_bwd_host call replayed the streamed pipeline (key 26), -1 if that
 * replay gave up waiting for its inputs and the step was redone on the direct path, 0 otherwise. */
int ga_debug_host_streamed(void);
/* Empty-kernel launch floor in microseconds (average over `reps` launches). */
int ga_probe_launch_floor(int reps, float* us, ga_stream_t stream);
/* Name of the kernel the calling thread launched last through this library ("" before the
 * first launch).  bench.py names the roofline kernel with it instead of assuming the dispatch. */
const char* ga_last_kernel(void);
/* Evidence for the tensor-core filter (forward variant 20, nn_mma.cuh): the raw filter values
 * h(q,t) ~ |t|^2 - 2 q.t of one cloud pair, out[q*m + t], n queries (xyz1), m <= 2048 targets
 * (xyz2), device pointers.  Tests compare it with the fp64 value against the documented bound. */
int ga_debug_mma_filter(int n, int m, const float* xyz1, const float* xyz2, float* out, ga_stream_t stream);
/* Same evidence for the tcgen05 / TMEM filter (forward variant 22, nn_distance_fwd_umma.cu). */
int ga_debug_umma_filter(int n, int m, const float* xyz1, const float* xyz2, float* out, ga_stream_t stream);
/* Development: per-warp clock totals of the following warp-specialised forward launches (variant 24) go to `buf`
 * (device memory, grid x 16 warps x 4 int64: waiting, working, staging, total); NULL switches it off. */
int ga_debug_ws_trace(long long* buf);

#ifdef __cplusplus
}
#endif
#endif /* GA_B200_H_ */
