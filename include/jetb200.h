/* jetb200.h — C ABI of the B200-native contraction engine behind the Jet API.
 *
 * This is the drop-in boundary for ONE path of XanaduAI/jet: pairwise tensor contraction
 * (index permutation + complex GEMM) driven per slice over a sliced tensor network.  Every entry
 * point cites the reference interface it replaces (paths relative to the reference repository).
 *
 * Conventions
 *   - All functions return 0 on success, non-zero on failure; jb_last_error() returns the message
 *     of the calling thread's last failure (the C++ layer rethrows it as Jet::Exception, mirroring
 *     include/jet/Abort.hpp:51-97).
 *   - dtype: JB_C64 = std::complex<float>, JB_C128 = std::complex<double> (the only two element
 *     types the reference supports, include/jet/Tensor.hpp:32-34).
 *   - Tensors are dense, row-major over their extents (last axis fastest), like Jet::Tensor
 *     (include/jet/Tensor.hpp:766-775).
 *   - Pointers named d_* are DEVICE pointers, h_* are HOST pointers.  Kernels never allocate: all
 *     buffers, including workspaces, are supplied by the caller (or owned by a jb_plan).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - All entry points are re-entrant; there is no hidden global stream or handle.
 *   - There is NO CPU fallback: every call fails loudly if no CUDA device is usable.
 */
#ifndef JETB200_H
#define JETB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_C64 0
#define JB_C128 1
#define JB_MAX_RANK 64

/* ---- status / device --------------------------------------------------------------------- */
const char *jb_last_error(void);
const char *jb_version(void);
int jb_device_count(int *count);
int jb_set_device(int device);
int jb_device_info(int device, int *sm_count, size_t *total_bytes, size_t *l2_bytes, int *cc_major,
                   int *cc_minor);

/* ---- device memory + streams (what cudaMalloc/cudaMemcpy are to include/jet/CudaTensor.hpp:158,
 *      185,307; here explicit, asynchronous on a stream, and never called inside an operator) --- */
int jb_malloc(void **d_ptr, size_t bytes);
int jb_free(void *d_ptr);
int jb_host_alloc(void **h_ptr, size_t bytes); /* pinned */
int jb_host_free(void *h_ptr);
int jb_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream);
int jb_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream);
int jb_memcpy_d2d(void *d_dst, const void *d_src, size_t bytes, void *stream);
int jb_memset_zero(void *d_dst, size_t bytes, void *stream);
int jb_stream_create(void **stream);
int jb_stream_destroy(void *stream);
int jb_stream_sync(void *stream);

/* ---- K1: index permutation ----------------------------------------------------------------
 * Replaces Permuter<Backend>::Transpose(data_in, shape, data_out, current_order, new_order)
 * (include/jet/permute/Permuter.hpp:50-79; back ends permute/QFlex.hpp:37-50 and
 * permute/Default.hpp:21-133), as called from Tensor::Transpose (include/jet/Tensor.hpp:594-611).
 * extent_in[i] is the extent of input axis i; output axis j is input axis perm[j].
 * Bit-exact (pure data movement).  in and out must not overlap. */
int jb_permute(int dtype, const void *d_in, void *d_out, int rank, const int64_t *extent_in,
               const int32_t *perm, void *stream);

/* ---- K2: row-major complex GEMM  C(MxN) = A(MxK) * B(KxN), alpha = 1, beta = 0 ---------------
 * Replaces gemmBinding / gemvBinding / dotuBinding -> cblas_{c,z}gemm / gemv / dotu_sub
 * (include/jet/TensorHelpers.hpp:49-61,79-89,103-111) as dispatched by MultiplyTensorData
 * (include/jet/TensorHelpers.hpp:131-168): lda = K, ldb = ldc = N; M == 1 and/or N == 1 are the
 * GEMV / DOTU (unconjugated) corners.  jb_gemm_ws_bytes gives the split-K workspace size
 * (may be 0); d_ws may be NULL when it is 0. */
size_t jb_gemm_ws_bytes(int dtype, int64_t m, int64_t n, int64_t k);
int jb_gemm(int dtype, int64_t m, int64_t n, int64_t k, const void *d_a, const void *d_b,
            void *d_c, void *d_ws, size_t ws_bytes, void *stream);

/* ---- fused pairwise contraction ---------------------------------------------------------------
 * Replaces Tensor::ContractTensors(A, B) (include/jet/Tensor.hpp:709-752): modes are integer
 * labels; equal labels are summed over (no conjugation).  The output is ordered
 * (modes of A not in B, in A's order) ++ (modes of B not in A, in B's order) exactly as
 * Tensor.hpp:732-741.  jb_contract_info reports the output rank/modes/extents, the GEMM view
 * (M, N, K), the workspace the chosen kernel needs and which kernel family was chosen
 * (0 = streaming FMA kernel with the small operand resident in shared memory,
 *  1 = permute + GEMM through the workspace). */
typedef struct jb_contract_info_t {
    int32_t rank_c;
    int32_t modes_c[JB_MAX_RANK];
    int64_t extent_c[JB_MAX_RANK];
    int64_t m, n, k;
    size_t ws_bytes;
    int32_t kernel;
} jb_contract_info_t;

int jb_contract_info(int dtype, int rank_a, const int64_t *extent_a, const int32_t *modes_a,
                     int rank_b, const int64_t *extent_b, const int32_t *modes_b,
                     jb_contract_info_t *info);
int jb_contract(int dtype, int rank_a, const int64_t *extent_a, const int32_t *modes_a,
                const void *d_a, int rank_b, const int64_t *extent_b, const int32_t *modes_b,
                const void *d_b, void *d_c, void *d_ws, size_t ws_bytes, void *stream);

/* ---- fused contraction chain ------------------------------------------------------------------
 * Replaces a RUN of Tensor::ContractTensors calls (include/jet/Tensor.hpp:709-752) of the kind
 * TensorNetwork::Contract issues on consecutive path steps (include/jet/TensorNetwork.hpp:301-328,
 * 394-421): X_i = ContractTensors(X_{i-1}, R_i) when x_is_left[i] != 0, else
 * ContractTensors(R_i, X_{i-1}), for i = 1..n_steps, with every R_i small (<= 256 elements, at
 * most 16 contracted and 16 free values) and all extents powers of two.  One kernel launch keeps
 * the intermediates X_1..X_{k-1} in shared memory; only X_0 and the R_i are read and X_k written.
 * The result (labels, extents, values within FMA rounding) equals the step-by-step calls.
 * jb_chain_info fails (non-zero, message "chain: ...") when the run does not fit one on-chip tile;
 * callers then split the run or fall back to jb_contract per step. */
typedef struct jb_chain_desc_t {
    int32_t dtype;
    int32_t n_steps;
    int32_t rank_x;
    const int64_t *extent_x; /* [rank_x] */
    const int32_t *modes_x;  /* [rank_x] */
    const int32_t *rank_r;   /* [n_steps] */
    const int64_t *extent_r; /* [sum rank_r] */
    const int32_t *modes_r;  /* [sum rank_r] */
    const int32_t *x_is_left; /* [n_steps] */
} jb_chain_desc_t;

typedef struct jb_chain_info_t {
    int32_t rank_c;
    int32_t modes_c[JB_MAX_RANK];
    int64_t extent_c[JB_MAX_RANK];
    int32_t log_tile;      /* the on-chip tile holds 2^log_tile elements */
    int32_t conflict_free; /* 1 if every shared-memory phase is bank-conflict-free by construction */
    int32_t n_stages;      /* barrier-separated stages (runs of K == N steps share one stage) */
    int32_t pad;
    double flops;          /* 8*M*N*K summed over the steps */
    double bytes;          /* sizeof(T)*(|X_0| + sum |R_i| + |X_k|): traffic of the fused launch */
    double step_bytes;     /* sizeof(T)*(MK+KN+MN) summed over the steps: traffic step by step */
} jb_chain_info_t;

int jb_chain_info(const jb_chain_desc_t *desc, jb_chain_info_t *info);
int jb_contract_chain(const jb_chain_desc_t *desc, const void *d_x, const void *const *d_r,
                      void *d_out, void *stream);
int jb_contract_chain_host(const jb_chain_desc_t *desc, const void *h_x, const void *const *h_r,
                           void *h_out);

/* ---- elementwise helpers --------------------------------------------------------------------
 * jb_add: c = a + b elementwise over n complex elements; the aligned-add at the core of
 * Tensor::AddTensors (include/jet/Tensor.hpp:433-451) — permute b first with jb_permute when its
 * index order differs.  jb_slice: fixes axis `axis` of `in` to `value` and writes the rank-1
 * smaller tensor, Tensor::SliceIndex (include/jet/Tensor.hpp:494-526).  jb_conj: elementwise
 * conjugate (Tensor::Conj, include/jet/Tensor.hpp:668-675). */
int jb_add(int dtype, int64_t n, const void *d_a, const void *d_b, void *d_c, void *stream);
int jb_slice(int dtype, const void *d_in, void *d_out, int rank, const int64_t *extent_in,
             int axis, int64_t value, void *stream);
int jb_conj(int dtype, int64_t n, const void *d_in, void *d_out, void *stream);

/* ---- host-buffer forms (what a Jet::Tensor method calls: host std::vector in, host std::vector
 *      out; H2D copy, kernel(s), D2H copy and a stream sync happen inside) ---------------------- */
int jb_permute_host(int dtype, const void *h_in, void *h_out, int rank, const int64_t *extent_in,
                    const int32_t *perm);
int jb_contract_host(int dtype, int rank_a, const int64_t *extent_a, const int32_t *modes_a,
                     const void *h_a, int rank_b, const int64_t *extent_b, const int32_t *modes_b,
                     const void *h_b, void *h_c);
int jb_gemm_host(int dtype, int64_t m, int64_t n, int64_t k, const void *h_a, const void *h_b,
                 void *h_c);
int jb_add_host(int dtype, int64_t n, const void *h_a, const void *h_b, void *h_c);
int jb_slice_host(int dtype, const void *h_in, void *h_out, int rank, const int64_t *extent_in,
                  int axis, int64_t value);
int jb_conj_host(int dtype, int64_t n, const void *h_in, void *h_out);

/* ---- contraction plan: a whole (sliced) network + path on one GPU ----------------------------
 * Replaces the per-task execution of TaskBasedContractor (AddContractionTasks / AddReductionTask /
 * AddDeletionTasks / Contract, include/jet/TaskBasedContractor.hpp:162-322) and the per-slice
 * network copies made by TensorNetwork::SliceIndices (include/jet/TensorNetwork.hpp:210-284):
 * the leaves live once in a device arena, slices are views selected on the device, every path
 * step is a CUDA-graph kernel node, intermediates get arena offsets from their lifetimes, and the
 * per-slice results are accumulated on the device in double precision.
 *
 * Network description: leaf i has rank[i] axes with extents extent[off_i..] and integer mode
 * labels mode[off_i..] (off_i = sum of earlier ranks), data h_data[i] (host, row-major, dtype).
 * path is num_steps pairs (a, b) of node ids; step s creates node num_leaves + s
 * (include/jet/TensorNetwork.hpp:301-328).  sliced_modes lists the modes fixed per slice; a slice
 * id is the row-major ravel over them, first listed mode slowest (TensorNetwork.hpp:210-230).
 */
typedef struct jb_plan jb_plan;

typedef struct jb_network_desc_t {
    int32_t dtype;
    int32_t device;
    int32_t num_leaves;
    const int32_t *rank;   /* [num_leaves] */
    const int64_t *extent; /* [sum rank] */
    const int32_t *mode;   /* [sum rank] */
    const void *const *h_data; /* [num_leaves] */
    int32_t num_steps;
    const int32_t *path; /* [2 * num_steps] */
    int32_t num_sliced;
    const int32_t *sliced_modes; /* [num_sliced] */
    int32_t flags;               /* JB_PLAN_* */
    int32_t batch;               /* slices contracted per launch: 0 = automatic (small per-slice tensors only), 1 = off,
                                    n = at most n (rounded down to a power of two, bounded by memory) */
} jb_network_desc_t;

#define JB_PLAN_KEEP_INTERMEDIATES 1 /* no buffer reuse: every step output stays readable */
#define JB_PLAN_NO_GRAPH 2           /* launch kernels directly instead of through a CUDA graph */
#define JB_PLAN_STORE_RESULTS 4      /* keep every slice's own result (for GetResults()) */
#define JB_PLAN_NO_FUSE 8            /* one kernel per path step: no fused contraction chains */
#define JB_PLAN_DRY_RUN 16           /* host-side planning only (no device needed): stats / steps / ops work,
                                        everything that would touch the GPU fails */

typedef struct jb_plan_stats_t {
    int64_t num_slices;        /* product of sliced extents */
    int64_t result_elems;      /* elements of the final tensor */
    int32_t result_rank;
    int32_t result_modes[JB_MAX_RANK];
    int64_t result_extent[JB_MAX_RANK];
    int32_t steps_total;       /* path steps */
    int32_t steps_shared;      /* slice-independent steps (run once, not per slice) */
    int32_t steps_stream;      /* per-slice steps on the streaming kernel */
    int32_t steps_ttgt;        /* per-slice steps on permute + GEMM */
    int32_t steps_chained;     /* per-slice steps executed inside fused chains */
    int32_t chains;            /* fused chains per slice */
    int32_t launches_per_slice; /* kernel launches in one slice's graph */
    double flops_per_slice;    /* 8*M*N*K summed over per-slice steps */
    double bytes_per_slice;    /* sizeof(T)*(MK+KN+MN) summed over per-slice steps */
    double fused_bytes_per_slice; /* what the launch units of one slice must move (chains fused) */
    double flops_shared, bytes_shared;
    double jet_flops_per_slice; /* 2*M*N*K over ALL steps: PathInfo::GetTotalFlops convention */
    size_t arena_bytes;        /* device memory reserved */
    int64_t max_step_elems;    /* largest intermediate */
    int32_t batch;             /* slices contracted per launch (slice batching; 1 = off) */
    int32_t pad;
} jb_plan_stats_t;

int jb_plan_create(const jb_network_desc_t *desc, jb_plan **plan);
int jb_plan_destroy(jb_plan *plan);
int jb_plan_stats(const jb_plan *plan, jb_plan_stats_t *stats);
/* Upload leaf data again (H2D) — the host->device leg of an end-to-end run. */
int jb_plan_upload(jb_plan *plan, const void *const *h_data);
/* Zero the accumulator, (re)run the shared steps if needed. */
int jb_plan_reset(jb_plan *plan);
/* Enqueue `count` slices: ids first, first+1, ... (asynchronous; no host sync). */
int jb_plan_run(jb_plan *plan, int64_t first_slice, int64_t count);
/* Enqueue an explicit list of slice ids. */
int jb_plan_run_list(jb_plan *plan, const int64_t *slice_ids, int64_t count);
/* Wait, then read the double-precision sum over all slices run since the last reset:
 * h_out receives result_elems (re, im) pairs of doubles. */
int jb_plan_result(jb_plan *plan, double *h_out);
/* With JB_PLAN_STORE_RESULTS: result of the ordinal-th slice run since reset, in dtype. */
int jb_plan_slice_result(jb_plan *plan, int64_t ordinal, void *h_out);
/* The same for `count` consecutive ordinals in one copy (what GetResults() of the task-based contractor reads). */
int jb_plan_slice_results(jb_plan *plan, int64_t first_ordinal, int64_t count, void *h_out);
/* With JB_PLAN_KEEP_INTERMEDIATES: copy node `node`'s tensor (last slice run) to the host. */
int jb_plan_node(jb_plan *plan, int32_t node, void *h_out, int64_t *elems);
/* Modes and extents of node `node` as the steps see it (a sliced leaf: after slicing); modes / extent may be NULL,
 * otherwise they receive up to JB_MAX_RANK entries. */
int jb_plan_node_info(const jb_plan *plan, int32_t node, int32_t *rank, int32_t *modes, int64_t *extent);
int jb_plan_sync(jb_plan *plan);
/* Device time (ms) between the first and last kernel of the most recent jb_plan_run*, measured
 * with CUDA events on the plan's stream. */
int jb_plan_last_ms(jb_plan *plan, float *ms);
/* The stream the plan launches on (cudaStream_t). */
int jb_plan_stream(jb_plan *plan, void **stream);
/* A second plan of the same network on `device` (any device of this process): the host-side plan (steps,
 * fused chains, arena offsets) is copied, the arena / stream / CUDA graphs are new; upload the leaves with
 * jb_plan_upload.  What LanePlans / SlicedContractor lanes and the multi-device set below are built from. */
int jb_plan_clone(const jb_plan *plan, int device, jb_plan **clone);
/* Device pointer of the plan's FP64 accumulator: `elems` (re, im) pairs of doubles, valid in stream order on
 * jb_plan_stream — the buffer a reduction over GPUs reads (jb_reduce_sum), with no host staging. */
int jb_plan_accumulator(jb_plan *plan, void **d_acc, int64_t *elems);
int jb_plan_device(const jb_plan *plan, int *device);
/* Per-step description for roofline accounting: fills up to cap entries. */
typedef struct jb_step_info_t {
    int32_t node_a, node_b, node_c;
    int32_t shared;  /* 1 = slice-independent */
    int32_t kernel;  /* 0 stream, 1 ttgt */
    int64_t m, n, k;
    double flops, bytes;
    int32_t op;      /* index of the launch unit (jb_plan_ops) that executes this step */
    int32_t pad;
} jb_step_info_t;
int jb_plan_steps(const jb_plan *plan, jb_step_info_t *steps, int32_t cap, int32_t *count);
/* Time every per-slice step individually (CUDA events, `reps` repetitions on the given slice);
 * ms[i] is the mean device time of step i (0 for shared steps). */
int jb_plan_profile(jb_plan *plan, int64_t slice, int reps, float *ms, int32_t cap);
/* Launch units of one slice in execution order: a unit is one path step or a fused chain of path
 * steps.  bytes = what the unit must move (for a chain: first input + small operands + last
 * output); step_bytes = the step-by-step figure sizeof(T)*(MK+KN+MN) summed over its steps. */
#define JB_GEMM_FMA 0      /* GemmKernel (+ SplitKReduceKernel): FP32 / FP64 FMA */
#define JB_GEMM_SMALL_MN 1 /* SmallMnKernel: DOTU / GEMV corner with a long K */
#define JB_GEMM_TCGEN05 2  /* GemmTf32x3Kernel: tcgen05 3xTF32 (complex64) */
#define JB_GEMM_DMMA 3     /* GemmDmmaKernel: FP64 tensor pipe (complex128) */
#define JB_GEMM_DOT_GATHER 4   /* DotGatherKernel: DOTU / GEMV corner, both operands read in their original layouts */
#define JB_GEMM_SMALL_GATHER 5 /* SmallGemmGatherKernel: M, N <= 16, moderate K, operands in place, batched over slices */
typedef struct jb_op_info_t {
    int32_t kernel;     /* 0 stream, 1 ttgt, 2 fused chain */
    int32_t n_steps;    /* path steps executed by this unit */
    int32_t first_step; /* path index of its first step */
    int32_t last_step;
    int32_t log_tile;   /* fused chain: log2 of the shared-memory tile (elements) */
    int32_t launches;
    int32_t n_stages;   /* fused chain: barrier-separated stages */
    int32_t gemm_kind;  /* ttgt: JB_GEMM_* — the GEMM kernel this unit launches */
    double flops, bytes, step_bytes;
    int32_t register_steps; /* fused chain: steps executed inside register stages (the rest go through shared memory) */
    int32_t pad;
} jb_op_info_t;
int jb_plan_ops(const jb_plan *plan, jb_op_info_t *ops, int32_t cap, int32_t *count);
/* Like jb_plan_profile, per launch unit: ms[i] is the mean device time of unit i. */
int jb_plan_profile_ops(jb_plan *plan, int64_t slice, int reps, float *ms, int32_t cap);


/* ---- one network over several plans: lanes x devices of ONE process -----------------------------------------
 * Replaces the Taskflow executor's worker pool (include/jet/TaskBasedContractor.hpp:322: tasks of different
 * slices run on different workers) and the tf::reduce over the per-slice results
 * (TaskBasedContractor.hpp:258-280): `lanes` plans per device keep that many slices in flight on one GPU
 * (one arena, stream and CUDA graph each), every listed device gets such a group, a run deals contiguous
 * blocks of the slice range to the plans in (device, lane) order, and the FP64 partial sums are added on the
 * devices in that fixed order (peer copies over NVLink) — deterministic for a given (devices, lanes).
 * devices == NULL means desc->device, desc->device + 1, ...; lanes == 0 picks the lane count from the plan's
 * arena size (4 lanes below 512 MiB, 2 below 8 GiB, else 1; never more than fit in free device memory). */
typedef struct jb_multi jb_multi;
int jb_multi_create(const jb_network_desc_t *desc, int num_devices, const int *devices, int lanes,
                    jb_multi **multi);
int jb_multi_destroy(jb_multi *multi);
int jb_multi_stats(const jb_multi *multi, jb_plan_stats_t *stats); /* of one plan */
int jb_multi_num_plans(const jb_multi *multi, int *num_devices, int *lanes);
int jb_multi_plan(jb_multi *multi, int index, jb_plan **plan); /* index = device_pos * lanes + lane */
int jb_multi_upload(jb_multi *multi, const void *const *h_data);
int jb_multi_reset(jb_multi *multi);
/* reset + enqueue slices [first, first + count) / the listed ids on all plans (asynchronous) */
int jb_multi_run(jb_multi *multi, int64_t first_slice, int64_t count);
int jb_multi_run_list(jb_multi *multi, const int64_t *slice_ids, int64_t count);
int jb_multi_sync(jb_multi *multi);
/* Sum over everything run since the reset (after jb_multi_reduce: over all ranks), as doubles. */
int jb_multi_result(jb_multi *multi, double *h_out);
/* With JB_PLAN_STORE_RESULTS: the result of the ordinal-th slice of the last run. */
int jb_multi_slice_result(jb_multi *multi, int64_t ordinal, void *h_out);
/* ... and of ALL slices of the last run, in run order (count * result_elems elements of dtype). */
int jb_multi_slice_results(jb_multi *multi, void *h_out);
int jb_multi_last_ms(jb_multi *multi, float *ms); /* the slowest plan */

/* ---- reduction across processes (one process per GPU): NCCL over NVLink / NVSwitch -----------------------------
 * The reference has no multi-process layer; its sliced runs sum per-slice results with tf::reduce inside one
 * process (TaskBasedContractor.hpp:258-280).  Here every rank contracts its block of slices and ONE
 * ncclReduce of the FP64 partial sums (enqueued on the plan's stream, device buffers in and out) ends the job.
 * NCCL is bound at run time (dlopen of libnccl.so.2; JB_NCCL_LIB overrides the path) — a process that already
 * holds NCCL (torch.distributed) shares its copy.  Rank 0 calls jb_comm_unique_id and distributes the
 * JB_COMM_ID_BYTES bytes by any means (MPI, a file, torch.distributed); then every rank calls jb_comm_create. */
#define JB_COMM_ID_BYTES 128
typedef struct jb_comm jb_comm;
int jb_comm_unique_id(void *id128);
int jb_comm_create(int world, int rank, const void *id128, int device, jb_comm **comm);
int jb_comm_destroy(jb_comm *comm);
int jb_comm_info(const jb_comm *comm, int *world, int *rank, int *nccl_version);
/* In-place sum of n_doubles doubles at d_buf over all ranks, into rank `root` (root < 0: into every rank),
 * asynchronous on `stream`. */
int jb_reduce_sum(jb_comm *comm, void *d_buf, int64_t n_doubles, int root, void *stream);
/* The set's on-device total (see jb_multi_result) reduced over the ranks of `comm`, which must live on the
 * set's first device; afterwards jb_multi_result returns the reduced sum on `root`. */
int jb_multi_reduce(jb_multi *multi, jb_comm *comm, int root);


/* ---- peak probes: roofline denominators measured on the device the caller is on ------------------------------
 * No reference counterpart (SURVEY 8(d) asks for FP64 / FP32 / TF32 peaks next to the HBM copy bandwidth).
 * value = TFLOP/s of a dependency-free register-only instruction stream on every SM (kinds 0-3) or GB/s of a
 * 1 GiB device copy, read + write (kind 4); best of three timed launches, CUDA events. */
#define JB_PEAK_FP32_FMA 0      /* fma.rn.f32x2 (FFMA2): what the chain / stream kernels issue */
#define JB_PEAK_FP64_FMA 1      /* fma.rn.f64 */
#define JB_PEAK_FP64_DMMA 2     /* mma.sync.m8n8k4.f64: the FP64 tensor pipe of GemmDmmaKernel */
#define JB_PEAK_TF32_MMA_SYNC 3 /* mma.sync.m16n8k8.tf32: the legacy warp-level tensor path (not tcgen05) */
#define JB_PEAK_HBM_COPY 4
int jb_probe_peak(int kind, double *value);

#ifdef __cplusplus
}
#endif
#endif /* JETB200_H */
