/* onika_b200.h -- C ABI of the B200-native onika data-parallel execution layer (libonika_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of Collab4exaNBody/onika: onika::parallel
 * (block_parallel_for / parallel_for dispatch on ParallelExecutionContext / ParallelExecutionStream / queue lanes),
 * its block reductions and the SOATL field-array streaming.  The reference has no C ABI of its own; its natural
 * seam is the non-template library code in src/parallel/parallel_execution_queue.cpp (schedule(), sync_and_remove())
 * calling into per-functor template code.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  The C++20 front end in include/onika/ (same names and signatures as the
 * reference headers) and the Python host mirror in onika_b200/ are both thin layers over exactly these symbols.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.  All pointers named *_dev must be device-accessible
 *     (cudaMalloc / managed / mapped pinned memory).  Pointers named *_host are ordinary host memory.
 *   - every function returns ONK_OK (0) or a negative ONK_ERR_* code; onk_last_error() gives the message.  The
 *     reference aborts on error (fatal_error(), src/core/log.cpp:195-204; ONIKA_CU_CHECK_ERRORS,
 *     include/onika/cuda/cuda_error.h:27-42); the C++ front end turns a non-zero status into fatal_error().
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with ONK_ERR_NO_DEVICE.
 *   - `lane` is an execution lane as in the reference (include/onika/parallel/constants.h:36-38):
 *     0..ONK_MAX_LANES-1, ONK_LANE_DEFAULT (-1) means lane 0.  Lane i <-> one non-blocking CUDA stream
 *     (src/cuda/cuda_context.cpp:70-82).  Work queued on one lane executes in order; lanes may overlap.
 *   - all compute entry points are asynchronous with respect to the host unless stated; use onk_lane_sync().
 *   - floating point: unfused IEEE-754 double operations, round-to-nearest (see DESIGN.md "parity rules").
 */
#ifndef ONIKA_B200_H
#define ONIKA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ONK_ABI_VERSION 2   /* 2: onk_launch_ex, onk_xchg_status, onk_cell_sum_fields_f64; exchange buffers of 8 KiB (flagged packets) */

/* ---- status ---- */
#define ONK_OK               0
#define ONK_ERR_NO_DEVICE   -1   /* no usable CUDA device / driver */
#define ONK_ERR_CUDA        -2   /* a CUDA runtime/driver call failed */
#define ONK_ERR_ARG         -3   /* invalid argument */
#define ONK_ERR_STATE       -4   /* onk_init() not called, or object in the wrong state */
#define ONK_ERR_NOMEM       -5

#define ONK_LANE_DEFAULT    -1   /* DEFAULT_EXECUTION_LANE, include/onika/parallel/constants.h:36 */
#define ONK_LANE_ALL        -2   /* UNDEFINED_EXECUTION_LANE used as "every lane" by wait()/schedule_all() (:37) */
#define ONK_MAX_LANES      256   /* MAX_EXECUTION_LANES (:38) */

#define ONK_MEM_DEVICE       0   /* cudaMalloc */
#define ONK_MEM_MANAGED      1   /* cudaMallocManaged: HostAllocationPolicy::CUDA_HOST, include/onika/memory/allocator.h:36-40 */
#define ONK_MEM_PINNED       2   /* cudaHostAlloc(mapped) : staging buffers and host result slots */

#define ONK_OP_MIN           0
#define ONK_OP_MAX           1
#define ONK_OP_SUM           2

int         onk_abi_version(void);
const char* onk_last_error(void);           /* thread-local message of the last failing call */

/* ------------------------------------------------------------------------------------------------------------
 * context / device       replaces CudaContext::default_cuda_ctx (src/cuda/cuda_context.cpp:42-58) and the device
 *                        selection of init_cuda (src/scg-builtin-components/init_cuda.cpp:84-111): one process drives
 *                        one device.
 * ------------------------------------------------------------------------------------------------------------ */
int onk_device_count(int* count);
int onk_init(int device);                   /* idempotent for the same device */
int onk_finalize(void);                     /* syncs and destroys lanes, frees scratch */
int onk_device_info(int* device, int* sm_count, size_t* total_mem, char* name, size_t name_len);
int onk_device_sync(void);                  /* ONIKA_CU_DEVICE_SYNCHRONIZE */

/* ------------------------------------------------------------------------------------------------------------
 * lanes -> streams       replaces CudaContext::getThreadStream(lane) (src/cuda/cuda_context.cpp:70-82),
 *                        ParallelExecutionStreamAutoAllocator (src/parallel/parallel_execution_queue.cpp:32-64),
 *                        ParallelExecutionStream::wait (include/onika/parallel/parallel_execution_stream.h:55-83)
 *                        and ParallelExecutionQueue::query_status (parallel_execution_queue.cpp:420-456).
 * ------------------------------------------------------------------------------------------------------------ */
int onk_lane_stream(int lane, void** cuda_stream);     /* lazily creates the lane's non-blocking stream */
int onk_lane_bind_stream(int lane, void* cuda_stream); /* run the lane on a caller-owned stream (e.g. a framework's) */
int onk_lane_sync(int lane);                           /* ONK_LANE_ALL: every lane */
int onk_lane_query(int lane, int* idle);               /* *idle = 1 when nothing is pending on the lane */
int onk_lane_wait_lane(int lane, int other_lane);      /* lane waits (on device) for what is queued on other_lane */
int onk_lane_host_func(int lane, void (*fn)(void*), void* user);  /* ONIKA_CU_STREAM_HOST_FUNC: hybrid CPU ops ordered in a lane
                                                                     (block_parallel_for_adapter.h:473-490).  fn runs on a worker
                                                                     thread of the library (not on a CUDA driver thread) while the
                                                                     lane waits; it must not call the CUDA API on this lane. */

/* ------------------------------------------------------------------------------------------------------------
 * memory                 replaces GenericHostAllocator::allocate/deallocate/is_gpu_addressable
 *                        (src/cuda/allocator.cpp:98-205) and CudaDeviceStorage<T>::New (include/onika/cuda/device_storage.h:48-58).
 *                        The reference's 12-byte {size,flags} trailer is replaced by a side table keyed by pointer.
 * ------------------------------------------------------------------------------------------------------------ */
int onk_alloc(size_t bytes, size_t align, int kind, void** ptr);
int onk_free(void* ptr, size_t bytes);                 /* bytes must match the allocation (0 = do not check) */
int onk_is_gpu_addressable(const void* ptr, int* yes);
int onk_memcpy(void* dst, const void* src, size_t bytes, int lane);   /* ONIKA_CU_MEMCPY: cudaMemcpyDefault, async */
int onk_memset_bytes(void* dst_dev, int byte, size_t bytes, int lane);/* ONIKA_CU_MEMSET */
int onk_prefetch(const void* ptr, size_t bytes, int to_device, int lane); /* ONIKA_CU_MEM_PREFETCH (managed memory) */

/* ------------------------------------------------------------------------------------------------------------
 * parallel-operation bracket   replaces the CUDA branch of ParallelExecutionQueue::schedule
 *                        (src/parallel/parallel_execution_queue.cpp:316-358): start event, H2D copy of the return-data
 *                        initial value, counter reset, [kernel], epilog callback, D2H copy of return data, stop event;
 *                        and the timing read-back of sync_and_remove (:387-392).
 *                        The device scratch is the reference's GPUKernelExecutionScratch (1 KiB: 8 x u64 counters +
 *                        960 B return data, include/onika/parallel/parallel_execution_context.h:56-66).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct onk_op onk_op;
#define ONK_OP_SCRATCH_BYTES      1024
#define ONK_OP_MAX_RETURN_BYTES    960
/* flags of onk_op_begin (its last argument; 0/1 keep their former meaning) */
#define ONK_OP_RESET_COUNTERS      1   /* zero the 8 scratch counters (work-stealing ticket ...) before the kernel */
#define ONK_OP_NO_TIMING           2   /* do not bracket the operation with timing events: onk_op_elapsed_ms reports 0 (saves ~3 us of
                                          host time per operation; the reference always times, src/parallel/parallel_execution_queue.cpp:328-356) */
int onk_op_create(onk_op** op);
int onk_op_destroy(onk_op* op);
int onk_op_device_scratch(onk_op* op, void** scratch_dev);   /* init_device_scratch(), parallel_execution_context.cpp:102-118 */
int onk_op_begin(onk_op* op, int lane, const void* return_init_host, size_t return_bytes, int flags);
int onk_op_end(onk_op* op, void* return_out_host, size_t return_bytes, void (*epilog)(void*), void* user);
int onk_op_elapsed_ms(onk_op* op, float* ms);                /* valid after the lane was synchronised */

/* generic kernel launch  replaces ONIKA_CU_LAUNCH_KERNEL (include/onika/cuda/cuda.h:264) at its three call sites
 *                        (block_parallel_for_adapter.h:186,197,206).  `cu_function` is a CUfunction handle
 *                        (cudaGetFuncBySymbol in the plugin's translation unit), so device code for user functors
 *                        stays in the plugin, as in the reference. */
int onk_launch(void* cu_function, const unsigned grid[3], const unsigned block[3], size_t shared_bytes, int lane, void** args);
/* the same with launch flags.  ONK_LAUNCH_SUB_BLOCKS: the CTA is block[0] x block[1] threads and holds block[1] independent
 * sub-blocks of block[0] threads, each standing for one "block" of the reference's one-block-per-item launch
 * (block_parallel_for_adapter.h:81-93); the kernel is launched as an explicit 1-CTA cluster, which is how device code
 * (the ONIKA_CU_* vocabulary, include/onika/cuda/cuda.h) tells the two kinds of CTA apart. */
#define ONK_LAUNCH_SUB_BLOCKS      1
int onk_launch_ex(void* cu_function, const unsigned grid[3], const unsigned block[3], size_t shared_bytes, int lane, void** args, int flags);
int onk_launch_count(uint64_t* n);          /* kernels launched by this library since onk_init (all entry points) */

/* lane sequence capture  CUDA-graph capture of everything queued on a set of lanes between begin and end
 *                        (the per-op overhead the reference pays on every schedule(), SURVEY 7 "hard parts" (e)). */
typedef struct onk_graph onk_graph;
int onk_graph_begin(int origin_lane, const int* other_lanes, int n_other);
int onk_graph_end(onk_graph** graph);
/* replays the captured sequence on `lane`.  The captured kernels keep using the reduction scratch of the lanes they were
 * captured on: do not queue reductions directly on those lanes while a replay may be running (wait for the replay first). */
int onk_graph_launch(onk_graph* graph, int lane);
int onk_graph_destroy(onk_graph* graph);

/* ------------------------------------------------------------------------------------------------------------
 * typed fast paths       (none in the reference: its GPU side is three 10-line trampolines, one functor call per
 *                        block: block_parallel_for_adapter.h:51-104.)  Each names the reference *behaviour* it
 *                        reproduces; results are bit-identical to the reference OpenMP path unless noted.
 * ------------------------------------------------------------------------------------------------------------ */

/* reductions.  MIN follows cpu_compute, tests/gpu_reduce_min_fp64.cu:46-55 (identity DBL_MAX, `if(d[i]<m) m=d[i]`,
 * NaNs ignored); MAX is its mirror (identity -DBL_MAX); SUM is the return-data accumulation idiom
 * with a fixed, run-to-run reproducible summation tree (|err| <= 1e-12 * sum|d| against any order).
 * out_dev receives one double.  n == 0 yields the identity. */
int onk_reduce_f64(int op, const double* in_dev, uint64_t n, double* out_dev, int lane);
/* fixed-order combine of `count` partials (cross-GPU: all-gathered per-rank partials; the counterpart of
 * mpi::all_reduce_multi, include/onika/mpi/all_reduce_multi.h:30-37) */
int onk_combine_f64(int op, const double* partials_dev, uint32_t count, double* out_dev, int lane);

/* ---- fused reduction + cross-GPU combine over NVLink peer memory ----
 * The counterpart of "reduce locally, then mpi::all_reduce_multi" (include/onika/mpi/all_reduce_multi.h:30-37) as ONE
 * kernel per rank: the last CTA of the local reduction stores its partial into every peer's exchange buffer (peer
 * stores over NVLink/NVSwitch) as a flagged 16-byte packet (payload and call epoch in the same words: no fences), polls its
 * own buffer for the peers' packets and folds the `world` partials in RANK ORDER -- every rank gets the same bits, sums are
 * reproducible, no collective library call.
 * Exchange buffers are ONK_XCHG_BYTES of device memory per rank (onk_xchg_create), shared between the processes of one
 * node through CUDA IPC: export the local handle, all-gather the handles with any host-side transport, open the peers.
 * All ranks must issue the same sequence of calls on an exchange (reductions, all-reduces, barriers: the call count, kept
 * on the device, is the packet epoch, so a captured lane sequence can be replayed), and all calls on one exchange must go
 * to the same lane (they complete in issue order); use one exchange per lane for independent sequences.
 * A peer that does not show up within ~10 s makes the call's results NaN and the exchange unusable: onk_xchg_status
 * reports the failed call once the lane was synchronised, every later call on the exchange returns ONK_ERR_STATE. */
#define ONK_XCHG_BYTES      8192
#define ONK_XCHG_MAX_WORLD    16
#define ONK_XCHG_MAX_VALUES    8
#define ONK_IPC_HANDLE_BYTES  64
typedef struct onk_xchg onk_xchg;
/* where the exchange buffers live.  DEVICE (default): device memory shared through CUDA IPC, peer stores over NVLink / NVSwitch.
 * HOST: pinned host memory shared through POSIX shared memory and mapped into every rank's device address space -- the same
 * packets travel over PCIe, a few microseconds slower per exchange, and no peer access between the GPUs is needed: the
 * fallback for boxes (or containers) where cudaIpcOpenMemHandle is refused.  All ranks of an exchange use the same transport;
 * onk_xchg_create takes it from the environment (ONK_XCHG_TRANSPORT=host), onk_xchg_create_ex from its argument. */
#define ONK_XCHG_TRANSPORT_DEVICE 0
#define ONK_XCHG_TRANSPORT_HOST   1
int onk_xchg_create(int rank, int world, onk_xchg** x);
int onk_xchg_create_ex(int rank, int world, int transport, onk_xchg** x);
int onk_xchg_export(onk_xchg* x, unsigned char handle[ONK_IPC_HANDLE_BYTES]);
int onk_xchg_connect(onk_xchg* x, const unsigned char* handles /* world x ONK_IPC_HANDLE_BYTES, rank order */);
int onk_xchg_destroy(onk_xchg* x);
int onk_reduce_xchg_f64(int op, const double* in_dev, uint64_t n, double* out_dev, onk_xchg* x, int lane);
/* mpi::all_reduce_multi (include/onika/mpi/all_reduce_multi.h:30-37: k scalars packed into one MPI_Allreduce with one
 * operator) for 1..ONK_XCHG_MAX_VALUES doubles in device memory, in place, folded in rank order; one 32-thread kernel
 * per rank, no collective library call.  Collective, counts as one epoch. */
int onk_xchg_allreduce_f64(onk_xchg* x, int op, double* values_dev, int count, int lane);
int onk_xchg_info(onk_xchg* x, int* rank, int* world);
int onk_xchg_status(onk_xchg* x, uint64_t* failed_call);   /* 0 = healthy; else the 1-based call in which a peer timed out */
/* device-side barrier between the ranks' lanes (same flag exchange, no payload): what every rank queued on `lane` before
 * it, peer stores included, is complete and visible once it has passed on any rank.  Collective, counts as one epoch. */
int onk_xchg_barrier(onk_xchg* x, int lane);

/* ---- id-based redistribution between the GPUs of a node ----
 * Replaces communication_scheme_from_ids (src/mpi/xs_data_move.cpp:42-462; include/onika/mpi/xs_data_move.h:31-50) +
 * data_move (include/onika/mpi/xs_data_move_impl.h:44-113): after[d][j] = before[s][i] wherever ids_after[d][j] ==
 * ids_before[s][i].  Instead of index tables + pack / MPI_Alltoallv / unpack, a plan holds one ROUTE (destination rank,
 * destination index) per local element, found through a directory distributed by id range over the ranks (the reference's
 * rendezvous ranks: xs_data_move_range_utils.h:39-63), and a move is one kernel per rank storing every element straight
 * into the destination rank's output window over NVLink peer memory.  Windows are device allocations mapped into every
 * peer process through CUDA IPC (export / all-gather the handles with any host transport / connect, like onk_xchg).
 * All calls below except the *_ptr / *_send_counts accessors are collective over the ranks of `x`.
 * onk_xs_plan_build returns ONK_ERR_ARG on EVERY rank (and the code in *status) when the ids are not a bijection over
 * [id_min,id_max): 1 id of the range without source, 2 id without destination, 3 duplicate id before, 4 duplicate id
 * after, 5 id out of range (the highest code present wins) -- the cases on which the reference aborts
 * (xs_data_move.cpp:286-289,331-334,366-369). */
typedef struct onk_window onk_window;
int onk_window_create(size_t bytes, int rank, int world, onk_window** w);
int onk_window_export(onk_window* w, unsigned char handle[ONK_IPC_HANDLE_BYTES]);
int onk_window_connect(onk_window* w, const unsigned char* handles /* world x ONK_IPC_HANDLE_BYTES, rank order */);
int onk_window_ptr(onk_window* w, void** local_dev, size_t* bytes);
int onk_window_destroy(onk_window* w);
typedef struct onk_xs_plan onk_xs_plan;
int onk_xs_plan_create(onk_xchg* x, uint64_t id_min /* included */, uint64_t id_max /* excluded */, onk_xs_plan** plan);
int onk_xs_plan_export(onk_xs_plan* plan, unsigned char handle[ONK_IPC_HANDLE_BYTES]);
int onk_xs_plan_connect(onk_xs_plan* plan, const unsigned char* handles);
int onk_xs_plan_build(onk_xs_plan* plan, const uint64_t* ids_before_dev, uint64_t n_before,
                      const uint64_t* ids_after_dev, uint64_t n_after, int lane, int* status);
int onk_xs_plan_send_counts(onk_xs_plan* plan, uint64_t* send_count_host /* world entries: the reference's send_count */);
int onk_xs_move(onk_xs_plan* plan, onk_window* out, uint64_t elem_bytes, const void* in_dev, int lane);
/* the same for several arrays that move with the same ids (the fields of a particle): one barrier before, one after, one
 * scatter kernel per array -- what the reference does by calling data_move once per array with the same tables */
int onk_xs_move_fields(onk_xs_plan* plan, int nfields, onk_window* const* outs, const uint64_t* elem_bytes, const void* const* ins_dev, int lane);
int onk_xs_plan_destroy(onk_xs_plan* plan);

/* parallel_for(N, y[i] += a*x[i])   include/onika/parallel/parallel_for.h:91-112 (config C1) */
int onk_axpy_f64(double* y_dev, const double* x_dev, double a, uint64_t n, int lane);
/* parallel_memset<T>                include/onika/parallel/memset.h:30-51 ; elem_bytes in {1,2,4,8}, value = low bytes of pattern */
int onk_parallel_memset(void* data_dev, uint64_t n, uint64_t pattern, int elem_bytes, int lane);
/* block_parallel_for(rows, BlockParallelValueAddFunctor)   include/onika/extras/block_parallel_value_add_functor.h:15-58 */
int onk_value_add_f64(double* array_dev, uint64_t rows, uint64_t cols, double value, int lane);

/* ---- SOATL (include/onika/soatl) ---- */
/* ChunkLogAllocationStrategy::update_capacity, include/onika/memory/alloc_strategy.h:44-67 (host arithmetic) */
uint64_t onk_soatl_update_capacity(uint64_t size, uint64_t capacity, uint64_t chunk);
/* column offsets of FieldArrays<A,C,ids...> for a given capacity, field_arrays.h:494-512; returns storage bytes */
uint64_t onk_soatl_layout(uint64_t align, const uint32_t* field_bytes, int nfields, uint64_t capacity, uint64_t* offsets);
/* soatl::parallel_apply_simd over nfields fp64 columns of one packed allocation: f_k[i] = f_k[i]*s + c_k
 * (config C3; compute.h:214-232 semantics incl. the ceil(n/chunk)*chunk padding sweep) */
int onk_soatl_rmw_f64(void* storage_dev, uint64_t align, uint64_t chunk, int nfields, uint64_t n, uint64_t capacity,
                      double s, const double* c_host, int lane);
/* the stock benchmark kernel tests/benchmark.cpp:96-104 (3 reads + 1 write), columns given by pointer */
int onk_soatl_dist_f64(double* e_dev, const double* rx_dev, const double* ry_dev, const double* rz_dev,
                       uint64_t n_padded, double ax, double ay, double az, int lane);
/* soatl::copy between two layouts of the same field list (include/onika/soatl/copy.h:29-97;
 * tests/soatlserializetest.cpp:86-104): e.g. <64,8,...> <-> the <1,1,...> wire format */
int onk_soatl_repack(void* dst_dev, uint64_t dst_align, uint64_t dst_capacity,
                     const void* src_dev, uint64_t src_align, uint64_t src_capacity,
                     const uint32_t* field_bytes, int nfields, uint64_t n, int lane);

/* ---- cell grid (config C4): cells = SOATL blocks packed back to back in one arena ---- */
/* capacity/offset of every cell (host arithmetic; cell c's block starts at cell_off[c], aligned to cell_align) */
uint64_t onk_cell_grid_layout(const uint32_t* cell_size, uint64_t ncells, uint64_t align, uint64_t chunk,
                              const uint32_t* field_bytes, int nfields, uint64_t cell_align,
                              uint32_t* cell_cap, uint64_t* cell_off);
/* block_parallel_for(ncells, per-cell sum of an fp64 field) + global sum.  cell_sum_dev[c] is the in-order sum of the
 * cell's particles for sizes <= 32 and a fixed tree above (|err| <= 1e-12 relative); total_dev (optional) is the
 * fixed-order sum of the cell sums.  field_offset_fn: the field is column `field` of FieldArrays<align,chunk,...>. */
int onk_cell_sum_f64(const void* arena_dev, const uint64_t* cell_off_dev, const uint32_t* cell_size_dev,
                     const uint32_t* cell_cap_dev, uint64_t ncells, uint64_t align,
                     const uint32_t* field_bytes, int nfields, int field,
                     double* cell_sum_dev, double* total_dev, int lane);
/* per-cell sums of EVERY field of the cell blocks (all fields fp64): cell_sums_dev[c*nfields + k] is the in-order sum of field k
 * over the particles of cell c (bit-identical to the host loop), totals_dev[k] (optional) the fixed-order sum over the cells.
 * With all fields wanted the per-cell layout is one contiguous stream (32 B per particle for rx,ry,rz,e: SURVEY 8d, C4 "all 4
 * fields streamed"); arena_bytes = size of the arena allocation (the layout value returned by onk_cell_grid_layout). */
int onk_cell_sum_fields_f64(const void* arena_dev, uint64_t arena_bytes, const uint64_t* cell_off_dev, const uint32_t* cell_size_dev,
                            const uint32_t* cell_cap_dev, uint64_t ncells, uint64_t align, const uint32_t* field_bytes, int nfields,
                            double* cell_sums_dev, double* totals_dev, int lane);
/* the same per-cell sum over a FIELD-MAJOR grid: `values_dev` is the summed field of all cells back to back, cell c owns
 * [cell_begin[c], cell_begin[c+1]) (ncells+1 offsets).  Same per-cell bits as the loop above (particle order); the layout
 * turns the kernel into a pure stream (8 B per particle + 16 B per cell of traffic, no padding between cells). */
int onk_cell_sum_packed_f64(const double* values_dev, uint64_t nvalues /* length of the column = cell_begin[ncells]; must not exceed the allocation (smaller is safe, only slower) */,
                            const uint64_t* cell_begin_dev, uint64_t ncells, double* cell_sum_dev, double* total_dev, int lane);

/* ------------------------------------------------------------------------------------------------------------
 * host-buffer entry points   what a reference plugin holding HOST data calls: input in ordinary host memory, result in
 *                        host memory, synchronous.  Host<->device copies are chunked and overlapped with the kernels on
 *                        two internal lanes (pinned staging).  These are the `e2e` calls of bench.py.
 * ------------------------------------------------------------------------------------------------------------ */
int onk_host_reduce_f64(int op, const double* in_host, uint64_t n, double* out_host);
int onk_host_axpy_f64(double* y_host, const double* x_host, double a, uint64_t n);
int onk_host_value_add_f64(double* array_host, uint64_t rows, uint64_t cols, double value);

#ifdef __cplusplus
}
#endif
#endif /* ONIKA_B200_H */
