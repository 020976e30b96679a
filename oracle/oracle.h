/* oracle/oracle.h -- TEST INFRASTRUCTURE: CPU restatement of the reference algorithms on the hot path.
 *
 * Plain C11 + OpenMP, one function per reference behaviour, each citing the reference file:line it follows
 * (paths relative to the reference tree root).  Pinned against the real reference (oracle/_ref, built from the
 * reference's own sources) and against tests/golden/ by tests/test_oracle_*.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs; `checks`, as the checker only) may load this
 * library -- as the checker, never as the thing measured or shipped.  The product path (onika_b200/) has
 * no CPU fallback and never links or loads anything under oracle/.
 */
#ifndef ONIKA_B200_ORACLE_H
#define ONIKA_B200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int orc_num_threads(void);
/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline asks for the cores it is allowed to use */
void orc_set_num_threads(int n);

/* deterministic hashed inputs in [lo,hi), filled with the same static OpenMP partition the reductions use, so pages are
 * first-touched by the thread that will read them (what cpu_init does: tests/gpu_reduce_min_fp64.cu:30-44).
 * Same values as tests/_inputs.py:lcg_uniform. */
void orc_fill_lcg_f64(double* d, uint64_t n, uint64_t seed, double lo, double hi);

/* ---- reductions ---- */
/* tests/gpu_reduce_min_fp64.cu:46-55 : m = DBL_MAX ; omp parallel for schedule(static) reduction(min:m) ; if(d[i]<m) m=d[i] */
double orc_reduce_min_f64(const double* d, uint64_t n);
/* mirror for max (include/onika/cuda/cuda_reduction.h:67-85 compares with '<' from the other side; start -DBL_MAX) */
double orc_reduce_max_f64(const double* d, uint64_t n);
/* sum idiom: every element atomically added into one accumulator -> order unspecified.
 * The oracle fixes one order: `nthreads` contiguous static chunks summed left to right, partials added in
 * thread order (what schedule(static) reduction(+) does when the runtime combines in thread order). */
double orc_reduce_sum_f64(const double* d, uint64_t n, int nthreads);
/* long-double reference sum used to state the error of any summation order */
double orc_reduce_sum_f64_ld(const double* d, uint64_t n);

/* ---- element streaming through parallel_for / block_parallel_for ---- */
/* include/onika/parallel/parallel_for.h:53-112 with functor y[i] += a*x[i] (unfused mul, add) */
void orc_axpy_f64(double* y, const double* x, double a, uint64_t n);
/* include/onika/parallel/memset.h:30-51 */
void orc_memset_f64(double* d, uint64_t n, double v);
/* include/onika/extras/block_parallel_value_add_functor.h:36-44 : a[i][j] += v, row per block */
void orc_value_add_f64(double* a, uint64_t rows, uint64_t cols, double v);
/* include/onika/parallel/parallel_for.h:114-155 + tutorial functor parallel_for_3d_benchmark.cu:50-65 :
 * m[(k*N+j)*N+i] += 1 for every (i,j,k) in [start,end) */
void orc_parallel_for_3d_f64(double* m, int64_t n, const int64_t start[3], const int64_t end[3]);

/* ---- SOATL packed field arrays layout ---- */
/* include/onika/memory/alloc_strategy.h:44-67 (ChunkLogAllocationStrategy) */
uint64_t orc_soatl_update_capacity(uint64_t s, uint64_t capacity, uint64_t chunk);
/* include/onika/soatl/field_arrays.h:494-501 (pfa_advance_field_ptr) == packed_field_arrays.h:105-118
 * offsets[k] = sum_{j<k} roundup(capacity*field_size[j], A); returns total storage bytes */
uint64_t orc_soatl_layout(uint64_t align, const uint32_t* field_sizes, int nfields, uint64_t capacity, uint64_t* offsets);
/* the closed form used by the reference's size calculator, packed_field_arrays.h:45-58,83-95 (must agree) */
uint64_t orc_soatl_storage_size_calc(uint64_t align, uint64_t chunk, const uint32_t* field_sizes, int nfields, uint64_t capacity);

/* ---- SOATL streaming ---- */
/* tests/benchmark.cpp:88-110 through soatl/compute.h:214-232 : iterates ceil(n/chunk)*chunk elements */
void orc_soatl_dist_f64(double* e, const double* rx, const double* ry, const double* rz, uint64_t n_padded,
                        double ax, double ay, double az);
/* f_k[i] = f_k[i]*s + c_k for k<nfields over one packed allocation; touches padding like
 * parallel_apply_simd does (compute.h:221-231) */
void orc_soatl_rmw_f64(uint8_t* storage, uint64_t align, uint64_t chunk, int nfields, uint64_t n, uint64_t capacity,
                       double s, const double* c);
/* tests/soatlserializetest.cpp:86-104 : per-column memcpy (soatl/copy.h:45) from one layout into another */
void orc_soatl_repack(uint8_t* dst, uint64_t dst_align, uint64_t dst_capacity,
                      const uint8_t* src, uint64_t src_align, uint64_t src_capacity,
                      const uint32_t* field_sizes, int nfields, uint64_t n);

/* ---- cell grid (config C4): cells packed back to back in one arena ---- */
/* arena layout: cell c occupies [cell_off[c], cell_off[c] + layout(cap[c])) with the SOATL layout of
 * FieldArrays<align,chunk, fields...>; sums field `field` (fp64) over size[c] particles in index order,
 * cell_sum[c] written, *total = sum over cells in cell order */
void orc_cell_sum_f64(const uint8_t* arena, const uint64_t* cell_off, const uint32_t* cell_size, const uint32_t* cell_cap,
                      uint64_t ncells, uint64_t align, const uint32_t* field_sizes, int nfields, int field,
                      double* cell_sum, double* total);
/* the same sums over a field-major grid: values of all cells back to back, cell c = [cell_begin[c], cell_begin[c+1]).
 * Equals oracle/_ref's ref_cell_sum (the reference's block_parallel_for over per-cell FieldArrays) on the same particles. */
void orc_cell_sum_packed_f64(const double* values, const uint64_t* cell_begin, uint64_t ncells, double* cell_sum, double* total);
/* builds cell_cap / cell_off for given sizes (capacity = update_capacity(size,0,chunk); offsets aligned to
 * `cell_align`); returns arena bytes */
uint64_t orc_cell_grid_layout(const uint32_t* cell_size, uint64_t ncells, uint64_t align, uint64_t chunk,
                              const uint32_t* field_sizes, int nfields, uint64_t cell_align,
                              uint32_t* cell_cap, uint64_t* cell_off);

/* ---- lanes (config C5) ---- */
/* tests/parallel_queue_lane.cu:98-104 "hybrid-lane": lane0 = a1+=v11 ; a1+=v12   lane1 = a2+=v21 ; a2+=v22 */
void orc_lane_sequence_f64(double* a1, double* a2, uint64_t rows, uint64_t cols,
                           double v11, double v12, double v21, double v22);

/* ---- id-based redistribution between ranks ---- */
/* include/onika/mpi/xs_data_move_range_utils.h:39-63 : rendezvous rank of an id and the id range a rank is in charge of */
int      orc_xs_process_for_id(uint64_t id, int nprocs, uint64_t id_min, uint64_t id_max);
uint64_t orc_xs_range_start(int rank, int nprocs, uint64_t id_min, uint64_t id_max);
/* src/mpi/xs_data_move.cpp:42-462 communication_scheme_from_ids, for all `world` ranks at once (the MPI_Alltoall(v) calls
 * become loops over ranks).  Rank-concatenated arrays: before_off/after_off hold world+1 prefix sums; ids_* the ids;
 * send_indices/recv_indices the scatter tables (position of a local element in its rank's send / receive buffer);
 * send_count/send_displ/recv_count/recv_displ are world x world, row = rank.  Returns 0, or -1 when an id has no source or
 * no destination (the reference aborts: .cpp:286-289,331-334,366-369). */
int orc_xs_comm_scheme(int world, uint64_t id_min, uint64_t id_max,
                       const uint64_t* before_off, const uint64_t* ids_before, const uint64_t* after_off, const uint64_t* ids_after,
                       int32_t* send_indices, int32_t* recv_indices,
                       int32_t* send_count, int32_t* send_displ, int32_t* recv_count, int32_t* recv_displ);
/* include/onika/mpi/xs_data_move_impl.h:44-113 data_move for all ranks: pack by send_indices, exchange blocks
 * (MPI_Alltoallv), unpack by recv_indices.  elem_bytes-sized opaque elements. */
void orc_xs_data_move(int world, const uint64_t* before_off, const uint64_t* after_off,
                      const int32_t* send_indices, const int32_t* recv_indices,
                      const int32_t* send_count, const int32_t* send_displ, const int32_t* recv_count, const int32_t* recv_displ,
                      uint64_t elem_bytes, const uint8_t* in, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
