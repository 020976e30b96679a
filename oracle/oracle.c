/* oracle/oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference (Collab4exaNBody/onika) algorithms
 * on the data-parallel hot path.  See oracle.h for the contract and who may load this.
 *
 * Arithmetic rule: unfused IEEE double operations (build with -ffp-contract=off), which is what a stock
 * x86-64 build of the reference computes (no -march in its CMake => no FMA to contract onto).
 * Pinned against oracle/_ref (the real reference) by tests/test_oracle_vs_reference.py and against
 * the committed fixtures under tests/golden/ by tests/test_oracle_golden.py.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <omp.h>

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

void orc_fill_lcg_f64(double* d, uint64_t n, uint64_t seed, double lo, double hi)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; i++)
  {
    uint64_t h = ((uint64_t)i * 2654435761ull + seed * 40503ull + 12345ull) & 0xFFFFFFFFull;
    h = ((h ^ (h >> 15)) * 2246822519ull) & 0xFFFFFFFFull;
    h = (h ^ (h >> 13)) & 0xFFFFFFFFull;
    d[i] = lo + (hi - lo) * ((double)h / 4294967296.0);
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* reductions                                                                                        */
/* ------------------------------------------------------------------------------------------------ */

/* reference: tests/gpu_reduce_min_fp64.cu:46-55
 *   double data_min = std::numeric_limits<double>::max();
 *   #pragma omp parallel for schedule(static) reduction(min:data_min)
 *   for(i) if( data[i] < data_min ) data_min = data[i];
 * NaNs never win (comparison false); ties keep the earlier value. */
double orc_reduce_min_f64(const double* d, uint64_t n)
{
  double m = DBL_MAX;
#pragma omp parallel for schedule(static) reduction(min:m)
  for (int64_t i = 0; i < (int64_t)n; i++)
  {
    if (d[i] < m) m = d[i];
  }
  return m;
}

/* same loop with the comparison mirrored, starting from the lowest finite double
 * (block_reduce_max, include/onika/cuda/cuda_reduction.h:67-85 uses `sdata[a] < sdata[b]`). */
double orc_reduce_max_f64(const double* d, uint64_t n)
{
  double m = -DBL_MAX;
#pragma omp parallel for schedule(static) reduction(max:m)
  for (int64_t i = 0; i < (int64_t)n; i++)
  {
    if (d[i] > m) m = d[i];
  }
  return m;
}

/* sum idiom: functor does ONIKA_CU_ATOMIC_ADD(*acc, v) per element (include/onika/cuda/cuda.h:200-207 host side =
 * omp atomic capture, include/onika/atomic.h:14-62).  Order unspecified in the reference; fixed here as
 * libgomp's schedule(static) partition (first n%T threads get one extra iteration), partials combined in
 * thread order. */
double orc_reduce_sum_f64(const double* d, uint64_t n, int nthreads)
{
  if (nthreads < 1) nthreads = 1;
  double* part = (double*)calloc((size_t)nthreads, sizeof(double));
  const uint64_t q = n / (uint64_t)nthreads, r = n % (uint64_t)nthreads;
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < nthreads; t++)
  {
    const uint64_t b = (uint64_t)t * q + ((uint64_t)t < r ? (uint64_t)t : r);
    const uint64_t e = b + q + ((uint64_t)t < r ? 1u : 0u);
    double s = 0.0;
    for (uint64_t i = b; i < e; i++) s += d[i];
    part[t] = s;
  }
  double s = 0.0;
  for (int t = 0; t < nthreads; t++) s += part[t];
  free(part);
  return s;
}

double orc_reduce_sum_f64_ld(const double* d, uint64_t n)
{
  /* blocked long double accumulation: error << 1e-16 relative for the sizes used */
  long double s = 0.0L;
  const uint64_t B = 1u << 16;
  for (uint64_t b = 0; b < n; b += B)
  {
    long double p = 0.0L;
    const uint64_t e = (b + B < n) ? b + B : n;
    for (uint64_t i = b; i < e; i++) p += (long double)d[i];
    s += p;
  }
  return (double)s;
}

/* ------------------------------------------------------------------------------------------------ */
/* element streaming                                                                                 */
/* ------------------------------------------------------------------------------------------------ */

/* reference dispatch: parallel_for(N, f, pec) -> ParallelForBlockAdapter (include/onika/parallel/parallel_for.h:53-84,
 * host: BLOCK_SIZE=1, THREAD_IDX=0 => f(offset+i) for i in [0,N)) -> block_parallel_for 1D OpenMP loop, default
 * schedule(guided) (block_parallel_for_adapter.h:289-296).  Functor: y[i] += a*x[i]. */
void orc_axpy_f64(double* y, const double* x, double a, uint64_t n)
{
#pragma omp parallel for schedule(guided)
  for (int64_t i = 0; i < (int64_t)n; i++)
  {
    const double p = a * x[i];
    y[i] = y[i] + p;
  }
}

/* reference: MemSetFunctor, include/onika/parallel/memset.h:30-40 : m_data[i] = m_value */
void orc_memset_f64(double* d, uint64_t n, double v)
{
#pragma omp parallel for schedule(guided)
  for (int64_t i = 0; i < (int64_t)n; i++) d[i] = v;
}

/* reference: BlockParallelValueAddFunctor::operator()(size_t i),
 * include/onika/extras/block_parallel_value_add_functor.h:36-44 : for j in [0,cols) a[i][j] += v ;
 * one call per row i from block_parallel_for(rows, ...) (block_parallel_for_adapter.h:279-300) */
void orc_value_add_f64(double* a, uint64_t rows, uint64_t cols, double v)
{
#pragma omp parallel for schedule(guided)
  for (int64_t i = 0; i < (int64_t)rows; i++)
  {
    double* row = a + (uint64_t)i * cols;
    for (uint64_t j = 0; j < cols; j++) row[j] += v;
  }
}

/* reference: parallel_for(ParallelExecutionSpace<3>) normalises start to 0 and carries the offset in the adapter
 * (include/onika/parallel/parallel_for.h:114-155); host loop order k,j,i (block_parallel_for_adapter.h:341-359);
 * functor Grid3DBenchmarkFunctor (plugins/onikaParallelTutorial/parallel_for_3d_benchmark.cu:50-65). */
void orc_parallel_for_3d_f64(double* m, int64_t n, const int64_t start[3], const int64_t end[3])
{
  for (int64_t k = start[2]; k < end[2]; k++)
    for (int64_t j = start[1]; j < end[1]; j++)
      for (int64_t i = start[0]; i < end[0]; i++)
        m[(k * n + j) * n + i] += 1.0;
}

/* ------------------------------------------------------------------------------------------------ */
/* SOATL layout                                                                                      */
/* ------------------------------------------------------------------------------------------------ */

/* reference: ChunkLogAllocationStrategy::update_capacity, include/onika/memory/alloc_strategy.h:44-67 */
uint64_t orc_soatl_update_capacity(uint64_t s, uint64_t capacity, uint64_t chunk)
{
  uint64_t nc;
  if (s > capacity)
  {
    if (capacity <= s / 2 || capacity == 0) nc = s;
    else nc = capacity * 2;
  }
  else if (s <= capacity / 2 || s == 0) nc = s;
  else nc = capacity;
  return ((nc + chunk - 1) / chunk) * chunk;
}

/* reference: FieldArraysWithAllocator::pfa_advance_field_ptr / set_storage_ptr,
 * include/onika/soatl/field_arrays.h:494-512 ; check_size_for_capacity, packed_field_arrays.h:105-118 */
uint64_t orc_soatl_layout(uint64_t align, const uint32_t* field_sizes, int nfields, uint64_t capacity, uint64_t* offsets)
{
  uint64_t off = 0;
  for (int k = 0; k < nfields; k++)
  {
    if (offsets) offsets[k] = off;
    const uint64_t bytes = capacity * (uint64_t)field_sizes[k];
    off += (bytes + align - 1) & ~(align - 1);
  }
  return off;
}

/* reference: PFACalc + SizeCounts::storage_size, include/onika/soatl/packed_field_arrays.h:36-58,83-95,120-144.
 * Fields are decomposed in units of CA=A/C bytes: a field of size sz counts sz/CA units of size CA plus one unit
 * of size sz%CA (if non zero); storage = C * sum_units N * roundup( (capacity/C)*SIZE , CA ). */
uint64_t orc_soatl_storage_size_calc(uint64_t align, uint64_t chunk, const uint32_t* field_sizes, int nfields, uint64_t capacity)
{
  const uint64_t CA = align / chunk;
  const uint64_t himask = ~(CA - 1);
  uint64_t* count = (uint64_t*)calloc((size_t)CA + 1, sizeof(uint64_t));
  for (int k = 0; k < nfields; k++)
  {
    count[field_sizes[k] % CA] += 1;
    count[CA] += field_sizes[k] / CA;
  }
  count[0] = 0; /* remainder 0 contributes nothing; the reference overwrites slot 0 with the CA-sized units */
  const uint64_t cap = capacity / chunk;
  uint64_t units = cap * CA * count[CA]; /* (cap*CA + CA-1) & himask == cap*CA */
  for (uint64_t r = 1; r < CA; r++) units += count[r] * ((cap * r + CA - 1) & himask);
  free(count);
  return chunk * units;
}

/* ------------------------------------------------------------------------------------------------ */
/* SOATL streaming                                                                                   */
/* ------------------------------------------------------------------------------------------------ */

/* reference kernel: COMPUTE_KERNEL, tests/benchmark.cpp:96-104, applied by
 * soatl::parallel_apply_simd(f, arrays, dist, rx, ry, rz) (include/onika/soatl/compute.h:214-232,268-278):
 * omp parallel for over ceil(N/C) chunks, omp simd over C lanes => caller passes n_padded = ceil(N/C)*C. */
void orc_soatl_dist_f64(double* e, const double* rx, const double* ry, const double* rz, uint64_t n_padded,
                        double ax, double ay, double az)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n_padded; i++)
  {
    double x = rx[i], y = ry[i], z = rz[i], d;
    x = x - ax;
    y = y - ay;
    z = z - az;
    d = sqrt(x * x + y * y + z * z);
    x /= d;
    y /= d;
    z /= d;
    d += x / (y * z);
    e[i] = d;
  }
}

/* config C3: f_k[i] = f_k[i]*s + c_k over FieldArrays<A,C,f0..f{K-1}> (all fp64), through
 * parallel_apply_simd => ceil(n/C)*C elements per column are touched (compute.h:221-231). */
void orc_soatl_rmw_f64(uint8_t* storage, uint64_t align, uint64_t chunk, int nfields, uint64_t n, uint64_t capacity,
                       double s, const double* c)
{
  const uint64_t n_padded = ((n + chunk - 1) / chunk) * chunk;
  const uint64_t col_bytes = (capacity * 8u + align - 1) & ~(align - 1);
  for (int k = 0; k < nfields; k++)
  {
    double* f = (double*)(storage + (uint64_t)k * col_bytes);
    const double ck = c[k];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n_padded; i++)
    {
      const double p = f[i] * s;
      f[i] = p + ck;
    }
  }
}

/* reference: soatl::copy(dst, src) = one memcpy of n*sizeof(T) per field (include/onika/soatl/copy.h:29-97),
 * as exercised by tests/soatlserializetest.cpp:86-104 between <64,8,...> and the <1,1,...> wire layout. */
void orc_soatl_repack(uint8_t* dst, uint64_t dst_align, uint64_t dst_capacity,
                      const uint8_t* src, uint64_t src_align, uint64_t src_capacity,
                      const uint32_t* field_sizes, int nfields, uint64_t n)
{
  uint64_t doff[64], soff[64];
  if (nfields > 64) return;
  orc_soatl_layout(dst_align, field_sizes, nfields, dst_capacity, doff);
  orc_soatl_layout(src_align, field_sizes, nfields, src_capacity, soff);
  for (int k = 0; k < nfields; k++) memcpy(dst + doff[k], src + soff[k], n * (uint64_t)field_sizes[k]);
}

/* ------------------------------------------------------------------------------------------------ */
/* cell grid                                                                                         */
/* ------------------------------------------------------------------------------------------------ */

uint64_t orc_cell_grid_layout(const uint32_t* cell_size, uint64_t ncells, uint64_t align, uint64_t chunk,
                              const uint32_t* field_sizes, int nfields, uint64_t cell_align,
                              uint32_t* cell_cap, uint64_t* cell_off)
{
  uint64_t off = 0;
  for (uint64_t c = 0; c < ncells; c++)
  {
    /* a freshly resized FieldArrays: capacity = update_capacity(size, 0, chunk) (field_arrays.h:200-211) */
    const uint64_t cap = orc_soatl_update_capacity(cell_size[c], 0, chunk);
    cell_cap[c] = (uint32_t)cap;
    cell_off[c] = off;
    uint64_t bytes = orc_soatl_layout(align, field_sizes, nfields, cap, NULL);
    off += (bytes + cell_align - 1) & ~(cell_align - 1);
  }
  return off;
}

/* config C4: block_parallel_for(ncells, functor) where the functor sums field `field` of its cell in particle
 * order (host: BLOCK_SIZE=1 => sequential, include/onika/cuda/cuda.h:180-229) and adds the cell sum into a global
 * accumulator.  The global accumulation order is unspecified in the reference (atomic); fixed here as cell order. */
void orc_cell_sum_f64(const uint8_t* arena, const uint64_t* cell_off, const uint32_t* cell_size, const uint32_t* cell_cap,
                      uint64_t ncells, uint64_t align, const uint32_t* field_sizes, int nfields, int field,
                      double* cell_sum, double* total)
{
  (void)nfields;
#pragma omp parallel for schedule(guided)
  for (int64_t c = 0; c < (int64_t)ncells; c++)
  {
    uint64_t foff = 0;
    for (int k = 0; k < field; k++)
    {
      const uint64_t bytes = (uint64_t)cell_cap[c] * field_sizes[k];
      foff += (bytes + align - 1) & ~(align - 1);
    }
    const double* e = (const double*)(arena + cell_off[c] + foff);
    double s = 0.0;
    for (uint32_t p = 0; p < cell_size[c]; p++) s += e[p];
    cell_sum[c] = s;
  }
  double t = 0.0;
  for (uint64_t c = 0; c < ncells; c++) t += cell_sum[c];
  *total = t;
}

/* ------------------------------------------------------------------------------------------------ */
/* lanes                                                                                             */
/* ------------------------------------------------------------------------------------------------ */

/* reference: tests/parallel_queue_lane.cu:98-104 ("hybrid-lane"): ops in one lane run in order, lanes are
 * independent (src/parallel/parallel_execution_queue.cpp:242-271).  With value-add kernels the result is the same
 * as running lane 0 then lane 1. */
void orc_lane_sequence_f64(double* a1, double* a2, uint64_t rows, uint64_t cols,
                           double v11, double v12, double v21, double v22)
{
  orc_value_add_f64(a1, rows, cols, v11);
  orc_value_add_f64(a2, rows, cols, v21);
  orc_value_add_f64(a1, rows, cols, v12);
  orc_value_add_f64(a2, rows, cols, v22);
}

/* ============================ id-based redistribution (src/mpi/xs_data_move.cpp) ============================ */

/* xs_data_move_range_utils.h:39-45 */
int orc_xs_process_for_id(uint64_t id, int nprocs, uint64_t id_min, uint64_t id_max)
{
  return (int)( ( ( id - id_min + 1 ) * (uint64_t)nprocs - 1 ) / ( id_max - id_min ) );
}

/* xs_data_move_range_utils.h:48-55 */
uint64_t orc_xs_range_start(int rank, int nprocs, uint64_t id_min, uint64_t id_max)
{
  const uint64_t range = id_max - id_min;
  return id_min + ( (uint64_t)rank * range ) / (uint64_t)nprocs;
}

/* What the three exchanges of communication_scheme_from_ids compute, rank by rank:
 *  (1) .cpp:76-247  every rank tells the rendezvous rank of each id where the id is before (owner, index) and where it must
 *      be after; the rendezvous ranks hold id_move[id - range_start] = {src_owner, src_index, dst_owner, dst_index};
 *  (2) .cpp:279-355 each rendezvous rank walks ITS id range in increasing id and hands every entry to the id's source
 *      owner; a source rank reads the lists of rendezvous ranks 0,1,... one after the other, i.e. its own elements in
 *      increasing id (ranges increase with the rank);
 *  (3) .cpp:357-416 the source rank gives the k-th element it meets for destination d the send-buffer slot
 *      send_displ[d]+k, and sends the matching dst_index list to d, which lays the lists of sources 0,1,... end to end;
 *  (4) .cpp:418-433 both tables are inverted into scatter form: table[local element] = buffer slot.
 * The union of the rendezvous tables is one directory over [id_min,id_max); one walk in increasing id serves (2)-(3). */
int orc_xs_comm_scheme(int world, uint64_t id_min, uint64_t id_max,
                       const uint64_t* before_off, const uint64_t* ids_before, const uint64_t* after_off, const uint64_t* ids_after,
                       int32_t* send_indices, int32_t* recv_indices,
                       int32_t* send_count, int32_t* send_displ, int32_t* recv_count, int32_t* recv_displ)
{
  const uint64_t range = id_max - id_min;
  const size_t W = (size_t)world;
  int32_t* src_owner = (int32_t*)malloc( sizeof(int32_t) * ( range ? range : 1 ) );
  int32_t* dst_owner = (int32_t*)malloc( sizeof(int32_t) * ( range ? range : 1 ) );
  uint64_t* src_index = (uint64_t*)malloc( sizeof(uint64_t) * ( range ? range : 1 ) );
  uint64_t* dst_index = (uint64_t*)malloc( sizeof(uint64_t) * ( range ? range : 1 ) );
  int32_t* cursor = (int32_t*)calloc( W * W , sizeof(int32_t) );
  int rc = 0;
  for(uint64_t k=0;k<range;k++) { src_owner[k] = -1; dst_owner[k] = -1; }      /* IdMove{0,0,-1,-1}, .cpp:218 */
  for(int r=0;r<world;r++)
  {
    for(uint64_t i=before_off[r];i<before_off[r+1];i++) { const uint64_t k = ids_before[i] - id_min; src_owner[k] = r; src_index[k] = i - before_off[r]; }
    for(uint64_t i=after_off[r];i<after_off[r+1];i++)   { const uint64_t k = ids_after[i] - id_min;  dst_owner[k] = r; dst_index[k] = i - after_off[r]; }
  }
  /* counts: send_count[s][d] = elements s sends to d ; recv_count[d][s] is the same number seen from d (.cpp:372,391) */
  for(size_t k=0;k<W*W;k++) { send_count[k] = 0; recv_count[k] = 0; }
  for(uint64_t k=0;k<range;k++)
  {
    if( src_owner[k] < 0 || dst_owner[k] < 0 ) { rc = -1; goto done; }
    ++ send_count[ (size_t)src_owner[k]*W + (size_t)dst_owner[k] ];
    ++ recv_count[ (size_t)dst_owner[k]*W + (size_t)src_owner[k] ];
  }
  for(int r=0;r<world;r++)
  {
    send_displ[(size_t)r*W] = 0; recv_displ[(size_t)r*W] = 0;                   /* .cpp:385-389, 423-427 */
    for(int p=1;p<world;p++)
    {
      send_displ[(size_t)r*W+p] = send_displ[(size_t)r*W+p-1] + send_count[(size_t)r*W+p-1];
      recv_displ[(size_t)r*W+p] = recv_displ[(size_t)r*W+p-1] + recv_count[(size_t)r*W+p-1];
    }
  }
  /* slots, in increasing id: the k-th element of the pair (s,d) sits at send_displ[s][d]+k on s and recv_displ[d][s]+k on d */
  for(uint64_t k=0;k<range;k++)
  {
    const size_t s = (size_t)src_owner[k], d = (size_t)dst_owner[k];
    const int32_t nth = cursor[s*W+d] ++;
    send_indices[ before_off[s] + src_index[k] ] = send_displ[s*W+d] + nth;
    recv_indices[ after_off[d] + dst_index[k] ]  = recv_displ[d*W+s] + nth;
  }
done:
  free(src_owner); free(dst_owner); free(src_index); free(dst_index); free(cursor);
  return rc;
}

/* xs_data_move_impl.h:74-111 */
void orc_xs_data_move(int world, const uint64_t* before_off, const uint64_t* after_off,
                      const int32_t* send_indices, const int32_t* recv_indices,
                      const int32_t* send_count, const int32_t* send_displ, const int32_t* recv_count, const int32_t* recv_displ,
                      uint64_t elem_bytes, const uint8_t* in, uint8_t* out)
{
  const size_t W = (size_t)world;
  const uint64_t total_before = before_off[world], total_after = after_off[world];
  uint8_t* comm_input  = (uint8_t*)malloc( ( total_before ? total_before : 1 ) * elem_bytes );   /* all ranks' send buffers, rank after rank */
  uint8_t* comm_output = (uint8_t*)malloc( ( total_after ? total_after : 1 ) * elem_bytes );
  for(int r=0;r<world;r++)                                                                      /* pack, :74-82 */
    for(uint64_t i=before_off[r];i<before_off[r+1];i++)
      memcpy( comm_input + ( before_off[r] + (uint64_t)send_indices[i] ) * elem_bytes , in + i*elem_bytes , elem_bytes );
  for(int d=0;d<world;d++)                                                                      /* MPI_Alltoallv, :88-94 */
    for(int s=0;s<world;s++)
      memcpy( comm_output + ( after_off[d] + (uint64_t)recv_displ[(size_t)d*W+s] ) * elem_bytes ,
              comm_input  + ( before_off[s] + (uint64_t)send_displ[(size_t)s*W+d] ) * elem_bytes ,
              (uint64_t)recv_count[(size_t)d*W+s] * elem_bytes );
  (void)send_count;
  for(int r=0;r<world;r++)                                                                      /* unpack, :96-103 */
    for(uint64_t j=after_off[r];j<after_off[r+1];j++)
      memcpy( out + j*elem_bytes , comm_output + ( after_off[r] + (uint64_t)recv_indices[j] ) * elem_bytes , elem_bytes );
  free(comm_input); free(comm_output);
}

/* field-major cell grid: per-cell fold in particle order, total in cell order (same functor as orc_cell_sum_f64) */
void orc_cell_sum_packed_f64(const double* values, const uint64_t* cell_begin, uint64_t ncells, double* cell_sum, double* total)
{
  double t = 0.0;
  for(uint64_t c=0;c<ncells;c++)
  {
    double s = 0.0;
    for(uint64_t p=cell_begin[c];p<cell_begin[c+1];p++) s += values[p];
    cell_sum[c] = s;
    t += s;
  }
  if( total != NULL ) *total = t;
}
