// onk_cells.cu -- cell-grid kernels for sm_100a (config C4): block_parallel_for over cells whose particles live in
// variable-length SOATL blocks.
//
// Reference behaviour: block_parallel_for(ncells, functor) launches ONE 128-thread block per cell
// (include/onika/parallel/block_parallel_for_adapter.h:81-93,195-198), the functor strides the cell's ~16 particles
// with ONIKA_CU_BLOCK_SIMD_FOR (include/onika/cuda/cuda.h:102) => 12.5 % lane use, 2 M block launches, one global
// atomic per cell for the grid-wide sum.  Each cell is its own cudaMallocManaged FieldArrays allocation.
//
// B200 design:
//   * HBM layout: all cells packed back to back in ONE arena (each block keeps the exact SOATL column layout
//     FieldArrays<A,C,...> would give it), with three small tables: byte offset, size, capacity per cell;
//   * a warp takes 32 consecutive cells per step: one coalesced load per table, then the variable-length columns
//     are STAGED THROUGH SHARED MEMORY 16 particles at a time: a half-warp fetches one cell's 16 x 8 B = one
//     128-byte line per load instruction (fully coalesced, CELL_ILP cells in flight per half-warp) into a padded
//     32 x 17 tile; then lane l folds row l (its own cell) in particle order with conflict-free LDS.64;
//   * in-cell sums are therefore sequential in particle order, bit-identical to the reference's host loop, and the
//     per-cell results leave as one coalesced 256-byte store per warp;
//   * the grid-wide total uses per-thread partials, a block reduction and the same ticketed two-pass finish as
//     onk_reduce.cu (fixed order, no atomics on data);
//   * persistent grid: CELL_CTAS_PER_SM x 148 CTAs, warps walk groups of 32 cells grid-stride.
#include "onk_internal.cuh"

namespace onk
{
  constexpr unsigned CELL_CTAS_PER_SM = 6;
  constexpr int CELL_MAX_FIELDS = 16;
  constexpr int CELL_ILP = 4;                       // cells in flight per half-warp
  constexpr int CELL_ROW = 17;                      // padded row length (doubles) of the staging tile
  struct CellFieldParams { uint32_t pre_bytes[CELL_MAX_FIELDS]; int npre; uint32_t align_mask; };

  __device__ __forceinline__ uint64_t cell_field_offset(uint32_t cap, const CellFieldParams& fp)
  {
    uint64_t off = 0;
    for (int k = 0; k < fp.npre; k++) off += ((uint64_t)cap * fp.pre_bytes[k] + fp.align_mask) & ~(uint64_t)fp.align_mask;
    return off;
  }

  __global__ void __launch_bounds__(THREADS, CELL_CTAS_PER_SM)
  cell_sum_f64_kernel(const uint8_t* __restrict__ arena, const uint64_t* __restrict__ cell_off, const uint32_t* __restrict__ cell_size,
                      const uint32_t* __restrict__ cell_cap, uint64_t ncells, const __grid_constant__ CellFieldParams fp,
                      double* __restrict__ cell_sum, double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ total)
  {
    __shared__ double stage[THREADS / 32][32 * CELL_ROW];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned hl = lane & 15u;            // lane within the half-warp
    const unsigned half = lane >> 4;           // which half-warp
    double* tile = stage[threadIdx.x >> 5];
    const uint64_t warp_global = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * THREADS) >> 5;
    double thread_total = 0.0;

    // a warp takes 32 consecutive cells per step: lane l owns cell c0+l (tables, in-order fold, result)
    for (uint64_t c0 = warp_global * 32; c0 < ncells; c0 += nwarps * 32)
    {
      const uint64_t cmine = c0 + lane;
      uint64_t off = 0; uint32_t sz = 0, cap = 0;
      if (cmine < ncells) { off = __ldg(cell_off + cmine); sz = __ldg(cell_size + cmine); cap = __ldg(cell_cap + cmine); }
      const uint64_t addr = off + cell_field_offset(cap, fp);
      uint32_t nmax = sz;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));

      double s = 0.0;
      for (uint32_t p0 = 0; p0 < nmax; p0 += 16)
      {
        // stage particles [p0, p0+16) of all 32 cells: half-warp `half` fetches cells half*16 + j
#pragma unroll
        for (int j0 = 0; j0 < 16; j0 += CELL_ILP)
        {
          double v[CELL_ILP];
#pragma unroll
          for (int q = 0; q < CELL_ILP; q++)
          {
            const int src = half * 16 + j0 + q;
            const uint64_t a = __shfl_sync(0xffffffffu, addr, src);
            const uint32_t n = __shfl_sync(0xffffffffu, sz, src);
            v[q] = (p0 + hl < n) ? ld_stream_ro1((const double*)(arena + a) + p0 + hl) : 0.0;
          }
#pragma unroll
          for (int q = 0; q < CELL_ILP; q++) tile[(half * 16 + j0 + q) * CELL_ROW + hl] = v[q];
        }
        __syncwarp();
        // particle order: s = ((s + e[p0]) + e[p0+1]) + ...   (+0.0 beyond the cell's size leaves s unchanged)
#pragma unroll
        for (int i = 0; i < 16; i++) s = __dadd_rn(s, tile[lane * CELL_ROW + i]);
        __syncwarp();
      }
      if (cmine < ncells) { cell_sum[cmine] = s; thread_total = __dadd_rn(thread_total, s); }
    }

    if (total == nullptr) return;
    const double blk = block_reduce<ONK_OP_SUM, THREADS>(thread_total);
    __shared__ bool is_last;
    if (threadIdx.x == 0)
    {
      partials[blockIdx.x] = blk;
      __threadfence();
      is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last)
    {
      __threadfence();
      double r = 0.0;
      for (unsigned i = threadIdx.x; i < gridDim.x; i += THREADS) r = __dadd_rn(r, __ldcg(partials + i));
      r = block_reduce<ONK_OP_SUM, THREADS>(r);
      if (threadIdx.x == 0) { *total = r; *ticket = 0u; }
    }
  }
}

using namespace onk;

extern "C" {

uint64_t onk_cell_grid_layout(const uint32_t* cell_size, uint64_t ncells, uint64_t align, uint64_t chunk,
                              const uint32_t* field_bytes, int nfields, uint64_t cell_align,
                              uint32_t* cell_cap, uint64_t* cell_off)
{
  // every cell is a freshly sized FieldArrays<align,chunk,...>: capacity = whole chunks covering its size
  // (field_arrays.h:200-211 with previous capacity 0); blocks are laid out in cell order
  if (cell_align == 0) cell_align = 1;
  uint64_t cursor = 0;
  for (uint64_t c = 0; c < ncells; c++)
  {
    const uint64_t cap = onk_soatl_update_capacity(cell_size[c], 0, chunk);
    if (cell_cap) cell_cap[c] = (uint32_t)cap;
    if (cell_off) cell_off[c] = cursor;
    const uint64_t bytes = onk_soatl_layout(align, field_bytes, nfields, cap, nullptr);
    cursor += ((bytes + cell_align - 1) / cell_align) * cell_align;
  }
  return cursor;
}

int onk_cell_sum_f64(const void* arena_dev, const uint64_t* cell_off_dev, const uint32_t* cell_size_dev,
                     const uint32_t* cell_cap_dev, uint64_t ncells, uint64_t align,
                     const uint32_t* field_bytes, int nfields, int field,
                     double* cell_sum_dev, double* total_dev, int lane)
{
  ONK_REQUIRE(field >= 0 && field < nfields && field <= CELL_MAX_FIELDS, "onk_cell_sum_f64: bad field index %d", field);
  ONK_REQUIRE(field_bytes != nullptr && field_bytes[field] == 8, "onk_cell_sum_f64: the summed field must be fp64");
  ONK_REQUIRE(align >= 8 && (align & (align - 1)) == 0, "onk_cell_sum_f64: alignment must be a power of two >= 8 (fp64 columns)");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  if (ncells == 0)
  {
    if (total_dev) ONK_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(double), L->stream));
    return ONK_OK;
  }
  ONK_REQUIRE(arena_dev && cell_off_dev && cell_size_dev && cell_cap_dev && cell_sum_dev, "onk_cell_sum_f64: null pointer");
  CellFieldParams fp{};
  fp.npre = field; fp.align_mask = (uint32_t)(align - 1);
  for (int k = 0; k < field; k++) fp.pre_bytes[k] = field_bytes[k];
  const unsigned cells_per_cta = (THREADS / 32) * 32;
  unsigned grid = grid_for(ncells, cells_per_cta, CELL_CTAS_PER_SM);
  if (grid > (unsigned)MAX_PARTIALS) grid = MAX_PARTIALS;
  cell_sum_f64_kernel<<<grid, THREADS, 0, L->stream>>>((const uint8_t*)arena_dev, cell_off_dev, cell_size_dev, cell_cap_dev, ncells, fp,
                                                       cell_sum_dev, L->ws.partials, L->ws.ticket, total_dev);
  ONK_CHECK_LAUNCH(); count_launch();
  return ONK_OK;
}

} // extern "C"
