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
//     are STAGED THROUGH SHARED MEMORY 16 particles at a time: FOUR lanes fetch one cell's 16 x 8 B = 128-byte line
//     with one 256-bit load each, so a single LDG.E.256 instruction brings in the lines of 8 cells (1 KiB) and the 4
//     instructions that cover the warp's 32 cells are all in flight together; data lands in a padded 32 x 18 tile with
//     conflict-free 128-bit shared stores; then lane l folds row l (its own cell) in particle order with 128-bit
//     shared loads.  ~190 warp instructions per 32 cells instead of ~610 for a half-warp-per-cell 8-byte version
//     (profiles/r1_stream_cell_kernels_full.csv), which was issue/latency bound at 52 % of DRAM peak;
//   * cells whose column is not 32-byte aligned or whose capacity is not a multiple of 4 take a scalar load path;
//   * in-cell sums are therefore sequential in particle order, bit-identical to the reference's host loop, and the
//     per-cell results leave as one coalesced 256-byte store per warp;
//   * the grid-wide total uses per-thread partials, a block reduction and the same ticketed two-pass finish as
//     onk_reduce.cu (fixed order, no atomics on data);
//   * persistent grid: CELL_CTAS_PER_SM x 148 CTAs, warps walk groups of 32 cells grid-stride.
#include "onk_internal.cuh"
#include <cuda.h>
#include <cstdlib>

namespace onk
{
  constexpr unsigned CELL_CTAS_PER_SM = 4;
  constexpr int CELL_MAX_FIELDS = 16;
  constexpr int CELL_ROW = 18;                      // padded row length (doubles) of the staging tile: 144 B, 16-byte aligned
  struct CellFieldParams { uint32_t pre_bytes[CELL_MAX_FIELDS]; int npre; uint32_t align_mask; };

  __device__ __forceinline__ uint64_t cell_field_offset(uint32_t cap, const CellFieldParams& fp)
  {
    uint64_t off = 0;
    for (int k = 0; k < fp.npre; k++) off += ((uint64_t)cap * fp.pre_bytes[k] + fp.align_mask) & ~(uint64_t)fp.align_mask;
    return off;
  }

  __device__ __forceinline__ void sts128(double* p, double a, double b)
  {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" :: "r"((unsigned)__cvta_generic_to_shared(p)), "d"(a), "d"(b) : "memory");
  }
  __device__ __forceinline__ void lds128(const double* p, double& a, double& b)
  {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  }


  // grid-wide total: per-thread partials -> block reduction -> ticketed two-pass finish (fixed order)
  template<unsigned NT = THREADS>
  __device__ __forceinline__ void cell_total_finish(double thread_total, double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ total)
  {
    if (total == nullptr) return;
    const double blk = block_reduce<ONK_OP_SUM, NT>(thread_total);
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
      for (unsigned i = threadIdx.x; i < gridDim.x; i += NT) r = __dadd_rn(r, __ldcg(partials + i));
      r = block_reduce<ONK_OP_SUM, NT>(r);
      if (threadIdx.x == 0) { *total = r; *ticket = 0u; }
    }
  }

  constexpr uint32_t CELL_SCALAR_FLAG = 0x80000000u;

  __global__ void __launch_bounds__(THREADS, CELL_CTAS_PER_SM)
  cell_sum_f64_kernel(const uint8_t* __restrict__ arena, const uint64_t* __restrict__ cell_off, const uint32_t* __restrict__ cell_size,
                      const uint32_t* __restrict__ cell_cap, uint64_t ncells, const __grid_constant__ CellFieldParams fp,
                      double* __restrict__ cell_sum, double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ total)
  {
    __shared__ __align__(16) double stage[THREADS / 32][32 * CELL_ROW];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned q = lane & 3u;              // which 32-byte quarter of a cell's 128-byte line this lane fetches
    const unsigned c8 = lane >> 2;             // cell within the group of 8 served by one load instruction
    double* tile = stage[threadIdx.x >> 5];
    const uint64_t warp_global = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * THREADS) >> 5;
    double thread_total = 0.0;

    // Software pipeline over the warp's steps (32 consecutive cells each; lane l owns cell c0+l):
    //   tables of step s+2 are loaded into registers, the data lines of step s+1 are prefetched into L2
    //   (prefetch.global.L2, one or two fire-and-forget requests per lane), step s is processed meanwhile.
    //   (Measured: a two-step prefetch distance is slower, 110 us vs 92 us on 128^3 cells -- with 4736 resident warps it
    //   keeps ~90 MB of speculative lines in the 126 MB L2 and thrashes it.)
    // This removes the two serialized DRAM round trips (tables, then data) from every step.
    const uint64_t stride = nwarps * 32;
    auto load_tables = [&](uint64_t c, uint64_t& off, uint32_t& sz, uint32_t& cap)
    {
      off = 0; sz = 0; cap = 0;
      if (c < ncells) { off = __ldg(cell_off + c); sz = __ldg(cell_size + c); cap = __ldg(cell_cap + c); }
    };
    auto prefetch_lines = [&](uint64_t off, uint32_t sz, uint32_t cap)
    {
      const uint8_t* pcol = arena + off + cell_field_offset(cap, fp);
      if (sz > 0)  asm volatile("prefetch.global.L2 [%0];" :: "l"(pcol));
      if (sz > 16) asm volatile("prefetch.global.L2 [%0];" :: "l"(pcol + 128));
    };
    uint64_t off0, off1, off2; uint32_t sz0, sz1, sz2, cap0, cap1, cap2;
    const uint64_t cfirst = warp_global * 32 + lane;
    load_tables(cfirst, off0, sz0, cap0);
    load_tables(cfirst + stride, off1, sz1, cap1);
    for (uint64_t c0 = warp_global * 32; c0 < ncells; c0 += stride)
    {
      const uint64_t cmine = c0 + lane;
      load_tables(cmine + 2 * stride, off2, sz2, cap2);
      prefetch_lines(off1, sz1, cap1);
      const uint64_t off = off0; const uint32_t sz = sz0, cap = cap0;
      const uint8_t* col = arena + off + cell_field_offset(cap, fp);
      // size with a flag in the top bit: this cell's column cannot be read with aligned 256-bit loads
      const uint32_t szf = sz | ((((uintptr_t)col & 31u) != 0 || (cap & 3u) != 0) ? CELL_SCALAR_FLAG : 0u);
      uint32_t nmax = sz;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));

      double s = 0.0;
      for (uint32_t p0 = 0; p0 < nmax; p0 += 16)
      {
        // stage particles [p0, p0+16) of all 32 cells: load instruction `it` serves cells it*8 .. it*8+7
        d4 v[4];
#pragma unroll
        for (int it = 0; it < 4; it++)
        {
          const int src = it * 8 + (int)c8;
          const uint8_t* a = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)col, src);
          const uint32_t nf = __shfl_sync(0xffffffffu, szf, src);
          const uint32_t n = nf & ~CELL_SCALAR_FLAG;
          const uint32_t e0 = p0 + 4u * q;                       // first of the 4 particles this lane fetches
          const double* src_ptr = (const double*)a + e0;
          v[it] = d4{0.0, 0.0, 0.0, 0.0};
          if (e0 < n)
          {
            if (!(nf & CELL_SCALAR_FLAG)) v[it] = ld_stream_ro(src_ptr);   // capacity is a multiple of 4 => the whole vector is inside the column
            else
            {
              v[it].a = ld_stream_ro1(src_ptr);
              if (e0 + 1 < n) v[it].b = ld_stream_ro1(src_ptr + 1);
              if (e0 + 2 < n) v[it].c = ld_stream_ro1(src_ptr + 2);
              if (e0 + 3 < n) v[it].d = ld_stream_ro1(src_ptr + 3);
            }
            // particles beyond the cell's size (padding up to the capacity) count as +0.0
            if (e0 + 1 >= n) v[it].b = 0.0;
            if (e0 + 2 >= n) v[it].c = 0.0;
            if (e0 + 3 >= n) v[it].d = 0.0;
          }
        }
#pragma unroll
        for (int it = 0; it < 4; it++)
        {
          double* row = tile + (it * 8 + (int)c8) * CELL_ROW + 4 * q;
          sts128(row, v[it].a, v[it].b);
          sts128(row + 2, v[it].c, v[it].d);
        }
        __syncwarp();
        // particle order: s = ((s + e[p0]) + e[p0+1]) + ...   (+0.0 beyond the cell's size leaves s unchanged)
        const double* mine = tile + lane * CELL_ROW;
#pragma unroll
        for (int i = 0; i < 16; i += 2)
        {
          double x0, x1;
          lds128(mine + i, x0, x1);
          s = __dadd_rn(s, x0);
          s = __dadd_rn(s, x1);
        }
        __syncwarp();
      }
      if (cmine < ncells) { cell_sum[cmine] = s; thread_total = __dadd_rn(thread_total, s); }
      // rotate the pipeline
      off0 = off1; sz0 = sz1; cap0 = cap1;
      off1 = off2; sz1 = sz2; cap1 = cap2;
    }

    cell_total_finish(thread_total, partials, ticket, total);
  }

  // ------------------------------------------------------------------------------------------------------------
  // Asynchronously staged variant (default).  The cell columns go global -> shared with cp.async (LDGSTS.128, L1
  // bypass) into a per-warp ring of ASY_STAGES stages of 32 rows: 8 lanes move one cell's 128-byte line as eight
  // 16-byte pieces, the src-size operand zero-fills whatever lies beyond the cell's size (ragged tails, empty cells),
  // so nothing is held in registers and ASY_STAGES steps of 32 cells are in flight per warp: the
  // table -> address -> data latency chain of one step overlaps the in-order folds of the previous ones.
  // (A variant with one TMA bulk copy per cell -- cp.async.bulk, variable byte count per lane -- was measured at
  // 146 us on 128^3 cells against 105 us for the register-staged kernel above: ~20 cycles per small bulk request and
  // SM, the request rate and not the bandwidth being the limit at 128-256 B per copy.)
  // Particles beyond ASY_RCAP, columns that are not 16-byte aligned: summed straight from global memory by their lane.
  // shape of the ring: NWARPS warps per CTA (one CTA per SM), STAGES steps in flight per warp, RCAP particles staged per
  // cell (first 128-byte line + whatever of the second one fits); rows are RCAP+2 doubles apart (16-byte multiple, staggers
  // the banks).  8 warps x 3 stages x 32 particles and 12 x 2 x 32 fill the same 204 KB; more resident warps hide more of
  // the table -> address -> data latency, deeper rings keep more bytes in flight per warp.
  constexpr unsigned ASY_CTAS_PER_SM = 1;
  template<int NWARPS, int STAGES, int RCAP> constexpr size_t asy_smem() { return (size_t)NWARPS * STAGES * 32 * (RCAP + 2) * sizeof(double); }

  __device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
  __device__ __forceinline__ void cp_async16_zfill(unsigned dst, const void* src, unsigned src_bytes)
  {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
  }
  __device__ __forceinline__ double lds64_u32(unsigned a)
  {
    double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v;
  }
  __device__ __forceinline__ void lds128_u32(unsigned a, double& x0, double& x1)
  {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(a) : "memory");
  }

  // PACKED: field-major grid -- `arena` is the summed column of all cells back to back and cell_off holds ncells+1 element
  // offsets (cell c = [cell_off[c], cell_off[c+1])); rows are staged from the even element at or below the cell's first
  // one (16-byte aligned source), the possible leading stranger is replaced by +0.0 in the fold.
  template<bool PACKED, int NWARPS, int ASY_STAGES, int ASY_RCAP>
  __global__ void __launch_bounds__(NWARPS * 32, ASY_CTAS_PER_SM)
  cell_sum_f64_async_kernel(const uint8_t* __restrict__ arena, const uint64_t* __restrict__ cell_off, const uint32_t* __restrict__ cell_size,
                            const uint32_t* __restrict__ cell_cap, uint64_t ncells, const __grid_constant__ CellFieldParams fp,
                            double* __restrict__ cell_sum, double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ total)
  {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    constexpr int ASY_ROW = ASY_RCAP + 2;
    constexpr unsigned NT = NWARPS * 32;
    static_assert(ASY_RCAP >= 16 && ASY_RCAP <= 32 && ASY_RCAP % 2 == 0, "first line fully staged, second line partly");
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const unsigned piece = lane & 7u;          // which 16-byte piece of a 128-byte line this lane moves
    const unsigned g4 = lane >> 3;             // cell within the group of 4 served by one cp.async instruction
    double* ring = reinterpret_cast<double*>(dyn_smem) + (size_t)w * ASY_STAGES * 32 * ASY_ROW;

    const uint64_t warp_global = ((uint64_t)blockIdx.x * NT + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * NT) >> 5;
    const uint64_t nsteps_total = (ncells + 31) / 32;
    const uint64_t my_steps = (warp_global < nsteps_total) ? (nsteps_total - warp_global + nwarps - 1) / nwarps : 0;
    double thread_total = 0.0;

    // per-stage bookkeeping in registers (stage index is a compile-time constant after unrolling)
    const uint8_t* col_s[ASY_STAGES]; uint32_t sz_s[ASY_STAGES]; uint32_t staged_s[ASY_STAGES]; uint32_t odd_s[ASY_STAGES];

    // tables of the NEXT step to be issued, loaded one issue ahead so that their DRAM latency is never on the critical path
    uint64_t nx_off = 0; uint32_t nx_sz = 0, nx_cap = 0;
    auto load_next_tables = [&](uint64_t step)
    {
      const uint64_t c = (warp_global + step * nwarps) * 32 + lane;
      nx_off = 0; nx_sz = 0; nx_cap = 0;
      if (step < my_steps && c < ncells)
      {
        if constexpr (PACKED)
        {
          const uint64_t b = __ldg(cell_off + c), e = __ldg(cell_off + c + 1);
          nx_cap = (uint32_t)(b & 1ull);                     // leading stranger
          nx_off = (b - nx_cap) * 8ull;                      // byte offset of the even element the row starts at
          nx_sz = (uint32_t)(e - b) + nx_cap;                // row length including the stranger
        }
        else { nx_off = __ldg(cell_off + c); nx_sz = __ldg(cell_size + c); nx_cap = __ldg(cell_cap + c); }
      }
    };
    load_next_tables(0);

    auto issue = [&](int stage, uint64_t step)
    {
      const uint64_t off = nx_off; const uint32_t sz = nx_sz, cap = nx_cap;
      load_next_tables(step + 1);                        // steps are issued in order: the next call is for step+1
      const uint8_t* col = PACKED ? arena + off : arena + off + cell_field_offset(cap, fp);
      if constexpr (PACKED) odd_s[stage] = cap;
      uint32_t n_st = sz < (uint32_t)ASY_RCAP ? sz : (uint32_t)ASY_RCAP;
      if ((((uintptr_t)col) & 15u) != 0) n_st = 0;                 // unaligned column: this lane reads global memory itself
      col_s[stage] = col; sz_s[stage] = sz; staged_s[stage] = n_st;
      double* tile = ring + (size_t)stage * 32 * ASY_ROW;
#pragma unroll
      for (int it = 0; it < 8; it++)
      {
        const int src_lane = it * 4 + (int)g4;
        const uint8_t* a = (const uint8_t*)__shfl_sync(0xffffffffu, (unsigned long long)col, src_lane);
        const uint32_t nb = __shfl_sync(0xffffffffu, n_st, src_lane) * 8u;      // staged bytes of that cell
        const unsigned dst = smem_u32(tile + (size_t)src_lane * ASY_ROW) + piece * 16u;
        // first line: always written (zero-filled beyond the cell's size, also for empty cells)
        {
          const uint32_t o = piece * 16u;
          const uint32_t valid = (nb > o) ? ((nb - o < 16u) ? nb - o : 16u) : 0u;
          cp_async16_zfill(dst, a + o, valid);
        }
        // second line: only when some cell of this instruction is longer than 16 particles
        if (__any_sync(0xffffffffu, nb > 128u))
        {
          const uint32_t o = 128u + piece * 16u;
          const uint32_t valid = (nb > o) ? ((nb - o < 16u) ? nb - o : 16u) : 0u;
          if (o < (uint32_t)ASY_RCAP * 8u) cp_async16_zfill(dst + 128u, a + o, valid);     // the row ends at RCAP particles
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // prologue: fill the ring (empty groups are committed too so that group counting stays uniform)
#pragma unroll
    for (int sgi = 0; sgi < ASY_STAGES; sgi++)
    {
      if ((uint64_t)sgi < my_steps) issue(sgi, (uint64_t)sgi);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }

    for (uint64_t k0 = 0; k0 < my_steps; k0 += ASY_STAGES)
    {
#pragma unroll
      for (int sgi = 0; sgi < ASY_STAGES; sgi++)
      {
        const uint64_t k = k0 + sgi;
        if (k < my_steps)                                  // warp-uniform
        {
          asm volatile("cp.async.wait_group %0;" :: "n"(ASY_STAGES - 1) : "memory");   // the oldest group (this stage) has landed
          __syncwarp();
          const double* row = ring + ((size_t)sgi * 32 + lane) * ASY_ROW;
          const uint32_t n_st = staged_s[sgi], sz = sz_s[sgi];
          double s = 0.0;
          // particle order; zero-filled slots add +0.0, which leaves s unchanged
#pragma unroll
          for (int i = 0; i < 16; i += 2)
          {
            double x0, x1; lds128(row + i, x0, x1);
            if constexpr (PACKED) { if (i == 0 && odd_s[sgi] != 0u && n_st != 0u) x0 = 0.0; }
            s = __dadd_rn(s, x0); s = __dadd_rn(s, x1);
          }
          if (n_st > 16u)
          {
#pragma unroll
            for (int i = 16; i < ASY_RCAP; i += 2) { double x0, x1; lds128(row + i, x0, x1); s = __dadd_rn(s, x0); s = __dadd_rn(s, x1); }
          }
          if (sz > n_st)                                   // long or unaligned cells: the rest straight from global memory
          {
            const double* e = reinterpret_cast<const double*>(col_s[sgi]);
            uint32_t i0 = n_st;
            if constexpr (PACKED) { if (n_st == 0u) i0 = odd_s[sgi]; }     // nothing staged (unaligned column): skip the stranger here
            for (uint32_t i = i0; i < sz; i++) s = __dadd_rn(s, ld_stream_ro1(e + i));
          }
          const uint64_t c = (warp_global + k * nwarps) * 32 + lane;
          if (c < ncells) { cell_sum[c] = s; thread_total = __dadd_rn(thread_total, s); }
          __syncwarp();                                    // every lane is done reading this stage
          if (k + ASY_STAGES < my_steps) issue(sgi, k + ASY_STAGES);
          else asm volatile("cp.async.commit_group;" ::: "memory");
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    cell_total_finish<NT>(thread_total, partials, ticket, total);
  }
}

namespace onk
{
  // ------------------------------------------------------------------------------------------------------------
  // FIELD-MAJOR cell grid: the summed field of ALL cells is one contiguous column, cell c owning the run
  // [cell_begin[c], cell_begin[c+1]) (what an exaNBody-style grid calls its cell offsets).  The per-cell blocks of the
  // reference layout force 128 useful bytes out of every 512 through the memory system (DRAM traffic 1.4x the
  // algorithmic bytes, kernel above); in this layout the kernel is a pure stream: a CTA of NT threads takes NT consecutive
  // cells per step, moves their values global -> shared with cp.async and thread t folds cell t in particle order from
  // shared memory -- same bits as the per-cell host loop.  17 values per cell are staged on average (Poisson(16) tiles
  // fit with 4 sigma to spare); whatever exceeds that is read from global memory by the owning thread.

  // Tile staging: values go global -> shared as 16-byte pieces (LDGSTS.128 bypassing L1; the 8-byte form drags three
  // no-op shared loads per copy along on sm_100), starting at the 16-byte aligned element at or below the tile's first
  // one.  Local value i lives in slot i + 2*(i/32): pairs stay together and 16-byte aligned, and runs that start a
  // multiple of 16 values apart (uniform 16-particle cells) spread over the banks instead of piling up on one.
  // STAGES tiles per CTA ring, CTAS resident CTAs per SM: the folds of one CTA overlap the copies of the others.
  __device__ __forceinline__ void cp_async16(unsigned dst, const void* src)
  {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
  }

  // In-order fold of doubles [qb, qe) of a tile staged with the 128-byte swizzle (16-byte chunk c of 128-byte row r at chunk
  // c ^ (r & 7); `tile` = shared address of the tile, 1024-byte aligned).  Row by row: the row base is 128-byte aligned, so the
  // address of chunk c is (base ^ ((r & 7) << 4)) ^ (c << 4) -- one XOR with an immediate per LDS.128 once the row term is
  // hoisted (the per-element form recomputes the whole slot expression: 7 integer instructions per load).  Whole rows are
  // unrolled.  Worth 6 % on the all-fields kernel, whose columns are whole rows (6.45 -> 6.83 TB/s at 256^3 cells); the
  // field-major kernel, whose cells start anywhere, keeps the per-element form (no gain, more registers).
  __device__ __forceinline__ double fold_swizzled(unsigned tile, uint32_t qb, uint32_t qe, double s)
  {
    uint32_t q = qb;
    if (q >= qe) return s;
    if (q & 1u)
    {
      const uint32_t r = q >> 4;
      s = __dadd_rn(s, lds64_u32(((tile + (r << 7)) ^ ((r & 7u) << 4)) ^ (((q >> 1) & 7u) << 4) | 8u));
      ++q;
    }
    while (q + 2u <= qe)
    {
      const uint32_t r = q >> 4;
      const unsigned bx = (tile + (r << 7)) ^ ((r & 7u) << 4);
      const uint32_t row_end = (r + 1u) << 4;
      if ((q & 15u) == 0u && row_end <= qe)
      {
#pragma unroll
        for (unsigned c = 0; c < 8u; c++)
        {
          double x0, x1; lds128_u32(bx ^ (c << 4), x0, x1);
          s = __dadd_rn(s, x0); s = __dadd_rn(s, x1);
        }
        q = row_end;
      }
      else
      {
        const uint32_t stop = row_end < qe ? row_end : qe;
        unsigned a = (q >> 1) & 7u;
#pragma unroll 4
        for (; q + 2u <= stop; q += 2u, a++)
        {
          double x0, x1; lds128_u32(bx ^ (a << 4), x0, x1);
          s = __dadd_rn(s, x0); s = __dadd_rn(s, x1);
        }
        if (q < stop) break;                    // a single element is left (only possible at qe)
      }
    }
    if (q < qe)
    {
      const uint32_t r = q >> 4;
      s = __dadd_rn(s, lds64_u32(((tile + (r << 7)) ^ ((r & 7u) << 4)) ^ (((q >> 1) & 7u) << 4)));
    }
    return s;
  }

  // One tile in flight per CTA, CTAS resident CTAs per SM (resident warps, not a deeper ring per CTA, hide the latency
  // here: 3 stages x 2 CTAs 77 us, 2 x 3 67 us, 1 x 6 66 us, 1 x 5 64 us on 128^3 cells).  Tiles are handed out
  // dynamically: the first one is the CTA's own index, further ones come from a global ticket drawn by thread 0 one tile
  // ahead (so the draw, like the offset tables, is never waited for).  Tiles hold equal numbers of cells, not of
  // particles: with uneven densities a static round robin would leave SMs idle; on the uniform Poisson(16) grid both
  // measure the same 64 us.  Every CTA draws until its first out-of-range ticket: exactly ntiles draws in total, and
  // whoever draws the last one puts the counter back to zero for the next launch.
  template<int CTAS, int NT>
  __global__ void __launch_bounds__(NT, CTAS)
  cell_sum_packed_f64_kernel(const double* __restrict__ values, const uint64_t* __restrict__ cell_begin, uint64_t ncells,
                             double* __restrict__ cell_sum, double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ total)
  {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ unsigned long long s_next_tile[2];
    const unsigned tile_smem = smem_u32(dyn_smem);
    constexpr int PK_CELLS = NT;                                             // cells per tile (one per thread)
    constexpr int PK_CAP = 17 * NT;                                          // values staged per tile
    constexpr int COPY_ITERS = (PK_CAP / 2 + NT - 1) / NT;       // 16-byte pieces per thread and tile
    const uint64_t ntiles = (ncells + PK_CELLS - 1) / PK_CELLS;
    const unsigned parity = (unsigned)(((uintptr_t)values >> 3) & 1u);   // element g is 16-byte aligned iff g + parity is even
    const unsigned pair_slot0 = (threadIdx.x + (threadIdx.x >> 4)) * 16u;  // byte offset of pair j = tid + NT*it: pair_slot0 + (NT + NT/16)*16*it
    unsigned* tile_counter = ticket + 1;
    double thread_total = 0.0;
    bool drawing = true;                                                  // thread 0: this CTA has not yet drawn an out-of-range ticket

    auto draw = [&](int slot)                                             // thread 0 only
    {
      unsigned long long t = ~0ull;
      if (drawing)
      {
        const unsigned old = atomicAdd(tile_counter, 1u);
        t = (unsigned long long)gridDim.x + old;
        if (t >= ntiles) drawing = false;
        if ((uint64_t)old + 1u == ntiles) atomicExch(tile_counter, 0u);   // the last draw of the launch
      }
      s_next_tile[slot] = t;
    };

    // offsets of a tile, loaded one tile ahead of their use
    uint64_t nx_b = 0, nx_e = 0, nx_base = 0, nx_end = 0;
    auto load_tables = [&](uint64_t tile)
    {
      nx_b = nx_e = nx_base = nx_end = 0;
      if (tile < ntiles)
      {
        const uint64_t c0 = tile * PK_CELLS, c = c0 + threadIdx.x;
        const uint64_t cend = (c0 + PK_CELLS < ncells) ? c0 + PK_CELLS : ncells;
        nx_base = __ldg(cell_begin + c0); nx_end = __ldg(cell_begin + cend);
        if (c < ncells) { nx_b = __ldg(cell_begin + c); nx_e = __ldg(cell_begin + c + 1); } else { nx_b = nx_e = nx_end; }
      }
    };

    uint32_t pb = 0, pe = 0, cnt = 0;
    // copies the tile whose offsets are in nx_*, then fetches the offsets of `following`
    auto issue = [&](uint64_t following)
    {
      // the staged window starts at the aligned element at or below the tile's first value: one stranger in front at most
      const uint64_t lead = (nx_base + parity) & 1ull;
      const uint64_t span = nx_end - nx_base + lead;
      const uint64_t rb = nx_b - nx_base + lead, re = nx_e - nx_base + lead;
      pb = rb < 0xffffffffull ? (uint32_t)rb : 0xffffffffu;
      pe = re < 0xffffffffull ? (uint32_t)re : 0xffffffffu;
      cnt = span < (uint64_t)PK_CAP ? (uint32_t)span : (uint32_t)PK_CAP;
      const uint4* src = reinterpret_cast<const uint4*>(values + ((int64_t)nx_base - (int64_t)lead)) + threadIdx.x;
      load_tables(following);
      const unsigned dst = tile_smem + pair_slot0;
      const uint32_t npairs = cnt >> 1;
#pragma unroll
      for (int it = 0; it < COPY_ITERS; it++)
        if (threadIdx.x + (unsigned)it * NT < npairs) cp_async16(dst + (unsigned)it * ((unsigned)(NT + NT / 16) * 16u), src + it * NT);
      if ((cnt & 1u) != 0u && threadIdx.x == (npairs & (unsigned)(NT - 1)))       // odd tail: 8 valid bytes, the rest zero-filled
        cp_async16_zfill(dst + (npairs / NT) * ((unsigned)(NT + NT / 16) * 16u), src + (npairs / NT) * NT, 8u);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint64_t cur = blockIdx.x;                                            // tile being copied / folded
    if (threadIdx.x == 0) draw(0);
    load_tables(cur);
    __syncthreads();
    uint64_t nxt = s_next_tile[0];
    if (cur < ntiles) issue(nxt);

    for (int slot = 1; cur < ntiles; slot ^= 1)                           // CTA-uniform
    {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();                                                    // every thread's pieces of this tile have landed
      if (threadIdx.x == 0) draw(slot);                                   // ticket of the tile after the next one
      const uint32_t staged_end = pe < cnt ? pe : cnt;
      double s = 0.0;
      if (pb < staged_end)                                                // particle order; value p sits in slot p + 2*(p/32)
      {
        uint32_t p = pb;
        unsigned a = tile_smem + (p + 2u * (p >> 5)) * 8u;
        if (p & 1u) { s = __dadd_rn(s, lds64_u32(a)); ++p; a += (p & 31u) ? 8u : 24u; }
        uint32_t npair = (staged_end - p) >> 1;
        uint32_t r = (32u - (p & 31u)) >> 1;                              // pairs left before the next skipped slot pair
        for (; npair != 0u; --npair)
        {
          double x0, x1; lds128_u32(a, x0, x1);
          s = __dadd_rn(s, x0); s = __dadd_rn(s, x1);
          a += 16u;
          if (--r == 0u) { a += 16u; r = 16u; }
        }
        if ((staged_end - p) & 1u) s = __dadd_rn(s, lds64_u32(a));
      }
      const uint64_t c = cur * PK_CELLS + threadIdx.x;
      if (pe > cnt && c < ncells)                                         // long tiles: the rest of the run straight from global memory
      {
        const uint64_t gb = __ldg(cell_begin + c), ge = __ldg(cell_begin + c + 1);
        const uint64_t staged_g = gb + (uint64_t)(staged_end > pb ? staged_end - pb : 0u);
        for (uint64_t g = staged_g; g < ge; g++) s = __dadd_rn(s, ld_stream_ro1(values + g));
      }
      if (c < ncells) { cell_sum[c] = s; thread_total = __dadd_rn(thread_total, s); }
      __syncthreads();                                                    // everybody is done reading the tile; the new ticket is visible
      cur = nxt;
      nxt = s_next_tile[slot];
      if (cur < ntiles) issue(nxt);
    }
    cell_total_finish<NT>(thread_total, partials, ticket, total);
  }
}

namespace onk
{
  // ---- field-major grid, TMA-staged: the tile is ONE tensor-map bulk copy ----
  // The value column is described to the TMA unit as a 2D tensor of 16-double (128-byte) rows; a tile of NT cells is the box
  // of BOXROWS rows starting at the row that holds the tile's first value, brought in by one cp.async.bulk.tensor issued by
  // one thread (no per-thread copy loop, no slot arithmetic on the way in) and laid out with the 128-byte swizzle: the
  // 16-byte chunk c of row r lands at chunk c ^ (r & 7), so cells that start a multiple of 16 values apart (uniform
  // 16-particle cells: every thread at chunk 0 of its own row) spread over the banks.  Rows past the last full row of the
  // column are never requested (out-of-bounds rows are zero-filled by the unit); their values, like those of tiles longer
  // than the box, are read from global memory by the owning thread.
  // The tensor map's box is SUBROWS (16) rows = 2 KiB; a tile is brought in by as many such copies as its values span (up to
  // the BOXROWS rows of the staging buffer), all completing on the same mbarrier: nothing beyond the tile's last row is
  // requested (one fixed box per tile had to be sized for the densest tile and re-read 25 % of the column through L2).
  constexpr int TMA_SUBROWS = 16;
  template<int CTAS, int NT, int BOXROWS, bool ROWFOLD = true>
  __global__ void __launch_bounds__(NT, CTAS)
  cell_sum_packed_tma_kernel(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ values, uint64_t nrows_full,
                             const uint64_t* __restrict__ cell_begin, uint64_t ncells,
                             double* __restrict__ cell_sum, double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ total)
  {
    extern __shared__ unsigned char dyn_smem[];
    __shared__ unsigned long long s_next_tile[2];
    __shared__ __align__(8) unsigned long long s_bar;
    static_assert(BOXROWS % TMA_SUBROWS == 0, "the staging buffer holds whole copies");
    const unsigned tile_smem = (smem_u32(dyn_smem) + 1023u) & ~1023u;     // the swizzle pattern is anchored at 1024-byte boundaries
    const unsigned bar = smem_u32(&s_bar);
    const uint64_t ntiles = (ncells + NT - 1) / NT;
    unsigned* tile_counter = ticket + 1;
    double thread_total = 0.0;
    bool drawing = true;
    if (threadIdx.x == 0)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    auto draw = [&](int slot)                                             // thread 0 only
    {
      unsigned long long t = ~0ull;
      if (drawing)
      {
        const unsigned old = atomicAdd(tile_counter, 1u);
        t = (unsigned long long)gridDim.x + old;
        if (t >= ntiles) drawing = false;
        if ((uint64_t)old + 1u == ntiles) atomicExch(tile_counter, 0u);   // the last draw of the launch
      }
      s_next_tile[slot] = t;
    };

    uint64_t nx_b = 0, nx_e = 0, nx_base = 0, nx_end = 0;
    auto load_tables = [&](uint64_t tile)
    {
      nx_b = nx_e = nx_base = nx_end = 0;
      if (tile < ntiles)
      {
        const uint64_t c0 = tile * NT, c = c0 + threadIdx.x;
        const uint64_t cl = (c0 + NT < ncells) ? c0 + NT : ncells;
        nx_base = __ldg(cell_begin + c0);
        nx_end = __ldg(cell_begin + cl);                                  // one past the tile's last value (same address for all threads)
        if (c < ncells) { nx_b = __ldg(cell_begin + c); nx_e = __ldg(cell_begin + c + 1); } else { nx_b = nx_e = nx_base; }
      }
    };

    uint32_t pb = 0, pe = 0, cnt = 0; uint64_t row0 = 0;
    auto issue = [&](uint64_t following)
    {
      row0 = nx_base >> 4;                                                // first row of the tile
      const uint64_t first = row0 << 4;
      const uint64_t rb = nx_b - first, re = nx_e - first;
      pb = rb < 0xffffffffull ? (uint32_t)rb : 0xffffffffu;
      pe = re < 0xffffffffull ? (uint32_t)re : 0xffffffffu;
      const uint64_t rows_avail = row0 < nrows_full ? nrows_full - row0 : 0;
      uint64_t rows = (nx_end - first + 15u) >> 4;                        // rows the tile's values span
      if (rows > rows_avail) rows = rows_avail;
      if (rows > (uint64_t)BOXROWS) rows = BOXROWS;
      const uint32_t ncopies = (uint32_t)((rows + TMA_SUBROWS - 1) / TMA_SUBROWS);
      cnt = (uint32_t)(rows << 4);                                        // staged values: whole rows that exist and fit the buffer
      if (threadIdx.x == 0)
      {
        // (a tile without any full row still arrives once, with nothing to wait for)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(ncopies * (unsigned)(TMA_SUBROWS * 128)) : "memory");
        for (uint32_t k = 0; k < ncopies; k++)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       :: "r"(tile_smem + k * (unsigned)(TMA_SUBROWS * 128)), "l"(&tmap), "r"(0), "r"((int)(row0 + (uint64_t)k * TMA_SUBROWS)), "r"(bar) : "memory");
      }
      load_tables(following);
    };

    uint64_t cur = blockIdx.x;
    if (threadIdx.x == 0) draw(0);
    load_tables(cur);
    __syncthreads();                                                      // barrier initialised, first ticket visible
    uint64_t nxt = s_next_tile[0];
    if (cur < ntiles) issue(nxt);
    unsigned phase = 0;

    for (int slot = 1; cur < ntiles; slot ^= 1)                           // CTA-uniform
    {
      unsigned landed = 0;
      while (!landed)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(landed) : "r"(bar), "r"(phase) : "memory");
      phase ^= 1u;
      if (threadIdx.x == 0) draw(slot);                                   // ticket of the tile after the next one
      const uint32_t staged_end = pe < cnt ? pe : cnt;
      double s = 0.0;
      if constexpr (ROWFOLD) s = fold_swizzled(tile_smem, pb, staged_end, 0.0);           // particle order
      else if (pb < staged_end)
      {
        auto slot_addr = [&](uint32_t q) -> unsigned { const uint32_t row = q >> 4; return tile_smem + (row << 7) + ((((q >> 1) ^ row) & 7u) << 4) + ((q & 1u) << 3); };
        uint32_t q = pb;
        if (q & 1u) { s = __dadd_rn(s, lds64_u32(slot_addr(q))); ++q; }
        for (; q + 2u <= staged_end; q += 2u)
        {
          double x0, x1; lds128_u32(slot_addr(q), x0, x1);
          s = __dadd_rn(s, x0); s = __dadd_rn(s, x1);
        }
        if (q < staged_end) s = __dadd_rn(s, lds64_u32(slot_addr(q)));
      }
      const uint64_t c = cur * NT + threadIdx.x;
      if (pe > cnt && c < ncells)                                         // beyond the box or in the last partial row: global memory
      {
        const uint64_t first = row0 << 4;
        const uint64_t gb = first + (uint64_t)(pb > staged_end ? pb : staged_end), ge = first + (uint64_t)pe;
        for (uint64_t g = gb; g < ge; g++) s = __dadd_rn(s, ld_stream_ro1(values + g));
      }
      if (c < ncells) { cell_sum[c] = s; thread_total = __dadd_rn(thread_total, s); }
      __syncthreads();                                                    // everybody is done reading the tile; the new ticket is visible
      cur = nxt;
      nxt = s_next_tile[slot];
      if (cur < ntiles) issue(nxt);
    }
    cell_total_finish<NT>(thread_total, partials, ticket, total);
  }
}

namespace onk
{
  // ---- per-cell sums of EVERY fp64 field of the cell blocks (config C4, "all four fields streamed": 32 B per particle) ----
  // When all the fields of a cell are wanted the per-cell layout IS a contiguous stream: cell blocks lie back to back in the
  // arena, so a tile of consecutive cells is one byte range.  The arena is described to the TMA unit as a 2D tensor of
  // 128-byte rows; a tile of NT/nfields cells is the run of rows from the one holding the tile's first byte to the one holding
  // its last -- 16-row (2 KiB) cp.async.bulk.tensor copies issued by one thread onto one mbarrier, exactly as many as the tile
  // spans -- laid out with the 128-byte swizzle (16-byte chunk c of row r at chunk c ^ (r & 7)) so that the threads, each
  // folding its own column from its own row(s), spread over the banks.  Thread t
  // owns column (cell t / nfields, field t % nfields) and folds it in particle order (bit-identical to the host loop); the
  // nfields sums of a cell leave as consecutive doubles (coalesced: thread t -> element t of the tile's output).  Columns
  // that reach beyond the box (dense tiles) or beyond the last full row of the arena finish from global memory.  Tiles are
  // drawn from a global ticket one tile ahead, as in the field-major kernel above.  Per-field totals: per-CTA partials in
  // shared memory (thread order), then the ticketed last CTA folds the CTAs' partials in index order.
  constexpr int CELL_ALLF_MAX = 8;
  struct CellAllFieldsParams { uint32_t nfields; uint32_t align_mask; };

  template<int CTAS, int NT, int BOXROWS>
  __global__ void __launch_bounds__(NT, CTAS)
  cell_sum_fields_tma_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ arena, uint64_t arena_bytes, uint64_t nrows_full,
                             const uint64_t* __restrict__ cell_off, const uint32_t* __restrict__ cell_size, const uint32_t* __restrict__ cell_cap,
                             uint64_t ncells, const __grid_constant__ CellAllFieldsParams prm,
                             double* __restrict__ cell_sums, double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ totals)
  {
    extern __shared__ unsigned char dyn_smem[];
    __shared__ unsigned long long s_next_tile[2];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ double s_thread_total[NT];
    __shared__ bool s_is_last;
    static_assert(BOXROWS % TMA_SUBROWS == 0, "the staging buffer holds whole copies");
    const unsigned tile_smem = (smem_u32(dyn_smem) + 1023u) & ~1023u;     // the swizzle pattern is anchored at 1024-byte boundaries
    const unsigned bar = smem_u32(&s_bar);
    const unsigned nf = prm.nfields;
    const unsigned cpt = (unsigned)NT / nf;                                // cells per tile
    const unsigned lc = threadIdx.x / nf, f = threadIdx.x - lc * nf;       // this thread's column: (local cell, field)
    const bool active = lc < cpt;
    const uint64_t ntiles = (ncells + cpt - 1) / cpt;
    unsigned* tile_counter = ticket + 1;
    double thread_total = 0.0;
    bool drawing = true;
    if (threadIdx.x == 0)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto draw = [&](int slot)                                             // thread 0 only
    {
      unsigned long long t = ~0ull;
      if (drawing)
      {
        const unsigned old = atomicAdd(tile_counter, 1u);
        t = (unsigned long long)gridDim.x + old;
        if (t >= ntiles) drawing = false;
        if ((uint64_t)old + 1u == ntiles) atomicExch(tile_counter, 0u);   // the last draw of the launch
      }
      s_next_tile[slot] = t;
    };

    // tables of the tile after the current one, loaded while the current one is in flight
    uint64_t nx_base = 0, nx_end = 0, nx_col = 0; uint32_t nx_sz = 0;
    auto load_tables = [&](uint64_t tile)
    {
      nx_base = nx_end = nx_col = 0; nx_sz = 0;
      if (tile < ntiles)
      {
        const uint64_t c0 = tile * cpt, c = c0 + lc;
        nx_base = __ldg(cell_off + c0);
        nx_end = (c0 + cpt < ncells) ? __ldg(cell_off + c0 + cpt) : arena_bytes;      // one past the tile's last byte
        if (active && c < ncells)
        {
          const uint32_t cap = __ldg(cell_cap + c);
          const uint64_t stride = ((uint64_t)cap * 8ull + prm.align_mask) & ~(uint64_t)prm.align_mask;    // fp64 column of FieldArrays<align,...>
          nx_col = __ldg(cell_off + c) + (uint64_t)f * stride;
          nx_sz = __ldg(cell_size + c);
        }
      }
    };

    uint64_t col = 0, row0 = 0; uint32_t sz = 0, cnt = 0;
    auto issue = [&](uint64_t following)
    {
      row0 = nx_base >> 7;                                                // first 128-byte row of the box
      col = nx_col; sz = nx_sz;
      const uint64_t rows_avail = row0 < nrows_full ? nrows_full - row0 : 0;
      uint64_t rows = (nx_end - (row0 << 7) + 127u) >> 7;                 // rows the tile's cell blocks span
      if (rows > rows_avail) rows = rows_avail;
      if (rows > (uint64_t)BOXROWS) rows = BOXROWS;
      const uint32_t ncopies = (uint32_t)((rows + TMA_SUBROWS - 1) / TMA_SUBROWS);
      cnt = (uint32_t)(rows << 4);                                        // staged doubles: whole rows that exist and fit the buffer
      if (threadIdx.x == 0)
      {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(ncopies * (unsigned)(TMA_SUBROWS * 128)) : "memory");
        for (uint32_t k = 0; k < ncopies; k++)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       :: "r"(tile_smem + k * (unsigned)(TMA_SUBROWS * 128)), "l"(&tmap), "r"(0), "r"((int)(row0 + (uint64_t)k * TMA_SUBROWS)), "r"(bar) : "memory");
      }
      load_tables(following);
    };

    uint64_t cur = blockIdx.x;
    if (threadIdx.x == 0) draw(0);
    load_tables(cur);
    __syncthreads();                                                      // barrier initialised, first ticket visible
    uint64_t nxt = s_next_tile[0];
    if (cur < ntiles) issue(nxt);
    unsigned phase = 0;

    for (int slot = 1; cur < ntiles; slot ^= 1)                           // CTA-uniform
    {
      unsigned landed = 0;
      while (!landed)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(landed) : "r"(bar), "r"(phase) : "memory");
      phase ^= 1u;
      if (threadIdx.x == 0) draw(slot);                                   // ticket of the tile after the next one
      const uint64_t c = cur * cpt + lc;
      if (active && c < ncells)
      {
        const uint64_t q0 = (col - (row0 << 7)) >> 3;                     // first double of the column, relative to the box
        const uint64_t qe = q0 + sz;
        const uint32_t staged_end = (uint32_t)(qe < (uint64_t)cnt ? qe : (q0 < (uint64_t)cnt ? (uint64_t)cnt : q0));
        double s = 0.0;
        if (q0 < (uint64_t)cnt) s = fold_swizzled(tile_smem, (uint32_t)q0, staged_end, 0.0);    // particle order
        if (qe > (uint64_t)staged_end)                                    // beyond the box / the last full row: global memory
        {
          const double* g = reinterpret_cast<const double*>(arena + col);
          const uint64_t done = (q0 < (uint64_t)cnt) ? (uint64_t)staged_end - q0 : 0ull;
          for (uint64_t i = done; i < (uint64_t)sz; i++) s = __dadd_rn(s, ld_stream_ro1(g + i));
        }
        cell_sums[c * nf + f] = s;
        thread_total = __dadd_rn(thread_total, s);
      }
      __syncthreads();                                                    // everybody is done reading the tile; the new ticket is visible
      cur = nxt;
      nxt = s_next_tile[slot];
      if (cur < ntiles) issue(nxt);
    }

    // per-field totals
    if (totals == nullptr) return;
    s_thread_total[threadIdx.x] = active ? thread_total : 0.0;
    __syncthreads();
    if (threadIdx.x < nf)
    {
      double r = 0.0;
      for (unsigned t = threadIdx.x; t < cpt * nf; t += nf) r = __dadd_rn(r, s_thread_total[t]);     // thread order: fixed
      partials[(size_t)blockIdx.x * CELL_ALLF_MAX + threadIdx.x] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      __threadfence();
      s_is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_is_last)
    {
      __threadfence();
      if (threadIdx.x < nf)
      {
        double r = 0.0;
        for (unsigned b = 0; b < gridDim.x; b++) r = __dadd_rn(r, __ldcg(partials + (size_t)b * CELL_ALLF_MAX + threadIdx.x));
        totals[threadIdx.x] = r;
      }
      if (threadIdx.x == 0) *ticket = 0u;
    }
  }

  // fallback for arenas the tensor map cannot describe (unaligned base, less than one row): one thread per column, global loads
  __global__ void __launch_bounds__(THREADS)
  cell_sum_fields_plain_kernel(const uint8_t* __restrict__ arena, const uint64_t* __restrict__ cell_off, const uint32_t* __restrict__ cell_size,
                               const uint32_t* __restrict__ cell_cap, uint64_t ncells, const __grid_constant__ CellAllFieldsParams prm,
                               double* __restrict__ cell_sums)
  {
    const uint64_t ncol = ncells * prm.nfields;
    for (uint64_t t = (uint64_t)blockIdx.x * THREADS + threadIdx.x; t < ncol; t += (uint64_t)gridDim.x * THREADS)
    {
      const uint64_t c = t / prm.nfields; const unsigned f = (unsigned)(t - c * prm.nfields);
      const uint64_t stride = ((uint64_t)__ldg(cell_cap + c) * 8ull + prm.align_mask) & ~(uint64_t)prm.align_mask;
      const double* g = reinterpret_cast<const double*>(arena + __ldg(cell_off + c) + (uint64_t)f * stride);
      const uint32_t n = __ldg(cell_size + c);
      double s = 0.0;
      for (uint32_t i = 0; i < n; i++) s = __dadd_rn(s, ld_stream_ro1(g + i));
      cell_sums[t] = s;
    }
  }

  // totals of the plain path: column f of the ncells x nfields result, one CTA, fixed order (rarely used: tiny or odd arenas)
  __global__ void __launch_bounds__(THREADS)
  cell_sum_fields_totals_kernel(const double* __restrict__ cell_sums, uint64_t ncells, unsigned nf, double* __restrict__ totals)
  {
    for (unsigned f = 0; f < nf; f++)
    {
      double r = 0.0;
      for (uint64_t c = threadIdx.x; c < ncells; c += THREADS) r = __dadd_rn(r, cell_sums[c * nf + f]);
      r = block_reduce<ONK_OP_SUM, THREADS>(r);
      if (threadIdx.x == 0) totals[f] = r;
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
  // "ldg" = register-staged kernel; default = cp.async ring (shapes below): this layout is bound by its DRAM access
  // pattern (0.42 GB of full 128-byte lines for 0.30 GB of payload), not by the SM.
  static const int variant = []{ const char* e = getenv("ONK_CELL_KERNEL"); return (e && e[0] == 'l') ? 0 : 1; }();
  if (variant == 1)
  {
#define ONK_ASY_GO(NW, ST, RC) do { \
    static bool configured = false; \
    if (!configured) { ONK_CUDA(cudaFuncSetAttribute(cell_sum_f64_async_kernel<false,NW,ST,RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asy_smem<NW,ST,RC>())); configured = true; } \
    grid = grid_for(ncells, NW * 32, ASY_CTAS_PER_SM); \
    cell_sum_f64_async_kernel<false,NW,ST,RC><<<grid, NW * 32, asy_smem<NW,ST,RC>(), L->stream>>>((const uint8_t*)arena_dev, cell_off_dev, cell_size_dev, cell_cap_dev, ncells, fp, \
                                                                                             cell_sum_dev, L->ws.partials, L->ws.ticket, total_dev); } while (0)
    // ring shape, measured on 128^3 / 256^3 cells (tools/exp_cell_ring.py, cold L2): 16 warps x 1 stage x 32 particles
    // 78.8 us / 540 us (default); 8 x 3 x 32 (round 1) 80.7 / 552; 12 x 2 x 32 80.9 / 542; 20 x 1 x 32 80.9 / 572;
    // 24 x 1 x 32 80.9 / 579; 16 x 1 x 24 82.9 / 558.  A compact ring (32 first lines + second lines only for the cells
    // that have one, 4 KB + 2.5 KB per stage instead of 8.7 KB, hence 12-16 warps with 2-3 stages) was built and measured at
    // 87-93 us: the extra bookkeeping of the compaction costs more than the resident warps bring.  The line pattern itself
    // can be fetched in 62-64 us by >= 12 resident warps (tools/exp_gather4.cu, list-driven ring, TMA gather4 or cp.async
    // alike); tables and results add 13 % of traffic: this kernel is within ~10 % of what the layout allows
    // (profiles/r2_c4_floor.md).
    static const int shape = []{ const char* e = getenv("ONK_CELL_SHAPE"); return e ? atoi(e) : 0; }();      // tuning / comparison only
    if (shape == 1) ONK_ASY_GO(8, 3, 32);
    else ONK_ASY_GO(16, 1, 32);
#undef ONK_ASY_GO
  }
  else
    cell_sum_f64_kernel<<<grid, THREADS, 0, L->stream>>>((const uint8_t*)arena_dev, cell_off_dev, cell_size_dev, cell_cap_dev, ncells, fp,
                                                         cell_sum_dev, L->ws.partials, L->ws.ticket, total_dev);
  ONK_CHECK_LAUNCH(); count_launch();
  return ONK_OK;
}

int onk_cell_sum_packed_f64(const double* values_dev, uint64_t nvalues, const uint64_t* cell_begin_dev, uint64_t ncells,
                            double* cell_sum_dev, double* total_dev, int lane)
{
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  if (ncells == 0)
  {
    if (total_dev) ONK_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(double), L->stream));
    return ONK_OK;
  }
  ONK_REQUIRE(cell_begin_dev && cell_sum_dev, "onk_cell_sum_packed_f64: null pointer");
  ONK_REQUIRE(((uintptr_t)values_dev & 7) == 0, "onk_cell_sum_packed_f64: values are not 8-byte aligned");
  // ONK_PK_CFG=rows (tuning only): the row-staged kernel of the per-cell layout fed from the packed column (83 us on 128^3 cells)
  static const bool rows_variant = []{ const char* e = getenv("ONK_PK_CFG"); return e && e[0] == 'r'; }();
  if (rows_variant)
  {
    static bool rows_configured = false;
    if (!rows_configured) { ONK_CUDA(cudaFuncSetAttribute(cell_sum_f64_async_kernel<true,8,3,32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asy_smem<8,3,32>())); rows_configured = true; }
    unsigned g = grid_for(ncells, (THREADS / 32) * 32, ASY_CTAS_PER_SM);
    if (g > (unsigned)MAX_PARTIALS) g = MAX_PARTIALS;
    cell_sum_f64_async_kernel<true,8,3,32><<<g, THREADS, asy_smem<8,3,32>(), L->stream>>>((const uint8_t*)values_dev, cell_begin_dev, nullptr, nullptr, ncells, CellFieldParams{},
                                                                       cell_sum_dev, L->ws.partials, L->ws.ticket, total_dev);
    ONK_CHECK_LAUNCH(); count_launch();
    return ONK_OK;
  }
  // TMA-staged kernel when the column can be described by a tensor map (16-byte aligned, at least one full 128-byte row);
  // ONK_PK_KERNEL=ldgsts forces the cp.async kernel below (tuning / comparison)
  static const bool allow_tma = []{ const char* e = getenv("ONK_PK_KERNEL"); return !(e && e[0] == 'l'); }();
  const uint64_t nrows_full = nvalues / 16;
  if (allow_tma && ((uintptr_t)values_dev & 15) == 0 && nrows_full > 0 && nrows_full < (1ull << 31))
  {
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_t encode = nullptr;
    if (!encode)
    {
      void* fn = nullptr; cudaDriverEntryPointQueryResult q;
      ONK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
      if (!fn) { set_error("onk_cell_sum_packed_f64: cuTensorMapEncodeTiled entry point not found"); return ONK_ERR_CUDA; }
      encode = (encode_t)fn;
    }
    constexpr int l2p = (int)CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
#define ONK_PK_TMA_GO(C, NTH, BR, RF) do { \
    CUtensorMap tmap; \
    const cuuint64_t gdim[2] = { 16, (cuuint64_t)nrows_full }; \
    const cuuint64_t gstride[1] = { 128 }; \
    const cuuint32_t box[2] = { 16, (cuuint32_t)TMA_SUBROWS }; \
    const cuuint32_t estr[2] = { 1, 1 }; \
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)values_dev, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, \
                              CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); \
    if (r != CUDA_SUCCESS) { set_error("onk_cell_sum_packed_f64: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return ONK_ERR_CUDA; } \
    constexpr size_t smem = (size_t)BR * 128 + 1024; \
    static bool configured = false; \
    if (!configured) { ONK_CUDA(cudaFuncSetAttribute(cell_sum_packed_tma_kernel<C, NTH, BR, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; } \
    unsigned grid = grid_for(ncells, NTH, C); \
    if (grid > (unsigned)MAX_PARTIALS) grid = MAX_PARTIALS; \
    cell_sum_packed_tma_kernel<C, NTH, BR, RF><<<grid, NTH, smem, L->stream>>>(tmap, values_dev, nrows_full, cell_begin_dev, ncells, cell_sum_dev, \
                                                                           L->ws.partials, L->ws.ticket, total_dev); } while (0)
    // shape: 8 resident CTAs of 128 threads, 160-row (20 KiB) staging buffers for tiles of 128 rows on average.  Measured with
    // the C ABI call, cold L2, 128^3 / 256^3 cells (tools/exp_field_major_fold.py): 8x128x160 56.4 us / 357 us (6.76 TB/s),
    // 9x128x160 with the row-wise fold 58.4 / 355, 10x128x160 (round 1: capped at 48 registers) 56.3 / 466, 7x128x160 58.4 / 374,
    // 6x128x160 62.5 / 400, 12x128x144 68.6 / 441, 6x192x256 60.4 / 365, 4x256x320 64.5 / 400: at scale FEWER resident CTAs
    // with uncapped registers win; L2 promotion none / 64 B / 128 B within 0.1 us, 256 B +2 us
    static const int shape = []{ const char* e = getenv("ONK_PK_FOLD"); return e ? atoi(e) : 0; }();      // tuning / comparison only
    if (shape == 1) ONK_PK_TMA_GO(10, 128, 160, false);
    else if (shape == 4) ONK_PK_TMA_GO(9, 128, 160, true);
    else ONK_PK_TMA_GO(8, 128, 160, false);
#undef ONK_PK_TMA_GO
    ONK_CHECK_LAUNCH(); count_launch();
    return ONK_OK;
  }
  // CTA shape (ONK_PK_SHAPE, tuning): resident CTAs per SM x threads.  Smaller CTAs pay less per barrier, too small ones
  // too much per tile: 10 x 128 59 us (default), 12 x 128 61, 5 x 256 61, 8 x 128 65, 20 x 64 65, 3 x 512 68, 2 x 512 76,
  // 1 x 1024 94 us on 128^3 cells (C ABI call with preallocated outputs, cold L2)
  static const int shape = []{ const char* e = getenv("ONK_PK_SHAPE"); return e ? atoi(e) : 0; }();
#define ONK_PK_GO(C, NTH) do { \
    constexpr size_t smem = (size_t)(17 * NTH + (17 * NTH) / 16 + 16) * 8; \
    static bool configured = false; \
    if (!configured) { ONK_CUDA(cudaFuncSetAttribute(cell_sum_packed_f64_kernel<C, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; } \
    unsigned grid = grid_for(ncells, NTH, C); \
    if (grid > (unsigned)MAX_PARTIALS) grid = MAX_PARTIALS; \
    cell_sum_packed_f64_kernel<C, NTH><<<grid, NTH, smem, L->stream>>>(values_dev, cell_begin_dev, ncells, cell_sum_dev, L->ws.partials, L->ws.ticket, total_dev); } while (0)
  if (shape == 1) ONK_PK_GO(5, 256); else if (shape == 2) ONK_PK_GO(3, 512); else ONK_PK_GO(10, 128);
#undef ONK_PK_GO
  ONK_CHECK_LAUNCH(); count_launch();
  return ONK_OK;
}

int onk_cell_sum_fields_f64(const void* arena_dev, uint64_t arena_bytes, const uint64_t* cell_off_dev, const uint32_t* cell_size_dev,
                            const uint32_t* cell_cap_dev, uint64_t ncells, uint64_t align, const uint32_t* field_bytes, int nfields,
                            double* cell_sums_dev, double* totals_dev, int lane)
{
  ONK_REQUIRE(nfields >= 1 && nfields <= CELL_ALLF_MAX && field_bytes != nullptr, "onk_cell_sum_fields_f64: 1..%d fields", CELL_ALLF_MAX);
  for (int k = 0; k < nfields; k++) ONK_REQUIRE(field_bytes[k] == 8, "onk_cell_sum_fields_f64: field %d is not fp64", k);
  ONK_REQUIRE(align >= 8 && (align & (align - 1)) == 0, "onk_cell_sum_fields_f64: alignment must be a power of two >= 8 (fp64 columns)");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  if (ncells == 0)
  {
    if (totals_dev) ONK_CUDA(cudaMemsetAsync(totals_dev, 0, sizeof(double) * nfields, L->stream));
    return ONK_OK;
  }
  ONK_REQUIRE(arena_dev && cell_off_dev && cell_size_dev && cell_cap_dev && cell_sums_dev, "onk_cell_sum_fields_f64: null pointer");
  ONK_REQUIRE(((uintptr_t)arena_dev & 7) == 0, "onk_cell_sum_fields_f64: arena is not 8-byte aligned");
  CellAllFieldsParams prm{ (uint32_t)nfields, (uint32_t)(align - 1) };
  const uint64_t nrows_full = arena_bytes / 128;
  static const bool allow_tma = []{ const char* e = getenv("ONK_CELLF_KERNEL"); return !(e && e[0] == 'p'); }();
  if (allow_tma && ((uintptr_t)arena_dev & 15) == 0 && nrows_full > 0 && nrows_full < (1ull << 31))
  {
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_t encode = nullptr;
    if (!encode)
    {
      void* fn = nullptr; cudaDriverEntryPointQueryResult q;
      ONK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
      if (!fn) { set_error("onk_cell_sum_fields_f64: cuTensorMapEncodeTiled entry point not found"); return ONK_ERR_CUDA; }
      encode = (encode_t)fn;
    }
    // 128 threads = 32 cells x 4 fields per tile: 32 x 4 x 1.43 rows = 183 rows on average for Poisson(16) cells in chunks of
    // 16, sigma 11: the 224-row staging buffer (28 KiB) holds all but 1 tile in 10^4 (the rest finishes from global memory)
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = { 16, (cuuint64_t)nrows_full };
    const cuuint64_t gstride[1] = { 128 };
    const cuuint32_t box[2] = { 16, (cuuint32_t)TMA_SUBROWS };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)arena_dev, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("onk_cell_sum_fields_f64: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return ONK_ERR_CUDA; }
#define ONK_CELLF_GO(CT, NTH, BR) do { \
    constexpr size_t smem = (size_t)BR * 128 + 1024; \
    static bool configured = false; \
    if (!configured) { ONK_CUDA(cudaFuncSetAttribute(cell_sum_fields_tma_kernel<CT, NTH, BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = true; } \
    const unsigned cpt = NTH / (unsigned)nfields; \
    unsigned grid = grid_for(ncells, cpt, CT); \
    if (grid > (unsigned)(MAX_PARTIALS / CELL_ALLF_MAX)) grid = MAX_PARTIALS / CELL_ALLF_MAX; \
    cell_sum_fields_tma_kernel<CT, NTH, BR><<<grid, NTH, smem, L->stream>>>(tmap, (const uint8_t*)arena_dev, arena_bytes, nrows_full, cell_off_dev, cell_size_dev, cell_cap_dev, \
                                                                            ncells, prm, cell_sums_dev, L->ws.partials, L->ws.ticket, totals_dev); } while (0)
    // 7 x 128 x 224 6134 / 6821 GB/s of arena stream at 128^3 / 256^3 cells (tools/exp_cell_fields_shape.py); 6 x 128 x 224
    // 6084 / 6865, 6 x 128 x 256 6044 / 6880, 3 x 256 x 448 6182 / 6850, 5 x 128 x 224 5823 / 6493, 4 x 128 x 224 5216 / 5686
    static const int shape = []{ const char* e = getenv("ONK_CELLF_SHAPE"); return e ? atoi(e) : 0; }();      // tuning / comparison only
    if (shape == 1) ONK_CELLF_GO(6, 128, 224);
    else if (shape == 4) ONK_CELLF_GO(3, 256, 448);
    else ONK_CELLF_GO(7, 128, 224);
#undef ONK_CELLF_GO
    ONK_CHECK_LAUNCH(); count_launch();
    return ONK_OK;
  }
  unsigned grid = grid_for(ncells * (uint64_t)nfields, THREADS, 8);
  cell_sum_fields_plain_kernel<<<grid, THREADS, 0, L->stream>>>((const uint8_t*)arena_dev, cell_off_dev, cell_size_dev, cell_cap_dev, ncells, prm, cell_sums_dev);
  ONK_CHECK_LAUNCH(); count_launch();
  if (totals_dev)
  {
    cell_sum_fields_totals_kernel<<<1, THREADS, 0, L->stream>>>(cell_sums_dev, ncells, (unsigned)nfields, totals_dev);
    ONK_CHECK_LAUNCH(); count_launch();
  }
  return ONK_OK;
}

} // extern "C"
