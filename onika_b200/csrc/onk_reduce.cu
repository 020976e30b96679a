// onk_reduce.cu -- fp64 min / max / sum reductions for sm_100a.
//
// Reference behaviour: cpu_compute, tests/gpu_reduce_min_fp64.cu:46-55 (the GPU half is a TODO there, :109-110);
// the unused shared-memory trees block_reduce_add/min/max, include/onika/cuda/cuda_reduction.h:30-85; the
// atomic return-data accumulation idiom.
//
// B200 design (HBM-bound, 8 B/element read, no reuse):
//   * one pass over the data with 256-bit non-coherent streaming loads (LDG.E.256, L1 no-allocate, L2 evict-first),
//     UNROLL independent loads in flight per thread, a grid of CTAS_PER_SM x 148 persistent CTAs walking
//     32 KiB tiles grid-stride, so every SM keeps >= 64 KiB of loads in flight;
//   * per-thread accumulators -> butterfly warp shuffle -> 8 warp partials in shared memory -> one partial per CTA;
//   * two-pass finish inside the same launch: the last CTA to retire (atomic ticket, self-resetting) folds the
//     per-CTA partials in a fixed order and writes the result.  No atomics on data, no second launch, and the
//     summation tree depends only on (n, grid) => run-to-run reproducible sums.
#include "onk_internal.cuh"

namespace onk
{
  constexpr int RED_UNROLL = 4;                                  // 4 x 32 B per thread per tile (tools/tune.cu sweep: every shape lands within 2 % of 7.3 TB/s)
  constexpr unsigned RED_TILE = THREADS * RED_UNROLL * 4;        // doubles per tile (4096 = 32 KiB)
  constexpr unsigned RED_CTAS_PER_SM = 4;

  template<int OP>
  __global__ void __launch_bounds__(THREADS, RED_CTAS_PER_SM)
  reduce_f64_kernel(const double* __restrict__ in, uint64_t n, uint64_t head, uint64_t ntiles,
                    double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ out)
  {
    using R = Red<OP>;
    double acc[RED_UNROLL * 4];
#pragma unroll
    for (int i = 0; i < RED_UNROLL * 4; i++) acc[i] = R::identity();

    // body: full tiles of 32-byte aligned vectors starting at in+head
    const double* body = in + head;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
    {
      const double* p = body + t * RED_TILE + threadIdx.x * 4;
      d4 v[RED_UNROLL];
#pragma unroll
      for (int u = 0; u < RED_UNROLL; u++) v[u] = ld_stream_ro(p + u * (THREADS * 4));
#pragma unroll
      for (int u = 0; u < RED_UNROLL; u++)
      {
        acc[u * 4 + 0] = R::apply(acc[u * 4 + 0], v[u].a);
        acc[u * 4 + 1] = R::apply(acc[u * 4 + 1], v[u].b);
        acc[u * 4 + 2] = R::apply(acc[u * 4 + 2], v[u].c);
        acc[u * 4 + 3] = R::apply(acc[u * 4 + 3], v[u].d);
      }
    }

    // edges (at most head + one tile of elements): scalar, grid-strided over the whole grid
    {
      const uint64_t tail_begin = head + ntiles * RED_TILE;
      const uint64_t nedge = head + (n - tail_begin);
      for (uint64_t e = (uint64_t)blockIdx.x * THREADS + threadIdx.x; e < nedge; e += (uint64_t)gridDim.x * THREADS)
      {
        const uint64_t i = (e < head) ? e : (tail_begin + (e - head));
        acc[0] = R::apply(acc[0], ld_stream_ro1(in + i));
      }
    }

    // fixed-order fold of the 16 per-thread accumulators
#pragma unroll
    for (int s = RED_UNROLL * 2; s > 0; s >>= 1)
#pragma unroll
      for (int i = 0; i < s; i++) acc[i] = R::apply(acc[i], acc[i + s]);

    const double blk = block_reduce<OP, THREADS>(acc[0]);

    __shared__ bool is_last;
    if (threadIdx.x == 0)
    {
      partials[blockIdx.x] = blk;
      __threadfence();
      const unsigned done = atomicAdd(ticket, 1u);
      is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last)
    {
      __threadfence();
      double r = R::identity();
      for (unsigned i = threadIdx.x; i < gridDim.x; i += THREADS) r = R::apply(r, __ldcg(partials + i));
      r = block_reduce<OP, THREADS>(r);
      if (threadIdx.x == 0) { *out = r; *ticket = 0u; }
    }
  }

  template<int OP>
  __global__ void __launch_bounds__(THREADS)
  combine_f64_kernel(const double* __restrict__ partials, unsigned count, double* __restrict__ out)
  {
    using R = Red<OP>;
    // strictly sequential in rank order: bit-reproducible irrespective of the communication library's tree
    if (threadIdx.x == 0)
    {
      double r = R::identity();
      for (unsigned i = 0; i < count; i++) r = R::apply(r, partials[i]);
      *out = r;
    }
  }

  template<int OP>
  static int launch_reduce(const double* in, uint64_t n, double* out, Lane* L)
  {
    // 32-byte alignment of the vector body; a pointer that is not even 8-byte aligned is rejected by the caller
    uint64_t head = ((32 - ((uintptr_t)in & 31)) & 31) / 8;
    if (head > n) head = n;
    const uint64_t ntiles = (n - head) / RED_TILE;
    const uint64_t edge = n - ntiles * RED_TILE;
    unsigned grid = grid_for(ntiles, 1, RED_CTAS_PER_SM);
    if (ntiles == 0) grid = (unsigned)((edge + THREADS - 1) / THREADS ? (edge + THREADS - 1) / THREADS : 1);
    if (grid > (unsigned)MAX_PARTIALS) grid = MAX_PARTIALS;
    reduce_f64_kernel<OP><<<grid, THREADS, 0, L->stream>>>(in, n, head, ntiles, L->ws.partials, L->ws.ticket, out);
    ONK_CHECK_LAUNCH();
    count_launch();
    return ONK_OK;
  }
}

using namespace onk;

extern "C" {

int onk_reduce_f64(int op, const double* in_dev, uint64_t n, double* out_dev, int lane)
{
  ONK_REQUIRE(out_dev != nullptr, "onk_reduce_f64: null output");
  ONK_REQUIRE(n == 0 || in_dev != nullptr, "onk_reduce_f64: null input");
  ONK_REQUIRE(((uintptr_t)in_dev & 7) == 0, "onk_reduce_f64: input is not 8-byte aligned");
  ONK_REQUIRE(op == ONK_OP_MIN || op == ONK_OP_MAX || op == ONK_OP_SUM, "onk_reduce_f64: unknown op %d", op);
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  switch (op)
  {
    case ONK_OP_MIN: return launch_reduce<ONK_OP_MIN>(in_dev, n, out_dev, L);
    case ONK_OP_MAX: return launch_reduce<ONK_OP_MAX>(in_dev, n, out_dev, L);
    default:         return launch_reduce<ONK_OP_SUM>(in_dev, n, out_dev, L);
  }
}

int onk_combine_f64(int op, const double* partials_dev, uint32_t count, double* out_dev, int lane)
{
  ONK_REQUIRE(out_dev != nullptr && (count == 0 || partials_dev != nullptr), "onk_combine_f64: null argument");
  ONK_REQUIRE(op == ONK_OP_MIN || op == ONK_OP_MAX || op == ONK_OP_SUM, "onk_combine_f64: unknown op %d", op);
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  switch (op)
  {
    case ONK_OP_MIN: combine_f64_kernel<ONK_OP_MIN><<<1, 32, 0, L->stream>>>(partials_dev, count, out_dev); break;
    case ONK_OP_MAX: combine_f64_kernel<ONK_OP_MAX><<<1, 32, 0, L->stream>>>(partials_dev, count, out_dev); break;
    default:         combine_f64_kernel<ONK_OP_SUM><<<1, 32, 0, L->stream>>>(partials_dev, count, out_dev); break;
  }
  ONK_CHECK_LAUNCH();
  count_launch();
  return ONK_OK;
}

} // extern "C"
