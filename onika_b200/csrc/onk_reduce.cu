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

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

namespace onk
{
  constexpr int RED_UNROLL = 4;                                  // 4 x 32 B per thread per tile (tools/tune.cu sweep: every shape lands within 2 % of 7.3 TB/s)
  constexpr unsigned RED_TILE = THREADS * RED_UNROLL * 4;        // doubles per tile (4096 = 32 KiB)
  constexpr unsigned RED_CTAS_PER_SM = 4;

  // ---- cross-GPU exchange (fused into the finish of the reduction kernel) ----
  // One buffer per rank, mapped into every peer process.  Values travel in FLAGGED PACKETS (the "LL" idea of collective
  // libraries): a double is sent as two 8-byte words, each carrying 32 payload bits and the 32-bit call epoch, written with
  // one 16-byte peer store.  An 8-byte word is never torn, so a reader that sees the epoch in both words holds the value:
  // no release fence on the producer, no acquire on the consumer, one NVLink round per exchange.  The epoch is a call
  // counter kept in the rank's own buffer (device-resident, so a captured CUDA graph replays correctly) and slots are
  // double-buffered by its parity: a rank can be at most one call ahead of the slowest peer.
  struct alignas(16) XchgPacket { unsigned long long lo, hi; };
  struct XchgBuffer
  {
    XchgPacket         slot[2][ONK_XCHG_MAX_WORLD][ONK_XCHG_MAX_VALUES];   // [parity][source rank][value]
    unsigned long long epoch;                      // calls completed on this exchange (only this rank's kernels touch it)
  };
  static_assert(sizeof(XchgBuffer) <= ONK_XCHG_BYTES, "exchange buffer layout");
  struct XchgParams
  {
    XchgBuffer* peer[ONK_XCHG_MAX_WORLD];          // peer[r] = rank r's buffer as seen from this device (peer[rank] is local)
    unsigned long long* error_host;                // mapped pinned word: epoch of the first call in which a peer did not show up
    int rank, world;
  };

  constexpr long long XCHG_TIMEOUT_CLOCKS = 20000000000ll;          // ~10 s: a peer is missing; fail instead of hanging the device

  __device__ __forceinline__ void xchg_send(XchgPacket* dst, double v, unsigned epoch32)
  {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), e = (unsigned long long)epoch32 << 32;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" :: "l"(dst), "l"((b & 0xffffffffull) | e), "l"((b >> 32) | e) : "memory");
  }
  // spins without backing off first (the peers are usually a few microseconds apart), then sleeps between polls
  __device__ __forceinline__ bool xchg_recv(const XchgPacket* src, unsigned epoch32, double* v)
  {
    unsigned long long lo, hi;
    long long t0 = 0;
    for (unsigned spins = 0;; spins++)
    {
      asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
      if ((unsigned)(lo >> 32) == epoch32 && (unsigned)(hi >> 32) == epoch32) break;
      if (spins == 2048u) t0 = clock64();
      if (spins > 2048u)
      {
        __nanosleep(128);
        if (clock64() - t0 > XCHG_TIMEOUT_CLOCKS) { *v = __longlong_as_double(0x7ff8000000000000ll); return false; }
      }
    }
    *v = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    return true;
  }
  __device__ __forceinline__ void xchg_fail(const XchgParams& xp, unsigned long long epoch)
  {
    if (*reinterpret_cast<volatile unsigned long long*>(xp.error_host) == 0ull) *reinterpret_cast<volatile unsigned long long*>(xp.error_host) = epoch;
    __threadfence_system();
  }
  // next epoch of this rank's exchange; call from ONE thread
  __device__ __forceinline__ unsigned long long xchg_next_epoch(const XchgParams& xp)
  {
    XchgBuffer* mine = xp.peer[xp.rank];
    const unsigned long long e = mine->epoch + 1ull;
    mine->epoch = e;
    return e;
  }

  // called by ONE CTA with the rank-local result in thread 0; returns the global result in thread 0.  A peer that does not
  // show up makes the result NaN for every operator and records the epoch in the exchange's host-visible error word.
  template<int OP>
  __device__ __forceinline__ double xchg_combine(double local, const XchgParams& xp)
  {
    using R = Red<OP>;
    __shared__ double s_local;
    __shared__ unsigned long long s_epoch;
    __shared__ double s_vals[ONK_XCHG_MAX_WORLD];
    __shared__ int s_failed;
    if (threadIdx.x == 0) { s_local = local; s_epoch = xchg_next_epoch(xp); s_failed = 0; }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const int par = (int)(epoch & 1ull);
    if ((int)threadIdx.x < xp.world)
    {
      const int peer = (int)threadIdx.x;
      xchg_send(&xp.peer[peer]->slot[par][xp.rank][0], s_local, (unsigned)epoch);            // my value into rank `peer`'s buffer
      double v;
      if (!xchg_recv(&xp.peer[xp.rank]->slot[par][peer][0], (unsigned)epoch, &v)) { s_failed = 1; xchg_fail(xp, epoch); }
      s_vals[peer] = v;
    }
    __syncthreads();
    double r = R::identity();
    if (threadIdx.x == 0)
    {
      for (int i = 0; i < xp.world; i++) r = R::apply(r, s_vals[i]);                          // rank order
      if (s_failed) r = __longlong_as_double(0x7ff8000000000000ll);
    }
    return r;
  }

  // block reduction of the per-thread values, one partial per CTA, two-pass finish inside the same launch: the last CTA to
  // retire (atomic ticket, self-resetting) folds the per-CTA partials in index order and, for XCHG, combines across ranks
  template<int OP, bool XCHG>
  __device__ __forceinline__ void reduce_finish(double thread_value, double* __restrict__ partials, unsigned* __restrict__ ticket,
                                                double* __restrict__ out, const XchgParams& xp)
  {
    using R = Red<OP>;
    const double blk = block_reduce<OP, THREADS>(thread_value);
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
      if constexpr (XCHG) r = xchg_combine<OP>(r, xp);
      if (threadIdx.x == 0) { *out = r; *ticket = 0u; }
    }
  }

  // What a reduction does before it may touch memory its predecessor on the lane wrote (programmatic dependent launch, see
  // launch_pdl).  Launched early -- the predecessor was a reduction that released its dependents -- the CTAs become resident
  // while the predecessor's last CTA is still folding partials and exchanging with the other GPUs; in that window they ask the
  // L2 for the first tiles they are going to read (a prefetch has no semantics: whatever the predecessor still writes lands in
  // the L2 as well, and the loads come after the wait), then wait for the predecessor's completion and memory flush, then
  // release their own dependents (after the wait: at most ONE successor is in flight behind a running reduction).
  constexpr unsigned PDL_RELEASE = 1u;              // release this kernel's dependents once it passed its own wait
  constexpr unsigned PDL_PREFETCH_SHIFT = 8;        // bits 8..15: tiles per CTA to prefetch into L2 before the wait
  __device__ __forceinline__ void pdl_prologue(const double* body, uint64_t ntiles, unsigned tile_bytes, unsigned flags)
  {
    const unsigned pf = (flags >> PDL_PREFETCH_SHIFT) & 0xffu;
    if (pf != 0 && threadIdx.x == 0)                   // bulk requests are issued from uniform registers: one thread, one tile after the other
      for (uint64_t k = 0, t = blockIdx.x; k < pf && t < ntiles; k++, t += gridDim.x)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(reinterpret_cast<const char*>(body) + t * tile_bytes), "r"(tile_bytes) : "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");          // no-op unless launched as a programmatic dependent
    if (flags & PDL_RELEASE) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }

  template<int OP, bool XCHG>
  __global__ void __launch_bounds__(THREADS, RED_CTAS_PER_SM)
  reduce_f64_kernel(const double* __restrict__ in, uint64_t n, uint64_t head, uint64_t ntiles,
                    double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ out,
                    const __grid_constant__ XchgParams xp, unsigned pdl_flags)
  {
    using R = Red<OP>;
    pdl_prologue(in + head, ntiles, RED_TILE * 8u, pdl_flags);

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

    reduce_finish<OP, XCHG>(acc[0], partials, ticket, out, xp);
  }

  // ---- large inputs: the body goes global -> shared with TMA bulk copies ----
  // One elected thread streams 64 KiB tiles into a 2-stage ring (cp.async.bulk + mbarrier complete_tx), all 256 threads fold
  // the landed tile from shared memory (conflict-free LDS.128), one CTA per SM.  128 KiB in flight per SM as two large
  // requests measured 7.28-7.32 TB/s on 2 GiB against 7.11-7.18 TB/s for the register-staged loop above in the same run
  // (tools/tune.cu r; deeper rings or 96-112 KiB tiles fall back below 7.0).  Edges and the finish are shared with it.
  constexpr unsigned TMA_TILE_BYTES = 65536;
  constexpr unsigned TMA_TILE = TMA_TILE_BYTES / 8;             // doubles
  constexpr int      TMA_STAGES = 2;
  constexpr size_t   TMA_SMEM = (size_t)TMA_STAGES * TMA_TILE_BYTES + 64;

  template<int OP, bool XCHG>
  __global__ void __launch_bounds__(THREADS, 1)
  reduce_f64_tma_kernel(const double* __restrict__ in, uint64_t n, uint64_t head, uint64_t ntiles,
                        double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ out,
                        const __grid_constant__ XchgParams xp, unsigned pdl_flags)
  {
    using R = Red<OP>;
    pdl_prologue(in + head, ntiles, TMA_TILE_BYTES, pdl_flags);

    extern __shared__ __align__(128) unsigned char tsm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + (size_t)TMA_STAGES * TMA_TILE_BYTES);
    const unsigned tid = threadIdx.x;
    if (tid == 0)
    {
      for (int s = 0; s < TMA_STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((unsigned)__cvta_generic_to_shared(bars + s)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const double* body = in + head;                            // 32-byte aligned
    const uint64_t mine = (blockIdx.x < ntiles) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](uint64_t k)
    {
      const int s = (int)(k % TMA_STAGES);
      const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + s);
      const unsigned dst = (unsigned)__cvta_generic_to_shared(tsm + (size_t)s * TMA_TILE_BYTES);
      const uint64_t tile = blockIdx.x + k * gridDim.x;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(TMA_TILE_BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(dst), "l"(body + tile * TMA_TILE), "r"(TMA_TILE_BYTES), "r"(bar) : "memory");
    };
    if (tid == 0) for (uint64_t k = 0; k < (uint64_t)TMA_STAGES && k < mine; k++) issue(k);

    double acc[4] = { R::identity(), R::identity(), R::identity(), R::identity() };
    for (uint64_t k = 0; k < mine; k++)
    {
      const int s = (int)(k % TMA_STAGES);
      const unsigned phase = (unsigned)((k / TMA_STAGES) & 1);
      const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + s);
      unsigned landed = 0;
      while (!landed)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(landed) : "r"(bar), "r"(phase) : "memory");
      const unsigned base = (unsigned)__cvta_generic_to_shared(tsm + (size_t)s * TMA_TILE_BYTES) + tid * 32u;
#pragma unroll
      for (unsigned i = 0; i < TMA_TILE_BYTES / (THREADS * 32u); i++)          // fixed order: reproducible sums
      {
        double x0, x1, x2, x3;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(base + i * (THREADS * 32u)) : "memory");
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x2), "=d"(x3) : "r"(base + i * (THREADS * 32u) + 16u) : "memory");
        acc[0] = R::apply(acc[0], x0); acc[1] = R::apply(acc[1], x1); acc[2] = R::apply(acc[2], x2); acc[3] = R::apply(acc[3], x3);
      }
      __syncthreads();                                                          // the stage may be overwritten
      if (tid == 0 && k + TMA_STAGES < mine) issue(k + TMA_STAGES);
    }

    // edges (head + what is left after the last full tile): scalar, grid-strided over the whole grid
    {
      const uint64_t tail_begin = head + ntiles * TMA_TILE;
      const uint64_t nedge = head + (n - tail_begin);
      for (uint64_t e = (uint64_t)blockIdx.x * THREADS + tid; e < nedge; e += (uint64_t)gridDim.x * THREADS)
      {
        const uint64_t i = (e < head) ? e : (tail_begin + (e - head));
        acc[0] = R::apply(acc[0], ld_stream_ro1(in + i));
      }
    }
    acc[0] = R::apply(acc[0], acc[2]); acc[1] = R::apply(acc[1], acc[3]);
    acc[0] = R::apply(acc[0], acc[1]);
    reduce_finish<OP, XCHG>(acc[0], partials, ticket, out, xp);
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

  // Programmatic dependent launch (ONK_PDL=0 turns it off): a reduction may be SCHEDULED while its predecessor on the lane
  // drains -- it waits for the predecessor's completion and memory flush in griddepcontrol.wait, its first instruction, before it
  // touches anything -- so part of the launch latency between back-to-back kernels of a lane is hidden.  A/B on one 2-GPU box
  // (bench.py --gpus 2, two runs each): weak step 0.2997 -> 0.2987 ms, 2^28 in total 0.1562 -> 0.1536 ms, two-lane C5 sequence
  // unchanged within noise.  By default the kernels do NOT release their dependents early: an early release parks the
  // dependent's CTAs on the SMs while the predecessor drains, which takes slots from the kernels of OTHER lanes.  ONK_PDL_EARLY=2
  // turns it on for the register-staged kernel, with an L2 prefetch of the dependent's first tiles while it waits (pdl_prologue):
  // one GPU, back to back, 128 MiB 22.5 -> 20.9 us, 256 MiB 42.0 -> 40.7 us; 8 GPUs, 2^28 doubles in total (tools/exp_strong.py,
  // two runs each): 46.2 / 46.4 -> 45.5 / 45.4 us per step (6.34x -> 6.47x), but the two-lane C5 sequence 0.5298 / 0.5310 ->
  // 0.5366 / 0.5333 ms: a 2 % gain on single-lane chains of medium reductions against 0.4-1.3 % on multi-lane work, hence opt-in.
  template<class... KArgs, class... Args>
  static cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args... args)
  {
    static const bool pdl = []{ const char* e = getenv("ONK_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (pdl)
    {
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  }

  template<int OP>
  static int launch_reduce(const double* in, uint64_t n, double* out, Lane* L, const XchgParams* xp = nullptr)
  {
    // 32-byte alignment of the vector body; a pointer that is not even 8-byte aligned is rejected by the caller
    uint64_t head = ((32 - ((uintptr_t)in & 31)) & 31) / 8;
    if (head > n) head = n;
    // large inputs take the TMA-staged body (ONK_RED_KERNEL=ldg forces the register-staged kernel: tuning / comparison)
    // ONK_RED_KERNEL (tuning / tests): "ldg" never takes the TMA-staged body, "tma" takes it from 2 tiles per SM on
    static const int red_mode = []{ const char* e = getenv("ONK_RED_KERNEL"); return (e && e[0] == 'l') ? 1 : (e && e[0] == 't') ? 2 : 0; }();
    const bool allow_tma = red_mode != 1;
    // the TMA-staged body pays off on long streams (1 GiB: 151.6 vs 162.7 us); up to 512 MiB the register-staged kernel with
    // its 4 CTAs per SM ramps up and drains a little faster (256 MiB back to back: 41.7 vs 43.1 us, 128 MiB: 20.7 vs 25.0 us;
    // tools/exp_reduce_small.py) -- the per-GPU shards of a strong-scaling run are in that range
    const uint64_t tma_tiles = (n - head) / TMA_TILE;
    // ONK_PDL_EARLY (A/B knob): 0 = dependents start when the predecessor has completed, 1 = released early, 2 = released early
    // and prefetching ONK_PDL_PREFETCH_MB (default 32) of their input into L2 while they wait; 1 and 2 concern the register-staged
    // kernel (shards below 768 MiB, where the gap between two kernels is a visible share of the step), 3 = 2 for both kernels
    static const int early = []{ const char* e = getenv("ONK_PDL_EARLY"); return e ? atoi(e) : 0; }();
    static const unsigned prefetch_mb = []{ const char* e = getenv("ONK_PDL_PREFETCH_MB"); const int v = e ? atoi(e) : 32; return (unsigned)(v < 0 ? 0 : v); }();
    auto flags_for = [&](unsigned grid, unsigned tile_bytes) -> unsigned
    {
      if (early <= 0) return 0u;
      unsigned pf = 0;
      if (early >= 2 && grid > 0) { pf = (unsigned)((((uint64_t)prefetch_mb << 20) + (uint64_t)grid * tile_bytes - 1) / ((uint64_t)grid * tile_bytes)); if (pf > 64) pf = 64; }
      return PDL_RELEASE | (pf << PDL_PREFETCH_SHIFT);
    };
    if (allow_tma && tma_tiles >= (red_mode == 2 ? 2ull * (uint64_t)rt().sm_count : 12288ull))
    {
      static bool configured = false;
      if (!configured)
      {
        ONK_CUDA(cudaFuncSetAttribute(reduce_f64_tma_kernel<OP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
        ONK_CUDA(cudaFuncSetAttribute(reduce_f64_tma_kernel<OP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
        configured = true;
      }
      const unsigned g = (unsigned)rt().sm_count;
      const unsigned pdl_flags = (early >= 3) ? flags_for(g, TMA_TILE_BYTES) : 0u;
      if (xp != nullptr && xp->world > 1)
        ONK_CUDA(launch_pdl(reduce_f64_tma_kernel<OP, true>, g, THREADS, TMA_SMEM, L->stream, in, n, head, tma_tiles, L->ws.partials, L->ws.ticket, out, *xp, pdl_flags));
      else
        ONK_CUDA(launch_pdl(reduce_f64_tma_kernel<OP, false>, g, THREADS, TMA_SMEM, L->stream, in, n, head, tma_tiles, L->ws.partials, L->ws.ticket, out, XchgParams{}, pdl_flags));
      ONK_CHECK_LAUNCH();
      count_launch();
      return ONK_OK;
    }
    const uint64_t ntiles = (n - head) / RED_TILE;
    const uint64_t edge = n - ntiles * RED_TILE;
    static const unsigned ctas_per_sm = []{ const char* e = getenv("ONK_RED_CTAS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= 8) ? (unsigned)v : RED_CTAS_PER_SM; }();   // tuning knob
    unsigned grid = grid_for(ntiles, 1, ctas_per_sm);
    if (ntiles == 0) grid = (unsigned)((edge + THREADS - 1) / THREADS ? (edge + THREADS - 1) / THREADS : 1);
    if (grid > (unsigned)MAX_PARTIALS) grid = MAX_PARTIALS;
    const unsigned pdl_flags = flags_for(grid, RED_TILE * 8u);
    if (xp != nullptr && xp->world > 1)
      ONK_CUDA(launch_pdl(reduce_f64_kernel<OP, true>, grid, THREADS, 0, L->stream, in, n, head, ntiles, L->ws.partials, L->ws.ticket, out, *xp, pdl_flags));
    else
      ONK_CUDA(launch_pdl(reduce_f64_kernel<OP, false>, grid, THREADS, 0, L->stream, in, n, head, ntiles, L->ws.partials, L->ws.ticket, out, XchgParams{}, pdl_flags));
    ONK_CHECK_LAUNCH();
    count_launch();
    return ONK_OK;
  }
}

using namespace onk;

namespace onk
{
  // all-reduce of `count` <= ONK_XCHG_MAX_VALUES doubles held by every rank, in place, same operator for all values:
  // thread (p,k) sends this rank's value k into rank p's buffer and receives rank p's value k; values are folded in rank order
  template<int OP>
  __global__ void __launch_bounds__(ONK_XCHG_MAX_WORLD * ONK_XCHG_MAX_VALUES)
  xchg_allreduce_kernel(double* __restrict__ values, int count, const __grid_constant__ XchgParams xp)
  {
    using R = Red<OP>;
    __shared__ double s_vals[ONK_XCHG_MAX_WORLD][ONK_XCHG_MAX_VALUES];
    __shared__ unsigned long long s_epoch;
    __shared__ int s_failed;
    if (threadIdx.x == 0) { s_epoch = xchg_next_epoch(xp); s_failed = 0; }
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const int par = (int)(epoch & 1ull);
    const int peer = (int)threadIdx.x / ONK_XCHG_MAX_VALUES, k = (int)threadIdx.x % ONK_XCHG_MAX_VALUES;
    if (peer < xp.world && k < count)
    {
      xchg_send(&xp.peer[peer]->slot[par][xp.rank][k], values[k], (unsigned)epoch);
      double v;
      if (!xchg_recv(&xp.peer[xp.rank]->slot[par][peer][k], (unsigned)epoch, &v)) { s_failed = 1; xchg_fail(xp, epoch); }
      s_vals[peer][k] = v;
    }
    __syncthreads();
    if ((int)threadIdx.x < count)
    {
      double r = R::identity();
      for (int p = 0; p < xp.world; p++) r = R::apply(r, s_vals[p][threadIdx.x]);
      values[threadIdx.x] = s_failed ? __longlong_as_double(0x7ff8000000000000ll) : r;
    }
  }

  // device-side barrier between the ranks' lanes: everything queued before it on every rank's lane, peer stores included, is
  // complete and visible when it returns on any rank.  The earlier kernels of the lane have completed when this one starts;
  // the system-scope fences order their (peer) writes before the packet and the packet before what follows.
  __global__ void xchg_barrier_kernel(const __grid_constant__ XchgParams xp)
  {
    __threadfence_system();
    (void) xchg_combine<ONK_OP_MIN>(0.0, xp);
    __threadfence_system();
  }
}

struct onk_xchg
{
  XchgBuffer* local = nullptr;                   // device transport: cudaMalloc'ed; host transport: device view of the local shared-memory segment
  XchgParams  params{};
  bool        connected = false;
  int         transport = ONK_XCHG_TRANSPORT_DEVICE;
  void*       opened[ONK_XCHG_MAX_WORLD] = {};   // device transport: IPC mappings of the peers' buffers
  unsigned long long* error_host = nullptr;      // pinned + mapped: written by the kernels on a timeout, read by the host
  // host transport: POSIX shared-memory segments (own + peers'), registered with the CUDA runtime
  void*       shm_map[ONK_XCHG_MAX_WORLD] = {};
  char        shm_name[48] = "";
};

namespace onk
{
  static const char XCHG_HOST_MAGIC[4] = { 'O', 'N', 'K', 'H' };

  // maps (and, for the owner, creates) a shared-memory segment and registers it with the CUDA runtime; returns the device view
  static int xchg_map_segment(const char* name, bool create, void** host_map, XchgBuffer** dev_view)
  {
    const int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) { set_error("exchange (host transport): shm_open(%s) failed", name); return ONK_ERR_STATE; }
    if (create && ftruncate(fd, ONK_XCHG_BYTES) != 0) { close(fd); shm_unlink(name); set_error("exchange (host transport): ftruncate failed"); return ONK_ERR_NOMEM; }
    void* p = mmap(nullptr, ONK_XCHG_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { if (create) shm_unlink(name); set_error("exchange (host transport): mmap failed"); return ONK_ERR_NOMEM; }
    if (create) memset(p, 0, ONK_XCHG_BYTES);
    cudaError_t e = cudaHostRegister(p, ONK_XCHG_BYTES, cudaHostRegisterMapped | cudaHostRegisterPortable);
    void* d = nullptr;
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&d, p, 0);
    if (e != cudaSuccess) { munmap(p, ONK_XCHG_BYTES); if (create) shm_unlink(name); ONK_CUDA(e); }
    *host_map = p; *dev_view = (XchgBuffer*)d;
    return ONK_OK;
  }
}

namespace onk
{
  // sticky failure: once a peer did not show up in a call, every later call on the exchange is refused
  static int xchg_usable(onk_xchg* x, const char* who)
  {
    if (x == nullptr || !x->connected) { set_error("%s: exchange is not connected", who); return ONK_ERR_ARG; }
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(x->error_host);
    if (e != 0ull) { set_error("%s: the exchange failed in call %llu (a peer rank did not show up within the timeout)", who, e); return ONK_ERR_STATE; }
    return ONK_OK;
  }
}

extern "C" {

int onk_xchg_create_ex(int rank, int world, int transport, onk_xchg** x)
{
  ONK_REQUIRE(x != nullptr, "onk_xchg_create: null argument");
  ONK_REQUIRE(world >= 1 && world <= ONK_XCHG_MAX_WORLD && rank >= 0 && rank < world, "onk_xchg_create: bad rank/world %d/%d", rank, world);
  ONK_REQUIRE(transport == ONK_XCHG_TRANSPORT_DEVICE || transport == ONK_XCHG_TRANSPORT_HOST, "onk_xchg_create: unknown transport %d", transport);
  if (int rc = require_init()) return rc;
  onk_xchg* h = new onk_xchg();
  h->transport = transport;
  cudaError_t e = cudaHostAlloc((void**)&h->error_host, 64, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) { delete h; ONK_CUDA(e); }
  *h->error_host = 0ull;
  if (transport == ONK_XCHG_TRANSPORT_DEVICE)
  {
    e = cudaMalloc((void**)&h->local, ONK_XCHG_BYTES);      // plain cudaMalloc: exportable through CUDA IPC
    if (e == cudaSuccess) e = cudaMemset(h->local, 0, ONK_XCHG_BYTES);
    if (e != cudaSuccess) { if (h->local) cudaFree(h->local); cudaFreeHost(h->error_host); delete h; ONK_CUDA(e); }
  }
  else
  {
    static std::atomic<unsigned> seq{0};
    snprintf(h->shm_name, sizeof(h->shm_name), "/onkx_%d_%u_%d", (int)getpid(), seq.fetch_add(1), rank);
    if (int rc = xchg_map_segment(h->shm_name, true, &h->shm_map[rank], &h->local)) { cudaFreeHost(h->error_host); delete h; return rc; }
  }
  h->params.rank = rank; h->params.world = world;
  h->params.peer[rank] = h->local;
  h->params.error_host = h->error_host;          // unified addressing: the mapped host pointer is valid on the device
  h->connected = (world == 1);
  *x = h;
  return ONK_OK;
}

int onk_xchg_create(int rank, int world, onk_xchg** x)
{
  // ONK_XCHG_TRANSPORT=host selects the shared-memory transport for every exchange of the process (boxes without peer access)
  static const int transport = []{ const char* e = getenv("ONK_XCHG_TRANSPORT"); return (e && e[0] == 'h') ? ONK_XCHG_TRANSPORT_HOST : ONK_XCHG_TRANSPORT_DEVICE; }();
  return onk_xchg_create_ex(rank, world, transport, x);
}

int onk_xchg_export(onk_xchg* x, unsigned char handle[ONK_IPC_HANDLE_BYTES])
{
  ONK_REQUIRE(x != nullptr && handle != nullptr, "onk_xchg_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == ONK_IPC_HANDLE_BYTES, "CUDA IPC handle size");
  if (x->transport == ONK_XCHG_TRANSPORT_HOST)
  {
    memset(handle, 0, ONK_IPC_HANDLE_BYTES);
    memcpy(handle, XCHG_HOST_MAGIC, 4);
    strncpy((char*)handle + 4, x->shm_name, ONK_IPC_HANDLE_BYTES - 5);
    return ONK_OK;
  }
  cudaIpcMemHandle_t h;
  ONK_CUDA(cudaIpcGetMemHandle(&h, x->local));
  memcpy(handle, &h, sizeof(h));
  return ONK_OK;
}

int onk_xchg_connect(onk_xchg* x, const unsigned char* handles)
{
  ONK_REQUIRE(x != nullptr && handles != nullptr, "onk_xchg_connect: null argument");
  for (int r = 0; r < x->params.world; r++)
  {
    if (r == x->params.rank) continue;
    const unsigned char* hr = handles + (size_t)r * ONK_IPC_HANDLE_BYTES;
    if (x->transport == ONK_XCHG_TRANSPORT_HOST)
    {
      ONK_REQUIRE(memcmp(hr, XCHG_HOST_MAGIC, 4) == 0, "onk_xchg_connect: rank %d did not export a host-transport handle (all ranks must use the same transport)", r);
      char name[ONK_IPC_HANDLE_BYTES]; memcpy(name, hr + 4, ONK_IPC_HANDLE_BYTES - 4); name[ONK_IPC_HANDLE_BYTES - 5] = 0;
      XchgBuffer* d = nullptr;
      if (int rc = xchg_map_segment(name, false, &x->shm_map[r], &d)) return rc;
      x->params.peer[r] = d;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, hr, sizeof(h));
    void* p = nullptr;
    ONK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->opened[r] = p;
    x->params.peer[r] = (XchgBuffer*)p;
  }
  x->connected = true;
  return ONK_OK;
}

int onk_xchg_destroy(onk_xchg* x)
{
  if (!x) return ONK_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < ONK_XCHG_MAX_WORLD; r++) if (x->opened[r]) cudaIpcCloseMemHandle(x->opened[r]);
  for (int r = 0; r < ONK_XCHG_MAX_WORLD; r++) if (x->shm_map[r]) { cudaHostUnregister(x->shm_map[r]); munmap(x->shm_map[r], ONK_XCHG_BYTES); }
  if (x->transport == ONK_XCHG_TRANSPORT_HOST) { if (x->shm_name[0]) shm_unlink(x->shm_name); }
  else if (x->local) cudaFree(x->local);
  if (x->error_host) cudaFreeHost(x->error_host);
  delete x;
  return ONK_OK;
}

int onk_xchg_status(onk_xchg* x, uint64_t* failed_call)
{
  ONK_REQUIRE(x != nullptr && failed_call != nullptr, "onk_xchg_status: null argument");
  *failed_call = *reinterpret_cast<volatile unsigned long long*>(x->error_host);
  return ONK_OK;
}

int onk_xchg_info(onk_xchg* x, int* rank, int* world)
{
  ONK_REQUIRE(x != nullptr, "onk_xchg_info: null exchange");
  if (rank) *rank = x->params.rank;
  if (world) *world = x->params.world;
  return ONK_OK;
}

int onk_xchg_barrier(onk_xchg* x, int lane)
{
  if (int rc = xchg_usable(x, "onk_xchg_barrier")) return rc;
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  xchg_barrier_kernel<<<1, 32, 0, L->stream>>>(x->params);
  ONK_CHECK_LAUNCH();
  count_launch();
  return ONK_OK;
}

int onk_xchg_allreduce_f64(onk_xchg* x, int op, double* values_dev, int count, int lane)
{
  if (int rc = xchg_usable(x, "onk_xchg_allreduce_f64")) return rc;
  ONK_REQUIRE(values_dev != nullptr && count >= 1 && count <= ONK_XCHG_MAX_VALUES, "onk_xchg_allreduce_f64: 1..%d values", ONK_XCHG_MAX_VALUES);
  ONK_REQUIRE(op == ONK_OP_MIN || op == ONK_OP_MAX || op == ONK_OP_SUM, "onk_xchg_allreduce_f64: unknown op %d", op);
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  constexpr unsigned T = ONK_XCHG_MAX_WORLD * ONK_XCHG_MAX_VALUES;
  switch (op)
  {
    case ONK_OP_MIN: xchg_allreduce_kernel<ONK_OP_MIN><<<1, T, 0, L->stream>>>(values_dev, count, x->params); break;
    case ONK_OP_MAX: xchg_allreduce_kernel<ONK_OP_MAX><<<1, T, 0, L->stream>>>(values_dev, count, x->params); break;
    default:         xchg_allreduce_kernel<ONK_OP_SUM><<<1, T, 0, L->stream>>>(values_dev, count, x->params); break;
  }
  ONK_CHECK_LAUNCH();
  count_launch();
  return ONK_OK;
}

int onk_reduce_xchg_f64(int op, const double* in_dev, uint64_t n, double* out_dev, onk_xchg* x, int lane)
{
  if (int rc = xchg_usable(x, "onk_reduce_xchg_f64")) return rc;
  ONK_REQUIRE(out_dev != nullptr && (n == 0 || in_dev != nullptr), "onk_reduce_xchg_f64: null pointer");
  ONK_REQUIRE(((uintptr_t)in_dev & 7) == 0, "onk_reduce_xchg_f64: input is not 8-byte aligned");
  ONK_REQUIRE(op == ONK_OP_MIN || op == ONK_OP_MAX || op == ONK_OP_SUM, "onk_reduce_xchg_f64: unknown op %d", op);
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  switch (op)
  {
    case ONK_OP_MIN: return launch_reduce<ONK_OP_MIN>(in_dev, n, out_dev, L, &x->params);
    case ONK_OP_MAX: return launch_reduce<ONK_OP_MAX>(in_dev, n, out_dev, L, &x->params);
    default:         return launch_reduce<ONK_OP_SUM>(in_dev, n, out_dev, L, &x->params);
  }
}

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
