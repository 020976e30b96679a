// onk_stream.cu -- element-streaming kernels for sm_100a: axpy, parallel_memset, value-add, SOATL column sweeps.
//
// Reference behaviour: the per-element / per-row functors that the reference pushes through its K1/K3/K4 trampolines
// (include/onika/parallel/block_parallel_for_adapter.h:81-104, parallel_for.h:53-84: one element per thread, scalar
// 8-byte accesses, ceil(N/128) blocks) and the host loops of soatl::parallel_apply_simd (soatl/compute.h:214-232).
//
// B200 design (all HBM-bound, no reuse => no shared memory staging; see DESIGN.md for the roofline of each):
//   * 256-bit vector loads/stores (LDG.E.256 / STG.E.256), L1 no-allocate, evict-first on loads;
//   * tiles of THREADS x UNROLL x 32 B per CTA, all loads of a tile issued before the first use;
//   * persistent grid = CTAS_PER_SM x 148, tiles walked grid-stride so the whole chip streams one DRAM region;
//   * unfused arithmetic (__dmul_rn / __dadd_rn) so results are bit-identical to the reference's x86-64 build.
// Pointers that are not 32-byte aligned take the scalar (8-byte, still coalesced) path of the same kernel.
#include "onk_internal.cuh"

namespace onk
{
  constexpr unsigned STREAM_CTAS_PER_SM = 4;

  template<int UNROLL> struct Tile { static constexpr unsigned ELEMS = THREADS * UNROLL * 4; };

  // ---------------------------------------------------------------- axpy: y[i] = y[i] + a*x[i]  (2R + 1W)
  constexpr int AXPY_UNROLL = 2;
  __global__ void __launch_bounds__(THREADS, STREAM_CTAS_PER_SM)
  axpy_f64_kernel(double* __restrict__ y, const double* __restrict__ x, double a, uint64_t n, uint64_t ntiles)
  {
    constexpr unsigned TILE = Tile<AXPY_UNROLL>::ELEMS;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
    {
      const uint64_t base = t * TILE + threadIdx.x * 4;
      d4 vx[AXPY_UNROLL], vy[AXPY_UNROLL];
#pragma unroll
      for (int u = 0; u < AXPY_UNROLL; u++) vx[u] = ld_stream_ro(x + base + u * (THREADS * 4));
#pragma unroll
      for (int u = 0; u < AXPY_UNROLL; u++) vy[u] = ld_stream_rw(y + base + u * (THREADS * 4));
#pragma unroll
      for (int u = 0; u < AXPY_UNROLL; u++)
      {
        vy[u].a = __dadd_rn(vy[u].a, __dmul_rn(a, vx[u].a));
        vy[u].b = __dadd_rn(vy[u].b, __dmul_rn(a, vx[u].b));
        vy[u].c = __dadd_rn(vy[u].c, __dmul_rn(a, vx[u].c));
        vy[u].d = __dadd_rn(vy[u].d, __dmul_rn(a, vx[u].d));
        st_stream(y + base + u * (THREADS * 4), vy[u]);
      }
    }
    for (uint64_t i = ntiles * TILE + (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS)
      y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i]));
  }

  // ---------------------------------------------------------------- a[i] = a[i] + v   (1R + 1W)   value-add
  // ---------------------------------------------------------------- a[i] = a[i]*s + c (1R + 1W)   SOATL column sweep
  // One kernel for both: `ncols` columns of `len` doubles, column k at base + k*col_stride (doubles).
  // tools/tune.cu sweep (profiles/r1_tune_sweep_width_rmw_axpy.log): a read+write mix is fastest with ~32 KB of loads in
  // flight per SM (6.39 TB/s); deeper unrolling (128 KB/SM) costs 4 % -- the opposite of the read-only kernels
  constexpr int RMW_UNROLL = 1;
  constexpr int RMW_MAX_COLS = 16;
  struct RmwParams { double s; double c[RMW_MAX_COLS]; };

  template<bool SCALE>
  __global__ void __launch_bounds__(THREADS, STREAM_CTAS_PER_SM)
  rmw_f64_kernel(double* __restrict__ base, uint64_t col_stride, unsigned ncols, uint64_t len, uint64_t tiles_per_col,
                 const __grid_constant__ RmwParams prm)
  {
    constexpr unsigned TILE = Tile<RMW_UNROLL>::ELEMS;
    const uint64_t ntiles = tiles_per_col * ncols;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
    {
      const unsigned k = (unsigned)(t / tiles_per_col);
      const uint64_t tc = t - (uint64_t)k * tiles_per_col;
      double* col = base + (uint64_t)k * col_stride;
      const double c = prm.c[k];
      const uint64_t off = tc * TILE + threadIdx.x * 4;
      d4 v[RMW_UNROLL];
#pragma unroll
      for (int u = 0; u < RMW_UNROLL; u++) v[u] = ld_stream_rw(col + off + u * (THREADS * 4));
#pragma unroll
      for (int u = 0; u < RMW_UNROLL; u++)
      {
        if (SCALE)
        {
          v[u].a = __dadd_rn(__dmul_rn(v[u].a, prm.s), c); v[u].b = __dadd_rn(__dmul_rn(v[u].b, prm.s), c);
          v[u].c = __dadd_rn(__dmul_rn(v[u].c, prm.s), c); v[u].d = __dadd_rn(__dmul_rn(v[u].d, prm.s), c);
        }
        else
        {
          v[u].a = __dadd_rn(v[u].a, c); v[u].b = __dadd_rn(v[u].b, c);
          v[u].c = __dadd_rn(v[u].c, c); v[u].d = __dadd_rn(v[u].d, c);
        }
        st_stream(col + off + u * (THREADS * 4), v[u]);
      }
    }
    // column tails (and whole columns when the vector path is disabled: tiles_per_col == 0)
    const uint64_t tail0 = tiles_per_col * TILE;
    const uint64_t tail_len = len - tail0;
    const uint64_t ntail = tail_len * ncols;
    for (uint64_t e = (uint64_t)blockIdx.x * THREADS + threadIdx.x; e < ntail; e += (uint64_t)gridDim.x * THREADS)
    {
      const unsigned k = (unsigned)(e / tail_len);
      const uint64_t i = tail0 + (e - (uint64_t)k * tail_len);
      double* p = base + (uint64_t)k * col_stride + i;
      *p = SCALE ? __dadd_rn(__dmul_rn(*p, prm.s), prm.c[k]) : __dadd_rn(*p, prm.c[k]);
    }
  }

  // ---------------------------------------------------------------- parallel_memset (1W)
  constexpr int SET_UNROLL = 4;
  __global__ void __launch_bounds__(THREADS, STREAM_CTAS_PER_SM)
  memset_b64_kernel(double* __restrict__ d, double v, uint64_t n, uint64_t ntiles)
  {
    constexpr unsigned TILE = Tile<SET_UNROLL>::ELEMS;
    const d4 vv{v, v, v, v};
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
    {
      const uint64_t base = t * TILE + threadIdx.x * 4;
#pragma unroll
      for (int u = 0; u < SET_UNROLL; u++) st_stream(d + base + u * (THREADS * 4), vv);
    }
    for (uint64_t i = ntiles * TILE + (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS) d[i] = v;
  }

  // fragments (< 32 bytes) before/after the 32-byte aligned body: bytes of the replicated pattern, phase preserved
  __global__ void memset_frag_kernel(uint8_t* __restrict__ p, unsigned n, uint64_t p64, unsigned phase)
  {
    if (threadIdx.x < n) p[threadIdx.x] = (uint8_t)(p64 >> (8 * ((phase + threadIdx.x) & 7)));
  }

  // ---------------------------------------------------------------- SOATL benchmark kernel (3R + 1W), tests/benchmark.cpp:96-104
  constexpr int DIST_UNROLL = 1;
  __device__ __forceinline__ double dist_one(double x, double y, double z, double ax, double ay, double az)
  {
    x = __dsub_rn(x, ax);
    y = __dsub_rn(y, ay);
    z = __dsub_rn(z, az);
    double d = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    x = __ddiv_rn(x, d);
    y = __ddiv_rn(y, d);
    z = __ddiv_rn(z, d);
    return __dadd_rn(d, __ddiv_rn(x, __dmul_rn(y, z)));
  }

  __global__ void __launch_bounds__(THREADS, STREAM_CTAS_PER_SM)
  dist_f64_kernel(double* __restrict__ e, const double* __restrict__ rx, const double* __restrict__ ry, const double* __restrict__ rz,
                  uint64_t n, uint64_t ntiles, double ax, double ay, double az)
  {
    constexpr unsigned TILE = Tile<DIST_UNROLL>::ELEMS;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
    {
      const uint64_t i = t * TILE + threadIdx.x * 4;
      const d4 x = ld_stream_ro(rx + i), y = ld_stream_ro(ry + i), z = ld_stream_ro(rz + i);
      d4 o;
      o.a = dist_one(x.a, y.a, z.a, ax, ay, az);
      o.b = dist_one(x.b, y.b, z.b, ax, ay, az);
      o.c = dist_one(x.c, y.c, z.c, ax, ay, az);
      o.d = dist_one(x.d, y.d, z.d, ax, ay, az);
      st_stream(e + i, o);
    }
    for (uint64_t i = ntiles * TILE + (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * THREADS)
      e[i] = dist_one(rx[i], ry[i], rz[i], ax, ay, az);
  }

  // ---------------------------------------------------------------- SOATL repack (soatl::copy between layouts)
  constexpr int PACK_MAX_FIELDS = 24;
  struct PackParams { uint64_t src_off[PACK_MAX_FIELDS]; uint64_t dst_off[PACK_MAX_FIELDS]; uint64_t bytes[PACK_MAX_FIELDS]; int nfields; };
  constexpr unsigned PACK_CHUNK = 16384;   // bytes per (field, chunk) work item

  __global__ void __launch_bounds__(THREADS)
  repack_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, const __grid_constant__ PackParams prm, uint64_t max_chunks)
  {
    const uint64_t nitems = max_chunks * prm.nfields;
    for (uint64_t it = blockIdx.x; it < nitems; it += gridDim.x)
    {
      const int k = (int)(it / max_chunks);
      const uint64_t b0 = (it - (uint64_t)k * max_chunks) * PACK_CHUNK;
      if (b0 >= prm.bytes[k]) continue;
      const uint64_t nb = (prm.bytes[k] - b0 < PACK_CHUNK) ? prm.bytes[k] - b0 : PACK_CHUNK;
      const uint8_t* s = src + prm.src_off[k] + b0;
      uint8_t* d = dst + prm.dst_off[k] + b0;
      // 1. bytes up to the first 16-byte boundary of the DESTINATION
      const unsigned dmis = (unsigned)((uintptr_t)d & 15u);
      uint64_t head = dmis ? 16u - dmis : 0u;
      if (head > nb) head = nb;
      for (uint64_t i = threadIdx.x; i < head; i += THREADS) d[i] = s[i];
      const uint8_t* s2 = s + head;
      uint8_t* d2 = d + head;
      const uint64_t rem = nb - head;
      const unsigned sh = (unsigned)((uintptr_t)s2 & 15u);     // source misalignment relative to the (now aligned) destination
      uint64_t done = 0;
      if (sh == 0)
      {
        const uint64_t nv = rem / 16;
        for (uint64_t i = threadIdx.x; i < nv; i += THREADS) ((uint4*)d2)[i] = __ldg((const uint4*)s2 + i);
        done = nv * 16;
      }
      else
      {
        // 2. shifted copy: every 16-byte destination word is cut out of two ALIGNED 16-byte source words (the wire layout
        //    <1,1,...> puts columns at arbitrary byte offsets).  The last word is left to the byte tail so that no read
        //    goes past the end of the range.
        uint64_t nv = rem / 16;
        if (nv > 0) nv -= 1;
        const uint4* S = (const uint4*)(s2 - sh);
        const unsigned q = sh >> 3, r = (sh & 7u) * 8u;
        for (uint64_t i = threadIdx.x; i < nv; i += THREADS)
        {
          const uint4 a = __ldg(S + i), b = __ldg(S + i + 1);
          unsigned long long w[4] = { (unsigned long long)a.x | ((unsigned long long)a.y << 32), (unsigned long long)a.z | ((unsigned long long)a.w << 32),
                                      (unsigned long long)b.x | ((unsigned long long)b.y << 32), (unsigned long long)b.z | ((unsigned long long)b.w << 32) };
          const unsigned long long lo0 = q ? w[1] : w[0], lo1 = q ? w[2] : w[1], lo2 = q ? w[3] : w[2];
          const unsigned long long o0 = r ? ((lo0 >> r) | (lo1 << (64u - r))) : lo0;
          const unsigned long long o1 = r ? ((lo1 >> r) | (lo2 << (64u - r))) : lo1;
          uint4 o; o.x = (unsigned)o0; o.y = (unsigned)(o0 >> 32); o.z = (unsigned)o1; o.w = (unsigned)(o1 >> 32);
          ((uint4*)d2)[i] = o;
        }
        done = nv * 16;
      }
      // 3. tail bytes
      for (uint64_t i = done + threadIdx.x; i < rem; i += THREADS) d2[i] = s2[i];
    }
  }

  static inline bool aligned32(const void* p) { return ((uintptr_t)p & 31) == 0; }
}

using namespace onk;

extern "C" {

int onk_axpy_f64(double* y_dev, const double* x_dev, double a, uint64_t n, int lane)
{
  if (n == 0) return ONK_OK;
  ONK_REQUIRE(y_dev != nullptr && x_dev != nullptr, "onk_axpy_f64: null pointer");
  ONK_REQUIRE((((uintptr_t)y_dev | (uintptr_t)x_dev) & 7) == 0, "onk_axpy_f64: pointers must be 8-byte aligned");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  const unsigned TILE = Tile<AXPY_UNROLL>::ELEMS;
  const uint64_t ntiles = (aligned32(y_dev) && aligned32(x_dev)) ? n / TILE : 0;
  const unsigned grid = ntiles ? grid_for(ntiles, 1, STREAM_CTAS_PER_SM) : grid_for(n, THREADS * 8, STREAM_CTAS_PER_SM);
  axpy_f64_kernel<<<grid, THREADS, 0, L->stream>>>(y_dev, x_dev, a, n, ntiles);
  ONK_CHECK_LAUNCH(); count_launch();
  return ONK_OK;
}

int onk_value_add_f64(double* array_dev, uint64_t rows, uint64_t cols, double value, int lane)
{
  const uint64_t n = rows * cols;
  if (n == 0) return ONK_OK;
  ONK_REQUIRE(array_dev != nullptr, "onk_value_add_f64: null pointer");
  ONK_REQUIRE(((uintptr_t)array_dev & 7) == 0, "onk_value_add_f64: pointer must be 8-byte aligned");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  // Array2D is dense row-major (include/onika/extras/array2d.h:13-24): rows x cols is one contiguous stream
  const unsigned TILE = Tile<RMW_UNROLL>::ELEMS;
  const uint64_t tiles = aligned32(array_dev) ? n / TILE : 0;
  RmwParams prm{}; prm.s = 1.0; prm.c[0] = value;
  const unsigned grid = tiles ? grid_for(tiles, 1, STREAM_CTAS_PER_SM) : grid_for(n, THREADS * 8, STREAM_CTAS_PER_SM);
  rmw_f64_kernel<false><<<grid, THREADS, 0, L->stream>>>(array_dev, 0, 1, n, tiles, prm);
  ONK_CHECK_LAUNCH(); count_launch();
  return ONK_OK;
}

int onk_parallel_memset(void* data_dev, uint64_t n, uint64_t pattern, int elem_bytes, int lane)
{
  if (n == 0) return ONK_OK;
  ONK_REQUIRE(data_dev != nullptr, "onk_parallel_memset: null pointer");
  ONK_REQUIRE(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, "onk_parallel_memset: element size %d not in {1,2,4,8}", elem_bytes);
  ONK_REQUIRE(((uintptr_t)data_dev % elem_bytes) == 0, "onk_parallel_memset: pointer not aligned on the element size");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  // replicate the pattern to 64 bits; a 64-bit aligned span is then filled with 256-bit stores
  uint64_t p64 = pattern;
  if (elem_bytes == 1) { p64 &= 0xff; p64 |= p64 << 8; p64 |= p64 << 16; p64 |= p64 << 32; }
  else if (elem_bytes == 2) { p64 &= 0xffff; p64 |= p64 << 16; p64 |= p64 << 32; }
  else if (elem_bytes == 4) { p64 &= 0xffffffffu; p64 |= p64 << 32; }
  const uint64_t total = n * (uint64_t)elem_bytes;
  uint8_t* p = (uint8_t*)data_dev;
  // Patterns that repeat every 32 bits (0.0, integers, bytes, ...) go to the device's fill engine: measured 7.3 TB/s
  // against ~6.3 TB/s for ANY SM-issued store shape, TMA bulk stores included (profiles/r1_tune_sweep_memset.log).
  // General 64-bit values (most doubles) need the kernel below.
  if ((uint32_t)p64 == (uint32_t)(p64 >> 32) && ((uintptr_t)p & 3) == 0 && (total & 3) == 0)
  {
    typedef int (*memset32_t)(unsigned long long, unsigned int, size_t, void*);
    static memset32_t fill32 = nullptr;
    if (!fill32)
    {
      void* fn = nullptr; cudaDriverEntryPointQueryResult q;
      ONK_CUDA(cudaGetDriverEntryPoint("cuMemsetD32Async", &fn, cudaEnableDefault, &q));
      fill32 = (memset32_t)fn;
    }
    if (fill32)
    {
      const int r = fill32((unsigned long long)(uintptr_t)p, (unsigned int)p64, total / 4, (void*)L->stream);
      if (r != 0) { set_error("onk_parallel_memset: cuMemsetD32Async failed with CUresult %d", r); return ONK_ERR_CUDA; }
      return ONK_OK;
    }
  }
  uint64_t head = ((32 - ((uintptr_t)p & 31)) & 31);
  if (head > total) head = total;
  const uint64_t body = (total - head) / 8;          // 64-bit words starting at a 32-byte boundary
  const uint64_t tail = total - head - body * 8;
  double v; memcpy(&v, &p64, 8);
  if (body)
  {
    const unsigned TILE = Tile<SET_UNROLL>::ELEMS;
    const uint64_t ntiles = body / TILE;
    const unsigned grid = ntiles ? grid_for(ntiles, 1, STREAM_CTAS_PER_SM) : grid_for(body, THREADS * 8, STREAM_CTAS_PER_SM);
    memset_b64_kernel<<<grid, THREADS, 0, L->stream>>>((double*)(p + head), v, body, ntiles);
    ONK_CHECK_LAUNCH(); count_launch();
  }
  // head/tail fragments (< 32 bytes each)
  if (head)
  {
    memset_frag_kernel<<<1, 32, 0, L->stream>>>(p, (unsigned)head, p64, 0u);
    ONK_CHECK_LAUNCH(); count_launch();
  }
  if (tail)
  {
    memset_frag_kernel<<<1, 32, 0, L->stream>>>(p + head + body * 8, (unsigned)tail, p64, (unsigned)((head + body * 8) & 7));
    ONK_CHECK_LAUNCH(); count_launch();
  }
  return ONK_OK;
}

int onk_soatl_rmw_f64(void* storage_dev, uint64_t align, uint64_t chunk, int nfields, uint64_t n, uint64_t capacity,
                      double s, const double* c_host, int lane)
{
  ONK_REQUIRE(nfields >= 0 && nfields <= RMW_MAX_COLS, "onk_soatl_rmw_f64: at most %d fields", RMW_MAX_COLS);
  ONK_REQUIRE(chunk > 0 && align > 0 && (align & (align - 1)) == 0, "onk_soatl_rmw_f64: bad align/chunk");
  const uint64_t n_padded = ((n + chunk - 1) / chunk) * chunk;   // parallel_apply_simd sweeps whole chunks (compute.h:221-231)
  ONK_REQUIRE(n_padded <= capacity, "onk_soatl_rmw_f64: size %llu exceeds capacity %llu", (unsigned long long)n, (unsigned long long)capacity);
  if (n_padded == 0 || nfields == 0) return ONK_OK;
  ONK_REQUIRE(storage_dev != nullptr && c_host != nullptr, "onk_soatl_rmw_f64: null pointer");
  const uint64_t col_bytes = (capacity * 8 + align - 1) & ~(align - 1);   // field_arrays.h:494-501
  ONK_REQUIRE((col_bytes & 7) == 0 && ((uintptr_t)storage_dev & 7) == 0, "onk_soatl_rmw_f64: fp64 columns must be 8-byte aligned");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  RmwParams prm{}; prm.s = s;
  for (int k = 0; k < nfields; k++) prm.c[k] = c_host[k];
  const unsigned TILE = Tile<RMW_UNROLL>::ELEMS;
  const bool vec = aligned32(storage_dev) && (col_bytes % 32) == 0;
  const uint64_t tiles = vec ? n_padded / TILE : 0;
  const unsigned grid = tiles ? grid_for(tiles * nfields, 1, STREAM_CTAS_PER_SM) : grid_for(n_padded * nfields, THREADS * 8, STREAM_CTAS_PER_SM);
  rmw_f64_kernel<true><<<grid, THREADS, 0, L->stream>>>((double*)storage_dev, col_bytes / 8, (unsigned)nfields, n_padded, tiles, prm);
  ONK_CHECK_LAUNCH(); count_launch();
  return ONK_OK;
}

int onk_soatl_dist_f64(double* e_dev, const double* rx_dev, const double* ry_dev, const double* rz_dev,
                       uint64_t n_padded, double ax, double ay, double az, int lane)
{
  if (n_padded == 0) return ONK_OK;
  ONK_REQUIRE(e_dev && rx_dev && ry_dev && rz_dev, "onk_soatl_dist_f64: null pointer");
  ONK_REQUIRE((((uintptr_t)e_dev | (uintptr_t)rx_dev | (uintptr_t)ry_dev | (uintptr_t)rz_dev) & 7) == 0, "onk_soatl_dist_f64: pointers must be 8-byte aligned");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  const unsigned TILE = Tile<DIST_UNROLL>::ELEMS;
  const bool vec = aligned32(e_dev) && aligned32(rx_dev) && aligned32(ry_dev) && aligned32(rz_dev);
  const uint64_t ntiles = vec ? n_padded / TILE : 0;
  const unsigned grid = ntiles ? grid_for(ntiles, 1, STREAM_CTAS_PER_SM) : grid_for(n_padded, THREADS * 4, STREAM_CTAS_PER_SM);
  dist_f64_kernel<<<grid, THREADS, 0, L->stream>>>(e_dev, rx_dev, ry_dev, rz_dev, n_padded, ntiles, ax, ay, az);
  ONK_CHECK_LAUNCH(); count_launch();
  return ONK_OK;
}

int onk_soatl_repack(void* dst_dev, uint64_t dst_align, uint64_t dst_capacity,
                     const void* src_dev, uint64_t src_align, uint64_t src_capacity,
                     const uint32_t* field_bytes, int nfields, uint64_t n, int lane)
{
  ONK_REQUIRE(nfields >= 0 && nfields <= PACK_MAX_FIELDS, "onk_soatl_repack: at most %d fields", PACK_MAX_FIELDS);
  ONK_REQUIRE(n <= dst_capacity && n <= src_capacity, "onk_soatl_repack: n exceeds a capacity");
  if (n == 0 || nfields == 0) return ONK_OK;
  ONK_REQUIRE(dst_dev && src_dev && field_bytes, "onk_soatl_repack: null pointer");
  Lane* L; if (int rc = lane_get(lane, &L)) return rc;
  PackParams prm{}; prm.nfields = nfields;
  onk_soatl_layout(src_align, field_bytes, nfields, src_capacity, prm.src_off);
  onk_soatl_layout(dst_align, field_bytes, nfields, dst_capacity, prm.dst_off);
  uint64_t maxb = 0;
  for (int k = 0; k < nfields; k++) { prm.bytes[k] = n * field_bytes[k]; if (prm.bytes[k] > maxb) maxb = prm.bytes[k]; }
  const uint64_t max_chunks = (maxb + PACK_CHUNK - 1) / PACK_CHUNK;
  const unsigned grid = grid_for(max_chunks * nfields, 1, 8);
  repack_kernel<<<grid, THREADS, 0, L->stream>>>((uint8_t*)dst_dev, (const uint8_t*)src_dev, prm, max_chunks);
  ONK_CHECK_LAUNCH(); count_launch();
  return ONK_OK;
}

/* ---- host arithmetic shared by every front end (no device needed) ---- */

uint64_t onk_soatl_update_capacity(uint64_t size, uint64_t capacity, uint64_t chunk)
{
  // growth policy of the reference's default allocation strategy (include/onika/memory/alloc_strategy.h:44-67):
  // grow to max(size, 2*capacity), shrink when at most half full, always a whole number of chunks
  if (chunk == 0) chunk = 1;
  uint64_t target = capacity;
  if (size > capacity) target = (capacity == 0 || capacity <= size / 2) ? size : 2 * capacity;
  else if (size == 0 || size <= capacity / 2) target = size;
  const uint64_t nchunks = (target + chunk - 1) / chunk;
  return nchunks * chunk;
}

uint64_t onk_soatl_layout(uint64_t align, const uint32_t* field_bytes, int nfields, uint64_t capacity, uint64_t* offsets)
{
  // one packed allocation; column k begins where column k-1 ends, rounded up to `align` (field_arrays.h:494-512)
  uint64_t cursor = 0;
  const uint64_t mask = align - 1;
  for (int k = 0; k < nfields; k++)
  {
    if (offsets != nullptr) offsets[k] = cursor;
    cursor += (capacity * (uint64_t)field_bytes[k] + mask) & ~mask;
  }
  return cursor;
}

} // extern "C"
