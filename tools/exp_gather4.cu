// exp_gather4.cu -- can the TMA unit's tile::gather4 mode (4 arbitrary 128-byte rows per request, sm_100) fetch the scattered
// `e` lines of per-cell SOATL blocks faster than per-lane cp.async?  Micro-benchmark: rows of 128 bytes, one wanted row out of
// every `stride_rows` (4 = a 512-byte cell block), per-warp rings of STAGES x 32 rows, lanes 0,4,...,28 issue one gather4 each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/exp_gather4 tools/exp_gather4.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)



__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k_read(const double* __restrict__ p, size_t n, double* out)
{ double a = 0; for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a += p[i]; if (a == 123.456) *out = a; }

// MODE 2: no staging at all -- 4 lanes per row, 32 bytes per lane straight into registers (the strided-line microbenchmark shape)
__global__ void __launch_bounds__(256) k_direct(const double* __restrict__ base, uint64_t nwanted, unsigned stride_rows, double* __restrict__ out)
{
  double acc = 0.0;
  const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = t0; i < nwanted * 4; i += nt)
  {
    const uint64_t c = i >> 2; const unsigned q = (unsigned)(i & 3);
    const double* p = base + ((c * stride_rows + (stride_rows - 1)) * 16 + q * 4);
    double a, b, cc, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(cc), "=d"(d) : "l"(p));
    acc += a + b + cc + d;
  }
  out[t0] = acc;
}

// the same two shapes driven by an explicit list of wanted rows (the real C4 pattern: 512- and 1024-byte blocks, 1 or 2 lines each)
__global__ void __launch_bounds__(256) k_direct_list(const double* __restrict__ base, const unsigned* __restrict__ rows, uint64_t nwanted, double* __restrict__ out)
{
  double acc = 0.0;
  const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = t0; i < nwanted * 4; i += nt)
  {
    const uint64_t c = i >> 2; const unsigned q = (unsigned)(i & 3);
    const double* p = base + ((uint64_t)__ldg(rows + c) * 16 + q * 4);
    double a, b, cc, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(cc), "=d"(d) : "l"(p));
    acc += a + b + cc + d;
  }
  out[t0] = acc;
}
template<int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32) k_ring_list(const double* __restrict__ base, const unsigned* __restrict__ rows, uint64_t nwanted, double* __restrict__ out)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char* ring = smem + (size_t)w * STAGES * 4096;
  const uint64_t warp_global = (uint64_t)blockIdx.x * WARPS + w, nwarps = (uint64_t)gridDim.x * WARPS;
  const uint64_t nsteps = (nwanted + 31) / 32;
  const uint64_t mine = warp_global < nsteps ? (nsteps - warp_global + nwarps - 1) / nwarps : 0;
  unsigned nx = 0;
  auto preload = [&](uint64_t k) { const uint64_t c = (warp_global + k * nwarps) * 32 + lane; nx = (k < mine && c < nwanted) ? __ldg(rows + c) : 0u; };
  preload(0);
  auto issue = [&](int stage, uint64_t k)
  {
    const unsigned row = nx; preload(k + 1);
    unsigned char* dst = ring + (size_t)stage * 4096;
    const unsigned piece = lane & 7, g4 = lane >> 3;
#pragma unroll
    for (int it = 0; it < 8; it++)
    {
      const int src_lane = it * 4 + (int)g4;
      const unsigned r = __shfl_sync(0xffffffffu, row, src_lane);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst + src_lane * 128 + ((piece ^ (src_lane & 7)) << 4))), "l"((const char*)base + (size_t)r * 128 + piece * 16) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int s = 0; s < STAGES; s++) { if ((uint64_t)s < mine) issue(s, s); else asm volatile("cp.async.commit_group;" ::: "memory"); }
  double acc = 0.0;
  for (uint64_t k = 0; k < mine; k++)
  {
    const int s = (int)(k % STAGES);
    asm volatile("cp.async.wait_group %0;" :: "n"(STAGES - 1) : "memory"); __syncwarp();
    const unsigned rowa = (smem_u32(ring + (size_t)s * 4096) + lane * 128u) ^ ((lane & 7u) << 4);
#pragma unroll
    for (unsigned c = 0; c < 8; c++) { double x0, x1; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(rowa ^ (c << 4)) : "memory"); acc += x0; acc += x1; }   // dependent chain, as the in-order fold
    __syncwarp();
    if (k + STAGES < mine) issue(s, k + STAGES); else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template<int MODE, int WARPS, int STAGES>   // 0 = gather4 , 1 = per-lane cp.async 16 B (8 lanes per row)
__global__ void __launch_bounds__(WARPS * 32) k_gather(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ base, uint64_t nwanted, unsigned stride_rows,
                                                       double* __restrict__ out)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[WARPS][STAGES];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char* ring = smem + (size_t)w * STAGES * 4096;
  if (lane == 0) for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[w][s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const uint64_t warp_global = (uint64_t)blockIdx.x * WARPS + w, nwarps = (uint64_t)gridDim.x * WARPS;
  const uint64_t nsteps = (nwanted + 31) / 32;
  const uint64_t mine = warp_global < nsteps ? (nsteps - warp_global + nwarps - 1) / nwarps : 0;
  auto issue = [&](int stage, uint64_t k)
  {
    const uint64_t c = (warp_global + k * nwarps) * 32 + lane;               // wanted row index of this lane
    const int row = (int)((c < nwanted ? c : nwanted - 1) * stride_rows + (stride_rows - 1));
    unsigned char* dst = ring + (size_t)stage * 4096;
    if (MODE == 0)
    {
      const int r1 = __shfl_down_sync(0xffffffffu, row, 1), r2 = __shfl_down_sync(0xffffffffu, row, 2), r3 = __shfl_down_sync(0xffffffffu, row, 3);
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bars[w][stage])), "r"(4096u) : "memory");
      __syncwarp();
      if ((lane & 3) == 0)
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     :: "r"(smem_u32(dst + lane * 128)), "l"(&tmap), "r"(0), "r"(row), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(&bars[w][stage])) : "memory");
    }
    else
    {
      const unsigned piece = lane & 7, g4 = lane >> 3;
#pragma unroll
      for (int it = 0; it < 8; it++)
      {
        const int src_lane = it * 4 + (int)g4;
        const int r = __shfl_sync(0xffffffffu, row, src_lane);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst + src_lane * 128 + piece * 16)), "l"((const char*)base + (size_t)r * 128 + piece * 16) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  for (int s = 0; s < STAGES; s++) { if ((uint64_t)s < mine) issue(s, s); else if (MODE == 1) asm volatile("cp.async.commit_group;" ::: "memory"); }
  double acc = 0.0;
  unsigned phase[STAGES]; for (int s = 0; s < STAGES; s++) phase[s] = 0;
  for (uint64_t k = 0; k < mine; k++)
  {
    const int s = (int)(k % STAGES);
    if (MODE == 0)
    {
      unsigned ok = 0;
      while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bars[w][s])), "r"(phase[s]) : "memory");
      phase[s] ^= 1u;
    }
    else { asm volatile("cp.async.wait_group %0;" :: "n"(STAGES - 1) : "memory"); __syncwarp(); }
    const double* rowp = reinterpret_cast<const double*>(ring + (size_t)s * 4096 + lane * 128);
#pragma unroll
    for (int i = 0; i < 16; i++) acc += rowp[(i + lane) & 15];            // (rotated start: no bank conflicts)
    __syncwarp();
    if (k + STAGES < mine) issue(s, k + STAGES); else if (MODE == 1) asm volatile("cp.async.commit_group;" ::: "memory");
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char** argv)
{
  const uint64_t nwanted = argc > 1 ? strtoull(argv[1], 0, 10) : 4194304ull;
  const unsigned stride_rows = argc > 2 ? atoi(argv[2]) : 4;
  const int boxrows = argc > 3 ? atoi(argv[3]) : 1;
  const uint64_t nrows = nwanted * stride_rows;
  double* a; CK(cudaMalloc(&a, nrows * 128));
  std::vector<double> h(nrows * 16);
  for (uint64_t r = 0; r < nrows; r++) for (int j = 0; j < 16; j++) h[r * 16 + j] = (r % stride_rows == stride_rows - 1) ? 1.0 : 1e9;     // wanted rows hold ones
  CK(cudaMemcpy(a, h.data(), nrows * 128, cudaMemcpyHostToDevice));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
  double* flush; CK(cudaMalloc(&flush, 1ull << 30));
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = { 16, nrows }; const cuuint64_t gstride[1] = { 128 }; const cuuint32_t box[2] = { 16, (cuuint32_t)boxrows }; const cuuint32_t estr[2] = { 1, 1 };
  CUresult r = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, a, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box {16,%d}: CUresult %d\n", boxrows, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](const char* name, auto launch, size_t nthreads)
  {
    float best = 1e9f;
    for (int rep = 0; rep < 6; rep++)
    {
      k_read<<<sms * 8, 256>>>(flush, (1ull << 30) / 8, out);          // flush L2 by READING 1 GiB (clean lines)
      CK(cudaMemsetAsync(out, 0, sizeof(double) * nthreads));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep > 0 && ms < best) best = ms;
    }
    std::vector<double> ho(nthreads);
    CK(cudaMemcpy(ho.data(), out, ho.size() * 8, cudaMemcpyDeviceToHost));
    double tot = 0; for (double v : ho) tot += v;
    printf("%-34s stride %u rows: %8.1f us  %7.1f GB/s useful ; sum %.0f\n", name, stride_rows, best * 1e3, nwanted * 128.0 / best / 1e6, tot);
  };
  CK(cudaFree(out)); CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
  run("direct LDG.256, 4 CTAs x 256", [&]{ k_direct<<<sms * 4, 256>>>(a, nwanted, stride_rows, out); }, (size_t)sms * 4 * 256);
  run("direct LDG.256, 8 CTAs x 256", [&]{ k_direct<<<sms * 8, 256>>>(a, nwanted, stride_rows, out); }, (size_t)sms * 8 * 256);
#define GO(MODE, W, S, C) do { const size_t smem = (size_t)W * S * 4096; \
    CK(cudaFuncSetAttribute(k_gather<MODE, W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    char nm[64]; snprintf(nm, sizeof nm, "%s %d warps x %d stages x %d CTAs", MODE == 0 ? "gather4 " : "cp.async", W, S, C); \
    run(nm, [&]{ k_gather<MODE, W, S><<<sms * C, W * 32, smem>>>(tmap, a, nwanted, stride_rows, out); }, (size_t)sms * C * W * 32); } while (0)
  GO(1, 8, 3, 2); GO(0, 8, 3, 2);
  GO(1, 8, 6, 1); GO(0, 8, 6, 1);
  GO(1, 16, 3, 1); GO(0, 16, 3, 1);
  GO(1, 4, 3, 4); GO(0, 4, 3, 4);
  GO(1, 8, 2, 3); GO(0, 8, 2, 3);
  GO(1, 4, 6, 2); GO(0, 4, 6, 2);
  GO(1, 2, 6, 4); GO(0, 2, 6, 4);
  GO(1, 8, 1, 6); GO(0, 8, 1, 6);
  GO(1, 16, 1, 3); GO(0, 16, 1, 3);
  // ---- the real C4 pattern: Poisson(16)-like cells, 57 % one line at the end of a 512-byte block, 43 % two lines at the end of a 1024-byte block ----
  {
    std::vector<unsigned> rows; rows.reserve(nwanted * 2);
    uint64_t cursor = 0, state = 88172645463325252ull; uint64_t ncell = 0;
    while (rows.size() + 2 <= nwanted && (cursor + 8) < nrows)
    {
      state ^= state << 13; state ^= state >> 7; state ^= state << 17;
      const bool two = (state % 1000) < 433;
      if (two) { rows.push_back((unsigned)(cursor + 6)); rows.push_back((unsigned)(cursor + 7)); cursor += 8; }
      else { rows.push_back((unsigned)(cursor + 3)); cursor += 4; }
      ncell++;
    }
    const uint64_t nl = rows.size();
    // wanted rows hold ones in the buffer only at residue stride-1 of the regular pattern; refill so that every listed row holds ones
    for (uint64_t r = 0; r < cursor; r++) for (int j = 0; j < 16; j++) h[r * 16 + j] = 1e9;
    for (unsigned r : rows) for (int j = 0; j < 16; j++) h[(uint64_t)r * 16 + j] = 1.0;
    CK(cudaMemcpy(a, h.data(), nrows * 128, cudaMemcpyHostToDevice));
    unsigned* drows; CK(cudaMalloc(&drows, nl * 4)); CK(cudaMemcpy(drows, rows.data(), nl * 4, cudaMemcpyHostToDevice));
    printf("C4 pattern: %llu cells, %llu lines over %.1f MB of blocks (expect sum %.0f)\n", (unsigned long long)ncell, (unsigned long long)nl, cursor * 128.0 / 1e6, nl * 16.0);
    const uint64_t keep = nwanted; (void)keep;
    auto runl = [&](const char* name, auto launch, size_t nthreads)
    {
      float best = 1e9f;
      for (int rep = 0; rep < 6; rep++)
      {
        k_read<<<sms * 8, 256>>>(flush, (1ull << 30) / 8, out);
        CK(cudaMemsetAsync(out, 0, sizeof(double) * nthreads));
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep > 0 && ms < best) best = ms;
      }
      std::vector<double> ho(nthreads); CK(cudaMemcpy(ho.data(), out, ho.size() * 8, cudaMemcpyDeviceToHost));
      double tot = 0; for (double v : ho) tot += v;
      printf("%-40s : %8.1f us  %7.1f GB/s useful, %6.2f G lines/s ; sum %.0f\n", name, best * 1e3, nl * 128.0 / best / 1e6, nl / (best * 1e6), tot);
    };
    runl("C4 list, direct LDG.256, 4 CTAs x 256", [&]{ k_direct_list<<<sms * 4, 256>>>(a, drows, nl, out); }, (size_t)sms * 4 * 256);
    runl("C4 list, direct LDG.256, 8 CTAs x 256", [&]{ k_direct_list<<<sms * 8, 256>>>(a, drows, nl, out); }, (size_t)sms * 8 * 256);
#define GOL(W, S, C) do { const size_t smem = (size_t)W * S * 4096; CK(cudaFuncSetAttribute(k_ring_list<W, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      char nm[64]; snprintf(nm, sizeof nm, "C4 list, cp.async ring %d warps x %d stages x %d", W, S, C); \
      runl(nm, [&]{ k_ring_list<W, S><<<sms * C, W * 32, smem>>>(a, drows, nl, out); }, (size_t)sms * C * W * 32); } while (0)
    GOL(8, 3, 1); GOL(8, 3, 2); GOL(16, 3, 1); GOL(8, 6, 1); GOL(4, 3, 4); GOL(12, 3, 1); GOL(16, 2, 1); GOL(12, 4, 1); GOL(24, 2, 1); GOL(32, 1, 1); GOL(16, 1, 2); GOL(10, 3, 1); GOL(6, 3, 2); GOL(11, 4, 1); GOL(8, 2, 2); GOL(8, 4, 1);
  }
  return 0;
}
