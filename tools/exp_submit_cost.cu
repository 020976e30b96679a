// experiment: host-side cost of the pieces of one scheduled operation (not part of the product)
#include <onika_b200.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__global__ void tiny(double* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.0; }
template<class F> double us(F f, int n = 2000) { auto t0 = std::chrono::steady_clock::now(); for (int i = 0; i < n; i++) f(); auto t1 = std::chrono::steady_clock::now(); return std::chrono::duration<double, std::micro>(t1 - t0).count() / n; }
int main()
{
  onk_init(0);
  void* s = nullptr; onk_lane_stream(0, &s); cudaStream_t st = (cudaStream_t)s;
  double* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  cudaEvent_t e; cudaEventCreate(&e);
  onk_op* op; onk_op_create(&op); void* scr; onk_op_device_scratch(op, &scr);
  void* fn = nullptr; cudaGetFuncBySymbol((cudaFunction_t*)&fn, (const void*)tiny);
  unsigned g[3] = {8,1,1}, b[3] = {128,1,1}; void* args[1] = { &d };
  auto sync = [&]{ cudaStreamSynchronize(st); };
  printf("cudaEventRecord            %.2f us\n", us([&]{ cudaEventRecord(e, st); })); sync();
  printf("tiny<<<>>>                 %.2f us\n", us([&]{ tiny<<<8,128,0,st>>>(d); })); sync();
  printf("onk_launch                 %.2f us\n", us([&]{ onk_launch(fn, g, b, 0, 0, args); })); sync();
  printf("onk_op_begin+end (no data) %.2f us\n", us([&]{ onk_op_begin(op, 0, nullptr, 0, 0); onk_op_end(op, nullptr, 0, nullptr, nullptr); })); sync();
  printf("begin+launch+end           %.2f us\n", us([&]{ onk_op_begin(op, 0, nullptr, 0, 0); onk_launch(fn, g, b, 0, 0, args); onk_op_end(op, nullptr, 0, nullptr, nullptr); })); sync();
  cudaStreamCaptureStatus cs; printf("cudaStreamIsCapturing      %.2f us\n", us([&]{ cudaStreamIsCapturing(st, &cs); }));
  float ms; printf("sync+elapsed (idle stream) %.2f us\n", us([&]{ onk_lane_sync(0); onk_op_elapsed_ms(op, &ms); }));
  printf("launch+sync (synchronous op) %.2f us\n", us([&]{ onk_op_begin(op, 0, nullptr, 0, 0); onk_launch(fn, g, b, 0, 0, args); onk_op_end(op, nullptr, 0, nullptr, nullptr); onk_lane_sync(0); onk_op_elapsed_ms(op, &ms); }, 500));
  printf("bare launch+sync             %.2f us\n", us([&]{ tiny<<<8,128,0,st>>>(d); cudaStreamSynchronize(st); }, 500));
  return 0;
}
