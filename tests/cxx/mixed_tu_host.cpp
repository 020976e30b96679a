// Host-compiler half of the mixed-compiler test (g++ -c, no nvcc): a plugin's host-only .cpp file that creates the default
// CUDA context and reads the device description.  The other half, mixed_tu_device.cu, is compiled by nvcc and must see the
// very same object (one layout for onika::cuda::CudaDevice / CudaContext in both compilers).
#include <onika/cuda/cuda_context.h>
#include <onika/parallel/parallel_execution_context.h>
#include <cstddef>

extern "C" void mixed_host_tu_probe(long out[8])
{
  auto* ctx = onika::cuda::CudaContext::default_cuda_ctx();      // built HERE, by host-compiled code
  out[0] = ( ctx != nullptr && ctx->has_devices() ) ? ctx->m_devices[0].m_deviceProp.multiProcessorCount : -1;
  out[1] = (long) sizeof(onika::cuda::CudaDevice);
  out[2] = (long) sizeof(onika::cuda::CudaContext);
  out[3] = (long) offsetof(onikaDeviceProp_t, multiProcessorCount);
  out[4] = (long) sizeof(onikaDeviceProp_t);
  out[5] = (long) sizeof(onikaDim3_t);
  out[6] = (long) sizeof(onika::parallel::ParallelExecutionContext);
  out[7] = (long) reinterpret_cast<std::uintptr_t>( ctx );
}
