// onika/cuda/cuda_error.h -- ONIKA_CU_CHECK_ERRORS: print file:line and abort on a CUDA error (the reference's policy).
#pragma once
#include <cstdio>
#include <cstdlib>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif

namespace onika { namespace cuda {
  inline void assert_success(int code, const char* what, const char* file, int line)
  {
    if (code != 0)
    {
#if defined(__CUDACC__)
      std::fprintf(stderr, "GPU error %d (%s) : %s @ %s:%d\n", code, cudaGetErrorString((cudaError_t)code), what, file, line);
#else
      std::fprintf(stderr, "GPU error %d : %s @ %s:%d\n", code, what, file, line);
#endif
      std::abort();
    }
  }
  // the reference's spelling (include/onika/cuda/cuda_error.h:34-41 there): the caller may ask not to abort
  inline void assertSuccess(int code, const char* file, int line, bool abort_on_failure = true)
  {
    if (code == 0) return;
    if (abort_on_failure) assert_success(code, "runtime call", file, line);
#if defined(__CUDACC__)
    std::fprintf(stderr, "GPU error %d (%s) @ %s:%d\n", code, cudaGetErrorString((cudaError_t)code), file, line);
#else
    std::fprintf(stderr, "GPU error %d @ %s:%d\n", code, file, line);
#endif
  }
} }
// a translation unit may bring its own policy by defining the macro before including this header, as in the reference
#ifndef ONIKA_CU_CHECK_ERRORS
#define ONIKA_CU_CHECK_ERRORS(expr) ::onika::cuda::assert_success( int(expr) , #expr , __FILE__ , __LINE__ )
#endif

// status of a C-ABI call (include/onika_b200.h) -> fatal error with the library's message
#define ONIKA_B200_CHECK(expr) do { if( (expr) != 0 ) { std::fprintf(stderr,"onika_b200: %s @ %s:%d\n", onk_last_error(), __FILE__, __LINE__); std::abort(); } } while(0)
