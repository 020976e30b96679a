// onika/extras/array2d.h -- dense row-major 2D array of doubles in host/device visible memory (Array2D) and the
// trivially copyable view of it that functors carry into kernels (Array2DReference).
// Both expose data() / operator[](row) / rows() / columns(); the row arithmetic lives in one CRTP mixin.
#pragma once
#include <cstddef>
#include <onika/cuda/cuda.h>
#include <onika/cuda/stl_adaptors.h>
#include <onika/memory/allocator.h>

namespace onika { namespace extras {

  namespace array2d_details
  {
    // Derived provides base_ptr() and the m_rows / m_columns members
    template<class Derived, class ElemPtr>
    struct RowMajor
    {
      ONIKA_HOST_DEVICE_FUNC inline size_t rows() const { return self().m_rows; }
      ONIKA_HOST_DEVICE_FUNC inline size_t columns() const { return self().m_columns; }
      ONIKA_HOST_DEVICE_FUNC inline ElemPtr row_ptr(size_t r) const { return self().base_ptr() + self().m_columns * r; }
    private:
      ONIKA_HOST_DEVICE_FUNC inline const Derived& self() const { return static_cast<const Derived&>(*this); }
    };
  }

  // owning container: managed memory, so the same pointer is valid in kernels and on the host
  struct Array2D : public array2d_details::RowMajor<Array2D,double*>
  {
    onika::memory::CudaMMVector<double> m_data;
    size_t m_rows = 0;
    size_t m_columns = 0;

    inline void resize(size_t r, size_t c) { m_data.resize( r * c ); m_rows = r; m_columns = c; }
    ONIKA_HOST_DEVICE_FUNC inline double* base_ptr() const { return const_cast<double*>( onika::cuda::vector_data( m_data ) ); }
    ONIKA_HOST_DEVICE_FUNC inline double* data() { return base_ptr(); }
    ONIKA_HOST_DEVICE_FUNC inline const double* data() const { return base_ptr(); }
    ONIKA_HOST_DEVICE_FUNC inline double* operator [] (size_t r) { return row_ptr(r); }
    ONIKA_HOST_DEVICE_FUNC inline const double* operator [] (size_t r) const { return row_ptr(r); }
  };

  // non-owning, trivially copyable: what BlockParallelValueAddFunctor & co. hold
  struct Array2DReference : public array2d_details::RowMajor<Array2DReference,double*>
  {
    double * const m_data_ptr = nullptr;
    size_t const m_rows = 0;
    size_t const m_columns = 0;

    Array2DReference() = default;
    Array2DReference(const Array2DReference&) = default;
    Array2DReference(Array2DReference&&) = default;
    ONIKA_HOST_DEVICE_FUNC inline Array2DReference(Array2D& a) : m_data_ptr( a.data() ) , m_rows( a.rows() ) , m_columns( a.columns() ) {}
    ONIKA_HOST_DEVICE_FUNC inline double* base_ptr() const { return m_data_ptr; }
    ONIKA_HOST_DEVICE_FUNC inline double* operator [] (size_t r) const { return row_ptr(r); }
  };

} }
