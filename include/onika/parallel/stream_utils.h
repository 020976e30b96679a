// onika/parallel/stream_utils.h -- output streams the parallel layer reports on.
#pragma once
#include <iostream>

namespace onika { namespace parallel {
  inline std::ostream& log_out() { return std::cout; }
  inline std::ostream& log_err() { return std::cerr; }
} }
