// onika/log.h -- lout / lerr / ldbg and fatal_error().  Errors on the hot path have no error codes and no
// exceptions in onika: fatal_error() << "..." << std::endl; prints and aborts.  The B200 front end turns any
// non-zero status of the C ABI (include/onika_b200.h) into exactly that.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace onika
{
  struct LogStream
  {
    std::ostream* m_out = &std::cout;
    bool m_enabled = true;
    template<class T> inline LogStream& operator << (const T& x) { if (m_enabled) (*m_out) << x; return *this; }
    inline LogStream& operator << (std::ostream& (*manip)(std::ostream&)) { if (m_enabled) manip(*m_out); return *this; }
  };
  inline LogStream lout { &std::cout, true };
  inline LogStream lerr { &std::cerr, true };
  inline LogStream ldbg { &std::cout, false };

  // collects the message, prints it and aborts when the temporary dies (end of the full expression)
  struct FatalErrorStream
  {
    std::ostringstream m_msg;
    FatalErrorStream() = default;
    FatalErrorStream(FatalErrorStream&& o) : m_msg(std::move(o.m_msg)) {}
    template<class T> inline FatalErrorStream& operator << (const T& x) { m_msg << x; return *this; }
    inline FatalErrorStream& operator << (std::ostream& (*manip)(std::ostream&)) { manip(m_msg); return *this; }
    [[noreturn]] inline ~FatalErrorStream()
    {
      std::cerr << "Fatal error: " << m_msg.str() << std::endl << std::flush;
      std::abort();
    }
  };
  inline FatalErrorStream fatal_error() { return FatalErrorStream{}; }
}
