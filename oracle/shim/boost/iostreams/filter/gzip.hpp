// TEST INFRASTRUCTURE ONLY (oracle build shim).  gzip_decompressor lives in
// filtering_streambuf.hpp of this shim (dxyWindow.cpp:19 includes this header).
#ifndef PGT_ORACLE_SHIM_GZIP_HPP
#define PGT_ORACLE_SHIM_GZIP_HPP
#include <boost/iostreams/filtering_streambuf.hpp>
#endif
