// TEST INFRASTRUCTURE ONLY (oracle build shim).  dxyWindow.cpp:18 includes
// <boost/iostreams/copy.hpp> but never calls boost::iostreams::copy; an empty
// header is sufficient.  See filtering_streambuf.hpp in this directory.
#ifndef PGT_ORACLE_SHIM_COPY_HPP
#define PGT_ORACLE_SHIM_COPY_HPP
#endif
