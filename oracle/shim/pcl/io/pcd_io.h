// oracle/shim/pcl/io/pcd_io.h — TEST INFRASTRUCTURE (PCD I/O is not exercised)
#pragma once
#include <string>
namespace pcl { namespace io {
template <typename C>
inline int loadPCDFile(const std::string&, C&) { return -1; }
}}
