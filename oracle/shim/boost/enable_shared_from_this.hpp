// oracle/shim/boost — TEST INFRASTRUCTURE: boost smart pointers mapped onto std
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::enable_shared_from_this;
using std::dynamic_pointer_cast;
using std::make_shared;
}  // namespace boost
