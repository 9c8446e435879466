#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::static_pointer_cast;
}  // namespace boost
