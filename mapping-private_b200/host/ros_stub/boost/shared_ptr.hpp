#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
}  // namespace boost
