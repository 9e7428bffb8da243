// ORACLE shim (test infrastructure): boost::shared_ptr as the std one
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
}
