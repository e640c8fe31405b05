// Stand-in for <boost/function.hpp> (std::function under the boost name) — TEST INFRASTRUCTURE.
#ifndef PLSVO_REFDEPS_BOOST_FUNCTION
#define PLSVO_REFDEPS_BOOST_FUNCTION
#include <functional>
#include "bind.hpp"  // the real boost headers pull boost::bind and its placeholders in transitively
namespace boost {
template <class S>
using function = std::function<S>;
}
#endif
