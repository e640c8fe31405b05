// Stand-in for <boost/shared_ptr.hpp> (std::shared_ptr under the boost name) — TEST INFRASTRUCTURE.
#ifndef PLSVO_REFDEPS_BOOST_SHARED_PTR
#define PLSVO_REFDEPS_BOOST_SHARED_PTR
#include <memory>
namespace boost {
template <class T>
using shared_ptr = std::shared_ptr<T>;
}
#endif
