// Stand-in for <boost/bind.hpp> (std::bind and its placeholders) — TEST INFRASTRUCTURE.
#ifndef PLSVO_REFDEPS_BOOST_BIND
#define PLSVO_REFDEPS_BOOST_BIND
#include <functional>
namespace boost {
using std::bind;
}
using namespace std::placeholders;
#endif
