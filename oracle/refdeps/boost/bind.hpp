// Stand-in for <boost/bind.hpp> — TEST INFRASTRUCTURE.  std::bind under the boost name, plus the one boost-only feature
// the reference uses: relational operators on bind expressions (src/frame_handler_mono.cpp:505 sorts with
// boost::bind(&pair::second, _1) > boost::bind(&pair::second, _2)).
#ifndef PLSVO_REFDEPS_BOOST_BIND
#define PLSVO_REFDEPS_BOOST_BIND
#include <functional>
#include <utility>
namespace boost {
template <class F>
struct bind_t {
  F f;
  template <class... A>
  auto operator()(A&&... a) const -> decltype(f(std::forward<A>(a)...)) {
    return f(std::forward<A>(a)...);
  }
};
template <class... X>
auto bind(X&&... x) -> bind_t<decltype(std::bind(std::forward<X>(x)...))> {
  return bind_t<decltype(std::bind(std::forward<X>(x)...))>{std::bind(std::forward<X>(x)...)};
}
template <class A, class B>
struct bind_greater {
  A a;
  B b;
  template <class... T>
  bool operator()(T&&... t) const {
    return a(t...) > b(t...);
  }
};
template <class A, class B>
bind_greater<bind_t<A>, bind_t<B>> operator>(bind_t<A> a, bind_t<B> b) {
  return bind_greater<bind_t<A>, bind_t<B>>{a, b};
}
}  // namespace boost
using namespace std::placeholders;
#endif
