// Stand-in for <boost/noncopyable.hpp> — TEST INFRASTRUCTURE.
#ifndef PLSVO_REFDEPS_BOOST_NONCOPYABLE
#define PLSVO_REFDEPS_BOOST_NONCOPYABLE
namespace boost {
class noncopyable {
 protected:
  noncopyable() {}
  ~noncopyable() {}
 private:
  noncopyable(const noncopyable&);
  noncopyable& operator=(const noncopyable&);
};
}  // namespace boost
#endif
