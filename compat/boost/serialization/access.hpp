// compat/boost/serialization/access.hpp — the friend hook libcf's classes
// declare, plus split_member.  The archives in compat/boost/archive call
// T::serialize(ar, version) through this class.
#ifndef CDAE_B200_COMPAT_BOOST_SERIALIZATION_ACCESS_HPP_
#define CDAE_B200_COMPAT_BOOST_SERIALIZATION_ACCESS_HPP_
#include <type_traits>
namespace boost {
namespace serialization {
class access {
 public:
  template <class Archive, class T>
  static void serialize(Archive& ar, T& t, const unsigned int version) {
    t.serialize(ar, version);
  }
  template <class Archive, class T>
  static void member_save(Archive& ar, const T& t, const unsigned int version) {
    t.save(ar, version);
  }
  template <class Archive, class T>
  static void member_load(Archive& ar, T& t, const unsigned int version) {
    t.load(ar, version);
  }
};
template <class Archive, class T>
inline void split_member(Archive& ar, T& t, const unsigned int version) {
  if (Archive::is_saving::value)
    access::member_save(ar, t, version);
  else
    access::member_load(ar, t, version);
}
template <class T>
struct array_wrapper {
  T* ptr;
  unsigned long n;
};
template <class T>
inline array_wrapper<T> make_array(T* p, unsigned long n) {
  return array_wrapper<T>{p, n};
}
}  // namespace serialization
}  // namespace boost
#endif
