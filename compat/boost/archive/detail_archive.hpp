// compat/boost/archive/detail_archive.hpp — a small, self-describing-free
// archive pair behind Boost.Serialization's spelling (ar & x, ar << x, ar >> x).
//
// Boost is not installed here, and a boost::archive::binary_oarchive byte
// stream is not reproducible without it, so cdae_b200 defines its OWN cache
// format behind the same names: a magic line
// "cdae_b200-archive v1 <binary|text>\n" followed by the fields in declaration
// order — binary: raw little-endian scalars, u64 lengths; text: one
// whitespace-separated token per scalar, strings as "<len> <bytes>".
// Caches written by the reference's own Boost build cannot be read, and vice
// versa; INTEGRATION.md says so.
#ifndef CDAE_B200_COMPAT_BOOST_ARCHIVE_DETAIL_HPP_
#define CDAE_B200_COMPAT_BOOST_ARCHIVE_DETAIL_HPP_

#include "../../std_prelude.h"

#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include <boost/serialization/access.hpp>

namespace boost {
namespace archive {
namespace compat_detail {

[[noreturn]] inline void archive_fail(const char* what) {
  std::cerr << "cdae_b200 archive error: " << what << std::endl;
  std::abort();
}

template <bool Binary>
class oarchive_impl {
 public:
  typedef std::true_type is_saving;
  typedef std::false_type is_loading;
  explicit oarchive_impl(std::ostream& os) : os_(os) {
    os_ << "cdae_b200-archive v1 " << (Binary ? "binary" : "text") << "\n";
  }
  template <class T>
  oarchive_impl& operator<<(const T& t) {
    put(t);
    return *this;
  }
  template <class T>
  oarchive_impl& operator&(const T& t) {
    put(t);
    return *this;
  }
  // split_member() compiles BOTH halves for every archive (runtime branch on is_saving); the
  // load half of an output archive is never executed.
  template <class T>
  oarchive_impl& operator>>(T&) {
    archive_fail("operator>> on an output archive");
    return *this;
  }

 private:
  template <class T>
  typename std::enable_if<std::is_arithmetic<T>::value>::type put(const T& v) {
    if (Binary) {
      os_.write(reinterpret_cast<const char*>(&v), sizeof(T));
    } else {
      if (std::is_floating_point<T>::value) {
        os_.precision(17);
        os_ << v << ' ';
      } else {
        os_ << +v << ' ';
      }
    }
  }
  template <class T>
  typename std::enable_if<std::is_enum<T>::value>::type put(const T& v) {
    put(static_cast<std::int32_t>(v));
  }
  void put(const std::string& s) {
    put(static_cast<std::uint64_t>(s.size()));
    os_.write(s.data(), static_cast<std::streamsize>(s.size()));
    if (!Binary) os_ << ' ';
  }
  template <class A, class B>
  void put(const std::pair<A, B>& p) {
    put(p.first);
    put(p.second);
  }
  template <class T>
  void put(const std::vector<T>& v) {
    put(static_cast<std::uint64_t>(v.size()));
    for (const auto& e : v) put(e);
  }
  void put(const std::vector<bool>& v) {
    put(static_cast<std::uint64_t>(v.size()));
    for (bool e : v) put(static_cast<std::uint8_t>(e));
  }
  template <class K, class V, class... R>
  void put(const std::map<K, V, R...>& m) {
    put(static_cast<std::uint64_t>(m.size()));
    for (const auto& e : m) {
      put(e.first);
      put(e.second);
    }
  }
  template <class K, class V, class... R>
  void put(const std::unordered_map<K, V, R...>& m) {
    put(static_cast<std::uint64_t>(m.size()));
    for (const auto& e : m) {
      put(e.first);
      put(e.second);
    }
  }
  template <class T>
  void put(const boost::serialization::array_wrapper<T>& a) {
    for (unsigned long i = 0; i < a.n; ++i) put(a.ptr[i]);
  }
  template <class T>
  typename std::enable_if<std::is_class<T>::value>::type put(const T& t) {
    boost::serialization::access::serialize(*this, const_cast<T&>(t), 0u);
  }
  std::ostream& os_;
};

template <bool Binary>
class iarchive_impl {
 public:
  typedef std::false_type is_saving;
  typedef std::true_type is_loading;
  explicit iarchive_impl(std::istream& is) : is_(is) {
    std::string line;
    std::getline(is_, line);
    const std::string want =
        std::string("cdae_b200-archive v1 ") + (Binary ? "binary" : "text");
    if (line != want) archive_fail("bad magic (not a cdae_b200 cache, or wrong binary/text mode)");
  }
  template <class T>
  iarchive_impl& operator>>(T& t) {
    get(t);
    return *this;
  }
  template <class T>
  iarchive_impl& operator&(T& t) {
    get(const_cast<typename std::remove_const<T>::type&>(t));  // save half streams const members
    return *this;
  }
  template <class T>
  iarchive_impl& operator<<(const T&) {
    archive_fail("operator<< on an input archive");
    return *this;
  }

 private:
  template <class T>
  typename std::enable_if<std::is_arithmetic<T>::value>::type get(T& v) {
    if (Binary) {
      is_.read(reinterpret_cast<char*>(&v), sizeof(T));
    } else if (sizeof(T) == 1) {
      int tmp = 0;
      is_ >> tmp;
      v = static_cast<T>(tmp);
    } else {
      is_ >> v;
    }
    if (!is_) archive_fail("truncated stream");
  }
  template <class T>
  typename std::enable_if<std::is_enum<T>::value>::type get(T& v) {
    std::int32_t x = 0;
    get(x);
    v = static_cast<T>(x);
  }
  void get(std::string& s) {
    std::uint64_t n = 0;
    get(n);
    if (!Binary) is_.get();  // the single separator written after the length
    s.resize(static_cast<size_t>(n));
    if (n) is_.read(&s[0], static_cast<std::streamsize>(n));
    if (!is_) archive_fail("truncated string");
  }
  template <class A, class B>
  void get(std::pair<A, B>& p) {
    get(p.first);
    get(p.second);
  }
  template <class T>
  void get(std::vector<T>& v) {
    std::uint64_t n = 0;
    get(n);
    v.clear();
    v.resize(static_cast<size_t>(n));
    for (auto& e : v) get(e);
  }
  void get(std::vector<bool>& v) {
    std::uint64_t n = 0;
    get(n);
    v.assign(static_cast<size_t>(n), false);
    for (size_t i = 0; i < v.size(); ++i) {
      std::uint8_t b = 0;
      get(b);
      v[i] = b != 0;
    }
  }
  template <class K, class V, class... R>
  void get(std::map<K, V, R...>& m) {
    std::uint64_t n = 0;
    get(n);
    m.clear();
    for (std::uint64_t i = 0; i < n; ++i) {
      K k;
      V v;
      get(k);
      get(v);
      m.emplace(std::move(k), std::move(v));
    }
  }
  template <class K, class V, class... R>
  void get(std::unordered_map<K, V, R...>& m) {
    std::uint64_t n = 0;
    get(n);
    m.clear();
    for (std::uint64_t i = 0; i < n; ++i) {
      K k;
      V v;
      get(k);
      get(v);
      m.emplace(std::move(k), std::move(v));
    }
  }
  template <class T>
  void get(const boost::serialization::array_wrapper<T>& a) {
    for (unsigned long i = 0; i < a.n; ++i) get(a.ptr[i]);
  }
  template <class T>
  typename std::enable_if<std::is_class<T>::value>::type get(T& t) {
    boost::serialization::access::serialize(*this, t, 0u);
  }
  std::istream& is_;
};

}  // namespace compat_detail

typedef compat_detail::oarchive_impl<true> binary_oarchive;
typedef compat_detail::iarchive_impl<true> binary_iarchive;
typedef compat_detail::oarchive_impl<false> text_oarchive;
typedef compat_detail::iarchive_impl<false> text_iarchive;

}  // namespace archive
}  // namespace boost
#endif
