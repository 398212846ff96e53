// compat/gflags/gflags.h — stand-in for the part of gflags libcf uses:
// DEFINE_{string,int32,int64,uint64,double,bool} -> FLAGS_name globals, and
// gflags::ParseCommandLineFlags / SetUsageMessage.  gflags is not installed
// in this image.  Accepted syntax (what apps/yelp/cdae.sh passes):
// --name=value, --name value, -name=value, --boolflag, --noboolflag, with
// bool values true/false/1/0/t/f/yes/no/y/n; "--" ends flag parsing.
#ifndef CDAE_B200_COMPAT_GFLAGS_GFLAGS_H_
#define CDAE_B200_COMPAT_GFLAGS_GFLAGS_H_

#include "../std_prelude.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

namespace gflags {

typedef std::int32_t int32;
typedef std::int64_t int64;
typedef std::uint64_t uint64;

namespace compat_detail {
enum Kind { K_BOOL, K_INT32, K_INT64, K_UINT64, K_DOUBLE, K_STRING };
struct Entry {
  Kind kind;
  void* ptr;
  const char* help;
};
inline std::map<std::string, Entry>& registry() {
  static std::map<std::string, Entry> r;
  return r;
}
inline std::string& usage() {
  static std::string u;
  return u;
}
struct Registerer {
  Registerer(const char* name, Kind k, void* p, const char* help) {
    registry()[name] = Entry{k, p, help};
  }
};
inline bool parse_bool(const std::string& v, bool* out) {
  static const char* t[] = {"1", "t", "true", "y", "yes", "True", "TRUE"};
  static const char* f[] = {"0", "f", "false", "n", "no", "False", "FALSE"};
  for (auto s : t)
    if (v == s) return *out = true, true;
  for (auto s : f)
    if (v == s) return *out = false, true;
  return false;
}
inline bool assign(const Entry& e, const std::string& v) {
  char* end = nullptr;
  switch (e.kind) {
    case K_BOOL: {
      bool b;
      if (!parse_bool(v, &b)) return false;
      *static_cast<bool*>(e.ptr) = b;
      return true;
    }
    case K_INT32:
      *static_cast<int32*>(e.ptr) = static_cast<int32>(std::strtol(v.c_str(), &end, 0));
      return end && *end == 0 && !v.empty();
    case K_INT64:
      *static_cast<int64*>(e.ptr) = std::strtoll(v.c_str(), &end, 0);
      return end && *end == 0 && !v.empty();
    case K_UINT64:
      *static_cast<uint64*>(e.ptr) = std::strtoull(v.c_str(), &end, 0);
      return end && *end == 0 && !v.empty();
    case K_DOUBLE:
      *static_cast<double*>(e.ptr) = std::strtod(v.c_str(), &end);
      return end && *end == 0 && !v.empty();
    case K_STRING:
      *static_cast<std::string*>(e.ptr) = v;
      return true;
  }
  return false;
}
[[noreturn]] inline void die(const std::string& msg) {
  std::fprintf(stderr, "ERROR: %s\n", msg.c_str());
  std::exit(1);
}
}  // namespace compat_detail

inline void SetUsageMessage(const std::string& u) { compat_detail::usage() = u; }

// Returns the index of the first non-flag argument (like gflags); when
// remove_flags is true argv is compacted to argv[0] + the non-flag arguments.
inline uint32_t ParseCommandLineFlags(int* argc, char*** argv, bool remove_flags) {
  using namespace compat_detail;
  int n = *argc;
  char** av = *argv;
  int keep = 1;
  int first_nonflag = n;
  for (int i = 1; i < n; ++i) {
    const char* a = av[i];
    if (a[0] != '-' || a[1] == 0) {  // positional
      if (first_nonflag == n) first_nonflag = i;
      av[keep++] = av[i];
      continue;
    }
    if (std::strcmp(a, "--") == 0) {
      for (int j = i + 1; j < n; ++j) av[keep++] = av[j];
      break;
    }
    std::string s(a + (a[1] == '-' ? 2 : 1));
    std::string name = s, value;
    bool has_value = false;
    size_t eq = s.find('=');
    if (eq != std::string::npos) {
      name = s.substr(0, eq);
      value = s.substr(eq + 1);
      has_value = true;
    }
    auto it = registry().find(name);
    if (it == registry().end() && !has_value && name.compare(0, 2, "no") == 0) {
      auto it2 = registry().find(name.substr(2));
      if (it2 != registry().end() && it2->second.kind == K_BOOL) {
        *static_cast<bool*>(it2->second.ptr) = false;
        continue;
      }
    }
    if (it == registry().end()) die("unknown command line flag '" + name + "'");
    if (!has_value) {
      if (it->second.kind == K_BOOL) {
        *static_cast<bool*>(it->second.ptr) = true;
        continue;
      }
      if (i + 1 >= n) die("flag '" + name + "' is missing its argument");
      value = av[++i];
    }
    if (!assign(it->second, value))
      die("illegal value '" + value + "' specified for flag '" + name + "'");
  }
  if (remove_flags) {
    *argc = keep;
    return 1;
  }
  return static_cast<uint32_t>(first_nonflag);
}

inline void ShutDownCommandLineFlags() {}

}  // namespace gflags

namespace google {
using gflags::ParseCommandLineFlags;
using gflags::SetUsageMessage;
}  // namespace google

#define CDAE_COMPAT_DEFINE_FLAG(type, kind, name, val, txt)                             \
  type FLAGS_##name = val;                                                              \
  static ::gflags::compat_detail::Registerer cdae_compat_flagreg_##name(                \
      #name, ::gflags::compat_detail::kind, &FLAGS_##name, txt)

#define DEFINE_bool(name, val, txt) CDAE_COMPAT_DEFINE_FLAG(bool, K_BOOL, name, val, txt)
#define DEFINE_int32(name, val, txt) \
  CDAE_COMPAT_DEFINE_FLAG(::gflags::int32, K_INT32, name, val, txt)
#define DEFINE_int64(name, val, txt) \
  CDAE_COMPAT_DEFINE_FLAG(::gflags::int64, K_INT64, name, val, txt)
#define DEFINE_uint64(name, val, txt) \
  CDAE_COMPAT_DEFINE_FLAG(::gflags::uint64, K_UINT64, name, val, txt)
#define DEFINE_double(name, val, txt) CDAE_COMPAT_DEFINE_FLAG(double, K_DOUBLE, name, val, txt)
#define DEFINE_string(name, val, txt) \
  CDAE_COMPAT_DEFINE_FLAG(std::string, K_STRING, name, val, txt)

#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_int32(name) extern ::gflags::int32 FLAGS_##name
#define DECLARE_int64(name) extern ::gflags::int64 FLAGS_##name
#define DECLARE_uint64(name) extern ::gflags::uint64 FLAGS_##name
#define DECLARE_double(name) extern double FLAGS_##name
#define DECLARE_string(name) extern std::string FLAGS_##name

#endif  // CDAE_B200_COMPAT_GFLAGS_GFLAGS_H_
