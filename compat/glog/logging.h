// compat/glog/logging.h — stand-in for the part of google-glog libcf uses:
// LOG(INFO|WARNING|ERROR|FATAL) streams, CHECK / CHECK_{EQ,NE,LT,LE,GT,GE}
// (stream-able, abort on failure), FLAGS_log_dir, google::InitGoogleLogging,
// google::SetLogDestination.  glog itself is not installed in this image.
//
// Behaviour kept from glog: FATAL (and a failed CHECK) prints the message and
// aborts the process — that is the reference's whole error convention
// (SURVEY.md §8b "Errors").  INFO lines go to the destination file when one
// was set and could be opened, else to stderr; GLOG_minloglevel=N in the
// environment drops lines below severity N.
#ifndef CDAE_B200_COMPAT_GLOG_LOGGING_H_
#define CDAE_B200_COMPAT_GLOG_LOGGING_H_

#include "../std_prelude.h"

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <sstream>
#include <string>

namespace google {

typedef int LogSeverity;
const LogSeverity GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3;
const LogSeverity INFO = GLOG_INFO, WARNING = GLOG_WARNING, ERROR = GLOG_ERROR, FATAL = GLOG_FATAL;

namespace compat_detail {
struct State {
  std::FILE* sink = nullptr;  // nullptr -> stderr
  int min_level = 0;
  State() {
    if (const char* e = std::getenv("GLOG_minloglevel")) min_level = std::atoi(e);
  }
};
inline State& state() {
  static State s;
  return s;
}
}  // namespace compat_detail

inline void InitGoogleLogging(const char* /*argv0*/) {}
inline void ShutdownGoogleLogging() {}
inline void SetLogDestination(LogSeverity /*severity*/, const char* base_filename) {
  auto& st = compat_detail::state();
  if (st.sink) {
    std::fclose(st.sink);
    st.sink = nullptr;
  }
  if (base_filename && *base_filename) st.sink = std::fopen(base_filename, "a");
}

class LogMessage {
 public:
  LogMessage(const char* file, int line, LogSeverity sev) : sev_(sev) {
    static const char tag[] = {'I', 'W', 'E', 'F'};
    const char* base = file;
    for (const char* p = file; *p; ++p)
      if (*p == '/') base = p + 1;
    ss_ << tag[sev < 0 ? 0 : (sev > 3 ? 3 : sev)] << ' ' << base << ':' << line << "] ";
  }
  ~LogMessage() {
    auto& st = compat_detail::state();
    if (sev_ >= st.min_level || sev_ >= GLOG_FATAL) {
      std::string s = ss_.str();
      if (s.empty() || s.back() != '\n') s.push_back('\n');
      std::FILE* out = (st.sink && sev_ < GLOG_ERROR) ? st.sink : stderr;
      std::fwrite(s.data(), 1, s.size(), out);
      std::fflush(out);
    }
    if (sev_ >= GLOG_FATAL) std::abort();
  }
  std::ostream& stream() { return ss_; }

 private:
  std::ostringstream ss_;
  LogSeverity sev_;
};

// Makes `cond ? (void)0 : Voidify() & stream << ...` type-check.
struct LogMessageVoidify {
  void operator&(std::ostream&) {}
};

}  // namespace google

#define LOG(severity) ::google::LogMessage(__FILE__, __LINE__, ::google::GLOG_##severity).stream()
#define LOG_IF(severity, cond) \
  !(cond) ? (void)0 : ::google::LogMessageVoidify() & LOG(severity)

#define CHECK(cond)            \
  (cond) ? (void)0             \
         : ::google::LogMessageVoidify() & LOG(FATAL) << "Check failed: " #cond " "

#define CDAE_COMPAT_CHECK_OP(a, b, op)                                                      \
  ((a)op(b)) ? (void)0                                                                      \
             : ::google::LogMessageVoidify() &                                              \
                   LOG(FATAL) << "Check failed: " #a " " #op " " #b " (" << (a) << " vs. " \
                              << (b) << ") "
#define CHECK_EQ(a, b) CDAE_COMPAT_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) CDAE_COMPAT_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) CDAE_COMPAT_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) CDAE_COMPAT_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) CDAE_COMPAT_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) CDAE_COMPAT_CHECK_OP(a, b, >=)
#define CHECK_NOTNULL(p) (p)
#define DCHECK(cond) CHECK(cond)

// glog exports its own flags; libcf's app assigns FLAGS_log_dir.
namespace google {
namespace compat_detail {
inline std::string& log_dir() {
  static std::string s;
  return s;
}
}  // namespace compat_detail
}  // namespace google
#define FLAGS_log_dir (::google::compat_detail::log_dir())

#endif  // CDAE_B200_COMPAT_GLOG_LOGGING_H_
