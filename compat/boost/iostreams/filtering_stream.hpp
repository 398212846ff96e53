// compat/boost/iostreams/filtering_stream.hpp — stand-in for the one pattern
// libcf uses (base/io/file.hpp:128-144): push(gzip filter); push(fstream);
// then stream an archive through it.  Output is buffered and (de)compressed
// with zlib in one shot when the stream is destroyed / first read.
#ifndef CDAE_B200_COMPAT_BOOST_IOSTREAMS_FILTERING_STREAM_HPP_
#define CDAE_B200_COMPAT_BOOST_IOSTREAMS_FILTERING_STREAM_HPP_

#include "../../std_prelude.h"

#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#if defined(__has_include)
#if __has_include(<zlib.h>) && !defined(CDAE_COMPAT_NO_ZLIB)
#include <zlib.h>
#define CDAE_COMPAT_HAVE_ZLIB 1
#endif
#endif

namespace boost {
namespace iostreams {

struct input {};
struct output {};
struct gzip_compressor {};
struct gzip_decompressor {};

namespace compat_detail {
[[noreturn]] inline void fail(const char* what) {
  std::cerr << "cdae_b200 iostreams error: " << what << std::endl;
  std::abort();
}
inline std::string gz_deflate(const std::string& in) {
#ifdef CDAE_COMPAT_HAVE_ZLIB
  z_stream zs{};
  if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK)
    fail("deflateInit2");
  std::string out;
  std::vector<unsigned char> buf(1 << 16);
  zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in.data()));
  zs.avail_in = static_cast<uInt>(in.size());
  int rc;
  do {
    zs.next_out = buf.data();
    zs.avail_out = static_cast<uInt>(buf.size());
    rc = deflate(&zs, Z_FINISH);
    out.append(reinterpret_cast<char*>(buf.data()), buf.size() - zs.avail_out);
  } while (rc == Z_OK);
  deflateEnd(&zs);
  if (rc != Z_STREAM_END) fail("deflate");
  return out;
#else
  return in;
#endif
}
inline std::string gz_inflate(const std::string& in) {
#ifdef CDAE_COMPAT_HAVE_ZLIB
  z_stream zs{};
  if (inflateInit2(&zs, 15 + 32) != Z_OK) fail("inflateInit2");
  std::string out;
  std::vector<unsigned char> buf(1 << 16);
  zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in.data()));
  zs.avail_in = static_cast<uInt>(in.size());
  int rc;
  do {
    zs.next_out = buf.data();
    zs.avail_out = static_cast<uInt>(buf.size());
    rc = inflate(&zs, Z_NO_FLUSH);
    if (rc != Z_OK && rc != Z_STREAM_END) {
      inflateEnd(&zs);
      fail("inflate (not a gzip stream?)");
    }
    out.append(reinterpret_cast<char*>(buf.data()), buf.size() - zs.avail_out);
  } while (rc != Z_STREAM_END);
  inflateEnd(&zs);
  return out;
#else
  return in;
#endif
}
}  // namespace compat_detail

template <class Mode>
class filtering_stream;

template <>
class filtering_stream<output> : public std::ostream {
 public:
  filtering_stream() : std::ostream(&buf_) {}
  ~filtering_stream() { finish(); }
  void push(const gzip_compressor&) { gzip_ = true; }
  void push(std::ostream& sink) { sink_ = &sink; }
  void reset() { finish(); }

 private:
  void finish() {
    if (!sink_) return;
    const std::string raw = buf_.str();
    const std::string z = gzip_ ? compat_detail::gz_deflate(raw) : raw;
    sink_->write(z.data(), static_cast<std::streamsize>(z.size()));
    sink_->flush();
    sink_ = nullptr;
  }
  std::stringbuf buf_;
  std::ostream* sink_ = nullptr;
  bool gzip_ = false;
};

template <>
class filtering_stream<input> : public std::istream {
 public:
  filtering_stream() : std::istream(&buf_) {}
  void push(const gzip_decompressor&) { gzip_ = true; }
  void push(std::istream& src) {
    std::ostringstream ss;
    ss << src.rdbuf();
    const std::string raw = ss.str();
    buf_.str(gzip_ ? compat_detail::gz_inflate(raw) : raw);
  }

 private:
  std::stringbuf buf_;
  bool gzip_ = false;
};

typedef filtering_stream<input> filtering_istream;
typedef filtering_stream<output> filtering_ostream;

}  // namespace iostreams
}  // namespace boost
#endif
