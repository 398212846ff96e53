// compat/boost/tokenizer.hpp — stand-in for boost::tokenizer<char_separator<char>>
// as used by libcf::split_line: split on ANY character of the delimiter string
// and DROP empty tokens (char_separator's default drop_empty_tokens policy).
// Known answer pinned by the reference: test/file_test.hpp:14-23.
#ifndef CDAE_B200_COMPAT_BOOST_TOKENIZER_HPP_
#define CDAE_B200_COMPAT_BOOST_TOKENIZER_HPP_
#include <string>
#include <vector>
namespace boost {
template <class Char>
class char_separator {
 public:
  explicit char_separator(const Char* dropped_delims) : delims_(dropped_delims) {}
  bool is_delim(Char c) const { return delims_.find(c) != std::basic_string<Char>::npos; }

 private:
  std::basic_string<Char> delims_;
};
template <class Separator>
class tokenizer {
 public:
  typedef std::vector<std::string>::const_iterator iterator;
  tokenizer(const std::string& s, const Separator& sep) {
    std::string cur;
    for (char c : s) {
      if (sep.is_delim(c)) {
        if (!cur.empty()) tokens_.push_back(cur);
        cur.clear();
      } else {
        cur.push_back(c);
      }
    }
    if (!cur.empty()) tokens_.push_back(cur);
  }
  iterator begin() const { return tokens_.begin(); }
  iterator end() const { return tokens_.end(); }

 private:
  std::vector<std::string> tokens_;
};
}  // namespace boost
#endif
