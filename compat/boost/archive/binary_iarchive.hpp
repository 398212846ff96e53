// compat stand-in, see detail_archive.hpp
#include <boost/archive/detail_archive.hpp>
