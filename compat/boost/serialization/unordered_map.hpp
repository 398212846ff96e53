// compat stand-in: container support lives in compat/boost/archive/detail_archive.hpp
#include <boost/serialization/access.hpp>
