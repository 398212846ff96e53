// compat stand-in, see filtering_stream.hpp
#include <boost/iostreams/filtering_stream.hpp>
