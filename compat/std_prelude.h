// compat/std_prelude.h — the standard headers the real glog / Boost / Eigen headers pull in
// transitively and the reference's sources silently rely on (e.g. base/io/file.hpp uses
// std::unique_ptr without <memory>).  Included by every stand-in entry header.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
