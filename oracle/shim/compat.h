// compat.h — force-included (-include) in front of every reference translation unit when the
// compiled-reference oracle is built with g++ on Linux. TEST INFRASTRUCTURE ONLY.
// The reference was written against MSVC and relies on transitive includes and a few MSVC-isms:
//   uint32_t without <cstdint> (geometry.h:264), FLT_MAX/FLT_EPSILON via -DkInfinity/-DkEpsilon
//   (cmakelists.txt:61-62), std::queue (integrator.h:311), std::unordered_map / std::unique_ptr
//   (scene.h:43-46), std::sqrtf / std::atanf (light.h:139,160; examples/vpt.cpp:43).
#pragma once
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <ctime>
#include <filesystem>
#include <iomanip>
#include <memory>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
namespace std {
using ::atanf;
using ::sqrtf;
} // namespace std
