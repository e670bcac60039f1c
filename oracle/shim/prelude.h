/* TEST INFRASTRUCTURE - force-included prelude for building the UNMODIFIED reference
 * hot-path translation units (read in place from /root/reference) with g++ on Linux.
 * Not reference code, not product code.  See SURVEY.md §8(c), H5.
 *
 * 1. <math.h>/<stdlib.h> FIRST: gives the global-namespace float overloads of
 *    abs/sqrt/cos/sin/atan that MSVC resolves unqualified calls to
 *    (UtilityFunctions.cpp:330 `abs(dot)`, ECMDataTypes.h:35 `sqrt`, ORCA.cpp:376 `atan`).
 * 2. headers MSVC/Boost pull in transitively.
 * 3. std::powf (Simulator.cpp:277) is missing from libstdc++ 13.
 * 4. Timer.h assigns high_resolution_clock::now() to a steady_clock time_point
 *    (legal only on MSVC where the two are the same type).
 */
#pragma once
#include <math.h>
#include <stdlib.h>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>
#include <limits>
#include <tuple>
#include <chrono>
#include <iostream>
#include <unordered_map>
#include <map>
#include <stack>
#include <queue>
#include <memory>
#include <random>
namespace std { using ::powf; }
#define high_resolution_clock steady_clock
