// MOCK of util/misc/time.h:12-35
#pragma once
#include <chrono>
#include <tiledarray.h>
namespace mpqc {
using time_point = std::chrono::high_resolution_clock::time_point;
inline time_point fenced_now(madness::World&) { return std::chrono::high_resolution_clock::now(); }
inline double duration_in_s(time_point const& a, time_point const& b) { return std::chrono::duration<double>(b - a).count(); }
}  // namespace mpqc
