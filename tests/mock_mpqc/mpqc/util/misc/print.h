// MOCK of mpqc/util/misc/print.h (the real one pulls units/exenv): print_par as ccsd_t.h calls it.
#pragma once
#include <tiledarray.h>
namespace mpqc {
namespace utility {
template <typename... Args>
void print_par(madness::World&, Args&&...) {}
}  // namespace utility
}  // namespace mpqc
