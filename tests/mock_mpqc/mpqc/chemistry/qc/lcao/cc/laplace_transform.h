// MOCK of cc/laplace_transform.h: nothing of it is instantiated by the GPU path; ccsd_t.h's Laplace code only names
// its function templates in dependent calls.
#pragma once
#include <tiledarray.h>
#include "mpqc/math/external/eigen/eigen.h"
