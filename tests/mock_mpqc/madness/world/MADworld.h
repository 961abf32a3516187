// MOCK: see ../../tiledarray.h
#pragma once
#include <tiledarray.h>
