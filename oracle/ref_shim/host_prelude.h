/*
 * Force-included (-include) ahead of every reference HOST source compiled by
 * oracle/build_ref.py: standard headers and std:: names that MSVC provides
 * implicitly and the reference relies on (e.g. std::powf in hdr_loader.cpp:105,
 * std::cosf/sinf in camera code).  TEST INFRASTRUCTURE.
 */
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
namespace std { using ::powf; using ::cosf; using ::sinf; }
