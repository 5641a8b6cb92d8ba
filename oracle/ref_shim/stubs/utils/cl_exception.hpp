/*
 * Portable stand-in for the reference's utils/cl_exception.hpp (which relies on the
 * MSVC-only std::exception(const char*) constructor, cl_exception.hpp:113), found
 * first on the include path when oracle/build_ref.py compiles src/scene/scene.cpp.
 * Same two names, same behaviour: throw on non-zero status.  TEST INFRASTRUCTURE.
 */
#pragma once
#include <stdexcept>
#include <string>

class CLException : public std::runtime_error
{
public:
    CLException(const char* msg, int status)
        : std::runtime_error(std::string(msg) + " (status " + std::to_string(status) + ")") {}
};

inline void ThrowIfFailed(int status, const char* msg)
{
    if (status != 0) throw CLException(msg, status);
}
