// api_guard.h — "the C ABI never throws" (include/b200rt.h), enforced: every int-returning extern "C" entry point is
// a function-try-block that ends in B200RT_CATCH, so a std::bad_alloc from a growing vector (or anything else) becomes
// an error code and a message instead of crossing the language boundary into a C / Go / Java host.
#pragma once
#include <exception>
#include <new>

#include "../../include/b200rt.h"

extern "C" void b200rt_internal_set_error(const char* msg);

#define B200RT_CATCH                                                                                           \
  catch (const std::bad_alloc&) { b200rt_internal_set_error("out of host memory"); return B200RT_E_NOMEM; }     \
  catch (const std::exception& e_) { b200rt_internal_set_error(e_.what()); return B200RT_E_INVALID; }           \
  catch (...) { b200rt_internal_set_error("unknown C++ exception"); return B200RT_E_INVALID; }
