// Internal declarations shared by the host-side translation units of libhsolve_cuda.
#pragma once

#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hsolve_cuda.h"

struct hs_error : public std::runtime_error {
  int32_t code;
  hs_error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// records the message for hs_last_error() and returns the code
int32_t hs_fail(int32_t code, const std::string& msg);

#define HS_TRY_BEGIN try {
#define HS_TRY_END                                                        \
  }                                                                       \
  catch (const hs_error& e) { return hs_fail(e.code, e.what()); }         \
  catch (const std::bad_alloc&) { return hs_fail(HS_ENOMEM, "host allocation failed"); } \
  catch (const std::exception& e) { return hs_fail(HS_EARG, e.what()); }
