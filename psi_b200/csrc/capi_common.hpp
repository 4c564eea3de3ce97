// capi_common.hpp -- error translation shared by the two halves of the C-ABI (capi_host.cpp: graph / paths / reads,
// no CUDA; capi.cu: the device entry points).
#ifndef PSI_B200_CAPI_COMMON_HPP
#define PSI_B200_CAPI_COMMON_HPP

#include <exception>
#include <new>
#include <stdexcept>
#include <string>

#include "../../include/psi_b200.h"

namespace psi_b200 {

// host-side error classes (the device engine's live in device/context.hpp and derive from std::runtime_error too)
std::string& capi_global_error();      // thread local

}  // namespace psi_b200

#define HOST_GUARD(body)                         \
  try { body; return PSI_B200_OK; }              \
  catch (const std::bad_alloc&) { ::psi_b200::capi_global_error() = "out of host memory"; return PSI_B200_ERR_NOMEM; } \
  catch (const std::exception& e) { ::psi_b200::capi_global_error() = e.what(); return PSI_B200_ERR_ARG; } \
  catch (...) { ::psi_b200::capi_global_error() = "unknown error"; return PSI_B200_ERR_ARG; }

#endif
