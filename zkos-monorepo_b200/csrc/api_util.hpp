// Entry-point boilerplate shared by the C-ABI translation units: exceptions -> return codes + thread-local message.
//   API_TRY    ... API_END : no lock, no device (state queries, the prover's own locking)
//   API_BEGIN  ... API_END : a primitive call on the primary device, serialised on that device's mutex
#pragma once
#include "../../include/zkgpu.h"
#include "context.cuh"

#define API_TRY try {
#define API_BEGIN try { zk::DeviceScope api_scope_(zk::rt().primary());
#define API_END                                                               \
    return ZKGPU_OK; }                                                        \
    catch (const zk::Error& e) { zk::g_last_error = e.what(); return e.code; } \
    catch (const std::exception& e) { zk::g_last_error = e.what(); return ZKGPU_ERR_INTERNAL; }
