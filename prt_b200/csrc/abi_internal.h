// abi_internal.h -- helpers shared by the translation units that implement the C ABI.
#pragma once
#include "../../include/prt_b200.h"
#include <cuda_runtime.h>
#include <string>

int prt_set_error(int code, const std::string &msg);   // records the thread-local message, returns code
cudaStream_t prt_ctx_stream(prt_ctx *);                // the context's own stream
int prt_ctx_sms(const prt_ctx *);
