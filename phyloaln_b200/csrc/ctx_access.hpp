// ctx_access.hpp -- the few pa_ctx fields the other translation units of the library need
// (pa_ctx itself is defined in api.cu).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "../../include/phyloaln_b200.h"

cudaStream_t pa_ctx_stream(pa_ctx *ctx);
int pa_ctx_device(pa_ctx *ctx);
int pa_ctx_num_sms(pa_ctx *ctx);
void pa_ctx_set_error(pa_ctx *ctx, const std::string &msg);
void pa_ctx_count_launches(pa_ctx *ctx, int n);
