// domain.cuh -- survivor stage (placeholder until the domain kernels land).
#pragma once
#include <cuda_runtime.h>
namespace pa {
struct DomainBufs {
  void release() {}
};
}  // namespace pa
