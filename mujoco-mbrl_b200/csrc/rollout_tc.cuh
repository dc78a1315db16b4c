// tcgen05 tensor-core rollout engine (placeholder until the kernel lands).
#pragma once
#include <string>

#include "common.cuh"
#include "philox.cuh"

namespace mbrl {

struct TcModel {
  int ready = 0;
};

inline bool tc_init(TcModel*, int, int, int, bool, size_t, std::string* why) {
  *why = "not built yet";
  return false;
}
inline bool tc_set_weights(TcModel*, const float*, const float*, const float*, const float*, const float*,
                           const float*, std::string* why) {
  *why = "not built yet";
  return false;
}
inline void tc_free(TcModel*) {}
inline cudaError_t tc_launch_rollout(TcModel*, const ModelDev&, const ActionSource&, const Shape&, const float*,
                                     float*, float*, float*, int, cudaStream_t) {
  return cudaErrorNotSupported;
}

}  // namespace mbrl
