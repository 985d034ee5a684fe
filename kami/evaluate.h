// kami::eval (kami/evaluate.h:6) -- arena gating of a candidate network.  Drives the same
// select / infer / expand kernels with two weight sets; SURVEY.md 8(f) #2, not built yet.
#pragma once
#include <stdexcept>

#include "nn/nn.h"

namespace kami {
inline bool eval(NN*, NN*, int) { throw std::runtime_error("kami::eval is not built yet (SURVEY.md 8(f) #2)"); }
}  // namespace kami
