// fuse.cpp -- FAST mode planner (pattern-matched fused plan).  Placeholder until the fused kernels land.
#include "internal.h"
namespace b200 {
bool build_fast_plan(Plan *, ggml_cgraph *) { return false; }
}  // namespace b200
