// cooperative_groups.h -- TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory). Blocks of a launch run one
// after the other here, so a grid-wide barrier cannot be met: cooperative launches are refused
// (cudaLaunchCooperativeKernel returns an error) and grid_group::sync() aborts if it is ever reached.
#pragma once
#include "cuda_runtime.h"

namespace cooperative_groups {
struct grid_group {
	void sync() const { ::cuda_on_cpu::ptx::unsupported("cooperative_groups::grid_group::sync (grid-wide barrier)"); }
};
inline grid_group this_grid() { return grid_group(); }
} // namespace cooperative_groups
