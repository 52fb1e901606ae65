// librmn_b200/csrc/ez_interp_fast.cuh -- weight-form ("arith = 1") specialisations of the interpolation kernel.
// Included by ez_interp.cu inside its translation unit.
#pragma once

namespace {
// returns false when no specialisation covers this launch (the caller falls back to k_interp_ref)
bool launch_fast(const InterpArgs &a, cudaStream_t s) {
    (void)a; (void)s;
    return false;
}
}  // namespace
