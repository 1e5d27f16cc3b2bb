// One translation unit per (mode, minor-slot groups) of the fast kernels; the Makefile compiles this file with
// -DRB_MODE=<SolveMode> -DRB_NGPT=<g-points> -DRB_NG=<1|2> [-DRB_NMU=<1|4>] -DRB_ENTRY=<entry point declared in
// solver_launch.cuh>.
#include "solver_launch.cuh"

#if !defined(RB_MODE) || !defined(RB_NGPT) || !defined(RB_NG) || !defined(RB_ENTRY)
#error "compile with -DRB_MODE= -DRB_NGPT= -DRB_NG= -DRB_ENTRY="
#endif

#ifndef RB_NMU
#define RB_NMU 1
#endif

namespace rb {
int RB_ENTRY(SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    return launch_fast_ng<RB_MODE, RB_NGPT, RB_NG, RB_NMU>(P, max_smem_optin, s);
}
}  // namespace rb
