// One translation unit per (mode, minor-slot groups) of the warp-specialised kernels (solver_ws.cuh); compiled with
// -DRB_MODE=<0|2> -DRB_NGPT=<g-points> -DRB_NG=<1|2> -DRB_ENTRY=<entry point declared in solver_launch.cuh>.
#include "solver_launch.cuh"

#if !defined(RB_MODE) || !defined(RB_NGPT) || !defined(RB_NG) || !defined(RB_ENTRY)
#error "compile with -DRB_MODE= -DRB_NGPT= -DRB_NG= -DRB_ENTRY="
#endif

namespace rb {
int RB_ENTRY(SolveParams<float>& P, int max_smem_optin, cudaStream_t s) {
    return launch_ws_ng<RB_MODE, RB_NGPT, RB_NG>(P, max_smem_optin, s);
}
}  // namespace rb
