// k_lik_tc.cu -- TMA + tcgen05/TMEM likelihood kernel (placeholder until the kernel lands)
#include "bvg_internal.cuh"
bool tc_available() { return false; }
int tc_setup(bvg_fit *f) { bvg_set_error("tcgen05 engine not built yet"); return BVG_ERR_ARG; }
void tc_destroy(bvg_fit *f) {}
int launch_lik_tc(bvg_fit *f, bool backward) { bvg_set_error("tcgen05 engine not built yet"); return BVG_ERR_ARG; }
