#include "hs_panel.cuh"

void hs_panel_setup_c64() { panel_setup<cplx>(); }
int hs_panel_width_c64(const hs_fac* f, int max_n, int nfronts) { return choose_width<cplx>(f, max_n, nfronts); }
void hs_panel_launch_c64(hs_fac* f, int W, int f0, int nact, int j0, int m, cudaStream_t st) { panel_dispatch<cplx>(f, W, f0, nact, j0, m, st); }
void hs_trsm_rows_c64(hs_fac* f, int W, int f0, int nact, int j0, int max_rows, cudaStream_t st) { trsm_rows_dispatch<cplx>(f, W, f0, nact, j0, max_rows, st); }
