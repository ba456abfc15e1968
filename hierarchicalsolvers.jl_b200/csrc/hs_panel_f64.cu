#include "hs_panel.cuh"

void hs_panel_setup_f64() { panel_setup<double>(); }
int hs_panel_width_f64(const hs_fac* f, int max_n, int nfronts) { return choose_width<double>(f, max_n, nfronts); }
void hs_panel_launch_f64(hs_fac* f, int W, int f0, int nact, int j0, int m, cudaStream_t st) { panel_dispatch<double>(f, W, f0, nact, j0, m, st); }
void hs_trsm_rows_f64(hs_fac* f, int W, int f0, int nact, int j0, int max_rows, cudaStream_t st) { trsm_rows_dispatch<double>(f, W, f0, nact, j0, max_rows, st); }
