// sph_sweeps_other.cuh -- WCSPH / PCISPH / IISPH sweeps (included by sph_sweeps.cu, both modes).
#pragma once

namespace SPH_NS {

void wc_phase(SphHandle *h, int phase, cudaStream_t st) { (void)h; (void)phase; (void)st; }
void pc_phase(SphHandle *h, int phase, cudaStream_t st) { (void)h; (void)phase; (void)st; }
void pc_precompute(SphHandle *h, cudaStream_t st) { (void)h; (void)st; }
void ii_phase(SphHandle *h, int phase, cudaStream_t st) { (void)h; (void)phase; (void)st; }

} // namespace SPH_NS
