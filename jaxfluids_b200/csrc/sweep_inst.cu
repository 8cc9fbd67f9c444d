// sweep_inst.cu -- explicit instantiations of launch_sweep (and with it the sweep kernels) for ONE (axis, RECON) pair:
// compiled once per pair with -DJXF_INST_A=<0..2> -DJXF_INST_RECON=<0..5> (jaxfluids_b200/build.py), in parallel.
#include "plan.cuh"

#ifndef JXF_INST_A
#error "compile with -DJXF_INST_A=<axis> -DJXF_INST_RECON=<recon>"
#endif

template int launch_sweep<JXF_INST_A, JXF_INST_RECON, RIEMANN_HLLC, 0>(const jxf_solver*, SweepArgs, cudaStream_t);
template int launch_sweep<JXF_INST_A, JXF_INST_RECON, RIEMANN_HLLC, 1>(const jxf_solver*, SweepArgs, cudaStream_t);
#ifndef JXF_TUNE_ONLY
template int launch_sweep<JXF_INST_A, JXF_INST_RECON, RIEMANN_RUSANOV, 0>(const jxf_solver*, SweepArgs, cudaStream_t);
template int launch_sweep<JXF_INST_A, JXF_INST_RECON, RIEMANN_RUSANOV, 1>(const jxf_solver*, SweepArgs, cudaStream_t);
#endif
