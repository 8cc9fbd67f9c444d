// sweep_inst.cu -- explicit instantiations of launch_sweep (and with it the sweep kernels) for ONE (axis, RECON) pair:
// compiled once per pair with -DJXF_INST_A=<0..2> -DJXF_INST_RECON=<0..5> (jaxfluids_b200/build.py), in parallel.
#include "plan.cuh"

#ifndef JXF_INST_A
#error "compile with -DJXF_INST_A=<axis> -DJXF_INST_RECON=<recon>"
#endif

#define JXF_INST_SWEEP(A, R, S, E) template int launch_sweep<A, R, S, E>(const jxf_solver*, SweepArgs, cudaStream_t);
#ifdef JXF_TUNE_ONLY
JXF_SWEEPS_TUNE(JXF_INST_SWEEP, JXF_INST_A)
#else
JXF_SWEEPS_OF(JXF_INST_SWEEP, JXF_INST_A, JXF_INST_RECON)
#if JXF_INST_RECON < 4
JXF_SWEEPS_PLAIN_OF(JXF_INST_SWEEP, JXF_INST_A, JXF_INST_RECON)
#endif
#endif
