#!/bin/bash
# Developer helper: builds a single-instance variant library for kernel experiments.
#   ./dev_variant.sh <out.so> <model> <n_lines> [extra nvcc -D flags...]
set -e
cd "$(dirname "$0")"
out=$1; m=$2; l=$3; shift 3
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
T=$(mktemp -d)
nvcc $FL "$@" -DLF_DEV_MODEL=$m -DLF_DEV_L=$l -c lucifit_capi.cu -o $T/capi.o &
nvcc $FL "$@" -c lucifit_reduce.cu -o $T/reduce.o &
nvcc $FL "$@" -c lucifit_prior.cu -o $T/prior.o &
nvcc $FL "$@" -Xptxas -v -DLF_INST_MODEL=$m -DLF_INST_L=$l -c lucifit_fit_inst.cu -o $T/inst.o 2> $T/ptxas.log &
wait
grep -A2 "lf_lm_kernelILi${m}ELi${l}ELi1E" $T/ptxas.log | grep -v Compiling || true
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out $T/capi.o $T/reduce.o $T/prior.o $T/inst.o -lcudart
rm -rf $T
