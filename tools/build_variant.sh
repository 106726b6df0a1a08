#!/bin/bash
# Development build of one kernel variant for A/B timing: tools/build_variant.sh NAME [extra nvcc flags ...]
# -> variants/NAME.so (4 lanes per env, no wind, no action tracking only: compiles in seconds).  tools/ab_variants.sh
# copies each variant over the in-tree library on the GPU box and benches it.
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart static \
     -I include -DATC_DEV_FAST "$@" -o variants/$NAME.so atc_reinforcement_learning_b200/csrc/atc_kernels.cu atc_reinforcement_learning_b200/csrc/atc_vecnorm.cu atc_reinforcement_learning_b200/csrc/atc_text.cu
