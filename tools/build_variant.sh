#!/bin/bash
# usage: tools/build_variant.sh NAME [extra nvcc flags ...]   -> exp/var_NAME.so  (development aid: A/B kernel experiments)
name=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ftz=true -prec-div=false -prec-sqrt=false -Xptxas --register-usage-level=10 \
  -Xcompiler -fPIC -shared "$@" -o exp/var_$name.so imitation-learning-rl_b200/csrc/ilrl_capi.cu imitation-learning-rl_b200/csrc/ilrl_policy.cu
