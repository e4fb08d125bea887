#!/bin/bash
# TEST INFRASTRUCTURE: compiles csrc/b2env.cu for the HOST with g++ (every CUDA thread a fiber, see
# cuda_emu.h) -> tools/emu/libb2env_emu.so.  Used by tests/test_emu_*.py to debug kernel logic without a GPU.
# The product never loads this library.
set -e
cd "$(dirname "$0")"
CXX=${CXX:-g++}
$CXX -O2 -g -std=c++17 -fPIC -shared -march=x86-64-v3 -ffp-contract=fast -Wno-unused-result \
  -DB2E_EMU_IMPL -include cuda_emu.h -x c++ ../../pybullet-robot-envs_b200/csrc/b2env.cu \
  -o ${OUT:-libb2env_emu.so} -lpthread "$@"
