#!/bin/bash
# build the library with several unroll factors of the chain kernel's stage loops and time each on the GPU
for u in "$@"; do
  touch walking-controllers_b200/csrc/mpc_chain.cu
  make -C walking-controllers_b200/csrc EXTRA=-DWC_CHAIN_UNROLL=$u > /dev/null 2>&1 || { echo build failed $u; continue; }
  cp walking-controllers_b200/libwc_b200.so /tmp/libwc_u$u.so
done
