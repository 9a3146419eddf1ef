#!/bin/bash
# Builds phyloaln_b200/libphyloaln_b200.so for sm_100a (B200); see Makefile. Extra nvcc flags: EXTRA="...".
set -e
cd "$(dirname "$0")"
make -j8 -s "$@"
