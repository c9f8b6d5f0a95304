#!/bin/bash
# microbenchmarks + quick bench (one GPU box visit)
cd tools/bin
echo "== rf_reuse_rates"; ./rf_reuse_rates
echo "== ceiling old"; ./pc_pair_ceiling_old
echo "== ceiling new"; ./pc_pair_ceiling
cd ../..
