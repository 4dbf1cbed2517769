#!/bin/bash
# Tuning sweeps of the shared-memory resident tridiagonalisation (k_sytrd_smem) on the GPU box; results of round 1 are in
# profiles/r01_resident_sytrd_{grid,rows}_sweep.txt.   usage: res_grid_sweep.sh grid | rows | stages
case "${1:-stages}" in
grid)    # CTAs of the cooperative grid (XTBB200_RES_GRID)
  for n in 552 900 1400; do
    for g in 12 16 24 32 48 64 96 148; do
      echo "n $n grid $g: $(XTBB200_RES_GRID=$g EIG_ONLY=random python tools/eig_check.py $n $n | grep '^n' | tail -1 | sed 's/.*sytrd/sytrd/')"
    done
  done ;;
rows)    # rows of a column per thread and smallest CTA (XTBB200_RES_RR, XTBB200_RES_TMIN)
  for n in 300 552 900 1400 1800; do
    for cfg in "2 256" "4 128" "4 256" "8 64" "8 128" "8 256"; do
      set -- $cfg
      echo "n $n rr $1 tmin $2: $(XTBB200_RES_RR=$1 XTBB200_RES_TMIN=$2 EIG_ONLY=random python tools/eig_check.py $n $n | grep '^n' | tail -1 | sed 's/.*resid/resid/')"
    done
  done ;;
*)       # stage times with the defaults, plus the accuracy tests
  EIG_ONLY=random python tools/eig_check.py 300 552 900 1400 1800 3000 | sed 's/.*dw/dw/'
  python -m pytest tests/test_gpu_eigh.py -q 2>&1 | tail -2 ;;
esac
