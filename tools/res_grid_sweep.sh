#!/bin/bash
# sweep of the resident tridiagonalisation's rows per thread / smallest CTA (XTBB200_RES_RR, XTBB200_RES_TMIN)
for n in 300 552 900 1400 1800; do
  for cfg in "2 256" "4 128" "4 256" "8 64" "8 128" "8 256"; do
    set -- $cfg
    echo "n $n rr $1 tmin $2: $(XTBB200_RES_RR=$1 XTBB200_RES_TMIN=$2 EIG_ONLY=random python tools/eig_check.py $n $n | grep '^n' | tail -1 | sed 's/.*resid/resid/')"
  done
done
