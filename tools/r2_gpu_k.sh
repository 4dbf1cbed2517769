#!/bin/bash
set -x
O=gpurun_out
export XTBB200_TUNING=1
( EIG_ONLY=random timeout 600 python tools/eig_check.py 6000 6000 ) 2>&1 | tail -3 > $O/k_eig.log
cat $O/k_eig.log
( XTBB200_SYTRD_PROF=1 EIG_ONLY=random timeout 300 python tools/eig_check.py 6000 ) 2>&1 | grep "sytrd phase" | head -10
( XTBB200_NO_SYM_SYTRD=1 XTBB200_SYTRD_PROF=1 EIG_ONLY=random timeout 300 python tools/eig_check.py 6000 ) 2>&1 | grep "sytrd phase" | head -10
