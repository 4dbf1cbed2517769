#!/bin/bash
( EIG_ONLY=random timeout 600 python tools/eig_check.py 552 552 1000 1659 1659 ) 2>&1 | tail -6
