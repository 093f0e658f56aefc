#!/bin/bash
for ab in 0 64 80 95; do echo -n "ablate $ab: "; FAMI_DCN_ABLATE=$ab BLOCKED=1 timeout 100 python tools/time_dcn.py 2>&1 | grep "sigma 0.5" | cut -d: -f2 | cut -d'>' -f1; done
