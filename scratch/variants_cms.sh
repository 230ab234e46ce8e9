#!/bin/bash
# usage: scratch/variants_cms.sh lib_a.so ... : CMS-scale iteration profile per build variant
for lib in "$@"; do
  echo "=== $lib"
  CELERITAS_B200_LIB=$PWD/celeritas_b200/$lib python scratch/iter_profile.py cms-scale 2>&1 | grep -v "warning" | grep "iterations\|active \[      0,     16)\|active \[  16384\|active \[ 262144\|active \[ 524288"
done
