#!/bin/bash
# CMS-scale iteration profile for several fused-launch thresholds
for f in "$@"; do
  echo "=== fuse_threshold $f"
  FUSE=$f python scratch/iter_profile.py cms-scale 2>&1 | grep -v "warning" | grep "iterations\|active"
done
