#!/bin/bash
cd $GRAFT_REPO_ROOT
T="timeout -k 5"
for pf in 0 1 2; do
  echo "== WSR_HALO_PREFETCH=$pf"
  for shape in "64 64 64 128 256" "64 192 64 128 256" "64 128 128 64 128" "64 384 128 64 128"; do
    WSR_HALO_PREFETCH=$pf $T 120 python tools/prof_fuse.py $shape 5
  done
  WSR_HALO_PREFETCH=$pf WSR_TC_DBG=3 $T 120 python tools/prof_fuse.py 64 64 64 128 256 5
done
