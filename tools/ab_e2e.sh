#!/bin/bash
# time rmd_scan end to end (compact transport) with several builds of the library: tools/ab_e2e.sh lib1.so lib2.so ...
for lib in "$@"; do
  echo "$lib: $(RMD_LIB=$lib python tools/profile_scan.py --mb 10 --steps 6 --e2e | grep e2e)"
done
