#!/bin/bash
# time the 10 Mb scan with several builds of the library: tools/ab_time.sh lib1.so lib2.so ...  (development aid)
for lib in "$@"; do
  echo "$lib: $(RMD_LIB=$lib python tools/profile_scan.py --mb 10 --steps 5 | cut -d'|' -f2,3)"
done
