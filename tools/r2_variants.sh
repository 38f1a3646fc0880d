#!/bin/bash
# cycle accounting of scan-kernel build variants (libsymphony_b200_<variant>.so, built with -DTC_PROFILE)
for V in ${VARIANTS:-prof nowalk grp}; do
  for F in ${FLUSHES:-32}; do
    echo "== variant $V flush $F"
    SYM_TC_FLUSH=$F SYM_B200_LIB=$PWD/symphonypy_b200/_C/libsymphony_b200_$V.so timeout 300 python tools/tc_profile.py 2>&1 | tail -4
  done
done
