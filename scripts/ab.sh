#!/bin/bash
# same-box comparison of library builds (build_ab/lib<V>.so): alternate them so clock /
# power-cap drift hits all.  usage: VARIANTS="A B" scripts/ab.sh [workload ...]
WLS=${@:-wind_map_e16}
VARIANTS=${VARIANTS:-A B}
for rep in 1 2; do
  for v in $VARIANTS; do
    cp build_ab/lib$v.so bayesnf_b200/libbnf_sm100.so
    for wl in $WLS; do
      TAG="$v $wl" WL=$wl MASKS=${MASKS:-0,0} timeout 200 python scripts/epi_experiment.py 2>&1 | tail -n ${TAILN:-1}
    done
  done
done
