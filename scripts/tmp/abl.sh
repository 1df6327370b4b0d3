bash scripts/tmp/quick.sh 2>&1 | grep -v "^+"
cp build_ab/libX.so bayesnf_b200/libbnf_sm100.so
TAG=X WL=wind_map_e16 MASKS=0,1,2,3,0 timeout 400 python scripts/epi_experiment.py 2>&1 | tail -n 6
