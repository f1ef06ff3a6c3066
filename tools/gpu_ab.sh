#!/bin/bash
# A/B of one environment switch on one box: bench.py twice each way.   usage: VAR=LSSVC_NO_HEAD TAG=r2o tools/gpu_ab.sh
TAG=${TAG:-r2}; VAR=${VAR:-LSSVC_NO_PDL}
for mode in on off on off; do
  if [ $mode = off ]; then export $VAR=1; else unset $VAR; fi
  timeout 300 python bench.py --steps 36 --no-cpu-baseline > gpurun_out/${TAG}_${VAR}_${mode}.json 2> gpurun_out/${TAG}_${VAR}_${mode}.err
  python -c "import json;d=json.load(open('gpurun_out/${TAG}_${VAR}_${mode}.json'));f=d['roofline']['in_frame'];print('$VAR unset' if '$mode'=='on' else '$VAR=1',d['value'],d['ms_per_step'],d['clocks']['sm_mhz'],f['p_frame_launch_ms'],f.get('head'),f['ms'],f['achieved'])"
done
