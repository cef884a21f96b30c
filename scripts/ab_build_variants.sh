#!/bin/bash
# A/B of build-time variants ON ONE BOX under the sustained (power-capped) rollout loop: for every argument (a string of extra
# nvcc flags, "" = the tree as it is) rebuild the library and run the headline bench twice.  usage: bash scripts/ab_build_variants.sh "" "-DVITMARL_WAIT_SLEEP_NS=64"
run() { python bench.py --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('[$1]', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'MHz', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'mlp', round(k['fused_mlp']['avg_launch_us'],1), 'attn', round(k['fused_attn_block']['avg_launch_us'],1))"; }
for v in "$@" "${1:-}"; do
  VITMARL_NVCC_EXTRA="$v" python -c "from vitmarl_b200 import _build; _build.build(force=True)" > /dev/null 2>&1 || { echo "build failed for [$v]"; continue; }
  run "$v"; run "$v"
done
