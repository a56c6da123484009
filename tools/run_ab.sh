# A/B of run-time switches on the per-launch minibatch timeline: bash tools/run_ab.sh NAME "ENV=1 ..." "ENV=0 ..."
name=$1; shift
i=0
for envs in "$@"; do env $envs python tools/mb_timeline.py 2>&1 | tail -13 > gpurun_out/${name}_$i.txt; echo "== $envs"; cat gpurun_out/${name}_$i.txt; i=$((i+1)); done
