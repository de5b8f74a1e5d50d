run2() { # label, env...
  label=$1; shift
  (CUDA_VISIBLE_DEVICES=0 env "$@" python bench.py --steps 2 --no-cpu-baseline $EXTRA > gpurun_out/two_${label}_0.json 2>/dev/null &
   CUDA_VISIBLE_DEVICES=1 env "$@" python bench.py --steps 2 --no-cpu-baseline $EXTRA > gpurun_out/two_${label}_1.json 2>/dev/null & wait)
  python - <<PY
import json
for i in (0,1):
    d=json.loads(open("gpurun_out/two_${label}_%d.json"%i).read().strip().splitlines()[-1])
    print("${label}", i, round(d["value"]), round(d["e2e"]["value"]), d["clocks"])
PY
}
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)"; nvidia-smi topo -m | head -8
EXTRA="" run2 plain A=1
EXTRA="" run2 omp1 OMP_NUM_THREADS=1
EXTRA="--no-clock-sampler" run2 nosampler A=1
