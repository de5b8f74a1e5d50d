set -u
out=gpurun_out
for occ in 8 10 12; do for mg in 0 1; do
  export SD_CCL_OCC=$occ SD_CCL_MERGE=$mg
  python -m pytest tests/test_gpu_seg.py -q -x -k "ccl or full_size or partition" > $out/s4_t_${occ}_${mg}.log 2>&1; echo "occ=$occ merge=$mg tests: $(tail -1 $out/s4_t_${occ}_${mg}.log)"
  python tools/ccl_bench.py 128 --stages > $out/s4_c128_${occ}_${mg}.json 2>&1
  python tools/ccl_bench.py 512 --stages > $out/s4_c512_${occ}_${mg}.json 2>&1
  python - <<PY
import json
for n in (128, 512):
    try:
        d = json.loads(open("$out/s4_c%d_${occ}_${mg}.json" % n).read().strip().splitlines()[-1])
        print(" ", n, "text", d["text_stages_us"], "lbl %.1f us frac %.3f | +stats %.1f us frac %.3f" % (1e3*d["text"]["label_ms"], d["text"]["label_frac"], 1e3*d["text"]["label_stats_ms"], d["text"]["label_stats_frac"]), "| dense", d["dense_stages_us"], "frac %.3f / %.3f" % (d["dense"]["label_frac"], d["dense"]["label_stats_frac"]))
    except Exception as e:
        print(" ", n, "failed", e)
PY
done; done
