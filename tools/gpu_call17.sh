set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -3
B2F_PATH=split timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "float_spectra or pol_modes or multi_if or requantised" 2>&1 | tail -3
for v in split legacy; do
  B2F_PATH=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --seconds 20 > gpurun_out/bench17_$v.json 2> gpurun_out/bench17_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for n in ("split","legacy"):
    try:
        d=json.loads(open(f"gpurun_out/bench17_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],2), "rt", round(d["rt_factor"],1), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["parity_check"]["ok"], "roofline", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(n, "ERR", e); print(open(f"gpurun_out/bench17_{n}.err").read()[-1500:])
PY
B2F_PATH=split python tools/prof_run.py 2 > gpurun_out/plain_split17.log 2>&1 &&
B2F_PATH=split ncu --set full --clock-control none --import-source on -k regex:"kt_row_tiles" -c 1 -o gpurun_out/r02_kt python tools/prof_run.py 2 > gpurun_out/ncu_kt17.log 2>&1
echo "ncu rc=$?"
