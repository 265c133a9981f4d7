# 2-bit base codes: the new GPU tests first, then the whole GPU suite, then config 2 end to end with both base forms
python -m pytest tests/test_gpu_qcode.py -m gpu -x -q -k "base_code or ships_base" > gpurun_out/b_new.log 2>&1; echo "rc=$?" >> gpurun_out/b_new.log; tail -3 gpurun_out/b_new.log
python -m pytest tests -m gpu -x -q > gpurun_out/b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/b_tests.log; tail -3 gpurun_out/b_tests.log
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 8"
$B --legs e2e_api > gpurun_out/b_codes.json 2> gpurun_out/b_codes.err; echo "codes rc=$?"
$B --legs main --base-form nibbles > gpurun_out/b_nib.json 2> gpurun_out/b_nib.err; echo "nibbles rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/b_codes.json", "gpurun_out/b_nib.json"):
    d = json.load(open(f))
    print(f, d["ms_per_step"], d["roofline"]["frac"], {k: d["e2e"][k] for k in ("ms_per_step", "h2d_bytes_per_step", "value", "batch_form")}, d["e2e"]["keep_masked_batch"], (d.get("e2e_api") or {}).get("ms_per_step"))
PY
