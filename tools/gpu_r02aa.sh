#!/bin/bash
# Round 2, session 3: one-sweep radix passes -- parity (every form against the super-tile form and the oracle), rate per form, launch lists
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eer.py -m gpu -q --tb=short -x -k "one_sweep or goldens or sizes" > gpurun_out/test_eer.log 2>&1
echo "eer tests exit $? $(tail -n 1 gpurun_out/test_eer.log)"
grep -h "FAILED\|Error\|assert" gpurun_out/test_eer.log | head -20
timeout 300 python tools/eer_forms.py 100000000 1 2 3 4 5 6 > gpurun_out/eer_forms.txt 2>&1
echo "forms exit $?"; cat gpurun_out/eer_forms.txt | tail -12
for form in 1 4; do
  EER_FORM=$form EER_N=100000000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/eer_launches_form$form.csv python tools/prof_eer_small.py > gpurun_out/ncu_form$form.log 2>&1
  echo "ncu form $form exit $?"
done
python - <<'PY'
import csv
for form in (1, 4):
    try:
        rows = list(csv.reader(l for l in open(f"gpurun_out/eer_launches_form{form}.csv") if l.startswith('"')))
    except Exception as e:
        print(form, e); continue
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    print("form", form)
    for r in rows[1:]:
        print("   %-60s %s %s" % (r[ki][:60], r[vi], r[ui]))
PY
