#!/bin/bash
# Round 2, session 3: one-sweep radix passes -- parity, rate per form, launch list of the default form
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eer.py -m gpu -q --tb=short -x > gpurun_out/test_eer.log 2>&1
echo "eer tests exit $? $(tail -n 1 gpurun_out/test_eer.log)"
grep -h "FAILED\|Error\|assert" gpurun_out/test_eer.log | head -20
timeout 300 python tools/eer_forms.py 100000000 ${FORMS:-1 2 4 5} > gpurun_out/eer_forms.txt 2>&1
echo "forms exit $?"; cat gpurun_out/eer_forms.txt | tail -12
for form in ${NCU_FORMS:-5}; do
  EER_FORM=$form EER_N=100000000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 15 --csv --log-file gpurun_out/eer_launches_form$form.csv python tools/prof_eer_small.py > gpurun_out/ncu_form$form.log 2>&1
  echo "ncu form $form exit $?"
  python - $form <<'PY'
import csv, sys
form = sys.argv[1]
rows = list(csv.reader(l for l in open(f"gpurun_out/eer_launches_form{form}.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
for r in rows[1:]:
    print("   %-70s %8.1f us" % (r[ki][:70], float(r[vi].replace(",", "")) / 1e3))
PY
done
