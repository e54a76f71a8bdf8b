#!/bin/bash
# Round 2, session 3: one-sweep radix passes with swizzled staging -- parity, rate per form, ncu --set full of a middle pass (forms 5 and 2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eer.py -m gpu -q --tb=short -x -k "one_sweep or goldens or sizes" > gpurun_out/test_eer.log 2>&1
echo "eer tests exit $? $(tail -n 1 gpurun_out/test_eer.log)"
grep -h "FAILED\|Error\|assert" gpurun_out/test_eer.log | head -20
timeout 300 python tools/eer_forms.py 100000000 1 2 4 5 > gpurun_out/eer_forms.txt 2>&1
echo "forms exit $?"; cat gpurun_out/eer_forms.txt | tail -12
for form in 5 2; do
  EER_FORM=$form EER_N=100000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"radix_onesweep_kernel" -s 5 -c 1 -f -o gpurun_out/prof_onesweep_form$form python tools/prof_eer_small.py > gpurun_out/ncu_onesweep_form$form.log 2>&1
  echo "ncu form $form exit $?"
done
