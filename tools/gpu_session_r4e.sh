#!/bin/bash
# round 2, session 4e: which change moved the step-0 losses of the trainer parity test (stem mma / lane priority / pool width)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "=== stem accuracy"; timeout 200 python tools/stem_accuracy.py 2>&1 | grep -v Warn | tee $O/r4e_stem_accuracy.txt
for v in "CAUSALGEN_B200_STEM_MMA=1" "CAUSALGEN_B200_STEM_MMA=0" "CAUSALGEN_B200_PRIO=0 CAUSALGEN_B200_SIDE_STREAMS=2" "CAUSALGEN_B200_SIDE_STREAMS=0" "CAUSALGEN_B200_STEM_MMA=0 CAUSALGEN_B200_SIDE_STREAMS=0"; do
  echo "=== trainer test, $v"; rm -f $O/parity_report.txt
  env $v timeout 300 python -m pytest tests/test_trainer_gpu.py -q -x -k "tiny_ukbb-3-False" 2>&1 | tail -2
  grep "step0\|step2 nll" $O/parity_report.txt | cut -c1-60,90-170
done 2>&1 | tee $O/r4e_trainer_variants.txt
